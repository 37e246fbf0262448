// Host-side check of csrc/fft_reg.cuh (the build container has no GPU): runs the lane-explicit stage functions for
// every lane of a team in sequence and compares with a naive float128-ish (long double) DFT.
//   nvcc -x cu -std=c++17 -I breathing-phase-classifier_b200/csrc tests/host/fft_host_test.cpp -o /tmp/fft_host_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fft_reg.cuh"

using namespace bpc;

template <int R>
static double check_team_fft() {
    constexpr int N = R * R;
    const long double PI = 3.14159265358979323846264338327950288L;
    std::vector<double2> x(N), X(N);
    for (int n = 0; n < N; ++n) x[n] = make_double2(std::sin(0.37 * n) + 0.01 * n, std::cos(1.3 * n * n * 0.001) - 0.5);
    std::vector<double2> xch(R * (R + 1));
    std::vector<std::vector<double2>> regs(R, std::vector<double2>(R));
    for (int h = 0; h < R; ++h) {
        double2 tw[R];
        for (int k1 = 0; k1 < R; ++k1) {
            const long double ang = -2.0L * PI * (long double)(h * k1) / N;
            tw[k1] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
        for (int j = 0; j < R; ++j) regs[h][j] = x[h + R * j];
        team_fft_stage_a<R>(regs[h].data(), tw, 1, xch.data(), h);
    }
    for (int h = 0; h < R; ++h) {
        team_fft_stage_b<R>(regs[h].data(), xch.data(), h);
        for (int k2 = 0; k2 < R; ++k2) X[h + R * k2] = regs[h][bitrev<R>(k2)];
    }
    double worst = 0.0;
    for (int k = 0; k < N; ++k) {
        long double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            const long double ang = -2.0L * PI * (long double)((long long)k * n % N) / N;
            re += x[n].x * cosl(ang) - x[n].y * sinl(ang);
            im += x[n].x * sinl(ang) + x[n].y * cosl(ang);
        }
        worst = std::fmax(worst, std::fabs((double)(re - X[k].x)));
        worst = std::fmax(worst, std::fabs((double)(im - X[k].y)));
    }
    // real split: 2N reals
    std::vector<double> y(2 * N);
    for (int n = 0; n < N; ++n) { y[2 * n] = x[n].x; y[2 * n + 1] = x[n].y; }
    for (int h = 0; h < R; ++h) {
        const long double ang = -2.0L * PI * h / (2.0L * N);
        const double2 wl = make_double2((double)sinl(ang), -(double)cosl(ang));      // -i * exp(i ang)
        const int hp = (R - h) % R;
        for (int k2 = 0; k2 < R; ++k2) {
            const int k2p = h ? R - 1 - k2 : (R - k2) % R;
            const double2 zk = regs[h][bitrev<R>(k2)], zn = regs[hp][bitrev<R>(k2p)];
            double2 t;
            // compile-time K2 dispatch
            auto call = [&](auto K) { t = rsplit_term<R, decltype(K)::value>(zk, zn, wl); };
            bool done = false;
            auto try_k = [&](auto K) { if (!done && decltype(K)::value == k2) { call(K); done = true; } };
            [&]<int... Is>(std::integer_sequence<int, Is...>) { (try_k(std::integral_constant<int, Is>{}), ...); }
            (std::make_integer_sequence<int, R>{});
            const int k = h + R * k2;
            long double re = 0, im = 0;
            for (int n = 0; n < 2 * N; ++n) {
                const long double a2 = -2.0L * PI * (long double)((long long)k * n % (2 * N)) / (2.0L * N);
                re += y[n] * cosl(a2);
                im += y[n] * sinl(a2);
            }
            worst = std::fmax(worst, std::fabs((double)(re - 0.5 * t.x)));
            worst = std::fmax(worst, std::fabs((double)(im - 0.5 * t.y)));
        }
    }
    return worst;
}

int main() {
    const double e16 = check_team_fft<16>(), e32 = check_team_fft<32>();
    std::printf("team_fft<16> max abs err %.3e\nteam_fft<32> max abs err %.3e\n", e16, e32);
    return (e16 < 1e-10 && e32 < 1e-9) ? 0 : 1;
}
