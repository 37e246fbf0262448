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


// r02-k: N = 1024 on a team of 64 threads (three register stages 16 x 16 x 4, two exchanges) + the pairing of the
// real-input split: the thread that holds Z[N - k] must be the partner LANE of the same warp.
static double check_team64() {
    constexpr int N = 1024;
    const long double PI = 3.14159265358979323846264338327950288L;
    std::vector<double2> x(N), X(N), Z(N);
    for (int n = 0; n < N; ++n) x[n] = make_double2(std::sin(0.37 * n) + 0.01 * n, std::cos(1.3 * n * n * 0.001) - 0.5);
    for (int k = 0; k < N; ++k) {
        long double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            const long double ang = -2.0L * PI * (long double)((long long)k * n % N) / N;
            re += x[n].x * cosl(ang) - x[n].y * sinl(ang);
            im += x[n].x * sinl(ang) + x[n].y * cosl(ang);
        }
        X[k] = make_double2((double)re, (double)im);
    }
    std::vector<double2> twa(16 * 64), twb(16 * 4);
    for (int k1 = 0; k1 < 16; ++k1)
        for (int j = 0; j < 64; ++j) {
            const long double ang = -2.0L * PI * (long double)(j * k1) / N;
            twa[k1 * 64 + j] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
    for (int k2 = 0; k2 < 16; ++k2)
        for (int j0 = 0; j0 < 4; ++j0) {
            const long double ang = -2.0L * PI * (long double)(j0 * k2) / 64;
            twb[k2 * 4 + j0] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
    std::vector<double2> buf(16 * kT64Pitch), buf2(16 * kT64Pitch);
    std::vector<std::vector<double2>> regs(64, std::vector<double2>(16));
    std::vector<int> written(16 * kT64Pitch, 0);
    for (int tid = 0; tid < 64; ++tid) {                               // stage A + exchange-1 stores
        for (int n1 = 0; n1 < 16; ++n1) regs[tid][n1] = x[tid + 64 * n1];
        team64_stage(regs[tid].data(), twa.data() + tid, 64);
        for (int k = 0; k < 16; ++k) { buf[t64_x1_store(k, tid)] = regs[tid][bitrev<16>(k)]; written[t64_x1_store(k, tid)]++; }
    }
    double worst = 0.0;
    int bad = 0;
    std::vector<int> k1_seen(16, 0);
    for (int tid = 0; tid < 64; ++tid) {                               // exchange-1 loads, stage B, exchange-2 stores
        const int k1 = t64_k1(tid), q = tid & 3;
        k1_seen[k1]++;
        for (int m = 0; m < 16; ++m) { regs[tid][m] = buf[t64_x1_load(k1, q, m)]; bad += written[t64_x1_load(k1, q, m)] != 1; }
        team64_stage(regs[tid].data(), twb.data() + q, 4);
        for (int k = 0; k < 16; ++k) buf2[t64_x2_store(k1, q, k)] = regs[tid][bitrev<16>(k)];
    }
    for (int k1 = 0; k1 < 16; ++k1) bad += k1_seen[k1] != 4;
    for (int tid = 0; tid < 64; ++tid) {                               // exchange-2 loads, stage C
        const int k1 = t64_k1(tid), q = tid & 3;
        for (int r = 0; r < 4; ++r)
            for (int j0 = 0; j0 < 4; ++j0) regs[tid][4 * r + j0] = buf2[t64_x2_load(k1, q, r, j0)];
        team64_stage_c(regs[tid].data());
        const int c = k1 + 16 * q;
        for (int u = 0; u < 16; ++u) {
            const double2 z = regs[tid][t64_zidx(u)];
            Z[c + 64 * u] = z;
            worst = std::fmax(worst, std::fabs(z.x - X[c + 64 * u].x));
            worst = std::fmax(worst, std::fabs(z.y - X[c + 64 * u].y));
        }
    }
    // bank pattern of the 8-byte exchange accesses: the sixteen lanes of a half-warp must hit sixteen different banks
    for (int hw = 0; hw < 4; ++hw) {
        for (int m = 0; m < 16; ++m) {
            int seen1 = 0, seen2 = 0, seen3 = 0;
            for (int l = 0; l < 16; ++l) {
                const int tid = 16 * hw + l, k1 = t64_k1(tid), q = tid & 3;
                seen1 |= 1 << (t64_x1_load(k1, q, m) & 15);
                seen2 |= 1 << (t64_x2_store(k1, q, m) & 15);
                seen3 |= 1 << (t64_x2_load(k1, q, m >> 2, m & 3) & 15);
            }
            bad += (seen1 != 0xffff) + (seen2 != 0xffff) + (seen3 != 0xffff);
        }
    }
    // pairing of the split: Z[N - k] for k = c + 64 u, u < 8 (and k = 512 on c = 0), comes from the partner lane's
    // register t64_zidx(15 - u) -- or, for c = 0, from this thread's own t64_zidx((16 - u) % 16)
    for (int tid = 0; tid < 64; ++tid) {
        const int k1 = t64_k1(tid), q = tid & 3, c = k1 + 16 * q;
        const int pt = (tid & 32) | t64_partner_lane(tid);
        for (int u = 0; u <= 8; ++u) {
            if (u == 8 && c != 0) continue;
            const int k = c + 64 * u;
            const double2 zn = c == 0 ? regs[tid][t64_zidx((16 - u) % 16)] : regs[pt][t64_zidx(15 - u)];
            const double2 want = Z[(N - k) % N];
            bad += !(zn.x == want.x && zn.y == want.y);
        }
    }
    if (bad) { std::printf("team64: %d index errors\n", bad); return 1.0; }
    return worst;
}

int main() {
    const double e16 = check_team_fft<16>(), e32 = check_team_fft<32>();
    const double e64 = check_team64();
    std::printf("team_fft<16> max abs err %.3e\nteam_fft<32> max abs err %.3e\nteam64 (1024 on two warps) max abs err %.3e\n", e16, e32, e64);
    return (e16 < 1e-10 && e32 < 1e-9 && e64 < 1e-9) ? 0 : 1;
}
