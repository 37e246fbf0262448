// Host run of csrc/fft20.cuh: the radix-20 Hilbert transform exactly as k_hilbert sequences it (three DIF passes, the
// pair split, three DIT passes), every butterfly executed in turn.  usage: fft20_host_test in.f32 out.f32
// in: 16000 float32 samples; out: 16000 float32 = imag(scipy.signal.hilbert(in)).
#include <cmath>
#include <cstdio>
#include <vector>
#include "fft20.cuh"
using namespace bpc;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::vector<float> y(16000), h(16000);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(y.data(), 4, 16000, f) != 16000) return 3;
    fclose(f);
    const double kPi = 3.14159265358979323846;
    std::vector<float2> twa(20 * 400), twb(20 * 20), ptw(8001), x(kH20Pitch, make_float2(0.f, 0.f));
    for (int k = 0; k < 20; ++k)
        for (int p = 0; p < 400; ++p) {
            const double a = -2.0 * kPi * double(k * p) / 8000.0;
            twa[k * 400 + p] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int k = 0; k < 20; ++k)
        for (int p = 0; p < 20; ++p) {
            const double a = -2.0 * kPi * double(k * p) / 400.0;
            twb[k * 20 + p] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int k = 0; k <= 8000; ++k) {
        const double a = -2.0 * kPi * double(k) / 16000.0;
        ptw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    for (int m = 0; m < kH20N; ++m) x[h20_pad(m)] = make_float2(y[2 * m], y[2 * m + 1]);
    for (int j = 0; j < kH20Bfly; ++j) h20_butterfly<8000, false>(x.data(), twa.data(), j);
    for (int j = 0; j < kH20Bfly; ++j) h20_butterfly<400, false>(x.data(), twb.data(), j);
    for (int j = 0; j < kH20Bfly; ++j) h20_butterfly<20, false>(x.data(), nullptr, j);
    for (int k = 0; k <= kH20N / 2; ++k)
        h20_split_pair(x.data(), k, h20_pad(h20_pos(k)), h20_pad(h20_pos((kH20N - k) % kH20N)), ptw[k]);
    for (int j = 0; j < kH20Bfly; ++j) h20_butterfly<20, true>(x.data(), nullptr, j);
    for (int j = 0; j < kH20Bfly; ++j) h20_butterfly<400, true>(x.data(), twb.data(), j);
    for (int j = 0; j < kH20Bfly; ++j) h20_butterfly<8000, true>(x.data(), twa.data(), j);
    for (int m = 0; m < kH20N; ++m) {
        const float2 o = x[h20_pad(m)];
        h[2 * m] = o.x * (1.0f / 8000.f);
        h[2 * m + 1] = -o.y * (1.0f / 8000.f);
    }
    f = fopen(argv[2], "wb");
    fwrite(h.data(), 4, 16000, f);
    fclose(f);
    return 0;
}
