// Float32 complex FFT of length 8000 = 20 x 20 x 20 for scipy.signal.hilbert(y) on a 16000-sample float32 segment
// (methods.py:72; scipy runs a float32 pocketfft on float32 input).  Three radix-20 passes instead of the six radix-5 /
// radix-4 passes of the first version (k_time.cu, r01 v0-v27): every butterfly keeps its 20 points in registers and does
// the 20-point DFT as a twiddle-free prime-factor (Good-Thomas) 4 x 5 transform, so the data cross shared memory 6 times
// per Hilbert transform instead of 12, and the per-pass twiddles come from tables laid out [k][pos] (coalesced).
//
// Storage: element i of the 8000-point array lives at h20_pad(i) = i + i / 20 (8400 float2).  With that pitch the three
// access patterns of the passes (lane stride 1 with element stride 420, lane stride 1 inside 20-blocks with element
// stride 21, lane stride 21) are all free of shared-memory bank conflicts for 8-byte accesses (16 lanes per wavefront).
//
// Forward = decimation in frequency (natural in, digit-reversed out: bin k lands at h20_pos(k)); the inverse runs the
// transposed decimation-in-time network on the spectrum where it lies (digit-reversed in, natural out).
// Everything is __host__ __device__ and butterfly-explicit so tests/host/fft20_host_test.cpp can run the exact index
// logic on the CPU (there is no GPU in the build container).
#pragma once
#include <cuda_runtime.h>

namespace bpc {

#ifndef BPC_HD
#define BPC_HD __host__ __device__ __forceinline__
#endif

constexpr int kH20N = 8000;                 // transform length
constexpr int kH20Pitch = 8400;             // padded storage, float2 elements
constexpr int kH20Bfly = 400;               // butterflies per pass

BPC_HD float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
BPC_HD float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
BPC_HD float2 f2mul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }

BPC_HD int h20_pad(int i) { return i + i / 20; }
// unpadded position of output bin k after the three DIF passes
BPC_HD int h20_pos(int k) {
    const int k1 = k % 20, r = k / 20;
    return k1 * 400 + (r % 20) * 20 + r / 20;
}

// forward 4-point DFT of (a, b, c, d), in place
BPC_HD void dft4f(float2& a, float2& b, float2& c, float2& d) {
    const float2 s0 = f2add(a, c), d0 = f2sub(a, c), s1 = f2add(b, d), t = f2sub(b, d);
    const float2 d1 = make_float2(t.y, -t.x);                       // -i (b - d)
    a = f2add(s0, s1); b = f2add(d0, d1); c = f2sub(s0, s1); d = f2sub(d0, d1);
}
// forward 5-point DFT, in place
BPC_HD void dft5f(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 t1 = f2add(a1, a4), t2 = f2add(a2, a3), t3 = f2sub(a1, a4), t4 = f2sub(a2, a3);
    const float2 m1 = make_float2(fmaf(c2, t2.x, fmaf(c1, t1.x, a0.x)), fmaf(c2, t2.y, fmaf(c1, t1.y, a0.y)));
    const float2 m2 = make_float2(fmaf(c1, t2.x, fmaf(c2, t1.x, a0.x)), fmaf(c1, t2.y, fmaf(c2, t1.y, a0.y)));
    const float2 n1 = make_float2(fmaf(s2, t4.x, s1 * t3.x), fmaf(s2, t4.y, s1 * t3.y));
    const float2 n2 = make_float2(fmaf(-s1, t4.x, s2 * t3.x), fmaf(-s1, t4.y, s2 * t3.y));
    a0 = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
    a1 = make_float2(m1.x + n1.y, m1.y - n1.x);                     // m1 - i n1
    a4 = make_float2(m1.x - n1.y, m1.y + n1.x);
    a2 = make_float2(m2.x + n2.y, m2.y - n2.x);
    a3 = make_float2(m2.x - n2.y, m2.y + n2.x);
}

// Forward 20-point DFT of register-resident points, natural order in and out.  Good-Thomas: n = (5 n1 + 4 n2) mod 20,
// k = (5 k1 + 16 k2) mod 20  =>  W20^(n k) = W4^(n1 k1) W5^(n2 k2): five 4-point and four 5-point DFTs, no twiddles.
BPC_HD void dft20f(float2 (&a)[20]) {
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2)
        dft4f(a[(4 * n2) % 20], a[(5 + 4 * n2) % 20], a[(10 + 4 * n2) % 20], a[(15 + 4 * n2) % 20]);
    // now a[(5 k1 + 4 n2) % 20] = V[k1][n2]
    float2 b[20];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 v0 = a[(5 * k1) % 20], v1 = a[(5 * k1 + 4) % 20], v2 = a[(5 * k1 + 8) % 20], v3 = a[(5 * k1 + 12) % 20],
               v4 = a[(5 * k1 + 16) % 20];
        dft5f(v0, v1, v2, v3, v4);
        b[(5 * k1) % 20] = v0;
        b[(5 * k1 + 16) % 20] = v1;
        b[(5 * k1 + 32) % 20] = v2;
        b[(5 * k1 + 48) % 20] = v3;
        b[(5 * k1 + 64) % 20] = v4;
    }
#pragma unroll
    for (int k = 0; k < 20; ++k) a[k] = b[k];
}

// Butterfly j (0 <= j < 400) of the pass with sub-transform length SPAN in {8000, 400, 20} on the padded array x.
// tw: this pass's table, tw[k * Q + pos] = exp(-2 pi i k pos / SPAN), Q = SPAN / 20 (unused for SPAN == 20).
// DIF (kDit = false): small DFT, then twiddle; DIT: twiddle, then small DFT.
template <int SPAN, bool kDit>
BPC_HD void h20_butterfly(float2* x, const float2* __restrict__ tw, int j) {
    constexpr int Q = SPAN / 20;
    constexpr int ES = Q + Q / 20;                         // element stride in the padded array: 420, 21, 1
    const int blk = j / Q, pos = j - blk * Q;
    const int base = h20_pad(blk * SPAN + pos);
    float2 a[20];
#pragma unroll
    for (int r = 0; r < 20; ++r) a[r] = x[base + r * ES];
    if (kDit && Q > 1) {
#pragma unroll
        for (int r = 1; r < 20; ++r) a[r] = f2mul(a[r], tw[r * Q + pos]);
    }
    dft20f(a);
    if (!kDit && Q > 1) {
#pragma unroll
        for (int r = 1; r < 20; ++r) a[r] = f2mul(a[r], tw[r * Q + pos]);
    }
#pragma unroll
    for (int r = 0; r < 20; ++r) x[base + r * ES] = a[r];
}

// Pair (k, N - k), 0 <= k <= N / 2, of the spectrum Z = FFT_8000(y[2m] + i y[2m+1]) lying digit-reversed in x:
// real-FFT split -> Y[k], Y[N-k] (bins of the 16000-point real transform);  G = -i Y on 0 < k < 8000, 0 at k = 0 and
// k = 8000;  inverse split -> the 8000-point spectrum whose inverse transform is h[2m] + i h[2m+1]; stored conjugated
// (the inverse is run as a forward transform of the conjugate).  w = exp(-2 pi i k / 16000).
// pk / pn: padded storage positions of bins k and (N - k) % N, h20_pad(h20_pos(.)) (the kernel reads them from a table).
BPC_HD void h20_split_pair(float2* x, int k, int pk, int pn, float2 w) {
    const float2 zk = x[pk], zn = x[pn];
    // Y[k] = E + w O ; Y[N-k] = conj(E - w O)
    const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
    const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
    const float2 wo = f2mul(w, o);
    const float2 yk = f2add(e, wo);
    const float2 ys = f2sub(e, wo);
    const float2 yn = make_float2(ys.x, -ys.y);
    float2 gk = make_float2(yk.y, -yk.x), gn = make_float2(yn.y, -yn.x);
    if (k == 0) { gk = make_float2(0.f, 0.f); gn = make_float2(0.f, 0.f); }
    // inverse split: E' = (G[k] + conj(G[N-k])) / 2, O' = (G[k] - conj(G[N-k])) / 2 * conj(w), Z' = E' + i O'
    const float2 e2 = make_float2(0.5f * (gk.x + gn.x), 0.5f * (gk.y - gn.y));
    const float2 d2 = make_float2(0.5f * (gk.x - gn.x), 0.5f * (gk.y + gn.y));
    const float2 o2 = f2mul(d2, make_float2(w.x, -w.y));
    const float2 z = make_float2(e2.x - o2.y, e2.y + o2.x);
    // partner: E'[N-k] = conj(E'[k]); O'[N-k] = (G[N-k] - conj(G[k])) / 2 * conj(w[N-k]), conj(w[N-k]) = -w[k]
    const float2 dn = make_float2(0.5f * (gn.x - gk.x), 0.5f * (gn.y + gk.y));
    const float2 on = f2mul(dn, make_float2(-w.x, -w.y));
    const float2 zn2 = make_float2(e2.x - on.y, -e2.y + on.x);
    x[pk] = make_float2(z.x, -z.y);
    if (pn != pk) x[pn] = make_float2(zn2.x, -zn2.y);             // k == 0 and k == N / 2 pair with themselves
}

}  // namespace bpc
