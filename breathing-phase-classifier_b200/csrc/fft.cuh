// FP64 in-place radix-4 decimation-in-frequency FFTs in shared memory.
//
// Why FP64: the reference STFT (librosa.stft, process.py:32,43,51 / methods.py:59-63,84,90,138) is a float64 FFT
// rounded once to complex64, and the parity budget is 1e-3 dB over an 80 dB window -- a float32 FFT measures
// 3e-4..9e-4 dB off on the fixture set ("precision").  B200 issues 64 DFMA/clk/SM, so the FFTs stay
// affordable.
//
// Layout: N = 4^M complex points; after the M passes X[k] sits at logical position rev4<M>(k) (base-4 digit reversal).
// Two variants:
//   fft_r4_dif       CTA- or warp-wide, plain indexing, twiddles tw[j] = exp(-2 pi i j / N) (used by k_autocorr)
//   warp_fft_r4      one warp, XOR-swizzled buffer (conflict-free in every pass, incl. quarter-span 4 and 1) and
//                    per-pass contiguous twiddle tables read through the read-only path (r01 ncu: 47 % of the
//                    shared wavefronts of the plain variant were bank conflicts)
#pragma once
#include <cuda_runtime.h>

namespace bpc {

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

template <int M>
__device__ __forceinline__ int rev4(int k) {
    unsigned r = __brev((unsigned)k) >> (32 - 2 * M);
    return (int)(((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u));
}

// 16-byte elements, 8 per 128-byte bank row: XOR the row-slot with the next three index bits.
__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 7); }

struct SyncWarp { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
struct SyncBlock { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// One radix-4 DIF butterfly at `base` with quarter-span q; twiddle index step `ts` (= N / span).  kSwz: the buffer uses
// the XOR-swizzled layout (swz()), which keeps the four accesses of every pass, including quarter-spans 4 and 1, on
// distinct banks.
template <bool kTwiddle, bool kSwz = false>
__device__ __forceinline__ void r4_butterfly(double2* x, int base, int q, const double2* tw, int pos_ts) {
    const int i0 = kSwz ? swz(base) : base, i1 = kSwz ? swz(base + q) : base + q;
    const int i2 = kSwz ? swz(base + 2 * q) : base + 2 * q, i3 = kSwz ? swz(base + 3 * q) : base + 3 * q;
    const double2 a0 = x[i0], a1 = x[i1], a2 = x[i2], a3 = x[i3];
    const double2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3);
    const double2 d = csub(a1, a3);
    const double2 b3 = make_double2(d.y, -d.x);            // -i * (a1 - a3)
    double2 y0 = cadd(b0, b2), y1 = cadd(b1, b3), y2 = csub(b0, b2), y3 = csub(b1, b3);
    if (kTwiddle) {
        y1 = cmul(y1, tw[pos_ts]);
        y2 = cmul(y2, tw[2 * pos_ts]);
        y3 = cmul(y3, tw[3 * pos_ts]);
    }
    x[i0] = y0;
    x[i1] = y1;
    x[i2] = y2;
    x[i3] = y3;
}

// Forward complex FFT of N = 4^M points by a team of NT threads (tid in [0, NT)); `sync` separates the passes.
template <int M, int NT, class Sync, bool kSwz = false>
__device__ __forceinline__ void fft_r4_dif(double2* x, const double2* tw, int tid, Sync sync) {
    constexpr int N = 1 << (2 * M);
#pragma unroll
    for (int p = 0; p < M; ++p) {
        const int span = N >> (2 * p);
        const int q = span >> 2;
        const int ts = N / span;
#pragma unroll
        for (int j0 = 0; j0 < N / 4; j0 += NT) {
            const int j = j0 + tid;
            if ((N / 4) % NT == 0 || j < N / 4) {
                const int pos = j & (q - 1);
                const int base = ((j - pos) << 2) + pos;
                if (p == M - 1) r4_butterfly<false, kSwz>(x, base, q, tw, 0);
                else r4_butterfly<true, kSwz>(x, base, q, tw, pos * ts);
            }
        }
        sync();
    }
}

// Real-input FFT post-processing: Z = FFT_N(z), z[m] = x[2m] + i x[2m+1]  ->  X[k] of the 2N-point real FFT,
// k in [0, N].  `ptw[k]` = exp(-2*pi*i*k/(2N)).  kSwz selects the swizzled buffer layout.
template <int M, bool kSwz = false>
__device__ __forceinline__ double2 rfft_bin(const double2* z, const double2* __restrict__ ptw, int k) {
    constexpr int N = 1 << (2 * M);
    const int ik = rev4<M>(k & (N - 1)), in = rev4<M>((N - k) & (N - 1));
    const double2 zk = z[kSwz ? swz(ik) : ik];
    const double2 zn = z[kSwz ? swz(in) : in];
    const double2 e = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
    const double2 o = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
    const double2 w = __ldg(ptw + k);
    return make_double2(e.x + (w.x * o.x - w.y * o.y), e.y + (w.x * o.y + w.y * o.x));
}

// Per-pass twiddle table layout for warp_fft_r4<M>: for pass p (quarter-span q = N / 4^(p+1), p < M-1) the block
// [off_p, off_p + 3q) holds W_N^(r * pos * N/(4q)) at index (r-1)*q + pos, r = 1..3.  Total 3 * (N/4 + N/16 + ... + 4).
constexpr int twp_size(int M) { return M <= 1 ? 0 : 3 * (1 << (2 * (M - 1))) + twp_size(M - 1); }

// One warp, swizzled buffer, per-pass twiddles from global memory (read-only path, L1 resident).
// The caller must __syncwarp() after filling x; the function ends with a __syncwarp().
template <int M>
__device__ __forceinline__ void warp_fft_r4(double2* x, const double2* __restrict__ twp, int lane) {
    constexpr int N = 1 << (2 * M);
    int toff = 0;
#pragma unroll
    for (int p = 0; p < M; ++p) {
        const int q = N >> (2 * (p + 1));
#pragma unroll 2
        for (int j = lane; j < N / 4; j += 32) {
            const int pos = j & (q - 1);
            const int base = ((j - pos) << 2) + pos;
            const int i0 = swz(base), i1 = swz(base + q), i2 = swz(base + 2 * q), i3 = swz(base + 3 * q);
            const double2 a0 = x[i0], a1 = x[i1], a2 = x[i2], a3 = x[i3];
            const double2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3);
            const double2 d = csub(a1, a3);
            const double2 b3 = make_double2(d.y, -d.x);
            double2 y0 = cadd(b0, b2), y1 = cadd(b1, b3), y2 = csub(b0, b2), y3 = csub(b1, b3);
            if (p < M - 1) {
                y1 = cmul(y1, __ldg(twp + toff + pos));
                y2 = cmul(y2, __ldg(twp + toff + q + pos));
                y3 = cmul(y3, __ldg(twp + toff + 2 * q + pos));
            }
            x[i0] = y0;
            x[i1] = y1;
            x[i2] = y2;
            x[i3] = y3;
        }
        toff += 3 * q;
        __syncwarp();
    }
}

// numpy: np.abs(complex64) == hypotf(re, im); glibc evaluates it in double and rounds once.
__device__ __forceinline__ float c64_abs(double2 x) {
    const float re = (float)x.x, im = (float)x.y;
    return (float)sqrt((double)re * (double)re + (double)im * (double)im);
}

}  // namespace bpc
