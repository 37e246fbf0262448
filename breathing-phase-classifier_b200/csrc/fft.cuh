// Small shared pieces of the FFT kernels.  The transforms themselves live in fft_reg.cuh (register-resident FP64 team
// FFTs: STFT-512, STFT-2048, CQT, autocorrelation) and k_time.cu (mixed-radix float32 Hilbert transforms).
//
// Why FP64: the reference STFT (librosa.stft, process.py:32,43,51 / methods.py:59-63,84,90,138) is a float64 FFT
// rounded once to complex64, and the parity budget is 1e-3 dB over an 80 dB window -- a float32 FFT measures
// 3e-4..9e-4 dB off on the fixture set ("precision").  B200 issues 64 DFMA/clk/SM, so the FFTs stay
// affordable.  (r01 v0/v1 ran shared-memory radix-4 passes here; v2 moved every transform into registers.)
#pragma once
#include <cuda_runtime.h>

namespace bpc {

// numpy: np.abs(complex64) == hypotf(re, im); glibc evaluates it in double and rounds once.
__device__ __forceinline__ float c64_abs(double2 x) {
    const float re = (float)x.x, im = (float)x.y;
    return (float)sqrt((double)re * (double)re + (double)im * (double)im);
}

// |re + i im| of a complex64 the way numpy / glibc do it: (float)sqrt((double)re * re + (double)im * im), evaluated in
// float32 pairs: s = hi + lo exactly (error-free products and sum), r = sqrt(hi) corrected by the residual.  With a
// correctly rounded reciprocal square root the result equals the double-precision evaluation on every one of 2 M random
// inputs; the hardware approximation (<= 2 ulp) leaves a second-order error of ~3e-14, i.e. a 1-ulp difference in about
// one result per million (tests/test_host.py::test_c64_abs_f32_algorithm emulates both cases on the CPU).
// (The DSQRT sequence of c64_abs was 6 % of k_frame2048's instructions.)
// exact path of c64_abs_f32, out of line: its DSQRT sequence is ~32 instructions per call site, and k_frame2048 has 32
// call sites -- inlined it pushed that kernel's code past the instruction cache (r01 v30: 180 KB of SASS, 11 of 12
// issue slots lost to "no instruction" stalls)
static __device__ __noinline__ float c64_abs_exact(float re, float im) {
    return (float)sqrt((double)re * (double)re + (double)im * (double)im);
}

// The float32-pair evaluation without its range check; *x receives max(|re|, |im|).  The result is only valid for
// 1e-18 < x < 1e18 (squares neither denormal nor infinite; x == 0 yields NaN): the caller tracks the range of x over a
// whole row and redoes the row with c64_abs_exact in the rare case (an all-zero frame) -- one warp-uniform branch per
// row instead of a branch per bin, which split the split loops of the STFT kernels into 33 basic blocks
// (r02: k_frame2048 2.55 -> 2.41 ms, k_stft512 0.365 -> 0.34 ms).
__device__ __forceinline__ float c64_abs_f32_unchecked(float re, float im, float* xmax_out) {
    const float a = fabsf(re), b = fabsf(im);
    const float x = fmaxf(a, b), y = fminf(a, b);
    const float p = __fmul_rn(x, x), pe = __fmaf_rn(x, x, -p);          // x^2 = p + pe
    const float q = __fmul_rn(y, y), qe = __fmaf_rn(y, y, -q);          // y^2 = q + qe
    const float hi = __fadd_rn(p, q);
    const float lo = __fadd_rn(__fadd_rn(__fsub_rn(p, hi), q), __fadd_rn(pe, qe));   // p >= q: fast two-sum
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(hi));           // ~2 ulp: the residual step absorbs it
    const float r0 = __fmul_rn(hi, rs);
    const float res = __fadd_rn(__fmaf_rn(-r0, r0, hi), lo);            // (hi + lo) - r0^2
    *xmax_out = x;
    return __fmaf_rn(res, __fmul_rn(0.5f, rs), r0);
}
__device__ __forceinline__ bool c64_abs_in_range(float x) { return x > 1e-18f && x < 1e18f; }

__device__ __forceinline__ float c64_abs_f32(float re, float im) {
    float x;
    float out = c64_abs_f32_unchecked(re, im, &x);
    if (!c64_abs_in_range(x)) out = c64_abs_exact(re, im);              // zero / denormal squares / overflow: exact path
    return out;
}

}  // namespace bpc
