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

}  // namespace bpc
