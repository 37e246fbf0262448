// Sample-rate conversion on load (process.py:28: `librosa.load(wav_path, sr=SR)` resamples any file that is not 16 kHz
// with libsoxr's "HQ" converter).  libsoxr cannot be restated bit for bit, so -- like the half-band decimator of the CQT
// -- oracle and device share ONE published filter definition instead (oracle/resample.py, tables.cpp::resample_filter):
// a Kaiser-windowed sinc (150 dB, pass band flat to 0.9125 of the lower Nyquist like soxr HQ, transition up to that
// Nyquist), evaluated as a polyphase filter.  out[m] lies at input time m q / p (p / q = sr_out / sr_in in lowest
// terms): phase f = (m q) mod p, first input sample k0 = (m q) div p - half + 1, 2 half taps, FP64 accumulation in
// ascending tap order, one rounding to float32.
#include "kernels.cuh"

namespace bpc {

// One thread per output sample; the 2 * half coefficients of a phase are contiguous (consecutive outputs cycle through
// the phases, so a warp reads up to 32 rows of the table: it is L1/L2 resident, at most p * 2 half doubles ~ 0.8 MB).
__global__ void __launch_bounds__(256) k_resample(const float* __restrict__ x, long long n_in,
                                                  const double* __restrict__ tab, int p, int q, int half,
                                                  float* __restrict__ out, long long n_out) {
    const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    const long long t = m * (long long)q;
    const long long k0 = t / p - half + 1;
    const int f = (int)(t - (t / p) * p);
    const double* c = tab + (size_t)f * 2 * half;
    double acc = 0.0;
    for (int j = 0; j < 2 * half; ++j) {
        const long long k = k0 + j;
        const double v = (k >= 0 && k < n_in) ? (double)__ldg(x + k) : 0.0;
        acc = fma(__ldg(c + j), v, acc);
    }
    out[m] = (float)acc;
}

void launch_resample(const float* x, long long n_in, const double* tab, int p, int q, int half, float* out,
                     long long n_out, cudaStream_t st) {
    if (n_out <= 0) return;
    k_resample<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(x, n_in, tab, p, q, half, out, n_out);
    note_launch();
}

}  // namespace bpc
