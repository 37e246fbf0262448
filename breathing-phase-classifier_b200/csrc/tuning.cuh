// Tuning estimation shared by the STFT-512 (chroma_stft, 12 bins/octave) and STFT-2048 (chroma_cens, 36 bins/octave)
// paths: librosa.estimate_tuning -> piptrack -> pitch_tuning (reached from process.py:52 and process.py:53).
#pragma once
#include <cfloat>
#include "common.cuh"

namespace bpc {

// ---------------------------------------------------------------------------- tuning estimation (shared with 2048)
// librosa.estimate_tuning -> piptrack -> pitch_tuning on candidate lists held in shared memory.
// cand_mag / cand_pitch: n candidates (mag + dskew, pitch in Hz, both float32).  sortbuf: >= next_pow2(n) floats.
// Returns the histogram bin (0..99); sets *empty when the frequency set is empty (tuning 0.0 == bin 50).
__device__ inline int tuning_from_candidates(const float* cand_mag, const float* cand_pitch, int n, float* sortbuf,
                                      int* hist, const double* __restrict__ edges, int bins_per_octave, bool* empty) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_best;
    if (n == 0) {
        *empty = true;
        return 50;                   // edges[50] == 0.0
    }
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = tid; i < n2; i += nt) sortbuf[i] = i < n ? cand_mag[i] : FLT_MAX;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += nt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = sortbuf[i], c = sortbuf[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > c) == asc) { sortbuf[i] = c; sortbuf[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // np.median: middle element, or float32 mean of the two middle elements
    const float thr = (n & 1) ? sortbuf[n >> 1] : __fmul_rn(__fadd_rn(sortbuf[(n >> 1) - 1], sortbuf[n >> 1]), 0.5f);
    for (int i = tid; i < 100; i += nt) hist[i] = 0;
    __syncthreads();
    const float bpo = (float)bins_per_octave;
    for (int i = tid; i < n; i += nt) {
        if (cand_mag[i] >= thr) {
            const float f = cand_pitch[i];
            // hz_to_octs: float32 log2(f / 27.5); residual = mod(bpo * octs, 1.0) in float32
            const float octs = (float)log2((double)__fdiv_rn(f, 27.5f));
            const float x = __fmul_rn(bpo, octs);
            float r = __fsub_rn(x, floorf(x));
            if (r >= 0.5f) r = __fsub_rn(r, 1.0f);
            const double v = (double)r;
            int bi = (int)floor((v + 0.5) * 100.0);
            bi = bi < 0 ? 0 : (bi > 99 ? 99 : bi);
            while (bi > 0 && v < edges[bi]) --bi;
            while (bi < 99 && v >= edges[bi + 1]) ++bi;
            atomicAdd(&hist[bi], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0, bc = hist[0];
        for (int i = 1; i < 100; ++i)
            if (hist[i] > bc) { bc = hist[i]; best = i; }
        s_best = best;
    }
    __syncthreads();
    *empty = false;
    return s_best;
}

// piptrack candidate test at bin k of a magnitude column (S[k-1], S[k], S[k+1] given), librosa semantics.
__device__ __forceinline__ bool piptrack_candidate(float sm1, float s0, float sp1, float ref, int k, double bin_hz,
                                                   float* pitch, float* magv) {
    const float m0 = s0 > ref ? s0 : 0.f, mm = sm1 > ref ? sm1 : 0.f, mp = sp1 > ref ? sp1 : 0.f;
    if (!(m0 > mm && m0 >= mp)) return false;
    // numba stencil: a, b evaluated in float64 from float32 sums / differences
    const double a = (double)__fadd_rn(sp1, sm1) - 2.0 * (double)s0;
    const double bb = (double)__fsub_rn(sp1, sm1) / 2.0;
    const float shift = (fabs(bb) >= fabs(a)) ? 0.f : (float)(-bb / a);
    const float avg = __fmul_rn(__fsub_rn(sp1, sm1), 0.5f);                 // np.gradient interior
    const float dskew = __fmul_rn(__fmul_rn(0.5f, avg), shift);
    *pitch = (float)(((double)k + (double)shift) * bin_hz);
    *magv = __fadd_rn(s0, dskew);
    return true;
}

constexpr int kMaxCand = 4096;


}  // namespace bpc
