// Tuning estimation shared by the STFT-512 (chroma_stft, 12 bins/octave) and STFT-2048 (chroma_cens, 36 bins/octave)
// paths: librosa.estimate_tuning -> piptrack -> pitch_tuning (reached from process.py:52 and process.py:53).
#pragma once
#include <cfloat>
#include "common.cuh"

namespace bpc {

// ---------------------------------------------------------------------------- tuning estimation (shared with 2048)
// librosa.estimate_tuning -> piptrack -> pitch_tuning on candidate lists held in shared memory.
// cand_mag / cand_pitch: n candidates (mag + dskew > 0, pitch in Hz, both float32).  sortbuf: >= 516 words of scratch.
// Returns the histogram bin (0..99); sets *empty when the frequency set is empty (tuning 0.0 == bin 50).
__device__ inline int tuning_from_candidates(const float* cand_mag, const float* cand_pitch, int n, float* sortbuf,
                                      int* hist, const double* __restrict__ edges, int bins_per_octave, bool* empty) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_best;
    if (n == 0) {
        *empty = true;
        return 50;                   // edges[50] == 0.0
    }
    // np.median needs the middle order statistic(s) only: 4-pass byte-wise radix select on the bit patterns of the
    // (positive) magnitudes for the two ranks (n-1)/2 and n/2 at once.  (v1..v4 ran a full bitonic sort of up to 4096
    // candidates: 40 % of k_even2048's instructions.)  sortbuf is reused as 2 x 256 counters + 4 words of state.
    unsigned* cnt = reinterpret_cast<unsigned*>(sortbuf);          // [2][256]
    unsigned* state = cnt + 512;                                   // prefix[2], rank[2]
    if (tid == 0) {
        state[0] = state[1] = 0u;
        state[2] = (unsigned)((n - 1) >> 1);
        state[3] = (unsigned)(n >> 1);
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 512; i += nt) cnt[i] = 0u;
        __syncthreads();
        const unsigned p0 = state[0], p1 = state[1];
        const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < n; i += nt) {
            const unsigned key = __float_as_uint(cand_mag[i]);
            const unsigned d = (key >> shift) & 0xffu;
            if ((key & himask) == p0) atomicAdd(&cnt[d], 1u);
            if ((key & himask) == p1) atomicAdd(&cnt[256 + d], 1u);
        }
        __syncthreads();
        if (tid < 64) {
            // warp w < 2 resolves target w: 8 counters per lane, warp prefix sum, the lane holding the rank finishes
            const int w = tid >> 5, ln = tid & 31;
            const unsigned* hc = cnt + 256 * w;
            unsigned local = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) local += hc[ln * 8 + ((i + (ln >> 2)) & 7)];   // rotation: conflict-free banks
            unsigned incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (ln >= o) incl += v;
            }
            const unsigned excl = incl - local, r = state[2 + w];
            __syncwarp();
            if (r >= excl && r < incl) {
                unsigned c = excl;
                for (int i = 0; i < 8; ++i) {
                    const unsigned hv = hc[ln * 8 + i];
                    if (r < c + hv) {
                        state[w] |= (unsigned)(ln * 8 + i) << shift;
                        state[2 + w] = r - c;
                        break;
                    }
                    c += hv;
                }
            }
        }
        __syncthreads();
    }
    // np.median: middle element, or float32 mean of the two middle elements
    const float lo_mid = __uint_as_float(state[0]), hi_mid = __uint_as_float(state[1]);
    const float thr = (n & 1) ? hi_mid : __fmul_rn(__fadd_rn(lo_mid, hi_mid), 0.5f);
    __syncthreads();
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

// Candidate capacities.  A piptrack candidate is a strict local maximum (S[k] > S[k-1], S[k] >= S[k+1]), so two
// neighbouring bins never both qualify: at most ceil(bins / 2) per frame.
//   STFT-512 : bins 5..127 (123) x 63 frames  -> <= 62 * 63 = 3906
//   STFT-2048: bins 20..511 (492) x 32 frames -> <= 246 * 32 = 7872   (white noise reaches ~5200: r01 v27 overflowed 4096)
constexpr int kMaxCand = 4096;
constexpr int kMaxCand2048 = 7936;
constexpr int kSelectWords = 1024;       // scratch of tuning_from_candidates (2 x 256 counters + 4 words of state)


}  // namespace bpc
