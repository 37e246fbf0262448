// chroma_cens rows of the chroma channel (process.py:53-57): librosa.feature.chroma_cens(y, sr=16000, hop_length=256)
//   = 7-octave / 36-bins-per-octave constant-Q transform (multi-rate recursion: rectangular-window STFT-512 of the
//     signal decimated by 2 per octave, times one sparse complex basis that is the same for every octave),
//     |.| / sqrt(filter length), fold 252 -> 12, L1 normalise, 4-level quantise, 41-tap Hann smoothing, L2 normalise.
// One CTA per segment; one warp owns a frame t and walks the 7 octaves, so the 12 chroma sums stay in registers.
// The decimator stands in for soxr_hq (absent library): the same 127-tap Kaiser half-band the oracle uses.
#include <cmath>
#include "kernels.cuh"
#include "fft.cuh"
#include "tables.hpp"

namespace bpc {

constexpr int kBinLo = 60, kBinHi = 144;               // rfft bins the sparsified bases can touch (measured 62..141)
constexpr int kBinSpan = kBinHi - kBinLo + 1;

struct CensSmem {
    float dec[8000 + 4000 + 2000 + 1000 + 500 + 250 + 64];   // decimated signals, octaves 1..6
    double2 fbuf[8][256];
    float2 spec[8][kBinSpan + 3];
    float cqmag[8][kCqtBinsPerOct];
    float chroma[12 * kMaxFrames];
    float quant[12 * kMaxFrames];
    double hb[kHalfbandTaps];
    double swin[43];
    double dscratch[32];
    float fscratch[32];
};

__global__ void __launch_bounds__(256) k_cens(const float* __restrict__ y, Geometry g, Tables tb, Workspace ws,
                                              float* feats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CensSmem& S = *reinterpret_cast<CensSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L, T = g.T;
    const float* yb = y + (size_t)b * L;

    for (int i = tid; i < kHalfbandTaps; i += 256) S.hb[i] = tb.halfband[i];
    if (tid < 43) {
        // scipy.signal.get_window('hann', 43, fftbins=False), normalised to unit sum in the smoothing loop below
        S.swin[tid] = 0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * (double)tid / 42.0);
    }
    __syncthreads();

    // ---- six cascaded 2:1 decimations: out[n] = f32( sum_k h[k] * in[2n + 63 - k] / sqrt(0.5) ), zero extension
    int len_in = L, off_out = 0;
    const float* in_g = yb;
    const float* in_s = nullptr;
    int offs[7], lens[7];
    offs[0] = -1; lens[0] = L;
    const double inv_s = 1.0 / sqrt(0.5);
    for (int o = 1; o <= 6; ++o) {
        const int len_out = (len_in + 1) / 2;
        for (int n = tid; n < len_out; n += 256) {
            const int c0 = 2 * n;
            double acc = 0.0;
            // centre tap
            {
                const float v = c0 < len_in ? (in_s ? in_s[c0] : __ldg(in_g + c0)) : 0.f;
                acc = S.hb[63] * (double)v;
            }
            // odd offsets (the even-offset taps of a half-band filter vanish)
#pragma unroll 4
            for (int d = 1; d <= 63; d += 2) {
                const int i0 = c0 - d, i1 = c0 + d;
                const float v0 = (i0 >= 0 && i0 < len_in) ? (in_s ? in_s[i0] : __ldg(in_g + i0)) : 0.f;
                const float v1 = (i1 >= 0 && i1 < len_in) ? (in_s ? in_s[i1] : __ldg(in_g + i1)) : 0.f;
                acc += S.hb[63 + d] * (double)v0 + S.hb[63 - d] * (double)v1;
            }
            S.dec[off_out + n] = (float)(acc * inv_s);
        }
        __syncthreads();
        offs[o] = off_out; lens[o] = len_out;
        in_s = S.dec + off_out;
        in_g = nullptr;
        off_out += len_out;
        len_in = len_out;
    }

    // ---- CQT -> chroma fold, one warp per frame
    const int tun = ws.tuning[b * 2 + 1];
    const int16_t* bcol = tb.cqt_col + (size_t)tun * kCqtBinsPerOct * kCqtEllWidth;
    const float* bre = tb.cqt_re + (size_t)tun * kCqtBinsPerOct * kCqtEllWidth;
    const float* bim = tb.cqt_im + (size_t)tun * kCqtBinsPerOct * kCqtEllWidth;
    const double* slen = tb.cqt_sqrt_len + (size_t)tun * kCqtBins;
    double2* buf = S.fbuf[warp];
    for (int t = warp; t < T; t += 8) {
        float csum = 0.f;                                        // lanes 0..11: chroma c = lane
        for (int o = 0; o < kCqtOctaves; ++o) {
            const int hop = g.hop >> o;
            const int lo = lens[o];
            const float* sg = o == 0 ? nullptr : S.dec + offs[o];
            const int g0 = t * hop - 256;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = lane + 32 * i;
                const int gi = g0 + 2 * m;
                float x0 = 0.f, x1 = 0.f;
                if (gi >= 0 && gi < lo) x0 = sg ? sg[gi] : __ldg(yb + gi);
                if (gi + 1 >= 0 && gi + 1 < lo) x1 = sg ? sg[gi + 1] : __ldg(yb + gi + 1);
                buf[swz(m)] = make_double2((double)x0, (double)x1);   // window = 'ones'
            }
            __syncwarp();
            warp_fft_r4<4>(buf, tb.twp256, lane);
            for (int k = kBinLo + lane; k <= kBinHi; k += 32) {
                const double2 X = rfft_bin<4, true>(buf, tb.ptw512, k);
                S.spec[warp][k - kBinLo] = make_float2((float)X.x, (float)X.y);   // complex64 STFT
            }
            __syncwarp();
            const double scale = sqrt((double)(1 << o));           // fft_basis *= sqrt(sr / my_sr)
            for (int r = lane; r < kCqtBinsPerOct; r += 32) {
                double cr = 0.0, ci = 0.0;
                for (int j = 0; j < kCqtEllWidth; ++j) {
                    const int col = bcol[r * kCqtEllWidth + j];
                    if (col < 0) break;
                    const float br = (float)((double)bre[r * kCqtEllWidth + j] * scale);
                    const float bi = (float)((double)bim[r * kCqtEllWidth + j] * scale);
                    const float2 d = S.spec[warp][col - kBinLo];
                    cr += (double)br * (double)d.x - (double)bi * (double)d.y;
                    ci += (double)br * (double)d.y + (double)bi * (double)d.x;
                }
                // complex64 response, then V /= sqrt(lengths) (complex128 math, complex64 store), then |V|
                const float r32 = (float)cr, i32 = (float)ci;
                const double sl = slen[kCqtBins - kCqtBinsPerOct * (o + 1) + r];
                S.cqmag[warp][r] = c64_abs(make_double2((double)r32 / sl, (double)i32 / sl));
            }
            __syncwarp();
            if (lane < 12) {
                // cq_to_chroma: chroma c sums bins {3c-1, 3c, 3c+1} (mod 36) of every octave
                const int j0 = (3 * lane + 35) % 36;
                csum += S.cqmag[warp][j0] + S.cqmag[warp][3 * lane] + S.cqmag[warp][3 * lane + 1];
            }
            __syncwarp();
        }
        if (lane < 12) S.chroma[lane * T + t] = csum;
    }
    __syncthreads();
    // ---- CENS post-processing per column: L1 normalise, quantise
    for (int t = tid; t < T; t += 256) {
        double l1 = 0.0;
        for (int c = 0; c < 12; ++c) l1 += fabs((double)S.chroma[c * T + t]);
        if (l1 < 1.17549435e-38) l1 = 1.0;
        for (int c = 0; c < 12; ++c) {
            const float v = (float)((double)S.chroma[c * T + t] / l1);
            S.quant[c * T + t] = 0.25f * (float)((v > 0.4f) + (v > 0.2f) + (v > 0.1f) + (v > 0.05f));
        }
    }
    __syncthreads();
    // 41 non-zero taps of hann(43) / sum, scipy.ndimage.convolve(mode='constant') along time
    double wsum = 0.0;
    for (int j = 0; j < 43; ++j) wsum += S.swin[j];
    for (int i = tid; i < 12 * T; i += 256) {
        const int c = i / T, t = i - c * T;
        double acc = 0.0;
        for (int j = 0; j < 43; ++j) {
            const int tt = t + 21 - j;
            if (tt >= 0 && tt < T) acc += (S.swin[j] / wsum) * (double)S.quant[c * T + tt];
        }
        S.chroma[i] = (float)acc;
    }
    __syncthreads();
    // L2 normalise each column
    for (int t = tid; t < T; t += 256) {
        double l2 = 0.0;
        for (int c = 0; c < 12; ++c) l2 += (double)S.chroma[c * T + t] * (double)S.chroma[c * T + t];
        l2 = sqrt(l2);
        if (l2 < 1.17549435e-38) l2 = 1.0;
        for (int c = 0; c < 12; ++c) S.quant[c * T + t] = (float)((double)S.chroma[c * T + t] / l2);
    }
    __syncthreads();
    if (ws.dbg_chroma_cens) {
        float* d = ws.dbg_chroma_cens + (size_t)b * 12 * T;
        for (int i = tid; i < 12 * T; i += 256) d[i] = S.quant[i];
    }
    // ---- row-wise z-score -> rows 12..23; pad rows 24..127 with the min over all 24 normalised rows
    float mn = FLT_MAX;
    float* o = plane_ptr(feats, b, BPC_CH_CHROMA, T);
    for (int r = warp; r < 12; r += 8) {
        const ZTerm z = np_row_zterm(S.quant + r * T, T, lane);
        for (int t = lane; t < T; t += 32) {
            const float v = z(S.quant[r * T + t]);
            o[(12 + r) * T + t] = v;
            mn = fminf(mn, v);
        }
    }
    mn = block_min(mn, S.fscratch);
    const float fill = fminf(mn, ws.chroma_min[b * 2 + 0]);
    for (int i = 24 * T + tid; i < kPlaneRows * T; i += 256) o[i] = fill;
    if (tid == 0) ws.chroma_min[b * 2 + 1] = mn;
}

void launch_cens(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                 cudaStream_t st) {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(k_cens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CensSmem));
        done = true;
    }
    k_cens<<<n, 256, sizeof(CensSmem), st>>>(y, g, tb, ws, feats);
    note_launch();
}

}  // namespace bpc
