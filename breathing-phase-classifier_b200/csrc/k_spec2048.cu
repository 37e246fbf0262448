// STFT-2048 group (librosa default n_fft, reached through methods.py:59-63,90 and process.py:74-75):
//   k_spec2048  one CTA per segment: per frame FP64 real FFT-2048 -> |X| -> spectral centroid / bandwidth / flatness /
//               contrast, mel-D power column; afterwards flux statistics, onset envelope, tempogram plane.
//   k_even2048  the hop-512 frames (= even hop-256 frames): spectral_rolloff with numpy's sequential float32 cumsum and
//               the 36-bins-per-octave tuning estimate chroma_cens needs.
#include <cmath>
#include "kernels.cuh"
#include "fft.cuh"
#include "tuning.cuh"

namespace bpc {

constexpr int kMag2048Stride = 1028;         // even-frame |X| workspace row stride (1025 valid)
constexpr int kTempoLags = 384;

// spectral_contrast sub-bands (librosa, fmin=200, n_bands=6, sr=16000, n_fft=2048): first bin, length, order count
__constant__ int c_band_lo[7] = {0, 25, 51, 102, 204, 409, 819};
__constant__ int c_band_len[7] = {25, 26, 51, 102, 205, 410, 206};
__constant__ int c_band_n[7] = {1, 1, 1, 2, 4, 8, 4};

// mean of the n smallest and n largest of magbuf[lo .. lo+len) by one warp (len <= 416, n <= 8).
// Ties are broken by position, so every element is selected at most once.
__device__ void warp_band_extremes(const float* magbuf, int lo, int len, int n, int lane, double* valley,
                                   double* peak) {
    float v[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int j = lane + 32 * i;
        v[i] = j < len ? magbuf[lo + j] : -1.f;          // magnitudes are >= 0; -1 marks "absent"
    }
    unsigned taken_hi = 0, taken_lo = 0;
    double sum_hi = 0.0, sum_lo = 0.0;
    for (int r = 0; r < n; ++r) {
        // largest remaining
        float best = -2.f;
        int bi = -1;
#pragma unroll
        for (int i = 0; i < 13; ++i)
            if (v[i] >= 0.f && !((taken_hi >> i) & 1u) && v[i] > best) { best = v[i]; bi = i; }
        int gidx = bi < 0 ? 0x7fffffff : lane + 32 * bi;
        float wb = best;
        int wi = gidx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, wb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (ob > wb || (ob == wb && oi < wi)) { wb = ob; wi = oi; }
        }
        if (wi == gidx && bi >= 0) taken_hi |= 1u << bi;
        sum_hi += (double)wb;
        // smallest remaining
        best = 3.0e38f;
        bi = -1;
#pragma unroll
        for (int i = 0; i < 13; ++i)
            if (v[i] >= 0.f && !((taken_lo >> i) & 1u) && v[i] < best) { best = v[i]; bi = i; }
        gidx = bi < 0 ? 0x7fffffff : lane + 32 * bi;
        wb = best;
        wi = gidx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, wb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (ob < wb || (ob == wb && oi < wi)) { wb = ob; wi = oi; }
        }
        if (wi == gidx && bi >= 0) taken_lo |= 1u << bi;
        sum_lo += (double)wb;
    }
    // np.mean of a float32 slice -> float32
    *peak = (double)(float)(sum_hi / (double)n);
    *valley = (double)(float)(sum_lo / (double)n);
}

struct Spec2048Frames {                       // live during the frame loop and the flux / onset stage
    double2 fbuf[1024];
    double2 tw[1024];
    float magbuf[1032];
    float melD[kPlaneRows * kMaxFrames];      // mel-D power [m*T + t]
    double cent[kMaxFrames], bw[kMaxFrames];
    float flat[kMaxFrames];
    double peak[7 * kMaxFrames], valley[7 * kMaxFrames];
};
struct Spec2048Smem {
    union {
        Spec2048Frames f;
        float tg[kTempoLags * kMaxFrames];    // tempogram autocorrelations [lag*T + t]; f is dead by then
    } u;
    float onset[kMaxFrames + 2 * 192 + 8];    // onset envelope with the tempogram's 192-sample pads
    float frame[kTempoLags + 8];
    float colmax[kMaxFrames];
    double dscratch[32];
    float fscratch[32];
};

__global__ void __launch_bounds__(256) k_spec2048(const float* __restrict__ y, Geometry g, Tables tb, Workspace ws,
                                                  float* feats, float* scalars) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Spec2048Smem& S = *reinterpret_cast<Spec2048Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, T = g.T, L = g.L, hop = g.hop;
    const float* yb = y + (size_t)b * L;

    for (int j = tid; j < 1024; j += 256) S.u.f.tw[j] = tb.tw1024[j];
    double w0[4], w1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = tid + 256 * i;
        w0[i] = tb.hann2048[2 * m];
        w1[i] = tb.hann2048[2 * m + 1];
    }
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const int g0 = t * hop - 1024;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = tid + 256 * i;
            const int gi = g0 + 2 * m;
            const float x0 = (gi >= 0 && gi < L) ? __ldg(yb + gi) : 0.f;
            const float x1 = (gi + 1 >= 0 && gi + 1 < L) ? __ldg(yb + gi + 1) : 0.f;
            S.u.f.fbuf[m] = make_double2((double)x0 * w0[i], (double)x1 * w1[i]);
        }
        __syncthreads();                      // also: previous frame's feature warps are done with magbuf
        fft_r4_dif<5, 256>(S.u.f.fbuf, S.u.f.tw, tid, SyncBlock());
        float* even_out = nullptr;
        if ((t & 1) == 0) even_out = ws.mag2048_even + ((size_t)b * ((T + 1) / 2) + (t >> 1)) * kMag2048Stride;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int k = tid + 256 * i;
            if (k <= 1024) {
                const float m = c64_abs(rfft_bin<5>(S.u.f.fbuf, tb.ptw2048, k));
                S.u.f.magbuf[k] = m;
                if (even_out) even_out[k] = m;
            }
        }
        __syncthreads();
        // ---- per-frame features, warp-specialised; no trailing barrier (the next frame's load only touches fbuf)
        if (warp == 0) {
            // spectral_centroid / bandwidth (methods.py:59-60) + flatness (methods.py:62)
            double sum = 0.0, slog = 0.0, spow = 0.0;
            for (int k = lane; k < 1025; k += 32) {
                const float m = S.u.f.magbuf[k];
                sum += (double)m;
                const float p = fmaxf(1e-10f, __fmul_rn(m, m));
                slog += (double)logf(p);
                spow += (double)p;
            }
            sum = warp_sum(sum);
            slog = warp_sum(slog);
            spow = warp_sum(spow);
            const double len = sum < 1.17549435e-38 ? 1.0 : sum;       // util.normalize(norm=1) threshold tiny(f32)
            double c = 0.0;
            for (int k = lane; k < 1025; k += 32) c += (double)k * 7.8125 * (double)(float)((double)S.u.f.magbuf[k] / len);
            c = warp_sum(c);
            double v = 0.0;
            for (int k = lane; k < 1025; k += 32) {
                const double d = fabs((double)k * 7.8125 - c);
                v += (double)(float)((double)S.u.f.magbuf[k] / len) * (d * d);
            }
            v = warp_sum(v);
            if (lane == 0) {
                S.u.f.cent[t] = c;
                S.u.f.bw[t] = sqrt(v);
                const float gmean = expf((float)(slog / 1025.0));
                const float amean = (float)(spow / 1025.0);
                S.u.f.flat[t] = __fdiv_rn(gmean, amean);
            }
        } else {
            // spectral_contrast order statistics (methods.py:63): one band per warp 1..7
            const int band = warp - 1;
            double va, pk;
            warp_band_extremes(S.u.f.magbuf, c_band_lo[band], c_band_len[band], c_band_n[band], lane, &va, &pk);
            if (lane == 0) { S.u.f.peak[band * T + t] = pk; S.u.f.valley[band * T + t] = va; }
        }
        // mel-D power column (n_fft 2048, 128 mels, fmax 8000): threads 128..255 take one row each
        if (tid >= 128) {
            const int m = tid - 128;
            const int s = tb.mel_d.start[m], c = tb.mel_d.count[m];
            const float* w = tb.mel_d.w + (size_t)m * tb.mel_d.width;
            float acc = 0.f;
            for (int j = 0; j < c; ++j) {
                const float mv = S.u.f.magbuf[s + j];
                acc = fmaf(__ldg(w + j), __fmul_rn(mv, mv), acc);
            }
            S.u.f.melD[m * T + t] = acc;
        }
    }
    __syncthreads();

    float* sc = scalars + (size_t)b * g.nscal;
    const int NP = kPlaneRows * T;
    // ---- centroid / bandwidth / flatness statistics (methods.py:64-68), warp 0..2
    if (warp < 3) {
        double s = 0.0, q = 0.0;
        for (int t = lane; t < T; t += 32) {
            const double v = warp == 0 ? S.u.f.cent[t] : (warp == 1 ? S.u.f.bw[t] : (double)S.u.f.flat[t]);
            s += v;
            q += v * v;
        }
        s = warp_sum(s);
        q = warp_sum(q);
        const double mean = s / T;
        double m2 = 0.0, m3 = 0.0;
        for (int t = lane; t < T; t += 32) {
            const double d = (warp == 0 ? S.u.f.cent[t] : (warp == 1 ? S.u.f.bw[t] : (double)S.u.f.flat[t])) - mean;
            m2 += d * d;
            m3 += d * d * d;
        }
        m2 = warp_sum(m2) / T;
        m3 = warp_sum(m3) / T;
        if (lane == 0) {
            const double sd = sqrt(m2);
            if (warp == 0) {
                sc[8] = (float)(mean / 8000.0);
                sc[9] = (float)(sd / 8000.0);
                sc[10] = (float)(m3 / (m2 * sqrt(m2)));                   // scipy.stats.skew (biased)
            } else if (warp == 1) {
                sc[11] = (float)(mean / 8000.0);
                sc[12] = (float)(sd / 8000.0);
            } else {
                sc[15] = (float)mean;
                sc[16] = (float)sd;
            }
        }
    }
    // ---- contrast = power_to_db(peak) - power_to_db(valley), each clamped at its own max - 80 (float64 arrays)
    {
        double pmax = -1e300, vmax = -1e300;
        for (int i = tid; i < 7 * T; i += 256) {
            const double p = 10.0 * log10(fmax(1e-10, S.u.f.peak[i]));
            const double v = 10.0 * log10(fmax(1e-10, S.u.f.valley[i]));
            S.u.f.peak[i] = p;
            S.u.f.valley[i] = v;
            pmax = fmax(pmax, p);
            vmax = fmax(vmax, v);
        }
        pmax = block_reduce(pmax, -1e300, OpMaxD(), S.dscratch);
        vmax = block_reduce(vmax, -1e300, OpMaxD(), S.dscratch);
        double s = 0.0, q = 0.0;
        for (int i = tid; i < 7 * T; i += 256) {
            const double c = fmax(S.u.f.peak[i], pmax - 80.0) - fmax(S.u.f.valley[i], vmax - 80.0);
            s += c;
            q += c * c;
        }
        s = block_sum(s, S.dscratch);
        q = block_sum(q, S.dscratch);
        if (tid == 0) {
            const double mean = s / (7.0 * T);
            sc[17] = (float)mean;
            sc[18] = (float)sqrt(fmax(0.0, q / (7.0 * T) - mean * mean));
        }
    }
    // ---- mel-D: L = 10 log10(max(1e-10, P)); flux uses ref=max (methods.py:90-92), onset uses ref=1 (process.py:74)
    float pmx = -FLT_MAX;
    for (int i = tid; i < NP; i += 256) pmx = fmaxf(pmx, S.u.f.melD[i]);
    pmx = block_max(pmx, S.fscratch);
    const float ref_db = (float)(10.0 * log10((double)fmaxf(1e-10f, pmx)));
    float lmax = -FLT_MAX;
    for (int i = tid; i < NP; i += 256) {
        const float l = __fmul_rn(10.0f, log10f(fmaxf(1e-10f, S.u.f.melD[i])));
        S.u.f.melD[i] = l;
        lmax = fmaxf(lmax, l);
    }
    lmax = block_max(lmax, S.fscratch);                                     // barrier inside: melD complete
    const float floor1 = __fsub_rn(lmax, 80.0f);                            // ref = 1.0 variant
    const float floorm = __fsub_rn(__fsub_rn(lmax, ref_db), 80.0f);         // ref = max variant
    // one warp per time step: flux[t] and onset difference d[t], t = 0..T-2
    float* flux = S.frame;                                                  // reuse (T-1 <= 384)
    for (int j = tid; j < T + 2 * 192 + 8; j += 256) S.onset[j] = 0.f;
    __syncthreads();
    for (int t = warp; t < T - 1; t += 8) {
        double f2 = 0.0, on = 0.0;
        for (int m = lane; m < kPlaneRows; m += 32) {
            const float l0 = S.u.f.melD[m * T + t], l1 = S.u.f.melD[m * T + t + 1];
            const float a0 = fmaxf(__fsub_rn(l0, ref_db), floorm), a1 = fmaxf(__fsub_rn(l1, ref_db), floorm);
            const float d = __fsub_rn(a1, a0);
            f2 += (double)__fmul_rn(d, d);
            const float o = __fsub_rn(fmaxf(l1, floor1), fmaxf(l0, floor1));
            on += (double)fmaxf(0.f, o);
        }
        f2 = warp_sum(f2);
        on = warp_sum(on);
        if (lane == 0) {
            flux[t] = sqrtf((float)f2);
            // onset_env = pad(mean over mels, (1 + 2048 // (2 * 256), 0))[:T]; stored at offset 192 (left tempogram pad)
            if (t + 5 < T) S.onset[192 + t + 5] = (float)(on / (double)kPlaneRows);
        }
    }
    __syncthreads();
    if (warp == 0) {
        double s = 0.0, q = 0.0;
        float mx = -FLT_MAX;
        for (int t = lane; t < T - 1; t += 32) {
            s += (double)flux[t];
            q += (double)flux[t] * (double)flux[t];
            mx = fmaxf(mx, flux[t]);
        }
        s = warp_sum(s);
        q = warp_sum(q);
        mx = warp_max(mx);
        if (lane == 0) {
            const double mean = s / (T - 1);
            sc[26] = (float)mean;
            sc[27] = (float)sqrt(fmax(0.0, q / (T - 1) - mean * mean));
            sc[28] = mx;
        }
    }
    if (ws.dbg_onset)
        for (int t = tid; t < T; t += 256) ws.dbg_onset[(size_t)b * T + t] = S.onset[192 + t];
    // ---- tempogram (process.py:75): linear-ramp pad 192, 384-sample Hann frames at hop 1, autocorrelation, /max
    if (tid < 192) {
        const float edge = S.onset[192 + T - 1];
        const float step = __fdiv_rn(edge, 192.0f);
        S.onset[192 + T + tid] = __fmul_rn((float)(191 - tid), step);       // np.pad(mode='linear_ramp', end 0)
    }
    __syncthreads();
    float* TG = S.u.tg;                                                    // [384 * T], aliases the FFT buffers
    for (int t = 0; t < T; ++t) {
        for (int n = tid; n < kTempoLags; n += 256) S.frame[n] = (float)((double)S.onset[t + n] * tb.hann384[n]);
        if (tid < 8) S.frame[kTempoLags + tid] = 0.f;
        __syncthreads();
        // structural zeros: onset[0..4] == 0 and the left pad is 0, so frame[n] == 0 for n < 197 - t
        const int n0 = (197 - t) > 0 ? (197 - t) : 0;
        for (int lag = tid; lag < kTempoLags; lag += 256) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int n = n0;
            const int nend = kTempoLags - lag;
            for (; n + 3 < nend; n += 4) {
                a0 = fmaf(S.frame[n], S.frame[n + lag], a0);
                a1 = fmaf(S.frame[n + 1], S.frame[n + 1 + lag], a1);
                a2 = fmaf(S.frame[n + 2], S.frame[n + 2 + lag], a2);
                a3 = fmaf(S.frame[n + 3], S.frame[n + 3 + lag], a3);
            }
            for (; n < nend; ++n) a0 = fmaf(S.frame[n], S.frame[n + lag], a0);
            TG[lag * T + t] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
    // column max |.| (util.normalize norm=inf), then whole-array z-score over all 384 rows (process.py:76)
    for (int t = warp; t < T; t += 8) {
        float mx = 0.f;
        for (int lag = lane; lag < kTempoLags; lag += 32) mx = fmaxf(mx, fabsf(TG[lag * T + t]));
        mx = warp_max(mx);
        if (lane == 0) S.colmax[t] = mx;
    }
    __syncthreads();
    double s = 0.0, q = 0.0;
    for (int i = tid; i < kTempoLags * T; i += 256) {
        const int t = i % T;
        const double len = (double)S.colmax[t] < 2.2250738585072014e-308 ? 1.0 : (double)S.colmax[t];
        const double v = (double)TG[i] / len;
        s += v;
        q += v * v;
    }
    s = block_sum(s, S.dscratch);
    q = block_sum(q, S.dscratch);
    const double mean = s / (double)(kTempoLags * T);
    const double sd = sqrt(fmax(0.0, q / (double)(kTempoLags * T) - mean * mean));
    float* o = plane_ptr(feats, b, BPC_CH_TEMPOGRAM, T);
    for (int i = tid; i < NP; i += 256) {                                    // pad_freq truncates to the first 128 lags
        const int t = i % T;
        const double len = (double)S.colmax[t] < 2.2250738585072014e-308 ? 1.0 : (double)S.colmax[t];
        o[i] = (float)(((double)TG[i] / len - mean) / (sd + 1e-8));
    }
}

void launch_spec2048(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                     float* scalars, cudaStream_t st) {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(k_spec2048, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Spec2048Smem));
        done = true;
    }
    k_spec2048<<<n, 256, sizeof(Spec2048Smem), st>>>(y, g, tb, ws, feats, scalars);
    note_launch();
}

// ------------------------------------------------------------------------------------ even frames: rolloff + tuning36
// spectral_rolloff (methods.py:61) is called without hop_length -> hop 512, i.e. the even hop-256 frames; its decision
// `cumsum(S) < 0.85 * cumsum(S)[-1]` is taken on numpy's *sequential float32* cumsum, which is reproduced here with
// one thread per frame (1025 dependent float32 adds).  chroma_cens (process.py:53) estimates its tuning from the same
// frames (estimate_tuning(y=y, bins_per_octave=36) -> piptrack n_fft 2048, hop 512).
struct Even2048Smem {
    float cand_mag[kMaxCand];
    float cand_pitch[kMaxCand];
    float sortbuf[kMaxCand];
    float colmax[kMaxFrames];
    float roll[kMaxFrames];
    int hist[100];
};

__global__ void __launch_bounds__(256) k_even2048(Geometry g, Tables tb, Workspace ws, float* scalars,
                                                  int32_t* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Even2048Smem& E = *reinterpret_cast<Even2048Smem*>(smem_raw);
    float* cand_mag = E.cand_mag;
    float* cand_pitch = E.cand_pitch;
    float* sortbuf = E.sortbuf;
    float* colmax = E.colmax;
    float* roll = E.roll;
    int* hist = E.hist;
    __shared__ int s_ncand;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, T = g.T, TE = (T + 1) / 2;
    const float* mag_b = ws.mag2048_even + (size_t)b * TE * kMag2048Stride;
    if (tid == 0) s_ncand = 0;
    // column maxima (piptrack threshold) -- warps 1..7
    if (warp > 0) {
        for (int f = warp - 1; f < TE; f += 7) {
            float mx = 0.f;
            for (int k = lane; k < 1025; k += 32) mx = fmaxf(mx, __ldg(mag_b + (size_t)f * kMag2048Stride + k));
            mx = warp_max(mx);
            if (lane == 0) colmax[f] = mx;
        }
    } else {
        // rolloff chains -- warp 0, one frame per lane (TE <= 32)
        for (int f = lane; f < TE; f += 32) {
            const float* col = mag_b + (size_t)f * kMag2048Stride;
            float c = 0.f;
            for (int k = 0; k < 1025; ++k) c = __fadd_rn(c, __ldg(col + k));
            const float thr = __fmul_rn(0.85f, c);
            float c2 = 0.f;
            int kk = 1024;
            for (int k = 0; k < 1025; ++k) {
                c2 = __fadd_rn(c2, __ldg(col + k));
                if (!(c2 < thr)) { kk = k; break; }
            }
            roll[f] = (float)kk * 7.8125f;
        }
    }
    __syncthreads();
    if (warp == 0) {
        double s = 0.0, q = 0.0;
        for (int f = lane; f < TE; f += 32) { s += (double)roll[f]; q += (double)roll[f] * (double)roll[f]; }
        s = warp_sum(s);
        q = warp_sum(q);
        if (lane == 0) {
            const double mean = s / TE;
            float* sc = scalars + (size_t)b * g.nscal;
            sc[13] = (float)(mean / 8000.0);
            sc[14] = (float)(sqrt(fmax(0.0, q / TE - mean * mean)) / 8000.0);
        }
    }
    // piptrack: bins 20..511 (150 Hz <= k * 7.8125 < 4000 Hz)
    const int nb = 492;
    for (int idx = tid; idx < nb * TE; idx += 256) {
        const int f = idx / nb, k = 20 + idx - f * nb;
        const float* col = mag_b + (size_t)f * kMag2048Stride;
        float pitch, mv;
        if (piptrack_candidate(__ldg(col + k - 1), __ldg(col + k), __ldg(col + k + 1), __fmul_rn(0.1f, colmax[f]), k,
                               7.8125, &pitch, &mv)) {
            const int slot = atomicAdd(&s_ncand, 1);
            if (slot < kMaxCand) { cand_mag[slot] = mv; cand_pitch[slot] = pitch; }
        }
    }
    __syncthreads();
    int n = s_ncand;
    unsigned flags = 0;
    if (n > kMaxCand) { n = kMaxCand; flags |= BPC_SEG_CAND_OVERFLOW; }
    bool empty = false;
    const int tbin = tuning_from_candidates(cand_mag, cand_pitch, n, sortbuf, hist, tb.hist_edges, 36, &empty);
    if (empty) flags |= BPC_SEG_TUNING_EMPTY;
    if (tid == 0) {
        ws.tuning[b * 2 + 1] = tbin;
        if (status && flags) atomicOr((unsigned int*)&status[b], flags);
    }
}

void launch_even2048(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* scalars,
                     int32_t* status, cudaStream_t st) {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(k_even2048, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Even2048Smem));
        done = true;
    }
    k_even2048<<<n, 256, sizeof(Even2048Smem), st>>>(g, tb, ws, scalars, status);
    note_launch();
}

}  // namespace bpc
