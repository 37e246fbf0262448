// STFT-512 group: ingest (pad_or_truncate), batched framing + Hann + FP64 real FFT-512 -> |X| workspace, and the four
// consumer roles of that spectrogram (one CTA per (segment, role)):
//   role 0  mel / mel_delta / mel_delta2 / mod_spec   (process.py:32-41, 69-72; methods.py:142-143)
//   role 1  mfcc (+delta, +delta2, row-wise z)          (process.py:43-49)
//   role 2  "gammatone" = log1p(mel64 @ |X|)            (process.py:59-62; methods.py:136-140)
//   role 3  chroma_stft rows + low-frequency ratio      (process.py:51-52; methods.py:84-88)
#include <atomic>
#include <cmath>
#include <cstdlib>
#include "kernels.cuh"
#include "fft.cuh"
#include "fft_reg.cuh"
#include "tuning.cuh"

namespace bpc {

static std::atomic<int64_t> g_launches{0};
int64_t launches_issued() { return g_launches.load(); }
void note_launch(int n) { g_launches.fetch_add(n); }

// ------------------------------------------------------------------------------------------------- ingest
// methods.py:24-28 pad_or_truncate (+ soundfile's int16 / 32768 when the caller passes PCM16).
__global__ void k_ingest(const void* __restrict__ wav, int dtype, long long L_in, float* __restrict__ y, int L,
                         long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / L;
        const int n = (int)(i - b * L);
        float v = 0.f;
        if (n < L_in) {
            if (dtype == 1) v = (float)((const short*)wav)[b * L_in + n] * (1.0f / 32768.0f);
            else v = ((const float*)wav)[b * L_in + n];
        }
        y[i] = v;
    }
}

void launch_ingest(const void* wav, int wav_dtype, int64_t L_in, float* y, int n, const Geometry& g, cudaStream_t st) {
    const long long total = (long long)n * g.L;
    const int blocks = (int)((total + 1023) / 1024 < 148 * 16 ? (total + 1023) / 1024 : 148 * 16);
    k_ingest<<<blocks, 256, 0, st>>>(wav, wav_dtype, (long long)L_in, y, g.L, total);
    note_launch();
}

// ------------------------------------------------------------------------------------------------ STFT-512
// One half-warp ("team") per frame: 16 lanes x 16 register-resident complex points, one shared-memory exchange
// (fft_reg.cuh).  librosa.stft semantics: zero centre padding, periodic Hann (float64) * float32 samples, float64 FFT,
// complex64 rounding, |.| as hypotf.
constexpr int kS512Threads = 128;

__global__ void __launch_bounds__(kS512Threads, 4) k_stft512(const float* __restrict__ y, Geometry g, Tables tb,
                                                          float* __restrict__ mag, int total_frames) {
    __shared__ double2 s_xch[kS512Threads / 16][16 * 17];
    const int lane = threadIdx.x & 31, h = lane & 15, team = threadIdx.x >> 4;
    const int partner = (lane & 16) | ((16 - h) & 15);
    const int T = g.T, L = g.L, hop = g.hop;
    double2* xch = s_xch[team];
    double2 tw[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tw[k1] = __ldg(tb.tw256 + ((h * k1) & 255));
    const double2 wp = __ldg(tb.ptw512 + h);
    const double2 wl = make_double2(wp.y, -wp.x);                         // -i * exp(-2 pi i h / 512)
    const double2* win2 = reinterpret_cast<const double2*>(tb.hann512);
    const int per_iter = gridDim.x * (kS512Threads / 16);
    // trip count from blockIdx only (every team runs every iteration, `valid` gates the stores): convergent shuffles
    for (int f0 = blockIdx.x * (kS512Threads / 16); f0 < total_frames; f0 += per_iter) {
        const int f = f0 + team;
        const bool valid = f < total_frames;
        const int fc = valid ? f : total_frames - 1;
        const int b = fc / T, t = fc - b * T;
        const float* yb = y + (size_t)b * L;
        const int g0 = t * hop - 256;
        double2 a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int m = h + 16 * j;
            const int gi = g0 + 2 * m;                         // even; L is even
            float2 v = make_float2(0.f, 0.f);
            if (gi >= 0 && gi < L) v = __ldg(reinterpret_cast<const float2*>(yb + gi));
            const double2 w = __ldg(win2 + m);
            a[j] = make_double2((double)v.x * w.x, (double)v.y * w.y);
        }
        team_fft<16>(a, tw, 1, xch, h);
        float* out = mag + (size_t)fc * kMagStride;
        // every conjugate pair once (bins 0 .. 256, the Nyquist bin as the second output of k = 0); float32-pair
        // magnitudes without a per-bin range check: the range is tracked over the row and the row redone exactly in the
        // rare case (an all-zero frame) -- one warp-uniform branch per row (fft.cuh)
        float xlo = 1.f, xhi = 1.f;
        auto emit = [&](int k, double2 t2) {
            float x;
            const float re = 0.5f * (float)t2.x;
            float v = c64_abs_f32_unchecked(re, 0.5f * (float)t2.y, &x);
            if (k == 256) { v = fabsf(re); x = 1.f; }         // X[N] is real: |.| without the hypot (and never out of range)
            xlo = fminf(xlo, x);
            xhi = fmaxf(xhi, x);
            if (valid) out[k] = v;
        };
        team_rsplit_pairs<16, 0>(a, wl, h, partner, emit);
        if (__any_sync(0xffffffffu, !(c64_abs_in_range(xlo) && c64_abs_in_range(xhi)))) {
            auto emit_exact = [&](int k, double2 t2) {
                const float re = 0.5f * (float)t2.x;
                const float v = k == 256 ? fabsf(re) : c64_abs_exact(re, 0.5f * (float)t2.y);
                if (valid) out[k] = v;
            };
            team_rsplit_pairs<16, 0>(a, wl, h, partner, emit_exact);
        }
    }
}

void launch_stft512(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, cudaStream_t st) {
    const int total = n * g.T;
    const int per_cta = kS512Threads / 16;
    int grid = (total + per_cta - 1) / per_cta;
    if (grid > 148 * 16) grid = 148 * 16;
    k_stft512<<<grid, kS512Threads, 0, st>>>(y, g, tb, ws.mag512, total);
    note_launch();
}

// ------------------------------------------------------------------------------------- shared device pieces
// librosa.power_to_db(S, ref, amin=1e-10, top_db=80) on a shared-memory array, in place (float32 like numpy).
__device__ void power_to_db_inplace(float* P, int n, bool ref_is_max, float* fscratch) {
    float ref_db = 0.f;
    if (ref_is_max) {
        float mx = -FLT_MAX;
        for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, P[i]);
        mx = block_max(mx, fscratch);
        ref_db = (float)(10.0 * log10((double)fmaxf(1e-10f, mx)));
    }
    float vmax = -FLT_MAX;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(1e-10f, P[i]))), ref_db);
        P[i] = v;
        vmax = fmaxf(vmax, v);
    }
    vmax = block_max(vmax, fscratch);
    const float floor_db = __fsub_rn(vmax, 80.0f);
    for (int i = threadIdx.x; i < n; i += blockDim.x) P[i] = fmaxf(P[i], floor_db);
    __syncthreads();
}

// librosa.feature.delta(width=9, order, mode='interp') == scipy savgol_filter(9, polyorder=order, deriv=order):
// interior correlation; the 4 edge samples on each side take the (constant) order-th derivative of the edge fit,
// i.e. the interior value at t=4 / t=T-5.  float64 accumulate (scaled by the reciprocal of the normaliser: the
// difference to a float64 division disappears in the float32 rounding), float32 store.
__device__ __forceinline__ double delta1_taps(double m4, double m3, double m2, double m1, double p1, double p2,
                                              double p3, double p4) {
    return fma(4.0, p4 - m4, fma(3.0, p3 - m3, fma(2.0, p2 - m2, p1 - m1))) * (1.0 / 60.0);
}
__device__ __forceinline__ double delta2_taps(double m4, double m3, double m2, double m1, double c, double p1, double p2,
                                              double p3, double p4) {
    return fma(28.0, p4 + m4, fma(7.0, p3 + m3, fma(-8.0, p2 + m2, fma(-17.0, p1 + m1, -20.0 * c)))) * (1.0 / 462.0);
}
__device__ __forceinline__ float delta_at(const float* row, int t, int T, int order) {
    const int tc = t < 4 ? 4 : (t > T - 5 ? T - 5 : t);
    const float* r = row + tc;
    if (order == 1)
        return (float)delta1_taps((double)r[-4], (double)r[-3], (double)r[-2], (double)r[-1], (double)r[1], (double)r[2],
                                  (double)r[3], (double)r[4]);
    return (float)delta2_taps((double)r[-4], (double)r[-3], (double)r[-2], (double)r[-1], (double)r[0], (double)r[1],
                              (double)r[2], (double)r[3], (double)r[4]);
}

// Both deltas of `rows` rows of length T (src[r * T + t]) by the whole CTA: a work item is one run of 8 consecutive
// centres of one row, whose 16 samples are widened to double ONCE and slide through registers (delta_at widens nine
// values per output and order: the quarter-rate F2F.F64.F32 conversions were a quarter of role_mel's stencil phase).
// d1 / d2 receive [rows, T]; the four edge frames on either side repeat the value of centre 4 / T - 5 (see above).
// If `sums` is set, sums[0..3] accumulate this thread's sum / sum of squares of the d1 and d2 values it stored.
// No barrier inside: the caller synchronises before anyone else reads d1 / d2.
__device__ __forceinline__ void delta_rows_block(const float* src, int rows, int T, float* d1, float* d2, double* sums) {
    const int runs = (T - 8 + 7) >> 3;                           // centres 4 .. T - 5 in runs of 8
    double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0;
    for (int item = threadIdx.x; item < rows * runs; item += blockDim.x) {
        const int r = item / runs, run = item - r * runs;
        const int c0 = 4 + 8 * run, nc = min(8, T - 4 - c0);    // centres c0 .. c0 + nc - 1
        const float* row = src + (size_t)r * T;
        double w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = (double)row[min(c0 - 4 + j, T - 1)];
        float* o1 = d1 + (size_t)r * T;
        float* o2 = d2 + (size_t)r * T;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nc) {
                const float v1 = (float)delta1_taps(w[j], w[j + 1], w[j + 2], w[j + 3], w[j + 5], w[j + 6], w[j + 7], w[j + 8]);
                const float v2 = (float)delta2_taps(w[j], w[j + 1], w[j + 2], w[j + 3], w[j + 4], w[j + 5], w[j + 6], w[j + 7], w[j + 8]);
                const int c = c0 + j;
                int reps = 1;
                o1[c] = v1;
                o2[c] = v2;
                if (c == 4) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { o1[e] = v1; o2[e] = v2; }
                    reps += 4;
                }
                if (c == T - 5) {
#pragma unroll
                    for (int e = 1; e <= 4; ++e) { o1[T - 5 + e] = v1; o2[T - 5 + e] = v2; }
                    reps += 4;
                }
                const double a = (double)v1, b = (double)v2, n = (double)reps;
                s1 = fma(n, a, s1); q1 = fma(n * a, a, q1);
                s2 = fma(n, b, s2); q2 = fma(n * b, b, q2);
            }
        }
    }
    if (sums) { sums[0] = s1; sums[1] = q1; sums[2] = s2; sums[3] = q2; }
}

// Band-form filterbank applied to |X| (power = false) or |X|^2 (power = true): out[m*T + t], float32 accumulate.
// ROWS (a power of two) consecutive threads own consecutive mel rows of one frame; the weights come from the transposed
// band table ([tap][row], one coalesced request per tap, zero beyond a row's count), the loop runs to the largest count
// of the warp's rows.
template <int ROWS>
__device__ void apply_bank(const BankDev& bank, const float* __restrict__ mag_b, int T, bool power, float* out) {
    // every thread owns one mel row of TWO frames (t, t + half): the tap weight is loaded once for both and the two
    // accumulation chains are independent
    const int half = (T + 1) >> 1;
    const int total = ROWS * half;
    const int iters = (total + blockDim.x - 1) / blockDim.x;
    for (int it = 0; it < iters; ++it) {
        const int idx = it * blockDim.x + threadIdx.x;
        const bool live = idx < total;
        const int m = idx & (ROWS - 1), t0 = live ? idx / ROWS : 0, t1 = t0 + half;
        const bool live1 = live && t1 < T;
        const int s = __ldg(bank.start + m), c = live ? __ldg(bank.count + m) : 0;
        const int cmax = __reduce_max_sync(0xffffffffu, c);
        const float* src0 = mag_b + (size_t)t0 * kMagStride;
        const float* src1 = mag_b + (size_t)(live1 ? t1 : t0) * kMagStride;
        const float* wt = bank.wt + m;
        float acc0 = 0.f, acc1 = 0.f;
        // taps in rounds of four (the transposed table carries four zero taps past the widest row, upload_bank): twelve
        // independent loads in flight per round instead of three -- the kernel waits on these loads (long scoreboard 3.8
        // warps per issue); same taps in the same order, a zero weight adds +-0
        for (int j = 0; j < cmax; j += 4) {
            float w[4], v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = min(s + j + u, 256);
                w[u] = __ldg(wt + (j + u) * ROWS);
                v0[u] = __ldg(src0 + k);
                v1[u] = __ldg(src1 + k);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc0 = fmaf(w[u], power ? __fmul_rn(v0[u], v0[u]) : v0[u], acc0);
                acc1 = fmaf(w[u], power ? __fmul_rn(v1[u], v1[u]) : v1[u], acc1);
            }
        }
        if (live) out[m * T + t0] = acc0;
        if (live1) out[m * T + t1] = acc1;
    }
    __syncthreads();
}

// C[k*T + t] = sum_n D[k*128 + n] * P[n*T + t], k < 40 (ortho DCT-II along the mel axis, first 40 rows).
// 256 threads = 8 row groups (5 rows, warp-uniform: D is read as broadcast float4) x 32 column pairs; float32 partial
// sums over 32 n combined in float64.  Dg: the [40, 128] matrix in global memory (every load is warp-uniform, i.e. one
// broadcast sector out of L1; staging it in shared memory cost 20 KB per CTA and one resident CTA per SM).
__device__ void dct_mel40(const float* __restrict__ Dg, const float* P, int T, float* C) {
    const int kb = threadIdx.x >> 5, tp = threadIdx.x & 31;
    for (int tb0 = 0; tb0 < T; tb0 += 64) {                    // one pass per 64 columns (a single pass when T <= 64)
        // columns tp and tp + 32 of the pass: the lanes of a warp read consecutive words of a P row (the r01 pairing
        // 2 tp, 2 tp + 1 was a two-way bank conflict on every load: half of this kernel's excess shared wavefronts)
        const int t0 = tb0 + tp;
        const bool has1 = t0 + 32 < T;
        const int t1 = has1 ? t0 + 32 : T - 1;
        if (kb < 8 && t0 < T) {
            double acc[5][2];
#pragma unroll
            for (int i = 0; i < 5; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                float part[5][2];
#pragma unroll
                for (int i = 0; i < 5; ++i) part[i][0] = part[i][1] = 0.f;
#pragma unroll 2
                for (int n = blk * 32; n < blk * 32 + 32; n += 4) {
                    float p0[4], p1[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { p0[q] = P[(n + q) * T + t0]; p1[q] = P[(n + q) * T + t1]; }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float4 d = __ldg(reinterpret_cast<const float4*>(Dg + (5 * kb + i) * 128 + n));
                        part[i][0] = fmaf(d.x, p0[0], part[i][0]); part[i][1] = fmaf(d.x, p1[0], part[i][1]);
                        part[i][0] = fmaf(d.y, p0[1], part[i][0]); part[i][1] = fmaf(d.y, p1[1], part[i][1]);
                        part[i][0] = fmaf(d.z, p0[2], part[i][0]); part[i][1] = fmaf(d.z, p1[2], part[i][1]);
                        part[i][0] = fmaf(d.w, p0[3], part[i][0]); part[i][1] = fmaf(d.w, p1[3], part[i][1]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 5; ++i) { acc[i][0] += (double)part[i][0]; acc[i][1] += (double)part[i][1]; }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                C[(5 * kb + i) * T + t0] = (float)acc[i][0];
                if (has1) C[(5 * kb + i) * T + t1] = (float)acc[i][1];
            }
        }
    }
    __syncthreads();
}

// C2[k*T + u] = sum_t DT[u][t] * C1[k*T + t] (ortho DCT-II along time); DTs is the transposed matrix [t][u] staged in
// shared memory.  Same 5 x 2 register tile; float32 partial sums over 16 t combined in float64.
__device__ __forceinline__ void dct_time40_tile(const float* DTs, const float* C1, int T, float* C2, int ub0) {
    const int kb = threadIdx.x >> 5, up = threadIdx.x & 31;
    const int u0 = ub0 + up;                                   // columns up and up + 32 of the tile (conflict-free rows of DTs)
    const bool has1 = u0 + 32 < T;
    const int u1 = has1 ? u0 + 32 : T - 1;
    if (kb < 8 && u0 < T) {
        double acc[5][2];
        float part[5][2];
#pragma unroll
        for (int i = 0; i < 5; ++i) { acc[i][0] = acc[i][1] = 0.0; part[i][0] = part[i][1] = 0.f; }
        const float* dp = DTs + u0;
        const int du = u1 - u0;
        for (int t = 0; t < T; ++t, dp += T) {
            const float d0 = dp[0], d1 = dp[du];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float c = C1[(5 * kb + i) * T + t];
                part[i][0] = fmaf(d0, c, part[i][0]);
                part[i][1] = fmaf(d1, c, part[i][1]);
            }
            if ((t & 15) == 15) {
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    acc[i][0] += (double)part[i][0]; acc[i][1] += (double)part[i][1];
                    part[i][0] = part[i][1] = 0.f;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            C2[(5 * kb + i) * T + u0] = (float)(acc[i][0] + (double)part[i][0]);
            if (has1) C2[(5 * kb + i) * T + u1] = (float)(acc[i][1] + (double)part[i][1]);
        }
    }
}

__device__ void dct_time40(const float* DTs, const float* C1, int T, float* C2) {
    for (int ub0 = 0; ub0 < T; ub0 += 64) dct_time40_tile(DTs, C1, T, C2, ub0);   // one pass per 64 output columns
    __syncthreads();
}

// copy a constant matrix into shared memory (all threads; caller syncs)
__device__ __forceinline__ void stage_matrix(float* dst, const float* __restrict__ src, int n) {
    // eight trips' loads in flight before the first store (the plain loop: one L2 round trip per trip, 14 trips)
    const int nt = blockDim.x;
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * nt) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + min(i0 + u * nt, n - 1));
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + u * nt < n) dst[i0 + u * nt] = v[u];
    }
}

// whole-array statistics of a shared-memory array
__device__ ZTerm zterm_of(const float* a, int n, double* dscratch, float* fscratch, float* min_out) {
    double s = 0.0, q = 0.0;
    float mn = FLT_MAX;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = (double)a[i];
        s += v;
        q += v * v;
        mn = fminf(mn, a[i]);
    }
    block_sum2(s, q, dscratch);
    if (min_out) *min_out = block_min(mn, fscratch);
    return make_zterm(s, q, (double)n);
}

// ------------------------------------------------------------------------------------------- role 0: mel3+mod
template <bool LONG>
__device__ void role_mel(int b, const Geometry g, const Tables& tb, const Workspace& ws, float* feats, float* mel3,
                         float* smem, double* dscratch, float* fscratch) {
    // `smem` is the role's private array space: shared memory for 1 s segments, a per-segment global scratch region in
    // long mode (LONG), where the time DCT matrix is also read from its global table instead of being staged
    const int T = g.T, NP = kPlaneRows * T;
    float* P = smem;                 // [128*T] mel power -> mel_db
    float* C1 = P + NP;              // [40*T]
    float* C2 = C1 + 40 * T;         // [40*T]
    const float* DTs = LONG ? tb.dct_time : C2 + 40 * T;      // [T*T] DCT matrix (time axis), transposed
    const float* mag_b = ws.mag512 + (size_t)b * T * kMagStride;

    if (!mel3 && !LONG) stage_matrix(C2 + 40 * T, tb.dct_time, T * T);
    apply_bank<128>(tb.mel_a, mag_b, T, true, P);
    power_to_db_inplace(P, NP, true, fscratch);                      // process.py:33
    if (ws.dbg_mel_db) {
        float* d = ws.dbg_mel_db + (size_t)b * NP;
        for (int i = threadIdx.x; i < NP; i += blockDim.x) d[i] = P[i];
    }
    float *o0, *o1, *o2;
    if (mel3) {
        o0 = mel3 + (size_t)b * 3 * NP; o1 = o0 + NP; o2 = o1 + NP;
    } else {
        o0 = plane_ptr(feats, b, BPC_CH_MEL, T);
        o1 = plane_ptr(feats, b, BPC_CH_MEL_DELTA, T);
        o2 = plane_ptr(feats, b, BPC_CH_MEL_DELTA2, T);
    }
    // statistics of mel_db, delta, delta2 (process.py:34-38).  The raw deltas are parked in their output planes and
    // normalised in place by the thread that wrote them (v1 evaluated the FP64 stencils twice).
    double s0 = 0, q0 = 0, s1, q1, s2, q2;
    {
        double sums[4];
        delta_rows_block(P, kPlaneRows, T, o1, o2, sums);
        s1 = sums[0]; q1 = sums[1]; s2 = sums[2]; q2 = sums[3];
    }
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const double v0 = (double)P[i];
        s0 += v0; q0 += v0 * v0;
    }
    block_sum2(s0, q0, dscratch);
    block_sum2(s1, q1, dscratch);
    block_sum2(s2, q2, dscratch);
    const ZTerm z0 = make_zterm(s0, q0, (double)NP), z1 = make_zterm(s1, q1, (double)NP),
                z2 = make_zterm(s2, q2, (double)NP);
    const bool stats = !LONG && !mel3 && ws.stats_acc != nullptr;
    StatAcc a0, a1, a2;
    a0.init(); a1.init(); a2.init();
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const float v0 = z0(P[i]), v1 = z1(o1[i]), v2 = z2(o2[i]);
        o0[i] = v0;
        o1[i] = v1;
        o2[i] = v2;
        if (stats) { a0.add(v0); a1.add(v1); a2.add(v2); }
    }
    if (stats) {
        stat_flush_block(a0, ws.stats_acc + 5 * BPC_CH_MEL, dscratch, fscratch);
        stat_flush_block(a1, ws.stats_acc + 5 * BPC_CH_MEL_DELTA, dscratch, fscratch);
        stat_flush_block(a2, ws.stats_acc + 5 * BPC_CH_MEL_DELTA2, dscratch, fscratch);
    }
    if (mel3) return;

    // mod_spec (methods.py:142-143): DCT-II ortho over mel (keep 40), then over time
    dct_mel40(tb.dct_mel, P, T, C1);
    // long mode: the [40 x T] . [T x T] time DCT is spread over a (column tile, segment) grid by the next two kernels
    if (LONG) return;
    dct_time40(DTs, C1, T, C2);
    if (ws.dbg_mod) {
        float* d = ws.dbg_mod + (size_t)b * 40 * T;
        for (int i = threadIdx.x; i < 40 * T; i += blockDim.x) d[i] = C2[i];
    }
    float mn;
    const ZTerm zm = zterm_of(C2, 40 * T, dscratch, fscratch, &mn);
    const float fill = zm(mn);                                         // pad_freq: min of the normalised array
    float* om = plane_ptr(feats, b, BPC_CH_MOD_SPEC, T);
    StatAcc am;
    am.init();
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const float v = (i < 40 * T) ? zm(C2[i]) : fill;
        om[i] = v;
        if (stats && i < 40 * T) am.add(v);
    }
    if (stats) {
        if (threadIdx.x == 0) am.add_n(fill, NP - 40 * T);
        stat_flush_block(am, ws.stats_acc + 5 * BPC_CH_MOD_SPEC, dscratch, fscratch);
    }
}

// ----------------------------------------------------------------------------------------------- role 1: mfcc
template <bool LONG>
__device__ void role_mfcc(int b, const Geometry g, const Tables& tb, const Workspace& ws, float* feats, float* smem,
                          double* dscratch, float* fscratch) {
    const int T = g.T, NP = kPlaneRows * T;
    float* P = smem;                 // [128*T]
    float* MF = P + NP;              // [40*T]
    float* OUT = MF + 40 * T;        // [120*T]
    const float* mag_b = ws.mag512 + (size_t)b * T * kMagStride;
    apply_bank<128>(tb.mel_b, mag_b, T, true, P);
    power_to_db_inplace(P, NP, false, fscratch);                      // librosa.feature.mfcc: power_to_db(ref=1.0)
    dct_mel40(tb.dct_mel, P, T, MF);
    if (ws.dbg_mfcc) {
        float* d = ws.dbg_mfcc + (size_t)b * 120 * T;
        for (int i = threadIdx.x; i < 120 * T; i += blockDim.x) {
            const int r = i / T, t = i - r * T;
            d[i] = r < 40 ? MF[i] : (r < 80 ? delta_at(MF + (r - 40) * T, t, T, 1) : delta_at(MF + (r - 80) * T, t, T, 2));
        }
    }
    // row-wise z-score of vstack([mfcc, delta, delta2]) (process.py:46-47): one warp per row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float mn = FLT_MAX;
    for (int i = threadIdx.x; i < 40 * T; i += blockDim.x) OUT[i] = MF[i];
    delta_rows_block(MF, 40, T, OUT + 40 * T, OUT + 80 * T, nullptr);
    __syncthreads();
    // (uniform trip count; a warp past the last row redoes row 119 without storing -- it may read that row while its
    // owner rewrites it, the result is dropped: the warp reductions inside np_row_zterm are then provably convergent,
    // plain SHFL instead of WARPSYNC.COLLECTIVE sequences)
    for (int r0 = 0; r0 < 120; r0 += nw) {
        const bool live = r0 + warp < 120;
        const int r = live ? r0 + warp : 119;
        const ZTerm z = np_row_zterm(OUT + r * T, T, lane);
        __syncwarp();
        if (live)
            for (int t = lane; t < T; t += 32) {
                const float v = z(OUT[r * T + t]);
                OUT[r * T + t] = v;
                mn = fminf(mn, v);
            }
    }
    mn = block_min(mn, fscratch);
    __syncthreads();
    float* o = plane_ptr(feats, b, BPC_CH_MFCC, T);
    const bool stats = !LONG && ws.stats_acc != nullptr;
    StatAcc am;
    am.init();
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const float v = (i < 120 * T) ? OUT[i] : mn;
        o[i] = v;
        if (stats && i < 120 * T) am.add(v);
    }
    if (stats) {
        if (threadIdx.x == 0) am.add_n(mn, NP - 120 * T);
        stat_flush_block(am, ws.stats_acc + 5 * BPC_CH_MFCC, dscratch, fscratch);
    }
}

// ------------------------------------------------------------------------------------------ role 2: gammatone
template <bool LONG>
__device__ void role_gammatone(int b, const Geometry g, const Tables& tb, const Workspace& ws, float* feats,
                               float* smem, double* dscratch, float* fscratch) {
    const int T = g.T, NP = kPlaneRows * T, NG = tb.mel_c.rows * T;
    float* G = smem;
    const float* mag_b = ws.mag512 + (size_t)b * T * kMagStride;
    apply_bank<64>(tb.mel_c, mag_b, T, false, G);
    for (int i = threadIdx.x; i < NG; i += blockDim.x) G[i] = log1pf(G[i]);
    __syncthreads();
    if (ws.dbg_gam) {
        float* d = ws.dbg_gam + (size_t)b * NG;
        for (int i = threadIdx.x; i < NG; i += blockDim.x) d[i] = G[i];
    }
    float mn;
    const ZTerm z = zterm_of(G, NG, dscratch, fscratch, &mn);
    const float fill = z(mn);
    float* o = plane_ptr(feats, b, BPC_CH_GAMMATONE, T);
    const bool stats = !LONG && ws.stats_acc != nullptr;
    StatAcc ag;
    ag.init();
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const float v = (i < NG) ? z(G[i]) : fill;
        o[i] = v;
        if (stats && i < NG) ag.add(v);
    }
    if (stats) {
        if (threadIdx.x == 0) ag.add_n(fill, NP - NG);
        stat_flush_block(ag, ws.stats_acc + 5 * BPC_CH_GAMMATONE, dscratch, fscratch);
    }
}

// ------------------------------------------------------------------------------------------ role 3: chroma_stft
template <bool LONG>
__device__ void role_chroma_stft(int b, const Geometry g, const Tables& tb, const Workspace& ws, float* feats,
                                 float* scalars, int32_t* status, float* smem, double* dscratch, float* fscratch) {
    const int T = g.T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    // candidate capacity: kMaxCand in shared memory, every (bin, frame) pair in long mode (global scratch)
    const int cap = LONG ? 123 * T : kMaxCand;
    float* cand_mag = smem;                       // [cap]
    float* cand_pitch = cand_mag + cap;           // [cap]
    float* sortbuf = cand_pitch + cap;            // [kSelectWords]
    float* colmax = sortbuf + kSelectWords;       // [T] (64 in shared memory)
    float* raw = colmax + (LONG ? T : 64); // [12*T]
    int* hist = (int*)(raw + 12 * T);             // [100]
    __shared__ int s_ncand;
    const float* mag_b = ws.mag512 + (size_t)b * T * kMagStride;
    if (tid == 0) s_ncand = 0;

    // column maxima + low-frequency energy ratio (methods.py:84-88: sum(|X|^2[:32]) / (sum(|X|^2) + 1e-8))
    double e_low = 0.0, e_tot = 0.0;
    for (int t0 = 0; t0 < T; t0 += nw) {              // uniform trip count: convergent warp reductions
        const bool live = t0 + warp < T;
        const int t = live ? t0 + warp : T - 1;
        float mx = 0.f;
        for (int k = lane; k < 257; k += 32) {
            const float v = __ldg(mag_b + (size_t)t * kMagStride + k);
            mx = fmaxf(mx, v);
            const double p = live ? (double)__fmul_rn(v, v) : 0.0;
            e_tot += p;
            if (k < 32) e_low += p;
        }
        mx = warp_max(mx);
        if (lane == 0 && live) colmax[t] = mx;
    }
    block_sum2(e_low, e_tot, dscratch);
    if (tid == 0) {
        const float lo = (float)e_low, tot = (float)e_tot;                  // np.sum of float32 arrays
        scalars[(size_t)b * g.nscal + 25] = __fdiv_rn(lo, __fadd_rn(tot, 1e-8f));
    }
    __syncthreads();

    // piptrack over bins 5..127 (150 Hz <= k * 31.25 < 4000 Hz), threshold 0.1 * column max
    const int nb = 123;
    for (int idx = tid; idx < nb * T; idx += blockDim.x) {
        const int t = idx / nb, k = 5 + idx - t * nb;
        const float* col = mag_b + (size_t)t * kMagStride;
        float pitch, mv;
        if (piptrack_candidate(__ldg(col + k - 1), __ldg(col + k), __ldg(col + k + 1), __fmul_rn(0.1f, colmax[t]), k,
                               31.25, &pitch, &mv)) {
            const int slot = atomicAdd(&s_ncand, 1);
            if (slot < cap) { cand_mag[slot] = mv; cand_pitch[slot] = pitch; }
        }
    }
    __syncthreads();
    int n = s_ncand;
    uint32_t flags = 0;
    if (n > cap) { n = cap; flags |= BPC_SEG_CAND_OVERFLOW; }
    bool empty = false;
    const int tbin = tuning_from_candidates(cand_mag, cand_pitch, n, sortbuf, hist, tb.hist_edges, 12, &empty);
    if (empty) flags |= BPC_SEG_TUNING_EMPTY;
    if (tid == 0) {
        ws.tuning[b * 2 + 0] = tbin;
        if (status && flags) atomicOr((unsigned int*)&status[b], flags);
    }

    // raw chroma = chromafb[tuning] @ |X| (float32), one warp per frame
    const float* fb = tb.chroma + (size_t)tbin * 12 * 257;
    for (int t0 = 0; t0 < T; t0 += nw) {              // uniform trip count: convergent warp reductions
        const bool live = t0 + warp < T;
        const int t = live ? t0 + warp : T - 1;
        float acc[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] = 0.f;
        for (int k = lane; k < 257; k += 32) {
            const float v = __ldg(mag_b + (size_t)t * kMagStride + k);
#pragma unroll
            for (int c = 0; c < 12; ++c) acc[c] = fmaf(__ldg(fb + c * 257 + k), v, acc[c]);
        }
        float cmax = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            acc[c] = warp_sum(acc[c]);
            cmax = fmaxf(cmax, fabsf(acc[c]));
        }
        // util.normalize(norm=inf): float64 division, float32 store; columns below tiny are left alone
        const double len = (cmax < 1.17549435e-38f) ? 1.0 : (double)cmax;
        if (lane < 12 && live) {
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < 12; ++c) if (c == lane) v = acc[c];
            raw[lane * T + t] = (float)((double)v / len);
        }
    }
    __syncthreads();
    if (ws.dbg_chroma_stft) {
        float* d = ws.dbg_chroma_stft + (size_t)b * 12 * T;
        for (int i = tid; i < 12 * T; i += blockDim.x) d[i] = raw[i];
    }
    // row-wise z-score (process.py:55), rows 0..11 of the chroma plane; the pad rows are filled by the CENS kernel
    float mn = FLT_MAX;
    float* o = plane_ptr(feats, b, BPC_CH_CHROMA, T);
    const bool stats = !LONG && ws.stats_acc != nullptr;
    StatAcc ac;
    ac.init();
    for (int r0 = 0; r0 < 12; r0 += nw) {
        const bool live = r0 + warp < 12;
        const int r = live ? r0 + warp : 11;
        const ZTerm z = np_row_zterm(raw + r * T, T, lane);
        if (live)
            for (int t = lane; t < T; t += 32) {
                const float v = z(raw[r * T + t]);
                o[r * T + t] = v;
                mn = fminf(mn, v);
                if (stats) ac.add(v);
            }
    }
    mn = block_min(mn, fscratch);
    if (tid == 0) ws.chroma_min[b * 2 + 0] = mn;
    if (stats) stat_flush_block(ac, ws.stats_acc + 5 * BPC_CH_CHROMA, dscratch, fscratch);   // rows 12..127: k_cens
}

// ----------------------------------------------------------------------------------------------- the kernel
// roles 0 / 1 (three CTAs per SM) and roles 2 / 3 (four per SM) are launched separately with their own footprints
constexpr int kConsumerSmemFloats = kPlaneRows * kMaxFrames + 40 * kMaxFrames + 120 * kMaxFrames;             // role_mfcc
static_assert(kPlaneRows * kMaxFrames + 80 * kMaxFrames + kMaxFrames * kMaxFrames <= kConsumerSmemFloats, "role_mel layout");
constexpr int kLightSmemFloats = 2 * kMaxCand + kSelectWords + 64 + 12 * kMaxFrames + 100 + 28;              // role_chroma_stft
static_assert(64 * kMaxFrames <= kLightSmemFloats, "role_gammatone layout");

// per-segment scratch floats of the four roles in long mode (each role owns a disjoint region)
__host__ __device__ inline size_t consumer_role_offset(int role, int T) {
    const size_t t = (size_t)T;
    const size_t o1 = 208 * t, o2 = o1 + 288 * t, o3 = o2 + 64 * t;
    return role == 0 ? 0 : (role == 1 ? o1 : (role == 2 ? o2 : o3));
}
size_t consumer_scratch_floats(int T) {
    return consumer_role_offset(3, T) + (size_t)2 * 123 * T + kSelectWords + (size_t)13 * T + 128;
}

// 1 s mode, roles 0 / 1: three CTAs per SM by shared memory; 288 threads (nine warps) is what 72 registers allow
constexpr int kHeavyThreads = 288;
template <bool LONG>
__global__ void __launch_bounds__(LONG ? 512 : kHeavyThreads, LONG ? 1 : 3) k_spec512_consumers(Geometry g, Tables tb, Workspace ws, float* feats,
                                                           float* scalars, int32_t* status, float* mel3,
                                                           int role_base) {
    extern __shared__ __align__(16) float smem_dyn[];
    __shared__ double dscratch[32];
    __shared__ float fscratch[32];
    const int b = blockIdx.x;
    const int role = blockIdx.y + role_base;
    // LONG is a template parameter so that the 1 s instantiation keeps provable shared-memory addressing (LDS / STS)
    float* smem = LONG ? ws.scratch + (size_t)b * ws.scratch_stride + consumer_role_offset(role, g.T) : smem_dyn;
    switch (role) {
        case 0: role_mel<LONG>(b, g, tb, ws, feats, mel3, smem, dscratch, fscratch); break;
        case 1: role_mfcc<LONG>(b, g, tb, ws, feats, smem, dscratch, fscratch); break;
        case 2: role_gammatone<LONG>(b, g, tb, ws, feats, smem, dscratch, fscratch); break;
        default: role_chroma_stft<LONG>(b, g, tb, ws, feats, scalars, status, smem, dscratch, fscratch); break;
    }
}

// long mode: C2[40, 64-column tile] = C1[40, T] . DT[T, tile]; C1 / C2 live in role 0's scratch region
__global__ void __launch_bounds__(256) k_modspec_time_long(Geometry g, Tables tb, Workspace ws) {
    const int b = blockIdx.y, T = g.T;
    float* base = ws.scratch + (size_t)b * ws.scratch_stride + consumer_role_offset(0, T);
    dct_time40_tile(tb.dct_time, base + (size_t)kPlaneRows * T, T, base + (size_t)(kPlaneRows + 40) * T, 64 * blockIdx.x);
}

__global__ void __launch_bounds__(256) k_modspec_finish_long(Geometry g, Workspace ws, float* feats) {
    __shared__ double dscratch[32];
    __shared__ float fscratch[32];
    const int b = blockIdx.x, T = g.T, NP = kPlaneRows * T;
    const float* C2 = ws.scratch + (size_t)b * ws.scratch_stride + consumer_role_offset(0, T) + (size_t)(kPlaneRows + 40) * T;
    if (ws.dbg_mod) {
        float* d = ws.dbg_mod + (size_t)b * 40 * T;
        for (int i = threadIdx.x; i < 40 * T; i += blockDim.x) d[i] = C2[i];
    }
    float mn;
    const ZTerm zm = zterm_of(C2, 40 * T, dscratch, fscratch, &mn);
    const float fill = zm(mn);                                         // pad_freq: min of the normalised array
    float* om = plane_ptr(feats, b, BPC_CH_MOD_SPEC, T);
    for (int i = threadIdx.x; i < NP; i += blockDim.x) om[i] = (i < 40 * T) ? zm(C2[i]) : fill;
}

// 1 s mode, roles 2 / 3 only (gammatone, chroma_stft + tuning): its own kernel so that the register tiles of the DCT
// roles do not set its register count (80 -> at most 64: four CTAs per SM instead of three)
__global__ void __launch_bounds__(256, 4) k_spec512_light(Geometry g, Tables tb, Workspace ws, float* feats, float* scalars,
                                                          int32_t* status) {
    extern __shared__ __align__(16) float smem_dyn[];
    __shared__ double dscratch[32];
    __shared__ float fscratch[32];
    const int b = blockIdx.x;
    if (blockIdx.y == 0) role_gammatone<false>(b, g, tb, ws, feats, smem_dyn, dscratch, fscratch);
    else role_chroma_stft<false>(b, g, tb, ws, feats, scalars, status, smem_dyn, dscratch, fscratch);
}

static void set_consumer_smem() {
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_spec512_consumers<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kConsumerSmemFloats * (int)sizeof(float));
    });
}

void launch_spec512_consumers(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                              float* scalars, int32_t* status, bool with_chroma, cudaStream_t st) {
    set_consumer_smem();
    static const char* only = std::getenv("BPC_ONLY_ROLE");      // profiling aid: time one role (outputs incomplete)
    if (only) {
        const int r = std::atoi(only);
        if (g.long_mode) k_spec512_consumers<true><<<dim3(n, 1), 512, 0, st>>>(g, tb, ws, feats, scalars, status, nullptr, r);
        else k_spec512_consumers<false><<<dim3(n, 1), 256, (r < 2 ? kConsumerSmemFloats : kLightSmemFloats) * sizeof(float), st>>>(
            g, tb, ws, feats, scalars, status, nullptr, r);
        note_launch();
        return;
    }
    if (g.long_mode) {
        // 512 threads: with one CTA per (segment, role) and <= 148 segments a launch, more loads in flight per SM
        k_spec512_consumers<true><<<dim3(n, 2), 512, 0, st>>>(g, tb, ws, feats, scalars, status, nullptr, 0);
        k_spec512_consumers<true><<<dim3(n, with_chroma ? 2 : 1), 512, 0, st>>>(g, tb, ws, feats, scalars, status, nullptr, 2);
    } else {
        k_spec512_consumers<false><<<dim3(n, 2), kHeavyThreads, kConsumerSmemFloats * sizeof(float), st>>>(
            g, tb, ws, feats, scalars, status, nullptr, 0);
        k_spec512_light<<<dim3(n, with_chroma ? 2 : 1), 256, kLightSmemFloats * sizeof(float), st>>>(g, tb, ws, feats, scalars,
                                                                                                    status);
    }
    note_launch(2);
    if (g.long_mode) {
        if (modspec_time_tc_enabled(tb)) {                     // tcgen05 3xTF32 (k_tc.cu); BPC_TC_DCT=0: FP32 SIMT tiles
            launch_modspec_time_tc(n, g, tb, ws, consumer_role_offset(0, g.T), st);
        } else {
            k_modspec_time_long<<<dim3((g.T + 63) / 64, n), 256, 0, st>>>(g, tb, ws);
            note_launch();
        }
        k_modspec_finish_long<<<n, 256, 0, st>>>(g, ws, feats);
        note_launch();
    }
}

void launch_logmel_only(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* mel3, cudaStream_t st) {
    set_consumer_smem();
    dim3 grid(n, 1);
    if (g.long_mode) k_spec512_consumers<true><<<grid, 512, 0, st>>>(g, tb, ws, nullptr, nullptr, nullptr, mel3, 0);
    else k_spec512_consumers<false><<<grid, kHeavyThreads, kConsumerSmemFloats * sizeof(float), st>>>(g, tb, ws, nullptr, nullptr,
                                                                                           nullptr, mel3, 0);
    note_launch();
}

// ------------------------------------------------------- config-2 stage output: power_to_db(|X|^2, ref=max) [257, T]
__global__ void __launch_bounds__(256) k_stft_db(Geometry g, const float* __restrict__ mag, float* __restrict__ out) {
    extern __shared__ __align__(16) float sP[];    // [257*T]
    __shared__ float fscratch[32];
    const int b = blockIdx.x, T = g.T;
    const float* mag_b = mag + (size_t)b * T * kMagStride;
    for (int idx = threadIdx.x; idx < 257 * T; idx += blockDim.x) {
        const int k = idx % 257, t = idx / 257;
        const float v = __ldg(mag_b + (size_t)t * kMagStride + k);
        sP[k * T + t] = __fmul_rn(v, v);
    }
    __syncthreads();
    power_to_db_inplace(sP, 257 * T, true, fscratch);
    float* o = out + (size_t)b * 257 * T;
    for (int i = threadIdx.x; i < 257 * T; i += blockDim.x) o[i] = sP[i];
}

// long mode: the [257, T] tile does not fit on chip; two sweeps over the magnitudes (L2-resident) instead.  Same float32
// operations as power_to_db_inplace (the clamp floor follows from the maximum power because 10 log10 is monotone).
__global__ void __launch_bounds__(256) k_stft_db_long(Geometry g, const float* __restrict__ mag, float* __restrict__ out) {
    __shared__ float fscratch[32];
    const int b = blockIdx.x, T = g.T;
    const float* mag_b = mag + (size_t)b * T * kMagStride;
    float mx = -FLT_MAX;
    for (int idx = threadIdx.x; idx < 257 * T; idx += blockDim.x) {
        const int t = idx / 257, k = idx - t * 257;
        const float v = __ldg(mag_b + (size_t)t * kMagStride + k);
        mx = fmaxf(mx, __fmul_rn(v, v));
    }
    mx = block_max(mx, fscratch);
    const float ref_db = (float)(10.0 * log10((double)fmaxf(1e-10f, mx)));
    const float vmax = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(1e-10f, mx))), ref_db);
    const float floor_db = __fsub_rn(vmax, 80.0f);
    float* o = out + (size_t)b * 257 * T;
    for (int i = threadIdx.x; i < 257 * T; i += blockDim.x) {
        const int k = i / T, t = i - k * T;
        const float v = __ldg(mag_b + (size_t)t * kMagStride + k);
        o[i] = fmaxf(__fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(1e-10f, __fmul_rn(v, v)))), ref_db), floor_db);
    }
}

void launch_stft_db(int n, const Geometry& g, const Workspace& ws, float* stft_db, cudaStream_t st) {
    if (g.long_mode) {
        k_stft_db_long<<<n, 256, 0, st>>>(g, ws.mag512, stft_db);
        note_launch();
        return;
    }
    const int bytes = 257 * g.T * (int)sizeof(float);
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_stft_db, cudaFuncAttributeMaxDynamicSharedMemorySize, 257 * kMaxFrames * 4);
    });
    k_stft_db<<<n, 256, bytes, st>>>(g, ws.mag512, stft_db);
    note_launch();
}

// ------------------------------------------- methods.py:142-143 on caller-provided mel_db ([n, 128, T] -> [n, 40, T])
__global__ void __launch_bounds__(256) k_modspec(Geometry g, Tables tb, const float* __restrict__ mel_db,
                                                 float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int T = g.T, NP = kPlaneRows * T, b = blockIdx.x;
    float* P = smem;
    float* C1 = P + NP;
    float* C2 = C1 + 40 * T;
    float* DTs = C2 + 40 * T;
    stage_matrix(DTs, tb.dct_time, T * T);
    for (int i = threadIdx.x; i < NP; i += blockDim.x) P[i] = mel_db[(size_t)b * NP + i];
    __syncthreads();
    dct_mel40(tb.dct_mel, P, T, C1);
    dct_time40(DTs, C1, T, C2);
    for (int i = threadIdx.x; i < 40 * T; i += blockDim.x) out[(size_t)b * 40 * T + i] = C2[i];
}

// long mode: mel_db -> role 0's scratch region -> mel DCT there; the tiled time DCT follows; then C2 is copied out
__global__ void __launch_bounds__(256) k_modspec_long_in(Geometry g, Tables tb, Workspace ws, const float* __restrict__ mel_db) {
    const int T = g.T, NP = kPlaneRows * T, b = blockIdx.x;
    float* P = ws.scratch + (size_t)b * ws.scratch_stride + consumer_role_offset(0, T);
    for (int i = threadIdx.x; i < NP; i += blockDim.x) P[i] = mel_db[(size_t)b * NP + i];
    __syncthreads();
    dct_mel40(tb.dct_mel, P, T, P + NP);
}
__global__ void __launch_bounds__(256) k_modspec_long_out(Geometry g, Workspace ws, float* __restrict__ out) {
    const int T = g.T, b = blockIdx.x;
    const float* C2 = ws.scratch + (size_t)b * ws.scratch_stride + consumer_role_offset(0, T) + (size_t)(kPlaneRows + 40) * T;
    for (int i = threadIdx.x; i < 40 * T; i += blockDim.x) out[(size_t)b * 40 * T + i] = C2[i];
}

void launch_modspec(int n, const Geometry& g, const Tables& tb, const Workspace& ws, const float* mel_db, float* out,
                    cudaStream_t st) {
    if (g.long_mode) {
        for (int off = 0; off < n; off += ws.cap) {                    // the scratch region holds ws.cap segments
            const int m = n - off < ws.cap ? n - off : ws.cap;
            const size_t NP = (size_t)kPlaneRows * g.T;
            k_modspec_long_in<<<m, 256, 0, st>>>(g, tb, ws, mel_db + off * NP);
            if (modspec_time_tc_enabled(tb)) launch_modspec_time_tc(m, g, tb, ws, consumer_role_offset(0, g.T), st);
            else k_modspec_time_long<<<dim3((g.T + 63) / 64, m), 256, 0, st>>>(g, tb, ws);
            k_modspec_long_out<<<m, 256, 0, st>>>(g, ws, out + (size_t)off * 40 * g.T);
            note_launch(3);
        }
        return;
    }
    set_consumer_smem();
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_modspec, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kConsumerSmemFloats * (int)sizeof(float));
    });
    k_modspec<<<n, 256, kConsumerSmemFloats * sizeof(float), st>>>(g, tb, mel_db, out);
    note_launch();
}

// ================================================================= BASELINE config 2, fused (1 s mode): k_logmel_fused
// log-power STFT [257, T] + mel / mel_delta / mel_delta2 [3, 128, T] of a segment in ONE kernel: nothing but the input
// (32 KB of PCM16, or 64 KB of float32) is read from HBM and nothing but the two outputs is written (225,532 B per
// segment with float32 input: SURVEY 8(d) config 2).  The three-kernel path it replaces moved the |X| workspace
// through L2 / HBM twice more (k_stft512 -> k_stft_db + k_spec512_consumers role 0).
//
// Persistent: one 512-thread CTA per SM walks segments b = blockIdx.x, + gridDim.x, ...
//   stage   the raw segment lands in shared memory by ONE bulk-TMA copy per <= 32 KB (cp.async.bulk + mbarrier,
//           SASS: UBLKCP) between two permanent 256-sample zero pads (librosa's centre padding: no bounds checks in
//           the frame loader); the copy for segment b + gridDim.x is issued as soon as the FFT phase of segment b has
//           consumed the buffer, so it overlaps the dB / filterbank / stencil / store phases
//   FFT     32 half-warp teams x 2 rounds: Hann * samples -> FP64 real FFT-512 (fft_reg.cuh, split exchange) ->
//           |X|^2 as float32 into a [T][257] tile (stride 257: conflict-free for the frame-major writes here and the
//           bin-major reads below); PCM16 samples become doubles by the 2^52 trick (one XOR + one DADD instead of two
//           quarter-rate conversions) and their 2^-15 is folded, exactly, into the final complex64 rounding
//   dB      power_to_db(ref = max, top_db = 80) from the tile, written once, 16-byte stores
//   mel     band-form mel-A filterbank from the tile -> [128][T] (in the FFT exchange space), dB, Savitzky-Golay
//           deltas into the (now dead) power tile, three whole-array z-scores, each plane written once
constexpr int kLmThreads = 512;
constexpr int kLmTeams = kLmThreads / 16;
constexpr int kLmPStride = 257;
constexpr int kLmPad = 256;                                   // n_fft / 2 zero samples on either side
constexpr int kLmXchDoubles = 16 * 17;                        // split exchange: one component at a time
constexpr int kLmRawBytesPcm = (2 * kLmPad + 16000) * 2, kLmRawBytesF32 = (2 * kLmPad + 16000) * 4;
constexpr int kLmTileBytes = ((63 * kLmPStride * 4 + 15) / 16) * 16;
constexpr int kLmXchBytes = kLmTeams * kLmXchDoubles * 8;
static_assert(kLmXchBytes >= kPlaneRows * 63 * 4, "the mel tile reuses the exchange space");
static_assert(kLmTileBytes >= 2 * kPlaneRows * 63 * 4, "the two delta tiles reuse the power tile");

template <bool PCM>
__global__ void __launch_bounds__(kLmThreads, 1) k_logmel_fused(const void* __restrict__ wav, int n, Geometry g, Tables tb,
                                                               float* __restrict__ stft_db, float* __restrict__ mel3) {
    extern __shared__ __align__(16) unsigned char lm_smem[];
    __shared__ double dscratch[32];
    __shared__ float fscratch[32];
    __shared__ __align__(8) uint64_t bar;
    constexpr int kRawBytes = PCM ? kLmRawBytesPcm : kLmRawBytesF32;
    constexpr int kSampleBytes = PCM ? 2 : 4;
    unsigned char* raw = lm_smem;
    float* tile = reinterpret_cast<float*>(lm_smem + kRawBytes);                       // [T][257] |X|^2, later d1 | d2
    double* xch_all = reinterpret_cast<double*>(lm_smem + kRawBytes + kLmTileBytes);   // exchange, later mel [128][T]
    float* melP = reinterpret_cast<float*>(xch_all);
    const int tid = threadIdx.x, lane = tid & 31, h = lane & 15, team = tid >> 4;
    const int partner = (lane & 16) | ((16 - h) & 15);
    const int T = 63, L = 16000, NP = kPlaneRows * T;
    const uint32_t seg_bytes = (uint32_t)L * kSampleBytes;

    // permanent zero pads; barrier; first copy
    for (int i = tid; i < kLmPad * kSampleBytes / 4; i += kLmThreads) {
        reinterpret_cast<uint32_t*>(raw)[i] = 0u;
        reinterpret_cast<uint32_t*>(raw + kLmPad * kSampleBytes + seg_bytes)[i] = 0u;
    }
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int b) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of the buffer precede the async write
        mbar_expect_tx(&bar, seg_bytes);
        const unsigned char* src = static_cast<const unsigned char*>(wav) + (size_t)b * seg_bytes;
        for (uint32_t off = 0; off < seg_bytes; off += 32000u) {
            const uint32_t len = seg_bytes - off < 32000u ? seg_bytes - off : 32000u;
            tma_bulk_g2s(raw + kLmPad * kSampleBytes + off, src + off, len, &bar);
        }
    };
    if (tid == 0 && (int)blockIdx.x < n) issue(blockIdx.x);

    // inter-stage twiddles from the [k1][h] table (L1-resident, 4 KB): registers are the 512-thread CTA's scarce resource
    const double2* twa = tb.twa256 + h;
    const double2 wp = __ldg(tb.ptw512 + h);
    const double2 wl = make_double2(wp.y, -wp.x);                         // -i * exp(-2 pi i h / 512)
    const double2* win2 = reinterpret_cast<const double2*>(tb.hann512);
    double* xr = xch_all + (size_t)team * kLmXchDoubles;
    // 2 X[k] comes out of the split; complex64 = float32 rounding of X[k] (x 2^-15 for PCM16 samples: exact)
    const float out_scale = PCM ? 0.5f / 32768.0f : 0.5f;
    uint32_t parity = 0;

    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        mbar_wait(&bar, parity);
        parity ^= 1u;
        // ---- FFT phase
        float pmax = 0.f;
        for (int t0 = 0; t0 < T; t0 += kLmTeams) {                        // uniform trip count: every shuffle is convergent
            const int t = t0 + team;
            const bool valid = t < T;
            const int tt = valid ? t : T - 1;
            double2 a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int m = h + 16 * j;
                const double2 w = __ldg(win2 + m);
                double x0, x1;
                if (PCM) {
                    const uint32_t pr = *reinterpret_cast<const uint32_t*>(raw + ((size_t)tt * 256 + 2 * m) * 2);
                    const int s0 = (int)(short)(pr & 0xffffu), s1 = (int)pr >> 16;
                    x0 = __hiloint2double(0x43300000, s0 ^ (int)0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
                    x1 = __hiloint2double(0x43300000, s1 ^ (int)0x80000000) - 4503601774854144.0;
                } else {
                    const float2 v = *reinterpret_cast<const float2*>(raw + ((size_t)tt * 256 + 2 * m) * 4);
                    x0 = (double)v.x;
                    x1 = (double)v.y;
                }
                a[j] = make_double2(x0 * w.x, x1 * w.y);
            }
            team_fft_split<16>(a, twa, 16, xr, h);
            float* row = tile + tt * kLmPStride;
            // every conjugate pair once (fft_reg.cuh::team_rsplit_pairs); bin 256 = X[N] is real: |.| without the hypot
            float xlo = 1.f, xhi = 1.f, pmax_row = 0.f;
            auto emit = [&](int k, double2 t2) {
                const float re = out_scale * (float)t2.x, im = out_scale * (float)t2.y;
                float x;
                float v = c64_abs_f32_unchecked(re, im, &x);             // range checked once per row (fft.cuh)
                if (k == 256) { v = fabsf(re); x = 1.f; }
                xlo = fminf(xlo, x);
                xhi = fmaxf(xhi, x);
                const float p = __fmul_rn(v, v);
                if (valid) { row[k] = p; pmax_row = fmaxf(pmax_row, p); }
            };
            team_rsplit_pairs<16, 0>(a, wl, h, partner, emit);
            if (__any_sync(0xffffffffu, !(c64_abs_in_range(xlo) && c64_abs_in_range(xhi)))) {
                pmax_row = 0.f;
                auto emit_exact = [&](int k, double2 t2) {
                    const float re = out_scale * (float)t2.x, im = out_scale * (float)t2.y;
                    const float v = k == 256 ? fabsf(re) : c64_abs_exact(re, im);
                    const float p = __fmul_rn(v, v);
                    if (valid) { row[k] = p; pmax_row = fmaxf(pmax_row, p); }
                };
                team_rsplit_pairs<16, 0>(a, wl, h, partner, emit_exact);
            }
            pmax = fmaxf(pmax, pmax_row);
        }
        const float mx = block_max(pmax, fscratch);               // has the barrier that ends the FFT phase
        __syncthreads();
        if (tid == 0 && b + (int)gridDim.x < n) issue(b + gridDim.x);   // the raw buffer is free: prefetch the next segment

        // ---- log-power STFT (process.py:32-33 semantics of power_to_db: ref = max, amin 1e-10, top_db 80)
        if (stft_db) {
            const float ref_db = (float)(10.0 * log10((double)fmaxf(1e-10f, mx)));
            const float vmax = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(1e-10f, mx))), ref_db);
            const float floor_db = __fsub_rn(vmax, 80.0f);
            float* o = stft_db + (size_t)b * 257 * T;
            auto db_at = [&](int i) {
                const int k = i / T, t = i - k * T;
                return fmaxf(__fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(1e-10f, tile[t * kLmPStride + k]))), ref_db), floor_db);
            };
            const int total = 257 * T;
            const int head = (int)((4 - ((reinterpret_cast<uintptr_t>(o) >> 2) & 3)) & 3);      // floats up to 16-byte alignment
            const int body4 = (total - head) >> 2;
            if (tid < head) o[tid] = db_at(tid);
            for (int v = tid; v < body4; v += kLmThreads) {
                const int i = head + 4 * v;
                __stcs(reinterpret_cast<float4*>(o + i), make_float4(db_at(i), db_at(i + 1), db_at(i + 2), db_at(i + 3)));
            }
            for (int i = head + 4 * body4 + tid; i < total; i += kLmThreads) o[i] = db_at(i);
        }

        // ---- mel-A filterbank on the power tile (apply_bank<128> with the tile as its source: same taps, same order)
        {
            const BankDev& bank = tb.mel_a;
            const int half = (T + 1) >> 1, total = kPlaneRows * half;
            for (int idx = tid; idx < total; idx += kLmThreads) {
                const int m = idx & (kPlaneRows - 1), t0 = idx / kPlaneRows, t1 = t0 + half;
                const bool live1 = t1 < T;
                const int s0 = __ldg(bank.start + m), c = __ldg(bank.count + m);
                const int cmax = __reduce_max_sync(0xffffffffu, c);
                const float* src0 = tile + t0 * kLmPStride;
                const float* src1 = tile + (live1 ? t1 : t0) * kLmPStride;
                const float* wt = bank.wt + m;
                float acc0 = 0.f, acc1 = 0.f;
                for (int j = 0; j < cmax; ++j) {
                    const int k = min(s0 + j, 256);
                    const float w = __ldg(wt + j * kPlaneRows);
                    acc0 = fmaf(w, src0[k], acc0);
                    acc1 = fmaf(w, src1[k], acc1);
                }
                melP[m * T + t0] = acc0;
                if (live1) melP[m * T + t1] = acc1;
            }
        }
        __syncthreads();                                           // mel tile complete; every reader of the power tile is done
        power_to_db_inplace(melP, NP, true, fscratch);            // process.py:33
        float* d1s = tile;                                         // the power tile is dead: park the raw deltas there
        float* d2s = tile + NP;
        double s0 = 0, q0 = 0, s1, q1, s2, q2;
        {
            double sums[4];
            delta_rows_block(melP, kPlaneRows, T, d1s, d2s, sums);
            s1 = sums[0]; q1 = sums[1]; s2 = sums[2]; q2 = sums[3];
        }
        for (int i = tid; i < NP; i += kLmThreads) {
            const double v0 = (double)melP[i];
            s0 += v0; q0 += v0 * v0;
        }
        block_sum2(s0, q0, dscratch);
        block_sum2(s1, q1, dscratch);
        block_sum2(s2, q2, dscratch);
        const ZTerm z0 = make_zterm(s0, q0, (double)NP), z1 = make_zterm(s1, q1, (double)NP),
                    z2 = make_zterm(s2, q2, (double)NP);
        float4* o0 = reinterpret_cast<float4*>(mel3 + (size_t)b * 3 * NP);
        float4* o1 = o0 + NP / 4;
        float4* o2 = o1 + NP / 4;
        for (int v = tid; v < NP / 4; v += kLmThreads) {           // every thread reads back what it wrote above or what
            const float4 a = reinterpret_cast<const float4*>(melP)[v];     // the block_sum2 barriers have published
            const float4 c1 = reinterpret_cast<const float4*>(d1s)[v], c2 = reinterpret_cast<const float4*>(d2s)[v];
            __stcs(o0 + v, make_float4(z0(a.x), z0(a.y), z0(a.z), z0(a.w)));
            __stcs(o1 + v, make_float4(z1(c1.x), z1(c1.y), z1(c1.z), z1(c1.w)));
            __stcs(o2 + v, make_float4(z2(c2.x), z2(c2.y), z2(c2.z), z2(c2.w)));
        }
        __syncthreads();                                           // tiles are rewritten by the next segment's FFT phase
    }
}

bool launch_logmel_fused(const void* wav, int wav_dtype, int n, const Geometry& g, const Tables& tb, float* stft_db,
                         float* mel3, cudaStream_t st) {
    static const char* env = std::getenv("BPC_FUSED_LOGMEL");
    if (g.long_mode || g.T != 63 || g.L != 16000 || (env && env[0] == '0')) return false;
    if ((reinterpret_cast<uintptr_t>(wav) | reinterpret_cast<uintptr_t>(mel3)) & 15) return false;
    static PerDeviceOnce once;
    static int sms = 148;
    const int bytes_pcm = kLmRawBytesPcm + kLmTileBytes + kLmXchBytes, bytes_f32 = kLmRawBytesF32 + kLmTileBytes + kLmXchBytes;
    once.run([&] {
        cudaFuncSetAttribute(k_logmel_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_pcm);
        cudaFuncSetAttribute(k_logmel_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_f32);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    });
    const int grid = n < sms ? n : sms;
    if (grid <= 0) return true;
    if (wav_dtype == BPC_WAV_PCM16) k_logmel_fused<true><<<grid, kLmThreads, bytes_pcm, st>>>(wav, n, g, tb, stft_db, mel3);
    else k_logmel_fused<false><<<grid, kLmThreads, bytes_f32, st>>>(wav, n, g, tb, stft_db, mel3);
    note_launch();
    return true;
}

}  // namespace bpc
