// STFT-2048 group (librosa default n_fft, reached through methods.py:59-63,90 and process.py:74-75), r01 v1 layout:
//   k_frame2048 one WARP per frame: Hann * samples -> FP64 real FFT-2048 with register-resident radix-32 stages -> |X| row
//               in shared memory -> spectral centroid / bandwidth / flatness / contrast order statistics and the mel-D
//               power column from that row (v1 split this in two kernels around a 259 KB/segment HBM workspace)
//   k_even2048  the hop-512 frames (= even hop-256 frames): spectral_rolloff with numpy's sequential float32 cumsum and
//               the 36-bins-per-octave tuning estimate chroma_cens needs
//   k_seg2048   one CTA per segment: statistics of the per-frame features, mel-D dB -> flux + onset envelope ->
//               tempogram plane
// (v0 was one 256-thread CTA per segment doing all of this behind CTA-wide barriers: 36 % of the step, FP64 pipe 9 %,
//  47 % of its shared wavefronts bank conflicts -- profiles/r01_*.)
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "kernels.cuh"
#include "fft.cuh"
#include "fft_reg.cuh"
#include "tuning.cuh"

namespace bpc {

constexpr int kTempoLags = 384;
constexpr int kFrameFeat = 20;               // doubles per frame: cent, bw, flat, peak[7], valley[7] (+3 pad)

// ================================================================================================ k_feat2048
// spectral_contrast sub-bands (librosa, fmin=200, n_bands=6, sr=16000, n_fft=2048): first bin, length, order count
// Batcher odd-even merge sort network on C register-resident floats (ascending), written for the power of two P >= C
// with the elements C .. P-1 taken as +inf: every compare-exchange that touches them is a no-op and is dropped.
template <int P, int C>
__device__ __forceinline__ void sort_net_asc(float (&a)[C]) {
#pragma unroll
    for (int p = 1; p < P; p <<= 1)
#pragma unroll
        for (int k = p; k >= 1; k >>= 1)
#pragma unroll
            for (int j = k % p; j + k < P; j += 2 * k)
#pragma unroll
                for (int i = 0; i < k; ++i)
                    if (i + j + k < C && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        const float lo = fminf(a[i + j], a[i + j + k]), hi = fmaxf(a[i + j], a[i + j + k]);
                        a[i + j] = lo;
                        a[i + j + k] = hi;
                    }
}

// Mean of the N smallest and of the N largest of row[lo .. lo+len) by one warp; C = ceil(len / 32) elements per lane.
// Every lane sorts its own elements once (registers); the N extraction rounds then only compare the lanes' current
// heads (one REDUX + one ballot per round).  Ties carry equal values, so which of them is taken does not matter.
// (v1 rescanned all 13 elements per lane in every round: 45 % of the executed instructions of this kernel.)
template <int P, int C, int N>
__device__ __forceinline__ void warp_band_extremes(const float* row, int lo, int len, int lane, double* valley,
                                                   double* peak) {
    static_assert(C > N, "every lane must hold more elements than are extracted");
    float s[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const int j = lane + 32 * i;
        s[i] = j < len ? row[lo + j] : __int_as_float(0x7f800000);       // +inf marks "absent" (only the last row)
    }
    sort_net_asc<P, C>(s);
    const bool full = lane + 32 * (C - 1) < len;
    unsigned L[N], H[N];                                                 // magnitudes are >= 0: bit patterns order like the values
#pragma unroll
    for (int r = 0; r < N; ++r) {
        L[r] = __float_as_uint(s[r]);
        H[r] = __float_as_uint(full ? s[C - 1 - r] : s[C - 2 - r]);
    }
    double sum_lo = 0.0, sum_hi = 0.0;
#pragma unroll 1                     // code size: the loop body of k_frame2048 must stream through the instruction caches
    for (int r = 0; r < N; ++r) {
        const unsigned mlo = __reduce_min_sync(0xffffffffu, L[0]);
        const unsigned blo = __ballot_sync(0xffffffffu, L[0] == mlo);
        if (lane == __ffs(blo) - 1) {
#pragma unroll
            for (int q = 0; q + 1 < N; ++q) L[q] = L[q + 1];
            L[N - 1] = 0x7f800000u;
        }
        sum_lo += (double)__uint_as_float(mlo);
        const unsigned mhi = __reduce_max_sync(0xffffffffu, H[0]);
        const unsigned bhi = __ballot_sync(0xffffffffu, H[0] == mhi);
        if (lane == __ffs(bhi) - 1) {
#pragma unroll
            for (int q = 0; q + 1 < N; ++q) H[q] = H[q + 1];
            H[N - 1] = 0u;
        }
        sum_hi += (double)__uint_as_float(mhi);
    }
    *peak = (double)(float)(sum_hi / (double)N);         // np.mean of a float32 slice -> float32
    *valley = (double)(float)(sum_lo / (double)N);
}

// bands whose order count is 1: plain min / max
__device__ __forceinline__ void warp_band_minmax(const float* row, int lo, int len, int lane, double* valley,
                                                 double* peak) {
    unsigned mn = 0x7f800000u, mx = 0u;
    for (int j = lane; j < len; j += 32) {
        const unsigned v = __float_as_uint(row[lo + j]);
        mn = min(mn, v);
        mx = max(mx, v);
    }
    *valley = (double)__uint_as_float(__reduce_min_sync(0xffffffffu, mn));
    *peak = (double)__uint_as_float(__reduce_max_sync(0xffffffffu, mx));
}

// ================================================================================================ k_frame2048
// One WARP (= one 32-thread CTA) per frame: Hann * samples -> team_fft<32> (fft_reg.cuh: 32 lanes x 32 register-resident
// complex points, one shared-memory exchange) -> real split -> |X| row (float32, 1025 bins) in the same shared memory ->
// every per-frame consumer of that row.  Nothing but the per-frame results leaves the SM; the even (hop-512) frames
// additionally store their row for k_even2048.
// r02 rewrite (r01: 7200 warp instructions per frame, 28 % of them FP64, 168 registers with spills):
//   * a CTA is ONE warp: the frame index depends on blockIdx only, so the compiler knows that every shuffle / barrier
//     is convergent (the four-warp CTAs compiled each of them as WARPSYNC.COLLECTIVE + ENDCOLLECTIVE: 630 of 8200
//     instructions of the loop body);
//   * decimation-in-time register DFTs with six-FMA butterflies (DitR), the real split once per conjugate pair with
//     the twiddle factored the same way (ten instead of twenty FP64 instructions per bin pair, table rs2048), the
//     factor 1/2 of the split folded into the window table;
//   * spectral moments with compile-time weights (sum m, sum i m, sum i^2 m per lane; k = lane + 32 i is put back
//     once per frame).
// Instruction fetch: the loop body is straight-line SASS that only runs at speed while the warps of an SM walk it in
// lock-step (every frame costs the same, so they do).  r01 v30-v32 moved the hop-512 work in here: the even frames then
// took longer than the odd ones, the warps drifted apart, and 11 of 12 issue slots went to "no instruction" stalls.
constexpr int kF2RowBytes = 32 * 33 * 8;             // exchange buffer (one component at a time), later the |X| row (1028 floats)
// 168 registers = 12 warps per SM, no spills.  Measured alternatives: 128 registers (16 warps, 75 spilled doubles per
// frame) 2.78 vs 2.71 ms; 144 registers (14 warps) 4.74 vs 2.58 ms; 8 warps per SM at 168 registers (a third of the
// register file left to the kernels of the side streams) 2.76 vs 2.43 ms for the kernel and 11.42 vs 11.18 ms per step.
constexpr int kF2WarpsPerSm = 12;
// WARPS = 1 (default): one-warp CTAs, 12 per SM, each walking its own contiguous range of frames (consecutive frames of
// a warp overlap by 7 / 8 of their samples).  WARPS = 12 (BPC_F2_WARPS=12): one CTA per SM whose warps take CONSECUTIVE
// frames of a contiguous range: twelve consecutive frames span 2048 + 11 * 256 samples, so all but 12 * 256 samples
// of a round are L1 hits.  Measured 2.375 ms (1) against 2.427 ms (12): the sample loads are not what the warps wait
// for, and independent warps drift into different phases of the frame, which fills the pipes better than twelve warps
// that start every round together.  The same reason sank the r02 attempt to stage a round's sample window by bulk TMA
// (one cp.async.bulk into a double-buffered shared window + mbarrier, every warp reading its frame with LDS): correct,
// UBLKCP in the SASS, but the CTA-wide barrier per round that frees the window put all warps into the same phase --
// 2.87 ms.  The trip count depends on blockIdx and kernel arguments only and warps past the end of the range redo its
// last frame without storing, so the collectives stay convergent (no WARPSYNC.COLLECTIVE) in the multi-warp CTA too.
template <int WARPS>
__device__ __forceinline__ void frame2048_body(const float* __restrict__ y, const Geometry& g, const Tables& tb,
                                               const Workspace& ws, int total_frames, int frames_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* smem_raw = smem_dyn + (size_t)warp * kF2RowBytes;
    double* xch = reinterpret_cast<double*>(smem_raw);
    float* row = reinterpret_cast<float*>(smem_raw);
    const int T = g.T, L = g.L, hop = g.hop, TE = (T + 1) / 2;
    const int partner = (32 - lane) & 31;
    const double2* win2 = reinterpret_cast<const double2*>(tb.hann2048h);   // 0.5 * Hann: the 1/2 of the real split
    const double2* twa = tb.twa1024 + lane;                               // [k1][lane]
    const double2* rs = tb.rs2048 + lane;                                 // split twiddles as (scale, tan / cot)
    const int f_lo = blockIdx.x * frames_per_cta;
    const int f_hi = f_lo + frames_per_cta < total_frames ? f_lo + frames_per_cta : total_frames;
    for (int f0 = f_lo; f0 < f_hi; f0 += WARPS) {
        const bool live = f0 + warp < f_hi;
        const int f = live ? f0 + warp : f_hi - 1;
        const int b = f / T, t = f - b * T;
        const float* yb = y + (size_t)b * L;
        const int g0 = t * hop - 1024;
        {
            double2 a[32];
            if (g0 >= 0 && g0 + 2048 <= L) {                   // interior frame (55 of 63): no bounds checks, one base pointer
                const float2* src = reinterpret_cast<const float2*>(yb + g0) + lane;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 v = __ldg(src + 32 * j);
                    const double2 w = __ldg(win2 + lane + 32 * j);
                    a[j] = make_double2((double)v.x * w.x, (double)v.y * w.y);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int m = lane + 32 * j;
                    const int gi = g0 + 2 * m;                 // even; L is even, so the pair is in or out together
                    float2 v = make_float2(0.f, 0.f);
                    if (gi >= 0 && gi < L) v = __ldg(reinterpret_cast<const float2*>(yb + gi));
                    const double2 w = __ldg(win2 + m);
                    a[j] = make_double2((double)v.x * w.x, (double)v.y * w.y);
                }
            }
            team_fft_split<32>(a, twa, 32, xch, lane);
            // real split, every conjugate pair once: with Z = FFT_1024 of the packed frame (already halved),
            //   s = Z[k] + conj(Z[N-k]),  d = Z[k] - conj(Z[N-k]),  p = w d,  w = -i exp(-2 pi i k / 2048),
            //   X[k] = s + p,  X[N-k] = conj(s - p);   w = c (t + i) for k < 256 and c (1 + i t) for k >= 256.
            const double z0re = a[0].x, z0im = a[0].y;
            // EXACT = false: float32-pair magnitudes without a per-bin range check, the range of max(|re|, |im|) is
            // tracked over the row; EXACT = true (the rare redo, e.g. an all-zero frame): the double-precision form.
            auto split_row = [&](auto exact_tag, float& xlo, float& xhi) {
                constexpr bool EXACT = decltype(exact_tag)::value;
#pragma unroll
                for (int K2 = 0; K2 <= 16; ++K2) {
                    const double2 zk = a[bitrev<32>(K2)];
                    double2 zn;
                    if (K2 < 16) {
                        const double2 src = a[bitrev<32>(31 - K2)];
                        zn.x = __shfl_sync(0xffffffffu, src.x, partner);
                        zn.y = __shfl_sync(0xffffffffu, src.y, partner);
                        if (lane == 0) zn = a[bitrev<32>((32 - K2) & 31)];
                    } else {
                        zn = zk;                                   // k = 512 (lane 0) pairs with itself
                    }
                    const double2 ct = __ldg(rs + 32 * K2);
                    const double sx = zk.x + zn.x, sy = zk.y - zn.y, dx = zk.x - zn.x, dy = zk.y + zn.y;
                    double qx, qy;
                    if (K2 < 8) { qx = fma(ct.y, dx, -dy); qy = fma(ct.y, dy, dx); }
                    else        { qx = fma(-ct.y, dy, dx); qy = fma(ct.y, dx, dy); }
                    const double x0 = fma(ct.x, qx, sx), y0 = fma(ct.x, qy, sy);
                    const double x1 = fma(-ct.x, qx, sx), y1 = fma(ct.x, qy, -sy);
                    const int k = lane + 32 * K2;
                    float m0, m1, xa = 1.f, xb = 1.f;
                    if (EXACT) {
                        m0 = c64_abs_exact((float)x0, (float)y0);
                        m1 = c64_abs_exact((float)x1, (float)y1);
                    } else {
                        m0 = c64_abs_f32_unchecked((float)x0, (float)y0, &xa);
                        m1 = c64_abs_f32_unchecked((float)x1, (float)y1, &xb);
                    }
                    if (K2 < 16) {
                        row[k] = m0;
                        xlo = fminf(xlo, xa);
                        xhi = fmaxf(xhi, xa);
                        if (K2 > 0 || lane > 0) {
                            row[1024 - k] = m1;
                            xlo = fminf(xlo, xb);
                            xhi = fmaxf(xhi, xb);
                        }
                    } else if (lane == 0) {
                        row[512] = m0;
                        xlo = fminf(xlo, xa);
                        xhi = fmaxf(xhi, xa);
                    }
                }
            };
            float xlo = 1.f, xhi = 1.f;
            split_row(std::false_type{}, xlo, xhi);
            if (__any_sync(0xffffffffu, !(c64_abs_in_range(xlo) && c64_abs_in_range(xhi))))
                split_row(std::true_type{}, xlo, xhi);
            if (lane == 0) row[1024] = fabsf((float)(2.0 * (z0re - z0im)));   // X[1024] = Re Z[0] - Im Z[0] (Z halved)
            // words 1025 .. 1091 follow the row: the mel-D loop below reads up to 51 words past a band's start with
            // zero weights, and what the exchange left there may be a NaN pattern
            row[1025 + lane] = 0.f;
            row[1057 + lane] = 0.f;
            if (lane < 3) row[1089 + lane] = 0.f;
        }
        __syncwarp();
        if ((t & 1) == 0 && live) {                            // hop-512 frame: keep the row for rolloff / tuning-36
            float4* dst = reinterpret_cast<float4*>(ws.mag_even + ((size_t)b * TE + (t >> 1)) * kMag2048Stride);
#pragma unroll
            for (int q = 0; q < (kMag2048Stride / 4 + 31) / 32; ++q) {
                const int i = lane + 32 * q;
                if (i < kMag2048Stride / 4) {
                    float4 v = reinterpret_cast<const float4*>(row)[i];
                    if (i == kMag2048Stride / 4 - 1) { v.y = 0.f; v.z = 0.f; v.w = 0.f; }
                    dst[i] = v;
                }
            }
        }
        // spectral_centroid / bandwidth (methods.py:59-60): moments of the L1-normalised column; flatness (:62) needs
        // sum(log(p)): the logs of up to 17 powers (each >= 1e-10) are taken as one double-precision log of their product.
        // Bin k = lane + 32 i: the lane accumulates sum m, sum i m, sum i^2 m with compile-time weights.
        double sm = 0.0, s1 = 0.0, s2 = 0.0, spow = 0.0, prod0 = 1.0, prod1 = 1.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float m = row[lane + 32 * i];
            const double dm = (double)m;
            sm += dm;
            if (i > 0) { s1 = fma(dm, (double)i, s1); s2 = fma(dm, (double)(i * i), s2); }
            const double p = (double)fmaxf(1e-10f, __fmul_rn(m, m));
            spow += p;
            if (i < 16) prod0 *= p; else prod1 *= p;
        }
        {
            const float m = lane == 0 ? row[1024] : 0.f;       // the Nyquist bin (i = 32) exists for lane 0 only
            const double dm = (double)m;
            sm += dm;
            s1 = fma(dm, 32.0, s1);
            s2 = fma(dm, 1024.0, s2);
            if (lane == 0) {
                const double p = (double)fmaxf(1e-10f, __fmul_rn(m, m));
                spow += p;
                prod1 *= p;
            }
        }
        double slog = log(prod0) + log(prod1);
        const double dl = (double)lane;
        const double smk = fma(32.0, s1, dl * sm);                                      // sum m k
        const double smk2 = fma(1024.0, s2, fma(64.0 * dl, s1, dl * dl * sm));          // sum m k^2
        const double smf = smk * 7.8125, smf2 = smk2 * (7.8125 * 7.8125);
        sm = warp_sum(sm);
        const double smf_w = warp_sum(smf), smf2_w = warp_sum(smf2);
        slog = warp_sum(slog);
        spow = warp_sum(spow);
        double* ff = ws.frame_feat + (size_t)f * kFrameFeat;
        if (lane == 0 && live) {
            const double len = sm < 1.17549435e-38 ? 1.0 : sm;     // util.normalize(norm=1): tiny(float32) guard
            const double c = smf_w / len;
            ff[0] = c;
            ff[1] = sqrt(fmax(0.0, smf2_w / len - 2.0 * c * (smf_w / len) + c * c * (sm / len)));
            const float gmean = expf((float)(slog / 1025.0));
            const float amean = (float)(spow / 1025.0);
            ff[2] = (double)__fdiv_rn(gmean, amean);
        }
        // spectral_contrast order statistics (methods.py:63); band table: first bin, length, order count
        //   {0,25,1} {25,26,1} {51,51,1} {102,102,2} {204,205,4} {409,410,8} {819,206,4}
        {
            double va[7], pk[7];
            warp_band_minmax(row, 0, 25, lane, &va[0], &pk[0]);
            warp_band_minmax(row, 25, 26, lane, &va[1], &pk[1]);
            warp_band_minmax(row, 51, 51, lane, &va[2], &pk[2]);
            warp_band_extremes<4, 4, 2>(row, 102, 102, lane, &va[3], &pk[3]);
            warp_band_extremes<8, 7, 4>(row, 204, 205, lane, &va[4], &pk[4]);
            warp_band_extremes<16, 13, 8>(row, 409, 410, lane, &va[5], &pk[5]);
            warp_band_extremes<8, 7, 4>(row, 819, 206, lane, &va[6], &pk[6]);
            if (lane < 7 && live) {
                double p = pk[0], v = va[0];
#pragma unroll
                for (int bnd = 1; bnd < 7; ++bnd) if (lane == bnd) { p = pk[bnd]; v = va[bnd]; }
                ff[3 + lane] = p;
                ff[10 + lane] = v;
            }
        }
        // mel-D power column (n_fft 2048, 128 mels, fmax 8000): methods.py:90 and process.py:74.  Weights are read
        // from the transposed band table ([tap][mel]: one coalesced request per tap); the squares are formed on the fly.
        float* md = ws.melD + (size_t)f * kPlaneRows;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = lane + 32 * i;
            const int s0 = __ldg(tb.mel_d.start + m), c = __ldg(tb.mel_d.count + m);
            const int cmax = __reduce_max_sync(0xffffffffu, c);
            const float* wp = tb.mel_d.wt + m;                   // [tap][mel], zero beyond a row's count and 4 zero taps
            const float* rp = row + s0;                          // past the widest row (upload_bank)
            float acc = 0.f;
            for (int j = 0; j < cmax; j += 4) {                  // same taps, same order as one tap per iteration
                const float w0 = __ldg(wp), w1 = __ldg(wp + kPlaneRows), w2 = __ldg(wp + 2 * kPlaneRows),
                            w3 = __ldg(wp + 3 * kPlaneRows);
                const float m0 = rp[0], m1 = rp[1], m2 = rp[2], m3 = rp[3];
                acc = fmaf(w0, __fmul_rn(m0, m0), acc);
                acc = fmaf(w1, __fmul_rn(m1, m1), acc);
                acc = fmaf(w2, __fmul_rn(m2, m2), acc);
                acc = fmaf(w3, __fmul_rn(m3, m3), acc);
                wp += 4 * kPlaneRows;
                rp += 4;
            }
            if (live) md[m] = acc;
        }
        __syncwarp();                                          // the row is the next frame's exchange buffer
    }
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, kF2WarpsPerSm / WARPS) k_frame2048(const float* __restrict__ y, Geometry g,
                                                                                 Tables tb, Workspace ws,
                                                                                 int total_frames, int frames_per_cta) {
    frame2048_body<WARPS>(y, g, tb, ws, total_frames, frames_per_cta);
}
// ============================================================================================ k_frame2048_w2 (r02-k)
// The same frame pipeline on TWO warps per frame (a 64-thread CTA = one frame at a time): 16 instead of 32
// register-resident points per thread (team64_fft, fft_reg.cuh: three register stages 16 x 16 x 4, the first exchange
// across the CTA, the second inside groups of four lanes), 128 instead of 168 registers, 16 instead of 12 warps per SM.
// Per frame the two warps execute what the one warp executed (the FFT has the same operation count; the second exchange
// adds 64 shared-memory instructions per thread), every per-frame consumer is split between them:
//   * real split: thread c = k1 + 16 q holds Z[c + 64 u]; its conjugate partner c' = (64 - c) % 64 is a lane of the
//     same warp by construction of the thread map (t64_partner_lane), so the pairing is two shuffles as before;
//   * moments / flatness: bins k = tid + 64 i, combined across the warps through 5 doubles of shared memory;
//   * spectral contrast: warp 0 takes the 410-bin band (sort network of 13, eight extraction rounds), warp 1 the six
//     others -- about the same instruction count;
//   * mel-D: warp 0 rows 0-31 and 96-127, warp 1 rows 32-95 (narrowest + widest bands against the two middle groups).
// Same formulas per bin as k_frame2048; the FFT factorisation differs, so |X| agrees to the last float32 bit or two
// (the kernel passes the whole GPU parity suite as the default path).
// MEASURED AND NOT THE DEFAULT (BPC_F2_W2=1 selects it; profiles/r02_k_frame2048_w2.txt): 2.62 ms against 2.39 ms for
// the one-warp kernel.  The question it answers is whether k_frame2048 is bound by its three warps per scheduler: it
// is not -- at 16 warps per SM the issue slots are as busy as at 12 (47.9 % against 48.2 %), while the two-warp form
// executes 11 % more instructions (per-frame prologue, twiddle and table loads per warp instead of per frame, the second
// exchange) and pushes the L1 data pipe from 50 % to 69 % of its wavefront peak (window, twiddle and split tables are
// re-read from L1 for every frame by every thread: ~130 KB through a 128 B / clk pipe per frame, plus 32 KB for the second
// exchange).  With 96 registers (20 warps per SM, 33 doubles spilled per frame) it runs at 2.84 ms.
#ifndef BPC_F2W2_CTAS
#define BPC_F2W2_CTAS 8
#endif
constexpr int kF2w2CtasPerSm = BPC_F2W2_CTAS;
#ifdef BPC_F2W2_REGS
__global__ void __maxnreg__(BPC_F2W2_REGS) k_frame2048_w2(
#else
__global__ void __launch_bounds__(64, kF2w2CtasPerSm) k_frame2048_w2(
#endif
    const float* __restrict__ y, Geometry g, Tables tb, Workspace ws, int total_frames, int frames_per_cta) {
    __shared__ __align__(16) double xr_s[16 * kT64Pitch];      // exchange buffer, then the |X| row (1092 floats)
    __shared__ double red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int k1 = t64_k1(tid), q = tid & 3, c = k1 + 16 * q;
    const int partner = t64_partner_lane(tid);
    float* row = reinterpret_cast<float*>(xr_s);
    const int T = g.T, L = g.L, hop = g.hop, TE = (T + 1) / 2;
    const double2* win2 = reinterpret_cast<const double2*>(tb.hann2048h) + tid;   // 0.5 * Hann: the 1/2 of the real split
    const double2* twa = tb.t64a + tid;
    const double2* twb = tb.t64b + q;
    const double2* rs = tb.rs2048 + c;                                    // split twiddles as (scale, tan / cot)
    const int f_lo = blockIdx.x * frames_per_cta;
    const int f_hi = f_lo + frames_per_cta < total_frames ? f_lo + frames_per_cta : total_frames;
    for (int f = f_lo; f < f_hi; ++f) {
        const int b = f / T, t = f - b * T;
        const float* yb = y + (size_t)b * L;
        const int g0 = t * hop - 1024;
        {
            double2 a[16];
            if (g0 >= 0 && g0 + 2048 <= L) {                   // interior frame (55 of 63): no bounds checks, one base pointer
                const float2* src = reinterpret_cast<const float2*>(yb + g0) + tid;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 v = __ldg(src + 64 * j);
                    const double2 wv = __ldg(win2 + 64 * j);
                    a[j] = make_double2((double)v.x * wv.x, (double)v.y * wv.y);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int gi = g0 + 2 * (tid + 64 * j);    // even; L is even, so the pair is in or out together
                    float2 v = make_float2(0.f, 0.f);
                    if (gi >= 0 && gi < L) v = __ldg(reinterpret_cast<const float2*>(yb + gi));
                    const double2 wv = __ldg(win2 + 64 * j);
                    a[j] = make_double2((double)v.x * wv.x, (double)v.y * wv.y);
                }
            }
            team64_fft(a, twa, twb, xr_s, tid, k1, q);
            __syncthreads();                                   // everyone has its exchange-2 values: the buffer becomes the row
            // real split, every conjugate pair once (see k_frame2048): bin k = c + 64 u, u < 8, pairs with N - k, which
            // the partner lane holds as its u' = 15 - u (c = 0: this thread's own (16 - u) % 16); k = 512 is c = 0, u = 8
            const double z0re = a[0].x, z0im = a[0].y;         // c = 0: Z[0]
            auto split_row = [&](auto exact_tag, float& xlo, float& xhi) {
                constexpr bool EXACT = decltype(exact_tag)::value;
#pragma unroll
                for (int U = 0; U <= 8; ++U) {
                    const double2 zk = a[t64_zidx(U)];
                    double2 zn;
                    if (U < 8) {
                        const double2 src = a[t64_zidx(15 - U)];
                        zn.x = __shfl_sync(0xffffffffu, src.x, partner);
                        zn.y = __shfl_sync(0xffffffffu, src.y, partner);
                        if (c == 0) zn = a[t64_zidx((16 - U) & 15)];
                    } else {
                        zn = zk;                                   // k = 512 pairs with itself
                    }
                    const double2 ct = __ldg(rs + 64 * U);
                    const double sx = zk.x + zn.x, sy = zk.y - zn.y, dx = zk.x - zn.x, dy = zk.y + zn.y;
                    double qx, qy;
                    if (U < 4) { qx = fma(ct.y, dx, -dy); qy = fma(ct.y, dy, dx); }       // k < 256
                    else       { qx = fma(-ct.y, dy, dx); qy = fma(ct.y, dx, dy); }
                    const double x0 = fma(ct.x, qx, sx), y0 = fma(ct.x, qy, sy);
                    const double x1 = fma(-ct.x, qx, sx), y1 = fma(ct.x, qy, -sy);
                    const int k = c + 64 * U;
                    float m0, m1, xa = 1.f, xb = 1.f;
                    if (EXACT) {
                        m0 = c64_abs_exact((float)x0, (float)y0);
                        m1 = c64_abs_exact((float)x1, (float)y1);
                    } else {
                        m0 = c64_abs_f32_unchecked((float)x0, (float)y0, &xa);
                        m1 = c64_abs_f32_unchecked((float)x1, (float)y1, &xb);
                    }
                    if (U < 8) {
                        row[k] = m0;
                        xlo = fminf(xlo, xa);
                        xhi = fmaxf(xhi, xa);
                        if (U > 0 || c > 0) {
                            row[1024 - k] = m1;
                            xlo = fminf(xlo, xb);
                            xhi = fmaxf(xhi, xb);
                        }
                    } else if (c == 0) {
                        row[512] = m0;
                        xlo = fminf(xlo, xa);
                        xhi = fmaxf(xhi, xa);
                    }
                }
            };
            float xlo = 1.f, xhi = 1.f;
            split_row(std::false_type{}, xlo, xhi);
            if (__any_sync(0xffffffffu, !(c64_abs_in_range(xlo) && c64_abs_in_range(xhi))))
                split_row(std::true_type{}, xlo, xhi);            // each warp redoes its own bins (all-zero frames)
            if (c == 0) row[1024] = fabsf((float)(2.0 * (z0re - z0im)));   // X[1024] = Re Z[0] - Im Z[0] (Z halved)
            // words 1025 .. 1091 follow the row: the mel-D loop reads up to 51 words past a band's start with zero
            // weights, and what the exchange left there may be a NaN pattern
            row[1025 + tid] = 0.f;
            if (tid < 3) row[1089 + tid] = 0.f;
        }
        __syncthreads();                                       // the row is complete
        if ((t & 1) == 0) {                                    // hop-512 frame: keep the row for rolloff / tuning-36
            float4* dst = reinterpret_cast<float4*>(ws.mag_even + ((size_t)b * TE + (t >> 1)) * kMag2048Stride);
#pragma unroll
            for (int r = 0; r < (kMag2048Stride / 4 + 63) / 64; ++r) {
                const int i = tid + 64 * r;
                if (i < kMag2048Stride / 4) {
                    float4 v = reinterpret_cast<const float4*>(row)[i];
                    if (i == kMag2048Stride / 4 - 1) { v.y = 0.f; v.z = 0.f; v.w = 0.f; }
                    dst[i] = v;
                }
            }
        }
        // spectral_centroid / bandwidth / flatness (methods.py:59-62): bin k = tid + 64 i, the thread accumulates
        // sum m, sum i m, sum i^2 m with compile-time weights and the product of its 16 (thread 0: 17) clamped powers
        {
            double sm = 0.0, s1 = 0.0, s2 = 0.0, spow = 0.0, prod = 1.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float m = row[tid + 64 * i];
                const double dm = (double)m;
                sm += dm;
                if (i > 0) { s1 = fma(dm, (double)i, s1); s2 = fma(dm, (double)(i * i), s2); }
                const double p = (double)fmaxf(1e-10f, __fmul_rn(m, m));
                spow += p;
                prod *= p;
            }
            {
                const float m = tid == 0 ? row[1024] : 0.f;     // the Nyquist bin (i = 16) exists for thread 0 only
                const double dm = (double)m;
                sm += dm;
                s1 = fma(dm, 16.0, s1);
                s2 = fma(dm, 256.0, s2);
                if (tid == 0) {
                    const double p = (double)fmaxf(1e-10f, __fmul_rn(m, m));
                    spow += p;
                    prod *= p;
                }
            }
            double slog = log(prod);
            const double dt = (double)tid;
            const double smk = fma(64.0, s1, dt * sm);                                       // sum m k
            const double smk2 = fma(4096.0, s2, fma(128.0 * dt, s1, dt * dt * sm));          // sum m k^2
            const double smf = smk * 7.8125, smf2 = smk2 * (7.8125 * 7.8125);
            sm = warp_sum(sm);
            const double smf_w = warp_sum(smf), smf2_w = warp_sum(smf2);
            slog = warp_sum(slog);
            spow = warp_sum(spow);
            if (lane == 0) {
                red[w][0] = sm; red[w][1] = smf_w; red[w][2] = smf2_w; red[w][3] = slog; red[w][4] = spow;
            }
        }
        double* ff = ws.frame_feat + (size_t)f * kFrameFeat;
        // spectral_contrast order statistics (methods.py:63); band table: first bin, length, order count
        //   {0,25,1} {25,26,1} {51,51,1} {102,102,2} {204,205,4} {409,410,8} {819,206,4}
        if (w == 0) {
            double va, pk;
            warp_band_extremes<16, 13, 8>(row, 409, 410, lane, &va, &pk);
            if (lane == 0) { ff[3 + 5] = pk; ff[10 + 5] = va; }
        } else {
            double va[7], pk[7];
            warp_band_minmax(row, 0, 25, lane, &va[0], &pk[0]);
            warp_band_minmax(row, 25, 26, lane, &va[1], &pk[1]);
            warp_band_minmax(row, 51, 51, lane, &va[2], &pk[2]);
            warp_band_extremes<4, 4, 2>(row, 102, 102, lane, &va[3], &pk[3]);
            warp_band_extremes<8, 7, 4>(row, 204, 205, lane, &va[4], &pk[4]);
            warp_band_extremes<8, 7, 4>(row, 819, 206, lane, &va[6], &pk[6]);
            va[5] = 0.0; pk[5] = 0.0;
            if (lane < 7 && lane != 5) {
                double p = pk[0], v = va[0];
#pragma unroll
                for (int bnd = 1; bnd < 7; ++bnd) if (lane == bnd) { p = pk[bnd]; v = va[bnd]; }
                ff[3 + lane] = p;
                ff[10 + lane] = v;
            }
        }
        // mel-D power column (n_fft 2048, 128 mels, fmax 8000): methods.py:90 and process.py:74.  Row groups of 32:
        // warp 0 takes groups 0 and 3, warp 1 groups 1 and 2.
        float* md = ws.melD + (size_t)f * kPlaneRows;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = lane + 32 * (i == 0 ? w : 3 - w);
            const int s0 = __ldg(tb.mel_d.start + m), cnt = __ldg(tb.mel_d.count + m);
            const int cmax = __reduce_max_sync(0xffffffffu, cnt);
            const float* wp = tb.mel_d.wt + m;                   // [tap][mel], zero beyond a row's count and 4 zero taps
            const float* rp = row + s0;                          // past the widest row (upload_bank)
            float acc = 0.f;
            for (int j = 0; j < cmax; j += 4) {                  // same taps, same order as one tap per iteration
                const float w0 = __ldg(wp), w1 = __ldg(wp + kPlaneRows), w2 = __ldg(wp + 2 * kPlaneRows),
                            w3 = __ldg(wp + 3 * kPlaneRows);
                const float m0 = rp[0], m1 = rp[1], m2 = rp[2], m3 = rp[3];
                acc = fmaf(w0, __fmul_rn(m0, m0), acc);
                acc = fmaf(w1, __fmul_rn(m1, m1), acc);
                acc = fmaf(w2, __fmul_rn(m2, m2), acc);
                acc = fmaf(w3, __fmul_rn(m3, m3), acc);
                wp += 4 * kPlaneRows;
                rp += 4;
            }
            md[m] = acc;
        }
        __syncthreads();                                       // red[] is complete; the row is the next frame's exchange buffer
        if (tid == 0) {
            const double sm = red[0][0] + red[1][0], smf_w = red[0][1] + red[1][1], smf2_w = red[0][2] + red[1][2];
            const double slog = red[0][3] + red[1][3], spow = red[0][4] + red[1][4];
            const double len = sm < 1.17549435e-38 ? 1.0 : sm;     // util.normalize(norm=1): tiny(float32) guard
            const double cc = smf_w / len;
            ff[0] = cc;
            ff[1] = sqrt(fmax(0.0, smf2_w / len - 2.0 * cc * (smf_w / len) + cc * cc * (sm / len)));
            const float gmean = expf((float)(slog / 1025.0));
            const float amean = (float)(spow / 1025.0);
            ff[2] = (double)__fdiv_rn(gmean, amean);
        }
    }
}

// ================================================================================================= k_seg2048
constexpr int kSegGroups = 3;                // tempogram frames in flight per CTA (96 threads x 4 lags each)
constexpr int kSegThreads = 96 * kSegGroups; // 9 warps (v41: two groups; the kernel runs four CTAs per SM either way)

struct Seg2048Smem {
    union {
        float melD[kMaxFrames * kPlaneRows];          // [t][m] power -> 10 log10
        float tg[kPlaneRows * kMaxFrames];            // tempogram rows 0..127, raw autocorrelation [lag][t]
    } u;
    double peak[7 * kMaxFrames], valley[7 * kMaxFrames];
    double cent[kMaxFrames], bw[kMaxFrames], flat[kMaxFrames];
    float onset[kMaxFrames + 2 * 192 + 8];            // onset envelope with the tempogram's 192-sample pads
    __align__(16) float frame[kSegGroups][kTempoLags + 8];
    float flux[kMaxFrames];
    double sumv[kMaxFrames], sumq[kMaxFrames];
    float ac0[kMaxFrames];
    double dscratch[32];
    float fscratch[32];
};

__device__ __forceinline__ void group_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool LONG>
__global__ void __launch_bounds__(kSegThreads) k_seg2048(Geometry g, Tables tb, Workspace ws, float* feats,
                                                          float* scalars, int phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Seg2048Smem& S = *reinterpret_cast<Seg2048Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = kSegThreads, NW = kSegThreads / 32;
    const int b = blockIdx.x, T = g.T, NP = kPlaneRows * T;
    float* sc = scalars + (size_t)b * g.nscal;
    // per-segment arrays: shared memory for 1 s segments, the segment's global scratch region in long mode
    struct { float *melD, *tg; double *peak, *valley, *cent, *bw, *flat, *sumv, *sumq; float *onset, *flux, *ac0; } V;
    if (LONG) {
        double* d = reinterpret_cast<double*>(ws.scratch + (size_t)b * ws.scratch_stride);
        V.peak = d; d += 7 * T; V.valley = d; d += 7 * T; V.cent = d; d += T; V.bw = d; d += T; V.flat = d; d += T;
        V.sumv = d; d += T; V.sumq = d; d += T;
        float* f = reinterpret_cast<float*>(d + (T & 1));                      // keep 16-byte alignment for the float4 staging
        V.melD = f; V.tg = f; f += (size_t)kPlaneRows * T;
        V.onset = f; f += T + 2 * 192 + 8; V.flux = f; f += T; V.ac0 = f;
    } else {
        V.melD = S.u.melD; V.tg = S.u.tg; V.peak = S.peak; V.valley = S.valley; V.cent = S.cent; V.bw = S.bw;
        V.flat = S.flat; V.sumv = S.sumv; V.sumq = S.sumq; V.onset = S.onset; V.flux = S.flux; V.ac0 = S.ac0;
    }

    // long mode runs three launches: phase 1 = statistics / flux / onset envelope (grid (segment)), phase 2 = the
    // tempogram frames of this CTA's share (grid (segment, part)), phase 3 = normalisation; phase 0 = all (1 s)
    if (!LONG || phase == 1) {
    {   // stage per-frame features and mel-D
        const float4* src = reinterpret_cast<const float4*>(ws.melD + (size_t)b * T * kPlaneRows);
        // (four trips' loads in flight before the first store: the plain copy loops were one L2 round trip per trip)
        const int n4 = T * kPlaneRows / 4;
        for (int i0 = tid; i0 < n4; i0 += 4 * NT) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(src + min(i0 + u * NT, n4 - 1));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * NT < n4) reinterpret_cast<float4*>(V.melD)[i0 + u * NT] = v[u];
        }
        const double* ff = ws.frame_feat + (size_t)b * T * kFrameFeat;
        for (int i0 = tid; i0 < T * 17; i0 += 4 * NT) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = min(i0 + u * NT, T * 17 - 1), t = i / 17;
                v[u] = __ldg(ff + t * kFrameFeat + (i - t * 17));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NT;
                if (i < T * 17) {
                    const int t = i / 17, j = i - t * 17;
                    if (j == 0) V.cent[t] = v[u];
                    else if (j == 1) V.bw[t] = v[u];
                    else if (j == 2) V.flat[t] = v[u];
                    else if (j < 10) V.peak[(j - 3) * T + t] = v[u];
                    else V.valley[(j - 10) * T + t] = v[u];
                }
            }
        }
    }
    __syncthreads();
    // ---- centroid / bandwidth / flatness statistics (methods.py:64-68), warps 0..2
    if (warp < 3) {
        const double* a = warp == 0 ? V.cent : (warp == 1 ? V.bw : V.flat);
        double s = 0.0;
        for (int t = lane; t < T; t += 32) s += a[t];
        s = warp_sum(s);
        const double mean = s / T;
        double m2 = 0.0, m3 = 0.0;
        for (int t = lane; t < T; t += 32) {
            const double d = a[t] - mean;
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
        for (int i = tid; i < 7 * T; i += NT) {
            const double p = 10.0 * log10(fmax(1e-10, V.peak[i]));
            const double v = 10.0 * log10(fmax(1e-10, V.valley[i]));
            V.peak[i] = p;
            V.valley[i] = v;
            pmax = fmax(pmax, p);
            vmax = fmax(vmax, v);
        }
        pmax = block_reduce(pmax, -1e300, OpMaxD(), S.dscratch);
        vmax = block_reduce(vmax, -1e300, OpMaxD(), S.dscratch);
        double s = 0.0, q = 0.0;
        for (int i = tid; i < 7 * T; i += NT) {
            const double c = fmax(V.peak[i], pmax - 80.0) - fmax(V.valley[i], vmax - 80.0);
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
    for (int i = tid; i < NP; i += NT) pmx = fmaxf(pmx, V.melD[i]);
    pmx = block_max(pmx, S.fscratch);
    const float ref_db = (float)(10.0 * log10((double)fmaxf(1e-10f, pmx)));
    float lmax = -FLT_MAX;
    for (int i = tid; i < NP; i += NT) {
        const float l = __fmul_rn(10.0f, log10f(fmaxf(1e-10f, V.melD[i])));
        V.melD[i] = l;
        lmax = fmaxf(lmax, l);
    }
    lmax = block_max(lmax, S.fscratch);                                     // barrier inside: melD complete
    const float floor1 = __fsub_rn(lmax, 80.0f);                            // ref = 1.0 variant
    const float floorm = __fsub_rn(__fsub_rn(lmax, ref_db), 80.0f);         // ref = max variant
    for (int j = tid; j < T + 2 * 192 + 8; j += NT) V.onset[j] = 0.f;
    __syncthreads();
    for (int tf0 = 0; tf0 < T - 1; tf0 += NW) {         // uniform trip count: convergent warp reductions
        const bool tlive = tf0 + warp < T - 1;
        const int t = tlive ? tf0 + warp : T - 2;
        double f2 = 0.0, on = 0.0;
        for (int m = lane; m < kPlaneRows; m += 32) {
            const float l0 = V.melD[t * kPlaneRows + m], l1 = V.melD[(t + 1) * kPlaneRows + m];
            const float a0 = fmaxf(__fsub_rn(l0, ref_db), floorm), a1 = fmaxf(__fsub_rn(l1, ref_db), floorm);
            const float d = __fsub_rn(a1, a0);
            f2 += (double)__fmul_rn(d, d);
            const float o = __fsub_rn(fmaxf(l1, floor1), fmaxf(l0, floor1));
            on += (double)fmaxf(0.f, o);
        }
        f2 = warp_sum(f2);
        on = warp_sum(on);
        if (lane == 0 && tlive) {
            V.flux[t] = sqrtf((float)f2);
            // onset_env = pad(mean over mels, (1 + 2048 // (2 * 256), 0))[:T]; stored at offset 192 (left tempogram pad)
            if (t + 5 < T) V.onset[192 + t + 5] = (float)(on / (double)kPlaneRows);
        }
    }
    __syncthreads();
    if (warp == 0) {
        double s = 0.0, q = 0.0;
        float mx = -FLT_MAX;
        for (int t = lane; t < T - 1; t += 32) {
            s += (double)V.flux[t];
            q += (double)V.flux[t] * (double)V.flux[t];
            mx = fmaxf(mx, V.flux[t]);
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
        for (int t = tid; t < T; t += NT) ws.dbg_onset[(size_t)b * T + t] = V.onset[192 + t];
    // ---- tempogram (process.py:75): linear-ramp pad 192, 384-sample Hann frames at hop 1, autocorrelation, /max
    if (tid < 192) {
        const float edge = V.onset[192 + T - 1];
        const float step = __fdiv_rn(edge, 192.0f);
        V.onset[192 + T + tid] = __fmul_rn((float)(191 - tid), step);       // np.pad(mode='linear_ramp', end 0)
    }
    for (int t = tid; t < T; t += NT) { V.sumv[t] = 0.0; V.sumq[t] = 0.0; }
    __syncthreads();
    }
    if (LONG && phase == 1) return;
    if (!LONG || phase == 2) {
    // kSegGroups frames in flight: group gidx (96 threads) owns frames gidx, gidx + kSegGroups, ...; thread j owns lags
    // 4j .. 4j + 3.
    const int gidx = tid / 96, j = tid - gidx * 96, l0 = 4 * j;
    float* F = S.frame[gidx];
    // trip count from blockIdx only; a group past the last frame redoes frame T - 1 and drops the result, so that the
    // warp reductions below are provably convergent (plain SHFL)
    const int t_step = kSegGroups * (LONG ? (int)gridDim.y : 1);
    for (int t0 = LONG ? kSegGroups * (int)blockIdx.y : 0; t0 < T; t0 += t_step) {
        const bool live = t0 + gidx < T;
        const int t = live ? t0 + gidx : T - 1;
        {
            for (int n = j; n < kTempoLags + 8; n += 96)
                F[n] = n < kTempoLags ? (float)((double)V.onset[t + n] * __ldg(tb.hann384 + n)) : 0.f;
        }
        group_bar(1 + gidx, 96);
        {
            // structural zeros: onset[0..4] == 0 and the left pad is 0, so F[n] == 0 for n < 197 - t
            int n = (197 - t) > 0 ? ((197 - t) & ~3) : 0;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            // the window F[n + l0 .. n + l0 + 7] slides by four per step: its upper half is carried over in registers
            // (two shared-memory loads per 16 FMAs instead of three; the kernel is bound by shared-memory wavefronts)
            float4 B0 = *reinterpret_cast<const float4*>(F + min(n + l0, kTempoLags + 4));
            for (; n + l0 < kTempoLags; n += 4) {
                const float4 A = *reinterpret_cast<const float4*>(F + n);
                const float4 B1 = *reinterpret_cast<const float4*>(F + n + l0 + 4);
                a0 = fmaf(A.x, B0.x, a0); a0 = fmaf(A.y, B0.y, a0); a0 = fmaf(A.z, B0.z, a0); a0 = fmaf(A.w, B0.w, a0);
                a1 = fmaf(A.x, B0.y, a1); a1 = fmaf(A.y, B0.z, a1); a1 = fmaf(A.z, B0.w, a1); a1 = fmaf(A.w, B1.x, a1);
                a2 = fmaf(A.x, B0.z, a2); a2 = fmaf(A.y, B0.w, a2); a2 = fmaf(A.z, B1.x, a2); a2 = fmaf(A.w, B1.y, a2);
                a3 = fmaf(A.x, B0.w, a3); a3 = fmaf(A.y, B1.x, a3); a3 = fmaf(A.z, B1.y, a3); a3 = fmaf(A.w, B1.z, a3);
                B0 = B1;
            }
            if (j == 0 && live) V.ac0[t] = a0;
            if (l0 < kPlaneRows && live) {
                V.tg[(l0 + 0) * T + t] = a0; V.tg[(l0 + 1) * T + t] = a1;
                V.tg[(l0 + 2) * T + t] = a2; V.tg[(l0 + 3) * T + t] = a3;
            }
            double sv = (double)a0 + (double)a1 + (double)a2 + (double)a3;
            double sq = (double)a0 * a0 + (double)a1 * a1 + (double)a2 * a2 + (double)a3 * a3;
            sv = warp_sum(sv);
            sq = warp_sum(sq);
            if (lane == 0 && live) { atomicAdd(&V.sumv[t], sv); atomicAdd(&V.sumq[t], sq); }
        }
        group_bar(1 + gidx, 96);
    }
    }
    if (LONG && phase == 2) return;
    __syncthreads();
    // util.normalize(norm=inf) divides each column by max |.| (= lag 0); z-score over all 384 x T values (process.py:76)
    double s = 0.0, q = 0.0;
    for (int t = tid; t < T; t += NT) {
        const double c = (double)V.ac0[t] < 2.2250738585072014e-308 ? 1.0 : (double)V.ac0[t];
        s += V.sumv[t] / c;
        q += V.sumq[t] / (c * c);
    }
    s = block_sum(s, S.dscratch);
    q = block_sum(q, S.dscratch);
    const double mean = s / (double)(kTempoLags * T);
    const double sd = sqrt(fmax(0.0, q / (double)(kTempoLags * T) - mean * mean));
    const double inv_sd = 1.0 / (sd + 1e-8);
    // one reciprocal per column instead of two FP64 divisions per element (they were 9 % of this kernel)
    for (int t = tid; t < T; t += NT) V.sumv[t] = 1.0 / ((double)V.ac0[t] < 2.2250738585072014e-308 ? 1.0 : (double)V.ac0[t]);
    __syncthreads();
    float* o = plane_ptr(feats, b, BPC_CH_TEMPOGRAM, T);
    int t = tid % T;
    const int tstep = NT % T;
    const bool stats = !LONG && ws.stats_acc != nullptr;
    StatAcc at;
    at.init();
    for (int i = tid; i < NP; i += NT) {                                     // pad_freq truncates to the first 128 lags
        const float v = (float)(((double)V.tg[i] * V.sumv[t] - mean) * inv_sd);
        o[i] = v;
        if (stats) at.add(v);
        t += tstep;
        if (t >= T) t -= T;
    }
    if (stats) stat_flush_block(at, ws.stats_acc + 5 * BPC_CH_TEMPOGRAM, S.dscratch, S.fscratch);
}

// ------------------------------------------------------------------------------------ even frames: rolloff + tuning36
// spectral_rolloff (methods.py:61) is called without hop_length -> hop 512, i.e. the even hop-256 frames; its decision
// `cumsum(S) < 0.85 * cumsum(S)[-1]` is taken on numpy's *sequential float32* cumsum, which is reproduced here with
// one thread per frame (1025 dependent float32 adds).  chroma_cens (process.py:53) estimates its tuning from the same
// frames (estimate_tuning(y=y, bins_per_octave=36) -> piptrack n_fft 2048, hop 512).
// The candidate lists (worst case 7936 entries, typically a few hundred) live in global memory: as 63.5 KB of shared
// memory they held this latency-bound kernel at three CTAs per SM (r01 v40).
struct Even2048Smem {
    float sortbuf[kSelectWords];
    float colmax[kMaxFrames];
    float roll[kMaxFrames];
    int hist[100];
};

template <bool LONG>
__global__ void __launch_bounds__(256) k_even2048(Geometry g, Tables tb, Workspace ws, float* scalars,
                                                  int32_t* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Even2048Smem& E = *reinterpret_cast<Even2048Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, T = g.T, TE = (T + 1) / 2;
    // candidate lists and per-frame arrays: shared memory (1 s), the segment's global scratch region in long mode
    const int cap = LONG ? 492 * TE : kMaxCand2048;
    float* lbase = ws.scratch + (size_t)b * ws.scratch_stride;
    float* cand_mag = LONG ? lbase : ws.cand36 + (size_t)b * 2 * kMaxCand2048;
    float* cand_pitch = cand_mag + cap;
    float* colmax = LONG ? lbase + 2 * (size_t)cap : E.colmax;
    float* roll = LONG ? colmax + TE : E.roll;
    float* sortbuf = E.sortbuf;
    int* hist = E.hist;
    __shared__ int s_ncand;
    // even frame f is row f of the segment's mag_even block (written by k_frame2048)
    const float* mag_b = ws.mag_even + (size_t)b * TE * kMag2048Stride;
    const size_t fstride = (size_t)kMag2048Stride;
    if (tid == 0) s_ncand = 0;
    // 1 s (32 even frames): warp 0 runs the sequential cumsums, one frame per lane, while warps 1..7 take the column
    // maxima.  Long mode (hundreds of frames): all eight warps do the maxima, then all 256 threads take one frame each.
    const int mx_first = LONG ? warp : warp - 1, mx_step = LONG ? 8 : 7;
    if (LONG || warp > 0) {
        for (int f = mx_first; f < TE; f += mx_step) {
            const float4* col4 = reinterpret_cast<const float4*>(mag_b + f * fstride);   // pad words 1025..1027 are 0
            float mx = 0.f;
            for (int i = lane; i < kMag2048Stride / 4; i += 32) {
                const float4 v = __ldg(col4 + i);
                mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            }
            mx = warp_max(mx);
            if (lane == 0) colmax[f] = mx;
        }
    }
    if (LONG || warp == 0) {
        // numpy's sequential float32 cumsum, one frame per lane; the loads run 32 elements ahead of the dependent adds
        for (int f = LONG ? tid : lane; f < TE; f += LONG ? 256 : 32) {
            const float* col = mag_b + f * fstride;
            const float4* col4 = reinterpret_cast<const float4*>(col);
            float c = 0.f;
            for (int q = 0; q < 256; q += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(col4 + q + u);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    c = __fadd_rn(c, v[u].x); c = __fadd_rn(c, v[u].y); c = __fadd_rn(c, v[u].z); c = __fadd_rn(c, v[u].w);
                }
            }
            c = __fadd_rn(c, __ldg(col + 1024));
            const float thr = __fmul_rn(0.85f, c);
            float c2 = 0.f;
            int kk = -1;
            for (int q = 0; q < 256 && kk < 0; q += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(col4 + q + u);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        c2 = __fadd_rn(c2, e[w]);
                        if (kk < 0 && !(c2 < thr)) kk = 4 * (q + u) + w;
                    }
                }
            }
            if (kk < 0) kk = 1024;
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
        const float* col = mag_b + f * fstride;
        float pitch, mv;
        if (piptrack_candidate(__ldg(col + k - 1), __ldg(col + k), __ldg(col + k + 1), __fmul_rn(0.1f, colmax[f]), k,
                               7.8125, &pitch, &mv)) {
            const int slot = atomicAdd(&s_ncand, 1);
            if (slot < cap) { cand_mag[slot] = mv; cand_pitch[slot] = pitch; }
        }
    }
    __syncthreads();
    int n = s_ncand;
    unsigned flags = 0;
    if (n > cap) { n = cap; flags |= BPC_SEG_CAND_OVERFLOW; }
    bool empty = false;
    const int tbin = tuning_from_candidates(cand_mag, cand_pitch, n, sortbuf, hist, tb.hist_edges, 36, &empty);
    if (empty) flags |= BPC_SEG_TUNING_EMPTY;
    if (tid == 0) {
        ws.tuning[b * 2 + 1] = tbin;
        if (status && flags) atomicOr((unsigned int*)&status[b], flags);
    }
}

// ==================================================================================================== launchers
void launch_spec2048(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                     float* scalars, cudaStream_t st) {
    static PerDeviceOnce once;
    static int sms = 148, warps = 1;
    once.run([&] {
        cudaFuncSetAttribute(k_seg2048<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Seg2048Smem));
        cudaFuncSetAttribute(k_seg2048<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Seg2048Smem));
        cudaFuncSetAttribute(k_frame2048<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kF2RowBytes);
        cudaFuncSetAttribute(k_frame2048<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
        cudaFuncSetAttribute(k_frame2048<12>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
        cudaFuncSetAttribute(k_frame2048_w2, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
        if (const char* e = getenv("BPC_F2_WARPS")) warps = atoi(e) == 12 ? 12 : 1;    // A/B switch
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    });
    const int total = n * g.T;
    if (total <= 0) return;
    // r02-k: two warps per frame, measured slower (see k_frame2048_w2) and kept behind BPC_F2_W2=1 (read per launch so
    // that a test can compare the two kernels in one process)
    const char* w2e = std::getenv("BPC_F2_W2");
    if (w2e && std::atoi(w2e) == 1) {
        int grid2 = sms * kF2w2CtasPerSm;
        const int per2 = (total + grid2 - 1) / grid2;
        grid2 = (total + per2 - 1) / per2;
        k_frame2048_w2<<<grid2, 64, 0, st>>>(y, g, tb, ws, total, per2);
        note_launch();
        return;
    }
    // persistent: the 12 warps an SM holds at 168 registers, each CTA walking a contiguous range of frames
    int grid = warps == 1 ? sms * kF2WarpsPerSm : sms;
    int per = (total + grid - 1) / grid;
    per = (per + warps - 1) / warps * warps;
    grid = (total + per - 1) / per;
    if (warps == 1) k_frame2048<1><<<grid, 32, kF2RowBytes, st>>>(y, g, tb, ws, total, per);
    else k_frame2048<12><<<grid, 32 * 12, 12 * kF2RowBytes, st>>>(y, g, tb, ws, total, per);
    note_launch();
}

void launch_seg2048(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats, float* scalars,
                    cudaStream_t st) {
    if (g.long_mode) {
        k_seg2048<true><<<dim3(n, 1), kSegThreads, sizeof(Seg2048Smem), st>>>(g, tb, ws, feats, scalars, 1);
        k_seg2048<true><<<dim3(n, 16), kSegThreads, sizeof(Seg2048Smem), st>>>(g, tb, ws, feats, scalars, 2);
        k_seg2048<true><<<dim3(n, 1), kSegThreads, sizeof(Seg2048Smem), st>>>(g, tb, ws, feats, scalars, 3);
        note_launch(2);
    } else {
        k_seg2048<false><<<n, kSegThreads, sizeof(Seg2048Smem), st>>>(g, tb, ws, feats, scalars, 0);
    }
    note_launch();
}

void launch_even2048(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* scalars,
                     int32_t* status, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_even2048<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Even2048Smem));
        cudaFuncSetAttribute(k_even2048<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Even2048Smem));
    });
    if (g.long_mode) k_even2048<true><<<n, 256, sizeof(Even2048Smem), st>>>(g, tb, ws, scalars, status);
    else k_even2048<false><<<n, 256, sizeof(Even2048Smem), st>>>(g, tb, ws, scalars, status);
    note_launch();
}

}  // namespace bpc
