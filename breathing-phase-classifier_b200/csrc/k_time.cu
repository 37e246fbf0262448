// Time-domain scalar features (methods.py:48-114), one CTA per segment:
//   k_time_basic  rms / zcr frame statistics (idx 0-7), skew / kurtosis / percentiles of |y| (idx 29-32), status bits
//   k_autocorr    normalised autocorrelation lags 160 / 320 and first-minimum index over lags < 800 (idx 33-35)
//   k_hilbert     |hilbert(y)| envelope statistics and scipy.signal.find_peaks count / heights (idx 19-24)
#include <cmath>
#include "kernels.cuh"
#include "fft.cuh"
#include "fft_reg.cuh"
#include "fft20.cuh"

namespace bpc {

// =============================================================================================== k_time_basic
struct TimeBasicSmem {
    float y[kMaxLen];
    unsigned hist[4][2048];
    double sq[kMaxFrames + 8];       // per-256-block sum of squares
    int cz[kMaxFrames + 8];          // per-256-block zero-crossing counts
    float rms[kMaxFrames];
    double zcr[kMaxFrames];
    double dscratch[32];
    float fscratch[32];
    unsigned prefix[4];
    unsigned rank[4];
    unsigned long long bar;
};

// numpy.percentile(method='linear') lerp
__device__ __forceinline__ float lerp_q(float a, float b, double g) {
    const double da = (double)a, db = (double)b;
    return (float)(g >= 0.5 ? db - (db - da) * (1.0 - g) : da + (db - da) * g);
}

// 512 threads: the kernel is latency-bound and shared memory (the staged segment + four histograms) allows two CTAs
// per SM, so the CTA size sets the occupancy (r01 v41: 256 -> 512 threads)
constexpr int kTimeBasicThreads = 512;
template <bool LONG>
__global__ void __launch_bounds__(kTimeBasicThreads) k_time_basic(const float* __restrict__ y, Geometry g, Workspace ws,
                                                    float* scalars, int32_t* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TimeBasicSmem& S = *reinterpret_cast<TimeBasicSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L, T = g.T;
    const float* yb = y + (size_t)b * L;
    // 1 s: the segment is staged in shared memory (TMA bulk copy); long mode reads it from global memory (L2) and keeps
    // the per-block / per-frame arrays in the segment's scratch region
    const float* ys = S.y;
    double* sqv = S.sq; int* czv = S.cz; float* rmsv = S.rms; double* zcrv = S.zcr;
    if (LONG) {
        ys = yb;
        double* d = reinterpret_cast<double*>(ws.scratch + (size_t)b * ws.scratch_stride);
        sqv = d; d += T + 8; zcrv = d; d += T;
        czv = reinterpret_cast<int*>(d); rmsv = reinterpret_cast<float*>(czv + T + 8);
    } else if ((L & 3) == 0 && ((size_t)yb & 15) == 0) {
        stage_segment_tma(S.y, yb, L, (uint64_t*)&S.bar);
    } else {
        for (int i = tid; i < L; i += kTimeBasicThreads) S.y[i] = yb[i];
    }
    __syncthreads();

    // ---- moments (scipy.stats.skew / kurtosis, biased) and input checks
    double s1 = 0.0;
    int bad = 0, nonzero = 0;
    for (int i = tid; i < L; i += kTimeBasicThreads) {
        const float v = ys[i];
        s1 += (double)v;
        bad |= !isfinite(v);
        nonzero |= (v != 0.f);
    }
    s1 = block_sum(s1, S.dscratch);
    bad = __syncthreads_or(bad);
    nonzero = __syncthreads_or(nonzero);
    const double mean = s1 / L;
    double m2 = 0.0, m3 = 0.0, m4 = 0.0;
    for (int i = tid; i < L; i += kTimeBasicThreads) {
        const double d = (double)ys[i] - mean;
        const double d2 = d * d;
        m2 += d2;
        m3 += d2 * d;
        m4 += d2 * d2;
    }
    m2 = block_sum(m2, S.dscratch) / L;
    m3 = block_sum(m3, S.dscratch) / L;
    m4 = block_sum(m4, S.dscratch) / L;
    float* sc = scalars + (size_t)b * g.nscal;
    if (tid == 0) {
        sc[29] = (float)(m3 / (m2 * sqrt(m2)));
        sc[30] = (float)(m4 / (m2 * m2) - 3.0);
        if (status) {
            unsigned f = 0;
            if (bad) f |= BPC_SEG_NONFINITE;
            if (!nonzero) f |= BPC_SEG_SILENT;
            if (f) atomicOr((unsigned int*)&status[b], f);
        }
    }

    // ---- per-256-sample block sums: squares (rms) and zero crossings (zcr)
    const int nblk = (L + 255) / 256;
    for (int j = warp; j < nblk; j += kTimeBasicThreads / 32) {
        double sq = 0.0;
        int cz = 0;
        for (int i = 256 * j + lane; i < 256 * j + 256 && i < L; i += 32) {
            const float v = ys[i];
            sq += (double)__fmul_rn(v, v);
            if (i >= 1) {
                // librosa.zero_crossings: values within +-1e-10 are clipped to +0, then signbit comparison
                const float a = fabsf(v) <= 1e-10f ? 0.f : v;
                const float pv = ys[i - 1];
                const float p = fabsf(pv) <= 1e-10f ? 0.f : pv;
                cz += (signbit(a) != signbit(p)) ? 1 : 0;
            }
        }
        sq = warp_sum(sq);
        cz = warp_sum(cz);
        if (lane == 0) { sqv[j] = sq; czv[j] = cz; }
    }
    __syncthreads();
    // frame t covers y[256 (t-4), 256 (t+4)) (frame_length 2048 centred, hop 256)
    for (int t = tid; t < T; t += kTimeBasicThreads) {
        double sq = 0.0;
        int cz = 0;
        for (int j = t - 4; j < t + 4; ++j)
            if (j >= 0 && j < nblk) { sq += sqv[j]; cz += czv[j]; }
        const int first = 256 * (t - 4);                  // first sample of the frame never counts as a crossing
        if (first >= 1 && first < L) {
            const float v = ys[first], pv = ys[first - 1];
            const float a = fabsf(v) <= 1e-10f ? 0.f : v, p = fabsf(pv) <= 1e-10f ? 0.f : pv;
            cz -= (signbit(a) != signbit(p)) ? 1 : 0;
        }
        rmsv[t] = sqrtf((float)(sq / 2048.0));
        zcrv[t] = (double)cz / 2048.0;
    }
    __syncthreads();
    if (warp < 2) {
        double s = 0.0, q = 0.0, mx = -1e300, mn = 1e300;
        for (int t = lane; t < T; t += 32) {
            const double v = warp == 0 ? (double)rmsv[t] : zcrv[t];
            s += v;
            q += v * v;
            mx = fmax(mx, v);
            mn = fmin(mn, v);
        }
        s = warp_sum(s);
        q = warp_sum(q);
        mx = warp_max(mx);
        mn = warp_min(mn);
        if (lane == 0) {
            const double mu = s / T;
            const double sd = sqrt(fmax(0.0, q / T - mu * mu));
            float* o = sc + (warp == 0 ? 0 : 4);
            o[0] = (float)mu; o[1] = (float)sd; o[2] = (float)mx; o[3] = (float)mn;
        }
    }

    // ---- np.percentile(|y|, 90 / 10): exact order statistics by 3-pass radix select on the float bit patterns
    const double v90 = 0.9 * (double)(L - 1), v10 = 0.1 * (double)(L - 1);
    if (tid == 0) {
        S.rank[0] = (unsigned)floor(v90); S.rank[1] = min((unsigned)floor(v90) + 1u, (unsigned)(L - 1));
        S.rank[2] = (unsigned)floor(v10); S.rank[3] = min((unsigned)floor(v10) + 1u, (unsigned)(L - 1));
        S.prefix[0] = S.prefix[1] = S.prefix[2] = S.prefix[3] = 0u;
    }
    const int shifts[3] = {21, 10, 0};
    const unsigned masks[3] = {0x7ffu, 0x7ffu, 0x3ffu};
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = tid; i < 4 * 2048; i += kTimeBasicThreads) (&S.hist[0][0])[i] = 0u;
        __syncthreads();
        const unsigned p0 = S.prefix[0], p1 = S.prefix[1], p2 = S.prefix[2], p3 = S.prefix[3];
        const unsigned hi_mask = pass == 0 ? 0u : (pass == 1 ? 0xffe00000u : 0xfffffc00u);
        if (pass == 0) {
            // all four targets still share the empty prefix: one histogram serves them (4x fewer shared atomics).  It
            // is kept as four partial histograms (one per warp pair): |y| of a breath sits in a handful of exponent
            // bins, and same-address shared atomics serialise
            for (int i = tid; i < L; i += kTimeBasicThreads)
                atomicAdd(&S.hist[warp & 3][__float_as_uint(fabsf(ys[i])) >> 21], 1u);
            __syncthreads();
            for (int i = tid; i < 2048; i += kTimeBasicThreads) S.hist[0][i] += S.hist[1][i] + S.hist[2][i] + S.hist[3][i];
        } else {
            for (int i = tid; i < L; i += kTimeBasicThreads) {
                const unsigned key = __float_as_uint(fabsf(ys[i]));
                const unsigned hi = key & hi_mask, d = (key >> shifts[pass]) & masks[pass];
                if (hi == p0) atomicAdd(&S.hist[0][d], 1u);
                if (hi == p1) atomicAdd(&S.hist[1][d], 1u);
                if (hi == p2) atomicAdd(&S.hist[2][d], 1u);
                if (hi == p3) atomicAdd(&S.hist[3][d], 1u);
            }
        }
        __syncthreads();
        if (warp < 4) {
            // find the digit whose cumulative count first exceeds rank; 64 bins per lane
            const unsigned* h = S.hist[pass == 0 ? 0 : warp];
            const int nb = (int)masks[pass] + 1, per = nb / 32;
            // each lane sums its `per` consecutive bins starting at a lane-dependent rotation: lane * per alone is a
            // multiple of 32 words, i.e. all 32 lanes on one bank (this loop was 80 % of the kernel's shared wavefronts)
            unsigned local = 0;
            for (int i = 0; i < per; ++i) local += h[lane * per + ((i + lane) & (per - 1))];
            unsigned incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const unsigned excl = incl - local;
            const unsigned r = S.rank[warp];
            const bool mine = r >= excl && r < incl;
            if (mine) {
                unsigned c = excl;
                for (int i = 0; i < per; ++i) {
                    const unsigned hv = h[lane * per + i];
                    if (r < c + hv) {
                        S.prefix[warp] |= ((unsigned)(lane * per + i)) << shifts[pass];
                        S.rank[warp] = r - c;
                        break;
                    }
                    c += hv;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        const float a90 = __uint_as_float(S.prefix[0]), b90 = __uint_as_float(S.prefix[1]);
        const float a10 = __uint_as_float(S.prefix[2]), b10 = __uint_as_float(S.prefix[3]);
        sc[31] = lerp_q(a90, b90, v90 - floor(v90));
        sc[32] = lerp_q(a10, b10, v10 - floor(v10));
    }
}

void launch_time_scalars(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws,
                         float* scalars, int32_t* status, cudaStream_t st);

// ================================================================================================= k_autocorr
// r[k] = sum_n y[n] y[n+k], k < 800 (methods.py:105-112; the reference computes all 2L-1 lags with np.correlate).
// Blocked Wiener-Khinchin in FP64: with A_b = y[1024 b : 1024 b + 1024] zero-padded to 2048 and F_b = rfft(A_b),
//   R[k] = sum_b conj(F_b[k]) * (F_b[k] + (-1)^k F_{b+1}[k]),   r = irfft(R)[0:1024]  (exact linear correlation).
// r01 v8 layout: one 8-warp CTA per segment; every warp transforms one block with team_fft<32> (32 register-resident
// points per lane, one exchange) and leaves its spectrum in its own exchange buffer, then all 256 threads fold the eight
// spectra of the round into thread-private bins of R.  Two rounds cover the 16 blocks; warp 0 runs the inverse.
// (v1..v7: 17 CTA-wide shared-memory radix-4 FFTs in sequence, 90 % of the shared-memory pipe, half of it bank conflicts.)
constexpr int kAcWarps = 4, kAcThreads = 32 * kAcWarps;   // r02: 4 warps x 3 CTAs per SM (168 registers) instead of 8 x 1 (255)
constexpr int kAcBins = (1025 + kAcThreads - 1) / kAcThreads;          // bins of R per thread (5)

struct AutocorrSmem {
    double2 xch[kAcWarps][32 * 33];       // exchange buffers; after a transform: the block's spectrum X[0..1024]
    double dscratch[32];
    int iscratch[32];
};

__global__ void __launch_bounds__(kAcThreads, 3) k_autocorr(const float* __restrict__ y, Geometry g, Tables tb,
                                                            int* __restrict__ ints, float* scalars) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AutocorrSmem& S = *reinterpret_cast<AutocorrSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L;
    const float* yb = y + (size_t)b * L;
    const int partner = (32 - lane) & 31;
    const double2 wp = __ldg(tb.ptw2048 + lane);
    const double2 wl = make_double2(wp.y, -wp.x);                         // -i * exp(-2 pi i lane / 2048)
    const double2* twa = tb.twa1024 + lane;
    double2* xch = S.xch[warp];

    double2 R[kAcBins], carry[kAcBins];
#pragma unroll
    for (int i = 0; i < kAcBins; ++i) { R[i] = make_double2(0.0, 0.0); carry[i] = make_double2(0.0, 0.0); }
    const double sgn = (tid & 1) ? -1.0 : 1.0;                            // (-1)^k for k = tid + 256 i
    const int nblk = (L + 1023) / 1024;
    const int rounds = (nblk + kAcWarps - 1) / kAcWarps;
    // blocks are taken from the top so that the spectrum of block b + 1 is known when block b is folded
    for (int r = 0; r < rounds; ++r) {
        const int base = nblk - kAcWarps * (r + 1);                       // lowest block of this round (may be < 0)
        const int blk = base + warp;
        {
            double2 a[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float2 v = make_float2(0.f, 0.f);
                if (j < 16 && blk >= 0) {
                    const int gi = blk * 1024 + 2 * (lane + 32 * j);     // even; L is even
                    if (gi < L) v = __ldg(reinterpret_cast<const float2*>(yb + gi));
                }
                a[j] = make_double2((double)v.x, (double)v.y);
            }
            team_fft<32>(a, twa, 32, xch, lane);
            // every conjugate pair once (bins 0 .. 1024; the second output of k = 0 is the real Nyquist bin X[1024])
            auto emit = [&](int k, double2 t2) { xch[k] = make_double2(0.5 * t2.x, 0.5 * t2.y); };
            team_rsplit_pairs<32, 0>(a, wl, lane, partner, emit);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kAcBins; ++i) {
            const int k = tid + kAcThreads * i;
            if (k <= 1024) {
                double2 nxt = carry[i];
#pragma unroll
                for (int w = kAcWarps - 1; w >= 0; --w) {
                    const double2 F = S.xch[w][k];
                    // conj(F) * (F + sgn * nxt)
                    const double gx = F.x + sgn * nxt.x, gy = F.y + sgn * nxt.y;
                    R[i].x += F.x * gx + F.y * gy;
                    R[i].y += F.x * gy - F.y * gx;
                    nxt = F;
                }
                carry[i] = nxt;
            }
        }
        __syncthreads();
    }
    // irfft(R, 2048) through one complex FFT-1024: Z[k] = E[k] + i O[k], feed conj(Z) to the forward transform
#pragma unroll
    for (int i = 0; i < kAcBins; ++i) {
        const int k = tid + kAcThreads * i;
        if (k <= 1024) S.xch[0][k] = R[i];
    }
    __syncthreads();
    double* rr = reinterpret_cast<double*>(S.xch[2]);                     // r[0 .. 1023]
    if (warp == 0) {
        const double2* Rs = S.xch[0];
        double2 a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int k = lane + 32 * j;
            const double2 xk = Rs[k], xn = Rs[1024 - k];
            const double2 e = make_double2(0.5 * (xk.x + xn.x), 0.5 * (xk.y - xn.y));
            const double2 d = make_double2(0.5 * (xk.x - xn.x), 0.5 * (xk.y + xn.y));
            const double2 w = __ldg(tb.ptw2048 + k);                       // exp(-i th); need exp(+i th) = conj
            const double2 o = make_double2(d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y);
            a[j] = make_double2(e.x - o.y, -(e.y + o.x));                  // conj(e + i o)
        }
        team_fft<32>(a, twa, 32, S.xch[1], lane);
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) {
            const int m = lane + 32 * k2;                                  // output index: r[2m], r[2m+1]
            if (m < 512) {
                const double2 o = a[bitrev<32>(k2)];
                *reinterpret_cast<double2*>(rr + 2 * m) = make_double2(o.x / 1024.0, -o.y / 1024.0);
            }
        }
    }
    __syncthreads();
    // normalise by r[0]; first index of the minimum over lags < sr // 20 (np.argmin; NaN -> first NaN)
    const double r0 = rr[0];
    const int nl = 800 < L ? 800 : L / 2;
    double best = 1e300;
    int bi = 0x7fffffff;
    for (int k = tid; k < nl; k += kAcThreads) {
        const float v = (float)(rr[k] / r0);
        if ((double)v < best) { best = (double)v; bi = k; }
    }
    // block argmin (value, then smallest index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { S.dscratch[warp] = best; S.iscratch[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kAcWarps; ++w)
            if (S.dscratch[w] < best || (S.dscratch[w] == best && S.iscratch[w] < bi)) { best = S.dscratch[w]; bi = S.iscratch[w]; }
        if (!(r0 == r0) || r0 == 0.0 || bi == 0x7fffffff) bi = 0;          // all-NaN row: argmin returns 0
        float* sc = scalars + (size_t)b * g.nscal;
        sc[33] = (float)(rr[160] / r0);
        sc[34] = (float)(rr[320] / r0);
        sc[35] = (float)((double)bi / 16000.0);
        ints[b * 2 + 1] = bi;
    }
}

// ================================================================================================== k_hilbert
// scipy.signal.hilbert(y) for L = 16000 through two float32 complex FFT-8000 (scipy runs float32 pocketfft on float32
// input):  forward: real-FFT split of y;  G[k] = -i Y[k] (0 < k < 8000), G[0] = G[8000] = 0;  h = irfft(G).
// 8000 = 20 x 20 x 20: three register-resident radix-20 passes each way (fft20.cuh; v0-v27 ran six radix-5 / radix-4
// passes each way, 53 % of whose shared-memory wavefronts were bank conflicts).  The forward transform is decimation in
// frequency (natural in, digit-reversed out), the inverse runs the transposed (decimation-in-time) network on the data
// where it lies, so no reordering pass is needed.
constexpr int kHN = 8000;
constexpr int kHilbertThreads = 416;        // 400 butterflies per pass; 2 CTAs / SM (shared memory) -> 78 registers

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }


template <int R>
__device__ __forceinline__ void dft_small(float2* a) {
    if (R == 4) {
        const float2 b0 = cadd(a[0], a[2]), b1 = csub(a[0], a[2]), b2 = cadd(a[1], a[3]);
        const float2 d = csub(a[1], a[3]);
        const float2 b3 = make_float2(d.y, -d.x);
        a[0] = cadd(b0, b2); a[1] = cadd(b1, b3); a[2] = csub(b0, b2); a[3] = csub(b1, b3);
    } else {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        const float2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
        const float2 m1 = make_float2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
        const float2 m2 = make_float2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
        const float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
        const float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
        a[0] = make_float2(a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y);
        a[1] = make_float2(m1.x + n1.y, m1.y - n1.x);      // m1 - i n1
        a[4] = make_float2(m1.x - n1.y, m1.y + n1.x);      // m1 + i n1
        a[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
        a[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
    }
}

constexpr int kHilbertXBytes = kH20Pitch * (int)sizeof(float2);      // 67200: padded FFT array, later the envelope

struct HilbertTail {
    double dscratch[32];
    float fscratch[32];
    int iscratch[32];
    float best_v;
    int best_i;
};

__global__ void __launch_bounds__(kHilbertThreads, 2) k_hilbert(const float* __restrict__ y, Geometry g, Tables tb,
                                                              Workspace ws, float* scalars) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* X = reinterpret_cast<float2*>(smem_raw);                         // [8400] padded; later env [16000] floats
    HilbertTail& S = *reinterpret_cast<HilbertTail*>(smem_raw + kHilbertXBytes + sizeof(unsigned short) * 16000);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L;                                        // L == 16000 (checked on the host)
    const float* yb = y + (size_t)b * L;
    load_f2_batched<(kHN + kHilbertThreads - 1) / kHilbertThreads>(reinterpret_cast<const float2*>(yb), kHN, kHN, tid,
                                                                   kHilbertThreads, [&](int m, float2 v) { X[h20_pad(m)] = v; });
    __syncthreads();
    const float2* twa = tb.tw20a;                                             // [20][400]: exp(-2 pi i k pos / 8000)
    const float2* twb = tb.tw20b;                                             // [20][20]:  exp(-2 pi i k pos / 400)
    if (tid < kH20Bfly) h20_butterfly<8000, false>(X, twa, tid);
    __syncthreads();
    if (tid < kH20Bfly) h20_butterfly<400, false>(X, twb, tid);
    __syncthreads();
    if (tid < kH20Bfly) h20_butterfly<20, false>(X, nullptr, tid);
    __syncthreads();
    // pairs (k, N-k): real-FFT split -> Y[k], Y[N-k]; G = -i Y; inverse split -> conj(Z), written back in place
    for (int k = tid; k <= kHN / 2; k += kHilbertThreads)
        h20_split_pair(X, k, __ldg(tb.h20pos + k), __ldg(tb.h20pos + kHN - k), __ldg(tb.ptw16000f + k));
    __syncthreads();
    if (tid < kH20Bfly) h20_butterfly<20, true>(X, nullptr, tid);
    __syncthreads();
    if (tid < kH20Bfly) h20_butterfly<400, true>(X, twb, tid);
    __syncthreads();
    if (tid < kH20Bfly) h20_butterfly<8000, true>(X, twa, tid);
    __syncthreads();
    // h[2m] = Re(conj(out[m])) / N, h[2m+1] = Im(conj(out[m])) / N; envelope = |y + i h| (float32 like complex64 abs)
    float* env = reinterpret_cast<float*>(smem_raw);                           // [16000], over X's storage
    {
        constexpr int kPer = (kHN + kHilbertThreads - 1) / kHilbertThreads;   // 20
        float e0[kPer], e1[kPer];
        float xlo = 1.f, xhi = 1.f;                       // range of max(|re|, |im|): checked once per segment (fft.cuh)
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int m = tid + kHilbertThreads * i;
            if (m < kHN) {
                const float2 o = X[h20_pad(m)];
                const float2 v = __ldg(reinterpret_cast<const float2*>(yb) + m);
                const float h0 = o.x * (1.0f / (float)kHN), h1 = -o.y * (1.0f / (float)kHN);
                float xa, xb;
                e0[i] = c64_abs_f32_unchecked(v.x, h0, &xa);                   // np.abs(complex64)
                e1[i] = c64_abs_f32_unchecked(v.y, h1, &xb);
                xlo = fminf(xlo, fminf(xa, xb));
                xhi = fmaxf(xhi, fmaxf(xa, xb));
            }
        }
        // (the barrier the envelope store below needs anyway) a segment with exact zeros -- digital silence, a zero
        // padded tail -- takes the checked form for every sample
        if (__syncthreads_or(!(c64_abs_in_range(xlo) && c64_abs_in_range(xhi)))) {
#pragma unroll                                             // (e0 / e1 must stay registers: no run-time indexing)
            for (int i = 0; i < kPer; ++i) {
                const int m = tid + kHilbertThreads * i;
                if (m < kHN) {
                    const float2 o = X[h20_pad(m)];
                    const float2 v = __ldg(reinterpret_cast<const float2*>(yb) + m);
                    const float h0 = o.x * (1.0f / (float)kHN), h1 = -o.y * (1.0f / (float)kHN);
                    e0[i] = c64_abs_f32(v.x, h0);
                    e1[i] = c64_abs_f32(v.y, h1);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int m = tid + kHilbertThreads * i;
            if (m < kHN) reinterpret_cast<float2*>(env)[m] = make_float2(e0[i], e1[i]);
        }
    }
    __syncthreads();
    double s = 0.0, q = 0.0;
    for (int c = tid; c < L / 4; c += kHilbertThreads) {
        const float4 v = reinterpret_cast<const float4*>(env)[c];
        const double a = (double)v.x, b2 = (double)v.y, c2 = (double)v.z, d = (double)v.w;
        s += (a + b2) + (c2 + d);
        q += fma(a, a, b2 * b2) + fma(c2, c2, d * d);
    }
    s = block_sum(s, S.dscratch);
    q = block_sum(q, S.dscratch);
    const float emean = (float)(s / L);
    const float estd = (float)sqrt(fmax(0.0, q / L - (s / L) * (s / L)));
    // scipy.signal.find_peaks(env, height=emean, distance=1600): local maxima (plateau mid-points), height filter.
    // Candidates are compacted into a list (r01 v3: the selection rounds rescanned all 16000 flags, 20 % of the kernel).
    unsigned short* clist = reinterpret_cast<unsigned short*>(smem_raw + kHilbertXBytes);          // after X / env
    constexpr int kMaxList = 16000;                                            // 32 KB: every sample could be listed
    constexpr int kGone = 0xffff;
    if (tid == 0) S.best_i = 0;                                                // list length
    __syncthreads();
    // four samples per thread and step; only rising edges that do not rise further (strict peaks and plateau starts)
    // take the general path
    for (int c = tid; c < L / 4; c += kHilbertThreads) {
        const float4 v = reinterpret_cast<const float4*>(env)[c];
        const int i0 = 4 * c;
        const float lft = i0 > 0 ? env[i0 - 1] : FLT_MAX, rgt = i0 + 4 < L ? env[i0 + 4] : FLT_MAX;
        unsigned m = (unsigned)(lft < v.x && v.x >= v.y) | ((unsigned)(v.x < v.y && v.y >= v.z) << 1) |
                     ((unsigned)(v.y < v.z && v.z >= v.w) << 2) | ((unsigned)(v.z < v.w && v.w >= rgt) << 3);
        while (m) {
            const int i = i0 + __ffs(m) - 1;
            m &= m - 1;
            const float e = env[i];
            int ahead = i + 1;
            while (ahead < L - 1 && env[ahead] == e) ++ahead;
            if (env[ahead] < e) {
                const int mid = (i + ahead - 1) / 2;
                if (env[mid] >= emean) {
                    const int slot = atomicAdd(&S.best_i, 1);
                    if (slot < kMaxList) clist[slot] = (unsigned short)mid;
                }
            }
        }
    }
    __syncthreads();
    const int nc = min(S.best_i, kMaxList);
    __syncthreads();
    int n_peaks = 0;
    double hs = 0.0, hq = 0.0;
    for (int iter = 0; iter < 64; ++iter) {
        // highest-priority remaining candidate (ties: later position first, as a stable ascending argsort would)
        float bv = -1.f;
        int bi = -1;
        for (int j = tid; j < nc; j += kHilbertThreads) {
            const int i = clist[j];
            if (i != kGone && (env[i] > bv || (env[i] == bv && i > bi))) { bv = env[i]; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > bv || (ob == bv && oi > bi)) { bv = ob; bi = oi; }
        }
        if (lane == 0) { S.fscratch[warp] = bv; S.iscratch[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kHilbertThreads / 32; ++w)
                if (S.fscratch[w] > bv || (S.fscratch[w] == bv && S.iscratch[w] > bi)) { bv = S.fscratch[w]; bi = S.iscratch[w]; }
            S.best_v = bv;
            S.best_i = bi;
        }
        __syncthreads();
        const int pi = S.best_i;
        if (pi < 0) break;
        const float pv = S.best_v;
        ++n_peaks;
        hs += (double)pv;
        hq += (double)pv * (double)pv;
        for (int j = tid; j < nc; j += kHilbertThreads) {
            const int i = clist[j];
            if (i != kGone && i > pi - 1600 && i < pi + 1600) clist[j] = (unsigned short)kGone;
        }
        __syncthreads();
    }
    if (tid == 0) {
        float* sc = scalars + (size_t)b * g.nscal;
        sc[19] = emean;
        sc[20] = estd;
        sc[21] = __fdiv_rn(emean, __fadd_rn(estd, 1e-8f));
        sc[22] = (float)n_peaks;
        const double hm = n_peaks > 0 ? hs / n_peaks : 0.0;
        sc[23] = (float)hm;
        sc[24] = n_peaks > 1 ? (float)sqrt(fmax(0.0, hq / n_peaks - hm * hm)) : 0.f;
        ws.ints[b * 2 + 0] = n_peaks;
    }
}

// =============================================================================================== k_hilbert_long
// Long mode (L = 16000 d): the same algorithm with the transform in the segment's global scratch region (L2-resident)
// and a run-time radix plan: N = L / 2 = 2^a 3^b 5^c is walked with radix-5, radix-3, radix-4 and at most one radix-2
// pass (decimation in frequency forward, the transposed decimation-in-time network for the inverse, so the spectrum
// is processed where it lies, in digit-reversed order).  float32 throughout, like scipy's transform of float32 input.
struct HilbertPlan {
    int n;                 // N = L / 2
    int npass;
    int radix[16];
};

__device__ __forceinline__ void dft2(float2* a) {
    const float2 t = a[0];
    a[0] = cadd(t, a[1]);
    a[1] = csub(t, a[1]);
}
__device__ __forceinline__ void dft3(float2* a) {
    const float c = -0.5f, sn = 0.86602540378443864676f;
    const float2 t1 = cadd(a[1], a[2]), t2 = csub(a[1], a[2]);
    const float2 m = make_float2(a[0].x + c * t1.x, a[0].y + c * t1.y);
    const float2 n = make_float2(sn * t2.x, sn * t2.y);
    a[0] = cadd(a[0], t1);
    a[1] = make_float2(m.x + n.y, m.y - n.x);              // m - i n
    a[2] = make_float2(m.x - n.y, m.y + n.x);              // m + i n
}

template <int R, bool kDit>
__device__ __forceinline__ void long_pass(float2* x, int N, int span, const float2* __restrict__ tw, int tid, int nthr) {
    const int q = span / R, ts = N / span;
    for (int j = tid; j < N / R; j += nthr) {
        const int blk = j / q, pos = j - blk * q;
        const int base = blk * span + pos;
        float2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = x[base + r * q];
        if (kDit && pos != 0) {
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], __ldg(tw + (size_t)r * pos * ts));
        }
        if (R == 2) dft2(a);
        else if (R == 3) dft3(a);
        else dft_small<R>(a);
        if (!kDit && pos != 0) {
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], __ldg(tw + (size_t)r * pos * ts));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) x[base + r * q] = a[r];
    }
    __syncthreads();
}

template <bool kDit>
__device__ __forceinline__ void long_pass_rt(float2* x, int N, int R, int span, const float2* tw, int tid, int nthr) {
    switch (R) {
        case 5: long_pass<5, kDit>(x, N, span, tw, tid, nthr); break;
        case 4: long_pass<4, kDit>(x, N, span, tw, tid, nthr); break;
        case 3: long_pass<3, kDit>(x, N, span, tw, tid, nthr); break;
        default: long_pass<2, kDit>(x, N, span, tw, tid, nthr); break;
    }
}

// position of output bin k after the DIF passes of the plan
__device__ __forceinline__ int plan_pos(const HilbertPlan& P, int k) {
    int p = 0, s = P.n;
    for (int i = 0; i < P.npass - 1; ++i) {
        const int r = P.radix[i];
        const int d = k % r;
        k /= r; s /= r;
        p += d * s;
    }
    return p + k;
}

constexpr int kHilbertLongThreads = 1024;
__global__ void __launch_bounds__(kHilbertLongThreads) k_hilbert_long(const float* __restrict__ y, Geometry g, Tables tb,
                                                                  Workspace ws, HilbertPlan P, float* scalars) {
    __shared__ HilbertTail S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = kHilbertLongThreads;
    const int b = blockIdx.x, L = g.L, N = P.n;
    const float* yb = y + (size_t)b * L;
    float* base = ws.scratch + (size_t)b * ws.scratch_stride;
    float2* X = reinterpret_cast<float2*>(base);                               // [N]; later env [L] floats in place
    int* clist = reinterpret_cast<int*>(base + L);                             // [L / 2] candidate peaks
    for (int m = tid; m < N; m += NT) X[m] = __ldg(reinterpret_cast<const float2*>(yb) + m);
    __syncthreads();
    const float2* tw = tb.tw_long;
    {
        int span = N;
        for (int i = 0; i < P.npass; ++i) { long_pass_rt<false>(X, N, P.radix[i], span, tw, tid, NT); span /= P.radix[i]; }
    }
    // same pair processing as k_hilbert: split, G = -i Y on 0 < k < N, inverse split, conj for the forward-as-inverse trick
    for (int k = tid; k <= N / 2; k += NT) {
        const int kn = (N - k) % N;
        const int pk = plan_pos(P, k), pn = plan_pos(P, kn);
        const float2 zk = X[pk], zn = X[pn];
        const float2 w = __ldg(tb.ptw_long + k);                               // exp(-2 pi i k / L)
        const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
        const float2 wo = cmul(w, o);
        float2 yk = cadd(e, wo);
        float2 yn = cconj(csub(e, wo));
        float2 gk = make_float2(yk.y, -yk.x), gn = make_float2(yn.y, -yn.x);
        if (k == 0) { gk = make_float2(0.f, 0.f); gn = make_float2(0.f, 0.f); }
        const float2 e2 = make_float2(0.5f * (gk.x + gn.x), 0.5f * (gk.y - gn.y));
        const float2 d2 = make_float2(0.5f * (gk.x - gn.x), 0.5f * (gk.y + gn.y));
        const float2 o2 = cmul(d2, cconj(w));
        const float2 z = make_float2(e2.x - o2.y, e2.y + o2.x);
        const float2 dn = make_float2(0.5f * (gn.x - gk.x), 0.5f * (gn.y + gk.y));
        const float2 on = cmul(dn, make_float2(-w.x, -w.y));
        const float2 zn2 = make_float2(e2.x - on.y, -e2.y + on.x);
        X[pk] = cconj(z);
        if (kn != k && k != 0) X[pn] = cconj(zn2);
    }
    __syncthreads();
    {
        int span = 1;
        for (int i = P.npass - 1; i >= 0; --i) { span *= P.radix[i]; long_pass_rt<true>(X, N, P.radix[i], span, tw, tid, NT); }
    }
    float* env = base;
    const float inv_n = 1.0f / (float)N;
    for (int m = tid; m < N; m += NT) {                                         // in place: X[m] -> env[2m], env[2m+1]
        const float2 o = X[m];
        const float2 v = __ldg(reinterpret_cast<const float2*>(yb) + m);
        const float h0 = o.x * inv_n, h1 = -o.y * inv_n;
        const float e0 = (float)sqrt((double)v.x * (double)v.x + (double)h0 * (double)h0);
        const float e1 = (float)sqrt((double)v.y * (double)v.y + (double)h1 * (double)h1);
        env[2 * m] = e0;
        env[2 * m + 1] = e1;
    }
    __syncthreads();
    double s = 0.0, q = 0.0;
    for (int i = tid; i < L; i += NT) { const double v = (double)env[i]; s += v; q += v * v; }
    s = block_sum(s, S.dscratch);
    q = block_sum(q, S.dscratch);
    const float emean = (float)(s / L);
    const float estd = (float)sqrt(fmax(0.0, q / L - (s / L) * (s / L)));
    constexpr int kGone = -1;
    const int max_list = L / 2;
    if (tid == 0) S.best_i = 0;
    __syncthreads();
    for (int i = tid + 1; i < L - 1; i += NT) {
        if (env[i - 1] < env[i]) {
            int ahead = i + 1;
            while (ahead < L - 1 && env[ahead] == env[i]) ++ahead;
            if (env[ahead] < env[i]) {
                const int mid = (i + ahead - 1) / 2;
                if (env[mid] >= emean) {
                    const int slot = atomicAdd(&S.best_i, 1);
                    if (slot < max_list) clist[slot] = mid;
                }
            }
        }
    }
    __syncthreads();
    const int nc = min(S.best_i, max_list);
    __syncthreads();
    int n_peaks = 0;
    double hs = 0.0, hq = 0.0;
    const int max_iter = L / 1600 + 2;                                          // peaks are >= 1600 samples apart
    for (int iter = 0; iter < max_iter; ++iter) {
        float bv = -1.f;
        int bi = -1;
        for (int j = tid; j < nc; j += NT) {
            const int i = clist[j];
            if (i != kGone && (env[i] > bv || (env[i] == bv && i > bi))) { bv = env[i]; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > bv || (ob == bv && oi > bi)) { bv = ob; bi = oi; }
        }
        if (lane == 0) { S.fscratch[warp] = bv; S.iscratch[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NT / 32; ++w)
                if (S.fscratch[w] > bv || (S.fscratch[w] == bv && S.iscratch[w] > bi)) { bv = S.fscratch[w]; bi = S.iscratch[w]; }
            S.best_v = bv;
            S.best_i = bi;
        }
        __syncthreads();
        const int pi = S.best_i;
        if (pi < 0) break;
        const float pv = S.best_v;
        ++n_peaks;
        hs += (double)pv;
        hq += (double)pv * (double)pv;
        for (int j = tid; j < nc; j += NT) {
            const int i = clist[j];
            if (i != kGone && i > pi - 1600 && i < pi + 1600) clist[j] = kGone;
        }
        __syncthreads();
    }
    if (tid == 0) {
        float* sc = scalars + (size_t)b * g.nscal;
        sc[19] = emean;
        sc[20] = estd;
        sc[21] = __fdiv_rn(emean, __fadd_rn(estd, 1e-8f));
        sc[22] = (float)n_peaks;
        const double hm = n_peaks > 0 ? hs / n_peaks : 0.0;
        sc[23] = (float)hm;
        sc[24] = n_peaks > 1 ? (float)sqrt(fmax(0.0, hq / n_peaks - hm * hm)) : 0.f;
        ws.ints[b * 2 + 0] = n_peaks;
    }
}

// ==================================================================================================== launchers
void launch_time_scalars(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws,
                         float* scalars, int32_t* status, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_time_basic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TimeBasicSmem));
        cudaFuncSetAttribute(k_time_basic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TimeBasicSmem));
        cudaFuncSetAttribute(k_autocorr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AutocorrSmem));
    });
    if (g.long_mode) k_time_basic<true><<<n, kTimeBasicThreads, sizeof(TimeBasicSmem), st>>>(y, g, ws, scalars, status);
    else k_time_basic<false><<<n, kTimeBasicThreads, sizeof(TimeBasicSmem), st>>>(y, g, ws, scalars, status);
    k_autocorr<<<n, kAcThreads, sizeof(AutocorrSmem), st>>>(y, g, tb, ws.ints, scalars);
    note_launch(2);
}

void launch_hilbert(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* scalars,
                    cudaStream_t st) {
    if (g.long_mode) {
        HilbertPlan P{};
        P.n = g.L / 2;
        int r = P.n;
        for (int f : {5, 3}) while (r % f == 0) { P.radix[P.npass++] = f; r /= f; }
        while (r % 4 == 0) { P.radix[P.npass++] = 4; r /= 4; }
        if (r == 2) { P.radix[P.npass++] = 2; r = 1; }
        k_hilbert_long<<<n, kHilbertLongThreads, 0, st>>>(y, g, tb, ws, P, scalars);
        note_launch();
        return;
    }
    const int bytes = (int)(kHilbertXBytes + sizeof(unsigned short) * 16000 + sizeof(HilbertTail));
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_hilbert, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    });
    k_hilbert<<<n, kHilbertThreads, bytes, st>>>(y, g, tb, ws, scalars);
    note_launch();
}

}  // namespace bpc
