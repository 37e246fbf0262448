// LPC channel (process.py:64-67, methods.py:116-134): pre-emphasis 0.97, 25 ms / 10 ms Hamming frames, Burg's method
// of order 12 as in librosa.lpc (float64), whole-array z-score over ALL frames, first T frames kept, rows padded.
// One CTA per segment, one warp per frame; forward / backward prediction errors live in shared memory.
#include <cmath>
#include "kernels.cuh"

namespace bpc {

constexpr int kLpcFrame = 400, kLpcShift = 160, kLpcOrder = 12, kLpcMaxFrames = 112;
constexpr int kLpcThreads = 224;                     // 7 warps: the 98 frames of a 1 s segment are 14 full rounds
constexpr int kLpcPer = 13;                            // samples per lane: 13 * 31 = 403 >= 400

struct LpcSmem {
    float coef[kLpcOrder * kLpcMaxFrames];
    double dscratch[32];
    float fscratch[32];
};

// Burg iteration I on register-resident error signals (r01 v2: v1 kept them in shared memory and was bound by its
// 94 %-busy load/store pipe).  Lane l owns samples n = 13 l + q.  B[q] = bwd[n]; the forward error that pairs with it,
// fwd[n + 1 + I], sits in F[(q + I) % 13]: the per-iteration shift fwd = fwd[1:] is a renaming plus one element handed
// down from the next lane.
// a / b to within an ulp (b > 0, normal): hardware reciprocal seed (2^-23), two Newton steps, one residual correction.
// The IEEE division sequence with its slow-path check was 27 % of this kernel's instructions (one per Burg order).
__device__ __forceinline__ double div_fast(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(r, fma(-b, q, a), q);
}

template <int I>
__device__ __forceinline__ void burg_step(double (&F)[kLpcPer], double (&B)[kLpcPer], double& a_lane, double& den,
                                          int lane) {
    constexpr int len = kLpcFrame - 1 - I;              // pairs n < len are live
    const double eps = 2.2250738585072014e-308;         // util.tiny(float64)
    // No per-element masks: every backward error with n >= len is zero (dropped below) and every forward slot beyond
    // sample 399 is zero, so dead pairs contribute 0 to the sums and stay 0 under the update.
    double num = 0.0;
#pragma unroll
    for (int q = 0; q < kLpcPer; ++q) num = fma(B[q], F[(q + I) % kLpcPer], num);
    num = warp_sum(num);
    const double k = div_fast(num * -2.0, den + eps);
    // Levinson update a[j] = a_prev[j] + k * a_prev[I - j + 1], j = 1 .. I + 1; lane j holds a[j] (a[0] = 1, rest 0)
    {
        const int src = I + 1 - lane;
        const double mirror = __shfl_sync(0xffffffffu, a_lane, src & 31);
        if (lane >= 1 && lane <= I + 1) a_lane = a_lane + k * mirror;
    }
#pragma unroll
    for (int q = 0; q < kLpcPer; ++q) {
        const double f = F[(q + I) % kLpcPer], bw = B[q];
        F[(q + I) % kLpcPer] = fma(k, bw, f);
        B[q] = fma(k, f, bw);
    }
    // den = (1 - k^2) den - bwd[-1]^2 - fwd[0]^2 with the updated errors
    const double f0 = __shfl_sync(0xffffffffu, F[I % kLpcPer], 0);
    const double bl = __shfl_sync(0xffffffffu, B[(len - 1) % kLpcPer], (len - 1) / kLpcPer);
    den = (1.0 - k * k) * den - bl * bl - f0 * f0;
    // bwd = bwd[:-1]: the last live backward error is dropped
    if (lane == (len - 1) / kLpcPer) B[(len - 1) % kLpcPer] = 0.0;
    // fwd = fwd[1:]: the slot of this lane's first element receives the next lane's first element
    F[I % kLpcPer] = __shfl_down_sync(0xffffffffu, F[I % kLpcPer], 1);
}

template <int I>
__device__ __forceinline__ void burg_all(double (&F)[kLpcPer], double (&B)[kLpcPer], double& a_lane, double& den,
                                         int lane) {
    if constexpr (I < kLpcOrder) {
        burg_step<I>(F, B, a_lane, den, lane);
        burg_all<I + 1>(F, B, a_lane, den, lane);
    }
}

// phase 0: everything (1 s).  Long mode: phase 1 = the Burg frames of this CTA's share (grid (segment, part)) into the
// scratch region, phase 2 = statistics + plane (grid (segment)).
template <bool LONG>
__global__ void __launch_bounds__(kLpcThreads, 3) k_lpc(const float* __restrict__ y, Geometry g, Tables tb, Workspace ws,
                                                float* feats, int phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LpcSmem& S = *reinterpret_cast<LpcSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L, T = g.T, F_ = g.lpc_frames;
    const float* yb = y + (size_t)b * L;
    // [12, F] coefficients: shared memory (1 s: F = 98), the segment's global scratch region in long mode
    float* coef = LONG ? ws.scratch + (size_t)b * ws.scratch_stride : S.coef;

    // The trip count depends on blockIdx only and every warp runs every iteration (warps past the last frame redo it
    // and drop the result): the compiler can then prove that the ~200 shuffles per frame are convergent.  With
    // `fr = warp; fr < F; fr += 4` each of them was a WARPSYNC.COLLECTIVE / ENDCOLLECTIVE pair.
    const int frb = LONG ? blockIdx.y * (kLpcThreads / 32) : 0;
    const int frs = LONG ? gridDim.y * (kLpcThreads / 32) : kLpcThreads / 32;
    for (int fr_base = frb; fr_base < F_ && phase != 2; fr_base += frs) {
        const bool fr_valid = fr_base + warp < F_;
        const int fr = fr_valid ? fr_base + warp : F_ - 1;
        const int start = fr * kLpcShift;
        double Bv[kLpcPer], Fv[kLpcPer];
#pragma unroll
        for (int q = 0; q < kLpcPer; ++q) {
            const int n = kLpcPer * lane + q, gi = start + n;
            double x = 0.0;
            if (n < kLpcFrame) {
                // y_emph = append(y[0], y[1:] - 0.97 * y[:-1])   (float32)
                const float e = gi == 0 ? __ldg(yb) : __fsub_rn(__ldg(yb + gi), __fmul_rn(0.97f, __ldg(yb + gi - 1)));
                x = (double)e * __ldg(tb.hamming400 + n);
            }
            Bv[q] = x;
        }
        // fwd[n] pairs with bwd[n]: F[q] = x[n + 1]
        const double nxt = __shfl_down_sync(0xffffffffu, Bv[0], 1);
#pragma unroll
        for (int q = 0; q < kLpcPer - 1; ++q) Fv[q] = Bv[q + 1];
        Fv[kLpcPer - 1] = lane < 31 ? nxt : 0.0;
        // bwd = x[:-1]: sample 399 is never a backward error
        if (lane == (kLpcFrame - 1) / kLpcPer) Bv[(kLpcFrame - 1) % kLpcPer] = 0.0;
        double den = 0.0;
#pragma unroll
        for (int q = 0; q < kLpcPer; ++q) den = fma(Fv[q], Fv[q], fma(Bv[q], Bv[q], den));
        den = warp_sum(den);
        double a_lane = lane == 0 ? 1.0 : 0.0;
        burg_all<0>(Fv, Bv, a_lane, den, lane);
        if (fr_valid && lane >= 1 && lane <= kLpcOrder) coef[(lane - 1) * F_ + fr] = (float)a_lane;
    }
    if (LONG && phase == 1) return;
    __syncthreads();
    const int F = F_;
    if (ws.dbg_lpc) {
        float* d = ws.dbg_lpc + (size_t)b * kLpcOrder * F;
        for (int i = tid; i < kLpcOrder * F; i += kLpcThreads) d[i] = coef[i];
    }
    // whole-array z over all F frames (process.py:65); pad_time keeps the first T columns; pad value = min of those
    double s = 0.0, q = 0.0;
    for (int i = tid; i < kLpcOrder * F; i += kLpcThreads) { const double v = (double)coef[i]; s += v; q += v * v; }
    s = block_sum(s, S.dscratch);
    q = block_sum(q, S.dscratch);
    const ZTerm z = make_zterm(s, q, (double)(kLpcOrder * F));
    const int Tk = T < F ? T : F;
    float mn = FLT_MAX;
    for (int i = tid; i < kLpcOrder * Tk; i += kLpcThreads) {
        const int c = i / Tk, t = i - c * Tk;
        mn = fminf(mn, z(coef[c * F + t]));
    }
    mn = block_min(mn, S.fscratch);
    float* o = plane_ptr(feats, b, BPC_CH_LPC, T);
    const bool stats = !LONG && ws.stats_acc != nullptr;
    StatAcc al;
    al.init();
    for (int i = tid; i < kPlaneRows * T; i += kLpcThreads) {
        const int c = i / T, t = i - c * T;
        const bool live = c < kLpcOrder && t < Tk;
        const float v = live ? z(coef[c * F + t]) : mn;
        o[i] = v;
        if (stats && live) al.add(v);
    }
    if (stats) {
        if (tid == 0) al.add_n(mn, kPlaneRows * T - kLpcOrder * Tk);
        stat_flush_block(al, ws.stats_acc + 5 * BPC_CH_LPC, S.dscratch, S.fscratch);
    }
}

void launch_lpc(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_lpc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpcSmem));
        cudaFuncSetAttribute(k_lpc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpcSmem));
    });
    if (g.long_mode) {
        k_lpc<true><<<dim3(n, 16), kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 1);
        k_lpc<true><<<dim3(n, 1), kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 2);
        note_launch();
    } else {
        k_lpc<false><<<n, kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 0);
    }
    note_launch();
}

// ------------------------------------------------------------------------------- dataset-level statistics + padding
// acc layout: [(9 + nscal)][5] doubles = {count, sum, sumsq, min, max}.  One CTA per segment, one WARP per plane (warp 9:
// the scalars): only warp shuffles, no block barriers; lane 0 of each warp issues the five atomics of its plane.
constexpr int kStatsThreads = 288;         // nine warps: one per plane

__global__ void __launch_bounds__(kStatsThreads) k_stats(Geometry g, const float* __restrict__ feats,
                                                         double* __restrict__ acc) {
    const int b = blockIdx.x, c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        // rows that carry data (api.cu::kLiveRows); rows live..127 of a plane repeat one pad value (pad_freq), which is
        // accounted for analytically instead of being read back: 772 of 1152 rows cross HBM
        constexpr int kLive[9] = {24, 64, 12, 128, 128, 128, 120, 40, 128};
        const int live = kLive[c];
        const int NP = live * g.T;                             // a multiple of 4 (live is), planes are 16-byte aligned
        const float* plane = feats + ((size_t)b * 9 + c) * kPlaneRows * g.T;
        const float4* p4 = reinterpret_cast<const float4*>(plane);
        double s = 0.0, q = 0.0;
        float mn = FLT_MAX, mx = -FLT_MAX;
        int cnt = 0;
#pragma unroll 8
        for (int i = lane; i < NP / 4; i += 32) {
            const float4 v4 = __ldg(p4 + i);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if ((__float_as_uint(v[e]) & 0x7f800000u) != 0x7f800000u) {
                    const double d = (double)v[e];
                    s += d;
                    q = fma(d, d, q);
                    mn = fminf(mn, v[e]);
                    mx = fmaxf(mx, v[e]);
                    ++cnt;
                }
            }
        }
        if (lane == 0 && live < kPlaneRows) {
            const float pv = __ldg(plane + NP);
            if ((__float_as_uint(pv) & 0x7f800000u) != 0x7f800000u) {
                const int m = (kPlaneRows - live) * g.T;
                s += (double)m * (double)pv;
                q += (double)m * ((double)pv * (double)pv);
                mn = fminf(mn, pv);
                mx = fmaxf(mx, pv);
                cnt += m;
            }
        }
        s = warp_sum(s);
        q = warp_sum(q);
        cnt = warp_sum(cnt);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0 && cnt > 0) {
            double* a = acc + (size_t)c * 5;
            atomicAdd(a + 0, (double)cnt);
            atomicAdd(a + 1, s);
            atomicAdd(a + 2, q);
            atomic_min_double(a + 3, (double)mn);
            atomic_max_double(a + 4, (double)mx);
        }
    }
}

// Statistics of the scalar vectors: thread (group, i) walks scalar i of every kStatGroups-th segment of this CTA's
// share, the groups are combined in shared memory, and each CTA issues five atomics per scalar.  (Up to v36 every
// segment issued them - 737 k atomics on 180 addresses per 4096-segment step, which took longer (0.14 ms) than the
// 0.8 GB the plane warps read.)
constexpr int kStatScalThreads = 256;
__global__ void __launch_bounds__(kStatScalThreads) k_stats_scalars(Geometry g, const float* __restrict__ scalars, int n,
                                                                    double* __restrict__ acc) {
    __shared__ double sh_s[kStatScalThreads], sh_q[kStatScalThreads];
    __shared__ float sh_mn[kStatScalThreads], sh_mx[kStatScalThreads];
    __shared__ int sh_c[kStatScalThreads];
    const int S = g.nscal, G = kStatScalThreads / S;            // S <= 64 (checked on the host)
    const int tid = threadIdx.x, grp = tid / S, i = tid - grp * S;
    double s = 0.0, q = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    int cnt = 0;
    if (grp < G) {
        for (int b = blockIdx.x * G + grp; b < n; b += gridDim.x * G) {
            const float v = __ldg(scalars + (size_t)b * S + i);
            if ((__float_as_uint(v) & 0x7f800000u) != 0x7f800000u) {
                const double d = (double)v;
                s += d;
                q = fma(d, d, q);
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
                ++cnt;
            }
        }
    }
    sh_s[tid] = s; sh_q[tid] = q; sh_mn[tid] = mn; sh_mx[tid] = mx; sh_c[tid] = cnt;
    __syncthreads();
    if (tid < S) {
        for (int k = 1; k < G; ++k) {
            const int o = k * S + tid;
            s += sh_s[o]; q += sh_q[o]; mn = fminf(mn, sh_mn[o]); mx = fmaxf(mx, sh_mx[o]); cnt += sh_c[o];
        }
        if (cnt > 0) {
            double* a = acc + (size_t)(9 + tid) * 5;
            atomicAdd(a + 0, (double)cnt);
            atomicAdd(a + 1, s);
            atomicAdd(a + 2, q);
            atomic_min_double(a + 3, (double)mn);
            atomic_max_double(a + 4, (double)mx);
        }
    }
}

void launch_stats(int n, const Geometry& g, const float* feats, const float* scalars, double* acc, bool planes,
                  cudaStream_t st) {
    if (planes) {                                              // fused mode: the producers did the planes
        k_stats<<<n, kStatsThreads, 0, st>>>(g, feats, acc);
        note_launch();
    }
    const int G = kStatScalThreads / g.nscal;
    int grid = (n + G - 1) / G;
    if (grid > 148) grid = 148;
    k_stats_scalars<<<grid, kStatScalThreads, 0, st>>>(g, scalars, n, acc);
    note_launch();
}

__global__ void k_pad_scalars(Geometry g, float* scalars, int n) {
    const int extra = g.nscal - BPC_NUM_SCALARS;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * extra) scalars[(size_t)(i / extra) * g.nscal + BPC_NUM_SCALARS + i % extra] = 0.f;
}

void launch_pad_scalars(int n, const Geometry& g, float* scalars, cudaStream_t st) {
    const int extra = g.nscal - BPC_NUM_SCALARS;
    if (extra <= 0) return;
    k_pad_scalars<<<(n * extra + 255) / 256, 256, 0, st>>>(g, scalars, n);
    note_launch();
}

}  // namespace bpc
