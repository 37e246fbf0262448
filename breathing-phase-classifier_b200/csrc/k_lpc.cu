// LPC channel (process.py:64-67, methods.py:116-134): pre-emphasis 0.97, 25 ms / 10 ms Hamming frames, Burg's method
// of order 12 as in librosa.lpc (float64), whole-array z-score over ALL frames, first T frames kept, rows padded.
// One CTA per segment, one warp per frame; forward / backward prediction errors live in shared memory.
#include <cmath>
#include <cstdlib>
#include <type_traits>
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

// Burg's method on one frame, one warp: lane j returns a[j] (j <= 12).  All 32 lanes must call it.
__device__ __forceinline__ double burg_frame(const float* __restrict__ yb, int start, const double* __restrict__ hamming,
                                             int lane) {
    double Bv[kLpcPer], Fv[kLpcPer];
#pragma unroll
    for (int q = 0; q < kLpcPer; ++q) {
        const int n = kLpcPer * lane + q, gi = start + n;
        double x = 0.0;
        if (n < kLpcFrame) {
            // y_emph = append(y[0], y[1:] - 0.97 * y[:-1])   (float32)
            const float e = gi == 0 ? __ldg(yb) : __fsub_rn(__ldg(yb + gi), __fmul_rn(0.97f, __ldg(yb + gi - 1)));
            x = (double)e * __ldg(hamming + n);
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
    return a_lane;
}

// ---------------------------------------------------------------------------------------------- k_lpc_fast (r02-i)
// Burg's reflection coefficients from lag products instead of three passes over the error signals per order.
// With f_i[n] = sum_j a_j x[n-j] and b_i[n-1] = sum_j a_j x[n-1-i+j] (a = the order-i predictor, a_0 = 1) the
// numerator of librosa's recursion is a quadratic form in a,
//     sum_{n=i+1}^{N-1} f_i[n] b_i[n-1] = sum_{j,l<=i} a_j a_l Phi[j][i+1-l]  +  (the terms n = i+1 .. M-1, explicit),
//     Phi[u][v] = sum_{n=M}^{N-1} x[n-u] x[n-v],   Phi[u+1][v+1] = Phi[u][v] + x[M-1-u] x[M-1-v] - x[N-1-u] x[N-1-v],
// so one pass over the frame (13 lag sums, Phi[0][.]) and O(M^3) scalar work replace 3 M passes: 5.4 k + ~2.5 k FP64
// operations per frame against 29 k.  The denominator follows librosa's own recursion
// den <- (1 - k^2) den - b_{i+1}[N-1]^2 - f_{i+1}[i+1]^2 with the two edge errors evaluated from the frame's first and
// last 13 samples.  Same reflection coefficients in exact arithmetic; in floating point the quadratic form cancels
// where the direct sums do not, by a factor ~ den_0 / den_i (the prediction gain): measured on 1100 frames (real
// fixtures, tones, chirps, 30 Hz low-passed noise) against oracle/librosa_shim's lpc the coefficients differ by
// <= 3e-9 wherever min_i den_i / den_0 > 1e-4, 2e-7 down to 1e-6, 6e-4 at 1e-10.  Frames below
// kLfGainFloor (6 % of the real fixture frames, none of the synthetic bench frames) are therefore queued and redone by
// the direct method (burg_frame, one warp per frame: k_lpc_redo), so every output is either the direct recursion or
// within 3e-9 of it -- two orders below the float32 rounding of the stored coefficients.
// One THREAD per frame: the 13 lag sums are 13 independent FMA chains over a register ring of the last 13
// samples (FP64-pipe bound, 82 % of the instructions of the pass are DFMA), the order recursion is fully unrolled so
// that Phi (91 doubles), a, the head and the tail samples are statically indexed registers.
constexpr int kLfThreads = 96, kLfM = kLpcOrder, kLfPhi = (kLfM + 1) * (kLfM + 2) / 2;
constexpr double kLfGainFloor = 1e-4;
__host__ __device__ constexpr int lf_idx(int u, int v) {        // u <= v
    return u * (kLfM + 1) - u * (u - 1) / 2 + (v - u);
}
static_assert(lf_idx(kLfM, kLfM) == kLfPhi - 1, "upper triangle, row major");
__constant__ double c_hamming400[kLpcFrame];

void upload_lpc_constants(const double* hamming400) {
    cudaMemcpyToSymbol(c_hamming400, hamming400, sizeof(double) * kLpcFrame);
}

struct LfState {
    double a[kLfM + 1];
    double head[kLfM + 1];      // x[0 .. 12]
    double tl[kLfM + 1];        // x[N-1-k], k = 0 .. 12
    double c[kLfM + 1];         // Phi[0][v]: the 13 lag sums
    double den, den0;
    bool redo;
};

// Phi[u][v] (symmetric): row 0 in registers, rows 1 .. 12 in shared memory, one column of 78 doubles per thread
#define LF_LO(u, v) ((u) < (v) ? (u) : (v))
#define LF_HI(u, v) ((u) < (v) ? (v) : (u))
// (volatile: without it the compiler forwards the 78 stored values to their uses, i.e. keeps them in registers it
// does not have -- 880 bytes of local-memory spills per thread)
#define LF_PHI(u, v) (LF_LO(u, v) == 0 ? s.c[LF_HI(u, v)] : static_cast<const volatile double*>(phi)[(lf_idx(LF_LO(u, v), LF_HI(u, v)) - (kLfM + 1)) * kLfThreads])

template <int I>
__device__ __forceinline__ void lf_order(LfState& s, const double* __restrict__ phi) {
    if constexpr (I < kLfM) {
        const double eps = 2.2250738585072014e-308;             // util.tiny(float64)
        // main part: Q = sum_{v=1}^{I+1} a[I+1-v] * (sum_{j<=I} a[j] Phi[j][v])
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int v = 1; v <= I + 1; ++v) {
            double g0 = 0.0, g1 = 0.0;
#pragma unroll
            for (int j = 0; j <= I; j += 2) {
                g0 = fma(s.a[j], LF_PHI(j, v), g0);
                if (j + 1 <= I) g1 = fma(s.a[j + 1], LF_PHI(j + 1, v), g1);
            }
            if (v & 1) q0 = fma(s.a[I + 1 - v], g0 + g1, q0);
            else q1 = fma(s.a[I + 1 - v], g0 + g1, q1);
        }
        // head part: the pairs n = I+1 .. M-1, from the first samples directly
        double hsum = 0.0;
#pragma unroll
        for (int n = I + 1; n < kLfM; ++n) {
            double f = 0.0, b = 0.0;
#pragma unroll
            for (int j = 0; j <= I; ++j) {
                f = fma(s.a[j], s.head[n - j], f);
                b = fma(s.a[j], s.head[n - 1 - I + j], b);
            }
            hsum = fma(f, b, hsum);
        }
        const double num = (q0 + q1) + hsum;
        const double k = div_fast(num * -2.0, s.den + eps);
        // Levinson: a[j] <- a[j] + k a[I+1-j], j = 1 .. I+1 (a[I+1] = 0 before)
#pragma unroll
        for (int j = 1; 2 * j <= I + 1; ++j) {
            const double t1 = s.a[j], t2 = s.a[I + 1 - j];
            s.a[j] = fma(k, t2, t1);
            if (j != I + 1 - j) s.a[I + 1 - j] = fma(k, t1, t2);
        }
        s.a[I + 1] = k;                                         // j = I+1 pairs with a[0] = 1
        // edge errors of the new order: f_{I+1}[I+1] and b_{I+1}[N-1]
        double fe = 0.0, be = 0.0;
#pragma unroll
        for (int j = 0; j <= I + 1; ++j) {
            fe = fma(s.a[j], s.head[I + 1 - j], fe);
            be = fma(s.a[j], s.tl[I + 1 - j], be);
        }
        s.den = (1.0 - k * k) * s.den - be * be - fe * fe;
        if (!(s.den > kLfGainFloor * s.den0)) s.redo = true;    // also catches NaN
        lf_order<I + 1>(s, phi);
    }
}

// One CTA per segment, 96 threads = frames 0 .. 95 (three full warps; the two remaining frames of a 1 s segment go to
// k_lpc_redo).  The segment is staged by bulk-TMA copies (cp.async.bulk + mbarrier, SASS UBLKCP): one copy per row of
// 160 samples (= the frame shift) at a row pitch of 164 floats, so that the float4 a thread loads for its frame --
// frame f starts at row f -- falls into a different 16-byte bank group for each of eight consecutive lanes: the
// strided per-frame reads (stride 640 B: 32 distinct L1 lines per request from global memory, which bound the first
// version of this kernel by its load pipe) become conflict-free LDS.128.
constexpr int kLfRow = kLpcShift, kLfPitch = kLfRow + 4, kLfRows = 16000 / kLfRow;
struct LfSmem {
    float y[kLfRows * kLfPitch];                                 // later: Phi rows 1 .. 12, [78][96] doubles
    uint64_t bar;
};
static_assert((kLfPhi - kLfM - 1) * kLfThreads * sizeof(double) <= kLfRows * kLfPitch * sizeof(float), "Phi fits the staging buffer");

__global__ void __launch_bounds__(kLfThreads, 3) k_lpc_fast(const float* __restrict__ y, Geometry g, Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LfSmem& S = *reinterpret_cast<LfSmem*>(smem_raw);
    const int F = g.lpc_frames, tid = threadIdx.x, b = blockIdx.x;
    const float* yb = y + (size_t)b * g.L;
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(&S.bar, (uint32_t)(kLfRows * kLfRow * sizeof(float)));
        __syncwarp();
        for (int r = tid; r < kLfRows; r += 32)
            tma_bulk_g2s(S.y + r * kLfPitch, yb + r * kLfRow, (uint32_t)(kLfRow * sizeof(float)), &S.bar);
    }
    mbar_wait(&S.bar, 0);
    const int fr = tid;                                          // F >= 96 in 1 s mode (98)
    const float* srow = S.y + fr * kLfPitch;                     // sample n of the frame: srow[n + 4 (n / 160)]
    LfState s;
    double ring[kLfM + 1], acc[kLfM + 1];
    // y_emph = append(y[0], y[1:] - 0.97 * y[:-1]) (float32), frame = y_emph * hamming (float64)
    float yprev = fr > 0 ? srow[-5] : 0.f;                       // sample 160 fr - 1 = column 159 of the previous row
    double e_head = 0.0;
#pragma unroll
    for (int n4 = 0; n4 < 12; n4 += 4) {                         // samples 0 .. 11: fill the ring
        const float4 q = *reinterpret_cast<const float4*>(srow + n4);
        const float ys[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int n = n4 + w;
            const float e = (fr == 0 && n == 0) ? ys[w] : __fsub_rn(ys[w], __fmul_rn(0.97f, yprev));
            yprev = ys[w];
            const double x = (double)e * c_hamming400[n];
            ring[n] = x;
            s.head[n] = x;
            e_head = fma(x, x, e_head);
        }
    }
#pragma unroll
    for (int v = 0; v <= kLfM; ++v) acc[v] = 0.0;
    // samples 12 .. 399: the 13 lag sums over a register ring (slot = n % 13: static inside a block of 52 = lcm(13, 4))
    auto block = [&](int n0, auto count_c, bool first) {
        constexpr int kCount = decltype(count_c)::value;         // samples in this block, a multiple of 4
#pragma unroll
        for (int u4 = 0; u4 < kCount; u4 += 4) {
            const int n = n0 + u4;
            BPC_ASSERT(fr * kLfPitch + n + 4 * ((n >= kLfRow) + (n >= 2 * kLfRow)) + 3 < kLfRows * kLfPitch);
            const float4 q = *reinterpret_cast<const float4*>(srow + n + 4 * ((n >= kLfRow) + (n >= 2 * kLfRow)));
            const float ys[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                constexpr int kDummy = 0; (void)kDummy;
                const int u = u4 + w;
                const float e = __fsub_rn(ys[w], __fmul_rn(0.97f, yprev));
                yprev = ys[w];
                const double x = (double)e * c_hamming400[n + w];
                const int slot = (12 + u) % 13;
                ring[slot] = x;
                if (u == 0 && first) s.head[kLfM] = x;           // x[12]
#pragma unroll
                for (int v = 0; v <= kLfM; ++v) acc[v] = fma(x, ring[(slot - v + 13) % 13], acc[v]);
            }
        }
    };
#pragma unroll 1
    for (int blk = 0; blk < 7; ++blk) block(12 + 52 * blk, std::integral_constant<int, 52>{}, blk == 0);
    block(376, std::integral_constant<int, 24>{}, false);
    // x[N-1-k] sits in ring slot (399 - k) % 13 = (9 - k + 13) % 13
#pragma unroll
    for (int k = 0; k <= kLfM; ++k) s.tl[k] = ring[(9 - k + 13) % 13];
    // Phi: first row = the lag sums (registers), rows 1 .. 12 by the edge recurrence into the staging buffer, which
    // every thread of the CTA has finished reading
#pragma unroll
    for (int v = 0; v <= kLfM; ++v) s.c[v] = acc[v];
    __syncthreads();
    double* phi = reinterpret_cast<double*>(S.y) + tid;
#pragma unroll
    for (int u = 0; u < kLfM; ++u)
#pragma unroll
        for (int v = u; v < kLfM; ++v) {
            const double prev = u == 0 ? s.c[v] : phi[(lf_idx(u, v) - (kLfM + 1)) * kLfThreads];
            phi[(lf_idx(u + 1, v + 1) - (kLfM + 1)) * kLfThreads] =
                fma(-s.tl[u], s.tl[v], fma(s.head[kLfM - 1 - u], s.head[kLfM - 1 - v], prev));
        }
    // den_0 = sum fwd^2 + sum bwd^2 = 2 sum x^2 - x[0]^2 - x[N-1]^2
    const double c0 = acc[0] + e_head;
    s.den0 = s.den = 2.0 * c0 - s.head[0] * s.head[0] - s.tl[0] * s.tl[0];
    s.redo = false;
#pragma unroll
    for (int j = 0; j <= kLfM; ++j) s.a[j] = j == 0 ? 1.0 : 0.0;
    // (an all-zero frame keeps a = [1, 0, ...]: every reflection coefficient is -2 * 0 / (0 + tiny) = 0)
    if (s.den0 != 0.0) lf_order<0>(s, phi);
    float* coef = ws.lpc_coef + (size_t)b * kLpcOrder * F;
#pragma unroll
    for (int j = 1; j <= kLfM; ++j) coef[(j - 1) * F + fr] = (float)s.a[j];
    BPC_ASSERT(F - kLfThreads <= kLfThreads && b < ws.cap);
    if (s.redo) ws.lpc_redo[1 + atomicAdd(ws.lpc_redo, 1)] = b * F + fr;
    // frames 96 .. F-1: the direct method (no fourth warp for two frames)
    if (tid < F - kLfThreads) ws.lpc_redo[1 + atomicAdd(ws.lpc_redo, 1)] = b * F + kLfThreads + tid;
}

// The frames k_lpc_fast queued, by the direct recursion: one warp per frame.
__global__ void __launch_bounds__(kLpcThreads, 3) k_lpc_redo(const float* __restrict__ y, Geometry g, Tables tb,
                                                             Workspace ws) {
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * (kLpcThreads / 32), w0 = blockIdx.x * (kLpcThreads / 32) + (threadIdx.x >> 5);
    const int count = ws.lpc_redo[0], F = g.lpc_frames;
    for (int i = w0; i < count; i += warps) {
        const int gid = ws.lpc_redo[1 + i];
        const int b = gid / F, fr = gid - b * F;
        BPC_ASSERT(count <= ws.cap * F && gid >= 0 && b < ws.cap);
        const double a_lane = burg_frame(y + (size_t)b * g.L, fr * kLpcShift, tb.hamming400, lane);
        if (lane >= 1 && lane <= kLpcOrder) ws.lpc_coef[((size_t)b * kLpcOrder + (lane - 1)) * F + fr] = (float)a_lane;
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
    // (1 s mode, phase 2: the coefficients k_lpc_fast / k_lpc_redo left in Workspace::lpc_coef)
    float* coef = LONG ? ws.scratch + (size_t)b * ws.scratch_stride
                       : (phase == 2 ? ws.lpc_coef + (size_t)b * kLpcOrder * F_ : S.coef);

    // The trip count depends on blockIdx only and every warp runs every iteration (warps past the last frame redo it
    // and drop the result): the compiler can then prove that the ~200 shuffles per frame are convergent.  With
    // `fr = warp; fr < F; fr += 4` each of them was a WARPSYNC.COLLECTIVE / ENDCOLLECTIVE pair.
    const int frb = LONG ? blockIdx.y * (kLpcThreads / 32) : 0;
    const int frs = LONG ? gridDim.y * (kLpcThreads / 32) : kLpcThreads / 32;
    for (int fr_base = frb; fr_base < F_ && phase != 2; fr_base += frs) {
        const bool fr_valid = fr_base + warp < F_;
        const int fr = fr_valid ? fr_base + warp : F_ - 1;
        const double a_lane = burg_frame(yb, fr * kLpcShift, tb.hamming400, lane);
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
        cudaFuncSetAttribute(k_lpc_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LfSmem));
    });
    if (g.long_mode) {
        k_lpc<true><<<dim3(n, 16), kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 1);
        k_lpc<true><<<dim3(n, 1), kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 2);
        note_launch();
    } else {
        // BPC_LPC_FAST=0: the direct recursion for every frame in one kernel (the r02-g form)
        static const bool fast = !(std::getenv("BPC_LPC_FAST") && std::atoi(std::getenv("BPC_LPC_FAST")) == 0);
        if (fast && ws.lpc_coef && g.L == kLfRows * kLfRow && g.lpc_frames >= kLfThreads &&
            (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
            cudaMemsetAsync(ws.lpc_redo, 0, sizeof(int), st);
            k_lpc_fast<<<n, kLfThreads, sizeof(LfSmem), st>>>(y, g, ws);
            k_lpc_redo<<<148 * 3, kLpcThreads, 0, st>>>(y, g, tb, ws);
            k_lpc<false><<<n, kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 2);
            note_launch(2);
        } else {
            k_lpc<false><<<n, kLpcThreads, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats, 0);
        }
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
