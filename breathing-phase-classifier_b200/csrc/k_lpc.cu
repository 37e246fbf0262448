// LPC channel (process.py:64-67, methods.py:116-134): pre-emphasis 0.97, 25 ms / 10 ms Hamming frames, Burg's method
// of order 12 as in librosa.lpc (float64), whole-array z-score over ALL frames, first T frames kept, rows padded.
// One CTA per segment, one warp per frame; forward / backward prediction errors live in shared memory.
#include <cmath>
#include "kernels.cuh"

namespace bpc {

constexpr int kLpcFrame = 400, kLpcShift = 160, kLpcOrder = 12, kLpcMaxFrames = 112;

struct LpcSmem {
    double fa[8][kLpcFrame];          // forward errors, indexed by sample
    double ba[8][kLpcFrame];          // backward errors
    float coef[kLpcOrder * kLpcMaxFrames];
    double dscratch[32];
    float fscratch[32];
};

__global__ void __launch_bounds__(256) k_lpc(const float* __restrict__ y, Geometry g, Tables tb, Workspace ws,
                                             float* feats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LpcSmem& S = *reinterpret_cast<LpcSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L, T = g.T, F = g.lpc_frames;
    const float* yb = y + (size_t)b * L;
    double* FA = S.fa[warp];
    double* BA = S.ba[warp];
    const double eps = 2.2250738585072014e-308;             // util.tiny(float64)

    for (int fr = warp; fr < F; fr += 8) {
        const int start = fr * kLpcShift;
        for (int n = lane; n < kLpcFrame; n += 32) {
            const int gi = start + n;
            // y_emph = append(y[0], y[1:] - 0.97 * y[:-1])   (float32)
            const float e = gi == 0 ? __ldg(yb) : __fsub_rn(__ldg(yb + gi), __fmul_rn(0.97f, __ldg(yb + gi - 1)));
            const double x = (double)e * tb.hamming400[n];
            FA[n] = x;
            BA[n] = x;
        }
        __syncwarp();
        // fwd[j] = FA[j + 1 + i], bwd[j] = BA[j], j in [0, 399 - i) at iteration i
        double den = 0.0;
        for (int j = lane; j < kLpcFrame - 1; j += 32) den += FA[j + 1] * FA[j + 1] + BA[j] * BA[j];
        den = warp_sum(den);
        double a_cur[kLpcOrder + 1], a_prev[kLpcOrder + 1];
#pragma unroll
        for (int j = 0; j <= kLpcOrder; ++j) { a_cur[j] = j == 0 ? 1.0 : 0.0; a_prev[j] = a_cur[j]; }
#pragma unroll
        for (int i = 0; i < kLpcOrder; ++i) {
            const int len = kLpcFrame - 1 - i;
            double num = 0.0;
            for (int j = lane; j < len; j += 32) num += BA[j] * FA[j + 1 + i];
            num = warp_sum(num);
            const double k = (num * -2.0) / (den + eps);
            // ar_coeffs_prev, ar_coeffs = ar_coeffs, ar_coeffs_prev ; then the Levinson update
#pragma unroll
            for (int j = 0; j <= kLpcOrder; ++j) { const double tmp = a_prev[j]; a_prev[j] = a_cur[j]; a_cur[j] = tmp; }
#pragma unroll
            for (int j = 1; j <= i + 1; ++j) a_cur[j] = a_prev[j] + k * a_prev[i - j + 1];
            for (int j = lane; j < len; j += 32) {
                const double f = FA[j + 1 + i], bw = BA[j];
                FA[j + 1 + i] = f + k * bw;
                BA[j] = bw + k * f;
            }
            __syncwarp();
            const double q = 1.0 - k * k;
            const double bl = BA[len - 1], f0 = FA[1 + i];
            den = q * den - bl * bl - f0 * f0;
            __syncwarp();
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < kLpcOrder; ++c) S.coef[c * F + fr] = (float)a_cur[c + 1];
        }
    }
    __syncthreads();
    if (ws.dbg_lpc) {
        float* d = ws.dbg_lpc + (size_t)b * kLpcOrder * F;
        for (int i = tid; i < kLpcOrder * F; i += 256) d[i] = S.coef[i];
    }
    // whole-array z over all F frames (process.py:65); pad_time keeps the first T columns; pad value = min of those
    double s = 0.0, q = 0.0;
    for (int i = tid; i < kLpcOrder * F; i += 256) { const double v = (double)S.coef[i]; s += v; q += v * v; }
    s = block_sum(s, S.dscratch);
    q = block_sum(q, S.dscratch);
    const ZTerm z = make_zterm(s, q, (double)(kLpcOrder * F));
    const int Tk = T < F ? T : F;
    float mn = FLT_MAX;
    for (int i = tid; i < kLpcOrder * Tk; i += 256) {
        const int c = i / Tk, t = i - c * Tk;
        mn = fminf(mn, z(S.coef[c * F + t]));
    }
    mn = block_min(mn, S.fscratch);
    float* o = plane_ptr(feats, b, BPC_CH_LPC, T);
    for (int i = tid; i < kPlaneRows * T; i += 256) {
        const int c = i / T, t = i - c * T;
        o[i] = (c < kLpcOrder && t < Tk) ? z(S.coef[c * F + t]) : mn;
    }
}

void launch_lpc(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                cudaStream_t st) {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(k_lpc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpcSmem));
        done = true;
    }
    k_lpc<<<n, 256, sizeof(LpcSmem), st>>>(y, g, tb, ws, feats);
    note_launch();
}

// ------------------------------------------------------------------------------- dataset-level statistics + padding
// acc layout: [(9 + nscal)][5] doubles = {count, sum, sumsq, min, max}
__global__ void __launch_bounds__(256) k_stats(Geometry g, const float* __restrict__ feats,
                                               const float* __restrict__ scalars, double* __restrict__ acc) {
    __shared__ double dscratch[32];
    const int b = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    if (c < 9) {
        const int NP = kPlaneRows * g.T;
        const float* p = feats + ((size_t)b * 9 + c) * NP;
        double s = 0.0, q = 0.0, mn = 1e300, mx = -1e300, cnt = 0.0;
        for (int i = tid; i < NP; i += 256) {
            const double v = (double)p[i];
            if (isfinite(v)) { s += v; q += v * v; mn = fmin(mn, v); mx = fmax(mx, v); cnt += 1.0; }
        }
        s = block_sum(s, dscratch);
        q = block_sum(q, dscratch);
        cnt = block_sum(cnt, dscratch);
        mn = block_reduce(mn, 1e300, OpMinD(), dscratch);
        mx = block_reduce(mx, -1e300, OpMaxD(), dscratch);
        if (tid == 0 && cnt > 0.0) {
            double* a = acc + (size_t)c * 5;
            atomicAdd(a + 0, cnt);
            atomicAdd(a + 1, s);
            atomicAdd(a + 2, q);
            atomic_min_double(a + 3, mn);
            atomic_max_double(a + 4, mx);
        }
    } else {
        for (int i = tid; i < g.nscal; i += 256) {
            const double v = (double)scalars[(size_t)b * g.nscal + i];
            if (isfinite(v)) {
                double* a = acc + (size_t)(9 + i) * 5;
                atomicAdd(a + 0, 1.0);
                atomicAdd(a + 1, v);
                atomicAdd(a + 2, v * v);
                atomic_min_double(a + 3, v);
                atomic_max_double(a + 4, v);
            }
        }
    }
}

void launch_stats(int n, const Geometry& g, const float* feats, const float* scalars, double* acc, cudaStream_t st) {
    dim3 grid(n, 10);
    k_stats<<<grid, 256, 0, st>>>(g, feats, scalars, acc);
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
