// Shared device helpers: reductions, TMA-bulk staging, z-score finalisation.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>
#include <mutex>

namespace bpc {

// Function attributes (the opt-in dynamic shared-memory size) are per DEVICE: a launcher runs its attribute block once
// per device it is used on, under a lock (handles on different devices may be driven from different host threads).
struct PerDeviceOnce {
    std::mutex m;
    unsigned long long mask = 0;
    template <class F> void run(F&& f) {
        int d = 0;
        cudaGetDevice(&d);
        std::lock_guard<std::mutex> l(m);
        if ((mask >> (d & 63)) & 1ull) return;
        f();
        mask |= 1ull << (d & 63);
    }
};


// Index assertions of the checked build (`make checked` -> gpurun_variants/lib_checked.so, -DBPC_CHECKED): the GPU pool
// has no compute-sanitizer, so the kernels added in r02 state their own bounds; a violated one traps the kernel and the
// next API call reports the launch failure.  Compiled out of the product library.
#ifdef BPC_CHECKED
#include <cassert>
#define BPC_ASSERT(c) assert(c)
#else
#define BPC_ASSERT(c) ((void)0)
#endif

constexpr int kPlaneRows = 128;
constexpr int kMagStride = 260;          // |STFT512| workspace row stride (257 valid bins, 16-byte aligned rows)
constexpr int kMag2048Stride = 1028;     // |STFT2048| workspace row stride (1025 valid bins)
constexpr int kMaxFrames = 64;           // on-chip kernels of this build hold T <= 64 frames (1 s @ 16 kHz, hop 256)
constexpr int kMaxLen = 16384;           // ... and L <= 16384 samples

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide reductions over blockDim.x threads (multiple of 32, <= 1024).  `scratch` holds >= 32 elements of T and is
// reused; every thread gets the result.  Contains __syncthreads(): call from all threads.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T ident, Op op, T* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    T r = (lane < nw) ? scratch[lane] : ident;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}
struct OpAdd { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpMaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMaxD { __device__ double operator()(double a, double b) const { return fmax(a, b); } };
struct OpMinD { __device__ double operator()(double a, double b) const { return fmin(a, b); } };

__device__ __forceinline__ double block_sum(double v, double* scratch) { return block_reduce(v, 0.0, OpAdd(), scratch); }
// two sums behind one pair of barriers (blockDim.x <= 512: 2 x 16 partials in the 32-element scratch array)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { scratch[w] = a; scratch[16 + w] = b; }
    __syncthreads();
    double ra = (lane < nw) ? scratch[lane] : 0.0, rb = (lane < nw) ? scratch[16 + lane] : 0.0;
    a = warp_sum(ra);
    b = warp_sum(rb);
}
__device__ __forceinline__ float block_max(float v, float* scratch) { return block_reduce(v, -FLT_MAX, OpMaxF(), scratch); }
__device__ __forceinline__ float block_min(float v, float* scratch) { return block_reduce(v, FLT_MAX, OpMinF(), scratch); }

// ------------------------------------------------------------------------------------ numpy-style z-score terms
// (x - mean) / (std + 1e-8) with mean / std already rounded to float32 (numpy float32 reductions), float32 ops.
struct ZTerm {
    float mean, denom, rden;
    // (x - mean) / denom: quotient from the reciprocal plus one residual correction (correctly rounded except in rare
    // double-rounding cases; a plain IEEE division costs ~10 instructions per output element with its slow path)
    __device__ __forceinline__ float operator()(float x) const {
        const float d = __fsub_rn(x, mean);
        const float q = __fmul_rn(d, rden);
        return __fmaf_rn(__fmaf_rn(-q, denom, d), rden, q);
    }
};
__device__ __forceinline__ ZTerm make_zterm(double sum, double sumsq, double n) {
    const double mean = sum / n;
    double var = sumsq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    ZTerm z;
    z.mean = (float)mean;
    z.denom = __fadd_rn((float)sqrt(var), 1e-8f);
    z.rden = __frcp_rn(z.denom);
    return z;
}

// numpy's float32 add-reduce of a contiguous run of n <= 128 elements (pairwise_sum with its 8 strided accumulators),
// reproduced operation for operation by one warp; every lane returns the sum.  Row-wise z-scores (process.py:47,55) take
// their mean from this, which matters when a row is constant: numpy's rounded mean then differs from the value and the
// "normalised" row becomes -1 / +1 instead of 0 (silent input -> MFCC row 0).
__device__ __forceinline__ float np_sum_f32_warp128(const float* a, int n, int lane) {
    float res;
    if (n < 8) {
        res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    const int body = n - (n & 7);
    float r = 0.f;
    if (lane < 8) {
        r = a[lane];
        for (int i = 8 + lane; i < body; i += 8) r = __fadd_rn(r, a[i]);
    }
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    res = __shfl_sync(0xffffffffu, r, 0);
    for (int i = body; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

// Any n: numpy's pairwise_sum recursion above 128 elements (n2 = n / 2 rounded down to a multiple of 8, left + right),
// walked with an explicit stack (uniform across the warp).  Rows longer than 128 only occur in long mode.
__device__ inline float np_sum_f32_warp(const float* a, int n, int lane) {
    if (n <= 128) return np_sum_f32_warp128(a, n, lane);
    float vals[12];
    int off[12], len[12], state[12];
    int sp = 0, tp = 0;
    off[0] = 0; len[0] = n; state[0] = 0; tp = 1;
    while (tp > 0) {
        const int f = tp - 1;
        if (len[f] <= 128) {
            vals[sp++] = np_sum_f32_warp128(a + off[f], len[f], lane);
            --tp;
        } else {
            int n2 = len[f] / 2;
            n2 -= n2 % 8;
            if (state[f] == 0) {
                state[f] = 1;
                off[tp] = off[f]; len[tp] = n2; state[tp] = 0; ++tp;
            } else if (state[f] == 1) {
                state[f] = 2;
                off[tp] = off[f] + n2; len[tp] = len[f] - n2; state[tp] = 0; ++tp;
            } else {
                const float r = vals[--sp], l = vals[--sp];
                vals[sp++] = __fadd_rn(l, r);
                --tp;
            }
        }
    }
    return vals[0];
}

// Row-wise z-score terms the numpy way: float32 mean from np_sum_f32_warp, deviations in float32, population std.
// `a` holds n floats (shared memory, or global scratch in long mode); one warp per row.
__device__ __forceinline__ ZTerm np_row_zterm(const float* a, int n, int lane) {
    const float mean = __fdiv_rn(np_sum_f32_warp(a, n, lane), (float)n);
    double q = 0.0;
    for (int t = lane; t < n; t += 32) {
        const float d = __fsub_rn(a[t], mean);
        q += (double)__fmul_rn(d, d);
    }
    q = warp_sum(q);
    ZTerm z;
    z.mean = mean;
    z.denom = __fadd_rn((float)sqrt(q / (double)n), 1e-8f);
    z.rden = __frcp_rn(z.denom);
    return z;
}

// ----------------------------------------------------------- 1-D TMA bulk copy global -> shared (SASS: UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Stage `n` floats of a segment (global, 16-byte aligned base, n % 4 == 0) into shared memory with one TMA bulk
// copy per <=32 KiB slice; all threads return after the data is visible.  `bar` is a shared uint64_t.
__device__ __forceinline__ void stage_segment_tma(float* dst, const float* src, int n, uint64_t* bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = (uint32_t)n * 4u;
        mbar_expect_tx(bar, total);
        for (uint32_t off = 0; off < total; off += 32768u) {
            const uint32_t len = (total - off) < 32768u ? (total - off) : 32768u;
            tma_bulk_g2s((char*)dst + off, (const char*)src + off, len, bar);
        }
    }
    mbar_wait(bar, 0);
}

// Thread-strided copy loop over n float2 elements of a segment with BATCH loads in flight per thread before the first
// use: sink(m, v) receives element m (elements >= n_valid read as zero).  The plain `for (m = tid; ...) dst[f(m)] = src[m]`
// loops of the CTA-per-segment kernels compile to one LDG -> one STS per trip, i.e. one DRAM round trip per element and
// thread (k_hilbert: 20 serialised round trips, 20 % of its stall samples; k_cens_dec: 19 %, profiles/r02_j_*).
template <int BATCH, class Sink>
__device__ __forceinline__ void load_f2_batched(const float2* __restrict__ src, int n, int n_valid, int tid, int nthreads,
                                                Sink&& sink) {
    for (int m0 = tid; m0 < n; m0 += BATCH * nthreads) {
        float2 v[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int m = m0 + u * nthreads;
            v[u] = m < n_valid ? __ldg(src + m) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int m = m0 + u * nthreads;
            if (m < n) sink(m, v[u]);
        }
    }
}

// float atomic min/max via ordered-int trick (values must not be NaN)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v));
    else atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

// ------------------------------------------------------------------------------ dataset statistics, fused (r01 v37)
// Running statistics of the values a thread stores into one output plane; flushed per CTA into the handle's
// [(9 + S), 5] accumulator {count, sum, sum of squares, min, max} (bpc_channel_stats; BASELINE config 3's all-reduce
// payload).  Up to v36 a separate kernel re-read the 0.8 GB of planes of every 4096-segment step for this.
struct StatAcc {
    double s, q;
    float mn, mx;
    int cnt;
    __device__ __forceinline__ void init() { s = 0.0; q = 0.0; mn = FLT_MAX; mx = -FLT_MAX; cnt = 0; }
    __device__ __forceinline__ void add(float v) {
        if ((__float_as_uint(v) & 0x7f800000u) != 0x7f800000u) {
            const double d = (double)v;
            s += d;
            q = fma(d, d, q);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
            ++cnt;
        }
    }
    __device__ __forceinline__ void add_n(float v, int m) {          // the same value m times (pad_freq rows)
        if ((__float_as_uint(v) & 0x7f800000u) != 0x7f800000u && m > 0) {
            const double d = (double)v;
            s += (double)m * d;
            q += (double)m * (d * d);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
            cnt += m;
        }
    }
};
// Call from all threads of a CTA of at most 10 warps; dscratch / fscratch: the CTA's 32-element reduction scratch arrays.
__device__ __forceinline__ void stat_flush_block(const StatAcc& a, double* acc_c, double* dscratch, float* fscratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double s = warp_sum(a.s), q = warp_sum(a.q);
    const int cnt = warp_sum(a.cnt);
    const float mn = warp_min(a.mn), mx = warp_max(a.mx);
    __syncthreads();
    if (lane == 0) {
        dscratch[w] = s; dscratch[10 + w] = q; dscratch[20 + w] = (double)cnt;
        fscratch[w] = mn; fscratch[16 + w] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, Q = 0.0, C = 0.0;
        float MN = FLT_MAX, MX = -FLT_MAX;
        for (int i = 0; i < nw; ++i) {
            S += dscratch[i]; Q += dscratch[10 + i]; C += dscratch[20 + i];
            MN = fminf(MN, fscratch[i]); MX = fmaxf(MX, fscratch[16 + i]);
        }
        if (C > 0.0) {
            atomicAdd(acc_c + 0, C);
            atomicAdd(acc_c + 1, S);
            atomicAdd(acc_c + 2, Q);
            atomic_min_double(acc_c + 3, (double)MN);
            atomic_max_double(acc_c + 4, (double)MX);
        }
    }
    __syncthreads();
}

}  // namespace bpc
