// The one GEMM-shaped stage of the path on the 5th-generation tensor cores: the time-axis DCT of `mod_spec`
// (methods.py:142-143, second `dct`) in long mode, C2[40 n_seg, T] = C1[40 n_seg, T] . D^T with D the [T, T] ortho DCT-II
// matrix (T = 1876 at 30 s: 140 MFLOP per segment, one [5440 x 1876] x [1876 x 1876] product per 136-segment launch).
//
// tcgen05.mma kind::tf32, cta_group::1, M = 128 x N = 128 tiles, accumulator in TMEM (128 lanes x 128 columns), both
// operands K-major in shared memory in the canonical no-swizzle layout (8-row x 16-byte core matrices), written by the
// CTA's own threads (the A rows are scattered over the per-segment scratch regions, so there is no TMA tensor map).
// Precision: 3xTF32 -- every float32 operand is split into hi = tf32(x) and lo = tf32(x - hi) and the product is
// accumulated as lo*hi + hi*lo + hi*hi in the FP32 accumulator, which recovers float32-level products (a single TF32
// pass has 10 mantissa bits: 1e-3 relative, two orders beyond the parity budget of this stage).
// The TMEM accumulator adds with truncation, a bias that grows with the number of K steps when the products share a
// sign -- and row 0 of C1 (the mel-axis DC term, ~ -600 dB-units in every frame) times the time-axis DC basis is exactly
// that: r02 measured 16 float32 ulps at T = 188 and 48 at T = 313.  So the row means of C1 are removed before the product
// (k_tc_row_means; the residual rows have mixed signs) and put back exactly in the epilogue as mean * sum_t D[u][t]
// (a [T] table; zero up to rounding for every u > 0).
#include <cstdlib>
#include "kernels.cuh"

namespace bpc {

size_t consumer_scratch_floats(int T);
constexpr int kTcM = 128, kTcN = 128, kTcK = 32;                 // CTA tile; K per stage (4 MMAs of K = 8)
constexpr int kTcOpFloats = kTcM * kTcK;                         // one operand stage: 128 rows x 32 k = 16 KB

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// canonical K-major, no swizzle: [k / 4][row / 8][row % 8][k % 4] floats -> 128-byte core matrices; consecutive 8-row
// groups are 128 B apart (SBO), the two 16-byte K chunks of one K = 8 MMA are rows/8 * 128 B apart (LBO)
__device__ __forceinline__ int tc_idx(int row, int k) { return (((k >> 2) * (kTcM / 8) + (row >> 3)) << 5) + ((row & 7) << 2) + (k & 3); }

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr_bytes) {
    constexpr uint32_t kLbo = (kTcM / 8) * 128, kSbo = 128;
    uint64_t d = (uint64_t)((saddr_bytes >> 4) & 0x3FFFu);
    d |= (uint64_t)((kLbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((kSbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                                      // descriptor version 1 (sm_100); layout type 0 = no swizzle
    return d;
}

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kTcIdesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128) k_modspec_time_tc(Geometry g, Tables tb, Workspace ws, int n_seg, size_t role0_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* Ahi = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* Alo = Ahi + kTcOpFloats;
    uint32_t* Bhi = Alo + kTcOpFloats;
    uint32_t* Blo = Bhi + kTcOpFloats;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = g.T, m0 = blockIdx.y * kTcM, n0 = blockIdx.x * kTcN, m_total = n_seg * 40;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTcN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    // thread `tid` owns row tid of both operand tiles: A row m (segment m / 40, mel-DCT row m % 40 of C1), B row u of D
    const int m = m0 + tid, u = n0 + tid;
    const float* a_row = nullptr;
    if (m < m_total) {
        const int seg = m / 40, r = m - seg * 40;
        a_row = ws.scratch + (size_t)seg * ws.scratch_stride + role0_off + (size_t)(kPlaneRows + r) * T;
    }
    const float* b_row = u < T ? tb.dct_time_n + (size_t)u * T : nullptr;

    uint32_t phase = 0;
    for (int k0 = 0; k0 < T; k0 += kTcK) {
#pragma unroll 8
        for (int kk = 0; kk < kTcK; ++kk) {
            const int k = k0 + kk, idx = tc_idx(tid, kk);
            const float a = (a_row && k < T) ? a_row[k] : 0.f;
            const float b = (b_row && k < T) ? __ldg(b_row + k) : 0.f;
            const uint32_t ah = tf32_rna(a), bh = tf32_rna(b);
            Ahi[idx] = ah;
            Alo[idx] = tf32_rna(a - __uint_as_float(ah));
            Bhi[idx] = bh;
            Blo[idx] = tf32_rna(b - __uint_as_float(bh));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA unit
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa_hi = smem_u32(Ahi), sa_lo = smem_u32(Alo), sb_hi = smem_u32(Bhi), sb_lo = smem_u32(Blo);
            constexpr uint32_t kStep = 2 * (kTcM / 8) * 128;              // two K chunks = one K = 8 MMA
#pragma unroll
            for (int term = 0; term < 3; ++term) {                        // small terms first: lo*hi, hi*lo, hi*hi
                const uint32_t sa = term == 0 ? sa_lo : sa_hi, sb = term == 1 ? sb_lo : sb_hi;
#pragma unroll
                for (int j = 0; j < kTcK / 8; ++j)
                    tc_mma(tmem, tc_desc(sa + j * kStep), tc_desc(sb + j * kStep), (k0 | term | j) != 0);
            }
            // arrives on the barrier when every MMA issued so far has completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(&bar, phase);                                           // the operand buffers may be refilled
        phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (= tile rows), 8 columns at a time
    const int mrow = m0 + warp * 32 + lane;
    float* c_row = nullptr;
    if (mrow < m_total) {
        const int seg = mrow / 40, r = mrow - seg * 40;
        c_row = ws.scratch + (size_t)seg * ws.scratch_stride + role0_off + (size_t)(kPlaneRows + 40 + r) * T;
    }
#pragma unroll 1
    for (int c = 0; c < kTcN; c += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c_row) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (n0 + c + q < T) c_row[n0 + c + q] = __uint_as_float(v[q]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcN));
}

// ------------------------------------------------------------------------------------------ pipelined version
// Both operands are first split (hi / lo) and written to global memory ALREADY in the canonical shared-memory layout,
// one contiguous 16 KB block per (row tile, K block, half); D's tiles are made once per handle, C1's once per launch.
// The GEMM kernel is then a pure copy-engine + tensor-core pipeline: a producer thread streams 4 x 16 KB per stage
// with bulk TMA copies (cp.async.bulk, mbarrier complete_tx), an MMA thread issues the 12 tcgen05.mma of the stage and
// releases it with tcgen05.commit, and the four warps read the accumulator out of TMEM at the end.
constexpr int kTcStages = 3;
constexpr int kTcStageFloats = 4 * kTcOpFloats;                  // A hi, A lo, B hi, B lo

// one warp per row of A: float32 mean (FP64 sum) of the T values of a C1 row
__global__ void __launch_bounds__(128) k_tc_row_means(Geometry g, Workspace ws, int rows_total, size_t role0_off) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31, T = g.T;
    if (row >= rows_total) return;
    const int seg = row / 40, r = row - seg * 40;
    const float* src = ws.scratch + (size_t)seg * ws.scratch_stride + role0_off + (size_t)(kPlaneRows + r) * T;
    double acc = 0.0;
    for (int t = lane; t < T; t += 32) acc += (double)src[t];
    acc = warp_sum(acc);
    if (lane == 0) ws.tc_mean[row] = (float)(acc / (double)T);
}

// rows of the operand: A (is_a): row m = segment m / 40, mel-DCT row m % 40 of C1 in the scratch regions; B: row u of D
__global__ void __launch_bounds__(128) k_tc_prep_tiles(Geometry g, Tables tb, Workspace ws, int is_a, int rows_total,
                                                       size_t role0_off, uint32_t* __restrict__ out) {
    const int T = g.T, kb = blockIdx.x, rt = blockIdx.y, tid = threadIdx.x;
    const int row = rt * kTcM + tid, KB = gridDim.x;
    const float* src = nullptr;
    if (row < rows_total) {
        if (is_a) {
            const int seg = row / 40, r = row - seg * 40;
            src = ws.scratch + (size_t)seg * ws.scratch_stride + role0_off + (size_t)(kPlaneRows + r) * T;
        } else {
            src = tb.dct_time_n + (size_t)row * T;
        }
    }
    uint32_t* hi = out + ((size_t)(rt * KB + kb) * 2) * kTcOpFloats;
    uint32_t* lo = hi + kTcOpFloats;
    const float mean = (is_a && src) ? ws.tc_mean[row] : 0.f;
#pragma unroll
    for (int c = 0; c < kTcK / 4; ++c) {                          // one 16-byte K chunk per store: coalesced across rows
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = kb * kTcK + 4 * c + q;
            const float x = (src && k < T) ? __fsub_rn(src[k], mean) : 0.f;
            h[q] = tf32_rna(x);
            l[q] = tf32_rna(x - __uint_as_float(h[q]));
        }
        const int idx = tc_idx(tid, 4 * c);
        *reinterpret_cast<uint4*>(hi + idx) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + idx) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void __launch_bounds__(128, 1) k_modspec_time_tc_pipe(Geometry g, Tables tb, Workspace ws, const uint32_t* __restrict__ a_tiles,
                                                                 const uint32_t* __restrict__ b_tiles, int n_seg, size_t role0_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* stage0 = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ __align__(8) uint64_t full[kTcStages], empty[kTcStages], done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = g.T, KB = (T + kTcK - 1) / kTcK, m_total = n_seg * 40;
    const int nt = blockIdx.x, mt = blockIdx.y;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTcN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {                                               // producer: bulk TMA copies, 64 KB per stage
        const uint32_t* a_src = a_tiles + (size_t)mt * KB * 2 * kTcOpFloats;
        const uint32_t* b_src = b_tiles + (size_t)nt * KB * 2 * kTcOpFloats;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % kTcStages;
            if (kb >= kTcStages) mbar_wait(&empty[s], (uint32_t)((kb / kTcStages - 1) & 1));
            uint32_t* dst = stage0 + (size_t)s * kTcStageFloats;
            mbar_expect_tx(&full[s], (uint32_t)(kTcStageFloats * sizeof(uint32_t)));
            tma_bulk_g2s(dst, a_src + (size_t)kb * 2 * kTcOpFloats, 2 * kTcOpFloats * sizeof(uint32_t), &full[s]);
            tma_bulk_g2s(dst + 2 * kTcOpFloats, b_src + (size_t)kb * 2 * kTcOpFloats, 2 * kTcOpFloats * sizeof(uint32_t), &full[s]);
        }
    } else if (tid == 32) {                                       // MMA issuer
        constexpr uint32_t kStep = 2 * (kTcM / 8) * 128;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % kTcStages;
            mbar_wait(&full[s], (uint32_t)((kb / kTcStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa_hi = smem_u32(stage0 + (size_t)s * kTcStageFloats), sa_lo = sa_hi + kTcOpFloats * 4;
            const uint32_t sb_hi = sa_lo + kTcOpFloats * 4, sb_lo = sb_hi + kTcOpFloats * 4;
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const uint32_t sa = term == 0 ? sa_lo : sa_hi, sb = term == 1 ? sb_lo : sb_hi;
#pragma unroll
                for (int j = 0; j < kTcK / 8; ++j)
                    tc_mma(tmem, tc_desc(sa + j * kStep), tc_desc(sb + j * kStep), (kb | term | j) != 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    __syncwarp();
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int mrow = mt * kTcM + warp * 32 + lane, n0 = nt * kTcN;
    float* c_row = nullptr;
    float mean = 0.f;
    if (mrow < m_total) {
        const int seg = mrow / 40, r = mrow - seg * 40;
        c_row = ws.scratch + (size_t)seg * ws.scratch_stride + role0_off + (size_t)(kPlaneRows + 40 + r) * T;
        mean = ws.tc_mean[mrow];
    }
#pragma unroll 1
    for (int c = 0; c < kTcN; c += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c_row) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (n0 + c + q < T)                                // the row mean comes back: mean * sum_t D[u][t]
                    c_row[n0 + c + q] = __fmaf_rn(mean, __ldg(tb.dct_colsum + n0 + c + q), __uint_as_float(v[q]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcN));
}

size_t tc_tile_words(int rows, int T) {                           // uint32 words of a pre-tiled hi / lo operand
    return (size_t)((rows + kTcM - 1) / kTcM) * ((T + kTcK - 1) / kTcK) * 2 * kTcOpFloats;
}

bool modspec_time_tc_enabled(const Tables& tb) {
    static const char* env = std::getenv("BPC_TC_DCT");
    return tb.dct_time_n != nullptr && !(env && std::atoi(env) == 0);
}

void launch_modspec_time_tc(int n, const Geometry& g, const Tables& tb, const Workspace& ws, size_t role0_off,
                            cudaStream_t st) {
    static PerDeviceOnce once;
    static int mode = 2;                                          // BPC_TC_DCT: 2 = pipelined (default), 1 = single stage
    const int bytes1 = 4 * kTcOpFloats * (int)sizeof(uint32_t), bytes2 = kTcStages * kTcStageFloats * (int)sizeof(uint32_t);
    once.run([&] {
        cudaFuncSetAttribute(k_modspec_time_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes1);
        cudaFuncSetAttribute(k_modspec_time_tc_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes2);
        const char* env = std::getenv("BPC_TC_DCT");
        if (env && std::atoi(env) == 1) mode = 1;
    });
    const int KB = (g.T + kTcK - 1) / kTcK, MT = (n * 40 + kTcM - 1) / kTcM, NT = (g.T + kTcN - 1) / kTcN;
    if (mode == 1 || !ws.tc_a || !tb.dct_tiles) {
        k_modspec_time_tc<<<dim3(NT, MT), 128, bytes1, st>>>(g, tb, ws, n, role0_off);
        note_launch();
        return;
    }
    k_tc_row_means<<<(n * 40 + 3) / 4, 128, 0, st>>>(g, ws, n * 40, role0_off);
    k_tc_prep_tiles<<<dim3(KB, MT), 128, 0, st>>>(g, tb, ws, 1, n * 40, role0_off, ws.tc_a);
    k_modspec_time_tc_pipe<<<dim3(NT, MT), 128, bytes2, st>>>(g, tb, ws, ws.tc_a, tb.dct_tiles, n, role0_off);
    note_launch(3);
}

// D's hi / lo tiles, once per handle (after dct_time_n is uploaded)
void launch_tc_prep_b(const Geometry& g, const Tables& tb, uint32_t* out, cudaStream_t st) {
    const int KB = (g.T + kTcK - 1) / kTcK, NT = (g.T + kTcN - 1) / kTcN;
    Workspace none{};
    k_tc_prep_tiles<<<dim3(KB, NT), 128, 0, st>>>(g, tb, none, 0, g.T, 0, out);
    note_launch();
}

}  // namespace bpc
