// chroma_cens rows of the chroma channel (process.py:53-57): librosa.feature.chroma_cens(y, sr=16000, hop_length=256)
//   = 7-octave / 36-bins-per-octave constant-Q transform (multi-rate recursion: rectangular-window STFT-512 of the
//     signal decimated by 2 per octave, times one sparse complex basis that is the same for every octave),
//     |.| / sqrt(filter length), fold 252 -> 12, L1 normalise, 4-level quantise, 41-tap Hann smoothing, L2 normalise.
// One CTA per segment (r01 v2 layout):
//   phase 1  six cascaded 2:1 half-band decimations, each thread producing 4 consecutive outputs from a register
//            window of FP64 samples (the signals live in shared memory de-interleaved into even / odd samples, which is
//            also the complex packing the real FFT wants);
//   phase 2  one half-warp per frame walks the 7 octaves: team_fft<16> (fft_reg.cuh) on 512 samples, real split of the
//            85 bins the sparsified bases touch, FP32 sparse complex product, fold to 12 chroma sums kept in registers;
//   phase 3  CENS post-processing on the [12, T] tile.
// The decimator stands in for soxr_hq (absent library): the same 127-tap Kaiser half-band the oracle uses.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "kernels.cuh"
#include "fft.cuh"
#include "fft_reg.cuh"
#include "tables.hpp"

namespace bpc {

constexpr int kBinLo = 60, kBinHi = 144;               // rfft bins the sparsified bases can touch (measured 62..141)
constexpr int kBinSpan = kBinHi - kBinLo + 1;
constexpr int kOutPerThread = 4;
constexpr int kWin = 63 + kOutPerThread;                // odd-sample window of one thread

// half-band taps in constant memory: centre tap and the 32 odd-offset taps g[m] = h[63 + 2m + 1] (the even offsets of
// a half-band filter are zero, the filter is symmetric)
__constant__ double c_hb_centre;
__constant__ double c_hb_odd[32];

void upload_cens_constants(const double* taps127) {
    double odd[32];
    for (int m = 0; m < 32; ++m) odd[m] = taps127[63 + 2 * m + 1];
    cudaMemcpyToSymbol(c_hb_centre, taps127 + 63, sizeof(double));
    cudaMemcpyToSymbol(c_hb_odd, odd, sizeof(odd));
}

// Position of sample q (q >= -kZPad) in a de-interleaved array: kZPad zeros on either side stand in for librosa's
// zero centre-padding (no bounds checks in the loaders), and one pad word per 32 keeps the stride-4 accesses of the
// decimator (thread i reads q = 4 i + c) on 32 distinct banks.
constexpr int kZPad = 128;
__device__ __forceinline__ int cens_dec_stride(const Workspace& ws) { return ws.dec_stride; }
#ifdef BPC_DEC_PAD32
// r01 layout: one pad word per 32 and scalar loads (thread i reads q = 4 i + c: 1.84 wavefronts per load on average over
// the 67 offsets of a window -- every one of k_cens_dec's 2.7e7 excess shared wavefronts, profiles/r02_j_*)
__device__ __forceinline__ int ppos(int q) { return (q + kZPad) + ((q + kZPad) >> 5); }
__host__ __device__ constexpr int plen(int n) { return (n + 2 * kZPad) + ((n + 2 * kZPad) >> 5) + 1; }
#else
// r02-k layout: no pad words; a thread's window of 67 odd samples starts at a multiple of four, so it is read as 17
// LDS.128 (the 32 lanes of a warp read 512 contiguous bytes: conflict-free) instead of 67 scalar loads
__device__ __forceinline__ int ppos(int q) { return q + kZPad; }
__host__ __device__ constexpr int plen(int n) { return (n + 2 * kZPad + 3) & ~3; }
#endif

// de-interleaved signals of octaves 1..6 (lengths 8000 .. 250 -> halves 4000 .. 125)
constexpr int kHalf1 = 4000, kHalf2 = 2000, kHalf3 = 1000, kHalf4 = 500, kHalf5 = 250, kHalf6 = 125;

// Global layout of the decimated signals of one segment (workspace `dec`): octave o = 1..6 in natural sample order,
// kGPad zero samples on either side (librosa's centre padding; the workspace is zeroed once at allocation and the pads
// are never written), so the CQT frame loader needs no bounds checks.
constexpr int kGPad = 256;
__host__ __device__ constexpr int goff(int o) {          // offset of octave o's first pad sample
    int off = 0;
    for (int i = 1; i < o; ++i) off += (16000 >> i) + 2 * kGPad;
    return off;
}
static_assert(goff(7) == 18822, "1 s layout: 18,822 floats per segment");
// same layout for a segment of L samples (long mode; L >> 6 is even for L = 16000 d)
__host__ __device__ inline int goff_len(int o, int L) {
    int off = 0;
    for (int i = 1; i < o; ++i) off += (L >> i) + 2 * kGPad;
    return off;
}
int cens_dec_floats_per_segment(int L) { return (goff_len(7, L) + 3) & ~3; }

// ---------------------------------------------------------------------------------------------- k_cens_dec
struct DecSmem {
    float a[2 * plen(8000)];                           // input (even | odd); later octaves 2..6
    float o1[2 * plen(kHalf1)];                        // octave 1
};
#ifndef BPC_DEC_THREADS
#define BPC_DEC_THREADS 256
#endif
constexpr int kDecThreads = BPC_DEC_THREADS;      // r01 (scalar window loads, 156-168 registers): 2 CTAs x 6 warps was what the registers allowed

// One decimation stage: in (de-interleaved, half-length hin, i.e. 2 hin samples) -> out (de-interleaved, hin samples)
// and to global memory in natural order.
// out[n] = f32( (h63 * E[n] + sum_m g[m] * (O[n-1-m] + O[n+m])) / sqrt(0.5) ),  E[q] = in[2q], O[q] = in[2q+1].
__device__ __forceinline__ void decimate_stage(const float* __restrict__ inE, const float* __restrict__ inO, int hin,
                                               float* __restrict__ outE, float* __restrict__ outO,
                                               float* __restrict__ outG, int tid) {
    const double inv_s = 1.0 / sqrt(0.5);
    for (int n0 = kOutPerThread * tid; n0 < hin; n0 += kOutPerThread * kDecThreads) {
#ifdef BPC_DEC_PAD32
        double w[kWin];
#pragma unroll
        for (int q = 0; q < kWin; ++q) w[q] = (double)inO[ppos(n0 - 32 + q)];   // zero pads cover q < 0 and q >= hin
        float ev[kOutPerThread];
#pragma unroll
        for (int p = 0; p < kOutPerThread; ++p) ev[p] = inE[ppos(n0 + p)];
#else
        static_assert(kOutPerThread == 4 && kZPad % 4 == 0 && kWin <= 68, "LDS.128 windows");
        double w[68];
        const float4* o4 = reinterpret_cast<const float4*>(inO + ppos(n0 - 32));   // zero pads cover q < 0 and q >= hin
#pragma unroll
        for (int q4 = 0; q4 < 17; ++q4) {
            const float4 t4 = o4[q4];
            w[4 * q4] = (double)t4.x; w[4 * q4 + 1] = (double)t4.y; w[4 * q4 + 2] = (double)t4.z; w[4 * q4 + 3] = (double)t4.w;
        }
        const float4 e4 = *reinterpret_cast<const float4*>(inE + ppos(n0));
        const float ev[kOutPerThread] = {e4.x, e4.y, e4.z, e4.w};
#endif
        float v[kOutPerThread];
#pragma unroll
        for (int p = 0; p < kOutPerThread; ++p) {
            double acc = c_hb_centre * (double)ev[p];
#pragma unroll
            for (int m = 0; m < 32; ++m) acc = fma(c_hb_odd[m], w[p + 31 - m] + w[p + 32 + m], acc);
            v[p] = (float)(acc * inv_s);
        }
#ifndef BPC_DEC_PAD32
        if (n0 + 3 < hin) {                                    // all four outputs exist: contiguous vector stores
            *reinterpret_cast<float2*>(outE + ppos(n0 >> 1)) = make_float2(v[0], v[2]);
            *reinterpret_cast<float2*>(outO + ppos(n0 >> 1)) = make_float2(v[1], v[3]);
            *reinterpret_cast<float4*>(outG + n0) = make_float4(v[0], v[1], v[2], v[3]);
            continue;
        }
#endif
        // hin is a multiple of 4 for every stage but the last (250 outputs): guard per pair
#pragma unroll
        for (int p = 0; p < kOutPerThread; p += 2) {
            const int n = n0 + p;
            if (n < hin) {
                outE[ppos(n >> 1)] = v[p];
                if (n + 1 < hin) outO[ppos(n >> 1)] = v[p + 1];
                if (n + 1 < hin) *reinterpret_cast<float2*>(outG + n) = make_float2(v[p], v[p + 1]);
                else outG[n] = v[p];
            }
        }
    }
}

__global__ void __launch_bounds__(kDecThreads, 2) k_cens_dec(const float* __restrict__ y, Geometry g, Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecSmem& S = *reinterpret_cast<DecSmem*>(smem_raw);
    const int tid = threadIdx.x, b = blockIdx.x, L = g.L;
    const float2* y2 = reinterpret_cast<const float2*>(y + (size_t)b * L);
    float* G = ws.dec + (size_t)b * cens_dec_stride(ws);
    for (int i = tid; i < 2 * plen(8000); i += kDecThreads) S.a[i] = 0.f;
    for (int i = tid; i < 2 * plen(kHalf1); i += kDecThreads) S.o1[i] = 0.f;
    __syncthreads();
    {
        float* E0 = S.a;
        float* O0 = S.a + plen(8000);
        load_f2_batched<8>(y2, 8000, (L + 1) / 2, tid, kDecThreads, [&](int m, float2 v) {
            E0[ppos(m)] = v.x;
            O0[ppos(m)] = v.y;
        });
    }
    __syncthreads();
    decimate_stage(S.a, S.a + plen(8000), 8000, S.o1, S.o1 + plen(kHalf1), G + goff(1) + kGPad, tid);
    __syncthreads();
    // octaves 2..6 reuse the input's storage (dead after stage 1); each needs its pads zeroed again
    for (int i = tid; i < 2 * plen(8000); i += kDecThreads) S.a[i] = 0.f;
    __syncthreads();
    float* E2 = S.a;                 float* O2 = E2 + plen(kHalf2);
    float* E3 = O2 + plen(kHalf2);   float* O3 = E3 + plen(kHalf3);
    float* E4 = O3 + plen(kHalf3);   float* O4 = E4 + plen(kHalf4);
    float* E5 = O4 + plen(kHalf4);   float* O5 = E5 + plen(kHalf5);
    float* E6 = O5 + plen(kHalf5);   float* O6 = E6 + plen(kHalf6);
    decimate_stage(S.o1, S.o1 + plen(kHalf1), kHalf1, E2, O2, G + goff(2) + kGPad, tid);
    __syncthreads();
    decimate_stage(E2, O2, kHalf2, E3, O3, G + goff(3) + kGPad, tid);
    __syncthreads();
    decimate_stage(E3, O3, kHalf3, E4, O4, G + goff(4) + kGPad, tid);
    __syncthreads();
    decimate_stage(E4, O4, kHalf4, E5, O5, G + goff(5) + kGPad, tid);
    __syncthreads();
    decimate_stage(E5, O5, kHalf5, E6, O6, G + goff(6) + kGPad, tid);
}

// Long mode: one launch per decimation stage, a thread per output sample, signals in global memory in natural order
// (same taps, same pairing and accumulation order as decimate_stage).  in: n_in samples starting at in[0] (no pads are
// assumed: out-of-range samples read as 0); out: n_in / 2 samples.
__global__ void __launch_bounds__(256) k_dec_stage_long(const float* __restrict__ in_base, size_t in_stride, int in_off,
                                                        int n_in, float* __restrict__ out_base, size_t out_stride,
                                                        int out_off) {
    const int b = blockIdx.y;
    const float* in = in_base + (size_t)b * in_stride + in_off;
    float* out = out_base + (size_t)b * out_stride + out_off;
    const double inv_s = 1.0 / sqrt(0.5);
    const int n_out = n_in >> 1;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_out; n += gridDim.x * blockDim.x) {
        auto O = [&](int q) -> double { const int i = 2 * q + 1; return (i >= 0 && i < n_in) ? (double)__ldg(in + i) : 0.0; };
        double acc = c_hb_centre * (double)__ldg(in + 2 * n);
#pragma unroll 4
        for (int m = 0; m < 32; ++m) acc = fma(c_hb_odd[m], O(n - 1 - m) + O(n + m), acc);
        out[n] = (float)(acc * inv_s);
    }
}

// Long mode, r02: the same stage as a tiled kernel -- a CTA produces 1024 consecutive outputs of one segment from a
// shared-memory copy of the 2 * 1024 + 128 input samples it needs (de-interleaved into even / odd samples, padded like
// the 1 s kernel's arrays), every thread four outputs from a 67-sample FP64 register window: 17 shared-memory loads per
// output instead of 65 global ones (k_dec_stage_long: 2.25 ms for the six stages of a step against 0.55 ms for the 1 s
// decimator; this kernel: see DESIGN section 8).  Same taps, pairing and accumulation order: bit-identical output.
constexpr int kDecTile = 1024, kDecTileThreads = kDecTile / kOutPerThread;
__host__ __device__ constexpr int dpad(int i) { return i + (i >> 5); }
__global__ void __launch_bounds__(kDecTileThreads) k_dec_tile_long(const float* __restrict__ in_base, size_t in_stride,
                                                                   int in_off, int n_in, float* __restrict__ out_base,
                                                                   size_t out_stride, int out_off) {
    __shared__ float sE[dpad(kDecTile) + 1];
    __shared__ float sO[dpad(kDecTile + 64) + 1];
    const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * kDecTile;
    const float* in = in_base + (size_t)b * in_stride + in_off;
    float* out = out_base + (size_t)b * out_stride + out_off;
    const int n_out = n_in >> 1;
    // pair q = (in[2q], in[2q+1]); the tile needs O[t0 - 32 .. t0 + 1024 + 31] and E[t0 .. t0 + 1023]
    for (int i = tid; i < kDecTile + 64; i += kDecTileThreads) {
        const int q = t0 - 32 + i;
        float2 v = make_float2(0.f, 0.f);
        if (q >= 0 && 2 * q + 1 < n_in) v = __ldg(reinterpret_cast<const float2*>(in) + q);
        sO[dpad(i)] = v.y;
        if (i >= 32 && i < kDecTile + 32) sE[dpad(i - 32)] = v.x;
    }
    __syncthreads();
    const double inv_s = 1.0 / sqrt(0.5);
    const int n0 = kOutPerThread * tid;
    double w[kWin];
#pragma unroll
    for (int q = 0; q < kWin; ++q) w[q] = (double)sO[dpad(n0 + q)];
    float v[kOutPerThread];
#pragma unroll
    for (int p = 0; p < kOutPerThread; ++p) {
        double acc = c_hb_centre * (double)sE[dpad(n0 + p)];
#pragma unroll
        for (int m = 0; m < 32; ++m) acc = fma(c_hb_odd[m], w[p + 31 - m] + w[p + 32 + m], acc);
        v[p] = (float)(acc * inv_s);
    }
#pragma unroll
    for (int p = 0; p < kOutPerThread; ++p)
        if (t0 + n0 + p < n_out) out[t0 + n0 + p] = v[p];
}

// ---------------------------------------------------------------------------------------------- k_cens (CQT)
constexpr int kCensThreads = 256, kCensTeams = kCensThreads / 16;
constexpr int kLoFirstOct = 4, kLoOcts = kCqtOctaves - kLoFirstOct;   // octaves of k_cens_lo (1 s mode)

// ---- the sparse basis product  C[r, t] = sum_k B[r, k] X_t[k]  with LANES = FRAMES (r02-h)
// The spectra of sixteen frames sit in shared memory at a row pitch of 89 float2 (178 words: the sixteen frames of a
// half-warp fall into sixteen different bank pairs), a thread takes one frame and THREE adjacent basis rows: one
// conflict-free data load feeds twelve FMAs, the weights are half-warp broadcasts.  The band of row r starts at bin
// cqt_start[r]; the three rows of a triple start within 7 bins of each other (all 100 tunings), so their weights are
// staged in rows padded with zeros on either side and walked over the union of the three bands -- a zero weight adds
// +-0, so every sum is bit-identical to the per-row product the r01 kernel ran with lanes = rows (whose data loads
// were 2.5-way bank conflicts and whose weight loads doubled the wavefronts: 0.19 wavefronts per complex MAC against
// 0.08 here).
constexpr int kCqSpecPitch = 89;
static_assert(kBinSpan + 3 <= kCqSpecPitch, "bands may read up to three bins past kBinHi (zero weights there)");
static_assert(kCqBinsPerOct == kCqtBinsPerOct && kCqOctaves == kCqtOctaves, "kernels.cuh mirrors tables.hpp");

// Stage the first `bytes` of the block of tuning `tun` (bulk-TMA, SASS UBLKCP).  `bar` must have been initialised
// (mbar_init by one thread + a CTA barrier); every thread waits for the data.
__device__ __forceinline__ void cq_stage(CqBlock& Q, const Tables& tb, int tun, uint32_t bytes, uint64_t* bar, int tid) {
    if (tid == 0) {
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(&Q, tb.cq_blocks + tun, bytes, bar);
    }
    mbar_wait(bar, 0);
}

// |CQ| of rows 3 p, 3 p + 1, 3 p + 2 of octave o for the frame whose spectrum row is `sp` (= spec row + tri_s[p]).
// w0 = wpad[o & 1] + 3 p kCqWPitch + kCqPadL; sl = 1 / sqrt(lengths) of row 3 p of this octave.
__device__ __forceinline__ void cq_triple(const float2* __restrict__ w0, int td1, int td2, int tu,
                                          const float2* __restrict__ sp, int o, const double* __restrict__ sl,
                                          float& m0, float& m1, float& m2) {
    const float2* w1 = w0 + kCqWPitch - td1;
    const float2* w2 = w0 + 2 * kCqWPitch - td2;
    BPC_ASSERT(td1 >= 0 && td2 >= td1 && td2 <= kCqPadL && tu >= 1 && tu <= kCqWPitch - kCqPadL);   // taps inside the padded rows
    float cr0 = 0.f, ci0 = 0.f, cr1 = 0.f, ci1 = 0.f, cr2 = 0.f, ci2 = 0.f;
#pragma unroll 4
    for (int u = 0; u < tu; ++u) {
        const float2 d = sp[u];
        const float2 a = w0[u], c = w1[u], e = w2[u];
        cr0 = fmaf(a.x, d.x, cr0); cr0 = fmaf(-a.y, d.y, cr0);
        ci0 = fmaf(a.x, d.y, ci0); ci0 = fmaf(a.y, d.x, ci0);
        cr1 = fmaf(c.x, d.x, cr1); cr1 = fmaf(-c.y, d.y, cr1);
        ci1 = fmaf(c.x, d.y, ci1); ci1 = fmaf(c.y, d.x, ci1);
        cr2 = fmaf(e.x, d.x, cr2); cr2 = fmaf(-e.y, d.y, cr2);
        ci2 = fmaf(e.x, d.y, ci2); ci2 = fmaf(e.y, d.x, ci2);
    }
    // complex64 response, then V /= sqrt(lengths) (complex128 math, complex64 store), then |V|
    const float pow2 = (float)(1 << (o >> 1));
    m0 = c64_abs_f32((float)((double)(cr0 * pow2) * sl[0]), (float)((double)(ci0 * pow2) * sl[0]));
    m1 = c64_abs_f32((float)((double)(cr1 * pow2) * sl[1]), (float)((double)(ci1 * pow2) * sl[1]));
    m2 = c64_abs_f32((float)((double)(cr2 * pow2) * sl[2]), (float)((double)(ci2 * pow2) * sl[2]));
}

// ~108 KB and <= 128 registers: two 256-thread CTAs per SM, so a 592-segment chunk is exactly two full waves (at three
// 128-thread CTAs per SM the second wave ran one third full)
struct CensSmem {
    union {
        double2 xch[kCensTeams][16 * 17];              // FFT exchange buffers (CQT phase)
        struct {                                       // CENS post-processing, after the CQT phase
            float chroma[12 * kMaxFrames];
            float quant[12 * kMaxFrames];
        } post;
    } u;
    float2 spec[kCensTeams][kCqSpecPitch];             // complex64 STFT bins kBinLo.. of the round's sixteen frames
    float m2[2][kCensTeams][kCqTriples];               // |CQ| of row 3 p + 2 (goes to chroma p + 1), by octave parity
    float csum[12 * kMaxFrames];                       // folded chroma sums of the CQT phase
    CqBlock cq;                                        // padded basis rows, row triples, 1 / sqrt(lengths), hann(43) / sum
    uint64_t bar;
    double dscratch[32];
    float fscratch[32];
};
static_assert(sizeof(CensSmem) <= 111 * 1024, "two CTAs per SM");

template <bool LONG>
__global__ void __launch_bounds__(kCensThreads, 2) k_cens(const float* __restrict__ y, Geometry g, Tables tb,
                                                          Workspace ws, float* feats, int phase, int n_oct) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CensSmem& S = *reinterpret_cast<CensSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, L = g.L, T = g.T;
    const float* yb = y + (size_t)b * L;
    const float2* y2 = reinterpret_cast<const float2*>(yb);
    const float* G = ws.dec + (size_t)b * cens_dec_stride(ws);
    // [12, T] tiles: shared memory (1 s), the segment's global scratch region in long mode
    float* lbase = ws.scratch + (size_t)b * ws.scratch_stride;
    float* p_csum = LONG ? lbase : S.csum;
    float* p_chroma = LONG ? lbase + 12 * (size_t)T : S.u.post.chroma;
    float* p_quant = LONG ? lbase + 24 * (size_t)T : S.u.post.quant;

    // ---- stage the tables of this segment's tuning (padded basis rows, 1 / sqrt(lengths), smoothing window)
    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();
    cq_stage(S.cq, tb, ws.tuning[b * 2 + 1], (uint32_t)sizeof(CqBlock), &S.bar, tid);

    // ---- CQT -> chroma fold.  A round is sixteen frames: per octave every half-warp transforms its frame
    // (team_fft<16>) and leaves the 85 touched bins in `spec`; then the basis product runs over the round's sixteen
    // spectra with lanes = frames (thread = frame pf, row triple pp; cq_triple above).
    const int h = lane & 15, team = tid >> 4;
    const int partner = (lane & 16) | ((16 - h) & 15);
    double2* xch = S.u.xch[team];
    double2 tw[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tw[k1] = __ldg(tb.tw256 + ((h * k1) & 255));
    const double2 wp = __ldg(tb.ptw512 + h);
    const double2 wl = make_double2(wp.y, -wp.x);                         // -i * exp(-2 pi i h / 512)
    float2* spec = S.spec[team];
    const int pf = tid & 15, pp = tid >> 4;                               // product: frame of the round, row triple
    const bool pact = pp < kCqTriples;                                    // warps 6, 7 only transform
    const int pq = pact ? pp : 0;
    const int ts = S.cq.tri_s[pq], td1 = S.cq.tri_d1[pq], td2 = S.cq.tri_d2[pq], tu = S.cq.tri_u[pq];
    // long mode: phase 1 = the CQT frames of this CTA's share (grid (segment, part)), phase 2 = post-processing
    // (trip count from blockIdx only: the barriers, shuffles and warp barriers inside are convergent)
    const int t_first = LONG ? (int)blockIdx.y * kCensTeams : 0;
    const int t_step = kCensTeams * (LONG ? (int)gridDim.y : 1);
    for (int t0 = t_first; t0 < T && phase != 2; t0 += t_step) {
        const int t = t0 + team;
        const bool valid = t < T;
        float csum = 0.f, pm0 = 0.f, pm1 = 0.f;                  // thread (pf, pp): chroma pp of frame t0 + pf
#pragma unroll 1
        for (int o = 0; o < n_oct; ++o) {
            const int c0 = (valid ? t : 0) * (128 >> o) - 128;   // first complex sample of the frame
            double2 a[16];
            if (o == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int q = c0 + h + 16 * j;
                    float2 v = make_float2(0.f, 0.f);
                    if (q >= 0 && 2 * q < L) v = __ldg(y2 + q);
                    a[j] = make_double2((double)v.x, (double)v.y);   // window = 'ones'
                }
            } else {
                // c0 >= -128 complex samples = -kGPad samples: the zero pads are the centre padding
                const float2* src = reinterpret_cast<const float2*>(G + (LONG ? goff_len(o, L) : goff(o)) + kGPad) + c0 + h;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 v = __ldg(src + 16 * j);
                    a[j] = make_double2((double)v.x, (double)v.y);
                }
            }
            team_fft<16>(a, tw, 1, xch, h);
            __syncthreads();                                     // the previous octave's product has read `spec`
            auto emit = [&](int k, double2 t2) {
                if (k >= kBinLo && k < kBinLo + kCqSpecPitch)    // bins past kBinHi only ever meet zero weights
                    spec[k - kBinLo] = make_float2((float)(0.5 * t2.x), (float)(0.5 * t2.y));   // complex64 STFT
            };
            team_rsplit<16, 3, 9>(a, wl, h, partner, emit);
            __syncthreads();
            if (pact) {
                // cq_to_chroma: chroma c sums bins {3c-1, 3c, 3c+1} (mod 36) of every octave, octave by octave
                if (o > 0) csum += S.m2[(o - 1) & 1][pf][(pp + kCqTriples - 1) % kCqTriples] + pm0 + pm1;
                float m2v;
                BPC_ASSERT(ts >= 0 && ts + tu <= kCqSpecPitch);
                cq_triple(S.cq.wpad[o & 1] + 3 * pp * kCqWPitch + kCqPadL, td1, td2, tu, &S.spec[pf][ts], o,
                          S.cq.inv_sl + (kCqtBins - kCqtBinsPerOct * (o + 1)) + 3 * pp, pm0, pm1, m2v);
                S.m2[o & 1][pf][pp] = m2v;
            }
        }
        __syncthreads();
        if (pact) {
            csum += S.m2[(n_oct - 1) & 1][pf][(pp + kCqTriples - 1) % kCqTriples] + pm0 + pm1;
            const int tt = t0 + pf;
            if (tt < T) {
                if (n_oct < kCqtOctaves) {                       // octaves 4-6 come from k_cens_lo, added in octave order
                    const float* lo = ws.cens_lo + ((size_t)b * kLoOcts * 12 + pp) * T + tt;
#pragma unroll
                    for (int q2 = 0; q2 < kLoOcts; ++q2) csum += lo[q2 * 12 * T];
                }
                p_csum[pp * T + tt] = csum;
            }
        }
    }
    if (LONG && phase == 1) return;
    __syncthreads();
    // ---- CENS post-processing per column: L1 normalise, quantise
    for (int t = tid; t < T; t += kCensThreads) {
        double l1 = 0.0;
        for (int c = 0; c < 12; ++c) l1 += fabs((double)p_csum[c * T + t]);
        if (l1 < 1.17549435e-38) l1 = 1.0;
        for (int c = 0; c < 12; ++c) {
            const float v = (float)((double)p_csum[c * T + t] / l1);
            p_quant[c * T + t] = 0.25f * (float)((v > 0.4f) + (v > 0.2f) + (v > 0.1f) + (v > 0.05f));
        }
    }
    __syncthreads();
    // 41 non-zero taps of hann(43) / sum, scipy.ndimage.convolve(mode='constant') along time
    for (int i = tid; i < 12 * T; i += kCensThreads) {
        const int c = i / T, t = i - c * T;
        double acc = 0.0;
        const int jlo = max(0, t + 21 - (T - 1)), jhi = min(42, t + 21);
        for (int j = jlo; j <= jhi; ++j) acc += S.cq.swin[j] * (double)p_quant[c * T + t + 21 - j];
        p_chroma[i] = (float)acc;
    }
    __syncthreads();
    // L2 normalise each column
    for (int t = tid; t < T; t += kCensThreads) {
        double l2 = 0.0;
        for (int c = 0; c < 12; ++c) l2 += (double)p_chroma[c * T + t] * (double)p_chroma[c * T + t];
        l2 = sqrt(l2);
        if (l2 < 1.17549435e-38) l2 = 1.0;
        for (int c = 0; c < 12; ++c) p_quant[c * T + t] = (float)((double)p_chroma[c * T + t] / l2);
    }
    __syncthreads();
    if (ws.dbg_chroma_cens) {
        float* d = ws.dbg_chroma_cens + (size_t)b * 12 * T;
        for (int i = tid; i < 12 * T; i += kCensThreads) d[i] = p_quant[i];
    }
    // ---- row-wise z-score -> rows 12..23; pad rows 24..127 with the min over all 24 normalised rows
    float mn = FLT_MAX;
    float* o = plane_ptr(feats, b, BPC_CH_CHROMA, T);
    const bool stats = !LONG && ws.stats_acc != nullptr;
    StatAcc ac;
    ac.init();
    for (int r = warp; r < 12; r += kCensThreads / 32) {
        const ZTerm z = np_row_zterm(p_quant + r * T, T, lane);
        for (int t = lane; t < T; t += 32) {
            const float v = z(p_quant[r * T + t]);
            o[(12 + r) * T + t] = v;
            mn = fminf(mn, v);
            if (stats) ac.add(v);
        }
    }
    mn = block_min(mn, S.fscratch);
    const float fill = fminf(mn, ws.chroma_min[b * 2 + 0]);
    for (int i = 24 * T + tid; i < kPlaneRows * T; i += kCensThreads) o[i] = fill;
    if (tid == 0) ws.chroma_min[b * 2 + 1] = mn;
    if (stats) {                                                   // rows 0..11 were counted by the chroma_stft role
        if (tid == 0) ac.add_n(fill, (kPlaneRows - 24) * T);
        stat_flush_block(ac, ws.stats_acc + 5 * BPC_CH_CHROMA, S.dscratch, S.fscratch);
    }
}

// ---------------------------------------------------------------------------------------------- k_cens_lo (r02-h)
// The three lowest octaves (o = 4, 5, 6: 1000 / 500 / 250 samples, hop h = 16 / 8 / 4 against n_fft = 512) without FFTs.
// Consecutive frames of these octaves overlap in 496 .. 508 of their 512 samples, and the sparsified bases touch only
// the 85 bins kBinLo .. kBinHi, so each of those bins is carried from frame to frame by the sliding-DFT recurrence of
// the rectangular window ('ones', which is what librosa's CQT uses):
//     X_{j+1}[k] = W^{-hk} ( X_j[k] + sum_{n<h} (x[s_j + 512 + n] - x[s_j + n]) W^{nk} ),   W = exp(-2 pi i / 512),
// started from the window [-512, 0) that lies wholly in the zero padding (X = 0) and brought to frame 0 (window
// [-256, 256)) in 16 hops of 16 samples.  Per bin and frame that is 2 h + 8 FP64 instructions with the twiddles W^{nk}
// (the same for every octave) in registers -- 40 / 24 / 4 against the ~6.9 k of an FFT-512 + split per frame shared by 85
// bins (81 per bin) -- and no shared-memory exchange at all.  In octave 6 every frame holds the whole 250-sample
// signal: after the start-up the recurrence is a pure rotation.
// Rounding: one unit rotation per hop, <= 79 hops: 3e-14 of the largest |X| the bin has seen, the same order as the
// FFT's own error and nine orders below the complex64 rounding that follows (A/B against the FFT form: 0 of 1,048,320
// chroma values differ on 130 segments, tools/cens_ab.py).
// Threads 0-95 own the bins of octave 4, threads 96-191 those of octaves 5 AND 6 (2844 against 2664 FP64 instructions
// per bin).  Frames are produced in batches of 16 into a shared spectrum tile; the basis product runs with LANES =
// FRAMES (row pitch 89 float2: the sixteen frames of a half-warp fall into sixteen bank pairs) and three adjacent basis
// rows per thread: one conflict-free data load feeds twelve FMAs, the weights are half-warp broadcasts from rows padded
// with zeros on either side (a zero weight adds +-0: the sums are those of the per-row product).  The three per-octave
// chroma sums go to Workspace::cens_lo and k_cens adds them after its own octaves 0-3 in the order the single-kernel
// version used (bit-identical folding).
constexpr int kLoGroup = 96, kLoThreads = 2 * kLoGroup;                             // 192
constexpr int kLoBatch = 16, kLoSpecPitch = kCqSpecPitch;
constexpr int kLoFrames = 63;                                                       // 1 s mode only: T = 63
constexpr int kLoTriples = kCqTriples;
constexpr int kLoD0 = 256 + kLoFrames * 16, kLoD1 = 256 + kLoFrames * 8, kLoD2 = 256;
static_assert(kBinSpan <= kLoSpecPitch && kLoSpecPitch <= kLoGroup, "one thread per touched bin");
static_assert(kLoOcts * kLoTriples * kLoBatch == 3 * kLoThreads, "basis product: three items per thread");

struct CensLoSmem {
    double d0[kLoD0], d1[kLoD1], d2[kLoD2];                      // x[i] - x[i - 512] per octave (octave 6: x[i], i < 256)
    float2 spec[kLoOcts][kLoBatch][kLoSpecPitch];
    alignas(16) unsigned char cq_bytes[kCqBlockLoBytes];         // the head of the tuning's CqBlock: basis rows, triples,
                                                                 // 1 / sqrt(lengths) of octaves 6, 5, 4
    float m2[kLoOcts][kLoBatch][kLoTriples];                     // |CQ| of row 3 p + 2 (goes to chroma p + 1)
    uint64_t bar;
};
static_assert(sizeof(CensLoSmem) <= 75 * 1024 + 640, "three CTAs per SM");

// `hops` hops of H samples of one bin: (store X as complex64,) then X <- rot (X + sum_n d[n] tw[n]).
template <int H, bool EMIT>
__device__ __forceinline__ void lo_hops(const double* __restrict__ d, const double2 (&tw)[16], double2 rot, double2& X,
                                        int hops, float2* __restrict__ out, bool store) {
#pragma unroll 2
    for (int i = 0; i < hops; ++i) {
        if (EMIT && store) out[i * kLoSpecPitch] = make_float2((float)X.x, (float)X.y);
        const double2* dp = reinterpret_cast<const double2*>(d + i * H);
        BPC_ASSERT((reinterpret_cast<uintptr_t>(dp) & 15) == 0);
        double ar0 = 0.0, ai0 = 0.0, ar1 = 0.0, ai1 = 0.0;
#pragma unroll
        for (int n = 0; n < H; n += 2) {
            const double2 dd = dp[n >> 1];
            ar0 = fma(dd.x, tw[n].x, ar0);         ai0 = fma(dd.x, tw[n].y, ai0);
            ar1 = fma(dd.y, tw[n + 1].x, ar1);     ai1 = fma(dd.y, tw[n + 1].y, ai1);
        }
        const double ar = X.x + (ar0 + ar1), ai = X.y + (ai0 + ai1);
        X.x = fma(ar, rot.x, -(ai * rot.y));
        X.y = fma(ar, rot.y, ai * rot.x);
    }
}

__global__ void __launch_bounds__(kLoThreads, 3) k_cens_lo(Geometry g, Tables tb, Workspace ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CensLoSmem& S = *reinterpret_cast<CensLoSmem*>(smem_raw);
    const CqBlock& cq = *reinterpret_cast<const CqBlock*>(S.cq_bytes);
    const int tid = threadIdx.x, b = blockIdx.x, T = g.T;
    const float* G = ws.dec + (size_t)b * cens_dec_stride(ws);
    // tables of this segment's tuning: the bulk copy runs while the difference signals are built
    if (tid == 0) {
        mbar_init(&S.bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&S.bar, (uint32_t)kCqBlockLoBytes);
        tma_bulk_g2s(S.cq_bytes, tb.cq_blocks + ws.tuning[b * 2 + 1], (uint32_t)kCqBlockLoBytes, &S.bar);
    }
#pragma unroll
    for (int q = 0; q < kLoOcts; ++q) {
        const int n = 16000 >> (kLoFirstOct + q);
        const float* x = G + goff(kLoFirstOct + q) + kGPad;
        double* d = q == 0 ? S.d0 : (q == 1 ? S.d1 : S.d2);
        const int len = q == 0 ? kLoD0 : (q == 1 ? kLoD1 : kLoD2);
        // (loads of a batch of four trips first: the plain loop waited for one L2 round trip per element)
        for (int i0 = tid; i0 < len; i0 += 4 * kLoThreads) {
            float a[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kLoThreads;
                a[u] = (i < len && i < n) ? __ldg(x + i) : 0.f;
                c[u] = (i < len && i >= 512 && i - 512 < n) ? __ldg(x + i - 512) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kLoThreads;
                if (i < len) d[i] = (double)a[u] - (double)c[u];
            }
        }
    }
    __syncthreads();                                            // d0 .. d2 and the barrier object are visible
    mbar_wait(&S.bar, 0);                                       // ... and so are the tables
    // this thread's bin: twiddles W^{nk}, n < 16, and the hop rotations W^{-hk}
    const int grp = tid / kLoGroup;                             // warp-uniform: 0 = octave 4, 1 = octaves 5 and 6
    const int kk = tid - grp * kLoGroup;                        // bin - kBinLo (lanes past the span idle along)
    const int k = kBinLo + kk;
    auto wpow = [&](int m) {                                    // exp(-2 pi i m / 512)
        m &= 511;
        const double2 w = __ldg(tb.ptw512 + (m <= 256 ? m : 512 - m));
        return m <= 256 ? w : make_double2(w.x, -w.y);
    };
    double2 tw[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) tw[n] = wpow(n * k);
    auto conj_w = [&](int m) { const double2 w = wpow(m); return make_double2(w.x, -w.y); };
    const double2 rot16 = conj_w(16 * k), rot8 = conj_w(8 * k), rot4 = conj_w(4 * k);
    const bool store = kk < kLoSpecPitch;                       // bins past kBinHi only ever meet zero weights
    const int ks = store ? kk : 0;
    double2 Xa = make_double2(0.0, 0.0), Xb = make_double2(0.0, 0.0);
    // start-up: from the all-zero window to frame 0 in 16 hops of 16 samples (every octave)
    lo_hops<16, false>(grp == 0 ? S.d0 : S.d1, tw, rot16, Xa, 16, nullptr, false);
    if (grp == 1) lo_hops<16, false>(S.d2, tw, rot16, Xb, 16, nullptr, false);
    float* lo = ws.cens_lo + (size_t)b * kLoOcts * 12 * T;
    const int f = tid & 15, p = tid >> 4;                       // basis product: frame of the batch, row triple
    const int ts = cq.tri_s[p], td1 = cq.tri_d1[p], td2 = cq.tri_d2[p], tu = cq.tri_u[p];
    for (int t0 = 0; t0 < T; t0 += kLoBatch) {
        const int nf = min(kLoBatch, T - t0);
        BPC_ASSERT(256 + (t0 + nf) * 16 <= kLoD0 && 256 + (t0 + nf) * 8 <= kLoD1 && t0 + nf <= T);
        if (grp == 0) {
            lo_hops<16, true>(S.d0 + 256 + t0 * 16, tw, rot16, Xa, nf, &S.spec[0][0][ks], store);
        } else {
            lo_hops<8, true>(S.d1 + 256 + t0 * 8, tw, rot8, Xa, nf, &S.spec[1][0][ks], store);
            for (int i = 0; i < nf; ++i) {                      // octave 6: rotation only
                if (store) S.spec[2][i][ks] = make_float2((float)Xb.x, (float)Xb.y);
                const double xr = Xb.x, xi = Xb.y;
                Xb.x = fma(xr, rot4.x, -(xi * rot4.y));
                Xb.y = fma(xr, rot4.y, xi * rot4.x);
            }
        }
        __syncthreads();
        float m0[kLoOcts], m1[kLoOcts];
#pragma unroll
        for (int q2 = 0; q2 < kLoOcts; ++q2) {
            const int o = kLoFirstOct + q2;
            float m2v;
            BPC_ASSERT(ts >= 0 && ts + tu <= kLoSpecPitch && f < kLoBatch && p < kLoTriples);
            cq_triple(cq.wpad[o & 1] + 3 * p * kCqWPitch + kCqPadL, td1, td2, tu, &S.spec[q2][f][ts], o,
                      cq.inv_sl + (kCqtOctaves - 1 - o) * kCqtBinsPerOct + 3 * p, m0[q2], m1[q2], m2v);
            S.m2[q2][f][p] = m2v;
        }
        __syncthreads();
        // cq_to_chroma per octave: chroma c = p sums rows {3c-1, 3c, 3c+1} (mod 36), in that order
        if (f < nf) {
#pragma unroll
            for (int q2 = 0; q2 < kLoOcts; ++q2)
                lo[(q2 * 12 + p) * T + t0 + f] = S.m2[q2][f][(p + kLoTriples - 1) % kLoTriples] + m0[q2] + m1[q2];
        }
        // the next batch's slide writes `spec` (free since the barrier above); its basis product, which writes `m2`,
        // comes after that slide's barrier
    }
}

// stage 1: the half-band decimations, 2: the CQT / CENS kernels, 0: both (the per-kernel timing leg launches them
// apart: 4 = k_cens_lo alone, 3 = k_cens alone)
void launch_cens(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                 cudaStream_t st, int stage) {
    static PerDeviceOnce once;
    once.run([&] {
        cudaFuncSetAttribute(k_cens_dec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem));
        cudaFuncSetAttribute(k_cens<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CensSmem));
        cudaFuncSetAttribute(k_cens<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CensSmem));
        cudaFuncSetAttribute(k_cens_lo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CensLoSmem));
    });
    const bool dec = stage == 0 || stage == 1;
    if (dec && g.long_mode) {
        const int L = g.L;
        static const bool tiled = !(std::getenv("BPC_DEC_TILED") && std::atoi(std::getenv("BPC_DEC_TILED")) == 0);
        for (int o = 1; o <= 6; ++o) {
            const int n_in = L >> (o - 1);
            const float* src = o == 1 ? y : ws.dec;
            const size_t sstride = o == 1 ? (size_t)L : (size_t)ws.dec_stride;
            const int soff = o == 1 ? 0 : goff_len(o - 1, L) + kGPad;
            if (tiled) {
                const int tiles = (n_in / 2 + kDecTile - 1) / kDecTile;
                k_dec_tile_long<<<dim3(tiles, n), kDecTileThreads, 0, st>>>(src, sstride, soff, n_in, ws.dec,
                                                                            (size_t)ws.dec_stride, goff_len(o, L) + kGPad);
            } else {
                const int blocks = std::min(64, (n_in / 2 + 255) / 256);
                k_dec_stage_long<<<dim3(blocks, n), 256, 0, st>>>(src, sstride, soff, n_in, ws.dec, (size_t)ws.dec_stride,
                                                                  goff_len(o, L) + kGPad);
            }
        }
        note_launch(6);
    } else if (dec) {
        k_cens_dec<<<n, kDecThreads, sizeof(DecSmem), st>>>(y, g, ws);
        note_launch();
    }
    if (stage == 1) return;
    if (g.long_mode) {
        // a CTA takes kCensTeams frames per round: no more parts than rounds (16 parts left half of the teams of a
        // 2 s segment, 126 frames, without a frame: 3.55 ms against 2.35 ms for the same samples at 30 s)
        const int parts = std::max(1, std::min(16, (g.T + kCensTeams - 1) / kCensTeams));
        k_cens<true><<<dim3(n, parts), kCensThreads, sizeof(CensSmem), st>>>(y, g, tb, ws, feats, 1, kCqtOctaves);
        k_cens<true><<<dim3(n, 1), kCensThreads, sizeof(CensSmem), st>>>(y, g, tb, ws, feats, 2, kCqtOctaves);
        note_launch(2);
    } else {
        // BPC_CENS_LO=0: all seven octaves by FFT in k_cens (the r02-g form)
        static const bool lo = !(std::getenv("BPC_CENS_LO") && std::atoi(std::getenv("BPC_CENS_LO")) == 0);
        const bool use_lo = lo && ws.cens_lo != nullptr && g.T == kLoFrames;
        if (use_lo && stage != 3) {
            k_cens_lo<<<n, kLoThreads, sizeof(CensLoSmem), st>>>(g, tb, ws);
            note_launch();
        }
        if (stage == 4) return;
        k_cens<false><<<n, kCensThreads, sizeof(CensSmem), st>>>(y, g, tb, ws, feats, 0, use_lo ? kLoFirstOct : kCqtOctaves);
        note_launch();
    }
}

}  // namespace bpc
