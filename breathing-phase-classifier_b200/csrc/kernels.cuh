// Device-side table / workspace descriptors and kernel launchers shared between api.cu and the k_*.cu files.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "common.cuh"
#include "../../include/bpc.h"

namespace bpc {

struct BankDev {            // band-form triangular filterbank (see tables.hpp::SparseBank)
    const int* start;       // [rows]
    const int* count;       // [rows]
    const float* w;         // [rows, width]
    const float* wt;        // [width, rows]  (transposed: lanes that own consecutive rows read consecutive words)
    int rows, width;
};

// ---- per-tuning constant block of the CQT kernels (k_cens, k_cens_lo), built on the host (api.cu::build_tables) in
// exactly the layout the kernels keep in shared memory, so that staging it is ONE bulk-TMA copy (r02-j; the kernels
// used to derive it from the band tables in their prologues: 21 % of k_cens's stall samples).
constexpr int kCqBinsPerOct = 36, kCqOctaves = 7, kCqBins = kCqBinsPerOct * kCqOctaves;
constexpr int kCqPadL = 8, kCqWPitch = 33;                 // basis rows padded with zeros: tap index in [-8, 25)
constexpr int kCqTriples = kCqBinsPerOct / 3;
struct alignas(16) CqBlock {
    float2 wpad[2][kCqBinsPerOct * kCqWPitch];             // [0]: basis, [1]: basis * sqrt(2) (odd octaves), (re, im)
    short tri_s[kCqTriples], tri_d1[kCqTriples], tri_d2[kCqTriples], tri_u[kCqTriples];   // row triples: first bin - 60,
                                                           // offsets of rows 3p+1 / 3p+2, taps of the union band
    double inv_sl[kCqBins];                                // 1 / sqrt(lengths), lowest octave first
    double swin[44];                                       // hann(43) / sum (CENS smoothing), one pad entry
};
static_assert(sizeof(CqBlock) % 16 == 0, "bulk copies move multiples of 16 bytes");
constexpr int kCqBlockLoBytes = (int)(offsetof(CqBlock, inv_sl) + 3 * kCqBinsPerOct * sizeof(double));   // k_cens_lo's share
static_assert(kCqBlockLoBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

struct Tables {
    // FFT
    const double* hann512;        // [512]
    const double* hann2048;       // [2048]
    const double* hann2048h;      // [2048] 0.5 * Hann (k_frame2048: the 1/2 of the real-input split folded into the window)
    const double2* rs2048;        // [576] split twiddle w_k = -i exp(-2 pi i k / 2048) factored for FMA butterflies:
                                  //       (c, t) with w = c (t + i) for k < 256 and w = c (1 + i t) for k >= 256
    const double2* tw256;         // exp(-2 pi i j / 256), j < 256
    const double2* ptw512;        // exp(-2 pi i k / 512),  k <= 256
    const double2* ptw2048;       // exp(-2 pi i k / 2048), k <= 1024
    const double2* twa1024;       // [k1][h] = exp(-2 pi i h k1 / 1024), 32 x 32: inter-stage twiddles of team_fft<32>
    const double2* twa256;        // [k1][h] = exp(-2 pi i h k1 / 256), 16 x 16: the same for team_fft<16> (k_logmel_fused)
    const double2* t64a;          // [k1][j] = exp(-2 pi i j k1 / 1024), 16 x 64: stage A -> B twiddles of team64_fft (k_frame2048_w2)
    const double2* t64b;          // [k2][j0] = exp(-2 pi i j0 k2 / 64), 16 x 4: stage B -> C twiddles of team64_fft
    // filterbanks
    BankDev mel_a, mel_b, mel_c, mel_d;
    const float* dct_mel;         // [40, 128]
    const float* dct_time;        // [T, T]  transposed: [t][u]
    const float* dct_time_n;      // [T, T]  [u][t] (K-major B operand of the tensor-core time DCT; long mode only)
    const uint32_t* dct_tiles;    // the same matrix split into tf32 hi / lo halves, pre-tiled in the canonical UMMA layout
    const float* dct_colsum;      // [T] sum_t D[u][t] (long mode: puts the removed row means back, k_tc.cu)
    const float* chroma;          // [100, 12, 257]
    const double* hist_edges;     // [101]
    // CQT
    const int16_t* cqt_start;     // [100, 36]     first rfft bin of each basis row's band
    const float* cqt_re;          // [100, 36, W]  band form: entry j of row r weighs bin cqt_start[r] + j (zeros in gaps)
    const float* cqt_im;          // [100, 36, W]
    const double* cqt_sqrt_len;   // [100, 252]
    const CqBlock* cq_blocks;     // [100]  the same tables per tuning, staged whole by the CQT kernels
    int cqt_gw[3];                // band width needed by the rows 20-35 / 4-19 / 0-3 over all tunings (<= kCqtEllWidth)
    // LPC
    const double* hamming400;     // [400]
    // Hilbert (FFT-8000 = 4^3 * 5^3)
    const float2* tw20a;          // [20][400] exp(-2 pi i k pos / 8000), float32: scipy.signal.hilbert runs a float32 FFT on float32 input
    const unsigned short* h20pos; // [8001] padded storage position of bin k % 8000 after the forward passes (fft20.cuh)
    const float2* tw20b;          // [20][20]  exp(-2 pi i k pos / 400)  (radix-20 passes of the FFT-8000, fft20.cuh)
    const float2* ptw16000f;      // exp(-2 pi i k / 16000), k <= 8000
    // long mode Hilbert (FFT-N, N = L / 2 = 2^a 3^b 5^c): exp(-2 pi i j / N), j < N and exp(-2 pi i k / L), k <= N
    const float2* tw_long;
    const float2* ptw_long;
    // tempogram
    const double* hann384;        // [384]
};

struct Workspace {            // per chunk of `cap` segments
    int cap;
    float* y;                 // [cap, L]   float32 waveform after pad_or_truncate (only when ingest is needed)
    float* mag512;            // [cap, T, kMagStride]
    float* mag_even;          // [cap, (T+1)/2, kMag2048Stride]  |STFT2048| rows of the hop-512 frames (1025 valid bins)
    float* cand36;            // [cap, 2, kMaxCand2048]  1 s mode: piptrack candidates (magnitude, pitch) of k_even2048
    double* frame_feat;       // [cap, T, 20]  per-frame centroid, bandwidth, flatness, contrast peaks / valleys
    float* melD;              // [cap, T, 128] mel-D power columns
    float* dec;               // [cap, dec_stride]  half-band decimated signals of the CQT octaves 1..6 (zero padded)
    int dec_stride;
    float* cens_lo;           // [cap, 3, 12, T] 1 s mode: chroma sums of the CQT octaves 4-6 (k_cens_lo -> k_cens)
    float* lpc_coef;          // [cap, 12, F] 1 s mode: LPC coefficients of every frame (k_lpc_fast / k_lpc_redo -> k_lpc)
    int* lpc_redo;            // [1 + cap F]  count, then the (segment F + frame) ids k_lpc_fast left to the direct method
    int* tuning;              // [cap, 2]   tuning bin for 12 / 36 bins per octave
    float* chroma_min;        // [cap, 2]   min of the normalised chroma_stft / chroma_cens rows
    int* ints;                // [cap, 2]   n_peaks, first-min index
    uint32_t* tc_a;           // long mode: C1 split into tf32 hi / lo tiles for the tensor-core time DCT (k_tc.cu)
    float* tc_mean;           // long mode: [cap * 40] row means of C1, removed before the tensor-core product (k_tc.cu)
    double* stats_acc;        // [(9 + S), 5] dataset statistics accumulated by the producers (1 s mode; nullptr: k_stats does it)
    float* scratch;           // [cap, scratch_stride]  long mode: what the 1 s kernels keep in shared memory
    size_t scratch_stride;
    // debug (raw, un-normalised stages of the last chunk)
    float* dbg_mel_db;        // [cap, 128, T]
    float* dbg_mfcc;          // [cap, 120, T]
    float* dbg_gam;           // [cap, 64, T]
    float* dbg_mod;           // [cap, 40, T]
    float* dbg_chroma_stft;   // [cap, 12, T]
    float* dbg_chroma_cens;   // [cap, 12, T]
    float* dbg_lpc;           // [cap, 12, F]
    float* dbg_onset;         // [cap, T]
};

struct Geometry {
    int L;                    // expected_len
    int T;                    // frames = L / hop + 1
    int hop;                  // 256
    int nscal;                // scalars per segment in the output (36 or padded)
    int lpc_frames;           // len(range(0, L - 400, 160))
    int long_mode;            // L > 16000 (BASELINE config 4): per-segment arrays live in Workspace::scratch, not on chip
};

// feats layout: [B, 9, 128, T]; plane pointer of channel c of segment b
__device__ __forceinline__ float* plane_ptr(float* feats, int b, int c, int T) {
    return feats + ((size_t)b * 9 + c) * (size_t)kPlaneRows * T;
}

// ---- launchers (all enqueue on `st`, no sync).  n = segments in this chunk.
void launch_ingest(const void* wav, int wav_dtype, int64_t L_in, float* y, int n, const Geometry& g, cudaStream_t st);
void launch_stft512(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, cudaStream_t st);
void launch_spec512_consumers(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                              float* scalars, int32_t* status, bool with_chroma, cudaStream_t st);
void launch_stft_db(int n, const Geometry& g, const Workspace& ws, float* stft_db, cudaStream_t st);
void launch_logmel_only(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* mel3, cudaStream_t st);
// BASELINE config 2 in one kernel (1 s mode, L_in == expected_len, 16-byte aligned buffers); false: use the staged path
bool launch_logmel_fused(const void* wav, int wav_dtype, int n, const Geometry& g, const Tables& tb, float* stft_db,
                         float* mel3, cudaStream_t st);
void launch_spec2048(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                     float* scalars, cudaStream_t st);
void launch_seg2048(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats, float* scalars,
                    cudaStream_t st);
void launch_even2048(int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* scalars,
                     int32_t* status, cudaStream_t st);
void launch_time_scalars(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws,
                         float* scalars, int32_t* status, cudaStream_t st);
void launch_hilbert(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* scalars,
                    cudaStream_t st);
void launch_lpc(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                cudaStream_t st);
void launch_cens(const float* y, int n, const Geometry& g, const Tables& tb, const Workspace& ws, float* feats,
                 cudaStream_t st, int stage = 0);
void launch_stats(int n, const Geometry& g, const float* feats, const float* scalars, double* acc, bool planes,
                  cudaStream_t st);
void launch_modspec(int n, const Geometry& g, const Tables& tb, const Workspace& ws, const float* mel_db, float* out,
                    cudaStream_t st);
void launch_pad_scalars(int n, const Geometry& g, float* scalars, cudaStream_t st);

void launch_collate(const float* store_feats, const float* store_scalars, const long long* ia, const long long* ib,
                    int n, int mode, float lam, float oml, int y1, int y2, int x1, int x2, int T, int nscal,
                    float* out_feats, float* out_scalars, cudaStream_t st);

void launch_pad_values(const float* feats, int T, int n, const int* live_dev, float* fill, cudaStream_t st);
void launch_compact_rows(const float* feats, int T, int n, const int* live_host, float* out, cudaStream_t st);

size_t consumer_scratch_floats(int T);
bool modspec_time_tc_enabled(const Tables& tb);
size_t tc_tile_words(int rows, int T);
void launch_tc_prep_b(const Geometry& g, const Tables& tb, uint32_t* out, cudaStream_t st);
void launch_modspec_time_tc(int n, const Geometry& g, const Tables& tb, const Workspace& ws, size_t role0_off,
                            cudaStream_t st);

void launch_resample(const float* x, long long n_in, const double* tab, int p, int q, int half, float* out,
                     long long n_out, cudaStream_t st);

struct WavItem { long long offset, frames; int channels, fmt; };   // payload of one file inside the blob (k_wav.cu)
void launch_wav_decode(const unsigned char* blob, long long blob_bytes, const WavItem* items, int n, int L, float* y,
                       cudaStream_t st);

void upload_cens_constants(const double* taps127);
void upload_lpc_constants(const double* hamming400);
int cens_dec_floats_per_segment(int L);
int64_t launches_issued();   // process-wide counter bumped by every launcher
void note_launch(int n = 1);

}  // namespace bpc
