/*
 * bpc.h -- C ABI of libbpc_b200.so: the B200-native `--precompute` feature-extraction path of
 * dohyeoplim/breathing-phase-classifier (reference paths below are relative to /root/reference/).
 *
 * The reference has no FFI: the path sits behind three Python functions and an on-disk format.  This ABI is what a
 * ctypes binding of those functions needs:
 *
 *   src/precompute/process.py:25-108   process_and_save_npz((file_id, wav_path, target_dir))   -> bpc_precompute*
 *   src/precompute/core.py:19-45       process_dataset_threaded(df, audio_dir, target_dir, ..) -> bpc_precompute* (batched)
 *   src/precompute/methods.py:24-28    pad_or_truncate                                          -> done on device (L_in vs expected_len)
 *   src/precompute/methods.py:48-114   extract_enhanced_scalar_features(y, sr) -> f32[36]       -> `scalars` output
 *   src/precompute/methods.py:116-143  extract_lpc / gammatone / spectral_modulation features   -> bpc_debug_copy (raw stages)
 *   src/precompute/process.py:12-23    module constants                                         -> bpc_params
 *   src/dataset.py:8,25-26,48          consumer contract: sorted-key stacking                    -> feats layout [B,9,128,T]
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the ABI.  Every function returns 0 on success or a negative
 *     bpc_status code; bpc_last_error() gives the message.
 *   - per-segment problems never fail the call (reference: process.py:107-108 returns (id, False, err)); they are
 *     reported as bit flags in `status[B]` (0 = ok).
 *   - a handle owns all constant tables and workspaces (allocated once in bpc_create).  Calls on one handle must be
 *     serialised by the caller; different handles are independent.  Device entry points enqueue on the given
 *     cudaStream_t and do not synchronise.
 *   - there is NO CPU fallback: without a CUDA device bpc_create fails with BPC_ERR_CUDA.
 */
#ifndef BPC_B200_H
#define BPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPC_ABI_VERSION 3   /* 2: streaming host calls (begin / wait); 3: 13 kernel timing ids (k_cens_lo) */

/* feats channel order = sorted .npz keys (dataset.py:26) */
enum bpc_channel {
    BPC_CH_CHROMA = 0, BPC_CH_GAMMATONE = 1, BPC_CH_LPC = 2, BPC_CH_MEL = 3, BPC_CH_MEL_DELTA = 4,
    BPC_CH_MEL_DELTA2 = 5, BPC_CH_MFCC = 6, BPC_CH_MOD_SPEC = 7, BPC_CH_TEMPOGRAM = 8, BPC_NUM_CHANNELS = 9
};
#define BPC_NUM_SCALARS 36   /* methods.py:54-112 appends 8+11+6+4+4+3 values */
#define BPC_PLANE_ROWS 128   /* N_MELS, process.py:15 */

enum bpc_wav_dtype { BPC_WAV_F32 = 0, BPC_WAV_PCM16 = 1 };

enum bpc_status {
    BPC_OK = 0,
    BPC_ERR_ARG = -1,          /* bad argument / unsupported parameter combination */
    BPC_ERR_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
    BPC_ERR_ALLOC = -3,
    BPC_ERR_UNSUPPORTED = -4   /* not sm_100, other constants than the reference's, expected_len not 16000 * 2^a 3^b 5^c */
    /* -5 is retired: the path's one collective (the all-gather of the 1.8 KB statistics accumulator) is issued by the
       host through torch.distributed on the pointer bpc_channel_stats_device returns, not by this library */
};

/* per-segment status bits */
#define BPC_SEG_NONFINITE      1u   /* input had NaN/Inf */
#define BPC_SEG_TUNING_EMPTY   2u   /* pitch_tuning saw an empty frequency set -> tuning 0.0 (librosa warns) */
#define BPC_SEG_CAND_OVERFLOW  4u   /* more piptrack candidates than the list holds (chroma invalid); the lists hold the
                                       combinatorial maximum, so this cannot occur for finite input */
#define BPC_SEG_SILENT         8u   /* all-zero segment */

/* process.py:12-23 / methods.py:10-22 module constants */
typedef struct bpc_params {
    int32_t sr;            /* 16000 */
    int32_t n_fft;         /* 512   */
    int32_t hop;           /* 256   */
    int32_t n_mels;        /* 128   */
    int32_t n_mfcc;        /* 40    */
    float   fmax;          /* 4500  */
    int32_t n_gammatone;   /* 64    */
    int32_t n_lpc;         /* 12    */
    int32_t expected_len;  /* int(SR * DURATION) = 16000; 16000 * d with d = 2^a 3^b 5^c <= 32 selects the long mode
                              (BASELINE config 4): same outputs at T = expected_len / 256 + 1 frames */
    int32_t pad_scalars_to;/* 0 = emit the reference's 36 scalars; 39 = append zeros (README's "39") */
} bpc_params;

typedef struct bpc_handle bpc_handle;

int  bpc_abi_version(void);
void bpc_default_params(bpc_params* p);
/* frames per segment: expected_len / hop + 1 (process.py:30) */
int  bpc_num_frames(const bpc_params* p);
/* number of scalars written per segment (36 or pad_scalars_to) */
int  bpc_num_scalars(const bpc_params* p);

/* max_batch: largest B a single call may pass; workspaces are sized for an internal chunk, not for max_batch. */
int  bpc_create(bpc_handle** out, const bpc_params* p, int device, int64_t max_batch);
void bpc_destroy(bpc_handle* h);
const char* bpc_last_error(const bpc_handle* h);   /* h may be NULL: last error of a failed bpc_create */

/* Full path, device buffers.  wav: [B, L_in] (float32 or int16), feats: [B, 9, 128, T] float32,
 * scalars: [B, bpc_num_scalars] float32, status: [B] int32 (may be NULL).  L_in != expected_len is truncated / zero
 * padded on device (methods.py:24-28). */
int  bpc_precompute(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in,
                    float* feats, float* scalars, int32_t* status, void* stream);

/* Same, HOST buffers (pageable or pinned): chunks are staged through pinned memory, H2D / compute / D2H overlapped on
 * internal streams; returns after the results are in the host buffers.  This is the call the Python mirror of
 * process_dataset_threaded makes (core.py:19-45). */
int  bpc_precompute_host(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in,
                         float* feats, float* scalars, int32_t* status);

/* Compact host layout (what actually crosses PCIe).  pad_freq (methods.py:39-46) fills rows live..127 of a plane with
 * ONE constant, so of the 9 * 128 = 1152 rows of a segment only BPC_LIVE_ROWS = 772 carry data:
 *   chroma 24 (process.py:54-57), gammatone 64, lpc 12, mel / mel_delta / mel_delta2 128 each, mfcc 120, mod_spec 40,
 *   tempogram 128 -- bpc_live_rows(channel).
 * rows: [B, 772, T] float32, the data rows of the nine planes back to back in channel order; pad: [B, 9] float32, the
 * constant of every plane's remaining rows (0 for planes without pad rows).  67 % of the bytes of the full tensor and
 * no host-side fill: ONE contiguous device->host copy per internal piece.  bpc_expand_compact rebuilds the full
 * [n, 9, 128, T] tensor on the host (n_threads host threads; bit-identical to bpc_precompute_host's output) for
 * consumers that want it (the .npz writer); the packed shard and its Dataset keep the compact form. */
#define BPC_LIVE_ROWS 772
int  bpc_live_rows(int channel);
int  bpc_precompute_host_compact(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in,
                                 float* rows, float* pad, float* scalars, int32_t* status);
int  bpc_expand_compact(const float* rows, const float* pad, int64_t n, int T, float* feats, int n_threads);

/* Streaming form of bpc_precompute_host_compact, for a caller that works through a dataset batch after batch (the
 * loop of process_dataset_threaded, core.py:33-43).  The pieces of a call join a ring that belongs to the handle:
 * _begin enqueues them behind whatever is still in flight, returns with (at most) the last two pieces of this call
 * unretired and hands back a ticket; bpc_host_wait(ticket) returns once every output of that call is in its host
 * buffers (ticket < 0: everything enqueued so far).  With two sets of output buffers -- begin(k + 1), then wait(k) --
 * the GPU and both copy engines stay busy across calls: the head of a call (nothing to copy until its first piece has
 * been computed) and its tail (the D2H of the last piece) overlap with the neighbouring calls.  Until its ticket has
 * been waited for, a call's input and output buffers belong to the library.  The synchronous entry points are
 * exactly begin + wait, and they may be mixed with this pair (a synchronous call retires older tickets first). */
int  bpc_precompute_host_compact_begin(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in,
                                       float* rows, float* pad, float* scalars, int32_t* status, int64_t* ticket);
int  bpc_host_wait(bpc_handle* h, int64_t ticket);

/* Pinned host memory placed on the NUMA node the handle's GPU hangs off (sysfs numa_node of its PCI address; mbind +
 * first touch from a thread bound to that node's CPUs, then cudaHostRegister).  Buffers passed to bpc_precompute_host*
 * need not come from here, but on multi-socket boxes device<->host copies into remote-node memory run at a fraction
 * of the PCIe rate.  *numa_node receives the node used (-1: unknown / single node).  Free with bpc_host_free. */
void* bpc_host_alloc(bpc_handle* h, int64_t bytes, int* numa_node);
void  bpc_host_free(bpc_handle* h, void* p);

/* BASELINE config 2 stage: log-power STFT (power_to_db(|X|^2, ref=max), [B, 1+n_fft/2, T], may be NULL) and the
 * normalised mel / mel_delta / mel_delta2 planes ([B, 3, 128, T]).  Device buffers. */
int  bpc_stage_logmel(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in,
                      float* stft_db, float* mel3, void* stream);

/* methods.py:142-143 extract_spectral_modulation_features on caller-provided mel_db: [n, 128, T] -> [n, 40, T]
 * (ortho DCT-II over the mel axis, first 40 rows, then ortho DCT-II over time).  Device buffers. */
int  bpc_modspec(bpc_handle* h, const float* mel_db, int64_t n, float* out, void* stream);

/* Dataset-level statistics accumulated over every bpc_precompute* call since the last reset:
 * stats[(9 + nscal)][5] doubles = {count, sum, sum of squares, min, max}, first the 9 channels (sorted order, over all
 * 128*T values of each plane), then each scalar.  Host output.  bpc_channel_stats_device returns the device pointer
 * of the same accumulator so a caller can reduce it across ranks (one all-gather of the 1.8 KB, then sum over cols 0-2,
 * min col 3, max col 4: bpc_b200/stats.py). */
int  bpc_channel_stats(bpc_handle* h, double* stats_host);
int  bpc_channel_stats_device(bpc_handle* h, double** stats_dev, int64_t* n_rows);
int  bpc_channel_stats_reset(bpc_handle* h);

/* Raw (un-normalised) intermediates of the LAST chunk processed, for parity tests.  `what`:
 *   "mag512" [n,T,260] |STFT512| (k fastest, 257 valid)      "mel_db" [n,128,T]         "mfcc_raw" [n,120,T]
 *   "gammatone_raw" [n,64,T]   "mod_spec_raw" [n,40,T]        "chroma_stft_raw" [n,12,T] "chroma_cens_raw" [n,12,T]
 *   "lpc_raw" [n,12,F]         "onset_env" [n,T]              "tuning" [n,2] int32 (bin index 0..99 for 12 / 36 bpo)
 *   "ints" [n,2] int32 (n_peaks, autocorr first-min index)
 * Copies min(cap_bytes, available) bytes to the HOST buffer `out` and stores the byte count in *got. */
int  bpc_debug_copy(bpc_handle* h, const char* what, void* out, int64_t cap_bytes, int64_t* got);

/* Debug stage buffers cost ~100 KB of extra writes per segment, so they are off by default. */
int  bpc_set_debug(bpc_handle* h, int on);

/* Host-only access to the constant tables (no GPU needed; used by the CPU test-suite to pin them against the oracle):
 *   "mel_a" [128,257] "mel_b" [128,257] "mel_c" [64,257] "mel_d" [128,1025] "dct_mel" [40,128] "dct_time" [T,T]
 *   "hann512" [512] f64 "hann2048" [2048] f64 "chroma" [12,257] (tuning_idx) "cqt_basis" [36,257,2] (tuning_idx)
 *   "cqt_sqrt_len" [252] f64 (tuning_idx) "halfband" [ntaps] f64 "hist_edges" [101] f64
 *   "mel_d_band" [128,82]: the band form of "mel_d" the STFT-2048 kernel walks (start, count, 80 weights per row; starts
 *   moved down and padded with leading zero weights so that 32 consecutive rows start in 32 different shared-memory banks)
 * Writes float32 unless noted; returns the element count or a negative status. */
int64_t bpc_table_copy(const bpc_params* p, const char* name, int tuning_idx, void* out, int64_t cap_elems);

/* ---- batch assembly on a device-resident feature store -------------------------------------
 * dataset.py:59-73 collate_fn (stack features / scalars of the drawn items) fused with augmentation.py:5-44 /
 * train.py:76-89.  store_feats [N, 9, 128, T], store_scalars [N, S] (S = bpc_num_scalars), idx_a / idx_b [n] int64
 * device arrays of store rows; out_feats [n, 9, 128, T], out_scalars [n, S] (may be NULL).
 *   BPC_MIX_NONE    out[i] = store[idx_a[i]]
 *   BPC_MIX_MIXUP   out[i] = lam * store[idx_a[i]] + (1 - lam) * store[idx_b[i]], features and scalars (float32 multiply,
 *                   multiply, add -- the rounding of the torch expression in train.py:84-85)
 *   BPC_MIX_CUTMIX  rows [y1, y2) x columns [x1, x2) of every plane come from store[idx_b[i]] (augmentation.py:27-28);
 *                   scalars are those of idx_a
 * The random draws (permutation, lam, box) stay with the caller, as in the reference. */
enum bpc_mix_mode { BPC_MIX_NONE = 0, BPC_MIX_MIXUP = 1, BPC_MIX_CUTMIX = 2 };
int  bpc_collate(bpc_handle* h, const float* store_feats, const float* store_scalars, int64_t N,
                 const int64_t* idx_a, const int64_t* idx_b, int64_t n, int mode, double lam,
                 int y1, int y2, int x1, int x2, float* out_feats, float* out_scalars, void* stream);

/* ---- host-side output writer (no GPU involved) ---------------------------------------------------
 * process.py:92-103: np.savez(<target_dir>/<file_id>.npz, mel=..., mfcc=..., chroma=..., mel_delta=..., mel_delta2=...,
 * gammatone=..., lpc=..., mod_spec=..., tempogram=..., scalars=...), an uncompressed zip of ten .npy members.
 * feats is the [9, 128, T] sorted-key slab of one segment, scalars its [nscal] vector.
 * bpc_npz_size: bytes of one such archive.  bpc_npz_pack: serialise into `out` (returns bytes written).
 * bpc_npz_write_batch: write n archives with n_threads host threads; ok[i] = 1 written, -1 skipped because
 * status[i] has BPC_SEG_NONFINITE (status may be NULL), -2 cannot open, -3 short write. */
int64_t bpc_npz_size(int T, int nscal);
int64_t bpc_npz_pack(const float* feats, const float* scalars, int T, int nscal, void* out, int64_t cap);
int  bpc_npz_write_batch(const char* target_dir, const char* const* file_ids, const float* feats, const float* scalars,
                         const int32_t* status, int64_t n, int T, int nscal, int n_threads, int32_t* ok);

/* ---- host-side ingest (no GPU involved) ----------------------------------------------------------
 * process.py:28-29: `y, _ = librosa.load(wav_path, sr=SR); y = pad_or_truncate(y, EXPECTED_LEN)` for a batch of files.
 * Reads RIFF/WAVE PCM16 mono files of sample rate expected_sr with n_threads host threads into out[n, L] (int16, zero
 * padded / truncated to L; feed it to bpc_precompute* as BPC_WAV_PCM16, the device applies soundfile's 1/32768).
 * Per file: sr[i], frames[i] (frames in the file) and code[i] = 0 or a bpc_wav_code; rows of failed files are zeroed.
 * BPC_WAV_ERR_UNSUPPORTED (stereo, other rates or sample formats) means "decode this one with a general reader". */
enum bpc_wav_code { BPC_WAV_ERR_OPEN = -10, BPC_WAV_ERR_FORMAT = -11, BPC_WAV_ERR_UNSUPPORTED = -12 };
int  bpc_wav_load_batch(const char* const* paths, int64_t n, int expected_sr, int64_t L, int16_t* out,
                        int32_t* sr, int32_t* frames, int32_t* code, int n_threads);

/* ---- GPU-side decode ---------------------------------------------------------------------------------
 * process.py:28 `librosa.load(wav_path, sr=SR)` = soundfile decode to float32 (int16 / 2^15, 24- and 32-bit / 2^31,
 * unsigned 8-bit (x - 128) / 2^7, IEEE float as stored) + librosa.to_mono (float32 mean over the channels) and
 * process.py:29 pad_or_truncate -- for any RIFF/WAVE sample format and channel count, on the device.  The host only
 * walks the chunk headers (bpc_wav_parse: no sample is touched) and copies the file images to the GPU as they are.
 * bpc_wav_parse: image of one file -> where its samples lie and how they are stored; returns BPC_OK or a bpc_wav_code
 *   (BPC_WAV_ERR_UNSUPPORTED: compressed formats, other bit depths, more than 7 channels).
 * bpc_wav_decode: blob = n file images in DEVICE memory, file i starting at byte file_offset[i] (host array);
 *   info[i] from bpc_wav_parse (host array); a payload that does not lie inside blob[0, blob_bytes) is rejected
 *   (BPC_ERR_ARG) before anything is launched.  Writes y[n, L] float32 (device): frames beyond the file are zero, frames
 *   beyond L are dropped.  Files of another sample rate are decoded at their own rate (resample with bpc_resample). */
enum bpc_sample_fmt { BPC_FMT_U8 = 1, BPC_FMT_PCM16 = 2, BPC_FMT_PCM24 = 3, BPC_FMT_PCM32 = 4, BPC_FMT_F32 = 5, BPC_FMT_F64 = 6 };
typedef struct bpc_wav_info {
    int64_t data_offset;   /* byte offset of the first sample inside the file image */
    int64_t frames;        /* sample frames present (a truncated data chunk counts what is there) */
    int32_t sr, channels, fmt, reserved;
} bpc_wav_info;
int  bpc_wav_parse(const void* image, int64_t n_bytes, bpc_wav_info* info);
int  bpc_wav_decode(bpc_handle* h, const void* blob, int64_t blob_bytes, const int64_t* file_offset,
                    const bpc_wav_info* info, int64_t n, int64_t L, float* y, void* stream);

/* ---- sample-rate conversion on load ---------------------------------------------------------------
 * process.py:28 `librosa.load(wav_path, sr=SR)` resamples files of any other rate with libsoxr "HQ", an un-vendored
 * dependency that cannot be restated bit for bit.  Stand-in (the same disclosure as the CQT's half-band decimator):
 * a 150 dB Kaiser-windowed sinc with soxr HQ's pass band (flat to 0.9125 of the lower Nyquist), evaluated as a
 * polyphase filter in FP64 on the device; oracle/resample.py is its CPU counterpart.
 * bpc_resample_len: ceil(n_in * sr_out / sr_in) (librosa.resample's output length).
 * bpc_resample: device buffers, in [n_in] float32 -> out [bpc_resample_len] float32 (out_cap elements available).
 * bpc_resample_filter (host only, no GPU): the polyphase table [p, 2 half] as doubles; returns the element count. */
int64_t bpc_resample_len(int64_t n_in, int sr_in, int sr_out);
int  bpc_resample(bpc_handle* h, const float* in, int64_t n_in, int sr_in, int sr_out, float* out, int64_t out_cap,
                  void* stream);
int64_t bpc_resample_filter(int sr_in, int sr_out, double* out, int64_t cap, int* p, int* q, int* half);

/* Segments processed per internal chunk (= per kernel launch; default min(4144, max_batch) for 1 s segments); env
 * BPC_CHUNK overrides the default at create time. */
int  bpc_chunk_size(const bpc_handle* h);

/* Kernel launches issued by this handle since creation (bench.py's `gpu_launches`). */
int64_t bpc_launch_count(const bpc_handle* h);

/* Per-kernel device times (CUDA events around every launch of the full path) for bench.py's roofline leg.
 * ids: 0 ingest, 1 stft512, 2 spec512 consumers (both launches), 3 frame2048, 4 even2048, 5 cens, 6 time_basic+autocorr,
 * 7 hilbert, 8 lpc, 9 stats, 10 seg2048, 11 cens_dec, 12 cens_lo.  bpc_kernel_times synchronises, sums the elapsed ms / launch counts
 * since the last call. */
#define BPC_NUM_KERNEL_IDS 13
int  bpc_set_kernel_timing(bpc_handle* h, int on);
int  bpc_kernel_times(bpc_handle* h, double* ms_out, int64_t* launches_out, int n_ids);
const char* bpc_kernel_name(int id);

#ifdef __cplusplus
}
#endif
#endif /* BPC_B200_H */
