// Host-side ingest of the precompute path: `librosa.load(path, sr=16000)` (process.py:28) for the files the reference
// is fed -- RIFF/WAVE, PCM16, mono -- read by a thread pool straight into one [n, L] int16 batch (pad_or_truncate,
// methods.py:24-28, is applied while copying; the 1/32768 scaling happens on the device in k_ingest).  Anything else
// (other sample rates, stereo, float / 24-bit data) is reported per file so that the caller can fall back to a general
// decoder for that file.  No GPU involved.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/bpc.h"

namespace {

uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// returns BPC_WAV_* code; on success fills out[0..L) and *frames (frames in the file, before pad / truncate)
int load_one(const char* path, int expected_sr, int64_t L, int16_t* out, int32_t* sr, int32_t* frames) {
    *sr = 0; *frames = 0;
    FILE* f = std::fopen(path, "rb");
    if (!f) return BPC_WAV_ERR_OPEN;
    unsigned char hdr[12];
    int rc = BPC_WAV_ERR_FORMAT;
    bool have_fmt = false;
    uint16_t tag = 0, channels = 0, bits = 0;
    if (std::fread(hdr, 1, 12, f) == 12 && !std::memcmp(hdr, "RIFF", 4) && !std::memcmp(hdr + 8, "WAVE", 4)) {
        for (;;) {
            unsigned char ch[8];
            if (std::fread(ch, 1, 8, f) != 8) break;
            const uint32_t size = rd32(ch + 4);
            if (!std::memcmp(ch, "fmt ", 4)) {
                unsigned char fm[40];
                const size_t want = size < sizeof(fm) ? size : sizeof(fm);
                if (size < 16 || std::fread(fm, 1, want, f) != want) break;
                tag = rd16(fm); channels = rd16(fm + 2); *sr = (int32_t)rd32(fm + 4); bits = rd16(fm + 14);
                if (tag == 0xFFFE && size >= 26) tag = rd16(fm + 24);            // WAVE_FORMAT_EXTENSIBLE: sub-format
                have_fmt = true;
                if (std::fseek(f, (long)(size - want + (size & 1)), SEEK_CUR)) break;
            } else if (!std::memcmp(ch, "data", 4)) {
                if (!have_fmt) break;
                if (tag != 1 || bits != 16 || channels != 1 || *sr != expected_sr) { rc = BPC_WAV_ERR_UNSUPPORTED; break; }
                const int64_t n = size / 2;
                *frames = (int32_t)(n > 0x7fffffff ? 0x7fffffff : n);
                const int64_t take = n < L ? n : L;
                const size_t got = std::fread(out, 2, (size_t)take, f);          // little-endian host (x86-64 / aarch64)
                if ((int64_t)got < take) {                                        // truncated file: soundfile reads what is there
                    *frames = (int32_t)got;
                }
                if ((int64_t)got < L) std::memset(out + got, 0, (size_t)(L - (int64_t)got) * 2);
                rc = BPC_OK;
                break;
            } else {
                if (std::fseek(f, (long)(size + (size & 1)), SEEK_CUR)) break;
            }
        }
    }
    std::fclose(f);
    return rc;
}

}  // namespace

// Chunk walk over a file image in memory (the GPU-side decode path: the samples themselves are only touched by
// k_wav_decode).  Same acceptance rules as soundfile / scipy.io.wavfile for uncompressed data.
extern "C" int bpc_wav_parse(const void* image, int64_t n_bytes, bpc_wav_info* info) {
    if (!image || !info || n_bytes < 0) return BPC_ERR_ARG;
    std::memset(info, 0, sizeof(*info));
    const unsigned char* p = static_cast<const unsigned char*>(image);
    if (n_bytes < 12 || std::memcmp(p, "RIFF", 4) || std::memcmp(p + 8, "WAVE", 4)) return BPC_WAV_ERR_FORMAT;
    int64_t pos = 12;
    bool have_fmt = false;
    uint16_t tag = 0, channels = 0, bits = 0;
    while (pos + 8 <= n_bytes) {
        const uint32_t size = rd32(p + pos + 4);
        const unsigned char* body = p + pos + 8;
        if (!std::memcmp(p + pos, "fmt ", 4)) {
            if (size < 16 || pos + 8 + 16 > n_bytes) return BPC_WAV_ERR_FORMAT;
            tag = rd16(body); channels = rd16(body + 2); info->sr = (int32_t)rd32(body + 4); bits = rd16(body + 14);
            if (tag == 0xFFFE && size >= 26 && pos + 8 + 26 <= n_bytes) tag = rd16(body + 24);   // WAVE_FORMAT_EXTENSIBLE
            have_fmt = true;
        } else if (!std::memcmp(p + pos, "data", 4)) {
            if (!have_fmt) return BPC_WAV_ERR_FORMAT;
            int fmt = 0;
            if (tag == 1) fmt = bits == 8 ? BPC_FMT_U8 : bits == 16 ? BPC_FMT_PCM16 : bits == 24 ? BPC_FMT_PCM24 : bits == 32 ? BPC_FMT_PCM32 : 0;
            else if (tag == 3) fmt = bits == 32 ? BPC_FMT_F32 : bits == 64 ? BPC_FMT_F64 : 0;
            if (!fmt || channels < 1 || channels > 7 || info->sr <= 0) return BPC_WAV_ERR_UNSUPPORTED;
            const int64_t avail = n_bytes - (pos + 8);
            const int64_t bytes = (int64_t)size < avail ? (int64_t)size : avail;          // truncated file: what is there
            info->data_offset = pos + 8;
            info->channels = channels;
            info->fmt = fmt;
            info->frames = bytes / ((int64_t)channels * (bits / 8));
            return BPC_OK;
        }
        pos += 8 + (int64_t)size + (size & 1);
    }
    return BPC_WAV_ERR_FORMAT;
}

extern "C" int bpc_wav_load_batch(const char* const* paths, int64_t n, int expected_sr, int64_t L, int16_t* out,
                                  int32_t* sr, int32_t* frames, int32_t* code, int n_threads) {
    if (!paths || !out || !sr || !frames || !code || n < 0 || L <= 0) return BPC_ERR_ARG;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            code[i] = load_one(paths[i], expected_sr, L, out + (size_t)i * L, sr + i, frames + i);
            if (code[i] != BPC_OK) std::memset(out + (size_t)i * L, 0, (size_t)L * 2);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return BPC_OK;
}
