// Host-side output writer of the precompute path: the reference's per-segment `.npz` (process.py:92-103, np.savez:
// an uncompressed zip of ten "<key>.npy" members) serialised straight from the [9,128,T] / [S] float32 slabs the
// device path produces, by a small thread pool.  No GPU involved; np.load / zipfile read the result (CRC-32 checked).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bpc.h"

namespace {

// np.savez member order of process.py:93-103; value = channel index in the sorted-key feats slab (bpc_channel)
struct Member { const char* key; int channel; };
const Member kMembers[9] = {
    {"mel", BPC_CH_MEL}, {"mfcc", BPC_CH_MFCC}, {"chroma", BPC_CH_CHROMA}, {"mel_delta", BPC_CH_MEL_DELTA},
    {"mel_delta2", BPC_CH_MEL_DELTA2}, {"gammatone", BPC_CH_GAMMATONE}, {"lpc", BPC_CH_LPC},
    {"mod_spec", BPC_CH_MOD_SPEC}, {"tempogram", BPC_CH_TEMPOGRAM}};

uint32_t g_crc[8][256];
std::atomic<bool> g_crc_ready{false};

void crc_init() {
    if (g_crc_ready.load(std::memory_order_acquire)) return;
    static std::atomic<bool> busy{false};
    bool expected = false;
    if (!busy.compare_exchange_strong(expected, true)) {
        while (!g_crc_ready.load(std::memory_order_acquire)) std::this_thread::yield();
        return;
    }
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crc[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int s = 1; s < 8; ++s) g_crc[s][i] = (g_crc[s - 1][i] >> 8) ^ g_crc[0][g_crc[s - 1][i] & 0xFF];
    g_crc_ready.store(true, std::memory_order_release);
}

// slice-by-8 CRC-32 (IEEE 802.3, the zip polynomial)
uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n) {
    crc = ~crc;
    while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { crc = g_crc[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8); --n; }
    while (n >= 8) {
        uint64_t w;
        std::memcpy(&w, p, 8);
        w ^= crc;
        crc = g_crc[7][w & 0xFF] ^ g_crc[6][(w >> 8) & 0xFF] ^ g_crc[5][(w >> 16) & 0xFF] ^ g_crc[4][(w >> 24) & 0xFF] ^
              g_crc[3][(w >> 32) & 0xFF] ^ g_crc[2][(w >> 40) & 0xFF] ^ g_crc[1][(w >> 48) & 0xFF] ^ g_crc[0][w >> 56];
        p += 8; n -= 8;
    }
    while (n--) crc = g_crc[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

// .npy v1.0 header of a C-ordered little-endian float32 array (numpy.lib.format: magic, version, uint16 length,
// dict literal padded with spaces so that the data starts on a 64-byte boundary, terminated by '\n')
std::string npy_header(const std::string& shape) {
    std::string d = "{'descr': '<f4', 'fortran_order': False, 'shape': (" + shape + "), }";
    const size_t unpadded = 10 + d.size() + 1;
    d.append((64 - unpadded % 64) % 64, ' ');
    d.push_back('\n');
    std::string h("\x93NUMPY\x01\x00", 8);
    h.push_back((char)(d.size() & 0xFF));
    h.push_back((char)(d.size() >> 8));
    return h + d;
}

void put16(uint8_t*& p, uint32_t v) { p[0] = v & 0xFF; p[1] = (v >> 8) & 0xFF; p += 2; }
void put32(uint8_t*& p, uint32_t v) { put16(p, v & 0xFFFF); put16(p, v >> 16); }

struct Layout {
    std::string hdr_plane, hdr_scal;
    size_t plane_bytes, scal_bytes;
    int64_t total;
};

Layout layout_of(int T, int nscal) {
    Layout l;
    l.hdr_plane = npy_header("128, " + std::to_string(T));
    l.hdr_scal = npy_header(std::to_string(nscal) + ",");
    l.plane_bytes = (size_t)BPC_PLANE_ROWS * T * 4;
    l.scal_bytes = (size_t)nscal * 4;
    int64_t tot = 0;
    for (const Member& m : kMembers) {
        const size_t name = std::strlen(m.key) + 4;
        tot += 30 + name + l.hdr_plane.size() + l.plane_bytes;     // local header + member
        tot += 46 + name;                                          // central directory entry
    }
    tot += 30 + 11 + l.hdr_scal.size() + l.scal_bytes + 46 + 11;   // "scalars.npy"
    tot += 22;                                                     // end of central directory
    l.total = tot;
    return l;
}

int64_t pack(const Layout& l, const float* feats, const float* scalars, int T, uint8_t* out) {
    struct Entry { std::string name; uint32_t crc, size, offset; };
    Entry ent[10];
    uint8_t* p = out;
    auto member = [&](int i, const std::string& name, const std::string& hdr, const void* data, size_t bytes) {
        Entry& e = ent[i];
        e.name = name;
        e.offset = (uint32_t)(p - out);
        e.size = (uint32_t)(hdr.size() + bytes);
        uint32_t c = crc32(0, reinterpret_cast<const uint8_t*>(hdr.data()), hdr.size());
        e.crc = crc32(c, static_cast<const uint8_t*>(data), bytes);
        put32(p, 0x04034b50); put16(p, 20); put16(p, 0); put16(p, 0);         // signature, version, flags, stored
        put16(p, 0); put16(p, 0x21);                                          // dos time 00:00:00, date 1980-01-01
        put32(p, e.crc); put32(p, e.size); put32(p, e.size);
        put16(p, (uint32_t)name.size()); put16(p, 0);
        std::memcpy(p, name.data(), name.size()); p += name.size();
        std::memcpy(p, hdr.data(), hdr.size()); p += hdr.size();
        std::memcpy(p, data, bytes); p += bytes;
    };
    for (int i = 0; i < 9; ++i)
        member(i, std::string(kMembers[i].key) + ".npy", l.hdr_plane,
               feats + (size_t)kMembers[i].channel * BPC_PLANE_ROWS * T, l.plane_bytes);
    member(9, "scalars.npy", l.hdr_scal, scalars, l.scal_bytes);
    const uint32_t cd_off = (uint32_t)(p - out);
    for (const Entry& e : ent) {
        put32(p, 0x02014b50); put16(p, 20); put16(p, 20); put16(p, 0); put16(p, 0);
        put16(p, 0); put16(p, 0x21);
        put32(p, e.crc); put32(p, e.size); put32(p, e.size);
        put16(p, (uint32_t)e.name.size()); put16(p, 0); put16(p, 0); put16(p, 0); put16(p, 0);
        put32(p, 0x01800000u);                                                // external attrs: regular file 0600
        put32(p, e.offset);
        std::memcpy(p, e.name.data(), e.name.size()); p += e.name.size();
    }
    const uint32_t cd_size = (uint32_t)(p - out) - cd_off;
    put32(p, 0x06054b50); put16(p, 0); put16(p, 0); put16(p, 10); put16(p, 10);
    put32(p, cd_size); put32(p, cd_off); put16(p, 0);
    return p - out;
}

}  // namespace

extern "C" {

int64_t bpc_npz_size(int T, int nscal) {
    if (T <= 0 || nscal <= 0) return BPC_ERR_ARG;
    return layout_of(T, nscal).total;
}

int64_t bpc_npz_pack(const float* feats, const float* scalars, int T, int nscal, void* out, int64_t cap) {
    if (!feats || !scalars || !out || T <= 0 || nscal <= 0) return BPC_ERR_ARG;
    crc_init();
    const Layout l = layout_of(T, nscal);
    if (cap < l.total) return BPC_ERR_ARG;
    return pack(l, feats, scalars, T, static_cast<uint8_t*>(out));
}

int bpc_npz_write_batch(const char* target_dir, const char* const* file_ids, const float* feats, const float* scalars,
                        const int32_t* status, int64_t n, int T, int nscal, int n_threads, int32_t* ok) {
    if (!target_dir || !file_ids || !feats || !scalars || !ok || n < 0 || T <= 0 || nscal <= 0) return BPC_ERR_ARG;
    crc_init();
    const Layout l = layout_of(T, nscal);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    std::atomic<int64_t> next{0};
    const std::string dir(target_dir);
    auto work = [&]() {
        std::vector<uint8_t> buf((size_t)l.total);
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            ok[i] = 0;
            if (status && (status[i] & BPC_SEG_NONFINITE)) { ok[i] = -1; continue; }   // reported by the caller as a failure
            const int64_t bytes = pack(l, feats + (size_t)i * 9 * BPC_PLANE_ROWS * T, scalars + (size_t)i * nscal, T,
                                       buf.data());
            const std::string path = dir + "/" + file_ids[i] + ".npz";
            FILE* f = std::fopen(path.c_str(), "wb");
            if (!f) { ok[i] = -2; continue; }
            const bool good = std::fwrite(buf.data(), 1, (size_t)bytes, f) == (size_t)bytes;
            if (std::fclose(f) != 0 || !good) { ok[i] = -3; continue; }
            ok[i] = 1;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return BPC_OK;
}

}  // extern "C"
