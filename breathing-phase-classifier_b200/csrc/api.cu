// C ABI of libbpc_b200.so (include/bpc.h): handle, constant tables, workspaces, chunked launch sequence, the
// host-buffer (pinned-staged, double-buffered) path, dataset statistics and debug accessors.
#include <cuda_runtime.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include "../../include/bpc.h"
#include "kernels.cuh"
#include "tables.hpp"
#include "fft20.cuh"
#include "tuning.cuh"

using namespace bpc;

namespace {

thread_local std::string g_create_error;
const double kPi = 3.141592653589793238462643383279502884;

constexpr int kSlots = 3;          // chunks in flight in the host-buffer path: one computing, one in D2H, one in host fill

struct Slot {                      // one in-flight chunk of the host-buffer path
    void* d_wav = nullptr;         // [chunk, L_in_max] raw input (f32 or pcm16)
    float* d_feats = nullptr;
    float* d_scalars = nullptr;
    int32_t* d_status = nullptr;
    void* h_wav = nullptr;         // pinned staging
    float* h_feats = nullptr;
    float* h_scalars = nullptr;
    int32_t* h_status = nullptr;
    cudaStream_t st = nullptr;
    cudaStream_t st_out = nullptr;    // highest-priority stream of the piece's output side: pad values, row compaction, D2H.
                                      // On the compute stream those two small kernels queued behind the NEXT piece's
                                      // kernels and every D2H started ~0.2 ms late (r02 timeline: 45 instead of 52 GB/s).
    cudaEvent_t done = nullptr;
    cudaEvent_t computed = nullptr;   // kernels of this chunk finished (the slots share one workspace)
    cudaEvent_t fill_ready = nullptr; // the pad values of this chunk are in h_fill
    float* d_fill = nullptr;          // [chunk, 9] pad value of every plane (host path: pad rows are not transferred)
    float* h_fill = nullptr;
    float* d_rows = nullptr;          // [chunk, 772, T] the data rows back to back (k_compact_rows): one contiguous D2H
};

struct HostOut {                   // exactly one of the two host layouts
    float* feats;                  // full  [B, 9, 128, T]
    float* rows;                   // compact [B, 772, T] ...
    float* pad;                    // ... + [B, 9]
};

struct Fly {                       // one piece of the host path that has been enqueued and not yet retired
    int slot;
    int64_t off;                   // first segment of the piece inside its call
    int n;
    HostOut out;
    float* scalars;
    int32_t* status;
    bool to_rows, live_only, pin_f, pin_p, pin_s;
    bool filled;                   // full layout: the pad rows have been written by the host threads
    bool last;                     // last piece of its call
    int64_t ticket;
};

// Rows of each plane that carry data; rows live..127 are one constant per plane (pad_freq, methods.py:39-46):
// chroma 12 + 12 (process.py:54-57), gammatone N_GAMMATONE, lpc N_LPC, mel x3 full, mfcc 3 * N_MFCC, mod_spec N_MFCC,
// tempogram full (truncated from 384 rows, process.py:78).
const int kLiveRows[9] = {24, 64, 12, 128, 128, 128, 120, 40, 128};
constexpr int kLiveTotal = BPC_LIVE_ROWS;

struct RowRun { int start, end; };     // [start, end) in the flattened 9 * 128 rows of one segment

std::vector<RowRun> live_runs() {
    std::vector<RowRun> r;
    for (int c = 0; c < 9; ++c) {
        const int s = c * kPlaneRows, e = s + kLiveRows[c];
        if (!r.empty() && r.back().end == s) r.back().end = e;
        else r.push_back({s, e});
    }
    return r;
}

// Constant fill with streaming (non-temporal) stores: the pad rows are written once and not read by this process, and
// a write-allocate fill would first READ every cache line it overwrites, doubling the host-memory traffic that the
// ranks of a multi-GPU box share.
inline void fill_stream(float* p, float* e, float v) {
#if defined(__SSE2__)
    while (p < e && (reinterpret_cast<uintptr_t>(p) & 15)) *p++ = v;
    const __m128 x = _mm_set1_ps(v);
    for (; p + 4 <= e; p += 4) _mm_stream_ps(p, x);
#endif
    while (p < e) *p++ = v;
}


// ---- NUMA placement of the host side (multi-GPU boxes are multi-socket: a device->host copy into memory of the other
// socket crosses the inter-socket link and runs at a fraction of the PCIe rate; tools/d2h_ceiling.py measures it).
struct NumaInfo {
    int node = -1;                   // node of the GPU's PCIe root (-1: unknown / not NUMA)
    std::vector<int> cpus;           // that node's CPUs that this process may run on
};

std::string read_small_file(const std::string& path) {
    std::string out;
    if (FILE* f = std::fopen(path.c_str(), "r")) {
        char buf[4096];
        const size_t n = std::fread(buf, 1, sizeof(buf) - 1, f);
        buf[n] = 0;
        out = buf;
        std::fclose(f);
    }
    while (!out.empty() && (out.back() == '\n' || out.back() == ' ')) out.pop_back();
    return out;
}

std::vector<int> parse_cpulist(const std::string& s) {
    std::vector<int> cpus;
    size_t i = 0;
    while (i < s.size()) {
        char* end = nullptr;
        const long a = std::strtol(s.c_str() + i, &end, 10);
        if (end == s.c_str() + i) break;
        long b = a;
        i = (size_t)(end - s.c_str());
        if (i < s.size() && s[i] == '-') {
            b = std::strtol(s.c_str() + i + 1, &end, 10);
            i = (size_t)(end - s.c_str());
        }
        for (long c = a; c <= b && c < 4096; ++c) cpus.push_back((int)c);
        if (i < s.size() && s[i] == ',') ++i;
    }
    return cpus;
}

NumaInfo numa_of_device(int device) {
    NumaInfo info;
#if defined(__linux__)
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return info; }
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    const std::string base = std::string("/sys/bus/pci/devices/") + bus;
    const std::string node = read_small_file(base + "/numa_node");
    if (!node.empty()) info.node = std::atoi(node.c_str());
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    const bool have_mask = sched_getaffinity(0, sizeof(allowed), &allowed) == 0;
    for (int c : parse_cpulist(read_small_file(base + "/local_cpulist")))
        if (!have_mask || (c < CPU_SETSIZE && CPU_ISSET(c, &allowed))) info.cpus.push_back(c);
    // one node only (or the kernel does not know): nothing to place
    if (read_small_file("/sys/devices/system/node/online").find_first_of(",-") == std::string::npos) info.node = -1;
#endif
    return info;
}

void bind_this_thread(const std::vector<int>& cpus) {
#if defined(__linux__)
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) if (c < CPU_SETSIZE) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof(set), &set);
#endif
}

// Page-aligned host memory on `numa.node` (mbind; containers that filter the syscall still get first-touch placement
// because the touching thread is bound to the node's CPUs), registered with CUDA.  Returns nullptr on failure.
void* numa_pinned_alloc(const NumaInfo& numa, size_t bytes, bool place) {
#if defined(__linux__)
    const size_t len = (bytes + 4095) & ~size_t(4095);
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    madvise(p, len, MADV_HUGEPAGE);
    cpu_set_t old;
    const bool restore = place && !numa.cpus.empty() && sched_getaffinity(0, sizeof(old), &old) == 0;
    if (place && numa.node >= 0 && numa.node < 64) {
        unsigned long mask = 1ul << numa.node;
        syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, &mask, 65ul, 0u);
    }
    if (restore) bind_this_thread(numa.cpus);
    std::memset(p, 0, len);                                   // first touch
    if (restore) sched_setaffinity(0, sizeof(old), &old);
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, len);
        return nullptr;
    }
    return p;
#else
    void* p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
#endif
}

void numa_pinned_free(void* p, size_t bytes) {
#if defined(__linux__)
    cudaHostUnregister(p);
    munmap(p, (bytes + 4095) & ~size_t(4095));
#else
    (void)bytes;
    cudaFreeHost(p);
#endif
}

// Minimal fork-join pool for the host-side work of the host path (pad-row fill, staging copies).
class HostPool {
public:
    explicit HostPool(int n, std::vector<int> cpus = {}) : cpus_(std::move(cpus)) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { bind_this_thread(cpus_); loop(i); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return (int)th_.size() + 1; }
    // runs fn(part, parts) on every worker and on the caller; returns when all are done
    void run(const std::function<void(int, int)>& fn) {
        { std::lock_guard<std::mutex> l(m_); fn_ = &fn; pending_ = (int)th_.size(); ++gen_; }
        cv_.notify_all();
        fn((int)th_.size(), size());
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }
private:
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* fn;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(id, size());
            { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::vector<int> cpus_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* fn_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace

// Everything one in-flight chunk needs besides its inputs / outputs: the intermediate workspace, the side streams of
// the fork / join inside a chunk and their events.  The device entry points use the handle's `main` context; the host
// path gives each of its slots an own context so that consecutive pieces overlap on the GPU (the tail of piece i's
// kernels runs next to the head of piece i + 1 instead of leaving SMs idle).
struct ChunkCtx {
    Workspace ws{};
    Workspace ws_dbg{};            // same workspace with the debug pointers populated (main context only)
    int cap = 0;
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_spec512 = nullptr, ev_time = nullptr, ev_f2048 = nullptr, ev_seg = nullptr;
    cudaEvent_t ws_free = nullptr; // end of the last enqueue that used this workspace (any stream)
    bool ws_used = false;
};

struct bpc_handle {
    bpc_params p{};
    Geometry g{};
    int device = 0;
    int64_t max_batch = 0;
    int chunk = 0;
    Tables tb{};
    ChunkCtx main;                 // context of the device entry points (and of the host path while debug is on)
    ChunkCtx slot_ctx[kSlots];     // contexts of the host path's slots (allocated with the slots)
    bool debug = false;
    double* stats_acc = nullptr;   // [(9 + nscal), 5]
    std::vector<void*> dev_allocs;
    std::vector<void*> host_allocs;
    Slot slot[kSlots];
    bool slots_ready = false;
    size_t slot_wav_bytes = 0;
    int* live_dev = nullptr;       // kLiveRows on the device
    HostPool* pool = nullptr;      // host threads of the host path (env BPC_HOST_THREADS; default min(8, cores / ranks))
    NumaInfo numa{};               // NUMA node / local CPUs of this GPU (env BPC_NUMA=0: no placement)
    bool numa_place = true;
    std::vector<std::pair<void*, size_t>> numa_allocs;   // numa_pinned_alloc blocks owned by the handle or handed out
    bool contig_d2h = true;        // compact host layout: k_compact_rows + ONE copy per piece (env BPC_D2H_MODE=2d: row runs)
    struct Resampler { int sr_in, sr_out, p, q, half; const double* tab; };
    std::vector<Resampler> resamplers;   // polyphase tables uploaded so far (bpc_resample)
    WavItem* wav_items = nullptr;        // device copy of the per-file descriptors of bpc_wav_decode
    int64_t wav_items_cap = 0;
    bool compact_d2h = true;       // host path transfers live rows only (env BPC_COMPACT_D2H=0: whole planes)
    int host_chunk = 0;            // piece size of the host path (env BPC_HOST_CHUNK, default chunk / 2: the D2H of a piece
                                   // can only start when its kernels are done, so smaller pieces shorten the ramp)
    bool taper_tail = true;        // host path halves the last pieces of a call (env BPC_TAPER=0: equal chunks)
    bool ramp_head = true;         // ... and starts with small pieces so that the first D2H starts early (env BPC_RAMP=0)
    bool slot_ctx_on = false;      // env BPC_SLOT_CTX=1: every slot computes in its own workspace (pieces overlap on the GPU)
    // The host path is a ring of pieces that outlives a call: bpc_precompute_host*_begin enqueues the pieces of a call
    // behind whatever is still in flight and returns with up to two pieces unretired; bpc_host_wait retires them.
    std::deque<Fly> fly;
    int64_t piece_seq = 0;         // pieces enqueued so far (slot = piece_seq % kSlots)
    int64_t ticket_next = 1, ticket_done = 0;
    double t_wait = 0.0, t_fill = 0.0;
    int last_n = 0;
    int64_t launches0 = 0;
    bool timing = false;           // per-kernel CUDA-event timing (bench.py roofline leg)
    // Fork / join inside a chunk: the STFT-512 branch, the time-domain branch and the per-segment STFT-2048 statistics
    // run on side streams next to the STFT-2048 -> tuning -> CENS chain on the caller's stream (run_chunk).
    bool multi_stream = true;      // env BPC_STREAMS=0 turns it off; the per-kernel timing leg always runs serially
    struct Ev { int id; cudaEvent_t a, b; };
    std::vector<Ev> evs;
    std::string err;
};

namespace {

#define BPC_CUDA(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                        \
            return BPC_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

template <typename T>
int upload(bpc_handle* h, const std::vector<T>& v, const T** out) {
    void* d = nullptr;
    BPC_CUDA(h, cudaMalloc(&d, std::max<size_t>(16, v.size() * sizeof(T))));
    h->dev_allocs.push_back(d);
    BPC_CUDA(h, cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T*>(d);
    return BPC_OK;
}

template <typename T>
int dalloc(bpc_handle* h, size_t count, T** out) {
    void* d = nullptr;
    BPC_CUDA(h, cudaMalloc(&d, std::max<size_t>(16, count * sizeof(T))));
    h->dev_allocs.push_back(d);
    BPC_CUDA(h, cudaMemset(d, 0, std::max<size_t>(16, count * sizeof(T))));
    *out = reinterpret_cast<T*>(d);
    return BPC_OK;
}

int upload_bank(bpc_handle* h, const SparseBank& b, BankDev* out) {
    const int* s; const int* c; const float* w;
    int rc;
    if ((rc = upload(h, b.start, &s))) return rc;
    if ((rc = upload(h, b.count, &c))) return rc;
    if ((rc = upload(h, b.w, &w))) return rc;
    std::vector<float> t((size_t)b.rows * (b.width + 4), 0.f);      // four zero taps past the widest row: readers may unroll by 4
    for (int r = 0; r < b.rows; ++r)
        for (int j = 0; j < b.width; ++j) t[(size_t)j * b.rows + r] = b.w[(size_t)r * b.width + j];
    const float* wt;
    if ((rc = upload(h, t, &wt))) return rc;
    out->start = s; out->count = c; out->w = w; out->wt = wt; out->rows = b.rows; out->width = b.width;
    return BPC_OK;
}

// Band starts moved down so that the 32 rows a warp reads together start in 32 different shared-memory banks.
// k_frame2048 walks the band of mel row m = lane + 32 i as row[start[m] + j], j = 0, 1, ...: with the natural starts of
// the n_fft 2048 bank 2-4 lanes of a group fall into the same bank at every tap (316 instead of 92 wavefronts per frame,
// all 5.8e7 excess shared wavefronts of the kernel: profiles/r02_j_ncu_summary.txt).  Row m gets start - r and r leading
// ZERO weights, r in [0, 32) chosen per group of 32 rows by a bipartite matching (rows x banks) that minimises the
// longest padded band; fma(0, p, +0) = +0 for the finite p the row holds, so every sum is bit-identical.
SparseBank deconflict_bank(const SparseBank& b, int group) {
    SparseBank o = b;
    std::vector<int> shift(b.rows, 0);
    for (int g0 = 0; g0 < b.rows; g0 += group) {
        const int n = std::min(group, b.rows - g0);
        int cmax = 0;
        for (int i = 0; i < n; ++i) cmax = std::max(cmax, b.count[g0 + i]);
        for (int cap = cmax; cap <= cmax + 32; ++cap) {
            std::vector<int> owner(32, -1);                                  // bank -> row of the group
            auto ok = [&](int i, int bank) {
                const int r = ((b.start[g0 + i] - bank) % 32 + 32) % 32;
                return b.start[g0 + i] - r >= 0 && b.count[g0 + i] + r <= cap;
            };
            std::function<bool(int, std::vector<char>&)> augment = [&](int i, std::vector<char>& seen) {
                for (int bank = 0; bank < 32; ++bank) {
                    if (seen[bank] || !ok(i, bank)) continue;
                    seen[bank] = 1;
                    if (owner[bank] < 0 || augment(owner[bank], seen)) { owner[bank] = i; return true; }
                }
                return false;
            };
            int matched = 0;
            for (int i = 0; i < n; ++i) {
                std::vector<char> seen(32, 0);
                matched += augment(i, seen) ? 1 : 0;
            }
            if (matched == n) {
                for (int bank = 0; bank < 32; ++bank)
                    if (owner[bank] >= 0) shift[g0 + owner[bank]] = ((b.start[g0 + owner[bank]] - bank) % 32 + 32) % 32;
                break;
            }
        }
    }
    o.width = 0;
    for (int r = 0; r < b.rows; ++r) {
        o.start[r] = b.start[r] - shift[r];
        o.count[r] = b.count[r] + shift[r];
        o.width = std::max(o.width, o.count[r]);
    }
    o.w.assign((size_t)o.rows * std::max(1, o.width), 0.f);
    for (int r = 0; r < b.rows; ++r)
        for (int j = 0; j < b.count[r]; ++j) o.w[(size_t)r * o.width + shift[r] + j] = b.w[(size_t)r * b.width + j];
    return o;
}

std::vector<double2> twiddles(int n, int count) {
    std::vector<double2> t(count);
    for (int j = 0; j < count; ++j) {
        const double ang = -2.0 * kPi * double(j) / double(n);
        t[j] = make_double2(std::cos(ang), std::sin(ang));
    }
    return t;
}

int check_params(const bpc_params* p, std::string* why) {
    if (!p) { *why = "params is NULL"; return BPC_ERR_ARG; }
    if (p->sr != 16000 || p->n_fft != 512 || p->hop != 256 || p->n_mels != 128 || p->n_mfcc != 40 ||
        p->n_gammatone != 64 || p->n_lpc != 12) {
        *why = "this build implements the reference constants only (sr 16000, n_fft 512, hop 256, n_mels 128, "
               "n_mfcc 40, n_gammatone 64, n_lpc 12; process.py:12-23)";
        return BPC_ERR_UNSUPPORTED;
    }
    if (!(p->fmax > 0.f && p->fmax <= 8000.f)) { *why = "fmax out of range"; return BPC_ERR_ARG; }
    {
        // DURATION 1.0 s is the reference's constant (process.py:13-14) and the on-chip path.  Whole multiples up to 32 s
        // run the same kernels in "long mode" (BASELINE config 4) with their per-segment arrays in a global scratch
        // region; the Hilbert FFT of length 8000 d is mixed-radix 2/3/5, so d must factor into those.
        int d = p->expected_len / 16000, r = d;
        for (int f : {2, 3, 5}) while (r > 1 && r % f == 0) r /= f;
        if (p->expected_len <= 0 || p->expected_len % 16000 != 0 || d < 1 || d > 32 || r != 1) {
            *why = "expected_len must be 16000 * d with d = 2^a 3^b 5^c <= 32 (DURATION 1.0 s is the reference's value)";
            return BPC_ERR_UNSUPPORTED;
        }
    }
    if (p->pad_scalars_to != 0 && p->pad_scalars_to < BPC_NUM_SCALARS) { *why = "pad_scalars_to < 36"; return BPC_ERR_ARG; }
    if (p->pad_scalars_to > 256) { *why = "pad_scalars_to > 256"; return BPC_ERR_ARG; }
    return BPC_OK;
}

int build_tables(bpc_handle* h) {
    const bpc_params& p = h->p;
    Tables& tb = h->tb;
    int rc;
    if ((rc = upload(h, hann_periodic(512), &tb.hann512))) return rc;
    if ((rc = upload(h, hann_periodic(2048), &tb.hann2048))) return rc;
    {
        std::vector<double> hw = hann_periodic(2048);
        for (double& v : hw) v *= 0.5;
        if ((rc = upload(h, hw, &tb.hann2048h))) return rc;
        std::vector<double2> t(576);
        for (int k = 0; k < 576; ++k) {
            const double th = 2.0 * kPi * double(k) / 2048.0;
            const double wr = -std::sin(th), wi = -std::cos(th);          // -i exp(-i th)
            t[k] = k < 256 ? make_double2(wi, wr / wi) : make_double2(wr, wi / wr);
        }
        if ((rc = upload(h, t, &tb.rs2048))) return rc;
    }
    if ((rc = upload(h, hann_periodic(384), &tb.hann384))) return rc;
    if ((rc = upload(h, hamming_sym(400), &tb.hamming400))) return rc;
    if ((rc = upload(h, twiddles(256, 256), &tb.tw256))) return rc;
    {
        std::vector<double2> t(32 * 32);
        for (int k1 = 0; k1 < 32; ++k1)
            for (int hh = 0; hh < 32; ++hh) {
                const double ang = -2.0 * kPi * double(hh * k1) / 1024.0;
                t[k1 * 32 + hh] = make_double2(std::cos(ang), std::sin(ang));
            }
        if ((rc = upload(h, t, &tb.twa1024))) return rc;
    }
    {
        std::vector<double2> t(16 * 16);
        for (int k1 = 0; k1 < 16; ++k1)
            for (int hh = 0; hh < 16; ++hh) {
                const double ang = -2.0 * kPi * double((hh * k1) & 255) / 256.0;
                t[k1 * 16 + hh] = make_double2(std::cos(ang), std::sin(ang));
            }
        if ((rc = upload(h, t, &tb.twa256))) return rc;
    }
    {
        std::vector<double2> ta(16 * 64), tbb(16 * 4);
        for (int k1 = 0; k1 < 16; ++k1)
            for (int j = 0; j < 64; ++j) {
                const double ang = -2.0 * kPi * double(j * k1) / 1024.0;
                ta[k1 * 64 + j] = make_double2(std::cos(ang), std::sin(ang));
            }
        for (int k2 = 0; k2 < 16; ++k2)
            for (int j0 = 0; j0 < 4; ++j0) {
                const double ang = -2.0 * kPi * double((j0 * k2) & 63) / 64.0;
                tbb[k2 * 4 + j0] = make_double2(std::cos(ang), std::sin(ang));
            }
        if ((rc = upload(h, ta, &tb.t64a))) return rc;
        if ((rc = upload(h, tbb, &tb.t64b))) return rc;
    }
    if ((rc = upload(h, twiddles(512, 257), &tb.ptw512))) return rc;
    if ((rc = upload(h, twiddles(2048, 1025), &tb.ptw2048))) return rc;
    {
        auto to_f = [](const std::vector<double2>& d) {
            std::vector<float2> f(d.size());
            for (size_t i = 0; i < d.size(); ++i) f[i] = make_float2((float)d[i].x, (float)d[i].y);
            return f;
        };
        for (int span : {8000, 400}) {                               // per-pass twiddles of the radix-20 FFT-8000
            const int q = span / 20;
            std::vector<float2> t((size_t)20 * q);
            for (int k = 0; k < 20; ++k)
                for (int pos = 0; pos < q; ++pos) {
                    const double ang = -2.0 * kPi * double(k * pos) / double(span);
                    t[(size_t)k * q + pos] = make_float2((float)std::cos(ang), (float)std::sin(ang));
                }
            if ((rc = upload(h, t, span == 8000 ? &tb.tw20a : &tb.tw20b))) return rc;
        }
        if ((rc = upload(h, to_f(twiddles(16000, 8001)), &tb.ptw16000f))) return rc;
        {
            std::vector<unsigned short> pos(kH20N + 1);
            for (int k = 0; k <= kH20N; ++k) pos[k] = (unsigned short)h20_pad(h20_pos(k % kH20N));
            if ((rc = upload(h, pos, &tb.h20pos))) return rc;
        }
    }
    tb.tw_long = nullptr;
    tb.ptw_long = nullptr;
    if (h->g.long_mode) {
        const int N = h->g.L / 2;
        std::vector<float2> a((size_t)N), b((size_t)N + 1);
        for (int j = 0; j < N; ++j) {
            const double ang = -2.0 * kPi * double(j) / double(N);
            a[j] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        for (int k = 0; k <= N; ++k) {
            const double ang = -2.0 * kPi * double(k) / double(2 * N);
            b[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        if ((rc = upload(h, a, &tb.tw_long))) return rc;
        if ((rc = upload(h, b, &tb.ptw_long))) return rc;
    }
    if ((rc = upload_bank(h, mel_bank(p.sr, 512, 128, 0.0, p.fmax), &tb.mel_a))) return rc;
    if ((rc = upload_bank(h, mel_bank(p.sr, 512, 128, 0.0, p.sr / 2.0), &tb.mel_b))) return rc;
    if ((rc = upload_bank(h, mel_bank(p.sr, 512, 64, 0.0, p.sr / 2.0), &tb.mel_c))) return rc;
    {
        // BPC_MELD_SHIFT=0: the natural band starts (A/B of the bank-conflict-free layout)
        const char* e = std::getenv("BPC_MELD_SHIFT");
        SparseBank md = mel_bank(p.sr, 2048, 128, 0.0, p.sr / 2.0);
        if (!(e && std::atoi(e) == 0)) md = deconflict_bank(md, 32);
        if ((rc = upload_bank(h, md, &tb.mel_d))) return rc;
    }
    if ((rc = upload(h, dct2_ortho(40, 128), &tb.dct_mel))) return rc;
    {
        const int T = h->g.T;
        std::vector<float> d = dct2_ortho(T, T), dt(size_t(T) * T);
        for (int u = 0; u < T; ++u)
            for (int t = 0; t < T; ++t) dt[size_t(t) * T + u] = d[size_t(u) * T + t];     // device wants [t][u]
        if ((rc = upload(h, dt, &tb.dct_time))) return rc;
        tb.dct_time_n = nullptr;
        tb.dct_tiles = nullptr;
        tb.dct_colsum = nullptr;
        if (h->g.long_mode) {
            if ((rc = upload(h, d, &tb.dct_time_n))) return rc;
            std::vector<float> cs((size_t)T);
            for (int u = 0; u < T; ++u) {
                double acc = 0.0;
                for (int t = 0; t < T; ++t) acc += (double)d[size_t(u) * T + t];
                cs[u] = (float)acc;
            }
            if ((rc = upload(h, cs, &tb.dct_colsum))) return rc;
            uint32_t* tiles = nullptr;
            if ((rc = dalloc(h, tc_tile_words(T, T), &tiles))) return rc;
            launch_tc_prep_b(h->g, tb, tiles, 0);
            BPC_CUDA(h, cudaDeviceSynchronize());
            tb.dct_tiles = tiles;
        }
    }
    std::vector<double> edges = tuning_edges();
    if ((rc = upload(h, edges, &tb.hist_edges))) return rc;
    {
        std::vector<float> all;
        all.reserve(size_t(kNumTunings) * 12 * 257);
        for (int i = 0; i < kNumTunings; ++i) {
            std::vector<float> c = chroma_bank(p.sr, 512, edges[i]);
            all.insert(all.end(), c.begin(), c.end());
        }
        if ((rc = upload(h, all, &tb.chroma))) return rc;
    }
    {
        // The sparsified basis rows (util.sparsify_rows, 8..16 non-zeros) are stored as dense bands: row r covers the
        // bins [start, start + kCqtEllWidth) with zeros in the few gaps, so the kernel walks a row without column
        // indices (a zero weight adds +-0 to the accumulators: the sums stay bit-identical to the sparse product).
        std::vector<int16_t> start;
        std::vector<float> re, im;
        std::vector<double> sl;
        int gw[3] = {1, 1, 1};
        for (int i = 0; i < kNumTunings; ++i) {
            CqtBasisEll e = cqt_basis(p.sr, edges[i]);
            if (e.col.empty()) { h->err = "CQT basis row wider than the ELL width"; return BPC_ERR_UNSUPPORTED; }
            std::vector<float> bre((size_t)kCqtBinsPerOct * kCqtEllWidth, 0.f), bim(bre.size(), 0.f);
            for (int r = 0; r < kCqtBinsPerOct; ++r) {
                int lo = 1 << 30, hi = -1;
                for (int j = 0; j < kCqtEllWidth; ++j) {
                    const int c = e.col[r * kCqtEllWidth + j];
                    if (c >= 0) { lo = std::min(lo, c); hi = std::max(hi, c); }
                }
                if (hi < 0) { lo = 64; hi = 64; }
                if (lo < 60 || hi > 144) { h->err = "CQT basis support outside the staged bin window"; return BPC_ERR_UNSUPPORTED; }
                if (hi - lo + 1 > kCqtEllWidth) { h->err = "CQT basis row wider than the band width"; return BPC_ERR_UNSUPPORTED; }
                for (int j = 0; j < kCqtEllWidth; ++j) {
                    const int c = e.col[r * kCqtEllWidth + j];
                    if (c >= 0) {
                        bre[(size_t)r * kCqtEllWidth + (c - lo)] = e.re[r * kCqtEllWidth + j];
                        bim[(size_t)r * kCqtEllWidth + (c - lo)] = e.im[r * kCqtEllWidth + j];
                    }
                }
                start.push_back((int16_t)lo);
                // rounds of the basis product in k_cens (cens_round_row): rows 20-35, 4-19, 0-3 -- the band widths grow
                // with the row, so the four-row round is the narrow one
                const int q = r >= 20 ? 0 : (r >= 4 ? 1 : 2);
                gw[q] = std::max(gw[q], hi - lo + 1);
            }
            re.insert(re.end(), bre.begin(), bre.end());
            im.insert(im.end(), bim.begin(), bim.end());
            sl.insert(sl.end(), e.sqrt_len.begin(), e.sqrt_len.end());
        }
        if ((rc = upload(h, start, &tb.cqt_start))) return rc;
        if ((rc = upload(h, re, &tb.cqt_re))) return rc;
        if ((rc = upload(h, im, &tb.cqt_im))) return rc;
        if ((rc = upload(h, sl, &tb.cqt_sqrt_len))) return rc;
        // the same tables as per-tuning blocks in the layout the CQT kernels keep in shared memory (kernels.cuh::CqBlock)
        {
            std::vector<CqBlock> blocks(kNumTunings);
            std::memset(blocks.data(), 0, blocks.size() * sizeof(CqBlock));
            const double sqrt2 = std::sqrt(2.0), pi = 3.14159265358979323846;
            double wsum = 0.0;                                          // hann(43) normalised to unit sum, in index order
            for (int j = 0; j < 43; ++j) wsum += 0.5 - 0.5 * std::cos(2.0 * pi * (double)j / 42.0);
            for (int i = 0; i < kNumTunings; ++i) {
                CqBlock& B = blocks[(size_t)i];
                const float* bre = re.data() + (size_t)i * kCqtBinsPerOct * kCqtEllWidth;
                const float* bim = im.data() + (size_t)i * kCqtBinsPerOct * kCqtEllWidth;
                int cnt[kCqtBinsPerOct];
                for (int r = 0; r < kCqtBinsPerOct; ++r) {
                    cnt[r] = 1;                                         // taps up to the row's last non-zero weight
                    for (int jj = 0; jj < kCqtEllWidth; ++jj) {
                        const float wr = bre[r * kCqtEllWidth + jj], wi = bim[r * kCqtEllWidth + jj];
                        const size_t o = (size_t)r * kCqWPitch + kCqPadL + jj;
                        B.wpad[0][o] = make_float2(wr, wi);
                        // fft_basis *= sqrt(sr / my_sr) rounded to complex64: odd octaves carry a factor sqrt(2); the
                        // remaining power of two is applied to the (linear) response, which is exact
                        B.wpad[1][o] = make_float2((float)((double)wr * sqrt2), (float)((double)wi * sqrt2));
                        if (jj >= 1 && (wr != 0.f || wi != 0.f)) cnt[r] = jj + 1;
                    }
                }
                for (int p3 = 0; p3 < kCqTriples; ++p3) {
                    const int16_t* bs = start.data() + (size_t)i * kCqtBinsPerOct + 3 * p3;
                    const int s0 = bs[0], d1 = bs[1] - s0, d2 = bs[2] - s0;
                    const int u = std::max(cnt[3 * p3], std::max(d1 + cnt[3 * p3 + 1], d2 + cnt[3 * p3 + 2]));
                    if (d1 < 0 || d2 < d1 || d2 > kCqPadL || u > kCqWPitch - kCqPadL || s0 - 60 < 0 || s0 - 60 + u > 89) {
                        h->err = "CQT basis rows do not fit the padded triple layout";
                        return BPC_ERR_UNSUPPORTED;
                    }
                    B.tri_s[p3] = (short)(s0 - 60);
                    B.tri_d1[p3] = (short)d1;
                    B.tri_d2[p3] = (short)d2;
                    B.tri_u[p3] = (short)u;
                }
                for (int k = 0; k < kCqtBins; ++k) B.inv_sl[k] = 1.0 / sl[(size_t)i * kCqtBins + k];
                // scipy.signal.get_window('hann', 43, fftbins=False) / sum
                for (int j = 0; j < 43; ++j) B.swin[j] = (0.5 - 0.5 * std::cos(2.0 * pi * (double)j / 42.0)) / wsum;
            }
            if ((rc = upload(h, blocks, &tb.cq_blocks))) return rc;
        }
        for (int q = 0; q < 3; ++q) tb.cqt_gw[q] = gw[q];
    }
    {
        std::vector<double> hb = halfband_taps();
        for (int d = 1; d <= 63; ++d)
            if (hb[63 + d] != hb[63 - d] || ((d & 1) == 0 && std::fabs(hb[63 + d]) > 1e-15)) {
                h->err = "half-band taps are not symmetric / half-band";
                return BPC_ERR_UNSUPPORTED;
            }
        upload_cens_constants(hb.data());
        upload_lpc_constants(hamming_sym(400).data());
    }
    return BPC_OK;
}

int build_workspace(bpc_handle* h, ChunkCtx& ctx, int cap) {
    const Geometry& g = h->g;
    Workspace& w = ctx.ws;
    const size_t C = (size_t)cap, T = (size_t)g.T;
    int rc;
    ctx.cap = cap;
    w.cap = cap;
    for (auto& s : ctx.side) BPC_CUDA(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&ctx.ev_fork, &ctx.ev_spec512, &ctx.ev_time, &ctx.ev_f2048, &ctx.ev_seg, &ctx.ws_free})
        BPC_CUDA(h, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if ((rc = dalloc(h, C * g.L, &w.y))) return rc;
    if ((rc = dalloc(h, C * T * kMagStride, &w.mag512))) return rc;
    if ((rc = dalloc(h, C * ((T + 1) / 2) * kMag2048Stride, &w.mag_even))) return rc;
    w.cand36 = nullptr;
    if (!g.long_mode && (rc = dalloc(h, C * 2 * (size_t)kMaxCand2048, &w.cand36))) return rc;
    if ((rc = dalloc(h, C * T * 20, &w.frame_feat))) return rc;
    if ((rc = dalloc(h, C * T * 128, &w.melD))) return rc;
    w.dec_stride = cens_dec_floats_per_segment(g.L);
    if ((rc = dalloc(h, C * (size_t)w.dec_stride, &w.dec))) return rc;
    w.cens_lo = nullptr;
    w.lpc_coef = nullptr;
    w.lpc_redo = nullptr;
    if (!g.long_mode && (rc = dalloc(h, C * 12 * (size_t)g.lpc_frames, &w.lpc_coef))) return rc;
    if (!g.long_mode && (rc = dalloc(h, 1 + C * (size_t)g.lpc_frames, &w.lpc_redo))) return rc;
    if (!g.long_mode && (rc = dalloc(h, C * 3 * 12 * T, &w.cens_lo))) return rc;
    if ((rc = dalloc(h, C * 2, &w.tuning))) return rc;
    if ((rc = dalloc(h, C * 2, &w.chroma_min))) return rc;
    if ((rc = dalloc(h, C * 2, &w.ints))) return rc;
    w.scratch = nullptr;
    w.scratch_stride = 0;
    w.tc_a = nullptr;
    if (g.long_mode && (rc = dalloc(h, tc_tile_words(cap * 40, g.T), &w.tc_a))) return rc;
    w.tc_mean = nullptr;
    if (g.long_mode && (rc = dalloc(h, (size_t)cap * 40, &w.tc_mean))) return rc;
    if (g.long_mode) {
        // per-segment scratch of the kernel that needs most (kernels of a chunk run one after the other in long mode)
        size_t need = consumer_scratch_floats(g.T);
        need = std::max(need, (size_t)1200 * T + 8192);
        w.scratch_stride = (need + 63) & ~size_t(63);
        if ((rc = dalloc(h, C * w.scratch_stride, &w.scratch))) return rc;
    }
    if (!h->stats_acc && (rc = dalloc(h, (size_t)(9 + g.nscal) * 5, &h->stats_acc))) return rc;
    // 1 s mode: the producers of the planes accumulate the dataset statistics themselves instead of k_stats re-reading
    // the planes (0.8 GB less DRAM traffic per 4096-segment step; +0.15 ms in the producers against 0.18 ms of
    // k_stats).  BPC_FUSED_STATS=0 and the long mode keep k_stats.
    {
        const char* fs = std::getenv("BPC_FUSED_STATS");
        w.stats_acc = (!g.long_mode && !(fs && fs[0] == '0')) ? h->stats_acc : nullptr;
    }
    w.dbg_mel_db = w.dbg_mfcc = w.dbg_gam = w.dbg_mod = w.dbg_chroma_stft = w.dbg_chroma_cens = w.dbg_lpc =
        w.dbg_onset = nullptr;
    ctx.ws_dbg = w;
    return BPC_OK;
}

int ensure_debug(bpc_handle* h) {
    Workspace& w = h->main.ws_dbg;
    if (w.dbg_mel_db) return BPC_OK;
    const Geometry& g = h->g;
    const size_t C = (size_t)h->chunk, T = (size_t)g.T;
    int rc;
    if ((rc = dalloc(h, C * 128 * T, &w.dbg_mel_db))) return rc;
    if ((rc = dalloc(h, C * 120 * T, &w.dbg_mfcc))) return rc;
    if ((rc = dalloc(h, C * 64 * T, &w.dbg_gam))) return rc;
    if ((rc = dalloc(h, C * 40 * T, &w.dbg_mod))) return rc;
    if ((rc = dalloc(h, C * 12 * T, &w.dbg_chroma_stft))) return rc;
    if ((rc = dalloc(h, C * 12 * T, &w.dbg_chroma_cens))) return rc;
    if ((rc = dalloc(h, C * 12 * (size_t)g.lpc_frames, &w.dbg_lpc))) return rc;
    if ((rc = dalloc(h, C * T, &w.dbg_onset))) return rc;
    return BPC_OK;
}

int reset_stats(bpc_handle* h, cudaStream_t st) {
    const int rows = 9 + h->g.nscal;
    std::vector<double> init((size_t)rows * 5, 0.0);
    for (int r = 0; r < rows; ++r) { init[r * 5 + 3] = 1e300; init[r * 5 + 4] = -1e300; }
    BPC_CUDA(h, cudaMemcpyAsync(h->stats_acc, init.data(), init.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    BPC_CUDA(h, cudaStreamSynchronize(st));
    return BPC_OK;
}

// The launch sequence for one chunk of n <= chunk segments; inputs / outputs are device pointers for this chunk.
int run_chunk(bpc_handle* h, ChunkCtx& cx, const void* wav, int wav_dtype, int64_t L_in, int n, float* feats,
              float* scalars, int32_t* status, cudaStream_t st) {
    const Geometry& g = h->g;
    const Workspace& ws = (h->debug && &cx == &h->main) ? cx.ws_dbg : cx.ws;
    // One workspace per handle: whatever stream the previous enqueue used, this one starts after it (calls on a handle
    // are serialised by the caller on the host, not necessarily on one stream).
    if (cx.ws_used) BPC_CUDA(h, cudaStreamWaitEvent(st, cx.ws_free, 0));
    const float* y;
    if (wav_dtype == BPC_WAV_F32 && L_in == g.L && (reinterpret_cast<uintptr_t>(wav) & 15) == 0) {
        y = static_cast<const float*>(wav);                   // pad_or_truncate is the identity: no copy
    } else {
        launch_ingest(wav, wav_dtype, L_in, ws.y, n, g, st);
        y = ws.y;
    }
    if (status) BPC_CUDA(h, cudaMemsetAsync(status, 0, sizeof(int32_t) * n, st));
    auto timed = [&](int id, auto&& fn) {
        if (!h->timing) { fn(); return; }
        bpc_handle::Ev e{id, nullptr, nullptr};
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
        cudaEventRecord(e.a, st);
        fn();
        cudaEventRecord(e.b, st);
        h->evs.push_back(e);
    };
    if (g.long_mode) {
        // long mode: one stream (the kernels share the scratch region)
        timed(1, [&] { launch_stft512(y, n, g, h->tb, ws, st); });
        timed(2, [&] { launch_spec512_consumers(n, g, h->tb, ws, feats, scalars, status, true, st); });
        timed(3, [&] { launch_spec2048(y, n, g, h->tb, ws, feats, scalars, st); });
        timed(4, [&] { launch_even2048(n, g, h->tb, ws, scalars, status, st); });
        timed(10, [&] { launch_seg2048(n, g, h->tb, ws, feats, scalars, st); });
        timed(11, [&] { launch_cens(y, n, g, h->tb, ws, feats, st, 1); });
        timed(5, [&] { launch_cens(y, n, g, h->tb, ws, feats, st, 2); });
        timed(8, [&] { launch_lpc(y, n, g, h->tb, ws, feats, st); });
        timed(6, [&] { launch_time_scalars(y, n, g, h->tb, ws, scalars, status, st); });
        timed(7, [&] { launch_hilbert(y, n, g, h->tb, ws, scalars, st); });
    } else if (h->timing || !h->multi_stream) {
        timed(1, [&] { launch_stft512(y, n, g, h->tb, ws, st); });
        timed(2, [&] { launch_spec512_consumers(n, g, h->tb, ws, feats, scalars, status, true, st); });
        timed(3, [&] { launch_spec2048(y, n, g, h->tb, ws, feats, scalars, st); });
        timed(4, [&] { launch_even2048(n, g, h->tb, ws, scalars, status, st); });
        timed(10, [&] { launch_seg2048(n, g, h->tb, ws, feats, scalars, st); });
        timed(11, [&] { launch_cens(y, n, g, h->tb, ws, feats, st, 1); });
        timed(12, [&] { launch_cens(y, n, g, h->tb, ws, feats, st, 4); });
        timed(5, [&] { launch_cens(y, n, g, h->tb, ws, feats, st, 3); });
        timed(6, [&] { launch_time_scalars(y, n, g, h->tb, ws, scalars, status, st); });
        timed(7, [&] { launch_hilbert(y, n, g, h->tb, ws, scalars, st); });
        timed(8, [&] { launch_lpc(y, n, g, h->tb, ws, feats, st); });
    } else {
        // Dependencies: stft512 -> consumers (mag512); spec2048 -> {even2048, seg2048} (mag_even, frame_feat, melD);
        // {consumers (chroma_min), even2048 (tuning-36)} -> cens; the time-domain kernels only read y.  Every branch
        // joins before the statistics kernel, so consecutive chunks (one shared workspace) stay ordered.
        cudaStream_t sA = cx.side[0], sC = cx.side[1], sD = cx.side[2];
        BPC_CUDA(h, cudaEventRecord(cx.ev_fork, st));
        BPC_CUDA(h, cudaStreamWaitEvent(sA, cx.ev_fork, 0));
        BPC_CUDA(h, cudaStreamWaitEvent(sC, cx.ev_fork, 0));
        launch_spec2048(y, n, g, h->tb, ws, feats, scalars, st);
        BPC_CUDA(h, cudaEventRecord(cx.ev_f2048, st));
        launch_stft512(y, n, g, h->tb, ws, sA);
        launch_spec512_consumers(n, g, h->tb, ws, feats, scalars, status, true, sA);
        BPC_CUDA(h, cudaEventRecord(cx.ev_spec512, sA));
        launch_even2048(n, g, h->tb, ws, scalars, status, st);
        BPC_CUDA(h, cudaStreamWaitEvent(sD, cx.ev_f2048, 0));
        launch_seg2048(n, g, h->tb, ws, feats, scalars, sD);
        BPC_CUDA(h, cudaEventRecord(cx.ev_seg, sD));
        launch_lpc(y, n, g, h->tb, ws, feats, sC);
        launch_time_scalars(y, n, g, h->tb, ws, scalars, status, sC);
        launch_hilbert(y, n, g, h->tb, ws, scalars, sC);
        BPC_CUDA(h, cudaEventRecord(cx.ev_time, sC));
        BPC_CUDA(h, cudaStreamWaitEvent(st, cx.ev_spec512, 0));
        launch_cens(y, n, g, h->tb, ws, feats, st);
        BPC_CUDA(h, cudaStreamWaitEvent(st, cx.ev_seg, 0));
        BPC_CUDA(h, cudaStreamWaitEvent(st, cx.ev_time, 0));
    }
    launch_pad_scalars(n, g, scalars, st);
    timed(9, [&] { launch_stats(n, g, feats, scalars, h->stats_acc, ws.stats_acc == nullptr, st); });
    h->last_n = n;
    BPC_CUDA(h, cudaGetLastError());
    BPC_CUDA(h, cudaEventRecord(cx.ws_free, st));
    cx.ws_used = true;
    return BPC_OK;
}

int host_pinned(bpc_handle* h, size_t bytes, void** out) {
    void* p = numa_pinned_alloc(h->numa, bytes, h->numa_place);
    if (!p) { h->err = "pinned host allocation failed"; return BPC_ERR_ALLOC; }
    h->numa_allocs.push_back({p, bytes});
    *out = p;
    return BPC_OK;
}

int ensure_slots(bpc_handle* h) {
    if (h->slots_ready) return BPC_OK;
    const Geometry& g = h->g;
    const char* env_hc = std::getenv("BPC_HOST_CHUNK");
    // pieces of the host path: small enough that the D2H of a piece starts early (1 s: 296 segments = two per SM)
    h->host_chunk = env_hc ? std::atoi(env_hc) : (h->g.long_mode ? std::max(1, h->chunk / 2) : std::min(296, h->chunk));
    if (h->host_chunk < 1 || h->host_chunk > h->chunk) h->host_chunk = h->chunk;
    const size_t C = (size_t)h->host_chunk;                             // the host path moves pieces of host_chunk segments
    h->slot_wav_bytes = C * (size_t)g.L * 4 * 2;                        // room for L_in up to 2 * L of float32
    const size_t feats_bytes = C * 9 * kPlaneRows * (size_t)g.T * 4, scal_bytes = C * (size_t)g.nscal * 4;
    const size_t rows_bytes = C * (size_t)kLiveTotal * (size_t)g.T * 4;
    int rc;
    for (int i = 0; i < kSlots; ++i) {
        Slot& s = h->slot[i];
        BPC_CUDA(h, cudaMalloc(&s.d_wav, h->slot_wav_bytes));
        BPC_CUDA(h, cudaMalloc((void**)&s.d_feats, feats_bytes));
        BPC_CUDA(h, cudaMalloc((void**)&s.d_scalars, scal_bytes));
        BPC_CUDA(h, cudaMalloc((void**)&s.d_status, C * 4));
        BPC_CUDA(h, cudaMalloc((void**)&s.d_rows, rows_bytes));
        BPC_CUDA(h, cudaMalloc((void**)&s.d_fill, C * 9 * 4));
        for (void* d : {(void*)s.d_wav, (void*)s.d_feats, (void*)s.d_scalars, (void*)s.d_status, (void*)s.d_rows, (void*)s.d_fill})
            h->dev_allocs.push_back(d);
        // pinned staging on the GPU's NUMA node (used when the caller's buffers are pageable)
        if ((rc = host_pinned(h, h->slot_wav_bytes, &s.h_wav))) return rc;
        if ((rc = host_pinned(h, feats_bytes, (void**)&s.h_feats))) return rc;
        if ((rc = host_pinned(h, scal_bytes, (void**)&s.h_scalars))) return rc;
        if ((rc = host_pinned(h, C * 4, (void**)&s.h_status))) return rc;
        if ((rc = host_pinned(h, C * 9 * 4, (void**)&s.h_fill))) return rc;
        BPC_CUDA(h, cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        {
            int prio_lo = 0, prio_hi = 0;
            BPC_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            BPC_CUDA(h, cudaStreamCreateWithPriority(&s.st_out, cudaStreamNonBlocking, prio_hi));
        }
        BPC_CUDA(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        BPC_CUDA(h, cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming));
        BPC_CUDA(h, cudaEventCreateWithFlags(&s.fill_ready, cudaEventDisableTiming));
    }
    {
        // r02: with a workspace per slot consecutive pieces overlap on the GPU, but then every piece FINISHES later and
        // its D2H starts later -- measured 19-20.5 ms per 4096-segment call against 17.9-18.6 ms serialised.  Off by default.
        const char* env_sc = std::getenv("BPC_SLOT_CTX");
        h->slot_ctx_on = env_sc && std::atoi(env_sc) == 1;
        const char* env_rp = std::getenv("BPC_RAMP");
        h->ramp_head = !(env_rp && std::atoi(env_rp) == 0);
    }
    for (int i = 0; h->slot_ctx_on && i < kSlots; ++i)
        if ((rc = build_workspace(h, h->slot_ctx[i], h->host_chunk))) return rc;
    BPC_CUDA(h, cudaMalloc((void**)&h->live_dev, sizeof(kLiveRows)));
    h->dev_allocs.push_back(h->live_dev);
    BPC_CUDA(h, cudaMemcpy(h->live_dev, kLiveRows, sizeof(kLiveRows), cudaMemcpyHostToDevice));
    // Host threads (pad-row fill of the full layout, staging copies of pageable buffers): this rank's share of the
    // cores, at most 8 -- the fill of a 4096-segment call is 0.39 GB of streaming stores, ~10 GB/s per thread.
    const char* env_t = std::getenv("BPC_HOST_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    const char* env_lws = std::getenv("LOCAL_WORLD_SIZE");
    const int lws = std::max(1, env_lws ? std::atoi(env_lws) : 1);
    int nt = env_t ? std::atoi(env_t) : std::min(8, std::max(1, (hw > 0 ? hw : 8) / lws));
    if (hw > 0 && nt > hw) nt = hw;
    if (nt < 1) nt = 1;
    h->pool = new HostPool(nt - 1, h->numa_place ? h->numa.cpus : std::vector<int>());
    const char* env_c = std::getenv("BPC_COMPACT_D2H");
    h->compact_d2h = !(env_c && std::atoi(env_c) == 0);
    const char* env_m = std::getenv("BPC_D2H_MODE");
    h->contig_d2h = !(env_m && std::string(env_m) == "2d");
    const char* env_tp = std::getenv("BPC_TAPER");
    h->taper_tail = !(env_tp && std::atoi(env_tp) == 0);
    h->slots_ready = true;
    return BPC_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

// ================================================================================================= ABI
extern "C" {

int bpc_abi_version(void) { return BPC_ABI_VERSION; }

void bpc_default_params(bpc_params* p) {
    if (!p) return;
    p->sr = 16000; p->n_fft = 512; p->hop = 256; p->n_mels = 128; p->n_mfcc = 40; p->fmax = 4500.f;
    p->n_gammatone = 64; p->n_lpc = 12; p->expected_len = 16000; p->pad_scalars_to = 0;
}

int bpc_num_frames(const bpc_params* p) { return p ? p->expected_len / p->hop + 1 : BPC_ERR_ARG; }
int bpc_num_scalars(const bpc_params* p) {
    if (!p) return BPC_ERR_ARG;
    return p->pad_scalars_to > BPC_NUM_SCALARS ? p->pad_scalars_to : BPC_NUM_SCALARS;
}

int bpc_create(bpc_handle** out, const bpc_params* p, int device, int64_t max_batch) {
    if (!out) { g_create_error = "out is NULL"; return BPC_ERR_ARG; }
    *out = nullptr;
    std::string why;
    int rc = check_params(p, &why);
    if (rc) { g_create_error = why; return rc; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        cudaGetLastError();
        return BPC_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return BPC_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return BPC_ERR_CUDA; }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        g_create_error = "libbpc_b200 is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) +
                         std::to_string(prop.minor);
        return BPC_ERR_UNSUPPORTED;
    }
    bpc_handle* h = new bpc_handle();
    h->p = *p;
    h->device = device;
    h->max_batch = max_batch > 0 ? max_batch : 1;
    h->g.L = p->expected_len;
    h->g.hop = p->hop;
    h->g.T = p->expected_len / p->hop + 1;
    h->g.nscal = bpc_num_scalars(p);
    h->g.lpc_frames = (p->expected_len - 400 + 159) / 160;
    h->g.long_mode = p->expected_len > 16000 ? 1 : 0;
    const char* env_chunk = std::getenv("BPC_CHUNK");
    // 28 segments per SM per launch: every kernel launch ends in a tail during which SMs drain, and fewer, longer
    // launches amortise it (592 -> 4096 segments per launch measured +4.4 %; the workspace is 0.33 MB per segment)
    int chunk = env_chunk ? std::atoi(env_chunk) : 4144;
    if (chunk < 1) chunk = 4144;
    // long mode: about the same samples per chunk, but never below one CTA-per-segment wave of the 148 SMs (workspace:
    // ~18 MB per 30 s segment)
    if (h->g.long_mode && !env_chunk) chunk = std::max(148, chunk / (p->expected_len / 16000));
    h->chunk = (int)std::min<int64_t>(chunk, h->max_batch);
    h->launches0 = launches_issued();
    const char* env_streams = std::getenv("BPC_STREAMS");
    h->multi_stream = !(env_streams && std::atoi(env_streams) == 0);
    {
        const char* env_numa = std::getenv("BPC_NUMA");
        h->numa_place = !(env_numa && std::atoi(env_numa) == 0);
        h->numa = numa_of_device(device);
    }
    if ((rc = build_tables(h)) || (rc = build_workspace(h, h->main, h->chunk)) || (rc = reset_stats(h, 0))) {
        g_create_error = h->err;
        bpc_destroy(h);
        return rc;
    }
    *out = h;
    return BPC_OK;
}

void bpc_destroy(bpc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kSlots; ++i) {
        if (h->slot[i].st) cudaStreamDestroy(h->slot[i].st);
        if (h->slot[i].st_out) cudaStreamDestroy(h->slot[i].st_out);
        if (h->slot[i].done) cudaEventDestroy(h->slot[i].done);
        if (h->slot[i].computed) cudaEventDestroy(h->slot[i].computed);
        if (h->slot[i].fill_ready) cudaEventDestroy(h->slot[i].fill_ready);
    }
    delete h->pool;
    for (ChunkCtx* cx : {&h->main, &h->slot_ctx[0], &h->slot_ctx[1], &h->slot_ctx[2]}) {
        for (auto& s : cx->side) if (s) cudaStreamDestroy(s);
        for (cudaEvent_t e : {cx->ev_fork, cx->ev_spec512, cx->ev_time, cx->ev_f2048, cx->ev_seg, cx->ws_free}) if (e) cudaEventDestroy(e);
    }
    for (void* d : h->dev_allocs) cudaFree(d);
    for (void* d : h->host_allocs) cudaFreeHost(d);
    for (auto& a : h->numa_allocs) numa_pinned_free(a.first, a.second);
    delete h;
}

const char* bpc_last_error(const bpc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int bpc_set_debug(bpc_handle* h, int on) {
    if (!h) return BPC_ERR_ARG;
    if (on) {
        int rc = ensure_debug(h);
        if (rc) return rc;
    }
    h->debug = on != 0;
    return BPC_OK;
}

int bpc_precompute(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, float* feats,
                   float* scalars, int32_t* status, void* stream) {
    if (!h) return BPC_ERR_ARG;
    if (!wav || !feats || !scalars || B < 0 || L_in <= 0 || (wav_dtype != BPC_WAV_F32 && wav_dtype != BPC_WAV_PCM16)) {
        h->err = "bpc_precompute: bad argument";
        return BPC_ERR_ARG;
    }
    BPC_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Geometry& g = h->g;
    const size_t esz = wav_dtype == BPC_WAV_F32 ? 4 : 2;
    for (int64_t off = 0; off < B; off += h->chunk) {
        const int n = (int)std::min<int64_t>(h->chunk, B - off);
        int rc = run_chunk(h, h->main, static_cast<const char*>(wav) + (size_t)off * L_in * esz, wav_dtype, L_in, n,
                           feats + (size_t)off * 9 * kPlaneRows * g.T, scalars + (size_t)off * g.nscal,
                           status ? status + off : nullptr, st);
        if (rc) return rc;
    }
    return BPC_OK;
}

}  // extern "C"

namespace {

// The host-buffer path: a three-stage software pipeline over pieces of host_chunk segments -- H2D of piece i+1 ||
// kernels of piece i || D2H (+ host fill, full layout only) of piece i-1 -- on three slots with their own streams.
// r02: the ring of pieces is a property of the HANDLE, not of a call.  host_begin enqueues the pieces of one call behind
// whatever is still in flight and returns with the last two pieces unretired; host_wait retires up to a ticket.  A caller
// that alternates two sets of host buffers (begin k+1, then wait k) keeps the GPU and the copy engines busy across
// calls: the ramp at the head of a call (nothing to copy until the first piece is computed) and the D2H of the last
// piece at its tail overlap with the neighbouring calls.  The synchronous entry points are begin + wait.
cudaError_t host_wait_event(bpc_handle* h, cudaEvent_t ev) {
    const auto t_a = std::chrono::steady_clock::now();
    const cudaError_t e = cudaEventSynchronize(ev);
    h->t_wait += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count();
    return e;
}

void host_job(bpc_handle* h, const std::function<void(int, int)>& job) {
    const auto t_a = std::chrono::steady_clock::now();
    h->pool->run(job);
    h->t_fill += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count();
}

// Full layout: the 380 constant pad rows of every segment of the piece, from its nine pad values (they leave the device
// before the bulk rows, so this overlaps the piece's D2H and the next piece's kernels).
int host_fill(bpc_handle* h, Fly& f) {
    if (f.to_rows || !f.live_only || f.filled) return BPC_OK;
    Slot& s = h->slot[f.slot];
    const int T = h->g.T, n = f.n;
    const size_t seg_feats = (size_t)9 * kPlaneRows * T;
    float* user = f.out.feats + (size_t)f.off * seg_feats;
    BPC_CUDA(h, host_wait_event(h, s.fill_ready));
    host_job(h, [&](int part, int parts) {
        for (int b = part; b < n; b += parts) {
            float* dst = user + (size_t)b * seg_feats;
            for (int c = 0; c < 9; ++c)
                if (kLiveRows[c] < kPlaneRows)
                    fill_stream(dst + ((size_t)c * kPlaneRows + kLiveRows[c]) * T,
                                dst + (size_t)(c + 1) * kPlaneRows * T, s.h_fill[(size_t)b * 9 + c]);
        }
#if defined(__SSE2__)
        _mm_sfence();
#endif
    });
    f.filled = true;
    return BPC_OK;
}

// Retire the oldest piece in flight: its bulk D2H has finished; staging -> user copies for pageable buffers.
int host_retire_front(bpc_handle* h) {
    Fly& f = h->fly.front();
    int rc = host_fill(h, f);
    if (rc) return rc;
    Slot& s = h->slot[f.slot];
    const Geometry& g = h->g;
    const int T = g.T, n = f.n;
    const int64_t off = f.off;
    const size_t seg_feats = (size_t)9 * kPlaneRows * T, seg_rows = (size_t)kLiveTotal * T;
    BPC_CUDA(h, host_wait_event(h, s.done));
    if (!f.pin_f) {                                                // pageable output: staging -> user buffer
        if (f.to_rows) {
            float* user = f.out.rows + (size_t)off * seg_rows;
            host_job(h, [&](int part, int parts) {
                for (int b = part; b < n; b += parts)
                    std::memcpy(user + (size_t)b * seg_rows, s.h_feats + (size_t)b * seg_rows, seg_rows * 4);
            });
        } else {
            float* user = f.out.feats + (size_t)off * seg_feats;
            const std::vector<RowRun> runs = live_runs();
            const bool live_only = f.live_only;
            host_job(h, [&](int part, int parts) {
                for (int b = part; b < n; b += parts) {
                    float* dst = user + (size_t)b * seg_feats;
                    const float* src = s.h_feats + (size_t)b * seg_feats;
                    if (live_only) {
                        for (const RowRun& r : runs)
                            std::memcpy(dst + (size_t)r.start * T, src + (size_t)r.start * T,
                                        (size_t)(r.end - r.start) * T * 4);
                    } else {
                        std::memcpy(dst, src, seg_feats * 4);
                    }
                }
            });
        }
    }
    if (f.to_rows && !f.pin_p) std::memcpy(f.out.pad + (size_t)off * 9, s.h_fill, (size_t)n * 9 * 4);
    if (!f.pin_s) std::memcpy(f.scalars + (size_t)off * g.nscal, s.h_scalars, (size_t)n * g.nscal * 4);
    if (f.status) std::memcpy(f.status + off, s.h_status, (size_t)n * 4);
    if (f.last) h->ticket_done = f.ticket;
    h->fly.pop_front();
    return BPC_OK;
}

// An error in the middle of the pipeline: copies into the caller's buffers may still be in flight -- let them land,
// then forget every piece (their tickets count as done; the error has been reported).
int host_abort(bpc_handle* h, int rc) {
    for (int i = 0; i < kSlots; ++i) {
        if (h->slot[i].st) cudaStreamSynchronize(h->slot[i].st);
        if (h->slot[i].st_out) cudaStreamSynchronize(h->slot[i].st_out);
    }
    cudaGetLastError();
    h->fly.clear();
    h->ticket_done = h->ticket_next - 1;
    return rc;
}

int host_wait(bpc_handle* h, int64_t ticket) {
    if (ticket < 0 || ticket >= h->ticket_next) ticket = h->ticket_next - 1;      // everything enqueued so far
    BPC_CUDA(h, cudaSetDevice(h->device));
    while (h->ticket_done < ticket && !h->fly.empty()) {
        const int rc = host_retire_front(h);
        if (rc) return host_abort(h, rc);
    }
    return BPC_OK;
}

struct HostTimeline { cudaEvent_t a, c, d; int n; };

int host_begin(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, HostOut out, float* scalars,
               int32_t* status, int64_t* ticket_out, std::vector<HostTimeline>* tl, int64_t* pieces_out) {
    BPC_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_slots(h);
    if (rc) return rc;
    const Geometry& g = h->g;
    const size_t esz = wav_dtype == BPC_WAV_F32 ? 4 : 2;
    if ((size_t)h->host_chunk * L_in * esz > h->slot_wav_bytes) { h->err = "L_in too large for the staging buffers"; return BPC_ERR_ARG; }
    const int T = g.T;
    const size_t seg_feats = (size_t)9 * kPlaneRows * T, seg_rows = (size_t)kLiveTotal * T;
    const bool to_rows = out.rows != nullptr;
    const bool pin_in = is_pinned(wav), pin_f = is_pinned(to_rows ? out.rows : out.feats), pin_s = is_pinned(scalars);
    const bool pin_p = to_rows && is_pinned(out.pad);
    const std::vector<RowRun> runs = live_runs();
    // Piece schedule: full pieces, then the last <= host_chunk segments in halves (not below 128 segments, about one CTA
    // wave): what is exposed at the end of a call is the D2H (+ fill) of the LAST piece only, so it should be small.
    struct Piece { int64_t off; int n; };
    std::vector<Piece> sched;
    {
        // Ramp: with an empty ring the D2H engine idles until the first piece has been computed, so the call starts with
        // pieces of about one and two CTA waves (148, 296 segments) before it settles at host_chunk.  Behind a call
        // that is still in flight there is nothing to ramp.
        int64_t off = 0;
        const bool cold = h->fly.empty();
        int ramp = (cold && h->ramp_head && !h->g.long_mode && B >= 4 * (int64_t)h->host_chunk) ? 148 : h->host_chunk;
        while (off < B) {
            const int64_t cap = std::min<int64_t>(ramp, h->host_chunk);
            int64_t rem = B - off, n = std::min<int64_t>(cap, rem);
            if (rem <= h->host_chunk && h->taper_tail && rem >= 256) n = std::max<int64_t>(128, rem / 2);
            sched.push_back({off, (int)n});
            off += n;
            ramp = (int)std::min<int64_t>(2 * (int64_t)ramp, h->host_chunk);
        }
    }
    if (to_rows && !h->contig_d2h)
        for (int64_t b = 0; b < B; ++b)
            for (int c = 0; c < 9; ++c)
                if (kLiveRows[c] == kPlaneRows) out.pad[b * 9 + c] = 0.f;
    const int64_t nchunks = (int64_t)sched.size();
    if (pieces_out) *pieces_out = nchunks;
    const bool live_only = to_rows || h->compact_d2h;        // only the 772 data rows cross PCIe
    const int64_t ticket = h->ticket_next++;
    if (ticket_out) *ticket_out = ticket;
    if (nchunks == 0) { if (h->fly.empty()) h->ticket_done = ticket; return BPC_OK; }
    // Piece i is enqueued while the two pieces before it are still in flight (the GPU always has the next piece
    // queued); then the pad rows of the piece before it are written (full layout) and the piece two back is retired.
    auto body = [&]() -> int {
        for (int64_t i = 0; i < nchunks; ++i) {
            while (h->fly.size() > 2) {                                    // (a begin right after an abort / odd ring states)
                const int rc2 = host_retire_front(h);
                if (rc2) return rc2;
            }
            const int slot_id = (int)(h->piece_seq % kSlots);
            Slot& s = h->slot[slot_id];
            const int64_t off = sched[i].off;
            const int n = sched[i].n;
            const char* src = static_cast<const char*>(wav) + (size_t)off * L_in * esz;
            const size_t in_bytes = (size_t)n * L_in * esz;
            if (tl) {
                HostTimeline e{nullptr, nullptr, nullptr, n};
                cudaEventCreate(&e.a); cudaEventCreate(&e.c); cudaEventCreate(&e.d);
                cudaEventRecord(e.a, s.st);
                tl->push_back(e);
            }
            if (pin_in) {
                BPC_CUDA(h, cudaMemcpyAsync(s.d_wav, src, in_bytes, cudaMemcpyHostToDevice, s.st));
            } else {
                std::memcpy(s.h_wav, src, in_bytes);
                BPC_CUDA(h, cudaMemcpyAsync(s.d_wav, s.h_wav, in_bytes, cudaMemcpyHostToDevice, s.st));
            }
            // The H2D above overlaps the previous piece's kernels; the kernels themselves are serialised by the
            // handle's workspace event (run_chunk), because all slots share the one workspace.
            // debug on: the raw stages of the last chunk are read from the main context afterwards (bpc_debug_copy)
            ChunkCtx& cx = (h->debug || !h->slot_ctx_on) ? h->main : h->slot_ctx[slot_id];
            int rc2 = run_chunk(h, cx, s.d_wav, wav_dtype, L_in, n, s.d_feats, s.d_scalars, s.d_status, s.st);
            if (rc2) return rc2;
            if (tl) cudaEventRecord(tl->back().c, s.st);
            BPC_CUDA(h, cudaEventRecord(s.computed, s.st));
            BPC_CUDA(h, cudaStreamWaitEvent(s.st_out, s.computed, 0));
            cudaStream_t so = s.st_out;                                // everything below: the piece's output side
            float* sdst = pin_s ? scalars + (size_t)off * g.nscal : s.h_scalars;
            if (to_rows) {
                float* pdst = pin_p ? out.pad + (size_t)off * 9 : s.h_fill;
                float* rdst = pin_f ? out.rows + (size_t)off * seg_rows : s.h_feats;
                if (h->contig_d2h) {
                    launch_pad_values(s.d_feats, T, n, h->live_dev, s.d_fill, so);
                    BPC_CUDA(h, cudaMemcpyAsync(pdst, s.d_fill, (size_t)n * 9 * 4, cudaMemcpyDeviceToHost, so));
                } else {
                    // copy-engine only (no kernel has to find a free SM next to the following piece's persistent
                    // CTAs): the pad value of plane c is the first element of its first pad row, a 4-byte column
                    // of the [n, 9 * 128 * T] matrix; planes without pad rows report 0 (memset once per call)
                    for (int c = 0; c < 9; ++c)
                        if (kLiveRows[c] < kPlaneRows)
                            BPC_CUDA(h, cudaMemcpy2DAsync(pdst + c, 9 * 4, s.d_feats + ((size_t)c * kPlaneRows + kLiveRows[c]) * T,
                                                          seg_feats * 4, 4, (size_t)n, cudaMemcpyDeviceToHost, so));
                }
                if (h->contig_d2h) {
                    launch_compact_rows(s.d_feats, T, n, kLiveRows, s.d_rows, so);
                    BPC_CUDA(h, cudaMemcpyAsync(rdst, s.d_rows, (size_t)n * seg_rows * 4, cudaMemcpyDeviceToHost, so));
                } else {
                    size_t row0 = 0;                               // first compact row of the run
                    for (const RowRun& r : runs) {
                        BPC_CUDA(h, cudaMemcpy2DAsync(rdst + row0 * T, seg_rows * 4, s.d_feats + (size_t)r.start * T,
                                                      seg_feats * 4, (size_t)(r.end - r.start) * T * 4, (size_t)n,
                                                      cudaMemcpyDeviceToHost, so));
                        row0 += (size_t)(r.end - r.start);
                    }
                }
            } else {
                float* fdst = pin_f ? out.feats + (size_t)off * seg_feats : s.h_feats;
                if (live_only) {
                    // PCIe carries only the rows that hold data (772 of the 1152 rows of a segment); the constant
                    // pad rows are re-created on the host from one value per plane.
                    launch_pad_values(s.d_feats, T, n, h->live_dev, s.d_fill, so);
                    BPC_CUDA(h, cudaMemcpyAsync(s.h_fill, s.d_fill, (size_t)n * 9 * 4, cudaMemcpyDeviceToHost, so));
                    BPC_CUDA(h, cudaEventRecord(s.fill_ready, so));
                    const size_t pitch = seg_feats * 4;
                    for (const RowRun& r : runs)
                        BPC_CUDA(h, cudaMemcpy2DAsync(fdst + (size_t)r.start * T, pitch, s.d_feats + (size_t)r.start * T,
                                                      pitch, (size_t)(r.end - r.start) * T * 4, (size_t)n,
                                                      cudaMemcpyDeviceToHost, so));
                } else {
                    BPC_CUDA(h, cudaMemcpyAsync(fdst, s.d_feats, (size_t)n * seg_feats * 4, cudaMemcpyDeviceToHost, so));
                }
            }
            BPC_CUDA(h, cudaMemcpyAsync(sdst, s.d_scalars, (size_t)n * g.nscal * 4, cudaMemcpyDeviceToHost, so));
            BPC_CUDA(h, cudaMemcpyAsync(s.h_status, s.d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
            BPC_CUDA(h, cudaEventRecord(s.done, so));
            if (tl) cudaEventRecord(tl->back().d, so);
            h->piece_seq++;
            h->fly.push_back(Fly{slot_id, off, n, out, scalars, status, to_rows, live_only, pin_f, pin_p, pin_s, false,
                                 i + 1 == nchunks, ticket});
            if (h->fly.size() >= 2) {                                      // pad rows of the piece before this one
                rc2 = host_fill(h, h->fly[h->fly.size() - 2]);
                if (rc2) return rc2;
            }
            while (h->fly.size() > 2) {                                    // retire the piece two back
                rc2 = host_retire_front(h);
                if (rc2) return rc2;
            }
        }
        return BPC_OK;
    };
    rc = body();
    if (rc) return host_abort(h, rc);
    return BPC_OK;
}

// Synchronous call = begin + wait (+ the optional trace lines).
int host_pipeline(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, HostOut out, float* scalars,
                  int32_t* status) {
    const auto t_call = std::chrono::steady_clock::now();
    const char* env_trace = std::getenv("BPC_HOST_TRACE");
    const bool timeline = env_trace && std::atoi(env_trace) >= 2;      // per-piece device timeline (debugging aid)
    std::vector<HostTimeline> tl;
    h->t_wait = h->t_fill = 0.0;
    int64_t ticket = 0, nchunks = 0;
    int rc = host_begin(h, wav, wav_dtype, B, L_in, out, scalars, status, &ticket, timeline ? &tl : nullptr, &nchunks);
    if (rc) return rc;
    rc = host_wait(h, ticket);
    if (rc) return rc;
    const bool to_rows = out.rows != nullptr;
    if (timeline) {
        cudaDeviceSynchronize();
        for (size_t i = 0; i < tl.size(); ++i) {
            float a = 0.f, c = 0.f, d = 0.f;
            cudaEventElapsedTime(&a, tl[0].a, tl[i].a);
            cudaEventElapsedTime(&c, tl[0].a, tl[i].c);
            cudaEventElapsedTime(&d, tl[0].a, tl[i].d);
            std::fprintf(stderr, "[bpc host]   piece %2zu n=%4d  stream start %7.3f  computed %7.3f  d2h done %7.3f ms\n", i, tl[i].n, a, c, d);
        }
        for (auto& e : tl) { cudaEventDestroy(e.a); cudaEventDestroy(e.c); cudaEventDestroy(e.d); }
        cudaGetLastError();
    }
    if (env_trace) {
        int ncpu = -1;
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) ncpu = CPU_COUNT(&set);
#endif
        std::fprintf(stderr, "[bpc host] B=%lld pieces=%lld layout=%s total %.2f ms, waiting on GPU/PCIe %.2f ms, host copy/fill %.2f ms "
                     "(%d threads, %d cpus in the affinity mask, GPU numa node %d with %d local cpus, placement %s)\n",
                     (long long)B, (long long)nchunks, to_rows ? (h->contig_d2h ? "compact/contiguous" : "compact/2d") : "full",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(), h->t_wait,
                     h->t_fill, h->pool->size(), ncpu, h->numa.node, (int)h->numa.cpus.size(), h->numa_place ? "on" : "off");
    }
    return BPC_OK;
}

}  // namespace

extern "C" {

int bpc_precompute_host(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, float* feats,
                        float* scalars, int32_t* status) {
    if (!h) return BPC_ERR_ARG;
    if (!wav || !feats || !scalars || B < 0 || L_in <= 0 || (wav_dtype != BPC_WAV_F32 && wav_dtype != BPC_WAV_PCM16)) {
        h->err = "bpc_precompute_host: bad argument";
        return BPC_ERR_ARG;
    }
    return host_pipeline(h, wav, wav_dtype, B, L_in, HostOut{feats, nullptr, nullptr}, scalars, status);
}

int bpc_precompute_host_compact(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, float* rows,
                                float* pad, float* scalars, int32_t* status) {
    if (!h) return BPC_ERR_ARG;
    if (!wav || !rows || !pad || !scalars || B < 0 || L_in <= 0 ||
        (wav_dtype != BPC_WAV_F32 && wav_dtype != BPC_WAV_PCM16)) {
        h->err = "bpc_precompute_host_compact: bad argument";
        return BPC_ERR_ARG;
    }
    return host_pipeline(h, wav, wav_dtype, B, L_in, HostOut{nullptr, rows, pad}, scalars, status);
}

int bpc_precompute_host_compact_begin(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, float* rows,
                                      float* pad, float* scalars, int32_t* status, int64_t* ticket) {
    if (!h) return BPC_ERR_ARG;
    if (!wav || !rows || !pad || !scalars || !ticket || B < 0 || L_in <= 0 ||
        (wav_dtype != BPC_WAV_F32 && wav_dtype != BPC_WAV_PCM16)) {
        h->err = "bpc_precompute_host_compact_begin: bad argument";
        return BPC_ERR_ARG;
    }
    return host_begin(h, wav, wav_dtype, B, L_in, HostOut{nullptr, rows, pad}, scalars, status, ticket, nullptr, nullptr);
}

int bpc_host_wait(bpc_handle* h, int64_t ticket) {
    if (!h) return BPC_ERR_ARG;
    if (!h->slots_ready) return BPC_OK;                     // nothing was ever enqueued
    return host_wait(h, ticket);
}

int bpc_live_rows(int channel) { return (channel >= 0 && channel < 9) ? kLiveRows[channel] : BPC_ERR_ARG; }

int bpc_expand_compact(const float* rows, const float* pad, int64_t n, int T, float* feats, int n_threads) {
    if (!rows || !pad || !feats || n < 0 || T <= 0) return BPC_ERR_ARG;
    const size_t seg_feats = (size_t)9 * kPlaneRows * T, seg_rows = (size_t)kLiveTotal * T;
    int nt = std::max(1, std::min(n_threads, 64));
    if ((int64_t)nt > n) nt = (int)std::max<int64_t>(1, n);
    auto work = [&](int part) {
        for (int64_t b = part; b < n; b += nt) {
            const float* src = rows + (size_t)b * seg_rows;
            float* dst = feats + (size_t)b * seg_feats;
            for (int c = 0; c < 9; ++c) {
                const size_t live = (size_t)kLiveRows[c] * T;
                std::memcpy(dst, src, live * 4);
                const float v = pad[(size_t)b * 9 + c];
                for (float* q = dst + live; q < dst + (size_t)kPlaneRows * T; ++q) *q = v;
                src += live;
                dst += (size_t)kPlaneRows * T;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    return BPC_OK;
}

void* bpc_host_alloc(bpc_handle* h, int64_t bytes, int* numa_node) {
    if (!h || bytes <= 0) return nullptr;
    if (cudaSetDevice(h->device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    void* p = nullptr;
    if (host_pinned(h, (size_t)bytes, &p)) return nullptr;
    if (numa_node) *numa_node = h->numa_place ? h->numa.node : -1;
    return p;
}

void bpc_host_free(bpc_handle* h, void* p) {
    if (!h || !p) return;
    for (auto it = h->numa_allocs.begin(); it != h->numa_allocs.end(); ++it)
        if (it->first == p) {
            cudaSetDevice(h->device);
            numa_pinned_free(it->first, it->second);
            h->numa_allocs.erase(it);
            return;
        }
}

int bpc_stage_logmel(bpc_handle* h, const void* wav, int wav_dtype, int64_t B, int64_t L_in, float* stft_db,
                     float* mel3, void* stream) {
    if (!h) return BPC_ERR_ARG;
    if (!wav || !mel3 || B < 0 || L_in <= 0) { h->err = "bpc_stage_logmel: bad argument"; return BPC_ERR_ARG; }
    BPC_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Geometry& g = h->g;
    const Workspace& ws = h->debug ? h->main.ws_dbg : h->main.ws;
    const size_t esz = wav_dtype == BPC_WAV_F32 ? 4 : 2;
    for (int64_t off = 0; off < B; off += h->chunk) {
        const int n = (int)std::min<int64_t>(h->chunk, B - off);
        const void* w = static_cast<const char*>(wav) + (size_t)off * L_in * esz;
        if (L_in == g.L && !h->debug &&
            launch_logmel_fused(w, wav_dtype, n, g, h->tb, stft_db ? stft_db + (size_t)off * 257 * g.T : nullptr,
                                mel3 + (size_t)off * 3 * kPlaneRows * g.T, st)) {
            h->last_n = n;
            continue;
        }
        const float* y;
        if (wav_dtype == BPC_WAV_F32 && L_in == g.L && (reinterpret_cast<uintptr_t>(w) & 15) == 0) y = static_cast<const float*>(w);
        else { launch_ingest(w, wav_dtype, L_in, ws.y, n, g, st); y = ws.y; }
        launch_stft512(y, n, g, h->tb, ws, st);
        if (stft_db) launch_stft_db(n, g, ws, stft_db + (size_t)off * 257 * g.T, st);
        launch_logmel_only(n, g, h->tb, ws, mel3 + (size_t)off * 3 * kPlaneRows * g.T, st);
        h->last_n = n;
    }
    BPC_CUDA(h, cudaGetLastError());
    return BPC_OK;
}

int bpc_modspec(bpc_handle* h, const float* mel_db, int64_t n, float* out, void* stream) {
    if (!h) return BPC_ERR_ARG;
    if (!mel_db || !out || n < 0) { h->err = "bpc_modspec: bad argument"; return BPC_ERR_ARG; }
    BPC_CUDA(h, cudaSetDevice(h->device));
    if (n > 0) launch_modspec((int)n, h->g, h->tb, h->main.ws, mel_db, out, static_cast<cudaStream_t>(stream));
    BPC_CUDA(h, cudaGetLastError());
    return BPC_OK;
}

int bpc_collate(bpc_handle* h, const float* store_feats, const float* store_scalars, int64_t N, const int64_t* idx_a,
                const int64_t* idx_b, int64_t n, int mode, double lam, int y1, int y2, int x1, int x2,
                float* out_feats, float* out_scalars, void* stream) {
    if (!h) return BPC_ERR_ARG;
    const Geometry& g = h->g;
    if (!store_feats || !idx_a || !out_feats || N <= 0 || n < 0 || n > 65535 || mode < BPC_MIX_NONE ||
        mode > BPC_MIX_CUTMIX || (mode != BPC_MIX_NONE && !idx_b) || (out_scalars && !store_scalars)) {
        h->err = "bpc_collate: bad argument (n <= 65535 per call)";
        return BPC_ERR_ARG;
    }
    if (mode == BPC_MIX_CUTMIX && (y1 < 0 || y2 > kPlaneRows || x1 < 0 || x2 > g.T)) {
        h->err = "bpc_collate: cut box outside the [128, T] plane";
        return BPC_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(store_feats) | reinterpret_cast<uintptr_t>(out_feats)) & 15) {
        h->err = "bpc_collate: feature buffers must be 16-byte aligned";
        return BPC_ERR_ARG;
    }
    BPC_CUDA(h, cudaSetDevice(h->device));
    // lam stays a double until here, like the Python float of the reference; torch multiplies a float32 tensor by it
    // after rounding it to float32, and (1 - lam) is formed in double first
    if (n > 0)
        launch_collate(store_feats, store_scalars, reinterpret_cast<const long long*>(idx_a),
                       reinterpret_cast<const long long*>(idx_b), (int)n, mode, (float)lam, (float)(1.0 - lam), y1, y2,
                       x1, x2, g.T, g.nscal, out_feats, out_scalars, static_cast<cudaStream_t>(stream));
    BPC_CUDA(h, cudaGetLastError());
    return BPC_OK;
}

int bpc_channel_stats(bpc_handle* h, double* stats_host) {
    if (!h || !stats_host) return BPC_ERR_ARG;
    BPC_CUDA(h, cudaSetDevice(h->device));
    BPC_CUDA(h, cudaDeviceSynchronize());
    BPC_CUDA(h, cudaMemcpy(stats_host, h->stats_acc, sizeof(double) * 5 * (9 + h->g.nscal), cudaMemcpyDeviceToHost));
    return BPC_OK;
}

int bpc_channel_stats_device(bpc_handle* h, double** stats_dev, int64_t* n_rows) {
    if (!h || !stats_dev) return BPC_ERR_ARG;
    *stats_dev = h->stats_acc;
    if (n_rows) *n_rows = 9 + h->g.nscal;
    return BPC_OK;
}

int bpc_channel_stats_reset(bpc_handle* h) {
    if (!h) return BPC_ERR_ARG;
    BPC_CUDA(h, cudaSetDevice(h->device));
    BPC_CUDA(h, cudaDeviceSynchronize());
    return reset_stats(h, 0);
}

int bpc_debug_copy(bpc_handle* h, const char* what, void* out, int64_t cap_bytes, int64_t* got) {
    if (!h || !what || !out || !got) return BPC_ERR_ARG;
    BPC_CUDA(h, cudaSetDevice(h->device));
    BPC_CUDA(h, cudaDeviceSynchronize());
    const Geometry& g = h->g;
    const Workspace& w = h->main.ws_dbg;
    const size_t n = (size_t)h->last_n, T = (size_t)g.T;
    const void* src = nullptr;
    size_t bytes = 0;
    const std::string k(what);
    if (k == "mag512") { src = w.mag512; bytes = n * T * kMagStride * 4; }
    else if (k == "mel_db") { src = w.dbg_mel_db; bytes = n * 128 * T * 4; }
    else if (k == "mfcc_raw") { src = w.dbg_mfcc; bytes = n * 120 * T * 4; }
    else if (k == "gammatone_raw") { src = w.dbg_gam; bytes = n * 64 * T * 4; }
    else if (k == "mod_spec_raw") { src = w.dbg_mod; bytes = n * 40 * T * 4; }
    else if (k == "chroma_stft_raw") { src = w.dbg_chroma_stft; bytes = n * 12 * T * 4; }
    else if (k == "chroma_cens_raw") { src = w.dbg_chroma_cens; bytes = n * 12 * T * 4; }
    else if (k == "lpc_raw") { src = w.dbg_lpc; bytes = n * 12 * (size_t)g.lpc_frames * 4; }
    else if (k == "onset_env") { src = w.dbg_onset; bytes = n * T * 4; }
    else if (k == "tuning") { src = w.tuning; bytes = n * 2 * 4; }
    else if (k == "ints") { src = w.ints; bytes = n * 2 * 4; }
    else { h->err = "bpc_debug_copy: unknown key " + k; return BPC_ERR_ARG; }
    if (!src) { h->err = "bpc_debug_copy: debug buffers are off (bpc_set_debug)"; return BPC_ERR_ARG; }
    bytes = std::min<size_t>(bytes, (size_t)cap_bytes);
    BPC_CUDA(h, cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    *got = (int64_t)bytes;
    return BPC_OK;
}

int64_t bpc_table_copy(const bpc_params* p, const char* name, int tuning_idx, void* out, int64_t cap_elems) {
    std::string why;
    if (check_params(p, &why) || !name || !out) return BPC_ERR_ARG;
    const std::string k(name);
    auto put_f = [&](const std::vector<float>& v) -> int64_t {
        if ((int64_t)v.size() > cap_elems) return BPC_ERR_ARG;
        std::memcpy(out, v.data(), v.size() * sizeof(float));
        return (int64_t)v.size();
    };
    auto put_d = [&](const std::vector<double>& v) -> int64_t {
        if ((int64_t)v.size() > cap_elems) return BPC_ERR_ARG;
        std::memcpy(out, v.data(), v.size() * sizeof(double));
        return (int64_t)v.size();
    };
    const int T = p->expected_len / p->hop + 1;
    std::vector<double> edges = tuning_edges();
    if (tuning_idx < 0 || tuning_idx >= kNumTunings) tuning_idx = 50;
    if (k == "mel_a") return put_f(mel_bank(p->sr, 512, 128, 0.0, p->fmax).dense);
    if (k == "mel_b") return put_f(mel_bank(p->sr, 512, 128, 0.0, p->sr / 2.0).dense);
    if (k == "mel_c") return put_f(mel_bank(p->sr, 512, 64, 0.0, p->sr / 2.0).dense);
    if (k == "mel_d") return put_f(mel_bank(p->sr, 2048, 128, 0.0, p->sr / 2.0).dense);
    if (k == "mel_d_band") {                                   // the band form k_frame2048 reads: [128, 2 + 80] = start, count, weights
        const SparseBank md = deconflict_bank(mel_bank(p->sr, 2048, 128, 0.0, p->sr / 2.0), 32);
        if (md.width > 80) return BPC_ERR_UNSUPPORTED;
        std::vector<float> v((size_t)md.rows * 82, 0.f);
        for (int r = 0; r < md.rows; ++r) {
            v[(size_t)r * 82] = (float)md.start[r];
            v[(size_t)r * 82 + 1] = (float)md.count[r];
            for (int j = 0; j < md.width; ++j) v[(size_t)r * 82 + 2 + j] = md.w[(size_t)r * md.width + j];
        }
        return put_f(v);
    }
    if (k == "dct_mel") return put_f(dct2_ortho(40, 128));
    if (k == "dct_time") return put_f(dct2_ortho(T, T));
    if (k == "hann512") return put_d(hann_periodic(512));
    if (k == "hann2048") return put_d(hann_periodic(2048));
    if (k == "hann384") return put_d(hann_periodic(384));
    if (k == "hamming400") return put_d(hamming_sym(400));
    if (k == "chroma") return put_f(chroma_bank(p->sr, 512, edges[tuning_idx]));
    if (k == "hist_edges") return put_d(edges);
    if (k == "halfband") return put_d(halfband_taps());
    if (k == "cqt_basis" || k == "cqt_sqrt_len") {
        std::vector<std::complex<float>> dense;
        CqtBasisEll e = cqt_basis(p->sr, edges[tuning_idx], &dense);
        if (e.col.empty()) return BPC_ERR_UNSUPPORTED;
        if (k == "cqt_sqrt_len") return put_d(e.sqrt_len);
        std::vector<float> flat(dense.size() * 2);
        for (size_t i = 0; i < dense.size(); ++i) { flat[2 * i] = dense[i].real(); flat[2 * i + 1] = dense[i].imag(); }
        return put_f(flat);
    }
    return BPC_ERR_ARG;
}

int64_t bpc_resample_len(int64_t n_in, int sr_in, int sr_out) {
    if (n_in < 0 || sr_in <= 0 || sr_out <= 0) return BPC_ERR_ARG;
    return (n_in * (int64_t)sr_out + sr_in - 1) / sr_in;
}

int64_t bpc_resample_filter(int sr_in, int sr_out, double* out, int64_t cap, int* p, int* q, int* half) {
    if (sr_in <= 0 || sr_out <= 0 || !out) return BPC_ERR_ARG;
    const ResampleFilter f = resample_filter(sr_in, sr_out);
    if ((int64_t)f.tab.size() > cap) return BPC_ERR_ARG;
    std::memcpy(out, f.tab.data(), f.tab.size() * sizeof(double));
    if (p) *p = f.p;
    if (q) *q = f.q;
    if (half) *half = f.half;
    return (int64_t)f.tab.size();
}

int bpc_resample(bpc_handle* h, const float* in, int64_t n_in, int sr_in, int sr_out, float* out, int64_t out_cap,
                 void* stream) {
    if (!h) return BPC_ERR_ARG;
    const int64_t n_out = bpc_resample_len(n_in, sr_in, sr_out);
    if (n_out == 0) return BPC_OK;                                    // empty waveform: nothing to write
    if (!in || !out || n_out < 0 || out_cap < n_out || sr_in > 768000 || sr_out > 768000) {
        h->err = "bpc_resample: bad argument";
        return BPC_ERR_ARG;
    }
    BPC_CUDA(h, cudaSetDevice(h->device));
    const bpc_handle::Resampler* r = nullptr;
    for (const auto& e : h->resamplers) if (e.sr_in == sr_in && e.sr_out == sr_out) r = &e;
    if (!r) {
        const ResampleFilter f = resample_filter(sr_in, sr_out);
        if (f.tab.size() > (size_t)64 << 20) { h->err = "bpc_resample: rate pair needs too large a polyphase table"; return BPC_ERR_UNSUPPORTED; }
        const double* d = nullptr;
        int rc = upload(h, f.tab, &d);
        if (rc) return rc;
        h->resamplers.push_back({sr_in, sr_out, f.p, f.q, f.half, d});
        r = &h->resamplers.back();
    }
    launch_resample(in, n_in, r->tab, r->p, r->q, r->half, out, n_out, static_cast<cudaStream_t>(stream));
    BPC_CUDA(h, cudaGetLastError());
    return BPC_OK;
}

int bpc_wav_decode(bpc_handle* h, const void* blob, int64_t blob_bytes, const int64_t* file_offset,
                   const bpc_wav_info* info, int64_t n, int64_t L, float* y, void* stream) {
    if (!h) return BPC_ERR_ARG;
    if (n == 0) return BPC_OK;
    if (!blob || !file_offset || !info || !y || n < 0 || n > 65535 || L <= 0 || L > 0x7fffffff) {
        h->err = "bpc_wav_decode: bad argument (at most 65535 files per call)";
        return BPC_ERR_ARG;
    }
    std::vector<WavItem> items((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const bpc_wav_info& f = info[i];
        if (f.fmt < BPC_FMT_U8 || f.fmt > BPC_FMT_F64 || f.channels < 1 || f.channels > 7 || f.frames < 0 || f.data_offset < 0) {
            h->err = "bpc_wav_decode: info[" + std::to_string(i) + "] does not come from a successful bpc_wav_parse";
            return BPC_ERR_ARG;
        }
        static const int width[7] = {0, 1, 2, 3, 4, 4, 8};
        const int64_t first = file_offset[i] + f.data_offset, bytes = f.frames * f.channels * width[f.fmt];
        if (file_offset[i] < 0 || first < 0 || first > blob_bytes || bytes > blob_bytes - first) {
            h->err = "bpc_wav_decode: the samples of file " + std::to_string(i) + " do not lie inside the blob";
            return BPC_ERR_ARG;
        }
        items[(size_t)i] = {(long long)first, (long long)f.frames, f.channels, f.fmt};
    }
    BPC_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->wav_items_cap < n) {                                    // grows; freed with the handle
        WavItem* d = nullptr;
        BPC_CUDA(h, cudaMalloc((void**)&d, sizeof(WavItem) * (size_t)n));
        h->dev_allocs.push_back(d);
        h->wav_items = d;
        h->wav_items_cap = n;
    }
    // pageable source: the runtime stages it before returning, `items` may die with this call
    BPC_CUDA(h, cudaMemcpyAsync(h->wav_items, items.data(), sizeof(WavItem) * (size_t)n, cudaMemcpyHostToDevice, st));
    launch_wav_decode(static_cast<const unsigned char*>(blob), (long long)blob_bytes, h->wav_items, (int)n, (int)L, y, st);
    BPC_CUDA(h, cudaGetLastError());
    return BPC_OK;
}

int64_t bpc_launch_count(const bpc_handle* h) { return h ? launches_issued() - h->launches0 : 0; }

int bpc_chunk_size(const bpc_handle* h) { return h ? h->chunk : BPC_ERR_ARG; }

int bpc_set_kernel_timing(bpc_handle* h, int on) {
    if (!h) return BPC_ERR_ARG;
    h->timing = on != 0;
    return BPC_OK;
}

int bpc_kernel_times(bpc_handle* h, double* ms_out, int64_t* launches_out, int n_ids) {
    if (!h || !ms_out || !launches_out || n_ids < BPC_NUM_KERNEL_IDS) return BPC_ERR_ARG;
    BPC_CUDA(h, cudaSetDevice(h->device));
    BPC_CUDA(h, cudaDeviceSynchronize());
    for (int i = 0; i < n_ids; ++i) { ms_out[i] = 0.0; launches_out[i] = 0; }
    for (auto& e : h->evs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        if (e.id >= 0 && e.id < n_ids) { ms_out[e.id] += (double)ms; launches_out[e.id] += 1; }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    h->evs.clear();
    return BPC_OK;
}

const char* bpc_kernel_name(int id) {
    static const char* names[BPC_NUM_KERNEL_IDS] = {"k_ingest", "k_stft512", "k_spec512_consumers",
                                                    "k_frame2048", "k_even2048", "k_cens",
                                                    "k_time_basic+k_autocorr", "k_hilbert", "k_lpc", "k_stats",
                                                    "k_seg2048", "k_cens_dec", "k_cens_lo"};
    return (id >= 0 && id < BPC_NUM_KERNEL_IDS) ? names[id] : "";
}

}  // extern "C"
