// Batch assembly on a device-resident feature store: the gather that `collate_fn` does on the host
// (dataset.py:59-73) fused with the CutMix / MixUp augmentations of augmentation.py:5-44 and train.py:76-89.
// Pure HBM traffic: one float4 stream in (two for a mix), one out; 290,304 B per segment and stream.
#include "kernels.cuh"

namespace bpc {

// mode 0: out[i] = store[ia[i]]
// mode 1: out[i] = lam * store[ia[i]] + oml * store[ib[i]]     (torch: two float32 multiplies, one add; no FMA)
// mode 2: out[i] = store[ib[i]] inside rows [y1,y2) x columns [x1,x2) of every plane, store[ia[i]] elsewhere
__global__ void __launch_bounds__(256) k_collate(const float4* __restrict__ store, const long long* __restrict__ ia,
                                                 const long long* __restrict__ ib, int mode, float lam, float oml,
                                                 int y1, int y2, int x1, int x2, int T, int vec_per_seg,
                                                 float4* __restrict__ out) {
    const int i = blockIdx.y;
    const float4* a = store + (size_t)ia[i] * vec_per_seg;
    const float4* b = mode ? store + (size_t)ib[i] * vec_per_seg : a;
    float4* o = out + (size_t)i * vec_per_seg;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vec_per_seg; v += gridDim.x * blockDim.x) {
        float4 r = __ldcs(a + v);                       // streamed once: keep it out of the way of L2-resident data
        if (mode == 1) {
            const float4 s = __ldcs(b + v);
            r.x = __fadd_rn(__fmul_rn(lam, r.x), __fmul_rn(oml, s.x));
            r.y = __fadd_rn(__fmul_rn(lam, r.y), __fmul_rn(oml, s.y));
            r.z = __fadd_rn(__fmul_rn(lam, r.z), __fmul_rn(oml, s.z));
            r.w = __fadd_rn(__fmul_rn(lam, r.w), __fmul_rn(oml, s.w));
        } else if (mode == 2) {
            // element e = 4 v + j of the segment: column e % T, row (e / T) % 128
            const int e0 = 4 * v, row0 = e0 / T;
            int w = e0 - row0 * T, h = row0 & (kPlaneRows - 1);
            bool in[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                in[j] = h >= y1 && h < y2 && w >= x1 && w < x2;
                if (++w == T) { w = 0; h = (h + 1) & (kPlaneRows - 1); }
            }
            if (in[0] | in[1] | in[2] | in[3]) {
                const float4 s = __ldcs(b + v);
                if (in[0]) r.x = s.x;
                if (in[1]) r.y = s.y;
                if (in[2]) r.z = s.z;
                if (in[3]) r.w = s.w;
            }
        }
        __stcs(o + v, r);
    }
}

// scalars: [n, S] (S not a multiple of 4 in general): one thread per element
__global__ void k_collate_scalars(const float* __restrict__ store, const long long* __restrict__ ia,
                                  const long long* __restrict__ ib, int mix, float lam, float oml, int S, long long total,
                                  float* __restrict__ out) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / S;
        const int j = (int)(e - i * S);
        float r = store[(size_t)ia[i] * S + j];
        if (mix) r = __fadd_rn(__fmul_rn(lam, r), __fmul_rn(oml, store[(size_t)ib[i] * S + j]));
        out[e] = r;
    }
}

void launch_collate(const float* store_feats, const float* store_scalars, const long long* ia, const long long* ib,
                    int n, int mode, float lam, float oml, int y1, int y2, int x1, int x2, int T, int nscal,
                    float* out_feats, float* out_scalars, cudaStream_t st) {
    const int vec = 9 * kPlaneRows * T / 4;             // 128 rows: always a multiple of 4 floats
    // 148 SMs x 8 resident 256-thread CTAs; x-blocks per segment so that the grid is a few waves at typical batches
    int bx = (148 * 8 * 4 + n - 1) / n;
    const int bx_max = (vec + 255) / 256;
    if (bx > bx_max) bx = bx_max;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n);
    k_collate<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(store_feats), ia, ib, mode, lam, oml, y1, y2, x1, x2,
                                    T, vec, reinterpret_cast<float4*>(out_feats));
    note_launch();
    if (out_scalars) {
        const long long total = (long long)n * nscal;
        const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        k_collate_scalars<<<blocks, 256, 0, st>>>(store_scalars, ia, ib, mode == 1, lam, oml, nscal, total, out_scalars);
        note_launch();
    }
}

// Pad values of a chunk: fill[b * 9 + c] = first element of the first pad row of plane c (pad_freq fills rows
// live[c]..127 with one constant per plane, methods.py:39-46); planes without pad rows get 0.  Used by the host path,
// which transfers only the live rows over PCIe and re-creates the constant rows on the host.
__global__ void k_pad_values(const float* __restrict__ feats, int T, int n, const int* __restrict__ live,
                             float* __restrict__ fill) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 9) return;
    const int b = i / 9, c = i - b * 9, lv = live[c];
    fill[i] = lv < kPlaneRows ? feats[((size_t)b * 9 + c) * kPlaneRows * T + (size_t)lv * T] : 0.f;
}

void launch_pad_values(const float* feats, int T, int n, const int* live_dev, float* fill, cudaStream_t st) {
    k_pad_values<<<(n * 9 + 255) / 256, 256, 0, st>>>(feats, T, n, live_dev, fill);
    note_launch();
}

// Compact copy of a chunk for the host path: out[b] = the 772 data rows of segment b back to back (chroma 24, gammatone
// 64, lpc 12, mel / mel_delta / mel_delta2 128 each, mfcc 120, mod_spec 40, tempogram 128 -- every count a multiple of
// four rows, so float4 granules never straddle a plane), so that ONE contiguous device->host copy per piece carries
// exactly the bytes that have to cross PCIe.  772 T * 4 B read + written per segment (0.39 MB at T = 63).
struct LiveLayout { int start4[10]; };          // prefix sums of live rows * T / 4 (float4 granules), 9 planes + total

__global__ void __launch_bounds__(256) k_compact_rows(const float4* __restrict__ feats, int plane4, LiveLayout lay,
                                                      float4* __restrict__ out) {
    const int b = blockIdx.y, total4 = lay.start4[9];
    const float4* src = feats + (size_t)b * 9 * plane4;
    float4* dst = out + (size_t)b * total4;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total4; v += gridDim.x * blockDim.x) {
        int c = 0;
#pragma unroll
        for (int q = 1; q < 9; ++q) c += v >= lay.start4[q];
        __stcs(dst + v, __ldcs(src + (size_t)c * plane4 + (v - lay.start4[c])));
    }
}

void launch_compact_rows(const float* feats, int T, int n, const int* live_host, float* out, cudaStream_t st) {
    LiveLayout lay;
    lay.start4[0] = 0;
    for (int c = 0; c < 9; ++c) lay.start4[c + 1] = lay.start4[c] + live_host[c] * T / 4;
    int bx = (148 * 8 * 2 + n - 1) / n;
    const int bx_max = (lay.start4[9] + 255) / 256;
    if (bx > bx_max) bx = bx_max;
    if (bx < 1) bx = 1;
    k_compact_rows<<<dim3(bx, n), 256, 0, st>>>(reinterpret_cast<const float4*>(feats), kPlaneRows * T / 4, lay,
                                               reinterpret_cast<float4*>(out));
    note_launch();
}

}  // namespace bpc
