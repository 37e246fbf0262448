// GPU-side wav decode: soundfile's sample scaling, librosa.to_mono and pad_or_truncate
// (process.py:28-29, methods.py:24-28) for a batch of RIFF/WAVE file images lying in device memory as they were read
// from disk.  One thread per output sample; sample words are assembled from bytes (the payload of a file starts at an
// arbitrary byte offset of the blob).  HBM-bound byte work: reads the payload once, writes [n, L] float32 once.
#include "kernels.cuh"

namespace bpc {

__device__ __forceinline__ float wav_sample(const unsigned char* __restrict__ p, int fmt) {
    switch (fmt) {
        case BPC_FMT_U8: return ((float)p[0] - 128.0f) * (1.0f / 128.0f);
        case BPC_FMT_PCM16: return (float)(short)(p[0] | (p[1] << 8)) * (1.0f / 32768.0f);
        case BPC_FMT_PCM24: {                                    // 24 bits fit a float32 mantissa: exact
            const int v = (int)((unsigned)p[0] << 8 | (unsigned)p[1] << 16 | (unsigned)p[2] << 24) >> 8;
            return (float)v * (1.0f / 8388608.0f);
        }
        case BPC_FMT_PCM32: {                                    // float32(float64(v) / 2^31): one rounding, then a power of two
            const int v = (int)((unsigned)p[0] | (unsigned)p[1] << 8 | (unsigned)p[2] << 16 | (unsigned)p[3] << 24);
            return __int2float_rn(v) * (1.0f / 2147483648.0f);
        }
        case BPC_FMT_F32: {
            const unsigned u = (unsigned)p[0] | (unsigned)p[1] << 8 | (unsigned)p[2] << 16 | (unsigned)p[3] << 24;
            return __uint_as_float(u);
        }
        default: {                                               // BPC_FMT_F64
            unsigned long long u = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) u |= (unsigned long long)p[i] << (8 * i);
            return __double2float_rn(__longlong_as_double((long long)u));
        }
    }
}

__global__ void __launch_bounds__(256) k_wav_decode(const unsigned char* __restrict__ blob, long long blob_bytes,
                                                    const WavItem* __restrict__ items, int L, float* __restrict__ y) {
    const int i = blockIdx.y;
    const WavItem it = items[i];
    const int bytes = it.fmt == BPC_FMT_U8 ? 1 : it.fmt == BPC_FMT_PCM16 ? 2 : it.fmt == BPC_FMT_PCM24 ? 3
                      : it.fmt == BPC_FMT_F64 ? 8 : 4;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (t < it.frames) {
            const unsigned char* p = blob + it.offset + (size_t)t * it.channels * bytes;
            BPC_ASSERT(it.offset + ((long long)t + 1) * it.channels * bytes <= blob_bytes);
            float acc = wav_sample(p, it.fmt);
            // librosa.to_mono = np.mean(axis = channels) in float32: sequential sum (fewer than 8 addends), then / count
            for (int c = 1; c < it.channels; ++c) acc = __fadd_rn(acc, wav_sample(p + c * bytes, it.fmt));
            v = it.channels > 1 ? __fdiv_rn(acc, (float)it.channels) : acc;
        }
        y[(size_t)i * L + t] = v;
    }
}

void launch_wav_decode(const unsigned char* blob, long long blob_bytes, const WavItem* items, int n, int L, float* y,
                       cudaStream_t st) {
    const int bx = (L + 255) / 256 < 64 ? (L + 255) / 256 : 64;
    k_wav_decode<<<dim3(bx, n), 256, 0, st>>>(blob, blob_bytes, items, L, y);
    note_launch();
}

}  // namespace bpc
