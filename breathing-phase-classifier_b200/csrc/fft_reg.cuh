// Register-resident FP64 FFTs for sm_100a: every lane keeps R complex points in registers, does a radix-R DFT on them
// with compile-time twiddles, and the lanes of a team exchange data ONCE through shared memory (four-step FFT:
// N = R x R).  Compared with the shared-memory radix-4 passes of fft.cuh this moves 2 * 16 B per point through the
// shared-memory pipe instead of 2 * 16 B per point per pass (5 passes for N = 1024) and has no per-butterfly index
// arithmetic -- the r01 v1 profile showed those kernels bound by shared-memory wavefronts / issue slots with the FP64
// pipe 10-20 % busy.
//
//   team_fft<16>  N = 256   16 lanes x 16 points  (two independent transforms per warp)   real FFT-512  (STFT-512, CQT)
//   team_fft<32>  N = 1024  32 lanes x 32 points  (one transform per warp)                real FFT-2048 (STFT-2048, autocorr)
//
// Everything is __host__ __device__ and lane-explicit so tests/fft_host_test.cpp can run the exact index logic on the
// CPU (there is no GPU in the build container).
#pragma once
#include <cuda_runtime.h>

namespace bpc {

#ifndef BPC_HD
#define BPC_HD __host__ __device__ __forceinline__
#endif

// cos(2 pi i / 64), i = 0..16 (quarter wave); everything else by symmetry.
BPC_HD constexpr double cos64_q(int i) {
    switch (i) {
        case 0: return 1.0;
        case 1: return 0.99518472667219688624;
        case 2: return 0.98078528040323044913;
        case 3: return 0.95694033573220886494;
        case 4: return 0.92387953251128675613;
        case 5: return 0.88192126434835502971;
        case 6: return 0.83146961230254523708;
        case 7: return 0.77301045336273696081;
        case 8: return 0.70710678118654752440;
        case 9: return 0.63439328416364549822;
        case 10: return 0.55557023301960222474;
        case 11: return 0.47139673682599764856;
        case 12: return 0.38268343236508977173;
        case 13: return 0.29028467725446236764;
        case 14: return 0.19509032201612826785;
        case 15: return 0.09801714032956060199;
        default: return 0.0;
    }
}
// cos / sin of 2 pi i / 64 for any integer i >= 0
BPC_HD constexpr double cos64(int i) {
    i &= 63;
    return i <= 16 ? cos64_q(i) : (i <= 32 ? -cos64_q(32 - i) : (i <= 48 ? -cos64_q(i - 32) : cos64_q(64 - i)));
}
BPC_HD constexpr double sin64(int i) { return cos64(i + 48); }   // sin(x) = cos(x - pi/2) = cos(x + 3 pi / 2)

BPC_HD double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
BPC_HD double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
BPC_HD double2 c_mul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

// d * exp(-2 pi i * I / R) with a compile-time twiddle
template <int R, int I>
BPC_HD double2 mul_tw(double2 d) {
    constexpr int q = (I * (64 / R)) & 63;                  // position on the 64-point circle
    if (q == 0) return d;
    if (q == 16) return make_double2(d.y, -d.x);            // -i
    if (q == 32) return make_double2(-d.x, -d.y);
    if (q == 48) return make_double2(-d.y, d.x);            // +i
    constexpr double c = cos64(q), s = -sin64(q);           // exp(-i th) = c + i s
    return make_double2(fma(d.x, c, -(d.y * s)), fma(d.x, s, d.y * c));
}

// In-place radix-2 decimation-in-frequency DFT of R register-resident points a[OFF .. OFF + R).
// On return a[OFF + bitrev_R(k)] = X[k].
template <int R, int OFF>
struct DifR {
    template <int I>
    static BPC_HD void bfly(double2* a) {
        const double2 u = a[OFF + I], v = a[OFF + I + R / 2];
        a[OFF + I] = c_add(u, v);
        a[OFF + I + R / 2] = mul_tw<R, I>(c_sub(u, v));
    }
    template <int I>
    static BPC_HD void loop(double2* a) {
        if constexpr (I < R / 2) {
            bfly<I>(a);
            loop<I + 1>(a);
        }
    }
    static BPC_HD void run(double2* a) {
        loop<0>(a);
        DifR<R / 2, OFF>::run(a);
        DifR<R / 2, OFF + R / 2>::run(a);
    }
};
template <int OFF>
struct DifR<1, OFF> {
    static BPC_HD void run(double2*) {}
};

template <int R>
BPC_HD constexpr int bitrev(int k) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (k & 1); k >>= 1; }
    return r;
}

// ------------------------------------------------------------------------------------ r02: decimation in time, FMA form
// Same contract as DifR (natural-order input, a[OFF + S * bitrev_R(k)] = X[k] on return), but the twiddle is applied
// BEFORE the butterfly, y = u +/- w v, and factored as w = c (1 + i t) (or c (t + i) when |Im w| > |Re w|) with
// compile-time c and t = tan / cot of the angle (Linzer-Feig): p = v (1 + i t) is two FMAs, u +/- c p four more --
// six FP64 instructions per butterfly instead of the eight of (u - v) w (4 DADD + 2 DMUL + 2 DFMA), all of them FMAs.
// Radix 32: 388 instead of 456 FP64 instructions per lane and transform.  |t| <= 1, so nothing is amplified.
template <int R, int OFF, int S>
struct DitR {
    template <int K>
    static BPC_HD void bfly(double2* a) {
        constexpr int pe = OFF + 2 * S * bitrev<R / 2>(K), po = pe + S;
        constexpr int q = (K * (64 / R)) & 63;              // w = exp(-2 pi i q / 64), q in [0, 32)
        const double2 u = a[pe], v = a[po];
        if constexpr (q == 0) {
            a[pe] = c_add(u, v);
            a[po] = c_sub(u, v);
        } else if constexpr (q == 16) {                     // w = -i: w v = (v.y, -v.x)
            a[pe] = make_double2(u.x + v.y, u.y - v.x);
            a[po] = make_double2(u.x - v.y, u.y + v.x);
        } else if constexpr (q == 8) {                      // w = c (1 - i): w v = c (v.x + v.y, v.y - v.x)
            constexpr double c = cos64(8);
            const double px = v.x + v.y, py = v.y - v.x;
            a[pe] = make_double2(fma(c, px, u.x), fma(c, py, u.y));
            a[po] = make_double2(fma(-c, px, u.x), fma(-c, py, u.y));
        } else if constexpr (q == 24) {                     // w = -c (1 + i): w v = -c (v.x - v.y, v.x + v.y)
            constexpr double c = cos64(8);
            const double px = v.x - v.y, py = v.x + v.y;
            a[pe] = make_double2(fma(-c, px, u.x), fma(-c, py, u.y));
            a[po] = make_double2(fma(c, px, u.x), fma(c, py, u.y));
        } else {
            constexpr double wr = cos64(q), wi = -sin64(q);
            if constexpr ((wr < 0 ? -wr : wr) >= (wi < 0 ? -wi : wi)) {
                constexpr double t = wi / wr;               // w = wr (1 + i t)
                const double px = fma(-t, v.y, v.x), py = fma(t, v.x, v.y);
                a[pe] = make_double2(fma(wr, px, u.x), fma(wr, py, u.y));
                a[po] = make_double2(fma(-wr, px, u.x), fma(-wr, py, u.y));
            } else {
                constexpr double t = wr / wi;               // w = wi (t + i)
                const double px = fma(t, v.x, -v.y), py = fma(t, v.y, v.x);
                a[pe] = make_double2(fma(wi, px, u.x), fma(wi, py, u.y));
                a[po] = make_double2(fma(-wi, px, u.x), fma(-wi, py, u.y));
            }
        }
    }
    template <int K>
    static BPC_HD void loop(double2* a) {
        if constexpr (K < R / 2) {
            bfly<K>(a);
            loop<K + 1>(a);
        }
    }
    static BPC_HD void run(double2* a) {
        DitR<R / 2, OFF, 2 * S>::run(a);
        DitR<R / 2, OFF + S, 2 * S>::run(a);
        loop<0>(a);
    }
};
template <int OFF, int S>
struct DitR<1, OFF, S> {
    static BPC_HD void run(double2*) {}
};

// The register DFT every team transform below uses (BPC_FFT_DIF=1 selects the r01 decimation-in-frequency form for A/B runs)
template <int R>
BPC_HD void reg_dft(double2* a) {
#ifdef BPC_FFT_DIF
    DifR<R, 0>::run(a);
#else
    DitR<R, 0, 1>::run(a);
#endif
}

// ------------------------------------------------------------------------------------------------ team FFT
// N = R * R points, a team of R lanes (lane id h in [0, R)), R points per lane.
//   in : a[j]            = x[h + R j]
//   out: a[bitrev_R(k2)] = X[h + R k2]
// xch: shared-memory exchange buffer of R * (R + 1) double2 per team (row pitch R + 1 keeps both the row-major stores
//      and the column reads free of bank conflicts);  tw: per-lane inter-stage twiddles, tw[k1] = exp(-2 pi i h k1 / N),
//      k1 = 1 .. R-1 (index 0 unused), typically registers filled once per kernel by team_fft_twiddles().
// The caller provides the team-wide execution barrier between the two halves (`__syncwarp()` on the device; the host
// test simply runs all lanes of stage A before stage B).
// tw points at this lane's first twiddle, element k1 is tw[k1 * tws]: a register array (tws = 1) or a table laid out
// [k1][h] (tw = table + h, tws = R).
template <int R>
BPC_HD void team_fft_stage_a(double2* a, const double2* tw, int tws, double2* xch, int h) {
    reg_dft<R>(a);
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
        double2 v = a[bitrev<R>(k1)];
        if (k1 > 0) v = c_mul(v, tw[k1 * tws]);
        xch[k1 * (R + 1) + h] = v;
    }
}
template <int R>
BPC_HD void team_fft_stage_b(double2* a, const double2* xch, int h) {
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = xch[h * (R + 1) + j];
    reg_dft<R>(a);
}

// ------------------------------------------------------------------------------- real-input split (rfft of 2N reals)
// The 2N real samples were packed as z[m] = x[2m] + i x[2m+1]; Z = FFT_N(z) sits in the team's registers as left by
// team_fft_stage_b.  X[k] = (Z[k] + conj(Z[N-k])) / 2 - i w^k (Z[k] - conj(Z[N-k])) / 2,  w = exp(-2 pi i / 2N).
// Bin k = h + R k2 pairs with N - k, which lives in lane (R - h) % R at k2' = R - 1 - k2 (h != 0) or (R - k2) % R (h == 0).
// rsplit_term<R, K2>(zk, zn, wl) returns 2 X[h + R K2] given zn = Z[N - k]; wl = -i * exp(-2 pi i h / 2N) (per lane).
template <int R, int K2>
BPC_HD double2 rsplit_term(double2 zk, double2 zn, double2 wl) {
    const double2 s = make_double2(zk.x + zn.x, zk.y - zn.y);           // zk + conj(zn)
    const double2 d = make_double2(zk.x - zn.x, zk.y + zn.y);           // zk - conj(zn)
    const double2 w = mul_tw<2 * R, K2>(wl);                            // -i * exp(-2 pi i (h + R K2) / 2N)
    return make_double2(s.x + fma(w.x, d.x, -(w.y * d.y)), s.y + fma(w.x, d.y, w.y * d.x));
}

// ------------------------------------------------------------------------ r02-k: N = 1024 on TWO warps (team of 64)
// 64 threads x 16 register-resident points, three register stages 16 x 16 x 4 around two exchanges (k_frame2048_w2:
// half the registers per thread of team_fft<32>, so 20 instead of 12 warps per SM).  With W = exp(-2 pi i / 1024),
// n = j + 64 n1 and k = k1 + 16 (k2 + 16 k3):
//   stage A  thread j           : B[j][k1]     = W^(j k1)        sum_n1 x[j + 64 n1]      W_16^(n1 k1)
//   stage B  thread (k1, j0)    : C[j0][k2]    = W_64^(j0 k2)    sum_m  B[j0 + 4 m][k1]   W_16^(m k2)
//   stage C  thread (k1, q)     : Z[k1 + 16 (q + 4 r) + 256 k3]  = sum_j0 C[j0][q + 4 r]  W_4^(j0 k3),   r, k3 < 4
// so the thread ends with Z[c + 64 u], c = k1 + 16 q, u = r + 4 k3 < 16, in register t64_zidx(u).  Exchange 1 goes
// through rows [k1][j] of pitch 68 (stage B reads j0 + 4 m of row k1: the sixteen (k1, j0) of a half-warp fall into
// sixteen different 8-byte banks because 68 = 4 mod 16 and the k1 of a half-warp are distinct mod 4); exchange 2 stays
// inside the four lanes that share k1 and reuses their row ([j0] at pitch 17).
// Thread -> (k1, q): lane = 4 g + q; the k1 of a warp are chosen so that the thread holding Z[N - k] (c' = (64 - c) % 64,
// i.e. k1' = (16 - k1) % 16) is in the SAME warp -- the real-input split then pairs bins with shuffles:
//   warp 0: k1 = 0 1 2 3 | 8 15 14 13      warp 1: k1 = 4 5 6 7 | 12 11 10 9      (g = 0..3 | 4..7)
constexpr int kT64Pitch = 68;
BPC_HD constexpr int t64_zidx(int u) { return 4 * (u & 3) + bitrev<4>(u >> 2); }
BPC_HD int t64_k1(int tid) {
    const int w = tid >> 5, g = (tid >> 2) & 7;
    return g < 4 ? 4 * w + g : ((w == 0 && g == 4) ? 8 : 16 - (4 * w + (g - 4)));
}
// lane (of the same warp) that holds row c' = (64 - c) % 64; Z[N - (c + 64 u)] is its register t64_zidx(15 - u), except
// for c = 0, where it is this thread's own t64_zidx((16 - u) % 16)
BPC_HD int t64_partner_lane(int tid) {
    const int l = tid & 31, g = l >> 2, q = l & 3, k1 = t64_k1(tid);
    if (k1 == 0) return (g << 2) | ((4 - q) & 3);
    if (k1 == 8) return (g << 2) | (3 - q);
    return ((g ^ 4) << 2) | (3 - q);
}
// stages A and B: radix-16 DFT of the 16 register points, then the inter-stage twiddles tw[k * tws], k = 1 .. 15
// (A: W^(j k1), table [k1][j], tw = table + j, tws = 64;  B: W_64^(j0 k2), table [k2][j0], tw = table + j0, tws = 4)
BPC_HD void team64_stage(double2* a, const double2* tw, int tws) {
    reg_dft<16>(a);
#pragma unroll
    for (int k = 1; k < 16; ++k) a[bitrev<16>(k)] = c_mul(a[bitrev<16>(k)], tw[k * tws]);
}
// stage C: four radix-4 DFTs over j0 (a[4 r + j0] -> a[4 r + bitrev_4(k3)])
BPC_HD void team64_stage_c(double2* a) {
    DitR<4, 0, 1>::run(a);
    DitR<4, 4, 1>::run(a);
    DitR<4, 8, 1>::run(a);
    DitR<4, 12, 1>::run(a);
}
// exchange indices (elements of the buffer, 16 * kT64Pitch of them)
BPC_HD constexpr int t64_x1_store(int k1, int j) { return k1 * kT64Pitch + j; }                 // stage A thread j, all k1
BPC_HD constexpr int t64_x1_load(int k1, int j0, int m) { return k1 * kT64Pitch + j0 + 4 * m; } // -> a[m]
BPC_HD constexpr int t64_x2_store(int k1, int j0, int k2) { return k1 * kT64Pitch + 17 * j0 + k2; }
BPC_HD constexpr int t64_x2_load(int k1, int q, int r, int j0) { return k1 * kT64Pitch + 17 * j0 + q + 4 * r; }   // -> a[4 r + j0]

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------ device wrappers
// Full transform for a team of R lanes inside one warp (R = 16: two teams per warp, R = 32: the warp).  All 32 lanes of
// the warp must call it together.  `xch` is the team's exchange buffer (R * (R + 1) double2).
template <int R>
__device__ __forceinline__ void team_fft(double2* a, const double2* tw, int tws, double2* xch, int h) {
    team_fft_stage_a<R>(a, tw, tws, xch, h);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = xch[h * (R + 1) + j];
    __syncwarp();                       // the buffer may be rewritten (next frame) once every lane has read its row
    reg_dft<R>(a);
}

// Same transform with the exchange done in two rounds (real parts, then imaginary parts) through a buffer of
// R * (R + 1) doubles: half the shared memory per team, which leaves the L1 enough room for the window / twiddle tables
// (k_frame2048: 16.9 KB per warp left the L1 ~50 KB and a 68 % hit rate).  Same wavefront count, 2 R more instructions.
template <int R>
__device__ __forceinline__ void team_fft_split(double2* a, const double2* tw, int tws, double* xr, int h) {
    reg_dft<R>(a);
#pragma unroll
    for (int k1 = 1; k1 < R; ++k1) a[bitrev<R>(k1)] = c_mul(a[bitrev<R>(k1)], tw[k1 * tws]);
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) xr[k1 * (R + 1) + h] = a[bitrev<R>(k1)].x;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j) a[j].x = xr[h * (R + 1) + j];
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) xr[k1 * (R + 1) + h] = a[bitrev<R>(k1)].y;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j) a[j].y = xr[h * (R + 1) + j];
    __syncwarp();                       // the buffer may be rewritten once every lane has read its row
    reg_dft<R>(a);
}

// Real-input split over bins k = h + R k2, k2 in [K2, K2HI]: calls emit(k, 2 X[k]).  `partner` is the warp lane that
// holds row (R - h) % R of the same team.  Must be called by all 32 lanes (shuffles).
template <int R, int K2, int K2HI, class Emit>
__device__ __forceinline__ void team_rsplit(const double2* a, double2 wl, int h, int partner, Emit& emit) {
    if constexpr (K2 <= K2HI) {
        const double2 zk = a[bitrev<R>(K2)];
        const double2 src = a[bitrev<R>(R - 1 - K2)];
        double2 zn;
        zn.x = __shfl_sync(0xffffffffu, src.x, partner);
        zn.y = __shfl_sync(0xffffffffu, src.y, partner);
        if (h == 0) zn = a[bitrev<R>((R - K2) % R)];
        emit(h + R * K2, rsplit_term<R, K2>(zk, zn, wl));
        team_rsplit<R, K2 + 1, K2HI>(a, wl, h, partner, emit);
    }
}
// Same split, every conjugate pair ONCE: X[k] and X[N - k] share s = Z[k] + conj(Z[N-k]) and p = w^k (Z[k] - conj(Z[N-k])),
//   2 X[k] = s + p,   2 X[N - k] = conj(s - p),
// so lane h evaluates k2 < R / 2 only and emits both bins (the partner lane covers the other half): half the shuffles
// and twiddle rotations, 12 instead of 20 FP64 operations per pair.  k = 0 yields X[N] (the Nyquist bin of the real
// transform) as its second output; k = N / 2 (lane 0, k2 = R / 2) pairs with itself.  emit(k, 2 X[k]) as above, for
// every k in [0, N] exactly once.  Must be called by all 32 lanes.
template <int R, int K2, class Emit>
__device__ __forceinline__ void team_rsplit_pairs(const double2* a, double2 wl, int h, int partner, Emit& emit) {
    if constexpr (K2 <= R / 2) {
        const double2 zk = a[bitrev<R>(K2)];
        double2 zn;
        if constexpr (K2 < R / 2) {
            const double2 src = a[bitrev<R>(R - 1 - K2)];
            zn.x = __shfl_sync(0xffffffffu, src.x, partner);
            zn.y = __shfl_sync(0xffffffffu, src.y, partner);
        }
        if (h == 0 || K2 == R / 2) zn = a[bitrev<R>((R - K2) % R)];
        const double2 s = make_double2(zk.x + zn.x, zk.y - zn.y);           // zk + conj(zn)
        const double2 d = make_double2(zk.x - zn.x, zk.y + zn.y);           // zk - conj(zn)
        const double2 w = mul_tw<2 * R, K2>(wl);                            // -i * exp(-2 pi i (h + R K2) / 2N)
        const double2 p = make_double2(fma(w.x, d.x, -(w.y * d.y)), fma(w.x, d.y, w.y * d.x));
        if (K2 < R / 2 || h == 0) {
            const int k = h + R * K2;
            emit(k, make_double2(s.x + p.x, s.y + p.y));
            if (K2 < R / 2) emit(R * R - k, make_double2(s.x - p.x, p.y - s.y));
        }
        team_rsplit_pairs<R, K2 + 1>(a, wl, h, partner, emit);
    }
}
// The whole transform for a CTA of exactly 64 threads (both warps call it together; it contains CTA barriers).
// in: a[n1] = x[tid + 64 n1]; out: a[t64_zidx(u)] = Z[c + 64 u], c = k1 + 16 q.  xr: 16 * kT64Pitch doubles, exchanged one
// component at a time (the other component waits in its registers).  twa = table [k1][j] + tid, twb = table [k2][j0] + q.
__device__ __forceinline__ void team64_fft(double2* a, const double2* twa, const double2* twb, double* xr, int tid,
                                           int k1, int q) {
    team64_stage(a, twa, 64);
#pragma unroll
    for (int k = 0; k < 16; ++k) xr[t64_x1_store(k, tid)] = a[bitrev<16>(k)].x;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m].x = xr[t64_x1_load(k1, q, m)];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) xr[t64_x1_store(k, tid)] = a[bitrev<16>(k)].y;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m].y = xr[t64_x1_load(k1, q, m)];
    __syncwarp();                       // row k1 is read by its own four lanes only: they may now rewrite it
    team64_stage(a, twb, 4);
#pragma unroll
    for (int k = 0; k < 16; ++k) xr[t64_x2_store(k1, q, k)] = a[bitrev<16>(k)].x;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j0 = 0; j0 < 4; ++j0) a[4 * r + j0].x = xr[t64_x2_load(k1, q, r, j0)];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) xr[t64_x2_store(k1, q, k)] = a[bitrev<16>(k)].y;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j0 = 0; j0 < 4; ++j0) a[4 * r + j0].y = xr[t64_x2_load(k1, q, r, j0)];
    team64_stage_c(a);
}
#endif



}  // namespace bpc
