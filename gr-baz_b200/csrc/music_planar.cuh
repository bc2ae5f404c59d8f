// music_planar.cuh - K1 for PLANAR input: one c64 stream per antenna, windows formed by pointer arithmetic
// (SURVEY.md section 8(f) rank 2).
//
// In a gr-baz flowgraph the MUSIC block is fed by M antenna streams that the CPU first interleaves
// (/root/reference/lib/baz_interleaver.cc:152-229, or blocks.interleave), cuts into vectors
// (stream_to_vector) and, for sliding windows, copies again (/root/reference/lib/baz_overlap.cc:107-129:
// every output item re-copies the overlapping samples and consume_each() advances by the hop).  Those copies
// produce exactly  x_w(r, c) = stream_r[w * hop + c],  r < M, c < N  - the matrix lib/baz_music_doa.cc:82-84
// reshapes out of the interleaved item.  This kernel reads that matrix straight from the M streams:
// window w starts hop snapshots after window w - 1 (hop == N: back-to-back vectors; hop < N: overlap N - hop,
// served from L2 instead of being copied), so no interleaved copy of the input ever exists.
//
// One warp per (window, 4x4 antenna tile) like cov_tile_kernel; lanes stride over snapshots, each load is a
// fully coalesced 256-byte row of one antenna's stream.  Antenna indices >= M (M not a multiple of 4) are
// treated as silent elements and never loaded or stored.
#pragma once
#include "music_kernels.cuh"

namespace music {

struct PlanarStreams {
    const float2 *p[MAXM];  // device pointers, one per antenna
};

__device__ __forceinline__ float2 ldg_stream2(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

template <bool OFF>
__global__ void __launch_bounds__(256) cov_planar_kernel(const PlanarStreams S, unsigned long long first_snapshot,
                                                         unsigned hop, double *__restrict__ R, int W, int N, int M)
{
    const int T = (M + 3) >> 2;
    const int tiles = OFF ? (T * (T - 1)) / 2 : T;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (item >= (long long)W * tiles) return;
    const int w = (int)(item / tiles);
    int t = (int)(item % tiles);
    int I, J;
    if (OFF) {
        I = 0;
        while (t >= T - 1 - I) { t -= T - 1 - I; ++I; }
        J = I + 1 + t;
    } else {
        I = J = t;
    }
    const unsigned long long s0 = first_snapshot + (unsigned long long)w * hop;
    const float2 *pa[4], *pb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pa[i] = (4 * I + i < M) ? S.p[4 * I + i] + s0 : nullptr;
        pb[i] = (OFF && 4 * J + i < M) ? S.p[4 * J + i] + s0 : nullptr;
    }

    double acc[OFF ? 32 : 16];
#pragma unroll
    for (int i = 0; i < (OFF ? 32 : 16); ++i) acc[i] = 0.0;

    constexpr int U = 4;  // snapshots in flight per lane
    for (int c0 = lane; c0 < N; c0 += 32 * U) {
        float2 xa[U][4], xb[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xa[u][i] = (c < N && pa[i]) ? ldg_stream2(pa[i] + c) : make_float2(0.f, 0.f);
                if (OFF) xb[u][i] = (c < N && pb[i]) ? ldg_stream2(pb[i] + c) : make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double ar[4], ai[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ar[i] = xa[u][i].x; ai[i] = xa[u][i].y; }
            if (OFF) {
                double br[4], bi[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { br[i] = xb[u][i].x; bi[i] = xb[u][i].y; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {  // x_i * conj(y_j)
                        acc[2 * (i * 4 + j)] = fma(ar[i], br[j], fma(ai[i], bi[j], acc[2 * (i * 4 + j)]));
                        acc[2 * (i * 4 + j) + 1] = fma(ai[i], br[j], fma(-ar[i], bi[j], acc[2 * (i * 4 + j) + 1]));
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fma(ar[i], ar[i], fma(ai[i], ai[i], acc[i]));
                int e = 4;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = i + 1; j < 4; ++j) {
                        acc[e] = fma(ar[i], ar[j], fma(ai[i], ai[j], acc[e]));
                        acc[e + 1] = fma(ai[i], ar[j], fma(-ar[i], ai[j], acc[e + 1]));
                        e += 2;
                    }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < (OFF ? 32 : 16); ++i) acc[i] = warp_sum(acc[i]);
    if (lane != 0) return;
    const double dn = (double)N;
    double *Rw = R + (size_t)w * M * M * 2;
    if (OFF) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                const int r = 4 * I + i, c = 4 * J + j;
                if (r >= M || c >= M) continue;
                const double re = acc[2 * (i * 4 + j)] / dn, im = acc[2 * (i * 4 + j) + 1] / dn;
                Rw[2 * (r * M + c)] = re;  Rw[2 * (r * M + c) + 1] = im;
                Rw[2 * (c * M + r)] = re;  Rw[2 * (c * M + r) + 1] = -im;
            }
    } else {
        for (int i = 0; i < 4; ++i) {
            const int r = 4 * I + i;
            if (r >= M) continue;
            Rw[2 * (r * M + r)] = acc[i] / dn;
            Rw[2 * (r * M + r) + 1] = 0.0;
        }
        int e = 4;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j, e += 2) {
                const int r = 4 * I + i, c = 4 * I + j;
                if (r >= M || c >= M) continue;
                const double re = acc[e] / dn, im = acc[e + 1] / dn;
                Rw[2 * (r * M + c)] = re;  Rw[2 * (r * M + c) + 1] = im;
                Rw[2 * (c * M + r)] = re;  Rw[2 * (c * M + r) + 1] = -im;
            }
    }
}

}  // namespace music
