// music_covn.cuh - K1 for M = 8 and M = 16: TMA-tiled covariance with the window staged ONCE per CTA group.
//
// The v1 tile kernels (cov_tile_kernel) give every 4x4 antenna tile its own warp that streams the
// window from global memory, i.e. a window is read 2x (M = 8) to 4x (M = 16) through L2 and the loads
// are not overlapped with the arithmetic.  Here a group of J warps (J = 2 for M = 8, J = 8 for M = 16)
// shares one ring of 4 KiB stages filled by cp.async.bulk.tensor (SASS UTMALDG) through a 3-D tensor
// map {32 floats = one 128-byte row (1 snapshot at M = 16, 2 at M = 8), N*M/16 rows, W windows}, box
// {32, 32, 1}, with CU_TENSOR_MAP_SWIZZLE_128B (rows narrower than the 128-byte swizzle span are padded
// to it in shared memory, hence the 128-byte row view): the hardware XOR-swizzle
// of the 16-byte units makes "lane <-> snapshot row, same unit" reads bank-conflict free (a plain 1-D
// copy would put all lanes on the same banks for 64/128-byte rows), and out-of-range rows of the last
// chunk of a window are zero-filled by TMA (zeros add nothing to x x^H).
//
// Jobs (equal cost: 64 DFMA per snapshot, 32 fp64 accumulators per lane):
//   OFF(I, J), I < J : full 4x4 complex block  R[4I..][4J..]
//   DIAG2(I1, I2)    : Hermitian halves of the two diagonal blocks I1 and I2
// M = 8: {OFF(0,1), DIAG2(0,1)};  M = 16: the 6 OFF blocks + DIAG2(0,1) + DIAG2(2,3) = 2 M^2 DFMA per
// snapshot in total, the minimum for a Hermitian R.  Windows are claimed dynamically per group
// (global ticket counter, self-resetting like the fused kernel's).
//
// Reference lines covered: /root/reference/lib/baz_music_doa.cc:74-85.
#pragma once
#include <cuda.h>

#include "music_kernels.cuh"

namespace music {

constexpr int CN_WARPS = 8;
constexpr int CN_STAGES_TOTAL = 24;  // 4 KiB stages per CTA (96 KiB), split between the groups
constexpr int CN_LAG = 2;            // a stage is refilled CN_LAG iterations after the producer released it

template <int M> struct CovNJobs;
template <> struct CovNJobs<8> {
    static constexpr int J = 2;
    __device__ static void get(int j, bool &off, int &a, int &b) { off = (j == 0); a = 0; b = 1; }
};
template <> struct CovNJobs<16> {
    static constexpr int J = 8;
    __device__ static void get(int j, bool &off, int &a, int &b)
    {
        const int oa[8] = {0, 0, 0, 1, 1, 2, 0, 2}, ob[8] = {1, 2, 3, 2, 3, 3, 1, 3};
        off = j < 6; a = oa[j]; b = ob[j];
    }
};

// 16-byte unit u of snapshot row r inside a 128B-swizzled stage (stage base is 1024-byte aligned)
template <int M>
__device__ __forceinline__ float4 covn_unit(const unsigned char *stage, int r, int u)
{
    const uint32_t off = (uint32_t)r * (8 * M) + (uint32_t)u * 16;
    return *reinterpret_cast<const float4 *>(stage + (off ^ ((off >> 3) & 0x70)));  // bits 4-6 ^= bits 7-9
}

__device__ __forceinline__ void covn_acc_off(double (&acc)[32], const float4 a0, const float4 a1, const float4 b0, const float4 b1)
{
    const double ar[4] = {a0.x, a0.z, a1.x, a1.z}, ai[4] = {a0.y, a0.w, a1.y, a1.w};
    const double br[4] = {b0.x, b0.z, b1.x, b1.z}, bi[4] = {b0.y, b0.w, b1.y, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // x_i * conj(y_j)
            acc[2 * (i * 4 + j)] = fma(ar[i], br[j], fma(ai[i], bi[j], acc[2 * (i * 4 + j)]));
            acc[2 * (i * 4 + j) + 1] = fma(ai[i], br[j], fma(-ar[i], bi[j], acc[2 * (i * 4 + j) + 1]));
        }
}

__device__ __forceinline__ void covn_acc_diag(double *acc /*16*/, const float4 a0, const float4 a1)
{
    const double r[4] = {a0.x, a0.z, a1.x, a1.z}, im[4] = {a0.y, a0.w, a1.y, a1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fma(r[i], r[i], fma(im[i], im[i], acc[i]));
    int e = 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            acc[e] = fma(r[i], r[j], fma(im[i], im[j], acc[e]));
            acc[e + 1] = fma(im[i], r[j], fma(-r[i], im[j], acc[e + 1]));
            e += 2;
        }
}

__device__ __forceinline__ void covn_write_diag(double *Rw, int M, int I, const double *acc, double dn)
{
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * I + i;
        Rw[2 * (r * M + r)] = acc[i] / dn;
        Rw[2 * (r * M + r) + 1] = 0.0;
    }
    int e = 4;
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) {
            const double re = acc[e] / dn, im = acc[e + 1] / dn;
            const int r = 4 * I + i, c = 4 * I + j;
            Rw[2 * (r * M + c)] = re;  Rw[2 * (r * M + c) + 1] = im;
            Rw[2 * (c * M + r)] = re;  Rw[2 * (c * M + r) + 1] = -im;
            e += 2;
        }
}

template <int M>
__global__ void __launch_bounds__(CN_WARPS * 32, 1)
covN_tma_kernel(const __grid_constant__ CUtensorMap tm, double *__restrict__ R, int W, int N, unsigned *__restrict__ work_ctr)
{
    // address of the tensor map in param space, taken in the kernel body (not inside a lambda: a by-reference
    // capture would make the compiler copy the map to local memory, which TMA cannot read)
    const unsigned long long tm_addr = reinterpret_cast<unsigned long long>(&tm);
    constexpr int J = CovNJobs<M>::J;
    constexpr int G = CN_WARPS / J;                 // window groups per CTA
    constexpr int SG = CN_STAGES_TOTAL / G;         // stages per group
    constexpr int ROWS = COV_CHUNK / (8 * M);       // snapshots per stage (64 or 32)
    extern __shared__ __align__(1024) unsigned char cn_smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned stages: realign explicitly (CN_SMEM carries the slack)
    unsigned char *cn_smem = cn_smem_raw + ((1024u - (smem_u32(cn_smem_raw) & 1023u)) & 1023u);
    // layout: [0, 24 * 4096) stages | full barriers [24] | empty barriers [24] | window-id rings [G][8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(cn_smem + CN_STAGES_TOTAL * COV_CHUNK);
    volatile int *wring_all = reinterpret_cast<volatile int *>(cn_smem + CN_STAGES_TOTAL * COV_CHUNK + 2 * CN_STAGES_TOTAL * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp / J, job = warp % J;
    const bool producer = (job == 0 && lane == 0);
    unsigned char *ring = cn_smem + (size_t)g * SG * COV_CHUNK;
    const uint32_t full0 = smem_u32(bars + g * SG), empty0 = smem_u32(bars + CN_STAGES_TOTAL + g * SG), ring0 = smem_u32(ring);
    volatile int *wring = wring_all + g * 8;
    if (producer) {
        for (int s = 0; s < SG; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, J); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int cpw = (N + ROWS - 1) / ROWS;  // chunks per window (the last one is zero-filled beyond N)
    bool off;
    int ta, tb;
    CovNJobs<M>::get(job, off, ta, tb);

    // producer state: next chunk to request = chunk iq of window iw; T_issued counts requests of this group
    int iq = 0, iw = -1, wr = 0;
    unsigned issued = 0;
    auto claim = [&]() {
        const unsigned tkt = atomicAdd(&work_ctr[0], 1u);
        iw = tkt < (unsigned)W ? (int)tkt : -1;
        wring[wr & 7] = iw;
        ++wr;
    };
    auto issue = [&]() {  // producer only; the slot must be free
        const int slot = (int)(issued % SG);
        mbar_expect_tx(full0 + 8 * slot, COV_CHUNK);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(ring0 + slot * COV_CHUNK), "l"(tm_addr), "r"(0), "r"(iq * 32), "r"(iw), "r"(full0 + 8 * slot)
                     : "memory");
        ++issued;
        if (++iq == cpw) { iq = 0; claim(); }
    };
    if (producer) {
        claim();
        for (int s = 0; s < SG - CN_LAG && iw >= 0; ++s) issue();
    }
    __syncwarp();
    // all warps of the group must see the first ring entry
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(J * 32) : "memory");

    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    unsigned consumed = 0;  // chunks consumed by this warp (same sequence in every warp of the group)
    for (int rd = 0;; ++rd) {
        const int wcur = wring[rd & 7];
        if (wcur < 0) break;
        for (int q = 0; q < cpw; ++q, ++consumed) {
            const int slot = (int)(consumed % SG);
            while (!mbar_try_wait(full0 + 8 * slot, (consumed / SG) & 1)) {}
            const unsigned char *stage = ring + (size_t)slot * COV_CHUNK;
#pragma unroll
            for (int rr = 0; rr < ROWS / 32; ++rr) {
                const int r = lane + 32 * rr;
                const float4 a0 = covn_unit<M>(stage, r, 2 * ta), a1 = covn_unit<M>(stage, r, 2 * ta + 1);
                const float4 b0 = covn_unit<M>(stage, r, 2 * tb), b1 = covn_unit<M>(stage, r, 2 * tb + 1);
                if (off) covn_acc_off(acc, a0, a1, b0, b1);
                else { covn_acc_diag(acc, a0, a1); covn_acc_diag(acc + 16, b0, b1); }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * slot) : "memory");
            if (producer && iw >= 0) {
                // refill the slot that will hold request number `issued`: it held request issued - SG, which every
                // warp of the group must have released (the producer itself released it CN_LAG iterations ago)
                if (issued >= (unsigned)SG) {
                    const unsigned old = issued - SG;
                    while (!mbar_try_wait(empty0 + 8 * (old % SG), (old / SG) & 1)) {}
                }
                issue();
            }
        }
        // the producer may have claimed the next window during this one: make the ring entry visible to the group
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(J * 32) : "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = warp_sum(acc[i]);
        if (lane == 0) {
            const double dn = (double)N;
            double *Rw = R + (size_t)wcur * M * M * 2;
            if (off) {
                for (int i = 0; i < 4; ++i)
                    for (int j = 0; j < 4; ++j) {
                        const double re = acc[2 * (i * 4 + j)] / dn, im = acc[2 * (i * 4 + j) + 1] / dn;
                        const int r = 4 * ta + i, c = 4 * tb + j;
                        Rw[2 * (r * M + c)] = re;  Rw[2 * (r * M + c) + 1] = im;
                        Rw[2 * (c * M + r)] = re;  Rw[2 * (c * M + r) + 1] = -im;
                    }
            } else {
                covn_write_diag(Rw, M, ta, acc, dn);
                covn_write_diag(Rw, M, tb, acc + 16, dn);
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    }
    // the last CTA to finish re-arms the ticket counter (launches of one handle are serialised by the host)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&work_ctr[1], 1u) == gridDim.x - 1) {
            work_ctr[0] = 0;
            work_ctr[1] = 0;
            __threadfence();
        }
    }
}

constexpr size_t CN_SMEM = (size_t)CN_STAGES_TOTAL * COV_CHUNK + 2 * CN_STAGES_TOTAL * 8 + 4 * 8 * 4 + 1024;

}  // namespace music
