// music_covn.cuh - K1 for M = 8 and M = 16: TMA-tiled covariance with the window staged ONCE per CTA group.
//
// The v1 tile kernels (cov_tile_kernel) give every 4x4 antenna tile its own warp that streams the
// window from global memory, i.e. a window is read 2x (M = 8) to 4x (M = 16) through L2 and the loads
// are not overlapped with the arithmetic.  Here a group of J warps (J = 1 for M = 8, J = 4 for M = 16)
// shares one ring of 4 KiB stages filled by cp.async.bulk.tensor (SASS UTMALDG) through a 3-D tensor
// map {32 floats = one 128-byte row (1 snapshot at M = 16, 2 at M = 8), N*M/16 rows, W windows}, box
// {32, 32, 1}, with CU_TENSOR_MAP_SWIZZLE_128B (rows narrower than the 128-byte swizzle span are padded
// to it in shared memory, hence the 128-byte row view): the hardware XOR-swizzle
// of the 16-byte units makes "lane <-> snapshot row, same unit" reads bank-conflict free (a plain 1-D
// copy would put all lanes on the same banks for 64/128-byte rows), and out-of-range rows of the last
// chunk of a window are zero-filled by TMA (zeros add nothing to x x^H).
//
// Jobs (equal cost: 128 DFMA per snapshot, 64 fp64 accumulators per lane).  The fp32 -> fp64 converts
// (F2F.F64.F32, 16 lanes/clk/SM on the XU pipe against 64 lanes/clk/SM of DFMA) are the second bound of
// this kernel, so the register blocks are as large as the register file allows: 8 x 4 complex entries
// per lane need 16 or 24 converts for 128 DFMA (a 4 x 4 block would need 16 for 64 and be XU-bound).
//   SYM(a, b)     : block R[4a..][4b..] + Hermitian halves of the diagonal blocks a and b (16 converts)
//   WIDE(a; b, c) : blocks R[4a..][4b..] and R[4a..][4c..]                                (24 converts)
// M = 8: {SYM(0,1)}, one warp per window;  M = 16: {SYM(0,1), SYM(2,3), WIDE(0;2,3), WIDE(1;2,3)} =
// 2 M^2 DFMA per snapshot in total, the minimum for a Hermitian R.  Windows are claimed dynamically per
// group (global ticket counter, self-resetting like the fused kernel's).
//
// Reference lines covered: /root/reference/lib/baz_music_doa.cc:74-85.
#pragma once
#include <cuda.h>

#include "music_kernels.cuh"

namespace music {

constexpr int CN_WARPS = 8;
constexpr int CN_RING_BYTES = 192 * 1024;  // TMA stages of one CTA, split between the groups
constexpr int CN_MAX_STAGES = 32;

template <int M> struct CovNJobs;
template <> struct CovNJobs<8> {
    static constexpr int J = 1;
    static constexpr int LAG = 0;     // the only consumer of a stage is the producer warp itself
    static constexpr int STEPS = 4;   // 32-snapshot steps per stage: 8 KiB stages
    static constexpr int SG = 3;      // stages per group (8 groups x 3 x 8 KiB = 192 KiB)
    __device__ static void get(int, bool &sym, int &a, int &b, int &c) { sym = true; a = 0; b = 1; c = 1; }
};
template <> struct CovNJobs<16> {
    static constexpr int J = 4;
    static constexpr int LAG = 2;     // a stage is refilled LAG iterations after the producer released it
    static constexpr int STEPS = 2;   // 8 KiB stages
    static constexpr int SG = 12;     // 2 groups x 12 x 8 KiB = 192 KiB
    __device__ static void get(int j, bool &sym, int &a, int &b, int &c)
    {
        sym = j < 2;
        a = (j == 0) ? 0 : (j == 1) ? 2 : (j == 2) ? 0 : 1;
        b = (j == 0) ? 1 : (j == 1) ? 3 : 2;
        c = 3;
    }
};

// 16-byte unit u of snapshot row r inside a 128B-swizzled stage (stage base is 1024-byte aligned)
template <int M>
__device__ __forceinline__ float4 covn_unit(const unsigned char *stage, int r, int u)
{
    const uint32_t off = (uint32_t)r * (8 * M) + (uint32_t)u * 16;
    return *reinterpret_cast<const float4 *>(stage + (off ^ ((off >> 3) & 0x70)));  // bits 4-6 ^= bits 7-9
}

struct CovnQuad {  // four consecutive antennas of one snapshot, converted once
    double r[4], i[4];
};
template <int M>
__device__ __forceinline__ CovnQuad covn_quad(const unsigned char *stage, int row, int blk)
{
    const float4 u0 = covn_unit<M>(stage, row, 2 * blk), u1 = covn_unit<M>(stage, row, 2 * blk + 1);
    CovnQuad q;
    q.r[0] = u0.x; q.i[0] = u0.y; q.r[1] = u0.z; q.i[1] = u0.w;
    q.r[2] = u1.x; q.i[2] = u1.y; q.r[3] = u1.z; q.i[3] = u1.w;
    return q;
}

__device__ __forceinline__ void covn_acc_off(double *acc /*32*/, const CovnQuad &a, const CovnQuad &b)
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // x_i * conj(y_j)
            acc[2 * (i * 4 + j)] = fma(a.r[i], b.r[j], fma(a.i[i], b.i[j], acc[2 * (i * 4 + j)]));
            acc[2 * (i * 4 + j) + 1] = fma(a.i[i], b.r[j], fma(-a.r[i], b.i[j], acc[2 * (i * 4 + j) + 1]));
        }
}

__device__ __forceinline__ void covn_acc_diag(double *acc /*16*/, const CovnQuad &a)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fma(a.r[i], a.r[i], fma(a.i[i], a.i[i], acc[i]));
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            const int e = 4 + 2 * (i * 3 - (i * (i - 1)) / 2 + (j - i - 1));  // 4 + 2 * index of (i, j) in the strict upper triangle
            acc[e] = fma(a.r[i], a.r[j], fma(a.i[i], a.i[j], acc[e]));
            acc[e + 1] = fma(a.i[i], a.r[j], fma(-a.r[i], a.i[j], acc[e + 1]));
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
    constexpr int SG = CovNJobs<M>::SG;             // stages per group
    constexpr int STEPS = CovNJobs<M>::STEPS;
    constexpr int ROWS = 32 * STEPS;                // snapshots per stage
    constexpr int STAGE = ROWS * 8 * M;             // bytes per stage
    static_assert(G * SG * STAGE <= CN_RING_BYTES && G * SG <= CN_MAX_STAGES && STAGE % 1024 == 0, "ring layout");
    extern __shared__ __align__(1024) unsigned char cn_smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned stages: realign explicitly (CN_SMEM carries the slack)
    unsigned char *cn_smem = cn_smem_raw + ((1024u - (smem_u32(cn_smem_raw) & 1023u)) & 1023u);
    // layout: [0, CN_RING_BYTES) stages | full barriers [32] | empty barriers [32] | window-id rings [G][8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(cn_smem + CN_RING_BYTES);
    volatile int *wring_all = reinterpret_cast<volatile int *>(cn_smem + CN_RING_BYTES + 2 * CN_MAX_STAGES * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp / J, job = warp % J;
    const bool producer = (job == 0 && lane == 0);
    unsigned char *ring = cn_smem + (size_t)g * SG * STAGE;
    const uint32_t full0 = smem_u32(bars + g * SG), empty0 = smem_u32(bars + CN_MAX_STAGES + g * SG), ring0 = smem_u32(ring);
    volatile int *wring = wring_all + g * 8;
    if (producer) {
        for (int s = 0; s < SG; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, J); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int cpw = (N + ROWS - 1) / ROWS;  // chunks per window (the last one is zero-filled beyond N)
    constexpr int LAG = CovNJobs<M>::LAG;
    bool sym;
    int ta, tb, tc;
    CovNJobs<M>::get(job, sym, ta, tb, tc);

    // producer state: next chunk to request = chunk iq of window iw; T_issued counts requests of this group
    int iq = 0, iw = -1, wr = 0;
    unsigned issued = 0;
    auto claim = [&]() {
        const unsigned tkt = atomicAdd(&work_ctr[0], 1u);
        iw = tkt < (unsigned)W ? (int)tkt : -1;
        wring[wr & 7] = iw;
        ++wr;
    };
    const uint64_t pol_stream = l2_policy_evict_first();
    auto issue = [&]() {  // producer only; the slot must be free
        const int slot = (int)(issued % SG);
        mbar_expect_tx(full0 + 8 * slot, STAGE);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                     ::"r"(ring0 + slot * STAGE), "l"(tm_addr), "r"(0), "r"(iq * (STAGE / 128)), "r"(iw), "r"(full0 + 8 * slot), "l"(pol_stream)
                     : "memory");  // evict_first: the stream is read once and must not push the steering table out of L2
        ++issued;
        if (++iq == cpw) { iq = 0; claim(); }
    };
    if (producer) {
        claim();
        for (int s = 0; s < SG - LAG && iw >= 0; ++s) issue();
    }
    __syncwarp();
    // all warps of the group must see the first ring entry
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(J * 32) : "memory");

    double acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.0;
    unsigned consumed = 0;  // chunks consumed by this warp (same sequence in every warp of the group)
    for (int rd = 0;; ++rd) {
        const int wcur = wring[rd & 7];
        if (wcur < 0) break;
        for (int q = 0; q < cpw; ++q, ++consumed) {
            const int slot = (int)(consumed % SG);
            while (!mbar_try_wait(full0 + 8 * slot, (consumed / SG) & 1)) {}
            const unsigned char *stage = ring + (size_t)slot * STAGE;
#pragma unroll
            for (int rr = 0; rr < STEPS; ++rr) {
                const int r = lane + 32 * rr;
                const CovnQuad a = covn_quad<M>(stage, r, ta), b = covn_quad<M>(stage, r, tb);
                covn_acc_off(acc, a, b);
                if (sym) {
                    covn_acc_diag(acc + 32, a);
                    covn_acc_diag(acc + 48, b);
                } else {
                    const CovnQuad c = covn_quad<M>(stage, r, tc);
                    covn_acc_off(acc + 32, a, c);
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * slot) : "memory");
            if (producer && iw >= 0) {
                // refill the slot that will hold request number `issued`: it held request issued - SG, which every
                // warp of the group must have released (the producer itself released it LAG iterations ago)
                if (issued >= (unsigned)SG) {
                    const unsigned old = issued - SG;
                    while (!mbar_try_wait(empty0 + 8 * (old % SG), (old / SG) & 1)) {}
                }
                issue();
            }
        }
        // the producer may have claimed the next window during this one: make the ring entry visible to the group
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(J * 32) : "memory");
        // transposing butterfly: 62 exchanges instead of 64 x 5, and lane l ends up owning entries 2l and 2l + 1
        // (= one complex entry of an off-diagonal block, or two reals / one complex entry of a diagonal half)
#pragma unroll
        for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const double send = upper ? acc[i] : acc[i + n];
                const double keep = upper ? acc[i + n] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        {
            const double dn = (double)N;
            const double v0 = acc[0] / dn, v1 = acc[1] / dn;
            double2 *Rw = reinterpret_cast<double2 *>(R) + (size_t)wcur * M * M;
            if (lane < 16 || !sym) {  // complex entry (i, j) of block (ta, lane < 16 ? tb : tc)
                const int e = lane & 15, r = 4 * ta + (e >> 2), c = 4 * (lane < 16 ? tb : tc) + (e & 3);
                Rw[r * M + c] = make_double2(v0, v1);
                Rw[c * M + r] = make_double2(v0, -v1);
            } else {  // diagonal half of block I: lanes 0-1 of the octet hold the 4 real diagonals, lanes 2-7 the 6 pairs
                const int I = lane < 24 ? ta : tb, e = lane & 7;
                if (e < 2) {
                    const int r = 4 * I + 2 * e;
                    Rw[r * M + r] = make_double2(v0, 0.0);
                    Rw[(r + 1) * M + r + 1] = make_double2(v1, 0.0);
                } else {
                    const int p = e - 2;  // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
                    const int i = p < 3 ? 0 : (p < 5 ? 1 : 2), j = p < 3 ? p + 1 : (p < 5 ? p - 1 : 3);
                    const int r = 4 * I + i, c = 4 * I + j;
                    Rw[r * M + c] = make_double2(v0, v1);
                    Rw[c * M + r] = make_double2(v0, -v1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0;
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

constexpr size_t CN_SMEM = (size_t)CN_RING_BYTES + 2 * CN_MAX_STAGES * 8 + 8 * 8 * 4 + 1024;

}  // namespace music
