// music_fused.cuh - FUSED persistent kernel for the headline shape class (M = 4, n = 1, peak
// outputs only): K1 + K2 + K3 in one launch, one CTA per SM, warp-specialised; R and the
// eigenvectors never leave shared memory.
//
//   warps 0..7   covariance: per-warp TMA ring exactly as cov4_tma_kernel; a finished window's R
//                is pushed into a 64-slot shared-memory queue (in-order publish).
//   warp  8      eigensolver: one lane per queued window (up to 32 at once), cyclic Jacobi
//                (herm_eig_body<4>), eigenvectors written next to the queue slot.
//   warps 9..15  pseudospectrum scan + peak pick of up to 4 windows per pass: thread <-> bin,
//                same hot loop as scan_peak1_kernel, results merged over the 7 warps and written
//                to global (angle, level, bin).
//
// The stages overlap in time on every SM: while the covariance warps keep HBM busy (their FP64
// demand is ~1/3 of the pipe), the scan warps fill the remaining FP64 issue slots.  Counters in
// shared memory (cov_pub >= eig_done >= scan_done) hand windows from stage to stage.
//
// Reference lines covered: /root/reference/lib/baz_music_doa.cc:74-155 (everything work() does
// per window except the optional spectrum port).
#pragma once
#include "music_kernels.cuh"

namespace music {

constexpr int FZ_COV_WARPS = 8;
constexpr int FZ_SCAN_WARPS = 7;
constexpr int FZ_THREADS = 32 * (FZ_COV_WARPS + 1 + FZ_SCAN_WARPS);  // 512
constexpr int FZ_SCAN_THREADS = 32 * FZ_SCAN_WARPS;                  // 224 bins per pass iteration
constexpr int FZ_Q = 64;        // window queue slots per CTA
constexpr int FZ_WPT = 4;       // windows per scan pass
constexpr int FZ_STAGES = 4;    // 4 KiB TMA stages per covariance warp
constexpr int FZ_TS = 3;        // steering-table tile stages (TMA ring shared by the scan warps)
constexpr int FZ_TCOMP = 9;     // doubles per table row for M = 4: Re/Im a_0..a_3, ||a||^2
constexpr int FZ_TILE_BYTES = FZ_TCOMP * FZ_SCAN_THREADS * 8;  // 16128 B: one 224-row tile, [comp][224]

struct FusedCtl {               // shared-memory control block
    unsigned cov_seq;           // tickets handed to covariance warps
    volatile unsigned cov_pub;  // windows whose R is in the queue
    volatile unsigned eig_done; // windows whose eigenvectors are in the queue
    volatile unsigned scan_done;// windows fully processed (queue slot free)
    volatile unsigned cov_finished;  // covariance warps that ran out of windows
    volatile unsigned batch_start, batch_cnt;
    unsigned pad;
};

constexpr size_t FZ_OFF_CTL = 1024;
constexpr size_t FZ_OFF_WIN = 1088;                           // int win[FZ_Q]
constexpr size_t FZ_OFF_RED = FZ_OFF_WIN + 4 * FZ_Q;          // reduction scratch: 7 warps x 4 x (double, int)
constexpr size_t FZ_OFF_RQ = 2048;                            // double Rq[FZ_Q][32]
constexpr size_t FZ_OFF_VQ = FZ_OFF_RQ + (size_t)FZ_Q * 256;  // double Vq[FZ_Q][32]
constexpr size_t FZ_OFF_TBL = FZ_OFF_VQ + (size_t)FZ_Q * 256;   // FZ_TS table tiles
constexpr size_t FZ_OFF_RING = (FZ_OFF_TBL + (size_t)FZ_TS * FZ_TILE_BYTES + 127) / 128 * 128;
constexpr size_t FZ_SMEM = FZ_OFF_RING + (size_t)FZ_COV_WARPS * FZ_STAGES * COV_CHUNK;
constexpr size_t FZ_OFF_TBAR = 512;                            // uint64 tfull[FZ_TS], tempty[FZ_TS]
static_assert(FZ_COV_WARPS * FZ_STAGES * 8 <= FZ_OFF_TBAR, "covariance barriers overlap the table barriers");
static_assert(FZ_SMEM <= 227 * 1024, "fused kernel shared memory");
static_assert(FZ_OFF_RED + 12 * FZ_SCAN_WARPS * FZ_WPT <= FZ_OFF_RQ, "control area overflow");

__device__ __forceinline__ void bar_sync_scan() { asm volatile("bar.sync 1, %0;" ::"n"(FZ_SCAN_THREADS) : "memory"); }

// Bins covered by the scan warps per window: niter * 224; the steering table must be padded to at
// least that many rows (prep_table_kernel pads whole TILE-row tiles with ||a||^2 = +inf).
__host__ __device__ inline int fused_scan_rows(int K) { return (K + FZ_SCAN_THREADS - 1) / FZ_SCAN_THREADS * FZ_SCAN_THREADS; }

// Steering table in the fused kernel's layout: tiles of 224 rows, [tile][comp][224] fp64, so that one
// 1-D bulk copy brings a whole tile; rows >= K are padding with ||a||^2 = +inf (never win).
__global__ void prep_table_fused_kernel(const float2 *__restrict__ tab, double *__restrict__ tbl, int K)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;  // row
    if (r >= fused_scan_rows(K)) return;
    double *base = tbl + (size_t)(r / FZ_SCAN_THREADS) * (FZ_TCOMP * FZ_SCAN_THREADS) + (r % FZ_SCAN_THREADS);
    double na = 0.0;
    for (int i = 0; i < 4; ++i) {
        double re = 0.0, im = 0.0;
        if (r < K) { const float2 a = tab[(size_t)r * 4 + i]; re = a.x; im = a.y; }
        base[(size_t)(2 * i) * FZ_SCAN_THREADS] = re;
        base[(size_t)(2 * i + 1) * FZ_SCAN_THREADS] = im;
        na = fma(re, re, fma(im, im, na));
    }
    base[(size_t)8 * FZ_SCAN_THREADS] = (r < K) ? na : __longlong_as_double(0x7ff0000000000000LL);
}

__global__ void __launch_bounds__(FZ_THREADS, 1)
music4_fused_kernel(const float *__restrict__ in, const double *__restrict__ tbl /* [niter][9][224] */, int W, int N, int K, PeakOut out,
                    long long *__restrict__ dbg /* optional [grid][16] clock64 trace, may be null */)
{
    const long long t_start = clock64();
    extern __shared__ __align__(128) unsigned char fz_smem[];
    FusedCtl *ctl = reinterpret_cast<FusedCtl *>(fz_smem + FZ_OFF_CTL);
    int *qwin = reinterpret_cast<int *>(fz_smem + FZ_OFF_WIN);
    double *Rq = reinterpret_cast<double *>(fz_smem + FZ_OFF_RQ);
    double *Vq = reinterpret_cast<double *>(fz_smem + FZ_OFF_VQ);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ctl->cov_seq = 0; ctl->cov_pub = 0; ctl->eig_done = 0; ctl->scan_done = 0; ctl->cov_finished = 0;
        ctl->batch_start = 0; ctl->batch_cnt = 0;
        const uint32_t tb0 = smem_u32(fz_smem + FZ_OFF_TBAR);
        for (int s = 0; s < FZ_TS; ++s) {
            mbar_init(tb0 + 8 * s, 1);                          // tfull: one producer arrival + tx bytes
            mbar_init(tb0 + 8 * (FZ_TS + s), FZ_SCAN_WARPS);    // tempty: one arrival per scan warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp < FZ_COV_WARPS) {
        // ================= covariance warps =================
        uint64_t *bars = reinterpret_cast<uint64_t *>(fz_smem) + warp * FZ_STAGES;
        unsigned char *ring = fz_smem + FZ_OFF_RING + (size_t)warp * FZ_STAGES * COV_CHUNK;
        const uint32_t bar0 = smem_u32(bars), ring0 = smem_u32(ring);
        const int gw = blockIdx.x * FZ_COV_WARPS + warp, total_warps = gridDim.x * FZ_COV_WARPS;
        const size_t win_bytes = (size_t)N * 32;
        const int cpw = (int)((win_bytes + COV_CHUNK - 1) / COV_CHUNK);
        const int nwin = gw < W ? (W - gw + total_warps - 1) / total_warps : 0;
        const long long total = (long long)nwin * cpw;
        const unsigned char *src0 = reinterpret_cast<const unsigned char *>(in);
        if (lane == 0) {
            for (int s = 0; s < FZ_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        auto issue = [&](long long c) {  // lane 0 only
            const int j = (int)(c / cpw), q = (int)(c % cpw);
            const size_t off = (size_t)q * COV_CHUNK;
            const uint32_t bytes = (uint32_t)min((size_t)COV_CHUNK, win_bytes - off);
            const int slot = (int)(c % FZ_STAGES);
            const unsigned char *src = src0 + ((size_t)gw + (size_t)j * total_warps) * win_bytes + off;
            mbar_expect_tx(bar0 + 8 * slot, bytes);
            bulk_g2s(ring0 + slot * COV_CHUNK, src, bytes, bar0 + 8 * slot);
        };
        if (lane == 0)
            for (long long c = 0; c < total && c < FZ_STAGES; ++c) issue(c);
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0;
        int q = 0, j = 0, slot = 0;
        uint32_t parity = 0;
        for (long long c = 0; c < total; ++c) {
            while (!mbar_try_wait(bar0 + 8 * slot, parity)) {}
            const size_t off = (size_t)q * COV_CHUNK;
            const int nsnap = (int)(min((size_t)COV_CHUNK, win_bytes - off) >> 5);
            const float4 *buf = reinterpret_cast<const float4 *>(ring + (size_t)slot * COV_CHUNK);
            if (nsnap == COV_CHUNK / 32) {
                float4 xa[4], xb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    xa[u] = buf[2 * (lane + 32 * u)];
                    xb[u] = buf[2 * (lane + 32 * u) + 1];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) cov4_accumulate(acc, xa[u], xb[u]);
            } else {
                for (int s = lane; s < nsnap; s += 32) cov4_accumulate(acc, buf[2 * s], buf[2 * s + 1]);
            }
            __syncwarp();
            if (lane == 0 && c + FZ_STAGES < total) issue(c + FZ_STAGES);
            if (++slot == FZ_STAGES) { slot = 0; parity ^= 1; }
            if (++q == cpw) {
                q = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = warp_sum(acc[i]);
                if (lane == 0) {
                    const unsigned seq = atomicAdd(&ctl->cov_seq, 1u);
                    while (seq - ctl->scan_done >= (unsigned)FZ_Q) __nanosleep(100);  // queue slot free?
                    const double dn = (double)N;
                    double *Rw = Rq + (size_t)(seq % FZ_Q) * 32;
                    Rw[0] = acc[0] / dn;   Rw[1] = 0.0;
                    Rw[10] = acc[1] / dn;  Rw[11] = 0.0;
                    Rw[20] = acc[2] / dn;  Rw[21] = 0.0;
                    Rw[30] = acc[3] / dn;  Rw[31] = 0.0;
                    int e = 4;
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = a + 1; b < 4; ++b) {
                            const double re = acc[e] / dn, im = acc[e + 1] / dn;
                            Rw[2 * (a * 4 + b)] = re;  Rw[2 * (a * 4 + b) + 1] = im;
                            Rw[2 * (b * 4 + a)] = re;  Rw[2 * (b * 4 + a) + 1] = -im;
                            e += 2;
                        }
                    qwin[seq % FZ_Q] = gw + j * total_warps;
                    while (ctl->cov_pub != seq) {}  // publish in ticket order
                    __threadfence_block();
                    ctl->cov_pub = seq + 1;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0;
                ++j;
            }
        }
        if (lane == 0) atomicAdd((unsigned *)&ctl->cov_finished, 1u);
        if (dbg && lane == 0) dbg[blockIdx.x * 16 + warp] = clock64() - t_start;  // this covariance warp is done
    } else if (warp == FZ_COV_WARPS) {
        // ================= eigensolver warp =================
        long long eig_busy = 0, eig_rounds = 0;
        for (;;) {
            const unsigned done = ctl->eig_done;
            const unsigned avail = ctl->cov_pub - done;
            if (avail == 0) {
                if (ctl->cov_finished == (unsigned)FZ_COV_WARPS && ctl->cov_pub == done) break;
                __nanosleep(200);
                continue;
            }
            __threadfence_block();
            const long long t0 = clock64();
            const unsigned cnt = min(avail, 32u);
            if ((unsigned)lane < cnt) {
                const unsigned slot = (done + lane) % FZ_Q;
                herm_eig_body<4, true>(Rq + (size_t)slot * 32, nullptr, Vq + (size_t)slot * 32, 4);
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) ctl->eig_done = done + cnt;
            __syncwarp();
            eig_busy += clock64() - t0;
            ++eig_rounds;
        }
        if (dbg && lane == 0) {
            dbg[blockIdx.x * 16 + 8] = clock64() - t_start;
            dbg[blockIdx.x * 16 + 9] = eig_busy;
            dbg[blockIdx.x * 16 + 10] = eig_rounds;
        }
    } else {
        // ================= scan warps =================
        constexpr int M = 4, vsz = 32;
        const int st = threadIdx.x - 32 * (FZ_COV_WARPS + 1);  // 0..223
        const int swarp = st >> 5;
        double *redP = reinterpret_cast<double *>(fz_smem + FZ_OFF_RED);  // [7][4]
        int *redk = reinterpret_cast<int *>(fz_smem + FZ_OFF_RED + 8 * FZ_SCAN_WARPS * FZ_WPT);
        const uint32_t Vq0 = smem_u32(Vq);
        const int niter = (K + FZ_SCAN_THREADS - 1) / FZ_SCAN_THREADS;
        const uint32_t tb0 = smem_u32(fz_smem + FZ_OFF_TBAR), tbuf0 = smem_u32(fz_smem + FZ_OFF_TBL);
        unsigned T = 0;  // table tiles consumed so far (identical in every scan thread)
        long long scan_busy = 0, scan_passes = 0;
        for (;;) {
            if (st == 0) {
                unsigned start, cnt;
                for (;;) {
                    start = ctl->scan_done;
                    const unsigned avail = ctl->eig_done - start;
                    if (avail > 0) { cnt = min(avail, (unsigned)FZ_WPT); break; }
                    if (ctl->cov_finished == (unsigned)FZ_COV_WARPS && ctl->cov_pub == start) { cnt = 0; break; }
                    __nanosleep(200);
                }
                ctl->batch_start = start;
                ctl->batch_cnt = cnt;
                __threadfence_block();
            }
            bar_sync_scan();
            const unsigned start = ctl->batch_start, cnt = ctl->batch_cnt;
            if (cnt == 0) break;
            const long long t0 = clock64();
            uint32_t ev[FZ_WPT];  // shared address of each window's eigenvectors (clamped duplicates beyond cnt)
#pragma unroll
            for (int b = 0; b < FZ_WPT; ++b) ev[b] = Vq0 + 8 * vsz * ((start + min((unsigned)b, cnt - 1)) % FZ_Q);
            PeakState<FZ_WPT> ps;
            ps.reset();

            // Steering-table tiles stream through a FZ_TS-deep TMA ring shared by the scan warps (the
            // table lives in L2, but under the covariance warps' HBM stream an L2 hit costs > 1000
            // cycles - far more than one tile of arithmetic).  T counts tiles since kernel start:
            // slot = T % FZ_TS, tfull parity = (T / FZ_TS) & 1; a tile is released by one arrival per
            // scan warp on tempty, which the producer (thread 0 of the scan group) awaits before refill.
            for (int it = 0; it < niter; ++it, ++T) {
                const int slot = (int)(T % FZ_TS);
                if (st == 0) {
                    // keep the ring full: tiles it .. it+FZ_TS-1 of this pass (prologue at it == 0)
                    for (int a = (it == 0 ? 0 : FZ_TS - 1); a < FZ_TS; ++a) {
                        const int ia = it + a;
                        if (ia >= niter) break;
                        const unsigned Ta = T + a;
                        const int sa = (int)(Ta % FZ_TS);
                        if (Ta >= FZ_TS) while (!mbar_try_wait(tb0 + 8 * (FZ_TS + sa), (uint32_t)((Ta / FZ_TS - 1) & 1))) {}
                        mbar_expect_tx(tb0 + 8 * sa, FZ_TILE_BYTES);
                        bulk_g2s(tbuf0 + sa * FZ_TILE_BYTES, tbl + (size_t)ia * (FZ_TCOMP * FZ_SCAN_THREADS), FZ_TILE_BYTES, tb0 + 8 * sa);
                    }
                }
                while (!mbar_try_wait(tb0 + 8 * slot, (uint32_t)((T / FZ_TS) & 1))) {}
                const int k = it * FZ_SCAN_THREADS + st;
                const uint32_t row = tbuf0 + slot * FZ_TILE_BYTES + 8 * st;
                double ar[M], ai[M], na;
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ar[i]) : "r"(row + 8 * FZ_SCAN_THREADS * (2 * i)));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ai[i]) : "r"(row + 8 * FZ_SCAN_THREADS * (2 * i + 1)));
                }
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(na) : "r"(row + 8 * FZ_SCAN_THREADS * 8));
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tb0 + 8 * (FZ_TS + slot)) : "memory");
                scan_bin<M, FZ_WPT>(ar, ai, na, k, ev, ps);
            }
#pragma unroll
            for (int b = 0; b < FZ_WPT; ++b) {
                int kk = ps.bestk[b];
                double P = kk >= 0 ? 1.0 / ps.bestd[b] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double Po = __shfl_xor_sync(0xffffffffu, P, o);
                    const int ko = __shfl_xor_sync(0xffffffffu, kk, o);
                    if (peak_better(Po, ko, P, kk)) { P = Po; kk = ko; }
                }
                if (lane == 0) { redP[swarp * FZ_WPT + b] = P; redk[swarp * FZ_WPT + b] = kk; }
            }
            bar_sync_scan();
            if ((unsigned)st < cnt) {
                const int b = st;
                double P = redP[b];
                int kk = redk[b];
                for (int q = 1; q < FZ_SCAN_WARPS; ++q)
                    if (peak_better(redP[q * FZ_WPT + b], redk[q * FZ_WPT + b], P, kk)) { P = redP[q * FZ_WPT + b]; kk = redk[q * FZ_WPT + b]; }
                const size_t o = (size_t)qwin[(start + b) % FZ_Q];
                if (kk >= 0) {
                    out.angles[o] = (float)((double)kk * 360.0 / (double)K);  // reference :134, :153
                    if (out.levels) out.levels[o] = (float)P;                 // reference :154
                } else {
                    out.angles[o] = 0.f;                                      // (0,0) initial pair, reference :95
                    if (out.levels) out.levels[o] = 0.f;
                }
                if (out.bins) out.bins[o] = kk;
            }
            bar_sync_scan();  // red[] and the queue slots may be reused from here on
            if (st == 0) {
                __threadfence_block();
                ctl->scan_done = start + cnt;
            }
            scan_busy += clock64() - t0;
            ++scan_passes;
        }
        if (dbg && st == 0) {
            dbg[blockIdx.x * 16 + 11] = clock64() - t_start;
            dbg[blockIdx.x * 16 + 12] = scan_busy;
            dbg[blockIdx.x * 16 + 13] = scan_passes;
        }
    }
}

}  // namespace music
