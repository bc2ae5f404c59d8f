// microbench.cu - B200 pipe rates that decide the MUSIC kernel design:
//   DFMA issue rate, F2F.F64.F32 rate, whether they overlap, an integer-ALU f32->f64 widening,
//   and DMMA (mma.sync m8n8k4 f64).  Prints lane-ops per clock per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double *out, double a, double b, long long *cyc)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH>
__global__ void k_dfma_chains(double *out, double a, double b, long long *cyc)
{
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 8 / CH; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_f2f(double *out, int seed, long long *cyc)
{
    int v[8];
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0x3f800000 + threadIdx.x * 8 + i + seed;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double d = (double)__int_as_float(v[i]);
            acc ^= __double2loint(d) ^ __double2hiint(d);
            v[i] += 0x101;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// R DFMA per F2F (the covariance kernel has M DFMA per conversion: R = 4, 8, 16)
template <int R>
__global__ void k_mix(double *out, int seed, double b, long long *cyc)
{
    int v[2];
    double acc[2 * R];
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) acc[i] = threadIdx.x + i;
    v[0] = 0x3f800000 + threadIdx.x + seed;
    v[1] = 0x3f900000 + threadIdx.x + seed;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double d = (double)__int_as_float(v[j]);
            v[j] += 0x101;
#pragma unroll
            for (int i = 0; i < R; ++i) acc[j * R + i] = fma(acc[j * R + i], d, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ double widen_alu(int x)
{   // exact f32 -> f64 for normal numbers with integer ops only
    unsigned t = (unsigned)x & 0x7fffffffu;
    unsigned hi = (t >> 3) + 0x38000000u;
    hi |= (unsigned)x & 0x80000000u;
    unsigned lo = (unsigned)x << 29;
    return __hiloint2double((int)hi, (int)lo);
}

template <int R>
__global__ void k_mix_alu(double *out, int seed, double b, long long *cyc)
{
    int v[2];
    double acc[2 * R];
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) acc[i] = threadIdx.x + i;
    v[0] = 0x3f800000 + threadIdx.x + seed;
    v[1] = 0x3f900000 + threadIdx.x + seed;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double d = widen_alu(v[j]);
            v[j] += 0x101;
#pragma unroll
            for (int i = 0; i < R; ++i) acc[j * R + i] = fma(acc[j * R + i], d, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * R; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// legacy mma.sync TF32 m16n8k8 (SASS HMMA.1688.F32.TF32): CH independent accumulators per warp
template <int CH>
__global__ void k_hmma_tf32(float *out, long long *cyc)
{
    float c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1.f; c[i][3] = 2.f; }
    unsigned a0 = 0x3f800000u + threadIdx.x * 8192u, a1 = a0 + 8192u, a2 = a0 + 16384u, a3 = a0 + 24576u, b0 = 0x3f000000u, b1 = 0x3f002000u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dmma(double *out, double a, double b, long long *cyc)
{
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char *name, F launch, double lane_ops_per_thread_iter, int threads, int ctas_per_sm, int sms, long long *d_cyc)
{
    int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(grid, threads);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(grid, threads);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = new long long[grid];
    cudaMemcpy(h, d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    double ops_per_sm = lane_ops_per_thread_iter * ITERS * threads * ctas_per_sm;
    printf("%-28s threads/SM %4d  %8.1f lane-ops/clk/SM (clock64)  %8.2f Tops/s (event, %.3f ms)  err=%s\n", name,
           threads * ctas_per_sm, ops_per_sm / avg, ops_per_sm * sms / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
    delete[] h;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, sm_%d%d\n", p.name, sms, p.major, p.minor);
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms * 8);
    // dependent-issue latency of DFMA: one warp per SMSP, CH independent chains -> lane-ops/clk/SM = 128*CH/L... (L = 4*32*CH/rate)
    run("DFMA 1 chain, 1 warp/SMSP", [&](int g, int t) { k_dfma_chains<1><<<g, t>>>(out, 1.0000001, 1e-9, cyc); }, 8, 128, 1, sms, cyc);
    run("DFMA 2 chains, 1 warp/SMSP", [&](int g, int t) { k_dfma_chains<2><<<g, t>>>(out, 1.0000001, 1e-9, cyc); }, 8, 128, 1, sms, cyc);
    run("DFMA 4 chains, 1 warp/SMSP", [&](int g, int t) { k_dfma_chains<4><<<g, t>>>(out, 1.0000001, 1e-9, cyc); }, 8, 128, 1, sms, cyc);
    run("DFMA 8 chains, 1 warp/SMSP", [&](int g, int t) { k_dfma_chains<8><<<g, t>>>(out, 1.0000001, 1e-9, cyc); }, 8, 128, 1, sms, cyc);
    // HMMA.1688.F32.TF32: "lane-ops" here = MMA instructions * 32, so MMAs/clk/SM = value / 32
    run("HMMA tf32 1 chain 1w/SMSP", [&](int g, int t) { k_hmma_tf32<1><<<g, t>>>((float *)out, cyc); }, 1, 128, 1, sms, cyc);
    run("HMMA tf32 4 chains 1w/SMSP", [&](int g, int t) { k_hmma_tf32<4><<<g, t>>>((float *)out, cyc); }, 4, 128, 1, sms, cyc);
    run("HMMA tf32 8 chains 1w/SMSP", [&](int g, int t) { k_hmma_tf32<8><<<g, t>>>((float *)out, cyc); }, 8, 128, 1, sms, cyc);
    run("HMMA tf32 8 chains 4w/SMSP", [&](int g, int t) { k_hmma_tf32<8><<<g, t>>>((float *)out, cyc); }, 8, 512, 1, sms, cyc);
    for (int tpb : {128, 1024}) {
        run("DFMA", [&](int g, int t) { k_dfma<<<g, t>>>(out, 1.0000001, 1e-9, cyc); }, 8, tpb, 1, sms, cyc);
        run("F2F.F64.F32", [&](int g, int t) { k_f2f<<<g, t>>>(out, 1, cyc); }, 8, tpb, 1, sms, cyc);
        run("DFMA:F2F 4:1 (count DFMA)", [&](int g, int t) { k_mix<4><<<g, t>>>(out, 1, 1e-9, cyc); }, 8, tpb, 1, sms, cyc);
        run("DFMA:F2F 8:1 (count DFMA)", [&](int g, int t) { k_mix<8><<<g, t>>>(out, 1, 1e-9, cyc); }, 16, tpb, 1, sms, cyc);
        run("DFMA:ALUwiden 4:1 (DFMA)", [&](int g, int t) { k_mix_alu<4><<<g, t>>>(out, 1, 1e-9, cyc); }, 8, tpb, 1, sms, cyc);
        run("DMMA m8n8k4 (FMA lanes)", [&](int g, int t) { k_dmma<<<g, t>>>(out, 1.0000001, 1e-9, cyc); }, 4 * 8.0, tpb, 1, sms, cyc);
    }
    return 0;
}
