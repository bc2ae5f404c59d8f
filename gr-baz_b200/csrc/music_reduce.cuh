// music_reduce.cuh - downstream reducers.
//
// The consumers of the block in gr-baz are GUI-rate sinks: the compass takes one angle per update
// (/root/reference/python/doa_compass_control.py:102-108, set_direction) and the plot sink one spectrum
// (/root/reference/python/plot_sink.py:38).  At millions of windows per second the sensible adapter is a reduction on
// the device: the CIRCULAR mean of each reported angle slot over a batch of windows (an arithmetic mean is wrong
// across the 0/360 wrap) together with the mean resultant length (1 = all windows agree, 0 = uniformly spread), and
// the mean pseudospectrum.  No reference counterpart; definition in oracle/music_oracle.py::reduce_angles /
// reduce_spectrum.  Deterministic: fixed-order tree reductions, no atomics.
#pragma once
#include "music_kernels.cuh"

namespace music {

constexpr int REDUCE_THREADS = 1024;

// One CTA per angle slot i < n.  A window counts iff levels == nullptr or levels[w][i] > 0 (the block leaves
// (0, 0) in slots it could not fill); weight = levels[w][i] if weighted else 1.
__global__ void __launch_bounds__(REDUCE_THREADS) reduce_angles_kernel(const float *__restrict__ angles, const float *__restrict__ levels,
                                                                      int W, int n, int weighted, float *__restrict__ mean_deg,
                                                                      float *__restrict__ resultant, float *__restrict__ weight_sum)
{
    __shared__ double sh[3][REDUCE_THREADS];
    const int i = blockIdx.x, t = threadIdx.x;
    double s = 0.0, c = 0.0, wsum = 0.0;
    for (int w = t; w < W; w += REDUCE_THREADS) {
        const double lv = levels ? (double)levels[(size_t)w * n + i] : 1.0;
        if (!(lv > 0.0)) continue;  // unfilled slot (or NaN level)
        const double wt = weighted ? lv : 1.0;
        double sn, cs;
        sincos((double)angles[(size_t)w * n + i] * (3.14159265358979323846 / 180.0), &sn, &cs);
        s = fma(wt, sn, s);
        c = fma(wt, cs, c);
        wsum += wt;
    }
    sh[0][t] = s; sh[1][t] = c; sh[2][t] = wsum;
    __syncthreads();
    for (int o = REDUCE_THREADS / 2; o > 0; o >>= 1) {
        if (t < o) {
            sh[0][t] += sh[0][t + o];
            sh[1][t] += sh[1][t + o];
            sh[2][t] += sh[2][t + o];
        }
        __syncthreads();
    }
    if (t == 0) {
        const double S = sh[0][0], C = sh[1][0], Wt = sh[2][0];
        double deg = 0.0, r = 0.0;
        if (Wt > 0.0) {
            deg = atan2(S, C) * (180.0 / 3.14159265358979323846);
            if (deg < 0.0) deg += 360.0;
            r = sqrt(fma(S, S, C * C)) / Wt;
        }
        mean_deg[i] = (float)deg;
        if (resultant) resultant[i] = (float)r;
        if (weight_sum) weight_sum[i] = (float)Wt;
    }
}

// mean over W windows of spectrum[w][k]; thread <-> bin (coalesced rows), fp64 accumulation in window order
__global__ void __launch_bounds__(256) reduce_spectrum_kernel(const float *__restrict__ spectrum, int W, int K, float *__restrict__ mean)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double acc = 0.0;
    for (int w = 0; w < W; ++w) acc += (double)spectrum[(size_t)w * K + k];
    mean[k] = (float)(acc / (double)W);
}

}  // namespace music
