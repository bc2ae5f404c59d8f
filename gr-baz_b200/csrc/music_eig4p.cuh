// music_eig4p.cuh - PRINCIPAL eigenvector of a 4 x 4 Hermitian covariance, four lanes per window (fused kernel, n = 1).
//
// For one source (n = 1) the reference needs from eig_sym (/root/reference/lib/baz_music_doa.cc:88-93) only the split
// "largest eigenvector | the other three": the noise subspace G = eigvec.cols(0, M-n-1) enters work() through
// ||G^H a||^2 alone (:110-119), and that is the same number for ANY orthonormal basis of the orthogonal complement of the
// principal eigenvector e.  A full Jacobi decomposition (herm_eig4_coop, ~26 k cycles per round of 8 windows, a chain of
// fp64 rsqrt / divisions) is therefore replaced on the fused path by
//   1. repeated squaring of the (power-of-two scaled) matrix, A <- A^2 / 2^k: the eigenvalue ratios square with every
//      step, so A becomes rank one (mu e e^H) to fp64 accuracy after log2(53 / log2(l1/l2)) steps - 3 at 20 dB, 5 at 0 dB.
//      Rank-one-ness is tested on ||A||_F^2 >= (1 - 1e-9) tr(A)^2 (i.e. sum_{i>1} mu_i / mu_1 <= 5e-10), after which
//      ONE more squaring pushes the residual components below 1e-18;
//   2. e = the column of A with the largest diagonal, then two power steps with the ORIGINAL (scaled) matrix, which
//      wash out the rounding noise accumulated by the squarings (each step shrinks it by l2/l1) - the result agrees with
//      LAPACK zheevd to <= 2e-15 down to -10 dB SNR (tools/emulate_eig4_principal.py);
//   3. a residual certificate ||A0 e - lambda e|| <= 1e-12 lambda;
//   4. the phase that makes e_0 real and >= 0 (the scan kernels rely on it) and the Householder reflector that maps the
//      first unit vector onto -e: its columns 1..3,  g_j = u_j - conj(e_j) / (1 + e_0) * (e + u_0),  are an orthonormal
//      basis of the complement of e (1 + e_0 is in [1, 2]: no cancellation), stored where the scan expects the three
//      noise eigenvectors.
// Every scaling is an exact power of two taken from the exponent of the trace, so a window scaled by 2^k gives bit-identical
// vectors.  A window is frozen as soon as it has converged: its result does not depend on the other windows of the warp.
// Windows that do not converge within EIGP_MAXSQ squarings (eigenvalue ratio < ~1.02: noise-only input), fail the
// certificate or hold NaN/Inf/zero report false and go through the Jacobi solver (herm_eig4_coop) instead.
#pragma once
#include "music_kernels.cuh"

namespace music {

constexpr int EIGP_MAXSQ = 12;

// 2^-floor(log2 t) for a positive, normal, finite t (ok = false otherwise)
__device__ __forceinline__ double eigp_pow2_scale(const double t, bool &ok)
{
    const int hi = __double2hiint(t);
    const int e = (hi >> 20) & 0x7ff;
    const int se = 2046 - e;
    ok = hi > 0 && e > 0 && e < 0x7ff && se > 0;
    return __hiloint2double((ok ? se : 1023) << 20, 0);
}

__device__ __forceinline__ double eigp_sum4(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// All 32 lanes call this together; lane group g = lane >> 2 works on one window, j = lane & 3 is the matrix row the lane
// holds.  Rw: 4 x 4 complex, row-major interleaved (shared memory); vw: 16 complex of shared scratch that receives
// Vt[rank][i] like herm_eig_body (ranks 0..2 = complement basis, rank 3 = principal eigenvector).  Returns true for the
// lanes of a window that was solved here; false -> the caller runs the Jacobi solver for that window.
__device__ __forceinline__ bool eig4_principal_coop(const double *Rw, double *vw, const bool active, const int j)
{
    constexpr unsigned FULL = 0xffffffffu;
    double2 *S = reinterpret_cast<double2 *>(vw);  // scratch: S[row * 4 + col]
    double a0r[4], a0i[4], ar[4], ai[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a0r[c] = active ? Rw[2 * (j * 4 + c)] : 0.0;
        a0i[c] = active ? Rw[2 * (j * 4 + c) + 1] : 0.0;
    }
    const double dj = j == 0 ? a0r[0] : j == 1 ? a0r[1] : j == 2 ? a0r[2] : a0r[3];  // own diagonal
    bool okc;
    const double sc0 = eigp_pow2_scale(eigp_sum4(dj), okc);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a0r[c] *= sc0; a0i[c] *= sc0;
        ar[c] = a0r[c]; ai[c] = a0i[c];
    }
    // st: 0 squaring, 1 rank-one test passed (one more squaring to go), 2 converged, 3 failed / not participating
    int st = (active && okc) ? 0 : 3;
    for (int it = 0; it < EIGP_MAXSQ; ++it) {
        if (!__any_sync(FULL, st < 2)) break;
        if (active) {  // (idle lane groups alias the last window's scratch: they must not store)
#pragma unroll
            for (int c = 0; c < 4; ++c) S[j * 4 + c] = make_double2(ar[c], ai[c]);
        }
        __syncwarp();
        double nr[4], ni[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double xr = 0.0, xi = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 b = S[k * 4 + c];
                xr = fma(ar[k], b.x, xr); xr = fma(-ai[k], b.y, xr);
                xi = fma(ar[k], b.y, xi); xi = fma(ai[k], b.x, xi);
            }
            nr[c] = xr; ni[c] = (c == j) ? 0.0 : xi;
        }
        __syncwarp();
        const double ndj = j == 0 ? nr[0] : j == 1 ? nr[1] : j == 2 ? nr[2] : nr[3];
        double fj = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) fj = fma(nr[c], nr[c], fma(ni[c], ni[c], fj));
        const double t = eigp_sum4(ndj), f = eigp_sum4(fj);
        bool oks;
        const double sc = eigp_pow2_scale(t, oks);
        const bool pass = f >= 0.999999999 * (t * t);  // false for NaN
        if (st < 2) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { ar[c] = nr[c] * sc; ai[c] = ni[c] * sc; }
            st = !oks ? 3 : (st == 1 ? 2 : (pass ? 1 : 0));
        }
    }
    bool ok = st == 2;
    // the column with the largest diagonal (ties: lowest index) is mu e conj(e_j*): take it from the row of lane j*
    // (row j* = conj of column j*, the matrix is Hermitian)
    const double d0 = __shfl_sync(FULL, ar[0], (threadIdx.x & 28) | 0), d1 = __shfl_sync(FULL, ar[1], (threadIdx.x & 28) | 1);
    const double d2 = __shfl_sync(FULL, ar[2], (threadIdx.x & 28) | 2), d3 = __shfl_sync(FULL, ar[3], (threadIdx.x & 28) | 3);
    int js = 0;
    double dm = d0;
    if (d1 > dm) { dm = d1; js = 1; }
    if (d2 > dm) { dm = d2; js = 2; }
    if (d3 > dm) { dm = d3; js = 3; }
    if (active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) S[j * 4 + c] = make_double2(ar[c], ai[c]);
    }
    __syncwarp();
    double ur[4], ui[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double2 b = S[js * 4 + c];
        ur[c] = b.x; ui[c] = -b.y;
    }
    __syncwarp();
    // two power steps with the original (scaled) matrix, then one more product for the certificate
    double wr = 0.0, wi = 0.0, lam = 0.0;
#pragma unroll 1  // (one copy of the step: instruction footprint, see music_fused.cuh)
    for (int step = 0; step < 3; ++step) {
        wr = 0.0; wi = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            wr = fma(a0r[c], ur[c], wr); wr = fma(-a0i[c], ui[c], wr);
            wi = fma(a0r[c], ui[c], wi); wi = fma(a0i[c], ur[c], wi);
        }
        const double urj = j == 0 ? ur[0] : j == 1 ? ur[1] : j == 2 ? ur[2] : ur[3];
        const double uij = j == 0 ? ui[0] : j == 1 ? ui[1] : j == 2 ? ui[2] : ui[3];
        if (step == 2) {
            // ||A0 e - lambda e||^2 <= 1e-24 lambda^2,  lambda = e^H A0 e
            lam = eigp_sum4(fma(urj, wr, uij * wi));
            const double rr = fma(-lam, urj, wr), ri = fma(-lam, uij, wi);
            const double res2 = eigp_sum4(fma(rr, rr, ri * ri));
            ok = ok && (res2 <= 1e-24 * (lam * lam)) && lam > 0.0;  // false for NaN
            break;
        }
        const double s2 = eigp_sum4(fma(wr, wr, wi * wi));
        const double inv = 1.0 / sqrt(s2);
        if (active) S[j] = make_double2(wr * inv, wi * inv);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) { ur[c] = S[c].x; ui[c] = S[c].y; }
        __syncwarp();
    }
    // phase: component 0 real and >= 0
    double pr, pi;
    eig_phase(ur[0], ui[0], pr, pi);
    double er[4], ei[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) eig_out4(ur[c], ui[c], pr, pi, er[c], ei[c]);
    ei[0] = 0.0;
    if (ok) {
        const double erj = j == 0 ? er[0] : j == 1 ? er[1] : j == 2 ? er[2] : er[3];
        const double eij = j == 0 ? ei[0] : j == 1 ? ei[1] : j == 2 ? ei[2] : ei[3];
        vw[2 * (3 * 4 + j)] = erj;
        vw[2 * (3 * 4 + j) + 1] = eij;
        const double h = 1.0 / (1.0 + er[0]);
#pragma unroll
        for (int q = 1; q < 4; ++q) {  // g_q[j], stored as rank q - 1
            double gr, gi;
            if (j == 0) {
                gr = -er[q]; gi = ei[q];                     // -conj(e_q)
            } else {
                // delta_jq - e_j conj(e_q) / (1 + e_0)
                const double pr2 = fma(erj, er[q], eij * ei[q]), pi2 = fma(eij, er[q], -(erj * ei[q]));
                gr = fma(-pr2, h, j == q ? 1.0 : 0.0);
                gi = -(pi2 * h);
            }
            vw[2 * ((q - 1) * 4 + j)] = gr;
            vw[2 * ((q - 1) * 4 + j) + 1] = gi;
        }
    }
    __syncwarp();
    return ok;
}

}  // namespace music
