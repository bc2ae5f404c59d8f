/*
 * CPU oracle (plain C, fp64) for gr-baz's MUSIC DOA block.  TEST INFRASTRUCTURE ONLY.
 *
 * A restatement of the reference algorithm, used (a) as the parity checker for the CUDA
 * path and (b) as the timed host baseline ("port").  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing in the product
 * path (gr-baz_b200/, lib/) links or calls this file.
 *
 * PARITY UNPINNED BY THE REFERENCE: it has no tests or golden vectors for this block and
 * cannot be compiled here (needs GNU Radio + Armadillo + LAPACK, all absent).  Armadillo's
 * eig_sym (LAPACK zheev/zheevd, unpinned version) is restated as a cyclic complex Jacobi
 * eigensolver; MUSIC only uses the noise-subspace projector, which is independent of the
 * eigenvector phases / degenerate-subspace rotations that differ between LAPACK and
 * Jacobi.  tests/test_oracle.py checks this file against the numpy/LAPACK restatement
 * (oracle/music_oracle.py) and against tests/golden.
 *
 * Reference lines followed (all under /root/reference):
 *   lib/baz_music_doa.cc:72-161   work()       -> music_oracle_work()
 *   lib/baz_music_doa.cc:74-77    new[] + widen c64 -> c128 per call (kept: it is part of
 *                                 what the reference costs per window)
 *   lib/baz_music_doa.cc:82-85    x(r,c) = in[c*M + r];  R = x x^H / N
 *   lib/baz_music_doa.cc:88-93    eigenvalues ascending; G = first M-n eigenvectors
 *   lib/baz_music_doa.cc:103-121  per step: widen a, 1/pow(norm(G^H a, 2), 2), spectrum
 *   lib/baz_music_doa.cc:129-141  top-n insertion, strict '>'
 *   lib/baz_music_doa.cc:146-155  float casts of angle / level
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "music_oracle.h"

/* ------------------------------------------------------------------------------------
 * Hermitian eigensolver: cyclic two-sided Jacobi, fp64.
 *   A: M x M Hermitian, row-major, interleaved (re, im); destroyed.
 *   V: M x M, columns are orthonormal eigenvectors on return (same layout).
 *   w: M eigenvalues, ASCENDING (stable w.r.t. column index on ties), V permuted to match.
 * Rotation J = [[c, s*e],[-s*conj(e), c]], e = a_pq/|a_pq|, t = s/c the smaller root of
 * t^2 + 2*theta*t - 1 = 0, theta = (a_qq - a_pp) / (2|a_pq|).
 * ------------------------------------------------------------------------------------ */
#define AR(i, j) A[2 * ((i) * M + (j))]
#define AI(i, j) A[2 * ((i) * M + (j)) + 1]
#define VR(i, j) V[2 * ((i) * M + (j))]
#define VI(i, j) V[2 * ((i) * M + (j)) + 1]

int music_oracle_herm_eig(unsigned M, double *A, double *V, double *w)
{
    unsigned i, j, k, p, q;
    int sweep;
    for (i = 0; i < M; i++)
        for (j = 0; j < M; j++) {
            VR(i, j) = (i == j) ? 1.0 : 0.0;
            VI(i, j) = 0.0;
        }
    for (sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, fro = 0.0;
        for (i = 0; i < M; i++)
            for (j = 0; j < M; j++) {
                double e2 = AR(i, j) * AR(i, j) + AI(i, j) * AI(i, j);
                fro += e2;
                if (i != j) off += e2;
            }
        if (off <= 1e-32 * fro || off == 0.0) break;
        for (p = 0; p + 1 < M; p++)
            for (q = p + 1; q < M; q++) {
                double gr = AR(p, q), gi = AI(p, q);
                double g = sqrt(gr * gr + gi * gi);
                double app = AR(p, p), aqq = AR(q, q);
                double theta, t, c, s, er, ei, swr, swi;
                if (g == 0.0) continue;
                er = gr / g;
                ei = gi / g;
                theta = (aqq - app) / (2.0 * g);
                t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (theta < 0.0) t = -t;
                c = 1.0 / sqrt(t * t + 1.0);
                s = t * c;
                swr = s * er; /* s*e */
                swi = s * ei;
                for (k = 0; k < M; k++) {
                    double kpr, kpi, kqr, kqi;
                    if (k == p || k == q) continue;
                    kpr = AR(k, p); kpi = AI(k, p);
                    kqr = AR(k, q); kqi = AI(k, q);
                    /* a_kp' = c a_kp - conj(s e) a_kq ;  a_kq' = (s e) a_kp + c a_kq */
                    AR(k, p) = c * kpr - (swr * kqr + swi * kqi);
                    AI(k, p) = c * kpi - (swr * kqi - swi * kqr);
                    AR(k, q) = c * kqr + (swr * kpr - swi * kpi);
                    AI(k, q) = c * kqi + (swr * kpi + swi * kpr);
                    AR(p, k) = AR(k, p); AI(p, k) = -AI(k, p);
                    AR(q, k) = AR(k, q); AI(q, k) = -AI(k, q);
                }
                AR(p, p) = app - t * g; AI(p, p) = 0.0;
                AR(q, q) = aqq + t * g; AI(q, q) = 0.0;
                AR(p, q) = 0.0; AI(p, q) = 0.0;
                AR(q, p) = 0.0; AI(q, p) = 0.0;
                for (k = 0; k < M; k++) {
                    double kpr = VR(k, p), kpi = VI(k, p), kqr = VR(k, q), kqi = VI(k, q);
                    VR(k, p) = c * kpr - (swr * kqr + swi * kqi);
                    VI(k, p) = c * kpi - (swr * kqi - swi * kqr);
                    VR(k, q) = c * kqr + (swr * kpr - swi * kpi);
                    VI(k, q) = c * kqi + (swr * kpi + swi * kpr);
                }
            }
    }
    /* ascending, stable insertion sort of (w, column) */
    {
        unsigned *perm = (unsigned *)malloc(M * sizeof(unsigned));
        double *tmp = (double *)malloc(2 * M * M * sizeof(double));
        if (!perm || !tmp) { free(perm); free(tmp); return -1; }
        for (i = 0; i < M; i++) perm[i] = i;
        for (i = 1; i < M; i++) {
            unsigned pi = perm[i];
            double wi = AR(pi, pi);
            j = i;
            while (j > 0 && AR(perm[j - 1], perm[j - 1]) > wi) { perm[j] = perm[j - 1]; j--; }
            perm[j] = pi;
        }
        memcpy(tmp, V, 2 * M * M * sizeof(double));
        for (j = 0; j < M; j++) {
            w[j] = AR(perm[j], perm[j]);
            for (i = 0; i < M; i++) {
                VR(i, j) = tmp[2 * (i * M + perm[j])];
                VI(i, j) = tmp[2 * (i * M + perm[j]) + 1];
            }
        }
        free(perm);
        free(tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * One window: lib/baz_music_doa.cc:72-161.
 * ------------------------------------------------------------------------------------ */
int music_oracle_work(const float *in_c64, unsigned m, unsigned n, unsigned nsamples,
                      const float *table_c64, unsigned resolution,
                      float *out_angle, float *out_level, float *out_spectrum,
                      int32_t *out_bins, double *out_P, double *out_R, double *out_eigvals,
                      double *out_eigvec)
{
    unsigned i, r, c, step, average_over;
    double *data, *R, *V, *w, *a, *doa_angle, *doa_strength;
    int32_t *doa_bin;
    if (!(m > 0 && n >= 1 && n < m && nsamples > 0 && (nsamples % m) == 0 && resolution > 0))
        return -2;
    average_over = nsamples / m; /* :83 */

    /* :75-77  per-call allocation and c64 -> c128 widening ("FIXME: Move outside") */
    data = (double *)malloc(2 * (size_t)nsamples * sizeof(double));
    R = (double *)calloc(2 * (size_t)m * m, sizeof(double));
    V = (double *)malloc(2 * (size_t)m * m * sizeof(double));
    w = (double *)malloc(m * sizeof(double));
    a = (double *)malloc(2 * (size_t)m * sizeof(double));
    doa_angle = (double *)calloc(n, sizeof(double));    /* :95 vDOAs(n, (0,0)) */
    doa_strength = (double *)calloc(n, sizeof(double));
    doa_bin = (int32_t *)malloc(n * sizeof(int32_t));
    if (!data || !R || !V || !w || !a || !doa_angle || !doa_strength || !doa_bin) {
        free(data); free(R); free(V); free(w); free(a); free(doa_angle); free(doa_strength); free(doa_bin);
        return -1;
    }
    for (i = 0; i < 2 * nsamples; i++) data[i] = (double)in_c64[i];
    for (i = 0; i < n; i++) doa_bin[i] = -1;

    /* :82-85  x(r,c) = data[c*m + r];  R = x x^H / N  (full matrix, as zgemm would) */
    for (c = 0; c < average_over; c++) {
        const double *x = data + 2 * (size_t)c * m;
        for (r = 0; r < m; r++) {
            double xr = x[2 * r], xi = x[2 * r + 1];
            for (i = 0; i < m; i++) {
                double yr = x[2 * i], yi = x[2 * i + 1]; /* x_r * conj(x_i) */
                R[2 * (r * m + i)] += xr * yr + xi * yi;
                R[2 * (r * m + i) + 1] += xi * yr - xr * yi;
            }
        }
    }
    for (i = 0; i < 2 * m * m; i++) R[i] /= (double)average_over;
    if (out_R) memcpy(out_R, R, 2 * (size_t)m * m * sizeof(double));

    /* :88-90 eig_sym -> ascending eigenvalues, eigenvectors in columns */
    if (music_oracle_herm_eig(m, R, V, w) != 0) {
        free(data); free(R); free(V); free(w); free(a); free(doa_angle); free(doa_strength); free(doa_bin);
        return -1;
    }
    if (out_eigvals) memcpy(out_eigvals, w, m * sizeof(double));
    if (out_eigvec) memcpy(out_eigvec, V, 2 * (size_t)m * m * sizeof(double));

    /* :93  G = eigvec.cols(0, m-n-1);  :103-141 scan */
    for (step = 0; step < resolution; step++) {
        const float *ar = table_c64 + 2 * (size_t)step * m;
        double acc = 0.0, nrm, strength;
        unsigned g;
        for (i = 0; i < 2 * m; i++) a[i] = (double)ar[i]; /* :110-112 */
        for (g = 0; g < m - n; g++) { /* (G^H a)_g = sum_i conj(V[i][g]) a_i */
            double vr = 0.0, vi = 0.0, mag;
            for (i = 0; i < m; i++) {
                double er = V[2 * (i * m + g)], ei = V[2 * (i * m + g) + 1];
                vr += er * a[2 * i] + ei * a[2 * i + 1];
                vi += er * a[2 * i + 1] - ei * a[2 * i];
            }
            mag = hypot(vr, vi); /* arma::norm(.,2) on complex: |v_g| via std::abs, then squared */
            acc += mag * mag;
        }
        nrm = sqrt(acc);
        strength = 1.0 / (nrm * nrm); /* 1.0 / pow(norm, 2) */
        if (out_spectrum) out_spectrum[step] = (float)strength; /* :120-121 */
        if (out_P) out_P[step] = strength;
        for (i = 0; i < n; i++) { /* :129-141 */
            if (strength > doa_strength[i]) {
                unsigned j;
                for (j = n - 1; j > i; j--) { /* insert at i, pop_back */
                    doa_angle[j] = doa_angle[j - 1];
                    doa_strength[j] = doa_strength[j - 1];
                    doa_bin[j] = doa_bin[j - 1];
                }
                doa_angle[i] = (double)step * 360.0 / (double)resolution; /* :134 */
                doa_strength[i] = strength;
                doa_bin[i] = (int32_t)step;
                break;
            }
        }
    }
    for (i = 0; i < n; i++) { /* :150-155 */
        out_angle[i] = (float)doa_angle[i];
        if (out_level) out_level[i] = (float)doa_strength[i];
        if (out_bins) out_bins[i] = doa_bin[i];
    }
    free(data); free(R); free(V); free(w); free(a); free(doa_angle); free(doa_strength); free(doa_bin);
    return 0; /* the reference returns 1 = one item produced, :160 */
}

/* W windows, one work() call each (the reference processes one item per call). */
int music_oracle_work_batch(const float *in_c64, unsigned nwindows, unsigned m, unsigned n,
                            unsigned nsamples, const float *table_c64, unsigned resolution,
                            float *out_angle, float *out_level, float *out_spectrum,
                            int32_t *out_bins, double *out_P)
{
    unsigned w;
    for (w = 0; w < nwindows; w++) {
        int rc = music_oracle_work(in_c64 + 2 * (size_t)w * nsamples, m, n, nsamples, table_c64, resolution,
                                   out_angle + (size_t)w * n, out_level ? out_level + (size_t)w * n : NULL,
                                   out_spectrum ? out_spectrum + (size_t)w * resolution : NULL,
                                   out_bins ? out_bins + (size_t)w * n : NULL,
                                   out_P ? out_P + (size_t)w * resolution : NULL, NULL, NULL, NULL);
        if (rc != 0) return rc;
    }
    return 0;
}
