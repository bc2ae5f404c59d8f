/* CPU oracle for gr-baz MUSIC DOA - TEST INFRASTRUCTURE ONLY (see music_oracle.c). */
#ifndef MUSIC_ORACLE_H
#define MUSIC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Hermitian eigendecomposition (cyclic Jacobi); A destroyed; eigenvalues ascending. */
int music_oracle_herm_eig(unsigned M, double *A, double *V, double *w);

/* One window, restating /root/reference/lib/baz_music_doa.cc:72-161.
 * in_c64: nsamples interleaved (re,im) floats; table_c64: [resolution][m] (re,im) floats.
 * out_level/out_spectrum/out_bins/out_P/out_R/out_eigvals/out_eigvec may be NULL.
 * Returns 0, -1 on allocation failure, -2 on invalid parameters. */
int music_oracle_work(const float *in_c64, unsigned m, unsigned n, unsigned nsamples,
                      const float *table_c64, unsigned resolution,
                      float *out_angle, float *out_level, float *out_spectrum,
                      int32_t *out_bins, double *out_P, double *out_R, double *out_eigvals,
                      double *out_eigvec);

int music_oracle_work_batch(const float *in_c64, unsigned nwindows, unsigned m, unsigned n,
                            unsigned nsamples, const float *table_c64, unsigned resolution,
                            float *out_angle, float *out_level, float *out_spectrum,
                            int32_t *out_bins, double *out_P);
#ifdef __cplusplus
}
#endif
#endif
