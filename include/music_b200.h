/*
 * music_b200.h - C ABI of the B200-native (sm_100a) MUSIC direction-of-arrival hot path.
 *
 * This is the drop-in boundary for gr-baz's `baz_music_doa` block: everything that
 * /root/reference/lib/baz_music_doa.cc does per window inside work() (lines 72-161) runs
 * behind these entry points as hand-written CUDA; the GNU Radio block (lib/baz_music_doa.cc
 * in this repo), the Python helper and the bench harness are the only callers.
 *
 * Plain C types only: pointers and sizes, no C++/torch types, no exceptions.  Every function
 * returns 0 on success or a negative MUSIC_B200_E* code; the message is available from
 * music_b200_last_error().  There is NO CPU fallback: create() fails if the device is not a
 * compute-capability 10.x GPU.
 *
 * Data layouts (identical to the reference's GNU Radio item layouts):
 *   input window : nsamples complex64 (re, im floats), antennas sample-interleaved,
 *                  x(r, c) = in[c*m + r]            (reference lib/baz_music_doa.cc:37, :82-84)
 *   array response table : [resolution][m] complex64  (reference lib/baz_music_doa.h:32-33,
 *                  marshalled by swig/baz_swig.i:564)
 *   angles / levels : n floats per window             (reference lib/baz_music_doa.cc:38, :146-155)
 *   spectrum        : resolution floats per window    (reference lib/baz_music_doa.cc:38, :120-121)
 *   bins            : n int32 per window - the integer peak-bin index k behind each angle
 *                     (angle = (float)(k*360.0/resolution), :134); -1 where the reference
 *                     would leave its (0, 0) initial pair in place (:95).
 */
#ifndef MUSIC_B200_H
#define MUSIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUSIC_B200_OK 0
#define MUSIC_B200_EINVAL (-1)   /* bad argument (same conditions as the reference's ctor asserts) */
#define MUSIC_B200_ECUDA (-2)    /* CUDA runtime error */
#define MUSIC_B200_ENODEVICE (-3) /* no sm_100-class device / device index out of range */
#define MUSIC_B200_ENOMEM (-4)

#define MUSIC_B200_MAX_M 16 /* antennas supported by this build */

typedef struct music_b200 music_b200;

/* ABI version of this header (bumped on any signature change). */
int music_b200_version(void);  /* 2: multi-device handles, fused bins all-gather */

/*
 * Replaces baz_make_music_doa() + the constructor
 * (/root/reference/lib/baz_music_doa.cc:29-53, lib/baz_music_doa.h:36-43).
 * Argument checks are the reference's asserts (:45-50) made real, plus 1 <= n < m
 * (m == n underflows eigvec.cols(0, m-n-1) at :93) and m <= MUSIC_B200_MAX_M.
 * table_c64: [resolution][m] interleaved (re, im) floats, copied (the block owns its table,
 * :42).  device: CUDA device ordinal.
 */
int music_b200_create(music_b200 **out, uint32_t m, uint32_t n, uint32_t nsamples,
                      uint32_t resolution, const float *table_c64, int device);

/*
 * The same block spread over `ndev` GPUs of one node (windows are independent, so they shard
 * round-robin, window w -> devices[w mod ndev], and every device keeps a full copy of the array response).  The handle
 * behaves like one from music_b200_create():
 *   process_host()     deals the call's windows round-robin at WINDOW granularity to the devices (strided 2-D copies,
 *                      one PCIe link per device, no interleaving pass on the host) and returns the outputs in stream
 *                      order - one work() call of one flowgraph block uses every GPU and every PCIe link;
 *   set_table() / set_geometry() / set_peak_mode()  apply to every device;
 *   process_device()   goes to the device the input pointer lives on;  process_device_sharded() (below) runs all of them;
 *   planar, reducer, stage-timing entries are served by devices[0].
 * lib/baz_music_doa.cc picks the devices from the environment variable BAZ_MUSIC_DOA_DEVICES ("0,1,2,3" or "all").
 */
int music_b200_create_multi(music_b200 **out, uint32_t m, uint32_t n, uint32_t nsamples, uint32_t resolution,
                            const float *table_c64, const int *devices, int ndev);
int music_b200_device_count(const music_b200 *h);

/*
 * Device-resident, sharded form of work() on a multi-device handle: device g (= devices[g] of create_multi) holds in
 * d_in_c64[g] the windows w = i * G + g of a stream of nwindows_total windows (i = local index, compacted) and receives
 * their angles / levels in d_angles[g] / d_levels[g] ([ceil((nwindows_total - g) / G)][n], levels may be NULL).
 * d_bins_all (may be NULL): G device arrays, d_bins_all[p] on device p, each int32 [nwindows_total][n] in STREAM order.
 * The all-gather of the peak bins is fused into the scan epilogue: the kernel of shard g stores bin w into
 * d_bins_all[p][w] for every p over NVLink (peer-mapped memory; no NCCL, no separate collective kernel), so when the
 * G streams have finished every device holds all peaks.  streams (may be NULL): one cudaStream_t per device.
 */
int music_b200_process_device_sharded(music_b200 *h, const float *const *d_in_c64, uint32_t nwindows_total, float *const *d_angles,
                                      float *const *d_levels, int32_t *const *d_bins_all, void *const *streams);

/*
 * The same fused all-gather when every GPU belongs to its own PROCESS (one rank per GPU, e.g. under torchrun): each rank
 *   1. gather_create(h, total_windows, handles)  allocates its stream-ordered gather buffer (int32 [total_windows][n]) and
 *      epoch flags and returns their two 64-byte CUDA IPC handles;
 *   2. exchanges the handles with the other ranks by any means (bench.py: one all_gather of 128 bytes at start-up);
 *   3. gather_attach(h, nranks, rank, all_handles)  maps the peers' buffers.
 * From then on every process_device() call of this handle treats its nwindows windows as the shard w = i * nranks + rank
 * and its scan epilogue stores the peak bins into every rank's gather buffer; the last CTA raises this rank's epoch flag
 * at every peer.  gather_wait(h, stream) enqueues a one-warp kernel that returns when all ranks' flags have reached this
 * rank's call count, i.e. when gather_buffer(h) holds every rank's bins of the latest call (all ranks must make the same
 * sequence of calls, as with any collective).  ncclAllGather stays the reference implementation the tests compare with.
 */
int music_b200_gather_create(music_b200 *h, uint32_t total_windows, unsigned char *ipc_handles /* [2][64] */);
int music_b200_gather_attach(music_b200 *h, int nranks, int rank, const unsigned char *all_handles /* [nranks][2][64] */);
int music_b200_gather_wait(music_b200 *h, void *stream);
const int32_t *music_b200_gather_buffer(const music_b200 *h);
int music_b200_gather_read(music_b200 *h, int32_t *host_out, uint32_t count); /* synchronous copy of the first `count` entries */

/*
 * Replaces baz_music_doa::set_array_response() (/root/reference/lib/baz_music_doa.cc:60-70).
 * Thread-safe against a concurrent process_*() call, like the reference's d_mutex (:67, :101):
 * a call in flight finishes with the old table, later calls see the new one.
 */
int music_b200_set_table(music_b200 *h, const float *table_c64);

/*
 * Peak rule.  MUSIC_B200_PEAKS_TOP_BINS (default) is the reference's: the n largest bins, ties to the
 * lower bin (/root/reference/lib/baz_music_doa.cc:129-141) - for n >= 2 usually neighbouring bins of one
 * peak.  MUSIC_B200_PEAKS_LOCAL_MAXIMA is an opt-in extension with no reference counterpart: the n
 * largest circular local maxima (P[k] > P[k-1] and P[k] >= P[k+1], P[k] > 0), each more than
 * exclusion_bins away from every stronger peak taken; outputs in descending strength, unfilled slots
 * (0, 0, bin -1).  Applies to later process_*() calls.
 */
#define MUSIC_B200_PEAKS_TOP_BINS 0
#define MUSIC_B200_PEAKS_LOCAL_MAXIMA 1
int music_b200_set_peak_mode(music_b200 *h, int mode, uint32_t exclusion_bins);

/*
 * Downstream reducers (extension, no reference counterpart): what a GUI-rate consumer such as the DOA
 * compass (/root/reference/python/doa_compass_control.py:102-108) or a plot sink
 * (/root/reference/python/plot_sink.py:38) needs from millions of windows per second.
 *   reduce_angles: for every angle slot i < n the CIRCULAR mean over nwindows of angles[w][i] (degrees in
 *     [0, 360)), the mean resultant length (1: all windows agree, 0: spread uniformly) and the total
 *     weight.  A window counts iff levels == NULL or levels[w][i] > 0 (unfilled slots carry level 0);
 *     weighted != 0 weights each window by its level.  resultant / weight_sum may be NULL.
 *   reduce_spectrum: mean over nwindows of spectrum[w][k], k < resolution.
 * _device: buffers in device memory, enqueued on `stream`; _host: host buffers (upload, reduce, download).
 */
int music_b200_reduce_angles_device(music_b200 *h, const float *d_angles, const float *d_levels, uint32_t nwindows, int weighted,
                                    float *d_mean_deg, float *d_resultant, float *d_weight_sum, void *stream);
int music_b200_reduce_spectrum_device(music_b200 *h, const float *d_spectrum, uint32_t nwindows, float *d_mean, void *stream);
int music_b200_reduce_angles_host(music_b200 *h, const float *angles, const float *levels, uint32_t nwindows, int weighted,
                                  float *mean_deg, float *resultant, float *weight_sum);
int music_b200_reduce_spectrum_host(music_b200 *h, const float *spectrum, uint32_t nwindows, float *mean);

/*
 * Planar input: one c64 stream per antenna instead of interleaved items.  Window w is
 *     x_w(r, c) = streams[r][w * hop + c],   r < m, c < N = nsamples / m,
 * i.e. what the flowgraph in front of the reference block builds on the CPU by interleaving the
 * antenna streams (/root/reference/lib/baz_interleaver.cc:152-229), cutting vectors and, for
 * sliding windows, re-copying the overlap (/root/reference/lib/baz_overlap.cc:107-129) before
 * baz_music_doa::work() reshapes it back (/root/reference/lib/baz_music_doa.cc:82-84).
 * hop == N: back-to-back windows; hop < N: windows overlap by N - hop snapshots; hop > N: gaps.
 * Each stream must hold (nwindows - 1) * hop + N samples.  Outputs as in process_host/_device.
 * Results equal those of process_*() on the interleaved windows up to fp64 rounding of R (the planar kernels sum the
 * snapshots in a different order for m > 4): P(theta) agrees to ~1e-12 relative, peak bins are equal except where two
 * bins' strengths tie to within that rounding (in practice the 90/270 degree pair of an x-axis ULA).
 *   _host  : streams[r] are host pointers (copied to the device in chunks, planar, no interleave)
 *   _device: d_streams is a HOST array of m device pointers (8-byte aligned), work is enqueued on `stream`
 */
int music_b200_process_planar_host(music_b200 *h, const float *const *streams, uint32_t hop, uint32_t nwindows,
                                   float *angles, float *levels, float *spectrum, int32_t *bins);
int music_b200_process_planar_device(music_b200 *h, const float *const *d_streams, uint32_t hop, uint32_t nwindows,
                                     float *d_angles, float *d_levels, float *d_spectrum, int32_t *d_bins, void *stream);

/*
 * Retune without marshalling a table: builds the array response ON THE DEVICE from the element
 * positions and the wavelength, i.e. replaces calculate_antenna_array_response() + the SWIG
 * complex128 -> complex64 conversion + set_array_response()
 * (/root/reference/python/music_doa_helper.py:29-46, :100-103; swig/baz_swig.i:564) by one call.
 * positions_xy: [m][2] doubles in metres, already multiplied by array_spacing
 * (music_doa_helper.py:56); wavelength = 299792458 / frequency (:55, :101).
 * The resulting table is bit-identical to the one the Python helper computes on this host:
 * entries whose float32 rounding could depend on the last bits of sin/cos are re-evaluated
 * with the host libm; *guarded (may be NULL) receives how many (typically ~3e-5 of 2*m*K).
 * Same locking and double buffering as music_b200_set_table().
 */
int music_b200_set_geometry(music_b200 *h, const double *positions_xy, double wavelength, uint32_t *guarded);

/*
 * Test hook (no device needed): the host-libm evaluation of one table entry that
 * music_b200_set_geometry() uses for the guarded entries; re_im = {Re, Im} as float32.
 */
void music_b200_steer_entry_host(const double *positions_xy, double wavelength, uint32_t resolution, uint32_t step,
                                 uint32_t antenna, float *re_im);

/* Copies the table currently in use (float[resolution][m][2]) back to host memory. */
int music_b200_get_table(music_b200 *h, float *table_c64);

/*
 * Replaces the body of baz_music_doa::work() (/root/reference/lib/baz_music_doa.cc:72-161)
 * for `nwindows` consecutive input items held in HOST memory (what the GNU Radio scheduler
 * hands to work(): input_items[0], output_items[0..2]).  Host<->device copies are done
 * inside (chunked, double-buffered).  levels / spectrum / bins may be NULL, mirroring
 * output_items.size() (:97-99, :148-149); NULL skips that work.
 * Pageable caller memory (GNU Radio's circular buffers, numpy arrays) of 1 MiB or more is pinned with cudaHostRegister
 * the first time it is seen and remembered in the handle (up to 16 ranges; released by destroy()), so that later calls
 * on the same buffers copy at the pinned rate; MUSIC_B200_HOSTREG=0 in the environment turns that off.
 */
int music_b200_process_host(music_b200 *h, const float *in_c64, uint32_t nwindows,
                            float *angles, float *levels, float *spectrum, int32_t *bins);

/*
 * Same computation with all buffers already resident in device memory (the streaming /
 * benchmark path, and the entry a device-resident upstream block would call).  Enqueued on
 * `stream` (a cudaStream_t, NULL = legacy default stream); does not synchronise.
 * d_in must be 16-byte aligned.
 * m = 4, n = 1 (and m = 8, n = 1 without the spectrum) run as ONE persistent kernel per call; consecutive calls of the
 * m = 4 kernel on one stream overlap their launch tails (programmatic dependent launch) while their results still
 * land in stream order.  d_spectrum != NULL (reference :120-121) keeps m = 4 / n = 1 on that kernel (its fp64 workers
 * write the spectrum); other shapes then take the three-kernel path.
 */
int music_b200_process_device(music_b200 *h, const float *d_in_c64, uint32_t nwindows,
                              float *d_angles, float *d_levels, float *d_spectrum,
                              int32_t *d_bins, void *stream);

/*
 * process_device plus optional fp64 internals for stage-by-stage parity tests against the
 * oracle (any of them may be NULL):
 *   d_P64     [nwindows][resolution]  strength = 1/||G^H a||^2 before the float cast (:114-119)
 *   d_R       [nwindows][m][m][2]     covariance R = x x^H / N (:85)
 *   d_eigvals [nwindows][m]           ascending eigenvalues (:88-90)
 */
int music_b200_process_device_ex(music_b200 *h, const float *d_in_c64, uint32_t nwindows,
                                 float *d_angles, float *d_levels, float *d_spectrum,
                                 int32_t *d_bins, double *d_P64, double *d_R,
                                 double *d_eigvals, void *stream);

/* Number of CUDA kernels this handle has launched so far (bench.py's gpu_launches). */
uint64_t music_b200_launch_count(const music_b200 *h);

/*
 * Optional per-stage device timing for the benchmark's roofline leg.  While enabled, every
 * internal chunk records CUDA events on the launch stream around K1 (covariance), K2
 * (eigendecomposition), K3 (scan) and the top-n kernel.  get_stage_times() synchronises on
 * them, returns the accumulated milliseconds since the last call in ms4[0..3] and the number
 * of chunks in *chunks (may be NULL), then resets the accumulation.
 */
int music_b200_set_stage_timing(music_b200 *h, int enable);
int music_b200_get_stage_times(music_b200 *h, double *ms4, uint64_t *chunks);

/* Debug: copy the fused kernel's per-CTA clock64 trace (16 int64 per CTA; only when the handle
 * was created with MUSIC_B200_TRACE=1 in the environment). */
int music_b200_debug_fused_trace(music_b200 *h, long long *host_out, int max_ctas);

/* Debug: cumulative window counts of the fused M = 8 kernel - out2[0] solved by the principal-eigenvector (squaring) solver,
 * out2[1] by the Jacobi fallback. */
int music_b200_debug_fused8_stats(music_b200 *h, uint64_t *out2);

/* Last error text for this handle; h == NULL returns the last create() failure. */
const char *music_b200_last_error(const music_b200 *h);

void music_b200_destroy(music_b200 *h);

#ifdef __cplusplus
}
#endif
#endif /* MUSIC_B200_H */
