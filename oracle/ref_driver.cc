// oracle/ref_driver.cc - TEST INFRASTRUCTURE ONLY.
// C entry point around the reference's own block (compiled from /root/reference/lib/baz_music_doa.cc against the
// stand-in headers, see oracle/ref_shim/armadillo and lib/gr_shim/): builds the block with the reference's factory and
// calls its work() once per window, the way the GNU Radio scheduler does (work() returns 1: one item per call,
// /root/reference/lib/baz_music_doa.cc:160).
#include <baz_music_doa.h>

#include <complex>
#include <vector>

extern "C" int ref_music_work_batch(const float *in_c64, int W, int m, int n, int nsamples, const float *table_c64, int K, float *angles,
                                    float *levels, float *spectrum /* may be NULL */)
{
    if (!in_c64 || !table_c64 || !angles || !levels || W < 0 || m <= 0 || n <= 0 || n >= m || K <= 0 || nsamples % m) return -1;
    array_response_t resp((size_t)K, antenna_response_t((size_t)m));
    for (int k = 0; k < K; ++k)
        for (int a = 0; a < m; ++a) resp[k][a] = gr_complex(table_c64[2 * ((size_t)k * m + a)], table_c64[2 * ((size_t)k * m + a) + 1]);
    baz_music_doa_sptr blk = baz_make_music_doa((unsigned)m, (unsigned)n, (unsigned)nsamples, resp, (unsigned)K);
    for (int w = 0; w < W; ++w) {
        gr_vector_const_void_star ins(1, in_c64 + (size_t)w * nsamples * 2);
        gr_vector_void_star outs;
        outs.push_back(angles + (size_t)w * n);
        outs.push_back(levels + (size_t)w * n);  // the reference dereferences output 1 unconditionally (:147-154)
        if (spectrum) outs.push_back(spectrum + (size_t)w * K);
        if (blk->work(1, ins, outs) != 1) return -2;
    }
    return 0;
}
