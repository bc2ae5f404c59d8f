/* -*- c++ -*- */
/*
 * SWIG fragment for the B200-native MUSIC DOA block.  It replaces the slice of
 * /root/reference/swig/baz_swig.i that binds the reference block (lines 77-79 and 560-574),
 * with the gate renamed from ARMADILLO_FOUND to MUSIC_B200_FOUND.  The Python-visible name,
 * factory signature and set_array_response() are unchanged, so
 *     baz.music_doa(m, n, nsamples, array_response, resolution)
 * keeps working from python/music_doa_helper.py and from GRC-generated flowgraphs.
 *
 * SWIG is not installed in this image, so this file is not compiled here; the same C ABI is
 * exercised from Python through ctypes (gr-baz_b200/_capi.py) and from C++ through
 * lib/test_block.cc.
 */

/* --- in the %{ ... %} header block of baz_swig.i (reference :77-79) --- */
#ifdef MUSIC_B200_FOUND
#include "baz_music_doa.h"
#endif // MUSIC_B200_FOUND

/* --- in the declarations section (reference :560-574) --- */
#ifdef MUSIC_B200_FOUND

GR_SWIG_BLOCK_MAGIC(baz,music_doa)

baz_music_doa_sptr baz_make_music_doa(unsigned int m, unsigned int n, unsigned int nsamples, const std::vector<std::vector<gr_complex> >& array_response, unsigned int resolution);

class baz_music_doa : public gr::sync_block
{
private:
	baz_music_doa(unsigned int m, unsigned int n, unsigned int nsamples, const array_response_t& array_response, unsigned int resolution);
public:
	void set_array_response(const std::vector<std::vector<gr_complex> >& array_response);
	void set_array_geometry(const std::vector<std::vector<double> >& positions_xy, double wavelength);
	void set_peak_mode(int mode, unsigned int exclusion_bins);
};

#endif // MUSIC_B200_FOUND
