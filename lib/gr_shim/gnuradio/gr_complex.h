// gr_shim: stand-in for <gnuradio/gr_complex.h> and <gnuradio/types.h> (compile-only, see README.md)
#ifndef GR_SHIM_GR_COMPLEX_H
#define GR_SHIM_GR_COMPLEX_H
#include <complex>
#include <vector>
typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;
typedef std::vector<int> gr_vector_int;
#endif
