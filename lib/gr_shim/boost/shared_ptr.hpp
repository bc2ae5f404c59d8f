// gr_shim: boost::shared_ptr stand-in (GNU Radio 3.7 block factories return boost::shared_ptr)
#ifndef GR_SHIM_BOOST_SHARED_PTR_HPP
#define GR_SHIM_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
using std::shared_ptr;
}
#endif
