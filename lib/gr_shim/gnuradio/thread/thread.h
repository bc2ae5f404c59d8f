// gr_shim: stand-in for <gnuradio/thread/thread.h> (compile-only, see README.md)
#ifndef GR_SHIM_THREAD_H
#define GR_SHIM_THREAD_H
#include <mutex>
namespace gr {
namespace thread {
typedef std::mutex mutex;
typedef std::unique_lock<std::mutex> scoped_lock;
}  // namespace thread
}  // namespace gr
#endif
