// gr_shim: stand-in for <gnuradio/sync_block.h> (compile-only, see README.md).
// Models what a gr::sync_block subclass sees: name(), unique_id(), the two signatures, the
// output-multiple hint, and the pure virtual work().  general_work()/consume_each() belong to
// the real scheduler and are not modelled.
#ifndef GR_SHIM_SYNC_BLOCK_H
#define GR_SHIM_SYNC_BLOCK_H
#include <gnuradio/gr_complex.h>
#include <gnuradio/io_signature.h>
#include <string>
namespace gr {
class sync_block
{
public:
    virtual ~sync_block() {}
    const std::string &name() const { return d_name; }
    long unique_id() const { return d_unique_id; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }
    void set_output_multiple(int multiple) { d_output_multiple = multiple; }
    int output_multiple() const { return d_output_multiple; }
    void set_min_noutput_items(int m) { d_min_noutput_items = m; }
    int min_noutput_items() const { return d_min_noutput_items; }
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;

protected:
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_unique_id(next_id()), d_in(in), d_out(out), d_output_multiple(1), d_min_noutput_items(1)
    {
    }

private:
    static long next_id()
    {
        static long id = 0;
        return ++id;
    }
    std::string d_name;
    long d_unique_id;
    io_signature::sptr d_in, d_out;
    int d_output_multiple, d_min_noutput_items;
};
}  // namespace gr
#endif
