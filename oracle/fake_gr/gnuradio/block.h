/* TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the slice of the
 * GNU Radio 3.7 runtime API that the gr-dvbt block sources touch, so that the
 * reference's own lib .cc files can be compiled *verbatim* (from /root/reference,
 * never copied) into oracle/_ref/libdvbt_ref.so and driven one general_work()
 * call at a time by oracle/ref_harness.cc.  No scheduler, no buffers: the
 * harness owns the item counters and the tag lists.
 *
 * libdvbt_b200.so (the product) never sees this header.  Its only other user is the compile check /
 * test build of the gr::block shims (gr_dvbt_b200/shim/Makefile -> libdvbt_b200_shim_test.so, loaded by
 * tests/test_shim_gpu.py alone): the shims are written against the real gnuradio/block.h, and this image has none.
 */
#ifndef DVBT_ORACLE_FAKE_GR_BLOCK_H
#define DVBT_ORACLE_FAKE_GR_BLOCK_H

#include <stdint.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace pmt {
/* a pmt is either a symbol (text) or a long in the code paths gr-dvbt uses */
struct pmt_t {
  std::string text;
  long number;
  pmt_t() : number(0) {}
};
inline bool operator==(const pmt_t &a, const pmt_t &b) { return a.text == b.text && a.number == b.number; }
inline pmt_t string_to_symbol(const std::string &s) { pmt_t p; p.text = s; return p; }
inline pmt_t from_long(long v) { pmt_t p; p.number = v; return p; }
inline long to_long(const pmt_t &p) { return p.number; }
inline std::string symbol_to_string(const pmt_t &p) { return p.text; }
}  // namespace pmt

namespace gr {

struct tag_t {
  uint64_t offset;
  pmt::pmt_t key;
  pmt::pmt_t value;
  tag_t() : offset(0) {}
};

class io_signature {
 public:
  typedef boost::shared_ptr<io_signature> sptr;
  static sptr make(int min_streams, int max_streams, int item_size) {
    sptr s(new io_signature);
    s->d_min = min_streams; s->d_max = max_streams; s->d_size = item_size;
    return s;
  }
  int min_streams() const { return d_min; }
  int max_streams() const { return d_max; }
  int sizeof_stream_item(int) const { return d_size; }
 private:
  int d_min, d_max, d_size;
};

class block {
 public:
  block() : h_nread(0), h_nwritten(0), h_consumed(0), h_relative_rate(1.0), h_output_multiple(1) {}
  block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
      : h_name(name), h_in(in), h_out(out), h_nread(0), h_nwritten(0), h_consumed(0),
        h_relative_rate(1.0), h_output_multiple(1) {}
  virtual ~block() {}

  io_signature::sptr input_signature() const { return h_in; }
  io_signature::sptr output_signature() const { return h_out; }
  std::string name() const { return h_name; }

  void set_relative_rate(double r) { h_relative_rate = r; }
  void set_output_multiple(int m) { h_output_multiple = m; }
  void set_alignment(int) {}
  bool is_unaligned() { return true; }
  void set_history(unsigned) {}
  void set_tag_propagation_policy(int) {}
  void set_min_noutput_items(int) {}
  void set_min_output_buffer(long) {}
  void set_min_output_buffer(int, long) {}

  void consume_each(int n) { h_consumed = n; }
  uint64_t nitems_read(unsigned) { return h_nread; }
  uint64_t nitems_written(unsigned) { return h_nwritten; }

  void add_item_tag(unsigned, uint64_t offset, const pmt::pmt_t &key, const pmt::pmt_t &value) {
    tag_t t; t.offset = offset; t.key = key; t.value = value;
    h_out_tags.push_back(t);
  }
  void get_tags_in_range(std::vector<tag_t> &v, unsigned, uint64_t start, uint64_t end, const pmt::pmt_t &key) {
    v.clear();
    for (size_t i = 0; i < h_in_tags.size(); i++)
      if (h_in_tags[i].offset >= start && h_in_tags[i].offset < end && h_in_tags[i].key.text == key.text)
        v.push_back(h_in_tags[i]);
  }

  virtual void forecast(int, gr_vector_int &) {}
  virtual int general_work(int, gr_vector_int &, gr_vector_const_void_star &, gr_vector_void_star &) { return 0; }

  /* harness-owned state (public on purpose) */
  std::string h_name;
  io_signature::sptr h_in, h_out;
  uint64_t h_nread, h_nwritten;
  int h_consumed;
  double h_relative_rate;
  int h_output_multiple;
  std::vector<tag_t> h_in_tags, h_out_tags;
};

class sync_interpolator : public block {
 public:
  sync_interpolator() : h_interp(1) {}
  sync_interpolator(const std::string &name, io_signature::sptr in, io_signature::sptr out, unsigned interp)
      : block(name, in, out), h_interp(interp) {}
  virtual int work(int, gr_vector_const_void_star &, gr_vector_void_star &) { return 0; }
  unsigned h_interp;
};

}  // namespace gr

namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}

#endif
