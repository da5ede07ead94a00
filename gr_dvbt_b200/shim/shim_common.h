// gr::block shims over the C ABI of libdvbt_b200 (include/dvbt_b200.h).
//
// Each shim derives from the reference's own public block class (include/dvbt/<block>.h) and
// defines that class's make(): built into libgnuradio-dvbt in place of lib/<block>_impl.cc, the
// GRC files, the SWIG module and every flowgraph in apps/ keep working unchanged.
// The scheduler still owns the stream buffers; a shim borrows them for the duration of one call
// (the library stages them through its own device buffers) and asks the scheduler for large work
// items, because one kernel launch per 768-bit block would waste the GPU.
#ifndef DVBT_B200_SHIM_COMMON_H
#define DVBT_B200_SHIM_COMMON_H

#include <gnuradio/block.h>
#include <gnuradio/io_signature.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "dvbt_b200.h"

namespace gr {
namespace dvbt {
namespace b200 {

inline void check(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + dvbt_b200_last_error());
}

inline const char *tag_name(int key) {
  return key == DVBT_TAG_SYNC_START ? "sync_start" : key == DVBT_TAG_SUPERFRAME_START ? "superframe_start" : "symbol_index";
}

// tags of one key inside [nread, nread + window) as offsets relative to the window
inline void collect_tags(gr::block *b, const char *key, int dvbt_key, uint64_t nread, uint64_t window, std::vector<dvbt_b200_tag> &out) {
  std::vector<gr::tag_t> tags;
  b->get_tags_in_range(tags, 0, nread, nread + window, pmt::string_to_symbol(key));
  for (size_t i = 0; i < tags.size(); i++) {
    dvbt_b200_tag t;
    t.offset = tags[i].offset - nread;
    t.key = dvbt_key;
    t.value = pmt::to_long(tags[i].value);
    out.push_back(t);
  }
}

inline void emit_tags(gr::block *b, uint64_t nwritten, const dvbt_b200_tag *tags, size_t n) {
  for (size_t i = 0; i < n; i++)
    b->add_item_tag(0, nwritten + tags[i].offset, pmt::string_to_symbol(tag_name(tags[i].key)), pmt::from_long((long)tags[i].value));
}

}  // namespace b200
}  // namespace dvbt
}  // namespace gr
#endif
