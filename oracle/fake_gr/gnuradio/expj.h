/* TEST INFRASTRUCTURE: gr_expj as GNU Radio 3.7 documents it (gnuradio-runtime
 * include/gnuradio/expj.h; not vendored under /root/reference): (cos, sin) of a
 * float phase via sincosf.  Parity at the last ulp is unpinned. */
#pragma once
#include <gnuradio/block.h>
static inline gr_complex gr_expj(float phase) {
  float s, c;
  sincosf(phase, &s, &c);
  return gr_complex(c, s);
}
