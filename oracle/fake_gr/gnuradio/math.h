/* TEST INFRASTRUCTURE: gr::fast_atan2f is a table approximation in GNU Radio
 * (gnuradio-runtime lib/math/fast_atan2f.cc; not vendored).  The oracle uses the
 * exact atan2f; the difference (< 1e-3 rad) only moves the fractional CFO
 * estimate, which the equaliser absorbs.  "parity unpinned" for this call. */
#pragma once
#include <gnuradio/block.h>
namespace gr {
static inline float fast_atan2f(float y, float x) { return atan2f(y, x); }
static inline float fast_atan2f(gr_complex z) { return atan2f(z.imag(), z.real()); }
}
