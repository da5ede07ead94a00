/* TEST INFRASTRUCTURE ONLY — CPU restatement of gr::dvbt::dvbt_demap.
 *   lib/dvbt_demap_impl.cc:117-165 make_constellation_points (Gray mapping, axis-bit
 *       de-interleave b0b2b4|b1b3b5, points scaled by gain*norm)
 *   lib/dvbt_demap_impl.cc:167-203 find_constellation_value (first strictly smallest squared
 *       distance; distances as VOLK's generic 32fc_x2_square_dist_32f computes them: complex
 *       subtract, re*re + im*im in float)
 *   lib/dvbt_config.cc:229-249 normalisation factors
 */
#include "dvbt_oracle.h"
#include <math.h>

static int gray(int v) { return (v >> 1) ^ v; } /* dvbt_demap_impl.cc:205-209 */

/* constellation: 0 QPSK, 1 QAM16, 2 QAM64; alpha 1,2,4.  points: 2*size floats (re,im). */
int dvbt_oracle_constellation(int constellation, int alpha, float gain, float *points) {
  int size = constellation == 0 ? 4 : constellation == 1 ? 16 : 64;
  int m = constellation == 0 ? 2 : constellation == 1 ? 4 : 6;
  int step = 2; /* dvbt_config.cc:127-147 */
  float norm;   /* dvbt_config.cc:229-249: computed in double, stored in a float member */
  if (m == 2) norm = (float)(1.0 / sqrt(2));
  else if (m == 4) norm = (float)(alpha == 1 ? 1.0 / sqrt(10) : alpha == 2 ? 1.0 / sqrt(20) : 1.0 / sqrt(52));
  else norm = (float)(alpha == 1 ? 1.0 / sqrt(42) : alpha == 2 ? 1.0 / sqrt(60) : 1.0 / sqrt(108));
  float g = gain * norm; /* dvbt_demap_impl.cc:73 */
  int bpa = m / 2;       /* bits_per_axis = log2(size)/2 */
  int spa = (1 << bpa) / 2 - 1; /* steps_per_axis = sqrt(size)/2 - 1 */
  for (int i = 0; i < size; i++) {
    int q = (i >> (2 * (bpa - 1))) & 3;
    int sign0 = (q >> 1) ? -1 : 1, sign1 = (q & 1) ? -1 : 1;
    int x = (i >> (bpa - 1)) & ((1 << (bpa - 1)) - 1);
    int y = i & ((1 << (bpa - 1)) - 1);
    int xval = alpha + (spa - x) * step;
    int yval = alpha + (spa - y) * step;
    int val = (gray(x) << (bpa - 1)) + gray(y);
    x = 0; y = 0;
    for (int j = 0; j < bpa - 1; j++) {
      x += ((val >> (1 + 2 * j)) & 1) << j;
      y += ((val >> (2 * j)) & 1) << j;
    }
    val = (q << (2 * (bpa - 1))) + (x << (bpa - 1)) + y;
    points[2 * val] = g * (float)(sign0 * xval);
    points[2 * val + 1] = g * (float)(sign1 * yval);
  }
  return size;
}

/* in: n complex cells (re,im interleaved floats) -> out: n bytes */
void dvbt_oracle_demap(const float *in, long n, int constellation, int alpha, float gain, uint8_t *out) {
  float pts[128];
  int size = dvbt_oracle_constellation(constellation, alpha, gain, pts);
  for (long c = 0; c < n; c++) {
    float re = in[2 * c], im = in[2 * c + 1];
    volatile float dr = re - pts[0], di = im - pts[1];
    volatile float a = dr * dr, b = di * di;
    float min_dist = a + b; /* std::norm(val - points[0]), :169 */
    int min_index = 0;
    for (int i = 0; i < size; i++) {
      dr = re - pts[2 * i]; di = im - pts[2 * i + 1];
      a = dr * dr; b = di * di;
      float d = a + b;
      if (d < min_dist) { min_dist = d; min_index = i; }
    }
    out[c] = (uint8_t)min_index;
  }
}
