/* TEST INFRASTRUCTURE: generic (scalar) statements of the seven VOLK kernels the
 * reference calls (ofdm_sym_acquisition_impl.cc:168-237,531; dvbt_demap_impl.cc:176),
 * following VOLK's documented *_generic semantics.  VOLK is not vendored under
 * /root/reference, so these are "parity unpinned" at float-rounding level. */
#pragma once
#include <complex>
#include <cstdlib>
typedef std::complex<float> lv_32fc_t;
static inline size_t volk_get_alignment() { return 16; }
#define DVBT_VOLK_BOTH(name) name##_a, name##_u
static inline void volk_32fc_magnitude_squared_32f_u(float *o, const lv_32fc_t *a, unsigned n) {
  for (unsigned i = 0; i < n; i++) o[i] = a[i].real() * a[i].real() + a[i].imag() * a[i].imag();
}
static inline void volk_32fc_x2_multiply_conjugate_32fc_u(lv_32fc_t *o, const lv_32fc_t *a, const lv_32fc_t *b, unsigned n) {
  for (unsigned i = 0; i < n; i++) o[i] = a[i] * std::conj(b[i]);
}
static inline void volk_32fc_magnitude_32f_u(float *o, const lv_32fc_t *a, unsigned n) {
  for (unsigned i = 0; i < n; i++) o[i] = sqrtf(a[i].real() * a[i].real() + a[i].imag() * a[i].imag());
}
static inline void volk_32f_s32f_multiply_32f_u(float *o, const float *a, float s, unsigned n) {
  for (unsigned i = 0; i < n; i++) o[i] = a[i] * s;
}
static inline void volk_32f_x2_subtract_32f_u(float *o, const float *a, const float *b, unsigned n) {
  for (unsigned i = 0; i < n; i++) o[i] = a[i] - b[i];
}
static inline void volk_32fc_x2_multiply_32fc_u(lv_32fc_t *o, const lv_32fc_t *a, const lv_32fc_t *b, unsigned n) {
  for (unsigned i = 0; i < n; i++) o[i] = a[i] * b[i];
}
static inline void volk_32fc_x2_square_dist_32f_u(float *o, const lv_32fc_t *s, const lv_32fc_t *p, unsigned n) {
  for (unsigned i = 0; i < n; i++) {
    lv_32fc_t d = s[0] - p[i];
    o[i] = d.real() * d.real() + d.imag() * d.imag();
  }
}
#define volk_32fc_magnitude_squared_32f_a volk_32fc_magnitude_squared_32f_u
#define volk_32fc_x2_multiply_conjugate_32fc_a volk_32fc_x2_multiply_conjugate_32fc_u
#define volk_32fc_magnitude_32f_a volk_32fc_magnitude_32f_u
#define volk_32f_s32f_multiply_32f_a volk_32f_s32f_multiply_32f_u
#define volk_32f_x2_subtract_32f_a volk_32f_x2_subtract_32f_u
#define volk_32fc_x2_multiply_32fc_a volk_32fc_x2_multiply_32fc_u
#define volk_32fc_x2_square_dist_32f_a volk_32fc_x2_square_dist_32f_u
