// dvbt_demap — hard-decision constellation demapper.
//
// Replaces gr::dvbt::dvbt_demap (lib/dvbt_demap_impl.cc:217-240 general_work,
// :167-203 find_constellation_value, :117-165 make_constellation_points).  The decision is
// the reference's: the first index with the strictly smallest squared distance, distances
// computed as VOLK's generic 32fc_x2_square_dist_32f does (complex subtract, re*re + im*im,
// every float operation rounded on its own: __fsub_rn/__fmul_rn/__fadd_rn forbid FMA
// contraction), scanned from index 0 upwards.  One thread per cell; the 4/16/64 points sit in
// the kernel parameter block (constant bank, read uniformly by the warp).  HBM bound:
// 8 B in + 1 B out per cell.
#include "demod.cuh"

#include <math.h>
#include <new>

namespace dvbt {

static int gray(int v) { return (v >> 1) ^ v; }

// dvbt_demap_impl.cc:117-165 with the normalisation of dvbt_config.cc:229-249
int make_demap_table(int constellation, int hierarchy, float gain, DemapTable *t) {
  if (constellation < DVBT_QPSK || constellation > DVBT_QAM64) return DVBT_B200_EINVAL;
  int alpha = hierarchy == DVBT_ALPHA2 ? 2 : hierarchy == DVBT_ALPHA4 ? 4 : 1;
  int m = 2 * (constellation + 1);
  int size = 1 << m;
  const int step = 2;
  float norm;
  if (m == 2) norm = (float)(1.0 / sqrt(2));
  else if (m == 4) norm = (float)(alpha == 1 ? 1.0 / sqrt(10) : alpha == 2 ? 1.0 / sqrt(20) : 1.0 / sqrt(52));
  else norm = (float)(alpha == 1 ? 1.0 / sqrt(42) : alpha == 2 ? 1.0 / sqrt(60) : 1.0 / sqrt(108));
  float g = gain * norm;
  int bpa = m / 2, spa = (1 << bpa) / 2 - 1;
  for (int i = 0; i < size; i++) {
    int q = (i >> (2 * (bpa - 1))) & 3;
    int sign0 = (q >> 1) ? -1 : 1, sign1 = (q & 1) ? -1 : 1;
    int x = (i >> (bpa - 1)) & ((1 << (bpa - 1)) - 1), y = i & ((1 << (bpa - 1)) - 1);
    int xval = alpha + (spa - x) * step, yval = alpha + (spa - y) * step;
    int val = (gray(x) << (bpa - 1)) + gray(y);
    x = 0; y = 0;
    for (int j = 0; j < bpa - 1; j++) {
      x += ((val >> (1 + 2 * j)) & 1) << j;
      y += ((val >> (2 * j)) & 1) << j;
    }
    val = (q << (2 * (bpa - 1))) + (x << (bpa - 1)) + y;
    t->pts[val] = make_float2(g * (float)(sign0 * xval), g * (float)(sign1 * yval));
  }
  t->size = size;
  // per-axis view (see demap_cell_near): levels g * n in ascending order and the index bits each one stands for
  t->g = g;
  t->alpha = alpha;
  t->inv_step = 1.0f / (2.0f * g);
  t->bx = t->by = 0;
  t->near_ok = (g > 0.f && alpha == 1) ? 1 : 0;   // the shortcut's level arithmetic assumes the uniform grid
  {
    const int H = m / 2, L = 1 << H, half = L / 2;
    for (int a = 0; a < L; a++) {
      int ix = 0, iy = 0;
      for (int j = 0; j < H; j++) {
        int bit = (a >> (H - 1 - j)) & 1;
        ix |= bit << (m - 1 - 2 * j);
        iy |= bit << (m - 2 - 2 * j);
      }
      int kx = -1, ky = -1;
      for (int k = 0; k < L; k++) {
        int n = k >= half ? alpha + 2 * (k - half) : -(alpha + 2 * (half - 1 - k));
        float lv = g * (float)n;
        if (lv == t->pts[ix].x) kx = k;
        if (lv == t->pts[iy].y) ky = k;
      }
      if (kx < 0 || ky < 0) { t->near_ok = 0; break; }
      t->bx |= (unsigned long long)ix << (8 * kx);
      t->by |= (unsigned long long)iy << (8 * ky);
    }
    // soft-decision view: the levels in ascending order and, per bit, the levels where it is 1
    for (int k = 0; k < 8; k++) { t->soft_lv[0][k] = t->soft_lv[1][k] = 0.f; t->soft_ones[k] = 0; }
    t->soft_scale = 4.0f;
    for (int k = 0; k < L; k++) {
      const int ix = (int)((t->bx >> (8 * k)) & 0xff), iy = (int)((t->by >> (8 * k)) & 0xff);
      t->soft_lv[0][k] = t->pts[ix].x;
      t->soft_lv[1][k] = t->pts[iy].y;
      for (int e = 0; e < m; e++) {
        const int idx = (e & 1) ? iy : ix;
        if ((idx >> (m - 1 - e)) & 1) t->soft_ones[e] |= (unsigned char)(1u << k);
      }
    }
    // every point must be the product of its two axis levels (separable constellation)
    for (int i = 0; i < size && t->near_ok; i++) {
      int xm = 0, ym = 0;
      for (int j = 0; j < H; j++) { xm |= 1 << (m - 1 - 2 * j); ym |= 1 << (m - 2 - 2 * j); }
      if (t->pts[i].x != t->pts[i & xm].x || t->pts[i].y != t->pts[i & ym].y) t->near_ok = 0;
    }
  }
  return 0;
}

__device__ __forceinline__ uint8_t demap_cell(const DemapTable &t, float2 v) { return demap_cell_any(t, v); }

__global__ void __launch_bounds__(256) demap_kernel(const float2 *__restrict__ in, uint8_t *__restrict__ out, long long n,
                                                    const __grid_constant__ DemapTable t) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 *p = reinterpret_cast<const float4 *>(in + i);
    float4 a = p[0], b = p[1];
    uchar4 r;
    r.x = demap_cell(t, make_float2(a.x, a.y));
    r.y = demap_cell(t, make_float2(a.z, a.w));
    r.z = demap_cell(t, make_float2(b.x, b.y));
    r.w = demap_cell(t, make_float2(b.z, b.w));
    *reinterpret_cast<uchar4 *>(out + i) = r;
  } else {
    for (; i < n; i++) out[i] = demap_cell(t, in[i]);
  }
}

int demap_launch(const DemapTable &t, const float2 *d_in, uint8_t *d_out, long long ncells, cudaStream_t st) {
  if (ncells <= 0) return 0;
  long long threads = (ncells + 3) / 4;
  unsigned grid = (unsigned)((threads + 255) / 256);
  demap_kernel<<<grid, 256, 0, st>>>(d_in, d_out, ncells, t);
  count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace dvbt

struct dvbt_b200_demap {
  int device = dvbt::current_device();
  dvbt_b200_demap_params par;
  dvbt::DemapTable table;
  cudaStream_t stream = nullptr;
  dvbt::DevBuf d_in, d_out;
  dvbt::Staging stg;
};

extern "C" {

int dvbt_b200_demap_create(const dvbt_b200_demap_params *p, dvbt_b200_demap **out) {
  if (!p || !out) { dvbt::set_error("demap_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  if (p->nsize <= 0) { dvbt::set_error("demap_create: nsize must be positive"); return DVBT_B200_EINVAL; }
  dvbt::DemapTable t;
  if (dvbt::make_demap_table(p->constellation, p->hierarchy, p->gain, &t)) {
    dvbt::set_error("demap_create: bad constellation %d", p->constellation);
    return DVBT_B200_EINVAL;
  }
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_demap *h = new (std::nothrow) dvbt_b200_demap();
  if (!h) { dvbt::set_error("demap_create: out of memory"); return DVBT_B200_ENOMEM; }
  h->par = *p;
  h->table = t;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    dvbt::set_error("demap_create: cannot create stream");
    delete h;
    return DVBT_B200_ECUDA;
  }
  *out = h;
  return 0;
}

void dvbt_b200_demap_destroy(dvbt_b200_demap *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  h->d_in.release();
  h->d_out.release();
  h->stg.release();
  delete h;
}

int dvbt_b200_demap_points(const dvbt_b200_demap *h, float *re_im, int capacity_points) {
  if (!h || !re_im || capacity_points < h->table.size) { dvbt::set_error("demap_points: bad argument"); return DVBT_B200_EINVAL; }
  for (int i = 0; i < h->table.size; i++) { re_im[2 * i] = h->table.pts[i].x; re_im[2 * i + 1] = h->table.pts[i].y; }
  return h->table.size;
}

int dvbt_b200_demap_run_dev(dvbt_b200_demap *h, const void *d_in, size_t ncells, uint8_t *d_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (ncells && (!d_in || !d_out))) { dvbt::set_error("demap_run_dev: bad argument"); return DVBT_B200_EINVAL; }
  int rc = dvbt::join_default_stream(h->stream);
  if (rc) return rc;
  rc = dvbt::demap_launch(h->table, (const float2 *)d_in, d_out, (long long)ncells, h->stream);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int dvbt_b200_demap_work(dvbt_b200_demap *h, const void *in, size_t n_in_items, uint8_t *out, size_t noutput_items,
                         size_t *consumed, size_t *produced) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !consumed || !produced) { dvbt::set_error("demap_work: null argument"); return DVBT_B200_EINVAL; }
  *consumed = *produced = 0;
  if (n_in_items < noutput_items) { dvbt::set_error("demap_work: %zu input items for %zu output items", n_in_items, noutput_items); return DVBT_B200_EINVAL; }
  if (noutput_items == 0) return 0;
  if (!in || !out) { dvbt::set_error("demap_work: null buffer"); return DVBT_B200_EINVAL; }
  size_t ncells = noutput_items * (size_t)h->par.nsize;
  int rc;
  if ((rc = h->d_in.reserve(ncells * 8))) return rc;
  if ((rc = h->d_out.reserve(ncells))) return rc;
  if ((rc = h->stg.h2d(h->d_in.p, in, ncells * 8, h->stream))) return rc;
  rc = dvbt::demap_launch(h->table, h->d_in.as<float2>(), h->d_out.as<uint8_t>(), (long long)ncells, h->stream);
  if (rc) return rc;
  if ((rc = h->stg.d2h(out, h->d_out.p, ncells, h->stream))) return rc;
  *consumed = *produced = noutput_items;  // 1:1 (dvbt_demap_impl.cc:211-215, :236-239)
  return 0;
}

}  // extern "C"
