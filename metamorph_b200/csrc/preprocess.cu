// metamorph_b200 — on-GPU SigLIP image pre-processing, bit-exact with the reference's CPU path.
//
// Reference call sites: metamorph/train/train.py:1189-1209 (expand2square to the processor mean, then
// processor.preprocess(...)['pixel_values'][0]) with the SigLIP processor of siglip_encoder.py:113-121
// (resize 384x384 BICUBIC through Pillow, x 1/255, normalise mean 0.5 / std 0.5, channels first).
// Arithmetic restated from the un-vendored dependencies (see oracle/preprocess.py): Pillow's ImagingResample
// (src/libImaging/Resample.c) is a separable two-pass uint8 convolution with 22-bit fixed-point coefficients and a
// rounding to uint8 after each pass; the HF slow processor then maps every byte through
// float32(float64(u) / 255) -> (x - 0.5) / 0.5, i.e. a 256-entry table.
//   mm_resize_coeff_build   host: the per-axis coefficient table (double precision, Resample.c operation order)
//   resize_h_kernel         horizontal pass over the (virtually padded) square image -> uint8 [S][out][3]
//   resize_v_lut_kernel     vertical pass + table lookup + HWC->CHW, fp32 or bf16 output [3][out][out]
// Byte work, HBM/L2-bound and tiny (<= 20 M MACs per image): no tensor cores, one thread per output pixel, the
// taps of neighbouring threads overlap and are served by L1.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;   // Resample.c
constexpr int RS_THREADS = 128;

struct ResizeCoeffHeader {
  int in_size, out_size, ksize, reserved;
};
// layout of a coefficient table: header | bounds[out][2] (first tap, tap count) | kk[out][ksize]

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

int coeff_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(2.0 * filterscale) * 2 + 1;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= RS_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// img: [H][W][3] uint8; the virtual input is the S x S square with the image pasted at (top, left) on a canvas of
// `fill` (expand2square); S = side. out: [side][out_size][3].
__global__ void __launch_bounds__(RS_THREADS)
resize_h_kernel(const uint8_t* __restrict__ img, int H, int W, int side, int top, int left, int fill,
                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int out_size,
                uint8_t* __restrict__ out) {
  const int xx = blockIdx.x * RS_THREADS + threadIdx.x;
  const int y = blockIdx.y;
  if (xx >= out_size) return;
  const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
  const int* k = kk + (size_t)xx * ksize;
  int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
  const int ry = y - top;
  if (ry < 0 || ry >= H) {
    for (int x = 0; x < cnt; ++x) {
      const int w = k[x] * fill;
      s0 += w; s1 += w; s2 += w;
    }
  } else {
    const uint8_t* row = img + (size_t)ry * W * 3;
    for (int x = 0; x < cnt; ++x) {
      const int rx = xmin + x - left;
      const int w = k[x];
      if (rx < 0 || rx >= W) {
        s0 += w * fill; s1 += w * fill; s2 += w * fill;
      } else {
        s0 += w * row[rx * 3]; s1 += w * row[rx * 3 + 1]; s2 += w * row[rx * 3 + 2];
      }
    }
  }
  uint8_t* o = out + ((size_t)y * out_size + xx) * 3;
  o[0] = (uint8_t)clip8(s0);
  o[1] = (uint8_t)clip8(s1);
  o[2] = (uint8_t)clip8(s2);
}

// tmp: [side][out_w][3] uint8 -> out [3][out_size][out_w] via lut[256] (fp32), optionally rounded to bf16
__global__ void __launch_bounds__(RS_THREADS)
resize_v_lut_kernel(const uint8_t* __restrict__ tmp, int out_w, const int* __restrict__ bounds,
                    const int* __restrict__ kk, int ksize, int out_size, const float* __restrict__ lut,
                    void* __restrict__ out, int out_bf16) {
  const int xx = blockIdx.x * RS_THREADS + threadIdx.x;
  const int yy = blockIdx.y;
  if (xx >= out_w) return;
  const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
  const int* k = kk + (size_t)yy * ksize;
  int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < cnt; ++y) {
    const uint8_t* px = tmp + ((size_t)(ymin + y) * out_w + xx) * 3;
    const int w = k[y];
    s0 += w * px[0]; s1 += w * px[1]; s2 += w * px[2];
  }
  const float v[3] = {lut[clip8(s0)], lut[clip8(s1)], lut[clip8(s2)]};
  const size_t plane = (size_t)out_size * out_w, o = (size_t)yy * out_w + xx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (out_bf16) reinterpret_cast<bf16*>(out)[c * plane + o] = __float2bfloat16(v[c]);
    else reinterpret_cast<float*>(out)[c * plane + o] = v[c];
  }
}

}  // namespace

MM_API long long mm_resize_coeff_bytes(int in_size, int out_size) {
  if (in_size < 1 || out_size < 1) return 0;
  return (long long)sizeof(ResizeCoeffHeader) + (long long)out_size * 2 * 4 +
         (long long)out_size * coeff_ksize(in_size, out_size) * 4;
}

// Fills a HOST buffer (mm_resize_coeff_bytes) with the table for one axis; the caller copies it to the device.
MM_API int mm_resize_coeff_build(void* host_buf, int in_size, int out_size) {
  MM_CHECK_ARG(host_buf && in_size >= 1 && out_size >= 1, "mm_resize_coeff_build: bad arguments");
  MM_CHECK_ARG(in_size <= (1 << 16), "mm_resize_coeff_build: axis of %d pixels is too long", in_size);
  const int ksize = coeff_ksize(in_size, out_size);
  ResizeCoeffHeader h = {in_size, out_size, ksize, 0};
  uint8_t* p8 = reinterpret_cast<uint8_t*>(host_buf);
  memcpy(p8, &h, sizeof(h));
  int* bounds = reinterpret_cast<int*>(p8 + sizeof(h));
  int* kk = bounds + (size_t)out_size * 2;
  // Resample.c precompute_coeffs (in0 = 0, in1 = in_size) + normalize_coeffs_8bpc, same operation order
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const double ss = 1.0 / filterscale;
  double* w = (double*)malloc(sizeof(double) * (size_t)ksize);
  MM_CHECK_ARG(w != nullptr, "mm_resize_coeff_build: out of host memory");
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
      if (x < xmax) {
        const double v = (ww != 0.0) ? w[x] / ww : w[x];
        k[x] = (v < 0) ? (int)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << RS_PRECISION_BITS));
      } else {
        k[x] = 0;
      }
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  free(w);
  return MM_OK;
}

// One RGB uint8 image [H][W][3] (device) -> [3][out][out] fp32 / bf16 (device). pad_square != 0: expand2square with
// `fill` first (train.py:1191-1203). coeff_x / coeff_y: DEVICE copies of mm_resize_coeff_build tables for the
// horizontal / vertical axis (in_size = the padded side, or W / H without padding). tmp: [rows][out][3] uint8 scratch
// with rows = padded side (or H). lut: 256 floats (device), the normalised value of every byte.
MM_API int mm_siglip_preprocess(const void* img, int H, int W, int pad_square, int fill, const void* coeff_x,
                                const void* coeff_y, int ksize_x, int ksize_y, int out_size, const float* lut,
                                void* tmp, void* out, int out_bf16, cudaStream_t stream) {
  MM_CHECK_ARG(img && coeff_x && coeff_y && lut && tmp && out, "mm_siglip_preprocess: null pointer");
  MM_CHECK_ARG(H >= 1 && W >= 1 && out_size >= 1 && out_size <= 4096, "mm_siglip_preprocess: bad sizes");
  const int side_w = pad_square ? (H > W ? H : W) : W;
  const int side_h = pad_square ? side_w : H;
  const int top = pad_square && W > H ? (W - H) / 2 : 0;
  const int left = pad_square && H > W ? (H - W) / 2 : 0;
  MM_CHECK_ARG(ksize_x == coeff_ksize(side_w, out_size) && ksize_y == coeff_ksize(side_h, out_size),
               "mm_siglip_preprocess: coefficient tables do not belong to these sizes");
  const uint8_t* cx = reinterpret_cast<const uint8_t*>(coeff_x) + sizeof(ResizeCoeffHeader);
  const uint8_t* cy = reinterpret_cast<const uint8_t*>(coeff_y) + sizeof(ResizeCoeffHeader);
  const int* bx = reinterpret_cast<const int*>(cx);
  const int* kx = bx + (size_t)out_size * 2;
  const int* by = reinterpret_cast<const int*>(cy);
  const int* ky = by + (size_t)out_size * 2;
  dim3 g1((out_size + RS_THREADS - 1) / RS_THREADS, side_h);
  resize_h_kernel<<<g1, RS_THREADS, 0, stream>>>((const uint8_t*)img, H, W, side_w, top, left, fill, bx, kx, ksize_x,
                                                 out_size, (uint8_t*)tmp);
  MM_CHECK_LAUNCH();
  dim3 g2((out_size + RS_THREADS - 1) / RS_THREADS, out_size);
  resize_v_lut_kernel<<<g2, RS_THREADS, 0, stream>>>((const uint8_t*)tmp, out_size, by, ky, ksize_y, out_size, lut, out,
                                                     out_bf16);
  MM_CHECK_LAUNCH();
  return MM_OK;
}
