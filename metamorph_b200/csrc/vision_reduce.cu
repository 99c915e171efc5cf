// metamorph_b200 — SigLIP feature reduction: 27x27 -> 8x8 bilinear + L2 normalise.
//
// Reference: SiglipVisionTower.forward (siglip_encoder.py:151-163, 206-208): view [N,h,w,C],
// permute to NCHW, F.interpolate(fp32, bilinear, align_corners=False), back to [N,n,C] in the input
// dtype, then F.normalize(p=2, dim=-1). That is 5 kernels + 4 copies; here it is one gather-lerp-
// normalise pass: each CTA produces one output token (C channels), reading its 2x2 source taps with
// 128-bit loads. Arithmetic follows ATen's upsample_bilinear2d (area_pixel_compute_source_index,
// same operation order) in fp32, result rounded to bf16 before the norm exactly like the reference.
#include "common.cuh"

namespace {

__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1,
                                          float& l0, float& l1) {
  float src = scale * (dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// x: [N, S*S, C]  ->  y: [N, T*T, C]
__global__ void __launch_bounds__(256)
bilinear_l2norm_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int S, int T, int C,
                       int normalize, float eps) {
  __shared__ float red[32];
  const int tok = blockIdx.x % (T * T);
  const int n = blockIdx.x / (T * T);
  const int oy = tok / T, ox = tok % T;
  const float scale = (float)S / (float)T;
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  src_index(oy, scale, S, y0, y1, hy0, hy1);
  src_index(ox, scale, S, x0, x1, wx0, wx1);
  const bf16* base = x + (size_t)n * S * S * C;
  const bf16* p00 = base + (size_t)(y0 * S + x0) * C;
  const bf16* p01 = base + (size_t)(y0 * S + x1) * C;
  const bf16* p10 = base + (size_t)(y1 * S + x0) * C;
  const bf16* p11 = base + (size_t)(y1 * S + x1) * C;
  const int nvec = C >> 3;
  // C <= 8 * 256 * 1 handled by one vector per thread (C=1152 -> 144 vectors)
  float v[8];
  float ss = 0.f;
  const int vi = threadIdx.x;
  if (vi < nvec) {
    const int4 a = *reinterpret_cast<const int4*>(p00 + vi * 8);
    const int4 b = *reinterpret_cast<const int4*>(p01 + vi * 8);
    const int4 c = *reinterpret_cast<const int4*>(p10 + vi * 8);
    const int4 d = *reinterpret_cast<const int4*>(p11 + vi * 8);
    const uint32_t ua[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
    const uint32_t ub[4] = {(uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
    const uint32_t uc[4] = {(uint32_t)c.x, (uint32_t)c.y, (uint32_t)c.z, (uint32_t)c.w};
    const uint32_t ud[4] = {(uint32_t)d.x, (uint32_t)d.y, (uint32_t)d.z, (uint32_t)d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(ua[j]), fb = unpack_bf16x2(ub[j]), fc = unpack_bf16x2(uc[j]),
                   fd = unpack_bf16x2(ud[j]);
      const float r0 = hy0 * (wx0 * fa.x + wx1 * fb.x) + hy1 * (wx0 * fc.x + wx1 * fd.x);
      const float r1 = hy0 * (wx0 * fa.y + wx1 * fb.y) + hy1 * (wx0 * fc.y + wx1 * fd.y);
      v[2 * j] = __bfloat162float(__float2bfloat16(r0));
      v[2 * j + 1] = __bfloat162float(__float2bfloat16(r1));
      ss += v[2 * j] * v[2 * j] + v[2 * j + 1] * v[2 * j + 1];
    }
  }
  float inv = 1.f;
  if (normalize) {
    ss = block_sum(ss, red);
    // torch computes the norm in the tensor dtype (bf16 result), clamps at eps, divides in bf16
    const float nrm = __bfloat162float(__float2bfloat16(sqrtf(ss)));
    inv = fmaxf(nrm, eps);
  }
  if (vi < nvec) {
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = normalize ? pack_bf16x2(v[2 * j] / inv, v[2 * j + 1] / inv)
                       : pack_bf16x2(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<int4*>(y + ((size_t)n * T * T + tok) * C + vi * 8) =
        make_int4(o[0], o[1], o[2], o[3]);
  }
}

// y[r,:] = x[r,:] / max(||x[r,:]||, eps) (bf16 semantics as above); one warp per row
__global__ void l2norm_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long R,
                                   int C, float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  for (long long r = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < R;
       r += (long long)gridDim.x * warps_per_block) {
    const bf16* xr = x + r * C;
    float ss = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      const int4 a = *reinterpret_cast<const int4*>(xr + v * 8);
      const uint32_t u[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(u[j]);
        ss += f.x * f.x + f.y * f.y;
      }
    }
    ss = warp_sum(ss);
    const float inv = fmaxf(__bfloat162float(__float2bfloat16(sqrtf(ss))), eps);
    for (int v = lane; v < nvec; v += 32) {
      const int4 a = *reinterpret_cast<const int4*>(xr + v * 8);
      const uint32_t u[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(u[j]);
        o[j] = pack_bf16x2(f.x / inv, f.y / inv);
      }
      *reinterpret_cast<int4*>(y + r * C + v * 8) = make_int4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace

MM_API int mm_bilinear_l2norm(const void* x, void* y, int n_img, int in_side, int out_side, int C,
                              int normalize, float eps, cudaStream_t stream) {
  MM_CHECK_ARG(n_img > 0 && in_side > 0 && out_side > 0 && C % 8 == 0 && C <= 2048,
               "mm_bilinear_l2norm: need C%%8==0 and C<=2048 (C=%d)", C);
  bilinear_l2norm_kernel<<<n_img * out_side * out_side, 256, 0, stream>>>(
      (const bf16*)x, (bf16*)y, in_side, out_side, C, normalize, eps);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_l2norm_rows(const void* x, void* y, long long R, int C, float eps, cudaStream_t stream) {
  MM_CHECK_ARG(R > 0 && C % 8 == 0, "mm_l2norm_rows: need C%%8==0");
  long long blocks = ceil_div64(R, 8);
  if (blocks > (long long)mm_num_sms() * 8) blocks = (long long)mm_num_sms() * 8;
  l2norm_rows_kernel<<<(int)blocks, 256, 0, stream>>>((const bf16*)x, (bf16*)y, R, C, eps);
  MM_CHECK_LAUNCH();
  return MM_OK;
}
