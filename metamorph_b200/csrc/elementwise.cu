// metamorph_b200 — HBM-bound elementwise kernels of the train step (128-bit accesses, grid-stride):
//   swiglu_bwd     : (gate|up interleaved, dact) -> act (recompute), d(gate|up)   [LlamaMLP bwd]
//   gelu fwd/bwd   : erf GELU of mm_projector / vision_head (projector builder.py:55-59)
//   colsum         : bias gradient (sum over rows)
//   im2col_patch14 : SigLIP patch-embed Conv2d(k=s=14) as GEMM operand (modeling_siglip.py:178-184)
//   add_pos_emb    : + learned position embedding
//   sumsq / scale  : gradient-norm pieces for clipping
#include "common.cuh"

namespace {

// gu: [M, 2I] with every 32-column chunk = [16 gate | 16 up]; dact/act: [M, I]
__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, const bf16* __restrict__ dact,
                                  bf16* __restrict__ dgu, bf16* __restrict__ act, long long M,
                                  long long I) {
  const long long chunks_per_row = I / 16;
  const long long total = M * chunks_per_row * 2;  // one thread = 8 gate + 8 up columns
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int halfsel = (int)(idx & 1);
    const long long c = (idx >> 1) % chunks_per_row;
    const long long row = (idx >> 1) / chunks_per_row;
    const bf16* gp = gu + row * 2 * I + c * 32 + halfsel * 8;
    const int4 graw = *reinterpret_cast<const int4*>(gp);
    const int4 uraw = *reinterpret_cast<const int4*>(gp + 16);
    const long long acol = c * 16 + halfsel * 8;
    const int4 draw = *reinterpret_cast<const int4*>(dact + row * I + acol);
    const uint32_t ug[4] = {(uint32_t)graw.x, (uint32_t)graw.y, (uint32_t)graw.z, (uint32_t)graw.w};
    const uint32_t uu[4] = {(uint32_t)uraw.x, (uint32_t)uraw.y, (uint32_t)uraw.z, (uint32_t)uraw.w};
    const uint32_t ud[4] = {(uint32_t)draw.x, (uint32_t)draw.y, (uint32_t)draw.z, (uint32_t)draw.w};
    uint32_t og[4], ou[4], oa[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = unpack_bf16x2(ug[j]);
      const float2 u = unpack_bf16x2(uu[j]);
      const float2 d = unpack_bf16x2(ud[j]);
      const float s0 = 1.f / (1.f + __expf(-g.x)), s1 = 1.f / (1.f + __expf(-g.y));
      const float a0 = g.x * s0, a1 = g.y * s1;  // silu(g)
      oa[j] = pack_bf16x2(a0 * u.x, a1 * u.y);
      og[j] = pack_bf16x2(d.x * u.x * (s0 + a0 * (1.f - s0)), d.y * u.y * (s1 + a1 * (1.f - s1)));
      ou[j] = pack_bf16x2(d.x * a0, d.y * a1);
    }
    bf16* dgp = dgu + row * 2 * I + c * 32 + halfsel * 8;
    *reinterpret_cast<int4*>(dgp) = make_int4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<int4*>(dgp + 16) = make_int4(ou[0], ou[1], ou[2], ou[3]);
    if (act != nullptr)
      *reinterpret_cast<int4*>(act + row * I + acol) = make_int4(oa[0], oa[1], oa[2], oa[3]);
  }
}

__global__ void gelu_fwd_kernel(const bf16* __restrict__ z, bf16* __restrict__ a, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int4 raw = *reinterpret_cast<const int4*>(z + i * 8);
    const uint32_t u[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(u[j]);
      o[j] = pack_bf16x2(gelu_erf(f.x), gelu_erf(f.y));
    }
    *reinterpret_cast<int4*>(a + i * 8) = make_int4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void gelu_bwd_kernel(const bf16* __restrict__ z, const bf16* __restrict__ da,
                                bf16* __restrict__ dz, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int4 raw = *reinterpret_cast<const int4*>(z + i * 8);
    const int4 graw = *reinterpret_cast<const int4*>(da + i * 8);
    const uint32_t u[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
    const uint32_t g[4] = {(uint32_t)graw.x, (uint32_t)graw.y, (uint32_t)graw.z, (uint32_t)graw.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(u[j]);
      const float2 d = unpack_bf16x2(g[j]);
      o[j] = pack_bf16x2(d.x * gelu_erf_grad(f.x), d.y * gelu_erf_grad(f.y));
    }
    *reinterpret_cast<int4*>(dz + i * 8) = make_int4(o[0], o[1], o[2], o[3]);
  }
}

// out[n] (+)= sum_r x[r, n]; one thread per column pair, rows strided over blockIdx.y
__global__ void colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long R,
                              long long N, long long ld) {
  const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= N) return;
  float s0 = 0.f, s1 = 0.f;
  for (long long r = blockIdx.y; r < R; r += gridDim.y) {
    const bf162 v = *reinterpret_cast<const bf162*>(x + r * ld + col);
    const float2 f = __bfloat1622float2(v);
    s0 += f.x;
    s1 += f.y;
  }
  atomicAdd(out + col, s0);
  if (col + 1 < N) atomicAdd(out + col + 1, s1);
}

// images [N,3,S,S] (bf16, NCHW) -> patches [N*(S/14)^2, ldp]; column = c*196 + ky*14 + kx
// (matches Conv2d weight [O, 3, 14, 14] flattened); columns >= 588 are zero padding.
__global__ void im2col_patch14_kernel(const bf16* __restrict__ img, bf16* __restrict__ out, int N,
                                      int S, int ldp) {
  const int G = S / 14;
  const long long rows = (long long)N * G * G;
  const long long total = rows * ldp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % ldp);
    const long long row = idx / ldp;
    bf16 v = __float2bfloat16(0.f);
    if (col < 588) {
      const int c = col / 196, rem = col % 196, ky = rem / 14, kx = rem % 14;
      const int n = (int)(row / (G * G)), p = (int)(row % (G * G)), py = p / G, px = p % G;
      v = img[(((size_t)n * 3 + c) * S + (py * 14 + ky)) * S + (px * 14 + kx)];
    }
    out[idx] = v;
  }
}

// x[r, :] += pos[r % P, :]
__global__ void add_pos_emb_kernel(bf16* __restrict__ x, const bf16* __restrict__ pos, long long R,
                                   int P, int H) {
  const long long nvec = R * (H / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (H / 8);
    const int v = (int)(i % (H / 8));
    const int4 a = *reinterpret_cast<const int4*>(x + r * H + v * 8);
    const int4 b = *reinterpret_cast<const int4*>(pos + (size_t)(r % P) * H + v * 8);
    const uint32_t ua[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
    const uint32_t ub[4] = {(uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(ua[j]);
      const float2 g = unpack_bf16x2(ub[j]);
      o[j] = pack_bf16x2(f.x + g.x, f.y + g.y);
    }
    *reinterpret_cast<int4*>(x + r * H + v * 8) = make_int4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void sumsq_bf16_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long n8) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int4 raw = ld_nc_int4(x + i * 8);
    const uint32_t u[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(u[j]);
      s += f.x * f.x + f.y * f.y;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

int ew_grid(long long work_items, int threads) {
  long long b = ceil_div64(work_items, threads);
  const long long cap = (long long)mm_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

MM_API int mm_swiglu_bwd(const void* gu, const void* dact, void* dgu, void* act, long long M,
                         long long I, cudaStream_t stream) {
  MM_CHECK_ARG(M > 0 && I > 0 && I % 16 == 0, "mm_swiglu_bwd: need I%%16==0 (I=%lld)", I);
  const long long total = M * (I / 16) * 2;
  swiglu_bwd_kernel<<<ew_grid(total, 256), 256, 0, stream>>>((const bf16*)gu, (const bf16*)dact,
                                                             (bf16*)dgu, (bf16*)act, M, I);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_gelu_fwd(const void* z, void* a, long long n, cudaStream_t stream) {
  MM_CHECK_ARG(n > 0 && n % 8 == 0, "mm_gelu_fwd: n%%8 != 0");
  gelu_fwd_kernel<<<ew_grid(n / 8, 256), 256, 0, stream>>>((const bf16*)z, (bf16*)a, n / 8);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_gelu_bwd(const void* z, const void* da, void* dz, long long n, cudaStream_t stream) {
  MM_CHECK_ARG(n > 0 && n % 8 == 0, "mm_gelu_bwd: n%%8 != 0");
  gelu_bwd_kernel<<<ew_grid(n / 8, 256), 256, 0, stream>>>((const bf16*)z, (const bf16*)da,
                                                           (bf16*)dz, n / 8);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_colsum_accum(const void* x, float* out, long long R, long long N, long long ld,
                           cudaStream_t stream) {
  MM_CHECK_ARG(R > 0 && N > 0 && N % 2 == 0 && ld % 2 == 0, "mm_colsum_accum: N, ld must be even");
  dim3 grid((unsigned)ceil_div64(N / 2, 128), (unsigned)(R < 64 ? R : 64));
  colsum_kernel<<<grid, 128, 0, stream>>>((const bf16*)x, out, R, N, ld);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_im2col_patch14(const void* img, void* out, int n_img, int image_size, int ldp,
                             cudaStream_t stream) {
  // Conv2d(k=14, s=14, padding="valid"): floor(S/14) patches per side (384 -> 27, last 6 px unused)
  MM_CHECK_ARG(n_img > 0 && image_size >= 14 && ldp >= 588 && ldp % 8 == 0,
               "mm_im2col_patch14: need image_size>=14, ldp>=588, ldp%%8==0");
  const int G = image_size / 14;
  const long long total = (long long)n_img * G * G * ldp;
  im2col_patch14_kernel<<<ew_grid(total, 256), 256, 0, stream>>>((const bf16*)img, (bf16*)out, n_img,
                                                                 image_size, ldp);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_add_pos_emb(void* x, const void* pos, long long R, int P, int H, cudaStream_t stream) {
  MM_CHECK_ARG(R > 0 && P > 0 && H % 8 == 0, "mm_add_pos_emb: H%%8 != 0");
  add_pos_emb_kernel<<<ew_grid(R * (H / 8), 256), 256, 0, stream>>>((bf16*)x, (const bf16*)pos, R, P, H);
  MM_CHECK_LAUNCH();
  return MM_OK;
}

MM_API int mm_sumsq_bf16_accum(const void* x, float* out, long long n, cudaStream_t stream) {
  MM_CHECK_ARG(n > 0 && n % 8 == 0, "mm_sumsq_bf16_accum: n%%8 != 0");
  sumsq_bf16_kernel<<<ew_grid(n / 8, 256), 256, 0, stream>>>((const bf16*)x, out, n / 8);
  MM_CHECK_LAUNCH();
  return MM_OK;
}
