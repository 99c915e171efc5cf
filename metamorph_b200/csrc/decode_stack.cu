// metamorph_b200 — whole-stack decode step in ONE persistent kernel.
//
// The per-op decode step (decode.cu) launches 8 kernels per layer; each weight-streaming GEMM there pays
// launch + pipeline-fill + tail for 34-235 MB of weights and reaches 3-4 TB/s of the 6.6 TB/s HBM roofline.
// Here all L decoder layers of one step (HF LlamaDecoderLayer, modeling_llama.py:292; called per step by the
// reference's greedy_decode, metamorph_llama.py:526-535) run in a single launch of one CTA per SM:
//   * warp 0 (one elected lane) is a TMA producer that streams EVERY weight matrix of EVERY layer, in the order
//     the tiles are consumed, through an 8 x 16 KB shared-memory ring. It never waits for activations, so the
//     HBM stream continues across phase boundaries (the ring refills while the consumers are in a grid barrier
//     or in the attention phase);
//   * warps 1..8 consume the ring with the weight-slab-as-A-operand mma.sync trick of skinny_gemm (16 weight
//     rows x 8 sequences per MMA), reduce across warps in shared memory and run the fused epilogues;
//   * phases per layer:  RMSNorm+QKV | RoPE + KV append + split-context attention (+ last-arriver combine) |
//     o_proj + residual | RMSNorm + gate/up + SwiGLU | down_proj + residual, separated by grid barriers
//     (monotonic counter, release/acquire at gpu scope);
//   * tiles are dealt round-robin (tile j*SMs + cta) for the whole rounds; the tiles left over after the last whole
//     round are shared stream-K style: their (subtile, k-stage) units are cut into one equal contiguous range per CTA,
//     and a tile that straddles a range boundary is completed by the CTA holding its first stage from fp32 partial
//     sums parked in the workspace (added in k order: deterministic). Every CTA streams the same number of weight
//     bytes (+- one 16 KB stage) in every phase. (Stream-K over the WHOLE phase balances equally well but streams
//     ~20 % slower on B200: 148 equally spaced contiguous ranges instead of a moving window of consecutive tiles.)
// Activations between phases ([8, H] rows) live in an L2-resident workspace and are read with ld.global.cg.
// The RMSNorm statistics reproduce rmsnorm_fwd_kernel's reduction order bit for bit.
#include "common.cuh"
#include <mutex>
#include <stdlib.h>
#include <string.h>

typedef CUresult (*PFN_encodeTiledDs)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);

namespace {

constexpr int DS_CONS = 256;                 // consumer threads (8 warps)
constexpr int DS_THREADS = 32 + DS_CONS;     // + producer warp
constexpr int DS_STAGE = 16384;              // [8 k64-blocks][16 rows][128 B], 128B-swizzled
constexpr int DS_NSTAGE = 8;
constexpr int DS_KS = 512;                   // k elements per stage
constexpr int DS_XMAXK = 4096;               // activations with K <= 4096 are staged in shared memory
constexpr int DS_XSTRIDE = DS_XMAXK * 2 + 64;  // bytes; +64 makes the (g, t) fragment reads conflict free
constexpr int DS_XBYTES = 8 * DS_XSTRIDE;
constexpr int DS_D = 128;                    // head dim
constexpr int DS_RED_BYTES = 2 * 8 * 16 * 8 * 4;
constexpr int DS_OFF_X = DS_NSTAGE * DS_STAGE;
constexpr int DS_OFF_RED = DS_OFF_X + DS_XBYTES;
constexpr int DS_OFF_MISC = DS_OFF_RED + DS_RED_BYTES;      // 64 floats of row statistics + flags
constexpr int DS_OFF_BAR = DS_OFF_MISC + 512;
constexpr int DS_SYNC_FLAGS = 384;           // word offset of the per-CTA partial-sum flags in the sync area
constexpr int DS_SMEM = DS_OFF_BAR + 2 * DS_NSTAGE * 8 + 1024;

struct DsParams {
  const CUtensorMap* maps;      // [L][4]: wqkv, wo, wgu, wd as (64 k, rows, K/64) tensors
  const bf16* const* ln1;       // [L]
  const bf16* const* ln2;       // [L]
  bf16* x;                      // [B][H] residual stream, in/out
  bf16* kcache;                 // [L][B][Hkv][Tmax][128]
  bf16* vcache;
  long long cache_layer_stride; // elements
  const int* pos;               // [B] index of the token being fed
  const float* cos_t;
  const float* sin_t;
  unsigned* sync;               // [0] barrier arrivals, [1] exits, [2] launch generation, [64 + b*Hkv + hk] combine
                                // counters, [DS_SYNC_FLAGS + cta] stream-K partial flags
  bf16* qkv;                    // [8][(Hq+2Hkv)*128]
  bf16* attn;                   // [8][Hq*128]
  bf16* hmid;                   // [8][H]
  bf16* act;                    // [8][I]
  float* part;                  // [B][Hkv][S][G][2+128]
  float* spart;                 // [SMs][2][128] stream-K partial sums (fp32)
  int L, B, H, Hq, Hkv, I, Tmax, S;
  float scale, eps;
  unsigned long long* trace;    // optional [L][16][nsm] globaltimer stamps (MM_DS_TRACE), else null
  int pf_stages;                // L2 prefetch distance of the weight stream, in 16 KB stages per CTA (MM_DS_PF)
  int dbg;                      // MM_DS_DBG timing ablations (results are wrong when set): 1 no attention,
                                // 2 no grid barriers, 4 no activation staging, 8 consumers skip the MMAs, 16 no K/V L2 prefetch
};

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c0,
                                            int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

// Grid-wide barrier over the consumer halves of all CTAs (the producers never take part: they only read weights).
// Arrival = fire-and-forget red.release on a monotonic counter, wait = ld.acquire polling of the same word (one L2
// round trip after the last arrival; a last-arriver-publishes variant measured slower because every CTA then waits
// for its atomic's return value). bar.sync orders the CTA's writes before thread 0's release (cumulativity), as in
// CUTLASS' arrive_inc / wait_eq. Bounded spin: a protocol bug or a non-resident CTA becomes a trap, not a hung GPU.
__device__ __forceinline__ void grid_barrier(unsigned* sync, unsigned idx, unsigned nsm, int ctid) {
  cons_sync();
  if (ctid == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sync) : "memory");
    const unsigned target = idx * nsm;
    const long long t0 = clock64();
    while (ld_acquire_u32(sync) < target) {
      if (clock64() - t0 > (1ll << 32)) {
        printf("decode_stack: grid barrier %u timeout (block %d, arrivals %u)\n", idx, (int)blockIdx.x,
               ld_acquire_u32(sync));
        __trap();
      }
    }
  }
  cons_sync();
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// K is only required to be a multiple of 64: the last 512-wide stage is zero-filled by TMA on the weight side
// (out-of-bounds k64 blocks) and here on the activation side.
__device__ __forceinline__ void zero_x_tail(uint8_t* xs, int K, int ctid) {
  const int kpad = (K + DS_KS - 1) / DS_KS * DS_KS;
  for (int v = (K >> 3) + ctid; v < (kpad >> 3); v += DS_CONS) {
#pragma unroll
    for (int b = 0; b < 8; ++b) *reinterpret_cast<int4*>(xs + b * DS_XSTRIDE + v * 16) = make_int4(0, 0, 0, 0);
  }
}

// ---- activation staging ---------------------------------------------------------------------------------------
// xs[b][k] (bf16, row stride DS_XSTRIDE) = RMSNorm(x[b])[k] * w[k]; same arithmetic and the same reduction order
// as rmsnorm_fwd_kernel (256 threads: thread t owns vectors t and t+256, shuffle tree, 8 warp partials, tree).
__device__ __forceinline__ void stage_x_norm(uint8_t* xs, float* stat, const bf16* __restrict__ x,
                                             const bf16* __restrict__ w, int B, int H, float eps, int ctid) {
  const int lane = ctid & 31, cw = ctid >> 5;
  const int nvec = H >> 3;
  int4 raw[8][2];
  float ss[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    ss[b] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = ctid + i * DS_CONS;
      raw[b][i] = (b < B && v < nvec) ? __ldcg(reinterpret_cast<const int4*>(x + (size_t)b * H + v * 8))
                                      : make_int4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t u[4] = {(uint32_t)raw[b][i].x, (uint32_t)raw[b][i].y, (uint32_t)raw[b][i].z,
                             (uint32_t)raw[b][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(u[j]);
        ss[b] += f.x * f.x + f.y * f.y;
      }
    }
    ss[b] = warp_sum(ss[b]);
    if (lane == 0) stat[b * 8 + cw] = ss[b];
  }
  // norm weights: issued before the reduction so that their latency hides behind it
  int4 wraw[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int v = ctid + i * DS_CONS;
    wraw[i] = (v < nvec) ? *reinterpret_cast<const int4*>(w + v * 8) : make_int4(0, 0, 0, 0);
  }
  cons_sync();
  float rstd[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float r = (lane < 8) ? stat[b * 8 + lane] : 0.f;
    r = warp_sum(r);
    rstd[b] = rsqrtf(r / (float)H + eps);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int v = ctid + i * DS_CONS;
    if (v < nvec) {
      const uint32_t wu[4] = {(uint32_t)wraw[i].x, (uint32_t)wraw[i].y, (uint32_t)wraw[i].z, (uint32_t)wraw[i].w};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const uint32_t u[4] = {(uint32_t)raw[b][i].x, (uint32_t)raw[b][i].y, (uint32_t)raw[b][i].z,
                               (uint32_t)raw[b][i].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // bf16(w * bf16(x * rstd)): the inner product is rounded to bf16 (one packed cvt), the outer product of
          // two bf16 values is exact in fp32, so the packed bf16 multiply rounds exactly like rmsnorm_fwd_kernel
          const float2 f = unpack_bf16x2(u[j]);
          const __nv_bfloat162 n2 = __floats2bfloat162_rn(f.x * rstd[b], f.y * rstd[b]);
          const __nv_bfloat162 w2 = *reinterpret_cast<const __nv_bfloat162*>(&wu[j]);
          const __nv_bfloat162 o2 = __hmul2(n2, w2);
          o[j] = *reinterpret_cast<const uint32_t*>(&o2);
        }
        *reinterpret_cast<int4*>(xs + b * DS_XSTRIDE + v * 16) = make_int4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  zero_x_tail(xs, H, ctid);
  cons_sync();
}

__device__ __forceinline__ void stage_x_plain(uint8_t* xs, const bf16* __restrict__ x, int ldx, int B, int K,
                                              int ctid) {
  const int nvec = K >> 3;
  for (int v = ctid; v < nvec; v += DS_CONS) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int4 r = (b < B) ? __ldcg(reinterpret_cast<const int4*>(x + (size_t)b * ldx + v * 8))
                             : make_int4(0, 0, 0, 0);
      *reinterpret_cast<int4*>(xs + b * DS_XSTRIDE + v * 16) = r;
    }
  }
  zero_x_tail(xs, K, ctid);
  cons_sync();
}

// ---- attention unit: one (sequence, kv head, context split) -------------------------------------------------
// Same arithmetic as decode_attn_kernel / decode_attn_combine_kernel (decode.cu); `sm` aliases the activation
// staging area. The CTA that completes the last split of a (sequence, kv head) merges the partials.
template <int G>
__device__ void attn_unit(const DsParams& p, float* sm, int* flag, int layer, int b, int hk, int sp, int ctid) {
  const int Hq = p.Hq, Hkv = p.Hkv, Tmax = p.Tmax, S = p.S;
  float* sq = sm;
  float* sknew = sq + G * DS_D;
  float* svnew = sknew + DS_D;
  float* sred = svnew + DS_D;
  float* s_m = sred + 8 * G * DS_D;
  float* s_l = s_m + 8;
  float* sscore = s_l + 8;
  const int pos = p.pos[b];
  const int n_ctx = pos + 1;
  const int chunk = (n_ctx + S - 1) / S;
  const int p0 = sp * chunk, p1 = min(n_ctx, p0 + chunk);
  const int cpad = ((Tmax + S - 1) / S + 4) & ~3;
  const int lane = ctid & 31, warp = ctid >> 5;
  const long long ldqkv = (long long)(Hq + 2 * Hkv) * DS_D;
  const bf16* row = p.qkv + (long long)b * ldqkv;
  bf16* kcb = p.kcache + (long long)layer * p.cache_layer_stride + ((long long)b * Hkv + hk) * Tmax * DS_D;
  bf16* vcb = p.vcache + (long long)layer * p.cache_layer_stride + ((long long)b * Hkv + hk) * Tmax * DS_D;
  const float* cp = p.cos_t + (long long)pos * (DS_D / 2);
  const float* sp_ = p.sin_t + (long long)pos * (DS_D / 2);
  const bool owns_new = (pos >= p0 && pos < p1);
  const float scale = p.scale;
  for (int i = ctid; i < G * (DS_D / 2); i += DS_CONS) {
    const int h = i / (DS_D / 2), j = i % (DS_D / 2);
    const bf16* qh = row + (long long)(hk * G + h) * DS_D;
    const float a = ldcg_bf16(qh + j), c = ldcg_bf16(qh + j + DS_D / 2);
    sq[h * DS_D + j] = __bfloat162float(__float2bfloat16(a * cp[j] - c * sp_[j])) * scale;
    sq[h * DS_D + j + DS_D / 2] = __bfloat162float(__float2bfloat16(c * cp[j] + a * sp_[j])) * scale;
  }
  if (owns_new && ctid < DS_D / 2) {
    const bf16* kh = row + (long long)(Hq + hk) * DS_D;
    const bf16* vh = row + (long long)(Hq + Hkv + hk) * DS_D;
    const float a = ldcg_bf16(kh + ctid), c = ldcg_bf16(kh + ctid + DS_D / 2);
    const bf16 k0 = __float2bfloat16(a * cp[ctid] - c * sp_[ctid]);
    const bf16 k1 = __float2bfloat16(c * cp[ctid] + a * sp_[ctid]);
    kcb[(long long)pos * DS_D + ctid] = k0;
    kcb[(long long)pos * DS_D + ctid + DS_D / 2] = k1;
    sknew[ctid] = __bfloat162float(k0);
    sknew[ctid + DS_D / 2] = __bfloat162float(k1);
    const float v0 = ldcg_bf16(vh + 2 * ctid), v1 = ldcg_bf16(vh + 2 * ctid + 1);
    vcb[(long long)pos * DS_D + 2 * ctid] = __float2bfloat16(v0);
    vcb[(long long)pos * DS_D + 2 * ctid + 1] = __float2bfloat16(v1);
    svnew[2 * ctid] = v0;
    svnew[2 * ctid + 1] = v1;
  }
  cons_sync();
  // ---- scores: a half-warp per cached position (16 lanes x 16 B = one 256-byte K row), 8 rows in flight per lane
  const int n_loc = max(p1 - p0, 0);
  const int half = lane >> 4, l16 = lane & 15;
  constexpr int AU = 8;
  {
    float qr[G][8];
#pragma unroll
    for (int h = 0; h < G; ++h)
#pragma unroll
      for (int j = 0; j < 8; ++j) qr[h][j] = sq[h * DS_D + l16 * 8 + j];
    for (int i0 = 0; i0 < n_loc; i0 += 2 * 8 * AU) {
      int4 raw[AU];
      int idx[AU];
#pragma unroll
      for (int u = 0; u < AU; ++u) {
        idx[u] = i0 + 2 * (warp + 8 * u) + half;
        const int qpos = p0 + idx[u];
        raw[u] = (idx[u] < n_loc && qpos != pos)
                     ? *reinterpret_cast<const int4*>(kcb + (long long)qpos * DS_D + l16 * 8)
                     : make_int4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < AU; ++u) {
        float kf[8];
        if (idx[u] < n_loc && p0 + idx[u] == pos) {
#pragma unroll
          for (int j = 0; j < 8; ++j) kf[j] = sknew[l16 * 8 + j];
        } else {
          const uint32_t w4[4] = {(uint32_t)raw[u].x, (uint32_t)raw[u].y, (uint32_t)raw[u].z, (uint32_t)raw[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w4[j]);
            kf[2 * j] = f.x;
            kf[2 * j + 1] = f.y;
          }
        }
#pragma unroll
        for (int h = 0; h < G; ++h) {
          float sdot = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) sdot += qr[h][j] * kf[j];
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
          if (l16 == 0 && idx[u] < n_loc) sscore[h * cpad + idx[u]] = sdot;
        }
      }
    }
  }
  cons_sync();
  for (int h = warp; h < G; h += DS_CONS / 32) {
    float mx = -INFINITY;
    for (int i = lane; i < n_loc; i += 32) mx = fmaxf(mx, sscore[h * cpad + i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < n_loc; i += 32) {
      const float e = __expf(sscore[h * cpad + i] - mx);
      sscore[h * cpad + i] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) { s_m[h] = mx; s_l[h] = sum; }
  }
  cons_sync();
  // ---- O_partial = P V: same half-warp-per-row mapping, the lane owns 8 output dims
  float o[G][8];
#pragma unroll
  for (int h = 0; h < G; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) o[h][j] = 0.f;
  for (int i0 = 0; i0 < n_loc; i0 += 2 * 8 * AU) {
    int4 raw[AU];
    int idx[AU];
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      idx[u] = i0 + 2 * (warp + 8 * u) + half;
      const int qpos = p0 + idx[u];
      raw[u] = (idx[u] < n_loc && qpos != pos)
                   ? *reinterpret_cast<const int4*>(vcb + (long long)qpos * DS_D + l16 * 8)
                   : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      if (idx[u] < n_loc) {
        float vf[8];
        if (p0 + idx[u] == pos) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vf[j] = svnew[l16 * 8 + j];
        } else {
          const uint32_t w4[4] = {(uint32_t)raw[u].x, (uint32_t)raw[u].y, (uint32_t)raw[u].z, (uint32_t)raw[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w4[j]);
            vf[2 * j] = f.x;
            vf[2 * j + 1] = f.y;
          }
        }
#pragma unroll
        for (int h = 0; h < G; ++h) {
          const float pr = sscore[h * cpad + idx[u]];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[h][j] += pr * vf[j];
        }
      }
    }
  }
#pragma unroll
  for (int h = 0; h < G; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[h][j] += __shfl_xor_sync(0xffffffffu, o[h][j], 16);
      if (half == 0) sred[(warp * G + h) * DS_D + l16 * 8 + j] = o[h][j];
    }
  cons_sync();
  float* pout = p.part + (((long long)b * Hkv + hk) * S + sp) * G * (2 + DS_D);
  for (int i = ctid; i < G * DS_D; i += DS_CONS) {
    const int h = i / DS_D, dcol = i % DS_D;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < DS_CONS / 32; ++w) acc += sred[(w * G + h) * DS_D + dcol];
    pout[h * (2 + DS_D) + 2 + dcol] = acc;
  }
  if (ctid < G) {
    pout[ctid * (2 + DS_D)] = n_loc > 0 ? s_m[ctid] : -INFINITY;
    pout[ctid * (2 + DS_D) + 1] = n_loc > 0 ? s_l[ctid] : 0.f;
  }
  // ---- last arriver merges the S partials of this (sequence, kv head)
  cons_sync();
  unsigned* cnt = p.sync + 64 + b * Hkv + hk;
  if (ctid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(cnt, 1u);
    *flag = (old == (unsigned)(S - 1)) ? 1 : 0;
    if (old == (unsigned)(S - 1)) {
      __threadfence();
      *cnt = 0u;                                     // every split has arrived: safe to re-arm for the next layer
    }
  }
  cons_sync();
  if (*flag) {
    const float* pin = p.part + (((long long)b * Hkv + hk) * S) * G * (2 + DS_D);
    for (int i = ctid; i < G * DS_D; i += DS_CONS) {
      const int h = i / DS_D, dcol = i % DS_D;
      float M = -INFINITY;
      for (int s = 0; s < S; ++s) M = fmaxf(M, __ldcg(pin + (s * G + h) * (2 + DS_D)));
      float Lsum = 0.f, acc = 0.f;
      for (int s = 0; s < S; ++s) {
        const float* ps = pin + (s * G + h) * (2 + DS_D);
        const float m_s = __ldcg(ps);
        const float w = (m_s == -INFINITY) ? 0.f : __expf(m_s - M);
        Lsum += __ldcg(ps + 1) * w;
        acc += __ldcg(ps + 2 + dcol) * w;
      }
      p.attn[(long long)b * (Hq * DS_D) + (long long)(hk * G + h) * DS_D + dcol] = __float2bfloat16(acc / Lsum);
    }
  }
  cons_sync();   // smem (sq, sscore, flag) is reused by the next unit
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DS_THREADS, 1) decode_stack_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + DS_OFF_BAR;
  const uint32_t bar_empty = bar_full + 8 * DS_NSTAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsm = gridDim.x, cta = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DS_NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, DS_CONS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int QKVN = (p.Hq + 2 * p.Hkv) * DS_D;
  // weight phases: rows (in 16-row tiles; gate/up as 32-row pairs), reduction length
  const int ntile[4] = {QKVN / 16, p.H / 16, p.I / 16, p.H / 16};
  const int nsub[4] = {1, 1, 2, 1};
  const int Kp[4] = {p.H, p.Hq * DS_D, p.H, p.I};

  if (warp == 0) {
    // =================================================== producer ===========================================
    if (lane == 0) {
      // Two cursors walk the same flattened (layer, phase, tile, k-stage) schedule: `ld` issues the TMA loads into the
      // ring as slots free up; `pf` runs p.pf_stages (MM_DS_PF, default 8) further ahead issuing L2 tensor prefetches,
      // so HBM keeps streaming for a while when the ring is full (grid barriers, staging, attention). Measured on B200
      // (512-step decode, ms/step): 0: 4.61, 4: 4.48, 8: 4.44, 12: 4.48, 16: 4.59, 24: 5.70 — a long run-ahead
      // evicts its own lines before they are used.
      // (all schedule arithmetic is 32-bit and incremental: a 64-bit division per stage on this single lane costs as
      // much as the stage itself)
      struct Cursor {
        int l, ph, j, nks;    // layer, phase, round (j < q: whole tile j*nsm+cta; j == q: stream-K share of the rest)
        int st, ks, left;     // current subtile, k-stage within it, stage units left in the current range
        bool done;
      };
      auto settle = [&](Cursor& c) {   // first non-empty range at or after (l, ph, j)
        while (!c.done) {
          c.nks = (Kp[c.ph] + DS_KS - 1) / DS_KS;
          const int TS = nsub[c.ph] * c.nks;
          const int q = ntile[c.ph] / nsm;
          if (c.j < q) {
            c.st = (c.j * nsm + cta) * nsub[c.ph];
            c.ks = 0;
            c.left = TS;
            return;
          }
          if (c.j == q) {
            const int Ur = (ntile[c.ph] - q * nsm) * TS;
            const int r0 = Ur * cta / nsm, r1 = Ur * (cta + 1) / nsm;
            if (r0 < r1) {
              c.st = q * nsm * nsub[c.ph] + r0 / c.nks;
              c.ks = r0 % c.nks;
              c.left = r1 - r0;
              return;
            }
          }
          c.j = 0;
          if (++c.ph == 4) { c.ph = 0; if (++c.l == p.L) c.done = true; }
        }
      };
      auto advance = [&](Cursor& c) {
        if (--c.left > 0) {
          if (++c.ks == c.nks) { c.ks = 0; ++c.st; }
          return;
        }
        ++c.j;
        settle(c);
      };
      auto kv_prefetch = [&](int l) {
        // K/V rows of this CTA's attention units of layer l -> L2 while the qkv weights of the layer stream
        const int units = p.B * p.Hkv * p.S;
        for (int u = cta; u < units; u += nsm) {
          const int sp = u % p.S, hk = (u / p.S) % p.Hkv, b = u / (p.S * p.Hkv);
          const int n_ctx = p.pos[b] + 1;
          const int chunk = (n_ctx + p.S - 1) / p.S;
          const int p0 = sp * chunk, p1 = min(n_ctx, p0 + chunk);
          if (p1 > p0) {
            const long long off = (long long)l * p.cache_layer_stride +
                                  (((long long)b * p.Hkv + hk) * p.Tmax + p0) * DS_D;
            const uint32_t bytes = (uint32_t)(p1 - p0) * DS_D * 2;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.kcache + off), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.vcache + off), "r"(bytes) : "memory");
          }
        }
      };
      auto prefetch_stage = [&](const Cursor& c) {
        const CUtensorMap* map = p.maps + c.l * 4 + c.ph;
        const int row0 = c.st * 16, kblk = c.ks * 8;
        asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                     ::"l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(row0), "r"(kblk) : "memory");
      };
      Cursor ld{0, 0, 0, 0, 0, 0, 0, false}, pf{0, 0, 0, 0, 0, 0, 0, false};
      settle(ld);
      settle(pf);
      int kv_layer = 0;
      if (!(p.dbg & 16)) kv_prefetch(0);
      for (int i = 0; i < p.pf_stages && !pf.done; ++i) { prefetch_stage(pf); advance(pf); }
      uint32_t sc = 0;
      while (!ld.done) {
        if (!(p.dbg & 16) && ld.l > kv_layer) {   // entering a new layer: its K/V arrive in L2 during the qkv phase
          kv_layer = ld.l;
          kv_prefetch(kv_layer);
        }
        const uint32_t slot = sc % DS_NSTAGE, par = (sc / DS_NSTAGE) & 1u;
        const CUtensorMap* map = p.maps + ld.l * 4 + ld.ph;
        const int row0 = ld.st * 16, kblk = ld.ks * 8;
        mbar_wait(bar_empty + 8 * slot, par ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * slot, DS_STAGE);
        tma_load_3d(base + slot * DS_STAGE, map, bar_full + 8 * slot, 0, row0, kblk);
        ++sc;
        advance(ld);
        if (!pf.done) { prefetch_stage(pf); advance(pf); }
      }
    }
    return;
  }

  // ===================================================== consumers ============================================
  const int ctid = threadIdx.x - 32;
  const int cw = ctid >> 5;                 // 0..7: k64-block of each stage
  const int g = lane >> 2, t4 = lane & 3;
  uint8_t* xs = sbase + DS_OFF_X;
  float* red = reinterpret_cast<float*>(sbase + DS_OFF_RED);
  float* stat = reinterpret_cast<float*>(sbase + DS_OFF_MISC);
  int* flag = reinterpret_cast<int*>(sbase + DS_OFF_MISC + 256);
  uint32_t sc = 0;
  unsigned bar_idx = 0;
  // flags of this launch: generation (bumped by the last CTA to leave) x phases per launch; compared for equality
  const unsigned epoch_base = ld_acquire_u32(p.sync + 2) * (unsigned)(4 * p.L + 1);
  int redbuf = 0;
  const int B = p.B;

#define DS_TRACE(ev)                                                                                  \
  do {                                                                                                \
    if (p.trace != nullptr && ctid == 0) p.trace[((size_t)l * 16 + (ev)) * nsm + cta] = globaltimer_ns(); \
  } while (0)
  for (int l = 0; l < p.L; ++l) {
    for (int ph = 0; ph < 4; ++ph) {
      // ---------------- phase prologue: activations
      const int K = Kp[ph];
      const bf16* xg = nullptr;      // K > DS_XMAXK: fragments straight from the L2-resident buffer
      if (p.dbg & 4) { if (ph == 3 && K > DS_XMAXK) xg = p.act; }
      else if (ph == 0) stage_x_norm(xs, stat, p.x, p.ln1[l], B, p.H, p.eps, ctid);
      else if (ph == 1) stage_x_plain(xs, p.attn, K, B, K, ctid);
      else if (ph == 2) stage_x_norm(xs, stat, p.hmid, p.ln2[l], B, p.H, p.eps, ctid);
      else {
        if (K <= DS_XMAXK) stage_x_plain(xs, p.act, K, B, K, ctid);
        else xg = (p.dbg & 32) ? nullptr : p.act;   // dbg 32: (wrong) fragments from shared memory, timing only
      }
      const int nks = (K + DS_KS - 1) / DS_KS;
      // Whole rounds of tiles, then a stream-K share of the rest: the rest's (subtile, k-stage) units are cut into nsm
      // equal contiguous ranges. A tile whose units straddle a range boundary is finished by the CTA that holds its
      // FIRST stage (it reaches that tile at the end of its range); the CTAs holding the later k-ranges reach it at
      // the start of theirs, park their fp32 partial sums in the workspace and raise a flag. Partials are added in
      // ascending-k order: deterministic.
      const int TS = nsub[ph] * nks;
      const int q = ntile[ph] / nsm;                       // whole rounds: tile j*nsm + cta (the access pattern that
      const int base_u = q * nsm * TS;                     // streams fastest); only the rest is shared stream-K style
      const int Ur = (ntile[ph] - q * nsm) * TS;
      const unsigned epoch = epoch_base + (unsigned)(l * 4 + ph + 1);
      DS_TRACE(ph * 4 + 0);        // staged, tiles start
      for (int j = 0; j <= q; ++j) {
      int u, u1;
      if (j < q) { u = (j * nsm + cta) * TS; u1 = u + TS; }
      else { u = base_u + Ur * cta / nsm; u1 = base_u + Ur * (cta + 1) / nsm; }
      while (u < u1) {
        const int T = u / TS;
        const int tile_end = (T + 1) * TS;
        const int end = u1 < tile_end ? u1 : tile_end;
        const bool head = (u == T * TS);
        float sums[2] = {0.f, 0.f};
        int uu = u;
        while (uu < end) {
          const int st = uu / nks, ks0 = uu % nks;
          const int sub = st - T * nsub[ph];
          const int sub_end = end < (st + 1) * nks ? end : (st + 1) * nks;
          const int ks1 = ks0 + (sub_end - uu);
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          auto consume = [&](const int4 (&xb)[2]) {
            const uint32_t slot = sc % DS_NSTAGE, par = (sc / DS_NSTAGE) & 1u;
            mbar_wait(bar_full + 8 * slot, par);
            const uint8_t* bx = sbase + slot * DS_STAGE + cw * 2048;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (p.dbg & 8) break;
              const int ch = c * 4 + t4;
              const int4 w0 = *reinterpret_cast<const int4*>(bx + g * 128 + ((ch ^ g) << 4));
              const int4 w1 = *reinterpret_cast<const int4*>(bx + (g + 8) * 128 + ((ch ^ g) << 4));
              const uint32_t a1[4] = {(uint32_t)w0.x, (uint32_t)w1.x, (uint32_t)w0.y, (uint32_t)w1.y};
              const uint32_t a2[4] = {(uint32_t)w0.z, (uint32_t)w1.z, (uint32_t)w0.w, (uint32_t)w1.w};
              mma_bf16_16816(acc, a1, (uint32_t)xb[c].x, (uint32_t)xb[c].y);
              mma_bf16_16816(acc, a2, (uint32_t)xb[c].z, (uint32_t)xb[c].w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
            ++sc;
          };
          if (xg == nullptr) {
            for (int ks = ks0; ks < ks1; ++ks) {
              const int kbase = (ks * DS_KS + cw * 64 + t4 * 8) & (DS_XMAXK - 1);   // (the mask only matters for dbg 32)
              const int4 xb[2] = {*reinterpret_cast<const int4*>(xs + g * DS_XSTRIDE + kbase * 2),
                                  *reinterpret_cast<const int4*>(xs + g * DS_XSTRIDE + (kbase + 32) * 2)};
              consume(xb);
            }
          } else {
            // activations too wide to stage whole (down_proj): every lane streams exactly the fragments it will feed to
            // its own MMAs through a private 8-deep cp.async ring in the (idle) staging area. Register prefetching
            // does not work here: in-flight loads share a handful of scoreboards, so waiting for the oldest one waits
            // for the newest too (measured: 35 us per layer for this phase vs 23 us with this ring).
            constexpr int DS_XP = 8;
            const uint32_t xring = smem_u32(xs) + ctid * 16;
            auto fetch_x = [&](int ks) {
              const uint32_t dst = xring + (ks % DS_XP) * 8192;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const int k0 = ks * DS_KS + cw * 64 + t4 * 8 + c * 32;
                const bool ok = (g < B && k0 < K && ks < ks1);
                cp_async16(dst + c * 4096, ok ? (const void*)(xg + (size_t)g * K + k0) : (const void*)xg, ok);
              }
              cp_async_commit();
            };
#pragma unroll
            for (int j = 0; j < DS_XP; ++j) fetch_x(ks0 + j);       // (stages past the end copy zeros: uniform groups)
            for (int ks = ks0; ks < ks1; ++ks) {
              cp_async_wait<DS_XP - 1>();
              const uint8_t* src = xs + ctid * 16 + (ks % DS_XP) * 8192;
              const int4 xb[2] = {*reinterpret_cast<const int4*>(src), *reinterpret_cast<const int4*>(src + 4096)};
              consume(xb);
              fetch_x(ks + DS_XP);
            }
            cp_async_wait<0>();
          }
          // ---- cross-warp reduction of this (part of a) subtile
          float* rb = red + redbuf * (8 * 16 * 8);
          redbuf ^= 1;
          rb[(cw * 16 + g) * 8 + 2 * t4] = acc[0];
          rb[(cw * 16 + g) * 8 + 2 * t4 + 1] = acc[1];
          rb[(cw * 16 + g + 8) * 8 + 2 * t4] = acc[2];
          rb[(cw * 16 + g + 8) * 8 + 2 * t4 + 1] = acc[3];
          cons_sync();
          if (ctid < 128) {
            const int r = ctid >> 3, b = ctid & 7;
            float psum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) psum += rb[(w * 16 + r) * 8 + b];
            if (sub == 0) sums[0] = psum; else sums[1] = psum;
          }
          uu = sub_end;
        }
        bool finish = head;
        if (!head) {
          // later k-range of a tile owned by an earlier CTA: park the partial sums, raise the flag
          if (ctid < 128) {
            p.spart[(cta * 2 + 0) * 128 + ctid] = sums[0];
            p.spart[(cta * 2 + 1) * 128 + ctid] = sums[1];
          }
          cons_sync();
          if (ctid == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sync + DS_SYNC_FLAGS + cta), "r"(epoch) : "memory");
        } else if (end != tile_end) {
          // owner of a tile whose later k-ranges belong to the next CTAs (c_first..c_last, possibly many when the
          // rest is smaller than one stage per CTA): poll their flags in parallel, then add the partials in k order
          const int c_first = cta + 1;
          int c_last = c_first;
          while (c_last + 1 < nsm && base_u + Ur * (c_last + 1) / nsm < tile_end) ++c_last;
          for (int c2 = c_first + ctid; c2 <= c_last; c2 += DS_CONS) {
            const int s2 = base_u + Ur * c2 / nsm, e2 = base_u + Ur * (c2 + 1) / nsm;
            if (e2 > s2) {
              const long long t0 = clock64();
              while (ld_acquire_u32(p.sync + DS_SYNC_FLAGS + c2) != epoch) {
                if (clock64() - t0 > (1ll << 32)) {
                  printf("decode_stack: partial-sum flag timeout (block %d waits for %d)\n", cta, c2);
                  __trap();
                }
              }
            }
          }
          cons_sync();
          if (ctid < 128) {
#pragma unroll 4
            for (int c2 = c_first; c2 <= c_last; ++c2) {
              const int s2 = base_u + Ur * c2 / nsm, e2 = base_u + Ur * (c2 + 1) / nsm;
              if (e2 > s2) {
                sums[0] += __ldcg(p.spart + (c2 * 2 + 0) * 128 + ctid);
                sums[1] += __ldcg(p.spart + (c2 * 2 + 1) * 128 + ctid);
              }
            }
          }
        }
        if (finish && ctid < 128) {
          const int r = ctid >> 3, b = ctid & 7;
          const int n = T * 16 + r;            // output feature (single-subtile phases) / SwiGLU channel (gate/up)
          if (ph == 0) {
            if (b < B) p.qkv[(size_t)b * QKVN + n] = __float2bfloat16(sums[0]);
          } else if (ph == 1) {
            if (b < B) p.hmid[(size_t)b * p.H + n] = __float2bfloat16(sums[0] + ldcg_bf16(p.x + (size_t)b * p.H + n));
          } else if (ph == 2) {
            if (b < B) p.act[(size_t)b * p.I + n] = __float2bfloat16(silu(sums[0]) * sums[1]);
          } else {
            if (b < B) p.x[(size_t)b * p.H + n] = __float2bfloat16(sums[0] + ldcg_bf16(p.hmid + (size_t)b * p.H + n));
          }
        }
        u = end;
      }
      }   // rounds
      DS_TRACE(ph * 4 + 1);        // tiles done
      const bool last = (l == p.L - 1 && ph == 3);
      if (!last && !(p.dbg & 2)) grid_barrier(p.sync, ++bar_idx, (unsigned)nsm, ctid);
      DS_TRACE(ph * 4 + 2);        // barrier passed
      if (ph == 0) {
        // ---------------- attention phase
        const int G = p.Hq / p.Hkv;
        const int units = B * p.Hkv * p.S;
        for (int u = cta; u < units && !(p.dbg & 1); u += nsm) {
          const int sp = u % p.S, hk = (u / p.S) % p.Hkv, b = u / (p.S * p.Hkv);
          float* sm = reinterpret_cast<float*>(xs);
          if (G == 4) attn_unit<4>(p, sm, flag, l, b, hk, sp, ctid);
          else if (G == 8) attn_unit<8>(p, sm, flag, l, b, hk, sp, ctid);
          else if (G == 2) attn_unit<2>(p, sm, flag, l, b, hk, sp, ctid);
          else attn_unit<1>(p, sm, flag, l, b, hk, sp, ctid);
        }
        DS_TRACE(3);               // attention units done
        if (!(p.dbg & 2)) grid_barrier(p.sync, ++bar_idx, (unsigned)nsm, ctid);
        DS_TRACE(7);               // barrier after attention passed
      }
    }
  }
  // the last CTA to leave re-arms the counters for the next launch
  cons_sync();
  if (ctid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(p.sync + 1, 1u);
    if (old == (unsigned)nsm - 1) {
      p.sync[0] = 0u;
      p.sync[1] = 0u;
      p.sync[2] = p.sync[2] + 1u;
      __threadfence();
    }
  }
}

PFN_encodeTiledDs ds_encoder() {
  static PFN_encodeTiledDs enc = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiledDs>(fn);
  });
  return enc;
}

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct DsWorkspace {
  long long sync, qkv, attn, hmid, act, part, spart, trace, total;
};
constexpr int DS_TRACE_LAYERS = 64;
DsWorkspace ds_workspace(int B, int H, int Hq, int Hkv, int I, int S) {
  DsWorkspace w;
  long long o = 0;
  w.sync = o; o = align_up(o + (DS_SYNC_FLAGS + mm_num_sms()) * 4, 256);
  w.qkv = o;  o = align_up(o + 8ll * (Hq + 2 * Hkv) * DS_D * 2, 256);
  w.attn = o; o = align_up(o + 8ll * Hq * DS_D * 2, 256);
  w.hmid = o; o = align_up(o + 8ll * H * 2, 256);
  w.act = o;  o = align_up(o + 8ll * I * 2, 256);
  w.part = o; o = align_up(o + (long long)B * Hkv * S * (Hq / Hkv) * (2 + DS_D) * 4, 256);
  w.spart = o; o = align_up(o + (long long)mm_num_sms() * 2 * 128 * 4, 256);
  w.trace = o; o = align_up(o + (long long)DS_TRACE_LAYERS * 16 * mm_num_sms() * 8, 256);
  w.total = o;
  return w;
}

int ds_splits(int B, int Hkv) {
  int s = mm_num_sms() / (B * Hkv);
  if (s < 1) s = 1;
  if (s > 16) s = 16;
  return s;
}

}  // namespace

// ---- C ABI ----------------------------------------------------------------------------------------------------
MM_API long long mm_decode_stack_plan_bytes(int n_layers) {
  return (long long)n_layers * 4 * (long long)sizeof(CUtensorMap) + (long long)n_layers * 2 * 8;
}

// Fills a HOST buffer of mm_decode_stack_plan_bytes(L) bytes (tensor maps of the four weight matrices of each
// layer + the two norm-weight pointers); the caller copies it to device memory (128-byte aligned) once.
MM_API int mm_decode_stack_plan_build(void* plan_host, int n_layers, const void* const* wqkv,
                                      const void* const* wo, const void* const* wgu, const void* const* wd,
                                      const void* const* ln1, const void* const* ln2, int hidden, int n_heads,
                                      int n_kv_heads, int head_dim, int intermediate) {
  MM_CHECK_ARG(plan_host && n_layers > 0, "mm_decode_stack_plan_build: bad arguments");
  MM_CHECK_ARG(head_dim == DS_D, "mm_decode_stack: head_dim must be 128");
  MM_CHECK_ARG(hidden % 64 == 0 && hidden <= DS_XMAXK && n_heads * head_dim <= DS_XMAXK && intermediate % 64 == 0,
               "mm_decode_stack: need hidden, heads*128 <= 4096 and hidden, intermediate %% 64 == 0");
  MM_CHECK_ARG(n_heads % n_kv_heads == 0, "mm_decode_stack: bad GQA grouping");
  const int G = n_heads / n_kv_heads;
  MM_CHECK_ARG(G == 1 || G == 2 || G == 4 || G == 8, "mm_decode_stack: GQA group must be 1, 2, 4 or 8");
  PFN_encodeTiledDs enc = ds_encoder();
  MM_CHECK_ARG(enc != nullptr, "mm_decode_stack: cuTensorMapEncodeTiled unavailable");
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(plan_host);
  const void** lnp = reinterpret_cast<const void**>(reinterpret_cast<uint8_t*>(plan_host) +
                                                    (size_t)n_layers * 4 * sizeof(CUtensorMap));
  const long long rows[4] = {(long long)(n_heads + 2 * n_kv_heads) * head_dim, hidden, 2ll * intermediate, hidden};
  const long long ks[4] = {hidden, (long long)n_heads * head_dim, hidden, intermediate};
  for (int l = 0; l < n_layers; ++l) {
    const void* w[4] = {wqkv[l], wo[l], wgu[l], wd[l]};
    for (int i = 0; i < 4; ++i) {
      MM_CHECK_ARG(w[i] != nullptr && ((uintptr_t)w[i] & 15) == 0, "mm_decode_stack: weight pointer unaligned");
      cuuint64_t dims[3] = {64, (cuuint64_t)rows[i], (cuuint64_t)(ks[i] / 64)};
      cuuint64_t strides[2] = {(cuuint64_t)ks[i] * 2, 128};
      cuuint32_t box[3] = {64, 16, 8};
      cuuint32_t estr[3] = {1, 1, 1};
      CUtensorMap tm;
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w[i]), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      MM_CHECK_ARG(r == CUDA_SUCCESS, "mm_decode_stack: cuTensorMapEncodeTiled failed (%d) layer %d matrix %d",
                   (int)r, l, i);
      memcpy(&maps[l * 4 + i], &tm, sizeof(CUtensorMap));
    }
    lnp[l] = ln1[l];
    lnp[n_layers + l] = ln2[l];
  }
  return MM_OK;
}

// byte offset of the MM_DS_TRACE region ([layers<=64][16 events][SMs] uint64 globaltimer stamps) in the workspace
MM_API long long mm_decode_stack_trace_offset(int B, int hidden, int n_heads, int n_kv_heads, int intermediate) {
  if (B < 1 || n_kv_heads < 1) return -1;
  return ds_workspace(B, hidden, n_heads, n_kv_heads, intermediate, ds_splits(B, n_kv_heads)).trace;
}

MM_API long long mm_decode_stack_workspace_bytes(int B, int hidden, int n_heads, int n_kv_heads, int intermediate) {
  if (B < 1 || n_kv_heads < 1) return 0;
  return ds_workspace(B, hidden, n_heads, n_kv_heads, intermediate, ds_splits(B, n_kv_heads)).total;
}

// One decode step through all layers. `workspace` must be zero-filled ONCE by the caller before the first call
// (barrier / combine counters; the kernel re-arms them itself). x [B][hidden] bf16 is updated in place.
MM_API int mm_decode_stack(const void* plan_dev, int n_layers, void* x, void* kcache, void* vcache,
                           long long cache_layer_stride, const int* pos, const float* cos_t, const float* sin_t,
                           int B, int hidden, int n_heads, int n_kv_heads, int intermediate, int Tmax, float scale,
                           float eps, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  MM_CHECK_ARG(plan_dev && ((uintptr_t)plan_dev & 127) == 0, "mm_decode_stack: plan must be 128-byte aligned");
  MM_CHECK_ARG(B >= 1 && B <= 8, "mm_decode_stack: batch must be in [1,8]");
  const int S = ds_splits(B, n_kv_heads);
  const int G = n_heads / n_kv_heads;
  const DsWorkspace ws = ds_workspace(B, hidden, n_heads, n_kv_heads, intermediate, S);
  MM_CHECK_ARG(workspace && ((uintptr_t)workspace & 255) == 0 && workspace_bytes >= ws.total,
               "mm_decode_stack: workspace too small or unaligned");
  const int cpad = ((Tmax + S - 1) / S + 4) & ~3;
  MM_CHECK_ARG((long long)(G * DS_D + 2 * DS_D + 8 * G * DS_D + 16 + G * cpad) * 4 <= DS_XBYTES,
               "mm_decode_stack: context of %d positions does not fit the attention scratch", Tmax);
  {  // the kernel's schedule arithmetic is 32-bit
    const long long k512 = (intermediate + DS_KS - 1) / DS_KS;
    const long long units = (long long)(intermediate / 16) * 2 * ((hidden + DS_KS - 1) / DS_KS) + (hidden / 16) * k512;
    MM_CHECK_ARG(units * mm_num_sms() < (1ll << 30), "mm_decode_stack: shape too large for the 32-bit tile schedule");
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(decode_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_SMEM);
  });
  MM_CHECK_CUDA(attr_err);
  DsParams p;
  uint8_t* w8 = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* plan8 = reinterpret_cast<const uint8_t*>(plan_dev);
  p.maps = reinterpret_cast<const CUtensorMap*>(plan8);
  p.ln1 = reinterpret_cast<const bf16* const*>(plan8 + (size_t)n_layers * 4 * sizeof(CUtensorMap));
  p.ln2 = p.ln1 + n_layers;
  p.x = (bf16*)x;
  p.kcache = (bf16*)kcache;
  p.vcache = (bf16*)vcache;
  p.cache_layer_stride = cache_layer_stride;
  p.pos = pos;
  p.cos_t = cos_t;
  p.sin_t = sin_t;
  p.sync = reinterpret_cast<unsigned*>(w8 + ws.sync);
  p.qkv = reinterpret_cast<bf16*>(w8 + ws.qkv);
  p.attn = reinterpret_cast<bf16*>(w8 + ws.attn);
  p.hmid = reinterpret_cast<bf16*>(w8 + ws.hmid);
  p.act = reinterpret_cast<bf16*>(w8 + ws.act);
  p.part = reinterpret_cast<float*>(w8 + ws.part);
  p.spart = reinterpret_cast<float*>(w8 + ws.spart);
  p.L = n_layers; p.B = B; p.H = hidden; p.Hq = n_heads; p.Hkv = n_kv_heads; p.I = intermediate; p.Tmax = Tmax;
  p.S = S; p.scale = scale; p.eps = eps;
  static const int dbg = getenv("MM_DS_DBG") ? atoi(getenv("MM_DS_DBG")) : 0;
  p.dbg = dbg;
  static const int pf_stages = getenv("MM_DS_PF") ? atoi(getenv("MM_DS_PF")) : 8;
  p.pf_stages = pf_stages;
  static const bool trace = getenv("MM_DS_TRACE") != nullptr;
  p.trace = (trace && n_layers <= DS_TRACE_LAYERS) ? reinterpret_cast<unsigned long long*>(w8 + ws.trace) : nullptr;
  // one CTA per SM, all co-resident (the grid barrier relies on it): the kernel needs > half of an SM's shared memory
  decode_stack_kernel<<<mm_num_sms(), DS_THREADS, DS_SMEM, stream>>>(p);
  MM_CHECK_LAUNCH();
  return MM_OK;
}
