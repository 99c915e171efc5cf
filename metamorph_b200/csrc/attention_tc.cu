// metamorph_b200 — flash attention FORWARD on 5th-gen tensor cores (tcgen05 + TMEM + TMA), head_dim 128.
// (LLaMA causal GQA attention, HF modeling_llama.py:199-220 / SDPA in 4.45.)
//
// One CTA = 128 query rows of one (batch, q-head); K/V tiles of 128 keys stream through shared memory.
//   warp 0      : TMA producer   (Q once; K_j, V_j single-buffered, released by tcgen05.commit)
//   warp 1      : MMA issuer     S = Q K_j^T (SS, both K-major)   ->  TMEM cols [0,128)
//                                O += P_j V_j (TS: A = P in TMEM, B = V MN-major from smem) -> cols [128,256)
//   warps 2..5  : softmax        thread r owns query row r (TMEM lane r): no shuffles. Two passes over S
//                                (row max, then exp2 + row sum), P written back to TMEM as packed bf16
//                                aliasing the S columns already consumed. O is rescaled lazily (only when
//                                the running max grows by > 2^8, FA4-style), so the common tile never
//                                touches O in TMEM.
// 256 TMEM columns and 96 KB smem per CTA -> 2 CTAs/SM: while one CTA does softmax the other's MMAs run.
#include "common.cuh"
#include <mutex>
#include <stdlib.h>

namespace {

constexpr int TC_BR = 128, TC_BC = 128, TC_D = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_TILE_BYTES = 128 * 128 * 2;  // 32 KB
constexpr int TC_SMEM = 3 * TC_TILE_BYTES + 1024 + 128;
constexpr float kLog2e = 1.4426950408889634f;

struct TcFwdParams {
  bf16* o;
  float* lse;
  const int* seqlens;
  long long ldo;
  int B, T, Hq, Hkv;
  float scale;
  int causal;
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// K-major SW128 operand tile stored as two [128 rows x 64 elem] TMA boxes (16 KB each)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int k16) {
  const uint32_t addr = tile + (uint32_t)(k16 >> 2) * 16384u + (uint32_t)(k16 & 3) * 32u;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// MN-major SW128 operand tile: two [128 k-rows x 64 mn-elem] boxes; LBO = 16 KB between MN chunks
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int k16) {
  const uint32_t addr = tile + (uint32_t)k16 * 2048u;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(16384 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(TC_THREADS, 2)
flash_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + TC_TILE_BYTES, sV = base + 2 * TC_TILE_BYTES;
  const uint32_t bar = base + 3 * TC_TILE_BYTES;
  const uint32_t q_full = bar, k_full = bar + 8, k_empty = bar + 16, v_full = bar + 24, v_empty = bar + 32,
                 s_full = bar + 40, p_full = bar + 48, o_done = bar + 56, tmem_slot = bar + 64;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy (late) query tiles first: better tail behaviour under causal masking
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int q0 = qt * TC_BR;
  const int kv_len = p.seqlens ? p.seqlens[b] : p.T;
  int kv_end = kv_len;
  if (p.causal) kv_end = min(kv_end, q0 + TC_BR);
  const int n_tiles = (kv_end + TC_BC - 1) / TC_BC;
  const int tok0 = b * p.T;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (n_tiles > 0) {  // never leave a TMA write in flight when nobody will wait for it
      mbar_arrive_expect_tx(q_full, TC_TILE_BYTES);
      tma_load_2d(sQ, &tmap_q, q_full, h * TC_D, tok0 + q0);
      tma_load_2d(sQ + 16384, &tmap_q, q_full, h * TC_D + 64, tok0 + q0);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      mbar_wait(k_empty, ph ^ 1);
      mbar_arrive_expect_tx(k_full, TC_TILE_BYTES);
      tma_load_2d(sK, &tmap_k, k_full, hk * TC_D, tok0 + j * TC_BC);
      tma_load_2d(sK + 16384, &tmap_k, k_full, hk * TC_D + 64, tok0 + j * TC_BC);
      mbar_wait(v_empty, ph ^ 1);
      mbar_arrive_expect_tx(v_full, TC_TILE_BYTES);
      tma_load_2d(sV, &tmap_v, v_full, hk * TC_D, tok0 + j * TC_BC);
      tma_load_2d(sV + 16384, &tmap_v, v_full, hk * TC_D + 64, tok0 + j * TC_BC);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TC_BC >> 3) << 17) |
                             (uint32_t(TC_BR >> 4) << 24);                  // A,B K-major
    const uint32_t idesc_o = idesc_s | (1u << 16);                          // B (= V) MN-major
    if (n_tiles > 0) mbar_wait(q_full, 0);
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      mbar_wait(k_full, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < TC_D / 16; ++k)
        umma_bf16_ss(tS, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc_s, k != 0 ? 1u : 0u);
      umma_commit(k_empty);
      umma_commit(s_full);
      mbar_wait(p_full, ph);
      mbar_wait(v_full, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < TC_BC / 16; ++k)
        umma_bf16_ts(tO, tS + (uint32_t)k * 8u, desc_mnmajor(sV, k), idesc_o, (j | k) != 0 ? 1u : 0u);
      umma_commit(v_empty);
      umma_commit(o_done);
    }
  } else if (warp >= 2) {
    // ---------------------------------------------------------------- softmax / epilogue (row per thread)
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int row = q0 + r;  // query index inside the sequence
    const float sl2 = p.scale * kLog2e;
    float m_used = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      mbar_wait(s_full, ph);
      tcgen05_fence_after();
      const int c0 = j * TC_BC;
      const bool need_mask = (c0 + TC_BC > kv_len) || (p.causal && c0 + TC_BC - 1 > q0);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_off + c * 32, v);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const int col = c0 + c * 32 + t;
            const bool ok = (col < kv_len) && (!p.causal || col <= row);
            mx = fmaxf(mx, ok ? __uint_as_float(v[t]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(v[t]));
        }
      }
      mx *= sl2;  // scale > 0
      // lazy rescale of O: only when some row of this warp grew its max by more than 2^8
      const bool grow = (j > 0) && (mx > m_used + 8.f);
      if (j == 0) m_used = mx;
      if (__any_sync(0xffffffffu, grow)) {
        mbar_wait(o_done, ph ^ 1);  // PV of tile j-1 has landed in TMEM
        tcgen05_fence_after();
        const float m_new = fmaxf(m_used, mx);
        const float f = (m_new == -INFINITY) ? 1.f : exp2f(m_used - m_new);
        l_sum *= f;
        m_used = m_new;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) * f);
          tmem_st_32x32b_x32(tO + lane_off + c * 32, v);
        }
        tmem_st_wait();
      }
      const float m_eff = (m_used == -INFINITY) ? 0.f : m_used;
      // pass 2: P = exp2(S*scale - m), row sum, packed bf16 back into the consumed S columns
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_off + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          float e0 = exp2f(__uint_as_float(v[t]) * sl2 - m_eff);
          float e1 = exp2f(__uint_as_float(v[t + 1]) * sl2 - m_eff);
          if (need_mask) {
            const int col = c0 + c * 32 + t;
            if (!((col < kv_len) && (!p.causal || col <= row))) e0 = 0.f;
            if (!((col + 1 < kv_len) && (!p.causal || col + 1 <= row))) e1 = 0.f;
          }
          l_sum += e0 + e1;
          pk[t >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x16(tS + lane_off + c * 16, pk);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue
    float inv = 0.f;
    if (n_tiles > 0) {
      mbar_wait(o_done, (uint32_t)((n_tiles - 1) & 1));
      tcgen05_fence_after();
      inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
    }
    const bool row_ok = row < p.T;
    bf16* orow = p.o + (long long)(tok0 + row) * p.ldo + (long long)h * TC_D;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      if (n_tiles > 0) {
        tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int t = 0; t < 32; t += 8) {
          int4 o;
          o.x = pack_bf16x2(__uint_as_float(v[t]) * inv, __uint_as_float(v[t + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[t + 2]) * inv, __uint_as_float(v[t + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[t + 4]) * inv, __uint_as_float(v[t + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[t + 6]) * inv, __uint_as_float(v[t + 7]) * inv);
          *reinterpret_cast<int4*>(orow + c * 32 + t) = o;
        }
      }
    }
    if (p.lse != nullptr && row_ok)
      p.lse[((long long)b * p.Hq + h) * p.T + row] =
          l_sum > 0.f ? (m_used + log2f(l_sum)) / kLog2e : -INFINITY;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------------
// Forward v2: 1 CTA/SM, 16 softmax warps (warp -> TMEM lane quadrant x 32-column chunk), S double-buffered
// in TMEM so the tensor core computes S_{j+1} while the CUDA cores run softmax_j, K/V double-buffered in
// smem. One TMEM pass over S per tile (values stay in registers), one named barrier per tile to exchange the
// row maxima of the four column chunks.
// ---------------------------------------------------------------------------------------------------
constexpr int TC2_THREADS = 576;
constexpr int TC2_SMEM = 5 * TC_TILE_BYTES + 2 * 4 * 128 * 4 + 4 * 128 * 4 + 256 + 1024;

__global__ void __launch_bounds__(TC2_THREADS, 1)
flash_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;
  const uint32_t sK[2] = {base + TC_TILE_BYTES, base + 2 * TC_TILE_BYTES};
  const uint32_t sV[2] = {base + 3 * TC_TILE_BYTES, base + 4 * TC_TILE_BYTES};
  float* sMax = reinterpret_cast<float*>(base_ptr + 5 * TC_TILE_BYTES);   // [2][4][128]
  float* sSum = sMax + 2 * 4 * 128;                                       // [4][128]
  const uint32_t bar = base + 5 * TC_TILE_BYTES + 2 * 4 * 128 * 4 + 4 * 128 * 4;
  const uint32_t q_full = bar, k_full0 = bar + 8, k_full1 = bar + 16, k_empty0 = bar + 24, k_empty1 = bar + 32,
                 v_full0 = bar + 40, v_full1 = bar + 48, v_empty0 = bar + 56, v_empty1 = bar + 64,
                 s_full0 = bar + 72, s_full1 = bar + 80, p_full0 = bar + 88, p_full1 = bar + 96,
                 o_done = bar + 104, tmem_slot = bar + 112;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + 5 * TC_TILE_BYTES + 2 * 4 * 128 * 4 + 4 * 128 * 4 + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int q0 = qt * TC_BR;
  const int kv_len = p.seqlens ? p.seqlens[b] : p.T;
  int kv_end = kv_len;
  if (p.causal) kv_end = min(kv_end, q0 + TC_BR);
  const int n_tiles = (kv_end + TC_BC - 1) / TC_BC;
  const int tok0 = b * p.T;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    for (int i = 0; i < 9; ++i) mbar_init(bar + 8 * i, 1);   // q_full .. v_empty1
    mbar_init(s_full0, 1);
    mbar_init(s_full1, 1);
    mbar_init(p_full0, 16);
    mbar_init(p_full1, 16);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tSb[2] = {tmem, tmem + 128};
  const uint32_t tO = tmem + 256;

  if (warp == 0 && lane == 0) {
    if (n_tiles > 0) {
      mbar_arrive_expect_tx(q_full, TC_TILE_BYTES);
      tma_load_2d(sQ, &tmap_q, q_full, h * TC_D, tok0 + q0);
      tma_load_2d(sQ + 16384, &tmap_q, q_full, h * TC_D + 64, tok0 + q0);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int bf = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      mbar_wait(bf ? k_empty1 : k_empty0, ph ^ 1);
      mbar_arrive_expect_tx(bf ? k_full1 : k_full0, TC_TILE_BYTES);
      tma_load_2d(sK[bf], &tmap_k, bf ? k_full1 : k_full0, hk * TC_D, tok0 + j * TC_BC);
      tma_load_2d(sK[bf] + 16384, &tmap_k, bf ? k_full1 : k_full0, hk * TC_D + 64, tok0 + j * TC_BC);
      mbar_wait(bf ? v_empty1 : v_empty0, ph ^ 1);
      mbar_arrive_expect_tx(bf ? v_full1 : v_full0, TC_TILE_BYTES);
      tma_load_2d(sV[bf], &tmap_v, bf ? v_full1 : v_full0, hk * TC_D, tok0 + j * TC_BC);
      tma_load_2d(sV[bf] + 16384, &tmap_v, bf ? v_full1 : v_full0, hk * TC_D + 64, tok0 + j * TC_BC);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TC_BC >> 3) << 17) |
                             (uint32_t(TC_BR >> 4) << 24);
    const uint32_t idesc_o = idesc_s | (1u << 16);
    auto issue_s = [&](int j) {
      const int bf = j & 1;
      mbar_wait(bf ? k_full1 : k_full0, (uint32_t)((j >> 1) & 1));
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < TC_D / 16; ++k)
        umma_bf16_ss(tSb[bf], desc_kmajor(sQ, k), desc_kmajor(sK[bf], k), idesc_s, k != 0 ? 1u : 0u);
      umma_commit(bf ? k_empty1 : k_empty0);
      umma_commit(bf ? s_full1 : s_full0);
    };
    if (n_tiles > 0) {
      mbar_wait(q_full, 0);
      issue_s(0);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int bf = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      if (j + 1 < n_tiles) {
        // S buffer (j+1)&1 last held P_{j-1}: make sure PV_{j-1} has retired before overwriting it
        if (j >= 1) mbar_wait(o_done, (uint32_t)((j - 1) & 1));
        issue_s(j + 1);
      }
      mbar_wait(bf ? p_full1 : p_full0, ph);
      mbar_wait(bf ? v_full1 : v_full0, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < TC_BC / 16; ++k)
        umma_bf16_ts(tO, tSb[bf] + (uint32_t)k * 8u, desc_mnmajor(sV[bf], k), idesc_o, (j | k) != 0 ? 1u : 0u);
      umma_commit(bf ? v_empty1 : v_empty0);
      umma_commit(o_done);
    }
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int c = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int row = q0 + r;
    const float sl2 = p.scale * kLog2e;
    float m_used = -INFINITY, l_part = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int bf = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      mbar_wait(bf ? s_full1 : s_full0, ph);
      tcgen05_fence_after();
      const int c0 = j * TC_BC + c * 32;
      const bool need_mask = (j * TC_BC + TC_BC > kv_len) || (p.causal && j * TC_BC + TC_BC - 1 > q0);
      uint32_t v[32];
      tmem_ld_32x32b_x32(tSb[bf] + lane_off + c * 32, v);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (need_mask) {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int col = c0 + t;
          const bool ok = (col < kv_len) && (!p.causal || col <= row);
          const float x = ok ? __uint_as_float(v[t]) : -INFINITY;
          v[t] = __float_as_uint(x);
          mx = fmaxf(mx, x);
        }
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(v[t]));
      }
      float* mrow = sMax + (j & 1) * 512;
      mrow[c * 128 + r] = mx;
      asm volatile("bar.sync 1, 512;" ::: "memory");   // all S loads done (P may alias) + maxima visible
      mx = fmaxf(fmaxf(mrow[r], mrow[128 + r]), fmaxf(mrow[256 + r], mrow[384 + r])) * sl2;
      const bool grow = (j > 0) && (mx > m_used + 8.f);
      if (j == 0) m_used = mx;
      if (__any_sync(0xffffffffu, grow)) {
        mbar_wait(o_done, (uint32_t)((j - 1) & 1));
        tcgen05_fence_after();
        const float m_new = fmaxf(m_used, mx);
        const float f = (m_new == -INFINITY) ? 1.f : fast_exp2(m_used - m_new);
        l_part *= f;
        m_used = m_new;
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + lane_off + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * f);
        tmem_st_32x32b_x32(tO + lane_off + c * 32, o);
      }
      const float m_eff = (m_used == -INFINITY) ? 0.f : m_used;
      uint32_t pk[16];
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(v[t]), sl2, -m_eff));      // masked: exp2(-inf) = 0
        const float e1 = fast_exp2(fmaf(__uint_as_float(v[t + 1]), sl2, -m_eff));
        l_part += e0 + e1;
        pk[t >> 1] = pack_bf16x2(e0, e1);
      }
      tmem_st_32x32b_x16(tSb[bf] + lane_off + c * 16, pk);
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bf ? p_full1 : p_full0);
    }
    // epilogue: combine the four partial row sums, normalise, store
    sSum[c * 128 + r] = l_part;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const float l_sum = sSum[r] + sSum[128 + r] + sSum[256 + r] + sSum[384 + r];
    float inv = 0.f;
    uint32_t o[32];
    if (n_tiles > 0) {
      mbar_wait(o_done, (uint32_t)((n_tiles - 1) & 1));
      tcgen05_fence_after();
      inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
      tmem_ld_32x32b_x32(tO + lane_off + c * 32, o);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) o[t] = 0u;
    }
    if (row < p.T) {
      bf16* orow = p.o + (long long)(tok0 + row) * p.ldo + (long long)h * TC_D + c * 32;
#pragma unroll
      for (int t = 0; t < 32; t += 8) {
        int4 w;
        w.x = pack_bf16x2(__uint_as_float(o[t]) * inv, __uint_as_float(o[t + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o[t + 2]) * inv, __uint_as_float(o[t + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o[t + 4]) * inv, __uint_as_float(o[t + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o[t + 6]) * inv, __uint_as_float(o[t + 7]) * inv);
        *reinterpret_cast<int4*>(orow + t) = w;
      }
      if (c == 0 && p.lse != nullptr)
        p.lse[((long long)b * p.Hq + h) * p.T + row] = l_sum > 0.f ? (m_used + log2f(l_sum)) / kLog2e : -INFINITY;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

int make_tmap_rows(CUtensorMap* tm, const void* base, long long width, long long rows, long long ld) {
  static PFN_encodeTiled enc = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiled>(fn);
  });
  if (!enc) {
    mm_set_error("cuTensorMapEncodeTiled unavailable");
    return MM_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mm_set_error("cuTensorMapEncodeTiled failed (%d) for attention operand", (int)r);
    return MM_ERR_CUDA;
  }
  return MM_OK;
}

}  // namespace

// tcgen05 flash-attention forward, head_dim 128. Same argument meaning as mm_attn_fwd.
MM_API int mm_attn_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse,
                          const int* seqlens, long long ldq, long long ldk, long long ldv, long long ldo,
                          int B, int T, int Hq, int Hkv, int head_dim, int causal, float scale,
                          cudaStream_t stream) {
  MM_CHECK_ARG(head_dim == 128, "mm_attn_fwd_tc: head_dim must be 128");
  MM_CHECK_ARG(B > 0 && T > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "mm_attn_fwd_tc: bad head counts");
  MM_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 &&
                   ((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 &&
                   ((uintptr_t)o & 15) == 0, "mm_attn_fwd_tc: 16-byte alignment / pitch %% 8 required");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_rows(&tq, q, (long long)Hq * 128, (long long)B * T, ldq))) return rc;
  if ((rc = make_tmap_rows(&tk, k, (long long)Hkv * 128, (long long)B * T, ldk))) return rc;
  if ((rc = make_tmap_rows(&tv, v, (long long)Hkv * 128, (long long)B * T, ldv))) return rc;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] {
    err = cudaFuncSetAttribute(flash_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  });
  MM_CHECK_CUDA(err);
  TcFwdParams p;
  p.o = (bf16*)o; p.lse = lse; p.seqlens = seqlens; p.ldo = ldo;
  p.B = B; p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.scale = scale; p.causal = causal;
  dim3 grid((T + TC_BR - 1) / TC_BR, Hq, B);
  static const bool use_v1 = getenv("MM_ATTN_FWD_V1") != nullptr;
  if (use_v1) {
    flash_fwd_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(tq, tk, tv, p);
  } else {
    static std::once_flag once2;
    static cudaError_t err2 = cudaSuccess;
    std::call_once(once2, [&] {
      err2 = cudaFuncSetAttribute(flash_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM);
    });
    MM_CHECK_CUDA(err2);
    flash_fwd_tc2_kernel<<<grid, TC2_THREADS, TC2_SMEM, stream>>>(tq, tk, tv, p);
  }
  MM_CHECK_LAUNCH();
  return MM_OK;
}

// =====================================================================================================
// BACKWARD on tcgen05 (head_dim 128, causal, GQA).  One CTA = one 128-key tile of one kv head; it loops
// over the G query heads of the group and the query tiles at/after the key tile, keeping dK/dV in TMEM.
// Everything is computed in the TRANSPOSED orientation (lanes = keys, columns = queries) so that
//   S^T  = K  Q^T      (SS, both K-major)                      -> tST  [0,128)
//   dP^T = V dO^T      (SS, both K-major)                      -> tDPT [128,256)
//   P^T  = exp2(S^T*scale - lse[q])          elementwise, packed bf16 over the consumed S^T columns
//   dS^T = P^T o (dP^T - delta[q]) * scale   elementwise, packed bf16 over the consumed dP^T columns
//   dV  += P^T dO      (TS: A = P^T in TMEM,  B = dO MN-major) -> tDV  [256,384)
//   dK  += dS^T Q      (TS: A = dS^T in TMEM, B = Q  MN-major) -> tDK  [384,512)
//   dQ   = dS K        (SS: A = dS^T tile in smem read MN-major, B = K MN-major) -> tST, then red.add to HBM
// need no register transposes and no row reductions (lse and delta = rowsum(dO o O) are precomputed).
// Q / dO tiles are double-buffered by TMA; the same smem bytes serve as K-major and MN-major operands.
// =====================================================================================================
namespace {

constexpr int BT_THREADS = 576;  // TMA warp + MMA warp + 16 elementwise warps
constexpr int BT_TILE = 32768;
constexpr int BT_SMEM = 7 * BT_TILE + 2048 /*lse,delta x2*/ + 256 /*barriers*/ + 768 /*align slack*/;

struct TcBwdParams {
  const float* lse;    // lse * log2(e)   [B,Hq,T]
  const float* delta;  // rowsum(dO o O) * softmax_scale
  float* dq_accum;  // [B*T, Hq*128] fp32 (zeroed)
  bf16* dk;
  bf16* dv;
  const int* seqlens;
  long long lddk, lddv;
  int B, T, Hq, Hkv;
  float scale;
  int Tp;   // row pitch of lse / delta: T rounded up to whole 128-query tiles
};

// 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_dst),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(BT_THREADS, 1)
flash_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                    TcBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  if (threadIdx.x == 0 && base - smem_u32(smem_raw) > 768u) {
    printf("flash_bwd_tc_kernel: dynamic smem base misaligned beyond the reserved slack\n");
    __trap();
  }
  const uint32_t sK = base, sV = base + BT_TILE;
  const uint32_t sQ[2] = {base + 2 * BT_TILE, base + 4 * BT_TILE};
  const uint32_t sdO[2] = {base + 3 * BT_TILE, base + 5 * BT_TILE};
  const uint32_t sdS = base + 6 * BT_TILE;
  float* sStat = reinterpret_cast<float*>(base_ptr + 7 * BT_TILE);   // [2 buffers][lse 128 | delta 128]
  const uint32_t uStat = base + 7 * BT_TILE;
  const uint32_t bar = base + 7 * BT_TILE + 2048;
  const uint32_t kv_full = bar, qdo_full0 = bar + 8, qdo_full1 = bar + 16, qdo_empty0 = bar + 24,
                 qdo_empty1 = bar + 32, st_full = bar + 40, p_full = bar + 48, dq_full = bar + 56,
                 dq_empty = bar + 64, fin_full = bar + 72, mma_sync = bar + 80, tmem_slot = bar + 88;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + 7 * BT_TILE + 2048 + 88);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (Hkv*B, key tiles): CTAs are dispatched x-fastest, so ALL (head, batch) instances of the heaviest key
  // tile (j = 0 sees every query tile) start first and the light tiles fill the tail (LPT-style schedule)
  const int jt = blockIdx.y, hk = blockIdx.x % p.Hkv, b = blockIdx.x / p.Hkv;
  const int G = p.Hq / p.Hkv;
  const int kv0 = jt * 128;
  const int kv_len = p.seqlens ? p.seqlens[b] : p.T;
  const int tok0 = b * p.T;
  const int qt_begin = jt;                       // causal: query tiles at or after this key tile
  const int n_qt = (p.T + 127) / 128 - qt_begin;
  const int n_it = n_qt * G;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_do);
    mbar_init(kv_full, 1);
    mbar_init(qdo_full0, 1);
    mbar_init(qdo_full1, 1);
    mbar_init(qdo_empty0, 1);
    mbar_init(qdo_empty1, 1);
    mbar_init(st_full, 1);
    mbar_init(p_full, 16);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 16);
    mbar_init(fin_full, 1);
    mbar_init(mma_sync, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 384;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    mbar_arrive_expect_tx(kv_full, 2 * BT_TILE);
    tma_load_2d(sK, &tmap_k, kv_full, hk * 128, tok0 + kv0);
    tma_load_2d(sK + 16384, &tmap_k, kv_full, hk * 128 + 64, tok0 + kv0);
    tma_load_2d(sV, &tmap_v, kv_full, hk * 128, tok0 + kv0);
    tma_load_2d(sV + 16384, &tmap_v, kv_full, hk * 128 + 64, tok0 + kv0);
    for (int it = 0; it < n_it; ++it) {
      const int buf = it & 1, use = it >> 1;
      const int hq = hk * G + it / n_qt;
      const int q0 = (qt_begin + it % n_qt) * 128;
      const uint32_t full = buf ? qdo_full1 : qdo_full0, empty = buf ? qdo_empty1 : qdo_empty0;
      mbar_wait(empty, (uint32_t)((use & 1) ^ 1));
      mbar_arrive_expect_tx(full, 2 * BT_TILE + 1024);
      {
        const long long off = ((long long)b * p.Hq + hq) * p.Tp + q0;
        bulk_load(uStat + buf * 1024, p.lse + off, 512, full);
        bulk_load(uStat + buf * 1024 + 512, p.delta + off, 512, full);
      }
      tma_load_2d(sQ[buf], &tmap_q, full, hq * 128, tok0 + q0);
      tma_load_2d(sQ[buf] + 16384, &tmap_q, full, hq * 128 + 64, tok0 + q0);
      tma_load_2d(sdO[buf], &tmap_do, full, hq * 128, tok0 + q0);
      tma_load_2d(sdO[buf] + 16384, &tmap_do, full, hq * 128 + 64, tok0 + q0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(128 >> 3) << 17) |
                              (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_kmn = idesc_kk | (1u << 16);                 // B MN-major
    const uint32_t idesc_mnmn = idesc_kk | (1u << 15) | (1u << 16);   // A and B MN-major
    mbar_wait(kv_full, 0);
    for (int it = 0; it < n_it; ++it) {
      const int buf = it & 1, use = it >> 1;
      const uint32_t ph = (uint32_t)(it & 1);
      mbar_wait(buf ? qdo_full1 : qdo_full0, (uint32_t)(use & 1));
      mbar_wait(dq_empty, ph ^ 1);  // dQ of the previous iteration has been drained from tST
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ss(tST, desc_kmajor(sK, k), desc_kmajor(sQ[buf], k), idesc_kk, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ss(tDPT, desc_kmajor(sV, k), desc_kmajor(sdO[buf], k), idesc_kk, k != 0 ? 1u : 0u);
      umma_commit(st_full);
      mbar_wait(p_full, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ts(tDV, tST + (uint32_t)k * 8u, desc_mnmajor(sdO[buf], k), idesc_kmn, (it | k) != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ts(tDK, tDPT + (uint32_t)k * 8u, desc_mnmajor(sQ[buf], k), idesc_kmn, (it | k) != 0 ? 1u : 0u);
      // dQ overwrites tST, whose first 64 columns hold P^T (read by the dV MMAs): drain them first
      umma_commit(mma_sync);
      mbar_wait(mma_sync, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ss(tST, desc_mnmajor(sdS, k), desc_mnmajor(sK, k), idesc_mnmn, k != 0 ? 1u : 0u);
      umma_commit(buf ? qdo_empty1 : qdo_empty0);
      umma_commit(dq_full);
    }
    umma_commit(fin_full);
  } else if (warp >= 2) {
    // ---------------------------------------------------------------- elementwise / drains
    // 16 warps: warp -> (TMEM lane quadrant = warp%4, 32-column chunk c = (warp-2)/4). Every thread owns
    // one key row x 32 query columns of the tile, so all TMEM traffic and the exp2 work run 16-wide.
    const int quad = warp & 3;
    const int c = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int key_idx = kv0 + r;
    const bool key_ok = key_idx < kv_len;
    const float sl2 = p.scale * kLog2e;
    uint8_t* dS_row = base_ptr + 6 * BT_TILE + r * 128;
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int hq = hk * G + it / n_qt;
      const int q0 = (qt_begin + it % n_qt) * 128;
      // lse*log2e and delta*scale of this query tile arrive with the Q/dO TMA transaction
      const int sbuf = it & 1;
      const float* sLse = sStat + sbuf * 256;
      const float* sDelta = sLse + 128;
      mbar_wait(sbuf ? qdo_full1 : qdo_full0, (uint32_t)((it >> 1) & 1));
      mbar_wait(st_full, ph);
      tcgen05_fence_after();
      // two halves of 16 query columns keep the live register set small (no spills at 576 threads/CTA)
      uint32_t s0[16], s1[16], dp0[16], dp1[16];
      tmem_ld_32x32b_x16(tST + lane_off + c * 32, s0);
      tmem_ld_32x32b_x16(tST + lane_off + c * 32 + 16, s1);
      tmem_ld_32x32b_x16(tDPT + lane_off + c * 32, dp0);
      tmem_ld_32x32b_x16(tDPT + lane_off + c * 32 + 16, dp1);
      tmem_ld_wait();
      // P^T / dS^T are packed over columns that other warps of this quadrant are still loading:
      // every load of the tile must have completed before the first store.
      asm volatile("bar.sync 2, 512;" ::: "memory");
      const bool diag = q0 < kv0 + 128;            // tile touches the causal boundary
      const bool tail = q0 + 128 > p.T;            // tile overhangs the sequence end
      const bool fast = !diag && !tail;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const uint32_t(&sv)[16] = hf ? s1 : s0;
        const uint32_t(&dv_)[16] = hf ? dp1 : dp0;
        const int cb = c * 32 + hf * 16;           // first query column of this half
        uint32_t pp[8], dd[8];
        if (fast) {
          if (key_ok) {
#pragma unroll
            for (int t = 0; t < 16; t += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(sLse + cb + t);
              const float4 d4 = *reinterpret_cast<const float4*>(sDelta + cb + t);
              const float e0 = fast_exp2(fmaf(__uint_as_float(sv[t]), sl2, -l4.x));
              const float e1 = fast_exp2(fmaf(__uint_as_float(sv[t + 1]), sl2, -l4.y));
              const float e2 = fast_exp2(fmaf(__uint_as_float(sv[t + 2]), sl2, -l4.z));
              const float e3 = fast_exp2(fmaf(__uint_as_float(sv[t + 3]), sl2, -l4.w));
              pp[t >> 1] = pack_bf16x2(e0, e1);
              pp[(t >> 1) + 1] = pack_bf16x2(e2, e3);
              dd[t >> 1] = pack_bf16x2(e0 * fmaf(__uint_as_float(dv_[t]), p.scale, -d4.x),
                                       e1 * fmaf(__uint_as_float(dv_[t + 1]), p.scale, -d4.y));
              dd[(t >> 1) + 1] = pack_bf16x2(e2 * fmaf(__uint_as_float(dv_[t + 2]), p.scale, -d4.z),
                                             e3 * fmaf(__uint_as_float(dv_[t + 3]), p.scale, -d4.w));
            }
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) pp[t] = dd[t] = 0u;
          }
        } else {
#pragma unroll
          for (int t = 0; t < 16; t += 2) {
            float pv[2], dvv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int qq = cb + t + u;
              const int q_idx = q0 + qq;
              const bool ok = key_ok && (q_idx < p.T) && (key_idx <= q_idx);
              const float e = ok ? fast_exp2(fmaf(__uint_as_float(sv[t + u]), sl2, -sLse[qq])) : 0.f;
              pv[u] = e;
              dvv[u] = ok ? e * fmaf(__uint_as_float(dv_[t + u]), p.scale, -sDelta[qq]) : 0.f;   // pad statistics are undefined
            }
            pp[t >> 1] = pack_bf16x2(pv[0], pv[1]);
            dd[t >> 1] = pack_bf16x2(dvv[0], dvv[1]);
          }
        }
        tmem_st_32x32b_x8(tST + lane_off + c * 16 + hf * 8, pp);
        tmem_st_32x32b_x8(tDPT + lane_off + c * 16 + hf * 8, dd);
        // dS^T row -> smem in the SW128 MN-major layout (two [128 keys x 64 q] boxes, 16-byte chunks
        // XOR-swizzled with key%8) so that the dQ MMA can read it as its A operand.
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c8 = c * 4 + hf * 2 + i;
          const int slot = (c8 & 7) ^ (r & 7);
          *reinterpret_cast<int4*>(dS_row + (c8 >> 3) * 16384 + slot * 16) =
              make_int4(dd[4 * i], dd[4 * i + 1], dd[4 * i + 2], dd[4 * i + 3]);
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- drain dQ (lanes are query rows now, this warp's 32 d-columns): TMEM -> registers ->
      // red.global.add.v4.f32 into the fp32 accumulator. (Measured on B200 at the train shape: 3.71 ms per
      // launch; staging through smem + cp.reduce.async.bulk was slower, 4.36 ms, because it delays the
      // release of the Q/dO buffers and adds a smem round trip.)
      mbar_wait(dq_full, ph);
      tcgen05_fence_after();
      const int q_idx = q0 + r;
      float* dq_row = p.dq_accum + ((long long)(tok0 + q_idx) * p.Hq + hq) * 128 + c * 32;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tST + lane_off + c * 32, v);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);   // tST may be overwritten by the next S^T
      if (q_idx < p.T) {
#pragma unroll
        for (int t = 0; t < 32; t += 4)
          red_add_v4(dq_row + t, __uint_as_float(v[t]), __uint_as_float(v[t + 1]), __uint_as_float(v[t + 2]),
                     __uint_as_float(v[t + 3]));
      }
    }
    // ---- final: dK (already scaled through dS) and dV
    mbar_wait(fin_full, 0);
    tcgen05_fence_after();
    const bool row_ok = key_idx < p.T;
    bf16* dk_row = p.dk + (long long)(tok0 + key_idx) * p.lddk + (long long)hk * 128 + c * 32;
    bf16* dv_row = p.dv + (long long)(tok0 + key_idx) * p.lddv + (long long)hk * 128 + c * 32;
    {
      uint32_t a[32], d[32];
      tmem_ld_32x32b_x32(tDK + lane_off + c * 32, a);
      tmem_ld_32x32b_x32(tDV + lane_off + c * 32, d);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int t = 0; t < 32; t += 8) {
          int4 o;
          o.x = pack_bf16x2(__uint_as_float(a[t]), __uint_as_float(a[t + 1]));
          o.y = pack_bf16x2(__uint_as_float(a[t + 2]), __uint_as_float(a[t + 3]));
          o.z = pack_bf16x2(__uint_as_float(a[t + 4]), __uint_as_float(a[t + 5]));
          o.w = pack_bf16x2(__uint_as_float(a[t + 6]), __uint_as_float(a[t + 7]));
          *reinterpret_cast<int4*>(dk_row + t) = o;
          o.x = pack_bf16x2(__uint_as_float(d[t]), __uint_as_float(d[t + 1]));
          o.y = pack_bf16x2(__uint_as_float(d[t + 2]), __uint_as_float(d[t + 3]));
          o.z = pack_bf16x2(__uint_as_float(d[t + 4]), __uint_as_float(d[t + 5]));
          o.w = pack_bf16x2(__uint_as_float(d[t + 6]), __uint_as_float(d[t + 7]));
          *reinterpret_cast<int4*>(dv_row + t) = o;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace

// tcgen05 flash-attention backward; same contract as mm_attn_bwd (workspace from mm_attn_bwd_workspace_bytes).
MM_API int mm_attn_bwd_tc(const void* q, const void* k, const void* v, const void* o, const void* dout,
                          const float* lse, void* dq, void* dk, void* dv, const int* seqlens, long long ldq,
                          long long ldk, long long ldv, long long ldo, long long lddo, long long lddq,
                          long long lddk, long long lddv, int B, int T, int Hq, int Hkv, int head_dim,
                          float scale, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  MM_CHECK_ARG(head_dim == 128, "mm_attn_bwd_tc: head_dim must be 128");
  MM_CHECK_ARG(B > 0 && T > 0 && Hq % Hkv == 0, "mm_attn_bwd_tc: bad shape");
  const int Tp = (T + 127) / 128 * 128;
  const long long delta_bytes = ((long long)B * Hq * Tp * 4 + 255) / 256 * 256;
  MM_CHECK_ARG(workspace != nullptr &&
                   workspace_bytes >= 2 * delta_bytes + (long long)B * T * Hq * 128 * 4 + 1024,
               "mm_attn_bwd_tc: workspace too small (use mm_attn_bwd_workspace_bytes)");
  MM_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 &&
                   lddk % 8 == 0 && lddv % 8 == 0, "mm_attn_bwd_tc: pitches %% 8");
  float* delta = reinterpret_cast<float*>(workspace);
  float* dq_accum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + delta_bytes);
  MM_CHECK_CUDA(cudaMemsetAsync(dq_accum, 0, (size_t)B * T * Hq * 128 * 4, stream));
  float* lse2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(dq_accum) + (size_t)B * T * Hq * 128 * 4);
  int rc;
  if ((rc = mm_attn_bwd_delta_launch(o, dout, delta, ldo, lddo, B, T, Hq, head_dim, scale, lse, lse2, Tp, stream)))
    return rc;
  CUtensorMap tq, tk, tv, tdo;
  if ((rc = make_tmap_rows(&tq, q, (long long)Hq * 128, (long long)B * T, ldq))) return rc;
  if ((rc = make_tmap_rows(&tk, k, (long long)Hkv * 128, (long long)B * T, ldk))) return rc;
  if ((rc = make_tmap_rows(&tv, v, (long long)Hkv * 128, (long long)B * T, ldv))) return rc;
  if ((rc = make_tmap_rows(&tdo, dout, (long long)Hq * 128, (long long)B * T, lddo))) return rc;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] {
    err = cudaFuncSetAttribute(flash_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM);
  });
  MM_CHECK_CUDA(err);
  TcBwdParams p;
  p.lse = lse2; p.delta = delta; p.dq_accum = dq_accum; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.seqlens = seqlens;
  p.lddk = lddk; p.lddv = lddv; p.B = B; p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.scale = scale;
  p.Tp = Tp;
  dim3 grid(Hkv * B, (T + 127) / 128);
  flash_bwd_tc_kernel<<<grid, BT_THREADS, BT_SMEM, stream>>>(tq, tk, tv, tdo, p);
  MM_CHECK_LAUNCH();
  return mm_attn_bwd_convert_launch(dq_accum, dq, (long long)B * T, Hq * 128, lddq, stream);
}
