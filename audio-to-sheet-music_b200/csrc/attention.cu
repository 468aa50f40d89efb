// Fused (flash-style) multi-head attention on tcgen05 for the CrossTransformerEncoder bottleneck
// (demucs transformer.py self / cross attention, 8 heads x 64, no mask, eval -> no dropout; call site
// /root/reference/src/models/stem_separation/ATHTDemucs_v2.py:228).  Replaces the unfused
// QK^T GEMM -> softmax -> PV GEMM sequence: scores never touch HBM.
//
// One CTA = one (segment, head, 128-query tile); 192 threads:
//   warp 0     : TMA producer  (Q once; K double-buffered, V single-buffered 128-key tiles, SWIZZLE_128B)
//   warp 1     : TMEM allocator + MMA issuer: S = Q K^T (M128 N128 K64) into TMEM cols [0,128),
//                PV = P V (M128 N64 K128, V consumed MN-major straight from its [key][d] tile) into cols [128,192)
//   warps 2..5 : online softmax, one query row per thread: tcgen05.ld S, running max / sum in fp32 with exp2,
//                P written as bf16 into a swizzled K-major smem tile, O accumulated in registers
// TMEM use is 256 columns and shared memory 97 KB, so two CTAs share an SM and one CTA's softmax
// overlaps the other's MMAs.
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace athtd {

struct FaParams {
  int Sq, Sk;
  float scale_log2;      // (1/sqrt(64)) * log2(e)
  bf16* O; long ldo;     // output rows [B*Sq, ldo], head h at columns h*64
};

static constexpr int FA_TILE = 16384;   // 128 rows x 64 bf16

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(192, 2)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + FA_TILE;            // 2 buffers
  uint8_t* sV = smem + 3 * FA_TILE;
  uint8_t* sP = smem + 4 * FA_TILE;        // 128 x 128 bf16 = two 64-key K-major atoms
  uint64_t* bars = (uint64_t*)(smem + 6 * FA_TILE);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 6,
           *s_full = bars + 7, *p_ready = bars + 8, *pv_full = bars + 9;
  uint32_t* tmem_slot = (uint32_t*)(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Sk + 127) / 128;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmV) : "memory");
    mbar_init(smem_u32(q_full), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&k_full[i]), 1); mbar_init(smem_u32(&k_empty[i]), 1); }
    mbar_init(smem_u32(v_full), 1); mbar_init(smem_u32(v_empty), 1);
    mbar_init(smem_u32(s_full), 1); mbar_init(smem_u32(p_ready), 128); mbar_init(smem_u32(pv_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), FA_TILE);
      tma_load_2d(smem_u32(sQ), &tmQ, smem_u32(q_full), h * 64, b * p.Sq + q0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(smem_u32(&k_empty[s]), (uint32_t)((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(smem_u32(&k_full[s]), FA_TILE);
        tma_load_2d(smem_u32(sK + s * FA_TILE), &tmK, smem_u32(&k_full[s]), h * 64, b * p.Sk + j * 128);
        mbar_wait(smem_u32(v_empty), (uint32_t)(j & 1) ^ 1u);
        mbar_expect_tx(smem_u32(v_full), FA_TILE);
        tma_load_2d(smem_u32(sV), &tmV, smem_u32(v_full), h * 64, b * p.Sk + j * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // S: M128 N128, A and B K-major.  PV: M128 N64, A K-major (P), B MN-major (V tile is [key][d], d contiguous)
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t dq = make_sw128_desc(smem_u32(sQ));
      const uint64_t dp = make_sw128_desc(smem_u32(sP));
      const uint64_t dv = make_sw128_desc(smem_u32(sV));
      mbar_wait(smem_u32(q_full), 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(smem_u32(&k_full[s]), (uint32_t)((j >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dk = make_sw128_desc(smem_u32(sK + s * FA_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
        umma_commit(smem_u32(&k_empty[s]));
        umma_commit(smem_u32(s_full));
        mbar_wait(smem_u32(p_ready), (uint32_t)(j & 1));
        mbar_wait(smem_u32(v_full), (uint32_t)(j & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: 16 keys = 32 B inside the (k/4)-th 64-key atom of P;  B: 16 key rows of V = 2048 B further
          const uint64_t da = dp + (uint64_t)((k >> 2) * (FA_TILE >> 4) + (k & 3) * 2);
          const uint64_t db = dv + (uint64_t)(k * (2048 >> 4));
          umma_bf16(tmem_PV, da, db, idesc_pv, k ? 1u : 0u);
        }
        umma_commit(smem_u32(v_empty));
        umma_commit(smem_u32(pv_full));
      }
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int nvalid = min(128, p.Sk - j * 128);
      mbar_wait(smem_u32(s_full), (uint32_t)(j & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float mx = -INFINITY;
      if (nvalid == 128) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), r);
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), r);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);
      const float mb = m_new * p.scale_log2;
      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), r);
        uint32_t pk[16];
        if (nvalid == 128) {          // all but the last key tile: no masking
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -mb));
            const float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -mb));
            l_tile += p0 + p1;
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            pk[i >> 1] = *(uint32_t*)&t;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < nvalid) ? ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -mb)) : 0.f;
            const float p1 = (c * 32 + i + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -mb)) : 0.f;
            l_tile += p0 + p1;
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            pk[i >> 1] = *(uint32_t*)&t;
          }
        }
        // K-major SWIZZLE_128B: row pitch 128 B, 16-byte chunk index XOR (row & 7); keys [64a, 64a+64) in atom a
        uint8_t* prow = sP + (c >> 1) * FA_TILE + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c & 1) * 4 + g) ^ (row & 7);
          *(uint4*)(prow + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA (async proxy)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(p_ready)) : "memory");
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      mbar_wait(smem_u32(pv_full), (uint32_t)(j & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_PV + lane_addr + (uint32_t)(c * 32), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(r[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (q0 + row < p.Sq) {
      const float inv = 1.0f / l_run;
      bf16* dst = p.O + ((long)b * p.Sq + q0 + row) * p.ldo + h * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 t = __floats2bfloat162_rn(o[8 * g + 2 * i] * inv, o[8 * g + 2 * i + 1] * inv);
          pk[i] = *(uint32_t*)&t;
        }
        *(uint4*)(dst + 8 * g) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

bool flash_attn_supported(long ldq, long ldkv, long ldo) {
  return tensor_map_api_available() && ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0;
}

// q: [B*Sq, ldq] (head h at columns h*64 of the pointer), k / v: [B*Sk, ldkv], o: [B*Sq, ldo]; all bf16.
int launch_flash_attn(const bf16* q, long ldq, const bf16* k, const bf16* v, long ldkv, int B, int Sq, int Sk, bf16* o,
                      long ldo, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  if (!make_tensor_map_2d(&tmQ, q, 512, (uint64_t)B * Sq, (uint64_t)ldq * 2, 64, 128)) return 2;
  if (!make_tensor_map_2d(&tmK, k, 512, (uint64_t)B * Sk, (uint64_t)ldkv * 2, 64, 128)) return 3;
  if (!make_tensor_map_2d(&tmV, v, 512, (uint64_t)B * Sk, (uint64_t)ldkv * 2, 64, 128)) return 4;
  FaParams p;
  p.Sq = Sq; p.Sk = Sk; p.scale_log2 = 0.125f * 1.4426950408889634f; p.O = o; p.ldo = ldo;
  const size_t smem = 1024 + 6 * FA_TILE + 128;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
  dim3 grid((Sq + 127) / 128, 8, B);
  flash_attn_kernel<<<grid, 192, smem, st>>>(tmQ, tmK, tmV, p);
  return 0;
}

}  // namespace athtd
