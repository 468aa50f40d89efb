// tcgen05 / TMEM / TMA GEMM for sm_100a: the tensor-core implementation of the conv-as-GEMM problem
// in its FLAT-ROW form.
//
//   acc[rho, n] = sum_{tap} sum_{k < Ktap}  A[rho + tapRow[tap], k] * B[n, tap*Ktap + k]          (bf16 in, fp32 accumulate)
//
// A is a 2-D view (rows x Ktap, row pitch >= 16 B) of a channels-last activation buffer; every conv of the
// path (1x1, strided k8s4, transposed k8s4 phases, dilated k3) is such a view plus 1..3 row-shifted taps, because
// the buffers carry zero pad rows / pad frames (DESIGN.md "flat-row GEMM").  Rows that fall in padding are
// computed and discarded by the epilogue's validity decode.  Ktap need not be a multiple of the 64-element K block:
// the TMA box is clipped by the tensor-map extent (zero fill) and only ceil(k/16) MMAs are issued.
//
// PERSISTENT, warp-specialised kernel: one CTA per SM loops over 128 x BN output tiles (n fastest, so the CTAs that
// share an A tile run together and it is fetched from HBM once):
//   warp 0      : TMA producer  - cp.async.bulk.tensor.2d (SWIZZLE_128B boxes 64 x 128 / 64 x BN) into a deep smem ring
//                 that runs ahead ACROSS tiles (hides the ~1 us load latency of the short-K layers)
//   warp 1      : TMEM allocator + MMA issuer - tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) into one of TWO TMEM
//                 accumulators; tcgen05.commit frees the smem stage / hands the accumulator to the epilogue
//   warps 2..17 : epilogue (4 warps per TMEM lane quarter, interleaved 32-column chunks; 2 per quarter in the
//                 two-CTAs-per-SM variant) - tcgen05.ld, fused bias / GroupNorm apply / GELU / GLU / LayerScale / residual /
//                 frequency-embedding / GroupNorm partial sums in packed fp32 pairs, 32-byte stores; overlaps the next tile's
//                 main loop through the second accumulator
// Both single-thread roles keep their warp converged around an elect.sync issuer and track the ring position with running
// counters: a runtime-divisor `%` / `/` per K block or a divergent `lane == 0` branch (descriptors through R2UR moves) made
// the issue thread, not the tensor pipe, the bottleneck of the long-K layers.
#include "gemm.cuh"
#include "tc_ptx.cuh"
#include <mutex>
#include <stdio.h>

namespace athtd {

static constexpr int TC_BM = 128;
static constexpr int TC_BK = 64;
static constexpr int TC_MAX_STAGES = 8;
// epilogue warps per CTA: the epilogue is instruction-issue / latency bound (2 warps per scheduler reached ~60 % issue
// utilisation and cost more than the MMAs of a K = 512 tile), so the one-CTA-per-SM variant runs 16 of them (4 per TMEM
// lane quarter); the two-CTAs-per-SM variant keeps 8 per CTA (16 per SM)
template <int MINB> struct TcCfg {
  // MINB == 3: one CTA per SM with 12 epilogue warps, for 192-wide tiles (6 chunks: two per warp; with 16 warps half of them took two
  // chunks and the other half waited at the per-tile barrier, and the per-tile row decode of the extra warps is a third of the work)
  static constexpr int EPI_WARPS = MINB == 1 ? 16 : MINB == 3 ? 12 : 8;
  static constexpr int CTAS = MINB == 2 ? 2 : 1;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};
static constexpr int TC_VEC = 256;       // staged per-tile column vectors (bias, GroupNorm weight / bias, column scale)

struct TcParams {
  int Mflat, N, BN, stages;
  int m_tiles, n_tiles;
  int ntaps, kb_per_tap, Ktap;
  int tapRow[3];
  // flat row -> (b, t', f') decode and validity
  int RpA, G2p, gpf, G2, vlo, vhi;
  // output row mapping: ((b*oG2p + t' + ogsh) * oRp + f' + orsh) * ldc
  int oG2p, ogsh, oRp, orsh;
  long ldc;
  void* C; int n_store; int no_store;
  const float* bias; const float* colscale;
  const void* res;
  const float* rowtab; float rowtab_scale;
  double* stats; int stat_mode; int statR;
  const float* gn_mr; const float* gn_w; const float* gn_b; int gn_mode;
  int convt_cout;
  int skip_lo, skip_hi;    // output columns [skip_lo, skip_hi) are computed (statistics) but not stored
  uint32_t fdRpA[2], fdG2p[2], fdNt[2];   // (multiplier, shift) of the division-free x / RpA, x / G2p, x / n_tiles (x < 2^31)
  float* fin_mr; double fin_count; int fin_n; unsigned* fin_counter;      // fused statistics finalisation (TcFlat)
  int b_res;               // weight-stationary: the whole-K B tile of this CTA's (fixed) n-tile is loaded ONCE; the ring carries only A
  int wide;                // 32-byte row accesses allowed (row pitch and base 32-byte aligned, skip range in 16-column units)
};

__device__ __forceinline__ void store8(bf16* dst, const float* v) {
  uint32_t pk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); pk[i] = *(uint32_t*)&t; }
  *(uint4*)dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}
// 32-byte (16 x bf16) accesses: one full sector per thread and instruction (sm_100 256-bit LDG / STG).  The row-per-thread
// epilogue cannot coalesce across lanes, so wider per-thread accesses halve the L2 request count of the store-heavy GEMMs.
__device__ __forceinline__ void store16(bf16* dst, const float* v) {
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); pk[i] = *(uint32_t*)&t; }
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
               "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
}
__device__ __forceinline__ void load16_add(const bf16* src, float* v) {
  uint32_t t[8];
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]),
               "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "l"(src));
#pragma unroll
  for (int e = 0; e < 8; ++e) { float2 f = __bfloat1622float2(*(const __nv_bfloat162*)&t[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
}
__device__ __forceinline__ void load8_add(const bf16* src, float* v) {
  uint4 t = *(const uint4*)src;
  const __nv_bfloat162* h = (const __nv_bfloat162*)&t;
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
}

// epilogue feature flags (compile-time): unused stages must not cost issue slots (an if-converted generic epilogue
// issued ~4400 instructions per 32-column chunk and dominated the kernel)
enum { EF_GELU = 1, EF_GLU = 2, EF_GN = 4, EF_POST = 8, EF_STATS = 16 };

// 1 / (1 + 2^(-x log2 e)) with the two approximate SFU ops (relative error ~2^-22; +-inf saturate to 0 / 1): the IEEE-rounded
// __frcp_rn / range-checked __expf pair cost 17 instructions and a branch per gate
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// per-row state of one epilogue thread for one tile
struct EpiRow {
  bool valid, edge_lo, edge_hi;      // edge_*: first / last input row of a transposed conv
  long orow;
  int b, m;
  float gmean, grstd;
};

// One 32-column chunk of one accumulator row: r[] = raw fp32 accumulators of columns [ncol, ncol+32).
// sv: per-tile column vectors staged in shared memory, indexed relative to the tile's first column c0 = ncol - n0
// ([0] bias, [1] GroupNorm weight, [2] GroupNorm bias: accumulator columns; [3] column scale: output columns)
template <int EF>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const EpiRow& er, const uint32_t (&r)[32], int ncol, int nc,
                                               const float* __restrict__ sv, int c0, float& ssum, float& ssq) {
  constexpr int NV = (EF & EF_GLU) ? 16 : 32;
  const int Nout = (EF & EF_GLU) ? p.N / 2 : p.N;
  float v[32];
#pragma unroll
  for (int j4 = 0; j4 < 32; j4 += 4) {                     // columns past N hold garbage and are never stored
    const float4 b4 = *(const float4*)(sv + c0 + j4);      // broadcast reads (every lane owns a row, same columns)
    // packed fp32 pairs (FADD2 / FFMA2): the epilogue is issue-bound, the accumulator registers already come in pairs
    float2 x01 = f2add(make_float2(__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1])), make_float2(b4.x, b4.y));
    float2 x23 = f2add(make_float2(__uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3])), make_float2(b4.z, b4.w));
    if (EF & EF_GN) {
      const float4 w4 = *(const float4*)(sv + TC_VEC + c0 + j4), o4 = *(const float4*)(sv + 2 * TC_VEC + c0 + j4);
      const float2 nm = f2splat(-er.gmean), rs = f2splat(er.grstd);
      x01 = f2fma(f2mul(f2add(x01, nm), rs), make_float2(w4.x, w4.y), make_float2(o4.x, o4.y));
      x23 = f2fma(f2mul(f2add(x23, nm), rs), make_float2(w4.z, w4.w), make_float2(o4.z, o4.w));
    }
    if (EF & EF_GELU) { x01 = gelu_fast2(x01); x23 = gelu_fast2(x23); }
    v[j4] = x01.x; v[j4 + 1] = x01.y; v[j4 + 2] = x23.x; v[j4 + 3] = x23.y;
  }
  int no = ncol, nco = nc;             // output column base / count
  if (EF & EF_GLU) {
    no = ncol >> 1; nco = nc >> 1;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = v[2 * j] * sigmoid_fast(v[2 * j + 1]);
  }
  if (EF & EF_POST) {
    if (p.colscale) {
      const float* cs = sv + 3 * TC_VEC + ((EF & EF_GLU) ? (c0 >> 1) : c0);
#pragma unroll
      for (int j4 = 0; j4 < NV; j4 += 4) {
        const float4 c4 = *(const float4*)(cs + j4);
        const float2 a = f2mul(make_float2(v[j4], v[j4 + 1]), make_float2(c4.x, c4.y));
        const float2 b = f2mul(make_float2(v[j4 + 2], v[j4 + 3]), make_float2(c4.z, c4.w));
        v[j4] = a.x; v[j4 + 1] = a.y; v[j4 + 2] = b.x; v[j4 + 3] = b.y;
      }
    }
    if (p.rowtab) {      // per-row table (frequency embedding): Nout % 4 == 0, 16-byte gathers instead of scalar ones
      const float4* tp = (const float4*)(p.rowtab + (long)er.m * Nout + no);
#pragma unroll
      for (int g = 0; g < NV / 4; ++g)
        if (4 * g < nco) {
          const float4 t4 = __ldg(tp + g);
          v[4 * g] += p.rowtab_scale * t4.x; v[4 * g + 1] += p.rowtab_scale * t4.y;
          v[4 * g + 2] += p.rowtab_scale * t4.z; v[4 * g + 3] += p.rowtab_scale * t4.w;
        }
    }
    if (p.res) {
      const bf16* rp = (const bf16*)p.res + er.orow * p.ldc + no;
      if (p.wide) {
#pragma unroll
        for (int g = 0; g < NV / 16; ++g) {
          if (16 * g + 16 <= nco) load16_add(rp + 16 * g, v + 16 * g);
          else if (16 * g < nco) load8_add(rp + 16 * g, v + 16 * g);
        }
      } else {
#pragma unroll
        for (int g = 0; g < NV / 8; ++g)
          if (8 * g < nco) load8_add(rp + 8 * g, v + 8 * g);
      }
    }
  }
  if (EF & EF_STATS) {
    // 16-column groups (N, the chunk starts and the transposed conv's C_out are multiples of 16, so a group never straddles
    // a phase).  Transposed conv: output rows -2, -1 (q == 0, phases 0, 1) and 4F, 4F+1 (last q, phases 2, 3) are cropped and
    // do not count; the phase of a group is uniform across the warp, so this is two selects per group, no divergence.
#pragma unroll
    for (int g = 0; g < NV / 16; ++g) {
      float2 s2 = f2splat(0.f), q2 = f2splat(0.f);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const bool full = 16 * g + 16 <= nco;                                               // ragged last group (N % 32 == 8 / GLU)
        const float2 x = make_float2(full || 16 * g + j < nco ? v[16 * g + j] : 0.f, full || 16 * g + j + 1 < nco ? v[16 * g + j + 1] : 0.f);
        s2 = f2add(s2, x); q2 = f2fma(x, x, q2);
      }
      const float s = s2.x + s2.y, q = q2.x + q2.y;
      bool counted = 16 * g < nco;
      if (p.convt_cout > 0) counted = counted && !((no + 16 * g < 2 * p.convt_cout) ? er.edge_lo : er.edge_hi);
      if (counted) { ssum += s; ssq += q; }
    }
  }
  if (!p.no_store) {
    bf16* cp = (bf16*)p.C + er.orow * p.ldc + no;
    const int nst_cols = min(nco, p.n_store - no);
    if (p.wide) {
#pragma unroll
      for (int g = 0; g < NV / 16; ++g) {
        const int c = no + 16 * g;
        if (16 * g + 16 <= nst_cols) {
          if (!(c >= p.skip_lo && c + 16 <= p.skip_hi)) store16(cp + 16 * g, v + 16 * g);
        } else if (16 * g < nst_cols) {
          if (!(c >= p.skip_lo && c + 8 <= p.skip_hi)) store8(cp + 16 * g, v + 16 * g);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < NV / 8; ++g) {
        const int c = no + 8 * g;
        if (8 * g < nst_cols && !(c >= p.skip_lo && c + 8 <= p.skip_hi)) store8(cp + 8 * g, v + 8 * g);
      }
    }
  }
}

// All chunks of one accumulator tile that belong to this epilogue warp (interleaved 32-column chunks, `step` columns apart).
template <int EF>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, const EpiRow& er, uint32_t tacc, int n0, int c_first, int step,
                                              const float* __restrict__ sv, float& ssum, float& ssq) {
  for (int c0 = c_first; c0 < p.BN; c0 += step) {
    uint32_t r[32];
    tmem_ld32(tacc + (uint32_t)c0, r);
    const int ncol = n0 + c0;
    const int nc = min(32, min(p.BN - c0, p.N - ncol));      // valid accumulator columns in this chunk (multiple of 8)
    if (er.valid && nc > 0) epilogue_chunk<EF>(p, er, r, ncol, nc, sv, c0, ssum, ssq);
  }
}

// Residual rows of the NEXT tile of this CTA are pulled into the L2 while the current tile is processed (the residual stream, e.g.
// the token buffer of out_proj, has left the L2 by the time it is read back: the row-per-thread loads then sat on the full DRAM
// latency, 37 % of the stall samples of those launches).  One 128-byte line per thread: `sub` walks the lines of the row segment.
__device__ __forceinline__ void prefetch_res_l2(const TcParams& p, bool glu, int rho, int n0, int sub, int nsub) {
  const int q2 = fast_div(rho, p.fdRpA);
  const int fp = rho - q2 * p.RpA;
  const int b = fast_div(q2, p.fdG2p);
  const int tp = q2 - b * p.G2p;
  if (!(rho < p.Mflat && fp >= p.vlo && fp < p.vhi && tp >= p.gpf && tp < p.gpf + p.G2)) return;
  const long orow = ((long)b * p.oG2p + tp + p.ogsh) * p.oRp + fp + p.orsh;
  const int Nout = glu ? p.N >> 1 : p.N, no = glu ? n0 >> 1 : n0, nco = min(glu ? p.BN >> 1 : p.BN, Nout - no);
  const char* rp = (const char*)((const bf16*)p.res + orow * p.ldc + no);
  for (int off = 128 * sub; off < 2 * nco; off += 128 * nsub) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + off));
}

// ------------------------------------------------------------------ kernel
template <int EF, int MINB>
__global__ void __launch_bounds__(TcCfg<MINB>::THREADS, TcCfg<MINB>::CTAS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  pdl_trigger();                 // the next kernel's CTAs may set up while this grid drains (programmatic dependent launch)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET on the shared-memory array: a round trip through uintptr_t makes every later access a
  // generic LD / ST (long-scoreboard, queued behind the global loads) instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stageA = TC_BM * TC_BK * 2;
  const int stageB = p.BN * TC_BK * 2;             // BN multiple of 8 -> multiple of 1024 B
  const int nst = p.stages;
  const int nkb_all = p.ntaps * p.kb_per_tap;
  uint8_t* sA = smem;
  uint8_t* sB = smem + nst * stageA;               // b_res: nkb_all resident K blocks of B instead of a ring
  uint64_t* bars = (uint64_t*)(sB + (p.b_res ? nkb_all : nst) * stageB);
  uint64_t* full = bars, *empty = bars + TC_MAX_STAGES, *tfull = bars + 2 * TC_MAX_STAGES, *tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  uint64_t* bres_full = bars + 2 * TC_MAX_STAGES + 6;      // (bytes 176..183 of the barrier block: free)
  float* svec = (float*)(bars + 2 * TC_MAX_STAGES + 8);        // [2 tiles][4 vectors][TC_VEC]
  constexpr int EPI_WARPS = TcCfg<MINB>::EPI_WARPS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.ntaps * p.kb_per_tap;
  const int n_total_tiles = p.m_tiles * p.n_tiles;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.BN)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < nst; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull[i]), 1); mbar_init(smem_u32(&tempty[i]), EPI_WARPS); }
    mbar_init(smem_u32(bres_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                    // barrier init / TMEM allocation above overlap the predecessor; its data is visible from here

  if (warp == 0) {
    {
      const bool issuer = elect_one();            // converged warp, one lane issues (see the MMA role)
      const uint32_t bytes = (uint32_t)(stageA + (p.b_res ? 0 : stageB));
      if (p.b_res && (int)blockIdx.x < n_total_tiles) {
        // weight-stationary: gridDim.x is a multiple of n_tiles, so t % n_tiles (this CTA's n-tile) never changes: its B tile is
        // fetched once for all K blocks and stays; short-K layers streamed more weight bytes than activation bytes per tile
        const int n0 = ((int)blockIdx.x - fast_div((int)blockIdx.x, p.fdNt) * p.n_tiles) * p.BN;
        if (issuer) {
          mbar_expect_tx(smem_u32(bres_full), (uint32_t)(nkb_all * stageB));
          int tap = 0, kin = 0;
          for (int kb = 0; kb < nkb_all; ++kb) {
            tma_load_2d(smem_u32(sB + kb * stageB), &tmB, smem_u32(bres_full), tap * p.Ktap + kin, n0);
            kin += TC_BK;
            if (kin >= p.Ktap) { kin = 0; ++tap; }
          }
        }
        __syncwarp();
      }
      // ring position as running counters: no integer divisions on the single-thread critical path (an `it % stages`,
      // `it / stages`, `kb / kb_per_tap` per K block cost more latency than the MMAs of a narrow tile take to execute)
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < n_total_tiles; t += gridDim.x) {
        const int mt_ = fast_div(t, p.fdNt);
        const int row0 = mt_ * TC_BM;
        const int n0 = (t - mt_ * p.n_tiles) * p.BN;
        int tap = 0, kin = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full[s]);
          if (issuer) {
            mbar_expect_tx(fb, bytes);
            tma_load_2d(smem_u32(sA + s * stageA), &tmA, fb, kin, row0 + p.tapRow[tap]);
            if (!p.b_res) tma_load_2d(smem_u32(sB + s * stageB), &tmB, fb, tap * p.Ktap + kin, n0);
          }
          __syncwarp();
          kin += TC_BK;
          if (kin >= p.Ktap) { kin = 0; ++tap; }
          if (++s == nst) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // the whole warp walks the loop (converged: uniform-datapath address arithmetic), one elected lane issues
      const bool issuer = elect_one();
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int i = 0, s = 0; uint32_t ph = 0;
      const uint64_t da0 = make_sw128_desc(smem_u32(sA)), db0 = make_sw128_desc(smem_u32(sB));
      const uint64_t dstepA = (uint64_t)(stageA >> 4), dstepB = (uint64_t)(stageB >> 4);     // descriptor address field is in 16-B units
      if (p.b_res && (int)blockIdx.x < n_total_tiles) {
        mbar_wait(smem_u32(bres_full), 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int t = blockIdx.x; t < n_total_tiles; t += gridDim.x, ++i) {
        const int buf = i & 1;
        mbar_wait(smem_u32(&tempty[buf]), ((uint32_t)(i >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(buf * p.BN);
        int kin = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&full[s]), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = da0 + (uint64_t)s * dstepA, db = db0 + (uint64_t)(p.b_res ? kb : s) * dstepB;
          const int nmma = (min(TC_BK, p.Ktap - kin) + 15) >> 4;     // columns past Ktap are zero-filled by TMA
          if (issuer) {
            for (int k = 0; k < nmma; ++k)      // advance 16 bf16 = 32 B inside the 128-B swizzle atom
              umma_bf16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            umma_commit(smem_u32(&empty[s]));
          }
          __syncwarp();
          kin += TC_BK;
          if (kin >= p.Ktap) kin = 0;
          if (++s == nst) { s = 0; ph ^= 1u; }
        }
        if (issuer) umma_commit(smem_u32(&tfull[buf]));
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue warps: thread <-> accumulator row (TMEM lane), EPI_WARPS / 4 warps per lane quarter
    // taking interleaved 32-column chunks
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int etid = threadIdx.x - 64;
    const int Nout_ = (EF & EF_GLU) ? p.N / 2 : p.N;
    // Per-tile column vectors (bias, GroupNorm weight / bias, column scale; clamped at the last column: columns past N are
    // never stored) are staged in shared memory ONE TILE AHEAD: the global loads of tile i+1 are issued before tile i's
    // accumulator is processed and stashed after it, so their latency is off the epilogue's per-tile dependency chain (that
    // chain, not the MMAs, bounds every short-K layer).  BN <= TC_VEC <= epilogue threads: one column per thread.
    float vr0 = 0.f, vr1 = 0.f, vr2 = 0.f, vr3 = 0.f;
    auto fetch_vecs = [&](int n0) {
      if (etid < p.BN) {
        const int n = min(n0 + etid, p.N - 1);
        vr0 = p.bias ? __ldg(p.bias + n) : 0.f;
        if (EF & EF_GN) { vr1 = __ldg(p.gn_w + n); vr2 = __ldg(p.gn_b + n); }
        if ((EF & EF_POST) && p.colscale) {
          const int no = (EF & EF_GLU) ? (n0 >> 1) + etid : n0 + etid;
          vr3 = __ldg(p.colscale + min(no, Nout_ - 1));
        }
      }
    };
    auto stash_vecs = [&](float* sv) {
      if (etid < p.BN) {
        sv[etid] = vr0;
        if (EF & EF_GN) { sv[TC_VEC + etid] = vr1; sv[2 * TC_VEC + etid] = vr2; }
        if ((EF & EF_POST) && p.colscale) sv[3 * TC_VEC + etid] = vr3;
      }
    };
    if ((int)blockIdx.x < n_total_tiles) {
      fetch_vecs(((int)blockIdx.x % p.n_tiles) * p.BN); stash_vecs(svec);
      if ((EF & EF_POST) && p.res) {      // the first tile's residual rows arrive in the L2 behind its main loop
        const int mt0 = fast_div((int)blockIdx.x, p.fdNt);
        prefetch_res_l2(p, (EF & EF_GLU) != 0, mt0 * TC_BM + row, ((int)blockIdx.x - mt0 * p.n_tiles) * p.BN, sub, EPI_WARPS / 4);
      }
    }
    int i = 0;
    for (int t = blockIdx.x; t < n_total_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      const int mt_ = fast_div(t, p.fdNt);
      const int row0 = mt_ * TC_BM;
      const int n0 = (t - mt_ * p.n_tiles) * p.BN;
      const float* sv = svec + buf * 4 * TC_VEC;
      EpiRow er;
      const int rho = row0 + row;                   // Mflat < 2^31: 32-bit row decode
      const int q2 = fast_div(rho, p.fdRpA);
      const int fp = rho - q2 * p.RpA;
      er.b = fast_div(q2, p.fdG2p);
      const int tp = q2 - er.b * p.G2p;
      er.valid = rho < p.Mflat && fp >= p.vlo && fp < p.vhi && tp >= p.gpf && tp < p.gpf + p.G2;
      er.orow = ((long)er.b * p.oG2p + tp + p.ogsh) * p.oRp + fp + p.orsh;
      er.m = fp - p.vlo;
      er.gmean = 0.f; er.grstd = 1.f;
      if ((EF & EF_GN) && er.valid) {
        const long gi = p.gn_mode == STAT_PER_G1_M ? (long)er.b * p.statR + er.m : (long)er.b;
        er.gmean = p.gn_mr[2 * gi]; er.grstd = p.gn_mr[2 * gi + 1];
      }
      // tile i's vectors (stashed during tile i-1) become visible; every warp is done reading the other buffer
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
      const int tn = t + (int)gridDim.x;
      if (tn < n_total_tiles) fetch_vecs((tn - fast_div(tn, p.fdNt) * p.n_tiles) * p.BN);
      er.edge_lo = er.m == 0; er.edge_hi = er.m == p.vhi - p.vlo - 1;
      float ssum = 0.f, ssq = 0.f;
      if ((EF & EF_POST) && p.res && tn < n_total_tiles) {
        const int mtn = fast_div(tn, p.fdNt);
        prefetch_res_l2(p, (EF & EF_GLU) != 0, mtn * TC_BM + row, (tn - mtn * p.n_tiles) * p.BN, sub, EPI_WARPS / 4);
      }
      mbar_wait(smem_u32(&tfull[buf]), (uint32_t)(i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + (uint32_t)(buf * p.BN) + ((uint32_t)(q * 32) << 16);
      epilogue_tile<EF>(p, er, tacc, n0, 32 * sub, 8 * EPI_WARPS, sv, ssum, ssq);
      // this warp is done reading the accumulator: release it to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[buf])) : "memory");
      if (tn < n_total_tiles) stash_vecs(svec + (buf ^ 1) * 4 * TC_VEC);
      if (EF & EF_STATS) {
        if (p.stat_mode == STAT_PER_G1_M) {
          if (er.valid) {
            double* st = p.stats + 2 * ((long)er.b * p.statR + er.m);
            atomicAdd(st, (double)ssum); atomicAdd(st + 1, (double)ssq);
          }
        } else {
          const int key = er.valid ? er.b : -1;
          const int key0 = __reduce_max_sync(0xffffffffu, key);
          const bool uniform = __all_sync(0xffffffffu, key == key0 || key == -1);
          if (key0 >= 0) {
            const int slot = (t + sub) % STAT_SLOTS;
            if (uniform) {
              // rows are combined in fp64 so that the result does not depend on which rows share a warp / tile,
              // i.e. on the position of a segment inside the batch (multi-GPU spans must reproduce one-GPU bits).
              // (Carrying per-thread fp64 sums across the tiles of a segment and reducing only when the segment changes was
              // measured: -13 us on the largest transposed-conv launch, +4 us on each of the ten linear2 launches -- not kept.)
              double a = er.valid ? (double)ssum : 0.0, c = er.valid ? (double)ssq : 0.0;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
              if (lane == 0) { double* sp = p.stats + 2 * ((long)key0 * STAT_SLOTS + slot); atomicAdd(sp, a); atomicAdd(sp + 1, c); }
            } else if (er.valid) {
              double* sp = p.stats + 2 * ((long)er.b * STAT_SLOTS + slot); atomicAdd(sp, (double)ssum); atomicAdd(sp + 1, (double)ssq);
            }
          }
        }
      }
    }
    if ((EF & EF_STATS) && p.fin_mr) __threadfence();      // this thread's statistics atomics are visible before the CTA's ticket
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
  if ((EF & EF_STATS) && p.fin_mr) {
    // The CTA that draws the last ticket finalises the per-segment statistics (same slot order and butterfly as
    // finalize_gn_slots_kernel: identical bits): one launch and its ~6 us less per GroupNorm of the transformer / decoder.
    int* s_last = (int*)(tmem_slot + 1);
    if (threadIdx.x == 0) *s_last = atomicAdd(p.fin_counter, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (*s_last) {
      __threadfence();
      for (int i = warp; i < p.fin_n; i += (int)(blockDim.x >> 5)) {
        double st[2] = {0.0, 0.0};
        for (int sl = lane; sl < STAT_SLOTS; sl += 32) {
          st[0] += __ldcg(p.stats + 2 * ((long)i * STAT_SLOTS + sl)); st[1] += __ldcg(p.stats + 2 * ((long)i * STAT_SLOTS + sl) + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { st[0] += __shfl_xor_sync(0xffffffffu, st[0], o); st[1] += __shfl_xor_sync(0xffffffffu, st[1], o); }
        if (lane == 0) { float m, r; stats_to_mean_rstd(st, p.fin_count, 1e-5f, m, r); p.fin_mr[2 * i] = m; p.fin_mr[2 * i + 1] = r; }
      }
    }
  }
}

// ------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// Two CTAs of a cluster (two SMs) share one 256 x BN tile: each loads ITS 128 rows of A and HALF of the B tile (BN / 2
// weight rows), the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256) which reads both halves of B from both
// CTAs' shared memory and writes each CTA's 128 accumulator rows into its own TMEM.  Per CTA and k-block that is 16 + 16 KB
// of operands instead of 16 + 32 KB: the K = 512 transformer GEMMs are bound by L2 -> SM operand traffic.
// Barriers: full[s] lives in the leader (armed with the bytes of both CTAs; both CTAs' TMA loads complete_tx on it), the
// leader's MMA commits arrive on empty[s] / tfull[buf] of BOTH CTAs (multicast), both CTAs' epilogue warps arrive on the
// leader's tempty[buf].
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {   // arrive on the same barrier of cluster rank 0
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(local_bar));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int EF>
__global__ void __launch_bounds__(TcCfg<1>::THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET on the shared-memory array: a round trip through uintptr_t makes every later access a
  // generic LD / ST (long-scoreboard, queued behind the global loads) instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stageA = TC_BM * TC_BK * 2;
  const int stageB = (p.BN / 2) * TC_BK * 2;        // this CTA's half of the B tile
  const int nst = p.stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + nst * stageA;
  uint64_t* bars = (uint64_t*)(sB + nst * stageB);
  uint64_t* full = bars, *empty = bars + TC_MAX_STAGES, *tfull = bars + 2 * TC_MAX_STAGES, *tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  float* svec = (float*)(bars + 2 * TC_MAX_STAGES + 8);
  constexpr int EPI_WARPS = TcCfg<1>::EPI_WARPS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int nkb = p.ntaps * p.kb_per_tap;
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int n_pair_tiles = m_pairs * p.n_tiles;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * p.BN)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < nst; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull[i]), 1); mbar_init(smem_u32(&tempty[i]), 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();        // both CTAs' barriers exist before any remote arrive / complete_tx
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      const bool issuer = elect_one();            // converged warp, one lane issues; ring position as running counters
      const uint32_t bytes = (uint32_t)(2 * (stageA + stageB));      // both CTAs' shares land on the leader's barrier
      int s = 0; uint32_t ph = 0;
      for (int u = cid; u < n_pair_tiles; u += ncl) {
        const int mp_ = fast_div(u, p.fdNt);
        const int row0 = (2 * mp_ + (int)rank) * TC_BM;
        const int nb0 = (u - mp_ * p.n_tiles) * p.BN + (int)rank * (p.BN / 2);
        int tap = 0, kin = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full[s]);
          if (issuer) {
            if (leader) mbar_expect_tx(fb, bytes);
            const uint32_t fbl = fb & 0xFEFFFFFFu;                     // the leader CTA's copy of this barrier
            tma_load_2d_pair(smem_u32(sA + s * stageA), &tmA, fbl, kin, row0 + p.tapRow[tap]);
            tma_load_2d_pair(smem_u32(sB + s * stageB), &tmB, fbl, tap * p.Ktap + kin, nb0);
          }
          __syncwarp();
          kin += TC_BK;
          if (kin >= p.Ktap) { kin = 0; ++tap; }
          if (++s == nst) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool issuer = elect_one();
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 across the pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int i = 0, s = 0; uint32_t ph = 0;
      const uint64_t da0 = make_sw128_desc(smem_u32(sA)), db0 = make_sw128_desc(smem_u32(sB));
      const uint64_t dstepA = (uint64_t)(stageA >> 4), dstepB = (uint64_t)(stageB >> 4);
      for (int u = cid; u < n_pair_tiles; u += ncl, ++i) {
        const int buf = i & 1;
        mbar_wait(smem_u32(&tempty[buf]), ((uint32_t)(i >> 1) & 1u) ^ 1u);     // both CTAs' epilogues drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(buf * p.BN);
        int kin = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&full[s]), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = da0 + (uint64_t)s * dstepA, db = db0 + (uint64_t)s * dstepB;
          const int nmma = (min(TC_BK, p.Ktap - kin) + 15) >> 4;
          if (issuer) {
            for (int k = 0; k < nmma; ++k)
              umma_bf16_pair(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            umma_commit_pair(smem_u32(&empty[s]));
          }
          __syncwarp();
          kin += TC_BK;
          if (kin >= p.Ktap) kin = 0;
          if (++s == nst) { s = 0; ph ^= 1u; }
        }
        if (issuer) umma_commit_pair(smem_u32(&tfull[buf]));
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int etid = threadIdx.x - 64;
    const int Nout_ = (EF & EF_GLU) ? p.N / 2 : p.N;
    // per-tile column vectors staged one tile ahead (as in gemm_tc_kernel)
    float vr0 = 0.f, vr1 = 0.f, vr2 = 0.f, vr3 = 0.f;
    auto fetch_vecs = [&](int n0) {
      if (etid < p.BN) {
        const int n = min(n0 + etid, p.N - 1);
        vr0 = p.bias ? __ldg(p.bias + n) : 0.f;
        if (EF & EF_GN) { vr1 = __ldg(p.gn_w + n); vr2 = __ldg(p.gn_b + n); }
        if ((EF & EF_POST) && p.colscale) {
          const int no = (EF & EF_GLU) ? (n0 >> 1) + etid : n0 + etid;
          vr3 = __ldg(p.colscale + min(no, Nout_ - 1));
        }
      }
    };
    auto stash_vecs = [&](float* sv) {
      if (etid < p.BN) {
        sv[etid] = vr0;
        if (EF & EF_GN) { sv[TC_VEC + etid] = vr1; sv[2 * TC_VEC + etid] = vr2; }
        if ((EF & EF_POST) && p.colscale) sv[3 * TC_VEC + etid] = vr3;
      }
    };
    if (cid < n_pair_tiles) { fetch_vecs((cid - fast_div(cid, p.fdNt) * p.n_tiles) * p.BN); stash_vecs(svec); }
    int i = 0;
    for (int u = cid; u < n_pair_tiles; u += ncl, ++i) {
      const int buf = i & 1;
      const int mp_ = fast_div(u, p.fdNt);
      const int row0 = (2 * mp_ + (int)rank) * TC_BM;
      const int n0 = (u - mp_ * p.n_tiles) * p.BN;
      const float* sv = svec + buf * 4 * TC_VEC;
      EpiRow er;
      const int rho = row0 + row;
      const int q2 = fast_div(rho, p.fdRpA);
      const int fp = rho - q2 * p.RpA;
      er.b = fast_div(q2, p.fdG2p);
      const int tp = q2 - er.b * p.G2p;
      er.valid = rho < p.Mflat && fp >= p.vlo && fp < p.vhi && tp >= p.gpf && tp < p.gpf + p.G2;
      er.orow = ((long)er.b * p.oG2p + tp + p.ogsh) * p.oRp + fp + p.orsh;
      er.m = fp - p.vlo;
      er.gmean = 0.f; er.grstd = 1.f;
      if ((EF & EF_GN) && er.valid) {
        const long gi = p.gn_mode == STAT_PER_G1_M ? (long)er.b * p.statR + er.m : (long)er.b;
        er.gmean = p.gn_mr[2 * gi]; er.grstd = p.gn_mr[2 * gi + 1];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
      const int un = u + ncl;
      if (un < n_pair_tiles) fetch_vecs((un - fast_div(un, p.fdNt) * p.n_tiles) * p.BN);
      er.edge_lo = er.m == 0; er.edge_hi = er.m == p.vhi - p.vlo - 1;
      float ssum = 0.f, ssq = 0.f;
      if ((EF & EF_POST) && p.res && un < n_pair_tiles) {
        const int mpn = fast_div(un, p.fdNt);
        prefetch_res_l2(p, (EF & EF_GLU) != 0, (2 * mpn + (int)rank) * TC_BM + row, (un - mpn * p.n_tiles) * p.BN, sub, EPI_WARPS / 4);
      }
      mbar_wait(smem_u32(&tfull[buf]), (uint32_t)(i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + (uint32_t)(buf * p.BN) + ((uint32_t)(q * 32) << 16);
      epilogue_tile<EF>(p, er, tacc, n0, 32 * sub, 8 * EPI_WARPS, sv, ssum, ssq);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[buf])) : "memory");
        else mbar_arrive_leader(smem_u32(&tempty[buf]));
      }
      if (un < n_pair_tiles) stash_vecs(svec + (buf ^ 1) * 4 * TC_VEC);
      if (EF & EF_STATS) {      // per-segment sums only (the pair kernel is used for the transformer linears)
        const int key = er.valid ? er.b : -1;
        const int key0 = __reduce_max_sync(0xffffffffu, key);
        const bool uniform = __all_sync(0xffffffffu, key == key0 || key == -1);
        if (key0 >= 0) {
          const int slot = (2 * u + (int)rank + sub) % STAT_SLOTS;
          if (uniform) {
            double a = er.valid ? (double)ssum : 0.0, c = er.valid ? (double)ssq : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
            if (lane == 0) { double* sp = p.stats + 2 * ((long)key0 * STAT_SLOTS + slot); atomicAdd(sp, a); atomicAdd(sp + 1, c); }
          } else if (er.valid) {
            double* sp = p.stats + 2 * ((long)er.b * STAT_SLOTS + slot); atomicAdd(sp, (double)ssum); atomicAdd(sp + 1, (double)ssq);
          }
        }
      }
    }
  }
  // the peer's shared memory and TMEM are read / written by the leader's MMAs: nobody leaves before everything has drained
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

bool make_tensor_map_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                        uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tensor_map_api_available() { return get_encode() != nullptr; }

static int g_tc_bn_cap = 256;
static bool g_tc_two_ctas = true;
static int g_tc_pair = -1;      // CTA-pair (cta_group::2) kernel for 256-wide tiles: tuning flag 0x40000 = every eligible launch,
                                // 0x80000 = only K >= 1536, neither = never (default).  Isolated it reaches 1427 vs 1290 TFLOP/s at K = 2048
                                // and 959 vs 1026 at K = 512; inside the forward (residual + statistics epilogue, cold operands) the
                                // K = 2048 launches measure the same 115-117 us either way (profiles/r01_summary.md).
static bool g_tc_w12 = true;             // 192-wide tiles on the 12-epilogue-warp variant (tuning flag 0x100000 turns it off)
static bool g_tc_fuse_fin = true;        // statistics finalised by the GEMM's last CTA (tuning flag 0x200000: separate finalize launch)
static bool g_tc_bres = true;            // weight-stationary B tiles for short-K layers (tuning flag 0x400000 turns it off)
static bool g_tc_halve_mid = false;      // tuning: N in (128, 256] as two N/2-wide tiles (two CTAs per SM) instead of one N-wide tile
void tc_set_bn_cap(int cap) {
  g_tc_bn_cap = cap & 0xffff; g_tc_two_ctas = !(cap & 0x10000); g_tc_halve_mid = (cap & 0x20000) != 0; g_tc_w12 = !(cap & 0x100000); g_tc_fuse_fin = !(cap & 0x200000); g_tc_bres = !(cap & 0x400000);
  g_tc_pair = (cap & 0x40000) ? 1 : (cap & 0x80000) ? 2 : 0;
}

int tc_pick_bn(int N) {
  if (N % 16) return 0;
  if (g_tc_halve_mid && N > 128 && N <= 256 && (N / 2) % 16 == 0) return N / 2;
  if (N <= g_tc_bn_cap) return N;
  if (g_tc_bn_cap >= 256 && N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;      // (measured: two 128-wide CTAs per SM beat one 192-wide tile for N = 384)
  if (N % 192 == 0 && g_tc_bn_cap >= 192) return 192;
  if (N % 64 == 0) return 64;
  if (N % 32 == 0) return 32;
  return 16;
}

bool tc_flat_supported(const TcFlat& f) {
  if (f.Ktap < 8 || f.Ktap % 8) return false;
  if (tc_pick_bn(f.N) == 0) return false;
  if (f.glu && (f.N % 16)) return false;
  if ((f.a_pitch * 2) % 16 || ((uintptr_t)f.A % 16) || ((uintptr_t)f.B % 16)) return false;
  if (f.ldc % 8 || f.c_is_f32 || f.alpha != 1.0f) return false;
  return get_encode() != nullptr;
}

static int num_sms() { return device_sm_count(); }

// (the CTA-pair kernel does not finalise; it is only reachable through the tuning flags)
bool tc_flat_fuses_finalize(const TcFlat& f) {
  return g_tc_fuse_fin && f.fin_mr != nullptr && f.fin_counter != nullptr && f.stats != nullptr && f.stat_mode == STAT_PER_G1 && g_tc_pair <= 0;
}

int launch_gemm_tc_flat(const TcFlat& f, cudaStream_t st) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.Mflat = (int)f.Mflat; p.N = f.N; p.BN = tc_pick_bn(f.N);
  p.ntaps = f.ntaps; p.Ktap = f.Ktap; p.kb_per_tap = (f.Ktap + TC_BK - 1) / TC_BK;
  for (int i = 0; i < 3; ++i) p.tapRow[i] = f.tapRow[i];
  p.RpA = f.RpA; p.G2p = f.G2p; p.gpf = f.gpf; p.G2 = f.G2; p.vlo = f.vlo; p.vhi = f.vhi;
  p.oG2p = f.oG2p; p.ogsh = f.ogsh; p.oRp = f.oRp; p.orsh = f.orsh; p.ldc = f.ldc;
  p.C = f.C; p.bias = f.bias;
  p.n_store = f.n_store > 0 ? f.n_store : (f.glu ? f.N / 2 : f.N);
  p.no_store = f.no_store;
  p.colscale = f.colscale; p.res = f.res; p.rowtab = f.rowtab; p.rowtab_scale = f.rowtab_scale;
  p.stats = f.stats; p.stat_mode = f.stat_mode; p.statR = f.statR; p.convt_cout = f.convt_cout;
  p.gn_mr = f.gn_mr; p.gn_w = f.gn_w; p.gn_b = f.gn_b; p.gn_mode = f.gn_mode;
  p.skip_lo = f.skip_lo; p.skip_hi = f.skip_hi;
  p.fin_mr = nullptr; p.fin_count = f.fin_count; p.fin_n = f.fin_n; p.fin_counter = f.fin_counter;
  p.wide = (f.ldc % 16 == 0) && ((uintptr_t)f.C % 32 == 0) && (!f.res || (uintptr_t)f.res % 32 == 0) && (f.skip_lo % 16 == 0) &&
           (f.skip_hi % 16 == 0);
  p.m_tiles = (int)((f.Mflat + TC_BM - 1) / TC_BM);
  p.n_tiles = (f.N + p.BN - 1) / p.BN;
  fast_div_init((uint32_t)p.RpA, p.fdRpA); fast_div_init((uint32_t)p.G2p, p.fdG2p); fast_div_init((uint32_t)p.n_tiles, p.fdNt);
  CUtensorMap tmA, tmB;
  // Two taps on ADJACENT rows of a dense activation buffer (row pitch == channels: the transposed-conv layers) are one
  // K = 2 * C view with overlapping rows [x[r], x[r+1]] (row stride C): no half-empty K blocks when C is not a multiple of 64
  // (C = 96: 3 full K blocks per tile instead of 4 half-padded ones).  The packed weights are already [tap0 | tap1] along K.
  bool merged = false;
  if (f.ntaps == 2 && f.tapRow[0] == 0 && f.tapRow[1] == 1 && f.a_pitch == f.Ktap && (f.Ktap % TC_BK) != 0 && f.a_rows > 1 &&
      make_tensor_map_2d(&tmA, f.A, (uint64_t)2 * f.Ktap, (uint64_t)f.a_rows - 1, (uint64_t)f.a_pitch * 2, TC_BK, TC_BM)) {
    merged = true;
    p.ntaps = 1; p.Ktap = 2 * f.Ktap; p.kb_per_tap = (p.Ktap + TC_BK - 1) / TC_BK; p.tapRow[1] = 0;
  }
  if (!merged &&
      !make_tensor_map_2d(&tmA, f.A, (uint64_t)f.Ktap, (uint64_t)f.a_rows, (uint64_t)f.a_pitch * 2, TC_BK, TC_BM)) return 2;
  // N = 384 (three 128-wide tiles on two CTAs per SM) cannot keep a whole-K weight tile next to the A ring in 112 KB; as two 192-wide
  // tiles on one CTA per SM (12 epilogue warps) it can when K <= 384: weight-stationary beats the narrower tiles there
  // (N = 384, K = 384 transposed conv of the frequency decoder: 734 -> 690 us)
  if (g_tc_bres && p.BN == 128 && f.N % 192 == 0 && f.N % 256 != 0) {
    const int nkb_all = p.ntaps * p.kb_per_tap;
    const int room = 227 * 1024 - 1024 - (512 + 2 * 4 * TC_VEC * 4) - nkb_all * 192 * TC_BK * 2;
    const long nt192 = f.N / 192, gx = ((long)num_sms() / nt192) * nt192;
    if (nkb_all >= 4 && room >= 3 * TC_BM * TC_BK * 2 && (long)p.m_tiles * nt192 >= 4 * gx) {      // (K = 192: measured 4 us slower)
      p.BN = 192; p.n_tiles = (int)nt192;
      fast_div_init((uint32_t)p.n_tiles, p.fdNt);
    }
  }
  if (!make_tensor_map_2d(&tmB, f.B, (uint64_t)f.ntaps * f.Ktap, (uint64_t)f.N, (uint64_t)f.ntaps * f.Ktap * 2, TC_BK, p.BN)) return 3;
  const int stage_bytes = TC_BM * TC_BK * 2 + p.BN * TC_BK * 2;
  // narrow tiles are epilogue-bound: two CTAs per SM (2 x 8 epilogue warps, 2 x 2 accumulators <= 512 TMEM columns)
  const int ctas_per_sm = (p.BN <= 128 && g_tc_two_ctas) ? 2 : 1;
  const int tail_bytes = 512 + 2 * 4 * TC_VEC * 4;          // barriers + staged column vectors
  const int smem_budget = (ctas_per_sm == 2 ? 112 : 227) * 1024 - 1024 - tail_bytes;
  p.stages = std::min(TC_MAX_STAGES, std::max(2, smem_budget / stage_bytes));
  size_t smem = 1024 + (size_t)p.stages * stage_bytes + tail_bytes;
  const long tiles = (long)p.m_tiles * p.n_tiles;
  long grid_x = std::min<long>(tiles, (long)num_sms() * ctas_per_sm);
  // weight-stationary variant: the whole-K B tile of one n-tile stays in shared memory when it leaves room for >= 3 A stages and
  // every CTA has several m-tiles to amortise it over (short-K layers: transposed convs, strided encoder convs, rewrites)
  {
    const int nkb_all = p.ntaps * p.kb_per_tap;
    const int stA = TC_BM * TC_BK * 2, stB = p.BN * TC_BK * 2;
    const long gx = ((long)num_sms() * ctas_per_sm / p.n_tiles) * p.n_tiles;
    const int a_stages = (smem_budget - nkb_all * stB) / stA;
    if (g_tc_bres && gx >= p.n_tiles && a_stages >= 3 && tiles >= 4 * gx) {
      p.b_res = 1;
      p.stages = std::min(TC_MAX_STAGES, a_stages);
      smem = 1024 + (size_t)p.stages * stA + (size_t)nkb_all * stB + tail_bytes;
      grid_x = gx;
    }
  }
  dim3 grid((unsigned)grid_x);
  if (g_tc_pair < 0) g_tc_pair = 0;
  // CTA pairs for the 256-wide tiles of long-M problems (transformer linears): 256 x 256 tile per cluster
  const bool pair = (g_tc_pair == 1 || (g_tc_pair == 2 && p.ntaps * p.kb_per_tap >= 24)) && p.BN == 256 && f.N % 256 == 0 && p.m_tiles >= 2 * num_sms() && f.stat_mode != STAT_PER_G1_M &&
                    (num_sms() % 2 == 0);
  if (!pair && tc_flat_fuses_finalize(f)) p.fin_mr = f.fin_mr;
  CUtensorMap tmBh;
  int pair_stages = 0;
  size_t pair_smem = 0;
  if (pair) {
    if (!make_tensor_map_2d(&tmBh, f.B, (uint64_t)f.ntaps * f.Ktap, (uint64_t)f.N, (uint64_t)f.ntaps * f.Ktap * 2, TC_BK, p.BN / 2)) return 3;
    const int sb = TC_BM * TC_BK * 2 + (p.BN / 2) * TC_BK * 2;
    pair_stages = std::min(TC_MAX_STAGES, (227 * 1024 - 1024 - tail_bytes) / sb);
    pair_smem = 1024 + (size_t)pair_stages * sb + tail_bytes;
  }
  int ef = 0;
  if (f.act == ACT_GELU) ef |= EF_GELU;
  if (f.glu) ef |= EF_GLU;
  if (f.gn_mr) ef |= EF_GN;
  if (f.colscale || f.rowtab || f.res) ef |= EF_POST;
  if (f.stat_mode != STAT_NONE) ef |= EF_STATS;
#define TC_LAUNCH(E)                                                                                                      \
  case E: {                                                                                                               \
    static PerDeviceOnce attr_set;                                                                                        \
    if (attr_set.first()) {                                                                                               \
      cudaFuncSetAttribute(gemm_tc_kernel<E, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                \
      cudaFuncSetAttribute(gemm_tc_kernel<E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);                \
      cudaFuncSetAttribute(gemm_tc_kernel<E, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                \
      cudaFuncSetAttribute(gemm_tc_pair_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);              \
    }                                                                                                                     \
    if (pair) {                                                                                                           \
      TcParams pp = p;                                                                                                    \
      pp.stages = pair_stages;                                                                                            \
      cudaLaunchConfig_t cfg;                                                                                             \
      memset(&cfg, 0, sizeof(cfg));                                                                                       \
      cfg.gridDim = dim3((unsigned)num_sms()); cfg.blockDim = dim3(TcCfg<1>::THREADS); cfg.dynamicSmemBytes = pair_smem;  \
      cfg.stream = st;                                                                                                    \
      cudaLaunchAttribute at[1];                                                                                          \
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1;             \
      at[0].val.clusterDim.z = 1;                                                                                         \
      cfg.attrs = at; cfg.numAttrs = 1;                                                                                   \
      return cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<E>, tmA, tmBh, pp) == cudaSuccess ? 0 : 6;                      \
    }                                                                                                                     \
    if (ctas_per_sm == 2) launch_pdl(gemm_tc_kernel<E, 2>, grid, dim3(TcCfg<2>::THREADS), smem, st, tmA, tmB, p);         \
    else if (p.BN == 192 && g_tc_w12) launch_pdl(gemm_tc_kernel<E, 3>, grid, dim3(TcCfg<3>::THREADS), smem, st, tmA, tmB, p);  \
    else launch_pdl(gemm_tc_kernel<E, 1>, grid, dim3(TcCfg<1>::THREADS), smem, st, tmA, tmB, p);                          \
    return 0;                                                                                                             \
  }
  switch (ef) {
    TC_LAUNCH(0)
    TC_LAUNCH(EF_GELU)
    TC_LAUNCH(EF_STATS)
    TC_LAUNCH(EF_POST)
    TC_LAUNCH(EF_POST | EF_STATS)
    TC_LAUNCH(EF_GLU)
    TC_LAUNCH(EF_GLU | EF_POST)
    TC_LAUNCH(EF_GLU | EF_GN | EF_POST)
    default: return 5;     // unsupported epilogue combination
  }
#undef TC_LAUNCH
}

}  // namespace athtd
