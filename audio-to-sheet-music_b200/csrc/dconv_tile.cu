// DConv residual branch (demucs demucs.py:DConv, depth 2, compress 8) for the layers whose
// GroupNorm(1, .) statistics span more than one CTA can hold: the whole TIME branch (statistics per segment over
// [C/8 | 2C, L_i]) and frequency levels 3-4 (per (segment, frequency row) over [., Tf], slabs too wide for one SM).
//
//   y += scale * GLU( GN2( W2 * GELU( GN1( conv_k3_dil(y) + b1 ) ) + b2 ) )
//
// runs as three bandwidth-bound passes; kernel boundaries are the two grid-wide reductions:
//   A  h = conv_k3_dil(y) + b1              -> narrow hidden buffer (C/8 channels, bf16) + GN1 partial sums
//   B  g = GELU(GN1(h)) in place            + GN2 partial sums of e = W2 g + b2 (e is never stored)
//   C  y += scale * GLU(GN2(W2 g + b2))     e recomputed from the narrow g (K = 16..48), y tile staged in shared memory
// The wide tensors move once in A (read y) and once in C (read + write y); the 2C-wide expand tensor never exists.
// All contractions are warp-level mma.sync on 16-row tiles (K = 3C -> C/8 -> 2C is far too small / too epilogue-heavy
// for a tcgen05 tile: the previous version ran them as three gemm_tc launches that were 5-10x off the HBM floor).
//
// Rows are addressed through a "logical row" space per segment: i = t * Rr + f, element offset
// base + b*seg + t*fs + f*C; the k3 taps shift i by +-dil*Rr (time branch: Rr = 1, a row shift; frequency branch:
// Rr = R rows per frame, a frame shift).  Out-of-range rows are zeros (the conv's zero padding).
#include "kernels.cuh"
#include "mma_sync.cuh"

namespace athtd {

struct DcGeom {
  int nT, Rr, rows;      // rows = nT * Rr logical rows per segment
  uint32_t fdRr[2];      // division-free i / Rr (common.cuh: fast_div)
  long seg, base, fs;
};

struct DcTileParams {
  DcGeom g;
  int dil, per_row;      // per_row: statistics per (segment, f) instead of per segment
  bf16* y; bf16* h;
  const bf16* w1; const float* b1; const float* g1w; const float* g1b;
  const bf16* w2; const float* b2; const float* g2w; const float* g2b; const float* scale;
  const bf16* ghi; const bf16* glo; const float* gv;      // pass B: G = W2^T W2 (hi + lo halves, [HP][HP]) and [2 W2^T b2 | W2^T 1 | sum b2 | sum b2^2]
  double* st1; double* st2;
  // optional HEncLayer tail fused into pass C of the second residual layer (C <= 96): out = GLU(rewrite_1x1(y) + rb)
  const bf16* rw; const float* rb; bf16* out;      // rw [2C][C], rb [2C]: GLU-interleaved rows like the expand
};

template <int C> struct DcDims {
  static constexpr int H = C / 8;
  static constexpr int HN = (H + 7) / 8;                          // hidden n-tiles that hold real channels
  static constexpr int HP = H <= 16 ? 16 : (H <= 32 ? 32 : 48);   // == hidden_pad(H), the packed K of the expand
  static constexpr int KS = HP / 16;
  static constexpr int XP = C + 8;                                // activation tile pitch (conflict-free ldmatrix)
  static constexpr int K1 = 3 * C, K1P = K1 + 8;
  static constexpr int K2P = HP + 8;
};

// 16-byte asynchronous global -> shared copies (LDGSTS): every thread issues all its copies back to back, so a tile load is
// bounded by bandwidth instead of by (copies per thread) x (global latency) as with register-staged loops.
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;                                     // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// logical rows [i_start, i_start + n) x CW channels starting at channel c0 of rows with C channels -> dst[n][CW + 8]
template <int C, int CW, bool ASYNC>
__device__ __forceinline__ void dc_load_rows(bf16* dst, const bf16* __restrict__ yb, const DcGeom& g, int i_start, int n, int c0 = 0) {
  constexpr int CV = CW / 8, XP = CW + 8;
  for (int idx = threadIdx.x; idx < n * CV; idx += blockDim.x) {
    const int row = idx / CV, cv = idx - row * CV;
    const int i = i_start + row;
    const bool valid = i >= 0 && i < g.rows;
    const int t = valid ? fast_div(i, g.fdRr) : 0, f = valid ? i - t * g.Rr : 0;
    const bf16* src = yb + (long)t * g.fs + (long)f * C + c0 + cv * 8;
    if (!ASYNC) {       // few copies per thread (C = 48 conv pass): register-staged loads beat LDGSTS + wait (measured)
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (valid) v = *(const uint4*)src;
      *(uint4*)(dst + row * XP + cv * 8) = v;
    } else {
      cp_async16(dst + row * XP + cv * 8, src, valid);
    }
  }
}

// warp 0: (sum, sumsq) fp64 slots -> mean / rstd (biased variance, eps 1e-5)
__device__ __forceinline__ void dc_mean_rstd(const double* st, int nslot, double count, float* out) {
  const int lane = threadIdx.x & 31;
  double a = 0.0, c = 0.0;
  for (int s = lane; s < nslot; s += 32) { a += st[2 * s]; c += st[2 * s + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if (lane == 0) {
    double tmp[2] = {a, c};
    stats_to_mean_rstd(tmp, count, 1e-5f, out[0], out[1]);
  }
}
// fills mr[2*f] for f < Rr (per_row) or mr[0..1] (per segment)
__device__ __forceinline__ void dc_stage_mean_rstd(const double* st, int b, int per_row, int Rr, double count, float* mr) {
  if (per_row) {
    for (int f = threadIdx.x; f < Rr; f += blockDim.x) {
      float m, r;
      stats_to_mean_rstd(st + 2 * ((long)b * Rr + f), count, 1e-5f, m, r);
      mr[2 * f] = m; mr[2 * f + 1] = r;
    }
  } else if (threadIdx.x < 32) {
    dc_mean_rstd(st + 2 * (long)b * STAT_SLOTS, STAT_SLOTS, count, mr);
  }
}

__device__ __forceinline__ void dc_block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
  __syncthreads();
  float x = l < nw ? red[2 * l] : 0.f, y = l < nw ? red[2 * l + 1] : 0.f;
  a = warp_sum(x); b = warp_sum(y);
}

// row-pair partial sums -> statistics accumulators
__device__ __forceinline__ void dc_row_stats(double* st, int b, int Rr, int r_lo, int r_hi, bool v_lo, bool v_hi, float s_lo,
                                             float q_lo, float s_hi, float q_hi) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, o); q_lo += __shfl_xor_sync(0xffffffffu, q_lo, o);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, o); q_hi += __shfl_xor_sync(0xffffffffu, q_hi, o);
  }
  if ((threadIdx.x & 3) == 0) {
    if (v_lo) { double* sp = st + 2 * ((long)b * Rr + r_lo % Rr); atomicAdd(sp, (double)s_lo); atomicAdd(sp + 1, (double)q_lo); }
    if (v_hi) { double* sp = st + 2 * ((long)b * Rr + r_hi % Rr); atomicAdd(sp, (double)s_hi); atomicAdd(sp + 1, (double)q_hi); }
  }
}

// ------------------------------------------------------------------ pass A
template <int C, int TM>
__global__ void __launch_bounds__(256) dconv_a_kernel(const DcTileParams p) {
  pdl_begin();
  typedef DcDims<C> D;
  constexpr int MW = TM / 16, NG = 8 / MW, NPW = (D::HN + NG - 1) / NG;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int halo = p.dil * p.g.Rr;
  const int nload = TM + 2 * halo;
  bf16* xs = (bf16*)smem_raw;                       // [nload][XP]: logical rows i0-halo .. i0+TM+halo
  bf16* w1s = xs + (size_t)nload * D::XP;           // [8*HN][K1P]
  float* red = (float*)(w1s + 8 * D::HN * D::K1P);  // 32 floats
  const int b = blockIdx.y, i0 = blockIdx.x * TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  dc_load_rows<C, C, (C >= 96)>(xs, p.y + p.g.base + (long)b * p.g.seg, p.g, i0 - halo, nload);
  for (int idx = tid; idx < 8 * D::HN * (D::K1 / 8); idx += blockDim.x) {
    const int n = idx / (D::K1 / 8), kc = idx - n * (D::K1 / 8);
    cp_async16(w1s + n * D::K1P + kc * 8, p.w1 + (long)n * D::K1 + kc * 8, true);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int mt = warp % MW, ng = warp / MW;
  float acc[NPW][4];
#pragma unroll
  for (int j = 0; j < NPW; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 3; ++tap) {
    const bf16* xr = xs + (size_t)tap * halo * D::XP;
#pragma unroll 4
    for (int kc = 0; kc < C / 16; ++kc) {
      uint32_t a[4];
      ldsm_a(xr, D::XP, mt * 16, kc * 16, lane, a);
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        const int nt = ng * NPW + j;
        if (nt < D::HN) {
          uint32_t bb[2];
          frag_b(w1s, D::K1P, nt * 8, tap * C + kc * 16, lane, bb);
          mma16816(acc[j], a, bb);
        }
      }
    }
  }
  const int r_lo = i0 + mt * 16 + g, r_hi = r_lo + 8;
  const bool v_lo = r_lo < p.g.rows, v_hi = r_hi < p.g.rows;
  bf16* hb = p.h + (long)b * p.g.rows * D::HP;
  float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
  for (int j = 0; j < NPW; ++j) {
    const int nt = ng * NPW + j;
    if (nt < D::HN) {
      const int c = nt * 8 + 2 * q;
      const float b0 = p.b1[c], b1 = p.b1[c + 1];      // zero beyond H, like the packed weight rows: pad channels are exactly 0
      const float x0 = acc[j][0] + b0, x1 = acc[j][1] + b1, x2 = acc[j][2] + b0, x3 = acc[j][3] + b1;
      if (v_lo) { s_lo += x0 + x1; q_lo += x0 * x0 + x1 * x1; *(uint32_t*)(hb + (long)r_lo * D::HP + c) = pack_bf16x2(x0, x1); }
      if (v_hi) { s_hi += x2 + x3; q_hi += x2 * x2 + x3 * x3; *(uint32_t*)(hb + (long)r_hi * D::HP + c) = pack_bf16x2(x2, x3); }
    }
  }
  if (ng == 0) {
    for (int c = 8 * D::HN + 2 * q; c < D::HP; c += 8) {   // K padding of the expand must be zero
      if (v_lo) *(uint32_t*)(hb + (long)r_lo * D::HP + c) = 0u;
      if (v_hi) *(uint32_t*)(hb + (long)r_hi * D::HP + c) = 0u;
    }
  }
  if (p.per_row) {
    dc_row_stats(p.st1, b, p.g.Rr, r_lo, r_hi, v_lo, v_hi, s_lo, q_lo, s_hi, q_hi);
  } else {
    float s = s_lo + s_hi, qq = q_lo + q_hi;
    dc_block_sum2(s, qq, red);
    if (tid == 0) {
      double* sp = p.st1 + 2 * ((long)b * STAT_SLOTS + blockIdx.x % STAT_SLOTS);
      atomicAdd(sp, (double)s); atomicAdd(sp + 1, (double)qq);
    }
  }
}

// GELU(GN1(.)) of one packed bf16 pair of the hidden buffer, written back in place; returns the packed result
template <int C>
__device__ __forceinline__ uint32_t dc_gn_gelu_pair(bf16* hp, int c, float mean, float rstd, const float* g1w, const float* g1b) {
  const float2 x = unpack_bf16x2(*(const uint32_t*)hp);
  // zero-padded affine beyond H: (x - m) * r * 0 + 0 = 0 and GELU(0) = 0, the K padding stays exactly zero
  const float2 y = gelu_fast2(f2fma(f2mul(f2add(x, f2splat(-mean)), f2splat(rstd)), make_float2(g1w[c], g1w[c + 1]),
                                     make_float2(g1b[c], g1b[c + 1])));
  const uint32_t r = pack_bf16x2(y.x, y.y);
  *(uint32_t*)hp = r;
  return r;
}

// ------------------------------------------------------------------ pass B
// g = GELU(GN1(h)) in place + the GroupNorm-2 partial sums of e = W2 g + b2 WITHOUT forming e: per row
//   sum_n e_n^2 = g^T G g + 2 (W2^T b2)^T g + sum b2^2 ,   sum_n e_n = (W2^T 1)^T g + sum b2 ,   G = W2^T W2  (packed once per model).
// Y = g G is HN x KS x 2 MMAs per 16-row tile (G as hi + lo activation-dtype halves: 2^-17 relative) whose accumulators start at the
// linear term and pair element by element with the A fragments of g; the 2C-wide product took 2C/8 x KS MMAs plus a packed
// square-accumulate per output pair (12 x more instructions at C = 48, 16 x at C = 384) and 86 KB of staged weights at C = 384.
template <int C>
__global__ void __launch_bounds__(256) dconv_b_kernel(const DcTileParams p) {
  pdl_begin();
  typedef DcDims<C> D;
  constexpr int GP = D::HP + 8;                     // pitch of the staged G halves (conflict-free fragment loads)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* ghs = (bf16*)smem_raw;                      // [HP][GP]  G hi (row n = output column j, k = i; G is symmetric)
  bf16* gls = ghs + D::HP * GP;                     // [HP][GP]  G lo
  float* gvs = (float*)(gls + D::HP * GP);          // [2 HP + 2]
  float* g1ws = gvs + 2 * D::HP + 2;                // [HP]
  float* g1bs = g1ws + D::HP;                       // [HP]
  float* mr = g1bs + D::HP;                         // [64]
  float* red = mr + 64;                             // [32]
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  for (int idx = tid; idx < 2 * D::HP * (D::HP / 8); idx += blockDim.x) {
    const int half = idx / (D::HP * (D::HP / 8)), r = idx - half * (D::HP * (D::HP / 8));
    const int n = r / (D::HP / 8), kc = r - n * (D::HP / 8);
    cp_async16((half ? gls : ghs) + n * GP + kc * 8, (half ? p.glo : p.ghi) + (long)n * D::HP + kc * 8, true);
  }
  cp_async_commit();
  for (int i = tid; i < 2 * D::HP + 2; i += blockDim.x) gvs[i] = p.gv[i];
  for (int i = tid; i < D::HP; i += blockDim.x) { g1ws[i] = p.g1w[i]; g1bs[i] = p.g1b[i]; }
  dc_stage_mean_rstd(p.st1, b, p.per_row, p.g.Rr, (double)D::H * (p.per_row ? p.g.nT : p.g.rows), mr);
  cp_async_wait<0>();
  __syncthreads();
  const float sb = gvs[2 * D::HP], sbb = gvs[2 * D::HP + 1];
  bf16* hb = p.h + (long)b * p.g.rows * D::HP;
  // per-thread sums over the tiles of this CTA in fp64: which tiles a CTA walks depends on the grid, i.e. on the batch size, and the
  // result must not (per-row partials are deterministic; fp64 adds of fp32 values are order-independent in practice)
  double ts = 0.0, tq = 0.0;
  for (int tile = blockIdx.x; tile * 128 < p.g.rows; tile += gridDim.x) {
    const int r_lo = tile * 128 + warp * 16 + g, r_hi = r_lo + 8;
    const bool v_lo = r_lo < p.g.rows, v_hi = r_hi < p.g.rows;
    const int f_lo = p.per_row ? r_lo - fast_div(r_lo, p.g.fdRr) * p.g.Rr : 0, f_hi = p.per_row ? r_hi - fast_div(r_hi, p.g.fdRr) * p.g.Rr : 0;
    const float m_lo = mr[2 * f_lo], rs_lo = mr[2 * f_lo + 1], m_hi = mr[2 * f_hi], rs_hi = mr[2 * f_hi + 1];
    uint32_t a[D::KS][4];
#pragma unroll
    for (int ks = 0; ks < D::KS; ++ks) {
      const int c0 = ks * 16 + 2 * q, c1 = c0 + 8;
      a[ks][0] = v_lo ? dc_gn_gelu_pair<C>(hb + (long)r_lo * D::HP + c0, c0, m_lo, rs_lo, g1ws, g1bs) : 0u;
      a[ks][1] = v_hi ? dc_gn_gelu_pair<C>(hb + (long)r_hi * D::HP + c0, c0, m_hi, rs_hi, g1ws, g1bs) : 0u;
      if (ks * 16 + 8 >= D::H) { a[ks][2] = 0u; a[ks][3] = 0u; continue; }     // pure K padding (zeros written by pass A)
      a[ks][2] = v_lo ? dc_gn_gelu_pair<C>(hb + (long)r_lo * D::HP + c1, c1, m_lo, rs_lo, g1ws, g1bs) : 0u;
      a[ks][3] = v_hi ? dc_gn_gelu_pair<C>(hb + (long)r_hi * D::HP + c1, c1, m_hi, rs_hi, g1ws, g1bs) : 0u;
    }
    // rows g (lo) and g + 8 (hi): this lane's share (columns 2q, 2q+1 of every hidden n-tile); lane q == 0 carries the constants
    float s_lo = q == 0 ? sb : 0.f, q_lo = q == 0 ? sbb : 0.f, s_hi = s_lo, q_hi = q_lo;
#pragma unroll
    for (int nt = 0; nt < D::HN; ++nt) {
      const int c = nt * 8 + 2 * q;
      const float2 v2 = *(const float2*)(gvs + c), u2 = *(const float2*)(gvs + D::HP + c);
      float d[4] = {v2.x, v2.y, v2.x, v2.y};
#pragma unroll
      for (int ks = 0; ks < D::KS; ++ks) {
        if (ks * 16 >= D::H) continue;
        uint32_t bb[2];
        frag_b(ghs, GP, nt * 8, ks * 16, lane, bb); mma16816(d, a[ks], bb);
        frag_b(gls, GP, nt * 8, ks * 16, lane, bb); mma16816(d, a[ks], bb);
      }
      const float2 gl = unpack_bf16x2(a[nt >> 1][(nt & 1) * 2]), gh = unpack_bf16x2(a[nt >> 1][(nt & 1) * 2 + 1]);
      q_lo = fmaf(d[0], gl.x, fmaf(d[1], gl.y, q_lo)); s_lo = fmaf(u2.x, gl.x, fmaf(u2.y, gl.y, s_lo));
      q_hi = fmaf(d[2], gh.x, fmaf(d[3], gh.y, q_hi)); s_hi = fmaf(u2.x, gh.x, fmaf(u2.y, gh.y, s_hi));
    }
    if (p.per_row) {
      dc_row_stats(p.st2, b, p.g.Rr, r_lo, r_hi, v_lo, v_hi, s_lo, q_lo, s_hi, q_hi);
    } else {
      if (v_lo) { ts += (double)s_lo; tq += (double)q_lo; }
      if (v_hi) { ts += (double)s_hi; tq += (double)q_hi; }
    }
  }
  if (!p.per_row) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ts += __shfl_xor_sync(0xffffffffu, ts, o); tq += __shfl_xor_sync(0xffffffffu, tq, o); }
    if (lane == 0) {
      double* sp = p.st2 + 2 * ((long)b * STAT_SLOTS + (blockIdx.x + warp) % STAT_SLOTS);
      atomicAdd(sp, ts); atomicAdd(sp + 1, tq);
    }
  }
}

// ------------------------------------------------------------------ pass C
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The GLU-interleaved packed expand rows (2j = value_j, 2j+1 = gate_j) are de-interleaved while they are staged
// (rows [0,C) values, [C,2C) gates): a lane then owns two ADJACENT value channels and their gates (one value n-tile + one
// gate n-tile per step) and updates the y tile with 4-byte accesses.  sigmoid(x) = 0.5 + 0.5 tanh(x/2) (one MUFU op).
// Per-segment statistics (PER_ROW = false) fold GroupNorm into one per-column FMA: e = d * alpha + beta.
// blockIdx.z selects a slice of CS value channels (and their gates): the wide layers run as several small CTAs per row tile
// (occupancy instead of one 150 KB CTA per SM).  The y tile arrives through cp.async while the expand is evaluated; the
// residual update happens after it has landed.
template <int C, int CS, int TM, bool PER_ROW, bool REWRITE>
__global__ void __launch_bounds__(256, (!REWRITE && C <= 96) ? 5 : 0) dconv_c_kernel(const DcTileParams p) {
  pdl_begin();
  typedef DcDims<C> D;
  constexpr int MW = TM / 16, NG = 8 / MW, CV = CS / 8, XPS = CS + 8;
  constexpr int NVT = (CV + NG - 1) / NG;           // value n-tiles per warp
  static_assert(!REWRITE || CS == C, "the fused rewrite needs all channels of a row in one CTA");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* ys = (bf16*)smem_raw;                       // [TM][XPS]
  bf16* w2s = ys + TM * XPS;                        // [2CS][K2P] de-interleaved: rows [0,CS) values, [CS,2CS) gates
  bf16* wrs = w2s + 2 * CS * D::K2P;                // REWRITE: [2C][XPS] de-interleaved rewrite rows, then out tile [TM][XPS]
  bf16* os = wrs + (REWRITE ? 2 * C * XPS : 0);
  float* al = (float*)(os + (REWRITE ? TM * XPS : 0));   // [2CS] alpha (PER_ROW: GroupNorm weight)
  float* be = al + 2 * CS;                          // [2CS] beta  (PER_ROW: GroupNorm bias)
  float* b2s = be + 2 * CS;                         // [2CS] expand bias (PER_ROW only)
  float* scs = b2s + 2 * CS;                        // [CS]
  float* mr = scs + CS;                             // [64]
  const int b = blockIdx.y, i0 = blockIdx.x * TM, ch0 = blockIdx.z * CS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  bf16* yb = p.y + p.g.base + (long)b * p.g.seg;
  for (int idx = tid; idx < 2 * CS * (D::HP / 8); idx += blockDim.x) {
    const int nd = idx / (D::HP / 8), kc = idx - nd * (D::HP / 8);
    const int n = nd < CS ? 2 * (ch0 + nd) : 2 * (ch0 + nd - CS) + 1;     // packed rows are GLU-interleaved
    cp_async16(w2s + nd * D::K2P + kc * 8, p.w2 + (long)n * D::HP + kc * 8, true);
  }
  cp_async_commit();
  dc_load_rows<C, CS, true>(ys, yb, p.g, i0, TM, ch0);
  if (REWRITE) {
    for (int idx = tid; idx < 2 * C * (C / 8); idx += blockDim.x) {
      const int nd = idx / (C / 8), kc = idx - nd * (C / 8);
      const int n = nd < C ? 2 * nd : 2 * (nd - C) + 1;
      cp_async16(wrs + nd * XPS + kc * 8, p.rw + (long)n * C + kc * 8, true);
    }
  }
  cp_async_commit();
  for (int i = tid; i < CS; i += blockDim.x) scs[i] = p.scale[ch0 + i];
  dc_stage_mean_rstd(p.st2, b, PER_ROW ? 1 : 0, p.g.Rr, (double)(2 * C) * (PER_ROW ? p.g.nT : p.g.rows), mr);
  __syncthreads();
  for (int nd = tid; nd < 2 * CS; nd += blockDim.x) {
    const int n = nd < CS ? 2 * (ch0 + nd) : 2 * (ch0 + nd - CS) + 1;
    const float half = nd < CS ? 1.0f : 0.5f;       // gates carry the x/2 of the tanh form
    if (PER_ROW) {
      al[nd] = half * p.g2w[n]; be[nd] = half * p.g2b[n]; b2s[nd] = p.b2[n];
    } else {
      const float a = mr[1] * p.g2w[n];
      al[nd] = half * a; be[nd] = half * ((p.b2[n] - mr[0]) * a + p.g2b[n]);
    }
  }
  cp_async_wait<1>();                               // expand weights have landed (the y tile may still be in flight)
  __syncthreads();
  const int mt = warp % MW, ng = warp / MW;
  const int r_lo = i0 + mt * 16 + g, r_hi = r_lo + 8;
  const bool v_lo = r_lo < p.g.rows, v_hi = r_hi < p.g.rows;
  const int f_lo = PER_ROW ? r_lo - fast_div(r_lo, p.g.fdRr) * p.g.Rr : 0, f_hi = PER_ROW ? r_hi - fast_div(r_hi, p.g.fdRr) * p.g.Rr : 0;
  const float m_lo = mr[2 * f_lo], rs_lo = mr[2 * f_lo + 1], m_hi = mr[2 * f_hi], rs_hi = mr[2 * f_hi + 1];
  const bf16* hb = p.h + (long)b * p.g.rows * D::HP;
  uint32_t a[D::KS][4];
#pragma unroll
  for (int ks = 0; ks < D::KS; ++ks) {
    const int c0 = ks * 16 + 2 * q, c1 = c0 + 8;
    a[ks][0] = v_lo ? ld_b32(hb + (long)r_lo * D::HP + c0) : 0u;
    a[ks][1] = v_hi ? ld_b32(hb + (long)r_hi * D::HP + c0) : 0u;
    if (ks * 16 + 8 >= D::H) { a[ks][2] = 0u; a[ks][3] = 0u; continue; }
    a[ks][2] = v_lo ? ld_b32(hb + (long)r_lo * D::HP + c1) : 0u;
    a[ks][3] = v_hi ? ld_b32(hb + (long)r_hi * D::HP + c1) : 0u;
  }
  float uo[NVT][4];
#pragma unroll
  for (int iv = 0; iv < NVT; ++iv) {
    const int vt = ng + iv * NG;
    if (vt < CV) {
      float dv[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < D::KS; ++ks) {
        uint32_t bvf[2], bgf[2];      // value + gate n-tile fragments of this k step through one ldmatrix.x4
        ldsm_b_pair(w2s, D::K2P, vt * 8, CS + vt * 8, ks * 16, lane, bvf, bgf);
        mma16816(dv, a[ks], bvf);
        mma16816(dg, a[ks], bgf);
      }
      const int j = vt * 8 + 2 * q;
      const float2 av = *(const float2*)(al + j), bv = *(const float2*)(be + j);
      const float2 ag = *(const float2*)(al + CS + j), bg = *(const float2*)(be + CS + j);
      const float2 sc = *(const float2*)(scs + j);
      float ev[4], eg[4];
      if (PER_ROW) {
        const float2 kv = *(const float2*)(b2s + j), kg = *(const float2*)(b2s + CS + j);
        ev[0] = fmaf((dv[0] + kv.x - m_lo) * rs_lo, av.x, bv.x); ev[1] = fmaf((dv[1] + kv.y - m_lo) * rs_lo, av.y, bv.y);
        ev[2] = fmaf((dv[2] + kv.x - m_hi) * rs_hi, av.x, bv.x); ev[3] = fmaf((dv[3] + kv.y - m_hi) * rs_hi, av.y, bv.y);
        eg[0] = fmaf((dg[0] + kg.x - m_lo) * rs_lo, ag.x, bg.x); eg[1] = fmaf((dg[1] + kg.y - m_lo) * rs_lo, ag.y, bg.y);
        eg[2] = fmaf((dg[2] + kg.x - m_hi) * rs_hi, ag.x, bg.x); eg[3] = fmaf((dg[3] + kg.y - m_hi) * rs_hi, ag.y, bg.y);
      } else {
        const float2 v01 = f2fma(make_float2(dv[0], dv[1]), av, bv), v23 = f2fma(make_float2(dv[2], dv[3]), av, bv);
        const float2 g01 = f2fma(make_float2(dg[0], dg[1]), ag, bg), g23 = f2fma(make_float2(dg[2], dg[3]), ag, bg);
        ev[0] = v01.x; ev[1] = v01.y; ev[2] = v23.x; ev[3] = v23.y; eg[0] = g01.x; eg[1] = g01.y; eg[2] = g23.x; eg[3] = g23.y;
      }
      {
        const float2 h2 = f2splat(0.5f);
        const float2 s01 = f2fma(h2, make_float2(tanh_approx(eg[0]), tanh_approx(eg[1])), h2);
        const float2 s23 = f2fma(h2, make_float2(tanh_approx(eg[2]), tanh_approx(eg[3])), h2);
        const float2 u01 = f2mul(sc, f2mul(make_float2(ev[0], ev[1]), s01)), u23 = f2mul(sc, f2mul(make_float2(ev[2], ev[3]), s23));
        uo[iv][0] = u01.x; uo[iv][1] = u01.y; uo[iv][2] = u23.x; uo[iv][3] = u23.y;
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  bf16* y_lo = ys + (mt * 16 + g) * XPS, *y_hi = y_lo + 8 * XPS;
#pragma unroll
  for (int iv = 0; iv < NVT; ++iv) {
    const int vt = ng + iv * NG;
    if (vt < CV) {
      const int j = vt * 8 + 2 * q;
      const float2 x0 = unpack_bf16x2(*(const uint32_t*)(y_lo + j)), x1 = unpack_bf16x2(*(const uint32_t*)(y_hi + j));
      *(uint32_t*)(y_lo + j) = pack_bf16x2(x0.x + uo[iv][0], x0.y + uo[iv][1]);
      *(uint32_t*)(y_hi + j) = pack_bf16x2(x1.x + uo[iv][2], x1.y + uo[iv][3]);
    }
  }
  __syncthreads();
  if (REWRITE) {
    // HEncLayer tail on the updated tile: out = GLU(Wr y + rb); y itself is not needed again and is not written back
    uint32_t ar[C / 16][4];
#pragma unroll
    for (int kc = 0; kc < C / 16; ++kc) ldsm_a(ys, XPS, mt * 16, kc * 16, lane, ar[kc]);
#pragma unroll
    for (int iv = 0; iv < NVT; ++iv) {
      const int vt = ng + iv * NG;
      if (vt < CV) {
        float dv[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kc = 0; kc < C / 16; ++kc) {
          uint32_t bvf[2], bgf[2];
          ldsm_b_pair(wrs, XPS, vt * 8, C + vt * 8, kc * 16, lane, bvf, bgf);
          mma16816(dv, ar[kc], bvf);
          mma16816(dg, ar[kc], bgf);
        }
        const int j = vt * 8 + 2 * q;
        const float rv0 = p.rb[2 * j], rg0 = 0.5f * p.rb[2 * j + 1], rv1 = p.rb[2 * j + 2], rg1 = 0.5f * p.rb[2 * j + 3];
        const float o0 = (dv[0] + rv0) * fmaf(0.5f, tanh_approx(fmaf(dg[0], 0.5f, rg0)), 0.5f);
        const float o1 = (dv[1] + rv1) * fmaf(0.5f, tanh_approx(fmaf(dg[1], 0.5f, rg1)), 0.5f);
        const float o2 = (dv[2] + rv0) * fmaf(0.5f, tanh_approx(fmaf(dg[2], 0.5f, rg0)), 0.5f);
        const float o3 = (dv[3] + rv1) * fmaf(0.5f, tanh_approx(fmaf(dg[3], 0.5f, rg1)), 0.5f);
        *(uint32_t*)(os + (mt * 16 + g) * XPS + j) = pack_bf16x2(o0, o1);
        *(uint32_t*)(os + (mt * 16 + g + 8) * XPS + j) = pack_bf16x2(o2, o3);
      }
    }
    __syncthreads();
  }
  const bf16* src_tile = REWRITE ? os : ys;
  bf16* dstb = REWRITE ? p.out + p.g.base + (long)b * p.g.seg : yb;
  for (int idx = tid; idx < TM * CV; idx += blockDim.x) {
    const int row = idx / CV, cv = idx - row * CV;
    const int i = i0 + row;
    if (i < p.g.rows) {
      const int t = fast_div(i, p.g.fdRr), f = i - t * p.g.Rr;
      *(uint4*)(dstb + (long)t * p.g.fs + (long)f * C + ch0 + cv * 8) = *(const uint4*)(src_tile + row * XPS + cv * 8);
    }
  }
}

// ------------------------------------------------------------------ host side
bool dconv_tile_supported(int C) { return C == 48 || C == 96 || C == 192 || C == 384; }

template <int C, int TM, int CS, int TMC>
static int dconv_tile_launch(const DcTileParams& p, int B, cudaStream_t st) {
  typedef DcDims<C> D;
  const int halo = p.dil * p.g.Rr;
  const size_t smA = (size_t)(TM + 2 * halo) * D::XP * 2 + (size_t)8 * D::HN * D::K1P * 2 + 32 * 4 + 16;
  const size_t smB = (size_t)2 * D::HP * (D::HP + 8) * 2 + (size_t)(2 * D::HP + 2 + 2 * D::HP + 64 + 32) * 4 + 16;
  const bool rewrite = p.rw != nullptr && CS == C;
  const size_t smC = (size_t)TMC * (CS + 8) * 2 + (size_t)2 * CS * D::K2P * 2 + (size_t)(7 * CS + 64) * 4 + 16 +
                     (rewrite ? (size_t)(2 * C + TMC) * (CS + 8) * 2 : 0);
  if (smA > 227 * 1024 || smC > 227 * 1024) return 1;
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(dconv_a_kernel<C, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(dconv_b_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(dconv_c_kernel<C, CS, TMC, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(dconv_c_kernel<C, CS, TMC, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if constexpr (CS == C)
      cudaFuncSetAttribute(dconv_c_kernel<C, CS, TMC, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int tiles = (p.g.rows + TM - 1) / TM;
  const int tiles_c = (p.g.rows + TMC - 1) / TMC;
  const int tiles_b = (p.g.rows + 127) / 128;
  const int gb = std::min(tiles_b, std::max(1, (device_sm_count() * 16 + B - 1) / B));
  launch_pdl(dconv_a_kernel<C, TM>, dim3(dim3(tiles, B)), dim3(256), smA, st, p);
  launch_pdl(dconv_b_kernel<C>, dim3(dim3(gb, B)), dim3(256), smB, st, p);
  if (p.per_row) launch_pdl(dconv_c_kernel<C, CS, TMC, true, false>, dim3(dim3(tiles_c, B, C / CS)), dim3(256), smC, st, p);
  else if (rewrite) {
    if constexpr (CS == C) launch_pdl(dconv_c_kernel<C, CS, TMC, false, true>, dim3(dim3(tiles_c, B, 1)), dim3(256), smC, st, p);
  } else launch_pdl(dconv_c_kernel<C, CS, TMC, false, false>, dim3(dim3(tiles_c, B, C / CS)), dim3(256), smC, st, p);
  return 0;
}

// ptrs: w1p [HP][3C] bf16, b1p, g1wp, g1bp [HP] fp32, w2p [2C][HP] bf16 (GLU-interleaved rows), b2i, g2wi, g2bi [2C] fp32
// (interleaved), scale [C] fp32.  rw / rb / out (optional, time branch with C <= 96): fuse out = GLU(rewrite(y)) into pass C.
// Three launches; returns 0 on success.
// (C = 96 also works but measured 16 us slower than the separate tcgen05 rewrite GEMM)
bool dconv_tile_can_rewrite(int C, bool freq) { return !freq && C == 48; }
int launch_dconv_tile(bf16* y, RowSpace ys, bf16* h, int dil, const bf16* w1p, const float* b1p, const float* g1wp, const float* g1bp,
                      const bf16* w2p, const float* b2i, const float* g2wi, const float* g2bi, const bf16* ghi, const bf16* glo, const float* gv,
                      const float* scale, double* st1,
                      double* st2, const bf16* rw, const float* rb, bf16* out, cudaStream_t st) {
  DcTileParams p;
  p.rw = rw; p.rb = rb; p.out = out;
  const bool freq = ys.G2 > 1;
  p.g.nT = freq ? ys.G2 : ys.R; p.g.Rr = freq ? ys.R : 1; p.g.rows = p.g.nT * p.g.Rr;
  fast_div_init((uint32_t)p.g.Rr, p.g.fdRr);
  p.g.seg = ys.g1_stride(); p.g.base = ys.origin(); p.g.fs = freq ? (long)ys.Rp * ys.C : ys.C;
  p.dil = dil; p.per_row = freq ? 1 : 0;
  p.y = y; p.h = h; p.w1 = w1p; p.b1 = b1p; p.g1w = g1wp; p.g1b = g1bp; p.w2 = w2p; p.b2 = b2i; p.g2w = g2wi; p.g2b = g2bi;
  p.scale = scale; p.st1 = st1; p.st2 = st2; p.ghi = ghi; p.glo = glo; p.gv = gv;
  const int B = ys.batch();
  if (freq && ys.R > 32) return 1;
  switch (ys.C) {
    case 48: return dconv_tile_launch<48, 128, 48, 128>(p, B, st);
    case 96: return dconv_tile_launch<96, 128, 96, 64>(p, B, st);
    case 192: return dconv_tile_launch<192, 128, 96, 64>(p, B, st);
    case 384: return dconv_tile_launch<384, 64, 96, 64>(p, B, st);
  }
  return 1;
}

}  // namespace athtd
