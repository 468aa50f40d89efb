// Fused frequency-encoder layer tail for (segment, frequency row) slabs, bf16 build:
//
//   [level 0 only] y = GELU(conv_k8s4(x))                      (K = 32: 8 frequency taps x 4 CaC channels)
//   y += scale * GLU(GN2(W2 GELU(GN1(conv_k3_dil(y)))))        x 2 (dilation 1, 2)     -- demucs DConv
//   out = GLU(rewrite_1x1(y)) (+ 2.0 * freq_emb[f] on level 0)                         -- demucs HEncLayer tail
//
// (demucs hdemucs.py:HEncLayer / demucs.py:DConv; ATHTDemucs_v2.py:210-215.)
// Inside a frequency-branch layer every GroupNorm statistic is local to one (segment, frequency row), so the whole
// chain runs on a [T x C] slab that lives in shared memory: the layer reads its input once and writes its output
// once (HBM-bound by design; the unfused sequence moved ~7 slab-sized tensors per DConv layer).
//
// PERSISTENT: one CTA per SM slot loops over slabs; every weight of the layer (conv taps, both DConv depths, rewrite) is
// copied ONCE per CTA from the pre-packed bf16 blob into padded shared-memory tiles (the first version re-staged and
// re-converted ~11k fp32 weights per slab, a third of its instructions).  GroupNorm + bias fold into one per-column FMA
// (alpha / beta recomputed per slab after the statistics), sigmoid is 0.5 + 0.5 tanh(x/2) (one MUFU op).
//
// The contractions are tiny (K = 32 / 3C / 16 / C on 16-row tiles of the resident slab), so they use the warp-level
// tensor-core MMA (mma.sync.m16n8k16 bf16 -> fp32) straight from shared memory: a tcgen05 / TMEM round trip per
// 259-row slab would cost more than the math.  The large, compute-bound GEMMs of the path stay on gemm_tc_kernel.
#include "kernels.cuh"
#include "mma_sync.cuh"

namespace athtd {

static constexpr int ER_THREADS = 288;     // 9 warps; a 6 s segment has 259 frames = 17 row tiles of 16

__device__ __forceinline__ float er_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// barrier of one 9-warp slab group (named barrier 1 + group; a CTA runs one or two groups that share the staged weights)
__device__ __forceinline__ void er_gsync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(ER_THREADS) : "memory"); }
__device__ __forceinline__ void er_block_sum2(float& a, float& b, float* red, int grp, int tid) {
  a = warp_sum(a); b = warp_sum(b);
  const int w = tid >> 5, l = tid & 31, nw = ER_THREADS >> 5;
  er_gsync(grp);
  if (l == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
  er_gsync(grp);
  float x = l < nw ? red[2 * l] : 0.f, y = l < nw ? red[2 * l + 1] : 0.f;
  a = warp_sum(x); b = warp_sum(y);
}
__device__ __forceinline__ void sts_pair(bf16* p, float a, float b) { *(uint32_t*)p = pack_bf16x2(a, b); }

template <int C> struct ErDims {
  static constexpr int H = C / 8;
  static constexpr int HN = (H + 7) / 8;          // hidden n-tiles (1 or 2)
  static constexpr int HP = 16;                   // packed hidden K (hidden_pad(H) for H <= 16)
  static constexpr int XP = C + 8;                // slab pitch: conflict-free fragment loads
  static constexpr int K1 = 3 * C, K1P = K1 + 8;  // conv-k3 weights [8*HN][K1P]
  static constexpr int K2P = HP + 8;              // expand weights  [2C][K2P]
  static constexpr int KRP = C + 8;               // rewrite weights [2C][KRP]
  static constexpr int AT = C / 8;                // value n-tiles (gates are tiles AT .. 2AT-1)
  static constexpr int PP = 40;                   // level-0 patch pitch (K = 32)
  static constexpr int VD = 48 + 7 * C;           // fp32 vectors per depth: b1 g1w g1b [16] | b2 g2w g2b [2C] | scale [C]
};

template <int C, bool FUSE_CONV, int G>
__global__ void __launch_bounds__(ER_THREADS * G, (C == 48 && G == 1 ? 2 : 1))
enc_row_kernel(const bf16* __restrict__ xin, RowSpace xis, const bf16* __restrict__ yin, bf16* __restrict__ out, RowSpace ys,
               const EncRowParams P) {
  pdl_trigger();
  typedef ErDims<C> D;
  const int Tn = ys.G2;
  const int MT = (Tn + 15) / 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // shared by all slab groups of the CTA: every weight of the layer
  bf16* w1t = (bf16*)smem_raw;                              // [2][8*HN][K1P]
  bf16* w2t = w1t + 2 * 8 * D::HN * D::K1P;                 // [2][2C][K2P]  rows [0,C) values, [C,2C) gates
  bf16* wrt = w2t + 2 * 2 * C * D::K2P;                     // [2C][KRP]     same split
  bf16* cwt = wrt + 2 * C * D::KRP;                         // FUSE_CONV: conv weights [C][PP]
  float* vec = (float*)(cwt + (FUSE_CONV ? C * D::PP : 0)); // [2][VD] | rb [2C] (gate half pre-scaled by 0.5) | cb [C]
  float* rbs = vec + 2 * D::VD;
  float* cbs = rbs + 2 * C;
  // GroupNorm-2 statistics without a pass over e = W2 h + b2: sum(e) and sum(e^2) over the slab are LINEAR in the second-moment
  // matrix D = h~^T h~ of the hidden tile with a constant-one column appended (h~ = [h, 1]), see the statistics step below
  float* gq = cbs + C;                                      // [2][16][16]  sum(e^2) = sum_ij gq[i][j] D[i][j]
  float* gs = gq + 2 * 256;                                 // [2][16]      sum(e)   = sum_j gs[j] D[H][j]
  // per slab group (9 warps): activation tiles and per-slab vectors
  constexpr int GF = 4 * C + C + 64;                        // al [2C] be [2C] embv [C] red [64]
  // FUSE_CONV: the hidden operand tile hb aliases the (larger) level-0 patch tile, which is dead after the strided conv
  const int slab_elems = (MT * 16 + 4) * D::XP + (FUSE_CONV ? MT * 16 * D::PP : MT * 16 * D::K2P);
  const int btid = threadIdx.x;
  const int grp = btid / ER_THREADS;
  const int tid = btid - grp * ER_THREADS, lane = tid & 31, warp = tid >> 5, nwarp = ER_THREADS >> 5;
  float* gvec = gs + 32 + grp * GF;
  float* al = gvec;                                         // [2C] per-slab GroupNorm-2 alpha (gate half x 0.5)
  float* be = al + 2 * C;                                   // [2C] beta
  float* embv = be + 2 * C;                                 // [C]
  float* red = embv + C;                                    // 64 floats
  bf16* xs = (bf16*)(gs + 32 + G * GF) + (size_t)grp * slab_elems;   // [(MT*16 + 4)][XP], rows shifted by +2 (zero halo)
  bf16* hb = xs + (MT * 16 + 4) * D::XP;                    // [MT*16][K2P]
  bf16* pst = hb;                                           // FUSE_CONV: patch [MT*16][PP] (aliases hb)
  const int g = lane >> 2, q = lane & 3;
  bf16* xsi = xs + 2 * D::XP;                               // interior row 0

  // ---- once per CTA: zero the activation tiles, stage every weight of the layer
  for (int i = tid; i < ((MT * 16 + 4) * D::XP) / 2; i += ER_THREADS) ((uint32_t*)xs)[i] = 0u;
  if (!FUSE_CONV)
    for (int i = tid; i < (MT * 16 * D::K2P) / 2; i += ER_THREADS) ((uint32_t*)hb)[i] = 0u;
  if (FUSE_CONV) {
    for (int i = btid; i < C * 4; i += blockDim.x) {         // conv taps [C][32] -> [C][PP]
      const int n = i >> 2, kc = i & 3;
      *(uint4*)(cwt + n * D::PP + kc * 8) = *(const uint4*)(P.cw + n * 32 + kc * 8);
    }
    for (int i = btid; i < C; i += blockDim.x) cbs[i] = P.cb[i];
  }
  for (int dd = 0; dd < 2; ++dd) {
    bf16* w1 = w1t + dd * 8 * D::HN * D::K1P;
    for (int i = btid; i < 8 * D::HN * (D::K1 / 8); i += blockDim.x) {
      const int n = i / (D::K1 / 8), kc = i - n * (D::K1 / 8);
      *(uint4*)(w1 + n * D::K1P + kc * 8) = *(const uint4*)(P.w1[dd] + (long)n * D::K1 + kc * 8);
    }
    bf16* w2 = w2t + dd * 2 * C * D::K2P;
    for (int i = btid; i < 2 * C * 2; i += blockDim.x) {      // packed rows are GLU-interleaved (2j value, 2j+1 gate)
      const int n = i >> 1, kc = i & 1, nd = (n & 1) * C + (n >> 1);
      *(uint4*)(w2 + nd * D::K2P + kc * 8) = *(const uint4*)(P.w2[dd] + (long)n * D::HP + kc * 8);
    }
    float* v = vec + dd * D::VD;
    for (int i = btid; i < 16; i += blockDim.x) { v[i] = P.b1[dd][i]; v[16 + i] = P.g1w[dd][i]; v[32 + i] = P.g1b[dd][i]; }
    for (int n = btid; n < 2 * C; n += blockDim.x) {
      const int nd = (n & 1) * C + (n >> 1);
      v[48 + nd] = P.b2[dd][n]; v[48 + 2 * C + nd] = P.g2w[dd][n]; v[48 + 4 * C + nd] = P.g2b[dd][n];
    }
    for (int i = btid; i < C; i += blockDim.x) v[48 + 6 * C + i] = P.scale[dd][i];
  }
  for (int i = btid; i < 2 * C * (C / 8); i += blockDim.x) {
    const int n = i / (C / 8), kc = i - n * (C / 8), nd = (n & 1) * C + (n >> 1);
    uint4 w = *(const uint4*)(P.rw + (long)n * C + kc * 8);
    if (n & 1) {      // gate rows x 0.5 (tanh form of the sigmoid; a power of two: exact in bf16)
      __nv_bfloat162* h = (__nv_bfloat162*)&w;
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(0.5f * __low2float(h[e]), 0.5f * __high2float(h[e]));
    }
    *(uint4*)(wrt + nd * D::KRP + kc * 8) = w;
  }
  for (int n = btid; n < 2 * C; n += blockDim.x) rbs[(n & 1) * C + (n >> 1)] = (n & 1) ? 0.5f * P.rb[n] : P.rb[n];
  __syncthreads();
  // With W = the staged (bf16) expand weights [2C][H], b = b2, c = H the constant column:
  //   sum_{t,n} e     = sum_j (sum_n W[n][j]) D[c][j] + (sum_n b[n]) D[c][c]
  //   sum_{t,n} e^2   = sum_{i,j<H} (W^T W)[i][j] D[i][j] + sum_j 2 (sum_n b[n] W[n][j]) D[c][j] + (sum_n b[n]^2) D[c][c]
  for (int idx = btid; idx < 2 * 272; idx += blockDim.x) {
    const int dd = idx / 272, e = idx - dd * 272;
    const bf16* w2 = w2t + dd * 2 * C * D::K2P;
    const float* b2v = vec + dd * D::VD + 48;
    float acc = 0.f;
    if (e < 256) {
      const int i = e >> 4, j = e & 15;
      if (i < D::H && j < D::H) {
        for (int n = 0; n < 2 * C; ++n) acc += __bfloat162float(w2[n * D::K2P + i]) * __bfloat162float(w2[n * D::K2P + j]);
      } else if (i == D::H && j < D::H) {
        for (int n = 0; n < 2 * C; ++n) acc += 2.0f * b2v[n] * __bfloat162float(w2[n * D::K2P + j]);
      } else if (i == D::H && j == D::H) {
        for (int n = 0; n < 2 * C; ++n) acc += b2v[n] * b2v[n];
      }
      gq[dd * 256 + e] = acc;
    } else {
      const int j = e - 256;
      if (j < D::H) { for (int n = 0; n < 2 * C; ++n) acc += __bfloat162float(w2[n * D::K2P + j]); }
      else if (j == D::H) { for (int n = 0; n < 2 * C; ++n) acc += b2v[n]; }
      gs[dd * 16 + j] = acc;
    }
  }
  __syncthreads();
  pdl_wait();        // the weight staging above (constant data) overlaps the predecessor; activations are read from here on

  const int n_slabs = ys.batch() * ys.R;
  for (int slab = blockIdx.x * G + grp; slab < n_slabs; slab += gridDim.x * G) {
    const int f = slab % ys.R, b = slab / ys.R;
    if (P.emb) for (int i = tid; i < C; i += ER_THREADS) embv[i] = P.emb_scale * P.emb[f * C + i];
    if (FUSE_CONV) {
      // level 0: y = GELU(conv_k8s4(x)): patch [T x 32] = 8 padded frequency rows x 4 channels, contiguous in xin
      if (P.zspec) {
        // straight from the fp32 spectrogram: chunk ch = patch rows 2ch, 2ch+1 (4 CaC channels each), normalised + rounded
        // exactly like the packed copy was (bf16((z - mean) * inv)); rows outside [0, zrows) are the conv's zero padding
        const float mean = P.ms_spec[2 * b], inv = 1.0f / (1e-5f + P.ms_spec[2 * b + 1]);
        for (int i = tid; i < MT * 16 * 4; i += ER_THREADS) {
          const int t = i >> 2, ch = i & 3;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (t < Tn) {
            const int r0 = 4 * f - 2 + 2 * ch;
            const float4* zp = (const float4*)P.zspec + ((long)(b * Tn + t) * P.zrows + r0);
            float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f), z1 = z0;
            const bool in0 = r0 >= 0 && r0 < P.zrows, in1 = r0 + 1 >= 0 && r0 + 1 < P.zrows;
            if (in0) z0 = zp[0];
            if (in1) z1 = zp[1];
            v.x = in0 ? pack_bf16x2((z0.x - mean) * inv, (z0.y - mean) * inv) : 0u;
            v.y = in0 ? pack_bf16x2((z0.z - mean) * inv, (z0.w - mean) * inv) : 0u;
            v.z = in1 ? pack_bf16x2((z1.x - mean) * inv, (z1.y - mean) * inv) : 0u;
            v.w = in1 ? pack_bf16x2((z1.z - mean) * inv, (z1.w - mean) * inv) : 0u;
          }
          *(uint4*)(pst + t * D::PP + ch * 8) = v;
        }
      } else
      for (int i = tid; i < MT * 16 * 4; i += ER_THREADS) {   // 4 x 16-byte chunks per frame; frames >= Tn are zero
        const int t = i >> 2, ch = i & 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t < Tn) v = *(const uint4*)(xin + xis.row_off_bt(b, t, 4 * f - 2) + ch * 8);
        *(uint4*)(pst + t * D::PP + ch * 8) = v;
      }
      er_gsync(grp);
      for (int mt = warp; mt < MT; mt += nwarp) {
        uint32_t a[2][4];
        ldsm_a(pst, D::PP, mt * 16, 0, lane, a[0]);
        ldsm_a(pst, D::PP, mt * 16, 16, lane, a[1]);
#pragma unroll
        for (int nt = 0; nt < D::AT; ++nt) {
          float d[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t b0[2], b1[2];
          ldsm_b2(cwt, D::PP, nt * 8, 0, lane, b0, b1);
          mma16816(d, a[0], b0);
          mma16816(d, a[1], b1);
          const int c = nt * 8 + 2 * q, r0 = mt * 16 + g;
          const float2 cb = *(const float2*)(cbs + c);
          if (r0 < Tn) { const float2 g2 = gelu_fast2(f2add(make_float2(d[0], d[1]), make_float2(cb.x, cb.y))); sts_pair(xsi + r0 * D::XP + c, g2.x, g2.y); }
          if (r0 + 8 < Tn) { const float2 g2 = gelu_fast2(f2add(make_float2(d[2], d[3]), make_float2(cb.x, cb.y))); sts_pair(xsi + (r0 + 8) * D::XP + c, g2.x, g2.y); }
        }
      }
    } else {
      for (int i = tid; i < Tn * (C / 8); i += ER_THREADS) {
        const int t = i / (C / 8), c = (i - t * (C / 8)) * 8;
        *(uint4*)(xsi + t * D::XP + c) = *(const uint4*)(yin + ys.row_off_bt(b, t, f) + c);
      }
    }

    if (FUSE_CONV) {
      // the patch tile becomes the hidden operand tile: its K padding (columns >= 8*HN of 16) must read as zero
      er_gsync(grp);
      for (int i = tid; i < MT * 16; i += ER_THREADS)
        for (int c = 8 * D::HN; c < 16; c += 8) *(uint4*)(hb + i * D::K2P + c) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int dd = 0; dd < 2; ++dd) {
      const int dil = 1 << dd;
      const bf16* w1 = w1t + dd * 8 * D::HN * D::K1P;
      const bf16* w2 = w2t + dd * 2 * C * D::K2P;
      const float* v = vec + dd * D::VD;
      const float* b1 = v, *g1w = v + 16, *g1b = v + 32, *b2 = v + 48, *g2w = b2 + 2 * C, *g2b = b2 + 4 * C, *scl = b2 + 6 * C;
      er_gsync(grp);
      // ---- h = conv_k3_dil(y) + b1 (taps = row-shifted A fragments of the resident slab); GroupNorm(1,H) partial sums
      float hacc[2][D::HN][4];
      float s1 = 0.f, q1 = 0.f;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mt = warp + mi * nwarp;
#pragma unroll
        for (int nt = 0; nt < D::HN; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) hacc[mi][nt][e] = 0.f;
        if (mt < MT) {
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
            const bf16* xr = xsi + (tap - 1) * dil * D::XP;
#pragma unroll
            for (int kc = 0; kc < C / 16; ++kc) {
              uint32_t a[4];
              ldsm_a(xr, D::XP, mt * 16, kc * 16, lane, a);
#pragma unroll
              for (int nt = 0; nt < D::HN; ++nt) {
                uint32_t bb[2];
                frag_b(w1, D::K1P, nt * 8, tap * C + kc * 16, lane, bb);
                mma16816(hacc[mi][nt], a, bb);
              }
            }
          }
#pragma unroll
          for (int nt = 0; nt < D::HN; ++nt) {
            const int c = nt * 8 + 2 * q, r0 = mt * 16 + g;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = c + (e & 1), rr = r0 + (e >> 1) * 8;
              hacc[mi][nt][e] += b1[cc];
              if (rr < Tn && cc < D::H) { s1 += hacc[mi][nt][e]; q1 += hacc[mi][nt][e] * hacc[mi][nt][e]; }
            }
          }
        }
      }
      er_block_sum2(s1, q1, red, grp, tid);
      {
        const float n = (float)(Tn * D::H);
        const float mean = s1 / n;
        const float rstd = rsqrtf(fmaxf(q1 / n - mean * mean, 0.f) + 1e-5f);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const int mt = warp + mi * nwarp;
          if (mt < MT) {
#pragma unroll
            for (int nt = 0; nt < D::HN; ++nt) {
              const int c = nt * 8 + 2 * q, r0 = mt * 16 + g;
              float y[4];      // column H carries the constant one of h~; rows past the slab are all-zero (they must not count)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int cc = c + (e & 1);
                const bool rv = r0 + (e >> 1) * 8 < Tn;
                y[e] = !rv ? 0.f : cc < D::H ? gelu_fast((hacc[mi][nt][e] - mean) * rstd * g1w[cc] + g1b[cc]) : cc == D::H ? 1.0f : 0.f;
              }
              sts_pair(hb + r0 * D::K2P + c, y[0], y[1]);
              sts_pair(hb + (r0 + 8) * D::K2P + c, y[2], y[3]);
            }
          }
        }
      }
      er_gsync(grp);

      // ---- e = W2 h + b2 : GroupNorm(1, 2C) statistics from D = h~^T h~ (one or two MMAs per 16-row tile instead of the 2C-wide
      //      product), then the GroupNorm + GLU + LayerScale + residual pass
      float s2 = 0.f, q2 = 0.f;
      {
        float dacc[D::HN][4];
#pragma unroll
        for (int nt = 0; nt < D::HN; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) dacc[nt][e] = 0.f;
        for (int mt = warp; mt < MT; mt += nwarp) {
          uint32_t t[4];
          ldsm_x4_trans(hb, D::K2P, mt * 16, 0, lane, t);
          { const uint32_t b0[2] = {t[0], t[2]}; mma16816(dacc[0], t, b0); }
          if (D::HN == 2) { const uint32_t b1[2] = {t[1], t[3]}; mma16816(dacc[D::HN - 1], t, b1); }
        }
        const float* gqd = gq + dd * 256;
        const float* gsd = gs + dd * 16;
#pragma unroll
        for (int nt = 0; nt < D::HN; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = g + (e >> 1) * 8, j = nt * 8 + 2 * q + (e & 1);
            if (i <= D::H) {       // (rows past the constant one are zero)
              q2 += gqd[i * 16 + j] * dacc[nt][e];
              if (i == D::H) s2 += gsd[j] * dacc[nt][e];
            }
          }
      }
      er_block_sum2(s2, q2, red, grp, tid);
      {
        const float n2 = (float)(Tn * 2 * C);
        const float mean2 = s2 / n2;
        const float rstd2 = rsqrtf(fmaxf(q2 / n2 - mean2 * mean2, 0.f) + 1e-5f);
        // e_norm = d * alpha + beta; the gate half carries the x/2 of tanh, the value half the LayerScale factor of its column
        for (int n = tid; n < 2 * C; n += ER_THREADS) {
          const float half = n >= C ? 0.5f : scl[n];
          const float a_ = rstd2 * g2w[n];
          al[n] = half * a_; be[n] = half * ((b2[n] - mean2) * a_ + g2b[n]);
        }
      }
      er_gsync(grp);
      {
        // the warp's two row tiles share the B fragments (one ldmatrix.x4: value + gate n-tile) and the per-column vectors
        const int mt0 = warp, mt1 = warp + nwarp;
        uint32_t a0[4], a1[4];
        if (mt0 < MT) ldsm_a(hb, D::K2P, mt0 * 16, 0, lane, a0);
        if (mt1 < MT) ldsm_a(hb, D::K2P, mt1 * 16, 0, lane, a1);
#pragma unroll
        for (int nt = 0; nt < D::AT; ++nt) {
          uint32_t bv_[2], bg_[2];
          ldsm_b_pair(w2, D::K2P, nt * 8, (nt + D::AT) * 8, 0, lane, bv_, bg_);
          const int c = nt * 8 + 2 * q;
          const float2 av = *(const float2*)(al + c), bv = *(const float2*)(be + c);
          const float2 ag = *(const float2*)(al + C + c), bg = *(const float2*)(be + C + c);
          auto tile = [&](const uint32_t (&a)[4], int mt) {
            float da[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(da, a, bv_);
            mma16816(dg, a, bg_);
            const int r0 = mt * 16 + g;
            // packed fp32 pairs: x + (value * alpha + beta) * (0.5 + 0.5 tanh(gate * alpha' + beta')); alpha, beta carry the LayerScale
            const float2 h2 = f2splat(0.5f);
            const float2 g01 = f2fma(make_float2(dg[0], dg[1]), ag, bg), g23 = f2fma(make_float2(dg[2], dg[3]), ag, bg);
            const float2 s01 = f2fma(h2, make_float2(er_tanh(g01.x), er_tanh(g01.y)), h2);
            const float2 s23 = f2fma(h2, make_float2(er_tanh(g23.x), er_tanh(g23.y)), h2);
            const float2 u01 = f2fma(make_float2(da[0], da[1]), av, bv), u23 = f2fma(make_float2(da[2], da[3]), av, bv);
            if (r0 < Tn) {
              const float2 o = f2fma(u01, s01, unpack_bf16x2(*(const uint32_t*)(xsi + r0 * D::XP + c)));
              sts_pair(xsi + r0 * D::XP + c, o.x, o.y);
            }
            if (r0 + 8 < Tn) {
              const float2 o = f2fma(u23, s23, unpack_bf16x2(*(const uint32_t*)(xsi + (r0 + 8) * D::XP + c)));
              sts_pair(xsi + (r0 + 8) * D::XP + c, o.x, o.y);
            }
          };
          if (mt0 < MT) tile(a0, mt0);
          if (mt1 < MT) tile(a1, mt1);
        }
      }
    }
    er_gsync(grp);

    // ---- rewrite 1x1 (C -> 2C) + GLU (+ frequency embedding): result overwrites the slab rows of the owning warp
    {
      // TPW row tiles per pass of a warp (C = 48: both of them) share the weight fragments of an n-tile; the gate rows were staged
      // pre-scaled by 0.5 (exact in bf16) and the accumulators start from the bias: GLU = value * (0.5 + 0.5 tanh(gate))
      constexpr int TPW = 1;
      for (int mtb = warp; mtb < MT; mtb += TPW * nwarp) {
        uint32_t a[TPW][C / 16][4];
#pragma unroll
        for (int ti = 0; ti < TPW; ++ti)
          if (mtb + ti * nwarp < MT) {
#pragma unroll
            for (int kc = 0; kc < C / 16; ++kc) ldsm_a(xsi, D::XP, (mtb + ti * nwarp) * 16, kc * 16, lane, a[ti][kc]);
          }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < D::AT; ++nt) {
          uint32_t bvv[C / 16][2], bgg[C / 16][2];
#pragma unroll
          for (int kc = 0; kc + 1 < C / 16; kc += 2) {
            ldsm_b2(wrt, D::KRP, nt * 8, kc * 16, lane, bvv[kc], bvv[kc + 1]);
            ldsm_b2(wrt, D::KRP, (nt + D::AT) * 8, kc * 16, lane, bgg[kc], bgg[kc + 1]);
          }
          if ((C / 16) & 1) ldsm_b_pair(wrt, D::KRP, nt * 8, (nt + D::AT) * 8, (C / 16 - 1) * 16, lane, bvv[C / 16 - 1], bgg[C / 16 - 1]);
          const int c = nt * 8 + 2 * q;
          const float2 rv = *(const float2*)(rbs + c), rg = *(const float2*)(rbs + C + c);
          float2 em = make_float2(0.f, 0.f);
          if (P.emb) em = *(const float2*)(embv + c);
#pragma unroll
          for (int ti = 0; ti < TPW; ++ti) {
            const int mt = mtb + ti * nwarp;
            if (mt < MT) {
              float da[4] = {rv.x, rv.y, rv.x, rv.y}, dg[4] = {rg.x, rg.y, rg.x, rg.y};
#pragma unroll
              for (int kc = 0; kc < C / 16; ++kc) { mma16816(da, a[ti][kc], bvv[kc]); mma16816(dg, a[ti][kc], bgg[kc]); }
              const int r0 = mt * 16 + g;
              const float2 h2 = f2splat(0.5f);
              const float2 s01 = f2fma(h2, make_float2(er_tanh(dg[0]), er_tanh(dg[1])), h2);
              const float2 s23 = f2fma(h2, make_float2(er_tanh(dg[2]), er_tanh(dg[3])), h2);
              const float2 p01 = f2fma(make_float2(da[0], da[1]), s01, em), p23 = f2fma(make_float2(da[2], da[3]), s23, em);
              if (r0 < Tn) sts_pair(xsi + r0 * D::XP + c, p01.x, p01.y);            // rows >= Tn stay zero (conv halo of the next slab)
              if (r0 + 8 < Tn) sts_pair(xsi + (r0 + 8) * D::XP + c, p23.x, p23.y);
            }
          }
        }
      }
    }
    er_gsync(grp);
    for (int i = tid; i < Tn * (C / 8); i += ER_THREADS) {
      const int t = i / (C / 8), c = (i - t * (C / 8)) * 8;
      *(uint4*)(out + ys.row_off_bt(b, t, f) + c) = *(const uint4*)(xsi + t * D::XP + c);
    }
    er_gsync(grp);
  }
}

template <int C>
static size_t enc_row_smem(int Tn, bool fuse, int G) {
  typedef ErDims<C> D;
  const int MT = (Tn + 15) / 16;
  const size_t wbf = (size_t)2 * 8 * D::HN * D::K1P + (size_t)2 * 2 * C * D::K2P + (size_t)2 * C * D::KRP + (fuse ? (size_t)C * D::PP : 0);
  const size_t slab = (size_t)(MT * 16 + 4) * D::XP + (fuse ? (size_t)MT * 16 * D::PP : (size_t)MT * 16 * D::K2P);
  return wbf * 2 + sizeof(float) * (2 * D::VD + 2 * C + C + 2 * 256 + 32 + (size_t)G * (4 * C + C + 64)) + (size_t)G * slab * 2 + 32;
}

bool enc_row_supported(int C, int Tn, bool fuse_conv) {
  if (Tn > 16 * 2 * (ER_THREADS / 32)) return false;          // two row tiles per warp
  if (C == 48) return enc_row_smem<48>(Tn, fuse_conv, 1) <= 113 * 1024;
  if (C == 96) return !fuse_conv && enc_row_smem<96>(Tn, false, 1) <= 227 * 1024;
  return false;
}

static int er_num_sms() { return device_sm_count(); }

void launch_enc_row(const bf16* xin, RowSpace xis, const bf16* yin, bf16* out, RowSpace ys, const EncRowParams& P, bool fuse_conv,
                    cudaStream_t st) {
  const int Tn = ys.G2;
  const int slabs = ys.batch() * ys.R;
  if (ys.C == 48) {
    // level 0 (fused conv): three 9-warp slab groups in one CTA per SM (27 warps) if they fit, else two CTAs of one group
    const bool g3 = fuse_conv && enc_row_smem<48>(Tn, true, 3) <= 226 * 1024;
    if (g3) {
      const size_t smem = enc_row_smem<48>(Tn, true, 3);
      cudaFuncSetAttribute(enc_row_kernel<48, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_pdl(enc_row_kernel<48, true, 3>, dim3(std::min((slabs + 2) / 3, er_num_sms())), dim3(3 * ER_THREADS), smem, st, xin, xis, yin, out, ys, P);
      return;
    }
    const size_t smem = enc_row_smem<48>(Tn, fuse_conv, 1);
    const int grid = std::min(slabs, 2 * er_num_sms());
    if (fuse_conv) {
      cudaFuncSetAttribute(enc_row_kernel<48, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_pdl(enc_row_kernel<48, true, 1>, dim3(grid), dim3(ER_THREADS), smem, st, xin, xis, yin, out, ys, P);
    } else {
      cudaFuncSetAttribute(enc_row_kernel<48, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_pdl(enc_row_kernel<48, false, 1>, dim3(grid), dim3(ER_THREADS), smem, st, xin, xis, yin, out, ys, P);
    }
  } else {
    // C = 96: the weights alone are 77 KB, so one CTA per SM; two 9-warp slab groups share them (18 warps per SM)
    const int G = enc_row_smem<96>(Tn, false, 2) <= 227 * 1024 ? 2 : 1;
    const size_t smem = enc_row_smem<96>(Tn, false, G);
    const int grid = std::min((slabs + G - 1) / G, er_num_sms());
    if (G == 2) {
      cudaFuncSetAttribute(enc_row_kernel<96, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_pdl(enc_row_kernel<96, false, 2>, dim3(grid), dim3(2 * ER_THREADS), smem, st, xin, xis, yin, out, ys, P);
    } else {
      cudaFuncSetAttribute(enc_row_kernel<96, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_pdl(enc_row_kernel<96, false, 1>, dim3(grid), dim3(ER_THREADS), smem, st, xin, xis, yin, out, ys, P);
    }
  }
}

}  // namespace athtd
