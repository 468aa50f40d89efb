// Fused frequency-encoder layer tail for one (segment, frequency row) slab, bf16 build:
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
// The contractions are tiny (K = 32 / 3C / 16 / C on 16-row tiles of the resident slab), so they use the warp-level
// tensor-core MMA (mma.sync.m16n8k16 bf16 -> fp32) straight from shared memory: a tcgen05 / TMEM round trip per
// 259-row slab would cost more than the math.  The large, compute-bound GEMMs of the path stay on gemm_tc_kernel.
#include "kernels.cuh"

namespace athtd {

static constexpr int ER_THREADS = 288;     // 9 warps; a 6 s segment has 259 frames = 17 row tiles of 16

struct EncRowParams {
  // level-0 conv (only if fuse_conv): weight [C][4][8], bias [C]
  const float* cw; const float* cb;
  // per depth: conv3 weight [H][C][3], bias [H], gn1 w/b [H], expand weight [2C][H], bias [2C], gn2 w/b [2C], scale [C]
  const float* w1[2]; const float* b1[2]; const float* g1w[2]; const float* g1b[2];
  const float* w2[2]; const float* b2[2]; const float* g2w[2]; const float* g2b[2]; const float* scale[2];
  // rewrite weight [2C][C], bias [2C]; optional per-frequency table [F][C] with scale
  const float* rw; const float* rb; const float* emb; float emb_scale;
};

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *(const uint32_t*)p; }
__device__ __forceinline__ void sts_pair(bf16* p, float a, float b) { *(__nv_bfloat162*)p = __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float sig_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }

__device__ __forceinline__ void er_block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
  __syncthreads();
  float x = l < nw ? red[2 * l] : 0.f, y = l < nw ? red[2 * l + 1] : 0.f;
  a = warp_sum(x); b = warp_sum(y);
}

// A fragment of a 16x16 tile of a row-major bf16 matrix in shared memory (pitch P elements)
__device__ __forceinline__ void load_a(const bf16* S, int P, int r0, int k0, int lane, uint32_t (&a)[4]) {
  const int g = lane >> 2, q = lane & 3;
  a[0] = lds32(S + (r0 + g) * P + k0 + 2 * q);
  a[1] = lds32(S + (r0 + g + 8) * P + k0 + 2 * q);
  a[2] = lds32(S + (r0 + g) * P + k0 + 2 * q + 8);
  a[3] = lds32(S + (r0 + g + 8) * P + k0 + 2 * q + 8);
}
// B fragment (16 x 8) from weights stored [n][k] (k contiguous, pitch KP)
__device__ __forceinline__ void load_b(const bf16* W, int KP, int n0, int k0, int lane, uint32_t (&b)[2]) {
  const int g = lane >> 2, q = lane & 3;
  b[0] = lds32(W + (n0 + g) * KP + k0 + 2 * q);
  b[1] = lds32(W + (n0 + g) * KP + k0 + 2 * q + 8);
}

template <int C, bool FUSE_CONV>
__global__ void __launch_bounds__(ER_THREADS, (C == 48 ? 2 : 1))
enc_row_kernel(const bf16* __restrict__ xin, RowSpace xis, const bf16* __restrict__ yin, bf16* __restrict__ out, RowSpace ys,
               EncRowParams P) {
  constexpr int H = C / 8;
  constexpr int HN = (H + 7) / 8;                 // hidden n-tiles (1 or 2)
  constexpr int HP = 16;                          // hidden K padded to one k-step
  constexpr int XP = C + 8;                       // slab pitch: conflict-free fragment loads
  constexpr int K1 = 3 * C, K1P = K1 + 8;         // conv-k3 weights [8*HN][K1P]
  constexpr int K2P = HP + 8;                     // expand weights  [2C][K2P]
  constexpr int KRP = C + 8;                      // rewrite weights [2C][KRP]
  constexpr int AT = C / 8;                       // value n-tiles (gates are tiles AT .. 2AT-1)
  const int Tn = ys.G2;
  const int MT = (Tn + 15) / 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* xs = (bf16*)smem_raw;                               // [(MT*16 + 4)][XP], rows shifted by +2 (zero halo)
  bf16* hb = xs + (MT * 16 + 4) * XP;                       // [MT*16][K2P]
  bf16* w1t = hb + MT * 16 * K2P;                           // [8*HN][K1P]
  bf16* w2t = w1t + 8 * HN * K1P;                           // [2C][K2P]
  bf16* wrt = w2t + 2 * C * K2P;                            // [2C][KRP]   (also conv weights [C][40] on level 0, before use)
  float* vec = (float*)(wrt + 2 * C * KRP);                 // b1[16] g1w[16] g1b[16] b2[2C] g2w[2C] g2b[2C] scale[C] rb[2C] emb[C] cb[C]
  float* red = vec + 48 + 10 * C;                           // 64 floats
  const int f = blockIdx.x % ys.R, b = blockIdx.x / ys.R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  bf16* xsi = xs + 2 * XP;                                  // interior row 0

  // ---- zero the slab (halo rows, rows >= T, pad columns)
  for (int i = tid; i < ((MT * 16 + 4) * XP) / 2; i += blockDim.x) ((uint32_t*)xs)[i] = 0u;

  if (FUSE_CONV) {
    // level 0: y = GELU(conv_k8s4(x)): patch [T x 32] = 8 padded frequency rows x 4 channels, contiguous in xin.
    // The patch and the conv weights are staged in the (not yet used) hidden-operand / weight area.
    constexpr int PP = 40;
    bf16* pstage = hb;                                      // [MT*16][PP]
    bf16* cwt = hb + MT * 16 * PP;                          // [C][PP]
    for (int i = tid; i < (MT * 16 * PP) / 2; i += blockDim.x) ((uint32_t*)pstage)[i] = 0u;
    __syncthreads();
    for (int i = tid; i < Tn * 4; i += blockDim.x) {        // 4 x 16-byte chunks per frame
      const int t = i >> 2, ch = i & 3;
      const uint4 v = *(const uint4*)(xin + xis.row_off(b * Tn + t, 4 * f - 2) + ch * 8);
      *(uint4*)(pstage + t * PP + ch * 8) = v;
    }
    for (int i = tid; i < C * 32; i += blockDim.x) {        // cwt[co][j*4+ci] = W[co][ci][j]
      const int co = i / 32, k = i % 32, j = k >> 2, ci = k & 3;
      cwt[co * PP + k] = __float2bfloat16_rn(P.cw[(co * 4 + ci) * 8 + j]);
    }
    for (int i = tid; i < C; i += blockDim.x) vec[48 + 9 * C + i] = P.cb[i];
    __syncthreads();
    const float* cb = vec + 48 + 9 * C;
    for (int mt = warp; mt < MT; mt += nwarp) {
      uint32_t a[2][4];
      load_a(pstage, PP, mt * 16, 0, lane, a[0]);
      load_a(pstage, PP, mt * 16, 16, lane, a[1]);
#pragma unroll
      for (int nt = 0; nt < AT; ++nt) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t bb[2];
        load_b(cwt, PP, nt * 8, 0, lane, bb); mma16816(d, a[0][0], a[0][1], a[0][2], a[0][3], bb[0], bb[1]);
        load_b(cwt, PP, nt * 8, 16, lane, bb); mma16816(d, a[1][0], a[1][1], a[1][2], a[1][3], bb[0], bb[1]);
        const int c = nt * 8 + 2 * q, r0 = mt * 16 + g;
        if (r0 < Tn) sts_pair(xsi + r0 * XP + c, gelu_fast(d[0] + cb[c]), gelu_fast(d[1] + cb[c + 1]));
        if (r0 + 8 < Tn) sts_pair(xsi + (r0 + 8) * XP + c, gelu_fast(d[2] + cb[c]), gelu_fast(d[3] + cb[c + 1]));
      }
    }
  } else {
    __syncthreads();
    for (int i = tid; i < Tn * (C / 8); i += blockDim.x) {
      const int t = i / (C / 8), c = (i % (C / 8)) * 8;
      *(uint4*)(xsi + t * XP + c) = *(const uint4*)(yin + ys.row_off(b * Tn + t, f) + c);
    }
  }

  __syncthreads();
  for (int i = tid; i < (MT * 16 * K2P) / 2; i += blockDim.x) ((uint32_t*)hb)[i] = 0u;     // hidden operand: K padding must be 0

  for (int dd = 0; dd < 2; ++dd) {
    const int dil = 1 << dd;
    __syncthreads();
    // ---- stage this depth's weights as bf16 [n][k] and the fp32 vectors
    for (int i = tid; i < 8 * HN * K1; i += blockDim.x) {
      const int j = i / K1, kk = i % K1, tap = kk / C, c = kk % C;
      w1t[j * K1P + kk] = __float2bfloat16_rn(j < H ? P.w1[dd][(j * C + c) * 3 + tap] : 0.f);
    }
    for (int i = tid; i < 2 * C * HP; i += blockDim.x) {
      const int n = i / HP, k = i % HP;
      w2t[n * K2P + k] = __float2bfloat16_rn(k < H ? P.w2[dd][n * H + k] : 0.f);
    }
    for (int i = tid; i < 16; i += blockDim.x) {
      vec[i] = i < H ? P.b1[dd][i] : 0.f; vec[16 + i] = i < H ? P.g1w[dd][i] : 0.f; vec[32 + i] = i < H ? P.g1b[dd][i] : 0.f;
    }
    for (int i = tid; i < 2 * C; i += blockDim.x) {
      vec[48 + i] = P.b2[dd][i]; vec[48 + 2 * C + i] = P.g2w[dd][i]; vec[48 + 4 * C + i] = P.g2b[dd][i];
    }
    for (int i = tid; i < C; i += blockDim.x) vec[48 + 6 * C + i] = P.scale[dd][i];
    __syncthreads();
    const float* b1 = vec, *g1w = vec + 16, *g1b = vec + 32, *b2 = vec + 48, *g2w = b2 + 2 * C, *g2b = b2 + 4 * C, *scl = b2 + 6 * C;

    // ---- h = conv_k3_dil(y) + b1 (taps = row-shifted A fragments of the resident slab); GroupNorm(1,H) partial sums
    float hacc[2][HN][4];
    float s1 = 0.f, q1 = 0.f;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int mt = warp + mi * nwarp;
#pragma unroll
      for (int nt = 0; nt < HN; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) hacc[mi][nt][e] = 0.f;
      if (mt < MT) {
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
          const bf16* xr = xsi + (tap - 1) * dil * XP;
#pragma unroll
          for (int kc = 0; kc < C / 16; ++kc) {
            uint32_t a[4];
            load_a(xr, XP, mt * 16, kc * 16, lane, a);
#pragma unroll
            for (int nt = 0; nt < HN; ++nt) {
              uint32_t bb[2];
              load_b(w1t, K1P, nt * 8, tap * C + kc * 16, lane, bb);
              mma16816(hacc[mi][nt], a[0], a[1], a[2], a[3], bb[0], bb[1]);
            }
          }
        }
#pragma unroll
        for (int nt = 0; nt < HN; ++nt) {
          const int c = nt * 8 + 2 * q, r0 = mt * 16 + g;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c + (e & 1), rr = r0 + (e >> 1) * 8;
            hacc[mi][nt][e] += b1[cc];
            if (rr < Tn && cc < H) { s1 += hacc[mi][nt][e]; q1 += hacc[mi][nt][e] * hacc[mi][nt][e]; }
          }
        }
      }
    }
    er_block_sum2(s1, q1, red);
    {
      const float n = (float)(Tn * H);
      const float mean = s1 / n;
      const float rstd = rsqrtf(fmaxf(q1 / n - mean * mean, 0.f) + 1e-5f);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mt = warp + mi * nwarp;
        if (mt < MT) {
#pragma unroll
          for (int nt = 0; nt < HN; ++nt) {
            const int c = nt * 8 + 2 * q, r0 = mt * 16 + g;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = c + (e & 1);
              v[e] = cc < H ? gelu_fast((hacc[mi][nt][e] - mean) * rstd * g1w[cc] + g1b[cc]) : 0.f;
            }
            sts_pair(hb + r0 * K2P + c, v[0], v[1]);
            sts_pair(hb + (r0 + 8) * K2P + c, v[2], v[3]);
          }
        }
      }
    }
    __syncthreads();

    // ---- e = W2 h + b2 : statistics pass (nothing stored), then GroupNorm + GLU + LayerScale + residual pass
    float s2 = 0.f, q2 = 0.f;
    for (int mt = warp; mt < MT; mt += nwarp) {
      uint32_t a[4];
      load_a(hb, K2P, mt * 16, 0, lane, a);
      const int r0 = mt * 16 + g;
      const bool v0 = r0 < Tn, v1 = r0 + 8 < Tn;
#pragma unroll 4
      for (int nt = 0; nt < 2 * AT; ++nt) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t bb[2];
        load_b(w2t, K2P, nt * 8, 0, lane, bb);
        mma16816(d, a[0], a[1], a[2], a[3], bb[0], bb[1]);
        const int c = nt * 8 + 2 * q;
        const float e0 = d[0] + b2[c], e1 = d[1] + b2[c + 1], e2 = d[2] + b2[c], e3 = d[3] + b2[c + 1];
        if (v0) { s2 += e0 + e1; q2 += e0 * e0 + e1 * e1; }
        if (v1) { s2 += e2 + e3; q2 += e2 * e2 + e3 * e3; }
      }
    }
    er_block_sum2(s2, q2, red);
    const float n2 = (float)(Tn * 2 * C);
    const float mean2 = s2 / n2;
    const float rstd2 = rsqrtf(fmaxf(q2 / n2 - mean2 * mean2, 0.f) + 1e-5f);
    for (int mt = warp; mt < MT; mt += nwarp) {
      uint32_t a[4];
      load_a(hb, K2P, mt * 16, 0, lane, a);
      const int r0 = mt * 16 + g;
#pragma unroll 2
      for (int nt = 0; nt < AT; ++nt) {
        float da[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t bb[2];
        load_b(w2t, K2P, nt * 8, 0, lane, bb); mma16816(da, a[0], a[1], a[2], a[3], bb[0], bb[1]);
        load_b(w2t, K2P, (nt + AT) * 8, 0, lane, bb); mma16816(dg, a[0], a[1], a[2], a[3], bb[0], bb[1]);
        const int c = nt * 8 + 2 * q;
        float u[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cc = c + (e & 1);
          const float ea = (da[e] + b2[cc] - mean2) * rstd2 * g2w[cc] + g2b[cc];
          const float eg = (dg[e] + b2[cc + C] - mean2) * rstd2 * g2w[cc + C] + g2b[cc + C];
          u[e] = scl[cc] * (ea * sig_fast(eg));
        }
        if (r0 < Tn) {
          const float2 x0 = __bfloat1622float2(*(const __nv_bfloat162*)(xsi + r0 * XP + c));
          sts_pair(xsi + r0 * XP + c, x0.x + u[0], x0.y + u[1]);
        }
        if (r0 + 8 < Tn) {
          const float2 x1 = __bfloat1622float2(*(const __nv_bfloat162*)(xsi + (r0 + 8) * XP + c));
          sts_pair(xsi + (r0 + 8) * XP + c, x1.x + u[2], x1.y + u[3]);
        }
      }
    }
  }
  __syncthreads();

  // ---- rewrite 1x1 (C -> 2C) + GLU (+ frequency embedding): result overwrites the slab rows of the owning warp
  for (int i = tid; i < 2 * C * C; i += blockDim.x) {
    const int n = i / C, k = i % C;
    wrt[n * KRP + k] = __float2bfloat16_rn(P.rw[i]);
  }
  for (int i = tid; i < 2 * C; i += blockDim.x) vec[48 + 7 * C + i] = P.rb[i];
  for (int i = tid; i < C; i += blockDim.x) vec[48 + 6 * C + i] = P.emb ? P.emb_scale * P.emb[f * C + i] : 0.f;
  __syncthreads();
  const float* rb = vec + 48 + 7 * C, *embv = vec + 48 + 6 * C;
  for (int mt = warp; mt < MT; mt += nwarp) {
    uint32_t a[C / 16][4];
#pragma unroll
    for (int kc = 0; kc < C / 16; ++kc) load_a(xsi, XP, mt * 16, kc * 16, lane, a[kc]);
    __syncwarp();
    const int r0 = mt * 16 + g;
#pragma unroll 2
    for (int nt = 0; nt < AT; ++nt) {
      float da[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kc = 0; kc < C / 16; ++kc) {
        uint32_t bb[2];
        load_b(wrt, KRP, nt * 8, kc * 16, lane, bb); mma16816(da, a[kc][0], a[kc][1], a[kc][2], a[kc][3], bb[0], bb[1]);
        load_b(wrt, KRP, (nt + AT) * 8, kc * 16, lane, bb); mma16816(dg, a[kc][0], a[kc][1], a[kc][2], a[kc][3], bb[0], bb[1]);
      }
      const int c = nt * 8 + 2 * q;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cc = c + (e & 1);
        o[e] = (da[e] + rb[cc]) * sig_fast(dg[e] + rb[cc + C]) + embv[cc];
      }
      sts_pair(xsi + r0 * XP + c, o[0], o[1]);
      sts_pair(xsi + (r0 + 8) * XP + c, o[2], o[3]);
    }
  }
  __syncthreads();
  for (int i = tid; i < Tn * (C / 8); i += blockDim.x) {
    const int t = i / (C / 8), c = (i % (C / 8)) * 8;
    *(uint4*)(out + ys.row_off(b * Tn + t, f) + c) = *(const uint4*)(xsi + t * XP + c);
  }
}

template <int C>
static size_t enc_row_smem(int Tn) {
  constexpr int H = C / 8, HN = (H + 7) / 8;
  const int MT = (Tn + 15) / 16;
  size_t bf = (size_t)(MT * 16 + 4) * (C + 8) + (size_t)MT * 16 * 24 + (size_t)8 * HN * (3 * C + 8) + (size_t)2 * C * 24 + (size_t)2 * C * (C + 8);
  return bf * 2 + sizeof(float) * (48 + 10 * C + 64) + 32;
}

bool enc_row_supported(int C, int Tn, bool fuse_conv) {
  if (Tn > 16 * 2 * (ER_THREADS / 32)) return false;          // two row tiles per warp
  // level-0 staging (patch [MT*16][40] + conv weights [48][40]) must fit in the hidden-operand + weight area
  const int MT = (Tn + 15) / 16;
  if (C == 48) return enc_row_smem<48>(Tn) <= 110 * 1024 &&
                      (!fuse_conv || MT * 16 * 40 + 48 * 40 <= MT * 16 * 24 + 8 * (3 * 48 + 8) + 2 * 48 * 24 + 2 * 48 * 56);
  if (C == 96) return !fuse_conv && enc_row_smem<96>(Tn) <= 220 * 1024;
  return false;
}

// ptrs: [cw, cb] + 2 x [w1, b1, g1w, g1b, w2, b2, g2w, g2b, scale] + [rw, rb, emb]
void launch_enc_row(const bf16* xin, RowSpace xis, const bf16* yin, bf16* out, RowSpace ys, const float* const* ptrs, float emb_scale,
                    bool fuse_conv, cudaStream_t st) {
  EncRowParams P;
  P.cw = ptrs[0]; P.cb = ptrs[1];
  for (int d = 0; d < 2; ++d) {
    const float* const* qd = ptrs + 2 + 9 * d;
    P.w1[d] = qd[0]; P.b1[d] = qd[1]; P.g1w[d] = qd[2]; P.g1b[d] = qd[3]; P.w2[d] = qd[4]; P.b2[d] = qd[5]; P.g2w[d] = qd[6];
    P.g2b[d] = qd[7]; P.scale[d] = qd[8];
  }
  P.rw = ptrs[20]; P.rb = ptrs[21]; P.emb = ptrs[22]; P.emb_scale = emb_scale;
  const int Tn = ys.G2;
  const int blocks = ys.batch() * ys.R;
  if (ys.C == 48) {
    const size_t smem = enc_row_smem<48>(Tn);
    if (fuse_conv) {
      cudaFuncSetAttribute(enc_row_kernel<48, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      enc_row_kernel<48, true><<<blocks, ER_THREADS, smem, st>>>(xin, xis, yin, out, ys, P);
    } else {
      cudaFuncSetAttribute(enc_row_kernel<48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      enc_row_kernel<48, false><<<blocks, ER_THREADS, smem, st>>>(xin, xis, yin, out, ys, P);
    }
  } else {
    const size_t smem = enc_row_smem<96>(Tn);
    cudaFuncSetAttribute(enc_row_kernel<96, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    enc_row_kernel<96, false><<<blocks, ER_THREADS, smem, st>>>(xin, xis, yin, out, ys, P);
  }
}

}  // namespace athtd
