// Fused DConv residual branch for the FREQUENCY encoder layers (demucs DConv, depth 2).
//
// Inside a frequency-branch HEncLayer the tensor is reshaped to [B*Fr, C, T], so both GroupNorm(1, .) statistics of
// every DConv layer are local to one (segment, frequency-row): a [T x C] slab.  One CTA owns one such slab, keeps it
// in shared memory and runs BOTH residual layers on it
//     x += scale * GLU( GN2( W2 * GELU( GN1( conv_k3_dil(x) ) ) ) )         (dilation 1, then 2)
// so the slab is read from and written to HBM exactly once (the unfused path moves it ~7 times per layer and the
// 2C-wide expand tensor twice).  The contractions are tiny (K = 3C -> C/8 -> 2C), so this kernel is bandwidth-bound
// and uses CUDA cores; the expand output is never materialised: it is evaluated once for the statistics and once
// more for the normalised GLU.
#include "kernels.cuh"

namespace athtd {

template <typename T> struct SmemT;
template <> struct SmemT<float> { typedef float type; };
template <> struct SmemT<bf16> { typedef bf16 type; };

__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
  __syncthreads();
  float x = l < nw ? red[2 * l] : 0.f, y = l < nw ? red[2 * l + 1] : 0.f;
  a = warp_sum(x); b = warp_sum(y);
}

struct DconvRowParams {
  // per depth d: conv3 weight [H][C][3], bias [H], gn1 w/b [H], expand weight [2C][H], bias [2C], gn2 w/b [2C], scale [C]
  const float* w1[2]; const float* b1[2]; const float* g1w[2]; const float* g1b[2];
  const float* w2[2]; const float* b2[2]; const float* g2w[2]; const float* g2b[2]; const float* scale[2];
};

template <typename T> struct Pair2;
template <> struct Pair2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *(const float2*)p; }
  static __device__ __forceinline__ void st(float* p, float2 v) { *(float2*)p = v; }
};
template <> struct Pair2<bf16> {
  static __device__ __forceinline__ float2 ld(const bf16* p) { return __bfloat1622float2(*(const __nv_bfloat162*)p); }
  static __device__ __forceinline__ void st(bf16* p, float2 v) { *(__nv_bfloat162*)p = __floats2bfloat162_rn(v.x, v.y); }
};

static constexpr int DCONV_THREADS = 288;     // >= 259 frames of a 6 s segment: the k3 conv runs one frame per thread

template <typename T, int C>
__global__ void __launch_bounds__(DCONV_THREADS, (C == 48 ? 3 : 2)) dconv_row_kernel(T* __restrict__ y, RowSpace ys, DconvRowParams P) {
  constexpr int H = C / 8;
  constexpr int XP = C + 2;                       // padded slab pitch (elements): odd word stride -> no bank conflicts
  constexpr int CB = 24 / H;                      // channels per thread in the statistics pass (24 weight registers)
  constexpr int CC = 2;                           // (value, gate) channel pairs per thread in the apply pass
  const int Tn = ys.G2;                           // frames
  extern __shared__ float smem_f[];
  float* w1s = smem_f;                            // [3][C][H]   (hidden channel fastest)
  float* w2s = w1s + H * 3 * C;                   // [2C][H]
  float* vec = w2s + 2 * C * H;                   // b1[H] g1w[H] g1b[H] b2[2C] g2w[2C] g2b[2C] scale[C]
  float* hs = vec + 3 * H + 7 * C;                // [Tn][H]
  float* red = hs + Tn * H;                       // 64 floats
  T* xs = (T*)(red + 64);                         // [Tn][XP]
  const int f = blockIdx.x % ys.R, b = blockIdx.x / ys.R;
  const int tid = threadIdx.x;

  // ---- load the slab: frame t, 8-channel chunk (16 B for bf16, 32 B for fp32)
  for (int i = tid; i < Tn * (C / 8); i += blockDim.x) {
    const int t = i / (C / 8), c = (i % (C / 8)) * 8;
    float v[8];
    VecIO<T, 8>::load(y + ys.row_off(b * Tn + t, f) + c, v);
#pragma unroll
    for (int k = 0; k < 8; k += 2) Pair2<T>::st(xs + t * XP + c + k, make_float2(v[k], v[k + 1]));
  }

  for (int d = 0; d < 2; ++d) {
    const int dil = 1 << d;
    __syncthreads();
    // ---- stage this depth's weights (fp32) in shared memory
    for (int i = tid; i < H * 3 * C; i += blockDim.x) {
      const int j = i % H, r = i / H, c = r % C, k = r / C;       // dst [k][c][j]
      w1s[i] = P.w1[d][(j * C + c) * 3 + k];
    }
    for (int i = tid; i < 2 * C * H; i += blockDim.x) w2s[i] = P.w2[d][i];
    for (int i = tid; i < H; i += blockDim.x) { vec[i] = P.b1[d][i]; vec[H + i] = P.g1w[d][i]; vec[2 * H + i] = P.g1b[d][i]; }
    for (int i = tid; i < 2 * C; i += blockDim.x) {
      vec[3 * H + i] = P.b2[d][i]; vec[3 * H + 2 * C + i] = P.g2w[d][i]; vec[3 * H + 4 * C + i] = P.g2b[d][i];
    }
    for (int i = tid; i < C; i += blockDim.x) vec[3 * H + 6 * C + i] = P.scale[d][i];
    __syncthreads();
    const float* b1 = vec, *g1w = vec + H, *g1b = vec + 2 * H, *b2 = vec + 3 * H, *g2w = b2 + 2 * C, *g2b = b2 + 4 * C,
               *scl = b2 + 6 * C;

    // ---- h = conv_k3_dil(x) + b1 : one frame per thread, all H hidden channels in registers
    float s1 = 0.f, q1 = 0.f;
    for (int t = tid; t < Tn; t += blockDim.x) {
      float acc[H];
#pragma unroll
      for (int j = 0; j < H; ++j) acc[j] = b1[j];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int tt = t + (k - 1) * dil;
        if (tt < 0 || tt >= Tn) continue;
        const T* xr = xs + tt * XP;
        const float4* wp = (const float4*)(w1s + k * C * H);
#pragma unroll 4
        for (int c = 0; c < C; c += 2) {
          const float2 xv = Pair2<T>::ld(xr + c);
          float w[2 * H];
#pragma unroll
          for (int g = 0; g < 2 * H / 4; ++g) {
            const float4 t4 = wp[(c * H) / 4 + g];
            w[4 * g] = t4.x; w[4 * g + 1] = t4.y; w[4 * g + 2] = t4.z; w[4 * g + 3] = t4.w;
          }
#pragma unroll
          for (int j = 0; j < H; ++j) { acc[j] = fmaf(w[j], xv.x, acc[j]); acc[j] = fmaf(w[H + j], xv.y, acc[j]); }
        }
      }
#pragma unroll
      for (int j = 0; j < H; ++j) { hs[t * H + j] = acc[j]; s1 += acc[j]; q1 += acc[j] * acc[j]; }
    }
    block_sum2(s1, q1, red);
    {
      const float n = (float)(Tn * H);
      const float mean = s1 / n;
      const float rstd = rsqrtf(fmaxf(q1 / n - mean * mean, 0.f) + 1e-5f);
      __syncthreads();
      for (int i = tid; i < Tn * H; i += blockDim.x) {
        const int j = i % H;
        hs[i] = gelu_act<T>((hs[i] - mean) * rstd * g1w[j] + g1b[j]);
      }
    }
    __syncthreads();

    // ---- e = W2 h + b2: statistics only (2C x T values), never stored.  Thread = CB output channels (weights in
    //      registers) x a strided set of frames.
    float s2 = 0.f, q2 = 0.f;
    {
      constexpr int NG = 2 * C / CB;
      const int nslot = blockDim.x / NG;
      const int cg = tid % NG, slot = tid / NG;
      if (slot < nslot) {
        float w[CB][H], bb[CB];
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          bb[cc] = b2[cg * CB + cc];
#pragma unroll
          for (int j = 0; j < H; ++j) w[cc][j] = w2s[(cg * CB + cc) * H + j];
        }
        for (int t = slot; t < Tn; t += nslot) {
          float hv[H];
#pragma unroll
          for (int j = 0; j < H; j += 2) { const float2 h2 = *(const float2*)(hs + t * H + j); hv[j] = h2.x; hv[j + 1] = h2.y; }
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            float e = bb[cc];
#pragma unroll
            for (int j = 0; j < H; ++j) e = fmaf(w[cc][j], hv[j], e);
            s2 += e; q2 = fmaf(e, e, q2);
          }
        }
      }
    }
    block_sum2(s2, q2, red);
    const float n2 = (float)(Tn * 2 * C);
    const float mean2 = s2 / n2;
    const float rstd2 = rsqrtf(fmaxf(q2 / n2 - mean2 * mean2, 0.f) + 1e-5f);

    // ---- x += scale * GN2(e)[c] * sigmoid(GN2(e)[c + C])      (e recomputed; CC (value, gate) pairs per thread)
    {
      constexpr int NG = C / CC;
      const int nslot = blockDim.x / NG;
      const int cg = tid % NG, slot = tid / NG;
      if (slot < nslot) {
        const int c0 = cg * CC;
        float wa[CC][H], wg[CC][H], ba[CC], bg[CC], ga[CC], gb_[CC], gg[CC], gh[CC], sc[CC];
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          const int c = c0 + cc;
          // fold GroupNorm's affine into the recomputed dot product: GN(e) = (e - mean) * rstd * w + b
          ga[cc] = rstd2 * g2w[c]; gb_[cc] = g2b[c] - mean2 * rstd2 * g2w[c];
          gg[cc] = rstd2 * g2w[c + C]; gh[cc] = g2b[c + C] - mean2 * rstd2 * g2w[c + C];
          ba[cc] = b2[c]; bg[cc] = b2[c + C]; sc[cc] = scl[c];
#pragma unroll
          for (int j = 0; j < H; ++j) { wa[cc][j] = w2s[c * H + j]; wg[cc][j] = w2s[(c + C) * H + j]; }
        }
        for (int t = slot; t < Tn; t += nslot) {
          float hv[H];
#pragma unroll
          for (int j = 0; j < H; j += 2) { const float2 h2 = *(const float2*)(hs + t * H + j); hv[j] = h2.x; hv[j + 1] = h2.y; }
          float upd[CC];
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            float ea = ba[cc], eg = bg[cc];
#pragma unroll
            for (int j = 0; j < H; ++j) { ea = fmaf(wa[cc][j], hv[j], ea); eg = fmaf(wg[cc][j], hv[j], eg); }
            ea = fmaf(ea, ga[cc], gb_[cc]);
            eg = fmaf(eg, gg[cc], gh[cc]);
            const float sg = sizeof(T) == 2 ? __frcp_rn(1.0f + __expf(-eg)) : sigmoid_acc(eg);
            upd[cc] = sc[cc] * (ea * sg);
          }
#pragma unroll
          for (int cc = 0; cc < CC; cc += 2) {
            T* xp = xs + t * XP + c0 + cc;
            float2 xv = Pair2<T>::ld(xp);
            xv.x += upd[cc]; xv.y += upd[cc + 1];
            Pair2<T>::st(xp, xv);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < Tn * (C / 8); i += blockDim.x) {
    const int t = i / (C / 8), c = (i % (C / 8)) * 8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k += 2) { const float2 p2 = Pair2<T>::ld(xs + t * XP + c + k); v[k] = p2.x; v[k + 1] = p2.y; }
    VecIO<T, 8>::store(y + ys.row_off(b * Tn + t, f) + c, v);
  }
}

template <typename T, int C>
static size_t dconv_row_smem(int Tn) {
  constexpr int H = C / 8;
  return sizeof(float) * (H * 3 * C + 2 * C * H + 3 * H + 7 * C + (size_t)Tn * H + 64) + sizeof(T) * (size_t)Tn * (C + 2) + 16;
}

template <typename T>
bool dconv_row_supported(int C, int Tn) {
  if (C == 48) return dconv_row_smem<T, 48>(Tn) <= 200 * 1024;
  if (C == 96) return dconv_row_smem<T, 96>(Tn) <= 200 * 1024;
  return false;
}

template <typename T>
void launch_dconv_row(T* y, RowSpace ys, const float* const* ptrs, cudaStream_t st) {
  DconvRowParams P;
  for (int d = 0; d < 2; ++d) {
    const float* const* q = ptrs + 9 * d;
    P.w1[d] = q[0]; P.b1[d] = q[1]; P.g1w[d] = q[2]; P.g1b[d] = q[3]; P.w2[d] = q[4]; P.b2[d] = q[5]; P.g2w[d] = q[6];
    P.g2b[d] = q[7]; P.scale[d] = q[8];
  }
  const int Tn = ys.G2;
  const int blocks = ys.batch() * ys.R;
  if (ys.C == 48) {
    const size_t smem = dconv_row_smem<T, 48>(Tn);
    cudaFuncSetAttribute(dconv_row_kernel<T, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dconv_row_kernel<T, 48><<<blocks, DCONV_THREADS, smem, st>>>(y, ys, P);
  } else {
    const size_t smem = dconv_row_smem<T, 96>(Tn);
    cudaFuncSetAttribute(dconv_row_kernel<T, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dconv_row_kernel<T, 96><<<blocks, DCONV_THREADS, smem, st>>>(y, ys, P);
  }
}

template bool dconv_row_supported<float>(int, int);
template bool dconv_row_supported<bf16>(int, int);
template void launch_dconv_row<float>(float*, RowSpace, const float* const*, cudaStream_t);
template void launch_dconv_row<bf16>(bf16*, RowSpace, const float* const*, cudaStream_t);

}  // namespace athtd
