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

template <typename T, int C>
__global__ void __launch_bounds__(256) dconv_row_kernel(T* __restrict__ y, RowSpace ys, DconvRowParams P) {
  constexpr int H = C / 8;
  constexpr int XP = C + 2;                       // padded slab pitch (elements): odd word stride -> no bank conflicts
  const int Tn = ys.G2;                           // frames
  extern __shared__ float smem_f[];
  float* w1s = smem_f;                            // [H][3][C]   (tap-major inside a row)
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
    for (int k = 0; k < 8; ++k) xs[t * XP + c + k] = from_f<T>(v[k]);
  }

  for (int d = 0; d < 2; ++d) {
    const int dil = 1 << d;
    __syncthreads();
    // ---- stage this depth's weights (fp32) in shared memory
    for (int i = tid; i < H * 3 * C; i += blockDim.x) {
      const int j = i / (3 * C), r = i % (3 * C), k = r / C, c = r % C;
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

    // ---- h = conv_k3_dil(x) + b1 ; partial sums for GroupNorm(1, H)
    float s1 = 0.f, q1 = 0.f;
    for (int i = tid; i < Tn * (H / 3); i += blockDim.x) {          // 3 hidden channels per item
      const int t = i % Tn, j0 = (i / Tn) * 3;
      float a0 = b1[j0], a1 = b1[j0 + 1], a2 = b1[j0 + 2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int tt = t + (k - 1) * dil;
        if (tt < 0 || tt >= Tn) continue;
        const T* xr = xs + tt * XP;
        const float* wa = w1s + (j0 * 3 + k) * C;
#pragma unroll 8
        for (int c = 0; c < C; ++c) {
          const float xv = to_f<T>(xr[c]);
          a0 = fmaf(wa[c], xv, a0); a1 = fmaf(wa[3 * C + c], xv, a1); a2 = fmaf(wa[6 * C + c], xv, a2);
        }
      }
      hs[t * H + j0] = a0; hs[t * H + j0 + 1] = a1; hs[t * H + j0 + 2] = a2;
      s1 += a0 + a1 + a2; q1 += a0 * a0 + a1 * a1 + a2 * a2;
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

    // ---- e = W2 h + b2: statistics only (2C x T values), never stored
    float s2 = 0.f, q2 = 0.f;
    for (int i = tid; i < Tn * (C / 4); i += blockDim.x) {           // 8 of the 2C channels per item
      const int t = i % Tn, c0 = (i / Tn) * 8;
      float hv[H];
#pragma unroll
      for (int j = 0; j < H; ++j) hv[j] = hs[t * H + j];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        float e = b2[c0 + cc];
#pragma unroll
        for (int j = 0; j < H; ++j) e = fmaf(w2s[(c0 + cc) * H + j], hv[j], e);
        s2 += e; q2 += e * e;
      }
    }
    block_sum2(s2, q2, red);
    const float n2 = (float)(Tn * 2 * C);
    const float mean2 = s2 / n2;
    const float rstd2 = rsqrtf(fmaxf(q2 / n2 - mean2 * mean2, 0.f) + 1e-5f);

    // ---- x += scale * GN2(e)[c] * sigmoid(GN2(e)[c + C])      (e recomputed)
    for (int i = tid; i < Tn * (C / 4); i += blockDim.x) {           // 4 channels per item
      const int t = i % Tn, c0 = (i / Tn) * 4;
      float hv[H];
#pragma unroll
      for (int j = 0; j < H; ++j) hv[j] = hs[t * H + j];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = c0 + cc;
        float ea = b2[c], eg = b2[c + C];
#pragma unroll
        for (int j = 0; j < H; ++j) { ea = fmaf(w2s[c * H + j], hv[j], ea); eg = fmaf(w2s[(c + C) * H + j], hv[j], eg); }
        ea = (ea - mean2) * rstd2 * g2w[c] + g2b[c];
        eg = (eg - mean2) * rstd2 * g2w[c + C] + g2b[c + C];
        const float sg = sizeof(T) == 2 ? __frcp_rn(1.0f + __expf(-eg)) : sigmoid_acc(eg);
        xs[t * XP + c] = from_f<T>(to_f<T>(xs[t * XP + c]) + scl[c] * (ea * sg));
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < Tn * (C / 8); i += blockDim.x) {
    const int t = i / (C / 8), c = (i % (C / 8)) * 8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = to_f<T>(xs[t * XP + c + k]);
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
    dconv_row_kernel<T, 48><<<blocks, 256, smem, st>>>(y, ys, P);
  } else {
    const size_t smem = dconv_row_smem<T, 96>(Tn);
    cudaFuncSetAttribute(dconv_row_kernel<T, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dconv_row_kernel<T, 96><<<blocks, 256, smem, st>>>(y, ys, P);
  }
}

template bool dconv_row_supported<float>(int, int);
template bool dconv_row_supported<bf16>(int, int);
template void launch_dconv_row<float>(float*, RowSpace, const float* const*, cudaStream_t);
template void launch_dconv_row<bf16>(bf16*, RowSpace, const float* const*, cudaStream_t);

}  // namespace athtd
