// Bandwidth-bound kernels of the path: input normalisation, GroupNorm/LayerNorm applies,
// GLU/LayerScale residuals, softmax, decoder resize+skip, weight packing.
// All activation tensors are channels-last inside padded row spaces (common.cuh: RowSpace).
#include "kernels.cuh"
#include <algorithm>

namespace athtd {

// measured on B200 inside the bench step (same gpurun call, two repetitions each): 27.44 / 27.50 ms with programmatic dependent
// launch, 26.31 / 26.78 ms without -- the early-resident CTAs of the next kernel cost more than the overlapped prologues save.
// Off by default; athtd_set_pdl(1) turns it on.
static bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void pdl_set_enabled(bool on) { g_pdl = on; }

// ------------------------------------------------------------------ per-sample sum / sumsq
__global__ void sum_sumsq_kernel(const float* __restrict__ x, long n_per_sample, double* __restrict__ stats) {
  pdl_begin();
  const int b = blockIdx.y;
  const float* p = x + (long)b * n_per_sample;
  float s = 0.f, ss = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_sample; i += (long)gridDim.x * blockDim.x) {
    float v = p[i]; s += v; ss += v * v;
  }
  __shared__ float sh[2][32];
  s = warp_sum(s); ss = warp_sum(ss);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = ss; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    double a = l < nw ? (double)sh[0][l] : 0.0, c = l < nw ? (double)sh[1][l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (l == 0) { atomicAdd(stats + 2 * b, a); atomicAdd(stats + 2 * b + 1, c); }
  }
}
void launch_sum_sumsq(const float* x, int B, long n_per_sample, double* stats, cudaStream_t st) {
  int blocks = (int)min((n_per_sample + 256 * 8 - 1) / (256 * 8), (long)296);
  if (blocks < 1) blocks = 1;
  launch_pdl(sum_sumsq_kernel, dim3(dim3(blocks, B)), dim3(256), 0, st, x, n_per_sample, stats);
}

// (sum, sumsq) double accumulators -> (mean, rstd) floats, biased variance (GroupNorm semantics)
__global__ void finalize_gn_kernel(const double* __restrict__ stats, double count, float* __restrict__ mr, long n, int nslot) {
  pdl_begin();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double st[2] = {stats[2 * i], stats[2 * i + 1]};
    float m, r; stats_to_mean_rstd(st, count, 1e-5f, m, r); mr[2 * i] = m; mr[2 * i + 1] = r;
  }
}
// slotted accumulators (nslot > 1): one warp per group, lanes over slots, fixed-order tree (a single thread walking the
// 64 slots was a 10 us dependent-load chain, 16 times per forward)
__global__ void finalize_gn_slots_kernel(const double* __restrict__ stats, double count, float* __restrict__ mr, long n, int nslot) {
  pdl_begin();
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  double st[2] = {0.0, 0.0};
  for (int s = lane; s < nslot; s += 32) { st[0] += stats[2 * (i * nslot + s)]; st[1] += stats[2 * (i * nslot + s) + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { st[0] += __shfl_xor_sync(0xffffffffu, st[0], o); st[1] += __shfl_xor_sync(0xffffffffu, st[1], o); }
  if (lane == 0) { float m, r; stats_to_mean_rstd(st, count, 1e-5f, m, r); mr[2 * i] = m; mr[2 * i + 1] = r; }
}
void launch_finalize_gn(const double* stats, double count, float* mr, long n, int nslot, cudaStream_t st) {
  if (nslot == 1) launch_pdl(finalize_gn_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, st, stats, count, mr, n, nslot);
  else launch_pdl(finalize_gn_slots_kernel, dim3((unsigned)((n * 32 + 127) / 128)), dim3(128), 0, st, stats, count, mr, n, nslot);
}

// torch.std default (unbiased) and the reference guard (x - mean) / (1e-5 + std)   ATHTDemucs_v2.py:268-275
__device__ __forceinline__ void unbiased_mean_std(const double* st, double n, float& mean, float& std_) {
  double m = st[0] / n;
  double var = (st[1] - st[0] * m) / (n - 1.0);
  if (var < 0.0) var = 0.0;
  mean = (float)m; std_ = (float)sqrt(var);
}

__global__ void finalize_meanstd_kernel(const double* __restrict__ stats, double n, float* __restrict__ out, int B) {
  pdl_begin();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { float m, s; unbiased_mean_std(stats + 2 * b, n, m, s); out[2 * b] = m; out[2 * b + 1] = s; }
}
void launch_finalize_meanstd(const double* stats, double n, float* out, int B, cudaStream_t st) {
  launch_pdl(finalize_meanstd_kernel, dim3((B + 127) / 128), dim3(128), 0, st, stats, n, out, B);
}

__global__ void set_meanstd_kernel(float* __restrict__ out, int B, float mean, float stdv) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { out[2 * b] = mean; out[2 * b + 1] = stdv; }
}
void launch_set_meanstd(float* out, int B, float mean, float stdv, cudaStream_t st) {
  set_meanstd_kernel<<<(B + 127) / 128, 128, 0, st>>>(out, B, mean, stdv);
}

// wav [B,2,L] fp32 -> normalised channels-last padded [B, Rp, 2]
template <typename T>
__global__ void pack_wav_kernel(const float* __restrict__ wav, const float* __restrict__ meanstd, T* __restrict__ out,
                                RowSpace rs, int L) {
  int b = blockIdx.y;
  float mean = meanstd[2 * b], inv = 1.0f / (1e-5f + meanstd[2 * b + 1]);
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
    float a = (wav[((long)b * 2 + 0) * L + l] - mean) * inv;
    float c = (wav[((long)b * 2 + 1) * L + l] - mean) * inv;
    long o = rs.row_off(b, l);
    out[o] = from_f<T>(a); out[o + 1] = from_f<T>(c);
  }
}
template <typename T>
void launch_pack_wav(const float* wav, const float* meanstd, T* out, RowSpace rs, int L, cudaStream_t st) {
  pack_wav_kernel<T><<<dim3(min((L + 255) / 256, 1024), rs.G), 256, 0, st>>>(wav, meanstd, out, rs, L);
}

// Z [B, Tf, 2048, 4] fp32 -> normalised padded [B*Tf, Rp, 4]
template <typename T>
__global__ void pack_spec_kernel(const float4* __restrict__ Z, const float* __restrict__ meanstd, T* __restrict__ out,
                                 RowSpace rs, int Tf) {
  int g = blockIdx.y;           // b*Tf + t
  int b = g / Tf;
  float mean = meanstd[2 * b], inv = 1.0f / (1e-5f + meanstd[2 * b + 1]);
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < rs.R; f += gridDim.x * blockDim.x) {
    float4 v = Z[(long)g * rs.R + f];
    long o = rs.row_off(g, f);
    out[o] = from_f<T>((v.x - mean) * inv); out[o + 1] = from_f<T>((v.y - mean) * inv);
    out[o + 2] = from_f<T>((v.z - mean) * inv); out[o + 3] = from_f<T>((v.w - mean) * inv);
  }
}
template <typename T>
void launch_pack_spec(const float* Z, const float* meanstd, T* out, RowSpace rs, int Tf, cudaStream_t st) {
  pack_spec_kernel<T><<<dim3((rs.R + 255) / 256, rs.G), 256, 0, st>>>((const float4*)Z, meanstd, out, rs, Tf);
}

// ------------------------------------------------------------------ GroupNorm(1,C) + GELU, in place
// stat index: per_row ? (g/G2)*R + r : g/G2 ; count given by the caller.
template <typename T, int VEC>
__global__ void gn_gelu_kernel(T* __restrict__ h, RowSpace rs, int G2, int per_row, const float* __restrict__ mr,
                               const float* __restrict__ w, const float* __restrict__ bvec) {
  const int CV = rs.C / VEC;
  long total = (long)rs.G * rs.R * CV;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % CV) * VEC; long row = i / CV;
    int r = (int)(row % rs.R); int g = (int)(row / rs.R);
    long si = per_row ? (long)(g / G2) * rs.R + r : (long)(g / G2);
    float mean = mr[2 * si], rstd = mr[2 * si + 1];
    long o = rs.row_off(g, r) + c;
    float v[VEC];
    VecIO<T, VEC>::load(h + o, v);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = gelu_act<T>((v[k] - mean) * rstd * w[c + k] + bvec[c + k]);
    VecIO<T, VEC>::store(h + o, v);
  }
}
template <typename T>
__global__ void gn_gelu_scalar_kernel(T* __restrict__ h, RowSpace rs, int G2, int per_row, const float* __restrict__ mr,
                                      const float* __restrict__ w, const float* __restrict__ bvec) {
  long total = (long)rs.G * rs.R * rs.C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % rs.C); long row = i / rs.C;
    int r = (int)(row % rs.R); int g = (int)(row / rs.R);
    long si = per_row ? (long)(g / G2) * rs.R + r : (long)(g / G2);
    float mean = mr[2 * si], rstd = mr[2 * si + 1];
    long o = rs.row_off(g, r) + c;
    float v = (to_f<T>(h[o]) - mean) * rstd * w[c] + bvec[c];
    h[o] = from_f<T>(gelu_act<T>(v));
  }
}
template <typename T>
void launch_gn_gelu(T* h, RowSpace rs, int G2, int per_row, const float* mr, const float* w,
                    const float* b, cudaStream_t st) {
  if (rs.C % 8 == 0) {
    long total = (long)rs.G * rs.R * (rs.C / 8);
    gn_gelu_kernel<T, 8><<<(int)min((total + 255) / 256, (long)148 * 32), 256, 0, st>>>(h, rs, G2, per_row, mr, w, b);
  } else {
    long total = (long)rs.G * rs.R * rs.C;
    gn_gelu_scalar_kernel<T><<<(int)min((total + 255) / 256, (long)148 * 32), 256, 0, st>>>(h, rs, G2, per_row, mr, w, b);
  }
}

// x <- x + scale[c] * ( GN(e)[c] * sigmoid(GN(e)[c+C]) )     DConv tail (demucs DConv layers 4..6)
template <typename T>
__global__ void gn_glu_res_kernel(T* __restrict__ x, RowSpace xs, const T* __restrict__ e, RowSpace es, int G2, int per_row,
                                  const float* __restrict__ mr, const float* __restrict__ w,
                                  const float* __restrict__ bvec, const float* __restrict__ scale) {
  const int C = xs.C;
  long total = (long)xs.G * xs.R * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long row = i / C;
    int r = (int)(row % xs.R); int g = (int)(row / xs.R);
    long si = per_row ? (long)(g / G2) * xs.R + r : (long)(g / G2);
    float mean = mr[2 * si], rstd = mr[2 * si + 1];
    long eo = es.row_off(g, r);
    float a = (to_f<T>(e[eo + c]) - mean) * rstd * w[c] + bvec[c];
    float gt = (to_f<T>(e[eo + c + C]) - mean) * rstd * w[c + C] + bvec[c + C];
    long o = xs.row_off(g, r) + c;
    x[o] = from_f<T>(to_f<T>(x[o]) + scale[c] * (a * sigmoid_acc(gt)));
  }
}
template <typename T>
void launch_gn_glu_res(T* x, RowSpace xs, const T* e, RowSpace es, int G2, int per_row, const float* mr,
                       const float* w, const float* b, const float* scale, cudaStream_t st) {
  long total = (long)xs.G * xs.R * xs.C;
  int blocks = (int)min((total + 255) / 256, (long)148 * 16);
  gn_glu_res_kernel<T><<<blocks, 256, 0, st>>>(x, xs, e, es, G2, per_row, mr, w, b, scale);
}

// ------------------------------------------------------------------ (GroupNorm apply) + LayerNorm (+pos-emb)
// One warp per token row of C <= 512 channels.  x: [rows, C] contiguous.
//   if gstats: x' = (x-mu_b)*rstd_b*gw + gb  is written to xout (MyGroupNorm over all tokens of a sample)
//   if lw:     y  = LN(x')*lw + lb (+ pe[row % S])   is written to y
template <typename T, int RPW>
__global__ void __launch_bounds__(256, 3) norm_rows_kernel(const T* __restrict__ x, T* __restrict__ xout, T* __restrict__ y, long rows, int C, int S,
                                 const float* __restrict__ gmr, const float* __restrict__ gw,
                                 const float* __restrict__ gb, const float* __restrict__ lw, const float* __restrict__ lb,
                                 const float* __restrict__ pe, RowSpace yrs, T* __restrict__ y2,
                                 const float* __restrict__ lw2, const float* __restrict__ lb2) {
  pdl_begin();
  // one warp per RPW consecutive rows; each lane owns up to two 8-channel chunks (C <= 512, C % 8 == 0): 16-byte loads /
  // stores, all RPW rows' loads in flight together, and every affine vector is fetched once per RPW rows (fetched per row,
  // the fp32 parameters were 4-6x the bf16 payload in L1 traffic)
  const int lane = threadIdx.x & 31;
  const long row0 = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (row0 >= rows) return;
  const int nchunk = C >> 3;
  float v[RPW][16];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long row = min(row0 + r, rows - 1);          // tail rows recompute the last row and are not stored
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (lane + 32 * i < nchunk) VecIO<T, 8>::load(x + row * C + (lane + 32 * i) * 8, v[r] + 8 * i);
  }
  if (gmr) {
    // (mean, rstd) of the row's segment: one 32-bit division per row (rows < 2^31), not a 64-bit one per row and chunk
    float gm_[RPW], gr_[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int seg = (int)min(row0 + r, rows - 1) / S;
      gm_[r] = gmr[2 * seg]; gr_[r] = gmr[2 * seg + 1];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (lane + 32 * i < nchunk) {
        float w8[8], b8[8];
        VecIO<float, 8>::load(gw + c, w8); VecIO<float, 8>::load(gb + c, b8);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const long row = row0 + r;
          const float gm = gm_[r], gr = gr_[r];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[r][8 * i + k] = (v[r][8 * i + k] - gm) * gr * w8[k] + b8[k];
          if (row < rows) VecIO<T, 8>::store(xout + row * C + c, v[r] + 8 * i);
        }
      }
    }
  }
  if (!lw) return;
  float mean[RPW], rstd[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (lane + 32 * i < nchunk) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[r][8 * i + k];
      }
    mean[r] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    mean[r] = mean[r] / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (lane + 32 * i < nchunk) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { float dlt = v[r][8 * i + k] - mean[r]; q += dlt * dlt; }
      }
    rstd[r] = q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) rstd[r] = rsqrtf(rstd[r] / C + 1e-5f);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ch = lane + 32 * i;
    if (ch < nchunk) {
      const int c = ch * 8;
      float w8[8], b8[8];
      VecIO<float, 8>::load(lw + c, w8); VecIO<float, 8>::load(lb + c, b8);
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const long row = row0 + r;
        if (row < rows) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[r][8 * i + k] - mean[r]) * rstd[r] * w8[k] + b8[k];
          if (pe) {
            float p8[8];
            VecIO<float, 8>::load(pe + (long)((int)row % S) * C + c, p8);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += p8[k];
          }
          long yoff = row * C;
          if (yrs.C > 0) yoff = yrs.row_off((int)row / yrs.R, (int)row % yrs.R);   // scatter into a padded row space (rows < 2^31)
          VecIO<T, 8>::store(y + yoff + c, o);
        }
      }
      if (y2) {      // a second LayerNorm of the same rows (same statistics, other affine): the next layer's other consumer
        VecIO<float, 8>::load(lw2 + c, w8); VecIO<float, 8>::load(lb2 + c, b8);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const long row = row0 + r;
          if (row < rows) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (v[r][8 * i + k] - mean[r]) * rstd[r] * w8[k] + b8[k];
            VecIO<T, 8>::store(y2 + row * C + c, o);
          }
        }
      }
    }
  }
}
// LayerNorm-only form for C == 512 (every transformer norm): the affine vectors of a lane's two fixed 8-channel chunks live in
// registers and the warp walks rows grid-stride with the NEXT TWO rows (2 x 2 x 16 B, still packed) already in flight.  The
// generic kernel re-read 4 KB of fp32 affine parameters from L1 per 1 KB row (4x the payload) and ran at ~2.5 TB/s; the
// one-row-ahead version of this kernel was launched with twice the CTAs its ~90 registers let an SM hold (two waves of a
// grid-stride loop) and had 1 KB per warp in flight (ncu: 23 % of the warp slots active, long-scoreboard bound, 2.1 TB/s).
// Now: one resident wave (2 CTAs of 8 warps per SM) and 2 KB per warp in flight.  Same operation order per row as the generic
// kernel (bit-identical results).
template <typename T, bool HAS2>
__global__ void __launch_bounds__(256, 2) ln512_rows_kernel(const T* __restrict__ x, T* __restrict__ y, long rows, int S,
                                                            const float* __restrict__ lw, const float* __restrict__ lb,
                                                            const float* __restrict__ pe, RowSpace yrs, T* __restrict__ y2,
                                                            const float* __restrict__ lw2, const float* __restrict__ lb2) {
  pdl_begin();
  constexpr int C = 512;
  if constexpr (sizeof(T) != 2) return;          // (instantiated for the fp32 build's template, never launched for it)
  else {
  const int lane = threadIdx.x & 31;
  const long nw = (long)gridDim.x * (blockDim.x >> 5);
  long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float w[16], b[16], w2[HAS2 ? 16 : 1], b2[HAS2 ? 16 : 1];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (lane + 32 * i) * 8;
    VecIO<float, 8>::load(lw + c, w + 8 * i); VecIO<float, 8>::load(lb + c, b + 8 * i);
    if (HAS2) { VecIO<float, 8>::load(lw2 + c, w2 + 8 * i); VecIO<float, 8>::load(lb2 + c, b2 + 8 * i); }
  }
  auto ldrow = [&](long r, uint4& a, uint4& c) {
    const uint4* p = (const uint4*)(x + r * C);
    a = p[lane]; c = p[lane + 32];
  };
  auto unpack = [&](const uint4& a, float* v) {
    const float2 f0 = __bfloat1622float2(*(const __nv_bfloat162*)&a.x), f1 = __bfloat1622float2(*(const __nv_bfloat162*)&a.y);
    const float2 f2 = __bfloat1622float2(*(const __nv_bfloat162*)&a.z), f3 = __bfloat1622float2(*(const __nv_bfloat162*)&a.w);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
  };
  uint4 c0a, c0b, n1a = make_uint4(0, 0, 0, 0), n1b = n1a;
  ldrow(row, c0a, c0b);
  if (row + nw < rows) ldrow(row + nw, n1a, n1b);
  for (; row < rows; row += nw) {
    uint4 n2a = make_uint4(0, 0, 0, 0), n2b = n2a;
    if (row + 2 * nw < rows) ldrow(row + 2 * nw, n2a, n2b);
    float v[16];
    unpack(c0a, v); unpack(c0b, v + 8);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    s = warp_sum(s);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { float dlt = v[k] - mean; q += dlt * dlt; }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / C + 1e-5f);
    long yoff = row * C;
    if (yrs.C > 0) yoff = yrs.row_off((int)row / yrs.R, (int)row % yrs.R);      // (rows < 2^31: 32-bit divisions)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = (lane + 32 * i) * 8;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (v[8 * i + k] - mean) * rstd * w[8 * i + k] + b[8 * i + k];
      if (pe) {
        float p8[8];
        VecIO<float, 8>::load(pe + (long)((int)row % S) * C + c, p8);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += p8[k];
      }
      VecIO<T, 8>::store(y + yoff + c, o);
      if (HAS2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[8 * i + k] - mean) * rstd * w2[8 * i + k] + b2[8 * i + k];
        VecIO<T, 8>::store(y2 + row * C + c, o);
      }
    }
    c0a = n1a; c0b = n1b; n1a = n2a; n1b = n2b;
  }
  }
}

template <typename T>
void launch_norm_rows(const T* x, T* xout, T* y, long rows, int C, int S, const float* gmr,
                      const float* gw, const float* gb, const float* lw, const float* lb, const float* pe,
                      RowSpace yrs, cudaStream_t st, T* y2, const float* lw2, const float* lb2) {
  if (C == 512 && !gmr && lw && sizeof(T) == 2) {
    const unsigned grid = (unsigned)std::min<long>((rows + 7) / 8, 2L * device_sm_count());      // one resident wave
    if (y2) launch_pdl(ln512_rows_kernel<T, true>, dim3(grid), dim3(256), 0, st, x, y, rows, S, lw, lb, pe, yrs, y2, lw2, lb2);
    else launch_pdl(ln512_rows_kernel<T, false>, dim3(grid), dim3(256), 0, st, x, y, rows, S, lw, lb, pe, yrs, y2, lw2, lb2);
    return;
  }
  constexpr int wpb = 8;
  if (C == 512)      // (measured: two rows per warp help the 512-channel GroupNorm + LayerNorm passes, not the narrower ones)
    launch_pdl(norm_rows_kernel<T, 2>, dim3((unsigned)((rows + wpb * 2 - 1) / (wpb * 2))), dim3(wpb * 32), 0, st, x, xout, y, rows, C, S, gmr, gw, gb, lw,
                                                                                              lb, pe, yrs, y2, lw2, lb2);
  else
    launch_pdl(norm_rows_kernel<T, 1>, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), 0, st, x, xout, y, rows, C, S, gmr, gw, gb, lw, lb, pe,
                                                                                    yrs, y2, lw2, lb2);
}

// ------------------------------------------------------------------ row softmax, in place (fp32 math)
template <typename T>
__global__ void softmax_rows_kernel(T* __restrict__ s, long rows, int n) {
  long row = blockIdx.x;
  T* p = s + row * n;
  __shared__ float sh[32];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, to_f<T>(p[i]));
  mx = warp_max(mx);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sh[w] = mx;
  __syncthreads();
  mx = l < nw ? sh[l] : -INFINITY; mx = warp_max(mx);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf(to_f<T>(p[i]) - mx);
  sum = warp_sum(sum);
  if (l == 0) sh[w] = sum;
  __syncthreads();
  sum = l < nw ? sh[l] : 0.f; sum = warp_sum(sum);
  float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = from_f<T>(expf(to_f<T>(p[i]) - mx) * inv);
}
template <typename T>
void launch_softmax_rows(T* s, long rows, int n, cudaStream_t st) {
  softmax_rows_kernel<T><<<(unsigned)rows, 256, 0, st>>>(s, rows, n);
}

// ------------------------------------------------------------------ y[b, r, :] = x[b, r, :] + vec[b*vstride, :]
template <typename T>
__global__ void add_rowvec_kernel(const T* __restrict__ x, T* __restrict__ y, long rows_per_b, int C, int B,
                                  const float* __restrict__ vec, long vstride) {
  pdl_begin();
  long total = (long)B * rows_per_b * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long b = i / ((long)rows_per_b * C);
    y[i] = from_f<T>(to_f<T>(x[i]) + vec[b * vstride + c]);
  }
}
// 16-byte form (C % VEC == 0, 16-byte aligned buffers): one vector of one row per thread and iteration
template <typename T>
__global__ void add_rowvec_vec_kernel(const T* __restrict__ x, T* __restrict__ y, long rows_per_b, int C, int B,
                                      const float* __restrict__ vec, long vstride) {
  pdl_begin();
  constexpr int VEC = 16 / (int)sizeof(T);
  const int cv = C / VEC;
  const int total = B * (int)rows_per_b * cv;      // (< 2^31 vectors, checked by the launcher: 32-bit index arithmetic)
  const int per_b = (int)rows_per_b * cv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / per_b;
    const int c = (i % cv) * VEC;
    uint4 raw = *(const uint4*)(x + (long)i * VEC);
    T* e = (T*)&raw;
    const float* vp = vec + (long)b * vstride + c;
#pragma unroll
    for (int k = 0; k < VEC; ++k) e[k] = from_f<T>(to_f<T>(e[k]) + __ldg(vp + k));
    *(uint4*)(y + (long)i * VEC) = raw;
  }
}
template <typename T>
void launch_add_rowvec(const T* x, T* y, long rows_per_b, int C, int B, const float* vec, long vstride, cudaStream_t st) {
  constexpr int VEC = 16 / (int)sizeof(T);
  if (C % VEC == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && (long)B * rows_per_b * (C / VEC) < (1L << 31)) {
    long total = (long)B * rows_per_b * (C / VEC);
    launch_pdl(add_rowvec_vec_kernel<T>, dim3((int)min((total + 255) / 256, (long)148 * 16)), dim3(256), 0, st, x, y, rows_per_b, C, B, vec, vstride);
    return;
  }
  long total = (long)B * rows_per_b * C;
  launch_pdl(add_rowvec_kernel<T>, dim3((int)min((total + 255) / 256, (long)148 * 16)), dim3(256), 0, st, x, y, rows_per_b, C, B, vec, vstride);
}

// ------------------------------------------------------------------ text conditioning vector
// cvec[row] = out_proj(in_proj_v(v_proj(emb[row]))): three chained fp32 linears of one 512-vector (TextCrossAttention with a
// single key: softmax == 1, ATHTDemucs_v2.py:38-58).  One block per (segment, prompt) row, a warp per
// output feature (coalesced float4 weight reads, shuffle reduction), the intermediates stay in shared memory: three
// latency-bound M = B*P GEMM launches (~55 us each) become one ~10 us launch.
__device__ __forceinline__ void tv_layer(const float* __restrict__ x, int K, const float* __restrict__ W, const float* __restrict__ b,
                                         float* __restrict__ y, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // four outputs per warp pass: 4 x K/128 independent 16-byte weight loads in flight per lane (the loop is L2-latency bound)
  for (int n0 = 4 * warp; n0 < N; n0 += 4 * nw) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
      const float4 v = *(const float4*)(x + 4 * k4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = __ldg((const float4*)(W + (long)min(n0 + j, N - 1) * K) + k4);
        acc[j] = fmaf(w.x, v.x, acc[j]); acc[j] = fmaf(w.y, v.y, acc[j]); acc[j] = fmaf(w.z, v.z, acc[j]); acc[j] = fmaf(w.w, v.w, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      if (lane == 0 && n0 + j < N) y[n0 + j] = acc[j] + b[n0 + j];
    }
  }
}
__global__ void __launch_bounds__(1024) text_vectors_kernel(const float* __restrict__ emb, const float* __restrict__ W1,
                                                           const float* __restrict__ b1, const float* __restrict__ W2,
                                                           const float* __restrict__ b2, const float* __restrict__ W3,
                                                           const float* __restrict__ b3, float* __restrict__ out) {
  pdl_begin();
  __shared__ __align__(16) float s0[512], s1[384], s2[384];
  const long row = blockIdx.x;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s0[i] = emb[row * 512 + i];
  __syncthreads();
  tv_layer(s0, 512, W1, b1, s1, 384);
  __syncthreads();
  tv_layer(s1, 384, W2, b2, s2, 384);
  __syncthreads();
  tv_layer(s2, 384, W3, b3, out + row * 384, 384);
}
void launch_text_vectors(const float* emb, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                         const float* b3, float* out, int rows, cudaStream_t st) {
  launch_pdl(text_vectors_kernel, dim3(rows), dim3(1024), 0, st, emb, W1, b1, W2, b2, W3, b3, out);
}

// ------------------------------------------------------------------ decoder layer tail
// out[g, d, c] = lerp_rows( act(GN(u))[g, :, c] )(d) + 0.1 * lerp_rows( skip[g, :, c] )(d)
//   u    : transposed-conv output in phase layout inside the INPUT's padded geometry (us, 4*Cu channels per row)
//   GN   : GroupNorm(1,C) statistics per sample (g / G2), biased variance, then exact GELU  (has_gn)
//   lerp : F.interpolate(mode=linear/bilinear along one axis, align_corners=False)   
// Follows FreqDecoder.forward / TimeDecoder.forward, ATHTDemucs_v2.py:82-104 / 125-139.
__device__ __forceinline__ void lerp_coords(int d, int in, int out, int& i0, int& i1, float& lam) {
  if (in == out) { i0 = d; i1 = d; lam = 0.f; return; }
  float scale = (float)in / (float)out;
  float src = scale * ((float)d + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  lam = src - (float)i0;
}
// One block = one group (segment x frame on the frequency branch, segment on the time branch) x TD output rows.
// Thread <-> fixed channel chunk x a contiguous RUN of output rows: GroupNorm scale / shift, the two interpolation
// source rows and the two skip rows live in registers and are reloaded only when the source index moves (the resize
// coordinates of the tile's rows are computed once per block into shared memory).
// STAGE (every input row is used, i.e. up-sampling or ~1:1 resizes): GELU(GN(u)) of the input rows the tile needs is
// evaluated ONCE into shared memory (fp32) and the outputs interpolate from there -- the direct form evaluated two GELUs
// per output element (16x redundant on the 32 -> 259 layer) and was ALU-bound.  The exact 4:1 layers (only phases 0 / 3 of
// the transposed conv are stored) keep the direct form: each stored row is used exactly once.
struct LerpRow { int i0, i1; float lam; int j0, j1; float mu; };
// a[k] = GELU(a[k] * sc[k] + sh[k]); bf16 build: packed pairs
template <typename T, int VEC>
__device__ __forceinline__ void dec_gn_gelu(float (&a)[VEC], const float (&sc)[VEC], const float (&sh)[VEC]) {
  if constexpr (sizeof(T) == 2 && VEC % 2 == 0) {
#pragma unroll
    for (int k = 0; k < VEC; k += 2) {
      const float2 r = gelu_fast2(f2fma(make_float2(a[k], a[k + 1]), make_float2(sc[k], sc[k + 1]), make_float2(sh[k], sh[k + 1])));
      a[k] = r.x; a[k + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) a[k] = gelu_act<T>(fmaf(a[k], sc[k], sh[k]));
  }
}
// v = lerp(a0, a1; lam) + 0.1 * lerp(s0, s1; mu)
template <typename T, int VEC>
__device__ __forceinline__ void dec_blend(float (&v)[VEC], const float (&a0)[VEC], const float (&a1)[VEC], const float (&s0)[VEC],
                                          const float (&s1)[VEC], float lam, float mu) {
  if constexpr (sizeof(T) == 2 && VEC % 2 == 0) {
    const float2 l0 = f2splat(1.f - lam), l1 = f2splat(lam), m0 = f2splat(0.1f * (1.f - mu)), m1 = f2splat(0.1f * mu);
#pragma unroll
    for (int k = 0; k < VEC; k += 2) {
      float2 r = f2mul(l0, make_float2(a0[k], a0[k + 1]));
      r = f2fma(l1, make_float2(a1[k], a1[k + 1]), r);
      r = f2fma(m0, make_float2(s0[k], s0[k + 1]), r);
      r = f2fma(m1, make_float2(s1[k], s1[k + 1]), r);
      v[k] = r.x; v[k + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = ((1.f - lam) * a0[k] + lam * a1[k]) + 0.1f * ((1.f - mu) * s0[k] + mu * s1[k]);
  }
}

template <typename T, int VEC, bool STAGE>
__global__ void __launch_bounds__(256, 4) dec_apply_kernel(const T* __restrict__ u, int Uin, RowSpace us, int Cu, T* __restrict__ out,
                                                        RowSpace os, int G2, int has_gn, const float* __restrict__ mr,
                                                        const float* __restrict__ gw, const float* __restrict__ gb,
                                                        const T* __restrict__ skip, RowSpace ss, int TD) {
  pdl_begin();
  extern __shared__ __align__(16) uint8_t dec_smem[];
  LerpRow* tab = (LerpRow*)dec_smem;                                 // [TD]
  float* act_s = (float*)(dec_smem + (((size_t)TD * sizeof(LerpRow) + 15) & ~(size_t)15));   // STAGE: [n_in][Cu]
  const int g = blockIdx.y;
  const int CV = os.C / VEC;
  const int RW = blockDim.x / CV;
  const int cv = threadIdx.x % CV, rw = threadIdx.x / CV, c = cv * VEC;
  const bool active = rw < RW;
  const int d0 = blockIdx.x * TD, d1 = min(d0 + TD, os.R);
  const T* ug = u + us.row_off(g, 0);
  T* og = out + os.row_off(g, 0);
  const T* sg = skip + ss.row_off(g, 0);
  const long urow = us.C, orow = os.C, srow = ss.C;
  for (int i = threadIdx.x; i < d1 - d0; i += blockDim.x) {
    LerpRow lr;
    lerp_coords(d0 + i, Uin, os.R, lr.i0, lr.i1, lr.lam);
    lerp_coords(d0 + i, ss.R, os.R, lr.j0, lr.j1, lr.mu);
    tab[i] = lr;
  }
  float sc[VEC], sh[VEC];
  if (has_gn) {
    const float mean = mr[2 * (g / G2)], rstd = mr[2 * (g / G2) + 1];
    float w[VEC], bb[VEC];
    VecIO<float, VEC>::load(gw + c, w); VecIO<float, VEC>::load(gb + c, bb);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { sc[k] = rstd * w[k]; sh[k] = bb[k] - mean * sc[k]; }
  }
  __syncthreads();
  const int in_lo = tab[0].i0;
  if (STAGE) {
    const int n_in = tab[d1 - d0 - 1].i1 - in_lo + 1;
    if (active)
      for (int r = rw; r < n_in; r += RW) {
        const int fo = in_lo + r;
        float a[VEC];
        // phase layout: output row fo = 4q + r - 2 lives in the row of x[q-1] (us geometry), columns r*Cu + c
        VecIO<T, VEC>::load(ug + (long)(((fo + 2) >> 2) - 1) * urow + ((fo + 2) & 3) * Cu + c, a);
        if (has_gn) dec_gn_gelu<T, VEC>(a, sc, sh);
        VecIO<float, VEC>::store(act_s + (long)r * Cu + c, a);
      }
    __syncthreads();
  }
  if (!active) return;
  if (!STAGE) {
    // direct form: rows strided over the block, every iteration independent (two GELU'd source rows + two skip rows)
    for (int d = d0 + rw; d < d1; d += RW) {
      const LerpRow lr = tab[d - d0];
      float a0[VEC], a1[VEC], s0[VEC], s1[VEC], v[VEC];
      VecIO<T, VEC>::load(ug + (long)(((lr.i0 + 2) >> 2) - 1) * urow + ((lr.i0 + 2) & 3) * Cu + c, a0);
      VecIO<T, VEC>::load(ug + (long)(((lr.i1 + 2) >> 2) - 1) * urow + ((lr.i1 + 2) & 3) * Cu + c, a1);
      VecIO<T, VEC>::load(sg + (long)lr.j0 * srow + c, s0);
      VecIO<T, VEC>::load(sg + (long)lr.j1 * srow + c, s1);
      if (has_gn) { dec_gn_gelu<T, VEC>(a0, sc, sh); dec_gn_gelu<T, VEC>(a1, sc, sh); }
      dec_blend<T, VEC>(v, a0, a1, s0, s1, lr.lam, lr.mu);
      VecIO<T, VEC>::store(og + (long)d * orow + c, v);
    }
    return;
  }
  const int RL = (d1 - d0 + RW - 1) / RW;                // contiguous run of output rows per thread
  const int da = d0 + rw * RL, db = min(da + RL, d1);
  int ci0 = -1, ci1 = -1, cj0 = -1, cj1 = -1;
  float a0[VEC], a1[VEC], s0[VEC], s1[VEC], v[VEC];
  for (int d = da; d < db; ++d) {
    const LerpRow lr = tab[d - d0];
    if (lr.i0 != ci0 || lr.i1 != ci1) {
      if (lr.i0 == ci1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) a0[k] = a1[k];
      } else if (STAGE) {
        VecIO<float, VEC>::load(act_s + (long)(lr.i0 - in_lo) * Cu + c, a0);
      } else {
        VecIO<T, VEC>::load(ug + (long)(((lr.i0 + 2) >> 2) - 1) * urow + ((lr.i0 + 2) & 3) * Cu + c, a0);
        if (has_gn) dec_gn_gelu<T, VEC>(a0, sc, sh);
      }
      if (lr.i1 == lr.i0) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) a1[k] = a0[k];
      } else if (STAGE) {
        VecIO<float, VEC>::load(act_s + (long)(lr.i1 - in_lo) * Cu + c, a1);
      } else {
        VecIO<T, VEC>::load(ug + (long)(((lr.i1 + 2) >> 2) - 1) * urow + ((lr.i1 + 2) & 3) * Cu + c, a1);
        if (has_gn) dec_gn_gelu<T, VEC>(a1, sc, sh);
      }
      ci0 = lr.i0; ci1 = lr.i1;
    }
    if (lr.j0 != cj0 || lr.j1 != cj1) {
      if (lr.j0 == cj1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) s0[k] = s1[k];
      } else {
        VecIO<T, VEC>::load(sg + (long)lr.j0 * srow + c, s0);
      }
      if (lr.j1 == lr.j0) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) s1[k] = s0[k];
      } else {
        VecIO<T, VEC>::load(sg + (long)lr.j1 * srow + c, s1);
      }
      cj0 = lr.j0; cj1 = lr.j1;
    }
    dec_blend<T, VEC>(v, a0, a1, s0, s1, lr.lam, lr.mu);
    VecIO<T, VEC>::store(og + (long)d * orow + c, v);
  }
}
template <typename T>
void launch_dec_apply(const T* u, int Uin, RowSpace us, int Cu, T* out, RowSpace os, int G2, int has_gn,
                      const float* mr, const float* gw, const float* gb, const T* skip, RowSpace ss,
                      cudaStream_t st) {
  const int vec = os.C % 8 == 0 ? 8 : 4;
  const double ratio = (double)Uin / (double)os.R;
  // all input rows are stored and used AND a staged row serves several outputs or the group is short; the long ~1:1 resizes of
  // the time branch take the direct form
  const bool stage = Uin <= os.R + 8 && Cu == os.C && !(ratio > 0.9 && os.R >= 2048);
  const int RW = 256 / (os.C / vec);
  // staged: as many output rows per block as 40 KB of staged fp32 input rows allow (amortises the load -> GELU -> sync
  // latency chain); direct: ~128 rows per block
  int TD = stage ? (int)(((40 * 1024) / (4 * Cu) - 4) / ratio) : 256;
  TD = std::max(1, std::min(std::min(TD, 1024), os.R));
  const int tiles = (os.R + TD - 1) / TD;
  TD = (os.R + tiles - 1) / tiles;                          // balanced tiles
  size_t smem = (((size_t)TD * sizeof(LerpRow) + 15) & ~(size_t)15);
  if (stage) smem += (size_t)((int)(TD * ratio) + 4) * Cu * sizeof(float);
  dim3 grid(tiles, os.G);
#define DEC_LAUNCH(V, S) launch_pdl(dec_apply_kernel<T, V, S>, dim3(grid), dim3(256), smem, st, u, Uin, us, Cu, out, os, G2, has_gn, mr, gw, gb, skip, ss, TD)
  if (vec == 8) { if (stage) DEC_LAUNCH(8, true); else DEC_LAUNCH(8, false); }
  else { if (stage) DEC_LAUNCH(4, true); else DEC_LAUNCH(4, false); }
#undef DEC_LAUNCH
}

// ------------------------------------------------------------------ weight packing (fp32 params -> T, GEMM layouts)
// kind 0: plain copy               dst[i] = src[i]
// kind 1: conv k-major             src [Co][Ci][K] -> dst [Co][K*Ci]              (k*Ci + ci)
// kind 2: conv-transpose phases    src [Ci][Co][8] -> dst [4*Co][2*Ci], row r*Co+co, col j*Ci+ci, tap = j==0 ? r+4 : r
// kind 3: GLU interleave rows      src [2C][K]     -> dst row 2j = src row j, row 2j+1 = src row j+C
// kind 4: replicate                src [d0]        -> dst[i] = src[i % d0]
// kind 5/6/7: zero-padded variants (hidden channels of the tensor-core DConv), see below
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, long n, int kind, int d0, int d1, int d2) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long s = i;
    if (kind == 1) {           // d0=Co d1=Ci d2=K ; dst index i = (co*K + k)*Ci + ci
      int ci = (int)(i % d1); long t = i / d1; int k = (int)(t % d2); int co = (int)(t / d2);
      s = ((long)co * d1 + ci) * d2 + k;
    } else if (kind == 2) {    // d0=Ci d1=Co ; dst index i = (r*Co+co)*(2*Ci) + j*Ci + ci
      int ci = (int)(i % d0); long t = i / d0; int j = (int)(t % 2); t /= 2; int co = (int)(t % d1); int r = (int)(t / d1);
      int tap = j == 0 ? r + 4 : r;
      s = ((long)ci * d1 + co) * 8 + tap;
    } else if (kind == 3) {    // d0=2C d1=K
      int k = (int)(i % d1); int row = (int)(i / d1); int j = row >> 1; int half = row & 1;
      s = (long)(j + half * (d0 / 2)) * d1 + k;
    } else if (kind == 4) {    // replicate a [d0] vector
      s = i % d0;
    } else if (kind == 5) {    // conv k-major with zero-padded output rows: d0=Co d1=Ci d2=K ; dst [Cop][K*Ci]
      int ci = (int)(i % d1); long t = i / d1; int k = (int)(t % d2); int co = (int)(t / d2);
      if (co >= d0) { dst[i] = from_f<T>(0.f); continue; }
      s = ((long)co * d1 + ci) * d2 + k;
    } else if (kind == 6) {    // zero-pad a [d0] vector
      if (i >= d0) { dst[i] = from_f<T>(0.f); continue; }
    } else if (kind == 7) {    // GLU-interleaved rows with zero-padded columns: src [d0=2C][d1=K] -> dst [2C][d2=Kp]
      int k = (int)(i % d2); int row = (int)(i / d2); int j = row >> 1; int half = row & 1;
      if (k >= d1) { dst[i] = from_f<T>(0.f); continue; }
      s = (long)(j + half * (d0 / 2)) * d1 + k;
    }
    dst[i] = from_f<T>(src[s]);
  }
}
template <typename T>
void launch_pack_weight(const float* src, T* dst, long n, int kind, int d0, int d1, int d2, cudaStream_t st) {
  pack_weight_kernel<T><<<(int)min((n + 255) / 256, (long)2048), 256, 0, st>>>(src, dst, n, kind, d0, d1, d2);
}

// kinds 8 / 9 / 10: derived from the DConv expand W [C2][H] (fp32 parameters, rounded to the activation dtype like the packed copy
// the MMAs read) and its bias b [C2]:  sum_n e_n^2 = g^T G g + 2 (W^T b)^T g + sum b^2,  sum_n e_n = (W^T 1)^T g + sum b   for e = W g + b.
//   8: hi half of G = W^T W as [HP][HP] (zero beyond H), 9: lo half (G - hi), 10: fp32 [2 W^T b (HP) | W^T 1 (HP) | sum b | sum b^2]
template <typename T>
__global__ void pack_gram_kernel(const float* __restrict__ w, const float* __restrict__ b, T* __restrict__ dst, float* __restrict__ dstf,
                                 int kind, int C2, int H, int HP) {
  const int n_out = kind == 10 ? 2 * HP + 2 : HP * HP;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_out; idx += gridDim.x * blockDim.x) {
    auto r = [&](int n, int j) { return to_f<T>(from_f<T>(w[(long)n * H + j])); };
    double acc = 0.0;
    if (kind == 10) {
      if (idx < HP) { if (idx < H) for (int n = 0; n < C2; ++n) acc += 2.0 * (double)b[n] * (double)r(n, idx); }
      else if (idx < 2 * HP) { const int j = idx - HP; if (j < H) for (int n = 0; n < C2; ++n) acc += (double)r(n, j); }
      else if (idx == 2 * HP) { for (int n = 0; n < C2; ++n) acc += (double)b[n]; }
      else { for (int n = 0; n < C2; ++n) acc += (double)b[n] * (double)b[n]; }
      dstf[idx] = (float)acc;
    } else {
      const int i = idx / HP, j = idx - i * HP;
      if (i < H && j < H) for (int n = 0; n < C2; ++n) acc += (double)r(n, i) * (double)r(n, j);
      const float g = (float)acc;
      const float hi = to_f<T>(from_f<T>(g));
      dst[idx] = kind == 8 ? from_f<T>(g) : from_f<T>(g - hi);
    }
  }
}
template <typename T>
void launch_pack_gram(const float* w, const float* b, void* dst, int kind, int C2, int H, int HP, cudaStream_t st) {
  const int n_out = kind == 10 ? 2 * HP + 2 : HP * HP;
  pack_gram_kernel<T><<<(n_out + 127) / 128, 128, 0, st>>>(w, b, (T*)dst, (float*)dst, kind, C2, H, HP);
}

#define INST(T)                                                                                                         \
  template void launch_pack_gram<T>(const float*, const float*, void*, int, int, int, int, cudaStream_t);               \
  template void launch_pack_wav<T>(const float*, const float*, T*, RowSpace, int, cudaStream_t);                        \
  template void launch_pack_spec<T>(const float*, const float*, T*, RowSpace, int, cudaStream_t);                       \
  template void launch_gn_gelu<T>(T*, RowSpace, int, int, const float*, const float*, const float*, cudaStream_t); \
  template void launch_gn_glu_res<T>(T*, RowSpace, const T*, RowSpace, int, int, const float*, const float*,            \
                                     const float*, const float*, cudaStream_t);                                         \
  template void launch_norm_rows<T>(const T*, T*, T*, long, int, int, const float*, const float*, const float*,         \
                                    const float*, const float*, const float*, RowSpace, cudaStream_t, T*, const float*, const float*); \
  template void launch_softmax_rows<T>(T*, long, int, cudaStream_t);                                                    \
  template void launch_add_rowvec<T>(const T*, T*, long, int, int, const float*, long, cudaStream_t);                   \
  template void launch_dec_apply<T>(const T*, int, RowSpace, int, T*, RowSpace, int, int, const float*,                     \
                                    const float*, const float*, const T*, RowSpace, cudaStream_t);                      \
  template void launch_pack_weight<T>(const float*, T*, long, int, int, int, int, cudaStream_t);
INST(float)
INST(bf16)

}  // namespace athtd
