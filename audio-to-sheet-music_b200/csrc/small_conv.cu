// Time-branch encoder level 0 (demucs hdemucs.py:HEncLayer conv, Conv1d(2, 48, k8, s4, p2) + GELU; ATHTDemucs_v2.py:204)
// fused with the input normalisation (ATHTDemucs_v2.py:272-275): reads the fp32 waveform, writes bf16 channels-last rows.
//
//   y[b, t, co] = GELU( bias[co] + sum_{j<8, ci<2} W[co, ci, j] * xn[b, ci, 4t - 2 + j] ),   xn = (wav - mean) / (1e-5 + std)
//
// K = 16: one mma.sync k-step per 16 output rows.  A tcgen05 tile for this layer was epilogue-bound (466 us at batch 32
// against a 41 us HBM floor); here a warp owns 16 rows, builds the A fragment straight from coalesced fp32 loads (lane
// (g, q) <-> row g, tap q / q+4, both channels), and writes its 16 x 48 tile as one contiguous 1536-byte block.
#include "kernels.cuh"
#include "mma_sync.cuh"

namespace athtd {

__global__ void __launch_bounds__(256) tenc0_conv_kernel(const float* __restrict__ wav, const float* __restrict__ meanstd, int L,
                                                         const bf16* __restrict__ w /*[48][16] k = tap*2 + ci*/,
                                                         const float* __restrict__ bias, bf16* __restrict__ y, RowSpace ys) {
  pdl_begin();
  __shared__ __align__(16) bf16 stage[8][16][56];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int t0 = (blockIdx.x * 8 + warp) * 16;
  if (t0 >= ys.R) return;
  const float mean = meanstd[2 * b], inv = 1.0f / (1e-5f + meanstd[2 * b + 1]);
  const float* w0 = wav + (long)b * 2 * L, *w1 = w0 + L;
  auto ld = [&](int row, int tap) -> uint32_t {
    const int i = 4 * row - 2 + tap;
    if (i < 0 || i >= L) return 0u;                          // zero padding of the NORMALISED signal
    return pack_bf16x2((w0[i] - mean) * inv, (w1[i] - mean) * inv);
  };
  uint32_t a[4];
  a[0] = ld(t0 + g, q); a[1] = ld(t0 + g + 8, q); a[2] = ld(t0 + g, q + 4); a[3] = ld(t0 + g + 8, q + 4);
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    uint32_t bb[2];
    frag_b(w, 16, nt * 8, 0, lane, bb);
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816(d, a, bb);
    const int c = nt * 8 + 2 * q;
    const float b0 = bias[c], b1 = bias[c + 1];
    const float2 ga = gelu_fast2(f2add(make_float2(d[0], d[1]), make_float2(b0, b1)));
    const float2 gb = gelu_fast2(f2add(make_float2(d[2], d[3]), make_float2(b0, b1)));
    *(uint32_t*)&stage[warp][g][c] = pack_bf16x2(ga.x, ga.y);
    *(uint32_t*)&stage[warp][g + 8][c] = pack_bf16x2(gb.x, gb.y);
  }
  __syncwarp();
  bf16* dst = y + ys.row_off(b, t0);
  const int nrows = min(16, ys.R - t0);
#pragma unroll
  for (int i = lane; i < 16 * 6; i += 32) {
    const int r = i / 6, ch = i - r * 6;
    if (r < nrows) *(uint4*)(dst + (long)r * 48 + ch * 8) = *(const uint4*)&stage[warp][r][ch * 8];
  }
}

void launch_tenc0_conv(const float* wav, const float* meanstd, int L, const bf16* w, const float* bias, bf16* y, RowSpace ys,
                       cudaStream_t st) {
  const int tiles = (ys.R + 127) / 128;
  launch_pdl(tenc0_conv_kernel, dim3(dim3(tiles, ys.batch())), dim3(256), 0, st, wav, meanstd, L, w, bias, y, ys);
}

}  // namespace athtd

namespace athtd {

__device__ __forceinline__ void lerp_coords_s(int d, int in, int out, int& i0, int& i1, float& lam) {
  if (in == out) { i0 = d; i1 = d; lam = 0.f; return; }
  float scale = (float)in / (float)out;
  float src = scale * ((float)d + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  lam = src - (float)i0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Last FreqDecoder layer (ATHTDemucs_v2.py:82-104, i = 3: ConvTranspose2d(48, 4, (8,1), (4,1), (2,0)), no norm / GELU) fused
// with its exact 4:1 resize and the skip: the resize reads only rows 4d+1 (phase 3 of input row d) and 4d+2 (phase 0 of
// input row d+1), so
//     out[d] = 0.5 (x[d-1] W[..,7] + x[d] (W[..,3] + W[..,4]) + x[d+1] W[..,0]) + b + 0.1 lerp(skip)[d]
// is a 3-tap, 48 -> 4 channel conv (K = 144, one mma.sync n-tile) instead of a 16-column GEMM + a separate apply pass.
// One CTA per (segment, frame) group; rows of a group are contiguous, so A row d is 144 contiguous bf16 from x[d-1].
__global__ void __launch_bounds__(288) dec_last_freq_kernel(const bf16* __restrict__ x, RowSpace xs, const float* __restrict__ w /*[48][4][8]*/,
                                                            const float* __restrict__ bias, const bf16* __restrict__ skip, RowSpace ss,
                                                            bf16* __restrict__ out, RowSpace os) {
  pdl_begin();
  __shared__ __align__(16) bf16 ws[8][152];
  const int g_ = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  for (int i = tid; i < 8 * 144; i += blockDim.x) {
    const int n = i / 144, k = i - n * 144, tap = k / 48, ci = k - tap * 48;
    float v = 0.f;
    if (n < 4) {
      const float* wp = w + (ci * 4 + n) * 8;
      v = tap == 0 ? 0.5f * wp[7] : (tap == 1 ? 0.5f * (wp[3] + wp[4]) : 0.5f * wp[0]);
    }
    ws[n][k] = __float2bfloat16_rn(v);
  }
  const int R = os.R;
  const bf16* xg = x + xs.row_off(g_, 0);
  const bf16* sg = skip + ss.row_off(g_, 0);
  bf16* og = out + os.row_off(g_, 0);
  __syncthreads();
  // A fragments straight from global memory (a shared-memory staged variant with ldmatrix measured 20 % slower: the extra
  // block barrier and staging pass cost more than the 4-byte gathers)
  for (int mt = warp; mt * 16 < R; mt += blockDim.x >> 5) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    const bf16* a_lo = xg + (long)(mt * 16 + g - 1) * 48, *a_hi = a_lo + 8 * 48;
#pragma unroll
    for (int kc = 0; kc < 9; ++kc) {
      uint32_t a[4], bb[2];
      a[0] = ld_b32(a_lo + kc * 16 + 2 * q); a[1] = ld_b32(a_hi + kc * 16 + 2 * q);
      a[2] = ld_b32(a_lo + kc * 16 + 8 + 2 * q); a[3] = ld_b32(a_hi + kc * 16 + 8 + 2 * q);
      frag_b(&ws[0][0], 152, 0, kc * 16, lane, bb);
      mma16816(d, a, bb);
    }
    if (q < 2) {
      const int co = 2 * q;
      const float b0 = bias[co], b1 = bias[co + 1];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int dr = mt * 16 + g + 8 * hh;
        if (dr < R) {
          int j0, j1; float mu;
          lerp_coords_s(dr, ss.R, R, j0, j1, mu);
          const float2 s0 = unpack_bf16x2(ld_b32(sg + (long)j0 * ss.C + co)), s1 = unpack_bf16x2(ld_b32(sg + (long)j1 * ss.C + co));
          const float o0 = d[2 * hh] + b0 + 0.1f * ((1.f - mu) * s0.x + mu * s1.x);
          const float o1 = d[2 * hh + 1] + b1 + 0.1f * ((1.f - mu) * s0.y + mu * s1.y);
          *(uint32_t*)(og + (long)dr * os.C + co) = pack_bf16x2(o0, o1);
        }
      }
    }
  }
}

// Last TimeDecoder layer (ATHTDemucs_v2.py:125-139, i = 3: ConvTranspose1d(48, 4, 8, 4, 2), identity resize) + 0.1 * the
// 4x up-sampled skip: the four output phases of input row q are 16 GEMM columns (K = 96: [x[q-1], x[q]]), written straight
// to output rows 4q + r - 2 (a warp's 16 input rows cover 64 consecutive output rows = one contiguous 512-byte block).
__global__ void __launch_bounds__(256) dec_last_time_kernel(const bf16* __restrict__ x, RowSpace xs, const bf16* __restrict__ w /*[16][96]*/,
                                                            const float* __restrict__ bias4 /*[16]*/, const bf16* __restrict__ skip,
                                                            RowSpace ss, bf16* __restrict__ out, RowSpace os) {
  pdl_begin();
  __shared__ __align__(16) bf16 ws[16][104];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  for (int i = tid; i < 16 * 12; i += blockDim.x) {
    const int n = i / 12, kc = i - n * 12;
    *(uint4*)&ws[n][kc * 8] = *(const uint4*)(w + n * 96 + kc * 8);
  }
  const int Rin = xs.R, L = os.R;
  const bf16* xb = x + xs.row_off(b, 0);
  const bf16* sb = skip + ss.row_off(b, 0);
  bf16* ob = out + os.row_off(b, 0);
  __syncthreads();
  const int q0 = (blockIdx.x * 8 + warp) * 16;
  if (q0 > Rin) return;
  float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const int r_lo = min(q0 + g, Rin), r_hi = min(q0 + g + 8, Rin);      // rows past Rin recompute row Rin (results discarded)
  const bf16* a_lo = xb + (long)(r_lo - 1) * 48, *a_hi = xb + (long)(r_hi - 1) * 48;
#pragma unroll
  for (int kc = 0; kc < 6; ++kc) {
    uint32_t a[4], bb[2];
    a[0] = ld_b32(a_lo + kc * 16 + 2 * q); a[1] = ld_b32(a_hi + kc * 16 + 2 * q);
    a[2] = ld_b32(a_lo + kc * 16 + 8 + 2 * q); a[3] = ld_b32(a_hi + kc * 16 + 8 + 2 * q);
    frag_b(&ws[0][0], 104, 0, kc * 16, lane, bb); mma16816(d[0], a, bb);
    frag_b(&ws[0][0], 104, 8, kc * 16, lane, bb); mma16816(d[1], a, bb);
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = nt * 8 + 2 * q, r = col >> 2, co = col & 3;      // phase r, channels co, co+1
    const float b0 = bias4[col], b1 = bias4[col + 1];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int qq = q0 + g + 8 * hh;
      const int s = 4 * qq + r - 2;
      if (qq <= Rin && s >= 0 && s < L) {
        int j0, j1; float mu;
        lerp_coords_s(s, ss.R, L, j0, j1, mu);
        const float2 s0 = unpack_bf16x2(ld_b32(sb + (long)j0 * ss.C + co)), s1 = unpack_bf16x2(ld_b32(sb + (long)j1 * ss.C + co));
        const float o0 = d[nt][2 * hh] + b0 + 0.1f * ((1.f - mu) * s0.x + mu * s1.x);
        const float o1 = d[nt][2 * hh + 1] + b1 + 0.1f * ((1.f - mu) * s0.y + mu * s1.y);
        *(uint32_t*)(ob + (long)s * os.C + co) = pack_bf16x2(o0, o1);
      }
    }
  }
}

void launch_dec_last_freq(const bf16* x, RowSpace xs, const float* w, const float* bias, const bf16* skip, RowSpace ss, bf16* out,
                          RowSpace os, cudaStream_t st) {
  launch_pdl(dec_last_freq_kernel, dim3(os.G), dim3(288), 0, st, x, xs, w, bias, skip, ss, out, os);
}
void launch_dec_last_time(const bf16* x, RowSpace xs, const bf16* w, const float* bias4, const bf16* skip, RowSpace ss, bf16* out,
                          RowSpace os, cudaStream_t st) {
  const int tiles = (xs.R + 1 + 127) / 128;
  launch_pdl(dec_last_time_kernel, dim3(dim3(tiles, os.batch())), dim3(256), 0, st, x, xs, w, bias4, skip, ss, out, os);
}

}  // namespace athtd
