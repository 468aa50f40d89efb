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
    *(uint32_t*)&stage[warp][g][c] = pack_bf16x2(gelu_fast(d[0] + b0), gelu_fast(d[1] + b1));
    *(uint32_t*)&stage[warp][g + 8][c] = pack_bf16x2(gelu_fast(d[2] + b0), gelu_fast(d[3] + b1));
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
  tenc0_conv_kernel<<<dim3(tiles, ys.batch()), 256, 0, st>>>(wav, meanstd, L, w, bias, y, ys);
}

}  // namespace athtd
