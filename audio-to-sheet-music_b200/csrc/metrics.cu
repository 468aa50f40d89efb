// On-device evaluation metrics of the path's callers: the sums behind sdr_loss / sisdr_loss /
// new_sdr_metric (/root/reference/src/loss.py:9-87; called per stem and track by benchmark.py:555-588 and :669-670
// on CPU copies of the separated waveforms).  One bandwidth-bound pass over (estimate, target) per item:
//   sums[item] = { sum t, sum e, sum t^2, sum e^2, sum e t, sum (t - e)^2 }     (fp64 accumulators)
// The dB values are closed forms of these sums (metrics.py); the separated stems never leave the device.
#include "kernels.cuh"
#include <algorithm>

namespace athtd {

__global__ void __launch_bounds__(256) sdr_sums_kernel(const float* __restrict__ est, const float* __restrict__ tgt, long n,
                                                       double* __restrict__ sums) {
  const int item = blockIdx.y;
  const float* e = est + (long)item * n;
  const float* t = tgt + (long)item * n;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long n4 = ((((uintptr_t)e | (uintptr_t)t) & 15) == 0) ? n / 4 : 0;      // 16-byte vector body when both rows are aligned
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 a = ((const float4*)e)[i], b = ((const float4*)t)[i];
    const float ev[4] = {a.x, a.y, a.z, a.w}, tv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = tv[k] - ev[k];
      s[0] += tv[k]; s[1] += ev[k]; s[2] = fmaf(tv[k], tv[k], s[2]); s[3] = fmaf(ev[k], ev[k], s[3]);
      s[4] = fmaf(ev[k], tv[k], s[4]); s[5] = fmaf(d, d, s[5]);
    }
  }
  for (long i = 4 * n4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float ev = e[i], tv = t[i], d = tv - ev;
    s[0] += tv; s[1] += ev; s[2] = fmaf(tv, tv, s[2]); s[3] = fmaf(ev, ev, s[3]); s[4] = fmaf(ev, tv, s[4]); s[5] = fmaf(d, d, s[5]);
  }
  // per-thread fp32 partials over <= a few thousand elements, then fp64 across the block and the grid
  __shared__ double red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = (double)s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(sums + 6L * item + threadIdx.x, v);
  }
}

void launch_sdr_sums(const float* est, const float* tgt, int items, long n, double* sums, cudaStream_t st) {
  cudaMemsetAsync(sums, 0, sizeof(double) * 6 * items, st);
  // ~4096 elements per thread at most, >= 4 CTAs per SM in total when the problem allows
  long bx = (n + 256L * 16 - 1) / (256L * 16);
  const long cap = std::max(1L, (148L * 8 + items - 1) / items);
  bx = std::max(1L, std::min(bx, cap));
  sdr_sums_kernel<<<dim3((unsigned)bx, items), 256, 0, st>>>(est, tgt, n, sums);
}

}  // namespace athtd
