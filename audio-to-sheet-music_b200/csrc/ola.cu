// Track-level segment gather and weighted overlap-add, replacing the python chunk loop of
// OurModel._chunked_inference (/root/reference/benchmark.py:155-204 == app.py:129-178).
// Gather formulation: every output sample has at most two contributing chunks (stride > len/2),
// accumulated in chunk order with un-fused fp32 multiply / add, then divided by the clamped weight
// sum -- the same operation order as the reference loop, hence bit-identical results for identical
// per-chunk model outputs.  Fade ramps come from torch.linspace on the host side (quirk Q6).
#include "kernels.cuh"

namespace athtd {

// segs[k, c, i] = track[c, starts[k] + i]  (zero beyond T: the reference zero-pads the tail chunk)
__global__ void gather_chunks_kernel(const float* __restrict__ track, long T, int C, const long* __restrict__ starts,
                                     int chunk_len, float* __restrict__ segs) {
  pdl_begin();
  int k = blockIdx.y / C, c = blockIdx.y % C;
  long s0 = starts[k];
  const float* src = track + (long)c * T;
  float* dst = segs + ((long)k * C + c) * chunk_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunk_len; i += gridDim.x * blockDim.x) {
    long s = s0 + i;
    dst[i] = s < T ? src[s] : 0.f;
  }
}
void launch_gather_chunks(const float* track, long T, int C, const long* starts, int n_chunks, int chunk_len, float* segs,
                          cudaStream_t st) {
  launch_pdl(gather_chunks_kernel, dim3(dim3(min((chunk_len + 255) / 256, 1024), n_chunks * C)), dim3(256), 0, st, track, T, C, starts, chunk_len, segs);
}

__device__ __forceinline__ float chunk_w(int i, int actual, int fade, int flags, const float* __restrict__ up,
                                         const float* __restrict__ down) {
  // fade-in mask x fade-out mask (torchaudio Fade order; identical to the reference's slice assignments whenever the two ramps do not
  // overlap, which benchmark.py:185 guarantees with fade_len = min(overlap, actual_len // 2); a factor 1.0f is exact)
  float w_in = 1.0f, w_out = 1.0f;
  if ((flags & 1) && i < fade) w_in = up[i];
  if ((flags & 2) && i >= actual - fade) w_out = down[i - (actual - fade)];
  return __fmul_rn(w_in, w_out);
}

// seg_out: model output of global chunk k at seg_out + (k - k_base)*seg_stride, layout [C, chunk_len].
// Writes out[c, s - t_begin ... ] for s in [t_begin, t_end); out has row pitch out_pitch and is indexed from t_begin.
__global__ void chunk_ola_kernel(const float* __restrict__ seg_out, long seg_stride, int k_base, int chunk_len,
                                 const long* __restrict__ starts, const int* __restrict__ actual_len,
                                 const int* __restrict__ fade_len, const int* __restrict__ flags, int n_chunks, long stride,
                                 const float* __restrict__ ramp_up, const float* __restrict__ ramp_down,
                                 const int* __restrict__ ramp_off, float* __restrict__ out, long out_pitch, int C,
                                 long t_begin, long t_end, int normalize) {
  pdl_begin();
  for (long s = t_begin + (long)blockIdx.x * blockDim.x + threadIdx.x; s < t_end; s += (long)gridDim.x * blockDim.x) {
    int k_hi = (int)(s / stride); if (k_hi > n_chunks - 1) k_hi = n_chunks - 1;
    int k_lo = k_hi;
    if (k_hi >= 1 && s < starts[k_hi - 1] + actual_len[k_hi - 1]) k_lo = k_hi - 1;
    float wsum = 0.f;
    float acc[2] = {0.f, 0.f};
    for (int k = k_lo; k <= k_hi; ++k) {
      int i = (int)(s - starts[k]);
      if (i < 0 || i >= actual_len[k]) continue;
      float w = chunk_w(i, actual_len[k], fade_len[k], flags[k], ramp_up + ramp_off[k], ramp_down + ramp_off[k]);
      const float* so = seg_out + (long)(k - k_base) * seg_stride;
      for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(so[(long)c * chunk_len + i], w));
      wsum = __fadd_rn(wsum, w);
    }
    wsum = fmaxf(wsum, 1e-8f);
    // normalize = 0: plain weighted sum (test_inference.py:113-141: faded chunks are added, never divided)
    for (int c = 0; c < C; ++c) out[(long)c * out_pitch + (s - t_begin)] = normalize ? __fdiv_rn(acc[c], wsum) : acc[c];
  }
}
void launch_chunk_ola(const float* seg_out, long seg_stride, int k_base, int chunk_len, const long* starts,
                      const int* actual_len, const int* fade_len, const int* flags, int n_chunks, long stride,
                      const float* ramp_up, const float* ramp_down, const int* ramp_off, float* out, int C, long t_begin,
                      long t_end, int normalize, cudaStream_t st) {
  long n = t_end - t_begin;
  if (n <= 0) return;
  launch_pdl(chunk_ola_kernel, dim3((int)min((n + 255) / 256, (long)148 * 16)), dim3(256), 0, st, seg_out, seg_stride, k_base, chunk_len, starts,
                                                                              actual_len, fade_len, flags, n_chunks, stride,
                                                                              ramp_up, ramp_down, ramp_off, out, n, C,
                                                                              t_begin, t_end, normalize);
}

}  // namespace athtd
