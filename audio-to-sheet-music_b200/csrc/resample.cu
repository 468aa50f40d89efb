// Host-side staging row f-1 of the scope table: the reference's `load_audio` (/root/reference/app.py:113-126) resamples the
// decoded file to 44.1 kHz with torchaudio.transforms.Resample (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) and
// repeats a mono channel to stereo before the chunk loop.  torchaudio evaluates the polyphase FIR as
//     xpad = pad(x, (width, width + o));  y[i * nw + p] = sum_k K[p][k] * xpad[i * o + k]        (conv1d, stride o)
// with o = orig / gcd, nw = new / gcd and K = _get_sinc_resample_kernel (built on the host in float64 by audio.py, the same
// closed form).  Here: one thread per output sample, the input window of a block staged in shared memory, the filter bank read
// k-major ([k][p], coalesced over the phases of a warp), channels duplicated on the fly (mono -> stereo).
#include "kernels.cuh"
#include <algorithm>

namespace athtd {

static constexpr int RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS) resample_kernel(const float* __restrict__ x, int C_in, long T_in,
                                                               const float* __restrict__ Kt, int o, int nw, int taps, int width,
                                                               float* __restrict__ y, int C_out, long T_out, int IB) {
  extern __shared__ float xs[];                       // xpad[i0 * o, i0 * o + IB * o + taps)
  const int c = blockIdx.y;
  const int ci = c < C_in ? c : C_in - 1;             // mono -> stereo: every output channel reads the last input channel
  const long i0 = (long)blockIdx.x * IB;
  const float* xc = x + (long)ci * T_in;
  const int span = IB * o + taps;
  for (int j = threadIdx.x; j < span; j += RS_THREADS) {
    const long src = i0 * o + j - width;              // index into the un-padded signal
    xs[j] = (src >= 0 && src < T_in) ? xc[src] : 0.f;
  }
  __syncthreads();
  float* yc = y + (long)c * T_out;
  for (int idx = threadIdx.x; idx < IB * nw; idx += RS_THREADS) {
    const int il = idx / nw, p = idx - il * nw;
    const long n = (i0 + il) * nw + p;
    if (n >= T_out) continue;
    const float* xw = xs + il * o;
    float acc = 0.f;
    for (int k = 0; k < taps; ++k) acc = fmaf(Kt[(long)k * nw + p], xw[k], acc);
    yc[n] = acc;
  }
}

// identity rate: channel copy / duplication only
__global__ void channel_copy_kernel(const float* __restrict__ x, int C_in, long T, float* __restrict__ y) {
  const int c = blockIdx.y;
  const int ci = c < C_in ? c : C_in - 1;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (long)gridDim.x * blockDim.x) y[(long)c * T + i] = x[(long)ci * T + i];
}

int launch_resample(const float* x, int C_in, long T_in, const float* Kt, int o, int nw, int taps, int width, float* y, int C_out,
                    long T_out, cudaStream_t st) {
  if (Kt == nullptr) {
    channel_copy_kernel<<<dim3((unsigned)std::min<long>((T_in + 255) / 256, 4096), C_out), 256, 0, st>>>(x, C_in, T_in, y);
    return 0;
  }
  // IB input strides per block: ~2048 outputs per block, shared memory (IB * o + taps) floats
  int IB = std::max(1, 2048 / nw);
  while (IB > 1 && (size_t)(IB * o + taps) * 4 > 96 * 1024) IB /= 2;
  const size_t smem = (size_t)(IB * o + taps) * 4;
  if (smem > 200 * 1024) return 1;
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const long n_i = (T_out + nw - 1) / nw;
  resample_kernel<<<dim3((unsigned)((n_i + IB - 1) / IB), C_out), RS_THREADS, smem, st>>>(x, C_in, T_in, Kt, o, nw, taps, width, y, C_out,
                                                                                          T_out, IB);
  return 0;
}

}  // namespace athtd
