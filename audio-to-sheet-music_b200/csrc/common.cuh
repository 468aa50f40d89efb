// Common device helpers for the AudioTextHTDemucs B200 path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include <string.h>

namespace athtd {

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// exact (erf) GELU: torch F.gelu default / nn.GELU(approximate='none')
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// bf16 build: erf-GELU through erf(x/sqrt2) ~ tanh(x * (a0 + a1 x^2 + a2 x^4)) (fit of the odd function atanh(erf), |GELU
// error| <= 2.9e-5 with an exact tanh) evaluated with tanh.approx.f32 (relative error 2^-11): 7 FMA-pipe instructions + 1 SFU
// op per value instead of the 17 of a pure polynomial -- the GEMM epilogues, the decoder tail and the per-slab encoder
// kernels are instruction-issue bound on this function.  Total |error| <= 0.5 |x| 4.9e-4, below the bf16 rounding of the
// result.  x^2 is clamped so that the quartic term cannot flip the sign for |x| > 8.  The fp32 build keeps erff().
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 32.0f);
  float p = fmaf(-3.5471127794746665e-4f, x2, 3.702029975737703e-2f);
  p = fmaf(p, x2, 7.975055041244574e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): exact fp32 arithmetic, two lanes per issue slot.  The elementwise kernels
// of this path are instruction-issue bound (ncu: 62-77 % issue-active, no unit saturated), not FP32-pipe bound.
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b), "l"(*(unsigned long long*)&c));
  return *(float2*)&d;
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b));
  return *(float2*)&d;
}
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b));
  return *(float2*)&d;
}
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b));
  return *(float2*)&d;
}
__device__ __forceinline__ float2 f2splat(float a) { return make_float2(a, a); }
// gelu_fast on a pair: the same operation sequence per lane (bit-identical to two scalar calls)
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  float2 x2 = f2mul(x, x);
  x2.x = fminf(x2.x, 32.0f); x2.y = fminf(x2.y, 32.0f);
  float2 p = f2fma(f2splat(-3.5471127794746665e-4f), x2, f2splat(3.702029975737703e-2f));
  p = f2fma(p, x2, f2splat(7.975055041244574e-1f));
  const float2 u = f2mul(x, p);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hx = f2mul(f2splat(0.5f), x);
  return f2fma(hx, t, hx);
}
template <typename T> __device__ __forceinline__ float gelu_act(float x);
template <> __device__ __forceinline__ float gelu_act<float>(float x) { return gelu_erf(x); }
template <> __device__ __forceinline__ float gelu_act<bf16>(float x) { return gelu_fast(x); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// mean / rstd from double (sum, sumsq) accumulators, biased variance (GroupNorm / LayerNorm)
__device__ __forceinline__ void stats_to_mean_rstd(const double* st, double count, float eps,
                                                   float& mean, float& rstd) {
  double m = st[0] / count;
  double var = st[1] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// 16-byte (8 x bf16) / 32-byte (8 x fp32) vector access helpers
template <typename T, int VEC> struct VecIO;
template <int VEC> struct VecIO<float, VEC> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) { float4 t = *(const float4*)(p + i); v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w; }
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) *(float4*)(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
};
template <int VEC> struct VecIO<bf16, VEC> {
  static __device__ __forceinline__ void load(const bf16* p, float* v) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      uint2 t = *(const uint2*)(p + i);
      float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&t.x), c = __bfloat1622float2(*(const __nv_bfloat162*)&t.y);
      v[i] = a.x; v[i + 1] = a.y; v[i + 2] = c.x; v[i + 3] = c.y;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* v) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v[i], v[i + 1]), c = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
      uint2 t; t.x = *(uint32_t*)&a; t.y = *(uint32_t*)&c;
      *(uint2*)(p + i) = t;
    }
  }
};

// One-time-per-device launch setup (cudaFuncSetAttribute is per device / context, and an Engine may live on any GPU of the
// process) and the SM count of the current device.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
inline int device_sm_count() {
  static int n[64] = {};
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (n[d] == 0) {
    cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
    if (n[d] <= 0) n[d] = 148;
  }
  return n[d];
}

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the stream is still
// draining -- its CTAs become resident as soon as every CTA of the predecessor has executed pdl_trigger() (or exited), run
// their prologue (barrier init, TMEM allocation, staging of CONSTANT data such as weights) and block in pdl_wait() until the
// predecessor has completed and its memory is visible.  Every kernel that is launched with launch_pdl() calls pdl_wait() before it
// touches anything a predecessor may have written.  Both instructions are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_begin() { pdl_trigger(); pdl_wait(); }
bool pdl_enabled();
void pdl_set_enabled(bool on);
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// A padded channels-last row space.  G = B*G2 interior groups (G2 groups per segment: frames on the
// frequency branch, 1 on the time branch); every segment stores G2p >= G2 groups (gpf zero groups in front) and
// every group stores Rp >= R rows (pf zero rows in front).  Pads are written once (zero) and never again, which
// is what lets every convolution of the path be a GEMM over row-shifted views of the same buffer.
struct RowSpace {
  int G;    // interior groups = B * G2
  int R;    // interior rows per group
  int Rp;   // stored rows per group
  int pf;   // front pad rows
  int C;    // channels per row (row pitch in elements)
  int G2;   // interior groups per segment
  int G2p;  // stored groups per segment
  int gpf;  // front pad groups
  __host__ __device__ long row_off(int g, int r) const {
    int b = g / G2, t = g - b * G2;
    return (((long)b * G2p + gpf + t) * Rp + pf + r) * (long)C;
  }
  // the same with the segment and the group inside it already split (no division: per-element loops of the row kernels)
  __host__ __device__ long row_off_bt(int b, int t, int r) const { return (((long)b * G2p + gpf + t) * Rp + pf + r) * (long)C; }
  __host__ __device__ long elems() const { return (long)(G / G2) * G2p * Rp * C; }
  __host__ __device__ long rows_total() const { return (long)(G / G2) * G2p * Rp; }
  __host__ __device__ long g1_stride() const { return (long)G2p * Rp * C; }
  __host__ __device__ long g2_stride() const { return (long)Rp * C; }
  __host__ __device__ long origin() const { return ((long)gpf * Rp + pf) * C; }
  __host__ __device__ int batch() const { return G / G2; }
};

// x / d for 0 <= x < 2^31 from a host-computed (multiplier, shift): 3 instructions instead of the ~35-instruction
// (~100-cycle dependent chain) integer division, four of which sat on the per-tile critical path of the epilogue
__device__ __forceinline__ int fast_div(int x, const uint32_t (&fd)[2]) {
  return (int)((__umulhi((uint32_t)x, fd[0]) + (uint32_t)x) >> fd[1]);
}
inline void fast_div_init(uint32_t d, uint32_t (&fd)[2]) {
  uint32_t shr = 0;
  while ((1u << shr) < d) ++shr;
  fd[0] = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << shr) - d)) / d + 1);
  fd[1] = shr;
}

}  // namespace athtd
