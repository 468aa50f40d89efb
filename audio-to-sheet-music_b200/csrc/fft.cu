// Spectral front / back end of the path, replacing htdemucs._spec / _magnitude / _ispec
// (call sites /root/reference/src/models/stem_separation/ATHTDemucs_v2.py:261-262, 297-310;
//  semantics: demucs 4.0.1 spec.py / htdemucs.py).
//
// One CTA of 256 threads transforms one (segment, frame): the stereo pair is packed as
// real=left, imag=right into ONE 4096-point complex FFT held in shared memory
// (radix-16 x radix-16 x radix-16, three register passes with padded transposes), and the two
// real spectra are recovered from the Hermitian symmetry.  Fused around it:
//   forward : single-level reflect padding, periodic hann window, 1/64 (normalized=True),
//             Nyquist-bin drop, complex-as-channels packing [B,Tf,2048,(Lre,Lim,Rre,Rim)],
//             per-segment sum / sum-of-squares partials for the input normalisation.
//   inverse : freq_out 1x1 conv, 259->2048 linear resize of the mask logits, sigmoid,
//             the signed-CaC mask*phase product (quirk Q4), Hermitian extension with zero
//             Nyquist, inverse FFT, window and 1/64 scale, 4-frame overlap-add in a shared-memory
//             ring (fixed order), the constant window-sum envelope 1/1.5, the de-normalised
//             time branch, [B,2,L] store: the windowed frames never reach memory.
#include "kernels.cuh"
#include <algorithm>

namespace athtd {

#define FFT_N 4096
#define FFT_PITCH 272          // 256 + 16: half-warps land on disjoint banks
#define FFT_SMEM (16 * FFT_PITCH)

// Complex arithmetic on packed fp32 pairs (re, im) = one 64-bit register pair: FADD2 / FMUL2 / FFMA2 (ptxas folds the
// scalar broadcasts and the re <-> im swaps into operand modifiers).  The transform is instruction-issue bound (~1600 SASS
// instructions per thread and frame, 850 of them fp32 arithmetic in the scalar form), not HBM bound.
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  // (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
  return f2fma(f2splat(a.y), make_float2(-b.y, b.x), f2mul(f2splat(a.x), b));
}
// v * (wr + i wi) with compile-time wr, wi: v * (wr, wr) + swap(v) * (-wi, wi)
__device__ __forceinline__ float2 cmul_const(float2 v, float wr, float wi) {
  return f2fma(make_float2(v.y, v.x), make_float2(-wi, wi), f2mul(v, f2splat(wr)));
}

// 4-point DFT, forward uses W4 = -i, inverse +i
template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = f2add(a, c), s1 = f2sub(a, c), s2 = f2add(b, d), s3 = f2sub(b, d);
  // forward: y1 = s1 - i*s3 ; y3 = s1 + i*s3.   (-i)*(x+iy) = (y, -x): swap(s3) * (1, -1)
  const float2 sw = make_float2(s3.y, s3.x);
  const float2 pm = INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f), mp = INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f);
  a = f2add(s0, s2);
  c = f2sub(s0, s2);
  b = f2fma(sw, pm, s1);
  d = f2fma(sw, mp, s1);
}

// 16-point DFT in registers, natural order in and out.
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
  // step 1: for each n2, DFT4 over n1 of v[4*n1 + n2]  -> t[n2][k1] stored back at v[4*k1 + n2]
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // step 2: twiddle W16^(n2*k1)
  const float sg = INV ? 1.f : -1.f;
  const float cs[10] = {1.f, C1, C2, S1, 0.f, -S1, -C2, -C1, -1.f, -C1};
  const float sn[10] = {0.f, S1, C2, C1, 1.f, C1, C2, S1, 0.f, -S1};
#pragma unroll
  for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
    for (int n2 = 1; n2 < 4; ++n2) {
      const int m = k1 * n2;
      v[4 * k1 + n2] = cmul_const(v[4 * k1 + n2], cs[m], sg * sn[m]);
    }
  // step 3: for each k1, DFT4 over n2 -> X[k1 + 4*k2] ; currently at v[4*k1 + n2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // now v[4*k1 + k2] = X[k1 + 4*k2] : transpose the 4x4 to natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) { float2 t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}

// w[k] = W^(base*k), k = 1..15, from FOUR table loads (k = 1, 2, 4, 8) and eleven complex products (product depth <= 3,
// relative error <= 4e-7).  Loading all fifteen entries made the kernel L1-wavefront bound: the strided gathers
// tw[base*k] touch ~240 cache lines per warp and pass, three times the shared-memory traffic of the transform itself.
template <bool INV>
__device__ __forceinline__ void twiddle_powers(const float2* __restrict__ tw, int base, float2 (&w)[16]) {
  w[1] = tw[base]; w[2] = tw[2 * base]; w[4] = tw[4 * base]; w[8] = tw[8 * base];
  if (INV) { w[1].y = -w[1].y; w[2].y = -w[2].y; w[4].y = -w[4].y; w[8].y = -w[8].y; }
  w[3] = cmul(w[1], w[2]); w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
#pragma unroll
  for (int k = 1; k < 8; ++k) w[8 + k] = cmul(w[8], w[k]);
}

// Passes 2 and 3 plus the natural-order write-back; on entry v holds pass-1 outputs V[k1]
// (already multiplied by W4096^(t*k1)) of thread t.  On exit sc holds X[k] at index k + (k>>4).
// The transposes go through ONE array of (re, im) pairs: 64-bit shared-memory accesses, half the LDS / STS instructions of
// separate real / imaginary planes (ncu of the split version: mio_throttle was the top stall, LSU the busiest pipe at 43 %);
// every access pattern below is conflict-free per half warp (consecutive pairs, or stride 17 pairs = 34 words).
template <bool INV>
__device__ __forceinline__ void fft4096_tail(float2 (&v)[16], float2* sc, const float2* __restrict__ tw) {
  const int t = threadIdx.x;
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) sc[k1 * FFT_PITCH + t] = v[k1];
  __syncthreads();
  const int hi = t >> 4, lo = t & 15;     // pass 2: (k1 = hi, m2 = lo)
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) v[m1] = sc[hi * FFT_PITCH + 16 * m1 + lo];
  dft16<INV>(v);
  {
    float2 w[16];
    twiddle_powers<INV>(tw, 16 * lo, w);
#pragma unroll
    for (int j1 = 1; j1 < 16; ++j1) v[j1] = cmul(v[j1], w[j1]);
  }
  __syncthreads();
#pragma unroll
  for (int j1 = 0; j1 < 16; ++j1) sc[hi * FFT_PITCH + lo * 17 + j1] = v[j1];
  __syncthreads();
  // pass 3: (k1 = hi, j1 = lo) reads over m2
#pragma unroll
  for (int m2 = 0; m2 < 16; ++m2) v[m2] = sc[hi * FFT_PITCH + m2 * 17 + lo];
  dft16<INV>(v);
  __syncthreads();
#pragma unroll
  for (int j2 = 0; j2 < 16; ++j2) {
    int k = hi + 16 * lo + 256 * j2;
    sc[k + (k >> 4)] = v[j2];
  }
  __syncthreads();
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256) stft_cac_kernel(const float* __restrict__ wav, int L, int Tf, float4* __restrict__ Z,
                                                        double* __restrict__ stats, const float2* __restrict__ tw,
                                                        const float* __restrict__ win) {
  pdl_begin();
  __shared__ float2 sc[FFT_SMEM];
  __shared__ float red[2][8];
  const int t = threadIdx.x, frame = blockIdx.x, b = blockIdx.y;
  const float* wl = wav + (long)b * 2 * L;
  const float* wr = wl + L;
  float2 v[16];
  const int base = frame * 1024 - 1536;        // kept frame `frame` of _spec == stft frame frame+2
  if (base >= 0 && base + FFT_N <= L) {        // interior frame (all but 2 + 3 per segment): no reflection, constant offsets
    const float* pl = wl + base + t;
    const float* pr = wr + base + t;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float w = win[256 * n1 + t];
      v[n1] = make_float2(pl[256 * n1] * w, pr[256 * n1] * w);
    }
  } else {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      int n = 256 * n1 + t;
      int idx = base + n;
      if (idx < 0) idx = -idx;
      if (idx >= L) idx = 2 * (L - 1) - idx;
      float w = win[n];
      v[n1] = make_float2(wl[idx] * w, wr[idx] * w);
    }
  }
  dft16<false>(v);
  {
    float2 w[16];
    twiddle_powers<false>(tw, t, w);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], w[k1]);
  }
  fft4096_tail<false>(v, sc, tw);
  float s = 0.f, ss = 0.f;
  float4* zo = Z + ((long)b * Tf + frame) * 2048;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int k = t + 256 * i;
    int kn = (FFT_N - k) & (FFT_N - 1);
    const float2 A = sc[k + (k >> 4)], Bc = sc[kn + (kn >> 4)];
    const float ar = A.x, ai = A.y, br = Bc.x, bi = Bc.y;
    float4 o;
    o.x = (ar + br) * (0.5f / 64.f);       // L.re
    o.y = (ai - bi) * (0.5f / 64.f);       // L.im   (exactly +0 at k = 0)
    o.z = (ai + bi) * (0.5f / 64.f);       // R.re
    o.w = (br - ar) * (0.5f / 64.f);       // R.im   (exactly +0 at k = 0)
    zo[k] = o;
    s += o.x + o.y + o.z + o.w;
    ss += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
  }
  s = warp_sum(s); ss = warp_sum(ss);
  if ((t & 31) == 0) { red[0][t >> 5] = s; red[1][t >> 5] = ss; }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
    atomicAdd(stats + 2 * b, a); atomicAdd(stats + 2 * b + 1, c);
  }
}

void launch_stft_cac(const float* wav, int B, int L, int Tf, float* Z, double* stats, const float2* tw, const float* win,
                     cudaStream_t st) {
  launch_pdl(stft_cac_kernel, dim3(dim3(Tf, B)), dim3(256), 0, st, wav, L, Tf, (float4*)Z, stats, tw, win);
}

// ------------------------------------------------------------------ inverse: mask + iFFT -> windowed frames
// `scale` = (float)in / (float)out, hoisted by the caller (one division per kernel instead of one per element and frame)
__device__ __forceinline__ void lerp_coords_f(int d, int in, int out, float scale, int& i0, int& i1, float& lam) {
  if (in == out) { i0 = d; i1 = d; lam = 0.f; return; }
  float src = scale * ((float)d + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  lam = src - (float)i0;
}

// Front end of the inverse transform for one (output batch bo, frame): reads the spectrogram row Z[bz, frame], applies the
// freq_out 1x1 conv + 259 -> 2048 resize + sigmoid mask with the signed-CaC phase product (use_mask) or nothing (plain _ispec
// of channels (L, R): the STFT -> iSTFT microbenchmark), and writes the Hermitian-extended 4096-point spectrum of the packed
// signal L + iR into sr / si (natural order, un-padded indices).  Ends with __syncthreads().
template <typename T>
__device__ __forceinline__ void masked_spectrum_to_smem(const float4* __restrict__ zi, const T* __restrict__ dec, const RowSpace& ds,
                                                        int g, int use_mask, const float (&fw)[8], const float (&fb)[2],
                                                        float lerp_scale, float2* sc) {
  const int t = threadIdx.x;
  const T* dg = use_mask ? dec + ds.row_off(g, 0) : nullptr;      // rows of one group are contiguous: row r at dg + r * C
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int k = t + 256 * i;
    float4 z = zi[k];
    float2 xl, xr;
    if (use_mask) {
      int i0, i1; float lam;
      lerp_coords_f(k, ds.R, 2048, lerp_scale, i0, i1, lam);
      float a0, a1, a2, a3, c0, c1, c2, c3;
      if constexpr (sizeof(T) == 2) {      // the 4 decoder channels of a row are one 8-byte load
        const uint2 ra = *(const uint2*)(dg + i0 * ds.C), rc = *(const uint2*)(dg + i1 * ds.C);
        const float2 ra0 = __bfloat1622float2(*(const __nv_bfloat162*)&ra.x), ra1 = __bfloat1622float2(*(const __nv_bfloat162*)&ra.y);
        const float2 rc0 = __bfloat1622float2(*(const __nv_bfloat162*)&rc.x), rc1 = __bfloat1622float2(*(const __nv_bfloat162*)&rc.y);
        a0 = ra0.x; a1 = ra0.y; a2 = ra1.x; a3 = ra1.y; c0 = rc0.x; c1 = rc0.y; c2 = rc1.x; c3 = rc1.y;
      } else {
        const float4 ra = *(const float4*)(dg + i0 * ds.C), rc = *(const float4*)(dg + i1 * ds.C);
        a0 = ra.x; a1 = ra.y; a2 = ra.z; a3 = ra.w; c0 = rc.x; c1 = rc.y; c2 = rc.z; c3 = rc.w;
      }
      float l0a = fw[0] * a0 + fw[1] * a1 + fw[2] * a2 + fw[3] * a3 + fb[0];
      float l1a = fw[4] * a0 + fw[5] * a1 + fw[6] * a2 + fw[7] * a3 + fb[1];
      float l0b = fw[0] * c0 + fw[1] * c1 + fw[2] * c2 + fw[3] * c3 + fb[0];
      float l1b = fw[4] * c0 + fw[5] * c1 + fw[6] * c2 + fw[7] * c3 + fb[1];
      const float q0 = (1.f - lam) * l0a + lam * l0b, q1 = (1.f - lam) * l1a + lam * l1b;
      float m0, m1, inv0, inv1;
      if constexpr (sizeof(T) == 2) {      // bf16 build: MUFU exp / reciprocal (relative error ~1e-6, far below the bf16 activations)
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m0) : "f"(1.0f + __expf(-q0)));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m1) : "f"(1.0f + __expf(-q1)));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv0) : "f"(z.x + 1e-8f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(z.y + 1e-8f));
      } else {
        m0 = sigmoid_acc(q0); m1 = sigmoid_acc(q1);
        inv0 = 1.0f / (z.x + 1e-8f); inv1 = 1.0f / (z.y + 1e-8f);
      }
      // quirk Q4 (ATHTDemucs_v2.py:303-309): "magnitudes" are the signed CaC planes L.re and L.im
      float ms0 = z.x * m0, ms1 = z.y * m1;
      xl = make_float2(ms0 * (z.x * inv0), ms0 * (z.y * inv0));
      xr = make_float2(ms1 * (z.z * inv1), ms1 * (z.w * inv1));
    } else {
      xl = make_float2(z.x, z.y); xr = make_float2(z.z, z.w);
    }
    if (k == 0) { xl.y = 0.f; xr.y = 0.f; }        // C2R ignores the imaginary part of DC
    // Y[k] = XL[k] + i XR[k] ; Y[N-k] = conj(XL[k]) + i conj(XR[k])
    sc[k] = make_float2(xl.x - xr.y, xl.y + xr.x);
    if (k > 0) sc[FFT_N - k] = make_float2(xl.x + xr.y, xr.x - xl.y);
  }
  if (t == 0) sc[2048] = make_float2(0.f, 0.f);   // Nyquist bin is zero-padded by _ispec
  __syncthreads();
}

// Inverse 4096-point transform of the spectrum masked_spectrum_to_smem left in sc; on exit sc[n + (n >> 4)] holds the
// (left, right) time samples n of the frame (before the window and the 1/64 scale).
__device__ __forceinline__ void ifft4096_smem(float2* sc, const float2* __restrict__ tw) {
  const int t = threadIdx.x;
  float2 v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = sc[256 * n1 + t];
  dft16<true>(v);
  {
    float2 w[16];
    twiddle_powers<true>(tw, t, w);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], w[k1]);
  }
  __syncthreads();
  fft4096_tail<true>(v, sc, tw);
}

// ------------------------------------------------------------------ inverse: mask + iFFT + window + overlap-add + time branch
// htdemucs._ispec (ATHTDemucs_v2.py:310; demucs spec.py ispectro) fused with everything around it:
//   out[bo, ch, s] = (sum_t w[n] * irfft(64 Z_t)[n], n = s + 1536 - 1024 t) / 1.5  +  (time_out(tdec)[ch] * std_t + mean_t)
// (the window-sum envelope is the constant 1.5 on the kept samples; time branch: ATHTDemucs_v2.py:313-324).
// The windowed frames never go to memory: a CTA walks a contiguous run of frames of the flat (output batch, frame) list with
// a four-hop accumulator ring in shared memory -- the frame t adds its four quarters to the hops t .. t+3 and completes hop
// t, which is scaled, combined with the time branch and stored.  A run that starts inside a segment first re-transforms the
// three frames before it (ring warm-up, no stores); runs are equal-sized so that one wave of CTAs fills the chip.
// Every output sample is the fixed-order sum ((f[t-3] + f[t-2]) + f[t-1]) + f[t] of its (up to) four frames.
#define IFFT_SMEM_BYTES ((FFT_SMEM + FFT_N) * 8)

template <typename T>
__global__ void __launch_bounds__(256, 3) istft_fused_kernel(const float4* __restrict__ Z, int Tf, int L, int Bout, int zb_div,
                                                           const T* __restrict__ dec, RowSpace ds, int use_mask,
                                                           const float* __restrict__ fo_w, const float* __restrict__ fo_b,
                                                           const T* __restrict__ tdec, RowSpace ts, const float* __restrict__ to_w,
                                                           const float* __restrict__ to_b, const float* __restrict__ meanstd_t,
                                                           int ms_div, float* __restrict__ out, long out_bstride,
                                                           const float2* __restrict__ tw, const float* __restrict__ win) {
  pdl_begin();
  extern __shared__ __align__(16) float2 fsm[];
  float2* sc = fsm;
  float2* ring = fsm + FFT_SMEM;                     // [4 hops][1024] (left, right) partial overlap-add sums
  const int t = threadIdx.x;
  float fw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, fb[2] = {0, 0}, tw8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tb[2] = {0, 0};
  if (use_mask) {
#pragma unroll
    for (int i = 0; i < 8; ++i) fw[i] = fo_w[i];
    fb[0] = fo_b[0]; fb[1] = fo_b[1];
  }
  if (tdec) {
#pragma unroll
    for (int i = 0; i < 8; ++i) tw8[i] = to_w[i];
    tb[0] = to_b[0]; tb[1] = to_b[1];
  }
  const float lerp_scale = use_mask ? (float)ds.R / 2048.0f : 1.0f;
  const long total = (long)Bout * Tf;
  long f_begin = total * blockIdx.x / gridDim.x, f_end = total * (blockIdx.x + 1) / gridDim.x;
  const bool vec_ok = (L & 3) == 0 && (out_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;

  while (f_begin < f_end) {
    const int bo = (int)(f_begin / Tf);
    const int ta = (int)(f_begin - (long)bo * Tf);
    const int tb_ = (int)min((long)Tf, (long)ta + (f_end - f_begin));      // piece [ta, tb_) of segment bo
    const int bz = bo / zb_div;
    float mean = 0.f, sd = 1.f;
    if (tdec) { mean = meanstd_t[2 * (bo / ms_div)]; sd = meanstd_t[2 * (bo / ms_div) + 1]; }
    const int t_start = max(0, ta - 3);
    float* ob = out + (long)bo * out_bstride;

    auto emit = [&](int h) {            // hop h of the overlap-add buffer is complete: positions p = 1024 h + j, sample s = p - 1536
      const float2* rg = ring + (h & 3) * 1024;
      const int s0 = 1024 * h - 1536 + 4 * t;
      if (s0 + 3 < 0 || s0 >= L) return;
      const float4 q0 = *(const float4*)(rg + 4 * t), q1 = *(const float4*)(rg + 4 * t + 2);
      float a0[4] = {q0.x, q0.z, q1.x, q1.z}, a1[4] = {q0.y, q0.w, q1.y, q1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { a0[e] *= (1.0f / 1.5f); a1[e] *= (1.0f / 1.5f); }
      if (tdec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int s = s0 + e;
          if (s < 0 || s >= L) continue;
          const T* r = tdec + ts.row_off(bo, s);
          float x0, x1, x2, x3;
          if constexpr (sizeof(T) == 2) {
            const uint2 rr = *(const uint2*)r;
            const float2 r0 = __bfloat1622float2(*(const __nv_bfloat162*)&rr.x), r1 = __bfloat1622float2(*(const __nv_bfloat162*)&rr.y);
            x0 = r0.x; x1 = r0.y; x2 = r1.x; x3 = r1.y;
          } else {
            const float4 rr = *(const float4*)r;
            x0 = rr.x; x1 = rr.y; x2 = rr.z; x3 = rr.w;
          }
          const float y0 = tw8[0] * x0 + tw8[1] * x1 + tw8[2] * x2 + tw8[3] * x3 + tb[0];
          const float y1 = tw8[4] * x0 + tw8[5] * x1 + tw8[6] * x2 + tw8[7] * x3 + tb[1];
          a0[e] += y0 * sd + mean; a1[e] += y1 * sd + mean;
        }
      }
      if (vec_ok && s0 >= 0 && s0 + 3 < L) {
        *(float4*)(ob + s0) = make_float4(a0[0], a0[1], a0[2], a0[3]);
        *(float4*)(ob + L + s0) = make_float4(a1[0], a1[1], a1[2], a1[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int s = s0 + e;
          if (s >= 0 && s < L) { ob[s] = a0[e]; ob[L + s] = a1[e]; }
        }
      }
    };

    for (int fr = t_start; fr < tb_; ++fr) {
      masked_spectrum_to_smem<T>(Z + ((long)bz * Tf + fr) * 2048, dec, ds, bo * Tf + fr, use_mask, fw, fb, lerp_scale, sc);
      ifft4096_smem(sc, tw);
      // quarter qq of the frame belongs to hop fr + qq; the first frame of the run and every quarter 3 open a new hop
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = t + 256 * i, qq = i >> 2;
        const float w = win[n] * (1.0f / 64.0f);
        const float2 y2 = sc[n + (n >> 4)];
        const float2 x = make_float2(y2.x * w, y2.y * w);
        float2* slot = ring + ((fr + qq) & 3) * 1024 + (n & 1023);
        if (qq == 3 || fr == t_start) *slot = x;
        else { float2 a = *slot; a.x += x.x; a.y += x.y; *slot = a; }
      }
      __syncthreads();
      if (fr >= ta) emit(fr);
      // no barrier needed here: the next frame's spectrum overwrites sr / si, which were last read BEFORE the barrier above,
      // and its quarter 3 overwrites the ring slot emit() is reading only after the barriers inside the next transform
    }
    if (tb_ == Tf)                       // end of the segment: hops Tf .. Tf + 2 only have older frames left
      for (int h = Tf; h < Tf + 3; ++h) emit(h);
    f_begin += tb_ - ta;
  }
}

template <typename T>
void launch_istft_fused(const float* Z, int Tf, int L, int Bout, int zb_div, const T* dec, RowSpace ds, int use_mask,
                        const float* fo_w, const float* fo_b, const T* tdec, RowSpace ts, const float* to_w, const float* to_b,
                        const float* meanstd_t, int ms_div, float* out, long out_bstride, const float2* tw, const float* win,
                        cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first()) cudaFuncSetAttribute(istft_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, IFFT_SMEM_BYTES);
  // equal runs of the flat (batch, frame) list: one wave of 3 CTAs per SM, at least ~12 frames per run (3 warm-up frames)
  const long total = (long)Bout * Tf;
  const long slots = 3L * device_sm_count();
  long grid = std::min(slots, std::max(1L, total / 12));
  launch_pdl(istft_fused_kernel<T>, dim3((unsigned)grid), dim3(256), IFFT_SMEM_BYTES, st, (const float4*)Z, Tf, L, Bout, zb_div, dec, ds, use_mask, fo_w,
                                                                     fo_b, tdec, ts, to_w, to_b, meanstd_t, ms_div, out, out_bstride,
                                                                     tw, win);
}

template void launch_istft_fused<float>(const float*, int, int, int, int, const float*, RowSpace, int, const float*, const float*,
                                        const float*, RowSpace, const float*, const float*, const float*, int, float*, long,
                                        const float2*, const float*, cudaStream_t);
template void launch_istft_fused<bf16>(const float*, int, int, int, int, const bf16*, RowSpace, int, const float*, const float*,
                                       const bf16*, RowSpace, const float*, const float*, const float*, int, float*, long,
                                       const float2*, const float*, cudaStream_t);

}  // namespace athtd
