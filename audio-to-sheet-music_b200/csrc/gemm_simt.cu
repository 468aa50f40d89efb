// SIMT (CUDA-core, fp32 accumulate) implementation of the generic conv-as-GEMM problem.
// It is the fp32 parity path and the fallback for shapes the tcgen05 kernel does not take
// (tiny K / N layers that are HBM-bound anyway).
#include "gemm.cuh"
#include <string.h>

namespace athtd {

template <typename T, typename TC, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc d) {
  constexpr int BK = 16;
  constexpr int NTX = BN / TN;
  constexpr int NTY = BM / TM;
  static_assert(NTX * NTY == 256, "256 threads");
  constexpr int AROWS = BM / 16;          // A rows loaded per thread
  constexpr int BPER = BN * BK / 256;     // B elements loaded per thread
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float red_s[BM], red_ss[BM];
  __shared__ int red_g[BM];

  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const T* __restrict__ A = (const T*)d.A;
  const T* __restrict__ Bp = (const T*)d.B;
  const int n0 = blockIdx.y * BN;
  const long Mtot = (long)d.G1 * d.G2 * d.Mg;

  // ---- row decode for the A loads
  const int a_k = tid % BK;
  long a_off[AROWS];
  int a_g2[AROWS];
#pragma unroll
  for (int i = 0; i < AROWS; ++i) {
    int r = tid / BK + 16 * i;
    int g, m;
    bool ok;
    if (d.grouped) {
      g = blockIdx.z; m = blockIdx.x * BM + r; ok = m < d.Mg;
    } else {
      long rid = (long)blockIdx.x * BM + r;
      ok = rid < Mtot;
      g = (int)(rid / d.Mg); m = (int)(rid - (long)g * d.Mg);
    }
    int g1 = g / d.G2, g2 = g - g1 * d.G2;
    a_off[i] = (long)g1 * d.sAg1 + (long)g2 * d.sAg2 + (long)m * d.sAm;
    a_g2[i] = ok ? g2 : -(1 << 28);
  }
  long b_goff = 0;
  if (d.grouped) {
    int g = blockIdx.z; int g1 = g / d.G2, g2 = g - g1 * d.G2;
    b_goff = (long)g1 * d.sBg1 + (long)g2 * d.sBg2;
  }
  const bool b_kfast = (d.sBk == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d.K; k0 += BK) {
    {  // A tile
      int kk = k0 + a_k;
      int tap = kk / d.Ktap;
      int kin = kk - tap * d.Ktap;
      bool kok = kk < d.K;
      int sg2 = 0; long toff = 0;
      if (kok) { sg2 = d.shG2[tap]; toff = (long)sg2 * d.sAg2 + (long)d.shM[tap] * d.sAm + kin; }
#pragma unroll
      for (int i = 0; i < AROWS; ++i) {
        int g2 = a_g2[i] + sg2;
        bool ok = kok && g2 >= 0 && g2 < d.G2;
        float v = 0.f;
        if (ok) v = to_f<T>(A[a_off[i] + toff]);
        As[a_k][tid / BK + 16 * i] = v;
      }
    }
    {  // B tile
#pragma unroll
      for (int j = 0; j < BPER; ++j) {
        int bn, bk;
        if (b_kfast) { bk = tid % BK; bn = tid / BK + 16 * j; }
        else { bn = tid % BN; bk = tid / BN + (256 / BN) * j; }
        int n = n0 + bn, kk = k0 + bk;
        float v = 0.f;
        if (n < d.N && kk < d.K) v = to_f<T>(Bp[b_goff + (long)n * d.sBn + (long)kk * d.sBk]);
        Bs[bk][bn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const int Nout = d.glu ? d.N / 2 : d.N;
  TC* __restrict__ C = (TC*)d.C;
  const TC* __restrict__ R = (const TC*)d.res;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int r = ty * TM + i;
    int g, m; bool ok;
    if (d.grouped) { g = blockIdx.z; m = blockIdx.x * BM + r; ok = m < d.Mg; }
    else {
      long rid = (long)blockIdx.x * BM + r; ok = rid < Mtot;
      g = (int)(rid / d.Mg); m = (int)(rid - (long)g * d.Mg);
    }
    int g1 = g / d.G2, g2 = g - g1 * d.G2;
    float s = 0.f, ss = 0.f;
    if (ok) {
      long coff = (long)g1 * d.sCg1 + (long)g2 * d.sCg2 + (long)m * d.sCm;
      long roff = (long)g1 * d.sRg1 + (long)g2 * d.sRg2 + (long)m * d.sRm;
      float v[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int n = n0 + tx * TN + j;
        float x = d.alpha * acc[i][j];
        if (n < d.N) {
          if (d.bias) x += d.bias[n];
          if (d.gbias) x += d.gbias[(long)g1 * d.sGb + n];
          if (d.act == ACT_GELU) x = gelu_erf(x);
        }
        v[j] = x;
      }
      if (d.glu) {
#pragma unroll
        for (int j = 0; j < TN; j += 2) {
          int n = n0 + tx * TN + j;
          if (n + 1 < d.N) {
            int no = n >> 1;
            float x = v[j] * sigmoid_acc(v[j + 1]);
            if (d.colscale) x *= d.colscale[no];
            if (d.rowtab) x += d.rowtab_scale * d.rowtab[(long)m * Nout + no];
            if (R) x += to_f<TC>(R[roff + no]);
            s += x; ss += x * x;
            C[coff + no] = from_f<TC>(x);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          int n = n0 + tx * TN + j;
          if (n < d.N) {
            float x = v[j];
            if (d.colscale) x *= d.colscale[n];
            if (d.rowtab) x += d.rowtab_scale * d.rowtab[(long)m * Nout + n];
            if (R) x += to_f<TC>(R[roff + n]);
            bool counted = true;
            if (d.convt_cout > 0) {   // rows -2,-1 (q==0, phases 0,1) and 4F,4F+1 (q==Mg-1, phases 2,3) are cropped
              int phase = n / d.convt_cout;
              if ((m == 0 && phase < 2) || (m == d.Mg - 1 && phase >= 2)) counted = false;
            }
            if (counted) { s += x; ss += x * x; }
            C[coff + n] = from_f<TC>(x);
          }
        }
      }
    }
    if (d.stat_mode != STAT_NONE) {
#pragma unroll
      for (int o = NTX / 2; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
      }
      if (tx == 0) {
        if (d.stat_mode == STAT_PER_G1_M) {
          if (ok) {
            double* st = d.stats + 2 * ((long)g1 * d.Mg + m);
            atomicAdd(st, (double)s); atomicAdd(st + 1, (double)ss);
          }
        } else {
          red_s[r] = s; red_ss[r] = ss; red_g[r] = ok ? g1 : -1;
        }
      }
    }
  }
  if (d.stat_mode == STAT_PER_G1) {
    __syncthreads();
    if (tid < 32) {
      int gfirst = red_g[0];
      double s = 0.0, ss = 0.0;
      for (int r = tid; r < BM; r += 32) {
        int g = red_g[r];
        if (g < 0) continue;
        if (g == gfirst) { s += red_s[r]; ss += red_ss[r]; }
        else {
          atomicAdd(d.stats + 2 * ((long)g * STAT_SLOTS + blockIdx.x % STAT_SLOTS), (double)red_s[r]);
          atomicAdd(d.stats + 2 * ((long)g * STAT_SLOTS + blockIdx.x % STAT_SLOTS) + 1, (double)red_ss[r]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
      }
      if (tid == 0 && gfirst >= 0) {
        atomicAdd(d.stats + 2 * ((long)gfirst * STAT_SLOTS + blockIdx.x % STAT_SLOTS), s);
        atomicAdd(d.stats + 2 * ((long)gfirst * STAT_SLOTS + blockIdx.x % STAT_SLOTS) + 1, ss);
      }
    }
  }
}

template <typename T, typename TC, int BM, int BN, int TM, int TN>
static void launch_cfg(const GemmDesc& d, cudaStream_t st) {
  dim3 grid;
  if (d.grouped) {
    grid = dim3((d.Mg + BM - 1) / BM, (d.N + BN - 1) / BN, d.G1 * d.G2);
  } else {
    long Mtot = (long)d.G1 * d.G2 * d.Mg;
    grid = dim3((unsigned)((Mtot + BM - 1) / BM), (d.N + BN - 1) / BN, 1);
  }
  gemm_simt_kernel<T, TC, BM, BN, TM, TN><<<grid, 256, 0, st>>>(d);
}

template <typename T, typename TC>
static void launch_by_shape(const GemmDesc& d, cudaStream_t st) {
  if (d.N <= 16) launch_cfg<T, TC, 128, 16, 4, 2>(d, st);
  else if (d.N <= 48 || (long)d.G1 * d.G2 * d.Mg <= 64) launch_cfg<T, TC, 64, 64, 4, 4>(d, st);
  else launch_cfg<T, TC, 128, 64, 8, 4>(d, st);
}

template <typename T>
void launch_gemm_simt(const GemmDesc& d, cudaStream_t st) {
  if (d.c_is_f32) launch_by_shape<T, float>(d, st);
  else launch_by_shape<T, T>(d, st);
}

template void launch_gemm_simt<float>(const GemmDesc&, cudaStream_t);
template void launch_gemm_simt<bf16>(const GemmDesc&, cudaStream_t);

}  // namespace athtd
