// Generic "conv-as-GEMM" problem descriptor shared by the SIMT (fp32-accurate) and the
// tcgen05 (bf16 tensor-core) GEMM kernels.
//
//   C[g1,g2,m,n] = epilogue( alpha * sum_{tap,k} A[g1, g2+shG2[tap], m+shM[tap], k] * B[g1,g2,n,tap*Ktap+k] )
//
// Row space: groups (g1 in [0,G1), g2 in [0,G2)), Mg rows per group.  A rows are addressed as
//   A + g1*sAg1 + g2*sAg2 + m*sAm + k   (k contiguous; sAm may be SMALLER than K: the strided
// convolutions and the transposed convolutions are plain GEMMs over overlapping row windows of a
// channels-last activation buffer, see DESIGN.md "conv as overlapping-row GEMM").
// A tap whose g2+shG2 falls outside [0,G2) contributes zeros (time-axis zero padding of DConv).
#pragma once
#include "common.cuh"

namespace athtd {

enum { ACT_NONE = 0, ACT_GELU = 1 };
enum { STAT_NONE = 0, STAT_PER_G1 = 1, STAT_PER_G1_M = 2 };
// per-segment statistics are spread over STAT_SLOTS accumulator slots (slot = CTA index mod STAT_SLOTS) so that the
// fp64 atomics of thousands of CTAs do not serialise on one L2 address; finalize_gn sums the slots.
static constexpr int STAT_SLOTS = 64;

struct GemmDesc {
  int G1, G2, Mg;
  int N, K;            // N = number of B rows (before GLU pairing); K = ntaps*Ktap
  int ntaps, Ktap;
  int shG2[3], shM[3];
  const void* A; long sAg1, sAg2, sAm;
  const void* B; long sBg1, sBg2, sBn, sBk;
  void* C; long sCg1, sCg2, sCm;
  int c_is_f32;        // C (and res) stored as fp32 regardless of the activation type
  float alpha;
  const float* bias;       // [N]
  int act;
  int glu;                 // adjacent column pairs (2j, 2j+1) -> a*sigmoid(b), N/2 outputs
  const float* colscale;   // [Nout]
  const void* res; long sRg1, sRg2, sRm;   // residual (same type as C), added last
  const float* rowtab; float rowtab_scale; // table[m][Nout] added (freq_emb), indexed by m
  const float* gbias; long sGb;            // per-g1 bias row: gbias[g1*sGb + n]
  double* stats; int stat_mode;            // accumulates (sum, sumsq) of the stored values
  int convt_cout;          // >0: transposed-conv phase layout, mask the two cropped rows per edge
  int grouped;             // 1: tiles never span groups (B or stats differ per group)
};

inline GemmDesc gemm_desc_zero() {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.G1 = d.G2 = 1; d.ntaps = 1; d.alpha = 1.0f; d.sBk = 1;
  return d;
}

template <typename T> void launch_gemm_simt(const GemmDesc& d, cudaStream_t st);

// Flat-row form of the same problem for the tcgen05 kernel (gemm_tc.cu): A is a 2-D [a_rows, Ktap] bf16 view with
// row pitch a_pitch elements, taps shift the row index; the epilogue decodes a flat row rho into
// (b, t', f') = (rho / RpA / G2p, rho / RpA % G2p, rho % RpA), keeps it iff f' in [vlo,vhi) and t' in [gpf,gpf+G2),
// and stores it at row ((b*oG2p + t' + ogsh)*oRp + f' + orsh) of C (row pitch ldc).
struct TcFlat {
  const void* A; long a_rows; long a_pitch; int Ktap; int ntaps; int tapRow[3];
  const void* B; int N;
  long Mflat;
  int RpA, G2p, gpf, G2, vlo, vhi;
  int oG2p, ogsh, oRp, orsh; long ldc; void* C; int c_is_f32;
  float alpha; const float* bias; int act; int glu; const float* colscale; const void* res;
  const float* rowtab; float rowtab_scale; double* stats; int stat_mode; int statR; int convt_cout;
  int n_store;          // store only the first n_store output columns (0 = all): zero-padded hidden channels
  int no_store;         // statistics-only pass (the GroupNorm'd result is recomputed by a second pass instead of stored)
  const float* gn_mr; const float* gn_w; const float* gn_b; int gn_mode;   // GroupNorm apply right after the bias
  void* dbg;            // unused (kept for the micro-benchmark ABI)
  int skip_lo, skip_hi; // output columns [skip_lo, skip_hi) feed the statistics but are not stored
  // fused finalisation of per-segment statistics (stat_mode == STAT_PER_G1): the CTA that finishes last (ticket on *fin_counter, zeroed
  // with the statistics) turns the fin_n x STAT_SLOTS slot sums into (mean, rstd) pairs in fin_mr -- no finalize launch
  float* fin_mr; double fin_count; int fin_n; unsigned* fin_counter;
};
void tc_set_bn_cap(int cap);   // 128 or 256: largest N tile (A/B tuning knob)
bool tensor_map_api_available();              // cuTensorMapEncodeTiled reachable through the runtime's driver entry point
bool tc_flat_supported(const TcFlat& f);
int launch_gemm_tc_flat(const TcFlat& f, cudaStream_t st);   // bf16 in, fp32 accumulate; 0 = launched
bool tc_flat_fuses_finalize(const TcFlat& f);                // whether that launch also writes f.fin_mr

}  // namespace athtd
