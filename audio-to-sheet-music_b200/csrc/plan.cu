// Host-side launch sequence of the segment-batched separation forward (one CUDA stream, no
// allocation, no synchronisation): replaces AudioTextHTDemucs.forward
// (/root/reference/src/models/stem_separation/ATHTDemucs_v2.py:250-326) for a batch of B segments
// and P prompts per segment.  The prompt-independent part (STFT, normalisation, both encoder
// branches, cross-transformer) runs once; text conditioning, both decoders, the mask/iSTFT tail
// and the time-branch sum run once per prompt over the shared encoder state.
#include "plan.h"
#include <string.h>
#include <stdio.h>

namespace athtd {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------- packed blob layout
PackLayout::PackLayout(int dtype) {
  items = build_pack_list();
  size_t esz = dtype == 0 ? 4 : 2;
  size_t off = 0;
  for (auto& it : items) {
    off = align_up(off, 256);
    it.offset_bytes = (long)off;
    off += (size_t)it.numel * (it.is_f32 ? 4 : esz);
    index[it.key] = (int)(&it - &items[0]);
  }
  total_bytes = (long)align_up(off, 256);
}

template <typename T>
int pack_weights(const ParamTable& pt, const PackLayout& pl, const float* params, void* packed, cudaStream_t st) {
  for (auto& it : pl.items) {
    const float* src = params + pt.off(it.src) + it.src_off;
    char* dst = (char*)packed + it.offset_bytes;
    if (it.kind >= 8) {      // derived from the weight and its bias
      std::string bn = it.src;
      bn.replace(bn.rfind("weight"), 6, "bias");
      launch_pack_gram<T>(src, params + pt.off(bn), dst, it.kind, it.d0, it.d1, it.d2, st);
      continue;
    }
    if (it.is_f32) launch_pack_weight<float>(src, (float*)dst, it.numel, it.kind, it.d0, it.d1, it.d2, st);
    else launch_pack_weight<T>(src, (T*)dst, it.numel, it.kind, it.d0, it.d1, it.d2, st);
  }
  return (int)cudaGetLastError();
}
template int pack_weights<float>(const ParamTable&, const PackLayout&, const float*, void*, cudaStream_t);
template int pack_weights<bf16>(const ParamTable&, const PackLayout&, const float*, void*, cudaStream_t);

// ---------------------------------------------------------------------------------- shapes
Shapes::Shapes(int B_, int L_, int P_) : B(B_), L(L_), P(P_) {
  Tf = (L + 1023) / 1024;
  Fr[0] = 2048; for (int i = 1; i <= 4; ++i) Fr[i] = Fr[i - 1] / 4;
  Lt[0] = L; for (int i = 1; i <= 4; ++i) Lt[i] = (Lt[i - 1] + 3) / 4;
  Sf = Tf * 8; St = Lt[4];
}

// ---------------------------------------------------------------------------------- plan
static RowSpace make_space(int B, int G2, int R, int C, bool padded) {
  RowSpace r;
  r.G = B * G2; r.G2 = G2; r.R = R; r.C = C;
  if (padded) {
    r.pf = 2;
    r.Rp = G2 > 1 ? R + 4 : (R + 8 + 3) / 4 * 4;   // multiple of 4: the k8/s4 conv reads the buffer as rows of 4*C
    r.G2p = G2 > 1 ? G2 + 4 : 1;                    // 2 zero frames each side: DConv taps along time read zeros
    r.gpf = G2 > 1 ? 2 : 0;
  } else {
    r.pf = 0; r.Rp = R; r.G2p = G2; r.gpf = 0;
  }
  return r;
}

// Row spaces for a batch of B segments.  Segment b of every buffer occupies [b, b + 1) * g1_stride(), whatever B is, so a plan
// laid out for `cap` segments runs any batch B <= cap in place (pads stay where they are).
template <typename T>
void PlanT<T>::spaces(int B) {
  const Shapes& s = sh;
  const int Tf = s.Tf;
  xf0_rs = make_space(B, Tf, 2048, 4, true);
  xt0_rs = make_space(B, 1, s.L, 2, true);
  for (int i = 0; i < 4; ++i) {
    yf_rs[i] = make_space(B, Tf, s.Fr[i + 1], kCh[i], true);
    yt_rs[i] = make_space(B, 1, s.Lt[i + 1], kCh[i], true);
  }
  xc_rs = make_space(B, Tf, 8, 384, true);
  xtc_rs = make_space(B, 1, s.St, 384, true);
  for (int i = 0; i < 4; ++i) {
    df_rs[i] = make_space(B, Tf, Tf, kDecCh[i + 1], true);
    dt_rs[i] = make_space(B, 1, s.Lt[3 - i], kDecCh[i + 1], true);
  }
}

template <typename T>
int PlanT<T>::set_batch(int B) {
  if (B < 1 || B > cap) return 1;
  sh.B = B;
  spaces(B);
  return 0;
}

template <typename T>
void PlanT<T>::layout(char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* {
    off = align_up(off, 256);
    char* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const Shapes& s = sh;
  const int B = cap, Tf = s.Tf;      // every size below is for the batch CAPACITY
  spaces(cap);
  // ---- zero-initialised region (pads must stay zero; interior is always overwritten)
  xf0 = (T*)take(xf0_rs.elems() * sizeof(T));
  xt0 = (T*)take(xt0_rs.elems() * sizeof(T));
  for (int i = 0; i < 4; ++i) {
    yf[i] = (T*)take(yf_rs[i].elems() * sizeof(T));
    ef[i] = (T*)take(yf_rs[i].elems() * sizeof(T));
    yt[i] = (T*)take(yt_rs[i].elems() * sizeof(T));
    et[i] = (T*)take(yt_rs[i].elems() * sizeof(T));
  }
  xc = (T*)take(xc_rs.elems() * sizeof(T));
  xtc = (T*)take(xtc_rs.elems() * sizeof(T));
  for (int i = 0; i < 4; ++i) {
    df[i] = (T*)take(df_rs[i].elems() * sizeof(T));
    dt[i] = (T*)take(dt_rs[i].elems() * sizeof(T));
  }
  zero_bytes = align_up(off, 256);
  // ---- statistics (zeroed every forward)
  off = zero_bytes;
  stats_begin = off;
  st_spec = (double*)take(sizeof(double) * 2 * B);
  st_wav = (double*)take(sizeof(double) * 2 * B);
  for (int i = 0; i < 4; ++i)
    for (int d = 0; d < 2; ++d)
      for (int k = 0; k < 2; ++k) {
        st_df[i][d][k] = (double*)take(sizeof(double) * 2 * B * s.Fr[i + 1]);
        st_dt[i][d][k] = (double*)take(sizeof(double) * 2 * B * STAT_SLOTS);
      }
  for (int l = 0; l < 5; ++l) { st_xf[l][0] = (double*)take(sizeof(double) * 2 * B * STAT_SLOTS); st_xf[l][1] = (double*)take(sizeof(double) * 2 * B * STAT_SLOTS); }
  fin_xf = (unsigned*)take(sizeof(unsigned) * 16);
  // decoder GroupNorm accumulators: LAST in the statistics block -- encode() clears [stats_begin, dec_stats_begin),
  // decode() clears [dec_stats_begin, stats end) so that one encode can be followed by any number of decodes
  off = align_up(off, 256);
  dec_stats_begin = off;
  st_dec = (double*)take(sizeof(double) * 2 * B * 6 * s.P * STAT_SLOTS);
  fin_dec = (unsigned*)take(sizeof(unsigned) * 6 * s.P);
  stats_bytes = align_up(off, 256) - stats_begin;
  off = stats_begin + stats_bytes;
  // ---- plain scratch
  mr = (float*)take(sizeof(float) * 2 * (size_t)B * 512);
  ms_spec = (float*)take(sizeof(float) * 2 * B);
  ms_wav = (float*)take(sizeof(float) * 2 * B);
  Z = (float*)take(sizeof(float) * (size_t)B * Tf * 2048 * 4);
  size_t hmax = 0, emax = 0, umax = 0;
  for (int i = 0; i < 4; ++i) {
    size_t rows = std::max((size_t)yf_rs[i].rows_total(), (size_t)yt_rs[i].rows_total());
    hmax = std::max(hmax, rows * (size_t)hidden_pad(kCh[i] / 8));
    emax = std::max(emax, rows * (2 * kCh[i]));
  }
  hbuf = (T*)take(hmax * sizeof(T));
  ebuf = (T*)take(emax * sizeof(T));
  const size_t Smax = std::max(s.Sf, s.St);
  tokf = (T*)take((size_t)B * s.Sf * 512 * sizeof(T));
  tokt = (T*)take((size_t)B * s.St * 512 * sizeof(T));
  for (int i = 0; i < 5; ++i) hn[i] = (T*)take((size_t)B * Smax * 512 * sizeof(T));
  qkv = (T*)take((size_t)B * Smax * 1536 * sizeof(T));
  kvb = (T*)take((size_t)B * Smax * 1024 * sizeof(T));
  obuf = (T*)take((size_t)B * Smax * 512 * sizeof(T));
  ffn = (T*)take((size_t)B * Smax * 2048 * sizeof(T));
  scores = (T*)take((size_t)B * 8 * Smax * Smax * sizeof(T));
  xenc = (T*)take((size_t)B * s.Sf * 384 * sizeof(T));
  xtenc = (T*)take((size_t)B * s.St * 384 * sizeof(T));
  cvec = (float*)take(sizeof(float) * (size_t)B * s.P * 384 * 3);
  t1 = (T*)take((size_t)B * Smax * 384 * sizeof(T));
  t2 = (T*)take((size_t)B * Smax * 384 * sizeof(T));
  // transposed-conv outputs in phase layout share the geometry of the layer INPUT with 4*Cout channels
  umax = std::max(umax, (size_t)xc_rs.rows_total() * 4 * kDecCh[1]);
  umax = std::max(umax, (size_t)xtc_rs.rows_total() * 4 * kDecCh[1]);
  for (int i = 1; i < 4; ++i) {
    umax = std::max(umax, (size_t)df_rs[i - 1].rows_total() * 4 * kDecCh[i + 1]);
    umax = std::max(umax, (size_t)dt_rs[i - 1].rows_total() * 4 * kDecCh[i + 1]);
  }
  ubuf = (T*)take(umax * sizeof(T));
  total_bytes = align_up(off, 256);
  spaces(sh.B);
}

template <typename T>
PlanT<T>::PlanT(int B, int L, int P, const ParamTable* pt_, const PackLayout* pl_, const float* params_, const void* packed_,
                void* workspace, const PlanConsts& c)
    : sh(B, L, P), pt(pt_), pl(pl_), params(params_), packed((const char*)packed_), consts(c), cap(B) {
  base_ = (char*)workspace;
  layout((char*)workspace);
}

template <typename T>
const float* PlanT<T>::P32(const std::string& name) const { return params + pt->off(name); }
template <typename T>
const T* PlanT<T>::PW(const std::string& key) const {
  auto it = pl->index.find(key);
  if (it == pl->index.end()) throw std::runtime_error("athtd: unknown packed weight " + key);
  return (const T*)(packed + pl->items[it->second].offset_bytes);
}
template <typename T>
const float* PlanT<T>::PA(const std::string& key) const {
  auto it = pl->index.find(key);
  if (it == pl->index.end()) throw std::runtime_error("athtd: unknown packed aux " + key);
  return (const float*)(packed + pl->items[it->second].offset_bytes);
}

template <typename T>
void PlanT<T>::prof_begin(double gflop, cudaStream_t st, const char* what, long m, int n, int k) {
  if (!profiling) return;      // per-launch CUDA-event timing of the GEMM kernels (bench.py roofline pass only)
  while (prof_ev.size() < prof_used + 2) { cudaEvent_t e; cudaEventCreate(&e); prof_ev.push_back(e); }
  prof_gflop += gflop;
  prof_items.push_back(gflop);
  char lab[96];
  snprintf(lab, sizeof(lab), "%-6s M %8ld N %5d K %5d", what, m, n, k);
  prof_labels.push_back(lab);
  cudaEventRecord(prof_ev[prof_used++], st);
}
template <typename T>
void PlanT<T>::prof_end(cudaStream_t st) {
  if (profiling) cudaEventRecord(prof_ev[prof_used++], st);
}

template <typename T>
void PlanT<T>::gemm(const GemmDesc& d, cudaStream_t st) {
  prof_begin(2.0 * (double)d.G1 * d.G2 * d.Mg * d.N * d.K * 1e-9, st, "simt", (long)d.G1 * d.G2 * d.Mg, d.N, d.K);
  launch_gemm_simt<T>(d, st);
  prof_end(st);
  ++n_launches;
}

template <typename T>
void PlanT<T>::get_profile(double* ms, double* gflop, int* n) {
  double total = 0.0;
  if (prof_used) cudaEventSynchronize(prof_ev[prof_used - 1]);
  for (size_t i = 0; i + 1 < prof_used; i += 2) { float t = 0.f; cudaEventElapsedTime(&t, prof_ev[i], prof_ev[i + 1]); total += t; }
  *ms = total; *gflop = prof_gflop; *n = (int)(prof_used / 2);
  if (getenv("ATHTD_PROFILE_DUMP")) {      // per-launch list for tuning: index, GFLOP, us, TFLOP/s
    for (size_t i = 0; i + 1 < prof_used; i += 2) {
      float t = 0.f; cudaEventElapsedTime(&t, prof_ev[i], prof_ev[i + 1]);
      fprintf(stderr, "gemm %3zu  %s  %9.3f GFLOP  %8.1f us  %7.1f TFLOP/s\n", i / 2, prof_labels[i / 2].c_str(), prof_items[i / 2], t * 1e3,
              prof_items[i / 2] / t);
    }
  }
}

// ---- lower one convolution-shaped op either to the tcgen05 flat-row kernel (bf16 build, supported shapes) or to
//      the SIMT kernel.  Both read the same buffers and produce the same layout.
template <typename T>
bool PlanT<T>::conv(const ConvOp<T>& o, cudaStream_t st) {
  const RowSpace& as = o.as;
  const RowSpace& cs = o.cs;
  const int C = as.C;
  const int B = as.batch();
  const bool freq = as.G2 > 1;
  const int K = o.mode == CONV_ROWS ? C : o.mode == CONV_K8S4 ? 8 * C : o.mode == CONV_T ? 2 * C : 3 * C;
  const int Mg = o.mode == CONV_K8S4 ? cs.R : o.mode == CONV_T ? as.R + 1 : as.R;
  const double gflop = 2.0 * (double)B * as.G2 * Mg * o.N * K * 1e-9;
  if (sizeof(T) == 2 && use_tc) {
    TcFlat f;
    memset(&f, 0, sizeof(f));
    f.B = o.w; f.N = o.N; f.alpha = 1.0f;
    f.G2p = as.G2p; f.gpf = as.gpf; f.G2 = as.G2;
    f.tapRow[0] = 0;
    if (o.mode == CONV_K8S4) {
      f.A = o.a; f.a_rows = as.rows_total() / 4; f.a_pitch = 4L * C; f.Ktap = 4 * C; f.ntaps = 2; f.tapRow[1] = 1;
      f.RpA = as.Rp / 4; f.vlo = 0; f.vhi = cs.R;
    } else {
      f.A = o.a; f.a_rows = as.rows_total(); f.a_pitch = C; f.Ktap = C; f.RpA = as.Rp;
      if (o.mode == CONV_ROWS) { f.ntaps = 1; f.vlo = as.pf; f.vhi = as.pf + as.R; }
      else if (o.mode == CONV_T) { f.ntaps = 2; f.tapRow[1] = 1; f.vlo = as.pf - 1; f.vhi = as.pf + as.R + 1; }
      else { f.ntaps = 3; const int sh_ = o.dil * (freq ? as.Rp : 1); f.tapRow[0] = -sh_; f.tapRow[1] = 0; f.tapRow[2] = sh_;
             f.vlo = as.pf; f.vhi = as.pf + as.R; }
    }
    f.Mflat = f.a_rows;
    f.oG2p = cs.G2p; f.ogsh = cs.gpf - as.gpf; f.oRp = cs.Rp; f.ldc = cs.C;
    f.orsh = (o.mode == CONV_T ? cs.pf - 1 : cs.pf) - f.vlo;   // transposed conv: row q sits at the row of x[q-1]
    f.C = o.c; f.c_is_f32 = 0; f.bias = o.bias; f.act = o.act; f.glu = o.glu; f.colscale = o.colscale; f.res = o.res;
    f.rowtab = o.rowtab; f.rowtab_scale = o.rowtab_scale; f.stats = o.stats; f.stat_mode = o.stat_mode; f.statR = as.R;
    f.convt_cout = o.mode == CONV_T ? o.N / 4 : 0;
    f.n_store = o.n_store; f.no_store = o.no_store; f.skip_lo = o.skip_lo; f.skip_hi = o.skip_hi; f.gn_mr = o.gn_mr; f.gn_w = o.gn_w; f.gn_b = o.gn_b; f.gn_mode = o.gn_mode;
    f.fin_mr = o.fin_mr; f.fin_count = o.fin_count; f.fin_n = B; f.fin_counter = o.fin_counter;
    const bool geom_ok = (o.mode != CONV_K8S4) || (as.Rp % 4 == 0 && as.pf == 2);
    if (geom_ok && tc_flat_supported(f)) {
      prof_begin(gflop, st, o.mode == CONV_ROWS ? "rows" : o.mode == CONV_K8S4 ? "k8s4" : o.mode == CONV_T ? "convT" : "k3", (long)B * as.G2 * Mg, o.N, K);
      int rc = launch_gemm_tc_flat(f, st);
      prof_end(st);
      if (rc != 0) throw std::runtime_error("athtd: cuTensorMapEncodeTiled failed for a tcgen05 GEMM");
      ++n_launches; ++n_tc;
      return tc_flat_fuses_finalize(f);
    }
  }
  if (o.no_store || o.gn_mr || o.n_store) throw std::runtime_error("athtd: fused DConv epilogue needs the tcgen05 kernel");
  GemmDesc d = gemm_desc_zero();
  d.G1 = B; d.G2 = as.G2; d.Mg = Mg; d.N = o.N; d.K = K; d.Ktap = K;
  d.sAg1 = as.g1_stride(); d.sAg2 = as.g2_stride();
  d.B = o.w; d.sBn = K; d.sBk = 1;
  if (o.mode == CONV_ROWS) { d.A = o.a + as.origin(); d.sAm = C; }
  else if (o.mode == CONV_K8S4) { d.A = o.a + as.origin() - 2L * C; d.sAm = 4L * C; }
  else if (o.mode == CONV_T) { d.A = o.a + as.origin() - C; d.sAm = C; }
  else {
    d.A = o.a + as.origin(); d.sAm = C; d.ntaps = 3; d.Ktap = C;
    for (int k = 0; k < 3; ++k) { d.shG2[k] = freq ? (k - 1) * o.dil : 0; d.shM[k] = freq ? 0 : (k - 1) * o.dil; }
  }
  // output rows: interior of cs, except the transposed conv whose row q sits one row before the interior (x[q-1])
  const long c_origin = o.mode == CONV_T ? cs.origin() - cs.C : cs.origin();
  d.C = o.c + c_origin; d.sCg1 = cs.g1_stride(); d.sCg2 = cs.g2_stride(); d.sCm = cs.C;
  d.bias = o.bias; d.act = o.act; d.glu = o.glu; d.colscale = o.colscale;
  if (o.res) { d.res = o.res + c_origin; d.sRg1 = d.sCg1; d.sRg2 = d.sCg2; d.sRm = d.sCm; }
  d.rowtab = o.rowtab; d.rowtab_scale = o.rowtab_scale;
  d.stats = o.stats; d.stat_mode = o.stat_mode;
  d.convt_cout = o.mode == CONV_T ? o.N / 4 : 0;
  gemm(d, st);
  return false;
}

template <typename T>
static ConvOp<T> conv_op(int mode, const T* a, RowSpace as, const T* w, int N, T* c, RowSpace cs) {
  ConvOp<T> o;
  memset(&o, 0, sizeof(o));
  o.mode = mode; o.a = a; o.as = as; o.w = w; o.N = N; o.c = c; o.cs = cs; o.dil = 1;
  return o;
}

// bf16 only: one DConv residual layer as three tiled mma.sync passes (dconv_tile.cu)
static int dconv_tile_run(bf16* y, RowSpace ys, bf16* h, int dil, const bf16* w1p, const float* b1p, const float* g1wp, const float* g1bp,
                          const bf16* w2p, const float* b2i, const float* g2wi, const float* g2bi, const bf16* ghi, const bf16* glo, const float* gv,
                      const float* scale, double* st1,
                          double* st2, const bf16* rw, const float* rb, bf16* out, cudaStream_t st) {
  return launch_dconv_tile(y, ys, h, dil, w1p, b1p, g1wp, g1bp, w2p, b2i, g2wi, g2bi, ghi, glo, gv, scale, st1, st2, rw, rb, out, st);
}
static int dconv_tile_run(float*, RowSpace, float*, int, const float*, const float*, const float*, const float*, const float*,
                          const float*, const float*, const float*, const float*, const float*, const float*, const float*, double*, double*,
                          const float*, const float*, float*, cudaStream_t) {
  return 1;
}

// bf16 only: last decoder layer (transposed conv 48 -> 4 + resize + skip) in one kernel (small_conv.cu)
static bool dec_last_run(bool freq, const bf16* x, RowSpace xs, const float* w32, const bf16* wp, const float* b, const float* b4,
                         const bf16* skip, RowSpace ss, bf16* out, RowSpace os, cudaStream_t st) {
  if (xs.C != 48 || os.C != 4) return false;
  if (freq) { if (os.R != xs.R) return false; launch_dec_last_freq(x, xs, w32, b, skip, ss, out, os, st); }
  else { if (os.R != 4 * xs.R) return false; launch_dec_last_time(x, xs, wp, b4, skip, ss, out, os, st); }
  return true;
}
static bool dec_last_run(bool, const float*, RowSpace, const float*, const float*, const float*, const float*, const float*, RowSpace,
                         float*, RowSpace, cudaStream_t) { return false; }

// bf16 only: time-branch level-0 conv fused with the waveform normalisation (small_conv.cu)
static bool tenc0_run(const float* wav, const float* ms, int L, const bf16* w, const float* bias, bf16* y, RowSpace ys, cudaStream_t st) {
  launch_tenc0_conv(wav, ms, L, w, bias, y, ys, st);
  return true;
}
static bool tenc0_run(const float*, const float*, int, const float*, const float*, float*, RowSpace, cudaStream_t) { return false; }

// bf16 only: fused per-slab layer tail (enc_row.cu).  run=false just answers whether it applies.
template <typename T>
bool PlanT<T>::enc_row_dispatch(bool run, int i, const T* x, RowSpace xin, const T* y, T* out, RowSpace ys, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    const int C = kCh[i];
    if (!(C == 48 || C == 96) || !enc_row_supported(C, sh.Tf, false)) return false;
    if (!run) return true;
    const bool fuse = i == 0 && enc_row_supported(C, sh.Tf, true);
    const std::string p = std::string("htdemucs.encoder.") + std::to_string(i);
    EncRowParams P;
    P.cw = (const bf16*)PW(p + ".conv.w"); P.cb = P32(p + ".conv.bias");
    for (int dd = 0; dd < 2; ++dd) {
      const std::string q = p + ".dconv.layers." + std::to_string(dd);
      P.w1[dd] = (const bf16*)PW(q + ".0.wp"); P.b1[dd] = PA(q + ".0.bp"); P.g1w[dd] = PA(q + ".1.wp"); P.g1b[dd] = PA(q + ".1.bp");
      P.w2[dd] = (const bf16*)PW(q + ".3.wp"); P.b2[dd] = PA(q + ".3.bi"); P.g2w[dd] = PA(q + ".4.wi"); P.g2b[dd] = PA(q + ".4.bi");
      P.scale[dd] = P32(q + ".6.scale");
    }
    P.rw = (const bf16*)PW(p + ".rewrite.w"); P.rb = PA(p + ".rewrite.b");
    P.emb = i == 0 ? P32("htdemucs.freq_emb.embedding.weight") : nullptr; P.emb_scale = 10.0f * 0.2f;
    P.zspec = fuse ? Z : nullptr; P.ms_spec = ms_spec; P.zrows = 2048;
    launch_enc_row((const bf16*)x, xin, (const bf16*)y, (bf16*)out, ys, P, fuse, st);
    ++n_launches;
    return true;
  } else {
    return false;      // fp32 build: SIMT kernels only
  }
}

// ---- one HEncLayer (demucs hdemucs.py:HEncLayer) on a channels-last row space.
//   x   : input activation (padded row space xin, C_in channels)
//   y   : conv+GELU output, updated in place by the two DConv residual layers (space ys, C channels)
//   out : rewrite+GLU output (same geometry as ys)
template <typename T>
void PlanT<T>::enc_layer(bool freq, int i, const T* x, RowSpace xin, T* y, RowSpace ys, T* out, cudaStream_t st) {
  const Shapes& s = sh;
  const int C = kCh[i], H = C / 8;
  const std::string p = std::string("htdemucs.") + (freq ? "encoder." : "tencoder.") + std::to_string(i);
  const int G2 = ys.G2;
  const int R = ys.R;
  // bf16 build, frequency levels 1-2: the whole layer tail (level 1: the k8s4 conv too) runs per (segment, frequency
  // row) slab in shared memory -- input read once, output written once (enc_row.cu)
  const bool row_fused = freq && use_fused_dconv && use_tc && enc_row_dispatch(false, i, nullptr, xin, nullptr, nullptr, ys, st);
  const bool row_conv = row_fused && i == 0 && enc_row_supported(C, s.Tf, true);
  const bool wav_conv = !freq && i == 0 && use_tc && use_fused_dconv && sizeof(T) == 2 && cur_wav != nullptr &&
                        tenc0_run(cur_wav, ms_wav, s.L, PW(p + ".conv.w"), P32(p + ".conv.bias"), y, ys, st);
  if (wav_conv) ++n_launches;
  if (!row_conv && !wav_conv) {  // strided conv k8 s4 p2 (+ right zero pad to a multiple of 4 on the time branch) + GELU
    ConvOp<T> o = conv_op<T>(CONV_K8S4, x, xin, PW(p + ".conv.w"), C, y, ys);
    o.bias = P32(p + ".conv.bias"); o.act = ACT_GELU;
    conv(o, st);
  }
  if (row_fused) { enc_row_dispatch(true, i, x, xin, y, out, ys, st); return; }
  const bool tc_dconv = sizeof(T) == 2 && use_tc && tensor_map_api_available();
  const int Hp = hidden_pad(H);
  RowSpace hs = ys; hs.C = tc_dconv ? Hp : H;
  RowSpace es = ys; es.C = 2 * C;
  if (freq && use_fused_dconv && dconv_row_supported<T>(C, s.Tf)) {
    // frequency branch: GroupNorm statistics are per (segment, frequency row) -> one CTA per row, both layers fused
    const float* ptrs[18];
    for (int dd = 0; dd < 2; ++dd) {
      const std::string q = p + ".dconv.layers." + std::to_string(dd);
      const char* names[9] = {".0.weight", ".0.bias", ".1.weight", ".1.bias", ".3.weight", ".3.bias", ".4.weight", ".4.bias", ".6.scale"};
      for (int k = 0; k < 9; ++k) ptrs[9 * dd + k] = P32(q + names[k]);
    }
    launch_dconv_row<T>(y, ys, ptrs, st); ++n_launches;
  } else
  for (int dd = 0; dd < 2; ++dd) {
    const std::string q = p + ".dconv.layers." + std::to_string(dd);
    double* st_h = freq ? st_df[i][dd][0] : st_dt[i][dd][0];
    double* st_e = freq ? st_df[i][dd][1] : st_dt[i][dd][1];
    const long nstat = freq ? (long)s.B * R : s.B;
    const int smode = freq ? STAT_PER_G1_M : STAT_PER_G1;
    if (tc_dconv && use_fused_dconv && dconv_tile_supported(C) && (!freq || R <= 32)) {
      // three bandwidth-bound mma.sync passes (conv3 + stats | GN+GELU + expand stats | expand + GN + GLU + residual)
      // second residual layer of a narrow time-branch level: the 1x1 rewrite + GLU runs on the updated tile inside pass C
      const bool fuse_rw = dd == 1 && dconv_tile_can_rewrite(C, freq);
      const int rc = dconv_tile_run(y, ys, hbuf, 1 << dd, PW(q + ".0.wp"), PA(q + ".0.bp"), PA(q + ".1.wp"), PA(q + ".1.bp"),
                                    PW(q + ".3.wp"), PA(q + ".3.bi"), PA(q + ".4.wi"), PA(q + ".4.bi"), PW(q + ".3.ghi"), PW(q + ".3.glo"),
                                    PA(q + ".3.gv"), P32(q + ".6.scale"), st_h, st_e,
                                    fuse_rw ? PW(p + ".rewrite.w") : nullptr, fuse_rw ? PA(p + ".rewrite.b") : nullptr,
                                    fuse_rw ? out : nullptr, st);
      if (rc != 0) throw std::runtime_error("athtd: dconv_tile launch failed");
      n_launches += 3;
      if (fuse_rw) return;
      continue;
    }
    if (tc_dconv) {
      // DConv layer on tensor cores in three passes over a narrow hidden buffer (C/8 channels zero-padded to Hp):
      //   1. h = conv3(y) + GroupNorm partial sums     2. h <- GELU(GN(h))
      //   3. e = W2 h: statistics only (not stored)    4. recompute e, GN, GLU, LayerScale, + y -> y
      {
        ConvOp<T> o = conv_op<T>(CONV_DIL3, y, ys, PW(q + ".0.wp"), Hp, hbuf, hs);
        o.dil = 1 << dd; o.bias = PA(q + ".0.bp"); o.stats = st_h; o.stat_mode = smode;
        conv(o, st);
      }
      launch_finalize_gn(st_h, (double)H * (freq ? s.Tf : R), mr, nstat, freq ? 1 : STAT_SLOTS, st); ++n_launches;
      launch_gn_gelu<T>(hbuf, hs, G2, freq ? 1 : 0, mr, PA(q + ".1.wp"), PA(q + ".1.bp"), st); ++n_launches;
      {
        ConvOp<T> o = conv_op<T>(CONV_ROWS, hbuf, hs, PW(q + ".3.wp"), 2 * C, y, ys);
        o.bias = PA(q + ".3.bi"); o.stats = st_e; o.stat_mode = smode; o.no_store = 1;
        conv(o, st);
      }
      launch_finalize_gn(st_e, (double)2 * C * (freq ? s.Tf : R), mr, nstat, freq ? 1 : STAT_SLOTS, st); ++n_launches;
      {
        ConvOp<T> o = conv_op<T>(CONV_ROWS, hbuf, hs, PW(q + ".3.wp"), 2 * C, y, ys);
        o.bias = PA(q + ".3.bi"); o.gn_mr = mr; o.gn_w = PA(q + ".4.wi"); o.gn_b = PA(q + ".4.bi"); o.gn_mode = smode;
        o.glu = 1; o.colscale = P32(q + ".6.scale"); o.res = y;
        conv(o, st);
      }
      continue;
    }
    {  // dilated k3 conv C -> C/8 along time (freq branch: taps shift the frame index)
      ConvOp<T> o = conv_op<T>(CONV_DIL3, y, ys, PW(q + ".0.w"), H, hbuf, hs);
      o.dil = 1 << dd; o.bias = P32(q + ".0.bias"); o.stats = st_h; o.stat_mode = smode;
      conv(o, st);
    }
    launch_finalize_gn(st_h, (double)H * (freq ? s.Tf : R), mr, nstat, freq ? 1 : STAT_SLOTS, st); ++n_launches;
    launch_gn_gelu<T>(hbuf, hs, G2, freq ? 1 : 0, mr, P32(q + ".1.weight"), P32(q + ".1.bias"), st); ++n_launches;
    {  // 1x1 expand C/8 -> 2C
      ConvOp<T> o = conv_op<T>(CONV_ROWS, hbuf, hs, PW(q + ".3.w"), 2 * C, ebuf, es);
      o.bias = P32(q + ".3.bias"); o.stats = st_e; o.stat_mode = smode;
      conv(o, st);
    }
    launch_finalize_gn(st_e, (double)2 * C * (freq ? s.Tf : R), mr, nstat, freq ? 1 : STAT_SLOTS, st); ++n_launches;
    launch_gn_glu_res<T>(y, ys, ebuf, es, G2, freq ? 1 : 0, mr, P32(q + ".4.weight"), P32(q + ".4.bias"),
                         P32(q + ".6.scale"), st); ++n_launches;
  }
  {  // 1x1 rewrite C -> 2C, GLU (+ frequency embedding after encoder 0, ATHTDemucs_v2.py:212-215)
    ConvOp<T> o = conv_op<T>(CONV_ROWS, y, ys, PW(p + ".rewrite.w"), 2 * C, out, ys);
    o.bias = PA(p + ".rewrite.b"); o.glu = 1;
    if (freq && i == 0) { o.rowtab = P32("htdemucs.freq_emb.embedding.weight"); o.rowtab_scale = 10.0f * 0.2f; }
    conv(o, st);
  }
}

// ---- attention for one branch: scores = (Q K^T)/sqrt(64), softmax, O = P V   (8 heads of 64)
static inline int flash_dispatch(const bf16* q, long ldq, const bf16* k, const bf16* v, long ldkv, int B, int Sq, int Sk, bf16* o,
                                 cudaStream_t st) {
  if (!flash_attn_supported(ldq, ldkv, 512)) return -1;
  return launch_flash_attn(q, ldq, k, v, ldkv, B, Sq, Sk, o, 512, st);
}
static inline int flash_dispatch(const float*, long, const float*, const float*, long, int, int, int, float*, cudaStream_t) { return -1; }

template <typename T>
void PlanT<T>::attention(const T* q, long ldq, const T* k, const T* v, long ldkv, int Sq, int Sk, T* o, cudaStream_t st) {
  const int B = sh.B;
  if (use_tc && use_flash) {
    prof_begin(4.0 * (double)B * 8 * Sq * Sk * 64 * 1e-9, st, "attn", (long)B * 8 * Sq, Sk, 64);
    int rc = flash_dispatch(q, ldq, k, v, ldkv, B, Sq, Sk, o, st);
    prof_end(st);
    if (rc == 0) { ++n_launches; ++n_tc; return; }
    if (rc > 0) throw std::runtime_error("athtd: cuTensorMapEncodeTiled failed for the attention kernel");
    if (profiling) { prof_used -= 2; }
  }
  {
    GemmDesc d = gemm_desc_zero();
    d.G1 = B; d.G2 = 8; d.Mg = Sq; d.N = Sk; d.K = 64; d.Ktap = 64; d.grouped = 1;
    d.A = q; d.sAg1 = (long)Sq * ldq; d.sAg2 = 64; d.sAm = ldq;
    d.B = k; d.sBg1 = (long)Sk * ldkv; d.sBg2 = 64; d.sBn = ldkv; d.sBk = 1;
    d.C = scores; d.sCg1 = 8L * Sq * Sk; d.sCg2 = (long)Sq * Sk; d.sCm = Sk;
    d.alpha = 0.125f;
    gemm(d, st);
  }
  launch_softmax_rows<T>(scores, (long)B * 8 * Sq, Sk, st); ++n_launches;
  {
    GemmDesc d = gemm_desc_zero();
    d.G1 = B; d.G2 = 8; d.Mg = Sq; d.N = 64; d.K = Sk; d.Ktap = Sk; d.grouped = 1;
    d.A = scores; d.sAg1 = 8L * Sq * Sk; d.sAg2 = (long)Sq * Sk; d.sAm = Sk;
    d.B = v; d.sBg1 = (long)Sk * ldkv; d.sBg2 = 64; d.sBn = 1; d.sBk = ldkv;
    d.C = o; d.sCg1 = (long)Sq * 512; d.sCg2 = 64; d.sCm = 512;
    gemm(d, st);
  }
}

template <typename T>
void PlanT<T>::linear(const T* a, int S, int K, const T* w, int N, const float* bias, int act, T* c, cudaStream_t st) {
  ConvOp<T> o = conv_op<T>(CONV_ROWS, a, make_space(sh.B, 1, S, K, false), w, N, c, make_space(sh.B, 1, S, N, false));
  o.bias = bias; o.act = act;
  conv(o, st);
}

// x <- x + gamma * (a W^T + b), optional per-sample (sum, sumsq) of the result
template <typename T>
bool PlanT<T>::linear_res(const T* a, int S, int K, const T* w, int N, const float* bias, const float* gamma, T* x,
                          double* stats, cudaStream_t st, unsigned* fin_counter) {
  ConvOp<T> o = conv_op<T>(CONV_ROWS, a, make_space(sh.B, 1, S, K, false), w, N, x, make_space(sh.B, 1, S, N, false));
  o.bias = bias; o.colscale = gamma; o.res = x;
  if (stats) { o.stats = stats; o.stat_mode = STAT_PER_G1; }
  if (stats && fin_counter) { o.fin_mr = mr; o.fin_count = (double)S * N; o.fin_counter = fin_counter; }
  return conv(o, st);
}

// FFN half of a transformer layer + norm_out (MyGroupNorm over all tokens of the segment).  The norm_out apply pass also
// emits the LayerNorm(s) the NEXT layer starts with (same row statistics, the next layer's affine): n1 = next layer's norm1
// of this branch, n2 = the other branch's next-layer norm2 applied to this branch (cross-attention keys / values).
template <typename T>
void PlanT<T>::xf_ffn_and_norm(const std::string& p, const char* ffn_norm, T* x, int S, double* stats, unsigned* fin_counter, T* n1,
                               const float* n1w, const float* n1b, T* n2, const float* n2w, const float* n2b, cudaStream_t st) {
  const long rows = (long)sh.B * S;
  const RowSpace none{};
  launch_norm_rows<T>(x, nullptr, hn[4], rows, 512, S, nullptr, nullptr, nullptr, P32(p + "." + ffn_norm + ".weight"),
                      P32(p + "." + ffn_norm + ".bias"), nullptr, none, st); ++n_launches;
  linear(hn[4], S, 512, PW(p + ".linear1.w"), 2048, P32(p + ".linear1.bias"), ACT_GELU, ffn, st);
  // (the tcgen05 GEMM finalises the per-segment sums itself: its last CTA turns the slots into (mean, rstd))
  if (!linear_res(ffn, S, 2048, PW(p + ".linear2.w"), 512, P32(p + ".linear2.bias"), P32(p + ".gamma_2.scale"), x, stats, st, fin_counter)) {
    launch_finalize_gn(stats, (double)S * 512, mr, sh.B, STAT_SLOTS, st); ++n_launches;
  }
  launch_norm_rows<T>(x, x, n1, rows, 512, S, mr, P32(p + ".norm_out.weight"), P32(p + ".norm_out.bias"), n1w, n1b,
                      nullptr, none, st, n2, n2w, n2b); ++n_launches;
}

template <typename T>
void PlanT<T>::cross_transformer(cudaStream_t st) {
  const Shapes& s = sh;
  const std::string xp = "htdemucs.crosstransformer";
  const int B = s.B;
  const RowSpace none{};
  T* X[2] = {tokf, tokt};
  const int S[2] = {s.Sf, s.St};
  const char* stacks[2] = {".layers.", ".layers_t."};
  // hn[br] = this branch's next-layer norm1 output, hn[2 + br] = LayerNorm of this branch with the OTHER stack's norm2 (read
  // by the other branch's cross-attention), hn[4] = scratch.  Both are written by the previous layer's norm_out pass.
  auto next_norms = [&](int l, int br, T*& n1, const float*& n1w, const float*& n1b, T*& n2, const float*& n2w, const float*& n2b) {
    n1 = nullptr; n1w = n1b = nullptr; n2 = nullptr; n2w = n2b = nullptr;
    if (l + 1 >= 5) return;
    const std::string pn = xp + stacks[br] + std::to_string(l + 1);
    n1 = hn[br]; n1w = P32(pn + ".norm1.weight"); n1b = P32(pn + ".norm1.bias");
    if ((l + 1) % 2 == 1) {      // next layer is a cross layer: the other stack normalises this branch with its norm2
      const std::string po = xp + stacks[1 - br] + std::to_string(l + 1);
      n2 = hn[2 + br]; n2w = P32(po + ".norm2.weight"); n2b = P32(po + ".norm2.bias");
    }
  };
  for (int l = 0; l < 5; ++l) {
    if (l % 2 == 0) {
      for (int br = 0; br < 2; ++br) {
        const std::string p = xp + stacks[br] + std::to_string(l);
        const long rows = (long)B * S[br];
        const T* h1 = hn[br];
        if (l == 0) {
          launch_norm_rows<T>(X[br], nullptr, hn[4], rows, 512, S[br], nullptr, nullptr, nullptr, P32(p + ".norm1.weight"),
                              P32(p + ".norm1.bias"), nullptr, none, st); ++n_launches;
          h1 = hn[4];
        }
        linear(h1, S[br], 512, PW(p + ".in_proj.w"), 1536, P32(p + ".self_attn.in_proj_bias"), ACT_NONE, qkv, st);
        attention(qkv, 1536, qkv + 512, qkv + 1024, 1536, S[br], S[br], obuf, st);
        linear_res(obuf, S[br], 512, PW(p + ".out_proj.w"), 512, P32(p + ".self_attn.out_proj.bias"),
                   P32(p + ".gamma_1.scale"), X[br], nullptr, st);
        T *n1, *n2; const float *n1w, *n1b, *n2w, *n2b;
        next_norms(l, br, n1, n1w, n1b, n2, n2w, n2b);
        xf_ffn_and_norm(p, "norm2", X[br], S[br], st_xf[l][br], fin_xf + 2 * l + br, n1, n1w, n1b, n2, n2w, n2b, st);
      }
    } else {
      // both branches read the PRE-update other branch (old_x, demucs transformer.py): hn[0..3] were all written by the
      // previous (self-attention) layer, before any update of this layer
      for (int br = 0; br < 2; ++br) {
        const std::string p = xp + stacks[br] + std::to_string(l);
        const int o = 1 - br;
        const T* w_in = PW(p + ".in_proj.w");
        const float* b_in = P32(p + ".cross_attn.in_proj_bias");
        linear(hn[br], S[br], 512, w_in, 512, b_in, ACT_NONE, qkv, st);
        linear(hn[2 + o], S[o], 512, w_in + 512L * 512, 1024, b_in + 512, ACT_NONE, kvb, st);
        attention(qkv, 512, kvb, kvb + 512, 1024, S[br], S[o], obuf, st);
        linear_res(obuf, S[br], 512, PW(p + ".out_proj.w"), 512, P32(p + ".cross_attn.out_proj.bias"),
                   P32(p + ".gamma_1.scale"), X[br], nullptr, st);
        T *n1, *n2; const float *n1w, *n1b, *n2w, *n2b;
        next_norms(l, br, n1, n1w, n1b, n2, n2w, n2b);
        xf_ffn_and_norm(p, "norm3", X[br], S[br], st_xf[l][br], fin_xf + 2 * l + br, n1, n1w, n1b, n2, n2w, n2b, st);
      }
    }
  }
}

template <typename T>
void PlanT<T>::encode(const float* wav, cudaStream_t st, const float* xnorm) {
  const Shapes& s = sh;
  const int B = s.B, Tf = s.Tf;
  const RowSpace none{};
  cudaMemsetAsync((char*)ws_base() + stats_begin, 0, dec_stats_begin - stats_begin, st);
  // spectral front end + input normalisation (ATHTDemucs_v2.py:261-275)
  if (xnorm) {
    // AudioTextHTDemucs._encode(x, xt) on its own (ATHTDemucs_v2.py:190-236): the caller supplies the ALREADY normalised
    // complex-as-channels spectrogram [B, Tf, 2048, 4] and waveform [B, 2, L]; they take the place of Z / wav with the
    // identity normalisation (mean 0, 1e-5 + std == 1 in fp32).  A decode after this would mask the wrong spectrogram.
    cudaMemcpyAsync(Z, xnorm, sizeof(float) * (size_t)B * Tf * 2048 * 4, cudaMemcpyDeviceToDevice, st);
    launch_set_meanstd(ms_spec, B, 0.0f, 0.99999f, st); ++n_launches;
    launch_set_meanstd(ms_wav, B, 0.0f, 0.99999f, st); ++n_launches;
  } else {
    launch_stft_cac(wav, B, s.L, Tf, Z, st_spec, consts.tw, consts.win, st); ++n_launches;
    launch_sum_sumsq(wav, B, 2L * s.L, st_wav, st); ++n_launches;
    launch_finalize_meanstd(st_spec, 4.0 * 2048.0 * Tf, ms_spec, B, st); ++n_launches;
    launch_finalize_meanstd(st_wav, 2.0 * s.L, ms_wav, B, st); ++n_launches;
  }
  // the packed bf16 copy of the normalised spectrogram is only needed when level 0 does not read Z itself (enc_row.cu)
  const bool z_direct = sizeof(T) == 2 && use_tc && use_fused_dconv && enc_row_dispatch(false, 0, nullptr, xf0_rs, nullptr, nullptr, yf_rs[0], st) &&
                        enc_row_supported(kCh[0], Tf, true);
  if (!z_direct) { launch_pack_spec<T>(Z, ms_spec, xf0, xf0_rs, Tf, st); ++n_launches; }
  cur_wav = wav;
  if (!(use_tc && use_fused_dconv && sizeof(T) == 2)) { launch_pack_wav<T>(wav, ms_wav, xt0, xt0_rs, s.L, st); ++n_launches; }
  // interleaved encoders (ATHTDemucs_v2.py:196-217)
  const T* xf = xf0; RowSpace xfs = xf0_rs;
  const T* xt = xt0; RowSpace xts = xt0_rs;
  for (int i = 0; i < 4; ++i) {
    enc_layer(false, i, xt, xts, yt[i], yt_rs[i], et[i], st);
    xt = et[i]; xts = yt_rs[i];
    enc_layer(true, i, xf, xfs, yf[i], yf_rs[i], ef[i], st);
    xf = ef[i]; xfs = yf_rs[i];
  }
  // bottleneck: 1x1 up-sample 384->512, norm_in + positional embeddings, 5 layers, 1x1 down-sample
  const std::string hp = "htdemucs.";
  {
    ConvOp<T> o = conv_op<T>(CONV_ROWS, ef[3], yf_rs[3], PW(hp + "channel_upsampler.w"), 512, hn[1], make_space(B, Tf, 8, 512, false));
    o.bias = P32(hp + "channel_upsampler.bias");
    conv(o, st);
    o = conv_op<T>(CONV_ROWS, et[3], yt_rs[3], PW(hp + "channel_upsampler_t.w"), 512, hn[2], make_space(B, 1, s.St, 512, false));
    o.bias = P32(hp + "channel_upsampler_t.bias");
    conv(o, st);
  }
  const std::string xp = "htdemucs.crosstransformer";
  launch_norm_rows<T>(hn[1], nullptr, tokf, (long)B * s.Sf, 512, s.Sf, nullptr, nullptr, nullptr, P32(xp + ".norm_in.weight"),
                      P32(xp + ".norm_in.bias"), consts.pe2d, none, st); ++n_launches;
  launch_norm_rows<T>(hn[2], nullptr, tokt, (long)B * s.St, 512, s.St, nullptr, nullptr, nullptr, P32(xp + ".norm_in_t.weight"),
                      P32(xp + ".norm_in_t.bias"), consts.pe1d, none, st); ++n_launches;
  cross_transformer(st);
  linear(tokf, s.Sf, 512, PW(hp + "channel_downsampler.w"), 384, P32(hp + "channel_downsampler.bias"), ACT_NONE, xenc, st);
  linear(tokt, s.St, 512, PW(hp + "channel_downsampler_t.w"), 384, P32(hp + "channel_downsampler_t.bias"), ACT_NONE, xtenc, st);
}

// text conditioning for prompt p (TextCrossAttention, ATHTDemucs_v2.py:38-58): with a single key the
// softmax is exactly 1, so attn_out is one 384-vector per (segment, prompt) -- SURVEY.md quirk Q3.
template <typename T>
void PlanT<T>::text_vectors(const float* emb, cudaStream_t st) {
  launch_text_vectors(emb, P32("text_attn.v_proj.weight"), P32("text_attn.v_proj.bias"),
                      P32("text_attn.attn.in_proj_weight") + 768L * 384, P32("text_attn.attn.in_proj_bias") + 768,
                      P32("text_attn.attn.out_proj.weight"), P32("text_attn.attn.out_proj.bias"), cvec, sh.B * sh.P, st);
  ++n_launches;
}

template <typename T>
void PlanT<T>::text_condition(int p, const T* x, int S, T* out, RowSpace outs, cudaStream_t st) {
  const long rows = (long)sh.B * S;
  launch_add_rowvec<T>(x, t1, S, 384, sh.B, cvec + (size_t)p * 384, (long)sh.P * 384, st); ++n_launches;
  linear(t1, S, 384, PW("text_attn.out_mlp.0.w"), 384, P32("text_attn.out_mlp.0.bias"), ACT_GELU, t2, st);
  {
    ConvOp<T> o = conv_op<T>(CONV_ROWS, t2, make_space(sh.B, 1, S, 384, false), PW("text_attn.out_mlp.2.w"), 384, t1,
                             make_space(sh.B, 1, S, 384, false));
    o.bias = P32("text_attn.out_mlp.2.bias"); o.res = t1;
    conv(o, st);
  }
  launch_norm_rows<T>(t1, nullptr, out, rows, 384, S, nullptr, nullptr, nullptr, P32("text_attn.norm_out.weight"),
                      P32("text_attn.norm_out.bias"), nullptr, outs, st); ++n_launches;
}

// one FreqDecoder / TimeDecoder layer (ATHTDemucs_v2.py:82-104, 125-139): transposed conv as ONE GEMM
// over [x[q-1], x[q]] row pairs producing the 4 output phases, GroupNorm statistics in the epilogue,
// then GN + GELU + linear resize to the target length + 0.1 * resized truncated skip.
template <typename T>
void PlanT<T>::dec_layer(bool freq, int i, int p, const T* x, RowSpace xs, T* out, RowSpace os, const T* skip, RowSpace ss,
                         cudaStream_t st) {
  const Shapes& s = sh;
  const int Cout = kDecCh[i + 1];
  const int G2 = xs.G2;
  const int Rin = xs.R;
  const std::string q = std::string(freq ? "freq_decoder" : "time_decoder") + ".layers." + std::to_string(i);
  if (i == 3 && use_tc && use_fused_dconv &&
      dec_last_run(freq, x, xs, P32(q + ".0.weight"), PW(q + ".0.w"), P32(q + ".0.bias"), PA(q + ".0.b4"), skip, ss, out, os, st)) {
    ++n_launches;
    return;
  }
  double* stt = st_dec + 2L * s.B * STAT_SLOTS * ((size_t)p * 6 + (freq ? 0 : 3) + (i < 3 ? i : 0));
  RowSpace us = xs; us.C = 4 * Cout;
  ConvOp<T> o = conv_op<T>(CONV_T, x, xs, PW(q + ".0.w"), 4 * Cout, ubuf, us);
  o.bias = PA(q + ".0.b4");
  if (i < 3) { o.stats = stt; o.stat_mode = STAT_PER_G1; }
  // exact 4:1 resize (4*Rin -> Rin rows) reads only rows 4d+1, 4d+2 = phases 3 and 0: phases 1, 2 feed the GroupNorm
  // statistics but are never stored; tensor-core path only
  if (sizeof(T) == 2 && use_tc && os.R == Rin) { o.skip_lo = Cout; o.skip_hi = 3 * Cout; }
  if (i < 3) { o.fin_mr = mr; o.fin_count = (double)Cout * 4.0 * Rin * G2; o.fin_counter = fin_dec + (size_t)p * 6 + (freq ? 0 : 3) + i; }
  const bool finalized = conv(o, st);
  if (i < 3 && !finalized) { launch_finalize_gn(stt, (double)Cout * 4.0 * Rin * G2, mr, s.B, STAT_SLOTS, st); ++n_launches; }
  launch_dec_apply<T>(ubuf, 4 * Rin, us, Cout, out, os, G2, i < 3 ? 1 : 0, mr, i < 3 ? P32(q + ".1.weight") : nullptr,
                      i < 3 ? P32(q + ".1.bias") : nullptr, skip, ss, st); ++n_launches;
}

template <typename T>
void PlanT<T>::decode(const float* emb, float* out, cudaStream_t st) {
  const Shapes& s = sh;
  const int B = s.B, Tf = s.Tf;
  cudaMemsetAsync((char*)ws_base() + dec_stats_begin, 0, stats_begin + stats_bytes - dec_stats_begin, st);
  text_vectors(emb, st);
  for (int p = 0; p < s.P; ++p) {
    text_condition(p, xenc, s.Sf, xc, xc_rs, st);
    text_condition(p, xtenc, s.St, xtc, xtc_rs, st);
    const T* x = xc; RowSpace xs = xc_rs;
    for (int i = 0; i < 4; ++i) { dec_layer(true, i, p, x, xs, df[i], df_rs[i], ef[3 - i], yf_rs[3 - i], st); x = df[i]; xs = df_rs[i]; }
    x = xtc; xs = xtc_rs;
    for (int i = 0; i < 4; ++i) { dec_layer(false, i, p, x, xs, dt[i], dt_rs[i], et[3 - i], yt_rs[3 - i], st); x = dt[i]; xs = dt_rs[i]; }
    // mask -> inverse STFT -> overlap-add -> + de-normalised time branch, one kernel (ATHTDemucs_v2.py:294-324)
    launch_istft_fused<T>(Z, Tf, s.L, B, 1, df[3], df_rs[3], 1, P32("freq_out.weight"), P32("freq_out.bias"), dt[3], dt_rs[3],
                          P32("time_out.weight"), P32("time_out.bias"), ms_wav, 1, out + (size_t)p * 2 * s.L, (long)s.P * 2 * s.L,
                          consts.tw, consts.win, st); ++n_launches;
  }
}

template <typename T>
void PlanT<T>::drop_graphs() {
  for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  graphs.clear();
}

template <typename T>
PlanT<T>::~PlanT() {
  drop_graphs();
  if (cap_stream) cudaStreamDestroy(cap_stream);
  for (auto e : prof_ev) cudaEventDestroy(e);
}

template <typename T>
int PlanT<T>::forward(const float* wav, const float* emb, float* out, cudaStream_t st) {
  if (use_graph && !profiling && !graph_broken) {
    GraphEntry* hit = nullptr;
    for (auto& g : graphs) if (g.B == sh.B && g.wav == wav && g.emb == emb && g.out == out) { hit = &g; break; }
    if (hit && hit->exec) {
      hit->stamp = ++graph_clock;
      n_launches = hit->n_launches; n_tc = hit->n_tc; ++n_graph_replays;
      return (int)cudaGraphLaunch(hit->exec, st);
    }
    if (hit) {      // second sighting of this argument set: capture the launch sequence on a private stream, then replay it
      if (!cap_stream && cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking) != cudaSuccess) { graph_broken = true; cudaGetLastError(); }
      if (!graph_broken && cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        cudaGraph_t graph = nullptr;
        try {
          n_launches = 0; n_tc = 0;
          encode(wav, cap_stream);
          n_enc_launches = n_launches; n_enc_tc = n_tc;
          decode(emb, out, cap_stream);
        } catch (...) {
          cudaStreamEndCapture(cap_stream, &graph);
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          graph_broken = true;
          throw;
        }
        cudaGraphExec_t exec = nullptr;
        if (cudaStreamEndCapture(cap_stream, &graph) == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          cudaGraphDestroy(graph);
          hit->exec = exec; hit->n_launches = n_launches; hit->n_tc = n_tc; hit->stamp = ++graph_clock;
          ++n_graph_replays;
          return (int)cudaGraphLaunch(exec, st);
        }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        graph_broken = true;          // capture is not available here: stay on the eager path for good
      } else {
        cudaGetLastError();
        graph_broken = true;
      }
    } else {
      if (graphs.size() >= 8) {       // forget the least recently used argument set
        size_t lru = 0;
        for (size_t i = 1; i < graphs.size(); ++i) if (graphs[i].stamp < graphs[lru].stamp) lru = i;
        if (graphs[lru].exec) cudaGraphExecDestroy(graphs[lru].exec);
        graphs.erase(graphs.begin() + lru);
      }
      graphs.push_back(GraphEntry{sh.B, wav, emb, out, nullptr, 0, 0, ++graph_clock});
    }
  }
  n_launches = 0; n_tc = 0;
  encode(wav, st);
  n_enc_launches = n_launches; n_enc_tc = n_tc;
  decode(emb, out, st);
  return (int)cudaGetLastError();
}

template <typename T>
bool PlanT<T>::tap(const std::string& name, TapInfo& ti) const {
  const Shapes& s = sh;
  auto set = [&](const void* p, int dt, long n, int d0, int d1, int d2, int d3) {
    ti.ptr = p; ti.dtype = dt; ti.numel = n; ti.dims[0] = d0; ti.dims[1] = d1; ti.dims[2] = d2; ti.dims[3] = d3; return true;
  };
  const int TD = sizeof(T) == 4 ? 0 : 1;
  auto rs = [&](const T* p, const RowSpace& r) { ti.geom[0] = r.G2; ti.geom[1] = r.G2p; ti.geom[2] = r.gpf; ti.geom[3] = r.R;
                                                    return set(p, TD, r.elems(), r.batch() * r.G2p, r.Rp, r.C, r.pf); };
  if (name == "Z") return set(Z, 0, (long)s.B * s.Tf * 2048 * 4, s.B, s.Tf, 2048, 4);
  if (name == "ms_spec") return set(ms_spec, 0, 2L * s.B, s.B, 2, 1, 1);
  if (name == "ms_wav") return set(ms_wav, 0, 2L * s.B, s.B, 2, 1, 1);
  if (name == "xf0") return rs(xf0, xf0_rs);
  if (name == "xt0") return rs(xt0, xt0_rs);
  for (int i = 0; i < 4; ++i) {
    std::string k = std::to_string(i);
    if (name == "yf" + k) return rs(yf[i], yf_rs[i]);
    if (name == "yt" + k) return rs(yt[i], yt_rs[i]);
    if (name == "enc" + k) return rs(ef[i], yf_rs[i]);
    if (name == "tenc" + k) return rs(et[i], yt_rs[i]);
    if (name == "fdec" + k) return rs(df[i], df_rs[i]);
    if (name == "tdec" + k) return rs(dt[i], dt_rs[i]);
  }
  if (name == "tokf") return set(tokf, TD, (long)s.B * s.Sf * 512, s.B, s.Sf, 512, 0);
  if (name == "tokt") return set(tokt, TD, (long)s.B * s.St * 512, s.B, s.St, 512, 0);
  if (name == "xenc") return set(xenc, TD, (long)s.B * s.Sf * 384, s.B, s.Sf, 384, 0);
  if (name == "xtenc") return set(xtenc, TD, (long)s.B * s.St * 384, s.B, s.St, 384, 0);
  if (name == "xc") return rs(xc, xc_rs);
  if (name == "xtc") return rs(xtc, xtc_rs);
  if (name == "cvec") return set(cvec, 0, (long)s.B * s.P * 384, s.B * s.P, 384, 1, 0);
  return false;
}

template struct PlanT<float>;
template struct PlanT<bf16>;

}  // namespace athtd
