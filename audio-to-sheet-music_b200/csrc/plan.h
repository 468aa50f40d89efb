// Plan = shapes + workspace layout + launch sequence for one (B segments, L samples, P prompts) shape.
#pragma once
#include <algorithm>
#include "kernels.cuh"
#include "model.cuh"

namespace athtd {

struct PackLayout {
  std::vector<PackItem> items;
  std::map<std::string, int> index;
  long total_bytes;
  explicit PackLayout(int dtype);
};

template <typename T>
int pack_weights(const ParamTable& pt, const PackLayout& pl, const float* params, void* packed, cudaStream_t st);

struct Shapes {
  int B, L, P, Tf;
  int Fr[5];   // frequency rows per level: 2048, 512, 128, 32, 8
  int Lt[5];   // time-branch lengths per level: L, ceil(L/4), ...
  int Sf, St;  // transformer tokens per branch
  Shapes(int B, int L, int P);
};

struct PlanConsts {
  const float2* tw;    // exp(-2 pi i k / 4096), k in [0,4096)
  const float* win;    // periodic hann(4096)
  const float* pe2d;   // [Sf, 512] 2-D sinusoidal embedding in (t1 fr) token order
  const float* pe1d;   // [St, 512]
};

struct TapInfo { const void* ptr; int dtype; long numel; int dims[4]; int geom[4]; };

enum { CONV_ROWS = 0, CONV_K8S4 = 1, CONV_T = 2, CONV_DIL3 = 3 };

// One convolution-shaped op on channels-last row spaces (lowered to the tcgen05 or the SIMT GEMM by PlanT::conv).
template <typename T>
struct ConvOp {
  int mode;                 // CONV_ROWS: 1x1 / linear; CONV_K8S4: kernel 8 stride 4 pad 2; CONV_T: transposed k8 s4 p2
                            // (4 phases, output in the input's geometry); CONV_DIL3: kernel 3, dilation dil, along time
  const T* a; RowSpace as;  // input buffer / space (as.C input channels)
  int dil;
  const T* w; int N;        // packed weights [N, K]
  T* c; RowSpace cs;        // output buffer / space
  const float* bias; int act; int glu; const float* colscale; const T* res;
  const float* rowtab; float rowtab_scale; double* stats; int stat_mode;
  int n_store, no_store, skip_lo, skip_hi; const float* gn_mr; const float* gn_w; const float* gn_b; int gn_mode;   // tensor-core path only
  float* fin_mr; double fin_count; unsigned* fin_counter;   // per-segment statistics finalised by the GEMM itself (gemm.cuh: TcFlat)
};

struct PlanBase {
  virtual ~PlanBase() {}
  virtual int forward(const float* wav, const float* emb, float* out, cudaStream_t st) = 0;
  virtual int encode_only(const float* wav, cudaStream_t st, const float* xnorm = nullptr) = 0;
  virtual int decode_only(const float* emb, float* out, cudaStream_t st) = 0;
  virtual bool tap(const std::string& name, TapInfo& ti) const = 0;
  virtual long workspace_bytes() const = 0;
  virtual long zero_region_bytes() const = 0;
  virtual int launches() const = 0;
  virtual int set_batch(int B) = 0;
  virtual void set_use_graph(bool on) = 0;
  virtual int graph_replays() const = 0;
  virtual void set_profile(bool on) = 0;
  virtual void set_use_tc(bool on) = 0;
  virtual void set_use_flash(bool on) = 0;
  virtual void set_use_fused_dconv(bool on) = 0;
  virtual int tc_launches() const = 0;
  virtual void get_profile(double* ms, double* gflop, int* n) = 0;
};

template <typename T>
struct PlanT : PlanBase {
  Shapes sh;            // sh.B = batch of the next forward (<= cap)
  const ParamTable* pt;
  const PackLayout* pl;
  const float* params;
  const char* packed;
  PlanConsts consts;
  int cap;              // batch capacity the workspace was laid out for
  char* base_ = nullptr;
  size_t total_bytes = 0, zero_bytes = 0, stats_begin = 0, stats_bytes = 0, dec_stats_begin = 0;
  int n_launches = 0, n_tc = 0, n_enc_launches = 0, n_enc_tc = 0;   // launches(): encode + the last decode
  bool use_tc = true, use_flash = true, use_fused_dconv = true;
  bool profiling = false;
  const float* cur_wav = nullptr;   // waveform of the forward in flight (time-branch level 0 reads it directly)
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  double prof_gflop = 0.0;
  std::vector<double> prof_items;
  std::vector<std::string> prof_labels;
  // CUDA-graph replay of the launch sequence.  A forward whose (batch, wav, emb, out) was seen
  // before is captured once on a private stream and replayed from then on: a track loop that reuses its staging buffers
  // pays ~160 kernel launches per batch once.  Callers with ever-changing pointers simply stay on the eager path.
  struct GraphEntry { int B; const float* wav; const float* emb; float* out; cudaGraphExec_t exec; int n_launches, n_tc; unsigned long stamp; };
  std::vector<GraphEntry> graphs;
  bool use_graph = true, graph_broken = false;
  cudaStream_t cap_stream = nullptr;
  unsigned long graph_clock = 0;
  int n_graph_replays = 0;
  void drop_graphs();
  ~PlanT() override;
  // buffers
  RowSpace xf0_rs, xt0_rs, yf_rs[4], yt_rs[4], xc_rs, xtc_rs, df_rs[4], dt_rs[4];
  T *xf0, *xt0, *yf[4], *ef[4], *yt[4], *et[4], *xc, *xtc, *df[4], *dt[4];
  double *st_spec, *st_wav, *st_df[4][2][2], *st_dt[4][2][2], *st_xf[5][2], *st_dec;
  unsigned *fin_xf, *fin_dec;      // tickets of the fused statistics finalisation: [5][2] / [P][6], zeroed with the statistics
  float *mr, *ms_spec, *ms_wav, *Z, *cvec;
  T *hbuf, *ebuf, *tokf, *tokt, *hn[5], *qkv, *kvb, *obuf, *ffn, *scores, *xenc, *xtenc, *t1, *t2, *ubuf;

  PlanT(int B, int L, int P, const ParamTable* pt, const PackLayout* pl, const float* params, const void* packed,
        void* workspace, const PlanConsts& c);
  void layout(char* base);
  void spaces(int B);
  int set_batch(int B) override;
  char* ws_base() const { return base_; }
  const float* P32(const std::string& name) const;
  const T* PW(const std::string& key) const;
  const float* PA(const std::string& key) const;
  void gemm(const GemmDesc& d, cudaStream_t st);
  bool conv(const ConvOp<T>& o, cudaStream_t st);      // true: the launch also finalised o.stats into o.fin_mr
  void prof_begin(double gflop, cudaStream_t st, const char* what, long m, int n, int k);
  void prof_end(cudaStream_t st);
  bool enc_row_dispatch(bool run, int i, const T* x, RowSpace xin, const T* y, T* out, RowSpace ys, cudaStream_t st);
  void enc_layer(bool freq, int i, const T* x, RowSpace xin, T* y, RowSpace ys, T* out, cudaStream_t st);
  void attention(const T* q, long ldq, const T* k, const T* v, long ldkv, int Sq, int Sk, T* o, cudaStream_t st);
  void linear(const T* a, int S, int K, const T* w, int N, const float* bias, int act, T* c, cudaStream_t st);
  bool linear_res(const T* a, int S, int K, const T* w, int N, const float* bias, const float* gamma, T* x, double* stats,
                  cudaStream_t st, unsigned* fin_counter = nullptr);      // true: stats already finalised into mr
  void xf_ffn_and_norm(const std::string& p, const char* ffn_norm, T* x, int S, double* stats, unsigned* fin_counter, T* n1,
                       const float* n1w, const float* n1b, T* n2, const float* n2w, const float* n2b, cudaStream_t st);
  void cross_transformer(cudaStream_t st);
  void encode(const float* wav, cudaStream_t st, const float* xnorm = nullptr);
  void text_vectors(const float* emb, cudaStream_t st);
  void text_condition(int p, const T* x, int S, T* out, RowSpace outs, cudaStream_t st);
  void dec_layer(bool freq, int i, int p, const T* x, RowSpace xs, T* out, RowSpace os, const T* skip, RowSpace ss,
                 cudaStream_t st);
  void decode(const float* emb, float* out, cudaStream_t st);

  int forward(const float* wav, const float* emb, float* out, cudaStream_t st) override;
  int encode_only(const float* wav, cudaStream_t st, const float* xnorm = nullptr) override {
    n_launches = 0; n_tc = 0;
    encode(wav, st, xnorm);
    n_enc_launches = n_launches; n_enc_tc = n_tc;
    return (int)cudaGetLastError();
  }
  int decode_only(const float* emb, float* out, cudaStream_t st) override {
    n_launches = n_enc_launches; n_tc = n_enc_tc;
    decode(emb, out, st);
    return (int)cudaGetLastError();
  }
  bool tap(const std::string& name, TapInfo& ti) const override;
  long workspace_bytes() const override { return (long)total_bytes; }
  long zero_region_bytes() const override { return (long)zero_bytes; }
  int launches() const override { return n_launches; }
  void set_profile(bool on) override { profiling = on; prof_used = 0; prof_gflop = 0.0; prof_items.clear(); prof_labels.clear(); }
  void set_use_tc(bool on) override { use_tc = on; drop_graphs(); }
  void set_use_flash(bool on) override { use_flash = on; drop_graphs(); }
  void set_use_fused_dconv(bool on) override { use_fused_dconv = on; drop_graphs(); }
  void set_use_graph(bool on) override { use_graph = on; if (!on) drop_graphs(); }
  int graph_replays() const override { return n_graph_replays; }
  int tc_launches() const override { return n_tc; }
  void get_profile(double* ms, double* gflop, int* n) override;
};

}  // namespace athtd
