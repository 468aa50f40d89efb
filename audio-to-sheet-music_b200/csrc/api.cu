// extern "C" boundary (include/athtd.h).  Never throws, never allocates device memory, never syncs.
#include "plan.h"
#include "../../include/athtd.h"
#include <string>
#include <mutex>
#include <string.h>

using namespace athtd;

static thread_local std::string g_err;
static const ParamTable& param_table() { static ParamTable t = build_param_table(); return t; }
static const PackLayout& pack_layout(int dtype) {
  static PackLayout l0(0), l1(1);
  return dtype == 0 ? l0 : l1;
}
static int fail(const std::string& m) { g_err = m; return 1; }
static int check_cuda(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}

extern "C" {

const char* athtd_last_error(void) { return g_err.c_str(); }
int athtd_version(void) { return 1; }

int athtd_param_count(void) { return (int)param_table().items.size(); }
const char* athtd_param_name(int i) { return param_table().items[i].name.c_str(); }
long athtd_param_numel(int i) { return param_table().items[i].numel; }
long athtd_param_offset(int i) { return param_table().items[i].offset; }
long athtd_params_total(void) { return param_table().total; }

long athtd_packed_bytes(int dtype) { return pack_layout(dtype).total_bytes; }

int athtd_pack_weights(const float* params_dev, void* packed_dev, int dtype, void* stream) {
  try {
    if (dtype == 0) pack_weights<float>(param_table(), pack_layout(0), params_dev, packed_dev, (cudaStream_t)stream);
    else if (dtype == 1) pack_weights<bf16>(param_table(), pack_layout(1), params_dev, packed_dev, (cudaStream_t)stream);
    else return fail("athtd_pack_weights: dtype must be 0 (fp32) or 1 (bf16)");
  } catch (const std::exception& e) { return fail(e.what()); }
  return check_cuda("athtd_pack_weights");
}

static int check_shape(int B, int L, int P, int dtype) {
  if (B < 1 || P < 1) return fail("athtd: B and P must be >= 1");
  if (L < 4096) return fail("athtd: segment length must be >= 4096 samples");
  if (dtype != 0 && dtype != 1) return fail("athtd: dtype must be 0 (fp32) or 1 (bf16)");
  return 0;
}

long athtd_workspace_bytes(int B, int L, int P, int dtype) {
  if (check_shape(B, L, P, dtype)) return -1;
  PlanConsts c{};
  if (dtype == 0) { PlanT<float> p(B, L, P, &param_table(), &pack_layout(0), nullptr, nullptr, nullptr, c); return p.workspace_bytes(); }
  PlanT<bf16> p(B, L, P, &param_table(), &pack_layout(1), nullptr, nullptr, nullptr, c);
  return p.workspace_bytes();
}

void* athtd_plan_create(int B, int L, int P, int dtype, const float* params_dev, const void* packed_dev, void* workspace_dev,
                        long workspace_bytes, const float* tw_dev, const float* win_dev, const float* pe2d_dev,
                        const float* pe1d_dev) {
  if (check_shape(B, L, P, dtype)) return nullptr;
  int dev = 0; cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    fail("athtd_plan_create: no CUDA device (this library has no CPU fallback)"); return nullptr;
  }
  if (prop.major != 10) { fail("athtd_plan_create: built for sm_100a (B200) only"); return nullptr; }
  PlanConsts c{(const float2*)tw_dev, win_dev, pe2d_dev, pe1d_dev};
  PlanBase* p = nullptr;
  try {
    if (dtype == 0) p = new PlanT<float>(B, L, P, &param_table(), &pack_layout(0), params_dev, packed_dev, workspace_dev, c);
    else p = new PlanT<bf16>(B, L, P, &param_table(), &pack_layout(1), params_dev, packed_dev, workspace_dev, c);
  } catch (const std::exception& e) { fail(e.what()); return nullptr; }
  if (p->workspace_bytes() > workspace_bytes) { delete p; fail("athtd_plan_create: workspace too small"); return nullptr; }
  return p;
}
void athtd_plan_destroy(void* plan) { delete (PlanBase*)plan; }

int athtd_plan_tokens(void* plan, int* Tf, int* Sf, int* St) {
  TapInfo ti;
  PlanBase* p = (PlanBase*)plan;
  if (!p->tap("tokf", ti)) return fail("tokf"); *Sf = ti.dims[1];
  if (!p->tap("tokt", ti)) return fail("tokt"); *St = ti.dims[1];
  *Tf = *Sf / 8;
  return 0;
}

#define GUARD(expr, tag)                                          \
  try { expr; } catch (const std::exception& e) { return fail(e.what()); } \
  return check_cuda(tag);

int athtd_forward(void* plan, const float* wav_dev, const float* emb_dev, float* out_dev, void* stream) {
  GUARD(((PlanBase*)plan)->forward(wav_dev, emb_dev, out_dev, (cudaStream_t)stream), "athtd_forward");
}
int athtd_encode(void* plan, const float* wav_dev, void* stream) {
  GUARD(((PlanBase*)plan)->encode_only(wav_dev, (cudaStream_t)stream), "athtd_encode");
}
int athtd_encode_normalized(void* plan, const float* x_cac_dev, const float* xt_dev, void* stream) {
  if (!x_cac_dev || !xt_dev) return fail("athtd_encode_normalized: null input");
  GUARD(((PlanBase*)plan)->encode_only(xt_dev, (cudaStream_t)stream, x_cac_dev), "athtd_encode_normalized");
}
int athtd_decode(void* plan, const float* emb_dev, float* out_dev, void* stream) {
  GUARD(((PlanBase*)plan)->decode_only(emb_dev, out_dev, (cudaStream_t)stream), "athtd_decode");
}
int athtd_plan_launches(void* plan) { return ((PlanBase*)plan)->launches(); }
int athtd_plan_set_batch(void* plan, int B) {
  if (((PlanBase*)plan)->set_batch(B)) return fail("athtd_plan_set_batch: B must be in [1, batch capacity of the plan]");
  return 0;
}
int athtd_plan_set_profile(void* plan, int on) { ((PlanBase*)plan)->set_profile(on != 0); return 0; }
int athtd_plan_get_profile(void* plan, double* gemm_ms, double* gemm_gflop, int* gemm_launches) {
  ((PlanBase*)plan)->get_profile(gemm_ms, gemm_gflop, gemm_launches);
  return 0;
}

int athtd_tap(void* plan, const char* name, const void** ptr, long* numel, int* dtype, int dims[8]) {
  TapInfo ti;
  memset(&ti, 0, sizeof(ti));
  if (!((PlanBase*)plan)->tap(name, ti)) return fail(std::string("athtd_tap: unknown buffer ") + name);
  *ptr = ti.ptr; *numel = ti.numel; *dtype = ti.dtype;
  for (int i = 0; i < 4; ++i) { dims[i] = ti.dims[i]; dims[4 + i] = ti.geom[i]; }
  return 0;
}
int athtd_plan_set_tc(void* plan, int on) { ((PlanBase*)plan)->set_use_tc(on != 0); return 0; }
int athtd_plan_set_flash(void* plan, int on) { ((PlanBase*)plan)->set_use_flash(on != 0); return 0; }
int athtd_plan_set_fused_dconv(void* plan, int on) { ((PlanBase*)plan)->set_use_fused_dconv(on != 0); return 0; }
int athtd_plan_tc_launches(void* plan) { return ((PlanBase*)plan)->tc_launches(); }
int athtd_plan_set_graph(void* plan, int on) { ((PlanBase*)plan)->set_use_graph(on != 0); return 0; }
int athtd_plan_graph_replays(void* plan) { return ((PlanBase*)plan)->graph_replays(); }

int athtd_memcpy_d2d(void* dst_dev, const void* src_dev, long bytes, void* stream) {
  cudaError_t e = cudaMemcpyAsync(dst_dev, src_dev, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(std::string("athtd_memcpy_d2d: ") + cudaGetErrorString(e));
  return 0;
}

static const ParamTable& clap_table() { static ParamTable t = build_clap_param_table(); return t; }
int athtd_clap_param_count(void) { return (int)clap_table().items.size(); }
const char* athtd_clap_param_name(int i) { return clap_table().items[i].name.c_str(); }
long athtd_clap_param_numel(int i) { return clap_table().items[i].numel; }
long athtd_clap_param_offset(int i) { return clap_table().items[i].offset; }
long athtd_clap_params_total(void) { return clap_table().total; }
long athtd_clap_workspace_bytes(int P, int S) { return clap_workspace_bytes(P, S); }
int athtd_clap_text_forward(const float* params_dev, const long* input_ids_dev, const long* attention_mask_dev, int P, int S,
                            void* workspace_dev, float* out_dev, int normalize, void* stream) {
  if (P <= 0 || S <= 0 || S > 512) return fail("athtd_clap_text_forward: need 1 <= S <= 512 tokens and P >= 1 prompts");
  try {
    int rc = clap_text_forward(clap_table(), params_dev, input_ids_dev, attention_mask_dev, P, S, workspace_dev, out_dev, normalize,
                               (cudaStream_t)stream);
    if (rc != 0) return fail(std::string("athtd_clap_text_forward: ") + cudaGetErrorString((cudaError_t)rc));
  } catch (const std::exception& e) { return fail(e.what()); }
  return 0;
}

int athtd_sdr_sums(const float* est_dev, const float* tgt_dev, int items, long n, double* sums_dev, void* stream) {
  if (items <= 0 || n <= 0) return fail("athtd_sdr_sums: items and n must be positive");
  launch_sdr_sums(est_dev, tgt_dev, items, n, sums_dev, (cudaStream_t)stream);
  return check_cuda("athtd_sdr_sums");
}

int athtd_stft_cac(const float* wav_dev, int B, int L, float* Z_dev, double* stats_dev, const float* tw_dev,
                   const float* win_dev, void* stream) {
  if (L < 4096) return fail("athtd_stft_cac: L must be >= 4096");
  launch_stft_cac(wav_dev, B, L, (L + 1023) / 1024, Z_dev, stats_dev, (const float2*)tw_dev, win_dev, (cudaStream_t)stream);
  return check_cuda("athtd_stft_cac");
}

int athtd_istft(const float* Z_dev, int B, int L, float* frames_dev, float* out_dev, const float* tw_dev,
                const float* win_dev, void* stream) {
  const int Tf = (L + 1023) / 1024;
  RowSpace none{};
  (void)frames_dev;      // kept in the signature for ABI stability: the fused inverse needs no frame scratch
  launch_istft_fused<float>(Z_dev, Tf, L, B, 1, nullptr, none, 0, nullptr, nullptr, nullptr, none, nullptr, nullptr, nullptr, 1,
                            out_dev, 2L * L, (const float2*)tw_dev, win_dev, (cudaStream_t)stream);
  return check_cuda("athtd_istft");
}

int athtd_gather_chunks(const float* track_dev, long T, int C, const long* starts_dev, int n_chunks, int chunk_len,
                        float* segs_dev, void* stream) {
  if (n_chunks < 0 || chunk_len < 1 || C < 1) return fail("athtd_gather_chunks: need n_chunks >= 0, chunk_len >= 1, C >= 1");
  if (n_chunks == 0) return 0;      // empty span (more ranks than chunks)
  launch_gather_chunks(track_dev, T, C, starts_dev, n_chunks, chunk_len, segs_dev, (cudaStream_t)stream);
  return check_cuda("athtd_gather_chunks");
}

int athtd_chunk_ola(const float* seg_out_dev, long seg_stride, int k_base, int chunk_len, const long* starts_dev,
                    const int* actual_len_dev, const int* fade_len_dev, const int* flags_dev, int n_chunks, long stride,
                    const float* ramp_up_dev, const float* ramp_down_dev, const int* ramp_off_dev, float* out_dev, int C,
                    long t_begin, long t_end, void* stream) {
  if (C < 1 || C > 2) return fail("athtd_chunk_ola: C must be 1 or 2");
  if (t_end <= t_begin) return 0;
  launch_chunk_ola(seg_out_dev, seg_stride, k_base, chunk_len, starts_dev, actual_len_dev, fade_len_dev, flags_dev, n_chunks,
                   stride, ramp_up_dev, ramp_down_dev, ramp_off_dev, out_dev, C, t_begin, t_end, 1, (cudaStream_t)stream);
  return check_cuda("athtd_chunk_ola");
}

int athtd_chunk_fade_add(const float* seg_out_dev, long seg_stride, int k_base, int chunk_len, const long* starts_dev,
                         const int* actual_len_dev, const int* fade_len_dev, const int* flags_dev, int n_chunks, long stride,
                         const float* ramp_up_dev, const float* ramp_down_dev, const int* ramp_off_dev, float* out_dev, int C,
                         long t_begin, long t_end, void* stream) {
  if (C < 1 || C > 2) return fail("athtd_chunk_fade_add: C must be 1 or 2");
  if (t_end <= t_begin) return 0;
  launch_chunk_ola(seg_out_dev, seg_stride, k_base, chunk_len, starts_dev, actual_len_dev, fade_len_dev, flags_dev, n_chunks,
                   stride, ramp_up_dev, ramp_down_dev, ramp_off_dev, out_dev, C, t_begin, t_end, 0, (cudaStream_t)stream);
  return check_cuda("athtd_chunk_fade_add");
}

int athtd_load_audio(const float* x_dev, int C_in, long T_in, const float* kernel_t_dev, int orig, int new_, int taps, int width,
                     float* y_dev, int C_out, long T_out, void* stream) {
  if (C_in < 1 || C_out < C_in || T_in < 1) return fail("athtd_load_audio: need 1 <= C_in <= C_out and T_in >= 1");
  if (kernel_t_dev) {
    if (orig < 1 || new_ < 1 || width < 0 || taps != 2 * width + orig) return fail("athtd_load_audio: taps must be 2 * width + orig (torchaudio kernel shape)");
    const long want = ((long)new_ * T_in + orig - 1) / orig;
    if (T_out != want) return fail("athtd_load_audio: T_out must be ceil(new * T_in / orig)");
  } else if (T_out != T_in) return fail("athtd_load_audio: without a filter bank T_out must equal T_in");
  if (launch_resample(x_dev, C_in, T_in, kernel_t_dev, orig, new_, taps, width, y_dev, C_out, T_out, (cudaStream_t)stream))
    return fail("athtd_load_audio: rate ratio needs more than 200 KB of shared memory per block");
  return check_cuda("athtd_load_audio");
}

int athtd_set_pdl(int on) { pdl_set_enabled(on != 0); return 0; }
int athtd_set_tc_tuning(int flags) { tc_set_bn_cap(flags); return 0; }
int athtd_attention_set_poly(int npoly) { flash_attn_set_poly(npoly); return 0; }

int athtd_attention_test(const void* q_dev, const void* k_dev, const void* v_dev, void* o_dev, int B, int Sq, int Sk, void* stream) {
  if (!flash_attn_supported(512, 512, 512)) return fail("athtd_attention_test: tensor-map API unavailable");
  int rc = launch_flash_attn((const bf16*)q_dev, 512, (const bf16*)k_dev, (const bf16*)v_dev, 512, B, Sq, Sk, (bf16*)o_dev, 512,
                             (cudaStream_t)stream);
  if (rc != 0) return fail("athtd_attention_test: cuTensorMapEncodeTiled failed");
  return check_cuda("athtd_attention_test");
}

int athtd_gemm_test(const void* A_dev, const void* B_dev, const float* bias_dev, void* C_dev, int M, int N, int K, int dtype,
                    int use_tensor_cores, void* stream) {
  GemmDesc d = gemm_desc_zero();
  d.Mg = M; d.N = N; d.K = K; d.Ktap = K; d.A = A_dev; d.sAm = K; d.B = B_dev; d.sBn = K; d.sBk = 1;
  d.C = C_dev; d.sCm = N; d.bias = bias_dev;
  if (use_tensor_cores) {
    if (dtype != 1) return fail("athtd_gemm_test: tensor-core path is bf16");
    TcFlat f;
    memset(&f, 0, sizeof(f));
    f.A = A_dev; f.a_rows = M; f.a_pitch = K; f.Ktap = K; f.ntaps = 1; f.B = B_dev; f.N = N; f.Mflat = M;
    f.RpA = M; f.G2p = 1; f.gpf = 0; f.G2 = 1; f.vlo = 0; f.vhi = M; f.oG2p = 1; f.ogsh = 0; f.oRp = M; f.orsh = 0;
    f.ldc = N; f.C = C_dev; f.alpha = 1.0f; f.bias = bias_dev;
    if (use_tensor_cores == 2) f.no_store = 1;     // micro-benchmark variants (tools/gemm_bench.py)
    if (use_tensor_cores == 3) { f.dbg = (void*)bias_dev; f.bias = nullptr; }   // bias_dev doubles as a long long[512*8] stamp buffer
    if (!tc_flat_supported(f)) return fail("athtd_gemm_test: shape not supported by the tcgen05 kernel");
    if (launch_gemm_tc_flat(f, (cudaStream_t)stream) != 0) return fail("athtd_gemm_test: cuTensorMapEncodeTiled failed");
  } else if (dtype == 0) launch_gemm_simt<float>(d, (cudaStream_t)stream);
  else launch_gemm_simt<bf16>(d, (cudaStream_t)stream);
  return check_cuda("athtd_gemm_test");
}

}  // extern "C"
