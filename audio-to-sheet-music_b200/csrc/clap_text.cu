// CLAP text tower: the encoder behind AudioTextHTDemucs._get_clap_embeddings
// (/root/reference/src/models/stem_separation/ATHTDemucs_v2.py:238-248; HF transformers ClapTextModelWithProjection /
// ClapModel.get_text_features, structure printed at src/models/stem_separation/AudioTextHTDemucs_Full.txt:630-823):
// RoBERTa-base (12 layers, hidden 768, 12 heads, FFN 3072, exact GELU, LayerNorm eps 1e-12, learned positions offset by
// padding_idx = 1) -> tanh pooler on the first token -> ClapProjectionLayer (768 -> 512, ReLU, 512 -> 512) [-> L2 norm].
// A handful of prompts x <= ~20 tokens: every GEMM has M = P*S rows, so this runs on the fp32 CUDA-core GEMM
// (gemm_simt.cu) with bias / GELU / residual epilogues; everything stays fp32 (embeddings feed every segment of a track).
#include "kernels.cuh"
#include "model.cuh"
#include <string>

namespace athtd {

static const int CL_H = 768, CL_L = 12, CL_NH = 12, CL_FF = 3072, CL_V = 50265, CL_POS = 514, CL_PAD = 1, CL_PROJ = 512;

ParamTable build_clap_param_table() {
  ParamTable t;
  auto wb = [&](const std::string& n, long w, long b) { t.add(n + ".weight", w); t.add(n + ".bias", b); };
  const std::string e = "text_model.embeddings.";
  t.add(e + "word_embeddings.weight", (long)CL_V * CL_H);
  t.add(e + "position_embeddings.weight", (long)CL_POS * CL_H);
  t.add(e + "token_type_embeddings.weight", CL_H);
  wb(e + "LayerNorm", CL_H, CL_H);
  for (int l = 0; l < CL_L; ++l) {
    const std::string p = "text_model.encoder.layer." + std::to_string(l) + ".";
    wb(p + "attention.self.query", (long)CL_H * CL_H, CL_H);
    wb(p + "attention.self.key", (long)CL_H * CL_H, CL_H);
    wb(p + "attention.self.value", (long)CL_H * CL_H, CL_H);
    wb(p + "attention.output.dense", (long)CL_H * CL_H, CL_H);
    wb(p + "attention.output.LayerNorm", CL_H, CL_H);
    wb(p + "intermediate.dense", (long)CL_FF * CL_H, CL_FF);
    wb(p + "output.dense", (long)CL_H * CL_FF, CL_H);
    wb(p + "output.LayerNorm", CL_H, CL_H);
  }
  wb("text_model.pooler.dense", (long)CL_H * CL_H, CL_H);
  wb("text_projection.linear1", (long)CL_PROJ * CL_H, CL_PROJ);
  wb("text_projection.linear2", (long)CL_PROJ * CL_PROJ, CL_PROJ);
  return t;
}

// embeddings: word[id] + position[pos] + token_type[0], LayerNorm; pos = (number of non-pad tokens up to and including
// this one) * (id != pad) + pad_idx   (RobertaEmbeddings.create_position_ids_from_input_ids).  One block per token.
__global__ void __launch_bounds__(256) clap_embed_kernel(const long* __restrict__ ids, int S, const float* __restrict__ word,
                                                         const float* __restrict__ posw, const float* __restrict__ typew,
                                                         const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                                                         float* __restrict__ out) {
  const int tok = blockIdx.x, p = tok / S, s = tok - p * S;
  __shared__ float red[2][8];
  __shared__ int pos_s;
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int j = 0; j <= s; ++j) cnt += ids[(long)p * S + j] != CL_PAD ? 1 : 0;
    pos_s = (ids[(long)p * S + s] != CL_PAD ? cnt : 0) + CL_PAD;
  }
  __syncthreads();
  const long id = ids[tok];
  const float* w = word + id * CL_H;
  const float* pw = posw + (long)pos_s * CL_H;
  float v[3];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { const int c = threadIdx.x + 256 * i; v[i] = w[c] + typew[c] + pw[c]; sum += v[i]; }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = sum;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[0][i];
  const float mean = tot / CL_H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { const float d = v[i] - mean; q += d * d; }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = q;
  __syncthreads();
  float qt = 0.f;
  for (int i = 0; i < 8; ++i) qt += red[1][i];
  const float rstd = rsqrtf(qt / CL_H + eps);
#pragma unroll
  for (int i = 0; i < 3; ++i) { const int c = threadIdx.x + 256 * i; out[(long)tok * CL_H + c] = (v[i] - mean) * rstd * lnw[c] + lnb[c]; }
}

// y = LayerNorm(x) (x already holds dense(.) + residual from the GEMM epilogue); one block of 256 per row of 768
__global__ void __launch_bounds__(256) clap_ln_kernel(const float* __restrict__ x, const float* __restrict__ lnw,
                                                      const float* __restrict__ lnb, float eps, float* __restrict__ y) {
  const long row = blockIdx.x;
  __shared__ float red[2][8];
  float v[3];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = x[row * CL_H + threadIdx.x + 256 * i]; sum += v[i]; }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = sum;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[0][i];
  const float mean = tot / CL_H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { const float d = v[i] - mean; q += d * d; }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = q;
  __syncthreads();
  float qt = 0.f;
  for (int i = 0; i < 8; ++i) qt += red[1][i];
  const float rstd = rsqrtf(qt / CL_H + eps);
#pragma unroll
  for (int i = 0; i < 3; ++i) { const int c = threadIdx.x + 256 * i; y[row * CL_H + c] = (v[i] - mean) * rstd * lnw[c] + lnb[c]; }
}

// scores [P, NH, S, S] (already scaled by 1/8): + (1 - mask[p, key]) * finfo.min, softmax over keys; one warp per row
__global__ void clap_softmax_kernel(float* __restrict__ sc, const long* __restrict__ mask, int S, long rows) {
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int p = (int)(row / ((long)CL_NH * S));
  float* r = sc + row * S;
  float mx = -INFINITY;
  for (int j = lane; j < S; j += 32) {
    const float v = r[j] + (mask[(long)p * S + j] ? 0.f : -3.4028234663852886e38f);
    r[j] = v; mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < S; j += 32) { const float e = expf(r[j] - mx); r[j] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < S; j += 32) r[j] *= inv;
}

__global__ void clap_act_kernel(float* __restrict__ x, long n, int mode) {       // 1 tanh, 2 relu
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    x[i] = mode == 1 ? tanhf(x[i]) : fmaxf(x[i], 0.f);
}
__global__ void clap_l2norm_kernel(float* __restrict__ x, int n) {               // one warp per row of n
  float* r = x + (long)blockIdx.x * n;
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 32) s += r[j] * r[j];
  s = warp_sum(s);
  const float inv = 1.0f / sqrtf(s);
  for (int j = threadIdx.x; j < n; j += 32) r[j] *= inv;
}

long clap_workspace_bytes(int P, int S) {
  const long rows = (long)P * S;
  // x, x2, q, k, v, ctx [rows,768] ; ffn [rows,3072] ; scores [P,12,S,S] ; pooled [P,768] ; proj [P,512]
  return sizeof(float) * (6 * rows * CL_H + rows * CL_FF + (long)P * CL_NH * S * S + (long)P * CL_H + (long)P * CL_PROJ) + 1024;
}

static void cl_linear(const float* a, long a_row_stride, long M, int K, const float* w, const float* b, int N, int act,
                      const float* res, float* c, cudaStream_t st) {
  GemmDesc d = gemm_desc_zero();
  d.Mg = (int)M; d.N = N; d.K = K; d.Ktap = K; d.A = a; d.sAm = a_row_stride; d.B = w; d.sBn = K; d.sBk = 1;
  d.C = c; d.sCm = N; d.bias = b; d.act = act; d.c_is_f32 = 1;
  if (res) { d.res = res; d.sRm = N; }
  launch_gemm_simt<float>(d, st);
}

int clap_text_forward(const ParamTable& pt, const float* params, const long* ids, const long* mask, int P, int S, void* workspace,
                      float* out, int normalize, cudaStream_t st) {
  const long rows = (long)P * S;
  float* x = (float*)workspace;
  float* x2 = x + rows * CL_H;
  float* q = x2 + rows * CL_H;
  float* k = q + rows * CL_H;
  float* v = k + rows * CL_H;
  float* ctx = v + rows * CL_H;
  float* ffn = ctx + rows * CL_H;
  float* sc = ffn + rows * CL_FF;
  float* pooled = sc + (long)P * CL_NH * S * S;
  float* proj = pooled + (long)P * CL_H;
  auto W = [&](const std::string& n) { return params + pt.off(n); };
  const std::string e = "text_model.embeddings.";
  clap_embed_kernel<<<(unsigned)rows, 256, 0, st>>>(ids, S, W(e + "word_embeddings.weight"), W(e + "position_embeddings.weight"),
                                                    W(e + "token_type_embeddings.weight"), W(e + "LayerNorm.weight"),
                                                    W(e + "LayerNorm.bias"), 1e-12f, x);
  for (int l = 0; l < CL_L; ++l) {
    const std::string p = "text_model.encoder.layer." + std::to_string(l) + ".";
    cl_linear(x, CL_H, rows, CL_H, W(p + "attention.self.query.weight"), W(p + "attention.self.query.bias"), CL_H, ACT_NONE, nullptr, q, st);
    cl_linear(x, CL_H, rows, CL_H, W(p + "attention.self.key.weight"), W(p + "attention.self.key.bias"), CL_H, ACT_NONE, nullptr, k, st);
    cl_linear(x, CL_H, rows, CL_H, W(p + "attention.self.value.weight"), W(p + "attention.self.value.bias"), CL_H, ACT_NONE, nullptr, v, st);
    {  // scores[p, h] = q_h k_h^T / 8
      GemmDesc d = gemm_desc_zero();
      d.G1 = P; d.G2 = CL_NH; d.Mg = S; d.N = S; d.K = 64; d.Ktap = 64; d.grouped = 1;
      d.A = q; d.sAg1 = (long)S * CL_H; d.sAg2 = 64; d.sAm = CL_H;
      d.B = k; d.sBg1 = (long)S * CL_H; d.sBg2 = 64; d.sBn = CL_H; d.sBk = 1;
      d.C = sc; d.sCg1 = (long)CL_NH * S * S; d.sCg2 = (long)S * S; d.sCm = S; d.alpha = 0.125f; d.c_is_f32 = 1;
      launch_gemm_simt<float>(d, st);
    }
    clap_softmax_kernel<<<(unsigned)(((long)P * CL_NH * S + 7) / 8), 256, 0, st>>>(sc, mask, S, (long)P * CL_NH * S);
    {  // ctx[p, :, h] = probs v_h
      GemmDesc d = gemm_desc_zero();
      d.G1 = P; d.G2 = CL_NH; d.Mg = S; d.N = 64; d.K = S; d.Ktap = S; d.grouped = 1;
      d.A = sc; d.sAg1 = (long)CL_NH * S * S; d.sAg2 = (long)S * S; d.sAm = S;
      d.B = v; d.sBg1 = (long)S * CL_H; d.sBg2 = 64; d.sBn = 1; d.sBk = CL_H;
      d.C = ctx; d.sCg1 = (long)S * CL_H; d.sCg2 = 64; d.sCm = CL_H; d.c_is_f32 = 1;
      launch_gemm_simt<float>(d, st);
    }
    cl_linear(ctx, CL_H, rows, CL_H, W(p + "attention.output.dense.weight"), W(p + "attention.output.dense.bias"), CL_H, ACT_NONE, x, x2, st);
    clap_ln_kernel<<<(unsigned)rows, 256, 0, st>>>(x2, W(p + "attention.output.LayerNorm.weight"), W(p + "attention.output.LayerNorm.bias"), 1e-12f, x);
    cl_linear(x, CL_H, rows, CL_H, W(p + "intermediate.dense.weight"), W(p + "intermediate.dense.bias"), CL_FF, ACT_GELU, nullptr, ffn, st);
    cl_linear(ffn, CL_FF, rows, CL_FF, W(p + "output.dense.weight"), W(p + "output.dense.bias"), CL_H, ACT_NONE, x, x2, st);
    clap_ln_kernel<<<(unsigned)rows, 256, 0, st>>>(x2, W(p + "output.LayerNorm.weight"), W(p + "output.LayerNorm.bias"), 1e-12f, x);
  }
  // pooler on the first token of every prompt (row stride S*768), projection head
  cl_linear(x, (long)S * CL_H, P, CL_H, W("text_model.pooler.dense.weight"), W("text_model.pooler.dense.bias"), CL_H, ACT_NONE, nullptr, pooled, st);
  clap_act_kernel<<<8, 256, 0, st>>>(pooled, (long)P * CL_H, 1);
  cl_linear(pooled, CL_H, P, CL_H, W("text_projection.linear1.weight"), W("text_projection.linear1.bias"), CL_PROJ, ACT_NONE, nullptr, proj, st);
  clap_act_kernel<<<8, 256, 0, st>>>(proj, (long)P * CL_PROJ, 2);
  cl_linear(proj, CL_PROJ, P, CL_PROJ, W("text_projection.linear2.weight"), W("text_projection.linear2.bias"), CL_PROJ, ACT_NONE, nullptr, out, st);
  if (normalize) clap_l2norm_kernel<<<P, 32, 0, st>>>(out, CL_PROJ);
  return (int)cudaGetLastError();
}

}  // namespace athtd
