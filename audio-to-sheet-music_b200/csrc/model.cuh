// Static description of the live parameters of AudioTextHTDemucs (state_dict layout of the
// reference; /root/reference/src/models/stem_separation/AudioTextHTDemucs_Full.txt)
// and of the packed GEMM-layout weight blob derived from it.
#pragma once
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

namespace athtd {

struct ParamInfo { std::string name; long numel; long offset; };

struct ParamTable {
  std::vector<ParamInfo> items;
  std::map<std::string, int> index;
  long total = 0;
  void add(const std::string& n, long numel) {
    index[n] = (int)items.size();
    items.push_back({n, numel, total});
    total += (numel + 3) / 4 * 4;      // keep every tensor 16-byte aligned
  }
  long off(const std::string& n) const {
    auto it = index.find(n);
    if (it == index.end()) throw std::runtime_error("athtd: unknown parameter " + n);
    return items[it->second].offset;
  }
};

static const int kCh[4] = {48, 96, 192, 384};
static const int kDecCh[5] = {384, 192, 96, 48, 4};

inline ParamTable build_param_table() {
  ParamTable t;
  auto wb = [&](const std::string& n, long w, long b) { t.add(n + ".weight", w); t.add(n + ".bias", b); };
  const char* br[2] = {"encoder", "tencoder"};
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < 4; ++i) {
      int cin = i == 0 ? (f == 0 ? 4 : 2) : kCh[i - 1];
      int c = kCh[i];
      std::string p = std::string("htdemucs.") + br[f] + "." + std::to_string(i);
      wb(p + ".conv", (long)c * cin * 8, c);
      wb(p + ".rewrite", (long)2 * c * c, 2 * c);
      for (int d = 0; d < 2; ++d) {
        std::string q = p + ".dconv.layers." + std::to_string(d);
        wb(q + ".0", (long)(c / 8) * c * 3, c / 8);
        wb(q + ".1", c / 8, c / 8);
        wb(q + ".3", (long)2 * c * (c / 8), 2 * c);
        wb(q + ".4", 2 * c, 2 * c);
        t.add(q + ".6.scale", c);
      }
    }
  t.add("htdemucs.freq_emb.embedding.weight", 512 * 48);
  wb("htdemucs.channel_upsampler", 512 * 384, 512);
  wb("htdemucs.channel_downsampler", 384 * 512, 384);
  wb("htdemucs.channel_upsampler_t", 512 * 384, 512);
  wb("htdemucs.channel_downsampler_t", 384 * 512, 384);
  std::string x = "htdemucs.crosstransformer";
  wb(x + ".norm_in", 512, 512);
  wb(x + ".norm_in_t", 512, 512);
  const char* stacks[2] = {"layers", "layers_t"};
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < 5; ++i) {
      std::string p = x + "." + stacks[s] + "." + std::to_string(i);
      std::string a = p + (i % 2 == 0 ? ".self_attn" : ".cross_attn");
      t.add(a + ".in_proj_weight", 1536 * 512);
      t.add(a + ".in_proj_bias", 1536);
      wb(a + ".out_proj", 512 * 512, 512);
      wb(p + ".linear1", 2048 * 512, 2048);
      wb(p + ".linear2", 512 * 2048, 512);
      wb(p + ".norm1", 512, 512);
      wb(p + ".norm2", 512, 512);
      if (i % 2 == 1) wb(p + ".norm3", 512, 512);
      wb(p + ".norm_out", 512, 512);
      t.add(p + ".gamma_1.scale", 512);
      t.add(p + ".gamma_2.scale", 512);
    }
  wb("text_attn.q_proj", 384 * 384, 384);
  wb("text_attn.k_proj", 384 * 512, 384);
  wb("text_attn.v_proj", 384 * 512, 384);
  t.add("text_attn.attn.in_proj_weight", 1152 * 384);
  t.add("text_attn.attn.in_proj_bias", 1152);
  wb("text_attn.attn.out_proj", 384 * 384, 384);
  wb("text_attn.out_mlp.0", 384 * 384, 384);
  wb("text_attn.out_mlp.2", 384 * 384, 384);
  wb("text_attn.norm_q", 384, 384);
  wb("text_attn.norm_out", 384, 384);
  const char* dec[2] = {"freq_decoder", "time_decoder"};
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < 4; ++i) {
      std::string p = std::string(dec[f]) + ".layers." + std::to_string(i);
      wb(p + ".0", (long)kDecCh[i] * kDecCh[i + 1] * 8, kDecCh[i + 1]);
      if (i < 3) wb(p + ".1", kDecCh[i + 1], kDecCh[i + 1]);
    }
  wb("freq_out", 8, 2);
  wb("time_out", 8, 2);
  return t;
}

// One entry of the packed blob: a GEMM-layout copy (activation dtype) or an fp32 auxiliary vector.
struct PackItem {
  std::string key;      // lookup key used by the forward
  std::string src;      // source parameter name
  long src_off;         // extra element offset inside the source tensor
  int kind;             // pack_weight_kernel kind (0 copy, 1 conv k-major, 2 convT phases, 3 GLU interleave, 4 replicate x4, ...);
                        // 8 / 9 / 10: derived from the weight AND its bias (pack_gram_kernel)
  int d0, d1, d2;
  long numel;
  bool is_f32;          // stored as fp32 (aux vectors) instead of the activation dtype
  long offset_bytes;    // filled by layout
};

inline int hidden_pad(int h) { return h <= 16 ? 16 : (h <= 32 ? 32 : (h + 15) / 16 * 16); }

inline std::vector<PackItem> build_pack_list() {
  std::vector<PackItem> v;
  auto add = [&](const std::string& key, const std::string& src, long so, int kind, int d0, int d1, int d2, long n, bool f32) {
    v.push_back({key, src, so, kind, d0, d1, d2, n, f32, 0});
  };
  const char* br[2] = {"encoder", "tencoder"};
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < 4; ++i) {
      int cin = i == 0 ? (f == 0 ? 4 : 2) : kCh[i - 1];
      int c = kCh[i];
      std::string p = std::string("htdemucs.") + br[f] + "." + std::to_string(i);
      add(p + ".conv.w", p + ".conv.weight", 0, 1, c, cin, 8, (long)c * cin * 8, false);
      add(p + ".rewrite.w", p + ".rewrite.weight", 0, 3, 2 * c, c, 0, (long)2 * c * c, false);
      add(p + ".rewrite.b", p + ".rewrite.bias", 0, 3, 2 * c, 1, 0, 2 * c, true);
      for (int d = 0; d < 2; ++d) {
        std::string q = p + ".dconv.layers." + std::to_string(d);
        add(q + ".0.w", q + ".0.weight", 0, 1, c / 8, c, 3, (long)(c / 8) * c * 3, false);
        add(q + ".3.w", q + ".3.weight", 0, 0, 0, 0, 0, (long)2 * c * (c / 8), false);
        // tensor-core DConv: hidden channels zero-padded to hp (multiple of 16), expand rows GLU-interleaved
        const int h = c / 8, hp = hidden_pad(h);
        add(q + ".0.wp", q + ".0.weight", 0, 5, h, c, 3, (long)hp * c * 3, false);
        add(q + ".0.bp", q + ".0.bias", 0, 6, h, 0, 0, hp, true);
        add(q + ".1.wp", q + ".1.weight", 0, 6, h, 0, 0, hp, true);
        add(q + ".1.bp", q + ".1.bias", 0, 6, h, 0, 0, hp, true);
        add(q + ".3.wp", q + ".3.weight", 0, 7, 2 * c, h, hp, (long)2 * c * hp, false);
        add(q + ".3.bi", q + ".3.bias", 0, 3, 2 * c, 1, 0, 2 * c, true);
        // GroupNorm-2 statistics of e = W2 g + b2 as a quadratic form of the narrow g (dconv_tile.cu pass B): G = W2^T W2 split into
        // two activation-dtype halves (hi + lo), and the fp32 vector [2 W2^T b2 (hp) | column sums of W2 (hp) | sum b2 | sum b2^2]
        add(q + ".3.ghi", q + ".3.weight", 0, 8, 2 * c, h, hp, (long)hp * hp, false);
        add(q + ".3.glo", q + ".3.weight", 0, 9, 2 * c, h, hp, (long)hp * hp, false);
        add(q + ".3.gv", q + ".3.weight", 0, 10, 2 * c, h, hp, (long)2 * hp + 2, true);
        add(q + ".4.wi", q + ".4.weight", 0, 3, 2 * c, 1, 0, 2 * c, true);
        add(q + ".4.bi", q + ".4.bias", 0, 3, 2 * c, 1, 0, 2 * c, true);
      }
    }
  const char* ud[4] = {"htdemucs.channel_upsampler", "htdemucs.channel_downsampler", "htdemucs.channel_upsampler_t",
                       "htdemucs.channel_downsampler_t"};
  for (int i = 0; i < 4; ++i) add(std::string(ud[i]) + ".w", std::string(ud[i]) + ".weight", 0, 0, 0, 0, 0, 512 * 384, false);
  std::string x = "htdemucs.crosstransformer";
  const char* stacks[2] = {"layers", "layers_t"};
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < 5; ++i) {
      std::string p = x + "." + stacks[s] + "." + std::to_string(i);
      std::string a = p + (i % 2 == 0 ? ".self_attn" : ".cross_attn");
      add(p + ".in_proj.w", a + ".in_proj_weight", 0, 0, 0, 0, 0, 1536 * 512, false);
      add(p + ".out_proj.w", a + ".out_proj.weight", 0, 0, 0, 0, 0, 512 * 512, false);
      add(p + ".linear1.w", p + ".linear1.weight", 0, 0, 0, 0, 0, 2048 * 512, false);
      add(p + ".linear2.w", p + ".linear2.weight", 0, 0, 0, 0, 0, 512 * 2048, false);
    }
  add("text_attn.out_mlp.0.w", "text_attn.out_mlp.0.weight", 0, 0, 0, 0, 0, 384 * 384, false);
  add("text_attn.out_mlp.2.w", "text_attn.out_mlp.2.weight", 0, 0, 0, 0, 0, 384 * 384, false);
  const char* dec[2] = {"freq_decoder", "time_decoder"};
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < 4; ++i) {
      std::string p = std::string(dec[f]) + ".layers." + std::to_string(i);
      int ci = kDecCh[i], co = kDecCh[i + 1];
      add(p + ".0.w", p + ".0.weight", 0, 2, ci, co, 0, (long)ci * co * 8, false);
      add(p + ".0.b4", p + ".0.bias", 0, 4, co, 0, 0, 4 * co, true);
    }
  return v;
}

}  // namespace athtd
