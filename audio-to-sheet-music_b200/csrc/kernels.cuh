// Launcher declarations for every kernel of the path (definitions: elementwise.cu, fft.cu, ola.cu,
// gemm_simt.cu, gemm_tc.cu, attention.cu).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace athtd {

// ---- elementwise.cu
void launch_sum_sumsq(const float* x, int B, long n_per_sample, double* stats, cudaStream_t st);
void launch_finalize_gn(const double* stats, double count, float* mr, long n, int nslot, cudaStream_t st);
void launch_finalize_meanstd(const double* stats, double n, float* out, int B, cudaStream_t st);
void launch_set_meanstd(float* out, int B, float mean, float stdv, cudaStream_t st);
template <typename T> void launch_pack_wav(const float* wav, const float* meanstd, T* out, RowSpace rs, int L, cudaStream_t st);
template <typename T> void launch_pack_spec(const float* Z, const float* meanstd, T* out, RowSpace rs, int Tf, cudaStream_t st);
template <typename T> void launch_gn_gelu(T* h, RowSpace rs, int G2, int per_row, const float* mr, const float* w,
                                          const float* b, cudaStream_t st);
template <typename T> void launch_gn_glu_res(T* x, RowSpace xs, const T* e, RowSpace es, int G2, int per_row, const float* mr,
                                             const float* w, const float* b, const float* scale, cudaStream_t st);
template <typename T> void launch_norm_rows(const T* x, T* xout, T* y, long rows, int C, int S, const float* gmr,
                                            const float* gw, const float* gb, const float* lw, const float* lb,
                                            const float* pe, RowSpace yrs, cudaStream_t st, T* y2 = nullptr,
                                            const float* lw2 = nullptr, const float* lb2 = nullptr);
template <typename T> void launch_softmax_rows(T* s, long rows, int n, cudaStream_t st);
void launch_text_vectors(const float* emb, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                         const float* b3, float* out, int rows, cudaStream_t st);
template <typename T> void launch_add_rowvec(const T* x, T* y, long rows_per_b, int C, int B, const float* vec, long vstride,
                                             cudaStream_t st);
template <typename T> void launch_dec_apply(const T* u, int Uin, RowSpace us, int Cu, T* out, RowSpace os, int G2,
                                            int has_gn, const float* mr, const float* gw, const float* gb, const T* skip,
                                            RowSpace ss, cudaStream_t st);
template <typename T> void launch_pack_weight(const float* src, T* dst, long n, int kind, int d0, int d1, int d2, cudaStream_t st);
template <typename T> void launch_pack_gram(const float* w, const float* b, void* dst, int kind, int C2, int H, int HP, cudaStream_t st);

// ---- resample.cu (load_audio: polyphase sinc resampler + mono -> stereo)
int launch_resample(const float* x, int C_in, long T_in, const float* Kt, int o, int nw, int taps, int width, float* y, int C_out,
                    long T_out, cudaStream_t st);

// ---- dconv_row.cu (fused DConv residual branch of the frequency encoder layers)
template <typename T> bool dconv_row_supported(int C, int Tn);
template <typename T> void launch_dconv_row(T* y, RowSpace ys, const float* const* ptrs, cudaStream_t st);

// ---- enc_row.cu (bf16: fused conv(level 0) + DConv + rewrite of a frequency encoder layer, warp-level MMA, persistent)
// all weight pointers are entries of the packed blob (model.cuh build_pack_list): conv taps [C][32], conv3 [HP][3C],
// expand [2C][HP] and rewrite [2C][C] with GLU-interleaved rows, fp32 vectors zero-padded / interleaved accordingly
struct EncRowParams {
  const bf16* cw; const float* cb;
  const bf16* w1[2]; const float* b1[2]; const float* g1w[2]; const float* g1b[2];
  const bf16* w2[2]; const float* b2[2]; const float* g2w[2]; const float* g2b[2]; const float* scale[2];
  const bf16* rw; const float* rb; const float* emb; float emb_scale;
  // level 0 with fused conv: read the fp32 spectrogram Z [B, Tf, 2048, 4] directly and normalise on the fly
  // ((z - mean) / (1e-5 + std), ATHTDemucs_v2.py:268-270) instead of reading a packed bf16 copy; nullptr = read xin
  const float* zspec; const float* ms_spec; int zrows;
};
bool enc_row_supported(int C, int Tn, bool fuse_conv);
void launch_enc_row(const bf16* xin, RowSpace xis, const bf16* yin, bf16* out, RowSpace ys, const EncRowParams& P, bool fuse_conv,
                    cudaStream_t st);

// ---- dconv_tile.cu (bf16: DConv residual branch as three tiled mma.sync passes; time branch + frequency levels 3-4)
bool dconv_tile_supported(int C);
int launch_dconv_tile(bf16* y, RowSpace ys, bf16* h, int dil, const bf16* w1p, const float* b1p, const float* g1wp, const float* g1bp,
                      const bf16* w2p, const float* b2i, const float* g2wi, const float* g2bi, const bf16* ghi, const bf16* glo, const float* gv,
                      const float* scale, double* st1,
                      double* st2, const bf16* rw, const float* rb, bf16* out, cudaStream_t st);
bool dconv_tile_can_rewrite(int C, bool freq);

// ---- small_conv.cu (bf16: time-branch level-0 conv + GELU fused with the waveform normalisation)
void launch_tenc0_conv(const float* wav, const float* meanstd, int L, const bf16* w, const float* bias, bf16* y, RowSpace ys,
                       cudaStream_t st);

// last decoder layers (48 -> 4 channels, no norm): transposed conv + resize + skip in one kernel (bf16)
void launch_dec_last_freq(const bf16* x, RowSpace xs, const float* w, const float* bias, const bf16* skip, RowSpace ss, bf16* out,
                          RowSpace os, cudaStream_t st);
void launch_dec_last_time(const bf16* x, RowSpace xs, const bf16* w, const float* bias4, const bf16* skip, RowSpace ss, bf16* out,
                          RowSpace os, cudaStream_t st);

// ---- fft.cu
void launch_stft_cac(const float* wav, int B, int L, int Tf, float* Z, double* stats, const float2* tw, const float* win,
                     cudaStream_t st);
template <typename T> void launch_istft_fused(const float* Z, int Tf, int L, int Bout, int zb_div, const T* dec, RowSpace ds,
                                              int use_mask, const float* fo_w, const float* fo_b, const T* tdec, RowSpace ts,
                                              const float* to_w, const float* to_b, const float* meanstd_t, int ms_div,
                                              float* out, long out_bstride, const float2* tw, const float* win, cudaStream_t st);

// ---- attention.cu (bf16 tcgen05 flash attention, 8 heads x 64)
void flash_attn_set_poly(int npoly);
bool flash_attn_supported(long ldq, long ldkv, long ldo);
int launch_flash_attn(const bf16* q, long ldq, const bf16* k, const bf16* v, long ldkv, int B, int Sq, int Sk, bf16* o,
                      long ldo, cudaStream_t st);

// ---- ola.cu  (track-level chunk gather / weighted overlap-add, benchmark.py:155-204)
void launch_gather_chunks(const float* track, long T, int C, const long* starts, int n_chunks, int chunk_len, float* segs,
                          cudaStream_t st);
void launch_chunk_ola(const float* seg_out, long seg_stride, int k_base, int chunk_len, const long* starts,
                      const int* actual_len, const int* fade_len, const int* flags, int n_chunks, long stride,
                      const float* ramp_up, const float* ramp_down, const int* ramp_off, float* out, int C, long t_begin,
                      long t_end, int normalize, cudaStream_t st);

// ---- metrics.cu  (sums behind sdr_loss / sisdr_loss / new_sdr_metric, src/loss.py:9-87)
void launch_sdr_sums(const float* est, const float* tgt, int items, long n, double* sums, cudaStream_t st);

}  // namespace athtd

// ---- clap_text.cu (CLAP text tower: RoBERTa-base + pooler + projection; ATHTDemucs_v2.py:238-248)
#include "model.cuh"
namespace athtd {
ParamTable build_clap_param_table();
long clap_workspace_bytes(int P, int S);
int clap_text_forward(const ParamTable& pt, const float* params, const long* ids, const long* mask, int P, int S, void* workspace,
                      float* out, int normalize, cudaStream_t st);
}  // namespace athtd
