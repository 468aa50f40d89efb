/* athtd.h -- C ABI of the B200-native (sm_100a) AudioTextHTDemucs separation path.
 *
 * Drop-in boundary for the hot path of savage-hacker14/audio-to-sheet-music:
 *   AudioTextHTDemucs.forward            src/models/stem_separation/ATHTDemucs_v2.py:250-326
 *   OurModel._chunked_inference          benchmark.py:155-204   (== app.py:129-178)
 * The reference is pure Python/PyTorch; a reference-side binding is a ctypes (or torch-extension)
 * stub that passes tensor.data_ptr() values -- see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every pointer named *_dev is a CUDA device pointer
 * on the current device; `stream` is a cudaStream_t passed as void*; no function allocates device
 * memory, synchronises, or throws.  Return value 0 = ok, non-zero = error (text via
 * athtd_last_error()).  dtype: 0 = fp32 activations + fp32 SIMT GEMMs (parity build),
 * 1 = bf16 activations, fp32 accumulation (performance build).  There is no CPU fallback.
 */
#ifndef ATHTD_H
#define ATHTD_H
#ifdef __cplusplus
extern "C" {
#endif

const char* athtd_last_error(void);
int athtd_version(void);

/* ---- parameters: replaces nn.Module.state_dict() traffic (SURVEY.md Appendix E key layout).
 * The live tensors are copied, by name, into ONE flat fp32 device buffer at these offsets. */
int athtd_param_count(void);
const char* athtd_param_name(int i);
long athtd_param_numel(int i);
long athtd_param_offset(int i);      /* in floats */
long athtd_params_total(void);       /* in floats */

/* GEMM-layout copies (k-major conv taps, transposed-conv phases, GLU-interleaved rewrites) in the
 * activation dtype.  Replaces nothing in the reference (PyTorch re-lays-out weights inside cuDNN). */
long athtd_packed_bytes(int dtype);
int athtd_pack_weights(const float* params_dev, void* packed_dev, int dtype, void* stream);

/* ---- plan: one (B segments, L samples, P prompts per segment) shape */
long athtd_workspace_bytes(int B, int L, int P, int dtype);
/* consts: tw_dev float2[4096] = exp(-2*pi*i*k/4096); win_dev float[4096] periodic hann;
 * pe2d_dev float[8*Tf][512], pe1d_dev float[St][512] (demucs create_2d_sin_embedding /
 * create_sin_embedding, token order "(t1 fr)").  workspace_dev must be zero-filled once by the caller. */
void* athtd_plan_create(int B, int L, int P, int dtype, const float* params_dev, const void* packed_dev,
                        void* workspace_dev, long workspace_bytes, const float* tw_dev, const float* win_dev,
                        const float* pe2d_dev, const float* pe1d_dev);
void athtd_plan_destroy(void* plan);
int athtd_plan_tokens(void* plan, int* Tf, int* Sf, int* St);

/* AudioTextHTDemucs.forward (ATHTDemucs_v2.py:250-326) with the CLAP call (:282) replaced by the
 * embedding: wav [B,2,L] fp32, emb [B,P,512] fp32 -> out [B,P,2,L] fp32. */
int athtd_forward(void* plan, const float* wav_dev, const float* emb_dev, float* out_dev, void* stream);
/* prompt-independent half (_spec, normalise, _encode; ATHTDemucs_v2.py:261-279) ... */
int athtd_encode(void* plan, const float* wav_dev, void* stream);
/* AudioTextHTDemucs._encode(x, xt) (ATHTDemucs_v2.py:190-236) on its own: x_cac [B, Tf, 2048, 4] fp32 = the NORMALISED
 * complex-as-channels spectrogram (channels-last: the reference's x [B,4,2048,Tf] permuted), xt [B,2,L] the normalised
 * waveform.  Outputs are read through athtd_tap ("xenc", "xtenc", "enc{i}", "tenc{i}").  Not followed by athtd_decode. */
int athtd_encode_normalized(void* plan, const float* x_cac_dev, const float* xt_dev, void* stream);
/* ... and the per-prompt half (text_attn, decoders, mask, _ispec, time branch; :282-324). */
int athtd_decode(void* plan, const float* emb_dev, float* out_dev, void* stream);
int athtd_plan_launches(void* plan);   /* kernels launched by the last forward (or the last encode + the last decode) */
/* A plan created for B segments is a batch CAPACITY: any 1 <= b <= B runs in the same workspace (per-segment buffers are
 * laid out segment-major, pads stay in place).  Applies to the following forward / encode / decode calls; wav / emb / out
 * are then [b, ...].  Lets a track's tail batch (54 chunks = 32 + 22) reuse the batch-32 workspace. */
int athtd_plan_set_batch(void* plan, int B);
/* measurement aid (bench.py roofline pass): CUDA-event pairs around every GEMM launch of subsequent forwards;
 * get_profile synchronises on the last event and returns summed GEMM ms, algorithmic GFLOP (2*M*N*K) and launches. */
int athtd_plan_set_profile(void* plan, int on);
int athtd_plan_get_profile(void* plan, double* gemm_ms, double* gemm_gflop, int* gemm_launches);

/* debug / test access to intermediate buffers of the last forward.  Row-space buffers report
 * dims = {stored groups, stored rows per group, channels, front pad rows, G2, G2p, gpf, R} (common.cuh RowSpace). */
/* (bf16 build: buffers that the fused kernels bypass -- "xf0", "xt0", "yf0", "yf1" -- are not written and hold stale data) */
int athtd_tap(void* plan, const char* name, const void** ptr, long* numel, int* dtype, int dims[8]);
/* bf16 build only: route the supported GEMMs through the tcgen05 kernel (default on) or keep everything on the
 * CUDA-core kernel (A/B measurements, kernel-level parity tests). */
int athtd_plan_set_tc(void* plan, int on);
int athtd_plan_set_flash(void* plan, int on);   /* fused tcgen05 attention (default on) vs GEMM-softmax-GEMM */
int athtd_plan_set_fused_dconv(void* plan, int on);   /* per-row fused DConv kernel (default on) vs GEMM passes */
int athtd_plan_tc_launches(void* plan);
/* CUDA-graph replay of athtd_forward (default on): the second call with the same (batch, wav, emb, out) pointers captures the
 * launch sequence on a private stream, later calls replay it with one cudaGraphLaunch on `stream`.  Results are those of the
 * eager sequence (same kernels, same arguments).  graph_replays counts the forwards served from a graph. */
int athtd_plan_set_graph(void* plan, int on);
int athtd_plan_graph_replays(void* plan);

/* async device-to-device copy on `stream` (lets tests read taps without a second CUDA binding) */
int athtd_memcpy_d2d(void* dst_dev, const void* src_dev, long bytes, void* stream);

/* ---- spectral front / back end on their own (htdemucs._spec+_magnitude / _ispec, BASELINE config 2)
 * Z layout: [B, Tf, 2048, 4] fp32 with channels (L.re, L.im, R.re, R.im); stats: double[2*B] zeroed by caller. */
int athtd_stft_cac(const float* wav_dev, int B, int L, float* Z_dev, double* stats_dev, const float* tw_dev,
                   const float* win_dev, void* stream);
/* inverse of channels (L, R) of Z: frames_dev scratch float[B*2*Tf*4096], out [B,2,L] */
int athtd_istft(const float* Z_dev, int B, int L, float* frames_dev, float* out_dev, const float* tw_dev,
                const float* win_dev, void* stream);

/* ---- track-level chunk loop (benchmark.py:155-204) */
int athtd_gather_chunks(const float* track_dev, long T, int C, const long* starts_dev, int n_chunks, int chunk_len,
                        float* segs_dev, void* stream);
/* seg_out_dev: per-chunk model outputs [C, chunk_len] of global chunk k at seg_out + (k-k_base)*seg_stride.
 * flags bit0 = fade-in, bit1 = fade-out; ramps are torch.linspace(0,1,n) / (1,0,n) tables concatenated,
 * ramp_off[k] = table offset for chunk k.  Writes out[c, s - t_begin] for s in [t_begin, t_end), pitch t_end-t_begin. */
int athtd_chunk_ola(const float* seg_out_dev, long seg_stride, int k_base, int chunk_len, const long* starts_dev,
                    const int* actual_len_dev, const int* fade_len_dev, const int* flags_dev, int n_chunks, long stride,
                    const float* ramp_up_dev, const float* ramp_down_dev, const int* ramp_off_dev, float* out_dev, int C,
                    long t_begin, long t_end, void* stream);

/* ---- load_audio (app.py:113-126): torchaudio.transforms.Resample(sr, 44100) (sinc_interp_hann, width 6, rolloff 0.99) followed
 * by mono -> stereo repetition, on the device.  x [C_in, T_in] fp32; kernel_t_dev = the torchaudio filter bank
 * _get_sinc_resample_kernel(orig, new, gcd) TRANSPOSED to [taps][new] fp32 with orig / new already divided by their gcd and
 * taps = 2 * width + orig (NULL: same rate, channel copy only); y [C_out, T_out] with T_out = ceil(new * T_in / orig); output
 * channels beyond C_in repeat the last input channel.  File decoding stays with the caller (torchaudio.load). */
int athtd_load_audio(const float* x_dev, int C_in, long T_in, const float* kernel_t_dev, int orig, int new_, int taps, int width,
                     float* y_dev, int C_out, long T_out, void* stream);

/* ---- CLAP text tower (ATHTDemucs_v2.py:238-248: tokenizer output -> ClapTextModelWithProjection(...).text_embeds, or
 * ClapModel.get_text_features with normalize = 1).  Parameters: the "text_model.*" / "text_projection.*" tensors of the HF
 * state_dict copied by name into one flat fp32 device buffer at athtd_clap_param_offset(i) (like athtd_param_*).
 * input_ids / attention_mask: int64 [P, S] (the tokenizer's padded batch); out: fp32 [P, 512]; fp32 throughout. */
int athtd_clap_param_count(void);
const char* athtd_clap_param_name(int i);
long athtd_clap_param_numel(int i);
long athtd_clap_param_offset(int i);
long athtd_clap_params_total(void);
long athtd_clap_workspace_bytes(int P, int S);
int athtd_clap_text_forward(const float* params_dev, const long* input_ids_dev, const long* attention_mask_dev, int P, int S,
                            void* workspace_dev, float* out_dev, int normalize, void* stream);

/* test_inference.py:113-141 variant of the loop: chunks are multiplied by torchaudio-Fade masks (ramp tables supplied by the
 * caller: linspace(0,1,n) and -linspace(0,1,n)+1) and ADDED, never divided by a weight sum; same arguments as athtd_chunk_ola. */
int athtd_chunk_fade_add(const float* seg_out_dev, long seg_stride, int k_base, int chunk_len, const long* starts_dev,
                         const int* actual_len_dev, const int* fade_len_dev, const int* flags_dev, int n_chunks, long stride,
                         const float* ramp_up_dev, const float* ramp_down_dev, const int* ramp_off_dev, float* out_dev, int C,
                         long t_begin, long t_end, void* stream);

/* ---- evaluation metrics on the device (src/loss.py:9-87 sdr_loss / sisdr_loss / new_sdr_metric; benchmark.py:555-588).
 * est / tgt: [items, n] fp32 rows; sums_dev: double[items][6] = {sum t, sum e, sum t^2, sum e^2, sum e*t, sum (t-e)^2}
 * (zeroed by the call).  The dB values are closed forms of these sums (audio-to-sheet-music_b200/metrics.py). */
int athtd_sdr_sums(const float* est_dev, const float* tgt_dev, int items, long n, double* sums_dev, void* stream);

/* tuning hook: how many of every 16 exp2 pairs of the attention softmax are evaluated by the FMA-pipe polynomial instead of the
 * SFU (0, 4, 5, 6 or 8; every setting computes the same softmax to bf16 accuracy); | 0x100 selects the one-thread-per-row softmax instead
 * of the default two threads per query row; | 0x200 runs only the exact running-maximum pass (the default runs an optimistic
 * fixed-reference pass and repeats a CTA with the exact pass when a row sum leaves its safe window); | 0x400 forces that repeat
 * (test hook).  Every combination returns the same softmax.  Process-wide. */
int athtd_attention_set_poly(int npoly);
/* programmatic dependent launch of the path's kernels (default OFF: measured 3 % slower on this path, profiles/r02_summary.md): a kernel's CTAs may set up (barrier init, TMEM allocation,
 * weight staging) while its predecessor drains; every kernel waits for the predecessor before touching its data.  Process-wide. */
int athtd_set_pdl(int on);
/* tuning hook of the tcgen05 GEMM tile selection (tools/): low 16 bits = widest N tile (256), 0x10000 = one CTA per SM only,
 * 0x20000 = N in (128, 256] as two N/2-wide tiles, 0x40000 / 0x80000 = route eligible 256-wide tiles (all / only K >= 1536) through
 * the cta_group::2 CTA-pair kernel, 0x100000 = 192-wide tiles on the 16- instead of the 12-epilogue-warp variant,
 * 0x200000 = per-segment statistics finalised by a separate launch instead of the GEMM's last CTA, 0x400000 = no weight-stationary
 * B tiles (short-K layers stream the weight tile with every K block again).  Every setting computes
 * the same GEMM.  Process-wide. */
int athtd_set_tc_tuning(int flags);

/* kernel-level parity test of the fused attention: q [B*Sq,512], k/v [B*Sk,512] bf16 (8 heads x 64) -> o [B*Sq,512] */
int athtd_attention_test(const void* q_dev, const void* k_dev, const void* v_dev, void* o_dev, int B, int Sq, int Sk,
                         void* stream);

/* ---- generic GEMM entry used by the kernel-level parity tests: C[M,N] = A[M,K] * B[N,K]^T (+bias), row-major */
int athtd_gemm_test(const void* A_dev, const void* B_dev, const float* bias_dev, void* C_dev, int M, int N, int K,
                    int dtype, int use_tensor_cores, void* stream);

#ifdef __cplusplus
}
#endif
#endif
