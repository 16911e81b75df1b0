/*
 * aldm_b200.h -- C-ABI of the B200-native AudioLDM2 sampling hot path.
 *
 * Drop-in boundary.  The reference (haoheliu/AudioLDM2) is pure Python/PyTorch;
 * its "FFI" for this path is the set of torch module calls listed below.  A Python host
 * (audioldm2_b200/engine.py, bound with ctypes -- see INTEGRATION.md) builds a flat table of
 * `aldm_op` records from the reference config dicts + state_dict and hands it to this library;
 * everything that touches the mel-latent tensor then runs as hand-written sm_100a kernels.
 *
 *   reference call (file:line)                                   replaced by
 *   -----------------------------------------------------------  ---------------------------------
 *   DiffusionWrapper.forward -> UNetModel.forward                aldm_program_run(unet program)
 *       latent_diffusion/models/ddpm.py:1821-1879,
 *       modules/diffusionmodules/openaimodel.py:837-885
 *   DDIMSampler.p_sample_ddim CFG combine + x_{t-1} update       aldm_ddim_step
 *       latent_diffusion/models/ddim.py:298-300,339-354
 *   masked blend + q_sample                                      aldm_ddim_step (mask != NULL)
 *       models/ddim.py:226-231, models/ddpm.py:430-436
 *   LatentDiffusion.decode_first_stage -> AutoencoderKL.decode   aldm_program_run(vae-decoder program)
 *       models/ddpm.py:922-926, latent_encoder/autoencoder.py:111-117,
 *       modules/diffusionmodules/model.py:653-686
 *   encode_first_stage -> Encoder.forward + quant_conv           aldm_program_run(vae-encoder program)
 *       models/ddpm.py:941-943, model.py:519-543, autoencoder.py:103-109
 *   first_stage_model.vocoder(mel)  (HiFi-GAN Generator.forward) aldm_program_run(vocoder program)
 *       models/ddpm.py:928-939, hifigan/models.py:149-165
 *   TacotronSTFT.mel_spectrogram                                 aldm_stft_mel
 *       utilities/audio/stft.py:159-178
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host_*; buffers are caller-owned (torch
 *     tensors), 16-byte aligned, and must stay alive while a program that references them exists;
 *   - every entry point returns 0 on success or a negative ALDM_E_* code; no exceptions, no
 *     abort; `aldm_last_error()` returns a thread-local message;
 *   - kernels are enqueued on the caller's stream and never synchronise; there is no CPU
 *     fallback: an unsupported shape is ALDM_E_UNSUPPORTED.
 *   - activations are channels-last: [B, H, W, C] fp32 ("f32 tensors") or, when they feed a
 *     tensor-core GEMM, "operand planes": two bf16 arrays hi,lo of shape [rows, Cp] with
 *     x ~= hi + lo (Cp = C rounded up to 8).  Weights are packed by the host into 128-byte
 *     swizzled tile images (audioldm2_b200/packing.py).
 */
#ifndef ALDM_B200_H_
#define ALDM_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALDM_ABI_VERSION 6
#define ALDM_MAX_TAPS 16

enum {
  ALDM_OK = 0,
  ALDM_E_ARG = -1,
  ALDM_E_SHAPE = -2,
  ALDM_E_ALIGN = -3,
  ALDM_E_CUDA = -4,
  ALDM_E_NOMEM = -5,
  ALDM_E_UNSUPPORTED = -6
};

/* ---- GEMM / implicit-GEMM convolution ---------------------------------------------------- */

enum { ALDM_GEMM_TC = 0, ALDM_GEMM_SIMT = 1, ALDM_GEMM_TC_V1 = 2 };   /* aldm_gemm_desc.impl: persistent tcgen05 | CUDA-core checker | one-tile-per-CTA tcgen05 */
/* OR-ed into aldm_gemm_desc.impl: w_packed is never written while the program runs (model weights), so the
 * kernel may start streaming it before its programmatic-dependency wait (overlapping the previous kernel's tail). */
#define ALDM_GEMM_STATIC_B (1 << 16)
enum { ALDM_ACT_NONE = 0, ALDM_ACT_GEGLU = 1, ALDM_ACT_TANH = 2, ALDM_ACT_SILU = 3 };
enum { ALDM_OUT_F32 = 0, ALDM_OUT_PLANES = 1, ALDM_OUT_NCHW = 2, ALDM_OUT_QKV = 3 };

/* out[row(m), n] = epilogue( sum_k A[m, k] * W[n, k] ),  k = tap * Cp + c
 * A is gathered from operand planes laid out [B, Hs, Ws, Cp]:
 *    m -> (b, oh, ow);  ih = oh*sy + dy[tap], iw = ow*sx + dx[tap]  (zero outside [0,H)x[0,W));
 *    source pixel = (ih >> up, iw >> up), Hs = H >> up, Ws = W >> up; source batch = b % bmod.
 * epilogue: v = acc + bias[n] + rowvec[b*ld_rowvec + n]; v = act(v); v += res[row(m)*ld_res + n];
 *           v *= alpha; if (accumulate) v += out_old;  store.
 * GEGLU: weights are packed so that tile columns [0,BN/2) are values and [BN/2,BN) their gates;
 *        the stored width is N/2.
 * Output row mapping: orow = (b*OHF + oh*osy + ooy)*OWF + ow ; element = out[orow*ldo + n]
 *        (ALDM_OUT_NCHW: out[((b*N + n)*OH + oh)*OW + ow]).
 * ALDM_OUT_F32 with out_hi/out_lo != NULL: dual output, the values are additionally stored as operand
 *        planes [orow, ldo] (saves the copy-prep kernel in front of the next GEMM).
 * ALDM_OUT_QKV (attention projections): columns n < n_split go to the planes out_hi/out_lo
 *        [row m, ldo]; columns n >= n_split (the V projection) are stored TRANSPOSED into
 *        out2_hi/out2_lo[((m / tok_per_batch) * (N - n_split) + (n - n_split)) * ld_t + m % tok_per_batch]
 *        so that the attention kernel finds V^T K-major (keys contiguous).  n_split % bn == 0. */
typedef struct aldm_gemm_desc {
  const void* a_hi;          /* bf16 [B_src, Hs, Ws, Cp] */
  const void* a_lo;
  const void* w_packed;      /* tile images, see packing.py */
  const float* w_plain;      /* optional fp32 [N, Kpad] (SIMT reference path) */
  const float* bias;         /* [N] or NULL */
  const float* rowvec;       /* [B, ld_rowvec] or NULL (timestep-embedding add) */
  const float* res;          /* residual, same row mapping as out, or NULL */
  float* out;                /* fp32 output (ALDM_OUT_F32 / NCHW) */
  void* out_hi;              /* operand-plane output (ALDM_OUT_PLANES), [rows, ldo] bf16 */
  void* out_lo;
  void* out2_hi;             /* ALDM_OUT_QKV: transposed planes of the columns >= n_split */
  void* out2_lo;
  float* ws;                 /* split-K workspace [splitk, Mpad, Npad] fp32 or NULL */
  int32_t B, H, W, Cp;       /* logical conv input (after nearest-upsample if up=1) */
  int32_t up, bmod;
  int32_t OH, OW, sy, sx;
  int32_t ntaps;
  int16_t dy[ALDM_MAX_TAPS];
  int16_t dx[ALDM_MAX_TAPS];
  int32_t N, K, Kpad, bn;    /* bn: N tile (32/64/128); Kpad multiple of 64 */
  int32_t ldo, ld_res, ld_rowvec;
  int32_t OHF, OWF, osy, ooy;
  int32_t act, out_mode, accumulate, splitk, impl;
  int32_t n_split, tok_per_batch, ld_t;
  float alpha;
} aldm_gemm_desc;

int aldm_gemm(const aldm_gemm_desc* d, void* stream);

/* ---- operand preparation (normalise / activate / split into bf16 hi+lo planes) ----------- */

enum { ALDM_PREP_COPY = 0, ALDM_PREP_SILU = 1, ALDM_PREP_LRELU = 2,
       ALDM_PREP_GN = 3, ALDM_PREP_GN_SILU = 4, ALDM_PREP_LN = 5 };

/* x = cat(src0[rows,c0], src1[rows,c1]) (fp32) -> f(x) -> planes hi/lo [rows, Cp].
 * GN: rows = B*HW, 32 groups over C = c0+c1, statistics per (b, group), eps as given
 *     (1e-5 UNet ResBlock/out, 1e-6 SpatialTransformer + VAE; SURVEY.md 8a').
 * LN: per-row statistics over C, eps 1e-5.  gamma/beta are [C]. */
typedef struct aldm_prep_desc {
  const float* src0; const float* src1;
  const float* gamma; const float* beta;
  void* out_hi; void* out_lo;
  double* scratch;           /* GN: B*(64*32*2*8 + 32*2*4 + 4) bytes: per-block double partials, (mean, rstd) floats, ticket
                                counters; must be ZERO before the first GroupNorm that uses it (the tickets reset themselves) */
  int32_t rows, c0, c1, Cp;
  int32_t B, HW, groups;
  int32_t mode;
  float eps, slope;
  int32_t src_nchw;          /* src0 is [B, C, HW] (NCHW) instead of [B*HW, C] */
} aldm_prep_desc;

int aldm_prep(const aldm_prep_desc* d, void* stream);

/* fp32 matrix -> weight tile images on the device (dynamic B operands: VAE attention K, V^T).
 * src is [N, K] with row stride lds (transpose=0) or [K, N] (transpose=1). */
int aldm_pack_b(const float* src, int32_t lds, int32_t transpose, int32_t N, int32_t K,
                int32_t bn, void* dst_packed, float* dst_plain, void* stream);

/* ---- attention --------------------------------------------------------------------------- */

/* softmax(scale * Q K^T + mask) V per (batch, head), head_dim = 32.
 * All operands are bf16 hi/lo planes written by the projection GEMMs (ALDM_OUT_QKV / ALDM_OUT_PLANES):
 *   Q : [B*Nq, ldq], head h at columns [q_col + h*32, +32)
 *   K : [Bkv*Nk, ldk], head h at columns [k_col + h*32, +32)
 *   Vt: [(bkv*heads*32 + h*32 + d), ld_t] keys contiguous (V transposed), ld_t >= Nk, ld_t % 8 == 0
 * kv batch index bkv = b % kv_bmod (0: bkv = b).  mask: [Bkv, Nk] floats (1 = keep) or NULL; entries != 1
 * are filled with -FLT_MAX before the softmax exactly like attention.py:356-360.
 * Output: operand planes [B*Nq, ldo], head h at columns [h*32, +32).
 * impl: ALDM_GEMM_TC = tcgen05 flash kernel, ALDM_GEMM_SIMT = CUDA-core checker. */
typedef struct aldm_attn_desc {
  const void* q_hi; const void* q_lo; const void* k_hi; const void* k_lo; const void* vt_hi; const void* vt_lo;
  const float* mask;
  void* out_hi; void* out_lo;
  int32_t B, heads, Nq, Nk, ldq, ldk, ld_t, ldo, q_col, k_col, kv_bmod, impl;
  float scale;
} aldm_attn_desc;

int aldm_attention(const aldm_attn_desc* d, void* stream);

/* row softmax in place: x[rows, n] (VAE AttnBlock, model.py:216-217), then split to planes */
int aldm_softmax_rows(const float* x, int32_t rows, int32_t n, float scale, void* out_hi, void* out_lo,
                      void* stream);

/* ---- small fused elementwise kernels ------------------------------------------------------ */

/* timestep_embedding (util.py:172-196): t[B] int64 -> planes [B, dim] (cos | sin);
 * freqs[dim/2] = exp(-ln(max_period)*i/(dim/2)) is tabulated by the host (util.py:183-187) */
int aldm_timestep_embedding(const int64_t* t, int32_t B, int32_t dim, const float* freqs,
                            void* out_hi, void* out_lo, void* stream);

/* K6: CFG combine + DDIM update (+ optional masked blend of the NEXT step's input is done by the
 * caller through `aldm_masked_blend`).  eps holds [2B, ...]: rows [0,B) uncond, [B,2B) cond.
 * x_prev = sqrt(a_prev)*pred_x0 + sqrt(1-a_prev-sigma^2)*e + sigma*noise, all fp32, n elements
 * per batch row; pred_x0 may be NULL. */
int aldm_ddim_step(const float* x, const float* eps_uncond, const float* eps_cond, const float* noise,
                   float* x_prev, float* pred_x0, int64_t n_total,
                   float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at, float guidance,
                   void* stream);

/* img = (sqrt_acp*x0 + sqrt_1m_acp*q_noise)*mask + (1-mask)*img; mask is [B,1,T,F] broadcast over C */
int aldm_masked_blend(float* img, const float* x0, const float* mask, const float* q_noise,
                      int32_t B, int32_t C, int32_t TF, float sqrt_acp, float sqrt_1m_acp, void* stream);

/* [B, C, HW] <-> [B, HW, C] fp32 */
int aldm_transpose_chw(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, int32_t to_nhwc,
                       void* stream);

/* out = scale * (mean + exp(0.5*clamp(logvar,-30,20)) * noise)  from NHWC moments [rows, 2*zc]
 * -> NCHW latent [B, zc, HW]   (distributions.py:24-41, ddpm.py:802) */
int aldm_posterior_sample(const float* moments, const float* noise_nchw, float* z_nchw,
                          int32_t B, int32_t zc, int32_t HW, float scale, void* stream);

/* ---- STFT + mel front end (K9) ------------------------------------------------------------ */

/* wav [B, T] fp32 in [-1,1] -> log-mel [B, frames, n_mels] (frames = T/hop + 1), reflect pad n_fft/2,
 * periodic Hann, radix-2 FFT (n_fft power of two <= 2048), magnitude, mel_basis [n_mels, n_fft/2+1]
 * GEMV, log(max(., 1e-5)).  (stft.py:52-81,159-178; audio_processing.py:85-91) */
int aldm_stft_mel(const float* wav, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                  const float* mel_basis, int32_t n_mels, float* out, int32_t out_frames, void* stream);

/* ---- programs: flat op tables replayed on a stream / as a CUDA graph ---------------------- */

enum { ALDM_OP_GEMM = 1, ALDM_OP_PREP = 2, ALDM_OP_ATTN = 3, ALDM_OP_SOFTMAX = 4, ALDM_OP_TEMB = 5,
       ALDM_OP_TRANSPOSE = 6, ALDM_OP_PACKB = 7, ALDM_OP_COPY = 8 };

typedef struct aldm_op {
  int32_t kind;
  int32_t tag;               /* free for the host (layer index, for profiling) */
  union {
    aldm_gemm_desc gemm;
    aldm_prep_desc prep;
    aldm_attn_desc attn;
    struct { const float* x; void* out_hi; void* out_lo; int32_t rows, n; float scale; } softmax;
    struct { const int64_t* t; const float* freqs; void* out_hi; void* out_lo; int32_t B, dim; } temb;
    struct { const float* src; float* dst; int32_t B, C, HW, to_nhwc; } transpose;
    struct { const float* src; void* dst_packed; float* dst_plain; int32_t lds, transpose, N, K, bn; } packb;
    struct { const void* src; void* dst; int64_t bytes; } copy;
  } u;
} aldm_op;

typedef struct aldm_program aldm_program;

int aldm_program_create(const aldm_op* ops, int32_t n_ops, aldm_program** out);
int aldm_program_run(aldm_program* p, void* stream);            /* plain launches */
int aldm_program_run_range(aldm_program* p, int32_t first, int32_t last, void* stream);
int aldm_program_capture(aldm_program* p, void* stream);        /* build + instantiate a CUDA graph */
int aldm_program_replay(aldm_program* p, void* stream);         /* cudaGraphLaunch */
int aldm_program_is_captured(aldm_program* p);                  /* 1 once aldm_program_capture succeeded */
int aldm_program_num_launches(aldm_program* p);                 /* kernels launched per run */
void aldm_program_destroy(aldm_program* p);

/* ---- engine: the reference's seams as single calls --------------------------
 * An engine ties the programs of one model instance to their fixed I/O slots (device addresses inside the
 * workspace the programs were resolved against) so that a caller who is not the Python host -- or the Python
 * host itself -- drives the hot path with the calls the reference makes:
 *   aldm_engine_set_conditioning  <- DiffusionWrapper.forward's cond-dict unpacking (ddpm.py:1821-1879), once per call
 *   aldm_engine_unet_eps          <- the two self.model.apply_model(x, t, c) calls of p_sample_ddim (ddim.py:293-296)
 *   aldm_engine_ddim_step         <- DDIMSampler.p_sample_ddim as a whole (ddim.py:265-355): UNet x2 + CFG + update
 *   aldm_engine_vae_decode        <- LatentDiffusion.decode_first_stage (ddpm.py:922-926)
 *   aldm_engine_vocoder           <- first_stage_model.vocoder(mel) in mel_spectrogram_to_waveform (ddpm.py:928-939)
 *   aldm_engine_vae_encode        <- encode_first_stage (ddpm.py:941-943), moments out
 * The engine borrows the programs (it never destroys them) and owns its descriptor copy plus the step graph
 * (all lanes as parallel branches) and the side streams / events used to capture it.  All
 * pointers are device pointers, fp32 contiguous NCHW as in the reference; everything is enqueued on `stream`;
 * nothing synchronises.  Single caller thread per engine. */
typedef struct aldm_engine aldm_engine;

#define ALDM_MAX_LANES 8

/* One UNet lane: an independent copy of the step / conditioning programs planned for B / n_lanes latent rows, with
 * its own workspace (the weight arena is shared).  Lanes are replayed as PARALLEL branches of one CUDA graph: the
 * UNet's deep levels are chains of ~10 us kernels that each fill a fraction of the 148 SMs, so independent
 * sub-batches overlap there while the large layers simply share the machine.  Samples are independent through the
 * whole path, so results do not depend on the lane count (up to split-K / tile-shape choices). */
typedef struct aldm_unet_lane {
  aldm_program* cond;           /* cross-attention K/V precompute (may be NULL: no cross-attention) */
  aldm_program* step;           /* one UNet evaluation of 2*Bl rows: rows [0,Bl) unconditional, [Bl,2Bl) conditional */
  float* x_slot;                /* [Bl, C, T, F] latent read by `step` */
  int64_t* t_slot;              /* [2Bl] DDPM timestep */
  float* eps_slot;              /* [2Bl, C, T, F] */
  float* ctx_slot[2];           /* [2Bl, ctx_len[i], ctx_dim[i]] zero-padded context i */
  float* mask_slot[2];          /* [2Bl, ctx_len[i]] 1 = attend */
  float* film_slot;             /* [2Bl, film_dim] or NULL */
} aldm_unet_lane;

typedef struct aldm_engine_desc {
  aldm_unet_lane lane[ALDM_MAX_LANES];   /* lane l owns latent rows [l*B/n_lanes, (l+1)*B/n_lanes) */
  int32_t n_lanes;              /* 1..ALDM_MAX_LANES, B % n_lanes == 0 */
  aldm_program* vae_dec;        /* may be NULL */
  aldm_program* vocoder;        /* may be NULL */
  aldm_program* vae_enc;        /* may be NULL */
  float* z_slot;                /* vae_dec input [B, C, T, F] */
  float* mel_slot;              /* vae_dec output [B, 1, T', F'] */
  float* voc_mel_slot;          /* vocoder input [B, T', F'] */
  float* wave_slot;             /* vocoder output [B, 1, L] */
  float* enc_mel_slot;          /* vae_enc input [B, 1, T', F'] */
  float* moments_slot;          /* vae_enc output [B, T, F, 2C] (channels-last) */
  int32_t B;                    /* latent batch the programs were planned for (all lanes together) */
  int32_t latent_elems;         /* C*T*F */
  int32_t mel_elems;            /* T'*F' */
  int32_t wave_len;             /* L */
  int32_t n_ctx;                /* 0..2 */
  int32_t ctx_len[2];
  int32_t ctx_dim[2];
  int32_t film_dim;
  int32_t use_graph;            /* 1: the lanes' step programs are captured into one graph on first use and replayed */
} aldm_engine_desc;

int aldm_engine_create(const aldm_engine_desc* d, aldm_engine** out);
void aldm_engine_destroy(aldm_engine* e);
/* which: 0 = unconditional half, 1 = conditional half.  ctx_i [B, len_i, ctx_dim[i]], mask_i [B, len_i] (fp32 0/1),
 * len_i <= ctx_len[i]; film_y [B, film_dim] or NULL.  Call for both halves, then aldm_engine_precompute once. */
int aldm_engine_set_conditioning(aldm_engine* e, int32_t which, const float* ctx0, const float* mask0, int32_t len0,
                                 const float* ctx1, const float* mask1, int32_t len1, const float* film_y, void* stream);
int aldm_engine_precompute(aldm_engine* e, void* stream);
int aldm_engine_unet_eps(aldm_engine* e, const float* x, int64_t t, float* eps_uncond, float* eps_cond, void* stream);
/* x_prev (and pred_x0 unless NULL) [B, C, T, F]; scalars as in aldm_ddim_step */
int aldm_engine_ddim_step(aldm_engine* e, const float* x, int64_t t, const float* noise, float a_t, float a_prev,
                          float sigma_t, float sqrt_one_minus_at, float guidance, float* x_prev, float* pred_x0,
                          void* stream);
int aldm_engine_vae_decode(aldm_engine* e, const float* z, float* mel, void* stream);
int aldm_engine_vocoder(aldm_engine* e, const float* mel, float* wave, void* stream);
int aldm_engine_vae_encode(aldm_engine* e, const float* mel, float* moments, void* stream);

/* ---- misc ---------------------------------------------------------------------------------- */

int aldm_abi_version(void);
size_t aldm_sizeof_op(void);
size_t aldm_sizeof_gemm_desc(void);
size_t aldm_sizeof_engine_desc(void);
size_t aldm_offsetof_gemm(int32_t field);     /* 0:B 1:ntaps 2:dy 3:N 4:ldo 5:act 6:alpha 7:n_split (layout self-check) */
const char* aldm_last_error(void);
int aldm_device_check(int32_t device);
int aldm_debug_timeline(long long* host_out, int32_t n);   /* profiling aid: per-stage clock64 stamps of CTA 0 (scripts/prof_ops.py --timeline) */        /* 0 if `device` is sm_100 and kernels can load */
int aldm_debug_umma_rate(int32_t N, int32_t mode, int32_t reps, long long* host_out, int32_t n_out);   /* profiling aid: cycles for `reps` tcgen05.mma 128 x N x 16 on each of n_out SMs (scripts/umma_rate.py) */
int aldm_debug_store_rate(int32_t n_cta, int32_t iters, int32_t mode, long long region_bytes, long long* host_out);   /* profiling aid: SM -> L2 store throughput, STG.128 (0) vs TMA bulk store (1) (scripts/store_rate.py) */

#ifdef __cplusplus
}
#endif
#endif /* ALDM_B200_H_ */
