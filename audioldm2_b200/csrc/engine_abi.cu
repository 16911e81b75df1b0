// Engine-level C-ABI: the reference's seams as single calls over borrowed programs and
// their fixed I/O slots.  Pure orchestration: async copies into / out of the slots, one graph replay (or
// eager run) and the K6 update; no kernels of its own except the timestep fill.
#include <string.h>

#include <new>

#include "common.cuh"

struct aldm_engine {
  aldm_engine_desc d;
};

namespace aldm {

__global__ void fill_i64_kernel(long long* p, int n, long long v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

static int copy_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0 || dst == src) return ALDM_OK;
  ALDM_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
  return ALDM_OK;
}

static int run_unet(aldm_engine* e, const float* x, int64_t t, cudaStream_t st) {
  const aldm_engine_desc& d = e->d;
  int rc = copy_async(d.x_slot, x, (size_t)d.B * d.latent_elems * sizeof(float), st);
  if (rc) return rc;
  fill_i64_kernel<<<(2 * d.B + 127) / 128, 128, 0, st>>>(reinterpret_cast<long long*>(d.t_slot), 2 * d.B, (long long)t);
  ALDM_CHECK_CUDA(cudaGetLastError());
  if (d.use_graph) {
    if (!aldm_program_is_captured(d.unet_step)) {
      // first use: one eager run (module load, attribute set-up), then capture on a private stream
      rc = aldm_program_run(d.unet_step, st);
      if (rc) return rc;
      ALDM_CHECK_CUDA(cudaStreamSynchronize(st));
      cudaStream_t cs;
      ALDM_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      rc = aldm_program_capture(d.unet_step, cs);
      cudaStreamDestroy(cs);
      return rc;            // the eager run above already produced eps for this call
    }
    return aldm_program_replay(d.unet_step, st);
  }
  return aldm_program_run(d.unet_step, st);
}

}  // namespace aldm

using namespace aldm;

extern "C" int aldm_engine_create(const aldm_engine_desc* d, aldm_engine** out) {
  ALDM_REQUIRE(d && out, ALDM_E_ARG, "engine_create: null argument");
  ALDM_REQUIRE(d->unet_step && d->x_slot && d->t_slot && d->eps_slot, ALDM_E_ARG, "engine_create: UNet program / slots missing");
  ALDM_REQUIRE(d->B > 0 && d->latent_elems > 0 && d->latent_elems % 4 == 0, ALDM_E_SHAPE, "engine_create: B=%d latent_elems=%d", d->B,
               d->latent_elems);
  ALDM_REQUIRE(d->n_ctx >= 0 && d->n_ctx <= 2, ALDM_E_SHAPE, "engine_create: n_ctx=%d", d->n_ctx);
  for (int i = 0; i < d->n_ctx; ++i)
    ALDM_REQUIRE(d->ctx_slot[i] && d->mask_slot[i] && d->ctx_len[i] > 0 && d->ctx_dim[i] > 0, ALDM_E_ARG,
                 "engine_create: context %d slots / sizes missing", i);
  ALDM_REQUIRE(!d->vae_dec || (d->z_slot && d->mel_slot && d->mel_elems > 0), ALDM_E_ARG, "engine_create: VAE decoder slots missing");
  ALDM_REQUIRE(!d->vocoder || (d->voc_mel_slot && d->wave_slot && d->wave_len > 0 && d->mel_elems > 0), ALDM_E_ARG,
               "engine_create: vocoder slots missing");
  ALDM_REQUIRE(!d->vae_enc || (d->enc_mel_slot && d->moments_slot), ALDM_E_ARG, "engine_create: VAE encoder slots missing");
  aldm_engine* e = new (std::nothrow) aldm_engine();
  ALDM_REQUIRE(e, ALDM_E_NOMEM, "engine_create: out of host memory");
  e->d = *d;
  *out = e;
  return ALDM_OK;
}

extern "C" void aldm_engine_destroy(aldm_engine* e) { delete e; }

extern "C" int aldm_engine_set_conditioning(aldm_engine* e, int32_t which, const float* ctx0, const float* mask0, int32_t len0,
                                            const float* ctx1, const float* mask1, int32_t len1, const float* film_y,
                                            void* stream) {
  ALDM_REQUIRE(e && (which == 0 || which == 1), ALDM_E_ARG, "set_conditioning: bad engine / half");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const aldm_engine_desc& d = e->d;
  const float* ctx[2] = {ctx0, ctx1};
  const float* msk[2] = {mask0, mask1};
  const int len[2] = {len0, len1};
  for (int i = 0; i < d.n_ctx; ++i) {
    ALDM_REQUIRE(ctx[i] && msk[i], ALDM_E_ARG, "set_conditioning: context %d missing", i);
    ALDM_REQUIRE(len[i] >= 1 && len[i] <= d.ctx_len[i], ALDM_E_SHAPE, "set_conditioning: context %d length %d > planned %d", i, len[i],
                 d.ctx_len[i]);
    float* cdst = d.ctx_slot[i] + (size_t)which * d.B * d.ctx_len[i] * d.ctx_dim[i];
    float* mdst = d.mask_slot[i] + (size_t)which * d.B * d.ctx_len[i];
    // zero-pad to the planned length (padded keys carry mask 0)
    ALDM_CHECK_CUDA(cudaMemsetAsync(cdst, 0, (size_t)d.B * d.ctx_len[i] * d.ctx_dim[i] * sizeof(float), st));
    ALDM_CHECK_CUDA(cudaMemsetAsync(mdst, 0, (size_t)d.B * d.ctx_len[i] * sizeof(float), st));
    ALDM_CHECK_CUDA(cudaMemcpy2DAsync(cdst, (size_t)d.ctx_len[i] * d.ctx_dim[i] * sizeof(float), ctx[i],
                                      (size_t)len[i] * d.ctx_dim[i] * sizeof(float), (size_t)len[i] * d.ctx_dim[i] * sizeof(float),
                                      d.B, cudaMemcpyDeviceToDevice, st));
    ALDM_CHECK_CUDA(cudaMemcpy2DAsync(mdst, (size_t)d.ctx_len[i] * sizeof(float), msk[i], (size_t)len[i] * sizeof(float),
                                      (size_t)len[i] * sizeof(float), d.B, cudaMemcpyDeviceToDevice, st));
  }
  if (d.film_slot) {
    ALDM_REQUIRE(film_y, ALDM_E_ARG, "set_conditioning: film_y missing");
    int rc = copy_async(d.film_slot + (size_t)which * d.B * d.film_dim, film_y, (size_t)d.B * d.film_dim * sizeof(float), st);
    if (rc) return rc;
  }
  return ALDM_OK;
}

extern "C" int aldm_engine_precompute(aldm_engine* e, void* stream) {
  ALDM_REQUIRE(e, ALDM_E_ARG, "precompute: null engine");
  if (!e->d.unet_cond) return ALDM_OK;
  return aldm_program_run(e->d.unet_cond, stream);
}

extern "C" int aldm_engine_unet_eps(aldm_engine* e, const float* x, int64_t t, float* eps_uncond, float* eps_cond, void* stream) {
  ALDM_REQUIRE(e && x, ALDM_E_ARG, "unet_eps: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = run_unet(e, x, t, st);
  if (rc) return rc;
  const size_t half = (size_t)e->d.B * e->d.latent_elems;
  if (eps_uncond && (rc = copy_async(eps_uncond, e->d.eps_slot, half * sizeof(float), st))) return rc;
  if (eps_cond && (rc = copy_async(eps_cond, e->d.eps_slot + half, half * sizeof(float), st))) return rc;
  return ALDM_OK;
}

extern "C" int aldm_engine_ddim_step(aldm_engine* e, const float* x, int64_t t, const float* noise, float a_t, float a_prev,
                                     float sigma_t, float sqrt_one_minus_at, float guidance, float* x_prev, float* pred_x0,
                                     void* stream) {
  ALDM_REQUIRE(e && x && noise && x_prev, ALDM_E_ARG, "ddim_step: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = run_unet(e, x, t, st);
  if (rc) return rc;
  const long long n = (long long)e->d.B * e->d.latent_elems;
  return aldm_ddim_step(x, e->d.eps_slot, e->d.eps_slot + n, noise, x_prev, pred_x0, n, a_t, a_prev, sigma_t, sqrt_one_minus_at,
                        guidance, stream);
}

extern "C" int aldm_engine_vae_decode(aldm_engine* e, const float* z, float* mel, void* stream) {
  ALDM_REQUIRE(e && z && e->d.vae_dec, ALDM_E_ARG, "vae_decode: engine has no decoder program / null z");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = copy_async(e->d.z_slot, z, (size_t)e->d.B * e->d.latent_elems * sizeof(float), st);
  if (rc) return rc;
  if ((rc = aldm_program_run(e->d.vae_dec, stream))) return rc;
  return mel ? copy_async(mel, e->d.mel_slot, (size_t)e->d.B * e->d.mel_elems * sizeof(float), st) : ALDM_OK;
}

extern "C" int aldm_engine_vocoder(aldm_engine* e, const float* mel, float* wave, void* stream) {
  ALDM_REQUIRE(e && mel && e->d.vocoder, ALDM_E_ARG, "vocoder: engine has no vocoder program / null mel");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = copy_async(e->d.voc_mel_slot, mel, (size_t)e->d.B * e->d.mel_elems * sizeof(float), st);
  if (rc) return rc;
  if ((rc = aldm_program_run(e->d.vocoder, stream))) return rc;
  return wave ? copy_async(wave, e->d.wave_slot, (size_t)e->d.B * e->d.wave_len * sizeof(float), st) : ALDM_OK;
}

extern "C" int aldm_engine_vae_encode(aldm_engine* e, const float* mel, float* moments, void* stream) {
  ALDM_REQUIRE(e && mel && e->d.vae_enc, ALDM_E_ARG, "vae_encode: engine has no encoder program / null mel");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = copy_async(e->d.enc_mel_slot, mel, (size_t)e->d.B * e->d.mel_elems * sizeof(float), st);
  if (rc) return rc;
  if ((rc = aldm_program_run(e->d.vae_enc, stream))) return rc;
  return moments ? copy_async(moments, e->d.moments_slot, (size_t)e->d.B * 2 * e->d.latent_elems * sizeof(float), st) : ALDM_OK;
}
