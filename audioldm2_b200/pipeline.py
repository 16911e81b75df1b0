"""Drop-in surface of ``audioldm2/pipeline.py`` for the sampling hot path.

``build_model`` / ``text_to_audio`` / ``super_resolution_and_inpainting`` keep the reference's signatures, positional
order and defaults (pipeline.py:142-267); extra keyword-only arguments configure what the reference obtains from the
network.  What differs, by design of this tier:

* the conditioning encoders (CLAP / Flan-T5 / AudioMAE-GPT2) are out of scope and need hub downloads that are
  unreachable offline: the UNet-boundary conditioning comes from ``latent_diffusion.cond_provider`` (``.cond(batch)``
  for the prompts, ``.uncond(n)`` for the unconditional branch); the default provider of a synthetic build is the
  seeded one of SURVEY.md 8d.  A provider may return the reference's keyed cond-dict or the unpacked form
  (model.unpack_cond_dict);
* candidate re-ranking (ddpm.py:1554-1568) uses ``latent_diffusion.ranker(waveform [n,1,L], texts) -> similarity [n]``
  (the reference's ``clap.cos_similarity``); without one the first candidate of each prompt is returned and a
  warning says so;
* the engine is planned per (latent batch, latent length); ``NativeAudioLDM2.engine`` re-plans on demand and caches.
"""
from __future__ import annotations

import warnings
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from . import arch, engine, frontend, model, parallel, synth
from .utils import seed_everything


class SyntheticConditioning:
    """Seeded conditioning at the UNet boundary: per-prompt rows for the conditional branch, one shared
    row for the unconditional one (T5("") / zero AudioMAE tokens / zero FiLM vector)."""

    def __init__(self, cfg: dict, seed: int = 77, t5_len: int = 32, device="cpu"):
        self.cfg, self.seed, self.t5_len, self.device = cfg, seed, t5_len, device

    def cond(self, batch: dict) -> dict:
        n = len(batch["text"])
        return synth.conditioning(self.cfg, n, seed=self.seed, t5_len=self.t5_len, device=self.device)[0]

    def uncond(self, n: int) -> dict:
        return synth.conditioning(self.cfg, n, seed=self.seed, t5_len=self.t5_len, device=self.device)[1]


def _tile(c, n_gen: int):
    """The n_gen tiling of generate_batch (ddpm.py:1516-1525): torch.cat([t] * n_gen) on every tensor, so the candidates
    of prompt i sit at rows i + k * batchsize."""
    if n_gen == 1 or c is None:
        return c
    if torch.is_tensor(c):
        return torch.cat([c] * n_gen, dim=0)
    if isinstance(c, dict):
        return {k: _tile(v, n_gen) for k, v in c.items()}
    if isinstance(c, (list, tuple)):
        return [_tile(v, n_gen) for v in c]
    return c


def select_best(waveform: np.ndarray, similarity, batchsize: int):
    """ddpm.py:1554-1564: per prompt i take the candidate (rows i, i+B, i+2B, ...) with the highest similarity."""
    sim = torch.as_tensor(similarity).reshape(-1)
    best_index = []
    for i in range(batchsize):
        candidates = sim[i::batchsize]
        best_index.append(i + int(torch.argmax(candidates).item()) * batchsize)
    return waveform[best_index], best_index


class NativeAudioLDM2:
    """What ``build_model`` returns: the reference's ``LatentDiffusion`` seen from pipeline.py (``generate_batch``,
    ``generate_batch_masked``, ``latent_t_size``) over per-shape native engines."""

    def __init__(self, cfg: dict, unet_sd, vae_sd, vocoder_sd, device, scale_factor: float = 1.0, ctx_max_len=None,
                 cond_provider=None, ranker: Optional[Callable] = None, **engine_kw):
        self.cfg, self.device = cfg, torch.device(device)
        self._sd = (unet_sd, vae_sd, vocoder_sd)
        self.scale_factor = scale_factor
        self.ctx_max_len = ctx_max_len
        self.engine_kw = engine_kw
        self.cond_provider, self.ranker = cond_provider, ranker
        self.latent_t_size = cfg["latent"][1]                 # pipeline.py:200 overwrites it per call
        self.cond_stage_key = "text"
        self._engines: Dict[Tuple[int, int, bool], model.NativeLatentDiffusion] = {}
        self._pinned: Dict[Tuple[int, ...], torch.Tensor] = {}

    # ---- engines -------------------------------------------------------------------------------------
    def engine(self, Bl: int, latent_t: Optional[int] = None, with_encoder: bool = False) -> model.NativeLatentDiffusion:
        T = int(latent_t or self.latent_t_size)
        for (b, t, enc), e in self._engines.items():
            if b == Bl and t == T and (enc or not with_encoder):
                return e
        if len(self._engines) >= 2:                          # keep the two most recent plans (1-3 GB of HBM each)
            self._engines.pop(next(iter(self._engines)))
            torch.cuda.empty_cache()
        cfg = dict(self.cfg)
        C_, _, F_ = self.cfg["latent"]
        assert T % 8 == 0, f"latent length {T} must be a multiple of 8 (three stride-2 levels)"
        cfg["latent"] = (C_, T, F_)
        n_cross = len([c for c in cfg["unet"]["context_dim"] if c is not None])
        lens = self.ctx_max_len or ((8, 128) if n_cross > 1 else (128,))
        e = model.NativeLatentDiffusion(cfg, *self._sd, Bl, self.device, scale_factor=self.scale_factor, ctx_max_len=lens,
                                        with_encoder=with_encoder, **self.engine_kw)
        self._engines[(Bl, T, with_encoder)] = e
        return e

    def _egress(self, wave: torch.Tensor) -> np.ndarray:
        """Waveform to host memory (ddpm.py:936 does a blocking ``.cpu().numpy()``): asynchronous copy into a cached
        pinned buffer on the current stream, one event wait, no device-wide synchronisation."""
        key = tuple(wave.shape)
        buf = self._pinned.get(key)
        if buf is None:
            buf = self._pinned[key] = torch.empty(wave.shape, dtype=torch.float32).pin_memory()
        buf.copy_(wave, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()
        return buf.numpy().copy()

    # ---- generate_batch (ddpm.py:1477-1570) -------------------------------------------------------------
    @torch.no_grad()
    def generate_batch(self, batch, ddim_steps=200, ddim_eta=1.0, x_T=None, n_gen=1, unconditional_guidance_scale=1.0,
                       unconditional_conditioning=None, use_plms=False, **kwargs):
        return self._generate(batch, ddim_steps, ddim_eta, x_T, n_gen, unconditional_guidance_scale,
                              unconditional_conditioning, use_plms, None, None)

    @torch.no_grad()
    def generate_batch_masked(self, batch, ddim_steps=200, ddim_eta=1.0, x_T=None, n_gen=1, unconditional_guidance_scale=1.0,
                              unconditional_conditioning=None, use_plms=False, time_mask_ratio_start_and_end=(0.25, 0.75),
                              freq_mask_ratio_start_and_end=(0.75, 1.0), **kwargs):
        return self._generate(batch, ddim_steps, ddim_eta, x_T, n_gen, unconditional_guidance_scale,
                              unconditional_conditioning, use_plms, time_mask_ratio_start_and_end, freq_mask_ratio_start_and_end)

    def _generate(self, batch, ddim_steps, ddim_eta, x_T, n_gen, guidance, uncond, use_plms, tmask, fmask):
        assert x_T is None and not use_plms, "the native path implements the DDIM sampler (pipeline.py never asks for PLMS)"
        shard = parallel.current_shard(len(batch["text"]))
        if shard is not None:              # one process per GPU: this rank generates prompts [lo, hi) of the call
            return self._generate_sharded(shard, batch, ddim_steps, ddim_eta, n_gen, guidance, uncond, tmask, fmask)
        return self._generate_local(batch, ddim_steps, ddim_eta, n_gen, guidance, uncond, tmask, fmask, None)

    def _generate_sharded(self, shard, batch, ddim_steps, ddim_eta, n_gen, guidance, uncond, tmask, fmask):
        """Every rank runs the same call with the same seed; prompts are cut into contiguous shards, the noise of the
        single-process latent batch (rows i + k * B) is drawn in full on every rank and sliced, and the selected waveforms are
        all-gathered at the end.  Results equal the single-process call row for row."""
        _, _, lo, hi = shard
        B = len(batch["text"])
        local = {k: (v[lo:hi] if (torch.is_tensor(v) or isinstance(v, list)) and len(v) == B else v) for k, v in batch.items()}
        rows = [i + k * B for k in range(n_gen) for i in range(lo, hi)]
        cond_l = parallel.shard_rows(self.cond_provider.cond(batch), lo, hi)       # conditioning of the whole call, this rank's rows
        out = self._generate_local(local, ddim_steps, ddim_eta, n_gen, guidance, uncond, tmask, fmask, (B, rows), cond_l)
        full = parallel.all_gather_rows(torch.from_numpy(out).to(self.device), B)
        return self._egress(full)

    def _generate_local(self, batch, ddim_steps, ddim_eta, n_gen, guidance, uncond, tmask, fmask, glob, cond_rows=None):
        masked = tmask is not None
        B = len(batch["text"])
        Bl = B * n_gen
        eng = self.engine(Bl, with_encoder=masked)
        C_, T, F_ = eng.latent
        # get_input -> encode_first_stage -> posterior.sample() (ddpm.py:845-846): a CPU torch.randn of the latent shape
        # (distributions.py:38).  Plain text_to_audio encodes an all-zero fbank only to read z.shape[0]; that encoder pass is
        # skipped here (345 GFLOP per prompt), the CPU draw is kept so every later CPU draw sees the reference's RNG state.
        x_T = noise_fn = None
        if glob is None:
            post_noise = torch.randn(B, C_, T, F_)
        else:                                 # sharded call: draw the global tensors, keep this rank's rows
            Bg, rows = glob
            post_noise = torch.randn(Bg, C_, T, F_)[rows[:B]]
            noise_fn = parallel.ShardedNoise(Bg * n_gen, 0, 0, (C_, T, F_), self.device, rows=rows,
                                             generator=torch.cuda.default_generators[self.device.index or 0])
            x_T = noise_fn.x_T()
        mask = x0 = None
        if masked:
            fbank = torch.as_tensor(batch["log_mel_spec"], dtype=torch.float32)           # [B, T', F']
            mel = _tile(fbank[:, None], n_gen).to(self.device)                            # torch.cat([z] * n_gen) (ddpm.py:1651)
            mom = eng.encode_first_stage_moments(mel)
            x0 = eng.get_first_stage_encoding(mom, _tile(post_noise, n_gen))
            mask = torch.ones(Bl, T, F_, device=self.device)                              # ddpm.py:1611-1617
            mask[:, int(T * tmask[0]):int(T * tmask[1]), :] = 0
            mask[:, :, int(F_ * fmask[0]):int(F_ * fmask[1])] = 0
            mask = mask[:, None].contiguous()
        cond = _tile(cond_rows if cond_rows is not None else self.cond_provider.cond(batch), n_gen)
        if guidance != 1.0 and uncond is None:
            uncond = self.cond_provider.uncond(Bl)                                        # ddpm.py:1529-1536
        texts = list(batch["text"]) * n_gen
        wave = eng.generate_waveform(cond, uncond, ddim_steps=ddim_steps, guidance=guidance, eta=ddim_eta, mask=mask, x0=x0,
                                     x_T=x_T, noise_fn=noise_fn)
        waveform = self._egress(wave)                                                     # ddpm.py:936
        if n_gen > 1:
            if self.ranker is None:
                warnings.warn("no ranker (the reference's clap.cos_similarity, ddpm.py:1556) is attached to this model: "
                              "returning the first of the n_candidate_gen_per_text candidates of every prompt")
                return waveform[:B]
            similarity = self.ranker(torch.from_numpy(waveform).squeeze(1), texts)
            waveform, _ = select_best(waveform, similarity, B)
        return waveform


def build_model(ckpt_path=None, config=None, device=None, model_name="audioldm2-full", *, synthetic: Optional[bool] = None,
                cond_provider=None, ranker: Optional[Callable] = None, t5_len: int = 32, ctx_max_len=None, **engine_kw):
    """pipeline.py:142-179.  ``ckpt_path`` is a reference ``<model_name>.pth`` (``["state_dict"]``, key layout of SURVEY.md
    8b); without it (no network here, utils.py:209-219) the seeded synthetic checkpoint is used.  Engines are planned lazily
    for the latent batch of each call (``batchsize * n_candidate_gen_per_text``)."""
    if device is None or device == "auto":
        device = torch.device("cuda:0")          # the native path has no CPU / MPS fallback
    cfg = arch.model_config(model_name) if config is None else config
    if isinstance(cfg, str):
        raise NotImplementedError("YAML configs are read by the reference's conditioning stack, which is out of scope; "
                                  "pass a dict from arch.model_config")
    if ckpt_path is None:
        if synthetic is False:
            raise RuntimeError("no checkpoint given and hub download is unavailable offline (utils.py:209-219)")
        un, vae, voc, sf = synth.unet_state_dict(cfg["unet"]), synth.vae_state_dict(cfg["vae"]), \
            synth.vocoder_state_dict(cfg["vocoder"]), 1.0
        if ctx_max_len is None:
            n_cross = len([c for c in cfg["unet"]["context_dim"] if c is not None])
            ctx_max_len = (8, t5_len) if n_cross > 1 else (t5_len,)
    else:
        sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]                      # pipeline.py:172
        un, vae, voc, sf = model.split_state_dict(sd)
    ld = NativeAudioLDM2(cfg, un, vae, voc, device, scale_factor=sf, ctx_max_len=ctx_max_len,
                         cond_provider=cond_provider or SyntheticConditioning(cfg, t5_len=t5_len, device=device), ranker=ranker,
                         **engine_kw)
    ld.model_name = model_name
    return ld


def make_batch_for_text_to_audio(text, transcription="", waveform=None, fbank=None, batchsize=1):
    """pipeline.py:84-124 restricted to the keys the hot path reads (the phoneme / kaldi-fbank entries feed the
    conditioning encoders).  ``text`` may also be a list of ``batchsize`` prompts."""
    text = [text] * batchsize if isinstance(text, str) else list(text)
    assert len(text) == batchsize, "a prompt list must have batchsize entries"
    if fbank is None:
        fbank = torch.zeros((batchsize, 1024, 64))          # not used (pipeline.py:93-96)
    else:
        fbank = torch.as_tensor(fbank, dtype=torch.float32)
        fbank = fbank.expand(batchsize, *fbank.shape[1:])
    if waveform is None:
        waveform = torch.zeros((batchsize, 160000))         # not used
    else:
        waveform = torch.as_tensor(waveform, dtype=torch.float32).expand(batchsize, -1)
    return {"text": text, "fname": [t.replace(" ", "_").replace("'", "_").replace('"', "_") for t in text],
            "waveform": waveform, "log_mel_spec": fbank, "transcription": [transcription] * batchsize}


def text_to_audio(latent_diffusion, text, transcription="", seed=42, ddim_steps=200, duration=10, batchsize=1,
                  guidance_scale=3.5, n_candidate_gen_per_text=3, latent_t_per_second=25.6, config=None):
    """pipeline.py:181-211 -> np.ndarray [batchsize, 1, samples] float32 in (-1, 1)."""
    seed_everything(int(seed))
    batch = make_batch_for_text_to_audio(text, transcription=transcription, waveform=None, batchsize=batchsize)
    latent_diffusion.latent_t_size = int(duration * latent_t_per_second)
    with torch.no_grad():
        waveform = latent_diffusion.generate_batch(batch, unconditional_guidance_scale=guidance_scale, ddim_steps=ddim_steps,
                                                   n_gen=n_candidate_gen_per_text, duration=duration)
    return waveform


def wav_to_fbank(latent_diffusion, original_audio_file_path=None, target_length=1024, waveform=None, sr=None):
    """tools.py:86-104 with the native front end: file / array -> read_wav_file's normalisation (tools.py:28-40) ->
    clip(-1, 1) -> aldm_stft_mel (K9) -> fbank [target_length, n_mels] (cropped like _pad_spec, tools.py:71-84)."""
    cfg = latent_diffusion.cfg
    vc = cfg["vocoder"]
    if waveform is None:
        waveform, sr = frontend.read_wav(original_audio_file_path)
    x = frontend.prepare_waveform(np.asarray(waveform, dtype=np.float32).reshape(-1), sr or vc["sampling_rate"],
                                  vc["sampling_rate"], target_length * vc["hop_size"])
    dev = latent_diffusion.device
    wav = torch.clip(torch.from_numpy(x), -1, 1).to(dev).contiguous()                      # get_mel_from_wav (tools.py:43-46)
    fb = engine.stft_mel(wav, vc["n_fft"], vc["hop_size"], frontend.mel_basis_for(cfg).to(dev), out_frames=target_length)[0]
    if fb.shape[-1] % 2 != 0:
        fb = fb[..., :-1]
    return fb, x


def super_resolution_and_inpainting(latent_diffusion, text, transcription="", original_audio_file_path=None, seed=42,
                                    ddim_steps=200, duration=None, batchsize=1, guidance_scale=2.5, n_candidate_gen_per_text=3,
                                    time_mask_ratio_start_and_end=(0.40, 0.6), freq_mask_ratio_start_and_end=(1.0, 1.0),
                                    latent_t_per_second=25.6, config=None, *, waveform=None, waveform_sr=None):
    """pipeline.py:213-267: STFT/mel front end (K9) -> VAE encoder -> masked DDIM -> decode -> vocoder.  ``waveform`` (+
    ``waveform_sr``) replaces the file read for callers that already hold the samples."""
    seed_everything(int(seed))
    if duration is None:
        duration = latent_diffusion.cfg["latent"][1] / latent_diffusion.cfg["latent_t_per_second"] if waveform is None \
            else len(np.reshape(waveform, -1)) / float(waveform_sr or latent_diffusion.cfg["sampling_rate"])
    frames_per_s = latent_diffusion.cfg["sampling_rate"] / latent_diffusion.cfg["vocoder"]["hop_size"]     # 102.4 at 16 kHz
    mel, _ = wav_to_fbank(latent_diffusion, original_audio_file_path, target_length=int(duration * frames_per_s),
                          waveform=waveform, sr=waveform_sr)
    batch = make_batch_for_text_to_audio(text, transcription=transcription, fbank=mel[None, ...], batchsize=batchsize)
    ds = 2 ** (len(latent_diffusion.cfg["vae"]["ch_mult"]) - 1)
    latent_diffusion.latent_t_size = mel.shape[0] // ds
    with torch.no_grad():
        waveform_out = latent_diffusion.generate_batch_masked(
            batch, unconditional_guidance_scale=guidance_scale, ddim_steps=ddim_steps, n_gen=n_candidate_gen_per_text,
            duration=duration, time_mask_ratio_start_and_end=time_mask_ratio_start_and_end,
            freq_mask_ratio_start_and_end=freq_mask_ratio_start_and_end)
    return waveform_out
