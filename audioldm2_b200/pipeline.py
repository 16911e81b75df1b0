"""Drop-in surface of ``audioldm2/pipeline.py`` for the sampling hot path.

Signatures follow the reference (pipeline.py:142-267).  What differs, by design of this tier:

* the conditioning encoders (CLAP / Flan-T5 / AudioMAE-GPT2) are out of scope and need hub
  downloads that are unreachable offline, so ``text_to_audio`` obtains the UNet-boundary
  conditioning from ``latent_diffusion.cond_provider(texts, batch) -> (cond, uncond)``; the default
  provider of ``build_model(synthetic=True)`` is the seeded synthetic one of SURVEY.md 8d;
* candidate re-ranking by CLAP similarity (ddpm.py:1554-1568) is not performed: with
  ``n_candidate_gen_per_text > 1`` the first candidate of each prompt is returned.
"""
from __future__ import annotations

import random
from typing import Callable, Optional

import numpy as np
import torch

from . import arch, engine, model, synth


def seed_everything(seed: int):
    """pipeline.py:20-31"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


def build_model(ckpt_path: Optional[str] = None, config=None, device=None, model_name: str = "audioldm2-full",
                batchsize: int = 1, n_candidate_gen_per_text: int = 1, synthetic: Optional[bool] = None,
                cond_provider: Optional[Callable] = None, t5_len: int = 32, **engine_kw):
    """pipeline.py:142-179.  ``ckpt_path`` is a reference ``<model_name>.pth`` (``["state_dict"]``,
    key layout of SURVEY.md 8b); without it (no network here) the seeded synthetic checkpoint is used."""
    if device is None or device == "auto":
        device = "cuda:0"
    cfg = arch.model_config(model_name)
    Bl = batchsize * n_candidate_gen_per_text
    n_cross = len([c for c in cfg["unet"]["context_dim"] if c is not None])
    lens = (8, 128) if n_cross > 1 else (128,)
    if ckpt_path is None:
        if synthetic is False:
            raise RuntimeError("no checkpoint given and hub download is unavailable offline (utils.py:209-219)")
        ld = model.build_synthetic(model_name, batch=Bl, device=device, t5_len=t5_len, **engine_kw)
    else:
        sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]          # pipeline.py:172
        un, vae, voc, sf = model.split_state_dict(sd)
        ld = model.NativeLatentDiffusion(cfg, un, vae, voc, Bl, device, scale_factor=sf, ctx_max_len=lens, **engine_kw)
    ld.model_name = model_name
    ld.cond_provider = cond_provider or (lambda texts, batch: synth.conditioning(cfg, batch, seed=77, t5_len=t5_len,
                                                                                 device=device))
    return ld


def text_to_audio(latent_diffusion, text, transcription="", seed=42, ddim_steps=200, duration=10, batchsize=1,
                  guidance_scale=3.5, n_candidate_gen_per_text=3, latent_t_per_second=25.6, config=None):
    """pipeline.py:181-211 -> np.ndarray [batchsize, 1, samples] float32 in (-1, 1)."""
    seed_everything(int(seed))
    Bl = batchsize * n_candidate_gen_per_text
    if Bl != latent_diffusion.batch:
        raise RuntimeError(f"engine was planned for latent batch {latent_diffusion.batch}, got {Bl} "
                           "(build_model(batchsize=..., n_candidate_gen_per_text=...))")
    assert int(duration * latent_t_per_second) == latent_diffusion.latent[1], "latent_t_size mismatch (pipeline.py:200)"
    texts = [text] * batchsize if isinstance(text, str) else list(text)
    cond, uncond = latent_diffusion.cond_provider(texts, Bl)
    wave = latent_diffusion.generate_waveform(cond, uncond, ddim_steps=ddim_steps, guidance=guidance_scale, eta=1.0)
    wave = wave.cpu().numpy()                                                  # ddpm.py:936
    if n_candidate_gen_per_text > 1:          # candidates of prompt i are rows i + k*batchsize (ddpm.py:1560-1562)
        wave = wave[:batchsize]
    return wave


def super_resolution_and_inpainting(latent_diffusion, text, original_audio_file_path=None, seed=42, ddim_steps=200,
                                    duration=None, batchsize=1, guidance_scale=2.5, n_candidate_gen_per_text=3,
                                    time_mask_ratio_start_and_end=(0.10, 0.15), freq_mask_ratio_start_and_end=(1.0, 1.0),
                                    config=None, waveform: Optional[torch.Tensor] = None, mel_basis: Optional[torch.Tensor] = None):
    """pipeline.py:213-267 with the native front end: waveform [B, T] (already resampled/normalised as
    read_wav_file does, tools.py:28-40) -> aldm_stft_mel -> VAE encoder -> masked DDIM -> waveform."""
    seed_everything(int(seed))
    ld = latent_diffusion
    Bl = batchsize * n_candidate_gen_per_text
    assert waveform is not None and mel_basis is not None, "pass the decoded waveform tensor and the mel filterbank"
    vc = ld.cfg["vocoder"]
    fb = engine.stft_mel(waveform.to(ld.device).contiguous(), vc["n_fft"], vc["hop_size"], mel_basis.to(ld.device),
                         out_frames=ld.mel_hw[0])                               # [B, T, F] == fbank (tools.py:86-104)
    mel = fb[:, None].expand(Bl // fb.shape[0] * fb.shape[0], 1, *fb.shape[1:]).contiguous()
    mom = ld.encode_first_stage_moments(mel)
    x0 = ld.get_first_stage_encoding(mom, torch.randn(Bl, ld.latent[0], ld.latent[1], ld.latent[2]))   # CPU randn (distributions.py:38)
    C_, T, F_ = ld.latent
    mask = torch.ones(Bl, 1, T, F_, device=ld.device)                          # ddpm.py:1611-1617
    mask[:, :, int(T * time_mask_ratio_start_and_end[0]):int(T * time_mask_ratio_start_and_end[1]), :] = 0
    mask[:, :, :, int(F_ * freq_mask_ratio_start_and_end[0]):int(F_ * freq_mask_ratio_start_and_end[1])] = 0
    texts = [text] * batchsize if isinstance(text, str) else list(text)
    cond, uncond = ld.cond_provider(texts, Bl)
    wave = ld.generate_waveform(cond, uncond, ddim_steps=ddim_steps, guidance=guidance_scale, eta=1.0, mask=mask, x0=x0)
    return wave.cpu().numpy()[:batchsize]
