"""NativeLatentDiffusion: the reference ``LatentDiffusion`` hot-path surface on the native engine.

Method names and argument meaning follow latent_diffusion/models/ddpm.py so the sampler and the
pipeline read like the reference:

    apply_model(x, t, cond)                  ddpm.py:1034-1042  (DiffusionWrapper.forward :1821-1879)
    decode_first_stage(z)                    ddpm.py:922-926
    mel_spectrogram_to_waveform(mel)         ddpm.py:928-939
    encode_first_stage(x) (+ posterior)      ddpm.py:941-943, 793-802
    q_sample / masked blend                  ddpm.py:430-436, ddim.py:226-231

plus ``p_sample_ddim`` (ddim.py:265-355) fused into one native step.  Conditioning is the dict the
reference's DiffusionWrapper unpacks: ``{"context_list": [...], "mask_list": [...], "y": ...}``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, arch, engine, plan
from .sampler import DDIMSampler, ddpm_tables


def split_state_dict(state_dict: Dict[str, torch.Tensor]):
    """Reference checkpoint keys -> (unet, vae, vocoder, scale_factor)."""
    un = {k[len("model.diffusion_model."):]: v for k, v in state_dict.items() if k.startswith("model.diffusion_model.")}
    voc = {k[len("first_stage_model.vocoder."):]: v for k, v in state_dict.items()
           if k.startswith("first_stage_model.vocoder.")}
    vae = {k[len("first_stage_model."):]: v for k, v in state_dict.items()
           if k.startswith("first_stage_model.") and not k.startswith("first_stage_model.vocoder.")
           and not k.startswith("first_stage_model.loss.")}
    sf = float(state_dict["scale_factor"]) if "scale_factor" in state_dict else 1.0
    return un, vae, voc, sf


class NativeLatentDiffusion:
    def __init__(self, cfg: dict, unet_sd, vae_sd, vocoder_sd, batch: int, device="cuda:0", scale_factor: float = 1.0,
                 ctx_max_len=(8, 128), impl: str = "tc", keep_plain: bool = False, use_graph: bool = True,
                 with_encoder: bool = False, arena_bcast=None, use_engine_abi: bool = False):
        """``batch`` is the latent batch B_l = batchsize * n_candidate_gen_per_text the programs are
        planned for.  ``arena_bcast(name, cpu_or_none, nbytes) -> device tensor`` lets parallel.py
        replace the H2D upload by an NCCL broadcast from rank 0."""
        self.cfg = cfg
        self.device = torch.device(device)
        self.batch = batch
        self.scale_factor = scale_factor
        self.use_graph = use_graph
        self.num_timesteps = cfg["timesteps"]
        self.parameterization = "eps"
        for k, v in ddpm_tables(cfg["linear_start"], cfg["linear_end"], cfg["timesteps"]).items():
            setattr(self, k, v)
        self.latent = tuple(cfg["latent"])
        C_, T, Fq = self.latent
        pk = dict(impl=impl, keep_plain=keep_plain)
        ds = 2 ** (len(cfg["vae"]["ch_mult"]) - 1)
        self.mel_hw = (T * ds, Fq * ds)

        def up(name, p: plan.Plan, ranges):
            dev_arena = arena_bcast(name, p.arena, p.arena.numel()) if arena_bcast else None
            return engine.DeviceProgram(p, self.device, ranges, dev_arena)

        pu = plan.build_unet(unet_sd, cfg["unet"], self.latent, batch, cfg_batched=True, ctx_max_len=ctx_max_len, **pk)
        self.unet = up("unet", pu, dict(cond=(pu.marks["cond_begin"], pu.marks["cond_end"]),
                                        step=(pu.marks["step_begin"], pu.marks["step_end"])))
        pd = plan.build_vae_decoder(vae_sd, cfg["vae"], self.latent, batch, scale_factor=scale_factor, **pk)
        self.vae_dec = up("vae_dec", pd, dict(all=(pd.marks["begin"], pd.marks["end"])))
        pv = plan.build_vocoder(vocoder_sd, cfg["vocoder"], self.mel_hw[0], batch, **pk)
        self.vocoder = up("vocoder", pv, dict(all=(pv.marks["begin"], pv.marks["end"])))
        self.vae_enc = None
        if with_encoder:
            pe = plan.build_vae_encoder(vae_sd, cfg["vae"], self.mel_hw, batch, **pk)
            self.vae_enc = up("vae_enc", pe, dict(all=(pe.marks["begin"], pe.marks["end"])))
        self.n_ctx = len([c for c in (cfg["unet"].get("context_dim") or []) if c is not None])
        self.film = cfg["unet"].get("extra_film_condition_dim") is not None
        self._t_host = torch.empty(2 * batch, dtype=torch.int64).pin_memory() if torch.cuda.is_available() else None
        # Engine-level C-ABI (include/aldm_b200.h, aldm_engine_*): the same programs and slots driven by one C call
        # per seam.  The Python orchestration below stays the default; tests/test_gpu_nets.py checks both agree.
        self.use_engine_abi = use_engine_abi
        self._engine = self._create_engine()

    def _create_engine(self):
        import ctypes as C
        L = _lib.lib()
        d = _lib.EngineDesc()
        u = self.unet
        d.unet_cond, d.unet_step = u.handles["cond"].value, u.handles["step"].value
        d.vae_dec, d.vocoder = self.vae_dec.handles["all"].value, self.vocoder.handles["all"].value
        d.vae_enc = self.vae_enc.handles["all"].value if self.vae_enc is not None else None
        d.x_slot, d.t_slot, d.eps_slot = u.view("x").data_ptr(), u.view("t").data_ptr(), u.view("eps").data_ptr()
        d.n_ctx = self.n_ctx
        for s in range(self.n_ctx):
            ctx = u.view(f"ctx{s}")
            d.ctx_slot[s], d.mask_slot[s] = ctx.data_ptr(), u.view(f"mask{s}").data_ptr()
            d.ctx_len[s], d.ctx_dim[s] = ctx.shape[1], ctx.shape[2]
        if self.film:
            y = u.view("y")
            d.film_slot, d.film_dim = y.data_ptr(), y.shape[1]
        d.z_slot, d.mel_slot = self.vae_dec.view("z").data_ptr(), self.vae_dec.view("mel").data_ptr()
        d.voc_mel_slot, d.wave_slot = self.vocoder.view("mel").data_ptr(), self.vocoder.view("wave").data_ptr()
        if self.vae_enc is not None:
            d.enc_mel_slot, d.moments_slot = self.vae_enc.view("mel").data_ptr(), self.vae_enc.view("moments").data_ptr()
        d.B = self.batch
        d.latent_elems = int(np.prod(self.latent))
        d.mel_elems = self.mel_hw[0] * self.mel_hw[1]
        d.wave_len = self.vocoder.view("wave").shape[-1]
        d.use_graph = int(self.use_graph)
        h = C.c_void_p()
        _lib.check(L.aldm_engine_create(C.byref(d), C.byref(h)), "engine_create")
        self._engine_desc = d
        return h

    def __del__(self):
        try:
            if getattr(self, "_engine", None):
                _lib.lib().aldm_engine_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    # conditioning (DiffusionWrapper.forward's dict unpacking, ddpm.py:1821-1879)
    # ------------------------------------------------------------------------------------------
    def set_conditioning(self, cond: dict, uncond: dict):
        """Rows [0,B) of every conditioning buffer hold the unconditional, [B,2B) the conditional branch.
        Cross-attention K/V of every layer are computed here once per call (step-invariant)."""
        B = self.batch
        if self.use_engine_abi:
            L, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
            keep = []
            for half, c in ((0, uncond), (1, cond)):
                a = [None, None, 0, None, None, 0]
                for s_ in range(self.n_ctx):
                    cl = c["context_list"][s_].to(self.device, torch.float32).contiguous()
                    ml = c["mask_list"][s_].to(self.device, torch.float32).contiguous()
                    keep += [cl, ml]
                    a[3 * s_], a[3 * s_ + 1], a[3 * s_ + 2] = cl.data_ptr(), ml.data_ptr(), cl.shape[1]
                y = c["y"].to(self.device, torch.float32).contiguous() if self.film else None
                keep.append(y)
                _lib.check(L.aldm_engine_set_conditioning(self._engine, half, *a, y.data_ptr() if y is not None else None, st),
                           "engine_set_conditioning")
            _lib.check(L.aldm_engine_precompute(self._engine, st), "engine_precompute")
            return
        for s in range(self.n_ctx):
            ctx, msk = self.unet.view(f"ctx{s}"), self.unet.view(f"mask{s}")
            ctx.zero_(); msk.zero_()
            for half, c in ((0, uncond), (1, cond)):
                cl, ml = c["context_list"][s], c["mask_list"][s]
                assert cl.shape[0] == B and cl.shape[1] <= ctx.shape[1], (cl.shape, ctx.shape)
                ctx[half * B:(half + 1) * B, :cl.shape[1]].copy_(cl, non_blocking=True)
                msk[half * B:(half + 1) * B, :ml.shape[1]].copy_(ml.float(), non_blocking=True)
        if self.film:
            y = self.unet.view("y")
            y[:B].copy_(uncond["y"], non_blocking=True); y[B:].copy_(cond["y"], non_blocking=True)
        self.unet.run("cond")

    def apply_model_pair(self, x: torch.Tensor, t: int):
        """Both apply_model calls of ddim.py:293-296 in one batched evaluation -> (eps_uncond, eps_cond)."""
        B = self.batch
        assert x.shape[0] == B and x.is_contiguous()
        if self.use_engine_abi:
            _lib.check(_lib.lib().aldm_engine_unet_eps(self._engine, x.data_ptr(), int(t), None, None,
                                                       torch.cuda.current_stream().cuda_stream), "engine_unet_eps")
            eps = self.unet.view("eps")
            return eps[:B], eps[B:]
        self.unet.view("x").copy_(x, non_blocking=True)
        self.unet.view("t").fill_(int(t))
        if self.use_graph:
            self.unet.replay("step")
        else:
            self.unet.run("step")
        eps = self.unet.view("eps")
        return eps[:B], eps[B:]

    def apply_model(self, x, t, cond: dict):
        """ddpm.py:1034-1042 for a single conditioning (runs the batched program with cond in both halves)."""
        self.set_conditioning(cond, cond)
        tv = int(t[0]) if torch.is_tensor(t) else int(t)
        return self.apply_model_pair(x, tv)[1].clone()

    def p_sample_ddim(self, x, st: dict, noise, guidance: float, out=None, pred_x0=None):
        """ddim.py:265-355: eps_u/eps_c, e = e_u + s (e_c - e_u), x_{t-1} update -- one graph replay + K6."""
        out = torch.empty_like(x) if out is None else out
        if self.use_engine_abi:
            assert x.is_contiguous() and noise.is_contiguous() and out.is_contiguous()
            _lib.check(_lib.lib().aldm_engine_ddim_step(
                self._engine, x.data_ptr(), int(st["t"]), noise.data_ptr(), st["a_t"], st["a_prev"], st["sigma_t"],
                st["sqrt_one_minus_at"], float(guidance), out.data_ptr(), pred_x0.data_ptr() if pred_x0 is not None else None,
                torch.cuda.current_stream().cuda_stream), "engine_ddim_step")
            return out
        e_u, e_c = self.apply_model_pair(x, st["t"])
        engine.ddim_step(x, e_u, e_c, noise, out, st["a_t"], st["a_prev"], st["sigma_t"], st["sqrt_one_minus_at"],
                         float(guidance), pred_x0)
        return out

    def masked_blend(self, img, x0, mask, q_noise, st: dict):
        return engine.masked_blend(img, x0, mask, q_noise, st["sqrt_acp_t"], st["sqrt_1m_acp_t"])

    # ------------------------------------------------------------------------------------------
    # first stage
    # ------------------------------------------------------------------------------------------
    def decode_first_stage(self, z: torch.Tensor) -> torch.Tensor:
        """ddpm.py:922-926 -> mel [B, 1, T', F'] (a view into the decoder workspace)."""
        if self.use_engine_abi:
            z = z.contiguous()
            _lib.check(_lib.lib().aldm_engine_vae_decode(self._engine, z.data_ptr(), None,
                                                         torch.cuda.current_stream().cuda_stream), "engine_vae_decode")
            return self.vae_dec.view("mel")
        self.vae_dec.view("z").copy_(z, non_blocking=True)
        self.vae_dec.run("all")
        return self.vae_dec.view("mel")

    def mel_spectrogram_to_waveform(self, mel: torch.Tensor) -> torch.Tensor:
        """ddpm.py:928-939 (without the .cpu().numpy()): mel [B,1,T,F] -> waveform [B,1,L] on the device."""
        B = mel.shape[0]
        if self.use_engine_abi:
            m = mel.reshape(B, mel.shape[-2], mel.shape[-1]).contiguous()
            _lib.check(_lib.lib().aldm_engine_vocoder(self._engine, m.data_ptr(), None,
                                                      torch.cuda.current_stream().cuda_stream), "engine_vocoder")
            return self.vocoder.view("wave")
        self.vocoder.view("mel").copy_(mel.reshape(B, mel.shape[-2], mel.shape[-1]), non_blocking=True)
        self.vocoder.run("all")
        return self.vocoder.view("wave")

    def encode_first_stage_moments(self, mel: torch.Tensor) -> torch.Tensor:
        if self.vae_enc is None:
            raise RuntimeError("engine was built without the VAE encoder (with_encoder=True)")
        self.vae_enc.view("mel").copy_(mel, non_blocking=True)
        self.vae_enc.run("all")
        return self.vae_enc.view("moments")

    def get_first_stage_encoding(self, moments: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """DiagonalGaussianDistribution.sample + scale (distributions.py:24-41, ddpm.py:793-802)."""
        return engine.posterior_sample(moments, noise.to(self.device).contiguous(), self.scale_factor)

    # ------------------------------------------------------------------------------------------
    # generate_batch (ddpm.py:1477-1570) minus the conditioning encoders and the CLAP re-ranker
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate_latent(self, cond: dict, uncond: dict, ddim_steps: int = 200, guidance: float = 3.5, eta: float = 1.0,
                        x_T=None, noise_fn=None, mask=None, x0=None):
        sampler = DDIMSampler(self)
        z, _ = sampler.sample(S=ddim_steps, batch_size=self.batch, shape=self.latent, conditioning=cond, eta=eta,
                              unconditional_guidance_scale=guidance, unconditional_conditioning=uncond, x_T=x_T,
                              noise_fn=noise_fn, mask=mask, x0=x0)
        return z

    @torch.no_grad()
    def generate_waveform(self, cond: dict, uncond: dict, ddim_steps: int = 200, guidance: float = 3.5, eta: float = 1.0,
                          x_T=None, noise_fn=None, mask=None, x0=None):
        z = self.generate_latent(cond, uncond, ddim_steps, guidance, eta, x_T, noise_fn, mask, x0)
        mel = self.decode_first_stage(z)
        return self.mel_spectrogram_to_waveform(mel)

    def launches_per_step(self) -> int:
        return self.unet.num_launches("step") + 1      # + K6

    def launches_decode(self) -> int:
        return self.vae_dec.num_launches("all") + self.vocoder.num_launches("all")


def build_synthetic(model_name: str = "audioldm2-full", batch: int = 1, device="cuda:0", cfg: Optional[dict] = None,
                    t5_len: int = 32, **kw) -> NativeLatentDiffusion:
    """Engine on the seeded synthetic checkpoint (no network: hub checkpoints are unreachable)."""
    from . import synth
    cfg = cfg or arch.model_config(model_name)
    lens = (8, t5_len) if len([c for c in cfg["unet"]["context_dim"] if c is not None]) > 1 else (t5_len,)
    return NativeLatentDiffusion(cfg, synth.unet_state_dict(cfg["unet"]), synth.vae_state_dict(cfg["vae"]),
                                 synth.vocoder_state_dict(cfg["vocoder"]), batch, device, ctx_max_len=lens, **kw)
