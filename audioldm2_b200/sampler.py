"""Host-side DDIM scheduler: mirrors ``DDIMSampler`` (latent_diffusion/models/ddim.py) -- the Python
loop, the schedule tables and the RNG draw order stay on the host exactly as in the reference;
each loop body (two UNet evaluations + CFG combine + x_{t-1} update) is one native call.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch


def ddpm_tables(linear_start: float = 0.0015, linear_end: float = 0.0195, timesteps: int = 1000) -> dict:
    """DDPM.register_schedule (ddpm.py:201-262), 'linear' beta schedule (util.py:23-29)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    ac = np.cumprod(1.0 - betas, axis=0)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f32(betas), alphas_cumprod=f32(ac), alphas_cumprod_prev=f32(np.append(1.0, ac[:-1])),
                sqrt_alphas_cumprod=f32(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=f32(np.sqrt(1.0 - ac)))


class DDIMSampler:
    """Same constructor / ``make_schedule`` / ``sample`` surface as the reference class
    (ddim.py:14-163); ``model`` is a ``NativeLatentDiffusion``."""

    def __init__(self, model, schedule: str = "linear", device=None, **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.device = device or model.device

    def make_schedule(self, ddim_num_steps: int, ddim_discretize: str = "uniform", ddim_eta: float = 0.0, verbose: bool = False):
        assert ddim_discretize == "uniform"
        n = self.ddpm_num_timesteps
        c = n // ddim_num_steps
        self.ddim_timesteps = np.asarray(list(range(0, n, c))) + 1                 # util.py:55-75
        ac = self.model.alphas_cumprod.clone().detach().to(torch.float32).cpu()      # ddim.py:47
        alphas = ac[self.ddim_timesteps]
        alphas_prev = np.asarray([ac[0]] + ac[self.ddim_timesteps[:-1]].tolist())   # util.py:78-81 (float64 ndarray)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sigmas = ddim_eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
                sqrt_1m = np.sqrt(1.0 - alphas)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sigmas, alphas, alphas_prev
        self.ddim_sqrt_one_minus_alphas = sqrt_1m
        # per-step fp32 scalars exactly as torch.full(...) casts them (ddim.py:330-335)
        f = lambda v: torch.full((1,), v).item()
        self.steps = []
        for i, step in enumerate(np.flip(self.ddim_timesteps)):
            idx = len(self.ddim_timesteps) - i - 1
            self.steps.append(dict(t=int(step), index=idx, a_t=f(alphas[idx]), a_prev=f(alphas_prev[idx]),
                                   sigma_t=f(sigmas[idx]), sqrt_one_minus_at=f(sqrt_1m[idx]),
                                   sqrt_acp_t=float(self.model.sqrt_alphas_cumprod[int(step)]),
                                   sqrt_1m_acp_t=float(self.model.sqrt_one_minus_alphas_cumprod[int(step)])))

    @torch.no_grad()
    def sample(self, S: int, batch_size: int, shape, conditioning=None, eta: float = 0.0, mask=None, x0=None,
               unconditional_guidance_scale: float = 1.0, unconditional_conditioning=None, x_T=None,
               noise_fn: Optional[Callable[[int, str], torch.Tensor]] = None, verbose: bool = False, **kwargs):
        """ddim.py:94-163.  ``noise_fn(i, kind)`` (kind in {"step", "q"}) lets tests inject recorded noise;
        by default noise is drawn with torch.randn on the device in the reference's order."""
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C_, T, F_ = shape
        size = (batch_size, C_, T, F_)
        samples = self.ddim_sampling(conditioning, size, x_T=x_T, mask=mask, x0=x0,
                                     unconditional_guidance_scale=unconditional_guidance_scale,
                                     unconditional_conditioning=unconditional_conditioning, noise_fn=noise_fn)
        return samples, None

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, mask=None, x0=None, unconditional_guidance_scale: float = 1.0,
                      unconditional_conditioning=None, noise_fn=None):
        """ddim.py:166-262 -- the hot loop."""
        m = self.model
        dev = m.device
        img = torch.randn(shape, device=dev) if x_T is None else x_T.to(dev).contiguous().clone()   # ddim.py:191
        # ddim.py:293-296: without an unconditional branch (or with scale 1.0) the model output is apply_model(x, t, c)
        # alone.  The engine always evaluates two halves, so the conditional branch is loaded into both: e_u == e_c
        # bit for bit (same kernels, same data) and e_u + s (e_c - e_u) == e_c exactly.
        single = unconditional_conditioning is None or unconditional_guidance_scale == 1.0
        m.set_conditioning(cond, None if single else unconditional_conditioning)
        nxt = torch.empty_like(img)
        for i, st in enumerate(self.steps):
            if mask is not None:                                                   # ddim.py:226-231
                qn = noise_fn(i, "q") if noise_fn else torch.randn_like(x0)        # q_sample noise (ddpm.py:431)
                m.masked_blend(img, x0, mask, qn, st)
            noise = noise_fn(i, "step") if noise_fn else torch.randn(shape, device=dev)   # ddim.py:351
            m.p_sample_ddim(img, st, noise, unconditional_guidance_scale, out=nxt)
            img, nxt = nxt, img
        return img
