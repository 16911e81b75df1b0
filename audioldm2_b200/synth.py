"""Seeded synthetic checkpoints with the reference ``state_dict`` layout.

There is no network on the build or GPU boxes, so hub checkpoints cannot be fetched
(utils.py:209-219).  Parity and throughput are therefore measured on deterministic
random weights that have exactly the key names / shapes of the reference modules
(SURVEY.md 8b/8d).  Rules (SURVEY.md 8d "synthetic inputs"):

* conv / linear weights ~ N(0, gain/sqrt(fan_in)) -- *including* the ones the reference
  zero-initialises with ``zero_module`` (openaimodel.py:255-257,810; attention.py:452-454),
  otherwise half of the graph would be multiplied by zero and left untested;
* biases and norm beta ~ 0.02 N(0,1); norm gamma ~ 1 + 0.1 N(0,1).

Tensors are generated one by one from a CPU ``torch.Generator`` in sorted-key order, so
the same (config, seed) gives bit-identical weights on every machine with this torch.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import arch


def _fill(shapes: Dict[str, Tuple[int, ...]], seed: int, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif len(shp) == 1:                      # norm gamma
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) * (gain / math.sqrt(fan_in))
        out[name] = t.contiguous()
    return out


def unet_state_dict(cfg: dict, seed: int = 1234) -> Dict[str, torch.Tensor]:
    return _fill(arch.unet_param_shapes(cfg), seed)


def vae_state_dict(cfg: dict, seed: int = 1235) -> Dict[str, torch.Tensor]:
    return _fill(arch.vae_param_shapes(cfg), seed)


def vocoder_state_dict(cfg: dict, seed: int = 1236) -> Dict[str, torch.Tensor]:
    sd = _fill(arch.vocoder_param_shapes(cfg), seed)
    # Keep activations O(1) through five upsampling stages and 15-20 residual blocks so the
    # final tanh is not saturated (the reference init N(0, 0.01) at
    # hifigan/models.py:10-13 gives vanishing outputs instead).
    for name in sd:
        w = sd[name]
        if name.startswith("ups.") and name.endswith(".weight"):
            # ConvTranspose1d weight is [Cin, Cout, k]; an output sample sees Cin*k/u taps.
            i = int(name.split(".")[1])
            u = cfg["upsample_rates"][i]
            cin, cout, k = w.shape
            sd[name] = (w * math.sqrt(cout * u / cin)).contiguous()
        elif name.startswith("resblocks.") and name.endswith(".weight"):
            sd[name] = (w * 0.7).contiguous()
        elif name == "conv_post.weight":
            sd[name] = (w * 0.5).contiguous()
    return sd


def conditioning(cfg: dict, batch: int, seed: int = 77, t5_len: int = 32, device="cpu"):
    """Synthetic conditioning at the UNet boundary.

    Returns (cond, uncond), each a dict with ``context_list`` / ``mask_list`` (lists of
    [B, L, D] / [B, L] float tensors) and ``y`` ([B, film_dim] or None), i.e. the keyword
    arguments of ``UNetModel.forward`` (openaimodel.py:837-845).
    """
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    ucfg = cfg["unet"]
    dims = [c for c in ucfg["context_dim"] if c is not None]
    cond = dict(context_list=[], mask_list=[], y=None)
    unc = dict(context_list=[], mask_list=[], y=None)
    for i, d in enumerate(dims):
        L = 8 if i == 0 and len(dims) > 1 else t5_len
        c = torch.randn(batch, L, d, generator=g)
        cond["context_list"].append(c.to(device)); cond["mask_list"].append(torch.ones(batch, L).to(device))
        if i == 0 and len(dims) > 1:        # AudioMAE tokens: uncond = zeros (encoders/modules.py:476-479)
            unc["context_list"].append(torch.zeros(batch, L, d).to(device))
            unc["mask_list"].append(torch.ones(batch, L).to(device))
        else:                               # T5(""): one token
            u = torch.randn(1, 1, d, generator=g).expand(batch, 1, d).contiguous()
            unc["context_list"].append(u.to(device)); unc["mask_list"].append(torch.ones(batch, 1).to(device))
    fd = ucfg.get("extra_film_condition_dim")
    if fd is not None:
        y = torch.randn(batch, fd, generator=g); y = y / y.norm(dim=-1, keepdim=True)
        cond["y"] = y.to(device)
        unc["y"] = torch.zeros(batch, fd).to(device)
    return cond, unc
