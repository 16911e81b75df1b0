#!/usr/bin/env python
"""HBM-bound kernels of the path (K6 CFG+DDIM update, K9 STFT/mel, K3 GroupNorm / LayerNorm prep):
achieved GB/s against the measured copy bandwidth (MEASURED_PEAKS.json), timed with CUDA events over
many back-to-back launches on inputs larger than L2 where the shape allows.  Also the ncu target for
these kernels (scripts/gpu_final.sh)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audioldm2_b200 import _lib, engine  # noqa: E402
from audioldm2_b200.plan import F32, Planner  # noqa: E402


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    dev = torch.device("cuda:0")
    peak = 6572.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {}
    # K6: workload shape (B=8 latents, 1 MB per tensor: L2 resident) and a 512 MB-per-tensor stream (HBM)
    for name, n in (("ddim_step_B8", 8 * 8 * 256 * 16), ("ddim_step_stream", 128 * 1024 * 1024)):
        x, eu, ec, nz, o = (torch.randn(n, device=dev) for _ in range(5))
        t = timed(lambda: engine.ddim_step(x, eu, ec, nz, o, 0.5, 0.6, 0.1, 0.7, 3.5), reps if n < 1e7 else 10)
        out[name] = dict(us=t * 1e6, algorithmic_bytes=20 * n, gbs=20 * n / t / 1e9, frac_of_measured_hbm=20 * n / t / 1e9 / peak)
        del x, eu, ec, nz, o
    # K9: 8 clips of 10.24 s @16 kHz, n_fft 1024 hop 160, 64 mels (config of pipeline.py:236-245)
    B, T = 8, 163840
    wav = (torch.rand(B, T, device=dev) - 0.5).contiguous()
    basis = torch.rand(64, 513, device=dev)
    t = timed(lambda: engine.stft_mel(wav, 1024, 160, basis, 1024), reps)
    by = 4 * B * T + 4 * B * 1024 * 64
    out["stft_mel_16k_B8"] = dict(us=t * 1e6, algorithmic_bytes=by, gbs=by / t / 1e9, frac_of_measured_hbm=by / t / 1e9 / peak)
    # K3: GroupNorm+SiLU prep on the largest UNet activation (2*B_l=16, 4096 px, 128 ch) and a LayerNorm
    for name, rows, Cc, mode, Bn in (("gn_silu_65536x128", 65536, 128, _lib.PREP_GN_SILU, 16),
                                     ("gn_silu_vae_524288x128", 524288, 128, _lib.PREP_GN_SILU, 8),
                                     ("ln_16384x256", 16384, 256, _lib.PREP_LN, 0)):
        P = Planner()
        src = F32(P.raw(rows * Cc * 4), rows, Cc)
        P.prep(mode, src, None, P.vec(torch.ones(Cc)), P.vec(torch.zeros(Cc)), eps=1e-5, B=Bn, HW=rows // max(Bn, 1))
        pl = P.finish(dict(src=("f32", src.ref, (rows, Cc))))
        prog = engine.DeviceProgram(pl, dev, dict(op=(0, len(pl.ops))))
        prog.view("src").copy_(torch.randn(rows, Cc, device=dev))
        t = timed(lambda: prog.run("op"), reps)
        by = rows * Cc * (4 + 4 + (4 if mode != _lib.PREP_LN else 0))      # GN reads x twice (stats, apply), writes 2 bf16 planes
        out[name] = dict(us=t * 1e6, algorithmic_bytes=by, gbs=by / t / 1e9, frac_of_measured_hbm=by / t / 1e9 / peak)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
