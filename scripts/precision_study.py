"""CPU precision study (test infrastructure; uses oracle/): what operand rounding does the 200-step waveform tolerate?

Runs the fp32 oracle DDIM loop of audioldm2-full (B=1, fixture noise) with the A operand (activations) and/or the B
operand (weights) of every UNet contraction rounded to a narrower format, and reports the relative L2 error of the final
latent / mel / waveform against the committed reference fixture (tests/golden/ddim_full_<S>.pt).

    python scripts/precision_study.py --steps 10 --mode f16a
modes: fp32 | f16a (activations fp16, weights exact) | f16aw (both fp16) | bf16a | tf32 (both 10-bit mantissa)
       f16a_attn32 (fp16 activations except attention QK^T / PV operands)
"""
import argparse, os, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioldm2_b200 import arch, synth
from oracle import functional as OF
from tests.golden import cases

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--mode", default="f16a")
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--only", default="", help="comma list of classes that get rounded: conv,lin,ff,attn,proj (default: all)")
ap.add_argument("--lnfold", action="store_true",
                help="LayerNorm folded into its consumer GEMM (DESIGN.md 8.3): the tensor core consumes fp16(x) of the UN-normalised "
                     "residual stream and the epilogue applies y = rstd (x_h W'^T - mu s) + b', i.e. the consumer sees "
                     "LN(x) + rstd * gamma * (fp16(x) - x) instead of fp16(LN(x))")
a = ap.parse_args()
torch.set_num_threads(a.threads)


def r_f16(x): return x.half().float()
def r_bf16(x): return x.bfloat16().float()
def r_tf32(x):   # round-to-nearest to 10 explicit mantissa bits
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)
ident = lambda x: x

mode = a.mode
ra = dict(fp32=ident, f16a=r_f16, f16aw=r_f16, bf16a=r_bf16, tf32=r_tf32, f16a_attn32=r_f16)[mode]
rw = dict(fp32=ident, f16a=ident, f16aw=r_f16, bf16a=ident, tf32=r_tf32, f16a_attn32=ident)[mode]
rattn = ident if mode in ("fp32", "f16a_attn32") else ra
ONLY = set(a.only.split(",")) if a.only else None
def cls_of(n, is_conv):
    if ".ff.net" in n: return "ff"
    if ".attn1." in n or ".attn2." in n: return "lin"
    if "proj_in" in n or "proj_out" in n: return "proj"
    return "conv" if is_conv else "lin"
LN_CONSUMERS = (".to_q", ".to_k", ".to_v", ".ff.net.0.proj")
def RA(n, is_conv=False):
    if a.lnfold and n.endswith(LN_CONSUMERS) and getattr(RA, "from_ln", False):
        return ident                      # the folded consumer reads fp16(x) itself: its error is injected by _ln below
    return ra if (ONLY is None or cls_of(n, is_conv) in ONLY) else ident
if ONLY is not None and "attn" not in ONLY: rattn = ident

_conv2d0, _lin0 = OF._conv2d, OF._lin
_ln0 = OF._ln
def _ln(sd, n, x):
    y = _ln0(sd, n, x)
    if not a.lnfold:
        return y
    rstd = torch.rsqrt(x.var(dim=-1, unbiased=False, keepdim=True) + 1e-5)
    return y + rstd * sd[n + ".weight"] * (r_f16(x) - x)
OF._ln = _ln
def _conv2d(sd, n, x, stride=1, padding=0):
    return F.conv2d(RA(n, True)(x), rw(sd[n + ".weight"]), sd.get(n + ".bias"), stride=stride, padding=padding)
def _lin(sd, n, x):
    RA.from_ln = n.endswith(".ff.net.0.proj")
    r = RA(n)
    RA.from_ln = False
    return F.linear(r(x), rw(sd[n + ".weight"]), sd.get(n + ".bias"))
def _cross_attention(sd, n, x, heads, context=None, mask=None):
    ctx = x if context is None else context
    RA.from_ln = True
    q = F.linear(RA(n + ".to_q")(x), rw(sd[n + ".to_q.weight"]))
    RA.from_ln = context is None          # K / V of a cross-attention come from the context, not from the LayerNorm
    k = F.linear(RA(n + ".to_k")(ctx), rw(sd[n + ".to_k.weight"]))
    v = F.linear(RA(n + ".to_v")(ctx), rw(sd[n + ".to_v.weight"]))
    RA.from_ln = False
    b, nq, c = q.shape
    d = c // heads
    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", rattn(q), rattn(k)) * (d ** -0.5)
    if mask is not None:
        m = mask.reshape(b, -1)
        m = m[:, None, None, :].expand(b, heads, 1, m.shape[-1]).reshape(b * heads, 1, -1)
        sim = sim.masked_fill(~(m == 1), -torch.finfo(sim.dtype).max)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", rattn(attn), rattn(v))
    out = out.reshape(b, heads, nq, d).permute(0, 2, 1, 3).reshape(b, nq, c)
    return _lin(sd, n + ".to_out.0", out)
OF._conv2d, OF._lin, OF._cross_attention = _conv2d, _lin, _cross_attention

cfg = arch.model_config("audioldm2-full")
S = a.steps
fx = cases.load(f"ddim_full_{S}")
usd = synth.unet_state_dict(cfg["unet"])
_, _, cond, unc = cases.unet_inputs(cfg, 1, t5_len=32)
x_T, noises, _ = cases.sampler_noise(cfg, 1, S)
t0 = time.time()
with torch.no_grad():
    z = OF.ddim_sample(usd, cfg["unet"], x_T, noises, cond, unc, S, 1.0, 3.5,
                       OF.ddpm_tables(cfg["linear_start"], cfg["linear_end"], cfg["timesteps"]))
    rel = lambda a_, b_: float(torch.linalg.norm(a_ - b_) / torch.linalg.norm(b_))
    e_lat = rel(z, fx["latent"])
    OF._conv2d, OF._lin = _conv2d0, _lin0
    mel = OF.vae_decode(synth.vae_state_dict(cfg["vae"]), cfg["vae"], z)
    wave = OF.vocoder_forward(synth.vocoder_state_dict(cfg["vocoder"]), cfg["vocoder"], mel.squeeze(1).permute(0, 2, 1))
print(f"RESULT only={a.only} mode={mode} steps={S} latent_rel={e_lat:.3e} mel_rel={rel(mel, fx['mel']):.3e} wave_rel={rel(wave, fx['wave']):.3e} "
      f"time={time.time() - t0:.0f}s", flush=True)
