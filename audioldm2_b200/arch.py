"""Architecture specs for the three networks on the AudioLDM2 sampling hot path.

The specs are plain Python data derived from the same config keys the reference
uses (``unet_config.params``, ``first_stage_config.params.ddconfig`` and the
HiFi-GAN config dict), so the planner, the weight packer and the oracle all walk
one description of the module tree.  Parameter names are the reference
``state_dict`` keys, relative to the sub-module prefix:

* UNet     : ``model.diffusion_model.``      (openaimodel.py:446-885)
* VAE      : ``first_stage_model.``          (model.py:419-686, autoencoder.py:103-117)
* vocoder  : ``first_stage_model.vocoder.``  (hifigan/models.py:112-174)

Nothing here touches torch; it is pure bookkeeping.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# --------------------------------------------------------------------------------------
# model-name -> config (restated from audioldm2/utils.py:116-702; only hot-path keys)
# --------------------------------------------------------------------------------------

_UNET_BASE = dict(
    in_channels=8, out_channels=8, model_channels=128, attention_resolutions=[8, 4, 2],
    num_res_blocks=2, channel_mult=[1, 2, 3, 5], num_head_channels=32,
    transformer_depth=1, context_dim=[768, 1024], extra_film_condition_dim=None,
    extra_sa_layer=True,
)

_VAE_16K = dict(ch=128, ch_mult=[1, 2, 4], num_res_blocks=2, z_channels=8, in_channels=1,
                out_ch=1, embed_dim=8, double_z=True, mel_bins=64)
_VAE_48K = dict(ch=128, ch_mult=[1, 2, 4, 8], num_res_blocks=2, z_channels=16, in_channels=1,
                out_ch=1, embed_dim=16, double_z=True, mel_bins=256)

_VOC_16K = dict(upsample_rates=[5, 4, 2, 2, 2], upsample_kernel_sizes=[16, 16, 8, 4, 4],
                upsample_initial_channel=1024, resblock_kernel_sizes=[3, 7, 11],
                resblock_dilation_sizes=[[1, 3, 5]] * 3, num_mels=64,
                n_fft=1024, hop_size=160, win_size=1024, sampling_rate=16000, fmin=0, fmax=8000)
_VOC_48K = dict(upsample_rates=[6, 5, 4, 2, 2], upsample_kernel_sizes=[12, 10, 8, 4, 4],
                upsample_initial_channel=1536, resblock_kernel_sizes=[3, 7, 11, 15],
                resblock_dilation_sizes=[[1, 3, 5]] * 4, num_mels=256,
                n_fft=2048, hop_size=480, win_size=2048, sampling_rate=48000, fmin=20, fmax=24000)


def model_config(model_name: str = "audioldm2-full") -> dict:
    """Hot-path subset of ``default_audioldm_config`` (utils.py:116-192).

    Returns ``{"unet", "vae", "vocoder", "latent": (C, T, F), "sampling_rate",
    "latent_t_per_second", "linear_start", "linear_end", "timesteps"}``.
    """
    unet = dict(_UNET_BASE)
    vae, voc = dict(_VAE_16K), dict(_VOC_16K)
    latent = (8, 256, 16)
    sr, tps = 16000, 25.6
    if "-large-" in model_name:                      # utils.py:118-120
        unet["context_dim"] = [768, 1024, None]
        unet["transformer_depth"] = 2
    if "-speech-" in model_name:                     # utils.py:121-123
        unet["context_dim"] = [768]
    if "48k" in model_name:                          # utils.py:413-561
        unet.update(in_channels=16, out_channels=16, context_dim=[None],
                    extra_film_condition_dim=512)
        vae, voc = dict(_VAE_48K), dict(_VOC_48K)
        latent = (16, 128, 32)
        sr, tps = 48000, 12.8
    if "t5" in model_name:                           # utils.py:563-702
        unet["context_dim"] = [1024]
    return dict(name=model_name, unet=unet, vae=vae, vocoder=voc, latent=latent,
                sampling_rate=sr, latent_t_per_second=tps,
                linear_start=0.0015, linear_end=0.0195, timesteps=1000)


def tiny_config(film: bool = False, variant: str = "") -> dict:
    """Shrunken configs with the same topology as the real ones (fast parity tests).

    variant "" / film : audioldm2-full topology (3 STs per site) / audioldm_48k UNet topology (FiLM, 2 self-attn STs)
    variant "large"   : audioldm2-full-large topology: context_dim [.., .., None], transformer_depth 2 (utils.py:118-120)
    variant "48k"     : audioldm_48k first stage: 4-level VAE (ch_mult [1,2,4,8]), HiFi-GAN with 4 MRF kernels
                        (3,7,11,15) and the 48 k upsampling plan (6,5,4,2,2) (utils.py:475-493, utilities/model.py:39-75)
    """
    unet = dict(_UNET_BASE)
    unet.update(model_channels=32, context_dim=[48, 64])
    if film:
        unet.update(context_dim=[None], extra_film_condition_dim=24)
    if variant == "large":
        unet.update(context_dim=[48, 64, None], transformer_depth=2)
    vae = dict(ch=32, ch_mult=[1, 2, 4], num_res_blocks=1, z_channels=8, in_channels=1,
               out_ch=1, embed_dim=8, double_z=True, mel_bins=32)
    voc = dict(upsample_rates=[5, 4, 2, 2, 2], upsample_kernel_sizes=[16, 16, 8, 4, 4],
               upsample_initial_channel=128, resblock_kernel_sizes=[3, 7, 11],
               resblock_dilation_sizes=[[1, 3, 5]] * 3, num_mels=32,
               n_fft=256, hop_size=40, win_size=256, sampling_rate=4000, fmin=0, fmax=2000)
    latent = (8, 32, 8)
    if variant == "48k":
        unet.update(in_channels=16, out_channels=16, context_dim=[None], extra_film_condition_dim=24)
        vae = dict(ch=32, ch_mult=[1, 2, 4, 8], num_res_blocks=1, z_channels=16, in_channels=1,
                   out_ch=1, embed_dim=16, double_z=True, mel_bins=64)
        voc = dict(upsample_rates=[6, 5, 4, 2, 2], upsample_kernel_sizes=[12, 10, 8, 4, 4],
                   upsample_initial_channel=192, resblock_kernel_sizes=[3, 7, 11, 15],
                   resblock_dilation_sizes=[[1, 3, 5]] * 4, num_mels=64,
                   n_fft=256, hop_size=40, win_size=256, sampling_rate=4000, fmin=0, fmax=2000)
        latent = (16, 16, 8)
    name = "tiny" + ("-film" if film else "") + (("-" + variant) if variant else "")
    return dict(name=name, unet=unet, vae=vae, vocoder=voc,
                latent=latent, sampling_rate=4000, latent_t_per_second=25.6,
                linear_start=0.0015, linear_end=0.0195, timesteps=1000)


# --------------------------------------------------------------------------------------
# UNet spec (openaimodel.py:576-811)
# --------------------------------------------------------------------------------------

@dataclass
class Layer:
    kind: str                 # conv | res | st | down | up
    name: str                 # state_dict prefix relative to the net, e.g. "input_blocks.4.0"
    cin: int = 0
    cout: int = 0
    heads: int = 0
    ctx_dim: Optional[int] = None   # construction-time context_dim of the ST (None => self-attn weights)
    depth: int = 1
    ctx_slot: int = -1              # index into context_list used at run time, -1 => None (self-attention)


@dataclass
class UNetSpec:
    cfg: dict
    emb_ch: int               # width of `emb` fed to ResBlocks (time_embed_dim or 2x with FiLM)
    time_embed_dim: int
    input_blocks: List[List[Layer]] = field(default_factory=list)
    middle: List[Layer] = field(default_factory=list)
    output_blocks: List[List[Layer]] = field(default_factory=list)
    skip_ch: List[int] = field(default_factory=list)   # channels pushed by each input block


def _route_contexts(layers: List[Layer], n_ctx: int) -> None:
    """TimestepEmbedSequential.forward (openaimodel.py:81-103): the i-th SpatialTransformer
    of a block gets ([None] + context_list)[i], or None when out of range."""
    st_id = 0
    for l in layers:
        if l.kind == "st":
            l.ctx_slot = (st_id - 1) if (1 <= st_id <= n_ctx) else -1
            st_id += 1


def unet_spec(cfg: dict) -> UNetSpec:
    mc = cfg["model_channels"]
    ted = mc * 4
    film = cfg.get("extra_film_condition_dim") is not None
    emb_ch = ted * 2 if film else ted
    context_dim = cfg.get("context_dim")
    if context_dim is None:
        context_dim = [None]
    if not isinstance(context_dim, list):
        context_dim = [context_dim]
    # number of *run-time* contexts: entries of context_dim that are real cross-attn dims
    n_ctx = len([c for c in context_dim if c is not None])
    nhc = cfg["num_head_channels"]
    depth = cfg.get("transformer_depth", 1)
    extra_sa = cfg.get("extra_sa_layer", True)
    spec = UNetSpec(cfg=cfg, emb_ch=emb_ch, time_embed_dim=ted)

    def st_layers(prefix: str, start: int, ch: int) -> List[Layer]:
        out = []
        idx = start
        if extra_sa:
            out.append(Layer("st", f"{prefix}.{idx}", ch, ch, ch // nhc, None, depth)); idx += 1
        for cd in context_dim:
            out.append(Layer("st", f"{prefix}.{idx}", ch, ch, ch // nhc, cd, depth)); idx += 1
        return out

    spec.input_blocks.append([Layer("conv", "input_blocks.0.0", cfg["in_channels"], mc)])
    chans = [mc]
    ch, ds, bi = mc, 1, 1
    cm = cfg["channel_mult"]
    for level, mult in enumerate(cm):
        for _ in range(cfg["num_res_blocks"]):
            layers = [Layer("res", f"input_blocks.{bi}.0", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg["attention_resolutions"]:
                layers += st_layers(f"input_blocks.{bi}", 1, ch)
            spec.input_blocks.append(layers); chans.append(ch); bi += 1
        if level != len(cm) - 1:
            spec.input_blocks.append([Layer("down", f"input_blocks.{bi}.0", ch, ch)])
            chans.append(ch); bi += 1; ds *= 2
    spec.skip_ch = list(chans)

    mid = [Layer("res", "middle_block.0", ch, ch)]
    mid += st_layers("middle_block", 1, ch)
    mid.append(Layer("res", f"middle_block.{len(mid)}", ch, ch))
    spec.middle = mid

    bo = 0
    for level, mult in list(enumerate(cm))[::-1]:
        for i in range(cfg["num_res_blocks"] + 1):
            ich = chans.pop()
            layers = [Layer("res", f"output_blocks.{bo}.0", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg["attention_resolutions"]:
                layers += st_layers(f"output_blocks.{bo}", 1, ch)
            if level and i == cfg["num_res_blocks"]:
                layers.append(Layer("up", f"output_blocks.{bo}.{len(layers)}", ch, ch))
                ds //= 2
            spec.output_blocks.append(layers); bo += 1

    for blk in spec.input_blocks + [spec.middle] + spec.output_blocks:
        _route_contexts(blk, n_ctx)
    return spec


def unet_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """name -> shape for every UNet parameter (the reference ``state_dict`` of UNetModel)."""
    s = unet_spec(cfg)
    mc, ted = cfg["model_channels"], s.time_embed_dim
    P: Dict[str, Tuple[int, ...]] = {}

    def lin(n, o, i, bias=True):
        P[n + ".weight"] = (o, i)
        if bias:
            P[n + ".bias"] = (o,)

    def conv(n, o, i, k):
        P[n + ".weight"] = (o, i, k, k); P[n + ".bias"] = (o,)

    def norm(n, c):
        P[n + ".weight"] = (c,); P[n + ".bias"] = (c,)

    lin("time_embed.0", ted, mc); lin("time_embed.2", ted, ted)
    if cfg.get("extra_film_condition_dim") is not None:
        lin("film_emb", ted, cfg["extra_film_condition_dim"])

    def add(l: Layer):
        n = l.name
        if l.kind == "conv":
            conv(n, l.cout, l.cin, 3)
        elif l.kind == "res":
            norm(n + ".in_layers.0", l.cin); conv(n + ".in_layers.2", l.cout, l.cin, 3)
            lin(n + ".emb_layers.1", l.cout, s.emb_ch)
            norm(n + ".out_layers.0", l.cout); conv(n + ".out_layers.3", l.cout, l.cout, 3)
            if l.cin != l.cout:
                conv(n + ".skip_connection", l.cout, l.cin, 1)
        elif l.kind == "down":
            conv(n + ".op", l.cout, l.cin, 3)
        elif l.kind == "up":
            conv(n + ".conv", l.cout, l.cin, 3)
        elif l.kind == "st":
            c = l.cin
            norm(n + ".norm", c); conv(n + ".proj_in", c, c, 1); conv(n + ".proj_out", c, c, 1)
            for d in range(l.depth):
                b = f"{n}.transformer_blocks.{d}"
                for a, cd in (("attn1", None), ("attn2", l.ctx_dim)):
                    kd = c if cd is None else cd
                    lin(f"{b}.{a}.to_q", c, c, bias=False)
                    lin(f"{b}.{a}.to_k", c, kd, bias=False)
                    lin(f"{b}.{a}.to_v", c, kd, bias=False)
                    lin(f"{b}.{a}.to_out.0", c, c)
                lin(f"{b}.ff.net.0.proj", 8 * c, c); lin(f"{b}.ff.net.2", c, 4 * c)
                for k in ("norm1", "norm2", "norm3"):
                    norm(f"{b}.{k}", c)

    for blk in s.input_blocks + [s.middle] + s.output_blocks:
        for l in blk:
            add(l)
    norm("out.0", mc); conv("out.2", cfg["out_channels"], mc, 3)
    return P


# --------------------------------------------------------------------------------------
# VAE spec (model.py:419-686) -- Decoder and Encoder, attn_resolutions == []
# --------------------------------------------------------------------------------------

def vae_param_shapes(cfg: dict, encoder: bool = True, decoder: bool = True) -> Dict[str, Tuple[int, ...]]:
    P: Dict[str, Tuple[int, ...]] = {}
    ch, cm, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    zc, ed = cfg["z_channels"], cfg["embed_dim"]

    def conv(n, o, i, k):
        P[n + ".weight"] = (o, i, k, k); P[n + ".bias"] = (o,)

    def norm(n, c):
        P[n + ".weight"] = (c,); P[n + ".bias"] = (c,)

    def res(n, i, o):
        norm(n + ".norm1", i); conv(n + ".conv1", o, i, 3)
        norm(n + ".norm2", o); conv(n + ".conv2", o, o, 3)
        if i != o:
            conv(n + ".nin_shortcut", o, i, 1)

    def attn(n, c):
        norm(n + ".norm", c)
        for k in ("q", "k", "v", "proj_out"):
            conv(f"{n}.{k}", c, c, 1)

    if decoder:
        bi = ch * cm[-1]
        conv("decoder.conv_in", bi, zc, 3)
        res("decoder.mid.block_1", bi, bi); attn("decoder.mid.attn_1", bi); res("decoder.mid.block_2", bi, bi)
        for lvl in reversed(range(len(cm))):
            bo = ch * cm[lvl]
            for ib in range(nrb + 1):
                res(f"decoder.up.{lvl}.block.{ib}", bi, bo); bi = bo
            if lvl != 0:
                conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
        norm("decoder.norm_out", bi); conv("decoder.conv_out", cfg["out_ch"], bi, 3)
        conv("post_quant_conv", zc, ed, 1)
    if encoder:
        conv("encoder.conv_in", ch, cfg["in_channels"], 3)
        in_mult = (1,) + tuple(cm)
        bi = ch
        for lvl in range(len(cm)):
            bi = ch * in_mult[lvl]; bo = ch * cm[lvl]
            for ib in range(nrb):
                res(f"encoder.down.{lvl}.block.{ib}", bi, bo); bi = bo
            if lvl != len(cm) - 1:
                conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
        res("encoder.mid.block_1", bi, bi); attn("encoder.mid.attn_1", bi); res("encoder.mid.block_2", bi, bi)
        norm("encoder.norm_out", bi)
        conv("encoder.conv_out", 2 * zc if cfg.get("double_z", True) else zc, bi, 3)
        conv("quant_conv", 2 * ed, 2 * zc, 1)
    return P


# --------------------------------------------------------------------------------------
# HiFi-GAN spec (hifigan/models.py:112-147); weight-norm already folded (utilities/model.py:139)
# --------------------------------------------------------------------------------------

def vocoder_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    P: Dict[str, Tuple[int, ...]] = {}
    c0 = cfg["upsample_initial_channel"]
    P["conv_pre.weight"] = (c0, cfg["num_mels"], 7); P["conv_pre.bias"] = (c0,)
    nk = len(cfg["resblock_kernel_sizes"])
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        P[f"ups.{i}.weight"] = (cin, ch, k); P[f"ups.{i}.bias"] = (ch,)   # ConvTranspose1d: [Cin,Cout,k]
        for j, ks in enumerate(cfg["resblock_kernel_sizes"]):
            for m in range(3):
                for cs in ("convs1", "convs2"):
                    P[f"resblocks.{i * nk + j}.{cs}.{m}.weight"] = (ch, ch, ks)
                    P[f"resblocks.{i * nk + j}.{cs}.{m}.bias"] = (ch,)
    P["conv_post.weight"] = (1, ch, 7); P["conv_post.bias"] = (1,)
    return P


def vocoder_out_len(cfg: dict, n_frames: int) -> int:
    L = n_frames
    for u, k in zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]):
        L = (L - 1) * u - 2 * ((k - u) // 2) + k
    return L
