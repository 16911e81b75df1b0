"""GPU parity tests of whole networks and of the sampler, through the native engine (C-ABI), against
(a) the committed golden fixtures = outputs of the unmodified reference modules, and
(b) the CPU oracle on the same seeded inputs.

Tolerances: single network evaluation 1e-4 relative L2 (measured ~1e-5: operand planes carry
2^-17 relative error, accumulation is fp32); end-to-end waveform after 10 / 200 DDIM steps 1e-3
relative L2 -- the tolerance BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

from audioldm2_b200 import arch, model, synth
from tests.conftest import rel_l2
from tests.golden import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NET_TOL = 1e-4
WAVE_TOL = 1e-3


def _to(c, dev):
    return dict(context_list=[t.to(dev) for t in c["context_list"]], mask_list=[t.to(dev) for t in c["mask_list"]],
                y=None if c["y"] is None else c["y"].to(dev))


def _engine(cfg, B, t5_len, **kw):
    lens = (8, t5_len) if len([c for c in cfg["unet"]["context_dim"] if c is not None]) > 1 else (t5_len,)
    return model.NativeLatentDiffusion(cfg, synth.unet_state_dict(cfg["unet"]), synth.vae_state_dict(cfg["vae"]),
                                       synth.vocoder_state_dict(cfg["vocoder"]), B, DEV, ctx_max_len=lens, **kw)


@pytest.fixture(scope="module")
def tiny_tc():
    return _engine(arch.tiny_config(), 2, 5, with_encoder=True)


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_unet_tiny(impl, tiny_tc):
    cfg = arch.tiny_config()
    eng = tiny_tc if impl == "tc" else _engine(cfg, 2, 5, impl="simt", use_graph=False)
    g = cases.load("unet_tiny")
    x, t, cond, unc = cases.unet_inputs(cfg, 2, t5_len=5)
    eng.set_conditioning(_to(cond, DEV), _to(unc, DEV))
    e_u, e_c = eng.apply_model_pair(x.to(DEV), int(t[0]))
    assert torch.isfinite(e_c).all()
    assert rel_l2(e_u, g["eps_uncond"]) < NET_TOL, impl
    assert rel_l2(e_c, g["eps_cond"]) < NET_TOL, impl
    # graph replay gives the same answer as the eager run
    e_u2, e_c2 = eng.apply_model_pair(x.to(DEV), int(t[0]))
    assert rel_l2(e_c2, g["eps_cond"]) < NET_TOL


def test_unet_tiny_film():
    cfg = arch.tiny_config(film=True)
    eng = _engine(cfg, 2, 32)
    g = cases.load("unet_tiny_film")
    x, t, cond, unc = cases.unet_inputs(cfg, 2)
    eng.set_conditioning(_to(cond, DEV), _to(unc, DEV))
    e_u, e_c = eng.apply_model_pair(x.to(DEV), int(t[0]))
    assert rel_l2(e_u, g["eps_uncond"]) < NET_TOL and rel_l2(e_c, g["eps_cond"]) < NET_TOL


def test_unet_tiny_large_topology():
    cfg = arch.tiny_config(variant="large")
    eng = _engine(cfg, 2, 5)
    g = cases.load("unet_tiny_large")
    x, t, cond, unc = cases.unet_inputs(cfg, 2, t5_len=5)
    eng.set_conditioning(_to(cond, DEV), _to(unc, DEV))
    e_u, e_c = eng.apply_model_pair(x.to(DEV), int(t[0]))
    assert rel_l2(e_u, g["eps_uncond"]) < NET_TOL and rel_l2(e_c, g["eps_cond"]) < NET_TOL


def test_tiny_48k_topology():
    """audioldm_48k: FiLM UNet (16-ch latent), 4-level VAE decoder/encoder, 48 k HiFi-GAN plan (k up to 15)."""
    from audioldm2_b200 import engine, plan
    cfg = arch.tiny_config(variant="48k")
    eng = _engine(cfg, 2, 32, with_encoder=True)
    g = cases.load("unet_tiny_48k")
    x, t, cond, unc = cases.unet_inputs(cfg, 2)
    eng.set_conditioning(_to(cond, DEV), _to(unc, DEV))
    e_u, e_c = eng.apply_model_pair(x.to(DEV), int(t[0]))
    assert rel_l2(e_u, g["eps_uncond"]) < NET_TOL and rel_l2(e_c, g["eps_cond"]) < NET_TOL
    gv = cases.load("vae_tiny_48k")
    assert rel_l2(eng.decode_first_stage(cases.latent(cfg, 2, seed=5).to(DEV)), gv["mel"]) < NET_TOL
    mom = eng.encode_first_stage_moments(cases.mel_input(cfg, 2).to(DEV))
    assert rel_l2(mom.permute(0, 3, 1, 2), gv["moments"]) < NET_TOL
    gw = cases.load("vocoder_tiny_48k")
    pv = plan.build_vocoder(synth.vocoder_state_dict(cfg["vocoder"]), cfg["vocoder"], 16, 2)
    prog = engine.DeviceProgram(pv, torch.device(DEV), dict(all=(0, len(pv.ops))))
    prog.view("mel").copy_(cases.vocoder_input(cfg, 2, 16).permute(0, 2, 1).contiguous().to(DEV))
    prog.run("all")
    assert rel_l2(prog.view("wave"), gw["wave"]) < NET_TOL


def test_vae_and_vocoder_tiny(tiny_tc):
    cfg = arch.tiny_config()
    g = cases.load("vae_tiny")
    mel = tiny_tc.decode_first_stage(cases.latent(cfg, 2, seed=5).to(DEV))
    assert rel_l2(mel, g["mel"]) < NET_TOL
    mom = tiny_tc.encode_first_stage_moments(cases.mel_input(cfg, 2).to(DEV))
    assert rel_l2(mom.permute(0, 3, 1, 2), g["moments"]) < NET_TOL
    gv = cases.load("vocoder_tiny")
    # vocoder program is planned for the decoder's frame count; run the fixture through a dedicated plan
    from audioldm2_b200 import engine, plan
    pv = plan.build_vocoder(synth.vocoder_state_dict(cfg["vocoder"]), cfg["vocoder"], 24, 2)
    prog = engine.DeviceProgram(pv, torch.device(DEV), dict(all=(0, len(pv.ops))))
    prog.view("mel").copy_(cases.vocoder_input(cfg, 2, 24).permute(0, 2, 1).contiguous().to(DEV))
    prog.run("all")
    assert rel_l2(prog.view("wave"), gv["wave"]) < NET_TOL


@pytest.mark.parametrize("masked", [False, True])
def test_ddim_tiny_vs_reference(masked, tiny_tc):
    cfg = arch.tiny_config()
    g = cases.load("ddim_tiny_masked" if masked else "ddim_tiny")
    _, _, cond, unc = cases.unet_inputs(cfg, 2, t5_len=5)
    x_T, noises, qn = cases.sampler_noise(cfg, 2, 5, masked=masked)
    mask = x0 = None
    if masked:
        mask, x0 = cases.inpaint_mask(cfg, 2)
        mask, x0 = mask.to(DEV), x0.to(DEV)
    nf = lambda i, kind: (qn[i] if kind == "q" else noises[i]).to(DEV)
    z = tiny_tc.generate_latent(_to(cond, DEV), _to(unc, DEV), ddim_steps=5, guidance=3.5, eta=1.0, x_T=x_T, noise_fn=nf,
                                mask=mask, x0=x0)
    assert rel_l2(z, g["latent"]) < 2e-4


# ---------------------------------------------------------------------------------------------
# full-size configuration (audioldm2-full), B = 1: the exact shapes of BASELINE config C1/C2
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full():
    return _engine(arch.model_config("audioldm2-full"), 1, 32)


def test_unet_full_vs_reference(full):
    cfg = arch.model_config("audioldm2-full")
    g = cases.load("unet_full")
    x, t, cond, unc = cases.unet_inputs(cfg, 1)
    full.set_conditioning(_to(cond, DEV), _to(unc, DEV))
    e_u, e_c = full.apply_model_pair(x.to(DEV), int(t[0]))
    assert rel_l2(e_u, g["eps_uncond"]) < NET_TOL
    assert rel_l2(e_c, g["eps_cond"]) < NET_TOL


def test_vae_vocoder_full_vs_reference(full):
    cfg = arch.model_config("audioldm2-full")
    g = cases.load("vae_full")
    mel = full.decode_first_stage(cases.latent(cfg, 1, seed=5).to(DEV))
    assert rel_l2(mel, g["mel"]) < NET_TOL
    gv = cases.load("vocoder_full")
    melin = cases.vocoder_input(cfg, 1, 1024).permute(0, 2, 1).contiguous()[:, None]      # [B,1,T,F]
    w = full.mel_spectrogram_to_waveform(melin.to(DEV))
    assert w.shape == (1, 1, 163872)
    assert rel_l2(w, gv["wave"]) < NET_TOL


@pytest.mark.parametrize("steps", [10, 200])
def test_end_to_end_waveform_vs_reference(steps, full):
    """x_T -> S x (2 UNet + update) -> VAE decode -> HiFi-GAN, identical noise; waveform within 1e-3."""
    cfg = arch.model_config("audioldm2-full")
    g = cases.load(f"ddim_full_{steps}")
    _, _, cond, unc = cases.unet_inputs(cfg, 1)
    x_T, noises, _ = cases.sampler_noise(cfg, 1, steps)
    nf = lambda i, kind: noises[i].to(DEV)
    z = full.generate_latent(_to(cond, DEV), _to(unc, DEV), ddim_steps=steps, guidance=3.5, eta=1.0, x_T=x_T, noise_fn=nf)
    e_lat = rel_l2(z, g["latent"])
    mel = full.decode_first_stage(z)
    e_mel = rel_l2(mel, g["mel"])
    wave = full.mel_spectrogram_to_waveform(mel)
    e_wav = rel_l2(wave, g["wave"])
    print(f"steps={steps}: latent {e_lat:.2e} mel {e_mel:.2e} waveform {e_wav:.2e}")
    assert e_lat < WAVE_TOL and e_mel < WAVE_TOL and e_wav < WAVE_TOL


# ---------------------------------------------------------------------------------------------
# the public pipeline surface (pipeline.py:142-267) on the tiny topology, against the CPU oracle fed with the
# replayed RNG draws (CUDA generator for x_T / step / q_sample noise, CPU generator for the posterior sample)
# ---------------------------------------------------------------------------------------------
def _replay_cuda_noise(seed, shape, S, masked):
    torch.manual_seed(seed); torch.cuda.manual_seed(seed)
    x_T = torch.randn(shape, device=DEV).cpu()
    noises, qn = [], []
    for _ in range(S):
        if masked:
            qn.append(torch.randn(shape, device=DEV).cpu())
        noises.append(torch.randn(shape, device=DEV).cpu())
    return x_T, noises, qn


def _oracle_wave(cfg, z):
    from oracle import functional as OF
    mel = OF.vae_decode(synth.vae_state_dict(cfg["vae"]), cfg["vae"], z)
    return OF.vocoder_forward(synth.vocoder_state_dict(cfg["vocoder"]), cfg["vocoder"], mel.squeeze(1).permute(0, 2, 1))


def test_pipeline_text_to_audio_tiny():
    import audioldm2_b200 as A
    from oracle import functional as OF
    cfg = arch.tiny_config()
    ld = A.build_model(config=cfg, t5_len=5)
    B, S, seed = 2, 4, 7
    wav = A.text_to_audio(ld, "a dog barking", seed=seed, ddim_steps=S, duration=1.25, batchsize=B, n_candidate_gen_per_text=1)
    assert isinstance(wav, np.ndarray) and wav.dtype == np.float32 and wav.shape[:2] == (B, 1)
    C_, T, F_ = cfg["latent"]
    x_T, noises, _ = _replay_cuda_noise(seed, (B, C_, T, F_), S, False)
    cond, unc = synth.conditioning(cfg, B, seed=77, t5_len=5)
    with torch.no_grad():
        z = OF.ddim_sample(synth.unet_state_dict(cfg["unet"]), cfg["unet"], x_T, noises, cond, unc, S, 1.0, 3.5,
                           OF.ddpm_tables(cfg["linear_start"], cfg["linear_end"], cfg["timesteps"]))
        ref = _oracle_wave(cfg, z)
    assert rel_l2(torch.from_numpy(wav), ref) < WAVE_TOL
    # n_candidate_gen_per_text > 1: candidates of prompt i are rows i + k*B; a ranker picks per prompt (ddpm.py:1554-1564)
    picked = {}

    def ranker(w, texts):
        assert w.shape[0] == 2 * B and len(texts) == 2 * B
        picked["sim"] = torch.tensor([0.0, 1.0, 1.0, 0.0])          # prompt 0 -> candidate 1 (row 2), prompt 1 -> candidate 0 (row 1)
        picked["w"] = w.clone()
        return picked["sim"]
    ld.ranker = ranker
    out = A.text_to_audio(ld, "a dog barking", seed=seed, ddim_steps=S, duration=1.25, batchsize=B, n_candidate_gen_per_text=2)
    assert out.shape == wav.shape
    assert np.array_equal(out[0, 0], picked["w"][2].numpy()) and np.array_equal(out[1, 0], picked["w"][1].numpy())
    # re-plan on a batch change (engine cache), reference call sequence with the API defaults' shape
    ld.ranker = None
    with pytest.warns(UserWarning):
        w1 = A.text_to_audio(ld, "x", seed=1, ddim_steps=2, duration=1.25, batchsize=1)       # n_candidate_gen_per_text=3 default
    assert w1.shape[:2] == (1, 1)


def test_pipeline_super_resolution_and_inpainting_tiny():
    import audioldm2_b200 as A
    from oracle import functional as OF
    from oracle import mel as OM
    cfg = arch.tiny_config()
    vc = cfg["vocoder"]
    ld = A.build_model(config=cfg, t5_len=5)
    B, S, seed = 2, 3, 11
    wav_in = cases.wav_input(5000).numpy()[0]                      # longer than the segment: cropped (tools.py:8-18)
    dur = 1.28                                                     # 128 mel frames at hop 40 / 4 kHz -> latent T = 32
    out = A.super_resolution_and_inpainting(ld, "x", seed=seed, ddim_steps=S, duration=dur, batchsize=B, n_candidate_gen_per_text=1,
                                            waveform=wav_in, waveform_sr=vc["sampling_rate"])
    # oracle: same front end, encoder, posterior, mask, masked DDIM
    from audioldm2_b200 import frontend
    x = np.clip(frontend.prepare_waveform(wav_in, 4000, 4000, 128 * vc["hop_size"]), -1, 1)
    logmel, _ = OM.stft_mel(x, vc["n_fft"], vc["hop_size"], vc["num_mels"], vc["sampling_rate"], vc["fmin"], vc["fmax"])
    fb = torch.from_numpy(logmel[0].T[:128]).float()                                     # [T', F']
    C_, T, F_ = cfg["latent"]
    torch.manual_seed(seed)
    post = torch.randn(B, C_, T, F_)                                                     # CPU draw (distributions.py:38)
    x_T, noises, qn = _replay_cuda_noise(seed, (B, C_, T, F_), S, True)
    vsd = synth.vae_state_dict(cfg["vae"])
    with torch.no_grad():
        mom = OF.vae_encode_moments(vsd, cfg["vae"], fb[None, None].expand(B, 1, -1, -1).contiguous())
        x0 = OF.posterior_sample(mom, post, 1.0)
        mask = torch.ones(B, 1, T, F_)
        mask[:, :, int(T * 0.40):int(T * 0.6), :] = 0                                    # pipeline.py:224 defaults
        cond, unc = synth.conditioning(cfg, B, seed=77, t5_len=5)
        z = OF.ddim_sample(synth.unet_state_dict(cfg["unet"]), cfg["unet"], x_T, noises, cond, unc, S, 1.0, 2.5,
                           OF.ddpm_tables(cfg["linear_start"], cfg["linear_end"], cfg["timesteps"]), mask=mask, x0=x0, q_noises=qn)
        ref = _oracle_wave(cfg, z)
    assert out.shape == tuple(ref.shape)
    assert rel_l2(torch.from_numpy(out), ref) < WAVE_TOL


def test_rank_shards_reproduce_single_process_batch(tiny_tc):
    """two ranks (B = 1 each, full-batch noise drawn and sliced) == one process with B = 2."""
    from audioldm2_b200 import parallel
    cfg = arch.tiny_config()
    S = 3
    cond, unc = synth.conditioning(cfg, 2, seed=77, t5_len=5)
    sn = parallel.ShardedNoise(2, 0, 2, cfg["latent"], DEV, seed=42)
    z_full = tiny_tc.generate_latent(_to(cond, DEV), _to(unc, DEV), ddim_steps=S, guidance=3.5, eta=1.0, x_T=sn.x_T(), noise_fn=sn).clone()
    e1 = _engine(cfg, 1, 5)
    for r in range(2):
        sr_ = parallel.ShardedNoise(2, r, r + 1, cfg["latent"], DEV, seed=42)
        c, u = parallel.shard_rows(cond, r, r + 1), parallel.shard_rows(unc, r, r + 1)
        z = e1.generate_latent(_to(c, DEV), _to(u, DEV), ddim_steps=S, guidance=3.5, eta=1.0, x_T=sr_.x_T(), noise_fn=sr_)
        assert rel_l2(z, z_full[r:r + 1]) < 2e-5, r
