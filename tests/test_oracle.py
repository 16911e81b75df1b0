"""Pin the CPU oracle (oracle/functional.py, oracle/mel.py) against the committed golden
fixtures, which hold outputs of the UNMODIFIED reference modules (tests/golden/make_golden.py).
Tolerance: the oracle restates the same fp32 torch ops in a different association order, so
agreement is at fp32 round-off: relative L2 <= 2e-5 (1e-4 after a multi-step sampler run)."""
import numpy as np
import pytest
import torch

from audioldm2_b200 import arch, synth
from oracle import functional as OF
from oracle import mel as OM
from tests.conftest import rel_l2
from tests.golden import cases

TOL = 2e-5


@pytest.mark.parametrize("name,film,t5,variant", [("unet_tiny", False, 5, ""), ("unet_tiny_film", True, 32, ""),
                                                  ("unet_tiny_large", False, 5, "large"), ("unet_tiny_48k", False, 32, "48k")])
def test_unet_tiny(name, film, t5, variant):
    cfg = arch.tiny_config(film=film, variant=variant)
    g = cases.load(name)
    sd = synth.unet_state_dict(cfg["unet"])
    x, t, cond, unc = cases.unet_inputs(cfg, 2, t5_len=t5)
    with torch.no_grad():
        for tag, c in (("cond", cond), ("uncond", unc)):
            y = OF.unet_forward(sd, cfg["unet"], x, t, c["context_list"], c["mask_list"], c["y"])
            assert rel_l2(y, g["eps_" + tag]) < TOL


def test_vae_tiny():
    cfg = arch.tiny_config()
    g = cases.load("vae_tiny")
    sd = synth.vae_state_dict(cfg["vae"])
    with torch.no_grad():
        mel = OF.vae_decode(sd, cfg["vae"], cases.latent(cfg, 2, seed=5))
        mom = OF.vae_encode_moments(sd, cfg["vae"], cases.mel_input(cfg, 2))
    assert rel_l2(mel, g["mel"]) < TOL
    assert rel_l2(mom, g["moments"]) < TOL


def test_first_stage_tiny_48k_topology():
    cfg = arch.tiny_config(variant="48k")
    g = cases.load("vae_tiny_48k")
    sd = synth.vae_state_dict(cfg["vae"])
    with torch.no_grad():
        assert rel_l2(OF.vae_decode(sd, cfg["vae"], cases.latent(cfg, 2, seed=5)), g["mel"]) < TOL
        assert rel_l2(OF.vae_encode_moments(sd, cfg["vae"], cases.mel_input(cfg, 2)), g["moments"]) < TOL
        gv = cases.load("vocoder_tiny_48k")
        w = OF.vocoder_forward(synth.vocoder_state_dict(cfg["vocoder"]), cfg["vocoder"], cases.vocoder_input(cfg, 2, 16))
    assert rel_l2(w, gv["wave"]) < TOL


def test_vocoder_tiny():
    cfg = arch.tiny_config()
    g = cases.load("vocoder_tiny")
    sd = synth.vocoder_state_dict(cfg["vocoder"])
    with torch.no_grad():
        w = OF.vocoder_forward(sd, cfg["vocoder"], cases.vocoder_input(cfg, 2, 24))
    assert w.shape[-1] == arch.vocoder_out_len(cfg["vocoder"], 24)
    assert rel_l2(w, g["wave"]) < TOL
    assert float(w.abs().max()) < 0.999      # tanh not saturated: the test is sensitive


@pytest.mark.parametrize("masked", [False, True])
def test_ddim_tiny(masked):
    cfg = arch.tiny_config()
    g = cases.load("ddim_tiny_masked" if masked else "ddim_tiny")
    sd = synth.unet_state_dict(cfg["unet"])
    _, _, cond, unc = cases.unet_inputs(cfg, 2, t5_len=5)
    x_T, noises, qn = cases.sampler_noise(cfg, 2, 5, masked=masked)
    mask = x0 = None
    if masked:
        mask, x0 = cases.inpaint_mask(cfg, 2)
    tables = OF.ddpm_tables(cfg["linear_start"], cfg["linear_end"], cfg["timesteps"])
    with torch.no_grad():
        z = OF.ddim_sample(sd, cfg["unet"], x_T, noises, cond, unc, 5, 1.0, 3.5, tables, mask, x0, qn)
    assert rel_l2(z, g["latent"]) < 1e-4


def test_ddim_schedule_values():
    """make_schedule facts from SURVEY.md 8(a') item 11: t in {1,6,...,996} for S=200."""
    st = OF.ddim_schedule(OF.ddpm_tables(), 200, 1.0)
    assert [s["t"] for s in st[:2]] == [996, 991] and st[-1]["t"] == 1
    assert st[0]["index"] == 199 and st[-1]["index"] == 0
    assert all(0 < s["a_t"] < 1 and s["sigma_t"] >= 0 for s in st)


@pytest.mark.parametrize("name,args,n", [("stft_tiny", (256, 40, 16, 4000, 0, 2000), 4000),
                                         ("stft_16k", (1024, 160, 64, 16000, 0, 8000), 163840)])
def test_stft_mel(name, args, n):
    g = cases.load(name)
    wav = cases.wav_input(n).numpy()
    logmel, mag = OM.stft_mel(wav, *args)
    assert logmel.shape == tuple(g["logmel"].shape)
    # the reference runs its DFT as an fp32 conv1d; the oracle uses a float64 FFT
    assert rel_l2(torch.from_numpy(logmel), g["logmel"]) < 1e-4
    assert abs(float(np.linalg.norm(mag)) / float(g["mag_l2"]) - 1) < 1e-5
    # independent check of the framing/window against torch.stft
    n_fft, hop = args[0], args[1]
    ts = torch.stft(torch.from_numpy(wav), n_fft, hop, n_fft, window=torch.hann_window(n_fft),
                    center=True, pad_mode="reflect", return_complex=True).abs()
    assert rel_l2(torch.from_numpy(mag), ts) < 1e-5


def test_mel_filterbank_properties():
    """Parity unpinned against librosa (absent); check the published properties instead:
    triangular, non-negative, Slaney area normalisation (each filter integrates to ~1 in Hz
    units -> sum(w) * bin_hz ~= 1 for filters well inside the band)."""
    sr, n_fft, n_mels = 16000, 1024, 64
    w = OM.mel_filterbank(sr, n_fft, n_mels, 0, 8000)
    assert w.shape == (64, 513) and w.dtype == np.float32 and (w >= 0).all()
    area = w.sum(axis=1) * (sr / n_fft)
    assert np.allclose(area[5:-1], 1.0, atol=0.15)
    peaks = w.argmax(axis=1)
    assert (np.diff(peaks) > 0).all()


@pytest.mark.slow
def test_full_size_against_reference_fixtures():
    cfg = arch.model_config("audioldm2-full")
    with torch.no_grad():
        g = cases.load("unet_full")
        sd = synth.unet_state_dict(cfg["unet"])
        x, t, cond, unc = cases.unet_inputs(cfg, 1)
        y = OF.unet_forward(sd, cfg["unet"], x, t, cond["context_list"], cond["mask_list"], None)
        assert rel_l2(y, g["eps_cond"]) < TOL
        g = cases.load("vae_full")
        sd = synth.vae_state_dict(cfg["vae"])
        assert rel_l2(OF.vae_decode(sd, cfg["vae"], cases.latent(cfg, 1, seed=5)), g["mel"]) < TOL
        g = cases.load("vocoder_full")
        sd = synth.vocoder_state_dict(cfg["vocoder"])
        w = OF.vocoder_forward(sd, cfg["vocoder"], cases.vocoder_input(cfg, 1, 1024))
        assert w.shape[-1] == 163872
        assert rel_l2(w, g["wave"]) < TOL
