"""Engine-level C-ABI (aldm_engine_*, include/aldm_b200.h): the reference's seams as single C calls.
The same programs driven through it must reproduce the Python-orchestrated path bit for bit, and the
reference fixture within the network tolerance."""
import ctypes as C

import pytest
import torch

from audioldm2_b200 import _lib, arch
from tests.conftest import rel_l2
from tests.golden import cases
from tests.test_gpu_nets import DEV, _engine, _to

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [True, False])
def test_engine_abi_matches_python_host(use_graph):
    cfg = arch.tiny_config()
    g = cases.load("ddim_tiny")
    _, _, cond, unc = cases.unet_inputs(cfg, 2, t5_len=5)
    x_T, noises, qn = cases.sampler_noise(cfg, 2, 5, masked=False)
    nf = lambda i, kind: noises[i].to(DEV)
    outs = []
    for abi in (False, True):
        eng = _engine(cfg, 2, 5, use_engine_abi=abi, use_graph=use_graph)
        z = eng.generate_latent(_to(cond, DEV), _to(unc, DEV), ddim_steps=5, guidance=3.5, eta=1.0, x_T=x_T, noise_fn=nf)
        wav = eng.mel_spectrogram_to_waveform(eng.decode_first_stage(z)).clone()
        outs.append((z.clone(), wav))
        if abi:
            # explicit output pointers of aldm_engine_unet_eps / aldm_engine_vae_decode / aldm_engine_vocoder
            L, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
            x = x_T.to(DEV).contiguous()
            eu, ec = torch.empty_like(x), torch.empty_like(x)
            _lib.check(L.aldm_engine_unet_eps(eng._engine, x.data_ptr(), 801, eu.data_ptr(), ec.data_ptr(), st))
            e2u, e2c = eng.apply_model_pair(x, 801)
            assert torch.equal(eu, e2u) and torch.equal(ec, e2c)
            mel = torch.empty(2, 1, *eng.mel_hw, device=DEV)
            _lib.check(L.aldm_engine_vae_decode(eng._engine, z.contiguous().data_ptr(), mel.data_ptr(), st))
            w2 = torch.empty_like(wav)
            _lib.check(L.aldm_engine_vocoder(eng._engine, mel.data_ptr(), w2.data_ptr(), st))
            assert torch.equal(w2, wav)
    (z0, w0), (z1, w1) = outs
    assert torch.equal(z0, z1), f"latent differs: {rel_l2(z1, z0):.3e}"
    assert torch.equal(w0, w1), f"waveform differs: {rel_l2(w1, w0):.3e}"
    assert rel_l2(z1, g["latent"]) < 2e-4
