"""Golden values for the mel filterbank (A12), produced by two INDEPENDENT published restatements of
``librosa.filters.mel`` (librosa==0.9.2 itself is not installable offline): torchaudio's ``melscale_fbanks`` and
transformers' ``mel_filter_bank``, both in their documented librosa-compatible mode (Slaney scale, Slaney norm).

    python tests/golden/make_mel_golden.py      # writes tests/golden/mel_basis.pt

Stored per config: the transformers (float64) filterbank rounded to float32 and its max deviation from torchaudio's."""
import os

import numpy as np
import torch
import torchaudio
from transformers.audio_utils import mel_filter_bank

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"16k": (16000, 1024, 64, 0.0, 8000.0), "48k": (48000, 2048, 256, 20.0, 24000.0), "tiny": (4000, 256, 16, 0.0, 2000.0)}

out = {}
for name, (sr, n_fft, n_mels, fmin, fmax) in CASES.items():
    nb = 1 + n_fft // 2
    tf = mel_filter_bank(nb, n_mels, fmin, fmax, sr, norm="slaney", mel_scale="slaney").T
    ta = torchaudio.functional.melscale_fbanks(nb, fmin, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
    dev = float(np.abs(tf - ta).max())
    assert dev < 1e-6, (name, dev)
    out[name] = dict(args=torch.tensor([sr, n_fft, n_mels, fmin, fmax], dtype=torch.float64),
                     shape=torch.tensor(tf.shape), nz_index=torch.from_numpy(np.stack(np.nonzero(tf))).to(torch.int32),
                     nz_value=torch.from_numpy(tf[np.nonzero(tf)].astype(np.float32)), torchaudio_max_dev=dev)   # sparse: 2 nonzeros per bin
    print(name, tf.shape, "transformers vs torchaudio max |diff|", dev)
torch.save(out, os.path.join(HERE, "mel_basis.pt"))
