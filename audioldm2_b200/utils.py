"""Host utilities the reference exports from ``audioldm2/__init__.py`` (utils.py:12-75): seed_everything, save_wave,
get_time, get_duration, read_list.  File output uses the standard library (``soundfile`` is not a dependency here):
16-bit PCM WAV, which is what ``sf.write(path, float32_array, samplerate)`` produces for a ``.wav`` path."""
from __future__ import annotations

import contextlib
import os
import random
import time
import wave

import numpy as np
import torch


def read_list(fname):
    """utils.py:12-18"""
    result = []
    with open(fname, "r", encoding="utf-8") as f:
        for each in f.readlines():
            result.append(each.strip("\n"))
    return result


def get_duration(fname):
    """utils.py:20-24"""
    with contextlib.closing(wave.open(fname, "r")) as f:
        return f.getnframes() / float(f.getframerate())


def get_time():
    """utils.py:33-35"""
    return time.strftime("%d_%m_%Y_%H_%M_%S", time.localtime())


def seed_everything(seed):
    """utils.py:38-49 (== pipeline.py:20-31)"""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = True


def _wav_name(name: str, i: int, many: bool) -> str:
    stem = os.path.basename(name) if ".wav" not in name else os.path.basename(name).split(".")[0]
    if many:
        return "%s_%s.wav" % (stem, i)
    fname = "%s.wav" % stem if ".wav" not in name else stem
    if len(fname) > 255:                  # avoid file names too long to be saved (utils.py:66-68)
        fname = f"{hex(hash(fname))}.wav"
    return fname


def save_wave(waveform, savepath, name="outwav", samplerate=16000):
    """utils.py:52-75: waveform [B, 1, L] float in (-1, 1) -> one PCM-16 WAV per row."""
    waveform = np.asarray(waveform)
    if type(name) is not list:
        name = [name] * waveform.shape[0]
    paths = []
    for i in range(waveform.shape[0]):
        path = os.path.join(savepath, _wav_name(name[i], i, waveform.shape[0] > 1))
        print("Save audio to %s" % path)
        pcm = np.clip(np.round(waveform[i, 0].astype(np.float64) * 32767.0), -32768, 32767).astype("<i2")
        with contextlib.closing(wave.open(path, "wb")) as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(int(samplerate))
            f.writeframes(pcm.tobytes())
        paths.append(path)
    return paths
