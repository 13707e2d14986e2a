"""Seeded synthetic inputs for BASELINE.json's configs (TEST INFRASTRUCTURE).

Definitions follow SURVEY.md 8(d).  numpy Generators only (bit-stable across
boxes); tensors are returned as float32 torch CPU tensors.
"""
from __future__ import annotations

import numpy as np
import torch


def wav(seconds: float, seed: int = 0, sr: int = 16000) -> np.ndarray:
    """config 1/4/5 audio: default_rng(seed).standard_normal(n).astype(f32) * 0.1."""
    n = int(round(seconds * sr))
    return (np.random.default_rng(seed).standard_normal(n).astype(np.float32) * np.float32(0.1))


def lnet_inputs(batch: int, seed: int = 0):
    """config 2: mel windows N(0,1) clipped to [-4,4] [B,1,80,16]; face U(0,1) [B,6,96,96],
    channels 0-2 rows 48: zeroed (mirrors inference.py:397)."""
    rng = np.random.default_rng(seed)
    mel = np.clip(rng.standard_normal((batch, 1, 80, 16)), -4, 4).astype(np.float32)
    face = rng.random((batch, 6, 96, 96), dtype=np.float32)
    face[:, :3, 48:] = 0
    return torch.from_numpy(mel), torch.from_numpy(face)


def dnet_inputs(batch: int, seed: int = 0, t: int = 26):
    """config 3: src U(-1,1) [B,3,256,256]; coeff N(0,1) [B,73,26]."""
    rng = np.random.default_rng(seed)
    src = (rng.random((batch, 3, 256, 256), dtype=np.float32) * 2 - 1).astype(np.float32)
    coeff = rng.standard_normal((batch, 73, t)).astype(np.float32)
    return torch.from_numpy(src), torch.from_numpy(coeff)


def warp_inputs(batch: int, seed: int = 0, c: int = 3, hw: int = 256, fhw: int = 64):
    """config 3(iii): src U(-1,1), flow N(0, 3^2)."""
    rng = np.random.default_rng(seed)
    src = (rng.random((batch, c, hw, hw), dtype=np.float32) * 2 - 1).astype(np.float32)
    flow = (rng.standard_normal((batch, 2, fhw, fhw)) * 3).astype(np.float32)
    return torch.from_numpy(src), torch.from_numpy(flow)
