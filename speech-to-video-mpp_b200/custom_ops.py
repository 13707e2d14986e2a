"""torch custom ops (``torch.ops.s2v.*``) over the C ABI - the form BASELINE.json's north_star asks for:
host code stays Python/PyTorch and reaches the sm_100a kernels as custom ops through the thin ctypes layer.

Stateless kernels are plain ops; the two networks are ops that take an integer engine handle (the packed
weights + per-batch plans live in the Python-side engine object registered under that handle).
"""
from __future__ import annotations

import weakref

import torch

from . import pipeline as _pipeline
from .futils import audio as _audio
from .futils import flow_util as _flow

# handle -> engine, held WEAKLY: the nn.Module owns its engine (packed weights, plans, workspaces); when the module is
# deleted or rebuilt the engine and its device memory go with it
_ENGINES: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_NEXT = [1]


def register_engine(engine) -> int:
    h = _NEXT[0]
    _NEXT[0] += 1
    _ENGINES[h] = engine
    return h


def release_engine(handle: int) -> None:
    _ENGINES.pop(handle, None)


@torch.library.custom_op("s2v::flow_warp", mutates_args=())
def flow_warp(source: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    return _flow.warp_flow(source, flow)


@flow_warp.register_fake
def _(source, flow):
    return torch.empty_like(source, dtype=torch.float32)


@torch.library.custom_op("s2v::melspectrogram", mutates_args=())
def melspectrogram(wav: torch.Tensor) -> torch.Tensor:
    return _audio.melspectrogram_device(wav)


@melspectrogram.register_fake
def _(wav):
    return wav.new_empty((80, 1 + wav.numel() // 200), dtype=torch.float32)


@torch.library.custom_op("s2v::mel_windows", mutates_args=())
def mel_windows(mel: torch.Tensor, fps: float, first: int, count: int) -> torch.Tensor:
    return _audio.mel_windows(mel, fps, first, count)


@mel_windows.register_fake
def _(mel, fps, first, count):
    return mel.new_empty((count, 1, 80, 16))


@torch.library.custom_op("s2v::glue_fake_to_face", mutates_args=())
def glue_fake_to_face(fake: torch.Tensor, size: int) -> torch.Tensor:
    return _pipeline.glue_fake_to_face(fake, size)


@glue_fake_to_face.register_fake
def _(fake, size):
    return fake.new_empty((fake.shape[0], 2 * fake.shape[1], size, size), dtype=torch.float32)


@torch.library.custom_op("s2v::lnet_forward", mutates_args=())
def lnet_forward(mel: torch.Tensor, face: torch.Tensor, handle: int) -> torch.Tensor:
    return _ENGINES[handle].forward(mel, face)


@lnet_forward.register_fake
def _(mel, face, handle):
    return face.new_empty((face.shape[0], 3, face.shape[2], face.shape[3]), dtype=torch.float32)


@torch.library.custom_op("s2v::dnet_forward", mutates_args=())
def dnet_forward(image: torch.Tensor, coeff: torch.Tensor, warp_only: bool, handle: int) -> list[torch.Tensor]:
    out = _ENGINES[handle].forward(image, coeff, "warp" if warp_only else None)
    res = [out["flow_field"], out["warp_image"]]
    if not warp_only:
        res.append(out["fake_image"])
    return res


@dnet_forward.register_fake
def _(image, coeff, warp_only, handle):
    b = image.shape[0]
    res = [image.new_empty((b, 2, 64, 64), dtype=torch.float32), torch.empty_like(image, dtype=torch.float32)]
    if not warp_only:
        res.append(torch.empty_like(image, dtype=torch.float32))
    return res
