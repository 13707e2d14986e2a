"""Generates tests/golden/semantic_golden.npz (TEST INFRASTRUCTURE; run in the build container only).

    python -m oracle.make_golden_semantic

futils/inference_utils.py of the reference cannot be imported here (its module top imports cv2, torchvision, the
face-detection and face3d packages), so the three functions on the path - obtain_seq_index (:73-76),
transform_semantic (:78-91), find_crop_norm_ratio (:93-99) - are cut out of the UNMODIFIED source file with ``ast``
and executed as they are, with only ``np`` and ``torch`` in their namespace.  Their outputs on seeded synthetic
coefficient tables (oracle/semantic.py synth_table) are the fixtures that pin the oracle and the CUDA kernel.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

from . import ref_shim, semantic, weights

NAMES = ("obtain_seq_index", "transform_semantic", "find_crop_norm_ratio")


def load_reference_functions():
    path = os.path.join(ref_shim.REF_ROOT, "futils", "inference_utils.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in NAMES:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in NAMES]


def main():
    _, transform_semantic, find_crop_norm_ratio = load_reference_functions()
    out = {}
    for tag, dtype, n in (("f32", np.float32, 40), ("f64", np.float64, 17)):
        table = semantic.synth_table(n, seed=3, dtype=dtype)
        ratio = find_crop_norm_ratio(table[0:1], table[1:])           # source = frame 0, targets = the rest
        frames = np.array([0, 1, 5, 12, 13, 14, n // 2, n - 14, n - 13, n - 2, n - 1], dtype=np.int32)
        out[tag + "_ratio"] = np.asarray(ratio)
        out[tag + "_frames"] = frames
        out[tag + "_plain"] = np.stack([transform_semantic(table, int(i)).numpy() for i in frames])
        out[tag + "_scaled"] = np.stack([transform_semantic(table, int(i), ratio).numpy() for i in frames])
        out[tag + "_zero_ratio"] = transform_semantic(table, 3, np.zeros(1, dtype)).numpy()
    np.savez_compressed(os.path.join(weights._GOLDEN, "semantic_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
