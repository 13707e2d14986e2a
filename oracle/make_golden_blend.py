"""Generates tests/golden/blend_golden.npz (TEST INFRASTRUCTURE; run in the build container only).

    python -m oracle.make_golden_blend

Laplacian_Pyramid_Blending_with_mask is cut out of the UNMODIFIED futils/inference_utils.py with ``ast`` (the module
imports the face-detection / face3d packages at its top and cannot be imported) and executed as it is with the real
``cv2`` and ``numpy`` in its namespace, on seeded inputs (oracle/blend.py synth_images).  Also records cv2.pyrDown /
cv2.pyrUp themselves on small edge-case shapes.
"""
from __future__ import annotations

import ast
import os

import cv2
import numpy as np

from . import blend, ref_shim, weights


def load_reference_function():
    path = os.path.join(ref_shim.REF_ROOT, "futils", "inference_utils.py")
    tree = ast.parse(open(path).read())
    ns = {"np": np, "cv2": cv2}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "Laplacian_Pyramid_Blending_with_mask":
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns["Laplacian_Pyramid_Blending_with_mask"]


def main():
    fn = load_reference_function()
    out = {"cv2_version": np.array(cv2.__version__)}
    A, B, m = blend.synth_images(64, 64, seed=0)
    out["blend64_l6"] = fn(A, B, m, 6)                              # the function's default depth
    out["blend64_l7"] = fn(A, B, m, 7)                              # down to 1 x 1, like 512 / 10 levels
    A, B, m = blend.synth_images(512, 512, seed=1)
    full = fn(A, B, m, 10)                                          # the call of inference.py:312
    out["blend512_l10_rows"] = full[::37].astype(np.float32)        # 14 of the 512 rows (the full image is 3 MB)
    out["blend512_l10_sum"] = np.array(full.astype(np.float64).sum())
    A, B, m = blend.synth_images(48, 80, seed=2)
    out["blend48x80_l4"] = fn(A, B, m, 4)
    rng = np.random.default_rng(3)
    for i, (h, w) in enumerate([(8, 8), (7, 9), (2, 2), (1, 1), (3, 1), (2, 1), (16, 24)]):
        x = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        f = rng.random((h, w), dtype=np.float32)
        out[f"down_u8_{i}_in"], out[f"down_u8_{i}"] = x, cv2.pyrDown(x).reshape((h + 1) // 2, (w + 1) // 2, 3)
        out[f"down_f32_{i}_in"], out[f"down_f32_{i}"] = f, cv2.pyrDown(f).reshape((h + 1) // 2, (w + 1) // 2)
        out[f"up_f32_{i}"] = cv2.pyrUp(f).reshape(2 * h, 2 * w)
    np.savez_compressed(os.path.join(weights._GOLDEN, "blend_golden.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith("blend")})


if __name__ == "__main__":
    main()
