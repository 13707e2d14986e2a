"""Development / measurement tool: Laplacian-pyramid blend throughput (512 x 512, 10 levels = inference.py:312) on the GPU
against the reference's own implementation (cv2 on the host cores) in the same run."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import s2v_b200
from oracle import blend as ob
from s2v_b200.futils import inference_utils as iu

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A, B, m = ob.synth_images(512, 512, seed=1, n=N)
a, b, mm = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(m).cuda()
for _ in range(3):
    out = iu.laplacian_blend(a, b, mm, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    out = iu.laplacian_blend(a, b, mm, 10)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# algorithmic bytes per frame: A, B, mask in; blended frame out (pyramids are intermediates)
alg = 512 * 512 * (3 + 3 + 4 + 12)
line = {"workload": "Laplacian_Pyramid_Blending_with_mask, 512x512x3 uint8 + float32 mask, 10 levels, batch %d" % N,
        "ms": round(ms, 3), "frames_per_s": round(N / ms * 1e3, 1), "GBps_algorithmic": round(N * alg / ms / 1e6, 1),
        "alg_bytes_per_frame": alg}
try:
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)

    def ref(A, B, m, num_levels):             # restated call sequence of inference_utils.py:181-222 on the real cv2
        GA, GB, GM = A.copy(), B.copy(), m.copy()
        gpA, gpB, gpM = [GA], [GB], [GM]
        for i in range(num_levels):
            GA, GB, GM = cv2.pyrDown(GA), cv2.pyrDown(GB), cv2.pyrDown(GM)
            gpA.append(np.float32(GA)); gpB.append(np.float32(GB)); gpM.append(np.float32(GM))
        lpA, lpB, gpMr = [gpA[num_levels - 1]], [gpB[num_levels - 1]], [gpM[num_levels - 1]]
        for i in range(num_levels - 1, 0, -1):
            lpA.append(np.subtract(gpA[i - 1], cv2.pyrUp(gpA[i]))); lpB.append(np.subtract(gpB[i - 1], cv2.pyrUp(gpB[i])))
            gpMr.append(gpM[i - 1])
        LS = [la * gm[:, :, None] + lb * (1.0 - gm[:, :, None]) for la, lb, gm in zip(lpA, lpB, gpMr)]
        ls_ = LS[0]
        for i in range(1, num_levels):
            ls_ = cv2.add(cv2.pyrUp(ls_), LS[i])
        return ls_
    k = min(N, 16)
    ref(A[0], B[0], m[0], 10)
    t0 = time.perf_counter()
    outs = [ref(A[i], B[i], m[i], 10) for i in range(k)]
    dt = time.perf_counter() - t0
    line["cpu_cv2_frames_per_s"] = round(k / dt, 1)
    line["cpu_cores"] = os.cpu_count()
    line["max_abs_vs_cv2"] = float(max(np.abs(out[i].cpu().numpy() - outs[i]).max() for i in range(k)))
except ImportError:
    line["cpu_cv2_frames_per_s"] = None
print(json.dumps(line))
