"""Development tool: per-op CUDA-event timing of the DNet plan at B=64 (eager replay)."""
import sys, os, re, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from oracle import synth, weights
from s2v_b200.models.DNet import DNet
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = DNet().to(dev).eval()
net.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
src, coeff = synth.dnet_inputs(B, seed=0)
src, coeff = src.to(dev), coeff.to(dev)
for _ in range(3):
    net(src, coeff)
eng = net.engine()
ent = eng._plans[(B, 26, "full")]
ops_l = ent["plan"].ops
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ops_l]
acc = {}
for r in range(3):
    torch.cuda.synchronize()
    for op, (a, b) in zip(ops_l, evs):
        a.record(); op.run(); b.record()
    torch.cuda.synchronize()
    if r == 0:
        continue
    for op, (a, b) in zip(ops_l, evs):
        key = re.sub(r"\.ph\d\d", ".ph*", op.name)
        key = re.sub(r"res(\d)\.res\d", r"res\1.*", key)
        d = acc.setdefault(key, [0.0, 0, 0.0])
        d[0] += a.elapsed_time(b) / 2; d[1] += 1 if r == 1 else 0; d[2] += getattr(op, "alg_flops", 0.0) if r == 1 else 0.0
if os.environ.get("BD_AFFINE", "0") == "1":
    rows = []
    for op, (a, b) in zip(ops_l, evs):
        if op.name == "affine_act":
            ms = a.elapsed_time(b)
            rows.append((ms, op.alg_bytes))
    print("affine_act launches (last pass): total %.3f ms, %.1f MB -> %.0f GB/s overall" % (
        sum(r[0] for r in rows), sum(r[1] for r in rows) / 1e6, sum(r[1] for r in rows) / sum(r[0] for r in rows) / 1e6))
    for ms, by in sorted(rows, reverse=True)[:12]:
        print("   %8.1f us  %8.1f MB  %6.0f GB/s" % (ms * 1e3, by / 1e6, by / ms / 1e6))
tot = sum(v[0] for v in acc.values())
print("sum of ops %.3f ms (B=%d), %d launches" % (tot, B, len(ops_l)))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-62s %7.3f ms %3d x %8.1f us %s" % (k, v[0], v[1], 1e3 * v[0] / v[1], ("%.0f TFLOP/s" % (v[2] / v[0] / 1e9)) if v[2] else ""))
