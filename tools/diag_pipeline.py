"""Development tool: where the full-path time goes (device events vs host wall clock per stage)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from oracle import synth, weights
from s2v_b200.futils import audio
from s2v_b200.models.DNet import DNet
from s2v_b200.models.LNet import LNet
from s2v_b200.pipeline import LipSyncPipeline, glue_fake_to_face
dev = torch.device("cuda", 0)
lnet = LNet().to(dev).eval(); lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
dnet = DNet().to(dev).eval(); dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
wav = torch.from_numpy(synth.wav(60.0, seed=0)).to(dev)
n = 1497
srcs, coeffs = synth.dnet_inputs(64, seed=1)
srcs = srcs.to(dev).repeat(24, 1, 1, 1)[:n].contiguous(); coeffs = coeffs.to(dev).repeat(24, 1, 1)[:n].contiguous()
pipe = LipSyncPipeline(lnet, dnet)
pipe.run(wav, srcs, coeffs); torch.cuda.synchronize()


def timed(name, fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(reps):
        fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%-44s device %.1f ms   host-issue %.1f ms   wall %.1f ms" % (name, a.elapsed_time(b) / reps, (t1 - t0) * 1e3 / reps, (t2 - t0) * 1e3 / reps), flush=True)


with torch.no_grad():
    timed("pipeline.run", lambda: pipe.run(wav, srcs, coeffs))
    timed("mel + windows", lambda: audio.mel_windows(audio.melspectrogram_device(wav), 25.0, 0, n))
    win = audio.mel_windows(audio.melspectrogram_device(wav), 25.0, 0, n)
    faces = torch.empty(n, 6, 96, 96, device=dev)

    def dloop():
        for s in range(0, n, 64):
            e = min(n, s + 64)
            faces[s:e] = glue_fake_to_face(dnet(srcs[s:e], coeffs[s:e])["fake_image"])
    timed("DNet loop (24 calls) + glue", dloop)

    def dloop_same():
        for s in range(0, 23):
            dnet(srcs[:64], coeffs[:64])
    timed("DNet 23 x B=64 same input, no glue", dloop_same)
    eng = dnet.engine()
    ent = eng._plans[(64, 26, "full")]
    timed("DNet 23 x raw engine replay", lambda: [eng._run(ent) for _ in range(23)])

    def lloop():
        for s in range(0, n, 128):
            e = min(n, s + 128)
            lnet(win[s:e], faces[s:e])
    timed("LNet loop (12 calls)", lloop)
    le = lnet.engine(); lent = le.plan_for(128)
    timed("LNet 12 x raw engine replay", lambda: [le._run(lent) for _ in range(12)])
    timed("DNet B=25 tail", lambda: dnet(srcs[:25], coeffs[:25]))
    timed("LNet B=89 tail", lambda: lnet(win[:89], faces[:89]))
