"""Development tool: where the overlapped host loop loses time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2v_b200
from oracle import synth, weights
from s2v_b200.models.LNet import LNet
from s2v_b200 import pipeline
dev = torch.device("cuda", 0)
net = LNet().to(dev).eval(); net.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
B, K = 128, 20
mel, face = synth.lnet_inputs(B, seed=0)
mel_h, face_h = mel.pin_memory(), face.pin_memory()
out_h = torch.empty(B, 3, 96, 96).pin_memory()
mel_d, face_d = mel.to(dev), face.to(dev)


def timed(name, fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-50s %.2f ms/step" % (name, dt * 1e3 / K), flush=True)


with torch.no_grad():
    timed("device-resident forward", lambda: [net(mel_d, face_d) for _ in range(K)])
    timed("sequential h2d+fwd+d2h", lambda: [out_h.copy_(net(mel_h.to(dev, non_blocking=True), face_h.to(dev, non_blocking=True)), non_blocking=True) for _ in range(K)])
    gen = lambda: (((mel_h, face_h), out_h) for _ in range(K))
    timed("stream_batches depth 2", lambda: pipeline.stream_batches(net, gen()))
    timed("stream_batches depth 3", lambda: pipeline.stream_batches(net, gen(), depth=3))
    # forward on a side stream while another stream does H2D copies in a loop
    s2 = torch.cuda.Stream()
    buf = torch.empty_like(face_d)

    def fwd_with_bg_copy():
        for _ in range(K):
            with torch.cuda.stream(s2):
                buf.copy_(face_h, non_blocking=True)
            net(mel_d, face_d)
    timed("device forward + concurrent H2D on stream 2", fwd_with_bg_copy)
    o2 = torch.empty(B, 3, 96, 96, device=dev)

    def fwd_with_bg_d2h():
        for _ in range(K):
            with torch.cuda.stream(s2):
                out_h.copy_(o2, non_blocking=True)
            net(mel_d, face_d)
    timed("device forward + concurrent D2H on stream 2", fwd_with_bg_d2h)

    # variants: which overlap stalls?
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()
    bufs = [(torch.empty_like(mel_d), torch.empty_like(face_d)) for _ in range(3)]
    outs = [torch.empty(B, 3, 96, 96, device=dev) for _ in range(3)]
    ev_in = [torch.cuda.Event() for _ in range(3)]
    ev_run = [torch.cuda.Event() for _ in range(3)]
    ev_out = [torch.cuda.Event() for _ in range(3)]

    def v_h2d_only():          # H2D on its own stream (prefetch one ahead), D2H on the compute stream
        for i in range(K + 1):
            if i < K:
                s = i % 3
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_run[s])
                    bufs[s][0].copy_(mel_h, non_blocking=True); bufs[s][1].copy_(face_h, non_blocking=True)
                    ev_in[s].record(s_in)
            if i > 0:
                s = (i - 1) % 3
                cur.wait_event(ev_in[s])
                r = net(*bufs[s])
                ev_run[s].record(cur)
                out_h.copy_(r, non_blocking=True)
    timed("H2D prefetch stream + D2H inline", v_h2d_only)

    def v_d2h_only():          # H2D inline, D2H on its own stream
        for i in range(K):
            s = i % 3
            cur.wait_event(ev_out[s])
            r = net(mel_h.to(dev, non_blocking=True), face_h.to(dev, non_blocking=True))
            outs[s].copy_(r)
            ev_run[s].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run[s])
                out_h.copy_(outs[s], non_blocking=True)
                ev_out[s].record(s_out)
        s_out.synchronize()
    timed("H2D inline + D2H stream", v_d2h_only)
    timed("stream_batches depth 3 (again)", lambda: pipeline.stream_batches(net, gen(), depth=3))
    K = 50
    gen = lambda: (((mel_h, face_h), out_h) for _ in range(K))
    timed("stream_batches depth 3, K=50", lambda: pipeline.stream_batches(net, gen(), depth=3))
    timed("sequential, K=50", lambda: [out_h.copy_(net(mel_h.to(dev, non_blocking=True), face_h.to(dev, non_blocking=True)), non_blocking=True) for _ in range(K)])
