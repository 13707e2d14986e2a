"""-m gpu test of the full per-frame path (mel -> windows -> DNet -> glue -> LNet, BASELINE configs[3] shape
at reduced length) against the oracle chain, plus shard equivalence: frames computed by "rank r of 2"
equal the unsharded frames bit-for-bit (frames are independent; no float atomics anywhere)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_full_path_vs_oracle_and_shard_equivalence():
    import gpu_util as G
    from oracle import mel as omel, nets, synth, weights
    from s2v_b200 import parallel
    from s2v_b200.models.DNet import DNet
    from s2v_b200.models.LNet import LNet
    from s2v_b200.pipeline import LipSyncPipeline, glue_fake_to_face
    G.lib()
    sd_l, sd_d = weights.make_state_dict("lnet", 0), weights.make_state_dict("dnet", 0)
    lnet, dnet = LNet().cuda().eval(), DNet().cuda().eval()
    lnet.load_state_dict(sd_l, strict=True)
    dnet.load_state_dict(sd_d, strict=True)
    wav = synth.wav(0.5, seed=0)                      # 8000 samples -> T = 41 -> 9 windows/frames
    n = len(omel.mel_window_starts(1 + len(wav) // 200))
    src, coeff = synth.dnet_inputs(n, seed=2)
    src, coeff = src.cuda(), coeff.cuda()
    pipe = LipSyncPipeline(lnet, dnet, lnet_batch=4, dnet_batch=4)
    assert pipe.n_frames(len(wav)) == n
    frames = pipe.run(torch.from_numpy(wav).cuda(), src, coeff)
    assert frames.shape == (n, 3, 96, 96)
    # oracle chain (fp64 mel restatement, fp32 torch nets on the GPU)
    mel = omel.melspectrogram(wav)
    win = torch.from_numpy(omel.mel_windows(mel)).cuda()
    fake = nets.dnet_forward({k: v.cuda() for k, v in sd_d.items()}, src, coeff)["fake_image"]
    face = nets.glue_dnet_to_lnet(fake)
    ref = nets.lnet_forward({k: v.cuda() for k, v in sd_l.items()}, win, face)
    m, _ = G.report("full path frames vs oracle", frames, ref)
    p = G.psnr(frames, ref, 1.0)
    print("full path PSNR %.2f dB" % p)
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write("full path (mel->DNet->glue->LNet) PSNR %.2f dB max_abs %.5f\n" % (p, m))
    assert p >= 45.0
    # glue kernel alone vs the oracle's glue
    g = glue_fake_to_face(fake)
    assert (g - face).abs().max().item() < 1e-5
    # shard equivalence, world = 2 emulated on one GPU
    parts = []
    for r in range(2):
        lo, hi = parallel.shard_range(n, r, 2)
        parts.append(pipe.run(torch.from_numpy(wav).cuda(), src[lo:hi], coeff[lo:hi], rank=r, world=2))
    assert torch.equal(torch.cat(parts, 0), frames)


def test_pipeline_builds_driving_windows_from_semantic_table():
    """pipeline.run(semantic=...) == pipeline.run(coeffs = the oracle's per-frame transform_semantic windows), bit-for-bit,
    unsharded and as rank r of 2 (the window of a shard's first frame reaches back into the previous shard's rows)."""
    import numpy as np
    import gpu_util as G
    from oracle import mel as omel, semantic as osem, synth, weights
    from s2v_b200 import parallel
    from s2v_b200.futils import inference_utils as iu
    from s2v_b200.models.DNet import DNet
    from s2v_b200.models.LNet import LNet
    from s2v_b200.pipeline import LipSyncPipeline
    G.lib()
    lnet, dnet = LNet().cuda().eval(), DNet().cuda().eval()
    lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
    dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
    wav = torch.from_numpy(synth.wav(0.5, seed=0)).cuda()
    n = len(omel.mel_window_starts(1 + wav.numel() // 200))
    src, _ = synth.dnet_inputs(n, seed=2)
    src = src.cuda()
    table = osem.synth_table(n, seed=7)
    table[:, 80:144] *= 0.3                      # expression coefficients at a realistic scale
    table[:, 257:262] = (table[:, 257:262] - 50.0) / 100.0
    ratio = iu.find_crop_norm_ratio(table[0:1], table)
    coeff = torch.from_numpy(np.stack([osem.transform_semantic(table, i, ratio) for i in range(n)])).cuda()
    pipe = LipSyncPipeline(lnet, dnet, lnet_batch=4, dnet_batch=4)
    a = pipe.run(wav, src, coeff)
    dev_table = iu.upload_semantic(table, "cuda")
    b = pipe.run(wav, src, None, semantic=dev_table, crop_norm_ratio=ratio)
    assert torch.equal(a, b)
    parts = []
    for r in range(2):
        lo, hi = parallel.shard_range(n, r, 2)
        parts.append(pipe.run(wav, src[lo:hi], None, rank=r, world=2, semantic=dev_table, crop_norm_ratio=ratio))
    assert torch.equal(torch.cat(parts, 0), a)


def test_stream_batches_matches_direct_forward():
    """pipeline.stream_batches (copy-in / forward / copy-out on three streams, two slots in flight) returns exactly what
    the plain per-batch loop returns, including a ragged last batch and more batches than slots."""
    import gpu_util as G
    from oracle import synth, weights
    from s2v_b200.models.LNet import LNet
    from s2v_b200.pipeline import stream_batches
    G.lib()
    lnet = LNet().cuda().eval()
    lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
    sizes = [8, 8, 8, 8, 5]
    ins, outs, refs = [], [], []
    for i, b in enumerate(sizes):
        mel, face = synth.lnet_inputs(b, seed=10 + i)
        ins.append((mel.pin_memory(), face.pin_memory()))
        outs.append(torch.full((b, 3, 96, 96), float("nan")).pin_memory())
        refs.append(lnet(mel.cuda(), face.cuda()).cpu())
    for depth in (3, 2, 1):
        for o in outs:
            o.fill_(float("nan"))
        n = stream_batches(lnet, zip(ins, outs), depth=depth)
        assert n == len(sizes)
        for o, r in zip(outs, refs):
            assert torch.equal(o, r), depth
    assert stream_batches(lnet, iter(())) == 0


def test_clip_batch_structure_vs_oracle_sample():
    """The benchmarked configuration: the 60 s clip of BASELINE.json configs[3] (1 497 frames) through LipSyncPipeline with
    its real batch structure (DNet 7 x 192 + a tail of 153 on the B=160 plan, LNet 5 x 256 + a tail of 217 on the B=224 plan,
    DNet and LNet on two streams), checked on a strided sample of frames against the oracle chain (fp32, TF32
    off); overlapped == single-stream bit-for-bit; rank 3 of 8 == its slice of the unsharded run."""
    import gpu_util as G
    from oracle import mel as omel, nets, synth, weights
    from s2v_b200 import parallel
    from s2v_b200.models.DNet import DNet
    from s2v_b200.models.LNet import LNet
    from s2v_b200.pipeline import LipSyncPipeline
    G.lib()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd_l, sd_d = weights.make_state_dict("lnet", 0), weights.make_state_dict("dnet", 0)
        lnet, dnet = LNet().cuda().eval(), DNet().cuda().eval()
        lnet.load_state_dict(sd_l, strict=True)
        dnet.load_state_dict(sd_d, strict=True)
        wav_np = synth.wav(60.0, seed=0)
        n = len(omel.mel_window_starts(1 + len(wav_np) // 200))
        assert n == 1497
        src64, co64 = synth.dnet_inputs(64, seed=1)
        idx = torch.arange(n) % 64
        src, coeff = src64[idx].cuda(), co64[idx].cuda()
        wav = torch.from_numpy(wav_np).cuda()
        pipe = LipSyncPipeline(lnet, dnet)
        frames = pipe.run(wav, src, coeff)
        assert frames.shape == (n, 3, 96, 96)
        assert torch.equal(LipSyncPipeline(lnet, dnet, overlap=False).run(wav, src, coeff), frames)
        # pinned-host inputs / outputs (bench.py's e2e form) give the same frames
        out_h = torch.empty(n, 3, 96, 96).pin_memory()
        f2 = pipe.run(torch.from_numpy(wav_np).pin_memory(), src.cpu().pin_memory(), coeff.cpu().pin_memory(), out_host=out_h)
        assert torch.equal(f2, frames) and torch.equal(out_h, frames.cpu())
        # strided sample (covers first / last frames, the batch seams 191|192, 255|256, the tail batches from 1344 / 1280 and the
        # tail mel window) vs the oracle chain
        sample = sorted(set(list(range(0, n, 97)) + [63, 64, 191, 192, 255, 256, 1279, 1280, 1343, 1344, n - 2, n - 1]))
        sidx = torch.tensor(sample)
        win = torch.from_numpy(omel.mel_windows(omel.melspectrogram(wav_np)))[sidx].cuda()
        sdd, sdl = {k: v.cuda() for k, v in sd_d.items()}, {k: v.cuda() for k, v in sd_l.items()}
        fake = nets.dnet_forward(sdd, src[sidx.cuda()], coeff[sidx.cuda()])["fake_image"]
        ref = nets.lnet_forward(sdl, win, nets.glue_dnet_to_lnet(fake))
        got = frames[sidx.cuda()]
        m, _ = G.report("60 s clip, %d sampled frames vs oracle" % len(sample), got, ref)
        p = G.psnr(got, ref, 1.0)
        worst = min(G.psnr(got[i], ref[i], 1.0) for i in range(len(sample)))
        with open("gpurun_out/parity_report.txt", "a") as f:
            f.write("60 s clip real batch structure: PSNR %.2f dB (worst frame %.2f) max_abs %.5f\n" % (p, worst, m))
        assert p >= 45.0 and worst >= 45.0
        lo, hi = parallel.shard_range(n, 3, 8)
        assert torch.equal(pipe.run(wav, src[lo:hi], coeff[lo:hi], rank=3, world=8), frames[lo:hi])
        info = dnet.engine().plan_cache_info()
        assert info["plans"] <= dnet.engine().max_plans
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
