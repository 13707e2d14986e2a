"""Development: lane-level CPU emulation of csrc/fft2d_mma.cu (fragment tables, ldmatrix.trans addressing, mma.m16n8k16 register
layouts, the in-place tile passes) against numpy's FFT.  No GPU needed:

    nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -o /tmp/dump tools/dump_fft_frags.cu
    /tmp/dump /tmp/frag.bin && python tools/emu_fft_mma.py /tmp/frag.bin          (tests/test_fft_mma_emulation.py does exactly this)

It transcribes the kernels' index arithmetic line by line; what it proves is that the arithmetic + the PTX fragment layouts as
documented give the transform, not that the hardware was programmed correctly (tests/test_gpu_kernels.py::test_fft2 does that)."""
import sys

import numpy as np

FRAG = np.fromfile(sys.argv[1] if len(sys.argv) > 1 else "/tmp/emu/frag.bin", dtype=np.uint32).reshape(90, 32, 4)


class Cfg:
    def __init__(s, S):
        s.S = S
        s.K1 = S // 2 + 1
        s.KP = (S + 15) // 16 * 16
        s.KT = s.KP // 16
        s.MT1 = (s.K1 + 7) // 8
        s.MT2 = (S + 15) // 16
        s.K2P = (2 * s.K1 + 15) // 16 * 16
        s.KT2 = s.K2P // 16
        s.RS = s.K1 * 32 + 16
        s.ROWS = s.KP
        s.TILE = s.ROWS * s.RS + 256
        s.N_F1, s.N_G, s.N_A2 = s.MT1 * s.KT, s.MT2 * s.KT, s.MT2 * s.KT2
        s.BASE = 0 if S == 48 else 60 if S == 24 else 84
        s.O_F1 = s.BASE
        s.O_GR = s.O_F1 + s.N_F1
        s.O_GI = s.O_GR + s.N_G
        s.O_WR = s.O_GI + s.N_G
        s.O_WI = s.O_WR + s.N_G
        s.O_A2 = s.O_WI + s.N_G


def h2f(u32):
    """uint32 register -> (lo, hi) floats"""
    a = np.array([u32 & 0xFFFF, u32 >> 16], dtype=np.uint16).view(np.float16)
    return float(a[0]), float(a[1])


def pack(lo, hi):
    a = np.array([lo, hi], dtype=np.float32).astype(np.float16).view(np.uint16)
    return int(a[0]) | (int(a[1]) << 16)


class Smem:
    def __init__(s, n):
        s.b = np.zeros(n, dtype=np.uint8)

    def row16(s, addr):
        assert addr % 16 == 0 and addr + 16 <= len(s.b), addr
        return s.b[addr:addr + 16].view(np.float16).astype(np.float32)

    def st32(s, addr, v):
        assert addr % 4 == 0 and addr + 4 <= len(s.b)
        s.b[addr:addr + 4] = np.array([v], dtype=np.uint32).view(np.uint8)

    def st128(s, addr, halves8):
        assert addr % 16 == 0 and addr + 16 <= len(s.b)
        s.b[addr:addr + 16] = np.ascontiguousarray(halves8, dtype=np.float16).view(np.uint8)


def ldsm_t(sm, addrs, nmat):
    """ldmatrix.trans: addrs[lane] row addresses (lanes 0 .. 8*nmat-1 used); returns regs[lane][nmat] as (lo, hi) float pairs"""
    out = [[None] * nmat for _ in range(32)]
    for m in range(nmat):
        mat = np.stack([sm.row16(addrs[8 * m + r]) for r in range(8)])      # [row][col]
        for lane in range(32):
            out[lane][m] = (mat[2 * (lane % 4)][lane // 4], mat[2 * (lane % 4) + 1][lane // 4])
    return out


def load_b(sm, a, stride, ktn):
    """returns b[kt] as dense [16 k][8 n] matrices (assembled from the per-lane registers per the B fragment layout)"""
    regs = [[None, None] for _ in range(ktn)]  # per kt: two reg sets [lane]
    kt = 0
    per_lane = [[[None, None] for _ in range(ktn)] for _ in range(32)]
    while kt + 1 < ktn:
        o = ldsm_t(sm, [a + (kt * 16 + lane) * stride for lane in range(32)], 4)
        for lane in range(32):
            per_lane[lane][kt][0], per_lane[lane][kt][1], per_lane[lane][kt + 1][0], per_lane[lane][kt + 1][1] = o[lane]
        kt += 2
    if ktn & 1:
        o = ldsm_t(sm, [a + ((ktn - 1) * 16 + (lane & 15)) * stride for lane in range(32)], 2)
        for lane in range(32):
            per_lane[lane][ktn - 1][0], per_lane[lane][ktn - 1][1] = o[lane]
    mats = []
    for kt in range(ktn):
        B = np.zeros((16, 8), dtype=np.float32)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            B[2 * t, g], B[2 * t + 1, g] = per_lane[lane][kt][0]            # b0,b1: k = 2t, 2t+1; n = g
            B[2 * t + 8, g], B[2 * t + 9, g] = per_lane[lane][kt][1]        # b2,b3: k = 2t+8, 2t+9
        mats.append(B)
    return mats


def a_frag(idx):
    """fragment idx of the table -> dense [16][16] A tile per the A fragment layout"""
    A = np.zeros((16, 16), dtype=np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        r = FRAG[idx, lane]
        A[g, 2 * t], A[g, 2 * t + 1] = h2f(int(r[0]))
        A[g + 8, 2 * t], A[g + 8, 2 * t + 1] = h2f(int(r[1]))
        A[g, 2 * t + 8], A[g, 2 * t + 9] = h2f(int(r[2]))
        A[g + 8, 2 * t + 8], A[g + 8, 2 * t + 9] = h2f(int(r[3]))
    return A


def acc_lane(Cm, lane):
    """dense accumulator [16][8] -> the 4 values lane holds"""
    g, t = lane >> 2, lane & 3
    return Cm[g, 2 * t], Cm[g, 2 * t + 1], Cm[g + 8, 2 * t], Cm[g + 8, 2 * t + 1]


def complex_pass(c, sm, k, ar, ai, base=0, pitch=None):
    pitch = c.RS if pitch is None else pitch
    bre = load_b(sm, base + (2 * k) * 16, pitch, c.KT)
    bim = load_b(sm, base + (2 * k + 1) * 16, pitch, c.KT)
    zr, zi = [], []
    for mt in range(c.MT2):
        r = np.zeros((16, 8), np.float32)
        i = np.zeros((16, 8), np.float32)
        for kt in range(c.KT):
            r += ar[mt][kt] @ bre[kt] + ai[mt][kt] @ (-bim[kt])
            i += ar[mt][kt] @ bim[kt] + ai[mt][kt] @ bre[kt]
        zr.append(r)
        zi.append(i)
    return zr, zi


def rfft2_emu(S, x, tma=False):
    """x [S][S][8] fp16 -> spec [S][K1][16] fp16.  tma: the TMA kernel's buffers - a dense X tile (pitch 16 S, + 256 zero bytes) in front of
    the padded Y tile - instead of the in-place tile of the cp.async kernel."""
    c = Cfg(S)
    xbytes = S * S * 16 + 256 if tma else 0
    xpitch = S * 16 if tma else c.RS
    yb = xbytes
    sm = Smem(xbytes + c.TILE)
    a1 = [[a_frag(c.O_F1 + mt * c.KT + kt) for kt in range(c.KT)] for mt in range(c.MT1)]
    for i in range(S * S):
        h, w = divmod(i, S)
        sm.st128(h * xpitch + w * 16, x[h, w])
    for h in range(S):
        row = yb + h * c.RS
        b = load_b(sm, h * xpitch, 16, c.KT)
        acc = [sum(a1[mt][kt] @ b[kt] for kt in range(c.KT)) for mt in range(c.MT1)]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for mt in range(c.MT1):
                k = mt * 8 + g
                if k < c.K1:
                    v = acc_lane(acc[mt], lane)
                    sm.st32(row + (2 * k) * 16 + t * 4, pack(v[0], v[1]))
                    sm.st32(row + (2 * k + 1) * 16 + t * 4, pack(v[2], v[3]))
    ar = [[a_frag(c.O_GR + mt * c.KT + kt) for kt in range(c.KT)] for mt in range(c.MT2)]
    ai = [[a_frag(c.O_GI + mt * c.KT + kt) for kt in range(c.KT)] for mt in range(c.MT2)]
    spec = np.zeros((S, c.K1, 16), dtype=np.float16)
    for k in range(c.K1):
        zr, zi = complex_pass(c, sm, k, ar, ai, base=yb)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for mt in range(c.MT2):
                r, i = acc_lane(zr[mt], lane), acc_lane(zi[mt], lane)
                kh0, kh1 = mt * 16 + g, mt * 16 + g + 8
                o = 2 * (2 * t)
                if kh0 < S:
                    spec[kh0, k, o:o + 4] = [r[0], i[0], r[1], i[1]]
                if kh1 < S:
                    spec[kh1, k, o:o + 4] = [r[2], i[2], r[3], i[3]]
    return spec


def irfft2_emu(S, spec, add, tma=False):
    """spec [S][K1][16] fp16, add [S][S][8] fp16 -> y [S][S][8] fp16.  tma: the TMA kernels' spec buffer - dense rows of 32 K1 bytes,
    KP rows + 256 zero bytes - instead of the padded in-place tile."""
    c = Cfg(S)
    pitch = c.K1 * 32 if tma else c.RS
    sm = Smem((c.KP * pitch + 256 + 127) // 128 * 128 if tma else c.TILE)
    for i in range(S * c.K1):
        kh, k = divmod(i, c.K1)
        u = spec[kh, k]
        sm.st128(kh * pitch + k * 32, u[0::2])
        sm.st128(kh * pitch + k * 32 + 16, u[1::2])
    ar = [[a_frag(c.O_WR + mt * c.KT + kt) for kt in range(c.KT)] for mt in range(c.MT2)]
    ai = [[a_frag(c.O_WI + mt * c.KT + kt) for kt in range(c.KT)] for mt in range(c.MT2)]
    for k in range(c.K1):
        zr, zi = complex_pass(c, sm, k, ar, ai, pitch=pitch)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for mt in range(c.MT2):
                r, i = acc_lane(zr[mt], lane), acc_lane(zi[mt], lane)
                h0, h1 = mt * 16 + g, mt * 16 + g + 8
                c0 = (2 * k) * 16 + t * 4
                if h0 < S:
                    sm.st32(c0 + h0 * pitch, pack(r[0], r[1]))
                    sm.st32(c0 + h0 * pitch + 16, pack(i[0], i[1]))
                if h1 < S:
                    sm.st32(c0 + h1 * pitch, pack(r[2], r[3]))
                    sm.st32(c0 + h1 * pitch + 16, pack(i[2], i[3]))
    a2 = [[a_frag(c.O_A2 + mt * c.KT2 + kt) for kt in range(c.KT2)] for mt in range(c.MT2)]
    y = np.zeros((S, S, 8), dtype=np.float16)
    for h in range(S):
        b = load_b(sm, h * pitch, 16, c.KT2)
        for mt in range(c.MT2):
            acc = sum(a2[mt][kt] @ b[kt] for kt in range(c.KT2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                v = acc_lane(acc, lane)
                w0, w1 = mt * 16 + g, mt * 16 + g + 8
                if w0 < S:
                    y[h, w0, 2 * t:2 * t + 2] = [v[0] + float(add[h, w0, 2 * t]), v[1] + float(add[h, w0, 2 * t + 1])]
                if w1 < S:
                    y[h, w1, 2 * t:2 * t + 2] = [v[2] + float(add[h, w1, 2 * t]), v[3] + float(add[h, w1, 2 * t + 1])]
    return y


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for S in (12, 24, 48):
        K1 = S // 2 + 1
        x = rng.standard_normal((S, S, 8)).astype(np.float16)
        ref = np.fft.rfft2(x.astype(np.float64), axes=(0, 1), norm="ortho")            # [S][K1][8]
        refi = np.stack((ref.real, ref.imag), axis=-1).reshape(S, K1, 16)
        z = np.maximum(rng.standard_normal((S, K1, 16)), 0).astype(np.float16)
        add = rng.standard_normal((S, S, 8)).astype(np.float16)
        zc = z.astype(np.float64).reshape(S, K1, 8, 2)
        refy = np.fft.irfft2(zc[..., 0] + 1j * zc[..., 1], s=(S, S), axes=(0, 1), norm="ortho") + add.astype(np.float64)
        outs = {}
        for tma in (False, True):              # the cp.async kernels' in-place tile / the TMA kernels' dense buffers
            spec = rfft2_emu(S, x, tma)
            e = np.abs(spec.astype(np.float64) - refi).max() / np.abs(refi).max()
            y = irfft2_emu(S, z, add, tma)
            ei = np.abs(y.astype(np.float64) - refy).max() / np.abs(refy).max()
            print("S=%d %-8s rfft2 rel-to-max err %.2e   irfft2 %.2e" % (S, "tma" if tma else "cp.async", e, ei))
            assert e < 2e-3 and ei < 2e-3
            outs[tma] = (spec, y)
        # same arithmetic in both buffer layouts: bit-identical results
        assert np.array_equal(outs[False][0], outs[True][0]) and np.array_equal(outs[False][1], outs[True][1])
    print("ok")
