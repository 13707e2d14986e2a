#!/usr/bin/env python
"""SASS opcode histogram per object of libs2v.so (profiles/r2_sass_opcodes.txt): static proof of which hardware paths the
kernels use (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA, ...).  Runs without a GPU.

    python tools/sass_hist.py > profiles/r2_sass_opcodes.txt
"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "UTCBAR", "UTCATOMSWS", "SYNCS",
        "UCGABAR_WAIT", "UCGABAR_ARV", "HMMA", "LDSM", "FFMA", "HFMA2", "DFMA", "DADD", "DMUL", "MUFU", "F2FP", "LDG", "STG", "LDS", "STS",
        "SHFL", "REDUX", "ATOMS", "ATOMG", "RED")
print("# SASS opcode histogram per object of libs2v.so (round 2, final build; tools/sass_hist.py)")
print("# cuobjdump -sass speech-to-video-mpp_b200/build/<obj>.o | opcode (modifiers stripped) | count   -- static instruction counts (loops counted once)")
print("# UTCHMMA = tcgen05.mma (kind::f16, incl. .2CTA pair form), LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAPF = TMA tensor load / store / L2 prefetch,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA + LDSM = mma.sync (the tensor-core DFT of fft2d_mma and the cross-check attention kernel), DFMA/DADD/DMUL = float64 statistics / resampler / cv2-double taps")
for obj in sorted(glob.glob(os.path.join(ROOT, "speech-to-video-mpp_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", out, re.M):
        if m.group(1) in KEEP:
            cnt[m.group(1)] += 1
    two_cta = len(re.findall(r"UTCHMMA\.2CTA", out))
    print("== " + os.path.basename(obj)[:-2])
    print(" ".join("%s:%d" % kv for kv in cnt.most_common()) + ("  (UTCHMMA.2CTA:%d)" % two_cta if two_cta else ""))
