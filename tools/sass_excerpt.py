"""Regenerates profiles/r2_sass_trunk_excerpt.txt and profiles/r2_sass_mnemonics.txt from the built library
(cuobjdump -sass; no GPU needed): the TMA producer, the MMA issue loops and both epilogues of the kernel the default step
launches, and mnemonic counts over the whole library.   python tools/sass_excerpt.py"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "leela_b200", "libleela_b200.so")
FUN = "_ZN3lb212trunk_kernelILb1ELb1ELi2EEEvNS_11TrunkParamsE"


def sass(fun=None):
    cmd = ["cuobjdump", "-sass"] + (["-fun", fun] if fun else []) + [LIB]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    lines = []
    for l in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", l)
        if m:
            lines.append("        /*%s*/  %s ;" % (m.group(1), m.group(2).rstrip()))
    return lines


def window(lines, pred, before, after, nth=0):
    idx = [i for i, l in enumerate(lines) if pred(l)]
    if not idx:
        return ["        (not found)"]
    i = idx[min(nth, len(idx) - 1)]
    return lines[max(0, i - before): i + after]


def main():
    L = sass(FUN)
    out = []
    out.append("SASS excerpts of trunk_kernel<kPair = true, kRes = true, kOutModes = 2> — the kernel bench.py's default step launches (policy fp16,")
    out.append("value lite), from `cuobjdump -sass leela_b200/libleela_b200.so` (sm_100a, nvcc 12.9; instruction encodings removed; regenerate with")
    out.append("`python tools/sass_excerpt.py`). %d instructions in the kernel." % len(L))
    out.append("Mnemonics: UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4 (e4m3), .2CTA = cta_group::2; UTCBAR = tcgen05.commit;")
    out.append("LDTM = tcgen05.ld; UTMALDG = cp.async.bulk.tensor (TMA; .2CTA form = the peer CTA's loads signalling the leader's mbarrier); UBLKCP = cp.async.bulk;")
    out.append("SYNCS = mbarrier; F2FP...E4M3 = cvt.rn.satfinite.e4m3x2.f32; FFMA2 / FADD2 / FMUL2 = packed fp32 pairs.")
    sections = [
        ("1. TMA producer: one stage = one 3-D tensor load of the activation slab (+ bulk copy of the weight unit when the layer changes)",
         window(L, lambda l: "UTMALDG" in l, 30, 45)),
        ("2. MMA issue loop, 3x3 layer, kind::f16 slab: 9 taps x 2 row halves, the next stage's barrier probed in the middle",
         window(L, lambda l: "UTCHMMA" in l, 12, 100)),
        ("3. MMA issue loop, lite mode's correction slab: kind::f8f6f4 (e4m3, K = 32) onto the same fp32 accumulator",
         window(L, lambda l: "UTCQMMA" in l, 8, 70)),
        ("4. tcgen05.commit at the end of a stage (frees the ring slot in both CTAs) and of an item (accumulator -> epilogue)",
         window(L, lambda l: "UTCBAR" in l, 6, 12)),
        ("5. Epilogue, lite layers: TMEM loads, scale + bias (FFMA2), ELU (FMUL2, MUFU.EX2, FADD2, FMNMX), fp16 row, e4m3 rows of the activations and of the residual x 2^12, 16-byte stores with L2 hints",
         window(L, lambda l: "E4M3" in l, 75, 110)),
        ("6. Epilogue, ordinary layers: fp16 row only",
         window(L, lambda l: "F2FP.F16.F32.PACK_AB" in l, 45, 60)),
    ]
    for title, body in sections:
        out.append("")
        out.append("=" * 110)
        out.append(title)
        out.append("=" * 110)
        out += body
    open(os.path.join(ROOT, "profiles", "r2_sass_trunk_excerpt.txt"), "w").write("\n".join(out) + "\n")
    # mnemonic counts over the whole library
    allsass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cnt = collections.Counter()
    for l in allsass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            cnt[m.group(1)] += 1
    keys = [k for k in cnt if re.match(r"(UTC|LDTM|STTM|UTMA|UBLKCP|UTCBAR|SYNCS|UCGABAR|MUFU|F2FP|FFMA2|FADD2|FMUL2|LDL|STL)", k)]
    with open(os.path.join(ROOT, "profiles", "r2_sass_mnemonics.txt"), "w") as f:
        f.write("Mnemonic counts over leela_b200/libleela_b200.so (all kernel instances), `python tools/sass_excerpt.py`:\n")
        for k in sorted(keys):
            f.write("%8d  %s\n" % (cnt[k], k))
    print("wrote profiles/r2_sass_trunk_excerpt.txt (%d lines) and r2_sass_mnemonics.txt" % len(out))


if __name__ == "__main__":
    main()
