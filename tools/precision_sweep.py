"""Which layers carry the operand-rounding error? Emulates the CUDA path's arithmetic in torch (fp32 convolutions on
operands rounded exactly where the kernels round them) over the 1024-position correctness set and compares every
per-layer precision assignment with the reference's own outputs (tests/golden/bench_golden.npz). TOOLING ONLY — it
needs a GPU for speed (python tools/precision_sweep.py out.json), nothing on the product path imports it.

Per-layer codes (one character per trunk layer, i.e. per conv with c_out > 1; the final C -> 1 conv and the inner
products are always fp32, as in the kernels):
  h  one fp16 term:            conv(hi, Wh)                        hi = fp16(a), Wh = fp16(w)
  x  three fp16 terms:         conv(hi, Wh) + conv(hi, Wl) + conv(lo, Wh)     lo = fp16(a - hi), Wl = fp16(w - Wh)
  w  weights split only:       conv(hi, Wh) + conv(hi, Wl)
  a  activations split only:   conv(hi, Wh) + conv(lo, Wh)
  e  fp16 term + fp8 (e4m3) correction terms:  conv(hi, Wh) + [conv(q8(a), q8(Wl * 2^s)) + conv(q8(lo * 2^s), q8(w))] / 2^s
  f  fp32 (no rounding)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leela_b200 import synth  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def rotate_idx(v, s):
    x, y = v % 19, v // 19
    if s & 4:
        x, y = y, x
    if s & 1:
        y = 18 - y
    if s & 2:
        x = 18 - x
    return y * 19 + x


ROT = np.array([[rotate_idx(v, s) for v in range(361)] for s in range(8)])
INV = {5: 6, 6: 5}
REV = np.array([[rotate_idx(v, INV.get(s, s)) for v in range(361)] for s in range(8)])


def expand(planes, rot):
    """uint32 [n,361] + symmetry -> float [n,32,19,19] (Network.cpp:765-773)."""
    n = planes.shape[0]
    src = planes[np.arange(n)[:, None], ROT[rot]]                      # [n,361] rotated
    bits = ((src[:, None, :] >> np.arange(32, dtype=np.uint32)[None, :, None]) & 1).astype(np.float32)
    return torch.from_numpy(bits.reshape(n, 32, 19, 19)).to(DEV)


def f16(t):
    return t.half().float()


def q8(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def elu(t):
    return torch.where(t > 0, t, torch.expm1(t))


def conv(a, w, pad):
    """Exact products, fp64 accumulation (cuDNN's fp32 algorithms for 128-channel 3x3 convs are off by ~1e-4 relative:
    the first version of this sweep had a 2.3e-3 floor on the policy from that), result as the fp32 the accumulator holds."""
    return F.conv2d(a.double(), w.double(), None, padding=pad).float()


def layer(a, w, b, code, fp8_shift=12):
    pad = w.shape[-1] // 2
    if code == "f":
        return elu(conv(a, w, pad) + b.view(1, -1, 1, 1))
    hi, wh = f16(a), f16(w)
    out = conv(hi, wh, pad)
    if code in "xw":
        out = out + conv(hi, f16(w - wh), pad)
    if code in "xa":
        out = out + conv(f16(a - hi), wh, pad)
    if code == "e":
        # weights: Wl is ~2^-12 of w, w ~ 2^-4 -> scale by 2^16 into e4m3's normal range; activations as they are
        sw, sa = 2.0 ** 16, 2.0 ** fp8_shift
        c = conv(q8(a), q8((w - wh) * sw), pad) / sw
        c = c + conv(q8((a - hi) * sa), q8(w * 16.0), pad) / (sa * 16.0)
        out = out + c
    return elu(out + b.view(1, -1, 1, 1))


def forward(net, x, codes):
    a = x
    L = len(net.conv_w) - 1
    for l in range(L):
        a = layer(a, net.w[l], net.b[l], codes[l])
    return conv(a, net.w[L], 1) + net.b[L].view(1, -1, 1, 1)   # from the unrounded last trunk output, no ELU yet


class Net:
    def __init__(self, nw):
        self.conv_w = nw.conv_w
        self.w = [torch.from_numpy(w).to(DEV) for w in nw.conv_w]
        self.b = [torch.from_numpy(b).to(DEV) for b in nw.conv_b]
        self.ip_w = [torch.from_numpy(w).to(DEV) for w in nw.ip_w]
        self.ip_b = [torch.from_numpy(b).to(DEV) for b in nw.ip_b]


def policy(net, planes, rot, temp, codes, chunk=256):
    out = []
    for lo in range(0, planes.shape[0], chunk):
        z = elu(forward(net, expand(planes[lo:lo + chunk], rot[lo:lo + chunk]), codes)).reshape(-1, 361)
        p = torch.softmax(z.double() / temp, 1).cpu().numpy()
        out.append(p[np.arange(p.shape[0])[:, None], REV[rot[lo:lo + chunk]]])
    return np.concatenate(out)


def value(net, planes, rot, codes, chunk=256):
    out = []
    for lo in range(0, planes.shape[0], chunk):
        v = elu(forward(net, expand(planes[lo:lo + chunk], rot[lo:lo + chunk]), codes)).reshape(-1, 361)
        h = elu(v @ net.ip_w[0].t() + net.ip_b[0])
        o = h @ net.ip_w[1].t() + net.ip_b[1]
        out.append(((1.0 + torch.tanh(o.double())) * 0.5).reshape(-1).cpu().numpy())
    return np.concatenate(out)


def stats(d):
    return {"max": float(d.max()), "p99.9": float(np.quantile(d, 0.999)), "mean": float(d.mean())}


def main():
    b = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "bench_golden.npz"))
    pp, vp, rot = b["policy_planes"], b["value_planes"], b["rotation"].astype(np.int64)
    temp = float(g["softmax_temp"])
    want_p, want_v = g["policy"].astype(np.float64), g["value"].astype(np.float64)
    pn, vn = Net(synth.policy_weights()), Net(synth.value_weights())
    pn1 = Net(synth.policy_weights(gain=1.0))
    LP, LV = 12, 11
    res = {"note": __doc__.split("\n")[0], "positions": int(pp.shape[0]), "policy": {}, "value": {}, "policy_gain1": {}}

    def pol(name, codes, net=pn, ref=want_p, key="policy"):
        p = policy(net, pp, rot, temp, codes)
        d = np.abs(p - ref)
        top = int((p.argmax(1) == ref.argmax(1)).sum())
        res[key][name] = dict(stats(d), codes=codes, top1_same=top, frac_pos_within_1e3=float((d.max(1) < 1e-3).mean()),
                              terms=sum({"h": 1, "x": 3, "w": 2, "a": 2, "e": 2, "f": 0}[c] * 1.0 for c in codes) / len(codes))
        print(key, name, codes, res[key][name], flush=True)

    def val(name, codes):
        v = value(vn, vp, rot, codes)
        d = np.abs(v - want_v)
        res["value"][name] = dict(stats(d), codes=codes, frac_within_1e3=float((d < 1e-3).mean()))
        print("value", name, codes, res["value"][name], flush=True)

    # sanity: the fp32 emulation reproduces the reference
    pol("fp32", "f" * LP)
    val("fp32", "f" * LV)
    for c, nm in (("h", "fp16 (default)"), ("x", "3 terms everywhere"), ("w", "weights split"), ("a", "activations split"),
                  ("e", "fp16 + fp8 corrections")):
        pol(nm, c * LP)
        val(nm, c * LV)
    # the last k layers precise / the first k layers precise / exactly one layer left in fp16
    for k in (2, 4, 6, 8, 10):
        pol(f"last {k} x", "h" * (LP - k) + "x" * k)
        pol(f"first {k} x", "x" * k + "h" * (LP - k))
    for k in (2, 4, 6, 8, 10):
        val(f"last {k} x", "h" * (LV - k) + "x" * k)
        val(f"first {k} x", "x" * k + "h" * (LV - k))
    for l in range(LP):
        pol(f"only layer {l + 1} fp16", "x" * l + "h" + "x" * (LP - 1 - l))
    for l in range(LV):
        val(f"only layer {l + 1} fp16", "x" * l + "h" + "x" * (LV - 1 - l))
    # mixed: fp8-corrected value with 3-term first layer, etc.
    val("x first, e rest", "x" + "e" * (LV - 1))
    val("w first, e rest", "w" + "e" * (LV - 1))
    pol("w first, e rest", "w" + "e" * (LP - 1))
    for k in (2, 4, 6, 8, 10):
        pol(f"last {k} e", "h" * (LP - k) + "e" * k)
    # gain 1 beside gain 2 (no reference golden for gain 1: the fp32 emulation, validated above, is the yardstick)
    ref1 = policy(pn1, pp, rot, temp, "f" * LP)
    for c, nm in (("h", "fp16 (default)"), ("x", "3 terms everywhere"), ("e", "fp16 + fp8 corrections")):
        pol(nm, c * LP, net=pn1, ref=ref1, key="policy_gain1")
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
