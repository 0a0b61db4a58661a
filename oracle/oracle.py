"""ctypes front-end of oracle/liboracle.so (oracle/leela_oracle.c). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import concurrent.futures as cf
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

ROUND_W, ROUND_ACT, ROUND_LAST = 1, 2, 4
P = 361

_f32p = C.POINTER(C.c_float)


class _Net(C.Structure):
    _fields_ = [("n_conv", C.c_int), ("k", C.c_int * 16), ("c_in", C.c_int * 16), ("c_out", C.c_int * 16),
                ("w", _f32p * 16), ("b", _f32p * 16),
                ("n_ip", C.c_int), ("ip_in", C.c_int * 4), ("ip_out", C.c_int * 4),
                ("ip_w", _f32p * 4), ("ip_b", _f32p * 4)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "leela_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.lb2o_rotate_nn_idx.restype = C.c_int
        L.lb2o_rev_rotate_nn_idx.restype = C.c_int
        L.lb2o_value_forward.restype = C.c_float
        _lib = L
    return _lib


def _ptr(a: np.ndarray, ty=_f32p):
    return a.ctypes.data_as(ty)


class OracleNet:
    """Holds a lb2o_net struct plus the numpy arrays that back its pointers."""

    def __init__(self, weights):
        self._keep = []
        n = _Net()
        n.n_conv = len(weights.convs)
        for i, c in enumerate(weights.convs):
            n.k[i], n.c_in[i], n.c_out[i] = c.k, c.c_in, c.c_out
            w = np.ascontiguousarray(weights.conv_w[i], dtype=np.float32)
            b = np.ascontiguousarray(weights.conv_b[i], dtype=np.float32)
            self._keep += [w, b]
            n.w[i], n.b[i] = _ptr(w), _ptr(b)
        n.n_ip = len(weights.ips)
        for j, ip in enumerate(weights.ips):
            n.ip_in[j], n.ip_out[j] = ip.n_in, ip.n_out
            w = np.ascontiguousarray(weights.ip_w[j], dtype=np.float32)
            b = np.ascontiguousarray(weights.ip_b[j], dtype=np.float32)
            self._keep += [w, b]
            n.ip_w[j], n.ip_b[j] = _ptr(w), _ptr(b)
        self.struct = n
        self.weights = weights


def _threads() -> int:
    return max(1, min(os.cpu_count() or 1, 32))


def policy_forward(net: OracleNet, planes: np.ndarray, rotation: np.ndarray, temperature: float = 0.75,
                   emulate: int = 0, want_logits: bool = False):
    """planes uint32 [n,361], rotation uint8 [n] -> probs float32 [n,361] (un-rotated, all points)."""
    L = lib()
    planes = np.ascontiguousarray(planes, dtype=np.uint32)
    rotation = np.ascontiguousarray(rotation, dtype=np.uint8)
    n = planes.shape[0]
    probs = np.zeros((n, P), dtype=np.float32)
    logits = np.zeros((n, P), dtype=np.float32)

    def one(i):
        L.lb2o_policy_forward(C.byref(net.struct), _ptr(planes[i], C.POINTER(C.c_uint32)), int(rotation[i]),
                              C.c_float(temperature), int(emulate), _ptr(probs[i]), _ptr(logits[i]))

    with cf.ThreadPoolExecutor(_threads()) as ex:
        list(ex.map(one, range(n)))
    return (probs, logits) if want_logits else probs


def value_forward(net: OracleNet, planes: np.ndarray, rotation: np.ndarray, emulate: int = 0) -> np.ndarray:
    L = lib()
    planes = np.ascontiguousarray(planes, dtype=np.uint32)
    rotation = np.ascontiguousarray(rotation, dtype=np.uint8)
    n = planes.shape[0]
    out = np.zeros(n, dtype=np.float32)

    def one(i):
        out[i] = L.lb2o_value_forward(C.byref(net.struct), _ptr(planes[i], C.POINTER(C.c_uint32)),
                                      int(rotation[i]), int(emulate))

    with cf.ThreadPoolExecutor(_threads()) as ex:
        list(ex.map(one, range(n)))
    return out


def trunk_activations(net: OracleNet, planes_one: np.ndarray, rotation: int, emulate: int = 0):
    """All conv outputs of one position: list of float32 [c_out, 361] (network orientation)."""
    L = lib()
    planes_one = np.ascontiguousarray(planes_one, dtype=np.uint32)
    acts = [np.zeros((c.c_out, P), dtype=np.float32) for c in net.weights.convs]
    arr = (_f32p * len(acts))(*[_ptr(a) for a in acts])
    L.lb2o_trunk_activations(C.byref(net.struct), _ptr(planes_one, C.POINTER(C.c_uint32)), int(rotation),
                             int(emulate), arr)
    return acts


def convolve(k, c_in, c_out, x, w, b, round_w=False, round_out=False) -> np.ndarray:
    """One conv + bias + ELU: x float32 [c_in,361], w OIHW, b [c_out] -> [c_out,361]."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32); w = np.ascontiguousarray(w, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.zeros((c_out, P), dtype=np.float32)
    L.lb2o_convolve(k, c_in, c_out, _ptr(x), _ptr(w), _ptr(b), _ptr(out), int(round_w), int(round_out))
    return out


def innerproduct(n_in, n_out, x, w, b) -> np.ndarray:
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32); w = np.ascontiguousarray(w, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.zeros(n_out, dtype=np.float32)
    L.lb2o_innerproduct(n_in, n_out, _ptr(x), _ptr(w), _ptr(b), _ptr(out))
    return out


def softmax(x, temperature=1.0) -> np.ndarray:
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros_like(x)
    L.lb2o_softmax(_ptr(x), _ptr(out), x.size, C.c_float(temperature))
    return out


def rotate_nn_idx(v: int, s: int) -> int:
    return lib().lb2o_rotate_nn_idx(int(v), int(s))


def rev_rotate_nn_idx(v: int, s: int) -> int:
    return lib().lb2o_rev_rotate_nn_idx(int(v), int(s))
