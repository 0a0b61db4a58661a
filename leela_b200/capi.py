"""ctypes binding of the C ABI in include/leela_b200.h (libleela_b200.so, built in-tree).

There is deliberately no fallback: if the shared library is missing or the GPU is not a B200
this raises, it never computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LB2_LIB: tuning builds of the same library (tools/ab_variants.py); the product is the in-tree one
LIB_PATH = os.environ.get("LB2_LIB") or os.path.join(HERE, "libleela_b200.so")

POLICY, VALUE = 0, 1
P = 361

# every symbol include/leela_b200.h declares
EXPORTS = [
    "lb2_init", "lb2_destroy", "lb2_net_create", "lb2_net_push_conv", "lb2_net_push_ip", "lb2_net_finalize",
    "lb2_eval_policy", "lb2_eval_value", "lb2_eval_both", "lb2_eval_ensemble", "lb2_eval_both_device", "lb2_submit_policy",
    "lb2_submit_value", "lb2_drain", "lb2_backend_name", "lb2_last_error", "lb2_device_count", "lb2_set_option",
    "lb2_get_option", "lb2_launch_count", "lb2_debug_trunk", "lb2_debug_read_trace", "lb2_planes_from_position", "lb2_eval_positions",
    "lb2_queue_error", "lb2_register_host_buffer", "lb2_unregister_host_buffer",
]

CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int)

_lib = None


class Lb2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lb2 error {code}: {msg}")
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -m leela_b200.build` (needs nvcc); "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    L.lb2_init.argtypes = [C.POINTER(C.c_int), ip, C.POINTER(vp)]
    L.lb2_destroy.argtypes = [vp]; L.lb2_destroy.restype = None
    L.lb2_net_create.argtypes = [vp, ip, C.POINTER(vp)]
    L.lb2_net_push_conv.argtypes = [vp, ip, ip, ip, vp, vp]
    L.lb2_net_push_ip.argtypes = [vp, ip, ip, vp, vp]
    L.lb2_net_finalize.argtypes = [vp]
    L.lb2_eval_policy.argtypes = [vp, vp, vp, ip, fp, vp]
    L.lb2_eval_value.argtypes = [vp, vp, vp, ip, vp]
    L.lb2_eval_both.argtypes = [vp, vp, vp, vp, ip, fp, vp, vp]
    L.lb2_eval_ensemble.argtypes = [vp, vp, vp, ip, fp, vp, vp]
    L.lb2_planes_from_position.argtypes = [vp, ip, ip, ip, ip, fp, vp, vp]
    L.lb2_eval_positions.argtypes = [vp, vp, vp, ip, fp, vp, vp]
    L.lb2_eval_both_device.argtypes = [vp, ip, vp, vp, vp, ip, fp, vp, vp, vp]
    L.lb2_submit_policy.argtypes = [vp, vp, vp, ip, fp, vp, CALLBACK, vp]
    L.lb2_submit_value.argtypes = [vp, vp, vp, ip, vp, CALLBACK, vp]
    L.lb2_drain.argtypes = [vp]
    L.lb2_backend_name.argtypes = [vp]; L.lb2_backend_name.restype = C.c_char_p
    L.lb2_last_error.argtypes = []; L.lb2_last_error.restype = C.c_char_p
    L.lb2_device_count.argtypes = [vp]
    L.lb2_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.lb2_get_option.argtypes = [vp, C.c_char_p]; L.lb2_get_option.restype = C.c_long
    L.lb2_launch_count.argtypes = [vp]; L.lb2_launch_count.restype = C.c_long
    L.lb2_debug_trunk.argtypes = [vp, ip, vp, vp, ip, ip, vp]
    L.lb2_debug_read_trace.argtypes = [vp, vp, C.c_long]
    L.lb2_queue_error.argtypes = [vp, C.c_char_p, ip]
    L.lb2_register_host_buffer.argtypes = [vp, vp, C.c_size_t]
    L.lb2_unregister_host_buffer.argtypes = [vp, vp]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise Lb2Error(rc, load().lb2_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def planes_from_position(stones, white_to_move, ko_point=-1, last_move=-1, prev_move=-1, komi=7.5):
    """Policy and value feature planes (uint32 [361] each) of a raw position; host code, no GPU needed."""
    L = load()
    st = np.ascontiguousarray(stones, dtype=np.uint8).reshape(361)
    pol, val = np.empty(361, dtype=np.uint32), np.empty(361, dtype=np.uint32)
    check(L.lb2_planes_from_position(_p(st), int(white_to_move), int(ko_point), int(last_move), int(prev_move), float(komi),
                                     _p(pol), _p(val)))
    return pol, val


# numpy mirror of struct lb2_position (include/leela_b200.h)
POSITION_DTYPE = np.dtype([("stones", np.uint8, 361), ("white_to_move", np.uint8), ("ko_point", np.int16),
                           ("last_move", np.int16), ("prev_move", np.int16), ("komi", np.float32)], align=True)
assert POSITION_DTYPE.itemsize == 372


class Evaluator:
    """A context with (optionally) a policy and a value net pushed layer by layer, exactly as
    Network::initialize pushes them into the OpenCL backend (Network.cpp:206-233)."""

    def __init__(self, policy=None, value=None, devices=None):
        L = load()
        self._L = L
        self.ctx = C.c_void_p()
        if devices is None:
            check(L.lb2_init(None, 0, C.byref(self.ctx)))
        else:
            arr = (C.c_int * len(devices))(*devices)
            check(L.lb2_init(arr, len(devices), C.byref(self.ctx)))
        self.nets = {}
        if policy is not None:
            self.push_net(POLICY, policy)
        if value is not None:
            self.push_net(VALUE, value)

    def push_net(self, kind, weights):
        L = self._L
        net = C.c_void_p()
        check(L.lb2_net_create(self.ctx, kind, C.byref(net)))
        for c, w, b in zip(weights.convs, weights.conv_w, weights.conv_b):
            w = np.ascontiguousarray(w, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
            check(L.lb2_net_push_conv(net, c.k, c.c_in, c.c_out, _p(w), _p(b)))
        for ip, w, b in zip(weights.ips, weights.ip_w, weights.ip_b):
            w = np.ascontiguousarray(w, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
            check(L.lb2_net_push_ip(net, ip.n_in, ip.n_out, _p(w), _p(b)))
        check(L.lb2_net_finalize(net))
        self.nets[kind] = net

    def close(self):
        if self.ctx:
            self._L.lb2_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------- host-buffer evaluation
    @staticmethod
    def _prep(planes, rotation):
        planes = np.ascontiguousarray(planes, dtype=np.uint32).reshape(-1, P)
        rotation = np.ascontiguousarray(rotation, dtype=np.uint8).reshape(-1)
        assert planes.shape[0] == rotation.shape[0]
        return planes, rotation

    def eval_policy(self, planes, rotation, temp=0.75):
        planes, rotation = self._prep(planes, rotation)
        n = planes.shape[0]
        out = np.empty((n, P), dtype=np.float32)
        check(self._L.lb2_eval_policy(self.ctx, _p(planes), _p(rotation), n, temp, _p(out)))
        return out

    def eval_value(self, planes, rotation):
        planes, rotation = self._prep(planes, rotation)
        n = planes.shape[0]
        out = np.empty(n, dtype=np.float32)
        check(self._L.lb2_eval_value(self.ctx, _p(planes), _p(rotation), n, _p(out)))
        return out

    def eval_both(self, policy_planes, value_planes, rotation, temp=0.75, probs_out=None, win_out=None):
        pp, rotation = self._prep(policy_planes, rotation)
        vp, _ = self._prep(value_planes, rotation)
        n = pp.shape[0]
        probs = probs_out if probs_out is not None else np.empty((n, P), dtype=np.float32)
        win = win_out if win_out is not None else np.empty(n, dtype=np.float32)
        check(self._L.lb2_eval_both(self.ctx, _p(pp), _p(vp), _p(rotation), n, temp, _p(probs), _p(win)))
        return probs, win

    def eval_positions(self, positions, rotation=None, temp=0.75):
        """positions: array of POSITION_DTYPE (raw boards); planes are built inside the library."""
        pos = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
        n = pos.shape[0]
        rot = None if rotation is None else np.ascontiguousarray(rotation, dtype=np.uint8)
        probs = np.empty((n, P), dtype=np.float32); win = np.empty(n, dtype=np.float32)
        check(self._L.lb2_eval_positions(self.ctx, _p(pos), _p(rot), n, temp, _p(probs), _p(win)))
        return probs, win

    def eval_ensemble(self, policy_planes=None, value_planes=None, temp=0.75):
        """AVERAGE_ALL on the device: mean over the 8 symmetries of every position."""
        pp = np.ascontiguousarray(policy_planes, dtype=np.uint32) if policy_planes is not None else None
        vp = np.ascontiguousarray(value_planes, dtype=np.uint32) if value_planes is not None else None
        n = (pp if pp is not None else vp).shape[0]
        probs = np.empty((n, P), dtype=np.float32) if pp is not None else None
        win = np.empty(n, dtype=np.float32) if vp is not None else None
        check(self._L.lb2_eval_ensemble(self.ctx, _p(pp) if pp is not None else None, _p(vp) if vp is not None else None, n, temp,
                                        _p(probs) if probs is not None else None, _p(win) if win is not None else None))
        return probs, win

    def eval_both_raw(self, pp_ptr, vp_ptr, rot_ptr, n, temp, probs_ptr, win_ptr):
        """Host pointers as integers (e.g. pinned torch tensors' data_ptr()); no numpy wrapping."""
        check(self._L.lb2_eval_both(self.ctx, pp_ptr, vp_ptr, rot_ptr, n, temp, probs_ptr, win_ptr))

    def eval_both_device(self, d_pp, d_vp, d_rot, n, temp, d_probs, d_win, stream=0, dev_index=0):
        """All arguments are device pointers (ints); asynchronous on `stream`."""
        check(self._L.lb2_eval_both_device(self.ctx, dev_index, d_pp, d_vp, d_rot, n, temp, d_probs, d_win,
                                           stream or None))

    def debug_trunk(self, kind, planes, rotation, n_layers, c_out):
        planes, rotation = self._prep(planes, rotation)
        n = planes.shape[0]
        out = np.empty((n, c_out, P), dtype=np.float32)
        check(self._L.lb2_debug_trunk(self.ctx, kind, _p(planes), _p(rotation), n, n_layers, _p(out)))
        return out

    def read_trace(self):
        """[n_cta, 96, 16] uint64 timeline of the last traced trunk launch (set_option('trace', 1))."""
        n = 256 * 96 * 16
        buf = np.zeros(n, dtype=np.uint64)
        rc = self._L.lb2_debug_read_trace(self.ctx, _p(buf), n)
        if rc < 0:
            check(rc)
        return buf[:rc * 96 * 16].reshape(rc, 96, 16)

    # -------------------------------------------------------------- misc
    def set_option(self, name, value):
        check(self._L.lb2_set_option(self.ctx, name.encode(), int(value)))

    def get_option(self, name):
        return int(self._L.lb2_get_option(self.ctx, name.encode()))

    def set_precision(self, policy, value):
        """Trunk precision of the two nets: 0 fp16 operands, 1 lite (fp16 + e4m3 corrections), 2 full split operands.
        (Lite and full cannot be mixed between the nets: go through fp16 so that no intermediate state is rejected.)"""
        self.set_option("policy_precision", 0)
        self.set_option("value_precision", value)
        self.set_option("policy_precision", policy)

    @property
    def launch_count(self):
        return int(self._L.lb2_launch_count(self.ctx))

    @property
    def backend(self):
        return self._L.lb2_backend_name(self.ctx).decode()

    def drain(self):
        check(self._L.lb2_drain(self.ctx))
