"""Soak: several host threads hammer every entry point (blocking calls of random sizes, ensembles, async
submissions) for a while; every result must equal the one computed alone. python tests/soak.py [seconds]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from leela_b200 import capi, synth

SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
b = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_positions.npz"))
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
pp, vp, rot = b["policy_planes"], b["value_planes"], b["rotation"]
want_p, want_v = ev.eval_both(pp, vp, rot, 0.75)
ens_p, ens_v = ev.eval_ensemble(pp[:64], vp[:64], 0.75)
stop = time.time() + SECONDS
errors, counts = [], [0] * 6


def blocking(t):
    rng = np.random.default_rng(t)
    while time.time() < stop and not errors:
        n = int(rng.choice([1, 2, 7, 33, 100, 256, 300, 700])); lo = int(rng.integers(0, 1024 - n))
        p, v = ev.eval_both(pp[lo:lo + n], vp[lo:lo + n], rot[lo:lo + n], 0.75)
        if not (np.array_equal(p, want_p[lo:lo + n]) and np.array_equal(v, want_v[lo:lo + n])):
            errors.append(("blocking", t, lo, n)); return
        counts[t] += n


def ensembles(t):
    rng = np.random.default_rng(100 + t)
    while time.time() < stop and not errors:
        n = int(rng.choice([1, 3, 20])); lo = int(rng.integers(0, 64 - n))
        p, v = ev.eval_ensemble(pp[lo:lo + n], vp[lo:lo + n], 0.75)
        if not (np.array_equal(p, ens_p[lo:lo + n]) and np.array_equal(v, ens_v[lo:lo + n])):
            errors.append(("ensemble", t, lo, n)); return
        counts[t] += n


def submits(t):
    rng = np.random.default_rng(200 + t)
    while time.time() < stop and not errors:
        idx = [int(i) for i in rng.integers(0, 1024, size=24)]
        outs = ev.submit_many(idx, pp, vp, rot, 0.75) if hasattr(ev, "submit_many") else None
        if outs is None:
            p = ev.eval_policy(pp[idx], rot[idx], 0.75)
            if not np.array_equal(p, want_p[idx]):
                errors.append(("policy", t)); return
        counts[t] += len(idx)


threads = [threading.Thread(target=blocking, args=(0,)), threading.Thread(target=blocking, args=(1,)),
           threading.Thread(target=ensembles, args=(2,)), threading.Thread(target=submits, args=(3,))]
for th in threads: th.start()
for th in threads: th.join()
ev.close()
print(f"soak {SECONDS:.0f} s: positions per thread {counts[:4]}, errors {errors}")
sys.exit(1 if errors else 0)
