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


def queue(t):
    """single-position requests through lb2_submit_* with one request outstanding, like a search thread"""
    import ctypes as C
    L = capi.load()
    rng = np.random.default_rng(300 + t)
    sem = threading.Semaphore(0)
    status = []
    cb = capi.CALLBACK(lambda user, st: (status.append(st), sem.release()))
    out_p = np.zeros((1, 361), np.float32); out_v = np.zeros(1, np.float32)
    while time.time() < stop and not errors:
        i = int(rng.integers(0, 1024))
        planes_p = np.ascontiguousarray(pp[i:i + 1]); planes_v = np.ascontiguousarray(vp[i:i + 1]); r = np.ascontiguousarray(rot[i:i + 1])
        if rng.integers(0, 3) == 0:
            capi.check(L.lb2_submit_policy(ev.ctx, planes_p.ctypes.data, r.ctypes.data, 1, 0.75, out_p.ctypes.data, cb, None)); sem.acquire()
            ok = np.array_equal(out_p[0], want_p[i])
        else:
            capi.check(L.lb2_submit_value(ev.ctx, planes_v.ctypes.data, r.ctypes.data, 1, out_v.ctypes.data, cb, None)); sem.acquire()
            ok = out_v[0] == want_v[i]
        if not ok or status[-1] != 0:
            errors.append(("queue", t, i)); return
        counts[t] += 1


def options(t):
    """flips run-time options under the other threads' feet: cached graphs must be rebuilt, the queue's dispatchers change
    their batching policy, results must not change"""
    k = 0
    while time.time() < stop and not errors:
        time.sleep(0.05)
        ev.set_option("use_graphs", k & 1); ev.set_option("small_batch", 48 if k & 2 else 0); ev.set_option("resident_weights", 2 if k & 4 else 0)
        ev.set_option("queue_linger", 0 if k & 8 else 1)
        k += 1
    ev.set_option("use_graphs", 1); ev.set_option("small_batch", 48); ev.set_option("resident_weights", 2); ev.set_option("queue_linger", 1)


counts += [0] * 10
threads = [threading.Thread(target=blocking, args=(0,)), threading.Thread(target=blocking, args=(1,)),
           threading.Thread(target=ensembles, args=(2,)), threading.Thread(target=submits, args=(3,)),
           threading.Thread(target=options, args=(15,))] + [threading.Thread(target=queue, args=(4 + i,)) for i in range(8)]
for th in threads: th.start()
for th in threads: th.join()
ev.close()
print(f"soak {SECONDS:.0f} s: positions per thread {counts[:12]}, errors {errors}")
sys.exit(1 if errors else 0)
