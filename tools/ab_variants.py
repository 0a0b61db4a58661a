"""A/B timing of tuning builds of libleela_b200.so on ONE box, interleaved (boxes differ by a few %).
Here:   python tools/ab_variants.py build name[@rev][:DEF=VAL,DEF2=VAL2] ...   (name 'head' = last commit's sources, name@rev = that commit's)
On GPU: python tools/ab_variants.py run [reps]       -> trunk/step us per variant, interleaved rounds
        python tools/ab_variants.py opts reps "name=value,name=value" "..." ...  -> the same for ONE library (LB2_LIB or the
        in-tree one) under different run-time option sets ("" = defaults), one fresh process per measurement"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "_variants")

def do_build(specs):
    from leela_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        if f.endswith(".so"):
            os.remove(os.path.join(VDIR, f))
    for spec in specs:
        name, _, defs = spec.partition(":")
        defines = [d for d in defs.split(",") if d]
        out = os.path.join(VDIR, f"{name}.so")
        rev = "HEAD"
        if "@" in name:
            name, rev = name.split("@")
            out = os.path.join(VDIR, f"{name}.so")
        if name == "head" or rev != "HEAD":
            with tempfile.TemporaryDirectory() as td:
                os.makedirs(os.path.join(td, "leela_b200", "csrc")); os.makedirs(os.path.join(td, "include"))
                for rel in ["leela_b200/csrc/lb2_api.cu", "leela_b200/csrc/lb2_kernels.cu", "leela_b200/csrc/lb2_kernels.cuh",
                            "leela_b200/csrc/lb2_ptx.cuh", "leela_b200/csrc/lb2_planes.cpp", "include/leela_b200.h"]:
                    open(os.path.join(td, rel), "wb").write(subprocess.check_output(["git", "show", f"{rev}:{rel}"], cwd=ROOT))
                build.build(defines=defines, out=out, csrc=os.path.join(td, "leela_b200", "csrc"))
        else:
            build.build(defines=defines, out=out)
        print("built", out, defines)

def one(lib, B, which):
    code = f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch
from leela_b200 import capi, synth
g = np.load(os.path.join({ROOT!r}, "tests", "golden", "bench_positions.npz"))
B = {B}
dev = torch.device("cuda", 0); st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev)
vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("max_batch", max(B, 512))
for kv in filter(None, os.environ.get("LB2_OPTS", "").split(",")):
    ev.set_option(kv.split("=")[0], int(kv.split("=")[1]))
a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr() if {which!r} != "value" else None, win.data_ptr() if {which!r} != "policy" else None)
for _ in range(10): ev.eval_both_device(*a, stream=st.cuda_stream)
torch.cuda.synchronize()
ev.set_option("profile_trunk", 1); ev.get_option("trunk_ns")
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(50): ev.eval_both_device(*a, stream=st.cuda_stream)
e1.record(st); torch.cuda.synchronize()
print("RES %.1f %.1f %.6f" % (e0.elapsed_time(e1) / 50 * 1e3, ev.get_option("trunk_ns") / 50 / 1e3, float(probs.sum()) + float(win.sum())))
"""
    env = dict(os.environ)
    if lib:
        env["LB2_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    for line in r.stdout.splitlines():
        if line.startswith("RES"):
            _, a, b, c = line.split()
            return float(a), float(b), float(c)
    return None, None, (r.stdout + r.stderr)[-400:]

def do_run(reps, B=256, which="both"):
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    res = {l: [] for l in libs}
    for r in range(reps):
        for l in libs:
            step, trunk, chk = one(os.path.join(VDIR, l), B, which)
            res[l].append((step, trunk))
            print(f"round {r} {l:28s} step {step} us trunk {trunk} us checksum {chk}", flush=True)
    print("--- median trunk us / step us")
    import statistics
    for l in libs:
        ok = [x for x in res[l] if x[0] is not None]
        if ok:
            print(f"{l:28s} trunk {statistics.median(x[1] for x in ok):8.1f}  step {statistics.median(x[0] for x in ok):8.1f}  (n={len(ok)})")

def do_opts(reps, optsets, B=256, which="both"):
    import statistics
    res = {o: [] for o in optsets}
    for r in range(reps):
        for o in optsets:
            os.environ["LB2_OPTS"] = o
            step, trunk, chk = one(os.environ.get("LB2_LIB", ""), B, which)
            res[o].append((step, trunk))
            print(f"round {r} {o or 'defaults':44s} step {step} us trunk {trunk} us checksum {chk}", flush=True)
    print("--- median trunk us / step us")
    for o in optsets:
        ok = [x for x in res[o] if x[0] is not None]
        if ok:
            print(f"{o or 'defaults':44s} trunk {statistics.median(x[1] for x in ok):8.1f}  step {statistics.median(x[0] for x in ok):8.1f}  (n={len(ok)})")

if __name__ == "__main__":
    if sys.argv[1] == "build":
        do_build(sys.argv[2:])
    elif sys.argv[1] == "opts":
        do_opts(int(sys.argv[2]), sys.argv[3:], int(os.environ.get("AB_B", "256")), os.environ.get("AB_WHICH", "both"))
    else:
        do_run(int(sys.argv[2]) if len(sys.argv) > 2 else 3, int(sys.argv[3]) if len(sys.argv) > 3 else 256, sys.argv[4] if len(sys.argv) > 4 else "both")
