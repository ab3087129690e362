"""A/B harness: builds one library per combination of the compile-time experiment switches (tc_common.cuh) under
profiles/ubench/variants/ (`python profiles/scripts/variants.py build`, CPU only) and measures every variant on the GPU
(`python profiles/scripts/variants.py run`): mean Langevin-step time over 1500 graph replays (late-trajectory edge
count) and the isolated, L2-flushed filter-network kernel."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "ubench", "variants")
VARIANTS = {"new": dict()}
TILES = {"new": (321, 481, 641)}


def build():
    from tsdiff_b200 import build as B
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-D%s=%d" % kv for kv in flags.items()] + B.sources() + ["-o", os.path.join(OUT, name + ".so")]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        print(name, "ok" if p.returncode == 0 else "FAILED\n" + out)


def run_one(name, tile):
    code = r'''
import sys, ctypes as C, json, torch
sys.path.insert(0, %r)
from tsdiff_b200 import _lib as L
L.LIB_PATH = %r
import bench
class A: pass
args = A(); args.batch = 100; args.network = 'condensenc'; args.math = 'tf32'; args.ld_steps = 5000; args.members = 1; args.mode = 'shard'
dev = torch.device('cuda:0')
lib = L.load()
if %d: lib.tsd_tune_node_tile(%d)
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
dd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
eng, runner = bench.build_runner(args, [model], dd, keep_traj=False)
runner.prepare(); runner.run(n_steps=1500); torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(1500): runner.graph.replay()
t1.record(); torch.cuda.synchronize()
step_us = t0.elapsed_time(t1) / 1500 * 1e3
peaks = bench.measured_peaks()
tr, hb = bench.kernel_rooflines(args, eng, peaks, dev)
print(json.dumps({"variant": %r, "tile": %d, "step_us": step_us, "filter_us_cold": tr["us_per_launch"], "agg_us_cold": hb["us_per_launch"]}))
''' % (ROOT, os.path.join(OUT, name + ".so"), tile, tile, name, tile)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    print(line[-1] if line else "FAILED %s: %s" % (name, r.stderr[-400:]), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        for name in VARIANTS:
            for tile in TILES[name]:
                run_one(name, tile)
