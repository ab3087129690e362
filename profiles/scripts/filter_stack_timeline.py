"""In-kernel %globaltimer timeline of k_filter_stack (CTA 0) and per-CTA phases of k_node_update inside a real Langevin
step.  Builds its own library copy with -DTSD_FS_DBG -DTSD_NODE_DBG under profiles/ubench/."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, '.')
import torch
from tsdiff_b200 import build as B, _lib as L
lib_dbg = os.path.join('profiles', 'ubench', 'libtsdiff_b200_dbg.so')
if '--build' in sys.argv or not os.path.exists(lib_dbg):
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + ['-DTSD_NODE_DBG', '-DTSD_FS_DBG'] + B.sources() + ['-o', lib_dbg])
    if '--build' in sys.argv:
        sys.exit(0)
L.LIB_PATH = lib_dbg
import bench
class A: pass
args = A(); args.batch = 100; args.network = 'condensenc'; args.math = 'tf32'; args.ld_steps = 5000
dev = torch.device('cuda:0')
lib = L.load()
lib.tsd_tune_filter_stack(0)
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
data_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
eng, runner = bench.build_runner(args, model, data_dev, keep_traj=False)
runner.prepare(); runner.run(n_steps=1500); torch.cuda.synchronize()
runner.use_graph = False
for _ in range(3):
    runner._one_step()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 256)()
lib.tsd_fs_dbg_read(buf)
names = {0: 'epi  acc1 ready', 1: 'epi  X quarter 0', 2: 'epi  X quarter 1', 3: 'epi  X quarter 2', 4: 'epi  X quarter 3',
         5: 'st   acc2 ready', 6: 'st   acc2 read, stores issued', 7: '-', 8: 'st   last store read', 9: 'mma  A first panel',
         10: 'mma  A issued', 11: 'mma  X quarter 0 seen', 12: 'mma  B issued', 13: 'tma  layer first panel', 14: 'tma  A last panel',
         15: 'tma  B last panel'}
t0 = min(v for v in buf if v)
ev = sorted((buf[i], i) for i in range(256) if buf[i])
print("k_filter_stack CTA 0 (us from its first stamp): layer  event")
for t, i in ev:
    print("  %7.2f  L%d  %s" % ((t - t0) * 1e-3, i // 16, names[i % 16]))
cta = (C.c_ulonglong * 1024)()
lib.tsd_node_cta_read(cta)
rows = [(cta[4 * i], cta[4 * i + 1], cta[4 * i + 2], cta[4 * i + 3]) for i in range(256) if cta[4 * i]]
k0 = min(r[0] for r in rows)
print("\nk_node_update (last launch of the step), per CTA: start, aggregation us, total us, in-edges")
for i, (a, b, c, e) in enumerate(rows):
    print("  cta %3d  start %6.2f  agg %6.2f  total %6.2f  in-edges %5d" % (i, (a - k0) * 1e-3, (b - a) * 1e-3, (c - a) * 1e-3, e))
print("kernel span %.2f us" % ((max(r[2] for r in rows) - k0) * 1e-3))
