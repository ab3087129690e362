"""In-kernel %globaltimer timeline of k_node_chain (CTA 0) inside a real Langevin step.  Builds its own library copy with
-DTSD_NODE_DBG -DTSD_FS_DBG under profiles/ubench/ (`--build` only builds)."""
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
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
data_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
eng, runner = bench.build_runner(args, model, data_dev, keep_traj=False)
runner.prepare(); runner.run(n_steps=1500); torch.cuda.synchronize()
runner.use_graph = False
for _ in range(3):
    runner._one_step()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 128)()
lib.tsd_node_chain_dbg_read(buf)
names = {0: 'block starts (previous epilogue done)', 1: 'grid barrier passed', 2: 'filter rows landed', 3: 'aggregated',
         4: 'stage 0 accumulator', 5: 'stage 1 accumulator', 6: 'stage 2 accumulator', 7: 'last epilogue done'}
ev = sorted((buf[i], i) for i in range(128) if buf[i])
t0 = ev[0][0]
print("k_node_chain CTA 0 (us from its first stamp): block  event")
for t, i in ev:
    print("  %7.2f  B%d  %s" % ((t - t0) * 1e-3, i // 8, names[i % 8]))
