"""In-kernel %globaltimer timeline of k_node_update (cluster 0's leader).  Needs a library built with -DTSD_NODE_DBG:
    nvcc ... -DTSD_NODE_DBG (python profiles/scripts/node_timeline.py builds its own copy under profiles/ubench/)."""
import ctypes as C, glob, os, subprocess, sys
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
lib.tsd_tune_filter_stack(int(os.environ.get('STACK_MODE', '0')))
lib.tsd_tune_node_pdl(int(os.environ.get('NODE_PDL', '1')))
data = bench.build_inputs(args, 0)
model, cfg = bench.make_models(args, dev)
data_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
names = {0: 'start', 1: 'setup done', 7: 'predecessor complete (pdl wait)', 2: 'agg done (tid 0)', 8: 's0 B ready', 9: 's0 first W', 10: 's0 issued', 11: 's0 acc', 3: 's0 epi done',
         12: 's1 B ready', 13: 's1 first W', 14: 's1 issued', 15: 's1 acc', 4: 's1 epi done', 16: 's2 B ready', 17: 's2 first W',
         18: 's2 issued', 19: 's2 acc', 5: 's2 epi done', 6: 'end'}
for tile in (int(t) for t in (sys.argv[1:] or ['64', '16'])):
    lib.tsd_tune_node_tile(tile)
    eng, runner = bench.build_runner(args, model, data_dev, keep_traj=False)
    runner.prepare(); runner.run(n_steps=1500); torch.cuda.synchronize()
    runner.use_graph = False
    for _ in range(3):
        runner._one_step()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 64)()
    
    lib.tsd_node_dbg_read(buf)
    t0 = buf[0]
    print("tile", tile, "(last k_node_update launch of the step = block 6: aggregation + lin2 + lin, eager launches)")
    for k in sorted(names, key=lambda k: buf[k]):
        if buf[k]:
            print("  %7.2f us  %s" % ((buf[k] - t0) * 1e-3, names[k]))
