"""Timing of the per-step edge build K2 (k_edge_count + k_edge_emit) with CUDA events.
usage: time_edge_build.py [num_reactions=100] [min_atoms=10] [max_atoms=25] [cutoff=10.0]
Algorithmic bytes: N*12 read + per edge 4 (row) + 4 (col) + 4 (length) + 4 + 4 (type codes) + 1 (in_b)
+ 4 + 4 (in_eid, in_src) written + 2 (N+1) 4 (row_ptr, in_ptr)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from tsdiff_b200 import engine as E
from tsdiff_b200.synthetic import make_batch
dev = 'cuda:0'
a = sys.argv[1:]
G = int(a[0]) if len(a) > 0 else 100
lo, hi = (int(a[1]), int(a[2])) if len(a) > 2 else (10, 25)
cutoff = float(a[3]) if len(a) > 3 else 10.0
g = make_batch(G, seed=1000, min_atoms=lo, max_atoms=hi)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()}
plan = E.BatchPlan(0, d['batch'], d['bond_index'], d['bond_type'], 4, 3)
pos = (d['pos_init'] * 2.0).contiguous()
for _ in range(3): plan.build_edges(pos, cutoff)
e, n = plan.edge_count(), plan.num_nodes
fw = torch.empty(256 << 20, dtype=torch.uint8, device=dev); fr = torch.zeros(64 << 20, device=dev)
cold = []
for _ in range(10):
    fw.zero_(); fr.sum()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); plan.build_edges(pos, cutoff); t1.record(); torch.cuda.synchronize(); cold.append(t0.elapsed_time(t1) * 1e3)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(50): plan.build_edges(pos, cutoff)
t1.record(); torch.cuda.synchronize()
nbytes = n * 12 + e * 29 + 2 * (n + 1) * 4
c = sorted(cold)[len(cold) // 2]
print("reactions", G, "N", n, "E", e, "edge build (2 kernels): cold median %.1f us (%.0f GB/s)  warm %.1f us" % (c, nbytes / c / 1e3, t0.elapsed_time(t1) * 1e3 / 50))
