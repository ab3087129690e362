"""Cold (L2 flushed, write+read pass) and warm timing of k_cfconv_aggregate.
usage: time_agg.py [num_reactions=100] [min_atoms=10] [max_atoms=25] [cutoff=10.0] [max_neighbors=32]
(BASELINE config 5, the stress case: 1000 55 65 15.0)"""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from tsdiff_b200 import _lib as L, engine as E
from tsdiff_b200.synthetic import make_batch
lib = L.load(); dev = 'cuda:0'
a = sys.argv[1:]
G = int(a[0]) if len(a) > 0 else 100
lo, hi = (int(a[1]), int(a[2])) if len(a) > 2 else (10, 25)
cutoff = float(a[3]) if len(a) > 3 else 10.0
g = make_batch(G, seed=1000, min_atoms=lo, max_atoms=hi)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()}
plan = E.BatchPlan(0, d['batch'], d['bond_index'], d['bond_type'], 4, 3)
plan.build_edges((d['pos_init'] * 2.0).contiguous(), cutoff)
e = plan.edge_count(); n = plan.num_nodes; h = 256
x1 = torch.randn(n, h, device=dev); filt = torch.randn(plan.edge_capacity, h, device=dev); agg = torch.empty(n, h, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fw = torch.empty(256 << 20, dtype=torch.uint8, device=dev); fr = torch.zeros(64 << 20, device=dev)
def run(): L.check(lib.tsd_cfconv_aggregate(C.byref(plan.c_batch), C.byref(plan.c_edges), h, L.ptr(x1), L.ptr(filt), L.ptr(agg), st), "agg")
for _ in range(3): run()
cold = []
for _ in range(10):
    fw.zero_(); fr.sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); cold.append(a.elapsed_time(b) * 1e3)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): run()
b.record(); torch.cuda.synchronize()
nbytes = e * h * 4 + 2 * n * h * 4 + e * 8 + (n + 1) * 4
c = sorted(cold)[len(cold) // 2]
print("reactions", G, "N", n, "E", e, "cold median %.1f us (%.0f GB/s)  warm %.1f us" % (c, nbytes / c / 1e3, a.elapsed_time(b) * 1e3 / 50))
