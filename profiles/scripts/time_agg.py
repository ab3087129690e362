"""Cold (L2 flushed, write+read pass) and warm timing of k_cfconv_aggregate on a batch-100 edge list."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from tsdiff_b200 import _lib as L, engine as E
from tsdiff_b200.synthetic import make_batch
lib = L.load(); dev = 'cuda:0'
g = make_batch(100, seed=1000)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()}
plan = E.BatchPlan(0, d['batch'], d['bond_index'], d['bond_type'], 4, 3)
plan.build_edges((d['pos_init'] * 2.0).contiguous(), 10.0)
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
print("E", e, "cold median %.1f us (%.0f GB/s)  warm %.1f us" % (c, nbytes / c / 1e3, a.elapsed_time(b) * 1e3 / 50))
