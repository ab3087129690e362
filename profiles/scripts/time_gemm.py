import ctypes as C, torch, sys, os, time
sys.path.insert(0, '.')
from tsdiff_b200 import _lib as L
lib = L.load()
dev='cuda:0'
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for math in (1, 0):
  for rows in (1826, 24000):
    x = torch.randn(rows, 256, device=dev); w = torch.randn(256,256,device=dev)/16; b = torch.zeros(256, device=dev)
    out = torch.empty(rows,256,device=dev); lin = L.linear(w,b)
    for it in range(5):
        L.check(lib.tsd_linear(rows, None, L.ptr(x), C.byref(lin), 3, L.ptr(out), math, st), "lin")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for it in range(n):
        L.check(lib.tsd_linear(rows, None, L.ptr(x), C.byref(lin), 3, L.ptr(out), math, st), "lin")
    e1.record(); torch.cuda.synchronize()
    print("math", math, "rows", rows, "us per launch (back-to-back, warm):", e0.elapsed_time(e1)*1000/n)
# empty-ish kernel launch rate for comparison
z = torch.zeros(1024, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(200): z.add_(1)
e1.record(); torch.cuda.synchronize(); print("torch tiny kernel us:", e0.elapsed_time(e1)*1000/200)
# CUDA-graph per-node cost: 50 tiny kernels captured and replayed
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): z.add_(1)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    for _ in range(50): z.add_(1)
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20): g.replay()
e1.record(); torch.cuda.synchronize(); print("graph of 50 tiny kernels: us per kernel node:", e0.elapsed_time(e1)*1000/20/50)
# graph of 50 tf32 gemms (24000 rows)
rows=24000
x = torch.randn(rows, 256, device=dev); w = torch.randn(256,256,device=dev)/16; b = torch.zeros(256, device=dev)
out = torch.empty(rows,256,device=dev); lin = L.linear(w,b)
for math in (1,0):
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        stc = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(20): L.check(lib.tsd_linear(rows, None, L.ptr(x), C.byref(lin), 3, L.ptr(out), math, stc), "lin")
    torch.cuda.synchronize()
    for _ in range(3): g2.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10): g2.replay()
    e1.record(); torch.cuda.synchronize(); print("graph of 20 gemms math", math, ": us per gemm:", e0.elapsed_time(e1)*1000/10/20)
import subprocess
print(subprocess.run(["nvidia-smi","-q","-d","COMPUTE,PERFORMANCE"],capture_output=True,text=True).stdout[:1500])
