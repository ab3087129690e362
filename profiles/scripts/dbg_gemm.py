"""Prints the in-kernel globaltimer timeline of CTA 0 of the tf32 GEMM (TSD_GEMM_DBG=1)."""
import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tsdiff_b200 import _lib as L
lib = L.load()
dev = 'cuda:0'
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rows in (1826, 24000):
    for act in (0, 3):
        x = torch.randn(rows, 256, device=dev); w = torch.randn(256, 256, device=dev) / 16; b = torch.zeros(256, device=dev)
        out = torch.empty(rows, 256, device=dev); lin = L.linear(w, b)
        for it in range(2):
            L.check(lib.tsd_linear(rows, None, L.ptr(x), C.byref(lin), act, L.ptr(out), 1, st), "lin")
        torch.cuda.synchronize()
