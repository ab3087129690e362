"""Post-sampling geometry metrics on the GPU -- host-side mirror of the reference's analysis helpers
clustering.py:98-105 (`calc_DMAE`) and :123-135 (`get_minimum_matches`), same names and argument meaning,
batched over many generated geometries.  fp64 like the reference (numpy / scipy.pdist).  All arithmetic runs in
libtsdiff_b200.so; there is no CPU fallback (CPU tensors raise)."""
import ctypes as C

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64_cuda(t, name):
    t = torch.as_tensor(t)
    if not t.is_cuda:
        raise L.TsdError("%s must be a CUDA tensor: tsdiff_b200 has no CPU path" % name)
    return t.to(torch.float64).contiguous()


def calc_DMAE(dm_ref, dm_guess, mape=False):
    """clustering.py:98-105.  dm_ref (n, n); dm_guess (n, n) -> 0-d tensor, or (B, n, n) -> (B,)."""
    dm_ref, dm_guess = _f64_cuda(dm_ref, "dm_ref"), _f64_cuda(dm_guess, "dm_guess")
    single = dm_guess.dim() == 2
    g = dm_guess.reshape(-1, dm_ref.size(0), dm_ref.size(0))
    out = torch.empty(g.size(0), dtype=torch.float64, device=g.device)
    with torch.cuda.device(g.device):
        L.check(L.load().tsd_dmae(dm_ref.size(0), g.size(0), L.ptr(dm_ref), L.ptr(g), int(bool(mape)), L.ptr(out), _stream()),
                "tsd_dmae")
    return out[0] if single else out


def calc_DMAE_from_positions(pos_ref, pos_guess, mape=False):
    """calc_DMAE of the distance matrices of pos_ref (n, 3) and pos_guess ((B,) n, 3) without materialising them."""
    pos_ref, pos_guess = _f64_cuda(pos_ref, "pos_ref"), _f64_cuda(pos_guess, "pos_guess")
    single = pos_guess.dim() == 2
    g = pos_guess.reshape(-1, pos_ref.size(0), 3)
    out = torch.empty(g.size(0), dtype=torch.float64, device=g.device)
    with torch.cuda.device(g.device):
        L.check(L.load().tsd_dmae_pos(pos_ref.size(0), g.size(0), L.ptr(pos_ref), L.ptr(g), int(bool(mape)), L.ptr(out),
                                      _stream()), "tsd_dmae_pos")
    return out[0] if single else out


def get_minimum_matches(ref, prb, matches, return_type="value"):
    """clustering.py:123-135 with its default metric sum((d_ref - d_prb) ** 2) over scipy.pdist order.
    ref (n, 3); prb (n, 3) or (B, n, 3); matches: (M, n) atom-index permutations.
    return_type 'value' -> the minimum (0-d / (B,)), otherwise the minimising permutation(s) ((n,) / (B, n))."""
    ref, prb = _f64_cuda(ref, "ref"), _f64_cuda(prb, "prb")
    matches = torch.as_tensor(matches)
    if matches.numel() == 0:
        raise ValueError("min() arg is an empty sequence")  # what the reference raises on an empty list
    matches = matches.to(device=ref.device, dtype=torch.int32).reshape(-1, ref.size(0)).contiguous()
    single = prb.dim() == 2
    p = prb.reshape(-1, ref.size(0), 3)
    lib = L.load()
    nd, ni = C.c_int64(), C.c_int64()
    L.check(lib.tsd_min_match_scratch(p.size(0), matches.size(0), C.byref(nd), C.byref(ni)), "tsd_min_match_scratch")
    sv = torch.empty(max(nd.value, 1), dtype=torch.float64, device=ref.device)
    si = torch.empty(max(ni.value, 1), dtype=torch.int32, device=ref.device)
    val = torch.empty(p.size(0), dtype=torch.float64, device=ref.device)
    idx = torch.empty(p.size(0), dtype=torch.int32, device=ref.device)
    with torch.cuda.device(ref.device):
        L.check(lib.tsd_min_match(ref.size(0), p.size(0), matches.size(0), L.ptr(ref), L.ptr(p), L.ptr(matches), L.ptr(sv),
                                  L.ptr(si), L.ptr(val), L.ptr(idx), _stream()), "tsd_min_match")
    if return_type == "value":
        return val[0] if single else val
    best = matches[idx.long()].long()
    return best[0] if single else best
