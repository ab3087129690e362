"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from tsdiff_b200 import _lib as L
from tsdiff_b200.config import AttrDict, QM9_DEFAULT_MODEL, TRAIN_CONFIG_MODEL
from tsdiff_b200.models.epsnet import get_model
from tsdiff_b200.synthetic import make_batch, shard_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tsdiff_b200.h")).read()
    declared = set(re.findall(r"\b(tsd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    if not os.path.exists(L.LIB_PATH):
        from tsdiff_b200 import build
        build.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert L.load().tsd_version() >= 100


def test_struct_layouts_match_header(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof with the ctypes mirrors."""
    import subprocess
    structs = {"tsd_linear_t": L.Linear, "tsd_batch_t": L.Batch, "tsd_edges_t": L.Edges,
               "tsd_edge_encoder_t": L.EdgeEncoder, "tsd_interaction_t": L.Interaction, "tsd_gine_t": L.Gine,
               "tsd_pair_mlp_t": L.PairMlp, "tsd_score_channel_t": L.ScoreChannel, "tsd_ld_params_t": L.LdParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tsdiff_b200.h"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_get_model_dispatch_and_errors():
    assert type(get_model(TRAIN_CONFIG_MODEL)).__name__ == "CondenseEncoderEpsNetwork"
    assert type(get_model(QM9_DEFAULT_MODEL)).__name__ == "DualEncoderEpsNetwork"
    with pytest.raises(NotImplementedError):
        get_model(AttrDict(network="nope"))


def test_config_access_patterns():
    cfg = AttrDict({"a": 1, "encoder": {"name": "schnet"}})
    assert cfg.a == 1 and cfg.get("b", 7) == 7 and cfg.get("encoder").name == "schnet"
    assert hasattr(cfg, "a") and not hasattr(cfg, "TS")


def test_no_cpu_fallback():
    m = get_model(TRAIN_CONFIG_MODEL)
    g = make_batch(1, seed=0)
    with pytest.raises(L.TsdError):
        m(g["atom_type"], g["r_feat"], g["p_feat"], g["pos_init"], g["bond_index"], g["bond_type"], g["batch"], None)
    args = (g["atom_type"], g["r_feat"], g["p_feat"], g["pos_init"], g["bond_index"], g["bond_type"], g["batch"])
    with pytest.raises(L.TsdError):  # gradients enabled: the training kernels are CUDA only as well
        m.get_loss(*args)
    with torch.no_grad(), pytest.raises(L.TsdError):  # forward value needs the CUDA path too
        m.get_loss(*args)


def test_checkpoint_roundtrip_strict():
    torch.manual_seed(0)
    a = get_model(TRAIN_CONFIG_MODEL)
    torch.manual_seed(1)
    b = get_model(TRAIN_CONFIG_MODEL)
    b.load_state_dict(a.state_dict(), strict=True)  # sampling.py:130
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k


def test_synthetic_batch_contract():
    g = make_batch(20, seed=1)
    n = g["atom_type"].numel()
    bi, bt = g["bond_index"], g["bond_type"]
    key = bi[0] * n + bi[1]
    assert bool((key[1:] > key[:-1]).all())  # sorted, no duplicates (datasets.py:495-498)
    pairs = set(map(tuple, bi.t().tolist()))
    assert all((c, r) in pairs for r, c in pairs)  # symmetric
    assert bool((g["batch"][bi[0]] == g["batch"][bi[1]]).all())
    assert bool((g["r_feat"].sum(1) == 8).all()) and g["r_feat"].shape == (n, 25)
    assert 10 <= int(g["num_nodes_per_graph"].min()) and int(g["num_nodes_per_graph"].max()) <= 25
    assert bool(((bt // 22 > 0) | (bt % 22 > 0)).all())


def test_shards_partition_the_batch():
    g = make_batch(10, seed=2)
    parts = [shard_batch(g, r, 4) for r in range(4)]
    assert sum(p["num_graphs"] for p in parts) == 10
    assert torch.equal(torch.cat([p["atom_type"] for p in parts]), g["atom_type"])
    off = 0
    for p in parts:
        assert p["atom_offset"] == off
        off += p["atom_type"].numel()
        assert int(p["bond_index"].min()) >= 0 and int(p["bond_index"].max()) < p["atom_type"].numel()


def test_workspace_bytes_is_host_arithmetic():
    """tsd_workspace_bytes needs no GPU: buffer sizes and counts for both networks / arithmetic modes."""
    import ctypes as C
    lib = L.load()
    eb, nb, ne, nn = C.c_uint64(), C.c_uint64(), C.c_int32(), C.c_int32()
    for network, math, want_nodes in ((0, 0, 4), (0, 1, 6), (1, 0, 7), (1, 1, 9)):
        rc = lib.tsd_workspace_bytes(1750, 28900, 256, network, math, C.byref(eb), C.byref(nb), C.byref(ne), C.byref(nn))
        assert rc == 0
        assert eb.value == 28900 * 256 * 4 and nb.value == 1750 * 256 * 4
        assert ne.value == 7 and nn.value == want_nodes
    assert lib.tsd_workspace_bytes(10, 10, 0, 0, 0, None, None, None, None) != 0  # invalid hidden


def test_ddpm_schedule_matches_oracle_coefficients():
    """engine.ddpm_schedule (the 8-column table K7 reads) against the oracle's per-step restatement of
    sampler.py:216-236, for the default range and for a denoise_from_time_t range that reaches t = 0."""
    from oracle import tsdiff_oracle as O
    from tsdiff_b200 import engine as E
    betas, _ = O.schedule_tensors(TRAIN_CONFIG_MODEL)
    for t_end, n_steps in ((5000, 7), (12, 12), (600, 3)):
        table = E.ddpm_schedule(betas, t_end, n_steps)
        assert table.shape == (n_steps, 8) and table.dtype == torch.float32
        seq = list(range(t_end - n_steps, t_end))
        seq_next = [-1] + seq[:-1]
        for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
            want = torch.cat(O.ddpm_coefficients(betas, i, j))
            assert torch.equal(table[k], want), (t_end, k)
    last = E.ddpm_schedule(betas, 12, 12)[-1]  # t = 0: no noise, atm1 = 1
    assert float(last[6]) == 0.0 and float(last[7]) == 1.0


def test_ld_schedule_matches_golden_constants():
    import json
    from oracle import tsdiff_oracle as O
    from tsdiff_b200 import engine as E
    _, alphas = O.schedule_tensors(TRAIN_CONFIG_MODEL)
    sched, sigmas = E.ld_schedule(alphas, 5000, 1e-7)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "schedule.json")))
    assert abs(float(sigmas[0]) - gold["sigma_0"]) < 1e-9 and abs(float(sigmas[-1]) - gold["sigma_last"]) < 1e-6
    assert sched.shape == (5000, 4) and float(sched[0, 1]) == float(sigmas[-1])  # first step uses sigma_T
    step = 1e-7 * (float(sigmas[-1]) / 0.01) ** 2
    assert abs(float(sched[0, 0]) - step) < 1e-6 * step


def test_dualenc_branch_schedule_matches_oracle_coefficients():
    """engine.dualenc_branch_schedule against the oracle's restatement of dualenc.py:861-944, including
    the last time index (t = 0: no noise)."""
    from oracle import tsdiff_oracle as O
    from tsdiff_b200 import engine as E
    betas, alphas = O.schedule_tensors(QM9_DEFAULT_MODEL)
    T = alphas.numel()
    for kind, eta in (("ddpm_noisy", 1.0), ("ddpm_det", 1.0), ("generalized", 1.0), ("generalized", 0.5)):
        n_steps = 6
        table = E.dualenc_branch_schedule(alphas, betas, n_steps, 1e-7, kind, eta=eta)
        seq = list(range(T - n_steps, T))
        seq_next = [-1] + seq[:-1]
        for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
            want = torch.cat([c.reshape(1) for c in O.dualenc_step_coefficients(alphas, betas, i, j, kind, 1e-7, eta)])
            assert torch.equal(table[k, :want.numel()], want), (kind, k)
            assert float(table[k, 6 if kind != "generalized" else 3]) == 1.0  # global channel on (no start sigma)
    full = E.dualenc_branch_schedule(alphas, betas, T, 1e-7, "ddpm_noisy")
    assert float(full[-1, 5]) == 0.0  # t == 0
