"""Generates the committed golden vectors by running the reference's OWN Python verbatim.

Run from the repo root, in the build container (needs /root/reference, CPU only):

    python tests/golden/make_golden.py

It imports /root/reference/models through oracle/ref_harness.py (third-party wheels
replaced by oracle/third_party.py), builds the reference networks with seeded default
initialisation, runs forward / Langevin sampling on the rxn_0 fixture graph and on small
synthetic batches, and stores inputs + outputs under tests/golden/.  The GPU box has no
reference tree: tests only read the files written here.
"""
import contextlib
import io
import json
import os
import pickle
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import ref_harness as rh  # noqa: E402
from tsdiff_b200.synthetic import make_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_rxn0():
    """birkholz_benchmark/rxn_0/samples_all.pkl holds 100 PyG Data objects of the same
    featurised reaction; PyG / rdkit classes are replaced by inert holders."""
    class Holder:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, s):
            self.__dict__["state"] = s

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("torch_geometric") or module.startswith("rdkit"):
                return type(name, (Holder,), {})
            return super().find_class(module, name)

    with open(os.path.join(rh.REFERENCE_ROOT, "birkholz_benchmark/rxn_0/samples_all.pkl"), "rb") as f:
        st = Unpickler(f).load()[0].state
    n = st["atom_type"].numel()
    return {
        "atom_type": st["atom_type"], "r_feat": st["r_feat"], "p_feat": st["p_feat"],
        "bond_index": st["edge_index"], "bond_type": st["edge_type"],
        "batch": torch.zeros(n, dtype=torch.long), "num_graphs": 1,
        "num_nodes_per_graph": torch.tensor([n]),
    }


def manifest(model):
    out = {}
    for k, v in model.state_dict().items():
        v64 = v.double()
        out[k] = {"shape": list(v.shape), "sum": float(v64.sum()), "abs_sum": float(v64.abs().sum())}
    return out


@contextlib.contextmanager
def injected_noise(noise):
    """Make the reference's `torch.randn_like(pos)` (sampler.py:213, dualenc.py:858) return
    pre-generated rows so the oracle / CUDA path can consume the identical stream."""
    it = iter(noise)
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: next(it).to(t)
    try:
        yield
    finally:
        torch.randn_like = orig


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def main():
    epsnet, sampler, common, geometry = rh.import_reference()
    cfg_b = rh.load_yaml_config("configs/train_config.yml").model
    cfg_a = rh.load_yaml_config("configs/geodiff_legacy/qm9_default.yml").model

    rxn0 = load_rxn0()
    syn4 = make_batch(4, seed=3, sizes=[10, 17, 25, 12])
    torch.save({k: v for k, v in rxn0.items()}, os.path.join(OUT, "rxn0_graph.pt"))
    torch.save(syn4, os.path.join(OUT, "syn4_graph.pt"))

    manifests = {}
    gold = {}

    # ------------------------------------------------------------------ path B (condensenc)
    torch.manual_seed(0)
    mb = epsnet.get_model(cfg_b)
    mb.eval()
    manifests["condensenc_train_config_seed0"] = manifest(mb)
    torch.manual_seed(1)
    mb1 = epsnet.get_model(cfg_b)
    manifests["condensenc_train_config_seed1"] = manifest(mb1)

    def fwd_b(model, g, pos):
        t = torch.zeros(g["num_graphs"], dtype=torch.long)
        with torch.no_grad():
            ei, idx, ln = model(g["atom_type"], g["r_feat"], g["p_feat"], pos, g["bond_index"],
                                g["bond_type"], g["batch"], t)
        return {"pos": pos, "edge_inv": ei, "edge_index": idx, "edge_length": ln}

    torch.manual_seed(2022)
    pos_a = torch.randn(13, 3)
    gold["b_rxn0_fwd"] = fwd_b(mb, rxn0, pos_a)
    torch.manual_seed(7)
    gold["b_rxn0_fwd_wide"] = fwd_b(mb, rxn0, torch.randn(13, 3) * 5.0)
    gold["b_syn4_fwd"] = fwd_b(mb, syn4, syn4["pos_init"] * 3.0)
    gold["b_syn4_fwd_wide"] = fwd_b(mb, syn4, syn4["pos_init"] * 6.0)

    # graph-builder outputs (common.py:115-223) at orders 4 and 3
    for name, g, pos in (("rxn0", rxn0, pos_a * 5.0), ("syn4", syn4, syn4["pos_init"] * 6.0)):
        n = g["atom_type"].numel()
        for order in (3, 4):
            with torch.no_grad():
                loc, tr, tp_ = common._extend_ts_graph_order(n, g["bond_index"], g["bond_type"], g["batch"], order=order)
                glob, loc2, tr2, tp2 = common.extend_ts_graph_order_radius(
                    n, pos, g["bond_index"], g["bond_type"], g["batch"], order=order, cutoff=10.0)
                ga, ta = common.extend_graph_order_radius(n, pos, g["bond_index"], g["bond_type"], g["batch"],
                                                          order=order, cutoff=10.0)
            gold["graph_%s_o%d" % (name, order)] = {
                "pos": pos, "local_index": loc, "local_type_r": tr, "local_type_p": tp_,
                "global_index": glob, "a_index": ga, "a_type": ta}

    # Langevin trajectories through EnsembleSampler.dynamic_sampling (sampler.py:118-257)
    def ld_b(models, g, pos_init, n_steps, seed):
        gen = torch.Generator().manual_seed(seed)
        noise = torch.randn(n_steps, pos_init.size(0), 3, generator=gen)
        ens = sampler.EnsembleSampler(models)
        with injected_noise(noise), quiet():
            pos, traj = ens.dynamic_sampling(
                g["atom_type"], g["r_feat"], g["p_feat"], pos_init, g["bond_index"], g["bond_type"],
                g["batch"], g["num_graphs"], extend_order=True, n_steps=n_steps, step_lr=1e-7,
                clip=1000, sampling_type="ld")
        return {"pos_init": pos_init, "noise": noise, "pos": pos, "traj": torch.stack(traj)}

    gold["b_rxn0_ld20"] = ld_b([mb], rxn0, pos_a, 20, 11)
    gold["b_syn4_ld10"] = ld_b([mb], syn4, syn4["pos_init"], 10, 12)
    gold["b_rxn0_ens2_ld5"] = ld_b([mb, mb1], rxn0, pos_a, 5, 13)
    t0 = torch.zeros(1, dtype=torch.long)
    with torch.no_grad():
        ens_out = sampler.EnsembleSampler([mb, mb1])(
            rxn0["atom_type"], rxn0["r_feat"], rxn0["p_feat"], pos_a, rxn0["bond_index"],
            rxn0["bond_type"], rxn0["batch"], t0)
    gold["b_rxn0_ens2_fwd"] = {"pos": pos_a, "edge_inv": ens_out[0], "edge_index": ens_out[1],
                               "edge_length": ens_out[2]}

    # --------------------------------------------------------------------- path A (dualenc)
    torch.manual_seed(0)
    ma = epsnet.get_model(cfg_a)
    ma.eval()
    manifests["dualenc_qm9_default_seed0"] = manifest(ma)

    def fwd_a(g, pos):
        t = torch.zeros(g["num_graphs"], dtype=torch.long)
        with torch.no_grad():
            out = ma(g["atom_type"], pos, g["bond_index"], g["bond_type"], g["batch"], t, return_edges=True)
        keys = ("edge_inv_global", "edge_inv_local", "edge_index", "edge_type", "edge_length", "local_edge_mask")
        d = dict(zip(keys, out))
        d["pos"] = pos
        return d

    gold["a_rxn0_fwd"] = fwd_a(rxn0, pos_a)
    # the max_norm=10 renorm has now been applied in place to the looked-up rows (schnet.py:152)
    manifests["dualenc_qm9_default_seed0_after_fwd"] = {
        "encoder_global.node_emb.weight": manifest(ma)["encoder_global.node_emb.weight"]}
    gold["a_rxn0_fwd_wide"] = fwd_a(rxn0, pos_a * 5.0)
    gold["a_syn4_fwd"] = fwd_a(syn4, syn4["pos_init"] * 3.0)
    gold["a_syn4_fwd_wide"] = fwd_a(syn4, syn4["pos_init"] * 6.0)

    def ld_a(g, pos_init, n_steps, seed, **kw):
        gen = torch.Generator().manual_seed(seed)
        noise = torch.randn(n_steps, pos_init.size(0), 3, generator=gen)
        with injected_noise(noise), quiet():
            pos, traj = ma.langevin_dynamics_sample(
                g["atom_type"], pos_init, g["bond_index"], g["bond_type"], g["batch"], g["num_graphs"],
                extend_order=True, n_steps=n_steps, step_lr=1e-7, sampling_type="ld", **kw)
        out = {"pos_init": pos_init, "noise": noise, "pos": pos, "traj": torch.stack(traj)}
        out.update({k: torch.tensor(float(v)) for k, v in kw.items()})
        return out

    gold["a_rxn0_ld10"] = ld_a(rxn0, pos_a, 10, 21, clip=10.0, clip_local=10.0)
    gold["a_syn4_ld5"] = ld_a(syn4, syn4["pos_init"], 5, 22, clip=10.0, clip_local=10.0, w_global=0.5)

    torch.save(gold, os.path.join(OUT, "golden_outputs.pt"))
    with open(os.path.join(OUT, "state_dict_manifests.json"), "w") as f:
        json.dump(manifests, f, indent=0, sort_keys=False)
    with open(os.path.join(OUT, "schedule.json"), "w") as f:
        sig = ((1.0 - mb.alphas).sqrt() / mb.alphas.sqrt())
        json.dump({"sigma_0": float(sig[0]), "sigma_last": float(sig[-1]), "sigma_2500": float(sig[2500]),
                   "beta_0": float(mb.betas[0]), "beta_last": float(mb.betas[-1]),
                   "alpha_last": float(mb.alphas[-1])}, f)
    size = sum(os.path.getsize(os.path.join(OUT, n)) for n in os.listdir(OUT))
    print("golden written: %d cases, %.1f KB total" % (len(gold), size / 1024))
    print("known answer b_rxn0_fwd: mean %.8f std %.8f E %d" % (
        gold["b_rxn0_fwd"]["edge_inv"].mean(), gold["b_rxn0_fwd"]["edge_inv"].std(),
        gold["b_rxn0_fwd"]["edge_inv"].numel()))


if __name__ == "__main__":
    main()
