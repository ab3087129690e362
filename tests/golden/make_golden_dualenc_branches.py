"""Golden vectors for the non-`ld` branches of DualEncoderEpsNetwork.langevin_dynamics_sample
(dualenc.py:861-944: `ddpm_noisy` -- the class's default --, `ddpm_det`, `generalized`), produced by the
reference's OWN Python like make_golden.py:

    python tests/golden/make_golden_dualenc_branches.py     -> tests/golden/golden_dualenc_branches.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, injected_noise, load_rxn0, quiet, rh  # noqa: E402
from tsdiff_b200.synthetic import make_batch  # noqa: E402


def main():
    epsnet, _, _, _ = rh.import_reference()
    cfg_a = rh.load_yaml_config("configs/geodiff_legacy/qm9_default.yml").model
    torch.manual_seed(0)
    ma = epsnet.get_model(cfg_a)
    ma.eval()
    rxn0 = load_rxn0()
    syn4 = make_batch(4, seed=3, sizes=[10, 17, 25, 12])
    torch.manual_seed(2022)
    pos_a = torch.randn(13, 3)

    def run(g, pos_init, n_steps, seed, **kw):
        gen = torch.Generator().manual_seed(seed)
        noise = torch.randn(n_steps, pos_init.size(0), 3, generator=gen)
        with injected_noise(noise), quiet():
            pos, traj = ma.langevin_dynamics_sample(
                g["atom_type"], pos_init, g["bond_index"], g["bond_type"], g["batch"], g["num_graphs"],
                extend_order=True, n_steps=n_steps, step_lr=1e-7, **kw)
        return {"pos_init": pos_init, "noise": noise, "pos": pos, "traj": torch.stack(traj)}

    gold = {
        "a_rxn0_ddpm_noisy10": run(rxn0, pos_a, 10, 41, clip=10.0, clip_local=10.0, sampling_type="ddpm_noisy"),
        "a_syn4_ddpm_det6": run(syn4, syn4["pos_init"], 6, 42, clip=10.0, clip_local=10.0, sampling_type="ddpm_det"),
        "a_rxn0_generalized8": run(rxn0, pos_a, 8, 43, clip=10.0, clip_local=10.0, sampling_type="generalized", eta=1.0),
        "a_syn4_generalized6_eta05": run(syn4, syn4["pos_init"], 6, 44, clip=10.0, clip_local=10.0, w_global=0.5,
                                         sampling_type="generalized", eta=0.5),
        # the whole schedule in 5000 steps would be too long: a short run over the LAST time indices needs
        # n_steps = T, so t == 0 (mask = 0) is exercised through the oracle-vs-engine table test instead
    }
    torch.save(gold, os.path.join(OUT, "golden_dualenc_branches.pt"))
    for k, v in gold.items():
        print(k, tuple(v["traj"].shape), "final |pos| max %.4f" % float(v["pos"].abs().max()))


if __name__ == "__main__":
    main()
