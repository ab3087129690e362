"""Golden vectors for get_loss (condensenc.py:267-328, dualenc.py:425-562), produced by the reference's OWN
Python.  The reference draws the time steps and the noise from torch's global generator; the script
replays the same draws (same seed, same call order) so that implementations can be given them explicitly:

    python tests/golden/make_golden_loss.py        -> tests/golden/golden_loss.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, load_rxn0, rh  # noqa: E402
from tsdiff_b200.synthetic import make_batch  # noqa: E402


def main():
    epsnet, _, _, _ = rh.import_reference()
    cfg_b = rh.load_yaml_config("configs/train_config.yml").model
    cfg_a = rh.load_yaml_config("configs/geodiff_legacy/qm9_default.yml").model
    torch.manual_seed(0)
    mb = epsnet.get_model(cfg_b)
    torch.manual_seed(0)
    ma = epsnet.get_model(cfg_a)
    syn4 = make_batch(4, seed=3, sizes=[10, 17, 25, 12])
    rxn0 = load_rxn0()
    gold = {}
    for name, g, pos, seed in (("syn4", syn4, syn4["pos_init"] * 1.5, 51), ("rxn0", rxn0, None, 52)):
        if pos is None:
            torch.manual_seed(2022)
            pos = torch.randn(13, 3) * 1.5
        G = g["num_graphs"]
        # path B: randint(t0, t1, (G//2+1,)) then randn(pos.size())   (condensenc.py:287-295)
        torch.manual_seed(seed)
        half = torch.randint(0, mb.num_timesteps, size=(G // 2 + 1,))
        t_b = torch.cat([half, mb.num_timesteps - 1 - half], dim=0)[:G]
        z_b = torch.randn(size=pos.size())
        torch.manual_seed(seed)
        with torch.no_grad():
            loss_b = mb.get_loss(g["atom_type"], g["r_feat"], g["p_feat"], pos, g["bond_index"], g["bond_type"],
                                 g["batch"], g["num_nodes_per_graph"], G)
        gold["b_" + name] = {"pos": pos, "time_step": t_b, "pos_noise": z_b, "loss": loss_b}
        # path A: randint(0, T, (G//2+1,)) then zeros(pos.size()).normal_()   (dualenc.py:441-451)
        torch.manual_seed(seed + 100)
        half = torch.randint(0, ma.num_timesteps, size=(G // 2 + 1,))
        t_a = torch.cat([half, ma.num_timesteps - half - 1], dim=0)[:G]
        z_a = torch.zeros(size=pos.size())
        z_a.normal_()
        torch.manual_seed(seed + 100)
        with torch.no_grad():
            loss, lg, ll = ma.get_loss(g["atom_type"], pos, g["bond_index"], g["bond_type"], g["batch"],
                                       g["num_nodes_per_graph"], G, return_unreduced_loss=True)
        gold["a_" + name] = {"pos": pos, "time_step": t_a, "pos_noise": z_a, "loss": loss, "loss_global": lg,
                             "loss_local": ll}
    torch.save(gold, os.path.join(OUT, "golden_loss.pt"))
    for k, v in gold.items():
        print(k, tuple(v["loss"].shape), "mean loss %.5f" % float(v["loss"].mean()), v["time_step"].tolist())


if __name__ == "__main__":
    main()
