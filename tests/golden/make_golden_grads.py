"""Reference gradients of the training loss (train.py:128-142: loss = get_loss(...).mean(); loss.backward()) for
the backward kernels of SURVEY.md section 8(f)-2, produced by the reference's OWN Python + autograd with the
draws of tests/golden/golden_loss.pt.  The full gradient is 2.8 M floats; the fixture keeps, per parameter,
(sum, abs-sum, L2 norm) in float64 and the full gradient of three small tensors:

    python tests/golden/make_golden_grads.py       -> tests/golden/golden_grads.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, rh  # noqa: E402
from tsdiff_b200.synthetic import make_batch  # noqa: E402

FULL = ("grad_dist_mlp.layers.2.weight", "edge_encoder.mlp.layers.0.weight", "encoder.interactions.6.conv.lin2.bias")


def main():
    epsnet, _, _, _ = rh.import_reference()
    cfg_b = rh.load_yaml_config("configs/train_config.yml").model
    torch.manual_seed(0)
    mb = epsnet.get_model(cfg_b)
    g = make_batch(4, seed=3, sizes=[10, 17, 25, 12])
    pos = g["pos_init"] * 1.5
    torch.manual_seed(51)  # the draws recorded in golden_loss.pt["b_syn4"]
    loss = mb.get_loss(g["atom_type"], g["r_feat"], g["p_feat"], pos, g["bond_index"], g["bond_type"], g["batch"],
                       g["num_nodes_per_graph"], g["num_graphs"]).mean()
    loss.backward()
    out = {"loss_mean": float(loss), "params": {}, "full": {}}
    for name, p in mb.named_parameters():
        if p.grad is None:
            continue
        gd = p.grad.double()
        out["params"][name] = {"sum": float(gd.sum()), "abs_sum": float(gd.abs().sum()), "norm": float(gd.norm())}
        if name in FULL:
            out["full"][name] = p.grad.flatten().tolist()
    with open(os.path.join(OUT, "golden_grads.json"), "w") as f:
        json.dump(out, f)
    print("loss.mean() = %.6f; %d parameters with gradients; total grad norm %.4f"
          % (out["loss_mean"], len(out["params"]), sum(v["norm"] ** 2 for v in out["params"].values()) ** 0.5))


if __name__ == "__main__":
    main()
