"""Golden vectors for the non-default branches of EnsembleSampler.dynamic_sampling (sampler.py:118-257),
produced by the reference's OWN Python like make_golden.py (same container requirements):

    python tests/golden/make_golden_ddpm.py          -> tests/golden/golden_ddpm.pt

Cases (path B, random-init seed 0, injected per-step noise so every implementation consumes the
same stream): the `ddpm` update from the default start, `ddpm` with the zero-noise start
(denoise_from_time_t), the from_ts_guess noising start (noise_from_time_t; its single torch.randn
draw is recorded) followed by `ld`, and `ld` with clip_pos.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, injected_noise, load_rxn0, quiet, rh  # noqa: E402
from tsdiff_b200.synthetic import make_batch  # noqa: E402


def main():
    epsnet, sampler, _, _ = rh.import_reference()
    cfg_b = rh.load_yaml_config("configs/train_config.yml").model
    torch.manual_seed(0)
    mb = epsnet.get_model(cfg_b)
    mb.eval()
    rxn0 = load_rxn0()
    syn4 = make_batch(4, seed=3, sizes=[10, 17, 25, 12])
    torch.manual_seed(2022)
    pos_a = torch.randn(13, 3)

    def run(g, pos_init, n_steps, seed, init_seed=None, **kw):
        gen = torch.Generator().manual_seed(seed)
        noise = torch.randn(n_steps, pos_init.size(0), 3, generator=gen)
        out = {"pos_init": pos_init, "noise": noise}
        if init_seed is not None:  # sampler.py:155 draws torch.randn(pos_init.size()) from the global generator
            torch.manual_seed(init_seed)
            out["init_noise"] = torch.randn(pos_init.size())
            torch.manual_seed(init_seed)
        ens = sampler.EnsembleSampler([mb])
        with injected_noise(noise), quiet():
            pos, traj = ens.dynamic_sampling(
                g["atom_type"], g["r_feat"], g["p_feat"], pos_init, g["bond_index"], g["bond_type"], g["batch"],
                g["num_graphs"], extend_order=True, n_steps=n_steps, step_lr=1e-7, clip=1000, **kw)
        out.update({"pos": pos, "traj": torch.stack(traj)})
        return out

    gold = {
        "b_rxn0_ddpm20": run(rxn0, pos_a, 20, 31, sampling_type="ddpm"),
        "b_syn4_ddpm10": run(syn4, syn4["pos_init"], 10, 32, sampling_type="ddpm"),
        # the last 12 time indices: exercises t == 0 (mask = 0, atm1 = 1)
        "b_rxn0_ddpm_t12": run(rxn0, pos_a * 0.5, 12, 33, sampling_type="ddpm", denoise_from_time_t=12),
        "b_rxn0_guess_ld8": run(rxn0, pos_a, 8, 34, init_seed=35, sampling_type="ld", denoise_from_time_t=3000,
                                noise_from_time_t=1500),
        "b_rxn0_guess_ddpm8": run(rxn0, pos_a, 8, 36, init_seed=37, sampling_type="ddpm", denoise_from_time_t=600,
                                  noise_from_time_t=0),
        "b_syn4_ld6_clip_pos": run(syn4, syn4["pos_init"], 6, 38, sampling_type="ld", clip_pos=20.0),
    }
    torch.save(gold, os.path.join(OUT, "golden_ddpm.pt"))
    for k, v in gold.items():
        print(k, tuple(v["traj"].shape), "final |pos| max %.4f" % float(v["pos"].abs().max()))


if __name__ == "__main__":
    main()
