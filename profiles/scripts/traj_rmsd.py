"""Trajectory-level parity of the tf32 tensor-core mode against the fp32 FFMA mode (same Philox
noise): per-reaction RMSD of the final geometries after a full LD trajectory."""
import sys, torch, time
sys.path.insert(0, '.')
from tsdiff_b200.synthetic import make_batch
from tsdiff_b200.config import TRAIN_CONFIG_MODEL
from tsdiff_b200.models.epsnet import get_model
from tsdiff_b200.models.sampler import EnsembleSampler
dev = 'cuda:0'
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
g = make_batch(100, seed=1000)
torch.manual_seed(0)
m = get_model(TRAIN_CONFIG_MODEL).to(dev)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()}
out = {}
for math in ("fp32", "tf32"):
    m.math = math
    ens = EnsembleSampler([m])
    t0 = time.time()
    pos, _ = ens.dynamic_sampling(d['atom_type'], d['r_feat'], d['p_feat'], d['pos_init'], d['bond_index'], d['bond_type'],
                                  d['batch'], 100, extend_order=True, n_steps=n_steps, step_lr=1e-7, clip=1000,
                                  sampling_type='ld', seed=2022, keep_traj=False)
    torch.cuda.synchronize()
    out[math] = pos.cpu()
    print(math, "seconds", time.time() - t0, "max |pos|", float(pos.abs().max()))
diff = (out['tf32'] - out['fp32'])
b = g['batch']
sq = torch.zeros(100).index_add_(0, b, (diff ** 2).sum(1))
rmsd = (sq / g['num_nodes_per_graph']).sqrt()
print("steps", n_steps, "RMSD tf32 vs fp32 per reaction: mean %.3e max %.3e median %.3e" % (rmsd.mean(), rmsd.max(), rmsd.median()))
