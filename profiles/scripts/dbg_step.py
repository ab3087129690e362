"""One eps-net evaluation of the batch-100 workload with TSD_GEMM_DBG=1: prints CTA-0 timelines of every tf32 GEMM."""
import sys, torch
sys.path.insert(0, '.')
from tsdiff_b200.synthetic import make_batch
from tsdiff_b200.config import TRAIN_CONFIG_MODEL
from tsdiff_b200.models.epsnet import get_model
dev = 'cuda:0'
g = make_batch(100, seed=1000)
torch.manual_seed(0)
m = get_model(TRAIN_CONFIG_MODEL).to(dev); m.math = 'tf32'
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()}
pos = (d['pos_init'] * 3.0).contiguous()
for it in range(2):
    print("==== pass", it, file=sys.stderr)
    m(d['atom_type'], d['r_feat'], d['p_feat'], pos, d['bond_index'], d['bond_type'], d['batch'], None)
    torch.cuda.synchronize()
