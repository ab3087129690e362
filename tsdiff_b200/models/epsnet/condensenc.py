"""CondenseEncoderEpsNetwork (path B) -- host-side mirror of the reference class
models/epsnet/condensenc.py:47-328: same constructor argument, attribute names, parameter
registration order (=> identical state_dict keys and seeded init), same `forward`
signature and return contract.  The arithmetic runs in the CUDA library."""
import os

import torch
from torch import nn

from ... import engine as E
from ..layers import (MultiLayerPerceptron, Marker, activation_name, get_edge_encoder, load_encoder,
                      schedule_parameters, NUM_BOND_TYPES)
from ._cache import EngineCache


class CondenseEncoderEpsNetwork(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.edge_encoder = get_edge_encoder(config)
        assert config.hidden_dim % 2 == 0
        self.atom_embedding = nn.Embedding(100, config.hidden_dim // 2)
        self.atom_feat_embedding = nn.Linear(config.feat_dim, config.hidden_dim // 2, bias=False)
        self.encoder = load_encoder(config, "encoder")
        self.grad_dist_mlp = MultiLayerPerceptron(
            2 * config.hidden_dim, [config.hidden_dim, config.hidden_dim // 2, 1], activation=config.mlp_act)
        # aliases registered a second time, exactly like condensenc.py:81-89 (state_dict carries
        # `model_embedding.*` / `model.*` duplicates of the same tensors)
        self.model_embedding = nn.ModuleList([self.atom_embedding, self.atom_feat_embedding])
        self.model = nn.ModuleList([self.edge_encoder, self.encoder, self.grad_dist_mlp])
        self.betas, self.alphas = schedule_parameters(config)
        self.num_timesteps = self.betas.size(0)
        self.num_bond_types = NUM_BOND_TYPES
        ch = self.edge_encoder.out_channels
        self.edge_cat = nn.Sequential(nn.Linear(ch * 2, ch), Marker(activation_name(config.edge_cat_act)),
                                      nn.Linear(ch, ch))
        self.math = os.environ.get("TSDIFF_B200_MATH", "fp32")
        self._cache = EngineCache()

    def _engine(self, atom_type, r_feat, p_feat, bond_index, bond_type, batch):
        E.require_cuda_inputs(batch=batch, atom_type=atom_type)
        with torch.cuda.device(batch.device):
            return self._cache.get(
                (atom_type, r_feat, p_feat, bond_index, bond_type, batch), (self.math,),
                lambda: E.CondensedScoreEngine([self], atom_type, r_feat, p_feat, bond_index, bond_type, batch,
                                               math=self.math), modules=(self,))

    @torch.no_grad()
    def forward(self, atom_type, r_feat, p_feat, pos, bond_index, bond_type, batch, time_step=None,
                return_edges=True, **kwargs):
        """condensenc.py:241-265.  `time_step` is accepted and ignored, as in the reference
        (condensenc.py:250-258).  Returns (edge_inv (E,1), edge_index (2,E) int64 row-major
        sorted, edge_length (E,1)) on the pred_edge_order graph."""
        eng = self._engine(atom_type, r_feat, p_feat, bond_index, bond_type, batch)
        pos = pos.detach().to(torch.float32).contiguous()
        eng.evaluate(pos)
        return condensed_outputs(eng, return_edges)

    def get_loss(self, atom_type, r_feat, p_feat, pos, bond_index, bond_type, batch, num_nodes_per_graph=None,
                 num_graphs=None, anneal_power=2.0, extend_order=True, extend_radius=True, time_step=None,
                 pos_noise=None):
        """condensenc.py:267-328.  Per-atom loss (N,1).  With gradients enabled (train.py:124-152) the loss carries
        the autograd graph over the parameters: forward and backward of every operator are kernels of the CUDA
        library (tsdiff_b200/training.py, fp32).  Under torch.no_grad() (the validation loop, train.py:150-175) the
        fused sampling kernels evaluate it.  Keyword-only extras: time_step (G,) and pos_noise (N,3) replace the
        reference's torch.randint / torch.randn draws (:287-295)."""
        dev = pos.device
        if num_graphs is None:
            num_graphs = int(batch.max().item()) + 1
        if time_step is None:
            t0, t1 = self.config.get("t0", 0), self.config.get("t1", self.num_timesteps)
            half_1 = torch.randint(t0, t1, size=(num_graphs // 2 + 1,), device=dev)
            time_step = torch.cat([half_1, t0 + t1 - 1 - half_1], dim=0)[:num_graphs]
        if pos_noise is None:
            pos_noise = torch.randn(size=pos.size(), device=dev)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from ... import training
            return training.condensed_loss(self, atom_type, r_feat, p_feat, pos, bond_index, bond_type, batch, time_step,
                                           pos_noise)
        with torch.no_grad():
            a = self.alphas.to(dev).index_select(0, time_step.to(dev))
            a_pos = a.index_select(0, batch).unsqueeze(-1)
            pos_perturbed = (pos + pos_noise.to(dev) * (1.0 - a_pos).sqrt() / a_pos.sqrt()).to(torch.float32).contiguous()
            edge_inv, edge_index, edge_length = self(atom_type, r_feat, p_feat, pos_perturbed, bond_index, bond_type,
                                                     batch, time_step)
            eng = self._engine(atom_type, r_feat, p_feat, bond_index, bond_type, batch)
            plan = eng.plan
            e = plan.edge_count()
            sel = plan.in_b[:e].bool() if eng.two_graphs else torch.ones(e, dtype=torch.bool, device=dev)
            a_edge = a.index_select(0, batch.index_select(0, edge_index[0])).unsqueeze(-1)
            d_gt = (pos[edge_index[0]] - pos[edge_index[1]]).norm(dim=-1).unsqueeze(-1)
            d_target = (d_gt - edge_length) / (1.0 - a_edge).sqrt() * a_edge.sqrt()

            def on_plan_edges(values):  # results live on the pred_edge_order edges: back to plan order
                full = torch.zeros(e, dtype=torch.float32, device=dev)
                full[sel] = values.reshape(-1)
                return full

            node_eq = E.eq_transform_directed(plan, pos_perturbed, on_plan_edges(edge_inv))
            pos_target = E.eq_transform_directed(plan, pos_perturbed, on_plan_edges(d_target))
            return torch.sum((node_eq - pos_target) ** 2, dim=-1, keepdim=True)


def condensed_outputs(eng, return_edges=True, divide=1):
    """Compacts the capacity-sized device buffers into the reference's return contract."""
    plan = eng.plan
    e = plan.edge_count()
    sel = plan.in_b[:e].bool() if eng.two_graphs else torch.ones(e, dtype=torch.bool, device=plan.device)
    edge_inv = eng.edge_inv_directed(e)[sel].unsqueeze(-1)
    if divide != 1:
        edge_inv = edge_inv / divide
    if not return_edges:
        return edge_inv
    edge_index = torch.stack([plan.row[:e][sel], plan.col[:e][sel]], dim=0).long()
    return edge_inv, edge_index, plan.length[:e][sel].unsqueeze(-1)
