"""DualEncoderEpsNetwork (path A) -- host-side mirror of models/epsnet/dualenc.py:62-374 and
its Langevin sampler :687-967 (`type: diffusion`, sampling_type 'ld').  Same constructor,
attribute names / state_dict layout, `forward` and `langevin_dynamics_sample` signatures;
compute is in the CUDA library."""
import os

import numpy as np
import torch
from torch import nn

from ... import engine as E
from ..layers import (GINEncoder, Marker, MultiLayerPerceptron, SchNetEncoder, activation_name, get_edge_encoder,
                      schedule_parameters, NUM_BOND_TYPES)
from ._cache import EngineCache


class DualEncoderEpsNetwork(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.edge_encoder_global = get_edge_encoder(config)
        self.edge_encoder_local = get_edge_encoder(config)
        self.encoder_global = SchNetEncoder(
            hidden_channels=config.hidden_dim, num_filters=config.hidden_dim, num_interactions=config.num_convs,
            edge_channels=self.edge_encoder_global.out_channels, cutoff=config.cutoff, smooth=config.smooth_conv,
            embedding=True)
        self.encoder_local = GINEncoder(hidden_dim=config.hidden_dim, num_convs=config.num_convs_local,
                                        embedding=True)
        self.grad_global_dist_mlp = MultiLayerPerceptron(
            2 * config.hidden_dim, [config.hidden_dim, config.hidden_dim // 2, 1], activation=config.mlp_act)
        self.grad_local_dist_mlp = MultiLayerPerceptron(
            2 * config.hidden_dim, [config.hidden_dim, config.hidden_dim // 2, 1], activation=config.mlp_act)
        self.model_type = config.type
        if self.model_type == "diffusion":
            self.betas, self.alphas = schedule_parameters(config)
            self.num_timesteps = self.betas.size(0)
        elif self.model_type == "dsm":  # denoising score matching, dualenc.py:144-159
            sigmas = torch.tensor(np.exp(np.linspace(np.log(config.sigma_begin), np.log(config.sigma_end),
                                                     config.num_noise_level)), dtype=torch.float32)
            self.sigmas = nn.Parameter(sigmas, requires_grad=False)
            self.num_timesteps = self.sigmas.size(0)
        else:
            raise NotImplementedError("model type %r (dualenc.py:123: 'diffusion' or 'dsm')" % (self.model_type,))
        self.TS = config.TS if hasattr(config, "TS") else False
        self.num_bond_types = NUM_BOND_TYPES
        global_modules = [self.edge_encoder_global, self.encoder_global, self.grad_global_dist_mlp]
        local_modules = [self.edge_encoder_local, self.encoder_local, self.grad_local_dist_mlp]
        if self.TS:
            ch = self.edge_encoder_global.out_channels
            act = activation_name(config.edge_cat_act)
            self.edge_cat_global = nn.Sequential(nn.Linear(ch * 2, ch), Marker(act), nn.Linear(ch, ch))
            self.edge_cat_local = nn.Sequential(nn.Linear(ch * 2, ch), Marker(act), nn.Linear(ch, ch))
            global_modules.append(self.edge_cat_global)
            local_modules.append(self.edge_cat_local)
        self.model_global = nn.ModuleList(global_modules)
        self.model_local = nn.ModuleList(local_modules)
        self.math = os.environ.get("TSDIFF_B200_MATH", "fp32")
        self._cache = EngineCache()

    def _engine(self, atom_type, bond_index, bond_type, batch, extend_order=True, extend_radius=True):
        E.require_cuda_inputs(batch=batch, atom_type=atom_type)
        with torch.cuda.device(batch.device):
            return self._cache.get((atom_type, bond_index, bond_type, batch),
                                   (self.math, bool(extend_order), bool(extend_radius)),
                                   lambda: E.DualScoreEngine(self, atom_type, bond_index, bond_type, batch, math=self.math,
                                                             extend_order=extend_order, extend_radius=extend_radius),
                                   modules=(self,))

    def _sigma_edge_inverse(self, time_step, batch, rows):
        """dsm: 1 / sigma_edge for the edges whose first atom is `rows` (dualenc.py:247-259)."""
        if time_step is None:
            raise ValueError("model type 'dsm' conditions on the noise level: time_step (G,) is required")
        noise_levels = self.sigmas.to(batch.device).index_select(0, time_step.to(batch.device))
        return (1.0 / noise_levels.index_select(0, batch.index_select(0, rows))).contiguous()

    @staticmethod
    def _row_scale(x, s):
        """x (E,1) * s (E,): tsd_row_scale (the product of dualenc.py:308-309 / :353-361)."""
        import ctypes as C
        from ... import _lib as L
        x = x.contiguous()
        out = torch.empty_like(x)
        L.check(L.load().tsd_row_scale(x.size(0), 1, L.ptr(x), L.ptr(s), L.ptr(out),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tsd_row_scale")
        return out

    @torch.no_grad()
    def forward(self, atom_type, pos, bond_index, bond_type, batch, time_step=None, edge_index=None,
                edge_type=None, edge_length=None, return_edges=False, extend_order=True, extend_radius=True,
                is_sidechain=None):
        """dualenc.py:206-374.  Returns (edge_inv_global (E,1), edge_inv_local (E_l,1)) and, with
        return_edges, (edge_index, edge_type, edge_length, local_edge_mask)."""
        if edge_index is not None or edge_type is not None or edge_length is not None:
            raise NotImplementedError("precomputed edges: the CUDA path always rebuilds the graph (dualenc.py:230)")
        if is_sidechain is not None:
            raise NotImplementedError("is_sidechain (the protein side-chain radius graph, common.py:345-366) is not built")
        eng = self._engine(atom_type, bond_index, bond_type, batch, extend_order, extend_radius)
        eng.refresh_embeddings()  # nn.Embedding(max_norm) renormalises on every lookup
        eng.evaluate(pos.detach().to(torch.float32).contiguous())
        plan = eng.plan
        e = plan.edge_count()
        etype = plan.tab0[:e].long()
        local = etype > 0
        inv_g = eng.directed(eng.edge_inv_global, e).unsqueeze(-1).clone()
        inv_l = eng.directed(eng.edge_inv_local, e)[local].unsqueeze(-1)
        if self.model_type == "dsm":
            rows = plan.row[:e].long()
            inv_g = self._row_scale(inv_g, self._sigma_edge_inverse(time_step, batch, rows))
            inv_l = self._row_scale(inv_l, self._sigma_edge_inverse(time_step, batch, rows[local]))
        if not return_edges:
            return inv_g, inv_l
        edge_index = torch.stack([plan.row[:e], plan.col[:e]], dim=0).long()
        return inv_g, inv_l, edge_index, etype, plan.length[:e].unsqueeze(-1).clone(), local

    def get_loss(self, atom_type, pos, bond_index, bond_type, batch, num_nodes_per_graph=None, num_graphs=None,
                 anneal_power=2.0, return_unreduced_loss=False, return_unreduced_edge_loss=False, extend_order=True,
                 extend_radius=True, is_sidechain=None, time_step=None, pos_noise=None):
        """dualenc.py:376-562 (model_type 'diffusion'), FORWARD VALUE ONLY: raises NotImplementedError when
        gradients are enabled (the backward kernels are not built).  Returns loss (N,1), or
        (loss, loss_global, loss_local) with return_unreduced_loss.  Keyword-only extras: time_step (G,)
        and pos_noise (N,3) replace the reference's draws (:441-451)."""
        E.require_no_grad(self, "DualEncoderEpsNetwork.get_loss")
        if is_sidechain is not None:
            raise NotImplementedError("is_sidechain is not built")
        if self.model_type == "dsm":
            return self._get_loss_dsm(atom_type, pos, bond_index, bond_type, batch, num_graphs, anneal_power,
                                      return_unreduced_loss, return_unreduced_edge_loss, extend_order, extend_radius,
                                      time_step, pos_noise)
        with torch.no_grad():
            dev = pos.device
            if num_graphs is None:
                num_graphs = int(batch.max().item()) + 1
            if time_step is None:
                half = torch.randint(0, self.num_timesteps, size=(num_graphs // 2 + 1,), device=dev)
                time_step = torch.cat([half, self.num_timesteps - half - 1], dim=0)[:num_graphs]
            if pos_noise is None:
                pos_noise = torch.zeros(size=pos.size(), device=dev)
                pos_noise.normal_()
            a = self.alphas.to(dev).index_select(0, time_step.to(dev))
            a_pos = a.index_select(0, batch).unsqueeze(-1)
            pos_perturbed = (pos + pos_noise.to(dev) * (1.0 - a_pos).sqrt() / a_pos.sqrt()).to(torch.float32).contiguous()
            inv_g, inv_l, edge_index, _, edge_length, local = self(atom_type, pos_perturbed, bond_index, bond_type, batch,
                                                                   time_step, return_edges=True, extend_order=extend_order,
                                                                   extend_radius=extend_radius)
            plan = self._engine(atom_type, bond_index, bond_type, batch, extend_order, extend_radius).plan
            a_edge = a.index_select(0, batch.index_select(0, edge_index[0])).unsqueeze(-1)
            d_gt = (pos[edge_index[0]] - pos[edge_index[1]]).norm(dim=-1).unsqueeze(-1)
            d_target = (d_gt - edge_length) / (1.0 - a_edge).sqrt() * a_edge.sqrt()
            global_mask = torch.logical_and(torch.logical_or(edge_length <= self.config.cutoff, local.unsqueeze(-1)),
                                            ~local.unsqueeze(-1))
            zero = torch.zeros_like(d_target)
            tgt_g = E.eq_transform_directed(plan, pos_perturbed, torch.where(global_mask, d_target, zero))
            eq_g = E.eq_transform_directed(plan, pos_perturbed, torch.where(global_mask, inv_g, zero))
            loss_global = torch.sum((eq_g - tgt_g) ** 2, dim=-1, keepdim=True)
            loc = local.unsqueeze(-1)
            inv_l_full = torch.zeros_like(d_target)
            inv_l_full[local] = inv_l
            tgt_l = E.eq_transform_directed(plan, pos_perturbed, torch.where(loc, d_target, zero))
            eq_l = E.eq_transform_directed(plan, pos_perturbed, inv_l_full)
            loss_local = torch.sum((eq_l - tgt_l) ** 2, dim=-1, keepdim=True)
            loss = (2 * loss_global + 5 * loss_local) / 7
            if return_unreduced_edge_loss:
                return None  # the reference's branch is `pass` (dualenc.py:556-557)
            if return_unreduced_loss:
                return loss, loss_global, loss_local
            return loss

    def _get_loss_dsm(self, atom_type, pos, bond_index, bond_type, batch, num_graphs, anneal_power, return_unreduced_loss,
                      return_unreduced_edge_loss, extend_order, extend_radius, time_step, pos_noise):
        """dualenc.py:969-1100 (forward value): noise levels sigma instead of alphas, target (d_gt - d) / sigma^2,
        loss weighted by sigma^anneal_power; loss = 2 * 0.5 * global + 5 * 0.5 * local."""
        with torch.no_grad():
            dev = pos.device
            if num_graphs is None:
                num_graphs = int(batch.max().item()) + 1
            if time_step is None:
                half = torch.randint(0, self.num_timesteps, size=(num_graphs // 2 + 1,), device=dev)
                time_step = torch.cat([half, self.num_timesteps - half - 1], dim=0)[:num_graphs]
            if pos_noise is None:
                pos_noise = torch.zeros(size=pos.size(), device=dev)
                pos_noise.normal_()
            noise_levels = self.sigmas.to(dev).index_select(0, time_step.to(dev))
            sigmas_pos = noise_levels.index_select(0, batch).unsqueeze(-1)
            pos_perturbed = (pos + pos_noise.to(dev) * sigmas_pos).to(torch.float32).contiguous()
            inv_g, inv_l, edge_index, _, edge_length, local = self(atom_type, pos_perturbed, bond_index, bond_type, batch,
                                                                   time_step, return_edges=True, extend_order=extend_order,
                                                                   extend_radius=extend_radius)
            plan = self._engine(atom_type, bond_index, bond_type, batch, extend_order, extend_radius).plan
            sigmas_edge = noise_levels.index_select(0, batch.index_select(0, edge_index[0])).unsqueeze(-1)
            d_gt = (pos[edge_index[0]] - pos[edge_index[1]]).norm(dim=-1).unsqueeze(-1)
            d_target = 1.0 / (sigmas_edge ** 2) * (d_gt - edge_length)
            loc = local.unsqueeze(-1)
            global_mask = torch.logical_and(torch.logical_or(edge_length <= self.config.cutoff, loc), ~loc)
            zero = torch.zeros_like(d_target)
            tgt_g = E.eq_transform_directed(plan, pos_perturbed, torch.where(global_mask, d_target, zero))
            eq_g = E.eq_transform_directed(plan, pos_perturbed, torch.where(global_mask, inv_g, zero))
            loss_global = 2 * torch.sum(0.5 * ((eq_g - tgt_g) ** 2) * (sigmas_pos ** anneal_power), dim=-1, keepdim=True)
            inv_l_full = torch.zeros_like(d_target)
            inv_l_full[local] = inv_l
            tgt_l = E.eq_transform_directed(plan, pos_perturbed, torch.where(loc, d_target, zero))
            eq_l = E.eq_transform_directed(plan, pos_perturbed, inv_l_full)
            loss_local = 5 * torch.sum(0.5 * ((eq_l - tgt_l) ** 2) * (sigmas_pos ** anneal_power), dim=-1, keepdim=True)
            loss = loss_global + loss_local
            if return_unreduced_edge_loss:
                return None  # the reference's branch is `pass` (dualenc.py:1094-1095)
            if return_unreduced_loss:
                return loss, loss_global, loss_local
            return loss

    def langevin_dynamics_sample(self, atom_type, pos_init, bond_index, bond_type, batch, num_graphs, extend_order,
                                 extend_radius=True, n_steps=100, step_lr=0.0000010, clip=1000, clip_local=None,
                                 clip_pos=None, min_sigma=0, is_sidechain=None, global_start_sigma=float("inf"),
                                 w_global=0.2, w_reg=1.0, **kwargs):
        """dualenc.py:687-967: sampling_type 'ld' (:946-952) and 'ddpm_noisy' (the default),
        'ddpm_det', 'generalized' with eta= (:861-944).  Extra keyword-only knobs that do not exist
        in the reference: noise= (n_steps,N,3) tensor replacing torch.randn_like, seed= Philox seed,
        keep_traj=, atom_offset= (global index of this shard's first atom)."""
        from ... import _lib as L
        if is_sidechain is not None:
            raise NotImplementedError("is_sidechain is not built")
        eng = self._engine(atom_type, bond_index, bond_type, batch, extend_order, extend_radius)
        eng.refresh_embeddings()
        if self.model_type == "dsm":
            # dualenc.py:1102-1203: annealed Langevin dynamics -- n_steps updates at EVERY noise level sigma >= min_sigma
            sched = E.dsm_schedule(self.sigmas, n_steps, step_lr, min_sigma, global_start_sigma)
            pos = pos_init.detach().to(torch.float32).contiguous().clone()
            ch0, ch1 = eng.score_channels(clip, clip_local, w_global)
            runner = E.LangevinRunner(eng, ch0, ch1, sched, pos, noise=kwargs.get("noise"),
                                      seed=E.resolve_seed(kwargs.get("seed")), atom_offset=kwargs.get("atom_offset", 0),
                                      clip_pos=clip_pos, keep_traj=kwargs.get("keep_traj", True),
                                      use_graph=kwargs.get("use_graph", True), rule=L.RULE_DSM)
            pos = runner.run()
            return pos, (list(runner.traj_cpu().unbind(0)) if runner.traj is not None else [])
        sampling_type = kwargs.get("sampling_type", "ddpm_noisy")
        if sampling_type not in ("ld", "ddpm_noisy", "ddpm_det", "generalized"):
            raise NotImplementedError("sampling_type %r (dualenc.py:861-952 has ld, ddpm_noisy, ddpm_det, generalized)"
                                      % (sampling_type,))
        sched, sigmas = E.ld_schedule(self.alphas, n_steps, step_lr, global_start_sigma)
        rule = L.RULE_LD
        if sampling_type != "ld":
            sched = E.dualenc_branch_schedule(self.alphas, self.betas, n_steps, step_lr, sampling_type,
                                              eta=kwargs.get("eta", 1.0), global_start_sigma=global_start_sigma)
            rule = L.RULE_GENERALIZED if sampling_type == "generalized" else L.RULE_DDPM_DUALENC
        pos = (pos_init.detach().to(torch.float32) * sigmas[-1].to(pos_init.device)).contiguous()
        ch0, ch1 = eng.score_channels(clip, clip_local, w_global)
        runner = E.LangevinRunner(eng, ch0, ch1, sched, pos, noise=kwargs.get("noise"),
                                  seed=E.resolve_seed(kwargs.get("seed")),
                                  atom_offset=kwargs.get("atom_offset", 0), clip_pos=clip_pos,
                                  keep_traj=kwargs.get("keep_traj", True), use_graph=kwargs.get("use_graph", True),
                                  rule=rule)
        pos = runner.run()
        traj = list(runner.traj_cpu().unbind(0)) if runner.traj is not None else []
        return pos, traj
