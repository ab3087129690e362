"""EnsembleSampler -- host-side mirror of models/sampler.py:44-257: same constructor,
attributes (`models`, `config`, `alphas`, `betas`, `num_timesteps`), `forward` and
`dynamic_sampling` signatures and return contract.  The Langevin loop (`sampling_type='ld'`,
sampler.py:238-244) runs as a replayed CUDA graph with no per-step host work."""
import torch
from torch import nn

from .. import engine as E
from .epsnet._cache import EngineCache
from .epsnet.condensenc import condensed_outputs


class EnsembleSampler(nn.Module):
    def __init__(self, models):
        super().__init__()
        self.models = models
        self.config = models[0].config
        self.alphas = models[0].alphas
        self.betas = models[0].betas
        self.num_timesteps = models[0].num_timesteps
        self.math = getattr(models[0], "math", "fp32")
        self._cache = EngineCache()

    def _engine(self, atom_type, r_feat, p_feat, bond_index, bond_type, batch):
        E.require_cuda_inputs(batch=batch, atom_type=atom_type)
        with torch.cuda.device(batch.device):
            return self._cache.get(
                (atom_type, r_feat, p_feat, bond_index, bond_type, batch), (self.math, len(self.models)),
                lambda: E.CondensedScoreEngine(self.models, atom_type, r_feat, p_feat, bond_index, bond_type, batch,
                                               math=self.math), modules=tuple(self.models))

    @torch.no_grad()
    def forward(self, atom_type, r_feat, p_feat, pos, bond_index, bond_type, batch, time_step=None,
                return_edges=True, **kwargs):
        """sampler.py:58-116: mean of edge_inv over the members; edges from member 0."""
        eng = self._engine(atom_type, r_feat, p_feat, bond_index, bond_type, batch)
        eng.evaluate(pos.detach().to(torch.float32).contiguous())
        return condensed_outputs(eng, return_edges, divide=len(self.models))

    def dynamic_sampling(self, atom_type, r_feat, p_feat, pos_init, bond_index, bond_type, batch, num_graphs,
                         extend_order, extend_radius=True, n_steps=100, step_lr=0.0000010, clip=1000, clip_pos=None,
                         denoise_from_time_t=None, noise_from_time_t=None, **kwargs):
        """sampler.py:118-257: the `ld` (:238-244) and `ddpm` (:215-236, the reference's default)
        updates and the three starts (:149-182: from_ts_guess noising, zero-noise start, default).
        Returns (pos on device, list of n_steps CPU (N,3) tensors).  Raises FloatingPointError when
        a NaN position appears.  Keyword-only extras absent from the reference: noise=
        (n_steps,N,3) tensor used instead of torch.randn_like; init_noise= (N,3) tensor used
        instead of the torch.randn draw of the from_ts_guess start; seed= Philox seed (default: drawn
        from torch's global generator on every call, like the reference's randn_like consumes it); keep_traj=; atom_offset= global index of the first atom of this
        shard; use_graph=; ensemble_group= a torch.distributed process group whose ranks each hold a
        DIFFERENT subset of the ensemble members (`self.models`) and the SAME batch: the per-atom
        scores are all-reduced every step, which reproduces the reference's per-step mean over all
        members (sampler.py:96-111) with one member per GPU (BASELINE config 3).  The exchange is fused into the
        update kernel (peer-memory stores over NVLink, engine.PeerExchange); ensemble_exchange="nccl" selects an
        NCCL all-reduce captured in the step graph instead."""
        from .. import _lib as L
        sampling_type = kwargs.get("sampling_type", "ddpm")
        if sampling_type not in ("ld", "ddpm"):
            raise NotImplementedError("sampling_type %r: 'ld' and 'ddpm' are built (SURVEY.md 8(f)-3)" % (sampling_type,))
        eng = self._engine(atom_type, r_feat, p_feat, bond_index, bond_type, batch)
        t_end = self.num_timesteps if denoise_from_time_t is None else int(denoise_from_time_t)
        assert t_end >= n_steps
        sched, sigmas = E.ld_schedule(self.alphas[:t_end], n_steps, step_lr)
        pos = pos_init.detach().to(torch.float32)
        if noise_from_time_t is not None:  # sampler.py:149-161
            assert denoise_from_time_t is not None and denoise_from_time_t >= noise_from_time_t >= 0
            z0 = kwargs.get("init_noise")
            if z0 is None:
                z0 = torch.randn(pos_init.size(), device=pos_init.device)
            alphas = self.alphas.to(pos.device)
            alpha_t = alphas[denoise_from_time_t - 1]
            alpha_s = alphas[noise_from_time_t - 1] if noise_from_time_t != 0 else 1
            pos = pos + z0.to(pos) * ((1.0 - (alpha_t / alpha_s)) / alpha_t).sqrt()
        elif denoise_from_time_t is None:
            pos = pos * sigmas[-1].to(pos.device)  # sampler.py:182
        pos = pos.contiguous().clone()
        rule = L.RULE_LD
        if sampling_type == "ddpm":
            sched, rule = E.ddpm_schedule(self.betas, t_end, n_steps), L.RULE_DDPM
        ch0, ch1 = eng.score_channels(clip)
        reduce, ensemble_size, exchange = None, None, None
        group = kwargs.get("ensemble_group")
        if group is not None:
            import torch.distributed as dist
            count = torch.tensor([len(self.models)], dtype=torch.int64, device=pos.device)
            dist.all_reduce(count, group=group)
            ensemble_size = int(count.item())
            if kwargs.get("ensemble_exchange", "fused") == "nccl" or dist.get_world_size(group) > L.MAX_EXCHANGE_RANKS:
                reduce = lambda t: dist.all_reduce(t, group=group)  # noqa: E731
            else:
                exchange = E.PeerExchange(eng.plan, group)
        runner = E.LangevinRunner(eng, ch0, ch1, sched, pos, reduce=reduce, ensemble_size=ensemble_size, exchange=exchange,
                                  noise=kwargs.get("noise"),
                                  seed=E.resolve_seed(kwargs.get("seed")),
                                  atom_offset=kwargs.get("atom_offset", 0), clip_pos=clip_pos,
                                  keep_traj=kwargs.get("keep_traj", True), use_graph=kwargs.get("use_graph", True),
                                  rule=rule)
        pos = runner.run()
        traj = list(runner.traj_cpu().unbind(0)) if runner.traj is not None else []
        if exchange is not None:
            del runner
            exchange.close()
        return pos, traj
