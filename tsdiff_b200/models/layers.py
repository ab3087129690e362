"""Parameter containers of the two eps-networks.

These modules only OWN parameters: they reproduce the reference's attribute names (hence
its checkpoint `state_dict` keys) and its construction order (hence
its RNG consumption, so seeded random-init weights are identical to the reference's).
They have no PyTorch forward: all arithmetic runs in the CUDA library through
tsdiff_b200.engine, which reads the live parameter storage in place.

Mirrors (structure only): models/common.py:46-90, models/encoder/edge.py:45-68,
models/encoder/schnet.py:74-171, models/encoder/gin.py:19-113.
"""
import numpy as np
import torch
from torch import nn

NUM_BOND_TYPES = 22  # len(utils.chem.BOND_TYPES), utils/chem.py:21


class _NoForward(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError("%s holds parameters only; compute goes through tsdiff_b200.engine (CUDA)"
                           % type(self).__name__)


class Marker(_NoForward):
    """Parameter-free placeholder that keeps nn.Sequential indices aligned with the
    reference (activation modules sit at index 1)."""

    def __init__(self, name):
        super().__init__()
        self.kind = name

    def extra_repr(self):
        return self.kind


def activation_name(name):
    """utils/__init__.py:6-11 activation_loader: the names the CUDA kernels implement."""
    table = {"swish": "swish", "ReLU": "relu", "relu": "relu", "Softplus": "softplus"}
    if name not in table:
        raise NotImplementedError("activation %r has no CUDA implementation" % (name,))
    return table[name]


class MultiLayerPerceptron(_NoForward):
    def __init__(self, input_dim, hidden_dims, activation="relu"):
        super().__init__()
        self.dims = [input_dim] + list(hidden_dims)
        self.act = activation_name(activation)
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(self.dims[:-1], self.dims[1:])])


class MLPEdgeEncoder(_NoForward):
    def __init__(self, hidden_dim, activation):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.bond_emb = nn.Embedding(100, hidden_dim)
        self.mlp = MultiLayerPerceptron(1, [hidden_dim, hidden_dim], activation=activation)

    @property
    def out_channels(self):
        return self.hidden_dim


def get_edge_encoder(cfg):
    if cfg.edge_encoder == "mlp":
        return MLPEdgeEncoder(cfg.hidden_dim, cfg.mlp_act)
    # the reference's 'gaussian' encoder raises NameError at construction (edge.py:24)
    raise NotImplementedError("Unknown edge encoder: %s" % cfg.edge_encoder)


class CFConv(_NoForward):
    def __init__(self, in_channels, out_channels, num_filters, filter_net, cutoff, smooth):
        super().__init__()
        self.lin1 = nn.Linear(in_channels, num_filters, bias=False)
        self.lin2 = nn.Linear(num_filters, out_channels)
        self.nn = filter_net
        self.cutoff = cutoff
        self.smooth = smooth
        nn.init.xavier_uniform_(self.lin1.weight)
        nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)


class InteractionBlock(_NoForward):
    def __init__(self, hidden_channels, num_gaussians, num_filters, cutoff, smooth):
        super().__init__()
        filter_net = nn.Sequential(nn.Linear(num_gaussians, num_filters), Marker("shifted_softplus"),
                                   nn.Linear(num_filters, num_filters))
        self.conv = CFConv(hidden_channels, hidden_channels, num_filters, filter_net, cutoff, smooth)
        self.act = Marker("shifted_softplus")
        self.lin = nn.Linear(hidden_channels, hidden_channels)


class SchNetEncoder(_NoForward):
    def __init__(self, hidden_channels=128, num_filters=128, num_interactions=6, edge_channels=100, cutoff=10.0,
                 smooth=False, embedding=False):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.num_filters = num_filters
        self.num_interactions = num_interactions
        self.cutoff = cutoff
        self.smooth = smooth
        self.embedding = embedding
        if embedding:
            self.node_emb = nn.Embedding(100, hidden_channels, max_norm=10.0)
        self.interactions = nn.ModuleList([
            InteractionBlock(hidden_channels, edge_channels, num_filters, cutoff, smooth)
            for _ in range(num_interactions)])

    @classmethod
    def from_config(cls, cfg):
        if cfg.edge_emb:
            raise NotImplementedError("encoder.edge_emb=True is broken in the reference (schnet.py:176)")
        return cls(hidden_channels=cfg.hidden_dim, num_filters=cfg.hidden_dim, num_interactions=cfg.num_convs,
                   edge_channels=cfg.hidden_dim, cutoff=cfg.cutoff, smooth=cfg.smooth_conv, embedding=False)


class GINEConv(_NoForward):
    def __init__(self, mlp, eps=0.0):
        super().__init__()
        self.nn = mlp
        self.register_buffer("eps", torch.Tensor([eps]))


class GINEncoder(_NoForward):
    def __init__(self, hidden_dim, num_convs=3, embedding=False):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_convs = num_convs
        self.embedding = embedding
        if embedding:
            self.node_emb = nn.Embedding(100, hidden_dim)
        self.convs = nn.ModuleList([
            GINEConv(MultiLayerPerceptron(hidden_dim, [hidden_dim, hidden_dim], activation="ReLU"))
            for _ in range(num_convs)])


def load_encoder(config, encoder_type="global_encoder"):
    """models/encoder/__init__.py:19-22 restricted to the encoders on the hot path."""
    cfg = config.get(encoder_type)
    if cfg.name == "schnet":
        return SchNetEncoder.from_config(cfg)
    raise NotImplementedError("encoder %r is outside the LD hot path (SURVEY.md section 2)" % cfg.name)


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    """Variance schedule, float64 like models/sampler.py:11-41."""
    n = num_diffusion_timesteps
    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(n, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        x = np.linspace(-6, 6, n)
        betas = 1 / (np.exp(-x) + 1) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (n,)
    return betas


def schedule_parameters(config):
    """betas/alphas exactly as condensenc.py:91-101: float64 -> float32, fp32 cumprod; both
    are frozen nn.Parameters and part of the state_dict."""
    betas = torch.from_numpy(get_beta_schedule(
        beta_schedule=config.beta_schedule, beta_start=config.beta_start, beta_end=config.beta_end,
        num_diffusion_timesteps=config.num_diffusion_timesteps)).float()
    alphas = (1.0 - betas).cumprod(dim=0)
    return nn.Parameter(betas, requires_grad=False), nn.Parameter(alphas, requires_grad=False)
