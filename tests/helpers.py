"""Shared test utilities: seeded product models, state_dict -> oracle params, comparisons."""
import torch

from tsdiff_b200.config import QM9_DEFAULT_MODEL, TRAIN_CONFIG_MODEL
from tsdiff_b200.models.epsnet import get_model


def make_model(kind, seed=0, device="cpu"):
    cfg = TRAIN_CONFIG_MODEL if kind == "condensenc" else QM9_DEFAULT_MODEL
    torch.manual_seed(seed)
    m = get_model(cfg)
    return m.to(device)


def oracle_params(model):
    """CPU copies of the state_dict for the functional oracle (detached clones: the oracle's
    max_norm renorm mutates its own copy)."""
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def max_rel_err(a, b):
    """Worst element error relative to the RMS magnitude of the reference vector.  (A plain
    element-wise |a-b|/|b| is ill-conditioned for the few edge scores that pass through 0.)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt().clamp(min=1e-30))


def to_dev(g, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in g.items()}
