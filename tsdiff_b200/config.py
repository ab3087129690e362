"""Attribute-dict config (stand-in for easydict.EasyDict, which the reference uses for the
YAML config: train.py:46-47, sampling.py:128).  Supports attribute access, `.get`,
`hasattr`, and nested dicts, which is everything the eps-nets ask of it."""
import yaml


class AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        for k, v in state.items():
            self[k] = v


def load_config(path):
    with open(path) as f:
        return AttrDict(yaml.safe_load(f))


# configs/train_config.yml `model:` block of the reference (the shipped TS network)
TRAIN_CONFIG_MODEL = AttrDict({
    "type": "diffusion", "network": "condensenc", "t0": 0, "t1": 5000,
    "edge_cutoff": 10.0, "edge_order": 4, "pred_edge_order": 3,
    "encoder": {"name": "schnet", "edge_emb": False, "num_convs": 7, "cutoff": 10.0, "smooth_conv": False,
                "mlp_act": "swish", "hidden_dim": 256},
    "feat_dim": 25, "hidden_dim": 256, "edge_encoder": "mlp", "mlp_act": "swish", "edge_cat_act": "swish",
    "beta_schedule": "sigmoid", "beta_start": 1.0e-7, "beta_end": 2.0e-3, "num_diffusion_timesteps": 5000,
})

# configs/geodiff_legacy/qm9_default.yml `model:` block (the DualEncoderEpsNetwork config)
QM9_DEFAULT_MODEL = AttrDict({
    "type": "diffusion", "network": "dualenc", "hidden_dim": 128, "num_convs": 6, "num_convs_local": 4,
    "cutoff": 10.0, "mlp_act": "ReLU", "beta_schedule": "sigmoid", "beta_start": 1.0e-7, "beta_end": 2.0e-3,
    "num_diffusion_timesteps": 5000, "edge_order": 3, "edge_encoder": "mlp", "smooth_conv": False,
})
