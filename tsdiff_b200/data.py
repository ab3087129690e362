"""Caller-side stand-ins for what `sampling.py` takes from torch_geometric:
`Data`, `Batch.from_data_list` / `to_data_list`, the `batching` / `repeat` helpers
(sampling.py:24-42) and the body of its sampling loop (sampling.py:169-225) as `sample_batch`, so
the reference's script logic runs without PyG / rdkit installed.  Pure host code: tensors stay
torch tensors, the device work happens inside `model.dynamic_sampling`.

Collation follows torch_geometric 1.7.2 (`Batch.from_data_list`, restated from its published
behaviour -- the wheel is not importable here): tensor attributes are concatenated along dim 0,
`edge_index`-like attributes (name contains "index") along the last dim after adding the running
node offset, python scalars become tensors, everything else (smiles strings, rdkit mols) becomes a
per-graph list; `batch` holds the graph id of every node.
"""
import copy
import pickle

import torch

_NODE_KEYS = ("atom_type", "pos", "x", "r_feat", "p_feat")  # attributes whose dim 0 is the node count


class Data:
    """Attribute bag with the slice of the PyG `Data` interface the sampling script touches."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    # ---- PyG-like protocol
    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not (k.startswith("__") and k.endswith("__"))]

    def __getitem__(self, key):
        return getattr(self, key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, self[k]

    @property
    def num_nodes(self):
        if "__num_nodes__" in self.__dict__ and self.__dict__["__num_nodes__"] is not None:
            return self.__dict__["__num_nodes__"]
        for k in _NODE_KEYS:
            v = self.__dict__.get(k)
            if torch.is_tensor(v):
                return v.size(0)
        ei = self.__dict__.get("edge_index")
        return int(ei.max()) + 1 if torch.is_tensor(ei) and ei.numel() else None

    @num_nodes.setter
    def num_nodes(self, n):
        self.__dict__["__num_nodes__"] = n

    def __cat_dim__(self, key, value):
        return -1 if "index" in key or "face" in key else 0

    def __inc__(self, key, value):
        return self.num_nodes if "index" in key or "face" in key else 0

    def apply(self, fn):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = fn(v)
        return self

    def to(self, device, *args, **kwargs):
        return self.apply(lambda t: t.to(device, *args, **kwargs))

    def cpu(self):
        return self.to("cpu")

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.copy(v)
        return out

    def __repr__(self):
        items = ["%s=%s" % (k, list(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in self]
        return "%s(%s)" % (self.__class__.__name__, ", ".join(items))


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        keys = []
        for d in data_list:
            for k in d.keys:
                if k not in keys:
                    keys.append(k)
        assert "batch" not in keys
        batch = cls()
        slices = {k: [0] for k in keys}
        cols = {k: [] for k in keys}
        batch_vec, offset = [], 0
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if torch.is_tensor(v):
                    inc = d.__inc__(k, v)
                    if inc:
                        v = v + offset
                    size = v.size(d.__cat_dim__(k, v)) if v.dim() > 0 else 1
                    if v.dim() == 0:
                        v = v.unsqueeze(0)
                elif isinstance(v, (int, float)) and not isinstance(v, bool):
                    v, size = torch.tensor([v]), 1
                else:
                    size = 1
                cols[k].append(v)
                slices[k].append(slices[k][-1] + size)
            batch_vec.append(torch.full((n,), g, dtype=torch.long))
            offset += n
        for k in keys:
            first = cols[k][0]
            if torch.is_tensor(first):
                batch.__dict__[k] = torch.cat(cols[k], dim=data_list[0].__cat_dim__(k, first))
            else:
                batch.__dict__[k] = cols[k]
        batch.__dict__["batch"] = torch.cat(batch_vec) if batch_vec else torch.zeros(0, dtype=torch.long)
        batch.__dict__["__slices__"] = slices
        batch.__dict__["__num_graphs__"] = len(data_list)
        batch.__dict__["__node_counts__"] = [d.num_nodes for d in data_list]
        batch.__dict__["__num_nodes__"] = offset
        batch.__dict__["__tensor_keys__"] = [k for k in keys if torch.is_tensor(cols[k][0])]
        batch.__dict__["__scalar_keys__"] = [k for k in keys if any(
            isinstance(d[k], (int, float)) and not isinstance(d[k], bool) for d in data_list)]
        return batch

    @property
    def num_graphs(self):
        return self.__dict__["__num_graphs__"]

    def to_data_list(self):
        slices = self.__dict__["__slices__"]
        out, offset = [], 0
        for g in range(self.num_graphs):
            d = Data()
            for k, sl in slices.items():
                v = self.__dict__[k]
                if torch.is_tensor(v):
                    dim = self.__cat_dim__(k, v)
                    item = v.narrow(dim, sl[g], sl[g + 1] - sl[g])
                    if "index" in k or "face" in k:
                        item = item - offset
                    if k in self.__dict__["__scalar_keys__"]:
                        item = item[0].item()
                    d.__dict__[k] = item
                else:
                    d.__dict__[k] = v[g]
            d.__dict__["__num_nodes__"] = self.__dict__["__node_counts__"][g]
            offset += self.__dict__["__node_counts__"][g]
            out.append(d)
        return out


def count_nodes_per_graph(data):
    """utils/transforms.py:188-196 (CountNodesPerGraph)."""
    if data.__dict__.get("__num_nodes__") is None:
        data.num_nodes = len(data.pos)
    data.num_nodes_per_graph = torch.LongTensor([data.num_nodes])
    return data


def repeat(iterable, num):
    """sampling.py:24-28."""
    out = []
    for x in iterable:
        out.extend([x.clone() for _ in range(num)])
    return out


def batching(iterable, batch_size, repeat_num=1):
    """sampling.py:33-42."""
    items = repeat(iterable, repeat_num)
    for cnt in range(0, len(items), batch_size):
        yield items[cnt: cnt + batch_size]


def sample_batch(model, batch, n_steps=5000, step_lr=1e-7, clip=1000.0, sampling_type="ld", eta=1.0,
                 noise_from_time_t=None, denoise_from_time_t=None, from_ts_guess=False, save_traj=False,
                 pos_init=None, **extras):
    """The body of sampling.py's loop for one collated batch (sampling.py:169-225): start geometry
    (random normal, or the TS guess divided by sqrt(alpha) when `from_ts_guess`), one
    `model.dynamic_sampling` call, the alpha-scaled trajectory, and the per-reaction `Data` objects
    with `pos_gen` set, on the CPU -- the list sampling.py appends to `results` and pickles.
    `model`: tsdiff_b200.models.sampler.EnsembleSampler.  Raises FloatingPointError like the reference
    (the script's retry loop stays with the caller)."""
    device = batch.atom_type.device
    if pos_init is None:
        if from_ts_guess:
            assert denoise_from_time_t is not None
            init_guess = batch.ts_guess if "ts_guess" in batch else batch.pos
            start_t = noise_from_time_t if noise_from_time_t is not None else denoise_from_time_t
            sqrt_a = model.alphas[start_t - 1].sqrt() if start_t != 0 else 1
            pos_init = (init_guess / sqrt_a).to(device)
        else:
            pos_init = torch.randn(batch.num_nodes, 3).to(device)
    pos_gen, pos_gen_traj = model.dynamic_sampling(
        atom_type=batch.atom_type, r_feat=batch.r_feat, p_feat=batch.p_feat, pos_init=pos_init,
        bond_index=batch.edge_index, bond_type=batch.edge_type, batch=batch.batch, num_graphs=batch.num_graphs,
        extend_order=True, n_steps=n_steps, step_lr=step_lr, clip=clip, sampling_type=sampling_type, eta=eta,
        noise_from_time_t=noise_from_time_t, denoise_from_time_t=denoise_from_time_t, keep_traj=save_traj, **extras)
    if save_traj:
        alphas = model.alphas.detach()
        if denoise_from_time_t is not None:
            alphas = alphas[denoise_from_time_t - n_steps: denoise_from_time_t]
        else:
            alphas = alphas[model.num_timesteps - n_steps: model.num_timesteps]
        alphas = alphas.flip(0).view(-1, 1, 1)
        traj = torch.stack(pos_gen_traj) * alphas.sqrt().cpu()
    results = []
    batch_vec_cpu = batch.batch.cpu()
    pos_gen_cpu = pos_gen.cpu()
    for j, data in enumerate(batch.to_data_list()):
        mask = batch_vec_cpu == j
        data.pos_gen = traj[:, mask] if save_traj else pos_gen_cpu[mask]
        results.append(data.to("cpu"))
    return results


class _InertHolder:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["state"] = state


class _SampleUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("torch_geometric") or module.startswith("rdkit"):
            return type(name, (_InertHolder,), {})
        return super().find_class(module, name)


def load_samples(path):
    """Reads a result / test-set pickle written by the reference (a list of PyG `Data`, e.g.
    birkholz_benchmark/rxn_0/samples_all.pkl) or by `save_samples` into a list of `Data`, without
    torch_geometric or rdkit (their classes are replaced by inert holders; rdkit mols are dropped)."""
    with open(path, "rb") as f:
        objs = _SampleUnpickler(f).load()
    out = []
    for o in objs:
        if isinstance(o, Data):
            out.append(o)
            continue
        d = Data()
        for k, v in o.__dict__.get("state", {}).items():
            if v is None or k == "rdmol":
                continue
            d.__dict__[k] = v
        out.append(d)
    return out


def save_samples(results, path):
    """sampling.py:219-221: the list of per-reaction results, pickled."""
    with open(path, "wb") as f:
        pickle.dump(results, f)
