"""TEST INFRASTRUCTURE ONLY -- never imported by the product (tsdiff_b200/).

Pure-PyTorch restatement of the *third-party* graph ops the reference's hot path
calls.  None of them is vendored under /root/reference; the reference pins them in
env.yaml:191-196 (pytorch-geometric 1.7.2, pytorch-scatter 2.0.8, pytorch-sparse
0.6.12, pytorch-cluster 1.5.9, pytorch 1.8.1).  The semantics below are restated
from the published behaviour of those versions ("parity unpinned" at this boundary:
the reference has no tests that hold known answers for them).

Reference call sites these functions stand in for:
  to_dense_adj / dense_to_sparse : models/common.py:158-190,299-310, condensenc.py:137-149
  coalesce                       : models/common.py:197-200,313-315
  radius_graph                   : models/common.py:344
  scatter_add / scatter_mean     : models/geometry.py:25-29, models/sampler.py:261
  MessagePassing.propagate       : models/encoder/schnet.py:102, models/encoder/gin.py:61
"""
import inspect

import torch

MAX_NUM_NEIGHBORS = 32  # PyG radius_graph default; common.py:344 does not override it


# ----------------------------------------------------------------------------- scatter
def scatter_add(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_add along dim 0 (the only dim the path uses)."""
    assert dim == 0
    if out is None:
        if dim_size is None:
            dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
        out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    return out


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_mean: sum / max(count, 1)."""
    total = scatter_add(src, index, dim=dim, dim_size=dim_size)
    count = scatter_add(torch.ones_like(index, dtype=src.dtype), index, dim=0, dim_size=total.shape[0])
    count = count.clamp(min=1)
    shape = [-1] + [1] * (total.dim() - 1)
    if total.is_floating_point():
        return total / count.view(shape)
    return torch.div(total, count.view(shape), rounding_mode="floor")


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    res = torch.full((dim_size,) + tuple(src.shape[1:]), torch.iinfo(src.dtype).min
                     if not src.is_floating_point() else float("-inf"), dtype=src.dtype)
    res = res.scatter_reduce(0, index.view([-1] + [1] * (src.dim() - 1)).expand_as(src), src, "amax")
    return res, None


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_add(src, index, dim=dim, out=out, dim_size=dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim=dim, dim_size=dim_size)
    raise NotImplementedError(reduce)


# ------------------------------------------------------------------------ dense <-> sparse
def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    """PyG 1.7.2 utils.to_dense_adj for batch=None: (1, n, n[, ...]); n = max_num_nodes
    or max index + 1; duplicate entries are SUMMED; entries are float32 ones when no
    edge_attr is given, else keep edge_attr's dtype."""
    assert batch is None, "the path never passes batch"
    n_seen = int(edge_index.max().item()) + 1
    n = n_seen if max_num_nodes is None else int(max_num_nodes)
    row, col = edge_index[0], edge_index[1]
    if max_num_nodes is not None and n_seen > n:
        keep = (row < n) & (col < n)
        row, col = row[keep], col[keep]
        edge_attr = None if edge_attr is None else edge_attr[keep]
    if edge_attr is None:
        edge_attr = torch.ones(row.numel(), device=edge_index.device)
    size = [1, n, n] + list(edge_attr.shape[1:])
    flat = torch.zeros([n * n] + list(edge_attr.shape[1:]), dtype=edge_attr.dtype, device=edge_index.device)
    flat.index_add_(0, row * n + col, edge_attr)
    return flat.view(size)


def dense_to_sparse(adj):
    """PyG 1.7.2 utils.dense_to_sparse: nonzero() in row-major order; a 3-D input
    offsets node indices by b*n."""
    assert adj.dim() in (2, 3) and adj.size(-1) == adj.size(-2)
    index = adj.nonzero(as_tuple=True)
    values = adj[index]
    if len(index) == 3:
        off = index[0] * adj.size(-1)
        index = (off + index[1], off + index[2])
    return torch.stack(index, dim=0), values


def coalesce(index, value, m, n, op="add"):
    """torch_sparse.coalesce: sort by row*n+col, duplicates summed."""
    key = index[0] * n + index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    out_index = torch.stack([uniq // n, uniq % n], dim=0)
    if value is None:
        return out_index, None
    out_val = torch.zeros((uniq.numel(),) + tuple(value.shape[1:]), dtype=value.dtype, device=value.device)
    out_val.index_add_(0, inv, value)
    return out_index, out_val


# --------------------------------------------------------------------------- radius graph
def pair_dist2(pos_a, pos_b):
    """Canonical fp32 squared distance used for the neighbour test: every operation
    individually rounded to fp32, summed as (dx*dx + dy*dy) + dz*dz.  The CUDA kernel
    mirrors this with __fmul_rn/__fadd_rn so edge sets are bit-exact."""
    d = pos_a - pos_b
    sq = d * d
    return (sq[..., 0] + sq[..., 1]) + sq[..., 2]


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=MAX_NUM_NEIGHBORS,
                 flow="source_to_target", num_workers=1):
    """PyG radius_graph -> torch_cluster.radius(x, x, r, batch, batch, max+1), CUDA rule:
    for every centre y scan the same-graph points x in index order, keep those with
    dist2 < r*r (strict) until max_num_neighbors(+1 without loops, self included) hits,
    emit (row = neighbour, col = centre), then drop row == col.  Output is grouped by
    centre in index order, neighbours ascending."""
    assert flow == "source_to_target"
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    # torch_cluster squares the (double) radius on the host and casts to scalar_t
    r2 = torch.tensor(float(r) * float(r), dtype=x.dtype)
    rows, cols = [], []
    # per graph dense tile (graphs are small); block-diagonal by construction
    counts = torch.bincount(batch, minlength=int(batch.max().item()) + 1 if n else 0)
    start = 0
    for cnt in counts.tolist():
        if cnt == 0:
            continue
        p = x[start:start + cnt]
        d2 = pair_dist2(p[None, :, :], p[:, None, :])  # [centre, neighbour]
        inr = d2 < r2
        rank = torch.cumsum(inr.to(torch.long), dim=1) - 1  # rank of neighbour among in-range, index order
        keep = inr & (rank < cap)
        if not loop:
            keep = keep & ~torch.eye(cnt, dtype=torch.bool, device=x.device)
        c_idx, n_idx = keep.nonzero(as_tuple=True)
        rows.append(n_idx + start)
        cols.append(c_idx + start)
        start += cnt
    if not rows:
        return torch.zeros((2, 0), dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=MAX_NUM_NEIGHBORS, num_workers=1):
    raise NotImplementedError("sidechain path (common.py:345-368) is outside the hot path")


# ------------------------------------------------------------------------ message passing
class MessagePassing(torch.nn.Module):
    """Minimal PyG MessagePassing(aggr='add', flow='source_to_target'): message() args
    named `<k>_j` are gathered from kwargs[k] at edge_index[0]; `<k>_i` at edge_index[1];
    everything else is passed through; messages are summed into edge_index[1]."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        names = [p for p in inspect.signature(self.message).parameters]
        args, num_nodes = {}, None
        for name in names:
            if name.endswith("_j") or name.endswith("_i"):
                data = kwargs[name[:-2]]
                if isinstance(data, (tuple, list)):
                    data = data[0] if name.endswith("_j") else data[1]
                num_nodes = data.size(0)
                args[name] = data[edge_index[0] if name.endswith("_j") else edge_index[1]]
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        return scatter_add(msg, edge_index[1], dim=0, dim_size=num_nodes)

    def message(self, x_j):
        return x_j
