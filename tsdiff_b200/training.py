"""Training step of CondenseEncoderEpsNetwork on the GPU -- SURVEY.md section 8(f)-2, BASELINE config 4.

`condensed_loss(model, ...)` is `get_loss` (models/epsnet/condensenc.py:267-328) with gradients: the operators of the
eps-net run UNFUSED through libtsdiff_b200.so (pre-activations kept), each wrapped in a torch.autograd.Function whose
backward is again a kernel of the library (csrc/train_ops.cu) -- torch is the tape, not the arithmetic; without the
library (or with CPU tensors) everything here raises.  fp32 throughout; reductions are deterministic.

`allreduce_gradients` is the data-parallel glue of train.py:140-145 for one process per GPU: the loss is a PER-NODE
mean (`loss.mean()` over the atoms of the batch), so a rank's gradients are weighted by its share of the atoms before
the NCCL sum -- the result equals the single-GPU gradient of the concatenated batch.
"""
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib as L
from . import engine as E

_ACT = {"none": 0, "relu": 1, "swish": 2, "ssp": 3, "softplus": 4}


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _linear_raw(x, weight, bias):
    """x (M, K) @ weight (N, K)^T + bias -> (M, N), fp32 FFMA (tsd_linear)."""
    out = torch.empty(x.size(0), weight.size(0), dtype=torch.float32, device=x.device)
    lin = L.linear(weight, bias)
    L.check(L.load().tsd_linear(x.size(0), None, L.ptr(x), C.byref(lin), 0, L.ptr(out), 0, _s()), "tsd_linear")
    return out


def _transpose(w):
    out = torch.empty(w.size(1), w.size(0), dtype=torch.float32, device=w.device)
    L.check(L.load().tsd_transpose(w.size(0), w.size(1), L.ptr(w), L.ptr(out), _s()), "tsd_transpose")
    return out


def _wgrad(dy, x, want_bias):
    m, n, k = dy.size(0), dy.size(1), x.size(1)
    lib = L.load()
    need = C.c_uint64()
    L.check(lib.tsd_linear_wgrad_scratch(m, n, k, C.byref(need)), "tsd_linear_wgrad_scratch")
    scratch = torch.empty(max(need.value, 1), dtype=torch.float32, device=dy.device)
    dw = torch.empty(n, k, dtype=torch.float32, device=dy.device)
    db = torch.empty(n, dtype=torch.float32, device=dy.device) if want_bias else None
    L.check(lib.tsd_linear_wgrad(m, n, k, L.ptr(dy), L.ptr(x), L.ptr(dw), L.ptr(db), L.ptr(scratch), _s()), "tsd_linear_wgrad")
    return dw, db


class Linear(Function):
    """nn.Linear: y = x W^T + b.  backward: dx = dy W (tsd_linear on W^T), dW = dy^T x, db = sum_m dy."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = _c(x), _c(weight)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return _linear_raw(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        dx = _linear_raw(dy, _transpose(weight), None) if ctx.needs_input_grad[0] else None
        dw, db = _wgrad(dy, x, ctx.has_bias)
        return dx, dw, db


class Act(Function):
    @staticmethod
    def forward(ctx, x, act):
        x = _c(x)
        ctx.save_for_backward(x)
        ctx.act = act
        y = torch.empty_like(x)
        L.check(L.load().tsd_act_forward(x.numel(), L.ptr(x), act, L.ptr(y), _s()), "tsd_act_forward")
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        L.check(L.load().tsd_act_backward(x.numel(), L.ptr(x), L.ptr(dy), ctx.act, L.ptr(dx), _s()), "tsd_act_backward")
        return dx, None


def _row_scale(x, s):
    out = torch.empty_like(x)
    L.check(L.load().tsd_row_scale(x.size(0), x.size(1), L.ptr(x), L.ptr(s), L.ptr(out), _s()), "tsd_row_scale")
    return out


class RowScale(Function):
    """x * C(len)[:, None]; the envelope has no gradient (positions are not differentiated, train.py:128-143)."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(s)
        return _row_scale(_c(x), s)

    @staticmethod
    def backward(ctx, dy):
        (s,) = ctx.saved_tensors
        return _row_scale(_c(dy), s), None


def _gate(a, table, code, shift):
    rows, h = code.numel(), table.size(1)
    out = torch.empty(rows, h, dtype=torch.float32, device=table.device)
    L.check(L.load().tsd_gate_rows(rows, h, L.ptr(a), L.ptr(table), L.ptr(code), shift, L.ptr(out), _s()), "tsd_gate_rows")
    return out


class GateRows(Function):
    """edge.py:66-68: a * table[code] (code = the low or high 16 bits of the packed pair-table value)."""

    @staticmethod
    def forward(ctx, a, table, code, shift):
        a, table = _c(a), _c(table)
        ctx.save_for_backward(a, table, code)
        ctx.shift = shift
        return _gate(a, table, code, shift)

    @staticmethod
    def backward(ctx, dy):
        a, table, code = ctx.saved_tensors
        dy = _c(dy)
        lib = L.load()
        da = _gate(dy, table, code, ctx.shift)
        # d table = onehot(code)^T (dy * a): a weight-gradient GEMM, deterministic
        g = torch.empty_like(dy)
        L.check(lib.tsd_mul(dy.numel(), L.ptr(dy), L.ptr(a), L.ptr(g), _s()), "tsd_mul")
        onehot = torch.empty(code.numel(), table.size(0), dtype=torch.float32, device=dy.device)
        L.check(lib.tsd_onehot(code.numel(), table.size(0), L.ptr(code), ctx.shift, L.ptr(onehot), _s()), "tsd_onehot")
        dtable, _ = _wgrad(onehot, g, False)
        return da, dtable, None, None


class Aggregate(Function):
    """schnet.py:102-107: agg_i = sum_{j->i} x1_j * filt_ji over the plan's directed edges."""

    @staticmethod
    def forward(ctx, x1, filt, plan, num_edges):
        x1, filt = _c(x1), _c(filt)
        ctx.save_for_backward(x1, filt)
        ctx.plan, ctx.num_edges = plan, num_edges
        agg = torch.empty_like(x1)
        L.check(L.load().tsd_cfconv_aggregate(C.byref(plan.c_batch), C.byref(plan.c_edges), x1.size(1), L.ptr(x1),
                                              L.ptr(filt), L.ptr(agg), _s()), "tsd_cfconv_aggregate")
        return agg

    @staticmethod
    def backward(ctx, dagg):
        x1, filt = ctx.saved_tensors
        plan = ctx.plan
        dagg = _c(dagg)
        dx1, dfilt = torch.empty_like(x1), torch.empty_like(filt)
        L.check(L.load().tsd_cfconv_aggregate_backward(C.byref(plan.c_batch), C.byref(plan.c_edges), ctx.num_edges, x1.size(1),
                                                       L.ptr(x1), L.ptr(filt), L.ptr(dagg), L.ptr(dx1), L.ptr(dfilt), _s()),
                "tsd_cfconv_aggregate_backward")
        return dx1, dfilt, None, None


class PairFeatures(Function):
    """common.py:226-229: cat[h_row * h_col, edge_attr]."""

    @staticmethod
    def forward(ctx, h, ea, plan, num_edges):
        h, ea = _c(h), _c(ea)
        ctx.save_for_backward(h)
        ctx.plan, ctx.num_edges = plan, num_edges
        out = torch.empty(num_edges, 2 * h.size(1), dtype=torch.float32, device=h.device)
        L.check(L.load().tsd_pair_features(C.byref(plan.c_edges), num_edges, h.size(1), L.ptr(h), L.ptr(ea), L.ptr(out), _s()),
                "tsd_pair_features")
        return out

    @staticmethod
    def backward(ctx, dout):
        (h,) = ctx.saved_tensors
        plan = ctx.plan
        dout = _c(dout)
        dh = torch.empty_like(h)
        L.check(L.load().tsd_pair_features_backward(C.byref(plan.c_batch), C.byref(plan.c_edges), h.size(1), L.ptr(h),
                                                    L.ptr(dout), L.ptr(dh), _s()), "tsd_pair_features_backward")
        return dh, dout[:, h.size(1):].contiguous(), None, None


def _eq_transform(plan, pos, inv, mask, mask_mode):
    out = torch.empty(max(plan.num_nodes, 1), 3, dtype=torch.float32, device=pos.device)
    ch = L.ScoreChannel(inv.data_ptr(), mask.data_ptr() if mask is not None else None, mask_mode, 0.0, 1.0, None)
    L.check(L.load().tsd_eq_transform(C.byref(plan.c_batch), C.byref(plan.c_edges), L.ptr(pos), C.byref(ch), 1.0, L.ptr(out),
                                      _s()), "tsd_eq_transform")
    return out[:plan.num_nodes]


class EqTransform(Function):
    """geometry.py:22-30 on the plan's directed edges selected by `mask`; linear in the edge scores."""

    @staticmethod
    def forward(ctx, inv, pos, plan, num_edges, mask, mask_mode):
        inv = _c(inv)
        ctx.save_for_backward(pos)
        ctx.plan, ctx.num_edges, ctx.mask, ctx.mask_mode = plan, num_edges, mask, mask_mode
        return _eq_transform(plan, pos, inv, mask, mask_mode)

    @staticmethod
    def backward(ctx, dnode):
        (pos,) = ctx.saved_tensors
        plan = ctx.plan
        dnode = _c(dnode)
        dinv = torch.empty(ctx.num_edges, dtype=torch.float32, device=pos.device)
        L.check(L.load().tsd_eq_transform_backward(C.byref(plan.c_edges), ctx.num_edges, L.ptr(pos), L.ptr(ctx.mask),
                                                   ctx.mask_mode, 1.0, L.ptr(dnode), L.ptr(dinv), _s()),
                "tsd_eq_transform_backward")
        return dinv, None, None, None, None, None


class NodeEmbed(Function):
    """condensenc.py:193-198: z = cat[emb[Z] + W r, W p - W r]."""

    @staticmethod
    def forward(ctx, emb_weight, feat_weight, atom_type, r_feat, p_feat):
        ctx.save_for_backward(atom_type, r_feat, p_feat)
        ctx.shapes = (emb_weight.shape, feat_weight.shape)
        n, half = atom_type.numel(), emb_weight.size(1)
        z = torch.empty(max(n, 1), 2 * half, dtype=torch.float32, device=emb_weight.device)
        L.check(L.load().tsd_condensed_node_embed(n, L.ptr(atom_type), L.ptr(r_feat), L.ptr(p_feat), feat_weight.size(1),
                                                  L.ptr(_c(emb_weight)), L.ptr(_c(feat_weight)), half, L.ptr(z), _s()),
                "tsd_condensed_node_embed")
        return z[:n]

    @staticmethod
    def backward(ctx, dz):
        atom_type, r_feat, p_feat = ctx.saved_tensors
        (types, half), (_, fdim) = ctx.shapes
        dz = _c(dz)
        demb = torch.empty(types, half, dtype=torch.float32, device=dz.device)
        dw = torch.empty(half, fdim, dtype=torch.float32, device=dz.device)
        L.check(L.load().tsd_condensed_node_embed_backward(atom_type.numel(), L.ptr(atom_type), L.ptr(r_feat), L.ptr(p_feat),
                                                           fdim, half, types, L.ptr(dz), L.ptr(demb), L.ptr(dw), _s()),
                "tsd_condensed_node_embed_backward")
        return demb, dw, None, None, None


class Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        out = torch.empty_like(a)
        L.check(L.load().tsd_add(a.numel(), L.ptr(a), L.ptr(b), L.ptr(out), _s()), "tsd_add")
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class SquaredError(Function):
    """condensenc.py:324-326: sum_d (a - b)^2 per atom; b (the target) has no gradient."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        ctx.save_for_backward(a, b)
        loss = torch.empty(a.size(0), dtype=torch.float32, device=a.device)
        L.check(L.load().tsd_sqerr_forward(a.size(0), L.ptr(a), L.ptr(b), L.ptr(loss), _s()), "tsd_sqerr_forward")
        return loss

    @staticmethod
    def backward(ctx, dloss):
        a, b = ctx.saved_tensors
        dloss = _c(dloss)
        da = torch.empty_like(a)
        L.check(L.load().tsd_sqerr_backward(a.size(0), L.ptr(a), L.ptr(b), L.ptr(dloss), L.ptr(da), _s()), "tsd_sqerr_backward")
        return da, None


def _lin(x, layer):
    return Linear.apply(x, layer.weight, layer.bias)


def condensed_loss(model, atom_type, r_feat, p_feat, pos, bond_index, bond_type, batch, time_step, pos_noise):
    """Per-atom loss (N, 1) of condensenc.py:267-328 WITH the autograd graph over the model's parameters.
    time_step (G,) and pos_noise (N, 3) are the draws of :287-295 (the caller makes them)."""
    cfg = model.config
    dev = pos.device
    for t, nm in ((pos, "pos"), (batch, "batch")):
        E._require_cuda(t, nm)
    lib = L.load()
    with torch.cuda.device(dev):
        a = model.alphas.to(dev).index_select(0, time_step.to(dev))
        a_pos = a.index_select(0, batch).unsqueeze(-1)
        pos = pos.to(torch.float32)
        pos_p = (pos + pos_noise.to(dev) * (1.0 - a_pos).sqrt() / a_pos.sqrt()).to(torch.float32).contiguous()
        # graph (positions are not differentiated): directed rows, one per edge -- no pair sharing in training
        plan = E.BatchPlan(0, batch, bond_index, bond_type, int(cfg.edge_order), int(cfg.pred_edge_order), upairs=False)
        plan.build_edges(pos_p, float(cfg.edge_cutoff))
        ne = plan.edge_count()
        two_graphs = int(cfg.edge_order) != int(cfg.pred_edge_order)
        length = plan.length[:ne].reshape(ne, 1).contiguous()
        code_a, code_b = plan.tab0[:ne].contiguous(), plan.tab1[:ne].contiguous()
        atom_type = atom_type.to(torch.long).contiguous()
        r_feat, p_feat = E._integer_features(r_feat, "r_feat"), E._integer_features(p_feat, "p_feat")

        z = NodeEmbed.apply(model.atom_embedding.weight, model.atom_feat_embedding.weight, atom_type, r_feat, p_feat)
        enc = model.edge_encoder
        act = _ACT[enc.mlp.act]
        t = Act.apply(_lin(length, enc.mlp.layers[0]), act)
        d_emb = _lin(t, enc.mlp.layers[1])
        cat_act = _ACT[E.L_act(cfg.edge_cat_act)]

        def edge_attr(code):  # condensenc.py:156-176
            gr = GateRows.apply(d_emb, enc.bond_emb.weight, code, 0)
            gp = GateRows.apply(d_emb, enc.bond_emb.weight, code, 16)
            c = _lin(torch.cat([gr, gp], dim=1), model.edge_cat[0])
            return _lin(Act.apply(c, cat_act), model.edge_cat[2])

        ea = edge_attr(code_a)
        h = z
        ssp = _ACT["ssp"]
        for blk in model.encoder.interactions:  # schnet.py:90-128
            env = torch.empty(ne, dtype=torch.float32, device=dev)
            L.check(lib.tsd_cutoff_envelope(ne, L.ptr(length), float(blk.conv.cutoff), int(bool(blk.conv.smooth)), L.ptr(env),
                                            _s()), "tsd_cutoff_envelope")
            w = RowScale.apply(_lin(Act.apply(_lin(ea, blk.conv.nn[0]), ssp), blk.conv.nn[2]), env)
            x1 = Linear.apply(h, blk.conv.lin1.weight, None)
            agg = Aggregate.apply(x1, w, plan, ne)
            y = _lin(Act.apply(_lin(agg, blk.conv.lin2), ssp), blk.lin)
            h = Add.apply(h, y)
        ea_out = edge_attr(code_b) if two_graphs else ea
        mlp = model.grad_dist_mlp
        mact = _ACT[mlp.act]
        o = PairFeatures.apply(h, ea_out, plan, ne)
        o = Act.apply(_lin(o, mlp.layers[0]), mact)
        o = Act.apply(_lin(o, mlp.layers[1]), mact)
        edge_inv = _lin(o, mlp.layers[2]).reshape(ne)
        mask = plan.in_b if two_graphs else None
        mode = 1 if two_graphs else 0
        node_eq = EqTransform.apply(edge_inv, pos_p, plan, ne, mask, mode)
        with torch.no_grad():  # the target (condensenc.py:309-322)
            row, col = plan.row[:ne].long(), plan.col[:ne].long()
            a_edge = a.index_select(0, batch.index_select(0, row))
            d_gt = (pos[row] - pos[col]).norm(dim=-1)
            d_target = ((d_gt - length.reshape(ne)) / (1.0 - a_edge).sqrt() * a_edge.sqrt()).contiguous()
            pos_target = _eq_transform(plan, pos_p, d_target, mask, mode)
        return SquaredError.apply(node_eq, pos_target).unsqueeze(-1)


def allreduce_gradients(parameters, num_local_nodes, group=None):
    """Data-parallel gradient reduction for `loss.mean().backward()` (train.py:140-143) with the batch sharded over the
    ranks of `group`: every rank holds d(mean over ITS atoms)/d theta, the single-GPU gradient is the atom-weighted
    mean, so grads are scaled by N_local / N_total and summed.  One flat bucket (2.77 M fp32 = 11 MB for the shipped
    model: a single NCCL all-reduce over NVLink; the backward is milliseconds, so there is nothing to overlap)."""
    import torch.distributed as dist
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return 0
    dev = params[0].grad.device
    count = torch.tensor([float(num_local_nodes)], dtype=torch.float64, device=dev)
    dist.all_reduce(count, group=group)
    weight = float(num_local_nodes) / float(count.item())
    flat = torch.cat([p.grad.reshape(-1) for p in params]) * weight
    dist.all_reduce(flat, group=group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off
