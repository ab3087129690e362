"""GPU: the training step (BASELINE config 4).

  * every backward kernel against torch.autograd of the same operator in fp32 (bound 1e-5 relative, fp32 re-association);
  * `get_loss(...).mean().backward()` of CondenseEncoderEpsNetwork against the reference's OWN autograd gradients
    (tests/golden/golden_grads.json: per parameter sum / abs-sum / L2 norm in fp64 for all 80 parameters + three full
    gradients) -- bound 1e-4 relative on the norms, 1e-4 of the tensor's RMS elementwise on the full ones;
  * loss values with gradients enabled equal the no_grad (fused sampling kernels) values;
  * after an Adam step the fused (no_grad) and the unfused (training) forward agree on the NEW weights (the engine
    cache follows the parameters).
"""
import ctypes as C
import json
import os

import pytest
import torch

from tsdiff_b200 import _lib as L
from tsdiff_b200 import training as T

from conftest import GOLDEN
from helpers import make_model, rel_err, to_dev

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, tol=1e-5):
    return rel_err(a, b) < tol


@pytest.mark.parametrize("m,k,n,bias", [(1000, 256, 256, True), (777, 512, 256, True), (300, 256, 128, True),
                                        (500, 128, 1, True), (500, 1, 256, True), (64, 25, 128, False), (2000, 256, 256, False)])
def test_linear_forward_backward_vs_torch(m, k, n, bias):
    torch.manual_seed(m + k)
    x = torch.randn(m, k, device=DEV, requires_grad=True)
    w = (torch.randn(n, k, device=DEV) / k ** 0.5).requires_grad_()
    b = torch.randn(n, device=DEV).requires_grad_() if bias else None
    gy = torch.randn(m, n, device=DEV)
    y = T.Linear.apply(x, w, b)
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone() if bias else None)
    x.grad = w.grad = None
    if bias:
        b.grad = None
    y2 = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    y2.backward(gy.double())
    assert _close(got[0], y2) and _close(got[1], x.grad) and _close(got[2], w.grad)
    if bias:
        assert _close(got[3], b.grad)


@pytest.mark.parametrize("act", ["swish", "ssp", "relu"])
def test_activation_backward_vs_torch(act):
    torch.manual_seed(1)
    x = (torch.randn(5000, device=DEV) * 3).requires_grad_()
    gy = torch.randn(5000, device=DEV)
    y = T.Act.apply(x, T._ACT[act])
    y.backward(gy)
    gx = x.grad.clone()
    x.grad = None
    xd = x.double()
    ref = {"swish": lambda t: t * torch.sigmoid(t), "ssp": lambda t: torch.nn.functional.softplus(t) - 0.6931471805599453,
           "relu": torch.relu}[act](xd)
    ref.backward(gy.double())
    assert _close(y, ref, 1e-6) and _close(gx, x.grad, 1e-6)


def test_gate_rows_backward_vs_torch():
    torch.manual_seed(2)
    e, h = 3000, 256
    a = torch.randn(e, h, device=DEV, requires_grad=True)
    table = torch.randn(100, h, device=DEV, requires_grad=True)
    lo, hi = torch.randint(0, 26, (e,), device=DEV), torch.randint(0, 26, (e,), device=DEV)
    code = (lo | (hi << 16)).to(torch.int32)
    gy = torch.randn(e, h, device=DEV)
    for shift, idx in ((0, lo), (16, hi)):
        a.grad = table.grad = None
        y = T.GateRows.apply(a, table, code, shift)
        y.backward(gy)
        ga, gt = a.grad.clone(), table.grad.clone()
        a.grad = table.grad = None
        ref = a.double() * table.double()[idx]
        ref.backward(gy.double())
        assert _close(y, ref, 1e-6) and _close(ga, a.grad, 1e-6) and _close(gt, table.grad)


def _plan(g, pos, cutoff=10.0):
    from tsdiff_b200 import engine as E
    d = to_dev(g, DEV)
    plan = E.BatchPlan(0, d["batch"], d["bond_index"], d["bond_type"], 4, 3, upairs=False)
    plan.build_edges(pos.to(DEV).contiguous(), cutoff)
    return plan, plan.edge_count()


def test_aggregate_pair_features_eq_transform_backward_vs_torch(syn4):
    pos = syn4["pos_init"] * 3.0
    plan, e = _plan(syn4, pos)
    n, h = plan.num_nodes, 256
    row, col = plan.row[:e].long(), plan.col[:e].long()
    torch.manual_seed(3)
    x1 = torch.randn(n, h, device=DEV, requires_grad=True)
    filt = torch.randn(e, h, device=DEV, requires_grad=True)
    gy = torch.randn(n, h, device=DEV)
    agg = T.Aggregate.apply(x1, filt, plan, e)
    agg.backward(gy)
    g1, g2 = x1.grad.clone(), filt.grad.clone()
    x1.grad = filt.grad = None
    ref = torch.zeros(n, h, dtype=torch.float64, device=DEV).index_add_(0, col, x1.double()[row] * filt.double())
    ref.backward(gy.double())
    assert _close(agg, ref, 1e-6) and _close(g1, x1.grad) and _close(g2, filt.grad)
    # pair features
    hh = torch.randn(n, h, device=DEV, requires_grad=True)
    ea = torch.randn(e, h, device=DEV, requires_grad=True)
    gp = torch.randn(e, 2 * h, device=DEV)
    pf = T.PairFeatures.apply(hh, ea, plan, e)
    pf.backward(gp)
    g1, g2 = hh.grad.clone(), ea.grad.clone()
    hh.grad = ea.grad = None
    ref = torch.cat([hh.double()[row] * hh.double()[col], ea.double()], dim=1)
    ref.backward(gp.double())
    assert _close(pf, ref, 1e-6) and _close(g1, hh.grad) and _close(g2, ea.grad)
    # eq_transform on the second graph's edges
    inv = torch.randn(e, device=DEV, requires_grad=True)
    p = pos.to(DEV).contiguous()
    gn = torch.randn(n, 3, device=DEV)
    node = T.EqTransform.apply(inv, p, plan, e, plan.in_b, 1)
    node.backward(gn)
    gi = inv.grad.clone()
    inv.grad = None
    sel = plan.in_b[:e].bool()
    u = (p.double()[row] - p.double()[col]) / plan.length[:e].double().unsqueeze(-1)
    contrib = u * (inv.double() * sel).unsqueeze(-1)
    ref = torch.zeros(n, 3, dtype=torch.float64, device=DEV).index_add_(0, row, contrib).index_add_(0, col, -contrib)
    ref.backward(gn.double())
    assert _close(node, ref, 1e-5) and _close(gi, inv.grad)


def test_node_embed_and_squared_error_backward_vs_torch(syn4):
    d = to_dev(syn4, DEV)
    n = d["atom_type"].numel()
    torch.manual_seed(4)
    emb = torch.randn(100, 128, device=DEV, requires_grad=True)
    wf = torch.randn(128, 25, device=DEV, requires_grad=True)
    gz = torch.randn(n, 256, device=DEV)
    z = T.NodeEmbed.apply(emb, wf, d["atom_type"], d["r_feat"], d["p_feat"])
    z.backward(gz)
    g1, g2 = emb.grad.clone(), wf.grad.clone()
    emb.grad = wf.grad = None
    r, p = d["r_feat"].double(), d["p_feat"].double()
    ref = torch.cat([emb.double()[d["atom_type"]] + r @ wf.double().t(), p @ wf.double().t() - r @ wf.double().t()], dim=1)
    ref.backward(gz.double())
    assert _close(z, ref, 1e-6) and _close(g1, emb.grad) and _close(g2, wf.grad)
    a = torch.randn(n, 3, device=DEV, requires_grad=True)
    b = torch.randn(n, 3, device=DEV)
    gl = torch.randn(n, device=DEV)
    loss = T.SquaredError.apply(a, b)
    loss.backward(gl)
    ga = a.grad.clone()
    a.grad = None
    ref = ((a.double() - b.double()) ** 2).sum(-1)
    ref.backward(gl.double())
    assert _close(loss, ref, 1e-6) and _close(ga, a.grad, 1e-6)


def test_get_loss_gradients_vs_reference_autograd(golden_loss, syn4):
    """train.py:128-143 on the golden batch: loss.mean() and the gradient of all 80 parameters against the reference's
    own autograd (tests/golden/make_golden_grads.py)."""
    gold = json.load(open(os.path.join(GOLDEN, "golden_grads.json")))
    ref = golden_loss["b_syn4"]
    m = make_model("condensenc", 0, DEV)
    m.train()
    d = to_dev(syn4, DEV)
    loss = m.get_loss(d["atom_type"], d["r_feat"], d["p_feat"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"], d["batch"],
                      d["num_nodes_per_graph"], syn4["num_graphs"], time_step=ref["time_step"].to(DEV),
                      pos_noise=ref["pos_noise"].to(DEV))
    assert loss.shape == ref["loss"].shape and loss.requires_grad
    assert rel_err(loss, ref["loss"]) < 1e-4
    mean = loss.mean()
    assert abs(float(mean) - gold["loss_mean"]) < 1e-4 * abs(gold["loss_mean"])
    mean.backward()
    named = dict(m.named_parameters())
    assert len(gold["params"]) == 80
    worst = 0.0
    for name, g in gold["params"].items():
        grad = named[name].grad
        assert grad is not None, name
        gd = grad.double()
        err = abs(float(gd.norm()) - g["norm"]) / max(g["norm"], 1e-12)
        worst = max(worst, err)
        assert err < 1e-4, (name, float(gd.norm()), g["norm"])
        assert abs(float(gd.sum()) - g["sum"]) < 1e-4 * max(g["abs_sum"], 1e-12), name
        assert abs(float(gd.abs().sum()) - g["abs_sum"]) < 1e-4 * max(g["abs_sum"], 1e-12), name
    for name, full in gold["full"].items():
        want = torch.tensor(full, dtype=torch.float64)
        got = named[name].grad.flatten().double().cpu()
        assert float((got - want).abs().max()) < 1e-4 * float(want.pow(2).mean().sqrt()), name
    # parameters that get no gradient in the reference (betas / alphas are frozen) get none here
    for name, p in named.items():
        if name not in gold["params"]:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, name
    print("worst relative gradient-norm error over 80 parameters: %.2e" % worst)


def test_loss_with_grad_equals_no_grad_value_and_follows_an_optimizer_step(golden_loss, syn4):
    ref = golden_loss["b_syn4"]
    m = make_model("condensenc", 0, DEV)
    d = to_dev(syn4, DEV)
    args = (d["atom_type"], d["r_feat"], d["p_feat"], ref["pos"].to(DEV), d["bond_index"], d["bond_type"], d["batch"],
            d["num_nodes_per_graph"], syn4["num_graphs"])
    kw = dict(time_step=ref["time_step"].to(DEV), pos_noise=ref["pos_noise"].to(DEV))
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    with torch.no_grad():
        before = m.get_loss(*args, **kw)
    loss = m.get_loss(*args, **kw)
    assert rel_err(loss, before) < 1e-5
    opt.zero_grad()
    loss.mean().backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 3000.0)
    opt.step()
    with torch.no_grad():
        after = m.get_loss(*args, **kw)           # fused kernels, fresh engine (the cache key follows the weights)
    again = m.get_loss(*args, **kw).detach()      # unfused training forward
    assert rel_err(after, again) < 1e-5
    assert rel_err(after, before) > 1e-3, "the step changed the weights: both paths must see the new ones"


def test_dualenc_training_is_not_built(syn4):
    m = make_model("dualenc", 0, DEV)
    d = to_dev(syn4, DEV)
    with pytest.raises(NotImplementedError):
        m.get_loss(d["atom_type"], syn4["pos_init"].to(DEV), d["bond_index"], d["bond_type"], d["batch"])
