"""Host-side driver of the CUDA library: owns device buffers (through torch), binds the
live nn.Parameter storage into the C structs, and issues the kernel sequence of one
eps-net evaluation / one Langevin step on torch's current stream (so the whole step can be
captured in a CUDA graph and replayed with no host round-trip).

torch is plumbing here (allocation, streams, graphs); every arithmetic op is a kernel of
libtsdiff_b200.so.  Nothing in this module falls back to PyTorch math.
"""
import ctypes as C

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_plan_device(fn):
    """Runs a method with the engine's device current: the library launches on torch's current stream of
    the CURRENT device and keeps per-process side streams, so a process that also drives another GPU must
    not call in with that one current.  (One process per GPU is the supported layout)"""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        with torch.cuda.device(self.plan.device):
            return fn(self, *args, **kwargs)
    return wrapped


def require_cuda_inputs(**tensors):
    for name, t in tensors.items():
        _require_cuda(t, name)


def resolve_seed(seed):
    """Philox seed of one sampling call.  The reference draws its noise with torch.randn_like
    (sampler.py:213, dualenc.py:858), which advances torch's global generator: two calls -- two batches,
    two `--repeat` copies, the retry after a FloatingPointError (sampling.py:171-236) -- never see the
    same stream.  Without an explicit seed the key is therefore drawn from the global generator per
    call: reproducible under torch.manual_seed, different for every call."""
    if seed is None:
        return int(torch.randint(0, 2 ** 62, (), dtype=torch.int64).item())
    return int(seed)


def _integer_features(t, name):
    """r_feat / p_feat are one-hot (preprocessing.py:151-164); the embedding kernel reads them as int64."""
    if t.is_floating_point():
        if not bool((t == t.round()).all()):
            raise L.TsdError("%s holds non-integer values: the CUDA node embedding takes one-hot / integer "
                             "features (int64), it would truncate them" % name)
    return t.to(torch.long).contiguous()


def _require_cuda(t, name):
    if not t.is_cuda:
        raise L.TsdError("%s must be a CUDA tensor: tsdiff_b200 has no CPU path" % name)


class BatchPlan:
    """Everything about a batch of reaction graphs that does not depend on positions:
    graph offsets, the bond-order pair tables (K1) and the capacity-sized edge arrays that
    K2 refills every step.  mode 0 = TS / condensenc tables, mode 1 = dualenc tables."""

    def __init__(self, mode, batch, bond_index, bond_type, order_a, order_b=0, ts_decode=False, upairs=False):
        lib = L.load()
        _require_cuda(batch, "batch")
        dev = batch.device
        self.device = dev
        self.mode = mode
        batch = batch.contiguous()
        n = batch.numel()
        counts = torch.bincount(batch) if n else torch.zeros(0, dtype=torch.long, device=dev)
        counts_h = counts.cpu()
        if n and not bool((batch[1:] >= batch[:-1]).all()):
            raise ValueError("`batch` must be sorted (PyG Batch contract)")
        if n and int(counts_h.min()) == 0:
            raise ValueError("empty graphs are not supported")
        g = counts_h.numel()
        self.num_nodes, self.num_graphs = n, g
        self.max_graph_nodes = int(counts_h.max()) if g else 1
        if self.max_graph_nodes > 256:
            raise L.TsdError("graphs with more than 256 atoms are not supported (got %d)" % self.max_graph_nodes)
        self.edge_capacity = int((counts_h * (counts_h - 1)).sum()) if g else 0
        self.num_pairs = int((counts_h * counts_h).sum()) if g else 0
        i32 = dict(dtype=torch.int32, device=dev)
        gp = torch.zeros(g + 1, dtype=torch.long)
        pp = torch.zeros(g + 1, dtype=torch.long)
        if g:
            gp[1:] = counts_h.cumsum(0)
            pp[1:] = (counts_h * counts_h).cumsum(0)
        self.graph_ptr = gp.to(**i32)
        self.pair_ptr = pp.to(**i32)
        self.node_graph = batch.to(torch.int32)
        self.counts = counts_h
        self.c_batch = L.Batch(n, g, self.max_graph_nodes, self.edge_capacity, self.graph_ptr.data_ptr(),
                               self.pair_ptr.data_ptr(), self.node_graph.data_ptr())

        # ---- K1: bond-order pair tables
        self.table_a = torch.zeros(max(self.num_pairs, 1), **i32)
        self.table_b = torch.zeros(max(self.num_pairs, 1), **i32)
        scratch = torch.empty(max(2 * self.num_pairs, 1), **i32)
        err = torch.zeros(1, **i32)
        bond_index = bond_index.to(device=dev, dtype=torch.long).contiguous()
        bond_type = bond_type.to(device=dev, dtype=torch.long).contiguous()
        L.check(lib.tsd_bond_order_build(mode, C.byref(self.c_batch), bond_index.size(1), L.ptr(bond_index),
                                         L.ptr(bond_type), order_a, max(order_b, 1), int(ts_decode),
                                         L.ptr(self.table_a), L.ptr(self.table_b), L.ptr(scratch), L.ptr(err),
                                         _stream()), "tsd_bond_order_build")
        code = int(err.item())
        if code:
            raise ValueError("bond_index/bond_type invalid for this batch (flag %d: 1 = index or type out of "
                             "range, 2 = bond across two graphs)" % code)

        # ---- K2 outputs, capacity sized
        cap = max(self.edge_capacity, 1)
        self.num_edges = torch.zeros(2, **i32)
        self.row = torch.zeros(cap, **i32)
        self.col = torch.zeros(cap, **i32)
        self.length = torch.ones(cap, dtype=torch.float32, device=dev)
        self.tab0 = torch.zeros(cap, **i32)
        self.tab1 = torch.zeros(cap, **i32)
        self.in_b = torch.zeros(cap, **i32)
        self.row_ptr = torch.zeros(n + 1, **i32)
        self.in_ptr = torch.zeros(n + 1, **i32)
        self.in_eid = torch.zeros(cap, **i32)
        self.in_src = torch.zeros(cap, **i32)
        self.graph_count = torch.zeros(max(g, 1), **i32)
        # ---- undirected pair list: the per-edge networks run once per unordered pair (the two directions
        # of an edge have bit-identical length and, for symmetric pair tables, types)
        self.upair_capacity = int((counts_h * (counts_h - 1)).sum()) // 2 if g else 0
        self.upairs = bool(upairs) and self._tables_symmetric()
        fields = [self.num_edges, self.row, self.col, self.length, self.tab0, self.tab1, self.in_b, self.row_ptr,
                  self.in_ptr, self.in_eid, self.in_src, self.graph_count]
        if self.upairs:
            ucap = max(self.upair_capacity, 1)
            self.num_upairs = torch.zeros(1, **i32)
            self.u_row = torch.zeros(ucap, **i32)
            self.u_col = torch.zeros(ucap, **i32)
            self.u_length = torch.ones(ucap, dtype=torch.float32, device=dev)
            self.u_tab0 = torch.zeros(ucap, **i32)
            self.u_tab1 = torch.zeros(ucap, **i32)
            self.edge_upair = torch.zeros(cap, **i32)
            self.in_upair = torch.zeros(cap, **i32)
            self.graph_ucount = torch.zeros(max(g, 1), **i32)
            fields += [self.num_upairs, self.u_row, self.u_col, self.u_length, self.u_tab0, self.u_tab1,
                       self.edge_upair, self.in_upair, self.graph_ucount]
            self.c_edges = L.Edges(*[t.data_ptr() for t in fields])
            # what the per-edge kernels see: one row per unordered pair; the in-CSR still walks the DIRECTED
            # in-edges of every node but addresses the pair's row (in_upair in the in_eid slot)
            self.c_work_edges = L.Edges(*[t.data_ptr() for t in (
                self.num_upairs, self.u_row, self.u_col, self.u_length, self.u_tab0, self.u_tab1, self.in_b,
                self.row_ptr, self.in_ptr, self.in_upair, self.in_src, self.graph_count)])
            self.c_work_batch = L.Batch(n, g, self.max_graph_nodes, self.upair_capacity, self.graph_ptr.data_ptr(),
                                        self.pair_ptr.data_ptr(), self.node_graph.data_ptr())
            self.work_capacity, self.work_tab0, self.work_tab1 = self.upair_capacity, self.u_tab0, self.u_tab1
        else:
            self.c_edges = L.Edges(*[t.data_ptr() for t in fields])
            self.c_work_edges, self.c_work_batch = self.c_edges, self.c_batch
            self.work_capacity, self.work_tab0, self.work_tab1 = self.edge_capacity, self.tab0, self.tab1

    def _tables_symmetric(self):
        """True when both pair tables equal their per-reaction transposes (a symmetric bond list, what
        datasets.py:495-498 produces).  Position independent: checked once per batch."""
        c = self.counts
        if c.numel() == 0:
            return True
        n_of = torch.repeat_interleave(c, c * c)                         # n_g of every table entry
        base = torch.repeat_interleave(self.pair_ptr[:-1].cpu().long(), c * c)
        local = torch.arange(int((c * c).sum())) - base
        i, j = local // n_of, local % n_of
        perm = (base + j * n_of + i).to(self.device)
        return bool(torch.equal(self.table_a, self.table_a[perm]) and torch.equal(self.table_b, self.table_b[perm]))

    def work_count(self):
        """Rows the per-edge kernels process (unordered pairs, or directed edges).  Synchronises."""
        return int(self.num_upairs[0].item()) if self.upairs else self.edge_count()

    def build_edges(self, pos, cutoff, max_neighbors=32):
        """K2: refill the edge list for the current positions (no sync)."""
        lib = L.load()
        assert pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape == (self.num_nodes, 3)
        L.check(lib.tsd_edge_build(C.byref(self.c_batch), L.ptr(pos), float(cutoff), int(max_neighbors),
                                   L.ptr(self.table_a), L.ptr(self.table_b), 1 if self.mode == 0 else 0,
                                   C.byref(self.c_edges), _stream()), "tsd_edge_build")

    def edge_count(self):
        return int(self.num_edges[0].item())  # synchronises; API-level use only


class _Scratch:
    """Capacity-sized activation buffers shared by all ensemble members."""

    def __init__(self, plan, hidden, n_edge_bufs, n_node_bufs, network=0, math="fp32"):
        dev = plan.device
        # the library states what its kernel sequence needs (tsd_workspace_bytes); keep the two in step
        ne, nn = C.c_int32(), C.c_int32()
        L.check(L.load().tsd_workspace_bytes(plan.num_nodes, plan.work_capacity, hidden, network, L.MATH[math], None, None,
                                             C.byref(ne), C.byref(nn)), "tsd_workspace_bytes")
        assert ne.value == n_edge_bufs and nn.value == n_node_bufs + (2 if math == "tf32" else 0), (ne.value, nn.value)
        cap = max(plan.work_capacity, 1)
        self.edge = [torch.empty(cap, hidden, dtype=torch.float32, device=dev) for _ in range(n_edge_bufs)]
        self.node = [torch.empty(max(plan.num_nodes, 1), hidden, dtype=torch.float32, device=dev)
                     for _ in range(n_node_bufs)]


class _WeightView:
    """Maps an nn.Linear weight to the tensor the GEMM reads.  fp32 mode: the live parameter
    itself.  tf32 mode: a round-to-nearest TF32 shadow copy (the tensor cores read fp32 operands
    by truncation; weights are constant while sampling, so they are rounded once per engine)."""

    def __init__(self, math):
        self.tf32 = math == "tf32"
        self.keep = []

    def __call__(self, w):
        if not self.tf32:
            return w
        shadow = torch.empty_like(w)
        L.check(L.load().tsd_round_tf32(L.ptr(w), L.ptr(shadow), w.numel(), _stream()), "tsd_round_tf32")
        self.keep.append(shadow)
        return shadow


def _node_pool(plan, hidden, math):
    """tf32 mode: the two extra (N, H) buffers of the encoder's split node update (tsd_schnet_encoder)."""
    if math != "tf32":
        return None, 0
    return torch.empty(2 * max(plan.num_nodes, 1) * hidden, dtype=torch.float32, device=plan.device), 2


def _filter_pool(plan, hidden, num_blocks, math):
    """tf32 mode: one filter buffer per interaction block beyond the two of the workspace, so the filter networks of
    all blocks (they only depend on edge_attr) run ahead of the serial node-side chain (tsd_schnet_encoder)."""
    extra = max(num_blocks - 2, 0)
    if math != "tf32" or extra == 0:
        return None, 0
    return torch.empty(extra * max(plan.work_capacity, 1) * hidden, dtype=torch.float32, device=plan.device), extra


def _edge_encoder_struct(enc, act, cat=None, cat_act="none", wv=lambda w: w):
    """enc: layers.MLPEdgeEncoder; cat: nn.Sequential(Linear, act, Linear) or None."""
    keep = []
    s = L.EdgeEncoder()
    s.lin0 = L.linear(enc.mlp.layers[0].weight, enc.mlp.layers[0].bias)
    s.lin1 = L.linear(wv(enc.mlp.layers[1].weight), enc.mlp.layers[1].bias)
    s.bond_emb = enc.bond_emb.weight.data_ptr()
    s.act = L.ACT[act]
    if cat is not None:
        c0 = L.linear(wv(cat[0].weight), cat[0].bias)
        c2 = L.linear(wv(cat[2].weight), cat[2].bias)
        keep += [c0, c2]
        s.cat0 = C.pointer(c0)
        s.cat2 = C.pointer(c2)
        s.cat_act = L.ACT[cat_act]
    return s, keep


def _interaction_struct(blk, wv=lambda w: w, next_blk=None):
    s = L.Interaction()
    s.nn0 = L.linear(wv(blk.conv.nn[0].weight), blk.conv.nn[0].bias)
    s.nn2 = L.linear(wv(blk.conv.nn[2].weight), blk.conv.nn[2].bias)
    s.lin1 = L.linear(wv(blk.conv.lin1.weight), None)
    s.lin2 = L.linear(wv(blk.conv.lin2.weight), blk.conv.lin2.bias)
    s.lin = L.linear(wv(blk.lin.weight), blk.lin.bias)
    s.cutoff = float(blk.conv.cutoff)
    s.smooth = int(bool(blk.conv.smooth))
    if next_blk is not None and getattr(wv, "tf32", False):
        # the next block's lin1 folded into this block's lin (schnet.py:101 after :128):
        # lin1_next(h + lin(y)) = lin1_next(h) + (W1 W_lin) y + W1 b_lin
        with torch.no_grad():
            w1 = next_blk.conv.lin1.weight.double()
            fw = (w1 @ blk.lin.weight.double()).float().contiguous()
            fb = (w1 @ blk.lin.bias.double()).float().contiguous()
        wv.keep.append(fb)
        s.fused_w = wv(fw).data_ptr()
        s.fused_b = fb.data_ptr()
    return s


def _interaction_structs(blocks, wv):
    blocks = list(blocks)
    return [_interaction_struct(b, wv, blocks[i + 1] if i + 1 < len(blocks) else None) for i, b in enumerate(blocks)]


def _interaction_array(blocks):
    """Contiguous C array of tsd_interaction_t for tsd_schnet_encoder."""
    return (L.Interaction * len(blocks))(*blocks)


def _pair_mlp_struct(mlp, wv=lambda w: w):
    s = L.PairMlp()
    s.l0 = L.linear(wv(mlp.layers[0].weight), mlp.layers[0].bias)
    s.l1 = L.linear(wv(mlp.layers[1].weight), mlp.layers[1].bias)
    s.l2 = L.linear(mlp.layers[2].weight, mlp.layers[2].bias)
    s.act = L.ACT[mlp.act]
    return s


def _gine_struct(conv, relu_after, wv=lambda w: w):
    s = L.Gine()
    s.nn0 = L.linear(wv(conv.nn.layers[0].weight), conv.nn.layers[0].bias)
    s.nn1 = L.linear(wv(conv.nn.layers[1].weight), conv.nn.layers[1].bias)
    s.eps = conv.eps.data_ptr()
    s.relu_after = int(relu_after)
    return s


class CondensedScoreEngine:
    """Path B: eps-net evaluation of an ensemble of CondenseEncoderEpsNetwork members on one
    BatchPlan (models/epsnet/condensenc.py:178-239 per member, models/sampler.py:58-116 for
    the ensemble sum).  `edge_inv` holds the SUM over members; consumers divide by M."""

    def __init__(self, models, atom_type, r_feat, p_feat, bond_index, bond_type, batch, math="fp32"):
        lib = L.load()
        cfg = models[0].config
        self.models = list(models)
        self.cfg = cfg
        self.math = L.MATH[math]
        for t, nm in ((atom_type, "atom_type"), (r_feat, "r_feat"), (p_feat, "p_feat"), (batch, "batch")):
            _require_cuda(t, nm)
        for m in self.models:
            if next(m.parameters()).device != batch.device:
                raise L.TsdError("model parameters and inputs must live on the same CUDA device")
        if int(cfg.pred_edge_order) > int(cfg.edge_order):
            # graph b is kept as a mask over graph a's edge list (graph_build.cu): that needs b's local pairs to
            # be a subset of a's.  The reference rebuilds the graph (condensenc.py:219-237) and has no such limit.
            raise NotImplementedError("pred_edge_order (%d) > edge_order (%d): the second graph must be a subset "
                                      "of the first" % (int(cfg.pred_edge_order), int(cfg.edge_order)))
        self.plan = BatchPlan(0, batch, bond_index, bond_type, int(cfg.edge_order), int(cfg.pred_edge_order), upairs=True)
        self.two_graphs = int(cfg.edge_order) != int(cfg.pred_edge_order)
        plan = self.plan
        h = int(cfg.hidden_dim)
        self.hidden = h
        self.cutoff = float(cfg.edge_cutoff)
        # d_emb, tmp, ea1, ea2, ef0, ef1, tmp2
        self.ws = _Scratch(plan, h, 7, 4, 0, math)
        self.side = torch.cuda.Stream(device=plan.device)  # second-graph edge embedding runs beside the encoder
        self.nf_pool, self.nf_pool_count = _node_pool(plan, h, math)
        self.ef_pool, self.ef_pool_count = _filter_pool(plan, h, len(models[0].encoder.interactions), math)
        self.edge_inv = torch.zeros(max(plan.work_capacity, 1), dtype=torch.float32, device=plan.device)
        # second graph: rows whose type codes differ from the first graph's (tsd_edge_embed_delta)
        cap = max(plan.work_capacity, 1)
        self.diff_rows = torch.zeros(cap, dtype=torch.int32, device=plan.device)
        self.diff_pos = torch.full((cap,), -1, dtype=torch.int32, device=plan.device)
        self.diff_count = torch.zeros(1, dtype=torch.int32, device=plan.device)
        atom_type = atom_type.to(torch.long).contiguous()
        r_feat = _integer_features(r_feat, "r_feat")
        p_feat = _integer_features(p_feat, "p_feat")
        self.members = []
        self.wv = _WeightView(math)
        for m in self.models:
            z = torch.empty(max(plan.num_nodes, 1), h, dtype=torch.float32, device=plan.device)
            L.check(lib.tsd_condensed_node_embed(plan.num_nodes, L.ptr(atom_type), L.ptr(r_feat), L.ptr(p_feat),
                                                 r_feat.size(1) if r_feat.dim() == 2 else int(cfg.feat_dim),
                                                 L.ptr(m.atom_embedding.weight), L.ptr(m.atom_feat_embedding.weight),
                                                 h // 2, L.ptr(z), _stream()), "tsd_condensed_node_embed")
            enc, keep = _edge_encoder_struct(m.edge_encoder, m.edge_encoder.mlp.act, m.edge_cat,
                                             L_act(cfg.edge_cat_act), self.wv)
            blocks = _interaction_array(_interaction_structs(m.encoder.interactions, self.wv))
            pair = _pair_mlp_struct(m.grad_dist_mlp, self.wv)
            # x1 of the first interaction block = lin1_0(z): position independent like z; the encoder fills it on the
            # first evaluation and skips the kernel afterwards (tsd_schnet_encoder's x1_first)
            x1_first = torch.empty(max(plan.num_nodes, 1), h, dtype=torch.float32, device=plan.device)
            self.members.append({"z": z, "enc": enc, "keep": keep, "blocks": blocks, "pair": pair, "x1_first": x1_first,
                                 "x1_valid": 0})

    @_on_plan_device
    def evaluate(self, pos):
        """One ensemble eps-net evaluation at `pos`; fills plan edges and self.edge_inv (sum
        over members, on the edges of graph a; consumers select graph b with plan.in_b)."""
        lib = L.load()
        plan, ws, s = self.plan, self.ws, _stream()
        b, e = C.byref(plan.c_work_batch), C.byref(plan.c_work_edges)  # rows = unordered pairs when plan.upairs
        d_emb, tmp, ea1, ea2, ef0, ef1, tmp2 = ws.edge
        hbuf, nf0, nf1, nf2 = ws.node
        plan.build_edges(pos, self.cutoff)
        main = torch.cuda.current_stream()
        for mi, mem in enumerate(self.members):
            enc = C.byref(mem["enc"])
            L.check(lib.tsd_edge_embed(b, e, L.ptr(plan.work_tab0), enc, 0, L.ptr(d_emb), L.ptr(tmp), L.ptr(ea1),
                                       self.math, s), "tsd_edge_embed")
            if self.two_graphs:
                fork = torch.cuda.Event()
                fork.record(main)
            L.check(lib.tsd_schnet_encoder(b, e, L.ptr(ea1), mem["blocks"], len(mem["blocks"]), L.ptr(mem["z"]),
                                           L.ptr(hbuf), L.ptr(ef0), L.ptr(ef1), L.ptr(nf0), L.ptr(nf1), L.ptr(nf2),
                                           L.ptr(self.nf_pool), self.nf_pool_count, L.ptr(self.ef_pool),
                                           self.ef_pool_count, L.ptr(mem["x1_first"]), mem["x1_valid"], self.math, s),
                    "tsd_schnet_encoder")
            if not torch.cuda.is_current_stream_capturing():
                mem["x1_valid"] = 1  # a call baked into a CUDA graph has to recompute it on every replay
            if self.two_graphs:
                # the pred_edge_order graph's edge embedding only needs d_emb and is only read by the pair
                # MLP: a side stream (a graph branch under capture) lets it fill the SMs the encoder leaves
                # idle.  It is ENQUEUED after the encoder so the first filter kernel is not queued behind it.
                self.side.wait_event(fork)
                with torch.cuda.stream(self.side):
                    L.check(lib.tsd_edge_embed_delta(b, e, L.ptr(plan.work_tab0), L.ptr(plan.work_tab1), enc, L.ptr(d_emb),
                                                     L.ptr(tmp2), L.ptr(ea2), L.ptr(self.diff_rows), L.ptr(self.diff_pos),
                                                     L.ptr(self.diff_count), self.math, _stream()), "tsd_edge_embed_delta")
                main.wait_stream(self.side)
                L.check(lib.tsd_pair_mlp_delta(b, e, L.ptr(hbuf), L.ptr(ea1), L.ptr(ea2), L.ptr(self.diff_pos),
                                               C.byref(mem["pair"]), 1 if mi > 0 else 0, L.ptr(ef0), L.ptr(self.edge_inv),
                                               self.math, s), "tsd_pair_mlp_delta")
            else:
                L.check(lib.tsd_pair_mlp(b, e, L.ptr(hbuf), L.ptr(ea1), C.byref(mem["pair"]), 1 if mi > 0 else 0,
                                         L.ptr(ef0), L.ptr(self.edge_inv), self.math, s), "tsd_pair_mlp")

    def score_channels(self, clip):
        ch0 = L.ScoreChannel(self.edge_inv.data_ptr(), self.plan.in_b.data_ptr(), 1 if self.two_graphs else 0,
                             float(clip) if clip is not None else 0.0, 1.0,
                             self.plan.edge_upair.data_ptr() if self.plan.upairs else None)
        return ch0, None

    def edge_inv_directed(self, e):
        """edge_inv (sum over members) of the first e directed edges."""
        if self.plan.upairs:
            return self.edge_inv[self.plan.edge_upair[:e].long()]
        return self.edge_inv[:e]

    @property
    def num_members(self):
        return len(self.members)


def eq_transform_directed(plan, pos, values):
    """models/geometry.py:22-30 on the plan's CURRENT directed edge list: `values` (e,) holds one scalar per
    directed edge in plan order (0 for edges that do not take part).  Returns (N,3).  Deterministic
    (per-atom sequential sums in edge order, K7's arithmetic)."""
    e = values.numel()
    full = torch.zeros(max(plan.edge_capacity, 1), dtype=torch.float32, device=plan.device)
    full[:e] = values.reshape(-1).to(torch.float32)
    out = torch.empty(max(plan.num_nodes, 1), 3, dtype=torch.float32, device=plan.device)
    ch = L.ScoreChannel(full.data_ptr(), None, 0, 0.0, 1.0, None)
    L.check(L.load().tsd_eq_transform(C.byref(plan.c_batch), C.byref(plan.c_edges), L.ptr(pos.contiguous()),
                                      C.byref(ch), 1.0, L.ptr(out), _stream()), "tsd_eq_transform")
    return out[:plan.num_nodes]


def require_no_grad(module, what):
    """The backward kernels are not built: loss values are available for evaluation only."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise NotImplementedError(
            "%s: forward (validation) value only -- call under torch.no_grad(); training needs the backward "
            "kernels (SURVEY.md section 8(f)-2, not built yet)" % what)


def L_act(name):
    from .models.layers import activation_name
    return activation_name(name)


class DualScoreEngine:
    """Path A: DualEncoderEpsNetwork evaluation (models/epsnet/dualenc.py:206-374, type
    'diffusion').  Global SchNet on all edges, local GIN on edges with type > 0."""

    def __init__(self, model, atom_type, bond_index, bond_type, batch, math="fp32", extend_order=True, extend_radius=True):
        lib = L.load()
        cfg = model.config
        self.model, self.cfg = model, cfg
        self.math = L.MATH[math]
        _require_cuda(atom_type, "atom_type")
        _require_cuda(batch, "batch")
        self.ts = bool(getattr(model, "TS", False))
        # common.py:387-417: extend_order=False keeps the bonds as they are (= hop order 1), extend_radius=False adds no
        # radius edges (a build cutoff of 0: d^2 < 0 never holds); CFConv keeps config.cutoff for its envelope
        self.plan = BatchPlan(1, batch, bond_index, bond_type, int(cfg.edge_order) if extend_order else 1, 0,
                              ts_decode=self.ts, upairs=True)
        plan = self.plan
        h = int(cfg.hidden_dim)
        self.hidden = h
        self.cutoff = float(cfg.cutoff) if extend_radius else 0.0
        self.ws = _Scratch(plan, h, 7, 7, 1, math)
        self.side = torch.cuda.Stream(device=plan.device)  # the local (GIN) branch runs beside the global one
        self.nf_pool, self.nf_pool_count = _node_pool(plan, h, math)
        self.ef_pool, self.ef_pool_count = _filter_pool(plan, h, len(model.encoder_global.interactions), math)
        cap = max(plan.work_capacity, 1)  # rows of the per-edge networks: unordered pairs when plan.upairs
        # TS variant (edge_cat): the local edge encoder needs its own d_emb / tmp scratch
        self.local_scratch = ([torch.empty(cap, h, dtype=torch.float32, device=plan.device) for _ in range(2)]
                              if self.ts else [None, None])
        self.edge_inv_global = torch.zeros(cap, dtype=torch.float32, device=plan.device)
        # lin1 of the first global interaction block applied to the (position independent) node embedding
        self.x1_first = torch.empty(max(plan.num_nodes, 1), h, dtype=torch.float32, device=plan.device)
        self.x1_valid = 0
        self.edge_inv_local = torch.zeros(cap, dtype=torch.float32, device=plan.device)
        self.atom_type = atom_type.to(torch.long).contiguous()
        act = model.edge_encoder_global.mlp.act
        cat_act = L_act(cfg.edge_cat_act) if self.ts else "none"
        self.wv = _WeightView(math)
        self.enc_g, self._k1 = _edge_encoder_struct(model.edge_encoder_global, act,
                                                    model.edge_cat_global if self.ts else None, cat_act, self.wv)
        self.enc_l, self._k2 = _edge_encoder_struct(model.edge_encoder_local, act,
                                                    model.edge_cat_local if self.ts else None, cat_act, self.wv)
        self.blocks = _interaction_array(_interaction_structs(model.encoder_global.interactions, self.wv))
        n_local = len(model.encoder_local.convs)
        self.gines = [_gine_struct(c, i < n_local - 1, self.wv) for i, c in enumerate(model.encoder_local.convs)]
        self.pair_g = _pair_mlp_struct(model.grad_global_dist_mlp, self.wv)
        self.pair_l = _pair_mlp_struct(model.grad_local_dist_mlp, self.wv)
        n = max(plan.num_nodes, 1)
        self.h0_global = torch.empty(n, h, dtype=torch.float32, device=plan.device)
        self.h0_local = torch.empty(n, h, dtype=torch.float32, device=plan.device)
        self.refresh_embeddings()

    @_on_plan_device
    def refresh_embeddings(self):
        """node_emb lookups (weights are constant during sampling).  The global embedding has
        max_norm=10: looked-up rows are renormalised IN PLACE like nn.Embedding does."""
        lib = L.load()
        m, plan = self.model, self.plan
        wg = m.encoder_global.node_emb.weight
        L.check(lib.tsd_embedding(plan.num_nodes, L.ptr(self.atom_type), L.ptr(wg), wg.size(0), wg.size(1), 10.0,
                                  L.ptr(self.h0_global), _stream()), "tsd_embedding")
        wl = m.encoder_local.node_emb.weight
        L.check(lib.tsd_embedding(plan.num_nodes, L.ptr(self.atom_type), L.ptr(wl), wl.size(0), wl.size(1), 0.0,
                                  L.ptr(self.h0_local), _stream()), "tsd_embedding")

    @_on_plan_device
    def evaluate(self, pos):
        lib = L.load()
        plan, ws, s = self.plan, self.ws, _stream()
        b, e = C.byref(plan.c_work_batch), C.byref(plan.c_work_edges)
        d_emb, tmp, ea_g, ea_l, ef0, ef1, efl = ws.edge
        hbuf, nf0, nf1, nf2, hloc, nf0l, nf1l = ws.node
        plan.build_edges(pos, self.cutoff)
        codes = L.ptr(plan.work_tab1)
        main = torch.cuda.current_stream()
        # local branch (independent of the global one) on the side stream: edge encoder on all edges,
        # GIN + pair MLP restricted to type > 0 by masks
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            ss = _stream()
            L.check(lib.tsd_edge_embed(b, e, codes, C.byref(self.enc_l), 0, L.ptr(self.local_scratch[0]),
                                       L.ptr(self.local_scratch[1]), L.ptr(ea_l), self.math, ss), "tsd_edge_embed")
            h_in = self.h0_local
            for gc in self.gines:
                L.check(lib.tsd_gine_layer(b, e, L.ptr(ea_l), C.byref(gc), L.ptr(h_in), L.ptr(hloc), L.ptr(nf0l),
                                           L.ptr(nf1l), self.math, ss), "tsd_gine_layer")
                h_in = hloc
            L.check(lib.tsd_pair_mlp(b, e, L.ptr(h_in), L.ptr(ea_l), C.byref(self.pair_l), 0, L.ptr(efl),
                                     L.ptr(self.edge_inv_local), self.math, ss), "tsd_pair_mlp")
        # global: edge encoder -> SchNet -> pair MLP on every edge
        L.check(lib.tsd_edge_embed(b, e, codes, C.byref(self.enc_g), 0, L.ptr(d_emb), L.ptr(tmp), L.ptr(ea_g),
                                   self.math, s), "tsd_edge_embed")
        L.check(lib.tsd_schnet_encoder(b, e, L.ptr(ea_g), self.blocks, len(self.blocks), L.ptr(self.h0_global),
                                       L.ptr(hbuf), L.ptr(ef0), L.ptr(ef1), L.ptr(nf0), L.ptr(nf1), L.ptr(nf2),
                                       L.ptr(self.nf_pool), self.nf_pool_count, L.ptr(self.ef_pool), self.ef_pool_count,
                                       L.ptr(self.x1_first), self.x1_valid, self.math, s), "tsd_schnet_encoder")
        if not torch.cuda.is_current_stream_capturing():
            self.x1_valid = 1
        L.check(lib.tsd_pair_mlp(b, e, L.ptr(hbuf), L.ptr(ea_g), C.byref(self.pair_g), 0, L.ptr(ef0),
                                 L.ptr(self.edge_inv_global), self.math, s), "tsd_pair_mlp")
        main.wait_stream(self.side)

    def score_channels(self, clip, clip_local, w_global):
        """dualenc.py:827-849: local score on type > 0 edges (+ optional clip_local); global
        score on the remaining edges, clipped, weighted by w_global."""
        index = self.plan.edge_upair.data_ptr() if self.plan.upairs else None
        ch0 = L.ScoreChannel(self.edge_inv_local.data_ptr(), self.plan.tab0.data_ptr(), 1,
                             float(clip_local) if clip_local is not None else 0.0, 1.0, index)
        ch1 = L.ScoreChannel(self.edge_inv_global.data_ptr(), self.plan.tab0.data_ptr(), 2,
                             float(clip) if clip is not None else 0.0, float(w_global), index)
        return ch0, ch1

    def directed(self, per_row, e):
        """A per-row result of the edge networks for the first e directed edges."""
        return per_row[self.plan.edge_upair[:e].long()] if self.plan.upairs else per_row[:e]

    num_members = 1


class PeerExchange:
    """Peer-mapped buffers of the one-shot score exchange fused into K7 (tsd_exchange_t): the ensemble-member-per-GPU
    mode (BASELINE config 3).  Every rank of `group` owns a (2, world, N, 3) score buffer and (2, world, G) flags in
    memory its peers can store into (CUDA IPC over NVLink); the 64-byte handles travel through the process group once.
    `PeerExchange.local(plan, world)` wires `world` emulated ranks of ONE process / ONE device together instead (used
    by the single-GPU protocol test: same kernel code, the "peers" are just other buffers)."""

    def __init__(self, plan, group=None, _local=None):
        import torch.distributed as dist
        lib = L.load()
        self.plan = plan
        n, g = max(plan.num_nodes, 1), max(plan.num_graphs, 1)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=plan.device)
        self._opened, self._owned, self._group = [], None, group
        if _local is not None:
            self.world, self.rank, bufs = _local
            self._keep = bufs
            ptrs = [(d.data_ptr(), f.data_ptr()) for d, f in bufs]
        else:
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            if self.world > L.MAX_EXCHANGE_RANKS:
                raise L.TsdError("the fused score exchange supports up to %d ranks" % L.MAX_EXCHANGE_RANKS)
            data_bytes = 2 * self.world * n * 3 * 4
            data_bytes = (data_bytes + 255) // 256 * 256
            total = data_bytes + 2 * self.world * g * 4
            base, handle = C.c_void_p(), (C.c_ubyte * 64)()
            with torch.cuda.device(plan.device):
                L.check(lib.tsd_peer_alloc(total, C.byref(base), handle), "tsd_peer_alloc")
                self._owned = base.value
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
                ptrs = []
                for p in range(self.world):
                    if p == self.rank:
                        addr = base.value
                    else:
                        q = C.c_void_p()
                        buf = (C.c_ubyte * 64).from_buffer_copy(handles[p])
                        L.check(lib.tsd_peer_open(buf, C.byref(q)), "tsd_peer_open")
                        self._opened.append(q.value)
                        addr = q.value
                    ptrs.append((addr, addr + data_bytes))
                dist.barrier(group=group)  # every rank has opened every buffer before anyone stores into them
        self.c = L.Exchange()
        self.c.world, self.c.rank, self.c.num_graphs = self.world, self.rank, plan.num_graphs
        for p, (d, f) in enumerate(ptrs):
            self.c.peer_data[p] = d
            self.c.peer_flags[p] = f
        self.c.epoch_base = self.epoch.data_ptr()

    @classmethod
    def local(cls, plans):
        """One PeerExchange per emulated rank (plans[r] = that rank's BatchPlan, all on one device)."""
        world = len(plans)
        n, g = max(plans[0].num_nodes, 1), max(plans[0].num_graphs, 1)
        dev = plans[0].device
        bufs = [(torch.zeros(2 * world * n * 3, dtype=torch.float32, device=dev),
                 torch.zeros(2 * world * g, dtype=torch.int32, device=dev)) for _ in range(world)]
        return [cls(plans[r], _local=(world, r, bufs)) for r in range(world)]

    def new_trajectory(self, n_steps):
        """Flag values are epoch_base + step + 1: moving the base past the last trajectory's values makes its
        leftovers (and those of a faster peer that already started the next one) unambiguous without any reset."""
        self.epoch.add_(int(n_steps) + 1)

    def close(self):
        """Collective over the group: every rank unmaps its peers' buffers, and only when ALL ranks have done so does
        anyone free its own (an exporter must not free memory a peer still has mapped)."""
        lib = L.load()
        if self._owned is None and not self._opened:
            return
        torch.cuda.synchronize(self.plan.device)
        for q in self._opened:
            lib.tsd_peer_close(C.c_void_p(q))
        self._opened = []
        if self._owned:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.barrier(group=self._group)
            lib.tsd_peer_free(C.c_void_p(self._owned))
            self._owned = None

_TRAJ_STAGING = {}  # device index -> ([two flat pinned float32 buffers], copy stream, weakref of the last user)


class LangevinRunner:
    """Runs the Langevin loop for either engine: per step [K2, eps-net kernels, K7], captured
    once in a CUDA graph and replayed; per-step scalars come from a device table indexed by
    a device step counter (the score net ignores the time step)."""

    def __init__(self, engine, ch0, ch1, sched, pos, noise=None, seed=0, atom_offset=0, clip_pos=None,
                 keep_traj=True, use_graph=True, rule=L.RULE_LD, reduce=None, ensemble_size=None, exchange=None):
        """reduce / ensemble_size: the ensemble-member-per-GPU mode.  `engine` holds this rank's members only;
        every step the rank's partial per-atom scores eq_transform(sum of local edge_inv / ensemble_size) are
        summed over the ranks by `reduce(tensor)` (in place, e.g. an NCCL all-reduce) before the update.  All
        ranks hold the same positions and draw the same noise, so they stay in lockstep without a broadcast.
        exchange: a PeerExchange -- the same mode with the exchange FUSED into K7 (peer-memory stores + flags over
        NVLink instead of a collective; no extra kernel, nothing but K7 in the step depends on the peers)."""
        assert sched.dim() == 2 and sched.size(1) == (8 if rule in (L.RULE_DDPM, L.RULE_DDPM_DUALENC) else 4)
        self.engine, self.plan = engine, engine.plan
        dev = self.plan.device
        self.n_steps = sched.size(0)
        self.sched = sched.to(device=dev, dtype=torch.float32).contiguous()
        self.pos = pos
        self.pos0 = pos.clone()
        self.noise = None if noise is None else noise.to(device=dev, dtype=torch.float32).contiguous()
        i32 = dict(dtype=torch.int32, device=dev)
        self.step_counter = torch.zeros(1, **i32)
        self.ticket = torch.zeros(1, **i32)
        self.nan_flag = torch.zeros(1, **i32)
        self.traj = (torch.empty(self.n_steps, self.plan.num_nodes, 3, dtype=torch.float32, device=dev)
                     if keep_traj else None)
        self.ch0, self.ch1 = ch0, ch1
        self.reduce = reduce
        self.node_eq = None
        inv_div = float(engine.num_members)
        assert reduce is None or exchange is None
        self.exchange = exchange
        if reduce is not None:
            assert ch1 is None and ensemble_size is not None and ensemble_size >= engine.num_members
            self.node_eq = torch.zeros(max(self.plan.num_nodes, 1), 3, dtype=torch.float32, device=dev)
            inv_div = float(ensemble_size)
        if exchange is not None:
            assert ch1 is None and ensemble_size is not None and ensemble_size >= engine.num_members
            inv_div = float(ensemble_size)
        self.inv_div = inv_div
        self.ld = L.LdParams(
            self.sched.data_ptr(), self.n_steps, self.step_counter.data_ptr(), self.ticket.data_ptr(),
            self.nan_flag.data_ptr(), self.noise.data_ptr() if self.noise is not None else None,
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(atom_offset), inv_div,
            float(clip_pos) if clip_pos is not None else 0.0,
            self.traj.data_ptr() if self.traj is not None else None, self.n_steps if keep_traj else 0, 0, int(rule),
            self.node_eq.data_ptr() if self.node_eq is not None else None,
            C.pointer(exchange.c) if exchange is not None else None)
        self.use_graph = use_graph
        self.graph = None
        # The trajectory goes to the host while the loop runs (the reference appends pos.cpu() every step,
        # sampler.py:246-247): finished slots in chunks, device -> one of two pinned staging buffers on a copy stream
        # beside the replays, then staging -> the caller's tensor on the host thread, which runs ahead of the GPU anyway.
        # Nothing of the 110 MB is left to copy when the loop ends.  The staging buffers and the stream are allocated once
        # per process and device (_TRAJ_STAGING): allocating / freeing pinned memory per call cost 20 - 600 ms depending
        # on the box (profiles/r3_e2e_phases.txt).  One trajectory at a time per device uses them.
        self._traj_host = None
        self._pending = []  # (first slot, last slot, staging index, copy-done event), oldest first
        self._done = 0      # steps issued since the last reset
        self._copied = 0    # trajectory slots handed to the copy stream

    TRAJ_CHUNK = 256

    def _staging(self):
        dev = torch.device(self.plan.device)
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        rows = 2 * self.TRAJ_CHUNK
        need = rows * max(self.plan.num_nodes, 1) * 3
        st = _TRAJ_STAGING.get(key)
        # another runner with chunks still in the buffers (two trajectories advanced alternately): finish its copies first
        other = st[2]() if st is not None and st[2] is not None else None
        if other is not None and other is not self and other._pending:
            other._drain(0)
            st = _TRAJ_STAGING.get(key)
        if st is None or st[0][0].numel() < need:
            cap = max(need, 1 << 22)  # >= 16 MB each: batches of up to ~2700 atoms share one allocation
            flat = [torch.empty(cap, dtype=torch.float32, pin_memory=True) for _ in range(2)]
            st = (flat, st[1] if st is not None else torch.cuda.Stream(device=dev), None)
        if st[2] is None or st[2]() is not self:
            import weakref
            st = (st[0], st[1], weakref.ref(self))
        _TRAJ_STAGING[key] = st
        bufs = [f[:need].view(rows, max(self.plan.num_nodes, 1), 3) for f in st[0]]
        return bufs, st[1]

    def _drain(self, keep):
        """staging -> host tensor for all but the newest `keep` chunks in flight"""
        while len(self._pending) > keep:
            a, b, i, done = self._pending.pop(0)
            done.synchronize()
            self._traj_host[a:b].copy_(self._staging()[0][i][: b - a])

    def _flush_traj(self):
        if self.traj is None:
            return
        bufs, copy_stream = self._staging()
        if self._traj_host is None:
            self._traj_host = torch.empty(self.traj.shape, dtype=self.traj.dtype)
        end = min(self._done, self.n_steps)
        while self._copied < end:
            a = self._copied
            b = min(end, a + bufs[0].size(0))
            self._drain(1)  # the staging buffer used two chunks ago is free again
            i = 0 if not self._pending else 1 - self._pending[-1][2]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            copy_stream.wait_event(ev)
            done = torch.cuda.Event()
            with torch.cuda.stream(copy_stream):
                bufs[i][: b - a].copy_(self.traj[a:b], non_blocking=True)
                done.record(copy_stream)
            self._pending.append((a, b, i, done))
            self._copied = b

    @_on_plan_device
    def traj_cpu(self):
        """The trajectory (n_steps, N, 3) on the host: the slots of the steps run so far are valid."""
        if self.traj is None:
            return None
        self._flush_traj()
        self._drain(0)
        return self._traj_host

    def _one_step(self):
        lib = L.load()
        self.engine.evaluate(self.pos)
        if self.reduce is not None:
            L.check(lib.tsd_eq_transform(C.byref(self.plan.c_batch), C.byref(self.plan.c_edges), L.ptr(self.pos),
                                         C.byref(self.ch0), self.inv_div, L.ptr(self.node_eq), _stream()),
                    "tsd_eq_transform")
            self.reduce(self.node_eq)
        L.check(lib.tsd_ld_step(C.byref(self.plan.c_batch), C.byref(self.plan.c_edges), L.ptr(self.pos),
                                C.byref(self.ch0), C.byref(self.ch1) if self.ch1 is not None else None,
                                C.byref(self.ld), _stream()), "tsd_ld_step")

    def _reset(self):
        if self.exchange is not None:
            self.exchange.new_trajectory(self.n_steps)
        self.pos.copy_(self.pos0)
        self._drain(0)
        self._done = self._copied = 0
        self.step_counter.zero_()
        self.ticket.zero_()
        self.nan_flag.zero_()

    @_on_plan_device
    def prepare(self):
        """Warm up (loads kernels, outside capture) and capture one step."""
        if not self.use_graph or self.graph is not None:
            return
        side = torch.cuda.Stream(device=self.plan.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._one_step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # with a collective in the step, other threads (the NCCL watchdog) issue CUDA calls during capture
        mode = dict(capture_error_mode="thread_local") if self.reduce is not None else {}
        with torch.cuda.graph(self.graph, **mode):
            self._one_step()
        self._reset()

    @_on_plan_device
    def run(self, n_steps=None, check_every=0):
        """Advance n_steps (default: all).  Raises FloatingPointError like sampler.py:248-250 /
        dualenc.py:959-961 if a NaN position was produced (checked after the chunk)."""
        n = self.n_steps if n_steps is None else n_steps
        self.prepare()
        for k in range(n):
            # (a graph of 8 consecutive steps saved 2 us per step of launch gap, but its capture / destruction cost more
            # per dynamic_sampling call than 5000 steps gain: profiles/r3_e2e_phases.txt)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._one_step()
            self._done += 1
            if self.traj is not None and self._done - self._copied >= self.TRAJ_CHUNK:
                self._flush_traj()
            if check_every and (k + 1) % check_every == 0 and int(self.nan_flag.item()):
                self._raise_flag()
        if int(self.nan_flag.item()):
            self._raise_flag()
        return self.pos

    def _raise_flag(self):
        if int(self.nan_flag.item()) & 2:
            raise L.TsdError("score exchange: a peer's contribution did not arrive within 5 s (rank missing or stalled)")
        raise FloatingPointError()


def ddpm_schedule(betas, t_end, n_steps):
    """(n_steps, 8) coefficient table of the `ddpm` update for i = t_end-1 ... t_end-n_steps and
    j = i-1 (-1 below the first index), evaluated with the reference's own fp32 tensor expressions
    (compute_alpha and sampler.py:216-236) on the CPU: sqrt(at), sqrt(1/at), sqrt(1/at - 1),
    sqrt(atm1) beta_t, sqrt(1 - beta_t) (1 - atm1), 1 - at, mask exp(0.5 log beta_t), sqrt(atm1)."""
    betas = betas.detach().float().cpu()
    cum = (1 - torch.cat([torch.zeros(1), betas], dim=0)).cumprod(dim=0)  # compute_alpha, sampler.py:138-141
    seq = list(range(t_end - n_steps, t_end))
    seq_next = [-1] + seq[:-1]
    rows = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        t = torch.tensor([i], dtype=torch.long)
        at = cum.index_select(0, t + 1)
        atm1 = cum.index_select(0, (torch.ones(1) * j).long() + 1)
        beta_t = 1 - at / atm1
        mask = 1 - (t == 0).float()
        rows.append(torch.cat([at.sqrt(), (1.0 / at).sqrt(), (1.0 / at - 1).sqrt(), atm1.sqrt() * beta_t,
                               (1 - beta_t).sqrt() * (1 - atm1), 1.0 - at, mask * torch.exp(0.5 * beta_t.log()),
                               atm1.sqrt()]))
    return torch.stack(rows).contiguous()


def dualenc_branch_schedule(alphas, betas, n_steps, step_lr, sampling_type, eta=1.0, global_start_sigma=float("inf")):
    """Coefficient table of dualenc.py:861-944 for i = T-1 ... T-n_steps, j = i-1 (-1 below the first index),
    evaluated with the reference's own fp32 tensor expressions on the CPU.
    `generalized`: (n_steps, 4) [step_size_pos, step_size_noise, 0, use_global];
    `ddpm_noisy` / `ddpm_det`: (n_steps, 8) [sqrt(1/at), sqrt(1/at - 1), sqrt(atm1) beta_t,
    sqrt(1 - beta_t) (1 - atm1), 1 - at, mask exp(0.5 logvar), use_global, 0]."""
    alphas, betas = alphas.detach().float().cpu(), betas.detach().float().cpu()
    sigmas = (1.0 - alphas).sqrt() / alphas.sqrt()
    cum = (1 - torch.cat([torch.zeros(1), betas], dim=0)).cumprod(dim=0)  # compute_alpha, dualenc.py:776-780
    t_end = alphas.numel()
    seq = list(range(t_end - n_steps, t_end))
    seq_next = [-1] + seq[:-1]
    rows = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        t = torch.tensor([i], dtype=torch.long)
        at = cum.index_select(0, t + 1)
        at_next = cum.index_select(0, (torch.ones(1) * j).long() + 1)
        use_global = (sigmas[i] < global_start_sigma).float().reshape(1)
        if sampling_type == "generalized":
            c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
            c2 = ((1 - at_next) - c1 ** 2).sqrt()
            pos_ld = step_lr * (sigmas[i] / 0.01) ** 2 / sigmas[i]
            pos_gen = 5 * ((1 - at).sqrt() / at.sqrt() - c2 / at_next.sqrt())
            step_pos = pos_ld if pos_ld < pos_gen else pos_gen
            noise_ld = torch.sqrt((step_lr * (sigmas[i] / 0.01) ** 2) * 2)
            noise_gen = 3 * (c1 / at_next.sqrt())
            step_noise = noise_ld if noise_ld < noise_gen else noise_gen
            rows.append(torch.cat([step_pos.reshape(1), step_noise.reshape(1), torch.zeros(1), use_global]))
        else:
            atm1 = at_next
            beta_t = 1 - at / atm1
            mask = 1 - (t == 0).float()
            logvar = (beta_t * (1 - atm1) / (1 - at)).log() if sampling_type == "ddpm_det" else beta_t.log()
            rows.append(torch.cat([(1.0 / at).sqrt(), (1.0 / at - 1).sqrt(), atm1.sqrt() * beta_t,
                                   (1 - beta_t).sqrt() * (1 - atm1), 1.0 - at, mask * torch.exp(0.5 * logvar),
                                   use_global, torch.zeros(1)]))
    return torch.stack(rows).contiguous().float()


def dsm_schedule(sigmas, n_steps, step_lr, min_sigma=0, global_start_sigma=float("inf")):
    """(levels * n_steps, 4) table [step_size, sigma, sqrt(2 step_size), use_global] of the annealed Langevin loop of
    dualenc.py:1102-1203: for every noise level sigma >= min_sigma (in order; the loop breaks at the first smaller one)
    n_steps rows with step_size = step_lr * (sigma / sigmas[-1]) ** 2, in the reference's fp32 tensor arithmetic."""
    sigmas = sigmas.detach().float().cpu()
    rows = []
    for sigma in sigmas:
        if sigma < min_sigma:
            break
        step_size = step_lr * (sigma / sigmas[-1]) ** 2
        row = torch.stack([step_size, sigma, torch.sqrt(step_size * 2), (sigma < global_start_sigma).float()])
        rows.extend([row] * int(n_steps))
    if not rows:
        return torch.zeros(0, 4)
    return torch.stack(rows).contiguous()


def ld_schedule(alphas, n_steps, step_lr, global_start_sigma=float("inf")):
    """(n_steps, 4) table [step_size, sigma, sqrt(2 step_size), use_global] for i = T-1 ... T-n_steps
    computed with the reference's own fp32 torch expressions (sampler.py:143,239-243):
    step_size = step_lr * (sigmas[i] / 0.01) ** 2."""
    alphas = alphas.detach().float().cpu()
    sigmas = (1.0 - alphas).sqrt() / alphas.sqrt()
    t = sigmas.numel()
    idx = torch.arange(t - 1, t - 1 - n_steps, -1)
    sig = sigmas[idx]
    step_size = step_lr * (sig / 0.01) ** 2
    out = torch.stack([step_size, sig, torch.sqrt(step_size * 2), (sig < global_start_sigma).float()], dim=1)
    return out.contiguous(), sigmas
