"""ctypes binding of include/tsdiff_b200.h (libtsdiff_b200.so).

The product path fails loudly when the CUDA library is missing or a call returns an
error; there is no CPU or PyTorch fallback anywhere behind these functions."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtsdiff_b200.so")

ACT = {"none": 0, "relu": 1, "ReLU": 1, "swish": 2, "ssp": 3, "softplus": 4, "Softplus": 4}
MATH = {"fp32": 0, "tf32": 1}

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)


class Linear(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("in_features", C.c_int32),
                ("out_features", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("num_graphs", C.c_int32), ("max_graph_nodes", C.c_int32),
                ("edge_capacity", C.c_int32), ("graph_ptr", C.c_void_p), ("pair_ptr", C.c_void_p),
                ("node_graph", C.c_void_p)]


class Edges(C.Structure):
    _fields_ = [("num_edges", C.c_void_p), ("row", C.c_void_p), ("col", C.c_void_p), ("length", C.c_void_p),
                ("tab0", C.c_void_p), ("tab1", C.c_void_p), ("in_b", C.c_void_p), ("row_ptr", C.c_void_p),
                ("in_ptr", C.c_void_p), ("in_eid", C.c_void_p), ("in_src", C.c_void_p), ("graph_count", C.c_void_p),
                ("num_upairs", C.c_void_p), ("u_row", C.c_void_p), ("u_col", C.c_void_p), ("u_length", C.c_void_p),
                ("u_tab0", C.c_void_p), ("u_tab1", C.c_void_p), ("edge_upair", C.c_void_p), ("in_upair", C.c_void_p),
                ("graph_ucount", C.c_void_p)]


class EdgeEncoder(C.Structure):
    _fields_ = [("lin0", Linear), ("lin1", Linear), ("bond_emb", C.c_void_p), ("act", C.c_int32),
                ("cat0", C.POINTER(Linear)), ("cat2", C.POINTER(Linear)), ("cat_act", C.c_int32)]


class Interaction(C.Structure):
    _fields_ = [("nn0", Linear), ("nn2", Linear), ("lin1", Linear), ("lin2", Linear), ("lin", Linear),
                ("cutoff", C.c_float), ("smooth", C.c_int32), ("fused_w", C.c_void_p), ("fused_b", C.c_void_p)]


class Gine(C.Structure):
    _fields_ = [("nn0", Linear), ("nn1", Linear), ("eps", C.c_void_p), ("relu_after", C.c_int32)]


class PairMlp(C.Structure):
    _fields_ = [("l0", Linear), ("l1", Linear), ("l2", Linear), ("act", C.c_int32)]


class ScoreChannel(C.Structure):
    _fields_ = [("inv", C.c_void_p), ("mask", C.c_void_p), ("mask_mode", C.c_int32), ("clip", C.c_float),
                ("weight", C.c_float), ("inv_index", C.c_void_p)]


MAX_EXCHANGE_RANKS = 8


class Exchange(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("num_graphs", C.c_int32),
                ("peer_data", C.c_void_p * MAX_EXCHANGE_RANKS), ("peer_flags", C.c_void_p * MAX_EXCHANGE_RANKS),
                ("epoch_base", C.c_void_p)]


class LdParams(C.Structure):
    _fields_ = [("sched", C.c_void_p), ("num_steps", C.c_int32), ("step_counter", C.c_void_p),
                ("ticket", C.c_void_p), ("nan_flag", C.c_void_p), ("noise", C.c_void_p), ("seed", C.c_uint64),
                ("atom_offset", C.c_int64), ("inv_div", C.c_float), ("clip_pos", C.c_float), ("traj", C.c_void_p),
                ("traj_steps", C.c_int32), ("traj_base_step", C.c_int32), ("rule", C.c_int32), ("node_score", C.c_void_p),
                ("exchange", C.POINTER(Exchange))]


RULE_LD, RULE_DDPM, RULE_DDPM_DUALENC, RULE_GENERALIZED, RULE_DSM = 0, 1, 2, 3, 4


# symbol -> argtypes; every function returns int (0 = TSD_OK) unless listed in _RESTYPES
_P = C.c_void_p
_SIGNATURES = {
    "tsd_version": [],
    "tsd_launch_count": [],
    "tsd_workspace_bytes": [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P],
    "tsd_linear": [C.c_int32, _P, _P, C.POINTER(Linear), C.c_int32, _P, C.c_int32, _P],
    "tsd_round_tf32": [_P, _P, C.c_int64, _P],
    "tsd_cfconv_aggregate": [C.POINTER(Batch), C.POINTER(Edges), C.c_int32, _P, _P, _P, _P],
    "tsd_last_cuda_error": [],
    "tsd_error_string": [C.c_int],
    "tsd_bond_order_build": [C.c_int, C.POINTER(Batch), C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P,
                             _P, _P, _P],
    "tsd_edge_build": [C.POINTER(Batch), _P, C.c_double, C.c_int32, _P, _P, C.c_int32, C.POINTER(Edges), _P],
    "tsd_condensed_node_embed": [C.c_int32, _P, _P, _P, C.c_int32, _P, _P, C.c_int32, _P, _P],
    "tsd_embedding": [C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_float, _P, _P],
    "tsd_edge_embed": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(EdgeEncoder), C.c_int32, _P, _P, _P,
                       C.c_int32, _P],
    "tsd_cfconv_layer": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(Interaction), _P, _P, _P, _P, _P, _P, _P,
                         C.c_int32, _P],
    "tsd_filter_network": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(Interaction), _P, _P, C.c_int32, _P],
    "tsd_filter_stack": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(Interaction), C.c_int32, C.POINTER(C.c_void_p), _P],
    "tsd_schnet_encoder": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(Interaction), C.c_int32, _P, _P, _P, _P,
                           _P, _P, _P, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P],
    "tsd_interaction_node_update": [C.POINTER(Batch), C.POINTER(Edges), C.POINTER(Interaction), C.POINTER(Linear), _P, _P, _P,
                                    _P, _P, _P],
    "tsd_gine_layer": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(Gine), _P, _P, _P, _P, C.c_int32, _P],
    "tsd_pair_mlp": [C.POINTER(Batch), C.POINTER(Edges), _P, _P, C.POINTER(PairMlp), C.c_int32, _P, _P, C.c_int32,
                     _P],
    "tsd_edge_embed_delta": [C.POINTER(Batch), C.POINTER(Edges), _P, _P, C.POINTER(EdgeEncoder), _P, _P, _P, _P, _P, _P,
                             C.c_int32, _P],
    "tsd_pair_mlp_delta": [C.POINTER(Batch), C.POINTER(Edges), _P, _P, _P, _P, C.POINTER(PairMlp), C.c_int32, _P, _P,
                           C.c_int32, _P],
    "tsd_ld_step": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(ScoreChannel), C.POINTER(ScoreChannel),
                    C.POINTER(LdParams), _P],
    "tsd_eq_transform": [C.POINTER(Batch), C.POINTER(Edges), _P, C.POINTER(ScoreChannel), C.c_float, _P, _P],
    "tsd_philox_normal": [C.c_int32, C.c_uint64, C.c_int32, C.c_int64, _P, _P],
    "tsd_act_forward": [C.c_int64, _P, C.c_int32, _P, _P],
    "tsd_act_backward": [C.c_int64, _P, _P, C.c_int32, _P, _P],
    "tsd_row_scale": [C.c_int32, C.c_int32, _P, _P, _P, _P],
    "tsd_cutoff_envelope": [C.c_int32, _P, C.c_float, C.c_int32, _P, _P],
    "tsd_gate_rows": [C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P],
    "tsd_onehot": [C.c_int32, C.c_int32, _P, C.c_int32, _P, _P],
    "tsd_transpose": [C.c_int32, C.c_int32, _P, _P, _P],
    "tsd_linear_wgrad_scratch": [C.c_int32, C.c_int32, C.c_int32, _P],
    "tsd_linear_wgrad": [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P],
    "tsd_cfconv_aggregate_backward": [C.POINTER(Batch), C.POINTER(Edges), C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P],
    "tsd_pair_features": [C.POINTER(Edges), C.c_int32, C.c_int32, _P, _P, _P, _P],
    "tsd_pair_features_backward": [C.POINTER(Batch), C.POINTER(Edges), C.c_int32, _P, _P, _P, _P],
    "tsd_eq_transform_backward": [C.POINTER(Edges), C.c_int32, _P, _P, C.c_int32, C.c_float, _P, _P, _P],
    "tsd_condensed_node_embed_backward": [C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P],
    "tsd_sqerr_forward": [C.c_int32, _P, _P, _P, _P],
    "tsd_sqerr_backward": [C.c_int32, _P, _P, _P, _P, _P],
    "tsd_add": [C.c_int64, _P, _P, _P, _P],
    "tsd_mul": [C.c_int64, _P, _P, _P, _P],
    "tsd_peer_alloc": [C.c_uint64, C.POINTER(C.c_void_p), _P],
    "tsd_peer_open": [_P, C.POINTER(C.c_void_p)],
    "tsd_peer_close": [_P],
    "tsd_peer_free": [_P],
    "tsd_dmae": [C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P],
    "tsd_dmae_pos": [C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P],
    "tsd_min_match_scratch": [C.c_int32, C.c_int32, _P, _P],
    "tsd_min_match": [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P],
}
_RESTYPES = {"tsd_error_string": C.c_char_p, "tsd_launch_count": C.c_int64}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class TsdError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it has not been built -- run
    `python -c "import __graft_entry__ as g; g.build()"` or `python tsdiff_b200/build.py`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TsdError("%s is missing: build it with `python tsdiff_b200/build.py` (needs nvcc); "
                       "tsdiff_b200 has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drifted apart
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.tsd_error_string(rc)
        raise TsdError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", rc))


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def linear(weight, bias=None):
    assert weight.is_contiguous() and (bias is None or bias.is_contiguous())
    return Linear(weight.data_ptr(), bias.data_ptr() if bias is not None else None, weight.shape[1], weight.shape[0])
