"""TEST INFRASTRUCTURE ONLY.

Imports the reference's own `models/` package VERBATIM from a reference checkout
(default /root/reference; read-only, never copied) by registering stand-ins for the
third-party wheels it needs but this image lacks (torch_geometric, torch_scatter,
torch_sparse, torch_cluster, rdkit, easydict, torchvision).  Real behaviour is supplied
only for the handful of ops the hot path executes (oracle/third_party.py); every other
imported name resolves to an inert dummy class.

Used by tests/golden/make_golden.py (to generate the committed golden vectors) and by
the optional cross-check tests that run only when the reference tree is present.  The
reference tree does not exist on the GPU box, so nothing on the GPU path imports this.
"""
import importlib
import os
import sys
import types

from . import third_party as tp

REFERENCE_ROOT = os.environ.get("TSDIFF_REFERENCE_ROOT", "/root/reference")

# rdkit 2020.09 Chem.rdchem.BondType.names has 22 entries (utils/chem.py:21 builds
# BOND_TYPES from it; models/encoder/edge.py:21 notes "NUM_BOND_TYPES = 22").
RDKIT_BOND_NAMES = [
    "UNSPECIFIED", "SINGLE", "DOUBLE", "TRIPLE", "QUADRUPLE", "QUINTUPLE", "HEXTUPLE",
    "ONEANDAHALF", "TWOANDAHALF", "THREEANDAHALF", "FOURANDAHALF", "FIVEANDAHALF",
    "AROMATIC", "IONIC", "HYDROGEN", "THREECENTER", "DATIVEONE", "DATIVE", "DATIVEL",
    "DATIVER", "OTHER", "ZERO",
]


class AttrDict(dict):
    """easydict.EasyDict stand-in: attribute access, nested dicts converted."""

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


class _StubModule(types.ModuleType):
    """Module whose unknown attributes resolve to fresh inert classes (subclassable,
    callable), so `from pkg import Anything` and `class X(pkg.Base)` both work."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def _module(name, **attrs):
    m = _StubModule(name)
    m.__path__ = []  # behave like a package
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    _module("torch_geometric")
    _module("torch_geometric.nn", MessagePassing=tp.MessagePassing, radius_graph=tp.radius_graph,
            radius=tp.radius)
    _module("torch_geometric.nn.conv", MessagePassing=tp.MessagePassing)
    _module("torch_geometric.nn.inits")
    _module("torch_geometric.nn.models")
    _module("torch_geometric.nn.models.schnet")
    _module("torch_geometric.utils", to_dense_adj=tp.to_dense_adj, dense_to_sparse=tp.dense_to_sparse)
    _module("torch_geometric.data")
    _module("torch_geometric.typing", OptPairTensor=object, Adj=object, OptTensor=object, Size=object)
    _module("torch_geometric.transforms")
    _module("torch_scatter", scatter=tp.scatter, scatter_add=tp.scatter_add,
            scatter_mean=tp.scatter_mean, scatter_max=tp.scatter_max)
    _module("torch_sparse", coalesce=tp.coalesce)
    _module("torch_cluster", radius_graph=tp.radius_graph, radius=tp.radius)
    _module("rdkit")
    _module("rdkit.Chem")
    bond_type = type("BondType", (), {"names": {n: i for i, n in enumerate(RDKIT_BOND_NAMES)}})
    _module("rdkit.Chem.rdchem", BondType=bond_type)
    for sub in ("Draw", "rdDepictor", "PeriodicTable", "rdMolAlign", "rdmolops", "Draw.rdMolDraw2D",
                "AllChem", "rdMolTransforms", "rdForceFieldHelpers"):
        _module("rdkit.Chem." + sub)
    _module("rdkit.RDLogger", DisableLog=lambda *a, **k: None)
    sys.modules["rdkit"].RDLogger = sys.modules["rdkit.RDLogger"]
    _module("torchvision")
    _module("torchvision.transforms")
    _module("torchvision.transforms.functional")
    _module("easydict", EasyDict=AttrDict)
    for name in ("networkx", "sidechainnet", "ase", "py3Dmol"):
        try:
            importlib.import_module(name)
        except Exception:
            _module(name)
    _installed = True


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "epsnet"))


def import_reference():
    """Returns (models.epsnet, models.sampler, models.common, models.geometry) imported
    verbatim from the reference tree."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's top-level packages are called `models` and `utils`
    for name in ("models", "utils"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            raise RuntimeError("a foreign module named %r is already imported" % name)
    epsnet = importlib.import_module("models.epsnet")
    sampler = importlib.import_module("models.sampler")
    common = importlib.import_module("models.common")
    geometry = importlib.import_module("models.geometry")
    return epsnet, sampler, common, geometry


def load_yaml_config(rel_path):
    import yaml
    with open(os.path.join(REFERENCE_ROOT, rel_path)) as f:
        return AttrDict(yaml.safe_load(f))
