"""One-entry cache of the engine of the last batch, so that calling `forward` repeatedly on the
same batch (as the reference's sampler loops do, models/sampler.py:194-206) does not rebuild the
bond-order tables every step.

The engine freezes state derived from the WEIGHTS (the condensed node embedding, the TF32 shadow
weights, the folded lin1.lin products, raw device pointers into the parameter storage), so the key
also carries every parameter's (data_ptr, _version): `load_state_dict`, an optimizer step or any
other in-place update bumps `_version`; `.to()` / `.float()` reallocate and change `data_ptr`.
Either way the next call builds a fresh engine (train.py:154-184 validates between optimizer steps)."""


def parameter_signature(modules):
    """(data_ptr, _version) of every parameter and buffer of the given modules, in registration order."""
    sig = []
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            sig.append((t.data_ptr(), t._version))
    return tuple(sig)


class EngineCache:
    def __init__(self):
        self.key = None
        self.tensors = None
        self.engine = None

    @staticmethod
    def _sig(t):
        return (t.data_ptr(), tuple(t.shape), t.dtype, t._version, str(t.device))

    def get(self, tensors, extra, build, modules=()):
        key = tuple(self._sig(t) for t in tensors) + tuple(extra) + parameter_signature(modules)
        if key != self.key:
            self.engine = None  # release the old engine's buffers before building the new one
            self.engine = build()
            self.key = key
            self.tensors = tensors  # keep them alive so data_ptr cannot be recycled
        return self.engine

    def clear(self):
        self.key = self.tensors = self.engine = None
