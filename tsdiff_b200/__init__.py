"""tsdiff_b200 -- B200-native drop-in for the Langevin-dynamics eps-net hot path of
seonghann/tsdiff.  Host side mirrors the reference's Python API
(`models.epsnet.get_model`, the eps-net `forward()` signatures, `EnsembleSampler`);
all compute runs in hand-written sm_100a CUDA kernels behind the C-ABI library
`libtsdiff_b200.so` (include/tsdiff_b200.h).  There is no CPU fallback."""

__version__ = "0.1.0"
