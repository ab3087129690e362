import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout (build container only)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return torch.load(os.path.join(GOLDEN, "golden_outputs.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_ddpm():
    """Non-default sampler branches (tests/golden/make_golden_ddpm.py)."""
    return torch.load(os.path.join(GOLDEN, "golden_ddpm.pt"), weights_only=False)


# keyword arguments every golden_ddpm case was generated with (make_golden_ddpm.py)
DDPM_CASES = {
    "b_rxn0_ddpm20": dict(sampling_type="ddpm"),
    "b_syn4_ddpm10": dict(sampling_type="ddpm"),
    "b_rxn0_ddpm_t12": dict(sampling_type="ddpm", denoise_from_time_t=12),
    "b_rxn0_guess_ld8": dict(sampling_type="ld", denoise_from_time_t=3000, noise_from_time_t=1500),
    "b_rxn0_guess_ddpm8": dict(sampling_type="ddpm", denoise_from_time_t=600, noise_from_time_t=0),
    "b_syn4_ld6_clip_pos": dict(sampling_type="ld", clip_pos=20.0),
}


@pytest.fixture(scope="session")
def golden_dualenc_branches():
    """dualenc.py:861-944 branches (tests/golden/make_golden_dualenc_branches.py)."""
    return torch.load(os.path.join(GOLDEN, "golden_dualenc_branches.pt"), weights_only=False)


DUALENC_BRANCH_CASES = {
    "a_rxn0_ddpm_noisy10": dict(clip=10.0, clip_local=10.0, sampling_type="ddpm_noisy"),
    "a_syn4_ddpm_det6": dict(clip=10.0, clip_local=10.0, sampling_type="ddpm_det"),
    "a_rxn0_generalized8": dict(clip=10.0, clip_local=10.0, sampling_type="generalized", eta=1.0),
    "a_syn4_generalized6_eta05": dict(clip=10.0, clip_local=10.0, w_global=0.5, sampling_type="generalized", eta=0.5),
}


@pytest.fixture(scope="session")
def golden_loss():
    """get_loss values (tests/golden/make_golden_loss.py)."""
    return torch.load(os.path.join(GOLDEN, "golden_loss.pt"), weights_only=False)


@pytest.fixture(scope="session")
def rxn0():
    return torch.load(os.path.join(GOLDEN, "rxn0_graph.pt"), weights_only=False)


@pytest.fixture(scope="session")
def syn4():
    return torch.load(os.path.join(GOLDEN, "syn4_graph.pt"), weights_only=False)


def graph_for(name, rxn0, syn4):
    return rxn0 if "rxn0" in name else syn4
