import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False))


def build_case_inputs(case):
    """Regenerate the seeded inputs a golden env case was minted from (oracle/make_goldens.py)."""
    from isaacgymloco_b200 import synthetic as S
    from oracle.make_goldens import case_cfg, input_checksum
    cfg = case_cfg(case)
    n = case["n"]
    hf = S.make_terrain(cfg, seed=case["seed"])
    state = S.make_state(cfg, n, hf, seed=case["seed"])
    noise = S.make_noise(n, seed=case["seed"] + 1000)
    targets = S.make_reset_targets(cfg, state, hf, seed=case["seed"] + 2000)
    g = torch.Generator().manual_seed(case["seed"] + 3000)
    delayed = 0.5 * torch.randn(n, 4, 12, generator=g)
    return cfg, hf, state, noise, targets, delayed, input_checksum(state)
