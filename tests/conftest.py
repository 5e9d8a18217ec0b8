import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_batch(g, i, device="cpu"):
    keys = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")
    return {k: torch.from_numpy(g[f"in/{i}/{k}"]).to(device) for k in keys}


def golden_sd(g, prefix):
    from collections import OrderedDict
    n = len(prefix) + 1
    return OrderedDict((k[n:], torch.from_numpy(v)) for k, v in g.items() if k.startswith(prefix + "/"))
