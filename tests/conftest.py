import os
import sys

# The loopback tests run several "ranks" as threads of this process on ONE GPU, each on its own stream, with kernels that
# wait for each other (the peer step).  CUDA maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8):
# with more streams than queues two ranks can share a queue and the spinning kernel of one blocks the kernels of the other
# behind it -- a false dependency that never resolves.  Must be set before the CUDA context exists.  (Test-only: a real
# job is one process per GPU with one compute stream.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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


# ---------------------------------------------------------------- kernel back-ends
# "cuda": the product library on a real GPU (marked gpu).
# "emul": the SAME kernel sources compiled for the test-only host emulator
#         (tests/emul) -- a debugging aid for the GPU-less build container; it is not
#         a product path and is skipped automatically if g++ cannot build it.
_EMUL = {}


def _emul_ops():
    if "ops" not in _EMUL:
        import subprocess
        from b200rl import native, ops
        script = os.path.join(ROOT, "tests", "emul", "build_emul.sh")
        r = subprocess.run(["bash", script], capture_output=True, text=True)
        if r.returncode != 0:
            _EMUL["ops"] = None
            _EMUL["err"] = r.stderr[-2000:]
        else:
            lib = native.Lib(os.path.join(ROOT, "tests", "emul", "_build", "libb200rl_emul.so"), emulated=True)
            _EMUL["ops"] = ops.Ops(lib, allow_cpu=True)
    if _EMUL["ops"] is None:
        import shutil
        if shutil.which("g++") is None:
            pytest.skip("no g++: the host emulator cannot be built")
        pytest.fail("emulator build failed (a kernel source does not compile for the host emulator):\n" + _EMUL.get("err", ""))
    return _EMUL["ops"]


@pytest.fixture(params=["emul", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """(ops, device) for each kernel back-end."""
    if request.param == "emul":
        return _emul_ops(), torch.device("cpu")
    from b200rl import ops
    return ops.default(), torch.device("cuda:0")
