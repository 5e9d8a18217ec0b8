"""The recurrent kernels' two tile widths (32- and 16-sequence tiles; `recurrent_tile16`, csrc/fwd_fused.cu, csrc/bwd_tc.cu).
In the ordinary emulator run a batch gets 16-sequence tiles only when it has at most two 32-sequence tiles (the emulated
device has 4 SMs), i.e. one tile per CTA.  Here `B200RL_TILE16` forces either width in a child process on a batch that gives
every CTA several tiles of a ragged batch, for PPO and V-MPO updates against the oracle (`ppo/learning.py:43-106`,
`v_mpo/learning.py:44-124`) and against each other."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, PKG

_CHILD = r"""
import os, sys
import numpy as np, torch
ROOT, PKG, algo, B, T, out = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import oracle
from b200rl import native, ops as O
from b200rl.engine import OnPolicyEngine, VMPOEngine
from b200rl.networks import models as M
from b200rl.synth import make_rollouts
lib = native.Lib(os.path.join(ROOT, "tests", "emul", "_build", "libb200rl_emul.so"), emulated=True)
ops = O.Ops(lib, allow_cpu=True)
torch.manual_seed(0)
model = M.MlpLSTMSingle(4, 2, T, 64)
batch = make_rollouts(B, T, 4, 2)
if algo == "V-MPO":
    eng = VMPOEngine(model.actor, oracle.DEFAULT_HYPER, ops=ops)
    for _ in range(2):
        eng.update(batch, 0.0075)
else:
    eng = OnPolicyEngine(model.actor, algo, oracle.DEFAULT_HYPER, ops=ops)
    for _ in range(2):
        eng.update(batch)
L = eng.losses()
params = torch.cat([v.reshape(-1) for v in model.actor.views_of(eng.flat).values()])
np.savez(out, params=params.numpy(), loss=L["loss"], head=eng.head_out.numpy())
"""


def _child(algo, B, T, tile16, tmp_path):
    out = str(tmp_path / f"tile{tile16}_{algo}.npz")
    env = dict(os.environ, B200RL_TILE16=str(tile16))
    r = subprocess.run([sys.executable, "-c", _CHILD, ROOT, PKG, algo, str(B), str(T), out], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


@pytest.mark.parametrize("algo", ["PPO", "V-MPO"])
def test_tile_widths_agree_with_the_oracle_and_with_each_other(algo, tmp_path):
    import oracle
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    from conftest import _emul_ops
    _emul_ops()                                               # builds the emulated library once
    B, T = 150, 5                                             # 10 ragged 16-sequence tiles / 5 32-sequence tiles over 4 CTAs, odd window
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(4, 2, T, 64)
    sd = {k: v.clone() for k, v in model.actor.state_dict().items()}
    orc = oracle.OracleLearner(algo, sd)
    batch = make_rollouts(B, T, 4, 2)
    for _ in range(2):
        want = orc.update(batch, coef_alpha=torch.tensor(0.0075)) if algo == "V-MPO" else orc.update(batch)
    got = {w: _child(algo, B, T, w, tmp_path) for w in (1, 0)}
    for w, g in got.items():
        np.testing.assert_allclose(g["loss"], float(want["loss"]), rtol=1e-4, atol=1e-7, err_msg=f"B200RL_TILE16={w}")
        np.testing.assert_allclose(g["params"], want["params"].numpy(), rtol=1e-4, atol=1e-6, err_msg=f"B200RL_TILE16={w}")
    # per sequence the two widths do the same arithmetic; only the order of the weight-gradient partial sums differs (one slab
    # row per CTA), so after two updates parameters and head outputs agree to rounding
    np.testing.assert_allclose(got[0]["head"], got[1]["head"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got[0]["params"], got[1]["params"], rtol=1e-5, atol=1e-7)
