"""CPU oracle for the learner-update path -- TEST INFRASTRUCTURE ONLY.

This package is a torch-CPU / fp32 restatement of the reference's learner update
(`/root/reference/agents/learner_module/*`, `networks/models.py`).  It exists to
CHECK the CUDA path; it is never on the product path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.

Parity status: the reference ships no tests / golden vectors of its own
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference
itself, imported and run in the build container by `tests/golden/make_golden.py`
(fixtures committed under `tests/golden/*.npz`, checked by
`tests/test_oracle_golden.py`).

All arithmetic the reference delegates to PyTorch (LSTMCell, Categorical, Normal,
smooth_l1_loss, clip_grad_norm_, RMSprop, Adam) is restated explicitly here so the
CUDA kernels have a formula-level specification to match.
"""
from .losses import compute_gae, compute_v_trace, soft_update, kldivergence  # noqa: F401
from .models import (  # noqa: F401
    init_params,
    actor_critic_forward,
    sac_actor_forward,
    sac_actor_continuous_forward,
    sac_critic_forward,
    sac_critic_continuous_forward,
)
from .steps import OracleLearner, DEFAULT_HYPER  # noqa: F401
