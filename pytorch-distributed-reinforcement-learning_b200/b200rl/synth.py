"""Seeded synthetic rollouts in the reference's batch layout.

The learner benchmark replaces Worker/Manager/LearnerStorage with pre-generated
rollouts of the shape `LearnerStorage.make_batch` would have written
(`agents/learner_storage.py:123-159`, widths from
`agents/storage_module/shared_batch.py:19-64`): every field fp32, batch-major
`(B, T, width)`; `hx`/`cx` are the LSTM state BEFORE step 0 of each window, so only
`(B, H)` is materialised (the reference ships `(B, T, H)` and reads `[:, 0]`,
`ppo/learning.py:37-40`).
"""
import math

import torch

FIELDS = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")


def make_rollouts(B, T, f, A, H=64, continuous=False, batch_index=0, act_dim=None):
    """Return a dict of CPU fp32 tensors.  Same seed -> same bytes on every host."""
    g = torch.Generator().manual_seed(1234 + batch_index)
    n = lambda *s: torch.randn(*s, generator=g)
    u = lambda *s: torch.rand(*s, generator=g)
    obs = n(B, T, f)
    rew = 0.1 * n(B, T, 1)
    if not continuous:
        act = torch.randint(0, A, (B, T, 1), generator=g).float()
        logits = torch.log_softmax(n(B, T, A), -1)          # behaviour policy: normalised log-probs (models.py:46-50)
        log_prob = logits.gather(-1, act.long())
    else:
        Ad = act_dim or A
        act = 2 * u(B, T, Ad) - 1
        mu = 2 * u(B, T, Ad) - 1
        sd = 0.5 + u(B, T, Ad)
        log_prob = (-((act - mu) ** 2) / (2 * sd * sd) - sd.log() - 0.5 * math.log(2 * math.pi))
        log_prob = log_prob[..., :1].contiguous() if Ad == 1 else log_prob
        logits = torch.zeros(B, T, Ad)                      # continuous nets publish zeros (models.py:90-94)
    is_fir = (u(B, T, 1) < 0.02).float()
    hx = 0.1 * n(B, H)
    cx = 0.1 * n(B, H)
    return dict(obs=obs, act=act, rew=rew, logits=logits, log_prob=log_prob, is_fir=is_fir, hx=hx, cx=cx)


def as_reference_tuple(batch, T):
    """The 8-tuple `sample_batch_from_sh_memory` returns (`agents/learner.py:197-233`),
    with hx/cx broadcast back to (B, T, H)."""
    hx = batch["hx"][:, None, :].expand(-1, T, -1).contiguous()
    cx = batch["cx"][:, None, :].expand(-1, T, -1).contiguous()
    return (batch["obs"], batch["act"], batch["rew"], batch["logits"], batch["log_prob"],
            batch["is_fir"], hx, cx)
