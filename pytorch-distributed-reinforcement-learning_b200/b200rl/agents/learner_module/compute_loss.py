"""Drop-in for the reference's `agents/learner_module/compute_loss.py`: same four
functions, same signatures, same shapes -- running the sm_100a kernels of
`csrc/scan_kernels.cu` through the C-ABI.  Pure functions of CUDA tensors; like the
reference call sites (`ppo/learning.py:48`, `impala/learning.py:48`) they carry no grad.
"""
import torch

from ... import ops as _ops


def compute_gae(deltas, gamma, lambda_):
    """(B, T-1, 1) -> (B, T-1, 1).  reference compute_loss.py:7-19"""
    return _ops.default().compute_gae(deltas.detach(), float(gamma), float(lambda_))


def compute_v_trace(behav_log_probs, target_log_probs, is_fir, rewards, values, gamma, rho_bar=0.8, c_bar=1.0):
    """-> (rho_clipped (B,T-1,1), advantages (B,T-1,1), values_target (B,T,1)).  reference compute_loss.py:22-66"""
    d = lambda t: t.detach()
    return _ops.default().compute_v_trace(d(behav_log_probs), d(target_log_probs), d(is_fir), d(rewards), d(values),
                                          float(gamma), float(rho_bar), float(c_bar))


def soft_update(critic, target_critic, tau=0.005):
    """theta' <- (1-tau) theta' + tau theta for every parameter pair.  reference compute_loss.py:69-71.
    Networks from `b200rl.networks.models` are updated with one launch over their flat buffers."""
    ops = _ops.default()
    if hasattr(critic, "flat_buffers") and hasattr(target_critic, "flat_buffers"):
        for s, t in zip(critic.flat_buffers(), target_critic.flat_buffers()):
            ops.soft_update_flat(s, t, float(tau))
        return
    for p, tp in zip(critic.parameters(), target_critic.parameters()):
        ops.soft_update_flat(p.data.contiguous().view(-1), tp.data.view(-1), float(tau))


def kldivergence(logits_p, logits_q):
    """KL(Cat(softmax p) || Cat(softmax q)).squeeze().  reference compute_loss.py:74-77"""
    return _ops.default().kldivergence(logits_p.detach(), logits_q.detach())
