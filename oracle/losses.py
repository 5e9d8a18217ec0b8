"""Oracle restatement of `agents/learner_module/compute_loss.py` (reference).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import torch


def compute_gae(deltas, gamma, lambda_):
    """Reverse scan g[t] = d[t] + gamma*lambda*g[t+1], g[T-1] = 0 beyond the end.

    Follows reference `compute_loss.py:7-19`.  NB the recurrence does NOT look at
    `is_fir` (quirk A.6-1); only the caller's td target does.
    deltas: (B, T-1, 1) -> (B, T-1, 1)
    """
    n = deltas.size(1)
    out = torch.empty_like(deltas)
    carry = torch.zeros_like(deltas[:, 0])
    k = gamma * lambda_
    for t in range(n - 1, -1, -1):
        carry = deltas[:, t] + k * carry
        out[:, t] = carry
    return out


def compute_v_trace(behav_log_probs, target_log_probs, is_fir, rewards, values,
                    gamma, rho_bar=0.8, c_bar=1.0):
    """V-trace targets.  Follows reference `compute_loss.py:22-66`.

    rho is clamped to [0.1, rho_bar] (both sides, quirk A.6-2), c to (-inf, c_bar].
    All inputs (B, T, 1).  Returns (rho_clipped (B,T-1,1), advantages (B,T-1,1),
    values_target (B,T,1)).
    """
    rho = torch.exp(target_log_probs[:, :-1] - behav_log_probs[:, :-1])
    rho_c = rho.clamp(min=0.1, max=rho_bar)
    c_c = rho.clamp(max=c_bar)
    disc = gamma * (1.0 - is_fir)                      # (B,T,1); used at index t+1
    td = rewards[:, :-1] + disc[:, 1:] * values[:, 1:]
    deltas = rho_c * (td - values[:, :-1])
    acc = torch.zeros_like(values)                     # vs - V, last step stays 0
    for t in range(deltas.size(1) - 1, -1, -1):
        acc[:, t] = deltas[:, t] + c_c[:, t] * disc[:, t + 1] * acc[:, t + 1]
    vs = values + acc
    adv = rho_c * (rewards[:, :-1] + disc[:, 1:] * vs[:, 1:] - values[:, :-1])
    return rho_c, adv, vs


def soft_update(params, target_params, tau=0.005):
    """Polyak update theta' <- (1-tau) theta' + tau theta, in place on target.

    Follows reference `compute_loss.py:69-71`.  `params` / `target_params` are
    dicts name -> tensor (they may alias: the reference's target critic IS the
    critic, quirk A.6-8).
    """
    with torch.no_grad():
        for k in params:
            p, tp = params[k], target_params[k]
            tp.copy_((1.0 - tau) * tp + tau * p)


def _categorical_from_probs(probs):
    """torch.distributions.Categorical(probs=...) restated: renormalise, then
    logits = log(clamp(p, eps, 1-eps)) (probs_to_logits)."""
    q = probs / probs.sum(-1, keepdim=True)
    eps = torch.finfo(q.dtype).eps
    logits = torch.log(q.clamp(min=eps, max=1.0 - eps))
    return q, logits


def kldivergence(logits_p, logits_q):
    """KL(Cat(softmax p) || Cat(softmax q)), `.squeeze()`d.

    Follows reference `compute_loss.py:74-77`; kl_divergence(Categorical,
    Categorical) = sum p*(logp - logq) with the 0/inf fix-ups of torch's
    `_kl_categorical_categorical`.
    """
    p, lp = _categorical_from_probs(torch.softmax(logits_p, -1))
    q, lq = _categorical_from_probs(torch.softmax(logits_q, -1))
    t = p * (lp - lq)
    t = torch.where(q == 0, torch.full_like(t, float("inf")), t)
    t = torch.where(p == 0, torch.zeros_like(t), t)
    return t.sum(-1).squeeze()
