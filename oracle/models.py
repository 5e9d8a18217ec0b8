"""Oracle restatement of `networks/models.py` (reference) in functional form.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Parameters are plain dicts keyed by the reference's state-dict names
(SURVEY.md 8b-B3): `body.0.weight (H,f)`, `body.0.bias`, `lstmcell.weight_ih (4H,H)`,
`lstmcell.weight_hh (4H,H)`, `lstmcell.bias_ih`, `lstmcell.bias_hh`, then the heads.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .losses import _categorical_from_probs

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- init
def _lin(out_f, in_f):
    m = nn.Linear(in_f, out_f)
    return m.weight.detach().clone(), m.bias.detach().clone()


def _lstm(H):
    m = nn.LSTMCell(H, H)
    return (m.weight_ih.detach().clone(), m.weight_hh.detach().clone(),
            m.bias_ih.detach().clone(), m.bias_hh.detach().clone())


def _put_lstm(p, t):
    p["lstmcell.weight_ih"], p["lstmcell.weight_hh"], p["lstmcell.bias_ih"], p["lstmcell.bias_hh"] = t


def init_params(kind, f, A, H=64):
    """Build one network's parameters, consuming the global torch RNG in the same
    order as the reference constructors do (so `torch.manual_seed(s)` followed by
    the same sequence of calls reproduces the reference's weights).

    kind: "single" (`MlpLSTMBase`, models.py:8-35), "single_continuous" (:103-112),
          "actor" (:121-134), "actor_continuous" (:162-180), "critic" (:234-252),
          "critic_continuous" (:273-294; `act_dim` = A generalises the hard-coded 1).
    """
    p = OrderedDict()
    w, b = _lin(H, f)                                  # body, models.py:16-19
    p["body.0.weight"], p["body.0.bias"] = w, b
    if kind in ("single", "single_continuous"):
        _put_lstm(p, _lstm(H))
        p["value.weight"], p["value.bias"] = _lin(1, H)
        p["logits.weight"], p["logits.bias"] = _lin(A, H)
        if kind == "single_continuous":
            p["logits_std.weight"], p["logits_std.bias"] = _lin(A, H)
    elif kind == "actor":
        _put_lstm(p, _lstm(H))
        p["logits.weight"], p["logits.bias"] = _lin(A, H)
    elif kind == "actor_continuous":
        _put_lstm(p, _lstm(H))
        p["logits.weight"], p["logits.bias"] = _lin(A, H)
        p["logits_log_std.weight"], p["logits_log_std.bias"] = _lin(A, H)
    elif kind == "critic":
        _put_lstm(p, _lstm(H))
        p["q_value.weight"], p["q_value.bias"] = _lin(A, H)
        p["body.0.weight"], p["body.0.bias"] = _lin(H, f)       # body re-created, models.py:239-242
    elif kind == "critic_continuous":
        _put_lstm(p, _lstm(H))
        p["encode_action.0.weight"], p["encode_action.0.bias"] = _lin(H // 2, A)
        p["q_value.weight"], p["q_value.bias"] = _lin(1, H)
        p["body.0.weight"], p["body.0.bias"] = _lin(H // 2, f)  # models.py:278-284
    else:
        raise ValueError(kind)
    return p


def init_twin(kind, f, A, H=64):
    """`MlpLSTMDoubleCritic(Continuous)` (models.py:325-342): q1 then q2."""
    out = OrderedDict()
    for name in ("q1", "q2"):
        for k, v in init_params(kind, f, A, H).items():
            out[f"{name}.{k}"] = v
    return out


def sub(params, prefix):
    n = len(prefix)
    return OrderedDict((k[n:], v) for k, v in params.items() if k.startswith(prefix))


# ------------------------------------------------------------------------ LSTM core
def lstm_cell(x, h, c, p):
    """nn.LSTMCell restated; PyTorch gate order i, f, g, o."""
    H = h.shape[-1]
    pre = (x @ p["lstmcell.weight_ih"].t() + p["lstmcell.bias_ih"]
           + h @ p["lstmcell.weight_hh"].t() + p["lstmcell.bias_hh"])
    i = torch.sigmoid(pre[:, 0 * H:1 * H])
    f = torch.sigmoid(pre[:, 1 * H:2 * H])
    g = torch.tanh(pre[:, 2 * H:3 * H])
    o = torch.sigmoid(pre[:, 3 * H:4 * H])
    c2 = f * c + i * g
    h2 = o * torch.tanh(c2)
    return h2, c2


def lstm_unroll(x, h, c, p, fused_cell=False):
    """T sequential cell steps over x (B,T,H); the state is NOT reset at is_fir
    (models.py:71-75, quirk A.6-6).  `fused_cell=True` routes through
    torch's own `_VF.lstm_cell` (what nn.LSTMCell calls) -- used by the timing
    baseline so its op stream equals the reference's."""
    outs = []
    for t in range(x.shape[1]):
        if fused_cell:
            h, c = torch._VF.lstm_cell(x[:, t], (h, c), p["lstmcell.weight_ih"], p["lstmcell.weight_hh"],
                                       p["lstmcell.bias_ih"], p["lstmcell.bias_hh"])
        else:
            h, c = lstm_cell(x[:, t], h, c, p)
        outs.append(h)
    return torch.stack(outs, 1)


def _torso(obs, p):
    B, T = obs.shape[:2]
    x = torch.relu(obs.reshape(B * T, -1) @ p["body.0.weight"].t() + p["body.0.bias"])
    return x.view(B, T, -1)


# ----------------------------------------------------------- on-policy actor-critic
def actor_critic_forward(p, obs, hx, cx, act, continuous=False, fused_cell=False):
    """`MlpLSTMBase.forward` / `MlpLSTMContinuous` (models.py:63-118).

    obs (B,T,f), hx/cx (B,H) = state BEFORE step 0, act (B,T,1).
    Returns (logits (B,T,A) = NORMALISED log-probs for the discrete net / zeros for
    the continuous one, log_prob(act) (B,T,1|A), entropy (B,T,1|A), value (B,T,1)).
    """
    B, T = obs.shape[:2]
    out = lstm_unroll(_torso(obs, p), hx, cx, p, fused_cell)
    value = out @ p["value.weight"].t() + p["value.bias"]
    z = out @ p["logits.weight"].t() + p["logits.bias"]
    if not continuous:
        q, logits = _categorical_from_probs(torch.softmax(z, -1))
        a = act.squeeze(-1).long()
        logp = logits.gather(-1, a.unsqueeze(-1))
        ent = -(logits.clamp(min=torch.finfo(logits.dtype).min) * q).sum(-1, keepdim=True)
        return logits, logp, ent, value
    mu = torch.tanh(z)
    std = F.softplus(out @ p["logits_std.weight"].t() + p["logits_std.bias"])
    logp = -((act - mu) ** 2) / (2 * std ** 2) - std.log() - LOG_SQRT_2PI
    ent = 0.5 + LOG_SQRT_2PI + std.log()
    return torch.zeros_like(act), logp, ent, value


# ---------------------------------------------------------------------- SAC networks
def sac_actor_forward(p, obs, hx, cx, fused_cell=False):
    """`MlpLSTMActor.forward` (models.py:135-159): probs, log(probs + 1e-8*[p==0])."""
    out = lstm_unroll(_torso(obs, p), hx, cx, p, fused_cell)
    probs = torch.softmax(out @ p["logits.weight"].t() + p["logits.bias"], -1)
    z = (probs == 0.0).float() * 1e-8
    return probs, torch.log(probs + z)


def sac_actor_continuous_forward(p, obs, hx, cx, eps, fused_cell=False):
    """`MlpLSTMActorContinuous.forward` (models.py:199-231) with the rsample noise
    `eps` (B,T,A) injected: x = mu + sigma*eps, a = tanh(x),
    logp = N(x; mu, sigma).log_prob - log(1 - a^2 + 1e-7)."""
    out = lstm_unroll(_torso(obs, p), hx, cx, p, fused_cell)
    mu = out @ p["logits.weight"].t() + p["logits.bias"]
    log_std = (out @ p["logits_log_std.weight"].t() + p["logits_log_std.bias"]).clamp(-20, 2)
    std = log_std.exp()
    x = mu + std * eps
    a = torch.tanh(x)
    logp = -((x - mu) ** 2) / (2 * std ** 2) - std.log() - LOG_SQRT_2PI
    logp = logp - torch.log(1 - a.pow(2) + 1e-7)
    return a, logp


def sac_critic_forward(p, obs, hx, cx, fused_cell=False):
    """`MlpLSTMCritic.forward` (models.py:254-270): Q (B,T,A)."""
    out = lstm_unroll(_torso(obs, p), hx, cx, p, fused_cell)
    return out @ p["q_value.weight"].t() + p["q_value.bias"]


def sac_critic_continuous_forward(p, obs, act, hx, cx, fused_cell=False):
    """`MlpLSTMCriticContinuous.forward` (models.py:296-322): obs->H/2, act->H/2,
    concat, LSTM, Q (B,T,1).  The action encoder width follows `act`'s last dim
    (the reference hard-codes 1, models.py:290)."""
    B, T = obs.shape[:2]
    xo = _torso(obs, p)
    xa = torch.relu(act.reshape(B * T, -1) @ p["encode_action.0.weight"].t() + p["encode_action.0.bias"]).view(B, T, -1)
    out = lstm_unroll(torch.cat([xo, xa], -1), hx, cx, p, fused_cell)
    return out @ p["q_value.weight"].t() + p["q_value.bias"]
