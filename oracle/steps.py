"""Oracle restatement of one learner update per algorithm.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Each `_*_update` follows the body of the reference's `learning()` coroutine
(`agents/learner_module/{ppo,impala,v_mpo,sac,sac_continuous}/learning.py`) for ONE
epoch, with the optimizers of `agents/learner.py:70,331-338,360-367` restated
explicitly (RMSprop / Adam / clip_grad_norm_).  Gradients come from torch autograd
on the restated forward math (the CUDA path hand-codes them).
"""
import math
from collections import OrderedDict

import torch

from . import models as M
from .losses import compute_gae, compute_v_trace, kldivergence, soft_update

# utils/parameters.json:1-31 (values restated)
DEFAULT_HYPER = dict(
    hidden_size=64, gamma=0.99, lmbda=0.95, eps_clip=0.1, policy_loss_coef=1.0,
    value_loss_coef=0.5, entropy_coef=0.00005, alpha=0.2, v_mpo_lagrange_multiplier_init=5,
    coef_eta=0.01, coef_alpha_upper=0.01, coef_alpha_below=0.005, K_epoch=1, lr=0.0001,
    max_grad_norm=40, reward_scale=0.1,
)


def smooth_l1(x, y):
    """F.smooth_l1_loss(x, y, beta=1) elementwise mean (ppo/learning.py:74)."""
    d = (x - y).abs()
    return torch.where(d < 1.0, 0.5 * d * d, d - 0.5).mean()


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ restated (ppo/learning.py:93-96):
    coef = min(1, max_norm / (||g||_2 + 1e-6)); returns the pre-clip norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


class RMSprop:
    """torch.optim.RMSprop(lr, alpha=0.99, eps, momentum=0, centered=False) restated
    (learner.py:70): v <- a v + (1-a) g^2 ; p <- p - lr g / (sqrt(v) + eps)."""

    def __init__(self, params, lr, alpha=0.99, eps=1e-5):
        self.p, self.lr, self.alpha, self.eps = list(params), lr, alpha, eps
        self.v = [torch.zeros_like(x) for x in self.p]

    def step(self, grads):
        with torch.no_grad():
            for p, v, g in zip(self.p, self.v, grads):
                v.mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)
                p.addcdiv_(g, v.sqrt().add_(self.eps), value=-self.lr)


class Adam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8) restated (learner.py:360-367)."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.p, self.lr, self.b1, self.b2, self.eps = list(params), lr, b1, b2, eps
        self.m = [torch.zeros_like(x) for x in self.p]
        self.v = [torch.zeros_like(x) for x in self.p]
        self.t = 0

    def step(self, grads):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        with torch.no_grad():
            for p, m, v, g in zip(self.p, self.m, self.v, grads):
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
                p.addcdiv_(m, denom, value=-self.lr / bc1)


def _flat(ts):
    return torch.cat([t.reshape(-1).float() for t in ts])


class OracleLearner:
    """One learner (model + optimizer state) for `algo` in
    {"PPO", "PPO-Continuous", "IMPALA", "V-MPO", "SAC", "SAC-Continuous"}.

    `params`: on-policy -> OrderedDict of the single net; SAC -> {"actor": ..., "critic": twin dict}.
    Tensors are cloned and owned by the learner (`self.params`).
    """

    def __init__(self, algo, params, hyper=None, fused_cell=False, separate_target=False):
        self.algo = algo
        self.h = dict(DEFAULT_HYPER)
        self.h.update(hyper or {})
        self.fused = fused_cell
        h = self.h
        if algo in ("SAC", "SAC-Continuous"):
            self.actor = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params["actor"].items())
            self.critic = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params["critic"].items())
            # learner.py:358: `target_critic = model.critic.to(device)` IS the critic (quirk A.6-8)
            self.target = (OrderedDict((k, v.detach().clone()) for k, v in self.critic.items())
                           if separate_target else self.critic)
            self.actor_opt = Adam(self.actor.values(), h["lr"])
            self.critic_opt = Adam(self.critic.values(), h["lr"])
            self.log_alpha = torch.tensor(math.log(h["alpha"]), dtype=torch.float64, requires_grad=True)
            self.alpha_opt = Adam([self.log_alpha], h["lr"])
            self.target_entropy = None  # = action_space, set at first update from shapes
        else:
            self.params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
            opt_params = list(self.params.values())
            if algo == "V-MPO":
                init = math.log(h["v_mpo_lagrange_multiplier_init"])
                self.log_eta = torch.tensor(init, dtype=torch.float64, requires_grad=True)
                self.log_alpha = torch.tensor(init, dtype=torch.float64, requires_grad=True)
                opt_params += [self.log_eta, self.log_alpha]
            self.opt = RMSprop(opt_params, h["lr"], eps=1e-5)

    # ------------------------------------------------------------------ on-policy
    def _fwd(self, b):
        cont = self.algo == "PPO-Continuous"
        return M.actor_critic_forward(self.params, b["obs"], b["hx"], b["cx"], b["act"],
                                      continuous=cont, fused_cell=self.fused)

    def _finish_on_policy(self, loss, out, extra=()):
        ps = list(self.params.values())
        grads = torch.autograd.grad(loss, ps + list(extra), allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, ps + list(extra))]
        mg = [g.clone() for g in grads[:len(ps)]]
        out["grad"] = _flat(mg)
        out["grad_norm"] = clip_grad_norm(mg, self.h["max_grad_norm"])   # model params only
        self.opt.step(mg + [g.clone() for g in grads[len(ps):]])
        out["extra_grads"] = [g.detach().clone() for g in grads[len(ps):]]
        out["loss"] = loss.detach()
        out["params"] = _flat([p.detach() for p in ps])
        return out

    def _ppo_update(self, b):
        h = self.h
        logits, logp, ent, value = self._fwd(b)
        with torch.no_grad():
            td = b["rew"][:, :-1] + h["gamma"] * (1 - b["is_fir"][:, 1:]) * value[:, 1:]
            delta = td - value[:, :-1]
            gae = compute_gae(delta, h["gamma"], h["lmbda"])
        ratio = torch.exp(logp[:, :-1] - b["log_prob"][:, :-1])
        s1 = ratio * gae
        s2 = ratio.clamp(1 - h["eps_clip"], 1 + h["eps_clip"]) * gae
        lp = -torch.min(s1, s2).mean()
        lv = smooth_l1(value[:, :-1], td)
        pe = ent[:, :-1].mean()
        loss = h["policy_loss_coef"] * lp + h["value_loss_coef"] * lv - h["entropy_coef"] * pe
        out = {"policy-loss": lp.detach(), "value-loss": lv.detach(), "policy-entropy": pe.detach(),
               "ratio": ratio.detach(), "gae": gae, "td_target": td, "value": value.detach(),
               "log_probs": logp.detach(), "entropy": ent.detach(), "logits": logits.detach()}
        return self._finish_on_policy(loss, out)

    def _impala_update(self, b):
        h = self.h
        logits, logp, ent, value = self._fwd(b)
        with torch.no_grad():
            rho, adv, vs = compute_v_trace(b["log_prob"], logp, b["is_fir"], b["rew"], value, h["gamma"])
        lp = -(logp[:, :-1] * adv).mean()
        lv = smooth_l1(value[:, :-1], vs[:, :-1])
        pe = ent[:, :-1].mean()
        loss = h["policy_loss_coef"] * lp + h["value_loss_coef"] * lv - h["entropy_coef"] * pe
        out = {"policy-loss": lp.detach(), "value-loss": lv.detach(), "policy-entropy": pe.detach(),
               "ratio": rho, "advantages": adv, "values_target": vs, "value": value.detach(),
               "log_probs": logp.detach(), "entropy": ent.detach(), "logits": logits.detach()}
        return self._finish_on_policy(loss, out)

    def _vmpo_update(self, b, coef_alpha):
        h = self.h
        logits, logp, _, value = self._fwd(b)
        with torch.no_grad():
            td = b["rew"][:, :-1] + h["gamma"] * (1 - b["is_fir"][:, 1:]) * value[:, 1:]
            delta = td - value[:, :-1]
            gae = compute_gae(delta, h["gamma"], h["lmbda"])
            top_gae, top_idx = torch.topk(gae, math.ceil(gae.size(0) / 2), 0)   # per-(t) over batch
            ratio = top_gae / (self.log_eta.exp().detach() + 1e-7)
        top_logp = logp[:, :-1].gather(0, top_idx)
        psi = torch.softmax(ratio.view(-1), 0)
        lp = -(psi.reshape(ratio.shape) * top_logp).sum()
        lv = smooth_l1(value[:, :-1], td)
        lt = self.log_eta.exp() * h["coef_eta"] + self.log_eta.exp() * (ratio.exp().mean().log())
        alpha = self.log_alpha.exp()
        kl = kldivergence(b["logits"][:, :-1], logits[:, :-1])
        la = (alpha * (coef_alpha - kl.detach()) + alpha.detach() * kl).mean()
        loss = h["policy_loss_coef"] * lp + h["value_loss_coef"] * lv + lt + la
        out = {"policy-loss": lp.detach(), "value-loss": lv.detach(), "loss-temperature": lt.detach(),
               "loss-alpha": la.detach(), "gae": gae, "td_target": td, "value": value.detach(),
               "kl": kl.detach(), "psi": psi.detach().reshape(ratio.shape), "top_idx": top_idx}
        out = self._finish_on_policy(loss, out, extra=(self.log_eta, self.log_alpha))
        out["log_eta"], out["log_alpha"] = self.log_eta.detach().clone(), self.log_alpha.detach().clone()
        return out

    # ------------------------------------------------------------------------ SAC
    def _twin(self, params, obs, hx, cx, act=None):
        q = []
        for name in ("q1.", "q2."):
            sp = M.sub(params, name)
            if act is None:
                q.append(M.sac_critic_forward(sp, obs, hx, cx, self.fused))
            else:
                q.append(M.sac_critic_continuous_forward(sp, obs, act, hx, cx, self.fused))
        return q

    def _sac_update(self, b, noise):
        h = self.h
        cont = self.algo == "SAC-Continuous"
        obs, hx, cx = b["obs"], b["hx"], b["cx"]
        A = b["act"].shape[-1] if cont else self.actor["logits.weight"].shape[0]
        target_entropy = float(A)                               # learner.py:363-365
        alpha32 = lambda: self.log_alpha.exp().detach().float()

        # -- actor step (sac/learning.py:36-62, sac_continuous/learning.py:36-55)
        if cont:
            a_pol, logp_pol = M.sac_actor_continuous_forward(self.actor, obs, hx, cx, noise[0], self.fused)
            q1, q2 = self._twin(self.critic, obs, hx, cx, a_pol)
            lp = (alpha32() * logp_pol - torch.min(q1, q2))[:, :-1].mean()
        else:
            probs, logp_pol = M.sac_actor_forward(self.actor, obs, hx, cx, self.fused)
            q1, q2 = self._twin(self.critic, obs, hx, cx)
            lp = (probs * (alpha32() * logp_pol - torch.min(q1, q2)))[:, :-1].sum(-1).mean()
        ag = [g.clone() for g in torch.autograd.grad(lp, list(self.actor.values()))]
        out = {"policy-loss": lp.detach(), "actor_grad": _flat(ag)}
        out["actor_grad_norm"] = clip_grad_norm(ag, h["max_grad_norm"])
        self.actor_opt.step(ag)

        # -- temperature step (sac/learning.py:64-74)
        if cont:
            ent_term = logp_pol[:, :-1].detach()
        else:
            ent_term = (probs * logp_pol)[:, :-1].sum(-1).detach()
        la = (self.log_alpha.exp() * (ent_term + target_entropy)).mean()
        (g_la,) = torch.autograd.grad(la, [self.log_alpha])
        self.alpha_opt.step([g_la])
        out["loss_alpha"] = la.detach()
        out["log_alpha_grad"] = g_la.detach().clone()

        # -- soft-Q target with the UPDATED actor and alpha (sac/learning.py:76-102)
        with torch.no_grad():
            if cont:
                a_cri, logp_cri = M.sac_actor_continuous_forward(self.actor, obs, hx, cx, noise[1], self.fused)
                tq1, tq2 = self._twin(self.target, obs, hx, cx, a_cri)
                soft = torch.min(tq1, tq2) - alpha32() * logp_cri
            else:
                probs_c, logp_cri = M.sac_actor_forward(self.actor, obs, hx, cx, self.fused)
                tq1, tq2 = self._twin(self.target, obs, hx, cx)
                soft = probs_c * (torch.min(tq1, tq2) - alpha32() * logp_cri)
            td = b["rew"][:, :-1] + (1 - b["is_fir"][:, 1:]) * h["gamma"] * soft[:, 1:].sum(-1, keepdim=True)

        # -- critic step (sac/learning.py:104-120)
        if cont:
            s1, s2 = self._twin(self.critic, obs, hx, cx, b["act"])
        else:
            s1, s2 = self._twin(self.critic, obs, hx, cx)
            s1 = s1.gather(-1, b["act"].long())
            s2 = s2.gather(-1, b["act"].long())
        lv = smooth_l1(s1[:, :-1], td) + smooth_l1(s2[:, :-1], td)
        cg = [g.clone() for g in torch.autograd.grad(lv, list(self.critic.values()))]
        out["critic_grad"] = _flat(cg)
        out["critic_grad_norm"] = clip_grad_norm(cg, h["max_grad_norm"])
        self.critic_opt.step(cg)
        soft_update(self.critic, self.target)                   # aliasing: theta <- (1-tau) theta + tau theta

        out.update({"value-loss": lv.detach(), "td_target": td, "alpha": self.log_alpha.exp().detach(),
                    "loss": (h["policy_loss_coef"] * lp + h["value_loss_coef"] * lv).detach(),
                    "actor_params": _flat([p.detach() for p in self.actor.values()]),
                    "critic_params": _flat([p.detach() for p in self.critic.values()]),
                    "log_alpha": self.log_alpha.detach().clone()})
        return out

    # --------------------------------------------------------------------- public
    def update(self, batch, coef_alpha=None, noise=None):
        a = self.algo
        if a in ("PPO", "PPO-Continuous"):
            return self._ppo_update(batch)
        if a == "IMPALA":
            return self._impala_update(batch)
        if a == "V-MPO":
            return self._vmpo_update(batch, coef_alpha)
        if a in ("SAC", "SAC-Continuous"):
            return self._sac_update(batch, noise)
        raise ValueError(a)
