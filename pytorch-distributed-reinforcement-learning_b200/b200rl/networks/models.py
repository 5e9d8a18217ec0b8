"""Host-side mirror of the reference's `networks/models.py` for the learner path.

Same class names, constructor signature `(f, n_outputs, sequence_length, hidden_size)`,
sub-module names and therefore the same `state_dict()` keys / shapes / order
(the state dict is a wire + checkpoint contract: workers
`load_state_dict` it, `worker.py:66-70`, and `main.py:128-146` resumes from it).

What differs: every parameter of a network is a VIEW into one flat fp32 buffer laid out
for the kernels (`ops.NetLayout`), `forward()` runs the sm_100a kernels through the C-ABI
(no autograd graph -- training goes through `b200rl.engine`, which hand-codes the
backward), and nothing loops over T in Python.
"""
import torch
import torch.nn as nn
from torch.distributions import Categorical, Normal

from .. import ops as _ops

H = _ops.H


class _FlatNet(nn.Module):
    """torso Linear+ReLU -> LSTMCell over T -> linear heads, parameters in one flat buffer."""

    continuous = False

    def __init__(self, f, n_outputs, sequence_length, hidden_size):
        super().__init__()
        if hidden_size != H:
            raise ValueError(f"b200rl kernels are built for hidden_size={H} (utils/parameters.json:10), got {hidden_size}")
        self.input_size, self.n_outputs = f, n_outputs
        self.sequence_length, self.hidden_size = sequence_length, hidden_size
        # same construction order as the reference => same weights under the same seed
        self.body = nn.Sequential(nn.Linear(f, hidden_size), nn.ReLU())
        self.after_torso()
        self._flat = None
        self.flatten_parameters()

    # --- layout ---------------------------------------------------------------
    def head_modules(self):
        raise NotImplementedError

    @property
    def torso_in(self):
        return self.input_size

    def layout(self):
        NH = sum(m.out_features for m in self.head_modules())
        return _ops.NetLayout(self.torso_in, NH)

    def _torso_blocks(self):
        """[(param, row0, col0)] blocks of the dense torso matrix Wt (H, Fin), and bias pieces."""
        lin = self.body[0]
        return [(lin.weight, 0, 0)], [(lin.bias, 0)]

    def flatten_parameters(self, into=None):
        """(Re)home all parameters into one flat buffer on their current device (`into`: a
        preallocated zeroed slice, used by the twin critic to keep q1|q2 contiguous)."""
        L = self.layout()
        dev = self.body[0].weight.device
        flat = torch.zeros(L.P, device=dev, dtype=torch.float32) if into is None else into
        Wt = L.view(flat, "Wt").view(H, L.Fin)
        bt = L.view(flat, "bt")
        wblocks, bblocks = self._torso_blocks()
        with torch.no_grad():
            for p, r0, c0 in wblocks:
                dst = Wt[r0:r0 + p.shape[0], c0:c0 + p.shape[1]]
                dst.copy_(p.data)
                p.data = dst
            for p, r0 in bblocks:
                dst = bt[r0:r0 + p.shape[0]]
                dst.copy_(p.data)
                p.data = dst
            for name, key in (("weight_ih", "W_ih"), ("weight_hh", "W_hh"), ("bias_ih", "b_ih"), ("bias_hh", "b_hh")):
                p = getattr(self.lstmcell, name)
                dst = L.view(flat, key).view(p.shape)
                dst.copy_(p.data)
                p.data = dst
            Wh = L.view(flat, "W_head").view(L.NH, H)
            bh = L.view(flat, "b_head")
            r = 0
            for m in self.head_modules():
                k = m.out_features
                dw, db = Wh[r:r + k], bh[r:r + k]
                dw.copy_(m.weight.data)
                db.copy_(m.bias.data)
                m.weight.data, m.bias.data = dw, db
                r += k
        self._flat = flat
        self._layout = L
        self._pinfo = {n: (q.data.storage_offset() - flat.storage_offset(), tuple(q.shape), tuple(q.stride()))
                       for n, q in self.named_parameters()}
        return flat

    def views_of(self, flat_like):
        """state-dict-ordered {key: view} into any flat tensor laid out like the parameters
        (gradients, optimizer state)."""
        from collections import OrderedDict
        self.flat_parameters()
        return OrderedDict((n, torch.as_strided(flat_like, shp, std, flat_like.storage_offset() + off))
                           for n, (off, shp, std) in self._pinfo.items())

    def export_state_dict(self):
        """Compact CPU copy of `state_dict()` (one D2H of the flat buffer, then per-key clones)
        -- what `pub_model` / checkpoints should ship instead of views of the flat buffer."""
        host = self.flat_parameters().detach().cpu()
        from collections import OrderedDict
        return OrderedDict((n, v.clone()) for n, v in self.views_of(host).items())

    def state_dict_from_flat(self, host_flat):
        """`state_dict()`-shaped clones cut out of a host copy of the flat buffer (see engine.snapshot_async)."""
        from collections import OrderedDict
        return OrderedDict((n, v.clone()) for n, v in self.views_of(host_flat).items())

    def flat_parameters(self):
        """The flat buffer (re-flattened first if something, e.g. unpickling, re-homed the views)."""
        L = self._layout
        w = self.lstmcell.weight_ih
        if self._flat is None or w.data_ptr() != self._flat.data_ptr() + 4 * L.off["W_ih"][0]:
            self.flatten_parameters()
        return self._flat

    def grad_mask(self):
        """None, or a flat 0/1 mask of entries that are real parameters (block-sparse torso)."""
        return None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._flat = None
        self.flatten_parameters()
        return self

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_flat"] = None
        return st

    # --- kernels ----------------------------------------------------------------
    def run_lstm(self, u, lstm_hxs, ops=None):
        """-> (head_out (B,T,NH), workspace) via the network-forward op (one fused kernel, or lstm_fwd + head_fwd)."""
        ops = ops or _ops.default()
        flat, L = self.flat_parameters(), self._layout
        hx, cx = lstm_hxs
        B, T = u.shape[0], u.shape[1]
        dev = u.device
        ws = dict(gates=torch.empty(B, T, H, 4, device=dev), c=torch.empty(B, T, H, device=dev),
                  h=torch.empty(B, T, H, device=dev))
        u = u.contiguous().float()
        head_out = torch.empty(B, T, L.NH, device=dev)
        ops.net_fwd(L, flat, u, hx.contiguous().float(), cx.contiguous().float(), ws, head_out)
        return head_out, ws

    # worker-side single-step inference (reference `act`, models.py:37-56): plain torch on
    # the shared parameter views; env-bound and outside the learner path.
    def _cell(self, obs, lstm_hxs):
        x = self.body(obs)
        return self.lstmcell(x, lstm_hxs)


class MlpLSTMBase(_FlatNet):
    """Discrete actor-critic (reference networks/models.py:8-100)."""

    def after_torso(self):
        self.lstmcell = nn.LSTMCell(self.hidden_size, self.hidden_size)
        self.value = nn.Linear(self.hidden_size, 1)
        self.logits = nn.Linear(self.hidden_size, self.n_outputs)
        self.CT = Categorical

    def head_modules(self):
        return [self.value, self.logits]

    def get_dist(self, x):
        return self.CT(torch.softmax(self.logits(x), dim=-1))

    def act(self, obs, lstm_hxs):
        with torch.no_grad():
            hx, cx = self._cell(obs, lstm_hxs)
            dist = self.get_dist(hx)
            action = dist.sample().detach()
            logits = dist.logits.detach() if hasattr(dist, "logits") else torch.zeros(action.shape)
        return action, logits, dist.log_prob(action).detach(), (hx.detach(), cx.detach())

    def act_batch(self, obs, lstm_hxs, ops=None):
        """Batched worker-side inference: `act` (reference networks/models.py:37-56, called once per
        environment step by every worker process, agents/worker.py:105) for N environments at once on the device -- ONE
        launch of the fused forward kernel with T = 1 (torso, LSTM cell and heads), then the sampling on the (N, A) head
        outputs.  obs (N, f), lstm_hxs = (hx, cx) each (N, H).  Returns what `act` returns, batched: action (N, 1) /
        (N, act_dim), logits (N, A) normalised log-probs (zeros for the continuous net), log_prob (N, 1) and the new (hx, cx)."""
        ops = ops or _ops.default()
        N = obs.shape[0]
        hx, cx = lstm_hxs
        head_out, ws = self.run_lstm(obs.reshape(N, 1, -1).float(), (hx, cx), ops)
        z = head_out[:, 0, 1:]
        with torch.no_grad():
            if self.continuous:
                dist = Normal(torch.tanh(z[:, :1]), torch.nn.functional.softplus(z[:, 1:2]))
                action = dist.sample()
                logits = torch.zeros_like(action)
                logp = dist.log_prob(action)
            else:
                dist = self.CT(torch.softmax(z, dim=-1))
                a = dist.sample()
                action = a.unsqueeze(-1).float()
                logits = dist.logits
                logp = dist.log_prob(a).unsqueeze(-1)
        return action, logits, logp, (ws["h"][:, 0].clone(), ws["c"][:, 0].clone())

    def forward(self, obs, lstm_hxs, behaviour_acts, ops=None):
        """-> (logits (B,T,A) normalised log-probs, log_probs (B,T,1), entropy (B,T,1), value (B,T,1))."""
        ops = ops or _ops.default()
        B, T = obs.shape[0], obs.shape[1]
        head_out, _ = self.run_lstm(obs.reshape(B, T, -1), lstm_hxs, ops)
        dev = obs.device
        A = self.n_outputs
        z = torch.zeros(B, T, 1, device=dev)
        batch = dict(act=behaviour_acts.contiguous().float(), rew=z, is_fir=z, log_prob=z)
        outs = dict(log_probs=torch.empty(B, T, 1, device=dev), entropy=torch.empty(B, T, 1, device=dev))
        if not self.continuous:
            outs["logits"] = torch.empty(B, T, A, device=dev)
        scal = torch.zeros(8, dtype=torch.float64, device=dev)
        hyper = dict(gamma=0.0, lmbda=0.0, eps_clip=0.0, policy_loss_coef=0.0, value_loss_coef=0.0, entropy_coef=0.0)
        ops.ppo_head(head_out, batch, torch.empty_like(head_out), scal, hyper, A, self.continuous, outs)
        logits = outs["logits"] if not self.continuous else torch.zeros_like(behaviour_acts)
        return logits, outs["log_probs"], outs["entropy"], head_out[..., 0:1].contiguous()


class MlpLSTMContinuous(MlpLSTMBase):
    """Gaussian actor-critic, mu = tanh(.), std = softplus(.) (reference models.py:103-118)."""

    continuous = True

    def after_torso(self):
        super().after_torso()
        self.logits_std = nn.Linear(self.hidden_size, self.n_outputs)
        self.NM = Normal

    def head_modules(self):
        return [self.value, self.logits, self.logits_std]

    def get_dist(self, x):
        return self.NM(torch.tanh(self.logits(x)), nn.functional.softplus(self.logits_std(x)))


class MlpLSTMSingle(nn.Module):
    """Actor and critic share one torso (reference models.py:345-351)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.actor = MlpLSTMBase(*args, **kwargs)
        self.critic = self.actor


class MlpLSTMSingleContinuous(nn.Module):
    """(reference models.py:354-360)"""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.actor = MlpLSTMContinuous(*args, **kwargs)
        self.critic = self.actor


# ================================================================================ SAC networks
class MlpLSTMActor(_FlatNet):
    """Discrete SAC actor: softmax policy head only (reference models.py:121-159)."""

    def after_torso(self):
        self.lstmcell = nn.LSTMCell(self.hidden_size, self.hidden_size)
        self.logits = nn.Linear(self.hidden_size, self.n_outputs)
        self.CT = Categorical

    def head_modules(self):
        return [self.logits]

    def get_dist(self, x):
        return self.CT(torch.softmax(self.logits(x), dim=-1))

    def act(self, obs, lstm_hxs):
        with torch.no_grad():
            hx, cx = self._cell(obs, lstm_hxs)
            dist = self.get_dist(hx)
            action = dist.sample().detach()
        return action, dist.logits.detach(), dist.log_prob(action).detach(), (hx.detach(), cx.detach())

    def forward(self, obs, lstm_hxs, ops=None):
        """-> (probs (B,T,A), log(probs + 1e-8*[probs==0]) (B,T,A))"""
        B, T = obs.shape[0], obs.shape[1]
        z, _ = self.run_lstm(obs.reshape(B, T, -1), lstm_hxs, ops)
        probs = torch.softmax(z, -1)
        return probs, torch.log(probs + (probs == 0.0).float() * 1e-8)


class MlpLSTMActorContinuous(_FlatNet):
    """tanh-Gaussian SAC actor: mu and clamped log-std heads (reference models.py:162-231)."""

    continuous = True

    def after_torso(self):
        self.lstmcell = nn.LSTMCell(self.hidden_size, self.hidden_size)
        self.logits = nn.Linear(self.hidden_size, self.n_outputs)
        self.logits_log_std = nn.Linear(self.hidden_size, self.n_outputs)
        self.NM = Normal

    def head_modules(self):
        return [self.logits, self.logits_log_std]

    def get_dist(self, x):
        return self.NM(self.logits(x), torch.clamp(self.logits_log_std(x), min=-20, max=2).exp())

    def action_log_prob(self, x):
        dist = self.get_dist(x)
        x_t = dist.rsample()
        action = torch.tanh(x_t)
        log_prob = dist.log_prob(x_t) - torch.log(1 - action.pow(2) + 1e-7)
        return dist, action, log_prob

    def act(self, obs, lstm_hxs):
        with torch.no_grad():
            hx, cx = self._cell(obs, lstm_hxs)
            _, action, log_prob = self.action_log_prob(hx)
        return action.detach(), torch.zeros(action.shape), log_prob.detach(), (hx.detach(), cx.detach())


class MlpLSTMCritic(_FlatNet):
    """Discrete Q network, Q(s, .) (reference models.py:234-270).  The reference constructs the
    torso twice (`MlpLSTMBase.__init__` then the override, models.py:239-242); the same RNG draws
    are consumed here so seeded weights agree."""

    def __init__(self, f, n_outputs, sequence_length, hidden_size):
        super().__init__(f, n_outputs, sequence_length, hidden_size)
        fresh = nn.Linear(f, hidden_size)
        with torch.no_grad():
            self.body[0].weight.copy_(fresh.weight)
            self.body[0].bias.copy_(fresh.bias)

    def after_torso(self):
        self.lstmcell = nn.LSTMCell(self.hidden_size, self.hidden_size)
        self.q_value = nn.Linear(self.hidden_size, self.n_outputs)

    def head_modules(self):
        return [self.q_value]

    def forward(self, obs, lstm_hxs, ops=None):
        B, T = obs.shape[0], obs.shape[1]
        q, _ = self.run_lstm(obs.reshape(B, T, -1), lstm_hxs, ops)
        return q


class MlpLSTMCriticContinuous(_FlatNet):
    """Q(s, a): obs -> H/2 and act -> H/2 encoders, concatenated, LSTM, scalar head
    (reference models.py:273-322).  Both encoders live in ONE dense torso matrix Wt (H, f + act_dim)
    with zero off-diagonal blocks, so the shared LSTM kernels apply; `grad_mask()` keeps the zero
    blocks out of the optimizer.  `act_dim` generalises the reference's hard-coded 1 (models.py:290,
    SURVEY.md 0.1); `n_outputs` is the action dimension as in the reference constructor."""

    def __init__(self, f, n_outputs, sequence_length, hidden_size, act_dim=1):
        self.act_dim = act_dim
        super().__init__(f, n_outputs, sequence_length, hidden_size)
        fresh = nn.Linear(f, hidden_size // 2)                      # models.py:278-284 (second construction)
        with torch.no_grad():
            self.body[0].weight.copy_(fresh.weight)
            self.body[0].bias.copy_(fresh.bias)

    def after_torso(self):
        # Base.__init__ built body as Linear(f, H) (RNG draw kept for parity); replace by the H/2 encoder
        with torch.random.fork_rng():                  # placeholder values; the real draw happens in __init__, in reference order
            self.body = nn.Sequential(nn.Linear(self.input_size, self.hidden_size // 2), nn.ReLU())
        self.lstmcell = nn.LSTMCell(self.hidden_size, self.hidden_size)
        self.encode_action = nn.Sequential(nn.Linear(self.act_dim, self.hidden_size // 2), nn.ReLU())
        self.q_value = nn.Linear(self.hidden_size, 1)

    @property
    def torso_in(self):
        return self.input_size + self.act_dim

    def head_modules(self):
        return [self.q_value]

    def _torso_blocks(self):
        half = self.hidden_size // 2
        return ([(self.body[0].weight, 0, 0), (self.encode_action[0].weight, half, self.input_size)],
                [(self.body[0].bias, 0), (self.encode_action[0].bias, half)])

    def grad_mask(self):
        L = self._layout
        m = torch.ones(L.P, device=self.flat_parameters().device)
        Wt = L.view(m, "Wt").view(H, L.Fin)
        half = self.hidden_size // 2
        Wt[:half, self.input_size:] = 0
        Wt[half:, :self.input_size] = 0
        m[L.n:] = 0
        return m

    def forward(self, obs, act, lstm_hxs, ops=None):
        B, T = obs.shape[0], obs.shape[1]
        u = torch.cat([obs.reshape(B, T, -1), act.reshape(B, T, -1)], -1)
        q, _ = self.run_lstm(u, lstm_hxs, ops)
        return q


class _Twin(nn.Module):
    """q1 and q2 side by side in ONE flat buffer [q1 | q2], so `clip_grad_norm_(critic.parameters())`
    (a joint norm over both networks, sac/learning.py:117-119) and Adam are one launch."""

    def forward(self, *args, **kwargs):
        return self.q1(*args, **kwargs), self.q2(*args, **kwargs)

    def flat_joint(self):
        P = self.q1._layout.P
        f1, f2 = self.q1.flat_parameters(), self.q2.flat_parameters()
        joint = getattr(self, "_joint", None)
        if joint is None or f1.data_ptr() != joint.data_ptr() or f2.data_ptr() != joint.data_ptr() + 4 * P:
            joint = torch.zeros(2 * P, device=f1.device, dtype=torch.float32)
            self.q1.flatten_parameters(into=joint[:P])
            self.q2.flatten_parameters(into=joint[P:])
            self._joint = joint
        return joint

    def flat_buffers(self):
        return [self.flat_joint()]

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_joint", None)
        return st


class MlpLSTMDoubleCritic(_Twin):
    """(reference models.py:335-342)"""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.q1 = MlpLSTMCritic(*args, **kwargs)
        self.q2 = MlpLSTMCritic(*args, **kwargs)


class MlpLSTMDoubleCriticContinuous(_Twin):
    """(reference models.py:325-332)"""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.q1 = MlpLSTMCriticContinuous(*args, **kwargs)
        self.q2 = MlpLSTMCriticContinuous(*args, **kwargs)


class MlpLSTMSeperate(nn.Module):
    """Separate actor and twin-critic torsos (reference models.py:363-369; the spelling is the reference's)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.actor = MlpLSTMActor(*args, **kwargs)
        self.critic = MlpLSTMDoubleCritic(*args, **kwargs)


class MlpLSTMSeperateContinuous(nn.Module):
    """(reference models.py:372-378).  `act_dim` (default 1 = the reference's domain) widens the action."""

    def __init__(self, f, n_outputs, sequence_length, hidden_size, act_dim=1):
        super().__init__()
        self.actor = MlpLSTMActorContinuous(f, n_outputs, sequence_length, hidden_size)
        self.critic = MlpLSTMDoubleCriticContinuous(f, n_outputs, sequence_length, hidden_size, act_dim=act_dim)
