"""Learner classes -- host-side mirror of the reference's `agents/learner.py`.

Kept from the reference: class names (`LearnerBase`,
`LearnerSinglePPO`, `LearnerSingleIMPALA`, `LearnerSingleVMPO`, `LearnerSeperate` (sic)),
the constructor signature (`learner.py:40-52`), `sample_batch_from_sh_memory`,
`put_batch_to_batch_q`, `learning_chain_<algo>`, `pub_model`, `log_loss_tensorboard`,
the shared-memory reader interface and the module-level `timer`.

Replaced: `self.optimizer` (torch RMSprop) and the autograd step become `self.engine`
(`b200rl.engine.OnPolicyEngine`: native kernels + fused clip/RMSprop on the flat
parameter vector); `self.to_gpu` ships only the fields the update reads, from pinned
memory.  ZeroMQ / tensorboardX are imported lazily and only if a socket / log dir is asked
for, so the learner update can be driven without them (the benchmark isolates the learner
from workers, as BASELINE.json's north_star asks).
"""
import asyncio
import os
import pickle

import numpy as np
import torch

from ..engine import OnPolicyEngine, SACEngine, VMPOEngine
from ..utils.utils import ExecutionTimer, HostStager, Params, Protocol, hyper_from_args, to_torch
from .storage_module.shared_batch import SMInterFace
from . import ppo_awrapper, v_mpo_awrapper, impala_awrapper, sac_awrapper, sac_continuous_awrapper

timer = ExecutionTimer(num_transition=Params.seq_len * Params.batch_size * 1)


def encode(protocol, data):
    """Wire frame of the reference (`utils/utils.py:244-245`): pickle + blosc2(clevel=1)."""
    import blosc2
    return pickle.dumps(protocol), blosc2.compress(pickle.dumps(data), clevel=1)


class LearnerBase(SMInterFace):
    def __init__(self, args, mutex, model, shm_ref, stop_event, learner_ip, learner_port, obs_shape,
                 shared_stat_array=None, heartbeat=None, process_group=None, ops=None):
        super().__init__(shm_ref=shm_ref)
        self.args = args
        self.mutex = mutex
        self.stop_event = stop_event
        self.obs_shape = obs_shape
        self.np_shared_stat_array = None
        if shared_stat_array is not None:
            self.np_shared_stat_array = np.frombuffer(buffer=shared_stat_array.get_obj(), dtype=np.float32, count=-1)
        self.heartbeat = heartbeat
        self.device = torch.device(self.args.device)
        self.model = model.to(self.device)
        self.process_group = process_group
        self._ops = ops                      # test hook (host-emulated library); None -> product library
        self.to_gpu = HostStager(self.device)
        self.pub_socket = None
        if learner_ip is not None:
            self.zeromq_set(learner_ip, learner_port)
        self.zero_copy = False               # shared arrays page-locked: batches are DMA'd straight out of the storage slots
        if shm_ref is not None:
            self.get_shared_memory_interface()
            if self.device.type == "cuda" and os.environ.get("B200RL_ZERO_COPY", "1") != "0":
                self.zero_copy = self.register_pinned()
        self.writer = None
        if getattr(args, "result_dir", None) is not None:
            from tensorboardX import SummaryWriter
            self.writer = SummaryWriter(log_dir=args.result_dir)
        self.idx = 0

    def __del__(self):
        if getattr(self, "pub_socket", None) is not None:
            self.pub_socket.close()
        if getattr(self, "zero_copy", False):
            try:
                self.unregister_pinned()
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    # ------------------------------------------------------------------ transport
    def zeromq_set(self, learner_ip, learner_port):
        import zmq
        context = zmq.Context()
        self.pub_socket = context.socket(zmq.PUB)
        self.pub_socket.bind(f"tcp://{learner_ip}:{int(learner_port) + 1}")

    def pub_model(self, model_state_dict):
        """Publish the model to the workers (reference learner.py:92-93: pickle + blosc2 + ZeroMQ PUB, every update).

        `args.publish_interval` = k > 1 publishes every k-th update only; `args.publish_async` moves pickle + compress + send
        to a single worker thread fed through a one-slot mailbox that always holds the NEWEST state dict (workers only ever
        want the latest model; an older unsent snapshot is dropped), so the learner loop never waits for the wire format
        (SURVEY.md 8f-3).  Defaults (k = 1, synchronous) are the reference's behaviour."""
        if self.pub_socket is None:
            return
        k = int(getattr(self.args, "publish_interval", 1) or 1)
        self._pub_count = getattr(self, "_pub_count", 0) + 1
        if k > 1 and (self._pub_count - 1) % k != 0:
            return
        if not getattr(self.args, "publish_async", False):
            self.pub_socket.send_multipart([*encode(Protocol.Model, model_state_dict)])
            return
        if getattr(self, "_pub_thread", None) is None:
            import threading
            self._pub_box, self._pub_cv, self._pub_stop = [None], threading.Condition(), False
            self.pub_dropped = 0

            def worker():
                while True:
                    with self._pub_cv:
                        while self._pub_box[0] is None and not self._pub_stop:
                            self._pub_cv.wait()
                        if self._pub_box[0] is None:
                            return
                        sd, self._pub_box[0] = self._pub_box[0], None
                    self.pub_socket.send_multipart([*encode(Protocol.Model, sd)])
            self._pub_thread = threading.Thread(target=worker, daemon=True)
            self._pub_thread.start()
        with self._pub_cv:
            if self._pub_box[0] is not None:
                self.pub_dropped += 1
            # the state dict views a pinned snapshot buffer that is reused two updates later: the mailbox owns a copy
            self._pub_box[0] = {k_: v.clone() for k_, v in model_state_dict.items()}
            self._pub_cv.notify()

    def pub_flush(self):
        """Wait until the asynchronous publisher has sent what is in its mailbox and stop it."""
        t = getattr(self, "_pub_thread", None)
        if t is not None:
            with self._pub_cv:
                self._pub_stop = True
                self._pub_cv.notify()
            t.join()
            self._pub_thread = None

    # -------------------------------------------------------------------- logging
    def log_loss_tensorboard(self, timer, loss, detached_losses):
        w = self.writer
        if w is None:
            return
        w.add_scalar("total-loss", float(loss.item()), self.idx)
        for key, tag in (("value-loss", "original-value-loss"), ("policy-loss", "original-policy-loss"),
                         ("policy-entropy", "original-policy-entropy")):
            if key in detached_losses:
                w.add_scalar(tag, float(detached_losses[key]), self.idx)
        if "ratio" in detached_losses:
            r = detached_losses["ratio"]
            w.add_scalar("min-ratio", float(r.min()), self.idx)
            w.add_scalar("max-ratio", float(r.max()), self.idx)
            w.add_scalar("avg-ratio", float(r.mean()), self.idx)
        for key in ("loss-temperature", "loss_alpha"):
            if key in detached_losses:
                w.add_scalar(key, float(detached_losses[key].mean()), self.idx)
        if "alpha" in detached_losses:
            w.add_scalar("alpha", float(detached_losses["alpha"]), self.idx)
        st = self.np_shared_stat_array
        if st is not None and st.size == 3 and bool(st[2]):
            w.add_scalar("50-game-mean-stat-of-epi-rew", st[1], st[0])
            st[2] = 0
        if isinstance(timer, ExecutionTimer):
            for k, v in timer.timer_dict.items():
                w.add_scalar(f"{k}-elapsed-mean-sec", sum(v) / (len(v) + 1e-6), self.idx)
            for k, v in timer.throughput_dict.items():
                w.add_scalar(f"{k}-transition-per-secs", sum(v) / (len(v) + 1e-6), self.idx)

    # ------------------------------------------------------ shared memory -> batch
    def sample_batch_from_sh_memory(self, sampling_method, copy=True):
        """The 8-tuple (obs, act, rew, logits, log_prob, is_fir, hx, cx) of host tensors (reference learner.py:168-233).
        `hx`/`cx` are returned as (B, H): only `[:, 0]` of the stored (B, T, H) is ever read by the update.  `copy=True` is
        the reference's defensive copy (fresh tensors); `copy=False` (on-policy only) returns VIEWS of the shared arrays for
        the zero-copy path -- the caller must not reset the slot counter before it has shipped them."""
        assert sampling_method in ("on-policy", "off-policy")
        sq, bn, sha = self.args.seq_len, self.args.batch_size, self.obs_shape
        hs, ac, buf = self.args.hidden_size, self.args.action_space, self.args.buffer_size
        stored = int(self.sh_obs_batch.shape[0] / (sq * sha[0]))
        if sampling_method == "off-policy":
            assert buf == stored
            idx, num = np.random.randint(0, buf, size=bn), buf
        else:
            assert bn == stored
            idx, num = slice(None), bn
        widths = (("obs", tuple(sha)), ("act", (1,)), ("rew", (1,)), ("logits", (ac,)), ("log_prob", (1,)),
                  ("is_fir", (1,)), ("hx", (hs,)), ("cx", (hs,)))
        out = []
        for name, w in widths:
            view = getattr(self, f"sh_{name}_batch").reshape((num, sq, *w))[idx]
            if name in ("hx", "cx"):
                view = view[:, 0]
            if copy or sampling_method == "off-policy":
                out.append(to_torch(np.array(view, dtype=np.float32, copy=True)))
            else:
                out.append(torch.from_numpy(view))
        return tuple(out)

    def is_sh_ready(self):
        return self.sh_data_num.value >= self.args.batch_size

    async def put_batch_to_batch_q(self):
        while not self.stop_event.is_set():
            if self.is_sh_ready():
                if self.zero_copy:
                    # DMA straight out of the page-locked slots on the copy stream; the slots go back to the storage process
                    # (counter reset) as soon as the copy engine has read them -- the queue carries device-resident batches
                    views = self.sample_batch_from_sh_memory(sampling_method="on-policy", copy=False)
                    shipped = self.to_gpu.ship(*views)
                    while shipped is None and not self.stop_event.is_set():      # ring full: the learner is behind
                        await asyncio.sleep(0.0005)
                        shipped = self.to_gpu.ship(*views)
                    if shipped is None:
                        break
                    dev, ev = shipped
                    while not ev.query():
                        await asyncio.sleep(0)
                    await self.batch_queue.put(dev)
                else:
                    await self.batch_queue.put(self.sample_batch_from_sh_memory(sampling_method="on-policy"))
                self.reset_data_num()
                await asyncio.sleep(0)         # a batch was just handed over: look again at once (the reference's fixed 1 ms nap
                continue                       # per batch is as long as a whole update here)
            await asyncio.sleep(0.001)

    # ----------------------------------------------------------------- coroutines
    @ppo_awrapper(timer=timer)
    def learning_ppo(self): ...

    @v_mpo_awrapper(timer=timer)
    def learning_v_mpo(self): ...

    @impala_awrapper(timer=timer)
    def learning_impala(self): ...

    @sac_awrapper(timer=timer)
    def learning_sac(self): ...

    @sac_continuous_awrapper(timer=timer)
    def learning_sac_continuous(self): ...

    async def _chain(self, queue_name, learn, feed):
        setattr(self, queue_name, asyncio.Queue(1024))
        await asyncio.gather(asyncio.create_task(learn()), asyncio.create_task(feed()))

    async def learning_chain_ppo(self):
        await self._chain("batch_queue", self.learning_ppo, self.put_batch_to_batch_q)

    async def learning_chain_v_mpo(self):
        await self._chain("batch_queue", self.learning_v_mpo, self.put_batch_to_batch_q)

    async def learning_chain_impala(self):
        await self._chain("batch_queue", self.learning_impala, self.put_batch_to_batch_q)

    async def learning_chain_sac(self):
        await self._chain("batch_buffer", self.learning_sac, self.put_batch_to_buffer_q)

    async def learning_chain_sac_continuous(self):
        await self._chain("batch_buffer", self.learning_sac_continuous, self.put_batch_to_buffer_q)


class LearnerSinglePPO(LearnerBase):
    """PPO / PPO-Continuous (reference learner.py:308-314)."""

    engine_algo = None      # None -> args.algo

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.model.actor is self.model.critic
        self.actor = self.model.actor
        self.critic = self.model.critic
        algo = self.engine_algo or self.args.algo
        self.engine = OnPolicyEngine(self.actor, algo, hyper_from_args(self.args), ops=self._ops,
                                     process_group=self.process_group)
        self.to_gpu.skip = {"logits"}          # behaviour logits are read by V-MPO only


class LearnerSingleIMPALA(LearnerSinglePPO):
    """(reference learner.py:317)"""
    engine_algo = "IMPALA"


class LearnerSingleVMPO(LearnerBase):
    """V-MPO (reference learner.py:320-348): the dual variables `log_eta`, `log_alpha` (fp64,
    initialised to log(v_mpo_lagrange_multiplier_init)) and their RMSprop groups live in the
    engine on the device; `get_coef_alpha()` draws eps_alpha log-uniformly in
    [coef_alpha_below, coef_alpha_upper] exactly as the reference does."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.model.actor is self.model.critic
        self.actor = self.model.actor
        self.critic = self.model.critic
        self.engine = VMPOEngine(self.actor, hyper_from_args(self.args), ops=self._ops, process_group=self.process_group)
        self.to_gpu.skip = {"log_prob"}          # V-MPO reads the behaviour logits, not log_prob

    @property
    def log_eta(self):
        return self.engine.duals[0]

    @property
    def log_alpha(self):
        return self.engine.duals[1]

    def get_coef_alpha(self):
        from torch.distributions import Uniform
        return Uniform(torch.tensor(self.args.coef_alpha_below).log(), torch.tensor(self.args.coef_alpha_upper).log()).sample().exp()


class LearnerSeperate(LearnerBase):
    """SAC / SAC-Continuous learner (reference learner.py:351-400; the spelling is the reference's):
    separate actor and twin critic, three Adam optimizers (fused kernels on the flat buffers),
    `log_alpha` = log(args.alpha) in fp64 on the device, `target_entropy = action_space`, and a
    replay-style readiness latch: once the shared buffer has filled once it stays ready."""

    def __init__(self, *args, separate_target=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.actor = self.model.actor
        self.critic = self.model.critic
        self.target_critic = self.critic                      # aliasing, as in the reference (learner.py:358)
        cont = bool(getattr(self.actor, "continuous", False))
        self.engine = SACEngine(self.model, hyper_from_args(self.args), cont, ops=self._ops, separate_target=separate_target,
                                process_group=self.process_group)
        self.target_entropy = self.args.action_space
        self.to_gpu.skip = {"logits", "log_prob"}
        self._latched = False

    @property
    def log_alpha(self):
        return self.engine.log_alpha

    def is_sh_ready(self):
        if self.sh_data_num.value >= self.args.buffer_size:
            self._latched = True
            self.reset_data_num()
        return self._latched

    async def put_batch_to_buffer_q(self):
        while not self.stop_event.is_set():
            if self.is_sh_ready():
                await self.batch_buffer.put(self.sample_batch_from_sh_memory(sampling_method="off-policy"))
                await asyncio.sleep(0)
                continue
            await asyncio.sleep(0.001)
