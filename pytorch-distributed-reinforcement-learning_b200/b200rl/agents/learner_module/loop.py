"""The learner iteration skeleton shared by every algorithm.

The reference repeats this ~100-line loop in each `learning.py`
(`ppo/learning.py:13-126`, `impala/learning.py:13-114`, ...): wait for a batch, H2D,
`K_epoch` updates, publish the model, log every `loss_log_interval`, checkpoint every
`model_save_interval`, heartbeat, yield to the batching coroutine.  Here it exists once;
each algorithm module contributes only its `update(parent, device_batch)`.

Contract kept from the reference: timer block names
(`learner-throughput` with throughput, `learner-batching-time`, `learner-forward-time`
enclosing `learner-backward-time`), `parent.idx`, `parent.pub_model(state_dict)` every
iteration, `parent.log_loss_tensorboard(timer, loss, detached_losses)`, checkpoint file
name `<algo>_<idx>.pt` (fires at idx 0), `parent.heartbeat`, a yield to the batching coroutine per iteration (`asyncio.sleep(0)`;
the reference sleeps 1 ms, which at a 5 ms update would cost 20 % of the throughput).
Difference: no per-epoch `print(loss.item())` -- that was the reference's only device sync.

Two schedules.  Engines without `snapshot_async` (V-MPO, SAC) synchronise the stream once per iteration and then
publish / log, exactly in the reference's order.  Engines with it (PPO / IMPALA) run one iteration AHEAD of the host:
after enqueuing update i the loop enqueues stream-ordered device->pinned-host copies of the parameters and the loss
scalars of update i, starts the H2D of batch i+1, and only then finishes iteration i-1 on the host (`pub_model`, logging,
checkpoint) while the GPU is busy with update i.  The host-side work (0.6 ms per iteration at B=4096, T=128: sync
latency, state-dict export, loss read-back, kernel enqueue) no longer leaves the GPU idle; what the workers receive is
bit-identical, one update (1.2 ms) later -- but ONLY while the queue already holds the next batch: whenever the loop is
about to wait on an empty queue it first finishes the pending iteration, so a slow feeder never sees a model that is a
whole batch-collection interval stale.  `B200RL_PIPELINE=0` selects the synchronous schedule.
"""
import asyncio
import os
import time

import torch


class RatioStats:
    """Stands in for the (B,T-1,1) `ratio` tensor the reference ships to the host only to
    log its min / max / mean (ppo/learning.py:87, learner.py:111-120)."""

    def __init__(self, mn, mx, mean):
        self._mn, self._mx, self._mean = mn, mx, mean

    def min(self):
        return self._mn

    def max(self):
        return self._mx

    def mean(self):
        return self._mean


class HostScalar(float):
    """float that also answers `.item()` / `.mean()` like a 0-dim tensor."""

    def item(self):
        return float(self)

    def mean(self):
        return self


async def _next_batch(queue, stop_event, poll=0.05):
    """`await queue.get()`, but a learner whose feeder has stopped leaves its loop instead of waiting forever (the
    reference blocks in `queue.get()` until the process is terminated): returns None once `stop_event` is set."""
    while True:
        try:
            return await asyncio.wait_for(queue.get(), timeout=poll)
        except asyncio.TimeoutError:
            if stop_event.is_set():
                return None


async def run_learning(parent, timer, queue_name, update, losses, publish_state, checkpoint_obj):
    assert hasattr(parent, queue_name)
    parent.idx = 0
    queue = getattr(parent, queue_name)
    args = parent.args
    cuda = torch.device(parent.device).type == "cuda"
    ahead = None                      # next host batch, already travelling to the device on the copy stream
    can_prefetch = cuda and hasattr(parent.to_gpu, "prefetch")
    engine = getattr(parent, "engine", None)
    pipelined = (cuda and hasattr(engine, "snapshot_async") and getattr(losses, "accepts_snapshot", False)
                 and os.environ.get("B200RL_PIPELINE", "1") != "0")
    pending = None                    # (idx, snapshot handle) of the iteration whose host side is still to be done

    def finish(idx, handle):
        """Host side of iteration `idx`, from its stream-ordered snapshot (never touches live device state)."""
        host_flat, L = engine.snapshot_wait(handle)
        parent.pub_model(parent.model.actor.state_dict_from_flat(host_flat))
        if L is not None:
            loss, detached = losses(parent, L)
            parent.log_loss_tensorboard(timer, loss, detached)
        if idx % args.model_save_interval == 0 and args.model_dir is not None:
            # checkpoint iterations are finished synchronously (below), so the live model IS this iteration's
            torch.save(checkpoint_obj(parent), os.path.join(args.model_dir, f"{args.algo}_{idx}.pt"))

    while not parent.stop_event.is_set():
        batch_args = None
        with timer.timer("learner-throughput", check_throughput=True):
            with timer.timer("learner-batching-time"):
                if ahead is not None:
                    batch_args, ahead = ahead, None
                else:
                    if pending is not None and queue.empty():
                        # nothing to overlap with: the feeder is slower than the learner (the normal case with real workers), so
                        # publish / log iteration i NOW instead of after the next batch has been collected -- otherwise the
                        # workers would gather all of batch i+1 under model i-1.  The wait for the GPU is cooperative: the
                        # batching coroutine keeps running (it ships the next batch while this update computes); if a batch
                        # arrives first, the update for it is enqueued before the host side of iteration i is finished.
                        while not engine.snapshot_ready(pending[1]) and queue.empty():
                            await asyncio.sleep(0)
                        if queue.empty():
                            finish(*pending)
                            pending = None
                    batch_args = await _next_batch(queue, parent.stop_event)
            if batch_args is not None:
                with timer.timer("learner-forward-time"):
                    dev_batch = parent.to_gpu(*batch_args)
                    for _ in range(args.K_epoch):
                        with timer.timer("learner-backward-time"):
                            update(parent, dev_batch)
                    if pipelined:
                        handle = engine.snapshot_async(with_losses=parent.idx % args.loss_log_interval == 0)
                    if can_prefetch and not queue.empty():       # overlap the next H2D with this update
                        ahead = queue.get_nowait()
                        if ahead is not None:
                            parent.to_gpu.prefetch(*ahead)
                    if cuda and not pipelined:
                        torch.cuda.current_stream(parent.device).synchronize()
                if pipelined:
                    if pending is not None:
                        finish(*pending)
                        pending = None
                    if parent.idx % args.model_save_interval == 0 and args.model_dir is not None:
                        finish(parent.idx, handle)               # rare: no later update may be in flight while the model is pickled
                    else:
                        pending = (parent.idx, handle)
                else:
                    parent.pub_model(publish_state(parent))
                    if parent.idx % args.loss_log_interval == 0:
                        loss, detached = losses(parent)
                        parent.log_loss_tensorboard(timer, loss, detached)
                    if parent.idx % args.model_save_interval == 0 and args.model_dir is not None:
                        torch.save(checkpoint_obj(parent), os.path.join(args.model_dir, f"{args.algo}_{parent.idx}.pt"))
                parent.idx += 1
            if parent.heartbeat is not None:
                parent.heartbeat.value = time.time()
        await asyncio.sleep(0)      # yield to the batching coroutine (the reference sleeps 1 ms here)
    if pending is not None:
        finish(*pending)


def on_policy_batch(dev_batch):
    obs, act, rew, logits, log_prob, is_fir, hx, cx = dev_batch
    return dict(obs=obs, act=act, rew=rew, logits=logits, log_prob=log_prob, is_fir=is_fir, hx=hx, cx=cx)


def on_policy_losses(parent, L=None):
    L = parent.engine.losses() if L is None else L
    detached = {"policy-loss": HostScalar(L["policy-loss"]), "value-loss": HostScalar(L["value-loss"]),
                "policy-entropy": HostScalar(L["policy-entropy"]),
                "ratio": RatioStats(L["ratio-min"], L["ratio-max"], L["ratio-avg"])}
    return HostScalar(L["loss"]), detached


on_policy_losses.accepts_snapshot = True      # may be fed the losses of a stream-ordered snapshot (pipelined schedule)
