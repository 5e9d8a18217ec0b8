"""IMPALA learner coroutine -- drop-in for the reference's
`agents/learner_module/impala/learning.py:13-114`.  The engine runs the V-trace head
(`csrc/head_kernels.cu: impala_head_kernel`) instead of the PPO head; everything else is
shared with PPO."""
from ..loop import run_learning, on_policy_batch, on_policy_losses


def update(parent, dev_batch):
    parent.engine.update(on_policy_batch(dev_batch))


async def learning(parent, timer):
    await run_learning(parent, timer, "batch_queue", update, on_policy_losses,
                       publish_state=lambda p: p.model.actor.export_state_dict(),
                       checkpoint_obj=lambda p: p.model)
