"""PPO / PPO-Continuous learner coroutine -- drop-in for the reference's
`agents/learner_module/ppo/learning.py:13-126` (`async def learning(parent, timer)`).
The per-epoch body (forward, TD/GAE, clipped surrogate + smooth-L1 + entropy, backward,
clip_grad_norm_, RMSprop) is `parent.engine.update`: seven native kernel launches."""
from ..loop import run_learning, on_policy_batch, on_policy_losses


def update(parent, dev_batch):
    parent.engine.update(on_policy_batch(dev_batch))


async def learning(parent, timer):
    await run_learning(parent, timer, "batch_queue", update, on_policy_losses,
                       publish_state=lambda p: p.model.actor.export_state_dict(),
                       checkpoint_obj=lambda p: p.model)
