"""SAC learner coroutine -- drop-in for the reference's
`agents/learner_module/sac/learning.py` (`async def learning(parent, timer)`): actor step,
temperature step, twin-critic step and `soft_update` are `parent.engine.update(batch)`."""
from ..loop import run_learning, on_policy_batch, HostScalar


def update(parent, dev_batch):
    parent.engine.update(on_policy_batch(dev_batch))


def losses(parent):
    L = parent.engine.losses()
    detached = {"policy-loss": HostScalar(L["policy-loss"]), "value-loss": HostScalar(L["value-loss"]),
                "loss_alpha": HostScalar(L["loss_alpha"]), "alpha": L["alpha"]}
    return HostScalar(L["loss"]), detached


async def learning(parent, timer):
    await run_learning(parent, timer, "batch_buffer", update, losses,
                       publish_state=lambda p: p.actor.export_state_dict(),
                       checkpoint_obj=lambda p: p.actor)
