"""V-MPO learner coroutine -- drop-in for the reference's
`agents/learner_module/v_mpo/learning.py:14-144`.  The epoch body (GAE, batch-dim top-k,
softmax weights, temperature / KL duals, backward, clip, RMSprop over model + duals) is
`parent.engine.update(batch, parent.get_coef_alpha())`."""
from ..loop import run_learning, on_policy_batch, HostScalar


def update(parent, dev_batch):
    parent.engine.update(on_policy_batch(dev_batch), float(parent.get_coef_alpha()))


def losses(parent):
    L = parent.engine.losses()
    detached = {"policy-loss": HostScalar(L["policy-loss"]), "value-loss": HostScalar(L["value-loss"]),
                "loss-temperature": HostScalar(L["loss-temperature"]), "loss-alpha": HostScalar(L["loss-alpha"])}
    return HostScalar(L["loss"]), detached


async def learning(parent, timer):
    await run_learning(parent, timer, "batch_queue", update, losses,
                       publish_state=lambda p: p.model.actor.export_state_dict(),
                       checkpoint_obj=lambda p: p.model)
