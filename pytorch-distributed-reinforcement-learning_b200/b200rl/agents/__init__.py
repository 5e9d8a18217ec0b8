"""Binds each algorithm's `learning()` coroutine onto `LearnerBase.learning_<algo>` with
the same decorator-factory names the reference exports (`agents/__init__.py:8-55`)."""
from .learner_module.ppo.learning import learning as alearning_ppo
from .learner_module.impala.learning import learning as alearning_impala
from .learner_module.v_mpo.learning import learning as alearning_v_mpo
from .learner_module.sac.learning import learning as alearning_sac
from .learner_module.sac_continuous.learning import learning as alearning_sac_continuous


def _awrapper(alearning):
    def factory(timer):
        def outer(func):
            async def inner(self, *args, **kwargs):
                return await alearning(self, timer, *args, **kwargs)
            return inner
        return outer
    return factory


ppo_awrapper = _awrapper(alearning_ppo)
impala_awrapper = _awrapper(alearning_impala)
v_mpo_awrapper = _awrapper(alearning_v_mpo)
sac_awrapper = _awrapper(alearning_sac)
sac_continuous_awrapper = _awrapper(alearning_sac_continuous)
