"""Storage-side buffers of the learner path."""
from .rollout_assembler import RolloutAssembler  # noqa: F401
