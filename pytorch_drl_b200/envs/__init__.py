"""Drop-ins for the stateful pieces of the reference's env-wrapper stack that sit on the learner
hot path: the observation `Normalizer` and the learner half of
`RandomNetworkDistillationWrapper` (predictor update + batched intrinsic reward)."""
from .normalize import Normalizer
from .rnd import RandomNetworkDistillation, TEACHER_KEYS, STUDENT_KEYS

__all__ = ['Normalizer', 'RandomNetworkDistillation', 'TEACHER_KEYS', 'STUDENT_KEYS']
