"""Drop-ins for the stateful pieces of the reference's env-wrapper stack that sit on the learner
hot path: the observation `Normalizer`, the learner half of
`RandomNetworkDistillationWrapper` (predictor update + batched intrinsic reward) and the reward normaliser
(`ReturnAcc` / `NormalizeRewardWrapper`) for all environments of a rank at once."""
from .normalize import Normalizer
from .normalize_reward import BatchedRewardNormalizer
from .rnd import RandomNetworkDistillation, TEACHER_KEYS, STUDENT_KEYS

__all__ = ['Normalizer', 'BatchedRewardNormalizer', 'RandomNetworkDistillation', 'TEACHER_KEYS', 'STUDENT_KEYS']
