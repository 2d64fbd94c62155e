from .credit_assignment import GAE, AdvantageEstimator, get_credit_assignment_op
from .samplers import IndependentSampler, IndicesIterator, Sampler
from .schedules import LinearSchedule
from .ppo import PPO, RolloutStorage
from .rollout import BatchedRolloutManager, DeviceRollout, MetadataManager
from .dqn import DQN

FusedPPO = PPO          # names to register next to the reference's own classes
FusedGAE = GAE

__all__ = ['PPO', 'FusedPPO', 'GAE', 'FusedGAE', 'AdvantageEstimator', 'IndependentSampler',
           'IndicesIterator', 'Sampler', 'LinearSchedule', 'RolloutStorage', 'get_credit_assignment_op',
           'BatchedRolloutManager', 'DeviceRollout', 'MetadataManager', 'DQN']
