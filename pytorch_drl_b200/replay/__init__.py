from .sumtree import PrioritizedReplayBuffer, PrioritizedReplayIndex, SumTree

__all__ = ['SumTree', 'PrioritizedReplayIndex', 'PrioritizedReplayBuffer']
