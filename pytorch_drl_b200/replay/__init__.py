from .sumtree import SumTree, PrioritizedReplayIndex

__all__ = ['SumTree', 'PrioritizedReplayIndex']
