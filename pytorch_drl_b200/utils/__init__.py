from .stats import standardize
from .nested import slice_nested_tensor, clone_nested_tensor, MinibatchView
from . import checkpoint

__all__ = ['standardize', 'slice_nested_tensor', 'clone_nested_tensor', 'MinibatchView', 'checkpoint']
