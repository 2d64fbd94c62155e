"""Drop-in for drl/utils/stats.py."""
import torch as tc

from .. import ops


def standardize(tensor: tc.Tensor) -> tc.Tensor:
    """Empirical z-scores with the unbiased std and no epsilon (drl/utils/stats.py:4-17), on the
    last dimension; a 2-D input is standardised per row (= per env)."""
    return ops.standardize(tensor)
