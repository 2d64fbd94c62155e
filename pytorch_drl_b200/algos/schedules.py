"""Drop-in for drl/algos/common/schedules.py."""


class LinearSchedule:
    """Linear anneal in global_step/max_steps (schedules.py:18-29); host-side scalar."""

    def __init__(self, initial_value: float, final_value: float, max_steps: int):
        self._initial_value = initial_value
        self._final_value = final_value
        self._max_steps = max_steps

    def value(self, global_step: int) -> float:
        frac_done = global_step / self._max_steps
        s = min(max(0., frac_done), 1.)
        return self._initial_value * (1. - s) + self._final_value * s
