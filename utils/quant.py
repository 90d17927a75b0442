"""Quantization policy utilities (reference `utils/quant.py`): bit-width selection under a data-movement constraint.

Pinned by `tests/golden/adaptive.json` (outputs of the reference's functions, `oracle/make_goldens.py`).
"""
import math
from typing import List, Tuple
import torch
from pipeedge_b200.quantization.basic_op import compression_factor
from .controller import AdaptiveIntegralXupController


def constrain_max_bitwidth(t_max: torch.Tensor, d_size: torch.Tensor, d_speed: torch.Tensor,
                           bw_max: torch.Tensor) -> torch.Tensor:
    """Largest bit-width whose packed payload moves within `t_max` (reference `quant.py:9-39`); 0 if none does.

    Packing is discrete: `floor(bw_max / b)` codes share one source word, so the payload shrinks by that integer
    factor, not by `b / bw_max`. `d_size / d_speed` is the un-quantised transfer time."""
    candidates = torch.arange(bw_max, -1, -1, dtype=torch.int)           # bw_max .. 0
    shrink = compression_factor(candidates[:-1]).to(dtype=torch.int)     # integer codes-per-word of bw_max .. 1
    fraction = torch.hstack((shrink.reciprocal(), torch.tensor(0)))      # payload fraction; bit-width 0 sends nothing
    affordable = torch.div(d_speed * t_max, d_size)                      # payload fraction the budget allows (0/0 -> nan)
    return candidates[affordable >= fraction][0]


class AdaptiveBitwidthPerformanceController(AdaptiveIntegralXupController):
    """Splits each window between two adjacent bit-widths so that the send rate meets `perf_constraint`
    (reference `quant.py:42-107`). Speed-up is modelled as `max_bitwidth / bitwidth` (perfect packing)."""

    def __init__(self, perf_constraint: float, bitwidths: List[int], bitwidth_start: int):
        self._bitwidths = sorted(bitwidths, reverse=True)
        self._speedups = [self._bitwidths[0] / b for b in self._bitwidths]
        super().__init__(perf_constraint, self._bitwidths[0] / bitwidth_start, u_max=self._speedups[-1])

    def __call__(self, perf_measured: float, window_len: int) -> Tuple[int, int, int]:
        """Measured performance -> `(bitwidth_1, bitwidth_2, iterations to spend in bitwidth_1)` for the next window."""
        target = super().__call__(perf_measured)
        slow = max(0, sum(1 for s in self._speedups if s <= target) - 1)
        fast = min(slow + 1, len(self._speedups) - 1)
        xup_slow, xup_fast = self._speedups[slow], self._speedups[fast]
        # time shares x, 1 - x of the two settings must average to the target period:
        #   1 / target = x / xup_slow + (1 - x) / xup_fast
        if math.isclose(xup_slow, xup_fast):
            share = 0
        else:
            share = (xup_slow * (xup_fast - target)) / (target * (xup_fast - xup_slow))
        return (self._bitwidths[slow], self._bitwidths[fast], round(window_len * share))
