"""Feedback-control helpers for the adaptive QuantPipe policy (reference `utils/controller.py`).

Same classes, constructor arguments and call signatures as the reference, so `utils/quant.py` and the driver's
adaptive hooks read identically; pinned by `tests/golden/adaptive.json` (traces produced by running the reference's
classes, `oracle/make_goldens.py`).
"""
from typing import Optional


class KalmanFilter:
    """Scalar Kalman filter (reference `controller.py:4-66`): estimates `x` from measurements `z = h * x + noise`.

    `Q` (process noise) and `R` (measurement noise) are public and constant, as in the reference."""
    # pylint: disable=invalid-name

    def __init__(self, x_hat_0: float = 0, p_0: float = 1):
        self._x_hat = x_hat_0
        self._p = p_0
        self.Q = 0.00001
        self.R = 0.01

    @property
    def x_hat(self) -> float:
        """The a-posteriori state estimate."""
        return self._x_hat

    def __call__(self, z: float, h: float = 1) -> float:
        """One time step with measurement `z` and measurement model `h`; returns the new estimate."""
        prior_cov = self._p + self.Q                                  # time update (the state model is x(k+1) = x(k))
        gain = (prior_cov * h) / ((h * prior_cov * h) + self.R)
        self._x_hat = self._x_hat + gain * (z - h * self._x_hat)      # measurement update
        self._p = (1.0 - gain * h) * prior_cov
        return self._x_hat


class AdaptiveIntegralXupController:
    """Adaptive integral controller of a speed-up ("X-up") signal (reference `controller.py:69-148`).

    The plant is modelled as `y = base_workload * u`; a Kalman filter tracks `base_workload`, whose reciprocal
    replaces the integral gain. `u` is clamped to `[1, u_max]` (anti-windup). `pole` in `[0, 1)` trades reaction
    speed for noise rejection."""
    # pylint: disable=invalid-name

    def __init__(self, reference: float, u_0: float, u_max: float = float('inf'), pole: float = 0,
                 kf_kwargs: Optional[dict] = None):
        # pylint: disable=too-many-arguments
        self.reference = reference
        self._u = u_0
        self._u_max = u_max
        self.pole = pole
        self._kalman_filt = KalmanFilter(**(kf_kwargs or {}))

    @property
    def pole(self) -> float:
        """The closed-loop pole."""
        return self._pole

    @pole.setter
    def pole(self, pole: float) -> None:
        if not 0 <= pole < 1:
            raise ValueError("pole must be in range [0, 1)")
        self._pole = pole

    def __call__(self, y: float) -> float:
        """Measured output `y(k)` -> control signal `u(k+1)`."""
        base_workload = self._kalman_filt(y, h=self._u)
        error = self.reference - y
        u_next = self._u + (1 - self._pole) * (error / base_workload)
        self._u = max(min(u_next, self._u_max), 1)
        return self._u
