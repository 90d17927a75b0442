"""Window statistics for the pipeline's monitoring keys - stand-in for the reference's `monitoring.py`.

The reference wraps two third-party libraries that are neither installed here nor vendored under `/root/reference`:
`apphb` (Application Heartbeats; unpinned in `pyproject.toml`) for the window arithmetic and `energymon` for energy.
This module keeps the reference's module-level API - `init`, `add_key`, `iteration_start`, `iteration`,
`get_locked_context` (with the `get_tag / get_window_size / get_window_perf / get_window_work / get_window_heartrate
/ ...` getters the adaptive-quantization hooks read, `runtime.py:121-216`) and `finish` - and restates the heartbeat
bookkeeping: every completed iteration is one heartbeat with a duration, a work amount and an accuracy value;
"instant" is the last heartbeat, "window" the sums over the last `window_size` heartbeats, "global" the sums over
all of them; every rate is a sum divided by the matching time sum. Parity of that arithmetic is unpinned (no apphb to
run); the hooks' decisions are tested on synthetic window values instead (`tests/test_adaptive_cpu.py`).

Difference that matters on a GPU: an iteration's duration may be SUPPLIED (`seconds=`), e.g. a hop's device-side
transfer time from CUDA events, instead of being the host time between `iteration_start` and `iteration`. Energy is
not measured (getters return 0). CSV logs (`<key>.csv`) are written only when `MONITORING_CSV=1`.
"""
import collections
import csv
import logging
import os
import threading
import time
from contextlib import contextmanager
from typing import Dict, Optional, Union

logger = logging.getLogger(__name__)

ENV_CSV: str = "MONITORING_CSV"
ENV_CSV_FILE_MODE: str = "CSV_FILE_MODE"


class _Heartbeats:
    """Heartbeat record of one key."""

    def __init__(self, window_size: int, log_name: Optional[str]):
        self.window_size = max(1, int(window_size))
        self.tag = 0
        self.window = collections.deque(maxlen=self.window_size)   # (seconds, work, accuracy)
        self.total = [0.0, 0, 0]
        self.log_name = log_name

    def beat(self, seconds: float, work, accuracy) -> None:
        self.window.append((seconds, work, accuracy))
        self.total[0] += seconds
        self.total[1] += work
        self.total[2] += accuracy
        self.tag += 1

    def sums(self, scope: str):
        if scope == 'instant':
            return self.window[-1] if self.window else (0.0, 0, 0)
        if scope == 'window':
            return tuple(sum(col) for col in zip(*self.window)) if self.window else (0.0, 0, 0)
        return tuple(self.total)

    def count(self, scope: str) -> int:
        return {'instant': min(1, self.tag), 'window': len(self.window), 'global': self.tag}[scope]


def _rate(amount, seconds: float) -> float:
    return amount / seconds if seconds > 0 else 0.0


class MonitorContext:
    """Keyed heartbeat records with the getters of the reference's `MonitorContext`
    (`src/pipeedge/monitoring/__init__.py:228-331`); `key` is a keyword argument there too."""

    def __init__(self):
        self._records: Dict[str, _Heartbeats] = {}

    def keys(self) -> tuple:
        """Monitored keys."""
        return tuple(self._records)

    def add_heartbeat(self, key: str, window_size: int, log_name: Optional[str] = None) -> None:
        """Start monitoring `key`."""
        if key in self._records:
            raise KeyError(f"key already exists: {key}")
        self._records[key] = _Heartbeats(window_size, log_name)
        if log_name is not None:
            with open(log_name, mode=os.getenv(ENV_CSV_FILE_MODE, 'w'), encoding='utf8', newline='') as f:
                csv.writer(f).writerow(['Tag', 'Time (s)', 'Work', 'Accuracy', 'Window Time (s)', 'Window Heart Rate (/s)',
                                        'Window Performance (/s)', 'Global Time (s)', 'Global Heart Rate (/s)',
                                        'Global Performance (/s)'])

    def heartbeat(self, key: str, seconds: float, work, accuracy) -> None:
        """Record one completed iteration of `key`."""
        rec = self._records[key]
        rec.beat(seconds, work, accuracy)
        if rec.log_name is not None:
            with open(rec.log_name, mode='a', encoding='utf8', newline='') as f:
                csv.writer(f).writerow([rec.tag - 1, f"{seconds:.9f}", work, accuracy, f"{self.get_window_time_s(key=key):.9f}",
                                        f"{self.get_window_heartrate(key=key):.6f}", f"{self.get_window_perf(key=key):.6f}",
                                        f"{self.get_global_time_s(key=key):.9f}", f"{self.get_global_heartrate(key=key):.6f}",
                                        f"{self.get_global_perf(key=key):.6f}"])

    def get_tag(self, key: str) -> int:
        """Number of heartbeats issued so far (the next heartbeat's tag)."""
        return self._records[key].tag

    def get_window_size(self, key: str) -> int:
        """Window period in heartbeats."""
        return self._records[key].window_size


def _install_getters() -> None:
    """`get_{instant,window,global}_{time_s,heartrate,work,perf,accuracy,accuracy_rate,energy_j,power_w}(key=...)`."""
    def make(scope: str, what: str):
        def getter(self, key: str) -> Union[int, float]:
            rec = self._records[key]   # pylint: disable=protected-access
            seconds, work, accuracy = rec.sums(scope)
            return {'time_s': seconds, 'heartrate': _rate(rec.count(scope), seconds), 'work': work,
                    'perf': _rate(work, seconds), 'accuracy': accuracy, 'accuracy_rate': _rate(accuracy, seconds),
                    'energy_j': 0.0, 'power_w': 0.0}[what]
        getter.__name__ = f"get_{scope}_{what}"
        return getter
    for scope in ('instant', 'window', 'global'):
        for what in ('time_s', 'heartrate', 'work', 'perf', 'accuracy', 'accuracy_rate', 'energy_j', 'power_w'):
            setattr(MonitorContext, f"get_{scope}_{what}", make(scope, what))


_install_getters()

_ctx: Optional[MonitorContext] = None
_ctx_lock = threading.RLock()
_key_locks: Dict[str, threading.Lock] = {}
_window_size = 1
_started: Dict[tuple, int] = {}     # (thread ident, key) -> monotonic_ns at iteration_start
_work_types: Dict[str, str] = {}
_acc_types: Dict[str, str] = {}


def _log_name(key: str) -> Optional[str]:
    return key + '.csv' if os.getenv(ENV_CSV, '0') == '1' else None


def init(key: str, window_size: int, work_type: str = 'items', acc_type: str = 'acc') -> None:
    """Create the monitoring context with a first key (`monitoring.py:95-123`)."""
    global _ctx, _window_size   # pylint: disable=global-statement
    with _ctx_lock:
        _ctx = MonitorContext()
        _window_size = window_size
    add_key(key, work_type=work_type, acc_type=acc_type)


def add_key(key: str, work_type: str = 'items', acc_type: str = 'acc') -> None:
    """Monitor another key with the context's window size (`monitoring.py:143-151`)."""
    with _ctx_lock:
        if _ctx is None:
            return
        _ctx.add_heartbeat(key, _window_size, log_name=_log_name(key))
        _key_locks[key] = threading.Lock()
        _work_types[key] = work_type
        _acc_types[key] = acc_type


def finish() -> None:
    """Log the global figures and destroy the context (`monitoring.py:125-141`)."""
    global _ctx   # pylint: disable=global-statement
    with _ctx_lock:
        if _ctx is None:
            return
        for key in _ctx.keys():
            logger.info("%s: Global Time: %s sec, Rate: %s microbatches/sec, Work: %s %s, Perf: %s %s/sec", key,
                        _ctx.get_global_time_s(key=key), _ctx.get_global_heartrate(key=key),
                        _ctx.get_global_work(key=key), _work_types[key], _ctx.get_global_perf(key=key), _work_types[key])
        _ctx = None
        _key_locks.clear()
        _started.clear()
        _work_types.clear()
        _acc_types.clear()


@contextmanager
def get_locked_context(key: str):
    """Yield the `MonitorContext` with `key` locked, for a consistent read of several metrics (`monitoring.py:153-158`)."""
    with _key_locks[key]:
        yield _ctx


def iteration_start(key: str) -> None:
    """Start an iteration of `key` on this thread (`monitoring.py:181-188`)."""
    if _ctx is None:
        return
    ident = (threading.get_ident(), key)
    if ident in _started:
        raise KeyError(f"Thread iteration context already exists for key: {key}")
    _started[ident] = time.monotonic_ns()


def iteration(key: str, work: Union[int, float] = 1, accuracy: Union[int, float] = 0, safe: bool = True,
              seconds: Optional[float] = None) -> None:
    """Complete an iteration (`monitoring.py:190-214`). `seconds`: the iteration's duration measured elsewhere (e.g. CUDA
    events); otherwise the host time since this thread's `iteration_start(key)`. Without either, `safe=False` treats the
    call as the start marker of a heartbeat series (as the reference does for the 'output' key)."""
    if _ctx is None:
        return
    now = time.monotonic_ns()
    began = _started.pop((threading.get_ident(), key), None)
    with _key_locks[key]:
        if seconds is None:
            if began is None:
                if safe:
                    raise KeyError(f"No thread iteration context for key: {key}")
                last = _started.get(('series', key))
                _started[('series', key)] = now
                if last is None:
                    return                      # first marker: nothing to report yet
                began = last
            seconds = (now - began) / 1e9
        _ctx.heartbeat(key, seconds, work, accuracy)
        if _ctx.get_tag(key=key) % _ctx.get_window_size(key=key) == 0:
            logger.debug("%s: Window Time: %s sec, Rate: %s /sec, Work: %s %s, Perf: %s %s/sec", key,
                         _ctx.get_window_time_s(key=key), _ctx.get_window_heartrate(key=key),
                         _ctx.get_window_work(key=key), _work_types[key], _ctx.get_window_perf(key=key), _work_types[key])


class DeviceIterations:
    """Iterations whose duration is measured ON THE DEVICE: `start(key)` / `finish(key, work, accuracy)` record CUDA
    events on the current stream around work that was only enqueued; the pair is turned into a heartbeat
    (`iteration(key, ..., seconds=elapsed)`) once the device has passed both, at a later `start`/`finish`/`harvest`
    call - nothing here synchronises the stream. `event_factory` is injectable for tests."""

    def __init__(self, event_factory=None):
        if event_factory is None:
            import torch   # pylint: disable=import-outside-toplevel

            def event_factory():
                evt = torch.cuda.Event(enable_timing=True)
                evt.record()
                return evt
        self._event = event_factory
        self._open = {}                        # (thread ident, key) -> start event
        self._pending = collections.deque()    # (key, start, end, work, accuracy), in completion order per stream
        self._lock = threading.Lock()

    def start(self, key: str) -> None:
        """Mark the start of an iteration of `key` on this thread's current stream."""
        self.harvest()
        self._open[(threading.get_ident(), key)] = self._event()

    def finish(self, key: str, work: Union[int, float] = 1, accuracy: Union[int, float] = 0) -> None:
        """Mark the end of the iteration started by `start(key)` on this thread."""
        began = self._open.pop((threading.get_ident(), key), None)
        if began is None:
            raise KeyError(f"No device iteration in flight for key: {key}")
        with self._lock:
            self._pending.append((key, began, self._event(), work, accuracy))
        self.harvest()

    def harvest(self, drain: bool = False) -> int:
        """Report every finished pair (all of them, waiting if necessary, with `drain`); returns how many."""
        done = []
        with self._lock:
            while self._pending:
                key, began, ended, work, accuracy = self._pending[0]
                if not drain and not ended.query():
                    break
                self._pending.popleft()
                done.append((key, began, ended, work, accuracy))
        for key, began, ended, work, accuracy in done:
            ended.synchronize()
            iteration(key, work=work, accuracy=accuracy, seconds=began.elapsed_time(ended) * 1e-3)
        return len(done)
