"""CPU: the adaptive-QuantPipe policy code (`utils/quant.py`, `utils/controller.py`, the driver's three policy hooks and
the window statistics they read) against `tests/golden/adaptive.json`, produced by running the reference's own
`utils/{quant,controller}.py` (`oracle/make_goldens.py adaptive`)."""
import json
import os
import sys
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from utils import controller, quant as quantutil  # noqa: E402

with open(os.path.join(ROOT, 'tests', 'golden', 'adaptive.json'), encoding='utf8') as _f:
    GOLD = json.load(_f)


def test_constrain_max_bitwidth_grid():
    """Integer results identical on the whole grid, including the reference's IndexError on NaN budgets (0/0, inf*0)."""
    for t_max, d_size, d_speed, bw_max, want in GOLD['constrain_max_bitwidth']:
        args = (torch.tensor(t_max), torch.tensor(d_size), torch.tensor(d_speed), torch.tensor(bw_max))
        if want < 0:
            with pytest.raises(IndexError):
                quantutil.constrain_max_bitwidth(*args)
        else:
            assert int(quantutil.constrain_max_bitwidth(*args)) == want, (t_max, d_size, d_speed, bw_max)


def test_kalman_filter_traces():
    for case in GOLD['kalman']:
        filt = controller.KalmanFilter(x_hat_0=case['x0'], p_0=case['p0'])
        for z, h, want in zip(case['z'], case['h'], case['x_hat']):
            assert filt(z, h) == pytest.approx(want, rel=1e-13, abs=1e-13)
        assert filt.x_hat == pytest.approx(case['x_hat'][-1], rel=1e-13)


def test_xup_controller_traces():
    for case in GOLD['xup']:
        ctl = controller.AdaptiveIntegralXupController(case['reference'], case['u_0'], u_max=case['u_max'],
                                                       pole=case['pole'])
        for y, want in zip(case['y'], case['u']):
            assert ctl(y) == pytest.approx(want, rel=1e-12)
    with pytest.raises(ValueError):
        controller.AdaptiveIntegralXupController(1.0, 1.0, pole=1.0)


def test_bitwidth_controller_traces():
    """(bitwidth_1, bitwidth_2, iterations) identical to the reference controller for 4 constraint / start settings."""
    for case in GOLD['bitwidth_controller']:
        ctl = quantutil.AdaptiveBitwidthPerformanceController(0, GOLD['bitwidths'], case['start'])
        ctl.reference = case['constraint']
        for perf, want in zip(case['perf'], case['out']):
            assert list(ctl(perf, case['window'])) == want


def test_runtime_bitwidth_table():
    """`runtime.BITWIDTHS` (largest bit-widths with distinct packing ratios) equals the reference's list."""
    import runtime
    assert runtime.BITWIDTHS == GOLD['bitwidths']


# ---------------------------------------------------------------- window statistics + the driver's policy hooks
class _Shard(torch.nn.Module):
    def __init__(self, quant_bit=0, rate=0.0):
        super().__init__()
        self.register_buffer('quant_bit', torch.tensor(quant_bit), persistent=False)
        self.register_buffer('rate_constraint', torch.tensor(float(rate)), persistent=False)


@pytest.fixture
def send_monitor():
    import monitoring
    import runtime
    monitoring.init(runtime.MONITORING_KEY_SEND, 10, work_type='Mbits')
    yield monitoring, runtime
    monitoring.finish()


def test_window_statistics():
    """Instant = last heartbeat, window = last `window_size` heartbeats, global = all; rates are sums over time sums."""
    import monitoring
    monitoring.init('k', 3, work_type='Mbits')
    try:
        for seconds, work in ((0.1, 10), (0.2, 10), (0.1, 20), (0.4, 40)):
            monitoring.iteration('k', work=work, accuracy=2, seconds=seconds)
        with monitoring.get_locked_context('k') as ctx:
            assert ctx.get_tag(key='k') == 4 and ctx.get_window_size(key='k') == 3
            assert ctx.get_window_work(key='k') == 70 and ctx.get_window_time_s(key='k') == pytest.approx(0.7)
            assert ctx.get_window_perf(key='k') == pytest.approx(100.0)
            assert ctx.get_window_heartrate(key='k') == pytest.approx(3 / 0.7)
            assert ctx.get_instant_perf(key='k') == pytest.approx(100.0)
            assert ctx.get_global_work(key='k') == 80 and ctx.get_global_heartrate(key='k') == pytest.approx(4 / 0.8)
            assert ctx.get_window_accuracy(key='k') == 6 and ctx.get_global_energy_j(key='k') == 0.0
        monitoring.iteration_start('k')            # host-timed iteration, as the reference's hooks use it
        monitoring.iteration('k', work=1)
        with pytest.raises(KeyError):
            monitoring.iteration('k', work=1)      # no iteration_start on this thread
        monitoring.iteration('k', work=0, safe=False)   # heartbeat-series marker (the 'output' key pattern)
    finally:
        monitoring.finish()
    monitoring.iteration('k', work=1)              # no context: a no-op, like the reference


@pytest.mark.parametrize('rate,start_bit,want', [(1e6, 0, 2), (2e4, 0, 8), (1e3, 0, 0), (0.0, 8, 0), (2e4, 8, 2),
                                                 (1e4, 0, 16), (3.5e4, 0, 6), (4.5e4, 0, 4)])
def test_heuristic_hook(send_monitor, rate, start_bit, want):
    """10 sends of 38.7 Mbit in 1 ms each: bandwidth 38.7 Gbit/s, 387 Mbit per window (hand-computed thresholds)."""
    monitoring, runtime = send_monitor
    shard = _Shard(start_bit, rate)
    out = torch.zeros(8, 4, 4)
    for i in range(10):
        if i:
            runtime.forward_hook_set_quant_bandwidth_heuristic(shard, None, out)
            assert int(shard.quant_bit) == start_bit          # only adapts at window boundaries
        monitoring.iteration(runtime.MONITORING_KEY_SEND, work=38.7, seconds=1e-3)
    runtime.forward_hook_set_quant_bandwidth_heuristic(shard, None, out)
    assert int(shard.quant_bit) == want


@pytest.mark.parametrize('rate,want', [(0.0, 0), (1e3, 0), (2e4, 10), (1e5, 2), (1e7, 2)])
def test_heuristic2_hook(send_monitor, rate, want):
    """One micro-batch of 8 x 197 x 768 fp32 (38.7 Mbit) against the measured 38.7 Gbit/s."""
    monitoring, runtime = send_monitor
    shard = _Shard(0, rate)
    out = torch.zeros(8, 197, 768)
    for _ in range(10):
        monitoring.iteration(runtime.MONITORING_KEY_SEND, work=out.numel() * 32e-6, seconds=1e-3)
    runtime.forward_hook_set_quant_bandwidth_heuristic_2(shard, None, out)
    t_budget = 8 / rate if rate else float('inf')
    expect = int(quantutil.constrain_max_bitwidth(torch.tensor(t_budget), out.numel() * 32e-6, 38730.752, torch.tensor(32)))
    assert int(shard.quant_bit) == max(2, expect) % 32 == want


def test_controller_hook(send_monitor):
    """The hook walks the module through the controller's (bitwidth1, bitwidth2, iterations) plan window by window."""
    monitoring, runtime = send_monitor
    shard = _Shard(0, 4000.0)                       # items/s wanted
    out = torch.zeros(8, 4, 4)
    ref_ctl = quantutil.AdaptiveBitwidthPerformanceController(0, runtime.BITWIDTHS, 32)
    ref_ctl.reference = 4000.0
    seen = []
    plan = (0, 0, 0)
    for window in range(4):
        seconds = (4e-3, 3e-3, 1.5e-3, 2.5e-3)[window]          # per send: 2000, 2667, 5333, 3200 items/s
        for _ in range(10):
            monitoring.iteration(runtime.MONITORING_KEY_SEND, work=1.0, seconds=seconds)
            with monitoring.get_locked_context(runtime.MONITORING_KEY_SEND) as ctx:
                tag = ctx.get_tag(key=runtime.MONITORING_KEY_SEND)
                rate = ctx.get_window_heartrate(key=runtime.MONITORING_KEY_SEND) * 8
            if tag % 10 == 0:
                plan = ref_ctl(rate, 10)
            bw1, bw2, iters = plan
            runtime.forward_hook_set_quant_controller(shard, None, out)
            seen.append(int(shard.quant_bit))
            want = (bw1 if iters > 0 else bw2) % 32
            assert seen[-1] == want, (window, tag, plan)
            plan = (bw1, bw2, max(0, iters - 1))
    assert len(set(seen)) > 1                                   # the policy did move


def test_device_iterations_with_fake_events():
    """Device-timed heartbeats: pairs are reported only once the 'device' has passed the end event, in order, with the
    elapsed time between the events; `harvest(drain=True)` waits for the rest."""
    import monitoring

    class FakeEvent:
        clock = 0.0

        def __init__(self):
            self.t_ms = FakeEvent.clock
            self.reached = False
            self.waited = False

        def query(self):
            return self.reached

        def synchronize(self):
            self.waited = True
            self.reached = True

        def elapsed_time(self, other):
            return other.t_ms - self.t_ms

    created = []

    def factory():
        created.append(FakeEvent())
        return created[-1]

    monitoring.init('shard', 4, work_type='tensors', acc_type='layers')
    try:
        iters = monitoring.DeviceIterations(event_factory=factory)
        for i in range(3):
            FakeEvent.clock = 10.0 * i
            iters.start('shard')
            FakeEvent.clock = 10.0 * i + 2.0 + i          # 2, 3, 4 ms of "device" work
            iters.finish('shard', work=8, accuracy=24)
        with monitoring.get_locked_context('shard') as ctx:
            assert ctx.get_tag(key='shard') == 0           # nothing has completed on the device yet
        created[1].reached = True                          # first end event passed
        assert iters.harvest() == 1
        with monitoring.get_locked_context('shard') as ctx:
            assert ctx.get_tag(key='shard') == 1 and ctx.get_instant_time_s(key='shard') == pytest.approx(2e-3)
        created[5].reached = True                          # third done, second not: order is preserved, nothing reported
        assert iters.harvest() == 0
        assert iters.harvest(drain=True) == 2
        with monitoring.get_locked_context('shard') as ctx:
            assert ctx.get_tag(key='shard') == 3 and ctx.get_global_time_s(key='shard') == pytest.approx(9e-3)
            assert ctx.get_global_work(key='shard') == 24 and ctx.get_global_accuracy(key='shard') == 72
        with pytest.raises(KeyError):
            iters.finish('shard')
    finally:
        monitoring.finish()
