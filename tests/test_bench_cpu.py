"""Host-side logic of bench.py that the multi-rank timing depends on (no GPU)."""
import importlib.util
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_is_initialised_before_the_barrier_and_samples_only_once_armed(monkeypatch):
    """start() carries the slow part (NVML initialisation) and must not record; arm() opens the window.  With start() after the
    opening barrier, rank 0's NVML initialisation sat inside the other ranks' timed region (they wait for rank 0 in the gradient
    exchange) — the N = 2 step time read 24.6 ms instead of 8.8 ms."""
    calls = {"init": 0, "clock": 0}
    nv = types.ModuleType("pynvml")
    nv.NVML_CLOCK_SM = 1

    def init():
        calls["init"] += 1
    nv.nvmlInit = init
    nv.nvmlDeviceGetHandleByIndex = lambda i: ("h", i)
    nv.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965

    def clock(h, k):
        calls["clock"] += 1
        return 1965
    nv.nvmlDeviceGetClockInfo = clock
    nv.nvmlDeviceGetPowerUsage = lambda h: 300000
    nv.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4            # sw_power_cap
    monkeypatch.setitem(sys.modules, "pynvml", nv)
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    time.sleep(0.08)
    assert calls["init"] == 1 and calls["clock"] == 0 and s.sm == []     # initialised, nothing recorded yet
    s.arm()
    time.sleep(0.12)
    out = s.stop()
    assert out["samples"] >= 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["source"].startswith("nvml")


def test_timed_region_arms_the_sampler_after_the_barrier():
    """Order in the bench source: sampler.start() -> barrier() -> sampler.arm() -> first event record."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    i_start, i_arm = src.index("sampler.start()"), src.index("sampler.arm()")
    i_barrier = src.index("barrier()", i_start)
    i_rec = src.index("e0.record()", i_arm)
    assert i_start < i_barrier < i_arm < i_rec
