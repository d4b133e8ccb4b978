"""Host-side pieces of bench.py that can be checked without a GPU: the nvidia-smi clock sampler (started before the warm-up,
windowed on the timed region -- a round-2 run whose sampler was started right before a 0.45 s region came back with zero
samples) and the best-effort NUMA binding of multi-rank end-to-end runs."""
import os
import stat
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _fake_nvidia_smi(tmp_path, startup_s):
    p = tmp_path / "nvidia-smi"
    p.write_text("#!/bin/bash\nsleep %s\nwhile true; do echo \"1965, 1965, 600.5, Not Active, Not Active, Not Active, Active\"; sleep 0.05; done\n" % startup_s)
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(tmp_path)


def test_clock_sampler_waits_for_the_first_row_and_windows_on_the_timed_region(tmp_path, monkeypatch):
    import bench

    monkeypatch.setenv("PATH", _fake_nvidia_smi(tmp_path, 0.6) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.05)          # "warm-up" shorter than nvidia-smi's start-up
    s.begin()                 # must block until rows arrive
    assert s.rows, "begin() returned before nvidia-smi delivered a row"
    n_before = len(s.rows)
    t0 = time.monotonic()
    time.sleep(0.3)           # the timed region
    t1 = time.monotonic()
    c = s.stop()
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["power_w_max"] == 600.5
    assert 3 <= c["samples"] <= 12                     # ~6 rows at 50 ms inside the 0.3 s window (+ the row that unblocked begin())
    w = s.summary(t0 + 0.1, t1)                        # a sub-window sees fewer rows
    assert 1 <= w["samples"] < c["samples"] + 1 and n_before >= 1


def test_clock_sampler_without_nvidia_smi(monkeypatch, tmp_path):
    import bench

    monkeypatch.setenv("PATH", str(tmp_path))          # no nvidia-smi anywhere
    s = bench.ClockSampler(0)
    s.start()
    s.begin()
    c = s.stop()
    assert c["sm_mhz"] is None and c["reasons"] == ["nvidia-smi unavailable"]


def test_numa_binding_is_best_effort(monkeypatch):
    import bench

    before = os.sched_getaffinity(0)
    info = bench.bind_to_gpu_numa_node(0)              # no GPU here: must report, not raise, and leave the affinity alone
    assert info["bound"] is False and ("error" in info or "pci" in info)
    assert os.sched_getaffinity(0) == before
    monkeypatch.setenv("SFB_BENCH_NUMA", "0")
    assert bench.bind_to_gpu_numa_node(0) == {"bound": False, "disabled": True}
