"""Host-side logic of bench.py that needs no GPU: the clock sampler's selection of samples, the shared
workload `config` object of the two arms."""
import importlib.util
import os
import sys

from conftest import ROOT

spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def _sampler(rows):
    s = bench.ClockSampler.__new__(bench.ClockSampler)

    class P:
        def terminate(self):
            pass
    s.p, s.rows = P(), rows
    return s


def row(mhz, cap="Not Active"):
    return [str(mhz), "1965", "700.0", "Not Active", "Not Active", "Not Active", cap]


def test_clock_sampler_prefers_samples_inside_the_timed_region():
    rows = [(10.0, row(1900)), (10.1, row(1650, "Active")), (10.2, row(1660, "Active")), (10.6, row(1950))]
    c = _sampler(rows).stop(10.08, 10.25)     # the window tolerates 50 ms before / 150 ms after the region
    assert c["samples"] == 2 and c["samples_inside_timed_region"] == 2
    assert c["sm_mhz"] in (1650.0, 1660.0) and c["reasons"] == ["sw_power_cap"] and c["sm_max_mhz"] == 1965.0


def test_clock_sampler_falls_back_to_the_nearest_samples_for_a_short_region():
    # a 40 ms timed region (20 steps at 8 GPUs) between two 100 ms samples: the nearest ones are reported, flagged
    rows = [(9.0, row(1900)), (9.85, row(1780, "Active")), (10.25, row(1775, "Active")), (12.5, row(1965))]
    c = _sampler(rows).stop(9.98, 10.02)
    assert c is not None and c["samples_inside_timed_region"] == 0 and c["samples"] >= 2
    assert 1775.0 <= c["sm_mhz"] <= 1780.0
    # nothing within a second of the region: no clocks rather than misleading ones
    assert _sampler([(1.0, row(1965))]).stop(9.98, 10.02) is None


def test_both_arms_describe_the_workload_with_the_same_config_object():
    mesh = __import__("importlib").import_module("fesom2-accelerate_b200.mesh")
    gm = mesh.make_workload("pi")
    a = bench.workload_config("pi", gm)
    b = bench.workload_config("pi", mesh.make_workload("pi"))
    assert a == b and a["workload"] == "pi" and a["node_level_updates"] == gm.S_n()
    assert a["alg_bytes_per_step"] == gm.bytes_alg() and "partitions" not in a and "device" not in a
