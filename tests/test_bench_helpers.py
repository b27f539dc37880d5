"""bench.py pieces that run without a GPU: the clock sampler degrades to an explicit "no sampler" record instead of failing, the
issue / traffic records of profiles/traffic.json are read per launch (a multi-launch sweep is averaged over its launches)."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_without_nvml_or_nvidia_smi():
    bench = load_bench()
    sampler = bench.ClockSampler(0)
    sampler.start()
    mark = sampler.mark()
    rec = sampler.stop(mark, sampler.mark())
    assert set(rec) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    assert rec["samples"] == 0 or rec["sm_mhz"] is not None  # either nothing could be sampled (this box) or real samples


def test_traffic_records_are_per_launch():
    bench = load_bench()
    rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
        assert name in rec and rec[name]["warp_inst_executed"] > 0 and rec[name]["report"].endswith(".ncu-rep")
    # cfg5's report holds both passes of the fused pedigree sweep: totals of the sweep, divided by the launches of a sweep
    assert rec["cfg5"]["launches_in_report"] == 2
    per_launch, report = bench.recorded_traffic("cfg5", 3)
    assert abs(per_launch * 3 - rec["cfg5"]["bytes_per_launch"]) < 1 and report == rec["cfg5"]["report"]
    issue = bench.recorded_issue("cfg5", 1e9, 3)
    assert abs(issue["warp_instructions_per_launch"] * 3 - rec["cfg5"]["warp_inst_executed"]) < 1
    one, _ = bench.recorded_traffic("cfg3", 36)
    assert one == rec["cfg3"]["bytes_per_launch"]
