"""bench.py contract on CPU: the reference arm (`--impl reference`, the oracle on the host cores)
prints exactly ONE JSON line on stdout with the keys the driver reads; ranks other than 0 of a
torchrun launch print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, SPPM_BENCH_SAMPLE_S="0.5", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ref-main-320x240",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run({})
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "photons traced+gathered/sec" and d["unit"] == "photons/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"] == "ref-main-320x240"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["port_value"] > 0
    rc = cb["reference_code"]                      # the reference's own PhotonMap, when oracle/_ref/libref.so is there
    if rc is not None:
        assert rc["value"] > 0 and "onetime as shipped" in rc["sample"], rc
        assert cb["value"] == max(cb["port_value"], rc["value"])          # the faster CPU implementation is the baseline
        assert cb["kind"] == ("reference" if rc["value"] > cb["port_value"] else "port")
    else:
        assert cb["kind"] == "port" and "threads" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_workload_config_is_shared_by_both_arms_and_traffic_is_digest_bound(tmp_path, monkeypatch):
    """`config` of the JSON line defines the workload and is the same dict in both arms (arm-specific settings live in
    `impl_config`); the ncu-derived `roofline.traffic` is only quoted while profiles/ncu_traffic.json was captured from
    the kernel sources as they are now."""
    sys.path.insert(0, ROOT)
    import bench
    a = bench.workload_config("teapot-room-1024x768", 786429, "reference")
    b = bench.workload_config("teapot-room-1024x768", 786429, "reference")
    assert a == b and a["workload"] == "teapot-room-1024x768" and a["film"] == [1024, 768] and a["photons_per_round"] == 2_000_000
    assert "l2" in a and "model" not in a
    assert bench.WORKLOAD_DEFAULTS["stress-4k"] == dict(scaling="strong", update="progressive")   # BASELINE config 5
    assert bench.WORKLOAD_DEFAULTS["caustics-1080p"]["scaling"] == "strong"                          # BASELINE config 4
    d = bench.csrc_digest()
    assert len(d) == 16 and d == bench.csrc_digest()
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        t = json.load(f)
    tr, src = bench.ncu_traffic("photon_step_kernel", "teapot-room-1024x768")
    if t["csrc_digest"] == d:
        assert tr == t["kernels"]["photon_step_kernel@teapot-room-1024x768"]["dram_bytes"] and src
    else:
        assert tr is None and src is None      # captured from other kernel sources: not quoted
    assert bench.ncu_traffic("no_such_kernel", "teapot-room-1024x768") == (None, None)
