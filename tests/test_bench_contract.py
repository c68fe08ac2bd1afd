"""The reference arm of bench.py runs on CPU: check the one-JSON-line contract (keys the driver reads)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3
    from oracle import ref_bench
    want_kind = "reference" if ref_bench.reference_root() is not None else "port"     # baseline/_ref staged?
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1
    if want_kind == "reference":      # the unmodified Python PTGEnv ran: its trajectory hash and all three arrangements
        ref = d["config"]["reference"]
        assert len(ref["trajectory_sha256"]) == 64 and ref["single_env_steps"] == 5323
        assert d["value"] == max(ref["subproc_steps_per_s"], ref["dummy6_steps_per_s"], ref["single_env_steps_per_s"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_bench_lines_are_self_consistent():
    """The round's final GPU lines under profiles/ (written by bench.py on B200 boxes): every key of the contract is
    there and the figures follow from each other -- value = envs / time, roofline.achieved = algorithmic bytes / time,
    frac = achieved / peak -- so the numbers quoted in README / DESIGN.md can be re-derived from the files."""
    import pytest
    for n in (1, 2, 4):
        with open(os.path.join(ROOT, "profiles", f"bench_r02d_{n}gpu.json")) as fh:
            d = json.loads(fh.read())
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
            assert key in d, key
        c, r, e = d["config"], d["roofline"], d["e2e"]
        assert d["n_gpus"] == n and d["metric"] == "env-steps/sec" and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["steps"] == 20 and d["warmup"] == 5 and d["gpu_launches"] == 20          # one kernel launch per step
        assert c["envs_per_gpu"] == 1048576 and c["envs_total"] == n * 1048576 and "workload" in c
        assert d["value"] == pytest.approx(c["envs_total"] / (d["ms_per_step"] * 1e-3), rel=1e-6)
        bytes_per = c["bytes_per_env_step"]
        b = c["bytes_per_env_step_breakdown"]
        assert bytes_per == pytest.approx(b["state_action_obs_reward_done"] + b["rng_state_per_draw"] * b["draws_per_env_step"], abs=0.01)
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6539.9
        assert r["achieved"] == pytest.approx(bytes_per * c["envs_per_gpu"] / (r["kernel_ms"] * 1e-3) / 1e9, rel=1e-3)
        assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0.5 < r["frac"] < 1.0
        assert r["traffic"] is None or r["traffic"] < 1.1 * bytes_per * c["envs_per_gpu"]   # no wasted re-reads
        assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]
        assert e["h2d_bytes_per_step"] >= c["envs_per_gpu"] and e["d2h_bytes_per_step"] > 30 * c["envs_per_gpu"]
        assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] == d["clocks"]["sm_max_mhz"]
        if n == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["trajectory_match"] is True
            assert cb["trajectory_sha256_reference"] == cb["trajectory_sha256_cuda"]
            assert d["cpu_baseline_port"]["kind"] == "port"
        ss = c["strong_scaling"]
        assert ss["envs_total"] == 1048576 and ss["envs_per_gpu"] * n == 1048576
        for mode in ss["modes"].values():
            assert mode["episodes_in_gathered_records"] == 1048576          # the collective's record is non-zero
