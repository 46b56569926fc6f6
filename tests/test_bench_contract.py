"""bench.py's line contract, as far as it can be exercised without a GPU: the reference arm (`--impl reference`: the restated
reference path on the host cores) prints one JSON line with the keys the driver reads, under torchrun only rank 0 prints, and
the product arm refuses to run without a CUDA device (no CPU fallback) and with the diagnostic skip switch set."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=240):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                          env=dict(os.environ, **(env or {})), cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--arch", "micro", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["arch"] == "micro" and "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_only_rank0_prints():
    r = _run(["--impl", "reference", "--arch", "micro", "--steps", "1", "--warmup", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the product arm is what `bench.py` runs; nothing to refuse
    r = _run(["--steps", "1", "--warmup", "1", "--arch", "micro"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_bench_refuses_the_diagnostic_skip_switch(monkeypatch):
    import pytest
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("OH_ENC_CHUNK", "16")
    monkeypatch.delenv("OH_SKIP_CATS", raising=False)
    assert bench.set_knobs().get("OH_ENC_CHUNK") == "16"  # every knob that is set is reported in config.knobs
    monkeypatch.setenv("OH_SKIP_CATS", "8")
    with pytest.raises(SystemExit, match="OH_SKIP_CATS"):
        bench.set_knobs()
