"""Host-side mirror of the reference's engine boundary: the cases of the reference's own unit tests
(src/engine/validation.rs:139-203, src/engine/whisper.rs:741-997, src/queue/worker.rs:319-538), plus the
multi-GPU sharding / record gather exercised with world_size-2 gloo on CPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from openhush_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- validation.rs tests -------------------------------------------------------------------------
def test_valid_audio():
    info = E.validate_audio(np.full(16000, 0.5, np.float32), 16000)
    assert abs(info.duration_secs - 1.0) < 1e-3 and info.sample_count == 16000
    assert abs(info.rms - 0.5) < 1e-3 and info.min_value == 0.5 and info.max_value == 0.5


def test_empty_audio():
    with pytest.raises(E.Empty):
        E.validate_audio(np.zeros(0, np.float32), 16000)


def test_too_short_and_too_long():
    with pytest.raises(E.TooShort):
        E.validate_audio(np.zeros(800, np.float32), 16000)  # 50 ms
    with pytest.raises(E.TooLong) as ei:
        E.validate_audio(np.zeros(16000 * 7201, np.float32), 16000)
    assert "exceeds maximum 7200.0s" in str(ei.value)


def test_nan_inf_counts_and_rate():
    x = np.zeros(16000, np.float32)
    x[100] = x[200] = np.nan
    with pytest.raises(E.ContainsNaN) as ei:
        E.validate_audio(x, 16000)
    assert ei.value.count == 2
    x = np.zeros(16000, np.float32)
    x[5] = np.inf
    x[6] = -np.inf
    x[7] = np.inf
    with pytest.raises(E.ContainsInfinite) as ei:
        E.validate_audio(x, 16000)
    assert ei.value.count == 3
    with pytest.raises(E.InvalidSampleRate) as ei:
        E.validate_audio(np.zeros(44100, np.float32), 44100)
    assert str(ei.value) == "Unexpected sample rate: 44100Hz (expected 16000Hz)"


def test_error_order_matches_reference():
    # empty wins over the sample rate; the sample rate wins over the duration; NaN is reported before Inf
    with pytest.raises(E.Empty):
        E.validate_audio(np.zeros(0, np.float32), 8000)
    with pytest.raises(E.InvalidSampleRate):
        E.validate_audio(np.zeros(10, np.float32), 8000)
    x = np.zeros(16000, np.float32)
    x[0], x[1] = np.inf, np.nan
    with pytest.raises(E.ContainsNaN):
        E.validate_audio(x, 16000)


# ---- whisper.rs tests ----------------------------------------------------------------------------
def test_model_parsing_and_filenames():
    M = E.WhisperModel
    assert M.from_str("tiny") is M.Tiny and M.from_str("LARGE") is M.LargeV3 and M.from_str("largev3") is M.LargeV3
    with pytest.raises(ValueError):
        M.from_str("tiny.en")  # no .en variants exist in the reference
    assert M.LargeV3.filename() == "ggml-large-v3.bin" and M.Tiny.filename() == "ggml-tiny.bin"
    assert M.Medium.size_bytes() == 1_500_000_000


def test_missing_model_is_model_not_found():
    with pytest.raises(E.ModelNotFound) as ei:
        E.WhisperEngine.new("/nonexistent/ggml-base.bin", "en", False, True)
    assert "Run 'openhush model download base'" in str(ei.value)


def test_error_display():
    assert str(E.LoadFailed("x")) == "Failed to load model: x"
    assert str(E.TranscriptionFailed("FailedToEncode")) == "Transcription failed: FailedToEncode"
    assert str(E.ValidationFailed(E.Empty())) == "Audio validation failed: Audio is empty (no samples)"
    assert E.LANG_CODES[0] == "en" and E.LANG_CODES[98] == "su" and len(E.LANG_CODES) == 99


def test_load_failed_without_gpu(micro_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.LoadFailed):
        E.WhisperEngine.new(micro_path, "en", False, True)
    with pytest.raises(E.LoadFailed):  # device = "cpu" in the config: this library has no CPU path
        E.WhisperEngine.new(micro_path, "en", False, False)


# ---- worker.rs tests -----------------------------------------------------------------------------
def test_worker_without_engine_still_answers():
    w = E.TranscriptionWorker(None)
    job = E.TranscriptionJob(E.AudioBuffer(np.zeros(16000, np.float32), 16000), sequence_id=7, chunk_id=2, is_final=False)
    r = w.process_job(job)
    assert (r.text, r.sequence_id, r.chunk_id, r.is_final, r.duration_secs) == ("", 7, 2, False, 0.0)  # worker.rs:125-142
    w.unload_engine()  # no-op without an engine (worker.rs test)


# ---- multi-GPU host logic ------------------------------------------------------------------------
def test_shard_and_merge_roundtrip():
    world, n_jobs = 3, 10
    per_rank = []
    for r in range(world):
        rec = []
        for i, job in enumerate(E.shard_jobs(n_jobs, r, world)):
            rec += [i, 2, 0, 100 + job, 2, job, job + 1, 100 + job, 300, 1, job]
        per_rank.append(rec)
    merged = E.merge_records(per_rank, world)
    assert sorted(merged) == list(range(n_jobs))
    for job, segs in merged.items():
        assert segs == [(0, 100 + job, [job, job + 1]), (100 + job, 300, [job])]


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["OH_ROOT"])
from openhush_b200 import engine as E
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_jobs = 7
mine = E.shard_jobs(n_jobs, rank, world)
rec = []
for i, job in enumerate(mine):                      # what oh_full_pack_records yields: item, n_seg, (t0, t1, n_tok, ids...)*
    rec += [i, 1, 10 * job, 10 * job + 5, 3, job, job * 2, job * 3]
t = torch.tensor(rec, dtype=torch.int64)
n = torch.tensor([t.numel()])
sizes = [torch.zeros_like(n) for _ in range(world)]
dist.all_gather(sizes, n)
pad = torch.zeros(int(max(int(s) for s in sizes)), dtype=torch.int64)
pad[:t.numel()] = t
allrec = [torch.zeros_like(pad) for _ in range(world)]
dist.all_gather(allrec, pad)
merged = E.merge_records([a[:int(s)].tolist() for a, s in zip(allrec, sizes)], world)
assert sorted(merged) == list(range(n_jobs)), merged
for job, segs in merged.items():
    assert segs == [(10 * job, 10 * job + 5, [job, job * 2, job * 3])], (job, segs)
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_record_gather_gloo_world2(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, OH_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout


# ---- C++ host mirror (csrc/host/engine.hpp) -------------------------------------------------------
def test_cpp_host_mirror_cpu(micro_path):
    from openhush_b200 import lib
    lib.build()
    exe = os.path.join(lib.CSRC, "host", "host_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "HOST_TEST_OK" in r.stdout, r.stdout + r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, micro_path, "32000", "1"], capture_output=True, text=True, timeout=60)
        assert "LOAD_FAILED Failed to load model:" in r.stdout and "no CPU fallback" in r.stdout


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the restated reference path on the host cores) on the micro model: one JSON line with
    the keys the driver reads; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OH_REF_BUDGET_S="30")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--arch", "micro", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio_seconds_per_second" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] >= 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--arch", "micro", "--gpus", "2"],
                          capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=600)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
