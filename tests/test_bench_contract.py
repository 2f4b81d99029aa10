"""bench.py's JSON contract on the CPU: the reference arm (`--impl reference` = the oracle port timed on the host cores, the one
place outside tests/ and smoke() that executes oracle/) prints one line with the keys the driver reads; under a multi-rank launch
only rank 0 works.  Tiny workload so that it takes a second; the GPU arm is exercised by the driver on a B200."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["--genome-len", "20000", "--read-len", "500", "--steps", "2", "--warmup", "1", "--cpu-sample-reads", "2"]


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *TINY, *args],
                          capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    p = _run()
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "PHMM forward-backward GCUPS" and d["unit"] == "GCUPS"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    # the reference arm runs on the device arm's `config`: one builder serves both, key for key
    sys.path.insert(0, ROOT)
    import bench
    old_argv, sys.argv = sys.argv, ["bench.py", "--impl", "reference", *TINY]
    try:
        args = bench.parse()
    finally:
        sys.argv = old_argv
    assert list(d["config"]) == ["workload", "n_nodes", "reads_per_gpu_per_step", "n_active_nodes", "n_warmup", "l2", "timing"]
    assert d["config"]["reads_per_gpu_per_step"] == bench.B200_WAVE_READS == 1332
    assert d["config"] == bench.workload_config(args, d["config"]["n_nodes"], bench.B200_WAVE_READS)
    assert d["cpu_baseline"]["sample_bases_per_read"] == 8       # 2 + 1 steps: the longer sample; 4 bases from 11 steps on
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ["--gpus", "2"])
    assert p.returncode == 0 and p.stdout.strip() == "", (p.stdout, p.stderr)


def test_reference_arm_ignores_torchrun_single_thread_default():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use the host cores (round 1 reported `cores: 1` at N > 1)."""
    if len(os.sched_getaffinity(0)) < 2:
        return
    p = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, ["--gpus", "2"])
    assert p.returncode == 0, p.stderr
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == 2          # = min(host cores, reads of the sample)
    for key in ("sample_reads", "sample_bases_per_read", "rows_covered", "cells_per_s_per_core", "reference_hint_cells_per_s_per_core"):
        assert key in d["cpu_baseline"], key
