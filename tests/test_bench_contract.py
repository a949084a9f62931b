"""The driver's benchmark contract (bench.py): the reference arm runs on CPU here and must print ONE JSON line with the
agreed keys; our arm must refuse to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built")
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gaussian_conv_pair_evals_per_s" and d["unit"] == "pair-evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["M"] == 1000000 and d["config"]["N"] == 1000000


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_uses_all_cores_under_torchrun_env():
    """``torch.distributed.run`` exports OMP_NUM_THREADS=1 to every rank; the reference arm must reset the OpenMP thread
    count itself (the N >= 2 reference lines ran on one thread) and report the count OpenMP really uses."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built")
    cores = len(os.sched_getaffinity(0))
    if cores < 2:
        pytest.skip("single-core container")
    vals = {}
    for tag, env in (("plain", {}), ("torchrun", {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})):
        r = _run(["--impl", "reference", "--gpus", "2" if env else "1", "--steps", "1", "--warmup", "0"], env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
        assert d["cpu_baseline"]["omp_max_threads"] == cores and d["cpu_baseline"]["cores"] == cores, d["cpu_baseline"]
        vals[tag] = d["value"]
    assert vals["torchrun"] > 0.6 * vals["plain"], vals       # same machine, same thread count: same rate
