"""bench.py's process-level contract, checked without a GPU: stdout carries the JSON line and nothing else, ranks other
than 0 of the reference arm exit quietly, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, code=None):
    e = dict(os.environ)
    e.update(env or {})
    cmd = [sys.executable, "-c", code] if code else [sys.executable, os.path.join(ROOT, "bench.py"), *args]
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_stdout_is_reserved_for_the_json_line():
    code = ("import os, sys, bench\n"
            "bench._claim_stdout()\n"
            "print('library chatter')\n"                      # Python-level print
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"   # C-level write to file descriptor 1
            "bench.emit({'metric': 'm', 'value': 1.5})\n"
            "print('more chatter')\n")
    r = _run([], code=code)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines() == ['{"metric": "m", "value": 1.5}']
    assert "NCCL version" in r.stderr and "library chatter" in r.stderr and "more chatter" in r.stderr


def test_reference_arm_runs_on_rank_zero_only():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and r.stdout == ""
    assert "CUDA" in r.stderr
