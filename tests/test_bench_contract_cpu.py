"""bench.py contract (CPU part): the reference arm runs without a GPU and prints ONE JSON line with the keys the driver
and the judge read; the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--workload", "bair64_b2_t4", "--steps", "1", "--warmup", "0", "--cpu-sample-frames", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "e2e",
              "cpu_baseline"):
        assert k in d, k
    assert d["value"] > 0 and d["config"]["workload"] == "bair64_b2_t4"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                      # on the GPU box the product arm is exercised by the driver itself
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0 and not any(l.startswith("{") and "\"value\"" in l for l in r.stdout.splitlines())
