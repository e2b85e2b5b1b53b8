"""bench.py contract checks that need no GPU: the reference arm prints the agreed JSON line; the product arm refuses to
run without a CUDA device (no silent CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_present():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--ref-row-stride", "295"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "permutations/s" and line["higher_is_better"] is True
    assert line["metric"] == "DTO permutations/sec at N=20k features" and line["config"]["features"] == 20000
    assert line["config"]["threshold_pairs"] == 589 * 589
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0 < line["value"] < 100  # the reference algorithm manages well under 100 permutations/s on any host


@pytest.mark.skipif(_gpu_present(), reason="a GPU is present")
def test_product_arm_fails_loudly_without_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--perms", "10"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "CUDA" in r.stderr or "NVIDIA" in r.stderr
    assert not r.stdout.strip()  # no JSON line is fabricated
