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


def test_ncu_capture_is_refused_for_other_kernel_sources(tmp_path, monkeypatch):
    """bench.py's roofline reads per-kernel instruction counts from an ncu capture; a capture taken from other kernel
    sources must be refused, not silently multiplied into this run's rate."""
    sys.path.insert(0, ROOT)
    import bench

    sha = bench.kernel_source_sha()
    assert len(sha) == 16 and sha == bench.kernel_source_sha()
    d, why = bench.load_ncu_captures("c3")
    if d is not None:
        assert d["kernel_source_sha"] == sha and {"scan", "sigma"} <= set(d)
    else:
        assert "refused" in why or "no capture" in why
    monkeypatch.setattr(bench, "kernel_source_sha", lambda: "0" * 16)
    d, why = bench.load_ncu_captures("c3")
    assert d is None and ("refused" in why or "no capture" in why)


def test_nominal_workload_sizes_and_sharding():
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    def ns(**kw):
        base = dict(config="c3", scaling="weak", perms=0, pairs=0, gpus=1)
        base.update(kw)
        return argparse.Namespace(**base)

    assert bench.nominal_permutations_per_step(ns()) == 100000
    assert bench.nominal_permutations_per_step(ns(gpus=8)) == 800000
    assert bench.nominal_permutations_per_step(ns(gpus=8, scaling="strong")) == 100000
    assert bench.nominal_permutations_per_step(ns(config="c5", gpus=8)) == 1000000
    assert bench.nominal_permutations_per_step(ns(config="c5", gpus=8, scaling="strong")) == 1000000
    assert bench.nominal_permutations_per_step(ns(config="c4", gpus=8)) == 2000000
    for total in (0, 1, 7, 100000, 100001):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = bench.shard(total, world, r)
                got.extend(range(lo, hi))
            assert got == list(range(total))
