"""N>1 path on CPU: world_size-2 gloo job (torchrun, 127.0.0.1) for sharding + the return-statistics
all-reduce, and bench.py's reference arm under torchrun (rank 0 alone runs and prints)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from gym_supplychain_b200.dist import shard_range, summarize_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(nproc, script_args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


def test_shard_range_is_a_contiguous_balanced_partition():
    for n, w in ((8_388_608, 8), (1001, 2), (7, 8), (16_777_216, 4), (5, 1)):
        cover, sizes = [], []
        for r in range(w):
            first, cnt = shard_range(n, r, w)
            cover += list(range(first, first + cnt)) if n < 5000 else []
            sizes.append(cnt)
            if r:
                assert first == prev_end
            prev_end = first + cnt
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1 and prev_end == n
        if n < 5000:
            assert cover == list(range(n))


def test_two_rank_gloo_allreduce_of_return_stats(tmp_path):
    res = _torchrun(2, [os.path.join(ROOT, "tests", "_dist_worker.py"), str(tmp_path)])
    assert res.returncode == 0, res.stderr[-2000:]
    out = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(2)]
    assert out[0]["first"] == 0 and out[0]["n_local"] == 501 and out[1]["first"] == 501 and out[1]["n_local"] == 500
    returns = -1e6 * (1.0 + np.sin(np.arange(1001) * 0.37))
    want = [1001, returns.sum(), (returns ** 2).sum(), returns.min(), returns.max()]
    for o in out:                                        # every rank ends with the job-wide statistics
        assert np.allclose(o["global"], want, rtol=1e-12)
        cnt, mean, std, mn, mx = o["summary"]
        assert cnt == 1001 and np.isclose(mean, returns.mean()) and np.isclose(std, returns.std(), rtol=1e-9)
        # bench.py's double-buffered exchange (consumed one step late, drained at the end) folds to the same totals
        assert o["pending_before_drain"] and np.allclose(o["overlapped"], want, rtol=1e-12)


def test_summarize_stats_empty():
    import torch
    cnt, mean, *_ = summarize_stats(torch.tensor([0.0, 0, 0, 1e300, -1e300], dtype=torch.float64))
    assert cnt == 0 and np.isnan(mean)


def test_bench_reference_arm_under_torchrun_prints_once():
    res = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--rollout-T", "8"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if "unavailable" in d:                               # baseline/_ref is installed by build() where /root/reference exists
        assert not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gym_supplychain"))
        return
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "env-steps/s" and d["config"]["id"] == "c4"
    # the `config` object is the one our arm prints for the same command line (the driver compares them)
    import bench
    want = bench.full_config("c4", bench.parse_args(["--gpus", "2", "--rollout-T", "8"]), 2)
    assert d["config"] == json.loads(json.dumps(want))


def test_reference_install_is_the_unmodified_reference():
    """baseline/_ref (git-ignored, made by build()) holds byte-identical copies of the reference's modules."""
    import filecmp
    import pytest
    ref, inst = "/root/reference/gym_supplychain/envs", os.path.join(ROOT, "baseline", "_ref", "gym_supplychain", "envs")
    if not (os.path.isdir(ref) and os.path.isdir(inst)):
        pytest.skip("reference tree or its install not present")
    names = [f for f in os.listdir(ref) if f.endswith(".py")]
    assert len(names) >= 8
    for f in names:
        assert filecmp.cmp(os.path.join(ref, f), os.path.join(inst, f), shallow=False), f
