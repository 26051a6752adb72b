"""Multi-rank path on CPU: world_size 2 (and 3) over gloo.  Clusters are sharded
cluster i -> rank i % world, each rank writes its own batch files, rank 0 writes the
aggregate; everything must equal the single-process reference golden."""
import glob
import os
import socket
import subprocess
import sys

import pytest

from fandom_search_b200 import parallel
from tests.util import compare_records, read_csv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_sharded_analyze_on_two_gpus(golden_dir, tmp_path):
    """Same as below on real devices: 2 ranks, one B200 each, NCCL process group."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run_world(golden_dir, tmp_path, 2, real=True)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_analyze_equals_single_process_golden(golden_dir, tmp_path, world):
    _run_world(golden_dir, tmp_path, world, real=False)


def _run_world(golden_dir, tmp_path, world, real):
    os.symlink(os.path.join(golden_dir, "fanworks"), tmp_path / "fanworks")
    os.symlink(os.path.join(golden_dir, "script.txt"), tmp_path / "script.txt")
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if real:
            env["FS_TEST_REAL_DEVICE"] = "1"
        else:
            env["CUDA_VISIBLE_DEVICES"] = ""
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"),
                                       golden_dir, str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    # each rank announced only its own clusters
    for rank, o in enumerate(outs):
        seen = [int(l.split()[2]) for l in o.splitlines() if l.startswith("Processing cluster")]
        assert seen == [i for i in range(3) if i % world == rank]
    aggs = glob.glob(str(tmp_path / "match-6gram-2*.csv"))
    assert len(aggs) == 1                                   # only rank 0 writes the aggregate
    got = read_csv(aggs[0])
    want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
    compare_records(got, want, tol=1e-12, basename=False)
    assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in want]
    for i in range(3):
        b = read_csv(str(tmp_path / ("match-6gram-batch-%d.csv" % i)), header=False)
        wb = read_csv(os.path.join(golden_dir, "golden_exhaustive.batch%d.csv" % i), header=False)
        assert [(r[0], r[1]) for r in b] == [(r[0], r[1]) for r in wb]


def test_cluster_owner_is_round_robin():
    assert [parallel.cluster_owner(i, 4) for i in range(9)] == [0, 1, 2, 3, 0, 1, 2, 3, 0]
