"""Host logic of the time-sharded path (physs_gp_b200/timeshard.py) on CPU: world_size 2 and 3 over gloo.
The compute building blocks are replaced by an oracle-backed stand-in (tests/dist/oracle_ops.py) so that
only the collective plumbing -- range split, all-gather of the range summaries, fold order, lml reduction,
terminal-state hand-off of the smoother -- is under test; the CUDA building blocks themselves are covered
by tests/test_gpu_pscan.py::test_time_shards_on_one_gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import filters as ofilters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_time_sharded_plumbing_matches_single_range(tmp_path, world):
    sys.path.insert(0, ROOT)
    from tests.dist.run_timeshard_gloo import problem
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29640 + world),
           os.path.join(ROOT, "tests", "dist", "run_timeshard_gloo.py"), str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=ROOT)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    prior, t, Y, R, *_ = problem()
    parts = [np.load(os.path.join(tmp_path, "rank%d.npz" % k)) for k in range(world)]
    assert [int(p["t0"]) for p in parts] == [0] + [int(p["t1"]) for p in parts[:-1]]
    mf = np.concatenate([p["mf"] for p in parts], axis=1)
    Pf = np.concatenate([p["Pf"] for p in parts], axis=1)
    ms = np.concatenate([p["ms"] for p in parts], axis=1)
    Ps = np.concatenate([p["Ps"] for p in parts], axis=1)
    for b in range(Y.shape[0]):
        lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(prior, t, Y[b], R[b], 0.0)
        ms_o, Ps_o = ofilters.smoother_sequential(prior, t, mf_o, Pf_o, full_state=True, jitter=0.0)
        for p in parts:                                   # every rank holds the whole-series lml
            assert abs(p["lml"][b] - lml_o) < 1e-9 * abs(lml_o)
        assert np.abs(mf[b] - mf_o[..., 0]).max() < 1e-9 and np.abs(Pf[b] - Pf_o).max() < 1e-9
        assert np.abs(ms[b] - ms_o[..., 0]).max() < 1e-9 and np.abs(Ps[b] - Ps_o).max() < 1e-9
