"""Multi-GPU checks (need >= 2 CUDA devices; run with `gpurun --gpus 2`):
N-sharded grid path == unsharded chain, bit for bit; chain split across ranks == single-rank run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        from bmm_mcmc_b200 import _lib
        return _lib.lib().bmm_device_count()
    except Exception:
        return 0


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["BMM_ROOT"])
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import bmm_mcmc_b200 as B
from bmm_mcmc_b200 import dist as bdist
bdist.init(rank, world, rank)
rng = np.random.default_rng(3)
N, P, K = 40_001, 64, 16
th = rng.uniform(0.1, 0.9, (K, P))
X = (rng.random((N, P)) < th[rng.integers(0, K, N)]).astype(np.int32)
lo, hi = bdist.shard_rows(N, world, rank)
out = {}
for prec in ("fp64", "fp32"):
    g = B.gibbs_stickbreaking(X[lo:hi], 8, K, alpha=1.0, burnin=1, seed=5, device=rank, precision=prec,
                              grid_path=True, n_global=N, row_offset=lo, probes=("counts",))
    out[prec] = g
# the same chain with relabelling: K x K cost partial sums of the two row blocks are all-reduced (fp64)
g = B.gibbs_full(X[lo:hi], 14, K, alpha=1.0, burnin=6, relabel=True, burnrelabel=3, seed=9, device=rank, precision="fp32",
                 grid_path=True, n_global=N, row_offset=lo)
out["rel"] = {k: g[k] for k in ("z", "z_original", "permutations", "theta")}
# a device-resident plan run three times: the captured sweep graph is replayed with fresh exchange numbers
from bmm_mcmc_b200 import api, _lib
ip = np.full((1, K), 1.0 / K); ith = np.ascontiguousarray(rng.uniform(0.2, 0.8, (1, P, K)))
plan = api.Plan(_lib.SAMPLER_STICKBREAKING, X[lo:hi], 8, K, 1.0, 0.5, 0.5, 1.0, 1.0, 1, False, 0, seed=5, device=rank,
                precision="fp32", grid_path=True, n_global=N, row_offset=lo, init_pi=ip, init_theta=ith)
for _ in range(3):
    plan.run()
g = plan.fetch()
plan.close()
out["plan"] = {k: g[k][0] for k in ("z", "theta", "pi")}
np.savez(os.path.join(os.environ["BMM_OUT"], "rank%d.npz" % rank), lo=lo, hi=hi,
         **{p + "_" + k: v for p, g in out.items() for k, v in g.items()})
bdist.finalize()
dist.destroy_process_group()
'''


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("p2p,cap", [("0", ""), ("1", ""), ("1", "4")])
def test_n_sharded_equals_unsharded(tmp_path, p2p, cap):
    """p2p = 1: counts exchanged by the one-shot push over IPC-mapped peer memory instead of NCCL.  cap = 4: the ranks'
    tensor kernels run on 4 CTAs each (~40 tiles per CTA: every mbarrier ring wraps), the unsharded comparison on 148."""
    import bmm_mcmc_b200 as B
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BMM_ROOT=ROOT, BMM_OUT=str(tmp_path), BMM_P2P=p2p)
    if cap:
        env["BMM_GRID_MAX_CTAS"] = cap
    for attempt in range(3):        # a probed "free" port can be taken again before torchrun binds it
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], env=env, timeout=300,
                           capture_output=True, text=True)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            break
    assert r.returncode == 0, r.stderr[-3000:]
    r = [np.load(tmp_path / ("rank%d.npz" % i)) for i in range(2)]
    rng = np.random.default_rng(3)
    N, P, K = 40_001, 64, 16
    th = rng.uniform(0.1, 0.9, (K, P))
    X = (rng.random((N, P)) < th[rng.integers(0, K, N)]).astype(np.int32)
    for prec in ("fp64", "fp32"):
        g = B.gibbs_stickbreaking(X, 8, K, alpha=1.0, burnin=1, seed=5, precision=prec, grid_path=True, probes=("counts",))
        z = np.concatenate([r[0][prec + "_z"], r[1][prec + "_z"]], axis=1)
        assert np.array_equal(z, g["z"]), prec
        for k in ("theta", "pi", "alpha", "counts"):
            assert np.array_equal(r[0][prec + "_" + k], g[k]) and np.array_equal(r[1][prec + "_" + k], g[k]), (prec, k)
    from bmm_mcmc_b200 import api, _lib
    ip = np.full((1, K), 1.0 / K); ith = np.ascontiguousarray(rng.uniform(0.2, 0.8, (1, P, K)))
    plan = api.Plan(_lib.SAMPLER_STICKBREAKING, X, 8, K, 1.0, 0.5, 0.5, 1.0, 1.0, 1, False, 0, seed=5, precision="fp32",
                    grid_path=True, init_pi=ip, init_theta=ith)
    plan.run()
    g = plan.fetch()
    plan.close()
    assert np.array_equal(np.concatenate([r[0]["plan_z"], r[1]["plan_z"]], axis=1), g["z"][0])
    for k in ("theta", "pi"):
        assert np.array_equal(r[0]["plan_" + k], g[k][0]) and np.array_equal(r[1]["plan_" + k], g[k][0]), k
    # relabelling on the sharded chain: the sweeps are bit-identical; the permutations come from all-reduced
    # costs whose float partial sums are grouped differently, so they agree wherever the optimum is not a
    # near-tie -- on this well-separated data everywhere -- and both ranks must hold the same ones
    g = B.gibbs_full(X, 14, K, alpha=1.0, burnin=6, relabel=True, burnrelabel=3, seed=9, precision="fp32", grid_path=True)
    z0 = np.concatenate([r[0]["rel_z_original"], r[1]["rel_z_original"]], axis=1)
    assert np.array_equal(z0, g["z_original"])
    assert np.array_equal(r[0]["rel_permutations"], r[1]["rel_permutations"])
    assert np.array_equal(r[0]["rel_permutations"], g["permutations"])
    assert np.array_equal(np.concatenate([r[0]["rel_z"], r[1]["rel_z"]], axis=1), g["z"])
    assert np.array_equal(r[0]["rel_theta"], g["theta"])
