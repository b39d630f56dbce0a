"""World-size-2 check of the row-sharded SGPR protocol with a real collective (gloo, CPU).

Each process owns half of the rows and drives the product's orchestration + C ABI (host model of the
device primitives, see tests/hostsim) through the six protocol steps of include/gpjax_b200.h; the two
all-reduces go through torch.distributed exactly as gpjax_b200/sgpr_ops.py issues them on NCCL.
Both ranks must end with the full-data ELBO and gradient of the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import ctypes as C, os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from gpjax_b200 import _abi
import oracle as o

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = _abi.declare(C.CDLL(os.path.join({here!r}, "hostsim", "libgpjax_b200_hostsim.so")))
p = lambda a: None if a is None else a.ctypes.data
rng = np.random.default_rng(7)
N, M, D, block = 420, 36, 3, 100
X = rng.uniform(-2, 2, (N, D)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
Z = np.ascontiguousarray(rng.uniform(-2, 2, (M, D)))
ell, var, sn, c = np.linspace(0.8, 1.4, D), np.array([1.3]), np.array([0.4]), np.array([0.2])
lo, hi = rank * N // world, (rank + 1) * N // world
Xr, yr = np.ascontiguousarray(X[lo:hi]), np.ascontiguousarray(y[lo:hi])
nbytes = lib.gpb_sgpr_workspace_bytes(M, D, block)
ws = np.zeros(nbytes // 8 + 8)
P = np.zeros(lib.gpb_sgpr_stats_count(M))
assert lib.gpb_sgpr_stats(None, 2, hi - lo, M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c), 1e-6, block,
                          p(ws), nbytes, p(P)) == 0
Pt = torch.from_numpy(P); dist.all_reduce(Pt)                      # exchange step 1
val, info = np.zeros(1), np.zeros(2, np.int32)
assert lib.gpb_sgpr_finish(None, 2, M, D, p(Z), D, p(ell), 0, p(var), p(sn), block, p(ws), nbytes, p(P), 1, p(val),
                           p(info)) == 0
flat = np.zeros(M * D + D + 1)
gZ, gl, gv = flat[:M * D], flat[M * D:M * D + D], flat[M * D + D:]
assert lib.gpb_sgpr_grad_local(None, 2, hi - lo, M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c), block,
                               p(ws), nbytes, p(gZ), p(gl), p(gv)) == 0
ft = torch.from_numpy(flat); dist.all_reduce(ft)                   # exchange step 2
gs, gc = np.zeros(1), np.zeros(1)
assert lib.gpb_sgpr_grad_finish(None, 2, M, D, p(Z), D, p(ell), 0, p(var), p(sn), block, p(ws), nbytes, None, p(gZ),
                                p(gl), p(gv), p(gs), p(gc)) == 0
ref, g = o.collapsed_elbo_value_and_grad_autodiff("matern52", X, y, Z, ell, var[0], sn[0], c[0])
err = dict(value=abs(val[0] - ref) / abs(ref),
           Z=float(np.max(np.abs(gZ.reshape(M, D) - g["inducing_inputs"])) / np.max(np.abs(g["inducing_inputs"]))),
           ell=float(np.max(np.abs(gl - g["lengthscale"])) / np.max(np.abs(g["lengthscale"]))),
           var=abs(gv[0] - g["variance"]) / abs(g["variance"]), sn=abs(gs[0] - g["obs_stddev"]) / abs(g["obs_stddev"]),
           c=abs(gc[0] - g["mean_const"]) / abs(g["mean_const"]), n=float(P[(M + 1) * (M + 2) + M + 1]))
vals = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(vals, torch.tensor([val[0]], dtype=torch.float64))
err["identical_across_ranks"] = bool(all(float(v) == float(vals[0]) for v in vals))
open(os.path.join({out!r}, f"result_{{rank}}.json"), "w").write(json.dumps(err))
dist.destroy_process_group()
'''


SVGP_WORKER = r'''
import ctypes as C, os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from gpjax_b200 import _abi
import oracle as o

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = _abi.declare(C.CDLL(os.path.join({here!r}, "hostsim", "libgpjax_b200_hostsim.so")))
p = lambda a: None if a is None else a.ctypes.data
rng = np.random.default_rng(11)                                    # the same stream on every rank: replicated parameters
N, M, D, block, B = 900, 24, 3, 64, 150                            # B rows per rank and step, drawn from the rank's own shard
X = rng.uniform(-2, 2, (N, D)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
Z = np.ascontiguousarray(rng.uniform(-2, 2, (M, D)))
mu = rng.standard_normal(M) * 0.3
W = np.ascontiguousarray(np.tril(rng.standard_normal((M, M)) * 0.1) + 0.7 * np.eye(M))
ell, var, sn, c = np.linspace(0.8, 1.4, D), np.array([1.3]), np.array([0.4]), np.array([0.2])
ndata, jitter = float(N), 1e-6
lo, hi = rank * N // world, (rank + 1) * N // world
idx = [lo_ + np.random.default_rng(100 + r).integers(0, hi_ - lo_, B)   # every rank can rebuild every rank's minibatch
       for r, (lo_, hi_) in enumerate((q * N // world, (q + 1) * N // world) for q in range(world))]
Xr, yr = np.ascontiguousarray(X[idx[rank]]), np.ascontiguousarray(y[idx[rank]])
nbytes = lib.gpb_sgpr_workspace_bytes(M, D, block)
ws = np.zeros(nbytes // 8 + 8)
P = np.zeros(lib.gpb_sgpr_stats_count(M))
assert lib.gpb_sgpr_stats(None, 1, B, M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c), jitter, block,
                          p(ws), nbytes, p(P)) == 0
Pt = torch.from_numpy(P); dist.all_reduce(Pt)                      # exchange step 1: statistics of the union batch
val, info = np.zeros(1), np.zeros(2, np.int32)
assert lib.gpb_svgp_finish(None, 1, M, D, p(Z), D, p(ell), 0, p(var), p(sn), p(c), p(mu), p(W), M, ndata, jitter, block,
                           p(ws), nbytes, p(P), 1, p(val), p(info)) == 0
flat = np.zeros(M * D + D + 1)
gZ, gl, gv = flat[:M * D], flat[M * D:M * D + D], flat[M * D + D:]
assert lib.gpb_sgpr_grad_local(None, 1, B, M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c), block,
                               p(ws), nbytes, p(gZ), p(gl), p(gv)) == 0
ft = torch.from_numpy(flat); dist.all_reduce(ft)                   # exchange step 2: data part of the gradient
gs, gc, gmu, gW = np.zeros(1), np.zeros(1), np.zeros(M), np.full((M, M), np.nan)
assert lib.gpb_svgp_grad_finish(None, 1, M, D, p(Z), D, p(ell), 0, p(var), p(sn), jitter, block, p(ws), nbytes, None,
                                p(W), M, p(gZ), p(gl), p(gv), p(gs), p(gc), p(gmu), p(gW), M) == 0
allidx = np.concatenate(idx)                                       # the step's effective batch: world * B rows
ref, g = o.svgp_elbo_value_and_grad_autodiff("matern32", X[allidx], y[allidx], Z, ell, var[0], sn[0], c[0], mu, W, ndata, jitter)
got = dict(lengthscale=gl, variance=gv[0], obs_stddev=gs[0], mean_const=gc[0], inducing_inputs=gZ.reshape(M, D),
           variational_mean=gmu, variational_root_covariance=gW)
err = dict(value=abs(val[0] - ref) / abs(ref))
for k in g:
    a, b = np.asarray(got[k]).reshape(np.shape(g[k])), np.asarray(g[k])
    err[k] = float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-8 * abs(ref)))
vals = [torch.zeros(1 + M, dtype=torch.float64) for _ in range(world)]
dist.all_gather(vals, torch.from_numpy(np.concatenate([val, gmu])))
err["identical_across_ranks"] = bool(all(torch.equal(v, vals[0]) for v in vals))
open(os.path.join({out!r}, f"result_{{rank}}.json"), "w").write(json.dumps(err))
dist.destroy_process_group()
'''


def _run_world2(tmp_path, worker):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "hostsim")], check=True)
    script = tmp_path / "worker.py"
    script.write_text(worker.format(root=ROOT, here=HERE, out=str(tmp_path)))
    import socket

    with socket.socket() as sk:  # ask the OS for a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    import json

    return [json.loads((tmp_path / f"result_{rank}.json").read_text()) for rank in range(2)]


def test_svgp_data_parallel_world2_gloo(tmp_path):
    """BASELINE config 5's multi-GPU form (SVGP row): every rank draws its own minibatch from its shard, the two
    all-reduces of the SGPR protocol make the step's effective batch the union, and the replicated finish must hand every rank the
    same ELBO and gradient -- those of the oracle's elbo (objectives.py:241-315) on the union batch."""
    for e in _run_world2(tmp_path, SVGP_WORKER):
        assert e.pop("identical_across_ranks")
        assert e.pop("value") <= 1e-9
        for k, v in e.items():
            assert v <= 1e-7, (k, v)


def test_sgpr_row_sharded_world2_gloo(tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "hostsim")], check=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, here=HERE, out=str(tmp_path)))
    import socket

    with socket.socket() as sk:  # ask the OS for a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    import json

    results = [json.loads((tmp_path / f"result_{rank}.json").read_text()) for rank in range(2)]
    for e in results:
        assert e["n"] == 420.0 and e["identical_across_ranks"]
        for k in ("value", "Z", "ell", "var", "sn", "c"):
            assert e[k] <= 1e-8, (k, e[k])
