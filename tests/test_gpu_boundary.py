"""Boundary contract of include/gpjax_b200.h on real hardware: stale-forward protection of the reused N x N buffer, refusal to
return silent zero gradients for the data, per-device launch state (two GPUs driven from ONE process), and the native NCCL
exchange step (gpb_allreduce_f64) of the row-sharded sparse path: sharded == unsharded."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a, device="cuda"):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2, 2, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return X, y


def test_backward_of_an_evicted_forward_is_replayed_not_mixed():
    """fwd A (N=100), fwd B (N=200: evicts A's buffer), fwd C (N=100: a NEW state object), backward A.  A per-state counter
    restarting at 0 would make A's token equal C's and combine A's alpha with C's factor; the process-wide token cannot."""
    from gpjax_b200 import ops

    def make(n, seed, ellv):
        X, y = data(n, 2, seed)
        p = [dev(np.array([ellv, ellv + 0.2])).requires_grad_(True), dev(1.1).requires_grad_(True), dev(0.3).requires_grad_(True)]
        return X, y, p

    XA, yA, pA = make(100, 1, 0.9)
    vA = ops.conjugate_mll_fused(0, dev(XA), dev(yA), pA[0], pA[1], pA[2], None, 1e-6)
    XB, yB, pB = make(200, 2, 1.0)
    ops.conjugate_mll_fused(0, dev(XB), dev(yB), pB[0], pB[1], pB[2], None, 1e-6)
    XC, yC, pC = make(100, 3, 1.4)
    ops.conjugate_mll_fused(0, dev(XC), dev(yC), pC[0], pC[1], pC[2], None, 1e-6)
    vA.backward()
    _, gref = o.conjugate_mll_value_and_grad_autodiff("rbf", XA, yA, np.array([0.9, 1.1]), 1.1, 0.3, 0.0)
    assert np.max(np.abs(pA[0].grad.cpu().numpy() - gref["lengthscale"])) <= 1e-8 * np.max(np.abs(gref["lengthscale"]))
    assert abs(pA[1].grad.item() - gref["variance"]) <= 1e-8 * abs(gref["variance"])
    assert abs(pA[2].grad.item() - gref["obs_stddev"]) <= 1e-8 * abs(gref["obs_stddev"])


def test_fused_objectives_refuse_data_gradients():
    from gpjax_b200 import ops
    from gpjax_b200.sgpr_ops import collapsed_elbo_fused

    X, y = data(64, 2, 5)
    Xg = dev(X).requires_grad_(True)
    v = ops.conjugate_mll_fused(0, Xg, dev(y), dev(np.array([1.0, 1.0])), dev(1.0), dev(0.3), None, 1e-6)
    with pytest.raises(NotImplementedError, match="data"):
        v.backward()
    yg = dev(y).requires_grad_(True)
    Z = dev(X[:8].copy()).requires_grad_(True)
    e = collapsed_elbo_fused(0, dev(X), yg, Z, dev(np.array([1.0, 1.0])), dev(1.0), dev(0.3), None, 1e-6, 32, None, "whitened")
    with pytest.raises(NotImplementedError, match="data"):
        e.backward()


def test_tensor_on_a_non_current_device_is_refused():
    from gpjax_b200 import ops

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    X = torch.zeros((8, 2), dtype=torch.float64, device="cuda:1")
    with torch.cuda.device(0), pytest.raises(RuntimeError, match="current device"):
        ops.gram_forward(0, X, X, torch.ones(2, dtype=torch.float64, device="cuda:1"), torch.ones((), dtype=torch.float64, device="cuda:1"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_driven_from_one_process():
    """XLA's default: one process, one executor thread per device.  Kernel attributes (dynamic shared memory opt-in), look-ahead
    streams and events are per device in the library, so the same calls work on cuda:1 after cuda:0 -- interleaved and from two
    host threads at once -- with results identical to the single-device run (N = 4500: int8 path, look-ahead, 129 KB leaves)."""
    import threading

    from gpjax_b200 import ops

    n = 4500
    X, y = data(n, 8, 77)
    ell = np.linspace(0.8, 1.6, 8)
    out, err = {}, []

    def run(idx, tag):
        try:
            with torch.cuda.device(idx):
                d = f"cuda:{idx}"
                p = [dev(ell, d).requires_grad_(True), dev(1.0, d).requires_grad_(True), dev(0.3, d).requires_grad_(True)]
                v = ops.conjugate_mll_fused(2, dev(X, d), dev(y, d), p[0], p[1], p[2], None, 1e-6)
                v.backward()
                torch.cuda.synchronize()
                out[tag] = (v.item(), p[0].grad.cpu().numpy().copy())
        except Exception as e:  # pragma: no cover
            err.append((tag, repr(e)))

    run(0, "d0")
    run(1, "d1")
    th = [threading.Thread(target=run, args=(i, f"t{i}")) for i in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    for tag in ("d1", "t0", "t1"):
        assert out[tag][0] == out["d0"][0], tag  # deterministic kernels: bit-identical value on either device
        assert np.array_equal(out[tag][1], out["d0"][1]), tag
    ops.release_buffers()


WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from gpjax_b200 import sgpr_ops
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
native = sgpr_ops.init_native_collective() if {native} else False
rng = np.random.default_rng(11)
N, M, D = 300_000, 1024, 8
X = rng.uniform(-2, 2, (N, D)); y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
Zn = rng.uniform(-2, 2, (M, D))
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
def run(Xs, ys, group_on):
    p = [dev(Zn).requires_grad_(True), dev(np.linspace(0.8, 1.6, D)).requires_grad_(True), dev(1.0).requires_grad_(True),
         dev(0.3).requires_grad_(True), dev(0.1).requires_grad_(True)]
    if not group_on:  # unsharded reference on this rank alone: hide the process group from the op
        w = sgpr_ops._world; sgpr_ops._world = lambda g: 1
    try:
        v = sgpr_ops.collapsed_elbo_fused(0, dev(Xs), dev(ys), p[0], p[1], p[2], p[3], p[4], 1e-6, 65536, None, "raw")
        v.backward()
    finally:
        if not group_on: sgpr_ops._world = w
    return v.item(), torch.cat([q.grad.reshape(-1) for q in p])
lo, hi = rank * N // world, (rank + 1) * N // world
vs, gs = run(X[lo:hi], y[lo:hi], True)
vf, gf = run(X, y, False)
res = dict(native=bool(native), value_rel=abs(vs - vf) / abs(vf), grad_rel=float((gs - gf).abs().max() / gf.abs().max()), value=vs)
vals = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(vals, torch.tensor([vs], dtype=torch.float64, device="cuda"))
res["identical_across_ranks"] = bool(all(float(v) == float(vals[0]) for v in vals))
open(os.path.join({out!r}, f"nccl_result_{{rank}}.json"), "w").write(json.dumps(res))
sgpr_ops.destroy_native_collectives()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("native", [True, False], ids=["gpb_allreduce_f64", "torch.distributed"])
def test_row_sharded_elbo_over_nccl_equals_unsharded(tmp_path, native):
    """2 ranks x 150,000 rows (3 blocks of 65,536 per rank: int8 statistics SYRK + pass 2) against the same evaluation on all
    300,000 rows on one GPU: value and every gradient within 1e-11 (only the summation order of the block statistics differs)."""
    import json
    import socket

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path), native=native))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for rank in range(2):
        e = json.loads((tmp_path / f"nccl_result_{rank}.json").read_text())
        assert e["native"] == native and e["identical_across_ranks"]
        assert e["value_rel"] <= 1e-11 and e["grad_rel"] <= 1e-11, e
