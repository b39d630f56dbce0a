"""Parity of the INT8-tensor-pipe (Ozaki scheme) path: the raw tcgen05 kind::i8 product is bit-exact against integer
matmul, the digit planes reconstruct the operand, the recombined fp64 product meets its truncation bound, and the
blocked Cholesky / conjugate_mll built on it agree with the FP64 DMMA path and the oracle."""
import numpy as np
import pytest
import torch

import oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_switch():
    from gpjax_b200 import ops

    before = ops.get_ozaki_slices()
    yield
    ops.set_ozaki_slices(before)


def test_driver_can_encode_tensor_maps():
    from gpjax_b200 import ops

    assert ops.ozaki_available()


@pytest.mark.parametrize("m,n,k,pad", [(128, 128, 128, 0), (1, 1, 128, 0), (300, 200, 256, 128), (1000, 520, 1024, 0),
                                       (4096, 4100, 2048, 1024), (129, 4500, 384, 0)])
def test_raw_int8_product_is_bit_exact(m, n, k, pad):
    from gpjax_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(m + 3 * n + 7 * k)
    Afull = torch.randint(-128, 128, (m, k + pad), dtype=torch.int8, device="cuda", generator=g)
    Bfull = torch.randint(-128, 128, (n, k + pad), dtype=torch.int8, device="cuda", generator=g)
    A, B = Afull[:, pad:], Bfull[:, :k]  # strided views: lda != k, non-zero column offset
    C = ops.igemm_i8(A, B)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T).to(torch.int64)  # exact: |sum| < 2^31 << 2^53
    assert torch.equal(C.to(torch.int64), ref)


@pytest.mark.parametrize("s", [4, 6, 7])
def test_digit_planes_reconstruct_the_operand(s):
    from gpjax_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(s)
    X = torch.randn(300, 256, dtype=torch.float64, device="cuda", generator=g)
    X *= torch.exp(3 * torch.randn(300, 1, dtype=torch.float64, device="cuda", generator=g))
    X[5] = 0.0
    X[7, 3] = 1.0  # exact power of two as the row maximum
    X[7, 4:] *= 1e-3
    # maxima just below a power of two (both signs): the carry into the top digit must stay inside int8
    X[9, :4] = torch.tensor([np.nextafter(2.0, 0), -np.nextafter(2.0, 0), 1.9765, -1.9765], dtype=torch.float64)
    X[9, 4:] *= 1e-2 / X[9, 4:].abs().max()
    X[11] = -np.nextafter(4.0, 0)
    Q, sc = ops.ozaki_slice(X, s)
    torch.cuda.synchronize()
    assert int(Q.min()) >= -128 and int(Q.max()) <= 127 and int(Q.to(torch.int32).abs().max()) > 64  # balanced radix-256 digits
    planes = Q.view(300, s, 256).double()
    w = torch.tensor([2.0 ** (-8 * (p + 1)) for p in range(s)], dtype=torch.float64, device="cuda")
    rec = (planes * w[None, :, None]).sum(1) * sc[:, None]
    err = (rec - X).abs() / sc[:, None]
    assert float(err.max()) <= 2.0 ** (-8 * s - 1)  # ONE rounding, to the last plane
    assert float(sc[5]) == 1.0 and int(Q[5].abs().max()) == 0
    assert torch.all(X.abs().amax(1)[sc > 0] < 0.5 * sc[sc > 0] + 1e-300)


def test_nan_row_poisons_its_scale_only():
    from gpjax_b200 import ops

    X = torch.ones(4, 128, dtype=torch.float64, device="cuda")
    X[2, 17] = float("nan")
    Q, sc = ops.ozaki_slice(X, 6)
    assert torch.isnan(sc[2]) and torch.isfinite(sc[[0, 1, 3]]).all()


@pytest.mark.parametrize("s,tol", [(4, 4e-7), (5, 2e-9), (6, 8e-12), (7, 4e-14)])
@pytest.mark.parametrize("m,n,k,lower", [(700, 300, 1024, False), (2500, 2500, 1024, True), (130, 260, 128, False)])
def test_recombined_product_meets_truncation_bound(s, tol, m, n, k, lower):
    from gpjax_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(m + n + k + s)
    A = torch.randn(m, k, dtype=torch.float64, device="cuda", generator=g)
    A *= torch.exp(torch.randn(m, 1, dtype=torch.float64, device="cuda", generator=g))
    B = A[:n] if lower else torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    Qa, sa = ops.ozaki_slice(A, s)
    Qb, sb = (Qa[:n], sa[:n]) if lower else ops.ozaki_slice(B, s)
    C = C0.clone()
    ops.ozaki_gemm_(C, Qa, sa, Qb, sb, k, s, alpha=-1.0, mask_lower=lower)
    torch.cuda.synchronize()
    ref = C0 - A @ B.T
    if lower:
        keep = torch.ones(m, n, dtype=torch.bool, device="cuda").tril()
        assert torch.equal(C[~keep], C0[~keep])  # masked entries untouched
        C, ref = C[keep], ref[keep]
        bound = (A.abs().amax(1)[:, None] * B.abs().amax(1)[None, :])[keep]
    else:
        bound = A.abs().amax(1)[:, None] * B.abs().amax(1)[None, :]
    # dropped orders: (s+1) 2^(-8 s - 2) k 2^(ea+eb), 2^e <= 8 (row max) at worst; measured far below -- the test pins the order
    # of magnitude
    assert float(((C - ref).abs() / (k * bound)).max()) <= tol / 16
    assert float(((C - ref).abs() / (A.norm(dim=1).max() * B.norm(dim=1).max())).max()) <= tol


def test_write_out_paths_agree_power_of_two_scales_generic_scales_and_poisoned_rows():
    """The default write-out builds the fp64 result with integer instructions (power-of-two scales = an exponent add).  Scales
    that are not powers of two (possible through the C ABI) take the FP64 write-out; a NaN scale poisons its row / column only."""
    from gpjax_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    m, n, k, s = 600, 500, 256, 6
    A = torch.randn(m, k, dtype=torch.float64, device="cuda", generator=g) * torch.exp(
        4 * torch.randn(m, 1, dtype=torch.float64, device="cuda", generator=g))
    B = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
    Qa, sa = ops.ozaki_slice(A, s)
    Qb, sb = ops.ozaki_slice(B, s)
    C0 = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    C1 = C0.clone()
    ops.ozaki_gemm_(C1, Qa, sa, Qb, sb, k, s, alpha=-1.0)
    ref = C0 - A @ B.T
    scale = A.abs().amax(1)[:, None] * B.abs().amax(1)[None, :] * k
    assert float(((C1 - ref).abs() / scale).max()) <= 1e-12
    # generic scales: 3 sa and sb / 3 give the same product up to the roundings of the FP64 write-out
    C2 = C0.clone()
    ops.ozaki_gemm_(C2, Qa, 3.0 * sa, Qb, sb / 3.0, k, s, alpha=-1.0)
    assert float(((C2 - C1).abs() / scale).max()) <= 1e-15
    # alpha that is not a power of two
    C3 = C0.clone()
    ops.ozaki_gemm_(C3, Qa, sa, Qb, sb, k, s, alpha=-0.3)
    assert float(((C3 - (C0 - 0.3 * (A @ B.T))).abs() / scale).max()) <= 1e-12
    # poisoned row and column
    sa2, sb2 = sa.clone(), sb.clone()
    sa2[17] = float("nan")
    sb2[401] = float("nan")
    C4 = C0.clone()
    ops.ozaki_gemm_(C4, Qa, sa2, Qb, sb2, k, s, alpha=-1.0)
    bad = torch.isnan(C4)
    expect = torch.zeros_like(bad)
    expect[17, :] = True
    expect[:, 401] = True
    assert torch.equal(bad, expect)
    assert torch.equal(C4[~bad], C1[~bad])


@pytest.mark.parametrize("n", [4096, 5000])
def test_cholesky_on_the_int8_pipe_matches_the_dmma_factor(n):
    from gpjax_b200 import ops

    rng = np.random.default_rng(n)
    X = rng.uniform(-2, 2, (n, 8))
    ell = np.linspace(0.8, 1.6, 8)
    Xd = torch.as_tensor(X, device="cuda")
    S = ops.gram_forward(0, Xd, Xd, torch.as_tensor(ell, device="cuda"), torch.tensor(1.0, dtype=torch.float64, device="cuda"),
                         diag_add=1e-6 + 0.09)
    ws = ops.FactorWorkspace(n, 8, potri=False, device="cuda")
    ops.set_ozaki_slices(0)
    L0 = S.clone()
    assert int(ops.potrf_lower_(L0, ws)) == 0
    ops.set_ozaki_slices(6)
    assert ops.get_ozaki_slices() == 6
    L7 = S.clone()
    assert int(ops.potrf_lower_(L7, ws)) == 0
    torch.cuda.synchronize()
    assert not torch.equal(L0, L7), "the switch did not change the arithmetic: int8 path not taken"
    # 6 radix-256 planes resolve 48 bits below the row maximum (round 1 certified 49 bits here at 1e-12); 7 planes (56 bits, what
    # the guard gives a bare matrix) must sit at the FP64 path's own level
    assert float((L0 - L7).abs().max() / L0.abs().max()) <= 3e-12
    rec = L7 @ L7.T
    assert float((rec - S).abs().max()) <= 1e-12 * float(S.abs().max()) * 8
    ops.set_ozaki_slices(7)
    L8 = S.clone()
    assert int(ops.potrf_lower_(L8, ws)) == 0
    assert float((L0 - L8).abs().max() / L0.abs().max()) <= 1e-13
    assert float((L8 @ L8.T - S).abs().max()) <= 2e-14 * float(S.abs().max()) * 8
    ops.set_ozaki_slices(4)
    L5 = S.clone()
    ops.potrf_lower_(L5, ws)
    e5 = float((L0 - L5).abs().max() / L0.abs().max())
    assert 1e-12 < e5 < 1e-6  # four planes (32 bits) are visibly coarser: the plane count really reaches the kernel


def test_conjugate_mll_value_and_gradient_on_the_int8_pipe_vs_oracle():
    from gpjax_b200 import ops

    n, d = 4096, 8
    rng = np.random.default_rng(41)
    X = rng.uniform(-2, 2, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    ell = np.linspace(0.8, 1.6, d)
    dev = lambda a: torch.as_tensor(np.asarray(a, np.float64), device="cuda")
    vref = o.conjugate_mll("rbf", X, y, ell, 1.0, 0.3, 0.0)
    gref = o.conjugate_mll_grad_closed_form("rbf", X, y, ell, 1.0, 0.3, 0.0)
    ops.set_ozaki_slices(7)
    p = [dev(ell).requires_grad_(True), dev(1.0).requires_grad_(True), dev(0.3).requires_grad_(True),
         dev(0.0).requires_grad_(True)]
    val = ops.conjugate_mll_fused(0, dev(X), dev(y), p[0], p[1], p[2], p[3], 1e-6)
    val.backward()
    assert abs(val.item() - vref) <= 1e-8 * abs(vref)
    assert np.max(np.abs(p[0].grad.cpu().numpy() - gref["lengthscale"])) <= 1e-8 * np.max(np.abs(gref["lengthscale"]))
    assert abs(p[1].grad.item() - gref["variance"]) <= 1e-8 * abs(gref["variance"])
    assert abs(p[2].grad.item() - gref["obs_stddev"]) <= 1e-8 * abs(gref["obs_stddev"])


@pytest.mark.parametrize("env", [{"GPB_OZ_KERNEL": "1"}, {"GPB_OZ_KERNEL": "2"}, {"GPB_OZ_KERNEL": "3"}, {"GPB_OZ_KERNEL": "3", "GPB_OZ_PAIR": "0"},
                                 {"GPB_OZ_KERNEL": "1", "GPB_OZ_PAIR": "0"}, {"GPB_OZ_KERNEL": "1", "GPB_OZ_PAIR": "0", "GPB_OZ_KBS1": "1"}])
def test_alternative_kernel_schedules_agree(env):
    """The measured-but-not-default kernels (variant 1: one CTA per 128 x 128 tile, the round-1 default; variant 2:
    plane-resident; variant 3: CTA pairs with N = 128 instructions; unpaired orders; one K-block per stage) are selected by
    environment variables read at first launch, so each runs in its own process.  The default is variant 4 (CTA pairs, order
    pairs side by side in N = 256 instructions), which every other test exercises."""
    import os
    import subprocess
    import sys

    code = (
        "import torch; from gpjax_b200 import ops\n"
        "g = torch.Generator(device='cuda').manual_seed(5)\n"
        "A = torch.randn(1500, 1024, dtype=torch.float64, device='cuda', generator=g)\n"
        "C0 = torch.randn(1500, 1500, dtype=torch.float64, device='cuda', generator=g)\n"
        "for s, tol in ((5, 2e-9), (6, 8e-12), (7, 4e-14)):\n"
        "    Q, sc = ops.ozaki_slice(A, s); C = C0.clone()\n"
        "    ops.ozaki_gemm_(C, Q, sc, Q, sc, 1024, s, alpha=-1.0, mask_lower=True)\n"
        "    ref = torch.where(torch.ones_like(C0, dtype=torch.bool).tril(), C0 - A @ A.T, C0)\n"
        "    err = float(((C - ref).abs() / (A.norm(dim=1).max() ** 2)).max())\n"
        "    assert err <= tol, (s, err)\n"
        "I = torch.randint(-128, 128, (300, 512), dtype=torch.int8, device='cuda', generator=g)\n"
        "assert torch.equal(ops.igemm_i8(I, I[:200]).long(), (I.double() @ I[:200].double().T).long())\n"
        "print('ok')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
