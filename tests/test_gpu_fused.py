"""The fused sub-tree kernel (csrc/fused_node.cuh: one cluster launch factors and inverts a diagonal block of up
to 512 / 1024 rows) against the launch-per-step recursion it replaces and against the oracle.  Both paths run
the same FP64 DMMA arithmetic in the same k order, so their results must agree to the last bits; the oracle
comparison is the parity gate proper (mogp fit / logposterior behind exauq/core/emulators.py:418-450)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exq():
    import exauq_toolbox_b200 as m

    m._lib.init(0)
    yield m
    m._lib.set_fused(512, 0)
    m._lib.set_int8_digits(6, 7)


@pytest.fixture(scope="module")
def orc():
    from oracle import gp_oracle

    return gp_oracle


def _run(exq, X, y, raw, thetas):
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    out = {"nll": dev.neg_log_likelihood, "loo": dev.loo()}
    out["val"], out["grad"], out["info"], _ = dev.logpost_batch(thetas, 0.0)
    dev.close()
    return out


def test_device_offers_clusters(exq):
    info = exq._lib.fused_info()
    assert info["max_rows"] == 512 and info["clusters_of_8"] >= 1


@pytest.mark.parametrize("N,d", [(200, 3), (500, 4), (600, 2), (1500, 5), (1025, 6), (4096, 8)])
def test_fused_equals_launch_per_step(exq, orc, N, d):
    """Shapes: one fused launch for everything (Np <= 512), odd splits (Np = 640: 384 + 256; 1536: 768 -> 384 + 384),
    a 1-row overhang (Np = 1152 = 9 blocks) and the benchmark size, with 1, 5 and 15 systems per launch."""
    lib = exq._lib
    X, y = orc.synthetic_problem(N, d)
    raw = np.full(d, -2 * np.log(0.5))
    rng = np.random.default_rng(N)
    R = 15 if N <= 1500 else 5
    thetas = np.concatenate([raw + rng.uniform(-0.6, 0.6, (R, d)), np.log(1.7) + rng.uniform(-0.4, 0.4, (R, 1))], axis=1)
    lib.set_fused(0, 0)
    ref = _run(exq, X, y, raw, thetas)
    for rows, G in [(512, 0), (512, 8), (256, 4), (1024, 8), (512, 1)]:
        lib.set_fused(rows, G)
        got = _run(exq, X, y, raw, thetas)
        assert np.all(got["info"] == 0)
        # a 1024-row fused block keeps the products with k = 512 on FP64 DMMA, the launch-per-step path sends them to
        # the 7-digit int8 kernel when the batch fills the chip: same values to 2^-50, not to the last bits
        tol = 1e-14 if rows <= 512 else 1e-12
        assert abs(got["nll"] - ref["nll"]) <= tol * abs(ref["nll"]), (rows, G)
        for a, b in zip(got["loo"], ref["loo"]):
            assert np.allclose(a, b, rtol=1e-12 if rows <= 512 else 1e-10, atol=1e-14), (rows, G)
        assert np.allclose(got["val"], ref["val"], rtol=tol, atol=0), (rows, G)
        assert np.allclose(got["grad"], ref["grad"], rtol=1e-10, atol=1e-12 * np.max(np.abs(ref["grad"]))), (rows, G)
    lib.set_fused(512, 0)


def test_fused_against_oracle(exq, orc):
    lib = exq._lib
    lib.set_fused(512, 0)
    N, d = 900, 4
    X, y = orc.synthetic_problem(N, d)
    raw = np.full(d, -2 * np.log(0.4))
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 1e-8)
    ref = orc.fit(X, y, "Matern52", raw, 1.7, 1e-8)
    onll = orc.neg_log_likelihood(ref)
    assert abs(dev.neg_log_likelihood - onll) <= 1e-9 * abs(onll)
    Xc = orc.synthetic_candidates(1000, d)
    m, v = dev.predict(Xc)
    mo, vo = orc.predict(ref, Xc)
    assert np.all(np.abs(m - mo) <= np.maximum(1e-8 * np.abs(mo), 1e-9))
    assert np.all(np.abs(v - vo) <= np.maximum(1e-8 * np.abs(vo), 1e-9))
    th = np.concatenate([raw, [np.log(1.7)]])
    val, grad, info, _ = dev.logpost_batch(th[None, :], 1e-8)
    oval, ograd = orc.neg_log_likelihood_grad(X, y, "Matern52", th, 1e-8)
    assert info[0] == 0 and abs(val[0] - oval) <= 1e-9 * abs(oval)
    assert np.all(np.abs(grad[0] - ograd) <= 1e-7 * np.max(np.abs(ograd)))
    dev.close()


@pytest.mark.parametrize("kernel,N,d,M,ell,nug", [("Matern52", 1500, 4, 5000, 0.45, 1e-9), ("SquaredExponential", 700, 3, 3001, 0.15, 1e-6),
                                                  ("ProductMat52", 300, 2, 777, 0.45, 1e-9), ("Matern52", 4096, 8, 40000, 0.45, 1e-9)])
def test_fused_pei_pipeline_equals_three_pass_pipeline_and_oracle(exq, orc, kernel, N, d, M, ell, nug):
    """csrc/pei_fused.cuh (assembly -> digit planes + per-tile mean / repulsion partials, no FP64 covariance chunk)
    against round 1's three passes on the same inputs -- identical digits, so identical variances; means and PEI
    differ only by summation order -- and both against the oracle's PEI (exauq/core/designers.py:522-706)."""
    lib = exq._lib
    X, y = orc.synthetic_problem(N, d)
    Xc = orc.synthetic_candidates(M, d)
    raw = np.full(d, -2 * np.log(ell))
    rep = np.random.default_rng(1).uniform(0, 1, (2 * d + 2 ** d, d))
    dev = exq.DeviceGP(X, y, kernel).fit(raw, 1.7, nug)
    out = {}
    for mode in (False, True):
        lib.set_pei_fused(mode)
        m, v = dev.predict(Xc)
        cand = exq.DeviceCandidates(Xc)
        cand.pei_eval(dev, rep, float(y.max()))
        idx, val = cand.batch(dev, 5)
        out[mode] = (m, v, cand.download("pei"), idx, val)
        cand.close()
    lib.set_pei_fused(True)
    # same digits -> same partial sums of squares; only the order in which the Np/32 partials are added differs
    assert np.allclose(out[True][1], out[False][1], rtol=0, atol=1e-13 * 1.7)
    # means: sums of N terms k_cj alpha_j in a different order -> a few ulp of sum_j |k_cj alpha_j| <= 1.7 |alpha|_1
    K = 1.7 * lib.corr(kernel, raw, X, X) + nug * np.eye(N)
    a1 = float(np.abs(np.linalg.solve(K, y)).sum()) * 1.7
    assert np.all(np.abs(out[True][0] - out[False][0]) <= 1e-14 * a1)
    pmax = float(out[False][2].max())
    assert np.allclose(out[True][2], out[False][2], rtol=1e-6, atol=1e-12 * pmax)   # EI amplifies the mean's last bits
    assert out[True][3].tolist() == out[False][3].tolist()
    if N <= 1500:
        ref = orc.fit(X, y, kernel, raw, 1.7, nug)
        mo, vo = orc.predict(ref, Xc)
        assert np.all(np.abs(out[True][0] - mo) <= np.maximum(1e-8 * np.abs(mo), 1e-9))
        assert np.all(np.abs(out[True][1] - vo) <= np.maximum(1e-8 * np.abs(vo), 1e-9))
        oi, ov, _ = orc.pei_batch_argmax(ref, kernel, Xc, rep, float(y.max()), 5)
        assert out[True][3].tolist() == oi.tolist()
        assert np.allclose(out[True][4], ov, rtol=1e-6)
    dev.close()


@pytest.mark.parametrize("kernel,N,d,R,fit_nugget", [("Matern52", 4096, 8, 15, False), ("SquaredExponential", 2300, 5, 6, False),
                                                     ("Matern52", 2048, 3, 4, True)])
def test_gradient_in_lauum_epilogue_equals_separate_kernel(exq, kernel, N, d, R, fit_nugget):
    """The int8 LAUUM kernel's C_GRAD epilogue (csrc/ozaki_gemm.cuh: K^-1 tiles contracted with dK/dtheta on the spot,
    never stored) against LAUUM + grad_fast_kernel on the same digits: the K^-1 entries are the same numbers, only the
    order of the FP64 summation differs.  The oracle comparison of the default (fused) path is
    tests/test_gpu_illcond.py / test_gpu_parity.py (mogp logpost_deriv, exauq/utilities/mogp_fitting.py:318-324)."""
    lib = exq._lib
    from oracle import gp_oracle as orc

    X, y = orc.synthetic_problem(N, d)
    rng = np.random.default_rng(N + d)
    raw = np.full(d, -2 * np.log(0.45 if kernel == "Matern52" else 0.2))
    cols = [raw + rng.uniform(-0.5, 0.5, (R, d)), np.log(1.7) + rng.uniform(-0.4, 0.4, (R, 1))]
    if fit_nugget:
        cols.append(np.log(1e-6) + rng.uniform(-1, 1, (R, 1)))
    thetas = np.concatenate(cols, axis=1)
    dev = exq.DeviceGP(X, y, kernel)
    out = {}
    for mode in (False, True):
        lib.set_grad_fused(mode)
        n0 = lib.launch_count()
        out[mode] = dev.logpost_batch(thetas, 1e-8, fit_nugget=fit_nugget)
        out[mode] += (lib.launch_count() - n0,)
    lib.set_grad_fused(True)
    assert np.all(out[True][2] == 0) and np.all(out[False][2] == 0)
    assert np.array_equal(out[True][0], out[False][0])                      # values do not involve the gradient path
    gmax = np.max(np.abs(out[False][1]), axis=1, keepdims=True)
    assert np.all(np.abs(out[True][1] - out[False][1]) <= 1e-11 * gmax)
    assert out[True][4] < out[False][4]                                     # one kernel less per evaluation round
    dev.close()
