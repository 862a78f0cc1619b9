"""categorical_logit_lpmf with one row of log odds per outcome.

CPU part: the oracle restatement (oracle/glm_oracle.c: oracle_categorical_logit_lpmf)
against (a) tests/golden/categorical_lpmf_golden.json, written by the unmodified
reference, (b) the reference live when oracle/_ref is present, (c) the pinned
categorical GLM oracle with beta = I (x beta = lin, d_x = d_lin), (d) numpy.
GPU part: smc_categorical_logit_lpmf through the C-ABI against the oracle and the
golden fixture, the error contract, determinism, and -- at 2e6 x 32 -- row-block
additivity plus the sum-to-zero property of every row of the partial.

Tolerances (BASELINE.json): rel 1e-10 on logp, 1e-9 on gradients."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "categorical_lpmf_golden.json")


def golden_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


CASES = golden_cases()
IDS = [f'{c["case"]}-{"propto" if c["propto"] else "full"}' for c in CASES]


def unpack(case):
    N, C = case["shape"]
    lin = np.asarray(case["lin"], dtype=np.float64).reshape((N, C), order="F")
    d = np.asarray(case["d_lin"], dtype=np.float64).reshape((N, C), order="F")
    return np.asarray(case["y"], dtype=np.int32), lin, case["logp"], d


def flags_of(case):
    return po.VAR_ALPHA | (po.PROPTO if case["propto"] else 0)


# ------------------------------------------------------------------ CPU: the oracle
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_matches_golden(case):
    y, lin, logp, d = unpack(case)
    o = po.categorical_logit_lpmf(y, lin, flags_of(case))
    assert o["rc"] == 0
    assert_logp(o["logp"], logp)
    assert_grad(o["d_lin"].ravel(), d.ravel(), "d_lin")


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built here")
def test_oracle_matches_reference_live():
    rng = np.random.default_rng(3)
    lin = rng.standard_normal((211, 17)) * 6.0
    y = rng.integers(1, 18, 211).astype(np.int32)
    o = po.categorical_logit_lpmf(y, lin)
    r = po.categorical_logit_lpmf(y, lin, impl="ref")
    assert r["rc"] == 0
    assert_logp(o["logp"], r["logp"])
    assert_grad(o["d_lin"].ravel(), r["d_lin"].ravel(), "d_lin")


def test_oracle_equals_categorical_glm_with_identity_beta():
    rng = np.random.default_rng(4)
    N, C = 97, 11
    lin = rng.standard_normal((N, C)) * 3.0
    y = rng.integers(1, C + 1, N).astype(np.int32)
    o = po.categorical_logit_lpmf(y, lin)
    g = po.categorical_logit_glm(y, lin, np.zeros(C), np.eye(C),
                                 flags=po.VAR_ALPHA | po.VAR_BETA | po.VAR_X)
    assert_logp(o["logp"], g["logp"])
    assert_grad(o["d_lin"].ravel(), g["d_x"].ravel(), "d_lin vs glm d_x")
    m = lin.max(axis=1, keepdims=True)
    ls = lin - m - np.log(np.exp(lin - m).sum(axis=1, keepdims=True))
    assert_logp(o["logp"], float(ls[np.arange(N), y - 1].sum()))


def test_oracle_error_codes():
    lin = np.zeros((3, 2))
    assert po.categorical_logit_lpmf([1, 2], lin)["rc"] == 1          # size
    assert po.categorical_logit_lpmf([1, 3, 1], lin)["rc"] == 2       # y out of [1, C]
    assert po.categorical_logit_lpmf([0], lin)["rc"] == 2
    bad = lin.copy()
    bad[1, 1] = np.inf
    assert po.categorical_logit_lpmf([1, 2, 1], bad)["rc"] == 2       # check_finite
    bad[1, 1] = np.nan
    assert po.categorical_logit_lpmf([1, 2, 1], bad, po.PROPTO)["rc"] == 2
    assert po.categorical_logit_lpmf([1, 2, 1], lin, po.PROPTO) == \
        {"rc": 0, "logp": 0.0, "d_lin": pytest.approx(np.zeros((3, 2)))}
    assert po.categorical_logit_lpmf(np.zeros(0, np.int32), np.zeros((0, 4)))["logp"] == 0.0


# ------------------------------------------------------------------ GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_gpu_matches_golden(gpu, case):
    mb = gpu
    y, lin, logp, d = unpack(case)
    yy = int(y[0]) if y.size == 1 and lin.shape[0] != 1 else mb.to_matrix_cuda(y)
    r = mb.lpmf.categorical_logit_lpmf(yy, mb.to_matrix_cuda(lin), propto=case["propto"])
    assert_logp(r.logp, logp)
    assert_grad(mb.from_matrix_cuda(r.d_theta).ravel(), d.ravel(), "d_lin")


@pytest.mark.gpu
@pytest.mark.parametrize("N,C", [(1, 1), (1, 5), (257, 2), (4099, 8), (4099, 9), (50021, 32),
                                 (3001, 33), (1153, 43), (2049, 64), (1025, 65), (513, 128),
                                 (777, 200)])
def test_gpu_matches_oracle(gpu, N, C):
    mb = gpu
    rng = np.random.default_rng(N + C)
    lin = np.asfortranarray(rng.standard_normal((N, C)) * 4.0)
    y = rng.integers(1, C + 1, N).astype(np.int32)
    o = po.categorical_logit_lpmf(y, lin)
    lin_d, y_d = mb.to_matrix_cuda(lin), mb.to_matrix_cuda(y)
    r = mb.lpmf.categorical_logit_lpmf(y_d, lin_d)
    assert_logp(r.logp, o["logp"])
    assert_grad(mb.from_matrix_cuda(r.d_theta).ravel(), o["d_lin"].ravel(), "d_lin")
    # data log odds: value only (another kernel: the row sums may associate differently);
    # propto with data log odds: nothing left
    v = mb.lpmf.categorical_logit_lpmf(y_d, lin_d, lin_var=False).logp
    assert abs(v - r.logp) <= 1e-13 * abs(r.logp)
    assert mb.lpmf.categorical_logit_lpmf(y_d, lin_d, propto=True, lin_var=False).logp == 0.0
    # bit-identical repeat (static schedule, fixed-order sums)
    r2 = mb.lpmf.categorical_logit_lpmf(y_d, lin_d)
    assert r2.logp == r.logp
    assert np.array_equal(mb.from_matrix_cuda(r2.d_theta), mb.from_matrix_cuda(r.d_theta))


@pytest.mark.gpu
def test_gpu_lsu_kernels_still_match_oracle(gpu):
    """The LSU kernels behind the TMA pipeline (layouts TMA cannot address) stay covered:
    the same comparison in a fresh process with SMC_CATL_TMA=0."""
    import subprocess, sys, os
    code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np, math_b200 as mb
from oracle import pyoracle as po
mb.runtime.set_device(0)
for N, C in ((4099, 9), (50021, 32), (2049, 64), (513, 128)):
    rng = np.random.default_rng(N + C)
    lin = np.asfortranarray(rng.standard_normal((N, C)) * 4.0)
    y = rng.integers(1, C + 1, N).astype(np.int32)
    o = po.categorical_logit_lpmf(y, lin)
    r = mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lin))
    assert abs(r.logp - o["logp"]) <= 1e-10 * abs(o["logp"]), (N, C)
    d = mb.from_matrix_cuda(r.d_theta)
    assert np.allclose(d, o["d_lin"], rtol=1e-9, atol=1e-13), (N, C)
print("ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, SMC_CATL_TMA="0"))
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_gpu_error_contract(gpu):
    mb = gpu
    lin = np.asfortranarray(np.random.default_rng(0).standard_normal((6, 3)))
    lin_d = mb.to_matrix_cuda(lin)
    y = np.array([1, 2, 3, 1, 2, 3], dtype=np.int32)
    with pytest.raises(mb.DomainError):
        mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(y + 1), lin_d)
    with pytest.raises(mb.DomainError):
        mb.lpmf.categorical_logit_lpmf(0, lin_d)
    with pytest.raises(ValueError):
        mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(y[:4]), lin_d)
    for v in (np.inf, -np.inf, np.nan):
        bad = lin.copy()
        bad[4, 1] = v
        for propto in (False, True):
            with pytest.raises(mb.DomainError):
                mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(bad),
                                               propto=propto, lin_var=False)
    e = mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(np.zeros(0, np.int32)),
                                       mb.MatrixCuda(0, 3, np.float64))
    assert e.logp == 0.0


@pytest.mark.gpu
def test_gpu_full_size_properties(gpu):
    """N = 2e6, C = 32 (the class count of BASELINE config 5a): the value is
    additive over row blocks, each row of the partial sums to zero
    (one-hot - softmax), and a sampled block matches the oracle."""
    mb = gpu
    N, C = 2_000_000, 32
    rng = np.random.default_rng(11)
    lin = np.asfortranarray(rng.standard_normal((N, C)) * 3.0)
    y = rng.integers(1, C + 1, N).astype(np.int32)
    r = mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lin))
    cut = 1_234_567
    parts = [mb.lpmf.categorical_logit_lpmf(mb.to_matrix_cuda(y[a:b]),
                                            mb.to_matrix_cuda(np.asfortranarray(lin[a:b])),
                                            lin_var=False).logp
             for a, b in ((0, cut), (cut, N))]
    assert_logp(r.logp, parts[0] + parts[1])
    d = mb.from_matrix_cuda(r.d_theta)
    assert np.abs(d.sum(axis=1)).max() < 1e-13
    a, b = 999_000, 1_001_000
    o = po.categorical_logit_lpmf(y[a:b], lin[a:b])
    assert_grad(d[a:b].ravel(), o["d_lin"].ravel(), "d_lin block")


# ------------------------------------------- GPU: the matrix products either side of it
@pytest.mark.gpu
@pytest.mark.parametrize("N,K,C", [(1, 1, 1), (5, 2, 3), (1531, 37, 5), (4099, 128, 32),
                                   (3001, 300, 8), (777, 21, 70), (2050, 64, 64),
                                   (1025, 513, 130)])
def test_gpu_matrix_product_and_adjoint(gpu, N, K, C):
    """lin = x beta + alpha^T and its reverse sweep (x^T adj, column sums) against
    numpy; the DMMA sweeps are fixed-order, so repeats are bit-identical."""
    mb = gpu
    rng = np.random.default_rng(N + K + C)
    x = np.asfortranarray(rng.standard_normal((N, K)))
    beta = rng.standard_normal((K, C)) / np.sqrt(K)
    alpha = rng.standard_normal(C)
    x_d = mb.to_matrix_cuda(x)
    lin = mb.from_matrix_cuda(mb.lpmf.multiply_matrix(x_d, beta, alpha))
    want = x @ beta + alpha
    assert lin.shape == (N, C)
    assert_grad(lin.ravel(), want.ravel(), "x beta + alpha",
                scale=(np.abs(x) @ np.abs(beta)).max())
    lin0 = mb.from_matrix_cuda(mb.lpmf.multiply_matrix(x_d, beta))
    assert_grad(lin0.ravel(), (x @ beta).ravel(), "x beta",
                scale=(np.abs(x) @ np.abs(beta)).max())
    adj = np.asfortranarray(rng.standard_normal((N, C)))
    adj_d = mb.to_matrix_cuda(adj)
    g, cs = mb.lpmf.multiply_matrix_adjoint(x_d, adj_d)
    assert_grad(g.ravel(), (x.T @ adj).ravel(), "x^T adj", scale=(np.abs(x).T @ np.abs(adj)).max())
    assert_grad(cs, adj.sum(axis=0), "column sums", scale=np.abs(adj).sum(axis=0).max())
    g2, cs2 = mb.lpmf.multiply_matrix_adjoint(x_d, adj_d)
    assert np.array_equal(g, g2) and np.array_equal(cs, cs2)


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,C", [(30011, 128, 32), (5003, 40, 7)])
def test_gpu_unfused_categorical_pipeline_equals_fused_glm(gpu, N, K, C):
    """multiply_matrix -> categorical_logit_lpmf -> multiply_matrix_adjoint gives the
    fused categorical GLM's value, d_alpha and d_beta, and the oracle's."""
    mb = gpu
    rng = np.random.default_rng(N + C)
    x = np.asfortranarray(rng.standard_normal((N, K)))
    beta = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(K))
    alpha = rng.standard_normal(C) * 0.3
    y = rng.integers(1, C + 1, N).astype(np.int32)
    x_d, y_d = mb.to_matrix_cuda(x), mb.to_matrix_cuda(y)
    fused = mb.categorical_logit_glm_lpmf(y_d, x_d, alpha, beta)
    un = mb.lpmf.categorical_logit_lpmf(y_d, mb.lpmf.multiply_matrix(x_d, beta, alpha))
    g, cs = mb.lpmf.multiply_matrix_adjoint(x_d, un.d_theta)
    o = po.categorical_logit_glm(y, x, alpha, beta)
    for ref_logp, ref_da, ref_db, tag in ((fused.logp, fused.d_alpha, fused.d_beta, "fused"),
                                          (o["logp"], o["d_alpha"], o["d_beta"], "oracle")):
        assert_logp(un.logp, ref_logp)
        assert_grad(cs, np.asarray(ref_da).ravel(), "d_alpha vs " + tag)
        assert_grad(g.ravel(order="F"), np.asarray(ref_db).ravel(order="F"), "d_beta vs " + tag)
