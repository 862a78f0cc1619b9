"""GPU: the CUDA path (through the C ABI) against tests/golden/glm_golden.json,
the outputs of the UNMODIFIED reference on the reference tests' own inputs and on
the device-test shapes (small_simple, big 153x71, broadcast_y, vector alpha)."""
import numpy as np
import pytest

from tests import golden_util as gu
from tests.util import assert_grad, assert_logp

pytestmark = pytest.mark.gpu

CASES = gu.load()


def _dev(gpu, v, scalar_ok=True):
    v = np.asarray(v)
    if v.size == 1 and scalar_ok:
        return v.ravel()[0].item()
    return gpu.to_matrix_cuda(v)


def _host(v):
    return v.to_host().ravel(order="F") if hasattr(v, "to_host") else np.atleast_1d(v)


@pytest.mark.parametrize("case", CASES, ids=gu.ids(CASES))
def test_cuda_matches_reference_golden(gpu, case):
    fam = case["family"]
    d = gu.inputs_of(case)
    x = gpu.to_matrix_cuda(d["x"])
    y = _dev(gpu, d["y"])
    for v in case["variants"]:
        e = v["expect"]
        var = ["alpha", "beta"] + (["x"] if v["x_var"] else [])
        if fam in ("bernoulli", "poisson"):
            fn = gpu.bernoulli_logit_glm_lpmf if fam == "bernoulli" else gpu.poisson_log_glm_lpmf
            r = fn(y, x, _dev(gpu, d["alpha"]), d["beta"], propto=v["propto"], var=var)
        elif fam == "binomial":
            r = gpu.binomial_logit_glm_lpmf(y, _dev(gpu, d["trials"]), x,
                                            _dev(gpu, d["alpha"]), d["beta"],
                                            propto=v["propto"], var=var)
        elif fam == "normal":
            r = gpu.normal_id_glm_lpdf(y, x, _dev(gpu, d["alpha"]), d["beta"],
                                       _dev(gpu, d["sigma"]), propto=v["propto"],
                                       var=var + ["sigma"] + (["y"] if v["x_var"] else []))
        elif fam == "neg_binomial":
            r = gpu.neg_binomial_2_log_glm_lpmf(y, x, _dev(gpu, d["alpha"]), d["beta"],
                                                _dev(gpu, d["phi"]), propto=v["propto"],
                                                var=var + ["phi"])
        elif fam == "ordered":
            r = gpu.ordered_logistic_glm_lpmf(y, x, d["beta"], d["cuts"],
                                              propto=v["propto"],
                                              var=["beta", "cuts"] + (["x"] if v["x_var"] else []))
        else:
            r = gpu.categorical_logit_glm_lpmf(y, x, d["alpha"], d["beta"],
                                               propto=v["propto"], var=var)
        assert_logp(r.logp, e["logp"])
        sc = float(np.max(np.abs(e["d_beta"]))) * 1e-2
        assert_grad(np.asarray(r.d_beta).ravel(order="F"), e["d_beta"], "d_beta")
        if "d_alpha" in e:
            assert_grad(_host(r.d_alpha), np.atleast_1d(e["d_alpha"]), "d_alpha", scale=sc)
        for key in ("d_sigma", "d_phi", "d_cuts"):
            if key in e:
                assert_grad(_host(r.d_aux), np.atleast_1d(e[key]), key, scale=sc)
        if "d_x" in e:
            assert_grad(r.d_x.to_host().ravel(order="F"), e["d_x"], "d_x")
        if "d_y" in e:
            assert_grad(_host(r.d_y), np.atleast_1d(e["d_y"]), "d_y", scale=sc)
