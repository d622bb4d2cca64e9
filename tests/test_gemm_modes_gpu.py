"""All arithmetic modes of the O(n p^2) products must pass the same parity tests: the main suite runs in the library default
(INT8 tensor cores with CTA pairs, mode 3); this file re-runs the parity subset with each mode selected explicitly."""
import pytest

from conftest import model_cases
from gp_grief_b200 import _native as nat
import test_native_gpu as tn
import test_api_gpu as ta

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 0, 5], ids=["int8", "fp64", "int8-cluster"])       # 5: CTA pairs on every product (3 = only >= 16 x 16 tiles)
def int8_mode(request):
    before = nat.lib().grief_get_default_option(nat.OPT_GEMM_MODE)
    nat.lib().grief_set_gemm_mode(request.param)
    try:
        yield request.param
    finally:
        nat.check(nat.lib().grief_set_default_option(nat.OPT_GEMM_MODE, before))


def test_default_mode_is_int8():
    assert nat.lib().grief_get_gemm_mode() == 1                                    # INT8 arithmetic ...
    assert nat.lib().grief_get_default_option(nat.OPT_GEMM_MODE) == 3              # ... on CTA pairs


@pytest.mark.gpu
@pytest.mark.parametrize("name", model_cases())
def test_phi_and_gram_int8(int8_mode, name):
    tn.test_phi_and_gram_match_reference(name, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", model_cases())
def test_lml_and_adjoint_gradient_int8(int8_mode, name):
    tn.test_lml_and_adjoint_gradient_match_reference(name)


@pytest.mark.gpu
def test_larger_shapes_int8(int8_mode):
    assert nat.lib().grief_get_gemm_mode() == (int8_mode & 1)
    tn.test_gram_larger_random_shape_vs_materialised_phi()
    tn.test_gram_and_quadform_over_several_slabs()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["syn_t2_n2000_d4_m8_p64", "syn_t2_matern52_n1500_d3_m10_p48"])
def test_type2_analytic_gradient_int8(int8_mode, name):
    ta.test_type2_analytic_gradient(name)


@pytest.mark.gpu
def test_models_int8(int8_mode):
    ta.test_reference_test_gp_grief_model()
    ta.test_checkgrad_and_optimize_type2()
    ta.test_grief_to_web_model_shares_statistics()


@pytest.mark.gpu
def test_non_finite_inputs_are_reported_not_swallowed():
    """INT8 mode cuts operands into integer digits: a NaN must surface as an error, not silently become zero."""
    import numpy as np
    import torch
    from gp_grief_b200 import device
    rng = np.random.default_rng(3)
    d, m, p, n = 3, 5, 40, 600
    xg = [np.linspace(0, 1, m) for _ in range(d)]
    basis = tn.orc.setup_inducing_cov(["RBF"] * d, [1.0] * d, [0.4] * d, xg, p)
    plan = tn._plan_from_basis(dict(d=d, names=["RBF"] * d, variances=[1.0] * d, lengthscales=[0.4] * d, xg=xg), basis)
    X = rng.random((n, d))
    X[17, 1] = np.nan
    T = plan.build_tables(torch.from_numpy(X).cuda())
    assert nat.lib().grief_get_gemm_mode() == 1
    with pytest.raises(ValueError, match="non-finite"):
        plan.gram(T, n)
    X[17, 1] = 0.5                                   # the flag is cleared: the next call works
    T = plan.build_tables(torch.from_numpy(X).cuda())
    A = plan.gram(T, n).cpu().numpy()
    assert np.all(np.isfinite(A))
