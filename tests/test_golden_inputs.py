"""A stale fixture must fail: every tests/golden/<solve case>.npz holds exactly the inputs that the committed seeded
generators (oracle/golden_inputs.py, torch only — the reference is not needed) produce for that case, bit for bit.
oracle/gen_golden.py feeds the same generators to the unmodified reference, so a fixture whose inputs match was
produced by the committed script on the committed configuration (lti8_T50_f64 was not)."""
import numpy as np
import pytest

import parity as P

from oracle import golden_inputs as GI  # noqa: E402


def test_every_solve_case_has_a_generator():
    assert set(P.SOLVE_CASES) == set(GI.CASES)


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_fixture_inputs_come_from_the_committed_generator(name):
    g = P.load(name)
    case = GI.make(name)
    want = GI.as_numpy(case)
    for k, v in want.items():
        assert g[k].dtype == v.dtype and g[k].shape == v.shape, (k, g[k].dtype, v.dtype, g[k].shape, v.shape)
        assert np.array_equal(g[k], v), f"{name}: fixture input {k} is not what oracle/golden_inputs.py generates"
    assert ("u_min" in g) == ("u_min" in want)
    for k, v in dict(nx="nx", nu="nu", T="T").items():
        assert int(g["meta_" + k]) == case[v]
    assert float(g["meta_dt"]) == case["dt"] and float(g["meta_reg_eps"]) == case.get("reg_eps", 1e-6)
    assert str(g["meta_grad_method"]) == case["grad_method"]
    for k, v in case["spec_extra"].items():
        assert np.array_equal(np.asarray(g["meta_" + k]), np.asarray(v)), k


def test_generators_do_not_depend_on_the_ambient_default_dtype():
    import torch
    a = GI.as_numpy(GI.make("lti8_T50_f64"))
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        b = GI.as_numpy(GI.make("lti8_T50_f64"))
    finally:
        torch.set_default_dtype(old)
    assert all(np.array_equal(a[k], b[k]) for k in a)
