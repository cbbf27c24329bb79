"""
GPU: the reference-side binding end to end.  The descriptors in tests/golden/reference_binding.npz were marshalled by
integration/eml_binding.py from the REAL reference's BasicModel / TwoLevelSystem objects (tests/golden/make_binding_golden.py);
here they go through eml_plan_create / eml_eval_batch_host exactly as the stub at basic_model.py:381 would send them, and
the result is compared with what the reference's own calculate_dynamics returned for the same models
(reference_operators.npz: expect_i, mesolve being the exact solution in the shim).  Tolerance: north_star's 1e-8 relative.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'integration'))

import eml_binding  # noqa: E402
from effective_model_learning_b200 import definitions  # noqa: E402
from tests.helpers import OBS3, assert_close, model_from_spec  # noqa: E402
from tests.test_binding import _fixture, _specs  # noqa: E402

pytestmark = pytest.mark.gpu


def test_reference_objects_through_the_c_abi_match_the_reference_dynamics():
    ref = np.load(os.path.join(HERE, 'golden', 'reference_operators.npz'))
    worst = 0.0
    for i, desc in enumerate(_fixture()):
        got = eml_binding.evaluate(desc)
        # the observables and rho0 of the fixture are Hermitian, for which mesolve reports the REAL part of Tr(O rho)
        # (models 17 and 19 have a non-Hermitian "H" -- sigmam in a coupling -- so the trace itself is complex there)
        want = ref[f'expect_{i}'].real
        assert_close(list(got.real), list(want), tol=1e-8, what=f'golden model {i}')
        worst = max(worst, float(np.abs(got.real - want).max()))
    assert worst < 1e-8


def test_expect_returns_what_mesolve_returned():
    """Types and shapes of `qutip.mesolve(...).expect` (basic_model.py:389): K arrays (T,), float64 for Hermitian e_ops."""
    specs, ts = _specs()
    ref = np.load(os.path.join(HERE, 'golden', 'reference_operators.npz'))
    for i in (0, 5, 13, 20):
        out = eml_binding.expect(model_from_spec(specs[i]), ts, OBS3, definitions.ops)
        assert len(out) == 3 and all(o.shape == (len(ts),) and o.dtype == np.float64 for o in out)
        assert_close(out, list(ref[f'expect_{i}'].real), tol=1e-8, what=f'golden model {i}')
    out = eml_binding.expect(model_from_spec(specs[3]), ts, ['sigmap', 'sigmaz'], definitions.ops)
    assert out[0].dtype == np.complex128 and out[1].dtype == np.float64
