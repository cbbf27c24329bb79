import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # the C-ABI library is a build artefact (git-ignored): build it once if this checkout has none yet
    lib = os.path.join(ROOT, 'effective_model_learning_b200', 'libeml_b200.so')
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture
def oracle_backend(monkeypatch):
    """Route EvalContext.evaluate to the CPU oracle (host-logic tests on a box without a GPU)."""
    from effective_model_learning_b200 import engine
    from tests.helpers import oracle_evaluate
    monkeypatch.setattr(engine.EvalContext, 'evaluate', oracle_evaluate)
    engine.clear_contexts()
    yield
    engine.clear_contexts()
