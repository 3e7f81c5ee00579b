import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle, the host-emulation library and the CUDA library exist (build is incremental)."""
    from bulletproofs_b200 import build as b
    b.build_oracle()
    b.build_emul()
    b.build_cuda()
    b.build_host()
    return True


@pytest.fixture(scope="session")
def orc(built):
    from oracle_binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "range_proof_v1.json")))


@pytest.fixture(scope="session")
def gpu_ctx(built):
    import bulletproofs_b200 as bp
    ctx = bp.Context(0)           # raises if there is no B200: GPU tests must not pass on a fallback
    yield ctx
    ctx.close()
