import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine_factory():
    """Builds CqlEngine instances on cuda:0; closes them at session end."""
    from replay_cql_b200.engine import CqlEngine, CqlHyperParams
    made = []

    def make(**kw):
        hp = CqlHyperParams(**kw)
        eng = CqlEngine(hp, device=0)
        made.append(eng)
        return eng

    yield make
    for e in made:
        e.close()
