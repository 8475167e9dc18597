import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def events_small():
    """First 300 K NEXMark events (seed 42) cut into 8 Ki-row batches: many batches, runs in seconds on the CPU."""
    from flock_b200 import nexgen
    return nexgen.generate(300_000, seed=42, batch_rows=8192)


@pytest.fixture(scope="session")
def events_seed7():
    from flock_b200 import nexgen
    return nexgen.generate(120_000, seed=7, batch_rows=65536)


@pytest.fixture(scope="session")
def gpu_ctx():
    """One GPU context for the whole session.  No skip: on a box without a device this fails loudly."""
    import flock_b200 as fb
    ctx = fb.Context(int(os.environ.get("FLOCKGPU_DEVICE", "0")))
    yield ctx
    ctx.close()


def sources_for(query: str, events: dict) -> list:
    """sources[relation][partition][batch] in the feed order of the query (one partition per relation)."""
    from flock_b200 import plans
    return [[events[r]] for r in plans.SOURCES[query]]
