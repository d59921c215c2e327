import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import slimtest as st

    st.build_oracle(ref=Path("/root/reference/src/libslim/estimate.c").exists())
    return st.Oracle()


@pytest.fixture(scope="session")
def ml100k():
    import slimtest as st

    return st.load_golden("ml100k")


@pytest.fixture(scope="session")
def automotive():
    import slimtest as st

    return st.load_golden("automotive")
