import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O

    O.build()
    return O


@pytest.fixture(scope="session")
def template_path(tmp_path_factory, oracle):
    """A BRIEF template file in the reference's text format, written from the repo's built-in table."""
    p = tmp_path_factory.mktemp("tmpl") / "brief_template.txt"
    return oracle.write_template_file(str(p))


@pytest.fixture(scope="session")
def have_ref(oracle):
    return oracle.have_ref()


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu on the B200 box; without a device they are skipped, never silently passed.
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
