import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a) GPU; run with -m gpu")


def _gpu_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] >= 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no B200 GPU in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist (built by __graft_entry__.build()); build if missing."""
    import importlib.util
    so = os.path.join(ROOT, "candle-vllm_b200", "libb200backend.so")
    if not os.path.exists(so):
        spec = importlib.util.spec_from_file_location("b200_build", os.path.join(ROOT, "candle-vllm_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    yield
