import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Collection order of the GPU files: the core rows of SURVEY.md section 8 (a1-a16 parity, the compiled reference, scenes, user
# kernels / stacks, update) run first, so that with `pytest -x` a late failure in a "next" row cannot hide them.
GPU_ORDER = ["test_gpu_parity", "test_gpu_reference", "test_gpu_scene", "test_gpu_user_kernels", "test_gpu_update", "test_gpu_configs",
             "test_gpu_golden", "test_gpu_batch", "test_gpu_import", "test_gpu_motion", "test_gpu_threads", "test_gpu_multilevel",
             "test_gpu_bitcode", "test_gpu_sbvh", "test_gpu_pathtrace", "test_gpu_multi_gpu"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return GPU_ORDER.index(name) if name in GPU_ORDER else len(GPU_ORDER)

    items.sort(key=key)  # stable: order inside a file is kept


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
