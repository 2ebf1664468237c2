import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)"
    )
    config.addinivalue_line(
        "markers", "gpu_multi: needs >= 2 CUDA devices of one node (gpurun --gpus N); skipped otherwise"
    )
