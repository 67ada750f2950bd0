import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # `-m gpu` tests need a real B200 and fail (not skip) without one: the product path has no CPU fallback.
    # `-m "not gpu"` covers the oracle against the reference's golden vectors, the host logic and the C-ABI surface.
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
