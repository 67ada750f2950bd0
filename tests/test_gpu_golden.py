"""The reference's own known-answer tests (tests/golden/reference_kats.py), replayed on the CUDA path through the C ABI."""
import pytest

from tests.golden.reference_kats import KATS
from tests.kat_runner import run_kat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from tests.cuda_backend import CudaBackend
    return CudaBackend()


@pytest.mark.parametrize("kat", KATS, ids=[k["id"] for k in KATS])
def test_cuda_kat(be, kat):
    run_kat(be, kat)


@pytest.mark.parametrize("kat", [k for k in KATS if k["kind"] in ("gemm", "gemm_grad", "tgemm_equiv", "gemmt_equiv")],
                         ids=lambda k: k["id"])
def test_cuda_kat_gemm_modes_small(kat):
    """tiny gemms route to the CUDA-core kernel in every mode: exact answers"""
    from tests.cuda_backend import CudaBackend
    import sliced_b200 as S
    for mode in (S.GEMM_3XTF32, S.GEMM_TF32, S.GEMM_SIMT):
        run_kat(CudaBackend(gemm_mode=mode), kat)
