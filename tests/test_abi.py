"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads without a GPU, exports exactly the
symbols include/sliced_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import subprocess

import pytest

import sliced_b200
from sliced_b200 import capi


def test_header_symbols_all_exported():
    lib = capi.load(build_if_missing=True)
    declared = capi.declared_symbols()
    assert len(declared) >= 60
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the binding covers the whole header
    unbound = [s for s in declared if s not in lib._sl_signatures]
    assert not unbound, f"declared in the header but not bound in capi.py: {unbound}"


def test_abi_version():
    assert capi.load(build_if_missing=True).sl_abi_version() == 1


def test_is_sm100a_native_code():
    """The shipped cubin is sm_100a and the gemm kernel really is tcgen05 + TMA (SASS mnemonics)."""
    capi.load(build_if_missing=True)
    out = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "no tcgen05.mma in the binary"
    assert "UTMALDG" in sass, "no TMA load in the binary"
    assert "LDTM" in sass, "no tcgen05.ld in the binary"


def test_no_cpu_fallback_without_device():
    if sliced_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sliced_b200.SlicedError) as e:
        sliced_b200.Context(0)
    assert e.value.code == capi.SL_ERR_NO_DEVICE
