"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads without a GPU, exports exactly the
symbols include/sliced_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import os
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


def test_rust_sys_crate_matches_header():
    """bindings/rust/sliced-b200-sys/src/lib.rs is generated from include/sliced_b200.h (tools/gen_rust_sys.py): it must be up to
    date and declare exactly the header's symbols (rustc is not in the image: this is the check the crate gets)."""
    import re
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    from sliced_b200 import capi
    lib_rs = open(os.path.join(root, "bindings", "rust", "sliced-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (sl_\w+)\(", lib_rs))
    assert declared == set(capi.declared_symbols())


def test_rust_patch_uses_only_declared_symbols():
    """the authored `cuda.rs` files of bindings/rust/reference-patch call only functions the header declares, with the C prototypes'
    argument counts"""
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    lib_rs = open(os.path.join(root, "bindings", "rust", "sliced-b200-sys", "src", "lib.rs")).read()
    arity = {m.group(1): (len([a for a in m.group(2).split(",") if a.strip()])) for m in re.finditer(r"pub fn (sl_\w+)\(([^)]*)\)", lib_rs)}
    patch = os.path.join(root, "bindings", "rust", "reference-patch", "src")
    used = 0
    for d, _, files in os.walk(patch):
        for f in files:
            if not f.endswith(".rs"):
                continue
            txt = open(os.path.join(d, f)).read()
            txt = re.sub(r"//.*", "", txt)
            txt = re.sub(r'"(?:[^"\\]|\\.)*"', '""', txt)   # string literals may mention entry points in prose
            for m in re.finditer(r"\b(sl_[a-z0-9_]+)\s*\(", txt):
                name = m.group(1)
                if name in ("sl_ctx",):
                    continue
                assert name in arity, f"{f}: {name} is not declared in include/sliced_b200.h"
                # count top-level commas of the call's argument list
                i, depth, commas, any_arg = m.end(), 1, 0, False
                while depth:
                    c = txt[i]
                    if c in "([{":
                        depth += 1
                    elif c in ")]}":
                        depth -= 1
                    elif c == "," and depth == 1:
                        commas += 1
                    elif not c.isspace():
                        any_arg = True
                    i += 1
                assert (commas + 1 if any_arg else 0) == arity[name], f"{f}: {name} called with {commas + 1} args, header has {arity[name]}"
                used += 1
    assert used >= 20
    # one authored `cuda.rs` per device trait family of SURVEY 8(b), plus the custos apply_fn / add_unary_grad shim
    want = ["binary_ew/cuda.rs", "gemm/cuda.rs", "gemm/grad/cuda.rs", "row_op/cuda.rs", "row_op/grad/cuda.rs", "col_op/cuda.rs", "col_op/grad/cuda.rs",
            "max/cuda.rs", "sum/cuda.rs", "mean/cuda.rs", "mean/grad/cuda.rs", "transpose/cuda.rs", "softmax/cuda.rs", "diagflat/cuda.rs",
            "diagflat/grad/cuda.rs", "onehot/cuda.rs", "onehot/grad/cuda.rs"]
    for w in want:
        assert os.path.exists(os.path.join(patch, "ops2", w)), f"reference-patch is missing src/ops2/{w}"
    assert os.path.exists(os.path.join(patch, "custos_shim", "apply_fn.rs"))
