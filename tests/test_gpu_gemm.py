"""
gemm / trans_gemm / gemm_grad parity on the GPU (through the C ABI).

Stated tolerances (BASELINE.json north_star: "K-scaled for gemm"), inputs U(-1,1):
  * SL_GEMM_SIMT   : BIT-EXACT vs the oracle's restatement (sequential-k mul+add, no FMA on either side)
  * SL_GEMM_3XTF32 : max |c - truth_fp64| <= 4 * K * 2^-24      (fp32-class; also held to <= 16x the oracle sgemm's own error)
  * SL_GEMM_TF32   : max |c - truth_fp64| <= 8 * sqrt(K) * 2^-11 (reported, loose gate: flagged fast mode)
"""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import sliced_b200 as S
    return S.Context(0)


@pytest.fixture(autouse=True)
def force_tensor_core_path(monkeypatch):
    """route every f32 gemm of this module through the tcgen05 kernel, whatever its size (default dispatch sends tiny
    problems to the CUDA-core kernel)"""
    monkeypatch.setenv("SLICED_GEMM_TC_FORCE", "1")


def truth(ta, tb, m, n, k, a, b):
    A = a.reshape(k, m).T if ta else a.reshape(m, k)
    B = b.reshape(n, k).T if tb else b.reshape(k, n)
    return (A.astype(np.float64) @ B.astype(np.float64)).ravel()


def run(ctx, ta, tb, m, n, k, a, b, mode, c0=None, accumulate=False):
    dc = ctx.array(c0) if c0 is not None else None
    return ctx.gemm_ex(ta, tb, m, n, k, ctx.array(a), ctx.array(b), dc, accumulate, mode).numpy()


SMALL = [(1, 1, 1), (4, 3, 2), (5, 10, 7), (64, 64, 64), (65, 63, 33), (130, 70, 257), (1000, 64, 64), (33, 129, 1000)]


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
@pytest.mark.parametrize("shape", SMALL)
@pytest.mark.parametrize("trans", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_simt_bit_exact_vs_oracle(ctx, dt, shape, trans):
    import sliced_b200 as S
    m, n, k = shape
    ta, tb = trans
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    if dt == np.int32:
        a, b = rng.integers(-9, 9, m * k).astype(dt), rng.integers(-9, 9, k * n).astype(dt)
        c0 = rng.integers(-9, 9, m * n).astype(dt)
    else:
        a, b = rng.uniform(-1, 1, m * k).astype(dt), rng.uniform(-1, 1, k * n).astype(dt)
        c0 = rng.uniform(-1, 1, m * n).astype(dt)
    ref = O.gemm_ex(ta, tb, m, n, k, a, b)
    got = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_SIMT)
    assert np.array_equal(got, ref), f"max diff {np.max(np.abs(got.astype(np.float64) - ref))}"
    ref_acc = O.gemm_ex(ta, tb, m, n, k, a, b, c0.copy(), accumulate=True)
    got_acc = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_SIMT, c0, True)
    assert np.array_equal(got_acc, ref_acc)


TC_SHAPES = [
    (1, 1, 1), (4, 3, 2), (17, 5, 9), (128, 128, 32), (128, 256, 64), (256, 128, 128), (128, 128, 4096),
    (64, 64, 1024), (130, 70, 260), (200, 300, 100), (1000, 64, 64), (384, 520, 36), (129, 257, 513),
    (512, 512, 512), (1024, 1024, 1024), (1000, 1000, 1000), (777, 1234, 555), (2048, 256, 64), (4096, 128, 32),
    (257, 1030, 31 + 33), (300, 200, 1002),  # K % 4 != 0 -> padded planes
]


@pytest.mark.parametrize("shape", TC_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("trans", [(0, 0), (0, 1), (1, 0)], ids=["NN", "NT", "TN"])
@pytest.mark.parametrize("mode", ["3xtf32", "tf32", "3xf16"])
def test_tensor_core_gemm(ctx, shape, trans, mode):
    import sliced_b200 as S
    m, n, k = shape
    ta, tb = trans
    md = {"3xtf32": S.GEMM_3XTF32, "tf32": S.GEMM_TF32, "3xf16": S.GEMM_3XF16}[mode]
    rng = np.random.default_rng(42 + m + n + k)
    a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
    t = truth(ta, tb, m, n, k, a, b)
    got = run(ctx, ta, tb, m, n, k, a, b, md)
    err = np.max(np.abs(got - t))
    if mode != "tf32":   # 3xf16 on shapes its kernel cannot take is 3xtf32; either way the fp32-class bound holds
        ref = O.gemm_ex(ta, tb, m, n, k, a, b)
        ref_err = np.max(np.abs(ref - t))
        assert err <= 4 * k * 2.0 ** -24, f"3xTF32 err {err} (oracle sgemm err {ref_err})"
        assert err <= 16 * ref_err + 2.0 ** -22, f"3xTF32 err {err} vs oracle sgemm err {ref_err}"
    else:
        assert err <= 8 * np.sqrt(k) * 2.0 ** -11, f"TF32 err {err}"


@pytest.mark.parametrize("cfg", ["1", "2", "3", "4"])
@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
def test_tile_configs(cfg, mode):
    """every tile configuration of the tcgen05 kernel (SLICED_GEMM_CFG) gives the same answers; tails included"""
    import sliced_b200 as S
    os.environ["SLICED_GEMM_CFG"] = cfg
    try:
        ctx = S.Context(0)
        md = S.GEMM_3XTF32 if mode == "3xtf32" else S.GEMM_TF32
        # NN / NT / TN: with cfg 4 (2-CTA) non-K-contiguous operands are consumed MN-major (no transposing pass) when their
        # extent is a multiple of 4, and through a transposing split otherwise (1111 x 2222) — both paths are exercised
        for (m, n, k) in [(128, 256, 128), (640, 768, 1024), (300, 700, 200), (4096, 512, 96), (1111, 2222, 333), (512, 1024, 40)]:
            for (ta, tb) in ((0, 0), (0, 1), (1, 0), (1, 1)):
                rng = np.random.default_rng(m + n + k)
                a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
                t = truth(ta, tb, m, n, k, a, b)
                got = run(ctx, ta, tb, m, n, k, a, b, md)
                err = np.max(np.abs(got - t))
                assert err <= (4 * k * 2.0 ** -24 if mode == "3xtf32" else 8 * np.sqrt(k) * 2.0 ** -11), (cfg, mode, ta, tb, m, n, k, err)
        ctx.close()
    finally:
        os.environ.pop("SLICED_GEMM_CFG", None)


F16_SHAPES = [(128, 256, 128), (256, 256, 64), (640, 768, 1024), (304, 704, 200), (4096, 512, 96), (1112, 2224, 336), (512, 1024, 40),
              (264, 8, 72), (8, 264, 4104), (2048, 2048, 2048),
              (256, 512, 16392)]   # rows longer than the register-cached limit of the row-wise split (8192)


@pytest.mark.parametrize("shape", F16_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_3xf16_kernel(shape, monkeypatch):
    """SL_GEMM_3XF16 through the kind::f16 2-CTA kernel itself (STRICT: no silent 3xTF32 substitution), all four operand
    layouts (K-major row-scaled and MN-major column-scaled planes), tails in M / N / K, same fp32-class bound as 3xTF32."""
    import sliced_b200 as S
    monkeypatch.setenv("SLICED_GEMM_CFG", "4")
    monkeypatch.setenv("SLICED_GEMM_F16_STRICT", "1")
    monkeypatch.setenv("SLICED_GEMM_TC_FORCE", "1")
    ctx = S.Context(0)
    m, n, k = shape
    for (ta, tb) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        rng = np.random.default_rng(m + n + k + ta * 2 + tb)
        a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
        t = truth(ta, tb, m, n, k, a, b)
        got = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_3XF16)
        err = np.max(np.abs(got - t))
        ref_err = np.max(np.abs(O.gemm_ex(ta, tb, m, n, k, a, b) - t))
        assert err <= 4 * k * 2.0 ** -24 and err <= 16 * ref_err + 2.0 ** -22, (ta, tb, m, n, k, err, ref_err)
        c0 = rng.uniform(-5, 5, m * n).astype(np.float32)
        got = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_3XF16, c0, True)
        assert np.max(np.abs(got - (t + c0))) <= 4 * k * 2.0 ** -24 + 2.0 ** -21
    ctx.close()


def test_3xf16_scaling_is_exact_and_per_output_index(monkeypatch):
    """Rows of A and columns of B spread over 2^-40 .. 2^40: every output element must keep fp32-class RELATIVE accuracy with
    respect to its own row/column scale (the power-of-two scaling is per output index and exact), zero rows / columns stay
    exactly zero, and scaling an operand by a power of two scales the result bit-exactly."""
    import sliced_b200 as S
    monkeypatch.setenv("SLICED_GEMM_CFG", "4")
    monkeypatch.setenv("SLICED_GEMM_F16_STRICT", "1")
    ctx = S.Context(0)
    m, n, k = 512, 768, 1024
    rng = np.random.default_rng(3)
    ra = np.ldexp(1.0, rng.integers(-40, 40, m)).astype(np.float32)
    cb = np.ldexp(1.0, rng.integers(-40, 40, n)).astype(np.float32)
    a0 = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    b0 = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    a0[7] = 0
    b0[:, 11] = 0
    for (ta, tb) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        a = a0 * ra[:, None]
        b = b0 * cb[None, :]
        am = np.ascontiguousarray(a.T if ta else a).ravel()
        bm = np.ascontiguousarray(b.T if tb else b).ravel()
        got = run(ctx, ta, tb, m, n, k, am, bm, S.GEMM_3XF16).reshape(m, n).astype(np.float64)
        t = a.astype(np.float64) @ b.astype(np.float64)
        rel = np.abs(got - t) / (ra[:, None].astype(np.float64) * cb[None, :].astype(np.float64))
        assert np.max(rel) <= 4 * k * 2.0 ** -24, (ta, tb, np.max(rel))
        assert np.all(got[7] == 0) and np.all(got[:, 11] == 0)
        base = run(ctx, ta, tb, m, n, k, np.ascontiguousarray(a0.T if ta else a0).ravel(), np.ascontiguousarray(b0.T if tb else b0).ravel(),
                   S.GEMM_3XF16).reshape(m, n)
        assert np.array_equal(got.astype(np.float32), base * ra[:, None] * cb[None, :])
    ctx.close()


def test_3xf16_wide_dynamic_range_inside_a_row(monkeypatch):
    """Elements far below their row's maximum lose low bits in fp16 (floor: 2^-40 of the row maximum); the result must still be
    fp32-class in the norm-wise sense |err| <= c K eps max|a_i| max|b_j|."""
    import sliced_b200 as S
    monkeypatch.setenv("SLICED_GEMM_CFG", "4")
    monkeypatch.setenv("SLICED_GEMM_F16_STRICT", "1")
    ctx = S.Context(0)
    m, n, k = 256, 512, 2048
    rng = np.random.default_rng(9)
    a = (rng.uniform(-1, 1, (m, k)) * np.ldexp(1.0, rng.integers(-30, 1, (m, k)))).astype(np.float32)
    b = (rng.uniform(-1, 1, (k, n)) * np.ldexp(1.0, rng.integers(-30, 1, (k, n)))).astype(np.float32)
    got = run(ctx, 0, 0, m, n, k, a.ravel(), b.ravel(), S.GEMM_3XF16).reshape(m, n).astype(np.float64)
    t = a.astype(np.float64) @ b.astype(np.float64)
    bound = 4 * k * 2.0 ** -24 * np.max(np.abs(a), axis=1)[:, None] * np.max(np.abs(b), axis=0)[None, :]
    assert np.all(np.abs(got - t) <= bound)
    # and against the element-wise fp32 bound  K eps sum|a||b|  (what a CPU sgemm guarantees): report, gate at 8x
    ew = k * 2.0 ** -24 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
    ratio = np.max(np.abs(got - t) / ew)
    print(f"3xf16 wide-range: max err / (K eps sum|a||b|) = {ratio:.3e}")
    assert ratio <= 8
    ctx.close()


@pytest.mark.parametrize("kc", ["0", "1", "2", "8"])
def test_promotion_chunk_lengths(kc, monkeypatch):
    """SLICED_GEMM_KC: k-blocks per TMEM chunk before promotion to fp32 registers (0 = one chunk, no promotion).
    Every setting is a correct gemm; shorter chunks are more accurate (round-to-nearest adds between chunks)."""
    import sliced_b200 as S
    monkeypatch.setenv("SLICED_GEMM_KC", kc)
    ctx = S.Context(0)
    for (m, n, k) in [(256, 512, 4096), (300, 700, 1000), (128, 256, 64)]:
        rng = np.random.default_rng(k)
        a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
        t = truth(0, 0, m, n, k, a, b)
        err = np.max(np.abs(run(ctx, 0, 0, m, n, k, a, b, S.GEMM_3XTF32) - t))
        print(f"kc={kc} {m}x{n}x{k} err={err:.3e}")
        assert err <= (4 * k * 2.0 ** -24 if kc != "0" else 16 * k * 2.0 ** -24), (kc, m, n, k, err)
    ctx.close()


def test_accumulate_and_set_semantics(ctx):
    """gemm is SET even into a junk-filled (Cached) buffer; accumulate=1 adds (the OpenCL reference's gemm_grad)"""
    import sliced_b200 as S
    m, n, k = 300, 260, 128
    rng = np.random.default_rng(0)
    a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
    c0 = rng.uniform(-5, 5, m * n).astype(np.float32)
    t = truth(0, 0, m, n, k, a, b)
    got = run(ctx, 0, 0, m, n, k, a, b, S.GEMM_3XTF32, c0, False)
    assert np.max(np.abs(got - t)) <= 4 * k * 2.0 ** -24
    got = run(ctx, 0, 0, m, n, k, a, b, S.GEMM_3XTF32, c0, True)
    assert np.max(np.abs(got - (t + c0))) <= 4 * k * 2.0 ** -24 + 2.0 ** -21


def test_gemm_grad_matches_oracle(ctx):
    import sliced_b200 as S
    m, k, n = 384, 200, 136
    rng = np.random.default_rng(11)
    lhs, rhs = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
    og = rng.uniform(-1, 1, m * n).astype(np.float32)
    lg_ref, rg_ref = np.zeros_like(lhs), np.zeros_like(rhs)
    O.gemm_grad(m, k, n, lhs, rhs, lg_ref, rg_ref, og)
    for mode in (S.GEMM_3XTF32, S.GEMM_SIMT):
        dl, dr = ctx.array(rng.uniform(-3, 3, m * k).astype(np.float32)), ctx.array(rng.uniform(-3, 3, k * n).astype(np.float32))
        ctx.gemm_grad(m, k, n, ctx.array(lhs), ctx.array(rhs), dl, dr, ctx.array(og), False, mode)  # SET over junk
        assert np.max(np.abs(dl.numpy() - lg_ref)) <= 4 * n * 2.0 ** -24 * 2
        assert np.max(np.abs(dr.numpy() - rg_ref)) <= 4 * m * 2.0 ** -24 * 2
    # requires_grad() == false on one side: NULL pointer skips it
    dl = ctx.array(np.zeros(m * k, np.float32))
    ctx.gemm_grad(m, k, n, ctx.array(lhs), ctx.array(rhs), dl, None, ctx.array(og), False, S.GEMM_3XTF32)
    assert np.max(np.abs(dl.numpy() - lg_ref)) <= 4 * n * 2.0 ** -24 * 2


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
def test_plane_scope_is_bit_identical_and_tracks_overwrites(mode):
    """sl_gemm_scope_begin/_end: operand planes are reused across gemms on the same buffer; results must be bit-identical to
    the unscoped calls, a gemm that overwrites a cached buffer must invalidate it, and the second iteration of the same call
    sequence (slots reused, no reallocation) must agree too."""
    import sliced_b200 as S
    md = S.GEMM_3XTF32 if mode == "3xtf32" else S.GEMM_TF32
    ctx = S.Context(0)
    m = k = n = 512  # big enough for the 2-CTA kernel with MN-major operands
    rng = np.random.default_rng(5)
    x = ctx.array(rng.uniform(-1, 1, m * k).astype(np.float32))
    w = ctx.array(rng.uniform(-1, 1, k * n).astype(np.float32))
    g = ctx.array(rng.uniform(-1, 1, m * n).astype(np.float32))

    def sequence():
        y = ctx.gemm(m, k, n, x, w, mode=md)          # x: A K-major, w: B MN-major
        dw = ctx.gemm_tn(k, n, m, x, g, mode=md)      # x: A MN-major (same flat planes), g: B MN-major
        dx = ctx.gemm_nt(m, k, n, g, w, mode=md)      # g: A K-major, w: B K-major
        y2 = ctx.gemm(m, n, n, y, w, mode=md)         # y was WRITTEN by a gemm above, then read
        ctx.gemm(m, k, n, x, w, out=y, mode=md)       # overwrite y (same values), planes of y must be re-derived next
        ctx.gemm(m, k, n, g, w, out=y, mode=md)       # now different values in y
        y3 = ctx.gemm(m, n, n, y, w, mode=md)
        return [t.numpy().copy() for t in (dw, dx, y2, y3)]

    plain = sequence()
    for _ in range(2):
        ctx.gemm_scope_begin()
        scoped = sequence()
        ctx.gemm_scope_end()
        for p, s in zip(plain, scoped):
            assert np.array_equal(p, s)
    # gemm_grad (implicit scope) == the two explicit gemms
    dl, dr = ctx.empty(m * k, np.float32), ctx.empty(k * n, np.float32)
    ctx.gemm_grad(m, k, n, x, w, dl, dr, g, False, md)
    assert np.array_equal(dl.numpy(), plain[1])
    assert np.array_equal(dr.numpy(), plain[0])
    ctx.close()


def test_3xtf32_is_fp32_class_on_ill_scaled_data(ctx):
    """values spanning 2^-10..2^10: plain TF32 loses ~3 decimal digits, 3xTF32 must not"""
    import sliced_b200 as S
    m = n = 256; k = 512
    rng = np.random.default_rng(3)
    a = (rng.uniform(-1, 1, m * k) * 2.0 ** rng.integers(-10, 10, m * k)).astype(np.float32)
    b = (rng.uniform(-1, 1, k * n) * 2.0 ** rng.integers(-10, 10, k * n)).astype(np.float32)
    t = truth(0, 0, m, n, k, a, b)
    scale = np.abs(a.reshape(m, k)).astype(np.float64) @ np.abs(b.reshape(k, n)).astype(np.float64)
    e3 = np.max(np.abs(run(ctx, 0, 0, m, n, k, a, b, S.GEMM_3XTF32) - t) / scale.ravel())
    e1 = np.max(np.abs(run(ctx, 0, 0, m, n, k, a, b, S.GEMM_TF32) - t) / scale.ravel())
    es = np.max(np.abs(O.gemm_ex(0, 0, m, n, k, a, b) - t) / scale.ravel())
    assert e3 <= 2 * es + 2.0 ** -23, (e3, e1, es)  # within 2x of the CPU sgemm's own error (componentwise relative)
    assert e1 > 16 * e3, "TF32 fast mode is not supposed to be this accurate: is the 3-term path really different?"


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
def test_full_size_sampled(ctx, mode):
    """BASELINE gemm sweep top size region: 8192^3 checked on a random sample of output entries against fp64 dot products"""
    import sliced_b200 as S
    m = n = k = 8192
    md = S.GEMM_3XTF32 if mode == "3xtf32" else S.GEMM_TF32
    rng = np.random.default_rng(42)
    a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
    da, db = ctx.array(a), ctx.array(b)
    c = ctx.gemm(m, k, n, da, db, mode=md).numpy().reshape(m, n)
    rows, cols = rng.integers(0, m, 48), rng.integers(0, n, 48)
    rows[:4] = [0, m - 1, 127, 128]; cols[:4] = [0, n - 1, 255, 256]
    t = a.reshape(m, k)[rows].astype(np.float64) @ b.reshape(k, n)[:, cols].astype(np.float64)
    err = np.max(np.abs(c[np.ix_(rows, cols)] - t))
    assert err <= (4 * k * 2.0 ** -24 if mode == "3xtf32" else 8 * np.sqrt(k) * 2.0 ** -11), err
    # linearity: (2A)B == 2(AB) exactly (power-of-two scaling commutes with every rounding step)
    c2 = ctx.gemm(m, k, n, ctx.array(a * 2), db, mode=md).numpy().reshape(m, n)
    assert np.array_equal(c2[rows], 2 * c[rows])


@pytest.mark.parametrize("case", [("NN", 0, 0, 1000, 10, 4096), ("NN", 0, 0, 257, 16, 300), ("NN", 0, 0, 4096, 1, 1000),
                                  ("TN", 1, 0, 512, 10, 8192), ("TN", 1, 0, 1030, 7, 5000), ("TN", 1, 0, 4096, 10, 2048),
                                  ("NT", 0, 1, 1000, 512, 10), ("NT", 0, 1, 333, 1026, 16), ("NT", 0, 1, 64, 64, 3),
                                  # odd small extents (zero-padded column), B too large for shared memory (v1 fallback), ragged row tails
                                  ("NN", 0, 0, 1001, 7, 512), ("NN", 0, 0, 300, 16, 16384), ("NN", 0, 0, 8200, 10, 4096),
                                  ("TN", 1, 0, 1028, 7, 5000), ("TN", 1, 0, 4096, 10, 65536), ("NT", 0, 1, 777, 1028, 9),
                                  ("NT", 0, 1, 5000, 4096, 10),
                                  # few output tiles over a deep contraction (sine_net's weight gradients): split-K CUDA-core kernel
                                  ("TN", 1, 0, 64, 64, 1000), ("TN", 1, 0, 1, 64, 1000), ("TN", 1, 0, 64, 1, 1000), ("TN", 1, 0, 200, 130, 5000),
                                  ("TN", 1, 0, 64, 64, 65536), ("TN", 1, 0, 67, 129, 300)],
                         ids=lambda c: f"{c[0]}-{c[3]}x{c[4]}x{c[5]}")
def test_skinny_shapes(case, monkeypatch):
    """the HBM-bound skinny kernels (the 10-class head of the nn.rs MLP): fp32 FMA accumulation, K-scaled tolerance"""
    import sliced_b200 as S
    monkeypatch.delenv("SLICED_GEMM_TC_FORCE", raising=False)   # default dispatch: these shapes never reach the tensor cores
    ctx = S.Context(0)
    _, ta, tb, m, n, k = case
    rng = np.random.default_rng(m + n + k)
    a, b = rng.uniform(-1, 1, m * k).astype(np.float32), rng.uniform(-1, 1, k * n).astype(np.float32)
    c0 = rng.uniform(-1, 1, m * n).astype(np.float32)
    t = truth(ta, tb, m, n, k, a, b)
    tol = 4 * k * 2.0 ** -24
    got = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_3XTF32)
    assert np.max(np.abs(got - t)) <= tol
    got = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_3XTF32, c0, True)
    assert np.max(np.abs(got - (t + c0))) <= tol + 2.0 ** -22
    exact = run(ctx, ta, tb, m, n, k, a, b, S.GEMM_SIMT)          # SIMT mode stays bit-identical to the oracle
    assert np.array_equal(exact, O.gemm_ex(ta, tb, m, n, k, a, b))
    ctx.close()


# ------------------------------------------------------------------------------------------------ the shapes the numbers are quoted on
def _rand32(rng, n, lo=-1.0, hi=1.0):
    return (rng.random(n, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def _sample_idx(rng, m, n, count=40):
    rows, cols = rng.integers(0, m, count), rng.integers(0, n, count)
    rows[:6] = [0, m - 1, 127, 128, 255, min(256, m - 1)]
    cols[:6] = [0, n - 1, 255, min(256, n - 1), 127, 128]
    return rows, cols


def _check_sampled(got, A64, B64, k, rows, cols, extra=None, what=""):
    """got[rows x cols] against fp64 dot products, 4 K 2^-24 and <= 16x the error of the oracle's own sequential fp32 sgemm on the
    same sampled entries (A64: [len(rows) x k], B64: [k x len(cols)] as float64 views of the fp32 operands)"""
    t = A64 @ B64
    if extra is not None:
        t = t + extra
    err = np.max(np.abs(got.astype(np.float64) - t))
    a32, b32 = np.ascontiguousarray(A64.astype(np.float32)).ravel(), np.ascontiguousarray(B64.astype(np.float32)).ravel()
    ref = O.gemm_ex(0, 0, len(rows), len(cols), k, a32, b32).reshape(len(rows), len(cols)).astype(np.float64)
    ref_err = np.max(np.abs(ref - A64 @ B64))
    assert err <= 4 * k * 2.0 ** -24, f"{what}: err {err} > 4K eps (oracle sgemm err {ref_err})"
    assert err <= 16 * ref_err + 2.0 ** -22, f"{what}: err {err} vs oracle sgemm err {ref_err}"
    return err, ref_err


MODES = {"3xf16": 3, "3xtf32": 0, "tf32": 1}


@pytest.mark.parametrize("mode", ["3xf16", "3xtf32"])
def test_headline_mlp_shapes_sampled(mode, monkeypatch):
    """BASELINE configs[4] (nn.rs MLP 4096-4096-4096-10, batch 65536): the three tensor-core products of one layer exactly as the
    fused step issues them — forward 65536x4096x4096 NN with bias + relu outputs, input gradient NT with the relu mask, weight
    gradient 4096x4096x65536 TN (K = 65536: the split-K path) with the fused bias gradient — sampled against fp64.
    3xf16 runs STRICT (no silent 3xTF32 substitution)."""
    import sliced_b200 as S
    monkeypatch.delenv("SLICED_GEMM_TC_FORCE", raising=False)   # default dispatch: this is what the benchmark runs
    if mode == "3xf16":
        monkeypatch.setenv("SLICED_GEMM_F16_STRICT", "1")
    ctx = S.Context(0)
    L, md = ctx.lib, MODES[mode]
    batch, d = 65536, 4096
    rng = np.random.default_rng(65536)
    x, w, bias = _rand32(rng, batch * d, 0, 1), _rand32(rng, d * d, -0.03, 0.03), _rand32(rng, d)
    dx, dw, db = ctx.array(x), ctx.array(w), ctx.array(bias)
    X, W = x.reshape(batch, d), w.reshape(d, d)
    rows, cols = _sample_idx(rng, batch, d)
    # ---- forward: z = x W + b, a = relu(z)
    z, a = ctx.empty(batch * d), ctx.empty(batch * d)
    S.capi.check(ctx.h, L.sl_linear_fwd(ctx.h, S.F32, batch, d, d, dx.ptr, dw.ptr, db.ptr, z.ptr, a.ptr, md))
    zh = z.numpy().reshape(batch, d)
    _check_sampled(zh[np.ix_(rows, cols)], X[rows].astype(np.float64), W[:, cols].astype(np.float64), d, rows, cols,
                   extra=bias[cols].astype(np.float64)[None, :], what=f"fwd {mode}")
    ah = a.numpy().reshape(batch, d)
    assert np.array_equal(ah[rows], np.where(zh[rows] >= 0, zh[rows], np.float32(0) * zh[rows]))
    del ah
    # ---- input gradient: gx = (z >= 0) * (g W^T)
    g = _rand32(rng, batch * d)
    dg = ctx.array(g)
    G = g.reshape(batch, d)
    gx = ctx.empty(batch * d)
    S.capi.check(ctx.h, L.sl_linear_bwd_input_relu(ctx.h, S.F32, batch, d, d, dw.ptr, dg.ptr, z.ptr, gx.ptr, md))
    gxh = gx.numpy().reshape(batch, d)
    mask = (zh[np.ix_(rows, cols)] >= 0).astype(np.float64)
    got = gxh[np.ix_(rows, cols)]
    assert np.all(got[mask == 0] == 0)
    t = G[rows].astype(np.float64) @ W[cols].astype(np.float64).T
    err = np.max(np.abs(got - mask * t))
    assert err <= 4 * d * 2.0 ** -24 * 0.03 * 4, f"dX {mode}: {err}"
    del gxh, gx
    # ---- weight + bias gradient: dW = x^T g (K = batch), db += colsum(g)
    wg, bg = ctx.empty(d * d), ctx.zeros(d)
    S.capi.check(ctx.h, L.sl_linear_bwd_params(ctx.h, S.F32, batch, d, d, dx.ptr, dg.ptr, wg.ptr, bg.ptr, md))
    wgh = wg.numpy().reshape(d, d)
    r2, c2 = _sample_idx(rng, d, d)
    _check_sampled(wgh[np.ix_(r2, c2)], X[:, r2].T.astype(np.float64), G[:, c2].astype(np.float64), batch, r2, c2, what=f"dW {mode}")
    bsum = G.astype(np.float64).sum(0)
    assert np.max(np.abs(bg.numpy() - bsum)) <= 4 * batch * 2.0 ** -24
    # the plain Tgemm entry point gives the same weights bit for bit
    wg2 = ctx.gemm_tn(d, d, batch, dx, dg, mode=md)
    assert np.array_equal(wg2.numpy().reshape(d, d)[r2], wgh[r2])
    ctx.close()


@pytest.mark.parametrize("size", [4096, 8192, 16384])
@pytest.mark.parametrize("mode", ["3xf16", "3xtf32", "tf32"])
def test_cube_sweep_sizes_sampled(size, mode, monkeypatch):
    """BASELINE configs[2] (gemm sweep M=N=K up to 16384): gemm, gemmT and Tgemm at the large sizes in all three modes, sampled
    against fp64 (4 K 2^-24 and <= 16x the oracle sgemm's error; TF32 fast mode: 8 sqrt(K) 2^-11)."""
    import sliced_b200 as S
    monkeypatch.delenv("SLICED_GEMM_TC_FORCE", raising=False)
    if mode == "3xf16":
        monkeypatch.setenv("SLICED_GEMM_F16_STRICT", "1")
    ctx = S.Context(0)
    md = MODES[mode]
    n = size
    rng = np.random.default_rng(42 + size)
    a, b = _rand32(rng, n * n), _rand32(rng, n * n)
    da, db = ctx.array(a), ctx.array(b)
    A, B = a.reshape(n, n), b.reshape(n, n)
    rows, cols = _sample_idx(rng, n, n)
    for name, ta, tb in (("NN", 0, 0), ("NT", 0, 1), ("TN", 1, 0)):
        c = ctx.gemm_ex(ta, tb, n, n, n, da, db, mode=md).numpy().reshape(n, n)
        A64 = (A[:, rows].T if ta else A[rows]).astype(np.float64)
        B64 = (B[cols].T if tb else B[:, cols]).astype(np.float64)
        got = c[np.ix_(rows, cols)]
        if mode == "tf32":
            err = np.max(np.abs(got - A64 @ B64))
            assert err <= 8 * np.sqrt(n) * 2.0 ** -11, (name, err)
        else:
            _check_sampled(got, A64, B64, n, rows, cols, what=f"{name} {size} {mode}")
        del c
    ctx.close()


@pytest.mark.parametrize("mode", ["3xtf32", "3xf16"])
def test_plane_scope_sees_non_gemm_writes(mode, monkeypatch):
    """Inside a gemm scope cached operand planes / column scales must die when ANY library call rewrites the buffer (element-wise
    op in place, sl_write, sl_clear, sgd) — not only when a gemm does: results stay bit-identical to the unscoped sequence."""
    import sliced_b200 as S
    monkeypatch.delenv("SLICED_GEMM_TC_FORCE", raising=False)
    md = MODES[mode]
    ctx = S.Context(0)
    m, k, n = 2048, 2304, 2560   # >= 37 pair tiles: 2-CTA kernel, MN-major operands, 3xFP16 eligible
    rng = np.random.default_rng(8)
    xh, wh, gh = _rand32(rng, m * k), _rand32(rng, k * n), _rand32(rng, m * n)
    x2h = _rand32(rng, m * k)

    def sequence(scoped):
        x, w, g = ctx.array(xh), ctx.array(wh), ctx.array(gh)
        if scoped:
            ctx.gemm_scope_begin()
        outs = [ctx.gemm(m, k, n, x, w, mode=md)]            # x row-split (leaves its column scales in the scope), w cached
        ctx.unary(S.UN_MUL_SCALAR, x, 3.0, 0.0, out=x)        # in-place element-wise write to x
        outs.append(ctx.gemm_tn(k, n, m, x, g, mode=md))      # must see 3x (column scales of the OLD x are stale)
        outs.append(ctx.gemm(m, k, n, x, w, mode=md))
        x.write(x2h)                                          # sl_write
        outs.append(ctx.gemm(m, k, n, x, w, mode=md))
        ctx.sgd_step(w, ctx.array(wh), 0.5)                   # w -= 0.5 w
        outs.append(ctx.gemm(m, k, n, x, w, mode=md))
        x.clear()                                             # sl_clear
        outs.append(ctx.gemm(m, k, n, x, w, mode=md))
        if scoped:
            ctx.gemm_scope_end()
        return [o.numpy() for o in outs]
    plain, scoped = sequence(False), sequence(True)
    for i, (p, s) in enumerate(zip(plain, scoped)):
        assert np.array_equal(p, s), i
    assert np.all(plain[-1] == 0)
    ctx.close()


@pytest.mark.parametrize("m,k,n", [(8192, 4096, 4096),    # 2-CTA kernel, bits in the epilogue (the MLP's shapes at 1/8 of the batch)
                                   (4096, 16384, 4096),   # split-K: the fold kernel takes / applies the bits
                                   (768, 512, 384),       # single-CTA tensor-core kernel: element-wise bit passes
                                   (100, 70, 50),         # CUDA-core kernel, ragged width (partial last mask word)
                                   (4096, 10, 4096),      # k = 10: forward through the skinny kn kernel
                                   (3000, 4096, 10)])     # n = 10: the head; its input gradient runs the skinny nt kernel with fused bits
@pytest.mark.parametrize("mode", ["3xf16", "3xtf32"])
def test_linear_mask_bits_variants_are_bit_identical(m, k, n, mode, monkeypatch):
    """sl_linear_fwd_bits / sl_linear_bwd_input_relu_bits (relu mask as one bit per element) against sl_linear_fwd /
    sl_linear_bwd_input_relu (mask = the pre-activation matrix): same activations, same input gradients, bit for bit, and the
    bits are exactly (z >= 0) — src/matrix.rs:181-188.  Default dispatch, every path the bits can take."""
    import ctypes as C
    import sliced_b200 as S
    monkeypatch.delenv("SLICED_GEMM_TC_FORCE", raising=False)
    ctx = S.Context(0)
    L, md = ctx.lib, MODES[mode]
    rng = np.random.default_rng(m + k + n)
    x, w, bias = _rand32(rng, m * k, 0, 1), _rand32(rng, k * n, -0.05, 0.05), _rand32(rng, n, -0.5, 0.5)
    if n >= 64:
        bias[:7] = 0; w.reshape(k, n)[:, :7] = 0   # exact zeros in z: the mask's >= matters
    dx, dw, db = ctx.array(x), ctx.array(w), ctx.array(bias)
    z, a, a2 = ctx.empty(m * n), ctx.empty(m * n), ctx.empty(m * n)
    nw = (n + 31) // 32
    bits = ctx.zeros(m * nw, np.int32)
    S.capi.check(ctx.h, L.sl_linear_fwd(ctx.h, S.F32, m, k, n, dx.ptr, dw.ptr, db.ptr, z.ptr, a.ptr, md))
    S.capi.check(ctx.h, L.sl_linear_fwd_bits(ctx.h, S.F32, m, k, n, dx.ptr, dw.ptr, db.ptr, a2.ptr, bits.ptr, md))
    zh = z.numpy().reshape(m, n)
    assert np.array_equal(a.numpy().view(np.uint32), a2.numpy().view(np.uint32))
    want = np.zeros((m, nw * 32), bool)
    want[:, :n] = zh >= 0
    want = np.packbits(want.reshape(m, nw, 32), axis=-1, bitorder="little").view(np.uint32).reshape(m, nw)
    assert np.array_equal(bits.numpy().view(np.uint32).reshape(m, nw), want)
    if n >= 64:
        assert np.all(zh[:, :7] == 0) and np.all(want[:, 0] & 0x7F == 0x7F)
    # input gradient of the NEXT layer's product through this relu: gx[m x n] = (z >= 0) * (g[m x p] W2[n x p]^T)
    for p_out in (10, 256):
        w2, g = _rand32(rng, n * p_out, -0.05, 0.05), _rand32(rng, m * p_out)
        dw2, dg = ctx.array(w2), ctx.array(g)
        gx, gx2 = ctx.empty(m * n), ctx.empty(m * n)
        S.capi.check(ctx.h, L.sl_linear_bwd_input_relu(ctx.h, S.F32, m, n, p_out, dw2.ptr, dg.ptr, z.ptr, gx.ptr, md))
        S.capi.check(ctx.h, L.sl_linear_bwd_input_relu_bits(ctx.h, S.F32, m, n, p_out, dw2.ptr, dg.ptr, bits.ptr, gx2.ptr, md))
        assert np.array_equal(gx.numpy().view(np.uint32), gx2.numpy().view(np.uint32)), p_out
    ctx.close()


def test_rowplane_reuse_inside_a_scope(monkeypatch):
    """3xFP16 inside a gemm scope: the row-scaled planes a forward gemm x[M x I] * w leaves behind serve the weight-gradient gemm
    x^T g of the same scope (which contracts over x's ROWS) — x is not split a second time; its row scales are folded, as exact
    powers of two, into the split of g.  Checked: (1) fp32-class against fp64 like every other gemm; (2) close to the same product
    computed outside any scope (per-column split of x); (3) sl_gemm_grad's mirror image (out_grad's planes reused as the B operand)
    and the bias-gradient column sums of sl_linear_bwd_params; (4) scaling rows of x by powers of two and the same rows of g by the
    inverse leaves the result BIT-identical (both factors disappear into the row scales)."""
    import sliced_b200 as S
    monkeypatch.delenv("SLICED_GEMM_TC_FORCE", raising=False)
    monkeypatch.setenv("SLICED_GEMM_F16_STRICT", "1")
    ctx = S.Context(0)
    L = ctx.lib
    M, I, Oo = 8192, 2048, 2048
    rng = np.random.default_rng(77)
    x, w, g = _rand32(rng, M * I, 0, 1), _rand32(rng, I * Oo, -0.05, 0.05), _rand32(rng, M * Oo, -1e-3, 1e-3)
    x.reshape(M, I)[5] = 0                       # an all-zero row (scale 1)
    x.reshape(M, I)[7] *= np.float32(2.0 ** -30)  # a row far below the others
    dx, dw, dg = ctx.array(x), ctx.array(w), ctx.array(g)
    X, G = x.reshape(M, I), g.reshape(M, Oo)
    rows, cols = _sample_idx(rng, I, Oo)
    plain = ctx.gemm_tn(I, Oo, M, dx, dg).numpy().reshape(I, Oo)                  # no scope: x split per column
    ctx.gemm_scope_begin()
    ctx.gemm(M, I, Oo, dx, dw)                                                    # forward: leaves x's row planes
    n0 = ctx.launches
    reused = ctx.gemm_tn(I, Oo, M, dx, dg).numpy().reshape(I, Oo)
    assert ctx.launches - n0 <= 5, "x must not be split again (colmax + colscale + cols of g, the gemm, at most a fold)"
    bg = ctx.zeros(Oo)
    wg = ctx.empty(I * Oo)
    S.capi.check(ctx.h, L.sl_linear_bwd_params(ctx.h, S.F32, M, I, Oo, dx.ptr, dg.ptr, wg.ptr, bg.ptr, -1))
    S.capi.check(ctx.h, L.sl_gemm_scope_end(ctx.h))
    for got, what in ((plain, "plain"), (reused, "reused")):
        _check_sampled(got[np.ix_(rows, cols)], X[:, rows].T.astype(np.float64), G[:, cols].astype(np.float64), M, rows, cols, what=what)
    scale = np.max(np.abs(plain))
    assert np.max(np.abs(reused - plain)) <= 4 * M * 2.0 ** -24 * 1e-3, np.max(np.abs(reused - plain)) / scale
    assert np.array_equal(wg.numpy().reshape(I, Oo), reused)
    assert np.max(np.abs(bg.numpy() - G.astype(np.float64).sum(0))) <= 4 * M * 2.0 ** -24 * 1e-3
    # mirror image: sl_gemm_grad splits out_grad per row for lhs_grad and reuses those planes as the B operand of rhs_grad
    lg, rg = ctx.empty(M * I), ctx.empty(I * Oo)
    ctx.gemm_grad(M, I, Oo, dx, dw, lg, rg, dg)
    _check_sampled(rg.numpy().reshape(I, Oo)[np.ix_(rows, cols)], X[:, rows].T.astype(np.float64), G[:, cols].astype(np.float64), M, rows, cols,
                   what="gemm_grad rhs_grad")
    r3, c3 = _sample_idx(rng, M, I)
    W = w.reshape(I, Oo)
    t = G[r3].astype(np.float64) @ W[c3].astype(np.float64).T
    assert np.max(np.abs(lg.numpy().reshape(M, I)[np.ix_(r3, c3)] - t)) <= 4 * Oo * 2.0 ** -24 * 0.05 * 1e-3 * 4
    # power-of-two row scalings cancel exactly
    p = np.exp2(rng.integers(-20, 21, M)).astype(np.float32)
    dx2, dg2 = ctx.array((X * p[:, None]).ravel()), ctx.array((G / p[:, None]).ravel())
    ctx.gemm_scope_begin()
    ctx.gemm(M, I, Oo, dx2, dw)
    reused2 = ctx.gemm_tn(I, Oo, M, dx2, dg2).numpy().reshape(I, Oo)
    S.capi.check(ctx.h, L.sl_gemm_scope_end(ctx.h))
    assert np.array_equal(reused2.view(np.uint32), reused.view(np.uint32))
    ctx.close()
