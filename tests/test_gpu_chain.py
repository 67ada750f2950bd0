"""
sl_fused_chain on the GPU (through the C ABI): the micro-op interpreter against the oracle's instruction-by-instruction replay of the
same program — BIT-EXACT for programs of single-IEEE-op instructions, rel 1e-6 where libm is involved — plus the identities that
make it the general form of the fixed-function kernels: sl_chained_fwd / sl_chained_bwd, sl_unary / sl_unary_grad and
sl_binary_ew / sl_binary_ew_grad are each reproduced bit for bit by the corresponding program.
"""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from tests import chain_cases as CC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import sliced_b200 as S
    return S.Context(0)


def bits_equal(a, b):
    u = np.uint32 if a.dtype.itemsize == 4 else np.uint64
    return np.array_equal(a.view(u), b.view(u))


@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("n", [1, 5, 4099, 262147])
def test_random_programs_match_oracle_replay(ctx, seed, n):
    from sliced_b200 import chain as CH
    rng = np.random.default_rng(1000 + seed)
    libm = seed % 3 == 0
    n_leaves, n_ops = int(rng.integers(1, 4)), int(rng.integers(2, 7))
    nodes = CC.random_graph(rng, n_leaves, n_ops, libm=libm, div=seed % 5 == 0)
    dt = np.float64 if seed % 4 == 1 else np.float32
    leaves = [(rng.uniform(0.3, 1.7, n) * (np.ones(n) if seed % 5 == 0 else rng.choice([-1.0, 1.0], n))).astype(dt) for _ in range(n_leaves)]
    vals = CC.forward_unfused(nodes, leaves)
    if not all(np.all(np.isfinite(v)) for v in vals):
        pytest.skip("overflowing random graph")
    ch, ex = CC.build_chain(nodes)
    last = len(nodes) - 1
    outs = list(dict.fromkeys([last] + ([int(rng.integers(n_leaves, last))] if last > n_leaves and seed % 2 else [])))
    wrt = list(range(n_leaves))
    for prog, ins, n_out in ((ch.forward([ex[o] for o in outs]), leaves, len(outs)),
                             (ch.backward([ex[o] for o in outs], [ex[k] for k in wrt]),
                              leaves + [rng.uniform(-1, 1, n).astype(dt) for _ in outs] + [rng.uniform(-1, 1, n).astype(dt) for _ in wrt], None)):
        L = prog.listing()
        n_out = len(L["out_reg"])
        # outputs: for the backward program the first len(wrt) outputs ARE the trailing inputs (in place)
        d_in = [ctx.array(a) for a in ins]
        if len(ins) > n_leaves:
            d_out = d_in[-len(wrt):] + [ctx.zeros(n, dt) for _ in range(n_out - len(wrt))]
            h_out = [a.copy() for a in ins[-len(wrt):]] + [np.zeros(n, dt) for _ in range(n_out - len(wrt))]
        else:
            d_out = [ctx.array(rng.uniform(-9, 9, n).astype(dt)) for _ in range(n_out)]   # junk: SET
            h_out = [np.zeros(n, dt) for _ in range(n_out)]
        CH.run(ctx, prog, d_in, d_out, n)
        scale = np.ones(n)
        O.chain_replay(L, ins, h_out, scale)
        for j, (d, h) in enumerate(zip(d_out, h_out)):
            got = d.numpy()
            if libm:   # CUDA libm vs glibc differ in the last ulps; later instructions amplify that by at most the register magnitudes
                tol = (2e-5 if dt == np.float32 else 1e-12) * np.maximum(scale, np.abs(h))
                assert np.all(np.abs(got.astype(np.float64) - h) <= tol), (seed, j, np.max(np.abs(got - h) / tol))
            else:
                assert bits_equal(got, h), (seed, n, j, np.max(np.abs(got - h)))


def test_chained_perf_programs_equal_the_hand_fused_kernels(ctx):
    """the fuser's programs for examples/chained_perf.rs:86-90 and the fixed-function sl_chained_fwd / sl_chained_bwd: same bits"""
    from sliced_b200 import chain as CH
    ch = CH.Chain()
    x, b = ch.inputs(2)
    squared = x.square()                 # registration order of chained_perf.rs:86-90 (the tape replays it in reverse)
    add = b + x
    mul_b = add * b
    out = squared * x + mul_b
    fwd, bwd = ch.forward([out]), ch.backward([out], [x, b])
    rng = np.random.default_rng(4)
    for n in (7, 1 << 20, (1 << 20) + 3):
        xs, bs, og = (rng.uniform(-2, 2, n).astype(np.float32) for _ in range(3))
        xg0, bg0 = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
        dx, db, dog = ctx.array(xs), ctx.array(bs), ctx.array(og)
        o1, o2 = ctx.chained_fwd(dx, db), ctx.zeros(n)
        CH.run(ctx, fwd, [dx, db], [o2])
        assert bits_equal(o1.numpy(), o2.numpy())
        xg1, bg1, xg2, bg2 = ctx.array(xg0), ctx.array(bg0), ctx.array(xg0), ctx.array(bg0)
        ctx.chained_bwd(dx, db, xg1, bg1, dog)
        CH.run(ctx, bwd, [dx, db, dog, xg2, bg2], [xg2, bg2])
        assert np.array_equal(xg1.numpy(), xg2.numpy()) and np.array_equal(bg1.numpy(), bg2.numpy())
    xs = np.full(16, 1.3, np.float32); bs = np.full(16, 2.1, np.float32)
    o = ctx.zeros(16)
    CH.run(ctx, fwd, [ctx.array(xs), ctx.array(bs)], [o])
    assert np.all(o.numpy().view(np.uint32) == 0x41156459)   # chained_perf.rs:91


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_single_op_programs_equal_the_fixed_function_entry_points(ctx, dt):
    """sl_unary / sl_unary_grad / sl_binary_ew / sl_binary_ew_grad re-expressed as programs: bit-identical"""
    import sliced_b200 as S
    from sliced_b200 import chain as CH
    rng = np.random.default_rng(11)
    n = 10007
    isint = dt == np.int32
    mk = (lambda: rng.integers(1, 40, n).astype(dt)) if isint else (lambda: rng.uniform(0.2, 1.9, n).astype(dt))
    x, y, og, g0, g1 = mk(), mk(), mk(), mk(), mk()
    dx, dy, dog = ctx.array(x), ctx.array(y), ctx.array(og)
    unops = [(S.UN_SQUARE, 0, 0), (S.UN_RELU, 0, 0), (S.UN_CLIP, 0.5, 1.5), (S.UN_NEG, 0, 0), (S.UN_MUL_SCALAR, 3.0, 0), (S.UN_ADD_SCALAR, 2.0, 0),
             (S.UN_NEG_DIV_SCALAR, 4.0, 0)]
    if not isint:
        unops += [(S.UN_POW, 3.0, 0), (S.UN_POW, 2.5, 0), (S.UN_TANH, 0, 0), (S.UN_SIGMOID, 0, 0), (S.UN_EXP, 0, 0), (S.UN_LN, 0, 0), (S.UN_NEG_LN, 0, 0)]
    for op, p0, p1 in unops:
        ch = CH.Chain()
        e = ch.input()
        f = e.unary(op, p0, p1)
        o = ctx.zeros(n, dt)
        CH.run(ctx, ch.forward([f]), [dx], [o])
        assert bits_equal(o.numpy(), ctx.unary(op, dx, p0, p1).numpy()), op
        a, b = ctx.array(g0), ctx.array(g0)
        ctx.unary_grad(op, dx, a, dog, p0, p1)
        CH.run(ctx, ch.backward([f], [e]), [dx, dog, b], [b])
        assert bits_equal(a.numpy(), b.numpy()), op
    for op in (S.ADD, S.SUB, S.MUL, S.DIV):
        ch = CH.Chain()
        l, r = ch.inputs(2)
        f = l._bin(op, r)
        o = ctx.zeros(n, dt)
        CH.run(ctx, ch.forward([f]), [dx, dy], [o])
        assert bits_equal(o.numpy(), ctx.binary_ew(op, dx, dy).numpy()), op
        a0, a1, b0, b1 = ctx.array(g0), ctx.array(g1), ctx.array(g0), ctx.array(g1)
        ctx.binary_ew_grad(op, dx, dy, a0, a1, dog)
        CH.run(ctx, ch.backward([f], [l, r]), [dx, dy, dog, b0, b1], [b0, b1])
        assert np.array_equal(a0.numpy(), b0.numpy()) and np.array_equal(a1.numpy(), b1.numpy()), op


def test_acc_outputs_misaligned_pointers_and_errors(ctx):
    import sliced_b200 as S
    from sliced_b200 import chain as CH
    rng = np.random.default_rng(2)
    n = 5003
    x, y, o0 = (rng.uniform(-1, 1, n + 1).astype(np.float32) for _ in range(3))
    ch = CH.Chain()
    a, b = ch.inputs(2)
    prog = ch.forward([a * b + a])
    prog.out_acc[0] = 1                                 # out = out + r
    dx, dy, do = ctx.array(x), ctx.array(y), ctx.array(o0)
    CH.run(ctx, prog, [dx.view(1, n), dy.view(1, n)], [do.view(1, n)], n)   # 4 bytes into the allocation: scalar path
    ref = o0.copy()
    ref[1:] = ref[1:] + (x[1:] * y[1:] + x[1:])
    assert bits_equal(do.numpy(), ref)
    # integer program with a transcendental opcode is refused, malformed programs are refused
    chi = CH.Chain()
    e = chi.input()
    bad = chi.forward([e.tanh()])
    xi = ctx.array(np.arange(8, dtype=np.int32))
    with pytest.raises(S.SlicedError):
        CH.run(ctx, bad, [xi], [ctx.zeros(8, np.int32)])
    prog2 = ch.forward([a + b])
    prog2.instr[0].a = 23                               # reads a register nobody wrote
    with pytest.raises(S.SlicedError):
        CH.run(ctx, prog2, [dx, dy], [do])
    CH.run(ctx, ch.forward([a + b]), [dx, dy], [do], 0)  # empty: a no-op
