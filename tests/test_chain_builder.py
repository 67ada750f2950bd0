"""
The op-list fuser on the CPU (no GPU needed): programs built by sliced_b200/host/chain_builder.cpp, replayed instruction by
instruction with the oracle's single-op loops, must give what the UNFUSED graph gives — forward values exactly, gradients exactly
(up to the sign of zero: an intermediate gradient that starts as `0 + c` in the reference starts as `c` in a register).
"""
import numpy as np
import pytest

import oracle as O
from tests import chain_cases as CC


def same(a, b):
    return np.array_equal(a, b)   # numeric equality: -0.0 == 0.0


@pytest.mark.parametrize("seed", range(40))
def test_random_chains_forward_and_backward(seed):
    rng = np.random.default_rng(seed)
    n_leaves, n_ops = int(rng.integers(1, 4)), int(rng.integers(2, 7))
    nodes = CC.random_graph(rng, n_leaves, n_ops, libm=seed % 3 == 0, div=seed % 5 == 0)
    n = 257
    dt = np.float64 if seed % 4 == 1 else np.float32
    leaves = [(rng.uniform(0.3, 1.7, n) * (np.ones(n) if seed % 5 == 0 else rng.choice([-1.0, 1.0], n))).astype(dt) for _ in range(n_leaves)]
    vals = CC.forward_unfused(nodes, leaves)
    if not all(np.all(np.isfinite(v)) for v in vals):
        pytest.skip("overflowing random graph")
    ch, ex = CC.build_chain(nodes)
    last = len(nodes) - 1
    outs = [last] + ([int(rng.integers(n_leaves, last))] if last > n_leaves and seed % 2 else [])
    outs = list(dict.fromkeys(outs))
    fwd = ch.forward([ex[o] for o in outs])
    got = [np.zeros(n, dt) for _ in outs]
    O.chain_replay(fwd.listing(), leaves, got)
    for o, g in zip(outs, got):
        assert same(g, vals[o]), (seed, o)
    # backward: every leaf takes a gradient except (sometimes) one; out-grads arrive at every materialised output
    wrt = [k for k in range(n_leaves) if not (seed % 7 == 3 and k == 0 and n_leaves > 1)]
    seeds = {o: rng.uniform(-1, 1, n).astype(dt) for o in outs}
    g0 = {k: rng.uniform(-1, 1, n).astype(dt) for k in wrt}
    ref = CC.backward_unfused(nodes, vals, seeds, {k: g0[k].copy() for k in wrt})
    bwd = ch.backward([ex[o] for o in outs], [ex[k] for k in wrt])
    L = bwd.listing()
    assert L["n_in"] == n_leaves + len(outs) + len(wrt)
    gl = [g0[k].copy() for k in wrt]
    extra = [np.zeros(n, dt) for _ in range(len(L["out_reg"]) - len(wrt))]
    O.chain_replay(L, leaves + [seeds[o] for o in outs] + gl, gl + extra)
    for k, g in zip(wrt, gl):
        assert same(g, ref[k]), (seed, k, np.max(np.abs(g - ref[k])))
    # seeds that are consumed inside the chain report their accumulated total (what `.grad()` of that buffer holds in the reference)
    internal = [o for o in outs if any((nd[0] == "bin" and o in (nd[2], nd[3])) or (nd[0] == "un" and nd[2] == o) for nd in nodes[o + 1:])]
    for o, e in zip([o for o in outs if o in internal], extra):
        assert same(e, ref[o]), (seed, o)


def test_chained_perf_graph_is_13_instructions_in_7_registers():
    """examples/chained_perf.rs:86-90 and its tape: the programs the fuser emits, and the reference's known answer"""
    from sliced_b200.chain import Chain
    ch = Chain()
    x, b = ch.inputs(2)
    squared = x.square()                 # registration order of chained_perf.rs:86-90 (the tape replays it in reverse)
    add = b + x
    mul_b = add * b
    out = squared * x + mul_b
    fwd, bwd = ch.forward([out]), ch.backward([out], [x, b])
    assert (fwd.n_instr, fwd.n_regs, fwd.n_in, fwd.n_out) == (5, 3, 2, 1)
    assert (bwd.n_instr, bwd.n_in, bwd.n_out) == (13, 5, 2) and bwd.n_regs <= 7   # 5 pinned inputs + 2 temporaries in tape order
    xs, bs, o = np.full(5, 1.3, np.float32), np.full(5, 2.1, np.float32), np.zeros(5, np.float32)
    O.chain_replay(fwd.listing(), [xs, bs], [o])
    assert o.view(np.uint32)[0] == 0x41156459          # == 9.336999f, chained_perf.rs:91
    xg, bg, og = np.zeros(5, np.float32), np.zeros(5, np.float32), np.ones(5, np.float32)
    O.chain_replay(bwd.listing(), [xs, bs, og, xg.copy(), bg.copy()], [xg, bg])
    xr, br = np.zeros(5, np.float32), np.zeros(5, np.float32)
    O.chained_bwd(xs, bs, xr, br, og)
    assert same(xg, xr) and same(bg, br)


def test_limits_are_reported_not_truncated():
    from sliced_b200 import SlicedError
    from sliced_b200.chain import Chain
    ch = Chain()
    x = ch.input()
    e = x
    for _ in range(40):
        e = e * x + x
    with pytest.raises(SlicedError):
        ch.forward([e])
    ch2 = Chain()
    ins = ch2.inputs(9)
    acc = ins[0]
    for i in ins[1:]:
        acc = acc + i
    with pytest.raises(SlicedError):
        ch2.forward([acc])
