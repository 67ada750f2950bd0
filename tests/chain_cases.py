"""Random element-wise expression graphs for the chain-fuser tests, with their unfused reference evaluation through the oracle's
single-op loops (forward op by op; backward = the tape replayed in reverse registration order into zero-initialised intermediate
gradients, exactly what the reference's closures do — src/ops.rs:115-187, 25-77)."""
import numpy as np

import oracle as O

EXACT_UN = [(O.UN_SQUARE, 0, 0), (O.UN_RELU, 0, 0), (O.UN_NEG, 0, 0), (O.UN_MUL_SCALAR, 0.75, 0), (O.UN_ADD_SCALAR, -0.5, 0), (O.UN_CLIP, -0.6, 0.8)]
LIBM_UN = [(O.UN_TANH, 0, 0), (O.UN_SIGMOID, 0, 0), (O.UN_EXP, 0, 0), (O.UN_POW, 3.0, 0), (O.UN_POW, 2.0, 0)]


def random_graph(rng, n_leaves, n_ops, libm=False, div=False):
    """nodes: list of ('in', k) | ('bin', op, a, b) | ('un', op, a, p0, p1); ids index into the list"""
    nodes = [("in", k) for k in range(n_leaves)]
    uns = EXACT_UN + (LIBM_UN if libm else [])
    bins = [O.ADD, O.SUB, O.MUL] + ([O.DIV] if div else [])
    for _ in range(n_ops):
        if rng.random() < 0.6:
            a, b = int(rng.integers(0, len(nodes))), int(rng.integers(0, len(nodes)))
            nodes.append(("bin", bins[int(rng.integers(0, len(bins)))], a, b))
        else:
            op, p0, p1 = uns[int(rng.integers(0, len(uns)))]
            nodes.append(("un", op, int(rng.integers(0, len(nodes))), p0, p1))
    return nodes


def build_chain(nodes):
    from sliced_b200.chain import Chain
    ch = Chain()
    ex = []
    for nd in nodes:
        if nd[0] == "in":
            ex.append(ch.input())
        elif nd[0] == "bin":
            ex.append(ex[nd[2]]._bin(nd[1], ex[nd[3]]))
        else:
            ex.append(ex[nd[2]].unary(nd[1], nd[3], nd[4]))
    return ch, ex


def forward_unfused(nodes, leaves):
    vals = []
    for nd in nodes:
        if nd[0] == "in":
            vals.append(leaves[nd[1]])
        elif nd[0] == "bin":
            vals.append(O.binary_ew(nd[1], vals[nd[2]], vals[nd[3]]))
        else:
            vals.append(O.unary(nd[1], vals[nd[2]], nd[3], nd[4]))
    return vals


def backward_unfused(nodes, vals, seeds, leaf_grads):
    """seeds: {node id: out_grad array}; leaf_grads: {leaf node id: grad array, updated in place}; returns the grads dict"""
    grads = {}
    for v, nd in enumerate(nodes):
        if nd[0] == "in":
            grads[v] = leaf_grads.get(v)
        else:
            grads[v] = seeds[v].copy() if v in seeds else None
    for v in range(len(nodes) - 1, -1, -1):
        nd = nodes[v]
        if nd[0] == "in" or grads[v] is None:
            continue

        def want(u):
            if nodes[u][0] == "in":
                return grads[u] is not None
            if grads[u] is None:
                grads[u] = np.zeros_like(vals[u])
            return True
        if nd[0] == "bin":
            a, b = nd[2], nd[3]
            wa, wb = want(a), want(b)
            O.binary_ew_grad(nd[1], vals[a], vals[b], grads[a] if wa else None, grads[b] if wb else None, grads[v])
        elif want(nd[2]):
            O.unary_grad(nd[1], vals[nd[2]], grads[nd[2]], grads[v], nd[3], nd[4])
    return grads
