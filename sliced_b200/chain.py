"""
Host face of the op-list fuser (sliced_b200/host/chain_builder.cpp) and of `sl_fused_chain`: build a chain of element-wise ops symbolically,
get the forward / backward micro-op programs, run them in one launch each.

    ch = Chain()
    x, b = ch.inputs(2)
    out = x.square() * x + (b + x) * b                 # examples/chained_perf.rs:86-90
    fwd = ch.forward([out])                            # sl_chain_prog: inputs (x, b) -> out
    bwd = ch.backward(seeds=[out], wrt=[x, b])         # inputs (x, b, out.grad, x.grad, b.grad) -> x.grad, b.grad (in place)

No arithmetic happens here: programs are built by the C++ builder (pure host code, works without a GPU) and executed by the CUDA
interpreter kernel through the C ABI.
"""
from __future__ import annotations

import ctypes as C

from . import capi

MAX_INSTRS, MAX_INPUTS, MAX_OUTPUTS, MAX_REGS = 32, 8, 4, 24
CH_ADD, CH_SUB, CH_MUL, CH_DIV, CH_RDIV_IMM, CH_CONST, CH_COPY, CH_UNARY_F, CH_UNARY_D = 0, 1, 2, 3, 4, 5, 6, 16, 48


class Instr(C.Structure):
    _fields_ = [("op", C.c_uint8), ("dst", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8), ("flags", C.c_uint8), ("pad_", C.c_uint8 * 3),
                ("imm0", C.c_double), ("imm1", C.c_double)]


class Prog(C.Structure):
    _fields_ = [("n_instr", C.c_int32), ("n_in", C.c_int32), ("n_out", C.c_int32), ("n_regs", C.c_int32), ("instr", Instr * MAX_INSTRS),
                ("out_reg", C.c_uint8 * MAX_OUTPUTS), ("out_acc", C.c_uint8 * MAX_OUTPUTS)]

    def listing(self):
        """[(op, dst, a, b, imm0, imm1)], out registers, acc flags — the neutral form the oracle's replay takes"""
        ins = [(i.op, i.dst, i.a, i.b, i.imm0, i.imm1) for i in self.instr[:self.n_instr]]
        return dict(n_in=self.n_in, n_regs=self.n_regs, instr=ins, out_reg=list(self.out_reg[:self.n_out]), out_acc=list(self.out_acc[:self.n_out]))


_declared = False


def _lib():
    global _declared
    lib = capi.load()
    if not _declared:
        P, i, d, vp = C.POINTER, C.c_int, C.c_double, C.c_void_p
        for name, args, res in (("slh_chain_new", [], vp), ("slh_chain_free", [vp], None), ("slh_chain_input", [vp], i),
                                ("slh_chain_binary", [vp, i, i, i], i), ("slh_chain_unary", [vp, i, i, d, d], i),
                                ("slh_chain_build_forward", [vp, P(i), i, P(Prog)], i),
                                ("slh_chain_build_backward", [vp, P(i), i, P(i), i, P(Prog)], i)):
            f = getattr(lib, name)
            f.argtypes, f.restype = args, res
        lib.sl_fused_chain.argtypes = [vp, i, P(Prog), P(vp), P(vp), C.c_size_t]   # refines the untyped binding of capi.py
        lib.sl_fused_chain.restype = i
        _declared = True
    return lib


class Expr:
    def __init__(self, chain: "Chain", vid: int):
        self.chain, self.id = chain, vid

    def _bin(self, op, other, swap=False):
        if not isinstance(other, Expr):
            raise TypeError("chain expressions combine with chain expressions (scalars: mul_scalar / add_scalar)")
        l, r = (other, self) if swap else (self, other)
        return Expr(self.chain, _lib().slh_chain_binary(self.chain.h, op, l.id, r.id))

    def __add__(self, o): return self._bin(capi.ADD, o)
    def __sub__(self, o): return self._bin(capi.SUB, o)
    def __mul__(self, o): return self._bin(capi.MUL, o)
    def __truediv__(self, o): return self._bin(capi.DIV, o)

    def unary(self, unop, p0=0.0, p1=0.0):
        return Expr(self.chain, _lib().slh_chain_unary(self.chain.h, unop, self.id, float(p0), float(p1)))

    def square(self): return self.unary(capi.UN_SQUARE)
    def pow(self, p): return self.unary(capi.UN_POW, p)
    def relu(self): return self.unary(capi.UN_RELU)
    def tanh(self): return self.unary(capi.UN_TANH)
    def sigmoid(self): return self.unary(capi.UN_SIGMOID)
    def exp(self): return self.unary(capi.UN_EXP)
    def ln(self): return self.unary(capi.UN_LN)
    def clip(self, lo, hi): return self.unary(capi.UN_CLIP, lo, hi)
    def neg(self): return self.unary(capi.UN_NEG)
    def mul_scalar(self, s): return self.unary(capi.UN_MUL_SCALAR, s)
    def add_scalar(self, s): return self.unary(capi.UN_ADD_SCALAR, s)


class Chain:
    def __init__(self):
        self.h = _lib().slh_chain_new()

    def __del__(self):
        try:
            if self.h:
                _lib().slh_chain_free(self.h)
                self.h = None
        except Exception:
            pass

    def input(self) -> Expr:
        return Expr(self, _lib().slh_chain_input(self.h))

    def inputs(self, n):
        return [self.input() for _ in range(n)]

    def forward(self, outs) -> Prog:
        ids = (C.c_int * len(outs))(*[o.id for o in outs])
        p = Prog()
        if _lib().slh_chain_build_forward(self.h, ids, len(outs), C.byref(p)) != 0:
            raise capi.SlicedError(capi.SL_ERR_INVALID_ARG, "chain does not fit the interpreter (inputs / outputs / instructions / registers)")
        return p

    def backward(self, seeds, wrt) -> Prog:
        s = (C.c_int * len(seeds))(*[o.id for o in seeds])
        w = (C.c_int * len(wrt))(*[o.id for o in wrt])
        p = Prog()
        if _lib().slh_chain_build_backward(self.h, s, len(seeds), w, len(wrt), C.byref(p)) != 0:
            raise capi.SlicedError(capi.SL_ERR_INVALID_ARG, "backward chain does not fit the interpreter")
        return p


def run(ctx, prog: Prog, inputs, outputs, n=None):
    """sl_fused_chain on a raw Context: inputs / outputs are DeviceArrays (an input may also be an output: in place)"""
    lib = _lib()
    n = outputs[0].size if n is None else n
    ins = (C.c_void_p * max(len(inputs), 1))(*[a.ptr for a in inputs])
    outs = (C.c_void_p * len(outputs))(*[a.ptr for a in outputs])
    capi.check(ctx.h, lib.sl_fused_chain(ctx.h, outputs[0].code, C.byref(prog), ins, outs, n))
