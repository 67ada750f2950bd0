"""
Python face of the C++ host layer (sliced_b200/host/*.cpp): the reference's device / Buffer / Matrix / *MayGrad
interface with the same names and argument order, so that tests read like the reference's own tests:

    device = CUDA()                                   # custos `CUDA::<Autograd<Base>>::new(0)` †
    lhs = device.buffer([1, 2, 3, 4, 5], np.int32)    # Buffer::from((&device, [..]))
    out = device.add(lhs, rhs)                        # BinaryOpsMayGrad::add
    out.backward()
    lhs.grad().read()

No arithmetic happens here; every call goes host layer -> C ABI -> CUDA kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import SlicedError
from .raw import dtype_code, _NP

_vp, _sz, _i, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
_declared = False


def _lib():
    global _declared
    lib = capi.load()
    if not _declared:
        P = C.POINTER
        sig = {
            "slh_last_error": ([], C.c_char_p),
            "slh_device_new": ([_i, _i], _vp),
            "slh_device_new_on_stream": ([_i, _i, _vp], _vp),
            "slh_device_free": ([_vp], None),
            "slh_device_ctx": ([_vp], _vp),
            "slh_device_sync": ([_vp], _i),
            "slh_range_begin": ([_vp], None),
            "slh_zero_grad": ([_vp], _i),
            "slh_tape_len": ([_vp], _i),
            "slh_n_grads": ([_vp], _i),
            "slh_set_fusion": ([_vp, _i], _i),
            "slh_flush": ([_vp], _i),
            "slh_fused_groups": ([_vp], _i),
            "slh_unfused_groups": ([_vp], _i),
            "slh_set_tape_enabled": ([_vp, _i], None),
            "slh_set_gemm_mode": ([_vp, _i], _i),
            "slh_buffer_new": ([_vp, _sz, _i], _vp),
            "slh_buffer_from_host": ([_vp, _vp, _sz, _i], _vp),
            "slh_buffer_wrap": ([_vp, _vp, _sz, _i], _vp),
            "slh_buffer_release": ([_vp], None),
            "slh_buffer_len": ([_vp], _sz),
            "slh_buffer_dtype": ([_vp], _i),
            "slh_buffer_dptr": ([_vp], _vp),
            "slh_buffer_id": ([_vp], C.c_uint64),
            "slh_buffer_set_requires_grad": ([_vp, _i], None),
            "slh_buffer_requires_grad": ([_vp], _i),
            "slh_buffer_read": ([_vp, _vp], _i),
            "slh_buffer_write": ([_vp, _vp], _i),
            "slh_grad": ([_vp], _vp),
            "slh_backward": ([_vp], _i),
            "slh_backward_with": ([_vp, _vp], _i),
            "slh_op": ([_vp, C.c_char_p, P(_vp), _i, P(_sz), _i, P(_d), _i], _vp),
            "slh_scalar_op": ([_vp, C.c_char_p, _vp, P(_d)], _i),
            "slh_sgd_step": ([_vp, _d], _i),
            "slh_mlp_new": ([_vp, _i, P(_sz), _i], _vp),
            "slh_mlp_free": ([_vp], None),
            "slh_mlp_n_params": ([_vp], _sz),
            "slh_mlp_metrics_dptr": ([_vp], _vp),
            "slh_mlp_weights": ([_vp, _i], _vp),
            "slh_mlp_bias": ([_vp, _i], _vp),
            "slh_mlp_grad_bucket": ([_vp], _vp),
            "slh_mlp_params": ([_vp], _vp),
            "slh_mlp_forward_backward": ([_vp, _vp, _vp, _vp, _sz, _sz, _i, P(_d), P(C.c_longlong)], _i),
            "slh_mlp_set_fused": ([_vp, _i], None),
            "slh_mlp_set_deferred": ([_vp, _i], _i),
            "slh_mlp_flush": ([_vp], _i),
            "slh_mlp_allreduce_grads": ([_vp], _i),
            "slh_mlp_sgd": ([_vp, _d], _i),
            "slh_mlp_step": ([_vp, _vp, _vp, _vp, _sz, _sz, _d, _i, P(_d), P(C.c_longlong)], _i),
            "slh_mlp_step_replay": ([_vp, _vp, _vp, _vp, _sz, _sz, _d, _i, P(_d), P(C.c_longlong)], _i),
            "slh_mlp_predict": ([_vp, _vp, _sz], _vp),
        }
        for name, (args, res) in sig.items():
            f = getattr(lib, name)
            f.argtypes = args
            f.restype = res
        _declared = True
    return lib


def _err():
    msg = _lib().slh_last_error()
    return SlicedError(-1, msg.decode() if msg else "?")


def _chk(rc):
    if rc != 0:
        raise _err()


class Buffer:
    """custos `Buffer<T, CUDA>` †: flat device memory taking part in autograd unless `.no_grad()`."""

    def __init__(self, device: "CUDA", handle):
        if not handle:
            raise _err()
        self.device = device
        self.h = handle

    def __del__(self):
        try:
            if self.h and self.device.h:
                _lib().slh_buffer_release(self.h)
        except Exception:
            pass

    def __len__(self):
        return _lib().slh_buffer_len(self.h)

    @property
    def dtype(self):
        return np.dtype(_NP[_lib().slh_buffer_dtype(self.h)])

    @property
    def ptr(self) -> int:
        return _lib().slh_buffer_dptr(self.h) or 0

    def read(self) -> np.ndarray:
        out = np.empty(len(self), self.dtype)
        if out.size:
            _chk(_lib().slh_buffer_read(self.h, out.ctypes.data))
        return out

    def write(self, host):
        host = np.ascontiguousarray(host, self.dtype).ravel()
        assert host.size == len(self)
        _chk(_lib().slh_buffer_write(self.h, host.ctypes.data))
        return self

    def no_grad(self):
        _lib().slh_buffer_set_requires_grad(self.h, 0)
        return self

    def require_grad(self):
        _lib().slh_buffer_set_requires_grad(self.h, 1)
        return self

    def requires_grad(self) -> bool:
        return bool(_lib().slh_buffer_requires_grad(self.h))

    def grad(self) -> "Buffer":
        return Buffer(self.device, _lib().slh_grad(self.h))

    def backward(self):
        _chk(_lib().slh_backward(self.h))

    def backward_with(self, seed: "Buffer"):
        _chk(_lib().slh_backward_with(self.h, seed.h))


class CUDA:
    """The device: custos `CUDA<Autograd<Base>>` (cached=False) or `CUDA<Autograd<Cached<Base>>>` (cached=True) †."""

    def __init__(self, index: int = 0, cached: bool = False, stream: int | None = None):
        lib = _lib()
        self.h = lib.slh_device_new(index, int(cached)) if stream is None else lib.slh_device_new_on_stream(index, int(cached), C.c_void_p(stream))
        if not self.h:
            self.h = None
            raise _err()

    def close(self):
        if self.h:
            _lib().slh_device_free(self.h)
            self.h = None

    @property
    def ctx_handle(self):
        return _lib().slh_device_ctx(self.h)

    @property
    def launches(self) -> int:
        return int(capi.load().sl_ctx_launch_count(self.ctx_handle))

    def sync(self):
        _chk(_lib().slh_device_sync(self.h))

    # ---- buffers
    def buffer(self, data, dtype=None) -> Buffer:
        host = np.ascontiguousarray(data, dtype=dtype).ravel()
        return Buffer(self, _lib().slh_buffer_from_host(self.h, host.ctypes.data, host.size, dtype_code(host.dtype)))

    def zeros(self, n, dtype=np.float32) -> Buffer:
        return Buffer(self, _lib().slh_buffer_new(self.h, n, dtype_code(dtype)))

    def wrap(self, ptr: int, n: int, dtype=np.float32) -> Buffer:
        return Buffer(self, _lib().slh_buffer_wrap(self.h, C.c_void_p(ptr), n, dtype_code(dtype)))

    # ---- custos Cached/Cursor and Autograd plumbing
    def range(self, n):
        """`for epoch in device.range(0..n)`: rewinds the Cached cursor every iteration (examples/nn.rs:184)."""
        for i in range(n):
            _lib().slh_range_begin(self.h)
            yield i

    def zero_grad(self):
        _chk(_lib().slh_zero_grad(self.h))

    def tape_len(self) -> int:
        return _lib().slh_tape_len(self.h)

    def set_fusion(self, on: bool = True):
        """custos `Lazy` + `optimize()` for element-wise chains: binary / unary ops are recorded and run as ONE sl_fused_chain launch
        when their values are needed (forward), and their grad closures as one fused backward launch."""
        _chk(_lib().slh_set_fusion(self.h, int(on)))
        return self

    def flush(self):
        _chk(_lib().slh_flush(self.h))

    @property
    def fused_groups(self) -> int: return _lib().slh_fused_groups(self.h)
    @property
    def unfused_groups(self) -> int: return _lib().slh_unfused_groups(self.h)

    def n_grads(self) -> int:
        """number of live gradient buffers (a buffer's gradient is released with the buffer)"""
        return _lib().slh_n_grads(self.h)

    def set_gemm_mode(self, mode: int):
        _chk(_lib().slh_set_gemm_mode(self.h, mode))

    # ---- ops
    def _op(self, name, bufs=(), dims=(), scalars=()) -> Buffer:
        B = (_vp * max(len(bufs), 1))(*[b.h for b in bufs])
        D = (_sz * max(len(dims), 1))(*[int(d) for d in dims])
        S = (_d * max(len(scalars), 1))(*[float(s) for s in scalars])
        return Buffer(self, _lib().slh_op(self.h, name.encode(), B, len(bufs), D, len(dims), S, len(scalars)))

    def add(self, lhs, rhs): return self._op("add", (lhs, rhs))
    def add2(self, lhs, rhs): return self._op("add2", (lhs, rhs))
    def sub(self, lhs, rhs): return self._op("sub", (lhs, rhs))
    def mul(self, lhs, rhs): return self._op("mul", (lhs, rhs))
    def div(self, lhs, rhs): return self._op("div", (lhs, rhs))
    def square(self, x): return self._op("square", (x,))
    def pow(self, x, rhs): return self._op("pow", (x,), (), (rhs,))
    def transpose(self, rows, cols, x): return self._op("transpose", (x,), (rows, cols))
    def gemm(self, m, k, n, lhs, rhs): return self._op("gemm", (lhs, rhs), (m, k, n))
    def add_row(self, rows, cols, lhs, rhs): return self._op("add_row", (lhs, rhs), (rows, cols))
    def add_row_mut(self, rows, cols, lhs, rhs): self._op("add_row_mut", (lhs, rhs), (rows, cols))
    def clip(self, x, lo, hi): return self._op("clip", (x,), (), (lo, hi))
    def exp(self, x): return self._op("exp", (x,))
    def max_cols(self, rows, cols, x): return self._op("max_cols", (x,), (rows, cols))
    def max_rows(self, cols, x): return self._op("max_rows", (x,), (cols,))
    def sum_rows(self, cols, x): return self._op("sum_rows", (x,), (cols,))
    def sum_cols(self, cols, x): return self._op("sum_cols", (x,), (cols,))
    def mean_cols(self, cols, x): return self._op("mean_cols", (x,), (cols,))
    def mean_rows(self, cols, x): return self._op("mean_rows", (x,), (cols,))
    def diagflat(self, x): return self._op("diagflat", (x,))
    def softmax(self, samples, features, x): return self._op("softmax", (x,), (samples, features))
    def relu(self, x): return self._op("relu", (x,))
    def tanh(self, x): return self._op("tanh", (x,))
    def sigmoid(self, x): return self._op("sigmoid", (x,))
    def apply_fn(self, x, unop, p0=0.0, p1=0.0): return self._op("apply_fn", (x,), (unop,), (p0, p1))
    def sub_cols(self, cols, lhs, rhs): return self._op("sub_cols", (lhs, rhs), (cols,))
    def div_cols(self, cols, lhs, rhs): return self._op("div_cols", (lhs, rhs), (cols,))
    def onehot(self, classes): return self._op("onehot", (classes,))

    def _scalar(self, name, x):
        out = _d(0)
        _chk(_lib().slh_scalar_op(self.h, name.encode(), x.h, C.byref(out)))
        return out.value

    def sum(self, x): return self._scalar("sum", x)
    def mean(self, x): return self._scalar("mean", x)
    def max(self, x): return self._scalar("max", x)

    def sgd_step(self, param: Buffer, lr: float):
        _chk(_lib().slh_sgd_step(param.h, lr))


class Matrix:
    """sliced `Matrix` (src/matrix.rs:21-25): (Buffer, rows, cols) with method sugar."""

    def __init__(self, device: CUDA, rows: int, cols: int, data=None, dtype=np.float32, buf: Buffer | None = None):
        self.device, self.rows, self.cols = device, rows, cols
        if buf is None:
            buf = device.zeros(rows * cols, dtype) if data is None else device.buffer(data, dtype)
        assert len(buf) == rows * cols, "data.len() != rows * cols"  # matrix/impl_from.rs:8
        self.buf = buf

    def _m(self, buf, rows, cols): return Matrix(self.device, rows, cols, buf=buf)
    def read(self): return self.buf.read()
    def grad(self): return self.buf.grad()
    def backward(self): self.buf.backward()
    def no_grad(self): self.buf.no_grad(); return self
    def require_grad(self): self.buf.require_grad(); return self
    def T(self): return self._m(self.device.transpose(self.rows, self.cols, self.buf), self.cols, self.rows)
    def gemm(self, rhs): return self._m(self.device.gemm(self.rows, self.cols, rhs.cols, self.buf, rhs.buf), self.rows, rhs.cols)
    def add(self, rhs): return self._m(self.device.add(self.buf, rhs.buf), self.rows, self.cols)
    def sub(self, rhs): return self._m(self.device.sub(self.buf, rhs.buf), self.rows, self.cols)
    def mul(self, rhs): return self._m(self.device.mul(self.buf, rhs.buf), self.rows, self.cols)
    def add_row(self, rhs): return self._m(self.device.add_row(self.rows, self.cols, self.buf, rhs.buf), self.rows, self.cols)
    def add_row_mut(self, rhs): self.device.add_row_mut(self.rows, self.cols, self.buf, rhs.buf)
    def relu(self): return self._m(self.device.relu(self.buf), self.rows, self.cols)
    def tanh(self): return self._m(self.device.tanh(self.buf), self.rows, self.cols)
    def sigmoid(self): return self._m(self.device.sigmoid(self.buf), self.rows, self.cols)
    def squared(self): return self._m(self.device.square(self.buf), self.rows, self.cols)
    def pow(self, rhs): return self._m(self.device.pow(self.buf, rhs), self.rows, self.cols)
    def sum_cols(self): return self._m(self.device.sum_cols(self.cols, self.buf), self.rows, 1)
    def l2_norm_cols(self): return self.squared().sum_cols().pow(0.5)
    def diagflat(self): return self._m(self.device.diagflat(self.buf), self.rows, self.rows)
    def softmax(self): return self._m(self.device.softmax(self.rows, self.cols, self.buf), self.rows, self.cols)


LOSS_SOFTMAX_CCE, LOSS_SQUARED = 0, 1


class Mlp:
    """examples/nn.rs (loss=LOSS_SOFTMAX_CCE) / examples/sine_net.rs (loss=LOSS_SQUARED) training loop on the device."""

    def __init__(self, device: CUDA, dims, loss=LOSS_SOFTMAX_CCE):
        self.device, self.dims = device, list(dims)
        arr = (_sz * len(dims))(*dims)
        self.h = _lib().slh_mlp_new(device.h, len(dims), arr, loss)
        if not self.h:
            raise _err()

    def __del__(self):
        try:
            if self.h and self.device.h:
                _lib().slh_mlp_free(self.h)
        except Exception:
            pass

    @property
    def n_layers(self): return len(self.dims) - 1
    @property
    def n_params(self): return _lib().slh_mlp_n_params(self.h)
    @property
    def metrics_ptr(self) -> int:
        """device address of the last step's [loss_sum f32][correct i32] (fetch with sl_read_async to avoid a blocking read per step)"""
        return _lib().slh_mlp_metrics_dptr(self.h) or 0
    def weights(self, l) -> Buffer: return Buffer(self.device, _lib().slh_mlp_weights(self.h, l))
    def bias(self, l) -> Buffer: return Buffer(self.device, _lib().slh_mlp_bias(self.h, l))
    def grad_bucket(self) -> Buffer: return Buffer(self.device, _lib().slh_mlp_grad_bucket(self.h))
    def params(self) -> Buffer: return Buffer(self.device, _lib().slh_mlp_params(self.h))

    def forward_backward(self, x: Buffer, y: Buffer, labels: Buffer | None, batch: int, grad_rows: int | None = None, want_metrics=True):
        loss, correct = _d(0), C.c_longlong(0)
        _chk(_lib().slh_mlp_forward_backward(self.h, x.h, y.h, labels.h if labels else None, batch, grad_rows or batch, int(want_metrics),
                                             C.byref(loss), C.byref(correct)))
        return loss.value, correct.value

    def set_fused(self, on: bool): _lib().slh_mlp_set_fused(self.h, int(on))
    def set_deferred(self, on: bool):
        """data-parallel pipelining across steps: each layer's gradient join + SGD update moves into the next forward pass (flush() applies what is owed)"""
        _chk(_lib().slh_mlp_set_deferred(self.h, int(on)))
    def flush(self): _chk(_lib().slh_mlp_flush(self.h))
    def allreduce_grads(self): _chk(_lib().slh_mlp_allreduce_grads(self.h))
    def sgd(self, lr): _chk(_lib().slh_mlp_sgd(self.h, lr))

    def step(self, x, y, labels, batch, lr, grad_rows=None, want_metrics=True):
        loss, correct = _d(0), C.c_longlong(0)
        _chk(_lib().slh_mlp_step(self.h, x.h, y.h, labels.h if labels else None, batch, grad_rows or batch, lr, int(want_metrics),
                                 C.byref(loss), C.byref(correct)))
        return loss.value, correct.value

    def step_replay(self, x, y, labels, batch, lr, grad_rows=None, want_metrics=True):
        """The step as a CUDA graph: recorded on the first call, one cudaGraphLaunch afterwards (custos `Lazy` + `run()`)."""
        loss, correct = _d(0), C.c_longlong(0)
        _chk(_lib().slh_mlp_step_replay(self.h, x.h, y.h, labels.h if labels else None, batch, grad_rows or batch, lr, int(want_metrics),
                                        C.byref(loss), C.byref(correct)))
        return loss.value, correct.value

    def predict(self, x: Buffer, batch: int) -> Buffer:
        return Buffer(self.device, _lib().slh_mlp_predict(self.h, x.h, batch))
