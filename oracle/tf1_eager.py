"""TEST INFRASTRUCTURE ONLY -- a minimal EAGER stand-in for the `tensorflow` 1.3 module, just large enough to execute the
reference's own graph-building code (nets/ColorHandPose3DNetwork.py, nets/PosePriorNetwork.py, utils/general.py,
utils/relative_trafo.py of lmb-freiburg/hand3d) function by function on numpy arrays.

Purpose: TensorFlow 1.3 cannot be installed here, so the reference graph cannot be run.  What CAN be done is to run the
reference's unmodified PYTHON code (layer lists, names, strides, concat order, crop arithmetic, Rodrigues formula, kinematic chain,
tuple orders ...) with every `tf.*` call bound to an eager numpy implementation.  The heavy ops (conv2d, pools, legacy bilinear
resize, soft-max, dilation2d, crop_and_resize) delegate to oracle/tf1_ops.py -- the same restatement of the published TF 1.3
kernels the oracle uses -- so the outputs pin the oracle's GRAPH restatement (oracle/hand3d_oracle.py) to the reference source,
not the op semantics (those stay pinned by the known-answer tests only: "parity unpinned" still applies to them).

Since round 2 it also carries the queue-reader stubs (FixedLengthRecordReader, string_input_producer, decode_raw, batch_join) and
the extra element-wise ops the reference's dataset readers need (data/BinaryDbReader.py, data/BinaryDbReaderSTB.py,
utils/canonical_trafo.py), so that their derived items can be generated from the reference source as well.

Used only by tests/golden/make_golden_reference_graph.py and tests/golden/make_golden_reference_reader.py.  Variables: `tf.get_variable` looks the full scoped name up in
`set_weights({name: ndarray})` and checks the shape the reference asks for.
"""
from __future__ import annotations

import contextlib
import types

import numpy as np

from . import tf1_ops as T

float32, int32, int64, bool = np.float32, np.int32, np.int64, np.bool_     # noqa: A001  (tf.bool)


class _Shape:
    def __init__(self, s):
        self._s = [int(v) for v in s]

    def as_list(self):
        return list(self._s)

    def __len__(self):
        return len(self._s)


class Tensor(np.ndarray):
    """ndarray with the two shape methods the reference calls on tf.Tensor."""
    def get_shape(self):
        return _Shape(self.shape)

    def set_shape(self, s):
        assert list(self.shape) == list(s), (self.shape, s)


def _w(x):
    return np.asarray(x).view(Tensor)


def _f(x):     # python floats / lists become float32 like in TF 1.x; existing arrays keep their dtype
    a = np.asarray(x)
    if a.dtype == np.float64 and not isinstance(x, np.ndarray):
        a = a.astype(np.float32)
    return a


# ------------------------------------------------------------------------------------------------- variables / scopes
_scopes: list = []
_weights: dict = {}
requested: list = []           # (name, shape) in the order the reference asked for them


def set_weights(w):
    _weights.clear(); _weights.update(w); requested.clear(); _scopes.clear()


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _scopes.append(name)
    try:
        yield
    finally:
        _scopes.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, collections=None):
    full = "/".join(_scopes + [name])
    if full not in _weights:
        raise KeyError("reference asked for variable %r which the weight dictionary does not hold" % full)
    v = np.asarray(_weights[full], np.float32)
    assert list(v.shape) == [int(s) for s in shape], (full, v.shape, shape)
    requested.append((full, tuple(v.shape)))
    return _w(v)


def constant_initializer(*a, **k):
    return None


contrib = types.SimpleNamespace(
    layers=types.SimpleNamespace(xavier_initializer_conv2d=lambda *a, **k: None, xavier_initializer=lambda *a, **k: None),
    framework=types.SimpleNamespace())


# ------------------------------------------------------------------------------------------------- element-wise / shape ops
def constant(v, dtype=None, **k):
    return _w(np.asarray(_f(v), dtype) if dtype else _f(v))


def cast(x, dtype, **k):
    return _w(np.asarray(x).astype(dtype))


def reshape(x, shape, **k):
    return _w(np.reshape(np.asarray(x), [int(s) for s in np.asarray(shape).reshape(-1)]))


def concat(values, axis, **k):
    return _w(np.concatenate([np.asarray(v) for v in values], axis))


def stack(values, axis=0, **k):
    return _w(np.stack([np.asarray(_f(v)) for v in values], axis))


def tile(x, multiples, **k):
    return _w(np.tile(np.asarray(x), multiples))


def transpose(x, perm=None, **k):
    return _w(np.transpose(np.asarray(x), perm))


def expand_dims(x, axis, **k):
    return _w(np.expand_dims(np.asarray(x), axis))


def squeeze(x, axis=None, **k):
    return _w(np.squeeze(np.asarray(x), tuple(axis) if axis is not None else None))


def ones(shape, dtype=np.float32, **k):
    return _w(np.ones(shape, dtype))


def zeros(shape, dtype=np.float32, **k):
    return _w(np.zeros(shape, dtype))


def ones_like(x, **k):
    return _w(np.ones_like(np.asarray(x)))


def zeros_like(x, **k):
    return _w(np.zeros_like(np.asarray(x)))


def range(*a, **k):      # noqa: A001
    return _w(np.arange(*a, dtype=np.int32))


def _pair(a, b):
    """TF converts python numbers / tuples to the dtype of the tensor operand (numpy would promote a tuple of ints to int64)."""
    if isinstance(a, np.ndarray) and not isinstance(b, np.ndarray):
        return a, np.asarray(b, a.dtype)
    if isinstance(b, np.ndarray) and not isinstance(a, np.ndarray):
        return np.asarray(a, b.dtype), b
    return a, b


def maximum(a, b, **k):
    return _w(np.maximum(*_pair(a, b)))


def minimum(a, b, **k):
    return _w(np.minimum(*_pair(a, b)))


def exp(x, **k):
    return _w(np.exp(x))


def logical_or(a, b, **k):
    return _w(np.logical_or(a, b))


def logical_not(a, **k):
    return _w(np.logical_not(a))


def slice(x, begin, size, **k):      # noqa: A001
    x = np.asarray(x)
    idx = tuple(np.s_[int(b):int(b) + int(n)] for b, n in zip(begin, size))
    return _w(x[idx])


def one_hot(indices, depth, on_value=1.0, off_value=0.0, dtype=np.float32, **k):
    idx = np.asarray(indices)
    out = np.full(idx.shape + (depth,), off_value, dtype)
    np.put_along_axis(out, idx[..., None].astype(np.int64), np.asarray(on_value, dtype), -1)
    return _w(out)


uint8 = np.uint8


def decode_raw(value, out_type, **k):
    return _w(np.frombuffer(bytes(value), dtype=out_type).copy())


# ---- queue-reader stubs: the records of the file named by string_input_producer are handed out one per read()
_record_files: dict = {}


class _Queue:
    def __init__(self, names):
        self.names = list(names)


class FixedLengthRecordReader:
    def __init__(self, header_bytes=0, record_bytes=0, **k):
        assert header_bytes == 0
        self.record_bytes = int(record_bytes)

    def read(self, queue):
        name = queue.names[0]
        st = _record_files.setdefault(name, {"pos": 0})
        with open(name, "rb") as f:
            f.seek(st["pos"] * self.record_bytes)
            value = f.read(self.record_bytes)
        assert len(value) == self.record_bytes, "ran out of records in %s" % name
        st["pos"] += 1
        return "%s:%d" % (name, st["pos"] - 1), value


def reset_readers():
    _record_files.clear()


def _batch_join(tensors_list, batch_size, capacity=None, enqueue_many=False, **k):
    assert batch_size == 1 and len(tensors_list) == 1 and not enqueue_many
    return [_w(np.expand_dims(np.asarray(t), 0)) for t in tensors_list[0]]


train = types.SimpleNamespace(string_input_producer=lambda names, **k: _Queue(names), batch_join=_batch_join)


def multiply(a, b, **k):
    return _w(np.multiply(a, b))


def square(x, **k):
    return _w(np.square(x))


def sqrt(x, **k):
    return _w(np.sqrt(x))


def sin(x, **k):
    return _w(np.sin(x))


def cos(x, **k):
    return _w(np.cos(x))


def atan(x, **k):
    return _w(np.arctan(x))


def round(x, **k):       # noqa: A001
    return _w(T.round_half_even(np.asarray(x)))


def equal(a, b, **k):
    return _w(np.equal(a, b))


def less(a, b, **k):
    return _w(np.less(a, b))


def greater(a, b, **k):
    return _w(np.greater(a, b))


def greater_equal(a, b, **k):
    return _w(np.greater_equal(a, b))


def logical_and(a, b, **k):
    return _w(np.logical_and(a, b))


def is_finite(x, **k):
    return _w(np.isfinite(x))


def where(c, a, b, **k):
    return _w(np.where(np.asarray(c), a, b))


def reduce_all(x, **k):
    return _w(np.all(x))


def reduce_sum(x, axis=None, **k):
    return _w(np.sum(np.asarray(x), axis))


def reduce_max(x, axis=None, **k):   # TF reduces an empty tensor to the identity of the op (-inf / +inf)
    x = np.asarray(x)
    if x.size == 0:
        return _w(np.float32(-np.inf))
    return _w(np.max(x, axis))


def reduce_min(x, axis=None, **k):
    x = np.asarray(x)
    if x.size == 0:
        return _w(np.float32(np.inf))
    return _w(np.min(x, axis))


def argmax(x, axis=None, dimension=None, **k):
    return _w(np.argmax(np.asarray(x), axis if axis is not None else dimension).astype(np.int64))


def boolean_mask(x, mask, **k):
    return _w(np.asarray(x)[np.asarray(mask)])


def sparse_to_dense(sparse_indices, output_shape, sparse_values, default_value=0, **k):
    out = np.full([int(s) for s in output_shape], default_value, np.float32)
    for idx in np.asarray(sparse_indices).reshape(-1, len(output_shape)):
        out[tuple(int(i) for i in idx)] = sparse_values
    return _w(out)


def dynamic_stitch(indices, data, **k):   # only the use of the reference: indices [[0],[1],...], data[i] of shape [1, B]
    n = len(indices)
    assert [list(i) for i in indices] == [[i] for i in np.arange(n)]
    return _w(np.concatenate([np.asarray(_f(d)) for d in data], 0))


def cond(pred, fn1, fn2, **k):
    return fn1() if np.asarray(pred).item() else fn2()


def check_numerics(x, message, **k):
    assert np.isfinite(np.asarray(x)).all(), message
    return x


def matmul(a, b, **k):
    a, b = np.asarray(a), np.asarray(b)
    if a.ndim == 2 and b.ndim == 2 and a.shape[1] > 64:      # FC layers: the oracle's kernel (same summation as its own path)
        return _w(T.fully_connected(a, b, np.zeros(b.shape[1], np.float32)))
    return _w(np.matmul(a, b))


def matrix_inverse(x, **k):
    return _w(np.linalg.inv(np.asarray(x)).astype(np.asarray(x).dtype))


# ------------------------------------------------------------------------------------------------- tf.nn / tf.image
def _conv2d(x, kernel, strides, padding, **k):
    assert padding == "SAME" and strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    kernel = np.asarray(kernel)
    return _w(T.conv2d_same(np.asarray(x), kernel, np.zeros(kernel.shape[3], np.float32), int(strides[1])))


def _bias_add(x, b, **k):
    return _w(np.asarray(x) + np.asarray(b))


def _max_pool(x, ksize, strides, padding, **k):
    assert list(ksize) == [1, 2, 2, 1] and list(strides) == [1, 2, 2, 1] and padding == "VALID"
    return _w(T.max_pool_2x2(np.asarray(x)))


def _avg_pool(x, ksize, strides, padding, **k):
    assert list(ksize) == [1, 8, 8, 1] and list(strides) == [1, 8, 8, 1] and padding == "SAME"
    return _w(T.avg_pool_8x8(np.asarray(x)))


def _softmax(x, **k):
    return _w(T.softmax_last(np.asarray(x)))


def _dropout(x, keep_prob, noise_shape=None, **k):
    assert keep_prob == 1.0, "only the evaluation branch (keep_prob 1.0 = identity) can be executed eagerly"
    return x


def _dilation2d(x, filt, strides, rates, padding, **k):
    x, filt = np.asarray(x), np.asarray(filt)
    assert padding == "SAME" and list(strides) == [1, 1, 1, 1] and list(rates) == [1, 1, 1, 1]
    assert filt.shape == (21, 21, 1) and np.all(filt == np.float32(1.0) / np.float32(441.0)) and x.shape[0] == 1 and x.shape[3] == 1
    return _w(T.dilation2d_21(x[0, :, :, 0])[None, :, :, None])


nn = types.SimpleNamespace(conv2d=_conv2d, bias_add=_bias_add, max_pool=_max_pool, avg_pool=_avg_pool, softmax=_softmax,
                           dropout=_dropout, dilation2d=_dilation2d)


def _resize_images(x, size, **k):
    x = np.asarray(x)
    if x.ndim == 3:      # tf.image.resize_images also takes a single [H,W,C] image (data/BinaryDbReader.py:372)
        return _w(T.resize_bilinear_tf1(x[None], int(size[0]), int(size[1]))[0])
    return _w(T.resize_bilinear_tf1(x, int(size[0]), int(size[1])))


def _crop_and_resize(image, boxes, box_ind, crop_size, **k):
    image = np.asarray(image)
    assert list(np.asarray(box_ind)) == list(np.arange(image.shape[0]))
    cs = np.asarray(crop_size)
    return _w(T.crop_and_resize(image, np.asarray(boxes), int(cs[0]), int(cs[1])))


image = types.SimpleNamespace(resize_images=_resize_images, crop_and_resize=_crop_and_resize)
