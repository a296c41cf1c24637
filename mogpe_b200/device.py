"""Device buffers and zero-copy tensor interchange (DLPack / __cuda_array_interface__) without any
tensor framework: TF, torch, cupy tensors and numpy arrays are all accepted at the API boundary."""
import ctypes as C

import numpy as np

from . import _lib


class DeviceBuffer:
    """Owning float64 device allocation made through the C-ABI."""

    def __init__(self, shape, device=0, zero=True):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape)) if not isinstance(shape, tuple) else shape
        self.device = device
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        p = C.c_void_p()
        _lib.check(_lib.load().mogpe_malloc(device, self.size * 8, C.byref(p)))
        self.ptr = p.value
        if zero:
            _lib.check(_lib.load().mogpe_memset_zero(device, self.ptr, self.size * 8, None))

    @classmethod
    def from_numpy(cls, a, device=0):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = cls(a.shape, device, zero=False)
        b.copy_from(a)
        return b

    def copy_from(self, a, offset=0):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert offset + a.size <= self.size
        _lib.check(_lib.load().mogpe_memcpy_h2d(self.device, self.ptr + offset * 8, a.ctypes.data, a.size * 8, None))
        _lib.check(_lib.load().mogpe_stream_sync(self.device, None))

    def numpy(self, offset=0, count=None, shape=None):
        count = self.size - offset if count is None else count
        out = np.empty(count, dtype=np.float64)
        _lib.check(_lib.load().mogpe_memcpy_d2h(self.device, out.ctypes.data, self.ptr + offset * 8, count * 8, None))
        if shape is not None:
            return out.reshape(shape)
        return out.reshape(self.shape) if (offset == 0 and count == self.size) else out

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f8", "data": (self.ptr, False), "version": 3, "strides": None}

    def free(self):
        if getattr(self, "ptr", None):
            _lib.load().mogpe_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host float64 buffer (cudaMallocHost) viewed as a numpy array."""

    def __init__(self, n):
        p = C.c_void_p()
        _lib.check(_lib.load().mogpe_malloc_host(int(n) * 8, C.byref(p)))
        self.ptr, self.size = p.value, int(n)
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_double)), shape=(self.size,))

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            _lib.load().mogpe_free_host(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- DLPack capsule parsing (dlpack.h, DLManagedTensor) -----------------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


_kDLCPU, _kDLCUDA, _kDLCUDAHost = 1, 2, 3


class TensorView:
    """Borrowed view of a float64 tensor: pointer + shape + where it lives.  Keeps its owner alive."""

    __slots__ = ("ptr", "shape", "on_device", "device_id", "_owner", "__weakref__")

    def __init__(self, ptr, shape, on_device, device_id, owner):
        self.ptr, self.shape, self.on_device, self.device_id, self._owner = ptr, tuple(shape), on_device, device_id, owner


_USED_NAME = b"used_dltensor"  # module-level: PyCapsule_SetName keeps the pointer, not a copy
_DELETER = C.CFUNCTYPE(None, C.c_void_p)


def _from_dlpack_capsule(capsule, owner):
    """Consume a DLPack capsule (the protocol of `__dlpack__`): read the DLManagedTensor, rename the capsule to
    "used_dltensor" so that its own destructor no longer frees the tensor, and call the producer's deleter when the returned
    view is garbage collected.  Zero-copy: the view points at the producer's memory."""
    import weakref

    api = C.pythonapi
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    api.PyCapsule_SetName.restype = C.c_int
    api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
    p = api.PyCapsule_GetPointer(capsule, b"dltensor")
    managed = C.cast(p, C.POINTER(_DLManagedTensor)).contents
    t = managed.dl_tensor
    if not (t.dtype.code == 2 and t.dtype.bits == 64 and t.dtype.lanes == 1):
        raise TypeError("mogpe_b200 expects float64 tensors (the reference runs in float64)")
    shape = [t.shape[i] for i in range(t.ndim)]
    if t.strides:
        expect = 1
        for i in reversed(range(t.ndim)):
            if shape[i] != 1 and t.strides[i] != expect:
                raise ValueError("tensor must be C-contiguous")
            expect *= shape[i]
    on_dev = t.device.device_type == _kDLCUDA
    if not on_dev and t.device.device_type not in (_kDLCPU, _kDLCUDAHost):
        raise TypeError(f"unsupported DLPack device type {t.device.device_type}")
    view = TensorView(t.data + t.byte_offset, shape, on_dev, t.device.device_id, (capsule, owner))
    deleter = managed.deleter
    api.PyCapsule_SetName(capsule, _USED_NAME)  # consumed: from here on WE release the tensor
    if deleter:
        weakref.finalize(view, _DELETER(deleter), p)
    return view


def as_view(x):
    """numpy / DeviceBuffer / anything with __cuda_array_interface__ or __dlpack__ -> TensorView."""
    if isinstance(x, TensorView):
        return x
    if isinstance(x, DeviceBuffer):
        return TensorView(x.ptr, x.shape, True, x.device, x)
    if isinstance(x, np.ndarray) or isinstance(x, (list, tuple, float, int)):
        a = np.ascontiguousarray(x, dtype=np.float64)
        return TensorView(a.ctypes.data, a.shape, False, 0, a)
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        if cai["typestr"] not in ("<f8", "=f8", "f8"):
            raise TypeError("mogpe_b200 expects float64 tensors (the reference runs in float64)")
        if cai.get("strides") is not None:
            expect = 8
            for s, st in zip(reversed(cai["shape"]), reversed(cai["strides"])):
                if s != 1 and st != expect:
                    raise ValueError("tensor must be C-contiguous")
                expect *= s
        return TensorView(cai["data"][0], cai["shape"], True, 0, x)
    if hasattr(x, "__dlpack__"):
        return _from_dlpack_capsule(x.__dlpack__(), x)
    if hasattr(x, "numpy"):
        return as_view(x.numpy())
    raise TypeError(f"cannot interpret {type(x)} as a float64 tensor")
