"""A minimal hand-rolled ``__dlpack__`` exporter that is NOT a torch tensor: device memory from libcudart through
ctypes, the DLManagedTensor built field by field.  Used by tests/test_gpu_dlpack.py (run in a subprocess that never
imports torch) to exercise the framework-agnostic entry point the way a TensorFlow / CuPy tensor would."""
import ctypes as C

import numpy as np


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    pass


DELETER = C.CFUNCTYPE(None, C.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p), ("deleter", DELETER)]

_rt = None


def cudart():
    global _rt
    if _rt is None:
        for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _rt = C.CDLL(name)
                break
            except OSError:
                continue
        if _rt is None:
            raise OSError("libcudart not found")
        _rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        _rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _rt.cudaFree.argtypes = [C.c_void_p]
    return _rt


class CudaBuffer:
    """float32 device array with the array-API DLPack protocol (``__dlpack__`` / ``__dlpack_device__``)."""

    def __init__(self, array=None, shape=None, device=0, strides=None, dtype_code=2):
        rt = cudart()
        assert rt.cudaSetDevice(device) == 0
        self.shape = tuple(array.shape if array is not None else shape)
        self.device = device
        self.nbytes = int(np.prod(self.shape)) * 4
        self.ptr = C.c_void_p()
        assert rt.cudaMalloc(C.byref(self.ptr), max(self.nbytes, 4)) == 0
        if array is not None:
            a = np.ascontiguousarray(array, np.float32)
            assert rt.cudaMemcpy(self.ptr, a.ctypes.data_as(C.c_void_p), self.nbytes, 1) == 0
        self.strides = strides
        self.dtype_code = dtype_code
        self.deleter_calls = 0
        self._live = {}

    def numpy(self):
        out = np.empty(self.shape, np.float32)
        assert cudart().cudaMemcpy(out.ctypes.data_as(C.c_void_p), self.ptr, self.nbytes, 2) == 0
        return out

    def __dlpack_device__(self):
        return (2, self.device)

    def __dlpack__(self, stream=None):
        m = DLManagedTensor()
        nd = len(self.shape)
        shape = (C.c_int64 * nd)(*self.shape)
        strides = (C.c_int64 * nd)(*self.strides) if self.strides is not None else None
        m.dl_tensor.data = self.ptr
        m.dl_tensor.device = DLDevice(2, self.device)
        m.dl_tensor.ndim = nd
        m.dl_tensor.dtype = DLDataType(self.dtype_code, 32, 1)
        m.dl_tensor.shape = shape
        m.dl_tensor.strides = strides if strides is not None else C.cast(None, C.POINTER(C.c_int64))
        m.dl_tensor.byte_offset = 0

        def _del(p):
            self.deleter_calls += 1
            self._live.pop(C.addressof(p.contents), None)

        cb = DELETER(_del)
        m.deleter = cb
        self._live[C.addressof(m)] = (m, shape, strides, cb)  # kept alive until the consumer calls the deleter
        new = C.pythonapi.PyCapsule_New
        new.restype = C.py_object
        new.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        return new(C.addressof(m), b"dltensor", None)
