"""Framework-agnostic entry point: the vote on ANY ``__dlpack__`` exporter, without importing PyTorch.

``ransac_voting_layer_all_masks_dlpack`` has the reference's argument list
(/root/reference/casapose/pose_estimation/ransac_voting.py:446-463) plus a caller-allocated ``out`` tensor and the
CUDA stream the caller works on.  Each tensor's ``__dlpack__()`` capsule is opened with the CPython capsule API,
the ``DLManagedTensor*`` inside goes straight to the C ABI (casa_ransac_vote_dlpack), which validates device, dtype,
shape and strides and consumes the capsule (its deleter runs once, after the GPU work has finished).  This module
imports ctypes and the library loader only — a TensorFlow process never needs torch for the voting path.
"""
import ctypes as C

from . import _lib

_DLTENSOR = b"dltensor"
_USED = b"used_dltensor"  # module-level: PyCapsule_SetName keeps the pointer, not a copy

_api = C.pythonapi
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_SetName.restype = C.c_int
_api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]


def _open(x, name, stream):
    """``x.__dlpack__()`` -> (capsule, DLManagedTensor* as int).  The capsule is renamed "used_dltensor" right away:
    from here on the C library owns the tensor and calls its deleter exactly once."""
    if not hasattr(x, "__dlpack__"):
        raise TypeError("%s: expected a __dlpack__ exporter, got %r" % (name, type(x)))
    try:
        cap = x.__dlpack__(stream=stream)
    except TypeError:
        cap = x.__dlpack__()
    if not _api.PyCapsule_IsValid(cap, _DLTENSOR):
        raise TypeError("%s.__dlpack__() did not return a 'dltensor' capsule" % name)
    ptr = _api.PyCapsule_GetPointer(cap, _DLTENSOR)
    _api.PyCapsule_SetName(cap, _USED)
    return cap, ptr


def ransac_voting_layer_all_masks_dlpack(mask, vertex, out, round_hyp_num, inlier_thresh=0.99, confidence=0.99,
                                         max_iter=20, min_num=5, max_num=30000, *, seed=0, image_offset=0, device=0,
                                         stream=0, seg_scores=False):
    """mask [b,h,w,oc] (or seg scores [b,h,w,1+oc]), vertex [b,h,w,vn,2] | [b,h,w,oc,vn,2], out [b,oc,vn,2]: float32
    CUDA tensors of any framework.  Writes the (x, y) keypoints into `out` (asynchronously on `stream`) and returns it.
    `stream` is the raw cudaStream_t (0 = the legacy default stream); DLPack's own convention for the producer side
    (1 = legacy default stream) is applied when the capsules are requested."""
    lib = _lib.lib()
    hdl = _lib.handle(device, stream)
    p = _lib.RansacParams(
        b=0, h=0, w=0, oc=0, vn=0, round_hyp_num=int(round_hyp_num), max_iter=int(max_iter),
        inlier_thresh=float(inlier_thresh), confidence=float(confidence), min_num=float(min_num), max_num=float(max_num),
        seed=int(seed) & 0xFFFFFFFFFFFFFFFF, image_offset=int(image_offset), pix_capacity=0, force_exact=0, vertex_per_class=0)
    dl_stream = 1 if not stream else int(stream)
    caps = [_open(mask, "mask", dl_stream), _open(vertex, "vertex", dl_stream), _open(out, "out", dl_stream)]
    rc = lib.casa_ransac_vote_dlpack(hdl, C.byref(p), caps[0][1], caps[1][1], caps[2][1], 1 if seg_scores else 0,
                                     C.c_void_p(int(stream) if stream else 0))
    _lib.check(rc)
    return out
