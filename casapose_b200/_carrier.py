"""Tensor carriers.  The product is the C-ABI library; PyTorch only carries device memory.

`as_cuda_f32` accepts a torch tensor or ANY object exporting ``__dlpack__`` (TensorFlow eager
tensors, CuPy, JAX, ...) and returns a zero-copy torch view of it.  Layout or dtype mismatches are
errors, not silent copies."""
import torch


def as_cuda(x, dtype, name):
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__"):
            x = torch.from_dlpack(x)
        else:
            raise TypeError("%s: expected a torch tensor or a __dlpack__ exporter, got %r" % (name, type(x)))
    if not x.is_cuda:
        raise ValueError("%s must live on a CUDA device (got %s); use the *_host entry point for host buffers" % (name, x.device))
    if x.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, x.dtype))
    if not x.is_contiguous():
        raise ValueError("%s must be C-contiguous (NHWC, like the reference's tensors)" % name)
    return x


def as_cuda_f32(x, name):
    return as_cuda(x, torch.float32, name)


def ptr(t):
    return None if t is None else t.data_ptr()


def current_stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
