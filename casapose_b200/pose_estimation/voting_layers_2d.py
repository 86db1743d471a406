"""Drop-in for casapose.pose_estimation.voting_layers_2d.CoordLSVotingWeighted.

Same constructor and call convention as the reference Keras layer
(/root/reference/casapose/pose_estimation/voting_layers_2d.py:5-30):

    layer = CoordLSVotingWeighted(name, num_classes, num_points=9, sigmoid_weights=False,
                                  filter_estimates=False, output_second_largest_component=False)
    coords = layer([seg, direct, w])        # [b, num_classes-1, num_points, 2], (y, x) pixels, float32

The body is one call into the sm_100a library (casa_ls_vote); there is no CPU / PyTorch fallback.
Forward only: the training-time gradient w.r.t. `direct` and `w` is out of scope."""
import ctypes as C

import torch

from .. import _lib
from .._carrier import as_cuda_f32, current_stream_ptr, ptr


class CoordLSVotingWeighted:
    def __init__(self, name, num_classes, num_points=9, sigmoid_weights=False, filter_estimates=False,
                 output_second_largest_component=False):
        self.name = name
        self.num_classes = num_classes
        self.num_points = num_points
        self.sigmoid_weights = sigmoid_weights
        self.filter_estimates = filter_estimates
        self.sigmoid_scale = 1.0
        self.output_second_largest_component = output_second_largest_component
        self.height = None
        self.width = None

    def build(self, input_shape):  # voting_layers_2d.py:23-25
        self.width = input_shape[0][2]
        self.height = input_shape[0][1]

    def __call__(self, inp, **kwargs):
        return self.call(inp, **kwargs)

    def call(self, inp, return_debug=False, check_finite=True, **kwargs):
        seg, direct, w = inp
        seg = as_cuda_f32(seg, "seg")
        direct = as_cuda_f32(direct, "direct")
        w = as_cuda_f32(w, "w")
        if seg.dim() != 4:
            raise ValueError("seg must be [b,h,w,num_classes], got %s" % (tuple(seg.shape),))
        b, h, wd, nc = seg.shape
        if self.height is None:
            self.build([seg.shape, direct.shape, w.shape])
        if nc != self.num_classes:
            raise ValueError("seg has %d channels, layer was built for num_classes=%d" % (nc, self.num_classes))
        vn = self.num_points
        if direct.dim() == 5:
            direct = direct.reshape(b, h, wd, 2 * vn)
        if tuple(direct.shape) != (b, h, wd, 2 * vn) or tuple(w.shape) != (b, h, wd, vn):
            raise ValueError("direct must be [b,h,w,%d] and w [b,h,w,%d]" % (2 * vn, vn))
        dev = seg.device
        out = torch.empty((b, nc - 1, vn, 2), dtype=torch.float32, device=dev)
        p = _lib.LsParams(b=b, h=h, w=wd, num_classes=nc, vn=vn, sigmoid_weights=int(bool(self.sigmoid_weights)),
                          filter_estimates=int(bool(self.filter_estimates)),
                          second_largest=int(bool(self.output_second_largest_component)), min_component=0,
                          check_finite=int(bool(check_finite)))
        dbg = None
        dbg_struct = None
        if return_debug:
            dbg = {
                "sums": torch.zeros((b, nc - 1, vn, 5), dtype=torch.float64, device=dev),
                "labels": torch.zeros((b, h, wd), dtype=torch.uint8, device=dev),
                "selected": torch.zeros((b, nc - 1), dtype=torch.int32, device=dev),
                "parent": torch.zeros((b, h, wd), dtype=torch.int32, device=dev),
                "tn": torch.zeros((b, nc - 1), dtype=torch.int32, device=dev),
            }
            dbg_struct = _lib.LsDebug(**{k: ptr(v) for k, v in dbg.items()})
        hdl = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
        with torch.cuda.device(dev):
            rc = _lib.lib().casa_ls_vote(hdl, C.byref(p), ptr(seg), ptr(direct), ptr(w), ptr(out),
                                         C.byref(dbg_struct) if dbg_struct is not None else None, current_stream_ptr(dev))
        _lib.check(rc)
        return (out, dbg) if return_debug else out
