"""Stand-in for the one tensorflow-addons call on the path: ``tfa.image.connected_components``
(/root/reference/casapose/pose_estimation/voting_layers_2d.py:53; TFA 0.17.0 is neither installed nor
vendored).  TEST INFRASTRUCTURE (see oracle/__init__.py and oracle/tf_standin/tensorflow/__init__.py).

Published contract of the TFA op: zero pixels map to 0, every other pixel to the id (> 0) of its
4-connected component of equal-valued pixels, ids contiguous in row-major order of each component's
first pixel.  scipy.ndimage.label with its default cross-shaped structure has the same contract for the
{0,1} images the layer feeds it; equal-valued-ness is enforced by labelling every distinct value
separately."""
import numpy as _np
from scipy import ndimage as _ndi


class _Image:
    @staticmethod
    def connected_components(images, name=None):
        images = _np.asarray(images)
        assert images.ndim == 2, "stand-in: the reference calls this per 2-D image (map_fn, :56)"
        values = [v for v in _np.unique(images) if v != 0]
        if len(values) <= 1:
            lab, _ = _ndi.label(images != 0)
            return lab.astype(_np.int32)
        # several distinct non-zero values: label each, then renumber in row-major order of first pixel
        out = _np.zeros(images.shape, _np.int64)
        nxt = 0
        for v in values:
            lab, n = _ndi.label(images == v)
            out[lab > 0] = lab[lab > 0] + nxt
            nxt += n
        flat = out.reshape(-1)
        ids, first = _np.unique(flat[flat > 0], return_index=True)
        order = _np.argsort(first)
        remap = _np.zeros(nxt + 1, _np.int64)
        remap[ids[order]] = _np.arange(1, len(ids) + 1)
        return remap[out].astype(_np.int32)


image = _Image()
