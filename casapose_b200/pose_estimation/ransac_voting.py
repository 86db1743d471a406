"""Drop-in for the voting entry point of casapose.pose_estimation.ransac_voting.

``ransac_voting_layer_all_masks`` keeps the reference signature
(/root/reference/casapose/pose_estimation/ransac_voting.py:446-463) and output
([b, oc, vn, 2] float32, (x, y) pixels); the body is one call into the sm_100a library.
Additive keyword-only extras (seed, idxs, selection, return_debug, ...) exist for testing."""
import ctypes as C

import torch

from .. import _lib
from .._carrier import as_cuda, as_cuda_f32, current_stream_ptr, ptr

DEBUG_FIELDS = ("tn0", "tn", "rounds", "counts", "win_idx", "hyps", "win_pts", "win_ratio", "ata", "atb",
                "refined", "pix", "pix_off", "stats")


def _params(b, h, w, oc, vn, round_hyp_num, inlier_thresh, confidence, max_iter, min_num, max_num, seed,
            image_offset, pix_capacity, force_exact):
    return _lib.RansacParams(
        b=b, h=h, w=w, oc=oc, vn=vn, round_hyp_num=int(round_hyp_num), max_iter=int(max_iter),
        inlier_thresh=float(inlier_thresh), confidence=float(confidence), min_num=float(min_num),
        max_num=float(max_num), seed=int(seed) & 0xFFFFFFFFFFFFFFFF, image_offset=int(image_offset),
        pix_capacity=int(pix_capacity), force_exact=int(bool(force_exact)), reserved=0)


def ransac_voting_layer_all_masks(
    mask,
    vertex,
    round_hyp_num,
    inlier_thresh=0.99,
    confidence=0.99,
    max_iter=20,
    min_num=5,
    max_num=30000,
    *,
    seed=0,
    image_offset=0,
    idxs=None,
    selection=None,
    return_debug=False,
    debug_hyps=False,
    pix_capacity=0,
    force_exact=False,
):
    """
    :param mask:      [b,h,w,oc]   float32 {0,1}
    :param vertex:    [b,h,w,vn,2] float32 (dy,dx)   (a [b,h,w,vn*2] tensor is viewed as such)
    :param round_hyp_num: hypotheses per round
    :return: [b,oc,vn,2] float32 (x,y) — and a dict of intermediates if return_debug
    """
    mask = as_cuda_f32(mask, "mask")
    vertex = as_cuda_f32(vertex, "vertex")
    if mask.dim() != 4:
        raise ValueError("mask must be [b,h,w,oc], got %s" % (tuple(mask.shape),))
    b, h, w, oc = mask.shape
    if vertex.dim() == 4:
        vertex = vertex.view(b, h, w, vertex.shape[3] // 2, 2)
    if vertex.dim() != 5 or tuple(vertex.shape[:3]) != (b, h, w) or vertex.shape[4] != 2:
        raise ValueError("vertex must be [b,h,w,vn,2] matching mask, got %s" % (tuple(vertex.shape),))
    if vertex.device != mask.device:
        raise ValueError("mask and vertex must be on the same device")
    vn = vertex.shape[3]
    dev = mask.device
    p = _params(b, h, w, oc, vn, round_hyp_num, inlier_thresh, confidence, max_iter, min_num, max_num, seed,
                image_offset, pix_capacity, force_exact)
    hn, mi = p.round_hyp_num, p.max_iter
    if idxs is not None:
        idxs = as_cuda(idxs, torch.int32, "idxs")
        if tuple(idxs.shape) != (b, oc, mi, hn, vn, 2):
            raise ValueError("idxs must be [b,oc,max_iter,hn,vn,2] = %s, got %s" % ((b, oc, mi, hn, vn, 2), tuple(idxs.shape)))
    if selection is not None:
        selection = as_cuda_f32(selection, "selection")
        if tuple(selection.shape) != (b, oc, h, w):
            raise ValueError("selection must be [b,oc,h,w]")
    out = torch.empty((b, oc, vn, 2), dtype=torch.float32, device=dev)
    dbg_struct = None
    dbg = None
    if return_debug:
        i32, f32 = torch.int32, torch.float32
        cap = pix_capacity if pix_capacity > 0 else h * w
        dbg = {
            "tn0": torch.zeros((b, oc), dtype=i32, device=dev),
            "tn": torch.zeros((b, oc), dtype=i32, device=dev),
            "rounds": torch.zeros((b, oc), dtype=i32, device=dev),
            "counts": torch.zeros((b, oc, mi, hn, vn), dtype=i32, device=dev),
            "win_idx": torch.zeros((b, oc, mi, vn), dtype=i32, device=dev),
            "win_pts": torch.zeros((b, oc, vn, 2), dtype=f32, device=dev),
            "win_ratio": torch.zeros((b, oc, vn), dtype=f32, device=dev),
            "ata": torch.zeros((b, oc, vn, 3), dtype=f32, device=dev),
            "atb": torch.zeros((b, oc, vn, 2), dtype=f32, device=dev),
            "refined": torch.zeros((b, oc), dtype=i32, device=dev),
            "pix": torch.zeros((b, cap), dtype=i32, device=dev),
            "pix_off": torch.zeros((b, oc), dtype=i32, device=dev),
            "stats": torch.zeros((4,), dtype=torch.int64, device=dev),
        }
        if debug_hyps:
            dbg["hyps"] = torch.zeros((b, oc, mi, hn, vn, 2), dtype=f32, device=dev)
        dbg_struct = _lib.RansacDebug(**{k: ptr(dbg.get(k)) for k in DEBUG_FIELDS})
    hdl = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
    with torch.cuda.device(dev):
        rc = _lib.lib().casa_ransac_vote(
            hdl, C.byref(p), ptr(mask), ptr(vertex), ptr(idxs), ptr(selection), ptr(out),
            C.byref(dbg_struct) if dbg_struct is not None else None, current_stream_ptr(dev))
    _lib.check(rc)
    if return_debug:
        st = C.c_uint32()
        _lib.check(_lib.lib().casa_last_status(hdl, C.byref(st)))
        dbg["status"] = st.value
        return out, dbg
    return out


def ransac_voting_layer_all_masks_host(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.99, max_iter=20,
                                       min_num=5, max_num=30000, *, seed=0, image_offset=0, device=0, out=None):
    """Same call with HOST tensors (numpy arrays or CPU torch tensors, ideally pinned): host->device copy,
    voting, device->host copy of the [b,oc,vn,2] result — all inside the library (casa_ransac_vote_host)."""
    mask_t = torch.as_tensor(mask)
    vertex_t = torch.as_tensor(vertex)
    if mask_t.is_cuda or vertex_t.is_cuda:
        raise ValueError("host entry point takes CPU buffers")
    if mask_t.dtype != torch.float32 or vertex_t.dtype != torch.float32:
        raise TypeError("mask / vertex must be float32")
    if not (mask_t.is_contiguous() and vertex_t.is_contiguous()):
        raise ValueError("mask / vertex must be C-contiguous")
    b, h, w, oc = mask_t.shape
    vn = vertex_t.shape[3] if vertex_t.dim() == 5 else vertex_t.shape[3] // 2
    p = _params(b, h, w, oc, vn, round_hyp_num, inlier_thresh, confidence, max_iter, min_num, max_num, seed,
                image_offset, 0, False)
    if out is None:
        out = torch.empty((b, oc, vn, 2), dtype=torch.float32)
    hdl = _lib.handle(device)
    rc = _lib.lib().casa_ransac_vote_host(hdl, C.byref(p), mask_t.data_ptr(), vertex_t.data_ptr(), out.data_ptr())
    _lib.check(rc)
    return out
