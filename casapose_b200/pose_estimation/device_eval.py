"""Device-resident evaluation loop: voting -> batched PnP -> ADD / ADD-S, one small read-back per batch.

`estimate_and_evaluate_poses` (pose_evaluation.py) keeps the reference's argument list and therefore re-uploads
its constant tables (3-D keypoints, cameras, evaluation meshes) on every call and hands numpy arrays between its
stages like the reference does between TF and OpenCV.  `DeviceEvaluator` is the same computation
(/root/reference/casapose/pose_estimation/pose_evaluation.py:11-101) with the tables uploaded once and every
intermediate left on the GPU: casa_ransac_vote_seg -> casa_pnp -> casa_pose_errors; the per-class sums are one
[oc, 8] read-back.  Same statistics as the drop-in with pnp_backend="cuda", metric_backend="cuda"."""
import numpy as np
import torch

from .ransac_voting import pnp_cuda, pose_errors_cuda, ransac_voting_layer_all_masks


class DeviceEvaluator:
    def __init__(self, object_points_3d, camera_matrix, diameters, evaluation_points=None, object_points_3d_count=None,
                 min_num=20, allowed_error_2d=5.0, round_hyp_num=512, device=None):
        """object_points_3d [oc,vn,3] keypoints; camera_matrix [3,3]; diameters [oc];
        evaluation_points [oc,P,3] + object_points_3d_count [oc] (mesh vertices for ADD / ADD-S) or None (:67-74)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        f = lambda x: torch.as_tensor(np.ascontiguousarray(np.asarray(x, np.float32))).to(dev)  # noqa: E731
        self.dev = dev
        self.kp3 = f(object_points_3d)
        self.oc, self.vn = self.kp3.shape[0], self.kp3.shape[1]
        self.cam = f(camera_matrix).reshape(3, 3)
        self.diam = f(diameters).reshape(self.oc)
        if evaluation_points is not None and object_points_3d_count is not None:
            self.models = f(evaluation_points)
            self.counts = torch.as_tensor(np.asarray(object_points_3d_count, np.int32).reshape(self.oc)).to(dev)
        else:
            self.models = self.kp3
            self.counts = torch.full((self.oc,), self.vn, dtype=torch.int32, device=dev)
        self.min_num, self.allowed_error_2d, self.hn = min_num, allowed_error_2d, round_hyp_num

    def __call__(self, output_seg, target_seg, output_vertex, poses_gt, offsets=None, **vote_kw):
        """output_seg / target_seg [b,h,w,1+oc], output_vertex [b,h,w,2*vn] (CUDA tensors), poses_gt [b,oc,3,4]
        -> (stats dict of [oc] numpy arrays, poses [b,oc,3,4] CUDA tensor, keypoints [b,oc,vn,2] CUDA tensor)."""
        b, h, w, _ = output_seg.shape
        oc, vn = self.oc, self.vn
        valid = ((target_seg[..., 1:] != 0).sum(dim=(1, 2)) > self.min_num).to(torch.int32)  # :30-33
        pts = ransac_voting_layer_all_masks(output_seg, output_vertex.reshape(b, h, w, vn, 2), self.hn, inlier_thresh=0.99,
                                            max_iter=20, min_num=self.min_num, max_num=30000, seg_scores=True, **vote_kw)
        n = b * oc
        off = None
        if offsets is not None:
            off = torch.as_tensor(offsets, dtype=torch.float32, device=self.dev).reshape(b, 1, 10).expand(b, oc, 10).reshape(n, 10).contiguous()
        poses = pnp_cuda(pts.reshape(n, vn, 2), self.kp3.unsqueeze(0).expand(b, oc, vn, 3).reshape(n, vn, 3).contiguous(),
                         self.cam.expand(n, 3, 3).contiguous(), off)
        gt = torch.as_tensor(poses_gt, dtype=torch.float32, device=self.dev).reshape(n, 3, 4)
        rows = pose_errors_cuda(poses, gt, self.cam.expand(n, 3, 3).contiguous(), self.models, self.counts,
                                self.diam.unsqueeze(0).expand(b, oc).reshape(n).contiguous(), valid.reshape(n),
                                self.allowed_error_2d).reshape(b, oc, 6)
        fp_mask = ((valid == 0) & (pts.reshape(b, oc, -1).sum(-1) > 0)).to(torch.float32)  # map_false_positive :517-522
        packed = torch.cat([rows.sum(dim=0), valid.sum(dim=0, keepdim=True).t().float(), fp_mask.sum(dim=0, keepdim=True).t()], dim=1)
        s = packed.cpu().numpy()  # the only read-back: [oc, 8]
        stats = {"err_2d": s[:, 0], "err_3d": s[:, 1], "valid_3d": s[:, 2], "valid_2d": s[:, 3], "missing_object": s[:, 4],
                 "false_positive_pose": s[:, 5], "valid_pose_count": s[:, 6], "false_positive_mask": s[:, 7]}
        return stats, poses.reshape(b, oc, 3, 4), pts
