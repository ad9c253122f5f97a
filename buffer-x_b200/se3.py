"""SE(3) helpers used by the hot path and by the success criterion.

Restates ``/root/reference/utils/SE3.py``: ``transform`` (:58-73),
``integrate_trans`` (:91-114), ``compute_rte`` (:134-147) and ``compute_rre``
(:150-165).  RTE is the Euclidean distance of the translations, RRE the
geodesic angle (degrees) of ``R_est^T R_gt`` with the reference's clipping.
"""
import math

import numpy as np
import torch


def transform(pts, trans):
    """R @ p + t for [n,3] or [b,n,3] points (torch or numpy)."""
    if pts.ndim == 3:
        out = trans[:, :3, :3] @ (pts.permute(0, 2, 1) if isinstance(pts, torch.Tensor) else pts.transpose(0, 2, 1))
        out = out + trans[:, :3, 3:4]
        return out.permute(0, 2, 1) if isinstance(out, torch.Tensor) else out.transpose(0, 2, 1)
    return (trans[:3, :3] @ pts.T + trans[:3, 3:4]).T


def integrate_trans(R, t):
    """Pack R [.,3,3] and t [.,3,1] into homogeneous 4x4 (batched or not)."""
    is_t = isinstance(R, torch.Tensor)
    if R.ndim == 3:
        T = torch.eye(4, device=R.device, dtype=R.dtype)[None].repeat(R.shape[0], 1, 1) if is_t \
            else np.tile(np.eye(4)[None], (R.shape[0], 1, 1))
        T[:, :3, :3] = R
        T[:, :3, 3:4] = t.reshape(-1, 3, 1)
    else:
        T = torch.eye(4, device=R.device, dtype=R.dtype) if is_t else np.eye(4)
        T[:3, :3] = R
        T[:3, 3:4] = t.reshape(3, 1)
    return T


def compute_rte(trans_est, trans_gt):
    return float(np.linalg.norm(np.asarray(trans_est)[:3, 3] - np.asarray(trans_gt)[:3, 3]))


def compute_rre(trans_est, trans_gt):
    R_est = np.asarray(trans_est)[:3, :3]
    R_gt = np.asarray(trans_gt)[:3, :3]
    c = (np.trace(R_est.T @ R_gt) - 1) / 2
    c = np.clip(c, -1 + 1e-16, 1 - 1e-16)
    return float(np.arccos(c) * 180 / math.pi)


def is_success(trans_est, trans_gt, rte_thresh, rre_thresh):
    """The reference's acceptance rule (``/root/reference/test.py:168-172``)."""
    return compute_rte(trans_est, trans_gt) < rte_thresh and compute_rre(trans_est, trans_gt) < rre_thresh
