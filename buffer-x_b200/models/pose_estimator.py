"""Pose estimation on the GPU.

Mirrors ``PoseEstimator`` of /root/reference/models/pose_estimator.py (:14-126): same constructor,
``estimate_pose(src_kpts, tgt_kpts, inlier_ind) -> (4x4 pose, num_inliers)`` and
``compute_confidence_score``.  The RANSAC back-end is ``bx_ransac`` (one-kernel-chain hypothesise /
verify / sequential bookkeeping on the device) instead of Open3D on host cores.  Sampling is an explicit
function of ``(seed, iteration)``; ``cfg.match.ransac_seed`` (default 0) selects the stream.
KISS-Matcher is not part of this path: like the reference when the package is missing
(pose_estimator.py:74-82) the estimator warns once and uses RANSAC.
"""
import numpy as np
import torch

from bufferx_b200 import ops


class PoseEstimator:
    def __init__(self, cfg):
        self.cfg = cfg
        self.pose_estimator = cfg.match.pose_estimator
        self.seed = int(cfg.match.get("ransac_seed", 0))
        self._ws = {}           # one RANSAC workspace per CUDA stream (pairs may be in flight on several streams)
        self._warned = False

    # ---- sync-free device path used by BufferX.forward ------------------------------------------------
    def enqueue(self, ss, tt, inlier_ind, d_I, maxI, seed=None):
        """Enqueue RANSAC on the current stream; returns the 18-double device result block."""
        if self.pose_estimator == "kiss_matcher" and not self._warned:
            print("Warning: KISS-Matcher back-end is not available in bufferx_b200. Falling back to RANSAC.")
            self._warned = True
            self.pose_estimator = "ransac"
        elif self.pose_estimator not in ("ransac", "kiss_matcher"):
            raise ValueError(f"Unknown pose estimator: {self.pose_estimator}")
        m = self.cfg.match
        key = (str(ss.device), torch.cuda.current_stream().cuda_stream)
        if key not in self._ws:
            self._ws[key] = ops.ransac_workspace(m.iter_n, ss.device)
        return ops.ransac(ss, tt, inlier_ind, d_I, maxI, m.dist_th, m.similar_th, m.confidence, m.iter_n,
                          self.seed if seed is None else seed, workspace=self._ws[key])

    # ---- reference-compatible entry point -------------------------------------------------------------
    def estimate_pose(self, src_kpts, tgt_kpts, inlier_ind, seed=None):
        dev = src_kpts.device if isinstance(src_kpts, torch.Tensor) and src_kpts.is_cuda else torch.device("cuda")
        ss = torch.as_tensor(src_kpts, dtype=torch.float32).to(dev).contiguous()
        tt = torch.as_tensor(tgt_kpts, dtype=torch.float32).to(dev).contiguous()
        ind = torch.as_tensor(np.asarray(inlier_ind.detach().cpu() if isinstance(inlier_ind, torch.Tensor) else inlier_ind),
                              dtype=torch.int32).to(dev).contiguous()
        n = int(ind.numel())
        if n == 0:
            ind = torch.zeros(1, dtype=torch.int32, device=dev)
        d_I = torch.tensor([n], dtype=torch.int32, device=dev)
        res = self.enqueue(ss, tt, ind, d_I, max(n, 1), seed)
        T, num_inliers, _, _ = ops.decode_ransac_result(res.cpu())
        return T, num_inliers

    def compute_confidence_score(self, num_inliers):
        return num_inliers >= self.cfg.match.get("early_exit_min_inliers", 15)
