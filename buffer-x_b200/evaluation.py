"""Evaluation / IO glue fed from the all-gathered per-pair records.

The reference's ``test.py`` is a single-process loop that evaluates every pair inline (test.py:148-192), writes 3DMatch
``.log`` trajectories (:151-166), per-sample and summary CSV files (utils/result_io.py:7-49, 80-124) and, for 3DMatch /
3DLoMatch, the RMSE-based recall of the benchmark (utils/tools.py:66-129, test.py:283-310).  Here the pairs are sharded
over one process per GPU (``driver.py``); every rank packs a 32-float record per pair, one all-gather collects them, and
this module turns the gathered records into exactly those artefacts on rank 0.  Same names, argument order, file formats
and number formatting as the reference functions they mirror (cited per function); NumPy only -- nothing here touches the
GPU, so it is covered by the CPU test-suite against fixtures generated from the reference's own functions
(tests/tools/gen_eval_golden.py).
"""
from __future__ import annotations

import csv
import os

import numpy as np

from .driver import unpack_record
from .se3 import compute_rre, compute_rte

__all__ = ["read_trajectory", "read_trajectory_info", "mat2quat", "computeTransformationErr", "evaluate_registration",
           "append_trajectory_entry", "pair_state", "states_from_records", "summarize_states", "save_per_sample_results",
           "save_full_results_csv", "format_final_results_summary", "print_final_results_summary", "evaluate_3dmatch_scenes"]


# --------------------------------------------------------------------------------------------------------------------
# 3DMatch benchmark files (reference utils/tools.py:66-96)
# --------------------------------------------------------------------------------------------------------------------
def read_trajectory(filename, dim=4):
    """``gt.log`` / estimated ``.log``: blocks of one key line "i \\t j \\t n" + ``dim`` matrix rows.
    -> (keys [P,3] str array, trajectories [P,dim,dim] float32)   (utils/tools.py:66-75)"""
    with open(filename) as f:
        lines = f.readlines()
    keys = lines[0::(dim + 1)]
    final_keys = [[k.split("\t")[0].strip(), k.split("\t")[1].strip(), k.split("\t")[2].strip()] for k in keys]
    traj = [lines[i].split("\t")[0:dim] for i in range(len(lines)) if i % (dim + 1) != 0]
    return np.asarray(final_keys), np.asarray(traj, dtype=np.float32).reshape(-1, dim, dim)


def read_trajectory_info(filename, dim=6):
    """``gt.info``: blocks of one key line + 6 rows of the 6x6 information matrix.
    -> (number of fragments, info [P,6,6] float32)   (utils/tools.py:78-96)"""
    with open(filename) as fid:
        contents = fid.readlines()
    n_pairs = len(contents) // 7
    info_list = []
    for i in range(n_pairs):
        rows = [np.array(contents[j].split(), dtype=np.float64).reshape(1, -1) for j in range(i * 7 + 1, i * 7 + 7)]
        info_list.append(np.vstack(rows))
    return int(contents[0].strip().split()[2]), np.asarray(info_list, dtype=np.float32).reshape(-1, dim, dim)


def mat2quat(M):
    """Rotation matrix -> unit quaternion (w, x, y, z), w >= 0: the eigenvector of Bar-Itzhack's symmetric K matrix for its
    largest eigenvalue -- the algorithm of ``nibabel.quaternions.mat2quat`` the reference calls (utils/tools.py:6, 101)."""
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, dtype=np.float64)[:3, :3].flat
    K = np.array([[Qxx - Qyy - Qzz, 0, 0, 0],
                  [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
                  [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                  [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)                 # uses the lower triangle
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = -q
    return q


def computeTransformationErr(trans, info):
    """e = [t, q_xyz];  e^T I e / I[0,0]   (utils/tools.py:99-103)"""
    t, r = trans[:3, 3], trans[:3, :3]
    q = mat2quat(r)
    er = np.concatenate([t, q[1:]], axis=0)
    return (er.reshape(1, 6) @ info @ er.reshape(6, 1) / info[0, 0]).item()


def evaluate_registration(num_fragment, result, result_pairs, gt_pairs, gt, gt_info, err2=0.2):
    """RMSE-based precision / recall of the 3DMatch benchmark over non-consecutive fragment pairs (utils/tools.py:104-129).
    -> (precision, recall, flags (0 good / 1 bad / 2 not in ground truth), per-pair error with NaN where unevaluated)"""
    err2 = err2 ** 2
    gt_mask = np.zeros((num_fragment, num_fragment), dtype=np.int64)
    flags, transformation_errors = [], np.full(result_pairs.shape[0], np.nan)
    for idx in range(gt_pairs.shape[0]):
        i, j = int(gt_pairs[idx, 0]), int(gt_pairs[idx, 1])
        if j - i > 1:
            gt_mask[i, j] = idx
    good, n_res, n_gt = 0, 0, np.sum(gt_mask > 0)
    for idx in range(result_pairs.shape[0]):
        i, j, pose = int(result_pairs[idx, 0]), int(result_pairs[idx, 1]), result[idx, :, :]
        if gt_mask[i, j] > 0:
            n_res += 1
            gt_idx = gt_mask[i, j]
            p = computeTransformationErr(np.linalg.inv(gt[gt_idx]) @ pose, gt_info[gt_idx])
            transformation_errors[idx] = p
            if p <= err2:
                good += 1
                flags.append(0)
            else:
                flags.append(1)
        else:
            flags.append(2)
    return good / max(n_res, 1e-6), good / n_gt, flags, transformation_errors


def append_trajectory_entry(path, src_id, tgt_id, trans_est):
    """One estimated pair in the benchmark's ``.log`` format: the INVERSE of the estimate, tab separated with the
    reference's exact spacing (test.py:156-166)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    trans = np.linalg.inv(trans_est)
    with open(path, "a+") as f:
        f.write(f"{src_id}\t {tgt_id}\t  1\n")
        for r in range(4):
            f.write(f"{trans[r, 0]}\t {trans[r, 1]}\t {trans[r, 2]}\t {trans[r, 3]}\t \n")


def evaluate_3dmatch_scenes(gt_root, est_root, timestr):
    """RMSE recall per scene and its mean (test.py:283-310): ``gt_root/<scene>/gt.log|gt.info`` against
    ``est_root/<scene>/<timestr>.log``.  -> (mean recall, {scene: (precision, recall)})"""
    scenes = sorted(os.listdir(gt_root))
    out = {}
    for scene in scenes:
        gt_pairs, gt_traj = read_trajectory(os.path.join(gt_root, scene, "gt.log"))
        n_fragments, gt_traj_cov = read_trajectory_info(os.path.join(gt_root, scene, "gt.info"))
        est_pairs, est_traj = read_trajectory(os.path.join(est_root, scene, f"{timestr}.log"))
        precision, recall, _, _ = evaluate_registration(n_fragments, est_traj, est_pairs, gt_pairs, gt_traj, gt_traj_cov)
        out[scene] = (precision, recall)
    return float(np.mean([v[1] for v in out.values()])) if out else float("nan"), out


# --------------------------------------------------------------------------------------------------------------------
# per-pair states (test.py:168-192) from forward() outputs or from gathered records
# --------------------------------------------------------------------------------------------------------------------
def pair_state(trans_est, relt_pose, num_inliers, num_mutual_inliers, num_inlier_ind, scales_used, times, rte_thresh, rre_thresh,
               data_time=0.0, model_time=None):
    """One row of the reference's ``states`` list: [success, rte, rre, num_inliers, num_mutual_inliers, num_inlier_ind,
    scales_used, data_time, model_time, desc_time, pose_time, pose_optim_time]  (test.py:168-192)."""
    trans_est = trans_est if trans_est is not None else np.eye(4)
    rte = compute_rte(trans_est, relt_pose)
    rre = compute_rre(trans_est, relt_pose)
    success = rte < rte_thresh and rre < rre_thresh
    times = list(times)[:3]
    model_time = float(sum(times)) if model_time is None else float(model_time)
    return [success, rte, rre, num_inliers, num_mutual_inliers, num_inlier_ind, scales_used, float(data_time), model_time, *times]


def states_from_records(records, data_times=None, model_times=None):
    """Gathered 32-float records (driver.pack_record, ordered by pair id) -> the ``states`` array of test.py:255."""
    rows = []
    for i, r in enumerate(np.asarray(records).reshape(-1, 32)):
        d = unpack_record(r)
        t = d["times"]
        rows.append([float(d["success"]), d["rte"], d["rre"], d["num_inliers"], d["num_mutual"], d["num_inlier_ind"], d["scales_used"],
                     0.0 if data_times is None else float(data_times[i]), float(sum(t)) if model_times is None else float(model_times[i]), *t])
    return np.array(rows, dtype=np.float64).reshape(-1, 12)


def summarize_states(states, dataset="synthetic"):
    """The aggregate the reference logs and returns (test.py:255-268, 312-338): recall, RTE/RRE over the SUCCESSFUL pairs,
    counter statistics, average times excluding the first five pairs."""
    states = np.asarray(states, dtype=np.float64)
    ok = states[:, 0] == 1
    first = 5
    eff = states[first:, 7:12] if len(states) > first else states[:, 7:12]
    return {
        "dataset": dataset,
        "recall": float(states[:, 0].sum() / states.shape[0]),
        "rte_mean_cm": float(states[ok, 1].mean() * 100) if ok.any() else float("nan"),
        "rte_std_cm": float(states[ok, 1].std() * 100) if ok.any() else float("nan"),
        "rre_mean_deg": float(states[ok, 2].mean()) if ok.any() else float("nan"),
        "rre_std_deg": float(states[ok, 2].std()) if ok.any() else float("nan"),
        "inliers_mean": float(states[:, 3].mean()), "inliers_std": float(states[:, 3].std()),
        "mutual_inliers_mean": float(states[:, 4].mean()), "mutual_inliers_std": float(states[:, 4].std()),
        "inlier_ind_mean": float(states[:, 5].mean()), "inlier_ind_std": float(states[:, 5].std()),
        "scales_used_mean": float(states[:, 6].mean()), "scales_used_std": float(states[:, 6].std()),
        "avg_data_time_s": float(eff[:, 0].mean()), "std_data_time_s": float(eff[:, 0].std()),
        "avg_model_time_s": float(eff[:, 1].mean()), "std_model_time_s": float(eff[:, 1].std()),
    }


# --------------------------------------------------------------------------------------------------------------------
# CSV / table writers (reference utils/result_io.py)
# --------------------------------------------------------------------------------------------------------------------
PER_SAMPLE_HEADER = ["sample_id", "success", "rte_m", "rre_deg", "num_inliers", "num_mutual_inliers", "num_inlier_ind", "scales_used",
                     "data_time_s", "model_time_s", "desc_time_s", "pose_time_s", "poseest_time_s", "pose_estimator", "early_exit"]
FULL_HEADER = ["dataset", "recall", "rte_mean_cm", "rte_std_cm", "rre_mean_deg", "rre_std_deg", "inliers_mean", "inliers_std",
               "mutual_inliers_mean", "mutual_inliers_std", "inlier_ind_mean", "inlier_ind_std", "scales_used_mean", "scales_used_std",
               "avg_data_time_s", "std_data_time_s", "avg_model_time_s", "std_model_time_s", "experiment_id", "timestamp"]


def save_per_sample_results(states, per_sample_file, pose_method, early_exit_status):
    """utils/result_io.py:7-49: one row per pair, times and errors with six decimals."""
    os.makedirs(os.path.dirname(per_sample_file) or ".", exist_ok=True)
    with open(per_sample_file, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(PER_SAMPLE_HEADER)
        for idx, s in enumerate(states):
            w.writerow([idx, int(s[0]), f"{s[1]:.6f}", f"{s[2]:.6f}", int(s[3]), int(s[4]), int(s[5]), int(s[6]), f"{s[7]:.6f}", f"{s[8]:.6f}",
                        f"{s[9]:.6f}", f"{s[10]:.6f}", f"{s[11]:.6f}", pose_method, early_exit_status])


def save_full_results_csv(results, experiment_id, timestr, num_points_per_patch, num_scales, num_fps, full_results_dir="full_results"):
    """utils/result_io.py:80-124: one row per dataset + experiment id / timestamp columns; returns the file path."""
    os.makedirs(full_results_dir, exist_ok=True)
    exp_name = experiment_id.rsplit("/", 1)[-1]
    path = f"{full_results_dir}/results_{exp_name}_{num_points_per_patch}_{num_scales}_{num_fps}_{timestr}.csv"
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=FULL_HEADER)
        w.writeheader()
        for row in results:
            r = dict(row)
            r["experiment_id"] = experiment_id
            r["timestamp"] = timestr
            w.writerow(r)
    return path


def format_final_results_summary(results):
    """The rows of the reference's final table (utils/result_io.py:52-77), four decimals each."""
    headers = ["Scene", "Recall", "RTE mean (cm)", "RTE std (cm)", "RRE mean (deg)", "RRE std (deg)", "Avg data t (s)", "Avg model t (s)"]
    rows = [[r["dataset"], f"{r['recall']:.4f}", f"{r['rte_mean_cm']:.4f}", f"{r['rte_std_cm']:.4f}", f"{r['rre_mean_deg']:.4f}",
             f"{r['rre_std_deg']:.4f}", f"{r['avg_data_time_s']:.4f}", f"{r['avg_model_time_s']:.4f}"] for r in results]
    return headers, rows


def print_final_results_summary(results):
    headers, rows = format_final_results_summary(results)
    try:
        from tabulate import tabulate
        table = tabulate(rows, headers=headers, tablefmt="grid")
    except ImportError:                       # same content, plain layout
        table = "\n".join(" | ".join(map(str, r)) for r in [headers] + rows)
    print("\n\033[1;32m========== Final Results Summary ==========")
    print(table, "\033[0m")
