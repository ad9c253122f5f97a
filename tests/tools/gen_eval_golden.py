"""Generate tests/golden/eval_golden.json from the REFERENCE's own evaluation / IO functions (build container only).

    python tests/tools/gen_eval_golden.py

Imports /root/reference/utils/tools.py and utils/result_io.py UNMODIFIED.  Their third-party imports that are absent offline
are stubbed: ``open3d`` (unused by the functions exercised) and ``nibabel.quaternions.mat2quat`` (provided by SciPy's
``Rotation.as_quat(canonical=True)`` re-ordered to (w, x, y, z): an implementation independent of ours).  A hand-made scene
(ground-truth ``gt.log`` / ``gt.info``, an estimated ``.log``) and a hand-made ``states`` table go through
``read_trajectory``, ``read_trajectory_info``, ``evaluate_registration``, ``save_per_sample_results`` and
``save_full_results_csv``; inputs and outputs are stored so that tests/test_evaluation_cpu.py can replay them against
``bufferx_b200.evaluation`` without the reference.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    from scipy.spatial.transform import Rotation

    def mat2quat(M):
        x, y, z, w = Rotation.from_matrix(np.asarray(M, dtype=np.float64)).as_quat(canonical=True)
        return np.array([w, x, y, z])

    _stub("open3d")
    nib = _stub("nibabel")
    nib.quaternions = _stub("nibabel.quaternions", mat2quat=mat2quat)
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    import utils.tools as RT
    import utils.result_io as RIO

    rng = np.random.default_rng(7)
    n_frag = 9
    pairs, gts, infos = [], [], []
    for i in range(n_frag):
        for j in range(i + 1, n_frag):
            if rng.random() < 0.55:
                R = Rotation.from_rotvec(rng.normal(scale=0.6, size=3)).as_matrix()
                T = np.eye(4)
                T[:3, :3], T[:3, 3] = R, rng.normal(scale=1.0, size=3)
                A = rng.normal(size=(6, 6))
                info = A @ A.T * 50 + np.eye(6) * 300
                pairs.append((i, j)); gts.append(T); infos.append(info)
    tmp = tempfile.mkdtemp()
    gt_log, gt_info, est_log = os.path.join(tmp, "gt.log"), os.path.join(tmp, "gt.info"), os.path.join(tmp, "est.log")
    with open(gt_log, "w") as f:
        for (i, j), T in zip(pairs, gts):
            f.write(f"{i}\t {j}\t {n_frag}\n")
            for r in range(4):
                f.write("\t".join(f"{v:.8e}" for v in T[r]) + "\n")
    with open(gt_info, "w") as f:
        for (i, j), I in zip(pairs, infos):
            f.write(f"{i}\t{j}\t{n_frag}\n")
            for r in range(6):
                f.write("\t".join(f"{v:.8e}" for v in I[r]) + "\n")
    # estimates: good, slightly off, wrong, and pairs that are not in the ground truth; written like test.py:156-166 (inverse pose)
    est = []
    for k, ((i, j), T) in enumerate(zip(pairs, gts)):
        if k % 5 == 4:
            continue
        noise = 0.005 if k % 3 else (0.5 if k % 2 else 0.08)
        D = np.eye(4)
        D[:3, :3] = Rotation.from_rotvec(rng.normal(scale=noise, size=3)).as_matrix()
        D[:3, 3] = rng.normal(scale=noise, size=3)
        est.append((i, j, T @ D))
    est.append((0, 1, np.eye(4)))
    est_poses_for_log = []
    with open(est_log, "w") as f:
        for i, j, T in est:
            trans_est = np.linalg.inv(T)              # so that the logged inverse is T
            est_poses_for_log.append(trans_est)
            trans = np.linalg.inv(trans_est)
            f.write(f"{i}\t {j}\t  1\n")
            for r in range(4):
                f.write(f"{trans[r, 0]}\t {trans[r, 1]}\t {trans[r, 2]}\t {trans[r, 3]}\t \n")
    gt_pairs, gt_traj = RT.read_trajectory(gt_log)
    n_fragments, gt_cov = RT.read_trajectory_info(gt_info)
    est_pairs, est_traj = RT.read_trajectory(est_log)
    prec, rec, flags, errs = RT.evaluate_registration(n_fragments, est_traj, est_pairs, gt_pairs, gt_traj, gt_cov)
    single = [RT.computeTransformationErr(np.linalg.inv(gt_traj[k]) @ est_traj[min(k, len(est_traj) - 1)], gt_cov[k]) for k in range(4)]

    # states table + CSV writers
    states = []
    for k in range(13):
        ok = k % 4 != 3
        states.append([ok, abs(rng.normal(0.03, 0.02)) + (0 if ok else 0.5), abs(rng.normal(1.0, 0.5)) + (0 if ok else 20), int(rng.integers(3, 60)),
                       int(rng.integers(800, 1300)), int(rng.integers(10, 90)), 3, rng.random() * 0.1, rng.random() * 0.02 + 0.005,
                       rng.random() * 0.01, rng.random() * 0.005, rng.random() * 0.002])
    ps = os.path.join(tmp, "per", "per_sample.csv")
    RIO.save_per_sample_results(np.array(states), ps, "RANSAC", "OFF")
    summary_row = {"dataset": "synthetic", "recall": 0.75, "rte_mean_cm": 3.5, "rte_std_cm": 1.25, "rre_mean_deg": 1.1, "rre_std_deg": 0.4,
                   "inliers_mean": 30.0, "inliers_std": 5.0, "mutual_inliers_mean": 1100.0, "mutual_inliers_std": 80.0, "inlier_ind_mean": 45.0,
                   "inlier_ind_std": 9.0, "scales_used_mean": 3.0, "scales_used_std": 0.0, "avg_data_time_s": 0.05, "std_data_time_s": 0.01,
                   "avg_model_time_s": 0.009, "std_model_time_s": 0.001}
    cwd = os.getcwd()
    os.chdir(tmp)
    full = RIO.save_full_results_csv([summary_row], "exp/threedmatch", "0101_0000", 512, 3, 1500)
    full_text = open(full).read()
    os.chdir(cwd)
    out = dict(gt_log=open(gt_log).read(), gt_info=open(gt_info).read(), est_log=open(est_log).read(),
               est_entries=[[i, j, np.asarray(P).tolist()] for (i, j, _), P in zip(est, est_poses_for_log)],
               n_fragments=int(n_fragments), gt_pairs=gt_pairs.tolist(), est_pairs=est_pairs.tolist(),
               gt_traj=gt_traj.astype(np.float64).tolist(), est_traj=est_traj.astype(np.float64).tolist(), gt_cov=gt_cov.astype(np.float64).tolist(),
               precision=prec, recall=rec, flags=[int(v) for v in flags], errors=[None if np.isnan(e) else float(e) for e in errs],
               single_errors=[float(v) for v in single], states=[[float(v) for v in s] for s in states], per_sample_csv=open(ps).read(),
               summary_row=summary_row, full_csv=full_text, full_csv_name=os.path.basename(full))
    with open(os.path.join(ROOT, "tests", "golden", "eval_golden.json"), "w") as f:
        json.dump(out, f)
    print("precision", prec, "recall", rec, "flags", flags)


if __name__ == "__main__":
    main()
