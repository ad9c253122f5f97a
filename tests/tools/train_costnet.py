"""Fixture generator (CPU, build container): fits the CostNet parameters of the synthetic checkpoint so that the
synthetic pairs actually register.

The trained reference checkpoint is not available offline; `synth.init_synthetic_weights` seeds random weights.
Random descriptor weights still give repeatable, matchable, azimuth-equivariant features (the network is a
deterministic function of the LRF-aligned patch), but a random CostNet predicts a random relative yaw, so every
pose hypothesis is wrong, the consensus set has ~10 members and RANSAC / refinement run on a toy problem.  This
script keeps the seeded descriptor weights, generates (source map, target map, true yaw bin) triples from
synthetic C2-shaped pairs with the CPU oracle (the same LRF / SPT / conv arithmetic the GPU path is tested
against) and fits ONLY `Pose.conv.*` to the objective of the reference's training (soft arg-max of the azimuth
logits = yaw between the two local frames, models/BUFFERX.py:66-69, 382-389).  Output:
buffer-x_b200/data/pose_synth_trained.npz (loaded by init_synthetic_weights(trained_pose=True)).

    python tests/tools/train_costnet.py [--pairs 24] [--steps 2500]
"""
import argparse
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bufferx_b200 as bx  # noqa: E402
from bufferx_b200.synth import init_synthetic_weights, make_pair, workload_cfg  # noqa: E402
from oracle import oracle as O  # noqa: E402

_COST_CONVS = [0, 3, 6, 9, 12, 15, 18, 21, 24, 27]


def costnet_logits(d1, d2, sd, azi_n=20, pfx="Pose.conv."):
    """oracle.cost_volume up to the logits (differentiable in sd)."""
    M = d1.shape[0]
    l = torch.arange(azi_n)
    idx = (l[None, :] - l[:, None]) % azi_n
    x = d1[:, :, :, idx.reshape(-1)].reshape(M, d1.shape[1], d1.shape[2], azi_n, azi_n)
    x = x.permute(0, 1, 3, 2, 4) - d2.unsqueeze(2)
    for i in _COST_CONVS:
        x = F.conv3d(x, sd[pfx + f"ops.{i}.weight"], sd[pfx + f"ops.{i}.bias"])
        if i != 27:
            x = F.relu(O._bn(x, sd, pfx + f"ops.{i + 1}", False))
    return x.reshape(M, azi_n)


def gen_pair_samples(sd, cfg, seed, n_kp):
    data = make_pair("C2", seed)
    src, tgt, T = data["src_fds_pcd"], data["tgt_fds_pcd"], data["relt_pose"].astype(np.float64)
    aligned = bool(data["is_aligned_to_global_z"])
    Kr = cfg.patch.num_points_radius_estimate
    si, ti = O.fps(src, Kr), O.fps(tgt, Kr)
    big, bk = (src, src[si]) if len(src) > len(tgt) else (tgt, tgt[ti])
    cum = O.radius_hist(bk, big)
    radii = [O.radius_estimation(src, src[si], tgt, tgt[ti], [th], cum=cum)[0] for th in cfg.patch.search_radius_thresholds]
    # corresponding key-points: source FPS points and the nearest target point of their ground-truth image
    sk = src[si[:n_kp]]
    img = sk.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    d2 = ((img[:, None, :] - tgt[None, :, :].astype(np.float64)) ** 2).sum(-1)
    nn = d2.argmin(1)
    keep = np.sqrt(d2[np.arange(len(sk)), nn]) < 0.03
    sk, tk = sk[keep], tgt[nn[keep]]
    perms = O.draw_perms(cfg, len(src), len(tgt), seed)
    out = []
    for i, r in enumerate(radii):
        a = O.describe(sd, cfg, src, sk, r, aligned, perms[i][0])
        b = O.describe(sd, cfg, tgt, tk, r, aligned, perms[i][1])
        A = b["R"].double().transpose(1, 2) @ torch.from_numpy(T[:3, :3]) @ a["R"].double()     # Rz(angle) = tt_R^T R_gt ss_R
        ang = torch.atan2(A[:, 1, 0], A[:, 0, 0]) % (2 * math.pi)
        ok = A[:, 2, 2] > 0.97                                        # local z axes agree: the pair is a pure yaw
        out.append((a["equi"][ok][:, :, 1:cfg.patch.ele_n - 1].clone(), b["equi"][ok][:, :, 1:cfg.patch.ele_n - 1].clone(),
                    (ang[ok] / (2 * math.pi / cfg.patch.azi_n)).float()))
    return out


def soft_target(bins, azi_n):
    lo = torch.floor(bins)
    w = bins - lo
    t = torch.zeros(len(bins), azi_n)
    t[torch.arange(len(bins)), lo.long() % azi_n] = 1 - w
    t[torch.arange(len(bins)), (lo.long() + 1) % azi_n] += w
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=24)
    ap.add_argument("--kp", type=int, default=400)
    ap.add_argument("--steps", type=int, default=2500)
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--out", default=os.path.join(ROOT, "buffer-x_b200", "data", "pose_synth_trained.npz"))
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    O.build()
    cfg = workload_cfg("C2")
    model = init_synthetic_weights(bx.BufferX(cfg))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    t0 = time.time()
    D1, D2, Y = [], [], []
    for s in range(args.pairs):
        for d1, d2, y in gen_pair_samples(sd, cfg, 100 + s, args.kp):     # seeds disjoint from the bench / test pairs
            D1.append(d1); D2.append(d2); Y.append(y)
        print(f"pair {s}: {sum(len(y) for y in Y)} samples, {time.time() - t0:.0f} s", flush=True)
    D1, D2, Y = torch.cat(D1), torch.cat(D2), torch.cat(Y)
    n = len(Y)
    n_val = max(64, n // 10)
    perm = torch.randperm(n)
    va, tr = perm[:n_val], perm[n_val:]
    names = [k for k in sd if k.startswith("Pose.conv.") and (k.endswith(".weight") or k.endswith(".bias")) and sd[k].dtype.is_floating_point
             and "running" not in k]
    params = []
    for k in names:
        sd[k] = sd[k].clone().requires_grad_(True)
        params.append(sd[k])
    opt = torch.optim.Adam(params, lr=1e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=2e-3, total_steps=args.steps)
    azi = cfg.patch.azi_n

    def circ_err(logits, y):
        ind = (F.softmax(logits, -1) * torch.arange(azi)[None]).sum(-1)
        e = (ind - y).abs()
        return torch.minimum(e, azi - e)

    for step in range(args.steps):
        b = tr[torch.randint(0, len(tr), (args.batch,))]
        logits = costnet_logits(D1[b], D2[b], sd, azi)
        loss = -(soft_target(Y[b], azi) * F.log_softmax(logits, -1)).sum(-1).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if step % 100 == 0 or step == args.steps - 1:
            with torch.no_grad():
                e = torch.cat([circ_err(costnet_logits(D1[va[i:i + 128]], D2[va[i:i + 128]], sd, azi), Y[va[i:i + 128]]) for i in range(0, n_val, 128)])
            print(f"step {step}: loss {loss.item():.3f}  val soft-arg-max error: median {e.median():.2f} bins, <1 bin {100 * (e < 1).float().mean():.0f} %  "
                  f"({time.time() - t0:.0f} s)", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, **{k: sd[k].detach().numpy().astype(np.float32) for k in sd if k.startswith("Pose.conv.") and sd[k].dtype.is_floating_point})
    print("wrote", args.out, os.path.getsize(args.out), "bytes")


if __name__ == "__main__":
    main()
