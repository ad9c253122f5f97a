#!/usr/bin/env python
"""bench.py -- registration pairs/sec of the BUFFER-X hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one synthetic pair (BASELINE config C2: 2x20000 points, 1500 FPS key-points, 512 points
per patch, 3 scales, 50000 RANSAC iterations, seeded synthetic weights) through the whole hot path
(FPS -> radius estimation -> 6x [patch gathering, LRF, SPT, conv stack, pooling] -> 3x [matching,
cost volume, hypotheses] -> consensus -> RANSAC -> refinement).
  value : pairs/s with the clouds + permutations already resident in HBM (device-event timed,
          max over ranks, summed over ranks: every rank runs its own K pairs = weak scaling)
  e2e   : the same metric through the public API ``BufferX.forward(data_source)`` with HOST (pinned)
          tensors: H2D of both clouds and the six permutations and D2H of the result block inside
          the timed region.
  roofline     : the dominant kernel (conv_gemm_kernel of the descriptor conv stack), algorithmic
                 FLOPs / CUDA-event time of its launches inside the timed region.
  kernels      : the two HBM-side kernels north_star names (neighbour gather, RANSAC), same method.
  cpu_baseline : the CPU oracle port timed on the host cores on a bounded sample of the same workload.
``--impl reference`` times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "registration pairs/sec (20k-pt clouds, 1500 kpts, 50k RANSAC)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf=1400.0, src="fallback")


def conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_tc_kernel launch, averaged over the eight layers of one
    batched descriptor pass, from the committed `ncu --set full` capture (profiles/r01_conv_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed regions: NVML in-process (10 ms period), nvidia-smi as the
    fallback (its first answer can take longer than a short timed region)."""
    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                h = pynvml.nvmlDeviceGetHandleByIndex(int(ids[gpu]) if gpu < len(ids) else gpu)
            self.nvml = (pynvml, h)
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml:
            nv, h = self.nvml
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            return [str(sm), str(mx)] + ["Active" if mask & bit else "Not Active" for _, bit in self.REASONS]
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [c.strip() for c in out.split(",")] if out else None

    def run(self):
        while not self.stop_flag:
            try:
                r = self.sample()
                if r:
                    self.rows.append(r)
            except Exception:
                pass
            time.sleep(0.01 if self.nvml else 0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples: NVML and nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
def cpu_sample(workload, cfg, sd, data, perms, frac=0.08):
    """Bounded CPU sample of one pair: full FPS / radius / matching / consensus / RANSAC / refinement, and the
    per-key-point stages (patch gathering, LRF, SPT, conv stack, cost volume) on a `frac` subset of the
    key-points, scaled back linearly.  Returns (seconds per pair estimate, cores, description, stage dict)."""
    from oracle import oracle as O
    O.build()
    # all the host threads the port can USE: beyond ~16 threads the small per-patch convolutions of the
    # torch-CPU stacks and the OpenMP loops over a few hundred key-points only lose time to oversubscription
    nthr = max(1, min(os.cpu_count() or 1, 16))
    torch.set_num_threads(nthr)
    O.lib().bxo_set_num_threads(nthr)
    src, tgt = data["src_fds_pcd"], data["tgt_fds_pcd"]
    Kr, K, S = cfg.patch.num_points_radius_estimate, cfg.patch.num_fps, cfg.patch.num_scales
    st = {}
    t0 = time.perf_counter()
    si, ti = O.fps(src, max(Kr, K)), O.fps(tgt, max(Kr, K))
    st["fps"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    big, bk = (src, src[si[:Kr]]) if len(src) > len(tgt) else (tgt, tgt[ti[:Kr]])
    cum = O.radius_hist(bk, big)
    radii = [O.radius_estimation(src, src[si[:Kr]], tgt, tgt[ti[:Kr]], [th], cum=cum)[0] for th in cfg.patch.search_radius_thresholds]
    st["radius_estimation"] = time.perf_counter() - t0
    ks = max(16, int(K * frac))
    sub = {}
    scale_k = K / ks
    desc_s = desc_t = None
    for i in range(S):
        a = O.describe(sd, cfg, src, src[si[:ks]], radii[i], bool(data["is_aligned_to_global_z"]), perms[i][0], timings=sub)
        b = O.describe(sd, cfg, tgt, tgt[ti[:ks]], radii[i], bool(data["is_aligned_to_global_z"]), perms[i][1], timings=sub)
        t0 = time.perf_counter()
        sm, tm, _, _ = O.mutual_nn(a["desc"].numpy(), b["desc"].numpy())
        sub["mutual_nn_sub"] = sub.get("mutual_nn_sub", 0.0) + time.perf_counter() - t0
        t0 = time.perf_counter()
        smi, tmi = torch.from_numpy(sm.astype(np.int64)), torch.from_numpy(tm.astype(np.int64))
        with torch.no_grad():
            O.cost_volume(a["equi"][smi][:, :, 1:cfg.patch.ele_n - 1], b["equi"][tmi][:, :, 1:cfg.patch.ele_n - 1], sd, cfg.patch.azi_n)
        sub["cost_volume"] = sub.get("cost_volume", 0.0) + time.perf_counter() - t0
    for k in ("ball_query_group", "lrf", "spt", "conv_desc", "cost_volume"):
        st[k] = sub.get(k, 0.0) * scale_k
    # full-size matching / consensus / RANSAC / refinement on synthetic descriptors / correspondences of the right size
    rng = np.random.default_rng(0)
    da = rng.normal(size=(K, 32)).astype(np.float32)
    db = rng.normal(size=(K, 32)).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(S):
        O.mutual_nn(da, db)
    st["mutual_nn"] = time.perf_counter() - t0
    Mc = int(0.35 * K) * S
    ss = rng.uniform(-3, 3, (Mc, 3)).astype(np.float32)
    tt = (ss + rng.normal(scale=0.02, size=(Mc, 3))).astype(np.float32)
    tt[Mc // 3:] = rng.uniform(-3, 3, (Mc - Mc // 3, 3))
    R = np.tile(np.eye(3, dtype=np.float32), (Mc, 1, 1))
    tv = rng.normal(scale=0.5, size=(Mc, 3)).astype(np.float32)
    tv[::5] = 0
    t0 = time.perf_counter()
    ind, _, _ = O.consensus(ss, tt, R, tv, cfg.patch.azi_n, cfg.match.inlier_th)
    st["consensus"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = O.ransac(ss, tt, ind, cfg.match.dist_th, cfg.match.similar_th, cfg.match.confidence, cfg.match.iter_n, 0)
    st["ransac"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.refine(ss, tt, r["T"].astype(np.float32), cfg.match.dist_th)
    st["refine"] = time.perf_counter() - t0
    total = sum(st.values())
    desc = (f"{workload}: FPS, radius estimation, matching, consensus, RANSAC, refinement at full size; patch gathering/LRF/SPT/"
            f"conv stack/cost volume on {ks} of {K} key-points per cloud and scale, scaled x{scale_k:.1f}")
    return total, O.num_threads(), desc, st


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path (the reference's own GPU path needs pointnet2_ops, knn_cuda,
    torch_batch_svd and open3d, none of which exist offline) on the host cores; rank 0 only."""
    if rank != 0:
        return
    import bufferx_b200 as bx
    from bufferx_b200.synth import init_synthetic_weights, make_pair, workload_cfg
    from oracle import oracle as O
    cfg = workload_cfg(args.workload)
    model = init_synthetic_weights(bx.BufferX(cfg), trained_pose=True)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    times = []
    cores, desc = 1, ""
    for s in range(args.warmup + args.steps):
        data = make_pair(args.workload, s % 4)
        perms = O.draw_perms(cfg, len(data["src_fds_pcd"]), len(data["tgt_fds_pcd"]), s)
        t, cores, desc, _ = cpu_sample(args.workload, cfg, sd, data, perms, frac=args.cpu_frac or 0.25)
        if s >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = 1.0 / sec
    line = {"metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: 2x{len(data['src_fds_pcd'])} pts, {cfg.patch.num_fps} kpts, {cfg.patch.num_points_per_patch} pts/patch, "
                                   f"{cfg.patch.num_scales} scales, {cfg.match.iter_n} RANSAC iters, seeded synthetic weights",
                       "note": "CPU oracle port of the reference path on the host cores (oracle/)"},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--cpu-frac", type=float, default=None,
                    help="fraction of the key-points the CPU sample runs the per-key-point stages on (default: 1.0 for the\n"
                         "cpu_baseline leg = whole pairs, 0.25 per step for --impl reference)")
    ap.add_argument("--depth", type=int, default=6,
                    help="pairs in flight per GPU (CUDA-graph slots on separate streams).  Measured on 1xB200 at the final commit "
                         "(20 steps): 4 -> 111.7, 6 -> 112.8, 8 -> 114.3 pairs/s; earlier 2 -> 92.6, 3 -> 97.3, 4 -> 100.3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--short", action="store_true", help="profiling runs under ncu: allow < 3 warm-up steps, skip the e2e leg")
    args = ap.parse_args()
    if args.impl == "ours" and not args.short:
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import bufferx_b200 as bx
    from bufferx_b200 import ops
    from bufferx_b200.driver import gather_records, pack_record
    from bufferx_b200.se3 import compute_rre, compute_rte
    from bufferx_b200.synth import init_synthetic_weights, make_pair, workload_cfg

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.load_library()
    cfg = workload_cfg(args.workload)
    model = init_synthetic_weights(bx.BufferX(cfg), trained_pose=True).to(dev)
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    S = cfg.patch.num_scales

    # ---- a pool of distinct synthetic pairs; rank r starts at pair r (round-robin sharding of a virtual list)
    pool = 4
    host, devd = [], []
    for j in range(pool):
        d = make_pair(args.workload, rank * pool + j)
        ns, nt = len(d["src_fds_pcd"]), len(d["tgt_fds_pcd"])
        st = np.random.RandomState(1000 + rank * pool + j)
        perms = [(st.choice(ns, ns, replace=False).astype(np.int32), st.choice(nt, nt, replace=False).astype(np.int32)) for _ in range(S)]
        h = dict(d)
        h["src_fds_pcd"] = torch.from_numpy(d["src_fds_pcd"]).pin_memory()
        h["tgt_fds_pcd"] = torch.from_numpy(d["tgt_fds_pcd"]).pin_memory()
        hp = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in perms]
        host.append((h, hp, d, perms))
        g = dict(d)
        g["src_fds_pcd"] = h["src_fds_pcd"].to(dev)
        g["tgt_fds_pcd"] = h["tgt_fds_pcd"].to(dev)
        devd.append((g, [(a.to(dev), b.to(dev)) for a, b in hp]))
    ns, nt = len(host[0][2]["src_fds_pcd"]), len(host[0][2]["tgt_fds_pcd"])
    h2d_bytes = (ns + nt) * 12 + S * (ns + nt) * 4
    d2h_bytes = (18 + S + 2 + 16) * 8
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    DEPTH = max(1, args.depth)   # pairs in flight per GPU (separate streams; one captured CUDA graph per slot)
    model.enable_cuda_graphs(True, slots_per_shape=DEPTH)

    def run_pipelined(mode, steps, timed):
        """mode 'dev': inputs resident in HBM; 'e2e': pinned host tensors through the public forward_async()."""
        recs, handles = [], []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for slots in model._slots.values():
            for sl in slots:
                sl.stream.wait_event(a)

        def collect(h, s):
            out = h.result()
            if timed:
                gt = host[s % pool][2]["relt_pose"]
                rte, rre = compute_rte(out[0], gt), compute_rre(out[0], gt)
                recs.append(pack_record(rank + world * s, out[0], out[1], out[2], out[3], out[4], out[5], rte, rre,
                                        float(rre < 15.0 and rte < 0.3)))   # 3DMatch success criterion of the reference

        for s in range(steps):
            j = s % pool
            if len(handles) == DEPTH:
                collect(*handles.pop(0))
            with torch.no_grad():
                if mode == "dev":
                    h = model.forward_async(devd[j][0], perms=devd[j][1])
                else:
                    h = model.forward_async(host[j][0], perms=host[j][1])
            handles.append((h, s))
        while handles:
            collect(*handles.pop(0))
        cur = torch.cuda.current_stream()
        for slots in model._slots.values():
            for sl in slots:
                cur.wait_stream(sl.stream)
        b.record()
        b.synchronize()
        return a.elapsed_time(b), recs

    def run_eager(steps):
        """Per-kernel event brackets (ops.Profiler) need eager launches: the roofline pass."""
        model.enable_cuda_graphs(False)
        ms = 0.0
        for s in range(steps):
            j = s % pool
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            with torch.no_grad():
                model(devd[j][0], perms=devd[j][1], ransac_seed=s)
            b.record()
            b.synchronize()
            ms += a.elapsed_time(b)
        model.enable_cuda_graphs(True, slots_per_shape=DEPTH)
        return ms

    l0 = ops.launch_count()
    run_eager(1)                                            # also sets every kernel attribute before graph capture
    launches_per_pair = ops.launch_count() - l0
    # ---- roofline pass: eager launches with per-kernel CUDA-event brackets (before the graph pools exist) ----
    run_eager(2)
    ops.profiler = ops.Profiler()
    n_eager = min(args.steps, 5)
    ms_eager = run_eager(n_eager)
    prof = ops.profiler.summary()
    ops.profiler = None
    run_pipelined("dev", max(args.warmup, DEPTH), False)    # captures the graphs, warms up
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM ----------------------------------------------------
    barrier()
    ms_dev, recs = run_pipelined("dev", args.steps, True)
    allrec = gather_records(np.stack(recs), world * args.steps, device=dev)   # the one collective of the path
    barrier()
    launches = launches_per_pair * args.steps
    # ---- timed region 2: host buffers through the public API --------------------------------------
    if args.short:
        ms_e2e = float("nan")
    else:
        run_pipelined("e2e", DEPTH, False)
        barrier()
        ms_e2e, _ = run_pipelined("e2e", args.steps, True)
        barrier()
    sampler.stop_flag = True

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        pk = measured_peaks()
        value = world * args.steps / (ms_dev / 1e3)
        e2e = world * args.steps / (ms_e2e / 1e3)
        cd = prof.get("conv_desc", dict(launches=0, ms=0.0, work=0.0))
        ach_tf = cd["work"] / (cd["ms"] / 1e3) / 1e12 if cd["ms"] > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (Cylindrical_Net layers; tcgen05 kind::tf32, 3xTF32 split, fp32-equivalent FLOPs)",
                "achieved": ach_tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf"],
                "peak_source": f"{pk['src']} bf16 dense (sustained); the 3-pass TF32 ceiling is ~375 TFLOP/s fp32-equivalent",
                "launches": cd["launches"], "avg_launch_ms": cd["ms"] / max(cd["launches"], 1),
                "share_of_step": cd["ms"] / ms_eager if ms_eager else None, "traffic": conv_traffic(),
                "measured_in": "eager (non-graph) pass of this run: per-kernel CUDA-event brackets need individual launches"}
        kern = {}
        sp = prof.get("select_patches")
        if sp and sp["ms"] > 0:
            gbs = sp["work"] / (sp["ms"] / 1e3) / 1e9
            kern["select_patches"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                                      "launches": sp["launches"], "avg_launch_ms": sp["ms"] / sp["launches"]}
        for k in ("conv_cost", "spt", "lrf", "fps", "ransac"):
            if k in prof:
                kern[k] = {"launches": prof[k]["launches"], "avg_ms": prof[k]["ms"] / max(prof[k]["launches"], 1),
                           "share_of_step": prof[k]["ms"] / ms_eager}
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: 2x{ns} pts, {cfg.patch.num_fps} kpts, {cfg.patch.num_points_per_patch} pts/patch, "
                                       f"{S} scales, {cfg.match.iter_n} RANSAC iters, seeded synthetic weights",
                           "pairs_per_rank": args.steps, "sharding": "pair i -> rank i mod world, one all_gather of 32-float records",
                           "l2": "per-pair working set (~1 GB of activations) exceeds the 126 MB L2; eager pass flushes 256 MB between steps",
                           "pairs_in_flight": DEPTH, "cuda_graphs": True, "eager_ms_per_step": ms_eager / max(n_eager, 1),
                           "weights": "seeded synthetic descriptor weights; CostNet fitted on disjoint synthetic pairs "
                                      "(tests/tools/train_costnet.py) so that the pairs register",
                           "registration_success": float(np.mean(allrec[:, 25])),
                           "median_rre_deg": float(np.median(allrec[:, 24])), "median_rte_m": float(np.median(allrec[:, 23])),
                           "mean_mutual_matches": float(np.mean(allrec[:, 20])),
                           "mean_consensus_inliers": float(np.mean(allrec[:, 21]))},
                "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "kernels": kern}
        if not args.no_cpu_baseline and world == 1:       # the CPU baseline is reported at N = 1 only
            from oracle import oracle as O
            t0 = time.perf_counter()
            secs, stages = [], {}
            for h in host[:2]:                              # two whole pairs: ~10 s of CPU work on the box
                sec, cores, desc, st = cpu_sample(args.workload, cfg, sd_cpu, h[2], h[3], frac=args.cpu_frac or 1.0)
                secs.append(sec)
                for k, v in st.items():
                    stages[k] = stages.get(k, 0.0) + v / 2
            sec = float(np.mean(secs))
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": f"{len(secs)} pairs; " + desc,
                                    "stage_seconds_per_pair": {k: round(v, 4) for k, v in stages.items()},
                                    "wall_s": round(time.perf_counter() - t0, 2)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
