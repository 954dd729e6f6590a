#!/usr/bin/env python
"""Benchmark of the radial keypoint-voting path (BASELINE.json: "voting frames/s (640x480, 3 kpts)").

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU

A *step* is one pass of the whole path over one batch of synthetic LINEMOD-shaped frames resident in
HBM: mask + back-projection + compaction (K1), grid prelude, sphere-shell voting with the fused peak
search (K2/K3), millimetre un-shift, Horn pose (K4) and -- with more than one rank -- the small
all_gather of the per-frame results.  Workload = BASELINE.json configs[2]: 4096 frames x 3 keypoints
per GPU (frames are independent, ranks share nothing: weak scaling).

Prints ONE JSON line (rank 0).  `value` is frames/s with inputs already in HBM; `e2e` is the same
metric through the C-ABI host entry point (rcv_vote_frames_host + rcv_horn_batch_host) with pinned
HOST buffers, host<->device copies inside the timed region.  `roofline` is for the dominant kernel
(k_vote) against the binding roofline of the north star: the shared-memory atomic rate, measured live
on this GPU; `roofline_hbm` gives the HBM view of the same step.  `cpu_baseline` times the oracle's
restatement of the reference's brute-force vote loop on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W, KPTS = 480, 640, 3
METRIC = "voting frames/s (640x480, 3 kpts)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-frames", type=int, default=6, help="frames in the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per step of --impl reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shard", default="cost", choices=["cost", "count"],
                    help="ycb workload: contiguous frame ranges balanced by the sum N*R^2 cost model (SURVEY 8e) or by frame count")
    ap.add_argument("--workload", default="linemod", choices=["linemod", "ycb"],
                    help="linemod = BASELINE configs[2] (the headline line); ycb = configs[3]-shaped frames (YCB camera, objects of 60-110 mm)")
    return ap.parse_args()


def config(frames, n_gpus, workload="linemod"):
    what = ("BASELINE configs[2]: batched voting, %d synthetic LINEMOD-shaped frames x 3 keypoints per GPU "
            "(640x480 uint16 depth + 3 float32 radius maps, object radius 40-70 mm at 0.7-1.1 m, sigma 0.01 dm, 2%% outliers)" % frames)
    if workload == "ycb":
        what = ("BASELINE configs[3]-shaped: batched voting, %d synthetic YCB-Video-shaped frames x 3 keypoints per GPU (YCB camera, "
                "640x480 uint16 depth + 3 float32 radius maps, object radius 60-110 mm at 0.7-1.1 m, sigma 0.01 dm, 2%% outliers; one global "
                "sequence ordered by decreasing distance, cut into contiguous per-rank ranges)" % frames)
    return {"workload": what,
            "frames_per_gpu": frames, "global_frames": frames * n_gpus, "keypoints": KPTS, "image": [H, W],
            "parallelism": "frames sharded over %d GPU(s), no data-path collective, one all_gather of results" % n_gpus,
            "cache": "inputs (%.1f GB per GPU) exceed L2 (126 MB); no reuse between steps" % (frames * H * W * (2 + 4 * KPTS) / 1e9)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle restatement of fast_for, all host threads)
# ------------------------------------------------------------------------------------------------
def cpu_items(frames):
    from rcvpose_b200 import synth
    items = []
    for fr in frames:
        for k in range(KPTS):
            items.append(synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][k]))
    return items


def host_threads():
    """Threads for the CPU arm: every core this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 to
    its workers, which would silently make the reference arm single-threaded, so the count is passed explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


CPU_PORT = ("oracle port of the reference's fast_for (float64 shell test per voxel, OpenMP over x-slices) with exact per-slice / per-row "
            "culls -- ~0.6x the time of the real numba fast_for on the same 8 cores (tools/cpu_calibrate.py, profiles/r02_cpu_calibration.json), "
            "so the ratio to it is conservative -- + lmshorn")


def cpu_run(items, threads=0):
    """Accumulator_3D (the reference's shell-test loop, OpenMP over slices) + lmshorn per frame; returns seconds, votes."""
    from oracle import oracle
    t0 = time.perf_counter()
    votes = 0
    centres = []
    for xyz, rl in items:
        c, info = oracle.Accumulator_3D(xyz, rl, method="brute", return_info=True, threads=threads)
        votes += info["votes"]
        centres.append(c[0])
    A = np.zeros((4, 4))
    for f in range(len(items) // KPTS):
        est = np.array(centres[KPTS * f: KPTS * f + KPTS])
        oracle.lmshorn(est + 1.0, est, KPTS, A)
    return time.perf_counter() - t0, votes, np.array(centres)


def run_reference(args):
    """The CPU arm.  Under torchrun only rank 0 works (the other ranks exit at once), with ALL host cores: the value is the
    whole box's CPU throughput whatever --gpus says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rcvpose_b200 import synth
    threads = host_threads()
    nf = args.ref_frames
    pools = [cpu_items([synth.config3_frame(10_000 + s * nf + f) for f in range(nf)]) for s in range(min(4, args.steps + args.warmup))]
    for w in range(args.warmup):
        cpu_run(pools[w % len(pools)][:KPTS], threads)
    tot, votes = 0.0, 0
    for s in range(args.steps):
        dt, v, _ = cpu_run(pools[s % len(pools)], threads)
        tot += dt
        votes += v
    value = args.steps * nf / tot
    sample = "%d frames x %d keypoints per step (%d frames timed in all, on rank 0 only), %d OpenMP threads: %s" % (
        nf, KPTS, nf * args.steps, threads, CPU_PORT)
    cfg = config(args.frames, args.gpus)
    cfg["reference_arm"] = "CPU only: %d frames per step on the host cores of the box; global_frames describes the GPU arm's step" % nf
    cfg["frames_timed_per_step"] = nf
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "gvotes_per_s": votes / tot / 1e9,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].strip().lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows else None,
                "power_w_max": max([float(r[2]) for r in rows if len(r) >= 3] or [0.0]), "samples": len(sm), "reasons": reasons}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rcvpose_b200 import pipeline, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d bench.py --gpus %d ..." % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.frames
    ycb = args.workload == "ycb"
    Knp = synth.ycb_K if ycb else synth.linemod_K
    counts, shard_info = [B] * world, None
    if ycb:
        # One global sequence of world * B frames, ordered like a camera approaching the object (the cost of a frame grows along
        # it), cut into contiguous per-rank ranges: by the sum N*R^2 cost model of SURVEY 8e (default) or by frame count.
        gp = synth.frame_params(B * world, KPTS, seed=1000, obj_radius_mm=(60.0, 110.0), approach=True)
        cost = synth.frame_cost(gp, Knp)
        ranges = pipeline.shard_by_cost(cost, world) if args.shard == "cost" else [pipeline.shard_range(B * world, r, world) for r in range(world)]
        sums = [float(cost[a:b].sum()) for a, b in ranges]
        counts = [b - a for a, b in ranges]
        shard_info = {"policy": args.shard, "frames_per_rank": counts, "cost_imbalance_max_over_mean": max(sums) / (sum(sums) / world),
                      "cost_model": "sum over keypoints of N * R^2 (N = pixels of the projected object, R = keypoint distance in voxels)"}
        lo, hi = ranges[rank]
        B = hi - lo
        data = synth.torch_batch(B, KPTS, seed=1000 + rank, device=dev, K=Knp, params={k: v[lo:hi] for k, v in gp.items()})
    else:
        data = synth.torch_batch(B, KPTS, seed=1000 + rank, device=dev, K=Knp, obj_radius_mm=(40.0, 70.0))
    depth, radius, model = data["depth"], data["radius"], data["model_mm"]
    K = torch.from_numpy(Knp).to(dev)
    pipe = pipeline.VotingPipeline(local, max_frames=B, n_kpts=KPTS, max_points_total=max(1 << 22, B * KPTS * (90000 if ycb else 12000)),
                                   max_grid=384 if ycb else 256)
    ctx = pipe.ctx
    atom_peak = ctx.measure_smem_atomic_peak()

    def step():
        return pipe.step_gathered(depth, radius, K, model, counts=counts)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    bad = int((out["status"] != 0).sum().item())
    votes_per_step = int(out["votes"].sum().item())
    points_per_step = int(out["n_points"].sum().item())
    l0 = ctx.launches
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launches - l0
    kt = ctx.vote_kernel_times(min(args.steps, 64))
    vote_ms = torch.tensor([sum(kt) / max(1, len(kt))], dtype=torch.float64, device=dev)
    tot = torch.tensor([votes_per_step, points_per_step, bad], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(vote_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    frames_global = sum(counts)
    value = frames_global / (ms_step * 1e-3)
    votes_global, points_global, bad_global = float(tot[0].item()), float(tot[1].item()), int(tot[2].item())

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # the e2e leg pins a host copy of the step's inputs on every rank; on a box whose RAM is shared by all ranks keep the
        # pinned total under half of what is available (whole 256-frame chunks; the rate is per frame, the note says so)
        Be = B
        try:
            import psutil
            per_frame = H * W * (2 + 4 * KPTS)
            fit = int(0.5 * psutil.virtual_memory().available / max(1, world) / per_frame)
            if fit < B:
                Be = max(256, fit // 256 * 256)
        except Exception:
            pass
        if world > 1:
            t_be = torch.tensor([Be], dtype=torch.int64, device=dev)
            dist.all_reduce(t_be, op=dist.ReduceOp.MIN)
            Be = int(t_be.item())
        hd = torch.empty((Be,) + tuple(depth.shape[1:]), dtype=depth.dtype, pin_memory=True)
        hr = torch.empty((Be,) + tuple(radius.shape[1:]), dtype=radius.dtype, pin_memory=True)
        hd.copy_(depth[:Be])
        hr.copy_(radius[:Be])
        hdn, hrn = hd.numpy().view(np.uint16), hr.numpy()
        hmodel = model[:Be].cpu().numpy()
        res = {k: torch.empty(s, dtype=t, pin_memory=True).numpy() for k, s, t in
               [("centre_mm", (Be, KPTS, 3), torch.float64), ("peak", (Be, KPTS), torch.int32), ("votes", (Be, KPTS), torch.int64),
                ("n_points", (Be, KPTS), torch.int32), ("grid", (Be, KPTS), torch.int32), ("status", (Be, KPTS), torch.int32)]}

        def e2e_step():
            o = ctx.vote_frames_host(hdn, hrn, Knp, mask_flags=1, frames_per_chunk=256, out=res)
            return o, ctx.horn_batch_host(hmodel, o["centre_mm"])

        o, RT = e2e_step()                                  # warm-up (allocates staging)
        assert np.array_equal(o["centre_mm"], out["centre_mm"][:Be].cpu().numpy()), "host and device entry points disagree"
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d_full = hdn.nbytes + hrn.nbytes + 72 + hmodel.nbytes + o["centre_mm"].nbytes
        h2d = ctx.last_h2d_bytes + hmodel.nbytes + o["centre_mm"].nbytes      # what crossed the bus: images cropped to their non-zero depth rows
        d2h = sum(v.nbytes for v in res.values()) + RT.nbytes
        e2e = {"value": Be * world * args.e2e_steps / float(dt.item()), "unit": UNIT, "frames_per_gpu_per_step": Be, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": args.e2e_steps, "h2d_bytes_per_step_uncropped": int(h2d_full),
               "h2d_gbs_per_gpu": h2d * args.e2e_steps / float(dt.item()) / 1e9,
               "api": "rcv_vote_frames_host + rcv_horn_batch_host (pinned host buffers, 256-frame chunks, copy/compute overlap; the entry point "
                      "copies only the rows between the first and last non-zero depth row of each frame, two copy streams; up to 3 helper threads scan the depth rows ahead of the copies)"}
        del hd, hr

    # ---- cpu_baseline + live parity spot check (rank 0, single GPU only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nf = min(args.cpu_frames, B)
        frames = [dict(K=Knp, depth=depth[f].cpu().numpy().view(np.uint16), radius=radius[f].cpu().numpy()) for f in range(nf)]
        items = cpu_items(frames)
        threads = host_threads()
        cpu_run(items[:1], threads)                         # warm-up (library load, thread pool)
        dt, v, centres = cpu_run(items, threads)
        got = out["centre_mm"][:nf].cpu().numpy().reshape(-1, 3)
        gv = int(out["votes"][:nf].sum().item())
        parity = {"frames": nf, "centres_bit_equal": bool(np.array_equal(got, centres)), "votes_equal": bool(gv == v)}
        cpu = {"value": nf / dt, "unit": UNIT, "cores": threads, "kind": "port", "gvotes_per_s": v / dt / 1e9,
               "sample": "%d frames x %d keypoints of this workload (%.1f s): %s" % (nf, KPTS, dt, CPU_PORT)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        vote_ms_avg = float(vote_ms.item())
        votes_per_launch = votes_global / world
        achieved = votes_per_launch / (vote_ms_avg * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes of one k_vote launch from the committed ncu capture (headline workload at its default size only)
            if args.workload == "linemod" and B == 4096:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "vote_kernel_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        alg_bytes = B * H * W * (2 + 4 * KPTS) * 2 + points_global / world * 36 * 2   # K1 reads maps twice; pool written + read
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 geometry / int32 votes",
                "data": "synthetic", "config": dict(config(args.frames, world, args.workload), global_frames=frames_global),
                "gvotes_per_s": votes_global / (ms_step * 1e-3) / 1e9, "votes_per_frame": votes_global / frames_global,
                "points_per_frame_kpt": points_global / frames_global / KPTS, "items_with_error_status": bad_global,
                "roofline": {"kernel": "k_vote", "bound": "smem_atomic", "achieved": achieved, "peak": atom_peak / 1e9, "unit": "Gvotes/s",
                             "frac": achieved / (atom_peak / 1e9), "traffic": traffic, "kernel_ms_per_launch": vote_ms_avg,
                             "kernel_share_of_step": vote_ms_avg / ms_step,
                             "peak_source": "conflict-free shared-memory atomicAdd rate measured on this GPU by rcv_ubench_smem_atomics (one vote = one smem atomic)"},
                "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                                 "algorithmic_bytes_per_step": alg_bytes},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "cpu_baseline": cpu, "parity_spot_check": parity}
        if shard_info:
            line["shard"] = shard_info
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints (NCCL's version banner, ...)
    was diverted to stderr in main()."""
    os.write(_JSON_OUT if _JSON_OUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _JSON_OUT
    args = parse()
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)          # keep stdout for the JSON line only
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
