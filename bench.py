#!/usr/bin/env python
"""bench.py — RGBD frames/s integrated (BASELINE.json metric) on the TUM1-shape workload
(config 2: 640x480, 5 mm voxels, tau 0.04 m, 300 synthetic frames per step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path (hash allocation + projective TSDF/colour update) over the
whole synthetic sequence.  Warm-up passes populate the map; timed passes are steady state.

  value     frames/s with the frames already resident in HBM, timed with CUDA events on the
            launching stream, max over ranks.
  e2e       the same metric through the public API (`B200TsdfVolume.integrate`, i.e. the C ABI)
            with PINNED HOST frames: H2D copies inside the timed region, one D2H read of the
            step's result (blocks touched / allocated) per step; wall clock around a full sync.
  roofline  integrate_kernel: algorithmic bytes (2*S*512*A_f + 7*W*H per frame, S = 20 B,
            SURVEY.md §8d) / CUDA-event duration of every launch, vs MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle port (OpenMP, all host threads) on a bounded sample, rank 0.

N > 1 (torchrun): the voxel-block hash space is sharded by BlockKeyHash % N; every rank sees every
frame and integrates only the blocks it owns (no data-path collective).  Total work is fixed, so
scaling is "strong".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyslam_b200 import synthetic as S  # noqa: E402

METRIC = "rgbd_frames_per_sec_integrated"
UNIT = "frames/s"
VOXEL_RECORD_BYTES = 20  # tsdf f32 + weight f32 + rgb 3 x f32 (SURVEY.md §8d "S")


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------

def _render_one(args):
    name, i = args
    d, c, T = S.render_frame(S.CONFIGS[name], i)
    return i, d, c, T


def load_frames(cfg_name: str, n_frames: int, rank: int, world: int, barrier=None):
    """Render (rank 0, all cores) or load the cached frames of the sequence."""
    cfg = S.CONFIGS[cfg_name]
    step = max(cfg.n_frames // n_frames, 1)
    idx = [k * step for k in range(n_frames)]
    cache = f"/tmp/b2v_frames_{cfg_name}_{n_frames}_{step}.npz"
    if rank == 0 and not os.path.exists(cache):
        import multiprocessing as mp
        procs = max(1, min(len(os.sched_getaffinity(0)), 48))
        with mp.get_context("fork").Pool(procs) as pool:
            res = sorted(pool.map(_render_one, [(cfg_name, i) for i in idx], chunksize=2))
        tmp = cache + f".tmp{os.getpid()}.npz"
        np.savez(tmp, depth=np.stack([r[1] for r in res]), color=np.stack([r[2] for r in res]),
                 Tcw=np.stack([r[3] for r in res]))
        os.replace(tmp, cache)
    if barrier is not None:
        barrier()
    z = np.load(cache)
    return cfg, z["depth"], z["color"], z["Tcw"]


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle = test infrastructure; this is one of the two places bench.py may run it)
# ------------------------------------------------------------------------------------------------

def cpu_port_fps(cfg, depth, color, Tcw, n_sample, threads, passes=2):
    """Steady-state frames/s of the CPU oracle port (OpenMP over touched blocks) on a bounded sample."""
    import oracle
    orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)
    n = min(n_sample, len(depth))
    for i in range(n):  # populate (allocation pass), untimed
        orc.integrate(depth[i], color[i], cfg.K, Tcw[i], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(passes):
        for i in range(n):
            orc.integrate(depth[i], color[i], cfg.K, Tcw[i], nthreads=threads)
    dt = time.perf_counter() - t0
    return passes * n / dt, n, passes


def cpu_ref_grid_fps(cfg, depth, color, Tcw, n_sample):
    """frames/s of the UNMODIFIED reference VoxelBlockGrid::integrate (point-average grid; 1 core,
    sequential branch) fed the world points the reference front-end derives from the same frames."""
    import oracle
    if not oracle.have_ref():
        return None
    g = oracle.RefGrid(cfg.voxel_size, 8)
    n = min(n_sample, len(depth))
    total = 0.0
    for i in range(n):
        d, c, T = depth[i], color[i], Tcw[i]
        valid = (d > 0.0) & (d < cfg.depth_trunc)
        z = d[valid].astype(np.float64)
        rows, cols = np.where(valid)
        x = (cols - cfg.cx) * z / cfg.fx
        y = (rows - cfg.cy) * z / cfg.fy
        Twc = S.inv_T(T)
        R, t = Twc[:3, :3], Twc[:3, 3]
        pw = np.stack([x * R[a, 0] + y * R[a, 1] + z * R[a, 2] + t[a] for a in range(3)], axis=1)
        total += g.integrate(pw.astype(np.float32), (c[valid] / 255.0).astype(np.float32))
    return n / total


def best_thread_count(cfg, depth, color, Tcw) -> int:
    """The port's per-frame allocation pass is serial (as Open3D's is), so more threads is not always
    faster: pick the OpenMP thread count with the best steady-state throughput on 4 frames."""
    import oracle
    hi = min(host_threads(), oracle.TsdfOracle.max_threads())
    cands = sorted({c for c in (4, 8, 16, 32, 64, hi) if c <= hi} | {hi})
    orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)
    n = min(4, len(depth))
    for i in range(n):
        orc.integrate(depth[i], color[i], cfg.K, Tcw[i], nthreads=hi)
    best, best_t = hi, float("inf")
    for rep in range(3):  # three rounds over the candidates, best time each: robust against host noise
        for c in cands:
            t0 = time.perf_counter()
            for i in range(n):
                orc.integrate(depth[i], color[i], cfg.K, Tcw[i], nthreads=c)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    return best


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------------

def run_reference_arm(args, rank):
    if rank != 0:
        return
    cfg, depth, color, Tcw = load_frames(args.config, args.frames, 0, 1)
    import oracle
    n_sample = min(args.cpu_frames, len(depth))
    threads = best_thread_count(cfg, depth, color, Tcw)
    orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)

    def one_step():
        for i in range(n_sample):
            orc.integrate(depth[i], color[i], cfg.K, Tcw[i], nthreads=threads)

    for _ in range(max(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    fps = args.steps * n_sample / dt
    sample = (f"first {n_sample} frames of the {args.config} sequence per step, steady state; "
              f"oracle port (C restatement of Open3D ScalableTSDFVolume under D1), OpenMP over blocks")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(cfg, args, n_sample, 1),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample, "cpu_model": cpu_model()},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cfg, args, frames_per_step, world):
    return {
        "workload": (f"{cfg.name} TUM1-shape {cfg.width}x{cfg.height} RGBD, voxel {cfg.voxel_size} m, "
                     f"sdf_trunc {cfg.sdf_trunc} m, depth_trunc {cfg.depth_trunc} m, block 8^3, "
                     f"alloc stride 4"),
        "frames_per_step": int(frames_per_step),
        "voxel_record_bytes": VOXEL_RECORD_BYTES,
        "parallelism": f"hash-shard x{world} (BlockKeyHash % N, no collective)" if world > 1 else "single GPU",
    }


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    from pyslam_b200 import B200TsdfVolume
    cfg, depth, color, Tcw = load_frames(args.config, args.frames, rank, world, barrier)
    F, H, W = depth.shape
    capacity = args.capacity
    vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=capacity,
                         device=local_rank, shard_rank=rank,
                         shard_count=args.shard_of if (world == 1 and args.shard_of > 1) else world)
    # a dedicated (non-default) stream: the legacy default stream has handle 0, which the C ABI reads
    # as "use the library's own stream" and which torch events would not observe
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    d_dev = torch.from_numpy(depth).cuda()
    c_dev = torch.from_numpy(color).cuda()
    # pinned host copies for the end-to-end path
    d_pin = torch.from_numpy(depth).pin_memory()
    c_pin = torch.from_numpy(color).pin_memory()
    d_pin_np, c_pin_np = d_pin.numpy(), c_pin.numpy()
    K = cfg.K

    def step_resident():
        # one C call enqueues the whole sequence (frames already in HBM) on torch's current stream
        vol.integrate_batch(d_dev, c_dev, K, Tcw, stream=stream.cuda_stream)

    def step_e2e():
        # public bulk API with pinned HOST frames: per-frame H2D on the copy stream overlaps the
        # kernels of the previous frame; then a D2H read of the step's result
        vol.integrate_batch(d_pin_np, c_pin_np, K, Tcw)
        return vol.last_frame_stats()  # syncs + D2H read of the device counters (64 B)

    # ---- warm-up (populates the map: steady state afterwards) ----
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    vol.synchronize()
    nb = vol.num_blocks()

    # ---- value: inputs resident in HBM, CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    for _ in range(2):  # keep the GPU under load while the sampler spins up
        step_resident()
    torch.cuda.synchronize()
    upd0, launches0 = vol.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    upd1, launches1 = vol.counters()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = args.steps * F / (ms_max * 1e-3)

    # ---- e2e: pinned host frames through the public API, H2D inside the timed region ----
    for _ in range(2):
        step_e2e()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
        vol.synchronize()   # every step ends with the D2H read of its result (the volume's counter block)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.steps * F / float(t.item())
    clocks = sampler.stop()  # sampled across the timed regions (resident + end-to-end)

    # ---- roofline: CUDA events around every integrate launch over passes of the same work ----
    def profile_pass(overlap, fusion):
        vol.set_overlap(overlap)
        vol.set_fusion(fusion)
        vol.profile_enable(True)
        u0, _ = vol.counters()
        b0 = vol.block_visits()
        for _ in range(min(args.steps, 3)):
            step_resident()
        torch.cuda.synchronize()
        a_ms, i_ms, nfr, nl = vol.profile_read()
        u1, _ = vol.counters()
        b1 = vol.block_visits()
        vol.profile_enable(False)
        upd, vis = u1 - u0, b1 - b0
        survey_bytes = 2 * VOXEL_RECORD_BYTES * 512 * upd + 7 * W * H * nfr       # SURVEY.md §8d formula
        moved_bytes = 2 * VOXEL_RECORD_BYTES * 512 * vis + 16 * W * H * nfr        # blocks visited + texels
        sec = i_ms * 1e-3
        return dict(alloc_ms=a_ms, integ_ms=i_ms, frames=nfr, launches=nl, updates=upd, visits=vis,
                    survey_bytes=survey_bytes, moved_bytes=moved_bytes,
                    gbs=survey_bytes / sec / 1e9 if sec > 0 else 0.0,
                    moved_gbs=moved_bytes / sec / 1e9 if sec > 0 else 0.0)

    situ = profile_pass(True, True)      # the timed-region schedule: fused groups, allocate overlapped
    iso = profile_pass(False, False)     # one frame per launch, kernels serialised: the HBM-bound kernel alone
    vol.set_overlap(True)
    vol.set_fusion(True)
    alloc_ms, integ_ms, nprof = iso["alloc_ms"], situ["integ_ms"], situ["frames"]
    block_updates, algo_bytes, achieved = situ["updates"], situ["survey_bytes"], situ["gbs"]
    peak, peak_src = 6650.0, "fallback"
    try:
        mp_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp_["hbm_gbs"]), "measured"
    except Exception:
        pass
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest.json")))
        traffic = prof.get("integrate_group_kernel", {}).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- mesh extraction, reported separately: marching cubes over the whole map ----
    mesh_info = None
    if world == 1:
        import ctypes as C
        nv_, nt_ = C.c_int64(0), C.c_int64(0)
        vol.synchronize()
        vol._L.b2v_extract_mesh(vol._h, C.byref(nv_), C.byref(nt_))   # warm-up (allocates scratch)
        t0 = time.perf_counter()
        vol._L.b2v_extract_mesh(vol._h, C.byref(nv_), C.byref(nt_))   # kernels + size read-back, no bulk copy
        mesh_ms = 1e3 * (time.perf_counter() - t0)
        mesh_info = {"ms_per_extract": mesh_ms, "vertices": int(nv_.value), "triangles": int(nt_.value),
                     "blocks": int(nb), "triangles_per_s": nt_.value / (mesh_ms * 1e-3),
                     "note": "b2v_extract_mesh on the populated map: neighbours, classify, scan, vertices, "
                             "triangles kernels + a 8-byte size read-back; arrays stay on the device"}

    # ---- secondary e2e figure: the same frames as RAW 16-bit depth (TUM-style payload, 5000 units per metre) ----
    e2e_u16 = None
    if world == 1:
        scale = np.float32(1.0 / 5000.0)
        raw16 = torch.from_numpy(np.round(depth * 5000.0).astype(np.uint16)).pin_memory()
        raw16_np = raw16.numpy()
        for _ in range(2):
            vol.integrate_batch(raw16_np, c_pin_np, K, Tcw, depth_scale=scale)
        vol.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            vol.integrate_batch(raw16_np, c_pin_np, K, Tcw, depth_scale=scale)
            vol.synchronize()
        dt16 = time.perf_counter() - t0
        e2e_u16 = {"value": args.steps * F / dt16, "unit": UNIT, "h2d_bytes_per_step": int(F * H * W * 5),
                   "d2h_bytes_per_step": 256,
                   "api": "B200TsdfVolume.integrate_batch(depths uint16, colors, K, poses, depth_scale) -> "
                          "b2v_integrate_batch_u16: raw 16-bit depth over PCIe, widened to float32 metres on the GPU "
                          "(bit-identical to depth.astype(float32) * depth_factor on the host)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    extra = {}
    if world == 1 and not args.no_cpu:
        threads = best_thread_count(cfg, depth, color, Tcw)
        fps, n_s, passes = cpu_port_fps(cfg, depth, color, Tcw, args.cpu_frames, threads)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"first {n_s} frames of the sequence, {passes} steady-state passes, oracle port "
                         f"(C restatement of Open3D ScalableTSDFVolume under D1), OpenMP over blocks"}
        fps1, _, _ = cpu_port_fps(cfg, depth, color, Tcw, min(args.cpu_frames, 16), 1, passes=1)
        extra["cpu_baseline_1core"] = {"value": fps1, "unit": UNIT, "cores": 1, "kind": "port"}
        rg = cpu_ref_grid_fps(cfg, depth, color, Tcw, 8)
        if rg is not None:
            extra["cpu_reference_point_average_grid"] = {
                "value": rg, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "8 frames, unmodified cpp/volumetric VoxelBlockGrid::integrate_raw<float,float> "
                          "(sequential branch; a different voxel model: point averaging, not TSDF)"}

    per_frame_blocks = block_updates / max(nprof, 1)
    fps_unfused = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(cfg, args, F, world),
                   "blocks_in_map": int(nb), "active_blocks_per_frame": per_frame_blocks,
                   "l2": (f"no flush: each step streams {nb * 10240 / 1e6:.0f} MB of voxel blocks "
                          f"(> 126 MB L2) between two visits of the same block"),
                   "timing": "CUDA events on the launching stream, max over ranks"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * H * W * 7),
                "d2h_bytes_per_step": 256, "timing": "wall clock around a full device sync",
                "api": "B200TsdfVolume.integrate_batch(depths, colors, K, poses) -> b2v_integrate_batch (pinned host frames)"},
        **({"e2e_u16_depth": e2e_u16} if e2e_u16 else {}),
        "gpu_launches": int(launches1 - launches0),
        "roofline": {
            "kernel": "integrate_group_kernel (up to 8 frames applied per block visit)", "bound": "hbm",
            "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "traffic": traffic,
            "algorithmic_bytes_per_launch": algo_bytes / max(situ["launches"], 1),
            "frames_per_launch": situ["frames"] / max(situ["launches"], 1),
            "avg_launch_us": 1e3 * integ_ms / max(situ["launches"], 1),
            "note": "achieved uses the per-frame SURVEY.md 8d formula 2*S*512*A_f + 7*W*H summed over the "
                    "frames of a launch; frac > 1 means the fused kernel moves fewer bytes than frame-by-frame "
                    "integration must (a block is read/written once per group instead of once per frame)",
            "bytes_moved_per_launch": situ["moved_bytes"] / max(situ["launches"], 1),
            "bytes_moved_gbs": situ["moved_gbs"],
            "block_visits_per_update": situ["visits"] / max(situ["updates"], 1),
            "measured": "in situ: CUDA events around every launch in the timed-region schedule "
                        "(allocate kernels of the next group run beside it)",
            "allocate_group_kernel_avg_us_in_situ": 1e3 * situ["alloc_ms"] / max(situ["launches"], 1),
            "per_frame_kernel": {
                "kernel": "integrate_kernel (one frame per launch, b2v_set_fusion 0, b2v_set_overlap 0)",
                "bound": "hbm", "achieved": iso["gbs"], "frac": iso["gbs"] / peak if peak else None,
                "avg_launch_us": 1e3 * iso["integ_ms"] / max(iso["launches"], 1),
                "allocate_kernel_avg_us": 1e3 * iso["alloc_ms"] / max(iso["frames"], 1)}},
        "clocks": clocks,
        **({"diagnostic": f"--shard-of {args.shard_of}: this process is rank 0 of a {args.shard_of}-way sharded job "
                          f"(value = that job's per-rank rate = its whole-job rate, ranks share nothing)"}
           if (world == 1 and args.shard_of > 1) else {}),
        "mesh": mesh_info,
        "cpu_baseline": cpu,
        **extra,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--frames", type=int, default=300, help="frames per step (the sequence length)")
    ap.add_argument("--capacity", type=int, default=1 << 19, help="block-pool capacity (10 KiB each)")
    ap.add_argument("--cpu-frames", type=int, default=32, help="frames in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shard-of", type=int, default=0,
                    help="diagnostic (N=1 only): act as rank 0 of a --shard-of-way sharded job on one GPU; ranks share "
                         "nothing, so this is the per-rank work of that job")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
