#!/usr/bin/env python
"""bench.py — headline benchmark of the pairwise SIFT-128 matching hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one JSON line on rank 0)
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's own CPU engine on the host cores

Metric (BASELINE.json): image-pairs/sec on 8k x 8k SIFT-128.  A *step* is one pass of the hot path over the whole
candidate pair list of the workload: BASELINE config #2, exhaustive matching of 100 web images x 8192 descriptors
(4,950 pairs) per GPU — 2-NN + ratio (0.85 "all" / 0.6 "good", fine_matching_graph.cc:42-43) + mutual cross-check.
  value : pairs/s with the packed descriptor table already resident in HBM (device-timed, CUDA events on the
          library's stream, L2 flushed between steps, max over ranks)
  e2e   : pairs/s through the public API from HOST buffers: per step pack+upload every image from pinned memory
          (H2D), replicate the table over NCCL when N > 1, match, and copy the match lists back (D2H)
  roofline: the tcgen05 matching kernel against the dense int8 tensor-core peak
  cpu_baseline: the reference's vendored exact kNN engine (oracle/_ref, nanoflann) on a bounded sample, same inputs
Multi-GPU (weak scaling): every rank owns 100 images and 4,950 pairs; the table is replicated once
(torch.distributed/NCCL broadcast into the library's arena), then ranks match independently (no data-path collective).
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

RATIO_ALL, RATIO_GOOD = 0.85, 0.6
INT8_SPEC_PEAK_TOPS = 4500.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--images", type=int, default=100, help="images per GPU")
    ap.add_argument("--rows", type=int, default=8192, help="descriptors per image")
    ap.add_argument("--cpu-sample-pairs", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mutual", type=int, default=1, help="1 = ratio + mutual cross-check (headline), 0 = ratio only")
    ap.add_argument("--no-int8-peak", action="store_true", help="skip the cuBLAS int8 GEMM peak measurement (rank 0, N = 1)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic_bytes(pairs_per_step: int, launches_per_step: float):
    """DRAM bytes per launch of the matching kernel, scaled from the committed `ncu --set full` capture
    (profiles/r1_match_kernel_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one forward launch and the
    number of pairs it covered).  None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "r1_match_kernel_traffic.json")
    if not os.path.exists(path) or launches_per_step <= 0:
        return None
    with open(path) as f:
        t = json.load(f)
    return t["dram_bytes_per_pair"] * pairs_per_step / launches_per_step


def measure_int8_peak(dev):
    """Dense int8 tensor-core rate of this GPU as cuBLAS delivers it (torch._int_mm, 8192^3, int32 accumulate), with the
    protocol of MEASURED_PEAKS.json: best of 10 (burst) and back to back for ~2 s (sustained).  TOP/s or None."""
    import torch
    try:
        n = 8192
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = max(10, int(2000.0 / best))
        e0.record()
        for _ in range(iters):
            torch._int_mm(a, b)
        e1.record(); e1.synchronize()
        ops = 2.0 * n ** 3
        return {"burst_tops": ops / (best * 1e-3) / 1e12, "sustained_tops": ops * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": "torch._int_mm 8192^3 (cuBLASLt int8, s32 accumulate): best of 10 and back to back for ~2 s"}
    except Exception as exc:  # noqa: BLE001
        return {"error": str(exc)[:200]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.lines, self.proc, self.enabled = index, [], None, enabled
        self.t0 = self.t1 = None   # timed window (perf_counter)

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self) -> dict:
        if not self.enabled:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampled on rank 0 only"]}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        # samples inside the timed window; nvidia-smi needs ~1 s to come up, so it is started before the warm-up steps and
        # a window shorter than its period falls back to the warm-up + timed span (same kernels, same load)
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        inside = [ln for t, ln in self.lines if t0 <= t <= t1 + 0.1]
        window = "timed region"
        if not inside:
            inside, window = [ln for t, ln in self.lines if t <= t1 + 0.1], "warm-up + timed region (timed region shorter than the sampling period)"
        sm, mx, pw, reasons = [], [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm),
                "window": window, "reasons": sorted(reasons)}


# ====================================================================================================== reference arm
def cpu_reference_sample(images, pairs, n_pairs: int):
    """Time the reference's vendored exact engine (nanoflann; oracle/_ref) — or the oracle port when it is absent —
    on the first n_pairs pairs of the workload, all host threads.  Returns (pairs_per_s, kind, cores, seconds)."""
    from oracle import oracle
    oracle.use_all_host_threads()   # torchrun exports OMP_NUM_THREADS=1; the reference uses every core (OpenMP default)
    use_ref = oracle.ref_available()
    cores = oracle.ref_lib().ref_nanoflann_max_threads() if use_ref else oracle.max_threads()
    f32 = {}
    t0 = time.perf_counter()
    for r, q in pairs[:n_pairs]:
        if use_ref:
            for i in (r, q):
                if i not in f32:
                    f32[i] = images[i].astype(np.float32)   # the reference's container: CV_32FC1 rows
            # tree on idx1, query idx2 rows, ratio test (feature_matching.cpp:319-342; fine_matching_graph.cc:72-133)
            oracle.ref_match(f32[r], f32[q], RATIO_ALL, 20)
        else:
            oracle.match_pair_u8(images[r], images[q], RATIO_ALL, mutual=True, ratio_good=RATIO_GOOD)
    dt = time.perf_counter() - t0
    return n_pairs / dt, ("reference" if use_ref else "port"), cores, dt


def flann_forest_sample(images, pairs, n_pairs: int):
    """Behavioural twin of the reference's PRODUCTION kNN (fine_matching_graph.cc:72-99: FLANN randomized KD forest, 8 trees,
    64 checks, k = 2) through the cv2.flann_Index build of this image: seconds per pair for index build (once per idx1 in
    the reference) and for the batch query of one partner.  The reference runs one such query per OpenMP thread
    (fine_matching_graph.cc:87).  None when cv2 is unavailable."""
    try:
        import cv2
    except Exception:  # noqa: BLE001
        return None
    build_s = query_s = 0.0
    for r, q in pairs[:n_pairs]:
        ref = np.ascontiguousarray(images[r], dtype=np.float32)
        qry = np.ascontiguousarray(images[q], dtype=np.float32)
        t0 = time.perf_counter()
        index = cv2.flann_Index(ref, dict(algorithm=1, trees=8))
        t1 = time.perf_counter()
        index.knnSearch(qry, 2, params=dict(checks=64))
        t2 = time.perf_counter()
        build_s += t1 - t0
        query_s += t2 - t1
    n = max(1, min(n_pairs, len(pairs)))
    return {"kd_forest_build_s_per_image": build_s / n, "kd_forest_query_s_per_pair": query_s / n,
            "what": "one cv2.flann_Index(KDTREE, trees=8) build + one knnSearch(k=2, checks=64) call per pair (approximate search; the "
                    "production path of fine_matching_graph.cc:72-99 runs one such query per OpenMP thread)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from metricsfm_b200 import synth
    col = synth.Collection(args.rows, seed=0)
    per_step = 2
    n_img = min(args.images, 2 * per_step * (args.steps + args.warmup))
    images = {i: col.image_u8(i) for i in range(n_img)}
    pairs = synth.exhaustive_pairs(n_img)
    vals = []
    kind, cores = "port", 1
    k = 0
    for s in range(args.warmup + args.steps):
        sel = [tuple(pairs[(k + j) % len(pairs)]) for j in range(per_step)]
        k += per_step
        v, kind, cores, dt = cpu_reference_sample(images, sel, per_step)
        if s >= args.warmup:
            vals.append((per_step, dt))
    total_pairs = sum(p for p, _ in vals)
    total_s = sum(d for _, d in vals)
    value = total_pairs / total_s
    line = {
        "impl": "reference", "metric": "image-pairs/sec (8k x 8k SIFT-128)", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"exhaustive {args.images} web images x {args.rows} SIFT-128 (BASELINE config #2), "
                               f"bounded sample of {per_step} pairs per step"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": f"{per_step} pairs of {args.rows}x{args.rows} per step x {args.steps} steps: nanoflann exact KD-tree "
                                   f"on idx1 + 2-NN of idx2 rows + ratio {RATIO_ALL} (feature_matching.cpp:319-342), OpenMP over queries"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ====================================================================================================== native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from metricsfm_b200 import distributed as D, scheduler, synth
    from metricsfm_b200.matcher import Matcher

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = {"pinned": False, "why": "single process: the CPU baseline keeps every host core"}
    if world > 1:
        numa = D.pin_to_gpu_numa_node(local_rank)   # before the page-locked staging buffers are allocated
        dist.init_process_group("nccl", device_id=dev)

    n_local, rows = args.images, args.rows
    n_global = n_local * world
    rows_padded = (rows + 255) // 256 * 256
    arena_rows = n_global * rows_padded

    # ---- synthetic inputs (seeded; identical on every rank for a given image id), staged in pinned host memory
    col = synth.Collection(rows, seed=0)
    my_ids = list(range(rank * n_local, (rank + 1) * n_local))
    host_desc = torch.empty((n_local, rows, 128), dtype=torch.uint8).pin_memory()
    for k, gid in enumerate(my_ids):
        host_desc[k].numpy()[:] = col.image_u8(gid)

    # ---- global pair list: exhaustive within each rank's block of images (4,950 pairs per block); LPT-sharded
    base = synth.exhaustive_pairs(n_local)
    pairs = np.concatenate([base + b * n_local for b in range(world)], axis=0)
    rows_per_image = np.full((n_global,), rows, np.int64)
    shards = scheduler.shard_pairs(pairs, rows_per_image, world)
    my_pairs = pairs[shards[rank]]

    # ---- packed table lives in torch-owned memory so NCCL can fill it during replication
    desc_arena = torch.empty((arena_rows, 128), dtype=torch.uint8, device=dev)
    norm_arena = torch.empty((arena_rows,), dtype=torch.int32, device=dev)
    m = Matcher(device=local_rank, max_images=n_global, arena_rows=arena_rows, external_desc_arena=desc_arena.data_ptr(),
                external_norm_arena=norm_arena.data_ptr())
    lib_stream = torch.cuda.ExternalStream(m.cuda_stream(), device=dev)

    owner, ranges = D.block_ranges(n_global, rows_padded, world)

    def stage_table():
        """Pack + upload this rank's images (H2D), reserve the others, replicate over NCCL.  Returns H2D bytes."""
        m.release_all()
        mine = [gid for gid in range(n_global) if owner[gid] == rank]
        before = [gid for gid in range(n_global) if gid < mine[0]]
        after = [gid for gid in range(n_global) if gid > mine[-1]]
        # identical allocation order on every rank => identical arena offsets; one call per block of foreign images
        if before:
            m.reserve_batch(before, [rows] * len(before))
        # page-locked host rows, no host wait: the transfer overlaps the pair planning of match_pairs
        m.upload_batch(mine, [host_desc[g - rank * n_local] for g in mine], wait=False)
        if after:
            m.reserve_batch(after, [rows] * len(after))
        if world > 1:
            lib_stream.synchronize()
            D.replicate_arena(desc_arena, norm_arena, ranges, dist)   # NCCL broadcast per owner block
            torch.cuda.synchronize()
        return n_local * rows * 128

    stage_table()
    kw = dict(ratio_good=RATIO_GOOD, mutual=bool(args.mutual), min_keypoints=20, orientation=0)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    sampler.start()
    for _ in range(args.warmup):
        m.match_pairs_resident(my_pairs, RATIO_ALL, **kw)
    barrier()
    sampler.mark_begin()
    step_ms, kern_ms, launches, match_launches, ops = [], [], 0, 0, 0
    n_matches = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        with torch.cuda.stream(lib_stream):
            flush.zero_()                       # L2 flush between timed iterations (not timed)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(lib_stream)
        n_matches = m.match_pairs_resident(my_pairs, RATIO_ALL, **kw)
        e1.record(lib_stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        t = m.timing()
        kern_ms.append(t["match_kernel_ms"])
        launches += t["total_launches"]
        match_launches += t["match_launches"]
        ops = t["int8_ops"]
    barrier()
    sampler.mark_end()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
        cnt = torch.tensor([len(my_pairs) * args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        pairs_done = float(cnt.item())
    else:
        pairs_done = float(len(my_pairs) * args.steps)
    value = pairs_done / (total_ms * 1e-3)

    # ---- end to end through the public API from host buffers
    e2e = None
    if not args.no_e2e:
        from metricsfm_b200.matcher import MatchResult
        e2e_steps = max(1, min(args.steps, 3))
        cap = len(my_pairs) * 2048
        out = MatchResult(offsets=np.zeros((len(my_pairs) + 1,), np.int64), ok=np.zeros((len(my_pairs),), np.int32),
                          matches=torch.empty((cap, 2), dtype=torch.int32).pin_memory().numpy(),
                          good=torch.empty((cap,), dtype=torch.uint8).pin_memory().numpy())
        d2h = 0
        h2d = 0
        for it in range(1 + e2e_steps):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            h2d = stage_table() + my_pairs.nbytes
            res = m.match_pairs(my_pairs, RATIO_ALL, out=out, **kw)
            d2h = m.timing()["d2h_bytes"]
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": (len(pairs) * e2e_steps) / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "timer": "host wall clock between barriers + cuda synchronize, max over ranks",
               "matches_per_step": int(len(res.matches))}

    # ---- CPU baseline (rank 0, N = 1 only): the reference's exact engine on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        images = {i: host_desc[i].numpy() for i in range(min(n_local, 2 * args.cpu_sample_pairs))}
        sel = [tuple(p) for p in pairs if p[0] in images and p[1] in images][: args.cpu_sample_pairs]
        v, kind, cores, secs = cpu_reference_sample(images, sel, len(sel))
        cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind,
               "sample": f"first {len(sel)} pairs of the workload ({rows}x{rows}), {secs:.1f} s: nanoflann exact KD-tree on idx1 + 2-NN of "
                         f"idx2 rows + ratio {RATIO_ALL}, OpenMP over queries (reference engine compiled from its own headers)"}
        forest = flann_forest_sample(images, sel, 2)
        if forest is not None:
            forest["extrapolated_pairs_per_s_one_query_per_core"] = cores / forest["kd_forest_query_s_per_pair"]
            cpu["production_kd_forest"] = forest

    int8_peak = None
    if rank == 0 and world == 1 and not args.no_int8_peak:
        m.close()                      # the library's scratch is not needed any more
        int8_peak = measure_int8_peak(dev)
    if rank == 0:
        peaks, peaks_src = measured_peaks()
        kern_avg_ms = float(np.mean(kern_ms))
        achieved_tops = ops / (kern_avg_ms * 1e-3) / 1e12
        # int8 dense rate = 2 x the bf16 rate on the same tensor pipes; the driver measures bf16 only.  The kernel is timed
        # inside a long step, so the sustained figure is the denominator.
        peak_tops = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        line = {
            "metric": "image-pairs/sec (8k x 8k SIFT-128)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"exhaustive matching of {n_local} web images x {rows} SIFT-128 descriptors ({len(base)} pairs) per GPU "
                                   f"(BASELINE config #2); 2-NN + ratio {RATIO_ALL}/{RATIO_GOOD}" + (" + mutual cross-check" if args.mutual else " (no mutual check)"),
                       "pairs_per_step_all_gpus": int(len(pairs)), "parallelism": f"pair-sharded x{world}, table replicated",
                       "l2": "flushed between timed steps (256 MiB write)", "matches_per_step_rank0": int(n_matches),
                       "host_numa_pinning_rank0": numa},
            "roofline": {"bound": "tensor", "achieved": achieved_tops, "peak": peak_tops, "unit": "TFLOP/s", "frac": achieved_tops / peak_tops,
                         "traffic": ncu_traffic_bytes(len(my_pairs), match_launches / max(args.steps, 1)),
                         "kernel": "match_pairs_kernel<4,64,8,1,2> (4 strips x N=64 tiles, 8 B stages, 2 TMEM buffers per strip)",
                         "note": "int8 tensor ops (2 per MAC), i.e. TOP/s; algorithmic ops = 2*M*N*128 per pair; peak = 2 x "
                                 f"bf16_tflops_sustained of MEASURED_PEAKS.json ({peaks_src}); spec dense int8 = 4500",
                         "frac_of_spec_int8": achieved_tops / INT8_SPEC_PEAK_TOPS, "kernel_ms_per_step": kern_avg_ms,
                         "kernel_share_of_step": kern_avg_ms / (float(np.mean(step_ms)))},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "match_kernel_launches": int(match_launches),
            "wall_s_timed_region": wall_s,
        }
        if int8_peak is not None:
            line["roofline"]["int8_gemm_measured"] = int8_peak
            if "sustained_tops" in int8_peak:
                line["roofline"]["frac_of_measured_int8_sustained"] = achieved_tops / int8_peak["sustained_tops"]
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=JSON_OUT, flush=True)
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


JSON_OUT = sys.stdout


def claim_stdout():
    """Rank 0 prints exactly ONE JSON line on stdout.  Native libraries (NCCL's version banner, ...) write to file
    descriptor 1 directly, so keep a private copy of the real stdout for the JSON line and point fd 1 at stderr."""
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
