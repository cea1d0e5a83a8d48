#!/usr/bin/env python
"""bench.py — the Raycaster micro-bench of BASELINE.json configs[1] on B200 (SURVEY.md §8d "C2").

Workload (one "step"): a 1,002,528-triangle displaced grid, reference-order BVH; 4,194,304 coherent primary rays and
4,194,304 incoherent bounce rays, each traced closest-hit (Raycaster::intersects) AND any-hit
(Raycaster::intersectsAny): 16,777,216 ray queries per step and per GPU. Metric: Mrays/s.

  value : device-resident — rays and hit buffers already in HBM, CUDA-event time on the launching stream.
  e2e   : the same step through the host-pointer C ABI (rb_raycaster_intersect / _intersect_any) with pinned host
          buffers: H2D of the rays and D2H of the Raycaster::Hit records inside the timed region.
  roofline     : dominant kernel, algorithmic bytes / live CUDA-event duration vs measured HBM peak.
  cpu_baseline : the reference's own CPU code (oracle/_ref, else the C oracle) on a bounded sample, host threads stated.

`--impl reference` times the reference CPU implementation alone (same metric/config, bounded sample per step).
Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU, weak scaling:
every rank traces the full ray sets against its own replica of the scene; no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_N = 708          # 2*708^2 = 1,002,528 triangles
RES = 2048            # 2048^2 = 4,194,304 rays per set
MINI, MAXI = 1e-4, 1e8
SAMPLE_ROW_STRIDE = 16  # CPU legs trace every 16th row of each set: 262,144 rays per set


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries the JSON line and nothing else: the reference library logs with std::cout ("[Raycaster] Mesh 0
# added ..."), so file descriptor 1 is pointed at stderr for the whole run and the line is written to a private
# duplicate of the original stdout.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _claim_stdout()
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs. The process is started well before the
    timed region (its NVML initialisation stalls kernel launches for ~100 ms) and only the samples between begin() and
    end() are reported."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc, self.i0, self.i1 = index, [], None, 0, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:  # wait for the first sample: initialisation is over
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def begin(self):
        self.i0 = len(self.lines)

    def end(self):
        time.sleep(0.06)
        self.i1 = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.i1 is None:
            self.end()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        window = self.lines[self.i0:self.i1] or self.lines[-2:]
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def sample_rows():
    return np.arange(0, RES, SAMPLE_ROW_STRIDE)


def sample_index():
    rows = sample_rows()
    return (rows[:, None] * RES + np.arange(RES)[None, :]).reshape(-1)


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_tracer(pos, idx):
    """The reference's own Raycaster when oracle/_ref was built, else the C restatement."""
    from oracle import ref
    if ref.available():
        r = ref.RefRaycaster()
        r.add_mesh(pos, idx)
        r.build()
        threads = ref.threads_default()  # the reference's policy: hardware_concurrency()-1 (System.hpp:82)

        def closest(o, d):
            return r.intersect(o, d, MINI, MAXI, threads=threads)

        def anyhit(o, d):
            return r.intersect_any(o, d, MINI, MAXI, threads=threads)[0]
        return "reference", threads, closest, anyhit
    from oracle import oracle as orc
    o_ = orc.OracleRaycaster()
    o_.add_mesh(pos, idx)
    o_.build()
    threads = max((os.cpu_count() or 2) - 1, 1)

    def closest(o, d):
        h = o_.intersect(o, d, MINI, MAXI, threads=threads)
        return dict(hit=h["hit"], dist=h["dist"], local_id=h["localId"])

    def anyhit(o, d):
        return o_.intersect_any(o, d, MINI, MAXI, threads=threads)
    return "port", threads, closest, anyhit


def cpu_sample_sets(pos, idx, closest):
    from rendu_b200 import workloads as wl
    sel = sample_index()
    co, cd = wl.coherent_rays(RES)
    co, cd = np.ascontiguousarray(co[sel]), np.ascontiguousarray(cd[sel])
    h = closest(co, cd)
    nrm = wl.triangle_normals(pos, idx, np.where(h["hit"] == 1, h["local_id"], 0))
    io, idr = wl.incoherent_rays(co, cd, h["hit"], h["dist"], nrm, seed=1234, index_base=0)
    return co, cd, np.ascontiguousarray(io), np.ascontiguousarray(idr)


def cpu_step(sets, closest, anyhit):
    co, cd, io, idr = sets
    t0 = time.perf_counter()
    closest(co, cd)
    anyhit(co, cd)
    closest(io, idr)
    anyhit(io, idr)
    return time.perf_counter() - t0, 4 * co.shape[0]


def config_dict(n_gpus):
    return {"workload": "C2 raycaster micro-bench: 1,002,528-triangle displaced grid (N=708), reference-order BVH, "
                        "4,194,304 coherent + 4,194,304 incoherent rays, closest-hit and any-hit (16,777,216 queries/step/GPU)",
            "triangles": 2 * GRID_N * GRID_N, "rays_per_set": RES * RES, "queries_per_step_per_gpu": 4 * RES * RES,
            "mini": MINI, "maxi": MAXI, "trace_mode": "exact (reference traversal order, bit-identical hits)",
            "parallelism": "replicated scene, %d independent rank(s), no collective" % n_gpus,
            "l2": "no explicit flush: one step streams ~0.7 GB (rays 201 MB + Hit records 403 MB + any flags 8 MB + 82 MB tree) through a 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rendu_b200 import workloads as wl
    pos, idx = wl.displaced_grid(GRID_N)
    kind, threads, closest, anyhit = cpu_tracer(pos, idx)
    sets = cpu_sample_sets(pos, idx, closest)
    for _ in range(args.warmup):
        cpu_step(sets, closest, anyhit)
    total_t, total_q = 0.0, 0
    for _ in range(args.steps):
        t, q = cpu_step(sets, closest, anyhit)
        total_t += t
        total_q += q
    value = total_q / total_t / 1e6
    sample = "every %dth row of both ray sets (%d rays/set, %d queries/step), closest+any" % (SAMPLE_ROW_STRIDE, sets[0].shape[0], 4 * sets[0].shape[0])
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import rendu_b200 as rb
    from rendu_b200 import workloads as wl
    from rendu_b200._capi import check, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — rendu_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    check(lib().rb_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t_setup = time.time()
    pos, idx = wl.displaced_grid(GRID_N)
    rc = rb.Raycaster()
    rc.addMesh(pos, idx)
    rc.updateHierarchy(rb.BUILD_REFERENCE)
    co, cd = wl.coherent_rays(RES)
    n = co.shape[0]
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def dev_f32(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    d_co, d_cd = dev_f32(co), dev_f32(cd)
    d_hit_c = torch.empty(n * 48, dtype=torch.uint8, device=dev)
    d_hit_i = torch.empty(n * 48, dtype=torch.uint8, device=dev)
    d_any_c = torch.empty(n, dtype=torch.uint8, device=dev)
    d_any_i = torch.empty(n, dtype=torch.uint8, device=dev)
    # Incoherent set: bounce rays from the coherent set's closest hits (traced by the CUDA path itself).
    rc.intersects_device(d_co.data_ptr(), d_cd.data_ptr(), n, d_hit_c.data_ptr(), MINI, MAXI, sh)
    torch.cuda.synchronize()
    hits = d_hit_c.cpu().numpy().view(rb.HIT_DTYPE)
    nrm = wl.triangle_normals(pos, idx, np.where(hits["hit"] == 1, hits["localId"], 0))
    io, idr = wl.incoherent_rays(co, cd, hits["hit"], hits["dist"], nrm, seed=1234, index_base=0)
    d_io, d_id = dev_f32(io), dev_f32(idr)
    log("[bench] rank %d setup %.1fs: %d tris, %d nodes, coherent hit rate %.3f" % (rank, time.time() - t_setup, rc.sizes()["tris"], rc.sizes()["nodes"], float((hits["hit"] == 1).mean())))

    launches = [
        ("closest_coherent", lambda: rc.intersects_device(d_co.data_ptr(), d_cd.data_ptr(), n, d_hit_c.data_ptr(), MINI, MAXI, sh)),
        ("any_coherent", lambda: rc.intersectsAny_device(d_co.data_ptr(), d_cd.data_ptr(), n, d_any_c.data_ptr(), MINI, MAXI, sh)),
        ("closest_incoherent", lambda: rc.intersects_device(d_io.data_ptr(), d_id.data_ptr(), n, d_hit_i.data_ptr(), MINI, MAXI, sh)),
        ("any_incoherent", lambda: rc.intersectsAny_device(d_io.data_ptr(), d_id.data_ptr(), n, d_any_i.data_ptr(), MINI, MAXI, sh)),
    ]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        for _, f in launches:
            f()
    barrier()

    # ---- timed region 1: device-resident (value)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in launches] for _ in range(args.steps)]
    l0 = rc.launch_count()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e_start.record(stream)
    for s in range(args.steps):
        for k, (_, f) in enumerate(launches):
            ev[s][k][0].record(stream)
            f()
            ev[s][k][1].record(stream)
    e_stop.record(stream)
    barrier()
    sampler.end()
    total_ms = e_start.elapsed_time(e_stop)
    gpu_launches = rc.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    per_kernel_ms = [float(np.mean([ev[s][k][0].elapsed_time(ev[s][k][1]) for s in range(args.steps)])) for k in range(len(launches))]

    # ---- timed region 2: end to end through the host-pointer ABI with pinned buffers (e2e)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_co, h_cd, h_io, h_id = pin(co), pin(cd), pin(io), pin(idr)
    h_hits = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
    h_any = torch.empty(n, dtype=torch.uint8).pin_memory()
    L = lib()

    # Two host threads, as the reference's callers do (intersects/intersectsAny are called concurrently from the
    # forParallel workers, PathTracer.cpp:186,204): one issues the closest-hit calls (download-heavy: 48 B Hit per ray),
    # the other the any-hit calls (upload-heavy), so both directions of the host link stay busy. ctypes drops the GIL.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=2)

    def e2e_closest():
        for o_, d_ in ((h_co, h_cd), (h_io, h_id)):
            check(L.rb_raycaster_intersect(rc._h, o_.data_ptr(), d_.data_ptr(), MINI, MAXI, n, h_hits.data_ptr()))

    def e2e_any():
        for o_, d_ in ((h_co, h_cd), (h_io, h_id)):
            check(L.rb_raycaster_intersect_any(rc._h, o_.data_ptr(), d_.data_ptr(), MINI, MAXI, n, h_any.data_ptr()))

    def e2e_step():
        if os.environ.get("RB_BENCH_E2E_THREADS", "2") == "1":
            e2e_closest()
            e2e_any()
            return
        fa, fb = pool.submit(e2e_closest), pool.submit(e2e_any)
        fa.result()
        fb.result()

    def copy_gbs(dst, src):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0_ = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return src.numel() * src.element_size() / (time.perf_counter() - t0_) / 1e9

    link = {"h2d_gbs": copy_gbs(d_hit_c, h_hits), "d2h_gbs": copy_gbs(h_hits, d_hit_c)}
    # Both directions at once (what the pipelined e2e step asks of the link): a second pinned buffer goes up on one
    # stream while the hit records come down on another.
    h_up = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n * 48, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    t0_ = time.perf_counter()
    with torch.cuda.stream(s_up):
        d_up.copy_(h_up, non_blocking=True)
    with torch.cuda.stream(s_dn):
        h_hits.copy_(d_hit_c, non_blocking=True)
    torch.cuda.synchronize()
    link["duplex_gbs_each_way"] = n * 48 / (time.perf_counter() - t0_) / 1e9
    del h_up, d_up
    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = 4 * n * 24
    d2h = 2 * n * 48 + 2 * n

    # ---- max over ranks
    times = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(times[0]), float(times[1])
    queries_step = 4 * n
    value = world * queries_step * args.steps / (total_ms * 1e-3) / 1e6
    e2e_value = world * queries_step * e2e_steps / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        # Work counters (instrumented kernel, outside the timed regions) -> algorithmic bytes per launch.
        work = []
        for (name, _), (o_, d_, anyh) in zip(launches, ((d_co, d_cd, False), (d_co, d_cd, True), (d_io, d_id, False), (d_io, d_id, True))):
            w = rc.count_work_device(o_.data_ptr(), d_.data_ptr(), n, anyh, MINI, MAXI, sh)
            out_b = 1 if anyh else 48
            w["bytes"] = n * (24 + out_b) + 64 * w["interior"] + 48 * w["tris"] + 32 * w["roots"]
            w["name"] = name
            work.append(w)
        dom = int(np.argmax(per_kernel_ms))
        peak, peak_src = measured_hbm_peak()
        achieved = work[dom]["bytes"] / (per_kernel_ms[dom] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:  # DRAM bytes of the same launch from the committed ncu --set full capture (not measured in this run)
            rec = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))
            traffic = rec[work[dom]["name"]]["dram_bytes_read"] + rec[work[dom]["name"]]["dram_bytes_write"]
            traffic_src = rec["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "rb_k_closest<exact>" if "closest" in work[dom]["name"] else "rb_k_any<exact>", "launch": work[dom]["name"],
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": work[dom]["bytes"], "avg_launch_ms": per_kernel_ms[dom],
                    "per_ray": {"interior_nodes": work[dom]["interior"] / n, "triangle_tests": work[dom]["tris"] / n,
                                "bytes": work[dom]["bytes"] / n},
                    "note": "the 82 MB tree is L2-resident, so bytes/s above is served mostly by L2, not HBM; the kernel's limiter is the L1 data pipe (l1tex__data_pipe_lsu_wavefronts at 85-90 % of one per cycle, profiles/r01_l1_data_pipe.txt), then issue slots (54 %)"}
        per_launch = {w["name"]: {"ms": per_kernel_ms[k], "mrays_s": n / (per_kernel_ms[k] * 1e-3) / 1e6,
                                  "interior_per_ray": w["interior"] / n, "tri_tests_per_ray": w["tris"] / n,
                                  "gbs": w["bytes"] / (per_kernel_ms[k] * 1e-3) / 1e9} for k, w in enumerate(work)}
        cpu = None
        if world == 1 and not args.no_cpu:
            kind, threads, closest, anyhit = cpu_tracer(pos, idx)
            sets = cpu_sample_sets(pos, idx, closest)
            cpu_step(sets, closest, anyhit)
            tt, qq = 0.0, 0
            while tt < 10.0:
                t, q = cpu_step(sets, closest, anyhit)
                tt += t
                qq += q
            cpu = {"value": qq / tt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind, "host_cpus": os.cpu_count(),
                   "sample": "every %dth row of both ray sets (%d rays/set), closest+any, repeated for %.1f s" % (SAMPLE_ROW_STRIDE, sets[0].shape[0], tt)}
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_dict(world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "api": "rb_raycaster_intersect / rb_raycaster_intersect_any with pinned host buffers, closest and any calls issued from two host threads", "host_link": link},
                "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "per_launch": per_launch}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ path-tracing workload (C3 / C4)
PT_W, PT_H, PT_DEPTH = 1920, 1080, 8


def pt_config(world, spp, split, trace_mode="exact"):
    return {"workload": "C3/C4 path trace: 1920x1080, %d spp, depth 8, synthetic textured GGX+Lambert Cornell room (cubemap environment, emissive panel, "
                        "shadowed point light, NEE on) with the 1,002,528-triangle displaced grid as floor" % spp,
            "width": PT_W, "height": PT_H, "spp": spp, "depth": PT_DEPTH, "split": split if world > 1 else "none", "trace_mode": trace_mode,
            "parallelism": "replicated scene, %d rank(s), %s" % (world, "tile split, bands gathered on rank 0" if split == "tile" else "sample split, one NCCL reduce of the fp32 sums"),
            "l2": "wavefront state (4M paths x 176 B) streams through HBM every bounce; no explicit flush"}


def run_pathtrace(args):
    import torch
    import rendu_b200 as rb
    from rendu_b200 import distributed as rd
    from rendu_b200 import pathtracer as rpt
    from rendu_b200 import workloads as wl
    from rendu_b200._capi import check, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — rendu_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    check(lib().rb_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    spp = args.spp
    desc = wl.cornell_synthetic(GRID_N if not args.small else 64, 512, 256)
    cam = wl.cornell_camera(PT_W / PT_H)
    scene = rpt.Scene(desc)
    pt = rpt.PathTracer(scene)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    d_sum = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32, device=dev)
    d_img = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32, device=dev)
    h_img = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32).pin_memory()

    tmode = {"exact": 0, "ordered": 1, "wide": 2}[args.trace_mode]
    if tmode == 2:  # experiment: closest-hit wavefront on the eight-wide quantised tree (shadow rays keep the binary one)
        scene.raycaster().build_wide()

    def opts_for(n_spp):
        if world > 1 and args.split == "tile":
            return rpt.make_opts(seed=1, tile=rd.tile_spec(rank, world), trace_mode=tmode, max_paths=args.max_paths)
        if world > 1:
            return rpt.make_opts(seed=1, samples=rd.sample_spec(rank, world), trace_mode=tmode, max_paths=args.max_paths)
        return rpt.make_opts(seed=1, trace_mode=tmode, max_paths=args.max_paths)

    def frame(n_spp, to_host=False):
        o = opts_for(n_spp)
        pt.render_device(cam, PT_W, PT_H, n_spp, PT_DEPTH, d_sum.data_ptr(), o, sh)
        final = None
        if world > 1 and args.split == "sample":
            total = rd.combine_sample_split(d_sum, rank, world)
            if rank == 0:
                d_img.zero_()
                pt.resolve_device(total.data_ptr(), PT_W, PT_H, n_spp, 3, d_img.data_ptr(), None, sh)
                final = d_img
        else:
            d_img.zero_()
            pt.resolve_device(d_sum.data_ptr(), PT_W, PT_H, n_spp, 3, d_img.data_ptr(), o, sh)
            if world > 1:
                rows = rd.local_rows(PT_H, rank, world).to(dev)
                final = rd.combine_tile_split(d_img[rows].contiguous(), PT_H, rank, world)
            else:
                final = d_img
        if to_host and rank == 0:
            h_img.copy_(final, non_blocking=True)
        return final

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        frame(16)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record(stream)
    for _ in range(args.steps):
        frame(spp)
    e1.record(stream)
    barrier()
    sampler.end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    st = pt.stats()
    # end to end: the same frame plus the D2H of the final image into pinned host memory on rank 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame(spp, to_host=True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms, e2e_ms, float(st["rays_closest"] + st["rays_shadow"]), float(st["kernel_launches"]), float(st["rays_closest"]), float(st["shaded"])], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms, e2e_ms = float(tmax[0]), float(tmax[1])
        rays_per_frame = float(t[2])
        img = h_img.numpy()
        line = {"metric": "1080p spp/s", "value": spp * args.steps / (ms * 1e-3), "unit": "spp/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": pt_config(world, spp, args.split, args.trace_mode), "clocks": clocks,
                "e2e": {"value": spp * args.steps / (e2e_ms * 1e-3), "unit": "spp/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": PT_W * PT_H * 3 * 4},
                "gpu_launches": int(float(t[3]) * args.steps), "mrays_s": rays_per_frame / (ms / args.steps * 1e-3) / 1e6,
                "msamples_s": PT_W * PT_H * spp / (ms / args.steps * 1e-3) / 1e6, "rays_per_sample": rays_per_frame / (PT_W * PT_H * spp),
                "closest_rays_per_sample": float(t[4]) / (PT_W * PT_H * spp), "shaded_hits_per_sample": float(t[5]) / (PT_W * PT_H * spp),
                "image_mean": float(img.mean()), "image_finite": bool(np.isfinite(img).all()), "roofline": None, "cpu_baseline": None}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_pathtrace_baseline(desc)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_pathtrace_baseline(desc):
    """The C oracle (a port: the reference needs its own Scene object, which cannot hold this synthetic scene) on a
    bounded sample of the frame: 960x540 at 64 spp (about 10 s), all host threads, scaled linearly (the loop is linear in pixels x spp)."""
    from oracle import oracle as orc
    from rendu_b200 import workloads as wl
    threads = max((os.cpu_count() or 2) - 1, 1)
    sc = orc.OracleScene(desc)
    w, h, s = 960, 540, 64
    t0 = time.perf_counter()
    _, rays = sc.render(wl.cornell_camera(PT_W / PT_H), w, h, s, PT_DEPTH, seed=1, rng="counter", threads=threads)
    dt = time.perf_counter() - t0
    msamples = w * h * s / dt / 1e6
    return {"value": msamples * 1e6 / (PT_W * PT_H), "unit": "spp/s", "cores": threads, "kind": "port", "msamples_s": msamples, "mrays_s": rays / dt / 1e6,
            "sample": "%dx%d at %d spp, depth %d (%.1f s), scaled to 1920x1080" % (w, h, s, PT_DEPTH, dt)}


# ------------------------------------------------------------------------------------------ LBVH workload (C5)
def run_lbvh(args):
    """BASELINE configs[4]: 10M-triangle displaced grid (N=2236 -> 9,999,392 triangles): GPU LBVH build, refit after
    re-displacing the vertices, then the C2-style coherent ray set; the reference's CPU updateHierarchy beside it."""
    import torch
    import rendu_b200 as rb
    from rendu_b200 import workloads as wl
    n_grid = 2236 if not args.small else 500
    pos, idx = wl.displaced_grid(n_grid)
    dev = torch.device("cuda", 0)
    out = {"triangles": int(idx.size // 3)}
    rc = rb.Raycaster()
    rc.addMesh(pos, idx)
    rc.updateHierarchy(rb.BUILD_LBVH)  # warm-up build (module load, allocator)
    times, dev_ms = [], []
    for _ in range(3):
        rc2 = rb.Raycaster()
        rc2.addMesh(pos, idx)
        t0 = time.perf_counter()
        rc2.updateHierarchy(rb.BUILD_LBVH)
        times.append(time.perf_counter() - t0)
        dev_ms.append(rc2.build_times()["build_device_ms"])
        del rc2
    out["lbvh_build_ms_incl_upload"] = 1e3 * min(times)
    moved = pos.copy()
    moved[:, 1] = (moved[:, 1] * np.float32(1.7)).astype(np.float32)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        rc.refit(moved)
        times.append(time.perf_counter() - t0)
    out["refit_ms_incl_upload"] = 1e3 * min(times)
    out.update(rc.build_times())
    out["build_device_ms"] = min(dev_ms)  # the first build of the process (rc) also pays module load and first-touch allocation
    out["depth"] = rc.sizes()["depth"]
    co, cd = wl.coherent_rays(RES)
    n = co.shape[0]
    d_o, d_d = torch.from_numpy(co).to(dev), torch.from_numpy(cd).to(dev)
    d_h = torch.empty(n * 48, dtype=torch.uint8, device=dev)
    sh = torch.cuda.current_stream().cuda_stream
    for mode, name in ((rb.TRACE_EXACT, "exact"), (rb.TRACE_ORDERED, "ordered")):
        rc.set_trace_mode(mode)
        for _ in range(3):
            rc.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), MINI, MAXI, sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            rc.intersects_device(d_o.data_ptr(), d_d.data_ptr(), n, d_h.data_ptr(), MINI, MAXI, sh)
        e1.record()
        torch.cuda.synchronize()
        out["closest_coherent_mrays_s_" + name] = n / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e6
    hits = d_h.cpu().numpy().view(rb.HIT_DTYPE)
    out["hit_rate"] = float((hits["hit"] == 1).mean())
    # ... and a path trace whose floor is that mesh: 1080p, 16 spp, depth 8, scene committed with the
    # GPU LBVH builder; scene_commit_ms covers flattening, upload, LBVH build and attribute tables.
    del rc, d_o, d_d, d_h
    from rendu_b200 import pathtracer as rpt
    desc = wl.cornell_synthetic(n_grid, 512, 256)
    t0 = time.perf_counter()
    sc = rpt.Scene(desc, build_mode=rb.BUILD_LBVH)
    out["scene_commit_ms"] = 1e3 * (time.perf_counter() - t0)
    pt = rpt.PathTracer(sc)
    cam = wl.cornell_camera(PT_W / PT_H)
    d_sum = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32, device=dev)
    for mode, name in ((0, "exact"), (1, "ordered")):
        o = rpt.make_opts(seed=1, trace_mode=mode)
        pt.render_device(cam, PT_W, PT_H, 16, PT_DEPTH, d_sum.data_ptr(), o, sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            pt.render_device(cam, PT_W, PT_H, 16, PT_DEPTH, d_sum.data_ptr(), o, sh)
        e1.record()
        torch.cuda.synchronize()
        out["pathtrace_1080p_spp_s_" + name] = 2 * 16 / (e0.elapsed_time(e1) * 1e-3)
    out["pathtrace_image_finite"] = bool(torch.isfinite(d_sum).all())
    if not args.no_cpu:
        from oracle import ref
        if ref.available():
            r = ref.RefRaycaster()
            r.add_mesh(pos, idx)
            out["reference_updateHierarchy_s"] = r.build()
            out["reference_nodes"] = r.sizes()["nodes"]
    emit({"workload": "C5 LBVH build + refit + trace", **out})


def run_c1(args):
    """BASELINE configs[0]: the reference's own CPU-runnable case — bundled scenes at 256x256, 8 spp, depth 5, scene
    camera — rendered through rb_pathtracer_render (host image in, host image out) next to the reference's own
    PathTracer::render on the box's host cores. The scenes come from the committed fixtures (tests/golden/scene_*.npz,
    flattened from the reference's loader by tests/golden/make_golden.py); the reference loads its own .scene files."""
    import rendu_b200 as rb  # noqa: F401
    from rendu_b200 import pathtracer as rpt
    out = {"workload": "C1 bundled scenes 256x256, 8 spp, depth 5", "scenes": {}}
    for name in ("cornellbox", "cornellbox-emissive"):
        desc = rpt.load_scene(os.path.join(ROOT, "tests", "golden", "scene_%s.npz" % name))
        g = np.load(os.path.join(ROOT, "tests", "golden", "render_%s.npz" % name))
        cam = dict(eye=g["cam_eye"], corner=g["cam_corner"], dx=g["cam_dx"], dy=g["cam_dy"])
        t0 = time.perf_counter()
        pt = rpt.PathTracer(rpt.Scene(desc))
        setup_s = time.perf_counter() - t0
        img = np.zeros((256, 256, 3), np.float32)
        times = []
        for k in range(13):
            img[:] = 0
            t0 = time.perf_counter()
            pt.render(cam, 8, 5, img, rpt.make_opts(seed=1))
            times.append(time.perf_counter() - t0)
        st = pt.stats()
        rec = {"gpu_render_ms": 1e3 * float(np.median(times[3:])), "gpu_scene_setup_ms": 1e3 * setup_s, "rays": st["rays_closest"] + st["rays_shadow"],
               "kernel_launches": st["kernel_launches"], "image_mean": float(img.mean())}
        if not args.no_cpu:
            from oracle import ref
            if ref.available():
                ref.add_default_resources()
                rs = ref.RefScene(name)
                ref.seed(1)
                _, dt, rays = rs.render(256, 256, 8, 5)
                _, dt2, _ = rs.render(256, 256, 8, 5)
                rec.update(reference_render_ms=1e3 * min(dt, dt2), reference_rays=rays, reference_threads=ref.threads_default())
                rec["speedup"] = rec["reference_render_ms"] / rec["gpu_render_ms"]
        out["scenes"][name] = rec
    emit(out)


def run_explore(args):
    """Developer aid (not a bench line): times scheduler / mode variants of the traversal kernels on the C2 sets and
    checks on the device that every variant returns the same hit records as the first one."""
    import torch
    import rendu_b200 as rb
    from rendu_b200 import workloads as wl
    dev = torch.device("cuda", 0)
    pos, idx = wl.displaced_grid(GRID_N)
    co, cd = wl.coherent_rays(RES)
    n = co.shape[0]
    sh = torch.cuda.current_stream().cuda_stream
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_co, d_cd = to(co), to(cd)
    variants = [("simple", {"RB_PERSISTENT": "0", "RB_LEAF_VOTES": "0", "RB_DUAL": "0"}, rb.TRACE_EXACT),
                ("persistent_t20", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "20", "RB_LEAF_VOTES": "0", "RB_DUAL": "0"}, rb.TRACE_EXACT)]
    for votes, thr, steps in ((8, 20, 1), (8, 20, 2), (8, 20, 4), (6, 12, 6), (6, 16, 6), (4, 12, 4), (6, 8, 6), (6, 12, 12)):
        variants.append(("phased_v%d_t%d_s%d" % (votes, thr, steps), {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": str(thr), "RB_LEAF_VOTES": str(votes), "RB_NODE_STEPS": str(steps), "RB_DUAL": "0"}, rb.TRACE_EXACT))
    for votes, thr, steps in ((6, 20, 6), (6, 24, 6), (6, 28, 6), (6, 32, 6), (4, 24, 4), (8, 24, 6), (6, 24, 4)):
        variants.append(("phased_v%d_t%d_s%d" % (votes, thr, steps), {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": str(thr), "RB_LEAF_VOTES": str(votes), "RB_NODE_STEPS": str(steps), "RB_DUAL": "0"}, rb.TRACE_EXACT))
    for thr in (12, 20, 24, 28, 32):  # prepared-ray ring (experiment): the refill costs no latency, so the threshold can go up
        variants.append(("ring_v6_t%d_s6" % thr, {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": str(thr), "RB_LEAF_VOTES": "6", "RB_NODE_STEPS": "6", "RB_DUAL": "0", "RB_RING": "1"}, rb.TRACE_EXACT))
    variants.append(("ring_v4_t28_s3", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "28", "RB_LEAF_VOTES": "4", "RB_NODE_STEPS": "3", "RB_DUAL": "0", "RB_RING": "1"}, rb.TRACE_EXACT))
    for thr in (12, 20, 28):  # eight-wide quantised tree (rb_wide.h): plain persistent scheduler, refill threshold as named
        variants.append(("wide_t%d" % thr, {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": str(thr), "RB_DUAL": "0"}, rb.TRACE_WIDE))
    variants.append(("wide_octant_t12", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "12", "RB_DUAL": "0", "RB_WIDE_OCTANT": "1"}, rb.TRACE_WIDE))  # closest hit front to back
    for votes, thr in ((6, 12), (12, 12), (6, 20), (16, 20)):  # the same tree under the phase-voting scheduler (node entries against triangle tests)
        variants.append(("wide_phased_v%d_t%d" % (votes, thr), {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": str(thr), "RB_LEAF_VOTES": str(votes), "RB_DUAL": "0", "RB_WIDE_PHASED": "1"}, rb.TRACE_WIDE))
    variants.append(("dual_v24_r16", {"RB_PERSISTENT": "1", "RB_DUAL": "1", "RB_LEAF_VOTES": "24", "RB_DUAL_REFILL": "16"}, rb.TRACE_EXACT))
    variants.append(("ordered_phased_v8_t20_s2", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "20", "RB_LEAF_VOTES": "8", "RB_NODE_STEPS": "2", "RB_DUAL": "0"}, rb.TRACE_ORDERED))
    variants.append(("ordered_phased_v6_t12_s6", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "12", "RB_LEAF_VOTES": "6", "RB_NODE_STEPS": "6", "RB_DUAL": "0"}, rb.TRACE_ORDERED))
    variants.append(("ordered_phased_v6_t12_s3", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "12", "RB_LEAF_VOTES": "6", "RB_NODE_STEPS": "3", "RB_DUAL": "0"}, rb.TRACE_ORDERED))
    variants.append(("ordered_persistent_t20", {"RB_PERSISTENT": "1", "RB_FETCH_THRESHOLD": "20", "RB_LEAF_VOTES": "0", "RB_DUAL": "0"}, rb.TRACE_ORDERED))
    only = os.environ.get("RB_EXPLORE_ONLY")  # comma-separated name prefixes: time a subset (the first one is the reference)
    if only:
        variants = [v for v in variants if any(v[0].startswith(p) for p in only.split(","))]
    base, d_io, d_id, results = None, None, None, {}
    for name, env, mode in variants:
        os.environ.update(env)
        os.environ["RB_RING"] = env.get("RB_RING", "0")
        os.environ["RB_WIDE_PHASED"] = env.get("RB_WIDE_PHASED", "0")
        os.environ["RB_WIDE_OCTANT"] = env.get("RB_WIDE_OCTANT", "0")
        rc = rb.Raycaster()
        rc.addMesh(pos, idx)
        rc.updateHierarchy()
        if mode == rb.TRACE_WIDE:
            rc.build_wide()
        rc.set_trace_mode(mode)
        bufs = {k: torch.empty(n * 48, dtype=torch.uint8, device=dev) for k in ("cc", "ci")}
        anys = {k: torch.empty(n, dtype=torch.uint8, device=dev) for k in ("ac", "ai")}
        if d_io is None:
            rc.intersects_device(d_co.data_ptr(), d_cd.data_ptr(), n, bufs["cc"].data_ptr(), MINI, MAXI, sh)
            torch.cuda.synchronize()
            hits = bufs["cc"].cpu().numpy().view(rb.HIT_DTYPE)
            nrm = wl.triangle_normals(pos, idx, np.where(hits["hit"] == 1, hits["localId"], 0))
            io, idr = wl.incoherent_rays(co, cd, hits["hit"], hits["dist"], nrm, seed=1234)
            d_io, d_id = to(io), to(idr)
        runs = [("cc", lambda: rc.intersects_device(d_co.data_ptr(), d_cd.data_ptr(), n, bufs["cc"].data_ptr(), MINI, MAXI, sh)),
                ("ac", lambda: rc.intersectsAny_device(d_co.data_ptr(), d_cd.data_ptr(), n, anys["ac"].data_ptr(), MINI, MAXI, sh)),
                ("ci", lambda: rc.intersects_device(d_io.data_ptr(), d_id.data_ptr(), n, bufs["ci"].data_ptr(), MINI, MAXI, sh)),
                ("ai", lambda: rc.intersectsAny_device(d_io.data_ptr(), d_id.data_ptr(), n, anys["ai"].data_ptr(), MINI, MAXI, sh))]
        res = {}
        for _ in range(6):  # bring the clocks up before timing anything of this variant
            for _, f in runs:
                f()
        for key, f in runs:
            f()
            torch.cuda.synchronize()
            times = []
            for _ in range(9):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                f()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            res[key + "_ms"] = float(np.median(times))
        res["total_ms"] = sum(res[k] for k in ("cc_ms", "ac_ms", "ci_ms", "ai_ms"))
        res["mrays_s"] = 4 * n / res["total_ms"] / 1e3
        snap = {k: v.clone() for k, v in list(bufs.items()) + list(anys.items())}
        if base is None:
            base = snap
        res["mismatch_vs_first"] = {k: int((snap[k].view(-1, 48 if k in bufs else 1) != base[k].view(-1, 48 if k in bufs else 1)).any(dim=1).sum()) for k in snap}
        results[name] = res
        log("[explore] %-24s %s" % (name, json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()})))
        del rc
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "explore.json"), "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--explore", action="store_true", help="developer aid: time kernel variants, no bench line")
    ap.add_argument("--workload", default="raycast", choices=["raycast", "pathtrace", "lbvh", "c1"], help="raycast = BASELINE configs[1] (default); pathtrace = configs[2]/[3]")
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--max-paths", type=int, default=0, help="pathtrace: paths in flight per wavefront batch (0 = library default)")
    ap.add_argument("--split", default="tile", choices=["tile", "sample"])
    ap.add_argument("--small", action="store_true", help="pathtrace: 8k-triangle floor instead of 1M (quick checks)")
    ap.add_argument("--trace-mode", default="exact", choices=["exact", "ordered", "wide"], help="pathtrace: traversal order of the wavefront kernels")
    args = ap.parse_args()
    _claim_stdout()
    if args.explore:
        run_explore(args)
    elif args.workload == "pathtrace" and args.impl != "reference":
        run_pathtrace(args)
    elif args.workload == "lbvh" and args.impl != "reference":
        run_lbvh(args)
    elif args.workload == "c1" and args.impl != "reference":
        run_c1(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
