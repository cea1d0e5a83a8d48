#!/usr/bin/env python
"""bench.py — Rendu's offline path tracer on B200: BASELINE.json's metric, Mrays/s (primary + bounce + shadow) and 1080p spp/s.

Headline workload (one "step" = one frame): 1920x1080, depth 8, the synthetic textured GGX+Lambert
Cornell room (cubemap environment, emissive panel, shadow-casting point light, NEE on) whose floor is the 1,002,528-
triangle displaced grid — 256 spp on one B200 (BASELINE configs[2]), 4096 spp strong-scaled over N GPUs (configs[3]: one
rank per GPU under torchrun, tile split by default, bands gathered on rank 0; the sample split with its NCCL reduce is
timed beside it). The same rate metric at every N, so the driver's v_N / (N v_1) is the strong-scaling efficiency.

  value        : device-timed Mrays/s of PathTracer::render's kernels (CUDA events on the launching stream), all ranks' rays.
  e2e          : the same frame through the public API — rb_pathtracer_render with the caller's host Image (N = 1), or the
                 rank-0 host image after the gather (N > 1): camera in, pixels out, copies inside the timed region.
  roofline     : per kernel class from one profiled frame (rb_pathtracer_profile_frame: CUDA events, counting kernels).
  cpu_baseline : the UNMODIFIED reference PathTracer::render (oracle/_ref) on the same scene, bounded sample, host threads stated.
  raycast      : the C2 Raycaster micro-bench (configs[1]: 4M coherent + 4M incoherent rays, closest-hit and any-hit against
                 the 1M-triangle mesh) with its own device / e2e numbers, roofline, reference baseline and `parity` block.

`--impl reference` times the reference's own CPU path tracer on the same scene (bounded sample per step).
Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import ctypes

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C_F3 = ctypes.c_float * 3
GRID_N = 708          # 2*708^2 = 1,002,528 triangles
RES = 2048            # 2048^2 = 4,194,304 rays per set
MINI, MAXI = 1e-4, 1e8
SAMPLE_ROW_STRIDE = 16  # CPU legs trace every 16th row of each set: 262,144 rays per set
PT_W, PT_H, PT_DEPTH = 1920, 1080, 8


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


REF_W, REF_H, REF_SPP = 480, 270, 16  # the CPU legs render this share of the frame (same camera, same aspect, same depth)


def reference_pathtracer(desc):
    """The unmodified reference PathTracer on the benchmark scene (oracle/_ref through oracle/ref.py RefBuiltScene), else the
    C restatement. Returns (kind, threads, step) with step() -> (seconds, rays, image)."""
    from rendu_b200 import workloads as wl
    from oracle import ref
    cam = wl.CORNELL_CAMERA
    if ref.available():
        ref.seed(1)
        rs = ref.RefBuiltScene(desc)
        threads = ref.threads_default()  # the reference's own policy: hardware_concurrency() - 1 workers (System.hpp:82)

        def step():
            img, dt, rays = rs.render(cam["eye"], cam["center"], cam["up"], cam["fov"], REF_W, REF_H, REF_SPP, PT_DEPTH)
            return dt, rays, img
        return "reference", threads, step
    from oracle import oracle as orc
    sc = orc.OracleScene(desc)
    threads = max((os.cpu_count() or 2) - 1, 1)

    def step():
        t0 = time.perf_counter()
        img, rays = sc.render(wl.cornell_camera(REF_W / REF_H), REF_W, REF_H, REF_SPP, PT_DEPTH, seed=1, rng="counter", threads=threads)
        return time.perf_counter() - t0, rays, img
    return "port", threads, step


def reference_sample_text():
    return "%dx%d at %d spp, depth %d of the same scene and camera (%.2f %% of the frame's samples per step)" % (
        REF_W, REF_H, REF_SPP, PT_DEPTH, 100.0 * REF_W * REF_H * REF_SPP / (PT_W * PT_H * 256))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rendu_b200 import workloads as wl
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    spp = args.spp or (256 if world == 1 else 4096)
    desc = wl.cornell_synthetic(GRID_N if not args.small else 64, 512, 256)
    kind, threads, step = reference_pathtracer(desc)
    for _ in range(args.warmup):
        step()
    total_t, total_rays = 0.0, 0
    for _ in range(args.steps):
        t, rays, _ = step()
        total_t += t
        total_rays += rays
    value = total_rays / total_t / 1e6
    spp_s = REF_W * REF_H * REF_SPP * args.steps / total_t / (PT_W * PT_H)
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": pt_config(world, spp, args.split), "spp_per_s": spp_s,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "host_cpus": os.cpu_count(), "sample": reference_sample_text(), "spp_per_s": spp_s},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm: shared plumbing
def init_ranks():
    import torch
    import rendu_b200 as rb
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
    return dict(rank=rank, world=world, local_rank=local_rank, dev=torch.device("cuda", local_rank), dist=dist)


def ncu_counters(name):
    """Counters of this round's ncu --set full capture of a kernel (profiles/r02_counters.json, written by
    tools/ncu_counters.py from the .ncu-rep of the same bench command): what names the limiter. None when absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_counters.json"))).get(name)
    except Exception:
        return None


def link_probe(torch, dev, nbytes=256 << 20, reps=5):
    """Host link, measured beside the e2e numbers it explains: each direction alone and both at once, warm streams and
    buffers, best of `reps` (CUDA events on the copy streams)."""
    h_up, h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_up, d_dn = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(up, dn):
        best = 1e9
        for _ in range(reps + 1):  # the first repetition warms streams and page mappings
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return nbytes / best / 1e9
    return {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "duplex_gbs_each_way": run(True, True), "bytes": nbytes, "reps": reps}


# ------------------------------------------------------------------------------------------ C2: Raycaster micro-bench
def raycast_block(args, R, steps, with_cpu):
    """BASELINE configs[1] on this rank's GPU (every rank traces the full sets against its own replica; no collective).
    Returns, on rank 0, the block with device / e2e throughput, per-launch times, roofline, parity and the CPU baseline."""
    import torch
    import rendu_b200 as rb
    from rendu_b200 import workloads as wl
    from rendu_b200._capi import check, lib
    rank, world, dev, dist = R["rank"], R["world"], R["dev"], R["dist"]

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
    log("[bench] rank %d raycast setup %.1fs: %d tris, %d nodes, coherent hit rate %.3f" % (rank, time.time() - t_setup, rc.sizes()["tris"], rc.sizes()["nodes"], float((hits["hit"] == 1).mean())))

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

    # Warm-up: at least the requested steps, and at least 0.3 s of work — a step is 5 ms, and an idle B200 needs longer than
    # three of them to bring its clocks up.
    t_w, k_w = time.perf_counter(), 0
    while k_w < max(args.warmup, 3) or time.perf_counter() - t_w < 0.3:
        for _, f in launches:
            f()
        torch.cuda.synchronize()
        k_w += 1
    barrier()

    # ---- device-resident: rays and hit buffers in HBM, CUDA events on the launching stream
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in launches] for _ in range(steps)]
    l0 = rc.launch_count()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record(stream)
    for s_ in range(steps):
        for k, (_, f) in enumerate(launches):
            ev[s_][k][0].record(stream)
            f()
            ev[s_][k][1].record(stream)
    e_stop.record(stream)
    barrier()
    total_ms = e_start.elapsed_time(e_stop)
    gpu_launches = rc.launch_count() - l0
    per_kernel_ms = [float(np.mean([ev[s_][k][0].elapsed_time(ev[s_][k][1]) for s_ in range(steps)])) for k in range(len(launches))]

    # ---- end to end through the host-pointer ABI with pinned buffers
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

    link = link_probe(torch, dev) if rank == 0 else None
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = 4 * n * 24
    d2h = 2 * n * 48 + 2 * n

    # The same step with the opt-in wire format for the coherent set: its rays ARE the primary rays of a pinhole camera
    # (eye (0,1.5,2.2), image plane y = 0 over [-1,1]^2), so rb_raycaster_intersect_primary generates them on the device
    # from 48 bytes of camera — bit for bit the rays of wl.coherent_rays, hence the same Hit records (checked below).
    from rendu_b200 import _capi
    cam = _capi.Camera()
    cam.eye, cam.corner = (C_F3)(0.0, 1.5, 2.2), (C_F3)(-1.0, 0.0, -1.0)
    cam.dx, cam.dy = (C_F3)(2.0, 0.0, 0.0), (C_F3)(0.0, 0.0, 2.0)
    h_hits2 = torch.empty(n * 48, dtype=torch.uint8).pin_memory()

    def e2e_closest_primary():
        check(L.rb_raycaster_intersect_primary(rc._h, ctypes.byref(cam), RES, RES, MINI, MAXI, h_hits2.data_ptr()))
        check(L.rb_raycaster_intersect(rc._h, h_io.data_ptr(), h_id.data_ptr(), MINI, MAXI, n, h_hits.data_ptr()))

    def e2e_any_primary():
        check(L.rb_raycaster_intersect_any_primary(rc._h, ctypes.byref(cam), RES, RES, MINI, MAXI, h_any.data_ptr()))
        check(L.rb_raycaster_intersect_any(rc._h, h_io.data_ptr(), h_id.data_ptr(), MINI, MAXI, n, h_any.data_ptr()))

    def e2e_step_primary():
        fa, fb = pool.submit(e2e_closest_primary), pool.submit(e2e_any_primary)
        fa.result()
        fb.result()

    # ... and with the asynchronous forms from ONE host thread: the four calls queue up in call order on the thread's ring —
    # closest-hit (download-heavy) first, any-hit (upload only) last — and overlap into max(up, down) on the link.
    h_hits_i = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
    h_any_i = torch.empty(n, dtype=torch.uint8).pin_memory()

    def e2e_step_async():
        check(L.rb_raycaster_intersect_async(rc._h, h_co.data_ptr(), h_cd.data_ptr(), MINI, MAXI, n, h_hits2.data_ptr()))
        check(L.rb_raycaster_intersect_async(rc._h, h_io.data_ptr(), h_id.data_ptr(), MINI, MAXI, n, h_hits_i.data_ptr()))
        check(L.rb_raycaster_intersect_any_async(rc._h, h_co.data_ptr(), h_cd.data_ptr(), MINI, MAXI, n, h_any.data_ptr()))
        check(L.rb_raycaster_intersect_any_async(rc._h, h_io.data_ptr(), h_id.data_ptr(), MINI, MAXI, n, h_any_i.data_ptr()))
        check(L.rb_raycaster_sync(rc._h))

    e2e_step_async()
    async_equal = bool(torch.equal(h_hits2, d_hit_c.cpu()) and torch.equal(h_hits_i, d_hit_i.cpu()) and torch.equal(h_any, d_any_c.cpu()) and torch.equal(h_any_i, d_any_i.cpu()))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step_async()
    barrier()
    e2e_async_s = time.perf_counter() - t0

    e2e_step_primary()
    primary_equal = bool(torch.equal(h_hits2, d_hit_c.cpu()))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step_primary()
    barrier()
    e2e_primary_s = time.perf_counter() - t0

    times = torch.tensor([total_ms, e2e_s * 1e3, e2e_primary_s * 1e3, e2e_async_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_primary_ms, e2e_async_ms = float(times[0]), float(times[1]), float(times[2]), float(times[3])
    queries_step = 4 * n
    value = world * queries_step * steps / (total_ms * 1e-3) / 1e6
    e2e_value = world * queries_step * steps / (e2e_ms * 1e-3) / 1e6
    if rank != 0:
        return None

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
    cnt = ncu_counters("raycast_" + work[dom]["name"])
    roofline = {"bound": "hbm", "kernel": "rb_k_trace_phased<%s>" % ("closest" if "closest" in work[dom]["name"] else "any"), "launch": work[dom]["name"],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": (cnt or {}).get("dram_bytes"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": work[dom]["bytes"], "avg_launch_ms": per_kernel_ms[dom],
                "per_ray": {"interior_nodes": work[dom]["interior"] / n, "triangle_tests": work[dom]["tris"] / n, "bytes": work[dom]["bytes"] / n},
                "limiter": "not HBM: the 82 MB tree is L2-resident, so the algorithmic bytes are served by L2 / L1. What ncu names is the L1 data pipe "
                           "(LSU wavefronts, one per cycle and SM at best), then issue slots — see `counters`",
                "counters": cnt, "counters_source": "profiles/r02_counters.json (ncu --set full --clock-control none of this bench command, this round)" if cnt else None}
    per_launch = {w["name"]: {"ms": per_kernel_ms[k], "mrays_s": n / (per_kernel_ms[k] * 1e-3) / 1e6,
                              "interior_per_ray": w["interior"] / n, "tri_tests_per_ray": w["tris"] / n,
                              "gbs": w["bytes"] / (per_kernel_ms[k] * 1e-3) / 1e9} for k, w in enumerate(work)}
    cpu, parity = None, None
    if with_cpu:
        # The reference traces the same rays (every 16th row of both sets): timed as the CPU baseline, and its records are
        # the parity check of what the GPU just produced in the timed region.
        kind, threads, closest, anyhit = cpu_tracer(pos, idx)
        sel = sample_index()
        sets = tuple(np.ascontiguousarray(a[sel]) for a in (co, cd, io, idr))
        ref_cc, ref_ci = closest(sets[0], sets[1]), closest(sets[2], sets[3])
        ref_ac, ref_ai = anyhit(sets[0], sets[1]), anyhit(sets[2], sets[3])
        tsel = torch.from_numpy(sel).to(dev)
        parity = {"rays_per_launch": int(sel.size), "mode": "exact (default trace mode)", "checked_against": kind}
        for name, dbuf, refh in (("closest_coherent", d_hit_c, ref_cc), ("closest_incoherent", d_hit_i, ref_ci)):
            g = dbuf.view(-1, 48)[tsel].cpu().numpy().view(rb.HIT_DTYPE).reshape(-1)
            hit_ref = np.asarray(refh["hit"]).astype(np.uint32)
            both = (g["hit"] == 1) & (hit_ref == 1)
            dref = np.asarray(refh["dist"], np.float32)
            rel = np.abs(g["dist"][both].astype(np.float64) - dref[both]) / np.maximum(np.abs(dref[both]), 1e-30)
            parity[name] = {"hit_flag_mismatch": int((g["hit"] != hit_ref).sum()),
                            "id_mismatch": int((g["localId"][both] != np.asarray(refh["local_id"])[both]).sum()),
                            "dist_bits_mismatch": int((g["dist"][both].view(np.uint32) != dref[both].view(np.uint32)).sum()),
                            "max_rel_dist": float(rel.max()) if rel.size else 0.0}
        for name, dbuf, refa in (("any_coherent", d_any_c, ref_ac), ("any_incoherent", d_any_i, ref_ai)):
            g = dbuf[tsel].cpu().numpy()
            parity[name] = {"flag_mismatch": int((g.astype(np.uint8) != np.asarray(refa).astype(np.uint8)).sum())}
        parity["id_mismatch"] = parity["closest_coherent"]["id_mismatch"] + parity["closest_incoherent"]["id_mismatch"]
        parity["rays"] = 4 * int(sel.size)
        parity["max_rel_dist"] = max(parity["closest_coherent"]["max_rel_dist"], parity["closest_incoherent"]["max_rel_dist"])
        cpu_step(sets, closest, anyhit)
        tt, qq = 0.0, 0
        while tt < 8.0:
            t, q = cpu_step(sets, closest, anyhit)
            tt += t
            qq += q
        cpu = {"value": qq / tt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind, "host_cpus": os.cpu_count(),
               "sample": "every %dth row of both ray sets (%d rays/set), closest+any, repeated for %.1f s" % (SAMPLE_ROW_STRIDE, sets[0].shape[0], tt)}
    return {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "steps": steps, "ms_per_step": total_ms / steps, "scaling": "weak (every rank traces the full sets; no collective)",
            "config": config_dict(world), "gpu_launches": int(gpu_launches),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
                    "api": "rb_raycaster_intersect / rb_raycaster_intersect_any with pinned host buffers, closest and any calls issued from two host threads", "host_link": link,
                    "asynchronous_calls_one_thread": {"value": world * queries_step * steps / (e2e_async_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                                      "api": "rb_raycaster_intersect_async x2, rb_raycaster_intersect_any_async x2, rb_raycaster_sync: same host rays in, same Raycaster::Hit records out",
                                                      "records_identical_to_resident_run": async_equal},
                    "with_device_generated_primary_rays": {"value": world * queries_step * steps / (e2e_primary_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * n * 24 + 96, "d2h_bytes_per_step": d2h,
                                                           "api": "coherent set through rb_raycaster_intersect_primary / _any_primary (camera in, Raycaster::Hit records out), incoherent set as above",
                                                           "records_identical_to_host_rays": primary_equal}},
            "roofline": roofline, "per_launch": per_launch, "parity": parity, "cpu_baseline": cpu}


def run_raycast(args):
    """Developer workload: the C2 block alone, printed as a line of its own."""
    R = init_ranks()
    sampler = ClockSampler(R["local_rank"])
    if R["rank"] == 0:
        sampler.start()
        sampler.begin()
    block = raycast_block(args, R, args.steps, R["world"] == 1 and not args.no_cpu)
    if R["rank"] == 0:
        sampler.end()
        block.update({"n_gpus": R["world"], "warmup": max(args.warmup, 3), "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "clocks": sampler.stop()})
        emit(block)
    if R["dist"] is not None:
        R["dist"].barrier()
        R["dist"].destroy_process_group()


# ------------------------------------------------------------------------------------------ path-tracing workload (C3 / C4): the headline


def pt_config(world, spp, split, trace_mode="exact"):
    return {"workload": "C3/C4 path trace (BASELINE configs[2]/[3]): 1920x1080, %d spp, depth 8, synthetic textured GGX+Lambert Cornell room (cubemap environment, emissive panel, "
                        "shadowed point light, NEE on) with the 1,002,528-triangle displaced grid as floor; PathTracer::render" % spp,
            "width": PT_W, "height": PT_H, "spp": spp, "depth": PT_DEPTH, "triangles": 2 * GRID_N * GRID_N + 12, "split": split if world > 1 else "none", "trace_mode": trace_mode,
            "parallelism": "replicated scene, %d rank(s)%s" % (world, "" if world == 1 else (", tile split (2-row bands dealt round-robin), bands gathered on rank 0" if split == "tile" else ", sample split, one NCCL reduce of the fp32 sums")),
            "l2": "no explicit flush: every bounce streams the wavefront state of a batch (32 Mi paths x 176 B = 5.9 GB) through HBM, far beyond the 126 MB L2",
            "frames": "one GPU renders the 256-spp frame of BASELINE configs[2], N > 1 the 4096-spp frame of configs[3] (override: --spp): same scene, same rate metric; "
                      "per-frame fixed costs (ray generation, accumulation, batch ramp-up) weigh 16 x less in the longer frame, so v_N / (N v_1) may exceed 1 by a few percent"}


def pt_roofline(pt, cam, opts, rays_scale=1.0):
    """Per kernel class, from ONE profiled frame outside the timed region (rb_pathtracer_profile_frame: a CUDA event before
    every kernel of a single-stream frame, then counting copies of the two traversal kernels): algorithmic bytes (DESIGN.md
    section 5) / summed launch durations against the measured HBM peak. The frame itself overlaps these classes on several
    streams, so `share` is of the serialised frame."""
    spp = 32
    p = pt.profile_frame(cam, PT_W, PT_H, spp, PT_DEPTH, opts)
    peak, peak_src = measured_hbm_peak()
    rc_, rs_, sh_ = p["rays_closest"], p["rays_shadow"], p["shaded_hits"]
    algo = {
        # queue slot + ray in (origin, direction) + hit record out, plus the tree: 64 B per interior visit, 48 B per triangle, 32 B per root
        "trace": rc_ * (4 + 32 + 16) + 64 * p["trace_interior"] + 48 * p["trace_tris"] + 32 * p["trace_roots"],
        # hit, ray, attenuation, colour in; next ray, attenuation, colour, shadow entry (48 B), two queue slots out (204 B of state), the
        # 144-byte per-triangle shading record and three 4-texel RGBA fp32 taps (192 B)
        "shade": sh_ * (204 + 144 + 192) + (rc_ - sh_) * 64,
        # shadow entry in (48 B), colour read-modify-write (32 B), plus the tree
        "shadow": rs_ * 80 + 64 * p["shadow_interior"] + 48 * p["shadow_tris"] + 32 * p["shadow_roots"],
        "raygen": p["samples"] * (32 + 16 + 16 + 4),
        "accumulate": p["samples"] * 16 + PT_W * PT_H * 24,
    }
    total_ms = sum(p["ms"].values())
    kernels = {}
    for k, ms in p["ms"].items():
        gbs = algo[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels[k] = {"ms_per_profiled_frame": ms, "launches": p["launches"][k], "avg_launch_ms": ms / max(p["launches"][k], 1), "share_of_serialised_frame": ms / total_ms,
                      "algorithmic_bytes": int(algo[k]), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    dom = max(p["ms"], key=lambda k: p["ms"][k])
    names = {"trace": "rb_k_pt_trace", "shade": "rb_k_pt_shade", "shadow": "rb_k_pt_shadow", "raygen": "rb_k_pt_raygen", "accumulate": "rb_k_pt_accumulate"}
    cnt = ncu_counters("pathtrace_" + dom)
    per_ray = {"closest_interior_nodes": p["trace_interior"] / max(rc_, 1), "closest_triangle_tests": p["trace_tris"] / max(rc_, 1),
               "shadow_interior_nodes": p["shadow_interior"] / max(rs_, 1), "shadow_triangle_tests": p["shadow_tris"] / max(rs_, 1)}
    return {"bound": "hbm", "kernel": names[dom], "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac_of_hbm_peak"],
            "traffic": (cnt or {}).get("dram_bytes"), "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo[dom] / max(p["launches"][dom], 1)),
            "avg_launch_ms": kernels[dom]["avg_launch_ms"], "profiled_frame": "1920x1080, %d spp, depth %d, single stream" % (spp, PT_DEPTH), "per_ray": per_ray, "kernels": kernels,
            "limiter": "not HBM for the traversal wavefronts (the tree is L2-resident: L1 data pipe and issue slots, as in the raycast block); shading is a chain of dependent "
                       "loads at 42 % occupancy (72 registers, 7 CTAs per SM) — see `counters`",
            "counters": cnt, "counters_source": "profiles/r02_counters.json (ncu --set full --clock-control none, this round)" if cnt else None}


def run_headline(args):
    import torch
    import rendu_b200 as rb  # noqa: F401
    from rendu_b200 import distributed as rd
    from rendu_b200 import pathtracer as rpt
    from rendu_b200 import workloads as wl

    R = init_ranks()
    rank, world, local_rank, dev, dist = R["rank"], R["world"], R["local_rank"], R["dev"], R["dist"]
    spp = args.spp or (256 if world == 1 else 4096)
    steps, warmup = args.steps, max(args.warmup, 3)
    # The C2 block runs first, on its own handle, and is gone before the path tracer allocates its wavefront state.
    block = None if args.no_raycast else raycast_block(args, R, steps if world == 1 else max(3, min(steps, 5)), world == 1 and not args.no_cpu)
    torch.cuda.empty_cache()
    t_setup = time.time()
    desc = wl.cornell_synthetic(GRID_N if not args.small else 64, 512, 256)
    cam = wl.cornell_camera(PT_W / PT_H)
    scene = rpt.Scene(desc)
    pt = rpt.PathTracer(scene)
    log("[bench] rank %d scene committed in %.1fs" % (rank, time.time() - t_setup))
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    d_sum = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32, device=dev)
    d_img = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32, device=dev)
    h_img = torch.zeros((PT_H, PT_W, 3), dtype=torch.float32).pin_memory()
    tmode = {"exact": 0, "ordered": 1, "wide": 2}[args.trace_mode]
    if tmode == 2:
        scene.raycaster().build_wide()
    rows = rd.local_rows(PT_H, rank, world).to(dev) if world > 1 else None
    combine_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def opts_for(split):
        if world > 1 and split == "tile":
            return rpt.make_opts(seed=1, tile=rd.tile_spec(rank, world), trace_mode=tmode, max_paths=args.max_paths)
        if world > 1:
            return rpt.make_opts(seed=1, samples=rd.sample_spec(rank, world), trace_mode=tmode, max_paths=args.max_paths)
        return rpt.make_opts(seed=1, trace_mode=tmode, max_paths=args.max_paths)

    def frame(n_spp, split, to_host=False):
        """One frame: every rank renders its share on the device; N > 1: the data-path collective (gather of bands /
        NCCL reduce of sums), resolve; optionally the final image into pinned host memory on rank 0."""
        o = opts_for(split)
        pt.render_device(cam, PT_W, PT_H, n_spp, PT_DEPTH, d_sum.data_ptr(), o, sh)
        final = None
        if world > 1 and split == "sample":
            combine_ev[0].record(stream)
            total = rd.combine_sample_split(d_sum, rank, world)
            combine_ev[1].record(stream)
            if rank == 0:
                d_img.zero_()
                pt.resolve_device(total.data_ptr(), PT_W, PT_H, n_spp, 3, d_img.data_ptr(), None, sh)
                final = d_img
        else:
            d_img.zero_()
            pt.resolve_device(d_sum.data_ptr(), PT_W, PT_H, n_spp, 3, d_img.data_ptr(), o, sh)
            if world > 1:
                combine_ev[0].record(stream)
                final = rd.combine_tile_split(d_img[rows].contiguous(), PT_H, rank, world)
                combine_ev[1].record(stream)
            else:
                final = d_img
        if to_host and rank == 0:
            h_img.copy_(final, non_blocking=True)
        return final

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(split, n_steps, to_host=False):
        """(device ms by CUDA events on the launching stream, wall ms) of n_steps frames, barrier + synchronize on both sides."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(n_steps):
            frame(spp, split, to_host)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    split = args.split if world > 1 else "none"
    for _ in range(warmup):
        frame(spp, split)
    barrier()
    # ---- timed region 1: device-resident (value)
    sampler.begin()
    ms, _ = timed(split, steps)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    st = pt.stats()
    combine_ms = combine_ev[0].elapsed_time(combine_ev[1]) if world > 1 else 0.0
    # ---- timed region 2: end to end. N = 1: the public API call with the caller's host image (upload of the image the
    # reference accumulates into, camera, render, download). N > 1: the frame above plus the rank-0 download.
    host_image = h_img.numpy()
    if world == 1:
        o1 = opts_for("none")
        host_image[:] = 0
        pt.render(cam, spp, PT_DEPTH, host_image, o1)
        e2e_each = []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            pt.render(cam, spp, PT_DEPTH, host_image, o1)  # accumulates into the caller's image like the reference: values grow, the work does not
            e2e_each.append((time.perf_counter() - t1) * 1e3)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        log("[bench] e2e frames (ms): " + " ".join("%.1f" % x for x in e2e_each) + "; device seconds of the last: %.4f" % pt.stats()["device_seconds"])
        host_image[:] = 0
        pt.render(cam, spp, PT_DEPTH, host_image, o1)
        e2e_h2d, e2e_d2h = PT_W * PT_H * 3 * 4 + 48, PT_W * PT_H * 3 * 4
        e2e_api = "rb_pathtracer_render(camera, 1920x1080 host Image): H2D of the caller's image (the reference accumulates into it), render, resolve, D2H"
    else:
        _, e2e_ms = timed(split, steps, to_host=True)
        e2e_h2d, e2e_d2h = 48, PT_W * PT_H * 3 * 4
        e2e_api = "rb_pathtracer_render_device per rank + %s + resolve, final image copied to pinned host memory on rank 0" % ("gather of the bands" if split == "tile" else "NCCL reduce")
    # ---- the other split, timed beside it (N > 1): fewer steps, same frame
    other = None
    if world > 1:
        osplit = "sample" if split == "tile" else "tile"
        frame(spp, osplit)
        n_o = max(2, min(steps, 4))
        oms, _ = timed(osplit, n_o)
        o_combine = combine_ev[0].elapsed_time(combine_ev[1])
        # the collective alone, every rank ready (what the wire costs, without waiting for the slowest rank)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        rd.combine_sample_split(d_sum, rank, world)
        c1.record(stream)
        barrier()
        other = {"split": osplit, "steps": n_o, "ms_per_step": oms / n_o, "combine_ms_last_frame_rank0": o_combine, "reduce_alone_ms": c0.elapsed_time(c1)}
    t = torch.tensor([ms, e2e_ms, float(st["rays_closest"] + st["rays_shadow"]), float(st["kernel_launches"]), float(st["rays_closest"]), float(st["shaded"]),
                      other["ms_per_step"] if other else 0.0], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms, e2e_ms = float(tmax[0]), float(tmax[1])
        rays_per_frame = float(t[2])
        frame_ms = ms / steps
        mrays = rays_per_frame / (frame_ms * 1e-3) / 1e6
        e2e_frame_ms = e2e_ms / steps
        img = host_image
        finite = np.isfinite(img)
        line = {"metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": frame_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": pt_config(world, spp, args.split, args.trace_mode), "clocks": clocks,
                "e2e": {"value": rays_per_frame / (e2e_frame_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h, "steps": steps,
                        "ms_per_step": e2e_frame_ms, "spp_per_s": spp / (e2e_frame_ms * 1e-3), "api": e2e_api},
                "gpu_launches": int(float(t[3]) * steps),
                "spp_per_s": spp / (frame_ms * 1e-3), "msamples_s": PT_W * PT_H * spp / (frame_ms * 1e-3) / 1e6,
                "rays_per_frame": rays_per_frame, "rays_per_sample": rays_per_frame / (PT_W * PT_H * spp),
                "closest_rays_per_sample": float(t[4]) / (PT_W * PT_H * spp), "shaded_hits_per_sample": float(t[5]) / (PT_W * PT_H * spp),
                "image": {"mean": float(img[finite].mean()), "non_finite_pixels": int((~finite).any(axis=2).sum()),
                          "sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
                          "note": "a non-finite pixel is reference-faithful (0 * inf in an unguarded path, PathTracer.cpp:277,288; DESIGN.md NaN fidelity)"}}
        if world > 1:
            line["multi_gpu"] = {"split": split, "combine_ms_last_frame_rank0": combine_ms,
                                 "combine_note": "CUDA events around the collective on rank 0: includes waiting for the slowest rank",
                                 "other_split": dict(other, ms_per_step=float(tmax[6]), spp_per_s=spp / (float(tmax[6]) * 1e-3), mrays_s=rays_per_frame / (float(tmax[6]) * 1e-3) / 1e6)}
        line["roofline"] = pt_roofline(pt, cam, opts_for("none"))
        line["cpu_baseline"], line["parity"] = None, None
        if world == 1 and not args.no_cpu:
            kind, threads, step = reference_pathtracer(desc)
            step()
            tt, rr, nn, ref_img = 0.0, 0, 0, None
            while tt < 10.0:
                dt, rays, ref_img = step()
                tt, rr, nn = tt + dt, rr + rays, nn + 1
            line["cpu_baseline"] = {"value": rr / tt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind, "host_cpus": os.cpu_count(),
                                    "spp_per_s": REF_W * REF_H * REF_SPP * nn / tt / (PT_W * PT_H),
                                    "sample": reference_sample_text() + ", repeated %d times (%.1f s)" % (nn, tt)}
            # Same share of the frame on the GPU: two independent estimates of one image differ by their noise only
            # (sigma_rel ~ 0.43 / sqrt(spp) per estimate); the means must agree far better.
            g = np.zeros((REF_H, REF_W, 3), np.float32)
            pt.render(wl.cornell_camera(REF_W / REF_H), REF_SPP, PT_DEPTH, g, rpt.make_opts(seed=1))
            ok = np.isfinite(g).all(axis=2) & np.isfinite(ref_img).all(axis=2)
            rel_l2 = float(np.sqrt(((g[ok] - ref_img[ok]) ** 2).sum() / (ref_img[ok] ** 2).sum()))
            line["parity"] = {"pathtrace": {"against": kind, "frame": "%dx%d, %d spp" % (REF_W, REF_H, REF_SPP), "rel_l2": rel_l2,
                                            "expected_noise_rel_l2": 0.43 * float(np.sqrt(2.0 / REF_SPP)), "mean_rel_diff": float(abs(g[ok].mean() - ref_img[ok].mean()) / ref_img[ok].mean())},
                              "raycast": block["parity"] if block else None}
        line["raycast"] = block
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    out["nodes"] = rc.sizes()["nodes"]
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs (profiling runs)")
    ap.add_argument("--explore", action="store_true", help="developer aid: time kernel variants, no bench line")
    ap.add_argument("--workload", default="headline", choices=["headline", "raycast", "lbvh", "c1"],
                    help="headline = path trace C3 (N=1) / C4 (N>1) with the C2 raycast block inside (default); raycast = the C2 block alone; lbvh = C5; c1 = bundled scenes")
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel of the frame (0 = 256 on one GPU, 4096 on several)")
    ap.add_argument("--max-paths", type=int, default=0, help="paths in flight per wavefront batch (0 = library default)")
    ap.add_argument("--split", default="tile", choices=["tile", "sample"], help="N>1: the split `value` is measured with (the other one is timed beside it)")
    ap.add_argument("--no-raycast", action="store_true", help="developer aid: skip the C2 raycast block (kernel sweeps of the path tracer)")
    ap.add_argument("--small", action="store_true", help="8k-triangle floor instead of 1M (quick checks)")
    ap.add_argument("--trace-mode", default="exact", choices=["exact", "ordered", "wide"], help="traversal order of the path tracer's wavefront kernels")
    args = ap.parse_args()
    _claim_stdout()
    if args.explore:
        run_explore(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "raycast":
        run_raycast(args)
    elif args.workload == "lbvh":
        run_lbvh(args)
    elif args.workload == "c1":
        run_c1(args)
    else:
        run_headline(args)


if __name__ == "__main__":
    main()
