"""developer aid (GPU box): where the host-pointer batch API spends its time — each of the four C2 calls alone, then pairs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rendu_b200 as rb
from rendu_b200 import workloads as wl
from rendu_b200._capi import check, lib
from concurrent.futures import ThreadPoolExecutor

pos, idx = wl.displaced_grid(708)
rc = rb.Raycaster(); rc.addMesh(pos, idx); rc.updateHierarchy()
co, cd = wl.coherent_rays(2048); n = co.shape[0]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_co, h_cd = pin(co), pin(cd)
h_hits = torch.empty(n * 48, dtype=torch.uint8).pin_memory(); h_hits2 = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
h_any = torch.empty(n, dtype=torch.uint8).pin_memory(); h_any2 = torch.empty(n, dtype=torch.uint8).pin_memory()
L = lib()
hits = np.zeros(n, rb.HIT_DTYPE)
check(L.rb_raycaster_intersect(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, h_hits.data_ptr()))
hh = h_hits.numpy().view(rb.HIT_DTYPE)
nrm = wl.triangle_normals(pos, idx, np.where(hh["hit"] == 1, hh["localId"], 0))
io, idr = wl.incoherent_rays(co, cd, hh["hit"], hh["dist"], nrm, seed=1234)
h_io, h_id = pin(io), pin(idr)
calls = {
    "closest_coherent": lambda out=h_hits: check(L.rb_raycaster_intersect(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, out.data_ptr())),
    "any_coherent": lambda out=h_any: check(L.rb_raycaster_intersect_any(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, out.data_ptr())),
    "closest_incoherent": lambda out=h_hits2: check(L.rb_raycaster_intersect(rc._h, h_io.data_ptr(), h_id.data_ptr(), 1e-4, 1e8, n, out.data_ptr())),
    "any_incoherent": lambda out=h_any2: check(L.rb_raycaster_intersect_any(rc._h, h_io.data_ptr(), h_id.data_ptr(), 1e-4, 1e8, n, out.data_ptr())),
}
def t(f, reps=8):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
for k, f in calls.items():
    print("%-20s alone: %.2f ms" % (k, t(f)))
pool = ThreadPoolExecutor(4)
def pair(a, b):
    def f():
        x, y = pool.submit(calls[a]), pool.submit(calls[b]); x.result(); y.result()
    return f
for a, b in (("closest_coherent", "any_coherent"), ("closest_coherent", "closest_incoherent"), ("any_coherent", "any_incoherent"), ("closest_incoherent", "any_incoherent")):
    print("%s + %s together: %.2f ms" % (a, b, t(pair(a, b))))
def allfour():
    fs = [pool.submit(f) for f in calls.values()]; [x.result() for x in fs]
print("all four on four threads: %.2f ms" % t(allfour))
def async4():
    check(L.rb_raycaster_intersect_async(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, h_hits.data_ptr()))
    check(L.rb_raycaster_intersect_async(rc._h, h_io.data_ptr(), h_id.data_ptr(), 1e-4, 1e8, n, h_hits2.data_ptr()))
    check(L.rb_raycaster_intersect_any_async(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, h_any.data_ptr()))
    check(L.rb_raycaster_intersect_any_async(rc._h, h_io.data_ptr(), h_id.data_ptr(), 1e-4, 1e8, n, h_any2.data_ptr()))
    check(L.rb_raycaster_sync(rc._h))
print("all four asynchronously from one thread (closest first): %.2f ms" % t(async4))
def async4b():
    check(L.rb_raycaster_intersect_async(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, h_hits.data_ptr()))
    check(L.rb_raycaster_intersect_any_async(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, h_any.data_ptr()))
    check(L.rb_raycaster_intersect_async(rc._h, h_io.data_ptr(), h_id.data_ptr(), 1e-4, 1e8, n, h_hits2.data_ptr()))
    check(L.rb_raycaster_intersect_any_async(rc._h, h_io.data_ptr(), h_id.data_ptr(), 1e-4, 1e8, n, h_any2.data_ptr()))
    check(L.rb_raycaster_sync(rc._h))
print("all four asynchronously, interleaved closest/any: %.2f ms" % t(async4b))
# raw copies of the same sizes for comparison
d = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
def cp_down(): h_hits.copy_(d, non_blocking=True); torch.cuda.synchronize()
def cp_up(): d[: n * 24].copy_(h_hits[: n * 24], non_blocking=True); torch.cuda.synchronize()
print("raw D2H 201 MB: %.2f ms, raw H2D 100 MB: %.2f ms" % (t(cp_down), t(cp_up)))
# the same transfer pattern without any kernel: per chunk two uploads (origins, directions) and one download, round-robin over S streams
chunk = 1 << 19
d_up = [torch.empty(chunk * 24, dtype=torch.uint8, device="cuda") for _ in range(8)]
d_dn = [torch.empty(chunk * 48, dtype=torch.uint8, device="cuda") for _ in range(8)]
h_up = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
for S in (3, 6):
    streams = [torch.cuda.Stream() for _ in range(S)]
    def pattern(down_bytes=48, ups=2):
        for c in range(n // chunk):
            k = c % S
            with torch.cuda.stream(streams[k]):
                for u in range(ups):
                    d_up[k][u * chunk * 12:(u + 1) * chunk * 12].copy_(h_up[c * chunk * 24 + u * chunk * 12: c * chunk * 24 + (u + 1) * chunk * 12], non_blocking=True)
                h_hits[c * chunk * down_bytes:(c + 1) * chunk * down_bytes].copy_(d_dn[k][: chunk * down_bytes], non_blocking=True)
        torch.cuda.synchronize()
    print("copies only, %d streams: one closest call (100 MB up, 201 MB down): %.2f ms" % (S, t(pattern)))
    def two():
        pattern(); 
    def both():
        for c in range(n // chunk):
            for rep, db in ((0, 48), (1, 1)):
                k = (2 * c + rep) % S
                with torch.cuda.stream(streams[k]):
                    d_up[k][: chunk * 24].copy_(h_up[c * chunk * 24:(c + 1) * chunk * 24], non_blocking=True)
                    h_hits[c * chunk * db:(c + 1) * chunk * db].copy_(d_dn[k][: chunk * db], non_blocking=True)
        torch.cuda.synchronize()
    print("copies only, %d streams: closest + any interleaved (201 MB up, 205 MB down): %.2f ms" % (S, t(both)))
