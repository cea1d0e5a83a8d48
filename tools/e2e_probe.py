import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rendu_b200 as rb
from rendu_b200 import workloads as wl
from rendu_b200._capi import check, lib
n = 2048 * 2048
pos, idx = wl.displaced_grid(708)
co, cd = wl.coherent_rays(2048)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_co, h_cd = pin(co), pin(cd)
h_hits = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
d = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
for name, f in (("h2d", lambda: d.copy_(h_hits, non_blocking=True)), ("d2h", lambda: h_hits.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(name, "%.1f GB/s" % (n * 48 / dt / 1e9))
for votes in ("0", "4"):
    os.environ["RB_LEAF_VOTES"] = votes
    rc = rb.Raycaster(); rc.addMesh(pos, idx); rc.updateHierarchy()
    L = lib()
    for rep in range(4):
        t = time.perf_counter()
        check(L.rb_raycaster_intersect(rc._h, h_co.data_ptr(), h_cd.data_ptr(), 1e-4, 1e8, n, h_hits.data_ptr()))
        dt = time.perf_counter() - t
        print("votes", votes, "host intersect %.2f ms -> %.0f Mrays/s" % (dt * 1e3, n / dt / 1e6))
    unp = np.ascontiguousarray(co); und = np.ascontiguousarray(cd); out = np.zeros(n, rb.HIT_DTYPE)
    t = time.perf_counter(); check(L.rb_raycaster_intersect(rc._h, unp.ctypes.data, und.ctypes.data, 1e-4, 1e8, n, out.ctypes.data)); dt = time.perf_counter() - t
    print("pageable host intersect %.2f ms" % (dt * 1e3))
