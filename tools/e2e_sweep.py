"""Times the host-buffer C-ABI step of bench.py for several pipeline chunk sizes (DRT_E2E_CHUNK) and prints the box's
pinned-memory PCIe rates for comparison (the step moves 805 MB up and 277 MB down)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from dartray_b200 import capi, scenes
P, idx = scenes.soup(512)
coh = scenes.coherent_rays(4096, 2048); inc = scenes.incoherent_rays(8388608)
c = capi.Context(0); c.set_triangles(P, idx); c.build_bvh()
h = [torch.from_numpy(a).pin_memory() for a in (*coh, *inc)]
hits = torch.empty((8388608, 4), dtype=torch.float32).pin_memory(); occ = torch.empty(8388608, dtype=torch.uint8).pin_memory()
hn = hits.numpy().view(capi.HIT_DTYPE).reshape(-1)
def step():
    c.trace_closest(h[0].numpy(), h[1].numpy(), out=hn); c.trace_closest(h[2].numpy(), h[3].numpy(), out=hn); c.trace_any(h[2].numpy(), h[3].numpy(), out=occ.numpy())
def pcie():
    n = 256 << 20
    hb, db = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8, device='cuda')
    hb2, db2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(4):
            if h2d:
                with torch.cuda.stream(s1): db.copy_(hb, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): hb2.copy_(db2, non_blocking=True)
        torch.cuda.synchronize(); return 4 * n / (time.perf_counter() - t) / 1e9
    run(True, True)
    print(f'PCIe pinned: H2D {run(True, False):.1f} GB/s, D2H {run(False, True):.1f} GB/s, both at once {run(True, True):.1f} GB/s each')
pcie()
for arg in sys.argv[1:]:
    os.environ['DRT_E2E_CHUNK'] = arg
    step(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): step()
    dt = (time.perf_counter() - t) / 5
    print(arg, f'{dt*1e3:.2f} ms  {3*8388608/dt/1e6:.0f} Mrays/s')
