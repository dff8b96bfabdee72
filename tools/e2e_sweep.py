"""Times the host-buffer C-ABI step of bench.py for several pipeline chunk sizes (DRT_E2E_CHUNK)."""
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
for chunk in sys.argv[1:]:
    os.environ['DRT_E2E_CHUNK'] = chunk
    step(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): step()
    dt = (time.perf_counter() - t) / 5
    print(chunk, f'{dt*1e3:.2f} ms  {3*8388608/dt/1e6:.0f} Mrays/s')
