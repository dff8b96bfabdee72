import sys, time, json
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from dartray_b200 import capi, scenes
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 23
P, idx = scenes.soup(ns)
c = capi.Context(0); c.set_triangles(P, idx); t=time.time(); c.build_bvh(); print('build', time.time()-t, c.bvh_info())
sets = {'incoherent': scenes.incoherent_rays(nr), 'coherent': scenes.coherent_rays(4096, nr // 4096)}
for name, (ro, rd) in sets.items():
    n = ro.shape[0]
    dro, drd = torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda()
    dh = torch.empty((n, 4), dtype=torch.float32, device='cuda'); docc = torch.empty(n, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for kind in ('closest', 'any'):
        f = (lambda: c.trace_closest_device(dro.data_ptr(), drd.data_ptr(), n, dh.data_ptr(), st)) if kind == 'closest' else (lambda: c.trace_any_device(dro.data_ptr(), drd.data_ptr(), n, docc.data_ptr(), st))
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{name} {kind}: {ms:.3f} ms  {n/ms/1e3:.1f} Mrays/s')
    c.set_counting(True); c.trace_closest_device(dro.data_ptr(), drd.data_ptr(), n, dh.data_ptr(), st); print(name, c.counters()); c.set_counting(False)
    t=time.time(); h = c.trace_closest(ro, rd); print('host-buffer e2e closest', time.time()-t, 'kernel ms', c.last_kernel_ms)
