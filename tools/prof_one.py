"""Three launches of the production closest-hit kernel (and any-hit with `any`) on the config-2 incoherent set, for ncu."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from dartray_b200 import capi, scenes
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
kind = sys.argv[2] if len(sys.argv) > 2 else 'closest'
P, idx = scenes.soup(512)
c = capi.Context(0); c.set_triangles(P, idx); c.build_bvh()
ro, rd = scenes.incoherent_rays(nr)
dro, drd = torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda()
dh = torch.empty((nr, 4), dtype=torch.float32, device='cuda')
docc = torch.empty(nr, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    if kind == 'any':
        c.trace_any_device(dro.data_ptr(), drd.data_ptr(), nr, docc.data_ptr(), st)
    else:
        c.trace_closest_device(dro.data_ptr(), drd.data_ptr(), nr, dh.data_ptr(), st)
torch.cuda.synchronize()
