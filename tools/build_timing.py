"""Host BVH build of soup(n) on a host-only context, with the builder's phase timings (DRT_BUILD_TIMING=1)."""
import sys, time; import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from dartray_b200 import capi, scenes
ns = int(sys.argv[1]) if len(sys.argv)>1 else 5120
t=time.time(); P, idx = scenes.soup(ns); print('gen', time.time()-t, idx.shape)
c = capi.Context(capi.DEVICE_NONE)
t=time.time(); c.set_triangles(P, idx); print('set', time.time()-t)
t=time.time(); c.build_bvh(); print('build', time.time()-t, c.bvh_info()['build_seconds'])
