#!/usr/bin/env python
"""Experiment: how much does ordering the incoherent ray set buy the production traversal kernel?
The rays are permuted on the HOST (numpy) by several candidate keys, then traced with
drt_trace_closest_device / drt_trace_any_device; results are order-independent per ray.
Writes gpurun_out/sort_experiment.json."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def part1by2(v):
    v = v.astype(np.uint64) & 0x1FFFFF
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton3(q, bits):
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


def quant(x, lo, hi, bits):
    s = (1 << bits) - 1
    return np.clip(((x - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, s).astype(np.uint64)


def keys(ro, rd):
    o, d = ro[:, :3].astype(np.float64), rd[:, :3].astype(np.float64)
    out = {}
    out["origin10"] = morton3(quant(o, -2.5, 2.5, 10), 10)
    # entry point into the scene box [-1.2, 1.2]^3 (slab), then direction
    inv = 1.0 / np.where(d == 0, 1e-30, d)
    t0 = (-1.2 - o) * inv
    t1 = (1.2 - o) * inv
    tn = np.minimum(t0, t1).max(axis=1)
    tf = np.maximum(t0, t1).min(axis=1)
    tn = np.where(tf >= np.maximum(tn, 0), np.maximum(tn, 0), 0.0)
    pe = o + d * tn[:, None]
    out["entry10"] = morton3(quant(pe, -1.25, 1.25, 10), 10)
    # interleaved: 3 origin bits / 3 direction bits alternately (6D Morton light): entry point 6 bits + dir 6 bits
    qe, qd = quant(pe, -1.25, 1.25, 7), quant(d, -1.0, 1.0, 7)
    k = np.zeros(o.shape[0], np.uint64)
    for b in range(7):
        for a in range(3):
            k |= ((qe[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(6 * b + a + 3)
            k |= ((qd[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(6 * b + a)
    out["entry7_dir7_interleaved"] = k
    # direction octant first, then entry point
    octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2))
    out["octant_entry10"] = (octant << np.uint64(30)) | out["entry10"]
    # mid point of the chord through the box
    pm = o + d * (0.5 * (tn + np.maximum(tf, tn)))[:, None]
    out["mid10"] = morton3(quant(pm, -1.25, 1.25, 10), 10)
    out["octant_mid10"] = (octant << np.uint64(30)) | out["mid10"]
    return out


def main():
    import torch
    from dartray_b200 import capi, scenes
    dev = torch.device("cuda", 0)
    P, idx = scenes.soup(512)
    ro, rd = scenes.incoherent_rays(8_388_608)
    ctx = capi.Context(0)
    ctx.set_triangles(P, idx)
    ctx.build_bvh(capi.SPLIT_SAH, 4)
    n = ro.shape[0]
    d_hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    d_occ = torch.empty(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}

    def run(name, perm):
        a = torch.from_numpy(np.ascontiguousarray(ro[perm])).to(dev) if perm is not None else torch.from_numpy(ro).to(dev)
        b = torch.from_numpy(np.ascontiguousarray(rd[perm])).to(dev) if perm is not None else torch.from_numpy(rd).to(dev)
        tc, ta = [], []
        st = torch.cuda.current_stream().cuda_stream
        for it in range(4):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            flush.fill_(it)
            e[0].record()
            ctx.trace_closest_device(a.data_ptr(), b.data_ptr(), n, d_hits.data_ptr(), st)
            e[1].record()
            flush.fill_(it)
            e[2].record()
            ctx.trace_any_device(a.data_ptr(), b.data_ptr(), n, d_occ.data_ptr(), st)
            e[3].record()
            torch.cuda.synchronize()
            tc.append(e[0].elapsed_time(e[1]))
            ta.append(e[2].elapsed_time(e[3]))
        res[name] = {"closest_ms": min(tc[1:]), "any_ms": min(ta[1:]), "closest_mrays": n / min(tc[1:]) / 1e3,
                     "any_mrays": n / min(ta[1:]) / 1e3}
        print(name, res[name], flush=True)

    run("unsorted", None)
    for name, k in keys(ro, rd).items():
        run(name, np.argsort(k, kind="stable"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sort_experiment.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
