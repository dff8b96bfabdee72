#!/usr/bin/env python
"""Benchmark of the ray-cast hot path (BASELINE.json configs[1]).

A step = one pass over one batch of synthetic rays on the `soup_1m` scene (1,015,810 triangles,
SAH BVH, maxnodeprims 4): closest-hit on 8,388,608 coherent primary rays, closest-hit on 8,388,608
incoherent rays, any-hit (shadow) on the incoherent set.  `value` = rays of all ranks / step time with
the rays resident in HBM; `e2e` = the same step through the host-buffer C-ABI calls
(drt_trace_closest / drt_trace_any) from pinned host memory, H2D + kernel + D2H inside the timed
region.  `--impl reference` times the CPU restatement of the reference path (oracle/) instead.
"""
from __future__ import annotations

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

METRIC = "Mrays/s (closest-hit + shadow)"
N_SPHERES = 512
COH_W, COH_H = 4096, 2048
N_INCOH = 8_388_608


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def microbench():
    """FP32 FMA rate and L2-resident read bandwidth by working-set size, measured on this pool's B200 by tools/microbench.cu
    (MEASURED_PEAKS.json holds neither): the newest profiles/r0N_microbench.json."""
    for name in ("r02_microbench.json", "r01_microbench.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), f"profiles/{name}"
        except Exception:
            continue
    return None, None


def l2_peak(working_set_bytes):
    """Measured L2 read bandwidth (GB/s) at the scene's working-set size: linear interpolation between the measured sizes."""
    mb, src = microbench()
    if mb is None:
        return None, None
    pts = sorted((float(k[:-2]), v) for k, v in mb["l2_read_gb_per_s"].items())  # "64MB" -> (64 MiB, GB/s)
    x = working_set_bytes / 2.0 ** 20
    if x <= pts[0][0]:
        return pts[0][1], src
    for (x0, y0), (x1, y1) in zip(pts, pts[1:]):
        if x <= x1:
            return y0 + (y1 - y0) * (x - x0) / (x1 - x0), src
    return pts[-1][1], src


def roofline_block(alg_bytes, fp_ops, ms, working_set_bytes, hbm_gbs, peak_src):
    """The contract's roofline object for a kernel whose working set is L2-resident: SURVEY 8d asks for both candidate
    rooflines (memory and FP32 issue) and to call the one that binds the bound.  ncu (profiles/r02_summary.md) shows DRAM
    at ~4 % of its peak and the L2 hit rate above 80 % for the 1 M-triangle BVH, so the memory roofline of this kernel is
    the L2's, measured at the working-set size; the HBM-copy and FP32 figures are kept beside it."""
    t = ms * 1e-3
    achieved = alg_bytes / t / 1e9
    mb, mb_src = microbench()
    l2, _ = l2_peak(working_set_bytes)
    out = {"bound": "l2" if l2 else "hbm", "achieved": achieved, "peak": l2 or hbm_gbs, "unit": "GB/s",
           "frac": achieved / (l2 or hbm_gbs),
           "peak_source": (f"L2 read bandwidth at a {working_set_bytes / 2.0 ** 20:.0f} MiB working set, tools/microbench.cu ({mb_src})"
                           if l2 else peak_src),
           "hbm": {"achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "peak_source": peak_src,
                   "note": "the same algorithmic bytes against the HBM copy peak: the working set is L2-resident, HBM is nearly idle"}}
    if mb is not None:
        fp = 2.0 * mb["fp32_fma_per_s"]
        out["fp32"] = {"achieved": fp_ops / t / 1e12, "peak": fp / 1e12, "unit": "TFLOP/s", "frac": fp_ops / t / fp,
                       "note": "20 x nodes + 51 x prims reference operations per ray (SURVEY 8d) against the measured FP32 FMA rate"}
    return out


class ClockSampler(threading.Thread):
    """Samples SM clocks / clock-event (throttle) reasons while the timed region runs: NVML polled every 2 ms (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints; spawning nvidia-smi takes
    longer than a step), falling back to nvidia-smi when pynvml is unavailable."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []  # (sm_mhz, sm_max_mhz, [active reason names])
        self.stop_flag = threading.Event()
        self.source = "nvml"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(v) for v in vis.split(",") if v.strip().isdigit()]
            phys = ids[index] if index < len(ids) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        bits = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        masks = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                 n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        self.rows.append((float(sm), float(mx), [nm for nm, m in zip(self.NAMES, masks) if bits & m]))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6 and r[0].replace(".", "").isdigit():
            self.rows.append((float(r[0]), float(r[1]), [nm for k, nm in enumerate(self.NAMES) if r[2 + k] == "Active"]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.002 if self._nvml is not None else 0.05)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = sorted({nm for r in self.rows for nm in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def make_workload(rank: int):
    from dartray_b200 import scenes
    P, idx = scenes.soup(N_SPHERES)
    coh = scenes.coherent_rays(COH_W, COH_H)
    # each rank traces its own incoherent set (weak scaling): PCG32 seeds 2/3 on rank 0 as SURVEY §8d pins
    inc = scenes.incoherent_rays(N_INCOH, seed_origin=2 + 16 * rank, seed_target=3 + 16 * rank)
    return P, idx, coh, inc


def cpu_sample(coh, inc, n_sample):
    """Bounded sample of the step for the CPU legs: every k-th ray of each set."""
    ks = max(1, coh[0].shape[0] // n_sample)
    ki = max(1, inc[0].shape[0] // n_sample)
    cs = (np.ascontiguousarray(coh[0][::ks]), np.ascontiguousarray(coh[1][::ks]))
    is_ = (np.ascontiguousarray(inc[0][::ki]), np.ascontiguousarray(inc[1][::ki]))
    return cs, is_


def run_cpu_step(orc, cs, is_, threads):
    t0 = time.perf_counter()
    orc.trace_closest(cs[0], cs[1], nthreads=threads)
    orc.trace_closest(is_[0], is_[1], nthreads=threads)
    orc.trace_any(is_[0], is_[1], nthreads=threads)
    dt = time.perf_counter() - t0
    return (cs[0].shape[0] + 2 * is_[0].shape[0]), dt


def config_dict(n_gpus, sample=None):
    d = {
        "workload": "soup_1m ray cast: 1,015,810-triangle procedural mesh, SAH BVH (maxnodeprims 4); per step "
                    "8,388,608 coherent closest-hit + 8,388,608 incoherent closest-hit + 8,388,608 incoherent "
                    "any-hit (shadow) rays per GPU",
        "baseline_config": "BASELINE.json configs[1]",
        "rays_per_step_per_gpu": COH_W * COH_H + 2 * N_INCOH,
        "parallelism": f"rays sharded, BVH replicated x{n_gpus}",
        "l2": "L2 flushed (256 MiB write) between timed steps; ray buffers (268 MB/launch) exceed the 126 MB L2",
    }
    if sample:
        d["sample"] = sample  # the CPU arm times a bounded sample of the step, not the whole step
    return d


# ---- render legs (BASELINE.json configs[2] and configs[3]) ------------------------------------------------
RENDER_RES = (1920, 1080)
AO_RAYS = 64
PATH_SPP = 256


def render_configs():
    from dartray_b200 import host
    ao = (host.Sampler(kind=host.SAMPLER_STRATIFIED, xs=1, ys=1, jitter=False),
          host.Integrator(kind=host.INTEGRATOR_AO, ao_nsamples=AO_RAYS))
    path = (host.Sampler(kind=host.SAMPLER_LD, spp=PATH_SPP), host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5))
    return ao, path


def render_legs(ctx_soup, soup_mesh, rank, world, barrier):
    """configs[2]: ambient occlusion on soup_1m; configs[3]: cornell_synth path tracing, 256 spp.  Pixels are
    sharded over the ranks (interleaved 1024-pixel blocks), the films summed with NCCL; timed as the blocking
    C-ABI call (wall clock between barriers, max over ranks)."""
    import torch
    import torch.distributed as dist
    from dartray_b200 import capi, distributed, host, scenes

    dev = torch.device("cuda", ctx_soup.device)
    out = {}
    (ao_s, ao_i), (pt_s, pt_i) = render_configs()
    cam_soup = host.PerspectiveCamera(host.look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), fov=40.0)
    sb, cam_cornell = scenes.cornell_synth()
    ctx_c = capi.Context(ctx_soup.device)
    host.upload_scene(ctx_c, sb.arrays())
    arrays_of = {"ao": None, "path": sb.arrays(), "path_f32": sb.arrays()}  # ao: the bare soup_1m mesh the ray-cast step uses
    # configs[3] twice: with the binary64 shading kernels that replay the reference's arithmetic (per-sample parity ~1e-6) and with
    # drt_set_shading_precision(DRT_PRECISION_F32), which north_star's 3-sigma bar for the path tracer admits (tests/test_precision_gpu.py)
    legs = [("ao", ctx_soup, cam_soup, ao_s, ao_i, None), ("path", ctx_c, cam_cornell, pt_s, pt_i, capi.PRECISION_F64),
            ("path_f32", ctx_c, cam_cornell, pt_s, pt_i, capi.PRECISION_F32)]
    for name, ctx, cam, smp, integ, precision in legs:
        film = host.Film(*RENDER_RES)
        if precision is not None:
            ctx.set_shading_precision(precision)
        # warm-up: one untimed render of the SAME configuration, so that the wavefront allocation (up to ~10 GB, sized by the
        # batch), the first launches and the NCCL channel are outside the timed region, as they are for every later frame
        host.configure_render(ctx, cam, film, smp, integ)
        distributed.render_sharded(ctx, rank, world)
        host.configure_render(ctx, cam, film, smp, integ)
        ctx.film_clear()
        l0 = ctx.kernel_launches
        barrier()
        t0 = time.perf_counter()
        parts = distributed.render_sharded(ctx, rank, world)
        t_done = time.perf_counter() - t0
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        # where the time goes: this rank's render, its film sum (which also waits for the slowest rank), its own total
        pr = torch.tensor([parts["render_s"], parts["film_sum_s"], t_done], dtype=torch.float64, device=dev)
        pr_min = pr.clone()
        if world > 1:
            dist.all_reduce(pr, op=dist.ReduceOp.MAX)
            dist.all_reduce(pr_min, op=dist.ReduceOp.MIN)
        st = ctx.render_stats()
        cnt = torch.tensor([st["camera_samples"], st["closest_rays"], st["shadow_rays"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        sec = float(dt.item())
        samples, closest, shadow = (float(v) for v in cnt.tolist())
        rgb = ctx.film_read()["rgb"]
        out[name] = {"seconds": sec, "camera_samples": int(samples), "samples_per_s": samples / sec,
                     "mrays_per_s": (closest + shadow) / sec / 1e6, "closest_rays": int(closest), "shadow_rays": int(shadow),
                     "launches_rank0": int(ctx.kernel_launches - l0), "mean_rgb": [float(v) for v in rgb.mean(axis=(0, 1))],
                     "per_rank": {"render_s_max": float(pr[0]), "render_s_min": float(pr_min[0]), "film_sum_s_max": float(pr[1]),
                                  "film_sum_s_min": float(pr_min[1]), "render_plus_sum_s_max": float(pr[2])}}
        # ---- roofline of the leg (rank 0's shard; untimed extra passes) ------------------------------------------------
        # (1) device time per kernel class: CUDA events after every launch of one more render of the same configuration
        ctx.set_render_profiling(capi.PROFILE_TIME)
        ctx.film_clear()
        ctx.render_shard(rank, world)
        prof = ctx.render_profile()
        my_samples = ctx.render_stats()["camera_samples"]
        # (2) reference work of the rays (slab and primitive tests of the reference's walk over its binary BVH),
        # counted by the counting kernel on every queue; the path leg counts a 16 spp render of the same film (1/16 of the
        # rays, the per-sample work is what scales)
        work_smp = smp if name == "ao" else host.Sampler(kind=host.SAMPLER_LD, spp=16)
        host.configure_render(ctx, cam, film, work_smp, integ)
        ctx.set_render_profiling(capi.PROFILE_WORK)
        ctx.film_clear()
        ctx.render_shard(rank, world)
        wk, wst = ctx.render_profile(), ctx.render_stats()
        ctx.set_render_profiling(0)
        floats_per_sample = int(ctx.pixel_samples(RENDER_RES[0] // 2, RENDER_RES[1] // 2).shape[1])
        host.configure_render(ctx, cam, film, smp, integ)
        ws = max(1, wst["camera_samples"])
        b_closest = 32 * wk["closest"]["nodes_visited"] + 36 * wk["closest"]["prims_tested"] + 48 * wk["closest"]["rays"]
        b_any = 32 * wk["any"]["nodes_visited"] + 36 * wk["any"]["prims_tested"] + 33 * wk["any"]["rays"]
        ops = 20 * (wk["closest"]["nodes_visited"] + wk["any"]["nodes_visited"]) + 51 * (wk["closest"]["prims_tested"] + wk["any"]["prims_tested"])
        per_sample_bytes = (b_closest + b_any) / ws + 4 * floats_per_sample + 16
        ms_all = sum(prof["ms"].values())
        dom = max(prof["ms"], key=prof["ms"].get)
        pk, pk_src = peaks()
        wset = ctx.bvh_info()["device_bytes"]
        rl = roofline_block(per_sample_bytes * my_samples, ops / ws * my_samples, ms_all, wset, pk["hbm_gbs"], pk_src)
        rl.update({
            "kernel": f"whole leg ({len([k for k, v in prof['launches'].items() if v])} kernel classes, {sum(prof['launches'].values())} launches); dominant class: {dom}",
            "per_sample": {"bytes": per_sample_bytes, "traversal_bytes": (b_closest + b_any) / ws, "sampler_bytes": 4 * floats_per_sample,
                           "film_bytes": 16, "fp_ops": ops / ws,
                           "closest_rays": wk["closest"]["rays"] / ws, "shadow_rays": wk["any"]["rays"] / ws,
                           "nodes_visited": (wk["closest"]["nodes_visited"] + wk["any"]["nodes_visited"]) / ws,
                           "prims_tested": (wk["closest"]["prims_tested"] + wk["any"]["prims_tested"]) / ws},
            "device_ms_by_class": prof["ms"], "launches_by_class": prof["launches"],
            "share_by_class": {k: v / ms_all for k, v in prof["ms"].items()},
            "work_counted_on": f"{wst['camera_samples']} camera samples ({'the timed configuration' if name == 'ao' else '16 spp of the same film'}), rank 0's shard",
            "traffic": None,
        })
        # the traversal kernels on their own: reference bytes of the class's rays over the class's device time
        scale = my_samples / ws
        trav = {}
        for cls, b in (("trace_closest", b_closest), ("trace_any", b_any)):
            if prof["ms"][cls] > 0:
                l2, _ = l2_peak(wset)
                a = b * scale / (prof["ms"][cls] * 1e-3) / 1e9
                trav[cls] = {"achieved": a, "unit": "GB/s", "frac_l2": a / l2 if l2 else None, "frac_hbm": a / pk["hbm_gbs"],
                             "mrays_per_s": wk["closest" if cls == "trace_closest" else "any"]["rays"] * scale / prof["ms"][cls] / 1e3}
        rl["traversal"] = trav
        out[name]["roofline"] = rl
        # ---- end to end through the C ABI from host arrays: scene upload (+ BVH build) + render + film read-back ------------
        arrays = arrays_of[name]
        barrier()
        t0 = time.perf_counter()
        ctx2 = capi.Context(ctx.device)
        if arrays is None:
            ctx2.set_triangles(*soup_mesh)
            ctx2.build_bvh(capi.SPLIT_SAH, 4)
        else:
            host.upload_scene(ctx2, arrays)
        host.configure_render(ctx2, cam, film, smp, integ)
        if precision is not None:
            ctx2.set_shading_precision(precision)
        t_up = time.perf_counter() - t0
        distributed.render_sharded(ctx2, rank, world)
        img = ctx2.film_read()["rgb"]
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        h2d = int(sum(v.nbytes for v in (arrays.values() if arrays is not None else soup_mesh) if isinstance(v, np.ndarray)))
        out[name]["e2e"] = {"seconds": float(t_e2e.item()), "value": samples / float(t_e2e.item()), "unit": "camera samples/s",
                            "upload_and_build_seconds_rank0": t_up, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(img.nbytes),
                            "what": "fresh context: drt_set_* scene arrays from host memory, drt_build_bvh, drt_render_shard, film sum, drt_film_read"}
        del ctx2
    out["ao"]["config"] = (f"BASELINE.json configs[2]: soup_1m, {RENDER_RES[0]}x{RENDER_RES[1]}, 1 camera sample/pixel at the "
                           f"pixel centre, {AO_RAYS} AO rays per hit")
    out["path"]["config"] = (f"BASELINE.json configs[3]: cornell_synth, {RENDER_RES[0]}x{RENDER_RES[1]}, lowdiscrepancy "
                             f"{PATH_SPP} spp, path maxdepth 5, box filter; pixel blocks sharded over {world} GPU(s), film "
                             "summed with NCCL; shading kernels in binary64 (DRT_PRECISION_F64: the reference's arithmetic, "
                             "per-sample parity ~1e-6)")
    out["path_f32"]["config"] = out["path"]["config"].replace(
        "binary64 (DRT_PRECISION_F64: the reference's arithmetic, per-sample parity ~1e-6)",
        "float32 (DRT_PRECISION_F32: per-pixel means within 3 sigma, tests/test_precision_gpu.py)")
    out["timing"] = "wall clock of the blocking drt_render_shard + film all-reduce, between barriers, max over ranks"
    return out


def cpu_render_sample(threads):
    """The oracle on a bounded sample of configs[3]: a 480x270 film of the same camera at 64 spp."""
    from dartray_b200 import host, scenes
    from tests.oracle_lib import Oracle
    sb, cam = scenes.cornell_synth()
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(480, 270), host.Sampler(kind=host.SAMPLER_LD, spp=64),
                          host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=5))
    t0 = time.perf_counter()
    o.render(0, 1, threads)
    dt = time.perf_counter() - t0
    st = o.render_stats()
    return {"path_samples_per_s": st["camera_samples"] / dt, "path_mrays_per_s": (st["closest_rays"] + st["shadow_rays"]) / dt / 1e6,
            "sample": f"cornell_synth 480x270 (same camera), 64 spp, maxdepth 5 ({st['camera_samples']} camera samples, {dt:.1f} s)"}


def reference_arm(args):
    """CPU restatement of the reference path (oracle/), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.oracle_lib import Oracle
    threads = os.cpu_count() or 1
    P, idx, coh, inc = make_workload(0)
    orc = Oracle()
    orc.set_triangles(P, idx)
    orc.build_bvh(2, 4)
    n_sample = 1 << 21  # rays per set and step: ~0.5 s of CPU work per step on 16 threads
    cs, is_ = cpu_sample(coh, inc, n_sample)
    for _ in range(args.warmup):
        run_cpu_step(orc, cs, is_, threads)
    rays = 0
    t = 0.0
    for _ in range(args.steps):
        r, dt = run_cpu_step(orc, cs, is_, threads)
        rays += r
        t += dt
    val = rays / t / 1e6
    sample = f"every {coh[0].shape[0] // cs[0].shape[0]}-th ray of each set ({cs[0].shape[0]} + 2x{is_[0].shape[0]} rays/step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus, sample),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the DartRay CPU path (oracle/), not the Dart VM: no Dart SDK in the image",
                         "render": cpu_render_sample(threads)},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the configs[2]/[3] render legs")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from dartray_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dartray_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout when the communicator is
        # created, so create it (first collective) with fd 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    P, idx, coh, inc = make_workload(rank)
    ctx = capi.Context(local)
    ctx.set_triangles(P, idx)
    ctx.build_bvh(capi.SPLIT_SAH, 4)
    info = ctx.bvh_info()

    n_coh, n_inc = coh[0].shape[0], inc[0].shape[0]
    d_coh = [torch.from_numpy(a).to(dev) for a in coh]
    d_inc = [torch.from_numpy(a).to(dev) for a in inc]
    d_hits = torch.empty((max(n_coh, n_inc), 4), dtype=torch.float32, device=dev)
    d_occ = torch.empty(n_inc, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def launch(kind):
        if kind == 0:
            ctx.trace_closest_device(d_coh[0].data_ptr(), d_coh[1].data_ptr(), n_coh, d_hits.data_ptr(), stream)
        elif kind == 1:
            ctx.trace_closest_device(d_inc[0].data_ptr(), d_inc[1].data_ptr(), n_inc, d_hits.data_ptr(), stream)
        else:
            ctx.trace_any_device(d_inc[0].data_ptr(), d_inc[1].data_ptr(), n_inc, d_occ.data_ptr(), stream)

    # algorithmic work of the dominant launch (incoherent closest-hit): counted by the kernel's
    # counting variant (identical to the oracle's counters, tests/test_trace_gpu.py)
    ctx.set_counting(True)
    work = []
    for kind in range(3):
        launch(kind)
        torch.cuda.synchronize()
        work.append(ctx.counters())
    ctx.set_counting(False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        for kind in range(3):
            launch(kind)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    clocks = ClockSampler(local)
    barrier()
    launches0 = ctx.kernel_launches
    clocks.start()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)  # L2 flush, outside the per-step event brackets
        ev[s][0].record()
        for kind in range(3):
            launch(kind)
            ev[s][kind + 1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks.stop_flag.set()
    clocks.join()
    gpu_launches = ctx.kernel_launches - launches0
    step_ms = [ev[s][0].elapsed_time(ev[s][3]) for s in range(args.steps)]
    per_kind_ms = [float(np.mean([ev[s][k].elapsed_time(ev[s][k + 1]) for s in range(args.steps)])) for k in range(3)]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    rays_per_step = n_coh + 2 * n_inc
    value = world * rays_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end through the host-buffer C ABI (pinned host memory in, host memory out) ----
    h_coh = [torch.from_numpy(a).pin_memory() for a in coh]
    h_inc = [torch.from_numpy(a).pin_memory() for a in inc]
    h_hits = torch.empty((max(n_coh, n_inc), 4), dtype=torch.float32).pin_memory()
    h_occ = torch.empty(n_inc, dtype=torch.uint8).pin_memory()
    hits_np = h_hits.numpy().view(capi.HIT_DTYPE).reshape(-1)

    def e2e_step():
        ctx.trace_closest(h_coh[0].numpy(), h_coh[1].numpy(), out=hits_np[:n_coh])
        ctx.trace_closest(h_inc[0].numpy(), h_inc[1].numpy(), out=hits_np[:n_inc])
        ctx.trace_any(h_inc[0].numpy(), h_inc[1].numpy(), out=h_occ.numpy())

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * rays_per_step * e2e_steps / float(e2e_t.item()) / 1e6

    render = None if args.no_render else render_legs(ctx, (P, idx), rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    w = work[1]
    alg_bytes = 32 * w["nodes_visited"] + 36 * w["prims_tested"] + 48 * w["rays"]
    achieved = alg_bytes / (per_kind_ms[1] * 1e-3) / 1e9
    # dram__bytes_read + dram__bytes_write of this launch from the round's `ncu --set full` capture (tools/capture_round.sh
    # writes the file next to the ncu metrics it was read from)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
        "breakdown": {
            "coherent_closest_mrays": n_coh / per_kind_ms[0] / 1e3, "incoherent_closest_mrays": n_inc / per_kind_ms[1] / 1e3,
            "incoherent_any_mrays": n_inc / per_kind_ms[2] / 1e3, "ms": per_kind_ms,
            "wall_s_timed_region": t_wall,
        },
        "roofline": dict(
            roofline_block(alg_bytes, 20 * w["nodes_visited"] + 51 * w["prims_tested"], per_kind_ms[1], info["device_bytes"],
                           pk["hbm_gbs"], pk_src),
            traffic=traffic, traffic_source=traffic_src,
            kernel="traceQKernel<closest, triangles-only leaf code> on the incoherent set (trace_fast2.cu)",
            algorithmic_bytes_per_launch=alg_bytes,
            per_ray={"nodes_visited": w["nodes_visited"] / w["rays"], "prims_tested": w["prims_tested"] / w["rays"],
                     "bytes": alg_bytes / w["rays"]},
            fp_ops_per_launch=20 * w["nodes_visited"] + 51 * w["prims_tested"],
            note="algorithmic bytes = 32*nodes + 36*prims + 48 per ray on the REFERENCE binary BVH (SURVEY 8d), whatever the "
                 "kernel's own node format fetches"),
        "e2e": {"value": e2e_val, "unit": "Mrays/s",
                "h2d_bytes_per_step": 32 * (n_coh + 2 * n_inc), "d2h_bytes_per_step": 16 * (n_coh + n_inc) + n_inc},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks.summary(),
        "bvh": {"nodes": info["n_nodes"], "device_bytes": info["device_bytes"], "build_seconds": info["build_seconds"]},
    }
    if render is not None:
        line["render"] = render
    if not args.no_cpu_baseline:
        from tests.oracle_lib import Oracle
        threads = os.cpu_count() or 1
        orc = Oracle()
        orc.set_triangles(P, idx)
        orc.build_bvh(2, 4)
        # bounded sample: the whole step (every ray of the three sets) repeated until >= 10 s of CPU work
        cs, is_ = cpu_sample(coh, inc, n_coh)
        run_cpu_step(orc, (cs[0][:4096], cs[1][:4096]), (is_[0][:4096], is_[1][:4096]), threads)
        r, dt, reps = 0, 0.0, 0
        while dt < 10.0 and reps < 8:
            r1, dt1 = run_cpu_step(orc, cs, is_, threads)
            r, dt, reps = r + r1, dt + dt1, reps + 1
        line["cpu_baseline"] = {
            "value": r / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"the full step ({r // reps} rays: all three ray sets) x {reps} passes, {dt:.1f} s of CPU time on {threads} threads",
            "note": "C++ restatement of the DartRay CPU path (oracle/), not the Dart VM: no Dart SDK in the image",
        }
        if render is not None:
            line["cpu_baseline"]["render"] = cpu_render_sample(threads)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
