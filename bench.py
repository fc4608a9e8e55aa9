#!/usr/bin/env python
"""bench.py — the two hot paths of Tomato27/Monte on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (top-level keys): photon histories/s on BASELINE config 2 — 325^3 label volume, 325x325
detector, 1e8 histories per view (947 photons x 105 625 pixels), primary + scatter tallies.  One
step = one view.  At N GPUs every rank runs 947 photons per pixel of the same view (weak scaling:
N x 1e8 histories per step) and the integer tallies are summed on rank 0 with one NCCL reduce.
`fdk` (nested object): FDK voxel-updates/s (GUPS) on BASELINE config 3 — 512^3 from 720 views of a
1024x768 detector; at N GPUs z-slabs + view-sharded filter + one all-gather (strong scaling).
`value` is timed with inputs resident in HBM (CUDA events per step, L2 flushed between steps,
max over ranks); `e2e` goes through the C-ABI host-buffer calls with pinned host memory, H2D and D2H
inside the timed region.  `--impl reference` times the CPU restatement of the reference (oracle/,
all host threads) on bounded samples of the same workloads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from monte_b200 import _abi, scenes  # noqa: E402
from monte_b200.dist import split_range  # noqa: E402

PER = 947                 # photons per pixel per view and per GPU: 947 * 325^2 = 1.0003e8 histories
SM_COUNT = 148
LANES_PER_SM = 128
# algorithmic lane-instructions per unit (restated in DESIGN.md)
MC_INSTR_PER_STEP, MC_INSTR_PER_INTERACTION = 56, 200
FDK_INSTR_PER_UPDATE = 35


def peaks():
    p = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), source="measured")
    except Exception:
        pass
    p["fp32_tlane_instr"] = SM_COUNT * LANES_PER_SM * p["sm_max_mhz"] * 1e6 / 1e12
    return p


def profile_traffic(name):
    """dram bytes per launch from the committed ncu --set full capture (profiles/*.json), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None
        try:                                   # CUDA_VISIBLE_DEVICES may renumber: address the GPU by UUID
            import torch
            self.device = "GPU-" + str(torch.cuda.get_device_properties(int(device)).uuid)
        except Exception:
            pass

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx = float(f[2])
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            top = sorted(sm)[len(sm) // 2:]           # samples under load = upper half
            out.update(sm_mhz=float(np.median(top)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# =============================================================================================
# reference arm: the CPU restatement of the reference on the host cores
# =============================================================================================
_C2 = {}


def cpu_mc_sample(ob, per_sample, view=0, threads=0, seed=11):
    if not _C2:
        g, vol, lab = scenes.config_c2()
        _C2.update(g=g, vol=vol, lab=lab, tb=ob.tables_from_xs(scenes.make_xs()))
    g, vol, lab, tb = _C2["g"], _C2["vol"], _C2["lab"], _C2["tb"]
    opts = ob.mc_opts(ob.RNG_MT, seed=seed, n_threads=threads)
    t = time.perf_counter()
    _, _, res, _, _ = ob.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0), opts, per_sample, views=(view, view + 1))
    dt = time.perf_counter() - t
    return res["histories"], dt, res


def cpu_fdk_sample(ob, z_slices=2, seed=0):
    """C3 geometry, backprojection of `z_slices` central slices from all 720 views (+ the filter of
    8 views, scaled) on all host threads"""
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    g.z_begin, g.z_end = 256 - z_slices // 2, 256 - z_slices // 2 + z_slices
    rng = np.random.default_rng(seed)
    filt = rng.random((g.n_views, g.nv, g.nu), dtype=np.float32)
    t = time.perf_counter()
    ob.fdk_backproject(g, filt)
    dt = time.perf_counter() - t
    return g.nx * g.ny * z_slices * g.n_views, dt


def as_shipped_reference(ob):
    """The UNMODIFIED reference programs (oracle/_ref, compiled from /root/reference's own sources by
    oracle/Makefile) timed as whole programs on one core, where they can express a workload at all:
    CBCT_real2 = 1e7 histories of one pencil through the 65^3-scale sphere phantom (scatter tally only),
    bp3d20 = 65x65x360 projections -> 5 columns of a 256^3 volume.  SURVEY 8(d) item (1)."""
    out = {}
    try:
        if ob.have_ref("CBCT_real2") and ob.have_ref("make_image01"):
            t = time.perf_counter()
            _, c, _ = ob.ref_cbct_real2()
            dt = time.perf_counter() - t
            out["mc"] = {"value": 1e7 / dt, "unit": "histories/s", "cores": 1, "kind": "reference",
                         "sample": "unmodified monte_cpp/CBCT_real2.cpp binary, as shipped: 1e7 histories, one pixel, one view, "
                                   "whole program in %.1f s (count = %d detected scatters)" % (dt, c["count"])}
        if ob.have_ref("bp3d20"):
            p = np.random.default_rng(0).random((360, 65, 65), dtype=np.float32)
            t = time.perf_counter()
            ob.ref_bp3d20(p)
            dt = time.perf_counter() - t
            out["fdk"] = {"value": 256 * 256 * 5 * 360 / dt / 1e9, "unit": "GUPS", "cores": 1, "kind": "reference",
                          "sample": "unmodified recon/bp3d20.cpp binary, as shipped: 65x65x360 -> 256x256x5 voxel columns, "
                                    "whole program incl. filter and raw-file I/O in %.1f s" % dt}
    except Exception as ex:                                  # the binaries are optional evidence, never a reason to fail the bench
        out["error"] = repr(ex)[:200]
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    ob.build(ref=False)
    cores = os.cpu_count() or 1
    per_sample = 8                                   # 8 photons/pixel of one C2 view = 845 000 histories per step
    n_hist, t_tot = 0, 0.0
    for k in range(args.warmup + args.steps):
        n, dt, _ = cpu_mc_sample(ob, per_sample, view=k % 360, seed=100 + k)
        if k >= args.warmup:
            n_hist += n
            t_tot += dt
    v = n_hist / t_tot
    upd, dtf = cpu_fdk_sample(ob, 2)
    sample = "C2 scene, %d photons/pixel of one view per step (%.3g histories/step), oracle MT19937 double" % (per_sample, n_hist / args.steps)
    line = {
        "impl": "reference", "metric": "photon_histories_per_s", "value": v, "unit": "histories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 325^3 labels, 325x325 detector, mono 140 keV, <=5 scatters; bounded sample per step"},
        "cpu_baseline": {"value": v, "unit": "histories/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fdk": {"metric": "fdk_voxel_updates_per_s", "value": upd / dtf / 1e9, "unit": "GUPS",
                "cpu_baseline": {"value": upd / dtf / 1e9, "unit": "GUPS", "cores": cores, "kind": "port",
                                 "sample": "C3 geometry, backprojection of 2 central z-slices from 720 views (%.3g updates)" % upd},
                "e2e": {"value": upd / dtf / 1e9, "unit": "GUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    shipped = as_shipped_reference(ob)
    if shipped:
        line["as_shipped"] = shipped
    print(json.dumps(line))


# =============================================================================================
# our arm
# =============================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--fdk-steps", type=int, default=3)
    ap.add_argument("--skip-fdk", action="store_true")
    ap.add_argument("--path", default="mc", choices=["mc", "fdk"],
                    help="which hot path provides the top-level keys (default: MC, BASELINE configs[1]); the other is nested")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-c4", action="store_true", help="skip the nested polyenergetic MC block (BASELINE configs[3])")
    ap.add_argument("--mc-tracking", type=int, default=0,
                    help="monte_mc_volume.tracking_mode of the headline scene: 0 = the reference's single-majorant loop (default), "
                         "1 CLEARANCE, 3 ADAPTIVE, 4 DIRECTIONAL (same physics, fewer tentative collisions; see DESIGN.md section 3)")
    ap.add_argument("--mc-cell-log2", type=int, default=3, help="clearance cells of 2^n voxels for --mc-tracking != 0")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from monte_b200 import api
    from monte_b200 import dist as mdist

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libmonte_gpu has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    api.init(local)
    pk = peaks()
    K, W = args.steps, max(args.warmup, 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ MC, config 2
    g, vol, lab = scenes.config_c2()
    if args.mc_tracking:
        vol.tracking_mode, vol.clearance_cell_log2 = args.mc_tracking, args.mc_cell_log2
    xs = scenes.make_xs()
    spec = scenes.mono_spectrum(140.0)
    scene = api.Scene(g, vol, lab, xs, spec)
    npix = g.ny * g.nx
    per_total = PER * ws
    im0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32, device=dev)
    im5 = torch.zeros_like(im0)
    stats = torch.zeros(16, dtype=torch.int64, device=dev)

    def run_local(a0, a5, per, views, n_range):
        scene.simulate_dev(a0, a5, per, seed=20261017, views=views, n_range=n_range, d_stats=stats)

    def mc_step(k):
        v = k % g.n_views
        mdist.mc_sharded_step(run_local, im0, im5, per_total, (v, v + 1))

    for k in range(W):
        mc_step(k)
    stats.zero_()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xFF)                       # L2 flush, outside the per-step events
        ev[k][0].record()
        mc_step(W + k)
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clk = clocks.stop()
    ms_mc = sum(a.elapsed_time(b) for a, b in ev)
    ms_mc = mdist.max_over_ranks(ms_mc, dev)
    st = api.unpack_stats(stats.cpu().numpy().astype(np.uint64))
    hist_rank = st["histories"]
    hist_total = npix * PER * ws * K
    assert hist_rank == npix * PER * K, (hist_rank, npix * PER * K)
    mc_value = hist_total / (ms_mc * 1e-3)
    steps_per_hist = st["woodcock_steps"] / hist_rank
    int_per_hist = st["interactions"] / hist_rank
    instr_per_hist = MC_INSTR_PER_STEP * steps_per_hist + MC_INSTR_PER_INTERACTION * int_per_hist
    achieved = (hist_rank / (ms_mc * 1e-3)) * instr_per_hist / 1e12            # per GPU, T lane-instr/s
    mc_roof = {"bound": "fp32-issue", "achieved": achieved, "peak": pk["fp32_tlane_instr"], "unit": "Tlane-instr/s",
               "frac": achieved / pk["fp32_tlane_instr"], "traffic": profile_traffic("mc_transport_kernel"),
               "kernel": "mc_transport_kernel_v3<false,5>", "kernel_ms_per_launch": ms_mc / K,
               "model": "%d instr/Woodcock step x %.3f steps/history + %d instr/interaction x %.3f (SURVEY 8d); "
                        "peak = 148 SM x 128 lanes x %.0f MHz (%s)" % (MC_INSTR_PER_STEP, steps_per_hist,
                                                                      MC_INSTR_PER_INTERACTION, int_per_hist,
                                                                      pk["sm_max_mhz"], pk["source"]),
               "hbm": {"achieved_gbs": (2 * npix * 4 + lab.size) * K / (ms_mc * 1e-3) / 1e9, "peak_gbs": pk["hbm_gbs"],
                       "note": "tally flush + one read of the label volume per view: the path is not HBM-bound"}}

    # e2e: host buffers through the C ABI (N=1) / the sharded pipeline with host staging (N>1)
    lab_pin = torch.from_numpy(lab).pin_memory()
    h0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32).pin_memory()
    h5 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32).pin_memory()
    nr = split_range(per_total, ws, rank)

    def mc_e2e_step(k):
        v = k % g.n_views
        if ws == 1:
            api.simulate(g, vol, lab_pin.numpy(), xs, spec, per_total, seed=20261017, views=(v, v + 1),
                         out=(h0.numpy(), h5.numpy()))
        else:                                     # H2D labels, kernel, NCCL reduce, D2H on rank 0
            sc = scene
            sc.update_labels(lab_pin.numpy())
            im0[v].zero_(); im5[v].zero_()
            sc.simulate_dev(im0, im5, per_total, seed=20261017, views=(v, v + 1), n_range=nr)
            dist.reduce(im0[v:v + 1], dst=0)
            dist.reduce(im5[v:v + 1], dst=0)
            if rank == 0:
                h0[v].copy_(im0[v], non_blocking=True)
                h5[v].copy_(im5[v], non_blocking=True)
            torch.cuda.synchronize()

    for k in range(2):
        mc_e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        mc_e2e_step(2 + k)
    barrier()
    t_e2e = mdist.max_over_ranks(time.perf_counter() - t0, dev)
    mc_e2e = {"value": hist_total / t_e2e, "unit": "histories/s",
              "h2d_bytes_per_step": int(lab.size + 2 * 201 * 16 + 201 * 4 + g.n_views * 8),
              "d2h_bytes_per_step": int(2 * npix * 4 + 128), "ms_per_step": 1e3 * t_e2e / K}
    scene.close()
    del im0, im5

    # cpu baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample
    mc_cpu = None
    if rank == 0 and ws == 1 and not args.skip_cpu:
        from oracle import binding as ob
        ob.build(ref=False)
        n, dt, _ = cpu_mc_sample(ob, 400)
        mc_cpu = {"value": n / dt, "unit": "histories/s", "cores": os.cpu_count(), "kind": "port",
                  "sample": "C2 scene, view 0, 400 photons/pixel = %d histories in %.1f s (oracle: double, MT19937, OpenMP over detector rows)" % (n, dt)}
        shipped = as_shipped_reference(ob)                   # the unmodified reference programs, one core, as shipped
        if "mc" in shipped:
            mc_cpu["as_shipped"] = shipped["mc"]

    # ------------------------------------------------------------------ FDK, config 3
    fdk = None
    if not args.skip_fdk:
        try:
            fdk = bench_fdk(args, api, mdist, torch, dist, dev, ws, rank, pk, flush, barrier)
        except Exception as e:                      # the headline line is still printed
            fdk = {"error": "%s: %s" % (type(e).__name__, e)}

    # ------------------------------------------------------------------ MC, config 4 (polyenergetic), nested
    c4 = None
    if not args.skip_c4:
        try:
            c4 = bench_mc_c4(args, api, mdist, torch, dist, dev, ws, rank, flush, barrier, (g, vol, lab, xs))
        except Exception as e:                      # the headline line is still printed
            c4 = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        line = {
            "metric": "photon_histories_per_s", "value": mc_value, "unit": "histories/s",
            "n_gpus": ws, "steps": K, "warmup": W, "ms_per_step": ms_mc / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 (BASELINE configs[1]): 325^3 uint8 label volume (water cylinder r=10 + 8 Ca rods), "
                                   "325x325 detector @0.1 cm, 947 photons/pixel = 1.0003e8 histories per view per GPU, "
                                   "mono 140 keV, pencil-per-pixel source, <=5 scatters, image0+image5 tallies",
                       "histories_per_step": npix * PER * ws, "parallelism": "photon-range x%d + 1 NCCL reduce/view" % ws if ws > 1 else "single GPU",
                       "l2": "256 MiB fill between steps (outside the per-step CUDA events); the 34 MB label volume is re-read from HBM each step",
                       "tracking_mode": int(args.mc_tracking),
                       "steps_per_history": steps_per_hist, "interactions_per_history": int_per_hist,
                       "primary_fraction": st["primaries"] / hist_rank, "scatter_detected_fraction": st["scatter_detected"] / hist_rank},
            "e2e": mc_e2e, "gpu_launches": K,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "samples": clk["samples"]},
            "roofline": mc_roof, "cpu_baseline": mc_cpu,
            "wall_s_timed_region": t_wall,
            "fdk": fdk,
            "mc_c4": c4,
        }
        if mc_cpu is not None and fdk and fdk.get("cpu_baseline") and "fdk" in shipped:
            fdk["cpu_baseline"]["as_shipped"] = shipped["fdk"]
        if args.path == "fdk" and fdk and "value" in fdk:      # FDK as the top-level line, MC nested
            top = dict(fdk)
            mc = {k: line[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "cpu_baseline", "config", "gpu_launches")}
            top.update(n_gpus=ws, higher_is_better=True, vs_baseline=None, data="synthetic", mc=mc)
            line = top
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


def bench_mc_c4(args, api, mdist, torch, dist, dev, ws, rank, flush, barrier, scene_parts):
    """BASELINE configs[3]: full-scatter MC over a 120 kVp polyenergetic spectrum, photon ranges split across the
    GPUs, one NCCL reduce per view.  Same C2 scene and per-view history count as the headline; timed for the
    reference's single-majorant Woodcock loop and for the two-level majorant (tracking_mode CLEARANCE)."""
    g, vol, lab, xs = scene_parts
    spec, keep = scenes.kramers_spectrum()
    npix = g.ny * g.nx
    per_total = PER * ws
    K, W = 6, 3
    im0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32, device=dev)
    im5 = torch.zeros_like(im0)
    stats = torch.zeros(16, dtype=torch.int64, device=dev)
    out = {"metric": "photon_histories_per_s", "unit": "histories/s", "scaling": "weak", "dtype": "f32", "steps": K, "warmup": W,
           "config": {"workload": "C4 (BASELINE configs[3]) physics on the C2 scene: 120 kVp Kramers spectrum hardened by 2.5 cm of "
                                  "water (0.5 keV bins), pencil-per-pixel source, <=5 scatters, %d histories per view per GPU" % (npix * PER),
                      "parallelism": "photon-range x%d + 1 NCCL reduce/view" % ws if ws > 1 else "single GPU",
                      "l2": "256 MiB fill between steps"}}
    old_mode, old_cell = vol.tracking_mode, vol.clearance_cell_log2
    try:
        for name, mode in (("reference_loop", _abi.TRACK_GLOBAL), ("clearance", _abi.TRACK_CLEARANCE)):
            vol.tracking_mode, vol.clearance_cell_log2 = mode, 2
            scene = api.Scene(g, vol, lab, xs, spec)

            def run_local(a0, a5, per, views, n_range):
                scene.simulate_dev(a0, a5, per, seed=20261017, views=views, n_range=n_range, d_stats=stats)

            for k in range(W):
                mdist.mc_sharded_step(run_local, im0, im5, per_total, (k, k + 1))
            stats.zero_()
            barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for k in range(K):
                flush.fill_(k & 0xFF)
                ev[k][0].record()
                mdist.mc_sharded_step(run_local, im0, im5, per_total, (W + k, W + k + 1))
                ev[k][1].record()
            barrier()
            ms = mdist.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev)
            st = api.unpack_stats(stats.cpu().numpy().astype(np.uint64))
            scene.close()
            hist = npix * PER * ws * K
            out[name] = {"value": hist / (ms * 1e-3), "ms_per_step": ms / K, "steps_per_history": st["woodcock_steps"] / st["histories"],
                         "scatter_detected_fraction": st["scatter_detected"] / st["histories"],
                         "seconds_for_1e11_histories": 1e11 / (hist / (ms * 1e-3))}
    finally:
        vol.tracking_mode, vol.clearance_cell_log2 = old_mode, old_cell
    out["value"] = out["clearance"]["value"]
    out["ms_per_step"] = out["clearance"]["ms_per_step"]
    out["tracking"] = "two-level majorant, 4-voxel clearance cells (tracking_mode CLEARANCE); reference_loop = single majorant"
    return out


def bench_fdk(args, api, mdist, torch, dist, dev, ws, rank, pk, flush, barrier):
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    Kf, Wf = max(args.fdk_steps, 1), 3
    v_lo, v_hi = split_range(g.n_views, ws, rank)
    # z-slabs of equal WORK, not equal thickness: at this cone angle the end slices see the detector in
    # few views or none (cuts on the backprojector's 16-slice blocks; a rank may own several ranges)
    z_ranges = mdist.fdk_z_partition(g, ws)
    my_z = z_ranges[rank]
    n_my = sum(b - a for a, b in my_z)
    gen = torch.Generator(device=dev).manual_seed(1234)
    proj = torch.rand((g.n_views, g.nu, g.nv), device=dev, generator=gen)     # same on every rank
    filt = torch.zeros(api.fdk_filtered_shape(g), device=dev)
    slab = torch.empty((n_my, g.ny, g.nx), device=dev)      # this rank's ranges, stacked in ascending z
    z_off, o = {}, 0
    for a, b in my_z:
        z_off[a] = o
        o += b - a

    exchange = os.environ.get("MONTE_BENCH_FDK_EXCHANGE", "band")

    def sharded(src):
        if exchange == "pipelined":
            # filter own views | broadcast the view pieces in ascending order (async, NCCL stream) |
            # backproject piece r as soon as pieces r and r+1 have landed, continuing the partial sums
            mdist.fdk_sharded_pipelined(lambda a, b: api.fdk_filter_dev(g, src, filt, a, b, pad=False),
                                        lambda a, b: api.fdk_pad_views_dev(g, filt, a, b),
                                        lambda z0, z1, a, b, cont: api.fdk_backproject_views_dev(
                                            g, filt, slab[z_off[z0]:z_off[z0] + z1 - z0], z0, z1, a, b, cont),
                                        filt, g.n_views, g.nv, g.nz, z_ranges=z_ranges)
        else:
            # filter own views | ONE all_to_all of the detector-row bands the peers' slabs read | backproject
            mdist.fdk_sharded_band(lambda a, b: api.fdk_filter_dev(g, src, filt, a, b, pad=False),
                                   lambda: api.fdk_pad_dev(g, filt),
                                   lambda z0, z1: api.fdk_backproject_dev(g, filt, slab[z_off[z0]:z_off[z0] + z1 - z0], z0, z1),
                                   lambda z0, z1: api.fdk_slab_rows(g, z0, z1), filt, g.n_views, g.nv, z_ranges)

    def fdk_step():
        sharded(proj)

    for _ in range(Wf):
        fdk_step()
    barrier()
    clocks = ClockSampler(dev.index)
    clocks.start()
    tot = 0.0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_filter = t_bp = 0.0
    for _ in range(Kf):
        flush.fill_(1)
        e[0].record()
        fdk_step()
        e[1].record()
        torch.cuda.synchronize()
        tot += e[0].elapsed_time(e[1])
    # split of one step into filter / backprojection (single GPU view of the kernels)
    e[0].record()
    api.fdk_filter_dev(g, proj, filt, v_lo, v_hi, pad=False)
    e[1].record()
    for a, b in my_z:
        api.fdk_backproject_dev(g, filt, slab[z_off[a]:z_off[a] + b - a], a, b)
    e[2].record()
    torch.cuda.synchronize()
    t_filter, t_bp = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    barrier()
    clk = clocks.stop()
    tot = mdist.max_over_ranks(tot, dev)
    updates = g.nx * g.ny * g.nz * g.n_views
    gups = updates * Kf / (tot * 1e-3) / 1e9
    upd_rank = g.nx * g.ny * n_my * g.n_views
    # (voxel, view) pairs that project onto the detector (the others are skipped, as in bp3d20.cpp:116):
    # Monte-Carlo estimate with the reference's projection formulas, 2e6 samples of this rank's slab
    rs = np.random.default_rng(7)
    ns = 2_000_000
    sx = g.x0 + g.vox * rs.integers(0, g.nx, ns)
    sy = g.y0 - g.vox * rs.integers(0, g.ny, ns)
    my_slices = np.concatenate([np.arange(a, b) for a, b in my_z])
    sz = g.z0 - g.vox * rs.choice(my_slices, ns)
    beta = np.deg2rad(g.angle0_deg + g.angle_step_deg * rs.integers(0, g.n_views, ns))
    kk = g.dsd / (sx * np.cos(beta) + sy * np.sin(beta) + g.dso)
    inside = float(np.mean((np.abs(kk * (-sx * np.sin(beta) + sy * np.cos(beta))) <= g.half_u) & (np.abs(kk * sz) <= g.half_v)))
    achieved = inside * upd_rank / (t_bp * 1e-3) * FDK_INSTR_PER_UPDATE / 1e12
    roof = {"bound": "fp32-issue", "achieved": achieved, "peak": pk["fp32_tlane_instr"], "unit": "Tlane-instr/s",
            "frac": achieved / pk["fp32_tlane_instr"], "traffic": profile_traffic("fdk_backproject_kernel"),
            "kernel": "fdk_backproject_kernel", "kernel_ms_per_launch": t_bp, "filter_ms_per_launch": t_filter,
            "traffic_note": "achieved and kernel_ms_per_launch cover one whole backprojection (the library splits it into L2-sized "
                            "view-chunk launches, 4 at C3 on one GPU); traffic is the ncu DRAM figure of ONE such chunk launch",
            "model": "%d lane-instr per voxel-update (SURVEY 8d) x %.4g updates per launch x %.3f of them on the detector "
                     "(off-detector pairs are skipped per column, as the reference skips them per voxel); the kernel's fast path needs 12 SASS "
                     "instructions per update, fewer than the model's %d, so frac can exceed 1" % (FDK_INSTR_PER_UPDATE, upd_rank, inside, FDK_INSTR_PER_UPDATE),
            "on_detector_fraction": inside,
            "hbm": {"algorithmic_bytes": 4 * g.nx * g.ny * n_my + 4 * g.n_views * g.nu * g.nv,
                    "achieved_gbs": (4 * g.nx * g.ny * n_my + 4 * g.n_views * g.nu * g.nv) / (t_bp * 1e-3) / 1e9,
                    "peak_gbs": pk["hbm_gbs"], "note": "projections are read once from HBM and re-read from L1/L2; not HBM-bound"}}

    # e2e through the C ABI with pinned host buffers (N=1), or H2D views + sharded pipeline + D2H slab (N>1)
    host_proj = torch.rand((g.n_views, g.nu, g.nv)).pin_memory() if ws == 1 else torch.rand((v_hi - v_lo, g.nu, g.nv)).pin_memory()
    host_vol = torch.empty((n_my, g.ny, g.nx)).pin_memory()
    del proj
    if ws > 1:
        proj_part = torch.empty((g.n_views, g.nu, g.nv), device=dev)

    def e2e_step():
        if ws == 1:
            api.fdk(g, host_proj.numpy(), want_filtered=False, out=host_vol.numpy())
        else:
            proj_part[v_lo:v_hi].copy_(host_proj, non_blocking=True)
            sharded(proj_part)
            host_vol.copy_(slab, non_blocking=True)
            torch.cuda.synchronize()

    if ws == 1:
        del filt, slab
        torch.cuda.empty_cache()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Kf):
        e2e_step()
    barrier()
    t_e2e = mdist.max_over_ranks(time.perf_counter() - t0, dev)
    e2e = {"value": updates * Kf / t_e2e / 1e9, "unit": "GUPS", "ms_per_step": 1e3 * t_e2e / Kf,
           "h2d_bytes_per_step": int(4 * g.n_views * g.nu * g.nv), "d2h_bytes_per_step": int(4 * g.nx * g.ny * g.nz)}
    cpu = None
    if rank == 0 and ws == 1 and not args.skip_cpu:
        from oracle import binding as ob
        upd, dt = cpu_fdk_sample(ob, 4)
        cpu = {"value": upd / dt / 1e9, "unit": "GUPS", "cores": os.cpu_count(), "kind": "port",
               "sample": "C3 geometry, backprojection of 4 central z-slices from 720 views = %.3g updates in %.1f s (oracle: double, OpenMP)" % (upd, dt)}
    return {"metric": "fdk_voxel_updates_per_s", "value": gups, "unit": "GUPS", "steps": Kf, "warmup": Wf,
            "ms_per_step": tot / Kf, "scaling": "strong", "dtype": "f32",
            "config": {"workload": "C3 (BASELINE configs[2]): 512^3 volume from 720 views of a 1024x768 detector, REFERENCE weights, "
                                   "weight+ramp filter + backprojection per step",
                       "parallelism": "z-ranges of equal work x%d %s, filter by views, %s" % (
                           ws, [[list(z) for z in zr] for zr in z_ranges],
                           "view pieces broadcast in order and overlapped with the backprojection" if exchange == "pipelined"
                           else "one all_to_all of the detector-row bands each slab reads") if ws > 1 else "single GPU",
                       "l2": "256 MiB fill between steps; projections (2.26 GB) exceed L2"},
            "e2e": e2e, "gpu_launches": 3 * Kf, "roofline": roof, "cpu_baseline": cpu,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "seconds_for_512cube_720views": tot / Kf * 1e-3}


if __name__ == "__main__":
    main()
