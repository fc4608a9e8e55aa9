#!/usr/bin/env python
"""bench.py — the two hot paths of Tomato27/Monte on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (top-level keys): photon histories/s on BASELINE configs[1] (C2) — 325^3 label volume, 325x325
detector, 1e8 histories per view (947 photons x 105 625 pixels), primary + scatter tallies.  One step = one
view.  At N GPUs every rank runs 947 photons per pixel of the same view (weak scaling: N x 1e8 histories per
step) and the integer tallies are summed on rank 0 with one NCCL reduce.
Nested objects, one per remaining BASELINE config, each with its own value / e2e / roofline:
  fdk    C3  FDK 512^3 from 720 views of 1024x768 (strong scaling: view-sharded filter, z-slabs of equal work)
  mc_c4  C4  full-scatter MC over a 120 kVp spectrum, 1e11 histories at N = 8 (1e10 at N = 1: stated)
  fdk_c5 C5  backprojection sweep 256^3 x 360, 512^3 x 720, 1024^3 x 1440
  c1     C1  the reference's own CPU-sized case (65^3, 65x65, 1e6 photons/view x 360 + 256^3 FDK) through the C ABI
`value` is timed with inputs resident in HBM (CUDA events per step, L2 flushed between steps, max over ranks,
one process per GPU + torch.distributed); `e2e` goes through the C-ABI host-buffer calls of libmonte_gpu with
pinned host memory, H2D and D2H inside the timed region — at N > 1 that is ONE call on rank 0 with the library
bound to all N devices (monte_gpu_init(N, ids): the sharding, the tally reduce and the row-band exchange happen
inside the library); the other ranks idle at a host barrier meanwhile.  `parity` (N >= 1): outside the timed
regions the sharded results are compared with a one-GPU computation of the same thing, bit for bit, and FDK
probe voxels with the CPU oracle; a mismatch makes the run exit non-zero.
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads, the count it really used
is printed) on bounded samples of the same workloads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from monte_b200 import _abi, scenes  # noqa: E402
from monte_b200.dist import split_range  # noqa: E402

PER = 947                 # photons per pixel per view and per GPU: 947 * 325^2 = 1.0003e8 histories
SM_COUNT = 148
LANES_PER_SM = 128
SEED = 20261017
# algorithmic lane-instructions per unit (restated in DESIGN.md)
MC_INSTR_PER_STEP, MC_INSTR_PER_INTERACTION = 56, 200
FDK_INSTR_PER_UPDATE = 35

# one workload string per BASELINE config, shared verbatim by both arms (the sample a CPU arm runs is stated apart)
WORKLOAD_C2 = ("C2 (BASELINE configs[1]): 325^3 uint8 label volume (water cylinder r=10 + 8 Ca rods), 325x325 detector @0.1 cm, "
               "360 views, mono 140 keV, pencil-per-pixel source, <=5 scatters, image0+image5 tallies; one step = the histories of one view")
WORKLOAD_C3 = ("C3 (BASELINE configs[2]): 512^3 volume from 720 views of a 1024x768 detector, REFERENCE weights, "
               "weight+ramp filter + backprojection per step")
WORKLOAD_C4 = ("C4 (BASELINE configs[3]): full-scatter MC on the C2 scene over a 120 kVp Kramers spectrum hardened by 2.5 cm of water "
               "(0.5 keV bins), pencil-per-pixel source, <=5 scatters, histories split across the GPUs, one tally reduce per view")
WORKLOAD_C5 = "C5 (BASELINE configs[4]): FDK sweep N^3 volume x V views, square detector of 1.5 N pixels (every voxel on the detector)"
WORKLOAD_C1 = ("C1 (BASELINE configs[0]): water/Ca cylinder phantom 65^3 @0.5 cm, 65x65 detector @0.5 cm, 1e6 photons/view (237 per pixel) x 360 "
               "views, then recon/bp3d20 geometry: 65x65x360 maps -> 256^3 volume")


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


def kernel_counters(name):
    """Counters of the committed ncu --set full capture of a kernel (profiles/kernel_counters.json): DRAM bytes, executed
    warp instructions and the active lanes per instruction of one captured launch, plus the units (histories / voxel
    updates on the detector) that launch processed.  None if the file has no entry."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def utilisation(roof, counters, units_per_s, pk, launches_per_unit_of_work=1.0):
    """issue_util = executed lane-instructions per second / (148 SM x 128 lanes x clock): what fraction of the FP32/INT
    issue slots the kernel really fills.  Executed lane-instructions per unit come from the committed ncu capture
    (smsp__inst_executed.sum x smsp__thread_inst_executed_per_inst_executed.ratio / units of the captured launch); the rate
    is this run's.  Unlike `frac` (algorithmic instructions of SURVEY 8d / peak) it cannot exceed 1."""
    if not counters:
        roof.update(issue_util=None, instr_per_unit_executed=None)
        return
    lane_instr = counters["inst_executed"] * counters["lanes_per_inst"]
    per_unit = lane_instr / counters["units"]
    roof.update(instr_per_unit_executed=per_unit, issue_util=units_per_s * per_unit / 1e12 / pk["fp32_tlane_instr"],
                issue_util_source=counters.get("source"))


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


def cpu_mc_sample(ob, per_sample, view=0, threads=0, seed=11, spectrum=None):
    if not _C2:
        g, vol, lab = scenes.config_c2()
        _C2.update(g=g, vol=vol, lab=lab, tb=ob.tables_from_xs(scenes.make_xs()))
    g, vol, lab, tb = _C2["g"], _C2["vol"], _C2["lab"], _C2["tb"]
    opts = ob.mc_opts(ob.RNG_MT, seed=seed, n_threads=threads)
    t = time.perf_counter()
    _, _, res, _, _ = ob.mc_run(g, vol, lab, tb, spectrum if spectrum is not None else scenes.mono_spectrum(140.0), opts, per_sample,
                                views=(view, view + 1))
    dt = time.perf_counter() - t
    return res["histories"], dt, res


def cpu_fdk_sample(ob, z_slices=2, seed=0):
    """C3 geometry, backprojection of `z_slices` central slices from all 720 views on all host threads"""
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    g.z_begin, g.z_end = 256 - z_slices // 2, 256 - z_slices // 2 + z_slices
    rng = np.random.default_rng(seed)
    filt = rng.random((g.n_views, g.nv, g.nu), dtype=np.float32)
    t = time.perf_counter()
    ob.fdk_backproject(g, filt)
    dt = time.perf_counter() - t
    return g.nx * g.ny * z_slices * g.n_views, dt


def cpu_c1(ob, threads, n_views=360):
    """BASELINE configs[0] at its stated size on the host cores: 1e6 photons/view x n_views views of the 65^3 phantom
    (oracle, double, MT19937), then the bp3d20 reconstruction 65x65x360 -> 256^3 (filter + full backprojection)."""
    g, vol, lab = scenes.config_c1()
    per = 237                                                   # 237 x 65^2 = 1.001e6 photons per view
    tb = ob.tables_from_xs(scenes.make_xs())
    t = time.perf_counter()
    _, _, res, _, _ = ob.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0), ob.mc_opts(ob.RNG_MT, seed=5, n_threads=threads), per,
                                views=(0, n_views))
    dt_mc = time.perf_counter() - t
    gf = _abi.bp3d20_geom().full_roi()
    gf.mask_r2 = 118 * 118
    proj = np.random.default_rng(0).random((360, 65, 65), dtype=np.float32)
    zs = 256 if n_views >= 360 else max(2, 256 * n_views // 360)   # shortened runs reconstruct fewer slices
    gf.z_begin, gf.z_end = 128 - zs // 2, 128 - zs // 2 + zs
    t = time.perf_counter()
    ob.fdk(gf, proj)
    dt_fdk = time.perf_counter() - t
    upd = 256 * 256 * zs * 360
    return {"mc": {"value": res["histories"] / dt_mc, "unit": "histories/s", "histories": res["histories"], "seconds": dt_mc},
            "fdk": {"value": upd / dt_fdk / 1e9, "unit": "GUPS", "voxel_updates": upd, "seconds": dt_fdk},
            "sample": "%d of 360 views (%.3g histories), %d of 256 z-slices" % (n_views, res["histories"], zs)}


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
    # torch.distributed.run exports OMP_NUM_THREADS=1: ask for the cores explicitly and report what OpenMP grants
    threads = ob.set_threads(cores)
    per_sample = 8                                   # 8 photons/pixel of one C2 view = 845 000 histories per step
    n_hist, t_tot = 0, 0.0
    for k in range(args.warmup + args.steps):
        n, dt, _ = cpu_mc_sample(ob, per_sample, view=k % 360, seed=100 + k, threads=threads)
        if k >= args.warmup:
            n_hist += n
            t_tot += dt
    v = n_hist / t_tot
    upd, dtf = cpu_fdk_sample(ob, 2)
    spec4, _keep = scenes.kramers_spectrum()
    n4, dt4, _ = cpu_mc_sample(ob, 4, view=0, seed=7, threads=threads, spectrum=spec4)
    sample = "%d photons/pixel of one view per step (%.3g histories/step), oracle: double, MT19937, OpenMP over detector rows" % (per_sample, n_hist / args.steps)
    line = {
        "impl": "reference", "metric": "photon_histories_per_s", "value": v, "unit": "histories/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_C2, "sample": sample, "histories_per_step": n_hist / args.steps},
        "cpu_baseline": {"value": v, "unit": "histories/s", "cores": threads, "host_cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fdk": {"metric": "fdk_voxel_updates_per_s", "value": upd / dtf / 1e9, "unit": "GUPS",
                "config": {"workload": WORKLOAD_C3, "sample": "backprojection of 2 central z-slices from 720 views (%.3g updates)" % upd},
                "cpu_baseline": {"value": upd / dtf / 1e9, "unit": "GUPS", "cores": threads, "kind": "port",
                                 "sample": "backprojection of 2 central z-slices from 720 views (%.3g updates)" % upd},
                "e2e": {"value": upd / dtf / 1e9, "unit": "GUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "mc_c4": {"metric": "photon_histories_per_s", "value": n4 / dt4, "unit": "histories/s",
                  "config": {"workload": WORKLOAD_C4, "sample": "4 photons/pixel of one view (%.3g histories), the reference's single-majorant loop" % n4},
                  "cpu_baseline": {"value": n4 / dt4, "unit": "histories/s", "cores": threads, "kind": "port"},
                  "e2e": {"value": n4 / dt4, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    if args.c1_views > 0:
        c1 = cpu_c1(ob, threads, args.c1_views)
        c1.update(config={"workload": WORKLOAD_C1, "sample": c1.pop("sample")}, cores=threads, kind="port")
        line["c1"] = c1
    shipped = None if args.skip_shipped else as_shipped_reference(ob)
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
    ap.add_argument("--skip-c5", action="store_true", help="skip the nested FDK sweep (BASELINE configs[4])")
    ap.add_argument("--skip-c1", action="store_true", help="skip the nested C1 block (BASELINE configs[0] through the C ABI)")
    ap.add_argument("--skip-parity", action="store_true", help="skip the (untimed) parity checks")
    ap.add_argument("--skip-shipped", action="store_true", help="reference arm: do not run the unmodified reference binaries (as_shipped block)")
    ap.add_argument("--c1-views", type=int, default=360, help="reference arm: views of the C1 run (360 = the stated size; 0 = skip)")
    ap.add_argument("--c4-histories", type=float, default=0.0, help="histories of the C4 run (default: 1e11 at 8 GPUs, else 1.25e10 per GPU... see mc_c4.config)")
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
    host_group = None
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")     # barriers that keep the GPUs idle (rank 0 drives all of them in the e2e legs)
    api.init(local)

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier(group=host_group)

    c = types.SimpleNamespace(args=args, torch=torch, dist=dist, api=api, mdist=mdist, ws=ws, rank=rank, local=local, dev=dev,
                              pk=peaks(), barrier=barrier, host_barrier=host_barrier,
                              flush=torch.empty(256 << 20, dtype=torch.uint8, device=dev),        # > 126 MB L2
                              parity={}, failures=[])

    def rebind_all():                                   # rank 0: the library on all N devices (in-process sharding)
        api.init(list(range(ws)))

    def rebind_own():
        api.init(local)
    c.rebind_all, c.rebind_own = rebind_all, rebind_own

    line = bench_mc_c2(c)
    blocks = (("fdk", bench_fdk_c3, args.skip_fdk), ("mc_c4", bench_mc_c4, args.skip_c4), ("fdk_c5", bench_fdk_c5, args.skip_c5 or args.skip_fdk),
              ("c1", bench_c1, args.skip_c1))
    for name, fn, skip in blocks:
        if skip:
            line[name] = None
            continue
        try:
            line[name] = fn(c)
        except Exception as e:                          # the headline line is still printed
            import traceback
            line[name] = {"error": "%s: %s" % (type(e).__name__, e), "where": traceback.format_exc()[-400:]}
            try:
                rebind_own()
            except Exception:
                pass
    # every rank's parity verdicts -> rank 0
    fails = c.failures
    if ws > 1:
        allf = [None] * ws
        dist.all_gather_object(allf, fails, group=host_group)
        fails = [f for fl in allf for f in fl]
    if rank == 0:
        c.parity["ok"] = not fails
        if fails:
            c.parity["failures"] = fails
        line["parity"] = c.parity
        if args.path == "fdk" and line.get("fdk") and "value" in line["fdk"]:      # FDK as the top-level line, MC nested
            top = dict(line["fdk"])
            mc = {k: line[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "cpu_baseline", "config", "gpu_launches")}
            top.update(n_gpus=ws, higher_is_better=True, vs_baseline=None, data="synthetic", mc=mc, parity=c.parity)
            line = top
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()
    if fails:
        raise SystemExit("bench.py: parity check failed: %s" % fails)


def timed_steps(c, step, K, W):
    """W warm-up steps, then K steps timed one by one with CUDA events on the current stream (L2 flushed before each,
    outside the events); returns the summed milliseconds, max over ranks"""
    torch = c.torch
    for k in range(W):
        step(k)
    c.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        c.flush.fill_(k & 0xFF)
        ev[k][0].record()
        step(W + k)
        ev[k][1].record()
    c.barrier()
    return c.mdist.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), c.dev)


def bench_mc_c2(c):
    torch, dist, api, mdist, args, ws, rank, dev, pk = c.torch, c.dist, c.api, c.mdist, c.args, c.ws, c.rank, c.dev, c.pk
    K, W = args.steps, max(args.warmup, 3)
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
        scene.simulate_dev(a0, a5, per, seed=SEED, views=views, n_range=n_range, d_stats=stats)

    def mc_step(k):
        v = k % g.n_views
        mdist.mc_sharded_step(run_local, im0, im5, per_total, (v, v + 1))

    for k in range(W):
        mc_step(k)
    stats.zero_()
    im0.zero_(); im5.zero_()                            # the timed steps visit views W .. W+K-1 once each (K <= 357)
    c.barrier()
    clocks = ClockSampler(c.local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall = time.perf_counter()
    for k in range(K):
        c.flush.fill_(k & 0xFF)                     # L2 flush, outside the per-step events
        if W + k >= g.n_views:                      # a revisited view starts from zero again (outside the events)
            im0[(W + k) % g.n_views].zero_(); im5[(W + k) % g.n_views].zero_()
        ev[k][0].record()
        mc_step(W + k)
        ev[k][1].record()
    c.barrier()
    t_wall = time.perf_counter() - t_wall
    clk = clocks.stop()
    ms_mc = mdist.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev)
    st = api.unpack_stats(stats.cpu().numpy().astype(np.uint64))
    hist_rank = st["histories"]
    hist_total = npix * PER * ws * K
    assert hist_rank == npix * PER * K, (hist_rank, npix * PER * K)
    mc_value = hist_total / (ms_mc * 1e-3)
    steps_per_hist = st["woodcock_steps"] / hist_rank
    int_per_hist = st["interactions"] / hist_rank
    instr_per_hist = MC_INSTR_PER_STEP * steps_per_hist + MC_INSTR_PER_INTERACTION * int_per_hist
    rate_rank = hist_rank / (ms_mc * 1e-3)
    achieved = rate_rank * instr_per_hist / 1e12            # per GPU, T lane-instr/s
    kc = kernel_counters("mc_transport_kernel")
    mc_roof = {"bound": "fp32-issue", "achieved": achieved, "peak": pk["fp32_tlane_instr"], "unit": "Tlane-instr/s",
               "frac": achieved / pk["fp32_tlane_instr"], "traffic": kc["dram_bytes"] if kc else None,
               "kernel": "mc_transport_kernel_v3<false,5>", "kernel_ms_per_launch": ms_mc / K,
               "instr_per_unit_model": instr_per_hist,
               "model": "%d instr/Woodcock step x %.3f steps/history + %d instr/interaction x %.3f (SURVEY 8d); "
                        "peak = 148 SM x 128 lanes x %.0f MHz (%s)" % (MC_INSTR_PER_STEP, steps_per_hist,
                                                                      MC_INSTR_PER_INTERACTION, int_per_hist,
                                                                      pk["sm_max_mhz"], pk["source"]),
               "hbm": {"achieved_gbs": (2 * npix * 4 + lab.size) * K / (ms_mc * 1e-3) / 1e9, "peak_gbs": pk["hbm_gbs"],
                       "note": "tally flush + one read of the label volume per view: the path is not HBM-bound"}}
    utilisation(mc_roof, kc, rate_rank, pk)

    # ---- parity (untimed): the N-rank sum of one view against the same view computed by rank 0 alone, bit for bit
    if not args.skip_parity:
        v = (W + K) % g.n_views
        im0[v].zero_(); im5[v].zero_()
        mdist.mc_sharded_step(run_local, im0, im5, per_total, (v, v + 1))
        c.barrier()
        if rank == 0:
            r0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32, device=dev)
            r5 = torch.zeros_like(r0)
            scene.simulate_dev(r0, r5, per_total, seed=SEED, views=(v, v + 1), n_range=(0, per_total))
            torch.cuda.synchronize()
            same = bool(torch.equal(r0[v], im0[v]) and torch.equal(r5[v], im5[v]))
            c.parity["mc_n_equals_1"] = same
            c.parity["mc_checked"] = "view %d, %d photons/pixel: %d-rank sum vs rank 0 alone, int32 images bit-equal" % (v, per_total, ws)
            c.parity["mc_image0_sum"] = int(r0[v].sum())
            if not same:
                c.failures.append("mc: %d-rank tallies differ from the one-GPU run" % ws)
            ref_h0, ref_h5 = r0[v].cpu().numpy(), r5[v].cpu().numpy()
            del r0, r5
        c.host_barrier()
    scene.close()
    del im0, im5

    # ---- e2e: host buffers through the C ABI.  N = 1: monte_gpu_simulate on this GPU.  N > 1: the same ONE call on
    # rank 0 with the library bound to all N devices; the other ranks wait at a host barrier with idle GPUs.
    lab_pin = torch.from_numpy(lab).pin_memory()
    h0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32).pin_memory()
    h5 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32).pin_memory()

    def e2e_loop(n):
        t0 = time.perf_counter()
        for k in range(n):
            v = k % g.n_views
            api.simulate(g, vol, lab_pin.numpy(), xs, spec, per_total, seed=SEED, views=(v, v + 1), out=(h0.numpy(), h5.numpy()))
        return time.perf_counter() - t0

    mc_e2e = None
    torch.cuda.empty_cache()
    c.host_barrier()
    if rank == 0:
        if ws > 1:
            c.rebind_all()
        res = {}
        for name, cache in (("strict", "0"), ("cached_labels", "1")):
            os.environ["MONTE_MC_LABEL_CACHE"] = cache
            e2e_loop(2)
            res[name] = e2e_loop(K)
        del os.environ["MONTE_MC_LABEL_CACHE"]
        if not args.skip_parity and ws > 1:             # the in-library N-device call against rank 0 alone, bit for bit
            v = (W + K) % g.n_views
            api.simulate(g, vol, lab_pin.numpy(), xs, spec, per_total, seed=SEED, views=(v, v + 1), out=(h0.numpy(), h5.numpy()))
            same = bool(np.array_equal(h0[v].numpy(), ref_h0) and np.array_equal(h5[v].numpy(), ref_h5))
            c.parity["mc_inlib_equals_1"] = same
            if not same:
                c.failures.append("mc: monte_gpu_simulate on %d devices differs from the one-GPU run" % ws)
        if ws > 1:
            c.rebind_own()
        mc_e2e = {"value": hist_total / res["strict"], "unit": "histories/s",
                  "h2d_bytes_per_step": int(lab.size + ws * (2 * 201 * 16 + 201 * 4 + g.n_views * 8)),
                  "d2h_bytes_per_step": int(2 * npix * 4 + ws * 128), "ms_per_step": 1e3 * res["strict"] / K,
                  "call": "monte_gpu_simulate (C ABI, pinned host buffers)" + (" with the library bound to %d devices on rank 0" % ws if ws > 1 else ""),
                  "note": "the label volume is uploaded on every call (MONTE_MC_LABEL_CACHE=0)" + (": every device takes 1/%d of it from the host and the rest from its peers over NVLink" % ws if ws > 1 else ""),
                  "cached_labels": {"value": hist_total / res["cached_labels"], "ms_per_step": 1e3 * res["cached_labels"] / K,
                                    "h2d_bytes_per_step": int(ws * (2 * 201 * 16 + 201 * 4 + g.n_views * 8)),
                                    "note": "MONTE_MC_LABEL_CACHE=1: the host buffer is hashed (8 threads) and re-uploaded only when its content changed (the default does that only where a clearance grid or a presence scan depends on the labels)"}}
    c.host_barrier()

    # cpu baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample
    mc_cpu = None
    if rank == 0 and ws == 1 and not args.skip_cpu:
        from oracle import binding as ob
        ob.build(ref=False)
        threads = ob.set_threads(os.cpu_count() or 1)
        n, dt, _ = cpu_mc_sample(ob, 400, threads=threads)
        mc_cpu = {"value": n / dt, "unit": "histories/s", "cores": threads, "kind": "port",
                  "sample": "C2 scene, view 0, 400 photons/pixel = %d histories in %.1f s (oracle: double, MT19937, OpenMP over detector rows)" % (n, dt)}
        c.shipped = as_shipped_reference(ob)                 # the unmodified reference programs, one core, as shipped
        if "mc" in c.shipped:
            mc_cpu["as_shipped"] = c.shipped["mc"]
        rc = ref_cuda_record()
        if rc:
            mc_cpu["ref_cuda"] = rc

    return {
        "metric": "photon_histories_per_s", "value": mc_value, "unit": "histories/s",
        "n_gpus": ws, "steps": K, "warmup": W, "ms_per_step": ms_mc / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_C2,
                   "histories_per_step": npix * PER * ws, "photons_per_pixel_per_gpu": PER,
                   "parallelism": "photon-range x%d + 1 NCCL reduce/view" % ws if ws > 1 else "single GPU",
                   "l2": "256 MiB fill between steps (outside the per-step CUDA events); the 34 MB label volume is re-read from HBM each step",
                   "tracking_mode": int(args.mc_tracking),
                   "steps_per_history": steps_per_hist, "interactions_per_history": int_per_hist,
                   "primary_fraction": st["primaries"] / hist_rank, "scatter_detected_fraction": st["scatter_detected"] / hist_rank},
        "e2e": mc_e2e, "gpu_launches": K,
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "samples": clk["samples"]},
        "roofline": mc_roof, "cpu_baseline": mc_cpu,
        "wall_s_timed_region": t_wall,
    }


def ref_cuda_record():
    """the reference's own CUDA kernel on a B200 of this pool (scripts/ref_cuda_run.py, committed under profiles/): a
    recorded figure, not re-measured in this run (the program takes minutes and has no timers)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ref_cuda.json")) as f:
            r = json.load(f)
        return {k: r[k] for k in ("binary", "histories", "wall_s", "whole_program_hist_per_s", "kernel_hist_per_s", "chi2_z",
                                  "ours_hist_per_s_kernel", "speedup_kernel", "chi2_image0", "chi2_dof") if k in r} | \
               {"kind": "reference (CUDA, sm_100)", "source": "profiles/r02_ref_cuda.json (recorded, same GPU model)"}
    except Exception:
        return None


def bench_mc_c4(c):
    """BASELINE configs[3]: full-scatter MC over a 120 kVp polyenergetic spectrum, 1e11 histories split across 8 GPUs,
    one tally reduce per view.  The run really executes the stated number of histories: 1e11 at N = 8; at fewer GPUs
    1.25e10 per GPU (the same per-GPU load), stated in config.  Timed for the reference's single-majorant Woodcock loop
    (a short sample) and for tracking_mode AUTO (-> DIRECTIONAL two-level majorant), which is the value."""
    torch, api, mdist, args, ws, rank, dev, pk = c.torch, c.api, c.mdist, c.args, c.ws, c.rank, c.dev, c.pk
    g, vol, lab = scenes.config_c2()
    xs = scenes.make_xs()
    spec, keep = scenes.kramers_spectrum()
    npix = g.ny * g.nx
    per_total = PER * ws
    hist_step = npix * PER * ws
    target = args.c4_histories if args.c4_histories > 0 else 1.25e10 * ws
    K_full = int(np.ceil(target / hist_step))
    W = 3
    im0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32, device=dev)
    im5 = torch.zeros_like(im0)
    stats = torch.zeros(16, dtype=torch.int64, device=dev)
    out = {"metric": "photon_histories_per_s", "unit": "histories/s", "scaling": "weak", "dtype": "f32", "warmup": W,
           "config": {"workload": WORKLOAD_C4, "histories_total": hist_step * K_full,
                      "histories_note": "1e11 (the stated size) at 8 GPUs; 1.25e10 per GPU at fewer" if args.c4_histories <= 0 else "--c4-histories",
                      "histories_per_step": hist_step,
                      "parallelism": "photon-range x%d + 1 NCCL reduce/view" % ws if ws > 1 else "single GPU",
                      "l2": "256 MiB fill between steps"}}
    for name, mode, K in (("reference_loop", _abi.TRACK_GLOBAL, 6), ("auto", _abi.TRACK_AUTO, K_full)):
        vol.tracking_mode, vol.clearance_cell_log2 = mode, 2
        scene = api.Scene(g, vol, lab, xs, spec)

        def run_local(a0, a5, per, views, n_range):
            scene.simulate_dev(a0, a5, per, seed=SEED, views=views, n_range=n_range, d_stats=stats)

        def step(k):
            v = k % g.n_views
            mdist.mc_sharded_step(run_local, im0, im5, per_total, (v, v + 1))

        for k in range(W):
            step(k)
        stats.zero_()
        c.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t_wall = time.perf_counter()
        for k in range(K):
            if k % 8 == 0:
                c.flush.fill_(k & 0xFF)              # (every 8th step: the label volume is re-read from HBM; 1e11 histories are ~1000 steps)
            ev[k][0].record()
            step(W + k)
            ev[k][1].record()
        c.barrier()
        t_wall = time.perf_counter() - t_wall
        ms = mdist.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev)
        st = api.unpack_stats(stats.cpu().numpy().astype(np.uint64))
        scene.close()
        hist = hist_step * K
        rate_rank = st["histories"] / (ms * 1e-3)
        sph, iph = st["woodcock_steps"] / st["histories"], st["interactions"] / st["histories"]
        out[name] = {"value": hist / (ms * 1e-3), "ms_per_step": ms / K, "steps": K, "histories": hist, "seconds_event_time": ms * 1e-3,
                     "seconds_wall": t_wall, "steps_per_history": sph, "interactions_per_history": iph,
                     "scatter_detected_fraction": st["scatter_detected"] / st["histories"]}
        if name == "auto":
            instr = MC_INSTR_PER_STEP * sph + MC_INSTR_PER_INTERACTION * iph
            out["roofline"] = {"bound": "fp32-issue", "achieved": rate_rank * instr / 1e12, "peak": pk["fp32_tlane_instr"], "unit": "Tlane-instr/s",
                               "frac": rate_rank * instr / 1e12 / pk["fp32_tlane_instr"], "traffic": None,
                               "kernel": "mc_transport_kernel_v3<false,5,3,2,false,CLEAR=true>", "kernel_ms_per_launch": ms / K,
                               "instr_per_unit_model": instr}
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_AUTO, 2
    mode, cell, ratio = api.resolve_tracking(xs, spec)
    out["value"] = out["auto"]["value"]
    out["ms_per_step"] = out["auto"]["ms_per_step"]
    out["steps"] = out["auto"]["steps"]
    out["seconds_for_the_run"] = out["auto"]["seconds_event_time"]
    out["tracking"] = "tracking_mode AUTO -> mode %d with %d-voxel clearance cells (mean majorant ratio %.2f); reference_loop = single majorant, 6-step sample" % (mode, 1 << cell, ratio)
    # e2e: the C-ABI host-buffer call (rank 0, all devices in-process at N > 1), 12 views
    del im0, im5
    torch.cuda.empty_cache()
    c.host_barrier()
    if rank == 0:
        if ws > 1:
            c.rebind_all()
        lab_pin = torch.from_numpy(lab).pin_memory()
        h0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32).pin_memory()
        h5 = torch.zeros_like(h0).pin_memory()
        os.environ["MONTE_MC_LABEL_CACHE"] = "0"
        Ke = 12
        for k in range(2 + Ke):
            if k == 2:
                t0 = time.perf_counter()
            api.simulate(g, vol, lab_pin.numpy(), xs, spec, per_total, seed=SEED, views=(k, k + 1), out=(h0.numpy(), h5.numpy()))
        te = time.perf_counter() - t0
        del os.environ["MONTE_MC_LABEL_CACHE"]
        out["e2e"] = {"value": hist_step * Ke / te, "unit": "histories/s", "ms_per_step": 1e3 * te / Ke, "steps": Ke,
                      "h2d_bytes_per_step": int(lab.size + ws * (2 * 201 * 16 + 481 * 4)), "d2h_bytes_per_step": int(2 * npix * 4 + ws * 128),
                      "note": "labels uploaded on every call; the clearance grids are rebuilt only when the labels change (content hash)"}
        if ws > 1:
            c.rebind_own()
    c.host_barrier()
    return out


def fdk_onboard_fraction(g, z_slices, n=2_000_000, seed=7):
    """fraction of the (voxel, view) pairs of the given slices that project onto the detector (the others are skipped, as in
    bp3d20.cpp:116): Monte-Carlo estimate with the reference's projection formulas"""
    rs = np.random.default_rng(seed)
    sx = g.x0 + g.vox * rs.integers(0, g.nx, n)
    sy = g.y0 - g.vox * rs.integers(0, g.ny, n)
    sz = g.z0 - g.vox * rs.choice(z_slices, n)
    beta = np.deg2rad(g.angle0_deg + g.angle_step_deg * rs.integers(0, g.n_views, n))
    kk = g.dsd / (sx * np.cos(beta) + sy * np.sin(beta) + g.dso)
    return float(np.mean((np.abs(kk * (-sx * np.sin(beta) + sy * np.cos(beta))) <= g.half_u) & (np.abs(kk * sz) <= g.half_v)))


def fdk_sharded_setup(c, g, proj):
    """buffers and the sharded step (filter own views | one all_to_all of the detector-row bands | backproject own z ranges)
    of the one-process-per-GPU FDK; returns (step(src), filt, slab, my_z, z_off, z_ranges, n_my)"""
    torch, api, mdist, ws, rank, dev = c.torch, c.api, c.mdist, c.ws, c.rank, c.dev
    z_ranges = mdist.fdk_z_partition(g, ws)
    my_z = z_ranges[rank]
    n_my = sum(b - a for a, b in my_z)
    filt = torch.zeros(api.fdk_filtered_shape(g), device=dev)
    slab = torch.empty((max(n_my, 1), g.ny, g.nx), device=dev)      # this rank's ranges, stacked in ascending z
    z_off, o = {}, 0
    for a, b in my_z:
        z_off[a] = o
        o += b - a
    exchange = os.environ.get("MONTE_BENCH_FDK_EXCHANGE", "peers" if ws > 1 else "band")
    peers = None
    for old in getattr(c, "peer_rows", []):                      # (an earlier set-up of this run: unmap before mapping anew)
        c.barrier()
        old.close()
    c.peer_rows = []
    if exchange == "peers":
        try:
            peers = mdist.PeerRows(api, filt, g.n_views)
        except Exception as ex:                                   # (no CUDA IPC between these processes: the all_to_all form)
            sys.stderr.write("bench: CUDA IPC set-up failed (%s); using the band all_to_all\n" % ex)
            exchange = "band"
        ok = torch.tensor([1 if exchange == "peers" else 0], device=dev)
        c.dist.all_reduce(ok, op=c.dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange = "band"
            if peers is not None:
                peers.close()
                peers = None
        if peers is not None:
            c.peer_rows.append(peers)

    def sharded(src):
        if exchange == "peers":
            mdist.fdk_sharded_peers(lambda a, b: api.fdk_filter_dev(g, src, filt, a, b, pad=False),
                                    lambda z0, z1, ptrs, ends: api.fdk_backproject_peers_dev(
                                        g, ptrs, ends, slab[z_off[z0]:z_off[z0] + z1 - z0], z0, z1),
                                    peers, g.n_views, z_ranges)
        elif exchange == "pipelined":
            mdist.fdk_sharded_pipelined(lambda a, b: api.fdk_filter_dev(g, src, filt, a, b, pad=False),
                                        lambda a, b: api.fdk_pad_views_dev(g, filt, a, b),
                                        lambda z0, z1, a, b, cont: api.fdk_backproject_views_dev(
                                            g, filt, slab[z_off[z0]:z_off[z0] + z1 - z0], z0, z1, a, b, cont),
                                        filt, g.n_views, g.nv, g.nz, z_ranges=z_ranges)
        else:
            mdist.fdk_sharded_band(lambda a, b: api.fdk_filter_dev(g, src, filt, a, b, pad=False),
                                   lambda: api.fdk_pad_dev(g, filt),
                                   lambda z0, z1: api.fdk_backproject_dev(g, filt, slab[z_off[z0]:z_off[z0] + z1 - z0], z0, z1),
                                   lambda z0, z1: api.fdk_slab_rows(g, z0, z1), filt, g.n_views, g.nv, z_ranges)
    sharded.peers = peers
    return sharded, filt, slab, my_z, z_off, z_ranges, n_my, exchange


def bench_fdk_c3(c):
    torch, dist, api, mdist, args, ws, rank, dev, pk = c.torch, c.dist, c.api, c.mdist, c.args, c.ws, c.rank, c.dev, c.pk
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    Kf, Wf = max(args.fdk_steps, 1), 3
    v_lo, v_hi = split_range(g.n_views, ws, rank)
    gen = torch.Generator(device=dev).manual_seed(1234)
    proj = torch.rand((g.n_views, g.nu, g.nv), device=dev, generator=gen)     # same on every rank
    sharded, filt, slab, my_z, z_off, z_ranges, n_my, exchange = fdk_sharded_setup(c, g, proj)

    for _ in range(Wf):
        sharded(proj)
    c.barrier()
    clocks = ClockSampler(dev.index)
    clocks.start()
    tot = 0.0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(Kf):
        c.flush.fill_(1)
        e[0].record()
        sharded(proj)
        e[1].record()
        torch.cuda.synchronize()
        tot += e[0].elapsed_time(e[1])
    # split of one step into filter / backprojection (this rank's kernels alone, no exchange)
    e[0].record()
    api.fdk_filter_dev(g, proj, filt, v_lo, v_hi, pad=False)
    e[1].record()
    torch.cuda.synchronize()
    t_filter = e[0].elapsed_time(e[1])
    c.barrier()
    sharded(proj)                                   # (restores every row this rank's slabs read)
    e[1].record()
    for a, b in my_z:
        if sharded.peers is not None:               # (the gather from the peers' rows is part of this rank's backprojection)
            api.fdk_backproject_peers_dev(g, sharded.peers.ptrs, sharded.peers.v_end, slab[z_off[a]:z_off[a] + b - a], a, b)
        else:
            api.fdk_backproject_dev(g, filt, slab[z_off[a]:z_off[a] + b - a], a, b)
    e[2].record()
    torch.cuda.synchronize()
    t_bp = e[1].elapsed_time(e[2])
    c.barrier()
    clk = clocks.stop()
    tot = mdist.max_over_ranks(tot, dev)
    t_filter_max, t_bp_max = mdist.max_over_ranks(t_filter, dev), mdist.max_over_ranks(t_bp, dev)
    updates = g.nx * g.ny * g.nz * g.n_views
    gups = updates * Kf / (tot * 1e-3) / 1e9
    upd_rank = g.nx * g.ny * n_my * g.n_views
    my_slices = np.concatenate([np.arange(a, b) for a, b in my_z]) if n_my else np.arange(1)
    inside = fdk_onboard_fraction(g, my_slices)
    rate_on = inside * upd_rank / (t_bp * 1e-3)             # on-detector updates per second, this rank's backprojector
    achieved = rate_on * FDK_INSTR_PER_UPDATE / 1e12
    kc = kernel_counters("fdk_backproject_kernel")
    roof = {"bound": "fp32-issue", "achieved": achieved, "peak": pk["fp32_tlane_instr"], "unit": "Tlane-instr/s",
            "frac": achieved / pk["fp32_tlane_instr"],
            "traffic": kc["dram_bytes_per_reconstruction"] if kc and ws == 1 else None,
            "kernel": "fdk_backproject_kernel", "kernel_ms_per_launch": t_bp, "filter_ms_per_launch": t_filter,
            "instr_per_unit_model": FDK_INSTR_PER_UPDATE,
            "traffic_note": "achieved, kernel_ms_per_launch and traffic cover one whole backprojection (the library splits it into L2-sized "
                            "view-chunk launches; traffic = sum of the ncu DRAM bytes of all of them at C3 on one GPU)",
            "model": "%d lane-instr per voxel-update (SURVEY 8d) x %.4g updates per reconstruction x %.3f of them on the detector "
                     "(off-detector pairs are skipped per column, as the reference skips them per voxel).  The kernel executes fewer "
                     "instructions per update than the model charges, so this model fraction can pass 1: issue_util is the utilisation" % (
                         FDK_INSTR_PER_UPDATE, upd_rank, inside),
            "on_detector_fraction": inside,
            "hbm": {"algorithmic_bytes": 4 * g.nx * g.ny * n_my + 4 * g.n_views * g.nu * g.nv,
                    "achieved_gbs": (4 * g.nx * g.ny * n_my + 4 * g.n_views * g.nu * g.nv) / (t_bp * 1e-3) / 1e9,
                    "peak_gbs": pk["hbm_gbs"], "note": "projections are read once from HBM and re-read from L1/L2; not HBM-bound"}}
    utilisation(roof, kc, rate_on, pk)

    # ---- parity (untimed): this rank's slab of the sharded run against (a) the same slab reconstructed by this rank alone
    # from all views it filtered itself, bit for bit; (b) the CPU oracle on 8 probe voxels per rank, 1e-4 of the maximum
    if not args.skip_parity and n_my:
        sharded(proj)
        torch.cuda.synchronize()
        got = slab[:n_my].clone()
        api.fdk_filter_dev(g, proj, filt, 0, g.n_views, pad=True)
        for a, b in my_z:
            api.fdk_backproject_dev(g, filt, slab[z_off[a]:z_off[a] + b - a], a, b)
        torch.cuda.synchronize()
        same = bool(torch.equal(got, slab[:n_my]))
        dense = torch.empty((g.n_views, g.nv, g.nu), dtype=torch.float32, device=dev)
        api.fdk_unpad_dev(g, filt, dense)
        torch.cuda.synchronize()
        dense_h = dense.cpu().numpy()
        del dense
        scale = float(got.abs().max())
        rel = 0.0
        from oracle import binding as ob
        ob.build(ref=False)
        ob.set_threads(4)
        rs = np.random.default_rng(100 + rank)
        for _ in range(8):
            z = int(rs.choice(my_slices)); t = int(rs.integers(0, g.ny)); s = int(rs.integers(0, g.nx))
            gp = g.copy()
            gp.s_begin, gp.s_end, gp.t_begin, gp.t_end, gp.z_begin, gp.z_end = s, s + 1, t, t + 1, z, z + 1
            ref = float(ob.fdk_backproject(gp, dense_h)[z, t, s])
            zi = next(z_off[a] + z - a for a, b in my_z if a <= z < b)
            rel = max(rel, abs(float(got[zi, t, s]) - ref) / scale)
        del dense_h, got
        if not same:
            c.failures.append("fdk: rank %d's slab of the sharded run differs from its one-GPU reconstruction" % rank)
        if rel > 1e-4:
            c.failures.append("fdk: rank %d's slab differs from the oracle by %.3g of the maximum" % (rank, rel))
        rel_max = mdist.max_over_ranks(rel, dev)
        same_all = mdist.max_over_ranks(0.0 if same else 1.0, dev) == 0.0
        c.parity.update(fdk_slab_bit_equal=same_all, fdk_slab_max_rel=rel_max,
                        fdk_checked="every rank: its z ranges, sharded run vs the rank alone (bit-equal) and 8 probe voxels vs the CPU oracle (<= 1e-4 of max)")
    elif not args.skip_parity:
        mdist.max_over_ranks(0.0, dev); mdist.max_over_ranks(0.0, dev)

    # ---- e2e through the C ABI with pinned host buffers: monte_gpu_fdk on this GPU (N = 1) or on rank 0 with the library
    # bound to all N devices
    del proj, filt, slab
    torch.cuda.empty_cache()
    c.host_barrier()
    e2e = None
    if rank == 0:
        if ws > 1:
            c.rebind_all()
        host_proj = torch.rand((g.n_views, g.nu, g.nv)).pin_memory()
        host_vol = torch.empty((g.nz, g.ny, g.nx)).pin_memory()
        _, _, _, stf = api.fdk(g, host_proj.numpy(), want_filtered=False, out=host_vol.numpy())
        t0 = time.perf_counter()
        for _ in range(Kf):
            _, _, _, stf = api.fdk(g, host_proj.numpy(), want_filtered=False, out=host_vol.numpy())
        t_e2e = time.perf_counter() - t0
        e2e = {"value": updates * Kf / t_e2e / 1e9, "unit": "GUPS", "ms_per_step": 1e3 * t_e2e / Kf,
               "h2d_bytes_per_step": int(4 * g.n_views * g.nu * g.nv), "d2h_bytes_per_step": int(4 * g.nx * g.ny * g.nz),
               "call": "monte_gpu_fdk (C ABI, pinned host buffers)" + (" with the library bound to %d devices on rank 0" % ws if ws > 1 else ""),
               "breakdown_ms": {"upload+filter": stf["ms_filter"],
                                ("backproject_after_the_last_filter" if ws > 1 else "gather+backproject"): stf["ms_backproject"],
                                "not_hidden_d2h_and_sync": stf["ms_d2h"], "total": stf["ms_total"]},
               "limiter": ("host links: %.2f GB in + %.2f GB out in %.1f ms = %.0f GB/s through one host's PCIe / memory system; the backprojection "
                           "runs underneath the uploads (views are dealt to the devices in interleaved chunks)" %
                           (4e-9 * g.n_views * g.nu * g.nv, 4e-9 * g.nx * g.ny * g.nz, 1e3 * t_e2e / Kf,
                            (4e-9 * g.n_views * g.nu * g.nv + 4e-9 * g.nx * g.ny * g.nz) / (t_e2e / Kf))) if ws > 1 else
                          "backprojection (the uploads run underneath it)"}
        if not args.skip_parity and ws > 1:             # the N-device call against the one-device call on the same host buffers
            got = host_vol.numpy().copy()
            c.rebind_own()
            api.fdk(g, host_proj.numpy(), want_filtered=False, out=host_vol.numpy())
            same = bool(np.array_equal(got, host_vol.numpy()))
            c.parity["fdk_inlib_equals_1"] = same
            if not same:
                c.failures.append("fdk: monte_gpu_fdk on %d devices differs from the one-GPU call" % ws)
            del got
        elif ws > 1:
            c.rebind_own()
        del host_proj, host_vol
    c.host_barrier()
    cpu = None
    if rank == 0 and ws == 1 and not args.skip_cpu:
        from oracle import binding as ob
        threads = ob.set_threads(os.cpu_count() or 1)
        upd, dt = cpu_fdk_sample(ob, 4)
        cpu = {"value": upd / dt / 1e9, "unit": "GUPS", "cores": threads, "kind": "port",
               "sample": "C3 geometry, backprojection of 4 central z-slices from 720 views = %.3g updates in %.1f s (oracle: double, OpenMP)" % (upd, dt)}
        if getattr(c, "shipped", None) and "fdk" in c.shipped:
            cpu["as_shipped"] = c.shipped["fdk"]
    limiter = max((("backprojection", t_bp_max), ("filter", t_filter_max), ("exchange+launch gaps", tot / Kf - t_bp_max - t_filter_max)), key=lambda q: q[1])
    return {"metric": "fdk_voxel_updates_per_s", "value": gups, "unit": "GUPS", "steps": Kf, "warmup": Wf,
            "ms_per_step": tot / Kf, "scaling": "strong", "dtype": "f32",
            "config": {"workload": WORKLOAD_C3,
                       "parallelism": "z-ranges of equal work x%d %s, filter by views, %s" % (
                           ws, [[list(z) for z in zr] for zr in z_ranges],
                           "view pieces broadcast in order and overlapped with the backprojection" if exchange == "pipelined"
                           else "no collective: the backprojector's pair conversion loads the detector-row band each slab reads out of the "
                                "peers' filtered rows (CUDA IPC, NVLink peer loads), two one-element all-reduces order the ranks" if exchange == "peers"
                           else "one all_to_all of the detector-row bands each slab reads") if ws > 1 else "single GPU",
                       "l2": "256 MiB fill between steps; projections (2.26 GB) exceed L2"},
            "breakdown_ms": {"filter_max_over_ranks": t_filter_max, "backproject_max_over_ranks": t_bp_max,
                             "exchange_and_gaps": tot / Kf - t_bp_max - t_filter_max, "step": tot / Kf, "limiter": limiter[0]},
            "e2e": e2e, "gpu_launches": 3 * Kf, "roofline": roof, "cpu_baseline": cpu,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "seconds_for_512cube_720views": tot / Kf * 1e-3}


def bench_fdk_c5(c):
    """BASELINE configs[4]: three points of the backprojection sweep, resident, sharded like C3 at N > 1"""
    torch, api, mdist, ws, rank, dev, pk = c.torch, c.api, c.mdist, c.ws, c.rank, c.dev, c.pk
    pts = []
    for n, views in ((256, 360), (512, 720), (1024, 1440)):
        det = 3 * n // 2
        g = _abi.generic_fdk_geom(views, det, det, n)
        gen = torch.Generator(device=dev).manual_seed(n)
        proj = torch.rand((views, det, det), device=dev, generator=gen)
        sharded, filt, slab, my_z, z_off, z_ranges, n_my, _ = fdk_sharded_setup(c, g, proj)
        K = 3 if n < 1024 else 2
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        sharded(proj)
        c.barrier()
        tot = 0.0
        for _ in range(K):
            c.flush.fill_(2)
            e[0].record()
            sharded(proj)
            e[1].record()
            torch.cuda.synchronize()
            tot += e[0].elapsed_time(e[1])
        e[0].record()
        for a, b in my_z:
            api.fdk_backproject_dev(g, filt, slab[z_off[a]:z_off[a] + b - a], a, b)
        e[1].record()
        torch.cuda.synchronize()
        t_bp = mdist.max_over_ranks(e[0].elapsed_time(e[1]), dev)
        tot = mdist.max_over_ranks(tot, dev)
        upd = float(n) ** 3 * views
        rate_on = float(g.nx * g.ny * n_my * views) / (e[0].elapsed_time(e[1]) * 1e-3) if n_my else 0.0
        pts.append({"volume": n, "views": views, "detector": det, "gups": upd * K / (tot * 1e-3) / 1e9, "ms_per_reconstruction": tot / K,
                    "backproject_ms_max_over_ranks": t_bp, "gups_backprojection_only": upd / (t_bp * 1e-3) / 1e9,
                    "roofline_frac_model": rate_on * FDK_INSTR_PER_UPDATE / 1e12 / pk["fp32_tlane_instr"]})
        del proj, filt, slab
        torch.cuda.empty_cache()
    big = pts[-1]
    return {"metric": "fdk_voxel_updates_per_s", "unit": "GUPS", "value": big["gups"], "scaling": "strong", "dtype": "f32",
            "config": {"workload": WORKLOAD_C5, "points": "256^3 x 360, 512^3 x 720, 1024^3 x 1440; value = the largest point",
                       "l2": "256 MiB fill between reconstructions"},
            "points": pts,
            "roofline": {"bound": "fp32-issue", "frac": big["roofline_frac_model"], "peak": pk["fp32_tlane_instr"], "unit": "Tlane-instr/s",
                         "achieved": big["roofline_frac_model"] * pk["fp32_tlane_instr"], "traffic": None, "kernel": "fdk_backproject_kernel",
                         "instr_per_unit_model": FDK_INSTR_PER_UPDATE},
            "e2e": None, "e2e_note": "resident sweep (13.6 GB of projections at the largest point); the host-buffer path is measured at C3"}


def bench_c1(c):
    """BASELINE configs[0] — the reference's own CPU-sized case — through the C ABI host-buffer calls on ONE GPU (rank 0):
    monte_gpu_simulate_maps over all 360 views in one call, then monte_gpu_fdk with the bp3d20 geometry over the whole 256^3."""
    api, torch, rank = c.api, c.torch, c.rank
    out = None
    c.host_barrier()
    if rank == 0:
        g, vol, lab = scenes.config_c1()
        xs, spec = scenes.make_xs(), scenes.mono_spectrum(140.0)
        per = 237
        api.simulate(g, vol, lab, xs, spec, per, seed=SEED, views=(0, 2))
        t0 = time.perf_counter()
        im0, im5, st, m0, m5 = api.simulate(g, vol, lab, xs, spec, per, seed=SEED, maps=True)
        t_mc = time.perf_counter() - t0
        gf = _abi.bp3d20_geom().full_roi()
        gf.mask_r2 = 118 * 118
        api.fdk(gf, m0.transpose(0, 1, 2).copy(), want_filtered=False)
        t0 = time.perf_counter()
        _, v, _, stf = api.fdk(gf, m0, want_filtered=False)
        t_fdk = time.perf_counter() - t0
        upd = 256 ** 3 * 360
        out = {"config": {"workload": WORKLOAD_C1, "parallelism": "single GPU (rank 0)"},
               "mc": {"metric": "photon_histories_per_s", "value": st["histories"] / t_mc, "unit": "histories/s", "histories": st["histories"],
                      "seconds": t_mc, "kernel_ms": st["ms_kernel"],
                      "e2e": {"value": st["histories"] / t_mc, "unit": "histories/s", "h2d_bytes_per_step": int(lab.size), "d2h_bytes_per_step": int(4 * im0.size * 4)}},
               "fdk": {"metric": "fdk_voxel_updates_per_s", "value": upd / t_fdk / 1e9, "unit": "GUPS", "seconds": t_fdk, "kernel_ms": stf["ms_total"],
                       "e2e": {"value": upd / t_fdk / 1e9, "unit": "GUPS", "h2d_bytes_per_step": int(m0.size * 4), "d2h_bytes_per_step": int(v.size * 4)}},
               "note": "wall clock around the two C-ABI calls (pageable numpy buffers); the MC maps of image0 feed the reconstruction"}
    c.host_barrier()
    return out


if __name__ == "__main__":
    main()
