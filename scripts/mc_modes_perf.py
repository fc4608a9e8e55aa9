"""Kernel time of one C2 view (325^3 labels, 325x325 detector, 947 photons/pixel = 1.0003e8 histories) through
the host-buffer C ABI call, per transport mode: the reference's forward coherent event (default), the
Rayleigh form-factor deflection (coherent_mode 1), and the two-level majorant (tracking_mode CLEARANCE) for
several cell sizes.  No torch: monte_mc_stats.ms_kernel is the CUDA-event time
of the transport launch inside monte_gpu_simulate."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api, scenes  # noqa: E402


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 947
    api.init(0)
    g, vol, lab = scenes.config_c2()
    xs = scenes.add_formfactors(scenes.make_xs())
    poly, keep = scenes.kramers_spectrum()
    for name, spec in (("mono140", scenes.mono_spectrum(140.0)), ("kramers120", poly)):
        for mode in (_abi.COHERENT_FORWARD, _abi.COHERENT_FORMFACTOR):
            g.coherent_mode = mode
            best, st = 1e30, None
            for it in range(3):
                _, _, st = api.simulate(g, vol, lab, xs, spec, per, seed=it + 1, views=(it, it + 1))
                if it:
                    best = min(best, st["ms_kernel"])
            n = st["histories"]
            print(json.dumps({"spectrum": name, "coherent_mode": mode, "histories": n, "ms_kernel": best,
                              "hist_per_s": n / best * 1e3, "steps_per_hist": st["woodcock_steps"] / n,
                              "coherent_per_hist": st["coherent"] / n,
                              "scatter_det_frac": st["scatter_detected"] / n}), flush=True)
    g.coherent_mode = _abi.COHERENT_FORWARD
    for name, spec in (("kramers120", poly), ("mono140", scenes.mono_spectrum(140.0)), ("mono60", scenes.mono_spectrum(60.0))):
        for cl in ((2, 3, 4) if name == "kramers120" else (3,)):
            vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, cl
            best, st = 1e30, None
            for it in range(3):
                _, _, st = api.simulate(g, vol, lab, xs, spec, per, seed=it + 1, views=(it, it + 1))
                if it:
                    best = min(best, st["ms_kernel"])
            n = st["histories"]
            print(json.dumps({"spectrum": name, "tracking": "clearance", "cell_voxels": 1 << cl, "histories": n,
                              "ms_kernel": best, "hist_per_s": n / best * 1e3, "steps_per_hist": st["woodcock_steps"] / n,
                              "ms_h2d_incl_grid_build": st["ms_h2d"], "scatter_det_frac": st["scatter_detected"] / n}), flush=True)
        if name != "kramers120":
            vol.tracking_mode = _abi.TRACK_GLOBAL
            _, _, st = api.simulate(g, vol, lab, xs, spec, per, seed=2, views=(1, 2))
            print(json.dumps({"spectrum": name, "tracking": "global", "ms_kernel": st["ms_kernel"],
                              "steps_per_hist": st["woodcock_steps"] / st["histories"]}), flush=True)
    api.shutdown()


if __name__ == "__main__":
    main()
