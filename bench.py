#!/usr/bin/env python
"""bench.py -- hex8 element-steps/s of the explicit step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, resident loop)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path on the host cores

A "step" is one explicit time step (Benchmarking-Parallel.cpp:106-171: both kicks, drift, BC,
GetForce, CalculateAccelerations, CheckEnergy, StableTimeStep) over the whole synthetic mesh.
Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement" for every definition.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SOFT = [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0]  # examples/Benchmarking-Parallel/materials.dat:1
BRAIN = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]  # examples/ex5/materials.dat:2
HGO = BRAIN[:4] + [10.0, 0, 0, 0, 0]
MATERIALS = {1: SOFT, 4: HGO, 5: BRAIN}
MAT_NAME = {1: "compressible neo-Hookean (mat 1)", 4: "HGO isotropic (mat 4)", 5: "HGO + 2-term Prony (mat 5)"}

# Algorithmic work per element-step (DESIGN.md "Work model"): bytes = compulsory HBM traffic of the
# design, flops = fp64 operations of the mode-basis formulation actually executed (FMA = 2).
def algorithmic_bytes(n, mat, energy, rho_n=None):
    if rho_n is None:
        rho_n = ((n + 1.0) / n) ** 3
    elem = 36 + 1 + 48 * rho_n + 192            # K_elem: conn+pid+eflag, X and u (unique nodes), f_e write
    node = 192 + (4 + 32) * rho_n + (8 + 2 + 72 + 72) * rho_n  # K_node: f_e read, CSR, m, flags, u v a read + write
    if energy:
        node += 48 * rho_n                       # fi: write + read next step (the displacement increment is rebuilt, not stored)
    hist = 2304 if mat == 5 else 0
    return elem + hist, node


def hbm_peak_gbs():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling recipe's fallback."""
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def step_roofline(value_total, n_gpus, mat, energy, rho_n):
    """`roofline` of a multi-GPU line: the STEP of one GPU against the HBM roof (the binding one since the current-Jacobian
    element kernel): algorithmic bytes per element-step x this GPU's element-steps/s."""
    b_elem, b_node = algorithmic_bytes(0, mat, energy, rho_n=rho_n)
    peak, src = hbm_peak_gbs()
    gbs = (b_elem + b_node) * (value_total / n_gpus) / 1e9
    return {"kernel": "whole step of one GPU (element kernel + node kernel + exchange)", "bound": "hbm", "achieved": gbs, "peak": peak,
            "unit": "GB/s", "frac": gbs / peak, "traffic": None, "peak_source": "hbm: " + src,
            "algorithmic_bytes_per_element": b_elem + b_node,
            "step_frac_of_hbm_roof_survey_bytes": 717.0 * (value_total / n_gpus) / 1e9 / peak,
            "note": "per-kernel fp64 / HBM views are in the N = 1 line; ncu cannot attach to a multi-process run"}


# (n, material, energy, injury) -> measured DRAM bytes per k_elem launch (ncu, see profiles/)
# key: (n, material, energy, injury, affine kernel)
NCU_TRAFFIC_BYTES = {(100, 1, True, False, False): 86507008 + 138745344,   # profiles/r01_k_elem_general_ncu_full.csv
                     (100, 1, True, False, True): 86602752 + 136579840,    # profiles/r01_k_elem_affine_ncu_full.csv
                     (100, 1, True, False, "cj"): 86508544 + 133630464}    # profiles/r02_k_elem_affine_cj_ncu_full.csv
# executed fp64 flops per element in K_elem (FMA = 2).  Material 1 from ncu (general kernel: 1208 DFMA + 463 DADD + 507 DMUL
# per element, profiles/r01_k_elem_general_ncu_full.csv); materials 4 and 5 = material 1 + the SASS difference of their
# material code
# Round 2 check (profiles/r02_k_elem_affine_ncu_full.csv, ftb_ln_rcp in the material code): 1082 DFMA + 320 DADD + 373 DMUL
# per element = 2857 flops -- the same work in 14 % fewer instructions (88.4 M instead of 102.4 M warp instructions)
ELEM_FLOPS = {1: 3386.0, 4: 4296.0, 5: 6446.0}
# k_elem_affine (parallelepiped reference geometry): no cofactor / determinant / reciprocal of J0 per Gauss point, F in 27
# FMAs, no coordinate modes and columns: 1060 DFMA + 357 DADD + 378 DMUL per element for material 1
# (profiles/r01_k_elem_affine_ncu_full.csv), i.e. 531 flops less than the general kernel
ELEM_FLOPS_AFFINE = {k: v - 531.0 for k, v in ELEM_FLOPS.items()}
# k_elem_affine_cj (round 2, current-Jacobian form of the parallelepiped element; the default for
# materials 1 and 4 without the strain outputs): material 1 from ncu -- 689 DFMA + 248 DADD + 251 DMUL per element
# (profiles/r02_k_elem_affine_cj_ncu_full.csv) = 1876 flops; material 4 = k_elem_affine<4> minus the SASS difference of
# the Gauss loop (47 DFMA + 4 DMUL fewer, 1 DADD more per point) plus the per-element M
ELEM_FLOPS_CJ = {1: 1876.0, 4: 3030.0}


def _nvml_sampler(stop, out, device_index):
    """Fast path: NVML from this process (a sample every ~2 ms, so the 25-60 ms timed regions are covered by several
    samples).  Returns False if NVML is not usable -- the caller then falls back to polling nvidia-smi."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the CUDA device
        try:
            import torch
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id
            dom = torch.cuda.get_device_properties(device_index).pci_domain_id
            dev = torch.cuda.get_device_properties(device_index).pci_device_id
            h = nv.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (dom, bus, dev)).encode())
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(device_index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))  # probe once before committing to this path
    except Exception:
        return False
    while not stop.is_set():
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            out.append([str(sm), str(mx)] + [("Active" if r & b else "Not Active") for b in bits.values()])
        except Exception:
            pass
        stop.wait(0.002)
    return True


def clocks_sampler(stop, out, device_index):
    try:
        if _nvml_sampler(stop, out, device_index):
            return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            out.append([x.strip() for x in r.stdout.strip().split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        if len(s) < 6:
            continue
        try:
            sm.append(float(s[0]))
            mx.append(float(s[1]))
        except ValueError:
            continue
        for nme, v in zip(names, s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(nme)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU side
def run_reference_cpu(n, mat, steps, warmup, ranks):
    """Times the reference's own explicit loop (oracle/_ref/ref_dump_fast = unmodified reference sources,
    built by oracle/ref/build_ref.sh) on a synthetic n^3 cube with `ranks` ranks through the ftmpi shim.
    Falls back to the C port (oracle/) when the reference build is absent."""
    from femtech_b200 import mesh
    refbin = os.path.join(ROOT, "oracle", "_ref", "ref_dump_fast")
    launcher = os.path.join(ROOT, "oracle", "_ref", "ftmpirun")
    X, conn, pid = mesh.cube_mesh(n)
    E = conn.shape[0]
    if os.path.exists(refbin) and os.path.exists(launcher):
        work = tempfile.mkdtemp(prefix="ftbench_")
        mesh.write_abaqus_inp(os.path.join(work, "cube.inp"), X, conn, pid)
        mesh.write_materials_dat(os.path.join(work, "materials.dat"), [mat], MATERIALS[mat])
        tMax = 0.1 if mat == 1 else 0.004
        cmd = [refbin, "cube.inp", "out", str(steps + warmup), repr(tMax), "0.007", "0.005", str(warmup), "nodump"]
        if ranks > 1:
            cmd = [launcher, "-np", str(ranks)] + cmd
        r = subprocess.run(cmd, cwd=work, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("REF_DUMP")]
        if r.returncode == 0 and line:
            kv = dict(t.split("=", 1) for t in line[0].split()[1:] if "=" in t)
            timed = int(kv["timed_steps"])
            loop_s = float(kv["loop_s"])
            return {"value": E * timed / loop_s, "unit": "element-steps/s", "cores": ranks, "kind": "reference",
                    "sample": "%d^3 hex8 cube (%d elements), %s, %d timed steps after %d warm-up, %d rank(s) via the "
                              "in-repo MPI shim; loop only (setup excluded)" % (n, E, MAT_NAME[mat], timed, warmup, ranks),
                    "loop_s": loop_s, "steps": timed}
        sys.stderr.write("reference run failed (%d): %s\n" % (r.returncode, r.stderr[-500:]))
    # port: the C restatement, one core
    from oracle import pyoracle as po
    kind, rate = mesh.benchmark_bc(X)
    o = po.OracleModel(X, conn, pid, [mat], MATERIALS[mat], fast=True)
    o.ShapeFunctions()
    o.AssembleLumpedMass()
    po.run_explicit([o], [kind], rate, 1e9, warmup, record=False)
    t0 = time.time()
    k, _, _ = po.run_explicit([o], [kind], rate, 1e9, steps, first_call=False, record=False)
    loop_s = time.time() - t0
    return {"value": E * k / loop_s, "unit": "element-steps/s", "cores": 1, "kind": "port",
            "sample": "%d^3 hex8 cube (%d elements), %s, %d timed steps, C port of the reference (oracle/), 1 core"
                      % (n, E, MAT_NAME[mat], k), "loop_s": loop_s, "steps": k}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(host_cores(), 32))
    n = args.ref_n
    res = run_reference_cpu(n, args.material, args.steps, args.warmup, cores)
    E_work = args.n ** 3
    out = {
        "impl": "reference", "metric": "hex8 element-steps/sec fp64", "value": res["value"], "unit": "element-steps/s",
        "n_gpus": args.gpus, "steps": res["steps"], "warmup": args.warmup,
        "ms_per_step": 1e3 * res["loop_s"] / max(res["steps"], 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %d^3 structured hex8 cube (%d elements), %s, benchmark BC, CheckEnergy every "
                               "step; CPU arm timed on a bounded %d^3 sample of it" % (args.n, E_work, MAT_NAME[args.material], n)},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ GPU side
def ours_single(args):
    import torch
    from femtech_b200 import mesh, solver
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    n, mat = args.n, args.material
    X, conn, pid = mesh.cube_mesh(n, jitter=args.jitter) if args.jitter else mesh.cube_mesh(n)
    E, N = conn.shape[0], X.shape[0]
    tMax = 1e30
    dMax_over_tMax = 0.07 if mat == 1 else 1.75  # the drivers' ramp rate (0.007/0.1) / a faster one for stiff parts
    kind, rate = mesh.benchmark_bc(X, dMax=dMax_over_tMax, tMax=1.0)
    energy = 0 if args.no_energy else 1

    m = solver.FemTech(X, conn, pid, [mat], MATERIALS[mat], device=dev)
    stream = torch.cuda.Stream(device=dev)
    m.set_stream(stream.cuda_stream)
    m.ShapeFunctions()
    n_affine = m.affine_elements
    m.AssembleLumpedMass()
    m.set_bc(kind, rate)
    fp64_peak, copy_peak = solver.measure_peaks(m, reps=5)

    # ---- device-resident timing (value) ---------------------------------------------------------
    m.explicit_begin(energy_every=energy)
    if args.injury:  # not the headline: the brain drivers' per-step injury criteria on top of the same loop
        m.InitInjuryCriterion()
    m.run_async(tMax, max(args.warmup, 3))
    torch.cuda.synchronize()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, dev), daemon=True)
    th.start()
    l0 = m.gpu_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        m.run_async(tMax, args.steps)
        ev1.record(stream)
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    launches = m.gpu_launches - l0
    # ---- same region again with per-kernel CUDA events (roofline) -------------------------------
    m.profile(True)
    with torch.cuda.stream(stream):
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record(stream)
        m.run_async(tMax, args.steps)
        ev3.record(stream)
    torch.cuda.synchronize()
    prof = m.profile_get()
    ms_total_prof = ev2.elapsed_time(ev3)
    m.profile(False)
    stop.set()
    th.join(timeout=2)
    m._poll()
    assert np.isfinite(m.Time) and m.steps_done >= 2 * args.steps, "time loop did not advance"

    value = E * args.steps / (ms_total * 1e-3)
    b_elem, b_node = algorithmic_bytes(n, mat, energy)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # which kernel the parallelepipeds run: the current-Jacobian form unless switched off or the strain outputs are on
    cj = mat in ELEM_FLOPS_CJ and not args.injury and os.environ.get("FTB200_NH", "1") != "0"
    flops_aff = ELEM_FLOPS_CJ[mat] if cj else ELEM_FLOPS_AFFINE[mat]
    flops = (n_affine * flops_aff + (E - n_affine) * ELEM_FLOPS[mat]) / E
    fused = prof["node_launches"] == 0  # one fused kernel per step (k_step): element and node work in the same launch
    elem_s = prof["elem_ms"] * 1e-3
    if fused:
        rho_n = ((n + 1.0) / n) ** 3
        # k_step: conn, pid, eflag + X,u,v,a,flags gathered (unique nodes) + felem W | felem R + ELL + m, flags, u,v,a R/W (+ fi R/W)
        b_step = 37 + 98 * rho_n + 192 + 192 + (32 + 8 + 2 + 72 + 72) * rho_n + (48 * rho_n if energy else 0) + (2304 if mat == 5 else 0)
        f_step = flops + 60 * rho_n
        tf, gbs = f_step * E / elem_s / 1e12, b_step * E / elem_s / 1e9
        fp64_bound = tf / fp64_peak >= gbs / hbm_peak
        roofline = {
            "kernel": "k_step (fused: element forces + dt on the fp64 pipe, node assembly/update hidden under it)",
            "bound": "fp64" if fp64_bound else "hbm", "achieved": tf if fp64_bound else gbs,
            "peak": fp64_peak if fp64_bound else hbm_peak, "unit": "TFLOP/s" if fp64_bound else "GB/s",
            "frac": (tf / fp64_peak) if fp64_bound else (gbs / hbm_peak), "traffic": None,
            "peak_source": "fp64: DFMA microbenchmark measured in this run; hbm: " + hbm_src,
            "launch_ms": prof["elem_ms"], "launches_timed": prof["elem_launches"],
            "algorithmic_flops_per_element": f_step, "algorithmic_bytes_per_element": b_step,
            "fp64_view": {"achieved": tf, "peak": fp64_peak, "frac": tf / fp64_peak, "unit": "TFLOP/s"},
            "hbm_view": {"achieved": gbs, "peak": hbm_peak, "frac": gbs / hbm_peak, "unit": "GB/s"},
            "step": {"element_steps_per_s": value, "roof_fp64": fp64_peak * 1e12 / f_step, "roof_hbm": hbm_peak * 1e9 / b_step,
                     "frac_of_min_roof": value / min(fp64_peak * 1e12 / f_step, hbm_peak * 1e9 / b_step)},
            "copy_gbs_measured_here": copy_peak,
            "kernel_share_of_step": prof["elem_ms"] / (ms_total_prof / args.steps),
        }
    else:
        node_s = prof["node_ms"] * 1e-3
        elem_tf = flops * E / elem_s / 1e12
        elem_gbs = b_elem * E / elem_s / 1e9
        node_gbs = b_node * E / node_s / 1e9
        fp64_bound = elem_tf / fp64_peak >= elem_gbs / hbm_peak
        roofline = {
            "kernel": (("k_elem_affine_cj" if cj else "k_elem_affine") if n_affine == E else "k_elem") +
                      " (fused gather, F, material, B^T sigma, element dt)",
            "bound": "fp64" if fp64_bound else "hbm",
            "achieved": elem_tf if fp64_bound else elem_gbs,
            "peak": fp64_peak if fp64_bound else hbm_peak,
            "unit": "TFLOP/s" if fp64_bound else "GB/s",
            "frac": (elem_tf / fp64_peak) if fp64_bound else (elem_gbs / hbm_peak),
            # DRAM bytes of one k_elem launch from the ncu --set full capture of this configuration
            # (profiles/r01_k_elem_final_ncu_full.csv: dram__bytes_read.sum + dram__bytes_write.sum); other configs: null
            "traffic": NCU_TRAFFIC_BYTES.get((n, mat, bool(energy), bool(args.injury), ("cj" if cj else True) if n_affine == E else False)),
            "traffic_unit": "bytes per launch (algorithmic: %d)" % int(b_elem * E),
            "traffic_source": "PINNED CONSTANT, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one "
                              "launch from the ncu --set full capture profiles/r02_k_elem_affine_cj_ncu_full.csv (round 2; "
                              "k_elem_affine: r01_k_elem_affine_ncu_full.csv, general kernel: r01_k_elem_general_ncu_full.csv); "
                              "null for configurations without a capture",
            # what north_star scores is the STEP: both roofs and the step's fraction of the slower one, first
            "step_frac_of_min_roof": value / min(fp64_peak * 1e12 / flops, hbm_peak * 1e9 / (b_elem + b_node)),
            "step_roof_fp64": fp64_peak * 1e12 / flops, "step_roof_hbm": hbm_peak * 1e9 / (b_elem + b_node),
            "step_frac_of_hbm_roof_survey_bytes": value / (hbm_peak * 1e9 / 717.0),
            "step_note": "roofs in element-steps/s: fp64 = measured DFMA peak / executed flops per element, hbm = measured copy "
                         "bandwidth / algorithmic bytes per element-step of this design (%.0f B; SURVEY 8(d) counts 717 B: "
                         "step_frac_of_hbm_roof_survey_bytes)" % (b_elem + b_node),
            "peak_source": "fp64: DFMA microbenchmark measured in this run; hbm: " + hbm_src,
            "launch_ms": prof["elem_ms"], "launches_timed": prof["elem_launches"],
            "algorithmic_flops_per_element": flops, "algorithmic_bytes_per_element": b_elem,
            "hbm_view": {"achieved": elem_gbs, "peak": hbm_peak, "frac": elem_gbs / hbm_peak, "unit": "GB/s"},
            "k_node": {"bound": "hbm", "achieved": node_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": node_gbs / hbm_peak,
                       "launch_ms": prof["node_ms"], "algorithmic_bytes_per_element": b_node},
            "step": {"element_steps_per_s": value,
                     "roof_fp64": fp64_peak * 1e12 / flops, "roof_hbm": hbm_peak * 1e9 / (b_elem + b_node),
                     "frac_of_min_roof": value / min(fp64_peak * 1e12 / flops, hbm_peak * 1e9 / (b_elem + b_node))},
            "copy_gbs_measured_here": copy_peak,
            "kernel_share_of_step": (prof["elem_ms"]) / (ms_total_prof / args.steps),
        }

    # ---- end to end through the public API with host buffers --------------------------------------
    # (a) ExplicitDynamics(): state uploaded from pinned host arrays, K steps with the per-step scalars
    #     (Time, dt, energies) read back every step, final state downloaded -- all inside the timed region
    e2e_steps = args.steps  # the same K steps as the device-timed region; upload and download amortised over them
    pin = {k: torch.zeros(3 * N, dtype=torch.float64).pin_memory() for k in ("u", "v", "a", "fi", "fn")}
    pinb = torch.zeros(3 * N, dtype=torch.int32).pin_memory()
    m.displacements, m.velocities, m.accelerations = pin["u"].numpy(), pin["v"].numpy(), pin["a"].numpy()
    m.fi, m.f_net, m.boundary = pin["fi"].numpy(), pin["fn"].numpy(), pinb.numpy()
    m.Time = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.explicit_begin(energy_every=energy)        # H2D of u, v, a, boundary + step 0
    for _ in range(e2e_steps):
        m.run_async(tMax, 1)
        m._poll()                                # D2H of the step's scalars (Time, dt, status)
        if energy:
            m.energy()                           # D2H of Wint, Wext, WKE, total
    m.sync_out()                                 # D2H of u, v, a, boundary, fi, f_net
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = (3 * 24 * N + 12 * N) / e2e_steps
    d2h = (5 * 24 * N + 12 * N) / e2e_steps + 200 + (32 if energy else 0)
    e2e_calls = {"value": E * e2e_steps / e2e_s, "unit": "element-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                 "api": "ExplicitDynamics (resident): pinned host state in, %d single-step calls each followed by a blocking "
                        "read-back of the step scalars, host state out; copies inside the timed region" % e2e_steps,
                 "steps": e2e_steps}
    # (a0) the headline: the same K steps as ONE ExplicitDynamics call.  The device writes each finished step's record
    #      (Time, dt, step, status, energies: 64 bytes) straight into the pinned host ring of ftb200_step_ring and the
    #      host consumes the K records as they arrive; state upload before and download after, all inside the timed region
    ring = m.step_ring(max(e2e_steps, 1))
    m.displacements[:] = 0.0; m.velocities[:] = 0.0; m.accelerations[:] = 0.0; m.boundary[:] = 0
    m.Time = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.explicit_begin(energy_every=energy)        # H2D of u, v, a, boundary + step 0
    m.run_async(tMax, e2e_steps)                 # enqueues the whole loop (CUDA graphs of 25 steps)
    recs = np.empty((e2e_steps, 8))
    for k in range(1, e2e_steps + 1):
        recs[k - 1] = m.wait_step(k)             # the step's result, written by the device over PCIe
    m.sync_out()                                 # D2H of u, v, a, boundary, fi, f_net
    torch.cuda.synchronize()
    e2e_ring_s = time.perf_counter() - t0
    ok_recs = bool(np.all(np.diff(recs[:, 0]) > 0) and np.all(recs[:, 2] == np.arange(1, e2e_steps + 1)) and
                   np.all(recs[:, 3] == 0) and (not energy or np.all(np.isfinite(recs[:, 4:]))))
    assert ok_recs, "step ring records inconsistent"
    m.step_ring(0)
    e2e = {"value": E * e2e_steps / e2e_ring_s, "unit": "element-steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": (5 * 24 * N + 12 * N) / e2e_steps + 64, "steps": e2e_steps, "records_consistent": ok_recs,
           "amortised_over_steps": e2e_steps,
           "amortisation": "the state upload (%.0f MB) and download (%.0f MB) cross PCIe ONCE per call and are spread over the %d "
                           "steps of the call: e2e approaches `value` as the call gets longer" % (h2d * e2e_steps / 1e6, (5 * 24 * N + 12 * N) / 1e6, e2e_steps),
           "api": "ExplicitDynamics (resident), one call for the %d steps: pinned host state in, every step's scalars "
                  "(Time, dt, status, energies) written by the device into a pinned host ring and consumed by the host "
                  "as they arrive, host state out; all copies inside the timed region" % e2e_steps}
    # (a') the same, but the per-step scalars are copied device -> host asynchronously into a pinned ring (one 64-byte
    #      D2H per step inside the timed region, consumed after a single synchronisation at the end)
    ring = torch.zeros(e2e_steps, 8, dtype=torch.float64).pin_memory()
    rnp = ring.numpy()
    m.displacements[:] = 0.0; m.velocities[:] = 0.0; m.accelerations[:] = 0.0; m.boundary[:] = 0
    m.Time = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.explicit_begin(energy_every=energy)
    for i in range(e2e_steps):
        m.run_async(tMax, 1)
        m.poll_async(rnp[i])
    m.sync_out()
    torch.cuda.synchronize()
    e2e_async_s = time.perf_counter() - t0
    ok_ring = bool(np.all(np.diff(rnp[:, 0]) > 0) and np.all(rnp[:, 2] == np.arange(1, e2e_steps + 1)))
    e2e_async = {"value": E * e2e_steps / e2e_async_s, "unit": "element-steps/s", "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": (5 * 24 * N + 12 * N) / e2e_steps + 64, "steps": e2e_steps, "ring_consistent": ok_ring,
                 "api": "as e2e, but the step scalars go device -> host with an asynchronous 64-byte copy per step into a pinned "
                        "ring and the host synchronises once at the end"}
    # (b) strict drop-in: the shipped drivers' four library calls per step, host arrays across PCIe every call
    leg_steps = min(args.steps, 10)
    bc = kind > 0
    m.boundary[bc] = 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(leg_steps):
        m.GetForce()
        m.CalculateAccelerations()
        m.CheckEnergy(m.Time, 1)
        m.dt = 0.8 * m.StableTimeStep()
    leg_s = time.perf_counter() - t0
    e2e_legacy = {"value": E * leg_steps / leg_s, "unit": "element-steps/s",
                  "h2d_bytes_per_step": (24 + 12 + 24 + 9 * 24 + 12 + 24 + 12) * N, "d2h_bytes_per_step": (24 + 24 + 24) * N,
                  "api": "legacy drop-in: GetForce + CalculateAccelerations + CheckEnergy + StableTimeStep on host arrays "
                         "(driver host loops not included)", "steps": leg_steps}

    # (c) N = 1 through the loop the multi-GPU runs take (split element launches, pack, dt through the peer-memory window,
    #     k_adv_p2p), zero neighbours: separates the cost of that loop from the cost of scaling in the N > 1 lines
    part_path = None
    try:
        m.Time = 0.0
        m.displacements[:] = 0.0; m.velocities[:] = 0.0; m.accelerations[:] = 0.0; m.boundary[:] = 0
        m.enable_partitioned_loop()
        m.explicit_begin(energy_every=energy)
        m.run_async(tMax, max(args.warmup, 3))
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev4.record(stream)
            m.run_async(tMax, args.steps)
            ev5.record(stream)
        torch.cuda.synchronize()
        ms_p = ev4.elapsed_time(ev5)
        part_path = {"value": E * args.steps / (ms_p * 1e-3), "unit": "element-steps/s", "ms_per_step": ms_p / args.steps,
                     "what": "the same mesh on 1 GPU through the partitioned (peer-memory) loop of bench.py --gpus N > 1 with no "
                             "neighbour: the N > 1 lines should be compared with this, `value` is the single-partition loop"}
    except Exception as ex:  # never fail the headline over the extra leg
        part_path = {"error": str(ex)[:200]}

    cpu = None
    if not args.no_cpu:
        cpu = run_reference_cpu(args.ref_n, mat, 40, 5, max(1, min(host_cores(), 32)))
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    out = {
        "metric": "hex8 element-steps/sec fp64", "value": value, "unit": "element-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %d^3 %s hex8 cube (%d elements, %d nodes), %s, benchmark BC "
                               "(Benchmarking-Parallel.cpp:184-244), %s, dt recomputed every step"
                               % (n, ("structured, nodes jittered by %g of the spacing" % args.jitter) if args.jitter else "structured",
                                  E, N, MAT_NAME[mat], "CheckEnergy every step" if energy else "no energy check"),
                   "element_kernel": "%d of %d hexahedra have a parallelepiped reference geometry and run %s; the rest run the "
                                     "general k_elem" % (n_affine, E, "k_elem_affine_cj (current-Jacobian form)" if cj else
                                                         "k_elem_affine (dN/dX once per element)"),
                   "injury_criteria": bool(args.injury),
                   "mode": "resident ExplicitDynamics loop, CUDA graph of 25 steps, " +
                           ("one fused kernel per step" if prof["node_launches"] == 0 else "element + node kernels per step"),
                   "l2": "per-step working set %.2f GB > 126 MB L2, no flush needed" % ((b_elem + b_node) * E / 1e9)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_per_step_calls": e2e_calls, "e2e_async": e2e_async,
        "e2e_legacy": e2e_legacy, "n1_partitioned_loop": part_path,
        "gpu_launches": launches, "clocks": summarize_clocks(samples),
        "ms_per_step_with_kernel_events": ms_total_prof / args.steps,
        "fp64_peak_tflops_measured": fp64_peak,
    }
    m.close()
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--edge", dest="n", type=int, default=100, help="cube edge in elements per GPU (BASELINE config 2: 100)")
    ap.add_argument("--material", type=int, default=1, choices=[1, 4, 5])
    ap.add_argument("--ref-n", type=int, default=40, help="edge of the bounded CPU sample")
    ap.add_argument("--no-energy", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="multi-GPU: weak = n^3 elements per GPU (default), "
                    "strong = one n^3 cube cut over the GPUs (BASELINE configs 3 and 4)")
    ap.add_argument("--no-validate", action="store_true", help="multi-GPU: skip the state check against the single-GPU run of the global mesh")
    ap.add_argument("--injury", action="store_true", help="also evaluate the injury criteria every step (ex5.cpp:240)")
    ap.add_argument("--jitter", type=float, default=0.0, help="move the interior nodes by this fraction of the spacing: "
                    "no element is a parallelepiped any more, so the general element kernel is measured")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from femtech_b200 import dist_bench
        dist_bench.run(args)
        return
    ours_single(args)


if __name__ == "__main__":
    main()
