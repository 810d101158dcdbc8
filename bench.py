#!/usr/bin/env python
"""bench.py -- short-range G interactions/s on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the hot path over one sub-volume: tree build + interaction lists + force
kernel (= one RCBForceTree constructor call of the reference, src/cpu/Particles.cxx:1313-1338).
  value     whole-job evaluated pairs per second with the particles already resident in HBM
  e2e       same, through the C ABI with HOST buffers: upload (H2D) + kick + download (D2H) per step
  roofline  force kernel only: 30 flop per evaluated pair / its CUDA-event time,
            against the FP32 FMA peak  n_SM * 128 lanes * 2 * sm_max_mhz
  cpu_baseline  the reference's own compiled sources (oracle/_ref) on a bounded cut-out of the same snapshot
`--impl reference` times that CPU reference alone (all host threads) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the CPU reference keeps its interaction lists on the stacks of its OpenMP workers (4 * VMAX floats, RCBForceTree.cxx:940);
# libgomp reads the variable once, when the first OpenMP runtime of the process starts (torch's import), so it is set here
os.environ.setdefault("OMP_STACKSIZE", "64M")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 30          # poly5 law, FMA = 2, rsqrt = 1, compare/select = 0
RSM, EDGE, THETA = 0.007, 3.2, 0.5   # reference indat:37-41
GHOST = 11                  # overload cells per face at the shipped spacing


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--np-side", type=int, default=256, help="alive particles per dimension per GPU")
    ap.add_argument("--ppn", type=int, default=512, help="leaf size (reference -N 512, run_hacc.sh:2)")
    ap.add_argument("--state", default="uniform", choices=["uniform", "clustered", "clumpy"],
                    help="uniform = z=50 Zel'dovich (configs[1]); clustered = shell-crossed Zel'dovich (configs[2]); clumpy = "
                         "clustered + 15 %% of the particles in 64 isothermal knots (stress case: lists beyond the reference's VMAX)")
    ap.add_argument("--sample-side", type=int, default=112,
                    help="cut-out side (cells) for the CPU baseline: 112^3 cells = 1.4 M particles = about 10 s on 16 cores")
    ap.add_argument("--arith", default="fused", choices=["fused", "x86"], help="pair-kernel arithmetic (include/haccsr.h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--subcycle", type=int, default=0, metavar="NSUB",
                    help="also time Particles::subCycle on the device (haccsr_subcycle: NSUB x [stream, out-of-box compaction, mass=1, "
                         "kick, stream], one upload and one download) and report it as the side block 'subcycle' (configs[4])")
    ap.add_argument("--cull", action="store_true", help="headline run with warp-level culling on (haccsr_set_culling); "
                    "by default culling is off and only a side measurement of it is reported under 'culled'")
    return ap.parse_args()


def ncu_traffic(kernel, np_side, state, arith):
    """DRAM bytes per launch of `kernel` from a committed `ncu --set full` capture of this bench command
    (profiles/traffic.json, written by tools/summarize_ncu.py --traffic); None when no capture matches."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        for e in json.load(f):
            if e["kernel"] == kernel and e["np_side"] == np_side and e["state"] == state and e.get("arith", "fused") == arith:
                return e["dram_bytes_per_launch"]
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([t.strip() for t in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class NumaLocal:
    """Pins the calling thread to the CPUs next to GPU `index` while host buffers are allocated (pages of page-locked
    memory land on the NUMA node of the allocating thread), then restores the affinity.  With 8 ranks on one box the
    host<->device copies of the e2e path otherwise cross the socket interconnect for half of the GPUs."""

    def __init__(self, index):
        self.index, self.saved = index, None

    def __enter__(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, cpus)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def make_snapshot(args, rank, device):
    from hacc_coral_b200 import synth
    boost = 1.0
    z = 50.0
    if args.state in ("clustered", "clumpy"):
        z, boost = 0.0, 0.35      # Zel'dovich pushed to shell crossing: sheets / filaments / knots
    try:
        p = synth.zeldovich_torch(args.np_side, z=z, seed=5009888 + rank, ghost=GHOST, growth_boost=boost, device=device)
    except Exception:
        p = synth.zeldovich(args.np_side, z=z, seed=5009888 + rank, ghost=GHOST, growth_boost=boost)
    if args.state == "clumpy":
        synth.add_clumps(p, float(args.np_side + 2 * GHOST), seed=99 + rank)
    return p


def run_reference_sample(p, nglt, args):
    """Time the compiled reference (oracle/_ref) on a cut-out of the snapshot.  Returns dict or None."""
    from hacc_coral_b200 import synth
    from oracle import refbind, oraclebind
    if not refbind.available():
        return None
    side = min(args.sample_side, nglt)
    lo = (nglt - side) // 2
    q = synth.cutout(p, lo, lo + side)
    box = ([0.0] * 3, [float(side)] * 3, [EDGE] * 3, [float(side) - EDGE] * 3)
    # pair count of the identical call from the plain-C restatement (walk only, no force)
    walk = oraclebind.run(q, *box, RSM, THETA, args.ppn, do_force=False)["stats"]
    cnt = walk["pairs_eval"]
    # the reference keeps its lists in fixed stack arrays of VMAX = 16384 entries and asserts on overflow
    # (RCBForceTree.cxx:921,1039,1070): where the cut-out's longest list comes close, the build with the #define raised
    # (oracle/build_ref.sh) is timed instead, and the sample says so
    big = walk["max_list"] >= 16000 and refbind.available(vmax=True)
    t0 = time.time()
    _, st, _ = refbind.rcb_kick(q, *box, RSM, THETA, args.ppn, fcoeff=1.0, law=refbind.LAW_POLY5, vmax=big)
    dt = time.time() - t0
    return {"pairs": int(cnt), "seconds": dt, "particles": int(q["x"].size), "side": side,
            "cores": os.cpu_count(), "wall_ctor_s": st["wall_s"], "note": ", VMAX raised to 1048576" if big else ""}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nglt = args.np_side + 2 * GHOST
    workload = "np=%d^3 alive per GPU + %d-cell overload shell (%d^3 grid units), %s Zel'dovich snapshot, ppn=%d, theta=%.1f, poly5" % (
        args.np_side, GHOST, nglt, {"uniform": "z=50 near-uniform", "clustered": "shell-crossed clustered", "clumpy": "shell-crossed + isothermal knots"}[args.state], args.ppn, THETA)
    config = {"workload": workload, "np_side": args.np_side, "ppn": args.ppn, "theta": THETA, "rsm": RSM,
              "force_law": "poly5", "arithmetic": args.arith, "state": args.state, "l2": "inputs larger than L2 (no flush needed)",
              "parallelism": "1 sub-volume per GPU, no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        import torch  # noqa: F401  (only for the snapshot generator's fallback path)
        # the reference arm only needs the cut-out: a snapshot of the same recipe (spacing, redshift, seed)
        # just large enough to contain it is generated on the CPU instead of the full field
        sub = argparse.Namespace(**vars(args))
        sub.np_side = min(args.np_side, max(64, args.sample_side + 8))
        p = make_snapshot(sub, 0, "cpu")
        vals = []
        info = None
        for it in range(args.warmup + args.steps):
            info = run_reference_sample(p, sub.np_side + 2 * GHOST, args)
            if info is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhaccref.so not built"}))
                return 0
            if it >= args.warmup:
                vals.append(info["pairs"] / info["seconds"] / 1e9)
        v = float(np.mean(vals))
        sample = "%d^3-cell cut-out (%d particles, %d pairs) of the same snapshot recipe, full RCBMonopoleForceTree ctor%s" % (
            info["side"], info["particles"], info["pairs"], info["note"])
        line = {"metric": "short-range G interactions/s", "value": v, "unit": "Ginteractions/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "impl": "reference", "cpu_baseline": {"value": v, "unit": "Ginteractions/s", "cores": info["cores"],
                                                      "kind": "reference", "sample": sample},
                "e2e": {"value": v, "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import hacc_coral_b200 as H
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; libhaccsr has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = "cuda:%d" % local
    p = make_snapshot(args, rank, dev)
    n = int(p["x"].size)
    # pinned host copies (the caller's arrays in the e2e path)
    pin = {}
    with NumaLocal(local):
        for k, v in p.items():
            t = torch.from_numpy(v).pin_memory()
            pin[k] = t.numpy()
    g = H.HaccSR(n, device=local, arith=H.ARITH_FUSED if args.arith == "fused" else H.ARITH_X86)
    g.set_force_law(H.LAW_SR_POLY, H.POLY5, RSM, H.RMAX)
    g.set_culling(args.cull)
    stream = torch.cuda.current_stream()
    g.set_stream(stream.cuda_stream)
    lo, hi = [0.0] * 3, [float(nglt)] * 3
    flo, fhi = [EDGE] * 3, [float(nglt) - EDGE] * 3
    g.upload(pin)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-input throughput -----------------------------------------------------------------
    for _ in range(args.warmup):
        st = g.kick(lo, hi, flo, fhi, THETA, args.ppn)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pairs = 0
    ms_force = ms_build = ms_walk = 0.0
    launches = force_launches = 0
    for _ in range(args.steps):
        st = g.kick(lo, hi, flo, fhi, THETA, args.ppn)
        pairs += st["pairs_evaluated"]
        ms_force += st["ms_force"]; ms_build += st["ms_build"]; ms_walk += st["ms_walk"]
        launches += st["total_launches"]; force_launches += st["force_launches"]
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    # ---- end-to-end through the C ABI with host buffers ------------------------------------------------
    # the call a user of the reference's constructor makes: haccsr_kick_host on caller-owned (page-locked) host
    # arrays = H2D of all ten arrays + build + walk + force + D2H of all ten arrays, every step.  The arrays are
    # kicked in place, so each step starts from the previous step's output (same particles, tree order).
    with NumaLocal(local):
        work = {k: torch.from_numpy(v.copy()).pin_memory().numpy() for k, v in pin.items()}
    g.kick_host(work, lo, hi, flo, fhi, THETA, args.ppn)   # warm
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    e2e_steps = max(1, min(args.steps, 3))
    pairs_e2e = 0
    for _ in range(e2e_steps):
        st2 = g.kick_host(work, lo, hi, flo, fhi, THETA, args.ppn)
        pairs_e2e += st2["pairs_evaluated"]
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    g.upload(pin)
    # one untimed pass that also counts the pairs inside the cutoff (honest-metric companion number)
    stc = g.kick(lo, hi, flo, fhi, THETA, args.ppn, count_in_cutoff=True)
    # side measurement: the same kick with warp-level culling (bit-identical result, fewer executed flops)
    culled = None
    if args.arith == "fused":
        g.set_culling(not args.cull)
        g.kick(lo, hi, flo, fhi, THETA, args.ppn)
        other = [g.kick(lo, hi, flo, fhi, THETA, args.ppn) for _ in range(3)]
        oc = g.kick(lo, hi, flo, fhi, THETA, args.ppn, count_in_cutoff=True)
        g.set_culling(args.cull)
        on, off_ms = (stc, np.mean([o["ms_force"] for o in other])) if args.cull else (oc, ms_force / args.steps)
        on_ms = ms_force / args.steps if args.cull else np.mean([o["ms_force"] for o in other])
        culled = {"ms_force": float(on_ms), "ms_force_unculled": float(off_ms), "speedup_force": float(off_ms / on_ms),
                  "pairs_force_law_frac": on["pairs_force_law"] / max(on["pairs_evaluated"], 1),
                  "note": "warp-level early exit after the cutoff test; results bit-identical; off in the headline unless --cull"}

    # side measurement: the full short-range sub-cycle loop of one long step, particles resident between the kicks
    subc = None
    if args.subcycle > 0:
        vmax = max(float(np.abs(pin[k]).max()) for k in ("vx", "vy", "vz"))
        # each half-stream moves the fastest particle 0.02 cells; the parity snapshots carry v = 0, then only the kicks move them
        pt = 0.02 / vmax if vmax > 0 else 0.01
        sub_args = (args.subcycle, pt, [float(nglt)] * 3, lo, hi, flo, fhi, THETA, args.ppn, 1e-3)
        g.upload(pin)
        g.subcycle(*sub_args)                         # warm
        barrier()
        s0, s1, s2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        s0.record(stream)
        g.upload(pin)
        s1.record(stream)
        sst = g.subcycle(*sub_args)
        s2.record(stream)
        g.download(out=work)
        s3 = torch.cuda.Event(enable_timing=True)
        s3.record(stream)
        barrier()
        subc = {"nsub": args.subcycle, "ms_resident": s1.elapsed_time(s2), "ms_with_transfers": s0.elapsed_time(s3),
                "pairs_evaluated": int(sst["pairs_evaluated"]), "ms_force": sst["ms_force"], "ms_build": sst["ms_build"],
                "value": sst["pairs_evaluated"] / (s1.elapsed_time(s2) * 1e-3) / 1e9,
                "e2e_value": sst["pairs_evaluated"] / (s0.elapsed_time(s3) * 1e-3) / 1e9, "unit": "Ginteractions/s",
                "note": "haccsr_subcycle = Particles::subCycle (Particles.cxx:1176-1201) on the device; rank 0's numbers"}
        g.upload(pin)

    tv = torch.tensor([ms, ms_e2e, ms_force], device=dev, dtype=torch.float64)
    sv = torch.tensor([float(pairs), float(pairs_e2e), float(launches)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(sv, op=dist.ReduceOp.SUM)
    ms_max, ms_e2e_max, ms_force_max = [float(t) for t in tv.cpu()]
    pairs_all, pairs_e2e_all, launches_all = [float(t) for t in sv.cpu()]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    pk, pk_kind = peaks()
    props = torch.cuda.get_device_properties(local)
    sm_max = float(pk.get("sm_max_mhz", 1965.0))
    fp32_peak = props.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12     # TFLOP/s
    achieved = FLOP_PER_PAIR * pairs / (ms_force * 1e-3) / 1e12                   # rank 0's force kernel
    value = pairs_all / (ms_max * 1e-3) / 1e9
    e2e_v = pairs_e2e_all / (ms_e2e_max * 1e-3) / 1e9
    bytes_pp = 42
    build_bytes = st["levels"] * 88.0 * n + 84.0 * n     # 88 B/particle/level + final gather
    line = {
        "metric": "short-range G interactions/s", "value": value, "unit": "Ginteractions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_v, "unit": "Ginteractions/s", "h2d_bytes_per_step": bytes_pp * n, "d2h_bytes_per_step": bytes_pp * n,
                "ms_per_step": ms_e2e_max / e2e_steps},
        "gpu_launches": int(launches_all),
        "roofline": {"bound": "fp32", "kernel": "k_force", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak, "traffic": ncu_traffic("k_force", args.np_side, args.state, args.arith), "peak_source": "%d SMs x 128 lanes x 2 x %.0f MHz (%s sm_max_mhz)" % (
                         props.multi_processor_count, sm_max, pk_kind),
                     "flop_per_interaction": FLOP_PER_PAIR, "ms_per_launch": ms_force / max(force_launches, 1)},
        "roofline_build": {"bound": "hbm", "kernel": "tree build (k_cm_tile + k_left_count + k_scatter + k_gather)",
                           "achieved": build_bytes / (ms_build / args.steps * 1e-3) / 1e9, "peak": pk.get("hbm_gbs"),
                           "unit": "GB/s", "frac": build_bytes / (ms_build / args.steps * 1e-3) / 1e9 / pk.get("hbm_gbs", 6650.0)},
        "phases_ms": {"build": ms_build / args.steps, "walk": ms_walk / args.steps, "force": ms_force / args.steps},
        "particles_per_gpu": n, "pairs_per_particle": st["pairs_evaluated"] / n,
        "pairs_in_cutoff_frac": stc["pairs_in_cutoff"] / max(stc["pairs_evaluated"], 1),
        "tree": {"nodes": st["nodes"], "leaves": st["leaves"], "mean_ppn": st["mean_ppn"], "levels": st["levels"],
                 "max_list": st["max_list"], "pseudo_particles": st["pseudo_particles"]},
        "clocks": sampler.summary(),
    }
    if culled:
        line["culled"] = culled
    if subc:
        line["subcycle"] = subc
    if args.cull:
        # with culling the kernel executes 30 flop only for the pairs that reach the force law and 9 (three differences,
        # the r2 chain, the softening add) for the rest: report the executed rate next to the algorithmic one
        ex = (30.0 * stc["pairs_force_law"] + 9.0 * (stc["pairs_evaluated"] - stc["pairs_force_law"])) / max(stc["pairs_evaluated"], 1)
        line["roofline"]["executed_flop_per_interaction"] = ex
        line["roofline"]["frac_executed"] = line["roofline"]["frac"] * ex / FLOP_PER_PAIR
        line["config"]["culling"] = "on"
    if not args.no_cpu_baseline and world == 1:      # the contract: rank 0 at N = 1 only (the reference arm covers every N)
        try:
            info = run_reference_sample(p, nglt, args)
        except Exception as ex:   # the baseline is a reported companion number; never fail the GPU line
            info = None
            line["cpu_baseline_error"] = repr(ex)
        if info:
            line["cpu_baseline"] = {
                "value": info["pairs"] / info["seconds"] / 1e9, "unit": "Ginteractions/s", "cores": info["cores"],
                "kind": "reference",
                "sample": "%d^3-cell cut-out (%d particles, %d pairs) of rank 0's snapshot, full RCBMonopoleForceTree ctor%s, %.1f s" % (
                    info["side"], info["particles"], info["pairs"], info["note"], info["seconds"])}
    print(json.dumps(line))
    g.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
