#!/usr/bin/env python
"""bench.py -- short-range G interactions/s on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the hot path over one sub-volume: tree build + interaction lists + force
kernel (= one RCBForceTree constructor call of the reference, src/cpu/Particles.cxx:1313-1338).
  value     whole-job evaluated pairs per second with the particles already resident in HBM
  e2e       same, through the C ABI with HOST buffers: upload (H2D) + kick + download (D2H) per step
            (haccsr_kick_host, the facade constructor's path); e2e.subcycle is the Level-2 path of INTEGRATION.md,
            haccsr_upload + haccsr_subcycle(nsub) + haccsr_download: the same copies amortised over nsub kicks
  roofline  force kernel only: 30 flop per evaluated pair / its CUDA-event time,
            against the FP32 FMA peak  n_SM * 128 lanes * 2 * sm_max_mhz
  roofline_build  tree build: N * (28 L + 84) algorithmic bytes / its CUDA-event time against the measured HBM peak
  clustered the same measurements on configs[2] (shell-crossed snapshot), the state north_star's target names
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
# libgomp reads the variables once, when the first OpenMP runtime of the process starts (torch's import), so they are set here
os.environ.setdefault("OMP_STACKSIZE", "64M")
if "reference" in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers when nproc > 1; the reference arm is the CPU
    # implementation with all the host threads it can use, on rank 0 alone
    try:
        _ncpu = len(os.sched_getaffinity(0))
    except Exception:
        _ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(_ncpu)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 30          # poly5 law, FMA = 2, rsqrt = 1, compare/select = 0
RSM, EDGE, THETA = 0.007, 3.2, 0.5   # reference indat:37-41
GHOST = 11                  # overload cells per face at the shipped spacing
STATE_NAME = {"uniform": "z=50 near-uniform", "clustered": "shell-crossed clustered", "clumpy": "shell-crossed + isothermal knots"}


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
    ap.add_argument("--no-clustered-block", action="store_true",
                    help="skip the side block 'clustered' (configs[2]) that the default uniform line carries")
    ap.add_argument("--sample-side", type=int, default=0,
                    help="cut-out side (cells) for the CPU baseline; 0 = sized from a calibration run so the arm stays within --cpu-budget")
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="seconds of CPU reference work allowed in total (reference arm)")
    ap.add_argument("--arith", default="fused", choices=["fused", "x86", "fused_rs3"], help="pair-kernel arithmetic (include/haccsr.h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--subcycle", type=int, default=5, metavar="NSUB",
                    help="sub-cycles per long step for the Level-2 end-to-end path (haccsr_subcycle; reference indat nsub = 5); 0 = skip")
    ap.add_argument("--cull", action="store_true", help="headline run with warp-level culling on (haccsr_set_culling); "
                    "by default culling is off and only a side measurement of it is reported under 'culled'")
    ap.add_argument("--no-refresh-block", action="store_true",
                    help="skip the side block 'refresh' (overload refresh of one global snapshot decomposed over the ranks, configs[3])")
    ap.add_argument("--tune-ppn", default="64,128,256", help="comma-separated leaf sizes for the side block 'tuned' (time to solution per "
                    "kick at other -N than the reference's shipped 512; '' = skip)")
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


def make_snapshot(args, state, rank, device, np_side=None):
    from hacc_coral_b200 import synth
    boost, z = 1.0, 50.0
    if state in ("clustered", "clumpy"):
        z, boost = 0.0, 0.35      # Zel'dovich pushed to shell crossing: sheets / filaments / knots
    side = np_side or args.np_side
    try:
        p = synth.zeldovich_torch(side, z=z, seed=5009888 + rank, ghost=GHOST, growth_boost=boost, device=device)
    except Exception:
        p = synth.zeldovich(side, z=z, seed=5009888 + rank, ghost=GHOST, growth_boost=boost)
    if state == "clumpy":
        synth.add_clumps(p, float(side + 2 * GHOST), seed=99 + rank)
    return p


def run_reference_sample(p, nglt, side, ppn):
    """Time the compiled reference (oracle/_ref) on a `side`^3-cell cut-out of the snapshot.  Returns dict or None."""
    from hacc_coral_b200 import synth
    from oracle import refbind, oraclebind
    if not refbind.available():
        return None
    side = min(side, nglt)
    lo = (nglt - side) // 2
    q = synth.cutout(p, lo, lo + side)
    box = ([0.0] * 3, [float(side)] * 3, [EDGE] * 3, [float(side) - EDGE] * 3)
    # pair count of the identical call from the plain-C restatement (walk only, no force)
    walk = oraclebind.run(q, *box, RSM, THETA, ppn, do_force=False)["stats"]
    cnt = walk["pairs_eval"]
    # the reference keeps its lists in fixed stack arrays of VMAX = 16384 entries and asserts on overflow
    # (RCBForceTree.cxx:921,1039,1070): where the cut-out's longest list comes close, the build with the #define raised
    # (oracle/build_ref.sh) is timed instead, and the sample says so
    big = walk["max_list"] >= 16000 and refbind.available(vmax=True)
    t0 = time.time()
    _, st, _ = refbind.rcb_kick(q, *box, RSM, THETA, ppn, fcoeff=1.0, law=refbind.LAW_POLY5, vmax=big)
    dt = time.time() - t0
    return {"pairs": int(cnt), "seconds": dt, "particles": int(q["x"].size), "side": side,
            "cores": int(os.environ.get("OMP_NUM_THREADS", 0)) or len(os.sched_getaffinity(0)),
            "wall_ctor_s": st["wall_s"], "note": ", VMAX raised to 1048576" if big else ""}


def reference_arm(args, config):
    """`--impl reference`: the reference's own compiled RCBForceTree on the host cores, on a cut-out of the same snapshot
    recipe sized so that warmup + steps repetitions stay within --cpu-budget seconds."""
    import torch  # noqa: F401  (only for the snapshot generator's fallback path)
    reps = args.warmup + args.steps
    calib_side = 56
    sub_side = min(args.np_side, 136)
    p = make_snapshot(args, args.state, 0, "cpu", np_side=sub_side)       # a snapshot just large enough to hold any cut-out
    nglt = sub_side + 2 * GHOST
    info = run_reference_sample(p, nglt, calib_side, args.ppn)            # calibration, untimed
    if info is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhaccref.so not built"}))
        return 0
    if args.sample_side > 0:
        side = args.sample_side
    else:
        per_particle = info["seconds"] / max(info["particles"], 1)
        budget = max(args.cpu_budget - 1.5 * info["seconds"], 5.0) / max(reps, 1)
        side = int(max(48, min(nglt, (budget / per_particle) ** (1.0 / 3.0))))
        side -= side % 8
    vals = []
    for it in range(reps):
        info = run_reference_sample(p, nglt, side, args.ppn)
        if it >= args.warmup:
            vals.append(info["pairs"] / info["seconds"] / 1e9)
    v = float(np.mean(vals))
    sample = "SAMPLED: %d^3-cell cut-out (%d particles, %d pairs) of the same snapshot recipe, full RCBMonopoleForceTree ctor%s, %d OpenMP threads" % (
        info["side"], info["particles"], info["pairs"], info["note"], info["cores"])
    line = {"metric": "short-range G interactions/s", "value": v, "unit": "Ginteractions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "impl": "reference", "cpu_baseline": {"value": v, "unit": "Ginteractions/s", "cores": info["cores"],
                                                  "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


class Bench:
    """One rank's measurements on one snapshot."""

    def __init__(self, args, rank, world, local, dist):
        import torch
        import hacc_coral_b200 as H
        self.torch, self.H, self.args, self.rank, self.world, self.local, self.dist = torch, H, args, rank, world, local, dist
        self.dev = "cuda:%d" % local
        self.nglt = args.np_side + 2 * GHOST
        self.lo, self.hi = [0.0] * 3, [float(self.nglt)] * 3
        self.flo, self.fhi = [EDGE] * 3, [float(self.nglt) - EDGE] * 3
        self.stream = torch.cuda.current_stream()
        self.g = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def events(self, k):
        return [self.torch.cuda.Event(enable_timing=True) for _ in range(k)]

    def load(self, state):
        """Generate the snapshot of `state`, page-lock it, create the context and upload."""
        torch, H, args = self.torch, self.H, self.args
        if self.g is not None:
            self.g.close()
        self.p = make_snapshot(args, state, self.rank, self.dev)
        self.n = int(self.p["x"].size)
        with NumaLocal(self.local):
            self.pin = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in self.p.items()}
            self.work = {k: torch.from_numpy(v.copy()).pin_memory().numpy() for k, v in self.p.items()}
        g = H.HaccSR(self.n, device=self.local, arith={"fused": H.ARITH_FUSED, "x86": H.ARITH_X86, "fused_rs3": H.ARITH_FUSED_RS3}[args.arith])
        g.set_force_law(H.LAW_SR_POLY, H.POLY5, RSM, H.RMAX)
        g.set_culling(args.cull)
        g.set_stream(self.stream.cuda_stream)
        g.upload(self.pin)
        self.g = g

    def kick(self, ppn=None, **kw):
        return self.g.kick(self.lo, self.hi, self.flo, self.fhi, THETA, ppn or self.args.ppn, **kw)

    def resident(self, steps, warmup, sampler=None):
        """`steps` kicks on resident particles between two events (after `warmup` untimed ones)."""
        for _ in range(warmup):
            st = self.kick()
        self.barrier()
        if sampler:
            sampler.start()
        e0, e1 = self.events(2)
        e0.record(self.stream)
        acc = {"pairs": 0, "ms_force": 0.0, "ms_build": 0.0, "ms_walk": 0.0, "launches": 0, "force_launches": 0}
        for _ in range(steps):
            st = self.kick()
            acc["pairs"] += st["pairs_evaluated"]
            acc["ms_force"] += st["ms_force"]; acc["ms_build"] += st["ms_build"]; acc["ms_walk"] += st["ms_walk"]
            acc["launches"] += st["total_launches"]; acc["force_launches"] += st["force_launches"]
        e1.record(self.stream)
        self.barrier()
        if sampler:
            sampler.stop_flag = True
        acc["ms"] = e0.elapsed_time(e1)
        acc["steps"] = steps
        acc["st"] = st
        return acc

    def e2e_facade(self, steps):
        """haccsr_kick_host on caller-owned page-locked arrays = H2D of all ten arrays + build + walk + force + D2H of all
        ten arrays, every step.  The arrays are kicked in place (each step starts from the previous step's output)."""
        self.g.kick_host(self.work, self.lo, self.hi, self.flo, self.fhi, THETA, self.args.ppn)   # warm
        self.barrier()
        e0, e1 = self.events(2)
        e0.record(self.stream)
        pairs = 0
        for _ in range(steps):
            st = self.g.kick_host(self.work, self.lo, self.hi, self.flo, self.fhi, THETA, self.args.ppn)
            pairs += st["pairs_evaluated"]
        e1.record(self.stream)
        self.barrier()
        return {"pairs": pairs, "ms": e0.elapsed_time(e1), "steps": steps}

    def e2e_subcycle(self, nsub, reps):
        """Level-2 path: one haccsr_upload, haccsr_subcycle(nsub) = nsub x [stream, out-of-box compaction, mass = 1, kick,
        stream] on the device, one haccsr_download (reference src/cpu/Particles.cxx:1176-1201)."""
        vmax = max(float(np.abs(self.pin[k]).max()) for k in ("vx", "vy", "vz"))
        # each half-stream moves the fastest particle 0.02 cells; the parity snapshots carry v = 0, then only the kicks move them
        pt = 0.02 / vmax if vmax > 0 else 0.01
        sub_args = (nsub, pt, [float(self.nglt)] * 3, self.lo, self.hi, self.flo, self.fhi, THETA, self.args.ppn, 1e-3)
        self.g.upload(self.pin)
        self.g.subcycle(*sub_args)                         # warm
        self.barrier()
        e = self.events(2)
        e[0].record(self.stream)
        pairs, ms_res = 0, 0.0
        for _ in range(reps):
            s = self.events(2)
            self.g.upload(self.pin)
            s[0].record(self.stream)
            sst = self.g.subcycle(*sub_args)
            s[1].record(self.stream)
            self.g.download(out=self.work)
            pairs += sst["pairs_evaluated"]
            self.torch.cuda.synchronize()
            ms_res += s[0].elapsed_time(s[1])
        e[1].record(self.stream)
        self.barrier()
        out = {"pairs": pairs, "ms": e[0].elapsed_time(e[1]), "ms_resident": ms_res, "reps": reps, "nsub": nsub}
        self.g.upload(self.pin)
        return out

    def e2e_long_step(self, nsub, reps):
        """Level-3 path: the particles never leave the GPU.  One long step = the particle side of the PM solve on the device
        (haccsr_cic: deposit, density grid to the host; haccsr_inverse_cic x 3: the three gradient grids from the host,
        Particles.cxx:589-714, mc3.cxx:304-379) around the short-range sub-cycle (haccsr_subcycle, nsub kicks).  Only the PM
        grids cross PCIe: nglt^3 floats down, 3 nglt^3 floats up per long step.  The FFT Poisson solve between deposit and
        gradient stays the reference's (src/dfft) and is not part of the timed region: the gradient grids are synthetic."""
        torch = self.torch
        ng = (self.nglt,) * 3
        ncell = self.nglt ** 3
        with NumaLocal(self.local):
            rho = torch.empty(ng, dtype=torch.float32).pin_memory().numpy()
            grads = [(torch.randn(ng, dtype=torch.float32) * 1e-4).pin_memory().numpy() for _ in range(3)]
        vmax = max(float(np.abs(self.pin[k]).max()) for k in ("vx", "vy", "vz"))
        pt = 0.02 / vmax if vmax > 0 else 0.01
        sub_args = (nsub, pt, [float(self.nglt)] * 3, self.lo, self.hi, self.flo, self.fhi, THETA, self.args.ppn, 1e-3)

        def step():
            self.g.cic(ng, 1.0, out=rho)
            for comp in range(3):
                self.g.inverse_cic(grads[comp], 1e-3, 1.0, comp)
            return self.g.subcycle(*sub_args)
        self.g.upload(self.pin)
        step()                                             # warm
        self.barrier()
        e = self.events(2)
        e[0].record(self.stream)
        pairs, ms_cic = 0, 0.0
        for _ in range(reps):
            c = self.events(3)
            c[0].record(self.stream)
            self.g.cic(ng, 1.0, out=rho)
            c[1].record(self.stream)
            for comp in range(3):
                self.g.inverse_cic(grads[comp], 1e-3, 1.0, comp)
            c[2].record(self.stream)
            pairs += self.g.subcycle(*sub_args)["pairs_evaluated"]
            self.torch.cuda.synchronize()
            ms_cic += c[0].elapsed_time(c[2])
        e[1].record(self.stream)
        self.barrier()
        out = {"pairs": pairs, "ms": e[0].elapsed_time(e[1]), "ms_pm_coupling": ms_cic / reps, "reps": reps, "nsub": nsub,
               "h2d": 3 * 4 * ncell, "d2h": 4 * ncell}
        self.g.upload(self.pin)
        return out

    def cic_block(self):
        """Device time of the deposit and of one interpolation with the grids already on the device (HBM-bound kernels)."""
        torch = self.torch
        import ctypes as C
        ng = (self.nglt,) * 3
        ncell = self.nglt ** 3
        grid = torch.randn(ng, dtype=torch.float32, device=self.dev)
        rho = torch.empty(ng, dtype=torch.float32, device=self.dev)
        ng3 = (C.c_int32 * 3)(*ng)
        lib, h = self.g.lib, self.g._h
        lib.haccsr_cic(h, ng3, 1.0, C.c_void_p(rho.data_ptr()), 1)
        lib.haccsr_inverse_cic(h, ng3, C.c_void_p(grid.data_ptr()), 1, 0.0, 1.0, 3)
        e = self.events(3)
        self.torch.cuda.synchronize()
        e[0].record(self.stream)
        for _ in range(3):
            lib.haccsr_cic(h, ng3, 1.0, C.c_void_p(rho.data_ptr()), 1)
        e[1].record(self.stream)
        for _ in range(3):
            lib.haccsr_inverse_cic(h, ng3, C.c_void_p(grid.data_ptr()), 1, 0.0, 1.0, 3)      # tau = 0: phi unchanged
        e[2].record(self.stream)
        self.torch.cuda.synchronize()
        t_cic, t_inv = e[0].elapsed_time(e[1]) / 3, e[1].elapsed_time(e[2]) / 3
        pk, _ = peaks()
        hbm = float(pk.get("hbm_gbs", 6650.0))
        # algorithmic bytes: deposit = 12 B of position per particle + 8 B accumulator zero + 8 B read + 4 B write per cell;
        # interpolation = 12 B position + 4 B read + 4 B write of the updated array per particle + 4 B per cell
        b_cic, b_inv = 12.0 * self.n + 20.0 * ncell, 20.0 * self.n + 4.0 * ncell
        return {"ms_cic": t_cic, "ms_inverse_cic": t_inv, "cic_GBps": b_cic / (t_cic * 1e-3) / 1e9, "inverse_cic_GBps": b_inv / (t_inv * 1e-3) / 1e9,
                "cic_frac_of_hbm": b_cic / (t_cic * 1e-3) / 1e9 / hbm, "inverse_cic_frac_of_hbm": b_inv / (t_inv * 1e-3) / 1e9 / hbm,
                "atomics_per_particle": 8, "grid": list(ng),
                "note": "haccsr_cic / haccsr_inverse_cic with device grids (Particles::cic / inverse_cic, Particles.cxx:589-714); the "
                        "deposit is bound by its 8 scattered 64-bit atomics per particle (L2), not by HBM"}

    def pcie(self):
        """Host<->device copy rates of the ten arrays with every rank copying at the same time."""
        self.barrier()
        e = self.events(3)
        e[0].record(self.stream)
        self.g.upload(self.pin)
        e[1].record(self.stream)
        self.g.download(out=self.work)
        e[2].record(self.stream)
        self.barrier()
        b = 42.0 * self.n
        return b / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9, b / (e[1].elapsed_time(e[2]) * 1e-3) / 1e9

    def culled(self, base_ms_force):
        """The same kick with warp-level culling toggled (bit-identical result, fewer executed flops)."""
        args = self.args
        self.g.set_culling(not args.cull)
        self.kick()
        other = [self.kick() for _ in range(3)]
        oc = self.kick(count_in_cutoff=True)
        # end to end through the facade path with culling on
        e2e = self.e2e_facade(2)
        self.g.set_culling(args.cull)
        self.g.upload(self.pin)
        oms = float(np.mean([o["ms_force"] for o in other]))
        on_ms, off_ms = (base_ms_force, oms) if args.cull else (oms, base_ms_force)
        return {"ms_force": on_ms, "ms_force_unculled": off_ms, "speedup_force": off_ms / on_ms,
                "ms_kick": float(np.mean([o["ms_total"] for o in other])),
                "pairs_force_law_frac": oc["pairs_force_law"] / max(oc["pairs_evaluated"], 1),
                "frac_executed_flop": (30.0 * oc["pairs_force_law"] + 9.0 * (oc["pairs_evaluated"] - oc["pairs_force_law"]))
                / (30.0 * max(oc["pairs_evaluated"], 1)),
                "e2e": {"value": e2e["pairs"] / (e2e["ms"] * 1e-3) / 1e9, "unit": "Ginteractions/s", "ms_per_step": e2e["ms"] / e2e["steps"],
                        "note": "haccsr_kick_host with culling %s, this rank" % ("off" if args.cull else "on")},
                "note": "warp-level early exit after the cutoff test; results bit-identical; off in the headline unless --cull"}

    def tuned(self, ppns, ref_pairs_incut):
        """Time to solution per kick for other leaf sizes (-N of the reference, src/simulation/MC3Options.cxx:91-136):
        the same physical in-cutoff pairs, fewer list pairs evaluated."""
        rows = []
        for ppn in ppns:
            self.kick(ppn=ppn)
            ks = [self.kick(ppn=ppn) for _ in range(3)]
            kc = self.kick(ppn=ppn, count_in_cutoff=True)
            ms = float(np.mean([k["ms_total"] for k in ks])); msf = float(np.mean([k["ms_force"] for k in ks]))
            ms_cull = None
            if self.args.arith != "x86":
                self.g.set_culling(not self.args.cull)
                self.kick(ppn=ppn)
                ms_cull = float(np.mean([self.kick(ppn=ppn)["ms_total"] for _ in range(3)]))
                self.g.set_culling(self.args.cull)
            rows.append({"ppn": ppn, "ms_kick": ms, "ms_kick_culling_%s" % ("off" if self.args.cull else "on"): ms_cull,
                         "ms_force": msf, "ms_build": float(np.mean([k["ms_build"] for k in ks])),
                         "pairs_evaluated": int(kc["pairs_evaluated"]), "pairs_in_cutoff": int(kc["pairs_in_cutoff"]),
                         "in_cutoff_Gpairs_per_s": kc["pairs_in_cutoff"] / (ms * 1e-3) / 1e9,
                         "roofline_frac": FLOP_PER_PAIR * kc["pairs_evaluated"] / (msf * 1e-3) / 1e12 / self.fp32_peak,
                         "levels": int(kc["levels"]), "mean_ppn": kc["mean_ppn"]})
        return rows


def block_for_state(B, args, steps, warmup, world, sampler=None):
    """Everything measured on the currently loaded snapshot; returns (local dict, reducible numbers)."""
    torch = B.torch
    acc = B.resident(steps, warmup, sampler)
    e2e = B.e2e_facade(max(1, min(steps, 3)))
    B.g.upload(B.pin)
    sub = B.e2e_subcycle(args.subcycle, 2 if steps > 2 else 1) if args.subcycle > 0 else None
    lng = B.e2e_long_step(args.subcycle, 2 if steps > 2 else 1) if args.subcycle > 0 else None
    h2d, d2h = B.pcie()
    stc = B.kick(count_in_cutoff=True)     # one untimed pass that also counts the pairs inside the cutoff
    tv = [acc["ms"], e2e["ms"], sub["ms"] if sub else 0.0, -h2d, -d2h, lng["ms"] if lng else 0.0]
    sv = [float(acc["pairs"]), float(e2e["pairs"]), float(sub["pairs"]) if sub else 0.0, float(acc["launches"]),
          float(lng["pairs"]) if lng else 0.0]
    tvt = torch.tensor(tv, device=B.dev, dtype=torch.float64)
    svt = torch.tensor(sv, device=B.dev, dtype=torch.float64)
    if B.dist is not None:
        B.dist.all_reduce(tvt, op=B.dist.ReduceOp.MAX)
        B.dist.all_reduce(svt, op=B.dist.ReduceOp.SUM)
    tv = [float(t) for t in tvt.cpu()]
    sv = [float(t) for t in svt.cpu()]
    st, n = acc["st"], B.n
    pk, pk_kind = peaks()
    achieved = FLOP_PER_PAIR * acc["pairs"] / (acc["ms_force"] * 1e-3) / 1e12          # this rank's force kernel
    L = int(st["levels"])
    alg_bytes = n * (28.0 * L + 84.0)                 # L * (16 B box/COM read + 4 B key read + 8 B index r/w) + 84 B
    ms_build = acc["ms_build"] / steps
    hbm = float(pk.get("hbm_gbs", 6650.0))
    out = {
        "value": sv[0] / (tv[0] * 1e-3) / 1e9, "ms_per_step": tv[0] / steps,
        "e2e": {"value": sv[1] / (tv[1] * 1e-3) / 1e9, "unit": "Ginteractions/s", "h2d_bytes_per_step": 42 * n, "d2h_bytes_per_step": 42 * n,
                "ms_per_step": tv[1] / e2e["steps"], "path": "haccsr_kick_host (facade constructor: all ten arrays up and down every kick)",
                "pcie_gbs_per_gpu_min_over_ranks": {"h2d": -tv[3], "d2h": -tv[4], "note": "haccsr_upload / haccsr_download of the ten arrays, all ranks copying at once"}},
        "gpu_launches": int(sv[3]),
        "roofline": {"bound": "fp32", "kernel": "k_force", "achieved": achieved, "peak": B.fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / B.fp32_peak, "traffic": ncu_traffic("k_force", args.np_side, B.state, args.arith),
                     "peak_source": B.peak_source, "flop_per_interaction": FLOP_PER_PAIR,
                     "ms_per_launch": acc["ms_force"] / max(acc["force_launches"], 1)},
        "roofline_build": {"bound": "hbm", "kernel": "tree build (k_split_pass per level + k_gather)",
                           "achieved": alg_bytes / (ms_build * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                           "frac": alg_bytes / (ms_build * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg_bytes,
                           "bytes_per_particle": "28 * levels + 84 (SURVEY.md 8(d))", "levels": L, "ms": ms_build,
                           "traffic": ncu_traffic("build", args.np_side, B.state, args.arith), "peak_source": pk_kind + " hbm_gbs"},
        "phases_ms": {"build": ms_build, "walk": acc["ms_walk"] / steps, "force": acc["ms_force"] / steps},
        "particles_per_gpu": n, "pairs_per_particle": st["pairs_evaluated"] / n,
        "pairs_in_cutoff_frac": stc["pairs_in_cutoff"] / max(stc["pairs_evaluated"], 1),
        "tree": {"nodes": st["nodes"], "leaves": st["leaves"], "mean_ppn": st["mean_ppn"], "levels": st["levels"],
                 "max_list": st["max_list"], "pseudo_particles": st["pseudo_particles"]},
    }
    if sub:
        kicks = sub["reps"] * sub["nsub"]
        out["e2e"]["subcycle"] = {
            "value": sv[2] / (tv[2] * 1e-3) / 1e9, "unit": "Ginteractions/s", "nsub": sub["nsub"], "ms_per_kick": tv[2] / kicks,
            "h2d_bytes_per_kick": 42 * n // sub["nsub"], "d2h_bytes_per_kick": 42 * n // sub["nsub"],
            "ms_per_kick_resident_this_rank": sub["ms_resident"] / kicks,
            "path": "haccsr_upload + haccsr_subcycle(nsub) + haccsr_download (INTEGRATION.md Level 2 = Particles::subCycle, Particles.cxx:1176-1201)"}
    if lng:
        kicks = lng["reps"] * lng["nsub"]
        out["e2e"]["long_step"] = {
            "value": sv[4] / (tv[5] * 1e-3) / 1e9, "unit": "Ginteractions/s", "nsub": lng["nsub"], "ms_per_kick": tv[5] / kicks,
            "h2d_bytes_per_step": lng["h2d"], "d2h_bytes_per_step": lng["d2h"], "ms_pm_coupling_this_rank": lng["ms_pm_coupling"],
            "path": "particles resident across the long step (INTEGRATION.md Level 3): haccsr_cic (density grid to the host) + "
                    "3 x haccsr_inverse_cic (gradient grids from the host) + haccsr_subcycle(nsub); only the PM grids cross PCIe; "
                    "the reference's FFT solve between deposit and gradient is outside the timed region (synthetic gradient grids)"}
    return out, acc, stc


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nglt = args.np_side + 2 * GHOST

    def config_for(state):
        workload = "np=%d^3 alive per GPU + %d-cell overload shell (%d^3 grid units), %s Zel'dovich snapshot, ppn=%d, theta=%.1f, poly5" % (
            args.np_side, GHOST, nglt, STATE_NAME[state], args.ppn, THETA)
        return {"workload": workload, "np_side": args.np_side, "ppn": args.ppn, "theta": THETA, "rsm": RSM,
                "force_law": "poly5", "arithmetic": args.arith, "state": state, "l2": "inputs larger than L2 (no flush needed)",
                "parallelism": "1 sub-volume per GPU, no data-path collective"}
    config = config_for(args.state)

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, config)

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; libhaccsr has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = Bench(args, rank, world, local, dist)
    pk, pk_kind = peaks()
    props = torch.cuda.get_device_properties(local)
    sm_max = float(pk.get("sm_max_mhz", 1965.0))
    B.fp32_peak = props.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12     # TFLOP/s
    B.peak_source = "%d SMs x 128 lanes x 2 x %.0f MHz (%s sm_max_mhz)" % (props.multi_processor_count, sm_max, pk_kind)

    # ---- headline state ------------------------------------------------------------------------------
    B.state = args.state
    B.load(args.state)
    sampler = ClockSampler(local)
    head, acc, stc = block_for_state(B, args, args.steps, args.warmup, world, sampler)
    culled = B.culled(acc["ms_force"] / args.steps) if args.arith != "x86" else None
    cic = B.cic_block()
    tuned = None
    if args.tune_ppn:
        rows = B.tuned([int(t) for t in args.tune_ppn.split(",")], stc["pairs_in_cutoff"])
        base = {"ppn": args.ppn, "ms_kick": head["ms_per_step"], "ms_force": acc["ms_force"] / args.steps, "ms_build": acc["ms_build"] / args.steps,
                "pairs_evaluated": int(stc["pairs_evaluated"]), "pairs_in_cutoff": int(stc["pairs_in_cutoff"]),
                "in_cutoff_Gpairs_per_s": stc["pairs_in_cutoff"] / (head["ms_per_step"] * 1e-3) / 1e9,
                "roofline_frac": head["roofline"]["frac"], "levels": int(stc["levels"]), "mean_ppn": stc["mean_ppn"]}
        best = min(rows + [base], key=lambda r: r["ms_kick"])
        tuned = {"sweep": [base] + rows, "best_ppn": best["ppn"], "ms_kick_best": best["ms_kick"],
                 "speedup_over_ppn_%d" % args.ppn: base["ms_kick"] / best["ms_kick"],
                 "note": "time to solution of one kick (build + walk + force) at other leaf sizes (-N, reference src/simulation/"
                         "MC3Options.cxx:91-136); the kicked set near the faces depends on leaf geometry, so parity at a tuned ppn is "
                         "gated on particles inside the force box (tests/test_gpu_parity.py::test_tuned_leaf_size_against_reference_at_512); "
                         "the headline stays at the reference's shipped ppn"}
    p_head, nglt_head = B.p, B.nglt
    if args.no_cpu_baseline or world != 1:
        p_head = None

    # ---- the clustered target state as a side block (configs[2]) -------------------------------------------
    clustered = None
    if args.state == "uniform" and not args.no_clustered_block:
        B.state = "clustered"
        B.load("clustered")
        steps_c = max(1, min(args.steps, 5))
        clustered, acc_c, _ = block_for_state(B, args, steps_c, 3, world)
        clustered["config"] = config_for("clustered")
        clustered["steps"] = steps_c
        if args.arith != "x86":
            clustered["culled"] = B.culled(acc_c["ms_force"] / steps_c)
    B.g.close()
    B.g = None
    del B.pin, B.work, B.p
    torch.cuda.empty_cache()

    # ---- overload refresh of ONE global snapshot cut into the ranks' sub-volumes (configs[3]; 1x1x1 / 2x1x1 / 2x2x1 / 2x2x2) ------
    refresh = None
    if not args.no_refresh_block and world in (1, 2, 4, 8):
        from tools import decomposed
        try:
            refresh = decomposed.refresh_block(args.np_side, GHOST, rank, world, local, dist)
        except Exception as ex:      # a side block must not take the headline down
            refresh = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    line = {"metric": "short-range G interactions/s", "value": head.pop("value"), "unit": "Ginteractions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config}
    line.update(head)
    line["clocks"] = sampler.summary()
    if culled:
        line["culled"] = culled
    if tuned:
        line["tuned"] = tuned
    line["cic"] = cic
    if clustered:
        line["clustered"] = clustered
    if refresh:
        line["refresh"] = refresh
    if args.cull:
        # with culling the kernel executes 30 flop only for the pairs that reach the force law and 9 (three differences,
        # the r2 chain, the softening add) for the rest: report the executed rate next to the algorithmic one
        ex = (30.0 * stc["pairs_force_law"] + 9.0 * (stc["pairs_evaluated"] - stc["pairs_force_law"])) / max(stc["pairs_evaluated"], 1)
        line["roofline"]["executed_flop_per_interaction"] = ex
        line["roofline"]["frac_executed"] = line["roofline"]["frac"] * ex / FLOP_PER_PAIR
        line["config"]["culling"] = "on"
    if not args.no_cpu_baseline and world == 1:      # the contract: rank 0 at N = 1 only (the reference arm covers every N)
        try:
            info = run_reference_sample(p_head, nglt_head, 112, args.ppn)
        except Exception as ex:   # the baseline is a reported companion number; never fail the GPU line
            info = None
            line["cpu_baseline_error"] = repr(ex)
        if info:
            line["cpu_baseline"] = {
                "value": info["pairs"] / info["seconds"] / 1e9, "unit": "Ginteractions/s", "cores": info["cores"],
                "kind": "reference",
                "sample": "SAMPLED: %d^3-cell cut-out (%d particles, %d pairs) of rank 0's snapshot, full RCBMonopoleForceTree ctor%s, %.1f s" % (
                    info["side"], info["particles"], info["pairs"], info["note"], info["seconds"])}
    # the JSON line is the last thing on stdout: NCCL (NCCL_DEBUG) may still write while the process group is torn down
    if dist is not None:
        dist.destroy_process_group()
    sys.stdout.flush()
    print(json.dumps(line))
    sys.stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
