#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 NUFFT backend (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3_t1|c3_t2|c3_t1_clustered]
    python bench.py --impl reference ...        # CPU arm: the oracle port on the host cores

A *step* is one pass of the hot path over one batch of synthetic input: one custom call as
jax-finufft issues it (run_nufft, lib/kernels.cc.cu:25-92) = bin-sort the points (setpts) +
execute one 3-D transform.  Default workload = BASELINE.json configs[2]: 3-D type 1, M=1e8
uniform random points, N=256^3 modes, eps=1e-6, complex64 (``c3_t1``); type 2 on the same
points and the clustered distribution are measured too and reported under ``also``.

N>1 (torchrun): the path shards over independent transforms (n_tot / vmapped batch): every
rank runs its own M=1e8 transform on its own GPU, no data-path collective -> weak scaling;
``value`` = N*M / max-over-ranks time.  The sharded configs BASELINE.json names are timed in the
same run and reported under ``also``: ``c4_stacked`` (config 4: 64 stacked 2-D type-1 transforms
sharing M=1e7 points, N=1024^2, the stack split across the ranks -- strong scaling, the N=1 run is
its own baseline) and ``c3_t1_points_sharded`` (ONE 3-D type 1 with its points split by range).
``also.ref_gpu`` times the unmodified reference cuFINUFFT (oracle/_ref, when it travelled to the
box) on the same tensors with the protocol of V/perftest/cuda/cuperftest.cu:183-303.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    #  name              type  M      N (JAX order)     eps   dist
    "c3_t1": (1, 10 ** 8, (256, 256, 256), 1e-6, "uniform"),
    "c3_t2": (2, 10 ** 8, (256, 256, 256), 1e-6, "uniform"),
    "c3_t1_clustered": (1, 10 ** 8, (256, 256, 256), 1e-6, "clustered"),
    "c3_t2_clustered": (2, 10 ** 8, (256, 256, 256), 1e-6, "clustered"),
    "small_t1": (1, 10 ** 6, (64, 64, 64), 1e-6, "uniform"),   # dev only
}
C4 = dict(n_transf=64, M=10 ** 7, nm=(1024, 1024), eps=1e-6)   # BASELINE.json configs[3]
METRIC = "NU points/sec, 3-D type-{t} (eps=1e-6, complex64, M=1e8, N=256^3), setpts+execute per step"
UNIT = "NU points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_t1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip also/e2e/cpu_baseline legs (profiling runs)")
    ap.add_argument("--cpu-sample", type=int, default=20_000_000, help="points of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-full", dest="cpu_full", action="store_false",
                    help="--impl reference: time --cpu-sample points per step instead of the full M (CPU-only smoke runs)")
    return ap.parse_args()


def make_config(workload, world):
    """The workload description both arms print (same keys, same values)."""
    typ, M, nm, eps, dist = WORKLOADS[workload]
    return {"workload": workload, "type": typ, "M_per_gpu": M, "N": list(nm), "eps": eps, "points": dist,
            "seed": "1+rank", "n_gpus": world}


def src_sha16():
    """Hash of the library's sources (csrc/ + include/): stamps which code an ncu capture measured."""
    import glob
    import hashlib
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "jax_finufft_b200", "csrc", "*.*")) + glob.glob(os.path.join(ROOT, "include", "*.h")))
    for f in files:
        if os.path.isfile(f):
            h.update(os.path.basename(f).encode())
            h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


# ----------------------------------------------------------------------------- synthetic inputs
def make_points(M, dist, nf, device, seed):
    """uniform = U[-pi,pi)^3 (V/perftest/cuda/cuperftest.cu:197-220);
    clustered = iid uniform in the corner box [-pi, -pi + 8h)^3, h = 2pi/nf (8 fine cells per dim)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    pts = []
    for d in range(3):
        u = torch.rand(M, device=device, generator=g)
        if dist == "uniform":
            pts.append((u * 2 - 1) * np.pi)
        else:
            pts.append(-np.pi + u * (8 * 2 * np.pi / nf))
    return pts, g


def make_complex(shape, device, g):
    import torch

    return torch.complex(torch.rand(shape, device=device, generator=g) * 2 - 1,
                         torch.rand(shape, device=device, generator=g) * 2 - 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _oracle_threads():
    import oracle

    try:  # all the host cores this process may use, whatever OMP_NUM_THREADS the launcher exported
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        oracle.set_num_threads(os.cpu_count() or 1)
    return oracle


def _cpu_runner(typ, nm, eps, M, seed=1):
    """-> run(n): the full N=256^3 oracle pipeline on the first n of M seeded points."""
    oracle = _oracle_threads()
    rng = np.random.default_rng(seed)
    x = rng.uniform(-np.pi, np.pi, size=(3, M))
    nmx = tuple(nm[::-1])
    if typ == 1:
        d = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
        return lambda n: oracle.nufft1(nmx, d[:n], *x[:, :n], eps=eps, prec=1)
    d = rng.uniform(-1, 1, nm) + 1j * rng.uniform(-1, 1, nm)
    return lambda n: oracle.nufft2(d, *x[:, :n], eps=eps, prec=1)


def _clock(run, n):
    t0 = time.perf_counter()
    run(n)
    return time.perf_counter() - t0


def cpu_port_sample(typ, M_full, nm, eps, sample):
    """cpu_baseline of the GPU arm: oracle/nufft_oracle.c (OpenMP, all host threads) on a BOUNDED
    sample -- the full N=256^3 pipeline on `sample` points and, for the M-independent cost (FFT +
    deconvolve), on sample/10 points, both warm; the per-point cost is the slope between the two,
    extrapolated linearly in M (flagged).  The reference arm (--impl reference) runs the full M."""
    import oracle

    run = _cpu_runner(typ, nm, eps, sample)
    small = max(1000, sample // 10)
    _clock(run, small)                            # first touch of the grids, thread pool, FFT tables
    t_small = min(_clock(run, small), _clock(run, small))
    _clock(run, sample)
    t_step = _clock(run, sample)
    per_pt = max(t_step - t_small, 0.0) / (sample - small)
    t_fixed = max(t_small - per_pt * small, 0.0)
    t_full = t_fixed + per_pt * M_full
    return {
        "value": M_full / t_full, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "extrapolated": True,
        "sample": (f"oracle/nufft_oracle.c (float64 restatement, OpenMP x{oracle.num_threads()}): the full N=256^3 pipeline on "
                   f"{sample} points ({t_step:.2f} s; {t_small:.2f} s on {small} points => {per_pt * 1e9:.0f} ns/point + "
                   f"{t_fixed:.2f} s M-independent FFT/deconvolve), extrapolated linearly in M to M={M_full}: {t_full:.1f} s. "
                   "`bench.py --impl reference` times the full M.  The reference CPU FINUFFT (xsimd+FFTW) cannot be built "
                   "offline (SURVEY.md 8c)"),
        "ms_per_step_extrapolated": t_full * 1e3,
    }


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL (NCCL_DEBUG=VERSION on some boxes) and other
    native code print to fd 1, so fd 1 is pointed at stderr for the whole run and the JSON line is
    written to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main_reference(a):
    """Reference arm: the CPU implementation of the path on the host cores (the oracle port; the
    reference's own FINUFFT cannot be built offline), all host threads, on the FULL workload --
    M points per step, nothing extrapolated.  A step costs tens of seconds, so steps/warm-up are
    capped (the line states what ran) to keep the run within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    typ, M, nm, eps, dist = WORKLOADS[a.workload]
    steps = max(1, min(a.steps, 2 if a.cpu_full else 3))
    warm = 1
    M_run = M if a.cpu_full else min(M, a.cpu_sample)
    run = _cpu_runner(typ, nm, eps, M_run)
    for _ in range(warm):
        _clock(run, M_run if not a.cpu_full else max(1000, M_run // 10))   # tables, thread pool, first touch
    ts = [_clock(run, M_run) for _ in range(steps)]
    t_step = statistics.mean(ts)
    cfg = make_config(a.workload, a.gpus)
    cb = {"value": M_run / t_step, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "extrapolated": False,
          "sample": (f"oracle/nufft_oracle.c (float64 restatement, OpenMP x{oracle.num_threads()}): {steps} timed steps of the full "
                     f"N=256^3 pipeline on M={M_run} points, {t_step:.2f} s/step (warm-up: one pass on "
                     f"{M_run if not a.cpu_full else max(1000, M_run // 10)} points); the reference CPU FINUFFT (xsimd+FFTW) cannot "
                     "be built offline (SURVEY.md 8c)")}
    if M_run != M:
        cfg["M_per_gpu"] = M_run
        cb["sample"] += f"; --no-cpu-full: bounded to {M_run} of {M} points"
    line = {
        "impl": "reference", "metric": METRIC.format(t=typ), "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------- reference GPU kernels
def ref_gpu_leg(workload, timed):
    import torch
    try:
        from oracle import ref_cufinufft as ref
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": f"oracle.ref_cufinufft: {ex}"}
    if not ref.available():
        return {"unavailable": "oracle/_ref/libcufinufft_ref.so not present (built by oracle/Makefile.ref where /root/reference exists)"}
    res = {"library": "oracle/_ref/libcufinufft_ref.so = vendor/finufft @ the reference's pin, unmodified, sm_100",
           "protocol": "V/perftest/cuda/cuperftest.cu:183-303"}
    for name in ("c3_t1", "c3_t2"):
        try:
            typ, M, nm, eps, distn, pts, data, step = workload(name)
            isign = 1 if typ == 1 else -1
            out = torch.empty((1,) + tuple(nm) if typ == 1 else (1, M), dtype=torch.complex64, device=data.device)

            def per_call():
                r = ref.RefPlan(typ, nm[::-1], n_trans=1, eps=eps, isign=isign)
                r.setpts(pts[2], pts[1], pts[0])
                r.execute(data[None], out=out)
                r.destroy()

            r = ref.RefPlan(typ, nm[::-1], n_trans=1, eps=eps, isign=isign)

            def kept():
                r.setpts(pts[2], pts[1], pts[0])
                r.execute(data[None], out=out)

            ms_kept, _ = timed(kept, 3, 1)
            r.destroy()
            ms_call, _ = timed(per_call, 3, 1)
            ms_ours, _ = timed(step, 3, 2)
            res[name] = {"ref_ms_plan_kept": ms_kept, "ref_ms_plan_per_call": ms_call, "ours_ms": ms_ours,
                         "speedup_plan_kept": ms_kept / ms_ours, "speedup_plan_per_call": ms_call / ms_ours}
            del pts, data, step, out
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            res[name] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    return res


# ----------------------------------------------------------------------------- GPU arm
def main_ours(a):
    import torch
    import torch.distributed as dist

    import jax_finufft_b200 as J
    from jax_finufft_b200 import _lib
    from jax_finufft_b200.plan import Plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for --impl ours")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    L.b2n_set_setpts_cache(0)  # every timed step re-sorts its points (B2N_SETPTS_CACHE in the environment is ignored)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, K, W):
        for _ in range(W):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.b2n_launch_count()
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / K, (L.b2n_launch_count() - n0)

    def workload(name):
        typ, M, nm, eps, distn = WORKLOADS[name]
        nf = 2 * nm[-1]
        pts, g = make_points(M, distn, nf, dev, seed=1 + rank)
        data = make_complex((M,) if typ == 1 else nm, dev, g)
        # JAX order: the LAST point array is the backend's x (lowering.py:96-105)
        if typ == 1:
            step = lambda: J.nufft1(nm, data, *pts, eps=eps, iflag=1)
        else:
            step = lambda: J.nufft2(data, *pts, eps=eps, iflag=-1)
        return typ, M, nm, eps, distn, pts, data, step

    typ, M, nm, eps, distn, pts, data, step = workload(a.workload)
    nftot = 1
    for n in nm:
        nftot *= 2 * n

    # ---- headline: device-resident inputs, K steps, device events, max over ranks
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed(step, a.steps, a.warmup)
    clocks = sampler.stop() if sampler else None
    value = world * M / (ms * 1e-3)

    line = {
        "metric": METRIC.format(t=typ), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(a.workload, world),
        "notes": {"step": "b2n_run = setpts (bin-sort) + execute, plan cached",
                  "l2": "inputs (2.0 GB points+strengths, 1.07 GB fine grid) exceed the 126 MB L2; no explicit flush",
                  "sharding": "independent transforms per GPU (n_tot split), no data-path collective" if world > 1 else "single GPU"},
        "gpu_launches": int(launches), "clocks": clocks,
    }

    if not a.no_extras:
        # ---- per-stage device times + roofline of the dominant kernel (CUDA events on the plan's
        # stream, inside the library: b2n_plan_timings)
        if rank == 0:
            p = Plan(typ, nm[::-1], eps=eps, isign=1 if typ == 1 else -1, debug=1)
            out = None
            KR = max(3, min(a.steps, 5))
            for it in range(KR + 1):
                p.setpts(pts[2], pts[1], pts[0])
                out = p.execute(data[None] if typ == 1 else data[None], out=out)
                if it == 0:
                    p.timings()  # drop warm-up
            st = {k: v / KR for k, v in p.timings().items()}
            p.destroy()
            del out
            kern = "spread" if typ == 1 else "interp"
            alg_bytes = (4 * 3 + 12) * M + 8 * nftot          # (4d+12)*M + 8*nf
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            peak = float(peaks.get("hbm_gbs", 6650.0))
            achieved = alg_bytes / (st[kern] * 1e-3) / 1e9
            # measured DRAM bytes of that kernel: one `ncu --set full` capture per change
            # (tools/gpu_round.sh -> tools/traffic_stamp.py), stamped with the hash of the sources it
            # was taken on; a capture of other sources is reported but flagged
            traffic, capture = None, None
            tf = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tf):
                ent = json.load(open(tf)).get(f"{a.workload}:{kern}")
                if isinstance(ent, dict):
                    traffic = ent.get("bytes")
                    capture = {"tag": ent.get("tag"), "kernel": ent.get("kernel"), "src_sha16": ent.get("src_sha16"),
                               "same_sources_as_this_run": ent.get("src_sha16") == src_sha16()}
            line["roofline"] = {"bound": "hbm", "kernel": "k_swr2_spread<7>" if typ == 1 else "k_swr2_interp<7>",
                                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "traffic": traffic, "traffic_capture": capture,
                                "algorithmic_bytes": alg_bytes, "kernel_ms": st[kern],
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                                "note": "3-D ns=7 spread/interp is FP32-pipe/issue-bound (343 complex cell updates per point, kept in "
                                        "registers), see DESIGN.md 4.2; HBM fraction reported as the contract asks",
                                # the binding unit: useful FMAs (2 * ns^3 per point) against 148 SMs x 128 FP32 lanes
                                "fp32": {"useful_fma_per_point": 686,
                                         "achieved_tfma_s": 686 * M / (st[kern] * 1e-3) / 1e12,
                                         "peak_tfma_s": 148 * 128 * float((clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6 / 1e12}}
            fp = line["roofline"]["fp32"]
            fp["frac"] = fp["achieved_tfma_s"] / fp["peak_tfma_s"]
            line["stages_ms"] = st
        barrier()

        # ---- e2e: the C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region
        host_pts = [torch.empty(M, dtype=torch.float32).pin_memory() for _ in range(3)]
        for h, d_ in zip(host_pts, pts):
            h.copy_(d_)
        host_in = torch.empty(data.shape, dtype=torch.complex64).pin_memory()
        host_in.copy_(data)
        out_shape = nm if typ == 1 else (M,)
        host_out = torch.empty(out_shape, dtype=torch.complex64).pin_memory()
        o = _lib.default_opts()
        o.upsampfac = 2.0
        n_k = (C.c_int64 * 3)(*nm[::-1])
        pp = (C.c_void_p * 3)(host_pts[2].data_ptr(), host_pts[1].data_ptr(), host_pts[0].data_ptr())

        def e2e_step():
            rc = L.b2n_run_host(typ, 3, 0, eps, 1 if typ == 1 else -1, 1, 1, M, n_k, C.byref(o),
                                C.c_void_p(host_in.data_ptr()), pp, None, C.c_void_p(host_out.data_ptr()))
            if rc > 1:
                raise RuntimeError(f"b2n_run_host failed: {rc}")

        KE = max(2, min(a.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(KE):
            e2e_step()
        torch.cuda.synchronize()
        t_e2e = max_over_ranks((time.perf_counter() - t0) / KE)
        h2d = sum(h.numel() * 4 for h in host_pts) + host_in.numel() * 8
        d2h = host_out.numel() * 8
        line["e2e"] = {"value": world * M / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": t_e2e * 1e3, "steps": KE, "api": "b2n_run_host (include/b200nufft.h), pinned host buffers"}
        del host_pts, host_in, host_out

        # ---- the other halves of the BASELINE metric (type 2; clustered points), fewer steps
        also = {}
        del pts, data, step
        torch.cuda.empty_cache()
        for name in ("c3_t1", "c3_t2", "c3_t1_clustered", "c3_t2_clustered"):
            if name == a.workload or not a.workload.startswith("c3"):
                continue
            w = workload(name)
            ms2, _ = timed(w[-1], 3, 2)
            also[name] = {"value": world * w[1] / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "steps": 3}
            del w
            torch.cuda.empty_cache()
        # ---- the same step when the points do not change between calls (solver
        # iterations on a fixed trajectory; forward + JVP + VJP of one step) with the setpts cache
        # on -- the bin-sort is replaced by a signature pass.  NOT the headline: the headline and
        # every other figure in this line re-sort on every step, as the reference does.
        if a.workload.startswith("c3"):
            L.b2n_set_setpts_cache(1)
            for name in ("c3_t1", "c3_t2"):
                w = workload(name)
                ms2, _ = timed(w[-1], 3, 2)
                also[name + "_fixed_points_setpts_cache"] = {"value": world * w[1] / (ms2 * 1e-3), "unit": UNIT,
                                                             "ms_per_step": ms2, "steps": 3}
                del w
                torch.cuda.empty_cache()
            L.b2n_set_setpts_cache(0)
            L.b2n_cache_clear()
        line["also"] = also

        # ---- BASELINE config 4: 64 stacked 2-D type-1 transforms sharing M=1e7 points, N=1024^2, the
        # stack split across the ranks (points replicated, every rank repeats the bin-sort, no
        # collective: ref README.md:364-367).  STRONG scaling: total work fixed, the N=1 run is the
        # baseline the driver's 1/2/4/8 sweep divides by.
        if a.workload == "c3_t1":
            from jax_finufft_b200 import parallel as P
            try:
                g4 = torch.Generator(device=dev).manual_seed(3)   # same points and strengths on every rank
                p4 = [(torch.rand(C4["M"], device=dev, generator=g4) * 2 - 1) * np.pi for _ in range(2)]
                c4 = make_complex((C4["n_transf"], C4["M"]), dev, g4)
                fn = lambda: P.nufft1_stacked(C4["nm"], c4, *p4, gather=False, eps=C4["eps"], iflag=1)
                ms4, _ = timed(fn, 5, 3)
                nloc = len(range(*P.shard_range(C4["n_transf"], world, rank)))
                also["c4_stacked"] = {"value": C4["n_transf"] * C4["M"] / (ms4 * 1e-3), "unit": "NU point-transforms/s",
                                      "ms_per_step": ms4, "steps": 5, "scaling": "strong", "n_transf": C4["n_transf"],
                                      "n_transf_per_gpu": nloc, "M": C4["M"], "N": list(C4["nm"]), "eps": C4["eps"],
                                      "step": "setpts (bin-sort, repeated per rank) + execute of this rank's share of the stack"}
                del p4, c4, fn
            except Exception as ex:  # noqa: BLE001  (auxiliary leg: must not cost the headline line)
                also["c4_stacked"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            torch.cuda.empty_cache()
            L.b2n_cache_clear()

        # ---- BASELINE config 5 sharded (ref tests/sharding_test.py:244-290): 3-D type 3,
        # 1e7 sources split across the ranks by index range, 1e7 targets in [-64,64)^3 replicated, the
        # partial sums all-reduced.  Strong scaling; N=1 is the plain single-GPU type 3.
        if a.workload == "c3_t1":
            from jax_finufft_b200 import parallel as P
            try:
                g5 = torch.Generator(device=dev).manual_seed(4)
                M5 = 10 ** 7
                x5 = [(torch.rand(M5, device=dev, generator=g5) * 2 - 1) * np.pi for _ in range(3)]
                s5 = [(torch.rand(M5, device=dev, generator=g5) * 2 - 1) * 64.0 for _ in range(3)]
                c5 = make_complex((M5,), dev, g5)
                lo, hi = P.shard_range(M5, world, rank)
                xl, cl = [t[lo:hi].contiguous() for t in x5], c5[lo:hi].contiguous()
                fn = lambda: P.nufft3_sharded_sources(cl, *xl, *s5, eps=1e-6, iflag=-1)
                ms5, _ = timed(fn, 3, 2)
                also["c5_t3_sources_sharded"] = {"value": M5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5, "steps": 3,
                                                 "scaling": "strong", "M_sources": M5, "N_targets": M5, "eps": 1e-6,
                                                 "targets": "[-64,64)^3 replicated", "collective": "all_reduce of the 1e7 targets (80 MB)"}
                del x5, s5, c5, xl, cl, fn
            except Exception as ex:  # noqa: BLE001
                also["c5_t3_sources_sharded"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            torch.cuda.empty_cache()
            L.b2n_cache_clear()

        # ---- the kernel to beat: the UNMODIFIED reference cuFINUFFT (oracle/_ref, built from
        # /root/reference by oracle/Makefile.ref) on the same B200 and the same tensors, timed the way
        # V/perftest/cuda/cuperftest.cu:183-303 does (CUDA events; setpts + execute with the plan kept)
        # and the way jax-finufft drives it (makeplan + setpts + execute + destroy per call,
        # lib/kernels.cc.cu:49-92).  Checker/baseline only: nothing of it is on our path.
        if rank == 0 and world == 1 and a.workload.startswith("c3"):
            also["ref_gpu"] = ref_gpu_leg(workload, timed)

        # ---- N > 1: the one path with a real exchange step: ONE 3-D type 1 with its
        # M points split across the ranks by index range (strong scaling of a single transform).
        # slab = spatial split (points exchanged by z-slab, spread into slab + halo, halo planes to the
        # neighbours); reduce_scatter = native path (private fine grids summed by NCCL reduce-scatter over
        # NVLink, slab/pencil FFT divided by N); psum = what jax-finufft gets from shard_map
        # (every rank runs the full FFT, all-reduce of the output modes).
        if world > 1 and a.workload == "c3_t1":
            from jax_finufft_b200 import parallel as P
            Ml = M // world
            w = workload("c3_t1")
            lp = [p_[:Ml].clone() for p_ in w[5]]
            lc = w[6][:Ml].clone()
            del w
            torch.cuda.empty_cache()
            # (auxiliary legs: a failure here is recorded, it must not cost the headline line)
            def sharded_leg(nmx, K, W):
                res = {}
                for mode in ("slab", "reduce_scatter", "psum"):
                    torch.cuda.empty_cache()
                    try:
                        fn = lambda: P.nufft1_sharded_points(nmx, lc, *lp, combine=mode, gather=False, eps=eps, iflag=1)
                        ms_s, _ = timed(fn, K, W)
                        res[mode] = {"value": world * Ml / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s,
                                     "M_total": world * Ml, "N": list(nmx), "scaling": "strong"}
                    except Exception as ex:  # noqa: BLE001
                        res[mode] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
                    L.b2n_cache_clear()
                    for sp in list(P._SPREAD_PLANS.values()):
                        sp.destroy()
                    P._SPREAD_PLANS.clear()
                return res

            line["also"]["c3_t1_points_sharded"] = sharded_leg(nm, 3, 3)
            # the same three paths where the uniform-grid work dominates: N = 512^3 (fine grid 1024^3 =
            # 8.6 GB, FFT ~ 9 ms on one GPU).  psum repeats that FFT on every rank and reduce_scatter
            # moves the whole private grid; the slab split divides both by the number of ranks.
            line["also"]["t1_n512_points_sharded"] = sharded_leg((512, 512, 512), 3, 3)
            del lp, lc

        # ---- CPU baseline: the oracle port on the host cores (rank 0, N=1 only)
        if rank == 0 and world == 1:
            line["cpu_baseline"] = cpu_port_sample(typ, M, nm, eps, a.cpu_sample)

    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
