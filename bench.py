#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: query points / second for the fused
approx_posterior + posterior_mean + MaxVar evaluation (+ acquisition argmax) at N=1000 training points, d=5.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): SURVEY.md section 8(d) config C3 -- d=5, p=4 outputs, N=1000, ARD Matern-5/2, seeded synthetic
problem; every GPU evaluates its contiguous shard of the 1e8 Philox candidates: 1.25e7 candidates per GPU per step (weak
scaling: the full 1e8 set at 8 GPUs).  A step = ONE pass of the hot path over the rank's shard:
  value : candidates already resident in HBM (device pointers), outputs (log_approx_post, log_post_mean, log_post_var)
          written to HBM, argmax of the acquisition taken in the same sweep, per-rank winners all-gathered.
  e2e   : the same call through the public API with HOST buffers: pinned-host candidates in, three pinned-host output
          vectors back, H2D/D2H inside the timed region.
One JSON line on stdout (rank 0).  `roofline` is for the dominant kernel, the variance contraction K* . L^-T + row sums:
  --engine auto|int8 (default): `vargemm_i8_kernel` -- tcgen05 INT8 tensor cores on error-free byte slices of both operands;
          achieved = the INT8 operations the slicing scheme needs (S(S+1)/2 slice GEMMs over the triangular N(N+1)/2 MACs) per
          second, peak = the tcgen05 kind::i8 issue-rate ceiling measured in this run (bosip_b200_int8_peak);
  --engine fp64: `vargemm_kernel` (mma.sync FP64 DMMA); peak = the DMMA microbenchmark measured in this run.
MEASURED_PEAKS.json holds neither an INT8 nor an FP64 figure.  `roofline.fp64_equivalent_tflops` is SURVEY 8(d)'s F_var per
second for either engine.  `cpu_baseline` is the oracle's restatement of the reference CPU path (NumPy/OpenBLAS, all host
cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_CANDIDATES = 10 ** 8
PER_GPU = TOTAL_CANDIDATES // 8          # 1.25e7 candidates per GPU per step
SEED = 20261019 + 3
CPU_SAMPLE = 65536                        # bounded CPU-baseline sample (one 65 536-point chunk, ~17 s)
METRIC = "query pts/sec, approx_posterior+MaxVar at N=1000 d=5"
UNIT = "points/s"


def flops_per_point(p, N, d, ck=9):
    """F_var = p N (N+1), F_mean = 2 p N, F_kern = p N (3d + c_k) as written (Matern-5/2: c_k = 9)."""
    return {"var": p * N * (N + 1), "mean": 2 * p * N, "kern": p * N * (3 * d + ck)}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 3 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": reasons}


def make_problem():
    import bosip_b200 as bb
    pr = bb.synth.make_problem("C3")
    model = bb.GaussianProcess(pr.kernel_id, pr.lam, pr.alpha, pr.sigma)
    like = bb.NormalLikelihood(pr.z_obs, pr.std_obs)
    prior = bb.ProductPrior(pr.prior_kind, pr.prior_params)
    return bb, pr, model, like, prior


def cpu_reference_pass(pr, Xq):
    """The reference's CPU path restated (oracle port): approx_posterior + MaxVar, GP predicted twice per variance call."""
    from oracle import bosip_oracle as orc
    if not hasattr(cpu_reference_pass, "fit"):
        cpu_reference_pass.fit = orc.gp_fit(pr.kernel_id, pr.X, pr.Y, pr.lam, pr.alpha, pr.sigma)
        cpu_reference_pass.like = orc.NormalLikelihood(pr.z_obs, pr.std_obs)
        cpu_reference_pass.prior = orc.ProductPrior(list(pr.prior_kind), [list(q) for q in pr.prior_params])
    return orc.cpu_baseline_posterior_maxvar(cpu_reference_pass.fit, cpu_reference_pass.like, cpu_reference_pass.prior, Xq, chunk=65536)


def config_dict(n_gpus):
    return {"workload": "C3: d=5, p=4, N=1000 ARD Matern-5/2 GP, NormalLikelihood, truncated-Normal prior; 1.25e7 Philox candidates "
                        "per GPU per step (1e8 at 8 GPUs), outputs log_approx_post/log_post_mean/log_post_var + MaxVar argmax",
            "points_per_gpu_per_step": PER_GPU, "global_points_per_step": PER_GPU * n_gpus, "d": 5, "p": 4, "N": 1000,
            "kernel": "Matern52", "parallelism": f"candidate-range shard x{n_gpus}, argmax all-gather",
            "l2": "inputs (500 MB candidates + 2.1-2.4 GB K* workspace per chunk sweep) exceed the 126 MB L2; no explicit flush needed"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; Julia/BOSS.jl cannot run in this
    image) on the host cores, rank 0 only, each step a bounded sample of the workload."""
    if rank != 0:
        return
    import bosip_b200 as bb
    pr = bb.synth.make_problem("C3")
    Xq = bb.synth.philox_uniform_candidates(SEED, 0, CPU_SAMPLE, pr.lb, pr.ub)
    for _ in range(args.warmup):
        cpu_reference_pass(pr, Xq[:, :8192])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(pr, Xq)
    dt = time.perf_counter() - t0
    v = args.steps * CPU_SAMPLE / dt
    cores = os.cpu_count()
    sample = f"{CPU_SAMPLE} of the {PER_GPU} candidates per step, chunked at 65536; NumPy/OpenBLAS, GP predicted twice per variance call as the reference does"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", default="auto", choices=["auto", "int8", "fp64"], help="variance engine (auto = int8 at N=1000)")
    ap.add_argument("--slices", type=int, default=0, help="byte slices of the INT8 engine (0 = default 7)")
    ap.add_argument("--quick-peaks", action="store_true", help="skip the 2 s sustained INT8 peak loop (ncu launch lists)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libbosip_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    # keep NCCL's own chatter (e.g. "NCCL version ...") off stdout: the driver parses ONE JSON line there
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    bb, pr, model, like, prior = make_problem()
    ctx = bb.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_variance_engine(args.engine, args.slices)
    use_i8 = args.engine != "fp64"
    S = args.slices or 7
    peak = ctx.fp64_peak()
    if args.quick_peaks:
        peak["int8_tops_burst"] = peak["int8_tops_sustained"] = ctx.int8_peak()
    else:
        peak["int8_tops_burst"], peak["int8_tops_sustained"] = ctx.int8_peak(sustained=True)
    post = bb.model_posterior(model, pr.X, pr.Y, ctx=ctx)           # every rank conditions the same seeded inputs
    start, count = bb.parallel.shard_range(PER_GPU * world, rank, world)
    src = bb.CandidateSource.philox(SEED, pr.lb, pr.ub, PER_GPU * world)

    # ---- device-resident inputs/outputs
    Xd = torch.empty((count, pr.d), dtype=torch.float64, device="cuda").t()
    bb.candidates(ctx, src, start, count, device_out=Xd)
    outs_d = {k: torch.empty(count, dtype=torch.float64, device="cuda") for k in ("log_approx_post", "log_post_mean", "log_post_var")}
    acq = bb.MaxVar()

    def step_device():
        r = bb.posterior_eval(post, like, prior, Xd, 7, acq=acq, index_offset=start, out=outs_d)
        return bb.parallel.argmax_allgather(r["best_idx"], r["best_val"], np.zeros(1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            last = None
            for _ in range(steps):
                last = fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timing_reset(True)
    l0 = ctx.launch_count()
    ms_total, winner = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    tm = ctx.timing_get(); ctx.timing_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps * PER_GPU * world / (ms_total * 1e-3)

    # ---- end to end through the public API with pinned HOST buffers
    Xh_t = torch.empty((count, pr.d), dtype=torch.float64).pin_memory()
    Xh_t.copy_(Xd.t())
    Xh = Xh_t.numpy().T                                               # d x count, column-major view of the pinned buffer
    outs_h_t = {k: torch.empty(count, dtype=torch.float64).pin_memory() for k in outs_d}
    outs_h = {k: v.numpy() for k, v in outs_h_t.items()}

    def step_host():
        r = bb.posterior_eval(post, like, prior, Xh, 7, acq=acq, index_offset=start, out=outs_h)
        return bb.parallel.argmax_allgather(r["best_idx"], r["best_val"], np.zeros(1))

    step_host()
    ms_e2e, winner_h = timed(step_host, max(1, min(args.steps, 3)))
    e2e_steps = max(1, min(args.steps, 3))
    e2e_value = e2e_steps * PER_GPU * world / (ms_e2e * 1e-3)
    assert winner_h[0] == winner[0], "host and device paths disagree on the argmax"
    assert np.array_equal(outs_h["log_post_var"][:1000], outs_d["log_post_var"][:1000].cpu().numpy())

    if rank == 0:
        fl = flops_per_point(pr.p, pr.N, pr.d)
        vg = tm["var_gemm"]
        pts_per_launch = args.steps * count / max(1, vg["launches"])
        vg_ms = vg["ms"] / max(1, vg["launches"])
        fp64_equiv = fl["var"] * pts_per_launch / (vg_ms * 1e-3) * 1e-12
        tname = "vargemm_i8_traffic.json" if use_i8 else "vargemm_traffic.json"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        if use_i8:
            int8_ops_per_point = S * (S + 1) // 2 * fl["var"]        # every kept slice pair repeats the triangular contraction
            achieved = int8_ops_per_point * pts_per_launch / (vg_ms * 1e-3) * 1e-12
            roof = {"kernel": f"vargemm_i8_kernel<{S}> (tcgen05 INT8 K* . L^-T on {S} byte slices per operand, TMEM accumulators, FP64 Horner "
                              "epilogue + row sum of squares)", "bound": "tensor", "achieved": achieved, "peak": peak["int8_tops_sustained"],
                    "unit": "TOP/s", "frac": achieved / peak["int8_tops_sustained"], "frac_of_burst_peak": achieved / peak["int8_tops_burst"],
                    "traffic": traffic,
                    "peak_source": "tcgen05.mma kind::i8 M=128 N=256 issue-rate microbenchmark measured in this run (bosip_b200_int8_peak): the "
                                   "sustained figure (2 s back to back under the 1 kW cap) because the kernel is timed inside a long step; "
                                   "MEASURED_PEAKS.json has no INT8 figure (2 x its bf16 sustained/burst would be 2807/3335); nominal dense INT8 4500",
                    "algorithmic_int8_ops_per_point": int8_ops_per_point, "slice_pairs": S * (S + 1) // 2}
        else:
            achieved = fp64_equiv
            roof = {"kernel": "vargemm_kernel (FP64 DMMA K* . L^-T + row sum of squares)", "bound": "tensor", "achieved": achieved,
                    "peak": peak["dmma_tflops"], "unit": "TFLOP/s", "frac": achieved / peak["dmma_tflops"], "traffic": traffic,
                    "peak_source": "DMMA m8n8k4 microbenchmark measured in this run (bosip_b200_fp64_peak); MEASURED_PEAKS.json has no FP64 figure; datasheet 37 TFLOP/s"}
        roof.update({"algorithmic_flops_per_point": fl["var"], "fp64_equivalent_tflops": fp64_equiv,
                     "fp64_equivalent_vs_dmma_peak": fp64_equiv / peak["dmma_tflops"], "points_per_launch": pts_per_launch,
                     "ms_per_launch": vg_ms, "launches": vg["launches"], "share_of_step": vg["ms"] / ms_total})
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config_dict(world), variance_engine=("int8x%d" % S) if use_i8 else "fp64"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(count * pr.d * 8), "d2h_bytes_per_step": int(count * 3 * 8),
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in tm.items()},
            "measured_peaks": peak,
            "whole_path_tflops_as_written": sum(fl.values()) * value / world * 1e-12,
            "argmax": {"global_index": int(winner[0]), "value": float(winner[1])},
        }
        if world > 1:
            out["cpu_baseline"] = None      # timed at N=1 only (the other ranks would sit in the barrier for 20-30 s of CPU work)
        elif not args.no_cpu_baseline:
            Xs = bb.synth.philox_uniform_candidates(SEED, 0, CPU_SAMPLE, pr.lb, pr.ub)
            cpu_reference_pass(pr, Xs[:, :8192])
            t0 = time.perf_counter()
            res = cpu_reference_pass(pr, Xs)
            dt = time.perf_counter() - t0
            # the bounded sample doubles as a live parity check of the first chunk
            chk = outs_d["log_post_mean"][:CPU_SAMPLE].cpu().numpy() if start == 0 else None
            if chk is not None:
                err = np.max(np.abs(chk - res[1]) / np.abs(res[1]))
                out["cpu_sample_max_relerr_log_post_mean"] = float(err)
            out["cpu_baseline"] = {"value": CPU_SAMPLE / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{CPU_SAMPLE} of the {PER_GPU} candidates of one step ({dt:.1f} s), chunked at 65536; oracle port of the "
                                             "reference CPU path (NumPy/OpenBLAS all cores, elementwise NumPy single-threaded), GP predicted twice "
                                             "per variance call as src/posterior/likelihood.jl:48-49 does"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
