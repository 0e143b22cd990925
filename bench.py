#!/usr/bin/env python
"""Headline benchmark: MNIST Neural-ODE train step, dopri5 + Taylor R3 (mnist.py --reg r3 --lam 6e-5,
batch 100 synthetic 28x28 inputs, rtol = atol = 1.4e-8 script default).

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torch.distributed.run)
    python bench.py --impl reference ...                   (CPU arm: the oracle port on all host cores)

One "step" = one full update(): PreODE, forward solve, loss head, adjoint solve, momentum update.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every field means.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_steps_per_s"
UNIT = "steps/s"
B, D, H, LAM, TOL = 100, 784, 100, 6e-5, 1.4e-8
WORKLOAD = "mnist.py Neural ODE classifier, dopri5, --reg r3 --lam 6e-5, batch 100 synthetic 28x28 (configs[0])"


def synth_batch(rank, step, pin=False):
    """Synthetic MNIST-shaped batch, a different one per (rank, step % 8)."""
    import torch
    g = torch.Generator().manual_seed(1234 + 97 * rank + (step % 8))
    images = torch.randint(0, 256, (B, 28, 28, 1), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 10, (B,), generator=g)
    eps = torch.randint(0, 2, (B, D), generator=g).float() * 2 - 1
    if pin:
        images, labels, eps = images.pin_memory(), labels.pin_memory(), eps.pin_memory()
    return images, labels, eps


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in list(self.lines):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bench_config(world, workload=None, strong=False):
    """The `config` object of the JSON line; IDENTICAL in both arms (the reference arm runs the b200 arm's config)."""
    if strong and world > 1:
        per, extra = divmod(B, world)
        return {"workload": workload or WORKLOAD, "global_batch": B, "per_gpu_batch": [per + (1 if r < extra else 0) for r in range(world)],
                "rtol": TOL, "atol": TOL,
                "parallelism": (f"dp{world}, STRONG scaling: the batch-{B} problem split in contiguous row blocks (uneven where {B} is not a "
                                "multiple of the GPU count), ONE global adaptive step sequence = the single-GPU sequence"),
                "l2": "flushed between steps (256 MiB device write, outside the per-step events)",
                "trajectory": "train steps 0..K-1 from the initial parameters (seeded), identical in every leg and arm"}
    return {"workload": workload or WORKLOAD, "global_batch": B * world, "per_gpu_batch": B, "rtol": TOL, "atol": TOL,
            "parallelism": (f"dp{world}: batch sharded {B}/GPU, ONE global adaptive step sequence (per-iteration exchange of the "
                            "error-norm partials and of the params_bar stage sums); the reference arm solves the "
                            "concatenated global batch on the host") if world > 1 else "single GPU",
            "l2": "flushed between steps (256 MiB device write, outside the per-step events)",
            "trajectory": "train steps 0..K-1 from the initial parameters (seeded), identical in every leg and arm"}


def traffic_from_profiles(kernel_key):
    """dram bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` summary
    (profiles/roofline_traffic.json, written by tools/ncu_summary.py); None when no capture is committed."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return rec.get(kernel_key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_train_steps(n_steps, warmup, threads, shards=1, budget_s=None):
    """The oracle port of one update() on host cores; returns (seconds per step, info).
    ``shards`` > 1: the global batch of a weak-scaling run (the per-GPU batches of all ranks, concatenated);
    ``budget_s``: stop after that many seconds of timed steps (bounded sample)."""
    import torch
    from oracle import dynamics_oracle as do
    torch.set_num_threads(threads)
    p_ode = do.init_mlp_params(D, H, seed=0)
    p_post = do.init_post_params(D, seed=1)
    vel = None
    times, info = [], {}
    for s in range(-warmup, n_steps):
        if budget_s is not None and times and sum(times) > budget_s:
            break
        parts = [synth_batch(r, max(s, 0)) for r in range(shards)]
        images, labels, eps = (torch.cat([p[k] for p in parts]) for k in range(3))
        t0 = time.perf_counter()
        out = do.train_step_grads(p_ode, p_post, do.pre_ode(images), labels, eps, LAM, reg="r3", rtol=TOL, atol=TOL)
        if s >= 0:   # warm-up evaluations (s < 0) do not move the parameters
            params = {"ode": p_ode, "post": p_post}
            if vel is None:
                vel = do.ode_oracle_tree_zeros(params)
            params, vel = do.momentum_update(params, vel, out["grads"], 1e-1)
            p_ode, p_post = params["ode"], params["post"]
            times.append(time.perf_counter() - t0)
        info = dict(nfe=float(out["nfe"]), fwd_iters=out["fwd_stats"]["n_iters"],
                    bwd_iters=out["bwd_stats"][0]["n_iters"], steps_timed=len(times))
    return sum(times) / len(times), info


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the box's host cores.  The reference itself (JAX,
    Haiku, mid-2020 API) cannot be installed in this image, so this is the oracle PORT (kind 'port')."""
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    steps = max(1, min(args.steps, 40))
    # same workload as the b200 arm at this N: weak scaling, the global batch is B per GPU
    strong = getattr(args, "scaling", "weak") == "strong"
    shards = 1 if strong else max(args.gpus, 1)     # strong scaling: the global batch stays the batch-100 problem
    sec, info = cpu_train_steps(steps, 1, threads, shards=shards, budget_s=150.0)
    steps = info["steps_timed"]
    # same unit as the b200 arm: per-GPU-batch (batch-100) train steps per second over the whole job - one step of
    # the concatenated global batch is `shards` of them
    val = shards / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(max(args.gpus, 1), strong=strong),
        "global_steps_per_s": 1.0 / sec, "samples_per_s": B * shards / sec,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"train steps 0..{steps - 1} from the initial parameters (forward solve + adjoint + "
                                   f"momentum update each) after 1 untimed warm-up evaluation, torch CPU fp32, {threads} threads; "
                                   f"last step nfe {info['nfe']}, adjoint iterations {info['bwd_iters']}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # the reference's own NFE counter (2 + 6 * iterations of the last timed step's forward solve) per second of
        # global solves; "sample_nfe_per_s" multiplies by the samples each evaluation covers
        "nfe_per_s": info["nfe"] / sec, "sample_nfe_per_s": info["nfe"] * B * shards / sec,
    }
    print(json.dumps(line), flush=True)


# ---- BASELINE configs[1]: ffjord_tabular.py, MINIBOONE-shaped 43-d data, --reg r2 --lam 1e-2, batch 1000 ---------------
FJ = dict(B=1000, D=43, HF=20, LAYERS=2, LAM=1e-2, LAM_W=1e-6, TOL=1.4e-8)
FJ_WORKLOAD = ("ffjord_tabular.py FFJORD, synthetic MINIBOONE-shaped 43-d data, ConcatSquash 43-860-860-43 softplus, dopri5, "
               "--reg r2 --lam 1e-2 --lam_w 1e-6, Hutchinson trace (N(0,1) probe), batch 1000 (configs[1])")


def fj_batch(rank, step, pin=False):
    import torch
    g = torch.Generator().manual_seed(4321 + 97 * rank + (step % 8))
    x = torch.randn((FJ["B"], FJ["D"]), generator=g)
    eps = torch.randn((FJ["B"], FJ["D"]), generator=g)
    if pin:
        x, eps = x.pin_memory(), eps.pin_memory()
    return x, eps


def fj_config(world, strong=False):
    gb = FJ["B"] if strong else FJ["B"] * world
    return {"workload": FJ_WORKLOAD, "global_batch": gb, "per_gpu_batch": gb // world, "rtol": FJ["TOL"], "atol": FJ["TOL"],
            "parallelism": "single GPU" if world == 1 else
                           (f"dp{world}: {'the batch-1000 problem split' if strong else 'batch sharded 1000/GPU'} in contiguous row "
                            "blocks, ONE global adaptive step sequence (per iteration: all-reduce of the stage combinations of the "
                            "params_bar block, all-gather of the per-rank error-norm sums; NCCL, captured in the iteration graph)"),
            "l2": "flushed between steps (256 MiB device write, outside the per-step events)",
            "trajectory": "train steps 0..K-1 from the initial parameters (seeded), identical in every leg and arm"}


def fj_cpu_steps(n_steps, warmup, threads, budget_s=None):
    """Oracle port of one update() of ffjord_tabular.py (547-553) on host cores -> (seconds per step, info)."""
    import torch
    from oracle import ffjord_oracle as fo
    torch.set_num_threads(threads)
    params = fo.init_css_params(FJ["D"], FJ["D"] * FJ["HF"], n_hidden_layers=FJ["LAYERS"], seed=0, scale=1.0)
    m, v = fo.adam_init(params)
    times, info = [], {}
    for s in range(-warmup, n_steps):
        if budget_s is not None and times and sum(times) > budget_s:
            break
        x, eps = fj_batch(0, max(s, 0))
        t0 = time.perf_counter()
        out = fo.train_step_grads(params, x, eps, lam=FJ["LAM"], lam_w=FJ["LAM_W"], reg="r2", rtol=FJ["TOL"], atol=FJ["TOL"])
        if s >= 0:
            params, m, v = fo.adam_update(params, m, v, out["grads"], s, 1e-3)
            times.append(time.perf_counter() - t0)
        info = dict(nfe=float(out["nfe"]), fwd_iters=out["fwd_stats"]["n_iters"], bwd_iters=out["bwd_stats"][0]["n_iters"],
                    steps_timed=len(times), loss=float(out["loss"]))
    return sum(times) / len(times), info


def fj_reference(args):
    threads = len(os.sched_getaffinity(0))
    steps = max(1, min(args.steps, 20))
    sec, info = fj_cpu_steps(steps, 1, threads, budget_s=120.0)
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": fj_config(1), "samples_per_s": FJ["B"] / sec,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"train steps 0..{info['steps_timed'] - 1} from the initial parameters, oracle port, torch CPU "
                                       f"fp32, {threads} threads; last step nfe {info['nfe']}, adjoint iterations {info['bwd_iters']}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "nfe_per_s": info["nfe"] / sec}
    print(json.dumps(line), flush=True)


def fj_main(args, result_out, rank, local_rank, world):
    """`--workload ffjord_tabular`: one update() of ffjord_tabular.py per step (forward solve of y, loss head, adjoint solve of
    aug_dynamics with the Hutchinson term and the K = 2 Taylor regulariser, Adam).  Contractions: tcgen05 3xTF32."""
    import torch
    import easy_neural_ode_b200 as eno
    from easy_neural_ode_b200 import ffjord_step, _cabi
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    def make(w):
        return ffjord_step.TabularFfjord(dim=FJ["D"], hdim_factor=FJ["HF"], num_layers=FJ["LAYERS"], reg="r2", lam=FJ["LAM"],
                                         lam_w=FJ["LAM_W"], rtol=FJ["TOL"], atol=FJ["TOL"], device=dev, seed=0, world=w)
    strong = args.scaling == "strong" and world > 1
    parity, sl = None, slice(0, FJ["B"])
    if world > 1:
        # batch rows sharded over the ranks, collective solves (tests/dist_check_flat.py); driver-visible parity: train step 0
        # against the single-device solve of the concatenated batch
        import torch.distributed as dist
        import easy_neural_ode_b200 as eno
        dist.init_process_group("nccl", device_id=dev)
        if FJ["B"] % world and strong:
            raise SystemExit("bench.py: --workload ffjord_tabular --scaling strong needs a world size that divides 1000")
        if strong:
            sl = eno.shard_rows(FJ["B"], world, rank)
        parts = [fj_batch(0 if strong else r, 0) for r in range(1 if strong else world)]
        full = tuple(torch.cat([q[i] for q in parts]).to(dev) for i in range(2))
        want = make(1).loss_and_grads(*full)
        eno.init_distributed()
        mine = tuple(t[sl] if strong else t[rank * FJ["B"]:(rank + 1) * FJ["B"]] for t in full)
        got = make(world).loss_and_grads(*mine)
        parity = {"against": f"single-GPU solve of the concatenated batch ({full[0].shape[0]} rows) of train step 0, same process, "
                             "before going collective",
                  "fwd_iters": [got["fwd_stats"]["n_iters"], want["fwd_stats"]["n_iters"]],
                  "bwd_iters": [got["bwd_stats"][0]["n_iters"], want["bwd_stats"][0]["n_iters"]],
                  "loss_rel": abs(float(got["loss"]) - float(want["loss"])) / abs(float(want["loss"])),
                  "grad_rel": float((got["grads"] - want["grads"]).abs().max() / want["grads"].abs().max())}
        del want, got, full
    model = make(world)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    src = 0 if strong else rank
    dev_b = [tuple(t[sl].to(dev) for t in fj_batch(src, s)) for s in range(8)]
    host_b = [tuple(t[sl].contiguous().pin_memory() for t in fj_batch(src, s)) for s in range(8)]
    snap = model.snapshot()
    for s in range(max(args.warmup, 3)):
        out = model.update(*dev_b[s % 8])
    torch.cuda.synchronize()
    model.restore(snap)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches, nfe_total, iters = 0, 0.0, []
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        evs[s][0].record()
        out = model.update(*dev_b[s % 8])
        evs[s][1].record()
        nfe_total += float(out["nfe"])
        iters.append((out["fwd_stats"]["n_iters"], out["bwd_stats"][0]["n_iters"]))
        launches += model.kernel_launches(out)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:                                             # device time of the slowest rank
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t)
    ms_per_step = dev_ms / args.steps
    units = 1 if (strong or world == 1) else world            # batch-1000 steps one step of the job advances
    loss_dev = float(out["loss"])

    model.restore(snap)
    def step_e2e(s):
        x, eps = host_b[s % 8]
        o = model.update(x.to(dev, non_blocking=True), eps.to(dev, non_blocking=True))
        return float(o["loss"])
    for s in range(2):
        step_e2e(s)
    model.restore(snap)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)

    # roofline of the dominant kernel: the tcgen05 GEMM launches of the solves, CUDA events on the launch stream
    L = _cabi.lib()
    h = _cabi.ctx(dev.index)
    model.restore(snap)
    L.eno_profile_enable(h, 1)
    n_prof = min(args.steps, 3)
    for s in range(n_prof):
        model.update(*dev_b[s % 8])
    torch.cuda.synchronize()
    L.eno_profile_enable(h, 0)
    gms, gfl, gcnt = C.c_double(), C.c_double(), C.c_int64()
    L.eno_profile_read_gemm(h, C.byref(gms), C.byref(gfl), C.byref(gcnt))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops", 1590.0))
    achieved = gfl.value / (gms.value * 1e-3) / 1e12 if gms.value > 0 else 0.0
    roofline = {"kernel": "k_gemm_tf32x3<BN, Epi> (tcgen05.mma kind::tf32, 3-pass split, TMA operands, TMEM accumulators; all "
                          "stage contractions of the forward and adjoint solves)",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "traffic": traffic_from_profiles("k_gemm_tf32x3"),
                "peak_source": ("measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1.59 PFLOP/s") +
                               "; fp32-exact work runs 3 tf32 passes at half the bf16 rate, so 1/6 of this peak is the ceiling "
                               "of the algorithmic rate",
                "avg_launch_ms": gms.value / max(gcnt.value, 1), "launches_timed": int(gcnt.value),
                "algorithmic_flop_per_launch_avg": gfl.value / max(gcnt.value, 1),
                "frac_of_3xtf32_ceiling": achieved / (tensor_peak / 6.0),
                "gemm_share_of_step": (gms.value / n_prof) / ms_per_step}
    if world > 1:
        eno.shutdown_distributed()
        dist.destroy_process_group()
        if rank != 0:
            return
    line = {"metric": METRIC, "value": units * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32 (tcgen05 3xTF32 contractions, fp32 accumulate)", "data": "synthetic",
            "config": fj_config(world, strong), "samples_per_s": units * FJ["B"] * 1e3 / ms_per_step, "clocks": clocks,
            "e2e": {"value": units * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": world * sum(x.numel() * x.element_size() for x in host_b[0]), "d2h_bytes_per_step": 4 * world},
            "gpu_launches": launches, "nfe_per_s": nfe_total / (dev_ms * 1e-3),
            "solver": {"iters_fwd_bwd_per_step": iters, "loss_last": loss_dev},
            "launch_mode": "polled solves; one loop iteration (6 right-hand sides + controller) replayed as a CUDA graph",
            "roofline": roofline}
    if parity:
        line["parity"] = parity
    if not args.no_cpu_baseline and world == 1:
        threads = len(os.sched_getaffinity(0))
        sec, info = fj_cpu_steps(min(args.steps, 3), 1, threads, budget_s=40.0)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"train steps 0..{info['steps_timed'] - 1} of the same trajectory, oracle port, torch CPU "
                                          f"fp32, {threads} threads; last step nfe {info['nfe']}, adjoint iterations "
                                          f"{info['bwd_iters']}, loss {info['loss']:.6f}"}
    print(json.dumps(line), file=result_out, flush=True)


# ---- BASELINE configs[3]: ffjord_mnist.py, 28x28 images, ConcatConv2D dynamics, --reg r2 --lam 3e-4, batch 200 -----------
FM = dict(B=200, IMG=(28, 28), C=64, LAM=3e-4, TOL=1e-5)
FM_WORKLOAD = ("ffjord_mnist.py FFJORD on synthetic 28x28 images, ConcatConv2D 1-64-64-64-1 (3x3) softplus, logit pre-flow, dopri5, "
               "--reg r2 --lam 3e-4, Hutchinson trace (Rademacher probe), rtol = atol = 1e-5 (script default), batch 200 (configs[3])")


def fm_batch(rank, step, pin=False, batch=None):
    import torch
    b = batch or FM["B"]
    g = torch.Generator().manual_seed(8765 + 97 * rank + (step % 8))
    images = torch.rand((b,) + FM["IMG"] + (1,), generator=g)
    eps = torch.randint(0, 2, (b,) + FM["IMG"] + (1,), generator=g).float() * 2 - 1
    if pin:
        images, eps = images.pin_memory(), eps.pin_memory()
    return images, eps


def fm_config(world, strong=False):
    gb = FM["B"] if strong else FM["B"] * world
    return {"workload": FM_WORKLOAD, "global_batch": gb, "per_gpu_batch": gb // world, "rtol": FM["TOL"], "atol": FM["TOL"],
            "parallelism": "single GPU" if world == 1 else
                           (f"dp{world}: {'the batch-200 problem split' if strong else 'batch sharded 200/GPU'} in contiguous row blocks, "
                            "ONE global adaptive step sequence (per iteration: all-reduce of the stage combinations of the params_bar "
                            "block, all-gather of the per-rank error-norm sums; NCCL, captured in the iteration graph)"),
            "l2": "flushed between steps (256 MiB device write, outside the per-step events)",
            "trajectory": "train steps 0..K-1 from the script's initial parameters (last layer zero-initialised, seeded), identical "
                          "in every leg and arm"}


def fm_cpu_steps(n_steps, threads, batch, budget_s=None):
    """Oracle port of one update() of ffjord_mnist.py (570-583) on host cores, on a shard of `batch` images of the
    same synthetic batches -> (seconds per step, info).  Adam on the host as in ffjord_oracle."""
    import torch
    from oracle import ffjord_mnist_oracle as fm, ffjord_oracle as fo
    torch.set_num_threads(threads)
    params = fm.init_conv_params(hidden=FM["C"], seed=0)
    m, v = fo.adam_init(params)
    times, info = [], {}
    for s in range(n_steps):
        if budget_s is not None and times and sum(times) > budget_s:
            break
        images, eps = fm_batch(0, s, batch=batch)
        t0 = time.perf_counter()
        out = fm.train_step_grads(params, images, eps, lam=FM["LAM"], reg="r2", rtol=FM["TOL"], atol=FM["TOL"])
        params, m, v = fo.adam_update(params, m, v, out["grads"], s, 1e-3 * min((s + 1.0) / 1e3, 1.0))
        times.append(time.perf_counter() - t0)
        info = dict(nfe=float(out["nfe"]), bwd_iters=out["bwd_stats"][0]["n_iters"], steps_timed=len(times), loss=float(out["loss"]))
    return sum(times) / len(times), info


def fm_reference(args):
    threads = len(os.sched_getaffinity(0))
    shard = 8
    sec, info = fm_cpu_steps(max(1, min(args.steps, 6)), threads, shard, budget_s=120.0)
    val = (shard / FM["B"]) / sec          # batch-200 train steps per second, extrapolated from the shard
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": fm_config(1), "samples_per_s": shard / sec,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"train steps 0..{info['steps_timed'] - 1} on a shard of {shard} of the 200 images (the cost of a "
                                       f"step is linear in the batch), scaled to 200; oracle port, torch CPU fp32, {threads} threads; "
                                       f"last step nfe {info['nfe']}, adjoint iterations {info['bwd_iters']}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "nfe_per_s": info["nfe"] * val}
    print(json.dumps(line), flush=True)


def fm_main(args, result_out, rank, local_rank, world):
    """`--workload ffjord_mnist`: one update() of ffjord_mnist.py per step (logit pre-flow, forward solve of y, bits/dim head,
    adjoint solve of aug_dynamics with the Hutchinson term and the K = 2 Taylor regulariser, clip + Adam)."""
    import torch
    from easy_neural_ode_b200 import ffjord_mnist_step
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    strong = args.scaling == "strong" and world > 1
    parity = None
    if world > 1:
        # batch rows sharded over the ranks, collective solves (tests/dist_check_flat.py); strong: BASELINE's 8-GPU line is
        # the batch-200 step itself split over the GPUs (25 images each at 8)
        import torch.distributed as dist
        import easy_neural_ode_b200 as eno
        dist.init_process_group("nccl", device_id=dev)
        if FM["B"] % world and strong:
            raise SystemExit("bench.py: --workload ffjord_mnist --scaling strong needs a world size that divides 200")
        sl = eno.shard_rows(FM["B"], world, rank) if strong else slice(0, FM["B"])
        do_parity = strong or world == 2                      # weak: the concatenated batch is 200 x world images
        if do_parity:
            parts = [fm_batch(0 if strong else r, 0) for r in range(1 if strong else world)]
            full = tuple(torch.cat([q[i] for q in parts]).to(dev) for i in range(2))
            single = ffjord_mnist_step.ImageFfjord(FM["IMG"], FM["C"], reg="r2", lam=FM["LAM"], rtol=FM["TOL"], atol=FM["TOL"],
                                                   device=dev, seed=0)
            single.update(*full)                              # one update first: the zero-initialised last layer has f == 0
            want = single.loss_and_grads(*full)
            snap1 = single.snapshot()
        eno.init_distributed()
    else:
        sl = slice(0, FM["B"])
    model = ffjord_mnist_step.ImageFfjord(FM["IMG"], FM["C"], reg="r2", lam=FM["LAM"], rtol=FM["TOL"], atol=FM["TOL"], device=dev, seed=0,
                                          world=world)
    if world > 1 and do_parity:
        # driver-visible parity: the collective solve of train step 1 against the single-device solve of the concatenated batch
        model.restore(snap1)
        mine = tuple(t[sl] if strong else t[rank * FM["B"]:(rank + 1) * FM["B"]] for t in full)
        got = model.loss_and_grads(*mine)
        gr = float((got["grads"] - want["grads"]).abs().max() / want["grads"].abs().max())
        parity = {"against": f"single-GPU solve of the concatenated batch ({full[0].shape[0]} images) of train step 1, same process, "
                             "before going collective",
                  "fwd_iters": [got["fwd_stats"]["n_iters"], want["fwd_stats"]["n_iters"]],
                  "bwd_iters": [got["bwd_stats"][0]["n_iters"], want["bwd_stats"][0]["n_iters"]],
                  "loss_rel": abs(float(got["loss"]) - float(want["loss"])) / abs(float(want["loss"])), "grad_rel": gr}
        model = ffjord_mnist_step.ImageFfjord(FM["IMG"], FM["C"], reg="r2", lam=FM["LAM"], rtol=FM["TOL"], atol=FM["TOL"], device=dev,
                                              seed=0, world=world)
        del single, want, full
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    src = 0 if strong else rank
    dev_b = [tuple(t[sl].to(dev) for t in fm_batch(src, s)) for s in range(8)]
    host_b = [tuple(t[sl].contiguous().pin_memory() for t in fm_batch(src, s)) for s in range(8)]
    snap = model.snapshot()
    for s in range(max(args.warmup, 3)):
        out = model.update(*dev_b[s % 8])
    torch.cuda.synchronize()
    model.restore(snap)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    nfe_total, iters, n_rhs = 0.0, [], 0
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        evs[s][0].record()
        out = model.update(*dev_b[s % 8])
        evs[s][1].record()
        nfe_total += float(out["nfe"])
        iters.append((out["fwd_stats"]["n_iters"], out["bwd_stats"][0]["n_iters"]))
        n_rhs += 2 + 6 * out["fwd_stats"]["n_iters"] + 3 + 6 * out["bwd_stats"][0]["n_iters"]
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:                                             # device time of the slowest rank
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t)
    ms_per_step = dev_ms / args.steps
    units = 1 if (strong or world == 1) else world            # batch-200 steps one step of the job advances
    loss_dev = float(out["loss"])
    model.restore(snap)

    def step_e2e(s):
        images, eps = host_b[s % 8]
        o = model.update(images.to(dev, non_blocking=True), eps.to(dev, non_blocking=True))
        return float(o["loss"])
    for s in range(2):
        step_e2e(s)
    model.restore(snap)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops", 1590.0))
    # roofline of the dominant kernels: the tcgen05 convolution launches of the solves (forward, backward data, weight
    # gradients), CUDA events on the launch stream, algorithmic flops (interior positions, not the three tf32 passes)
    from easy_neural_ode_b200 import _cabi
    L = _cabi.lib()
    h = _cabi.ctx(dev.index)
    model.restore(snap)
    L.eno_profile_enable(h, 1)
    n_prof = min(args.steps, 2)
    for s in range(n_prof):
        model.update(*dev_b[s % 8])
    torch.cuda.synchronize()
    L.eno_profile_enable(h, 0)
    gms, gfl, gcnt = C.c_double(), C.c_double(), C.c_int64()
    L.eno_profile_read_gemm(h, C.byref(gms), C.byref(gfl), C.byref(gcnt))
    model.restore(snap)
    achieved = gfl.value / (gms.value * 1e-3) / 1e12 if gms.value > 0 else 0.0
    roofline = {"kernel": "k_gemm_tf32x3<64, EpiConv*> (the 64 -> 64 ConcatConv2D layers as 9-tap K loops: forward, backward data "
                          "(mirrored taps) and weight gradients (MN-major operands) on tcgen05.mma kind::tf32, 3-pass split, TMA "
                          "operands, TMEM accumulators)",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "traffic": None,
                "peak_source": ("measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1.59 PFLOP/s") +
                               "; fp32-exact work runs 3 tf32 passes at half the bf16 rate, so 1/6 of this peak is the ceiling "
                               "of the algorithmic rate",
                "avg_launch_ms": gms.value / max(gcnt.value, 1), "launches_timed": int(gcnt.value),
                "algorithmic_flop_per_launch_avg": gfl.value / max(gcnt.value, 1),
                "frac_of_3xtf32_ceiling": achieved / (tensor_peak / 6.0),
                "gemm_share_of_step": (gms.value / n_prof) / ms_per_step}
    if world > 1:
        eno.shutdown_distributed()
        dist.destroy_process_group()
        if rank != 0:
            return
    line = {"metric": METRIC, "value": units * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32 (tcgen05 3xTF32 convolutions, fp32 accumulate)", "data": "synthetic",
            "config": fm_config(world, strong), "samples_per_s": units * FM["B"] * 1e3 / ms_per_step, "clocks": clocks,
            "e2e": {"value": units * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": world * sum(x.numel() * x.element_size() for x in host_b[0]), "d2h_bytes_per_step": 4 * world},
            "gpu_launches": 20 * n_rhs, "nfe_per_s": nfe_total / (dev_ms * 1e-3),
            "solver": {"iters_fwd_bwd_per_step": iters, "loss_last": loss_dev},
            "launch_mode": "polled solves; one loop iteration (6 right-hand sides + controller) replayed as a CUDA graph",
            "roofline": roofline}
    if parity:
        line["parity"] = parity
    if not args.no_cpu_baseline and world == 1:
        threads = len(os.sched_getaffinity(0))
        shard = 8
        sec, info = fm_cpu_steps(min(args.steps, 3), threads, shard, budget_s=40.0)
        line["cpu_baseline"] = {"value": (shard / FM["B"]) / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"train steps 0..{info['steps_timed'] - 1} on a shard of {shard} of the 200 images, scaled to 200; "
                                          f"oracle port, torch CPU fp32, {threads} threads; last step nfe {info['nfe']}, adjoint iterations "
                                          f"{info['bwd_iters']}, loss {info['loss']:.6f}"}
    print(json.dumps(line), file=result_out, flush=True)


# ---- BASELINE configs[2]: latent_ode.py generative ODE, per-trajectory solves (development workload; --gpus N shards the series) --
LT = dict(S=3, B=50, D=20, U=50, LAYERS=3, T=49, DATA=37, LAM=1e-2, TOL=1.4e-8)
LT_WORKLOAD = ("latent_ode.py generative ODE, synthetic Physionet-shaped series (37 features, 49 time points on [0, 1]), "
               "GenDynamics 20-50x4-20 tanh, dopri5, --reg r3 --lam 1e-2, 3 samples x batch 50 = 150 per-trajectory solves "
               "with 49 output times each, float64 (configs[2])")


def lt_batch(step, pin=False, rank=0):
    import torch
    g = torch.Generator().manual_seed(977 + 97 * rank + (step % 8))
    z0 = 0.1 * torch.randn((LT["S"], LT["B"], LT["D"]), generator=g, dtype=torch.float64)
    data = torch.randn((LT["B"], LT["T"], LT["DATA"]), generator=g, dtype=torch.float64)
    mask = (torch.rand((LT["B"], LT["T"], LT["DATA"]), generator=g) < 0.1).double()
    if pin:
        z0, data, mask = z0.pin_memory(), data.pin_memory(), mask.pin_memory()
    return z0, data, mask


def lt_config(world=1, strong=False):
    gb = LT["S"] * LT["B"] * (1 if strong else world)
    return {"workload": LT_WORKLOAD, "global_batch": gb, "per_gpu_batch": gb / world, "rtol": LT["TOL"],
            "atol": LT["TOL"],
            "parallelism": "single GPU (trajectories are independent: replicas only, no exchange)" if world == 1 else
                           (f"dp{world}: the batch axis {'of the 50-series problem split' if strong else 'sharded 50 series/GPU'}; "
                            "trajectories are independent (no collective in the solves), the likelihood sums and the parameter "
                            "gradient are all-reduced once per step"),
            "l2": "flushed between steps (256 MiB device write, outside the per-step events)",
            "trajectory": "train steps 0..K-1 from the initial parameters (seeded), identical in every leg and arm"}


def lt_cpu_sample(n_traj, threads):
    """Oracle port on host cores: forward solve + complete adjoint of `n_traj` of the 150 trajectories of train step 0
    (float64, nested-jvp Taylor closure), -> seconds per trajectory and the iteration counts."""
    import math
    import torch
    from oracle import latent_oracle as lo, ode_oracle as oo
    torch.set_num_threads(threads)
    p = lo.init_gen_params(LT["D"], LT["U"], LT["LAYERS"], seed=0, dtype=torch.float64)
    cl = lo.make_latent_closures("r3")
    z0, _, _ = lt_batch(0)
    z0 = z0.reshape(-1, LT["D"])
    ts = torch.linspace(0., 1., LT["T"], dtype=torch.float64)
    g = torch.Generator().manual_seed(5)
    t0 = time.perf_counter()
    nfe, bwd_iters = [], []
    for k in range(n_traj):
        y0 = (z0[k:k + 1], torch.zeros((), dtype=torch.float64))
        flat0, unravel = oo.ravel(y0)
        func = lambda yf, t, params: oo.ravel(cl["aug_dynamics"](unravel(yf), t, params))[0]
        ys, n, _ = oo.odeint(func, flat0, ts, p, rtol=LT["TOL"], atol=LT["TOL"], return_stats=True)
        gflat = torch.randn(ys.shape, generator=g, dtype=torch.float64)
        st = []
        oo.odeint_rev(func, LT["TOL"], LT["TOL"], math.inf, ys, ts, (p,), gflat, st)
        nfe.append(n)
        bwd_iters.append(sum(x["n_iters"] for x in st))
    return (time.perf_counter() - t0) / n_traj, nfe, bwd_iters


def lt_reference(args):
    threads = len(os.sched_getaffinity(0))
    n = 2
    sec, nfe, bwd = lt_cpu_sample(n, threads)
    n_all = LT["S"] * LT["B"]
    val = 1.0 / (sec * n_all)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * n_all * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": lt_config(),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} of the {n_all} trajectories of train step 0 (forward solve + complete adjoint, oracle "
                                       f"port, torch CPU float64, {threads} threads), scaled to {n_all}; NFE {nfe}, adjoint "
                                       f"iterations {bwd}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def lt_main(args, result_out, local_rank, rank=0, world=1):
    """`--workload latent_ode`: the generative half of one update() of latent_ode.py per step - 150 per-trajectory forward
    solves (one launch), decoder + likelihood, 150 complete adjoints of 48 segments each (one launch), Adam."""
    import torch
    from easy_neural_ode_b200 import latent_step, latent
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    strong = args.scaling == "strong" and world > 1
    sl, parity = slice(0, LT["B"]), None
    ts = torch.linspace(0., 1., LT["T"], dtype=torch.float64, device=dev)

    def make(w):
        return latent_step.LatentGen(LT["D"], LT["LAYERS"], LT["U"], LT["DATA"], reg="r3", lam=LT["LAM"], rtol=LT["TOL"],
                                     atol=LT["TOL"], device=dev, seed=0, world=w)

    def cut(b, rows):                                         # z0 [S, B, D], data / mask [B, T, F]: rows of the batch axis
        return b[0][:, rows].contiguous(), b[1][rows].contiguous(), b[2][rows].contiguous()
    if world > 1:
        import torch.distributed as dist
        import easy_neural_ode_b200 as eno
        dist.init_process_group("nccl", device_id=dev)
        if strong:
            sl = eno.shard_rows(LT["B"], world, rank)
        # driver-visible parity: train step 0 sharded vs the same step on one GPU over the concatenated batch
        parts = [lt_batch(0, rank=0 if strong else r) for r in range(1 if strong else world)]
        full = tuple(t.to(dev) for t in (torch.cat([q[0] for q in parts], dim=1), torch.cat([q[1] for q in parts]),
                                         torch.cat([q[2] for q in parts])))
        want = make(1).loss_and_grads(full[0], ts, full[1], full[2])
        rows = sl if strong else slice(rank * LT["B"], (rank + 1) * LT["B"])
        got = make(world).loss_and_grads(*[x for x in (cut(full, rows)[0], ts) + cut(full, rows)[1:]])
        parity = {"against": f"the same train step on one GPU over the concatenated batch ({full[1].shape[0]} series)",
                  "loss_rel": abs(float(got["loss"]) - float(want["loss"])) / abs(float(want["loss"])),
                  "grad_rel": max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(got["grads"], want["grads"])),
                  "nfe_identical": bool(torch.equal(got["nfe"], want["nfe"][:, rows])) if want["nfe"].dim() == 2 else None}
        del want, got, full
    model = make(world)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    src = 0 if strong else rank
    dev_b = [tuple(t.to(dev) for t in cut(lt_batch(s, rank=src), sl)) for s in range(8)]
    host_b = [tuple(t.pin_memory() for t in cut(lt_batch(s, rank=src), sl)) for s in range(8)]
    snap = model.snapshot()
    for s in range(max(args.warmup, 3)):
        out = model.update(dev_b[s % 8][0], ts, *dev_b[s % 8][1:])
    torch.cuda.synchronize()
    model.restore(snap)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    nfe_total, rhs_f, rhs_b, it_b = 0.0, 0, 0, 0
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        evs[s][0].record()
        out = model.update(dev_b[s % 8][0], ts, *dev_b[s % 8][1:])
        evs[s][1].record()
        nfe_total += float(out["nfe"].sum())
        rhs_f += int(2 * out["nfe"].numel() + 6 * out["fwd_iters"][:, 0].sum())
        it_b += int(out["bwd_iters"][..., 0].sum())
        rhs_b += int(2 * out["bwd_iters"][..., 0].numel() + 6 * out["bwd_iters"][..., 0].sum())
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:                                             # device time of the slowest rank
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t)
    ms_per_step = dev_ms / args.steps
    units = 1 if (strong or world == 1) else world            # 150-trajectory steps one step of the job advances
    loss_dev = float(out["loss"])

    model.restore(snap)
    def step_e2e(s):
        z0, data, mask = host_b[s % 8]
        o = model.update(z0.to(dev, non_blocking=True), ts, data.to(dev, non_blocking=True), mask.to(dev, non_blocking=True))
        return float(o["loss"])
    for s in range(2):
        step_e2e(s)
    model.restore(snap)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
        dist.destroy_process_group()
        if rank != 0:
            return

    # the dominant kernel (the adjoint of all trajectories, one launch) timed alone on the training state of step 0
    model.restore(snap)
    z0, data, mask = dev_b[0]
    spec, p = model.spec, model.params
    zs, rs, _, it_f, _ = latent.traj_fwd(spec, True, z0.reshape(-1, LT["D"]).contiguous(), ts, p, LT["TOL"], LT["TOL"], 0)
    g_z = torch.randn_like(zs)
    g_r = torch.zeros_like(rs)
    g_r[:, -1] = LT["LAM"] / rs.shape[0]
    reps, kms, fms = 5, 0.0, 0.0
    for i in range(reps + 1):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        flush.fill_(i)
        a.record()
        latent.traj_fwd(spec, True, z0.reshape(-1, LT["D"]).contiguous(), ts, p, LT["TOL"], LT["TOL"], 0)
        b.record()
        res = latent.traj_bwd(spec, True, zs, rs, ts, p, g_z, g_r, LT["TOL"], LT["TOL"], 0)
        c.record()
        torch.cuda.synchronize()
        if i:
            fms += a.elapsed_time(b)
            kms += b.elapsed_time(c)
    it1 = int(res[5][..., 0].sum())
    segs = res[5][..., 0].numel()
    P, nst = spec.n_params, 2 * 21 + 1 + spec.n_params
    mv = 2 * sum(a * b for a, b in zip(spec.dims[:-1], spec.dims[1:]))      # flops of one pass through the stack
    rhs1 = 2 * segs + 6 * it1
    # algorithmic work of the adjoint launch: per right-hand side 3 Taylor streams forward, their reverse sweep (again 3
    # transposed products) and 3 outer products per layer for the parameter cotangent; per iteration ~70 flops per state
    # element for the RK combination, error estimate and norm (SURVEY.md section 8d's count)
    flop = rhs1 * 9 * mv + it1 * 70 * nst
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops", 1590.0))
    achieved = flop / (kms / reps * 1e-3) / 1e12
    fp64_peak = 37.0
    roofline = {"kernel": "k_traj_bwd<double> (the complete adjoint of all 150 trajectories, 48 segments each, in ONE launch: "
                          "one CTA per trajectory, weights resident in shared memory, float64 FMA)",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "traffic": None, "peak_source": ("measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1.59 PFLOP/s"),
                "avg_launch_ms": kms / reps, "launches_timed": reps, "algorithmic_flop_per_launch": flop,
                "fp64_fma_peak_tflops": fp64_peak, "frac_of_fp64_fma": achieved / fp64_peak,
                "forward_launch_ms": fms / reps, "adjoint_iterations_per_launch": it1, "adjoint_rhs_per_launch": rhs1,
                "note": "float64 work on the non-tensor pipe (the script runs with jax_enable_x64); the tensor pipe is the contract's "
                        "denominator, 37 TFLOP/s is B200's nominal float64 vector peak (not in MEASURED_PEAKS.json).  150 sequential "
                        "per-trajectory solves of 2,500-FMA products are latency bound: one trajectory cannot fill an SM"}
    line = {"metric": METRIC, "value": units * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": lt_config(world, strong),
            "samples_per_s": units * LT["S"] * LT["B"] * 1e3 / ms_per_step, "clocks": clocks,
            "e2e": {"value": units * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": world * sum(x.numel() * x.element_size() for x in host_b[0]), "d2h_bytes_per_step": 8 * world},
            "gpu_launches": 3 * args.steps, "nfe_per_s": nfe_total / (dev_ms * 1e-3),
            "adjoint_rhs_per_s": rhs_b / (dev_ms * 1e-3),
            "solver": {"mean_nfe_per_trajectory": nfe_total / args.steps / (LT["S"] * LT["B"]),
                       "adjoint_iterations_per_step": it_b / args.steps, "forward_rhs_per_step": rhs_f / args.steps, "loss_last": loss_dev},
            "launch_mode": "two solver launches per step (all forward solves; all adjoints) + the per-trajectory sum; decoder, "
                           "likelihood and Adam are float64 torch ops",
            "roofline": roofline}
    if parity:
        line["parity"] = parity
    if not args.no_cpu_baseline and world == 1:
        threads = len(os.sched_getaffinity(0))
        n = 2
        sec, nfe, bwd = lt_cpu_sample(n, threads)
        n_all = LT["S"] * LT["B"]
        line["cpu_baseline"] = {"value": 1.0 / (sec * n_all), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} of the {n_all} trajectories of train step 0 (forward solve + complete adjoint, oracle "
                                          f"port, torch CPU float64, {threads} threads), scaled to {n_all}; NFE {nfe}, adjoint "
                                          f"iterations {bwd}"}
    print(json.dumps(line), file=result_out, flush=True)


def main():
    global B
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="polled solves instead of the whole-step CUDA graph")
    ap.add_argument("--batch", type=int, default=B, help="per-GPU batch (development sweeps only)")
    ap.add_argument("--reg-type", default="our", choices=["our", "fin"],
                    help="development only: 'fin' times the --reg_type fin arm (odeint_sepaux, lam_fro = lam_kin = 1e-2); "
                         "the headline metric is the default 'our' (--reg r3)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): batch 100 PER GPU; strong: the batch-100 problem itself split "
                         "over the GPUs in contiguous row blocks (100 rows on 8 GPUs = 13,13,13,13,12,12,12,12: uneven shards)")
    ap.add_argument("--workload", default="mnist", choices=["mnist", "ffjord_tabular", "latent_ode", "ffjord_mnist"],
                    help="mnist = BASELINE configs[0], the configuration the metric is quoted on (default); "
                         "ffjord_tabular = configs[1] (--gpus N: batch rows sharded, collective solves); latent_ode = configs[2] (--gpus N: series sharded, independent solves, gradients all-reduced); ffjord_mnist = configs[3] "
                         "(--gpus N: batch rows sharded, collective solves; --scaling strong splits the batch-200 step itself)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.workload == "ffjord_tabular":
            return fj_reference(args) if rank == 0 else None
        if args.workload == "latent_ode":
            return lt_reference(args) if rank == 0 else None
        if args.workload == "ffjord_mnist":
            return fm_reference(args) if rank == 0 else None
        return run_reference(args, rank, world)
    # stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr, so
    # that banners printed by native libraries (NCCL prints its version on communicator creation) cannot
    # end up in front of the result
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import easy_neural_ode_b200 as eno
    from easy_neural_ode_b200 import mnist_step, _cabi

    if args.workload == "ffjord_tabular":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
        return fj_main(args, result_out, rank, local_rank, world)
    if args.workload == "latent_ode":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
        return lt_main(args, result_out, local_rank, rank, world)
    if args.workload == "ffjord_mnist":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
        return fm_main(args, result_out, rank, local_rank, world)
    B = args.batch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    single = None
    strong = args.scaling == "strong" and world > 1
    sl = eno.shard_rows(B, world, rank) if strong else slice(0, B)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if args.reg_type == "our":
            # parity evidence for the sharded solve, untimed: every rank first solves the CONCATENATED global batch of
            # train step 0 on its own GPU, non-collectively (the library is not in collective mode yet) ...
            ref_model = mnist_step.MnistNode(reg="r3", lam=LAM, rtol=TOL, atol=TOL, device=dev, dim=D, hidden=H, world=1)
            parts = [synth_batch(r, 0) for r in range(1 if strong else world)]
            o = ref_model.loss_and_grads(*(torch.cat([q[k] for q in parts]).to(dev) for k in range(3)))
            mine = sl if strong else slice(rank * B, (rank + 1) * B)
            single = dict(y1=o["y1"][mine].clone(), g=o["grads"]["ode"].clone(), loss=float(o["loss"]),
                          fwd=o["fwd_stats"]["n_iters"], bwd=o["bwd_stats"][0]["n_iters"])
            del ref_model, o
        eno.init_distributed()   # solves become collective: one global adaptive step sequence
        if strong:
            eno.set_shards(B)    # ranks hold different numbers of rows; the error mean runs over the B rows of the problem

    if args.reg_type == "fin":
        model = mnist_step.MnistNode(reg="none", rtol=TOL, atol=TOL, device=dev, dim=D, hidden=H, world=world,
                                     reg_type="fin", lam_fro=1e-2, lam_kin=1e-2)
    else:
        model = mnist_step.MnistNode(reg="r3", lam=LAM, rtol=TOL, atol=TOL, device=dev, dim=D, hidden=H, world=world)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    if strong:
        model.rows_global = B
        dev_batches = [tuple(t[sl].contiguous().to(dev) for t in synth_batch(0, s)) for s in range(8)]
        host_batches = [tuple(t[sl].contiguous().pin_memory() for t in synth_batch(0, s)) for s in range(8)]
    else:
        dev_batches = [tuple(t.to(dev) for t in synth_batch(rank, s)) for s in range(8)]
        host_batches = [synth_batch(rank, s, pin=True) for s in range(8)]

    parity = None
    if single is not None:
        # ... then the same step as a collective solve over the shards; compared block by block
        o = model.loss_and_grads(*dev_batches[0])
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
        loss_g = o["loss"].detach().double().reshape(1).clone()      # mean over this rank's rows -> mean over the global batch
        if strong:
            loss_g *= world * (sl.stop - sl.start) / B                # uneven shards: weight by the rows held
        dist.all_reduce(loss_g)
        errs = torch.tensor([rel(o["y1"], single["y1"]), rel(o["grads"]["ode"], single["g"]),
                             abs(float(loss_g) / world - single["loss"]) / abs(single["loss"]),
                             float(o["fwd_stats"]["n_iters"] != single["fwd"]), float(o["bwd_stats"][0]["n_iters"] != single["bwd"])],
                            device=dev, dtype=torch.float64)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        parity = {"against": f"single-GPU solve of the concatenated batch ({B if strong else B * world} rows) of train step 0, same process, before "
                             "the library entered collective mode; max over ranks",
                  "fwd_iters": [o["fwd_stats"]["n_iters"], single["fwd"]], "bwd_iters": [o["bwd_stats"][0]["n_iters"], single["bwd"]],
                  "iteration_counts_identical_on_all_ranks": bool(errs[3] == 0 and errs[4] == 0),
                  "y1_rel": float(errs[0]), "grad_rel": float(errs[1]), "loss_rel": float(errs[2])}
        single = None

    use_graph = not args.no_graph

    def step(s, batch=None):
        images, labels, eps = batch if batch is not None else dev_batches[s % 8]
        if use_graph:
            # the whole update() as ONE CUDA-graph launch (deferred-mode solves, device-guarded update)
            return model.update_graphed(images, labels, eps)
        # world > 1: the batch is sharded over ranks; inside the solves the error norm and the
        # batch-summed adjoint blocks are exchanged over NCCL every iteration, so params_bar comes back
        # globally reduced and every rank applies the same update
        return model.update(images, labels, eps)

    # Every measured leg (device-timed, e2e, profiled, CPU) replays the SAME training trajectory:
    # steps 0..K-1 from the initial parameters, so the arms do identical work (the solver takes
    # more iterations as training stiffens the dynamics).
    snap = model.snapshot()
    for s in range(max(args.warmup, 3)):
        out = step(s)
    torch.cuda.synchronize()
    model.restore(snap)

    # ---- timed region: device-resident inputs, L2 flushed between steps, CUDA events, max over ranks
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    nfe_total = 0.0
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        evs[s][0].record()
        out = step(s)
        evs[s][1].record()
        if use_graph:
            launches += model._launches
            nfe_total += float(out["nfe"])
        else:
            launches += out["fwd_stats"]["n_launches"] + sum(b["n_launches"] for b in out["bwd_stats"])
            nfe_total += float(out["fwd_stats"]["nfe"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    units = 1 if strong else world                # batch-100 problems one step of the job advances
    value = units * 1e3 / ms_per_step             # batch-100 train steps per second, all ranks

    # ---- e2e: host (pinned) buffers through the public API, H2D + D2H inside the timed region
    def step_e2e(s):
        images, labels, eps = host_batches[s % 8]
        di, dl, de = images.to(dev, non_blocking=True), labels.to(dev, non_blocking=True), eps.to(dev, non_blocking=True)
        o = step(s, (di, dl, de))
        return float(o["loss"])                    # device -> host read of the step's result
    for s in range(3):
        step_e2e(s)
    model.restore(snap)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dbg = []
    for s in range(args.steps):
        t1 = time.perf_counter()
        step_e2e(s)
        dbg.append(round((time.perf_counter() - t1) * 1e3, 2))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("ENO_BENCH_DEBUG"):
        print("e2e per-step ms:", dbg, file=sys.stderr)
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units * args.steps / float(t.item())
    h2d = sum(x.numel() * x.element_size() for x in host_batches[0])

    # ---- roofline of the dominant kernel (adjoint per-sample step kernel), separate profiled pass
    L = _cabi.lib()
    h = _cabi.ctx(dev.index)
    model.restore(snap)
    use_graph = False                              # per-kernel events need the polled launch path
    L.eno_profile_enable(h, 1)
    for s in range(args.steps):
        step(s)
    torch.cuda.synchronize()
    L.eno_profile_enable(h, 0)
    ms = (C.c_double * 5)()
    cnt = (C.c_int64 * 5)()
    L.eno_profile_read(h, ms, cnt)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1.59 PFLOP/s"
    ffma = C.c_double()
    L.eno_ffma_peak(h, C.byref(ffma), None)
    rows_ms = ms[1] / max(cnt[1], 1)
    pgemm_ms = ms[4] / max(cnt[4], 1)
    flop_rows = 6.0 * B * 12 * 2 * D * H + 70.0 * 2 * B * D   # DESIGN.md "Algorithmic work"
    achieved = flop_rows / (rows_ms * 1e-3) / 1e12 if rows_ms > 0 else 0.0
    roofline = {"kernel": "k_cl_adj_rows<STEP> (adjoint dopri5 iteration, per-sample blocks; k_adj_rows on the streamed path)", "bound": "tensor",
                "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "traffic": traffic_from_profiles("k_cl_adj_rows_step"),   # ncu --set full, per launch (profiles/roofline_traffic.json)
                "peak_source": peak_src, "avg_launch_ms": rows_ms, "launches_timed": int(cnt[1]),
                "fp32_ffma_peak_tflops": ffma.value, "frac_of_fp32_ffma": achieved / ffma.value if ffma.value else None,
                "other_kernels_ms": {"k_fwd_step": ms[0] / max(cnt[0], 1),
                                     "adjoint parameter gradients per iteration (batched tcgen05 3xTF32 GEMM + k_adj_pcombine; "
                                     "k_adj_pgrad on the streamed path)": ms[2] / max(cnt[2], 1),
                                     **({"k_gemm_tf32x3<128, EpiPgradStage> alone": pgemm_ms} if cnt[4] else {}),
                                     **({"exchange (finish kernel incl. the peer-memory or NCCL exchange and the wait for the slowest rank), per iteration, fwd and adjoint":
                                         ms[3] / max(cnt[3], 1)} if world > 1 else {})},
                "note": "fp32 FFMA kernel; the tensor pipe is the contract's denominator, the fp32 pipe is the one it uses"}
    if cnt[4]:
        # the tensor-core kernel of the step: 12 contractions [D x K] x [K x H], K = 3 B, per launch (algorithmic 2 M N K flops)
        flop_pg = 12.0 * 2 * D * H * 3 * B
        roofline["tensor_kernel"] = {
            "kernel": "k_gemm_tf32x3<128, EpiPgradStage> (adjoint parameter-gradient contraction: tcgen05.mma kind::tf32, 3-pass "
                      "split, TMA operands, TMEM accumulators; one batched launch per adjoint iteration)",
            "avg_launch_ms": pgemm_ms, "launches_timed": int(cnt[4]), "algorithmic_flop_per_launch": flop_pg,
            "achieved": flop_pg / (pgemm_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
            "frac_of_3xtf32_ceiling": flop_pg / (pgemm_ms * 1e-3) / 1e12 / (tensor_peak / 6.0),
            "note": "fp32-exact work runs 3 tf32 passes at half the bf16 rate: 1/6 of the bf16 peak is the ceiling; at B = 100 the "
                    "launch is 84 tiles of 128 x 128 x 300 - latency bound, not throughput bound"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": bench_config(world, strong=strong, workload=None if args.reg_type == "our" else
                               "mnist.py Neural ODE classifier, dopri5, --reg_type fin --lam_fro 1e-2 --lam_kin 1e-2, batch 100 "
                               "synthetic 28x28 (development run, not the headline config)"),
        "global_steps_per_s": 1e3 / ms_per_step, "samples_per_s": B * units * 1e3 / ms_per_step,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "nfe_per_s": nfe_total / (dev_ms * 1e-3), "sample_nfe_per_s": nfe_total * B * units / (dev_ms * 1e-3),
        "solver": ({"fwd_iters": out["fwd_iters"], "bwd_iters": out["bwd_iters"], "graph_redone": out["redone"]}
                   if not args.no_graph else {"fwd": out["fwd_stats"], "bwd": out["bwd_stats"][0]}),
        "launch_mode": "whole update() replayed as one CUDA graph (deferred-mode solves)" if not args.no_graph
                       else "polled solves",
        "roofline": roofline,
        "wall_s": wall,
    }
    if parity is not None:
        line["parity"] = parity
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        n_cpu = min(args.steps, 8)
        sec, info = cpu_train_steps(n_cpu, 1, threads)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"train steps 0..{n_cpu - 1} of the same trajectory (the GPU legs time steps "
                                          f"0..{args.steps - 1}), oracle port, torch CPU fp32, {threads} threads; "
                                          f"last step nfe {info['nfe']}, adjoint iterations {info['bwd_iters']}"}
    if rank == 0:
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        # Tear down in dependency order: the captured graph holds NCCL work of both communicators.  Drop it,
        # drain the device, meet the other ranks once more, then leave without running communicator
        # destructors (a destructor waiting on a peer that already left is the classic exit hang).
        model._graphs = {}
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
