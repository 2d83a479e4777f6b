#!/usr/bin/env python
"""Benchmark of the routed interaction stack (BASELINE.json metric: routed-interaction samples/sec,
forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # this repo's CUDA stack
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the reference algorithm on CPU

Workload (BASELINE.json configs[1]): both branch stacks (text branch + image branch, K=6 cells, R=3 routing
layers, text 128 + 50 image tokens, hidden 768), bf16 arithmetic with fp32 accumulation, batch 256 per
GPU, one step = forward + backward of both stacks (loss = out.sum() + sim_paths.sum() per branch,
SURVEY §8d) + the data-parallel gradient all-reduce when N > 1.  Synthetic N(0,1) inputs, reference
default-init weights.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step driven from
pinned HOST buffers through the nn.Module API with the H2D copy of the inputs and a D2H read of the loss
inside the timed region.  `roofline`: all tcgen05 GEMM launches of a step, algorithmic FLOPs (2*m*n*k per
problem, no padding) over their CUDA-event durations, against the measured bf16 peak in MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port (oracle/d2r_oracle.py, with the reference's discarded reverse-attention
branch executed so that it pays the reference's real cost) on this box's host cores, bounded sample.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LT, LI, D, KC, R = 128, 50, 768, 6, 3
BATCH_PER_GPU = 256
# DRAM traffic of the dominant GEMM launch (ncu --set full, profiles/r01_ncu_gemm_k768_pair.txt): the m=32768, n=768,
# k=768 projection reads 51.6 MB and writes 10.4 MB per launch against 101.8 MB of algorithmic operand bytes (the
# bf16 output mostly stays in L2): no wasted re-reads.
GEMM_TRAFFIC_BYTES = 62.0e6
GEMM_TRAFFIC_NOTE = ("dram__bytes_read+write of one m=32768 n=768 k=768 launch (60% of the GEMM time is this shape "
                     "class); ncu capture of round 1, see profiles/r01_ncu_gemm_k768_pair.txt")
AGG_TRAFFIC_BYTES = 461.3e6
CPU_SAMPLE_BATCH = 8


def make_args():
    return argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                              bert_name="bert-base-uncased", vit_name="clip-vit-base-patch32")


def useful_flops_per_sample(Lt=LT, Li=LI, layers=R):
    """useful forward FLOPs per sample for both branches (dead branch excluded); bwd = 2x fwd."""
    def layer(Lq, Lc):
        lin = lambda L_: 2 * L_ * D * D
        cma = lin(Lq) + 2 * lin(Lc) + 4 * Lq * Lc * D
        imrc = 5 * lin(Lq) + 4 * Lq * Lq * D
        glac = cma + 2 * lin(Lq) + 8 * D * D + 4 * (Lq + 1) * D
        cmrc = cma + 4 * lin(Lq)
        crcmc = cma + 4 * lin(Lq) + 4 * Lq * Lq * D
        gesc = 8 * D * D
        router = 6 * (2 * D * 768 + 2 * 768 * KC + Lq * D)
        agg = 2 * KC * KC * Lq * D
        return cma * 0 + imrc + glac + cmrc + crcmc + gesc + router + agg
    return layers * (layer(Lt, Li) + layer(Li, Lt))


# ----------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_fn(batch, threads=None):
    """One fwd+bwd of both branch stacks with the oracle port (reference algorithm incl. its dead branch)."""
    import torch
    from oracle import d2r_oracle as O
    if threads:
        torch.set_num_threads(threads)
    Pt = O.make_params(2023, R, KC)
    Pi = O.make_params(2024, R, KC)
    for P in (Pt, Pi):
        for k, v in P.items():
            if v.is_floating_point() and "running" not in k and not O.is_dead_param(k):
                v.requires_grad_(True)
    text, image = O.make_inputs(2023, batch, LT, LI)
    text.requires_grad_(True)
    image.requires_grad_(True)

    def step():
        for P in (Pt, Pi):
            for v in P.values():
                v.grad = None
        text.grad = None
        image.grad = None
        o1, s1, _ = O.stack_forward(Pt, text, image, R, KC, False, True, {}, dead_branch=True)
        o2, s2, _ = O.stack_forward(Pi, text, image, R, KC, True, True, {}, dead_branch=True)
        loss = o1[0].sum() + s1.sum() + o2[0].sum() + s2.sum()
        loss.backward()
        return float(loss)
    return step


def run_cpu_reference(steps, warmup, batch=CPU_SAMPLE_BATCH):
    import torch
    cores = os.cpu_count() or 1
    step = cpu_reference_step_fn(batch, cores)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return dict(value=batch * steps / total, ms_per_step=1e3 * total / steps, cores=cores, threads=torch.get_num_threads(),
                sample=f"{steps} timed steps of batch {batch} (both branch stacks fwd+bwd, fp32, torch CPU, "
                       f"dead reverse-attention branch executed), {warmup} warm-up")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    steps = max(1, args.steps)
    # bound the run: each step is ~B=8 samples; never more than ~3 minutes of CPU work
    r = run_cpu_reference(min(steps, 40), min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": "routed-interaction samples/sec fwd+bwd", "value": r["value"],
        "unit": "samples/s", "n_gpus": args.gpus, "steps": min(steps, 40), "warmup": min(args.warmup, 3),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus) | {"cpu_sample_batch": CPU_SAMPLE_BATCH},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(n_gpus):
    return {"workload": "D2R routed interaction stack, both branches (text+image), bf16 fwd+bwd, K=6 cells, "
                        "R=3 routing layers, text 128 + 50 image tokens, hidden 768 (BASELINE configs[1])",
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * n_gpus, "text_len": LT, "image_tokens": LI,
            "hidden": D, "cells": KC, "routing_layers": R, "parallelism": f"dp{n_gpus}",
            "l2": "no explicit flush: the per-step working set (several GB of activations) is far larger "
                  "than the 126 MB L2"}


# ----------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["stdbuf", "-oL", "nvidia-smi", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True, bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------- own arm
def own_arm(args):
    import torch
    import torch.distributed as dist
    from d2r_b200 import kernels as K
    from d2r_b200.dp import GradAllReducer, InputPrefetcher
    from d2r_b200.interaction import InteractionModule, Reversed_InteractionModule, run_pair

    import d2r_b200.lanes as LN
    if args.cell_lanes is not None:
        LN.CELL_LANES = max(1, args.cell_lanes)
    if args.priority:
        LN.PRIORITIZE_FIRST_BLOCK = True
    if args.no_aux_bias:
        LN.AUX_BIAS = False
    if args.aux_wgrad:
        LN.AUX_WGRAD = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = BATCH_PER_GPU

    torch.manual_seed(2023)
    a = make_args()
    mt = InteractionModule(a, R, KC, 128).to(dev).train()
    mi = Reversed_InteractionModule(a, R, KC, 128).to(dev).train()
    reducer = GradAllReducer([mt, mi])

    g = torch.Generator().manual_seed(2023 + rank)
    h_text = torch.randn(B, LT, D, generator=g).pin_memory()
    h_image = torch.randn(B, LI, D, generator=g).pin_memory()
    d_text = torch.empty(B, LT, D, device=dev, requires_grad=True)
    d_image = torch.empty(B, LI, D, device=dev, requires_grad=True)
    h_loss = torch.empty(1).pin_memory()
    with torch.no_grad():
        d_text.copy_(h_text)
        d_image.copy_(h_image)

    # gradient all-reduce: layer-wise, launched from inside the backward (overlapped), or one collective per step
    overlap = args.overlap_allreduce
    if overlap:
        reducer.install()

    def fwd_bwd(serial=args.serial_branches, reduce=True):
        reducer.active = overlap and reduce
        d_text.grad = None
        d_image.grad = None
        for m in (mt, mi):
            for p in m.parameters():
                p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if serial:
                o1, s1 = mt(d_text, d_image)
                o2, s2 = mi(d_text, d_image)
            else:
                # the two branch stacks of modeling_unimo.py:842-843, issued on two CUDA streams
                (o1, s1), (o2, s2) = run_pair(mt, mi, d_text, d_image)
        loss = o1[0].sum() + s1.sum() + o2[0].sum() + s2.sum()
        loss.backward()
        if reducer.active:
            reducer.wait()
        elif reduce:
            reducer.pack()
        return loss

    # --- optional CUDA graph of fwd+bwd (+ gradient packing) --------------------------------------
    graph, static_loss = None, None
    side = torch.cuda.Stream()
    use_graph = not args.no_graph
    with torch.cuda.stream(side):
        for _ in range(2):
            loss = fwd_bwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = fwd_bwd()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:   # report and continue eagerly
            sys.stderr.write(f"[bench] CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}\n")
            graph = None
            torch.cuda.synchronize()

    prefetch = InputPrefetcher([d_text, d_image])

    def step(from_host, more=False):
        """One fwd+bwd(+all-reduce).  from_host: this step's inputs come from the pinned host batch through the
        prefetcher (H2D on a copy stream, overlapped with the previous step's compute); `more`: start the next
        step's H2D copy before computing."""
        if from_host:
            prefetch.commit()
            if more:
                prefetch.fetch([h_text, h_image])
        if graph is not None:
            graph.replay()
            loss = static_loss
        else:
            loss = fwd_bwd()
        if not overlap:
            reducer.all_reduce()
        if from_host:
            h_loss.copy_(loss.detach().float().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = K.L.launch_count()
        e0.record()
        # `ncu --profile-from-start off` captures exactly the HBM-resident timed region (all threads: the
        # backward kernels are launched from autograd's worker thread, so an NVTX range would miss them)
        prof = bool(os.environ.get("D2R_PROFILE_RANGE")) and not from_host
        if prof:
            torch.cuda.cudart().cudaProfilerStart()
        if from_host:
            prefetch.fetch([h_text, h_image])     # step 0's inputs: inside the timed region, not overlapped
        for i in range(steps):
            step(from_host, more=i + 1 < steps)
        if prof:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = K.L.launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples cover warm-up + the timed region (both under full load)
    for _ in range(max(args.warmup, 3)):
        step(False)
    ms, launches = timed(False, args.steps)
    if ms < 1500.0:
        # keep the GPUs under the same load a little longer so that nvidia-smi (100 ms period) sees it; `ms` is the
        # max over ranks, so every rank runs the same number of extra steps (they contain a collective)
        for _ in range(int(1500.0 / max(ms / args.steps, 1e-3)) + 1):
            step(False)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    prefetch.fetch([h_text, h_image])
    step(True, more=True)
    step(True)
    ms_e2e, _ = timed(True, args.steps)
    eager_launches_per_step = None
    if graph is not None:
        # kernels launched from a replayed graph do not pass through the C ABI: count one eager step
        n0 = K.L.launch_count()
        fwd_bwd(reduce=False)
        torch.cuda.synchronize()
        eager_launches_per_step = K.L.launch_count() - n0
        launches = eager_launches_per_step * args.steps

    # --- per-kernel roofline pass (separate, eager, CUDA events around every C-ABI GEMM launch) -----
    # (branches back to back here: kernels of concurrent streams would overlap inside each other's event pairs)
    roof, hbm = kernel_roofline(lambda: fwd_bwd(serial=True, reduce=False), K, torch) if rank == 0 else (None, None)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json, sustained bf16)" if peaks else "fallback"
        samples = B * world * args.steps
        value = samples / (ms / 1e3)
        flops_step = 3 * useful_flops_per_sample() * B
        cpu = None
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-sample"], capture_output=True,
                                 text=True, timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
            cpu = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as e:
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"failed: {e}"}
        line = {
            "metric": "routed-interaction samples/sec fwd+bwd", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world) | {"cuda_graph": graph is not None,
                                                  "branch_streams": 1 if args.serial_branches else 2,
                                                  "cell_lanes": LN.CELL_LANES,
                                                  "allreduce": "layer-wise, inside the backward" if overlap
                                                  else "one per step"},
            "clocks": clocks,
            "e2e": {"value": samples / (ms_e2e / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": (h_text.numel() + h_image.numel()) * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, all launches of one fwd+bwd step)",
                "achieved": roof["tflops"], "peak": tc_peak, "unit": "TFLOP/s", "frac": roof["tflops"] / tc_peak,
                "traffic": GEMM_TRAFFIC_BYTES, "traffic_note": GEMM_TRAFFIC_NOTE, "peak_source": peak_src, "launches_per_step": roof["launches"],
                "kernel_ms_per_step": roof["ms"], "single_stream_step_ms": roof["serial_step_ms"],
                "kernel_share_of_step": roof["ms"] / roof["serial_step_ms"],
                "step_algorithmic_tflops": flops_step / (ms / args.steps / 1e3) / 1e12,
                "step_frac_of_peak": flops_step / (ms / args.steps / 1e3) / 1e12 / tc_peak,
            },
            "roofline_hbm": {"bound": "hbm", "kernel": "agg_fwd_kernel / agg_bwd_kernel (aggregation epilogue)",
                             "achieved": hbm["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm["gbs"] / hbm_peak,
                             "launches_per_step": hbm["launches"], "kernel_ms_per_step": hbm["ms"],
                             "traffic": AGG_TRAFFIC_BYTES,
                             "traffic_note": "agg_fwd_kernel text non-final launch: 503.3 MB algorithmic, 461.3 MB "
                                             "dram (profiles/r01_ncu_agg_fwd.txt)"},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def kernel_roofline(fwd_bwd, K, torch):
    """Run two eager steps with CUDA events around each tcgen05 GEMM and each aggregation launch."""
    recs = {"gemm": [], "agg": []}
    orig_gemm, orig_af, orig_ab = K.gemm, K.aggregate_fwd, K.aggregate_bwd

    def timed_call(kind, work, fn, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        recs[kind].append((e0, e1, work))
        return r

    def gemm(a, b, c, **kw):
        if a.dtype != torch.bfloat16:
            return orig_gemm(a, b, c, **kw)
        return timed_call("gemm", 2.0 * kw["m"] * kw["n"] * kw["k"] * kw.get("batch", 1), orig_gemm, a, b, c, **kw)

    def agg_f(full, bvec, P, gate, final, inputs=None, want_pooled=True):
        x0 = full[0]
        nfull = sum(f is not None for f in full)
        n_out = 1 if final else len(full)
        byts = (nfull + n_out) * x0.numel() * x0.element_size()
        return timed_call("agg", byts, orig_af, full, bvec, P, gate, final, inputs, want_pooled)

    def agg_b(full, bvec, P, gate, final, d_outs, d_pooled, inputs=None):
        x0 = full[0]
        nfull = sum(f is not None for f in full)
        n_out = 1 if final else len(full)
        byts = (n_out + 2 * nfull) * x0.numel() * x0.element_size()
        return timed_call("agg", byts, orig_ab, full, bvec, P, gate, final, d_outs, d_pooled, inputs)

    import d2r_b200.lanes as LN
    lanes_were = LN.ENABLED
    LN.ENABLED = False                  # one stream: concurrent lanes would overlap inside each other's event pairs
    K.gemm, K.aggregate_fwd, K.aggregate_bwd = gemm, agg_f, agg_b
    serial_ms = 0.0
    try:
        fwd_bwd()                       # warm
        torch.cuda.synchronize()
        for v in recs.values():
            v.clear()
        nsteps = 2
        for _ in range(nsteps):
            # park the GPU for ~80 ms so that the host enqueues the whole step ahead of it: the events then
            # bracket back-to-back kernel execution, not host launch latency
            torch.cuda._sleep(int(0.08 * 1.9e9))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            fwd_bwd()
            s1.record()
            torch.cuda.synchronize()
            serial_ms += s0.elapsed_time(s1) / nsteps
    finally:
        K.gemm, K.aggregate_fwd, K.aggregate_bwd = orig_gemm, orig_af, orig_ab
        LN.ENABLED = lanes_were
    gms = sum(e0.elapsed_time(e1) for e0, e1, _ in recs["gemm"]) / nsteps
    gfl = sum(w for _, _, w in recs["gemm"]) / nsteps
    ams = sum(e0.elapsed_time(e1) for e0, e1, _ in recs["agg"]) / nsteps
    aby = sum(w for _, _, w in recs["agg"]) / nsteps
    return ({"tflops": gfl / (gms / 1e3) / 1e12, "ms": gms, "launches": len(recs["gemm"]) // nsteps,
             "serial_step_ms": serial_ms},
            {"gbs": aby / (ams / 1e3) / 1e9, "ms": ams, "launches": len(recs["agg"]) // nsteps})


_JSON_OUT = None


def reserve_stdout():
    """Keep the process' stdout for the ONE JSON line: libraries (NCCL's version banner, ...) write to file
    descriptor 1 directly, so fd 1 is pointed at stderr and the original stdout is kept aside for emit()."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--serial-branches", action="store_true",
                    help="call the two branch modules back to back instead of run_pair (two CUDA streams)")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="layer-wise gradient all-reduces launched from inside the backward (GradAllReducer.install) "
                         "instead of one collective after it; measured equal at 2 GPUs in round 1, not the default")
    ap.add_argument("--aux-wgrad", action="store_true", help="weight-gradient GEMMs on the helper stream as well")
    ap.add_argument("--no-aux-bias", action="store_true",
                    help="bias-gradient column sums on the GEMMs' own stream instead of a helper stream")
    ap.add_argument("--priority", action="store_true",
                    help="run_pair with high stream priority for the first (text) stack")
    ap.add_argument("--batch-per-gpu", type=int, default=None,
                    help="diagnostic only: override the per-GPU batch (the benchmark configuration is 256)")
    ap.add_argument("--cell-lanes", type=int, default=None,
                    help="CUDA streams per routing layer (default: d2r_b200.lanes.CELL_LANES; 1 = one stream)")
    ap.add_argument("--no-graph", action="store_true", help="run eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.batch_per_gpu:
        global BATCH_PER_GPU
        BATCH_PER_GPU = args.batch_per_gpu
    if args.cpu_sample:
        r = run_cpu_reference(steps=4, warmup=1)
        emit({"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]})
        return 0
    if args.impl == "reference":
        return reference_arm(args)
    return own_arm(args)


if __name__ == "__main__":
    sys.exit(main())
