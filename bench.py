#!/usr/bin/env python
"""Benchmark of the routed interaction stack (BASELINE.json metric: routed-interaction samples/sec,
forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # this repo's CUDA stack, BASELINE configs[1]
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the UNMODIFIED reference on the host cores
    python bench.py --config deep        # BASELINE configs[3]: R=4, K=6, text 256 + 197 image tokens, fwd+bwd
    python bench.py --config eval-sweep  # BASELINE configs[4]: eval / no_grad, batch 2 .. 4096 per GPU

Default workload (BASELINE.json configs[1]): both branch stacks (text branch + image branch, K=6 cells, R=3 routing
layers, text 128 + 50 image tokens, hidden 768), bf16 arithmetic with fp32 accumulation, batch 256 per GPU, one step
= forward + backward of both stacks (loss = out.sum() + sim_paths.sum() per branch) + the data-parallel
gradient all-reduce when N > 1.  Synthetic N(0,1) inputs, reference default-init weights.

One JSON line on stdout (rank 0).
  value        inputs resident in HBM, CUDA-graph replay, CUDA events, max over ranks
  e2e          the same step driven from pinned HOST buffers through the nn.Module API, H2D copy of the inputs and a
               D2H read of the loss inside the timed region
  roofline     all tensor-core launches of a step (tcgen05 GEMMs + fused attention kernels): algorithmic FLOPs
               (2*m*n*k per problem, no padding) over their CUDA-event durations, against the measured sustained bf16
               peak of MEASURED_PEAKS.json
  roofline_hbm the aggregation kernels (north_star (c)): algorithmic bytes over CUDA-event durations against the
               measured copy bandwidth
  cpu_baseline the UNMODIFIED reference modules (baseline/_ref, staged by __graft_entry__.build()) on this box's host
               cores, fp32, GPU hidden, bounded sample (batch 8); kind "reference".  Falls back to the oracle port
               (kind "port") only if the staged copy is missing.
  stock_pytorch_b200   the same unmodified reference modules on this B200 under torch.autocast(bfloat16), same batch,
               fwd+bwd, CUDA events: the honest "before" 
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

D, KC = 768, 6
CONFIGS = {
    # name: Lt, Li, R, batch per GPU, train?, BASELINE.json config index, description
    "stack": dict(Lt=128, Li=50, R=3, B=256, train=True, idx=1,
                  desc="D2R routed interaction stack, both branches (text+image), bf16 fwd+bwd, K=6 cells, "
                       "R=3 routing layers, text 128 + 50 image tokens, hidden 768 (BASELINE configs[1])"),
    "deep": dict(Lt=256, Li=197, R=4, B=256, train=True, idx=3,
                 desc="deep routing: both branches, bf16 fwd+bwd, K=6 cells, R=4 routing layers, text 256 + 197 image "
                      "tokens (ViT-B/16 patch count), hidden 768 (BASELINE configs[3])"),
    "full": dict(Lt=128, Li=50, R=3, B=64, train=True, idx=2,
                 desc="FULL D2R model (UnimoModelF: BERT-base + CLIP-ViT-B/32 towers on stock PyTorch feeding the routed "
                      "stacks, random init), data-parallel training step fwd+bwd+gradient all-reduce, bf16 autocast, "
                      "MVSA-shaped synthetic batches (max_seq 128, 224 px images -> 50 image tokens), R=3 "
                      "(BASELINE configs[2])"),
    "eval-sweep": dict(Lt=128, Li=50, R=3, B=256, train=False, idx=4,
                       desc="inference-only (eval, no_grad) throughput sweep, both branches, bf16, K=6, R=3, text 128 "
                            "+ 50 image tokens, batch 2..4096 per GPU (BASELINE configs[4])"),
}
SWEEP_BATCHES = (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096)
# DRAM traffic of the dominant launches from `ncu --set full` captures (see profiles/): per launch, like `achieved`
GEMM_TRAFFIC_BYTES = 62.3e6
GEMM_TRAFFIC_NOTE = ("dram__bytes_read+write of one m=32768 n=768 k=768 launch (51.6 MB read + 10.7 MB written; the K=768 "
                     "projections are 60% of the GEMM time); 101.8 MB algorithmic operand bytes, the output is mostly "
                     "still in L2 at kernel end; profiles/r02_ncu_gemm_k768_pair.txt")
AGG_TRAFFIC_BYTES = 454.6e6
AGG_TRAFFIC_NOTE = ("agg_fwd_kernel text non-final launch: 503.3 MB algorithmic, 454.6 MB dram (203.0 read + 251.6 "
                    "written); agg_bwd_kernel: 704.6 MB algorithmic, 687.5 MB dram (profiles/r02_ncu_agg_fwd_bwd.txt)")
CPU_SAMPLE_BATCH = 8


def make_args():
    return argparse.Namespace(embed_size=768, hid_router=768, hid_IMRC=768, num_head_IMRC=16,
                              raw_feature_norm_CMRC="clipped_l2norm", lambda_softmax_CMRC=4.0, alpha=0, margin=0.1,
                              bert_name="bert-base-uncased", vit_name="clip-vit-base-patch32")


def useful_flops_per_sample(Lt, Li, layers):
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
        return imrc + glac + cmrc + crcmc + gesc + router + agg
    return layers * (layer(Lt, Li) + layer(Li, Lt))


def metric_name(config):
    if config == "full":
        return "full-model samples/sec fwd+bwd (routed stack on d2r_b200, encoders stock PyTorch)"
    if config == "eval-sweep":
        return "routed-interaction samples/sec forward (eval, no_grad)"
    return "routed-interaction samples/sec fwd+bwd"


def workload_config(name, n_gpus, batch):
    c = CONFIGS[name]
    return {"workload": c["desc"], "batch_per_gpu": batch, "global_batch": batch * n_gpus, "text_len": c["Lt"],
            "image_tokens": c["Li"], "hidden": D, "cells": KC, "routing_layers": c["R"], "parallelism": f"dp{n_gpus}",
            "l2": "no explicit flush: the per-step working set (several GB of activations) is far larger than the "
                  "126 MB L2"}


# ----------------------------------------------------------------------------------- CPU arms
def run_reference_subprocess(device, batch, steps, warmup, cfg, bf16=False, timeout=900):
    """baseline/run_reference.py in a subprocess (GPU hidden for the CPU arm) -> its JSON line (dict) or raises."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_reference.py"), "--device", device, "--batch", str(batch),
           "--steps", str(steps), "--warmup", str(warmup), "--layers", str(cfg["R"]), "--text-len", str(cfg["Lt"]),
           "--image-tokens", str(cfg["Li"])]
    if bf16:
        cmd.append("--bf16")
    if not cfg["train"]:
        cmd.append("--eval")
    if cfg.get("idx") == 2:
        cmd.append("--full")
    env = dict(os.environ)
    if device == "cpu":
        env["CUDA_VISIBLE_DEVICES"] = ""
    # the arm must not inherit a launcher's rank variables or its thread caps (torchrun exports OMP_NUM_THREADS=1)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID", "GROUP_RANK",
              "ROLE_RANK", "LOCAL_WORLD_SIZE", "OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS",
              "GOMP_CPU_AFFINITY", "KMP_AFFINITY"):
        env.pop(k, None)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if out.returncode != 0:
        raise RuntimeError(f"run_reference.py rc={out.returncode}: {out.stderr.strip().splitlines()[-1:]}")
    return json.loads(out.stdout.strip().splitlines()[-1])


def oracle_port_step_fn(batch, cfg, threads=None):
    """Fallback CPU arm when baseline/_ref is not staged: the oracle port (reference algorithm incl. its dead branch)."""
    import torch
    from oracle import d2r_oracle as O
    if threads:
        torch.set_num_threads(threads)
    R = cfg["R"]
    Pt, Pi = O.make_params(2023, R, KC), O.make_params(2024, R, KC)
    for P in (Pt, Pi):
        for k, v in P.items():
            if v.is_floating_point() and "running" not in k and not O.is_dead_param(k):
                v.requires_grad_(True)
    text, image = O.make_inputs(2023, batch, cfg["Lt"], cfg["Li"])
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


def cpu_baseline(cfg, steps, warmup):
    """-> cpu_baseline dict.  Unmodified reference (kind "reference") when staged, else the oracle port."""
    from baseline import ref_loader as RL
    cores = os.cpu_count() or 1
    what = "fwd+bwd" if cfg["train"] else "eval/no_grad forward"
    if cfg.get("idx") == 2:
        what = "fwd+bwd of the WHOLE model (encoders + stacks + head)"
    if RL.available():
        try:
            r = run_reference_subprocess("cpu", CPU_SAMPLE_BATCH, steps, warmup, cfg)
            return {"value": r["samples_per_s_mean"], "unit": "samples/s", "cores": r["cores"], "threads": r["threads"],
                    "kind": "reference", "ms_per_step": 1e3 * r["mean_s_per_step"],
                    "sample": f"{steps} timed steps of batch {CPU_SAMPLE_BATCH} ({what} of both UNMODIFIED reference "
                              f"stacks, fp32, torch {r['torch']} CPU, GPU hidden), {warmup} warm-up; "
                              f"median {r['samples_per_s']:.2f} samples/s"}
        except Exception as e:      # fall through to the port, say why
            sys.stderr.write(f"[bench] reference CPU arm failed ({e}); timing the oracle port instead\n")
    step = oracle_port_step_fn(CPU_SAMPLE_BATCH, cfg, cores)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": CPU_SAMPLE_BATCH * steps / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "ms_per_step": 1e3 * dt / steps,
            "sample": f"{steps} timed steps of batch {CPU_SAMPLE_BATCH} (oracle port of both stacks, fwd+bwd, fp32, "
                      f"dead reverse-attention branch executed), {warmup} warm-up; baseline/_ref not staged"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # each step is a bounded sample (batch 8) of the workload: ~1-2 s of host time per step
    r = cpu_baseline(cfg, steps, warmup)
    line = {
        "impl": "reference", "metric": metric_name(args.config), "value": r["value"],
        "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, args.gpus, cfg["B"]),
        "cpu_baseline": r,
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU arm: `config` names the workload of the GPU arm; each timed step here is a bounded batch-8 sample of "
                "it (cpu_baseline.sample), run by rank 0 only on the host cores",
    }
    emit(line)
    return 0


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
    cfg = CONFIGS[args.config]
    LT, LI, R = cfg["Lt"], cfg["Li"], cfg["R"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch_per_gpu or cfg["B"]

    torch.manual_seed(2023)
    a = make_args()
    mt = InteractionModule(a, R, KC, 128).to(dev).train(cfg["train"])
    mi = Reversed_InteractionModule(a, R, KC, 128).to(dev).train(cfg["train"])
    if args.config == "eval-sweep":
        return eval_sweep(args, cfg, mt, mi, dev, rank, world, local)
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
    overlap = (args.overlap_allreduce or world > 1) and not args.no_overlap_allreduce
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
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples cover warm-up + the timed region (both under full load)
    for _ in range(max(args.warmup, 3)):
        step(False)
    ms = timed(False, args.steps)
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
    ms_e2e = timed(True, args.steps)
    # kernels launched from a replayed graph do not pass through the C ABI's counter: count one eager step
    n0 = K.L.launch_count()
    fwd_bwd(reduce=False)
    torch.cuda.synchronize()
    launches_per_step = K.L.launch_count() - n0

    # --- per-kernel roofline pass (separate, eager, CUDA events around every tensor-core / aggregation launch) -----
    # (branches back to back here: kernels of concurrent streams would overlap inside each other's event pairs)
    peaks = load_peaks()
    tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roof, hbm = (kernel_roofline(lambda: fwd_bwd(serial=True, reduce=False), K, torch, tc_peak, hbm_peak)
                 if rank == 0 else (None, None))

    if rank == 0:
        peak_src = "measured (MEASURED_PEAKS.json, sustained bf16)" if peaks else "fallback"
        samples = B * world * args.steps
        value = samples / (ms / 1e3)
        flops_step = 3 * useful_flops_per_sample(LT, LI, R) * B
        graph_used = graph is not None
        graph = static_loss = None            # release the graph's private pool before the stock-PyTorch run
        torch.cuda.empty_cache()
        # the two baselines are measured at N = 1 only (rank 0 would keep N - 1 GPUs idle for a minute or two; and under
        # torchrun the CPU arm inherits the rank's restricted CPU affinity: 0.3-0.4 samples/s instead of 25 in round 2)
        skipped = {"value": None, "skipped": "measured at --gpus 1 only"}
        cpu = cpu_baseline(cfg, steps=5, warmup=1) if world == 1 else dict(skipped)
        stock = stock_pytorch_b200(cfg, B) if world == 1 else dict(skipped)
        if stock and stock.get("value"):
            stock["speedup_of_this_repo"] = value / world / stock["value"]      # per GPU vs one B200
        line = {
            "metric": metric_name(args.config), "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.config, world, B),
            "engine": {"cuda_graph": graph_used, "branch_streams": 1 if args.serial_branches else 2,
                       "cell_lanes": LN.CELL_LANES,
                       "allreduce": "layer-wise, inside the backward" if overlap else "one per step"},
            "clocks": clocks,
            "e2e": {"value": samples / (ms_e2e / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": (h_text.numel() + h_image.numel()) * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "roofline": {
                "bound": "tensor",
                "kernel": "gemm_tc2_kernel / gemm_tc_kernel: the tensor-bound tcgen05 launches of one fwd+bwd step "
                          "(arithmetic intensity above the ridge point: the projections, their data- and weight-gradients)",
                "achieved": roof["classes"]["tensor"]["achieved_tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                "frac": roof["classes"]["tensor"]["frac_of_tensor_peak"],
                "traffic": GEMM_TRAFFIC_BYTES, "traffic_note": GEMM_TRAFFIC_NOTE, "peak_source": peak_src,
                "launches_per_step": roof["classes"]["tensor"]["launches"],
                "kernel_ms_per_step": roof["classes"]["tensor"]["ms"],
                "single_stream_step_ms": roof["serial_step_ms"],
                "kernel_share_of_step": roof["classes"]["tensor"]["ms"] / roof["serial_step_ms"],
                "ridge_flop_per_byte": roof["ridge_flop_per_byte"],
                "classes": roof["classes"],
                "all_tcgen05_launches": {"launches_per_step": roof["launches"], "kernel_ms_per_step": roof["ms"],
                                         "achieved": roof["tflops"], "frac": roof["tflops"] / tc_peak,
                                         "note": "round-1 definition: FLOPs of every tcgen05 launch (HBM- and "
                                                 "latency-bound ones included) over their summed time"},
                "vendor_calibration": "cuBLAS on this GPU, same shapes, isolated: 944 TF/s (m=32768 n=768 k=768; this "
                                      "kernel 899), 737 (m=12800; 567), 1490 (k=4608; 1326), 1333 (n=4608; 1326): "
                                      "profiles/r02_vendor_calibration.txt",
                "step_algorithmic_tflops": flops_step / (ms / args.steps / 1e3) / 1e12,
                "step_frac_of_peak": flops_step / (ms / args.steps / 1e3) / 1e12 / tc_peak,
            },
            "roofline_hbm": {"bound": "hbm", "kernel": "agg_fwd_kernel / agg_bwd_kernel (aggregation epilogue)",
                             "achieved": hbm["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm["gbs"] / hbm_peak,
                             "launches_per_step": hbm["launches"], "kernel_ms_per_step": hbm["ms"],
                             "traffic": AGG_TRAFFIC_BYTES, "traffic_note": AGG_TRAFFIC_NOTE},
            "cpu_baseline": cpu,
            "stock_pytorch_b200": stock,
        }
        emit(line)
    # every rank drops its CUDA graph (it holds captured NCCL kernels in the overlapped mode) before the
    # communicator goes away: destroying the process group under a live graph was round 1's "unexplained error
    # after the result line"
    graph = static_loss = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def full_arm(args):
    """BASELINE configs[2]: the reference's whole model with the B200 stack swapped in (d2r_b200.integration.
    accelerate), one process per GPU, batch sharded, one flat gradient all-reduce per step.  Rank 0 also times the
    UNMODIFIED model on the same GPU (`stock_pytorch_b200`) before the swap.  Eager execution (the encoders are the
    reference's stock PyTorch code); `value` has the batch resident in HBM, `e2e` copies it from pinned host memory."""
    import torch
    import torch.distributed as dist
    from baseline.full_model import build_reference_model, synthetic_batch
    from d2r_b200 import kernels as K
    from d2r_b200.dp import BucketedGradReducer, FlatGradReducer
    from d2r_b200.integration import accelerate
    cfg = CONFIGS["full"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch_per_gpu or cfg["B"]
    model, _ = build_reference_model(cfg["R"], seed=2023)
    model = model.to(dev).train()
    host = [t.pin_memory() for t in synthetic_batch(B, cfg["Lt"], seed=2023 + rank)]
    devb = [t.to(dev) for t in host]

    def fwd_bwd(batch):
        for p in model.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, logits = model(*batch)
        loss.backward()
        return loss

    def time_steps(fn, steps, warm):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    stock = None
    if rank == 0:
        try:
            ms_stock = 0.0
            for _ in range(2):
                fwd_bwd(devb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                float(fwd_bwd(devb).detach())
            e1.record()
            torch.cuda.synchronize()
            ms_stock = e0.elapsed_time(e1) / 5
            stock = {"value": B / (ms_stock / 1e3), "unit": "samples/s", "ms_per_step": ms_stock, "batch": B,
                     "dtype": "torch.autocast(bfloat16)",
                     "note": "the UNMODIFIED UnimoModelF (baseline/_ref) on this GPU, eager PyTorch, fwd+bwd+loss read, "
                             "5 steps after 2 warm-up, no all-reduce"}
        except Exception as e:
            stock = {"value": None, "unavailable": f"{type(e).__name__}: {e}"}
    accelerate(model, graph=not args.no_graph)
    graphed = not args.no_graph
    try:
        fwd_bwd(devb)                 # first call builds the CUDA graphs of the two stacks
        torch.cuda.synchronize()
    except Exception as e:
        sys.stderr.write(f"[bench] graphing the stacks failed ({type(e).__name__}: {e}); eager stacks\n")
        torch.cuda.synchronize()
        model.model.itr_module.__dict__["_d2r_graph"] = False
        graphed = False
        fwd_bwd(devb)
        torch.cuda.synchronize()
    # Default: FlatGradReducer -- one packing copy + one all-reduce of the live gradients after the backward (the
    # variant measured on the GPU at N = 1 and N = 2).  --bucketed-allreduce: gradients are views of one flat buffer and
    # bucket all-reduces are launched by hooks from inside the backward (gloo-tested; on the GPU only run at N = 1).
    bucketed = bool(getattr(args, "bucketed_allreduce", False))
    if bucketed:
        reducer = BucketedGradReducer(model.parameters(), bucket_mb=128.0)
        reducer.plan()
    else:
        reducer = FlatGradReducer(model.parameters())
    h_loss = torch.empty(1).pin_memory()

    def fwd_bwd_dp(batch):
        if not bucketed:
            loss = fwd_bwd(batch)
            reducer.step()
            return loss
        reducer.begin_step()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, logits = model(*batch)
        loss.backward()
        reducer.finish()
        return loss

    def step(from_host=False):
        if from_host:
            for d, h in zip(devb, host):
                d.copy_(h, non_blocking=True)
        loss = fwd_bwd_dp(devb)
        if from_host:
            h_loss.copy_(loss.detach().float().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = K.L.launch_count()
    step()
    torch.cuda.synchronize()
    launches_per_step = K.L.launch_count() - n0
    ms = time_steps(step, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = time_steps(lambda: step(True), args.steps, 2)
    # share of the routed stacks in the step: the two stacks alone, same batch, same precision
    ms_stack = None
    if bucketed:
        reducer.remove()              # (no collectives from the rank-0-only measurement below)
    if rank == 0:
        bb = model.model
        t_in = torch.randn(B, cfg["Lt"], D, device=dev, requires_grad=True)
        i_in = torch.randn(B, cfg["Li"], D, device=dev, requires_grad=True)

        def stack_only():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                o1, s1 = bb.itr_module(t_in, i_in)
                o2, s2 = bb.Reversed_itr_module(t_in, i_in)
            (o1[0].sum() + s1.sum() + o2[0].sum() + s2.sum()).backward()
        for _ in range(3):
            stack_only()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            stack_only()
        e1.record()
        torch.cuda.synchronize()
        ms_stack = e0.elapsed_time(e1) / 5
    if rank == 0:
        samples = B * world * args.steps
        value = samples / (ms / 1e3)
        if stock and stock.get("value"):
            stock["speedup_of_this_repo"] = value / world / stock["value"]
        nparams = sum(p.numel() for p in model.parameters())
        line = {
            "metric": metric_name("full"), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config("full", world, B),
            "engine": {"cuda_graph": "the two stacks (forward graph + backward graph); encoders eager" if graphed else False,
                       "swap": "d2r_b200.integration.accelerate (stacks via run_pair, CLS poolers, Block fusion, js_div)",
                       "allreduce": (f"{len(reducer.buckets)} buckets of one flat fp32 gradient buffer, all-reduced from "
                                     f"inside the backward; {reducer.dead} never-used tensors excluded") if bucketed
                       else f"one flat fp32 bucket per step, {reducer.dead} never-used tensors excluded",
                       "parameters_M": nparams / 1e6},
            "clocks": clocks,
            "e2e": {"value": samples / (ms_e2e / 1e3), "unit": "samples/s",
                    "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "stack_only_ms_per_step": ms_stack,
            "stack_share_of_step": (ms_stack / (ms / args.steps)) if ms_stack else None,
            "roofline": {"bound": "tensor", "kernel": "whole step, reference-executed FLOPs of the model (SURVEY §7: 52.5 "
                                                      "GFLOP/sample forward, x3 for fwd+bwd)",
                         "achieved": value / world * 3 * 52.5e9 / 1e12,
                         "peak": load_peaks().get("bf16_tflops_sustained", 1400.0), "unit": "TFLOP/s",
                         "frac": value / world * 3 * 52.5e9 / 1e12 / load_peaks().get("bf16_tflops_sustained", 1400.0),
                         "traffic": None},
            "stock_pytorch_b200": stock,
            "cpu_baseline": cpu_baseline(cfg, steps=3, warmup=1) if world == 1
            else {"value": None, "skipped": "measured at --gpus 1 only"},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def stock_pytorch_b200(cfg, batch):
    """The unmodified reference modules on this B200 (stock PyTorch eager: cuBLAS + aten) under autocast(bf16)."""
    from baseline import ref_loader as RL
    if not RL.available():
        return {"value": None, "unavailable": "baseline/_ref not staged"}
    try:
        r = run_reference_subprocess("cuda", batch, 5, 2, cfg, bf16=True, timeout=600)
        return {"value": r["samples_per_s"], "unit": "samples/s", "ms_per_step": 1e3 * r["median_s_per_step"],
                "batch": batch, "dtype": "torch.autocast(bfloat16)", "impl": r["impl"], "torch": r["torch"],
                "peak_mem_gb": r["peak_mem_gb"],
                "note": "UNMODIFIED reference InteractionModule + Reversed_InteractionModule, eager PyTorch on the same "
                        "GPU, same shapes, median of 5 steps after 2 warm-up, CUDA events around forward+backward+"
                        "loss read"}
    except Exception as e:
        return {"value": None, "unavailable": f"{type(e).__name__}: {e}"}


def eval_sweep(args, cfg, mt, mi, dev, rank, world, local):
    """BASELINE configs[4]: eval / no_grad forward of both stacks, batch 2..4096 per GPU, each replayed from its own
    CUDA graph.  `value` = the best whole-job samples/s of the sweep; every point is in `sweep`."""
    import torch
    import torch.distributed as dist
    from d2r_b200 import kernels as K
    from d2r_b200.interaction import run_pair
    LT, LI, R = cfg["Lt"], cfg["Li"], cfg["R"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    points = []
    h_out = torch.empty(1).pin_memory()
    for B in SWEEP_BATCHES:
        try:
            g = torch.Generator().manual_seed(2023 + rank)
            h_text = torch.randn(B, LT, D, generator=g).pin_memory()
            h_image = torch.randn(B, LI, D, generator=g).pin_memory()
            d_text, d_image = h_text.to(dev), h_image.to(dev)

            def fwd():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    (o1, s1), (o2, s2) = run_pair(mt, mi, d_text, d_image)
                return o1[0].sum() + o2[0].sum()
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(2):
                    fwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            n0 = K.L.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                res = fwd()
            launches = K.L.launch_count() - n0
            for _ in range(max(args.warmup, 3)):
                graph.replay()

            def timed(from_host):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    if from_host:
                        d_text.copy_(h_text, non_blocking=True)
                        d_image.copy_(h_image, non_blocking=True)
                    graph.replay()
                    if from_host:
                        h_out.copy_(res.detach().reshape(1), non_blocking=True)
                        torch.cuda.current_stream().synchronize()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ms], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                return ms
            ms, ms_h = timed(False), timed(True)
            points.append({"batch_per_gpu": B, "samples_per_s": B * world * args.steps / (ms / 1e3),
                           "ms_per_step": ms / args.steps, "e2e_samples_per_s": B * world * args.steps / (ms_h / 1e3),
                           "gpu_launches_per_step": int(launches)})
            del graph, res, d_text, d_image
            torch.cuda.empty_cache()
        except torch.OutOfMemoryError:
            points.append({"batch_per_gpu": B, "samples_per_s": None, "note": "out of memory"})
            torch.cuda.empty_cache()
            break
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        ok = [p for p in points if p.get("samples_per_s")]
        best = max(ok, key=lambda p: p["samples_per_s"])
        peaks = load_peaks()
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        fl = useful_flops_per_sample(LT, LI, R)
        for p in ok:
            p["frac_of_tensor_peak"] = p["samples_per_s"] / world * fl / 1e12 / tc_peak
        cpu = cpu_baseline(cfg, steps=5, warmup=1) if world == 1 else {"value": None, "skipped": "measured at --gpus 1 only"}
        line = {"metric": metric_name("eval-sweep"), "value": best["samples_per_s"],
                "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args.config, world, best["batch_per_gpu"]), "sweep": points, "clocks": clocks,
                "e2e": {"value": best["e2e_samples_per_s"], "unit": "samples/s",
                        "h2d_bytes_per_step": best["batch_per_gpu"] * (LT + LI) * D * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": best["gpu_launches_per_step"] * args.steps,
                "roofline": {"bound": "tensor", "kernel": "whole forward step (useful FLOPs, SURVEY §8d)",
                             "achieved": best["samples_per_s"] / world * fl / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                             "frac": best["samples_per_s"] / world * fl / 1e12 / tc_peak, "traffic": None},
                "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def kernel_roofline(fwd_bwd, K, torch, tc_peak, hbm_peak):
    """Two eager single-stream steps with CUDA events around every tensor-core launch (d2r_gemm on bf16 operands,
    fused attention) and every aggregation launch.  Each tensor-core launch is put in a roofline class by its
    ARITHMETIC INTENSITY (algorithmic FLOPs / algorithmic bytes, every distinct operand read or written once):
      tensor   intensity >= ridge point (measured bf16 peak / measured HBM bandwidth, ~209 FLOP/B): the projections
      hbm      intensity below the ridge: the attention products (34 .. 130 FLOP/B) -- bandwidth-bound by the model
      latency  launches whose roofline time is under 4 us (the [B,768] global-branch GEMMs): bound by launch and
               pipeline-fill latency, reported with their count and time only
    """
    ridge = tc_peak * 1e12 / (hbm_peak * 1e9)
    recs = {"gemm": [], "agg": []}
    orig = {n: getattr(K, n) for n in ("gemm", "aggregate_fwd", "aggregate_bwd")}
    fused = {n: getattr(K, n) for n in ("attn_fused_fwd", "attn_fused_bwd") if hasattr(K, n)}

    def timed_call(kind, work, fn, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        recs[kind].append((e0, e1, work))
        return r

    def gemm(a, b, c, **kw):
        if a.dtype != torch.bfloat16:
            return orig["gemm"](a, b, c, **kw)
        m, n, k, z = kw["m"], kw["n"], kw["k"], kw.get("batch", 1)
        byts = z * ((m * k + n * k) * 2 + m * n * c.element_size())
        if kw.get("residual") is not None:
            byts += z * m * n * kw["residual"].element_size()
        return timed_call("gemm", (2.0 * m * n * k * z, float(byts)), orig["gemm"], a, b, c, **kw)

    def agg_f(full, bvec, P, gate, final, inputs=None, want_pooled=True):
        x0 = full[0]
        nfull = sum(f is not None for f in full)
        n_out = 1 if final else len(full)
        byts = (nfull + n_out) * x0.numel() * x0.element_size()
        return timed_call("agg", byts, orig["aggregate_fwd"], full, bvec, P, gate, final, inputs, want_pooled)

    def agg_b(full, bvec, P, gate, final, d_outs, d_pooled, inputs=None):
        x0 = full[0]
        nfull = sum(f is not None for f in full)
        n_out = 1 if final else len(full)
        byts = (n_out + 2 * nfull) * x0.numel() * x0.element_size()
        return timed_call("agg", byts, orig["aggregate_bwd"], full, bvec, P, gate, final, d_outs, d_pooled, inputs)

    def fused_wrap(fn, fwd):
        # fused attention: 2 (forward: QK^T, PV) or 4 (backward: dP, dV, dQ, dK) products of 2 B Lq Lc D FLOP each
        def w(*a, **kw):
            B_, Lq, Lc, D_, H = kw["B"], kw["Lq"], kw["Lc"], kw["D"], kw["heads"]
            if fwd:
                el = Lq * D_ * (3 if kw.get("residual") is not None else 2) + 2 * Lc * D_ + H * Lq * Lc
            else:
                el = 3 * Lq * D_ + 4 * Lc * D_ + H * Lq * Lc
            return timed_call("gemm", ((2 if fwd else 4) * 2.0 * B_ * Lq * Lc * D_, 2.0 * B_ * el), fn, *a, **kw)
        return w

    import d2r_b200.lanes as LN
    lanes_were = LN.ENABLED
    LN.ENABLED = False                  # one stream: concurrent lanes would overlap inside each other's event pairs
    K.gemm, K.aggregate_fwd, K.aggregate_bwd = gemm, agg_f, agg_b
    for n, fn in fused.items():
        setattr(K, n, fused_wrap(fn, n.endswith("fwd")))
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
        K.gemm, K.aggregate_fwd, K.aggregate_bwd = orig["gemm"], orig["aggregate_fwd"], orig["aggregate_bwd"]
        for n, fn in fused.items():
            setattr(K, n, fn)
        LN.ENABLED = lanes_were
    cls = {c: {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0} for c in ("tensor", "hbm", "latency")}
    for e0, e1, (fl, by) in recs["gemm"]:
        t_roof_us = max(fl / (tc_peak * 1e12), by / (hbm_peak * 1e9)) * 1e6
        c = "latency" if t_roof_us < 4.0 else ("tensor" if fl / by >= ridge else "hbm")
        d = cls[c]
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    for d in cls.values():
        for k in d:
            d[k] /= nsteps
    gms = sum(d["ms"] for d in cls.values())
    gfl = sum(d["flops"] for d in cls.values())
    ams = sum(e0.elapsed_time(e1) for e0, e1, _ in recs["agg"]) / nsteps
    aby = sum(w for _, _, w in recs["agg"]) / nsteps
    t = cls["tensor"]
    h = cls["hbm"]
    return ({"tflops": gfl / (gms / 1e3) / 1e12, "ms": gms, "launches": int(sum(d["launches"] for d in cls.values())),
             "serial_step_ms": serial_ms, "ridge_flop_per_byte": ridge,
             "classes": {
                 "tensor": {"launches": int(t["launches"]), "ms": t["ms"],
                            "achieved_tflops": t["flops"] / (t["ms"] / 1e3) / 1e12 if t["ms"] else None,
                            "frac_of_tensor_peak": t["flops"] / (t["ms"] / 1e3) / 1e12 / tc_peak if t["ms"] else None},
                 "hbm": {"launches": int(h["launches"]), "ms": h["ms"],
                         "achieved_gbs": h["bytes"] / (h["ms"] / 1e3) / 1e9 if h["ms"] else None,
                         "frac_of_hbm_peak": h["bytes"] / (h["ms"] / 1e3) / 1e9 / hbm_peak if h["ms"] else None,
                         "achieved_tflops": h["flops"] / (h["ms"] / 1e3) / 1e12 if h["ms"] else None},
                 "latency": {"launches": int(cls["latency"]["launches"]), "ms": cls["latency"]["ms"]}}},
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
    ap.add_argument("--config", default="stack", choices=sorted(CONFIGS),
                    help="stack = BASELINE configs[1] (the headline); full = configs[2] (whole model, encoders on stock "
                         "PyTorch); deep = configs[3]; eval-sweep = configs[4]")
    ap.add_argument("--serial-branches", action="store_true",
                    help="call the two branch modules back to back instead of run_pair (two CUDA streams)")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="(default when N > 1) layer-wise gradient all-reduces launched from inside the backward "
                         "(GradAllReducer.install), captured in the CUDA graph")
    ap.add_argument("--no-overlap-allreduce", action="store_true",
                    help="one gradient all-reduce per step, issued after the graph replay (round-1 behaviour)")
    ap.add_argument("--bucketed-allreduce", action="store_true",
                    help="--config full: BucketedGradReducer (bucket all-reduces launched by hooks from inside the "
                         "backward) instead of one flat all-reduce after it")
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
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.config == "full":
        return full_arm(args)
    return own_arm(args)


if __name__ == "__main__":
    sys.exit(main())
