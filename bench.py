#!/usr/bin/env python
"""Benchmark of the SPARF ray-marching hot path (BASELINE.json metric: rays/s, fwd+bwd, 128 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine auto|simt_fp32|tc_3x]

One "step" = one pass of the hot path over one synthetic ray batch of BASELINE config 2 ("DTU 3-view,
fixed GT poses, 1024 rays x 128 samples": 3 x 341 = 1023 rays of 300x400 views, coarse network):
ray generation -> depth sampling -> positional encoding -> 8x256 MLP + colour head -> compositing ->
photometric Huber loss -> backward to every MLP weight (gradients zeroed each step).

Timing: W warm-up steps, then K steps, each bracketed by CUDA events on the launching stream with an
L2 flush (256 MiB memset) between steps; ms_per_step = mean of the K intervals; multi-GPU: barrier +
synchronize on both sides and the MAX over ranks.  Clocks/throttle reasons are sampled with nvidia-smi
during the timed region.  Prints ONE JSON line (rank 0).
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
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

# algorithmic work: MACs per MLP sample-evaluation, FLOP fwd+bwd = 3 * 2 * MACs
MACS_PER_SAMPLE = 63 * 256 + 3 * 256 * 256 + 319 * 256 + 2 * 256 * 256 + 256 * 257 + 283 * 128 + 128 * 3  # 527 872
FLOP_PER_SAMPLE_FWD_BWD = 6 * MACS_PER_SAMPLE  # 3 167 232
B_VIEWS, H_IMG, W_IMG, RAYS_PER_VIEW, S_COARSE = 3, 300, 400, 341, 128
WORKLOAD = "DTU-shaped 3 views 300x400, fixed GT poses, 3x341=1023 rays x 128 coarse samples, photometric loss, fwd+bwd"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    hbm_gbs=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self, wait_s=5.0):
        """Launch `nvidia-smi -lms 50` and block until its first row arrives (it needs ~0.5 s to come up; the timed
        region of a short run would otherwise be over before the first sample)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < wait_s:
                time.sleep(0.02)
            self.rows.clear()   # samples from here on fall inside the measured regions
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------ workload
def make_workload(device, seed=0):
    import common
    from sparf_b200.renderer import Graph
    opt = common.make_opt(S=S_COARSE, fine=False, rand_rays=1024, stratified=True, noise=False)
    torch.manual_seed(seed)
    np.random.seed(seed)
    data = common.make_scene(seed, B_VIEWS, H_IMG, W_IMG, focal=400.0)
    data.depth_range = torch.tensor([[1.2, 5.2]] * B_VIEWS)
    for k in ("image", "intr", "pose", "depth_range", "idx"):
        data[k] = data[k].to(device)
    net = Graph(opt, device)
    net.train()
    return opt, data, net


def flat_grads(net):
    """Point every parameter's .grad at a view of ONE flat fp32 buffer (single all-reduce, single zero_); the MLP
    backward kernels then accumulate straight into it (sparf_b200.distributed.FlatGradients)."""
    from sparf_b200.distributed import FlatGradients
    return FlatGradients([net]).flat


def run_ours(args):
    import sparf_b200
    from sparf_b200 import _lib, ops
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sparf_b200.set_engine(args.engine)
    L = _lib.lib()
    opt, data, net = make_workload(dev, seed=rank)
    if world > 1:  # identical replicas of the MLP on every rank
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    flat = flat_grads(net)
    image_flat = data.image.reshape(B_VIEWS, 3, -1).permute(0, 2, 1).contiguous()  # [B,HW,3]
    HW = H_IMG * W_IMG
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_total = args.warmup + args.steps
    # per-step ray indices (each rank its own shard of the global batch: weak scaling, 1023 rays / GPU)
    idx_host = [torch.randperm(HW, generator=g)[:RAYS_PER_VIEW].pin_memory() for _ in range(2 * n_total)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.zeros((), pin_memory=True)

    def step(ray_idx_dev):
        flat.zero_()
        out = net.render_image_at_specific_rays(opt, data, iter=0, ray_idx=ray_idx_dev, mode="train")
        gt = image_flat[:, ray_idx_dev]
        loss = ops.huber2(out.rgb, gt)
        loss.backward()
        if world > 1:
            dist.all_reduce(flat)  # one NCCL all-reduce of [d theta] per step
        return loss

    def local_step(ray_idx_dev):   # everything of a step except the cross-rank exchange
        flat.zero_()
        out = net.render_image_at_specific_rays(opt, data, iter=0, ray_idx=ray_idx_dev, mode="train")
        gt = image_flat[:, ray_idx_dev]
        loss = ops.huber2(out.rgb, gt)
        loss.backward()
        return loss.detach()

    graphed = None
    if args.graph:
        from sparf_b200.graphs import GraphedStep
        static_idx = idx_host[0].to(dev)
        c0 = L.sparf_launch_count()
        graphed = GraphedStep(local_step, (static_idx,), warmup=3)
        launches_per_graph = (L.sparf_launch_count() - c0) // 4   # 3 eager warm-ups + the capture

    def graph_step(ray_idx_src):
        loss = graphed(ray_idx_src)
        if world > 1:
            dist.all_reduce(flat)
        return loss

    def timed(n_warm, n_steps, e2e, use_graph=False):
        times = []
        for i in range(n_warm + n_steps):
            src = idx_host[(n_total if e2e else 0) + i]
            if not e2e:
                ray_idx_dev = src.to(dev, non_blocking=True)
            flush.zero_()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if e2e:  # host buffers in, host scalar out, inside the timed region
                ray_idx_dev = src if use_graph else src.to(dev, non_blocking=True)
            loss = graph_step(ray_idx_dev) if use_graph else step(ray_idx_dev)
            if e2e:
                loss_host.copy_(loss.detach(), non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            if i >= n_warm:
                times.append(e0.elapsed_time(e1))
        t = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), times

    rays_per_step = B_VIEWS * RAYS_PER_VIEW * world
    ops.PROFILE.clear()
    timed(args.warmup, 0, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.sparf_launch_count()
    ops.PROFILE.clear()
    ops.PROFILE_ON[0] = True
    total_ms, times = timed(0, args.steps, False)
    ops.PROFILE_ON[0] = False
    launches = L.sparf_launch_count() - launches0
    mlp_ms = ops.profile_total_ms()   # MLP kernel groups timed with CUDA events in the eager pass (same kernels)
    eager_ms_per_step = total_ms / args.steps
    if graphed is not None:           # the reported step: one CUDA-graph replay per step
        timed(args.warmup, 0, False, use_graph=True)
        total_ms, times = timed(0, args.steps, False, use_graph=True)
        launches = launches_per_graph * args.steps
    e2e_ms, _ = timed(args.warmup, args.steps, True, use_graph=graphed is not None)
    clocks = sampler.stop() if rank == 0 else None   # sampled over the device-timed AND the end-to-end region
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    ms_per_step = total_ms / args.steps
    value = rays_per_step / (ms_per_step * 1e-3)
    flop_step_gpu = B_VIEWS * RAYS_PER_VIEW * S_COARSE * FLOP_PER_SAMPLE_FWD_BWD   # per GPU
    mlp_ms_per_step = mlp_ms / args.steps if mlp_ms else None
    achieved = flop_step_gpu / (mlp_ms_per_step * 1e-3) / 1e12 if mlp_ms_per_step else None
    roofline = dict(bound="tensor", kernel="MLP fwd+bwd kernels of one step (sparf_mlp_forward + sparf_mlp_backward)",
                    achieved=achieved, peak=peaks["bf16_tflops"], unit="TFLOP/s",
                    frac=(achieved / peaks["bf16_tflops"]) if achieved else None, peak_source=peaks["source"] + ", burst bf16",
                    frac_of_sustained=(achieved / peaks["bf16_tflops_sustained"]) if achieved and peaks["bf16_tflops_sustained"] else None,
                    flop_per_launch_group=flop_step_gpu, ms_per_launch_group=mlp_ms_per_step,
                    # dram__bytes_read.sum + dram__bytes_write.sum of the three dominant kernels of the group (taped
                    # forward 1.217 GB, dgrad 1.142 GB, wgrad 2.393 GB), one ncu --set full capture at this shape
                    traffic=4.752e9 if args.engine in ("auto", "tc_3x") else None,
                    traffic_source="profiles/r01_ncu_chain.md (ncu --set full, per step)",
                    engine=args.engine)
    cpu = cpu_baseline(sample_steps=2) if world == 1 else None   # reported on rank 0 at N = 1 only
    line = dict(metric="rays/sec (fwd+bwd, 128 samples/ray)", value=value, unit="rays/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 in/out; GEMMs: " + args.engine, data="synthetic",
                config=dict(workload=WORKLOAD, rays_per_gpu=B_VIEWS * RAYS_PER_VIEW, global_rays=rays_per_step,
                            samples_per_ray=S_COARSE, l2_flush_between_steps=True,
                            launch="one CUDA-graph replay per step" if graphed is not None else "eager",
                            eager_ms_per_step=eager_ms_per_step,
                            timing="mean of per-step CUDA-event intervals, max over ranks",
                            parallelism="dp%d (ray sharding, one NCCL all-reduce of MLP grads per step)" % world),
                clocks=clocks,
                e2e=dict(value=rays_per_step / (e2e_ms / args.steps * 1e-3), unit="rays/s",
                         h2d_bytes_per_step=RAYS_PER_VIEW * 8, d2h_bytes_per_step=4),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_step_fn():
    """The reference's CPU path, restated (oracle/sparf_oracle.py): same batch, fwd + photometric loss + bwd."""
    import common
    from oracle import sparf_oracle as O
    opt = common.make_opt(S=S_COARSE, fine=False, stratified=True)
    data = common.make_scene(0, B_VIEWS, H_IMG, W_IMG, focal=400.0)
    torch.manual_seed(0)
    sd = common.det_weights(opt, 0)
    params = {k: v.clone().requires_grad_(k != "progress") for k, v in sd.items()}
    drange = torch.tensor([1.2, 5.2])
    g = torch.Generator().manual_seed(0)

    def step():
        for p in params.values():
            p.grad = None
        ray_idx = torch.randperm(H_IMG * W_IMG, generator=g)[:RAYS_PER_VIEW]
        center, ray = O.rays_from_ray_idx(data.pose, data.intr, H_IMG, W_IMG, ray_idx)
        rand = torch.rand(B_VIEWS, RAYS_PER_VIEW, S_COARSE, 1, generator=g)
        out = O.render(opt, params, None, center, ray, drange, mode="train", rand=rand)
        loss = O.photometric_loss(out, data.image, ray_idx)
        loss.backward()
        return loss

    return step


def pick_cpu_threads(step):
    """torch CPU ops of this size do not scale to every core of a 100+ core host (the reference has the
    same behaviour): time one step at a few thread counts and keep the fastest."""
    cores = os.cpu_count() or 1
    best = (None, float("inf"))
    for n in sorted({min(cores, 16), min(cores, 32), min(cores, 64), cores}):
        torch.set_num_threads(n)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (n, dt)
    torch.set_num_threads(best[0])
    return best[0]


def cpu_baseline(sample_steps=2):
    cores = os.cpu_count() or 1
    step = oracle_step_fn()
    used = pick_cpu_threads(step)
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = (time.perf_counter() - t0) / sample_steps
    return dict(value=B_VIEWS * RAYS_PER_VIEW / dt, unit="rays/s", cores=used, host_cores=cores, kind="port",
                sample="%d full steps of the same 1023-ray x 128-sample batch through oracle/sparf_oracle.py "
                       "(torch CPU fp32, best of {16,32,64,all} threads = %d), %.2f s/step" % (sample_steps, used, dt))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step = oracle_step_fn()
    cores = pick_cpu_threads(step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = B_VIEWS * RAYS_PER_VIEW / dt
    line = dict(impl="reference", metric="rays/sec (fwd+bwd, 128 samples/ray)", value=v, unit="rays/s",
                n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=args.steps, warmup=args.warmup, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD, note="reference is pure Python (cannot travel to the GPU box): timed on its "
                                                    "torch-CPU restatement oracle/sparf_oracle.py, pinned bit-exact to "
                                                    "the reference by tests/golden"),
                cpu_baseline=dict(value=v, unit="rays/s", cores=cores, kind="port",
                                  sample="each step = the full 1023-ray x 128-sample batch, fwd + loss + bwd"),
                e2e=dict(value=v, unit="rays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as one CUDA graph (default), 0: eager")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5  # bounded CPU sample: ~1 s per step per 8 cores
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
