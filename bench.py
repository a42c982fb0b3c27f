#!/usr/bin/env python
"""Benchmark of the NDJIR per-ray rendering hot path (BASELINE.json: train rays/s forward+backward at default.yaml;
grid-query GB/s vs HBM peak).

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one process per GPU, torchrun for N>1)
  python bench.py --impl reference [...]                     the reference path restated on the host CPU (oracle/)

One "step" = one loss.forward() + loss.backward() of the reference (python/train.py:135-140) over one batch of
B=4 views x R=512 rays of a synthetic DTU-shaped scene: sample placement (4 SDF-guided rounds), 512^3 x 4 voxel
feature query, geometric / material / light MLPs, NeuS compositing, shading, all losses, and the full backward
including the double-backward through the SDF normal, plus zeroing of every gradient buffer (the 2 GiB grid gradient
included).  Multi-GPU: rays are sharded (weak scaling: every rank renders its own 2048 rays), parameters replicated,
gradients all-reduced with NCCL inside the step.

JSON line keys follow the driver's contract; `roofline` describes the dominant kernel (the MLP product kernel),
`grid_query` the voxel-grid gather kernel against the measured HBM peak, `cpu_baseline` the oracle timed on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def profiled_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/r1_traffic.json), or {}."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def grid_label(conf):
    v = conf.geometric_network.voxel
    if v.type == "voxel":
        return f"voxel {v.grid_size}^3 x {v.feature_size}"
    if v.type == "triplaneline":
        return f"triplane 3 x {v.grid_size}^2 x {v.feature_size} + triline 3 x {v.grid_size} x {v.feature_size}"
    return "none (MLP-only SDF)"


def workload(conf, name="default"):
    r, tr = conf.renderer, conf.train
    return dict(workload=f"{name}.yaml train step (fwd+bwd), synthetic DTU-shaped rays",
                views=tr.batch_size, rays_per_view=tr.n_rays,
                fg_samples=r.n_samples0 + r.n_upsamples * r.n_samples1, bg_samples=r.n_bg_samples,
                light_dirs=2 * r.n_thetas * 2 * r.n_thetas, grid=grid_label(conf),
                l2="working set per step (~15 GB of activations + grid and grid-gradient tables) exceeds the 126 MB L2")


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the reference path; nnabla is not installable
# here) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_run(conf, steps, warmup, rays_per_step, views=1):
    import torch
    from ndjir_b200 import scene
    from oracle import cpu_render as CR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = scene.init_params(conf, seed=313)
    gen = torch.Generator().manual_seed(313)
    grids = {k: (P["grid"].get(k) if P["grid"].get(k) is not None else (torch.randn(shp, generator=gen) * 1e-3).numpy())
             for k, shp in scene.grid_shapes(conf).items()}
    model = CR.Model(conf, P, dtype=torch.float32, grids=grids or None)
    B, R = views, rays_per_step
    times = []
    for s in range(warmup + steps):
        camloc, raydir, color_gt = scene.make_batch(conf, step=s, B=B, R=R)
        rnd = scene.make_randoms(conf, B, R, step=s)
        t0 = time.perf_counter()
        CR.train_step(model, camloc, raydir, color_gt, 0.0, rnd)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return dict(value=B * R / (ms / 1e3), unit="rays/s", cores=cores, kind="port",
                sample=f"{steps} timed step(s) after {warmup} warm-up of {B} view(s) x {R} rays (full per-ray config and "
                       f"network widths, grid: {grid_label(conf)}, fwd+bwd incl. the dense grid gradient, whose fixed "
                       f"per-step cost is amortised over {B * R} rays here and over "
                       f"{conf.train.batch_size * conf.train.n_rays} in the full batch) with torch CPU fp32 on {cores} "
                       f"threads; oracle/cpu_render.py",
                ms_per_step=ms, warmup=warmup, steps=steps)


def run_reference(args, conf):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rays = args.ref_rays
    ref_warmup = min(args.warmup, 1)          # one warm-up step is enough on the CPU; the line reports what was run
    cb = cpu_reference_run(conf, max(1, args.steps), ref_warmup, rays)
    cfg = workload(conf, args.config)
    cfg["sample"] = cb["sample"]
    if not args.no_ref_full:
        # the SAME batch as the GPU arm (all views x all rays), once: the bounded sample above amortises the dense grid
        # gradient over fewer rays, this step does not
        tr = conf.train
        full = cpu_reference_run(conf, 1, 0, tr.n_rays, views=tr.batch_size)
        cfg["full_batch_step"] = {"value": full["value"], "unit": "rays/s", "ms_per_step": full["ms_per_step"],
                                  "views": tr.batch_size, "rays_per_view": tr.n_rays, "steps": 1, "warmup": 0}
    line = {"impl": "reference", "metric": "train_rays_per_sec_fwd_bwd", "value": cb["value"], "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": ref_warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------
def grid_query_roofline(eng, pk, torch):
    """voxel-grid gather at the micro-benchmark shape (2^24 uniform points, the training grid): algorithmic
    156 B/point (12 query + 16 output + 8 corners x 16 B) / CUDA-event time, vs the measured HBM copy peak."""
    from ndjir_b200 import _lib
    conf = eng.conf
    v = conf.geometric_network.voxel
    if v.type != "voxel":
        return None
    G, D = v.grid_size, v.feature_size
    n = 1 << 24
    gen = torch.Generator(device="cuda").manual_seed(412)
    q = torch.rand((n, 3), device="cuda", generator=gen) * 2 - 1
    out = torch.empty((n, D), device="cuda")
    F = eng.params.grid["voxel"]
    st = torch.cuda.current_stream().cuda_stream

    def once():
        _lib.call("ndjir_voxel_query_on_voxel", n, out.data_ptr(), q.data_ptr(), F.data_ptr(), [G, G, G], D, [-1.0] * 3,
                  [1.0] * 3, 0, st)
    def timed():
        for _ in range(3):
            once()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            once()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms = timed()                                     # default dispatch: brick-ordered sweep (csrc/voxel_binned.cu)
    _lib.call("ndjir_set_option", "voxel_binned", 0)
    ms_direct = timed()                              # one pass over the points in caller order (csrc/voxel.cu)
    _lib.call("ndjir_set_option", "voxel_binned", -1)
    bytes_pt = 12 + 4 * D + 8 * 4 * D
    ach = bytes_pt * n / (ms * 1e-3) / 1e9
    prof = profiled_traffic()
    tr = prof.get("voxel_binned_query_dram_bytes_per_call")
    tr_direct = prof.get("gather4_kernel_dram_bytes_per_launch")
    return {"kernel": "voxel gather (query_on_voxel): bin_count + bin_scan + bin_place + brick-ordered gather sweep",
            "points": n, "bytes_per_point": bytes_pt, "ms": ms,
            "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
            "traffic": tr, "dram_gbs_from_traffic": (tr / (ms * 1e-3) / 1e9 if tr else None),
            "direct_kernel": {"ms": ms_direct, "achieved": bytes_pt * n / (ms_direct * 1e-3) / 1e9,
                              "frac": bytes_pt * n / (ms_direct * 1e-3) / 1e9 / pk["hbm"], "traffic": tr_direct},
            "note": "uniform random 16-byte cells: DRAM fetches 128-byte lines, so the direct kernel moves ~3.6x the "
                    "algorithmic bytes at ~0.9 of the HBM copy peak; sorting the points by 16 MB table brick first "
                    "lets L2 serve the reuse inside a brick (DRAM traffic per call 9.7 -> 4.1 GB incl. the sort)",
            "peak_source": pk["source"], "l2": "2 GiB table and 2^24 random points: far larger than L2"}


def optimizer_roofline(eng, pk, torch):
    """Fused optimizer pass (csrc/optimizer.cu) over every parameter of the config - for
    default.yaml the 2 GiB voxel grid dominates: weight decay + non-finite check + Adam + gradient zeroing in one pass,
    32 algorithmic bytes per parameter (read w, g, m, v; write w, m, v, g)."""
    from ndjir_b200.solver import Solvers
    sv = Solvers(eng.conf, eng)
    sv.set_parameters()
    sv.update_learning_rate(100)
    n = sum(w.numel() for _, w, _ in sv._groups)
    for _ in range(3):
        sv.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        sv.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ach = 32.0 * n / (ms * 1e-3) / 1e9
    del sv
    return {"kernel": "optimizer::adam_kernel (+ non-finite scan of the MLP gradients)", "parameters": n, "ms": ms,
            "bytes_per_parameter": 32, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
            "frac": ach / pk["hbm"], "note": "not part of the timed fwd+bwd step (BASELINE metric); the reference runs "
            "zero_grad, weight_decay, check_inf_or_nan_grad and Adam as separate passes (~25 GB per iteration)"}


def run_ours(args, conf):
    import torch
    import torch.distributed as dist
    from ndjir_b200 import scene
    from ndjir_b200.engine import Engine, LOSS_NAMES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    eng = Engine(conf, device=f"cuda:{local}", world_size=world, process_group=pg)
    P = scene.init_params(conf, seed=313)
    eng.params.load_reference(P)
    eng.params.init_grid_on_device(scene.grid_shapes(conf), std=1e-3, seed=313)
    tr = conf.train
    B, R = tr.batch_size, tr.n_rays
    nsteps = args.warmup + args.steps

    # synthetic batches; host copies are pinned for the e2e arm.  weak scaling: every rank renders its own full batch;
    # strong scaling: ONE batch per step, sharded along the ray axis (parallel.shard_rays: rank k gets rays
    # [k R / world, (k + 1) R / world) of every view)
    strong = args.scaling == "strong" and world > 1
    host = []
    for s in range(nsteps):
        seed = s if strong else s * world + rank
        camloc, raydir, color_gt = scene.make_batch(conf, step=seed)
        rnd = scene.make_randoms(conf, B, R, step=seed)
        if strong:
            n = R // world
            assert n * world == R, "rays per view must divide by the number of GPUs"
            raydir, color_gt = raydir[:, rank * n:(rank + 1) * n], color_gt[:, rank * n:(rank + 1) * n]
            rnd = {k: v[:, rank * n:(rank + 1) * n] for k, v in rnd.items()}
        item = {"camloc": camloc, "raydir": raydir, "color_gt": color_gt, **rnd}
        host.append({k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in item.items()})
    h2d_bytes = sum(v.numel() * 4 for v in host[0].values())
    resident = [{k: v.cuda(non_blocking=True) for k, v in item.items()} for item in host]
    torch.cuda.synchronize()

    use_graph = not args.no_graph      # CUDA-graph replay of the step (Engine.train_step_graphed), NCCL exchanges included
    rnd_keys = [k for k in host[0] if k not in ("camloc", "raydir", "color_gt")]

    def run_step(item):
        if use_graph:
            return eng.train_step_graphed(item["camloc"], item["raydir"], item["color_gt"], {k: item[k] for k in rnd_keys},
                                          cos_anneal_ratio=0.0)
        return eng.train_step(item["camloc"], item["raydir"], item["color_gt"], item, cos_anneal_ratio=0.0)

    def step_resident(item):
        return run_step(item)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- value: inputs resident in HBM ----
    for s in range(args.warmup):
        step_resident(resident[s])
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.warmup, nsteps):
        losses = step_resident(resident[s])
    e1.record()
    barrier()
    ms_total = maxreduce(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    rays_per_step = B * R if strong else B * R * world
    value = rays_per_step / (ms_step * 1e-3)
    loss_host = losses.detach().cpu().numpy()

    # ---- e2e: host buffers in, loss out, copies inside the timed region ----
    dev_bufs = {k: torch.empty_like(v, device="cuda") for k, v in host[0].items()}
    loss_pinned = torch.empty(len(LOSS_NAMES), dtype=torch.float32).pin_memory()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for s in range(args.warmup, nsteps):
        if use_graph:       # the graphed step copies the pinned host tensors straight into its static device inputs
            l_ = run_step(host[s])
        else:
            for k, v in host[s].items():
                dev_bufs[k].copy_(v, non_blocking=True)
            l_ = run_step(dev_bufs)
        loss_pinned.copy_(l_, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the caller reads the loss every step (train.py:141-146)
    e3.record()
    barrier()
    ms_e2e = maxreduce(e2.elapsed_time(e3)) / args.steps
    e2e_value = rays_per_step / (ms_e2e * 1e-3)

    # ---- instrumented step: launches and product-kernel time (CUDA events around every ndjir_gemm) ----
    eng.profile = True
    eng.prof_events, eng.n_launches = [], 0
    it = resident[-1]
    eng.train_step(it["camloc"], it["raydir"], it["color_gt"], it, cos_anneal_ratio=0.0)     # eager: events per launch
    torch.cuda.synchronize()
    eng.profile = False
    gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in eng.prof_events)
    gemm_flops = sum(f for _, _, f, _ in eng.prof_events)
    buckets = {}
    for a_, b_, f_, tag in eng.prof_events:
        d = buckets.setdefault(tag, [0, 0.0, 0.0])
        d[0] += 1; d[1] += a_.elapsed_time(b_); d[2] += f_
    if os.environ.get("NDJIR_BENCH_DUMP"):
        with open(os.environ["NDJIR_BENCH_DUMP"], "w") as f:
            for k, v in sorted(buckets.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{k:<34s} n={v[0]:<4d} ms={v[1]:8.3f} tflops={v[2] / max(v[1], 1e-9) / 1e9:7.1f}\n")
    top = sorted(buckets.items(), key=lambda kv: -kv[1][1])[:12]
    breakdown = [{"shape": k, "launches": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / max(v[1], 1e-9) / 1e9, 1)}
                 for k, v in top]
    launches = eng.n_launches
    pk = peaks()
    roof = None
    if gemm_ms > 0:
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
        roof = {"kernel": "ndjir::gemmh::gemm_h_kernel (tcgen05 kind::f16 products on TMA-fed split-fp16 operands, three "
                          "tensor-core products per algorithmic product, fused epilogues; <= 8-wide shapes on the "
                          "memory-bound corner kernels)",
                "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_sustained"],
                "traffic": profiled_traffic().get("gemm_h_kernel_fwd_hidden_layer_dram_bytes_per_launch"),
                "traffic_note": "DRAM bytes of one 262144x256x256 forward product (algorithmic 537 MB: 4 B per "
                                "element in, 4 B out)",
                "ceiling_3_products": pk["tf_sustained"] / 3.0, "frac_of_3_product_ceiling": ach / (pk["tf_sustained"] / 3.0),
                "peak_source": pk["source"] + " (bf16 sustained)",
                "launches_per_step": len(eng.prof_events), "ms_per_step_in_kernel": gemm_ms,
                "share_of_step": gemm_ms / ms_step, "breakdown": breakdown,
                "how": "CUDA events around every product launch of one instrumented step right after the timed region; "
                       "achieved = algorithmic 2*M*N*K of all launches / summed duration"}
    gq = grid_query_roofline(eng, pk, torch) if rank == 0 else None
    opt = optimizer_roofline(eng, pk, torch) if rank == 0 else None

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(conf, 2, 1, args.ref_rays)
        cfg = workload(conf, args.config)
        cfg["launch"] = ("one CUDA-graph replay per step (Engine.train_step_graphed)" if use_graph
                         else "eager: every kernel enqueued from the host")
        if world > 1:
            cfg["grid_gradient_exchange"] = ("sparse: all-gather of the per-sample scatter inputs, replicated scatter"
                                             if eng._sparse_grid() else "dense all-reduce of the table gradient")
        cfg["parallelism"] = (f"ray-sharded x{world} ({'one batch split along the rays' if strong else 'one full batch per rank'}), "
                              f"replicated parameters, NCCL gradient all-reduce") if world > 1 else "1 GPU"
        if strong:
            cfg["rays_per_view_per_rank"] = R // world
        line = {"metric": "train_rays_per_sec_fwd_bwd", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": len(LOSS_NAMES) * 4, "ms_per_step": ms_e2e},
                "gpu_launches": launches * args.steps,
                "roofline": roof, "grid_query": gq, "optimizer_step": opt,
                "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None),
                "loss": {k: float(v) for k, v in zip(LOSS_NAMES, loss_host)}}
        emit(line)
    if world > 1:
        # The captured graphs hold NCCL work: tearing the communicator down underneath them can block at interpreter
        # exit.  Everything has been printed; leave without the collective teardown once every rank is done.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if _REAL_STDOUT is not None:
            _REAL_STDOUT.flush()
        os._exit(0)


def main():
    # Libraries (NCCL's version banner, torchrun) write to stdout; the contract is ONE JSON line there.  Everything else
    # goes to stderr: fd 1 is pointed at fd 2 for the run and the line is written to the saved descriptor at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="default")
    ap.add_argument("--ref-rays", type=int, default=256, help="rays per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-full", action="store_true", help="reference arm: skip the one full-batch step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every rank renders its own full batch; strong = one batch sharded along the rays")
    ap.add_argument("--rays", type=int, default=0, help="override rays per view (debugging)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from the host instead of replaying "
                    "the captured CUDA graph of the step")
    ap.add_argument("--grid", type=int, default=0, help="override voxel grid size (debugging)")
    args = ap.parse_args()
    from ndjir_b200.config import make_conf
    over = {}
    if args.rays:
        over["train"] = {"n_rays": args.rays}
    if args.grid:
        over["geometric_network"] = {"voxel": {"grid_size": args.grid}}
    conf = make_conf(args.config, **over)
    if args.impl == "reference":
        run_reference(args, conf)
    else:
        run_ours(args, conf)


if __name__ == "__main__":
    main()
