#!/usr/bin/env python
"""bench.py -- Recall@K queries/s of the gallery-sharded retrieval scorer on B200.

Workload (BASELINE.json configs[4], the configuration the 1/2/4/8-GPU metric is quoted on; SURVEY.md
section 8(d) config 5): 25 000 text queries x 5 000 000 image embeddings, d = 512, bf16 rows, top-10 per
query + Recall@1/5/10.  One "step" = one full pass of the scorer over the whole gallery: similarity GEMM with
the top-k kept in the epilogue (score matrix never written), k-way merge, and -- for N > 1 -- one NCCL
all-gather of the per-shard candidates (gallery rows sharded contiguously over the ranks; total work fixed,
so "scaling" is "strong").

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (torchrun launches N > 1)
  python bench.py --impl reference ...                   # the reference's CPU procedure on the host cores

`value`   : queries/s with inputs resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`     : same metric through the public API with PINNED HOST buffers: every step copies the queries and the
            gallery shard host->device (chunked, overlapped with the scoring of earlier chunks) and reads the
            top-k lists + recall counts back.
`roofline`: the dominant kernel (gemm_tn_kernel<MODE_TOPK>) against the measured bf16 tensor peak.
`cpu_baseline`: oracle port of the reference's procedure (dense fp32 matmul on all host cores + np.argsort per
            row, sprompt.py:509,559-567) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 512
TOPK = 10
CPU_SAMPLE_QUERIES = 8          # queries per CPU-baseline step (full gallery each)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.lines = []
        self.proc = None
        self.sel = str(device_index)
        try:
            import torch
            u = getattr(torch.cuda.get_device_properties(device_index), "uuid", None)
            if u is not None:
                self.sel = "GPU-" + str(u)
        except Exception:
            pass

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.sel, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_step(q_host, g_host_f32, gt_rows, oracle):
    """One bounded-sample step of the reference's procedure: dense scores + argsort ranks."""
    s = oracle.dense_scores(q_host, g_host_f32).numpy()
    ranks = oracle.reference_ranks_argsort(s, [[int(g)] for g in gt_rows])
    return s, ranks


def run_reference(args):
    import torch
    from lpi_b200 import synthetic as S
    from oracle import lpi_oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    shard, q, gt = S.make_gallery_shard(args.gallery, 0, args.gallery, args.queries, DIM, device=dev)
    g_host = shard.cpu().float()
    del shard
    nq = CPU_SAMPLE_QUERIES
    times = []
    for it in range(args.warmup + args.steps):
        lo = (it * nq) % max(1, args.queries - nq)
        qs = q[lo:lo + nq].cpu()
        t0 = time.perf_counter()
        cpu_reference_step(qs, g_host, gt[lo:lo + nq].tolist(), O)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = nq / (ms / 1e3)
    sample = f"{nq} of {args.queries} queries per step against the full {args.gallery}-row gallery"
    line = {"impl": "reference", "metric": "recall_at_k_queries_per_sec", "value": val, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(1, args.gpus)),
            "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"large-gallery Recall@K sweep: {args.queries} text queries x {args.gallery} image embeddings, d={DIM}, "
                        f"bf16 rows, top-{TOPK} + Recall@1/5/10 (BASELINE.json configs[4])",
            "queries": args.queries, "gallery": args.gallery, "dim": DIM, "topk": TOPK,
            "parallelism": f"gallery-sharded x{world}", "l2": "inputs larger than L2 (gallery shard >= 640 MB vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from lpi_b200 import ops, retrieval as R, synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    peaks = _peaks()
    nq, ng = args.queries, args.gallery
    lo, hi = R.shard_bounds(ng, world, rank, align=256)
    shard, q, gt = S.make_gallery_shard(ng, lo, hi, nq, DIM, device=dev)
    ptr, idx = R.gt_csr([[int(g)] for g in gt.tolist()])
    ptr, idx = ptr.to(dev), idx.to(dev)
    task = torch.zeros(nq, dtype=torch.int32, device=dev)

    kern_events = []

    def step(record=False):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ps, pi = ops.sim_topk(q, shard, TOPK, lo, merge=False)
        if record:
            e1.record()
            kern_events.append((e0, e1))
        sc, ix = ops.topk_merge(ps, pi) if ps.shape[0] > 1 else (ps[0], pi[0])
        if world > 1:
            sc, ix = R.merge_across_ranks(sc, ix, group)
        counts = ops.recall_counts(ix, ptr, idx, task, 1)
        return sc, ix, counts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ops.KERNEL_LAUNCHES
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        sc, ix, counts = step(record=True)
    t1.record()
    barrier()
    launches = ops.KERNEL_LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["sampled_over"] = "device-resident timed steps"
    ms_total = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    kern_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kern_events) / len(kern_events)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total) / args.steps
    kern_ms = float(kern_ms)

    # ---------------- e2e: pinned host buffers in, results out, every step
    q_host = q.cpu().pin_memory()
    g_host = shard.cpu().pin_memory()
    n_parts = max(1, min(args.e2e_chunks, (hi - lo) // 65536))
    bounds = [R.shard_bounds(hi - lo, n_parts, c, align=256) for c in range(n_parts)]
    g_dev = torch.empty_like(shard)
    q_dev = torch.empty_like(q)
    copy_streams = [torch.cuda.Stream() for _ in range(max(1, args.e2e_streams))]
    out_ix = torch.empty(nq, TOPK, dtype=torch.int32).pin_memory()
    out_counts = torch.empty(1, 4, dtype=torch.int32).pin_memory()

    def e2e_step():
        main = torch.cuda.current_stream()
        evs = []
        for cs in copy_streams:
            cs.wait_stream(main)
        with torch.cuda.stream(copy_streams[0]):
            q_dev.copy_(q_host, non_blocking=True)
            q_ev = torch.cuda.Event()
            q_ev.record(copy_streams[0])
        for c, (a, b) in enumerate(bounds):                      # chunk c rides copy stream c mod n: several DMA queues keep the link busy
            cs = copy_streams[c % len(copy_streams)]
            with torch.cuda.stream(cs):
                g_dev[a:b].copy_(g_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                evs.append(ev)
        main.wait_event(q_ev)
        lists_s, lists_i, thr = [], [], None
        for (a, b), ev in zip(bounds, evs):
            main.wait_event(ev)
            # a row of chunk c only matters in the merged list if it beats the best k-th score seen so far: one pre-pass, in chunk 0
            s_, i_ = ops.sim_topk(q_dev, g_dev[a:b], TOPK, lo + a, init_thr=thr)
            kth = s_[:, TOPK - 1]
            thr = kth.contiguous() if thr is None else torch.maximum(thr, kth)      # short lists end in -inf: the running bound stays
            lists_s.append(s_); lists_i.append(i_)
        if len(lists_s) > 1:
            s_, i_ = ops.topk_merge(torch.stack(lists_s), torch.stack(lists_i))
        if world > 1:
            s_, i_ = R.merge_across_ranks(s_, i_, group)
        c_ = ops.recall_counts(i_, ptr, idx, task, 1)
        out_ix.copy_(i_, non_blocking=True)
        out_counts.copy_(c_, non_blocking=True)
        return c_

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    sampler2 = None
    if rank == 0 and clocks is not None and (clocks.get("samples") or 0) < 3:     # short timed region (many GPUs): sample the e2e steps too
        sampler2 = ClockSampler(local)
        sampler2.start()
    w0 = time.perf_counter()
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    barrier()
    wall = (time.perf_counter() - w0) * 1e3
    e2e_ms = torch.tensor([max(t0.elapsed_time(t1), wall)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms) / e2e_steps
    if sampler2 is not None:
        c2 = sampler2.stop()
        if (c2.get("samples") or 0) > (clocks.get("samples") or 0):
            clocks = c2
            clocks["sampled_over"] = "e2e timed steps (the device-resident region was shorter than 3 samples)"
    e2e_counts = out_counts.clone()
    h2d = q_host.numel() * 2 + g_host.numel() * 2
    d2h = out_ix.numel() * 4 + out_counts.numel() * 4

    # ---------------- parity spot-check + CPU baseline (rank 0, N = 1 only for the timing)
    line_extra = {}
    if rank == 0:
        from oracle import lpi_oracle as O          # checker only (bench cpu_baseline leg)
        assert torch.equal(e2e_counts.cpu(), counts.cpu()), "e2e and device-resident passes disagree"
        if world == 1 and not args.skip_cpu:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            g_f32 = g_host.float()
            ns = CPU_SAMPLE_QUERIES
            times, mism, near = [], 0, 0
            for it in range(3):
                qlo = it * ns
                qs = q_host[qlo:qlo + ns]
                c0 = time.perf_counter()
                s, ranks = cpu_reference_step(qs, g_f32, gt[qlo:qlo + ns].tolist(), O)
                times.append(time.perf_counter() - c0)
                _, want = O.topk_lowest_index(s, TOPK)
                got = ix[qlo:qlo + ns].cpu().numpy()
                for r in range(ns):
                    verdict = O.audit_topk(got[r], want[r], qs[r], g_f32)
                    near += verdict == "near"
                    mism += verdict == "bad"
            cpu_s = sum(times[1:]) / len(times[1:])
            line_extra["cpu_baseline"] = {"value": ns / cpu_s, "unit": "queries/s", "cores": cores, "kind": "port",
                                          "sample": f"{ns} of {nq} queries per step against the full {ng}-row gallery: dense fp32 "
                                                    f"matmul on {cores} threads + np.argsort per row (sprompt.py:509,559-567); mean of 2 steps"}
            line_extra["parity"] = {"checked_queries": 3 * ns, "topk_mismatch": mism, "near_tie_swaps": near}
        c = counts.cpu().tolist()[0]
        line_extra["recall"] = {"r1": 100.0 * c[0] / c[3], "r5": 100.0 * c[1] / c[3], "r10": 100.0 * c[2] / c[3]}

    if rank == 0:
        flops = 2.0 * nq * (hi - lo) * DIM
        ach = flops / (kern_ms * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "scorer_traffic.json")
        if os.path.isfile(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get(str(world), {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": "recall_at_k_queries_per_sec", "value": nq / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms, "steps": e2e_steps, "chunks": n_parts, "copy_streams": len(copy_streams)},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"bound": "tensor", "kernel": "gemm_pair_kernel<MODE_TOPK> (cta_group::2, resident query tile)", "achieved": ach, "peak": peaks["tflops"],
                             "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": traffic,
                             "peak_source": peaks["source"] + " (sustained cuBLAS bf16)", "kernel_ms": kern_ms,
                             "flops_per_launch": flops,
                             "algorithmic_bytes_per_launch": (hi - lo) * DIM * 2 + nq * DIM * 2 + nq * TOPK * 8}}
        line.update(line_extra)
    train = None
    if not args.skip_train:
        del shard, g_dev, g_host                     # free the gallery before building the towers
        torch.cuda.empty_cache()
        try:
            train = train_leg(dev, world, rank, group, max(3, min(args.steps, 10)), 3)
        except Exception as e:                       # the secondary leg must never take the headline line down
            train = {"error": repr(e)[:300]}
    if rank == 0:
        line["train"] = train
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_leg(dev, world, rank, group, steps, warmup, batch_per_gpu=64):
    """Secondary metric of BASELINE.json ('prompted-CLIP train pairs/sec', configs[2]): COCO-shaped LPI training step, batch 64 per
    GPU, data-parallel over the ranks (global InfoNCE over all-gathered features, all-reduced prompt gradient), SGD step included."""
    import torch
    import torch.distributed as dist
    from lpi_b200 import lpi_step, ops, synthetic as S
    from lpi_b200.engine import TextEngine, VisionEngine

    sd = S.make_clip_state_dict(0)
    vision, text = VisionEngine(sd, dev), TextEngine(sd, dev)
    fac = {k: v.to(dev) for k, v in S.make_prompt_factors(0).items()}
    opt = lpi_step.PromptSGD(fac, 0.05)
    images = S.make_images(batch_per_gpu, rank).to(dev)
    tokens_host = S.make_tokens(batch_per_gpu, rank)
    text_len = int(tokens_host.argmax(dim=-1).max()) + 1      # host-side, as the tokenizer provides it: positions after the last EOT are dead
    tokens = tokens_host.to(dev)

    def eager_step():
        r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, group=group, text_len=text_len)
        opt.step(r["grads"])
        return r

    step, mode = eager_step, "eager"
    if os.environ.get("LPI_TRAIN_GRAPH", "1") != "0":
        try:                                               # one cudaGraphLaunch per step instead of ~400 Python-issued launches
            graphed = lpi_step.GraphedTrainStep(vision, text, fac, opt, images, tokens, 1 / 0.07, group=group, text_len=text_len)
            step, mode = graphed.step, "cuda-graph"
        except Exception as e:                             # capture is an optimisation: fall back to the eager step
            print(f"[bench] CUDA-graph capture of the training step failed, running eagerly: {e!r}", file=sys.stderr)
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = ops.KERNEL_LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    gb = batch_per_gpu * world
    return {"metric": "prompted_clip_train_pairs_per_sec", "value": gb / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
            "global_batch": gb, "parallelism": f"dp{world}", "launch_mode": mode, "launches_per_step": (ops.KERNEL_LAUNCHES - n0) / steps,
            "algorithmic_tflops": gb * 89.7e9 / (ms * 1e-3) / 1e12, "loss": float(r["losses"]["base_loss"]),
            "precision": "vision bf16 / text fp16 operands, fp32 accumulate",
            "text_positions_executed": text_len, "text_positions_note": "77-token captions; the text tower runs on the positions up to the batch's last "
            "EOT (output-exact under the causal mask); algorithmic_tflops counts the reference's full 77",
            "workload": "BASELINE.json configs[2]: ViT-B/16 + 12-layer text, "
            "224x224 synthetic images, 77-token captions, random init, fwd + 3 losses + dgrad to 5 284 prompt scalars + SGD"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gallery", type=int, default=5_000_000)
    ap.add_argument("--queries", type=int, default=25_000)
    ap.add_argument("--e2e-chunks", type=int, default=8, help="host->device pipeline depth of the e2e leg (gallery shard copied in this many pieces)")
    ap.add_argument("--e2e-streams", type=int, default=1, help="copy streams the e2e leg spreads its host->device chunks over")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: no cpu_baseline / parity leg")
    ap.add_argument("--skip-train", action="store_true", help="skip the secondary train pairs/s leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
