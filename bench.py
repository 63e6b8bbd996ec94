#!/usr/bin/env python
"""bench.py -- Recall@K queries/s of the gallery-sharded retrieval scorer on B200 (+ the prompted-CLIP train pairs/s leg).

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
            top-k lists + recall counts back.  `e2e_resident`: queries in / lists out with the gallery resident
            (the reference's real use: the gallery features come from the GPU encoders, sprompt.py:456-509).
`roofline`: the dominant kernel (gemm_pair_kernel<MODE_TOPK>) against the measured bf16 tensor peak.
`cpu_baseline`: the reference's procedure (dense fp32 matmul on all host cores + np.argsort per row,
            sprompt.py:509,559-567) on a bounded sample of the same workload.
`parity`  : at every N: sampled oracle top-k, and the merged sharded lists == the unsharded single-GPU lists.
`train`   : the second BASELINE metric (prompted-CLIP train pairs/s, configs[2]) with its own roofline, e2e,
            cpu_baseline (the REAL reference on the host cores), gpu_eager (the real reference on the B200 through
            stock torch, fp16 and bf16 -- the kernel to beat) and, at N > 1, data-parallel == single-process parity.
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
CPU_SAMPLE_QUERIES = 128        # queries per CPU-baseline step (full gallery each): enough rows for a compute-bound host matmul
REAL_REFERENCE_MAX_GALLERY = 100_000      # up to here the reference arm runs the REAL SPrompts.itm_eval (dense [Q, N] scores fit)
TRAIN_GFLOP_PER_PAIR = 89.7     # 2 x linear + 3 x attention + patch embed, 77 text positions


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


# ------------------------------------------------------------------------------------------ reference arm (host cores)
def cpu_reference_step(q_host, g_host_f32, gt_rows, oracle):
    """One bounded-sample step of the reference's procedure (oracle port): dense scores + argsort ranks."""
    s = oracle.dense_scores(q_host, g_host_f32).numpy()
    ranks = oracle.reference_ranks_argsort(s, [[int(g)] for g in gt_rows])
    return s, ranks


def real_reference_step(q_host_f32, g_host_f32, gt_rows, ref_ns):
    """The REAL reference on the host: `score_t2i = (image_feats @ text_feats.t()).t()` (sprompt.py:509) then the unmodified
    `SPrompts.itm_eval` (sprompt.py:550-646).  itm_eval always walks both directions, so the image->text side gets a single row."""
    import types

    import numpy as np

    s_t2i = (g_host_f32 @ q_host_f32.t()).t().contiguous().numpy()          # [n_queries, n_gallery]
    s_i2t = np.ascontiguousarray(s_t2i[:1, :1])
    n = s_t2i.shape[0]
    return ref_ns.sprompt.SPrompts.itm_eval(types.SimpleNamespace(cur_id=0), s_i2t, s_t2i, {t: int(gt_rows[t]) for t in range(n)}, {0: [0]},
                                            [0], np.zeros(n, dtype=np.int64))


def reference_sample_queries(args) -> int:
    """Queries per reference-arm step: 128 (compute-bound host matmul), fewer when K + W is large so the whole run ends in minutes."""
    n = CPU_SAMPLE_QUERIES * 12 // max(12, args.steps + args.warmup)
    return max(8, min(args.queries, (max(32, n) // 8) * 8))


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
    nq = reference_sample_queries(args)
    kind, ref_ns = "port", None
    if args.gallery <= REAL_REFERENCE_MAX_GALLERY:
        from oracle import reference_loader as RL
        if RL.reference_available():
            ref_ns = RL.load_reference()
            kind = "reference"
    times = []
    for it in range(args.warmup + args.steps):
        lo = (it * nq) % max(1, args.queries - nq + 1)
        qs = q[lo:lo + nq].cpu()
        t0 = time.perf_counter()
        if ref_ns is not None:
            from oracle import reference_loader as RL
            with RL.in_reference_cwd():
                real_reference_step(qs.float(), g_host, gt[lo:lo + nq].tolist(), ref_ns)
        else:
            cpu_reference_step(qs, g_host, gt[lo:lo + nq].tolist(), O)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = nq / (ms / 1e3)
    how = ("the REAL reference: (image_feats @ text_feats.t()).t() + unmodified SPrompts.itm_eval (baseline/_ref)" if kind == "reference"
           else "oracle port of sprompt.py:509,559-567 (dense fp32 matmul + np.argsort per row; the real itm_eval needs the dense "
                f"[Q, N] matrix and is run for galleries <= {REAL_REFERENCE_MAX_GALLERY} rows)")
    sample = f"{nq} of {args.queries} queries per step against the full {args.gallery}-row gallery, {cores} host threads; {how}"
    line = {"impl": "reference", "metric": "recall_at_k_queries_per_sec", "value": val, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(1, args.gpus)),
            "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample},
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

    n_chunks = ops.sim_topk_chunks(nq, hi - lo)
    xbuf, xs, xi = R.exchange_buffer(n_chunks, nq, TOPK, dev)     # this rank's per-chunk lists, laid out for the one all-gather

    def step(record=False):
        """threshold pre-pass + scoring pass (top-k in the GEMM epilogue, chunks sharing their thresholds) -> [N > 1: ONE all-gather of the
        exchange buffer] -> ONE kernel: k-way merge over chunks and ranks + Recall@1/5/10 counters."""
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ops.sim_topk(q, shard, TOPK, lo, n_chunks, merge=False, out=(xs, xi))
        if record:
            e1.record()
            kern_events.append((e0, e1))
        sc, ix, counts = R.merge_recall(xbuf, ptr, idx, task, 1, group)
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

    def upload(on_chunk=None):
        """queries + the gallery shard host -> device in `n_parts` pieces spread over the copy streams; returns the per-chunk events."""
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
        return q_ev, evs

    def e2e_step():
        main = torch.cuda.current_stream()
        q_ev, evs = upload()
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

    def e2e_resident_step():
        """queries in (pinned host), lists + counts out; the gallery shard stays resident in HBM."""
        q_dev.copy_(q_host, non_blocking=True)
        ops.sim_topk(q_dev, shard, TOPK, lo, n_chunks, merge=False, out=(xs, xi))
        s_, i_, c_ = R.merge_recall(xbuf, ptr, idx, task, 1, group)
        out_ix.copy_(i_, non_blocking=True)
        out_counts.copy_(c_, non_blocking=True)
        return c_

    def timed(fn, n_steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        w0 = time.perf_counter()
        t0.record()
        for _ in range(n_steps):
            fn()
        t1.record()
        barrier()
        wall = (time.perf_counter() - w0) * 1e3
        ms = torch.tensor([max(t0.elapsed_time(t1), wall)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n_steps

    e2e_steps = max(1, min(args.steps, 5))
    sampler2 = None
    if rank == 0 and clocks is not None and (clocks.get("samples") or 0) < 3:     # short timed region (many GPUs): sample the e2e steps too
        sampler2 = ClockSampler(local)
        sampler2.start()
    e2e_ms = timed(e2e_step, e2e_steps, max(1, min(args.warmup, 2)))
    if sampler2 is not None:
        c2 = sampler2.stop()
        if (c2.get("samples") or 0) > (clocks.get("samples") or 0):
            clocks = c2
            clocks["sampled_over"] = "e2e timed steps (the device-resident region was shorter than 3 samples)"
    e2e_counts = out_counts.clone()
    h2d = q_host.numel() * 2 + g_host.numel() * 2
    d2h = out_ix.numel() * 4 + out_counts.numel() * 4
    # the host -> device ceiling of this box: the same bytes copied by every rank at once with nothing else running
    def upload_only():
        q_ev, evs = upload()
        main = torch.cuda.current_stream()
        main.wait_event(q_ev)
        for ev in evs:
            main.wait_event(ev)
    h2d_ms = timed(upload_only, 3, 1)
    res_ms = timed(e2e_resident_step, e2e_steps, 2)
    res_counts = out_counts.clone()

    # ---------------- parity (rank 0, every N) + CPU baseline (rank 0, N = 1 only for the timing)
    line_extra = {}
    if rank == 0:
        from oracle import lpi_oracle as O          # checker only (bench cpu_baseline leg)
        assert torch.equal(e2e_counts.cpu(), counts.cpu()), "e2e and device-resident passes disagree"
        assert torch.equal(res_counts.cpu(), counts.cpu()), "e2e_resident and device-resident passes disagree"
        parity = {}
        full = shard
        if world > 1:                                # the unsharded single-GPU result on the same rows: must equal the merged shard lists
            g_dev = None
            full, _, _ = S.make_gallery_shard(ng, 0, ng, nq, DIM, device=dev)
            sc1, ix1 = ops.sim_topk(q, full, TOPK, 0)
            parity["lists_equal_unsharded"] = bool(torch.equal(ix1, ix) and torch.equal(sc1, sc))
            parity["index_checksum"] = int(ix.to(torch.int64).sum())
            parity["index_checksum_unsharded"] = int(ix1.to(torch.int64).sum())
            assert parity["lists_equal_unsharded"], "merged sharded top-k lists differ from the unsharded run"
        else:
            parity["index_checksum"] = int(ix.to(torch.int64).sum())
        if not args.skip_cpu:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            g_f32 = (g_host if world == 1 else full.cpu()).float()
            # N = 1: a short warm step (page faults of the score buffer, thread pools), then ONE timed step of 128 queries (~23 s on 16 cores)
            plan = [16, CPU_SAMPLE_QUERIES] if world == 1 else [24]
            times, mism, near, qlo = [], 0, 0, 0
            for ns in plan:
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
                qlo += ns
            parity.update({"checked_queries": qlo, "topk_mismatch": int(mism), "near_tie_swaps": int(near)})
            if world == 1:
                cpu_s = times[-1]
                line_extra["cpu_baseline"] = {"value": ns / cpu_s, "unit": "queries/s", "cores": cores, "kind": "port",
                                              "sample": f"{ns} of {nq} queries per step against the full {ng}-row gallery: dense fp32 "
                                                        f"matmul on {cores} threads + np.argsort per row (oracle port of sprompt.py:509,"
                                                        f"559-567; the real itm_eval runs in the `sweep` leg at 100 k rows); one timed step after a 16-query warm step"}
            del g_f32
        line_extra["parity"] = parity
        c = counts.cpu().tolist()[0]
        line_extra["recall"] = {"r1": 100.0 * c[0] / c[3], "r5": 100.0 * c[1] / c[3], "r10": 100.0 * c[2] / c[3]}
        del full

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
                        "ms_per_step": e2e_ms, "steps": e2e_steps, "chunks": n_parts, "copy_streams": len(copy_streams),
                        "h2d_gbs_per_gpu": h2d / (e2e_ms * 1e-3) / 1e9,
                        "h2d_only_ms": h2d_ms, "h2d_only_gbs_per_gpu": h2d / (h2d_ms * 1e-3) / 1e9,
                        "note": "h2d_only = the same pinned host -> device copies by every rank at once with no scoring: the host-link "
                                "ceiling of this box for this step"},
                "e2e_resident": {"value": nq / (res_ms * 1e-3), "unit": "queries/s", "ms_per_step": res_ms, "h2d_bytes_per_step": q_host.numel() * 2,
                                 "d2h_bytes_per_step": d2h, "note": "queries in from pinned host memory, lists + counts out; gallery shard resident "
                                 "in HBM (the reference's use: gallery features come from the GPU encoders, sprompt.py:456-509)"},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"bound": "tensor", "kernel": "gemm_pair_kernel<MODE_TOPK> (cta_group::2, resident query tile)", "achieved": ach, "peak": peaks["tflops"],
                             "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": traffic,
                             "peak_source": peaks["source"] + " (sustained cuBLAS bf16)", "kernel_ms": kern_ms,
                             "flops_per_launch": flops,
                             "algorithmic_bytes_per_launch": (hi - lo) * DIM * 2 + nq * DIM * 2 + nq * TOPK * 8}}
        line.update(line_extra)
    del shard, g_host
    g_dev = None
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.skip_sweep:
        try:
            line["sweep"] = sweep_leg(dev, args)
        except Exception as e:
            line["sweep"] = {"error": repr(e)[:300]}
    train = None
    if not args.skip_train:
        try:
            train = train_leg(dev, world, rank, group, max(3, min(args.steps, 10)), 3, peaks, args)
        except Exception as e:                       # the secondary leg must never take the headline line down
            train = {"error": repr(e)[:300]}
    if train is not None and "error" not in train and not args.skip_train and args.train_batch == 64:
        # BASELINE.json configs[2] spans batch 64 to 512: the same step at 256 pairs per GPU (timing + roofline only)
        try:
            torch.cuda.empty_cache()
            t256 = train_leg(dev, world, rank, group, 5, 3, peaks, args, batch_per_gpu=256, light=True)
            train["batch_256_per_gpu"] = {k: t256[k] for k in ("value", "unit", "ms_per_step", "global_batch", "launch_mode", "algorithmic_tflops",
                                                                "executed_tflops", "roofline", "e2e") if k in t256}
        except Exception as e:
            train["batch_256_per_gpu"] = {"error": repr(e)[:300]}
    if rank == 0:
        line["train"] = train
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_leg(dev, args):
    """BASELINE.json configs[4] smaller galleries (100 k / 500 k / 1 M rows x 25 k queries, device-resident) and configs[1] (Flickr30K-shaped
    1 k images x 5 k captions, both directions, through the public feature-level API), with the REAL reference beside the 100 k case and
    the 1 k x 5 k case.  Informational keys; the headline stays the 5 M-row run."""
    import numpy as np
    import torch
    from lpi_b200 import ops, retrieval as R, synthetic as S

    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for n in (100_000, 500_000, 1_000_000):
        g, q, gt = S.make_gallery_shard(n, 0, n, args.queries, DIM, device=dev)
        for _ in range(3):
            ops.sim_topk(q, g, TOPK, 0)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(5):
            ops.sim_topk(q, g, TOPK, 0)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        out[str(n)] = {"queries_per_s": args.queries / (ms * 1e-3), "ms_per_step": ms, "tflops": 2.0 * args.queries * n * DIM / (ms * 1e-3) / 1e12}
        if n == 100_000 and not args.skip_cpu:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gallery", str(n), "--queries", str(args.queries),
                                "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                out[str(n)]["cpu_reference"] = {k: d["cpu_baseline"][k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception:
                out[str(n)]["cpu_reference"] = {"error": (r.stderr or r.stdout)[-200:]}
        del g, q
    # configs[1]: 1 k x 5 k, Recall@1/5/10 in both directions, fp32-exact scoring (6-term split) through itm_eval_features
    img, txt, img2txt, txt2img, cat_i, cat_t = S.make_retrieval_set()
    img_d, txt_d = img.to(dev), txt.to(dev)
    for _ in range(2):
        res = R.itm_eval_features(img_d, txt_d, txt2img, img2txt, cat_i, cat_t, 5, precision="fp32")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        res = R.itm_eval_features(img_d, txt_d, txt2img, img2txt, cat_i, cat_t, 5, precision="fp32")
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    fl = {"queries": 6000, "ms_per_eval": ms, "queries_per_s": 6000 / (ms * 1e-3),
          "note": "wall clock of the whole call incl. host-side ground-truth CSR building, 2 sim+top-k launches, counters, result dict"}
    if not args.skip_cpu:
        from oracle import lpi_oracle as O
        s = (img @ txt.t()).numpy()
        c0 = time.perf_counter()
        want = O.itm_eval(s, np.ascontiguousarray(s.T), txt2img, img2txt, cat_i, cat_t, 5)
        fl["cpu_port_ms"] = (time.perf_counter() - c0) * 1e3
        fl["recall_equal_oracle"] = bool(want == res)
    out["flickr_1k_x_5k"] = fl
    return out


# ------------------------------------------------------------------------------------------ train leg
def train_flops_per_pair(text_len: int = 77, last_block_rows: bool = True):
    """(algorithmic, executed) GFLOP per image-text pair of one training step: 2 x linear + 3 x attention (fwd + dgrad, no wgrad) + the
    patch embedding forward.  `algorithmic` is the reference's own work (77 text positions, every row of every
    block); `executed` counts what the towers actually run: the text positions up to the batch's last EOT, and in the LAST block of each
    tower only the one row per sample the head reads for everything after the qkv projection (engine.Tower.last_block_rows: per
    tower 18 w^2 (L - 1) linear and 4 w L (L - 1) attention FLOP less per pass)."""
    def tower(width, L, trimmed_last):
        lin, att = 24.0 * width * width * L * 12, 4.0 * L * L * width * 12
        if trimmed_last:
            lin -= 18.0 * width * width * (L - 1)
            att -= 4.0 * width * L * (L - 1)
        return lin, att
    out = []
    for lt, trimmed in ((77, False), (text_len, last_block_rows)):
        v_lin, v_att = tower(768, 213, trimmed)
        t_lin, t_att = tower(512, lt, trimmed)
        out.append((2 * (v_lin + t_lin) + 3 * (v_att + t_att) + 2.0 * 196 * 768 * 768) / 1e9)
    return out[0], out[1]


def _kernel_shares(step_fn, n_steps=2):
    """CUDA-event time per kernel class over `n_steps` eager single-stream steps (every lpi_b200.ops wrapper bracketed by an event pair)."""
    import torch
    from lpi_b200 import ops

    classes = {"gemm": "gemm", "gemm_tf32": "gemm", "attn_fwd": "attention", "attn_bwd": "attention", "layernorm_fwd": "layernorm",
               "layernorm_bwd": "layernorm", "im2col_patches": "front/head", "assemble_vision": "front/head", "assemble_text": "front/head",
               "assemble_vision_bwd": "front/head", "sum_prompt_rows": "front/head", "head_fwd": "front/head", "head_bwd": "front/head",
               "prompt_fwd": "prompt+loss+sgd", "prompt_bwd": "prompt+loss+sgd", "sgemm": "prompt+loss+sgd", "clip_loss_logits": "prompt+loss+sgd",
               "row_mean": "prompt+loss+sgd", "add_rowconst": "prompt+loss+sgd", "task_loss": "prompt+loss+sgd", "sgd_momentum_step": "prompt+loss+sgd"}
    pairs, saved = [], {}
    for name, cls in classes.items():
        fn = getattr(ops, name, None)
        if fn is None:
            continue
        saved[name] = fn

        def wrap(fn=fn, cls=cls):
            def inner(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*a, **k)
                e1.record()
                pairs.append((cls, e0, e1))
                return r
            return inner
        setattr(ops, name, wrap())
    try:
        step_fn()
        torch.cuda.synchronize()
        pairs.clear()
        for _ in range(n_steps):
            step_fn()
        torch.cuda.synchronize()
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
    tot = {}
    for cls, e0, e1 in pairs:
        tot[cls] = tot.get(cls, 0.0) + e0.elapsed_time(e1) / n_steps
    s = sum(tot.values()) or 1.0
    return {k: {"ms": round(v, 4), "share": round(v / s, 4)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}, s


def train_leg(dev, world, rank, group, steps, warmup, peaks, args, batch_per_gpu=64, light=False):
    """Secondary metric of BASELINE.json ('prompted-CLIP train pairs/sec', configs[2]): COCO-shaped LPI training step, batch 64 per
    GPU, data-parallel over the ranks (global InfoNCE over all-gathered features, all-reduced prompt gradient), SGD step included."""
    import torch
    import torch.distributed as dist
    from lpi_b200 import lpi_step, ops, synthetic as S
    from lpi_b200.engine import TextEngine, VisionEngine

    sd = S.make_clip_state_dict(0)
    vision, text = VisionEngine(sd, dev), TextEngine(sd, dev)
    fac = {k: v.to(dev) for k, v in S.make_prompt_factors(0).items()}
    fac0 = {k: v.clone() for k, v in fac.items()}
    opt = lpi_step.PromptSGD(fac, 0.05)
    from lpi_b200 import tokenizer as T

    images_host = S.make_images(batch_per_gpu, rank).pin_memory()
    # the same captions the reference legs get, through PromptLearner's template "X X ... X <caption>." (prompt_learner.py:128-132)
    tokens_host = T.tokenize([" ".join(["X"] * 16) + " " + c + "." for c in S.make_captions(batch_per_gpu, rank)])
    text_len = int(tokens_host.argmax(dim=-1).max()) + 1      # host-side, as the tokenizer provides it: positions after the last EOT are dead
    if world > 1:                                             # one bound for all ranks (a longer bound is still exact)
        tl = torch.tensor([text_len], device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        text_len = int(tl)
    tokens_pinned = tokens_host.pin_memory()
    images, tokens = images_host.to(dev), tokens_host.to(dev)

    # ---- parity of the data-parallel step: the same global batch in ONE process on rank 0 (N > 1 only), before any SGD step
    parity = None
    if world > 1 and not light:
        r_dp = lpi_step.train_step(vision, text, fac0, images, tokens, 1 / 0.07, group=group, text_len=text_len)
        if rank == 0:
            gi = torch.cat([S.make_images(batch_per_gpu, r) for r in range(world)]).to(dev)
            gtk = torch.cat([T.tokenize([" ".join(["X"] * 16) + " " + c + "." for c in S.make_captions(batch_per_gpu, r)]) for r in range(world)]).to(dev)
            r_1 = lpi_step.train_step(vision, text, fac0, gi, gtk, 1 / 0.07, text_len=text_len)
            rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
            parity = {"global_batch": batch_per_gpu * world,
                      "base_loss_rel_diff": abs(float(r_dp["losses"]["base_loss"]) - float(r_1["losses"]["base_loss"])) / abs(float(r_1["losses"]["base_loss"])),
                      "grad_rel_diff_max": max(rel(r_dp["grads"][k], r_1["grads"][k]) for k in lpi_step.FACTOR_NAMES),
                      "note": "data-parallel step (all-gathered features, all-reduced 5 284-float gradient) vs the same global batch in one "
                              "process on rank 0; differences are fp32 reduction order only"}
            assert parity["base_loss_rel_diff"] < 1e-5 and parity["grad_rel_diff_max"] < 1e-3, parity
            del gi, gtk, r_1
        del r_dp
        torch.cuda.empty_cache()

    # first-step losses on the untouched factors (the reference legs report theirs at the same point: same images, captions and weights)
    r_first = lpi_step.train_step(vision, text, fac0, images, tokens, 1 / 0.07, group=group, text_len=text_len)
    loss_first = {k: float(v) for k, v in r_first["losses"].items()}
    del r_first

    def eager_step():
        r = lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, group=group, text_len=text_len)
        opt.step(r["grads"])
        return r

    graphed = None
    step, mode = eager_step, "eager"
    if os.environ.get("LPI_TRAIN_GRAPH", "1") != "0":
        try:                                               # one cudaGraphLaunch per step instead of ~400 Python-issued launches
            graphed = lpi_step.GraphedTrainStep(vision, text, fac, opt, images, tokens, 1 / 0.07, group=group, text_len=text_len)
            step, mode = graphed.step, "cuda-graph"
        except Exception as e:                             # capture is an optimisation: fall back to the eager step
            print(f"[bench] CUDA-graph capture of the training step failed, running eagerly: {e!r}", file=sys.stderr)
            torch.cuda.synchronize()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        for _ in range(warmup):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        sync()
        wall = (time.perf_counter() - w0) * 1e3 / n
        ms = torch.tensor([e0.elapsed_time(e1) / n, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]), r

    n0 = ops.KERNEL_LAUNCHES
    ms, _, r = timed(step, steps)
    launches_per_step = (ops.KERNEL_LAUNCHES - n0) / (steps + warmup)

    # ---- e2e: this step's images + token ids come from pinned host memory, the loss is read back, every step
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e_step():
        if graphed is not None:
            # every step's batch crosses PCIe from pinned host memory: the copy of the NEXT step's batch is started before this step's
            # replay (GraphedTrainStep.prefetch, copy stream) and overlaps it; the step itself starts from the batch staged a step earlier
            if getattr(graphed, "_staged", None) is None:
                graphed.prefetch(images_host, tokens_pinned)
            out = graphed.step()
            graphed.prefetch(images_host, tokens_pinned)
        else:
            images.copy_(images_host, non_blocking=True)
            tokens.copy_(tokens_pinned, non_blocking=True)
            out = eager_step()
        loss_host.copy_(out["losses"]["base_loss"].view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()         # the caller looks at the loss (sprompt.py:313-316)
        return out

    _, e2e_wall, _ = timed(e2e_step, steps)
    gb = batch_per_gpu * world
    g_alg, g_exe = train_flops_per_pair(text_len, vision.tower.last_block_rows)
    out = {"metric": "prompted_clip_train_pairs_per_sec", "value": gb / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
           "global_batch": gb, "parallelism": f"dp{world}", "launch_mode": mode, "launches_per_step": launches_per_step,
           "algorithmic_tflops": gb * g_alg / (ms * 1e-3) / 1e3, "executed_tflops": gb * g_exe / (ms * 1e-3) / 1e3,
           "loss": float(r["losses"]["base_loss"]), "loss_first_step": loss_first, "loss_first_step_sum": sum(loss_first.values()),
           "precision": f"vision {vision.precision} / text fp16 operands, fp32 accumulate, fp32 residual stream / LayerNorm / heads / losses",
           "text_positions_executed": text_len, "text_positions_note": "77-token captions; the text tower runs on the positions up to the batch's last "
           "EOT (output-exact under the causal mask); algorithmic figures count the reference's full 77, executed figures the positions run",
           "last_block_rows": bool(vision.tower.last_block_rows), "last_block_note": "the last block of each tower computes its attention / out_proj / "
           "MLP only for the row the head reads (CLS, EOT): the other rows of that block reach neither the features nor any gradient; "
           "algorithmic figures count the reference's full block, executed figures what runs",
           "workload": "BASELINE.json configs[2]: ViT-B/16 + 12-layer text, "
           "224x224 synthetic images, 77-token captions, random init, fwd + 3 losses + dgrad to 5 284 prompt scalars + SGD",
           "e2e": {"value": gb / (e2e_wall * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_wall,
                   "h2d_bytes_per_step": images_host.numel() * 4 + tokens_pinned.numel() * 8, "d2h_bytes_per_step": 4,
                   "note": "per step: pinned host images + token ids -> device (one batch per step, copied on a copy stream while the previous step runs), graph replay, loss read back (host wall clock, max over ranks)"}}
    per_gpu_alg = batch_per_gpu * g_alg / (ms * 1e-3) / 1e3
    per_gpu_exe = batch_per_gpu * g_exe / (ms * 1e-3) / 1e3
    out["roofline"] = {"bound": "tensor", "achieved": per_gpu_alg, "executed": per_gpu_exe, "peak": peaks["tflops"], "unit": "TFLOP/s per GPU",
                       "frac": per_gpu_alg / peaks["tflops"], "frac_executed": per_gpu_exe / peaks["tflops"],
                       "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                       "gflop_per_pair": {"algorithmic": g_alg, "executed": g_exe}}
    if parity is not None:
        out["parity"] = parity
    if rank == 0 and world == 1 and not light:
        try:                                               # where the step's time goes: event pairs around every kernel wrapper, eager, one stream
            shares, total = _kernel_shares(lambda: (lpi_step.train_step(vision, text, fac, images, tokens, 1 / 0.07, text_len=text_len,
                                                                        overlap_towers=False), None)[1])
            out["roofline"]["kernel_shares"] = shares
            out["roofline"]["kernel_shares_note"] = (f"CUDA events around every kernel wrapper over 2 eager single-stream steps: {total:.2f} ms serialised "
                                                     f"vs {ms:.2f} ms for the two-stream graph replay")
        except Exception as e:
            out["roofline"]["kernel_shares"] = {"error": repr(e)[:200]}
    del vision, text, graphed
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.skip_cpu and not light:
        for leg in ("gpu-eager", "cpu-train-ref"):         # own processes: the CPU oracle patches torch.Tensor.cuda, the GPU leg must not see that
            try:
                rr = subprocess.run([sys.executable, os.path.abspath(__file__), "--leg", leg, "--train-batch", str(batch_per_gpu)],
                                    capture_output=True, text=True, timeout=900)
                res = json.loads(rr.stdout.strip().splitlines()[-1])
            except Exception as e:
                res = {"error": repr(e)[:200]}
            out["gpu_eager" if leg == "gpu-eager" else "cpu_baseline"] = res
    return out


# ---- the REAL reference (baseline/_ref copy of retrieval/, or /root/reference in the build container), timed in its own process
def _reference_train_setup(batch, cuda, dtype=None):
    import torch
    from lpi_b200 import synthetic as S
    from oracle import reference_loader as RL

    ns = RL.load_reference(cuda=cuda)
    sd = S.make_clip_state_dict(0)
    if cuda:
        def make_clip(_args):
            m = ns.clip_model.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval()
            m.load_state_dict(sd)
            if dtype == "fp16":
                ns.clip_model.convert_weights(m)            # what build_model does for every GPU run (model.py:394-415, 522)
            return m.cuda()                                 # bf16: fp32 weights under torch.autocast (the reference's LayerNorm subclass rejects bf16 weights)
        ns.slinet.load_clip_to_cpu = make_clip
        with RL.in_reference_cwd():
            args = RL.reference_args()
            args["device"] = [torch.device("cuda")]
            net = ns.slinet.SliNet(args).cuda()
    else:
        net = RL.build_reference_slinet(sd, None, numtask=0)
    with torch.no_grad():
        for k, v in S.make_prompt_factors(0).items():
            getattr(net.prompts[0], k).copy_(v)
    net.update_fc(0)
    net.train()
    for n, p in net.named_parameters():                     # freeze policy, sprompt.py:229-237
        p.requires_grad_("prompts.0." in n)
    opt = torch.optim.SGD(net.parameters(), momentum=0.9, lr=0.05, weight_decay=2e-4)       # sprompt.py:253
    images, captions = S.make_images(batch, 0), S.make_captions(batch, 0)
    if cuda:
        images = images.cuda()

    import contextlib

    def step(n=batch):
        # the loop body of train_function, sprompt.py:300-311
        with (torch.autocast("cuda", dtype=torch.bfloat16) if (cuda and dtype == "bf16") else contextlib.nullcontext()):
            img_f, txt_f, vp, tp = net(images[:n], captions[:n])
            with RL.in_reference_cwd():
                out = net.cal_loss(img_f, txt_f, vp, tp)
            loss = sum(l for l in out["loss"].values())
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)
    return step


def leg_cpu_train_ref(batch):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _reference_train_setup(batch, cuda=False)
    step(2)                                                 # warm-up on 2 pairs (thread pools, allocator)
    t0 = time.perf_counter()
    loss = step()
    dt = time.perf_counter() - t0
    print(json.dumps({"value": batch / dt, "unit": "pairs/s", "cores": cores, "kind": "reference", "ms_per_step": dt * 1e3, "loss": loss,
                      "sample": f"1 step at B = {batch} after a 2-pair warm-up: the REAL reference (unmodified retrieval/ package) SliNet.forward + "
                                f"cal_loss + backward + SGD.step (sprompt.py:300-311) in fp32 on {cores} host threads"}), flush=True)


def leg_gpu_eager(batch):
    import torch

    out = {"note": "the REAL reference modules on the same B200 through stock torch (cuBLAS / torch attention), weights as build_model "
                   "ships them to a GPU (fp16 via convert_weights) and in bf16; captions tokenised on the host every step as the reference does"}
    for dtype in ("fp16", "bf16"):
        # one dtype per process would be cleaner, but the reference module cache is per process: rebuild the network only
        try:
            step = _reference_train_setup(batch, cuda=True, dtype=dtype)
            for _ in range(3):
                loss = step()
            torch.cuda.synchronize()
            n = 5
            t0 = time.perf_counter()
            for _ in range(n):
                loss = step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            out[dtype] = {"value": batch / dt, "unit": "pairs/s", "ms_per_step": dt * 1e3, "loss": loss,
                          "algorithmic_tflops": batch * TRAIN_GFLOP_PER_PAIR / dt / 1e3}
            del step
            torch.cuda.empty_cache()
        except Exception as e:
            out[dtype] = {"error": repr(e)[:300]}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gallery", type=int, default=5_000_000)
    ap.add_argument("--queries", type=int, default=25_000)
    ap.add_argument("--e2e-chunks", type=int, default=16, help="host->device pipeline depth of the e2e leg (gallery shard copied in this many pieces)")
    ap.add_argument("--e2e-streams", type=int, default=1, help="copy streams the e2e leg spreads its host->device chunks over")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: no cpu_baseline / reference legs")
    ap.add_argument("--skip-train", action="store_true", help="skip the secondary train pairs/s leg")
    ap.add_argument("--skip-sweep", action="store_true", help="skip the smaller-gallery / Flickr-shaped informational leg")
    ap.add_argument("--leg", default=None, choices=["cpu-train-ref", "gpu-eager"], help="internal: one reference leg of the train metric in its own process")
    ap.add_argument("--train-batch", type=int, default=64)
    args = ap.parse_args()
    if args.leg == "cpu-train-ref":
        return leg_cpu_train_ref(args.train_batch)
    if args.leg == "gpu-eager":
        return leg_gpu_eager(args.train_batch)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
