"""GPU bring-up ladder: each rung runs in its own process under a timeout (a trap or hang in one rung
must not take the others down).  Usage on the box:  python tools/gpu_ladder.py [rung ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rung_simt():
    import torch
    from lpi_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 512, generator=g).cuda()
    y = ops.l2_normalize(x)
    ref = x / x.norm(dim=-1, keepdim=True)
    print("l2norm maxerr", (y - ref).abs().max().item())
    s = torch.randn(50, 4000, generator=g).cuda()
    s[:, 100] = s[:, 7]          # exact ties
    v, i = ops.topk_rows(s, 10)
    rv, ri = torch.sort(s, dim=1, descending=True, stable=True)
    print("topk_rows idx equal", bool((i.long() == ri[:, :10]).all()), "val equal", bool((v == rv[:, :10]).all()))
    sp = ops.split_bf16(x, 6, 0).float().view(300, 6, 512)
    rec = sp[:, 0] + sp[:, 2] + sp[:, 5]
    print("split recon maxerr", (rec - x).abs().max().item())


def _gemm_case(M, N, K, epi, tile_n, seed=0, verbose=True):
    import torch
    from lpi_b200 import ops
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(M, K, generator=g)).cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    aux = torch.randn(M, N, generator=g).cuda().bfloat16()
    ref = a.float() @ w.float().t()
    kw = {}
    if epi == ops.EPI_BIAS_BF16:
        want = ref + bias
        kw = dict(bias=bias)
    elif epi == ops.EPI_BIAS_GELU_BF16:
        z = ref + bias
        want = z * torch.sigmoid(1.702 * z)
        kw = dict(bias=bias, out2=torch.empty(M, N, device="cuda", dtype=torch.bfloat16))
    elif epi == ops.EPI_BIAS_RESID_F32:
        want = resid + ref + bias
        kw = dict(bias=bias, resid=resid, out2=torch.empty(M, N, device="cuda", dtype=torch.bfloat16))
    elif epi == ops.EPI_F32:
        want = ref
    elif epi == ops.EPI_BIAS_F32:
        want = ref + bias
        kw = dict(bias=bias)
    elif epi == ops.EPI_ACC_F32:
        want = resid + ref
        kw = dict(out=resid.clone())
    elif epi == ops.EPI_DGELU_BF16:
        zf = aux.float()
        s = torch.sigmoid(1.702 * zf)
        want = ref * (s * (1 + 1.702 * zf * (1 - s)))
        kw = dict(aux=aux)
    elif epi == ops.EPI_BF16:
        want = ref
    out = ops.gemm(a, w, epi, tile_n=tile_n, **kw)
    torch.cuda.synchronize()
    err = (out.float() - want).abs()
    tol = 2e-2 * want.abs().max().item() if out.dtype == torch.bfloat16 else 1e-3 * max(1.0, want.abs().max().item())
    ok = bool(err.max().item() <= tol)
    if verbose or not ok:
        print(f"gemm M={M} N={N} K={K} epi={epi} bn={tile_n}: maxerr={err.max().item():.3e} tol={tol:.2e} ok={ok}")
    if not ok:
        bad = (err > tol)
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows", rows[:16].tolist(), "... n=", len(rows), " bad cols", cols[:16].tolist(), "... n=", len(cols))
        print("   out[0,:8]", out[0, :8].float().tolist())
        print("   ref[0,:8]", want[0, :8].tolist())
        print("   out[1,:8]", out[1 % M, :8].float().tolist())
        print("   ref[1,:8]", want[1 % M, :8].tolist())
    if epi == ops.EPI_BIAS_GELU_BF16:
        e2 = (kw["out2"].float() - (ref + bias)).abs().max().item()
        print("   preact out2 maxerr", e2)
    if epi == ops.EPI_BIAS_RESID_F32:
        e2 = (kw["out2"].float() - want).abs().max().item()
        print("   bf16 shadow maxerr", e2)
    return ok


def rung_gemm_min():
    from lpi_b200 import ops
    _gemm_case(128, 128, 64, ops.EPI_F32, 128)
    _gemm_case(128, 256, 64, ops.EPI_F32, 256)
    _gemm_case(128, 128, 128, ops.EPI_F32, 128)
    _gemm_case(128, 256, 512, ops.EPI_F32, 256)


def rung_gemm_shapes():
    from lpi_b200 import ops
    ok = True
    for (M, N, K) in [(100, 128, 64), (300, 512, 512), (4928, 1536, 512), (13632, 2304, 768), (13632, 768, 3072),
                      (2000, 3072, 768)]:
        for bn in (128, 256, 0):
            if bn and N % bn:
                continue
            ok &= _gemm_case(M, N, K, ops.EPI_F32, bn)
    for epi in range(8):
        ok &= _gemm_case(1000, 768, 768, epi, 0)
        ok &= _gemm_case(333, 512, 2048, epi, 128)
    print("ALL_OK" if ok else "SOME_FAILED")


def _topk_ref(q, g, k):
    import torch
    s = q.float() @ g.float().t()
    v, i = torch.sort(s, dim=1, descending=True, stable=True)
    return v[:, :k], i[:, :k], s


def rung_scorer():
    import torch
    from lpi_b200 import ops
    gen = torch.Generator().manual_seed(1)
    for (nq, ng, dim, chunks) in [(128, 256, 64, 1), (100, 1000, 512, 1), (300, 5000, 512, 3), (1000, 20000, 512, 0),
                                  (257, 70001, 3072, 0)]:
        q = torch.randn(nq, dim, generator=gen).cuda().bfloat16()
        g = torch.randn(ng, dim, generator=gen).cuda().bfloat16()
        g[5] = g[3]            # exact ties -> lowest index must win
        v, i = ops.sim_topk(q, g, 10, 0, chunks)
        torch.cuda.synchronize()
        rv, ri, s = _topk_ref(q, g, 10)
        same = (i.long() == ri)
        # a mismatch is only legitimate when the fp32 scores are within accumulation-order noise
        gap = (v - rv).abs().max().item()
        print(f"scorer nq={nq} ng={ng} dim={dim} chunks={chunks}: idx match {same.float().mean().item():.6f} "
              f"score maxdiff {gap:.3e}")
        if not same.all():
            bad = (~same).any(1).nonzero().flatten()[:4]
            for b in bad.tolist():
                print("   q", b, "got", i[b].tolist(), "want", ri[b].tolist())
                print("     got", v[b].tolist())
                print("    want", rv[b].tolist())


def rung_bench_gemm():
    import torch
    from lpi_b200 import ops
    for (M, N, K) in [(13632, 2304, 768), (13632, 768, 768), (13632, 3072, 768), (13632, 768, 3072), (4928, 1536, 512),
                      (4928, 512, 2048), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        for bn in (128, 256, 512):
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            for _ in range(3):
                ops.gemm(a, w, ops.EPI_BF16, out=out, tile_n=bn)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10):
                ops.gemm(a, w, ops.EPI_BF16, out=out, tile_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"gemm {M}x{N}x{K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
        out = a @ w.t()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            out = a @ w.t()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"cublas {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")


def rung_bench_scorer():
    import torch
    from lpi_b200 import ops
    for (nq, ng) in [(25000, 100000), (25000, 1000000), (25000, 5000000)]:
        q = torch.randn(nq, 512, device="cuda").bfloat16()
        g = torch.randn(ng, 512, device="cuda").bfloat16()
        for _ in range(2):
            ops.sim_topk(q, g, 10)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            ops.sim_topk(q, g, 10)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"scorer {nq}x{ng}: {ms:.2f} ms  {2*nq*ng*512/ms/1e9:.1f} TFLOP/s  chunks={ops.sim_topk_chunks(nq, ng)}")


RUNGS = {k[5:]: v for k, v in list(globals().items()) if k.startswith("rung_")}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        RUNGS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(RUNGS)
    for n in names:
        t = time.time()
        print(f"===== rung {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=300,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-6000:])
            print(f"===== rung {n} exit={r.returncode} {time.time()-t:.1f}s", flush=True)
        except subprocess.TimeoutExpired as e:
            print((e.stdout or "")[-3000:] if isinstance(e.stdout, str) else e.stdout)
            print(f"===== rung {n} TIMEOUT", flush=True)
