"""-m gpu checks of the tcgen05/TMA GEMM with its fused epilogues against an fp32 torch matmul of the same
bf16 operands (floating-point kernel: tolerance = bf16 output rounding, written below)."""
import pytest
import torch

from lpi_b200 import ops

pytestmark = pytest.mark.gpu


def _case(M, N, K, epi, tile_n, seed=0, half=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).cuda().to(half)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda().to(half)
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    aux = torch.randn(M, N, generator=g).cuda().to(half)
    ref = a.float() @ w.float().t()
    kw, second = {}, None
    if epi == ops.EPI_BIAS_BF16:
        want, kw = ref + bias, dict(bias=bias)
    elif epi == ops.EPI_BIAS_GELU_BF16:
        z = ref + bias
        want = z * torch.sigmoid(1.702 * z)
        kw = dict(bias=bias, out2=torch.empty(M, N, device="cuda", dtype=half))
        second = (kw["out2"], z)
    elif epi == ops.EPI_BIAS_RESID_F32:
        want = resid + ref + bias
        kw = dict(bias=bias, resid=resid, out2=torch.empty(M, N, device="cuda", dtype=half))
        second = (kw["out2"], want)
    elif epi == ops.EPI_F32:
        want = ref
    elif epi == ops.EPI_BIAS_F32:
        want, kw = ref + bias, dict(bias=bias)
    elif epi == ops.EPI_ACC_F32:
        want, kw = resid + ref, dict(out=resid.clone())
    elif epi == ops.EPI_DGELU_BF16:
        zf = aux.float()
        s = torch.sigmoid(1.702 * zf)
        want, kw = ref * (s * (1 + 1.702 * zf * (1 - s))), dict(aux=aux)
    else:
        want = ref
    out = ops.gemm(a, w, epi, tile_n=tile_n, **kw)
    scale = max(1.0, want.abs().max().item())
    hr = 2 ** -8 if half == torch.bfloat16 else 2 ** -10               # output rounding of the 16-bit type (+ fast-sigmoid error)
    if half == torch.float16 and epi in (ops.EPI_BIAS_GELU_BF16, ops.EPI_DGELU_BF16):
        hr = 2 ** -9                                                    # fp16 towers evaluate (d)QuickGELU on packed halves: ~2 ulp(fp16)
    tol = (hr if out.dtype == half else 2e-5) * scale                   # 16-bit rounding / fp32 accumulation order
    assert (out.float() - want).abs().max().item() <= tol
    if second is not None:
        assert (second[0].float() - second[1]).abs().max().item() <= hr * max(1.0, second[1].abs().max().item())


@pytest.mark.parametrize("M,N,K", [(100, 128, 64), (300, 512, 512), (4928, 1536, 512), (13632, 2304, 768),
                                   (13632, 768, 3072), (1, 128, 64)])
@pytest.mark.parametrize("tile_n", [0, 128, 256, 512, 1192, 1128, 2256, 2192, 2128])
def test_gemm_shapes(M, N, K, tile_n):
    if tile_n and N % {128: 128, 256: 256, 512: 256, 1192: 192, 1128: 128, 2256: 256, 2192: 192, 2128: 128}[tile_n]:
        pytest.skip("N not a multiple of the tile")
    _case(M, N, K, ops.EPI_F32, tile_n)


@pytest.mark.parametrize("epi", range(8))
def test_gemm_epilogues(epi):
    _case(1000, 768, 768, epi, 0)
    _case(333, 512, 2048, epi, 128, seed=1)
    _case(1000, 768, 768, epi, 512, seed=2)          # CTA-pair kernel (cta_group::2), odd number of M tiles
    _case(13632, 768, 3072, epi, 512, seed=3)


@pytest.mark.parametrize("epi", range(8))
@pytest.mark.parametrize("tile_n", [1192, 1128, 2256, 2192, 2128])
def test_gemm_narrow_pair_tiles(epi, tile_n):
    """CTA-pair kernel with 256 x 192 / 256 x 128 cluster tiles (picked automatically when 256-wide tiles would leave the last wave
    mostly empty, e.g. N = 768 at B = 64) and the 4-CTA clusters (2xxx: two pairs on adjacent row blocks of one N tile, the B tile
    multicast between them): every epilogue, ragged M (odd numbers of row blocks and of block pairs), both 16-bit operand types."""
    _case(1000, 768, 768, epi, tile_n, seed=7)
    _case(13632, 768, 3072, epi, tile_n, seed=8)
    _case(2496, 1536, 512, epi, tile_n, seed=9, half=torch.float16)


@pytest.mark.parametrize("epi", range(8))
def test_gemm_f16_epilogues(epi):
    """fp16 operands (text tower): same epilogues, every 16-bit tensor is fp16, tolerance = fp16 output rounding."""
    _case(1000, 768, 768, epi, 0, seed=4, half=torch.float16)
    _case(4928, 512, 2048, epi, 0, seed=5, half=torch.float16)
    _case(4928, 1536, 512, epi, 512, seed=6, half=torch.float16)


def test_gemm_f16_rejects_narrow_n():
    """fp16 operands run on the CTA-pair kernel only (cluster tiles of 128 / 192 / 256 columns); the 1-CTA tiles are bf16 / TF32."""
    from lpi_b200._lib import LpiError

    a = torch.zeros(64, 64, device="cuda", dtype=torch.float16)
    w = torch.zeros(64, 64, device="cuda", dtype=torch.float16)
    with pytest.raises(LpiError):
        ops.gemm(a, w, ops.EPI_F32)
    w = torch.zeros(128, 64, device="cuda", dtype=torch.float16)
    with pytest.raises(LpiError):
        ops.gemm(a, w, ops.EPI_F32, tile_n=128)


def test_gemm_rejects_bad_shapes():
    from lpi_b200._lib import LpiError

    a = torch.zeros(8, 60, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(128, 60, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(LpiError):
        ops.gemm(a, w, ops.EPI_F32)


@pytest.mark.parametrize("M,N,K", [(77 * 3, 512, 512), (4928, 1536, 512), (1000, 512, 2048), (5, 128, 32)])
def test_gemm_tf32(M, N, K):
    """fp32 operands on the TF32 path: error bound = 2^-11 operand rounding (vs 2^-8 for bf16)."""
    g = torch.Generator().manual_seed(K)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    ref = a.double() @ w.double().t()
    out = ops.gemm_tf32(a, w, ops.EPI_F32)
    tol = 2 ** -9 * float(ref.abs().max())          # two operands rounded to 11 bits each
    assert (out.double() - ref).abs().max() < tol
    out = ops.gemm_tf32(a, w, ops.EPI_BIAS_RESID_F32, bias=bias, resid=resid)
    assert (out.double() - (ref + bias + resid)).abs().max() < tol
    z = torch.empty(M, N, device="cuda")
    out = ops.gemm_tf32(a, w, ops.EPI_BIAS_GELU_F32, bias=bias, out2=z)
    pre = ref + bias
    assert (z.double() - pre).abs().max() < tol and (out.double() - pre * torch.sigmoid(1.702 * pre)).abs().max() < 2 * tol
    out = ops.gemm_tf32(a, w, ops.EPI_DGELU_F32, aux=z)
    s = torch.sigmoid(1.702 * z.double())
    assert (out.double() - ref * (s * (1 + 1.702 * z.double() * (1 - s)))).abs().max() < 2 * tol
    outb = ops.gemm_tf32(a, w, ops.EPI_BIAS_BF16, bias=bias)
    assert outb.dtype == torch.bfloat16 and (outb.double() - pre).abs().max() < 2 ** -8 * float(pre.abs().max())


@pytest.mark.parametrize("epi", [ops.EPI_BIAS_BF16, ops.EPI_BIAS_GELU_BF16, ops.EPI_BIAS_RESID_F32, ops.EPI_DGELU_BF16, ops.EPI_BF16])
def test_gemm_one_row_per_sample(epi):
    """M = batch size: the GEMMs of a tower's last block, which runs on the one row per sample the head reads (engine.Tower,
    last_block_rows): a single, mostly empty 256-row tile per N tile."""
    for M in (64, 5, 300):
        _case(M, 768, 768, epi, 0, seed=11, half=torch.float16)
        _case(M, 3072, 768, epi, 0, seed=12, half=torch.float16)
        _case(M, 768, 3072, epi, 0, seed=13, half=torch.float16)


def test_gemm_dynamic_tile_scheduler_clc():
    """LPI_GEMM_CLC=1 (opt-in): one cluster per tile in the grid, the resident CTA pairs take over the tiles of the clusters that
    have not been launched (clusterlaunchcontrol.try_cancel).  Same results as the static schedule; the switch is read once per
    process, hence the child process."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import importlib.util, torch\n"
            "spec = importlib.util.spec_from_file_location('gemm_cases', 'tests/test_gpu_gemm.py')\n"
            "t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)\n"
            "for epi in range(8):\n"
            "    t._case(13632, 2304, 768, epi, 0, seed=21, half=torch.float16)\n"
            "    t._case(13632, 768, 3072, epi, 0, seed=22)\n"
            "    t._case(54528, 768, 768, epi, 1192, seed=23, half=torch.float16)\n"
            "t._case(2496, 1536, 512, 0, 0, seed=24, half=torch.float16)\n"
            "torch.cuda.synchronize(); print('clc-ok')\n")
    env = dict(os.environ, LPI_GEMM_CLC="1", PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clc-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("B,L,H,K,half", [(64, 213, 12, 768, torch.float16), (3, 39, 8, 512, torch.float16), (5, 77, 8, 512, torch.bfloat16),
                                          (2, 197, 12, 768, torch.bfloat16), (1, 1, 2, 64, torch.float16)])
def test_gemm_do_delta(B, L, H, K, half):
    """lpi_gemm_do_delta: the out_proj dgrad (EPI_BF16) whose epilogue also leaves delta[b, h, l] = sum_d dO * O, the row sums of the
    attention backward (model.py:172,183-185 under autograd), on the caller-zeroed buffer; every cluster tile width."""
    g = torch.Generator().manual_seed(B * 7 + L)
    M, N = B * L, H * 64
    a = torch.randn(M, K, generator=g).cuda().to(half)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).cuda().to(half)
    o = torch.randn(M, N, generator=g).cuda().to(half)
    delta = torch.zeros(B * H * L, device="cuda")
    out = ops.gemm_do_delta(a, w, o, delta, L)
    ref = a.float() @ w.float().t()
    hr = 2 ** -8 if half == torch.bfloat16 else 2 ** -10
    assert out.dtype == half and (out.float() - ref).abs().max() <= hr * max(1.0, ref.abs().max().item())
    assert torch.equal(out, ops.gemm(a, w, ops.EPI_BF16))                      # the GEMM result itself is unchanged
    want = (ref * o.float()).view(B, L, H, 64).sum(-1).permute(0, 2, 1).reshape(-1)
    assert (delta - want).abs().max() <= 2e-5 * max(1.0, want.abs().max().item()) + 1e-4
    delta2 = torch.zeros_like(delta)
    ops.gemm_do_delta(a, w, o, delta2, L)
    assert torch.equal(delta, delta2)                                          # two addends per element: order-independent
