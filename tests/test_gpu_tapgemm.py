"""GPU parity of the tcgen05 tap-GEMM kernel (lr_tapgemm; rows a5 / f4 and the input GEMM of a13):
  * plain GEMM (store mode 3) against torch fp32 on the bf16-rounded operands;
  * the position-map CNN plan launch by launch against oracle/tapgemm.py (the CPU statement of the kernel): both
    accumulate bf16 products in fp32 and round the stored activation to bf16, so they differ by summation order only
    (one bf16 ulp = 2^-8 relative where a rounding boundary is crossed, accumulated over the layers);
  * the whole CNN at the reference's resolution (256x256) against `prnet.ResFcn256` in fp32 — the restatement of
    src/models/face/prnet.py:211-280 — within the bf16 bar (2e-2 of MaxPos)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N", [(1000, 192, 48), (19200, 1728, 1536), (77, 64, 16), (4096, 520, 512)])
def test_plain_gemm_against_fp32(native_lib, cuda, M, K, N):
    from lipreading_b200 import functional as LF
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    out = LF.tap_linear(a, w, bias)
    ref = a.float() @ w.float().t() + bias
    assert out.shape == (M, N) and out.dtype == torch.float32
    assert float((out - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))     # fp32 accumulation order only


@pytest.mark.parametrize("layout", ["packed", "fused", "plain", "streamed"])
def test_cnn_plan_launch_by_launch_against_cpu_statement(native_lib, cuda, layout):
    from oracle import tapgemm as OT
    from test_prnet_plan import compile_with_layout, randomized_net
    net = randomized_net(3)
    B, R = 3, 64
    x = torch.rand(B, R, R, 3, generator=torch.Generator().manual_seed(5))
    cpu = compile_with_layout(layout, net, B, R, "cpu")
    dev = compile_with_layout(layout, net, B, R, cuda)
    want = OT.run_plan(cpu, x)
    got = dev.run(x.to(cuda))
    torch.cuda.synchronize()
    for (name, vc), (_, vd) in zip(cpu.marks, dev.marks):
        c, d = vc.t.float(), vd.t.float().cpu()
        assert float((c - d).abs().max()) <= 2e-2 * float(c.abs().max()), name          # borders and padded channels too
        assert float(((c - d).abs() > 2 ** -6 * c.abs().max()).float().mean()) < 1e-3, name
    assert float((got.cpu() - want).abs().max()) <= 1e-2 * R * 1.1


def test_cnn_at_256_against_fp32_module(native_lib, cuda):
    from lipreading_b200 import prnet as P
    from test_prnet_plan import randomized_net
    net = randomized_net(7)
    pred = P.PosPrediction(device=cuda)                      # default engine on CUDA: tcgen05
    assert pred.engine == "tcgen05"
    pred.network.load_state_dict(net.state_dict())
    pred._configure()
    x = torch.rand(2, 256, 256, 3, generator=torch.Generator().manual_seed(11)).to(cuda)
    n0 = native_lib.lr_launch_count()
    got = pred.predict_batch(x)
    assert native_lib.lr_launch_count() - n0 == 54           # image packing + 53 tap GEMMs, nothing else
    with torch.no_grad():
        ref = net.to(cuda)(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1) * pred.MaxPos
    assert got.shape == (2, 256, 256, 3) and got.dtype == torch.float32
    assert float((got - ref).abs().max()) <= 2e-2 * pred.MaxPos
    # a second batch size compiles its own plan; numpy in -> numpy out like the reference's predict()
    one = pred.predict(x[0].cpu().numpy())
    assert float(abs(one - got[0].cpu().numpy()).max()) <= 1e-3 * pred.MaxPos


def test_recurrent_layer_gemms_on_tapgemm_match_library(native_lib, cuda):
    """functional.GEMM_TCGEN05: x @ W_ih^T + b_ih, dX, dW_ih, dW_hh of the recurrent layer (better_model.py:47-49,74) on
    lr_tapgemm give what the library GEMMs give on the same bf16 operands (fp32 accumulation order only)."""
    from lipreading_b200 import functional as LF
    g = torch.Generator().manual_seed(9)
    B, T, I, H = 32, 20, 192, 128
    x = torch.randn(B, T, I, generator=g).to(cuda).to(torch.bfloat16)
    lens = torch.full((B,), T, dtype=torch.int32, device=cuda)
    ws = []
    for _ in range(2):
        ws += [(torch.randn(3 * H, I, generator=g) / I ** 0.5).to(cuda), (torch.randn(3 * H, H, generator=g) / H ** 0.5).to(cuda),
               torch.randn(3 * H, generator=g).to(cuda) * 0.1, torch.randn(3 * H, generator=g).to(cuda) * 0.1]
    saved = LF.GEMM_DTYPE, LF.GEMM_TCGEN05
    out = {}
    try:
        for tc in (False, True):
            LF.GEMM_DTYPE, LF.GEMM_TCGEN05 = torch.bfloat16, tc
            xs = x.clone().requires_grad_(True)
            wl = [w.clone().requires_grad_(True) for w in ws]
            n0 = native_lib.lr_launch_count()
            hidden, h_n = LF.rnn_layer(xs, lens, "GRU", wl)
            (hidden.square().sum() + h_n.sum()).backward()
            out[tc] = [hidden.detach(), xs.grad.float()] + [w.grad for w in wl]
            out[(tc, "launches")] = native_lib.lr_launch_count() - n0
    finally:
        LF.GEMM_DTYPE, LF.GEMM_TCGEN05 = saved
    assert out[(True, "launches")] == out[(False, "launches")] + 5          # gi, dX, dW_ih, 2 x dW_hh
    for a, b in zip(out[False], out[True]):
        assert float((a - b).abs().max()) <= 1e-2 * max(1e-3, float(a.abs().max()))
