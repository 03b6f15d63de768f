"""tcgen05 kernels alone (through the C ABI probes): operands are rounded to fp16 exactly as the kernel does,
so the only difference left is fp32 accumulation order — gate 2e-5 relative to the output scale."""
import pytest
import torch

from tests.helpers import model_and_sd, rel_max

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from emotiongestures_b200 import TED
    from emotiongestures_b200.engine import Engine
    return Engine(TED, "cuda:0", precision="tc")


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 256, 256), (1000, 126, 126), (77, 1536, 256),
                                   (64, 256, 8704), (4096, 1024, 256), (129, 8, 64), (5, 64, 256),
                                   # tall problems take the weights-resident variant: BN 256; BN 128 where N % 256 != 0
                                   # (also with a K = 576 slice of 144 KB); N % 256 == 0 with K too long streams 128 x 256 tiles
                                   (40000, 512, 256), (76000, 126, 256), (40000, 384, 576), (40000, 256, 576)])
@pytest.mark.parametrize("epi", ["plain", "bias_relu", "addend_mod"])
def test_linear_tc(eng, m, n, k, epi):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g) / k ** 0.5
    bias = torch.randn(n, generator=g) if epi != "plain" else None
    addend = torch.randn(34, n, generator=g) if epi == "addend_mod" else None
    got = eng.debug_linear_tc(a, w, bias, addend, 34 if addend is not None else 0, relu=(epi == "bias_relu")).cpu()
    ref = a.half().double() @ w.half().double().t()
    if bias is not None:
        ref = ref + bias.double()
    if epi == "bias_relu":
        ref = ref.clamp_min(0)
    if addend is not None:
        ref = ref + addend.double()[torch.arange(m) % 34]
    assert rel_max(got, ref) <= 2e-5


@pytest.mark.parametrize("m,k", [(34, 512), (300, 1024), (5000, 512), (40000, 1024), (129, 256)])
@pytest.mark.parametrize("offset", [0.0, 50.0])
def test_linear_residual_layernorm_epilogue(eng, m, k, offset):
    """LayerNorm(A W^T + b + residual) computed in the GEMM epilogue (one thread owns the 256-wide row) against
    float64 on the fp16-rounded operands; `offset` puts the row mean 50 standard deviations from zero (the moments
    are taken about the row's first value, so nothing cancels)."""
    g = torch.Generator().manual_seed(m + k)
    a = torch.randn(m, k, generator=g)
    w = torch.randn(256, k, generator=g) / k ** 0.5
    bias = torch.randn(256, generator=g) * 0.1
    res = torch.randn(m, 256, generator=g) + offset
    ln_g, ln_b = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    got32, got16 = eng.debug_linear_ln_tc(a, w, bias, res, ln_g, ln_b)
    pre = a.half().double() @ w.half().double().t() + bias.double() + res.double()
    ref = torch.nn.functional.layer_norm(pre, (256,), ln_g.double(), ln_b.double(), eps=1e-6)
    assert rel_max(got32.cpu(), ref) <= (2e-5 if offset == 0 else 2e-4)
    assert torch.equal(got16.cpu(), got32.cpu().half())


@pytest.mark.parametrize("m,d_inner", [(34, 1024), (128, 1024), (129, 256), (1000, 512), (40000, 1024)])
def test_fused_feed_forward_block(eng, m, d_inner):
    """k_ffn_tc.cu: LayerNorm(x + W2 relu(W1 x + b1) + b2) in one kernel against float64 on the fp16-rounded operands;
    the hidden activations are rounded to fp16 between the two GEMMs exactly as the unfused chain stores them."""
    g = torch.Generator().manual_seed(m + d_inner)
    x = torch.randn(m, 256, generator=g)
    w1 = torch.randn(d_inner, 256, generator=g) / 16
    b1 = torch.randn(d_inner, generator=g) * 0.1
    w2 = torch.randn(256, d_inner, generator=g) / d_inner ** 0.5
    b2 = torch.randn(256, generator=g) * 0.1
    ln_g, ln_b = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    got32, got16 = eng.debug_ffn_tc(x, w1, b1, w2, b2, ln_g, ln_b)
    hid = (x.half().double() @ w1.half().double().t() + b1.double()).clamp_min(0).float().half().double()
    pre = hid @ w2.half().double().t() + b2.double() + x.double()
    ref = torch.nn.functional.layer_norm(pre, (256,), ln_g.double(), ln_b.double(), eps=1e-6)
    # the fp16 rounding of a hidden value can flip on an fp32-accumulation-order difference: one half-ulp of one
    # of d_inner terms, far below the 2e-3 budget but above pure re-association noise
    assert rel_max(got32.cpu(), ref) <= 1e-4
    assert torch.equal(got16.cpu(), got32.cpu().half())


@pytest.mark.parametrize("cin,cout,h,w,ks,stride,nchw", [
    (32, 32, 128, 70, 3, 1, False), (64, 64, 64, 35, 3, 1, False), (128, 128, 32, 18, 3, 1, False),
    (128, 34, 32, 18, 3, 1, True), (128, 60, 32, 31, 3, 1, True), (64, 64, 64, 62, 3, 1, False),
    (32, 64, 128, 70, 3, 2, False), (64, 128, 64, 35, 3, 2, False), (32, 64, 128, 70, 1, 2, False),
    (64, 128, 64, 35, 1, 2, False), (32, 32, 9, 5, 3, 1, False)])
def test_conv_tc(eng, cin, cout, h, w, ks, stride, nchw):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(cin + cout + h + w + ks + stride)
    b = 3
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5
    bias = torch.randn(cout, generator=g) if nchw else None
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    relu_first = not nchw
    got = eng.debug_conv_tc(x, wt, scale, shift, bias, stride, relu_first, nchw).cpu()
    ref = F.conv2d(x.half().double(), wt.half().double(), None if bias is None else bias.double(),
                   stride=stride, padding=ks // 2)
    if relu_first:
        ref = ref.clamp_min(0)
    ref = ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    assert got.shape == ref.shape
    assert not torch.isnan(got).any(), "unwritten output elements"
    # output is stored as fp16: half an ulp of the output scale on top of accumulation order
    assert rel_max(got, ref) <= 1.5e-3
    assert ((got - ref).abs() <= 1e-3 * ref.abs() + 2e-3).all()


@pytest.mark.parametrize("c,h,w", [(32, 128, 70), (64, 64, 35), (128, 32, 18), (64, 64, 62)])
def test_conv_tc_se_partial_sums(eng, c, h, w):
    """conv2's epilogue also emits per-tile channel sums (the SE squeeze); they must add up to the
    channel sums of the fp32 conv+BN output (not of its fp16 rounding)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(c + h)
    b = 5
    x = torch.randn(b, c, h, w, generator=g)
    wt = torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5
    scale = torch.rand(c, generator=g) + 0.5
    shift = torch.randn(c, generator=g) * 0.1
    got, sums = eng.debug_conv_tc(x, wt, scale, shift, se_sums=True)
    ref = F.conv2d(x.half().double(), wt.half().double(), padding=1) * scale.double().view(1, -1, 1, 1) \
        + shift.double().view(1, -1, 1, 1)
    assert rel_max(got.cpu(), ref) <= 1.5e-3
    mean_err = (sums.cpu().double() / (h * w) - ref.mean(dim=(2, 3))).abs().max().item()
    assert mean_err <= 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("b,l", [(3, 34), (7, 34), (1, 34), (4, 60), (5, 60), (2, 17), (130, 34)])
def test_attention_tc(eng, b, l):
    """Clip packing (3 x 34 / 2 x 60 rows per tile), ragged last group, block-diagonal masking."""
    g = torch.Generator().manual_seed(b * 100 + l)
    q, k, v = (torch.randn(b, 8, l, 64, generator=g) * s for s in (1.5, 1.5, 1.0))
    got = eng.debug_attention_tc(q, k, v).cpu()
    qd, kd, vd = q.half().double(), k.half().double(), v.half().double()
    ref = torch.softmax((qd / 8.0) @ kd.transpose(2, 3), dim=-1) @ vd
    assert not torch.isnan(got).any()
    assert rel_max(got, ref) <= 2e-3          # fp16 P and fp16 output
