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
                                   (64, 256, 8704), (4096, 1024, 256), (129, 8, 64), (5, 64, 256)])
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
