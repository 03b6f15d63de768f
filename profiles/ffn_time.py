"""Time the fused feed-forward kernel alone (attribution builds read EGX_FFN_CLUSTER / EGX_FFN_DEBUG):
    python profiles/ffn_time.py [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emotiongestures_b200 import TED
from emotiongestures_b200.engine import Engine, _ptr

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 34
eng = Engine(TED, "cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, 256, generator=g, device="cuda")
w1 = torch.randn(1024, 256, generator=g, device="cuda") / 16
w2 = torch.randn(256, 1024, generator=g, device="cuda") / 32
b1, b2 = torch.randn(1024, device="cuda") * 0.1, torch.randn(256, device="cuda") * 0.1
lg, lb = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda") * 0.1
x16, w1h, w2h = x.half(), w1.half(), w2.half()
o32, o16 = torch.empty(M, 256, device="cuda"), torch.empty(M, 256, device="cuda", dtype=torch.float16)
import ctypes as C
lib = eng.lib
# the debug probe converts its operands on every call; time it around the conversions by differencing two row counts is
# not needed: the conversions are ~0.1 ms against several ms of repeated calls below
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    eng.debug_ffn_tc(x, w1, b1, w2, b2, lg, lb)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    eng.debug_ffn_tc(x, w1, b1, w2, b2, lg, lb)
e1.record()
torch.cuda.synchronize()
print("rows %d: %.1f us per call (incl. operand conversion), cluster=%s debug=%s" % (
    M, e0.elapsed_time(e1) * 100, os.environ.get("EGX_FFN_CLUSTER"), os.environ.get("EGX_FFN_DEBUG")))
