import sys, os
sys.path.insert(0, os.getcwd())
import torch
from tests.test_gpu_parity import _engine
from emotiongestures_b200 import LOGMEL_LOG_IN
eng, _ = _engine("ted", 0, "fp32")
b, n, cols = 4096, 36267, 70
a = (0.1 * torch.randn(b, n, device="cuda")).clamp_(-1, 1)
for gt in (False, True, False):
    for _ in range(3): eng.logmel(a, LOGMEL_LOG_IN, True, n_cols=cols, _global_tile=gt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.logmel(a, LOGMEL_LOG_IN, True, n_cols=cols, _global_tile=gt)
    e1.record(); torch.cuda.synchronize()
    print("global_tile" if gt else "smem_tile", e0.elapsed_time(e1) / 20)
