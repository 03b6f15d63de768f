"""One generator-inference step (TED shape) between cudaProfilerStart/Stop, for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/step_once.py [clips]
Warm-up steps run before the profiled one; nothing here is a benchmark number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emotiongestures_b200 import LOGMEL_LOG_IN, TED, Transformer, randomize_norm_stats_

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = Transformer.from_config(TED).eval()
randomize_norm_stats_(gen, 1)
gen = gen.to(dev)
eng = gen.engine("tc")
g = torch.Generator(device=dev).manual_seed(1000)
audio = (0.1 * torch.randn(B, TED.n_audio, generator=g, device=dev)).clamp_(-1, 1)
prior = torch.randn(B, TED.prior_frames, TED.pose_dim, generator=g, device=dev)


def step():
    spec = eng.logmel(audio, LOGMEL_LOG_IN, True)
    return eng.generator_forward(spec, prior, None)[0]


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
