"""GPU, world_size 2 over NCCL (gpurun --gpus 2): the sharded path end to end.  N-GPU poses must equal the 1-GPU
poses bit for bit (per-clip results do not depend on the batch a clip sits in), the pose gather handles ragged shards,
and the FGD [n | sum | gram] all-reduce gives every rank the single-process statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    import torch.distributed as dist
    from emotiongestures_b200 import LOGMEL_LOG_IN, TED, fgd
    from emotiongestures_b200.engine import Engine
    from emotiongestures_b200.sharding import PeerGather, all_gather_poses, shard_bounds
    from oracle import synth
    from tests.helpers import model_and_sd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        _, sd = model_and_sd("ted", 0)
        eng = Engine(TED, dev, precision="tc")
        eng.load_state_dict(sd)
        audio = torch.from_numpy(synth.synth_audio(n_clips, TED.n_audio, seed=21)).to(dev)     # the global batch
        prior = torch.from_numpy(synth.synth_prior(n_clips, TED.prior_frames, TED.pose_dim, 21)).to(dev)
        lo, hi = shard_bounds(n_clips, rank, world)
        local = eng.generator_forward(eng.logmel(audio[lo:hi], LOGMEL_LOG_IN, True), prior[lo:hi])[0]
        gathered = all_gather_poses(local, n_clips)
        whole = eng.generator_forward(eng.logmel(audio, LOGMEL_LOG_IN, True), prior)[0]      # 1-GPU result on this rank
        # FGD statistics of a feature stand-in (the poses' first 128 coordinates of frame 0): shard + all-reduce
        feats = gathered[:, 0, :64].contiguous()
        shift = feats[:4].double().mean(0)
        acc = fgd.new_accumulator(64, dev)
        eng.fgd_accumulate(feats[lo:hi], acc, shift)
        fgd.all_reduce_stats(acc)
        mu, sigma = fgd.finalize_stats(acc, 64, shift)
        # the copy-engine gather (equal shards): same bytes as the NCCL all-gather, several rounds over both slots
        peer_ok = True
        if n_clips % world == 0:
            pg = PeerGather(local.numel(), local.dtype, dev)
            comm = torch.cuda.Stream(dev)
            for it in range(4):
                shard = (local + float(it)).contiguous()
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(dev))
                comm.wait_event(ready)
                full = pg.gather(shard, it & 1, comm)
                comm.synchronize()
                peer_ok &= bool(torch.equal(full.view(n_clips, *local.shape[1:]), whole + float(it)))
        q.put((rank, bool(torch.equal(gathered, whole)) and peer_ok, whole.cpu().numpy(), mu, sigma))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [16, 11])
def test_world2_nccl_sharded_poses_equal_single_gpu_poses(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(res[0][2], res[1][2]), "the two GPUs disagree on the same batch"
    feats = res[0][2][:, 0, :64].astype(np.float64)
    for rank, same, _, mu, sigma in res:
        assert same, f"rank {rank}: gathered (sharded) poses differ from the 1-GPU poses"
        np.testing.assert_allclose(mu, feats.mean(0), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(sigma, np.cov(feats, rowvar=False), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(res[0][3], res[1][3])
    np.testing.assert_array_equal(res[0][4], res[1][4])
