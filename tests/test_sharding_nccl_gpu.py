"""world_size-2 NCCL test of the multi-GPU plumbing on real GPUs (skipped on a one-GPU box): the flat-arena
checkpoint broadcast and the gather helper must work on the backend the product uses (gather_clips built
CPU buffers, which NCCL rejects), including a rank whose shard is empty."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pantomatrix_b200 import sharding
    torch.manual_seed(100 + rank)                                   # every rank starts from different weights
    net = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.BatchNorm1d(19), torch.nn.Linear(19, 5)).cuda()
    nbytes = sharding.broadcast_checkpoint(net, src=0)
    probe = sum(t.double().sum() for t in net.state_dict().values())
    sums = [torch.zeros((), dtype=torch.double, device="cuda") for _ in range(world)]
    dist.all_gather(sums, probe)
    assert all(torch.equal(s, sums[0]) for s in sums), "checkpoint broadcast did not equalise the ranks"
    # 3 clips over 2 ranks (2 + 1), then 1 clip over 2 ranks (rank 1 has nothing to send)
    s, e = sharding.shard_range(3, rank, world)
    local = torch.arange(s, e, device="cuda", dtype=torch.float32)[:, None, None].expand(e - s, 4, 3).contiguous()
    got3 = sharding.gather_clips(local, 3)
    s, e = sharding.shard_range(1, rank, world)
    got1 = sharding.gather_clips(torch.full((1, 2), 7, device="cuda", dtype=torch.int64) if e > s else None, 1)
    if rank == 0:
        torch.save({"got3": got3.cpu(), "got1": got1.cpu(), "nbytes": nbytes}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_nccl_broadcast_and_gather(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    out_path = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    assert got["got3"].shape == (3, 4, 3) and torch.equal(got["got3"][:, 0, 0], torch.tensor([0.0, 1.0, 2.0]))
    assert got["got1"].dtype == torch.int64 and got["got1"].tolist() == [[7, 7]]
    assert got["nbytes"] == sum(t.numel() * t.element_size() for t in
                                torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.BatchNorm1d(19), torch.nn.Linear(19, 5)).state_dict().values())
