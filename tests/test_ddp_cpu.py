"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, flat-bucket gradient averaging,
parameter broadcast.  The averaged bucket must equal the mean of the per-rank gradients the oracle computes
independently on each shard."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpgan import ddp
from oracle.gan import GANOracle, synthetic_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w, _ = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    comm = ddp.GradComm()
    # per-rank oracle gradients of the discriminator on this rank's shard
    torch.manual_seed(0)
    model = GANOracle("final", dims=2, spatial=32, n_unet_blocks=1)
    if rank == 1:  # replicas start different on purpose; broadcast must fix that
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)

    class _Net:  # minimal stand-in with the attributes broadcast_parameters walks
        def __init__(self, m):
            self.m, self.runtime = m, type("R", (), {"flat": None})()

        def parameters(self):
            return self.m.parameters()

        def buffers(self):
            return self.m.buffers()

    holder = type("M", (), {"generator": _Net(model.generator), "discriminator": _Net(model.discriminator)})()
    comm.broadcast_parameters(holder)
    glob = synthetic_batch(4, 2, 32, seed=1)
    shard = ddp.shard_batch(glob, rank, world)
    assert shard["t1w"].shape[0] == 2 and torch.equal(shard["t1w"], glob["t1w"][2 * rank:2 * rank + 2])
    for p in model.generator.parameters():
        p.requires_grad_(False)
    loss = model.training_step(shard, 0, 1)
    loss.backward()
    ps = list(model.discriminator.parameters())
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    torch.save({"local": flat.clone(), "w0": ps[0].detach().clone()}, os.path.join(result_dir, f"local{rank}.pt"))
    local = flat.clone()
    comm.allreduce(flat)
    torch.save(flat, os.path.join(result_dir, f"avg{rank}.pt"))
    assert comm.calls == 1 and comm.bytes == flat.numel() * 4
    # the overlapped form the fused step uses: two slices of one flat buffer, started one after the other, finished later
    split = (local.numel() // 3) // 8 * 8
    h_hi = comm.start(local[split:])
    h_lo = comm.start(local[:split])
    comm.finish(h_hi), comm.finish(h_lo)
    assert torch.allclose(local, flat, rtol=1e-6, atol=1e-8) and comm.calls == 3
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    l0, l1 = torch.load(tmp_path / "local0.pt"), torch.load(tmp_path / "local1.pt")
    a0, a1 = torch.load(tmp_path / "avg0.pt"), torch.load(tmp_path / "avg1.pt")
    assert torch.equal(l0["w0"], l1["w0"]), "broadcast_parameters did not synchronise the replicas"
    assert not torch.equal(l0["local"], l1["local"])          # different shards -> different gradients
    assert torch.equal(a0, a1)                                 # same bucket on both ranks
    assert torch.allclose(a0, (l0["local"] + l1["local"]) / 2, rtol=1e-6, atol=1e-8)


def test_shard_batch_rejects_ragged():
    b = {"t1w": torch.zeros(5, 1, 4, 4)}
    with pytest.raises(ValueError):
        ddp.shard_batch(b, 0, 2)
