"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

The reference trains with ``accelerator='dp'`` on a single GPU (GAN_final.py:481-485): replicas normalise their
own shard (per-replica BatchNorm statistics) and only gradients are exchanged.  Here each rank owns full replicas
of G and D in flat fp32 buffers, the global batch is sharded over ranks, and each network's flat gradient
buffer is averaged with ONE all-reduce per optimizer pass -- 9.7 MB for G, 10.4 MB for D.
Batch-norm running statistics stay per replica (rank 0's are the ones checkpointed, like DataParallel's replica 0).
"""
import os

import torch
import torch.distributed as dist


class GradComm:
    """Averages a flat gradient buffer across the process group."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.backend = dist.get_backend(group)
        self.calls = 0
        self.bytes = 0

    def allreduce(self, flat):
        self.calls += 1
        self.bytes += flat.numel() * flat.element_size()
        if self.world == 1:
            return flat
        if self.backend == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo (CPU tests of the host logic): no AVG
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        return flat

    def broadcast_parameters(self, model):
        """Make every replica start from rank 0's weights (DataParallel replicates module 0)."""
        for net in (model.generator, model.discriminator):
            rt = net.runtime
            if rt.flat is not None:
                dist.broadcast(rt.flat, 0, group=self.group)
                rt.mark_dirty()
                rt.refresh_shadows(force=True)
            else:
                for p in net.parameters():
                    dist.broadcast(p.data, 0, group=self.group)
            for b in net.buffers():
                dist.broadcast(b.data, 0, group=self.group)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* variables."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def attach(model, group=None):
    """Give ``model`` (mpgan.GAN) a gradient communicator; fit_batch / fused_step call it after each backward."""
    model.comm = GradComm(group)
    return model.comm


def shard_batch(batch, rank, world):
    """Contiguous shard of the global batch for this rank (global batch must divide evenly)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        if n % world:
            raise ValueError(f"global batch {n} is not divisible by world size {world}")
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out
