"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

The reference trains with ``accelerator='dp'`` on a single GPU (GAN_final.py:481-485): replicas normalise their
own shard (per-replica BatchNorm statistics) and only gradients are exchanged.  Here each rank owns full replicas
of G and D in flat fp32 buffers, the global batch is sharded over ranks, and each network's flat gradient
buffer is averaged per optimizer pass -- 9.7 MB for G, 10.4 MB for D -- overlapped with compute
in the fused step (``GAN.fused_step``): G's bucket reduces on the NCCL stream while the discriminator's real-batch
forward (independent of G's update) runs; D's bucket is split at layer 3: {D3, D4, Linear} (92 % of D's parameters) starts
the moment layer 3's weight gradient of the last backward has been enqueued and reduces under the backward of layers
2 and 1; the {D1, D2} remainder (0.3 MB) follows the pass.
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

    # ---- overlapped form: start() returns at once, finish() makes the CURRENT stream wait for the result ----
    def start(self, flat, stream=None):
        """Begin averaging ``flat`` (a contiguous slice of a flat gradient buffer) and return a handle.  The collective
        runs on the process group's own NCCL stream, ordered after everything already enqueued on ``stream`` (default:
        the current stream) -- kernels the caller enqueues next on its stream overlap with it.  Inside a CUDA-graph
        capture the NCCL stream becomes a parallel branch of the graph."""
        self.calls += 1
        self.bytes += flat.numel() * flat.element_size()
        if self.world == 1:
            return None
        avg = self.backend == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        if stream is not None and flat.is_cuda:
            with torch.cuda.stream(stream):
                work = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        return (work, flat, avg)

    def finish(self, handle):
        """The current stream (and, for gloo, the host) waits for the collective started by ``start``."""
        if handle is None:
            return
        work, flat, avg = handle
        work.wait()
        if not avg:
            flat.div_(self.world)

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
