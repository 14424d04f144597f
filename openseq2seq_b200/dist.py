"""torch.distributed (NCCL over NVLink) stand-in for the three Horovod/MPI collectives the reference
uses on this path:

  C1 hvd.allreduce(grad) per variable   -> ONE all-reduce (sum) over the flat fp32 gradient buffer,
                                           optionally split into buckets launched on a side stream;
                                           the 1/N is folded into the optimizer's unscale factor
  C2 hvd.broadcast of every global var  -> broadcast of the flat master / momentum / BN buffers
  C3 MPI gather of scalars              -> all_reduce of a scalar
"""
import os

import torch
import torch.distributed as dist


class TorchDistHvd(object):
    def __init__(self, rank, size, local_rank):
        self._rank, self._size, self._local_rank = rank, size, local_rank

    @classmethod
    def init(cls, backend=None):
        """backend: "nccl" on GPUs (default when CUDA is available); "gloo" is used by the CPU tests of
        the host-side logic (world_size 2, no GPU)."""
        rank = int(os.environ["RANK"])
        world = int(os.environ["WORLD_SIZE"])
        local = int(os.environ.get("LOCAL_RANK", rank))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # the all-reduce overlaps persistent one-CTA-per-SM conv kernels: every SM NCCL occupies is one
            # the convolutions lose, and 16 channels already saturate NVLink for 128 MB buckets
            # (tools/n2_sweep.sh: 43.5 -> 42.8 ms/step at N = 2); the user's own setting wins
            os.environ.setdefault("NCCL_MAX_CTAS", "16")
            torch.cuda.set_device(local)
            if not dist.is_initialized():
                dist.init_process_group("nccl", rank=rank, world_size=world,
                                        device_id=torch.device("cuda", local))
        elif not dist.is_initialized():
            dist.init_process_group(backend, rank=rank, world_size=world)
        obj = cls(rank, world, local)
        obj._device = "cuda" if backend == "nccl" else "cpu"
        return obj

    @classmethod
    def single(cls):
        return cls(0, 1, 0)

    def size(self):
        return self._size

    def rank(self):
        return self._rank

    def local_rank(self):
        return self._local_rank

    def allreduce_(self, flat):
        if self._size > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def broadcast_parameters(self, engine):
        if self._size > 1:
            for t in [engine.master, engine.mom, engine.fstate, engine.istate] + list(engine.moving.values()):
                dist.broadcast(t, src=0)
            engine.sync_half_copies()

    def sum_scalar(self, x):
        if self._size == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=getattr(self, "_device", "cuda"))
        dist.all_reduce(t)
        return float(t[0])

    def barrier(self):
        if self._size > 1:
            dist.barrier()
