"""Stand-in for the three Horovod/MPI collectives the reference uses on this path;
torch.distributed (NCCL) provides the rendezvous and the fallbacks:

  C1 hvd.allreduce(grad) per variable   -> sum of the flat fp32 gradient buffer, bucket by bucket on a side
                                           stream: over NVLink peer memory (PeerGradExchange / csrc/peer.cu) on
                                           one node, else NCCL all-reduce; the 1/N is folded into the optimizer's
                                           unscale factor
  C2 hvd.broadcast of every global var  -> broadcast of the flat master / momentum / BN buffers
  C3 MPI gather of scalars              -> all_reduce of a scalar
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

from . import _lib as L


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

    def make_peer_exchange(self, grad, buckets, timeout_s=None):
        """Gradient sum over NVLink peer memory (csrc/peer.cu) for this process group, or None when it cannot be
        set up (not NCCL / not one node / CUDA IPC refused / the self-test does not reproduce the closed-form sums) -- the decision is
        taken jointly so every rank uses the same transport."""
        if self._size < 2 or getattr(self, "_device", "cpu") != "cuda":
            return None
        mode = os.environ.get("OS2S_GRAD_EXCHANGE", "auto")
        if mode == "nccl":
            return None
        if int(os.environ.get("LOCAL_WORLD_SIZE", self._size)) != self._size:
            return None          # more than one node: peer memory does not reach
        px = PeerGradExchange(self, grad, buckets,
                              float(os.environ.get("OS2S_PEER_TIMEOUT_S", timeout_s or 600.0)))
        if not px.ok:
            if self._rank == 0:
                sys.stderr.write("[os2s] peer-memory gradient exchange unavailable (%s); using the NCCL all-reduce\n"
                                 % px.why)
            if mode == "peer":
                raise RuntimeError("OS2S_GRAD_EXCHANGE=peer but the peer-memory exchange failed: %s" % px.why)
            return None
        return px


def split_bucket(start, end, world):
    """Slices [lo, hi) of bucket [start, end) owned by each rank (16-byte aligned chunks; csrc/peer.cu)."""
    n = end - start
    chunk = (-(-n // world) + 3) & ~3
    out = []
    for r in range(world):
        lo = min(end, start + r * chunk)
        out.append((lo, min(end, lo + chunk)))
    return out


class PeerGradExchange(object):
    """Host side of csrc/peer.cu: exports this rank's flat gradient buffer and a staging buffer through CUDA IPC,
    maps the other ranks' buffers, and enqueues the per-bucket exchange.  torch.distributed is used for the
    rendezvous of the 64-byte handles only."""

    def __init__(self, hvd, grad, buckets, timeout_s):
        self.ok, self.why = False, ""
        self.rank, self.world = hvd.rank(), hvd.size()
        self.buckets = [(int(a), int(b)) for a, b in buckets]
        self.grad = grad
        self._ctx = ctypes.c_void_p(0)
        self._opened = {}
        lib = L.load()
        self._lib = lib
        nb = len(self.buckets)
        LL = ctypes.c_longlong * nb
        starts, ends = LL(*[a for a, _ in self.buckets]), LL(*[b for _, b in self.buckets])
        lib.os2s_peer_stage_bytes.restype = ctypes.c_longlong
        mine, err = None, ""
        try:
            nbytes = lib.os2s_peer_stage_bytes(self.world, nb, starts, ends)
            if nbytes <= 0:
                raise RuntimeError("bad bucket list")
            self.stage = torch.zeros(int(nbytes), dtype=torch.uint8, device=grad.device)
            torch.cuda.synchronize()
            mine = self._export(grad) + self._export(self.stage)
        except Exception as e:       # noqa: BLE001 -- every failure is reported to the other ranks below
            err = "rank %d: %s" % (self.rank, e)
        table = [None] * self.world
        dist.all_gather_object(table, (mine, err))
        errs = [e for _, e in table if e]
        if errs:
            self.why = "; ".join(errs)
            return
        err = ""
        try:
            gp, sp = [], []
            for r, (h, _) in enumerate(table):
                if r == self.rank:
                    gp.append(grad.data_ptr())
                    sp.append(self.stage.data_ptr())
                else:
                    gp.append(self._open(h[0]) + h[1])
                    sp.append(self._open(h[2]) + h[3])
            VP = ctypes.c_void_p * self.world
            L.check(lib.os2s_peer_create(self.rank, self.world, VP(*gp), VP(*sp), nb, starts, ends,
                                         ctypes.c_double(timeout_s), ctypes.byref(self._ctx)), "os2s_peer_create")
        except Exception as e:       # noqa: BLE001
            err = "rank %d: %s" % (self.rank, e)
        dist.all_gather_object(table, err)
        errs = [e for e in table if e]
        if errs:
            self.why = "; ".join(errs)
            return
        torch.cuda.synchronize()
        dist.barrier()               # every staging buffer is zeroed and mapped before the first flag is raised
        # the ranks are barrier-aligned here: a flag that does not arrive within seconds never will
        L.check(lib.os2s_peer_set_timeout(self._ctx, ctypes.c_double(min(timeout_s, 8.0))), "os2s_peer_set_timeout")
        good = self._self_test()
        L.check(lib.os2s_peer_set_timeout(self._ctx, ctypes.c_double(timeout_s)), "os2s_peer_set_timeout")
        t = torch.tensor([1 if good else 0], dtype=torch.int32, device=grad.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t[0]) != 1:
            self.why = "self-test mismatch or time-out"
            return
        self.ok = True

    def _export(self, t):
        h = (ctypes.c_ubyte * 64)()
        off = ctypes.c_longlong(0)
        L.check(self._lib.os2s_ipc_export(ctypes.c_void_p(t.data_ptr()), h, ctypes.byref(off)), "os2s_ipc_export")
        return (bytes(h), int(off.value))

    def _open(self, handle):
        if handle not in self._opened:        # one allocation may hold several exported tensors
            base = ctypes.c_void_p(0)
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
            L.check(self._lib.os2s_ipc_open(buf, ctypes.byref(base)), "os2s_ipc_open")
            self._opened[handle] = int(base.value)
        return self._opened[handle]

    def exchange_bucket(self, index, stream):
        """Enqueue the two-phase sum of bucket `index` on `stream` (a torch stream)."""
        L.check(self._lib.os2s_peer_exchange_bucket(self._ctx, int(index), ctypes.c_void_p(stream.cuda_stream)),
                "os2s_peer_exchange_bucket")

    def finish(self, stream):
        """Enqueue the wait for the summed slices of every bucket: afterwards grad holds the sum on this rank."""
        L.check(self._lib.os2s_peer_finish(self._ctx, ctypes.c_void_p(stream.cuda_stream)), "os2s_peer_finish")

    def timed_out(self):
        f = ctypes.c_int(0)
        L.check(self._lib.os2s_peer_timed_out(self._ctx, ctypes.byref(f)), "os2s_peer_timed_out")
        return bool(f.value)

    def allreduce_all(self, stream=None):
        """All buckets + finish on one stream (tests, iter_size > 1)."""
        stream = stream or torch.cuda.current_stream()
        for b in range(len(self.buckets)):
            self.exchange_bucket(b, stream)
        self.finish(stream)

    def _self_test(self):
        """Sum small integers (exact in fp32) through the exchange and compare with the closed form."""
        g = self.grad
        saved = g.clone()
        try:
            n = g.numel()
            base = (torch.arange(n, device=g.device, dtype=torch.int64) % 251).to(torch.float32)
            g.copy_(base + float(3 * self.rank + 1))
            torch.cuda.synchronize()
            dist.barrier()
            self.allreduce_all()
            torch.cuda.synchronize()
            if self.timed_out():
                return False
            want = base * float(self.world) + float(sum(3 * r + 1 for r in range(self.world)))
            lo = min(a for a, _ in self.buckets)
            hi = max(b for _, b in self.buckets)
            good = bool(torch.equal(g[lo:hi], want[lo:hi]))
            # outside the buckets nothing may change
            good = good and bool(torch.equal(g[:lo], (base + float(3 * self.rank + 1))[:lo]))
            dist.barrier()
            return good
        except Exception:            # noqa: BLE001
            return False
        finally:
            g.copy_(saved)
            torch.cuda.synchronize()
