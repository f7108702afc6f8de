"""Multi-GPU plumbing: one process per GPU, trajectories sharded over ranks, one result exchange per control step -- fused into the
SSE kernel over peer memory (FusedGather) or as a separate NCCL all-gather (pack_block / all_gather_block).

The reference has no distributed code: its only parallelism is one OS process per trajectory
(`num_of_actors` forked actors, quartic main_parallel.py:254-369).  Here the batch axis shards over ranks with NO data-path
collective (trajectories are independent); the only exchange is the all-gather of the per-control-step result
block [B_local, K + 4 + 1] (moments, aux, flags) to every rank (the learner reads rank 0's copy).  In-kernel Philox noise is
keyed by the GLOBAL trajectory id, so results do not depend on the number of ranks.
"""
import os


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` trajectories owned by `rank` (first `total % world` ranks get one extra)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def pack_block(moments, aux, flags):
    """[B_local, K] f64, [B_local, 4] f64, [B_local] u8 -> one contiguous [B_local, K+5] f64 block (one collective, not three)."""
    import torch
    return torch.cat([moments, aux, flags.to(torch.float64).unsqueeze(1)], dim=1).contiguous()


def unpack_block(block, K):
    import torch
    return block[:, :K], block[:, K:K + 4], block[:, K + 4].to(torch.uint8)


def all_gather_block(block, world, out=None):
    """All-gather equal-sized per-rank blocks along dim 0.  Returns the [world*B_local, C] tensor (every rank)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return block
    if out is None:
        out = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block)
    return out


# ---- fused result exchange over peer memory (include/qcart.h: qc_set_gather) ---------------------------------------------------------
class _DeviceArray:
    """Raw device memory as a __cuda_array_interface__ object, so that torch can view it without owning it."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class FusedGather:
    """The per-control-step result block [world * B, K + 5] without a collective: each rank's SSE kernel stores its rows into its own
    gather area and publishes a sequence flag in every rank's flag array (CUDA-IPC peer memory over NVLink); `wait()` enqueues the
    consumer kernel that waits for the flags and pulls the peers' rows.  Replaces pack_block + all_gather_block (one NCCL collective and
    three small kernels per control step).

    Multi-process use (one rank per GPU, torch.distributed initialised):   fg = FusedGather(sim, rank, world)
    Single-process use (tests; several sims acting as ranks on one device): fgs = FusedGather.local_group(sims)
    """

    NBUF = 4        # QC_GATHER_BUFS: buffer = sequence number mod 4 (allows "enqueue step k, then consume step k-1", see include/qcart.h)

    def __init__(self, sim, rank, world, _wire=True):
        import ctypes as C
        from . import _lib as L
        self.sim, self.rank, self.world, self.lib = sim, int(rank), int(world), L.load()
        self.cols, self.rows = sim.K + L.QC_AUX_COUNT + 1, self.world * sim.B
        self._C, self._L = C, L
        self.peer_g, self.peer_f, self._opened = [None] * self.world, [None] * self.world, []
        self.gather_ptr = self.flag_ptr = None
        err = None
        try:
            self.gather_ptr, self.gather_handle = self._alloc(self.NBUF * self.rows * self.cols * 8)
            self.flag_ptr, self.flag_handle = self._alloc(max(self.world, 1) * 8)
            self.peer_g[self.rank], self.peer_f[self.rank] = self.gather_ptr, self.flag_ptr
        except Exception as e:                      # noqa: BLE001  (still take part in the handle exchange below: it is a collective)
            err = e
        if _wire:
            if self.world > 1:
                import torch.distributed as dist
                handles = [None] * self.world
                dist.all_gather_object(handles, None if err is not None else (self.gather_handle, self.flag_handle, int(sim.B)))
                if err is None and any(h is None for h in handles):
                    err = RuntimeError("a peer rank could not allocate its gather area")
                if err is None and any(h[2] != int(sim.B) for h in handles):
                    # the kernel addresses row rank * B + b in EVERY rank's area: unequal shards would store out of bounds in peer memory
                    err = ValueError("FusedGather needs the same batch size on every rank, got %s (pad the shards, e.g. dist.shard_range on a multiple of world)" % [h[2] for h in handles])
                if err is None:
                    try:
                        for r, (hg, hf, _b) in enumerate(handles):
                            if r != self.rank:
                                self.peer_g[r], self.peer_f[r] = self._open(hg), self._open(hf)
                    except Exception as e:          # noqa: BLE001
                        err = e
            if err is not None:
                self.close(collective=False)
                raise err
            self._activate()
        elif err is not None:
            raise err

    def _alloc(self, nbytes):
        C = self._C
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        self._L.check(self.lib.qc_peer_alloc(self.sim.device, int(nbytes), C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def _open(self, handle):
        C = self._C
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._L.check(self.lib.qc_peer_open(self.sim.device, buf, C.byref(ptr)))
        self._opened.append(ptr.value)
        return ptr.value

    def _activate(self):
        C = self._C
        g = (C.c_void_p * self.world)(*self.peer_g)
        f = (C.c_void_p * self.world)(*self.peer_f)
        self._L.check(self.lib.qc_set_gather(self.sim.h, self.rank, self.world, g, f))

    @classmethod
    def local_group(cls, sims):
        """Several sims on the devices of ONE process play the ranks (no IPC: the pointers are passed directly)."""
        if len({int(s.B) for s in sims}) != 1:
            raise ValueError("FusedGather needs the same batch size on every rank")
        fgs = [cls(s, r, len(sims), _wire=False) for r, s in enumerate(sims)]
        for a in fgs:
            for r, b in enumerate(fgs):
                a.peer_g[r], a.peer_f[r] = b.gather_ptr, b.flag_ptr
            a._activate()
        return fgs

    def seq(self):
        return int(self.lib.qc_gather_seq(self.sim.h))

    def wait(self, seq=None):
        """Enqueue (on the current stream) the wait for every rank's rows of control step `seq` (default: the last qc_step of this rank)."""
        self._L.check(self.lib.qc_gather_wait(self.sim.h, self.seq() if seq is None else int(seq), self.sim._stream()))

    def failed_ranks(self):
        """Ranks a wait() gave up on (bounded spin of the consumer kernel); synchronises the device.  Empty list = healthy."""
        mask = self._C.c_uint32(0)
        self._L.check(self.lib.qc_gather_error(self.sim.h, self._C.byref(mask)))
        return [r for r in range(self.world) if mask.value >> r & 1]

    def block(self, seq=None):
        """[world * B, K + 5] float64 view of the buffer holding control step `seq`; valid after wait(seq) in stream order."""
        import torch
        seq = self.seq() if seq is None else int(seq)
        off = (seq % self.NBUF) * self.rows * self.cols * 8
        return torch.as_tensor(_DeviceArray(self.gather_ptr + off, (self.rows, self.cols), "<f8"), device="cuda:%d" % self.sim.device)

    def close(self, collective=True):
        """collective=True (normal shutdown after steps ran): every rank must call it, two barriers keep mappings alive until nobody stores
        into them any more.  collective=False: abort before the exchange was ever used (no barriers)."""
        import torch
        if getattr(self, "sim", None) is None:
            return
        torch.cuda.synchronize(self.sim.device)
        sync = collective and self.world > 1 and bool(self._opened)      # (single-process groups have no IPC mappings)
        if sync:
            import torch.distributed as dist
            dist.barrier()                           # nobody may still be storing into a mapping that is about to go away
        self.lib.qc_set_gather(self.sim.h, 0, 0, None, None)
        for p in self._opened:
            self.lib.qc_peer_close(self.sim.device, self._C.c_void_p(p))
        if sync:
            import torch.distributed as dist
            dist.barrier()
        for ptr in (self.gather_ptr, self.flag_ptr):
            if ptr:
                self.lib.qc_peer_free(self.sim.device, self._C.c_void_p(ptr))
        self.sim = None
