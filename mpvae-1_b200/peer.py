"""Peer-memory plumbing of the data-parallel g_R sum (csrc/peer_reduce.cu): one process per GPU on one node; every rank
allocates a partial buffer, a g_R buffer and a flag block through the library (cudaMalloc + CUDA IPC), exchanges the
handles over the torch.distributed group and maps the others'.  `PeerRing.fill(params)` then puts the tables into
`mpvae_probit_params`, and `mpvae_probit_backward` leaves the SUM of g_R over all ranks in `ring.g_r`: every rank owns
a chunk of g_R, pulls it from all ranks over NVLink, adds in rank order and stores the result to all ranks -- one
kernel inside the backward instead of an NCCL all-reduce after it.

torch.distributed is used for exactly one thing here: `all_gather_object` of the 64-byte handles at construction.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib


class _DevMem:
    """A cudaMalloc'ed buffer of the library exposed to torch without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr, "version": 3,
                                         "strides": None}


class _PeerBuffers:
    """Allocation + handle exchange shared by the rings below: `sizes` = {name: bytes}; afterwards `self.ptrs[name][r]`
    is rank r's buffer as seen from this process."""

    MAX_WORLD = 8

    def _setup(self, sizes, device, group):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError(f"{type(self).__name__} needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= self.MAX_WORLD:
            raise ValueError(f"{type(self).__name__}: world size {self.world} not in [2, {self.MAX_WORLD}]")
        lib = _lib.lib()
        self.device = torch.device(device)
        self._local, self._remote, handles = {}, {}, {}
        # Every collective below is entered by EVERY rank whatever happened locally, and the outcome is agreed on
        # collectively: either all ranks own a working ring or all of them raise (no rank is left waiting).
        error = None
        with torch.cuda.device(self.device):
            try:
                if os.environ.get("MPVAE_PEER_TEST_FAIL") == str(self.rank):     # test hook: exercise the collective fallback
                    raise RuntimeError("simulated set-up failure (MPVAE_PEER_TEST_FAIL)")
                for name, nbytes in sizes.items():
                    ptr, h = C.c_void_p(), C.create_string_buffer(64)
                    _lib.check(lib.mpvae_peer_alloc(nbytes, C.byref(ptr), h), "mpvae_peer_alloc")
                    self._local[name] = ptr.value
                    handles[name] = h.raw
            except Exception as e:      # noqa: BLE001
                error, handles = e, None
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handles, group=group)
            self.ptrs = {name: [0] * self.world for name in sizes}
            if error is None and all(h is not None for h in gathered):
                try:
                    for r, hs in enumerate(gathered):
                        for name in sizes:
                            if r == self.rank:
                                self.ptrs[name][r] = self._local[name]
                                continue
                            ptr = C.c_void_p()
                            _lib.check(lib.mpvae_peer_open(hs[name], C.byref(ptr)), "mpvae_peer_open")
                            self.ptrs[name][r] = ptr.value
                            self._remote.setdefault(name, []).append(ptr.value)
                except Exception as e:  # noqa: BLE001
                    error = e
            elif error is None:
                error = RuntimeError("a peer rank could not allocate its buffers")
            verdicts = [None] * self.world
            dist.all_gather_object(verdicts, None if error is None else str(error), group=group)
            if any(v is not None for v in verdicts):
                self._release()
                raise RuntimeError(f"{type(self).__name__}: set-up failed on rank(s) " +
                                   ", ".join(f"{r}: {v}" for r, v in enumerate(verdicts) if v is not None))

    def _release(self):
        lib = _lib.lib()
        for ptrs in self._remote.values():
            for ptr in ptrs:
                lib.mpvae_peer_close(C.c_void_p(ptr))
        self._remote = {}
        for ptr in self._local.values():
            lib.mpvae_peer_free(C.c_void_p(ptr))
        self._local = {}

    def check(self):
        """Raise if a flag wait of this rank has ever timed out (a peer was more than MPVAE_PEER_TIMEOUT_S late): the
        sums of that step are invalid and the caller should fall back to the NCCL all-reduce.  Synchronises the stream."""
        lib = _lib.lib()
        step = C.c_uint32(0)
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.mpvae_peer_error(C.c_void_p(self.ptrs["flags"][self.rank]), C.byref(step), stream), "mpvae_peer_error")
        if step.value:
            raise RuntimeError(f"{type(self).__name__}: rank {self.rank} gave up waiting for a peer at exchange step {step.value}")

    def enable_graph_replay(self):
        """Flag values come from a DEVICE counter the exchange kernels advance themselves, so that a captured CUDA graph
        (which replays identical launch arguments) keeps them increasing.  Call on every rank before capturing; from
        then on the host hands out the constant base step."""
        if getattr(self, "step_dev", None) is None:
            self.step_dev = torch.full((1,), self.step, dtype=torch.int32, device=self.device)
        return self.step_dev

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)      # nobody unmaps while a peer may still be reading
        self._release()


class PeerBucket(_PeerBuffers):
    """One flat fp32 buffer per rank in peer-mapped memory, summed over the ranks IN PLACE (a range at a time) by the
    library's exchange kernel: what the data-parallel step keeps its gradient bucket in when it runs without NCCL
    (train.DataParallelStep(peer_all=True)), which is also what lets the whole step be captured as a CUDA graph on
    every rank (NCCL collectives cannot be captured on this stack)."""

    def __init__(self, numel: int, device, group=None):
        lib = _lib.lib()
        self.numel = int(numel)
        self._setup({"buf": self.numel * 4, "flags": int(lib.mpvae_peer_flag_bytes())}, device, group)
        with torch.cuda.device(self.device):
            self.flat = torch.as_tensor(_DevMem(self._local["buf"], (self.numel,), "<f4"), device=self.device)
        self.step = 0
        dist.barrier(group=group)

    def allreduce(self, first: int = 0, n: int = None):
        """flat[first : first + n] = its sum over the ranks, on every rank (bit-identical); `first` a multiple of 4."""
        n = self.numel - first if n is None else int(n)
        if first % 4 or first < 0 or n <= 0 or first + n > self.numel:
            raise ValueError(f"PeerBucket.allreduce: bad range [{first}, {first + n}) of {self.numel}")
        lib = _lib.lib()
        step_dev = getattr(self, "step_dev", None)
        if step_dev is None:
            self.step += 1
            step = self.step
        else:
            step = 1
        table = (C.c_void_p * self.world)(*[p + 4 * first for p in self.ptrs["buf"]])
        flags = (C.c_void_p * self.world)(*self.ptrs["flags"])
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.mpvae_peer_allreduce_dev(table, table, flags, self.world, self.rank, step,
                                                    C.c_void_p(step_dev.data_ptr()) if step_dev is not None else C.c_void_p(0),
                                                    n, stream), "mpvae_peer_allreduce_dev")
        return self.flat

    def _release(self):
        self.flat = None
        super()._release()


class PeerRing(_PeerBuffers):

    def __init__(self, L: int, Z: int, device, group=None):
        lib = _lib.lib()
        self.L, self.Z = int(L), int(Z)
        # "tiles": per-tile completion counters of the exchange beside / inside the g_R product (monotonic, zeroed once here)
        sizes = {"part": self.L * self.Z * 4, "g_r": self.L * self.Z * 4, "flags": int(lib.mpvae_peer_flag_bytes()),
                 "tiles": _lib.PEER_TILE_BYTES}
        self._setup(sizes, device, group)
        with torch.cuda.device(self.device):
            self.g_r = torch.as_tensor(_DevMem(self._local["g_r"], (self.L, self.Z), "<f4"), device=self.device)
            self.part = torch.as_tensor(_DevMem(self._local["part"], (self.L, self.Z), "<f4"), device=self.device)
        self.step = 0
        dist.barrier(group=group)          # everybody has mapped everybody before the first kernel touches a peer

    def applies(self, S: int, B: int, L: int, Z: int, flags: int) -> bool:
        return (L, Z) == (self.L, self.Z) and B > 0

    def fill(self, p) -> torch.Tensor:
        """Put the peer tables of the next step into `p` (a _lib.ProbitParams); returns the tensor g_R lands in."""
        step_dev = getattr(self, "step_dev", None)
        if step_dev is None:
            self.step += 1
            p.peer_world, p.peer_rank, p.peer_step = self.world, self.rank, self.step
        else:
            # device counter: flag value = 1 + counter, the kernels bump the counter after every exchange
            p.peer_world, p.peer_rank, p.peer_step = self.world, self.rank, 1
            p.peer_step_dev = step_dev.data_ptr()
        for r in range(self.world):
            p.peer_part[r] = self.ptrs["part"][r]
            p.peer_g_r[r] = self.ptrs["g_r"][r]
            p.peer_flags[r] = self.ptrs["flags"][r]
            if "tiles" in self.ptrs:
                p.peer_tile_done[r] = self.ptrs["tiles"][r]
        return self.g_r

    def allreduce(self, n: int = None):
        """Stand-alone exchange: `self.g_r` (flat, first n floats) = sum over ranks of `self.part`."""
        lib = _lib.lib()
        n = self.L * self.Z if n is None else int(n)
        if getattr(self, "step_dev", None) is not None:
            raise RuntimeError("PeerRing.allreduce is not available once enable_graph_replay() has moved the step counter "
                               "to the device")
        self.step += 1
        tables = []
        for name in ("part", "g_r", "flags"):
            arr = (C.c_void_p * self.world)(*self.ptrs[name])
            tables.append(arr)
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.mpvae_peer_allreduce(tables[0], tables[1], tables[2], self.world, self.rank,
                                                self.step, n, stream),
                       "mpvae_peer_allreduce")
        return self.g_r

    def _release(self):
        self.g_r = self.part = None
        super()._release()


class NvlsRing(PeerRing):
    """Same exchange with the reduction done INSIDE the NVSwitch (csrc/peer_reduce.cu::peer_reduce_nvls_kernel):
    `multimem.ld_reduce` of the owner's chunk from a multicast address returns the sum over all ranks, `multimem.st`
    writes it to every rank (measured: as fast as the pull kernel, not faster; opt-in).  The multicast mapping of the buffers
    between the processes is torch's symmetric memory (`torch.distributed._symmetric_memory`: plumbing only -- the
    kernels are the library's)."""

    def __init__(self, L: int, Z: int, device, group=None):   # noqa: D401 -- deliberately does not call PeerRing.__init__
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NvlsRing needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= self.MAX_WORLD:
            raise ValueError(f"NvlsRing: world size {self.world} not in [2, {self.MAX_WORLD}]")
        self.L, self.Z = int(L), int(Z)
        self.device = torch.device(device)
        pg = group if group is not None else dist.group.WORLD
        n = self.L * self.Z
        lib = _lib.lib()
        error = None
        try:
            with torch.cuda.device(self.device):
                self._part_t = symm_mem.empty(n, dtype=torch.float32, device=self.device)
                self._g_r_t = symm_mem.empty(n, dtype=torch.float32, device=self.device)
                self._flags_t = symm_mem.empty(int(lib.mpvae_peer_flag_bytes()) // 4, dtype=torch.int32, device=self.device)
                self._flags_t.zero_()
                hp, hg, hf = (symm_mem.rendezvous(t, pg) for t in (self._part_t, self._g_r_t, self._flags_t))
                if not (hp.multicast_ptr and hg.multicast_ptr):
                    raise RuntimeError("no multicast (NVLS) support for this group")
                self.ptrs = {"part": list(hp.buffer_ptrs), "g_r": list(hg.buffer_ptrs), "flags": list(hf.buffer_ptrs)}
                self.mc_part, self.mc_g_r = int(hp.multicast_ptr), int(hg.multicast_ptr)
                self._handles = (hp, hg, hf)
        except Exception as e:      # noqa: BLE001
            error = e
        verdicts = [None] * self.world
        dist.all_gather_object(verdicts, None if error is None else str(error), group=group)
        if any(v is not None for v in verdicts):
            raise RuntimeError("NvlsRing: set-up failed on rank(s) " +
                               ", ".join(f"{r}: {v}" for r, v in enumerate(verdicts) if v is not None))
        self.part = self._part_t.view(self.L, self.Z)
        self.g_r = self._g_r_t.view(self.L, self.Z)
        self._local, self._remote = {}, {}
        self.step = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def fill(self, p) -> torch.Tensor:
        g = super().fill(p)
        p.peer_mc_part, p.peer_mc_g_r = self.mc_part, self.mc_g_r
        return g

    def allreduce(self, n: int = None):
        lib = _lib.lib()
        n = self.L * self.Z if n is None else int(n)
        self.step += 1
        tables = [(C.c_void_p * self.world)(*self.ptrs[name]) for name in ("part", "g_r", "flags")]
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.mpvae_peer_allreduce_nvls(tables[0], tables[1], tables[2], C.c_void_p(self.mc_part),
                                                     C.c_void_p(self.mc_g_r), self.world, self.rank,
                                                     self.step, n, stream), "mpvae_peer_allreduce_nvls")
        return self.g_r

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        self.g_r = self.part = None
        self._handles = None
        self._part_t = self._g_r_t = self._flags_t = None
