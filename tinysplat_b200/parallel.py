"""Data-parallel rendering over cameras (new: the reference renders one camera per step on one
device [REF scripts/train.py:54-55]).

One process per GPU, each holding a full replica of the Gaussian parameters.  Every rank
renders its own view(s) forward+backward; the only exchange step is a sum of the per-Gaussian
parameter gradients (59 floats per Gaussian at SH degree 3).  Each parameter's all-reduce is
launched from a post-accumulate-grad hook the moment autograd has produced that gradient, so
the largest payload (colors_rest, 76% of the bytes, ready right after SH-backward) crosses
NVLink while projection-backward is still running.  Forward-only rendering needs no
collective at all (replicas only).

Second exchange strategy (`PackedGradExchange`, used through the fused node): instead of
all-reducing the finished gradients (236 B per Gaussian), ranks exchange blend-backward's packed
rows (48 B per view and Gaussian) with one all-to-all, every rank runs projection-/SH-backward
for ALL views on its own shard of the Gaussians, and the shard results are all-gathered: 42 MB +
206 MB per rank instead of the 413 MB a ring all-reduce of 236 MB moves at 8 ranks.

The densification statistic is the per-view norm of d loss / d xy summed over views
[REF tinysplat/splatting/model_gaussian.py:130-132] — NOT the norm of the reduced gradient —
so it gets its own [N] all-reduce (`reduce_densify_stat`)."""
from __future__ import annotations

import os

from typing import Callable, Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


class GradientAllReducer:
    """All-reduce (mean or sum) of .grad for a fixed list of leaf tensors.

    Launch happens from the post-accumulate-grad hook of the LAST parameter to receive its
    gradient, i.e. as early as autograd allows.  Gradients that are views of one flat buffer
    (the fused node allocates them that way) are reduced with a single collective over the
    buffer's span; anything else falls back to one collective per tensor.

    Contract: exactly ONE backward per finish().  A second backward before finish() (gradient
    accumulation, retain_graph) would add into gradients that are already being reduced — it raises
    instead.  A parameter that receives no gradient in a step simply leaves the launch to finish()."""

    def __init__(self, params: Iterable[Tensor], process_group=None, average: bool = True,
                 overlap: bool = True):
        self.params: List[Tensor] = list(params)
        self.group = process_group
        self.average = average
        self.overlap = overlap
        self._works = []       # (tensor that was reduced, work handle)
        self._handles = []
        self._ready = 0
        self._launched = False
        self.last_num_collectives = 0
        self._enabled = dist.is_available() and dist.is_initialized() and \
            dist.get_world_size(process_group) > 1
        self._native_avg = self._enabled and dist.get_backend(process_group) == "nccl"
        if self._enabled and overlap:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if self._enabled else 1

    def _hook(self, p: Tensor) -> None:
        if self._launched:
            raise RuntimeError("GradientAllReducer: a gradient arrived while the all-reduce of this step is in "
                               "flight — call finish() after every backward (one backward per finish())")
        self._ready += 1
        if self._ready == len(self.params):
            self._launch()

    def _spans(self):
        """Group gradients by storage; a group whose members tile one span densely becomes one
        flat tensor over that span."""
        groups = {}
        for p in self.params:
            if p.grad is not None:
                groups.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p.grad)
        out = []
        for grads in groups.values():
            if len(grads) > 1 and all(g.is_contiguous() for g in grads):
                lo = min(g.storage_offset() for g in grads)
                hi = max(g.storage_offset() + g.numel() for g in grads)
                used = sum(g.numel() for g in grads)
                if hi - lo <= used + 4 * len(grads):        # only alignment padding in between
                    flat = torch.empty(0, dtype=grads[0].dtype, device=grads[0].device)
                    flat.set_(grads[0].untyped_storage(), lo, (hi - lo,))
                    out.append(flat)
                    continue
            out.extend(grads)
        return out

    def _launch(self) -> None:
        op = dist.ReduceOp.AVG if (self.average and self._native_avg) else dist.ReduceOp.SUM
        for t in self._spans():
            self._works.append((t, dist.all_reduce(t, op=op, group=self.group, async_op=True)))
        self.last_num_collectives = len(self._works)
        self._launched = True

    def finish(self) -> None:
        """Wait for the in-flight reductions (launching them now if the hooks did not)."""
        if not self._enabled:
            return
        if not self._launched:
            self._launch()
        for t, w in self._works:
            w.wait()
            if self.average and not self._native_avg:
                t.div_(self.world_size)
        self._works.clear()
        self._ready = 0
        self._launched = False

    def payload_bytes(self) -> int:
        return sum(p.numel() * p.element_size() for p in self.params)

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


class PackedGradExchange:
    """Collectives of the packed-row gradient exchange (see module docstring); the shard kernels
    are launched by the fused autograd node (tinysplat_b200.fused), which receives this object.

    world ranks, rank r owns Gaussians [r*Ns, (r+1)*Ns) with Ns = shard_rows(N) (a multiple of
    128 so that every shard slice of every parameter stays 16-byte aligned)."""

    CAM_FLOATS = 32      # include/tinysplat_b200.h: 3x4 view | 4x4 full projection | fx fy | pad

    def __init__(self, process_group=None, average: bool = True):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PackedGradExchange needs an initialised process group")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.average = average
        self._coalesce = dist.get_backend(process_group) == "nccl"
        self._buffers = {}
        self.last_bytes_sent = 0

    def shard_rows(self, n_gaussians: int) -> int:
        per = -(-n_gaussians // self.world)
        return max(128, (per + 127) // 128 * 128)

    def out_scale(self) -> float:
        return 1.0 / self.world if self.average else 1.0

    def buffer(self, key, shape, dtype, device, zero: bool = False, tag=None) -> Tensor:
        """Persistent buffer for everything NCCL touches.  c10d runs collectives on its own
        stream and marks their tensors as in use there, so the caching allocator cannot hand a
        freed block back until that stream has passed it; with the host running several steps
        ahead of the GPU, per-step allocations of the 48-236 MB exchange buffers degenerate into
        cudaMalloc/cudaFree stalls (measured: 30 ms per step instead of 0.4 ms)."""
        t, old_tag = self._buffers.get(key, (None, None))
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != device or \
                old_tag != tag:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self._buffers[key] = (t, tag)        # replaces the previous buffer of this key
        return t

    def gather_cameras(self, cam_row: Tensor) -> Tensor:
        """cam_row[32] of this rank's view -> [world, 32] with every rank's camera (a fresh
        tensor: backward reads it long after the next forward has gathered the next cameras)."""
        src = self.buffer("cam_src", (cam_row.numel(),), cam_row.dtype, cam_row.device)
        out = self.buffer("cam_all", (self.world * cam_row.numel(),), cam_row.dtype, cam_row.device)
        src.copy_(cam_row.reshape(-1))
        dist.all_gather_into_tensor(out, src, group=self.group)
        return out.view(self.world, cam_row.numel()).clone()

    def all_to_all_rows(self, send: Tensor) -> Tensor:
        """send[world*Ns, F] (row block j goes to rank j) -> recv[world, Ns, F] (block i came from
        rank i: view i's rows of this rank's shard).  `send` should come from buffer(); the result
        is a persistent buffer that the next call overwrites."""
        rows = send.shape[0] // self.world
        recv = self.buffer("recv", (self.world, rows, send.shape[1]), send.dtype, send.device)
        dist.all_to_all_single(recv, send, group=self.group)
        self.last_bytes_sent = send.numel() * send.element_size() * (self.world - 1) // self.world
        return recv

    def all_gather_shards(self, shards: List[Tensor]) -> List[Tensor]:
        """[Ns, ...] per parameter -> [world*Ns, ...] per parameter, one NCCL group call.  The
        results are persistent buffers that the next call overwrites: clone what must survive."""
        outs = [self.buffer(("out", i), (self.world * t.shape[0],) + tuple(t.shape[1:]), t.dtype, t.device)
                for i, t in enumerate(shards)]
        if self._coalesce:
            with dist._coalescing_manager(group=self.group, device=shards[0].device, async_ops=False):
                for o, t in zip(outs, shards):
                    dist.all_gather_into_tensor(o, t, group=self.group)
        else:
            for o, t in zip(outs, shards):
                dist.all_gather_into_tensor(o, t, group=self.group)
        self.last_bytes_sent += sum(t.numel() * t.element_size() for t in shards) * (self.world - 1)
        return outs


class PeerLayout:
    """Byte layout of one rank's peer-visible allocation (identical on every rank) for a capacity of
    `cap_rows` Gaussians with K stored SH bases at `world` ranks.  Pure host arithmetic.

        flags | err | cams[world][32] | geo[world][Ns][8] | rgb[world][world*Ns][3] |
        v_rest[(K-1)*3 per row] | v_dc[3] | v_means[3] | v_scales[3] | v_quats[4] | v_logit[1]

    Every segment starts on a 256-byte boundary; the gradient segments are sized for cap_rows rows
    so that the six parameter gradients handed to autograd are views of this buffer."""

    SHARD_ALIGN = 256          # rows; keeps every shard slice of every tensor 16-byte aligned
    MAX_CHUNKS = 7             # the rows are pushed in up to this many pieces (two barrier slots each + a final one)

    def __init__(self, world: int, cap_rows: int, K: int, flag_bytes: int):
        self.world, self.cap_rows, self.K = world, cap_rows, K
        self.cap_shard = self.shard_rows(cap_rows, world)
        seg = {}
        off = 0

        def add(name, nbytes):
            nonlocal off
            seg[name] = (off, nbytes)
            off += (nbytes + 255) // 256 * 256
        add("flags", flag_bytes)
        add("err", 256)
        add("cams", world * PackedGradExchange.CAM_FLOATS * 4)
        # geometry rows are laid out per pushed chunk, each chunk padded to its own shard size
        add("geo", world * (self.cap_shard + self.MAX_CHUNKS * self.SHARD_ALIGN) * 8 * 4)
        add("rgb", world * world * self.cap_shard * 3 * 4)
        for name, width in (("rest", (K - 1) * 3), ("dc", 3), ("means", 3), ("scales", 3), ("quats", 4), ("logit", 1)):
            add("g_" + name, max(1, cap_rows * width) * 4)
        self.seg = seg
        self.total_bytes = off

    @classmethod
    def shard_rows(cls, n_gaussians: int, world: int) -> int:
        per = -(-max(n_gaussians, 1) // world)
        return (per + cls.SHARD_ALIGN - 1) // cls.SHARD_ALIGN * cls.SHARD_ALIGN

    def shard_of(self, rank: int, n_gaussians: int):
        """(first row, rows) of `rank`'s shard for the CURRENT number of Gaussians."""
        ns_all = self.shard_rows(n_gaussians, self.world)
        s0 = rank * ns_all
        return s0, max(0, min(n_gaussians, s0 + ns_all) - s0), ns_all

    def chunks(self, n_gaussians: int, n_chunks: int):
        """Splits the rows into <= n_chunks pieces that are pushed, signalled and consumed one after the
        other (the transfer of piece c+1 overlaps the shard backward of piece c).  Every piece is sharded
        over the ranks on its own.  -> [(first row, rows, shard rows of the piece, first row of the
        piece's geometry block in the geo segment)], piece sizes multiples of world * SHARD_ALIGN."""
        n_chunks = max(1, min(int(n_chunks), self.MAX_CHUNKS))
        unit = self.world * self.SHARD_ALIGN
        per = -(-max(n_gaussians, 1) // n_chunks)
        per = (per + unit - 1) // unit * unit
        out, r0, g0 = [], 0, 0
        while r0 < n_gaussians:
            n = min(per, n_gaussians - r0)
            ns = self.shard_rows(n, self.world)
            out.append((r0, n, ns, g0))
            r0 += n
            g0 += ns          # rows per source rank; the block of a piece holds world * ns rows
        return out

    GRAD_WIDTH = {"rest": None, "dc": 3, "means": 3, "scales": 3, "quats": 4, "logit": 1}

    def width(self, name: str) -> int:
        return (self.K - 1) * 3 if name == "rest" else self.GRAD_WIDTH[name]


class _DeviceSpan:
    """Exposes a raw device allocation through __cuda_array_interface__ (torch.as_tensor aliases it)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


class PeerGradExchange:
    """Gradient exchange over NVLink peer memory, done by the kernels themselves (csrc/peer.cu): no
    collective call inside the step.  Used through the fused autograd node like PackedGradExchange.

    Every rank owns one cudaMalloc allocation laid out by PeerLayout and maps every other rank's
    allocation through CUDA IPC (handles exchanged once per (re)allocation over the process group).
    Per backward: ts_dp_push (geometry rows -> owner, colour cotangents + camera -> everyone),
    ts_peer_barrier, SH-backward over all views (local), shard projection-backward with stores into
    every rank's gradient segment, ts_peer_barrier.  The gradients handed to autograd are views of
    the local segment: they are overwritten by the next backward (consume or copy them before)."""

    peer = True

    def __init__(self, process_group=None, average: bool = True, headroom: float = 0.125,
                 timeout_s: float = 20.0, n_chunks: Optional[int] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGradExchange needs an initialised process group")
        from . import _lib
        self._lib = _lib
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        if self.world > _lib.load().ts_peer_max_ranks():
            raise RuntimeError(f"PeerGradExchange supports up to {_lib.load().ts_peer_max_ranks()} ranks")
        self.average = average
        self.headroom = headroom
        self.timeout_s = timeout_s
        # pieces the rows are pushed in (the transfer of one overlaps the shard backward of the previous)
        # (measured on 2 and 8 B200s: 1, 2 and 4 pieces are within 1 % of each other — profiles/README.md)
        self.n_chunks = int(n_chunks if n_chunks is not None else os.environ.get("TINYSPLAT_B200_PEER_CHUNKS", "1"))
        # geometry rows pushed and signalled before the colour rows (the shard projection-backward runs under
        # the colour transfer): 8 GPUs 1.3332 vs 1.3357 ms, 2 GPUs 1.145 vs 1.130 ms (4 GPUs: not measured)
        # -> on at 8 ranks only
        if "TINYSPLAT_B200_PEER_SPLIT" not in os.environ:
            _lib.load().ts_dp_exchange_split(1 if self.world >= 8 else 0)
        self.layout: Optional[PeerLayout] = None
        self.device = None
        self._base = None            # my allocation
        self._peer_bases: List[int] = []
        self._opened: List[int] = []
        self._span = None            # torch uint8 view of my allocation
        self._tables = {}
        self.epoch = 0
        self.last_bytes_sent = 0

    def out_scale(self) -> float:
        return 1.0 / self.world if self.average else 1.0

    # -- allocation + rendezvous --------------------------------------------------------------
    def ensure(self, n_gaussians: int, K: int, device) -> PeerLayout:
        """(Re)allocates and re-maps when the capacity or K changes; every rank must call this with
        the same arguments in the same step (N changes only when the model densifies)."""
        L = self.layout
        if L is not None and L.K == K and n_gaussians <= L.cap_rows and device == self.device:
            return L
        import ctypes as C
        lib = self._lib.load()
        self.close()
        cap = int(n_gaussians * (1.0 + self.headroom)) + 1024
        L = PeerLayout(self.world, cap, K, lib.ts_peer_flag_bytes())
        with torch.cuda.device(device):
            base = C.c_void_p()
            self._lib.check(lib.ts_peer_alloc(L.total_bytes, C.byref(base)), "ts_peer_alloc")
            hbytes = lib.ts_peer_ipc_handle_bytes()
            hbuf = (C.c_ubyte * hbytes)()
            self._lib.check(lib.ts_peer_ipc_get(base, hbuf), "ts_peer_ipc_get")
            mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=device)
            every = torch.empty(self.world * hbytes, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(every, mine, group=self.group)
            every = every.cpu().view(self.world, hbytes)
            bases, opened = [], []
            for r in range(self.world):
                if r == self.rank:
                    bases.append(base.value)
                    continue
                h = (C.c_ubyte * hbytes)(*every[r].tolist())
                p = C.c_void_p()
                self._lib.check(lib.ts_peer_ipc_open(h, C.byref(p)), "ts_peer_ipc_open")
                bases.append(p.value)
                opened.append(p.value)
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)      # everybody mapped everybody before the first push
        self.layout, self.device = L, device
        self._base, self._peer_bases, self._opened = base.value, bases, opened
        self._span = torch.as_tensor(_DeviceSpan(base.value, L.total_bytes), device=device)
        self.epoch = 0
        self._tables = {}
        return L

    def _table(self, key, ptrs):
        import ctypes as C
        t = self._tables.get(key)
        if t is None:
            if len(self._tables) > 4096:      # offsets change with N (densification): do not grow for ever
                self._tables.clear()
            t = (C.c_void_p * len(ptrs))(*ptrs)
            self._tables[key] = t
        return t

    def seg_ptrs(self, name: str, row_offset_bytes: int = 0):
        """ctypes table: the address of segment `name` (+ offset) in every rank's allocation."""
        off = self.layout.seg[name][0] + row_offset_bytes
        return self._table((name, row_offset_bytes), [b + off for b in self._peer_bases])

    def local_ptr(self, name: str) -> int:
        return self._base + self.layout.seg[name][0]

    def local_view(self, name: str, n_rows: int, *shape) -> Tensor:
        """fp32 view [n_rows, *shape] of one of my gradient segments."""
        off, _ = self.layout.seg[name]
        numel = n_rows
        for d in shape:
            numel *= d
        return self._span[off:off + 4 * numel].view(torch.float32).view(n_rows, *shape)

    SEGMENTS = ("flags", "err", "cams", "geo", "rgb", "g_rest", "g_dc", "g_means", "g_scales", "g_quats", "g_logit")

    def bases_table(self):
        """ctypes table of every rank's allocation base (ts_dp_exchange_peer)."""
        return self._table("bases", self._peer_bases)

    def seg_offsets_table(self):
        import ctypes as C
        t = self._tables.get("seg_offsets")
        if t is None:
            t = (C.c_int64 * len(self.SEGMENTS))(*[self.layout.seg[n][0] for n in self.SEGMENTS])
            self._tables["seg_offsets"] = t
        return t

    def piece_plan(self, n_gaussians: int):
        """(ctypes int32 [pieces][4] = first row, rows, shard rows, first geometry row per source rank;
        number of pieces; bytes this rank sends over NVLink: geometry rows + colours + shard gradients)."""
        import ctypes as C
        key = ("plan", n_gaussians, self.n_chunks)
        hit = self._tables.get(key)
        if hit is None:
            pieces = self.layout.chunks(n_gaussians, self.n_chunks) if n_gaussians > 0 else []
            flat, sent, w = [], 0, self.world
            for r0, n, ns_c, g0 in pieces:
                flat += [r0, n, ns_c, g0]
                ns = max(0, min(n, (self.rank + 1) * ns_c) - self.rank * ns_c)
                sent += (n * 32 * (w - 1)) // w + n * 12 * (w - 1) + ns * 44 * (w - 1)
            hit = ((C.c_int32 * max(len(flat), 1))(*flat), len(flat), sent)
            if len(self._tables) > 4096:
                self._tables.clear()
            self._tables[key] = hit
        arr, nflat, sent = hit
        return arr, nflat // 4, sent

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    SIGNAL, WAIT = 1, 2

    def barrier(self, slot: int, epoch: int, stream_ptr: int, mode: int = 3) -> None:
        """mode: SIGNAL (what this stream did so far is visible to ranks that wait for `epoch`), WAIT
        (until every rank has signalled), or both.  The two halves may sit on different streams."""
        lib = self._lib.load()
        self._lib.check(lib.ts_peer_barrier(self.world, self.rank, self.seg_ptrs("flags"), slot, epoch,
                                            self.local_ptr("err"), float(self.timeout_s), int(mode), stream_ptr),
                        "ts_peer_barrier")

    def check(self) -> None:
        """Raises when a barrier timed out (a peer never arrived).  Synchronises the device."""
        if self._span is None:
            return
        off, _ = self.layout.seg["err"]
        code = int(self._span[off:off + 4].view(torch.int32).item())
        if code != 0:
            raise RuntimeError(f"peer barrier timed out waiting for rank {code - 1} (rank {self.rank})")

    def close(self) -> None:
        if self._base is None:
            return
        lib = self._lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)      # nobody is still writing into a buffer about to go away
            self._span = None
            for p in self._opened:
                lib.ts_peer_ipc_close(p)
            lib.ts_peer_free(self._base)
        self._base, self._peer_bases, self._opened, self.layout, self._tables = None, [], [], None, {}


class DataParallelRenderer:
    """rank r renders cameras[r::world] with `rasterizer`, backpropagates `loss_fn`, and leaves
    the view-averaged gradient in every parameter's .grad on every rank."""

    def __init__(self, rasterizer: Callable, params: Iterable[Tensor], process_group=None,
                 average: bool = True, overlap: bool = True, strategy: str = "allreduce"):
        """strategy: "allreduce" (any rasterizer), "packed" or "peer" (the fused pipeline of
        tinysplat_b200.rasterizer.GaussianRasterizer: gradients leave backward already reduced;
        "peer" moves the data with the kernels' own NVLink stores instead of NCCL calls)."""
        if strategy not in ("allreduce", "packed", "peer"):
            raise ValueError("strategy must be 'allreduce', 'packed' or 'peer'")
        self.rasterizer = rasterizer
        self.group = process_group
        self.strategy = strategy
        enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        if strategy in ("packed", "peer") and enabled:
            if getattr(rasterizer, "pipeline", None) != "fused":
                raise ValueError(f"strategy='{strategy}' needs GaussianRasterizer(pipeline='fused')")
            cls = PackedGradExchange if strategy == "packed" else PeerGradExchange
            rasterizer.grad_exchange = cls(process_group, average)
            self.reducer = GradientAllReducer([], process_group, average, overlap=False)
        else:
            self.reducer = GradientAllReducer(params, process_group, average, overlap)

    def step(self, camera, dims, sh_degree: int, loss_fn: Callable[[Tensor, dict], Tensor]):
        img, extras = self.rasterizer(camera, dims, sh_degree)
        loss = loss_fn(img, extras)
        loss.backward()
        self.reducer.finish()
        return loss.detach(), img.detach(), extras

    def render_only(self, camera, dims, sh_degree: int):
        with torch.no_grad():
            return self.rasterizer(camera, dims, sh_degree)

    def reduce_densify_stat(self, extras: dict) -> Optional[Tensor]:
        g = extras["xys"].grad
        if g is None:
            return None
        stat = g.norm(dim=-1)
        if self.reducer._enabled:
            dist.all_reduce(stat, op=dist.ReduceOp.SUM, group=self.group)
        return stat
