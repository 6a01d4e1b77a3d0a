"""Device engine: drives libpyrhe_b200 over the jackknife blocks of one rank.

One process per GPU.  Jackknife blocks are split over ranks in contiguous ranges
(mirrors /root/reference/pyrhe/src/base/base.py:530-533), each rank keeps the
`.bed` rows of its own blocks resident in HBM, and there is exactly one exchange
step: an all-reduce (sum) of the running totals S = sum_j P_j and of the small
per-block Gram pieces.  Leave-one-out vectors are S - P_j formed
inside the Gram kernel (base.py:483-486 does the subtraction on the host).

PyTorch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np
import torch

from . import _lib
from .assemble import PathPlan
from .hostmath import block_ranges, impute_uniforms, rhs_matrix


def _round_up(a, b):
    return (a + b - 1) // b * b


def shard_sizes(num_jack: int, world: int, weights=None):
    """Blocks per rank.  Default: ceil(J / world) per rank, exactly how the reference hands block ranges to its worker
    processes (base.py:530-533).  With `weights` (one positive number per rank, e.g. the measured host-to-device rate
    of each rank's link when a pass is bound by the upload): shares proportional to the weights, largest remainders
    first, and at least one block per rank when there are enough blocks."""
    if weights is None:
        per = -(-num_jack // world)
        return [max(0, min((r + 1) * per, num_jack) - min(r * per, num_jack)) for r in range(world)]
    w = np.asarray(weights, dtype=np.float64)
    if w.shape != (world,) or not np.all(np.isfinite(w)) or np.any(w <= 0):
        raise ValueError("shard weights: one positive finite number per rank")
    raw = w / w.sum() * num_jack
    sizes = np.floor(raw).astype(np.int64)
    order = np.argsort(-(raw - sizes), kind="stable")
    sizes[order[: num_jack - int(sizes.sum())]] += 1
    while num_jack >= world and sizes.min() == 0:              # nobody idles while another rank holds several blocks
        sizes[int(np.argmax(sizes))] -= 1
        sizes[int(np.argmin(sizes))] += 1
    return [int(x) for x in sizes]


def shard_blocks(num_jack: int, world: int, rank: int, weights=None):
    """Contiguous range [j0, j1) of jackknife blocks owned by `rank` (`shard_sizes`)."""
    sizes = shard_sizes(num_jack, world, weights)
    j0 = sum(sizes[:rank])
    return j0, j0 + sizes[rank]


def allreduce_sum(tensors, group=None):
    """The path's one exchange step: sum the rank-local totals (and small Gram pieces) over all ranks.
    NCCL over NVLink for CUDA tensors; gloo in the CPU tests of the sharding logic."""
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class BlockStreamer:
    """Bounded-memory, pipelined `.bed` ingest of one rank (base.py:338-345 reads one block at a time).

    host rows (numpy array / memmap of the `.bed` payload) --staging threads--> pinned ring --copy stream--> device slot
    (+ the block's allele counts, `rhe_block_stats`, on the same stream).  A source that already lives in pinned host
    memory (it offers `pinned_rows(a, b)` -> pinned uint8 tensor [b - a, row_bytes]) is copied from where it lies,
    without the staging hop.  One background thread drives the staging
    and queues the copies, so the caller's thread only enqueues kernels: staging of block n+1 overlaps the PCIe copy
    of block n and the kernels of block n-1.  When the engine holds a ring of R device slots, the copy of block n
    waits (on the device) for the kernels of block n-R, so HBM use is bounded by R blocks whatever the size of the
    `.bed` share.  File-backed sources are read with `preadv` straight into the pinned ring (no page-fault walk over
    a mapping); in-memory arrays are copied by the same threads.
    """

    def __init__(self, eng: "RheEngine", packed: np.ndarray, n_workers: Optional[int] = None, ring_host: int = 3):
        import os
        self.eng = eng
        self.packed = packed
        cores = os.cpu_count() or 4
        self.n_workers = max(1, n_workers if n_workers else min(16, max(2, cores // max(eng.world, 1))))
        self.ring_host = max(1, min(ring_host, max(len(eng.own), 1)))
        self.copy_stream = torch.cuda.Stream(eng.device)
        # the ingest kernels of a block (allele counts, individual-major copy, re-tiling: ~2 ms at config 5) run on a
        # stream of their own behind the block's copy, so that the next block's H2D copy starts right away
        self.ingest_stream = torch.cuda.Stream(eng.device)
        # the genotype buffer was zeroed on the caller's stream: order that memset before the first copy
        self.copy_stream.wait_stream(torch.cuda.current_stream(eng.device))
        self.ingest_stream.wait_stream(torch.cuda.current_stream(eng.device))
        for st in (self.copy_stream, self.ingest_stream):
            eng.bed.record_stream(st)
            eng.counts.record_stream(st)
            for t in eng.gt.values():
                t.record_stream(st)
        self._fd, self._file_off = None, 0
        fname = getattr(packed, "filename", None)
        if fname is not None and isinstance(packed, np.memmap) and packed.flags.c_contiguous \
                and packed.shape[1] == eng.row_bytes:
            self._fd = os.open(fname, os.O_RDONLY)
            self._file_off = int(packed.offset)
        self.pinned_source = hasattr(packed, "pinned_rows")
        # the pinned ring: slot 0 now, the others on a helper thread while the first block is being staged (pinning
        # host memory runs at 1-2 GB/s: at config 5 three 1.25 GB slots would otherwise cost seconds up front)
        self.slots, self.views = [], []
        self._slot_ready = []
        self._slot_thread = None
        self.errors = []
        if not self.pinned_source:
            import threading
            self._slot_ready = [threading.Event() for _ in range(self.ring_host)]
            self.slots = [None] * self.ring_host
            self.views = [None] * self.ring_host

            def make_slots(ks):
                try:
                    with torch.cuda.device(eng.device):
                        for k in ks:
                            self.slots[k] = torch.empty((eng.max_m, eng.row_bytes), dtype=torch.uint8, pin_memory=True)
                            self.views[k] = self.slots[k].numpy()
                            self._slot_ready[k].set()
                except Exception as exc:                       # surfaced by the worker through check()
                    self.errors.append(exc)
                    for e in self._slot_ready:
                        e.set()
            make_slots([0])
            if self.ring_host > 1:
                self._slot_thread = threading.Thread(target=make_slots, args=(range(1, self.ring_host),), daemon=True)
                self._slot_thread.start()
        self.thread = None
        self.bytes_staged = 0
        self.passes = 0
        self.copy_seconds = None

    # ---- staging
    def _read_chunk(self, dst: np.ndarray, first_row: int):
        """Rows [first_row, first_row + len(dst)) of the payload into a contiguous piece of a pinned slot."""
        import os
        if self._fd is not None:
            mv = memoryview(dst).cast("B")
            off = self._file_off + first_row * self.eng.row_bytes
            done = 0
            while done < len(mv):
                n = os.preadv(self._fd, [mv[done:]], off + done)
                if n <= 0:
                    raise IOError("short read from the .bed file")
                done += n
        else:
            np.copyto(dst, self.packed[first_row: first_row + dst.shape[0]])

    def _worker(self, order):
        from concurrent.futures import ThreadPoolExecutor
        eng = self.eng
        R = eng.ring_blocks
        t_start = time.perf_counter()
        try:
            with torch.cuda.device(eng.device), ThreadPoolExecutor(self.n_workers) as pool:
                slot_free = [None] * self.ring_host            # event of the last H2D copy that read the pinned slot
                for n, j in enumerate(order):
                    a, b = eng.ranges[j]
                    m = b - a
                    k = n % self.ring_host
                    if self.pinned_source:
                        src = self.packed.pinned_rows(a, b)
                    else:
                        self._slot_ready[k].wait()
                        if self.errors:
                            return
                        if slot_free[k] is not None:
                            slot_free[k].synchronize()
                        step = -(-m // self.n_workers)
                        futs = [pool.submit(self._read_chunk, self.views[k][r0:min(m, r0 + step)], a + r0)
                                for r0 in range(0, m, step)]
                        for f in futs:
                            f.result()
                        src = self.slots[k][:m]
                    self.bytes_staged += m * eng.row_bytes
                    if R is not None and n >= R:               # the device slot still belongs to block n - R
                        prev = order[n - R]
                        self._released[prev].wait()
                        if self.errors:
                            return
                        self.copy_stream.wait_event(self._done[prev])
                    eng.upload_block(j, src, stream=self.copy_stream, count=False)
                    copied = torch.cuda.Event()
                    copied.record(self.copy_stream)
                    slot_free[k] = copied                      # the pinned slot is free once the copy has read it
                    self.ingest_stream.wait_event(copied)
                    eng.count_block(j, self.ingest_stream)
                    ev = torch.cuda.Event()
                    ev.record(self.ingest_stream)
                    self._events[j] = ev
                    self._ready[j].set()
                self.copy_stream.synchronize()
                #: host seconds from the start of the last pass to the arrival of its last block on the device
                self.copy_seconds = time.perf_counter() - t_start
                self.ingest_stream.synchronize()
        except Exception as exc:          # surfaced by acquire() / check()
            self.errors.append(exc)
            for e in self._ready.values():
                e.set()

    # ---- protocol used by RheEngine._pass
    def start(self, order):
        import threading
        self.join()
        self._ready = {j: threading.Event() for j in order}
        self._released = {j: threading.Event() for j in order}
        self._events, self._done = {}, {}
        self.passes += 1
        self.thread = threading.Thread(target=self._worker, args=(list(order),), daemon=True)
        self.thread.start()

    def acquire(self, j):
        """Blocks the host until block j's copy (and count) has been queued; returns the event it signals."""
        self._ready[j].wait()
        self.check()
        return self._events[j]

    def release(self, j, stream):
        ev = torch.cuda.Event()
        ev.record(stream)
        self._done[j] = ev
        self._released[j].set()

    def check(self):
        if self.errors:
            for e in getattr(self, "_released", {}).values():
                e.set()
            raise self.errors[0]

    def join(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None

    def close(self):
        import os
        for e in getattr(self, "_released", {}).values():
            e.set()
        self.join()
        if self._slot_thread is not None:
            self._slot_thread.join()
            self._slot_thread = None
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None
        self.slots, self.views = [], []


class RheEngine:
    def __init__(self, plan: PathPlan, *, n_indv: int, keep: np.ndarray, annot: np.ndarray, num_jack: int,
                 impute: str = "binary", seed: int = 0, device: Optional[torch.device] = None,
                 kernel_path: Optional[int] = None, rank: int = 0, world: int = 1,
                 store_partials: bool = True, process_group=None, retile: bool = False, shard_weights=None):
        if not torch.cuda.is_available():
            raise _lib.RheError("pyrhe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.plan = plan
        if kernel_path is None:      # int8 tcgen05 kernels whenever the layout fits them (asked of the library)
            kernel_path = _lib.PATH_TCGEN05
            why = _lib.tcgen05_unsupported_reason(plan)
            if why:
                import warnings
                warnings.warn("pyrhe_b200: falling back to the CUDA-core kernels (about 50x slower than the tcgen05 "
                              f"path): {why}", RuntimeWarning, stacklevel=2)
                kernel_path = _lib.PATH_SIMT
        self.kernel_path = kernel_path
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rank, self.world, self.pg = rank, world, process_group
        self.store_partials = store_partials
        self.n_indv = int(n_indv)
        self.keep = np.asarray(keep, dtype=bool)
        self.n_kept = int(self.keep.sum())
        self.J = int(num_jack)
        self.M_snps, K = annot.shape
        assert K == plan.K
        self.row_bytes = (self.n_indv + 3) // 4
        self.pitch = _round_up(self.row_bytes, 128)
        self.Np = 4 * self.pitch
        self.ranges = block_ranges(self.M_snps, self.J)
        self.shard_weights = None if shard_weights is None else [float(x) for x in shard_weights]
        self.j0, self.j1 = shard_blocks(self.J, world, rank, self.shard_weights)
        self.own = list(range(self.j0, self.j1))
        self.max_m = max(b - a for a, b in self.ranges)

        # --- per-block bin row lists (base.py:315-336) and the M table (rhe.py:16)
        member = annot != 0
        E, E_reg = plan.E, plan.E_reg
        self.Mjk = np.zeros((self.J + 1, E), dtype=np.int64)
        self.Mjk[self.J, :E_reg] = np.tile((annot == 1).sum(axis=0), plan.n_groups)
        for j, (a, b) in enumerate(self.ranges):
            self.Mjk[j, :E_reg] = self.Mjk[self.J, :E_reg] - np.tile(member[a:b].sum(axis=0), plan.n_groups)
        if plan.has_nxe:
            self.Mjk[:, E_reg] = 1                                  # genie.py:79-82
        rows, self._offs, self._row_slices = [], {}, {}
        cursor = 0
        for j in self.own:
            a, b = self.ranges[j]
            lists = [np.nonzero(member[a:b, k])[0].astype(np.int32) for k in range(K)]
            off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int32)
            self._offs[j] = (C.c_int32 * (K + 1))(*off.tolist())
            n = int(off[-1])
            self._row_slices[j] = (cursor, cursor + n)
            rows.append(np.concatenate(lists) if n else np.zeros(0, np.int32))
            cursor += n
        with torch.cuda.device(self.device):
            all_rows = np.concatenate(rows) if rows else np.zeros(0, np.int32)
            self.bin_rows = torch.from_numpy(np.concatenate([all_rows, np.zeros(1, np.int32)])).to(self.device)

            cfg = _lib.RheConfig(device=self.device.index, n_indv=self.n_indv, n_kept=self.n_kept,
                                 pitch_bytes=self.pitch, n_cols_set=plan.Rs, n_sets=plan.n_sets, n_ops=plan.n_ops,
                                 n_vec=plan.B, n_bins=plan.K, max_block_snps=self.max_m,
                                 impute_binary=1 if impute == "binary" else 0, kernel_path=kernel_path)
            self._ctx = C.c_void_p()
            _lib.check(self.lib.rhe_ctx_create(C.byref(self._ctx), C.byref(cfg)))
            if impute == "binary":
                self.uniforms = torch.from_numpy(impute_uniforms(seed, self.max_m)).to(self.device)
                _lib.check(self.lib.rhe_set_uniforms(self._ctx, _lib.ptr(self.uniforms), self.max_m))
            # annotation metadata of every own block, resident before the first block is seen: the per-block call
            # neither allocates nor synchronises (include/pyrhe_b200.h)
            self._plans = {}
            for j in self.own:
                lo, _ = self._row_slices[j]
                handle = C.c_void_p()
                _lib.check(self.lib.rhe_block_plan_create(
                    self._ctx, self.ranges[j][1] - self.ranges[j][0], C.c_void_p(self.bin_rows.data_ptr() + 4 * lo),
                    self._offs[j], self._stream(), C.byref(handle)))
                self._plans[j] = handle
        self.bed = None
        self.counts = None
        self.ring_blocks = None
        self._counted = set()
        #: hand the ingest-time allele counts to rhe_block_accumulate (False: every call re-counts the block)
        self.use_resident_counts = True
        #: use the individual-major copies where they exist (False: pass B always gathers SNP-major rows)
        self.use_fast_layout = True
        self.gt = {}
        #: re-tile the SNP-major rows of every block that owns an individual-major copy once both ingest products are
        #: taken (`rhe_block_retile`: contiguous 16 KB boxes of imputed counts for pass A).  Such a block can then be
        #: read by `run()` only: no decode / recount / gather-kernel toggles (the models switch this on, the test hooks
        #: of the engine leave it off)
        self.retile = bool(retile) and kernel_path == _lib.PATH_TCGEN05
        self._tiled = set()
        self._retile_scratch = None
        # pinned twins of the small results of a pass and their copy stream, created before any large device allocation:
        # pinning host memory for the first time in a process costs tens of milliseconds (measured: + 50 ms on the first
        # pass when it happened there, next to 180 GB of mapped HBM)
        with torch.cuda.device(self.device):
            self._d2h_stream = torch.cuda.Stream(self.device)
            self._pinned = {name: torch.empty(shape, dtype=torch.float64, pin_memory=True)
                            for name, shape in (("G_blk", (self.J, plan.E_reg, plan.Rs, plan.Rs)),
                                                ("XX", (self.J + 1, plan.E, plan.E)))}
        self.tail_seconds = None
        self._keep2_host = None
        #: with stored partials: S = sum_j P_j in one pass after the blocks (`rhe_sum_partials`: totals that are
        #: bit-reproducible from run to run) instead of RED into S from every pass B.  Off by default: at 13 config-5
        #: blocks per rank the extra pass costs 0.35 ms and the RED-free pass B saves 0.01 ms per block
        self.sum_stored_partials = False
        self.S = self.P_all = None
        self._row_off = {}
        cur = 0
        for j in self.own:
            self._row_off[j] = cur
            cur += self.ranges[j][1] - self.ranges[j][0]
        self.m_own = cur
        self._slot_off = dict(self._row_off)
        self.nxe_S = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_ctx", None):
            torch.cuda.synchronize(self.device)
            for handle in getattr(self, "_plans", {}).values():
                self.lib.rhe_block_plan_destroy(self._ctx, handle)
            self._plans = {}
            self.lib.rhe_ctx_destroy(self._ctx)
            self._ctx = None
            # the genotype residency and the accumulators go back to the allocator with the context
            self.bed = self.counts = self.S = self.P_all = self._retile_scratch = None
            self.gt = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        return int(self.lib.rhe_launch_count(self._ctx))

    # ------------------------------------------------------------------ inputs
    def set_rhs(self, Z: np.ndarray, W, Y_res: np.ndarray, env=None):
        """Right-hand sides [Z | W | y_res] (+ env-scaled set) -> fp32 on the device (mat_mul.py:12)."""
        plan = self.plan
        # fp32, padded to the device row length, in one pass over the inputs
        Rp, rs = rhs_matrix(plan, Z, W, Y_res, env, self.keep, dtype=np.float32, width=self.Np)
        if self._keep2_host is None:                           # 2-bit keep mask per packed word (the keep set is fixed)
            bits = np.zeros(self.Np, dtype=np.uint32)
            bits[: self.n_indv] = self.keep.astype(np.uint32) * 3
            keep2 = np.zeros(self.Np // 16, dtype=np.uint32)
            for t in range(16):
                keep2 |= bits[t::16] << np.uint32(2 * t)
            self._keep2_host = keep2
        keep2 = self._keep2_host
        with torch.cuda.device(self.device):
            self.R = torch.from_numpy(Rp).to(self.device)
            self.rowscale = torch.from_numpy(rs).to(self.device)
            self.keep2 = torch.from_numpy(keep2.view(np.int32)).to(self.device)
            _lib.check(self.lib.rhe_set_rhs(self._ctx, _lib.ptr(self.R), _lib.ptr(self.rowscale),
                                            _lib.ptr(self.keep2), self._stream()))
            if plan.has_nxe:                                         # X = diag(env): XXz = env^2 * z
                full = np.zeros((plan.B, self.Np), dtype=np.float32)
                full[:, np.nonzero(self.keep)[0]] = ((np.asarray(env, np.float64) ** 2)[:, None] * Z).T
                self.nxe_S = torch.from_numpy(full).to(self.device)

    # ------------------------------------------------------------------ genotype residency
    def alloc_genotypes(self, ring_blocks: Optional[int] = None, fast_layout: bool = True, reserve_bytes: float = 6e9):
        """Device buffer for `.bed` rows, pitch padded to 128 bytes, zeroed (the padding stays zero).

        ring_blocks=None: every own block resident (`m_own x pitch` bytes).  ring_blocks=R: a ring of R block-sized
        slots for the bounded-memory ingest (`stream_genotypes`), the way the reference holds one block at a time
        (/root/reference/pyrhe/src/base/base.py:338-345, base_streaming.py:85-104): block n of the streaming order
        lives in slot n % R until its kernels have run.  Next to the rows sit the per-SNP allele counts
        (`rhe_block_stats`), filled when a block becomes resident."""
        with torch.cuda.device(self.device):
            self._tiled = set()
            self._retile_scratch = None
            if ring_blocks is None and fast_layout and self.retile:
                # every block starts on a 128-row boundary and owns whole 128-row tiles (rhe_block_retile)
                self.ring_blocks = None
                self._slot_off, rows = {}, 0
                for j in self.own:
                    self._slot_off[j] = rows
                    rows += -(-(self.ranges[j][1] - self.ranges[j][0]) // 128) * 128
                rows = max(rows, 1)
            elif ring_blocks is None:
                self.ring_blocks = None
                rows = max(self.m_own, 1)
                self._slot_off = dict(self._row_off)
            else:
                self.ring_blocks = R = max(1, min(int(ring_blocks), max(len(self.own), 1)))
                rows = R * self.max_m
                self._slot_off = {j: (n % R) * self.max_m for n, j in enumerate(self.own)}
            self.bed = torch.zeros((rows, self.pitch), dtype=torch.uint8, device=self.device)
            self.counts = torch.zeros((rows, 4), dtype=torch.int32, device=self.device)
            self._counted = set()
            self.gt = {}
            self.reserve_state()
            if self.world > 1 and not getattr(self, "_collective_warm", False):
                # the first all-reduce of a communicator sets up its channels and staging buffers: pay for that here, on
                # the buffers of the real exchange, so that the first pass over the blocks runs like every later one
                self.S.zero_()
                plan = self.plan
                allreduce_sum([self.S, torch.zeros((self.J, plan.E_reg, plan.Rs, plan.Rs), dtype=torch.float64,
                                                   device=self.device)], self.pg)
                self._collective_warm = True
            if ring_blocks is None and fast_layout:
                self._alloc_fast_layout(reserve_bytes)
                if self.retile and self.gt and int(self.lib.rhe_block_tiled_bytes(self._ctx, self._plans[self.own[0]])) > 0:
                    self._retile_scratch = torch.empty(-(-self.max_m // 128) * 128 * self.pitch, dtype=torch.uint8,
                                                       device=self.device)
        return self.bed

    def _alloc_fast_layout(self, reserve_bytes: float):
        """Individual-major copies (`rhe_block_transpose`) for as many resident blocks as the HBM holds next to the
        SNP-major rows, the stored partials and the totals: pass B of those blocks feeds the tensor cores from tensor
        memory instead of staging every genotype through shared memory.  A block without a copy simply
        takes the gather kernel; results are identical either way."""
        free_b, _ = torch.cuda.mem_get_info(self.device)
        state = self.plan.E * self.plan.B * self.Np * 4
        budget = free_b - reserve_bytes - 3 * state      # S and the stored partials are allocated already (reserve_state)
        if self.retile:
            budget -= -(-self.max_m // 128) * 128 * self.pitch      # the scratch block of rhe_block_retile
        for j in self.own:
            need = int(self.lib.rhe_block_fast_bytes(self._ctx, self._plans[j]))
            if need <= 0 or need > budget:
                break
            self.gt[j] = torch.empty(need, dtype=torch.uint8, device=self.device)
            budget -= need

    def reserve_state(self):
        """The accumulators of a pass -- totals `S [E, B, Np]` and, with stored partials, `P_all [own blocks, E, B, Np]`
        (16 GB at config 5 on one GPU) -- allocated once and reused by every `run()`; `alloc_genotypes` calls this
        before it sizes the individual-major copies, so the first pass of a model run allocates nothing."""
        plan = self.plan
        with torch.cuda.device(self.device):
            if self.S is None:
                self.S = torch.empty((plan.E, plan.B, self.Np), dtype=torch.float32, device=self.device)
            if self.store_partials and self.P_all is None:
                self.P_all = torch.empty((max(len(self.own), 1), plan.E, plan.B, self.Np), dtype=torch.float32,
                                         device=self.device)
        return self.S, (self.P_all if self.store_partials else None)

    def genotype_bytes(self) -> int:
        if self.bed is None:
            return 0
        return self.bed.numel() + self.counts.numel() * 4 + sum(t.numel() for t in self.gt.values())

    def block_view(self, j: int):
        m = self.ranges[j][1] - self.ranges[j][0]
        off = self._slot_off[j]
        return self.bed[off: off + m], m

    def count_block(self, j: int, stream=None):
        """Per-SNP allele counts of resident block j (`rhe_block_stats`): the one read of the block that depends on
        nothing but the genotypes, done when the block lands so that no later pass repeats it."""
        if j in self._tiled:                           # counts and copies of a re-tiled block are final
            return
        rows, m = self.block_view(j)
        st = self._stream() if stream is None else C.c_void_p(stream.cuda_stream)
        cnt = self.counts[self._slot_off[j]: self._slot_off[j] + m]
        _lib.check(self.lib.rhe_block_stats(self._ctx, C.c_void_p(rows.data_ptr()), m, _lib.ptr(cnt), st))
        if j in self.gt:                               # ingest also writes the block's individual-major copy
            _lib.check(self.lib.rhe_block_transpose(self._ctx, C.c_void_p(rows.data_ptr()), self._plans[j], _lib.ptr(cnt),
                                                    _lib.ptr(self.gt[j]), st))
            if self._retile_scratch is not None:       # ... and, last, re-tiles the rows themselves for pass A
                if stream is not None:
                    self._retile_scratch.record_stream(stream)
                _lib.check(self.lib.rhe_block_retile(self._ctx, C.c_void_p(rows.data_ptr()), self._plans[j], _lib.ptr(cnt),
                                                     _lib.ptr(self._retile_scratch), st))
                self._tiled.add(j)
        self._counted.add(j)

    def count_all(self):
        for j in self.own:
            self.count_block(j)

    def upload_block(self, j: int, host_rows, stream=None, count: bool = True):
        """Host `.bed` rows of block j ([m_j, row_bytes] uint8 numpy / pinned torch) -> its device slot (+ counts)."""
        a, b = self.ranges[j]
        m = b - a
        st = self._stream() if stream is None else C.c_void_p(stream.cuda_stream)
        self._tiled.discard(j)                         # fresh PLINK rows
        dst = self.bed[self._slot_off[j]]
        _lib.check(self.lib.rhe_upload_rows(_lib.ptr(host_rows), self.row_bytes, m, C.c_void_p(dst.data_ptr()),
                                            self.pitch, st))
        if count:
            self.count_block(j, stream)

    def load_genotypes(self, packed: np.ndarray):
        """packed [M, row_bytes] (numpy / memmap of the .bed payload): uploads this rank's blocks (synchronous)."""
        if self.bed is None or self.ring_blocks is not None:
            self.alloc_genotypes()
        for j in self.own:
            a, b = self.ranges[j]
            self.upload_block(j, np.ascontiguousarray(packed[a:b]))
        torch.cuda.current_stream(self.device).synchronize()

    def plan_residency(self, reserve_bytes: float = 8e9) -> Optional[int]:
        """None when this rank's whole `.bed` share (plus the stored partials, if kept) fits the free HBM, else the
        number of ring slots to stream through."""
        free_b, _ = torch.cuda.mem_get_info(self.device)
        part = len(self.own) * self.plan.E * self.plan.B * self.Np * 4 if self.store_partials else 0
        need = self.m_own * (self.pitch + 16) + part + 3 * self.plan.E * self.plan.B * self.Np * 4
        return None if need + reserve_bytes < free_b else 4

    def stream_genotypes(self, packed: np.ndarray, n_workers: Optional[int] = None, ring_host: int = 3,
                         ring_blocks="auto"):
        """Pipelined ingest: `.bed` rows (numpy array or memmap) -> pinned staging ring -> device.
        Returns the `BlockStreamer` to hand to `run(upload=...)`."""
        if ring_blocks == "auto":
            ring_blocks = self.plan_residency()
        if self.bed is None or self.ring_blocks != ring_blocks:
            self.alloc_genotypes(ring_blocks)
        return BlockStreamer(self, packed, n_workers=n_workers, ring_host=ring_host)

    load_genotypes_async = stream_genotypes          # round-1 name

    # ------------------------------------------------------------------ the path
    def _accumulate(self, j, P_out, S_accum, gram_out):
        rows, m = self.block_view(j)
        cnt = gt = None
        layout = _lib.ROWS_PLINK
        if j in self._tiled:
            if not (self.use_resident_counts and self.use_fast_layout):
                raise _lib.RheError(f"block {j} was re-tiled at ingest (retile=True): its rows serve pass A only, the "
                                    "recount / gather-kernel toggles need an engine built with retile=False")
            layout = _lib.ROWS_TILED
        if j in self._counted and self.use_resident_counts:
            cnt = C.c_void_p(self.counts.data_ptr() + 16 * self._slot_off[j])
            if self.use_fast_layout and j in self.gt:
                gt = _lib.ptr(self.gt[j])
        _lib.check(self.lib.rhe_block_accumulate(
            self._ctx, C.c_void_p(rows.data_ptr()), self._plans[j], cnt, gt, layout, _lib.ptr(P_out), _lib.ptr(S_accum),
            _lib.ptr(gram_out), self._stream()))

    def _pass(self, upload, body):
        """One sweep over the own blocks in order; with a streamer, block j is consumed as soon as its own copy has
        landed and its device slot is handed back right after its kernels are queued."""
        cur = torch.cuda.current_stream(self.device)
        if upload is not None:
            upload.start(self.own)
        for jl, j in enumerate(self.own):
            if upload is not None:
                cur.wait_event(upload.acquire(j))
            body(jl, j)
            if upload is not None:
                upload.release(j, cur)

    def _host_buf(self, name: str, like):
        """Pinned host twin of a small device result (allocated once per engine): its D2H copy is asynchronous."""
        buf = self._pinned.get(name)
        if buf is None or buf.shape != like.shape:
            buf = self._pinned[name] = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
        return buf

    def run(self, upload=None, gram_hook=None) -> dict:
        """All own blocks -> totals -> all-reduce -> leave-one-out Grams.

        Returns host arrays XX [J+1, E, E] and G_blk [J, E_reg, Rs, Rs] (identical on all ranks).  `upload` is the
        `BlockStreamer` of `stream_genotypes`; without it the blocks must already be resident (`load_genotypes`).
        `gram_hook(G_blk)` (optional) runs on the host as soon as the per-bin Gram pieces have arrived, while the
        device still reduces `S` and forms the leave-one-out Grams (the covariate terms of the normal equations need
        nothing else: `assemble.normal_equations_prepare`); its return value comes back as `out["gram_hook"]`."""
        plan = self.plan
        E, E_reg, B, Rs, Np, J = plan.E, plan.E_reg, plan.B, plan.Rs, self.Np, self.J
        dev = self.device
        if upload is None and self.ring_blocks is not None and self.ring_blocks < len(self.own):
            raise _lib.RheError("genotypes live in a ring: run() needs the streamer (upload=...)")
        with torch.cuda.device(dev):
            S, P_all = self.reserve_state()            # the same buffers every run: no allocation inside a pass
            S.zero_()
            G_blk = torch.zeros((J, E_reg, Rs, Rs), dtype=torch.float64, device=dev)
            XX = torch.zeros((J + 1, E, E), dtype=torch.float64, device=dev)
            if P_all is not None and E > E_reg:
                # every (estimate, column) row of a block partial is fully written by pass B; only the NxE row
                # (no genotype contribution) has to be zeroed
                P_all[:, E_reg:].zero_()
            if self.store_partials and self.sum_stored_partials and len(self.own) > 0:
                # the totals from the stored partials in one streaming pass (block order: reproducible), instead of a
                # read-modify-write of S inside every block's pass B
                self._pass(upload, lambda jl, j: self._accumulate(j, P_all[jl], None, G_blk[j]))
                _lib.check(self.lib.rhe_sum_partials(self._ctx, _lib.ptr(P_all), E * B * Np, len(self.own), E_reg * B * Np,
                                                     _lib.ptr(S), self._stream()))
            else:
                self._pass(upload, lambda jl, j: self._accumulate(j, P_all[jl] if self.store_partials else None, S, G_blk[j]))
            if self.world > 1:
                allreduce_sum([G_blk], self.pg)
            # the small Gram pieces leave for the host first (pinned, asynchronous): the host works on them while the
            # device reduces S and forms the leave-one-out Grams
            st = torch.cuda.current_stream(dev)
            G_host = self._host_buf("G_blk", G_blk)
            self._d2h_stream.wait_stream(st)
            with torch.cuda.stream(self._d2h_stream):          # a side stream: the kernels that follow do not wait for it
                G_host.copy_(G_blk, non_blocking=True)
                ev_gram = torch.cuda.Event()
                ev_gram.record(self._d2h_stream)
            if self.world > 1:
                allreduce_sum([S], self.pg)
            if plan.has_nxe:
                S[E_reg].copy_(self.nxe_S)
            length = B * Np
            if self.store_partials and len(self.own) > 0:
                # stored partials are contiguous, and so are the XX slots of this rank's (contiguous) blocks: up to
                # four blocks per launch share one read of S
                j0 = self.own[0]
                _lib.check(self.lib.rhe_loo_gram_multi(
                    self._ctx, _lib.ptr(S), _lib.ptr(P_all), E * B * Np, len(self.own), E, length,
                    _lib.ptr(XX[j0]), E * E, self._stream()))
            elif len(self.own) > 0:
                # streaming policy (base_streaming.py:110-144): a second sweep recomputes every block's partial and
                # forms its leave-one-out Gram at once; in ring mode the rows are streamed from the host again
                scratch = torch.zeros((E, B, Np), dtype=torch.float32, device=dev)
                gscratch = torch.zeros((E_reg, Rs, Rs), dtype=torch.float64, device=dev)

                def second(jl, j):
                    self._accumulate(j, scratch, None, gscratch)
                    _lib.check(self.lib.rhe_loo_gram(self._ctx, _lib.ptr(S), _lib.ptr(scratch), E, length,
                                                     _lib.ptr(XX[j]), self._stream()))
                resident = self.ring_blocks is None or self.ring_blocks >= len(self.own)
                self._pass(None if resident else upload, second)
            if self.rank == self.world - 1:
                _lib.check(self.lib.rhe_loo_gram(self._ctx, _lib.ptr(S), None, E, length, _lib.ptr(XX[J]),
                                                 self._stream()))
            if self.world > 1:
                allreduce_sum([XX], self.pg)
            self.S, self.P_all = S, P_all
            XX_host = self._host_buf("XX", XX)
            XX_host.copy_(XX, non_blocking=True)
            ev_xx = torch.cuda.Event()
            ev_xx.record(st)
            t0 = time.perf_counter()
            ev_gram.synchronize()
            t1 = time.perf_counter()
            out = dict(G_blk=G_host.numpy().copy(), M=self.Mjk)
            if gram_hook is not None:
                out["gram_hook"] = gram_hook(out["G_blk"])
            t2 = time.perf_counter()
            ev_xx.synchronize()
            st.wait_stream(self._d2h_stream)                   # G_blk is reused by the next run
            out["XX"] = XX_host.numpy().copy()
            #: host seconds of the last run's tail: waiting for the Gram pieces, inside gram_hook, waiting for XX
            self.tail_seconds = (t1 - t0, t2 - t1, time.perf_counter() - t2)
            if upload is not None:
                upload.check()
        return out

    def decode_rows(self, packed_rows: np.ndarray) -> np.ndarray:
        """Decode arbitrary `.bed` rows ([m, row_bytes] uint8) -> int8 [m, N0], 3 = missing (no imputation)."""
        m = packed_rows.shape[0]
        out = np.empty((m, self.n_indv), dtype=np.int8)
        with torch.cuda.device(self.device):
            for a in range(0, m, self.max_m):
                b = min(m, a + self.max_m)
                dev = torch.zeros((b - a, self.pitch), dtype=torch.uint8, device=self.device)
                _lib.check(self.lib.rhe_upload_rows(_lib.ptr(packed_rows[a:b]), self.row_bytes, b - a,
                                                    C.c_void_p(dev.data_ptr()), self.pitch, self._stream()))
                dec = torch.empty((b - a, self.Np), dtype=torch.int8, device=self.device)
                _lib.check(self.lib.rhe_decode_block(self._ctx, C.c_void_p(dev.data_ptr()), b - a, 0, _lib.ptr(dec),
                                                     self._stream()))
                out[a:b] = dec.cpu().numpy()[:, : self.n_indv]
        return out

    # ------------------------------------------------------------------ test hooks
    def _plink_rows(self, j: int):
        if j in self._tiled:
            raise _lib.RheError(f"block {j} was re-tiled at ingest (retile=True): its PLINK rows are gone")
        return self.block_view(j)

    def decode_block(self, j: int, apply_impute: bool) -> np.ndarray:
        rows, m = self._plink_rows(j)
        out = torch.empty((m, self.Np), dtype=torch.int8, device=self.device)
        _lib.check(self.lib.rhe_decode_block(self._ctx, C.c_void_p(rows.data_ptr()), m, int(apply_impute),
                                             _lib.ptr(out), self._stream()))
        return out.cpu().numpy()[:, : self.n_indv]

    def block_stats(self, j: int) -> np.ndarray:
        rows, m = self._plink_rows(j)
        out = torch.empty((m, 4), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.rhe_block_stats(self._ctx, C.c_void_p(rows.data_ptr()), m, _lib.ptr(out), self._stream()))
        return out.cpu().numpy()
