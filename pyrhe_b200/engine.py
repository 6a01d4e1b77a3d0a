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
from typing import Optional

import numpy as np
import torch

from . import _lib
from .assemble import PathPlan
from .hostmath import block_ranges, impute_uniforms, rhs_matrix


def _round_up(a, b):
    return (a + b - 1) // b * b


def shard_blocks(num_jack: int, world: int, rank: int):
    """Contiguous range [j0, j1) of jackknife blocks owned by `rank`: ceil(J / world) per rank, exactly
    how the reference hands block ranges to its worker processes (base.py:530-533)."""
    per = -(-num_jack // world)
    return min(rank * per, num_jack), min((rank + 1) * per, num_jack)


def allreduce_sum(tensors, group=None):
    """The path's one exchange step: sum the rank-local totals (and small Gram pieces) over all ranks.
    NCCL over NVLink for CUDA tensors; gloo in the CPU tests of the sharding logic."""
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class RheEngine:
    def __init__(self, plan: PathPlan, *, n_indv: int, keep: np.ndarray, annot: np.ndarray, num_jack: int,
                 impute: str = "binary", seed: int = 0, device: Optional[torch.device] = None,
                 kernel_path: Optional[int] = None, rank: int = 0, world: int = 1,
                 store_partials: bool = True, process_group=None):
        if not torch.cuda.is_available():
            raise _lib.RheError("pyrhe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.plan = plan
        if kernel_path is None:      # int8 tcgen05 kernels whenever the layout fits one TMEM allocation
            kernel_path = _lib.PATH_TCGEN05 if _lib.tcgen05_supported(plan) else _lib.PATH_SIMT
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
        self.j0, self.j1 = shard_blocks(self.J, world, rank)
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
        self.bed = None
        self._row_off = {}
        cur = 0
        for j in self.own:
            self._row_off[j] = cur
            cur += self.ranges[j][1] - self.ranges[j][0]
        self.m_own = cur
        self.nxe_S = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.rhe_ctx_destroy(self._ctx)
            self._ctx = None

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
        R, rowscale = rhs_matrix(plan, Z, W, Y_res, env, self.keep)
        Rp = np.zeros((R.shape[0], self.Np), dtype=np.float32)
        Rp[:, : self.n_indv] = R
        rs = np.zeros((plan.n_sets, self.Np), dtype=np.float32)
        rs[:, : self.n_indv] = rowscale
        bits = np.zeros(self.Np, dtype=np.uint32)
        bits[: self.n_indv] = self.keep.astype(np.uint32) * 3
        keep2 = np.zeros(self.Np // 16, dtype=np.uint32)
        for t in range(16):
            keep2 |= bits[t::16] << np.uint32(2 * t)
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

    def alloc_genotypes(self):
        """Zeroed device buffer for this rank's `.bed` rows, pitch padded to 128 bytes."""
        with torch.cuda.device(self.device):
            self.bed = torch.zeros((max(self.m_own, 1), self.pitch), dtype=torch.uint8, device=self.device)
        return self.bed

    def upload_block(self, j: int, host_rows, stream=None):
        """Host `.bed` rows of block j ([m_j, row_bytes] uint8 numpy / pinned torch) -> device."""
        a, b = self.ranges[j]
        m = b - a
        st = self._stream() if stream is None else C.c_void_p(stream.cuda_stream)
        dst = self.bed[self._row_off[j]]
        _lib.check(self.lib.rhe_upload_rows(_lib.ptr(host_rows), self.row_bytes, m, C.c_void_p(dst.data_ptr()),
                                            self.pitch, st))

    def load_genotypes(self, packed: np.ndarray):
        """packed [M, row_bytes] (numpy / memmap of the .bed payload): uploads this rank's blocks."""
        if self.bed is None:
            self.alloc_genotypes()
        for j in self.own:
            a, b = self.ranges[j]
            self.upload_block(j, np.ascontiguousarray(packed[a:b]))
        torch.cuda.current_stream(self.device).synchronize()

    def load_genotypes_async(self, packed: np.ndarray, n_workers: int = 4, ring: int = 3):
        """Pipelined ingest: `.bed` rows (numpy array or memmap) -> pinned staging ring -> device.

        A pool of `n_workers` threads copies row ranges of block j from `packed` (page cache / disk) into one of `ring`
        pinned slots (numpy copies release the GIL, so the reads fault in parallel); the uploader thread then queues
        one asynchronous 2-D copy per block on a side stream and records an event.  Staging of block j+1 therefore
        overlaps the PCIe copy of block j and the kernels of block j-1.  Pass the returned handle to
        `run(upload=...)`; block j is consumed as soon as its own copy has finished."""
        import threading
        from concurrent.futures import ThreadPoolExecutor
        if self.bed is None:
            self.alloc_genotypes()
        copy_stream = torch.cuda.Stream(self.device)
        ready = {j: threading.Event() for j in self.own}
        events = {}
        errors = []
        max_m = max((self.ranges[j][1] - self.ranges[j][0] for j in self.own), default=0)
        ring = max(1, min(ring, len(self.own)))
        n_workers = max(1, n_workers)

        def worker():
            try:
                with torch.cuda.device(self.device):
                    slots = [torch.empty((max_m, self.row_bytes), dtype=torch.uint8).pin_memory() for _ in range(ring)]
                    views = [sl.numpy() for sl in slots]
                    slot_free = [None] * ring                      # event of the last H2D copy that read the slot
                    with ThreadPoolExecutor(n_workers) as pool:
                        for n, j in enumerate(self.own):
                            a, b = self.ranges[j]
                            m = b - a
                            k = n % ring
                            if slot_free[k] is not None:
                                slot_free[k].synchronize()
                            step = -(-m // n_workers)
                            futs = [pool.submit(np.copyto, views[k][r0:min(m, r0 + step)], packed[a + r0:a + min(m, r0 + step)])
                                    for r0 in range(0, m, step)]
                            for f in futs:
                                f.result()
                            self.upload_block(j, slots[k][:m], stream=copy_stream)
                            ev = torch.cuda.Event()
                            ev.record(copy_stream)
                            events[j] = ev
                            slot_free[k] = ev
                            ready[j].set()
                    copy_stream.synchronize()                      # the pinned slots die with this thread
            except Exception as exc:          # surfaced by run()
                errors.append(exc)
                for e in ready.values():
                    e.set()

        thread = threading.Thread(target=worker, daemon=True)
        thread.start()
        return dict(ready=ready, events=events, errors=errors, thread=thread)

    def block_view(self, j: int):
        m = self.ranges[j][1] - self.ranges[j][0]
        return self.bed[self._row_off[j]: self._row_off[j] + m], m

    # ------------------------------------------------------------------ the path
    def _accumulate(self, j, P_out, S_accum, gram_out):
        rows, m = self.block_view(j)
        lo, _ = self._row_slices[j]
        _lib.check(self.lib.rhe_block_accumulate(
            self._ctx, C.c_void_p(rows.data_ptr()), m, C.c_void_p(self.bin_rows.data_ptr() + 4 * lo),
            self._offs[j], _lib.ptr(P_out), _lib.ptr(S_accum), _lib.ptr(gram_out), self._stream()))

    def run(self, upload_events=None, upload=None) -> dict:
        """All own blocks -> totals -> all-reduce -> leave-one-out Grams.

        Returns host arrays XX [J+1, E, E] and G_blk [J, E_reg, Rs, Rs] (identical on all ranks).
        `upload_events[j]`, when given, is a CUDA event the block's genotype upload signals; `upload` is the handle
        of `load_genotypes_async` (events appear as the background thread records them)."""
        plan = self.plan
        E, E_reg, B, Rs, Np, J = plan.E, plan.E_reg, plan.B, plan.Rs, self.Np, self.J
        dev = self.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            S = torch.zeros((E, B, Np), dtype=torch.float32, device=dev)
            G_blk = torch.zeros((J, E_reg, Rs, Rs), dtype=torch.float64, device=dev)
            XX = torch.zeros((J + 1, E, E), dtype=torch.float64, device=dev)
            P_all = None
            if self.store_partials:
                # every (estimate, column) row of a block partial is fully written by pass B; only the NxE row
                # (no genotype contribution) has to be zeroed
                P_all = torch.empty((max(len(self.own), 1), E, B, Np), dtype=torch.float32, device=dev)
                if E > E_reg:
                    P_all[:, E_reg:].zero_()
            for jl, j in enumerate(self.own):
                if upload is not None:
                    upload["ready"][j].wait()
                    if upload["errors"]:
                        raise upload["errors"][0]
                    cur.wait_event(upload["events"][j])
                if upload_events is not None:
                    cur.wait_event(upload_events[j])
                self._accumulate(j, P_all[jl] if self.store_partials else None, S, G_blk[j])
            if self.world > 1:
                allreduce_sum([S, G_blk], self.pg)
            if plan.has_nxe:
                S[E_reg].copy_(self.nxe_S)
            length = B * Np
            scratch = None
            if self.store_partials and len(self.own) > 0:
                # stored partials are contiguous, and so are the XX slots of this rank's (contiguous) blocks: up to
                # four blocks per launch share one read of S
                j0 = self.own[0]
                _lib.check(self.lib.rhe_loo_gram_multi(
                    self._ctx, _lib.ptr(S), _lib.ptr(P_all), E * B * Np, len(self.own), E, length,
                    _lib.ptr(XX[j0]), E * E, self._stream()))
            for jl, j in enumerate(self.own if not self.store_partials else []):
                if self.store_partials:
                    Pj = P_all[jl]
                else:                                               # streaming policy: recompute the block
                    if scratch is None:
                        scratch = torch.zeros((E, B, Np), dtype=torch.float32, device=dev)
                        gscratch = torch.zeros((E_reg, Rs, Rs), dtype=torch.float64, device=dev)
                    self._accumulate(j, scratch, None, gscratch)
                    Pj = scratch
                _lib.check(self.lib.rhe_loo_gram(self._ctx, _lib.ptr(S), _lib.ptr(Pj), E, length, _lib.ptr(XX[j]),
                                                 self._stream()))
            if self.rank == self.world - 1:
                _lib.check(self.lib.rhe_loo_gram(self._ctx, _lib.ptr(S), None, E, length, _lib.ptr(XX[J]),
                                                 self._stream()))
            if self.world > 1:
                allreduce_sum([XX], self.pg)
            self.S, self.P_all = S, P_all
            out = dict(XX=XX.cpu().numpy(), G_blk=G_blk.cpu().numpy(), M=self.Mjk)
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
    def decode_block(self, j: int, apply_impute: bool) -> np.ndarray:
        rows, m = self.block_view(j)
        out = torch.empty((m, self.Np), dtype=torch.int8, device=self.device)
        _lib.check(self.lib.rhe_decode_block(self._ctx, C.c_void_p(rows.data_ptr()), m, int(apply_impute),
                                             _lib.ptr(out), self._stream()))
        return out.cpu().numpy()[:, : self.n_indv]

    def block_stats(self, j: int) -> np.ndarray:
        rows, m = self.block_view(j)
        out = torch.empty((m, 4), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.rhe_block_stats(self._ctx, C.c_void_p(rows.data_ptr()), m, _lib.ptr(out), self._stream()))
        return out.cpu().numpy()
