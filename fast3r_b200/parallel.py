"""Sequence-parallel inference of the fusion decoder over the GPUs of one box (one process per GPU).

The reference has no counterpart (no TP/PP/SP anywhere); the single-device result is the oracle.
Partitioning: rank r owns a contiguous range of views, i.e. a contiguous token range of the
N*P-token sequence.  Encoder blocks, LayerNorm, all linears and the DPT heads are token/view-local and need no
communication; only the global attention couples ranks: each decoder layer all-gathers K|V (bf16, S_local x 2D)
over NCCL/NVLink and every rank attends its local queries against all keys.  The image-index ids drawn by rank 0 are
broadcast to all ranks so the result equals the single-device forward whatever the per-rank RNG states are.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_views(num_views: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced view ranges: the first (num_views % world) ranks get one extra view."""
    base, rem = divmod(num_views, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def assemble_kv(gathered: torch.Tensor, batch: int, rows: List[int]) -> torch.Tensor:
    """gathered: (world, batch * max_rows, C) padded per-rank K|V blocks, each laid out (b, s_local).
    Returns (batch * sum(rows), C) laid out (b, s_global) with ranks concatenated in order."""
    world, _, C = gathered.shape
    mx = max(rows)
    if batch == 1 and all(r == mx for r in rows):
        return gathered.reshape(world * mx, C)
    g = gathered.view(world, batch, mx, C)
    parts = [g[r, :, : rows[r]] for r in range(world)]  # (batch, rows_r, C)
    return torch.cat(parts, dim=1).reshape(-1, C)


class SequenceParallel:
    def __init__(self, group=None, gather_preds: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (torchrun) before enabling sequence parallel")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.gather_preds = gather_preds
        self._ranges = None
        self.bytes_exchanged = 0

    def view_range(self, num_views: int) -> Tuple[int, int]:
        self._ranges = shard_views(num_views, self.world)
        if any(hi - lo == 0 for lo, hi in self._ranges):
            raise ValueError(f"sequence parallel needs at least one view per rank ({num_views} views, {self.world} ranks)")
        return self._ranges[self.rank]

    def broadcast_ids(self, ids: torch.Tensor, device) -> torch.Tensor:
        """Every rank uses the image ids drawn by rank 0 of the group (the single-device stream), whatever its own CPU
        RNG state is; each rank has still consumed its own draw, like the reference does per forward."""
        if self.world == 1:
            return ids
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        on_gpu = dist.get_backend(self.group) == "nccl"
        t = ids.to(device) if on_gpu else ids.clone()
        dist.broadcast(t, src=src, group=self.group)
        return t.cpu()

    def make_kv_exchange(self, batch: int, s_local: int, dim: int):
        """Returns kv_exchange(kv_local (batch*s_local, 2*dim) bf16) -> (kv_all (batch*s_total, 2*dim), s_total)."""
        tok_per_view = s_local // (self._ranges[self.rank][1] - self._ranges[self.rank][0])
        rows = [(hi - lo) * tok_per_view for lo, hi in self._ranges]
        mx, s_total = max(rows), sum(rows)
        even = all(r == mx for r in rows)
        state = {}

        def kv_exchange(kv_local: torch.Tensor):
            C = kv_local.shape[-1]
            if "buf" not in state:
                state["buf"] = torch.empty(self.world, batch * mx, C, dtype=kv_local.dtype, device=kv_local.device)
                state["pad"] = None if even else torch.zeros(batch * mx, C, dtype=kv_local.dtype,
                                                             device=kv_local.device)
            src = kv_local
            if not even:
                pad = state["pad"]
                pad.view(batch, mx, C)[:, :s_local] = kv_local.view(batch, s_local, C)
                src = pad
            dist.all_gather_into_tensor(state["buf"].view(-1, C), src.contiguous(), group=self.group)
            self.bytes_exchanged += state["buf"].numel() * state["buf"].element_size()
            return assemble_kv(state["buf"], batch, rows), s_total

        return kv_exchange

    def gather_results(self, final_results, num_views, batch, H, W, device):
        """All ranks end up with the preds of every view (API parity with the single-device forward)."""
        keys = [k for k in ("pts3d_in_other_view", "conf", "pts3d_local", "conf_local")
                if k in final_results[self._ranges[self.rank][0]]]
        mxv = max(hi - lo for lo, hi in self._ranges)
        lo, hi = self._ranges[self.rank]
        out = [dict() for _ in range(num_views)]
        for k in keys:
            loc = torch.cat([final_results[i][k] for i in range(lo, hi)], dim=0)  # (n_loc*B, ...)
            tail = loc.shape[1:]
            send = torch.zeros((mxv * batch,) + tuple(tail), dtype=loc.dtype, device=device)
            send[: loc.shape[0]] = loc
            recv = torch.empty((self.world, mxv * batch) + tuple(tail), dtype=loc.dtype, device=device)
            dist.all_gather_into_tensor(recv.view((-1,) + tuple(tail)), send, group=self.group)
            for r, (a, b) in enumerate(self._ranges):
                for i in range(a, b):
                    out[i][k] = recv[r, (i - a) * batch:(i - a + 1) * batch]
        return out


def enable_sequence_parallel(model, group=None, gather_preds: bool = True) -> SequenceParallel:
    sp = SequenceParallel(group, gather_preds)
    model.sp_group = sp
    return sp
