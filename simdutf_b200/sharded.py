"""Multi-GPU composition of the hot path (BASELINE.json config 5).

One process per GPU.  The buffer is cut into contiguous shards at code-point boundaries — the rule of
simdutf::trim_partial_utf8 (reference src/scalar/utf8.h:257-288) exactly as benchmarks/threaded.cpp:69-74
uses it for two threads — so every shard is an independent call of the single-GPU entry points: no halo, no
data exchange.  What is exchanged is two tiny collectives over NCCL (NVLink/NVSwitch):

  * all_gather of (input length, output length) per rank  -> each rank's global input / output offset and
    the global count;
  * all_reduce(MIN) of the packed first-error key (global position << 8 | error_code; INT64_MAX = none)
    -> the same global `result{error, count}` on every rank.

Outputs stay shard-local at the globally known offsets (the path has no bulk exchange step, so none is
invented).  The same code runs under gloo on CPU tensors for the world_size-2 tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import torch
import torch.distributed as dist

NO_ERROR_KEY = (1 << 63) - 1


def utf8_cut_before(peek: Callable[[int], int], pos: int, lo: int = 0) -> int:
    """Largest cut <= pos that does not split a character: back up over continuation bytes, at most 3
    (if there is no lead within 3 bytes the data is invalid there and any cut reports the same error)."""
    cut = pos
    for _ in range(3):
        if cut <= lo or (peek(cut) & 0xC0) != 0x80:
            break
        cut -= 1
    return cut


def utf8_shard_bounds(peek: Callable[[int], int], total_len: int, world: int) -> list[int]:
    """world+1 cut points: 0 = c_0 <= c_1 <= ... <= c_world = total_len, c_k = k*N/G backed up to a lead."""
    cuts = [0]
    for k in range(1, world):
        c = utf8_cut_before(peek, total_len * k // world, cuts[-1])
        cuts.append(max(c, cuts[-1]))
    cuts.append(total_len)
    return cuts


def utf16_shard_bounds(peek: Callable[[int], int], total_units: int, world: int) -> list[int]:
    """Same for UTF-16LE: never cut between a high and a low surrogate (reference src/scalar/utf16.h:114-124)."""
    cuts = [0]
    for k in range(1, world):
        c = total_units * k // world
        if c > cuts[-1] and c < total_units and (peek(c) & 0xFC00) == 0xDC00 and (peek(c - 1) & 0xFC00) == 0xD800:
            c -= 1
        cuts.append(max(c, cuts[-1]))
    cuts.append(total_units)
    return cuts


@dataclass
class ShardedResult:
    error: int          # global simdutf::error_code
    count: int          # global: total output elements on success, global input position on error
    in_offset: int      # this rank's first input element in the global buffer
    out_offset: int     # this rank's first output element in the global output
    local_count: int    # this rank's own result.count


def combine(local_error: int, local_count: int, local_in_len: int, device, group=None,
            count_is_length: bool = False) -> ShardedResult:
    """Turn per-shard `result{error,count}` into the global result.
    count_is_length: the operation's success count is the validated input length (validate_*), not an
    output size."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local_out = 0 if (local_error != 0 or count_is_length) else local_count
    mine = torch.tensor([local_in_len, local_out], dtype=torch.int64, device=device)
    if world > 1:
        allv = torch.empty(2 * world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allv, mine, group=group)
        allv = allv.view(world, 2).cpu()
    else:
        allv = mine.view(1, 2).cpu()
    in_off = int(allv[:rank, 0].sum().item())
    out_off = int(allv[:rank, 1].sum().item())
    key = NO_ERROR_KEY if local_error == 0 else (((in_off + local_count) << 8) | local_error)
    k = torch.tensor([key], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(k, op=dist.ReduceOp.MIN, group=group)
    gkey = int(k.item())
    if gkey == NO_ERROR_KEY:
        total = int(allv[:, 0].sum().item()) if count_is_length else int(allv[:, 1].sum().item())
        return ShardedResult(0, total, in_off, out_off, local_count)
    return ShardedResult(gkey & 0xFF, gkey >> 8, in_off, out_off, local_count)
