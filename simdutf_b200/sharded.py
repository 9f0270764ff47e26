"""Multi-GPU composition of the hot path (BASELINE.json config 5).

One process per GPU.  The buffer is cut into contiguous shards at code-point boundaries — the rule of
simdutf::trim_partial_utf8 (reference src/scalar/utf8.h:257-288) exactly as benchmarks/threaded.cpp:69-74
uses it for two threads — so every shard is an independent call of the single-GPU entry points: no halo, no
data exchange.  What is exchanged is ONE tiny collective over NCCL (NVLink/NVSwitch):

  * all_gather of the triplet (input length, result.error, result.count) per rank.  From the gathered triplets
    every rank derives, locally, its global input / output offset, the global count and the global first error
    — the minimum over ranks of the packed key (global position << 8 | error_code), i.e. exactly the value a
    min-allreduce of that key returns (`combine(..., use_allreduce=True)` still runs that second collective, for
    callers that want the north-star formulation verbatim; the results are identical).

Outputs stay shard-local at the globally known offsets (the path has no bulk exchange step, so none is
invented).  `combine` works on host integers (one device->host read, for callers that need Python ints); the
timed path uses `DeviceCombiner`, which keeps everything on the device: the kernels write their result straight
into the buffer that is gathered, and one launch of the library's own b200_sharded_combine_async turns the
gathered triplets into the global result and this rank's offsets — no torch arithmetic kernels, no host sync.
The same code runs under gloo on CPU tensors for the world_size-2 tests (host flavour only).
"""
from __future__ import annotations

import ctypes
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


def fold_triplets(triplets, rank: int, count_is_length: bool = False):
    """(error, count, in_offset, out_offset) from the gathered per-shard (in_len, error, count) triplets, in shard
    order.  The same arithmetic as the device kernel k_sharded_combine (csrc/k_sharded.cu)."""
    in_off = out_off = my_in = my_out = 0
    best = NO_ERROR_KEY
    for r, (n, err, cnt) in enumerate(triplets):
        n, err, cnt = int(n), int(err) & 0xFFFFFFFF, int(cnt)
        if r == rank:
            my_in, my_out = in_off, out_off
        if err != 0:
            best = min(best, ((in_off + cnt) << 8) | (err & 0xFF))
        elif not count_is_length:
            out_off += cnt
        in_off += n
    if best == NO_ERROR_KEY:
        return 0, (in_off if count_is_length else out_off), my_in, my_out
    return best & 0xFF, best >> 8, my_in, my_out


def combine(local_error: int, local_count: int, local_in_len: int, device, group=None,
            count_is_length: bool = False, use_allreduce: bool = False) -> ShardedResult:
    """Turn per-shard `result{error,count}` into the global result (host flavour: one small device->host read).
    count_is_length: the operation's success count is the validated input length (validate_*), not an
    output size."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = torch.tensor([local_in_len, local_error, local_count], dtype=torch.int64, device=device)
    if world > 1:
        allv = torch.empty(3 * world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allv, mine, group=group)
    else:
        allv = mine
    trip = allv.view(world, 3).cpu().tolist()
    err, cnt, in_off, out_off = fold_triplets(trip, rank, count_is_length)
    if use_allreduce and world > 1:  # the north-star's second collective; must agree with the local fold
        key = NO_ERROR_KEY if local_error == 0 else (((in_off + local_count) << 8) | local_error)
        k = torch.tensor([key], dtype=torch.int64, device=device)
        dist.all_reduce(k, op=dist.ReduceOp.MIN, group=group)
        gkey = int(k.item())
        assert (gkey == NO_ERROR_KEY and err == 0) or (gkey & 0xFF, gkey >> 8) == (err, cnt)
    return ShardedResult(err, cnt, in_off, out_off, local_count)


class DeviceCombiner:
    """Device-resident combining step for the timed path.  `triplet` (3 x int64 on the device) is
    [input length, b200_result]: pass `result_ptr` to the *_async entry point so that the kernel writes its result
    in place.  `step()` enqueues the all_gather and the library's combine kernel on the current stream; `read()`
    synchronises and returns a ShardedResult."""

    def __init__(self, lib, device, in_len: int, group=None, count_is_length: bool = False):
        self.lib, self.device, self.group = lib, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.count_is_length = 1 if count_is_length else 0
        self.triplet = torch.zeros(3, dtype=torch.int64, device=device)
        self.triplet[0] = in_len
        self.gathered = torch.zeros(3 * self.world, dtype=torch.int64, device=device)
        self.out = torch.zeros(4, dtype=torch.int64, device=device)  # b200_sharded_result

    @property
    def result_ptr(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(self.triplet.data_ptr() + 8)

    def step(self, stream_ptr: ctypes.c_void_p) -> None:
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.triplet, group=self.group)
            src = self.gathered
        else:
            src = self.triplet
        st = self.lib.b200_sharded_combine_async(ctypes.c_void_p(src.data_ptr()), self.world, self.rank,
                                                 self.count_is_length, ctypes.c_void_p(self.out.data_ptr()), stream_ptr)
        if st:
            raise RuntimeError("b200_sharded_combine_async failed: " + self.lib.b200_last_error().decode())

    def read(self) -> ShardedResult:
        o = self.out.cpu().tolist()
        return ShardedResult(int(o[0]) & 0xFFFFFFFF, int(o[1]), int(o[2]), int(o[3]), int(self.triplet[2].item()))
