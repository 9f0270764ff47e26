"""GPU test (-m gpu) of the sharded path (BASELINE.json config 5) on real hardware: two ranks, ONE
global buffer of the config-5 stream cut at a code-point boundary (the nominal cut lands inside a character), the
CUDA kernels per shard through the C ABI, the combining step, and the whole-buffer answer of the UNMODIFIED reference
(oracle/_ref, icelake / haswell) as the expectation: (error, position) — including an error injected into shard 1 —
the global count, the output offsets and the concatenated output, byte for byte.

With >= 2 GPUs the ranks run on cuda:0 / cuda:1 over NCCL with the device-side combiner (all_gather +
b200_sharded_combine_async: exactly bench.py's timed step).  On a one-GPU box both ranks share cuda:0 — NCCL refuses
two ranks on one device — and exchange the triplets over gloo (sharded.combine, host flavour); the kernels, the cut
and the combining arithmetic are the same.
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import ctypes, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import simdutf_b200 as b
from simdutf_b200 import sharded, synth
from tests._oracle import Oracle, Reference
port, rank, ngpu = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
world = 2
nccl = ngpu >= 2
local = rank if nccl else 0
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl" if nccl else "gloo", init_method="tcp://127.0.0.1:%s" % port, rank=rank, world_size=world,
                        **({"device_id": device} if nccl else {}))
lib = b.load(); b.set_device(local)
ref = Reference.load_or_none(); o = Oracle()
def whole(data):   # the expectation: one call of the reference on the whole buffer
    if ref is not None:
        return ref.convert_utf8_to_utf16le_with_errors(ref.best, data)
    return o.convert_utf8_to_utf16le_with_errors(data)
BLOCK = (1 << 20) + 1
seed = 5
found = None
for nominal in range(24 << 20, (24 << 20) + 64, 2):      # pick a size whose midpoint falls inside a character
    total = synth.stream_total_len(seed, nominal, "cpu", BLOCK)
    mid = synth.stream_range(seed, total // 2, total // 2 + 1, "cpu", BLOCK)
    if (int(mid[0]) & 0xC0) == 0x80:
        found = (nominal, total); break
assert found, "no mid-character cut found"
nominal, total = found
data = synth.stream_range(seed, 0, total, "cpu", BLOCK).numpy().copy()
cuts = sharded.utf8_shard_bounds(lambda i: int(data[i]), total, world)
assert cuts[1] < total // 2, "the cut must have been backed up to a lead byte"
stream = torch.cuda.current_stream(device); sp = ctypes.c_void_p(stream.cuda_stream)
for case in ("valid", "err_shard1", "err_both", "truncated_tail"):
    d = data.copy()
    if case in ("err_shard1", "err_both"): d[cuts[1] + 1000003] = 0xFF        # HEADER_BITS inside shard 1
    if case == "err_both": d[777] = 0xC0                                         # an earlier error in shard 0 wins
    if case == "truncated_tail": d = d[: total - 1] if (int(d[total - 1]) & 0xC0) == 0x80 else np.concatenate([d, np.array([0xE4], np.uint8)])
    n_total = int(d.size)
    mine = d[cuts[rank]: (cuts[rank + 1] if rank + 1 < world else n_total)]
    d_in = torch.from_numpy(mine.copy()).to(device)
    n = int(d_in.numel())
    units = b.utf16_length_from_utf8(d_in)
    d_out = torch.full((units + 64,), 0x5A5A, dtype=torch.int16, device=device)
    (werr, wcnt), wout = whole(d)
    if nccl:
        comb = sharded.DeviceCombiner(lib, device, n)
        st = lib.b200_convert_utf8_to_utf16le_async(ctypes.c_void_p(d_in.data_ptr()), n, ctypes.c_void_p(d_out.data_ptr()), comb.result_ptr, sp)
        assert st == 0, lib.b200_last_error()
        comb.step(sp)
        g = comb.read()
        lerr, lcnt = int(comb.triplet[1].item()) & 0xFFFFFFFF, int(comb.triplet[2].item())
    else:
        lerr, lcnt = b.convert_utf8_to_utf16le_with_errors(d_in, d_out)
        g = sharded.combine(lerr, lcnt, n, torch.device("cpu"), use_allreduce=True)
    assert (g.error, g.count) == (werr, wcnt), (case, rank, g, werr, wcnt)
    assert g.in_offset == cuts[rank], (case, g, cuts)
    assert bool((d_out[units:] == 0x5A5A).all()), "output overrun"
    if werr == 0:
        assert lerr == 0 and lcnt == units
        got = d_out[:units].cpu().numpy().view(np.uint16)
        assert np.array_equal(wout[g.out_offset: g.out_offset + units], got), (case, rank)
        tot = torch.tensor([units], dtype=torch.int64, device=device if nccl else "cpu")
        dist.all_reduce(tot)
        assert int(tot.item()) == wcnt == wout.size
    elif lerr == 0:   # a shard in front of the first error: its output is the reference's prefix
        (perr, pcnt), pout = whole(d[: cuts[rank + 1]]) if rank + 1 < world else ((1, 0), None)
        if perr == 0:
            got = d_out[:units].cpu().numpy().view(np.uint16)
            assert np.array_equal(pout[g.out_offset: g.out_offset + units], got), (case, rank)
dist.barrier(); dist.destroy_process_group()
print("OK", rank, "nccl" if nccl else "gloo", "ref" if ref is not None else "oracle")
'''


def test_two_rank_cut_buffer_vs_reference(tmp_path):
    import torch
    ngpu = torch.cuda.device_count()
    assert ngpu >= 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(ngpu)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o[-4000:]


# ------------------------------------------------------------------------------------------------------------
# One process, several devices: the native b200_mgpu_* entry points and the multi-device host-pointer path.
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def b():
    import simdutf_b200
    simdutf_b200.load()
    assert simdutf_b200.device_count() >= 1
    simdutf_b200.set_device(0)
    return simdutf_b200


def _whole(ref, oracle, data):
    if ref is not None:
        return ref.convert_utf8_to_utf16le_with_errors(ref.best, data)
    return oracle.convert_utf8_to_utf16le_with_errors(data)


@pytest.mark.parametrize("layout", ["distinct_devices", "three_shards_device0", "distinct_devices_no_nccl"])
def test_mgpu_entry_points_vs_reference(b, oracle, ref, layout):
    """b200_mgpu_convert_utf8_to_utf16le / _validate / _utf16_length on a cut buffer, shards on different devices
    (triplets over ncclAllGather) or several on one device (peer copies), against one whole-buffer reference call."""
    import numpy as np
    import torch
    from simdutf_b200 import sharded, synth
    ndev = b.device_count()
    if layout.startswith("distinct"):
        devs = list(range(min(ndev, 4)))
    else:
        devs = [0, 0, 0]
    b.set_tuning("no_nccl", 1 if layout.endswith("no_nccl") else 0)
    try:
        data = synth.mixed_utf8(6 << 20, seed=21).numpy().copy()
        for case in ("valid", "error_in_last_shard", "error_in_first_and_last"):
            d = data.copy()
            cuts = sharded.utf8_shard_bounds(lambda i: int(data[i]), d.size, len(devs))
            if case != "valid":
                d[cuts[-2] + 4321] = 0xF8
            if case == "error_in_first_and_last":
                d[99] = 0x80
            shards, outs = [], []
            for k, dev in enumerate(devs):
                t_in = torch.from_numpy(d[cuts[k]:cuts[k + 1]].copy()).to(f"cuda:{dev}")
                units = oracle.utf16_length_from_utf8(d[cuts[k]:cuts[k + 1]])
                t_out = torch.full((units + 64,), 0x5A5A, dtype=torch.int16, device=f"cuda:{dev}")
                shards.append((dev, t_in, t_out))
                outs.append((t_out, units))
            (werr, wcnt), wout = _whole(ref, oracle, d)
            res = b.mgpu("convert_utf8_to_utf16le", shards)
            want_gather = 1 if (layout == "distinct_devices") else 2
            assert b.load().b200_mgpu_last_gather() == want_gather
            for k, (e, c, in_off, out_off) in enumerate(res):
                assert (e, c) == (werr, wcnt), (layout, case, k, res)
                assert in_off == cuts[k]
                t_out, units = outs[k]
                assert bool((t_out[units:] == 0x5A5A).all()), "output overrun"
                if werr == 0:
                    got = t_out[:units].cpu().numpy().view(np.uint16)
                    assert np.array_equal(wout[out_off:out_off + units], got), (layout, case, k)
            vres = b.mgpu("validate_utf8_with_errors", [(dv, ti, None) for dv, ti, _ in shards])
            assert all((e, c) == oracle.validate_utf8_with_errors(d) for e, c, _, _ in vres), vres
            lres = b.mgpu("utf16_length_from_utf8", [(dv, ti, None) for dv, ti, _ in shards])
            assert all((e, c) == (0, oracle.utf16_length_from_utf8(d)) for e, c, _, _ in lres), lres
    finally:
        b.set_tuning("no_nccl", 0)


def test_host_path_over_all_devices_and_threads(b, oracle, ref):
    """b200_host_set_devices(n): a host buffer cut into segments dealt round-robin to every visible device gives the
    same (error, count) and the same output bytes as the reference; two threads calling at once do not disturb each
    other (per-thread host paths)."""
    import threading
    import numpy as np
    from simdutf_b200 import synth
    ndev = b.device_count()
    b.set_tuning("segment_mb", 2)
    try:
        data = synth.mixed_utf8(40 << 20, seed=22).numpy().copy()
        bad = data.copy()
        bad[31 << 20] = 0xFF
        (werr, wcnt), wout = _whole(ref, oracle, data)
        want_bad = oracle.validate_utf8_with_errors(bad)
        results = {}

        def work(tag, devices):
            b.set_device(0)
            b.host_set_devices(devices)
            out = np.zeros(wcnt + 64, dtype=np.uint16)
            r = b.convert_utf8_to_utf16le_with_errors(data, out)
            rb = b.convert_utf8_to_utf16le_with_errors(bad, np.zeros(wcnt + 64, dtype=np.uint16))
            n16 = b.utf16_length_from_utf8(data)
            results[tag] = (r, rb, n16, bool(np.array_equal(out[:wcnt], wout)), bool((out[wcnt:] == 0).all()))

        ts = [threading.Thread(target=work, args=(f"t{i}", ndev if i == 0 else 1)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for tag, (r, rb, n16, same, clean) in results.items():
            assert r == (werr, wcnt) and rb == want_bad and n16 == wcnt and same and clean, (tag, r, rb, n16, same, clean)
        assert len(results) == 2
    finally:
        b.set_tuning("segment_mb", 0)
        b.host_set_devices(1)


def test_device_pointers_through_the_host_flavour(b, oracle):
    """The C++ virtuals only have b200_host_*: device-resident data handed to them is recognised and processed in place
    (cudaPointerGetAttributes), and a call never leaves the caller's current CUDA device changed."""
    import ctypes
    import numpy as np
    import torch
    from simdutf_b200 import synth
    lib = b.load()
    data = synth.mixed_utf8(300_000, seed=23)
    dev = b.device_count() - 1
    d = data.to(f"cuda:{dev}")
    units = oracle.utf16_length_from_utf8(data.numpy())
    out = torch.zeros(units, dtype=torch.int16, device=f"cuda:{dev}")
    torch.cuda.set_device(0)
    res = b.Result()
    st = lib.b200_host_convert_utf8_to_utf16le(ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(out.data_ptr()), ctypes.byref(res))
    assert st == 0 and res.astuple() == (0, units)
    (_, _), wout = oracle.convert_utf8_to_utf16le_with_errors(data.numpy())
    assert np.array_equal(out.cpu().numpy().view(np.uint16), wout)
    assert torch.cuda.current_device() == 0
    # a tensor on the LAST device through the device-pointer flavour, without b200_set_device
    assert b.validate_utf8_with_errors(d) == (0, d.numel())
    assert torch.cuda.current_device() == 0
