#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 backend (BASELINE.json metric: input GB/s).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the reference's own CPU kernels

Workload (config.workload):
  N = 1   BASELINE.json configs[1]: utf16_length_from_utf8 + convert_utf8_to_utf16le_with_errors on 1 GiB of
          synthetic mixed 1-4-byte UTF-8 (25% ASCII / Latin / CJK / emoji by code point, seed 2).
  N > 1   configs[4] run weak: every rank owns one 1 GiB shard of that distribution (cut on a code-point
          boundary), runs the same two calls, then the two tiny collectives of the sharded path
          (all_gather of lengths, all_reduce-min of the first-error key) — see simdutf_b200/sharded.py.
A step = one pass of that path over the batch.  value = all ranks' input bytes / max-over-ranks device time
(CUDA events on the launching stream), inputs resident in HBM.  e2e = the same two calls through the
host-pointer C ABI (b200_host_*), pinned host buffers, H2D and D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GIB = 1 << 30
METRIC = "input GB/s, utf16_length_from_utf8 + convert_utf8_to_utf16le_with_errors (UTF-8 -> UTF-16LE)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own kernels (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_run(sample_bytes: int, threads: int, steps: int, warmup: int, seed: int = 2, stream: bool = False):
    """Times utf16_length_from_utf8 + convert_utf8_to_utf16le_with_errors of the UNMODIFIED reference library
    (best kernel of this host: icelake or haswell) on a `sample_bytes` sample of the config-2 distribution,
    split over `threads` host threads with the recipe of the reference's benchmarks/threaded.cpp:69-88.
    Falls back to the oracle port (1 thread) only if oracle/_ref did not travel."""
    import numpy as np
    from simdutf_b200 import synth
    from tests._oracle import Oracle, Reference
    import torch
    # generation is plumbing: done on the GPU when there is one (seconds instead of minutes for 1 GiB; the same generator
    # and seed as this repo's arm, so the two arms time the very same bytes), then moved to the host
    gdev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else torch.device("cpu")
    if stream:
        total = synth.stream_total_len(seed, sample_bytes, gdev)
        data = synth.stream_range(seed, 0, total, gdev).cpu().numpy()
    else:
        data = synth.mixed_utf8(sample_bytes, seed=seed, device=gdev).cpu().numpy()
    n = int(data.size)
    ref = Reference.load_or_none()
    out = np.empty(n + 64, dtype=np.uint16)
    times = []
    if ref is not None:
        kind, impl = "reference", ref.best
        fn = ref.L.ref_mt_utf16_length_then_convert_utf8_to_utf16le
        args = (impl.encode(), ctypes.c_void_p(data.ctypes.data), ctypes.c_size_t(n), ctypes.c_void_p(out.ctypes.data), threads)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            units = fn(*args)
            dt = time.perf_counter() - t0
            assert units > 0
            if i >= warmup:
                times.append(dt)
    else:
        kind, impl, threads = "port", "oracle.c", 1
        o = Oracle()
        small = data[: min(n, 32 << 20)]
        cut = o.trim_partial_utf8(small)
        small = small[:cut]
        n = int(small.size)
        for i in range(max(1, min(warmup, 1)) + max(1, min(steps, 3))):
            t0 = time.perf_counter()
            o.utf16_length_from_utf8(small)
            o.convert_utf8_to_utf16le_with_errors(small)
            dt = time.perf_counter() - t0
            if i >= 1:
                times.append(dt)
    sec = sum(times) / len(times)
    return {"value": n / sec / 1e9, "unit": "GB/s", "cores": threads, "kind": kind, "impl": impl,
            "sample": f"{n} bytes of the config-2 mixed UTF-8 distribution (seed {seed}), {len(times)} timed passes, "
                      f"mean; utf16_length_from_utf8 + convert_utf8_to_utf16le_with_errors per pass",
            "ms_per_step": sec * 1e3, "bytes": n}


def threaded_cpp_run(sample_bytes: int, seed: int = 2):
    """The reference's own benchmarks/threaded.cpp, compiled as is into oracle/_ref/threaded (oracle/Makefile), on a
    file dump of the config-2 sample: convert_utf8_to_utf16le on one thread and split over two (its hard-wired
    design, reference benchmarks/threaded.cpp:36-94).  Returns None when the binary is absent."""
    exe = os.path.join(ROOT, "oracle", "_ref", "threaded")
    if not os.path.exists(exe):
        return None
    import tempfile
    import torch
    from simdutf_b200 import synth
    data = synth.mixed_utf8(sample_bytes, seed=seed, device=torch.device("cpu")).numpy()
    # threaded.cpp:70-73 backs its midpoint up only over continuation bytes 0x80..0x9F; give it a file whose midpoint
    # is a character start (drop a few trailing characters until it is) so that both halves are valid UTF-8
    for _ in range(64):
        if (int(data[data.size // 2]) & 0xC0) != 0x80:
            break
        cut = data.size - 1
        while (int(data[cut]) & 0xC0) == 0x80:
            cut -= 1
        data = data[:cut]
    n = int(data.size)
    with tempfile.NamedTemporaryFile(suffix=".utf8", delete=False) as f:
        f.write(data.tobytes())
        path = f.name
    try:
        p = subprocess.run([exe, path], capture_output=True, text=True, timeout=300)
    finally:
        os.unlink(path)
    ns = {}
    for line in p.stdout.splitlines():
        for key in ("singlethread", "doublethread"):
            if line.startswith(key + ":"):
                ns[key] = float(line.split(":")[1])
    if len(ns) != 2:
        return {"unavailable": (p.stderr or p.stdout)[-200:].strip().replace("\n", " | ")}
    return {"single_thread_gbs": n / ns["singlethread"], "two_threads_gbs": n / ns["doublethread"], "unit": "GB/s",
            "kind": "reference", "what": "benchmarks/threaded.cpp as is: convert_utf8_to_utf16le only (no length query)",
            "sample": f"{n} bytes of the config-2 mixed UTF-8 distribution (seed {seed}) from a file"}


def workload_config(world: int, shard_bytes: int, total_bytes: int | None):
    """The `config` object, identical for this repo's arm and the reference arm (the driver compares them)."""
    if world == 1 and total_bytes is None:
        return {
            "workload": "configs[1]: utf16_length_from_utf8 + convert_utf8_to_utf16le_with_errors on 1 GiB of synthetic mixed "
                        "1-4-byte UTF-8 (seed 2), bit-exact output",
            "nominal_bytes_per_gpu": shard_bytes, "total_nominal_bytes": shard_bytes, "seed": 2,
            "char_mix": "25% each 1/2/3/4-byte code points (ASCII / Latin / CJK / emoji), i.i.d.",
            "l2_policy": "inputs and outputs (>= 1 GiB each) far exceed the 126 MB L2; no flush needed",
            "parallelism": "single GPU",
        }
    total = total_bytes if total_bytes is not None else world * shard_bytes
    return {
        "workload": f"configs[4]: ONE global buffer of {total / GIB:g} GiB of synthetic mixed 1-4-byte UTF-8 (counter-based block "
                    f"stream, seed 5; 16 GiB at 8 GPUs) cut into {world} shards at k*N/G backed up to a code-point boundary "
                    "(simdutf_b200.sharded.utf8_shard_bounds); per shard utf16_length_from_utf8 + "
                    "convert_utf8_to_utf16le_with_errors; one NCCL all_gather of (length, error, count) per step; global "
                    "first error and output offsets derived on the device from the gathered triplets",
        "nominal_bytes_per_gpu": total // world, "total_nominal_bytes": total, "seed": 5,
        "char_mix": "25% each 1/2/3/4-byte code points (ASCII / Latin / CJK / emoji), i.i.d.",
        "l2_policy": "inputs and outputs (>= 1 GiB each) far exceed the 126 MB L2; no flush needed",
        "parallelism": f"shards x{world}",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    threads = os.cpu_count() or 1
    shard = args.shard_bytes if args.shard_bytes else (GIB if world == 1 else 2 * GIB)
    cfg = workload_config(world, shard, args.total_bytes)
    # N = 1: the whole 1 GiB buffer of configs[1] (same bytes as this repo's arm when a GPU is there to generate it);
    # N > 1: a bounded 1 GiB sample of the config-5 stream
    sample = args.cpu_sample_bytes if args.cpu_sample_bytes else min(GIB, cfg["total_nominal_bytes"])
    r = cpu_reference_run(sample, threads, args.steps, args.warmup, seed=2 if world == 1 and args.total_bytes is None else 5,
                          stream=not (world == 1 and args.total_bytes is None))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "kernel": r["impl"], "bytes_per_step": r["bytes"]},
        "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not add to
    it: keep a private handle to the real stdout and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


_REAL_STDOUT = None
EXTRA_WARMUP_STEPS = 512  # fixed count (never a rank-local clock: every rank must issue the same collectives)


def make_shard(world, rank, shard_bytes, total_bytes, device, synth, sharded):
    """This rank's input.  N = 1: configs[1] (1 GiB, seed 2).  N > 1 (or --total-bytes): this rank's slice of the ONE
    global config-5 stream, cut at k*N/G backed up to a code-point boundary.  Returns (tensor, cuts or None)."""
    if world == 1 and total_bytes is None:
        return synth.mixed_utf8(shard_bytes, seed=2, device=device), None
    seed = 5
    nominal = total_bytes if total_bytes is not None else world * shard_bytes
    total = synth.stream_total_len(seed, nominal, device)
    cache = {}

    def peek(i):
        if i not in cache:
            lo = max(0, i - 4)
            vals = synth.stream_range(seed, lo, min(total, i + 4), device).cpu().tolist()
            for j, v in enumerate(vals):
                cache[lo + j] = v
        return cache[i]

    cuts = sharded.utf8_shard_bounds(peek, total, world)
    return synth.stream_range(seed, cuts[rank], cuts[rank + 1], device), cuts


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shard-bytes", type=int, default=0, help="input bytes per GPU (default: 1 GiB at N=1, 2 GiB at N>1)")
    ap.add_argument("--total-bytes", type=int, default=None, help="size of the ONE global config-5 buffer (default N x shard)")
    ap.add_argument("--cpu-sample-bytes", type=int, default=0, help="reference arm / cpu_baseline sample (default: 1 GiB)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-extras", action="store_true", help="skip the config-3/4 and §8f side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import simdutf_b200 as b
    from simdutf_b200 import sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        cpu_group = dist.new_group(backend="gloo")  # host-side barrier for the phases in which rank 0 drives every GPU itself
    lib = b.load()
    b.set_device(local_rank)
    for kv in filter(None, os.environ.get("B200_BENCH_TUNE", "").split(",")):  # experiments only: "conv_minb=4,segment_mb=64"
        k, v = kv.split("=")
        b.set_tuning(k, int(v))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    shard_bytes = args.shard_bytes if args.shard_bytes else (GIB if world == 1 else 2 * GIB)
    cfg = workload_config(world, shard_bytes, args.total_bytes)

    # ---- inputs resident in HBM ------------------------------------------------------------------------------
    d_in, cuts = make_shard(world, rank, shard_bytes, args.total_bytes, device, synth, sharded)
    n = int(d_in.numel())
    units = b.utf16_length_from_utf8(d_in)
    d_out = torch.empty(units, dtype=torch.int16, device=device)
    d_cnt = torch.zeros(1, dtype=torch.int64, device=device)
    # the convert kernel writes its b200_result straight into the triplet that is all-gathered: the multi-GPU step
    # adds one collective and one launch of the library's combine kernel, no copy, no host synchronisation
    comb = sharded.DeviceCombiner(lib, device, n)
    stream = torch.cuda.current_stream(device)
    sp = ctypes.c_void_p(stream.cuda_stream)
    in_p, out_p = ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_out.data_ptr())
    cnt_p, res_p = ctypes.c_void_p(d_cnt.data_ptr()), comb.result_ptr
    sharded_run = cuts is not None

    def step(ev=None):
        if ev:
            ev[0].record(stream)
        st = lib.b200_utf16_length_from_utf8_async(in_p, n, cnt_p, sp)
        if ev:
            ev[1].record(stream)
        st |= lib.b200_convert_utf8_to_utf16le_async(in_p, n, out_p, res_p, sp)
        if ev:
            ev[2].record(stream)
        if st:
            raise RuntimeError("b200 launch failed: " + lib.b200_last_error().decode())
        if sharded_run:
            comb.step(sp)

    # nvidia-smi samples every 200 ms and a step takes ~1-3 ms: start sampling before the warm-up and keep the GPU
    # under the same load for a FIXED number of untimed steps, so that the sampler sees the clocks under load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W + EXTRA_WARMUP_STEPS):
        step()
    torch.cuda.synchronize(device)
    local = comb.triplet.cpu().tolist()
    assert int(d_cnt.item()) == units and local[1] & 0xFFFFFFFF == 0 and local[2] == units

    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = b.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    t_beg.record(stream)
    for i in range(K):
        step(evs[i])
    t_end.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    clocks = sampler.stop() if rank == 0 else None
    launches = b.launch_count() - launches0
    ms_total = t_beg.elapsed_time(t_end)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    tot = torch.tensor([n, units], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    total_bytes, total_units = (int(x) for x in tot.tolist())
    ms_per_step = ms_total / K
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    conv_ms = [e[1].elapsed_time(e[2]) for e in evs]
    len_ms = [e[0].elapsed_time(e[1]) for e in evs]
    conv_avg = sum(conv_ms) / len(conv_ms)
    peak, peak_src = measured_peak()
    algo_bytes = n + 2 * units  # input read once + output written once
    achieved = algo_bytes / (conv_avg * 1e-3) / 1e9
    traffic = traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get("convert_utf8_to_utf16le", {})
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    except Exception:
        pass
    shard_check = None
    if sharded_run:  # the combined result every rank holds: global SUCCESS, global count, this rank's offsets
        g = comb.read()
        assert (g.error, g.count) == (0, total_units), (g, total_units)
        assert g.in_offset == cuts[rank] and cuts[-1] == total_bytes
        offs = torch.tensor([g.out_offset], dtype=torch.int64, device=device)
        alloffs = torch.empty(world, dtype=torch.int64, device=device)
        allunits = torch.empty(world, dtype=torch.int64, device=device)
        if world > 1:
            dist.all_gather_into_tensor(alloffs, offs)
            dist.all_gather_into_tensor(allunits, torch.tensor([units], dtype=torch.int64, device=device))
        else:
            alloffs, allunits = offs, torch.tensor([units], dtype=torch.int64, device=device)
        ex = (torch.cumsum(allunits, 0) - allunits).tolist()
        assert alloffs.tolist() == ex, (alloffs.tolist(), ex)
        shard_check = {"cuts": cuts, "cut_backups": [int(cuts[-1] * k // world - cuts[k]) for k in range(1, world)],
                       "global_result": [g.error, g.count], "out_offsets": ex}
    else:
        assert local[1] & 0xFFFFFFFF == 0 and local[2] == units

    # ---- e2e: host-pointer C ABI, pinned host buffers, copies inside the timed region ---------------------------
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in)
    h_out = torch.empty(units, dtype=torch.int16, pin_memory=True)
    hres, hcnt = b.Result(), ctypes.c_uint64()
    hin_p, hout_p = ctypes.c_void_p(h_in.data_ptr()), ctypes.c_void_p(h_out.data_ptr())

    def e2e_step():
        st = lib.b200_host_utf16_length_from_utf8(hin_p, n, ctypes.byref(hcnt))
        st |= lib.b200_host_convert_utf8_to_utf16le(hin_p, n, hout_p, ctypes.byref(hres))
        if st:
            raise RuntimeError("b200 host call failed: " + lib.b200_last_error().decode())

    e2e_step()
    assert hcnt.value == units and hres.astuple() == (0, units)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_bytes / float(te.item()) / 1e9
    same = bool(torch.equal(h_out[: 1 << 20], d_out[: 1 << 20].cpu()))
    assert same, "host path and device path disagree"
    del h_in, h_out

    # ---- e2e at N > 1, the drop-in way: ONE process (rank 0) hands the WHOLE global buffer, in host memory, to the host-
    # pointer C ABI, which deals its segments to all N devices (b200_host_set_devices).  The other ranks wait on a
    # host-side (gloo) barrier so that no NCCL kernel of theirs spins on the GPUs meanwhile.
    e2e_one = None
    if sharded_run and world > 1:
        del d_out
        torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)
        if rank == 0:
            e2e_one = one_process_e2e(b, lib, synth, torch, device, world, total_bytes, total_units, args.e2e_steps)
        dist.barrier(group=cpu_group)

    # ---- the other half of BASELINE.json's metric: validate_utf8_with_errors (config 1 ASCII, and the mixed buffer) ----
    val = {}
    if rank == 0 and world == 1:
        d_out = None
        val = validate_measurements(lib, synth, torch, device, stream, sp, peak, K, d_in)
    extras = {}
    if not args.no_extras and rank == 0 and world == 1:
        del d_in
        extras = side_measurements(b, lib, synth, torch, device, stream, sp, peak, K)

    cpu = cpu1 = cpu_thr = None
    if rank == 0 and world == 1:  # the CPU baseline is reported at N = 1 only
        sample = args.cpu_sample_bytes if args.cpu_sample_bytes else GIB
        cpu = cpu_reference_run(sample, os.cpu_count() or 1, 3, 1)
        cpu1 = cpu_reference_run(min(sample, 128 << 20), 1, 2, 1)
        cpu_thr = threaded_cpp_run(min(sample, 64 << 20))
    if rank == 0:
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": CONVERT_KERNEL_NAME,
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": conv_avg,
                "best_launch_ms": min(conv_ms), "median_launch_ms": statistics.median(conv_ms),
                "peak_source": peak_src,
                "length_kernel_ms": sum(len_ms) / len(len_ms),
                "length_kernel_gbs": n / (sum(len_ms) / len(len_ms) * 1e-3) / 1e9,
                "length_kernel_frac": n / (sum(len_ms) / len(len_ms) * 1e-3) / 1e9 / peak}
        roof.update(val)
        # BASELINE.json's configs 3 and 4, so that a record that keeps only the roofline dict carries them too
        for src, dst in (("config3_convert_utf16le_to_utf8_2GiB", "convert_utf16le_to_utf8"),
                         ("config4_base64_to_binary_2GiB", "base64_to_binary")):
            row = extras.get(src)
            if isinstance(row, dict):
                for k_src, k_dst in (("ms", "_ms"), ("input_gbs", "_input_gbs"), ("frac_of_peak", "_frac")):
                    if k_src in row:
                        roof[dst + k_dst] = row[k_src]
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "bytes_per_step": total_bytes, "utf16_units_per_step": total_units,
            "roofline": roof,
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
            "cpu_baseline_1thread": ({k: cpu1[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu1 else None),
            "cpu_baseline_threaded_cpp": cpu_thr,
            "e2e": ({"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": 2 * total_bytes,
                     "d2h_bytes_per_step": 2 * total_units + 24 * world,
                     "api": "b200_host_utf16_length_from_utf8 + b200_host_convert_utf8_to_utf16le (pinned host buffers)",
                     "ms_per_step": float(te.item()) * 1e3} if e2e_one is None else
                    dict(e2e_one, per_rank_processes={"value": e2e_value, "ms_per_step": float(te.item()) * 1e3,
                                                      "what": "the same two calls issued by N processes, each on its own shard and GPU"})),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "sharded": shard_check,
            "extra": extras,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def one_process_e2e(b, lib, synth, torch, device, world, total_bytes, total_units, steps):
    """utf16_length_from_utf8 + convert_utf8_to_utf16le on the whole global buffer through b200_host_* from ONE process
    over `world` devices.  Host buffers are pinned; every byte crosses PCIe inside the timed region."""
    seed = 5
    h_in = torch.empty(total_bytes, dtype=torch.uint8, pin_memory=True)
    for lo in range(0, total_bytes, GIB):
        hi = min(total_bytes, lo + GIB)
        h_in[lo:hi].copy_(synth.stream_range(seed, lo, hi, device))
    h_out = torch.empty(total_units, dtype=torch.int16, pin_memory=True)
    torch.cuda.empty_cache()
    hres, hcnt = b.Result(), ctypes.c_uint64()
    hin_p, hout_p = ctypes.c_void_p(h_in.data_ptr()), ctypes.c_void_p(h_out.data_ptr())
    b.host_set_devices(world)

    def step():
        st = lib.b200_host_utf16_length_from_utf8(hin_p, total_bytes, ctypes.byref(hcnt))
        st |= lib.b200_host_convert_utf8_to_utf16le(hin_p, total_bytes, hout_p, ctypes.byref(hres))
        if st:
            raise RuntimeError("b200 host call failed: " + lib.b200_last_error().decode())

    try:
        step()
        assert hcnt.value == total_units and hres.astuple() == (0, total_units), (hcnt.value, hres.astuple(), total_units)
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        sec = (time.perf_counter() - t0) / steps
    finally:
        b.host_set_devices(1)
    return {"value": total_bytes / sec / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 2 * total_bytes,
            "d2h_bytes_per_step": 2 * total_units + 24,
            "api": f"b200_host_utf16_length_from_utf8 + b200_host_convert_utf8_to_utf16le on the whole buffer from ONE process, "
                   f"segments dealt to {world} devices (b200_host_set_devices), pinned host buffers",
            "ms_per_step": sec * 1e3}


CONVERT_KERNEL_NAME = ("k_utf8_transcode_v3 (convert_utf8_to_utf16le_with_errors = ONE launch: bit-plane transcoder, one CTA of 16 worker warps + "
                       "a scan warp per SM; output offsets from a decoupled look-back that runs two tiles ahead of the copy-out; "
                       "the input crosses HBM once)")


def _timeit(torch, device, stream, fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / reps


def validate_measurements(lib, synth, torch, device, stream, sp, peak, K, mixed):
    """validate_utf8_with_errors, the first half of BASELINE.json's metric: config 1 (1 GiB ASCII) and the mixed buffer of
    config 2.  Device-resident, CUDA events; algorithmic bytes = the input, read once.  Goes into the `roofline` object."""
    d_res = torch.zeros(4, dtype=torch.int64, device=device)
    res_p = ctypes.c_void_p(d_res.data_ptr())
    out = {}
    a = synth.ascii_text(GIB, seed=1, device=device)
    n = int(a.numel())
    ap_ = ctypes.c_void_p(a.data_ptr())
    ms = _timeit(torch, device, stream, lambda: lib.b200_validate_utf8_with_errors_async(ap_, n, res_p, sp), K)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == n
    out["validate_utf8_ascii_gbs"] = n / ms / 1e6
    out["validate_utf8_ascii_frac"] = n / ms / 1e6 / peak
    out["validate_utf8_ascii_ms"] = ms
    del a
    n = int(mixed.numel())
    mp = ctypes.c_void_p(mixed.data_ptr())
    ms = _timeit(torch, device, stream, lambda: lib.b200_validate_utf8_with_errors_async(mp, n, res_p, sp), K)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == n
    out["validate_utf8_mixed_gbs"] = n / ms / 1e6
    out["validate_utf8_mixed_frac"] = n / ms / 1e6 / peak
    out["validate_utf8_mixed_ms"] = ms
    return out


def side_measurements(b, lib, synth, torch, device, stream, sp, peak, K):
    """Other BASELINE.json configs, timed the same way (device-resident, CUDA events), reported under "extra"."""
    out = {}

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / reps

    d_res = torch.zeros(4, dtype=torch.int64, device=device)
    res_p = ctypes.c_void_p(d_res.data_ptr())
    # config 1: validate_utf8_with_errors, 1 GiB ASCII
    a = synth.ascii_text(GIB, seed=1, device=device)
    n = int(a.numel())
    ap_ = ctypes.c_void_p(a.data_ptr())
    ms = timeit(lambda: lib.b200_validate_utf8_with_errors_async(ap_, n, res_p, sp), K)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == n
    out["config1_validate_utf8_ascii_1GiB"] = {"input_gbs": n / ms / 1e6, "ms": ms, "frac_of_peak": n / ms / 1e6 / peak}
    ms = timeit(lambda: lib.b200_count_utf8_async(ap_, n, res_p, sp), K)
    out["count_utf8_ascii_1GiB"] = {"input_gbs": n / ms / 1e6, "ms": ms, "frac_of_peak": n / ms / 1e6 / peak}
    del a
    # validate on the mixed buffer (no fast path applies)
    m = synth.mixed_utf8(GIB, seed=2, device=device)
    n = int(m.numel())
    mp = ctypes.c_void_p(m.data_ptr())
    ms = timeit(lambda: lib.b200_validate_utf8_with_errors_async(mp, n, res_p, sp), K)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == n
    out["validate_utf8_mixed_1GiB"] = {"input_gbs": n / ms / 1e6, "ms": ms, "frac_of_peak": n / ms / 1e6 / peak}
    chars = b.count_utf8(m)
    o32 = torch.empty(chars, dtype=torch.int32, device=device)
    ms = timeit(lambda: lib.b200_convert_utf8_to_utf32_async(mp, n, ctypes.c_void_p(o32.data_ptr()), res_p, sp), max(3, K // 2))
    out["convert_utf8_to_utf32_mixed_1GiB"] = {"input_gbs": n / ms / 1e6, "ms": ms, "frac_of_peak": (n + 4 * chars) / ms / 1e6 / peak}
    # §8f rank 1: UTF-8 -> UTF-16BE (same kernel, byte planes swapped)
    units = b.utf16_length_from_utf8(m)
    o16 = torch.empty(units, dtype=torch.int16, device=device)
    ms = timeit(lambda: lib.b200_convert_utf8_to_utf16be_async(mp, n, ctypes.c_void_p(o16.data_ptr()), res_p, sp), max(3, K // 2))
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == units
    out["next_convert_utf8_to_utf16be_mixed_1GiB"] = {"input_gbs": n / ms / 1e6, "ms": ms, "frac_of_peak": (n + 2 * units) / ms / 1e6 / peak}
    del m, o32, o16
    # config 3: UTF-16LE -> UTF-8, 2 GiB
    u = synth.mixed_utf16le(GIB, seed=3, device=device)
    nu = int(u.numel())
    up = ctypes.c_void_p(u.data_ptr())
    nb = b.utf8_length_from_utf16le(u)
    o8 = torch.empty(nb, dtype=torch.uint8, device=device)
    ms = timeit(lambda: lib.b200_convert_utf16le_to_utf8_async(up, nu, ctypes.c_void_p(o8.data_ptr()), res_p, sp), max(3, K // 2))
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == nb
    out["config3_convert_utf16le_to_utf8_2GiB"] = {"input_gbs": 2 * nu / ms / 1e6, "ms": ms, "frac_of_peak": (2 * nu + nb) / ms / 1e6 / peak}
    ms = timeit(lambda: lib.b200_count_utf16le_async(up, nu, res_p, sp), K)
    out["config3_count_utf16le_2GiB"] = {"input_gbs": 2 * nu / ms / 1e6, "ms": ms, "frac_of_peak": 2 * nu / ms / 1e6 / peak}
    ms = timeit(lambda: lib.b200_utf8_length_from_utf16le_async(up, nu, res_p, sp), K)
    out["config3_utf8_length_from_utf16le_2GiB"] = {"input_gbs": 2 * nu / ms / 1e6, "ms": ms, "frac_of_peak": 2 * nu / ms / 1e6 / peak}
    # §8f rank 1: change_endianness_utf16, then UTF-16BE -> UTF-8 on the swapped buffer
    ube = torch.empty_like(u)
    ubp = ctypes.c_void_p(ube.data_ptr())
    ms = timeit(lambda: lib.b200_change_endianness_utf16_async(up, nu, ubp, res_p, sp), K)
    out["next_change_endianness_utf16_2GiB"] = {"input_gbs": 2 * nu / ms / 1e6, "ms": ms, "frac_of_peak": 4 * nu / ms / 1e6 / peak}
    ms = timeit(lambda: lib.b200_convert_utf16be_to_utf8_async(ubp, nu, ctypes.c_void_p(o8.data_ptr()), res_p, sp), max(3, K // 2))
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == nb
    out["next_convert_utf16be_to_utf8_2GiB"] = {"input_gbs": 2 * nu / ms / 1e6, "ms": ms, "frac_of_peak": (2 * nu + nb) / ms / 1e6 / peak}
    del u, o8, ube
    # config 4: base64 decode, 2 GiB of text with CRLF every 76 + sparse whitespace
    text, payload = synth.base64_text(2 * GIB, seed=4, device=device)
    nt = int(text.numel())
    ob = torch.empty(nt // 4 * 3 + 3, dtype=torch.uint8, device=device)
    ms = timeit(lambda: lib.b200_base64_to_binary_async(ctypes.c_void_p(text.data_ptr()), nt, ctypes.c_void_p(ob.data_ptr()), 0, 0, res_p, sp),
                max(3, K // 2))
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[2].item()) == int(payload.numel())
    out["config4_base64_to_binary_2GiB"] = {"input_gbs": nt / ms / 1e6, "ms": ms,
                                            "frac_of_peak": (nt + int(payload.numel())) / ms / 1e6 / peak}
    # §8f rank 2: binary_to_base64 of the decoded payload
    npay = int(payload.numel())
    enc = torch.empty((npay + 2) // 3 * 4, dtype=torch.uint8, device=device)
    ms = timeit(lambda: lib.b200_binary_to_base64_async(ctypes.c_void_p(payload.data_ptr()), npay, ctypes.c_void_p(enc.data_ptr()), 0, res_p, sp), K)
    out["next_binary_to_base64_1.5GiB"] = {"input_gbs": npay / ms / 1e6, "ms": ms, "frac_of_peak": (npay + int(enc.numel())) / ms / 1e6 / peak}
    del text, payload, ob, enc

    def rec(key, ms, nin, nout):
        out[key] = {"input_gbs": nin / ms / 1e6, "ms": ms, "frac_of_peak": (nin + nout) / ms / 1e6 / peak}

    def p_(t):
        return ctypes.c_void_p(t.data_ptr())

    # §8f rank 1 (second part): the UTF-32 family on the code points of 1 GiB of mixed UTF-8
    m = synth.mixed_utf8(GIB, seed=2, device=device)
    n8 = int(m.numel())
    cps = b.count_utf8(m)
    units = b.utf16_length_from_utf8(m)
    u32 = torch.empty(cps, dtype=torch.int32, device=device)
    assert b.convert_utf8_to_utf32_with_errors(m, u32) == (0, cps)
    u16 = torch.empty(units, dtype=torch.int16, device=device)
    assert b.convert_utf8_to_utf16le_with_errors(m, u16) == (0, units)
    del m
    half = max(3, K // 2)
    rec("next_validate_utf32", timeit(lambda: lib.b200_validate_utf32_with_errors_async(p_(u32), cps, res_p, sp), K), 4 * cps, 0)
    rec("next_utf8_length_from_utf32", timeit(lambda: lib.b200_utf8_length_from_utf32_async(p_(u32), cps, res_p, sp), K), 4 * cps, 0)
    o8 = torch.empty(n8, dtype=torch.uint8, device=device)
    rec("next_convert_utf32_to_utf8", timeit(lambda: lib.b200_convert_utf32_to_utf8_async(p_(u32), cps, p_(o8), res_p, sp), half), 4 * cps, n8)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == n8
    del o8
    o16 = torch.empty(units, dtype=torch.int16, device=device)
    rec("next_convert_utf32_to_utf16le", timeit(lambda: lib.b200_convert_utf32_to_utf16le_async(p_(u32), cps, p_(o16), res_p, sp), half), 4 * cps, 2 * units)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == units
    o32 = torch.empty(cps, dtype=torch.int32, device=device)
    rec("next_convert_utf16le_to_utf32", timeit(lambda: lib.b200_convert_utf16le_to_utf32_async(p_(u16), units, p_(o32), res_p, sp), half), 2 * units, 4 * cps)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == cps
    # §8f rank 4: to_well_formed_utf16le (a map) and detect_encodings (three validations of one buffer)
    rec("next_to_well_formed_utf16le", timeit(lambda: lib.b200_to_well_formed_utf16le_async(p_(u16), units, p_(o16), res_p, sp), K), 2 * units, 2 * units)
    nb16 = 2 * units // 4 * 4
    rec("next_detect_encodings_utf16_text", timeit(lambda: lib.b200_detect_encodings_async(p_(u16), nb16, res_p, sp), half), nb16, 0)
    del u32, u16, o16, o32
    # §8f rank 3: the Latin-1 / ASCII family on 1 GiB of Latin-1 with 30 % bytes >= 0x80
    g = torch.Generator(device=device).manual_seed(59)
    lat = torch.randint(0, 0x80, (GIB,), dtype=torch.uint8, device=device, generator=g)
    lat = torch.where(torch.rand(GIB, device=device, generator=g) < 0.3, lat | 0x80, lat)
    nl = GIB
    asc = lat & 0x7F  # validate_ascii on ASCII (on `lat` it would time the error path: 30 % of the bytes are errors)
    rec("next_validate_ascii", timeit(lambda: lib.b200_validate_ascii_with_errors_async(p_(asc), nl, res_p, sp), K), nl, 0)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == nl
    del asc
    rec("next_utf8_length_from_latin1", timeit(lambda: lib.b200_utf8_length_from_latin1_async(p_(lat), nl, res_p, sp), K), nl, 0)
    n8 = b.utf8_length_from_latin1(lat)
    o8 = torch.empty(n8, dtype=torch.uint8, device=device)
    rec("next_convert_latin1_to_utf8", timeit(lambda: lib.b200_convert_latin1_to_utf8_async(p_(lat), nl, p_(o8), res_p, sp), half), nl, n8)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == n8
    ol = torch.empty(nl, dtype=torch.uint8, device=device)
    rec("next_convert_utf8_to_latin1", timeit(lambda: lib.b200_convert_utf8_to_latin1_async(p_(o8), n8, p_(ol), res_p, sp), half), n8, nl)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == nl
    del o8
    o16 = torch.empty(nl, dtype=torch.int16, device=device)
    rec("next_convert_latin1_to_utf16le", timeit(lambda: lib.b200_convert_latin1_to_utf16le_async(p_(lat), nl, p_(o16), res_p, sp), K), nl, 2 * nl)
    rec("next_convert_utf16le_to_latin1", timeit(lambda: lib.b200_convert_utf16le_to_latin1_async(p_(o16), nl, p_(ol), res_p, sp), K), 2 * nl, nl)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == nl
    del o16
    o32 = torch.empty(nl, dtype=torch.int32, device=device)
    rec("next_convert_latin1_to_utf32", timeit(lambda: lib.b200_convert_latin1_to_utf32_async(p_(lat), nl, p_(o32), res_p, sp), K), nl, 4 * nl)
    rec("next_convert_utf32_to_latin1", timeit(lambda: lib.b200_convert_utf32_to_latin1_async(p_(o32), nl, p_(ol), res_p, sp), K), 4 * nl, nl)
    assert int(d_res[0].item()) & 0xFFFFFFFF == 0 and int(d_res[1].item()) == nl
    # §8f rank 4: the batched front end — one million strings of 64 bytes of the mixed distribution in ONE launch per
    # operation (device flavour), next to the same strings validated one host call at a time (a sample of them)
    nstr, slen = 1 << 20, 64
    one = torch.tensor([0xE4, 0xB8, 0x80] * 10 + [0xC3, 0xA9] * 8 + [0xF0, 0x9F, 0x98, 0x80] * 4 + [0x61, 0x62], dtype=torch.uint8, device=device)
    assert one.numel() == slen  # a valid 64-byte string with 1-, 2-, 3- and 4-byte characters
    packed = one.repeat(nstr).contiguous()
    offs = (torch.arange(nstr + 1, dtype=torch.int64, device=device) * slen).contiguous()
    bres = torch.zeros((nstr, 2), dtype=torch.int64, device=device)
    ms = timeit(lambda: lib.b200_validate_utf8_batch_async(p_(packed), p_(offs), nstr, p_(bres), sp), K)
    out["next_validate_utf8_batch_1Mi_strings_x_64B"] = {"input_gbs": nstr * slen / ms / 1e6, "ms": ms, "strings_per_s": nstr / ms * 1e3,
                                                          "frac_of_peak": nstr * slen / ms / 1e6 / peak}
    bunits = torch.empty(nstr * slen, dtype=torch.int16, device=device)
    ms = timeit(lambda: lib.b200_convert_utf8_to_utf16le_batch_async(p_(packed), p_(offs), nstr, p_(bunits), None, p_(bres), sp), half)
    out["next_convert_utf8_to_utf16le_batch_1Mi_strings_x_64B"] = {"input_gbs": nstr * slen / ms / 1e6, "ms": ms, "strings_per_s": nstr / ms * 1e3}
    hs = packed[: 2000 * slen].cpu().numpy().tobytes()
    hstr = [hs[k * slen:(k + 1) * slen] for k in range(2000)]
    t0 = time.perf_counter()
    for x in hstr:
        b.validate_utf8_with_errors(x)
    one_by_one = (time.perf_counter() - t0) / len(hstr)
    t0 = time.perf_counter()
    b.validate_utf8_batch(hstr)
    out["next_host_validate_utf8_2000_strings_x_64B"] = {"us_per_string_single_calls": one_by_one * 1e6,
                                                          "us_per_string_one_batch_call": (time.perf_counter() - t0) / len(hstr) * 1e6}
    del packed, offs, bres, bunits
    return out


if __name__ == "__main__":
    sys.exit(main())
