// sp_device.cuh — device-side plumbing of the SINGLE-PASS transcoders (k_utf8_to_utf16.cu, k_utf16_to_utf8.cu): a
// persistent grid of CTAs of NW worker warps + one scan warp takes CTA-tiles in increasing order through an atomic
// ticket; the workers hand their warp totals to the scan warp through rings in shared memory, the scan warp resolves
// the tile's global output offset with a decoupled look-back over epoch-tagged descriptors and hands every worker its
// offset back.  Why each hand-off is the way it is (who publishes the aggregate, who takes the ticket, why the look-back
// loads are coalesced) is written up, with the measurements, in the header of k_utf8_transcode_v3 and in DESIGN.md §4.
#pragma once
#include "bp_device.cuh"
#include "device_common.cuh"

namespace b200 {
namespace sp {

// Hand-offs go through mbarrier objects in shared memory (one arrival releases any number of waiters, and waiters do not
// wait for EACH OTHER the way the threads of a bar.sync do): a worker that is ahead never waits for a slower worker.
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {  // release: this thread's earlier shared-memory writes are visible to waiters
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_hint(uint32_t addr, uint32_t parity) {  // acquire; the hardware parks the thread
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
  } while (!done);
}
// The scan warp's flavour: poll, then really sleep.  try_wait's suspend-time hint compiles to NANOSLEEP.SYNCS, which any
// mbarrier traffic of the SM ends at once: ncu counted ~480 polls per tile and scan warp, 4 instructions each (23 % of all
// warp-instructions of the base64 decoder, 32 % of UTF-16 -> UTF-8).  They turned out to be issue slots nobody else
// wanted — a plain nanosleep between polls changes little: measured on B200 with 0 / 256 / 512 / 1024 ns, ms per GiB:
// UTF-8 -> UTF-32 1.045 / 1.026 / 1.011 / 0.998, UTF-8 -> UTF-16 0.786 / 0.782 / 0.780 / 0.781, the 7-worker kernels
// (UTF-16 -> UTF-8, base64, UTF-32 sources) within 0.3 %.  512 ns: half a microsecond more for the look-back, which the
// two-tile deferral of the copy-out absorbs, and the profile counts instructions that do work.
constexpr uint32_t kScanSleepNs = 512u;
__device__ __forceinline__ void mbar_wait_sleep(uint32_t addr, uint32_t parity) {
  for (;;) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (done) break;
    __nanosleep(kScanSleepNs);
  }
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint32_t r;
  asm volatile("{ .reg .b16 t; ld.shared.b16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void stg_cs_u32(void *p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_u32(unsigned int *p, uint32_t v) {
  uint32_t r;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
}

// Hand-off rings, slot i & 3 for the CTA's i-th tile.
struct Rings {
  uint32_t tot[4][32];                          // workers -> scan warp: the warp totals
  unsigned long long goff[4][32];               // scan warp -> workers: every worker's global output offset
  uint32_t ticket[4];                           // the CTA-tile index
  uint32_t acc[4];                              // arrivals << 24 | sum of the warp totals
  alignas(8) unsigned long long mbar[12];       // [0,4) ticket posted, [4,8) offsets posted, [8,12) totals in
};
constexpr uint32_t kMbTicket = 0u, kMbGoff = 32u, kMbTotals = 64u;  // byte offsets into Rings::mbar

__device__ __forceinline__ void init_rings(Rings &rg, int nw) {  // one thread, before the CTA's first __syncthreads
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(rg.mbar);
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    mbar_init(mb + kMbTicket + 8u * k, 1);
    mbar_init(mb + kMbGoff + 8u * k, 1);
    mbar_init(mb + kMbTotals + 8u * k, (uint32_t)nw);
    rg.acc[k] = 0;
  }
}

// The scan warp's whole life (all 32 lanes).  `dbg` / `ts`: clock64 / globaltimer instrumentation (tools/dbg_timing.py),
// nullptr in the product instantiations.
template <int NW, int AHEAD>
__device__ __forceinline__ void scan_warp(Rings &rg, unsigned long long *desc, uint32_t epoch, uint32_t num_cta_tiles,
                                          Scratch *scr, unsigned long long *dbg, unsigned long long *ts) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(rg.mbar);
  // the CTA's first AHEAD tiles; every later one is reserved by a worker, AHEAD tiles before it is due (AHEAD = 2
  // removes the workers' 0.4 us wait for the ticket and costs as much again on the offsets: 0.884 vs 0.867 ms per GiB)
  uint32_t t = 0;
  if (lane < (unsigned)AHEAD) {
    t = atomicAdd(&scr->ticket, 1u);
    rg.ticket[lane] = t;
    mbar_arrive(mb + kMbTicket + 8u * lane);
  }
  long long dbg_wait = 0, dbg_lb = 0, dbg_lbmax = 0, dbg_polls = 0, dbg_n = 0, dbg_late = 0, dbg_seen = 0, dbg_start = 0;
  for (uint32_t iter = 0;; iter++) {
    const uint32_t slot = iter & 3u, ph = (iter >> 2) & 1u;
    mbar_wait_sleep(mb + kMbTicket + 8u * slot, ph);
    t = rg.ticket[slot];
    if (t >= num_cta_tiles) break;
    const long long c0 = dbg ? clock64() : 0;
    mbar_wait_sleep(mb + kMbTotals + 8u * slot, ph);
    const long long c1 = dbg ? clock64() : 0;
    unsigned long long now0 = 0;
    if (dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now0));
    const uint32_t mine = lane < (unsigned)NW ? rg.tot[slot][lane] : 0u;
    const uint32_t incl = bpd::warp_inclusive_u32(mine);
    const uint32_t agg = __shfl_sync(kFull, incl, 31);
    unsigned long long sum = 0;
    if (t == 0) {
      if (lane == 0) st_relaxed_u64(desc, desc_pack(epoch, kStatusPrefix, 0, agg));
    } else {
      // Look-back, COALESCED: descriptor base - 32 j - lane goes to lane `lane` of load j, so a warp load touches 256
      // contiguous bytes (8 sectors), and a lane polls only a descriptor that is not ready yet.  (A persistent grid
      // of equal tiles drifts into lockstep: then no predecessor of the current wave has its prefix yet and every
      // CTA reads the whole in-flight window, G descriptors G times per wave, on a few dozen L2 lines.)
      constexpr int kR = 4;
      long long base = (long long)t - 1;
      unsigned long long d[kR];
#pragma unroll
      for (int j = 0; j < kR; j++) {
        const long long idx = base - 32ll * j - (long long)lane;
        d[j] = idx >= 0 ? ld_relaxed_u64(desc + idx) : desc_pack(epoch, kStatusPrefix, 0, 0);
      }
      bool done = false;
      while (!done) {
#pragma unroll
        for (int j = 0; j < kR; j++) {
          if (!done) {  // warp-uniform
            const long long idx = base - 32ll * j - (long long)lane;
            uint32_t spins = 0;
            while (__any_sync(kFull, desc_epoch(d[j]) != epoch || desc_status(d[j]) == 0u)) {
              if (desc_epoch(d[j]) != epoch || desc_status(d[j]) == 0u) d[j] = ld_relaxed_u64(desc + idx);  // idx >= 0 here
              dbg_polls++;
              if (++spins > (1u << 24)) {  // cannot happen (tickets are handed out in order); never hang the device on a logic error
                report_error(scr, err_key(0, kOther));
                break;
              }
            }
            if (dbg && j == 0 && base == (long long)t - 1 && t >= 64) {
              unsigned long long now;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
              const long long mine_ts = (long long)ld_relaxed_u64(ts + t), pred_ts = (long long)ld_relaxed_u64(ts + idx);
              long long late = pred_ts - mine_ts;  // > 0: this predecessor published after me
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                const long long x = __shfl_xor_sync(kFull, late, o);
                late = x > late ? x : late;
              }
              dbg_late += late; dbg_seen += (long long)now - (mine_ts + (late > 0 ? late : 0));
              dbg_start += (long long)now0 - mine_ts;
            }
            const unsigned pm = __ballot_sync(kFull, desc_status(d[j]) == kStatusPrefix);
            const unsigned first = pm ? (unsigned)(__ffs((int)pm) - 1) : 32u;
            sum += (unsigned long long)__reduce_add_sync(kFull, lane < first ? (uint32_t)desc_value(d[j]) : 0u);
            if (pm) {
              const uint32_t lo = __shfl_sync(kFull, (uint32_t)desc_value(d[j]), first);
              const uint32_t hi = __shfl_sync(kFull, (uint32_t)(desc_value(d[j]) >> 32), first);
              sum += ((unsigned long long)hi << 32) | lo;
              done = true;
            }
          }
        }
        if (!done) {
          base -= 32 * kR;
#pragma unroll
          for (int j = 0; j < kR; j++) {
            const long long idx = base - 32ll * j - (long long)lane;
            d[j] = idx >= 0 ? ld_relaxed_u64(desc + idx) : desc_pack(epoch, kStatusPrefix, 0, 0);
          }
        }
      }
      if (lane == 0) st_relaxed_u64(desc + t, desc_pack(epoch, kStatusPrefix, 0, sum + agg));
    }
    if (lane < (unsigned)NW) rg.goff[slot][lane] = sum + (incl - mine);
    __syncwarp();
    if (lane == 0) mbar_arrive(mb + kMbGoff + 8u * slot);
    if (dbg) {
      const long long c2 = clock64();
      dbg_wait += c1 - c0; dbg_lb += c2 - c1; dbg_lbmax = (c2 - c1) > dbg_lbmax ? (c2 - c1) : dbg_lbmax; dbg_n++;
    }
  }
  if (dbg && lane == 0) {
    unsigned long long *o = dbg + 16ull * blockIdx.x;
    o[0] = dbg_n; o[1] = dbg_wait; o[2] = dbg_lb; o[3] = dbg_lbmax; o[4] = dbg_polls; o[5] = dbg_late; o[6] = dbg_seen; o[7] = dbg_start;
  }
}


// ---- worker side ----
// Hands the warp total of this warp's tile (the CTA's tile `ct`, ring slot `slot`) to the scan warp.  The worker whose
// total arrives LAST publishes the CTA aggregate itself (it never waits for the scan warp) and reserves the CTA's next
// tile: returns true (warp-uniformly) in that warp, the ticket request in flight in lane 0's `tn`.
template <int NW>
__device__ __forceinline__ bool post_totals(Rings &rg, uint32_t slot, unsigned warp, unsigned lane, uint32_t wtot, uint32_t ct,
                                            unsigned long long *desc, uint32_t epoch, Scratch *scr, uint32_t &tn) {
  bool took = false;
  if (lane == 0) {
    rg.tot[slot][warp] = wtot;
    const uint32_t old = atomicAdd(&rg.acc[slot], (1u << 24) | wtot);
    if ((old >> 24) == (uint32_t)NW - 1u) {
      tn = atom_add_u32(&scr->ticket, 1u);
      took = true;
      rg.acc[slot] = 0;
      st_relaxed_u64(desc + ct, desc_pack(epoch, kStatusAggregate, 0, (old & 0xFFFFFFu) + wtot));
    }
    mbar_arrive((uint32_t)__cvta_generic_to_shared(rg.mbar) + kMbTotals + 8u * slot);
  }
  return __any_sync(kFull, took);
}
// Posts the reserved ticket as the CTA's tile of iteration `iter_next`; returns it (warp-uniform).
__device__ __forceinline__ uint32_t post_ticket(Rings &rg, uint32_t iter_next, uint32_t tn, unsigned lane) {
  tn = __shfl_sync(kFull, tn, 0);
  if (lane == 0) {
    rg.ticket[iter_next & 3u] = tn;
    mbar_arrive((uint32_t)__cvta_generic_to_shared(rg.mbar) + kMbTicket + 8u * (iter_next & 3u));
  }
  return tn;
}
__device__ __forceinline__ uint32_t wait_ticket(Rings &rg, uint32_t iter) {
  mbar_wait_hint((uint32_t)__cvta_generic_to_shared(rg.mbar) + kMbTicket + 8u * (iter & 3u), (iter >> 2) & 1u);
  return rg.ticket[iter & 3u];
}
// The global output offset of this warp's tile of iteration `iter` (waits for the scan warp; a worker never runs ahead
// of the offsets ring, whether or not it has anything to copy).
__device__ __forceinline__ unsigned long long wait_goff(Rings &rg, uint32_t iter, unsigned warp) {
  mbar_wait_hint((uint32_t)__cvta_generic_to_shared(rg.mbar) + kMbGoff + 8u * (iter & 3u), (iter >> 2) & 1u);
  return rg.goff[iter & 3u][warp];
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t r;
  asm volatile("{ .reg .b16 t; ld.shared.u8 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ldg_v8(const void *p, uint32_t *r) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}

// staging (bytes [0, n) at alignment zero) -> dst[0 .. n)
__device__ __forceinline__ void copy_out_bytes(uint32_t stage_addr, uint32_t n, uint8_t *dst, unsigned lane) {
  uint32_t h = (uint32_t)(0u - (uint32_t)reinterpret_cast<uintptr_t>(dst)) & 3u;  // bytes in front of the first aligned word
  if (h > n) h = n;
  if (lane < h) dst[lane] = (uint8_t)lds_u8(stage_addr + lane);
  const uint32_t nw = (n - h) >> 2;
  uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + h);
  const uint32_t sel = 0x3210u + 0x1111u * h;  // word w of the destination = staging bytes [h + 4w, h + 4w + 4)
#pragma unroll 4
  for (uint32_t w = lane; w < nw; w += 32u) {
    const uint32_t a = lds_u32(stage_addr + 4u * w), b = lds_u32(stage_addr + 4u * w + 4u);
    stg_cs_u32(d32 + w, __byte_perm(a, b, sel));
  }
  const uint32_t done = h + 4u * nw;
  if (lane < n - done) dst[done + lane] = (uint8_t)lds_u8(stage_addr + done + lane);
}

// staging (bytes [0, n) at alignment zero, 16-byte-aligned buffer with >= 16 bytes of slack behind) -> dst[0 .. n).
// The destination's own 16-byte vectors: vector v of gbase = dst - mis holds staging bytes [16 v - mis, 16 v - mis + 16),
// i.e. two 128-bit shared loads, four byte permutes and one 128-bit streaming store per 16 bytes (the first version moved
// 32-bit words: two shared loads, a permute and a store per FOUR bytes — ncu: 11 % of the stall samples of UTF-16 ->
// UTF-8).  The partial first and last vectors go bytewise, lanes 0-15 and 16-31.
__device__ __forceinline__ void copy_out_bytes_v16(uint32_t stage_addr, uint32_t n, uint8_t *dst, unsigned lane) {
  const uint32_t mis = (uint32_t)reinterpret_cast<uintptr_t>(dst) & 15u;
  uint8_t *gbase = dst - mis;
  const uint32_t end = mis + n;
  const uint32_t v_lo = (mis + 15u) >> 4, v_hi = end >> 4;  // whole vectors [v_lo, v_hi)
  uint4 *gv = reinterpret_cast<uint4 *>(gbase);
  if (mis == 0u) {
    for (uint32_t v = lane; v < v_hi; v += 32u) {
      uint4 o;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(stage_addr + 16u * v));
      stg_stream_v4(gv + v, o);
    }
  } else {
    const uint32_t sh = 16u - mis;                       // staging byte offset of vector v is 16 (v - 1) + sh
    const uint32_t wsel = sh >> 2;
    const uint32_t psel = 0x3210u + 0x1111u * (sh & 3u);
    for (uint32_t v = v_lo + lane; v < v_hi; v += 32u) {
      uint4 lo, hi;
      const uint32_t a = stage_addr + 16u * (v - 1u);
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(a));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(a + 16u));
      uint32_t t0, t1, t2, t3, t4;
      switch (wsel) {  // warp-uniform
        case 0: t0 = lo.x; t1 = lo.y; t2 = lo.z; t3 = lo.w; t4 = hi.x; break;
        case 1: t0 = lo.y; t1 = lo.z; t2 = lo.w; t3 = hi.x; t4 = hi.y; break;
        case 2: t0 = lo.z; t1 = lo.w; t2 = hi.x; t3 = hi.y; t4 = hi.z; break;
        default: t0 = lo.w; t1 = hi.x; t2 = hi.y; t3 = hi.z; t4 = hi.w; break;
      }
      uint4 o;
      o.x = __byte_perm(t0, t1, psel);
      o.y = __byte_perm(t1, t2, psel);
      o.z = __byte_perm(t2, t3, psel);
      o.w = __byte_perm(t3, t4, psel);
      stg_stream_v4(gv + v, o);
    }
  }
  const uint32_t head_end = 16u * v_lo < end ? 16u * v_lo : end;
  if (lane < 16u) {
    const uint32_t e = mis + lane;
    if (e < head_end) gbase[e] = (uint8_t)lds_u8(stage_addr + lane);
  } else if (v_hi >= v_lo) {
    const uint32_t e = 16u * v_hi + (lane - 16u);
    if (e >= mis && e < end) gbase[e] = (uint8_t)lds_u8(stage_addr + e - mis);
  }
}


}  // namespace sp
}  // namespace b200
