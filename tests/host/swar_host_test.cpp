// tests/host/swar_host_test.cpp — CPU emulation of the kernels' per-granule pipeline.
//
// There is no GPU in the build container, so the exact `__host__ __device__` code the kernels run
// (simdutf_b200/csrc/swar.h) is driven here by a sequential stand-in for the CUDA plumbing (granule loads
// with zero filler, neighbour words, per-tile carry) and compared with the oracle (oracle/oracle.c) on
// seeded random and adversarial inputs, at every buffer misalignment.  This is a TEST: the product never
// runs this path.
//
// usage: swar_host_test [iterations] [seed]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <algorithm>
#include <vector>

#include "../../oracle/oracle.h"
#include "../../simdutf_b200/csrc/swar.h"
#include "../../simdutf_b200/csrc/bitplane.h"

using namespace b200;

static std::mt19937_64 rng;
static uint32_t rnd(uint32_t n) { return (uint32_t)(rng() % n); }

// A buffer placed at `misalign` bytes past a 16-byte boundary, viewed as granules with zero filler.
struct View {
  std::vector<uint8_t> mem;  // aligned storage
  uint64_t vbeg, vend;
  View(const uint8_t *data, size_t len, unsigned misalign) {
    vbeg = misalign;
    vend = misalign + len;
    mem.assign(((vend + 15) / 16 + 2) * 16, 0);
    if (len) memcpy(mem.data() + vbeg, data, len);
  }
  uint64_t ngran() const { return (vend + 15) / 16; }
  uint32_t word(long long wi) const {  // zero outside [vbeg, vend)
    uint32_t r = 0;
    for (int b = 0; b < 4; b++) {
      long long pos = wi * 4 + b;
      if (pos >= (long long)vbeg && pos < (long long)vend) r |= (uint32_t)mem[pos] << (8 * b);
    }
    return r;
  }
  void granule(uint64_t g, uint32_t w[4]) const { for (int k = 0; k < 4; k++) w[k] = word((long long)g * 4 + k); }
  uint32_t inrange_word(uint64_t g, int k) const {
    uint32_t m = 0;
    for (int b = 0; b < 4; b++) { uint64_t pos = g * 16 + 4 * k + b; if (pos >= vbeg && pos < vend) m |= 0x80u << (8 * b); }
    return m;
  }
};

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; if (failures < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

static std::string hex(const std::vector<uint8_t> &d) { std::string s; char b[4]; for (uint8_t c : d) { snprintf(b, 4, "%02x", c); s += b; } return s; }

// ---- UTF-8 ---------------------------------------------------------------------------------------------
static void test_utf8(const std::vector<uint8_t> &d, unsigned misalign) {
  const size_t len = d.size();
  View v(d.data(), len, misalign);
  const oracle_result want = oracle_validate_utf8_with_errors(d.data(), len);
  auto at = [&](uint64_t j) -> uint32_t { return d[j]; };
  // min over ALL local verdicts == oracle (the claim the kernels rely on: SURVEY.md A.1)
  {
    int code = 0; uint64_t pos = len;
    for (uint64_t i = 0; i < len; i++) { int c = u8_verdict(at, i, len); if (c) { code = c; pos = i; break; } }
    CHECK(code == want.error && (code == 0 || pos == want.count), "verdict-min mismatch %s", hex(d).c_str());
  }
  // the byte-SWAR count predicates of the reduction kernels (k_utf8.cu K2)
  uint64_t c8 = 0, c16 = 0;
  for (uint64_t g = 0; g < v.ngran(); g++) {
    uint32_t w[4]; v.granule(g, w);
    for (int k = 0; k < 4; k++) {
      uint32_t r = v.inrange_word(g, k);
      c8 += popc(u8_noncont(w[k]) & r);
      c16 += popc((u8_noncont(w[k]) | (u8_ge_f0(w[k]) >> 1)) & (r | (r >> 1)));
    }
  }
  CHECK(c8 == oracle_count_utf8(d.data(), len), "count_utf8 %s", hex(d).c_str());
  CHECK(c16 == oracle_utf16_length_from_utf8(d.data(), len), "utf16_length %s", hex(d).c_str());
}

// ---- UTF-8 -> UTF-16 through the bit-plane transcoder (bitplane.h; k_utf8_to_utf16.cu) --------------------
// Emulates the kernel's plumbing: warp tiles of 32 lane regions of K blocks of 32 bytes, every region processed
// front to back with the carry seeded from the 4 bytes before it.
static uint32_t range_mask32(const View &v, uint64_t b0) {
  uint32_t m = 0;
  for (int p = 0; p < 32; p++) if (b0 + p >= v.vbeg && b0 + p < v.vend) m |= 1u << p;
  return m;
}
static void test_utf8_bitplane(const std::vector<uint8_t> &d, unsigned misalign, int K) {
  const size_t len = d.size();
  View v(d.data(), len, misalign);
  const oracle_result want = oracle_validate_utf8_with_errors(d.data(), len);
  const bool poison = len > 0 && (d[0] & 0xC0) == 0x80;
  std::vector<uint16_t> out;
  std::vector<uint32_t> out32;
  uint64_t best_pos = ~0ull; int best_code = 0; bool any_flag = false;
  auto at = [&](uint64_t j) -> uint32_t { return d[j]; };
  auto vbyte = [&](uint64_t pos) -> uint32_t { return (pos >= v.vbeg && pos < v.vend) ? v.mem[pos] : 0u; };
  const uint64_t region = 32ull * K, nregions = (v.vend + region - 1) / region + 1;  // one extra: truncation shows on the filler
  for (uint64_t r = 0; r < nregions; r++) {
    const uint64_t r0 = r * region;
    bp::Carry carry = bp::carry_from_word(v.word((long long)(r0 / 4) - 1));
    bp::Carry carry32 = carry;
    bp::VCarry vc = bp::vcarry_from_word(v.word((long long)(r0 / 4) - 1));
    uint32_t prev_l4 = carry.l4;
    for (int j = 0; j < K; j++) {
      const uint64_t b0 = r0 + 32ull * j;
      uint32_t B[8];
      for (int i = 0; i < 8; i++) B[i] = v.word((long long)(b0 / 4) + i);
      uint32_t raw[8]; memcpy(raw, B, sizeof raw);
      bp::transpose_in(B);
      for (int k = 0; k < 8; k++) for (int p = 0; p < 32; p++)
        CHECK(((B[k] >> p) & 1) == ((vbyte(b0 + p) >> k) & 1), "transpose_in plane %d pos %d", k, p);
      const uint32_t nb = vbyte(b0 + 32);
      const uint32_t next_nc = (nb & 0xC0) != 0x80;
      uint32_t em = bp::emit16_mask(B, prev_l4, next_nc) & range_mask32(v, b0);
      if (poison) em = 0;
      uint32_t U[16];
      const uint32_t err = bp::utf8_to_utf16_block<true>(B, carry, U);
      const uint32_t err2 = bp::utf8_check_block(B, vc);
      CHECK(err == err2, "check_block differs from transcoder detector");
      prev_l4 = carry.l4;
      bp::transpose_out16(U);
      for (int p = 0; p < 32; p++) if ((em >> p) & 1) out.push_back((uint16_t)(p < 16 ? U[p] : U[p - 16] >> 16));
      uint32_t em32 = bp::emit32_mask(B, next_nc) & range_mask32(v, b0);
      if (poison) em32 = 0;
      uint32_t C[32];
      const uint32_t err3 = bp::utf8_to_utf32_block<true>(B, carry32, C);
      CHECK(err3 == err, "utf32 detector differs");
      bp::transpose_out21(C);
      for (int p = 0; p < 32; p++) if ((em32 >> p) & 1) out32.push_back(C[p]);
      bool flagged = err != 0;
      if (b0 < v.vend && v.vend <= b0 + 32 && len > 0) {
        uint32_t b1 = d[len - 1], b2 = len >= 2 ? d[len - 2] : 0, b3 = len >= 3 ? d[len - 3] : 0;
        flagged = flagged || u8_incomplete_tail(b1, b2, b3);
      }
      if (flagged) {
        any_flag = true;
        long long a = (long long)b0 - 3, b = (long long)b0 + 32;
        if (a < (long long)v.vbeg) a = (long long)v.vbeg;
        if (b > (long long)v.vend) b = (long long)v.vend;
        for (long long p = a; p < b; p++) {
          uint64_t i = (uint64_t)p - v.vbeg;
          int code = u8_verdict(at, i, len);
          if (code) { if (i < best_pos) { best_pos = i; best_code = code; } break; }
        }
      }
    }
  }
  if (want.error == 0) {
    CHECK(!any_flag, "bitplane detector flagged valid input mis=%u K=%d %s", misalign, K, hex(d).c_str());
    std::vector<uint16_t> w16(2 * len + 8);
    oracle_result r16 = oracle_convert_utf8_to_utf16le_with_errors(d.data(), len, w16.data());
    CHECK(out.size() == r16.count && memcmp(out.data(), w16.data(), 2 * r16.count) == 0, "bitplane utf16 output mis=%u K=%d %s", misalign, K, hex(d).c_str());
    std::vector<uint32_t> w32(len + 8);
    oracle_result r32 = oracle_convert_utf8_to_utf32_with_errors(d.data(), len, w32.data());
    CHECK(out32.size() == r32.count && memcmp(out32.data(), w32.data(), 4 * r32.count) == 0, "bitplane utf32 output mis=%u K=%d %s", misalign, K, hex(d).c_str());
  } else {
    CHECK(best_code == want.error && best_pos == want.count, "bitplane error mismatch got (%d,%llu) want (%d,%llu) mis=%u K=%d %s", best_code,
          (unsigned long long)best_pos, want.error, (unsigned long long)want.count, misalign, K, hex(d).c_str());
  }
  CHECK(out32.size() <= oracle_count_utf8(d.data(), len), "bitplane utf32 overrun");
  CHECK(out.size() <= oracle_utf16_length_from_utf8(d.data(), len), "bitplane overrun %zu > %llu mis=%u %s", out.size(),
        (unsigned long long)oracle_utf16_length_from_utf8(d.data(), len), misalign, hex(d).c_str());
}

// ---- UTF-16 --------------------------------------------------------------------------------------------
static void test_utf16(const std::vector<uint16_t> &u, unsigned misalign_units) {
  const size_t len = u.size();
  View v(reinterpret_cast<const uint8_t *>(u.data()), 2 * len, 2 * misalign_units);
  uint64_t cnt = 0, bytes = 0; uint64_t bad_pos = ~0ull;
  std::vector<uint8_t> out;
  for (uint64_t g = 0; g < v.ngran(); g++) {
    uint32_t w[4]; v.granule(g, w);
    uint32_t pw = v.word((long long)g * 4 - 1), nw = v.word((long long)g * 4 + 4);
    for (int i = 0; i < 8; i++) {
      uint64_t pos = g * 16 + 2 * i;
      if (pos < v.vbeg || pos >= v.vend) continue;
      uint32_t x = u16_unit(w, i);
      uint32_t pu = i == 0 ? (pw >> 16) : u16_unit(w, i - 1);
      uint32_t nu = i == 7 ? (nw & 0xFFFF) : u16_unit(w, i + 1);
      cnt += (x & 0xFC00) != 0xDC00;
      bytes += u16_utf8_bytes(x);
      if (u16_bad(x, pu, true, nu, true)) { uint64_t idx = (pos - v.vbeg) / 2; if (idx < bad_pos) bad_pos = idx; }
    }
  }
  CHECK(cnt == oracle_count_utf16le(u.data(), len), "count_utf16le");
  CHECK(bytes == oracle_utf8_length_from_utf16le(u.data(), len), "utf8_length_from_utf16le");
  std::vector<uint8_t> want(3 * len + 8);
  oracle_result r = oracle_convert_utf16le_to_utf8_with_errors(u.data(), len, want.data());
  oracle_result rv = oracle_validate_utf16le_with_errors(u.data(), len);
  if (r.error == 0) {
    CHECK(bad_pos == ~0ull, "utf16 false error");
    CHECK(bytes == r.count, "utf16->utf8 length");
  } else {
    CHECK(bad_pos == r.count, "utf16 error pos got %llu want %llu", (unsigned long long)bad_pos, (unsigned long long)r.count);
    CHECK(rv.error == r.error && rv.count == r.count, "oracle validate16 vs convert");
  }
}


// ---- UTF-16 -> UTF-8 through the bit-plane transcoder (bitplane.h; k_utf16_to_utf8.cu) ------------------------
static void test_utf16_bitplane(const std::vector<uint16_t> &u, unsigned misalign_units) {
  const size_t len = u.size();
  View v(reinterpret_cast<const uint8_t *>(u.data()), 2 * len, 2 * misalign_units);
  auto vunit = [&](long long idx) -> uint32_t {  // virtual unit index (from the aligned base); zero outside
    long long pos = idx * 2;
    if (pos < (long long)v.vbeg || pos >= (long long)v.vend) return 0u;
    return (uint32_t)v.mem[pos] | ((uint32_t)v.mem[pos + 1] << 8);
  };
  std::vector<uint8_t> out;
  uint64_t bad_pos = ~0ull;
  const uint64_t nblocks = (v.vend + 63) / 64 + 1;
  for (uint64_t blk = 0; blk < nblocks; blk++) {
    const long long u0 = (long long)blk * 32;
    uint32_t W[16];
    for (int i = 0; i < 16; i++) W[i] = vunit(u0 + 2 * i) | (vunit(u0 + 2 * i + 1) << 16);
    bp::transpose_in16(W);
    for (int k = 0; k < 16; k++) for (int s = 0; s < 32; s++)
      CHECK(((W[k] >> bp::split_pos(s)) & 1) == ((vunit(u0 + s) >> k) & 1), "transpose_in16 plane %d unit %d", k, s);
    bp::Carry16 c = bp::carry16_from_unit(vunit(u0 - 1));
    uint32_t X[32], e1, e2;
    const uint32_t err = bp::utf16_to_utf8_block(W, c, X, e1, e2);
    bp::transpose_out_n<24>(X);
    for (int s = 0; s < 32; s++) {
      const long long pos = (u0 + s) * 2;
      if (pos < (long long)v.vbeg || pos >= (long long)v.vend) continue;
      const int p = bp::split_pos(s);
      out.push_back((uint8_t)X[p]);
      if ((e1 >> p) & 1) out.push_back((uint8_t)(X[p] >> 8));
      if ((e2 >> p) & 1) out.push_back((uint8_t)(X[p] >> 16));
    }
    bool flagged = err != 0;
    // a buffer that ends with a high surrogate and has no filler unit behind it inside this block
    if (len > 0 && (long long)(v.vend / 2) - 1 >= u0 && (long long)(v.vend / 2) - 1 < u0 + 32 && (u[len - 1] & 0xFC00) == 0xD800) flagged = true;
    if (flagged) {
      for (long long i = u0 - 1; i < u0 + 32; i++) {
        const long long pos = i * 2;
        if (pos < (long long)v.vbeg || pos >= (long long)v.vend) continue;
        const uint64_t idx = (uint64_t)(pos - (long long)v.vbeg) / 2;
        const bool hp = idx > 0, hn = idx + 1 < len;
        if (u16_bad(u[idx], hp ? u[idx - 1] : 0, hp, hn ? u[idx + 1] : 0, hn)) { if (idx < bad_pos) bad_pos = idx; break; }
      }
    }
  }
  std::vector<uint8_t> want(3 * len + 8);
  oracle_result r = oracle_convert_utf16le_to_utf8_with_errors(u.data(), len, want.data());
  CHECK(out.size() == oracle_utf8_length_from_utf16le(u.data(), len), "utf16 bitplane size %zu vs %llu", out.size(), (unsigned long long)oracle_utf8_length_from_utf16le(u.data(), len));
  if (r.error == 0) {
    CHECK(bad_pos == ~0ull, "utf16 bitplane false error at %llu", (unsigned long long)bad_pos);
    CHECK(out.size() == r.count && memcmp(out.data(), want.data(), r.count) == 0, "utf16 bitplane output");
  } else {
    CHECK(bad_pos == r.count, "utf16 bitplane error pos got %llu want %llu", (unsigned long long)bad_pos, (unsigned long long)r.count);
  }
}

// ---- base64 --------------------------------------------------------------------------------------------
static void test_b64(const std::vector<uint8_t> &d, uint64_t options, uint64_t last_chunk) {
  const size_t len = d.size();
  const bool url = options & 1, both = options & 8, garbage = (options == 4 || options == 5 || options == 12);
  // main pass over [0, len): sextets (compacted), first invalid index
  std::vector<uint8_t> sx; uint64_t first_bad = ~0ull;
  for (size_t i = 0; i < len; i++) {
    uint32_t c = b64_class(d[i], url, both);
    if (c <= 63) sx.push_back((uint8_t)c);
    else if (c != 64 && !garbage && first_bad == ~0ull) first_bad = i;
  }
  const uint64_t V = sx.size();
  std::vector<uint8_t> out(b64_bytes_from_sextets(V) + 4);
  for (uint64_t b = 0; b < b64_bytes_from_sextets(V); b++) {
    uint64_t q = b / 3, m = b % 3, r = 4 * q + m;
    out[b] = (uint8_t)((sx[r] << (2 + 2 * m)) | (sx[r + 1] >> (4 - 2 * m)));
  }
  // epilogue exactly as k_base64.cu
  uint64_t srclen = len, equallocation = len; uint32_t equalsigns = 0;
  auto find_last = [&](long long end, bool sextet_only) -> long long {
    for (long long j = end - 1; j >= 0; j--) { uint32_t c = b64_class(d[j], url, both); if (sextet_only ? c <= 63 : c != 64) return j; }
    return -1;
  };
  if (!garbage) {
    srclen = (uint64_t)(find_last((long long)srclen, false) + 1);
    equallocation = srclen;
    if (srclen > 0 && d[srclen - 1] == '=') {
      equallocation = srclen - 1; srclen--; equalsigns = 1;
      srclen = (uint64_t)(find_last((long long)srclen, false) + 1);
      if (srclen > 0 && d[srclen - 1] == '=') { equallocation = srclen - 1; srclen--; equalsigns = 2; }
    }
  }
  int error; uint64_t in_count, out_count;
  const bool invalid = first_bad != ~0ull && first_bad < srclen;
  if (invalid) { error = kInvalidBase64Character; in_count = first_bad; out_count = 0; }
  else {
    uint32_t tail_val[3] = {0, 0, 0}; uint64_t tail_pos[3] = {0, 0, 0};
    long long end = (long long)srclen;
    for (uint32_t k = 0; k < (V & 3) && srclen > 0; k++) { long long f = find_last(end, true); if (f < 0) break; tail_pos[k] = f; tail_val[k] = b64_class(d[f], url, both); end = f; }
    b64_finish(srclen, equalsigns, equallocation, V, garbage, last_chunk, tail_val, tail_pos, &error, &in_count, &out_count);
  }
  std::vector<uint8_t> want(len + 8);
  oracle_full_result r = oracle_base64_to_binary_details(d.data(), len, want.data(), options, last_chunk);
  bool same = r.error == error && r.input_count == in_count;
  if (r.error != kInvalidBase64Character && r.error != kBase64ExtraBits) {
    same = same && r.output_count == out_count && out_count <= b64_bytes_from_sextets(V) && memcmp(out.data(), want.data(), out_count) == 0;
  }
  CHECK(same, "b64 opt=%llu lc=%llu got (%d,%llu,%llu) want (%d,%llu,%llu) in=%s", (unsigned long long)options, (unsigned long long)last_chunk, error,
        (unsigned long long)in_count, (unsigned long long)out_count, r.error, (unsigned long long)r.input_count, (unsigned long long)r.output_count, hex(d).c_str());
}


// ---- base64 classification / sextet values in bit-plane form (bitplane.h; k_base64.cu) -------------------------
static void test_b64_bitplane(const std::vector<uint8_t> &d, uint64_t options) {
  const bool url = options & 1, both = options & 8;
  const uint32_t plus_ok = (both || !url) ? ~0u : 0u, minus_ok = (both || url) ? ~0u : 0u;
  for (size_t b0 = 0; b0 < d.size(); b0 += 32) {
    uint32_t B[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t raw[32] = {0};
    for (size_t i = 0; i < 32 && b0 + i < d.size(); i++) raw[i] = d[b0 + i];
    memcpy(B, raw, 32);
    bp::transpose_in(B);
    uint32_t S[6];
    const bp::B64Class c = bp::base64_classify<true>(B, plus_ok, plus_ok, minus_ok, minus_ok, S);
    uint32_t U[8] = {S[0], S[1], S[2], S[3], S[4], S[5], 0, 0};
    bp::transpose_out8(U);
    uint8_t sx[32];
    memcpy(sx, U, 32);
    for (int p = 0; p < 32; p++) {
      const uint32_t cls = b64_class(raw[p], url, both);
      CHECK(((c.valid >> p) & 1) == (cls <= 63), "b64 plane valid byte %02x opt %llu", raw[p], (unsigned long long)options);
      CHECK(((c.ws >> p) & 1) == (cls == 64), "b64 plane ws byte %02x", raw[p]);
      if (cls <= 63) CHECK(sx[p] == cls, "b64 plane value byte %02x got %u want %u", raw[p], sx[p], cls);
    }
  }
}

// ---- the single-pass decoder's tile pipeline (k_base64.cu: k_b64_decode_v3) with a sequential stand-in for the CUDA
// plumbing: per tile of `tile` characters the sextets are compacted into a staging buffer at alignment zero (16 bytes of
// headroom), the tile's rank goff tells how many sextets of its first quantum precede it (pad = goff & 3, fetched by
// walking the input backwards), groups of 16 sextets are read at byte offset 16 - pad with the kernel's byte-permute
// selectors and packed by swar.h:b64_pack_quanta4, a quantum belongs to the tile that holds its last sextet, and the
// epilogue writes the 1-2 bytes of a trailing partial quantum.  The bytes must equal the oracle's decode.
static void test_b64_single_pass(const std::vector<uint8_t> &d, uint64_t options, size_t tile) {
  const bool url = options & 1, both = options & 8, garbage = (options == 4 || options == 5 || options == 12);
  std::vector<uint8_t> want(d.size() + 8);
  const oracle_full_result r = oracle_base64_to_binary_details(d.data(), d.size(), want.data(), options, 0);
  if (r.error != kSuccess) return;  // output is pinned for successful decodes only
  std::vector<uint8_t> out(d.size() + 64, 0xEE);
  uint64_t goff = 0;
  for (size_t t0 = 0; t0 < d.size(); t0 += tile) {
    std::vector<uint8_t> stage(16 + tile + 32, 0xAA);  // stale bytes behind the sextets, as in shared memory
    uint32_t wtot = 0;
    for (size_t i = t0; i < std::min(d.size(), t0 + tile); i++) {
      const uint32_t c = b64_class(d[i], url, both);
      if (c <= 63) stage[16 + wtot++] = (uint8_t)c;
    }
    const uint32_t pad = (uint32_t)(goff & 3);
    if (wtot) {
      uint32_t need = pad;
      for (long long pos = (long long)t0 - 1; need && pos >= 0; pos--) {
        const uint32_t c = b64_class(d[pos], url, both);
        if (c <= 63) stage[16 - pad + (--need)] = (uint8_t)c;
      }
      CHECK(need == 0, "b64 single pass: carried sextets missing");
      const uint32_t have = pad + wtot, nq = have >> 2, nb = 3 * nq;
      const uint32_t ngroups = (nq + 3) >> 2;
      const uint32_t sel = 0x3210u + 0x1111u * (4u - pad);
      std::vector<uint8_t> so(12 * ngroups + 16);
      for (uint32_t g = 0; g < ngroups; g++) {
        uint32_t x[5];
        memcpy(x, &stage[16 + 16 * g - 4], 20);
        uint32_t w[4], o3[3];
        for (int k = 0; k < 4; k++) w[k] = prmt(x[k], x[k + 1], sel) & 0x3F3F3F3Fu;
        b64_pack_quanta4(w, o3);
        memcpy(&so[12 * g], o3, 12);
      }
      memcpy(&out[3 * (goff >> 2)], so.data(), nb);
    }
    goff += wtot;
  }
  const uint64_t V = goff;
  if ((V & 3) >= 2) {  // the epilogue's tail bytes: tail_val[0] is the stream's last sextet
    uint32_t tv[3] = {0, 0, 0};
    long long end = (long long)d.size();
    if (!garbage) { while (end > 0 && (b64_class(d[end - 1], url, both) == 64 || d[end - 1] == '=')) end--; }
    for (uint32_t k = 0; k < (V & 3); k++) {
      long long f = end - 1;
      while (f >= 0 && b64_class(d[f], url, both) > 63) f--;
      if (f < 0) break;
      tv[k] = b64_class(d[f], url, both);
      end = f;
    }
    uint8_t *tail = &out[3 * (V >> 2)];
    if ((V & 3) == 2) tail[0] = (uint8_t)((tv[1] << 2) | (tv[0] >> 4));
    else { tail[0] = (uint8_t)((tv[2] << 2) | (tv[1] >> 4)); tail[1] = (uint8_t)((tv[1] << 4) | (tv[0] >> 2)); }
  }
  CHECK(memcmp(out.data(), want.data(), r.output_count) == 0, "b64 single pass: output differs, tile %zu opt %llu in=%s", tile,
        (unsigned long long)options, hex(d).c_str());
}

// ---- generators ----------------------------------------------------------------------------------------
static const uint8_t kSpecial[] = {0x20, 0x41, 0x7F, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED, 0xEE, 0xEF, 0xF0, 0xF1, 0xF3, 0xF4, 0xF5, 0xF7, 0xF8, 0xFF};
static void push_cp(std::vector<uint8_t> &o, uint32_t cp) {
  if (cp < 0x80) o.push_back(cp);
  else if (cp < 0x800) { o.push_back(0xC0 | cp >> 6); o.push_back(0x80 | (cp & 63)); }
  else if (cp < 0x10000) { o.push_back(0xE0 | cp >> 12); o.push_back(0x80 | (cp >> 6 & 63)); o.push_back(0x80 | (cp & 63)); }
  else { o.push_back(0xF0 | cp >> 18); o.push_back(0x80 | (cp >> 12 & 63)); o.push_back(0x80 | (cp >> 6 & 63)); o.push_back(0x80 | (cp & 63)); }
}
static uint32_t rand_cp() {
  switch (rnd(6)) {
    case 0: return rnd(0x80);
    case 1: return 0x80 + rnd(0x780);
    case 2: { uint32_t c = 0x800 + rnd(0xF800); return (c >= 0xD800 && c < 0xE000) ? 0x4E00 : c; }
    case 3: return 0x10000 + rnd(0x100000);
    case 4: { static const uint32_t edge[] = {0x7F, 0x80, 0x7FF, 0x800, 0xD7FF, 0xE000, 0xFFFF, 0x10000, 0x10FFFF, 0xFFF, 0x1000, 0x3FFFF, 0x40000}; return edge[rnd(13)]; }
    default: return 0x20 + rnd(0x5F);
  }
}
static std::vector<uint8_t> gen_utf8(size_t n) {
  std::vector<uint8_t> o;
  const uint32_t mode = rnd(5);
  if (mode == 0) { for (size_t i = 0; i < n; i++) o.push_back(kSpecial[rnd(sizeof(kSpecial))]); return o; }
  if (mode == 1) { for (size_t i = 0; i < n; i++) o.push_back((uint8_t)rnd(256)); return o; }
  while (o.size() < n) push_cp(o, rand_cp());
  if (mode == 3 && !o.empty()) { for (uint32_t k = 0, e = 1 + rnd(2); k < e; k++) o[rnd((uint32_t)o.size())] = kSpecial[rnd(sizeof(kSpecial))]; }
  if (mode == 4 && !o.empty()) o.resize(o.size() - rnd((uint32_t)std::min<size_t>(o.size(), 4)));  // truncate mid-character
  return o;
}
static std::vector<uint16_t> gen_utf16(size_t n) {
  std::vector<uint16_t> o;
  const uint32_t mode = rnd(3);
  static const uint16_t edge[] = {0x41, 0x7F, 0x80, 0x7FF, 0x800, 0xD7FF, 0xD800, 0xDBFF, 0xDC00, 0xDFFF, 0xE000, 0xFFFF};
  while (o.size() < n) {
    if (mode == 0) { o.push_back(edge[rnd(12)]); continue; }
    uint32_t cp = rand_cp();
    if (cp < 0x10000) o.push_back((uint16_t)cp);
    else { cp -= 0x10000; o.push_back(0xD800 + (cp >> 10)); o.push_back(0xDC00 + (cp & 0x3FF)); }
    if (mode == 2 && rnd(40) == 0) o.push_back(0xD800 + rnd(0x800));
  }
  if (rnd(4) == 0 && !o.empty()) o.pop_back();
  return o;
}
static std::vector<uint8_t> gen_b64(size_t n) {
  static const char abc[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/-_";
  static const char ws[] = " \t\n\r\f";
  static const char junk[] = "=*\x80\xff\x00.\x0b";
  std::vector<uint8_t> o;
  const uint32_t mode = rnd(4);
  for (size_t i = 0; i < n; i++) {
    uint32_t x = rnd(100);
    if (x < 80) o.push_back(mode == 1 ? abc[rnd(62) < 60 ? rnd(62) : 64 + rnd(2)] : abc[rnd(64)]);
    else if (x < 93) o.push_back(ws[rnd(5)]);
    else if (x < 97 && mode >= 2) o.push_back(rnd(2) ? junk[rnd(7)] : abc[62 + rnd(4)]);
    else o.push_back(abc[rnd(62)]);
  }
  static const char *tails[] = {"", "=", "==", " = = ", "= ", "=\n=", "===", " ", "  \n", "= =="};
  const char *t = tails[rnd(10)];
  for (; *t; t++) o.push_back(*t);
  return o;
}

// ---- SWAR screens of the widened rows (swar.h: u16_pairing_screen, u8l1_screen64, l1_high_count64) ---------------
// u16_pairing_screen over 8 units with both neighbours must be zero exactly when none of the 8 units is a bad
// surrogate (u16_bad); a lone low surrogate right behind or a lone high surrogate right in front of the group may
// also raise it (conservative: those are reported by their own groups).
static void test_u16_screen() {
  static const uint16_t pool[] = {0x0041, 0x4E2D, 0xD7FF, 0xD800, 0xDBFF, 0xDC00, 0xDFFF, 0xE000, 0xFFFD};
  uint16_t u[10];
  for (auto &x : u) x = rnd(3) ? pool[rnd(9)] : (uint16_t)(0xD800 + rnd(0x800));
  if (rnd(4) == 0) for (int i = 1; i + 1 < 10; i += 2) { u[i] = (uint16_t)(0xD800 + rnd(0x400)); u[i + 1] = (uint16_t)(0xDC00 + rnd(0x400)); }
  uint32_t w[4];
  for (int k = 0; k < 4; k++) w[k] = (uint32_t)u[1 + 2 * k] | ((uint32_t)u[2 + 2 * k] << 16);
  const uint32_t pw = ((uint32_t)u[0] << 16) | rnd(0x10000), nw = (uint32_t)u[9] | (rnd(0x10000) << 16);
  bool any_bad = false;
  for (int i = 1; i <= 8; i++) any_bad = any_bad || b200::u16_bad(u[i], u[i - 1], true, u[i + 1], true);
  const bool next_lone_low = ((u[9] & 0xFC00) == 0xDC00 && (u[8] & 0xFC00) != 0xD800) ||
                             ((u[0] & 0xFC00) == 0xD800 && (u[1] & 0xFC00) != 0xDC00);  // or: the unit before is a lone high
  const uint32_t wrong = b200::u16_pairing_screen(w, pw, nw);
  CHECK(!(any_bad && wrong == 0), "u16 screen missed a bad surrogate");
  CHECK(!(wrong != 0 && !any_bad && !next_lone_low), "u16 screen raised without cause %08x: %04x | %04x %04x %04x %04x %04x %04x %04x %04x | %04x", wrong,
        u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7], u[8], u[9]);  // the two-tag form (to_well_formed_utf16) gives the same answer
  CHECK((b200::u16_pairing_screen_tags(w, pw, nw) != 0) == (wrong != 0), "the two forms of the u16 screen disagree");
}
// u8l1_screen64: the continuation count is exact; suspect == false implies the oracle's walk over the 64 bytes
// (with the byte before and after) reports nothing inside them.
static void test_u8l1_screen() {
  static const uint8_t pool[] = {0x00, 0x41, 0x7F, 0x80, 0xA9, 0xBF, 0xC0, 0xC1, 0xC2, 0xC3, 0xC4, 0xDF, 0xE0, 0xEF, 0xF0, 0xF7, 0xF8, 0xFF};
  uint8_t b[66];
  const int mode = rnd(4);
  for (int i = 0; i < 66;) {
    if (mode == 0) { b[i++] = pool[rnd(18)]; continue; }
    if (rnd(3) == 0 && i + 1 < 66) { b[i++] = (uint8_t)(0xC2 + rnd(2)); b[i++] = (uint8_t)(0x80 + rnd(64)); }
    else b[i++] = (uint8_t)rnd(0x80);
  }
  if (mode == 1) b[rnd(66)] = pool[rnd(18)];
  if (mode == 2) { const int k = rnd(66); b[k] = (uint8_t)rnd(256); }
  uint32_t w[16];
  for (int j = 0; j < 16; j++) w[j] = (uint32_t)b[1 + 4 * j] | ((uint32_t)b[2 + 4 * j] << 8) | ((uint32_t)b[3 + 4 * j] << 16) | ((uint32_t)b[4 + 4 * j] << 24);
  bool suspect = false;
  const uint32_t conts = b200::u8l1_screen64(w, b[0], b[65], &suspect);
  uint32_t want_conts = 0, want_high = 0;
  for (int i = 1; i <= 64; i++) { want_conts += (b[i] & 0xC0) == 0x80; want_high += b[i] >> 7; }
  CHECK(conts == want_conts, "u8l1 continuation count %u vs %u", conts, want_conts);
  CHECK(b200::l1_high_count64(w) == want_high, "l1 high count");
  // the word-at-a-time emitters against the byte rules
  {
    std::vector<uint8_t> l8;
    for (int j = 0; j < 16; j++) {
      uint32_t f, c;
      const uint32_t hi = b200::l1u8_word(w[j], &f, &c);
      for (int k = 0; k < 4; k++) {
        l8.push_back((uint8_t)(f >> (8 * k)));
        if ((hi >> (8 * k)) & 1u) l8.push_back((uint8_t)(c >> (8 * k)));
      }
    }
    std::vector<uint8_t> want8(140);
    const uint64_t n8 = oracle_convert_latin1_to_utf8(b + 1, 64, want8.data());
    CHECK(l8.size() == n8 && std::equal(l8.begin(), l8.end(), want8.begin()), "l1u8_word");
  }
  if (!suspect && b[0] < 0xE0) {
    std::vector<uint8_t> got;
    for (int j = 0; j < 16; j++) {
      uint32_t keep7;
      const uint32_t o = b200::u8l1_word(w[j], j < 15 ? w[j + 1] : (uint32_t)b[65], &keep7);
      for (int k = 0; k < 4; k++) if ((keep7 >> (8 * k + 7)) & 1u) got.push_back((uint8_t)(o >> (8 * k)));
    }
    // the oracle on the characters that START inside the 64 bytes (a leading continuation byte belongs to the lane before;
    // a trailing C2/C3 lead takes its second byte from b[65])
    const int from = ((b[1] & 0xC0) == 0x80) ? 2 : 1;
    const int to = ((b[64] & 0xE0) == 0xC0) ? 66 : 65;
    std::vector<uint8_t> want(80);
    const oracle_result rr = oracle_convert_utf8_to_latin1_with_errors(b + from, to - from, want.data());
    CHECK(rr.error == ORACLE_SUCCESS && rr.count == got.size() && std::equal(got.begin(), got.end(), want.begin()),
          "u8l1_word: %d, %llu vs %zu", (int)rr.error, (unsigned long long)rr.count, got.size());
  }
  // the reference's walk over all 66 bytes: the first error, if any, must not lie in [1, 64] unless suspect
  std::vector<uint8_t> out(80);
  // start the walk on a character boundary: if b[0] is a lead, include it; if it is a lone continuation the walk errs at 0
  const oracle_result r = oracle_convert_utf8_to_latin1_with_errors(b, 66, out.data());
  // (a byte >= 0xE0 in front is an error of the lane before, which sorts first: nothing to prove about this lane then)
  if (!suspect && b[0] < 0xE0) {
    // walk again from byte 1 when byte 0 itself is the problem (it belongs to the previous lane)
    oracle_result r1 = r;
    if (r.error != ORACLE_SUCCESS && r.count == 0) {
      const int skip = ((b[0] & 0xE0) == 0xC0 && (b[1] & 0xC0) == 0x80) ? 2 : 1;
      r1 = oracle_convert_utf8_to_latin1_with_errors(b + skip, 66 - skip, out.data());
      if (r1.error != ORACLE_SUCCESS) r1.count += skip;
    }
    CHECK(r1.error == ORACLE_SUCCESS || r1.count > 64, "u8l1 screen missed error %d at %llu", (int)r1.error, (unsigned long long)r1.count);
  }
}

int main(int argc, char **argv) {
  const long iters = argc > 1 ? atol(argv[1]) : 20000;
  rng.seed(argc > 2 ? atoll(argv[2]) : 12345);
  for (long it = 0; it < iters; it++) {
    const size_t n8 = rnd(4) ? rnd(120) : rnd(700);
    std::vector<uint8_t> d = gen_utf8(n8);
    test_utf8(d, rnd(16));
    test_utf8_bitplane(d, rnd(16), 1 + rnd(4));
    std::vector<uint16_t> u = gen_utf16(rnd(4) ? rnd(60) : rnd(300));
    test_utf16(u, rnd(8));
    test_utf16_bitplane(u, rnd(8));
    std::vector<uint8_t> b = gen_b64(rnd(4) ? rnd(100) : rnd(400));
    static const uint64_t opts[] = {0, 1, 2, 3, 4, 5, 8, 12};
    test_b64(b, opts[rnd(8)], rnd(3));
    test_b64_bitplane(b, opts[rnd(8)]);
    { static const size_t tiles[] = {1, 3, 4, 5, 16, 17, 32, 64}; test_b64_single_pass(b, opts[rnd(8)], tiles[rnd(8)]); }
    if (it < 256) { std::vector<uint8_t> all(256); for (int i = 0; i < 256; i++) all[i] = (uint8_t)(i + it); test_b64_bitplane(all, opts[it & 7]); }
    for (int k = 0; k < 8; k++) { test_u16_screen(); test_u8l1_screen(); }
    if (failures > 50) break;
  }
  // known-answer edge cases
  test_utf8({}, 0);
  test_utf8({0x80}, 3);
  for (unsigned mis = 0; mis < 16; mis++) {
    std::vector<uint8_t> d(64, 0x20); d.push_back(0xFF); test_utf8(d, mis); test_utf8_bitplane(d, mis, 2);
    std::vector<uint8_t> e(64, 0x20); e.push_back(0xA9); test_utf8(e, mis);
    for (int cut = 1; cut <= 3; cut++) { std::vector<uint8_t> f(29 + mis, 'a'); push_cp(f, 0x1F600); f.resize(f.size() - cut); test_utf8(f, mis); test_utf8_bitplane(f, mis, 1 + mis % 4); }
  }
  printf("swar_host_test: %ld iterations, %d failures\n", iters, failures);
  return failures ? 1 : 0;
}
