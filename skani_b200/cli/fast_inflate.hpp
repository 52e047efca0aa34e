// fast_inflate.hpp -- raw DEFLATE (RFC 1951) decoder for the FASTA/FASTQ ingestion path of the host CLI.
//
// Why: reading `.fa.gz` inputs is bounded by inflate (the reference uses flate2 behind
// needletail, src/file_io.rs:141-362).  zlib's inflate keeps its state machine resumable at every byte, which costs it most
// of its speed; here the whole compressed file and the whole output buffer are in memory (the file is mmap'ed, the text is
// parsed in place afterwards), so the decoder can be a tight loop: a 64-bit bit buffer refilled with one unaligned load,
// one table look-up per literal / length / distance (11-bit and 8-bit first-level tables with sub-tables for longer codes),
// several literals per refill, and 8-byte-at-a-time match copies.  Anything unusual (a code set that is not complete, a
// distance beyond the start of the output, a truncated stream) makes the function return false and the caller falls back to
// zlib, which then reports the error in its own terms.
//
// Checked against zlib by tests/emu/emu_inflate.cpp: random and FASTA-like inputs at every compression level and strategy
// (stored, fixed and dynamic blocks, long matches, distance-1 runs), truncations, multi-member files, the gzip fixtures of
// tests/golden/.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <stdlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <new>
#ifdef SK_INFLATE_TRACE
#include <chrono>
#include <cstdio>
#endif
#include <string>
#include <vector>

namespace sk_inflate {

// Growable byte buffer that does NOT initialise new space (std::string::resize zero-fills, which for a multi-GB text is a
// single-threaded pass over memory the decoder is about to overwrite anyway): malloc / realloc, size <= capacity.
template <class T>
class RawBuf {
 public:
  RawBuf() {}
  RawBuf(const RawBuf&) = delete;
  RawBuf& operator=(const RawBuf&) = delete;
  RawBuf(RawBuf&& o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
  RawBuf& operator=(RawBuf&& o) noexcept { if (this != &o) { free(p_); p_ = o.p_; n_ = o.n_; cap_ = o.cap_; o.p_ = nullptr; o.n_ = o.cap_ = 0; } return *this; }
  ~RawBuf() { free(p_); }
  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  T* data() { return p_; }
  const T* data() const { return p_; }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }
  void clear() { n_ = 0; }
  void release() { free(p_); p_ = nullptr; n_ = cap_ = 0; }
  void reserve(size_t c) {
    if (c <= cap_) return;
    T* q = (T*)realloc(p_, c * sizeof(T));
    if (!q) throw std::bad_alloc();
    p_ = q; cap_ = c;
  }
  void resize(size_t n) { if (n > cap_) reserve(std::max(n, cap_ + cap_ / 2)); n_ = n; }     // new elements are uninitialised
  void push_back(T v) { if (n_ == cap_) reserve(std::max<size_t>(cap_ * 2, 4096)); p_[n_++] = v; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};
typedef RawBuf<char> TextBuf;
typedef RawBuf<uint16_t> SymBuf;

constexpr int LT_BITS = 11;     // first-level bits of the literal/length table
constexpr int DT_BITS = 8;      // ... of the distance table
// table entry: [31:16] payload (literal byte, length / distance base, or sub-table start), [15:12] kind, [11:8] extra bits (or
// sub-table bits), [7:0] bits to consume = code length + extra bits (for sub-table pointers: LT_BITS / DT_BITS; inside a
// sub-table: the bits beyond the first level + extra bits)
enum : uint32_t { K_LIT = 1u << 12, K_BASE = 2u << 12, K_EOB = 4u << 12, K_SUB = 8u << 12 };

struct Tables {
  uint32_t lt[(1 << LT_BITS) + 1024];    // 288 symbols, codes <= 15 bits: sub-tables need < 1024 extra entries
  uint32_t dt[(1 << DT_BITS) + 512];
};

static inline uint32_t rev_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1); code >>= 1; }
  return r;
}

// canonical Huffman code lengths -> look-up table; returns false unless the code is complete (Kraft sum exactly 1)
static inline bool build_table(const uint8_t* lens, int n_sym, int table_bits, uint32_t* table, size_t table_cap,
                               const uint32_t* sym_entry /* entry without the length field, per symbol */) {
  int count[16] = {0};
  for (int s = 0; s < n_sym; s++) count[lens[s]]++;
  count[0] = 0;
  uint32_t kraft = 0;
  for (int l = 1; l <= 15; l++) kraft += (uint32_t)count[l] << (15 - l);
  if (kraft != (1u << 15)) return false;
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
  // longest code behind every first-level prefix (codes longer than table_bits share their low table_bits bits)
  const uint32_t prim = 1u << table_bits;
  std::vector<uint8_t> sub_len(prim, 0);
  std::vector<uint32_t> codes(n_sym, 0);
  {
    uint32_t nc[16];
    memcpy(nc, next_code, sizeof(nc));
    for (int s = 0; s < n_sym; s++) {
      const int l = lens[s];
      if (!l) continue;
      const uint32_t r = rev_bits(nc[l]++, l);
      codes[s] = r;
      if (l > table_bits) { uint8_t& m = sub_len[r & (prim - 1)]; if (l > m) m = (uint8_t)l; }
    }
  }
  size_t next_free = prim;
  std::vector<uint32_t> sub_start(prim, 0);
  for (uint32_t p = 0; p < prim; p++)
    if (sub_len[p]) {
      const int sb = sub_len[p] - table_bits;
      if (next_free + ((size_t)1 << sb) > table_cap) return false;
      sub_start[p] = (uint32_t)next_free;
      table[p] = ((uint32_t)next_free << 16) | K_SUB | ((uint32_t)sb << 8) | (uint32_t)table_bits;
      next_free += (size_t)1 << sb;
    }
  for (int s = 0; s < n_sym; s++) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t xb = (sym_entry[s] & K_BASE) ? ((sym_entry[s] >> 8) & 15) : 0;     // extra bits are consumed together with the code
    const uint32_t e = sym_entry[s] | ((uint32_t)l + xb), r = codes[s];
    if (l <= table_bits) {
      for (uint32_t i = r; i < prim; i += 1u << l) table[i] = e;
    } else {
      const uint32_t p = r & (prim - 1);
      const int sb = sub_len[p] - table_bits;
      const uint32_t hi = r >> table_bits, step = 1u << (l - table_bits);
      // inside the sub-table the entry's length field counts only the bits beyond the first level
      const uint32_t es = sym_entry[s] | ((uint32_t)(l - table_bits) + xb);
      for (uint32_t i = hi; i < (1u << sb); i += step) table[sub_start[p] + i] = es;
    }
  }
  return true;
}

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline bool build_litlen(const uint8_t* lens, int n, Tables& t) {
  uint32_t ent[288];
  for (int s = 0; s < 288; s++) {
    if (s < 256) ent[s] = ((uint32_t)s << 16) | K_LIT;
    else if (s == 256) ent[s] = K_EOB;
    else if (s < 286) ent[s] = ((uint32_t)LEN_BASE[s - 257] << 16) | K_BASE | ((uint32_t)LEN_EXTRA[s - 257] << 8);
    else ent[s] = 0;          // 286, 287: never valid in a stream (kind 0 = error)
  }
  uint8_t l[288] = {0};
  memcpy(l, lens, (size_t)n);
  if (!build_table(l, 288, LT_BITS, t.lt, sizeof(t.lt) / 4, ent)) return false;
  // Two literals per look-up where both codes fit the first-level index (nucleotide text has 2-3 bit codes for A/C/G/T, and
  // the chain "look up -> shift -> look up" is what bounds a Huffman decoder): bit 8 of a literal entry says that the payload
  // holds two bytes and the length field covers both codes.
  uint32_t first[1 << LT_BITS];
  memcpy(first, t.lt, sizeof(first));
  for (uint32_t i = 0; i < (1u << LT_BITS); i++) {
    const uint32_t e = first[i];
    if (!(e & K_LIT)) continue;
    const uint32_t l1 = e & 0xFF;
    if (l1 >= (uint32_t)LT_BITS) continue;
    const uint32_t e2 = first[i >> l1];
    if (!(e2 & K_LIT) || (e2 & 0xFF) > (uint32_t)LT_BITS - l1) continue;
    t.lt[i] = ((((e >> 16) & 0xFF) | (((e2 >> 16) & 0xFF) << 8)) << 16) | K_LIT | (1u << 8) | (l1 + (e2 & 0xFF));
  }
  return true;
}
static inline bool build_dist(const uint8_t* lens, int n, Tables& t) {
  uint32_t ent[32];
  for (int s = 0; s < 32; s++) ent[s] = s < 30 ? (((uint32_t)DIST_BASE[s] << 16) | K_BASE | ((uint32_t)DIST_EXTRA[s] << 8)) : 0;
  uint8_t l[32] = {0};
  memcpy(l, lens, (size_t)n);
  return build_table(l, 32, DT_BITS, t.dt, sizeof(t.dt) / 4, ent);
}

static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts only (x86-64)

// Decodes ONE raw deflate stream starting at in[0].  Output is appended to `out` (which also serves as the window: matches may
// reach back into bytes that were there before, never before out[base]).  On success *in_used = compressed bytes consumed.
inline bool inflate_raw(const uint8_t* in, size_t in_len, TextBuf& out, size_t base, size_t* in_used) {
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint64_t bb = 0;       // bit buffer, LSB first
  int bc = 0;            // valid bits in bb
  size_t pos = out.size();
  Tables* T = new Tables;
  struct Del { Tables* t; ~Del() { delete t; } } del{T};
  auto ensure = [&](size_t extra) {   // room for `extra` more bytes (+ slack for the word-wise match copy)
    if (pos + extra + 16 > out.size()) out.resize(std::max(out.size() * 2, pos + extra + 65536));
  };
  // byte-wise refill (headers, stored blocks, the last bytes of the stream)
  auto need = [&](int n) -> bool {
    while (bc < n) {
      if (ip >= in_end) return false;
      bb |= (uint64_t)*ip++ << bc;
      bc += 8;
    }
    return true;
  };
  for (;;) {
    if (!need(3)) return false;
    const uint32_t bfinal = (uint32_t)bb & 1, btype = ((uint32_t)bb >> 1) & 3;
    bb >>= 3; bc -= 3;
    if (btype == 0) {                 // stored: to the byte boundary, LEN / NLEN, raw bytes
      const int drop = bc & 7;
      bb >>= drop; bc -= drop;
      if (!need(32)) return false;
      const uint32_t len = (uint32_t)bb & 0xFFFF, nlen = ((uint32_t)(bb >> 16)) & 0xFFFF;
      bb >>= 32; bc -= 32;
      if ((len ^ nlen) != 0xFFFF) return false;
      ensure(len);
      uint32_t left = len;
      while (left && bc >= 8) { out[pos++] = (char)(bb & 0xFF); bb >>= 8; bc -= 8; left--; }
      if (left) {                       // bc == 0 here; bits above bc may mirror bytes at ip that are skipped now: forget them
        bb = 0;
        if ((size_t)(in_end - ip) < left) return false;
        memcpy(&out[pos], ip, left);
        pos += left; ip += left;
      }
    } else if (btype == 1 || btype == 2) {
      uint8_t ll[288 + 32];
      int hlit = 288, hdist = 32;
      if (btype == 1) {
        for (int i = 0; i < 144; i++) ll[i] = 8;
        for (int i = 144; i < 256; i++) ll[i] = 9;
        for (int i = 256; i < 280; i++) ll[i] = 7;
        for (int i = 280; i < 288; i++) ll[i] = 8;
        for (int i = 0; i < 32; i++) ll[288 + i] = 5;
        // the fixed distance code has 32 codes of 5 bits (30, 31 never occur): complete as it stands
        uint32_t ent[32];
        for (int s = 0; s < 32; s++) ent[s] = s < 30 ? (((uint32_t)DIST_BASE[s] << 16) | K_BASE | ((uint32_t)DIST_EXTRA[s] << 8)) : 0;
        if (!build_litlen(ll, 288, *T) || !build_table(ll + 288, 32, DT_BITS, T->dt, sizeof(T->dt) / 4, ent)) return false;
      } else {
        if (!need(14)) return false;
        hlit = 257 + ((int)bb & 31); hdist = 1 + ((int)(bb >> 5) & 31);
        const int hclen = 4 + ((int)(bb >> 10) & 15);
        bb >>= 14; bc -= 14;
        if (hlit > 286 || hdist > 30) return false;
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        for (int i = 0; i < hclen; i++) {
          if (!need(3)) return false;
          cl[order[i]] = (uint8_t)(bb & 7);
          bb >>= 3; bc -= 3;
        }
        // code-length code: 7-bit table built directly
        uint16_t clt[128];
        {
          int count[8] = {0};
          for (int s = 0; s < 19; s++) count[cl[s]]++;
          count[0] = 0;
          uint32_t kraft = 0;
          for (int l = 1; l <= 7; l++) kraft += (uint32_t)count[l] << (7 - l);
          if (kraft != 128) return false;
          uint32_t nc[8], code = 0;
          for (int l = 1; l <= 7; l++) { code = (code + (uint32_t)count[l - 1]) << 1; nc[l] = code; }
          for (int s = 0; s < 19; s++) {
            const int l = cl[s];
            if (!l) continue;
            const uint32_t r = rev_bits(nc[l]++, l);
            for (uint32_t i = r; i < 128; i += 1u << l) clt[i] = (uint16_t)((s << 4) | l);
          }
        }
        int n = 0;
        const int total = hlit + hdist;
        while (n < total) {
          while (bc < 56 && ip < in_end) { bb |= (uint64_t)*ip++ << bc; bc += 8; }   // near the end fewer bits may be left: checked per symbol
          const uint16_t e = clt[bb & 127];
          const int l = e & 15, sym = e >> 4;
          if (l > bc) return false;
          bb >>= l; bc -= l;
          if (sym < 16) { ll[n++] = (uint8_t)sym; continue; }
          int rep, val = 0, eb;
          if (sym == 16) { if (n == 0) return false; val = ll[n - 1]; eb = 2; rep = 3; }
          else if (sym == 17) { eb = 3; rep = 3; }
          else { eb = 7; rep = 11; }
          if (!need(eb)) return false;
          rep += (int)(bb & ((1u << eb) - 1));
          bb >>= eb; bc -= eb;
          if (n + rep > total) return false;
          while (rep--) ll[n++] = (uint8_t)val;
        }
        if (ll[256] == 0) return false;                          // no end-of-block code
        if (!build_litlen(ll, hlit, *T)) return false;
        if (!build_dist(ll + hlit, hdist, *T)) return false;      // incomplete distance codes (a block without matches from some encoders): zlib decides
      }
      // ---- symbols
      const uint32_t* const lt = T->lt;
      const uint32_t* const dt = T->dt;
      for (;;) {
        ensure(4096);
        const size_t out_lim = out.size() - 16 - 258 - 9;         // a whole iteration (up to 8 literals + scratch byte, or one match + copy slack) fits below this
        char* const o = &out[0];
        bool eob = false;
        // Fast loop.  Invariants at the top: >= 16 input bytes ahead of ip, the bit buffer was just refilled (bc >= 56) and e is
        // the entry for its low bits.  The look-up for the NEXT symbol is issued before a match is copied, so its latency
        // overlaps the copy (the serial chain look-up -> shift -> look-up is what bounds a Huffman decoder).
#define SK_REFILL() do { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
#define SK_PUT_LIT(e)                                                                          \
  do {                                                                                         \
    const uint16_t two__ = (uint16_t)((e) >> 16);                                              \
    memcpy(o + pos, &two__, 2);          /* one or two literals: the second byte is scratch if there is one */ \
    pos += 1 + (((e) >> 8) & 1);                                                               \
    bb >>= ((e) & 0xFF); bc -= (int)((e) & 0xFF);                                              \
  } while (0)
        if (pos < out_lim && (size_t)(in_end - ip) >= 16) {
          SK_REFILL();
          uint32_t e = lt[bb & ((1u << LT_BITS) - 1)];
          for (;;) {
            if (e & K_LIT) {                     // up to 3 look-ups (<= 11 bits each) after one refill, then refill and go on
              SK_PUT_LIT(e);
              e = lt[bb & ((1u << LT_BITS) - 1)];
              if (e & K_LIT) {
                SK_PUT_LIT(e);
                e = lt[bb & ((1u << LT_BITS) - 1)];
                if (e & K_LIT) {
                  SK_PUT_LIT(e);
                  e = lt[bb & ((1u << LT_BITS) - 1)];
                }
              }
              if (!(pos < out_lim && (size_t)(in_end - ip) >= 16)) break;     // e is dropped: the careful loop looks it up again
              SK_REFILL();                                                     // adds high bits only: e stays the entry of the low bits
              continue;
            }
            if (e & K_SUB) {
              bb >>= LT_BITS; bc -= LT_BITS;
              e = lt[(e >> 16) + (bb & ((1u << ((e >> 8) & 15)) - 1))];
              if (e & K_LIT) {
                SK_PUT_LIT(e);
                if (!(pos < out_lim && (size_t)(in_end - ip) >= 16)) break;
                SK_REFILL();
                e = lt[bb & ((1u << LT_BITS) - 1)];
                continue;
              }
            }
            if (e & K_EOB) { bb >>= (e & 0xFF); bc -= (int)(e & 0xFF); eob = true; break; }
            if (!(e & K_BASE)) return false;
            // code + extra bits leave the buffer with ONE shift; the extra bits are read from the saved copy off the critical chain
            const uint64_t sv = bb;
            const uint32_t tb = e & 0xFF, xb = (e >> 8) & 15;
            bb >>= tb; bc -= (int)tb;
            const uint32_t len = (e >> 16) + ((uint32_t)(sv >> (tb - xb)) & ((1u << xb) - 1));
            uint32_t d = dt[bb & ((1u << DT_BITS) - 1)];
            if (d & K_SUB) {
              bb >>= DT_BITS; bc -= DT_BITS;
              d = dt[(d >> 16) + (bb & ((1u << ((d >> 8) & 15)) - 1))];
            }
            if (!(d & K_BASE)) return false;
            const uint64_t sd = bb;
            const uint32_t td = d & 0xFF, db = (d >> 8) & 15;
            bb >>= td; bc -= (int)td;
            const size_t dist = (d >> 16) + ((uint32_t)(sd >> (td - db)) & ((1u << db) - 1));
            if (dist > pos - base) return false;
            char* dst = o + pos;
            const char* src = dst - dist;
            pos += len;
            const bool more = pos < out_lim && (size_t)(in_end - ip) >= 16;
            if (more) { SK_REFILL(); e = lt[bb & ((1u << LT_BITS) - 1)]; }       // next symbol's look-up in flight during the copy
            if (dist >= 8) {
              memcpy(dst, src, 8); memcpy(dst + 8, src + 8, 8);                    // in order: valid for any distance >= 8
              if (len > 16) {
                char* const end = dst + len;
                dst += 16; src += 16;
                do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < end);
              }
            } else if (dist == 1) {
              memset(dst, (unsigned char)*src, len);
            } else {
              for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
            }
            if (!more) break;
          }
        }
#undef SK_PUT_LIT
#undef SK_REFILL
        if (eob) break;
        // careful loop: one symbol at a time with byte-wise refill (end of input, or the output buffer has to grow)
        {
          // (bits above bc in bb always mirror the bytes at ip, so byte-wise refills continue seamlessly after the fast loop)
          while (bc < 56 && ip < in_end) { bb |= (uint64_t)*ip++ << bc; bc += 8; }
          if (bc == 0) return false;
          uint32_t e = lt[bb & ((1u << LT_BITS) - 1)];
          int used = 0;
          if (e & K_SUB) { const uint32_t sb = (e >> 8) & 15; e = lt[(e >> 16) + ((bb >> LT_BITS) & ((1u << sb) - 1))]; used = LT_BITS; }
          used += (int)(e & 0xFF);                               // code (+ extra bits of a length symbol)
          if (used > bc) return false;
          const uint64_t sv = bb;
          bb >>= used; bc -= used;
          if (e & K_LIT) { ensure(2); out[pos++] = (char)(e >> 16); if (e & (1u << 8)) out[pos++] = (char)(e >> 24); continue; }
          if (e & K_EOB) break;
          if (!(e & K_BASE)) return false;
          const int xb = (int)((e >> 8) & 15);
          const uint32_t len = (e >> 16) + ((uint32_t)(sv >> (used - xb)) & ((1u << xb) - 1));
          while (bc < 56 && ip < in_end) { bb |= (uint64_t)*ip++ << bc; bc += 8; }
          uint32_t d = dt[bb & ((1u << DT_BITS) - 1)];
          used = 0;
          if (d & K_SUB) { const uint32_t sb = (d >> 8) & 15; d = dt[(d >> 16) + ((bb >> DT_BITS) & ((1u << sb) - 1))]; used = DT_BITS; }
          if (!(d & K_BASE)) return false;
          used += (int)(d & 0xFF);
          if (used > bc) return false;
          const int db = (int)((d >> 8) & 15);
          const size_t dist = (d >> 16) + ((uint32_t)(bb >> (used - db)) & ((1u << db) - 1));
          bb >>= used; bc -= used;
          if (dist > pos - base) return false;
          ensure(len);
          for (uint32_t i = 0; i < len; i++) { out[pos] = out[pos - dist]; pos++; }
        }
      }
    } else {
      return false;
    }
    if (bfinal) break;
  }
  // whole bytes still sitting in the bit buffer were not part of the stream
  const size_t unused = (size_t)(bc >> 3);
  *in_used = (size_t)(ip - in) - unused;
  out.resize(pos);
  return true;
}

// crc: function computing the CRC-32 of a buffer continuing from a previous value (zlib's crc32 signature), or nullptr to skip
typedef unsigned long (*crc_fn)(unsigned long, const unsigned char*, unsigned int);

// gzip file (one or several members back to back, flate2 MultiGzDecoder semantics; bytes after the last member that do not
// start a new one end the stream, as zlib's gzread does) -> text.  Returns false on anything unexpected; `out` is then
// unspecified and the caller decodes the file with zlib instead.
inline bool gunzip(const uint8_t* p, size_t n, TextBuf& out, crc_fn crc) {
  out.clear();
  if (n >= 18) {   // size hint: ISIZE of the last member (exact for single-member files below 4 GB)
    const uint8_t* t = p + n - 4;
    const size_t isz = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
    if (isz <= n * 1100 + 65536) out.reserve(isz + 65536 + 512);       // deflate cannot expand more than ~1032x
  }
  size_t o = 0;
  size_t members = 0;
  while (o < n) {
    if (n - o < 18 || p[o] != 0x1f || p[o + 1] != 0x8b) break;
    if (p[o + 2] != 8) return false;
    const uint8_t flg = p[o + 3];
    if (flg & 0xE0) return false;
    size_t h = o + 10;
    if (flg & 4) { if (h + 2 > n) return false; h += 2 + (size_t)(p[h] | (p[h + 1] << 8)); }
    if (flg & 8) { while (h < n && p[h]) h++; h++; }
    if (flg & 16) { while (h < n && p[h]) h++; h++; }
    if (flg & 2) h += 2;
    if (h + 8 > n) return false;
    const size_t base = out.size();
    size_t used = 0;
    if (!inflate_raw(p + h, n - h, out, base, &used)) return false;
    const size_t t = h + used;
    if (t + 8 > n) return false;
    const uint32_t want_crc = (uint32_t)p[t] | ((uint32_t)p[t + 1] << 8) | ((uint32_t)p[t + 2] << 16) | ((uint32_t)p[t + 3] << 24);
    const uint32_t want_len = (uint32_t)p[t + 4] | ((uint32_t)p[t + 5] << 8) | ((uint32_t)p[t + 6] << 16) | ((uint32_t)p[t + 7] << 24);
    const size_t got = out.size() - base;
    if ((uint32_t)got != want_len) return false;
    if (crc) {
      unsigned long c = crc(0, nullptr, 0);
      for (size_t q = 0; q < got;) { const size_t m = std::min<size_t>(got - q, 1u << 30); c = crc(c, (const unsigned char*)out.data() + base + q, (unsigned int)m); q += m; }
      if ((uint32_t)c != want_crc) return false;
    }
    o = t + 8;
    members++;
  }
  return members > 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// Block-parallel decoding of ONE gzip member (a single large `.fa.gz`, e.g. a multi-FASTA of contigs read with -i).
//
// A deflate stream has no index, but its blocks can be found and decoded out of order (the two-pass scheme of pugz /
// rapidgzip, restated here for nucleotide TEXT):
//   1. the compressed payload is cut into chunks; every chunk but the first looks for the first bit position at which a
//      non-final dynamic-Huffman block header parses (complete code sets, an end-of-block code), the whole block decodes,
//      every literal is a text byte and the next block header is plausible too;
//   2. every chunk is decoded from its block start to the next chunk's start into 16-bit SYMBOLS: a byte, or 256 + w for a
//      byte copied from position w of the (still unknown) 32 KB window before the chunk;
//   3. the windows are resolved front to back (only the last 32 KB of a chunk matter for the next one), then all chunks are
//      resolved into the output in parallel.
// A wrong guess in step 1 cannot survive: the chunk before it must arrive at exactly that bit position, and the member's
// ISIZE and CRC-32 are checked at the end.  On any failure the function returns false and the caller decodes serially.
// ---------------------------------------------------------------------------------------------------------------------
struct BitPos {
  const uint8_t* in; size_t nbytes; size_t pos;       // pos = absolute bit position
  inline uint64_t peek() const {                      // >= 57 valid bits while >= 8 bytes are left; zeros past the end
    const size_t by = pos >> 3;
    uint64_t v = 0;
    if (by + 8 <= nbytes) memcpy(&v, in + by, 8);
    else for (size_t i = 0; by + i < nbytes; i++) v |= (uint64_t)in[by + i] << (8 * i);
    return v >> (pos & 7);
  }
  inline bool has(size_t nbits) const { return pos + nbits <= nbytes * 8; }
};

static inline bool text_byte(uint32_t c) { return c == 9 || c == 10 || c == 13 || (c >= 32 && c < 127); }

// header of a dynamic block at br.pos (after the 3 header bits): code sets -> tables
static inline bool parse_dynamic(BitPos& br, Tables& T) {
  if (!br.has(14)) return false;
  uint64_t v = br.peek();
  const int hlit = 257 + (int)(v & 31), hdist = 1 + (int)((v >> 5) & 31), hclen = 4 + (int)((v >> 10) & 15);
  br.pos += 14;
  if (hlit > 286 || hdist > 30) return false;
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  uint8_t cl[19] = {0};
  if (!br.has((size_t)hclen * 3)) return false;
  v = br.peek();                                      // 19 * 3 = 57 bits at most
  for (int i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)((v >> (3 * i)) & 7);
  br.pos += (size_t)hclen * 3;
  uint16_t clt[128];
  {
    int count[8] = {0};
    for (int s2 = 0; s2 < 19; s2++) count[cl[s2]]++;
    count[0] = 0;
    uint32_t kraft = 0;
    for (int l = 1; l <= 7; l++) kraft += (uint32_t)count[l] << (7 - l);
    if (kraft != 128) return false;
    uint32_t nc[8], code = 0;
    for (int l = 1; l <= 7; l++) { code = (code + (uint32_t)count[l - 1]) << 1; nc[l] = code; }
    for (int s2 = 0; s2 < 19; s2++) {
      const int l = cl[s2];
      if (!l) continue;
      const uint32_t r = rev_bits(nc[l]++, l);
      for (uint32_t i = r; i < 128; i += 1u << l) clt[i] = (uint16_t)((s2 << 4) | l);
    }
  }
  uint8_t ll[288 + 32];
  int n = 0;
  const int total = hlit + hdist;
  while (n < total) {
    if (!br.has(1)) return false;
    v = br.peek();
    const uint16_t e = clt[v & 127];
    const int l = e & 15, sym = e >> 4;
    if (!br.has((size_t)l)) return false;
    br.pos += (size_t)l; v >>= l;
    if (sym < 16) { ll[n++] = (uint8_t)sym; continue; }
    int rep, val = 0, eb;
    if (sym == 16) { if (n == 0) return false; val = ll[n - 1]; eb = 2; rep = 3; }
    else if (sym == 17) { eb = 3; rep = 3; }
    else { eb = 7; rep = 11; }
    if (!br.has((size_t)eb)) return false;
    rep += (int)(v & ((1u << eb) - 1));
    br.pos += (size_t)eb;
    if (n + rep > total) return false;
    while (rep--) ll[n++] = (uint8_t)val;
  }
  if (ll[256] == 0) return false;
  return build_litlen(ll, hlit, T) && build_dist(ll + hlit, hdist, T);
}
static inline bool build_fixed(Tables& T) {
  uint8_t ll[288 + 32];
  for (int i = 0; i < 144; i++) ll[i] = 8;
  for (int i = 144; i < 256; i++) ll[i] = 9;
  for (int i = 256; i < 280; i++) ll[i] = 7;
  for (int i = 280; i < 288; i++) ll[i] = 8;
  for (int i = 0; i < 32; i++) ll[288 + i] = 5;
  uint32_t ent[32];
  for (int s2 = 0; s2 < 32; s2++) ent[s2] = s2 < 30 ? (((uint32_t)DIST_BASE[s2] << 16) | K_BASE | ((uint32_t)DIST_EXTRA[s2] << 8)) : 0;
  return build_litlen(ll, 288, T) && build_table(ll + 288, 32, DT_BITS, T.dt, sizeof(T.dt) / 4, ent);
}

// symbols of ONE Huffman-coded block (tables built) appended to out; `text_only` rejects literals that are not text;
// `limit` bounds the block's output (a trial decode must not run away)
static inline bool decode_block_symbols(BitPos& br, const Tables& T, SymBuf& out, bool text_only, size_t limit) {
  const size_t start = out.size();
  for (;;) {
    if (!br.has(1)) return false;
    uint64_t v = br.peek();
    uint32_t e = T.lt[v & ((1u << LT_BITS) - 1)];
    uint32_t used = 0;
    if (e & K_SUB) { const uint32_t sb = (e >> 8) & 15; e = T.lt[(e >> 16) + ((v >> LT_BITS) & ((1u << sb) - 1))]; used = LT_BITS; }
    used += e & 0xFF;
    if (!br.has(used)) return false;
    if (e & K_LIT) {
      const uint32_t c0 = (e >> 16) & 0xFF;
      if (text_only && !text_byte(c0)) return false;
      out.push_back((uint16_t)c0);
      if (e & (1u << 8)) { const uint32_t c1 = (e >> 24) & 0xFF; if (text_only && !text_byte(c1)) return false; out.push_back((uint16_t)c1); }
      br.pos += used;
      if (out.size() - start > limit) return false;
      continue;
    }
    if (e & K_EOB) { br.pos += used; return true; }
    if (!(e & K_BASE)) return false;
    const uint32_t xb = (e >> 8) & 15;
    const uint32_t len = (e >> 16) + ((uint32_t)(v >> (used - xb)) & ((1u << xb) - 1));
    br.pos += used;
    v = br.peek();
    uint32_t d = T.dt[v & ((1u << DT_BITS) - 1)];
    used = 0;
    if (d & K_SUB) { const uint32_t sb = (d >> 8) & 15; d = T.dt[(d >> 16) + ((v >> DT_BITS) & ((1u << sb) - 1))]; used = DT_BITS; }
    if (!(d & K_BASE)) return false;
    used += d & 0xFF;
    if (!br.has(used)) return false;
    const uint32_t db = (d >> 8) & 15;
    const size_t dist = (d >> 16) + ((uint32_t)(v >> (used - db)) & ((1u << db) - 1));
    br.pos += used;
    if (dist > 32768) return false;
    const size_t q = out.size();
    if (q - start + len > limit) return false;
    out.resize(q + len);
    uint16_t* o = out.data();
    for (uint32_t i = 0; i < len; i++) {
      const size_t at = q + i;
      if (at >= dist) o[at] = o[at - dist];
      else o[at] = (uint16_t)(256 + (32768 - (dist - at)));        // position in the unknown window before the chunk
    }
  }
}

// The same without the per-literal checks, for the real pass: symbols written through a raw pointer with one capacity check
// per symbol group, up to three look-ups per 64-bit load, word copies for matches inside the chunk.  The last bytes of the
// input go through the careful routine above.
static inline bool decode_block_symbols_fast(BitPos& br, const Tables& T, SymBuf& out) {
  const uint8_t* const in = br.in;
  const size_t nbytes = br.nbytes;
  size_t pos = br.pos;
  for (;;) {
    if (out.capacity() < out.size() + 4096) out.reserve(std::max(out.capacity() * 2, out.size() + (1u << 20)));
    uint16_t* const o = out.data();
    size_t q = out.size();
    const size_t q_lim = out.capacity() - 600;
    bool eob = false;
    while (q < q_lim && (pos >> 3) + 16 <= nbytes) {
      uint64_t v = load64(in + (pos >> 3)) >> (pos & 7);             // >= 57 valid bits
      uint32_t e = T.lt[v & ((1u << LT_BITS) - 1)];
      if (e & K_LIT) {
        uint32_t used = 0;
        for (int k = 0; k < 3 && (e & K_LIT); k++) {                 // 3 x 11 bits of the 57
          o[q] = (uint16_t)((e >> 16) & 0xFF); o[q + 1] = (uint16_t)(e >> 24);
          q += 1 + ((e >> 8) & 1);
          const uint32_t l = e & 0xFF;
          used += l; v >>= l;
          e = T.lt[v & ((1u << LT_BITS) - 1)];
        }
        pos += used;
        continue;
      }
      uint32_t used = 0;
      if (e & K_SUB) { const uint32_t sb = (e >> 8) & 15; e = T.lt[(e >> 16) + ((v >> LT_BITS) & ((1u << sb) - 1))]; used = LT_BITS; }
      used += e & 0xFF;
      if (e & K_LIT) { o[q++] = (uint16_t)((e >> 16) & 0xFF); pos += used; continue; }       // sub-table literals are single
      if (e & K_EOB) { pos += used; eob = true; break; }
      if (!(e & K_BASE)) return false;
      const uint32_t xb = (e >> 8) & 15;
      const uint32_t len = (e >> 16) + ((uint32_t)(v >> (used - xb)) & ((1u << xb) - 1));
      pos += used; v >>= used;                                        // <= 20 bits gone: >= 37 left, a distance needs <= 28
      uint32_t d = T.dt[v & ((1u << DT_BITS) - 1)];
      used = 0;
      if (d & K_SUB) { const uint32_t sb = (d >> 8) & 15; d = T.dt[(d >> 16) + ((v >> DT_BITS) & ((1u << sb) - 1))]; used = DT_BITS; }
      if (!(d & K_BASE)) return false;
      used += d & 0xFF;
      const uint32_t db = (d >> 8) & 15;
      const size_t dist = (d >> 16) + ((uint32_t)(v >> (used - db)) & ((1u << db) - 1));
      pos += used;
      if (dist > 32768) return false;
      if (q >= dist) {
        const uint16_t* src = o + q - dist;
        uint16_t* dst = o + q;
        if (dist >= 4) {                                              // 8-byte words of 4 symbols, in order
          uint16_t* const end = dst + len;
          do { memcpy(dst, src, 8); dst += 4; src += 4; } while (dst < end);
        } else for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
      } else {
        for (uint32_t i = 0; i < len; i++) {
          const size_t at = q + i;
          o[at] = at >= dist ? o[at - dist] : (uint16_t)(256 + (32768 - (dist - at)));
        }
      }
      q += len;
    }
    out.resize(q);
    br.pos = pos;
    if (eob) return true;
    // tail of the input, or the buffer has to grow: one careful step, then back to the fast loop
    if ((pos >> 3) + 16 > nbytes) return decode_block_symbols(br, T, out, false, (size_t)-1 >> 1);
  }
}

// one block (any type) at br.pos -> symbols; *final = BFINAL
static inline bool decode_any_block(BitPos& br, Tables& T, SymBuf& out, bool* final, bool text_only, size_t limit) {
  if (!br.has(3)) return false;
  const uint64_t v = br.peek();
  *final = (v & 1) != 0;
  const uint32_t btype = (uint32_t)(v >> 1) & 3;
  br.pos += 3;
  if (btype == 0) {
    br.pos = (br.pos + 7) & ~(size_t)7;
    if (!br.has(32)) return false;
    const uint64_t w = br.peek();
    const uint32_t len = (uint32_t)w & 0xFFFF, nlen = (uint32_t)(w >> 16) & 0xFFFF;
    if ((len ^ nlen) != 0xFFFF) return false;
    br.pos += 32;
    if (!br.has((size_t)len * 8) || len > limit) return false;
    const uint8_t* src = br.in + (br.pos >> 3);
    for (uint32_t i = 0; i < len; i++) { if (text_only && !text_byte(src[i])) return false; out.push_back(src[i]); }
    br.pos += (size_t)len * 8;
    return true;
  }
  if (btype == 1) { if (!build_fixed(T)) return false; }
  else if (btype == 2) { if (!parse_dynamic(br, T)) return false; }
  else return false;
  return text_only ? decode_block_symbols(br, T, out, true, limit) : decode_block_symbols_fast(br, T, out);
}

// first bit position >= from (and < to) at which a non-final dynamic block of text starts, or SIZE_MAX
static inline size_t find_block_start(const uint8_t* in, size_t nbytes, size_t from, size_t to, Tables& T, SymBuf& scratch) {
  for (size_t b = from; b < to; b++) {
    BitPos br{in, nbytes, b};
    if (!br.has(3 + 14)) return (size_t)-1;
    const uint64_t v = br.peek();
    if ((v & 7) != 4) continue;                                     // BFINAL = 0, BTYPE = 2 (bits: 0, then 01 LSB first = value 4)
    if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue;     // HLIT <= 286, HDIST <= 30
    scratch.clear();
    bool fin = false;
    if (!decode_any_block(br, T, scratch, &fin, true, 1u << 22)) continue;
    if (scratch.size() < 1024) continue;                            // real blocks of a large text file are tens of KB
    // the block that follows must look like a block too
    BitPos nx = br;
    if (!nx.has(3)) continue;
    const uint64_t w = nx.peek();
    const uint32_t bt = (uint32_t)(w >> 1) & 3;
    if (bt == 3) continue;
    if (bt == 2) { nx.pos += 3; Tables* T2 = new Tables; const bool okh = parse_dynamic(nx, *T2); delete T2; if (!okh) continue; }
    else if (bt == 0) {
      nx.pos += 3; nx.pos = (nx.pos + 7) & ~(size_t)7;
      if (!nx.has(32)) continue;
      const uint64_t x = nx.peek();
      if ((((uint32_t)x & 0xFFFF) ^ ((uint32_t)(x >> 16) & 0xFFFF)) != 0xFFFF) continue;
    }
    return b;
  }
  return (size_t)-1;
}

typedef unsigned long (*crc_combine_fn)(unsigned long, unsigned long, long);

// ONE gzip member decoded by `threads` threads (see above).  run(n, f) must execute f(0) .. f(n-1), possibly concurrently,
// and return when all are done.  Returns false if the file is not a single member, is too small to bother, is not text, or
// anything does not add up: the caller then uses the serial decoder.
template <class ParallelFor>
inline bool gunzip_parallel(const uint8_t* p, size_t n, TextBuf& out, int threads, crc_fn crc, crc_combine_fn crc_combine,
                            ParallelFor&& run, size_t min_chunk = 8u << 20) {
  if (threads < 2 || n < 18 + 3 * min_chunk || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return false;
  const uint8_t flg = p[3];
  if (flg & 0xE0) return false;
  size_t h = 10;
  if (flg & 4) { if (h + 2 > n) return false; h += 2 + (size_t)(p[h] | (p[h + 1] << 8)); }
  if (flg & 8) { while (h < n && p[h]) h++; h++; }
  if (flg & 16) { while (h < n && p[h]) h++; h++; }
  if (flg & 2) h += 2;
  if (h + 8 >= n) return false;
  const size_t payload_end = n - 8;                                  // if this is the only member, its trailer is the last 8 bytes
  if ((payload_end - h) / min_chunk < 3) return false;
  const uint32_t want_crc = (uint32_t)p[n - 8] | ((uint32_t)p[n - 7] << 8) | ((uint32_t)p[n - 6] << 16) | ((uint32_t)p[n - 5] << 24);
  const uint32_t want_len = (uint32_t)p[n - 4] | ((uint32_t)p[n - 3] << 8) | ((uint32_t)p[n - 2] << 16) | ((uint32_t)p[n - 1] << 24);
  struct Chunk { size_t start_bit = (size_t)-1, end_bit = 0; SymBuf sym; bool ok = false, final_seen = false; unsigned long crc = 0; size_t out_off = 0; };
  out.clear();
  { const size_t hint = (size_t)want_len; if (hint <= (payload_end - h) * 1100 + 65536) out.reserve(hint + 64); }
  // The member is processed in ROUNDS of up to 2 x threads chunks, so that the symbol buffers (2 bytes per output byte) stay
  // bounded for files of any size: a round starts at an exactly known bit (the end of the previous round) with a known window.
  size_t round_bit = h * 8;
  std::vector<uint8_t> round_win;                                    // the 32 KB of output before the round (empty for the first)
  unsigned long crc_all = crc ? crc(0, nullptr, 0) : 0;
  bool finished = false;
  std::vector<SymBuf> pool((size_t)threads * 2 + 1);             // symbol buffers live across rounds: their pages are touched once
#ifdef SK_INFLATE_TRACE
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
#endif
  while (!finished) {
#ifdef SK_INFLATE_TRACE
    const double tt0 = tnow();
#endif
    const size_t byte0 = round_bit >> 3;
    if (byte0 >= payload_end) return false;
    const size_t remaining = payload_end - byte0;
    const size_t K = std::max<size_t>(1, std::min<size_t>((size_t)threads * 2, (remaining + min_chunk - 1) / min_chunk));
    const size_t round_end = std::min(payload_end, byte0 + K * min_chunk);
    bool last_round = round_end == payload_end;
    std::vector<Chunk> ch(K + 1);                                    // ch[K] = the block start the round's last chunk stops at (next round's first)
    ch[0].start_bit = round_bit;
    for (size_t k = 0; k < K; k++) { ch[k].sym = std::move(pool[k]); ch[k].sym.clear(); }
    // 1. block starts (chunk k looks inside its own byte range; the sentinel looks right after the round)
    run(K + 1, [&](size_t k) {
      if (k == 0 || (k == K && last_round)) return;
      const size_t from = std::max(round_bit + 1, (byte0 + k * min_chunk) * 8);
      const size_t to = std::min(payload_end, byte0 + (k + (k == K ? 4 : 1)) * min_chunk) * 8;   // the sentinel may look further (a block can be longer than a small chunk)
      if (from >= to) return;
      Tables* T = new Tables;
      SymBuf scratch;
      scratch.reserve(1u << 20);
      ch[k].start_bit = find_block_start(p, payload_end, from, to, *T, scratch);
      delete T;
    });
    if (!last_round && ch[K].start_bit == (size_t)-1) {
      // no qualifying block start after the round: fine if the search reached the end of the member (only the final block, or a
      // few small ones, are left: the round's last chunk decodes them too), otherwise give up (serial decoder)
      if (byte0 + (K + 4) * min_chunk < payload_end) return false;
      last_round = true;
    }
    std::vector<size_t> use;                                          // chunks without a block start are merged into their predecessor
    for (size_t k = 0; k < K; k++) if (ch[k].start_bit != (size_t)-1) use.push_back(k);
#ifdef SK_INFLATE_TRACE
    const double tt1 = tnow();
#endif
    // 2. symbols
    run(use.size(), [&](size_t u) {
      Chunk& c = ch[use[u]];
      const size_t stop = u + 1 < use.size() ? ch[use[u + 1]].start_bit : (last_round ? (size_t)-1 : ch[K].start_bit);
      Tables* T = new Tables;
      BitPos br{p, payload_end, c.start_bit};
      c.sym.reserve((size_t)((stop == (size_t)-1 ? payload_end * 8 - c.start_bit : stop - c.start_bit) / 8 * 4) + 4096);
      bool good = true, fin = false;
      while (good) {
        if (br.pos == stop) break;
        if (br.pos > stop) { good = false; break; }                  // ran past the next start: that start was not a block boundary
        good = decode_any_block(br, *T, c.sym, &fin, false, (size_t)-1 >> 1);
        if (good && fin) { c.final_seen = true; break; }
      }
      if (good && stop != (size_t)-1 && (c.final_seen || br.pos != stop)) good = false;
      if (good && stop == (size_t)-1 && !c.final_seen) good = false;
      c.end_bit = br.pos;
      c.ok = good;
      delete T;
    });
#ifdef SK_INFLATE_TRACE
    const double tt2 = tnow();
#endif
    const size_t out0 = out.size();
    size_t total = 0;
    for (size_t u = 0; u < use.size(); u++) { Chunk& c = ch[use[u]]; if (!c.ok) return false; c.out_off = out0 + total; total += c.sym.size(); }
    if (last_round) {
      // the member must end right before its trailer (otherwise: more members, or garbage -- the serial path sorts that out)
      if (!ch[use.back()].final_seen || ((ch[use.back()].end_bit + 7) >> 3) != payload_end) return false;
      finished = true;
    }
    // 3a. windows, front to back: win[u] = the 32 KB of output before chunk u (only tails are resolved here)
    std::vector<std::vector<uint8_t>> win(use.size() + 1);
    win[0] = round_win;
    for (size_t u = 0; u < use.size(); u++) {
      const Chunk& c = ch[use[u]];
      const std::vector<uint8_t>& w = win[u];
      std::vector<uint8_t>& nw = win[u + 1];
      nw.assign(32768, 0);
      const size_t m = c.sym.size();
      const size_t take = std::min<size_t>(m, 32768);
      for (size_t i = 0; i < 32768 - take; i++) nw[i] = w.empty() ? 0 : w[i + take];
      for (size_t i = 0; i < take; i++) {
        const uint16_t sy = c.sym[m - take + i];
        if (sy < 256) nw[32768 - take + i] = (uint8_t)sy;
        else { if (w.empty()) return false; nw[32768 - take + i] = w[sy - 256]; }   // a reference before the start of the member
      }
    }
    // 3b. all chunks of the round into the output, in parallel; CRC per chunk
    out.resize(out0 + total);
    std::vector<char> bad(use.size(), 0);
    run(use.size(), [&](size_t u) {
      Chunk& c = ch[use[u]];
      const std::vector<uint8_t>& w = win[u];
      char* o = &out[c.out_off];
      const uint16_t* sy = c.sym.data();
      const size_t m = c.sym.size();
      size_t i = 0;
#if defined(__SSE2__)
      for (; i + 16 <= m; i += 16) {                                  // 16 symbols at a time; groups without window references are packed
        const __m128i a = _mm_loadu_si128((const __m128i*)(sy + i)), b = _mm_loadu_si128((const __m128i*)(sy + i + 8));
        if (_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_srli_epi16(_mm_or_si128(a, b), 8), _mm_setzero_si128())) == 0xFFFF) {
          _mm_storeu_si128((__m128i*)(o + i), _mm_packus_epi16(a, b));
        } else {
          for (size_t k = i; k < i + 16; k++) {
            if (sy[k] < 256) o[k] = (char)sy[k];
            else if (w.empty()) { bad[u] = 1; o[k] = 0; }
            else o[k] = (char)w[sy[k] - 256];
          }
        }
      }
#endif
      for (; i < m; i++) {
        if (sy[i] < 256) o[i] = (char)sy[i];
        else if (w.empty()) { bad[u] = 1; o[i] = 0; }
        else o[i] = (char)w[sy[i] - 256];
      }
      if (crc) {
        unsigned long cc = crc(0, nullptr, 0);
        for (size_t q = 0; q < m;) { const size_t mm = std::min<size_t>(m - q, 1u << 30); cc = crc(cc, (const unsigned char*)o + q, (unsigned int)mm); q += mm; }
        c.crc = cc;
      }
    });
    for (size_t k = 0; k < K; k++) pool[k] = std::move(ch[k].sym);
    for (char b : bad) if (b) return false;
    if (crc && crc_combine)
      for (size_t u = 0; u < use.size(); u++) {
        const Chunk& c = ch[use[u]];
        const size_t m = (u + 1 < use.size() ? ch[use[u + 1]].out_off : out0 + total) - c.out_off;
        crc_all = crc_combine(crc_all, c.crc, (long)m);
      }
#ifdef SK_INFLATE_TRACE
    fprintf(stderr, "[gunzip_parallel] round of %zu chunks: find %.1f ms, symbols %.1f ms, resolve+crc %.1f ms, %zu bytes\n", use.size(),
            (tt1 - tt0) * 1e3, (tt2 - tt1) * 1e3, (tnow() - tt2) * 1e3, total);
#endif
    if (!finished) { round_bit = ch[K].start_bit; round_win = win[use.size()]; }
  }
  if ((uint32_t)out.size() != want_len) return false;
  if (crc && crc_combine && (uint32_t)crc_all != want_crc) return false;
  return true;
}

}  // namespace sk_inflate
