// fastx.hpp -- FASTA / FASTQ(.gz) record reader of the host CLI (host only).
// Record rules follow needletail 0.5.1 as skani uses it (src/file_io.rs:158-176): format sniffed from the
// first byte ('>' FASTA, '@' FASTQ), transparent gzip (zlib; multi-member streams), FASTA sequences with line breaks
// ("\n" / "\r\n") removed, id = the whole header line without the leading symbol, 4-line FASTQ records; an empty or
// non-FASTX file is an error (the caller warns and skips it, src/file_io.rs:159-166).
//
// Ingestion speed: a file is read whole, inflated into one buffer and split into lines with
// memchr (no per-byte state machine).  Block-gzipped files (BGZF: bgzip / htslib, every member carries its compressed
// size in a 'BC' extra field) are inflated member-parallel with `inflate_threads` threads; ordinary single-member gzip has
// no block index: by the whole-buffer decoder of fast_inflate.hpp (1.7x zlib on nucleotide text), and when threads are to
// spare (few large files) block-parallel by its two-pass scheme (blocks located by trial decoding, decoded against an unknown
// window, resolved afterwards; ISIZE + CRC-32 verified); zlib as the fallback.  All of it checked against zlib in
// tests/emu/emu_inflate.cpp.
#pragma once
#include <zlib.h>

#include "fast_inflate.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace fastx {

struct Record { std::string id; std::string seq; };

// read-only view of a whole file (mmap; empty files map to a null view)
struct FileView {
  const unsigned char* p = nullptr;
  size_t n = 0;
  bool ok = false;
  explicit FileView(const std::string& path) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); return; }
    n = (size_t)st.st_size;
    ok = true;
    if (n) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { ok = false; n = 0; }
      else { p = (const unsigned char*)m; madvise(m, n, MADV_SEQUENTIAL); }
    }
    close(fd);
  }
  ~FileView() { if (p) munmap((void*)p, n); }
  FileView(const FileView&) = delete;
  FileView& operator=(const FileView&) = delete;
};

// BGZF member at `p` (n bytes available)?  returns the member's total compressed size (BSIZE + 1), or 0
inline size_t bgzf_member_size(const unsigned char* p, size_t n) {
  if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
  const size_t xlen = p[10] | (p[11] << 8);
  if (12 + xlen > n) return 0;
  for (size_t o = 12; o + 4 <= 12 + xlen;) {
    const size_t slen = p[o + 2] | (p[o + 3] << 8);
    if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) return (size_t)(p[o + 4] | (p[o + 5] << 8)) + 1;
    o += 4 + slen;
  }
  return 0;
}

// gzip bytes -> decompressed text: multi-member streams via zlib; BGZF member-parallel.
struct RawSpan { const unsigned char* p; size_t n; const unsigned char* data() const { return p; } size_t size() const { return n; } const unsigned char& operator[](size_t i) const { return p[i]; } };
using TextBuf = sk_inflate::TextBuf;
inline bool inflate_gzip(const RawSpan raw, TextBuf& out, int inflate_threads = 1) {
  out.clear();
  // ---- BGZF: every member announces its size, so the members can be located without inflating and inflated in parallel
  if (bgzf_member_size(raw.data(), raw.size())) {
    struct Mem { size_t off, csize, isize, out_off; };
    std::vector<Mem> mem;
    size_t o = 0, total = 0;
    bool all = true;
    while (o < raw.size()) {
      const size_t cs = bgzf_member_size(raw.data() + o, raw.size() - o);
      if (!cs || o + cs > raw.size() || cs < 26) { all = false; break; }
      const unsigned char* t = raw.data() + o + cs - 4;
      const size_t isz = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
      mem.push_back(Mem{o, cs, isz, total});
      total += isz;
      o += cs;
    }
    if (all && !mem.empty()) {
      out.resize(total);
      std::atomic<size_t> next{0};
      std::atomic<bool> ok{true};
      const bool use_fast = getenv("SK_ZLIB_INFLATE") == nullptr;
      auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < mem.size();) {
          const Mem& m = mem[i];
          if (m.isize == 0) continue;
          const unsigned char* p = raw.data() + m.off;
          const size_t xlen = p[10] | (p[11] << 8), hdr = 12 + xlen;
          if (hdr + 8 > m.csize) { ok = false; continue; }
          if (use_fast) {   // whole-buffer decoder into a per-thread scratch string (a member is <= 64 KB), zlib if it declines
            static thread_local TextBuf scratch;
            scratch.clear();
            size_t used = 0;
            if (sk_inflate::inflate_raw(p + hdr, raw.size() - (m.off + hdr), scratch, 0, &used) && scratch.size() == m.isize &&
                used + hdr + 8 <= m.csize) {
              memcpy(&out[m.out_off], scratch.data(), m.isize);
              continue;
            }
          }
          z_stream zs;
          memset(&zs, 0, sizeof(zs));
          if (inflateInit2(&zs, -15) != Z_OK) { ok = false; continue; }
          zs.next_in = (Bytef*)(p + hdr); zs.avail_in = (uInt)(m.csize - hdr - 8);
          zs.next_out = (Bytef*)&out[m.out_off]; zs.avail_out = (uInt)m.isize;
          const int rc = inflate(&zs, Z_FINISH);
          if (rc != Z_STREAM_END || zs.avail_out != 0) ok = false;
          inflateEnd(&zs);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < inflate_threads && (size_t)t < mem.size(); t++) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
      return ok.load();
    }
  }
  // ---- ordinary gzip (possibly several members back to back: flate2 MultiGzDecoder semantics): the whole-buffer decoder of
  //      fast_inflate.hpp first (1.6x zlib on nucleotide text); anything it does not like is decoded again by zlib below, which
  //      then accepts or rejects the file in its own terms.  SK_ZLIB_INFLATE=1 forces the zlib path (A/B, tests).
  if (getenv("SK_ZLIB_INFLATE") == nullptr) {
    // ONE large member and threads to spare (a multi-FASTA of contigs read with -i): block-parallel two-pass decoding
    // (fast_inflate.hpp gunzip_parallel); it declines anything that is not a single text member, or that does not verify
    // (worth its 2 bytes of symbol buffer per output byte from ~30 MB of compressed data and 4 threads on)
    if (inflate_threads >= (getenv("SK_INFLATE_MIN_CHUNK") ? 2 : 4) && getenv("SK_SERIAL_INFLATE") == nullptr) {
      auto run = [&](size_t n, const std::function<void(size_t)>& f) {
        std::atomic<size_t> next{0};
        auto w = [&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); };
        std::vector<std::thread> th;
        for (int t = 1; t < inflate_threads && (size_t)t < n; t++) th.emplace_back(w);
        w();
        for (auto& t : th) t.join();
      };
      const char* mc = getenv("SK_INFLATE_MIN_CHUNK");            // test hook: small chunks on small files
      if (sk_inflate::gunzip_parallel(raw.data(), raw.size(), out, inflate_threads, (sk_inflate::crc_fn)crc32,
                                      (sk_inflate::crc_combine_fn)crc32_combine, run, mc ? (size_t)atoll(mc) : (8u << 20))) {
        if (getenv("SK_TRACE")) fprintf(stderr, "[inflate] block-parallel: %zu -> %zu bytes with %d threads\n", raw.size(), out.size(), inflate_threads);
        return true;
      }
      if (getenv("SK_TRACE")) fprintf(stderr, "[inflate] block-parallel declined, serial\n");
      out.clear();
    }
    if (sk_inflate::gunzip(raw.data(), raw.size(), out, (sk_inflate::crc_fn)crc32)) return true;
    out.clear();
  }
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 15 + 16) != Z_OK) return false;
  zs.next_in = (Bytef*)raw.data(); zs.avail_in = (uInt)std::min<size_t>(raw.size(), 1u << 30);
  size_t consumed = 0;
  out.resize(std::max<size_t>(raw.size() * 4, 1 << 16));
  size_t produced = 0;
  bool ok = true;
  for (;;) {
    if (produced == out.size()) out.resize(out.size() * 2);
    zs.next_out = (Bytef*)&out[produced];
    zs.avail_out = (uInt)std::min<size_t>(out.size() - produced, 1u << 30);
    const uInt in0 = zs.avail_in, out0 = zs.avail_out;
    const int rc = inflate(&zs, Z_NO_FLUSH);
    consumed += in0 - zs.avail_in;
    produced += out0 - zs.avail_out;
    if (zs.avail_in == 0 && consumed < raw.size()) { zs.next_in = (Bytef*)raw.data() + consumed; zs.avail_in = (uInt)std::min<size_t>(raw.size() - consumed, 1u << 30); }
    if (rc == Z_STREAM_END) {
      if (consumed >= raw.size()) break;
      // another member?  (trailing garbage that is not a gzip header ends the stream, as zlib's gzread does)
      if (raw.size() - consumed < 2 || raw[consumed] != 0x1f || raw[consumed + 1] != 0x8b) break;
      if (inflateReset(&zs) != Z_OK) { ok = false; break; }
      continue;
    }
    if (rc != Z_OK && rc != Z_BUF_ERROR) { ok = false; break; }
    if (rc == Z_BUF_ERROR && zs.avail_in == 0 && consumed >= raw.size()) { ok = false; break; }   // truncated stream
  }
  inflateEnd(&zs);
  out.resize(produced);
  return ok;
}

// A record located inside a text buffer, not copied: its sequence is the lines of text[seq_b, seq_e) with the line breaks
// ("\n" / "\r\n") removed, n_bases bytes in all.  The CLI lays all records out first and then strips the lines straight into
// the one flat buffer the GPU library takes (one copy per base instead of text -> record string -> flat buffer).
struct RecordView { std::string id; size_t seq_b = 0, seq_e = 0, n_bases = 0; };

inline bool scan_fastx(const char* data, size_t n, std::vector<RecordView>& out) {
  enum { START, FA_SEQ, FQ_SEQ, FQ_PLUS, FQ_QUAL } st = START;
  bool ok = true, any = false, open = false;
  RecordView cur;
  size_t fq_len = 0;
  size_t b = 0;
  while (ok && b < n) {
    const char* nl = (const char*)memchr(data + b, '\n', n - b);
    const size_t e = nl ? (size_t)(nl - data) : n;
    const char* p = data + b;
    size_t len = e - b;
    if (len && p[len - 1] == '\r') len--;
    switch (st) {
      case START:
        if (len == 0) { if (!any) ok = false; break; }
        if (p[0] == '>' || p[0] == '@') {
          cur = RecordView(); cur.id.assign(p + 1, len - 1); cur.seq_b = cur.seq_e = std::min(e + 1, n);
          st = p[0] == '>' ? FA_SEQ : FQ_SEQ; any = true; open = true;
        } else ok = false;
        break;
      case FA_SEQ:
        if (len && p[0] == '>') {
          out.push_back(std::move(cur));
          cur = RecordView(); cur.id.assign(p + 1, len - 1); cur.seq_b = cur.seq_e = std::min(e + 1, n);
        } else { cur.n_bases += len; cur.seq_e = std::min(e + 1, n); }
        break;
      case FQ_SEQ: cur.seq_b = b; cur.seq_e = std::min(e + 1, n); cur.n_bases = len; fq_len = len; st = FQ_PLUS; break;
      case FQ_PLUS: if (len == 0 || p[0] != '+') ok = false; st = FQ_QUAL; break;
      case FQ_QUAL:
        if (len != fq_len) ok = false;
        out.push_back(std::move(cur)); cur = RecordView(); open = false; st = START;
        break;
    }
    b = e + 1;
  }
  if (ok && st == FA_SEQ && open) out.push_back(std::move(cur));
  if (ok && (st == FQ_SEQ || st == FQ_PLUS || st == FQ_QUAL)) ok = false;
  if (!any) ok = false;  // empty file (needletail: EmptyFile error)
  return ok;
}

// the sequence of a located record, line breaks removed; dst must hold r.n_bases bytes
inline void copy_sequence(const char* data, const RecordView& r, char* dst) {
  size_t b = r.seq_b;
  while (b < r.seq_e) {
    const char* nl = (const char*)memchr(data + b, '\n', r.seq_e - b);
    const size_t e = nl ? (size_t)(nl - data) : r.seq_e;
    size_t len = e - b;
    if (len && data[b + len - 1] == '\r') len--;
    memcpy(dst, data + b, len);
    dst += len;
    b = e + 1;
  }
}

// records of an in-memory FASTA / FASTQ text (copies)
inline bool parse_fastx(const char* data, size_t n, std::vector<Record>& out) {
  std::vector<RecordView> v;
  if (!scan_fastx(data, n, v)) return false;
  out.reserve(out.size() + v.size());
  for (auto& r : v) {
    Record rec;
    rec.id = std::move(r.id);
    rec.seq.resize(r.n_bases);
    if (r.n_bases) copy_sequence(data, r, &rec.seq[0]);
    out.push_back(std::move(rec));
  }
  return true;
}

// a file opened for the two-step read: the text stays alive (the mapping, or the inflated buffer) until the sequences are copied
struct LoadedFile {
  std::unique_ptr<FileView> fv;
  TextBuf text;
  const char* data = nullptr;
  size_t n = 0;
  std::vector<RecordView> recs;
};
inline bool open_fastx(const std::string& path, LoadedFile& f, int inflate_threads = 1) {
  f.fv.reset(new FileView(path));
  if (!f.fv->ok) return false;
  if (f.fv->n >= 2 && f.fv->p[0] == 0x1f && f.fv->p[1] == 0x8b) {
    if (!inflate_gzip(RawSpan{f.fv->p, f.fv->n}, f.text, inflate_threads)) return false;
    f.fv.reset();                                               // the compressed bytes are not needed any more
    f.data = f.text.data(); f.n = f.text.size();
  } else { f.data = (const char*)f.fv->p; f.n = f.fv->n; }     // plain text: scanned straight from the mapping
  return scan_fastx(f.data, f.n, f.recs);
}

inline bool read_fastx(const std::string& path, std::vector<Record>& out, int inflate_threads = 1) {
  LoadedFile f;
  if (!open_fastx(path, f, inflate_threads)) return false;
  out.reserve(out.size() + f.recs.size());
  for (auto& r : f.recs) {
    Record rec;
    rec.id = std::move(r.id);
    rec.seq.resize(r.n_bases);
    if (r.n_bases) copy_sequence(f.data, r, &rec.seq[0]);
    out.push_back(std::move(rec));
  }
  return true;
}

// allocator that leaves trivially constructible elements uninitialised: a std::vector<uint8_t> of tens of GB is not zero-filled
// by one thread before the parallel copies write (and first-touch) it
template <class T>
struct no_init_alloc : std::allocator<T> {
  template <class U> struct rebind { using other = no_init_alloc<U>; };
  template <class U> void construct(U* p) noexcept { ::new ((void*)p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};

}  // namespace fastx
