// fastx.hpp -- FASTA / FASTQ(.gz) record reader of the host CLI (host only).
// Record rules follow needletail 0.5.1 as skani uses it (src/file_io.rs:158-176): format sniffed from the
// first byte ('>' FASTA, '@' FASTQ), transparent gzip (zlib; multi-member streams), FASTA sequences with line breaks
// ("\n" / "\r\n") removed, id = the whole header line without the leading symbol, 4-line FASTQ records; an empty or
// non-FASTX file is an error (the caller warns and skips it, src/file_io.rs:159-166).
#pragma once
#include <zlib.h>

#include <string>
#include <vector>

namespace fastx {

struct Record { std::string id; std::string seq; };

// ---- FASTA / FASTQ reader (streaming over zlib; plain files pass through gzread unchanged) -----------------
inline bool read_fastx(const std::string& path, std::vector<Record>& out) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) return false;
  gzbuffer(f, 1 << 20);
  std::vector<char> buf(1 << 22);
  std::string line, pending;
  enum { START, FA_SEQ, FQ_SEQ, FQ_PLUS, FQ_QUAL } st = START;
  bool ok = true, any = false;
  Record cur;
  size_t fq_len = 0;
  auto flush_line = [&](std::string& ln) {
    if (!ln.empty() && ln.back() == '\r') ln.pop_back();
    switch (st) {
      case START:
        if (ln.empty()) { if (any) return; ok = false; return; }
        if (ln[0] == '>') { cur = Record(); cur.id = ln.substr(1); st = FA_SEQ; any = true; }
        else if (ln[0] == '@') { cur = Record(); cur.id = ln.substr(1); st = FQ_SEQ; any = true; }
        else ok = false;
        break;
      case FA_SEQ:
        if (!ln.empty() && ln[0] == '>') { out.push_back(std::move(cur)); cur = Record(); cur.id = ln.substr(1); }
        else cur.seq += ln;
        break;
      case FQ_SEQ: cur.seq = ln; fq_len = ln.size(); st = FQ_PLUS; break;
      case FQ_PLUS: if (ln.empty() || ln[0] != '+') ok = false; st = FQ_QUAL; break;
      case FQ_QUAL:
        if (ln.size() != fq_len) ok = false;
        out.push_back(std::move(cur)); cur = Record(); st = START;
        break;
    }
  };
  while (ok) {
    int got = gzread(f, buf.data(), (unsigned)buf.size());
    if (got < 0) { ok = false; break; }
    if (got == 0) break;
    size_t b = 0;
    for (int i = 0; i < got; i++) {
      if (buf[i] == '\n') {
        pending.append(buf.data() + b, i - b);
        flush_line(pending);
        pending.clear();
        b = i + 1;
        if (!ok) break;
      }
    }
    if (ok) pending.append(buf.data() + b, got - b);
  }
  gzclose(f);
  if (ok && !pending.empty()) flush_line(pending);
  if (ok && st == FA_SEQ) out.push_back(std::move(cur));
  if (ok && (st == FQ_SEQ || st == FQ_PLUS || st == FQ_QUAL)) ok = false;
  if (!any) ok = false;  // empty file (needletail: EmptyFile error)
  return ok;
}

}  // namespace fastx
