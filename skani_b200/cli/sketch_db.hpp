// sketch_db.hpp -- skani v0.3.0 on-disk sketch formats (host only; no CUDA).
//
// Byte layout = bincode 1.3 with default options (little endian, fixed-width integers, usize -> u64, u64 length
// prefixes, Option tag u8, bool u8, String = length + UTF-8 bytes) applied to the reference's serde structs:
//   (SketchParams, Sketch)                each `.sketch` file (src/sketch.rs:85) and each entry of `sketches.db`
//                                         (src/sketch_db.rs:45-47); SketchParams src/params.rs:137-146, Sketch
//                                         src/types.rs:253-277, SeedPosition src/types.rs:125-128
//   Vec<IndexEntry{file_name, offset, length}>   `index.db` (src/sketch_db.rs:10-15, 72-77)
//   (SketchParams, Vec<Sketch>)           `markers.bin`, sketches reduced by Sketch::get_markers_only
//                                         (src/types.rs:322-340, src/sketch.rs:141-146)
// The k-mer map is a Rust HashMap, so entry order in a file is arbitrary and carries no meaning; this writer emits
// ascending k-mers.  Map value (src/types.rs:207-244): bit 0 = 1 -> one position packed as
// ((pos << 31 | contig_index_canonical) << 1) | 1; bit 0 = 0 -> (index into multi_position_storage) << 1.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace skdb {

struct DiskParams {            // SketchParams, src/params.rs:137-146
  uint64_t c = 125, k = 15, marker_c = 1000;
  bool use_syncs = false, use_aa = false;
  uint64_t orf_size = 30;      // ORF_SIZE src/params.rs:32
  bool operator==(const DiskParams& o) const {
    return c == o.c && k == o.k && marker_c == o.marker_c && use_syncs == o.use_syncs && use_aa == o.use_aa && orf_size == o.orf_size;
  }
};

struct HostSketch {            // Sketch, src/types.rs:253-277 (records flattened: one entry per seed position)
  std::string file_name;
  bool has_seeds = true;       // kmer_seeds_k is Some(..) (false for get_markers_only)
  std::vector<uint32_t> kmer, pos, cc;      // cc = contig_index << 1 | canonical
  std::vector<std::string> contigs;
  uint64_t total_len = 0;
  std::vector<uint32_t> contig_lengths;
  uint64_t repetitive_kmers = 0;
  std::vector<uint64_t> markers;
  uint64_t marker_c = 125, c = 125, k = 15;  // marker_c field = c (quirk, src/types.rs:347)
  uint64_t contig_order = 0;
  bool individual_contig = false, amino_acid = false;
};

struct IndexEntry { std::string file_name; uint64_t offset = 0, length = 0; };

// ---------------------------------------------------------------- writer
struct Out {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u32(uint32_t v) { uint8_t t[4]; memcpy(t, &v, 4); b.insert(b.end(), t, t + 4); }
  void u64(uint64_t v) { uint8_t t[8]; memcpy(t, &v, 8); b.insert(b.end(), t, t + 8); }
  void str(const std::string& s) { u64(s.size()); b.insert(b.end(), s.begin(), s.end()); }
};

// DNA_TO_AA (src/types.rs:27-28) and its integer encoding (src/params.rs:150-180; the duplicated 'R' key keeps the later value 15)
inline const char* dna_to_aa() { return "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"; }
inline uint64_t aa_code(char a) {
  switch (a) {
    case 'A': return 0; case 'R': return 15; case 'N': return 2; case 'D': return 3; case 'C': return 4; case 'E': return 5;
    case 'F': return 6; case 'G': return 7; case 'H': return 8; case 'I': return 9; case 'K': return 10; case 'L': return 11;
    case 'M': return 12; case 'P': return 13; case 'Q': return 14; case 'S': return 16; case 'T': return 17; case 'V': return 18;
    case 'W': return 19; case 'Y': return 20; default: return 21;   // '*' = STOP_CODON src/params.rs:14
  }
}

inline void put_params(Out& o, const DiskParams& p) {
  o.u64(p.c); o.u64(p.k); o.u64(p.marker_c); o.u8(p.use_syncs); o.u8(p.use_aa);
  o.u64(64);
  for (int i = 0; i < 64; i++) o.u64(aa_code(dna_to_aa()[i]));
  o.u64(64);
  for (int i = 0; i < 64; i++) o.u8((uint8_t)dna_to_aa()[i]);
  o.u64(p.orf_size);
}

// records must be grouped by k-mer (any order inside a group); this is what sk_sketch_set_export returns
inline void put_sketch(Out& o, const HostSketch& s) {
  o.str(s.file_name);
  const size_t n = s.kmer.size();
  if (!s.has_seeds) {
    o.u8(0);
    o.u64(0);                                   // multi_position_storage
  } else {
    o.u8(1);
    size_t n_keys = 0, n_multi = 0;
    for (size_t i = 0; i < n;) {
      size_t j = i + 1;
      while (j < n && s.kmer[j] == s.kmer[i]) j++;
      n_keys++;
      if (j - i > 1) n_multi++;
      i = j;
    }
    o.u64(n_keys);
    size_t storage = 0;
    for (size_t i = 0; i < n;) {
      size_t j = i + 1;
      while (j < n && s.kmer[j] == s.kmer[i]) j++;
      o.u32(s.kmer[i]);
      if (j - i == 1) o.u64(((((uint64_t)s.pos[i] << 31) | (uint64_t)s.cc[i]) << 1) | 1ull);
      else o.u64((uint64_t)(storage++) << 1);
      i = j;
    }
    o.u64(n_multi);
    for (size_t i = 0; i < n;) {
      size_t j = i + 1;
      while (j < n && s.kmer[j] == s.kmer[i]) j++;
      if (j - i > 1) {
        o.u64(j - i);
        for (size_t t = i; t < j; t++) { o.u32(s.pos[t]); o.u32(s.cc[t]); }
      }
      i = j;
    }
  }
  o.u64(s.contigs.size());
  for (auto& c : s.contigs) o.str(c);
  o.u64(s.total_len);
  o.u64(s.contig_lengths.size());
  for (uint32_t l : s.contig_lengths) o.u32(l);
  o.u64(s.repetitive_kmers);
  o.u64(s.markers.size());
  for (uint64_t m : s.markers) o.u64(m);
  o.u64(s.marker_c); o.u64(s.c); o.u64(s.k); o.u64(s.contig_order);
  o.u8(s.individual_contig); o.u8(s.amino_acid);
}

inline HostSketch markers_only(const HostSketch& s) {   // Sketch::get_markers_only, src/types.rs:322-340
  HostSketch m;
  m.file_name = s.file_name; m.has_seeds = false; m.contigs = s.contigs; m.total_len = s.total_len;
  m.repetitive_kmers = s.repetitive_kmers; m.markers = s.markers; m.marker_c = s.marker_c; m.c = s.c; m.k = s.k;
  m.contig_order = s.contig_order; m.individual_contig = s.individual_contig; m.amino_acid = s.amino_acid;
  return m;
}

// ---------------------------------------------------------------- reader
struct In {
  const uint8_t* p; const uint8_t* e;
  In(const uint8_t* b, size_t n) : p(b), e(b + n) {}
  void need(size_t n) const { if ((size_t)(e - p) < n) throw std::runtime_error("truncated sketch data"); }
  uint8_t u8() { need(1); return *p++; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  uint64_t u64() { need(8); uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
  uint64_t len(size_t elem) { uint64_t n = u64(); if (elem && n > (uint64_t)(e - p) / elem) throw std::runtime_error("corrupt length prefix"); return n; }
  std::string str() { uint64_t n = len(1); std::string s((const char*)p, (size_t)n); p += n; return s; }
};

inline DiskParams get_params(In& in) {
  DiskParams p;
  p.c = in.u64(); p.k = in.u64(); p.marker_c = in.u64(); p.use_syncs = in.u8() != 0; p.use_aa = in.u8() != 0;
  uint64_t n = in.len(8); in.need(n * 8); in.p += n * 8;
  n = in.len(1); in.need(n); in.p += n;
  p.orf_size = in.u64();
  return p;
}

// seeds = false skips materialising the records (markers.bin entries have none anyway)
inline HostSketch get_sketch(In& in, bool seeds = true) {
  HostSketch s;
  s.file_name = in.str();
  const uint8_t tag = in.u8();
  if (tag > 1) throw std::runtime_error("corrupt Option tag (a pre-0.3 .sketch file?)");
  s.has_seeds = tag == 1;
  std::vector<uint32_t> keys;
  std::vector<uint64_t> vals;
  if (tag == 1) {
    uint64_t n = in.len(12);
    keys.resize(n); vals.resize(n);
    for (uint64_t i = 0; i < n; i++) { keys[i] = in.u32(); vals[i] = in.u64(); }
  }
  const uint64_t n_multi = in.len(8);
  std::vector<const uint8_t*> multi_at(n_multi);
  std::vector<uint64_t> multi_len(n_multi);
  for (uint64_t i = 0; i < n_multi; i++) {
    multi_len[i] = in.len(8);
    multi_at[i] = in.p;
    in.p += multi_len[i] * 8;
  }
  if (seeds) {
    for (size_t i = 0; i < keys.size(); i++) {
      if (vals[i] & 1) {
        const uint64_t packed = vals[i] >> 1;
        s.kmer.push_back(keys[i]); s.pos.push_back((uint32_t)(packed >> 31)); s.cc.push_back((uint32_t)(packed & 0x7FFFFFFFull));
      } else {
        const uint64_t si = vals[i] >> 1;
        if (si >= n_multi) throw std::runtime_error("multi-position index out of range");
        for (uint64_t t = 0; t < multi_len[si]; t++) {
          uint32_t a, b; memcpy(&a, multi_at[si] + 8 * t, 4); memcpy(&b, multi_at[si] + 8 * t + 4, 4);
          s.kmer.push_back(keys[i]); s.pos.push_back(a); s.cc.push_back(b);
        }
      }
    }
  }
  uint64_t n = in.len(8);
  for (uint64_t i = 0; i < n; i++) s.contigs.push_back(in.str());
  s.total_len = in.u64();
  n = in.len(4);
  s.contig_lengths.resize(n);
  for (uint64_t i = 0; i < n; i++) s.contig_lengths[i] = in.u32();
  s.repetitive_kmers = in.u64();
  n = in.len(8);
  s.markers.resize(n);
  for (uint64_t i = 0; i < n; i++) s.markers[i] = in.u64();
  s.marker_c = in.u64(); s.c = in.u64(); s.k = in.u64(); s.contig_order = in.u64();
  s.individual_contig = in.u8() != 0; s.amino_acid = in.u8() != 0;
  return s;
}

// ---------------------------------------------------------------- files
inline bool read_file(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
  fclose(f);
  return got == out.size();
}
inline bool write_file(const std::string& path, const std::vector<uint8_t>& b) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  size_t w = b.empty() ? 0 : fwrite(b.data(), 1, b.size(), f);
  return fclose(f) == 0 && w == b.size();
}

// consolidated database writer (SketchDbWriter, src/sketch_db.rs:17-84 + markers.bin, src/sketch.rs:141-146)
struct DbWriter {
  std::string dir;
  DiskParams params;
  FILE* concat = nullptr;
  std::vector<IndexEntry> index;
  std::vector<HostSketch> marker_sketches;
  uint64_t offset = 0;
  bool open(const std::string& d, const DiskParams& p) {
    dir = d; params = p;
    concat = fopen((dir + "/sketches.db").c_str(), "wb");
    return concat != nullptr;
  }
  bool add(const HostSketch& s) {
    Out o;
    put_params(o, params);
    put_sketch(o, s);
    if (fwrite(o.b.data(), 1, o.b.size(), concat) != o.b.size()) return false;
    index.push_back(IndexEntry{s.file_name, offset, (uint64_t)o.b.size()});
    offset += o.b.size();
    marker_sketches.push_back(markers_only(s));
    return true;
  }
  bool finalize() {
    if (fclose(concat) != 0) return false;
    concat = nullptr;
    Out ix;
    ix.u64(index.size());
    for (auto& e : index) { ix.str(e.file_name); ix.u64(e.offset); ix.u64(e.length); }
    if (!write_file(dir + "/index.db", ix.b)) return false;
    Out mk;
    put_params(mk, params);
    mk.u64(marker_sketches.size());
    for (auto& m : marker_sketches) put_sketch(mk, m);
    return write_file(dir + "/markers.bin", mk.b);
  }
};

// (SketchParams, Vec<Sketch>) of markers.bin (file_io::marker_sketches_from_marker_file, src/file_io.rs:719-729)
inline void read_markers_bin(const std::string& path, DiskParams& params, std::vector<HostSketch>& out) {
  std::vector<uint8_t> b;
  if (!read_file(path, b)) throw std::runtime_error("cannot read " + path);
  In in(b.data(), b.size());
  params = get_params(in);
  uint64_t n = in.len(8);
  out.clear();
  for (uint64_t i = 0; i < n; i++) out.push_back(get_sketch(in, false));
}

inline void read_index_db(const std::string& path, std::vector<IndexEntry>& out) {
  std::vector<uint8_t> b;
  if (!read_file(path, b)) throw std::runtime_error("cannot read " + path);
  In in(b.data(), b.size());
  uint64_t n = in.len(8);
  out.clear();
  for (uint64_t i = 0; i < n; i++) { IndexEntry e; e.file_name = in.str(); e.offset = in.u64(); e.length = in.u64(); out.push_back(e); }
}

// one (SketchParams, Sketch) blob: a `.sketch` file or a slice of sketches.db
inline HostSketch read_blob(const uint8_t* p, size_t n, DiskParams* params = nullptr) {
  In in(p, n);
  DiskParams dp = get_params(in);
  if (params) *params = dp;
  return get_sketch(in, true);
}

}  // namespace skdb
