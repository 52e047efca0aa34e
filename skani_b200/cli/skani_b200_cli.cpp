// skani_b200_cli.cpp -- `skani-b200 triangle|dist`: host driver over the C ABI (include/skani_b200.h).
//
// Mirrors the reference's command drivers for the two commands whose pair loops are the hot path:
//   triangle  src/triangle.rs:13-169  (flags src/cli.rs:236-330, defaults src/parse.rs:790-921)
//   dist      src/dist.rs:12-190      (flags src/cli.rs:100-232, defaults src/parse.rs:628-788)
// and their writers (TSV src/file_io.rs:15-139,608-678; phylip + .af matrices src/file_io.rs:364-539).
// FASTA/FASTQ(.gz) record rules follow needletail as skani uses it (ids = whole header line, sequences with line
// breaks removed, records < 500 bp dropped, src/file_io.rs:141-362).  `sketch` / `search` (on-disk sketch DB formats,
// src/sketch_db.rs) are not implemented yet.  All heavy work happens on the GPU through libskani_b200.so.
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/skani_b200.h"

namespace {

struct Record { std::string id; std::string seq; };

// ---- FASTA / FASTQ reader (streaming over zlib; plain files pass through gzread unchanged) -----------------
bool read_fastx(const std::string& path, std::vector<Record>& out) {
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

struct Genome {          // one Sketch-to-be (src/types.rs:253-277 metadata kept on the host)
  std::string file_name;
  std::vector<std::string> contigs;  // header lines
  uint64_t contig_order = 0;
  uint64_t total_len = 0;
};

struct Inputs {
  std::vector<Genome> genomes;       // sorted by (file_name, contig_order) (src/types.rs:360-364)
  std::vector<uint8_t> bases;
  std::vector<uint64_t> contig_off{0};
  std::vector<uint32_t> genome_of_contig;
};

// file_io::fastx_to_sketches / fastx_to_multiple_sketch_rewrite record rules (src/file_io.rs:141-362)
void load_inputs(std::vector<std::string> files, bool individual, int threads, Inputs& in) {
  std::sort(files.begin(), files.end());                      // final order = (file_name, contig_order)
  std::vector<std::vector<Record>> recs(files.size());
  std::vector<int> status(files.size(), 0);
  std::vector<std::thread> pool;
  std::vector<size_t> next(1, 0);
  auto worker = [&](int tid) {
    for (size_t i = tid; i < files.size(); i += threads) status[i] = read_fastx(files[i], recs[i]) ? 1 : -1;
  };
  for (int t = 0; t < threads; t++) pool.emplace_back(worker, t);
  for (auto& t : pool) t.join();
  for (size_t i = 0; i < files.size(); i++) {
    if (status[i] < 0) { fprintf(stderr, "WARN %s is not a valid fasta/fastq file; skipping.\n", files[i].c_str()); continue; }
    size_t kept = 0;
    if (!individual) {
      Genome g; g.file_name = files[i];
      for (auto& r : recs[i]) {
        if (r.seq.size() < 500) continue;                     // MIN_LENGTH_CONTIG (src/params.rs:42, src/file_io.rs:176)
        g.contigs.push_back(r.id); g.total_len += r.seq.size();
        in.bases.insert(in.bases.end(), r.seq.begin(), r.seq.end());
        in.contig_off.push_back(in.bases.size());
        in.genome_of_contig.push_back((uint32_t)in.genomes.size());
        kept++;
      }
      if (kept) in.genomes.push_back(std::move(g));
      else fprintf(stderr, "WARN File %s consists of only contigs < 500 bp. Skipping this file.\n", files[i].c_str());
    } else {
      bool warned = false;
      for (auto& r : recs[i]) {
        if (r.seq.size() < 500) {
          if (!warned) { fprintf(stderr, "WARN At least one sequence in file %s has < 500 bp. These sequences will be skipped.\n", files[i].c_str()); warned = true; }
          continue;
        }
        Genome g; g.file_name = files[i]; g.contigs.push_back(r.id); g.total_len = r.seq.size(); g.contig_order = kept++;
        in.bases.insert(in.bases.end(), r.seq.begin(), r.seq.end());
        in.contig_off.push_back(in.bases.size());
        in.genome_of_contig.push_back((uint32_t)in.genomes.size());
        in.genomes.push_back(std::move(g));
      }
    }
    recs[i].clear(); recs[i].shrink_to_fit();
  }
}

std::vector<std::string> read_list(const std::string& path) {
  std::vector<std::string> v;
  FILE* f = fopen(path.c_str(), "r");
  if (!f) { fprintf(stderr, "ERROR cannot open list file %s\n", path.c_str()); exit(1); }
  char line[1 << 16];
  while (fgets(line, sizeof(line), f)) {
    std::string s(line);
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
    if (!s.empty()) v.push_back(s);
  }
  fclose(f);
  return v;
}

std::string short_name(const std::string& s, bool short_header) {   // truncate_contig_name (src/types.rs:197-203)
  if (!short_header) return s;
  size_t b = s.find_first_not_of(" \t");
  if (b == std::string::npos) return s;
  size_t e = s.find_first_of(" \t", b);
  return s.substr(b, e == std::string::npos ? std::string::npos : e - b);
}

std::string f32_display(float v) {   // Rust `{}` / `{:0}` for an f32 holding an integer value
  if (v == std::floor(v) && std::fabs(v) < 1e15) { char b[64]; snprintf(b, sizeof(b), "%.0f", (double)v); return b; }
  char b[64]; snprintf(b, sizeof(b), "%g", (double)v); return b;
}

struct Opts {
  std::string cmd, out;
  std::vector<std::string> files, queries, refs;
  uint32_t c = 125, k = 15, m = 1000;
  bool c_set = false, m_set = false;
  double s = 0.0, min_af = -1e9, both_min_af = -1.0;
  bool individual = false, qi = false, ri = false, sparse = false, full_matrix = false, diagonal = false, ci = false, detailed = false,
       short_header = false, distance = false, robust = false, median = false, no_learned = false, faster_small = false,
       small_genomes = false, fast = false, medium = false, slow = false, no_marker_index = false;
  uint64_t n = 1000000000000ull;
  int threads = 3, device = 0;
};

void write_header(FILE* o, bool ci, bool detailed) {   // src/file_io.rs:15-23
  if (!ci && !detailed) fprintf(o, "Ref_file\tQuery_file\tANI\tAlign_fraction_ref\tAlign_fraction_query\tRef_name\tQuery_name\n");
  else if (!detailed) fprintf(o, "Ref_file\tQuery_file\tANI\tAlign_fraction_ref\tAlign_fraction_query\tRef_name\tQuery_name\tANI_5_percentile\tANI_95_percentile\n");
  else fprintf(o, "Ref_file\tQuery_file\tANI\tAlign_fraction_ref\tAlign_fraction_query\tRef_name\tQuery_name\tNum_ref_contigs\tNum_query_contigs\t"
                  "ANI_5_percentile\tANI_95_percentile\tStandard_deviation\tRef_90_ctg_len\tRef_50_ctg_len\tRef_10_ctg_len\tQuery_90_ctg_len\t"
                  "Query_50_ctg_len\tQuery_10_ctg_len\tAvg_chain_len\tTotal_bases_covered\n");
}
void write_row(FILE* o, const sk_ani_result& r, const Genome& ref, const Genome& qry, const Opts& op) {   // write_ani_res, src/file_io.rs:83-139
  fprintf(o, "%s\t%s\t%.2f\t%.2f\t%.2f\t%s\t%s", ref.file_name.c_str(), qry.file_name.c_str(), (double)(r.ani * 100.f),
          (double)(r.af_ref * 100.f), (double)(r.af_query * 100.f), short_name(ref.contigs[0], op.short_header).c_str(),
          short_name(qry.contigs[0], op.short_header).c_str());
  if (op.detailed) {
    fprintf(o, "\t%u\t%u\t%.2f\t%.2f\t%.2f\t%s\t%s\t%s\t%s\t%s\t%s\t%u\t%u", r.num_contigs_r, r.num_contigs_q, (double)(r.ci_lower * 100.f),
            (double)(r.ci_upper * 100.f), (double)(r.std * 100.f), f32_display(r.q90_r).c_str(), f32_display(r.q50_r).c_str(),
            f32_display(r.q10_r).c_str(), f32_display(r.q90_q).c_str(), f32_display(r.q50_q).c_str(), f32_display(r.q10_q).c_str(),
            r.avg_chain_int_len, r.total_bases_covered);
  } else if (op.ci) {
    fprintf(o, "\t%.2f\t%.2f", (double)(r.ci_lower * 100.f), (double)(r.ci_upper * 100.f));
  }
  fputc('\n', o);
}
void write_perfect(FILE* o, const Genome& g, const Opts& op) {   // write_ani_res_perfect, src/file_io.rs:25-81
  std::string nm = short_name(g.contigs[0], op.short_header);
  fprintf(o, "%s\t%s\t100.00\t100.00\t100.00\t%s\t%s", g.file_name.c_str(), g.file_name.c_str(), nm.c_str(), nm.c_str());
  if (op.detailed) fprintf(o, "\t%zu\t%zu\t100.00\t100.00\t0.00\t-1\t-1\t-1\t-1\t-1\t-1\t0\t%llu", g.contigs.size(), g.contigs.size(), (unsigned long long)g.total_len);
  else if (op.ci) fprintf(o, "\t100.00\t100.00");
  fputc('\n', o);
}

#define CK(ctx, call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "ERROR %s failed (%d): %s\n", #call, rc__, sk_last_error(ctx)); exit(1); } } while (0)

sk_sketch_set* sketch(sk_ctx* ctx, const Inputs& in, const sk_sketch_params& sp) {
  sk_sketch_set* set = nullptr;
  CK(ctx, sk_sketch_batch(ctx, in.bases.data(), in.contig_off.data(), (uint32_t)in.genome_of_contig.size(), in.genome_of_contig.data(),
                          (uint32_t)in.genomes.size(), &sp, &set));
  return set;
}

void resolve_presets(Opts& op) {   // src/parse.rs:829-853 / 680-710
  if (op.fast && op.slow) { fprintf(stderr, "ERROR Both --slow and --fast were set. This is not allowed.\n"); exit(1); }
  if (op.fast) op.c = 200;
  if (op.slow) op.c = 30;
  if (op.medium) op.c = 70;
  if (op.small_genomes) { op.c = 30; op.m = 200; }
  if (op.c > op.m) { fprintf(stderr, "ERROR We currently don't allow c (%u) > m (%u). -m should be larger than c.\n", op.c, op.m); exit(1); }  // src/params.rs:183
}

int run_triangle(Opts& op) {
  resolve_presets(op);
  if (op.files.empty()) { fprintf(stderr, "ERROR No reference inputs found.\n"); return 1; }
  Inputs in;
  load_inputs(op.files, op.individual, std::max(op.threads, 1), in);
  if (in.genomes.empty()) { fprintf(stderr, "ERROR No genomes/sketches found.\n"); return 1; }   // src/triangle.rs:46-49
  if (in.genomes.size() > 500 && !op.sparse) fprintf(stderr, "WARN > 500 genomes detected. The output matrix will be large. Consider using -E or --sparse for a tsv output instead.\n");
  sk_ctx* ctx = nullptr;
  if (sk_ctx_create(op.device, &ctx) != 0) { fprintf(stderr, "ERROR a CUDA device is required (no CPU fallback)\n"); return 1; }
  sk_sketch_params sp{op.c, op.k, op.m};
  sk_map_params mp{};
  mp.screen_val = op.s / 100.0;
  mp.min_aligned_frac = (op.min_af > -1e8 ? op.min_af : 15.0) / 100.0;
  mp.both_min_aligned_frac = op.both_min_af / 100.0;
  mp.robust = op.robust; mp.median = op.median;
  mp.rescue_small = !op.faster_small && !op.small_genomes;
  mp.learned_ani = !op.no_learned && op.c >= 70 && !op.individual && !op.median;   // regression::use_learned_ani (src/regression.rs:8-10)
  if (mp.learned_ani) fprintf(stderr, "INFO Learned ANI mode detected. ANI may be adjusted according to a regression model trained on MAGs.\n");
  sk_sketch_set* set = sketch(ctx, in, sp);
  {  // file-name order for the switch_qr tie-break (src/chain.rs:19-21): with -i all records of a file share its name
    std::vector<uint64_t> ranks(in.genomes.size());
    uint64_t rank = 0;
    for (size_t i = 0; i < in.genomes.size(); i++) {
      if (i && in.genomes[i].file_name != in.genomes[i - 1].file_name) rank++;
      ranks[i] = rank;
    }
    sk_sketch_set_set_name_ranks(set, ranks.data());
  }
  uint64_t* pairs = nullptr; uint64_t np = 0;
  CK(ctx, sk_screen_triangle(ctx, set, &mp, &pairs, &np));
  std::vector<sk_ani_result> res(np);
  CK(ctx, sk_chain_pairs(ctx, set, set, pairs, np, &mp, res.data()));
  sk_free(pairs);
  const size_t N = in.genomes.size();
  FILE* o = op.out.empty() ? stdout : fopen(op.out.c_str(), "w");
  if (!o) { fprintf(stderr, "ERROR cannot open %s\n", op.out.c_str()); return 1; }
  if (op.sparse) {   // write_sparse_matrix (src/file_io.rs:541-606); rows emitted in (i, j) order (the reference's order is arbitrary)
    write_header(o, op.ci, op.detailed);
    if (op.diagonal) for (auto& g : in.genomes) write_perfect(o, g, op);
    for (auto& r : res) if (r.ani > 0.1f) write_row(o, r, in.genomes[r.ref_id], in.genomes[r.query_id], op);
  } else {           // write_phyllip_matrix (src/file_io.rs:364-539)
    std::map<std::pair<uint32_t, uint32_t>, const sk_ani_result*> m;
    for (auto& r : res) if (r.ani > 0.1f) m[{r.ref_id, r.query_id}] = &r;
    const double perfect = op.distance ? 0. : 100., none = 100. - perfect;
    std::string af_name = op.out.empty() ? "skani_matrix.af" : op.out + ".af";
    FILE* af = fopen(af_name.c_str(), "w");
    fprintf(o, "%zu\n", N);
    if (af) fprintf(af, "%zu\n", N);
    for (size_t i = 0; i < N; i++) {
      const std::string& name = op.individual ? in.genomes[i].contigs[0] : in.genomes[i].file_name;
      fputs(name.c_str(), o);
      if (af) fputs(name.c_str(), af);
      for (size_t j = 0; j < N; j++) {
        bool full_cond = op.full_matrix || (i > j);
        auto it = m.find({(uint32_t)std::min(i, j), (uint32_t)std::max(i, j)});
        if (i == j) {
          if (full_cond || op.diagonal) fprintf(o, "\t%.2f", perfect);
          if (af) fprintf(af, "\t%.2f", 100.);
          continue;
        }
        if (it == m.end()) {
          if (full_cond) fprintf(o, "\t%.2f", none);
          if (af) fprintf(af, "\t%.2f", 0.);
        } else {
          if (full_cond) { double val = (double)(it->second->ani * 100.f); fprintf(o, "\t%.2f", op.distance ? 100. - val : val); }
          if (af) fprintf(af, "\t%.2f", (double)((j > i ? it->second->af_ref : it->second->af_query) * 100.f));
        }
      }
      fputc('\n', o);
      if (af) fputc('\n', af);
    }
    if (af) fclose(af);
    fprintf(stderr, "INFO Aligned fraction matrix written to %s\n", af_name.c_str());
  }
  if (o != stdout) fclose(o);
  sk_sketch_set_free(set);
  sk_ctx_destroy(ctx);
  return 0;
}

int run_dist(Opts& op) {
  resolve_presets(op);
  if (op.refs.empty() || op.queries.empty()) { fprintf(stderr, "ERROR No reference sketches/genomes or query sketches/genomes found.\n"); return 1; }
  Inputs rin, qin;
  load_inputs(op.refs, op.ri, std::max(op.threads, 1), rin);
  load_inputs(op.queries, op.qi, std::max(op.threads, 1), qin);
  if (rin.genomes.empty() || qin.genomes.empty()) { fprintf(stderr, "ERROR No reference sketches/genomes or query sketches/genomes found.\n"); return 1; }
  sk_ctx* ctx = nullptr;
  if (sk_ctx_create(op.device, &ctx) != 0) { fprintf(stderr, "ERROR a CUDA device is required (no CPU fallback)\n"); return 1; }
  sk_sketch_params sp{op.c, op.k, op.m};
  sk_map_params mp{};
  mp.screen_val = op.s / 100.0;
  mp.min_aligned_frac = (op.min_af > -1e8 ? op.min_af : 15.0) / 100.0;
  mp.both_min_aligned_frac = op.both_min_af / 100.0;
  mp.robust = op.robust; mp.median = op.median;
  mp.rescue_small = !op.faster_small && !op.small_genomes;
  mp.learned_ani = !op.no_learned && op.c >= 70 && !op.qi && !op.ri && !op.median;
  if (mp.learned_ani) fprintf(stderr, "INFO Learned ANI mode detected. ANI may be adjusted according to a regression model trained on MAGs.\n");
  const bool use_index = (op.queries.size() > 50 || op.qi) && !op.no_marker_index;   // FULL_INDEX_THRESH (src/parse.rs:750)
  sk_sketch_set* rset = sketch(ctx, rin, sp);
  sk_sketch_set* qset = sketch(ctx, qin, sp);
  // file-name order for the switch_qr tie-break (src/chain.rs:19-21): rank all names together
  {
    std::vector<std::pair<std::string, std::pair<int, size_t>>> names;
    for (size_t i = 0; i < rin.genomes.size(); i++) names.push_back({rin.genomes[i].file_name, {0, i}});
    for (size_t i = 0; i < qin.genomes.size(); i++) names.push_back({qin.genomes[i].file_name, {1, i}});
    std::sort(names.begin(), names.end(), [](auto& a, auto& b) { return a.first < b.first; });
    std::vector<uint64_t> rr(rin.genomes.size()), qr(qin.genomes.size());
    uint64_t rank = 0;
    for (size_t i = 0; i < names.size(); i++) {
      if (i && names[i].first != names[i - 1].first) rank++;
      (names[i].second.first ? qr : rr)[names[i].second.second] = rank;
    }
    sk_sketch_set_set_name_ranks(rset, rr.data());
    sk_sketch_set_set_name_ranks(qset, qr.data());
  }
  uint64_t* pairs = nullptr; uint64_t np = 0;
  CK(ctx, sk_screen_query_ref(ctx, rset, qset, &mp, use_index ? 2 : 0, &pairs, &np));
  std::vector<sk_ani_result> res(np);
  CK(ctx, sk_chain_pairs(ctx, rset, qset, pairs, np, &mp, res.data()));
  sk_free(pairs);
  // write_query_ref_list (src/file_io.rs:608-678): group by the query's first contig name, sort each group by ANI desc, top n
  std::map<std::string, std::vector<const sk_ani_result*>> groups;
  for (auto& r : res) if (r.ani > 0.1f) groups[qin.genomes[r.query_id].contigs[0]].push_back(&r);
  FILE* o = op.out.empty() ? stdout : fopen(op.out.c_str(), "w");
  if (!o) { fprintf(stderr, "ERROR cannot open %s\n", op.out.c_str()); return 1; }
  write_header(o, op.ci, op.detailed);
  for (auto& kv : groups) {
    auto v = kv.second;
    std::stable_sort(v.begin(), v.end(), [](const sk_ani_result* a, const sk_ani_result* b) { return a->ani > b->ani; });
    for (size_t i = 0; i < v.size() && i < op.n; i++) write_row(o, *v[i], rin.genomes[v[i]->ref_id], qin.genomes[v[i]->query_id], op);
  }
  if (o != stdout) fclose(o);
  sk_sketch_set_free(rset); sk_sketch_set_free(qset);
  sk_ctx_destroy(ctx);
  return 0;
}

void usage() {
  fprintf(stderr,
          "skani-b200 (Blackwell implementation of skani v0.3.0's ANI hot path)\n"
          "  skani-b200 triangle [fasta ... | -l list] [-i] [-E|--sparse] [-o out] [--full-matrix] [--diagonal] [--distance]\n"
          "  skani-b200 dist [query] [refs ...] [-q ...] [-r ...] [--ql list] [--rl list] [--qi] [--ri] [-n N] [-o out]\n"
          "  common: -c C -m M -k K -s SCREEN%% --min-af P --both-min-af P --robust --median --no-learned-ani --faster-small\n"
          "          --small-genomes --fast --medium --slow --ci --detailed --short-header --no-marker-index -t THREADS --device D\n");
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { usage(); return 2; }
  Opts op;
  op.cmd = argv[1];
  if (op.cmd != "triangle" && op.cmd != "dist") { usage(); return 2; }
  std::vector<std::string> positional;
  enum { NONE, QS, RS } multi = NONE;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "ERROR missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a[0] != '-') {
      if (multi == QS) op.queries.push_back(a); else if (multi == RS) op.refs.push_back(a); else positional.push_back(a);
      continue;
    }
    multi = NONE;
    if (a == "-c") { op.c = (uint32_t)atoi(val().c_str()); op.c_set = true; }
    else if (a == "-m") { op.m = (uint32_t)atof(val().c_str()); op.m_set = true; }
    else if (a == "-k") op.k = (uint32_t)atoi(val().c_str());
    else if (a == "-s") op.s = atof(val().c_str());
    else if (a == "-t") op.threads = atoi(val().c_str());
    else if (a == "-o") op.out = val();
    else if (a == "-n") op.n = strtoull(val().c_str(), nullptr, 10);
    else if (a == "-l") { auto v = read_list(val()); op.files.insert(op.files.end(), v.begin(), v.end()); }
    else if (a == "--ql") { auto v = read_list(val()); op.queries.insert(op.queries.end(), v.begin(), v.end()); }
    else if (a == "--rl") { auto v = read_list(val()); op.refs.insert(op.refs.end(), v.begin(), v.end()); }
    else if (a == "-q") multi = QS;
    else if (a == "-r") multi = RS;
    else if (a == "-i") op.individual = true;
    else if (a == "--qi") op.qi = true;
    else if (a == "--ri") op.ri = true;
    else if (a == "-E" || a == "--sparse") op.sparse = true;
    else if (a == "--full-matrix") op.full_matrix = true;
    else if (a == "--diagonal") op.diagonal = true;
    else if (a == "--min-af") op.min_af = atof(val().c_str());
    else if (a == "--both-min-af") op.both_min_af = atof(val().c_str());
    else if (a == "--ci") op.ci = true;
    else if (a == "--detailed") op.detailed = true;
    else if (a == "--short-header") op.short_header = true;
    else if (a == "--distance") op.distance = true;
    else if (a == "--robust") op.robust = true;
    else if (a == "--median") op.median = true;
    else if (a == "--no-learned-ani") op.no_learned = true;
    else if (a == "--faster-small") op.faster_small = true;
    else if (a == "--small-genomes") op.small_genomes = true;
    else if (a == "--fast") op.fast = true;
    else if (a == "--medium") op.medium = true;
    else if (a == "--slow") op.slow = true;
    else if (a == "--no-marker-index") op.no_marker_index = true;
    else if (a == "--device") op.device = atoi(val().c_str());
    else if (a == "-v" || a == "--debug" || a == "--trace") {}
    else { fprintf(stderr, "ERROR unknown option %s\n", a.c_str()); usage(); return 2; }
  }
  if (op.cmd == "triangle") {
    op.files.insert(op.files.end(), positional.begin(), positional.end());
    return run_triangle(op);
  }
  // dist: first positional is the query, the rest are references (src/cli.rs:115-121)
  if (!positional.empty()) {
    if (op.queries.empty()) { op.queries.push_back(positional[0]); op.refs.insert(op.refs.end(), positional.begin() + 1, positional.end()); }
    else op.refs.insert(op.refs.end(), positional.begin(), positional.end());
  }
  return run_dist(op);
}
