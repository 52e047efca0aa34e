// ============================================================================
// oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points over skani_oracle.cpp
// for the Python test-suite / bench CPU-baseline leg (ctypes).  Not part of the product.
// ============================================================================
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <algorithm>
#include <malloc.h>

#include "skani_oracle.hpp"

// The CPU baseline should not be handicapped by glibc malloc defaults under 100+ threads: multi-100 KB vectors would be
// mmap'ed/munmap'ed on every pair (serialising on the process' mm lock).  Keep them on the per-thread arenas instead.
namespace {
struct MallocTuning {
  MallocTuning() {
    mallopt(M_MMAP_THRESHOLD, 512 << 20);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 512);
  }
} g_malloc_tuning;
}  // namespace

using namespace orc;

extern "C" {

struct orc_result {  // same field order as sk_ani_result in include/skani_b200.h
  float ani, af_query, af_ref, ci_lower, ci_upper, std;
  float q90_q, q90_r, q50_q, q50_r, q10_q, q10_r;
  uint32_t num_contigs_q, num_contigs_r, avg_chain_int_len, total_bases_covered;
  uint32_t ref_id, query_id;
};

struct orc_cmd {
  double screen_val, min_aligned_frac, both_min_aligned_frac;
  int32_t robust, median, learned_ani, rescue_small;
};

static CommandParams to_cp(const orc_cmd* c) {
  CommandParams cp;
  cp.screen_val = c->screen_val;
  cp.min_aligned_frac = c->min_aligned_frac;
  cp.both_min_aligned_frac = c->both_min_aligned_frac;
  cp.robust = c->robust; cp.median = c->median; cp.learned_ani = c->learned_ani; cp.rescue_small = c->rescue_small;
  return cp;
}
static void to_res(const AniEstResult& a, uint32_t rid, uint32_t qid, orc_result* o) {
  o->ani = a.ani; o->af_query = a.align_fraction_query; o->af_ref = a.align_fraction_ref;
  o->ci_lower = a.ci_lower; o->ci_upper = a.ci_upper; o->std = a.std;
  o->q90_q = a.quant_90_contig_len_q; o->q90_r = a.quant_90_contig_len_r;
  o->q50_q = a.quant_50_contig_len_q; o->q50_r = a.quant_50_contig_len_r;
  o->q10_q = a.quant_10_contig_len_q; o->q10_r = a.quant_10_contig_len_r;
  o->num_contigs_q = a.num_contigs_q; o->num_contigs_r = a.num_contigs_r;
  o->avg_chain_int_len = a.avg_chain_int_len; o->total_bases_covered = a.total_bases_covered;
  o->ref_id = rid; o->query_id = qid;
}

uint64_t orc_mm_hash64(uint64_t x) { return mm_hash64(x); }
int orc_set_avx2_intrinsics(int on) { set_avx2_intrinsics(on != 0); return avx2_intrinsics() ? 1 : 0; }

// ---- sketching ------------------------------------------------------------------------------------
// Sketch every file; returns a malloc'd array of handles in (file_name, contig_order) order.
int orc_sketch_files(const char** paths, int n, uint64_t c, uint64_t k, uint64_t marker_c, int individual,
                     int avx2sem, int threads, void*** out, int* n_out, int* n_warn) {
  std::vector<std::string> files(paths, paths + n);
  SketchParams sp; sp.c = c; sp.k = k; sp.marker_c = marker_c;
  if (c > marker_c) return -1;  // params.rs:183-185 (reference panics)
  std::vector<std::string> warn;
  std::vector<Sketch> v = fastx_to_sketches(files, sp, individual != 0, avx2sem != 0, threads, &warn);
  void** arr = (void**)malloc(sizeof(void*) * (v.size() + 1));
  for (size_t i = 0; i < v.size(); i++) arr[i] = new Sketch(std::move(v[i]));
  *out = arr; *n_out = (int)v.size();
  if (n_warn) *n_warn = (int)warn.size();
  return 0;
}
void orc_free_array(void* p) { free(p); }

// contigs given as one concatenated ASCII buffer + n_contigs+1 offsets
void* orc_sketch_from_contigs(const char* name, const uint8_t* bases, const uint64_t* off, uint32_t n_contigs,
                              uint64_t c, uint64_t k, uint64_t marker_c, int avx2sem) {
  SketchParams sp; sp.c = c; sp.k = k; sp.marker_c = marker_c;
  std::vector<std::pair<const uint8_t*, size_t>> ctgs;
  for (uint32_t i = 0; i < n_contigs; i++) ctgs.push_back({bases + off[i], (size_t)(off[i + 1] - off[i])});
  return new Sketch(sketch_from_contigs(name, ctgs, nullptr, sp, avx2sem != 0));
}
void orc_sketch_free(void* s) { delete (Sketch*)s; }

// time seeding only (CPU baseline leg): seeds n_contigs contigs `reps` times, returns seconds
double orc_time_seeding(const uint8_t* bases, const uint64_t* off, uint32_t n_contigs, uint64_t c, uint64_t k,
                        uint64_t marker_c, int threads) {
  SketchParams sp; sp.c = c; sp.k = k; sp.marker_c = marker_c;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads > 0 ? threads : 1)
  for (long i = 0; i < (long)n_contigs; i++) {
    Sketch sk;
    sk.c = c; sk.k = k;
    fmh_seeds_avx2sem(bases + off[i], (size_t)(off[i + 1] - off[i]), sp, 0, sk);
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

uint64_t orc_sketch_n_records(void* s) {
  Sketch* sk = (Sketch*)s;
  uint64_t n = 0;
  const KmerSeeds& m = sk->kmer_seeds_k;
  for (size_t i = 0; i < m.capacity(); i++) {
    if (!m.slot_used(i)) continue;
    uint64_t t = m.slot_val(i);
    n += (t & 1) ? 1 : sk->multi_position_storage[(size_t)(t >> 1)].size();
  }
  return n;
}
uint64_t orc_sketch_n_kmers(void* s) { return ((Sketch*)s)->kmer_seeds_k.size(); }
uint64_t orc_sketch_n_markers(void* s) { return ((Sketch*)s)->marker_seeds.size(); }
uint64_t orc_sketch_n_contigs(void* s) { return ((Sketch*)s)->contigs.size(); }
uint64_t orc_sketch_total_len(void* s) { return ((Sketch*)s)->total_sequence_length; }
const char* orc_sketch_file_name(void* s) { return ((Sketch*)s)->file_name.c_str(); }
const char* orc_sketch_contig_name(void* s, uint64_t i) { return ((Sketch*)s)->contigs[i].c_str(); }
uint64_t orc_sketch_contig_order(void* s) { return ((Sketch*)s)->contig_order; }

// export as flat arrays: records sorted by (kmer, contig, pos); markers ascending; contig lengths in order
void orc_sketch_export(void* s, uint32_t* kmer, uint32_t* pos, uint32_t* cc, uint64_t* markers, uint32_t* contig_lengths) {
  Sketch* sk = (Sketch*)s;
  struct Rec { uint32_t kmer, pos, cc; };
  std::vector<Rec> recs;
  const KmerSeeds& m = sk->kmer_seeds_k;
  SeedPosition tmp;
  for (size_t i = 0; i < m.capacity(); i++) {
    if (!m.slot_used(i)) continue;
    const SeedPosition* p;
    size_t n = sk->get_seed_positions(m.slot_key(i), &p, &tmp);
    for (size_t a = 0; a < n; a++) recs.push_back({m.slot_key(i), p[a].pos, p[a].contig_index_canonical});
  }
  std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
    if (a.kmer != b.kmer) return a.kmer < b.kmer;
    if ((a.cc >> 1) != (b.cc >> 1)) return (a.cc >> 1) < (b.cc >> 1);
    if (a.pos != b.pos) return a.pos < b.pos;
    return a.cc < b.cc;
  });
  for (size_t i = 0; i < recs.size(); i++) { kmer[i] = recs[i].kmer; pos[i] = recs[i].pos; cc[i] = recs[i].cc; }
  std::vector<uint64_t> mk;
  const MarkerSet& ms = sk->marker_seeds;
  for (size_t i = 0; i < ms.capacity(); i++) if (ms.slot_used(i)) mk.push_back(ms.slot_key(i));
  std::sort(mk.begin(), mk.end());
  for (size_t i = 0; i < mk.size(); i++) markers[i] = mk[i];
  for (size_t i = 0; i < sk->contig_lengths.size(); i++) contig_lengths[i] = sk->contig_lengths[i];
}

// ---- screen -----------------------------------------------------------------------------------------
int orc_check_markers_quickly(void* ref, void* query, double screen_val, int rescue_small) {
  return check_markers_quickly(*(Sketch*)ref, *(Sketch*)query, screen_val, rescue_small != 0) ? 1 : 0;
}
// rows of the triangle screen: for every i the ascending list of j > i that pass screen_refs(i). CSR output.
int orc_screen_triangle(void** sk, int n, double screen_val, int rescue_small, uint64_t** row_off, uint32_t** cols) {
  std::vector<const Sketch*> v;
  for (int i = 0; i < n; i++) v.push_back((Sketch*)sk[i]);
  double sv = screen_val == 0. ? SEARCH_ANI_CUTOFF_DEFAULT : screen_val;
  KmerToSketch* idx = kmer_to_sketch_from_refs(v);
  std::vector<std::vector<uint32_t>> rows(n);
  for (int i = 0; i + 1 < n; i++) {
    auto pass = screen_refs(sv, *idx, *v[i], v, rescue_small != 0);
    for (uint32_t j : pass) if ((int)j > i) rows[i].push_back(j);
  }
  kmer_to_sketch_free(idx);
  uint64_t* ro = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
  uint64_t tot = 0;
  for (int i = 0; i < n; i++) { ro[i] = tot; tot += rows[i].size(); }
  ro[n] = tot;
  uint32_t* cs = (uint32_t*)malloc(sizeof(uint32_t) * (tot + 1));
  for (int i = 0; i < n; i++) memcpy(cs + ro[i], rows[i].data(), rows[i].size() * 4);
  *row_off = ro; *cols = cs;
  return 0;
}

// ---- chain ------------------------------------------------------------------------------------------
int orc_chain(void* ref, void* query, const orc_cmd* cmd, orc_result* out) {
  const Sketch& r = *(Sketch*)ref;
  const Sketch& q = *(Sketch*)query;
  CommandParams cp = to_cp(cmd);
  MapParams mp = map_params_from_sketch(r, cp, get_model_id(r.c, cp.learned_ani));
  to_res(chain_seeds(r, q, mp), 0, 0, out);
  return 0;
}

struct orc_debug {
  ChainDebug d;
  orc_result res;
};
void* orc_chain_debug(void* ref, void* query, const orc_cmd* cmd) {
  const Sketch& r = *(Sketch*)ref;
  const Sketch& q = *(Sketch*)query;
  CommandParams cp = to_cp(cmd);
  MapParams mp = map_params_from_sketch(r, cp, get_model_id(r.c, cp.learned_ani));
  auto* d = new orc_debug();
  to_res(chain_seeds(r, q, mp, &d->d), 0, 0, &d->res);
  return d;
}
void orc_debug_free(void* d) { delete (orc_debug*)d; }
void orc_debug_result(void* d, orc_result* out) { *out = ((orc_debug*)d)->res; }
int orc_debug_switched(void* d) { return ((orc_debug*)d)->d.switched ? 1 : 0; }
uint64_t orc_debug_n_anchors(void* d) { return ((orc_debug*)d)->d.anchors.size(); }
uint64_t orc_debug_n_chunks(void* d) { auto& c = ((orc_debug*)d)->d.chunk_first; return c.empty() ? 0 : c.size() - 1; }
uint64_t orc_debug_n_intervals(void* d) { return ((orc_debug*)d)->d.intervals_all.size(); }
uint64_t orc_debug_n_ests(void* d) { return ((orc_debug*)d)->d.ani_ests.size(); }
// anchors: 5 x u32 per anchor (qcontig,qpos,rcontig,rpos,rev); score i64; pointer u32 (chunk-local)
void orc_debug_anchors(void* dd, uint32_t* a5, int64_t* score, uint32_t* ptr) {
  auto& d = ((orc_debug*)dd)->d;
  for (size_t i = 0; i < d.anchors.size(); i++) {
    const Anchor& a = d.anchors[i];
    a5[5 * i] = a.query_contig; a5[5 * i + 1] = a.query_pos; a5[5 * i + 2] = a.ref_contig; a5[5 * i + 3] = a.ref_pos;
    a5[5 * i + 4] = a.reverse_match;
    score[i] = (int64_t)d.score[i];
    ptr[i] = d.pointer[i];
  }
}
void orc_debug_chunks(void* dd, uint32_t* first /* n_chunks+1 */, uint32_t* nseeds) {
  auto& d = ((orc_debug*)dd)->d;
  for (size_t i = 0; i < d.chunk_first.size(); i++) first[i] = d.chunk_first[i];
  for (size_t i = 0; i < d.chunk_nseeds.size(); i++) nseeds[i] = d.chunk_nseeds[i];
}
// intervals (descending sort order): 11 x i64 per interval: score,num_anchors,q0,q1,r0,r1,ref_contig,query_contig,chunk,rev,kept
void orc_debug_intervals(void* dd, int64_t* f) {
  auto& d = ((orc_debug*)dd)->d;
  for (size_t i = 0; i < d.intervals_all.size(); i++) {
    const ChainInterval& c = d.intervals_all[i];
    int64_t* o = f + 11 * i;
    o[0] = (int64_t)c.score; o[1] = (int64_t)c.num_anchors; o[2] = c.q0; o[3] = c.q1; o[4] = c.r0; o[5] = c.r1;
    o[6] = (int64_t)c.ref_contig; o[7] = (int64_t)c.query_contig; o[8] = (int64_t)c.chunk_id; o[9] = c.reverse_chain;
    o[10] = d.interval_kept[i];
  }
}
void orc_debug_ests(void* dd, double* est, uint64_t* weight) {
  auto& d = ((orc_debug*)dd)->d;
  for (size_t i = 0; i < d.ani_ests.size(); i++) { est[i] = d.ani_ests[i].first; weight[i] = d.ani_ests[i].second; }
}

float orc_gbdt_predict(int model, const float* x) { return gbdt_predict(model, x); }

// ---- drivers ----------------------------------------------------------------------------------------
static int emit(const std::vector<PairResult>& v, orc_result** out, uint64_t* n) {
  orc_result* arr = (orc_result*)malloc(sizeof(orc_result) * (v.size() + 1));
  for (size_t i = 0; i < v.size(); i++) to_res(v[i].r, v[i].ref_id, v[i].query_id, &arr[i]);
  *out = arr; *n = v.size();
  return 0;
}
int orc_triangle(void** sk, int n, const orc_cmd* cmd, int threads, orc_result** out, uint64_t* n_out,
                 uint64_t* n_chained, double* t_screen, double* t_chain) {
  std::vector<const Sketch*> v;
  for (int i = 0; i < n; i++) v.push_back((Sketch*)sk[i]);
  return emit(triangle(v, to_cp(cmd), threads, n_chained, t_screen, t_chain), out, n_out);
}
int orc_dist(void** refs, int nr, void** queries, int nq, const orc_cmd* cmd, int use_index, int threads,
             orc_result** out, uint64_t* n_out) {
  std::vector<const Sketch*> r, q;
  for (int i = 0; i < nr; i++) r.push_back((Sketch*)refs[i]);
  for (int i = 0; i < nq; i++) q.push_back((Sketch*)queries[i]);
  return emit(dist(r, q, to_cp(cmd), use_index != 0, threads), out, n_out);
}
int orc_search(void** refs, int nr, void** queries, int nq, const orc_cmd* cmd, int use_index, int threads,
               orc_result** out, uint64_t* n_out) {
  std::vector<const Sketch*> r, q;
  for (int i = 0; i < nr; i++) r.push_back((Sketch*)refs[i]);
  for (int i = 0; i < nq; i++) q.push_back((Sketch*)queries[i]);
  return emit(search(r, q, to_cp(cmd), use_index != 0, threads), out, n_out);
}

}  // extern "C"

// seed ONE contig of any length with either semantics (no MIN_LENGTH_CONTIG rule; used for the reference's
// unit vectors tests/tests.rs:130-157 which call the seeders directly)
extern "C" void* orc_seed_one_contig(const uint8_t* s, uint64_t n, uint64_t c, uint64_t k, uint64_t marker_c, int avx2sem) {
  orc::SketchParams sp; sp.c = c; sp.k = k; sp.marker_c = marker_c;
  orc::Sketch* sk = new orc::Sketch();
  sk->c = c; sk->k = k; sk->marker_c = c;
  sk->contigs.push_back("contig0");
  sk->contig_lengths.push_back((uint32_t)n);
  sk->total_sequence_length = n;
  if (avx2sem) orc::fmh_seeds_avx2sem(s, n, sp, 0, *sk);
  else orc::fmh_seeds_scalar(s, n, sp, 0, *sk);
  return sk;
}

// sketch n_genomes genomes laid out like sk_sketch_batch's input, threads over genomes (the reference's own parallel
// structure, src/file_io.rs:149); handles are written to out[0..n_genomes)
extern "C" int orc_sketch_many(const uint8_t* bases, const uint64_t* contig_off, const uint32_t* genome_of_contig, uint32_t n_contigs,
                               uint32_t n_genomes, uint64_t c, uint64_t k, uint64_t marker_c, int threads, void** out) {
  orc::SketchParams sp; sp.c = c; sp.k = k; sp.marker_c = marker_c;
  std::vector<uint32_t> first(n_genomes + 1, n_contigs);
  for (uint32_t i = n_contigs; i-- > 0;) first[genome_of_contig[i]] = i;
  for (uint32_t g = n_genomes; g-- > 0;) if (first[g] == n_contigs && g + 1 <= n_genomes) first[g] = first[g + 1];
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
  for (long g = 0; g < (long)n_genomes; g++) {
    std::vector<std::pair<const uint8_t*, size_t>> ctgs;
    for (uint32_t i = first[g]; i < first[g + 1]; i++) ctgs.push_back({bases + contig_off[i], (size_t)(contig_off[i + 1] - contig_off[i])});
    char name[32];
    snprintf(name, sizeof(name), "g%06ld", g);
    out[g] = new orc::Sketch(orc::sketch_from_contigs(name, ctgs, nullptr, sp, true));
  }
  return 0;
}
