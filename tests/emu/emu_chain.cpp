// Host emulation of the chaining path's per-item logic (skani_b200/csrc/chain_core.cuh: the __host__ __device__ functions
// the CUDA kernels call) checked against the CPU oracle's parity taps.  Development/test harness only: it validates the
// closed forms without a GPU; it is not a product path.
//   1. chunk assignment: FirstOp / MinOp segmented scans + chunk_need + chunk_local_of, driven sequentially in the order
//      chunk_kernel (chain.cu) applies them, against the oracle's chunk boundaries (sequential loop of src/chain.rs:738-836),
//      including anchor-free stretches > 20 kb ("catch-up" singleton chunks) and multi-contig queries;
//   2. interval order (IntervalKey / interval_before) and the greedy non-overlap decisions (overlap_contrib /
//      overlap_accept) against the oracle's sorted interval list and kept flags (src/chain.rs:1008-1099);
//   3. wyrand_at / lemire_below (random access) against a sequential WyRand + Lemire;
//   4. gbdt_eval on the flattened tables against the oracle's tree walk, bit-exact f32.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <vector>

#include "../../skani_b200/csrc/chain_core.cuh"
#include "../../oracle/skani_oracle.hpp"

namespace tables {
#include "../../skani_b200/csrc/gbdt_tables.inc"
}

static int failures = 0;
static long catchup_anchors = 0;   // anchors whose chunk is held back below `need` by the catch-up rule
#define CHECK(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static std::vector<uint8_t> random_seq(std::mt19937_64& rng, size_t n) {
  std::vector<uint8_t> s(n);
  for (auto& b : s) b = "ACGT"[rng() & 3];
  return s;
}
static std::vector<uint8_t> mutate(std::mt19937_64& rng, const std::vector<uint8_t>& a, double rate) {
  std::vector<uint8_t> s = a;
  std::uniform_real_distribution<double> u(0, 1);
  for (auto& b : s) if (u(rng) < rate) b = "ACGT"[rng() & 3];
  return s;
}
static orc::Sketch sketch_of(const char* name, const std::vector<std::vector<uint8_t>>& ctgs, const orc::SketchParams& sp) {
  std::vector<std::pair<const uint8_t*, size_t>> v;
  for (auto& c : ctgs) v.push_back({c.data(), c.size()});
  return orc::sketch_from_contigs(name, v, nullptr, sp, true);
}

// ---- 1. chunk ids from the sorted anchor list, with the device functions in chunk_kernel's order of application
static std::vector<uint32_t> emu_chunk_first(const std::vector<orc::Anchor>& an) {
  std::vector<uint32_t> first;
  sk::FirstState carryF; carryF.valid = 0; carryF.ctg = 0; carryF.p0 = 0; carryF.a0 = 0;
  sk::MinState carryM; carryM.valid = 0; carryM.ctg = 0; carryM.v = 0;
  uint32_t prev_ctg = 0xFFFFFFFFu, prev_cl = 0;
  for (size_t a = 0; a < an.size();) {
    size_t e = a + 1;      // one hit record = the anchors of one query seed position
    while (e < an.size() && an[e].query_contig == an[a].query_contig && an[e].query_pos == an[a].query_pos) e++;
    const uint32_t nh = (uint32_t)(e - a);
    sk::FirstState f; f.valid = 1; f.ctg = an[a].query_contig; f.p0 = an[a].query_pos; f.a0 = (uint32_t)a;
    carryF = sk::FirstOp()(carryF, f);                       // inclusive scan
    const uint32_t need = sk::chunk_need(an[a].query_pos, carryF.p0);
    const uint32_t al = (uint32_t)a - carryF.a0;
    const bool has_prev = carryM.valid && carryM.ctg == an[a].query_contig;   // exclusive scan state
    for (uint32_t t = 0; t < nh; t++) {
      const uint32_t cl = sk::chunk_local_of((uint64_t)al + t, has_prev, carryM.v, need);
      if (cl != need) catchup_anchors++;
      if (a + t == 0 || prev_ctg != an[a].query_contig || prev_cl != cl) first.push_back((uint32_t)(a + t));
      prev_ctg = an[a].query_contig; prev_cl = cl;
    }
    sk::MinState m; m.valid = 1; m.ctg = an[a].query_contig; m.v = (int64_t)need - (int64_t)al - (int64_t)(nh - 1);
    carryM = sk::MinOp()(carryM, m);
    a = e;
  }
  first.push_back((uint32_t)an.size());
  return first;
}

// ---- 2. interval order + greedy filter
static void check_intervals(const orc::ChainDebug& d, const char* what) {
  std::vector<sk::IntervalKey> keys;
  for (auto& iv : d.intervals_all)
    keys.push_back(sk::make_interval((int32_t)iv.score, (uint32_t)iv.num_anchors, iv.q0, iv.q1, iv.r0, iv.r1, (uint32_t)iv.ref_contig,
                                     (uint32_t)iv.query_contig, (uint32_t)iv.chunk_id, iv.reverse_chain ? 1u : 0u));
  for (size_t i = 0; i + 1 < keys.size(); i++) {
    CHECK(!sk::interval_before(keys[i + 1], keys[i]), "%s: interval %zu sorts after %zu", what, i, i + 1);
    CHECK((double)sk::iv_score(keys[i]) == d.intervals_all[i].score, "%s: score not an integer", what);
  }
  // the device sorts with interval_before: a std::sort of shuffled keys must give the oracle's order back
  std::vector<size_t> perm(keys.size());
  for (size_t i = 0; i < perm.size(); i++) perm[i] = i;
  std::mt19937_64 rng(11);
  std::shuffle(perm.begin(), perm.end(), rng);
  std::sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return sk::interval_before(keys[x], keys[y]); });
  for (size_t i = 0; i < perm.size(); i++)
    CHECK(memcmp(&keys[perm[i]], &keys[i], sizeof(sk::IntervalKey)) == 0, "%s: sorted position %zu differs", what, i);
  std::vector<size_t> acc;
  for (size_t i = 0; i < keys.size(); i++) {
    uint32_t sr = 0, hr = 0, sq = 0, hq = 0;
    for (size_t j : acc) sk::overlap_contrib(keys[i], keys[j], &sr, &hr, &sq, &hq);
    const bool keep = sk::overlap_accept(keys[i], sr, hr, sq, hq);
    CHECK(keep == (d.interval_kept[i] != 0), "%s: greedy decision of interval %zu", what, i);
    if (keep) acc.push_back(i);
  }
}

struct SeqWyRand {   // fastrand 1.9.0, sequential form
  uint64_t state;
  uint64_t next() {
    state += 0xA0761D6478BD642Full;
    __uint128_t t = (__uint128_t)state * (__uint128_t)(state ^ 0xE7037ED1A0B428DBull);
    return (uint64_t)t ^ (uint64_t)(t >> 64);
  }
};

int main() {
  std::mt19937_64 rng(20260924);
  int pairs_checked = 0, chunks_checked = 0, intervals_checked = 0;
  for (uint64_t c : {125ull, 30ull}) {
    orc::SketchParams sp; sp.c = c; sp.k = 15; sp.marker_c = c == 30 ? 200 : 1000;
    orc::CommandParams cp;
    const size_t L = 400000;
    std::vector<uint8_t> base = random_seq(rng, L);
    std::vector<std::pair<std::string, orc::Sketch>> sk;
    sk.push_back({"plain", sketch_of("a_plain", {mutate(rng, base, 0.01)}, sp)});
    sk.push_back({"divergent", sketch_of("b_div", {mutate(rng, base, 0.06)}, sp)});
    {  // 90 / 65 / 24 kb replaced by unrelated sequence: anchor-free stretches in BOTH roles -> the catch-up rule (src/chain.rs:744-793)
      std::vector<uint8_t> g = mutate(rng, base, 0.02);
      const size_t from[3] = {60000, 220000, 330000}, to[3] = {150000, 285000, 354000};
      for (int i = 0; i < 3; i++) {
        std::vector<uint8_t> junk = random_seq(rng, to[i] - from[i]);
        std::copy(junk.begin(), junk.end(), g.begin() + from[i]);
      }
      sk.push_back({"gappy", sketch_of("c_gappy", {g}, sp)});
    }
    {  // 12 contigs of uneven length, two of them reverse-complemented, one with a duplicated 40 kb block (repeats)
      std::vector<uint8_t> g = mutate(rng, base, 0.03);
      std::vector<std::vector<uint8_t>> ctgs;
      size_t p = 0;
      for (int i = 0; i < 12 && p < L; i++) {
        size_t len = std::min<size_t>(L - p, 5000 + (rng() % 70000));
        std::vector<uint8_t> ctg(g.begin() + p, g.begin() + p + len);
        if (i % 5 == 1) {
          std::reverse(ctg.begin(), ctg.end());
          for (auto& b : ctg) b = b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : 'A';
        }
        if (i == 3) ctg.insert(ctg.end(), g.begin() + 20000, g.begin() + 60000);
        ctgs.push_back(std::move(ctg));
        p += len;
      }
      sk.push_back({"contigs", sketch_of("d_contigs", ctgs, sp)});
    }
    for (size_t i = 0; i < sk.size(); i++)
      for (size_t j = 0; j < sk.size(); j++) {
        if (i == j) continue;
        orc::ChainDebug d;
        orc::MapParams mp = orc::map_params_from_sketch(sk[i].second, cp, orc::get_model_id(c, true));
        orc::chain_seeds(sk[i].second, sk[j].second, mp, &d);
        std::string what = "c=" + std::to_string(c) + " " + sk[i].first + " x " + sk[j].first;
        CHECK(d.anchors.size() > 500, "%s: only %zu anchors", what.c_str(), d.anchors.size());
        std::vector<uint32_t> first = emu_chunk_first(d.anchors);
        CHECK(first == d.chunk_first, "%s: chunk boundaries differ (%zu vs %zu chunks)", what.c_str(), first.size(), d.chunk_first.size());
        CHECK(!d.intervals_all.empty(), "%s: no intervals", what.c_str());
        check_intervals(d, what.c_str());
        pairs_checked++; chunks_checked += (int)d.chunk_first.size() - 1; intervals_checked += (int)d.intervals_all.size();
      }
  }
  // ---- 3. random-access WyRand == sequential stream; Lemire without the rejection loop
  {
    SeqWyRand s{7};
    for (uint64_t n = 0; n < 200000; n++) {
      uint64_t r = s.next();
      CHECK(r == sk::wyrand_at(7, n), "wyrand draw %llu", (unsigned long long)n);
      uint64_t bound = 1 + (rng() % 100000);
      bool rej = false;
      uint64_t v = sk::lemire_below(r, bound, &rej);
      __uint128_t m = (__uint128_t)r * bound;
      CHECK(v == (uint64_t)(m >> 64) && v < bound, "lemire value");
      bool would = (uint64_t)m < bound && (uint64_t)m < (0 - bound) % bound;
      CHECK(rej == would, "lemire rejection flag");
    }
  }
  // ---- 4. GBDT: flattened complete depth-3 trees == the oracle's predict, both models
  {
    auto b2f = [](uint32_t b) { float f; memcpy(&f, &b, 4); return f; };
    std::vector<float> thr[2], leaf[2];
    for (int i = 0; i < 1365; i++) { thr[0].push_back(b2f(tables::SK_GBDT_C125_THR[i])); thr[1].push_back(b2f(tables::SK_GBDT_C200_THR[i])); }
    for (int i = 0; i < 1560; i++) { leaf[0].push_back(b2f(tables::SK_GBDT_C125_LEAF[i])); leaf[1].push_back(b2f(tables::SK_GBDT_C200_LEAF[i])); }
    const unsigned char* feat[2] = {tables::SK_GBDT_C125_FEAT, tables::SK_GBDT_C200_FEAT};
    const float shrink[2] = {b2f(SK_GBDT_C125_SHRINK_BITS), b2f(SK_GBDT_C200_SHRINK_BITS)};
    const float bias[2] = {b2f(SK_GBDT_C125_BIAS_BITS), b2f(SK_GBDT_C200_BIAS_BITS)};
    const int ntrees[2] = {SK_GBDT_C125_NTREES, SK_GBDT_C200_NTREES};
    std::uniform_real_distribution<float> ani(88.f, 100.f), sd(0.f, 6.f), ql(500.f, 3e6f), cl(200.f, 20000.f);
    for (int it = 0; it < 200000; it++) {
      float x[5] = {ani(rng), sd(rng), ql(rng), ql(rng), cl(rng)};
      if (it % 7 == 0) x[2] = std::floor(x[2]);
      for (int m = 0; m < 2; m++) {
        float a = sk::gbdt_eval(feat[m], thr[m].data(), leaf[m].data(), ntrees[m], shrink[m], bias[m], x);
        float b = orc::gbdt_predict(m, x);
        CHECK(memcmp(&a, &b, 4) == 0, "gbdt model %d: %.9g vs %.9g", m, (double)a, (double)b);
      }
    }
  }
  CHECK(catchup_anchors > 0, "no input exercised the catch-up rule");
  printf("%d pairs, %d chunks, %d intervals, %ld catch-up anchors, %d failures\n", pairs_checked, chunks_checked, intervals_checked,
         catchup_anchors, failures);
  return failures ? 1 : 0;
}
