// synth.cpp -- deterministic synthetic genome generator for tests and bench.py.
// Not part of the product library and not part of the oracle: it only manufactures INPUTS.
//
//   PRNG        xoshiro256** seeded through splitmix64(seed ^ stream id)
//   clusters    genomes g belong to cluster g / G; the cluster ancestor is i.i.d. uniform ACGT of length L
//   member m    (= g % G) is the ancestor with i.i.d. substitutions at rate d_m ~ U[dmin, dmax]
//               (substituted base uniform over the other three); members with m % 4 == 1 additionally carry one
//               inversion (reverse complement) of a uniform 100-500 kb segment; members with m % 4 == 2 are cut
//               into 50 contigs at uniform cut points (every contig >= 1000 bp)
//   layout      genome g occupies out[(g - g_begin) * L, +L); contigs are consecutive slices of it
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
};
const char ACGT[4] = {'A', 'C', 'G', 'T'};

void gen_ancestor(uint64_t seed, uint64_t cluster, uint64_t L, uint8_t* out) {
  Rng r(seed ^ (0xA5A5ull << 32) ^ (cluster * 0x9E3779B97F4A7C15ull));
  uint64_t i = 0;
  while (i + 32 <= L) {
    uint64_t x = r.next();
    for (int j = 0; j < 32; j++) { out[i + j] = ACGT[x & 3]; x >>= 2; }
    i += 32;
  }
  if (i < L) { uint64_t x = r.next(); for (; i < L; i++) { out[i] = ACGT[x & 3]; x >>= 2; } }
}
inline int code(uint8_t b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3; }

void cut_points(uint64_t seed, uint64_t g, uint64_t L, int n_ctg, std::vector<uint64_t>& cuts) {
  // n_ctg contigs, each >= min_len: draw cut points on the slack and add spacing
  Rng r(seed ^ (0xC7C7ull << 32) ^ (g * 0xD1B54A32D192ED03ull));
  const uint64_t min_len = 1000;
  uint64_t slack = L - (uint64_t)n_ctg * min_len;
  std::vector<uint64_t> p(n_ctg - 1);
  for (auto& v : p) v = r.below(slack + 1);
  std::sort(p.begin(), p.end());
  cuts.assign(1, 0);
  for (int i = 0; i < n_ctg - 1; i++) cuts.push_back(p[i] + (uint64_t)(i + 1) * min_len);
  cuts.push_back(L);
}
}  // namespace

extern "C" {

// number of contigs of genome g
uint32_t synth_n_contigs(uint64_t g, uint64_t L, uint32_t G) {
  uint32_t m = (uint32_t)(g % G);
  return (m % 4 == 2 && L >= 100000) ? 50u : 1u;
}

// contig layout of genomes [g_begin, g_end): contig_off (relative to the first genome's first base) gets
// total_contigs + 1 entries, genome_of_contig is relative to g_begin.  Returns the number of contigs.
uint64_t synth_layout(uint64_t seed, uint64_t g_begin, uint64_t g_end, uint64_t L, uint32_t G, uint64_t* contig_off,
                      uint32_t* genome_of_contig) {
  uint64_t nc = 0;
  std::vector<uint64_t> cuts;
  for (uint64_t g = g_begin; g < g_end; g++) {
    uint32_t k = synth_n_contigs(g, L, G);
    uint64_t base = (g - g_begin) * L;
    if (k == 1) { if (contig_off) { contig_off[nc] = base; genome_of_contig[nc] = (uint32_t)(g - g_begin); } nc++; }
    else {
      cut_points(seed, g, L, (int)k, cuts);
      for (uint32_t i = 0; i < k; i++) { if (contig_off) { contig_off[nc] = base + cuts[i]; genome_of_contig[nc] = (uint32_t)(g - g_begin); } nc++; }
    }
  }
  if (contig_off) contig_off[nc] = (g_end - g_begin) * L;
  return nc;
}

// member g of its cluster, derived from the cluster ancestor `anc`
static void gen_member(uint64_t seed, uint64_t g, uint64_t L, uint32_t G, double dmin, double dmax, const uint8_t* anc, uint8_t* dst) {
  const uint32_t m = (uint32_t)(g % G);
  memcpy(dst, anc, L);
  Rng r(seed ^ (0x3C3Cull << 32) ^ (g * 0x9E3779B97F4A7C15ull));
  double d = dmin + (dmax - dmin) * r.uniform();
  if (m == 0) d = 0.0 + dmin * 0.0;  // member 0 is the ancestor itself
  if (d > 0) {
    double lg = std::log1p(-d);
    uint64_t i = 0;
    while (true) {
      double u = r.uniform();
      if (u <= 0) u = 1e-300;
      uint64_t skip = (uint64_t)(std::log(u) / lg);  // geometric gap
      i += skip;
      if (i >= L) break;
      int b = code(dst[i]);
      dst[i] = ACGT[(b + 1 + r.below(3)) & 3];
      i++;
    }
  }
  if (m % 4 == 1 && L >= 1000000) {  // one inversion of 100-500 kb
    uint64_t len = 100000 + r.below(400001);
    uint64_t a = r.below(L - len);
    std::reverse(dst + a, dst + a + len);
    for (uint64_t i = a; i < a + len; i++) dst[i] = ACGT[3 - code(dst[i])];
  }
}

// ASCII bases of genomes [g_begin, g_end) into out (size (g_end - g_begin) * L)
void synth_generate(uint64_t seed, uint64_t g_begin, uint64_t g_end, uint64_t L, uint32_t G, double dmin, double dmax,
                    uint8_t* out, int threads) {
  if (g_end <= g_begin) return;
  uint64_t c_begin = g_begin / G, c_end = (g_end - 1) / G + 1;
  if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
  {
    std::vector<uint8_t> anc(L);
#pragma omp for schedule(dynamic, 1)
    for (long c = (long)c_begin; c < (long)c_end; c++) {
      gen_ancestor(seed, (uint64_t)c, L, anc.data());
      for (uint32_t m = 0; m < G; m++) {
        uint64_t g = (uint64_t)c * G + m;
        if (g < g_begin || g >= g_end) continue;
        gen_member(seed, g, L, G, dmin, dmax, anc.data(), out + (g - g_begin) * L);
      }
    }
  }
}

// Arbitrary genome order (bench.py --shuffle-order: file order unrelated to relatedness): slot p of the output holds
// synthetic genome ids[p].  Same bytes per genome as synth_generate; the ancestor of a cluster is generated once per call.
uint64_t synth_layout_ids(uint64_t seed, const uint64_t* ids, uint64_t n, uint64_t L, uint32_t G, uint64_t* contig_off,
                          uint32_t* genome_of_contig) {
  uint64_t nc = 0;
  std::vector<uint64_t> cuts;
  for (uint64_t p = 0; p < n; p++) {
    const uint64_t g = ids[p];
    uint32_t k = synth_n_contigs(g, L, G);
    uint64_t base = p * L;
    if (k == 1) { if (contig_off) { contig_off[nc] = base; genome_of_contig[nc] = (uint32_t)p; } nc++; }
    else {
      cut_points(seed, g, L, (int)k, cuts);
      for (uint32_t i = 0; i < k; i++) { if (contig_off) { contig_off[nc] = base + cuts[i]; genome_of_contig[nc] = (uint32_t)p; } nc++; }
    }
  }
  if (contig_off) contig_off[nc] = n * L;
  return nc;
}

void synth_generate_ids(uint64_t seed, const uint64_t* ids, uint64_t n, uint64_t L, uint32_t G, double dmin, double dmax,
                        uint8_t* out, int threads) {
  if (n == 0) return;
  std::vector<std::pair<uint64_t, uint64_t>> byc(n);   // (genome id, slot), sorted => grouped by cluster
  for (uint64_t p = 0; p < n; p++) byc[p] = {ids[p], p};
  std::sort(byc.begin(), byc.end());
  std::vector<uint64_t> starts;
  for (uint64_t i = 0; i < n; i++) if (i == 0 || byc[i].first / G != byc[i - 1].first / G) starts.push_back(i);
  starts.push_back(n);
  if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
  {
    std::vector<uint8_t> anc(L);
#pragma omp for schedule(dynamic, 1)
    for (long ci = 0; ci < (long)starts.size() - 1; ci++) {
      gen_ancestor(seed, byc[starts[ci]].first / G, L, anc.data());
      for (uint64_t i = starts[ci]; i < starts[ci + 1]; i++) gen_member(seed, byc[i].first, L, G, dmin, dmax, anc.data(), out + byc[i].second * L);
    }
  }
}

}  // extern "C"
