// Host emulation of the seeding kernels' per-unit arithmetic (skani_b200/csrc/sk_core.cuh) checked against the
// CPU oracle.  Development/test harness only: it validates bit tricks without a GPU; it is not a product path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <tuple>
#include <vector>

#include "../../skani_b200/csrc/sk_core.cuh"
#include "../../oracle/skani_oracle.hpp"

using Rec = std::tuple<uint32_t, uint32_t, uint32_t>;  // kmer, pos, canon

static void emu_seed(const std::vector<uint8_t>& s, uint32_t c, uint32_t k, uint32_t marker_c, std::vector<Rec>& recs,
                     std::vector<uint64_t>& markers, bool scalar = false) {
  uint32_t n = (uint32_t)s.size();
  uint32_t nu = (n + 31) / 32;
  std::vector<uint64_t> P(nu, 0);
  std::vector<uint32_t> NM(nu, 0);
  for (uint32_t i = 0; i < n; i += 4) {   // through pack_word, as pack_kernel does (4 bytes at a time, zero padded)
    uint32_t x = 0;
    for (uint32_t b = 0; b < 4 && i + b < n; b++) x |= (uint32_t)s[i + b] << (8 * b);
    uint32_t c8, n4;
    sk::pack_word(x, c8, n4, scalar);
    for (uint32_t b = 0; b < 4 && i + b < n; b++) {
      P[(i + b) / 32] |= (uint64_t)((c8 >> (2 * b)) & 3) << (2 * ((i + b) % 32));
      NM[(i + b) / 32] |= (uint32_t)((n4 >> b) & 1) << ((i + b) % 32);
    }
  }
  uint64_t seed_mask = ~0ull >> (64 - 2 * k);
  uint64_t thr = ~0ull / c, thr_m = ~0ull / marker_c;
  for (uint32_t ul = 0; ul < nu; ul++) {
    uint64_t lo = ul ? P[ul - 1] : 0, hi = P[ul];
    uint32_t nlo = ul ? NM[ul - 1] : 0, nhi = NM[ul];
    uint32_t pass;
    if (scalar) pass = sk::unit_pass_mask_fast(lo, hi, nlo, nhi, n, ul, (uint32_t)seed_mask, thr, k);
    else {
      pass = sk::unit_pass_mask(lo, hi, nlo, nhi, n, ul, seed_mask, thr);
      if (pass != sk::unit_pass_mask_fast(lo, hi, nlo, nhi, n, ul, (uint32_t)seed_mask, thr)) { fprintf(stderr, "fast/slow pass mask mismatch\n"); exit(2); }
    }
    sk::WindowCtx w = sk::make_window_ctx(lo, hi);
    for (uint32_t j = 0; j < 32; j++) {
      if (!((pass >> j) & 1)) continue;
      bool canon;
      uint32_t seed = sk::window_seed(w, j, seed_mask, &canon);
      recs.push_back({seed, 32 * ul + j, canon ? 1u : 0u});
      if (sk::mm_hash64(seed) < thr_m) markers.push_back(sk::window_marker(w, j));
    }
  }
  std::sort(markers.begin(), markers.end());
  markers.erase(std::unique(markers.begin(), markers.end()), markers.end());
  std::sort(recs.begin(), recs.end());
}

static void oracle_seed(const std::vector<uint8_t>& s, uint32_t c, uint32_t k, uint32_t marker_c, std::vector<Rec>& recs,
                        std::vector<uint64_t>& markers, bool scalar = false) {
  orc::SketchParams sp; sp.c = c; sp.k = k; sp.marker_c = marker_c;
  orc::Sketch sk;
  if (scalar) orc::fmh_seeds_scalar(s.data(), s.size(), sp, 0, sk);
  else orc::fmh_seeds_avx2sem(s.data(), s.size(), sp, 0, sk);
  const orc::KmerSeeds& m = sk.kmer_seeds_k;
  orc::SeedPosition tmp;
  for (size_t i = 0; i < m.capacity(); i++) {
    if (!m.slot_used(i)) continue;
    const orc::SeedPosition* p;
    size_t cnt = sk.get_seed_positions(m.slot_key(i), &p, &tmp);
    for (size_t a = 0; a < cnt; a++) recs.push_back({m.slot_key(i), p[a].pos, p[a].contig_index_canonical & 1});
  }
  for (size_t i = 0; i < sk.marker_seeds.capacity(); i++)
    if (sk.marker_seeds.slot_used(i)) markers.push_back(sk.marker_seeds.slot_key(i));
  std::sort(markers.begin(), markers.end());
  std::sort(recs.begin(), recs.end());
}

int main() {
  std::mt19937_64 rng(20260924);
  const char alpha[] = "ACGT";
  int fails = 0, cases = 0;
  for (int t = 0; t < 400; t++) {
    uint32_t n;
    if (t < 60) n = 20 + t;            // around the 42-base minimum, every residue of (n-20) mod 4
    else if (t < 120) n = 480 + (t - 60);
    else n = 500 + (uint32_t)(rng() % 20000);
    std::vector<uint8_t> s(n);
    for (auto& ch : s) ch = alpha[rng() % 4];
    int flavour = t % 8;
    if (flavour == 1) for (int r = 0; r < 5; r++) s[rng() % n] = 'N';
    if (flavour == 2) { uint32_t q = (n > 20 ? (n - 20) / 4 : 0); for (uint32_t l = 0; l < 4 && q; l++) for (int d = -2; d < 24; d++) { int64_t p = (int64_t)l * q + d + (int64_t)(rng() % 3); if (p >= 0 && p < n && rng() % 3 == 0) s[p] = 'N'; } }
    if (flavour == 3) for (uint32_t i = 0; i < n; i++) if (rng() % 7 == 0) s[i] = (uint8_t)tolower(s[i]);
    if (flavour == 4) for (int r = 0; r < 30; r++) s[rng() % n] = "RYKMSWnuUBDHV-*\x01\x02\x03\x00"[rng() % 19];
    if (flavour == 5) { uint32_t a = rng() % n, b = std::min<uint32_t>(n, a + 1 + rng() % 200); for (uint32_t i = a; i < b; i++) s[i] = 'N'; }
    if (flavour == 6) { for (uint32_t i = 0; i < n; i++) s[i] = "AC"[(i / 3) % 2]; }  // low complexity / repeats
    if (flavour == 7) for (uint32_t i = n > 30 ? n - 30 : 0; i < n; i++) if (rng() % 4 == 0) s[i] = 'N';
    uint32_t cs[] = {125, 30, 10, 200, 1};
    uint32_t c = cs[t % 5], k = (t % 11 == 0) ? 13 : ((t % 13 == 0) ? 16 : 15), mc = std::max(c, (t % 3 == 0) ? 200u : 1000u);
    std::vector<Rec> r1, r2;
    std::vector<uint64_t> m1, m2;
    emu_seed(s, c, k, mc, r1, m1);
    oracle_seed(s, c, k, mc, r2, m2);
    cases++;
    if (r1 != r2 || m1 != m2) {
      fails++;
      fprintf(stderr, "MISMATCH case %d n=%u c=%u k=%u flavour=%d: recs %zu vs %zu, markers %zu vs %zu\n", t, n, c, k, flavour,
              r1.size(), r2.size(), m1.size(), m2.size());
    }
    // scalar fmh_seeds semantics (src/seeding.rs:225-323): one lane, no tail drop, 'N' and 'n' suppress k windows
    std::vector<Rec> s1, s2;
    std::vector<uint64_t> n1, n2;
    emu_seed(s, c, k, mc, s1, n1, true);
    oracle_seed(s, c, k, mc, s2, n2, true);
    if (s1 != s2 || n1 != n2) {
      fails++;
      fprintf(stderr, "SCALAR MISMATCH case %d n=%u c=%u k=%u flavour=%d: recs %zu vs %zu, markers %zu vs %zu\n", t, n, c, k, flavour,
              s1.size(), s2.size(), n1.size(), n2.size());
    }
  }
  // table check: ascii_code vs BYTE_TO_SEQ for all 256 bytes
  for (int b = 0; b < 256; b++) {
    uint32_t v = sk::ascii_code(b);
    if ((v & 3) != orc::BYTE_TO_SEQ[b] || ((v >> 2) != (b == 78))) { fails++; fprintf(stderr, "ascii_code mismatch at %d\n", b); }
  }
  printf("emu_seed: %d cases, %d failures\n", cases, fails);
  return fails ? 1 : 0;
}
