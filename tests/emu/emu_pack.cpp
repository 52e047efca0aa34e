// Host packer (skani_b200/csrc/host_pack.hpp) against sk::ascii_code (sk_core.cuh), the function pack_kernel applies on
// the device: same units, same N mask, for every byte value, odd lengths and unaligned starts.  Also prints the
// single-thread packing rate.  Development/test harness only.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../skani_b200/csrc/host_pack.hpp"
#include "../../skani_b200/csrc/sk_core.cuh"

int main() {
  std::mt19937_64 rng(7);
  int failures = 0;
  long cases = 0;
  for (uint32_t b = 0; b < 256; b++)
    if (sk_host::code_of(b) != sk::ascii_code(b)) { failures++; fprintf(stderr, "code_of(%u)\n", b); }
  const char* alphabet = "ACGTacgtNnUuRYKMSWBDHV-*.";
  for (int t = 0; t < 4000; t++) {
    const size_t n = t < 300 ? (size_t)t : 1 + rng() % 5000;
    const size_t shift = rng() % 32;
    std::vector<uint8_t> buf(n + shift + 64);
    const int mode = t % 4;
    for (auto& c : buf) {
      if (mode == 0) c = (uint8_t)"ACGT"[rng() & 3];
      else if (mode == 1) c = (uint8_t)alphabet[rng() % 25];
      else if (mode == 2) c = (uint8_t)(rng() & 0xFF);                       // every byte value, incl. 0..3 and >= 0x80
      else c = (rng() % 50 == 0) ? (uint8_t)(rng() & 3) : (uint8_t)"ACGTN"[rng() % 5];
    }
    const uint8_t* s = buf.data() + shift;
    const size_t nu = (n + 31) / 32;
    std::vector<uint64_t> P(nu + 1, 0xDEADBEEFull), Ps(nu + 1, 0), Pr(nu + 1, 0);
    std::vector<uint32_t> M(nu + 1, 0xABCDu), Ms(nu + 1, 0), Mr(nu + 1, 0);
    const bool any_n = sk_host::pack_contig(s, n, P.data(), M.data());
    sk_host::pack_contig_scalar(s, n, Ps.data(), Ms.data());
    { uint32_t want = 0; for (size_t j = 0; j < nu; j++) want |= Ms[j]; if (any_n != (want != 0)) { failures++; fprintf(stderr, "case %d: any-N flag\n", t); } }
    for (size_t i = 0; i < n; i++) {
      const uint32_t v = sk::ascii_code(s[i]);
      Pr[i / 32] |= (uint64_t)(v & 3) << (2 * (i % 32));
      Mr[i / 32] |= (v >> 2) << (i % 32);
    }
    for (size_t j = 0; j < nu; j++)
      if (P[j] != Pr[j] || M[j] != Mr[j] || Ps[j] != Pr[j] || Ms[j] != Mr[j]) { failures++; fprintf(stderr, "case %d unit %zu\n", t, j); break; }
    if (P[nu] != 0xDEADBEEFull || M[nu] != 0xABCDu) { failures++; fprintf(stderr, "case %d wrote past the end\n", t); }
    cases++;
  }
  // device-side SIMD-in-register conversion (sk::pack_word, used by pack_kernel) against ascii_code per byte
  for (int t = 0; t < 2000000; t++) {
    uint32_t x = 0;
    const int mode = t % 3;
    for (int b = 0; b < 4; b++) {
      uint8_t c = mode == 0 ? (uint8_t)"ACGTacgtUu"[rng() % 10] : mode == 1 ? (uint8_t)alphabet[rng() % 25] : (uint8_t)(rng() & 0xFF);
      x |= (uint32_t)c << (8 * b);
    }
    uint32_t c8, n4, rc = 0, rn = 0;
    sk::pack_word(x, c8, n4);
    for (int b = 0; b < 4; b++) { const uint32_t v = sk::ascii_code((x >> (8 * b)) & 0xFF); rc |= (v & 3) << (2 * b); rn |= (v >> 2) << b; }
    if (c8 != rc || n4 != rn) { failures++; if (failures < 10) fprintf(stderr, "pack_word(%08x): %x/%x vs %x/%x\n", x, c8, n4, rc, rn); }
  }
  // rate: 256 MB of ACGT, single thread
  {
    const size_t n = 256u << 20;
    std::vector<uint8_t> g(n);
    for (size_t i = 0; i < n; i++) g[i] = (uint8_t)"ACGT"[(i * 2654435761u >> 13) & 3];
    std::vector<uint64_t> P(n / 32);
    std::vector<uint32_t> M(n / 32);
    sk_host::pack_contig(g.data(), n, P.data(), M.data());
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < 3; r++) sk_host::pack_contig(g.data(), n, P.data(), M.data());
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 3;
    printf("pack rate %.2f GB/s of ASCII per thread\n", n / dt / 1e9);
  }
  printf("%ld cases, %d failures\n", cases, failures);
  return failures ? 1 : 0;
}
