// CPU check of skani_b200/cli/fast_inflate.hpp against zlib: every compression level and strategy (stored / fixed / dynamic
// blocks, RLE, Huffman only), FASTA-like and random data, long runs (distance 1), sizes around the loop margins, multi-member
// gzip, trailing garbage, truncations and corruptions (must fail or fall back, never crash or return wrong bytes), the gzip
// fixtures given on the command line.  Prints "N cases, F failures" and the decode rates of both decoders.
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <functional>
#include <random>
#include <thread>
#include <string>
#include <vector>

#include "../../skani_b200/cli/fast_inflate.hpp"

static std::string gz(const std::string& in, int level, int strategy, int wbits = 15 + 16) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  deflateInit2(&zs, level, Z_DEFLATED, wbits, 8, strategy);
  std::string out(deflateBound(&zs, in.size()) + 64, 0);
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
  deflate(&zs, Z_FINISH);
  out.resize(zs.total_out);
  deflateEnd(&zs);
  return out;
}
static bool zlib_gunzip(const std::string& in, std::string& out) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  inflateInit2(&zs, 15 + 16);
  out.assign(std::max<size_t>(in.size() * 8, 1 << 16), 0);
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  size_t produced = 0;
  for (;;) {
    if (produced == out.size()) out.resize(out.size() * 2);
    zs.next_out = (Bytef*)&out[produced]; zs.avail_out = (uInt)(out.size() - produced);
    const int rc = inflate(&zs, Z_NO_FLUSH);
    produced = out.size() - zs.avail_out;
    if (rc == Z_STREAM_END) {
      if (zs.avail_in >= 2 && zs.next_in[0] == 0x1f && zs.next_in[1] == 0x8b) { inflateReset(&zs); continue; }
      break;
    }
    if (rc != Z_OK) { inflateEnd(&zs); return false; }
  }
  inflateEnd(&zs);
  out.resize(produced);
  return true;
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static bool same(const sk_inflate::TextBuf& a, const std::string& b) { return a.size() == b.size() && (b.empty() || memcmp(a.data(), b.data(), b.size()) == 0); }

int main(int argc, char** argv) {
  std::mt19937_64 rng(20260924);
  int cases = 0, failures = 0;
  auto check = [&](const std::string& plain, const std::string& comp, const char* what) {
    sk_inflate::TextBuf got;
    cases++;
    if (!sk_inflate::gunzip((const uint8_t*)comp.data(), comp.size(), got, (sk_inflate::crc_fn)crc32) || !same(got, plain)) {
      failures++;
      fprintf(stderr, "FAIL %s: plain %zu comp %zu got %zu\n", what, plain.size(), comp.size(), got.size());
    }
  };
  auto fasta = [&](size_t n) {
    std::string s;
    while (s.size() < n) {
      s += ">contig_" + std::to_string(rng() % 100000) + " len=" + std::to_string(rng() % 9999) + "\n";
      const size_t L = 60 + rng() % 21, m = 200 + rng() % 20000;
      std::string seq(m, 'A');
      for (auto& c : seq) c = "ACGT"[rng() & 3];
      if (rng() % 3 == 0 && m > 3000) memmove(&seq[m / 2], &seq[100], std::min<size_t>(1000 + rng() % 1500, m - m / 2));   // repeats -> long matches
      if (rng() % 5 == 0) for (size_t i = 0; i < std::min<size_t>(m, 500); i++) seq[i] = 'N';      // runs -> distance 1
      for (size_t i = 0; i < m; i += L) { s.append(seq, i, L); s += '\n'; }
    }
    s.resize(n);
    return s;
  };
  const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
  for (int t = 0; t < 600; t++) {
    size_t n;
    if (t < 80) n = (size_t)t;                                  // tiny, incl. empty
    else if (t < 160) n = 250 + (size_t)(t - 80) * 7;           // around the 258 + slack margins
    else n = 1000 + rng() % 400000;
    std::string plain;
    switch (t % 5) {
      case 0: plain = fasta(n); break;
      case 1: plain.resize(n); for (auto& c : plain) c = (char)(rng() & 0xFF); break;              // incompressible: stored blocks at level 0..1
      case 2: plain.assign(n, 'x'); break;                                                             // one long run
      case 3: plain = fasta(n); for (size_t i = 0; i < plain.size(); i += 97) plain[i] = (char)(rng() & 0xFF); break;
      default: { plain.resize(n); for (size_t i = 0; i < n; i++) plain[i] = (char)("ab"[(i / (1 + i % 7)) & 1] + (rng() % 50 == 0)); }
    }
    const int level = (int)(t % 10), strat = strategies[(t / 10) % 5];
    check(plain, gz(plain, level, strat), "single member");
    if (t % 7 == 0) {                                           // small windows (distances stay short), still valid gzip
      check(plain, gz(plain, 6, Z_DEFAULT_STRATEGY, 9 + 16), "window 512");
    }
    if (t % 11 == 0) {                                          // several members, then bytes that are not a member
      std::string a = fasta(1 + rng() % 5000), b = fasta(rng() % 300), c;
      std::string comp = gz(plain, level, strat) + gz(a, 9, Z_DEFAULT_STRATEGY) + gz(b, 1, Z_DEFAULT_STRATEGY) + gz(c, 6, Z_DEFAULT_STRATEGY);
      check(plain + a + b, comp, "multi member");
      check(plain + a + b, comp + std::string("\0\0\0garbage", 10), "trailing bytes");
    }
    if (t % 13 == 0 && n > 50) {                                // truncated / corrupted: false or the right bytes, never garbage
      std::string comp = gz(plain, 6, Z_DEFAULT_STRATEGY);
      sk_inflate::TextBuf got;
      for (int k = 0; k < 6; k++) {
        std::string bad = comp;
        if (k < 3) bad.resize(bad.size() - 1 - rng() % std::min<size_t>(bad.size() - 1, 40));
        else bad[10 + rng() % (bad.size() - 10)] ^= (char)(1 << (rng() % 8));
        cases++;
        const bool ok = sk_inflate::gunzip((const uint8_t*)bad.data(), bad.size(), got, (sk_inflate::crc_fn)crc32);
        std::string ref;
        const bool zok = zlib_gunzip(bad, ref);
        if (ok && (!zok || !same(got, ref))) { failures++; fprintf(stderr, "FAIL corrupt case accepted: t=%d k=%d\n", t, k); }
      }
    }
  }
  for (int a = 1; a < argc; a++) {                              // fixtures: must equal zlib's output
    FILE* f = fopen(argv[a], "rb");
    if (!f) { failures++; continue; }
    std::string comp;
    char buf[65536];
    size_t r;
    while ((r = fread(buf, 1, sizeof(buf), f)) > 0) comp.append(buf, r);
    fclose(f);
    std::string ref;
    if (!zlib_gunzip(comp, ref)) { fprintf(stderr, "zlib failed on %s\n", argv[a]); failures++; continue; }
    check(ref, comp, argv[a]);
  }
  // ---- block-parallel decoding of one member (two-pass scheme): must equal the plain text, or decline
  auto prun = [](size_t n, const std::function<void(size_t)>& f) {
    std::vector<std::thread> th;
    std::atomic<size_t> next{0};
    auto w = [&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); };
    for (int t = 1; t < 6; t++) th.emplace_back(w);
    w();
    for (auto& t : th) t.join();
  };
  {
    int declined = 0, decoded = 0;
    for (int t = 0; t < 36; t++) {
      const size_t n = (t < 30 ? (1u << 20) : (10u << 20)) + rng() % (1u << 20);
      std::string plain = fasta(n);
      if (t % 10 == 9) for (auto& c : plain) c = (char)(rng() & 0xFF);              // binary: no text block will be found
      const int level = 1 + t % 9;
      std::string comp = gz(plain, level, (t % 4 == 3) ? Z_FILTERED : Z_DEFAULT_STRATEGY);
      if (t % 10 == 8) comp += gz(fasta(5000), 6, Z_DEFAULT_STRATEGY);                  // two members: must decline
      sk_inflate::TextBuf got;
      cases++;
      const size_t min_chunk = t < 30 ? (32u << 10) : (512u << 10);
      const bool ok = sk_inflate::gunzip_parallel((const uint8_t*)comp.data(), comp.size(), got, 6, (sk_inflate::crc_fn)crc32,
                                                  (sk_inflate::crc_combine_fn)crc32_combine, prun, min_chunk);
      if (ok) { decoded++; if (!same(got, plain) || t % 10 == 8) { failures++; fprintf(stderr, "FAIL parallel gunzip t=%d\n", t); } }
      else declined++;
      if (!ok && t % 10 < 8) { failures++; fprintf(stderr, "FAIL parallel gunzip declined plain text t=%d level %d\n", t, level); }
      if (t % 10 == 5) {                                                                  // corruption in the middle: decline (CRC) or decode like zlib
        std::string bad = comp;
        bad[bad.size() / 2] ^= 0x10;
        sk_inflate::TextBuf g2;
        std::string ref;
        cases++;
        const bool ok2 = sk_inflate::gunzip_parallel((const uint8_t*)bad.data(), bad.size(), g2, 6, (sk_inflate::crc_fn)crc32,
                                                     (sk_inflate::crc_combine_fn)crc32_combine, prun, min_chunk);
        if (ok2 && (!zlib_gunzip(bad, ref) || !same(g2, ref))) { failures++; fprintf(stderr, "FAIL parallel gunzip accepted a corrupt stream t=%d\n", t); }
      }
    }
    printf("parallel gunzip: %d decoded, %d declined\n", decoded, declined);
  }
  // rates on 64 MB of FASTA at level 6 (best of 3 each)
  {
    const std::string plain = fasta(64u << 20), comp = gz(plain, 6, Z_DEFAULT_STRATEGY);
    double best_f = 1e9, best_n = 1e9, best_z = 1e9;
    for (int r = 0; r < 2; r++) {
      sk_inflate::TextBuf o1, o3;
      std::string o2;
      double t0 = now(); const bool a = sk_inflate::gunzip((const uint8_t*)comp.data(), comp.size(), o1, (sk_inflate::crc_fn)crc32); double t1 = now();
      const bool b = zlib_gunzip(comp, o2); double t2 = now();
      sk_inflate::gunzip((const uint8_t*)comp.data(), comp.size(), o3, nullptr); double t3 = now();
      if (!a || !b || !same(o1, o2) || !same(o3, o2)) failures++;
      best_f = std::min(best_f, t1 - t0); best_z = std::min(best_z, t2 - t1); best_n = std::min(best_n, t3 - t2);
    }
    printf("rates (64 MB FASTA, ratio %.2f): fast_inflate %.0f MB/s (%.0f without CRC), zlib %.0f MB/s\n", (double)plain.size() / comp.size(),
           plain.size() / 1e6 / best_f, plain.size() / 1e6 / best_n, plain.size() / 1e6 / best_z);
    double best_p = 1e9;
    for (int r = 0; r < 2; r++) {
      sk_inflate::TextBuf o4;
      double t0 = now();
      const bool okp = sk_inflate::gunzip_parallel((const uint8_t*)comp.data(), comp.size(), o4, 6, (sk_inflate::crc_fn)crc32,
                                                   (sk_inflate::crc_combine_fn)crc32_combine, prun, 1u << 20);
      best_p = std::min(best_p, now() - t0);
      if (!okp || !same(o4, plain)) failures++;
    }
    printf("parallel gunzip (6 threads, one member): %.0f MB/s\n", plain.size() / 1e6 / best_p);
  }
  printf("%d cases, %d failures\n", cases, failures);
  return failures ? 1 : 0;
}
