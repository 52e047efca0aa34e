// host_pack.hpp -- ASCII -> 2-bit + N-mask packing of one contig on the HOST, in exactly the unit layout pack_kernel
// (seeding.cu) produces on the device: unit j covers bases 32j .. 32j+31 of the contig, P[j] holds base i at bits 2i..2i+1,
// NM[j] bit i = (byte == 'N').  Byte semantics = sk::ascii_code (sk_core.cuh; reference src/types.rs:40-49 BYTE_TO_SEQ and
// the 'N' test of src/avx2_seeding.rs:115-126).
//
// Why: the end-to-end triangle is PCIe-bound (50 GB of ASCII = 903 ms at the measured 55 GB/s); the
// same genomes are 12.5 GB as 2-bit units (+ an N mask only for contigs that contain 'N').  sk_sketch_batch packs a share
// of every sub-batch on the host while the previous one is in flight (api.cu, the share adapts to the measured packing and
// PCIe rates) and sk_sketch_batch_2bit takes sequences that are already packed.  Checked on the CPU by tests/emu/emu_pack.cpp
// against sk::ascii_code, on the GPU by tests/test_gpu_twobit.py (bit-exact sketches through both entries).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace sk_host {

inline uint32_t code_of(uint32_t b) {   // == sk::ascii_code, restated so this header stays free of CUDA qualifiers
  const uint32_t u = b & 0xDFu;
  uint32_t v = 0;
  if (u == 'C') v = 1;
  else if (u == 'G') v = 2;
  else if (u == 'T' || u == 'U') v = 3;
  if (b < 4) v = b;
  if ((b & 0xC0u) != 0x40u && b >= 4) v = 0;
  if (b == 78) v |= 4;
  return v;
}

inline void pack_unit_scalar(const uint8_t* s, uint32_t nvalid, uint64_t* p, uint32_t* nm) {
  uint64_t packed = 0;
  uint32_t m = 0;
  for (uint32_t i = 0; i < nvalid; i++) {
    const uint32_t v = code_of(s[i]);
    packed |= (uint64_t)(v & 3u) << (2 * i);
    m |= (v >> 2) << i;
  }
  *p = packed; *nm = m;
}

inline void pack_contig_scalar(const uint8_t* s, size_t n, uint64_t* P, uint32_t* NM) {
  const size_t nu = (n + 31) / 32;
  for (size_t j = 0; j < nu; j++) pack_unit_scalar(s + 32 * j, (uint32_t)(n - 32 * j < 32 ? n - 32 * j : 32), P + j, NM + j);
}

#if defined(__x86_64__)
// 32 bases per iteration: five byte compares give the two code bit-planes and the N plane as 32-bit masks; PDEP
// interleaves the planes into the 64-bit unit.  Bytes 0..3 (the table's identity rows) are left to the scalar path.
__attribute__((target("avx2,bmi2"))) inline void pack_contig_avx2(const uint8_t* s, size_t n, uint64_t* P, uint32_t* NM) {
  const size_t full = n / 32;
  const __m256i fold = _mm256_set1_epi8((char)0xDF), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T'),
                cU = _mm256_set1_epi8('U'), cN = _mm256_set1_epi8('N'), low = _mm256_set1_epi8((char)0xFC), zero = _mm256_setzero_si256();
  for (size_t j = 0; j < full; j++) {
    const __m256i x = _mm256_loadu_si256((const __m256i*)(s + 32 * j));
    if (__builtin_expect(_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(x, low), zero)) != 0, 0)) {
      pack_unit_scalar(s + 32 * j, 32, P + j, NM + j);
      continue;
    }
    const __m256i u = _mm256_and_si256(x, fold);
    const __m256i tu = _mm256_or_si256(_mm256_cmpeq_epi8(u, cT), _mm256_cmpeq_epi8(u, cU));
    const uint32_t b0 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpeq_epi8(u, cC), tu));
    const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpeq_epi8(u, cG), tu));
    P[j] = _pdep_u64(b0, 0x5555555555555555ull) | _pdep_u64(b1, 0xAAAAAAAAAAAAAAAAull);
    NM[j] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, cN));
  }
  if (n % 32) pack_unit_scalar(s + 32 * full, (uint32_t)(n % 32), P + full, NM + full);
}
// 64 bases per iteration with AVX-512BW: the byte compares produce the bit planes directly as 64-bit masks.
__attribute__((target("avx512f,avx512bw,bmi2"))) inline void pack_contig_avx512(const uint8_t* s, size_t n, uint64_t* P, uint32_t* NM) {
  const size_t full = n / 64;
  const __m512i fold = _mm512_set1_epi8((char)0xDF), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T'),
                cU = _mm512_set1_epi8('U'), cN = _mm512_set1_epi8('N'), four = _mm512_set1_epi8(4);
  for (size_t j = 0; j < full; j++) {
    const __m512i x = _mm512_loadu_si512((const void*)(s + 64 * j));
    if (__builtin_expect(_mm512_cmplt_epu8_mask(x, four) != 0, 0)) {            // table rows 0..3: scalar path
      pack_unit_scalar(s + 64 * j, 32, P + 2 * j, NM + 2 * j);
      pack_unit_scalar(s + 64 * j + 32, 32, P + 2 * j + 1, NM + 2 * j + 1);
      continue;
    }
    const __m512i u = _mm512_and_si512(x, fold);
    const uint64_t tu = _mm512_cmpeq_epi8_mask(u, cT) | _mm512_cmpeq_epi8_mask(u, cU);
    const uint64_t b0 = _mm512_cmpeq_epi8_mask(u, cC) | tu, b1 = _mm512_cmpeq_epi8_mask(u, cG) | tu;
    const uint64_t nn = _mm512_cmpeq_epi8_mask(x, cN);
    P[2 * j] = _pdep_u64(b0 & 0xFFFFFFFFull, 0x5555555555555555ull) | _pdep_u64(b1 & 0xFFFFFFFFull, 0xAAAAAAAAAAAAAAAAull);
    P[2 * j + 1] = _pdep_u64(b0 >> 32, 0x5555555555555555ull) | _pdep_u64(b1 >> 32, 0xAAAAAAAAAAAAAAAAull);
    NM[2 * j] = (uint32_t)nn; NM[2 * j + 1] = (uint32_t)(nn >> 32);
  }
  if (n % 64) pack_contig_avx2(s + 64 * full, n % 64, P + 2 * full, NM + 2 * full);
}
#endif

// P and NM must hold (n + 31) / 32 entries
inline void pack_contig(const uint8_t* s, size_t n, uint64_t* P, uint32_t* NM) {
#if defined(__x86_64__)
  static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  static const bool fast512 = fast && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
  if (fast512) { pack_contig_avx512(s, n, P, NM); return; }
  if (fast) { pack_contig_avx2(s, n, P, NM); return; }
#endif
  pack_contig_scalar(s, n, P, NM);
}

}  // namespace sk_host
