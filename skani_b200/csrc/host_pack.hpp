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
#include <stdio.h>
#include <stdlib.h>
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
// 64 bases per iteration with AVX-512 VBMI: ONE byte permute looks the 2-bit code of every base up in a 64-entry table indexed
// by the low six bits of the byte (bytes outside 0x40..0x7F are zeroed by the mask: only letters carry a code), two
// multiply-adds gather 4 codes into a byte and a down-convert leaves the 128 packed bits -- 11 micro-ops per 64 bases where the
// compare-based variant above needs ~28 cycles (measured 5 GB/s per thread; this one is bound by memory instead).
__attribute__((target("avx512f,avx512bw,avx512vbmi,sse4.1"))) inline bool pack_contig_avx512vbmi(const uint8_t* s, size_t n, uint64_t* P, uint32_t* NM) {   // returns: any 'N' seen
  uint64_t any_n = 0;
  alignas(64) static const uint8_t lut[64] = {
      0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,     // @ A B C D E F G ... T U ...
      0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};    // ` a b c d e f g ... t u ...
  const __m512i LUT = _mm512_load_si512((const void*)lut);
  const __m512i hi2 = _mm512_set1_epi8((char)0xC0), letter = _mm512_set1_epi8(0x40), cN = _mm512_set1_epi8('N'), four = _mm512_set1_epi8(4);
  const __m512i w8 = _mm512_set1_epi16(0x0401);       // bytes (1, 4): code0 + 4 * code1
  const __m512i w16 = _mm512_set1_epi32(0x00100001);  // words (1, 16): + 16 * (code2 + 4 * code3)
  const size_t full = n / 64;
  for (size_t j = 0; j < full; j++) {
    // the sequence streams through once: prefetch a page ahead (the hardware streamer stops at 4 KB page boundaries) and
    // write the units with non-temporal stores (no read-for-ownership of lines nobody on the CPU reads again: the consumer
    // is the GPU's DMA engine) -- 7.0 instead of 5.6 GB/s per thread from DRAM
    _mm_prefetch((const char*)(s + 64 * j + 4096), _MM_HINT_T0);
    const __m512i x = _mm512_loadu_si512((const void*)(s + 64 * j));
    if (__builtin_expect(_mm512_cmplt_epu8_mask(x, four) != 0, 0)) {            // table rows 0..3 (identity): scalar path
      pack_unit_scalar(s + 64 * j, 32, P + 2 * j, NM + 2 * j);
      pack_unit_scalar(s + 64 * j + 32, 32, P + 2 * j + 1, NM + 2 * j + 1);
      any_n |= NM[2 * j] | NM[2 * j + 1];
      continue;
    }
    const __mmask64 is_letter = _mm512_cmpeq_epi8_mask(_mm512_and_si512(x, hi2), letter);
    const __m512i code = _mm512_maskz_permutexvar_epi8(is_letter, x, LUT);     // vpermb uses index bits 5..0 only
    const __m512i q = _mm512_madd_epi16(_mm512_maddubs_epi16(code, w8), w16);   // one byte (4 bases) in the low byte of every dword
    const __m128i r = _mm512_cvtepi32_epi8(q);
    const uint64_t nn = _mm512_cmpeq_epi8_mask(x, cN);
    _mm_stream_si64((long long*)(P + 2 * j), _mm_cvtsi128_si64(r));            // unit arrays are only 8- / 4-byte aligned
    _mm_stream_si64((long long*)(P + 2 * j + 1), _mm_extract_epi64(r, 1));
    _mm_stream_si32((int*)(NM + 2 * j), (int)(uint32_t)nn);
    _mm_stream_si32((int*)(NM + 2 * j + 1), (int)(uint32_t)(nn >> 32));
    any_n |= nn;
  }
  _mm_sfence();
  if (n % 64) {
    pack_contig_avx2(s + 64 * full, n % 64, P + 2 * full, NM + 2 * full);
    for (size_t u = 2 * full; u < (n + 31) / 32; u++) any_n |= NM[u];
  }
  return any_n != 0;
}
// the table-driven variant is checked against the scalar definition once per process (every byte value, every position of a
// unit) before it is used; a mismatch disables it loudly instead of corrupting sequence
inline bool vbmi_packer_ok() {
  uint8_t buf[256 * 5 + 37];
  for (size_t i = 0; i < sizeof(buf); i++) buf[i] = (uint8_t)((i * 7 + i / 256) & 0xFF);
  for (size_t i = 0; i < 256; i++) buf[256 * 4 + (i % 256)] = (uint8_t)"ACGTNacgtnUuRYKM"[i % 16];
  const size_t n = sizeof(buf), nu = (n + 31) / 32;
  uint64_t p0[64], p1[64];
  uint32_t m0[64], m1[64];
  if (nu > 64) return false;
  pack_contig_scalar(buf, n, p0, m0);
  pack_contig_avx512vbmi(buf, n, p1, m1);
  return memcmp(p0, p1, nu * 8) == 0 && memcmp(m0, m1, nu * 4) == 0;
}
#endif

// P and NM must hold (n + 31) / 32 entries; returns whether the sequence contains an 'N' (the mask has a bit set)
inline bool pack_contig(const uint8_t* s, size_t n, uint64_t* P, uint32_t* NM) {
  const size_t nu_all = (n + 31) / 32;
  auto scan = [&]() { uint32_t any = 0; for (size_t j = 0; j < nu_all; j++) any |= NM[j]; return any != 0; };
#if defined(__x86_64__)
  static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  static const bool fast512 = fast && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
  static const bool vbmi = fast512 && __builtin_cpu_supports("avx512vbmi") && getenv("SK_PACK_NO_VBMI") == nullptr &&
                           (vbmi_packer_ok() || (fprintf(stderr, "skani_b200: AVX-512 VBMI packer failed its self-check, using the compare-based one\n"), false));
  if (vbmi) return pack_contig_avx512vbmi(s, n, P, NM);    // streams its output past the caches: the flag comes with it
  if (fast512) { pack_contig_avx512(s, n, P, NM); return scan(); }
  if (fast) { pack_contig_avx2(s, n, P, NM); return scan(); }
#endif
  pack_contig_scalar(s, n, P, NM);
  return scan();
}

// which implementation pack_contig uses on this machine (reported by sk_sketch_batch's trace and checked by the GPU-box tests)
inline const char* pack_impl_name() {
#if defined(__x86_64__)
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) {
    if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) {
      if (__builtin_cpu_supports("avx512vbmi") && getenv("SK_PACK_NO_VBMI") == nullptr && vbmi_packer_ok()) return "avx512vbmi";
      return "avx512bw";
    }
    return "avx2";
  }
#endif
  return "scalar";
}

}  // namespace sk_host
