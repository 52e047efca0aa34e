// triangle.cu -- whole `skani triangle` hot path from host buffers (reference src/triangle.rs:13-105):
// sketch every genome, marker screen (rows i, columns j > i), chain every passing pair, keep ani > 0.1.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sk_internal.h"

extern "C" int sk_triangle(sk_ctx* ctx, const uint8_t* bases, const uint64_t* contig_off, uint32_t n_contigs,
                           const uint32_t* genome_of_contig, uint32_t n_genomes, const sk_sketch_params* sp,
                           const sk_map_params* mp, sk_ani_result** out, uint64_t* n_out, sk_triangle_stats* stats) {
  if (!ctx || !out || !n_out || !sp || !mp) return SK_ERR_PARAM;
  *out = nullptr; *n_out = 0;
  SK_CUDA(cudaSetDevice(ctx->device));
  cudaEvent_t ev[4];
  for (auto& e : ev) SK_CUDA(cudaEventCreate(&e));
  struct EG { cudaEvent_t* e; ~EG() { for (int i = 0; i < 4; i++) cudaEventDestroy(e[i]); } } eg{ev};
  SK_CUDA(cudaEventRecord(ev[0], ctx->stream));
  sk_sketch_set* set = nullptr;
  SK_TRY(sk_sketch_batch(ctx, bases, contig_off, n_contigs, genome_of_contig, n_genomes, sp, &set));
  struct SG { sk_sketch_set* s; ~SG() { sk_sketch_set_free(s); } } sg{set};
  SK_CUDA(cudaEventRecord(ev[1], ctx->stream));
  uint64_t* pairs = nullptr;
  uint64_t np = 0;
  SK_TRY(sk_screen_triangle(ctx, set, mp, &pairs, &np));
  struct PG { uint64_t* p; ~PG() { free(p); } } pg{pairs};
  SK_CUDA(cudaEventRecord(ev[2], ctx->stream));
  std::vector<sk_ani_result> res(np);
  SK_TRY(sk_chain_pairs(ctx, set, set, pairs, np, mp, res.data()));
  SK_CUDA(cudaEventRecord(ev[3], ctx->stream));
  SK_CUDA(cudaEventSynchronize(ev[3]));
  uint64_t kept = 0;
  for (auto& r : res) if (r.ani > 0.1f) kept++;                    // src/triangle.rs:99 (NaN and -1 fail)
  sk_ani_result* o = (sk_ani_result*)malloc(sizeof(sk_ani_result) * (kept ? kept : 1));
  if (!o) return SK_ERR_NOMEM;
  uint64_t w = 0;
  for (auto& r : res) if (r.ani > 0.1f) o[w++] = r;
  *out = o; *n_out = kept;
  if (stats) {
    float a, b, c, t;
    cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
    cudaEventElapsedTime(&t, ev[0], ev[3]);
    stats->t_sketch = a * 1e-3; stats->t_screen = b * 1e-3; stats->t_chain = c * 1e-3; stats->t_total = t * 1e-3;
    stats->n_pairs_screened = np; stats->n_pairs_kept = kept;
  }
  return SK_OK;
}
