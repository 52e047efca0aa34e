// temporary: entry points not implemented yet return SK_ERR_STATE (removed as they land)
#include "sk_internal.h"
extern "C" {
#define NOTYET(ctx) do { if (ctx) (ctx)->err = "not implemented yet"; return SK_ERR_STATE; } while (0)
int sk_screen_triangle(sk_ctx* ctx, const sk_sketch_set*, const sk_map_params*, uint64_t**, uint64_t*) { NOTYET(ctx); }
int sk_screen_query_ref(sk_ctx* ctx, const sk_sketch_set*, const sk_sketch_set*, const sk_map_params*, int, uint64_t**, uint64_t*) { NOTYET(ctx); }
int sk_chain_pairs(sk_ctx* ctx, const sk_sketch_set*, const sk_sketch_set*, const uint64_t*, uint64_t, const sk_map_params*, sk_ani_result*) { NOTYET(ctx); }
int sk_chain_pair_debug(sk_ctx* ctx, const sk_sketch_set*, const sk_sketch_set*, uint64_t, const sk_map_params*, sk_chain_debug*) { NOTYET(ctx); }
void sk_chain_debug_free(sk_chain_debug*) {}
int sk_triangle(sk_ctx* ctx, const uint8_t*, const uint64_t*, uint32_t, const uint32_t*, uint32_t, const sk_sketch_params*, const sk_map_params*, sk_ani_result**, uint64_t*, sk_triangle_stats*) { NOTYET(ctx); }
}
