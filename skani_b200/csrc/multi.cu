// multi.cu -- `skani triangle` over several GPUs from ONE host process (the reference's pair loop,
// src/triangle.rs:71-105, is a single rayon process).  One host thread per context:
//   1. contiguous genome blocks (balanced by bases); every device runs the pipelined single-GPU triangle on its own block
//      (sk_triangle_local: upload / seed / screen / chain overlapped) and keeps its sketch set;
//   2. the MARKERS of every block are exchanged device-to-device (peer copies over NVLink when the devices differ, plain
//      device copies when contexts share a device) and every device screens the whole triangle -> the same sorted pair
//      list everywhere, from which the pairs that lie inside one block (already chained in step 1) are dropped;
//   3. the remaining cross-block pairs are cut into equal contiguous slices; a device fetches the sketches its slice
//      touches as sub-blobs INCLUDING their k-mer hash tables (sk_sketch_set_pack_subset, SK_PACK_TABLES) and chains them.
// The same steps run as one process per GPU over torch.distributed/NCCL in skani_b200/multi_gpu.py (bench.py --gpus N).
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "sk_internal.h"

extern "C" int sk_triangle_local(sk_ctx* ctx, const uint8_t* bases, const uint64_t* contig_off, uint32_t n_contigs,
                                 const uint32_t* genome_of_contig, uint32_t n_genomes, const sk_sketch_params* sp,
                                 const sk_map_params* mp, const uint64_t* name_ranks, sk_ani_result** out, uint64_t* n_out,
                                 sk_triangle_stats* stats, sk_sketch_set** set_out);

namespace {

struct PhaseBarrier {   // all threads meet; a failure reported by any of them makes everybody leave at the same barrier
  std::mutex mu;
  std::condition_variable cv;
  uint32_t n, count = 0;
  uint64_t gen = 0;
  bool failed = false;
  explicit PhaseBarrier(uint32_t n_) : n(n_) {}
  bool sync(bool ok) {
    std::unique_lock<std::mutex> lk(mu);
    if (!ok) failed = true;
    const uint64_t g = gen;
    if (++count == n) { count = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
    return !failed;
  }
};

struct Blob { void* d = nullptr; uint64_t bytes = 0; std::vector<uint64_t> meta; };

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

// Cross-block pairs -> one list per device.  The pairs of one connected component of the pair graph (a cluster of related
// genomes) stay together, so a device fetches that cluster's sketches once; with a genome order unrelated to relatedness a
// contiguous slice of the sorted list touches ~5x more genomes (skani_b200/multi_gpu.py partition_pairs is the same rule).
// Components above half a device's fair share are cut into runs of consecutive pairs; items go to the least loaded device,
// largest first (ties: first pair).  Deterministic; every list comes out sorted.
static void partition_pairs(const std::vector<uint64_t>& sorted_pairs, uint32_t W, uint32_t n_genomes, std::vector<std::vector<uint64_t>>& out) {
  out.assign(W, {});
  const size_t n = sorted_pairs.size();
  if (n == 0) return;
  if (W == 1) { out[0] = sorted_pairs; return; }
  std::vector<uint32_t> parent(n_genomes);
  for (uint32_t g = 0; g < n_genomes; g++) parent[g] = g;
  auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
  for (uint64_t p : sorted_pairs) {
    const uint32_t a = find((uint32_t)(p >> 32)), b = find((uint32_t)p);
    if (a != b) parent[std::max(a, b)] = std::min(a, b);            // root = smallest genome of the component
  }
  // pairs grouped by component root (stable: sorted inside a group)
  std::vector<std::pair<uint32_t, uint64_t>> keyed(n);
  for (size_t i = 0; i < n; i++) keyed[i] = {find((uint32_t)(sorted_pairs[i] >> 32)), sorted_pairs[i]};
  std::sort(keyed.begin(), keyed.end());
  const size_t cap = std::max<size_t>(1, (n + 2 * (size_t)W - 1) / (2 * (size_t)W));
  struct Item { size_t size, start; };
  std::vector<Item> items;
  for (size_t i = 0; i < n;) {
    size_t j = i;
    while (j < n && keyed[j].first == keyed[i].first) j++;
    for (size_t s0 = i; s0 < j; s0 += cap) items.push_back(Item{std::min(cap, j - s0), s0});
    i = j;
  }
  std::sort(items.begin(), items.end(), [&](const Item& a, const Item& b) { return a.size != b.size ? a.size > b.size : keyed[a.start].second < keyed[b.start].second; });
  std::vector<size_t> load(W, 0);
  for (const Item& it : items) {
    uint32_t r = 0;
    for (uint32_t k = 1; k < W; k++) if (load[k] < load[r]) r = k;
    for (size_t k = it.start; k < it.start + it.size; k++) out[r].push_back(keyed[k].second);
    load[r] += it.size;
  }
  for (auto& v : out) std::sort(v.begin(), v.end());
}

extern "C" int sk_triangle_multi(sk_ctx* const* ctxs, uint32_t n_ctx, const uint8_t* bases, const uint64_t* contig_off, uint32_t n_contigs,
                                 const uint32_t* genome_of_contig, uint32_t n_genomes, const sk_sketch_params* sp,
                                 const sk_map_params* mp, const uint64_t* name_ranks, sk_ani_result** out, uint64_t* n_out,
                                 sk_triangle_stats* stats) {
  if (!ctxs || n_ctx == 0 || !ctxs[0] || !out || !n_out || !sp || !mp || !contig_off || (n_contigs && !genome_of_contig)) return SK_ERR_PARAM;
  *out = nullptr; *n_out = 0;
  sk_ctx* ctx = ctxs[0];
  for (uint32_t d = 0; d < n_ctx; d++) if (!ctxs[d]) { ctx->err = "null context"; return SK_ERR_PARAM; }
  const double t_begin = now_s();
  const uint32_t W = std::min<uint32_t>(n_ctx, std::max<uint32_t>(n_genomes, 1));
  if (W == 1) {
    sk_sketch_set* set = nullptr;
    int rc = sk_triangle_local(ctx, bases, contig_off, n_contigs, genome_of_contig, n_genomes, sp, mp, name_ranks, out, n_out, stats, &set);
    if (set) sk_sketch_set_free(set);
    return rc;
  }
  // ---- genome blocks balanced by bases (every block gets at least one genome)
  std::vector<uint64_t> gbytes(n_genomes, 0);
  for (uint32_t i = 0; i < n_contigs; i++) {
    if (genome_of_contig[i] >= n_genomes || (i && genome_of_contig[i] < genome_of_contig[i - 1])) { ctx->err = "genome_of_contig must be non-decreasing and < n_genomes"; return SK_ERR_PARAM; }
    gbytes[genome_of_contig[i]] += contig_off[i + 1] - contig_off[i];
  }
  uint64_t total = 0;
  for (uint64_t b : gbytes) total += b;
  std::vector<uint32_t> gb(W + 1, 0), cb(W + 1, 0);
  {
    uint64_t acc = 0;
    uint32_t d = 1;
    for (uint32_t g = 0; g < n_genomes && d < W; g++) {
      acc += gbytes[g];
      const uint32_t left_blocks = W - d, left_genomes = n_genomes - (g + 1);
      if (acc * W >= total * d || left_genomes <= left_blocks) { gb[d++] = g + 1; }
    }
    while (d < W) { gb[d] = gb[d - 1]; d++; }
    gb[W] = n_genomes;
    for (uint32_t k = 1; k <= W; k++) gb[k] = std::max(gb[k], gb[k - 1]);
  }
  for (uint32_t d = 0; d <= W; d++) cb[d] = (uint32_t)(std::lower_bound(genome_of_contig, genome_of_contig + n_contigs, gb[d]) - genome_of_contig);
  // peer access between distinct devices (ignored when unsupported: the copies then stage through the host)
  for (uint32_t a = 0; a < W; a++)
    for (uint32_t b = 0; b < W; b++)
      if (ctxs[a]->device != ctxs[b]->device) {
        cudaSetDevice(ctxs[a]->device);
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device) == cudaSuccess && can) cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
        cudaGetLastError();
      }

  PhaseBarrier bar(W);
  std::vector<sk_sketch_set*> local(W, nullptr);
  std::vector<std::vector<sk_ani_result>> results(W);
  std::vector<Blob> mk(W);
  std::vector<std::vector<Blob>> sub(W, std::vector<Blob>(W));     // sub[src][dst]
  std::vector<std::vector<uint64_t>> cross_part(W);                 // cross-block pairs found by each device's share of the screen
  std::vector<int> rcs(W, SK_OK);
  std::vector<uint64_t> screened(W, 0);
  const bool trace = getenv("SK_TRACE") != nullptr;

  auto run = [&](uint32_t d) -> int {
    sk_ctx* c = ctxs[d];
    int rc = cudaSetDevice(c->device) == cudaSuccess ? SK_OK : SK_ERR_CUDA;   // failures travel to the next barrier: nobody waits for a thread that left
    c->cpu_share = (int)W;
    const double t0 = now_s();
    // ---- 1. own block
    if (rc == SK_OK) {
      std::vector<uint32_t> gl(cb[d + 1] - cb[d]);
      for (uint32_t i = cb[d]; i < cb[d + 1]; i++) gl[i - cb[d]] = genome_of_contig[i] - gb[d];
      sk_ani_result* r = nullptr; uint64_t nr = 0;
      sk_triangle_stats st;
      rc = sk_triangle_local(c, bases, contig_off + cb[d], cb[d + 1] - cb[d], gl.data(), gb[d + 1] - gb[d], sp, mp,
                             name_ranks ? name_ranks + gb[d] : nullptr, &r, &nr, &st, &local[d]);
      if (rc == SK_OK) {
        results[d].assign(r, r + nr);
        for (auto& x : results[d]) { x.ref_id += gb[d]; x.query_id += gb[d]; }
        screened[d] += st.n_pairs_screened;
      }
      if (r) sk_free(r);
    }
    const double t1 = now_s();
    // ---- 2. markers of every block -> every device
    if (rc == SK_OK) {
      uint64_t words = 0;
      rc = sk_sketch_set_subset_blob_size(local[d], nullptr, 0, SK_PACK_MARKERS_ONLY, &mk[d].bytes, &words);
      if (rc == SK_OK && c->arena.alloc(&mk[d].d, mk[d].bytes) != cudaSuccess) rc = SK_ERR_NOMEM;
      if (rc == SK_OK) { mk[d].meta.resize(words); rc = sk_sketch_set_pack_subset(local[d], nullptr, 0, SK_PACK_MARKERS_ONLY, mk[d].d, mk[d].meta.data()); }
    }
    if (!bar.sync(rc == SK_OK)) return rc;
    sk_sketch_set* mkset = nullptr;
    {
      std::vector<uint64_t> offs(W + 1, 0);
      for (uint32_t r = 0; r < W; r++) offs[r + 1] = offs[r] + ((mk[r].bytes + 255) & ~255ull);
      void* all = nullptr;
      if (c->arena.alloc(&all, std::max<uint64_t>(offs[W], 256)) != cudaSuccess) rc = SK_ERR_NOMEM;
      for (uint32_t r = 0; r < W && rc == SK_OK; r++)
        if (cudaMemcpyAsync((uint8_t*)all + offs[r], mk[r].d, mk[r].bytes, cudaMemcpyDefault, c->stream) != cudaSuccess) rc = SK_ERR_CUDA;
      if (rc == SK_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = SK_ERR_CUDA;
      if (rc == SK_OK) {
        std::vector<const void*> bp(W);
        std::vector<const uint64_t*> mp_(W);
        for (uint32_t r = 0; r < W; r++) { bp[r] = (uint8_t*)all + offs[r]; mp_[r] = mk[r].meta.data(); }
        rc = sk_sketch_set_unpack(c, W, bp.data(), mp_.data(), &mkset);
      }
      if (all) c->arena.release(all);
    }
    if (!bar.sync(rc == SK_OK)) { if (mkset) sk_sketch_set_free(mkset); return rc; }   // every peer has copied: the marker blobs may go
    c->arena.release(mk[d].d); mk[d].d = nullptr;
    // ---- sharded screen: this device screens the rows of ITS block against every genome before them (one sort of the markers of
    //      the blocks up to its own, rows of one block) and keeps the pairs that cross a block boundary; the W partial lists are
    //      exchanged through host memory and merged, so that every device holds the same sorted cross-block pair list
    uint64_t* pairs = nullptr; uint64_t np = 0;
    rc = sk_screen_triangle_block(c, mkset, gb[d], gb[d + 1], mp, &pairs, &np);
    sk_sketch_set_free(mkset);
    if (rc == SK_OK) {
      cross_part[d].clear();
      for (uint64_t i = 0; i < np; i++) if ((uint32_t)(pairs[i] >> 32) < gb[d]) cross_part[d].push_back(pairs[i]);
      sk_free(pairs);
    }
    if (!bar.sync(rc == SK_OK)) return rc;
    std::vector<uint64_t> cross;
    for (uint32_t r = 0; r < W; r++) cross.insert(cross.end(), cross_part[r].begin(), cross_part[r].end());
    std::sort(cross.begin(), cross.end());
    const double t2 = now_s();
    // split over the devices by connected component of the pair graph (every thread computes the same split)
    std::vector<std::vector<uint64_t>> share;
    partition_pairs(cross, W, n_genomes, share);
    auto genomes_of_slice = [&](uint32_t r, std::vector<uint32_t>& need) {
      need.clear();
      for (uint64_t pr : share[r]) { need.push_back((uint32_t)(pr >> 32)); need.push_back((uint32_t)pr); }
      std::sort(need.begin(), need.end());
      need.erase(std::unique(need.begin(), need.end()), need.end());
    };
    // ---- 3. pack what every device needs from this block (the pair list is identical everywhere, so no request round)
    std::vector<uint32_t> need, mine;
    for (uint32_t r = 0; r < W && rc == SK_OK; r++) {
      genomes_of_slice(r, need);
      if (r == d) mine = need;
      std::vector<uint32_t> loc;
      for (uint32_t g : need) if (g >= gb[d] && g < gb[d + 1]) loc.push_back(g - gb[d]);
      Blob& b = sub[d][r];
      uint64_t words = 0;
      uint32_t dummy = 0;
      rc = sk_sketch_set_subset_blob_size(local[d], loc.empty() ? &dummy : loc.data(), (uint32_t)loc.size(), SK_PACK_TABLES, &b.bytes, &words);
      if (rc == SK_OK && c->arena.alloc(&b.d, b.bytes) != cudaSuccess) rc = SK_ERR_NOMEM;
      if (rc == SK_OK) { b.meta.resize(words); rc = sk_sketch_set_pack_subset(local[d], loc.empty() ? &dummy : loc.data(), (uint32_t)loc.size(), SK_PACK_TABLES, b.d, b.meta.data()); }
    }
    if (!bar.sync(rc == SK_OK)) return rc;
    sk_sketch_set* work = nullptr;
    uint64_t remote_bytes = 0;
    {
      std::vector<uint64_t> offs(W + 1, 0);
      for (uint32_t r = 0; r < W; r++) offs[r + 1] = offs[r] + ((sub[r][d].bytes + 255) & ~255ull);
      void* all = nullptr;
      if (c->arena.alloc(&all, std::max<uint64_t>(offs[W], 256)) != cudaSuccess) rc = SK_ERR_NOMEM;
      for (uint32_t r = 0; r < W && rc == SK_OK; r++) {
        if (cudaMemcpyAsync((uint8_t*)all + offs[r], sub[r][d].d, sub[r][d].bytes, cudaMemcpyDefault, c->stream) != cudaSuccess) rc = SK_ERR_CUDA;
        if (r != d) remote_bytes += sub[r][d].bytes;
      }
      if (rc == SK_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = SK_ERR_CUDA;
      if (rc == SK_OK) {
        std::vector<const void*> bp(W);
        std::vector<const uint64_t*> mp_(W);
        for (uint32_t r = 0; r < W; r++) { bp[r] = (uint8_t*)all + offs[r]; mp_[r] = sub[r][d].meta.data(); }
        rc = sk_sketch_set_unpack(c, W, bp.data(), mp_.data(), &work);     // source-block order = ascending global ids
      }
      if (all) c->arena.release(all);
    }
    if (!bar.sync(rc == SK_OK)) { if (work) sk_sketch_set_free(work); return rc; }     // every peer has copied its sub-blobs
    for (uint32_t r = 0; r < W; r++) if (sub[d][r].d) { c->arena.release(sub[d][r].d); sub[d][r].d = nullptr; }
    sk_sketch_set_free(local[d]); local[d] = nullptr;
    const double t3 = now_s();
    // ---- chain the slice on the working set (ids -> working indices and back)
    const std::vector<uint64_t>& my_pairs = share[d];
    const uint64_t lo = 0, hi = my_pairs.size();
    if (rc == SK_OK && hi > lo) {
      if (sk_sketch_set_n_genomes(work) != mine.size()) { c->err = "fetch plan mismatch"; rc = SK_ERR_STATE; }
      std::vector<uint64_t> ranks(mine.size());
      for (size_t i = 0; i < mine.size(); i++) ranks[i] = name_ranks ? name_ranks[mine[i]] : mine[i];
      if (rc == SK_OK) rc = sk_sketch_set_set_name_ranks(work, ranks.data());
      std::vector<uint64_t> lp(hi - lo);
      for (uint64_t i = lo; i < hi; i++) {
        const uint64_t a = std::lower_bound(mine.begin(), mine.end(), (uint32_t)(my_pairs[i] >> 32)) - mine.begin();
        const uint64_t b = std::lower_bound(mine.begin(), mine.end(), (uint32_t)my_pairs[i]) - mine.begin();
        lp[i - lo] = (a << 32) | b;
      }
      std::vector<sk_ani_result> res(lp.size());
      if (rc == SK_OK) rc = sk_chain_pairs(c, work, work, lp.data(), lp.size(), mp, res.data());
      if (rc == SK_OK)
        for (auto& x : res)
          if (x.ani > 0.1f) { x.ref_id = mine[x.ref_id]; x.query_id = mine[x.query_id]; results[d].push_back(x); }   // src/triangle.rs:99
      screened[d] += hi - lo;
    }
    if (work) sk_sketch_set_free(work);
    if (trace) fprintf(stderr, "[sk_triangle_multi] device slot %u (gpu %d): block %u..%u local %.1f ms, markers+screen %.1f ms, fetch %.1f ms "
                               "(%zu genomes, %.1f MB remote), chain %.1f ms (%llu cross-block pairs of %zu)\n", d, c->device, gb[d], gb[d + 1],
                       (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, mine.size(), remote_bytes / 1e6, (now_s() - t3) * 1e3,
                       (unsigned long long)(hi - lo), cross.size());
    return rc;
  };

  std::vector<std::thread> th;
  for (uint32_t d = 0; d < W; d++)
    th.emplace_back([&, d] {
      rcs[d] = run(d);
    });
  for (auto& t : th) t.join();
  int rc = SK_OK;
  for (uint32_t d = 0; d < W; d++) {
    if (rcs[d] != SK_OK && rc == SK_OK) { rc = rcs[d]; if (d) ctx->err = "device slot " + std::to_string(d) + ": " + ctxs[d]->err; }
    cudaSetDevice(ctxs[d]->device);
    if (local[d]) sk_sketch_set_free(local[d]);
    if (mk[d].d) ctxs[d]->arena.release(mk[d].d);
    for (uint32_t r = 0; r < W; r++) if (sub[d][r].d) ctxs[d]->arena.release(sub[d][r].d);
  }
  cudaSetDevice(ctx->device);
  if (rc != SK_OK) return rc;
  size_t nres = 0;
  for (auto& v : results) nres += v.size();
  sk_ani_result* o = (sk_ani_result*)malloc(sizeof(sk_ani_result) * std::max<size_t>(nres, 1));
  if (!o) return SK_ERR_NOMEM;
  size_t k = 0;
  for (auto& v : results) { if (!v.empty()) memcpy(o + k, v.data(), v.size() * sizeof(sk_ani_result)); k += v.size(); }
  *out = o; *n_out = nres;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->t_total = now_s() - t_begin;
    for (uint64_t s : screened) stats->n_pairs_screened += s;
    stats->n_pairs_kept = nres;
  }
  return SK_OK;
}
