// Native executor for one fused LinK block forward (ELKBlock.forward, linkencoder.py:124-185):
// a single C-ABI call enqueues the whole kernel sequence on the stream
//
//   hash -> table build -> submanifold kernel map        (skipped when the caller passes a map)
//   conv plan (tile-skipping order of the kernel map)
//   block keys + radix sort/unique (key packing fused into the first pass) -> neighbour-block table
//   (+ zeroing of the block sums in the same launch)
//   -- wait for the features if they are still being uploaded (cudaStreamWaitEvent) --
//   pre_mix  = LN(x W^T)                                  (tcgen05 / FFMA)
//   local    = SubM 3^3 conv(x)                           (tcgen05 / FFMA)
//   segmented pre-aggregation (block order) -> window mean -> apply (+LN, +LN(local), add, ReLU)
//
// so the host pays one FFI crossing and one workspace allocation per block instead of ~27
// Python-level launches (the reference: ~150-180 launches, >= 4 device syncs, 4 cudaMalloc/Free).
// Nothing here synchronises or allocates: the caller provides one workspace arena.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

static inline int64_t al256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct BlockWs {
  int64_t table, kmap, plan_ws, fin, local, uniq, inverse, order, srank, seg, num, nbr, sums, mean, sort_ws, total;
  int64_t table_cap;
};

static BlockWs plan(int64_t n, int c, int kc, int r3, int kvol, bool need_kmap, bool need_plan = true) {
  BlockWs w;
  int64_t o = 0;
  const int64_t cp_bytes = need_plan ? al256(lk_conv_plan_ws_bytes(n)) : 0;
  w.table_cap = lk_table_capacity(n);
  w.table = o;   o += need_kmap ? al256(w.table_cap * 16) : 0;
  w.kmap = o;
  w.plan_ws = o; o += cp_bytes;
  w.fin = o;     o += al256(n * c * 4);
  w.local = o;   o += al256(n * c * 4);
  w.uniq = o;    o += al256(n * 8);
  w.inverse = o; o += al256(n * 4);
  w.order = o;   o += al256(n * 4);
  w.srank = o;   o += al256(n * 4);
  w.seg = o;     o += al256((n + 1) * 4);
  w.num = o;     o += 256;
  w.nbr = o;     o += al256(n * (int64_t)r3 * 4);
  w.sums = o;    o += al256(n * (int64_t)kc * 4);
  w.mean = o;    o += al256(n * (int64_t)kc * 4);
  w.sort_ws = o; o += al256(lk_sort_unique_ws_bytes(n));
  w.total = o;
  return w;
}

extern "C" int64_t lk_elk_block_ws_bytes(int64_t n, int c, int op, int r3, int kvol, int need_kmap) {
  int kc = (op == LK_OP_COSX ? 3 : 2) * c;
  return plan(n, c, kc, r3, kvol, need_kmap != 0).total;
}

#define LK_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != LK_OK) return rc__; \
  } while (0)

// side stream + fork/join events of the two-chain schedule, created once per device on first use
// (the only CUDA objects the library owns; no memory is allocated, nothing is synchronised)
struct BlockSide {
  cudaStream_t stream;
  cudaEvent_t fork, join, table;
};
static BlockSide* block_side() {
  static BlockSide sides[64];
  static bool made[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!made[dev]) {
    // lowest priority: when both chains have work ready, the CTAs of the conv (critical path) go first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char* pe = getenv("LINKB200_SIDE_PRIORITY");
    const int prio = (pe && pe[0] == '0') ? 0 : prio_lo;
    if (cudaStreamCreateWithPriority(&sides[dev].stream, cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&sides[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&sides[dev].join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&sides[dev].table, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    made[dev] = true;
  }
  return &sides[dev];
}

extern "C" int lk_elk_block_fwd(const lk_elk_block_args_t* a, lk_stream_t s) {
  LK_REQUIRE(a && a->n >= 0 && a->d_coords && a->d_feats && a->d_out && a->d_ws,
             "lk_elk_block_fwd: null argument");
  const int64_t n = a->n;
  const int c = a->gen.c, op = a->gen.op;
  const int kc = (op == LK_OP_COSX ? 3 : 2) * c;
  const bool need_kmap = a->build_kmap != 0;
  LK_REQUIRE(a->d_kmap, "lk_elk_block_fwd: d_kmap buffer is required");
  if (n == 0) return LK_OK;
  BlockWs w = plan(n, c, kc, a->r3, a->kvol, need_kmap);
  if (a->ws_bytes < w.total) {
    lk_set_error("lk_elk_block_fwd: workspace %lld < %lld bytes", (long long)a->ws_bytes, (long long)w.total);
    return LK_ENOSPC;
  }
  char* ws = (char*)a->d_ws;
  const int32_t* kmap = a->d_kmap;
  const bool tc_conv = a->use_tensor_cores && lk_conv_tc_supported(c, c) && a->d_conv_wt;
  const bool planned = tc_conv && a->d_plan_perm && a->d_plan_mask && a->kvol <= 32;
  uint64_t* uniq = (uint64_t*)(ws + w.uniq);
  int32_t* inverse = (int32_t*)(ws + w.inverse);
  int32_t* seg = (int32_t*)(ws + w.seg);
  int32_t* num = (int32_t*)(ws + w.num);
  int32_t* nbr = (int32_t*)(ws + w.nbr);
  int32_t* order = (int32_t*)(ws + w.order);
  int32_t* srank = (int32_t*)(ws + w.srank);
  float* sums = (float*)(ws + w.sums);
  float* mean = (float*)(ws + w.mean);
  float* fin = (float*)(ws + w.fin);
  float* local = (float*)(ws + w.local);

  // Two independent chains meet only in the last kernel, so they are enqueued on two streams:
  //   main : hash table -> kernel map -> conv plan -> [features ready] -> local_mix conv
  //   side : block keys + sort/unique -> neighbour table + zero -> [features ready] -> pre_mix ->
  //          pre-aggregation -> window mean
  //   join : apply (+ both LayerNorms, add, ReLU)
  // The index kernels of both chains are small, latency-bound grids (59-1184 CTAs of work that
  // leave most of the 148 SMs idle), so running the chains side by side hides one of them almost
  // completely.  Fork and join are stream-ordered events: no host or device synchronisation, and
  // still CUDA-graph capturable.
  lk_stream_t side = s;
  BlockSide* bs = a->single_stream ? nullptr : block_side();
  if (bs) {
    side = (lk_stream_t)bs->stream;
    LK_CUDA(cudaEventRecord(bs->fork, (cudaStream_t)s));
    LK_CUDA(cudaStreamWaitEvent(bs->stream, bs->fork, 0));
  }
  // ---- side chain ----
  // LINKB200_TABLE_SIDE=1 builds the hash table of the kernel map HERE, ahead of the sort, and lets the main
  // stream pick it up through an event (the main chain, table -> map -> plan -> conv, is the longer one).
  // Measured SLOWER on the same box (0.213 vs 0.197 ms per step, twice each, scripts/ab_table_side.sh): the
  // insert kernel's atomics then compete with the radix sort for the same L2 slices, and the sort gates
  // the pre-aggregation chain.  Off by default; kept as a knob.
  static const bool table_side_env = [] {
    const char* e = getenv("LINKB200_TABLE_SIDE");
    return e && e[0] == '1';
  }();
  const bool table_on_side = bs && need_kmap && table_side_env;
  if (table_on_side) {
    LK_TRY(lk_table_build_coords(a->d_coords, n, ws + w.table, w.table_cap, side));
    LK_CUDA(cudaEventRecord(bs->table, bs->stream));
  }
  LK_TRY(lk_sort_unique_coords(a->d_coords, &a->keyspec, n, a->key_bits, uniq, inverse, order, seg,
                               nullptr, num, srank, ws + w.sort_ws, lk_sort_unique_ws_bytes(n), side));
  LK_TRY(lk_block_neighbors_zero(uniq, num, n, &a->keyspec, a->d_block_offsets, a->r3, nbr, sums, kc, side));
  if (a->feats_ready) LK_CUDA(cudaStreamWaitEvent((cudaStream_t)side, (cudaEvent_t)a->feats_ready, 0));
  if (a->use_tensor_cores && (c == 32 || c == 64))
    LK_TRY(lk_linear_ln_tc_fwd(a->d_feats, a->d_premix_w, a->d_premix_g, a->d_premix_b, a->premix_eps, n, c, fin, side));
  else
    LK_TRY(lk_linear_ln_fwd(a->d_feats, a->d_premix_w, a->d_premix_g, a->d_premix_b, a->premix_eps, n, c, fin, side));
  LK_TRY(lk_link_preagg_seg_fwd(fin, a->d_coords, order, srank, n, &a->gen, sums, side));
  LK_TRY(lk_link_window_mean_seg(sums, seg, nbr, num, n, a->r3, kc, mean, side));
  if (bs) LK_CUDA(cudaEventRecord(bs->join, bs->stream));
  // ---- main chain ----
  if (need_kmap) {
    if (!table_on_side) LK_TRY(lk_table_build_coords(a->d_coords, n, ws + w.table, w.table_cap, s));
    LK_TRY(lk_kmap_query_subm_ev(a->d_coords, n, a->d_conv_offsets, a->kvol, ws + w.table, w.table_cap,
                                 a->d_kmap, table_on_side ? (void*)bs->table : nullptr, s));
  }
  if (planned && a->build_plan)
    LK_TRY(lk_conv_plan(kmap, n, a->kvol, a->d_conv_offsets, a->d_plan_perm, a->d_plan_mask,
                        ws + w.plan_ws, lk_conv_plan_ws_bytes(n), s));
  if (a->feats_ready && bs) LK_CUDA(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)a->feats_ready, 0));
  if (tc_conv) {
    lk_conv_epilogue_t ep = {nullptr, nullptr, nullptr, 0, a->conv_precision};
    if (planned)
      LK_TRY(lk_conv_tc_fwd_plan(a->d_feats, a->d_conv_wt, kmap, a->d_plan_perm, a->d_plan_mask, n,
                                 a->kvol, c, c, &ep, local, s));
    else
      LK_TRY(lk_conv_tc_fwd_ex(a->d_feats, a->d_conv_wt, kmap, n, a->kvol, c, c, &ep, local, s));
  } else {
    LK_REQUIRE(a->d_conv_w, "lk_elk_block_fwd: FFMA conv needs the untransposed weights");
    LK_TRY(lk_conv_fwd(a->d_feats, a->d_conv_w, kmap, n, a->kvol, c, c, nullptr, local, s));
  }
  // ---- join ----
  if (bs) LK_CUDA(cudaStreamWaitEvent((cudaStream_t)s, bs->join, 0));
  LK_TRY(lk_link_apply_fwd(mean, fin, a->d_coords, inverse, n, &a->gen, 1, local, a->d_g1, a->d_b1,
                           a->d_g2, a->d_b2, a->d_out, s));
  return LK_OK;
}

// ABI self-check for FFI bindings: sizeof of the argument structs as compiled into the library
extern "C" int lk_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(lk_keyspec_t);
    case 1: return (int)sizeof(lk_kernelgen_t);
    case 2: return (int)sizeof(lk_elk_block_args_t);
    case 3: return (int)sizeof(lk_conv_layer_t);
    case 4: return (int)sizeof(lk_enc_level_t);
    case 5: return (int)sizeof(lk_elk_encoder_args_t);
    default: return -1;
  }
}
