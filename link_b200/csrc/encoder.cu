// Native executor for the inference forward of the LinK encoder's backbone (ELKEncoder.forward,
// segmentation/core/models/semantic_kitti/linkencoder.py:339-375, without the classifier head): ONE
// C-ABI call enqueues
//
//   coordinate pyramid : 4 x (packed keys -> radix sort -> unique -> unpack) on four streams, every level
//                        straight from the input coordinates; the four sizes are read back once
//   index stream       : per level  strided 2^3 kernel map (+ plan), hash table, submanifold 3^3
//                        kernel map (+ plan) -- ahead of the feature kernels that need them
//   feature stream     : stem (2 convs) and per level  down conv -> 4 stage convs -> tail conv,
//                        every conv with its folded BatchNorm / shortcut / ReLU epilogue
//   block stream       : the level's LinK block (lk_elk_block_fwd) next to the level's conv stage
//   merge              : elk_tail conv with the conv branch as residual + ReLU (linkencoder.py:350)
//
// The python layer issues the same kernels as ~34 conv calls, 9 kernel-map builds and 4 block calls
// (~2 ms of interpreter + FFI time per scan, which paces the coarse levels where a conv is ~26 us of
// device time); here the host cost is ~190 launches from C.  The only host<->device synchronisations
// are the four output sizes of the strided convs, read while only index kernels are in flight.  (A
// level-by-level software pipeline -- sizes read while the previous level's convs run -- measured
// slower, 2.07 vs 1.89 ms: the persistent conv CTAs fill every SM's shared memory, the sort kernels of
// the next level only get SMs between two convs, and the host waits on them.)
// No allocation: the caller provides the level outputs and one workspace arena, both sized for n0
// rows per level (a strided level never has more voxels than its input).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

static inline int64_t en_al(int64_t x) { return (x + 255) & ~(int64_t)255; }

#define LK_TRY(call)                \
  do {                              \
    int rc__ = (call);              \
    if (rc__ != LK_OK) return rc__; \
  } while (0)

namespace {

struct EncWs {
  int64_t table[LK_ENC_MAX_LEVELS + 1], kmap3[LK_ENC_MAX_LEVELS + 1], perm3[LK_ENC_MAX_LEVELS + 1],
      mask3[LK_ENC_MAX_LEVELS + 1];
  int64_t kmap2[LK_ENC_MAX_LEVELS + 1], perm2[LK_ENC_MAX_LEVELS + 1], mask2[LK_ENC_MAX_LEVELS + 1];
  int64_t num, idx_ws, idx_ws2, idx_ws_bytes, ds_ws[LK_ENC_MAX_LEVELS], ds_ws_bytes, feat[5], elk_ws, elk_ws_bytes, total;
  int64_t table_bytes;
};

EncWs enc_plan(int64_t n0, int levels, int c_max, int elk_op, int r3) {
  EncWs w;
  int64_t o = 0;
  const int64_t tiles = (n0 + 127) / 128;
  w.table_bytes = en_al(lk_table_capacity(n0) * 16);
  for (int l = 0; l <= levels; ++l) {
    w.table[l] = o; o += w.table_bytes;
    w.kmap3[l] = o; o += en_al(27 * n0 * 4);
    w.perm3[l] = o; o += en_al(n0 * 4);
    w.mask3[l] = o; o += en_al(tiles * 4);
    w.kmap2[l] = o; o += l ? en_al(8 * n0 * 4) : 0;
    w.perm2[l] = o; o += l ? en_al(n0 * 4) : 0;
    w.mask2[l] = o; o += l ? en_al(tiles * 4) : 0;
  }
  w.num = o; o += 256;
  const int64_t a = lk_kmap_build_ws_bytes(n0, n0), b = lk_downsample_ws_bytes(n0);
  w.idx_ws_bytes = en_al(a > b ? a : b);
  w.idx_ws = o; o += w.idx_ws_bytes;       // level 0 map (feature stream)
  w.idx_ws2 = o; o += w.idx_ws_bytes;      // maps of the strided levels (index stream)
  w.ds_ws_bytes = en_al(b);
  for (int l = 0; l < LK_ENC_MAX_LEVELS; ++l) { w.ds_ws[l] = o; o += l < levels ? w.ds_ws_bytes : 0; }
  for (int i = 0; i < 5; ++i) { w.feat[i] = o; o += en_al(n0 * (int64_t)c_max * 4); }
  w.elk_ws_bytes = en_al(lk_elk_block_ws_bytes(n0, c_max, elk_op, r3, 27, 0));
  w.elk_ws = o; o += w.elk_ws_bytes;
  w.total = o;
  return w;
}

// library-owned streams / events of the three-chain schedule, once per device (no memory, no sync)
struct EncSide {
  cudaStream_t idx, blk, ds[LK_ENC_MAX_LEVELS];
  cudaEvent_t start, lvl[LK_ENC_MAX_LEVELS + 1], fork[LK_ENC_MAX_LEVELS + 1], join[LK_ENC_MAX_LEVELS + 1];
};
EncSide* enc_side() {
  static EncSide sides[64];
  static bool made[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!made[dev]) {
    EncSide& e = sides[dev];
    if (cudaStreamCreateWithFlags(&e.idx, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&e.blk, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int l = 0; l < LK_ENC_MAX_LEVELS; ++l)
      if (cudaStreamCreateWithFlags(&e.ds[l], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&e.start, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int l = 0; l <= LK_ENC_MAX_LEVELS; ++l) {
      if (cudaEventCreateWithFlags(&e.lvl[l], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&e.fork[l], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&e.join[l], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    made[dev] = true;
  }
  return &sides[dev];
}

int run_conv(const lk_conv_layer_t& L, const float* in, const int32_t* nbr, const int32_t* perm,
             const uint32_t* mask, int64_t n_out, int k, const float* residual, int precision, float* out,
             lk_stream_t s) {
  lk_conv_epilogue_t ep = {L.d_scale, L.d_shift, residual, L.relu, precision};
  return lk_conv_tc_fwd_plan(in, L.d_wimg, nbr, perm, mask, n_out, k, L.c_in, L.c_out, &ep, out, s);
}

}  // namespace

extern "C" int64_t lk_elk_encoder_ws_bytes(int64_t n0, int levels, int c_max, int elk_op, int r3) {
  if (levels < 0 || levels > LK_ENC_MAX_LEVELS) return -1;
  return enc_plan(n0, levels, c_max, elk_op, r3).total;
}

extern "C" int lk_elk_encoder_fwd(lk_elk_encoder_args_t* a, lk_stream_t s) {
  LK_REQUIRE(a && a->n0 >= 0 && a->levels >= 0 && a->levels <= LK_ENC_MAX_LEVELS, "lk_elk_encoder_fwd: bad arguments");
  const int NL = a->levels;
  const int64_t n0 = a->n0;
  for (int l = 0; l <= NL; ++l) a->n_out[l] = 0;
  if (n0 == 0) return LK_OK;
  LK_REQUIRE(a->d_coords0 && a->d_feats0 && a->d_out0 && a->d_off3_0 && a->d_ws, "lk_elk_encoder_fwd: null pointer");
  LK_REQUIRE(a->c_max > 0 && a->stem[1].c_out <= a->c_max, "lk_elk_encoder_fwd: c_max");
  for (int i = 0; i < 2; ++i)
    LK_REQUIRE(lk_conv_tc_supported(a->stem[i].c_in, a->stem[i].c_out) && a->stem[i].d_wimg,
               "lk_elk_encoder_fwd: stem conv %d: unsupported width %d -> %d", i, a->stem[i].c_in, a->stem[i].c_out);
  for (int l = 1; l <= NL; ++l) {
    const lk_enc_level_t& L = a->level[l - 1];
    LK_REQUIRE(L.d_out && L.d_coords && L.d_off2 && L.d_off3, "lk_elk_encoder_fwd: level %d: null pointer", l);
    const lk_conv_layer_t* all[7] = {&L.down, &L.stage[0], &L.stage[1], &L.stage[2], &L.stage[3], &L.tail, &L.elk_tail};
    for (int i = 0; i < 7; ++i)
      LK_REQUIRE(all[i]->d_wimg && lk_conv_tc_supported(all[i]->c_in, all[i]->c_out) && all[i]->c_out <= a->c_max &&
                     all[i]->c_in <= a->c_max,
                 "lk_elk_encoder_fwd: level %d conv %d: unsupported width %d -> %d", l, i, all[i]->c_in, all[i]->c_out);
    LK_REQUIRE(L.elk.gen.c == L.down.c_out && L.elk.d_conv_wt, "lk_elk_encoder_fwd: level %d: block width", l);
  }
  const int elk_op = NL ? a->level[0].elk.gen.op : LK_OP_COS;
  const int r3 = NL ? a->level[0].elk.r3 : 27;
  const EncWs w = enc_plan(n0, NL, a->c_max, elk_op, r3);
  if (a->ws_bytes < w.total) {
    lk_set_error("lk_elk_encoder_fwd: workspace %lld < %lld bytes", (long long)a->ws_bytes, (long long)w.total);
    return LK_ENOSPC;
  }
  EncSide* es = enc_side();
  LK_REQUIRE(es, "lk_elk_encoder_fwd: could not create the side streams");
  char* ws = (char*)a->d_ws;
  cudaStream_t main = (cudaStream_t)s;
  const lk_stream_t idx = a->single_stream ? s : (lk_stream_t)es->idx;
  const lk_stream_t blk = (a->single_stream || !a->overlap_branches) ? s : (lk_stream_t)es->blk;

  // ---- coordinate pyramid.  The sites of level l are the distinct values of floor(c0 / 2^l) 2^l, so all
  // levels derive from the INPUT coordinates (n0 is known): four independent pack -> sort -> unique ->
  // unpack chains on four streams, no level waits for the size of the one before it, and the host
  // reads the four sizes once.  (Level by level -- the reference's order -- the chains shrink with the
  // levels but each needs the previous size on the host: four dependent round trips.)  The kernel map
  // of level 0 is built on the feature stream meanwhile. ----
  const int32_t* coords[LK_ENC_MAX_LEVELS + 1];
  int64_t n[LK_ENC_MAX_LEVELS + 1];
  coords[0] = a->d_coords0;
  n[0] = n0;
  a->n_out[0] = n0;
  int32_t* d_num = (int32_t*)(ws + w.num);
  const bool multi = idx != s;
  if (multi) LK_CUDA(cudaEventRecord(es->start, main));
  for (int l = 1; l <= NL; ++l) {
    lk_enc_level_t& L = a->level[l - 1];
    const lk_stream_t ds = multi ? (lk_stream_t)es->ds[l - 1] : s;
    if (multi) LK_CUDA(cudaStreamWaitEvent((cudaStream_t)ds, es->start, 0));
    LK_TRY(lk_downsample(coords[0], n0, &L.down_spec, L.down_bits, L.d_coords, d_num + 8 * l, ws + w.ds_ws[l - 1],
                         w.ds_ws_bytes, ds));
  }
  LK_TRY(lk_kmap_build(coords[0], n0, coords[0], n0, a->d_off3_0, 27, 1, ws + w.table[0], lk_table_capacity(n0), 1,
                       (int32_t*)(ws + w.kmap3[0]), (int32_t*)(ws + w.perm3[0]), (uint32_t*)(ws + w.mask3[0]),
                       ws + w.idx_ws, w.idx_ws_bytes, s));
  if (multi) LK_CUDA(cudaEventRecord(es->lvl[0], main));
  for (int l = 1; l <= NL; ++l) {
    cudaStream_t ds = multi ? es->ds[l - 1] : main;
    int32_t h_num = 0;
    LK_CUDA(cudaMemcpyAsync(&h_num, d_num + 8 * l, 4, cudaMemcpyDeviceToHost, ds));
    LK_CUDA(cudaStreamSynchronize(ds));
    LK_REQUIRE(h_num > 0 && h_num <= n[l - 1], "lk_elk_encoder_fwd: level %d has %d output sites", l, (int)h_num);
    coords[l] = a->level[l - 1].d_coords;
    n[l] = h_num;
    a->n_out[l] = h_num;
  }

  // ---- index stream: kernel maps, hash tables and tile plans of the strided levels ----
  if (multi) LK_CUDA(cudaStreamWaitEvent((cudaStream_t)idx, es->lvl[0], 0));      // table of level 0
  for (int l = 1; l <= NL; ++l) {
    // strided map: sites of level l <- voxels of level l-1 (its table exists already)
    LK_TRY(lk_kmap_build(coords[l - 1], n[l - 1], coords[l], n[l], a->level[l - 1].d_off2, 8, 0, ws + w.table[l - 1],
                         lk_table_capacity(n[l - 1]), 0, (int32_t*)(ws + w.kmap2[l]), (int32_t*)(ws + w.perm2[l]),
                         (uint32_t*)(ws + w.mask2[l]), ws + w.idx_ws2, w.idx_ws_bytes, idx));
    LK_TRY(lk_kmap_build(coords[l], n[l], coords[l], n[l], a->level[l - 1].d_off3, 27, 1, ws + w.table[l],
                         lk_table_capacity(n[l]), 1, (int32_t*)(ws + w.kmap3[l]), (int32_t*)(ws + w.perm3[l]),
                         (uint32_t*)(ws + w.mask3[l]), ws + w.idx_ws2, w.idx_ws_bytes, idx));
    if (multi) LK_CUDA(cudaEventRecord(es->lvl[l], (cudaStream_t)idx));
  }

  // ---- feature stream ----
  float* t[5];
  for (int i = 0; i < 5; ++i) t[i] = (float*)(ws + w.feat[i]);
  const int prec = a->conv_precision;
  auto k3 = [&](int l) { return (const int32_t*)(ws + w.kmap3[l]); };
  auto p3 = [&](int l) { return (const int32_t*)(ws + w.perm3[l]); };
  auto m3 = [&](int l) { return (const uint32_t*)(ws + w.mask3[l]); };
  if (a->feats_ready) LK_CUDA(cudaStreamWaitEvent(main, (cudaEvent_t)a->feats_ready, 0));
  LK_TRY(run_conv(a->stem[0], a->d_feats0, k3(0), p3(0), m3(0), n0, 27, nullptr, prec, t[1], s));
  LK_TRY(run_conv(a->stem[1], t[1], k3(0), p3(0), m3(0), n0, 27, nullptr, prec, a->d_out0, s));
  const float* cur = a->d_out0;
  for (int l = 1; l <= NL; ++l) {
    lk_enc_level_t& L = a->level[l - 1];
    const int64_t nl = n[l];
    if (multi) LK_CUDA(cudaStreamWaitEvent(main, es->lvl[l], 0));
    float* x_in = t[0];
    LK_TRY(run_conv(L.down, cur, (const int32_t*)(ws + w.kmap2[l]), (const int32_t*)(ws + w.perm2[l]),
                    (const uint32_t*)(ws + w.mask2[l]), nl, 8, nullptr, prec, x_in, s));
    // LinK block on x_in (block stream), next to the conv stage (feature stream)
    lk_elk_block_args_t e = L.elk;
    e.n = nl;
    e.d_coords = coords[l];
    e.d_feats = x_in;
    e.d_out = t[4];
    e.d_conv_offsets = L.d_off3;
    e.d_kmap = (int32_t*)(ws + w.kmap3[l]);
    e.build_kmap = 0;
    e.build_plan = 0;
    e.d_plan_perm = (int32_t*)(ws + w.perm3[l]);
    e.d_plan_mask = (uint32_t*)(ws + w.mask3[l]);
    e.d_ws = ws + w.elk_ws;
    e.ws_bytes = w.elk_ws_bytes;
    e.feats_ready = nullptr;
    e.conv_precision = prec;
    if (blk != s) {
      LK_CUDA(cudaEventRecord(es->fork[l], main));
      LK_CUDA(cudaStreamWaitEvent((cudaStream_t)blk, es->fork[l], 0));
      LK_TRY(lk_elk_block_fwd(&e, blk));
      LK_CUDA(cudaEventRecord(es->join[l], (cudaStream_t)blk));
    }
    LK_TRY(run_conv(L.stage[0], x_in, k3(l), p3(l), m3(l), nl, 27, nullptr, prec, t[1], s));
    LK_TRY(run_conv(L.stage[1], t[1], k3(l), p3(l), m3(l), nl, 27, x_in, prec, t[2], s));
    LK_TRY(run_conv(L.stage[2], t[2], k3(l), p3(l), m3(l), nl, 27, nullptr, prec, t[1], s));
    LK_TRY(run_conv(L.stage[3], t[1], k3(l), p3(l), m3(l), nl, 27, t[2], prec, t[3], s));
    LK_TRY(run_conv(L.tail, t[3], k3(l), p3(l), m3(l), nl, 27, nullptr, prec, t[1], s));     // x_conv
    if (blk != s)
      LK_CUDA(cudaStreamWaitEvent(main, es->join[l], 0));
    else
      LK_TRY(lk_elk_block_fwd(&e, s));
    LK_TRY(run_conv(L.elk_tail, t[4], k3(l), p3(l), m3(l), nl, 27, t[1], prec, L.d_out, s));
    cur = L.d_out;
  }
  return LK_OK;
}
