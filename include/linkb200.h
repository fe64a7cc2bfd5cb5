/*
 * linkb200.h -- C ABI of liblinkb200.so: B200 (sm_100a) kernels for the LinK hot path.
 *
 * This is the drop-in boundary for the reference's native module `torchsparse.backend`
 * (reference: segmentation/torchsparse-u/torchsparse/backend/pybind_cuda.cpp:18-39, the
 * pybind11 table of 10 ops x {cpu,cuda}).  Each entry point below names the reference
 * function it replaces.  Differences from the reference interface, by design:
 *
 *   - plain C: raw device pointers + sizes, no torch types.  The CALLER owns every
 *     buffer (outputs and workspaces); the library never calls cudaMalloc/cudaFree and
 *     never synchronises the device, so every call is CUDA-graph capturable.
 *   - every call takes the CUDA stream to launch on (the reference launches on the
 *     legacy default stream, e.g. hash_cuda.cu:59,64).
 *   - return value: 0 on success, negative LK_E* on error; lk_last_error() gives a
 *     thread-local message.  (Reference: C++ exceptions / no CUDA error checks.)
 *   - fused entry points (lk_link_*, lk_kmap_*, lk_conv_*) have no single reference
 *     counterpart; each cites the reference python/CUDA lines it subsumes.
 *
 * Pointer naming: d_* = device memory, everything else is passed by value.
 * All feature matrices are row-major [rows, channels], fp32 unless stated.
 * Coordinates are int32 [rows, 4] = (x, y, z, batch)  (reference tensor.py:10-20).
 */
#ifndef LINKB200_H_
#define LINKB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LK_OK 0
#define LK_EINVAL (-1)   /* bad argument (shape, alignment, unsupported width) */
#define LK_ECUDA (-2)    /* a CUDA runtime call / launch failed */
#define LK_ENOSPC (-3)   /* caller workspace too small */

typedef void* lk_stream_t; /* cudaStream_t */

const char* lk_last_error(void);
int lk_version(void);
/* Number of kernel launches (and memsets) issued through this library by this process. */
int64_t lk_launch_count(void);
/* PCI address "dddd:bb:dd.f" of CUDA device `device` (cudaDeviceGetPCIBusId) -- the name of its sysfs
 * node; the host side pins its staging buffers on that device's NUMA node (link_b200/sharding.py:
 * bind_host_to_gpu).  No counterpart in the reference (its loaders leave placement to the OS). */
int lk_device_pci_bus_id(int device, char* buf, int len);
/* Pinned host memory for the upload path (cudaHostAlloc); write_combined = 1 for staging buffers the host
 * only writes (cudaHostAllocWriteCombined: uncached on the CPU side, DMA reads not snooped).  The reference
 * leaves pinning to torch's DataLoader (pin_memory=True, segmentation/train.py). */
int lk_host_alloc(int64_t bytes, int write_combined, void** out);
int lk_host_free(void* p);

/* ------------------------------------------------------------------------------------
 * Hashing -- replaces hash_cuda / kernel_hash_cuda (backend/hash/hash_cuda.cu:10-84).
 * Bit-exact 64-bit FNV-1a over the 4 int32 fields folded to 60 bits.
 * ---------------------------------------------------------------------------------- */
int lk_hash(const int32_t* d_coords, int64_t n, int64_t* d_out, lk_stream_t s);
/* d_out layout [K, N] (hash_cuda.cu:53); offsets int32 [K,3] are added to x,y,z. */
int lk_kernel_hash(const int32_t* d_coords, int64_t n, const int32_t* d_offsets, int k,
                   int64_t* d_out, lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * Hash table -- replaces hash_query_cuda + CuckooHashTableCuda_Multi
 * (backend/others/query_cuda.cu:9-58, backend/hashmap/hashmap_cuda.cu:9-204).
 * Open addressing over 16-byte slots {int64 key, int32 value}; capacity is a power of
 * two >= 2n.  Duplicate keys keep the LOWEST value (the reference CPU map keeps the first
 * insert, query_cpu.cpp:22-26).  Keys must not equal -1 (the empty marker); sphash
 * outputs are 60-bit non-negative.
 * ---------------------------------------------------------------------------------- */
int64_t lk_table_capacity(int64_t n);
/* lk_hash + lk_table_build in one launch: the keys are the FNV hashes of d_coords [n,4], computed on
 * the fly (same table contents). */
int lk_table_build_coords(const int32_t* d_coords, int64_t n, void* d_table, int64_t capacity,
                          lk_stream_t s);                 /* slots; bytes = 16 * slots */
int lk_table_build(const int64_t* d_keys, int64_t n, void* d_table, int64_t capacity,
                   lk_stream_t s);                    /* value of key i is i */
/* d_out[i] = value of d_queries[i] or -1 (the reference returns value+1 / 0 and python
 * subtracts 1: nn/functional/query.py:32). */
int lk_table_query(const int64_t* d_queries, int64_t nq, const void* d_table,
                   int64_t capacity, int64_t* d_out, lk_stream_t s);

/* Fused pieces of upsample_voxel (segmentation/core/models/utils.py:327-340): hash of
 * (floor(x/div), floor(y/div), floor(z/div), b), and a table probe with that hash. */
int lk_hash_div(const int32_t* d_coords, int64_t n, int div, int64_t* d_out, lk_stream_t s);
int lk_table_query_div(const int32_t* d_coords, int64_t n, int div, const void* d_table,
                       int64_t capacity, int64_t* d_out, lk_stream_t s);

/* count_cuda (backend/others/count_cuda.cu:10-31): histogram of idx>=0 into d_out[num]. */
int lk_count(const int32_t* d_idx, int64_t n, int32_t* d_out, int64_t num, lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * voxelize / devoxelize -- replace voxelize_{forward,backward}_cuda
 * (backend/voxelize/voxelize_cuda.cu:12-80) and devoxelize_{forward,backward}_cuda
 * (backend/devoxelize/devoxelize_cuda.cu:11-101).  `r3` is R = r^3 neighbours per row.
 * ---------------------------------------------------------------------------------- */
int lk_voxelize_fwd(const float* d_feats, const int32_t* d_idx, const int32_t* d_counts,
                    int64_t n, int64_t m, int c, float* d_out /*[m,c]*/, lk_stream_t s);
int lk_voxelize_bwd(const float* d_top /*[m,c]*/, const int32_t* d_idx, const int32_t* d_counts,
                    int64_t n, int64_t m, int c, float* d_bottom /*[n,c]*/, lk_stream_t s);
int lk_devoxelize_fwd(const float* d_feat /*[n,c]*/, const int32_t* d_idx /*[N,r3]*/,
                      const float* d_w /*[N,r3]*/, int64_t N, int r3, int c,
                      float* d_out /*[N,c]*/, lk_stream_t s);
int lk_devoxelize_bwd(const float* d_top /*[N,c]*/, const int32_t* d_idx, const float* d_w,
                      int64_t N, int r3, int c, int64_t n, float* d_bottom /*[n,c]*/,
                      lk_stream_t s);

/* out[i, g*c:(g+1)*c] = relu?(src_g[idx_g[i], :] + bias[g*c:(g+1)*c]) for g < groups (<= 8).
 * d_src / d_idx are HOST arrays of device pointers; idx_g == NULL is the identity map, a negative
 * index yields a zero row.  Head of ELKEncoder (linkencoder.py:371-379): nearest-parent upsample
 * (utils.py:327-340) + concat + bias + ReLU in one pass.  c % 4 == 0. */
int lk_gather_concat(const float* const* d_src, const int64_t* const* d_idx, int groups, int c,
                     int64_t n, const float* d_bias, int relu, float* d_out, lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * Key packing + radix sort/unique: the device-side replacement for torch.unique(dim=0)
 * used by voxel_to_aux (segmentation/core/models/utils.py:47), spdownsample
 * (nn/functional/downsample.py:48-50) and initial_voxelize (utils.py:239).
 *
 * A packed key is sum_f ((q[order[f]] - lo[order[f]]) << shift_f), most significant field
 * first, where q[a] = floor(coord[a] / div[a]) for a < 3 and q[3] = batch.  Ascending key
 * order == ascending signed lexicographic order on (q[order[0]], ..., q[order[3]]).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t div[3];   /* floor-divisor per spatial axis (>= 1) */
  int32_t mul[3];   /* multiplier applied on unpack (spdownsample: stride; blocks: 1) */
  int32_t order[4]; /* field order, most significant first; e.g. {0,1,2,3} or {3,0,1,2} */
  int32_t lo[4];    /* minimum of q per field (host-known bound) */
  int32_t bits[4];  /* field widths; sum <= 64 */
} lk_keyspec_t;

int lk_pack_keys(const int32_t* d_coords, int64_t n, const lk_keyspec_t* spec,
                 uint64_t* d_keys, lk_stream_t s);
int lk_unpack_keys(const uint64_t* d_keys, const int32_t* d_count /*device scalar, or NULL*/,
                   int64_t n, const lk_keyspec_t* spec, int32_t* d_coords /*[n,4]*/,
                   lk_stream_t s);

int64_t lk_sort_unique_ws_bytes(int64_t n);
/* Stable LSD radix sort of (key, row id) on the low `key_bits` bits, then unique.
 * Outputs (any may be NULL): d_unique[n] ascending distinct keys, d_inverse[n] rank of each
 * input row's key, d_order[n] row ids in sorted order (stable), d_seg[n+1] start of each
 * distinct key's run in d_order (d_seg[M] = n), d_counts[n] run lengths, d_num device
 * scalar M.  Entries past M are unspecified. */
int lk_sort_unique(const uint64_t* d_keys, int64_t n, int key_bits, uint64_t* d_unique,
                   int32_t* d_inverse, int32_t* d_order, int32_t* d_seg, int32_t* d_counts,
                   int32_t* d_num, void* d_ws, int64_t ws_bytes, lk_stream_t s);
/* lk_pack_keys + lk_sort_unique_ex with the key packing fused into the first histogram launch. */
int lk_sort_unique_coords(const int32_t* d_coords, const lk_keyspec_t* spec, int64_t n, int key_bits,
                          uint64_t* d_unique, int32_t* d_inverse, int32_t* d_order, int32_t* d_seg,
                          int32_t* d_counts, int32_t* d_num, int32_t* d_sorted_rank, void* d_ws,
                          int64_t ws_bytes, lk_stream_t s);
/* Same, plus d_sorted_rank[n] (may be NULL): rank of the key at each SORTED position, i.e.
 * d_sorted_rank[i] == d_inverse[d_order[i]] (the block row of the i-th voxel in block order;
 * feeds the segmented pre-aggregation, lk_link_preagg_seg_fwd). */
int lk_sort_unique_ex(const uint64_t* d_keys, int64_t n, int key_bits, uint64_t* d_unique,
                      int32_t* d_inverse, int32_t* d_order, int32_t* d_seg, int32_t* d_counts,
                      int32_t* d_num, int32_t* d_sorted_rank, void* d_ws, int64_t ws_bytes,
                      lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * LinK block -- fused replacement for the kernel generator + voxel_to_aux + aux_to_voxel
 * (segmentation/core/models/semantic_kitti/linkencoder.py:124-185,
 *  segmentation/core/models/utils.py:44-84, detection/det3d/models/utils/ts_elk.py:68-230).
 * ---------------------------------------------------------------------------------- */
/* Neighbour-block table (utils.py:65-73): for each of the M sorted unique block keys and
 * each of the R offsets (int32 [R,3], get_kernel_offsets order) the row of the neighbour
 * block or -1.  d_nbr is [capacity rows, R]; rows >= M are not written. */
int lk_block_neighbors(const uint64_t* d_unique, const int32_t* d_num, int64_t capacity,
                       const lk_keyspec_t* spec, const int32_t* d_offsets, int r3,
                       int32_t* d_nbr, lk_stream_t s);
/* Same, and the first M rows of d_zero [capacity, row_floats] are cleared in the same launch
 * (lk_zero_rows fused: the block-sum buffer of the pre-aggregation). */
int lk_block_neighbors_zero(const uint64_t* d_unique, const int32_t* d_num, int64_t capacity,
                            const lk_keyspec_t* spec, const int32_t* d_offsets, int r3,
                            int32_t* d_nbr, float* d_zero, int row_floats, lk_stream_t s);

#define LK_OP_COS 0   /* planes [cos, sin]           linkencoder.py:150-162 */
#define LK_OP_SIN 1   /* planes [sin, cos]           linkencoder.py:135-148 */
#define LK_OP_COSX 2  /* planes [cos, sin, linear]   linkencoder.py:164-176 */

typedef struct {
  int32_t op;          /* LK_OP_* */
  int32_t c;           /* channels C (multiple of 4, <= 256) */
  int32_t wrows;       /* rows of pos_weight: channel ch uses row ch % wrows (C/groups) */
  float coord_scale;   /* coords are divided by this before the Linear (cos_x encoder:
                          tensor stride, linkencoder.py:165; otherwise 1) */
  const float* d_pos_weight; /* [wrows, 3]   pos_weight.0.weight */
  const float* d_alpha;      /* [wrows] or NULL  (cos_x) */
  int32_t accurate_trig;     /* 0: Cody-Waite reduction + SFU sin/cos (abs err 2^-20.9, default);
                                1: libdevice sincosf (~1 ulp) */
  int32_t reserved;
} lk_kernelgen_t;

/* Zero the first *d_num rows of a [capacity, row_floats] fp32 buffer (row_floats % 4 == 0).
 * Lets block-level buffers be allocated for the worst case M = N without touching dead rows. */
int lk_zero_rows(float* d_buf, const int32_t* d_num, int64_t capacity, int row_floats,
                 lk_stream_t s);

/* Pass 1: per-block sums of the weighted planes.  d_sums [M, k*C] fp32 must be zeroed by
 * the caller, e.g. with lk_zero_rows (k = 2, or 3 for COSX).  Reads voxels in storage order; runs of equal block
 * index are reduced in registers and flushed with vector red.global.add. */
int lk_link_preagg_fwd(const float* d_fin /*[n,C]*/, const int32_t* d_coords,
                       const int32_t* d_blk /*[n] voxel -> block row*/, int64_t n,
                       const lk_kernelgen_t* gen, float* d_sums, lk_stream_t s);
/* Pass 1, segmented form (the one the block executor uses): voxels are visited in BLOCK order
 * through the sort permutation (d_order[i] = voxel row at sorted position i, d_sorted_rank[i] =
 * its block row; both from lk_sort_unique_ex).  Every lane group walks one contiguous range of
 * sorted positions (sized so that the whole grid is resident in one wave), reduces runs of equal
 * block row in registers (a warp-level segmented reduction over the variable-length voxel lists of
 * the blocks) and issues one vector reduction per block it touches: ~8x fewer L2 atomics than the
 * storage-order form, and a block that lies inside one lane group's range is summed in a fixed
 * order (deterministic).  d_sums must be zeroed.  Environment knobs (tuning only):
 * LINKB200_PREAGG=smem selects the previous one-step-per-warp kernel, LINKB200_PREAGG_Q=<q> raises
 * the range length per lane group. */
int lk_link_preagg_seg_fwd(const float* d_fin /*[n,C]*/, const int32_t* d_coords,
                           const int32_t* d_order, const int32_t* d_sorted_rank, int64_t n,
                           const lk_kernelgen_t* gen, float* d_sums, lk_stream_t s);
/* Pass 2a: window mean per block: (sum over the R neighbour blocks of sums) / (sum of
 * counts).  d_mean [M, k*C].  M read from d_num (device scalar). */
int lk_link_window_mean(const float* d_sums, const int32_t* d_counts, const int32_t* d_nbr,
                        const int32_t* d_num, int64_t capacity, int r3, int kc, float* d_mean,
                        lk_stream_t s);
/* Pass 2a with the block populations given as segment starts (n[b] = d_seg[b+1] - d_seg[b], the
 * d_seg output of lk_sort_unique*), which saves the separate counts pass. */
int lk_link_window_mean_seg(const float* d_sums, const int32_t* d_seg, const int32_t* d_nbr,
                            const int32_t* d_num, int64_t capacity, int r3, int kc, float* d_mean,
                            lk_stream_t s);
/* lk_link_window_mean_seg that also writes the window populations T[b] = sum of the neighbour blocks'
 * voxel counts as floats (d_tot [M], may be NULL): kept with d_mean for the backward pass. */
int lk_link_window_mean_tot(const float* d_sums, const int32_t* d_seg, const int32_t* d_nbr,
                            const int32_t* d_num, int64_t capacity, int r3, int kc, float* d_mean,
                            float* d_tot, lk_stream_t s);
/* Pass 2b: per voxel combine with its own phase.  Writes d_out [n,C]:
 *   fuse_norm == 0 : pre-LayerNorm value (linkencoder.py:162 / 148 / 176)
 *   fuse_norm == 1 : relu(LN(value; g1,b1) + LN(local; g2,b2)), eps 1e-6
 *                    (linkencoder.py:178-181); d_local [n,C] is local_mix.F. */
int lk_link_apply_fwd(const float* d_mean, const float* d_fin /*cos_x only, else NULL*/,
                      const int32_t* d_coords, const int32_t* d_blk, int64_t n,
                      const lk_kernelgen_t* gen, int fuse_norm, const float* d_local,
                      const float* d_g1, const float* d_b1, const float* d_g2,
                      const float* d_b2, float* d_out, lk_stream_t s);

/* Hand-written backward of the linear-kernel path with the fused norms (ops cos / sin, C in
 * {16, 32, 64, 128}); replaces autograd through devoxelize_backward (devoxelize_cuda.cu:38-59, float
 * atomics), the [N,kC] index/cat temporaries and voxelize_backward (voxelize_cuda.cu:28-42):
 *   lk_link_bwd_norm : recomputes the pre-norm value from the saved window means, both LayerNorms
 *     and the ReLU mask per voxel; writes d_dy [n,C] (gradient of the pre-norm value), d_dlocal
 *     [n,C] (gradient of local_mix.F), d_gsum [cap, 2C] = (block sums of the weighted d_dy) / T, and
 *     ADDS (dgamma1, dbeta1, dgamma2, dbeta2) into d_dparam [4,C] (caller zeroes it);
 *   lk_link_bwd_apply: sums d_gsum over the TRANSPOSED neighbourhood (d_nbr_t: the neighbour table
 *     of the negated offsets; equal to d_nbr for odd r), writes d_dfin [n,C] (gradient of F_input)
 *     and ADDS the gradient of pos_weight into d_dw [wrows,3] (caller zeroes it).
 * One warp owns a block, so the block sums need no atomics and are deterministic. */
int lk_link_bwd_supported(int c);
int lk_link_bwd_norm(const float* d_mean, const float* d_tot, const int32_t* d_seg,
                     const int32_t* d_order, const int32_t* d_num, int64_t capacity,
                     const int32_t* d_coords, const lk_kernelgen_t* gen, const float* d_local,
                     const float* d_dout, const float* d_g1, const float* d_b1, const float* d_g2,
                     const float* d_b2, float* d_dy, float* d_dlocal, float* d_gsum, float* d_dparam,
                     lk_stream_t s);
int lk_link_bwd_apply(const float* d_gsum, const float* d_mean, const int32_t* d_nbr_t,
                      const int32_t* d_seg, const int32_t* d_order, const int32_t* d_num,
                      int64_t capacity, int r3, const int32_t* d_coords, const lk_kernelgen_t* gen,
                      const float* d_fin, const float* d_dy, float* d_dfin, float* d_dw, lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * Native executor: one call enqueues the whole fused ELKBlock forward
 * (linkencoder.py:124-185): [hash -> table -> kernel map] -> pre_mix (Linear+LN) -> local_mix
 * (3^3 SubM conv) -> block keys -> sort/unique -> neighbour table -> pre-aggregation -> window
 * mean -> apply (+ both LayerNorms, add, ReLU).  One workspace arena, no sync, no allocation.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t n;                     /* active voxels */
  const int32_t* d_coords;       /* [n,4] */
  const float* d_feats;          /* [n,C]  block input (st.F) */
  float* d_out;                  /* [n,C]  block output */
  const float* d_premix_w;       /* [C,C]  pre_mix.0.weight */
  const float* d_premix_g;       /* [C]    pre_mix.1.weight */
  const float* d_premix_b;       /* [C]    pre_mix.1.bias */
  float premix_eps;
  int32_t kvol;                  /* kernel volume of local_mix (27) */
  const float* d_conv_w;         /* [kvol,C,C]  local_mix.0.kernel (FFMA path), may be NULL */
  const float* d_conv_wt;        /* packed tensor-core image of the kernel (lk_conv_tc_pack_weights) */
  const int32_t* d_conv_offsets; /* [kvol,3] kernel offsets (already scaled by the tensor stride) */
  int32_t* d_kmap;               /* caller-owned [kvol,n] kernel map of local_mix */
  int32_t build_kmap;            /* 1: build it here (hash -> table -> query) into d_kmap; 0: use it */
  int32_t build_plan;            /* 1: run lk_conv_plan on d_kmap into the three buffers below; 0: use them */
  int32_t* d_plan_perm;          /* caller-owned tile-skipping plan of the kernel map (lk_conv_plan): */
  uint32_t* d_plan_mask;         /*   [n], [ceil(n/128)]; both NULL = tensor-core conv without a plan */
  lk_keyspec_t keyspec;          /* block-key layout (div = block edge) */
  int32_t key_bits;
  int32_t r3;                    /* r^3 */
  const int32_t* d_block_offsets;/* [r3,3] neighbour-block offsets */
  lk_kernelgen_t gen;
  const float* d_g1; const float* d_b1;   /* norm        weight / bias */
  const float* d_g2; const float* d_b2;   /* norm_local  weight / bias */
  int32_t use_tensor_cores;
  int32_t conv_precision;        /* LK_PREC_FP32 (0, default) / LK_PREC_TF32 for the local_mix conv */
  void* d_ws; int64_t ws_bytes;
  int32_t single_stream;         /* 0 (default): the sort / pre-aggregation chain runs on a library-owned
                                    side stream next to the kernel-map / conv chain (fork/join by
                                    events); 1: everything on the caller's stream */
  int32_t reserved1;
  void* feats_ready;             /* cudaEvent_t or NULL: d_feats is still being uploaded on another
                                    stream; the executor enqueues all index-only work first and
                                    makes the stream wait for this event (cudaStreamWaitEvent, no
                                    host or device synchronisation) before the first feature kernel */
} lk_elk_block_args_t;
/* sizeof of the structs above as compiled (0: lk_keyspec_t, 1: lk_kernelgen_t,
 * 2: lk_elk_block_args_t, 3: lk_conv_layer_t, 4: lk_enc_level_t, 5: lk_elk_encoder_args_t): lets an FFI
 * binding verify its struct layouts at load time. */
int lk_abi_sizeof(int which);
int64_t lk_elk_block_ws_bytes(int64_t n, int c, int op, int r3, int kvol, int need_kmap);
int lk_elk_block_fwd(const lk_elk_block_args_t* args, lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * Native executor for the inference forward of the encoder backbone (ELKEncoder.forward without the
 * classifier head, linkencoder.py:339-375): stem, four strided levels of  down conv -> (conv stage ||
 * LinK block) -> merge, every Conv3d + eval-mode BatchNorm [+ shortcut] [+ ReLU] group as one tensor-core
 * conv launch with a fused epilogue, all kernel maps / hash tables / tile plans built on a library-
 * owned index stream ahead of the feature kernels, the LinK block of a level on a third stream next
 * to the level's conv stage.  One call per scan; the only synchronisations are the four 4-byte
 * read-backs of the strided levels' sizes.  Level outputs (d_out0, level[l].d_out / d_coords) and the
 * workspace are caller-owned and sized for n0 rows (a strided level never outgrows its input).
 * ---------------------------------------------------------------------------------- */
#define LK_ENC_MAX_LEVELS 4
typedef struct {
  const float* d_wimg;     /* packed tensor-core image of the kernel (lk_conv_tc_pack_weights[_ex]) */
  const float* d_scale;    /* folded eval-mode BatchNorm, [c_out] each, or NULL */
  const float* d_shift;
  int32_t c_in, c_out;     /* as packed: 32 / 64 / 128 (narrower layers zero padded by the caller) */
  int32_t relu;
  int32_t reserved;
} lk_conv_layer_t;
typedef struct {
  lk_conv_layer_t down;        /* BasicConvolutionBlock(ks = 2, stride = 2)            linkencoder.py:228 */
  lk_conv_layer_t stage[4];    /* two ResidualBlocks: conv a, conv b (+ identity shortcut) each  :231 */
  lk_conv_layer_t tail;        /* stageN_tail: Conv3d + BatchNorm                                :216 */
  lk_conv_layer_t elk_tail;    /* elkN_tail, merged with the conv branch: relu(x_conv + .)       :350 */
  lk_elk_block_args_t elk;     /* the level's LinK block: parameter fields, keyspec, key_bits, r3, gen,
                                  d_block_offsets filled by the caller; n / buffers / map by the executor */
  lk_keyspec_t down_spec;      /* key layout of the level's output sites: floor(c / 2^l) 2^l, order (b,x,y,z) */
  int32_t down_bits;
  int32_t reserved;
  const int32_t* d_off2;       /* [8,3]  offsets of the 2^3 strided conv at the INPUT level's stride */
  const int32_t* d_off3;       /* [27,3] offsets of the 3^3 convs at THIS level's stride */
  float* d_out;                /* [n0, c] level output x_l (first n_out[l] rows valid) */
  int32_t* d_coords;           /* [n0, 4] level coordinates (first n_out[l] rows valid) */
} lk_enc_level_t;
typedef struct {
  int64_t n0;
  const int32_t* d_coords0;    /* [n0,4] */
  const float* d_feats0;       /* [n0, stem[0].c_in] input features, channel-padded by the caller */
  void* feats_ready;           /* cudaEvent_t or NULL (see lk_elk_block_args_t) */
  int32_t levels;              /* strided levels, <= LK_ENC_MAX_LEVELS */
  int32_t c_max;               /* widest feature row of any layer */
  int32_t conv_precision;      /* LK_PREC_FP32 / LK_PREC_TF32 */
  int32_t single_stream;       /* 1: everything on the caller's stream */
  int32_t overlap_branches;    /* 1: LinK block of a level on its own stream next to the conv stage */
  int32_t reserved;
  lk_conv_layer_t stem[2];
  const int32_t* d_off3_0;     /* [27,3] offsets at stride 1 */
  float* d_out0;               /* [n0, stem[1].c_out] */
  lk_enc_level_t level[LK_ENC_MAX_LEVELS];
  void* d_ws; int64_t ws_bytes;
  int64_t n_out[LK_ENC_MAX_LEVELS + 1];   /* written (host): active voxels per level */
} lk_elk_encoder_args_t;
int64_t lk_elk_encoder_ws_bytes(int64_t n0, int levels, int c_max, int elk_op, int r3);
int lk_elk_encoder_fwd(lk_elk_encoder_args_t* args, lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * Training-mode BatchNorm over sparse feature rows, fused with the shortcut add and the ReLU that follow it:
 *     y = relu?( (x - mean) invstd gamma + beta [+ residual] ),   x, y, residual [n, c] fp32, c % 4 == 0.
 * Replaces spnn.BatchNorm (= nn.BatchNorm1d over .feats, torchsparse/nn/modules/norm.py:10-13) + spnn.ReLU /
 * the ResidualBlock add (linkencoder.py:26-37, 64-91) and nn.BatchNorm1d + nn.ReLU on .features in the
 * detection backbone (scn.py:64-107), with nn.BatchNorm1d's semantics: biased batch variance for the
 * normalisation, running_var updated with the unbiased one, running = (1 - momentum) running + momentum
 * batch, num_batches_tracked += 1 (pointers may be NULL: track_running_stats = False).  d_save_mean /
 * d_save_invstd [c] are kept for the backward.  Backward: g = dy * (y > 0) when d_y (the saved OUTPUT of
 * the forward, ReLU case) is given, else g = dy;  d_dgamma = sum g xhat, d_dbeta = sum g,
 * d_dx = gamma invstd (g - dbeta / n - xhat dgamma / n),  d_dresidual = g (optional).
 * d_ws: lk_bn_ws_bytes(c) bytes (per-channel double accumulators; zeroed by the call).
 * ---------------------------------------------------------------------------------- */
int lk_bn_supported(int c);
int64_t lk_bn_ws_bytes(int c);
int lk_bn_train_fwd(const float* d_x, const float* d_residual /*or NULL*/, int64_t n, int c,
                    const float* d_gamma /*or NULL*/, const float* d_beta /*or NULL*/, float eps, float momentum,
                    int relu, float* d_running_mean, float* d_running_var, int64_t* d_num_batches_tracked,
                    float* d_save_mean, float* d_save_invstd, float* d_y, void* d_ws, int64_t ws_bytes,
                    lk_stream_t s);
int lk_bn_train_bwd(const float* d_dy, const float* d_x, const float* d_y /*or NULL*/, int64_t n, int c,
                    const float* d_save_mean, const float* d_save_invstd, const float* d_gamma /*or NULL*/,
                    float* d_dx, float* d_dresidual /*or NULL*/, float* d_dgamma /*or NULL*/,
                    float* d_dbeta /*or NULL*/, void* d_ws, int64_t ws_bytes, lk_stream_t s);

/* Fused bias-free Linear + LayerNorm: out = LN(x @ W^T; gamma, beta, eps); x, out [n, c],
 * W [c, c] (nn.Linear layout).  ELKBlock.pre_mix (linkencoder.py:112-115).  c in {16,32,64,128}. */
int lk_linear_ln_fwd(const float* d_x, const float* d_w, const float* d_gamma, const float* d_beta,
                     float eps, int64_t n, int c, float* d_out, lk_stream_t s);

/* Same contract on the tcgen05 tensor cores (kind::tf32 with the 3xTF32 split => fp32-level
 * accuracy, accumulators in TMEM, LayerNorm in the TMEM epilogue).  c in {32,64}. */
int lk_linear_ln_tc_fwd(const float* d_x, const float* d_w, const float* d_gamma,
                        const float* d_beta, float eps, int64_t n, int c, float* d_out,
                        lk_stream_t s);

/* ------------------------------------------------------------------------------------
 * Kernel maps + sparse convolution -- replace the python kmap build
 * (nn/functional/conv.py:103-122) and convolution_{forward,backward}_cuda
 * (backend/convolution/convolution_cuda.cu:53-278).
 *
 * Our kernel map is OUTPUT-STATIONARY: d_nbr int32 [K, n_out], d_nbr[k, o] = input row
 * feeding output row o through kernel offset k, or -1 (this is the reference's `results`
 * tensor, conv.py:114).  The reference's (nbmaps, nbsizes) pair is the compaction of the
 * non-negative entries in (k, o) order.
 * ---------------------------------------------------------------------------------- */
/* d_table: hash table built over lk_hash(input coords).  offsets int32 [K,3]. */
int lk_kmap_query(const int32_t* d_out_coords, int64_t n_out, const int32_t* d_offsets, int k,
                  const void* d_table, int64_t capacity, int32_t* d_nbr, lk_stream_t s);
/* Submanifold special case (output coords == input coords, odd kernel => offsets[K-1-k] ==
 * -offsets[k]): only the first K/2 offsets are probed and each hit (i, j) fills both nbr[k, i] = j
 * and nbr[K-1-k, j] = i; the centre offset is the identity.  Same result as lk_kmap_query. */
int lk_kmap_query_subm(const int32_t* d_coords, int64_t n, const int32_t* d_offsets, int k,
                       const void* d_table, int64_t capacity, int32_t* d_nbr, lk_stream_t s);
/* Same, with the hash table built on another stream: the query waits for `table_ready` (a cudaEvent_t
 * recorded after lk_table_build* there, or NULL) after it has cleared the map on `s`. */
int lk_kmap_query_subm_ev(const int32_t* d_coords, int64_t n, const int32_t* d_offsets, int k,
                          const void* d_table, int64_t capacity, int32_t* d_nbr, void* table_ready,
                          lk_stream_t s);
/* Transposed relation: d_inv [K, n_in] (prefilled with -1 by this call):
 * d_inv[k, i] = o  whenever d_nbr[k, o] = i. */
int lk_kmap_invert(const int32_t* d_nbr, int64_t n_out, int k, int64_t n_in, int32_t* d_inv,
                   lk_stream_t s);
/* out[o, :] = sum_k in[nbr[k, o], :] @ W[k]   (+ bias).  W is [K, c_in, c_out] row-major
 * (module parameter layout, nn/modules/conv.py:33-38).  No atomics, out fully written. */
int lk_conv_fwd(const float* d_in, const float* d_w, const int32_t* d_nbr, int64_t n_out,
                int k, int c_in, int c_out, const float* d_bias /*or NULL*/, float* d_out,
                lk_stream_t s);
/* Fused epilogue for both convolution kernels:  y = relu?( acc * scale + shift + residual ).
 * Folds an eval-mode BatchNorm (scale = gamma / sqrt(var + eps), shift = beta - mean * scale), a
 * bias, the residual add of ResidualBlock (linkencoder.py:89-91) and the ReLU into the single
 * output store.  Any pointer may be NULL (scale -> 1, shift -> 0, residual -> none). */
typedef struct {
  const float* d_scale;     /* [c_out] or NULL */
  const float* d_shift;     /* [c_out] or NULL */
  const float* d_residual;  /* [n_out, c_out] or NULL */
  int32_t relu;
  int32_t precision;        /* tensor-core kernels only: LK_PREC_FP32 (0, default) = 3xTF32, fp32-level accuracy;
                               LK_PREC_TF32 (1) = single-pass TF32 (operands truncated to tf32, ~1e-3 relative) */
} lk_conv_epilogue_t;
#define LK_PREC_FP32 0
#define LK_PREC_TF32 1
int lk_conv_fwd_ex(const float* d_in, const float* d_w, const int32_t* d_nbr, int64_t n_out, int k,
                   int c_in, int c_out, const lk_conv_epilogue_t* ep, float* d_out, lk_stream_t s);
int lk_conv_tc_fwd_ex(const float* d_in, const float* d_wimg, const int32_t* d_nbr, int64_t n_out,
                      int k, int c_in, int c_out, const lk_conv_epilogue_t* ep, float* d_out,
                      lk_stream_t s);

/* Same contraction on the tcgen05 tensor cores (kind::tf32, 3xTF32 split => fp32-level accuracy):
 * weight-stationary persistent CTAs, the accumulators of up to 512/c_out output tiles resident in
 * TMEM, gathered rows written straight into TMEM (A operand), weights fetched as packed images
 * (d_wimg, see lk_conv_tc_pack_weights).  lk_conv_tc_supported: c_in, c_out in {32, 64}. */
int lk_conv_tc_supported(int c_in, int c_out);
int lk_conv_tc_fwd(const float* d_in, const float* d_wimg, const int32_t* d_nbr, int64_t n_out,
                   int k, int c_in, int c_out, const float* d_bias /*or NULL*/, float* d_out,
                   lk_stream_t s);
/* Tile-skipping plan for lk_conv_tc_fwd_plan (K <= 32).  Output rows are grouped by the signs of
 * the offsets they use (a counting sort on a <= 8-bit class derived from d_offsets int32 [K,3];
 * with K <= 8 or d_offsets == NULL the class is the K-bit presence mask itself), so that 128-row
 * tiles are homogeneous and (offset, tile) steps without a single pair can be skipped.
 * Outputs: d_perm [n_out] plan position -> output row (tile t = positions [128 t, 128 t + 128));
 * d_tile_mask [ceil(n_out/128)] bit k set iff some row of the tile has a neighbour at offset k.
 * Results of the convolution do not depend on the plan (same per-row sums in the same order). */
int64_t lk_conv_plan_ws_bytes(int64_t n_out);
int lk_conv_plan(const int32_t* d_nbr, int64_t n_out, int k, const int32_t* d_offsets /*or NULL*/,
                 int32_t* d_perm, uint32_t* d_tile_mask, void* d_ws, int64_t ws_bytes,
                 lk_stream_t s);
/* Weights of the tensor-core conv as the kernel consumes them: per offset k one contiguous image
 * of the shared-memory B operand (tf32 hi plane then lo plane, each [c_in/32] K-blocks of
 * [c_out rows x 128 bytes] in the SWIZZLE_128B pattern), so that the kernel fetches W[k] with ONE
 * asynchronous bulk copy.  d_wt [K, c_out, c_in] (per-offset transpose of the module parameter);
 * d_img: K * 8 * c_in * c_out bytes.  Packed once per weight update (cached by the caller). */
int lk_conv_tc_pack_weights(const float* d_wt, int k, int c_in, int c_out, float* d_img,
                            lk_stream_t s);
/* General form: d_w is [K, src_c_in, src_c_out] (layout 1: the module parameter, transposed on the fly) or
 * [K, src_c_out, src_c_in] (layout 0), zero padded to the image's c_in x c_out, offsets reversed when
 * flip_k (the input gradient of a submanifold conv): no transposed / padded / flipped copy is made. */
int lk_conv_tc_pack_weights_ex(const float* d_w, int k, int c_in, int c_out, int src_c_in, int src_c_out,
                               int layout, int flip_k, float* d_img, lk_stream_t s);
/* The tensor-core conv on packed weights, optionally with a plan: d_perm and d_tile_mask as
 * produced by lk_conv_plan (both NULL = identity order, no skipping). */
int lk_conv_tc_fwd_plan(const float* d_in, const float* d_wimg, const int32_t* d_nbr,
                        const int32_t* d_perm, const uint32_t* d_tile_mask, int64_t n_out, int k,
                        int c_in, int c_out, const lk_conv_epilogue_t* ep, float* d_out,
                        lk_stream_t s);
/* The same kernel on bf16 feature rows: d_in [n_in, c_in], ep->d_residual [n_out, c_out] and d_out
 * [n_out, c_out] are bf16 (scale / shift stay fp32); accumulation in fp32, weights rounded to tf32 (one
 * MMA per k-slice: a bf16 value is exact in tf32).  c_in in {32, 64, 128}, c_out in {32, 64}.  The reduced-
 * precision path of BASELINE config 3 (the reference runs this conv in fp16 under autocast,
 * torchsparse/nn/functional/conv.py:19). */
int lk_conv_tc_bf16_supported(int c_in, int c_out);
int lk_conv_tc_fwd_bf16(const void* d_in, const float* d_wimg, const int32_t* d_nbr,
                        const int32_t* d_perm, const uint32_t* d_tile_mask, int64_t n_out, int k,
                        int c_in, int c_out, const lk_conv_epilogue_t* ep, void* d_out, lk_stream_t s);
/* Composite index builds (one call = the launches of several entry points above; they exist to
 * keep the host side off the critical path of the encoder forward).
 * lk_kmap_build: [hash(d_in_coords) -> table build ->] kernel-map query [-> lk_conv_plan]
 *   (the kmap branch of F.conv3d, nn/functional/conv.py:103-121).  d_table is caller-owned
 *   (lk_table_capacity(n_in) * 16 bytes) and reusable by later maps of the same input level with
 *   build_table = 0; subm = 1 selects the symmetric submanifold query; plan pointers may be NULL.
 * lk_downsample: output coordinates of a strided conv whose kernel equals its stride
 *   (F.spdownsample, nn/functional/downsample.py:11-51): packed keys -> radix sort -> unique ->
 *   unpack; d_out_coords [n,4] with *d_num valid rows in the reference's torch.unique order. */
int64_t lk_kmap_build_ws_bytes(int64_t n_in, int64_t n_out);
int lk_kmap_build(const int32_t* d_in_coords, int64_t n_in, const int32_t* d_out_coords, int64_t n_out,
                  const int32_t* d_offsets, int k, int subm, void* d_table, int64_t capacity,
                  int build_table, int32_t* d_nbr, int32_t* d_plan_perm, uint32_t* d_plan_mask,
                  void* d_ws, int64_t ws_bytes, lk_stream_t s);
int64_t lk_downsample_ws_bytes(int64_t n);
int lk_downsample(const int32_t* d_coords, int64_t n, const lk_keyspec_t* spec, int key_bits,
                  int32_t* d_out_coords, int32_t* d_num, void* d_ws, int64_t ws_bytes, lk_stream_t s);
/* ------------------------------------------------------------------------------------
 * Voxelisation front-end of the detection pipeline: points -> voxels with the exact semantics of
 * det3d/ops/point_cloud/point_cloud_ops.py:112-183 (`points_to_voxel`, reverse_index=True): voxel
 * ids in order of first appearance, at most max_voxels voxels, the first max_points points of a
 * voxel in point order, coordinates (z, y, x).  voxel_size[3] / coors_range[6] are HOST arrays
 * (float32, like the numpy arguments of the reference).  d_voxels [max_voxels, max_points, ndim]
 * (rows of existing voxels fully written, zero padded), d_coors [max_voxels, 3], d_num_points
 * [max_voxels], d_voxel_num device scalar.  Deterministic (two stable radix sorts, no atomics).
 * ---------------------------------------------------------------------------------- */
int64_t lk_points_to_voxel_ws_bytes(int64_t n);
int lk_points_to_voxel(const float* d_points, int64_t n, int ndim, const float* voxel_size,
                       const float* coors_range, int max_points, int max_voxels, float* d_voxels,
                       int32_t* d_coors, int32_t* d_num_points, int32_t* d_voxel_num, void* d_ws,
                       int64_t ws_bytes, lk_stream_t s);
/* Reference-layout kernel map -> the output-stationary map of the conv kernels: d_nbmaps int32 [P, 2]
 * (input row, output row) ordered by offset, h_nbsizes int32 [K] ON THE HOST -- exactly the
 * (neighbor_map, neighbor_offset) arguments of convolution_{forward,backward}_cuda
 * (torchsparse/backend/pybind_cuda.cpp:20-21, convolution_cuda.cu:53-57).  d_nbr [K, n_rows] = partner
 * row of the pair whose column row_col is the row, or -1; identity_mid = 1: the centre offset maps
 * every row to itself (the reference's precompute_mid shortcut, convolution_cuda.cu:74-88). */
int lk_kmap_from_pairs(const int32_t* d_nbmaps, const int32_t* h_nbsizes, int k, int64_t n_rows,
                       int row_col, int identity_mid, int32_t* d_nbr, lk_stream_t s);
/* Candidate output sites of a strided / padded sparse conv that creates new active sites (spconv
 * SparseConv3d as used by detection/det3d/models/backbones/scn.py:494-566; spconv itself is not in the
 * reference tree): d_indices int32 [n,4] = (batch, z, y, x); kernel / stride / padding /
 * out_shape are HOST int32[3] in (z, y, x) order; d_cand int32 [n * cap, 4] = (x, y, z, batch) with
 * cap = prod ceil(k_a / s_a); invalid slots are (0, 0, 0, batch_size) and set *d_any_invalid (device int).
 * Sort-unique d_cand by (batch, z, y, x) and drop the trailing sentinel to get the output sites. */
int lk_strided_candidates(const int32_t* d_indices, int64_t n, const int32_t* kernel3, const int32_t* stride3,
                          const int32_t* padding3, const int32_t* out_shape3, int batch_size,
                          int32_t* d_cand, int32_t* d_any_invalid, lk_stream_t s);
/* ------------------------------------------------------------------------------------
 * Sparse -> dense BEV scatter of the detection backbone output and its transpose (the backward).
 * Replaces spconv's `.dense()` + permute in `ret = self.extra_conv(x).dense(); ret.view(N, C*D, H, W)`
 * (detection/det3d/models/backbones/scn.py:612-617).  d_indices int32 [n,4] = (batch, z, y, x), rows
 * unique; d_dense [batch, c, depth, height, width] channels-first (== [batch, c*depth, height, width]):
 * lk_bev_scatter zero-fills it and writes every row; lk_bev_gather reads grad rows back. */
int lk_bev_scatter(const float* d_feats, const int32_t* d_indices, int64_t n, int c, int batch,
                   int depth, int height, int width, float* d_dense, lk_stream_t s);
int lk_bev_gather(const float* d_dense, const int32_t* d_indices, int64_t n, int c, int batch,
                  int depth, int height, int width, float* d_feats, lk_stream_t s);
/* HOST helper (no device work): column-wise min / max of a host int32 [n,4] coordinate array -> lo4, hi4
 * (the bounds that size the packed sort keys; SparseTensor.from_host calls it while the upload is in flight). */
int lk_host_coord_bounds(const int32_t* h_coords, int64_t n, int32_t* lo4, int32_t* hi4);
/* grad_w[k] = sum_o in[nbr[k,o]]^T @ grad_out[o]; d_gw [K,c_in,c_out] zeroed by this call. */
int lk_conv_bwd_weight(const float* d_in, const float* d_gout, const int32_t* d_nbr,
                       int64_t n_out, int k, int c_in, int c_out, float* d_gw, lk_stream_t s);
/* The same product on tcgen05 (kind::tf32, 3xTF32; C_in, C_out in {32, 64, 128}, K <= 32) --
 * replaces the per-offset gather + torch::mm(in^T, grad_out) of convolution_backward_cuda
 * (torchsparse/backend/convolution/convolution_cuda.cu:217-278).  Two calls:
 *   lk_conv_wgrad_prepass   once per relation (kernel map direction): d_masks [ceil(n/64)] = offsets
 *                           with at least one pair per 64-row tile; with a tile order d_perm [n]
 *                           (lk_conv_plan) also d_nbrp [K, n] = d_nbr[:, d_perm] (both NULL otherwise);
 *   lk_conv_wgrad_tc        d_gw [K, c_in, c_out] (zeroed by the call) from d_in rows gathered through
 *                           d_nbrp (= d_nbr when there is no order) and d_gout rows d_perm[pos];
 *                           slots_per_cta <= 512 / c_out accumulator slots per CTA (0 = as many as fit). */
int lk_conv_wgrad_tc_supported(int c_in, int c_out);
int lk_conv_wgrad_prepass(const int32_t* d_nbr, const int32_t* d_perm, int64_t n, int k,
                          int32_t* d_nbrp, uint32_t* d_masks, lk_stream_t s);
int lk_conv_wgrad_tc(const float* d_in, const float* d_gout, const int32_t* d_nbrp,
                     const int32_t* d_perm, const uint32_t* d_masks, int64_t n, int k, int c_in,
                     int c_out, float* d_gw, int slots_per_cta, lk_stream_t s);
/* ------------------------------------------------------------------------------------
 * Rotated bird's-eye-view IoU of 3D boxes (x, y, z, dx, dy, dz, heading): d_out [n, m] =
 * IoU(d_a[i], d_b[j]).  Replaces iou3d_nms_cuda.boxes_iou_bev_gpu / the IoU inside nms_gpu
 * (detection/det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-414, iou3d_nms_api.cpp:11-16).
 * lk_boxes_iou_bev_hostcheck evaluates the SAME __host__ __device__ arithmetic on host arrays: a
 * test hook that pins the kernel's arithmetic to the reference's iou3d_cpu.cpp fixture without a
 * GPU -- not a fallback.
 * ---------------------------------------------------------------------------------- */
int lk_boxes_iou_bev(const float* d_a, int64_t n, const float* d_b, int64_t m, float* d_out,
                     lk_stream_t s);
/* Greedy rotated NMS over boxes already sorted by descending score: d_keep [n] (uint8) = 1 for the
 * boxes that survive (a box is dropped when an earlier KEPT box overlaps it with BEV IoU > thresh).
 * Replaces nms_gpu (iou3d_nms_kernel.cu:328-414 + the host-side greedy loop of iou3d_nms.cpp): the
 * 64 x 64 overlap bitmasks AND the greedy scan run on the device, nothing is copied to the host.
 * d_ws: lk_nms_bev_ws_bytes(n) bytes (the bitmask matrix), 8-byte aligned. */
int64_t lk_nms_bev_ws_bytes(int64_t n);
int lk_nms_bev(const float* d_boxes_sorted, int64_t n, float thresh, void* d_ws, int64_t ws_bytes,
               uint8_t* d_keep, lk_stream_t s);
/* Centre-distance NMS of CenterPoint (det3d/core/utils/circle_nms_jit.py:4-28, a numba loop on the host
 * in the reference): d_centers_sorted float32 [n, 2] = (x, y) of the boxes sorted by descending score;
 * box j is dropped when an earlier KEPT box lies within `thresh` in SQUARED distance.  Same workspace
 * and d_keep convention as lk_nms_bev. */
int lk_nms_circle(const float* d_centers_sorted, int64_t n, float thresh, void* d_ws, int64_t ws_bytes,
                  uint8_t* d_keep, lk_stream_t s);
int lk_boxes_iou_bev_hostcheck(const float* a, int64_t n, const float* b, int64_t m, float* out);

#ifdef __cplusplus
}
#endif
#endif /* LINKB200_H_ */
