/*
 * b200mp.h -- C ABI of the B200-native message-passing aggregation engine.
 *
 * This is the drop-in boundary.  The reference
 * (pyg-team/pytorch_geometric v2.9.0) has no FFI of its own: it late-binds a small set of
 * operator signatures (torch_scatter.*, torch.ops.torch_sparse.spmm_*, pyg_lib.ops.*) and a few
 * Python functions.  Each entry point below states which of those it replaces (file:line under
 * /root/reference/torch_geometric).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain pointers and sizes; no torch types.  All pointers are DEVICE pointers unless the
 *    name ends in _host.  Buffers are caller-owned; nothing is allocated inside.
 *  - feature matrices are row-major, contiguous, [rows, feat]; `val_dtype` selects the element
 *    type (B200MP_F32 / B200MP_BF16); accumulation is always fp32.
 *  - index arrays (`rowptr`, `col`, `index`, `perm`) share one `idx_dtype` per call
 *    (B200MP_I32 / B200MP_I64).  The reference uses int64; int32 halves index traffic and is
 *    what the engine's own graph cache stores when N, E < 2^31.
 *  - edge weights / attention values are always fp32.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *    enqueue work; they never synchronise unless documented.
 *  - return value: 0 on success, a negative B200MP_ERR_* code otherwise.  Nothing throws
 *    across the ABI.  A kernel cannot raise: out-of-range indices are undefined behaviour
 *    unless the caller checks them first with b200mp_index_stats() (min / max / sortedness in one
 *    pass); the host-side mirror raises the reference's "valid indices" IndexError
 *    (nn/conv/message_passing.py:269-290) from that result.
 */
#ifndef B200MP_H_
#define B200MP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MP_VERSION "0.1.0"

/* element / index dtypes */
enum { B200MP_F32 = 0, B200MP_BF16 = 1 };
enum { B200MP_I32 = 0, B200MP_I64 = 1 };
/* reductions (same codes as oracle/mp_oracle.c) */
enum { B200MP_SUM = 0, B200MP_MEAN = 1, B200MP_MIN = 2, B200MP_MAX = 3, B200MP_MUL = 4 };
/* errors */
enum {
    B200MP_OK = 0,
    B200MP_ERR_INVALID_ARG = -1,   /* NULL pointer, negative size, misaligned buffer */
    B200MP_ERR_UNSUPPORTED = -2,   /* dtype / reduce combination not implemented */
    B200MP_ERR_CUDA = -3,          /* a CUDA API call or launch failed; see b200mp_last_error() */
    B200MP_ERR_WORKSPACE = -4      /* workspace too small */
};

const char* b200mp_version(void);
const char* b200mp_last_error(void);          /* thread-local, human readable */
int b200mp_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);
/* Runtime switches for measurements (A/B of kernel variants; the defaults are the measured best).
 *   "spmm_impl"    0 = auto (default), 1 = lane-group-per-row kernel only, 2 = persistent TMA-fed kernel wherever legal
 *   "attn_staged"  2 = cp.async-staged attention sweeps with one-warp CTAs (default), 1 = 4-warp CTAs, 0 = register form;
 *                  0 also turns the multi-aggregation hit-bit path off
 *   "multi_tune"   6 = one-warp CTAs for the multi-aggregation row sweeps (default), 5 = 128-thread CTAs
 *   "gemm_*", "spmm_tune"   tuning knobs of the GEMM / gather kernels (see csrc/core.cu) */
int b200mp_set_option(const char* name, int value);

/* ------------------------------------------------------------------ graph structure (integer work, bit-exact)
 * Replaces: utils/_degree.py:9-31 (degree), index.py:27-37 (ptr2index / index2ptr ==
 * torch._convert_indices_from_coo_to_csr / repeat_interleave), utils/_index_sort.py:10-32 and
 * pyg_lib.ops.index_sort (stable radix sort by key), EdgeIndex.get_csr/get_csc/_sort_by_transpose
 * (edge_index.py:589-696), utils/loop.py:585-657 (add_remaining_self_loops) and :71-131/:382-492
 * (remove_self_loops + add_self_loops), nn/conv/gcn_conv.py:95-113 (gcn_norm). */

/* deg[i] = #(index == i); deg has idx_dtype. */
int b200mp_degree(const void* index, int64_t n_index, int64_t n_nodes, void* deg, int idx_dtype,
                  void* stream);
/* ptr[i] = #(index < i) for a SORTED index; ptr has n_nodes + 1 entries. */
int b200mp_index2ptr(const void* index_sorted, int64_t n_index, int64_t n_nodes, void* ptr,
                     int idx_dtype, void* stream);
/* index[e] = i for ptr[i] <= e < ptr[i+1]. */
int b200mp_ptr2index(const void* ptr, int64_t n_nodes, int64_t n_index, void* index,
                     int idx_dtype, void* stream);
/* min / max / sortedness of an index array in one pass: out_host-less, writes 3 int64 to DEVICE
 * memory stats[3] = {min, max, is_sorted(0/1)} (n_index == 0 -> {0, -1, 1}). */
int b200mp_index_stats(const void* index, int64_t n_index, int64_t* stats, int idx_dtype,
                       void* stream);
/* Stable sort of keys in [0, n_nodes): writes keys_sorted (optional, may be NULL), perm (the
 * stable argsort, idx_dtype) and ptr (optional, n_nodes+1, the CSR pointer of the sorted keys).
 * Workspace size from b200mp_sort_workspace_bytes(). */
int64_t b200mp_sort_workspace_bytes(int64_t n_index, int64_t n_nodes, int idx_dtype);
int b200mp_sort_by_key(const void* keys, int64_t n_index, int64_t n_nodes, void* keys_sorted,
                       void* perm, void* ptr, void* workspace, int64_t workspace_bytes,
                       int idx_dtype, void* stream);
/* out[i] = in[perm[i]] for 4- or 8-byte elements (elem_bytes), perm has idx_dtype. */
int b200mp_permute(const void* in, const void* perm, void* out, int64_t n, int elem_bytes,
                   int idx_dtype, void* stream);
/* Narrowing / widening copy between index dtypes (int64 <-> int32). */
int b200mp_convert_index(const void* in, int in_dtype, void* out, int out_dtype, int64_t n,
                         void* stream);
/* Self-loop handling.  Output order is the reference's: all non-loop edges in input order, then
 * (i,i) for i in [0, n_nodes).  w_in/w_out may be NULL.  mode 0 = add_remaining_self_loops
 * (existing loop weights override fill_value; duplicate loops: the LAST in input order wins,
 * which is the reference's CPU behaviour), mode 1 = remove_self_loops + add_self_loops(fill).
 * n_out_dev (DEVICE int64) receives E' = #nonloops + n_nodes.  row_out/col_out/w_out need
 * n_edges + n_nodes entries.  Workspace from b200mp_self_loops_workspace_bytes(). */
int64_t b200mp_self_loops_workspace_bytes(int64_t n_edges, int64_t n_nodes, int idx_dtype);
int b200mp_self_loops(const void* row, const void* col, const float* w_in, int64_t n_edges,
                      int64_t n_nodes, float fill_value, int mode, void* row_out, void* col_out,
                      float* w_out, int64_t* n_out_dev, void* workspace, int64_t workspace_bytes,
                      int idx_dtype, void* stream);
/* gcn_norm weights on a destination-sorted (CSR over dst) edge list:
 *   deg[i]  = sum of w over the CSR row i, in order (bit-identical to the reference's CPU
 *             scatter_add_, which visits edges in input order, because the sort is stable);
 *   dinv    = deg^-0.5 with inf -> 0;   w_out[e] = dinv[src[e]] * w[e] * dinv[dst(e)].
 * w may be NULL (all ones).  deg_inv_sqrt (n_nodes floats) is an output too. */
int b200mp_gcn_norm_csr(const void* rowptr, const void* src, const float* w, int64_t n_nodes,
                        int64_t n_edges, float* deg_inv_sqrt, float* w_out, int idx_dtype,
                        void* stream);

/* Long-row plan: rows with more than `chunk` edges are split into chunks of `chunk` edges so no
 * warp ever walks a power-law hub alone.  Two calls: count (writes counts_dev[2] =
 * {n_long_rows, n_chunks} to DEVICE memory; the caller reads them back once per graph), then
 * fill: long_rows[n_long_rows] (ascending row ids) and chunk_ptr[n_long_rows + 1] (exclusive scan
 * of ceil(deg / chunk)), both int64.  Workspace from b200mp_csr_plan_workspace_bytes(). */
int b200mp_csr_plan_count(const void* rowptr, int64_t n_rows, int64_t chunk, int64_t* counts_dev,
                          int idx_dtype, void* stream);
int64_t b200mp_csr_plan_workspace_bytes(int64_t n_rows, int64_t n_long_rows, int idx_dtype);
int b200mp_csr_plan_fill(const void* rowptr, int64_t n_rows, int64_t chunk, int64_t n_long_rows,
                         int64_t* long_rows, int64_t* chunk_ptr, void* workspace,
                         int64_t workspace_bytes, int idx_dtype, void* stream);

/* ------------------------------------------------------------------ gather + segmented reduce (the hot path)
 * out[i, :] = REDUCE_{e in [rowptr[i], rowptr[i+1])} val[e] * x[col[e], :]
 * Replaces: MessagePassing._collect/_lift index_select + message + aggregate
 * (nn/conv/message_passing.py:263-333,577-595; collect.jinja:118-139; aggr/base.py:173-185),
 * utils/_spmm.py:12-136, EdgeIndex.matmul (edge_index.py:1903-1970),
 * torch.ops.torch_sparse.spmm_{sum,mean,min,max} (edge_index.py:1798-1810).
 * Semantics are the reference's: empty rows -> 0 for every reduce; mean divides by
 * max(deg, 1); val (fp32, nullable) multiplies before the reduction; the weighted product is
 * rounded before the add (no FMA contraction) so that rows walked by a single lane group in CSR
 * order reproduce the reference's CPU results bit for bit.
 * x: [n_cols, feat], out: [n_rows, feat], both val_dtype.
 * Long rows: pass the plan (long_rows, chunk_ptr, n_long_rows, n_chunks, chunk) and a partials
 * workspace of n_chunks * feat fp32, or n_long_rows = 0 to walk every row with one lane group.
 * bias (nullable, feat fp32) is added to every output row in the epilogue (GCNConv's `out + bias`,
 * gcn_conv.py:265-266, without a second pass over [N, F]).
 * x_halo (nullable): second source segment for node-sharded runs -- column ids >= n_local_cols
 * are read from x_halo[c - n_local_cols, :] (the rows received by the halo all_to_all) so local
 * and remote rows are never concatenated (n_cols = n_local_cols + #halo rows).
 * flags bit 0 (accumulate, sum only, no bias): out[i,:] += result for rows that have edges and
 * rows without edges are left untouched -- adds the halo-edge part after the local-edge sweep.
 * peer_ptrs (nullable, DEVICE array of uint64 addresses, one per GPU) + peer_rows: the source
 * matrix is sharded by contiguous row ranges of peer_rows rows over the GPUs of the box and
 * peer_ptrs[r] is rank r's peer-mapped base address (torch symmetric memory / CUDA IPC); column c
 * is then gathered from peer_ptrs[c / peer_rows] + (c % peer_rows) * row_bytes, i.e. remote rows
 * come straight over NVLink inside this kernel -- the collective is fused into the gather.
 * relu_mask (nullable, with flags bit 0): a [n_rows, feat] matrix of val_dtype; after the accumulate, out[i,f] is
 * zeroed where relu_mask[i,f] <= 0 (rows without edges too) -- the ReLU backward of the producing layer fused into the
 * last writer of its gradient (a layer's input x = relu(pre) is its own mask). */
int b200mp_spmm_csr(const void* rowptr, const void* col, const float* val, const void* x,
                    void* out, int64_t n_rows, int64_t n_cols, int64_t feat, int reduce,
                    const int64_t* long_rows, const int64_t* chunk_ptr, int64_t n_long_rows,
                    int64_t n_chunks, int64_t chunk, float* partials, const float* bias,
                    const void* x_halo, int64_t n_local_cols, int flags, const void* peer_ptrs,
                    int64_t peer_rows, const void* relu_mask, int idx_dtype, int val_dtype, void* stream);

/* Segmented reduce without gather: out[i,:] = REDUCE_{e in [ptr[i], ptr[i+1])} src[e,:].
 * Replaces utils/_segment.py:11-50 (torch._segment_reduce / torch_scatter.segment_csr) and the
 * sorted-index case of utils/_scatter.py:14-138.  Same empty-segment and +-inf -> 0 rules.
 * Long segments (hub destinations): optional long-row plan as in b200mp_spmm_csr (n_long_rows = 0: none). */
int b200mp_segment_csr(const void* ptr, const void* src, void* out, int64_t n_rows, int64_t n_src,
                       int64_t feat, int reduce, const int64_t* long_rows, const int64_t* chunk_ptr,
                       int64_t n_long_rows, int64_t n_chunks, int64_t chunk, float* partials,
                       int idx_dtype, int val_dtype, void* stream);

/* Backward of min/max aggregation (ATen scatter_reduce rule: the gradient is split evenly among
 * tied extrema, and the zero-initialised output counts as one more tie when the extremum is
 * exactly 0 -- see oracle/mp_oracle.c oracle_scatter_backward).  Works on the TRANSPOSED CSR
 * (rows = source nodes, colT[e] = destination of that edge, valT = its weight):
 *   grad_x[j,:] = sum_{e in rowT(j)} [valT[e]*x[j,:] == out[colT[e],:]] * valT[e] * g[colT[e],:] / ties[colT[e],:]
 * ties [n_dst, feat] fp32 is produced by b200mp_minmax_ties on the forward CSR. */
int b200mp_minmax_ties(const void* rowptr, const void* col, const float* val, const void* x,
                       const void* out, float* ties, int64_t n_rows, int64_t feat,
                       int count_self_zero, int idx_dtype, int val_dtype, void* stream);
int b200mp_minmax_backward(const void* rowptr_t, const void* col_t, const float* val_t,
                           const void* x, const void* out, const void* grad_out, const float* ties,
                           void* grad_x, int64_t n_src, int64_t feat, int idx_dtype,
                           int val_dtype, void* stream);

/* Edge-wise dot product (SDDMM): dot[e] = sum_f a[row_of(e), f] * b[col[e], f] over a CSR.
 * This is the gradient of b200mp_spmm_csr(sum) wrt val: a = grad_out, b = x.
 * Replaces the value-gradient branch of _scatter_spmm (edge_index.py:1950-1953). */
int b200mp_sddmm_csr(const void* rowptr, const void* col, const void* a, const void* b, float* dot,
                     int64_t n_rows, int64_t feat, int idx_dtype, int val_dtype, void* stream);

/* ------------------------------------------------------------------ COO scatter fallback (atomics)
 * out[index[e], :] (+)= src[e, :] for an UNSORTED index.  Replaces utils/_scatter.py:14-138
 * (aten::scatter_add_ / scatter_reduce_, torch_scatter.scatter).  fp32 only.  `count` is a
 * caller-provided n_rows fp32 scratch (used by mean/min/max to detect empty rows).  out is fully
 * written (initialised inside).  sum uses red.global.add.v4.f32 where feat % 4 == 0. */
int b200mp_scatter_coo(const float* src, const void* index, float* out, float* count,
                       int64_t n_src, int64_t n_rows, int64_t feat, int reduce, int idx_dtype,
                       void* stream);
/* out[index[e], :] += src[e, :] into an EXISTING fp32 out (no initialisation; atomics): the
 * return leg of the multi-GPU halo exchange (aten::index_add_). */
int b200mp_index_add_rows(const float* src, const void* index, float* out, int64_t n_src,
                          int64_t feat, int idx_dtype, void* stream);
/* Gather rows: out[e,:] = x[index[e],:]  (aten::index_select, message_passing.py:263-290) --
 * only used by the unfused compatibility path and by backward of scatter. scale (nullable, one
 * fp32 per row of out) multiplies each gathered row. */
int b200mp_gather_rows(const void* x, const void* index, const float* scale, void* out,
                       int64_t n_out, int64_t feat, int idx_dtype, int val_dtype, void* stream);

/* ------------------------------------------------------------------ segment softmax
 * out[e,h] = exp(src[e,h] - max_g) / (sum_g exp(src - max_g) + 1e-16) over CSR groups.
 * Replaces utils/_softmax.py:12-92 and pyg_lib.ops.softmax_csr.  fp32.  backward:
 * grad_src = out * (grad_out - sum_g(grad_out * out)). */
int b200mp_softmax_csr(const void* ptr, const float* src, float* out, int64_t n_rows,
                       int64_t n_src, int64_t heads, int idx_dtype, void* stream);
int b200mp_softmax_csr_backward(const void* ptr, const float* out, const float* grad_out,
                                float* grad_src, int64_t n_rows, int64_t n_src, int64_t heads,
                                int idx_dtype, void* stream);
/* Edge-parallel pieces of the hub-safe segment softmax (groups longer than the long-row chunk): with
 * d = dst_of_edge[e], row [n_rows, heads] fp32,
 *   op 0: out = exp(a - row[d])   op 1: out = a / (row[d] + 1e-16)   op 2: out = a * b   op 3: out = a * (b - row[d]).
 * The host composes _softmax.py:82-88 as segment-max, op 0, segment-sum, op 1 (backward: op 2, segment-sum,
 * op 3) with b200mp_segment_csr and its long-row plan, so no group is ever walked by a single lane group. */
int b200mp_softmax_edge_op(int op, const float* a, const float* b, const float* row, const void* dst_of_edge,
                           float* out, int64_t n_src, int64_t heads, int idx_dtype, void* stream);


/* ------------------------------------------------------------------ bias gradient
 * out[f] = sum_i x[i, f] (fp32): the gradient of the layer bias that autograd derives for `out + bias`
 * (nn/conv/gcn_conv.py:263-264), as one deterministic two-launch column sum instead of ATen's generic
 * reduction.  partials: caller-owned fp32 workspace [n_parts, feat], n_parts from b200mp_column_sum_parts. */
int64_t b200mp_column_sum_parts(int64_t n_rows);
int b200mp_column_sum(const void* x, float* out, float* partials, int64_t n_parts, int64_t n_rows, int64_t feat,
                      int val_dtype, void* stream);

/* ------------------------------------------------------------------ multi-aggregation (one sweep, k outputs)
 * Replaces FusedAggregation.forward (nn/aggr/fused.py:191-336: one scatter per base reduction plus the
 * shared count) and the [sum, mean, min, max, var, std] members of MultiAggregation (nn/aggr/multi.py):
 * every requested output of row i is derived from ONE walk over rowptr[i]:rowptr[i+1].
 *   col != NULL: gather mode, row e reads x[col[e], :]  (x: [n_src, feat]);
 *   col == NULL: segment mode, x is the destination-sorted message matrix [n_src = E, feat].
 * out_* are nullable [n_rows, feat] tensors of val_dtype (NULL = not requested); ties_min / ties_max
 * (nullable, fp32 [n_rows, feat]) receive the number of edges attaining the extremum, plus one where the
 * extremum is 0 if count_self_zero (ATen's scatter_reduce backward rule) -- what the backward divides by.
 * mean = sum / max(deg,1); var = sumsq / max(deg,1) - mean^2; std = sqrt(max(var,1e-5)), 0 where that is
 * <= sqrt(1e-5) (fused.py:319-323); empty rows give 0 everywhere.  Hub rows: long-row plan of
 * b200mp_csr_plan_* with partials of n_chunks * 6 * feat fp32.
 * hit_mask (nullable; gather mode, fp32, feat % 4 == 0, with a ties plane): uint8 [n_edges, feat / 4] in CSR edge
 * order, bit i of byte (e, v) = x[col[e], 4v + i] equals the row's min, bit 4 + i = equals the row's max -- the
 * comparison the backward needs, taken while the row's sources are still in L2, so that the backward reads one byte
 * per edge and vector instead of the destination's min and max vectors (ask b200mp_multi_aggr_mask_supported). */
int b200mp_multi_aggr_mask_supported(int64_t feat, int val_dtype, int segment_mode);
int b200mp_multi_aggr_csr(const void* rowptr, const void* col, const void* x, void* out_sum, void* out_mean,
                          void* out_min, void* out_max, void* out_var, void* out_std, float* ties_min,
                          float* ties_max, void* hit_mask, int64_t n_rows, int64_t n_src, int64_t feat, int count_self_zero,
                          const int64_t* long_rows, const int64_t* chunk_ptr, int64_t n_long_rows,
                          int64_t n_chunks, int64_t chunk, float* partials, int idx_dtype, int val_dtype,
                          void* stream);
/* Backward of the above in one kernel.  The caller folds the output gradients into per-destination fp32
 * rows (all nullable, [n_dst, feat]):  term_a (added), term_b (multiplies the message value),
 * g_min = grad_min / ties_min with out_min (val_dtype), g_max / out_max likewise; then
 *   grad(x_e) = sum over the destinations d of the value:  term_a[d] + x * term_b[d]
 *               + [x == out_min[d]] * g_min[d] + [x == out_max[d]] * g_max[d].
 * segment_mode != 0: n_items = E messages, idx = dst_of_edge [E] (ptr unused), x / grad_x: [E, feat];
 * segment_mode == 0: n_items = n_src source rows, (ptr, idx) = transposed CSR (rowptr_t, col_t).
 * hit_mask + t2csr (both nullable, gather mode only): the forward's hit bits and the CSR slot of every transposed
 * slot; when given and supported they replace the reads of out_min / out_max (which must still be passed: shapes the
 * masked kernel does not cover compare against them). */
/* Elementwise prologue of the backward: folds the output gradients (nullable, [n_rows, feat] val_dtype) and
 * the saved mean / std / tie counts into term_a, term_b, g_min / ties_min, g_max / ties_max (fp32).
 * cnt = max(rowptr[i+1] - rowptr[i], 1); semi_grad drops the 2 x / cnt term of var / std (basic.py:106-110). */
int b200mp_multi_aggr_prepare_backward(const void* rowptr, const void* g_sum, const void* g_mean,
                                       const void* g_var, const void* g_std, const void* g_min,
                                       const void* g_max, const void* mean, const void* std,
                                       const float* ties_min, const float* ties_max, float* term_a,
                                       float* term_b, float* gmin_out, float* gmax_out, int64_t n_rows,
                                       int64_t feat, int semi_grad, int idx_dtype, int val_dtype, void* stream);
int b200mp_multi_aggr_backward(const void* ptr, const void* idx, const void* x, const float* term_a,
                               const float* term_b, const void* out_min, const float* g_min,
                               const void* out_max, const float* g_max, const void* hit_mask, const void* t2csr,
                               void* grad_x, int64_t n_items, int64_t feat, int segment_mode, int idx_dtype,
                               int val_dtype, void* stream);

/* ------------------------------------------------------------------ node-level attention terms
 * s_a[n,h] = sum_c x[n,h,c] * att_a[h,c] (and s_b with att_b from the same read of x; att_b / s_b nullable together):
 * GATConv's alpha_src / alpha_dst = (x * att).sum(-1) (nn/conv/gat_conv.py:330-331) without the [N,H,C] product tensor.
 * x: [n_rows, heads*chan] of val_dtype (fp32 / bf16), att_*: fp32 [heads*chan], s_*: fp32 [n_rows, heads].
 * Supported when chan * sizeof(val) is a multiple of 16 and a power-of-two number (<= 32) of 16-byte vectors, and a row has
 * at most 256 vectors (b200mp_head_dot_supported); other shapes are the caller's business.
 * Backward in one pass: grad_x = g_a (x) att_a + g_b (x) att_b (+ add, nullable: e.g. the attention's own grad_v), and
 * per-CTA partial rows part_* [n_parts, heads*chan] fp32 of grad_att_* = sum_n g_*[n,h] * x[n,h,c] (fold them with
 * b200mp_column_sum; n_parts from b200mp_head_dot_parts).  grad_x nullable (x does not need a gradient). */
int b200mp_head_dot_supported(int64_t heads, int64_t chan, int val_dtype);
int64_t b200mp_head_dot_parts(int64_t n_rows, int64_t heads, int64_t chan, int val_dtype);
int b200mp_head_dot(const void* x, const float* att_a, const float* att_b, float* s_a, float* s_b, int64_t n_rows,
                    int64_t heads, int64_t chan, int val_dtype, void* stream);
int b200mp_head_dot_backward(const void* x, const float* att_a, const float* att_b, const float* g_a, const float* g_b,
                             const void* add, void* grad_x, float* part_a, float* part_b, int64_t n_parts,
                             int64_t n_rows, int64_t heads, int64_t chan, int val_dtype, void* stream);

/* ------------------------------------------------------------------ fused GAT attention + aggregation
 * One sweep over the destination-sorted CSR per (node, head):
 *   logit_e = leaky_relu(a_src[col[e],h] + a_dst[i,h], slope)
 *   alpha_e = softmax over the row;  out[i,h,:] = sum_e alpha_e * xh[col[e],h,:]
 * Replaces GATConv.edge_update + message + aggregate (nn/conv/gat_conv.py:387-409,
 * utils/_softmax.py:82-88) -- 12 kernels and three [E,H,C] tensors in the reference.
 * xh: [n_src, heads*chan] val_dtype; a_src [n_src, heads], a_dst [n_rows, heads] fp32;
 * out: [n_rows, heads*chan] val_dtype; row_max/row_den [n_rows, heads] fp32 (saved for backward);
 * alpha_out (nullable) [n_edges, heads] fp32 in CSR order. */
int b200mp_gat_fused_csr(const void* rowptr, const void* col, const void* dst_of_edge, const void* xh,
                         const float* a_src, const float* a_dst, void* out, float* row_max,
                         float* row_den, float* alpha_out, int64_t n_rows, int64_t n_edges,
                         int64_t heads, int64_t chan, float slope, const int64_t* long_rows,
                         const int64_t* chunk_ptr, int64_t n_long_rows, int64_t n_chunks, int64_t chunk,
                         float* part_acc, float* part_ms, int idx_dtype, int val_dtype, void* stream);
/* (dst_of_edge = ptr2index(rowptr), only needed with alpha_out.  Hub rows: pass the long-row plan of
 * b200mp_csr_plan_* plus part_acc [n_chunks, heads*chan] and part_ms [n_chunks, heads, 2] fp32; every
 * chunk keeps an online-softmax state that is merged with exp(m_c - M) rescaling.)
 *
 * Backward of b200mp_gat_fused_csr, attention recomputed from row_max/row_den:
 *  row dots    rowdot[i,h] = <g[i,h,:], out[i,h,:]>               (scratch [n_rows, heads])
 *  edge sweep  grad_pre[e,h] (CSR order, scratch [n_edges, heads]) -- one thread per (edge, head), so
 *              no row is walked serially -- and grad_a_dst[i,h] = its (chunked) segmented sum
 *              (partials: n_chunks * heads fp32 when the plan is given)
 *  source sweep (rowptr_t/col_t, t2csr[e] = CSR slot of transposed slot e): grad_xh[j,h,:] (the
 *              message term only; the a_src/a_dst terms flow back through the caller's
 *              (xh*att).sum(-1)) and grad_a_src[j,h]. */
int b200mp_gat_fused_csr_backward(const void* rowptr, const void* col, const void* dst_of_edge,
                                  const void* rowptr_t, const void* col_t, const void* t2csr,
                                  const void* xh, const float* a_src, const float* a_dst,
                                  const float* row_max, const float* row_den, const void* out,
                                  const void* grad_out, float* grad_pre, float* rowdot, void* grad_xh,
                                  float* grad_a_src, float* grad_a_dst, int64_t n_rows, int64_t n_src,
                                  int64_t n_edges, int64_t heads, int64_t chan, float slope,
                                  const int64_t* long_rows, const int64_t* chunk_ptr,
                                  int64_t n_long_rows, int64_t n_chunks, int64_t chunk, float* partials,
                                  int idx_dtype, int val_dtype, void* stream);

/* ------------------------------------------------------------------ argmin / argmax outputs (csrc/arg.cu)
 * The (out, arg) operator signatures the reference binds to: torch_scatter.scatter_max / scatter_min
 * (utils/_scatter.py:147-156) and torch.ops.torch_sparse.spmm_min / spmm_max (edge_index.py:1798-1810).
 * `out` is what b200mp_scatter_coo / b200mp_spmm_csr produced (fp32, min or max); these passes find its producer:
 *   b200mp_scatter_arg   arg[i,f] = smallest e with index[e] == i and src[e,f] == out[i,f]; n_src for empty groups
 *   b200mp_spmm_csr_arg  arg[i,f] = first CSR slot e of row i with val[e]*x[col[e],f] == out[i,f]; nnz for empty rows
 * arg: int64 [n_rows, feat]. */
int b200mp_scatter_arg(const float* src, const void* index, const float* out, int64_t* arg, int64_t n_src,
                       int64_t n_rows, int64_t feat, int idx_dtype, void* stream);
int b200mp_spmm_csr_arg(const void* rowptr, const void* col, const float* val, const float* x, const float* out,
                        int64_t* arg, int64_t n_rows, int64_t feat, int64_t nnz, int idx_dtype, void* stream);

/* ------------------------------------------------------------------ fused attention family (csrc/attention.cu)
 * One kernel skeleton for the three score functions of the reference's attention convolutions, forward and
 * backward, over the destination-sorted CSR (alpha = softmax over the in-edges of i, utils/_softmax.py:82-88;
 * out[i,h,:] = sum_e alpha_e,h v[col[e],h,:]):
 *   mode 0 GAT    s = leaky_relu(s_src[j,h] + s_dst[i,h] (+ s_edge[e,h]))       nn/conv/gat_conv.py:387-409
 *   mode 1 GATv2  s = sum_c att[h,c] leaky_relu(v[j,h,c] + q[i,h,c])           nn/conv/gatv2_conv.py:358-378
 *   mode 2 DOT    s = scale <q[i,h,:], k[j,h,:]>                               nn/conv/transformer_conv.py:263-275
 * v (GAT xh / GATv2 x_l / value rows), k (DOT keys): [n_src, heads*chan] val_dtype with row strides v_stride /
 * k_stride in ELEMENTS (0 = heads*chan; k and v may be the halves of one [N, 2*H*C] product); q (GATv2 x_r / DOT
 * queries): [n_rows, heads*chan], stride q_stride; s_src [n_src, heads], s_dst [n_rows, heads], att [heads*chan],
 * s_edge [n_edges, heads] (CSR order, nullable) fp32.  out [n_rows, heads*chan]; row_max / row_den [n_rows, heads]
 * fp32 are saved for the backward; alpha_out (nullable) [n_edges, heads] fp32 in CSR order.
 * dropout_p in [0, 1): attention dropout (gat_conv.py:404, gatv2_conv.py:376, transformer_conv.py:268 -- F.dropout on the
 * normalised coefficients): (edge, head) pairs are dropped by a counter-based hash of (dropout_seed, CSR slot, head) that the
 * forward and the backward evaluate identically (pass the same p and seed to both); kept coefficients are scaled by
 * 1 / (1 - p); alpha_out returns the dropped coefficients like the reference.  0 = no dropout.
 * edge_feat (nullable; GATv2 and dot modes): per-edge feature rows [n_edges, heads*chan] of val_dtype in CSR order, i.e.
 * lin_edge(edge_attr) of `edge_dim` layers: GATv2 adds them inside the leaky_relu (gatv2_conv.py:358-360), the dot mode to
 * the key and to the value (transformer_conv.py:258-272).  The backward writes grad_edge_feat (same shape; required when
 * edge_feat is given).  (GAT's edge_dim term is the scalar s_edge.)
 * Shapes: b200mp_attn_supported(heads, chan, val_dtype) != 0 (rows of whole 16-byte vectors, <= 1 KB, a head =
 * a power-of-two number of vectors), else B200MP_ERR_UNSUPPORTED.  Hub rows: the long-row plan of
 * b200mp_csr_plan_* plus part_acc [n_chunks, heads*chan] and part_ms [n_chunks, heads, 2] fp32.
 *
 * Backward: destination sweep (softmax backward per edge, grad_q / grad_s_dst / grad_att, and the per-edge scratch
 * pair [n_edges, heads, 2] fp32 = (alpha, grad_score) in CSR order -- for GAT pair[...,1] is also the gradient of
 * s_edge) then source sweep on the transposed CSR (t2csr[e] = CSR slot of transposed slot e): grad_v, grad_k (DOT),
 * grad_s_src (GAT).  Long-row partials: partials [n_chunks, b200mp_attn_backward_partial_width(mode,H,C,0)] and
 * partials_t [n_chunks_t, ...(mode,H,C,1)] fp32; gatt_part (GATv2) [b200mp_attn_gatt_rows(), heads*chan] fp32. */
int b200mp_attn_supported(int64_t heads, int64_t chan, int val_dtype);
int b200mp_attn_csr_forward(int mode, const void* rowptr, const void* col, const void* v, const void* k,
                            const void* q, const float* s_src, const float* s_dst, const float* att,
                            const float* s_edge, int64_t v_stride, int64_t k_stride, int64_t q_stride,
                            void* out, float* row_max, float* row_den, float* alpha_out, int64_t n_rows,
                            int64_t n_edges, int64_t heads, int64_t chan, float slope, float scale,
                            const int64_t* long_rows, const int64_t* chunk_ptr, int64_t n_long_rows,
                            int64_t n_chunks, int64_t chunk, float* part_acc, float* part_ms, float dropout_p,
                            unsigned long long dropout_seed, const void* edge_feat, int idx_dtype, int val_dtype,
                            void* stream);
int64_t b200mp_attn_backward_partial_width(int mode, int64_t heads, int64_t chan, int transposed);
int64_t b200mp_attn_gatt_rows(void);
int b200mp_attn_csr_backward(int mode, const void* rowptr, const void* col, const void* rowptr_t,
                             const void* col_t, const void* t2csr, const void* v, const void* k, const void* q,
                             const float* s_src, const float* s_dst, const float* att, const float* s_edge,
                             int64_t v_stride, int64_t k_stride, int64_t q_stride, const float* row_max,
                             const float* row_den, const void* out, const void* grad_out, float* pair,
                             void* grad_v, void* grad_k, void* grad_q, float* grad_s_src, float* grad_s_dst,
                             float* grad_att, float* gatt_part, int64_t n_rows, int64_t n_src, int64_t n_edges,
                             int64_t heads, int64_t chan, float slope, float scale, const int64_t* long_rows,
                             const int64_t* chunk_ptr, int64_t n_long_rows, int64_t n_chunks, int64_t chunk,
                             float* partials, const int64_t* long_rows_t, const int64_t* chunk_ptr_t,
                             int64_t n_long_rows_t, int64_t n_chunks_t, float* partials_t, float dropout_p,
                             unsigned long long dropout_seed, const void* edge_feat, void* grad_edge_feat, int idx_dtype,
                             int val_dtype, void* stream);

/* ------------------------------------------------------------------ dense transform on tensor cores
 * fp32-accurate 3xTF32 GEMMs (tcgen05 + TMEM + TMA, csrc/gemm_tf32x3.cu) for the layer's
 * Linear (nn/dense/linear.py:121-127: F.linear, run by the reference as strict-fp32 cuBLAS):
 *   b200mp_linear_tf32x3            y [M,N]  = x [M,K] . w[N,K]^T
 *   b200mp_linear_grad_input_tf32x3 gx[M,K]  = g [M,N] . w[N,K]
 *   b200mp_linear_grad_weight_tf32x3 gw[N,K] = g [M,N]^T . x[M,K]   (deterministic split-K)
 * w_hi/w_lo come from b200mp_split_tf32 (w = w_hi + w_lo, w_hi = rn_tf32(w)); w_lo == NULL means w_hi is
 * the UNSPLIT weight and the kernel splits each B tile in shared memory (one L2 read of W per tile instead
 * of two; needs an output width that is a multiple of 128).  All matrices
 * row-major, contiguous, 16-byte aligned.  Shape limits (else B200MP_ERR_UNSUPPORTED and the caller
 * uses a library GEMM): reduction dim % 32 == 0, output width in {64, 128} or a multiple of 256,
 * and for grad_weight N % 128 == 0. */
int b200mp_split_tf32(const float* w, float* w_hi, float* w_lo, int64_t n, void* stream);
int b200mp_linear_tf32x3(const float* x, const float* w_hi, const float* w_lo, float* y, int64_t m,
                         int64_t n, int64_t k, void* stream);
int b200mp_linear_grad_input_tf32x3(const float* g, const float* w_hi, const float* w_lo, float* gx,
                                    int64_t m, int64_t n, int64_t k, void* stream);
int64_t b200mp_linear_grad_weight_workspace_bytes(int64_t m, int64_t n, int64_t k);
int b200mp_linear_grad_weight_tf32x3(const float* g, const float* x, float* gw, int64_t m, int64_t n,
                                     int64_t k, void* workspace, int64_t workspace_bytes,
                                     void* stream);

/* Pair form of the TS-mode kernel (A operand in tensor memory): two A streams accumulate into ONE TMEM accumulator,
 * the epilogue adds a bias and applies ReLU, and the output columns may be split over two matrices:
 *     [c1 | c2] [M, n1+n2] = act( [a1 | a2] [M, k1+k2] . B + bias )
 * b_layout 0: B = w [n1+n2, k1+k2] row-major (y = A w^T: SAGEConv's lin_l(agg) + lin_r(x) with w = [W_l | W_r],
 * sage_conv.py:134-141, in one launch instead of two GEMMs and an add); b_layout 1: B = w [k1+k2, n1+n2] row-major
 * (y = A w: both input gradients of such a pair from one read of g, RGCNConv's [H | x] . [W_1;..;W_R; root],
 * rgcn_conv.py:257-280).  a2 / c2 / bias nullable (k2 = 0 / n2 = 0).  k1, k2 % 32 == 0; n1, n2 % 128 == 0;
 * b_lo == NULL: b_hi is the unsplit matrix (the kernel splits B tiles itself). */
int b200mp_gemm_pair_tf32x3(const float* a1, int64_t k1, const float* a2, int64_t k2, const float* b_hi,
                            const float* b_lo, int b_layout, const float* bias, int relu, float* c1, int64_t n1,
                            float* c2, int64_t n2, int64_t m, void* stream);

/* Grouped form: pyg_lib.ops.segment_matmul(inputs, ptr, other) (nn/dense/linear.py:248-255, nn/conv/rgcn_conv.py:288;
 * HeteroLinear, HeteroDictLinear, RGCNConv's sorted-by-type path).
 *     c[ptr[r] : ptr[r+1], :] = a[ptr[r] : ptr[r+1], :] . B_r        for r in [0, n_seg)
 * in ONE persistent launch over (segment, 128-row tile, 128-column tile) work items; ptr [n_seg+1] int64 stays on the
 * device (the tile -> segment map is rebuilt in shared memory), partial tiles at segment ends are stored row-masked.
 * b_layout 1: B_r = w[r] of a [n_seg, K, N] weight (c = a w[r]); b_layout 0: B_r = w[r] of a [n_seg, N, K] weight
 * (c = a w[r]^T: the input gradient of layout 1).  b_hi / b_lo from b200mp_split_tf32 over the whole stack.
 * k % 32 == 0, n % 128 == 0, n_seg <= 120. */
int b200mp_segment_matmul_tf32x3(const float* a, const int64_t* ptr, int64_t n_seg, const float* b_hi,
                                 const float* b_lo, int b_layout, float* c, int64_t m, int64_t k, int64_t n,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MP_H_ */
