/*
 * mp_oracle.c -- CPU restatement of the reference's message-passing aggregation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pytorch_geometric_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg use it, and only as the checker / the timed CPU arm.
 *
 * The reference (pyg-team/pytorch_geometric v2.9.0, /root/reference) is pure Python over
 * ATen.  Each function below restates one reference function in plain C (serial, fp32
 * arithmetic in the same order the reference's CPU path uses: edges visited in input
 * order) and cites the reference file:line it follows.  Parity is PINNED: the restatement
 * is checked against golden vectors produced by importing the reference itself in the
 * build container (tests/golden/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py)
 * and against the reference's own known-answer tests (test/utils/test_softmax.py:12-27,
 * test_degree.py, test_loop.py:220-291, test_segment.py:13-31, test_scatter.py:111-120).
 *
 * Conventions: row-major fp32 feature matrices, int64 indices (the reference's dtype),
 * reduce codes shared with include/b200mp.h: 0 sum, 1 mean, 2 min, 3 max, 4 mul.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { R_SUM = 0, R_MEAN = 1, R_MIN = 2, R_MAX = 3, R_MUL = 4 };

/* ---------------------------------------------------------------- integer work */

/* torch_geometric/utils/_degree.py:9-31 -- degree = scatter_add of ones over index. */
void oracle_degree(const int64_t* index, int64_t E, int64_t N, int64_t* deg) {
    memset(deg, 0, sizeof(int64_t) * (size_t)N);
    for (int64_t e = 0; e < E; ++e) deg[index[e]] += 1;
}

/* torch_geometric/index.py:32-37 -- index2ptr: torch._convert_indices_from_coo_to_csr
 * on a SORTED index: ptr[i] = #elements < i, ptr[N] = E. */
void oracle_index2ptr(const int64_t* index, int64_t E, int64_t N, int64_t* ptr) {
    memset(ptr, 0, sizeof(int64_t) * (size_t)(N + 1));
    for (int64_t e = 0; e < E; ++e) ptr[index[e] + 1] += 1;
    for (int64_t i = 0; i < N; ++i) ptr[i + 1] += ptr[i];
}

/* torch_geometric/index.py:27-30 -- ptr2index: repeat_interleave(arange(N), ptr.diff()). */
void oracle_ptr2index(const int64_t* ptr, int64_t N, int64_t* index) {
    for (int64_t i = 0; i < N; ++i)
        for (int64_t e = ptr[i]; e < ptr[i + 1]; ++e) index[e] = i;
}

/* torch_geometric/utils/_index_sort.py:10-32 with stable=True semantics (the reference's
 * tests accept any permutation within equal keys, test/test_edge_index.py:205-208; the
 * stable one is the canonical member of that set).  Counting sort: perm is the stable
 * argsort of keys, ptr the CSR pointer of the sorted keys. */
void oracle_stable_sort_by_key(const int64_t* keys, int64_t E, int64_t N, int64_t* perm,
                               int64_t* ptr) {
    oracle_index2ptr(keys, E, N, ptr); /* histogram + scan does not need sortedness */
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N + 1));
    memcpy(cursor, ptr, sizeof(int64_t) * (size_t)(N + 1));
    for (int64_t e = 0; e < E; ++e) perm[cursor[keys[e]]++] = e;
    free(cursor);
}

/* torch_geometric/utils/loop.py:585-657 -- add_remaining_self_loops.
 * Output order: all non-loop edges in input order, then (i,i) for i in arange(N).
 * Weights: loop_attr = fill_value everywhere (compute_loop_attr, loop.py:742-758), then
 * existing self-loop weights overwrite it in input order (last one wins, loop.py:641-644).
 * w_in may be NULL (then w_out is not touched).  Returns E' = #nonloops + N. */
int64_t oracle_add_remaining_self_loops(const int64_t* row, const int64_t* col, const float* w_in,
                                        int64_t E, int64_t N, float fill_value, int64_t* row_out,
                                        int64_t* col_out, float* w_out) {
    int64_t k = 0;
    for (int64_t e = 0; e < E; ++e) {
        if (row[e] != col[e]) {
            row_out[k] = row[e];
            col_out[k] = col[e];
            if (w_in) w_out[k] = w_in[e];
            ++k;
        }
    }
    for (int64_t i = 0; i < N; ++i) {
        row_out[k + i] = i;
        col_out[k + i] = i;
        if (w_in) w_out[k + i] = fill_value;
    }
    if (w_in)
        for (int64_t e = 0; e < E; ++e)
            if (row[e] == col[e]) w_out[k + row[e]] = w_in[e];
    return k + N;
}

/* torch_geometric/utils/loop.py:71-131 (remove_self_loops) followed by :382-492
 * (add_self_loops, no edge_attr) -- GATConv's preprocessing (gat_conv.py:342-346). */
int64_t oracle_remove_then_add_self_loops(const int64_t* row, const int64_t* col, int64_t E,
                                          int64_t N, int64_t* row_out, int64_t* col_out) {
    int64_t k = 0;
    for (int64_t e = 0; e < E; ++e)
        if (row[e] != col[e]) {
            row_out[k] = row[e];
            col_out[k] = col[e];
            ++k;
        }
    for (int64_t i = 0; i < N; ++i) {
        row_out[k + i] = i;
        col_out[k + i] = i;
    }
    return k + N;
}

/* ---------------------------------------------------------------- scatter / segment */

static inline float red_init(int reduce) {
    switch (reduce) {
        case R_MIN: return INFINITY;
        case R_MAX: return -INFINITY;
        case R_MUL: return 1.0f;
        default: return 0.0f;
    }
}

/* torch_geometric/utils/_scatter.py:60-138 -- scatter(src[E,F], index[E], dim=0, dim_size=N).
 * sum: zeros.scatter_add_ (:68-70).  mean: sum / clamp(count,1) (:72-80).
 * min/max: zeros.scatter_reduce_(amin/amax, include_self=False) -> empty groups stay 0 (:98-100).
 * mul: ones.scatter_reduce_(prod, include_self=True) -> empty groups are 1 (:128-132). */
void oracle_scatter(const float* src, const int64_t* index, int64_t E, int64_t F, int64_t N,
                    int reduce, float* out) {
    const float init = red_init(reduce);
    for (int64_t i = 0; i < N * F; ++i) out[i] = init;
    int64_t* cnt = (int64_t*)calloc((size_t)(N > 0 ? N : 1), sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) {
        const int64_t d = index[e];
        cnt[d] += 1;
        const float* s = src + e * F;
        float* o = out + d * F;
        switch (reduce) {
            case R_SUM:
            case R_MEAN:
                for (int64_t f = 0; f < F; ++f) o[f] += s[f];
                break;
            case R_MIN:
                for (int64_t f = 0; f < F; ++f) o[f] = (s[f] < o[f] || isnan(s[f])) ? s[f] : o[f];
                break;
            case R_MAX:
                for (int64_t f = 0; f < F; ++f) o[f] = (s[f] > o[f] || isnan(s[f])) ? s[f] : o[f];
                break;
            case R_MUL:
                for (int64_t f = 0; f < F; ++f) o[f] *= s[f];
                break;
        }
    }
    for (int64_t d = 0; d < N; ++d) {
        float* o = out + d * F;
        if (reduce == R_MEAN) {
            const float c = (float)(cnt[d] < 1 ? 1 : cnt[d]);
            for (int64_t f = 0; f < F; ++f) o[f] = o[f] / c;
        } else if ((reduce == R_MIN || reduce == R_MAX) && cnt[d] == 0) {
            for (int64_t f = 0; f < F; ++f) o[f] = 0.0f;
        }
    }
    free(cnt);
}

/* Backward of scatter (autograd of the ATen ops the reference calls):
 *  sum  : grad_src[e] = g[index[e]]                       (gather, scatter_add_ backward)
 *  mean : grad_src[e] = g[index[e]] / clamp(count,1)
 *  min/max: scatter_reduce_ backward splits the gradient EVENLY among tied extrema
 *           (verified by running the reference).  ATen quirk kept on purpose:
 *           the tie count is N = (self == result) + #(src == result) where `self` is the
 *           zero-initialised output of _scatter.py:98 -- so when the extremum is exactly 0
 *           (common after ReLU) the gradient is divided by #ties + 1 (golden: scatter_F5 'min').
 *  mul is not differentiated here. */
void oracle_scatter_backward(const float* grad_out, const float* src, const float* out,
                             const int64_t* index, int64_t E, int64_t F, int64_t N, int reduce,
                             float* grad_src) {
    if (reduce == R_SUM || reduce == R_MEAN) {
        int64_t* cnt = (int64_t*)calloc((size_t)(N > 0 ? N : 1), sizeof(int64_t));
        for (int64_t e = 0; e < E; ++e) cnt[index[e]] += 1;
        for (int64_t e = 0; e < E; ++e) {
            const int64_t d = index[e];
            const float c = reduce == R_MEAN ? (float)(cnt[d] < 1 ? 1 : cnt[d]) : 1.0f;
            for (int64_t f = 0; f < F; ++f) grad_src[e * F + f] = grad_out[d * F + f] / c;
        }
        free(cnt);
        return;
    }
    float* ties = (float*)calloc((size_t)(N * F > 0 ? N * F : 1), sizeof(float));
    for (int64_t i = 0; i < N * F; ++i) ties[i] = (out[i] == 0.0f) ? 1.0f : 0.0f;
    for (int64_t e = 0; e < E; ++e) {
        const int64_t d = index[e];
        for (int64_t f = 0; f < F; ++f)
            if (src[e * F + f] == out[d * F + f]) ties[d * F + f] += 1.0f;
    }
    for (int64_t e = 0; e < E; ++e) {
        const int64_t d = index[e];
        for (int64_t f = 0; f < F; ++f)
            grad_src[e * F + f] =
                (src[e * F + f] == out[d * F + f]) ? grad_out[d * F + f] / ties[d * F + f] : 0.0f;
    }
    free(ties);
}

/* torch_geometric/utils/_segment.py:37-50 -- _torch_segment: torch._segment_reduce over
 * ptr ranges; mean uses initial=0; min/max map any +-inf RESULT to 0 (:48-49). */
void oracle_segment(const float* src, const int64_t* ptr, int64_t N, int64_t F, int reduce,
                    float* out) {
    for (int64_t i = 0; i < N; ++i) {
        float* o = out + i * F;
        const int64_t b = ptr[i], e_ = ptr[i + 1];
        const float init = red_init(reduce);
        for (int64_t f = 0; f < F; ++f) o[f] = init;
        for (int64_t e = b; e < e_; ++e) {
            const float* s = src + e * F;
            for (int64_t f = 0; f < F; ++f) {
                switch (reduce) {
                    case R_SUM:
                    case R_MEAN: o[f] += s[f]; break;
                    case R_MIN: o[f] = (s[f] < o[f] || isnan(s[f])) ? s[f] : o[f]; break;
                    case R_MAX: o[f] = (s[f] > o[f] || isnan(s[f])) ? s[f] : o[f]; break;
                    case R_MUL: o[f] *= s[f]; break;
                }
            }
        }
        if (reduce == R_MEAN) {
            const float c = (float)((e_ - b) < 1 ? 1 : (e_ - b));
            for (int64_t f = 0; f < F; ++f) o[f] = (e_ > b) ? o[f] / c : 0.0f;
        }
        if (reduce == R_MIN || reduce == R_MAX)
            for (int64_t f = 0; f < F; ++f)
                if (isinf(o[f])) o[f] = 0.0f;
    }
}

/* torch_geometric/utils/_softmax.py:82-92 (index path; the ptr path :60-81 is the same
 * arithmetic on contiguous groups): out = exp(src - max_g) / (sum_g exp(src - max_g) + 1e-16),
 * max_g from scatter(max) (so an empty group's max is 0, irrelevant as it has no members). */
void oracle_softmax(const float* src, const int64_t* index, int64_t E, int64_t H, int64_t N,
                    float* out) {
    float* mx = (float*)malloc(sizeof(float) * (size_t)(N * H > 0 ? N * H : 1));
    float* sm = (float*)malloc(sizeof(float) * (size_t)(N * H > 0 ? N * H : 1));
    oracle_scatter(src, index, E, H, N, R_MAX, mx);
    for (int64_t e = 0; e < E; ++e)
        for (int64_t h = 0; h < H; ++h) out[e * H + h] = expf(src[e * H + h] - mx[index[e] * H + h]);
    oracle_scatter(out, index, E, H, N, R_SUM, sm);
    for (int64_t e = 0; e < E; ++e)
        for (int64_t h = 0; h < H; ++h) out[e * H + h] = out[e * H + h] / (sm[index[e] * H + h] + 1e-16f);
    free(mx);
    free(sm);
}

/* Backward of softmax through autograd of the ops in _softmax.py:84-92 (max is detached):
 * grad_src = out * (g - sum_g(g * out)) up to the 1e-16 term, which autograd keeps:
 * y = u / (s + eps), u = exp(x - m):  dL/dx_e = y_e * (g_e - sum_{k in g} g_k y_k). */
void oracle_softmax_backward(const float* grad_out, const float* out, const int64_t* index,
                             int64_t E, int64_t H, int64_t N, float* grad_src) {
    float* dot = (float*)calloc((size_t)(N * H > 0 ? N * H : 1), sizeof(float));
    for (int64_t e = 0; e < E; ++e)
        for (int64_t h = 0; h < H; ++h) dot[index[e] * H + h] += grad_out[e * H + h] * out[e * H + h];
    for (int64_t e = 0; e < E; ++e)
        for (int64_t h = 0; h < H; ++h)
            grad_src[e * H + h] = out[e * H + h] * (grad_out[e * H + h] - dot[index[e] * H + h]);
    free(dot);
}

/* ---------------------------------------------------------------- gather + aggregate */

/* The unfused reference path of MessagePassing.propagate for a plain [2,E] edge_index:
 *   x_j = x.index_select(0, src)                (message_passing.py:263-290, collect.jinja:137)
 *   msg = w.view(-1,1) * x_j  (or x_j)          (gcn_conv.py:270-271 / sage_conv.py:146)
 *   out = scatter(msg, dst, 0, N_dst, reduce)   (aggr/base.py:173-185 -> _scatter.py:60-138)
 * which is also torch_geometric/edge_index.py:1903-1922 (_scatter_spmm).
 * x is [N_src, F]; out is [N_dst, F]. */
void oracle_gather_scatter(const float* x, const int64_t* src, const int64_t* dst, const float* w,
                           int64_t E, int64_t F, int64_t N_dst, int reduce, float* out) {
    float* msg = (float*)malloc(sizeof(float) * (size_t)(E * F > 0 ? E * F : 1));
    for (int64_t e = 0; e < E; ++e) {
        const float* xs = x + src[e] * F;
        if (w)
            for (int64_t f = 0; f < F; ++f) msg[e * F + f] = w[e] * xs[f];
        else
            for (int64_t f = 0; f < F; ++f) msg[e * F + f] = xs[f];
    }
    oracle_scatter(msg, dst, E, F, N_dst, reduce, out);
    free(msg);
}

/* Backward of oracle_gather_scatter wrt x (index_select backward = index_add_) and wrt w
 * (mul backward: sum_f x_j * grad_msg).  grad_w may be NULL.  out is the forward result
 * (needed for min/max). */
void oracle_gather_scatter_backward(const float* grad_out, const float* x, const float* out,
                                    const int64_t* src, const int64_t* dst, const float* w,
                                    int64_t E, int64_t F, int64_t N_src, int64_t N_dst, int reduce,
                                    float* grad_x, float* grad_w) {
    float* msg = (float*)malloc(sizeof(float) * (size_t)(E * F > 0 ? E * F : 1));
    float* gmsg = (float*)malloc(sizeof(float) * (size_t)(E * F > 0 ? E * F : 1));
    for (int64_t e = 0; e < E; ++e)
        for (int64_t f = 0; f < F; ++f) msg[e * F + f] = (w ? w[e] : 1.0f) * x[src[e] * F + f];
    oracle_scatter_backward(grad_out, msg, out, dst, E, F, N_dst, reduce, gmsg);
    memset(grad_x, 0, sizeof(float) * (size_t)(N_src * F));
    for (int64_t e = 0; e < E; ++e) {
        float acc = 0.0f;
        for (int64_t f = 0; f < F; ++f) {
            grad_x[src[e] * F + f] += (w ? w[e] : 1.0f) * gmsg[e * F + f];
            acc += x[src[e] * F + f] * gmsg[e * F + f];
        }
        if (grad_w) grad_w[e] = acc;
    }
    free(msg);
    free(gmsg);
}

/* torch_geometric/utils/_spmm.py:12-136 / edge_index.py:1925-1970 -- CSR SpMM with reduce;
 * value-for-value this is oracle_gather_scatter on the CSR-expanded COO, kept separately so
 * CSR inputs (rowptr, col, val) are exercised in the order torch.sparse.mm visits them. */
void oracle_spmm_csr(const int64_t* rowptr, const int64_t* col, const float* val, const float* x,
                     int64_t N, int64_t F, int reduce, float* out) {
    for (int64_t i = 0; i < N; ++i) {
        float* o = out + i * F;
        const int64_t b = rowptr[i], e_ = rowptr[i + 1];
        const float init = (reduce == R_MIN) ? INFINITY : (reduce == R_MAX ? -INFINITY : 0.0f);
        for (int64_t f = 0; f < F; ++f) o[f] = init;
        for (int64_t e = b; e < e_; ++e) {
            const float* xs = x + col[e] * F;
            const float wv = val ? val[e] : 1.0f;
            for (int64_t f = 0; f < F; ++f) {
                const float m = val ? wv * xs[f] : xs[f];
                if (reduce == R_MIN) o[f] = (m < o[f] || isnan(m)) ? m : o[f];
                else if (reduce == R_MAX) o[f] = (m > o[f] || isnan(m)) ? m : o[f];
                else o[f] += m;
            }
        }
        if (reduce == R_MEAN) {
            const float c = (float)((e_ - b) < 1 ? 1 : (e_ - b));
            for (int64_t f = 0; f < F; ++f) o[f] /= c;
        }
        if ((reduce == R_MIN || reduce == R_MAX) && e_ == b)
            for (int64_t f = 0; f < F; ++f) o[f] = 0.0f;
    }
}

/* ---------------------------------------------------------------- GCN */

/* torch_geometric/nn/conv/gcn_conv.py:95-113 -- gcn_norm on a [2,E] tensor,
 * flow=source_to_target: add_remaining_self_loops(fill = improved ? 2 : 1); w = ones if None;
 * deg = scatter(w, col, N, 'sum'); dinv = deg^-0.5 (inf -> 0); w' = dinv[row] * w * dinv[col].
 * Outputs row_out/col_out/w_out sized E + N; returns E'. */
int64_t oracle_gcn_norm(const int64_t* row, const int64_t* col, const float* w_in, int64_t E,
                        int64_t N, int improved, int add_self_loops, int64_t* row_out,
                        int64_t* col_out, float* w_out) {
    const float fill = improved ? 2.0f : 1.0f;
    int64_t Ep;
    float* w_tmp = (float*)malloc(sizeof(float) * (size_t)(E + N + 1));
    if (add_self_loops) {
        Ep = oracle_add_remaining_self_loops(row, col, w_in, E, N, fill, row_out, col_out, w_tmp);
        if (!w_in)
            for (int64_t e = 0; e < Ep; ++e) w_tmp[e] = 1.0f;
    } else {
        Ep = E;
        for (int64_t e = 0; e < E; ++e) {
            row_out[e] = row[e];
            col_out[e] = col[e];
            w_tmp[e] = w_in ? w_in[e] : 1.0f;
        }
    }
    float* deg = (float*)calloc((size_t)(N > 0 ? N : 1), sizeof(float));
    for (int64_t e = 0; e < Ep; ++e) deg[col_out[e]] += w_tmp[e];
    for (int64_t i = 0; i < N; ++i) {
        float d = powf(deg[i], -0.5f);
        if (isinf(d) && d > 0) d = 0.0f;
        deg[i] = d;
    }
    for (int64_t e = 0; e < Ep; ++e) w_out[e] = deg[row_out[e]] * w_tmp[e] * deg[col_out[e]];
    free(deg);
    free(w_tmp);
    return Ep;
}

/* out[M,N] = a[M,K] @ b[N,K]^T (+ bias[N]) -- torch_geometric/nn/dense/linear.py:121-127
 * (F.linear).  Accumulates in double so the checker is not the noisy side. */
void oracle_linear(const float* a, const float* b, const float* bias, int64_t M, int64_t K,
                   int64_t Nn, float* out) {
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < Nn; ++n) {
            double acc = 0.0;
            for (int64_t k = 0; k < K; ++k) acc += (double)a[m * K + k] * (double)b[n * K + k];
            out[m * Nn + n] = (float)acc + (bias ? bias[n] : 0.0f);
        }
}

/* torch_geometric/nn/conv/gcn_conv.py:227-268 -- GCNConv.forward on a [2,E] tensor:
 * gcn_norm -> x W^T -> propagate(sum of w * x_j over dst) -> + bias. */
void oracle_gcn_conv(const float* x, const int64_t* row, const int64_t* col, const float* w_in,
                     const float* weight, const float* bias, int64_t N, int64_t E, int64_t Fin,
                     int64_t Fout, int improved, int add_self_loops, float* out) {
    int64_t* r2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + N + 1));
    int64_t* c2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + N + 1));
    float* w2 = (float*)malloc(sizeof(float) * (size_t)(E + N + 1));
    float* xw = (float*)malloc(sizeof(float) * (size_t)(N * Fout + 1));
    const int64_t Ep = oracle_gcn_norm(row, col, w_in, E, N, improved, add_self_loops, r2, c2, w2);
    oracle_linear(x, weight, NULL, N, Fin, Fout, xw);
    oracle_gather_scatter(xw, r2, c2, w2, Ep, Fout, N, R_SUM, out);
    if (bias)
        for (int64_t i = 0; i < N; ++i)
            for (int64_t f = 0; f < Fout; ++f) out[i * Fout + f] += bias[f];
    free(r2);
    free(c2);
    free(w2);
    free(xw);
}

/* Backward of oracle_gcn_conv given grad_out [N,Fout]:
 *   grad_bias = sum_i g_i;  G = A^T-aggregate(g) ;  grad_W = G^T x ;  grad_x = G W.
 * grad_x may be NULL. */
void oracle_gcn_conv_backward(const float* grad_out, const float* x, const int64_t* row,
                              const int64_t* col, const float* w_in, const float* weight, int64_t N,
                              int64_t E, int64_t Fin, int64_t Fout, int improved,
                              int add_self_loops, float* grad_x, float* grad_weight,
                              float* grad_bias) {
    int64_t* r2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + N + 1));
    int64_t* c2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + N + 1));
    float* w2 = (float*)malloc(sizeof(float) * (size_t)(E + N + 1));
    float* G = (float*)malloc(sizeof(float) * (size_t)(N * Fout + 1));
    const int64_t Ep = oracle_gcn_norm(row, col, w_in, E, N, improved, add_self_loops, r2, c2, w2);
    /* d(xw) = scatter of w * g[dst] into src: the same gather-scatter with roles swapped */
    oracle_gather_scatter(grad_out, c2, r2, w2, Ep, Fout, N, R_SUM, G);
    if (grad_bias)
        for (int64_t f = 0; f < Fout; ++f) {
            double acc = 0.0;
            for (int64_t i = 0; i < N; ++i) acc += grad_out[i * Fout + f];
            grad_bias[f] = (float)acc;
        }
    if (grad_weight)
        for (int64_t o = 0; o < Fout; ++o)
            for (int64_t k = 0; k < Fin; ++k) {
                double acc = 0.0;
                for (int64_t i = 0; i < N; ++i) acc += (double)G[i * Fout + o] * (double)x[i * Fin + k];
                grad_weight[o * Fin + k] = (float)acc;
            }
    if (grad_x)
        for (int64_t i = 0; i < N; ++i)
            for (int64_t k = 0; k < Fin; ++k) {
                double acc = 0.0;
                for (int64_t o = 0; o < Fout; ++o) acc += (double)G[i * Fout + o] * (double)weight[o * Fin + k];
                grad_x[i * Fin + k] = (float)acc;
            }
    free(r2);
    free(c2);
    free(w2);
    free(G);
}

/* ---------------------------------------------------------------- SAGE / GIN */

/* torch_geometric/nn/conv/sage_conv.py:120-152 -- SAGEConv.forward (project=False,
 * normalize=False): out = lin_l(mean_{j in N(i)} x_j) + lin_r(x_i).  bias belongs to lin_l. */
void oracle_sage_conv(const float* x, const int64_t* row, const int64_t* col, const float* w_l,
                      const float* b_l, const float* w_r, int64_t N, int64_t E, int64_t Fin,
                      int64_t Fout, int reduce, float* out) {
    float* agg = (float*)malloc(sizeof(float) * (size_t)(N * Fin + 1));
    float* tmp = (float*)malloc(sizeof(float) * (size_t)(N * Fout + 1));
    oracle_gather_scatter(x, row, col, NULL, E, Fin, N, reduce, agg);
    oracle_linear(agg, w_l, b_l, N, Fin, Fout, out);
    if (w_r) {
        oracle_linear(x, w_r, NULL, N, Fin, Fout, tmp);
        for (int64_t i = 0; i < N * Fout; ++i) out[i] += tmp[i];
    }
    free(agg);
    free(tmp);
}

/* torch_geometric/nn/conv/gin_conv.py:73-98 -- GINConv aggregation part:
 * h = sum_{j in N(i)} x_j + (1 + eps) * x_i   (the MLP after it is plain dense code). */
void oracle_gin_aggregate(const float* x, const int64_t* row, const int64_t* col, int64_t N,
                          int64_t E, int64_t F, float eps, float* out) {
    oracle_gather_scatter(x, row, col, NULL, E, F, N, R_SUM, out);
    for (int64_t i = 0; i < N * F; ++i) out[i] += (1.0f + eps) * x[i];
}

/* ---------------------------------------------------------------- GAT */

/* torch_geometric/nn/conv/gat_conv.py:330-409 -- the attention part of GATConv.forward
 * after xh = lin(x).view(N,H,C):
 *   a_src = (xh * att_src).sum(-1), a_dst likewise                     (:330-331)
 *   remove_self_loops + add_self_loops                                 (:342-346)
 *   alpha = softmax_dst(leaky_relu(a_src[j] + a_dst[i], slope))        (:387-406)
 *   out_i = sum_e alpha_e * xh_j                                       (:408-409)
 * Outputs: out [N,H*C] (concat=True, no bias), alpha [E',H], and the E' edge list. */
int64_t oracle_gat_attention(const float* xh, const float* att_src, const float* att_dst,
                             const int64_t* row, const int64_t* col, int64_t N, int64_t E,
                             int64_t H, int64_t C, float slope, int add_self_loops,
                             int64_t* row_out, int64_t* col_out, float* alpha, float* out) {
    int64_t Ep;
    if (add_self_loops)
        Ep = oracle_remove_then_add_self_loops(row, col, E, N, row_out, col_out);
    else {
        Ep = E;
        memcpy(row_out, row, sizeof(int64_t) * (size_t)E);
        memcpy(col_out, col, sizeof(int64_t) * (size_t)E);
    }
    float* a_s = (float*)malloc(sizeof(float) * (size_t)(N * H + 1));
    float* a_d = (float*)malloc(sizeof(float) * (size_t)(N * H + 1));
    for (int64_t i = 0; i < N; ++i)
        for (int64_t h = 0; h < H; ++h) {
            float s = 0.0f, d = 0.0f;
            for (int64_t c = 0; c < C; ++c) {
                s += xh[(i * H + h) * C + c] * att_src[h * C + c];
                d += xh[(i * H + h) * C + c] * att_dst[h * C + c];
            }
            a_s[i * H + h] = s;
            a_d[i * H + h] = d;
        }
    float* logit = (float*)malloc(sizeof(float) * (size_t)(Ep * H + 1));
    for (int64_t e = 0; e < Ep; ++e)
        for (int64_t h = 0; h < H; ++h) {
            const float v = a_s[row_out[e] * H + h] + a_d[col_out[e] * H + h];
            logit[e * H + h] = v > 0.0f ? v : v * slope;
        }
    oracle_softmax(logit, col_out, Ep, H, N, alpha);
    memset(out, 0, sizeof(float) * (size_t)(N * H * C));
    for (int64_t e = 0; e < Ep; ++e)
        for (int64_t h = 0; h < H; ++h)
            for (int64_t c = 0; c < C; ++c)
                out[(col_out[e] * H + h) * C + c] += alpha[e * H + h] * xh[(row_out[e] * H + h) * C + c];
    free(a_s);
    free(a_d);
    free(logit);
    return Ep;
}

/* ---------------------------------------------------------------- RGCN */

/* torch_geometric/nn/conv/rgcn_conv.py:257-280 -- the per-relation loop (the semantic
 * ground truth): for r: h_r = propagate(edges of type r, aggr) ; out += h_r @ W[r];
 * then out += x @ root + bias.  weight is [R,Fin,Fout] (NOT transposed), root [Fin,Fout]. */
void oracle_rgcn_conv(const float* x, const int64_t* row, const int64_t* col,
                      const int64_t* edge_type, const float* weight, const float* root,
                      const float* bias, int64_t N, int64_t E, int64_t R, int64_t Fin,
                      int64_t Fout, int reduce, float* out) {
    int64_t* r2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + 1));
    int64_t* c2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + 1));
    float* h = (float*)malloc(sizeof(float) * (size_t)(N * Fin + 1));
    memset(out, 0, sizeof(float) * (size_t)(N * Fout));
    for (int64_t r = 0; r < R; ++r) {
        int64_t k = 0;
        for (int64_t e = 0; e < E; ++e)
            if (edge_type[e] == r) {
                r2[k] = row[e];
                c2[k] = col[e];
                ++k;
            }
        oracle_gather_scatter(x, r2, c2, NULL, k, Fin, N, reduce, h);
        const float* W = weight + r * Fin * Fout;
        for (int64_t i = 0; i < N; ++i)
            for (int64_t o = 0; o < Fout; ++o) {
                double acc = 0.0;
                for (int64_t kk = 0; kk < Fin; ++kk) acc += (double)h[i * Fin + kk] * (double)W[kk * Fout + o];
                out[i * Fout + o] += (float)acc;
            }
    }
    for (int64_t i = 0; i < N; ++i)
        for (int64_t o = 0; o < Fout; ++o) {
            double acc = 0.0;
            if (root)
                for (int64_t kk = 0; kk < Fin; ++kk) acc += (double)x[i * Fin + kk] * (double)root[kk * Fout + o];
            out[i * Fout + o] += (float)acc + (bias ? bias[o] : 0.0f);
        }
    free(r2);
    free(c2);
    free(h);
}
