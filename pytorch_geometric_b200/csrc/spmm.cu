// spmm.cu -- C ABI for the gather + segmented-reduce path (b200mp_spmm_csr) and its backward
// helpers (min/max tie counting, SDDMM for the edge-weight gradient).
#include "csr_dispatch.cuh"

namespace b200mp {

// ---------------------------------------------------------------- min/max backward
// ties[i,f] = [count_self_zero && out[i,f] == 0] + #{e in row i : val[e]*x[col[e],f] == out[i,f]}
template <typename T, typename I>
__global__ void __launch_bounds__(256)
minmax_ties_kernel(const I* __restrict__ rowptr, const I* __restrict__ col,
                   const float* __restrict__ val, const T* __restrict__ x, const T* __restrict__ out,
                   float* __restrict__ ties, int64_t n_rows, int64_t feat, int g, bool count_self_zero) {
    const int lig = threadIdx.x & (g - 1);
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / g;
    if (row >= n_rows) return;
    const int64_t begin = rowptr[row], end = rowptr[row + 1];
    for (int64_t f = lig; f < feat; f += g) {
        const float o = ElemTraits<T>::to_float(out[row * feat + f]);
        float cnt = (count_self_zero && o == 0.0f) ? 1.0f : 0.0f;
        for (int64_t e = begin; e < end; e += 4) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = 0.f;
                if (e + u < end) {
                    const int64_t c = col[e + u];
                    const float xv = ElemTraits<T>::to_float(x[c * feat + f]);
                    // the forward rounds the product to T before comparing (out is stored as T)
                    v[u] = ElemTraits<T>::to_float(ElemTraits<T>::from_float(val ? __fmul_rn(__ldg(val + e + u), xv) : xv));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e + u < end && v[u] == o) cnt += 1.0f;
        }
        ties[row * feat + f] = cnt;
    }
}

// grad_x[j,f] = sum_{e in rowT(j)} [valT*x[j,f] == out[d,f]] * valT * g[d,f] / ties[d,f],  d = colT[e]
template <typename T, typename I>
__global__ void __launch_bounds__(256)
minmax_backward_kernel(const I* __restrict__ rowptr_t, const I* __restrict__ col_t,
                       const float* __restrict__ val_t, const T* __restrict__ x,
                       const T* __restrict__ out, const T* __restrict__ grad_out,
                       const float* __restrict__ ties, T* __restrict__ grad_x, int64_t n_src,
                       int64_t feat, int g) {
    const int lig = threadIdx.x & (g - 1);
    const int64_t j = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / g;
    if (j >= n_src) return;
    const int64_t begin = rowptr_t[j], end = rowptr_t[j + 1];
    for (int64_t f = lig; f < feat; f += g) {
        const float xv = ElemTraits<T>::to_float(x[j * feat + f]);
        float acc = 0.0f;
        for (int64_t e = begin; e < end; ++e) {
            const int64_t d = col_t[e];
            const float w = val_t ? __ldg(val_t + e) : 1.0f;
            const float m = ElemTraits<T>::to_float(ElemTraits<T>::from_float(val_t ? __fmul_rn(w, xv) : xv));
            const float o = ElemTraits<T>::to_float(out[d * feat + f]);
            if (m == o) {
                const float gd = __fdiv_rn(ElemTraits<T>::to_float(grad_out[d * feat + f]), ties[d * feat + f]);
                acc = __fadd_rn(acc, val_t ? __fmul_rn(w, gd) : gd);
            }
        }
        grad_x[j * feat + f] = ElemTraits<T>::from_float(acc);
    }
}

// ---------------------------------------------------------------- SDDMM
// dot[e] = <a[row,:], b[col[e],:]>, one warp per CSR row, a[row] held in registers (up to 8
// values per lane, re-read from L1 beyond that), 5-step xor-shuffle reduction per edge.
template <typename T, typename I>
__global__ void __launch_bounds__(256)
sddmm_kernel(const I* __restrict__ rowptr, const I* __restrict__ col, const T* __restrict__ a,
             const T* __restrict__ b, float* __restrict__ dot, int64_t n_rows, int64_t feat) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int64_t begin = rowptr[row], end = rowptr[row + 1];
    const T* ar = a + row * feat;
    for (int64_t e = begin; e < end; e += 2) {
        const bool has1 = e + 1 < end;
        const T* b0 = b + static_cast<int64_t>(col[e]) * feat;
        const T* b1 = has1 ? b + static_cast<int64_t>(col[e + 1]) * feat : b0;
        float s0 = 0.f, s1 = 0.f;
        for (int64_t f = lane; f < feat; f += 32) {
            const float av = ElemTraits<T>::to_float(ar[f]);
            s0 = fmaf(av, ElemTraits<T>::to_float(b0[f]), s0);
            s1 = fmaf(av, ElemTraits<T>::to_float(b1[f]), s1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) {
            dot[e] = s0;
            if (has1) dot[e + 1] = s1;
        }
    }
}

template <typename T, typename I>
int spmm_typed(const void* rowptr, const void* col, const float* val, const void* x, void* out,
               int64_t n_rows, int64_t feat, int reduce, LongRowPlan plan, const float* bias,
               cudaStream_t stream) {
    return csr_reduce_auto<T, I, true>(static_cast<const I*>(rowptr), static_cast<const I*>(col), val,
                                        static_cast<const T*>(x), static_cast<T*>(out), n_rows, feat,
                                        reduce, false, plan, bias, stream);
}

inline int group_width(int64_t feat) {
    int g = 1;
    while (g < 32 && g < feat) g <<= 1;
    return g;
}

}  // namespace b200mp

using namespace b200mp;

#define DISPATCH_T_I(FN, ...)                                                        \
    do {                                                                             \
        if (val_dtype == B200MP_F32 && idx_dtype == B200MP_I32) return FN<float, int32_t>(__VA_ARGS__);        \
        if (val_dtype == B200MP_F32 && idx_dtype == B200MP_I64) return FN<float, int64_t>(__VA_ARGS__);        \
        if (val_dtype == B200MP_BF16 && idx_dtype == B200MP_I32) return FN<__nv_bfloat16, int32_t>(__VA_ARGS__); \
        if (val_dtype == B200MP_BF16 && idx_dtype == B200MP_I64) return FN<__nv_bfloat16, int64_t>(__VA_ARGS__); \
        set_error("unsupported dtype combination val=%d idx=%d", val_dtype, idx_dtype);                         \
        return B200MP_ERR_UNSUPPORTED;                                                                          \
    } while (0)

extern "C" int b200mp_spmm_csr(const void* rowptr, const void* col, const float* val, const void* x,
                               void* out, int64_t n_rows, int64_t n_cols, int64_t feat, int reduce,
                               const int64_t* long_rows, const int64_t* chunk_ptr,
                               int64_t n_long_rows, int64_t n_chunks, int64_t chunk, float* partials,
                               const float* bias, const void* x_halo, int64_t n_local_cols, int flags,
                               const void* peer_ptrs, int64_t peer_rows, const void* relu_mask, int idx_dtype,
                               int val_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && feat >= 0);
    B200MP_CHECK_ARG((flags & ~1) == 0 && (!(flags & 1) || (reduce == B200MP_SUM && !bias)));
    B200MP_CHECK_ARG(!relu_mask || (flags & 1));
    if (n_rows == 0 || feat == 0) return B200MP_OK;
    B200MP_CHECK_ARG(rowptr && out);
    B200MP_CHECK_ARG(x || n_cols == 0 || peer_ptrs);
    B200MP_CHECK_ARG(!peer_ptrs || (peer_rows > 0 && !x_halo && (feat * (val_dtype == B200MP_BF16 ? 2 : 4)) % 16 == 0));
    B200MP_CHECK_ARG(n_long_rows >= 0 && n_chunks >= 0);
    B200MP_CHECK_ARG(n_long_rows == 0 || (long_rows && chunk_ptr && partials && chunk > 0));
    B200MP_CHECK_ARG(!x_halo || (n_local_cols >= 0 && n_local_cols <= n_cols));
    LongRowPlan plan{long_rows, chunk_ptr, n_long_rows, n_long_rows ? n_chunks : 0, chunk, partials,
                     x_halo, x_halo ? n_local_cols : 0, flags & 1,
                     static_cast<const unsigned long long*>(peer_ptrs), peer_ptrs ? peer_rows : 0, relu_mask};
    DISPATCH_T_I(spmm_typed, rowptr, col, val, x, out, n_rows, feat, reduce, plan, bias,
                 static_cast<cudaStream_t>(stream));
}

namespace b200mp {
template <typename T, typename I>
int ties_typed(const void* rowptr, const void* col, const float* val, const void* x, const void* out,
               float* ties, int64_t n_rows, int64_t feat, int count_self_zero, cudaStream_t stream) {
    const int g = group_width(feat);
    const int64_t blocks = ceil_div(n_rows, 256 / g);
    minmax_ties_kernel<T, I><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const I*>(rowptr), static_cast<const I*>(col), val, static_cast<const T*>(x),
        static_cast<const T*>(out), ties, n_rows, feat, g, count_self_zero != 0);
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}
template <typename T, typename I>
int mmbwd_typed(const void* rowptr_t, const void* col_t, const float* val_t, const void* x,
                const void* out, const void* grad_out, const float* ties, void* grad_x, int64_t n_src,
                int64_t feat, cudaStream_t stream) {
    const int g = group_width(feat);
    const int64_t blocks = ceil_div(n_src, 256 / g);
    minmax_backward_kernel<T, I><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const I*>(rowptr_t), static_cast<const I*>(col_t), val_t, static_cast<const T*>(x),
        static_cast<const T*>(out), static_cast<const T*>(grad_out), ties, static_cast<T*>(grad_x), n_src,
        feat, g);
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}
template <typename T, typename I>
int sddmm_typed(const void* rowptr, const void* col, const void* a, const void* b, float* dot,
                int64_t n_rows, int64_t feat, cudaStream_t stream) {
    const int64_t blocks = ceil_div(n_rows, 8);
    sddmm_kernel<T, I><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const I*>(rowptr), static_cast<const I*>(col), static_cast<const T*>(a),
        static_cast<const T*>(b), dot, n_rows, feat);
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}
}  // namespace b200mp

extern "C" int b200mp_minmax_ties(const void* rowptr, const void* col, const float* val, const void* x,
                                  const void* out, float* ties, int64_t n_rows, int64_t feat,
                                  int count_self_zero, int idx_dtype, int val_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && feat >= 0);
    if (n_rows == 0 || feat == 0) return B200MP_OK;
    B200MP_CHECK_ARG(rowptr && out && ties);
    DISPATCH_T_I(ties_typed, rowptr, col, val, x, out, ties, n_rows, feat, count_self_zero,
                 static_cast<cudaStream_t>(stream));
}

extern "C" int b200mp_minmax_backward(const void* rowptr_t, const void* col_t, const float* val_t,
                                      const void* x, const void* out, const void* grad_out,
                                      const float* ties, void* grad_x, int64_t n_src, int64_t feat,
                                      int idx_dtype, int val_dtype, void* stream) {
    B200MP_CHECK_ARG(n_src >= 0 && feat >= 0);
    if (n_src == 0 || feat == 0) return B200MP_OK;
    B200MP_CHECK_ARG(rowptr_t && x && grad_x);
    DISPATCH_T_I(mmbwd_typed, rowptr_t, col_t, val_t, x, out, grad_out, ties, grad_x, n_src, feat,
                 static_cast<cudaStream_t>(stream));
}

extern "C" int b200mp_sddmm_csr(const void* rowptr, const void* col, const void* a, const void* b,
                                float* dot, int64_t n_rows, int64_t feat, int idx_dtype, int val_dtype,
                                void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && feat >= 0);
    if (n_rows == 0) return B200MP_OK;
    B200MP_CHECK_ARG(rowptr && dot);
    DISPATCH_T_I(sddmm_typed, rowptr, col, a, b, dot, n_rows, feat, static_cast<cudaStream_t>(stream));
}
