// softmax.cu -- segment (edge) softmax over CSR groups, forward and backward, fp32.
//   out[e,h] = exp(src[e,h] - max_g) / (sum_g exp(src - max_g) + 1e-16)       (_softmax.py:60-92)
// One lane group (power-of-two width >= heads, <= 32) per CSR row; lanes run over heads so every
// access to the [E,H] matrix is a contiguous H*4-byte segment.  Three passes over the row (max,
// sum, normalise); the row is L1/L2 resident after the first.
//
// Groups with hub destinations (a power-law graph has rows of 10^5..10^6 edges) must not be walked by
// one lane group: for those the host composes the reference's own sequence (_softmax.py:82-88) from the
// chunked segmented reduce (b200mp_segment_csr with a long-row plan) and the edge-parallel
// b200mp_softmax_edge_op below -- segment max, exp(x - max[d]), segment sum, divide.
#include "common.cuh"

namespace b200mp {

constexpr int kSoftT = 256;

template <typename I>
__global__ void __launch_bounds__(kSoftT)
softmax_csr_kernel(const I* __restrict__ ptr, const float* __restrict__ src, float* __restrict__ out,
                   int64_t n_rows, int64_t heads, int g) {
    const int lig = threadIdx.x & (g - 1);
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / g;
    if (row >= n_rows) return;
    const int64_t b = ptr[row], e_ = ptr[row + 1];
    for (int64_t h = lig; h < heads; h += g) {
        float m = -__builtin_inff();
        for (int64_t e = b; e < e_; ++e) {
            const float v = __ldg(src + e * heads + h);
            m = (v > m || v != v) ? v : m;
        }
        float s = 0.0f;
        for (int64_t e = b; e < e_; ++e) s = __fadd_rn(s, expf(__ldg(src + e * heads + h) - m));
        s = __fadd_rn(s, 1e-16f);
        for (int64_t e = b; e < e_; ++e) out[e * heads + h] = __fdiv_rn(expf(__ldg(src + e * heads + h) - m), s);
    }
}

// grad_src[e,h] = out[e,h] * (g[e,h] - sum_{k in group} g[k,h] * out[k,h])
template <typename I>
__global__ void __launch_bounds__(kSoftT)
softmax_csr_backward_kernel(const I* __restrict__ ptr, const float* __restrict__ out,
                            const float* __restrict__ grad_out, float* __restrict__ grad_src, int64_t n_rows,
                            int64_t heads, int g) {
    const int lig = threadIdx.x & (g - 1);
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / g;
    if (row >= n_rows) return;
    const int64_t b = ptr[row], e_ = ptr[row + 1];
    for (int64_t h = lig; h < heads; h += g) {
        float dot = 0.0f;
        for (int64_t e = b; e < e_; ++e) dot = fmaf(__ldg(grad_out + e * heads + h), __ldg(out + e * heads + h), dot);
        for (int64_t e = b; e < e_; ++e)
            grad_src[e * heads + h] = __ldg(out + e * heads + h) * (__ldg(grad_out + e * heads + h) - dot);
    }
}

// Edge-parallel pieces of the hub-safe path; d = dst_of_edge[e].
//   op 0: out = exp(a - row[d])          op 1: out = a / (row[d] + 1e-16)
//   op 2: out = a * b                    op 3: out = a * (b - row[d])
template <typename I>
__global__ void __launch_bounds__(256)
softmax_edge_op_kernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                       const float* __restrict__ row, const I* __restrict__ dst, float* __restrict__ out,
                       int64_t n, int64_t heads) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t e = i / heads, h = i - e * heads;
    const float av = a[i];
    float r = 0.f;
    if (op != 2) r = __ldg(row + static_cast<int64_t>(dst[e]) * heads + h);
    float o;
    if (op == 0) o = expf(av - r);
    else if (op == 1) o = __fdiv_rn(av, __fadd_rn(r, 1e-16f));
    else if (op == 2) o = av * b[i];
    else o = av * (b[i] - r);
    out[i] = o;
}

inline int soft_group(int64_t heads) {
    int g = 1;
    while (g < 32 && g < heads) g <<= 1;
    return g;
}

}  // namespace b200mp

using namespace b200mp;

extern "C" int b200mp_softmax_csr(const void* ptr, const float* src, float* out, int64_t n_rows, int64_t n_src,
                                  int64_t heads, int idx_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && n_src >= 0 && heads >= 0);
    if (n_rows == 0 || n_src == 0 || heads == 0) return B200MP_OK;
    B200MP_CHECK_ARG(ptr && src && out);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int g = soft_group(heads);
    const unsigned blocks = static_cast<unsigned>(ceil_div(n_rows, kSoftT / g));
    if (idx_dtype == B200MP_I32)
        softmax_csr_kernel<int32_t><<<blocks, kSoftT, 0, s>>>(static_cast<const int32_t*>(ptr), src, out, n_rows, heads, g);
    else if (idx_dtype == B200MP_I64)
        softmax_csr_kernel<int64_t><<<blocks, kSoftT, 0, s>>>(static_cast<const int64_t*>(ptr), src, out, n_rows, heads, g);
    else {
        set_error("bad idx_dtype %d", idx_dtype);
        return B200MP_ERR_UNSUPPORTED;
    }
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}

extern "C" int b200mp_softmax_csr_backward(const void* ptr, const float* out, const float* grad_out,
                                           float* grad_src, int64_t n_rows, int64_t n_src, int64_t heads,
                                           int idx_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && n_src >= 0 && heads >= 0);
    if (n_rows == 0 || n_src == 0 || heads == 0) return B200MP_OK;
    B200MP_CHECK_ARG(ptr && out && grad_out && grad_src);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int g = soft_group(heads);
    const unsigned blocks = static_cast<unsigned>(ceil_div(n_rows, kSoftT / g));
    if (idx_dtype == B200MP_I32)
        softmax_csr_backward_kernel<int32_t><<<blocks, kSoftT, 0, s>>>(static_cast<const int32_t*>(ptr), out, grad_out, grad_src, n_rows, heads, g);
    else if (idx_dtype == B200MP_I64)
        softmax_csr_backward_kernel<int64_t><<<blocks, kSoftT, 0, s>>>(static_cast<const int64_t*>(ptr), out, grad_out, grad_src, n_rows, heads, g);
    else {
        set_error("bad idx_dtype %d", idx_dtype);
        return B200MP_ERR_UNSUPPORTED;
    }
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}

extern "C" int b200mp_softmax_edge_op(int op, const float* a, const float* b, const float* row,
                                      const void* dst_of_edge, float* out, int64_t n_src, int64_t heads,
                                      int idx_dtype, void* stream) {
    B200MP_CHECK_ARG(op >= 0 && op <= 3 && n_src >= 0 && heads >= 0);
    if (n_src == 0 || heads == 0) return B200MP_OK;
    B200MP_CHECK_ARG(a && out && (op == 2 || (row && dst_of_edge)) && (op < 2 || b));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n = n_src * heads;
    const unsigned blocks = static_cast<unsigned>(ceil_div(n, 256));
    if (idx_dtype == B200MP_I32)
        softmax_edge_op_kernel<int32_t><<<blocks, 256, 0, s>>>(op, a, b, row, static_cast<const int32_t*>(dst_of_edge), out, n, heads);
    else if (idx_dtype == B200MP_I64)
        softmax_edge_op_kernel<int64_t><<<blocks, 256, 0, s>>>(op, a, b, row, static_cast<const int64_t*>(dst_of_edge), out, n, heads);
    else {
        set_error("bad idx_dtype %d", idx_dtype);
        return B200MP_ERR_UNSUPPORTED;
    }
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}
