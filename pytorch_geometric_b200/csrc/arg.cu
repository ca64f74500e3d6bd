// arg.cu -- argmin / argmax outputs for the extension-operator signatures the reference binds to
// (SURVEY.md section 8(b)):
//   torch_scatter.scatter_max / scatter_min (src, index, dim, dim_size) -> (out, arg)     utils/_scatter.py:147-156
//   torch.ops.torch_sparse.spmm_min / spmm_max (rowptr, col, value?, mat) -> (out, arg)   edge_index.py:1798-1810
// `out` comes from the reduce kernels (scatter.cu / csr_reduce.cuh); these passes find WHO produced it:
//   COO: arg[i,f] = smallest e with index[e] == i and src[e,f] == out[i,f]   (n_src when group i is empty)
//   CSR: arg[i,f] = first CSR slot e of row i with val[e] * x[col[e],f] == out[i,f]        (nnz when row i is empty)
// -- the conventions of torch_scatter / torch_sparse (first extremum wins, sentinel = number of inputs).
// fp32 only (equality against an output that was not re-rounded).  HBM-bound: one more read of the inputs.
#include "common.cuh"

namespace b200mp {

constexpr int kArgT = 256;

__global__ void fill_i64_kernel(int64_t* p, int64_t n, int64_t v) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) p[i] = v;
}

template <typename I>
__global__ void __launch_bounds__(kArgT)
coo_arg_kernel(const float* __restrict__ src, const I* __restrict__ index, const float* __restrict__ out, int64_t* __restrict__ arg,
               int64_t n_src, int64_t n_rows, int64_t feat) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_src * feat) return;
    const int64_t e = t / feat, f = t - e * feat;
    const int64_t d = index[e];
    if (static_cast<uint64_t>(d) >= static_cast<uint64_t>(n_rows)) return;
    if (src[t] == out[d * feat + f]) atomicMin(reinterpret_cast<long long*>(arg + d * feat + f), static_cast<long long>(e));
}

template <typename I>
__global__ void __launch_bounds__(kArgT)
csr_arg_kernel(const I* __restrict__ rowptr, const I* __restrict__ col, const float* __restrict__ val, const float* __restrict__ x,
               const float* __restrict__ out, int64_t* __restrict__ arg, int64_t n_rows, int64_t feat, int64_t nnz) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_rows * feat) return;
    const int64_t row = t / feat, f = t - row * feat;
    const int64_t b = rowptr[row], e1 = rowptr[row + 1];
    const float o = out[t];
    int64_t a = nnz;
    for (int64_t e = b; e < e1; ++e) {
        const float xv = __ldg(x + static_cast<int64_t>(col[e]) * feat + f);
        const float v = val ? __fmul_rn(__ldg(val + e), xv) : xv;
        if (v == o) { a = e; break; }
    }
    arg[t] = a;
}

}  // namespace b200mp

using namespace b200mp;

extern "C" int b200mp_scatter_arg(const float* src, const void* index, const float* out, int64_t* arg, int64_t n_src,
                                  int64_t n_rows, int64_t feat, int idx_dtype, void* stream) {
    B200MP_CHECK_ARG(n_src >= 0 && n_rows >= 0 && feat >= 0);
    if (n_rows == 0 || feat == 0) return B200MP_OK;
    B200MP_CHECK_ARG(out && arg && (n_src == 0 || (src && index)));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n_out = n_rows * feat;
    unsigned fb = static_cast<unsigned>(ceil_div(n_out, kArgT));
    if (fb > 148u * 16u) fb = 148u * 16u;
    fill_i64_kernel<<<fb, kArgT, 0, s>>>(arg, n_out, n_src);
    if (n_src > 0) {
        const unsigned blocks = static_cast<unsigned>(ceil_div(n_src * feat, kArgT));
        if (idx_dtype == B200MP_I32) coo_arg_kernel<int32_t><<<blocks, kArgT, 0, s>>>(src, static_cast<const int32_t*>(index), out, arg, n_src, n_rows, feat);
        else if (idx_dtype == B200MP_I64) coo_arg_kernel<int64_t><<<blocks, kArgT, 0, s>>>(src, static_cast<const int64_t*>(index), out, arg, n_src, n_rows, feat);
        else { set_error("bad idx_dtype %d", idx_dtype); return B200MP_ERR_UNSUPPORTED; }
    }
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}

extern "C" int b200mp_spmm_csr_arg(const void* rowptr, const void* col, const float* val, const float* x, const float* out,
                                   int64_t* arg, int64_t n_rows, int64_t feat, int64_t nnz, int idx_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && feat >= 0 && nnz >= 0);
    if (n_rows == 0 || feat == 0) return B200MP_OK;
    B200MP_CHECK_ARG(rowptr && out && arg && (nnz == 0 || (col && x)));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>(ceil_div(n_rows * feat, kArgT));
    if (idx_dtype == B200MP_I32) csr_arg_kernel<int32_t><<<blocks, kArgT, 0, s>>>(static_cast<const int32_t*>(rowptr), static_cast<const int32_t*>(col), val, x, out, arg, n_rows, feat, nnz);
    else if (idx_dtype == B200MP_I64) csr_arg_kernel<int64_t><<<blocks, kArgT, 0, s>>>(static_cast<const int64_t*>(rowptr), static_cast<const int64_t*>(col), val, x, out, arg, n_rows, feat, nnz);
    else { set_error("bad idx_dtype %d", idx_dtype); return B200MP_ERR_UNSUPPORTED; }
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}
