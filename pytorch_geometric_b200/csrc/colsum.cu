// colsum.cu -- column sum of a row-major matrix: the bias gradient of a layer, d bias = sum_i grad_out[i, :]
// (what autograd derives for `out + bias`, nn/conv/gcn_conv.py:263-264).  HBM-bound: n_rows * feat * s bytes
// read once.  Deterministic: each CTA sums one contiguous slab of rows into an fp32 partial row, a second
// launch folds the partial rows in slab order (no atomics).
#include "common.cuh"

namespace b200mp {

template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, float* __restrict__ partials, int64_t n_rows, int64_t feat,
                      int64_t rows_per_part) {
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_part;
    const int64_t r1 = r0 + rows_per_part < n_rows ? r0 + rows_per_part : n_rows;
    for (int64_t c = threadIdx.x; c < feat; c += blockDim.x) {      // a warp reads 32 consecutive columns of a row
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t r = r0;
        for (; r + 8 <= r1; r += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ElemTraits<T>::to_float(__ldg(x + (r + u) * feat + c));
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += v[u];
        }
        for (; r < r1; ++r) a[0] += ElemTraits<T>::to_float(__ldg(x + r * feat + c));
        partials[static_cast<int64_t>(blockIdx.x) * feat + c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partials, float* __restrict__ out, int64_t n_parts, int64_t feat) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= feat) return;
    float acc = 0.f;
    for (int64_t p = 0; p < n_parts; ++p) acc += partials[p * feat + c];
    out[c] = acc;
}

}  // namespace b200mp

using namespace b200mp;

extern "C" int64_t b200mp_column_sum_parts(int64_t n_rows) {
    if (n_rows <= 0) return 1;
    const int64_t want = static_cast<int64_t>(num_sms()) * 8;      // 8 resident 256-thread CTAs per SM
    const int64_t by_rows = ceil_div(n_rows, 64);                  // at least 64 rows per slab
    return want < by_rows ? want : by_rows;
}

extern "C" int b200mp_column_sum(const void* x, float* out, float* partials, int64_t n_parts, int64_t n_rows,
                                 int64_t feat, int val_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && feat >= 0 && n_parts >= 1);
    if (feat == 0) return B200MP_OK;
    B200MP_CHECK_ARG(out && partials && (x || n_rows == 0));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t rows_per_part = ceil_div(n_rows > 0 ? n_rows : 1, n_parts);
    if (val_dtype == B200MP_F32)
        colsum_partial_kernel<float><<<static_cast<unsigned>(n_parts), 256, 0, s>>>(static_cast<const float*>(x), partials,
                                                                                     n_rows, feat, rows_per_part);
    else if (val_dtype == B200MP_BF16)
        colsum_partial_kernel<__nv_bfloat16><<<static_cast<unsigned>(n_parts), 256, 0, s>>>(
            static_cast<const __nv_bfloat16*>(x), partials, n_rows, feat, rows_per_part);
    else {
        set_error("column_sum: unsupported val_dtype %d", val_dtype);
        return B200MP_ERR_UNSUPPORTED;
    }
    B200MP_LAUNCH_CHECK();
    colsum_final_kernel<<<static_cast<unsigned>(ceil_div(feat, 256)), 256, 0, s>>>(partials, out, n_parts, feat);
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}
