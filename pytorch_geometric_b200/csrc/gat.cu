// gat.cu -- fused GAT attention + aggregation over the destination-sorted CSR.
//
// Reference (nn/conv/gat_conv.py:387-409, utils/_softmax.py:82-88): alpha_j + alpha_i gathers,
// leaky_relu, scatter-max, exp, scatter-sum, two more gathers, a division, an [E,H,C] message
// tensor and a scatter-add -- about twelve kernels and three E x H x C materialisations.
// Here: ONE sweep.  Forward keeps a running (max, sum, weighted accumulator) per head in
// registers (online softmax), so every source row is read exactly once:
//     HBM bytes per edge = H*C*s (row) + H*4 (a_src gather) + idx;   per node = H*C*s + 3*H*4.
// Backward is two sweeps (destination CSR, then source CSR) with the attention recomputed from the
// saved per-(node, head) max and denominator instead of storing alpha [E,H].
#include "common.cuh"

namespace b200mp {

constexpr int kGatT = 256;

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.0f ? v : v * slope; }

// ---------------------------------------------------------------- forward, vectorised
// Lane group of G lanes per destination row; lane owns VPL 16-byte vectors; every vector lies
// inside one head (C % EPV == 0).
template <typename T, typename I, int G, int VPL>
__global__ void __launch_bounds__(kGatT)
gat_fwd_vec_kernel(const I* __restrict__ rowptr, const I* __restrict__ col, const T* __restrict__ xh,
                   const float* __restrict__ a_src, const float* __restrict__ a_dst, T* __restrict__ out,
                   float* __restrict__ row_max, float* __restrict__ row_den, int64_t n_rows, int heads,
                   int chan, int n_vec, float slope) {
    constexpr int EPV = ElemTraits<T>::kPerVec;
    constexpr int UNR = VPL == 1 ? 4 : 2;
    const int lig = threadIdx.x & (G - 1);
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    if (row >= n_rows) return;
    const int64_t begin = rowptr[row], end = rowptr[row + 1];
    const size_t row_bytes = static_cast<size_t>(n_vec) * 16;
    const char* xb = reinterpret_cast<const char*>(xh);

    int head[VPL];
    bool valid[VPL];
    float ad[VPL], m[VPL], s[VPL], acc[VPL][EPV];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int v = lig + k * G;
        valid[k] = v < n_vec;
        head[k] = valid[k] ? (v * EPV) / chan : 0;
        ad[k] = valid[k] ? __ldg(a_dst + row * heads + head[k]) : 0.0f;
        m[k] = -__builtin_inff();
        s[k] = 0.0f;
#pragma unroll
        for (int i = 0; i < EPV; ++i) acc[k][i] = 0.0f;
    }
    for (int64_t e = begin; e < end; e += UNR) {
        Vec16 buf[UNR][VPL];
        float as[UNR][VPL];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (e + u < end) {
                const int64_t c = col[e + u];
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    if (valid[k]) {
                        buf[u][k] = ldg_row16(xb + static_cast<size_t>(c) * row_bytes + static_cast<size_t>(lig + k * G) * 16);
                        as[u][k] = __ldg(a_src + c * heads + head[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (e + u < end) {
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    if (valid[k]) {
                        const float l = leaky(as[u][k] + ad[k], slope);
                        const float mn = fmaxf(m[k], l);
                        const float sc = expf(m[k] - mn);      // 0 on the first edge (m = -inf)
                        const float p = expf(l - mn);
                        s[k] = fmaf(s[k], sc, p);
                        float f[EPV];
                        ElemTraits<T>::unpack(buf[u][k], f);
#pragma unroll
                        for (int i = 0; i < EPV; ++i) acc[k][i] = fmaf(acc[k][i], sc, p * f[i]);
                        m[k] = mn;
                    }
                }
            }
        }
    }
    char* ob = reinterpret_cast<char*>(out) + static_cast<size_t>(row) * row_bytes;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        if (!valid[k]) continue;
        const int v = lig + k * G;
        const float den = s[k] + 1e-16f;                       // _softmax.py:87 "+ 1e-16"
        float f[EPV];
#pragma unroll
        for (int i = 0; i < EPV; ++i) f[i] = (end > begin) ? acc[k][i] / den : 0.0f;
        stg_stream16(ob + static_cast<size_t>(v) * 16, ElemTraits<T>::pack(f));
        if ((v * EPV) % chan == 0) {                           // first vector of its head
            row_max[row * heads + head[k]] = (end > begin) ? m[k] : 0.0f;
            row_den[row * heads + head[k]] = den;
        }
    }
}

// ---------------------------------------------------------------- forward, generic (any H, C)
// One thread per (destination, head): exact two-pass softmax statistics, then CC channels at a time.
template <typename T, typename I, int CC>
__global__ void __launch_bounds__(kGatT)
gat_fwd_scalar_kernel(const I* __restrict__ rowptr, const I* __restrict__ col, const T* __restrict__ xh,
                      const float* __restrict__ a_src, const float* __restrict__ a_dst, T* __restrict__ out,
                      float* __restrict__ row_max, float* __restrict__ row_den, int64_t n_rows, int heads,
                      int chan, float slope) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_rows * heads) return;
    const int64_t row = t / heads;
    const int h = static_cast<int>(t - row * heads);
    const int64_t begin = rowptr[row], end = rowptr[row + 1];
    const float ad = a_dst[t];
    float m = -__builtin_inff();
    for (int64_t e = begin; e < end; ++e) m = fmaxf(m, leaky(__ldg(a_src + static_cast<int64_t>(col[e]) * heads + h) + ad, slope));
    float s = 0.0f;
    for (int64_t e = begin; e < end; ++e) s += expf(leaky(__ldg(a_src + static_cast<int64_t>(col[e]) * heads + h) + ad, slope) - m);
    const float den = s + 1e-16f;
    row_max[t] = end > begin ? m : 0.0f;
    row_den[t] = den;
    const int64_t hc = static_cast<int64_t>(heads) * chan;
    for (int c0 = 0; c0 < chan; c0 += CC) {
        float acc[CC];
#pragma unroll
        for (int i = 0; i < CC; ++i) acc[i] = 0.0f;
        for (int64_t e = begin; e < end; ++e) {
            const int64_t c = col[e];
            const float p = expf(leaky(__ldg(a_src + c * heads + h) + ad, slope) - m) / den;
            const T* xr = xh + c * hc + static_cast<int64_t>(h) * chan + c0;
#pragma unroll
            for (int i = 0; i < CC; ++i)
                if (c0 + i < chan) acc[i] = fmaf(p, ElemTraits<T>::to_float(xr[i]), acc[i]);
        }
        T* orow = out + row * hc + static_cast<int64_t>(h) * chan + c0;
#pragma unroll
        for (int i = 0; i < CC; ++i)
            if (c0 + i < chan) orow[i] = ElemTraits<T>::from_float(acc[i]);
    }
}

// alpha[e,h] in CSR order from the saved statistics (return_attention_weights, gat_conv.py:374-383)
template <typename I>
__global__ void __launch_bounds__(kGatT)
gat_alpha_kernel(const I* __restrict__ rowptr, const I* __restrict__ col, const float* __restrict__ a_src,
                 const float* __restrict__ a_dst, const float* __restrict__ row_max,
                 const float* __restrict__ row_den, float* __restrict__ alpha, int64_t n_rows, int heads,
                 float slope) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_rows * heads) return;
    const int64_t row = t / heads;
    const int h = static_cast<int>(t - row * heads);
    const float ad = a_dst[t], m = row_max[t], den = row_den[t];
    for (int64_t e = rowptr[row]; e < rowptr[row + 1]; ++e)
        alpha[e * heads + h] = expf(leaky(__ldg(a_src + static_cast<int64_t>(col[e]) * heads + h) + ad, slope) - m) / den;
}

// ---------------------------------------------------------------- backward, destination sweep
// thread (i,h):  D = <g[i,h,:], out[i,h,:]>;  per edge: alpha, dot = <g[i,h,:], xh[j,h,:]>,
//   grad_logit = alpha * (dot - D);  grad_pre = grad_logit * leaky'(pre);  grad_a_dst[i,h] = sum.
template <typename T, typename I>
__global__ void __launch_bounds__(kGatT)
gat_bwd_dst_kernel(const I* __restrict__ rowptr, const I* __restrict__ col, const T* __restrict__ xh,
                   const float* __restrict__ a_src, const float* __restrict__ a_dst,
                   const float* __restrict__ row_max, const float* __restrict__ row_den,
                   const T* __restrict__ out, const T* __restrict__ grad_out, float* __restrict__ grad_pre,
                   float* __restrict__ grad_a_dst, int64_t n_rows, int heads, int chan, float slope) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_rows * heads) return;
    const int64_t row = t / heads;
    const int h = static_cast<int>(t - row * heads);
    const int64_t hc = static_cast<int64_t>(heads) * chan;
    const T* g = grad_out + row * hc + static_cast<int64_t>(h) * chan;
    const T* o = out + row * hc + static_cast<int64_t>(h) * chan;
    float D = 0.0f;
    for (int c = 0; c < chan; ++c) D = fmaf(ElemTraits<T>::to_float(g[c]), ElemTraits<T>::to_float(o[c]), D);
    const float ad = a_dst[t], m = row_max[t], den = row_den[t];
    float sum = 0.0f;
    for (int64_t e = rowptr[row]; e < rowptr[row + 1]; ++e) {
        const int64_t j = col[e];
        const float pre = __ldg(a_src + j * heads + h) + ad;
        const float alpha = expf(leaky(pre, slope) - m) / den;
        const T* xr = xh + j * hc + static_cast<int64_t>(h) * chan;
        float dot = 0.0f;
        for (int c = 0; c < chan; ++c) dot = fmaf(ElemTraits<T>::to_float(g[c]), ElemTraits<T>::to_float(xr[c]), dot);
        const float gp = alpha * (dot - D) * (pre > 0.0f ? 1.0f : slope);
        grad_pre[e * heads + h] = gp;
        sum += gp;
    }
    grad_a_dst[t] = sum;
}

// ---------------------------------------------------------------- backward, source sweep
// thread (j,h) on the transposed CSR: grad_xh[j,h,:] = sum_e alpha_e * g[d_e,h,:];
// grad_a_src[j,h] = sum_e grad_pre[csr_slot(e), h].
template <typename T, typename I, int CC>
__global__ void __launch_bounds__(kGatT)
gat_bwd_src_kernel(const I* __restrict__ rowptr_t, const I* __restrict__ col_t, const I* __restrict__ t2csr,
                   const float* __restrict__ a_src, const float* __restrict__ a_dst,
                   const float* __restrict__ row_max, const float* __restrict__ row_den,
                   const T* __restrict__ grad_out, const float* __restrict__ grad_pre,
                   T* __restrict__ grad_xh, float* __restrict__ grad_a_src, int64_t n_src, int heads, int chan,
                   float slope) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_src * heads) return;
    const int64_t j = t / heads;
    const int h = static_cast<int>(t - j * heads);
    const int64_t hc = static_cast<int64_t>(heads) * chan;
    const int64_t begin = rowptr_t[j], end = rowptr_t[j + 1];
    const float as = a_src[t];
    float sum = 0.0f;
    for (int64_t e = begin; e < end; ++e) sum += __ldg(grad_pre + static_cast<int64_t>(t2csr[e]) * heads + h);
    grad_a_src[t] = sum;
    for (int c0 = 0; c0 < chan; c0 += CC) {
        float acc[CC];
#pragma unroll
        for (int i = 0; i < CC; ++i) acc[i] = 0.0f;
        for (int64_t e = begin; e < end; ++e) {
            const int64_t d = col_t[e];
            const float alpha = expf(leaky(as + __ldg(a_dst + d * heads + h), slope) - __ldg(row_max + d * heads + h)) /
                                __ldg(row_den + d * heads + h);
            const T* g = grad_out + d * hc + static_cast<int64_t>(h) * chan + c0;
#pragma unroll
            for (int i = 0; i < CC; ++i)
                if (c0 + i < chan) acc[i] = fmaf(alpha, ElemTraits<T>::to_float(g[i]), acc[i]);
        }
        T* gx = grad_xh + j * hc + static_cast<int64_t>(h) * chan + c0;
#pragma unroll
        for (int i = 0; i < CC; ++i)
            if (c0 + i < chan) gx[i] = ElemTraits<T>::from_float(acc[i]);
    }
}

template <typename T, typename I>
int gat_fwd_typed(const void* rowptr_, const void* col_, const void* xh_, const float* a_src, const float* a_dst,
                  void* out_, float* row_max, float* row_den, float* alpha_out, int64_t n_rows, int64_t heads,
                  int64_t chan, float slope, cudaStream_t s) {
    constexpr int EPV = ElemTraits<T>::kPerVec;
    const I* rowptr = static_cast<const I*>(rowptr_);
    const I* col = static_cast<const I*>(col_);
    const T* xh = static_cast<const T*>(xh_);
    T* out = static_cast<T*>(out_);
    const size_t row_bytes = static_cast<size_t>(heads * chan) * sizeof(T);
    const bool vec_ok = row_bytes % 16 == 0 && chan % EPV == 0 && aligned16(xh) && aligned16(out) && row_bytes / 16 <= 128;
    if (vec_ok) {
        const int n_vec = static_cast<int>(row_bytes / 16);
#define GAT_LV(G_, V_)                                                                                      \
    gat_fwd_vec_kernel<T, I, G_, V_><<<static_cast<unsigned>(ceil_div(n_rows, kGatT / G_)), kGatT, 0, s>>>(  \
        rowptr, col, xh, a_src, a_dst, out, row_max, row_den, n_rows, static_cast<int>(heads),              \
        static_cast<int>(chan), n_vec, slope)
        if (n_vec <= 1) GAT_LV(1, 1);
        else if (n_vec <= 2) GAT_LV(2, 1);
        else if (n_vec <= 4) GAT_LV(4, 1);
        else if (n_vec <= 8) GAT_LV(8, 1);
        else if (n_vec <= 16) GAT_LV(16, 1);
        else if (n_vec <= 32) GAT_LV(32, 1);
        else if (n_vec <= 64) GAT_LV(32, 2);
        else GAT_LV(32, 4);
#undef GAT_LV
    } else {
        gat_fwd_scalar_kernel<T, I, 8><<<static_cast<unsigned>(ceil_div(n_rows * heads, kGatT)), kGatT, 0, s>>>(
            rowptr, col, xh, a_src, a_dst, out, row_max, row_den, n_rows, static_cast<int>(heads),
            static_cast<int>(chan), slope);
    }
    B200MP_LAUNCH_CHECK();
    if (alpha_out) {
        gat_alpha_kernel<I><<<static_cast<unsigned>(ceil_div(n_rows * heads, kGatT)), kGatT, 0, s>>>(
            rowptr, col, a_src, a_dst, row_max, row_den, alpha_out, n_rows, static_cast<int>(heads), slope);
        B200MP_LAUNCH_CHECK();
    }
    return B200MP_OK;
}

template <typename T, typename I>
int gat_bwd_typed(const void* rowptr, const void* col, const void* rowptr_t, const void* col_t, const void* t2csr,
                  const void* xh, const float* a_src, const float* a_dst, const float* row_max,
                  const float* row_den, const void* out, const void* grad_out, float* grad_pre, void* grad_xh,
                  float* grad_a_src, float* grad_a_dst, int64_t n_rows, int64_t n_src, int64_t heads,
                  int64_t chan, float slope, cudaStream_t s) {
    if (n_rows > 0)
        gat_bwd_dst_kernel<T, I><<<static_cast<unsigned>(ceil_div(n_rows * heads, kGatT)), kGatT, 0, s>>>(
            static_cast<const I*>(rowptr), static_cast<const I*>(col), static_cast<const T*>(xh), a_src, a_dst,
            row_max, row_den, static_cast<const T*>(out), static_cast<const T*>(grad_out), grad_pre, grad_a_dst,
            n_rows, static_cast<int>(heads), static_cast<int>(chan), slope);
    if (n_src > 0)
        gat_bwd_src_kernel<T, I, 8><<<static_cast<unsigned>(ceil_div(n_src * heads, kGatT)), kGatT, 0, s>>>(
            static_cast<const I*>(rowptr_t), static_cast<const I*>(col_t), static_cast<const I*>(t2csr), a_src,
            a_dst, row_max, row_den, static_cast<const T*>(grad_out), grad_pre, static_cast<T*>(grad_xh),
            grad_a_src, n_src, static_cast<int>(heads), static_cast<int>(chan), slope);
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}

}  // namespace b200mp

using namespace b200mp;

#define GAT_DISPATCH(FN, ...)                                                                                   \
    do {                                                                                                        \
        if (val_dtype == B200MP_F32 && idx_dtype == B200MP_I32) return FN<float, int32_t>(__VA_ARGS__);         \
        if (val_dtype == B200MP_F32 && idx_dtype == B200MP_I64) return FN<float, int64_t>(__VA_ARGS__);         \
        if (val_dtype == B200MP_BF16 && idx_dtype == B200MP_I32) return FN<__nv_bfloat16, int32_t>(__VA_ARGS__); \
        if (val_dtype == B200MP_BF16 && idx_dtype == B200MP_I64) return FN<__nv_bfloat16, int64_t>(__VA_ARGS__); \
        set_error("gat: unsupported dtype combination val=%d idx=%d", val_dtype, idx_dtype);                    \
        return B200MP_ERR_UNSUPPORTED;                                                                          \
    } while (0)

extern "C" int b200mp_gat_fused_csr(const void* rowptr, const void* col, const void* xh, const float* a_src,
                                    const float* a_dst, void* out, float* row_max, float* row_den,
                                    float* alpha_out, int64_t n_rows, int64_t heads, int64_t chan, float slope,
                                    int idx_dtype, int val_dtype, void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && heads > 0 && chan > 0);
    if (n_rows == 0) return B200MP_OK;
    B200MP_CHECK_ARG(rowptr && a_dst && out && row_max && row_den);
    GAT_DISPATCH(gat_fwd_typed, rowptr, col, xh, a_src, a_dst, out, row_max, row_den, alpha_out, n_rows, heads,
                 chan, slope, static_cast<cudaStream_t>(stream));
}

extern "C" int b200mp_gat_fused_csr_backward(const void* rowptr, const void* col, const void* rowptr_t,
                                             const void* col_t, const void* t2csr, const void* xh,
                                             const float* a_src, const float* a_dst, const float* row_max,
                                             const float* row_den, const void* out, const void* grad_out,
                                             float* grad_pre, void* grad_xh, float* grad_a_src,
                                             float* grad_a_dst, int64_t n_rows, int64_t n_src, int64_t heads,
                                             int64_t chan, float slope, int idx_dtype, int val_dtype,
                                             void* stream) {
    B200MP_CHECK_ARG(n_rows >= 0 && n_src >= 0 && heads > 0 && chan > 0);
    B200MP_CHECK_ARG(rowptr && rowptr_t && grad_xh && grad_a_src && grad_a_dst);
    GAT_DISPATCH(gat_bwd_typed, rowptr, col, rowptr_t, col_t, t2csr, xh, a_src, a_dst, row_max, row_den, out,
                 grad_out, grad_pre, grad_xh, grad_a_src, grad_a_dst, n_rows, n_src, heads, chan, slope,
                 static_cast<cudaStream_t>(stream));
}
