// gemm_tf32x3_ts.cuh -- "TS" variant of the 3xTF32 GEMM: the A operand lives in TENSOR MEMORY.
//
// Why (profiles/r1_summary.md section 3): with both operands in shared memory (SS mode) every K = 8
// step of tcgen05.mma.kind::tf32 re-reads 4 KB of A and 8 KB of B as fp32, and the three products of
// the 3xTF32 split triple that: 272 KB of shared-memory traffic per 32-wide k-block against
// 1536 tensor-pipe cycles -- the SS kernel is shared-memory-bandwidth bound at 45 % tensor activity.
// Here the splitter warps (which already touch every element of the A tile to produce hi/lo) write
// a_hi and a_lo straight into TMEM with tcgen05.st (row m -> TMEM lane m, k -> column), and the MMA
// takes A from TMEM:  tcgen05.mma.cta_group::1.kind::tf32 [d], [a_tmem], b_desc, idesc, p.
// Shared memory then only feeds B (4 KB per step at N = 128) and receives the TMA fills.
//
// Tile: BM = 128, BN = 128, BK = 32; 4 smem stages of 48 KB (A raw 16 KB + B_hi 16 KB + B_lo 16 KB);
// TMEM: 2 accumulator stages x 128 columns + 4 A slots x (32 hi + 32 lo) columns = 512 columns.
// A is K-major (row-major [M, K]); B is K-major or MN-major, either pre-split (W_hi, W_lo) or -- B_SPLIT --
// loaded raw and split in place by the splitter warps.  Why B_SPLIT: every output tile re-reads its B tiles
// from L2, and at M = 10^7 rows the kernel's L2 -> SM traffic (A twice, B_hi + B_lo per tile, C once: 70 GB
// for the 256 x 256 layer) runs into the L2 slice throughput (~6300 B/clk chip-wide) before the tensor pipe
// saturates; loading W once per tile instead of W_hi and W_lo removes 20 GB of it.
#pragma once

namespace b200mp {

constexpr int kTsBN = 128;
constexpr int kTsStages = 4;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMA store of one [32 rows x 32 fp32] box from shared memory (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// GROUPED: segment_matmul -- the row space is cut into segments (args.seg_ptr), every segment multiplies with its own B
// block; tiles never straddle a segment (the last tile of a segment is partial and is stored with row-masked writes).
constexpr int kMaxSegments = 120;   // (120 + 2) ints stay inside ONE 1 KB granule next to 225.25 KB of dynamic stages (227 KB per CTA)
struct GroupedTile {
    int m0, n0, seg, rows;     // first row, first column, segment, valid rows (<= 128)
};
template <bool GROUPED>
__device__ __forceinline__ bool ts_decode(int w, const GemmArgs& args, const int* tile_prefix, GroupedTile& t) {
    const int mt = w / args.n_tiles_n;
    t.n0 = (w % args.n_tiles_n) * kTsBN;
    if (!GROUPED) {
        t.m0 = mt * kBM;
        t.seg = 0;
        t.rows = kBM;
        return true;
    }
    int r = 0;
    while (r < args.n_seg && tile_prefix[r + 1] <= mt) ++r;          // few segments: a linear scan of shared memory
    if (r >= args.n_seg) return false;
    const int64_t s0 = args.seg_ptr[r], s1 = args.seg_ptr[r + 1];
    const int64_t m0 = s0 + static_cast<int64_t>(mt - tile_prefix[r]) * kBM;
    t.m0 = static_cast<int>(m0);
    t.seg = r;
    t.rows = static_cast<int>(s1 - m0 < kBM ? s1 - m0 : kBM);
    return true;
}

template <bool B_MN, bool B_SPLIT, bool GROUPED = false>
__global__ void __launch_bounds__(kGemmThreadsTs, 1)
gemm_tf32x3_ts_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                      const __grid_constant__ CUtensorMap tmap_b_hi, const __grid_constant__ CUtensorMap tmap_b_lo,
                      const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2, GemmArgs args) {
    constexpr int BN = kTsBN, BK = 32, kStages = kTsStages;
    constexpr uint32_t kSlab = BK * 128;
    constexpr uint32_t kABytes = kBM * BK * 4;                  // raw fp32 A tile (K-major, 128B swizzle)
    constexpr uint32_t kBBytes = BN * BK * 4;
    constexpr uint32_t kStageBytes = kABytes + 2 * kBBytes;     // 48 KB
    constexpr uint32_t kTxBytes = B_SPLIT ? kABytes + kBBytes : kStageBytes;
    constexpr uint32_t kAccCols = kAccStages * BN;              // 256
    constexpr uint32_t kASlotCols = 2 * BK;                     // hi | lo
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdesc = instr_desc(kBM, BN, false, B_MN);

    extern __shared__ __align__(1024) unsigned char gemm_smem[];
    const uint32_t smem_base = (s2u(gemm_smem) + 1023u) & ~1023u;
    unsigned char* smem_gen = gemm_smem + (smem_base - s2u(gemm_smem));
    // epilogue staging: per epilogue warp two [32 x 32] fp32 boxes (128-byte rows, 128B swizzle)
    const uint32_t staging = smem_base + kStages * kStageBytes;
    const uint32_t bars = staging + 4u * 2u * 4096u;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_split = [&](int s) { return bars + 8u * (kStages + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto bar_tfull = [&](int a) { return bars + 8u * (3 * kStages + a); };
    auto bar_tempty = [&](int a) { return bars + 8u * (3 * kStages + kAccStages + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * kStages + 2 * kAccStages);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    __shared__ int tile_prefix[GROUPED ? kMaxSegments + 2 : 1];   // (the dynamic stages leave ~1.7 KB of the 227 KB)
    if (GROUPED && threadIdx.x == 64) {                          // (an idle warp) tiles of 128 rows per segment, exclusive prefix
        int acc = 0;
        for (int r = 0; r < args.n_seg; ++r) {
            tile_prefix[r] = acc;
            acc += static_cast<int>((args.seg_ptr[r + 1] - args.seg_ptr[r] + kBM - 1) / kBM);
        }
        tile_prefix[args.n_seg] = acc;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(bar_full(s), 1);
            bar_init(bar_split(s), 128);
            bar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            bar_init(bar_tfull(a), 1);
            bar_init(bar_tempty(a), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    const uint32_t tmem_a0 = tmem_base + kAccCols;              // A slots start after the accumulators

    // n fastest: the two halves of a row block are adjacent
    const int n_work = (GROUPED ? tile_prefix[args.n_seg] : args.n_tiles_m) * args.n_tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                GroupedTile gt;
                ts_decode<GROUPED>(w, args, tile_prefix, gt);
                const int m0 = gt.m0, n0 = gt.n0;
                const int b_row0 = GROUPED ? gt.seg * args.b_seg_rows : 0;       // this segment's block of the stacked B
                for (int kb = 0; kb < args.k_blocks; ++kb) {
                    if (!GROUPED && args.prefetch > 0 && n0 == 0) {
                        // A tiles of this row block's later k-blocks / of this CTA's next row block into L2.
                        // (n fastest: a row block is visited n_tiles_n times in a row, prefetch it once)
                        int pk = kb + args.prefetch, pm0 = m0;
                        bool ok = pk < args.k_blocks;
                        if (!ok) {
                            const int wn = w + static_cast<int>(gridDim.x);
                            pk -= args.k_blocks;
                            pm0 = (wn / args.n_tiles_n) * kBM;
                            ok = wn < n_work && pk < args.k_blocks && pm0 != m0;
                        }
                        if (ok && pk < args.k_blocks_a1) tma_prefetch_2d(&tmap_a, pk * BK, pm0);
                    }
                    bar_wait(bar_empty(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    const uint32_t sb_hi = sa + kABytes;
                    const uint32_t sb_lo = sb_hi + kBBytes;
                    bar_expect_tx(bar_full(stage), kTxBytes);
                    // A = [a1 | a2] along K: the two products of a pair (agg W_l^T + x W_r^T) share one accumulator
                    if (kb < args.k_blocks_a1) tma_load_2d(sa, &tmap_a, kb * BK, m0, bar_full(stage));
                    else tma_load_2d(sa, &tmap_a2, (kb - args.k_blocks_a1) * BK, m0, bar_full(stage));
                    // B_SPLIT: tmap_b_hi describes the unsplit matrix; its tile lands in the hi buffer
                    if (B_MN) {
#pragma unroll
                        for (int s = 0; s < BN / 32; ++s) {
                            tma_load_2d(sb_hi + s * kSlab, &tmap_b_hi, n0 + 32 * s, b_row0 + kb * BK, bar_full(stage));
                            if (!B_SPLIT) tma_load_2d(sb_lo + s * kSlab, &tmap_b_lo, n0 + 32 * s, b_row0 + kb * BK, bar_full(stage));
                        }
                    } else {
                        tma_load_2d(sb_hi, &tmap_b_hi, kb * BK, b_row0 + n0, bar_full(stage));
                        if (!B_SPLIT) tma_load_2d(sb_lo, &tmap_b_lo, kb * BK, b_row0 + n0, bar_full(stage));
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: the loop runs warp-wide, one elected lane issues (see elect_one in gemm_tf32x3.cu)
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            bar_wait(bar_tempty(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + static_cast<uint32_t>(acc * BN);
            uint32_t accumulate = 0;
            for (int kb = 0; kb < args.k_blocks; ++kb) {
                bar_wait(bar_full(stage), phase);
                bar_wait(bar_split(stage), phase);
                tc_fence_after();
                const uint32_t sb_hi = smem_base + stage * kStageBytes + kABytes;
                const uint64_t b_hi0 = smem_desc<B_MN, BK>(sb_hi, kSlab);
                const uint64_t b_lo0 = smem_desc<B_MN, BK>(sb_hi + kBBytes, kSlab);
                const uint32_t a_hi = tmem_a0 + static_cast<uint32_t>(stage) * kASlotCols;
                const uint32_t a_lo = a_hi + BK;
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < BK / 8; ++j) {
                        const uint64_t bo = static_cast<uint64_t>((B_MN ? j * 1024u : j * 32u) >> 4);
                        umma_tf32_ts(d, a_lo + j * 8, b_hi0 + bo, kIdesc, accumulate);
                        umma_tf32_ts(d, a_hi + j * 8, b_lo0 + bo, kIdesc, 1u);
                        umma_tf32_ts(d, a_hi + j * 8, b_hi0 + bo, kIdesc, 1u);
                        accumulate = 1u;
                    }
                    umma_commit(bar_empty(stage));
                }
                accumulate = 1u;
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(bar_tfull(acc));
            __syncwarp();
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
    } else if ((warp >= 4 && warp < 8) || warp >= 12) {
        // ===== splitter: smem row m (K-major, 128B swizzle) -> (hi, lo) -> TMEM lane m =====
        // Two sets of four warps (4-7 and 12-15) take alternate k-blocks, so that one set's chain (mbarrier
        // wait -> LDS -> split -> tcgen05.st -> wait::st -> arrive) overlaps the other's.  Measured perf-neutral
        // (profiles/r1_gemm_ts.md): the kernel sits at 70 % tensor-pipe activity with or without it.
        const int set = warp >= 12 ? 1 : 0;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            for (int kb = 0; kb < args.k_blocks; ++kb, ++it) {
                if ((it & 1) != set) {
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    continue;
                }
                bar_wait(bar_full(stage), phase);
                const unsigned char* row = smem_gen + stage * kStageBytes + m * 128;
                const uint32_t t_hi = tmem_a0 + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(stage) * kASlotCols;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int chunk = half * 4 + c;                        // 16-byte chunk = 4 consecutive k
                        const float4 v = *reinterpret_cast<const float4*>(row + ((chunk ^ (m & 7)) << 4));
                        const float h0 = rn_tf32(v.x), h1 = rn_tf32(v.y), h2 = rn_tf32(v.z), h3 = rn_tf32(v.w);
                        hi[c * 4 + 0] = __float_as_uint(h0); lo[c * 4 + 0] = __float_as_uint(v.x - h0);
                        hi[c * 4 + 1] = __float_as_uint(h1); lo[c * 4 + 1] = __float_as_uint(v.y - h1);
                        hi[c * 4 + 2] = __float_as_uint(h2); lo[c * 4 + 2] = __float_as_uint(v.z - h2);
                        hi[c * 4 + 3] = __float_as_uint(h3); lo[c * 4 + 3] = __float_as_uint(v.w - h3);
                    }
                    tmem_st16(t_hi + half * 16, hi);
                    tmem_st16(t_hi + BK + half * 16, lo);
                }
                if (B_SPLIT) {
                    // raw B tile -> (hi in place, lo next to it): element-wise, so the swizzled layout is kept
                    unsigned char* bh = smem_gen + stage * kStageBytes + kABytes;
                    const int tid = m;
#pragma unroll
                    for (int off = tid * 16; off < static_cast<int>(kBBytes); off += 128 * 16) {
                        const float4 v = *reinterpret_cast<const float4*>(bh + off);
                        const float4 h = make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w));
                        *reinterpret_cast<float4*>(bh + off) = h;
                        *reinterpret_cast<float4*>(bh + kBBytes + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> MMA (async proxy) reads
                }
                tmem_st_wait();
                tc_fence_before();
                bar_arrive(bar_split(stage));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            GroupedTile gt;
            ts_decode<GROUPED>(w, args, tile_prefix, gt);
            const int64_t m0 = gt.m0;
            const int n0 = gt.n0;
            const bool partial = GROUPED && gt.rows < kBM;      // last tile of a segment: rows beyond it belong to the next one
            bar_wait(bar_tfull(acc), acc_phase);
            tc_fence_after();
            // TMEM -> registers -> swizzled smem box -> TMA store: every global write is a full 128-byte
            // line issued by the copy engine (the thread-per-row direct stores cost 29 % of the kernel).
            const uint32_t my_stage = staging + static_cast<uint32_t>(q) * 8192u;
            const bool second = (w % args.n_tiles_n) >= args.n_tiles_c1;                // which of the two outputs
            const CUtensorMap* cmap = second ? &tmap_c2 : &tmap_c;
            const int cn0 = second ? n0 - args.n_tiles_c1 * BN : n0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const uint32_t buf = my_stage + static_cast<uint32_t>((c0 >> 5) & 1) * 4096u;
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + c0), r);
                if (args.bias) {                                                      // fused epilogue: + bias[n] (, ReLU)
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(args.bias + n0 + c0 + j));
                }
                if (args.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.0f));
                }
                if (partial) {
                    // row-masked direct stores (a thread owns one row of the 32 x 32 box)
                    const int row = q * 32 + lane;
                    if (row < gt.rows) {
                        float* dst = args.c + (m0 + row) * args.ldc + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                    continue;
                }
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this box's previous store has read it
                __syncwarp();
                if (!(args.debug & 4)) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        sts16(buf + static_cast<uint32_t>(lane) * 128u + static_cast<uint32_t>((c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1],
                              r[4 * c + 2], r[4 * c + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tma_store_2d(cmap, cn0 + c0, static_cast<int>(m0) + q * 32, buf);
                }
            }
            tc_fence_before();
            bar_arrive(bar_tempty(acc));
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
    }
    if (warp >= 8 && warp < 12 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all output boxes written
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <bool B_MN, bool B_SPLIT, bool GROUPED = false>
static int launch_gemm_ts(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tbh, const CUtensorMap& tbl,
                          const CUtensorMap& tc, const CUtensorMap& tc2, const GemmArgs& args, cudaStream_t stream) {
    constexpr size_t smem = kTsStages * (kBM * 32 * 4 + 2 * kTsBN * 32 * 4) + 4 * 2 * 4096 + 256 + 1024;
    auto kfn = gemm_tf32x3_ts_kernel<B_MN, B_SPLIT, GROUPED>;
    B200MP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int n_work = args.n_tiles_m * args.n_tiles_n;
    const int grid = n_work < num_sms() ? n_work : num_sms();
    kfn<<<grid, kGemmThreadsTs, smem, stream>>>(ta, ta2, tbh, tbl, tc, tc2, args);
    B200MP_LAUNCH_CHECK();
    return B200MP_OK;
}

}  // namespace b200mp
