// tcgen05 projection GEMM:  C[M,N] (fp32) = A_hi.W_hi^T + A_lo.W_hi^T (+ A_hi.W_lo^T) + bias
//
// Computes the LSTM input projections x_t.W_ih^T + b_ih + b_hh for all timesteps and both
// directions (nn.LSTM of PyanNet2.py:95,170) in split precision: activations arrive as two fp16
// planes (hi, lo: hi + lo = the fp32 value to ~22 bits), weights as fp16 hi (+ lo for layer 0,
// whose inputs are un-normalised log-mel values), accumulation in fp32 in TMEM.
//
// Persistent, warp-specialised, one CTA per SM (grid = multiple of the number of 256-column
// ranges): CTA c owns column range c % n_ranges and keeps that range's weights RESIDENT in shared
// memory for its whole life (W is the B operand: K-major, 128B-swizzled, written once by TMA), and
// streams 128-row A tiles.  Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (single thread),
// warps 2-5 = epilogue.  The 128x256 fp32 accumulator is double-buffered in TMEM (2 x 256 columns),
// so the epilogue of tile i (tcgen05.ld -> +bias -> swizzled smem -> TMA store) overlaps the MMAs of
// tile i+1.  The CTAs that share an A tile (the n_ranges column ranges of one row block) run
// concurrently, so A is fetched from HBM once and hits L2 for the others.
// Roofline: 2 x 13.4 GB of fp32 output per layer at B*T = 3.28 M rows -> HBM-write bound (~2 ms).
#include "kernels.cuh"
#include "tc05.cuh"

namespace b200vad {

using namespace tc;

constexpr int TBM = 128;                 // rows per tile (UMMA M)
constexpr int TBK = 64;                  // k per smem tile (128 bytes of fp16 = one swizzle atom row)
constexpr int A_TILE_BYTES = TBM * TBK * 2;          // 16 KB
constexpr int EPI_CHUNK = 32;                          // fp32 columns per TMA store (128 bytes)
constexpr int EPI_BYTES = TBM * EPI_CHUNK * 4;         // 16 KB
constexpr int GEMM_TC_THREADS = 192;

struct GemmTcParams {
    const float* bias;       // [N]
    int num_m_tiles;
    int n_ranges;            // N / BN
    int kb;                  // k-blocks of 64
    int nw;                  // weight planes resident: 1 (hi) or 2 (hi, lo)
    int stages;              // A pipeline depth
};

template <int BN>
__global__ void __launch_bounds__(GEMM_TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const __grid_constant__ CUtensorMap tm_c, GemmTcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    constexpr int W_TILE_BYTES = BN * TBK * 2;
    const uint32_t w_base = smem_base;                                        // [nw][kb] tiles of BN x 64
    const uint32_t a_base = w_base + p.nw * p.kb * W_TILE_BYTES;              // [stages][hi, lo] tiles of 128 x 64
    const uint32_t epi_base = a_base + p.stages * 2 * A_TILE_BYTES;          // [2] staging tiles 128 x 32 fp32
    const uint32_t bar_base = epi_base + 2 * EPI_BYTES;
    // barriers: w_full | a_full[8] | a_empty[8] | acc_full[2] | acc_empty[2]
    const uint32_t bar_w = bar_base;
    auto bar_a_full = [&](int s) { return bar_base + 8 + 8 * s; };
    auto bar_a_empty = [&](int s) { return bar_base + 8 + 64 + 8 * s; };
    auto bar_acc_full = [&](int b) { return bar_base + 8 + 128 + 8 * b; };
    auto bar_acc_empty = [&](int b) { return bar_base + 8 + 144 + 8 * b; };
    const uint32_t tmem_slot = bar_base + 8 + 160;
    const uint32_t bias_smem = bar_base + 256;                                // BN floats (16-byte aligned)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int range = blockIdx.x % p.n_ranges;
    const int n0 = range * BN;
    const int tile0 = blockIdx.x / p.n_ranges;
    const int tile_step = gridDim.x / p.n_ranges;

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), 128); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BN; i += 128) {
            float b = p.bias ? p.bias[n0 + i] : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + 4 * i), "f"(b) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_w_hi); tma_prefetch_desc(&tm_c);
            mbar_expect_tx(bar_w, p.nw * p.kb * W_TILE_BYTES);
            for (int w = 0; w < p.nw; ++w)
                for (int kb = 0; kb < p.kb; ++kb)
                    tma_load_2d(w_base + (w * p.kb + kb) * W_TILE_BYTES, w == 0 ? &tm_w_hi : &tm_w_lo, kb * TBK, n0, bar_w);
            int s = 0;
            uint32_t ph = 0;
            for (int t = tile0; t < p.num_m_tiles; t += tile_step) {
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_a_full(s), 2 * A_TILE_BYTES);
                    tma_load_2d(a_base + (2 * s) * A_TILE_BYTES, &tm_a_hi, kb * TBK, t * TBM, bar_a_full(s));
                    tma_load_2d(a_base + (2 * s + 1) * A_TILE_BYTES, &tm_a_lo, kb * TBK, t * TBM, bar_a_full(s));
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(TBM, BN);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = tile0; t < p.num_m_tiles; t += tile_step, ++it) {
                const int ab = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(bar_acc_empty(ab), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ab * BN;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = a_base + (2 * s) * A_TILE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t w_hi = w_base + kb * W_TILE_BYTES, w_lo = w_base + (p.kb + kb) * W_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < TBK / 16; ++k) {
                        const uint64_t da_hi = smem_desc_sw128(a_hi + k * 32), da_lo = smem_desc_sw128(a_lo + k * 32);
                        const uint64_t dw_hi = smem_desc_sw128(w_hi + k * 32);
                        mma_f16(d_tmem, da_lo, dw_hi, idesc, (kb | k) != 0);          // small terms first
                        if (p.nw == 2) mma_f16(d_tmem, da_hi, smem_desc_sw128(w_lo + k * 32), idesc, 1);
                        mma_f16(d_tmem, da_hi, dw_hi, idesc, 1);
                    }
                    mma_commit(bar_a_empty(s));                                 // frees the A stage when the MMAs retire
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                mma_commit(bar_acc_full(ab));                                   // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;                                          // row of the tile == TMEM lane
        const int et = threadIdx.x - 64;                                        // 0..127
        int it = 0;
        int chunk_ctr = 0;
        for (int t = tile0; t < p.num_m_tiles; t += tile_step, ++it) {
            const int ab = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(bar_acc_full(ab), acc_ph);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / EPI_CHUNK; ++c, ++chunk_ctr) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BN + c * EPI_CHUNK, v);
                tmem_ld_wait();
                if (c == BN / EPI_CHUNK - 1) {
                    tc_fence_before();
                    mbar_arrive(bar_acc_empty(ab));                             // 128 arrivals release the accumulator
                }
                const int buf = chunk_ctr & 1;
                // the TMA store that last read staging[buf] (two chunks ago) must have finished reading
                if (et == 0) tma_store_wait_read<1>();
                named_bar_sync(1, 128);
                const uint32_t stage = epi_base + buf * EPI_BYTES + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o;
                    float4 bb;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                                 : "r"(bias_smem + 4 * (c * EPI_CHUNK + 4 * j)));
                    o.x = v[4 * j] + bb.x; o.y = v[4 * j + 1] + bb.y; o.z = v[4 * j + 2] + bb.z; o.w = v[4 * j + 3] + bb.w;
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + ((j ^ (row & 7)) << 4)), "f"(o.x), "f"(o.y),
                                 "f"(o.z), "f"(o.w) : "memory");
                }
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (et == 0) {
                    tma_store_2d(&tm_c, n0 + c * EPI_CHUNK, t * TBM, epi_base + buf * EPI_BYTES);
                    tma_store_commit();
                }
            }
        }
        if (et == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------- fp32 -> fp16 (hi, lo) planes
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, int64_t n, __half* __restrict__ hi,
                                                           __half* __restrict__ lo) {
    int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
        __half h[4], l[4];
        split_f16(v.x, h[0], l[0]); split_f16(v.y, h[1], l[1]); split_f16(v.z, h[2], l[2]); split_f16(v.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<uint2*>(h);
        *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<uint2*>(l);
    } else {
        for (; i < n; ++i) split_f16(x[i], hi[i], lo[i]);
    }
}
int split_planes_launch(const float* x, int64_t n, __half* hi, __half* lo, cudaStream_t st) {
    if (n <= 0) return B200VAD_OK;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(hi) & 7) || (reinterpret_cast<uintptr_t>(lo) & 7)) {
        set_error("split_planes: unaligned pointer");
        return B200VAD_EINVAL;
    }
    split_planes_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(x, n, hi, lo);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static size_t dtype_bytes(CUtensorMapDataType t) { return t == CU_TENSOR_MAP_DATA_TYPE_FLOAT16 ? 2 : 4; }

int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, uint64_t inner, uint64_t outer,
                 uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return B200VAD_ESTATE;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    (void)dtype_bytes;
    CUresult r = enc(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(2d) failed: %d (inner %llu outer %llu pitch %llu box %u x %u)", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes, box_inner, box_outer);
        return B200VAD_ECUDA;
    }
    return B200VAD_OK;
}
int make_tmap_3d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return B200VAD_ESTATE;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
        return B200VAD_ECUDA;
    }
    return B200VAD_OK;
}

// ---------------------------------------------------------------- launcher
// a_hi/a_lo: [M, K] fp16 (row pitch lda elements, lda % 8 == 0);  w_hi/w_lo: [N, Kp] fp16 (Kp % 64 == 0, zero padded);
// c: [M, N] fp32 (ldc % 4 == 0).  N must be a multiple of 128.  w_lo may be null (2-term product).
int gemm_tc_launch(const __half* a_hi, const __half* a_lo, int64_t lda, int64_t M, int K, const __half* w_hi, const __half* w_lo,
                   int Kp, int N, const float* bias, float* c, int64_t ldc, int num_sms, cudaStream_t st) {
    if (M <= 0) return B200VAD_OK;
    if (N % 128 != 0 || Kp % 64 != 0 || lda % 8 != 0 || ldc % 4 != 0 || K > Kp) {
        set_error("gemm_tc: unsupported shape N=%d K=%d Kp=%d lda=%lld ldc=%lld", N, K, Kp, (long long)lda, (long long)ldc);
        return B200VAD_EINVAL;
    }
    GemmTcParams p;
    p.bias = bias;
    p.num_m_tiles = (int)((M + TBM - 1) / TBM);
    p.kb = Kp / TBK;
    p.nw = w_lo ? 2 : 1;
    // widest column range whose weights (all planes, all of K) stay resident and leave >= 2 A stages
    const int max_smem = 227 * 1024;
    int BN = 0, fixed = 0, w_bytes = 0;
    for (int cand : {256, 128}) {
        if (N % cand) continue;
        w_bytes = p.nw * p.kb * cand * TBK * 2;
        fixed = w_bytes + 2 * EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/ + cand * 4 + 256;
        if ((max_smem - fixed) / (2 * A_TILE_BYTES) >= 2) { BN = cand; break; }
    }
    if (!BN) {
        set_error("gemm_tc: weights (%d planes x %d x K=%d) do not fit in shared memory", p.nw, N, Kp);
        return B200VAD_EINVAL;
    }
    p.n_ranges = N / BN;
    p.stages = (max_smem - fixed) / (2 * A_TILE_BYTES);
    if (p.stages > 8) p.stages = 8;
    const int smem = fixed + p.stages * 2 * A_TILE_BYTES;
    CUtensorMap tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo, tm_c;
    int rc;
    if ((rc = make_tmap_2d(&tm_a_hi, a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, M, lda * 2, TBK, TBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d(&tm_a_lo, a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, M, lda * 2, TBK, TBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d(&tm_w_hi, w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Kp, N, (uint64_t)Kp * 2, TBK, BN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d(&tm_w_lo, w_lo ? w_lo : w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Kp, N, (uint64_t)Kp * 2, TBK, BN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d(&tm_c, c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, N, M, ldc * 4, EPI_CHUNK, TBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    int grid = (num_sms / p.n_ranges) * p.n_ranges;
    if (grid < p.n_ranges) grid = p.n_ranges;
    int max_grid = p.num_m_tiles * p.n_ranges;
    if (grid > max_grid) grid = max_grid;
    prof_begin(1, st);
    if (BN == 256) {
        static bool attr = false;
        if (!attr) { B200VAD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); attr = true; }
        gemm_tc_kernel<256><<<grid, GEMM_TC_THREADS, smem, st>>>(tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo, tm_c, p);
    } else {
        static bool attr = false;
        if (!attr) { B200VAD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); attr = true; }
        gemm_tc_kernel<128><<<grid, GEMM_TC_THREADS, smem, st>>>(tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo, tm_c, p);
    }
    prof_end(1, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
