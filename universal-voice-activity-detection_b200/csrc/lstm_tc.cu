// tcgen05 LSTM recurrence (one layer, both directions) for sm_100a.
//
// nn.LSTM semantics (PyanNet2.py:95,170): gates = xg_t + W_hh h_{t-1}; i,f,o = sigmoid, g = tanh;
// c_t = f*c_{t-1} + i*g; h_t = o*tanh(c_t).  xg (input projection + both biases) comes from gemm_tc.cu.
//
// Mapping: the GATE dimension is the MMA M (4 blocks of 128 rows = i, f, g, o of the 128 hidden
// units), the BATCH is the MMA N.  So unit u of every gate lands in TMEM lane u, and the thread that
// owns lane u reads i,f,g,o of a cell from four column ranges of its own lane -- no cross-thread
// exchange.  One CTA = one direction x 64 sequences:
//   * W_hh (512 x 128 fp16, 128 KB) is the A operand, resident in shared memory for all T steps
//     (8 K-major 128B-swizzled tiles written once by TMA);
//   * h_{t-1} is the B operand: a [64 x 128] K-major swizzled tile that the pointwise threads
//     write directly in operand layout, as fp16 hi and lo planes (two accumulating MMAs: the
//     recurrence sees h to ~22 bits, see DESIGN.md "precision");
//   * the 64 columns are two independent halves ping-ponged between two warpgroups: while
//     warpgroup 0 does the pointwise update of half 0, the tensor core runs half 1's MMAs, so the
//     MMA latency hides under the MUFU-bound cell update (5 ex2 + 2 rcp per cell);
//   * cell state lives in registers (fp32); xg for the next 8 columns is prefetched from HBM while
//     the current 8 are computed.
// Per step and CTA: 2 x 64 tcgen05.mma (M128 N32 K16), 8192 cells.  Algorithmic FLOPs:
// 2*128*512 per (sequence, frame, direction).
#include "kernels.cuh"
#include "tc05.cuh"

namespace b200vad {

using namespace tc;

constexpr int LNB = 64;                    // sequences per CTA
constexpr int LHALF = 32;                  // columns per warpgroup
constexpr int LCH = 8;                     // columns per inner chunk
constexpr int LTC_THREADS = 288;           // 2 warpgroups + 1 control warp
constexpr int W_TILE = 128 * 64 * 2;       // 16 KB
constexpr int H_TILE = LNB * 64 * 2;       // 8 KB

struct LstmTcParams {
    const float* xg;       // [B][T][2][512]
    __half* y_hi;          // [B][T][256] (planes mode) or null
    __half* y_lo;
    float* y_f32;          // [B][T][256] (fp32 mode) or null
    int B, T;
};

__device__ __forceinline__ float ex2f(float x) { return fast_ex2(x); }

template <bool F32OUT>
__global__ void __launch_bounds__(LTC_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tm_whh, LstmTcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t w_base = smem_base;                       // [4 gates][2 kb] tiles 128 x 64
    const uint32_t h_base = w_base + 8 * W_TILE;             // [hi, lo][2 kb] tiles 64 x 64
    const uint32_t bar_base = h_base + 4 * H_TILE;
    const uint32_t bar_w = bar_base;
    auto bar_h_ready = [&](int h) { return bar_base + 8 + 8 * h; };
    auto bar_acc_ready = [&](int h) { return bar_base + 24 + 8 * h; };
    const uint32_t tmem_slot = bar_base + 40;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int b0 = blockIdx.x * LNB;
    const int T = p.T;

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int h = 0; h < 2; ++h) { mbar_init(bar_h_ready(h), 128); mbar_init(bar_acc_ready(h), 1); }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 8) {
        // ===================== control warp: weight load + MMA issue =====================
        if (elect_one()) {
            mbar_expect_tx(bar_w, 8 * W_TILE);
            for (int q = 0; q < 4; ++q)
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(w_base + (q * 2 + kb) * W_TILE, &tm_whh, kb * 64, dir * kGates + q * kHidden, bar_w);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            constexpr uint32_t idesc = idesc_f16(128, LHALF);
            for (int s = 0; s < T; ++s) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(bar_h_ready(h), s & 1);
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t d = tmem_base + q * LNB + h * LHALF;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t da = smem_desc_sw128(w_base + (q * 2 + kb) * W_TILE + k * 32);
                                const uint32_t hb = h_base + kb * H_TILE + h * (LHALF * 128) + k * 32;
                                mma_f16(d, da, smem_desc_sw128(hb + 2 * H_TILE), idesc, (kb | k) != 0);     // h_lo first
                                mma_f16(d, da, smem_desc_sw128(hb), idesc, 1);                            // h_hi
                            }
                    }
                    mma_commit(bar_acc_ready(h));
                }
            }
        }
    } else {
        // ===================== pointwise warpgroups =====================
        const int wg = warp >> 2;                       // half owned by this warpgroup
        const int u = (warp & 3) * 32 + lane;           // hidden unit == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        // zero this warpgroup's half of the h tiles (h_{-1} = 0): rows [wg*32, wg*32+32) of the 4 tiles
        for (int i = threadIdx.x & 127; i < 4 * LHALF * 128 / 16; i += 128) {
            int tile = i / (LHALF * 8), rem = i % (LHALF * 8);
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(h_base + tile * H_TILE + wg * LHALF * 128 + rem * 16), "r"(0u) : "memory");
        }
        fence_proxy_async();
        mbar_arrive(bar_h_ready(wg));

        // smem byte offset of (row n, unit u) inside a tile, without the row term: chunk swizzle needs n & 7
        const uint32_t kb_off = (u >> 6) * H_TILE;
        const uint32_t uc = (u & 63) >> 3, ub = (u & 7) * 2;
        float c[LHALF];
#pragma unroll
        for (int j = 0; j < LHALF; ++j) c[j] = 0.f;

        const int64_t row_stride = (int64_t)T * 2 * kGates;      // floats between consecutive sequences in xg
        const float* xg_u = p.xg + (int64_t)dir * kGates + u;
        auto xg_ptr = [&](int col, int t) {
            int b = min(b0 + wg * LHALF + col, p.B - 1);
            return xg_u + (int64_t)b * row_stride + (int64_t)t * 2 * kGates;
        };
        float xr[4][LCH], xn[4][LCH];
        auto prefetch = [&](float (&dst)[4][LCH], int ch, int t) {
#pragma unroll
            for (int j = 0; j < LCH; ++j) {
                const float* q = xg_ptr(ch * LCH + j, t);
#pragma unroll
                for (int g = 0; g < 4; ++g) dst[g][j] = __ldg(q + g * kHidden);
            }
        };
        prefetch(xr, 0, dir == 0 ? 0 : T - 1);

        for (int s = 0; s < T; ++s) {
            const int t = dir == 0 ? s : T - 1 - s;
            const int tn = dir == 0 ? t + 1 : t - 1;
            mbar_wait(bar_acc_ready(wg), s & 1);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < LHALF / LCH; ++ch) {
                float a[4][LCH];
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld8(lane_addr + g * LNB + wg * LHALF + ch * LCH, a[g]);
                // prefetch the next chunk's xg (next step's first chunk after the last one)
                if (ch + 1 < LHALF / LCH) prefetch(xn, ch + 1, t);
                else if (s + 1 < T) prefetch(xn, 0, tn);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < LCH; ++j) {
                    const int col = ch * LCH + j;
                    const float L2E = 1.4426950408889634f;
                    float gi = fminf(fmaxf(a[0][j] + xr[0][j], -20.f), 20.f);
                    float gf = fminf(fmaxf(a[1][j] + xr[1][j], -20.f), 20.f);
                    float gg = fminf(fmaxf(a[2][j] + xr[2][j], -10.f), 10.f);
                    float go = fminf(fmaxf(a[3][j] + xr[3][j], -20.f), 20.f);
                    float ei = ex2f(-L2E * gi), ef = ex2f(-L2E * gf), eo = ex2f(-L2E * go), eg = ex2f(2.f * L2E * gg);
                    // c' = c/(1+ef) + (eg-1)/((1+ei)(eg+1))  with one reciprocal
                    float di = 1.f + ei, df = 1.f + ef, dg = eg + 1.f;
                    float dig = di * dg;
                    float num = fmaf(c[col], dig, (eg - 1.f) * df);
                    float cn = num * fast_rcp(df * dig);
                    c[col] = cn;
                    float cc = fminf(fmaxf(cn, -10.f), 10.f);
                    float ec = ex2f(2.f * L2E * cc);
                    float hv = (ec - 1.f) * fast_rcp((1.f + eo) * (ec + 1.f));
                    __half hh = __float2half_rn(hv);
                    __half hl = __float2half_rn(hv - __half2float(hh));
                    const int n = wg * LHALF + col;
                    const uint32_t off = kb_off + n * 128 + ((uc ^ (n & 7)) << 4) + ub;
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(h_base + off), "h"(__half_as_ushort(hh)) : "memory");
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(h_base + 2 * H_TILE + off), "h"(__half_as_ushort(hl)) : "memory");
                    const int b = b0 + n;
                    if (b < p.B) {
                        const int64_t o = ((int64_t)b * T + t) * (2 * kHidden) + dir * kHidden + u;
                        if (F32OUT) p.y_f32[o] = hv;
                        else { p.y_hi[o] = hh; p.y_lo[o] = hl; }
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int j = 0; j < LCH; ++j) xr[g][j] = xn[g][j];
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_h_ready(wg));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<256>(tmem_base);
}

// whh: [2][512][128] fp16 (gate-major rows, as nn.LSTM stores weight_hh).  Exactly one of
// (y_hi, y_lo) / y_f32 is written.
int lstm_tc_launch(const float* xg, const __half* whh, __half* y_hi, __half* y_lo, float* y_f32, int B, int T, cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, whh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, kHidden, 2 * kGates, kHidden * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    LstmTcParams p{xg, y_hi, y_lo, y_f32, B, T};
    const int smem = 8 * W_TILE + 4 * H_TILE + 1024 + 256;
    dim3 grid((B + LNB - 1) / LNB, 2);
    prof_begin(0, st);
    if (y_f32) {
        static bool attr = false;
        if (!attr) { B200VAD_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
        lstm_tc_kernel<true><<<grid, LTC_THREADS, smem, st>>>(tm, p);
    } else {
        static bool attr = false;
        if (!attr) { B200VAD_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
        lstm_tc_kernel<false><<<grid, LTC_THREADS, smem, st>>>(tm, p);
    }
    prof_end(0, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
