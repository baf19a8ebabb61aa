// tcgen05 LSTM recurrence (one layer, both directions) for sm_100a.
//
// nn.LSTM semantics (PyanNet2.py:95,170): gates = xg_t + W_hh h_{t-1}; i,f,o = sigmoid, g = tanh;
// c_t = f*c_{t-1} + i*g; h_t = o*tanh(c_t).  xg (input projection + both biases) comes from gemm_tc.cu.
//
// Mapping: the GATE dimension is the MMA M (4 blocks of 128 rows = i, f, g, o of the 128 hidden
// units), the BATCH is the MMA N.  So unit u of every gate lands in TMEM lane u, and the thread that
// owns lane u reads i,f,g,o of a cell from four column ranges of its own lane -- no cross-thread
// exchange.  One CTA = one direction x 64 sequences, 18 warps:
//   * W_hh (512 x 128 fp16) is the A operand and lives in TENSOR MEMORY for all T steps (256
//     columns: lane = unit, two K-elements per 32-bit column), next to the 256 accumulator columns:
//     an MMA re-reads no weights from shared memory (in SS mode every M128 K16 MMA re-reads a 4 KB
//     A slice, which made the tensor pipe the serial bottleneck -- profiles/r01_lstm_ablation.md);
//   * h_{t-1} is the B operand: a [64 x 128] K-major 128B-swizzled tile that the pointwise threads
//     write directly in operand layout, as fp16 hi and lo planes (two accumulating MMAs: the
//     recurrence sees h to ~22 bits, DESIGN.md "precision"); the tiles are double-buffered by step
//     parity and are also the source of the TMA stores of the layer output planes y_hi / y_lo;
//   * the 64 columns are two independent halves: while the pointwise warpgroups of half 0 update
//     their cells, the tensor core (issued by warp 16) runs half 1's MMAs.  Each half is shared by
//     two warpgroups of 16 columns (4 warps per scheduler on the ex2/rcp dependency chains);
//     warps 17..20 are the xg TMA producers, one per pointwise warpgroup;
//   * cell state c lives in shared memory (fp32, conflict-free 8-byte accesses: two columns of a unit are adjacent),
//     so the column loop is a real loop; the cell update runs on packed fp32 pairs (FADD2 / FMUL2 / FFMA2), two
//     columns per issue slot -- only the MUFU ops, the clamps and the fp16 conversions are scalar;
//   * xg is the only HBM read stream (4 KB per sequence-step-direction): warp 17 feeds it through a
//     4-stage TMA ring per warpgroup (4 columns x 512 gates per stage, 128 KB in flight per SM); xg is stored
//     step-blocked: the 128 KB a CTA consumes per step are one contiguous record [gate][column group][unit][4 columns],
//     so a stage is one TMA box and a thread reads its 4 columns of a gate with one 16-byte shared-memory load
//     (an L2 prefetch ahead of the ring measured no gain -- profiles/r01_lstm_ablation.md).
// Per step and CTA: 2 x 64 tcgen05.mma (M128 N32 K16, A in TMEM), 8192 cells, 5 ex2 + 2 rcp per cell.
// Algorithmic FLOPs: 2*128*512 per (sequence, frame, direction); algorithmic HBM bytes per
// (sequence, frame, direction): 2048 (xg read) + 512 (y planes written).
#include "kernels.cuh"
#include "tc05.cuh"
#include <stdlib.h>

namespace b200vad {

using namespace tc;

constexpr int XBLK = 64;                   // sequences per xg record block (gemm_ts mode 3 layout)
// The 64 columns are split into PARTS (2 or 4) independently pipelined groups of 64 / PARTS columns: while the
// pointwise warpgroups of one part update their cells, the tensor core runs another part's MMAs.
constexpr int LPARTS_DEFAULT = 4;          // pipelined column groups (B200VAD_LSTM_PARTS=2 selects the two-half schedule)
constexpr int LWG = 4;                     // pointwise warpgroups
constexpr int LCH = 4;                     // columns per ring stage / inner chunk
constexpr int LSTAGES = 4;                 // xg ring depth per warpgroup
constexpr int LTC_THREADS = (LWG * 4 + 1 + LWG) * 32;   // 16 pointwise warps + MMA warp + one xg producer warp per warpgroup = 672
constexpr int X_STAGE = 4 * kHidden * LCH * 4;   // one TMA box: 4 gates x 128 units x 4 columns (fp32) = 8 KB
// Sequences per CTA (NB): 64 for throughput (4096-row batches fill 128 SMs), 16 for latency (small batches such as the
// 256 streams of the streaming mode spread over 4x as many SMs and a step shrinks to one MMA group + one 4-column chunk
// per warpgroup).  Derived sizes:
template <int NB> struct LstmGeom {
    static constexpr int LWCOLS = NB / LWG;             // columns per warpgroup (16 / 4)
    static constexpr int LNCH = LWCOLS / LCH;           // chunks per step and warpgroup (4 / 1)
    static constexpr int H_TILE = NB * 64 * 2;          // one [NB x 64] fp16 operand tile (8 KB / 2 KB)
    static constexpr int C_BYTES = NB * kHidden * 4;    // cell state (32 KB / 8 KB)
    static constexpr int TMEM_W = 4 * NB;               // first TMEM column of W_hh (after the 4 x NB accumulator columns)
    static constexpr int SMEM = 8 * H_TILE + C_BYTES + LWG * LSTAGES * X_STAGE + 1024 + 512;
};

static int g_lstm_tile = 0;               // 0 = automatic, 16 / 64 = forced (b200vad_set_lstm_tile)
int lstm_tc_set_tile(int nb) {
    if (nb != 0 && nb != 16 && nb != 64) return B200VAD_EINVAL;
    g_lstm_tile = nb;
    return B200VAD_OK;
}

struct LstmTcParams {
    const __half* whh;     // [2][512][128]
    int B, T;
    int flags;             // debug ablations (B200VAD_LSTM_DEBUG): 1 = no xg, 2 = no MMAs, 4 = no h_lo MMAs, 8 = first reciprocal on the FMA pipe
    float h1_scale;        // first plane of h = fp16(h1_scale * h), second = fp16(h - first): 1 (hi / lo) or 1 - kPlaneScale
};

template <int PARTS, int NB>
__global__ void __launch_bounds__(LTC_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tm_xg, const __grid_constant__ CUtensorMap tm_yhi,
               const __grid_constant__ CUtensorMap tm_ylo, LstmTcParams p) {
    constexpr int LNB = NB;
    constexpr int LWCOLS = LstmGeom<NB>::LWCOLS, LNCH = LstmGeom<NB>::LNCH, H_TILE = LstmGeom<NB>::H_TILE;
    constexpr int C_BYTES = LstmGeom<NB>::C_BYTES, TMEM_W = LstmGeom<NB>::TMEM_W;
    static_assert(NB % (16 * PARTS) == 0 && LWCOLS % LCH == 0, "MMA N = NB / PARTS must be a multiple of 16");
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* const smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t h_base = smem_base;                       // [step parity][hi, lo][2 kb] tiles 64 x 64
    const uint32_t c_off = 8 * H_TILE;                       // cell state [64 cols][128 units] fp32
    const uint32_t x_off = c_off + C_BYTES;                  // [LWG][LSTAGES] xg stages
    const uint32_t bar_off = x_off + LWG * LSTAGES * X_STAGE;
    const uint32_t bar_base = smem_base + bar_off;
    constexpr int LPN = LNB / PARTS;                          // columns per part (MMA N)
    constexpr int WG_PER_PART = LWG / PARTS;
    auto bar_h_ready = [&](int h) { return bar_base + 8 * h; };
    auto bar_acc_ready = [&](int h) { return bar_base + 32 + 8 * h; };
    auto bar_x_full = [&](int wg, int st) { return bar_base + 64 + 8 * (wg * LSTAGES + st); };
    auto bar_x_empty = [&](int wg, int st) { return bar_base + 64 + 8 * (LWG * LSTAGES + wg * LSTAGES + st); };
    const uint32_t tmem_slot = bar_base + 64 + 8 * 2 * LWG * LSTAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int b0 = blockIdx.x * LNB;
    const int T = p.T;

    if (threadIdx.x == 0) {
        for (int h = 0; h < PARTS; ++h) { mbar_init(bar_h_ready(h), 128 * WG_PER_PART); mbar_init(bar_acc_ready(h), 1); }
        for (int g = 0; g < LWG; ++g)
            for (int st = 0; st < LSTAGES; ++st) { mbar_init(bar_x_full(g, st), 1); mbar_init(bar_x_empty(g, st), 4); }
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp >= 17) {
        // ===================== xg producers: warp 17 + g feeds warpgroup g =====================
        // (one warp each: four lanes of one warp spinning on different barriers serialise each other)
        if (elect_one() && !(p.flags & 1)) {
            const int wg = warp - 17;
            int st = 0;
            uint32_t ph = 0;
            for (int s = 0; s < T; ++s) {
                const int t = dir == 0 ? s : T - 1 - s;
#pragma unroll 1
                for (int ch = 0; ch < LNCH; ++ch) {
                    mbar_wait(bar_x_empty(wg, st), ph ^ 1);
                    const uint32_t dst = smem_base + x_off + (wg * LSTAGES + st) * X_STAGE;
                    mbar_expect_tx(bar_x_full(wg, st), X_STAGE);
                    // one box = 4 gates x (128 units x 4 columns) of this step's 128 KB record
                    tma_load_4d(dst, &tm_xg, 0, (b0 % XBLK) / LCH + wg * LNCH + ch, 0, ((b0 / XBLK) * 2 + dir) * T + t, bar_x_full(wg, st));
                    if (++st == LSTAGES) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 16) {
        // ===================== MMA issue + layer-output TMA stores =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(128, LPN);
            // iteration s: h_ready(h) phase s = "h_{s-1} is in tile buffer (s & 1)" (phase 0 = zero state + W_hh in TMEM)
            for (int s = 0; s <= T; ++s) {
                const uint32_t hbuf = h_base + (s & 1) * 4 * H_TILE;
#pragma unroll
                for (int h = 0; h < PARTS; ++h) {
                    mbar_wait(bar_h_ready(h), s & 1);
                    tc_fence_after();
                    if (s < T) {
                        if (!(p.flags & 2)) {
                            // k-major issue order (the four gate accumulators alternate; the descriptors of a k-step are built once)
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) {
                                const uint32_t hb = hbuf + (kk >> 2) * H_TILE + h * (LPN * 128) + (kk & 3) * 32;
                                const uint64_t d_lo = smem_desc_sw128(hb + 2 * H_TILE), d_hi = smem_desc_sw128(hb);
                                if (!(p.flags & 4)) {
#pragma unroll
                                    for (int q = 0; q < 4; ++q)                                                       // h_lo first
                                        mma_f16_ts(tmem_base + q * LNB + h * LPN, tmem_base + TMEM_W + q * 64 + kk * 8, d_lo, idesc, kk != 0);
                                }
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    mma_f16_ts(tmem_base + q * LNB + h * LPN, tmem_base + TMEM_W + q * 64 + kk * 8, d_hi, idesc,
                                               (p.flags & 4) ? (uint32_t)(kk != 0) : 1u);
                            }
                        }
                        // once acc_ready fires the pointwise warps rewrite buffer (s+1)&1, last read by the store group of
                        // iteration s-1 for this part: every group but the PARTS-1 most recent ones must have finished reading
                        tma_store_wait_read<PARTS - 1>();
                        mma_commit(bar_acc_ready(h));
                    }
                    if (s > 0) {
                        // h_{s-1} of this half -> y planes at time t(s-1); rows beyond B are clipped by TMA
                        const int t = dir == 0 ? s - 1 : T - s;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            tma_store_3d(&tm_yhi, dir * kHidden + kb * 64, t, b0 + h * LPN, hbuf + kb * H_TILE + h * (LPN * 128));
                            tma_store_3d(&tm_ylo, dir * kHidden + kb * 64, t, b0 + h * LPN, hbuf + (2 + kb) * H_TILE + h * (LPN * 128));
                        }
                        tma_store_commit();
                    }
                }
            }
            tma_store_wait_all<0>();
        }
    } else {
        // ===================== pointwise warpgroups =====================
        const int wg = warp >> 2;                       // columns [wg*16, wg*16+16)
        const int half = wg / WG_PER_PART;              // the part this warpgroup belongs to
        const int u = (warp & 3) * 32 + lane;           // hidden unit == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        // W_hh -> TMEM: warpgroup g stores gate g; lane u gets row (dir, g, u): 128 fp16 = 64 packed words
        {
            const uint4* wrow = reinterpret_cast<const uint4*>(p.whh + ((size_t)dir * kGates + wg * kHidden + u) * kHidden);
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                uint32_t r[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 v = __ldg(wrow + part * 4 + i);
                    r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
                }
                tmem_st16(lane_addr + TMEM_W + wg * 64 + part * 16, r);
            }
            tmem_st_wait();
        }
        // zero this warpgroup's rows of the 8 h tiles (h_{-1} = 0) and its cell-state columns
        for (int i = threadIdx.x & 127; i < 8 * LWCOLS * 128 / 16; i += 128) {
            int tile = i / (LWCOLS * 8), rem = i % (LWCOLS * 8);
            *reinterpret_cast<uint4*>(smem_gen + tile * H_TILE + wg * LWCOLS * 128 + rem * 16) = make_uint4(0, 0, 0, 0);
        }
        float* const cst = reinterpret_cast<float*>(smem_gen + c_off) + 2 * u;  // [column pair][unit][2]: + (col >> 1) * 256 + (col & 1)
#pragma unroll
        for (int j = 0; j < LWCOLS; j += 2)
            *reinterpret_cast<float2*>(cst + ((wg * LWCOLS + j) >> 1) * (2 * kHidden)) = make_float2(0.f, 0.f);
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(bar_h_ready(half));

        unsigned char* const h_hi0 = smem_gen + (u >> 6) * H_TILE + (u & 7) * 2;   // + row*128 + swizzled chunk
        const uint32_t uc = (u & 63) >> 3;
        const float* const xring = reinterpret_cast<const float*>(smem_gen + x_off + wg * LSTAGES * X_STAGE) + u * LCH;
        int st = 0;
        uint32_t xph = 0;
        const float L2E = 1.4426950408889634f;

        for (int s = 0; s < T; ++s) {
            unsigned char* const h_hi = h_hi0 + ((s + 1) & 1) * 4 * H_TILE;     // h_s goes to buffer (s+1)&1
            mbar_wait(bar_acc_ready(half), s & 1);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < LNCH; ++ch) {
                const int col0 = wg * LWCOLS + ch * LCH;
                float a[4][LCH], x[4][LCH];
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld4(lane_addr + g * LNB + col0, a[g]);
                if (p.flags & 1) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
#pragma unroll
                        for (int j = 0; j < LCH; ++j) x[g][j] = 0.1f * g;
                } else {
                    // xg chunk from the ring: stage = [gate 4][unit 128][column 4] -> one 16-byte load per gate
                    mbar_wait(bar_x_full(wg, st), xph);
                    const float4* xs = reinterpret_cast<const float4*>(xring + st * (X_STAGE / 4));
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 v = xs[g * 128];
                        x[g][0] = v.x; x[g][1] = v.y; x[g][2] = v.z; x[g][3] = v.w;
                    }
                }
                tmem_ld_wait();
                // two columns per issue slot (FADD2 / FMUL2 / FFMA2); MUFU, min and the fp16 conversions stay scalar
                const f32x2 one = pack2(1.f, 1.f), mone = pack2(-1.f, -1.f), k2 = pack2(2.f * L2E, 2.f * L2E);
#pragma unroll
                for (int jp = 0; jp < LCH / 2; ++jp) {
                    const int n = col0 + 2 * jp;
                    // pre-activations arrive multiplied by -log2 e (i, f, o) / 2 log2 e (g): the packed weights carry the scale;
                    // one-sided clamps keep every exponential finite (<= 2^29); exp(-inf) = 0 is exact
                    const f32x2 ei = ex2_clamped2(add2(pack2(a[0][2 * jp], a[0][2 * jp + 1]), pack2(x[0][2 * jp], x[0][2 * jp + 1])));
                    const f32x2 ef = ex2_clamped2(add2(pack2(a[1][2 * jp], a[1][2 * jp + 1]), pack2(x[1][2 * jp], x[1][2 * jp + 1])));
                    const f32x2 eg = ex2_clamped2(add2(pack2(a[2][2 * jp], a[2][2 * jp + 1]), pack2(x[2][2 * jp], x[2][2 * jp + 1])));
                    const f32x2 eo = ex2_clamped2(add2(pack2(a[3][2 * jp], a[3][2 * jp + 1]), pack2(x[3][2 * jp], x[3][2 * jp + 1])));
                    // c' = c/(1+ef) + (eg-1)/((1+ei)(eg+1))  with one reciprocal
                    const f32x2 di = add2(ei, one), df = add2(ef, one), dg = add2(eg, one);
                    const f32x2 dig = mul2(di, dg);
                    float2* const cptr = reinterpret_cast<float2*>(cst + (n >> 1) * (2 * kHidden));
                    const float2 cold = *cptr;
                    const f32x2 den = mul2(df, dig);
                    const f32x2 cn = mul2(fma2(pack2(cold.x, cold.y), dig, mul2(add2(eg, mone), df)), (p.flags & 8) ? rcp2_fma(den) : rcp2(den));
                    float cn0, cn1;
                    unpack2(cn, cn0, cn1);
                    *cptr = make_float2(cn0, cn1);
                    const f32x2 ec = ex2_clamped2(mul2(cn, k2));
                    const f32x2 hv2 = mul2(add2(ec, mone), rcp2(mul2(add2(eo, one), add2(ec, one))));
                    float hv0, hv1;
                    unpack2(hv2, hv0, hv1);
                    // planes: fp16(h1_scale * h) and the remainder (h1_scale = 1: hi / lo; 1 - 2^-6: the 2-MMA projection's split)
                    const __half hh0 = __float2half_rn(hv0 * p.h1_scale), hh1 = __float2half_rn(hv1 * p.h1_scale);
                    const __half hl0 = __float2half_rn(hv0 - __half2float(hh0)), hl1 = __float2half_rn(hv1 - __half2float(hh1));
                    unsigned char* dst0 = h_hi + n * 128 + ((uc ^ (n & 7)) << 4);
                    unsigned char* dst1 = h_hi + (n + 1) * 128 + ((uc ^ ((n + 1) & 7)) << 4);
                    *reinterpret_cast<__half*>(dst0) = hh0;
                    *reinterpret_cast<__half*>(dst0 + 2 * H_TILE) = hl0;
                    *reinterpret_cast<__half*>(dst1) = hh1;
                    *reinterpret_cast<__half*>(dst1 + 2 * H_TILE) = hl1;
                }
                if (!(p.flags & 1)) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_x_empty(wg, st));
                    if (++st == LSTAGES) { st = 0; xph ^= 1; }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_h_ready(half));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem_base);
}

// xg: step-blocked [ceil(B/64)][2][T][4][16][128][4] fp32 as gemm_ts_xg_launch writes it; whh: [2][512][128] fp16 (gate-major rows, as nn.LSTM stores weight_hh).
// Output: the layer output as fp16 planes y_hi / y_lo [B][T][256] (hi / lo split, or the scaled (x1, x2) split of the
// 2-MMA projection when scaled_planes is set; either way the planes sum to the fp32 value).
int lstm_tc_launch(const float* xg, const __half* whh, __half* y_hi, __half* y_lo, int B, int T, int scaled_planes, cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    CUtensorMap tm_x, tm_yh, tm_yl;
    // xg is step-blocked (gemm_ts mode 3): [sequence block][dir][t][gate 4][column group 16][unit 128][4] fp32.  Viewed as
    // 8-byte elements: 256 per (gate, column group) run of 2 KB; one box = the 4 gates of one column group of one step.
    const uint64_t nblk = (uint64_t)((B + XBLK - 1) / XBLK);
    const uint64_t xdims[4] = {256, 16, 4, nblk * 2 * (uint64_t)T};
    const uint64_t xpitch[3] = {2048, 32768, 131072};
    const uint32_t xbox[4] = {256, 1, 4, 1};
    static_assert(LCH == 4 && XBLK == 64, "xg layout assumes 4-column groups and 64-sequence blocks");
    int rc = make_tmap_4d(&tm_x, xg, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, xdims, xpitch, xbox, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    const void* yh = y_hi;
    const void* yl = y_lo;
    const uint64_t yT = (uint64_t)T, yB = (uint64_t)B;
    if ((int64_t)nblk * 2 * T >= (1LL << 31)) {
        set_error("lstm_tc: too many (sequence block, step) records for one launch (B=%d T=%d)", B, T);
        return B200VAD_EINVAL;
    }
    static int parts = -1;
    if (parts < 0) { const char* e = getenv("B200VAD_LSTM_PARTS"); parts = (e && atoi(e) == 2) ? 2 : LPARTS_DEFAULT; }
    const int nb_env = g_lstm_tile;
    // 16 sequences per CTA while that still fits one wave of CTAs (latency mode), else 64 (throughput mode)
    int sms = 148, dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        sms = 148;
    }
    int nb = ((B + 15) / 16) * 2 <= sms ? 16 : 64;
    if (nb_env == 16 || nb_env == 64) nb = nb_env;
    const uint32_t lpn = nb == 16 ? 16 : XBLK / parts;
    rc = make_tmap_3d(&tm_yh, yh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2 * kHidden, yT, yB, (uint64_t)2 * kHidden * 2,
                      yT * 2 * kHidden * 2, 64, 1, lpn, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_3d(&tm_yl, yl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2 * kHidden, yT, yB, (uint64_t)2 * kHidden * 2,
                      yT * 2 * kHidden * 2, 64, 1, lpn, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("B200VAD_LSTM_DEBUG"); dbg = e ? atoi(e) : 0; }
    LstmTcParams p{whh, B, T, dbg, scaled_planes ? 1.f - kPlaneScale : 1.f};
    typedef void (*KernFn)(CUtensorMap, CUtensorMap, CUtensorMap, LstmTcParams);
    static const KernFn kerns[3] = {lstm_tc_kernel<2, 64>, lstm_tc_kernel<4, 64>, lstm_tc_kernel<1, 16>};
    static const int smems[3] = {LstmGeom<64>::SMEM, LstmGeom<64>::SMEM, LstmGeom<16>::SMEM};
    const int ki = nb == 16 ? 2 : (parts == 4 ? 1 : 0);
    const int smem = smems[ki];
    dim3 grid((B + nb - 1) / nb, 2);
    prof_begin(0, st);
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(kerns[ki]), smem))) return rc;
    kerns[ki]<<<grid, LTC_THREADS, smem, st>>>(tm_x, tm_yh, tm_yl, p);
    prof_end(0, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
