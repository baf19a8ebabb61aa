// Fused bidirectional LSTM layer for sm_100a: input projection + recurrence in ONE kernel, no xg tensor in HBM.
//
// nn.LSTM semantics (PyanNet2.py:95,170; PyanNet.py:105,181): gates = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh;
// i,f,o = sigmoid, g = tanh; c_t = f*c_{t-1} + i*g; h_t = o*tanh(c_t).
//
// The two-kernel layer (gemm_xg_pair_kernel -> xg in HBM -> lstm_tc_kernel) moved 8 KB of fp32 xg per frame and layer
// (107 of the 140 GB a 4096 x 8 s step moved) and had room for only ONE fp16 plane of W_hh in tensor memory.  Here a
// CLUSTER OF 4 CTAs owns a block of sequences of one direction and splits the HIDDEN UNITS: CTA r computes the four
// gates of units [32 r, 32 r + 32) = 128 gate rows = one MMA M tile.  Per CTA, tensor memory (512 columns) holds
//   * W_hh rows of its units as TWO fp16 planes (W_hi, W' -- the scaled split below): 128 columns,
//   * W_ih rows of its units as two planes (W_hi + W_lo or W_hi + W'):                <= 256 columns (D <= 256),
//   * the gate accumulators of up to 8 parts of 16 sequences:                           128 columns,
// so all weights of the layer stay resident for all T steps and nothing but x_t (fp16 planes, 1 KB per sequence-step) is
// read from HBM and nothing but h_t (the layer output planes) is written.
// Per step and part (16 sequences) the single MMA-issuing thread of a CTA issues
//   x-MMAs:  acc  = W_ih . x_t        (A = weights in TMEM, B = the TMA-loaded x tile; issued one step AHEAD, as soon as the
//                                      pointwise warps have read the previous step's accumulator),
//   h-MMAs:  acc += W_hh . h_{t-1}    (B = the h operand tile in shared memory) once the cluster has exchanged h_{t-1},
// the pointwise warps turn the accumulator into h_t for the CTA's 32 units, write it into the CTA's own slice of the h
// operand tile and to HBM, and one thread sends that 2 KB slice to the three peers with cp.async.bulk over distributed
// shared memory (the peers' mbarriers count the bytes).  x tiles are TMA-multicast: each CTA fetches a quarter of a tile
// for the whole cluster.  8 parts per cluster are in flight, so the exchange of one part hides behind the MMAs of the others.
//
// Precision (DESIGN.md "Precision"): every product is split-precision with fp32 accumulation.  Activations travel as the
// scaled split x1 = fp16((1 - s) x), x2 = fp16(x - x1), s = 2^-6, weights as W_hi = fp16(W), W' = fp16(W_hi + W_lo / s):
// x1 . W_hi + x2 . W' = x . W to ~2^-18 with TWO fp16 MMAs.  The recurrent product uses the same split for h, so W_hh has
// the accuracy of two planes (the single-plane W_hh of lstm_tc.cu was the dominant error of the old path: 1.3e-3 on p at
// logit spread 2).  Layer 0 may take (hi, lo) planes and three MMAs (x_lo.W_hi + x_hi.W_lo + x_hi.W_hi).
//
// K ordering of the recurrent product: the MMA K index is free as long as A and B agree.  K block r (64 slots = one
// 128-byte swizzled row) is [h1 of units 32r..32r+31 | h2 of the same units], i.e. exactly what CTA r produces, so a
// CTA's contribution to a part's operand tile is ONE contiguous 2 KB region (16 sequences x 128 bytes).
//
// Algorithmic work per (sequence, frame, direction): FLOPs 2 * 512 * (D + 128); HBM bytes 2 * 2 * D8 (x planes read,
// shared by both directions through L2) + 512 (y planes written).
#include "kernels.cuh"
#include "tc05.cuh"
#include <stdlib.h>
#include <algorithm>
#include <mutex>

namespace b200vad {

using namespace tc;

constexpr int FC = 4;                       // CTAs per cluster
constexpr int FU = kHidden / FC;            // hidden units per CTA (32)
constexpr int FPN = 16;                     // sequences per part (MMA N)
constexpr int FMAXP = 8;                    // parts per work item (128 sequences per cluster)
constexpr int FWG = 4;                      // pointwise warpgroups; warpgroup w owns parts w and w + 4
constexpr int F_THREADS = (FWG * 4 + 2) * 32;   // 16 pointwise warps + TMA producer warp + MMA warp = 576
constexpr int F_BOX = FPN * 128;            // one x box / one h k-block: 16 rows x 128 bytes = 2 KB
constexpr int F_HTILE = FC * F_BOX;         // h operand tile of a part: 4 k-blocks = 8 KB
constexpr int F_EXCH = 4 * FPN * FU * 4;    // gate transposition buffer of a warpgroup: [gate][sequence][unit] fp32 = 8 KB
constexpr int F_ACC_COL = 0;                // TMEM columns [0, 128): accumulators, part p at 16 p
constexpr int F_WHH_COL = FMAXP * FPN;      // [128, 256): W_hh, k-step j at 128 + 8 j
constexpr int F_WIH_COL = F_WHH_COL + 128;  // [256, 256 + Dp): W_ih plane a then plane b
constexpr int F_MAX_STAGES = 16;
constexpr int F_SMEM_FIXED = FMAXP * F_HTILE + FWG * F_EXCH;   // 96 KB
constexpr int F_SMEM_MAX = 232448;          // 227 KB per CTA

struct FusedParams {
    const __half* wih_hi;    // [2 * 512][ldw]  (gate-scaled, api.cu packing)
    const __half* wih_lo;
    const __half* whh_hi;    // [2][512][128]
    const __half* whh_lo;
    const float* bias;       // [2 * 512] b_ih + b_hh, gate-scaled
    __half* y_a;             // layer output planes [B][T][256]
    __half* y_b;
    int B, T;
    int nk;                  // k-steps of 16 of the input projection (D rounded up to 16) / 16
    int kblocks;             // 64-wide k blocks per x plane (TMA boxes per plane)
    int ldw;                 // row pitch of wih planes (elements)
    int terms;               // 2: planes are the scaled (x1, x2) split, plane b of W_ih is W';  3: (hi, lo) planes, plane b is W_lo
    int y_scaled;            // 1: y planes use the scaled split (feeds a 2-MMA product), 0: plain hi / lo (feeds the head)
    int items_per_dir;       // work items per direction; parts are spread evenly over them
    int stages;              // x ring depth
    int lag;                 // x-MMAs of a part are issued `lag` part slots after its h-MMAs
    int flags;               // debug (B200VAD_FUSED_DEBUG): 1 = skip the h exchange (timing probe, wrong results)
};

#define FUSED_WAIT(bar, parity, tag) mbar_wait_tag(bar, parity, tag)
__device__ __forceinline__ void mbar_wait_tag(uint32_t bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {   // ~2 s
            printf("b200vad lstm_fused: wait timed out (block %d thread %d tag %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, tag, bar,
                   parity);
            __trap();
        }
    }
}

__global__ void __cluster_dims__(FC, 1, 1) __launch_bounds__(F_THREADS, 1)
lstm_fused_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, FusedParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* const smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t h_off = 0;                                  // [part 8][k-block 4][16 rows][128 B]
    const uint32_t e_off = FMAXP * F_HTILE;                    // [warpgroup 4][gate 4][sequence 16][unit 32] fp32
    const uint32_t x_off = e_off + FWG * F_EXCH;               // [stage][plane 2][k-block][16 rows][128 B]
    const uint32_t stage_bytes = 2u * p.kblocks * F_BOX;
    const uint32_t bar_base = smem_base + x_off + p.stages * stage_bytes;
    auto bar_x_full = [&](int s) { return bar_base + 8 * s; };
    auto bar_x_empty = [&](int s) { return bar_base + 128 + 8 * s; };
    auto bar_acc_ready = [&](int q) { return bar_base + 256 + 8 * q; };
    auto bar_acc_free = [&](int q) { return bar_base + 320 + 8 * q; };
    auto bar_h_ready = [&](int q) { return bar_base + 384 + 8 * q; };
    auto bar_h_free = [&](int q) { return bar_base + 448 + 8 * q; };
    const uint32_t tmem_slot = bar_base + 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / FC, num_clusters = gridDim.x / FC;
    const int T = p.T;
    const int P = (p.B + FPN - 1) / FPN;                      // parts per direction
    const int ipd = p.items_per_dir;
    const int num_items = 2 * ipd;
    const int base_parts = P / ipd, rem_parts = P % ipd;

    if (threadIdx.x == 0) {
        for (int s = 0; s < F_MAX_STAGES; ++s) { mbar_init(bar_x_full(s), 1); mbar_init(bar_x_empty(s), FC); }
        for (int q = 0; q < FMAXP; ++q) {
            mbar_init(bar_acc_ready(q), 1);
            mbar_init(bar_acc_free(q), 4);
            mbar_init(bar_h_ready(q), 1);
            mbar_init(bar_h_free(q), FC);
        }
        mbar_fence_init();
    }
    if (warp == 17) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                                       // every CTA's barriers exist before any remote arrive / copy
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // persistent role state (phases carry over from item to item)
    uint32_t xcount = 0;                                      // producer / MMA thread: x ring uses so far
    uint32_t ph_a = 0, ph_b = 0;                              // per-part phase bits (meaning depends on the role)
    int loaded_dir = -1;

    for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int dir = item / ipd, ii = item - dir * ipd;
        const int nparts = base_parts + (ii < rem_parts ? 1 : 0);
        const int part0 = ii * base_parts + min(ii, rem_parts);
        const int seq0 = part0 * FPN;
        if (nparts == 0) continue;                            // (uniform over the cluster)
        const int total = T * nparts;

        // ---------------- weights of this direction -> tensor memory (pointwise warps; only when the direction changes)
        if (warp < 16 && loaded_dir != dir) {
            const int wg = warp >> 2, g = warp & 3;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(g * 32) << 16);
            const int grow = dir * kGates + g * kHidden + (int)rank * FU + lane;          // packed weight row of this TMEM lane
            // W_hh: k-step j holds plane (j >> 1) & 1 of units 32 (j >> 2) + 16 (j & 1) .. + 15
            for (int j = wg; j < 16; j += FWG) {
                const int u0 = 32 * (j >> 2) + 16 * (j & 1), plane = (j >> 1) & 1;
                const uint4* hp = reinterpret_cast<const uint4*>(p.whh_hi + (size_t)grow * kHidden + u0);
                const uint4* lp = reinterpret_cast<const uint4*>(p.whh_lo + (size_t)grow * kHidden + u0);
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint4 vh = __ldg(hp + i), vl = __ldg(lp + i);
                    const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (plane == 0) {
                            r[4 * i + e] = hw[e];
                        } else {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                            const __half2 wp = __floats2half2_rn(fmaf(fl.x, 1.f / kPlaneScale, fh.x), fmaf(fl.y, 1.f / kPlaneScale, fh.y));
                            r[4 * i + e] = *reinterpret_cast<const uint32_t*>(&wp);
                        }
                    }
                }
                tmem_st8(lane_addr + F_WHH_COL + 8 * j, r);
            }
            // W_ih: plane a = W_hi, plane b = W_lo (terms 3) or W' (terms 2); k-step j of plane pl at F_WIH_COL + pl * 8 nk + 8 j
            for (int j = wg; j < 2 * p.nk; j += FWG) {
                const int pl = j >= p.nk, kj = pl ? j - p.nk : j;
                const uint4* hp = reinterpret_cast<const uint4*>(p.wih_hi + (size_t)grow * p.ldw + 16 * kj);
                const uint4* lp = reinterpret_cast<const uint4*>(p.wih_lo + (size_t)grow * p.ldw + 16 * kj);
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint4 vh = __ldg(hp + i), vl = __ldg(lp + i);
                    const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (!pl) {
                            r[4 * i + e] = hw[e];
                        } else if (p.terms == 3) {
                            r[4 * i + e] = lw[e];
                        } else {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                            const __half2 wp = __floats2half2_rn(fmaf(fl.x, 1.f / kPlaneScale, fh.x), fmaf(fl.y, 1.f / kPlaneScale, fh.y));
                            r[4 * i + e] = *reinterpret_cast<const uint32_t*>(&wp);
                        }
                    }
                }
                tmem_st8(lane_addr + F_WIH_COL + 8 * j, r);
            }
            tmem_st_wait();
        }
        loaded_dir = dir;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (warp == 16) {
            // ===================== TMA producer: x tiles of (step, part) in issue order, multicast to the cluster =====================
            if (elect_one()) {
                const int nboxes = 2 * p.kblocks;
                for (int k = 0; k < total; ++k) {
                    const int s = k / nparts, q = k - s * nparts;
                    const int t = dir == 0 ? s : T - 1 - s;
                    const int st = (int)(xcount % (uint32_t)p.stages);
                    const uint32_t ph = (xcount / (uint32_t)p.stages) & 1u;
                    FUSED_WAIT(bar_x_empty(st), ph ^ 1u, 1);                  // all 4 CTAs' MMAs are done with this stage
                    mbar_expect_tx(bar_x_full(st), stage_bytes);
                    const uint32_t dst = smem_base + x_off + st * stage_bytes;
                    for (int bi = (int)rank; bi < nboxes; bi += FC) {
                        const int pl = bi >= p.kblocks, kb = pl ? bi - p.kblocks : bi;
                        tma_load_3d_mc(dst + bi * F_BOX, pl ? &tm_b : &tm_a, kb * 64, t, seq0 + q * FPN, bar_x_full(st), (uint16_t)0xF);
                    }
                    ++xcount;
                }
            }
        } else if (warp == 17) {
            // ===================== MMA issuer =====================
            if (elect_one()) {
                constexpr uint32_t idesc = idesc_f16(128, FPN);
                const int lag = min(p.lag, nparts - 1);
                // ph_a: h_ready phase bits, ph_b: acc_free phase bits
                auto issue_x = [&](int k) {
                    const int s = k / nparts, q = k - s * nparts;
                    if (s > 0) {                                                   // the pointwise warps have read acc(s - 1, q)
                        FUSED_WAIT(bar_acc_free(q), (ph_b >> q) & 1u, 2);
                        ph_b ^= 1u << q;
                    }
                    const int st = (int)(xcount % (uint32_t)p.stages);
                    const uint32_t ph = (xcount / (uint32_t)p.stages) & 1u;
                    FUSED_WAIT(bar_x_full(st), ph, 3);
                    tc_fence_after();
                    const uint32_t xa = smem_base + x_off + st * stage_bytes, xb = xa + p.kblocks * F_BOX;
                    const uint32_t d = tmem_base + F_ACC_COL + q * FPN;
                    const uint32_t wa0 = tmem_base + F_WIH_COL, wb0 = wa0 + 8 * p.nk;
                    for (int j = 0; j < p.nk; ++j) {
                        const uint32_t off = (j >> 2) * F_BOX + (j & 3) * 32;
                        const uint64_t da = smem_desc_sw128(xa + off), db = smem_desc_sw128(xb + off);
                        if (p.terms == 3) {
                            mma_f16_ts(d, wa0 + 8 * j, db, idesc, j != 0);         // x_lo . W_hi (small terms first)
                            mma_f16_ts(d, wb0 + 8 * j, da, idesc, 1);              // x_hi . W_lo
                        } else {
                            mma_f16_ts(d, wb0 + 8 * j, db, idesc, j != 0);         // x2 . W'
                        }
                        mma_f16_ts(d, wa0 + 8 * j, da, idesc, 1);                  // x1 . W_hi
                    }
                    mma_commit_mc(bar_x_empty(st), (uint16_t)0xF);
                    ++xcount;
                    if (s == 0) {                                                  // h_{-1} = 0: no recurrent product at the first step
                        mma_commit(bar_acc_ready(q));
                        if (T > 1) mma_commit_mc(bar_h_free(q), (uint16_t)0xF);
                    }
                };
                for (int k = 0; k < nparts; ++k) issue_x(k);
                for (int j = 0; j < total; ++j) {
                    const int s = j / nparts, q = j - s * nparts;
                    if (s > 0) {
                        FUSED_WAIT(bar_h_ready(q), (ph_a >> q) & 1u, 4);           // h_{s-1} of this part: all four slices landed
                        ph_a ^= 1u << q;
                        tc_fence_after();
                        const uint32_t hb = smem_base + h_off + q * F_HTILE;
                        const uint32_t d = tmem_base + F_ACC_COL + q * FPN;
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            const uint64_t dh = smem_desc_sw128(hb + (jj >> 2) * F_BOX + (jj & 3) * 32);
                            mma_f16_ts(d, tmem_base + F_WHH_COL + 8 * jj, dh, idesc, 1);
                        }
                        mma_commit(bar_acc_ready(q));
                        if (s < T - 1) mma_commit_mc(bar_h_free(q), (uint16_t)0xF);   // every CTA may now overwrite this part's h tile
                    }
                    const int k = j - lag + nparts;
                    if (j >= lag && k < total) issue_x(k);
                }
                // the accumulators of the last step must be read before the next item overwrites them
                for (int q = 0; q < nparts; ++q) {
                    FUSED_WAIT(bar_acc_free(q), (ph_b >> q) & 1u, 5);
                    ph_b ^= 1u << q;
                }
            }
        } else {
            // ===================== pointwise warpgroups =====================
            const int wg = warp >> 2, g = warp & 3;                               // g: gate read from TMEM == warp index in the warpgroup
            const uint32_t lane_addr = tmem_base + ((uint32_t)(g * 32) << 16);
            const float bias = __ldg(p.bias + dir * kGates + g * kHidden + (int)rank * FU + lane);
            // after the transposition: 4 units (quad q4) of one sequence n
            const int n = g + 4 * (lane >> 3), q4 = lane & 7;
            float* const exw = reinterpret_cast<float*>(smem_gen + e_off + wg * F_EXCH) + g * (FPN * FU) + lane;          // + 32 * sequence
            const float4* const exr = reinterpret_cast<const float4*>(smem_gen + e_off + wg * F_EXCH) + n * (FU / 4) + q4;   // + gate * 128
            const bool tid0 = (threadIdx.x & 127) == 0;
            const float s1 = 1.f - kPlaneScale;
            const float L2E2 = 2.f * kLog2e;
            float cst[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) cst[a][b] = 0.f;
            // own slice of the h operand tile: row n, chunk (q4 >> 1) for h1 and 4 + (q4 >> 1) for h2, 8-byte half (q4 & 1)
            const uint32_t hrow = (uint32_t)rank * F_BOX + n * 128 + ((q4 & 1) << 3);
            const uint32_t hc1 = (uint32_t)(((q4 >> 1) ^ (n & 7)) << 4), hc2 = (uint32_t)(((4 + (q4 >> 1)) ^ (n & 7)) << 4);
            // ph_a: acc_ready phase bits, ph_b: h_free phase bits
            for (int s = 0; s < T; ++s) {
                const int t = dir == 0 ? s : T - 1 - s;
#pragma unroll
                for (int pi = 0; pi < 2; ++pi) {
                    const int q = wg + FWG * pi;
                    if (q >= nparts) break;
                    FUSED_WAIT(bar_acc_ready(q), (ph_a >> q) & 1u, 6);
                    ph_a ^= 1u << q;
                    tc_fence_after();
                    float z[16];
                    tmem_ld16(lane_addr + F_ACC_COL + q * FPN, z);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_free(q));
                    // 2^(pre-activation) of this thread's gate row for the 16 sequences -> [gate][sequence][unit]
#pragma unroll
                    for (int j = 0; j < 16; ++j) exw[j * FU] = fast_ex2(fminf(z[j] + bias, 29.f));
                    named_bar_sync(1 + wg, 128);
                    const float4 vi = exr[0], vf = exr[FPN * FU / 4], vg = exr[2 * FPN * FU / 4], vo = exr[3 * FPN * FU / 4];
                    const float ei[4] = {vi.x, vi.y, vi.z, vi.w}, ef[4] = {vf.x, vf.y, vf.z, vf.w};
                    const float eg[4] = {vg.x, vg.y, vg.z, vg.w}, eo[4] = {vo.x, vo.y, vo.z, vo.w};
                    float hv[4];
                    const f32x2 one = pack2(1.f, 1.f), mone = pack2(-1.f, -1.f), k2 = pack2(L2E2, L2E2);
#pragma unroll
                    for (int jp = 0; jp < 2; ++jp) {
                        const f32x2 pei = pack2(ei[2 * jp], ei[2 * jp + 1]), pef = pack2(ef[2 * jp], ef[2 * jp + 1]);
                        const f32x2 peg = pack2(eg[2 * jp], eg[2 * jp + 1]), peo = pack2(eo[2 * jp], eo[2 * jp + 1]);
                        // c' = c / (1 + ef) + (eg - 1) / ((1 + ei)(eg + 1)) with one reciprocal (lstm_tc.cu)
                        const f32x2 di = add2(pei, one), df = add2(pef, one), dg = add2(peg, one);
                        const f32x2 dig = mul2(di, dg);
                        const f32x2 den = mul2(df, dig);
                        const f32x2 cn = mul2(fma2(pack2(cst[pi][2 * jp], cst[pi][2 * jp + 1]), dig, mul2(add2(peg, mone), df)), rcp2(den));
                        unpack2(cn, cst[pi][2 * jp], cst[pi][2 * jp + 1]);
                        const f32x2 ec = ex2_clamped2(mul2(cn, k2));
                        const f32x2 h2v = mul2(add2(ec, mone), rcp2(mul2(add2(peo, one), add2(ec, one))));
                        unpack2(h2v, hv[2 * jp], hv[2 * jp + 1]);
                    }
                    // recurrent operand planes: always the scaled split (h1 . W_hi + h2 . W')
                    __half a1[4], a2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_scaled_f16(hv[e], s1, a1[e], a2[e]);
                    uint2 w1, w2;
                    w1.x = (uint32_t)__half_as_ushort(a1[0]) | ((uint32_t)__half_as_ushort(a1[1]) << 16);
                    w1.y = (uint32_t)__half_as_ushort(a1[2]) | ((uint32_t)__half_as_ushort(a1[3]) << 16);
                    w2.x = (uint32_t)__half_as_ushort(a2[0]) | ((uint32_t)__half_as_ushort(a2[1]) << 16);
                    w2.y = (uint32_t)__half_as_ushort(a2[2]) | ((uint32_t)__half_as_ushort(a2[3]) << 16);
                    const bool exchange = s + 1 < T;
                    if (exchange) {
                        // the peers' MMAs of this step have finished reading their copy of this part's tile, and therefore
                        // (their h_ready of the previous step completed) last step's copies out of our slice have landed
                        FUSED_WAIT(bar_h_free(q), (ph_b >> q) & 1u, 7);
                        ph_b ^= 1u << q;
                        unsigned char* const hp = smem_gen + h_off + q * F_HTILE + hrow;
                        *reinterpret_cast<uint2*>(hp + hc1) = w1;
                        *reinterpret_cast<uint2*>(hp + hc2) = w2;
                        fence_proxy_async();
                    }
                    // layer output planes (B, T, 256): scaled split (next layer's 2-MMA projection) or hi / lo (head)
                    const int b = seq0 + q * FPN + n;
                    if (b < p.B) {
                        if (!p.y_scaled) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) split_f16(hv[e], a1[e], a2[e]);
                            w1.x = (uint32_t)__half_as_ushort(a1[0]) | ((uint32_t)__half_as_ushort(a1[1]) << 16);
                            w1.y = (uint32_t)__half_as_ushort(a1[2]) | ((uint32_t)__half_as_ushort(a1[3]) << 16);
                            w2.x = (uint32_t)__half_as_ushort(a2[0]) | ((uint32_t)__half_as_ushort(a2[1]) << 16);
                            w2.y = (uint32_t)__half_as_ushort(a2[2]) | ((uint32_t)__half_as_ushort(a2[3]) << 16);
                        }
                        const size_t yo = ((size_t)b * T + t) * (2 * kHidden) + dir * kHidden + (int)rank * FU + 4 * q4;
                        *reinterpret_cast<uint2*>(p.y_a + yo) = w1;
                        *reinterpret_cast<uint2*>(p.y_b + yo) = w2;
                    }
                    named_bar_sync(1 + wg, 128);                                   // slice complete (and the exchange buffer is free again)
                    if (exchange && tid0) {
                        const uint32_t src = smem_base + h_off + q * F_HTILE + rank * F_BOX;
                        if (p.flags & 1) {
                            mbar_arrive(bar_h_ready(q));
                        } else {
                            mbar_expect_tx(bar_h_ready(q), (FC - 1) * F_BOX);      // own arrival + the three peers' slices
#pragma unroll
                            for (uint32_t d = 1; d < FC; ++d) {
                                const uint32_t peer = (rank + d) & (FC - 1);
                                dsmem_bulk_copy(mapa_shared(src, peer), src, F_BOX, mapa_shared(bar_h_ready(q), peer));
                            }
                        }
                    }
                }
            }
        }
        // item boundary: every role of every CTA is done with this item's tiles and barriers
        tc_fence_before();
        cluster_sync_all();
        tc_fence_after();
    }
    tc_fence_before();
    cluster_sync_all();                                       // no CTA exits while a peer may still signal or copy into it
    if (warp == 17) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------- launcher
static int g_fused_clusters[64];
static std::once_flag g_fused_once[64];

static int fused_smem_bytes(int kblocks, int* stages_out) {
    const int stage = 2 * kblocks * F_BOX;
    int stages = (F_SMEM_MAX - 1024 - F_SMEM_FIXED - 1024) / stage;
    if (stages > F_MAX_STAGES) stages = F_MAX_STAGES;
    *stages_out = stages;
    return 1024 + F_SMEM_FIXED + stages * stage + 1024;
}

// resident clusters of the kernel on this device (cudaOccupancyMaxActiveClusters); cached per device
static int fused_max_clusters(int smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::call_once(g_fused_once[dev], [&] {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(FC * 64);
        cfg.blockDim = dim3(F_THREADS);
        cfg.dynamicSmemBytes = F_SMEM_MAX - 1024;            // worst case
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = FC; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = 0;
        cudaFuncSetAttribute(reinterpret_cast<const void*>(lstm_fused_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_MAX - 1024);
        if (cudaOccupancyMaxActiveClusters(&n, reinterpret_cast<const void*>(lstm_fused_kernel), &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            n = std::max(1, sms / FC - 4);
        }
        const char* e = getenv("B200VAD_FUSED_CLUSTERS");
        if (e && atoi(e) > 0) n = atoi(e);
        g_fused_clusters[dev] = n;
    });
    (void)smem;
    return g_fused_clusters[dev];
}

int lstm_fused_supported(int D) { return D >= 1 && D <= 256; }
int lstm_fused_clusters() { int st; return fused_max_clusters(fused_smem_bytes(4, &st)); }

// One bidirectional LSTM layer.  x_a / x_b: input planes (B, T, lda) fp16 (terms 2: scaled split; terms 3: hi / lo);
// weights as packed by api.cu (gate-scaled): wih planes [1024][ldw], whh planes [2][512][128], bias [1024];
// y_a / y_b: output planes (B, T, 256), scaled split if y_scaled else hi / lo.
int lstm_fused_launch(const __half* x_a, const __half* x_b, int64_t lda, int B, int T, int D, const __half* wih_hi,
                      const __half* wih_lo, int ldw, const __half* whh_hi, const __half* whh_lo, const float* bias, int terms,
                      __half* y_a, __half* y_b, int y_scaled, cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    if (!lstm_fused_supported(D) || (terms != 2 && terms != 3) || lda % 8 != 0 || ldw % 8 != 0 || ldw < (D + 15) / 16 * 16) {
        set_error("lstm_fused: unsupported shape (D=%d lda=%lld ldw=%d terms=%d)", D, (long long)lda, ldw, terms);
        return B200VAD_EINVAL;
    }
    FusedParams p;
    p.wih_hi = wih_hi; p.wih_lo = wih_lo; p.whh_hi = whh_hi; p.whh_lo = whh_lo; p.bias = bias; p.y_a = y_a; p.y_b = y_b;
    p.B = B; p.T = T; p.nk = (D + 15) / 16; p.kblocks = (D + 63) / 64; p.ldw = ldw; p.terms = terms; p.y_scaled = y_scaled;
    int stages = 0;
    const int smem = fused_smem_bytes(p.kblocks, &stages);
    p.stages = stages;
    static int dbg = -1, lag_env = -1;
    if (dbg < 0) { const char* e = getenv("B200VAD_FUSED_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (lag_env < 0) { const char* e = getenv("B200VAD_FUSED_LAG"); lag_env = e ? atoi(e) : 3; }
    p.flags = dbg; p.lag = lag_env;
    const int nc = fused_max_clusters(smem);
    // work items: parts of 16 sequences, spread evenly over items_per_dir items per direction (<= 8 parts each); choose the
    // count that minimises waves x step cost (a step costs ~ max(parts, 3) part slots: below ~3 parts the per-part
    // MMA -> pointwise -> exchange chain is the critical path)
    const int P = (B + FPN - 1) / FPN;
    int best_ipd = (P + FMAXP - 1) / FMAXP;
    double best_cost = 1e30;
    for (int ipd = (P + FMAXP - 1) / FMAXP; ipd <= P; ++ipd) {
        const int maxp = (P + ipd - 1) / ipd;
        const int waves = (2 * ipd + nc - 1) / nc;
        const double cost = waves * std::max<double>(maxp, 3.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_ipd = ipd; }
        if (maxp == 1) break;
    }
    p.items_per_dir = best_ipd;
    const int grid = FC * std::min(nc, 2 * best_ipd);
    CUtensorMap tm_a, tm_b;
    int rc;
    if ((rc = make_tmap_3d(&tm_a, x_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, D, T, B, lda * 2, (uint64_t)T * lda * 2, 64, 1, FPN,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_b, x_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, D, T, B, lda * 2, (uint64_t)T * lda * 2, 64, 1, FPN,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(lstm_fused_kernel), smem))) return rc;
    prof_begin(0, st);
    lstm_fused_kernel<<<grid, F_THREADS, smem, st>>>(tm_a, tm_b, p);
    prof_end(0, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
