// Fused bidirectional LSTM layer for sm_100a, CTA-PAIR form: the same layer as lstm_fused.cu (input projection + recurrence in
// one launch, all weights resident in tensor memory, no xg tensor in HBM; nn.LSTM semantics of PyanNet2.py:95,170 /
// PyanNet.py:105,181), with every product issued as tcgen05.mma.cta_group::2 by the even CTA of a pair.
//
// Why: in lstm_fused.cu every CTA of the 4-CTA cluster needs the WHOLE operand of every product in its own shared memory --
// all 128 sequences of x_t (128 KB per step, TMA multicast to all four CTAs) and all 128 sequences of h_{t-1} (each CTA
// stores its 16 KB slice into three peers: 48 KB out, 48 KB in per step over distributed shared memory).  The ablations
// (profiles/r02_lstm_fused.md) put ~25 % of the layer time on those two inbound streams.  A cta_group::2 MMA splits the B
// operand (the N = sequence dimension) over the two CTAs of a pair: each CTA keeps only ITS HALF of the sequences of every
// operand tile and the tensor cores fetch the other half from the peer themselves.  So per CTA and step
//   * x_t:      64 KB instead of 128 (the tile of a CTA is multicast to the one CTA of the other pair with the same half),
//   * h_{t-1}:  24 KB in / 24 KB out instead of 48 / 48 (a CTA's slice of a part goes to TWO CTAs per half, one of which
//               is itself for its own half).
//
// Cluster of 4 CTAs = 2 pairs; CTA r owns hidden units [32 r, 32 r + 32) exactly as in lstm_fused.cu (TMEM lane 32 g + l =
// gate l & 3 of unit 8 g + (l >> 2); W_hh two planes in columns [128, 256), W_ih two planes in [256, 256 + 16 nk),
// accumulators of 8 parts x 16 sequences in [0, 128)).  Pair p = CTAs (2 p, 2 p + 1): its MMAs have M = 256 = the gate rows of
// units [64 p, 64 p + 64).  hf = r & 1 is the CTA's half of the N dimension:
//   * recurrent product of part q (N = 16): CTA hf holds sequences [8 hf, 8 hf + 8) of the part: [k-block 4][8 rows][128 B];
//   * input product of a pair of parts (N = 32): CTA hf holds the 16 sequences of part 2 pp + hf.
// Only the even CTA (the leader) of a pair issues MMAs.  Its barriers collect both CTAs' readiness: operand tiles that land in
// the odd CTA complete the odd CTA's own mbarrier (TMA / st.async credit the destination CTA's barrier), and a forwarder
// thread there (the warps that issue MMAs in the leader) passes the phase on with one remote release-arrive.  Completion
// travels back by multicast tcgen05.commit.
#include "kernels.cuh"
#include "tc05.cuh"
#include <stdlib.h>
#include <algorithm>
#include <mutex>

namespace b200vad {

using namespace tc;

constexpr int PC = 4;                       // CTAs per cluster
constexpr int PU = kHidden / PC;            // hidden units per CTA (32)
constexpr int PPN = 16;                     // sequences per part
constexpr int PMAXP = 8;                    // parts per work item (128 sequences per cluster)
constexpr int P_PW_WARPS = 16;              // pointwise warps: warp w owns TMEM lane quarter w & 3 of parts (w >> 2) and (w >> 2) + 4
constexpr int P_W_PROD = P_PW_WARPS;        // TMA producer of the x tiles
constexpr int P_W_MMA = P_W_PROD + 1;       // leader: 2 recurrent-product issuers; odd CTA: 2 h_ready forwarders (also owns the TMEM allocation)
constexpr int P_W_MMAX = P_W_PROD + 3;      // leader: 2 input-product issuers; odd CTA: 2 x_full forwarders
constexpr int P_W_SEND = P_W_PROD + 5;      // 4 exchange senders
constexpr int P_THREADS = (P_W_PROD + 9) * 32;   // 25 warps
constexpr int P_HBOX = 8 * 128;             // one h k-block of a part in one CTA: 8 sequences x 128 bytes = one swizzle atom
constexpr int P_HTILE = PC * P_HBOX;        // h operand tile of a part in one CTA: 4 k-blocks (one per source CTA) = 4 KB
constexpr int P_SLICE = PPN * 128;          // own h slice of a part: 16 sequences x (32 units x 2 planes) = 2 KB
constexpr int P_STAGING = 2 * PMAXP * P_SLICE;   // [step parity][part]: 32 KB
constexpr int P_XBOX = PPN * 128;           // one x TMA box: this CTA's 16 sequences of a pair of parts, 64 k-values of one plane
constexpr int P_SCRATCH = 32 * 16 * 4;      // per pointwise warp: gate transposition scratch (2 KB)
constexpr int P_ACC_COL = 0;
constexpr int P_WHH_COL = PMAXP * PPN;      // 128
constexpr int P_WIH_COL = P_WHH_COL + 128;  // 256
constexpr int P_MAX_STAGES = 16;
constexpr int P_SMEM_FIXED = 2 * PMAXP * P_HTILE + P_STAGING + P_PW_WARPS * P_SCRATCH;   // 128 KB (two h tile sets)
constexpr int P_SMEM_MAX = 232448;

struct PairParams {
    const __half* wih_hi;
    const __half* wih_lo;
    const __half* whh_hi;
    const __half* whh_lo;
    const float* bias;
    __half* y_a;
    __half* y_b;
    int B, T;
    int nk, kblocks, ldw, terms;
    int items_per_dir;
    int stages;
    int opt;                 // 1: the input-product issuers yield to a pending recurrent batch; 2: double-buffered h tiles (no h_free hand-shake)
};

__device__ volatile int* g_pair_err_host = nullptr;

// bounded wait (a protocol bug traps instead of hanging the GPU box); CLUSTER: acquire at cluster scope (the phase was
// completed by another CTA's release-arrive)
// wait-time probe (opt & 64, tools/pair_waits.py): per (CTA, warp, wait tag) cycles spent in waits that were not already satisfied, and their count
constexpr int P_DBG_TAGS = 8, P_DBG_WARPS = 25, P_DBG_MAX_CTAS = 8;
__device__ long long g_pair_dbg[P_DBG_MAX_CTAS * P_DBG_WARPS * P_DBG_TAGS * 2];

template <bool CLUSTER>
__device__ __forceinline__ void pair_wait(uint32_t bar, uint32_t parity, int tag, long long* wacc = nullptr) {
    if (wacc ? mbar_test_wait(bar, parity) : (CLUSTER ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
        if (CLUSTER && wacc) mbar_try_wait_cluster(bar, parity);   // (the acquire)
        return;                                                    // probe mode: test_wait never suspends, so every wait that is not already satisfied is counted
    }
    const long long t0 = wacc ? clock64() : 0;
    struct Acc { long long* w; long long t0; int tag; __device__ ~Acc() { if (w) { w[2 * tag] += clock64() - t0; w[2 * tag + 1] += 1; } } } acc{wacc, t0, tag};
    unsigned tries = 0;
    long long tb = 0;
    while (!(CLUSTER ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait_hint(bar, parity, 20000u))) {
        if ((++tries & 63u) != 0) continue;
        const long long now = clock64();
        if (tb == 0) tb = now;
        if (now - tb > 4000000000LL) {   // ~2 s
            volatile int* e = g_pair_err_host;
            if (e && e[0] == 0) {
                e[0] = 2; e[1] = (int)blockIdx.x; e[2] = (int)threadIdx.x; e[3] = tag; e[4] = (int)bar; e[5] = (int)parity; e[6] = (int)gridDim.x;
                __threadfence_system();
            }
            printf("b200vad lstm_pair: wait timed out (block %d thread %d tag %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, tag, bar, parity);
            __trap();
        }
    }
}
#define PWAIT(bar, parity, tag) pair_wait<false>(bar, parity, tag, wacc)
#define PWAIT_CL(bar, parity, tag) pair_wait<true>(bar, parity, tag, wacc)

constexpr uint32_t P_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, descriptor version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t pdesc(uint32_t lo) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(P_DESC_HI));
    return d;
}
__device__ __forceinline__ uint32_t pdesc_lo(uint32_t smem_addr) { return (smem_addr & 0x3FFFFu) >> 4; }

template <int NK, int TERMS, bool PROBE>
__global__ void __cluster_dims__(PC, 1, 1) __launch_bounds__(P_THREADS, 1)
lstm_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, PairParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* const smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t h_off = 0;                                  // [step parity 2][part 8][k-block 4][8 rows][128 B]
    const uint32_t g_off = 2 * PMAXP * P_HTILE;                  // staging [step parity 2][part 8][16 rows][128 B]
    const uint32_t c_off = g_off + P_STAGING;                  // scratch [pointwise warp 16][2 KB]
    const uint32_t x_off = c_off + P_PW_WARPS * P_SCRATCH;     // [stage][plane 2][k-block][16 rows][128 B]
    const int nk = NK > 0 ? NK : p.nk;
    const int terms = NK > 0 ? TERMS : p.terms;
    const uint32_t stage_bytes = 2u * p.kblocks * P_XBOX;
    const uint32_t bar_base = smem_base + x_off + p.stages * stage_bytes;
    // leader: own tile (tx) + the odd CTA's forward; odd: own tile.  One set per input issuer / forwarder (slots of pair pp belong to
    // issuer pp & 1), so that every barrier is waited on by ONE thread, phase after phase (lstm_fused.cu has the reason)
    auto bar_x_full = [&](int me, int s) { return bar_base + (me ? 896 : 0) + 8 * s; };
    auto bar_x_empty = [&](int s) { return bar_base + 128 + 8 * s; };   // both pairs' MMAs are done with the stage (multicast commits)
    auto bar_acc_ready = [&](int q) { return bar_base + 256 + 8 * q; }; // the pair's commit
    auto bar_acc_free = [&](int q) { return bar_base + 320 + 8 * q; };  // leader's: 4 + 4 pointwise warps of the pair
    // leader: own tile (tx) + the odd CTA's forward; odd: own tile.  One set per h tile set: with two tile sets the bytes of
    // h_{s+1} may land while a slower CTA's phase of h_s is still open, so consecutive steps must not share a barrier
    auto bar_h_ready = [&](int set, int q) { return bar_base + (set ? 768 : 384) + 8 * q; };
    auto bar_h_free = [&](int q) { return bar_base + 448 + 8 * q; };    // both pairs' recurrent MMAs of the step are done
    const uint32_t tmem_slot = bar_base + 512;
    auto bar_slice = [&](int q) { return bar_base + 528 + 8 * q; };
    auto bar_x_done = [&](int pp) { return bar_base + 592 + 8 * pp; };  // leader's, per pair of parts
    const uint32_t busy_flags = bar_base + 704;                         // leader's: [2] a recurrent issuer is inside its (critical-path) MMA batch
    const bool opt_yield = (p.opt & 1) != 0, opt_hdouble = (p.opt & 2) != 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t hf = rank & 1u, leader = rank & ~1u;
    const bool is_leader = hf == 0;
    const uint16_t pair_mask = (uint16_t)(3u << leader);
    const int cluster_id = blockIdx.x / PC, num_clusters = gridDim.x / PC;
    const int T = p.T;
    const int P = (p.B + PPN - 1) / PPN;
    const int ipd = p.items_per_dir;
    const int num_items = 2 * ipd;
    const int base_parts = P / ipd, rem_parts = P % ipd;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P_MAX_STAGES; ++s) { mbar_init(bar_x_full(0, s), is_leader ? 2 : 1); mbar_init(bar_x_full(1, s), is_leader ? 2 : 1); mbar_init(bar_x_empty(s), 2); }
        for (int q = 0; q < PMAXP; ++q) {
            mbar_init(bar_acc_ready(q), 1);
            mbar_init(bar_acc_free(q), 8);
            mbar_init(bar_h_ready(0, q), is_leader ? 2 : 1);
            mbar_init(bar_h_ready(1, q), is_leader ? 2 : 1);
            mbar_init(bar_h_free(q), 2);
            mbar_init(bar_slice(q), 4);
            mbar_init(bar_x_done(q), 1);
        }
        asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %1};" ::"r"(busy_flags), "r"(0u) : "memory");
        mbar_fence_init();
    }
    if (warp == P_W_MMA) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    long long* const wacc = (PROBE && (p.opt & 64) && lane == 0 && blockIdx.x < P_DBG_MAX_CTAS)
                                ? &g_pair_dbg[((size_t)blockIdx.x * P_DBG_WARPS + warp) * (2 * P_DBG_TAGS)] : nullptr;
    if (wacc) {
        for (int i = 0; i < 2 * P_DBG_TAGS; ++i) wacc[i] = 0;
        wacc[0] = -clock64();
    }
    int xst = 0;
    uint32_t xph = 0;
    uint32_t ph_a = 0, ph_b = 0;
    uint32_t xfph = 0;                                        // input issuer / x forwarder: phase bits of its own x_full set, per stage
    int loaded_dir = -1;

    for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int dir = item / ipd, ii = item - dir * ipd;
        const int nparts = base_parts + (ii < rem_parts ? 1 : 0);
        const int part0 = ii * base_parts + min(ii, rem_parts);
        const int seq0 = part0 * PPN;
        if (nparts == 0) continue;

        // ---------------- weights of this direction -> this CTA's tensor memory (pointwise warps)
        if (warp < P_PW_WARPS && loaded_dir != dir) {
            const int wg = warp >> 2, g = warp & 3;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(g * 32) << 16);
            const int grow = dir * kGates + (lane & 3) * kHidden + (int)rank * PU + 8 * g + (lane >> 2);
            // W_hh: k-step j holds plane (j >> 1) & 1 of units 32 (j >> 2) + 16 (j & 1) .. + 15  (K block j >> 2 = source CTA)
            for (int j = wg; j < 16; j += 4) {
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
                tmem_st8(lane_addr + P_WHH_COL + 8 * j, r);
            }
            for (int j = wg; j < 2 * nk; j += 4) {
                const int pl = j >= nk, kj = pl ? j - nk : j;
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
                        } else if (terms == 3) {
                            r[4 * i + e] = lw[e];
                        } else {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                            const __half2 wp = __floats2half2_rn(fmaf(fl.x, 1.f / kPlaneScale, fh.x), fmaf(fl.y, 1.f / kPlaneScale, fh.y));
                            r[4 * i + e] = *reinterpret_cast<const uint32_t*>(&wp);
                        }
                    }
                }
                tmem_st8(lane_addr + P_WIH_COL + 8 * j, r);
            }
            tmem_st_wait();
        }
        loaded_dir = dir;
        tc_fence_before();
        cluster_sync_all();                                    // the leader's MMAs read BOTH CTAs' weights
        tc_fence_after();

        const int npairs = (nparts + 1) >> 1;
        if (warp == P_W_PROD) {
            // ===================== TMA producer: this CTA's half of the x tile of (step, pair of parts) =====================
            // the two CTAs with the same half (ranks hf and hf + 2) need the same 16 sequences: each fetches every other box and
            // multicasts it to both
            if (elect_one()) {
                const int nboxes = 2 * p.kblocks;
                const int sub = (int)(rank >> 1);
                const uint16_t mc = (uint16_t)(5u << hf);
                for (int s = 0; s < T; ++s) {
                    const int t = dir == 0 ? s : T - 1 - s;
                    for (int pp = 0; pp < npairs; ++pp) {
                        PWAIT(bar_x_empty(xst), xph ^ 1u, 1);
                        if (PROBE && (p.opt & 32)) {                                          // timing probe: no x loads (stale tiles)
                            mbar_arrive(bar_x_full(pp & 1, xst));
                            if (++xst == p.stages) { xst = 0; xph ^= 1u; }
                            continue;
                        }
                        mbar_expect_tx(bar_x_full(pp & 1, xst), stage_bytes);
                        const uint32_t dst = smem_base + x_off + xst * stage_bytes;
                        for (int bi = sub; bi < nboxes; bi += 2) {
                            const int pl = bi >= p.kblocks, kb = pl ? bi - p.kblocks : bi;
                            tma_load_3d_mc(dst + bi * P_XBOX, pl ? &tm_b : &tm_a, kb * 64, t, seq0 + pp * 2 * PPN + (int)hf * PPN, bar_x_full(pp & 1, xst), mc);
                        }
                        if (++xst == p.stages) { xst = 0; xph ^= 1u; }
                    }
                }
            }
        } else if (warp == P_W_MMAX || warp == P_W_MMAX + 1) {
            if (elect_one()) {
                const int me = warp - P_W_MMAX;
                if (is_leader) {
                    // ===================== input-product issuer: acc(s, pair) = W_ih . x_s  (M = 256, N = 32) =====================
                    constexpr uint32_t idesc = idesc_f16(256, 2 * PPN);
                    const uint32_t wa0 = tmem_base + P_WIH_COL, wb0 = wa0 + 8 * nk;
                    const uint32_t plane_lo = (uint32_t)(p.kblocks * P_XBOX) >> 4;
                    const uint16_t self_mask = (uint16_t)(1u << rank);
                    for (int s = 0; s < T; ++s) {
                        for (int pp = 0; pp < npairs; ++pp) {
                            if ((pp & 1) != me) {
                                if (++xst == p.stages) { xst = 0; xph ^= 1u; }
                                continue;
                            }
                            if (s > 0) {                                           // both CTAs' pointwise warps have read acc(s - 1, .)
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const int q = 2 * pp + h;
                                    if (q < nparts) {
                                        PWAIT_CL(bar_acc_free(q), (ph_b >> q) & 1u, 2);
                                        ph_b ^= 1u << q;
                                    }
                                }
                            }
                            PWAIT_CL(bar_x_full(me, xst), (xfph >> xst) & 1u, 3);
                            xfph ^= 1u << xst;
                            tc_fence_after();
                            const uint32_t xa = pdesc_lo(smem_base + x_off + xst * stage_bytes), xb = xa + plane_lo;
                            const uint32_t d = tmem_base + P_ACC_COL + pp * 2 * PPN;
                            if (NK > 0) {
#pragma unroll
                                for (int j = 0; j < (NK > 0 ? NK : 1); ++j) {
                                    if (PROBE && (p.opt & 4) && j > 0) break;               // timing probe: one k-step of the input product
                                    if (opt_yield && (j & 1) == 0) {               // the recurrent products are the critical path: let them pass
                                        uint32_t b0, b1;
                                        do {
                                            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(busy_flags) : "memory");
                                        } while (b0 | b1);
                                    }
                                    const uint32_t off = (uint32_t)((j >> 2) * (P_XBOX >> 4) + (j & 3) * 2);
                                    if (TERMS == 3) {
                                        mma_f16_ts_pair(d, wa0 + 8 * j, pdesc(xb + off), idesc, j != 0);
                                        mma_f16_ts_pair(d, wb0 + 8 * j, pdesc(xa + off), idesc, 1);
                                    } else {
                                        mma_f16_ts_pair(d, wb0 + 8 * j, pdesc(xb + off), idesc, j != 0);
                                    }
                                    mma_f16_ts_pair(d, wa0 + 8 * j, pdesc(xa + off), idesc, 1);
                                }
                            } else {
                                for (int j = 0; j < nk; ++j) {
                                    const uint32_t off = (uint32_t)((j >> 2) * (P_XBOX >> 4) + (j & 3) * 2);
                                    if (terms == 3) {
                                        mma_f16_ts_pair(d, wa0 + 8 * j, pdesc(xb + off), idesc, j != 0);
                                        mma_f16_ts_pair(d, wb0 + 8 * j, pdesc(xa + off), idesc, 1);
                                    } else {
                                        mma_f16_ts_pair(d, wb0 + 8 * j, pdesc(xb + off), idesc, j != 0);
                                    }
                                    mma_f16_ts_pair(d, wa0 + 8 * j, pdesc(xa + off), idesc, 1);
                                }
                            }
                            mma_commit_pair_mc(bar_x_empty(xst), (uint16_t)0xF);
                            mma_commit_pair_mc(bar_x_done(pp), self_mask);
                            if (++xst == p.stages) { xst = 0; xph ^= 1u; }
                        }
                    }
                    for (int q = 0; q < nparts; ++q) {
                        if (((q >> 1) & 1) != me) continue;
                        PWAIT_CL(bar_acc_free(q), (ph_b >> q) & 1u, 5);
                        ph_b ^= 1u << q;
                    }
                } else {
                    // ===================== x forwarder (odd CTA): own half tile landed -> one arrive on the leader's x_full =====================
                    const uint32_t remote0 = mapa_shared(bar_x_full(me, 0), leader);
                    for (int s = 0; s < T; ++s) {
                        for (int pp = 0; pp < npairs; ++pp) {
                            if ((pp & 1) == me) {
                                PWAIT(bar_x_full(me, xst), (xfph >> xst) & 1u, 3);
                                xfph ^= 1u << xst;
                                mbar_arrive_cluster(remote0 + 8 * xst);
                            }
                            if (++xst == p.stages) { xst = 0; xph ^= 1u; }
                        }
                    }
                }
            }
        } else if (warp == P_W_MMA || warp == P_W_MMA + 1) {
            if (elect_one()) {
                const int me = warp - P_W_MMA;
                // ph_a: h_ready phase bits [set 2][part 8]; a barrier is armed (expect_tx) before the step whose pointwise pass fills its tile
                for (int q = me; q < nparts; q += 2) {
                    if (T > 1) mbar_expect_tx(bar_h_ready(0, q), P_HTILE);
                    if (opt_hdouble && T > 2) mbar_expect_tx(bar_h_ready(1, q), P_HTILE);
                }
                if (is_leader) {
                    // ===================== recurrent-product issuer: acc(s, q) += W_hh . h_{s-1}  (M = 256, N = 16) =====================
                    constexpr uint32_t idesc = idesc_f16(256, PPN);
                    for (int s = 0; s < T; ++s) {
                        for (int q = me; q < nparts; q += 2) {
                            PWAIT(bar_x_done(q >> 1), (ph_b >> (q >> 1)) & 1u, 2);
                            ph_b ^= 1u << (q >> 1);
                            if (s > 0) {
                                const int hs = opt_hdouble ? ((s - 1) & 1) : 0;
                                PWAIT_CL(bar_h_ready(hs, q), (ph_a >> (8 * hs + q)) & 1u, 4);   // own half (tx) and the odd CTA's forward
                                ph_a ^= 1u << (8 * hs + q);
                                if ((opt_hdouble ? s + 1 : s) < T - 1) mbar_expect_tx(bar_h_ready(hs, q), P_HTILE);   // its next use
                                fence_proxy_async();
                                tc_fence_after();
                                const uint32_t hb = pdesc_lo(smem_base + h_off + ((opt_hdouble ? ((s - 1) & 1) * PMAXP : 0) + q) * P_HTILE);
                                const uint32_t d = tmem_base + P_ACC_COL + q * PPN;
                                if (opt_yield) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(busy_flags + 4 * me), "r"(1u) : "memory");
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj)
                                    if (!(PROBE && (p.opt & 8)))                              // timing probe: no recurrent MMAs
                                    mma_f16_ts_pair(d, tmem_base + P_WHH_COL + 8 * jj, pdesc(hb + (uint32_t)((jj >> 2) * (P_HBOX >> 4) + (jj & 3) * 2)), idesc, 1);
                                if (opt_yield) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(busy_flags + 4 * me), "r"(0u) : "memory");
                            }
                            mma_commit_pair_mc(bar_acc_ready(q), pair_mask);
                            // single h tile set: every CTA's copy of the part's tile may be overwritten once both pairs' MMAs have read it.
                            // With two sets the hand-shake is implied: h_s exists only after every pair's MMAs of step s - 1 (the last
                            // readers of the set it goes to) completed, because each CTA's h_{s-1} -- an input of ALL MMAs of step s --
                            // is computed from an accumulator of step s - 1.
                            if (!opt_hdouble && s < T - 1) mma_commit_pair_mc(bar_h_free(q), (uint16_t)0xF);
                        }
                    }
                } else {
                    // ===================== h forwarder (odd CTA): own half of h_{s-1} landed -> one arrive on the leader's h_ready =====================
                    const uint32_t remote0 = mapa_shared(bar_h_ready(0, 0), leader), remote1 = mapa_shared(bar_h_ready(1, 0), leader);
                    for (int s = 1; s < T; ++s) {
                        for (int q = me; q < nparts; q += 2) {
                            const int hs = opt_hdouble ? ((s - 1) & 1) : 0;
                            PWAIT(bar_h_ready(hs, q), (ph_a >> (8 * hs + q)) & 1u, 4);
                            ph_a ^= 1u << (8 * hs + q);
                            if ((opt_hdouble ? s + 1 : s) < T - 1) mbar_expect_tx(bar_h_ready(hs, q), P_HTILE);
                            fence_proxy_async();
                            mbar_arrive_cluster((hs ? remote1 : remote0) + 8 * q);
                        }
                    }
                }
            }
        } else if (warp >= P_W_SEND) {
            // ===================== exchange senders: own 2 KB slice of a part's h_s -> k-block `rank` of the part's tile =====================
            // rows (sequences) 0-7 go to the CTAs with hf = 0 (ranks 0, 2), rows 8-15 to those with hf = 1 (ranks 1, 3); the copy
            // into the own tile is a plain shared-memory store
            uint32_t cta_delta[PC];
#pragma unroll
            for (uint32_t d = 0; d < PC; ++d) cta_delta[d] = mapa_shared(smem_base, d) - smem_base;
            for (int s = 0; s + 1 < T; ++s) {
                for (int q = warp - P_W_SEND; q < nparts; q += 4) {
                    if (lane == 0) {
                        PWAIT(bar_slice(q), (ph_a >> q) & 1u, 1);
                        if (!opt_hdouble) PWAIT(bar_h_free(q), (ph_b >> q) & 1u, 7);
                    }
                    ph_a ^= 1u << q;
                    ph_b ^= 1u << q;
                    __syncwarp();
                    const uint32_t src = smem_base + g_off + ((s & 1) * PMAXP + q) * P_SLICE + lane * 16;
                    const uint32_t dst = smem_base + h_off + ((opt_hdouble ? (s & 1) * PMAXP : 0) + q) * P_HTILE + rank * P_HBOX + lane * 16;
                    const uint32_t bar = bar_h_ready(opt_hdouble ? (s & 1) : 0, q);
#pragma unroll
                    for (int c = 0; c < P_SLICE / 512; ++c) {                      // 512 B = 4 sequences
                        uint4 v;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + c * 512));
                        const uint32_t da = dst + (c & 1) * 512;
#pragma unroll
                        for (uint32_t d = 0; d < 2; ++d) {
                            const uint32_t dr = (uint32_t)(c >> 1) + 2 * d;       // destination rank (static)
                            if (dr == rank)
                                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(da), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                            else
                                st_async_v4(da + cta_delta[dr], v, bar + cta_delta[dr]);
                        }
                    }
                    fence_proxy_async();                                           // own copy: generic stores before the tensor cores' reads
                    __syncwarp();
                    if (lane == 0) mbar_complete_tx(bar, P_HBOX);
                }
            }
        } else if ((warp >> 2) < nparts) {
            // ===================== pointwise warps (as lstm_fused.cu) =====================
            const int pg = warp >> 2, g = warp & 3;
            const int uu = lane >> 2, jj = lane & 3;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(g * 32) << 16) + P_ACC_COL;
            const float bias = __ldg(p.bias + dir * kGates + jj * kHidden + (int)rank * PU + 8 * g + uu);
            unsigned char* const scr = smem_gen + c_off + warp * P_SCRATCH;
            auto scr_at = [&](int row, int c) -> unsigned char* {
                return scr + (((row * 64) + ((c ^ ((row >> 1) & 3)) << 4)) ^ (((row >> 2) & 1) << 6));
            };
            const uint32_t acc_free0 = mapa_shared(bar_acc_free(0), leader);
            const float s1 = 1.f - kPlaneScale;
            const float L2E2 = 2.f * kLog2e;
            float cst[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) cst[a][b] = 0.f;
            const int yrow = lane & 15;
            __half* const yplane = (lane >> 4) ? p.y_b : p.y_a;
            for (int s = 0; s < T; ++s) {
                const int t = dir == 0 ? s : T - 1 - s;
                const bool exchange = s + 1 < T;
#pragma unroll
                for (int pi = 0; pi < 2; ++pi) {
                    const int q = pg + 4 * pi;
                    if (q >= nparts) break;
                    PWAIT(bar_acc_ready(q), (ph_a >> pi) & 1u, 6);
                    ph_a ^= 1u << pi;
                    tc_fence_after();
                    float z[16];
                    tmem_ld16(lane_addr + q * PPN, z);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_relaxed(acc_free0 + 8 * q);
#pragma unroll
                    for (int j = 0; j < 16; ++j) z[j] = fast_ex2(fminf(z[j] + bias, 29.f));
                    float ev[4][4];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<float4*>(scr_at(lane, c)) = make_float4(z[4 * c], z[4 * c + 1], z[4 * c + 2], z[4 * c + 3]);
                    __syncwarp();
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float4 v = *reinterpret_cast<const float4*>(scr_at((lane & ~3) + b, jj));
                        ev[b][0] = v.x; ev[b][1] = v.y; ev[b][2] = v.z; ev[b][3] = v.w;
                    }
                    __syncwarp();
                    float hv[4];
                    const f32x2 one = pack2(1.f, 1.f), mone = pack2(-1.f, -1.f), k2 = pack2(L2E2, L2E2);
#pragma unroll
                    for (int jp = 0; jp < 2; ++jp) {
                        const f32x2 pei = pack2(ev[0][2 * jp], ev[0][2 * jp + 1]), pef = pack2(ev[1][2 * jp], ev[1][2 * jp + 1]);
                        const f32x2 peg = pack2(ev[2][2 * jp], ev[2][2 * jp + 1]), peo = pack2(ev[3][2 * jp], ev[3][2 * jp + 1]);
                        const f32x2 di = add2(pei, one), df = add2(pef, one), dg = add2(peg, one);
                        const f32x2 dig = mul2(di, dg);
                        const f32x2 den = mul2(df, dig);
                        const f32x2 cn = mul2(fma2(pack2(cst[pi][2 * jp], cst[pi][2 * jp + 1]), dig, mul2(add2(peg, mone), df)), rcp2(den));
                        unpack2(cn, cst[pi][2 * jp], cst[pi][2 * jp + 1]);
                        const f32x2 ec = ex2_clamped2(mul2(cn, k2));
                        const f32x2 h2v = mul2(add2(ec, mone), rcp2(mul2(add2(peo, one), add2(ec, one))));
                        unpack2(h2v, hv[2 * jp], hv[2 * jp + 1]);
                    }
                    unsigned char* const stg = smem_gen + g_off + ((s & 1) * PMAXP + q) * P_SLICE;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int row = 4 * jj + c;
                        __half a1, a2;
                        split_scaled_f16(hv[c], s1, a1, a2);
                        unsigned char* const rp = stg + row * 128 + uu * 2;
                        *reinterpret_cast<__half*>(rp + ((g ^ (row & 7)) << 4)) = a1;
                        *reinterpret_cast<__half*>(rp + (((4 + g) ^ (row & 7)) << 4)) = a2;
                    }
                    __syncwarp();
                    if (exchange && lane == 0) mbar_arrive(bar_slice(q));
                    const uint4 v = *reinterpret_cast<const uint4*>(stg + yrow * 128 + ((((lane >> 4) * 4 + g) ^ (yrow & 7)) << 4));
                    const int b = seq0 + q * PPN + yrow;
                    if (b < p.B)
                        *reinterpret_cast<uint4*>(yplane + ((size_t)b * T + t) * (2 * kHidden) + dir * kHidden + (int)rank * PU + 8 * g) = v;
                }
            }
        }
        tc_fence_before();
        cluster_sync_all();
        tc_fence_after();
    }
    // drain: the last `stages` multicast commits onto this CTA's x_empty barriers must have landed before it exits (lstm_fused.cu)
    if (warp == P_W_PROD && elect_one()) {
        for (int k = 0; k < p.stages; ++k) {
            PWAIT(bar_x_empty(xst), xph ^ 1u, 1);
            if (++xst == p.stages) { xst = 0; xph ^= 1u; }
        }
    }
    if (wacc) wacc[0] += clock64();
    tc_fence_before();
    cluster_sync_all();
    if (warp == P_W_MMA) tmem_dealloc_pair<512>(tmem_base);
}

// ---------------------------------------------------------------- launcher
static int* g_pair_err_hostbuf = nullptr;
int lstm_pair_last_timeout(int* out7) {
    for (int i = 0; i < 7; ++i) out7[i] = g_pair_err_hostbuf ? g_pair_err_hostbuf[i] : 0;
    return B200VAD_OK;
}
static int g_pair_opt = 3;
void lstm_pair_set_opt(int opt) { g_pair_opt = opt; }
static int g_pair_clusters[64];
static std::once_flag g_pair_once[64];

typedef void (*PairKern)(CUtensorMap, CUtensorMap, PairParams);
static PairKern pair_pick(int nk, int terms, int probe) {
    if (probe) {                                             // timing probes of the bench shapes (tools/pair_waits.py, tools/lstm_modes_timing.py)
        if (nk == 16 && terms == 2) return lstm_pair_kernel<16, 2, true>;
        if (nk == 5 && terms == 3) return lstm_pair_kernel<5, 3, true>;
    }
    if (nk == 16) return terms == 2 ? lstm_pair_kernel<16, 2, false> : lstm_pair_kernel<16, 3, false>;
    if (nk == 5) return terms == 2 ? lstm_pair_kernel<5, 2, false> : lstm_pair_kernel<5, 3, false>;
    if (nk == 4) return terms == 2 ? lstm_pair_kernel<4, 2, false> : lstm_pair_kernel<4, 3, false>;
    return lstm_pair_kernel<0, 0, false>;
}
int lstm_pair_read_debug(long long* host, int n) {
    const size_t bytes = sizeof(long long) * (size_t)std::min<long long>(n, (long long)P_DBG_MAX_CTAS * P_DBG_WARPS * P_DBG_TAGS * 2);
    B200VAD_CUDA(cudaMemcpyFromSymbol(host, g_pair_dbg, bytes));
    return B200VAD_OK;
}

static int pair_smem_bytes(int kblocks, int* stages_out) {
    const int stage = 2 * kblocks * P_XBOX;
    int stages = (P_SMEM_MAX - 1024 - P_SMEM_FIXED - 1024) / stage;
    if (stages > P_MAX_STAGES) stages = P_MAX_STAGES;
    *stages_out = stages;
    return 1024 + P_SMEM_FIXED + stages * stage + 1024;
}

static int pair_max_clusters() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::call_once(g_pair_once[dev], [&] {
        const void* fn = reinterpret_cast<const void*>(lstm_pair_kernel<16, 2, false>);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(PC * 64);
        cfg.blockDim = dim3(P_THREADS);
        cfg.dynamicSmemBytes = P_SMEM_MAX - 1024;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = PC; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = 0;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (set_max_dynamic_smem(fn, P_SMEM_MAX - 1024) != B200VAD_OK ||
            cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = std::max(1, sms / PC - 4);
        }
        const char* e = getenv("B200VAD_FUSED_CLUSTERS");
        if (e && atoi(e) > 0) n = atoi(e);
        g_pair_clusters[dev] = n;
    });
    return g_pair_clusters[dev];
}

int lstm_pair_launch(const __half* x_a, const __half* x_b, int64_t lda, int B, int T, int D, const __half* wih_hi,
                     const __half* wih_lo, int ldw, const __half* whh_hi, const __half* whh_lo, const float* bias, int terms,
                     __half* y_a, __half* y_b, cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    if (!lstm_fused_supported(D) || (terms != 2 && terms != 3) || lda % 8 != 0 || ldw % 8 != 0 || ldw < (D + 15) / 16 * 16) {
        set_error("lstm_pair: unsupported shape (D=%d lda=%lld ldw=%d terms=%d)", D, (long long)lda, ldw, terms);
        return B200VAD_EINVAL;
    }
    PairParams p;
    p.wih_hi = wih_hi; p.wih_lo = wih_lo; p.whh_hi = whh_hi; p.whh_lo = whh_lo; p.bias = bias; p.y_a = y_a; p.y_b = y_b;
    p.B = B; p.T = T; p.nk = (D + 15) / 16; p.kblocks = (D + 63) / 64; p.ldw = ldw; p.terms = terms;
    int stages = 0;
    const int smem = pair_smem_bytes(p.kblocks, &stages);
    p.stages = stages;
    static int opt_env = -2;
    if (opt_env == -2) { const char* e = getenv("B200VAD_PAIR_OPT"); opt_env = e ? atoi(e) : -1; }
    p.opt = opt_env >= 0 ? opt_env : g_pair_opt;
    const int nc = pair_max_clusters();
    const int P = (B + PPN - 1) / PPN;
    int best_ipd = (P + PMAXP - 1) / PMAXP;
    double best_cost = 1e30;
    static const double kStepUs[PMAXP + 1] = {0.0, 1.60, 1.93, 2.17, 2.41, 2.78, 3.16, 3.53, 3.90};   // step time by parts in flight (lstm_fused.cu)
    for (int ipd = (P + PMAXP - 1) / PMAXP; ipd <= P; ++ipd) {
        const int maxp = (P + ipd - 1) / ipd;
        const int waves = (2 * ipd + nc - 1) / nc;
        const double cost = waves * kStepUs[maxp];
        if (cost < best_cost - 1e-9) { best_cost = cost; best_ipd = ipd; }
        if (maxp == 1) break;
    }
    p.items_per_dir = best_ipd;
    const int grid = PC * std::min(nc, 2 * best_ipd);
    CUtensorMap tm_a, tm_b;
    int rc;
    if ((rc = make_tmap_3d(&tm_a, x_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, D, T, B, lda * 2, (uint64_t)T * lda * 2, 64, 1, PPN,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_b, x_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, D, T, B, lda * 2, (uint64_t)T * lda * 2, 64, 1, PPN,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const PairKern kern = pair_pick(p.nk, terms, (p.opt & ~3) != 0);
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(kern), smem))) return rc;
    if (!g_pair_err_hostbuf) {
        int* hb = nullptr;
        if (cudaHostAlloc(&hb, 64, cudaHostAllocMapped) == cudaSuccess) {
            for (int i = 0; i < 16; ++i) hb[i] = 0;
            int* dp = nullptr;
            if (cudaHostGetDevicePointer(&dp, hb, 0) == cudaSuccess &&
                cudaMemcpyToSymbol(g_pair_err_host, &dp, sizeof(dp)) == cudaSuccess) g_pair_err_hostbuf = hb;
        }
        cudaGetLastError();
    }
    prof_begin(0, st);
    kern<<<grid, P_THREADS, smem, st>>>(tm_a, tm_b, p);
    prof_end(0, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
