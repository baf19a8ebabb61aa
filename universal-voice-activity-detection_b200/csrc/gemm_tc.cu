// tcgen05 projection GEMM:  C[M,N] = X_hi.W_hi^T + X_lo.W_hi^T (+ X_hi.W_lo^T) + bias   (fp32 accumulate in TMEM)
//
// Computes the LSTM input projections x_t.W_ih^T + b_ih + b_hh for all timesteps and both
// directions (nn.LSTM of PyanNet2.py:95,170) and the head linears (PyanNet2.py:183-185) in split
// precision: activations arrive as two fp16 planes (hi, lo: hi + lo = the fp32 value to ~22
// bits), weights as fp16 hi + lo planes.
//
// The OUTPUT FEATURE dimension is the MMA M: a CTA owns 128 output features for its whole life and
// keeps their weight rows (hi and lo planes, up to K = 256) RESIDENT IN TENSOR MEMORY as the A
// operand (lane = feature, two fp16 k-elements per column).  Activation rows are the MMA N: 128-row
// tiles stream through a TMA ring as the (K-major, 128B-swizzled) B operand.  An SS-mode MMA would
// re-read a 4 KB weight slice from shared memory per instruction; with the weights in TMEM the
// shared-memory traffic is the streamed operand only (v1 of this kernel kept the weights in smem
// and ran at 45 % tensor-pipe utilisation, bound by shared-memory bandwidth -- profiles/r01_gemm.md).
// Persistent, warp-specialised, one CTA per SM: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
// (single thread), warps 2-5 = epilogue.  The 128 x 128 fp32 accumulator is double-buffered in TMEM,
// so the epilogue of tile i overlaps the MMAs of tile i+1.  In the accumulator a TMEM lane is an
// output feature and a column is a row, so the epilogue thread of lane f writes feature f of 32
// consecutive rows: each store instruction of a warp covers one contiguous 128-byte run of a row --
// straight from registers, no shared-memory staging.  The n_blocks CTAs that share an activation
// tile run concurrently, so it is fetched from HBM once and hits L2 for the others.
#include "kernels.cuh"
#include "tc05.cuh"

namespace b200vad {

using namespace tc;

constexpr int SBM = 128;                 // activation rows per tile (UMMA N)
constexpr int SBK = 64;                  // k per smem tile (128 bytes of fp16 = one swizzle atom row)
constexpr int S_TILE_BYTES = SBM * SBK * 2;          // 16 KB
constexpr int GEMM_TS_THREADS = 192;
constexpr int ACC_COL = 256;             // TMEM columns [0, 256): weights, [256, 512): two 128-column accumulators

struct GemmTsParams {
    const __half* w_hi;      // [N][Kp] fp16, zero padded
    const __half* w_lo;      // [N][Kp] or null
    const float* bias;       // [N] or null
    float* c;                // modes 0, 2: [M][ldc] fp32
    __half* o_hi;            // mode 1: [M][ldc] fp16 planes
    __half* o_lo;
    int64_t ldc;
    int64_t M;
    int num_m_tiles;
    int ldw;                 // weight row pitch in elements (>= Kp)
    int accumulate;          // modes 0, 3: add to the existing C (K split over several launches); bias is then ignored
    int T, tiles_per_blk;    // mode 3: rows are (sequence, step) pairs; a tile = 64 sequences x 2 steps
    const float* wc;         // mode 4: classifier weight [128] and bias [1]; c = prob [M]
    const float* bc;
    // batched overlapping-row activations (the SincNet convolutions): a tile = 128 rows of one batch item, fetched as a
    // 3-D box; tile row j of tile (b, rt) is output row  b * out_batch_rows + (rt * 128 + j) * out_row_step + out_row_off
    int rows3d, rows_per_batch, tiles_per_batch;
    int64_t out_batch_rows;
    int out_row_step, out_row_off;
    int n_valid;             // features >= n_valid are padding (weights zero) and are not stored
    int act_abs;             // mode 0: store |x| (after the last K pass)
    int n_blocks;            // N / 128
    int kb;                  // k-blocks of 64
    int nw;                  // weight planes: 1 (hi) or 2 (hi, lo)
    int stages;              // activation pipeline depth
    int Kp;
};

// MODE 0: C = acc + bias (fp32)      MODE 1: leaky_relu(acc + bias) -> fp16 (hi, lo) planes      MODE 2: leaky_relu -> fp32
// MODE 4: leaky_relu, then the classifier of PyanNet2.py:187 fused into the epilogue: prob[row] = sigmoid(wc . z + bc)
//   (N must be 128: one CTA owns all features of a row; the dot product is a register transpose-reduction over the
//   32 lanes of a warp plus a 4-warp exchange through shared memory).
// MODE 3: LSTM input projection in the recurrence's step-blocked layout.  Activations are a (B, T, K) tensor; a
//   tile is 64 sequences x 2 steps (tile row j = sequence * 2 + step, fetched as one 3-D TMA box), and
//   C = xg[sequence block][direction][t][gate][column group 16][unit 128][4 columns] fp32, feature = (direction,
//   gate, unit): every (block, direction, t) is one contiguous 128 KB record, which is what one recurrence CTA
//   consumes per step; the epilogue thread of a unit writes 4 sequences of one step as one 16-byte store.
template <int MODE>
__global__ void __launch_bounds__(GEMM_TS_THREADS, 1)
gemm_ts_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo, GemmTsParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;                                        // [stages][hi, lo] tiles of 128 x 64
    const uint32_t bar_base = a_base + p.stages * 2 * S_TILE_BYTES;
    auto bar_a_full = [&](int s) { return bar_base + 8 * s; };
    auto bar_a_empty = [&](int s) { return bar_base + 64 + 8 * s; };
    auto bar_acc_full = [&](int b) { return bar_base + 128 + 8 * b; };
    auto bar_acc_empty = [&](int b) { return bar_base + 144 + 8 * b; };
    const uint32_t bar_w = bar_base + 160;
    const uint32_t tmem_slot = bar_base + 168;
    const uint32_t part_smem = bar_base + 192;                                // mode 4: [2][4 warps][32 rows] fp32 partial dots

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x % p.n_blocks;
    const int tile0 = blockIdx.x / p.n_blocks;
    const int tile_step = gridDim.x / p.n_blocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), 128); }
        mbar_init(bar_w, 128);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int wcols = p.Kp >> 1;                                              // TMEM columns per weight plane

    if (warp == 0) {
        // ===================== TMA producer: activation tiles =====================
        if (elect_one()) {
            tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
            int s = 0;
            uint32_t ph = 0;
            for (int t = tile0; t < p.num_m_tiles; t += tile_step) {
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_a_full(s), 2 * S_TILE_BYTES);
                    if (MODE == 3) {
                        const int bblk = t / p.tiles_per_blk, tp = t - bblk * p.tiles_per_blk;
                        tma_load_3d(a_base + (2 * s) * S_TILE_BYTES, &tm_a_hi, kb * SBK, tp * 2, bblk * 64, bar_a_full(s));
                        tma_load_3d(a_base + (2 * s + 1) * S_TILE_BYTES, &tm_a_lo, kb * SBK, tp * 2, bblk * 64, bar_a_full(s));
                    } else if (p.rows3d) {
                        const int bi = t / p.tiles_per_batch, rt = t - bi * p.tiles_per_batch;
                        tma_load_3d(a_base + (2 * s) * S_TILE_BYTES, &tm_a_hi, kb * SBK, rt * SBM, bi, bar_a_full(s));
                        tma_load_3d(a_base + (2 * s + 1) * S_TILE_BYTES, &tm_a_lo, kb * SBK, rt * SBM, bi, bar_a_full(s));
                    } else {
                        tma_load_2d(a_base + (2 * s) * S_TILE_BYTES, &tm_a_hi, kb * SBK, t * SBM, bar_a_full(s));
                        tma_load_2d(a_base + (2 * s + 1) * S_TILE_BYTES, &tm_a_lo, kb * SBK, t * SBM, bar_a_full(s));
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: D[out, row] += W[out, k] . X[row, k] =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(128, SBM);
            mbar_wait(bar_w, 0);                                                // weights are in TMEM
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = tile0; t < p.num_m_tiles; t += tile_step, ++it) {
                const int ab = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(bar_acc_empty(ab), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ACC_COL + ab * SBM;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t x_hi = a_base + (2 * s) * S_TILE_BYTES, x_lo = x_hi + S_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        const uint32_t w_hi = tmem_base + (kb * 4 + k) * 8, w_lo = w_hi + wcols;
                        const uint64_t dx_hi = smem_desc_sw128(x_hi + k * 32), dx_lo = smem_desc_sw128(x_lo + k * 32);
                        mma_f16_ts(d_tmem, w_hi, dx_lo, idesc, (kb | k) != 0);            // small terms first
                        if (p.nw == 2) mma_f16_ts(d_tmem, w_lo, dx_hi, idesc, 1);
                        mma_f16_ts(d_tmem, w_hi, dx_hi, idesc, 1);
                    }
                    mma_commit(bar_a_empty(s));                                 // frees the stage when the MMAs retire
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                mma_commit(bar_acc_full(ab));                                   // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== weights -> TMEM, then epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
        const int q = warp & 3;
        const int out = blk * 128 + q * 32 + lane;                              // output feature == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int w = 0; w < p.nw; ++w) {
            const uint4* wrow = reinterpret_cast<const uint4*>((w == 0 ? p.w_hi : p.w_lo) + (size_t)out * p.ldw);
            for (int part = 0; part < p.Kp / 32; ++part) {                      // 32 fp16 = 16 packed columns
                uint32_t r[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 v = __ldg(wrow + part * 4 + i);
                    r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
                }
                tmem_st16(lane_addr + w * wcols + part * 16, r);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_w);
        const float bias = (p.bias && !p.accumulate && out < p.n_valid) ? __ldg(p.bias + out) : 0.f;
        int it = 0;
        for (int t = tile0; t < p.num_m_tiles; t += tile_step, ++it) {
            const int ab = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(bar_acc_full(ab), acc_ph);
            tc_fence_after();
            if (MODE == 3) {
                // feature = (dir, gate, unit): blk = dir * 4 + gate, unit = q * 32 + lane
                const int bblk = t / p.tiles_per_blk, tp = t - bblk * p.tiles_per_blk;
                const int dir = blk >> 2, gate = blk & 3;
                const int t0 = tp * 2;
                // float offset of (bblk, dir, t0, gate, cgrp 0, unit, 0); one step further = 4 * 8192 floats
                float* base = p.c + ((((int64_t)bblk * 2 + dir) * p.T + t0) * 4 + gate) * 8192 + (q * 32 + lane) * 4;
                const bool t1_ok = t0 + 1 < p.T;
#pragma unroll 1
                for (int c = 0; c < SBM / 32; ++c) {
                    float v[32];
                    tmem_ld32(lane_addr + ACC_COL + ab * SBM + c * 32, v);
                    tmem_ld_wait();
                    if (c == SBM / 32 - 1) {
                        tc_fence_before();
                        mbar_arrive(bar_acc_empty(ab));
                    }
                    // accumulator column j = sequence * 2 + step; 8 columns = 4 sequences (one column group) x 2 steps
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
#pragma unroll
                        for (int tl = 0; tl < 2; ++tl) {
                            if (tl == 1 && !t1_ok) continue;
                            float4* dst = reinterpret_cast<float4*>(base + (int64_t)tl * 4 * 8192 + ((c * 4 + g4) * 128) * 4);
                            float4 o = make_float4(v[8 * g4 + tl] + bias, v[8 * g4 + 2 + tl] + bias, v[8 * g4 + 4 + tl] + bias,
                                                   v[8 * g4 + 6 + tl] + bias);
                            if (p.accumulate) {
                                const float4 old = *dst;
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *dst = o;
                        }
                    }
                }
                continue;
            }
            int64_t row0 = (int64_t)t * SBM;
            int nrows = (int)min((int64_t)SBM, p.M - row0);
            int64_t row_step = 1;
            if (p.rows3d) {
                const int bi = t / p.tiles_per_batch, rt = t - bi * p.tiles_per_batch;
                nrows = min(SBM, p.rows_per_batch - rt * SBM);
                row_step = p.out_row_step;
                row0 = (int64_t)bi * p.out_batch_rows + (int64_t)rt * SBM * row_step + p.out_row_off;
            }
            const bool out_ok = out < p.n_valid;
#pragma unroll 1
            for (int c = 0; c < SBM / 32; ++c) {
                float v[32];
                tmem_ld32(lane_addr + ACC_COL + ab * SBM + c * 32, v);
                tmem_ld_wait();
                if (c == SBM / 32 - 1) {
                    tc_fence_before();
                    mbar_arrive(bar_acc_empty(ab));                             // 128 arrivals release the accumulator
                }
                if (MODE == 4) {
                    // v[j] <- wc[unit] * leaky_relu(z2[unit][row j]); then sum over the 128 units of every row
                    const float w = __ldg(p.wc + out);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[j] + bias;
                        v[j] = w * (x > 0.f ? x : 0.01f * x);
                    }
                    // transpose-reduction: after the step with offset o a lane keeps the rows whose bit o equals its own
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float keep = up ? v[i + o] : v[i];
                            const float send = up ? v[i] : v[i + o];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    // v[0] = this warp's 32-unit partial for row c * 32 + lane; combine the 4 warps through smem
                    const int pb = (it * (SBM / 32) + c) & 1;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(part_smem + 4 * ((pb * 4 + q) * 32 + lane)), "f"(v[0]) : "memory");
                    named_bar_sync(1, 128);
                    if (q == 0) {
                        float s = __ldg(p.bc);
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            float pv;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pv) : "r"(part_smem + 4 * ((pb * 4 + w4) * 32 + lane)));
                            s += pv;
                        }
                        if (c * 32 + lane < nrows) p.c[row0 + c * 32 + lane] = 1.f / (1.f + __expf(-s));
                    }
                    continue;
                }
                // lane = output feature, register j = row: every store instruction writes one contiguous run per warp.
                // Running pointers and a full-chunk fast path keep this at ~3 instructions per element: with one epilogue
                // warp per scheduler the instruction stream itself is the tile's critical path (profiles/r01_gemm.md).
                const int64_t base = (row0 + (int64_t)c * 32 * row_step) * p.ldc + out;
                const int64_t jstride = row_step * p.ldc;
                const int lim = out_ok ? nrows - c * 32 : 0;
                if (lim <= 0) continue;
                if (MODE == 0 && p.accumulate) {
                    // K split over two launches: fetch the 32 partial sums first (independent loads), then add
                    const float* src = p.c + base;
                    float old[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { old[j] = (j < lim) ? __ldcg(src) : 0.f; src += jstride; }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += old[j];
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = v[j] + bias;
                    if (MODE == 0 && p.act_abs) x = fabsf(x);
                    if (MODE != 0) x = x > 0.f ? x : 0.01f * x;
                    v[j] = x;
                }
                if (MODE == 1) {
                    __half* dh = p.o_hi + base;
                    __half* dl = p.o_lo + base;
                    if (lim >= 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { split_f16(v[j], *dh, *dl); dh += jstride; dl += jstride; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < lim) split_f16(v[j], *dh, *dl); dh += jstride; dl += jstride; }
                    }
                } else {
                    float* dst = p.c + base;
                    if (lim >= 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { *dst = v[j]; dst += jstride; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < lim) *dst = v[j]; dst += jstride; }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

__global__ void zero_i32_kernel(int* __restrict__ p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// ---------------------------------------------------------------- 2-MMA input projection (layers >= 1)
// xg = x . W^T with x and W as fp16 (hi, lo) planes needs three fp16 products for ~2^-21 accuracy (x_lo.W_hi + x_hi.W_lo
// + x_hi.W_hi) and that makes the K = 256 projections tensor bound.  Two products into ONE accumulator are enough when
// the activation planes are split with a scale s = 2^-6 instead of at the fp16 rounding point (lstm_tc_kernel writes
// them that way when asked):
//     x1 = fp16((1 - s) x),  x2 = fp16(x - x1) = s x + r      (r = rounding residual of x1, 2^-12 |x|; x1 + x2 = x to 2^-18)
//     W' = fp16(W_hi + W_lo / s)                              (once per CTA, on the way into TMEM)
//     x1.W_hi + x2.W' = x.((1 - s) W_hi + s W') + r.(W' - W_hi)
// (1 - s) W_hi + s W' equals W up to s * rounding(W') = 2^-18 |W|, and |W' - W_hi| = |W_lo| / s <= 2^-6 |W| multiplies
// r: every term is accurate to ~2^-18 relative (3-term: 2^-22; dropping W_lo: 2^-12).  The recurrence (W_hh single fp16)
// and the head's 3-term product consume the same (x1, x2) planes unchanged.  End to end the probabilities move by
// ~2.6e-5 relative (tools/two_mma_precision.py), 10x below the W_hh rounding that dominates.  Layer 0 (un-normalised
// log-mel input in hi / lo planes, HBM-write bound anyway) stays 3-term.
//
// Structure as gemm_ts_kernel<3> (tile = 64 sequences x 2 steps, weights in TMEM, double-buffered accumulator).  With
// 2/3 of the tensor work the kernel is bound by the 4 KB / frame xg write, and the 8 CTAs of a group (the 8 feature
// blocks) that read the same activation tiles drift apart unless they are kept in lockstep (see the producer).
// (A first version converted hi / lo planes to (x1, x2)-like operands in shared memory with a transform warpgroup: the
// extra 48 KB of shared-memory traffic per stage made it slower than the 3-term kernel -- profiles/r01_gemm.md.)
constexpr int X2_THREADS = 320;                       // TMA, MMA, 8 epilogue warps
constexpr int X2_STAGES = 6;
constexpr float X2_S = kPlaneScale;

struct GemmX2Params {
    const __half* w_hi;      // [1024][ldw]
    const __half* w_lo;
    const float* bias;       // [1024]
    float* xg;
    int T, tiles_per_blk, num_tiles, kb, ldw, terms;
    int* sync;               // [groups][sync_stride] window arrival counters (zeroed per launch) or null
    int sync_stride;
};
constexpr int X2_WINDOW = 2;                          // tiles per lockstep window

// Epilogue store of one 32-column accumulator chunk c of a (64 sequences x 2 steps) tile: accumulator column j =
// sequence * 2 + step; 8 columns = 4 sequences (one column group) x 2 steps; one 16-byte store = 4 sequences of a unit.
__device__ __forceinline__ void xg_store_chunk(float* base, int c, const float (&v)[32], float bias, bool t1_ok) {
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
            if (tl == 1 && !t1_ok) continue;
            float4* dst = reinterpret_cast<float4*>(base + (int64_t)tl * 4 * 8192 + ((c * 4 + g4) * 128) * 4);
            *dst = make_float4(v[8 * g4 + tl] + bias, v[8 * g4 + 2 + tl] + bias, v[8 * g4 + 4 + tl] + bias, v[8 * g4 + 6 + tl] + bias);
        }
    }
}

__global__ void __launch_bounds__(X2_THREADS, 1)
gemm_xg2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo, GemmX2Params p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;                                        // [stages][x1, x2] tiles of 128 x 64
    const uint32_t bar_base = a_base + X2_STAGES * 2 * S_TILE_BYTES;
    auto bar_a_full = [&](int s) { return bar_base + 8 * s; };                // TMA landed        (count 1 + tx)
    auto bar_a_empty = [&](int s) { return bar_base + 128 + 8 * s; };         // MMAs retired      (count 1)
    auto bar_acc_full = [&](int b) { return bar_base + 192 + 8 * b; };
    auto bar_acc_empty = [&](int b) { return bar_base + 208 + 8 * b; };
    const uint32_t bar_w = bar_base + 224;
    const uint32_t tmem_slot = bar_base + 232;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x & 7;                                           // feature block = dir * 4 + gate
    const int tile0 = blockIdx.x >> 3;
    const int tile_step = gridDim.x >> 3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < X2_STAGES; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), 256); }
        mbar_init(bar_w, 256);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int wcols = p.kb * 32;                                              // TMEM columns per weight plane (K / 2)

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
            int s = 0;
            uint32_t ph = 0;
            // The 8 CTAs of a group read the same activation tiles; only if they stay within a few tiles of each other does
            // L2 serve 7 of the 8 reads.  The kernel is memory bound, so free-running CTAs drift (measured: 15.5 GB of DRAM
            // reads for 3.4 GB of activations).  Lockstep: a producer announces every window of X2_WINDOW tiles it starts
            // and does not start window w + 1 before all 8 have started window w.  The wait is bounded (a late CTA only
            // costs bandwidth, never correctness), so co-residency of the group is not a requirement.
            int* const sync = p.sync ? p.sync + (size_t)tile0 * p.sync_stride : nullptr;
            int ti = 0;
            bool lock_ok = true;
            for (int t = tile0; t < p.num_tiles; t += tile_step, ++ti) {
                if (sync && (ti % X2_WINDOW) == 0) {
                    const int w = ti / X2_WINDOW;
                    if (w < p.sync_stride) {
                        atomicAdd(sync + w, 1);
                        if (w > 0 && lock_ok) {
                            // bounded: a group that is not co-resident (fewer free SMs than CTAs) times out ONCE and then
                            // free-runs -- lockstep only ever saves bandwidth
                            const long long t0c = clock64();
                            while (*reinterpret_cast<volatile int*>(sync + w - 1) < 8) {
                                if (clock64() - t0c > 400000LL) { lock_ok = false; break; }
                            }
                        }
                    }
                }
                const int bblk = t / p.tiles_per_blk, tp = t - bblk * p.tiles_per_blk;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_a_full(s), 2 * S_TILE_BYTES);
                    tma_load_3d(a_base + (2 * s) * S_TILE_BYTES, &tm_a_hi, kb * SBK, tp * 2, bblk * 64, bar_a_full(s));
                    tma_load_3d(a_base + (2 * s + 1) * S_TILE_BYTES, &tm_a_lo, kb * SBK, tp * 2, bblk * 64, bar_a_full(s));
                    if (++s == X2_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(128, SBM);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = tile0; t < p.num_tiles; t += tile_step, ++it) {
                const int ab = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(bar_acc_empty(ab), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ACC_COL + ab * SBM;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t x_1 = a_base + (2 * s) * S_TILE_BYTES, x_2 = x_1 + S_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        const uint32_t w_hi = tmem_base + (kb * 4 + k) * 8, w_p = w_hi + wcols;
                        const uint64_t dx_1 = smem_desc_sw128(x_1 + k * 32), dx_2 = smem_desc_sw128(x_2 + k * 32);
                        if (p.terms == 3) {
                            mma_f16_ts(d_tmem, w_hi, dx_2, idesc, (kb | k) != 0);              // x_lo.W_hi: small terms first
                            mma_f16_ts(d_tmem, w_p, dx_1, idesc, 1);                           // x_hi.W_lo
                        } else {
                            mma_f16_ts(d_tmem, w_p, dx_2, idesc, (kb | k) != 0);               // x2.W'
                        }
                        mma_f16_ts(d_tmem, w_hi, dx_1, idesc, 1);                              // x_hi.W_hi / x1.W_hi
                    }
                    mma_commit(bar_a_empty(s));
                    if (++s == X2_STAGES) { s = 0; ph ^= 1; }
                }
                mma_commit(bar_acc_full(ab));
            }
        }
    } else {
        // ===================== weights -> TMEM, then epilogue =====================
        // 8 warps: warp w owns TMEM lane quarter w & 3 (hardware rule) and column half (w - 2) >> 2 of every accumulator
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int unit = q * 32 + lane;
        const int out = blk * 128 + unit;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        {
            const uint4* whi = reinterpret_cast<const uint4*>(p.w_hi + (size_t)out * p.ldw);
            const uint4* wlo = reinterpret_cast<const uint4*>(p.w_lo + (size_t)out * p.ldw);
            for (int part = half; part < p.kb * 2; part += 2) {                 // 32 fp16 = 16 packed columns
                uint32_t ra[16], rb[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 vh = __ldg(whi + part * 4 + i), vl = __ldg(wlo + part * 4 + i);
                    const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ra[4 * i + e] = hw[e];
                        if (p.terms == 3) {
                            rb[4 * i + e] = lw[e];
                        } else {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                            const __half2 wp = __floats2half2_rn(fmaf(fl.x, 1.f / X2_S, fh.x), fmaf(fl.y, 1.f / X2_S, fh.y));
                            rb[4 * i + e] = *reinterpret_cast<const uint32_t*>(&wp);
                        }
                    }
                }
                tmem_st16(lane_addr + part * 16, ra);
                tmem_st16(lane_addr + wcols + part * 16, rb);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_w);
        const float bias = p.bias ? __ldg(p.bias + out) : 0.f;
        const int dir = blk >> 2, gate = blk & 3;
        int it = 0;
        for (int t = tile0; t < p.num_tiles; t += tile_step, ++it) {
            const int ab = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(bar_acc_full(ab), acc_ph);
            tc_fence_after();
            const int bblk = t / p.tiles_per_blk, tp = t - bblk * p.tiles_per_blk;
            const int t0 = tp * 2;
            float* base = p.xg + ((((int64_t)bblk * 2 + dir) * p.T + t0) * 4 + gate) * 8192 + unit * 4;
            const bool t1_ok = t0 + 1 < p.T;
            // both 32-column chunks of this warp's half are fetched before the first store is issued
            float v0[32], v1[32];
            tmem_ld32(lane_addr + ACC_COL + ab * SBM + half * 64, v0);
            tmem_ld32(lane_addr + ACC_COL + ab * SBM + half * 64 + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_acc_empty(ab));
            xg_store_chunk(base, half * 2, v0, bias, t1_ok);
            xg_store_chunk(base, half * 2 + 1, v1, bias, t1_ok);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------- CTA-pair input projection (cta_group::2)
// Every activation tile is consumed by the 8 feature-block CTAs of a group, so 8x the activation bytes cross the L2
// fabric into shared memory (26.8 GB per K = 256 launch next to 13.4 GB of xg writes), and the chip sits at its power cap
// during these kernels: the single-CTA 3-term and 2-MMA variants measured the same ~4.0 ms (profiles/r01_gemm.md).
// Here two feature blocks form a CTA pair on one TPC: each CTA keeps its own 128 features' weights in its own TMEM and
// loads HALF of the activation tile (32 sequences x 2 steps); one tcgen05.mma.cta_group::2 (M = 256, N = 128) issued
// by the even CTA multiplies both CTAs' weights with both halves.  Activation bytes through L2 and per-CTA shared-memory
// fill are halved; accumulators, epilogue and xg layout are those of gemm_ts_kernel<3>.
// terms = 3: planes are (hi, lo), weights (W_hi, W_lo);  terms = 2: planes are the scaled (x1, x2) split, weights
// (W_hi, W') -- see gemm_xg2_kernel.
constexpr int XP_THREADS = 320;                       // TMA, MMA, 8 epilogue warps
constexpr int XP_STAGES = 10;
constexpr int XP_HALF_BYTES = 64 * SBK * 2;           // 8 KB: 32 sequences x 2 steps x 64 k

struct GemmXpParams {
    const __half* w_hi;      // [1024][ldw]
    const __half* w_lo;
    const float* bias;       // [1024]
    float* xg;
    int T, tiles_per_blk, num_tiles, kb, ldw, terms;
    int k16;                 // k-steps of 16 that hold real inputs (K rounded up to 16): the zero padding up to kb * 64 is skipped
    int* sync;               // [groups][sync_stride] window arrival counters (zeroed per launch) or null
    int sync_stride;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XP_THREADS, 1)
gemm_xg_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, GemmXpParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;                                        // [stages][plane a, plane b] half tiles of 64 x 64
    const uint32_t bar_base = a_base + XP_STAGES * 2 * XP_HALF_BYTES;
    auto bar_a_full = [&](int s) { return bar_base + 8 * s; };                // even CTA's: both CTAs' TMA bytes (tx) + 1
    auto bar_a_empty = [&](int s) { return bar_base + 128 + 8 * s; };         // both CTAs': MMAs retired (multicast commit)
    auto bar_acc_full = [&](int b) { return bar_base + 256 + 8 * b; };        // both CTAs' (multicast commit)
    auto bar_acc_empty = [&](int b) { return bar_base + 272 + 8 * b; };       // even CTA's: 512 epilogue threads of the pair
    const uint32_t bar_w = bar_base + 288;                                    // even CTA's: 512 weight loaders
    const uint32_t tmem_slot = bar_base + 296;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                                  // == blockIdx.x & 1
    const int blk = blockIdx.x & 7;                                           // feature block = dir * 4 + gate
    const int tile0 = blockIdx.x >> 3;
    const int tile_step = gridDim.x >> 3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < XP_STAGES; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), 512); }
        mbar_init(bar_w, 512);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                                                       // barriers of both CTAs exist before any remote use
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int wcols = p.kb * 32;                                              // TMEM columns per weight plane (K / 2)

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; bytes are credited to the even CTA's barrier) =====================
        if (elect_one()) {
            tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
            int s = 0;
            uint32_t ph = 0;
            // lockstep of the 8 CTAs of a group, as in gemm_xg2_kernel: bounded waits, bandwidth only
            int* const sync = p.sync ? p.sync + (size_t)tile0 * p.sync_stride : nullptr;
            int ti = 0;
            bool lock_ok = true;
            for (int t = tile0; t < p.num_tiles; t += tile_step, ++ti) {
                if (sync && (ti % X2_WINDOW) == 0) {
                    const int w = ti / X2_WINDOW;
                    if (w < p.sync_stride) {
                        atomicAdd(sync + w, 1);
                        if (w > 0 && lock_ok) {
                            // bounded: a group that is not co-resident (fewer free SMs than CTAs) times out ONCE and then
                            // free-runs -- lockstep only ever saves bandwidth
                            const long long t0c = clock64();
                            while (*reinterpret_cast<volatile int*>(sync + w - 1) < 8) {
                                if (clock64() - t0c > 400000LL) { lock_ok = false; break; }
                            }
                        }
                    }
                }
                const int bblk = t / p.tiles_per_blk, tp = t - bblk * p.tiles_per_blk;
                const int seq0 = bblk * 64 + (int)rank * 32;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    if (rank == 0) mbar_expect_tx(bar_a_full(s), 4 * XP_HALF_BYTES);
                    tma_load_3d_pair(a_base + (2 * s) * XP_HALF_BYTES, &tm_a, kb * SBK, tp * 2, seq0, bar_a_full(s));
                    tma_load_3d_pair(a_base + (2 * s + 1) * XP_HALF_BYTES, &tm_b, kb * SBK, tp * 2, seq0, bar_a_full(s));
                    if (++s == XP_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (even CTA only, for the pair) =====================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = idesc_f16(256, SBM);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = tile0; t < p.num_tiles; t += tile_step, ++it) {
                const int ab = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(bar_acc_empty(ab), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ACC_COL + ab * SBM;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t x_a = a_base + (2 * s) * XP_HALF_BYTES, x_b = x_a + XP_HALF_BYTES;
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        if (kb * 4 + k >= p.k16) break;                                        // zero padding of K
                        const uint32_t w_a = tmem_base + (kb * 4 + k) * 8, w_b = w_a + wcols;
                        const uint64_t dx_a = smem_desc_sw128(x_a + k * 32), dx_b = smem_desc_sw128(x_b + k * 32);
                        if (p.terms == 3) {
                            mma_f16_ts_pair(d_tmem, w_a, dx_b, idesc, (kb | k) != 0);          // x_lo.W_hi: small terms first
                            mma_f16_ts_pair(d_tmem, w_b, dx_a, idesc, 1);                      // x_hi.W_lo
                        } else {
                            mma_f16_ts_pair(d_tmem, w_b, dx_b, idesc, (kb | k) != 0);          // x2.W'
                        }
                        mma_f16_ts_pair(d_tmem, w_a, dx_a, idesc, 1);                          // x_hi.W_hi / x1.W_hi
                    }
                    mma_commit_pair(bar_a_empty(s));
                    if (++s == XP_STAGES) { s = 0; ph ^= 1; }
                }
                mma_commit_pair(bar_acc_full(ab));
            }
        }
    } else {
        // ===================== weights -> own TMEM, then epilogue =====================
        // 8 warps: warp w owns TMEM lane quarter w & 3 (hardware rule) and column half (w - 2) >> 2 of every accumulator
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int unit = q * 32 + lane;
        const int out = blk * 128 + unit;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        {
            const uint4* whi = reinterpret_cast<const uint4*>(p.w_hi + (size_t)out * p.ldw);
            const uint4* wlo = reinterpret_cast<const uint4*>(p.w_lo + (size_t)out * p.ldw);
            for (int part = half; part < p.kb * 2; part += 2) {                 // 32 fp16 = 16 packed columns
                uint32_t ra[16], rb[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 vh = __ldg(whi + part * 4 + i), vl = __ldg(wlo + part * 4 + i);
                    const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ra[4 * i + e] = hw[e];
                        if (p.terms == 3) {
                            rb[4 * i + e] = lw[e];
                        } else {
                            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                            const __half2 wp = __floats2half2_rn(fmaf(fl.x, 1.f / X2_S, fh.x), fmaf(fl.y, 1.f / X2_S, fh.y));
                            rb[4 * i + e] = *reinterpret_cast<const uint32_t*>(&wp);
                        }
                    }
                }
                tmem_st16(lane_addr + part * 16, ra);
                tmem_st16(lane_addr + wcols + part * 16, rb);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_cluster(mapa_shared(bar_w, 0));
        const float bias = p.bias ? __ldg(p.bias + out) : 0.f;
        const int dir = blk >> 2, gate = blk & 3;
        const uint32_t acc_empty0 = mapa_shared(bar_acc_empty(0), 0), acc_empty1 = mapa_shared(bar_acc_empty(1), 0);
        int it = 0;
        for (int t = tile0; t < p.num_tiles; t += tile_step, ++it) {
            const int ab = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(bar_acc_full(ab), acc_ph);
            tc_fence_after();
            const int bblk = t / p.tiles_per_blk, tp = t - bblk * p.tiles_per_blk;
            const int t0 = tp * 2;
            float* base = p.xg + ((((int64_t)bblk * 2 + dir) * p.T + t0) * 4 + gate) * 8192 + unit * 4;
            const bool t1_ok = t0 + 1 < p.T;
            // both 32-column chunks of this warp's half are fetched before the first store is issued
            // (accumulator columns [0, 64) come from the even CTA's half tile)
            float v0[32], v1[32];
            tmem_ld32(lane_addr + ACC_COL + ab * SBM + half * 64, v0);
            tmem_ld32(lane_addr + ACC_COL + ab * SBM + half * 64 + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive_cluster_relaxed(ab ? acc_empty1 : acc_empty0);
            xg_store_chunk(base, half * 2, v0, bias, t1_ok);
            xg_store_chunk(base, half * 2 + 1, v1, bias, t1_ok);
        }
    }
    tc_fence_before();
    cluster_sync_all();                                                       // the peer's TMEM / barriers stay alive until both are done
    if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

// Loads the two weight planes of output row `wrow_*` into TMEM (lane = this thread's row): plane 0 = W_hi, plane 1 = W_lo
// (terms = 3) or W' = fp16(W_hi + W_lo / s) (terms = 2, the scaled-plane product described at gemm_xg2_kernel).
__device__ __forceinline__ void load_weight_planes(const __half* w_hi_row, const __half* w_lo_row, int parts, uint32_t lane_addr, int wcols, int terms) {
    const uint4* whi = reinterpret_cast<const uint4*>(w_hi_row);
    const uint4* wlo = reinterpret_cast<const uint4*>(w_lo_row);
    for (int part = 0; part < parts; ++part) {                                  // 32 fp16 = 16 packed columns
        uint32_t ra[16], rb[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 vh = __ldg(whi + part * 4 + i), vl = __ldg(wlo + part * 4 + i);
            const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ra[4 * i + e] = hw[e];
                if (terms == 3) {
                    rb[4 * i + e] = lw[e];
                } else {
                    const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                    const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                    const __half2 wp = __floats2half2_rn(fmaf(fl.x, 1.f / kPlaneScale, fh.x), fmaf(fl.y, 1.f / kPlaneScale, fh.y));
                    rb[4 * i + e] = *reinterpret_cast<const uint32_t*>(&wp);
                }
            }
        }
        tmem_st16(lane_addr + part * 16, ra);
        tmem_st16(lane_addr + wcols + part * 16, rb);
    }
}

// ---------------------------------------------------------------- sinc convolution + |x| + MaxPool1d(3) + InstanceNorm sums
// The stride-10 sinc layer (sincnet.py:50-61, 95-99) as ONE overlapping-row GEMM whose epilogue pools.  Rows t = 4 m + r of
// residue class r come from their own tensor map (80-byte row stride on the waveform copy shifted by 10 r mod 8 samples,
// see sincnet.cu); a tile takes a 32-row box from each of the four classes into smem rows [32 r, 32 r + 32), so accumulator
// column 32 r + m is convolution row t0 + 4 m + r: the MMA does not care about row order, and the epilogue thread of a
// feature holds all 128 columns in registers, where the three rows of every pooling group meet.  A tile advances by 120
// rows (30 per class: 40 pooling groups; box rows 30, 31 are recomputed by the next tile).  Written: pooled (B, P, C) fp32
// and the per-(item, channel) sum / sum of squares for InstanceNorm -- the (B, L, C) convolution output (4 MB per 8 s row)
// never exists.
constexpr int SP_THREADS = 192;
constexpr int SP_STAGES = 6;
constexpr int SP_TILE_ROWS = 120;                     // convolution rows per tile
struct alignas(64) SincMaps {
    CUtensorMap hi[4], lo[4];
};
struct SincPoolParams {
    const __half* w_hi;      // [128][ldw], rows >= n_valid zero
    const __half* w_lo;
    float* pooled;           // [B][P][ldc]
    double* stats;           // [B][n_valid][2]
    int64_t P;
    int ldc, n_valid, ldw, kb, tiles_per_batch, num_tiles, terms;
};

__global__ void __launch_bounds__(SP_THREADS, 1)
sinc_pool_gemm_kernel(const __grid_constant__ SincMaps maps, SincPoolParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;                                        // [stages][hi, lo] tiles of 128 x 64
    const uint32_t bar_base = a_base + SP_STAGES * 2 * S_TILE_BYTES;
    auto bar_a_full = [&](int s) { return bar_base + 8 * s; };
    auto bar_a_empty = [&](int s) { return bar_base + 64 + 8 * s; };
    auto bar_acc_full = [&](int b) { return bar_base + 128 + 8 * b; };
    auto bar_acc_empty = [&](int b) { return bar_base + 144 + 8 * b; };
    const uint32_t bar_w = bar_base + 160;
    const uint32_t tmem_slot = bar_base + 168;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SP_STAGES; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), 128); }
        mbar_init(bar_w, 128);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int wcols = p.kb * 32;                                              // TMEM columns per weight plane

    if (warp == 0) {
        // ===================== TMA producer: four 32-row class boxes per plane and k-block =====================
        if (elect_one()) {
            for (int r = 0; r < 4; ++r) { tma_prefetch_desc(&maps.hi[r]); tma_prefetch_desc(&maps.lo[r]); }
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int bi = t / p.tiles_per_batch, rt = t - bi * p.tiles_per_batch;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_a_full(s), 2 * S_TILE_BYTES);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        tma_load_3d(a_base + (2 * s) * S_TILE_BYTES + r * 4096, &maps.hi[r], kb * SBK, rt * 30, bi, bar_a_full(s));
                        tma_load_3d(a_base + (2 * s + 1) * S_TILE_BYTES + r * 4096, &maps.lo[r], kb * SBK, rt * 30, bi, bar_a_full(s));
                    }
                    if (++s == SP_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (three split-precision products, as gemm_ts_kernel) =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(128, SBM);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                const int ab = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(bar_acc_empty(ab), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ACC_COL + ab * SBM;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t x_hi = a_base + (2 * s) * S_TILE_BYTES, x_lo = x_hi + S_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        const uint32_t w_hi = tmem_base + (kb * 4 + k) * 8, w_lo = w_hi + wcols;
                        const uint64_t dx_hi = smem_desc_sw128(x_hi + k * 32), dx_lo = smem_desc_sw128(x_lo + k * 32);
                        if (p.terms == 3) {
                            mma_f16_ts(d_tmem, w_hi, dx_lo, idesc, (kb | k) != 0);        // small terms first
                            mma_f16_ts(d_tmem, w_lo, dx_hi, idesc, 1);
                        } else {
                            mma_f16_ts(d_tmem, w_lo, dx_lo, idesc, (kb | k) != 0);        // x2 . W' (scaled planes)
                        }
                        mma_f16_ts(d_tmem, w_hi, dx_hi, idesc, 1);
                    }
                    mma_commit(bar_a_empty(s));
                    if (++s == SP_STAGES) { s = 0; ph ^= 1; }
                }
                mma_commit(bar_acc_full(ab));
            }
        }
    } else {
        // ===================== weights -> TMEM, then the pooling epilogue (warps 2..5 -> lane quarters 2,3,0,1) =====================
        const int q = warp & 3;
        const int out = q * 32 + lane;                                          // output channel == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        load_weight_planes(p.w_hi + (size_t)out * p.ldw, p.w_lo + (size_t)out * p.ldw, p.kb * 2, lane_addr, wcols, p.terms);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_w);
        const bool out_ok = out < p.n_valid;
        int it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int ab = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            const int bi = t / p.tiles_per_batch, rt = t - bi * p.tiles_per_batch;
            mbar_wait(bar_acc_full(ab), acc_ph);
            tc_fence_after();
            float v[4][32];                                                     // v[r][m] = row 4 m + r of the tile
#pragma unroll
            for (int r = 0; r < 4; ++r) tmem_ld32(lane_addr + ACC_COL + ab * SBM + r * 32, v[r]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_acc_empty(ab));
            if (!out_ok) continue;
            const int64_t p0 = (int64_t)rt * (SP_TILE_ROWS / 3);
            float* dst = p.pooled + ((int64_t)bi * p.P + p0) * p.ldc + out;
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int g = 0; g < SP_TILE_ROWS / 3; ++g) {
                // rows 3 g, 3 g + 1, 3 g + 2 of the tile: compile-time (class, slot) after unrolling
                const float a0 = fabsf(v[(3 * g) & 3][(3 * g) >> 2]);
                const float a1 = fabsf(v[(3 * g + 1) & 3][(3 * g + 1) >> 2]);
                const float a2 = fabsf(v[(3 * g + 2) & 3][(3 * g + 2) >> 2]);
                const float m = fmaxf(fmaxf(a0, a1), a2);
                if (p0 + g < p.P) {
                    dst[(int64_t)g * p.ldc] = m;
                    sum += m;
                    sq = fmaf(m, m, sq);
                }
            }
            atomicAdd(p.stats + ((int64_t)bi * p.n_valid + out) * 2, exact_partial<32>((double)sum));   // exact -> order-independent (common.cuh)
            atomicAdd(p.stats + ((int64_t)bi * p.n_valid + out) * 2 + 1, exact_partial<32>((double)sq));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------- Conv1d(k = 5) + bias + MaxPool1d(3) + InstanceNorm sums
// The 5-tap convolutions of SincNet (sincnet.py:62-71, 100-101) over channel-last fp16 planes: row t of the GEMM is the
// 5 * Cin contiguous values starting at frame t (overlapping rows, one 3-D tensor map), K = 400 / 320.  Both weight
// planes of the WHOLE K stay in tensor memory (2 x Kp / 2 <= 448 columns), which leaves room for 64-column accumulators
// only -- one (K = 448) or two (K = 320) of them -- but removes the second, read-add-store pass of a K-split launch pair.  A
// tile is 64 rows, advances by 63 (21 pooling groups; the last row is recomputed by the next tile), and the epilogue thread
// of a channel pools its 64 accumulator columns in registers: written are pooled (B, P, C) fp32 and the InstanceNorm
// sums; the convolution output itself never exists.
constexpr int CP_THREADS = 192;
constexpr int CP_STAGES = 8;
constexpr int CP_ROWS = 64;                           // MMA N
constexpr int CP_TILE_BYTES = CP_ROWS * SBK * 2;      // 8 KB
constexpr int CP_ADV = 63;                            // rows a tile advances by (21 pooling groups)
struct ConvPoolParams {
    const __half* w_hi;      // [128][ldw], rows >= n_valid and columns >= K zero
    const __half* w_lo;
    const float* bias;       // [n_valid]
    float* pooled;           // [B][P][ldc]
    double* stats;           // [B][n_valid][2]
    int64_t P;
    int ldc, n_valid, ldw, kb, nacc, tiles_per_batch, num_tiles, terms;
};

__global__ void __launch_bounds__(CP_THREADS, 1)
conv_pool_gemm_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, ConvPoolParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;                                        // [stages][hi, lo] tiles of 64 x 64
    const uint32_t bar_base = a_base + CP_STAGES * 2 * CP_TILE_BYTES;
    auto bar_a_full = [&](int s) { return bar_base + 8 * s; };
    auto bar_a_empty = [&](int s) { return bar_base + 64 + 8 * s; };
    auto bar_acc_full = [&](int b) { return bar_base + 128 + 8 * b; };
    auto bar_acc_empty = [&](int b) { return bar_base + 144 + 8 * b; };
    const uint32_t bar_w = bar_base + 160;
    const uint32_t tmem_slot = bar_base + 168;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CP_STAGES; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), 128); }
        mbar_init(bar_w, 128);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int wcols = p.kb * 32;                                              // TMEM columns per weight plane
    const uint32_t acc_col = 2 * wcols;                                       // accumulators follow the two weight planes

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            tma_prefetch_desc(&tm_hi); tma_prefetch_desc(&tm_lo);
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int bi = t / p.tiles_per_batch, rt = t - bi * p.tiles_per_batch;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_a_full(s), 2 * CP_TILE_BYTES);
                    tma_load_3d(a_base + (2 * s) * CP_TILE_BYTES, &tm_hi, kb * SBK, rt * CP_ADV, bi, bar_a_full(s));
                    tma_load_3d(a_base + (2 * s + 1) * CP_TILE_BYTES, &tm_lo, kb * SBK, rt * CP_ADV, bi, bar_a_full(s));
                    if (++s == CP_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(128, CP_ROWS);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                const int ab = it % p.nacc;
                const uint32_t acc_ph = (it / p.nacc) & 1;
                mbar_wait(bar_acc_empty(ab), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc_col + ab * CP_ROWS;
                for (int kb = 0; kb < p.kb; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t x_hi = a_base + (2 * s) * CP_TILE_BYTES, x_lo = x_hi + CP_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        const uint32_t w_hi = tmem_base + (kb * 4 + k) * 8, w_lo = w_hi + wcols;
                        const uint64_t dx_hi = smem_desc_sw128(x_hi + k * 32), dx_lo = smem_desc_sw128(x_lo + k * 32);
                        if (p.terms == 3) {
                            mma_f16_ts(d_tmem, w_hi, dx_lo, idesc, (kb | k) != 0);        // small terms first
                            mma_f16_ts(d_tmem, w_lo, dx_hi, idesc, 1);
                        } else {
                            mma_f16_ts(d_tmem, w_lo, dx_lo, idesc, (kb | k) != 0);        // x2 . W' (scaled planes)
                        }
                        mma_f16_ts(d_tmem, w_hi, dx_hi, idesc, 1);
                    }
                    mma_commit(bar_a_empty(s));
                    if (++s == CP_STAGES) { s = 0; ph ^= 1; }
                }
                mma_commit(bar_acc_full(ab));
            }
        }
    } else {
        // ===================== weights -> TMEM, then the pooling epilogue =====================
        const int q = warp & 3;
        const int out = q * 32 + lane;                                          // output channel == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        load_weight_planes(p.w_hi + (size_t)out * p.ldw, p.w_lo + (size_t)out * p.ldw, p.kb * 2, lane_addr, wcols, p.terms);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_w);
        const bool out_ok = out < p.n_valid;
        const float bias = (p.bias && out_ok) ? __ldg(p.bias + out) : 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int ab = it % p.nacc;
            const uint32_t acc_ph = (it / p.nacc) & 1;
            const int bi = t / p.tiles_per_batch, rt = t - bi * p.tiles_per_batch;
            mbar_wait(bar_acc_full(ab), acc_ph);
            tc_fence_after();
            float v[2][32];
            tmem_ld32(lane_addr + acc_col + ab * CP_ROWS, v[0]);
            tmem_ld32(lane_addr + acc_col + ab * CP_ROWS + 32, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_acc_empty(ab));
            if (!out_ok) continue;
            const int64_t p0 = (int64_t)rt * (CP_ADV / 3);
            float* dst = p.pooled + ((int64_t)bi * p.P + p0) * p.ldc + out;
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int g = 0; g < CP_ADV / 3; ++g) {
                // max commutes with the bias add
                const float m = fmaxf(fmaxf(v[(3 * g) >> 5][(3 * g) & 31], v[(3 * g + 1) >> 5][(3 * g + 1) & 31]),
                                      v[(3 * g + 2) >> 5][(3 * g + 2) & 31]) + bias;
                if (p0 + g < p.P) {
                    dst[(int64_t)g * p.ldc] = m;
                    sum += m;
                    sq = fmaf(m, m, sq);
                }
            }
            atomicAdd(p.stats + ((int64_t)bi * p.n_valid + out) * 2, exact_partial<32>((double)sum));   // exact -> order-independent (common.cuh)
            atomicAdd(p.stats + ((int64_t)bi * p.n_valid + out) * 2 + 1, exact_partial<32>((double)sq));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------- fused head: Linear(256,128) + LeakyReLU -> Linear(128,128) + LeakyReLU
//                                                                   -> Linear(128,1) + Sigmoid   (PyanNet2.py:183-187)
// One kernel instead of two: both weight matrices (hi + lo planes: 256 + 128 TMEM columns) stay in tensor memory next to two
// 64-column accumulators.  Per tile of 64 frames: acc1 = W1 . y^T (y planes through a TMA ring); the epilogue warps turn acc1
// into z1 = leaky_relu(acc1 + b1) as fp16 (hi, lo) planes written straight into a swizzled K-major operand tile in shared
// memory (the way the recurrence writes h); acc2 = W2 . z1^T; the epilogue applies bias, LeakyReLU, the classifier dot
// product (register transpose-reduction + 4-warp exchange) and the sigmoid.  The z1 planes (1 KB per frame written and read
// back) never touch HBM.  Same products in the same order as gemm_ts_kernel<1> + <4>: bit-identical probabilities.
constexpr int HF_THREADS = 320;                      // TMA, MMA, 4 phase-1 warps, 4 phase-2 warps
constexpr int HF_STAGES = 8;
constexpr int HF_ROWS = 64;
constexpr int HF_TILE_BYTES = HF_ROWS * SBK * 2;      // 8 KB
constexpr int HF_W2_COL = 256, HF_ACC1_COL = 384, HF_ACC2_COL = 448;
struct HeadFusedParams {
    const __half* w1_hi;     // [128][256]
    const __half* w1_lo;
    const __half* w2_hi;     // [128][128]
    const __half* w2_lo;
    const float* b1;         // [128]
    const float* b2;         // [128]
    const float* wc;         // [128]
    const float* bc;         // [1]
    float* prob;             // [M]
    int64_t M;
    int num_tiles;
};

__global__ void __launch_bounds__(HF_THREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tm_y_hi, const __grid_constant__ CUtensorMap tm_y_lo, HeadFusedParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* const smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t a_base = smem_base;                                        // [stages][hi, lo] tiles of 64 x 64
    const uint32_t z_off = HF_STAGES * 2 * HF_TILE_BYTES;                     // z1 operand: [plane 2][kb 2] tiles of 64 x 64
    const uint32_t z_base = smem_base + z_off;
    const uint32_t bar_base = z_base + 4 * HF_TILE_BYTES;
    auto bar_a_full = [&](int s) { return bar_base + 8 * s; };
    auto bar_a_empty = [&](int s) { return bar_base + 64 + 8 * s; };
    const uint32_t bar_acc1_full = bar_base + 128, bar_acc1_empty = bar_base + 136, bar_z_full = bar_base + 144,
                   bar_z_empty = bar_base + 152, bar_acc2_full = bar_base + 160, bar_acc2_empty = bar_base + 168,
                   bar_w = bar_base + 176, tmem_slot = bar_base + 184;
    const uint32_t part_smem = bar_base + 192;                                // [2][4 warps][32 rows] fp32 partial dots
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < HF_STAGES; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_empty(s), 1); }
        mbar_init(bar_acc1_full, 1); mbar_init(bar_acc1_empty, 128); mbar_init(bar_z_full, 128); mbar_init(bar_z_empty, 1);
        mbar_init(bar_acc2_full, 1); mbar_init(bar_acc2_empty, 128); mbar_init(bar_w, 128);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer: y planes, 4 k-blocks per tile =====================
        if (elect_one()) {
            tma_prefetch_desc(&tm_y_hi); tma_prefetch_desc(&tm_y_lo);
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_a_full(s), 2 * HF_TILE_BYTES);
                    tma_load_2d(a_base + (2 * s) * HF_TILE_BYTES, &tm_y_hi, kb * SBK, t * HF_ROWS, bar_a_full(s));
                    tma_load_2d(a_base + (2 * s + 1) * HF_TILE_BYTES, &tm_y_lo, kb * SBK, t * HF_ROWS, bar_a_full(s));
                    if (++s == HF_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_f16(128, HF_ROWS);
            mbar_wait(bar_w, 0);
            tc_fence_after();
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            auto first_gemm = [&](int i) {                                      // acc1 = W1 . y^T of this CTA's i-th tile
                mbar_wait(bar_acc1_empty, (i & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(bar_a_full(s), ph);
                    tc_fence_after();
                    const uint32_t x_hi = a_base + (2 * s) * HF_TILE_BYTES, x_lo = x_hi + HF_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        const uint32_t w_hi = tmem_base + (kb * 4 + k) * 8, w_lo = w_hi + 128;
                        const uint64_t dx_hi = smem_desc_sw128(x_hi + k * 32), dx_lo = smem_desc_sw128(x_lo + k * 32);
                        mma_f16_ts(tmem_base + HF_ACC1_COL, w_hi, dx_lo, idesc, (kb | k) != 0);
                        mma_f16_ts(tmem_base + HF_ACC1_COL, w_lo, dx_hi, idesc, 1);
                        mma_f16_ts(tmem_base + HF_ACC1_COL, w_hi, dx_hi, idesc, 1);
                    }
                    mma_commit(bar_a_empty(s));
                    if (++s == HF_STAGES) { s = 0; ph ^= 1; }
                }
                mma_commit(bar_acc1_full);
            };
            if (blockIdx.x < p.num_tiles) first_gemm(0);
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                // the next tile's first GEMM needs only acc1 to be drained (the start of phase 1 of this tile): issued first, it
                // runs while the phase-1 warps are still writing this tile's z1
                if (t + (int)gridDim.x < p.num_tiles) first_gemm(it + 1);
                // acc2 = W2 . z1^T once z1 of this tile is in shared memory
                mbar_wait(bar_z_full, it & 1);
                mbar_wait(bar_acc2_empty, (it & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        const uint32_t w_hi = tmem_base + HF_W2_COL + (kb * 4 + k) * 8, w_lo = w_hi + 64;
                        const uint64_t dz_hi = smem_desc_sw128(z_base + kb * HF_TILE_BYTES + k * 32);
                        const uint64_t dz_lo = smem_desc_sw128(z_base + (2 + kb) * HF_TILE_BYTES + k * 32);
                        mma_f16_ts(tmem_base + HF_ACC2_COL, w_hi, dz_lo, idesc, (kb | k) != 0);
                        mma_f16_ts(tmem_base + HF_ACC2_COL, w_lo, dz_hi, idesc, 1);
                        mma_f16_ts(tmem_base + HF_ACC2_COL, w_hi, dz_hi, idesc, 1);
                    }
                }
                mma_commit(bar_z_empty);
                mma_commit(bar_acc2_full);
            }
        }
    } else {
        // ===================== two epilogue warpgroups: warps 2..5 = phase 1 (weights -> TMEM, then acc1 -> z1 operand tile),
        //                       warps 6..9 = phase 2 (acc2 -> classifier -> sigmoid); consecutive tiles overlap =====================
        const int q = warp & 3;
        const int out = q * 32 + lane;                                          // feature == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (warp < 6) {
            for (int w = 0; w < 2; ++w) {
                const uint4* w1 = reinterpret_cast<const uint4*>((w == 0 ? p.w1_hi : p.w1_lo) + (size_t)out * 256);
                for (int part = 0; part < 8; ++part) {
                    uint32_t r[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { uint4 v = __ldg(w1 + part * 4 + i); r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w; }
                    tmem_st16(lane_addr + w * 128 + part * 16, r);
                }
                const uint4* w2 = reinterpret_cast<const uint4*>((w == 0 ? p.w2_hi : p.w2_lo) + (size_t)out * 128);
                for (int part = 0; part < 4; ++part) {
                    uint32_t r[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { uint4 v = __ldg(w2 + part * 4 + i); r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w; }
                    tmem_st16(lane_addr + HF_W2_COL + w * 64 + part * 16, r);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_w);
            const float b1 = __ldg(p.b1 + out);
            // z1 element (row j, k = out): tile kb = out / 64, 16-byte chunk (out % 64) / 8 XOR (j % 8), byte (out % 8) * 2
            unsigned char* const z_hi0 = smem_gen + z_off + (out >> 6) * HF_TILE_BYTES + (out & 7) * 2;
            const uint32_t zc = (out & 63) >> 3;
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                mbar_wait(bar_acc1_full, it & 1);
                tc_fence_after();
                float v[2][32];
                tmem_ld32(lane_addr + HF_ACC1_COL, v[0]);
                tmem_ld32(lane_addr + HF_ACC1_COL + 32, v[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(bar_acc1_empty);
                mbar_wait(bar_z_empty, (it & 1) ^ 1);                           // the previous tile's second GEMM has read z1
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = c * 32 + j;
                        float x = v[c][j] + b1;
                        x = x > 0.f ? x : 0.01f * x;
                        __half h, l;
                        split_f16(x, h, l);
                        unsigned char* dst = z_hi0 + n * 128 + ((zc ^ (n & 7)) << 4);
                        *reinterpret_cast<__half*>(dst) = h;
                        *reinterpret_cast<__half*>(dst + 2 * HF_TILE_BYTES) = l;
                    }
                }
                fence_proxy_async();
                mbar_arrive(bar_z_full);
            }
        } else {
            const float b2 = __ldg(p.b2 + out), wc = __ldg(p.wc + out);
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
                const int64_t row0 = (int64_t)t * HF_ROWS;
                const int nrows = (int)min((int64_t)HF_ROWS, p.M - row0);
                mbar_wait(bar_acc2_full, it & 1);
                tc_fence_after();
                float v[2][32];
                tmem_ld32(lane_addr + HF_ACC2_COL, v[0]);
                tmem_ld32(lane_addr + HF_ACC2_COL + 32, v[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(bar_acc2_empty);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[c][j] + b2;
                        v[c][j] = wc * (x > 0.f ? x : 0.01f * x);
                    }
                    // transpose-reduction: after the step with offset o a lane keeps the rows whose bit o equals its own
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float keep = up ? v[c][i + o] : v[c][i];
                            const float send = up ? v[c][i] : v[c][i + o];
                            v[c][i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    const int pb = (it * 2 + c) & 1;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(part_smem + 4 * ((pb * 4 + q) * 32 + lane)), "f"(v[c][0]) : "memory");
                    named_bar_sync(1, 128);
                    if (q == 0) {
                        float sacc = __ldg(p.bc);
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            float pv;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pv) : "r"(part_smem + 4 * ((pb * 4 + w4) * 32 + lane)));
                            sacc += pv;
                        }
                        if (c * 32 + lane < nrows) p.prob[row0 + c * 32 + lane] = 1.f / (1.f + __expf(-sacc));
                    }
                }
            }
        }
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
// rows of D values -> planes with row pitch D8 (D rounded up to 8: 16-byte rows for TMA), padding zeroed
__global__ void __launch_bounds__(256) split_planes_pad_kernel(const float* __restrict__ x, int64_t rows, int D, int D8,
                                                               __half* __restrict__ hi, __half* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * D8) return;
    const int64_t r = i / D8;
    const int k = (int)(i - r * D8);
    __half h = __float2half_rn(0.f), l = h;
    if (k < D) split_f16(x[r * D + k], h, l);
    hi[i] = h;
    lo[i] = l;
}
int split_planes_pad_launch(const float* x, int64_t rows, int D, int D8, __half* hi, __half* lo, cudaStream_t st) {
    if (rows <= 0) return B200VAD_OK;
    split_planes_pad_kernel<<<(unsigned)((rows * D8 + 255) / 256), 256, 0, st>>>(x, rows, D, D8, hi, lo);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
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

int make_tmap_4d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, const uint64_t dims[4], const uint64_t pitch_bytes[3],
                 const uint32_t box[4], CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return B200VAD_ESTATE;
    }
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t st[3] = {pitch_bytes[0], pitch_bytes[1], pitch_bytes[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, dtype, 4, const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(4d) failed: %d", (int)r);
        return B200VAD_ECUDA;
    }
    return B200VAD_OK;
}

// ---------------------------------------------------------------- launchers
static int gemm_ts_run(int mode, const CUtensorMap& tm_a_hi, const CUtensorMap& tm_a_lo, GemmTsParams& p, int N, int num_sms,
                       cudaStream_t st) {
    p.n_blocks = N / 128;
    p.stages = 6;
    const int smem = 1024 + p.stages * 2 * S_TILE_BYTES + 192 + 1024 + 64;
    int grid = (num_sms / p.n_blocks) * p.n_blocks;
    if (grid < p.n_blocks) grid = p.n_blocks;
    const int64_t max_grid = (int64_t)p.num_m_tiles * p.n_blocks;
    if (grid > max_grid) grid = (int)max_grid;
    typedef void (*KernFn)(CUtensorMap, CUtensorMap, GemmTsParams);
    static const KernFn kerns[5] = {gemm_ts_kernel<0>, gemm_ts_kernel<1>, gemm_ts_kernel<2>, gemm_ts_kernel<3>, gemm_ts_kernel<4>};
    if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(kerns[mode]), smem)) return rc;
    const int prof_kind = (mode == 0 || mode == 3) ? 1 : 2;
    prof_begin(prof_kind, st);
    kerns[mode]<<<grid, GEMM_TS_THREADS, smem, st>>>(tm_a_hi, tm_a_lo, p);
    prof_end(prof_kind, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

static int gemm_ts_check(int K, int Kp, int ldw, int N, int64_t lda, int nw) {
    if (N % 128 != 0 || Kp % 64 != 0 || lda % 8 != 0 || K > Kp || ldw < Kp || ldw % 8 != 0 || nw * Kp > 512) {
        set_error("gemm_ts: unsupported shape N=%d K=%d Kp=%d ldw=%d planes=%d lda=%lld", N, K, Kp, ldw, nw, (long long)lda);
        return B200VAD_EINVAL;
    }
    return B200VAD_OK;
}

// a_hi/a_lo: [M, K] fp16 (row pitch lda elements, lda % 8 == 0);  w_hi/w_lo: [N, ldw] fp16, columns [K, Kp) zero (Kp % 64 == 0);
// N must be a multiple of 128 and planes * Kp <= 512 (weights resident in 256 TMEM columns).  w_lo may be null.
// mode 0: c[M, ldc] fp32 = acc + bias (accumulate: c += acc);  mode 1: leaky_relu -> fp16 planes o_hi / o_lo [M, ldc];
// mode 2: leaky_relu -> c fp32;  mode 4 (N == 128): leaky_relu -> sigmoid(wc . z + bc) -> c = prob [M].
int gemm_ts_launch(const __half* a_hi, const __half* a_lo, int64_t lda, int64_t M, int K, const __half* w_hi, const __half* w_lo,
                   int Kp, int ldw, int N, const float* bias, int mode, int accumulate, float* c, __half* o_hi, __half* o_lo,
                   int64_t ldc, int num_sms, cudaStream_t st, const float* wc, const float* bc) {
    if (M <= 0) return B200VAD_OK;
    const int nw = w_lo ? 2 : 1;
    int rc = gemm_ts_check(K, Kp, ldw, N, lda, nw);
    if (rc) return rc;
    if (!(mode == 0 || mode == 1 || mode == 2 || (mode == 4 && N == 128 && wc && bc))) {
        set_error("gemm_ts: bad mode %d (mode 4 needs N == 128 and the classifier weights)", mode);
        return B200VAD_EINVAL;
    }
    GemmTsParams p = {};
    p.wc = wc; p.bc = bc; p.n_valid = N;
    p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.c = c; p.o_hi = o_hi; p.o_lo = o_lo; p.ldc = ldc; p.M = M;
    p.num_m_tiles = (int)((M + SBM - 1) / SBM);
    p.kb = Kp / SBK; p.nw = nw; p.Kp = Kp; p.ldw = ldw; p.accumulate = accumulate;
    CUtensorMap tm_a_hi, tm_a_lo;
    if ((rc = make_tmap_2d(&tm_a_hi, a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, M, lda * 2, SBK, SBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d(&tm_a_lo, a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, M, lda * 2, SBK, SBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    return gemm_ts_run(mode, tm_a_hi, tm_a_lo, p, N, num_sms, st);
}

// Batched overlapping-row GEMM (the SincNet convolutions as GEMMs over sliding windows of a channel-last signal):
// planes a_hi / a_lo hold, per batch item (batch_stride elements apart), a signal whose row r is the K consecutive
// elements starting at r * row_stride (rows overlap; row_stride * 2 bytes must be a multiple of 16).  C row
// b * out_batch_rows + r * out_row_step + out_row_off, feature f < n_valid  (+)=  sum_k a[b][r][k] * w[f][k] (+ bias),
// optionally |.| on the way out.  Weights: [128][ldw] fp16 hi / lo, rows >= n_valid and columns >= K zero.
int gemm_ts_rows_launch(const __half* a_hi, const __half* a_lo, int64_t row_stride, int64_t batch_stride, int B, int rows_per_batch,
                        int K, const __half* w_hi, const __half* w_lo, int Kp, int ldw, int n_valid, const float* bias,
                        int accumulate, int act_abs, float* c, int64_t ldc, int64_t out_batch_rows, int out_row_step,
                        int out_row_off, int num_sms, cudaStream_t st) {
    if (B <= 0 || rows_per_batch <= 0) return B200VAD_OK;
    int rc = gemm_ts_check(K, Kp, ldw, 128, 8, w_lo ? 2 : 1);
    if (rc) return rc;
    if ((row_stride * 2) % 16 != 0 || (batch_stride * 2) % 16 != 0 || n_valid < 1 || n_valid > 128 ||
        (reinterpret_cast<uintptr_t>(a_hi) & 15) || (reinterpret_cast<uintptr_t>(a_lo) & 15)) {
        set_error("gemm_ts_rows: row / batch strides and base pointers must be 16-byte aligned, 1 <= n_valid <= 128");
        return B200VAD_EINVAL;
    }
    GemmTsParams p = {};
    p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.c = c; p.ldc = ldc;
    p.rows3d = 1; p.rows_per_batch = rows_per_batch; p.tiles_per_batch = (rows_per_batch + SBM - 1) / SBM;
    p.out_batch_rows = out_batch_rows; p.out_row_step = out_row_step; p.out_row_off = out_row_off;
    p.n_valid = n_valid; p.act_abs = act_abs; p.accumulate = accumulate;
    p.M = (int64_t)B * rows_per_batch;
    const int64_t tiles = (int64_t)B * p.tiles_per_batch;
    if (tiles >= (1LL << 31)) { set_error("gemm_ts_rows: too many tiles"); return B200VAD_EINVAL; }
    p.num_m_tiles = (int)tiles;
    p.kb = Kp / SBK; p.nw = w_lo ? 2 : 1; p.Kp = Kp; p.ldw = ldw;
    CUtensorMap tm_a_hi, tm_a_lo;
    if ((rc = make_tmap_3d(&tm_a_hi, a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, rows_per_batch, B, row_stride * 2, batch_stride * 2,
                           SBK, SBM, 1, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_a_lo, a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, rows_per_batch, B, row_stride * 2, batch_stride * 2,
                           SBK, SBM, 1, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    return gemm_ts_run(0, tm_a_hi, tm_a_lo, p, 128, num_sms, st);
}

// Sinc convolution + |x| + MaxPool1d(3) + InstanceNorm sums in one launch (sinc_pool_gemm_kernel).  wn_hi / wn_lo: the four
// shifted copies of the normalised waveform, copy e at + e * B * Np (sincnet.cu); L1 = convolution rows per item;
// pooled (B, L1 / 3, ldc) fp32 and stats (B, n_valid, 2) -- zeroed by the caller -- are written.
int sinc_pool_gemm_launch(const __half* wn_hi, const __half* wn_lo, int64_t Np, int B, int64_t L1, const __half* w_hi,
                          const __half* w_lo, int Kp, int ldw, int n_valid, int terms, float* pooled, int ldc, double* stats,
                          int num_sms, cudaStream_t st) {
    const int64_t P = L1 / 3;
    if (B <= 0 || P <= 0) return B200VAD_OK;
    int rc = gemm_ts_check(Kp, Kp, ldw, 128, 8, 2);
    if (rc) return rc;
    if ((terms != 2 && terms != 3) || n_valid < 1 || n_valid > 128 || (Np * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(wn_hi) & 15) || (reinterpret_cast<uintptr_t>(wn_lo) & 15)) {
        set_error("sinc_pool_gemm: bad arguments");
        return B200VAD_EINVAL;
    }
    SincMaps maps;
    for (int r = 0; r < 4; ++r) {
        // rows t = 4 m + r start at sample 40 m + 10 r = (40 m + 8 floor(10 r / 8)) + (10 r mod 8): copy (10 r mod 8) / 2
        const int64_t rows_r = (L1 - r + 3) / 4;
        const int e = ((10 * r) % 8) / 2, q = ((10 * r) / 8) * 8;
        const __half* bh = wn_hi + (int64_t)e * B * Np + q;
        const __half* bl = wn_lo + (int64_t)e * B * Np + q;
        if ((rc = make_tmap_3d(&maps.hi[r], bh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Kp, std::max<int64_t>(rows_r, 1), B, 80, Np * 2, SBK, 32, 1,
                               CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tmap_3d(&maps.lo[r], bl, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Kp, std::max<int64_t>(rows_r, 1), B, 80, Np * 2, SBK, 32, 1,
                               CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    }
    SincPoolParams p;
    p.w_hi = w_hi; p.w_lo = w_lo; p.pooled = pooled; p.stats = stats; p.P = P; p.ldc = ldc; p.n_valid = n_valid; p.ldw = ldw;
    p.kb = Kp / SBK; p.terms = terms;
    p.tiles_per_batch = (int)((P + SP_TILE_ROWS / 3 - 1) / (SP_TILE_ROWS / 3));
    const int64_t tiles = (int64_t)B * p.tiles_per_batch;
    if (tiles >= (1LL << 31)) { set_error("sinc_pool_gemm: too many tiles"); return B200VAD_EINVAL; }
    p.num_tiles = (int)tiles;
    const int smem = 1024 + SP_STAGES * 2 * S_TILE_BYTES + 256;
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(sinc_pool_gemm_kernel), smem))) return rc;
    const int grid = (int)std::min<int64_t>(num_sms, tiles);
    prof_begin(1, st);
    sinc_pool_gemm_kernel<<<grid, SP_THREADS, smem, st>>>(maps, p);
    prof_end(1, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// Conv1d(k = 5) + bias + MaxPool1d(3) + InstanceNorm sums in one launch (conv_pool_gemm_kernel).  a_hi / a_lo: channel-last
// planes (B, rows_in, row_stride) whose GEMM row t is the K contiguous values from element t * row_stride; L = convolution
// rows per item; weights [128][ldw] with Kp = ldw rounded-up K columns resident (Kp <= 448).  pooled (B, L / 3, ldc) fp32 and
// stats (B, n_valid, 2) -- zeroed by the caller -- are written.
int conv_pool_gemm_launch(const __half* a_hi, const __half* a_lo, int64_t row_stride, int64_t batch_stride, int B, int64_t L, int K,
                          const __half* w_hi, const __half* w_lo, int Kp, int ldw, int n_valid, const float* bias, int terms,
                          float* pooled, int ldc, double* stats, int num_sms, cudaStream_t st) {
    const int64_t P = L / 3;
    if (B <= 0 || P <= 0) return B200VAD_OK;
    if ((terms != 2 && terms != 3) || Kp % 64 != 0 || Kp > 448 || K > Kp || ldw < Kp || ldw % 8 != 0 || n_valid < 1 || n_valid > 128 || (row_stride * 2) % 16 != 0 ||
        (batch_stride * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(a_hi) & 15) || (reinterpret_cast<uintptr_t>(a_lo) & 15)) {
        set_error("conv_pool_gemm: unsupported shape K=%d Kp=%d ldw=%d n_valid=%d", K, Kp, ldw, n_valid);
        return B200VAD_EINVAL;
    }
    CUtensorMap tm_hi, tm_lo;
    int rc;
    if ((rc = make_tmap_3d(&tm_hi, a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, L, B, row_stride * 2, batch_stride * 2, SBK, CP_ROWS, 1,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_lo, a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, L, B, row_stride * 2, batch_stride * 2, SBK, CP_ROWS, 1,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    ConvPoolParams p;
    p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.pooled = pooled; p.stats = stats; p.P = P; p.ldc = ldc; p.n_valid = n_valid;
    p.ldw = ldw; p.kb = Kp / SBK; p.terms = terms;
    p.nacc = (Kp + 2 * CP_ROWS <= 512) ? 2 : 1;                 // 2 * (Kp / 2) weight columns + nacc * 64 accumulator columns <= 512
    p.tiles_per_batch = (int)((P + CP_ADV / 3 - 1) / (CP_ADV / 3));
    const int64_t tiles = (int64_t)B * p.tiles_per_batch;
    if (tiles >= (1LL << 31)) { set_error("conv_pool_gemm: too many tiles"); return B200VAD_EINVAL; }
    p.num_tiles = (int)tiles;
    const int smem = 1024 + CP_STAGES * 2 * CP_TILE_BYTES + 256;
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(conv_pool_gemm_kernel), smem))) return rc;
    const int grid = (int)std::min<int64_t>(num_sms, tiles);
    prof_begin(1, st);
    conv_pool_gemm_kernel<<<grid, CP_THREADS, smem, st>>>(tm_hi, tm_lo, p);
    prof_end(1, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// Fused head (head_fused_kernel): y planes [M][256] -> prob [M].  W1 [128][256], W2 [128][128] as fp16 hi / lo planes.
int head_fused_launch(const __half* y_hi, const __half* y_lo, int64_t M, const __half* w1_hi, const __half* w1_lo, const __half* w2_hi,
                      const __half* w2_lo, const float* b1, const float* b2, const float* wc, const float* bc, float* prob, int num_sms,
                      cudaStream_t st) {
    if (M <= 0) return B200VAD_OK;
    const int64_t tiles = (M + HF_ROWS - 1) / HF_ROWS;
    if (tiles >= (1LL << 31)) { set_error("head_fused: too many tiles"); return B200VAD_EINVAL; }
    CUtensorMap tm_hi, tm_lo;
    int rc;
    if ((rc = make_tmap_2d(&tm_hi, y_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 256, M, 256 * 2, SBK, HF_ROWS, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d(&tm_lo, y_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 256, M, 256 * 2, SBK, HF_ROWS, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    HeadFusedParams p;
    p.w1_hi = w1_hi; p.w1_lo = w1_lo; p.w2_hi = w2_hi; p.w2_lo = w2_lo; p.b1 = b1; p.b2 = b2; p.wc = wc; p.bc = bc; p.prob = prob;
    p.M = M; p.num_tiles = (int)tiles;
    const int smem = 1024 + HF_STAGES * 2 * HF_TILE_BYTES + 4 * HF_TILE_BYTES + 192 + 1024 + 64;
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(head_fused_kernel), smem))) return rc;
    const int grid = (int)std::min<int64_t>(num_sms, tiles);
    prof_begin(2, st);
    head_fused_kernel<<<grid, HF_THREADS, smem, st>>>(tm_hi, tm_lo, p);
    prof_end(2, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// 2-MMA input projection (K <= 256, both weight planes given): same contract as gemm_ts_xg_launch without `accumulate`;
// sync / sync_bytes: optional 4-byte aligned device scratch for the lockstep counters (skipped when too small)
int gemm_xg2_launch(const __half* x_hi, const __half* x_lo, int64_t lda, int B, int T, int K, const __half* w_hi,
                    const __half* w_lo, int Kp, int ldw, const float* bias, int terms, float* xg, int* sync,
                    size_t sync_bytes, int num_sms, cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    int rc = gemm_ts_check(K, Kp, ldw, 2 * kGates, lda, 2);
    if (rc) return rc;
    if (!w_lo || (terms != 2 && terms != 3)) { set_error("gemm_xg2: needs both weight planes and terms 2 or 3"); return B200VAD_EINVAL; }
    GemmX2Params p;
    p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.xg = xg; p.T = T; p.kb = Kp / SBK; p.ldw = ldw; p.terms = terms;
    p.tiles_per_blk = (T + 1) / 2;
    const int64_t tiles = (int64_t)((B + 63) / 64) * p.tiles_per_blk;
    if (tiles >= (1LL << 31)) { set_error("gemm_xg2: too many tiles"); return B200VAD_EINVAL; }
    p.num_tiles = (int)tiles;
    CUtensorMap tm_a_hi, tm_a_lo;
    if ((rc = make_tmap_3d(&tm_a_hi, x_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, T, B, lda * 2, (uint64_t)T * lda * 2, SBK, 2, 64,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_a_lo, x_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, T, B, lda * 2, (uint64_t)T * lda * 2, SBK, 2, 64,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const int smem = 1024 + X2_STAGES * 2 * S_TILE_BYTES + 512;
    int grid = (num_sms / 8) * 8;
    if (grid < 8) grid = 8;
    if ((int64_t)grid > tiles * 8) grid = (int)(tiles * 8);
    // lockstep counters: one int per (group, window) in caller-provided scratch, zeroed on the launch stream
    const int groups = grid / 8;
    const int64_t stride = ((tiles + groups - 1) / groups + X2_WINDOW - 1) / X2_WINDOW + 1;
    p.sync = nullptr; p.sync_stride = (int)std::min<int64_t>(stride, 1 << 30);
    if (sync && stride < (1 << 30) && sizeof(int) * (size_t)groups * stride <= sync_bytes) {
        p.sync = sync;
        // (a kernel, not cudaMemsetAsync: a memset may run on a copy engine and queue behind the host session's H2D copies)
        zero_i32_kernel<<<(unsigned)(((int64_t)groups * stride + 255) / 256), 256, 0, st>>>(sync, (int64_t)groups * stride);
        B200VAD_LAUNCH_CHECK();
    }
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(gemm_xg2_kernel), smem))) return rc;
    prof_begin(1, st);
    gemm_xg2_kernel<<<grid, X2_THREADS, smem, st>>>(tm_a_hi, tm_a_lo, p);
    prof_end(1, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// CTA-pair input projection (K <= 256, both weight planes): terms = 3 (hi / lo planes) or 2 (scaled planes, see
// gemm_xg2_kernel); same contract as gemm_xg2_launch.  The grid is a multiple of 8 CTAs = 4 pairs per activation tile.
int gemm_xg_pair_launch(const __half* x_a, const __half* x_b, int64_t lda, int B, int T, int K, const __half* w_hi,
                        const __half* w_lo, int Kp, int ldw, const float* bias, int terms, float* xg, int* sync,
                        size_t sync_bytes, int num_sms, cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    int rc = gemm_ts_check(K, Kp, ldw, 2 * kGates, lda, 2);
    if (rc) return rc;
    if (!w_lo || (terms != 2 && terms != 3)) { set_error("gemm_xg_pair: needs both weight planes and terms 2 or 3"); return B200VAD_EINVAL; }
    GemmXpParams p;
    p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.xg = xg; p.T = T; p.kb = Kp / SBK; p.ldw = ldw; p.terms = terms;
    p.k16 = (K + 15) / 16;
    p.tiles_per_blk = (T + 1) / 2;
    const int64_t tiles = (int64_t)((B + 63) / 64) * p.tiles_per_blk;
    if (tiles >= (1LL << 31)) { set_error("gemm_xg_pair: too many tiles"); return B200VAD_EINVAL; }
    p.num_tiles = (int)tiles;
    CUtensorMap tm_a, tm_b;
    if ((rc = make_tmap_3d(&tm_a, x_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, T, B, lda * 2, (uint64_t)T * lda * 2, SBK, 2, 32,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_b, x_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, T, B, lda * 2, (uint64_t)T * lda * 2, SBK, 2, 32,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const int smem = 1024 + XP_STAGES * 2 * XP_HALF_BYTES + 512;
    int grid = (num_sms / 8) * 8;
    if (grid < 8) grid = 8;
    if ((int64_t)grid > tiles * 8) grid = (int)(tiles * 8);
    const int groups = grid / 8;
    const int64_t stride = ((tiles + groups - 1) / groups + X2_WINDOW - 1) / X2_WINDOW + 1;
    p.sync = nullptr; p.sync_stride = (int)std::min<int64_t>(stride, 1 << 30);
    if (sync && stride < (1 << 30) && sizeof(int) * (size_t)groups * stride <= sync_bytes) {
        p.sync = sync;
        // (a kernel, not cudaMemsetAsync: a memset may run on a copy engine and queue behind the host session's H2D copies)
        zero_i32_kernel<<<(unsigned)(((int64_t)groups * stride + 255) / 256), 256, 0, st>>>(sync, (int64_t)groups * stride);
        B200VAD_LAUNCH_CHECK();
    }
    if ((rc = set_max_dynamic_smem(reinterpret_cast<const void*>(gemm_xg_pair_kernel), smem))) return rc;
    prof_begin(1, st);
    gemm_xg_pair_kernel<<<grid, XP_THREADS, smem, st>>>(tm_a, tm_b, p);
    prof_end(1, st);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// LSTM input projection (mode 3): x planes (B, T, K) with row pitch lda -> xg in the step-blocked layout described at
// gemm_ts_kernel (N = 1024 features = 2 directions x 4 gates x 128 units; xg holds ceil(B / 64) * 64 sequences).
int gemm_ts_xg_launch(const __half* x_hi, const __half* x_lo, int64_t lda, int B, int T, int K, const __half* w_hi,
                      const __half* w_lo, int Kp, int ldw, const float* bias, int accumulate, float* xg, int num_sms,
                      cudaStream_t st) {
    if (B <= 0 || T <= 0) return B200VAD_OK;
    const int nw = w_lo ? 2 : 1;
    int rc = gemm_ts_check(K, Kp, ldw, 2 * kGates, lda, nw);
    if (rc) return rc;
    GemmTsParams p = {};
    p.n_valid = 2 * kGates;
    p.w_hi = w_hi; p.w_lo = w_lo; p.bias = bias; p.c = xg; p.M = (int64_t)B * T;
    p.T = T; p.tiles_per_blk = (T + 1) / 2;
    p.num_m_tiles = ((B + 63) / 64) * p.tiles_per_blk;
    p.kb = Kp / SBK; p.nw = nw; p.Kp = Kp; p.ldw = ldw; p.accumulate = accumulate;
    CUtensorMap tm_a_hi, tm_a_lo;
    if ((rc = make_tmap_3d(&tm_a_hi, x_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, T, B, lda * 2, (uint64_t)T * lda * 2, SBK, 2, 64,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_3d(&tm_a_lo, x_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, K, T, B, lda * 2, (uint64_t)T * lda * 2, SBK, 2, 64,
                           CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    return gemm_ts_run(3, tm_a_hi, tm_a_lo, p, 2 * kGates, num_sms, st);
}

}  // namespace b200vad
