// Bidirectional LSTM stack + linear head of PyanNet2 / PyanNet
// (src/models/segmentation/PyanNet2.py:95,169-187; PyanNet.py:105,181-197; torch nn.LSTM
// semantics: gates i,f,g,o; g = W_ih x + b_ih + W_hh h + b_hh; c = f*c + i*g; h = o*tanh(c)).
//
// Per layer: (1) the input projection for all timesteps and both directions is one
// split-precision tensor-core GEMM (gemm.cu) writing xg[B*T, 2*512] fp32 with both biases
// folded in; (2) a persistent recurrent kernel: one CTA owns 32 sequences of one direction,
// W_hh (512x128 fp16) stays resident in shared memory for all T steps, each of 16 warps owns
// 8 hidden units x 4 gates so that i,f,g,o of a cell land in the same thread's accumulators,
// the cell state lives in registers (fp32), h_t goes to shared memory (fp16 operand of the
// next step) and to HBM (fp32 layer output: rounding the layer outputs to fp16 is the
// dominant error term of the stack, see DESIGN.md 'precision').  xg for step t+1 is prefetched during step t.
#include "kernels.cuh"

namespace b200vad {

#ifdef B200VAD_VALIDATE   // warp-MMA recurrence: cross-validation path of round 1, `make VALIDATE=1` only
constexpr int RB = 32;                 // sequences per CTA
constexpr int RLD = kHidden + 8;       // padded smem row (halves)
constexpr int RTHREADS = 512;

struct RecSmem {
    __half w[kGates * RLD];            // W_hh, rows = gate*128 + unit
    __half h[2][RB * RLD];             // double-buffered hidden state
};

// xg: [B][T][2][512] fp32;  y: [B][T][256] fp32;  whh: [2][512][128] fp16
__global__ void __launch_bounds__(RTHREADS, 1)
lstm_recurrent_kernel(const float* __restrict__ xg, const __half* __restrict__ whh, float* __restrict__ y, int B, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RecSmem& sm = *reinterpret_cast<RecSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dir = blockIdx.y;
    const int b0 = blockIdx.x * RB;
    const int g = lane >> 2, t4 = lane & 3;
    const int u0 = warp * 8;           // this warp's hidden units [u0, u0+8)

    // W_hh -> smem (16-byte chunks), h0 = 0
    const __half* wsrc = whh + (size_t)dir * kGates * kHidden;
    for (int i = tid; i < kGates * kHidden / 8; i += RTHREADS) {
        int r = i / (kHidden / 8), c = (i % (kHidden / 8)) * 8;
        *reinterpret_cast<uint4*>(&sm.w[r * RLD + c]) = __ldg(reinterpret_cast<const uint4*>(wsrc + r * kHidden + c));
    }
    for (int i = tid; i < 2 * RB * RLD / 2; i += RTHREADS) reinterpret_cast<uint32_t*>(&sm.h[0][0])[i] = 0u;

    // per-thread rows: (mt, hh) -> row mt*16 + g + hh*8
    int64_t rowbase[2][2];
    bool rowok[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            int b = b0 + mt * 16 + g + hh * 8;
            rowok[mt][hh] = b < B;
            rowbase[mt][hh] = (int64_t)min(b, B - 1) * T;
        }
    const int col = dir * kGates + u0 + 2 * t4;     // + q*128 for gate q

    float cst[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) cst[a][b][0] = cst[a][b][1] = 0.f;

    float2 nxt[2][2][4];
    auto prefetch = [&](int t) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float* p = xg + (rowbase[mt][hh] + t) * (2 * kGates) + col;
#pragma unroll
                for (int q = 0; q < 4; ++q) nxt[mt][hh][q] = __ldg(reinterpret_cast<const float2*>(p + q * kHidden));
            }
    };
    prefetch(dir == 0 ? 0 : T - 1);
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        const __half* hc = sm.h[s & 1];
        __half* hn = sm.h[(s + 1) & 1];
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[mt][q][0] = nxt[mt][0][q].x; acc[mt][q][1] = nxt[mt][0][q].y;
                acc[mt][q][2] = nxt[mt][1][q].x; acc[mt][q][3] = nxt[mt][1][q].y;
            }
        if (s + 1 < T) prefetch(dir == 0 ? t + 1 : t - 1);

#pragma unroll
        for (int ks = 0; ks < kHidden; ks += 16) {
            uint32_t a[2][4], bf[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                ldmatrix_x4(a[mt], smem_u32(&hc[(mt * 16 + (lane & 15)) * RLD + ks + (lane >> 4) * 8]));
#pragma unroll
            for (int q = 0; q < 4; q += 2) {
                int r = (q + (lane >> 4)) * kHidden + u0 + (lane & 7), c = ks + ((lane >> 3) & 1) * 8;
                uint32_t tt[4];
                ldmatrix_x4(tt, smem_u32(&sm.w[r * RLD + c]));
                bf[q][0] = tt[0]; bf[q][1] = tt[1]; bf[q + 1][0] = tt[2]; bf[q + 1][1] = tt[3];
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int q = 0; q < 4; ++q) mma_16816(acc[mt][q], a[mt], bf[q]);
        }

        // pointwise cell update
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float hv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float ig = sigmoid_pre(acc[mt][0][2 * hh + e]);
                    float fg = sigmoid_pre(acc[mt][1][2 * hh + e]);
                    float gg = tanh_pre(acc[mt][2][2 * hh + e]);
                    float og = sigmoid_pre(acc[mt][3][2 * hh + e]);
                    float c = fmaf(fg, cst[mt][hh][e], ig * gg);
                    cst[mt][hh][e] = c;
                    hv[e] = og * tanh_acc(c);
                }
                __half2 h2 = __floats2half2_rn(hv[0], hv[1]);
                int r = mt * 16 + g + hh * 8;
                *reinterpret_cast<__half2*>(&hn[r * RLD + u0 + 2 * t4]) = h2;
                if (rowok[mt][hh])
                    *reinterpret_cast<float2*>(&y[(rowbase[mt][hh] + t) * (2 * kHidden) + dir * kHidden + u0 + 2 * t4]) =
                        make_float2(hv[0], hv[1]);
            }
        __syncthreads();
    }
}

int lstm_recurrent_launch(const float* xg, const __half* whh, float* y, int B, int T, cudaStream_t stream) {
    if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(lstm_recurrent_kernel), (int)sizeof(RecSmem))) return rc;
    dim3 grid((B + RB - 1) / RB, 2);
    prof_begin(0, stream);
    lstm_recurrent_kernel<<<grid, RTHREADS, sizeof(RecSmem), stream>>>(xg, whh, y, B, T);
    prof_end(0, stream);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}
#else
int lstm_recurrent_launch(const float*, const __half*, float*, int, int, cudaStream_t) {
    set_error("the warp-MMA validation kernels are not in this build (make VALIDATE=1)");
    return B200VAD_ESTATE;
}
#endif

// ---------------------------------------------------------------- classifier: sigmoid(wc . z + bc), one warp per row
__global__ void __launch_bounds__(256) classifier_kernel(const float* __restrict__ z, int64_t rows, const float* __restrict__ wc,
                                                         const float* __restrict__ bc, float* __restrict__ prob) {
    int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    int lane = threadIdx.x & 31;
    float4 v = __ldg(reinterpret_cast<const float4*>(z + row * kHidden) + lane);
    float4 w = __ldg(reinterpret_cast<const float4*>(wc) + lane);
    float s = warp_sum(v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w);
    if (lane == 0) prob[row] = 1.f / (1.f + expf(-(s + __ldg(bc))));
}

int classifier_launch(const float* z, int64_t rows, const float* wc, const float* bc, float* prob, cudaStream_t stream) {
    if (rows == 0) return B200VAD_OK;
    classifier_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(z, rows, wc, bc, prob);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// fp16 gate-major copy of W_hh
// W_hh and the summed biases carry the gate exponent scaling (common.cuh lstm_gate_scale)
__global__ void pack_whh_kernel(const float* __restrict__ w, __half* __restrict__ out, __half* __restrict__ out_lo, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        __half hi, lo;
        split_f16(w[i] * lstm_gate_scale(i / kHidden), hi, lo);
        out[i] = hi;
        if (out_lo) out_lo[i] = lo;                      // second plane: the fused layer (lstm_fused.cu) forms W' from it
    }
}
__global__ void add_bias_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (a[i] + b[i]) * lstm_gate_scale(i);
}

int pack_whh(const float* w, __half* out, cudaStream_t stream, __half* out_lo) {
    int n = kGates * kHidden;
    pack_whh_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, out, out_lo, n);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}
int add_bias(const float* a, const float* b, float* out, int n, cudaStream_t stream) {
    add_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, out, n);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
