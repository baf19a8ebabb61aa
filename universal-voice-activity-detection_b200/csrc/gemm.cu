// Row-streaming GEMM  C[M,N] = act(A[M,K] . W[N,K]^T + bias)  with fp16 tensor-core
// operands, fp32 accumulation and optional split-precision terms:
//   terms = 1 : A_hi*W_hi                       (A already fp16, or precision not needed)
//   terms = 2 : A_hi*W_hi + A_hi*W_lo           (fp16 activations, ~fp32 weights)
//   terms = 3 : A_hi*W_hi + A_hi*W_lo + A_lo*W_hi  (fp32 activations split on the fly)
// A rows are addressed as  A + (m / rows_per_batch) * batch_stride + (m % rows_per_batch) * lda,
// which makes a Conv1d over a channel-last (B, T, C) tensor a plain GEMM with overlapping
// rows (lda = stride*C, K = kernel*C): used for the LSTM input projections
// (PyanNet2.py:95,170), the head linears (PyanNet2.py:183-187) and the SincNet
// convolutions (sincnet.py:50-69).
#include "kernels.cuh"

namespace b200vad {

// The warp-MMA (mma.sync) GEMM is the cross-validation path of round 1 (b200vad_set_impl(1)); it is compiled only with
// `make VALIDATE=1` (-DB200VAD_VALIDATE): the product library carries the tcgen05 kernels only.
#ifdef B200VAD_VALIDATE
constexpr int GBM = 128, GBK = 32, GPAD = 8, GLD = GBK + GPAD;   // smem row = 40 halves (80 B)

template <int BN, int TERMS, typename AT, int VEC>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs p) {
    constexpr int WARPS_N = BN / 32, WARPS_M = 8 / WARPS_N, WTM = GBM / WARPS_M, MT = WTM / 16, NT = 4;
    constexpr bool A_LO = (TERMS == 3);
    constexpr bool W_LO = (TERMS >= 2);
    __shared__ __align__(16) __half As_hi[GBM * GLD];
    __shared__ __align__(16) __half As_lo[A_LO ? GBM * GLD : 8];
    __shared__ __align__(16) __half Ws_hi[BN * GLD];
    __shared__ __align__(16) __half Ws_lo[W_LO ? BN * GLD : 8];
    __shared__ int64_t row_off[GBM];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int64_t m0 = (int64_t)blockIdx.y * GBM;
    const int n0 = blockIdx.x * BN;
    const AT* A = reinterpret_cast<const AT*>(p.A);

    if (tid < GBM) {
        int64_t m = m0 + tid;
        if (m >= p.M) m = p.M - 1;   // clamp (results for m >= M are never stored)
        row_off[tid] = (m / p.rows_per_batch) * p.a_batch_stride + (m % p.rows_per_batch) * p.lda;
    }
    __syncthreads();

    // global->register staging layout
    constexpr int A_ELEMS = GBM * GBK / 256;        // 16 elements per thread
    constexpr int A_ITERS = A_ELEMS / VEC;
    constexpr int A_TPR = GBK / VEC;                 // threads per row
    constexpr int A_RSTEP = 256 / A_TPR;
    AT a_reg[A_ELEMS];
    constexpr int W_ITERS = BN * GBK / 8 / 256;      // uint4 (8 halves) per thread: BN=128 -> 2, 64 -> 1
    uint4 w_hi_reg[W_ITERS], w_lo_reg[W_LO ? W_ITERS : 1];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int r = tid / A_TPR + i * A_RSTEP, c = (tid % A_TPR) * VEC;
            const AT* src = A + row_off[r] + k0 + c;
            if (VEC == 1) {
                a_reg[i] = (k0 + c < p.K) ? src[0] : AT(0);
            } else if (k0 + c + VEC <= p.K) {
                if (sizeof(AT) * VEC == 16) *reinterpret_cast<uint4*>(&a_reg[i * VEC]) = __ldg(reinterpret_cast<const uint4*>(src));
                else *reinterpret_cast<uint2*>(&a_reg[i * VEC]) = __ldg(reinterpret_cast<const uint2*>(src));
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) a_reg[i * VEC + v] = (k0 + c + v < p.K) ? src[v] : AT(0);
            }
        }
#pragma unroll
        for (int i = 0; i < W_ITERS; ++i) {
            int idx = tid + i * 256;
            int r = idx / 4, c = (idx % 4) * 8;
            int n = n0 + r;
            uint4 z = make_uint4(0, 0, 0, 0);
            w_hi_reg[i] = (n < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.W_hi + (int64_t)n * p.Kp + k0 + c)) : z;
            if (W_LO) w_lo_reg[i] = (n < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.W_lo + (int64_t)n * p.Kp + k0 + c)) : z;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int r = tid / A_TPR + i * A_RSTEP, c = (tid % A_TPR) * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (sizeof(AT) == 2) {
                    As_hi[r * GLD + c + v] = *reinterpret_cast<const __half*>(&a_reg[i * VEC + v]);
                } else {
                    float x = *reinterpret_cast<const float*>(&a_reg[i * VEC + v]);
                    __half hi, lo;
                    split_f16(x, hi, lo);
                    As_hi[r * GLD + c + v] = hi;
                    if (A_LO) As_lo[r * GLD + c + v] = lo;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < W_ITERS; ++i) {
            int idx = tid + i * 256;
            int r = idx / 4, c = (idx % 4) * 8;
            *reinterpret_cast<uint4*>(&Ws_hi[r * GLD + c]) = w_hi_reg[i];
            if (W_LO) *reinterpret_cast<uint4*>(&Ws_lo[r * GLD + c]) = w_lo_reg[i];
        }
    };

    float acc[MT][NT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

    const int nk = (p.K + GBK - 1) / GBK;
    load_tiles(0);
    for (int kt = 0; kt < nk; ++kt) {
        __syncthreads();          // previous tile fully consumed
        store_tiles();
        __syncthreads();
        if (kt + 1 < nk) load_tiles((kt + 1) * GBK);
#pragma unroll
        for (int ks = 0; ks < GBK; ks += 16) {
            uint32_t a_hi[MT][4], a_lo[A_LO ? MT : 1][4], b_hi[NT][2], b_lo[W_LO ? NT : 1][2];
            // A fragments: ldmatrix x4 -> (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15)
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                int r = wm * WTM + i * 16 + (lane & 15), c = ks + (lane >> 4) * 8;
                ldmatrix_x4(a_hi[i], smem_u32(&As_hi[r * GLD + c]));
                if (A_LO) ldmatrix_x4(a_lo[i], smem_u32(&As_lo[r * GLD + c]));
            }
            // B fragments: W rows are n, k contiguous; x4 -> two n-tiles x two k-halves
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
                int r = wn * 32 + j * 8 + (lane & 7) + ((lane >> 4) << 3), c = ks + ((lane >> 3) & 1) * 8;
                uint32_t t[4];
                ldmatrix_x4(t, smem_u32(&Ws_hi[r * GLD + c]));
                b_hi[j][0] = t[0]; b_hi[j][1] = t[1]; b_hi[j + 1][0] = t[2]; b_hi[j + 1][1] = t[3];
                if (W_LO) {
                    ldmatrix_x4(t, smem_u32(&Ws_lo[r * GLD + c]));
                    b_lo[j][0] = t[0]; b_lo[j][1] = t[1]; b_lo[j + 1][0] = t[2]; b_lo[j + 1][1] = t[3];
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (A_LO) mma_16816(acc[i][j], a_lo[i], b_hi[j]);   // small terms first
                    if (W_LO) mma_16816(acc[i][j], a_hi[i], b_lo[j]);
                    mma_16816(acc[i][j], a_hi[i], b_hi[j]);
                }
        }
    }

    // epilogue
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            int n = n0 + wn * 32 + j * 8 + 2 * t4;
            if (n >= p.N) continue;
            float b0 = p.bias ? __ldg(p.bias + n) : 0.f;
            float b1 = (p.bias && n + 1 < p.N) ? __ldg(p.bias + n + 1) : 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int64_t m = m0 + wm * WTM + i * 16 + g + h * 8;
                if (m >= p.M) continue;
                float v0 = acc[i][j][2 * h] + b0, v1 = acc[i][j][2 * h + 1] + b1;
                if (p.act == 1) {
                    v0 = v0 > 0.f ? v0 : 0.01f * v0;
                    v1 = v1 > 0.f ? v1 : 0.01f * v1;
                } else if (p.act == 2) {
                    v0 = fabsf(v0);
                    v1 = fabsf(v1);
                }
                if (p.c_half) {
                    __half* c = reinterpret_cast<__half*>(p.C) + m * p.ldc + n;
                    if (n + 1 < p.N && ((p.ldc & 1) == 0)) *reinterpret_cast<__half2*>(c) = __floats2half2_rn(v0, v1);
                    else {
                        c[0] = __float2half_rn(v0);
                        if (n + 1 < p.N) c[1] = __float2half_rn(v1);
                    }
                } else {
                    float* c = reinterpret_cast<float*>(p.C) + m * p.ldc + n;
                    if (n + 1 < p.N && ((p.ldc & 1) == 0)) *reinterpret_cast<float2*>(c) = make_float2(v0, v1);
                    else {
                        c[0] = v0;
                        if (n + 1 < p.N) c[1] = v1;
                    }
                }
            }
        }
}
#endif  // B200VAD_VALIDATE

// weights fp32 [N][K] -> fp16 hi / lo [N][Kp], zero padded
__global__ void split_weights_kernel(const float* __restrict__ w, int N, int K, int Kp, __half* __restrict__ hi,
                                     __half* __restrict__ lo, int gate_scale) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * Kp) return;
    int n = (int)(idx / Kp), k = (int)(idx % Kp);
    float x = (k < K) ? w[(int64_t)n * K + k] : 0.f;
    if (gate_scale) x *= lstm_gate_scale(n);
    __half h, l;
    split_f16(x, h, l);
    hi[idx] = h;
    if (lo) lo[idx] = l;
}

int split_weights(const float* w, int N, int K, int Kp, __half* hi, __half* lo, cudaStream_t stream, int gate_scale) {
    int64_t tot = (int64_t)N * Kp;
    split_weights_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(w, N, K, Kp, hi, lo, gate_scale);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

#ifdef B200VAD_VALIDATE
template <int BN, int TERMS, typename AT, int VEC>
static int launch_one(const GemmArgs& a, cudaStream_t stream) {
    dim3 grid((a.N + BN - 1) / BN, (unsigned)((a.M + GBM - 1) / GBM));
    prof_begin(2, stream);
    gemm_kernel<BN, TERMS, AT, VEC><<<grid, 256, 0, stream>>>(a);
    prof_end(2, stream);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// a_half: A is fp16 (terms 1 or 2) else fp32 (terms 1 or 3)
int gemm_launch(const GemmArgs& a, int a_half, int terms, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return B200VAD_OK;
    if (a.M / GBM >= 65535) {
        set_error("gemm: M too large for one launch (%lld)", (long long)a.M);
        return B200VAD_EINVAL;
    }
    const bool wide = a.N > 64;
    if (a_half) {
        bool vec = (a.lda % 8 == 0) && (a.a_batch_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
        if (!vec) {
            set_error("gemm: fp16 A must be 16-byte aligned with lda %% 8 == 0");
            return B200VAD_EINVAL;
        }
        if (terms >= 2) return wide ? launch_one<128, 2, __half, 8>(a, stream) : launch_one<64, 2, __half, 8>(a, stream);
        return wide ? launch_one<128, 1, __half, 8>(a, stream) : launch_one<64, 1, __half, 8>(a, stream);
    }
    bool vec = (a.lda % 4 == 0) && (a.a_batch_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    if (terms >= 3) {
        if (vec) return wide ? launch_one<128, 3, float, 4>(a, stream) : launch_one<64, 3, float, 4>(a, stream);
        return wide ? launch_one<128, 3, float, 1>(a, stream) : launch_one<64, 3, float, 1>(a, stream);
    }
    if (vec) return wide ? launch_one<128, 1, float, 4>(a, stream) : launch_one<64, 1, float, 4>(a, stream);
    return wide ? launch_one<128, 1, float, 1>(a, stream) : launch_one<64, 1, float, 1>(a, stream);
}

#else
int gemm_launch(const GemmArgs&, int, int, cudaStream_t) {
    set_error("the warp-MMA validation kernels are not in this build (make VALIDATE=1)");
    return B200VAD_ESTATE;
}
#endif

}  // namespace b200vad
