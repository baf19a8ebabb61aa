// Threshold + median smoothing + run-length segment extraction on the GPU.
//
// Replaces src/utils/helper.py:66-97 (torch.where(x < 0.5, 0, 1) -> .cpu() -> scipy
// medfilt(k) per row -> .to("cuda")) and the per-frame Python loop of
// src/scripts/predict.py:472-490.  On 0/1 data the zero-padded median of an odd window k
// is 1 iff the window sum >= (k+1)/2, so the filter is a sliding popcount: ballot words ->
// prefix sums in shared memory -> two lookups per frame.  HBM traffic: 4 B/frame in,
// 1 B (or 8 B for the reference's int64 view) out.
#include "kernels.cuh"

namespace b200vad {

constexpr int kMedTile = 2048;      // frames per CTA
constexpr int kMedMaxHalo = 512;    // supports kernel sizes up to 1025

// prob: [B][T] fp32.  out: [B][T] uint8 (elem=1) or int64 (elem=8).  NaN -> 1 (x < thr is false).
__global__ void __launch_bounds__(256)
threshold_median_kernel(const float* __restrict__ prob, int B, int64_t T, float thr, int half, void* __restrict__ out, int elem,
                        int32_t* __restrict__ near_count, float near_tol) {
    __shared__ int pre[kMedTile + 2 * kMedMaxHalo + 1];
    __shared__ int warp_tot[8];
    const int b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * kMedTile;
    const int span = kMedTile + 2 * half;           // frames [t0 - half, t0 + tile + half)
    const float* row = prob + (int64_t)b * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int near = 0;
    // exclusive prefix sum of decisions over the span, 256 frames per round
    int carry = 0;
    for (int base = 0; base < span; base += 256) {
        int s = base + tid;
        int64_t t = t0 - half + s;
        int d = 0;
        if (s < span && t >= 0 && t < T) {
            float p = __ldg(row + t);
            d = (p < thr) ? 0 : 1;
            if (near_count && s >= half && s < half + kMedTile && fabsf(p - thr) <= near_tol) near++;
        }
        unsigned m = __ballot_sync(0xffffffffu, d);
        int incl = __popc(m & (0xffffffffu >> (31 - lane)));
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int off = carry;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            int v = warp_tot[w];
            if (w < warp) off += v;
            carry += v;
        }
        if (s < span) pre[s + 1] = off + incl;
        __syncthreads();
    }
    if (tid == 0) pre[0] = 0;
    __syncthreads();
    const int need = half + 1;                      // (k+1)/2 with k = 2*half+1
    for (int i = tid; i < kMedTile; i += 256) {
        int64_t t = t0 + i;
        if (t >= T) break;
        int sum = pre[i + 2 * half + 1] - pre[i];   // window [t-half, t+half]
        int v = sum >= need;
        if (elem == 1) reinterpret_cast<uint8_t*>(out)[(int64_t)b * T + t] = (uint8_t)v;
        else reinterpret_cast<int64_t*>(out)[(int64_t)b * T + t] = (int64_t)v;
    }
    if (near_count) {
        near = (int)warp_sum((float)near);
        if (lane == 0 && near) atomicAdd(near_count, near);
    }
}

int threshold_median_launch(const float* prob, int B, int64_t T, float thr, int kernel, void* out, int elem,
                            int32_t* near_count, float near_tol, cudaStream_t stream) {
    if (B == 0 || T == 0) return B200VAD_OK;
    if (kernel < 1 || (kernel & 1) == 0 || kernel / 2 > kMedMaxHalo) {
        set_error("threshold_median: kernel size must be odd and <= %d", 2 * kMedMaxHalo + 1);
        return B200VAD_EINVAL;
    }
    int half = kernel / 2;
    dim3 grid((unsigned)((T + kMedTile - 1) / kMedTile), B);
    threshold_median_kernel<<<grid, 256, 0, stream>>>(prob, B, T, thr, half, out, elem, near_count, near_tol);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

// ---------------------------------------------------------------- run-length segments
// dec: flat uint8 stream; stream r = dec[offsets[r], offsets[r+1]).  A run of >= min_run
// consecutive non-zero frames yields (r, first, last_inclusive), frame indices relative to the
// stream start (predict.py:472-490 keeps a run iff round(end,2) - round(start,2) > 0, i.e.
// length >= 2).  Pass 1 counts per stream, an exclusive scan orders the output by (r, first),
// pass 2 writes.  One CTA per stream, 32 frames per thread-iteration via ballot.
template <bool WRITE>
__global__ void __launch_bounds__(256)
segments_kernel(const uint8_t* __restrict__ dec, const int64_t* __restrict__ offsets, int64_t uniform_T, int R, int min_run,
                int row_base, int32_t* __restrict__ counts, const int64_t* __restrict__ seg_off, int32_t* __restrict__ seg, int64_t cap) {
    const int r = blockIdx.x;
    const int64_t beg = offsets ? offsets[r] : (int64_t)r * uniform_T;
    const int64_t len = offsets ? offsets[r + 1] - beg : uniform_T;
    const uint8_t* d = dec + beg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int wsum[8];
    __shared__ int s_carry;
    __shared__ int s_open_start;     // start frame of a run still open at the chunk boundary (or -1)
    if (tid == 0) { s_carry = 0; s_open_start = -1; }
    __syncthreads();
    int64_t out_base = WRITE ? seg_off[r] : 0;
    // every chunk = 256 frames, one frame per thread; runs are closed at their END frame:
    // an end at frame e (d[e]=1, d[e+1]=0 or e=len-1) emits (start, e) where start is found by a
    // backward scan limited to the chunk, falling back to the carried open start.
    for (int64_t base = 0; base < len; base += 256) {
        int64_t t = base + tid;
        int cur = (t < len) ? (d[t] != 0) : 0;
        int nxt = (t + 1 < len) ? (d[t + 1] != 0) : 0;
        unsigned m = __ballot_sync(0xffffffffu, cur);
        __shared__ unsigned words[8];
        if (lane == 0) words[warp] = m;
        __syncthreads();
        int is_end = cur && !nxt;
        int start = -1;
        int valid = 0;
        if (is_end) {
            // backward scan for the first zero before tid within this chunk
            int pos = tid;           // position in chunk
            int w = pos >> 5;
            unsigned inv = ~words[w] & (0xffffffffu >> (31 - (pos & 31)));   // zeros at or before pos in word
            int zero_at = -1;
            while (true) {
                if (inv) { zero_at = (w << 5) + 31 - __clz(inv); break; }
                if (--w < 0) break;
                inv = ~words[w];
            }
            int64_t st;
            if (zero_at >= 0) st = base + zero_at + 1;
            else st = (s_open_start >= 0) ? (int64_t)s_open_start : base;    // run spans the chunk start
            start = (int)st;
            valid = (t - st + 1) >= min_run;
        }
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        int incl = __popc(vm & (0xffffffffu >> (31 - lane)));
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int off = s_carry, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) off += wsum[w];
            tot += wsum[w];
        }
        if (WRITE && valid) {
            int64_t o = out_base + off + incl - 1;
            if (o < cap) {
                seg[o * 3 + 0] = r + row_base;
                seg[o * 3 + 1] = start;
                seg[o * 3 + 2] = (int)t;
            }
        }
        __syncthreads();
        if (tid == 255) {
            // s_open_start := start of the run still open at the end of this chunk (or -1)
            if (!cur) {
                s_open_start = -1;
            } else {
                int w = 7, zero_at = -1;
                unsigned inv = ~words[7];
                while (true) {
                    if (inv) { zero_at = (w << 5) + 31 - __clz(inv); break; }
                    if (--w < 0) break;
                    inv = ~words[w];
                }
                if (zero_at >= 0) s_open_start = (int)(base + zero_at + 1);
                else if (s_open_start < 0) s_open_start = (int)base;
            }
            s_carry += tot;
        }
        __syncthreads();
    }
    if (!WRITE && tid == 0) counts[r] = s_carry;
}

// exclusive scan of counts[R] -> seg_off[R+1] (int64); single CTA, R up to millions is fine (serial chunks)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ counts, int R, int64_t* __restrict__ seg_off) {
    __shared__ long long wtot[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < R; base += 1024) {
        int i = base + tid;
        long long v = (i < R) ? counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        long long off = carry;
        for (int w = 0; w < warp; ++w) off += wtot[w];
        if (i < R) seg_off[i] = off + incl - v;
        __syncthreads();
        if (tid == 1023) carry = off + incl;
        __syncthreads();
    }
    if (tid == 0) seg_off[R] = carry;
}

int segments_launch(const uint8_t* dec, const int64_t* offsets, int64_t uniform_T, int R, int min_run, int row_base,
                    int32_t* counts, int64_t* seg_off, int32_t* seg, int64_t cap, cudaStream_t stream) {
    if (R == 0) return B200VAD_OK;
    segments_kernel<false><<<R, 256, 0, stream>>>(dec, offsets, uniform_T, R, min_run, row_base, counts, nullptr, nullptr, 0);
    B200VAD_LAUNCH_CHECK();
    scan_counts_kernel<<<1, 1024, 0, stream>>>(counts, R, seg_off);
    B200VAD_LAUNCH_CHECK();
    segments_kernel<true><<<R, 256, 0, stream>>>(dec, offsets, uniform_T, R, min_run, row_base, counts, seg_off, seg, cap);
    B200VAD_LAUNCH_CHECK();
    return B200VAD_OK;
}

}  // namespace b200vad
