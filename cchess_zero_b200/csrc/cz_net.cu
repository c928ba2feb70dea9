// cz_net.cu -- hand-written sm_100a kernels for the two ends of the policy-value network
// (policy_value_network.py:45-74): the first convolution evaluated straight from board bytes, and the
// fused policy / value heads.  The residual tower in between stays on the library tcgen05 path.
//
//   k_first_conv : canonical board bytes -> conv3x3(14->128)+bias+ReLU output, fp16 NHWC [B][90][128].
//                  The 14-plane input is one-hot and <= 32 of its 1260 cells are set, so the convolution is a
//                  gather-add of weight rows: out[cell][:] = b + sum over the 3x3 neighbourhood of W[tap][piece][:].
//                  The [9][10][14] tensor (and the reference's rank*9+file indexing, main.py:550-555) is never
//                  materialised: image cell (r, f) reads canonical board byte r*9+f.
//   k_head_conv  : conv1x1(128->3)+bias+ReLU (policy 2 ch + value 1 ch), then the value MLP
//                  90 -> 256 ReLU -> 1 tanh; writes hp fp16 [B][192] (flatten order (h, w, c), zero padded).
//   k_policy_fc  : logits[B][2086] = hp . Wp^T + bp with mma.sync m16n8k16 (fp16 in, fp32 accumulate, fp32 out).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cchess_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// first convolution from board bytes
// ------------------------------------------------------------------------------------------
constexpr int FC_POS = 4;        // positions per CTA
constexpr int FC_THREADS = 256;  // 64 channel-pair lanes x 4 cell groups

__global__ void __launch_bounds__(FC_THREADS) k_first_conv(const uint8_t *__restrict__ boards, int B, const __half2 *__restrict__ w /* [9][14][64] pairs */,
                                                            const float2 *__restrict__ bias /* [64] */, __half2 *__restrict__ out /* [B][90][64] */) {
    __shared__ __half2 sw[9 * 14 * 64];   // 31.5 KB: fp16 weights (the tensor-core path rounds them the same way), fp32 accumulation
    __shared__ uint8_t sb[FC_POS][96];
    for (int i = threadIdx.x; i < 9 * 14 * 64; i += FC_THREADS) sw[i] = w[i];
    const int p0 = blockIdx.x * FC_POS;
    for (int i = threadIdx.x; i < FC_POS * 24; i += FC_THREADS) {
        const int p = i / 24, wd = i - p * 24;
        reinterpret_cast<uint32_t *>(sb[p])[wd] = p0 + p < B ? reinterpret_cast<const uint32_t *>(boards + (size_t)(p0 + p) * 96)[wd] : 0u;
    }
    __syncthreads();
    const int cp = threadIdx.x & 63, grp = threadIdx.x >> 6;   // channel pair, cell group
    const float2 bv = bias[cp];
    for (int idx = grp; idx < FC_POS * 90; idx += 4) {
        const int p = idx / 90, cell = idx - p * 90;
        if (p0 + p >= B) break;
        const int r = cell / 10, f = cell - r * 10;
        float2 acc = bv;
#pragma unroll
        for (int dr = -1; dr <= 1; dr++) {
            const int rr = r + dr;
            if (rr < 0 || rr > 8) continue;
#pragma unroll
            for (int df = -1; df <= 1; df++) {
                const int ff = f + df;
                if (ff < 0 || ff > 9) continue;
                const int pc = sb[p][rr * 9 + ff];            // the reference's indexing: cell (rank, file) <- s[rank*9+file]
                if (pc) {
                    const float2 wv = __half22float2(sw[(((dr + 1) * 3 + (df + 1)) * 14 + (pc - 1)) * 64 + cp]);
                    acc.x += wv.x;
                    acc.y += wv.y;
                }
            }
        }
        out[((size_t)(p0 + p) * 90 + cell) * 64 + cp] = __floats2half2_rn(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f));
    }
}

// ------------------------------------------------------------------------------------------
// heads: 1x1 conv (3 outputs) + value MLP
// ------------------------------------------------------------------------------------------
constexpr int HC_POS = 4;
constexpr int HC_THREADS = 256;

__global__ void __launch_bounds__(HC_THREADS) k_head_conv(const __half *__restrict__ x /* [B][90][128] */, int B, const float *__restrict__ wh /* [3][128] */,
                                                           const float *__restrict__ bh /* [3] */, const float *__restrict__ w1t /* [90][256] */,
                                                           const float *__restrict__ b1 /* [256] */, const float *__restrict__ w2 /* [256] */, float b2,
                                                           __half *__restrict__ hp /* [B][192] */, float *__restrict__ value /* [B] */) {
    __shared__ float swh[3][128];
    __shared__ float hv[HC_POS][96];
    __shared__ float red[HC_POS][8];
    for (int i = threadIdx.x; i < 3 * 128; i += HC_THREADS) swh[i / 128][i % 128] = wh[i];
    __syncthreads();
    const int p0 = blockIdx.x * HC_POS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half_id = lane >> 4, l16 = lane & 15;   // two cells per warp iteration, 16 lanes x 8 channels each
    float wreg[3][8];
#pragma unroll
    for (int o = 0; o < 3; o++)
#pragma unroll
        for (int k = 0; k < 8; k++) wreg[o][k] = swh[o][l16 * 8 + k];
    const float bh0 = bh[0], bh1 = bh[1], bh2 = bh[2];
    for (int it = warp; it < HC_POS * 45; it += 8) {
        const int cidx = it * 2 + half_id;           // 0 .. HC_POS*90-1
        const int p = cidx / 90, cell = cidx - p * 90;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (p0 + p < B) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(x + ((size_t)(p0 + p) * 90 + cell) * 128 + l16 * 8);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 v = __half22float2(h2[k]);
                s0 += v.x * wreg[0][2 * k] + v.y * wreg[0][2 * k + 1];
                s1 += v.x * wreg[1][2 * k] + v.y * wreg[1][2 * k + 1];
                s2 += v.x * wreg[2][2 * k] + v.y * wreg[2][2 * k + 1];
            }
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (l16 == 0 && p0 + p < B) {
            // flatten order of tf.reshape on NHWC (policy_value_network.py:62, 72): index = cell*2 + c
            const __half2 pv = __floats2half2_rn(fmaxf(s0 + bh0, 0.f), fmaxf(s1 + bh1, 0.f));
            *reinterpret_cast<__half2 *>(hp + (size_t)(p0 + p) * 192 + cell * 2) = pv;
            hv[p][cell] = fmaxf(s2 + bh2, 0.f);
        }
    }
    // zero padding of hp columns 180..191
    if (threadIdx.x < HC_POS * 12) {
        const int p = threadIdx.x / 12, c = threadIdx.x % 12;
        if (p0 + p < B) hp[(size_t)(p0 + p) * 192 + 180 + c] = __float2half(0.f);
    }
    __syncthreads();
    // value head: fc1 (90 -> 256) + ReLU, fc2 (256 -> 1), tanh   (policy_value_network.py:73-74)
    const int t = threadIdx.x;
    float a[HC_POS];
#pragma unroll
    for (int p = 0; p < HC_POS; p++) a[p] = b1[t];
    for (int k = 0; k < 90; k++) {
        const float wv = w1t[k * 256 + t];
#pragma unroll
        for (int p = 0; p < HC_POS; p++) a[p] += wv * hv[p][k];
    }
    const float w2v = w2[t];
#pragma unroll
    for (int p = 0; p < HC_POS; p++) {
        float s = fmaxf(a[p], 0.f) * w2v;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[p][warp] = s;
    }
    __syncthreads();
    if (t < HC_POS && p0 + t < B) {
        float s = b2;
#pragma unroll
        for (int wq = 0; wq < 8; wq++) s += red[t][wq];
        value[p0 + t] = tanhf(s);
    }
}

// ------------------------------------------------------------------------------------------
// policy FC on tensor cores (legacy mma.sync path: the GEMM is 0.77 GFLOP and bound by its 8.5 MB output)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int NPAD = 2112;  // 33 * 64 >= 2086

// grid (ceil(B/64), 33), 128 threads: warp w -> rows [64*bx + 16w, +16), cols [64*by, +64)
__global__ void __launch_bounds__(128) k_policy_fc(const __half *__restrict__ hp /* [Bpad][192] */, int B, const __half *__restrict__ wp /* [NPAD][192] */,
                                                    const float *__restrict__ bp /* [NPAD] */, float *__restrict__ logits /* [B][2086] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int row0 = blockIdx.x * 64 + warp * 16, col0 = blockIdx.y * 64;
    float acc[8][4];
#pragma unroll
    for (int n = 0; n < 8; n++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[n][k] = 0.f;
    const int ra = min(row0 + g, B - 1), rb = min(row0 + g + 8, B - 1);   // clamp: rows >= B are computed but never stored
    const uint32_t *A0 = reinterpret_cast<const uint32_t *>(hp + (size_t)ra * 192);
    const uint32_t *A1 = reinterpret_cast<const uint32_t *>(hp + (size_t)rb * 192);
#pragma unroll 4
    for (int ks = 0; ks < 12; ks++) {
        uint32_t a[4];
        a[0] = A0[ks * 8 + t];
        a[1] = A1[ks * 8 + t];
        a[2] = A0[ks * 8 + 4 + t];
        a[3] = A1[ks * 8 + 4 + t];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const uint32_t *Bp = reinterpret_cast<const uint32_t *>(wp + (size_t)(col0 + n * 8 + g) * 192);
            uint32_t b[2];
            b[0] = Bp[ks * 8 + t];
            b[1] = Bp[ks * 8 + 4 + t];
            mma16816(acc[n], a, b);
        }
    }
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const int col = col0 + n * 8 + t * 2;
        if (col >= CZ_NLABEL) continue;
        const float b0 = bp[col], b1 = bp[col + 1];
        if (row0 + g < B) *reinterpret_cast<float2 *>(logits + (size_t)(row0 + g) * CZ_NLABEL + col) = make_float2(acc[n][0] + b0, acc[n][1] + b1);
        if (row0 + g + 8 < B) *reinterpret_cast<float2 *>(logits + (size_t)(row0 + g + 8) * CZ_NLABEL + col) = make_float2(acc[n][2] + b0, acc[n][3] + b1);
    }
}

}  // namespace

extern "C" {

int cz_net_first_conv(const uint8_t *canon_boards, int B, const void *w1, const float *b1, void *out, void *stream) {
    if (!canon_boards || !w1 || !b1 || !out || B <= 0) return CZ_EINVAL;
    k_first_conv<<<(B + FC_POS - 1) / FC_POS, FC_THREADS, 0, (cudaStream_t)stream>>>(canon_boards, B, reinterpret_cast<const __half2 *>(w1),
                                                                                      reinterpret_cast<const float2 *>(b1), reinterpret_cast<__half2 *>(out));
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

int cz_net_heads(const void *x, int B, const float *wh, const float *bh, const float *w1t, const float *b1, const float *w2, float b2,
                 const void *wp, const float *bp, void *hp_scratch, float *logits, float *value, void *stream) {
    if (!x || !wh || !bh || !w1t || !b1 || !w2 || !wp || !bp || !hp_scratch || !logits || !value || B <= 0) return CZ_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    k_head_conv<<<(B + HC_POS - 1) / HC_POS, HC_THREADS, 0, st>>>((const __half *)x, B, wh, bh, w1t, b1, w2, b2, (__half *)hp_scratch, value);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    dim3 grid((B + 63) / 64, NPAD / 64);
    k_policy_fc<<<grid, 128, 0, st>>>((const __half *)hp_scratch, B, (const __half *)wp, bp, logits);
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

}  // extern "C"
