// cz_net.cu -- hand-written sm_100a kernels for the two ends of the policy-value network
// (policy_value_network.py:45-74): the first convolution evaluated straight from board bytes, and the
// fused policy / value heads.  The batch-1024 residual tower in between stays on the library tcgen05 path.
//
//   k_first_conv      : canonical board bytes -> conv3x3(14->128)+bias+ReLU output, fp16 NHWC [B][90][128].
//                       The 14-plane input is one-hot and <= 32 of its 1260 cells are set, so the convolution is a
//                       gather-add of weight rows: out[cell][:] = b + sum over the 3x3 neighbourhood of W[tap][piece][:].
//                       The [9][10][14] tensor (and the reference's rank*9+file indexing, main.py:550-555) is never
//                       materialised: image cell (r, f) reads canonical board byte r*9+f.
//   k_first_conv_tc   : the same layer on tcgen05 + TMEM (one-hot im2col operand built in shared memory).
//   k_head_conv_mma   : conv1x1(128->3)+bias+ReLU (policy 2 ch + value 1 ch) as one streaming pass on mma.sync (hi+lo fp16
//                       weight split: fp32-weight accuracy); writes the policy features hp either row-major fp16 [B][192]
//                       (flatten order (h, w, c), zero padded) or directly as tcgen05 operand tiles.  k_head_conv: round-1 SIMT form.
//   k_value_mlp       : 90 -> 256 ReLU -> 1 tanh, 8 positions per CTA, all 90 weights of a hidden unit requested up front.
//   k_policy_fc_tc    : logits[B][2086] = hp . Wp^T + bp on tcgen05 + TMEM; both operands arrive as 48 KB bulk async copies
//                       (cp.async.bulk) already in the UMMA layout; fp32 logits stored 128 B per warp and position.
//   k_policy_fc       : the same on mma.sync m16n8k16 (small batches, row-major hp).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cchess_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// first convolution from board bytes: one CTA per position, a half-warp per output cell (128 channels, 8 per lane)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_first_conv(const uint8_t *__restrict__ boards, int B, const __half *__restrict__ w /* [9][14][128] */,
                                                     const float4 *__restrict__ bias /* [32] */, __half *__restrict__ out /* [B][90][128] */) {
    __shared__ uint8_t pb[11 * 12 + 12];   // zero-bordered image: pb[(r+1)*12 + f+1] = canonical board byte r*9+f (r < 9, f < 10)
    const int pos = blockIdx.x;
    const uint8_t *bd = boards + (size_t)pos * 96;
    if (threadIdx.x < 132) {
        const int r = threadIdx.x / 12 - 1, f = threadIdx.x % 12 - 1;
        pb[threadIdx.x] = (r >= 0 && r < 9 && f >= 0 && f < 10) ? bd[r * 9 + f] : (uint8_t)0;   // the reference's cell <- s[rank*9+file]
    }
    __syncthreads();
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;   // 16 half-warps, lane owns channels 8*l16 .. 8*l16+7
    const float4 b0 = bias[l16 * 2], b1 = bias[l16 * 2 + 1];
    for (int cell = hw; cell < 90; cell += 16) {
        const int r = cell / 10, f = cell - r * 10;
        const uint8_t *c0 = pb + r * 12 + f;          // top-left of the 3x3 window
        int pc[9];
#pragma unroll
        for (int t = 0; t < 9; t++) pc[t] = c0[(t / 3) * 12 + (t % 3)];
        float acc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int t = 0; t < 9; t++) {
            if (pc[t]) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(w + ((size_t)(t * 14 + pc[t] - 1) * 128 + l16 * 8)));
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 v = __half22float2(h2[k]);
                    acc[2 * k] += v.x;
                    acc[2 * k + 1] += v.y;
                }
            }
        }
        uint4 o;
        __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
        for (int k = 0; k < 4; k++) oh[k] = __floats2half2_rn(fmaxf(acc[2 * k], 0.f), fmaxf(acc[2 * k + 1], 0.f));
        *reinterpret_cast<uint4 *>(out + ((size_t)pos * 90 + cell) * 128 + l16 * 8) = o;
    }
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------------------
// first convolution on mma.sync with the ONE-HOT operand built in registers (alternative; measured 20.4 us vs 12.8 us for the gather-add
// at 1024 positions: each warp streams the whole 36 KB of weight fragments from shared memory per 16-cell tile -- not adopted).
//   D[16 cells][128 ch] = A[16][144] . B[144][128],  K = 9 taps x 16 piece slots (slot 0 = empty -> zero weight row; slot 15 of the
//   centre tap is a constant 1 against the bias row, as in the tcgen05 variant).  A never exists anywhere: lane (g, t) of the warp
//   derives its m16n8k16 fragment words for tap `tap` from the two piece codes of its rows g and g + 8 -- a 1.0 in the half that
//   matches the code, zero otherwise (~10 ALU instructions per tap).  B is the weight matrix pre-arranged on the host in FRAGMENT
//   order [k-step 9][n-tile 16][lane 32][2 words], copied once per CTA into shared memory: one conflict-free LDS.64 per MMA.
//   A CTA (8 warps) handles 4 positions = 24 row tiles, 3 per warp.
// ------------------------------------------------------------------------------------------
constexpr int FCM_POS = 4;
__global__ void __launch_bounds__(256) k_first_conv_mma(const uint8_t *__restrict__ boards, int B, const uint2 *__restrict__ wfrag /* [9][16][32] */,
                                                         __half *__restrict__ out /* [B][90][128] */) {
    __shared__ uint2 sW[9 * 16 * 32];                          // 36 864 B
    __shared__ uint8_t pb[FCM_POS][11 * 12 + 12];              // zero-bordered images: pb[(r+1)*12 + f+1] = canonical board byte r*9+f (r < 9, f < 10)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int pos0 = blockIdx.x * FCM_POS;
    for (int i = tid; i < 9 * 16 * 32 / 2; i += 256) reinterpret_cast<uint4 *>(sW)[i] = __ldg(reinterpret_cast<const uint4 *>(wfrag) + i);
    for (int i = tid; i < FCM_POS * 144; i += 256) {
        const int p = i / 144, j = i - p * 144;
        uint8_t v = 0;
        if (j < 132 && pos0 + p < B) {
            const int r = j / 12 - 1, f = j % 12 - 1;
            if (r >= 0 && r < 9 && f >= 0 && f < 10) v = boards[(size_t)(pos0 + p) * 96 + r * 9 + f];   // the reference's cell <- s[rank*9+file]
        }
        pb[p][j] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int tile = warp; tile < FCM_POS * 6; tile += 8) {
        const int p = tile / 6, mt = tile - p * 6;
        if (pos0 + p >= B) break;
        const int c0 = mt * 16 + g, c1 = c0 + 8;               // this lane's two rows (cells); rows >= 90 are padding
        const int r0 = c0 / 10, f0 = c0 - r0 * 10, r1 = c1 / 10, f1 = c1 - r1 * 10;
        const uint8_t *q0 = pb[p] + r0 * 12 + f0, *q1 = pb[p] + r1 * 12 + f1;   // top-left of the 3x3 windows
        float acc[16][4];
#pragma unroll
        for (int nt = 0; nt < 16; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
        for (int tap = 0; tap < 9; tap++) {
            const int off = (tap / 3) * 12 + (tap % 3);
            const int pc0 = c0 < 90 ? q0[off] : 0, pc1 = c1 < 90 ? q1[off] : 0;
            uint32_t a[4];
            const uint32_t one0 = 0x3C00u << ((pc0 & 1) * 16), one1 = 0x3C00u << ((pc1 & 1) * 16);
            a[0] = (pc0 >> 1) == t ? one0 : 0u;                // k = 2t, 2t+1   <-> codes 0..7
            a[1] = (pc1 >> 1) == t ? one1 : 0u;
            a[2] = (pc0 >> 1) == t + 4 ? one0 : 0u;            // k = 2t+8, 2t+9 <-> codes 8..15
            a[3] = (pc1 >> 1) == t + 4 ? one1 : 0u;
            if (tap == 4 && t == 3) { if (c0 < 90) a[2] |= 0x3C000000u; if (c1 < 90) a[3] |= 0x3C000000u; }   // slot 15 of the centre tap: 1 x bias row
            const uint2 *wk = sW + tap * 16 * 32 + lane;
#pragma unroll
            for (int nt = 0; nt < 16; nt++) {
                const uint2 bw = wk[nt * 32];
                const uint32_t b[2] = {bw.x, bw.y};
                mma16816(acc[nt], a, b);
            }
        }
        __half *o0 = out + ((size_t)(pos0 + p) * 90 + c0) * 128 + t * 2, *o1 = out + ((size_t)(pos0 + p) * 90 + c1) * 128 + t * 2;
#pragma unroll
        for (int nt = 0; nt < 16; nt++) {
            if (c0 < 90) *reinterpret_cast<__half2 *>(o0 + nt * 8) = __floats2half2_rn(fmaxf(acc[nt][0], 0.f), fmaxf(acc[nt][1], 0.f));
            if (c1 < 90) *reinterpret_cast<__half2 *>(o1 + nt * 8) = __floats2half2_rn(fmaxf(acc[nt][2], 0.f), fmaxf(acc[nt][3], 0.f));
        }
    }
}

// ------------------------------------------------------------------------------------------
// first convolution on the 5th-generation tensor cores (tcgen05 + TMEM), one 128-cell tile per CTA.
//   D[128 cells][128 ch] (f32, TMEM) = A[128][144] . B[144][128],  K = 9 taps x 16 "piece slots"
//   A is ONE-HOT and never exists in global memory: thread r builds row r (its cell's 3x3 neighbourhood, one 16-wide
//   slot per tap with a 1.0 at the piece code; code 0 = empty / off-board hits an all-zero weight row) straight into
//   shared memory in the canonical K-major no-swizzle UMMA layout; B (the folded conv weights, same layout, prepared once
//   on the host) is copied from L2.  One elected thread issues 9 tcgen05.mma (M128 N128 K16, kind::f16, f32 accumulate),
//   commits to an mbarrier; the four warps read their TMEM lane quarter back with tcgen05.ld, add bias, ReLU, store fp16.
// Shared-memory operand layout (both A and B): [k-chunk of 8 halves (18)][8-row group (16)][row in group (8)][8 halves]
//   -> core matrix = 128 contiguous bytes, SBO (next 8-row group) = 128 B, LBO (next k-chunk) = 2048 B.
// ------------------------------------------------------------------------------------------
constexpr int TC_TILE_BYTES = 18 * 16 * 128;   // 36 864 B per operand

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t smem_addr) {
    // cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 (no swizzle) [61,64)
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

// Persistent: each CTA copies B and allocates TMEM once, then walks tiles blockIdx.x, +gridDim.x, ... (three CTAs per SM so that one
// CTA's epilogue overlaps the others' operand build).  The bias rides in the GEMM: slot 15 of the centre tap is a constant 1 in A
// and the bias row in B, so the epilogue is ReLU + fp16 convert only.
__global__ void __launch_bounds__(128) k_first_conv_tc(const uint8_t *__restrict__ boards, int B, const uint4 *__restrict__ wB /* TC_TILE_BYTES */,
                                                        __half *__restrict__ out /* [B][90][128] */) {
    extern __shared__ __align__(128) unsigned char smem_tc[];
    unsigned char *sA = smem_tc, *sB = smem_tc + TC_TILE_BYTES;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long total = (long long)B * 90;
    const long long tiles = (total + 127) / 128;

    if (warp == 0) {   // TMEM: 128 columns x 128 lanes of f32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < TC_TILE_BYTES / 16; i += 128) reinterpret_cast<uint4 *>(sB)[i] = __ldg(wB + i);   // B operand, once
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b F16 K-major, N=128 (16<<17), M=128 (8<<24)
    const uint32_t idesc = (1u << 4) | (16u << 17) | (8u << 24);
    const uint64_t da = umma_desc_kmajor_noswizzle(smem_u32(sA)), db = umma_desc_kmajor_noswizzle(smem_u32(sB));
    unsigned char *rowp = sA + (tid >> 3) * 128 + (tid & 7) * 16;
    uint32_t phase = 0;

    // piece codes of the 3x3 neighbourhood of cell c (0 = empty / outside the board / beyond the batch)
    auto load_codes = [&](long long c, int (&pc)[9]) {
#pragma unroll
        for (int t = 0; t < 9; t++) pc[t] = 0;
        if (c < total) {
            const int pos = (int)(c / 90), cell = (int)(c - (long long)pos * 90);
            const int r = cell / 10, f = cell - r * 10;
            const uint8_t *bd = boards + (size_t)pos * 96;
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int rr = r + t / 3 - 1, ff = f + t % 3 - 1;
                if (rr >= 0 && rr < 9 && ff >= 0 && ff < 10) pc[t] = __ldg(bd + rr * 9 + ff);   // the reference's cell <- s[rank*9+file]
            }
        }
    };
    int pc[9], pcn[9];
    load_codes((long long)blockIdx.x * 128 + tid, pc);

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        // ---- A operand: one-hot row of this thread's cell (row `tid` of the tile) ----
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const uint32_t one = 0x3C00u << ((pc[t] & 1) * 16);      // fp16 1.0 in the low or high half of a word
            const int w = pc[t] >> 1;                                  // word 0..7 inside the 16-wide slot
            uint4 lo, hi;
            lo.x = w == 0 ? one : 0u; lo.y = w == 1 ? one : 0u; lo.z = w == 2 ? one : 0u; lo.w = w == 3 ? one : 0u;
            hi.x = w == 4 ? one : 0u; hi.y = w == 5 ? one : 0u; hi.z = w == 6 ? one : 0u; hi.w = w == 7 ? one : 0u;
            if (t == 4) hi.w |= 0x3C000000u;                           // slot 15 of the centre tap: constant 1 -> bias row of B
            *reinterpret_cast<uint4 *>(rowp + (2 * t) * 2048) = lo;
            *reinterpret_cast<uint4 *>(rowp + (2 * t + 1) * 2048) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy smem writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const uint64_t a = da + (uint64_t)((t * 4096) >> 4), b = db + (uint64_t)((t * 4096) >> 4);   // two k-chunks per MMA
                const uint32_t acc = t > 0 ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        load_codes((tile + gridDim.x) * 128 + tid, pcn);             // next tile's board bytes fly while the tensor core works
        {   // wait for this tile's MMAs
            const uint32_t bar = smem_u32(&mbar);
            asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                         ::"r"(bar), "r"(phase) : "memory");
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 = tile rows; thread = one row, 4 x 32 columns.  The fp16 row is
        // staged in shared memory (the A tile is dead once the mbarrier fired) with an XOR swizzle on the 16-byte chunk index,
        // then written with fully coalesced 512-byte warp stores (two rows per instruction). ----
        unsigned char *stage = sA + warp * (32 * 256);                 // this warp's 32 rows x 256 B
        const int lane = tid & 31;
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint4 o;
                __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    oh[k] = __floats2half2_rn(fmaxf(__uint_as_float(v[j * 8 + k * 2]), 0.f), fmaxf(__uint_as_float(v[j * 8 + k * 2 + 1]), 0.f));
                const int chunk = q * 4 + j;
                *reinterpret_cast<uint4 *>(stage + lane * 256 + ((chunk ^ (lane & 15)) << 4)) = o;
            }
        }
        __syncwarp();
        {
            const long long row0 = tile * 128 + warp * 32;             // first cell of this warp's 32 rows
            const int half = lane >> 4, ch = lane & 15;
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int r = 2 * i + half;
                if (row0 + r < total) {
                    const uint4 o = *reinterpret_cast<const uint4 *>(stage + r * 256 + ((ch ^ (r & 15)) << 4));
                    *reinterpret_cast<uint4 *>(out + (size_t)(row0 + r) * 128 + ch * 8) = o;
                }
            }
        }
        // the next tile overwrites A (its MMAs are complete: mbarrier) and the accumulators (every warp must have drained its lanes)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 9; t++) pc[t] = pcn[t];
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

// ------------------------------------------------------------------------------------------
// heads, stage 1: conv1x1 (128 -> 3) + bias + ReLU.  One CTA per position; all loads of a warp are issued first.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_head_conv(const __half *__restrict__ x /* [B][90][128] */, int B, const float *__restrict__ wh /* [3][128] */,
                                                    const float *__restrict__ bh /* [3] */, __half *__restrict__ hp /* [B][192] */,
                                                    float *__restrict__ hv /* [B][96] */) {
    const int pos = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half_id = lane >> 4, l16 = lane & 15;   // two cells per warp iteration, 16 lanes x 8 channels each
    float wreg[3][8];
#pragma unroll
    for (int o = 0; o < 3; o++)
#pragma unroll
        for (int k = 0; k < 8; k++) wreg[o][k] = __ldg(wh + o * 128 + l16 * 8 + k);
    const float bh0 = bh[0], bh1 = bh[1], bh2 = bh[2];
    uint4 raw[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {                     // cells: (warp + 8j)*2 + half_id, 45 pairs over 8 warps
        const int cell = (warp + 8 * j) * 2 + half_id;
        raw[j] = cell < 90 ? __ldg(reinterpret_cast<const uint4 *>(x + ((size_t)pos * 90 + cell) * 128 + l16 * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const int cell = (warp + 8 * j) * 2 + half_id;
        const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw[j]);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 v = __half22float2(h2[k]);
            s0 += v.x * wreg[0][2 * k] + v.y * wreg[0][2 * k + 1];
            s1 += v.x * wreg[1][2 * k] + v.y * wreg[1][2 * k + 1];
            s2 += v.x * wreg[2][2 * k] + v.y * wreg[2][2 * k + 1];
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (l16 == 0 && cell < 90) {
            // flatten order of tf.reshape on NHWC (policy_value_network.py:62, 72): index = cell*2 + c
            *reinterpret_cast<__half2 *>(hp + (size_t)pos * 192 + cell * 2) = __floats2half2_rn(fmaxf(s0 + bh0, 0.f), fmaxf(s1 + bh1, 0.f));
            hv[(size_t)pos * 96 + cell] = fmaxf(s2 + bh2, 0.f);
        }
    }
    if (threadIdx.x < 12) hp[(size_t)pos * 192 + 180 + threadIdx.x] = __float2half(0.f);   // K padding of the policy GEMM
}

// heads, stage 2a: value MLP 90 -> 256 ReLU -> 1 tanh (policy_value_network.py:73-74), 8 positions per CTA.
// Thread t owns hidden unit t: its 90 first-layer weights are requested up front (90 independent coalesced loads, one L2 round trip
// instead of nine), the positions' features are broadcast from shared memory.
constexpr int VM_POS = 8;
__global__ void __launch_bounds__(256) k_value_mlp(const float *__restrict__ hv /* [B][96] */, int B, const float *__restrict__ w1t /* [90][256] */,
                                                    const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2, float *__restrict__ value) {
    __shared__ float sh[VM_POS][96];
    __shared__ float red[VM_POS][8];
    const int p0 = blockIdx.x * VM_POS, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    float wv[90];
#pragma unroll
    for (int k = 0; k < 90; k++) wv[k] = __ldg(w1t + k * 256 + t);
    for (int i = t; i < VM_POS * 96; i += 256) {
        const int p = i / 96;
        sh[p][i - p * 96] = p0 + p < B ? hv[(size_t)(p0 + p) * 96 + (i - p * 96)] : 0.f;
    }
    const float bb = b1[t], w2v = w2[t];
    __syncthreads();
    float a[VM_POS];
#pragma unroll
    for (int p = 0; p < VM_POS; p++) a[p] = bb;
#pragma unroll
    for (int k = 0; k < 90; k++)
#pragma unroll
        for (int p = 0; p < VM_POS; p++) a[p] += wv[k] * sh[p][k];
#pragma unroll
    for (int p = 0; p < VM_POS; p++) {
        float s = fmaxf(a[p], 0.f) * w2v;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[p][warp] = s;
    }
    __syncthreads();
    if (t < VM_POS && p0 + t < B) {
        float s = __ldg(b2);
#pragma unroll
        for (int wq = 0; wq < 8; wq++) s += red[t][wq];
        value[p0 + t] = tanhf(s);
    }
}

// ------------------------------------------------------------------------------------------
// heads, stage 2b: policy FC on tensor cores (legacy mma.sync path: 0.77 GFLOP, bound by its 8.5 MB f32 output)
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// heads, stage 1 on the (legacy) tensor path: the same conv1x1 (128 -> 3) as one streaming pass.
// k_head_conv above spends 5.5 M warp instructions on shuffles and conversions to read 23.6 MB (13 us, 22 % of DRAM peak).  Here a CTA
// (one position, 6 warps) stages its 90 x 128 fp16 cells in shared memory with cp.async (rows padded to 272 B: conflict-free
// ldmatrix), each warp multiplies its 16 cells with mma.sync m16n8k16 against the head weights held in registers as B fragments
// (N = 8: policy 2 + value 1 + 5 zero columns).  The f32 weights enter as hi + lo fp16 pairs (two MMAs per k-step), so the result
// equals the fp32-weight dot product to ~1e-7 relative: no precision is traded for the speed.  ~2 instructions per cell.
// ------------------------------------------------------------------------------------------
constexpr int HC_ROW = 136;   // halves per staged row (128 + 8 pad = 272 B)
// TILED: hp is written in the UMMA operand layout k_policy_fc_tc consumes: [position tile of 128][k-chunk 24][128 positions][8 halves].
template <bool TILED>
__global__ void __launch_bounds__(192) k_head_conv_mma(const __half *__restrict__ x /* [B][90][128] */, int B, const float *__restrict__ wh /* [3][128] */,
                                                        const float *__restrict__ bh /* [3] */, __half *__restrict__ hp /* [B][192] or tiled */,
                                                        float *__restrict__ hv /* [B][96] */) {
    __shared__ __align__(16) __half sx[96 * HC_ROW];
    const int pos = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const __half *src = x + (size_t)pos * 90 * 128;
    for (int i = tid; i < 96 * 16; i += 192) {                   // 16-byte chunks: row i / 16, chunk i % 16
        const int r = i >> 4, c = i & 15;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sx + r * HC_ROW + c * 8);
        if (r < 90) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + r * 128 + c * 8) : "memory");
        else *reinterpret_cast<uint4 *>(sx + r * HC_ROW + c * 8) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // B fragments (col-major K x N): lane (g = lane >> 2 -> output n, t = lane & 3) holds k = 16*ks + 2t, 2t+1 and + 8
    const int g = lane >> 2, t = lane & 3;
    uint32_t bhi[8][2], blo[8][2];
#pragma unroll
    for (int ks = 0; ks < 8; ks++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = ks * 16 + h * 8 + t * 2;
            const float w0 = g < 3 ? __ldg(wh + g * 128 + k) : 0.f, w1 = g < 3 ? __ldg(wh + g * 128 + k + 1) : 0.f;
            const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
            const __half l0 = __float2half_rn(w0 - __half2float(h0)), l1 = __float2half_rn(w1 - __half2float(h1));
            __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
            bhi[ks][h] = *reinterpret_cast<uint32_t *>(&hh);
            blo[ks][h] = *reinterpret_cast<uint32_t *>(&ll);
        }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int row0 = warp * 16;
    // ldmatrix.x4: lanes 0-15 address rows row0 + (lane & 15) at k-offset 0, lanes 16-31 the same rows at k-offset 8
    const uint32_t abase = (uint32_t)__cvta_generic_to_shared(sx + (row0 + (lane & 15)) * HC_ROW + (lane >> 4) * 8);
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        uint32_t a[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(abase + ks * 32));
        mma16816(acc, a, bhi[ks]);
        mma16816(acc, a, blo[ks]);
    }
    // accumulator layout: rows row0 + g and row0 + g + 8, columns 2t, 2t + 1
    const float b0 = bh[0], b1 = bh[1], b2 = bh[2];
#pragma unroll
    for (int hlf = 0; hlf < 2; hlf++) {
        const int cell = row0 + g + hlf * 8;
        if (cell < 90) {
            // flatten order of tf.reshape on NHWC (policy_value_network.py:62, 72): index = cell*2 + c
            if (t == 0) {
                const __half2 v = __floats2half2_rn(fmaxf(acc[hlf * 2] + b0, 0.f), fmaxf(acc[hlf * 2 + 1] + b1, 0.f));
                const int k = cell * 2;                              // feature index (even): chunk k / 8, offset k % 8
                if (TILED) *reinterpret_cast<__half2 *>(hp + ((size_t)(pos >> 7) * 24 + (k >> 3)) * 1024 + (size_t)(pos & 127) * 8 + (k & 7)) = v;
                else *reinterpret_cast<__half2 *>(hp + (size_t)pos * 192 + k) = v;
            }
            if (t == 1) hv[(size_t)pos * 96 + cell] = fmaxf(acc[hlf * 2] + b2, 0.f);
        }
    }
    if (tid < 12) {                                                  // K padding of the policy GEMM (features 180..191)
        const int k = 180 + tid;
        if (TILED) hp[((size_t)(pos >> 7) * 24 + (k >> 3)) * 1024 + (size_t)(pos & 127) * 8 + (k & 7)] = __float2half(0.f);
        else hp[(size_t)pos * 192 + k] = __float2half(0.f);
    }
}

// ------------------------------------------------------------------------------------------
// heads, stage 2b on the 5th-generation tensor cores: logits[pos][label] = hp[pos] . wp[label] + bp[label]   (tcgen05 + TMEM)
//   D[128 labels][128 positions] (f32, TMEM) = A[128][192] . B[192][128]: A = a 128-label tile of the FC weights, B = a 128-position
//   tile of the head features, both already in the canonical K-major no-swizzle UMMA layout in global memory (weights: prepared once
//   on the host; features: written that way by k_head_conv_mma<true>), so each operand is ONE 48 KB bulk async copy
//   (cp.async.bulk, SASS UBLKCP) signalling an mbarrier.  One elected thread issues 12 tcgen05.mma (M128 N128 K16), commits; the four
//   warps read their TMEM lane quarter (lane = label) 32 positions at a time and store logits[pos][label0 .. label0+31] -- 128
//   contiguous bytes per warp and position.  136 CTAs for 1024 positions x 2086 labels, one tile each.
// ------------------------------------------------------------------------------------------
constexpr int FC_TILE_BYTES = 24 * 128 * 16;   // 49 152 B per operand tile
__global__ void __launch_bounds__(128) k_policy_fc_tc(const uint4 *__restrict__ hp_tiled, int B, const uint4 *__restrict__ wp_tiled,
                                                       const float *__restrict__ bp /* [>= 2176] */, float *__restrict__ logits /* [B][2086] */) {
    extern __shared__ __align__(128) unsigned char smem_fc_tc[];
    unsigned char *sA = smem_fc_tc, *sB = smem_fc_tc + FC_TILE_BYTES;
    __shared__ __align__(8) uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.y, nt = blockIdx.x;                     // label tile, position tile
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_ld)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_mma)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_ld)), "r"(2u * FC_TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sA)), "l"(reinterpret_cast<const unsigned char *>(wp_tiled) + (size_t)mt * FC_TILE_BYTES), "r"(FC_TILE_BYTES), "r"(smem_u32(&bar_ld)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sB)), "l"(reinterpret_cast<const unsigned char *>(hp_tiled) + (size_t)nt * FC_TILE_BYTES), "r"(FC_TILE_BYTES), "r"(smem_u32(&bar_ld)) : "memory");
        {   // operands landed
            const uint32_t bar = smem_u32(&bar_ld);
            asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                         ::"r"(bar), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (16u << 17) | (8u << 24);          // f32 accumulate, f16 x f16, N = 128, M = 128
        const uint64_t da = umma_desc_kmajor_noswizzle(smem_u32(sA)), db = umma_desc_kmajor_noswizzle(smem_u32(sB));
#pragma unroll
        for (int ks = 0; ks < 12; ks++) {
            const uint64_t a = da + (uint64_t)((ks * 4096) >> 4), b = db + (uint64_t)((ks * 4096) >> 4);   // two k-chunks per MMA
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    {
        const uint32_t bar = smem_u32(&bar_mma);
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                     ::"r"(bar), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int label = mt * 128 + warp * 32 + lane;
    const float bias = bp[label];
    const bool lab_ok = label < CZ_NLABEL;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * 32);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const int pos = nt * 128 + q * 32 + j;
            if (lab_ok && pos < B) logits[(size_t)pos * CZ_NLABEL + label] = __uint_as_float(v[j]) + bias;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

constexpr int NPAD = 2112;   // 33 * 64 >= 2086
constexpr int LDS_ROW = 200; // halves per staged row (192 + 8 pad): fragment loads hit 32 distinct banks

// grid (ceil(B/64), 33), 128 threads.  A tile (64 positions x 192) and B tile (64 labels x 192) are staged through shared
// memory with 16-byte coalesced loads, all in flight at once; warp w owns rows [16w, 16w+16) x 64 columns.
__global__ void __launch_bounds__(128) k_policy_fc(const __half *__restrict__ hp /* [B][192] */, int B, const __half *__restrict__ wp /* [NPAD][192] */,
                                                    const float *__restrict__ bp /* [NPAD] */, float *__restrict__ logits /* [B][2086] */) {
    extern __shared__ __align__(16) __half smem_fc[];
    __half *sA = smem_fc, *sB = smem_fc + 64 * LDS_ROW;
    const int row0 = blockIdx.x * 64, col0 = blockIdx.y * 64;
    for (int i = threadIdx.x; i < 64 * 24; i += 128) {          // 24 x 16-byte chunks per 192-half row
        const int r = i / 24, c = i - r * 24;
        const int ra = min(row0 + r, B - 1);                     // clamp: rows >= B are computed but never stored
        *reinterpret_cast<uint4 *>(sA + r * LDS_ROW + c * 8) = __ldg(reinterpret_cast<const uint4 *>(hp + (size_t)ra * 192 + c * 8));
        *reinterpret_cast<uint4 *>(sB + r * LDS_ROW + c * 8) = __ldg(reinterpret_cast<const uint4 *>(wp + (size_t)(col0 + r) * 192 + c * 8));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float acc[8][4];
#pragma unroll
    for (int n = 0; n < 8; n++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[n][k] = 0.f;
    const uint32_t *A0 = reinterpret_cast<const uint32_t *>(sA + (warp * 16 + g) * LDS_ROW);
    const uint32_t *A1 = reinterpret_cast<const uint32_t *>(sA + (warp * 16 + g + 8) * LDS_ROW);
#pragma unroll
    for (int ks = 0; ks < 12; ks++) {
        uint32_t a[4];
        a[0] = A0[ks * 8 + t];
        a[1] = A1[ks * 8 + t];
        a[2] = A0[ks * 8 + 4 + t];
        a[3] = A1[ks * 8 + 4 + t];
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const uint32_t *Bp = reinterpret_cast<const uint32_t *>(sB + (n * 8 + g) * LDS_ROW);
            uint32_t b[2];
            b[0] = Bp[ks * 8 + t];
            b[1] = Bp[ks * 8 + 4 + t];
            mma16816(acc[n], a, b);
        }
    }
    const int r0 = row0 + warp * 16 + g;
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const int col = col0 + n * 8 + t * 2;
        if (col >= CZ_NLABEL) continue;
        const float b0 = bp[col], b1 = bp[col + 1];
        if (r0 < B) *reinterpret_cast<float2 *>(logits + (size_t)r0 * CZ_NLABEL + col) = make_float2(acc[n][0] + b0, acc[n][1] + b1);
        if (r0 + 8 < B) *reinterpret_cast<float2 *>(logits + (size_t)(r0 + 8) * CZ_NLABEL + col) = make_float2(acc[n][2] + b0, acc[n][3] + b1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Operand split for the fp32-accurate inference mode (net.py: SplitTf32Plan).  y f32 [P][128] (NHWC activations of one convolution)
//   -> hi f32 [P][128] = tf32(y): rounded to the 10-bit TF32 mantissa (nearest, ties away; the low 13 bits are zero, so whatever
//      conversion the tensor-core kernel applies to it is the identity), and
//   -> x2 fp16 [P][256] = { (y - hi) * 2^11 | hi }: both halves have <= 11 significant bits, i.e. they are EXACT in fp16 (up to
//      fp16's range: |hi| <= 65504, residues below 2^-24 * 2^11 flush).
// A TF32 convolution hi(x) * hi(w) (K = 1152: the only long accumulation chain of full-size terms) plus an fp16 convolution
// x2 * { hi(w) | lo(w) * 2^11 } (the two cross terms, 2^-11 smaller, scaled back by 2^-11 in the epilogue) give the product to
// O(2^-22): dropped are lo*lo and the 11-bit rounding of the two lo operands.  Streaming kernel: 32 B in, 64 B out per thread.
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
#define SPLIT_SCALE 2048.0f

__global__ void __launch_bounds__(256) k_split_tf32(const float4 *__restrict__ y, float4 *__restrict__ hi, uint4 *__restrict__ x2, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(y + 2 * i), b = __ldg(y + 2 * i + 1);
        const float4 ha = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
        const float4 hb = make_float4(tf32_hi(b.x), tf32_hi(b.y), tf32_hi(b.z), tf32_hi(b.w));
        hi[2 * i] = ha; hi[2 * i + 1] = hb;
        uint4 l, h;
        *reinterpret_cast<__half2 *>(&l.x) = __floats2half2_rn((a.x - ha.x) * SPLIT_SCALE, (a.y - ha.y) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&l.y) = __floats2half2_rn((a.z - ha.z) * SPLIT_SCALE, (a.w - ha.w) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&l.z) = __floats2half2_rn((b.x - hb.x) * SPLIT_SCALE, (b.y - hb.y) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&l.w) = __floats2half2_rn((b.z - hb.z) * SPLIT_SCALE, (b.w - hb.w) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&h.x) = __floats2half2_rn(ha.x, ha.y);
        *reinterpret_cast<__half2 *>(&h.y) = __floats2half2_rn(ha.z, ha.w);
        *reinterpret_cast<__half2 *>(&h.z) = __floats2half2_rn(hb.x, hb.y);
        *reinterpret_cast<__half2 *>(&h.w) = __floats2half2_rn(hb.z, hb.w);
        uint4 *o = x2 + (i >> 4) * 32 + (i & 15);          // row of 256 halves = 32 uint4: { lo: 16 | hi: 16 }
        o[0] = l; o[16] = h;
    }
}

// One pass per convolution: v = ReLU(t + 2^-11 s + bias [+ skip]) from the two library convolutions' raw results, then the split of v
// for the NEXT convolution.  t f32 [P][128] (hi*hi), s fp16 [P][128] or null (cross terms, scaled), skip f32 or null; outputs (each
// optional): x f32 (the value itself: next block's skip / the heads' input; may alias skip), hi f32, x2 fp16 [P][256] = { lo 2^11 | hi }.
// Streaming: 32-80 B in, 64-96 B out per thread (8 channels).
__global__ void __launch_bounds__(256) k_epilogue_split(const float4 *__restrict__ t, const uint4 *__restrict__ s, const float4 *__restrict__ bias,
                                                        const float4 *skip, float4 *x, float4 *__restrict__ hi, uint4 *__restrict__ x2, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float4 a = __ldg(t + 2 * i), b = __ldg(t + 2 * i + 1);
        if (s) {
            const uint4 sv = __ldg(s + i);
            const float2 s0 = __half22float2(*reinterpret_cast<const __half2 *>(&sv.x)), s1 = __half22float2(*reinterpret_cast<const __half2 *>(&sv.y));
            const float2 s2 = __half22float2(*reinterpret_cast<const __half2 *>(&sv.z)), s3 = __half22float2(*reinterpret_cast<const __half2 *>(&sv.w));
            const float r = 1.0f / SPLIT_SCALE;
            a.x += s0.x * r; a.y += s0.y * r; a.z += s1.x * r; a.w += s1.y * r;
            b.x += s2.x * r; b.y += s2.y * r; b.z += s3.x * r; b.w += s3.y * r;
        }
        const float4 ba = __ldg(bias + 2 * (i & 15)), bb = __ldg(bias + 2 * (i & 15) + 1);
        a.x += ba.x; a.y += ba.y; a.z += ba.z; a.w += ba.w;
        b.x += bb.x; b.y += bb.y; b.z += bb.z; b.w += bb.w;
        if (skip) {
            const float4 ka = skip[2 * i], kb = skip[2 * i + 1];
            a.x += ka.x; a.y += ka.y; a.z += ka.z; a.w += ka.w;
            b.x += kb.x; b.y += kb.y; b.z += kb.z; b.w += kb.w;
        }
        a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
        b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
        if (x) { x[2 * i] = a; x[2 * i + 1] = b; }
        if (!hi) continue;
        const float4 ha = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
        const float4 hb = make_float4(tf32_hi(b.x), tf32_hi(b.y), tf32_hi(b.z), tf32_hi(b.w));
        hi[2 * i] = ha; hi[2 * i + 1] = hb;
        uint4 l, h;
        *reinterpret_cast<__half2 *>(&l.x) = __floats2half2_rn((a.x - ha.x) * SPLIT_SCALE, (a.y - ha.y) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&l.y) = __floats2half2_rn((a.z - ha.z) * SPLIT_SCALE, (a.w - ha.w) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&l.z) = __floats2half2_rn((b.x - hb.x) * SPLIT_SCALE, (b.y - hb.y) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&l.w) = __floats2half2_rn((b.z - hb.z) * SPLIT_SCALE, (b.w - hb.w) * SPLIT_SCALE);
        *reinterpret_cast<__half2 *>(&h.x) = __floats2half2_rn(ha.x, ha.y);
        *reinterpret_cast<__half2 *>(&h.y) = __floats2half2_rn(ha.z, ha.w);
        *reinterpret_cast<__half2 *>(&h.z) = __floats2half2_rn(hb.x, hb.y);
        *reinterpret_cast<__half2 *>(&h.w) = __floats2half2_rn(hb.z, hb.w);
        uint4 *o = x2 + (i >> 4) * 32 + (i & 15);
        o[0] = l; o[16] = h;
    }
}

}  // namespace

extern "C" {

int cz_net_first_conv(const uint8_t *canon_boards, int B, const void *w1, const float *b1, void *out, void *stream) {
    if (!canon_boards || !w1 || !b1 || !out || B <= 0) return CZ_EINVAL;
    k_first_conv<<<B, 256, 0, (cudaStream_t)stream>>>(canon_boards, B, reinterpret_cast<const __half *>(w1), reinterpret_cast<const float4 *>(b1),
                                                      reinterpret_cast<__half *>(out));
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

int cz_net_first_conv_mma(const uint8_t *canon_boards, int B, const void *w_frag, void *out, void *stream) {
    if (!canon_boards || !w_frag || !out || B <= 0) return CZ_EINVAL;
    k_first_conv_mma<<<(B + FCM_POS - 1) / FCM_POS, 256, 0, (cudaStream_t)stream>>>(canon_boards, B, reinterpret_cast<const uint2 *>(w_frag),
                                                                                     reinterpret_cast<__half *>(out));
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

int cz_net_first_conv_tc(const uint8_t *canon_boards, int B, const void *w_umma, const float *b1, void *out, void *stream) {
    if (!canon_boards || !w_umma || !b1 || !out || B <= 0) return CZ_EINVAL;
    const int smem = 2 * TC_TILE_BYTES;   // 73 728 B > 48 KB default: opt in
    if (cudaFuncSetAttribute(k_first_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return CZ_ECUDA;
    const long long tiles = ((long long)B * 90 + 127) / 128;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = tiles < 3LL * sms ? tiles : 3LL * sms;    // persistent: three CTAs per SM (3 x 74 KB smem, 3 x 128 TMEM columns)
    (void)b1;                                                         // the bias is row (centre tap, slot 15) of w_umma
    k_first_conv_tc<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(canon_boards, B, reinterpret_cast<const uint4 *>(w_umma),
                                                                        reinterpret_cast<__half *>(out));
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

// value MLP (side stream) || policy FC on already computed head features hp / hv
int cz_net_heads_fc(const void *hp, const float *hv, int B, const float *w1t, const float *b1, const float *w2, const float *b2,
                    const void *wp, const float *bp, float *logits, float *value, void *stream) {
    if (!hp || !hv || !w1t || !b1 || !w2 || !b2 || !wp || !bp || !logits || !value || B <= 0) return CZ_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 2 * 64 * LDS_ROW * (int)sizeof(__half);   // 51200 B > the 48 KB default: opt in (per device, so every call)
    if (cudaFuncSetAttribute(k_policy_fc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return CZ_ECUDA;
    // The value MLP and the policy FC are independent: fork the value MLP onto a side stream (event fork/join, which
    // CUDA-graph capture records as two parallel branches) so that the two small kernels overlap.
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return CZ_ECUDA;
    static cudaStream_t side[64] = {nullptr};
    static cudaEvent_t ev_fork[64] = {nullptr}, ev_join[64] = {nullptr};
    if (!side[dev]) {   // created on the first (eager, warm-up) call, never during capture
        if (cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking) != cudaSuccess) return CZ_ECUDA;
        if (cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming) != cudaSuccess) return CZ_ECUDA;
        if (cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming) != cudaSuccess) return CZ_ECUDA;
    }
    if (cudaEventRecord(ev_fork[dev], st) != cudaSuccess) return CZ_ECUDA;
    if (cudaStreamWaitEvent(side[dev], ev_fork[dev], 0) != cudaSuccess) return CZ_ECUDA;
    k_value_mlp<<<(B + VM_POS - 1) / VM_POS, 256, 0, side[dev]>>>(hv, B, w1t, b1, w2, b2, value);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    if (cudaEventRecord(ev_join[dev], side[dev]) != cudaSuccess) return CZ_ECUDA;
    dim3 grid((B + 63) / 64, NPAD / 64);
    k_policy_fc<<<grid, 128, smem, st>>>((const __half *)hp, B, (const __half *)wp, bp, logits);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    if (cudaStreamWaitEvent(st, ev_join[dev], 0) != cudaSuccess) return CZ_ECUDA;
    return CZ_OK;
}

int cz_net_heads(const void *x, int B, const float *wh, const float *bh, const float *w1t, const float *b1, const float *w2, const float *b2,
                 const void *wp, const float *bp, void *hp_scratch, float *hv_scratch, float *logits, float *value, void *stream) {
    if (!x || !wh || !bh || !hp_scratch || !hv_scratch || B <= 0) return CZ_EINVAL;
    static const bool legacy = getenv("CCHESS_HEAD_CONV") && !strcmp(getenv("CCHESS_HEAD_CONV"), "simt");
    if (legacy) k_head_conv<<<B, 256, 0, (cudaStream_t)stream>>>((const __half *)x, B, wh, bh, (__half *)hp_scratch, hv_scratch);
    else k_head_conv_mma<false><<<B, 192, 0, (cudaStream_t)stream>>>((const __half *)x, B, wh, bh, (__half *)hp_scratch, hv_scratch);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    return cz_net_heads_fc(hp_scratch, hv_scratch, B, w1t, b1, w2, b2, wp, bp, logits, value, stream);
}


// The heads for large batches: conv1x1 on mma.sync writing the policy features in the UMMA-tiled layout, value MLP on a side stream,
// policy FC on tcgen05 (k_policy_fc_tc).  wp_tiled: dev fp16 [17 label tiles][24 k-chunks][128 labels][8] (labels >= 2086 zero),
// bp: dev f32 [2176]; hp_tiled scratch: fp16, ceil(B/128) * 49152 bytes, ZERO-INITIALISED by the caller (rows beyond B stay zero).
int cz_net_heads_tc(const void *x, int B, const float *wh, const float *bh, const float *w1t, const float *b1, const float *w2, const float *b2,
                    const void *wp_tiled, const float *bp, void *hp_tiled_scratch, float *hv_scratch, float *logits, float *value, void *stream) {
    if (!x || !wh || !bh || !w1t || !b1 || !w2 || !b2 || !wp_tiled || !bp || !hp_tiled_scratch || !hv_scratch || !logits || !value || B <= 0) return CZ_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 2 * FC_TILE_BYTES;
    if (cudaFuncSetAttribute(k_policy_fc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return CZ_ECUDA;
    k_head_conv_mma<true><<<B, 192, 0, st>>>((const __half *)x, B, wh, bh, (__half *)hp_tiled_scratch, hv_scratch);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return CZ_ECUDA;
    static cudaStream_t side[64] = {nullptr};
    static cudaEvent_t ev_fork[64] = {nullptr}, ev_join[64] = {nullptr};
    if (!side[dev]) {   // created on the first (eager, warm-up) call, never during capture
        if (cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking) != cudaSuccess) return CZ_ECUDA;
        if (cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming) != cudaSuccess) return CZ_ECUDA;
        if (cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming) != cudaSuccess) return CZ_ECUDA;
    }
    if (cudaEventRecord(ev_fork[dev], st) != cudaSuccess) return CZ_ECUDA;
    if (cudaStreamWaitEvent(side[dev], ev_fork[dev], 0) != cudaSuccess) return CZ_ECUDA;
    k_value_mlp<<<(B + VM_POS - 1) / VM_POS, 256, 0, side[dev]>>>(hv_scratch, B, w1t, b1, w2, b2, value);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    if (cudaEventRecord(ev_join[dev], side[dev]) != cudaSuccess) return CZ_ECUDA;
    dim3 grid((B + 127) / 128, 17);
    k_policy_fc_tc<<<grid, 128, smem, st>>>((const uint4 *)hp_tiled_scratch, B, (const uint4 *)wp_tiled, bp, logits);
    if (cudaGetLastError() != cudaSuccess) return CZ_ECUDA;
    if (cudaStreamWaitEvent(st, ev_join[dev], 0) != cudaSuccess) return CZ_ECUDA;
    return CZ_OK;
}

// y dev f32 [n_pix][128] -> hi dev f32 [n_pix][128] = tf32(y), x2 dev fp16 [n_pix][256] = { (y - hi) * 2^11 | hi } (see k_split_tf32)
int cz_net_split_tf32(const float *y, float *hi, void *x2, long long n_pix, void *stream) {
    if (!y || !hi || !x2 || n_pix <= 0) return CZ_EINVAL;
    const long long n8 = n_pix * 16;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (n8 + 255) / 256;
    if (blocks > 8LL * sms) blocks = 8LL * sms;          // grid-stride, 8 resident CTAs of 256 threads per SM
    k_split_tf32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(y), reinterpret_cast<float4 *>(hi),
                                                                    reinterpret_cast<uint4 *>(x2), n8);
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

// The f32 epilogue of one three-product convolution fused with the operand split for the next one (k_epilogue_split):
//   v = ReLU(t + 2^-11 s + bias [+ skip]);  x (optional) = v;  hi / x2 (optional, both or neither) = split of v as in cz_net_split_tf32.
// t dev f32 [n_pix][128]; s dev fp16 [n_pix][128] or NULL; bias dev f32 [128]; skip dev f32 [n_pix][128] or NULL (x may alias skip).
int cz_net_epilogue_split(const float *t, const void *s, const float *bias, const float *skip, float *x, float *hi, void *x2, long long n_pix, void *stream) {
    if (!t || !bias || n_pix <= 0 || (!x && !hi) || ((hi == nullptr) != (x2 == nullptr))) return CZ_EINVAL;
    const long long n8 = n_pix * 16;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (n8 + 255) / 256;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    k_epilogue_split<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(t), reinterpret_cast<const uint4 *>(s),
                                                                        reinterpret_cast<const float4 *>(bias), reinterpret_cast<const float4 *>(skip),
                                                                        reinterpret_cast<float4 *>(x), reinterpret_cast<float4 *>(hi),
                                                                        reinterpret_cast<uint4 *>(x2), n8);
    return cudaGetLastError() == cudaSuccess ? CZ_OK : CZ_ECUDA;
}

}  // extern "C"
