// cz_rules.cuh -- warp-cooperative xiangqi rules for sm_100a: move generation in the
// reference's emission order, move application, flip and the 14-plane encode.
//
// One warp owns one position.  The 90-byte mailbox board sits in shared memory (one
// 96-byte slab per warp); each lane owns squares {lane, lane+32, lane+64}.  Move lists
// are generated once into per-square slots and compacted with an exclusive warp scan over the
// squares so that the output order is exactly the reference's y-major / x-minor piece scan with its per-piece
// direction order (GameBoard.get_legal_moves, main.py:743-1109).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define CZ_FULL 0xffffffffu
// pure per-lane functions are also compiled for the host so that the CPU test tier can run THIS source against the golden vectors
// (tests/host_rules_harness.cu); the qualifier does not change the device code
#define CZ_HD __host__ __device__ __forceinline__

namespace cz {

// piece kinds after folding colour: 1..7 = K A R B N P C  (pieces_order, main.py:208)
enum { K_ = 1, A_ = 2, R_ = 3, B_ = 4, N_ = 5, P_ = 6, C_ = 7 };

CZ_HD bool piece_red(int p) { return p >= 1 && p <= 7; }
CZ_HD int piece_kind(int p) { return p > 7 ? p - 7 : p; }

// Generates the moves of the piece on `sq` for `side` (0 red / 1 black) in reference order into
// out[0..17] (a piece has at most 17 moves) and returns the count.
// Kings are handled here too (palace steps); the flying-general capture is appended by the caller.
#define CZ_OK_TARGET(q) ((q) == 0 || (piece_red(q) != red))            /* validate_move, main.py:727-740 */
#define CZ_EMIT(dst) do { out[n] = (uint16_t)(sq | ((dst) << 7)); n++; } while (0)
CZ_HD int gen_piece(const uint8_t *b, int sq, int side, uint16_t *out) {
    const int p = b[sq];
    if (p == 0) return 0;
    const bool red = piece_red(p);
    if (red != (side == 0)) return 0;
    const int y = sq / 9, x = sq - y * 9;
    const int kind = piece_kind(p);
    int n = 0;
    if (kind == R_ || kind == C_) {
        // rays: left, right, towards y-1, towards y+1 (main.py:757-833 / 947-1062)
        const bool cannon = kind == C_;
#pragma unroll 1
        for (int d = 0; d < 4; d++) {
            const int step = d == 0 ? -1 : d == 1 ? 1 : d == 2 ? -9 : 9;
            const int len = d == 0 ? x : d == 1 ? 8 - x : d == 2 ? y : 9 - y;
            int t = sq;
            bool screen = false;
#pragma unroll 1
            for (int k = 0; k < len; k++) {
                t += step;
                const int q = b[t];
                if (!screen) {
                    if (q == 0) { CZ_EMIT(t); }
                    else if (!cannon) { if (piece_red(q) != red) CZ_EMIT(t); break; }
                    else screen = true;
                } else if (q != 0) {
                    if (piece_red(q) != red) CZ_EMIT(t);
                    break;
                }
            }
        }
    } else if (kind == N_) {
        // i in (-1,+1), j in (-1,+1): (y+2i, x+j) leg (y+i, x); then (y+i, x+2j) leg (y, x+j)  (835-856)
#pragma unroll 1
        for (int ij = 0; ij < 4; ij++) {
            const int i = (ij & 2) ? 1 : -1, j = (ij & 1) ? 1 : -1;
            int ty = y + 2 * i, tx = x + j;
            if (ty >= 0 && ty < 10 && tx >= 0 && tx < 9) {
                const int q = b[ty * 9 + tx];
                if (CZ_OK_TARGET(q) && b[(y + i) * 9 + x] == 0) CZ_EMIT(ty * 9 + tx);
            }
            ty = y + i; tx = x + 2 * j;
            if (ty >= 0 && ty < 10 && tx >= 0 && tx < 9) {
                const int q = b[ty * 9 + tx];
                if (CZ_OK_TARGET(q) && b[y * 9 + x + j] == 0) CZ_EMIT(ty * 9 + tx);
            }
        }
    } else if (kind == B_) {
        // i in (-2,+2): (y+i, x+i) then (y+i, x-i); own half only; eye must be empty (857-888)
#pragma unroll 1
        for (int i = -2; i <= 2; i += 4) {
            const int h = i / 2, ty = y + i;
            if (ty < 0 || ty > 9 || (red ? ty > 4 : ty < 5)) continue;
            int tx = x + i;
            if (tx >= 0 && tx < 9) {
                const int q = b[ty * 9 + tx];
                if (CZ_OK_TARGET(q) && b[(y + h) * 9 + x + h] == 0) CZ_EMIT(ty * 9 + tx);
            }
            tx = x - i;
            if (tx >= 0 && tx < 9) {
                const int q = b[ty * 9 + tx];
                if (CZ_OK_TARGET(q) && b[(y + h) * 9 + x - h] == 0) CZ_EMIT(ty * 9 + tx);
            }
        }
    } else if (kind == A_) {
        // i in (-1,+1): (y+i, x+i) then (y+i, x-i); palace only (889-918)
#pragma unroll 1
        for (int i = -1; i <= 1; i += 2) {
            const int ty = y + i;
            if (ty < 0 || ty > 9 || (red ? ty > 2 : ty < 7)) continue;
            int tx = x + i;
            if (tx >= 3 && tx <= 5) { const int q = b[ty * 9 + tx]; if (CZ_OK_TARGET(q)) CZ_EMIT(ty * 9 + tx); }
            tx = x - i;
            if (tx >= 3 && tx <= 5) { const int q = b[ty * 9 + tx]; if (CZ_OK_TARGET(q)) CZ_EMIT(ty * 9 + tx); }
        }
    } else if (kind == K_) {
        // (y, x-1), (y, x+1), (y-1, x), (y+1, x) inside the own palace (919-946)
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            const int ty = y + (k == 2 ? -1 : k == 3 ? 1 : 0), tx = x + (k == 0 ? -1 : k == 1 ? 1 : 0);
            if (ty < 0 || ty > 9 || tx < 3 || tx > 5 || (red ? ty > 2 : ty < 7)) continue;
            const int q = b[ty * 9 + tx];
            if (CZ_OK_TARGET(q)) CZ_EMIT(ty * 9 + tx);
        }
    } else {  // P_
        // forward (red y+1 / black y-1); after the river x+1 then x-1 (1063-1095)
        const int ty = red ? y + 1 : y - 1;
        if (ty >= 0 && ty < 10) { const int q = b[ty * 9 + x]; if (CZ_OK_TARGET(q)) CZ_EMIT(ty * 9 + x); }
        if (red ? y > 4 : y < 5) {
            if (x < 8) { const int q = b[sq + 1]; if (CZ_OK_TARGET(q)) CZ_EMIT(sq + 1); }
            if (x > 0) { const int q = b[sq - 1]; if (CZ_OK_TARGET(q)) CZ_EMIT(sq - 1); }
        }
    }
    return n;
}
#undef CZ_EMIT
#undef CZ_OK_TARGET

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(CZ_FULL, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(CZ_FULL, inc, 31);
    return inc - v;
}

// Per-warp scratch of the move generator: every square gets an 18-entry slot and a count.
struct MoveScratch {
    uint16_t slot[96 * 18];
    uint8_t cnt[96];
};

// Warp-cooperative GameBoard.get_legal_moves.  b: 90-byte board in shared memory, moves: shared
// uint16[>=136].  One generation pass into per-square slots, then an exclusive warp scan over the
// squares (y-major, x-minor = the reference's scan order) compacts them.  Returns the move count
// (uniform across the warp).  All 32 lanes must call, converged.
__device__ __noinline__ int warp_legal_moves(const uint8_t *b, int side, uint16_t *moves, MoveScratch &T, int lane) {
    int Ksq = -1, ksq = -1;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const int sq = lane + 32 * k;
        int p = 0, c = 0;
        if (sq < 90) {
            p = b[sq];
            c = gen_piece(b, sq, side, T.slot + sq * 18);
        }
        T.cnt[sq] = (uint8_t)c;
        __syncwarp();
        const unsigned Kmask = __ballot_sync(CZ_FULL, p == 1), kmask = __ballot_sync(CZ_FULL, p == 8);
        if (Kmask) Ksq = 32 * k + __ffs(Kmask) - 1;   // at most one king of each colour
        if (kmask) ksq = 32 * k + __ffs(kmask) - 1;
    }
    __syncwarp();
    int n = 0;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const int sq = lane + 32 * k;
        const int c = T.cnt[sq];
        int tot;
        const int off = n + warp_excl_scan(c, lane, tot);
        for (int j = 0; j < c; j++)
            if (off + j < 136) moves[off + j] = T.slot[sq * 18 + j];
        n += tot;
    }
    // flying general: same file, nothing strictly between; the mover's king takes, appended LAST (1097-1107)
    if (Ksq >= 0 && ksq >= 0 && (Ksq % 9) == (ksq % 9)) {
        bool face = true;
        for (int s = Ksq + 9; s < ksq; s += 9)
            if (b[s] != 0) face = false;
        if (face) {
            if (lane == 0 && n < 136) moves[n] = side == 0 ? (uint16_t)(Ksq | (ksq << 7)) : (uint16_t)(ksq | (Ksq << 7));
            n++;
        }
    }
    __syncwarp();
    return n;
}

// swap colour of a piece code (try_flip's swapcase, main.py:566-572)
CZ_HD int swap_colour(int p) { return p == 0 ? 0 : (p <= 7 ? p + 7 : p - 7); }

template <typename T> CZ_HD T enc_one();
template <> CZ_HD float enc_one<float>() { return 1.0f; }
template <> CZ_HD __nv_bfloat16 enc_one<__nv_bfloat16>() { return __float2bfloat16(1.0f); }
template <> CZ_HD __half enc_one<__half>() { return __float2half(1.0f); }

// Warp-cooperative generate_inputs (main.py:531-557): flip for black (rows reversed, colours
// swapped), then T[rank][file][plane] for rank < 9, file < 10 reads board cell rank*9+file --
// the reference's indexing, which drops squares 82..89 and reads 8 cells twice.
// out: 1260 elements of T in global memory (row of the NN batch), 16-byte aligned.
template <typename T>
__host__ __device__ void warp_encode(const uint8_t *b, int side, T *out, int lane) {
    constexpr int VEC = 16 / sizeof(T);           // elements per 16-byte store
    constexpr int NV = 1260 / VEC;                // 315 (f32) or 157.5 -> handled below
    static_assert(1260 % (VEC / 2) == 0, "row must be a multiple of 8 bytes");
    if constexpr (sizeof(T) == 4) {
        for (int v = lane; v < NV; v += 32) {
            float4 o;
            float *of = reinterpret_cast<float *>(&o);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = v * 4 + e, cell = k / 14, plane = k - cell * 14;
                const int rank = cell / 10, file = cell - rank * 10, s = rank * 9 + file;
                int p;
                if (side == 0) p = b[s];
                else { const int yy = s / 9, xx = s - yy * 9; p = swap_colour(b[(9 - yy) * 9 + xx]); }
                of[e] = (p - 1 == plane) ? 1.0f : 0.0f;
            }
            reinterpret_cast<float4 *>(out)[v] = o;
        }
    } else {
        // 2-byte element types: 1260 * 2 = 2520 B = 315 8-byte stores
        for (int v = lane; v < 315; v += 32) {
            T o[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = v * 4 + e, cell = k / 14, plane = k - cell * 14;
                const int rank = cell / 10, file = cell - rank * 10, s = rank * 9 + file;
                int p;
                if (side == 0) p = b[s];
                else { const int yy = s / 9, xx = s - yy * 9; p = swap_colour(b[(9 - yy) * 9 + xx]); }
                o[e] = (p - 1 == plane) ? enc_one<T>() : T(0.0f);
            }
            reinterpret_cast<uint2 *>(out)[v] = *reinterpret_cast<uint2 *>(o);
        }
    }
}

}  // namespace cz
