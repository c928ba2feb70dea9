// TEST INFRASTRUCTURE.  Compiles the per-lane rule functions of cchess_zero_b200/csrc/cz_rules.cuh FOR THE HOST so that the
// CPU test tier can run the product's own move-generation / encode source against the reference's golden vectors without a GPU.
// Only the warp plumbing (ballots, scans) is replaced by its serial meaning here: squares visited in order, lanes 0..31 in turn.
#include <stdint.h>
#include <string.h>

#include "../cchess_zero_b200/csrc/cz_rules.cuh"

extern "C" int hr_legal_moves(const uint8_t *board, int side, uint16_t *out) {
    int n = 0, Ksq = -1, ksq = -1;
    for (int sq = 0; sq < 90; sq++) {                 // exclusive scan over squares == concatenation in square order
        uint16_t slot[18];
        const int c = cz::gen_piece(board, sq, side, slot);
        for (int j = 0; j < c; j++) out[n++] = slot[j];
        if (board[sq] == 1) Ksq = sq;
        if (board[sq] == 8) ksq = sq;
    }
    if (Ksq >= 0 && ksq >= 0 && (Ksq % 9) == (ksq % 9)) {   // flying general, appended last (same test as warp_legal_moves)
        bool face = true;
        for (int s = Ksq + 9; s < ksq; s += 9)
            if (board[s] != 0) face = false;
        if (face) out[n++] = side == 0 ? (uint16_t)(Ksq | (ksq << 7)) : (uint16_t)(ksq | (Ksq << 7));
    }
    return n;
}

extern "C" void hr_encode_f32(const uint8_t *board, int side, float *out) {
    for (int lane = 0; lane < 32; lane++) cz::warp_encode<float>(board, side, out, lane);
}
