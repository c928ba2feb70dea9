// TEST INFRASTRUCTURE.  Compiles the per-lane rule functions of cchess_zero_b200/csrc/cz_rules.cuh FOR THE HOST so that the
// CPU test tier can run the product's own move-generation / encode source against the reference's golden vectors without a GPU.
// Only the warp plumbing (ballots that build the bitboards, the piece compaction and the scan) is replaced by its serial meaning here: squares visited in order, lanes 0..31 in turn.
#include <stdint.h>
#include <string.h>

#include "../cchess_zero_b200/csrc/cz_rules.cuh"

extern "C" int hr_legal_moves(const uint8_t *board, int side, uint16_t *out) {
    cz::Bits P;
    cz::bits_from_board(board, P);                    // the device builds the same three 90-bit sets with ballots
    int n = 0, Ksq = -1, ksq = -1;
    for (int sq = 0; sq < 90; sq++) {                 // exclusive scan over the mover's pieces == concatenation in square order
        const int p = board[sq];
        if (p == 1) Ksq = sq;
        if (p == 8) ksq = sq;
        if (p == 0 || cz::piece_red(p) != (side == 0)) continue;
        uint16_t slot[18];
        const int c = cz::gen_piece_bits(P, p, sq, slot);
        for (int j = 0; j < c; j++) out[n++] = slot[j];
    }
    if (cz::kings_face(P, Ksq, ksq))                  // flying general, appended last (same test as warp_legal_moves)
        out[n++] = side == 0 ? (uint16_t)(Ksq | (ksq << 7)) : (uint16_t)(ksq | (Ksq << 7));
    return n;
}

extern "C" void hr_encode_f32(const uint8_t *board, int side, float *out) {
    for (int lane = 0; lane < 32; lane++) cz::warp_encode<float>(board, side, out, lane);
}
