// The 8 symmetries of a Tak board (reference: tak/src/symm.rs:5-55).  The reference enumerates them as
//   [id, rot, rot^2, rot^3, mirror, mirror.rot, mirror.rot^2, mirror.rot^3]      (mirror first, then k%4 rotations)
// through takparse's Square::{rotate,mirror} and Direction::{rotate,mirror}.  takparse 0.5.5 is not vendored, so the
// sense of `rotate` / axis of `mirror` is parity-unpinned; any consistent choice yields the
// same SET of 8 augmented examples.  Here: rotate (col,row) -> (row, N-1-col), mirror col -> N-1-col, and directions
// transform as the unit vectors they are.  Moves cross the ABI as u16: bits 0-5 square (row*N+col), bits 6-7
// piece / direction (Up, Down, Left, Right = 0..3), bits 8-15 pattern mask (0 = placement).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace tb {

TB_HD void sym_square(int n, int k, int col, int row, int* out_col, int* out_row) {
    int c = col, r = row;
    if (k >= 4) c = n - 1 - c;
    for (int i = 0; i < (k & 3); ++i) {
        const int t = c;
        c = r;
        r = n - 1 - t;
    }
    *out_col = c;
    *out_row = r;
}
// source square whose image under symmetry k is (col,row)
TB_HD void sym_square_inv(int n, int k, int col, int row, int* out_col, int* out_row) {
    int c = col, r = row;
    for (int i = 0; i < (k & 3); ++i) {
        const int t = c;
        c = n - 1 - r;
        r = t;
    }
    if (k >= 4) c = n - 1 - c;
    *out_col = c;
    *out_row = r;
}
TB_HD int sym_direction(int k, int dir) {
    int dc = dir == 2 ? -1 : dir == 3 ? 1 : 0;   // Left / Right
    int dr = dir == 0 ? 1 : dir == 1 ? -1 : 0;   // Up / Down
    if (k >= 4) dc = -dc;
    for (int i = 0; i < (k & 3); ++i) {
        const int t = dc;
        dc = dr;
        dr = -t;
    }
    return dr == 1 ? 0 : dr == -1 ? 1 : dc == -1 ? 2 : 3;
}
// Symmetry::<N>::symmetries(move)[k] (symm.rs:41-55): the square moves, a spread's direction turns, the pattern stays
TB_HD uint16_t sym_move(int n, int k, uint16_t mv) {
    const int sq = mv & 63;
    int c, r;
    sym_square(n, k, sq % n, sq / n, &c, &r);
    int kind = (mv >> 6) & 3;
    if (mv >> 8) kind = sym_direction(k, kind);
    return uint16_t((mv & 0xFF00) | (kind << 6) | (r * n + c));
}

}  // namespace tb
