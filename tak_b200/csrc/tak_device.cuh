// Warp-cooperative Tak game state for sm_100a: one warp holds one game, lane l owns squares l and l+32.
//
// Device-side replacement for tak::Game<N> / Board<N> / Tile (reference: tak/src/game.rs:25-35, board.rs:8-10,
// tile.rs:7-10) and the three hot functions possible_moves (move_gen.rs:7-102), play (game.rs:121-209) and
// result (game.rs:220-267, board.rs:61-113).
//
// HBM record ("packed state", S bytes, S = 288/384/1152 for N = 5/6/8):
//   cols[NSQ]   stack colours per square, bit i = piece i is Black (bit 0 = bottom); u64 for N<=6, u128 above
//   hts[NSQ]    stack heights (u8)
//   walls,caps  bitboards of the top-piece kind
//   scalars     to_move, ply, reserves, half_komi, reversible_plies
//   occ,blk     DERIVED bitboards (occupied squares, squares whose top piece is Black), rewritten by every store:
//               a thread-per-state kernel can classify / count a position from the last 96 bytes of its record
// Squares are stored in MOVE-GENERATION order o = col*N + row (the reference enumerates `for x {for y}`),
// so a warp prefix sum over per-square move counts yields the reference's move order directly, and a
// warp's loads/stores of cols[] are one coalesced 256-byte access.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

namespace tb {

constexpr unsigned FULL = 0xFFFFFFFFu;

__host__ __device__ constexpr int tak_stones(int n) {
    return n == 3 ? 10 : n == 4 ? 15 : n == 5 ? 21 : n == 6 ? 30 : n == 7 ? 40 : 50;
}
__host__ __device__ constexpr int tak_caps(int n) { return n <= 4 ? 0 : n <= 6 ? 1 : 2; }

template <int N>
struct StateLayout {
    static constexpr int NSQ = N * N;
    using Col = std::conditional_t<(N <= 6), uint64_t, unsigned __int128>;
    static constexpr int COLS_BYTES = NSQ * int(sizeof(Col));
    static constexpr int HTS_OFF = (COLS_BYTES + 15) / 16 * 16;
    static constexpr int HTS_BYTES = (NSQ + 15) / 16 * 16;
    static constexpr int BB_OFF = HTS_OFF + HTS_BYTES;   // walls, caps
    static constexpr int SC_OFF = BB_OFF + 16;           // 16 bytes of scalars
    static constexpr int DER_OFF = SC_OFF + 16;          // derived: occupied, black tops
    static constexpr int RAW = DER_OFF + 16;
    static constexpr int S = (RAW + 31) / 32 * 32;
};

struct StateScalars {  // 16 bytes
    uint8_t to_move;   // 0 White 1 Black
    uint8_t ws, wc, bs, bc;
    int8_t half_komi;
    uint8_t reversible;
    uint8_t _pad;
    uint16_t ply;
    uint16_t _pad2[3];
};
static_assert(sizeof(StateScalars) == 16, "scalars must be 16 bytes");

// Move counts of one stack in one direction, summed over the pickup sizes p = 1..maxp:
//   d_sum_le[maxp][f]   = sum_p #(compositions of p into at most f parts)        (f = free squares along the ray)
//   d_sum_flat[maxp][f] = sum_p #(compositions of p into exactly f+1 parts whose last part is 1)   (a capstone flattening
//                         the wall that ends the ray; [p][0] counts p == 1)
// The lanes of a warp index these with different (maxp, f), so they live in global memory and are read through L1
// (__ldg): a __constant__ table serialises divergent indices.
static __device__ uint16_t d_sum_le[9][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0},       {0, 1, 1, 1, 1, 1, 1, 1},        {0, 2, 3, 3, 3, 3, 3, 3},
    {0, 3, 6, 7, 7, 7, 7, 7},       {0, 4, 10, 14, 15, 15, 15, 15},  {0, 5, 15, 25, 30, 31, 31, 31},
    {0, 6, 21, 41, 56, 62, 63, 63}, {0, 7, 28, 63, 98, 119, 126, 127}, {0, 8, 36, 92, 162, 218, 246, 254}};
static __device__ uint16_t d_sum_flat[9][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0}, {1, 0, 0, 0, 0, 0, 0, 0},  {1, 1, 0, 0, 0, 0, 0, 0},
    {1, 2, 1, 0, 0, 0, 0, 0}, {1, 3, 3, 1, 0, 0, 0, 0},  {1, 4, 6, 4, 1, 0, 0, 0},
    {1, 5, 10, 10, 5, 1, 0, 0}, {1, 6, 15, 20, 15, 6, 1, 0}, {1, 7, 21, 35, 35, 21, 7, 1}};
__device__ __forceinline__ int sum_le(int maxp, int f) { return __ldg(&d_sum_le[maxp][f]); }
__device__ __forceinline__ int sum_flat(int maxp, int f) { return __ldg(&d_sum_flat[maxp][f]); }

// GameResult byte (include/taknative.h): 0 ongoing, 1 white, 2 black, 3 draw, |0x10 road / reversible
enum : uint8_t { RES_ONGOING = 0, RES_WHITE = 1, RES_BLACK = 2, RES_DRAW = 3, RES_FLAG = 0x10 };

template <int N>
struct WarpGame {
    using L = StateLayout<N>;
    using Col = typename L::Col;
    static constexpr int NSQ = L::NSQ;
    static constexpr bool TWO = NSQ > 32;
    static constexpr uint64_t ALL = NSQ == 64 ? ~0ull : ((1ull << NSQ) - 1);

    Col c0, c1;        // colours of squares lane / lane+32
    int h0, h1;        // heights
    uint64_t walls, caps;  // uniform
    int to_move, ply, ws, wc, bs, bc, half_komi, reversible;  // uniform

    __device__ __forceinline__ static int lane() { return threadIdx.x & 31; }
    __device__ __forceinline__ static int row_of(int o) { return o % N; }
    __device__ __forceinline__ static int col_of(int o) { return o / N; }

    __device__ __forceinline__ void load(const uint8_t* rec) {
        const int l = lane();
        const Col* cols = reinterpret_cast<const Col*>(rec);
        const uint8_t* hts = rec + L::HTS_OFF;
        c0 = 0; c1 = 0; h0 = 0; h1 = 0;
        if (l < NSQ) { c0 = cols[l]; h0 = hts[l]; }
        if (TWO && l + 32 < NSQ) { c1 = cols[l + 32]; h1 = hts[l + 32]; }
        const uint64_t* bb = reinterpret_cast<const uint64_t*>(rec + L::BB_OFF);
        walls = bb[0];
        caps = bb[1];
        const uint4 sv = *reinterpret_cast<const uint4*>(rec + L::SC_OFF);
        StateScalars sc;
        *reinterpret_cast<uint4*>(&sc) = sv;
        to_move = sc.to_move; ply = sc.ply; ws = sc.ws; wc = sc.wc; bs = sc.bs; bc = sc.bc;
        half_komi = sc.half_komi; reversible = sc.reversible;
    }
    __device__ __forceinline__ void store(uint8_t* rec) const {
        const int l = lane();
        Col* cols = reinterpret_cast<Col*>(rec);
        uint8_t* hts = rec + L::HTS_OFF;
        if (l < NSQ) { cols[l] = c0; hts[l] = uint8_t(h0); }
        if (TWO && l + 32 < NSQ) { cols[l + 32] = c1; hts[l + 32] = uint8_t(h1); }
        const uint64_t occ = occupied(), blk = black_tops();   // whole-warp ballots
        if (l == 0) {
            uint64_t* bb = reinterpret_cast<uint64_t*>(rec + L::BB_OFF);
            bb[0] = walls;
            bb[1] = caps;
            uint64_t* der = reinterpret_cast<uint64_t*>(rec + L::DER_OFF);
            der[0] = occ;
            der[1] = blk;
            StateScalars sc{};
            sc.to_move = uint8_t(to_move); sc.ply = uint16_t(ply);
            sc.ws = uint8_t(ws); sc.wc = uint8_t(wc); sc.bs = uint8_t(bs); sc.bc = uint8_t(bc);
            sc.half_komi = int8_t(half_komi); sc.reversible = uint8_t(reversible);
            *reinterpret_cast<uint4*>(rec + L::SC_OFF) = *reinterpret_cast<uint4*>(&sc);
        }
    }
    __device__ __forceinline__ void reset(int hk) {
        c0 = 0; c1 = 0; h0 = 0; h1 = 0; walls = 0; caps = 0;
        to_move = 0; ply = 0; ws = bs = tak_stones(N); wc = bc = tak_caps(N); half_komi = hk; reversible = 0;
    }

    // bitboard from per-lane predicates of the two owned squares
    __device__ __forceinline__ uint64_t bb(bool p0, bool p1) const {
        uint64_t lo = __ballot_sync(FULL, p0);
        if (!TWO) return lo & ALL;
        uint64_t hi = __ballot_sync(FULL, p1);
        return (lo | (hi << 32)) & ALL;
    }
    __device__ __forceinline__ static bool top_black(Col c, int h) { return h > 0 && ((c >> (h - 1)) & 1); }
    __device__ __forceinline__ uint64_t occupied() const { return bb(h0 > 0, h1 > 0); }
    __device__ __forceinline__ uint64_t black_tops() const { return bb(top_black(c0, h0), top_black(c1, h1)); }

    // ---- Game::result (game.rs:220-267) ----------------------------------------------------------------
    __device__ __forceinline__ static bool has_road(uint64_t road) {
        // o = col*N + row: +1 is Up (row+1), +N is Right (col+1)
        uint64_t row0 = 0, rowL = 0;
#pragma unroll
        for (int c = 0; c < N; ++c) { row0 |= 1ull << (c * N); rowL |= 1ull << (c * N + N - 1); }
        const uint64_t col0 = (1ull << N) - 1, colL = col0 << (N * (N - 1));
        // vertical: from row 0 to row N-1 (board.rs:79-86); horizontal: col 0 to col N-1 (board.rs:88-93)
        uint64_t reach = road & row0, reach2 = road & col0;
        for (;;) {
            uint64_t a = reach | ((reach << 1) & ~row0) | ((reach >> 1) & ~rowL) | (reach << N) | (reach >> N);
            uint64_t b = reach2 | ((reach2 << 1) & ~row0) | ((reach2 >> 1) & ~rowL) | (reach2 << N) | (reach2 >> N);
            a &= road; b &= road;
            if (a == reach && b == reach2) break;
            reach = a; reach2 = b;
        }
        return (reach & rowL) != 0 || (reach2 & colL) != 0;
    }
    __device__ __forceinline__ uint8_t result() const {
        const uint64_t occ = occupied();
        const uint64_t blk = black_tops();
        const uint64_t wht = occ & ~blk;
        const uint64_t road_w = wht & ~walls, road_b = blk & ~walls;
        const uint64_t road_prev = to_move == 0 ? road_b : road_w;  // player who just moved
        const uint64_t road_cur = to_move == 0 ? road_w : road_b;
        if (has_road(road_prev)) return uint8_t((to_move == 0 ? RES_BLACK : RES_WHITE) | RES_FLAG);
        if (has_road(road_cur)) return uint8_t((to_move == 0 ? RES_WHITE : RES_BLACK) | RES_FLAG);
        if ((wc == 0 && ws == 0) || (bc == 0 && bs == 0) || occ == ALL) {
            const uint64_t flat = ~(walls | caps);
            const int fd = __popcll(wht & flat) - __popcll(blk & flat);
            const int k = half_komi / 2;  // truncating, as i8 division
            if (fd > k) return RES_WHITE;
            if (fd < k) return RES_BLACK;
            return (half_komi % 2 == 0) ? RES_DRAW : RES_BLACK;
        }
        if (reversible >= 50) return uint8_t(RES_DRAW | RES_FLAG);
        return RES_ONGOING;
    }
    __device__ __forceinline__ int flat_diff() const {
        const uint64_t occ = occupied();
        const uint64_t blk = black_tops();
        const uint64_t flat = ~(walls | caps);
        return __popcll(occ & ~blk & flat) - __popcll(blk & flat);
    }

    // ---- move generation (move_gen.rs:7-102) -----------------------------------------------------------
    // free run of droppable squares from o in direction d (0 Up 1 Down 2 Left 3 Right); wall_after = the
    // run is ended by a wall (a capstone on top of the moving stack may flatten it)
    __device__ __forceinline__ void free_run(int o, int d, int& free, bool& wall_after) const {
        const int r = row_of(o), c = col_of(o);
        const int dist = d == 0 ? N - 1 - r : d == 1 ? r : d == 2 ? c : N - 1 - c;
        const int delta = d == 0 ? 1 : d == 1 ? -1 : d == 2 ? -N : N;
        const uint64_t blockers = walls | caps;
        free = 0;
        wall_after = false;
        int q = o;
        for (int s = 0; s < dist; ++s) {
            q += delta;
            if ((blockers >> q) & 1) {
                wall_after = (walls >> q) & 1;
                break;
            }
            ++free;
        }
    }
    __device__ __forceinline__ int my_stones() const { return to_move == 0 ? ws : bs; }
    __device__ __forceinline__ int my_caps() const { return to_move == 0 ? wc : bc; }
    __device__ __forceinline__ bool mover_black() const { return ply < 2 ? (to_move == 0) : (to_move == 1); }

    __device__ __forceinline__ int count_square(int o, Col c, int h) const {
        if (o >= NSQ) return 0;
        if (ply < 2) return h == 0 ? 1 : 0;
        if (h == 0) return (my_stones() > 0 ? 2 : 0) + (my_caps() > 0 ? 1 : 0);
        if (top_black(c, h) != (to_move == 1)) return 0;
        const bool is_cap = (caps >> o) & 1;
        const int maxp = h < N ? h : N;
        int total = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            int free;
            bool wall_after;
            free_run(o, d, free, wall_after);
            const bool flatten = wall_after && is_cap;
            total += sum_le(maxp, free) + (flatten ? sum_flat(maxp, free) : 0);
        }
        return total;
    }
    // per-lane move counts of its two squares
    __device__ __forceinline__ void count_moves(int& n0, int& n1) const {
        n0 = count_square(lane(), c0, h0);
        n1 = TWO ? count_square(lane() + 32, c1, h1) : 0;
    }
    __device__ __forceinline__ static uint16_t abi_square(int o) { return uint16_t(row_of(o) * N + col_of(o)); }

    template <class Emit>
    __device__ __forceinline__ void emit_square(int o, Col c, int h, int base, Emit&& emit) const {
        if (o >= NSQ) return;
        const uint16_t sq = abi_square(o);
        int k = base;
        if (ply < 2) {
            if (h == 0) emit(k, sq);
            return;
        }
        if (h == 0) {
            if (my_stones() > 0) { emit(k++, sq); emit(k++, uint16_t(sq | (1u << 6))); }
            if (my_caps() > 0) emit(k++, uint16_t(sq | (2u << 6)));
            return;
        }
        if (top_black(c, h) != (to_move == 1)) return;
        const bool is_cap = (caps >> o) & 1;
        const int maxp = h < N ? h : N;
        for (int d = 0; d < 4; ++d) {
            int free;
            bool wall_after;
            free_run(o, d, free, wall_after);
            const bool flatten = wall_after && is_cap;
            if (free == 0 && !flatten) continue;
            for (int p = 1; p <= maxp; ++p) {
                // drop sequences in descending lexicographic order == odd p-bit patterns in ascending order
                for (unsigned v = 1; v < (1u << p); v += 2) {
                    const int parts = __popc(v);
                    const bool ok = parts <= free || (flatten && parts == free + 1 && (p == 1 || (v & 2)));
                    if (ok) emit(k++, uint16_t(sq | (unsigned(d) << 6) | ((v << (8 - p)) << 8)));
                }
            }
        }
    }
    // total number of legal moves; emit(k, move) is called for move k in reference order
    template <class Emit>
    __device__ __forceinline__ int generate(Emit&& emit) const {
        int n0, n1;
        count_moves(n0, n1);
        // exclusive scan over squares 0..31 then 32..63
        int inc0 = n0;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            int t = __shfl_up_sync(FULL, inc0, s);
            if (lane() >= s) inc0 += t;
        }
        const int tot0 = __shfl_sync(FULL, inc0, 31);
        emit_square(lane(), c0, h0, inc0 - n0, emit);
        int total = tot0;
        if (TWO) {
            int inc1 = n1;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                int t = __shfl_up_sync(FULL, inc1, s);
                if (lane() >= s) inc1 += t;
            }
            total += __shfl_sync(FULL, inc1, 31);
            emit_square(lane() + 32, c1, h1, tot0 + inc1 - n1, emit);
        }
        return total;
    }
    __device__ __forceinline__ int count_total() const {
        int n0, n1;
        count_moves(n0, n1);
        int s = n0 + n1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        return s;
    }
    // possible_moves()[pick(len)] without materialising the list: only the lane whose square owns index `pick`
    // enumerates its moves.  `total` receives possible_moves().len(); pick(total) must be in [0, total).
    template <class PickFn>
    __device__ __forceinline__ uint16_t select_move(PickFn&& pick_of, int& total) const {
        int n0, n1;
        count_moves(n0, n1);
        int inc0 = n0, inc1 = n1;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t0 = __shfl_up_sync(FULL, inc0, s);
            const int t1 = __shfl_up_sync(FULL, inc1, s);
            if (lane() >= s) { inc0 += t0; inc1 += t1; }
        }
        const int tot0 = __shfl_sync(FULL, inc0, 31);
        total = tot0 + __shfl_sync(FULL, inc1, 31);
        if (total == 0) return 0xFFFF;
        const int pick = pick_of(total);
        uint16_t chosen = 0xFFFF;
        auto emit = [&](int k, uint16_t mv) { if (k == pick) chosen = mv; };
        const int b0 = inc0 - n0, b1 = tot0 + inc1 - n1;
        if (pick >= b0 && pick < b0 + n0) emit_square(lane(), c0, h0, b0, emit);
        if (TWO && pick >= b1 && pick < b1 + n1) emit_square(lane() + 32, c1, h1, b1, emit);
        const unsigned m = __ballot_sync(FULL, chosen != 0xFFFF);
        return uint16_t(__shfl_sync(FULL, int(chosen), m ? __ffs(m) - 1 : 0));
    }

    // ---- Game::play (game.rs:121-209) ---------------------------------------------------------------------
    // CHECK = validate like the reference and return its PlayError code; without CHECK the move must be legal.
    template <bool CHECK>
    __device__ __forceinline__ int play(uint16_t mv) {
        const int sq = mv & 63;
        const unsigned mask = mv >> 8;
        const int kind = (mv >> 6) & 3;
        if (CHECK && sq >= NSQ) return -1;  // OutOfBounds
        const int o = (sq % N) * N + (sq / N);
        const int l = lane();
        const bool swapped = ply < 2;
        if (mask == 0) {
            // ---- execute_place (game.rs:147-169)
            if (CHECK) {
                const uint64_t occ = occupied();
                if ((occ >> o) & 1) return -2;
                if (kind == 2 && my_caps() == 0) return -3;
                if (kind <= 1 && my_stones() == 0) return -4;
                if (kind == 3) return -32;
                if (swapped && kind != 0) return -5;
            }
            const bool black = mover_black();
            if (l == (o & 31)) {
                if (o < 32) { c0 = black ? 1 : 0; h0 = 1; }
                else { c1 = black ? 1 : 0; h1 = 1; }
            }
            if (kind == 1) walls |= 1ull << o;
            if (kind == 2) caps |= 1ull << o;
            if (kind <= 1) {
                if ((to_move == 0) != swapped) ws -= 1; else bs -= 1;
            } else {
                if (to_move == 0) wc -= 1; else bc -= 1;
            }
            reversible = 0;
        } else {
            // ---- execute_spread (game.rs:171-209)
            const int p = 8 - (__ffs(mask) - 1);   // pattern.count_pieces()
            const int drops = __popc(mask);
            const int d = kind;
            const int r = row_of(o), cc = col_of(o);
            const int delta = d == 0 ? 1 : d == 1 ? -1 : d == 2 ? -N : N;
            // source column / height, broadcast from the owning lane
            const int src_lane = o & 31;
            Col sc = (o < 32) ? c0 : c1;
            int sh = (o < 32) ? h0 : h1;
            if constexpr (sizeof(Col) == 8) {
                sc = __shfl_sync(FULL, sc, src_lane);
            } else {
                uint64_t lo = uint64_t(sc), hi = uint64_t(sc >> 64);
                lo = __shfl_sync(FULL, lo, src_lane);
                hi = __shfl_sync(FULL, hi, src_lane);
                sc = (Col(hi) << 64) | lo;
            }
            sh = __shfl_sync(FULL, sh, src_lane);
            const bool src_cap = (caps >> o) & 1, src_wall = (walls >> o) & 1;
            if (CHECK) {
                if (sh == 0) return -6;                                      // EmptySquare
                if (top_black(sc, sh) != mover_black()) return -7;           // StackNotOwned
                if (p > N) return -11;                                       // TakeError::CarryLimit
                if (p > sh) return -12;                                      // TakeError::StackSize
                const int dist = d == 0 ? N - 1 - r : d == 1 ? r : d == 2 ? cc : N - 1 - cc;
                // walk the drops in order and report the first error the reference would hit
                int q = o;
                unsigned m = mask;
                for (int t = 1; t <= drops; ++t) {
                    if (t > dist) return -13;                                // SpreadOutOfBounds
                    q += delta;
                    const int lead = __clz(m << 24);                         // zeros before this drop's one bit
                    const int dt = lead + 1;
                    m = (m << dt) & 0xFF;
                    const bool last_single_cap = (t == drops) && dt == 1 && src_cap;
                    if ((caps >> q) & 1) return -9;                          // StackError::Cap
                    if (((walls >> q) & 1) && !last_single_cap) return -8;   // StackError::Wall
                }
            }
            const Col carry = (sc >> (sh - p)) & ((Col(1) << p) - 1);  // bit 0 = bottom-most carried piece
            // my squares: am I the source, or at distance t along the direction?
#pragma unroll
            for (int half = 0; half < (TWO ? 2 : 1); ++half) {
                const int q = l + 32 * half;
                if (q >= NSQ) continue;
                Col& qc = half ? c1 : c0;
                int& qh = half ? h1 : h0;
                if (q == o) {
                    qh = sh - p;
                    qc = sc & ((Col(1) << qh) - 1);
                    continue;
                }
                const int diff = q - o;
                int t = 0;
                if (d == 0 && col_of(q) == cc && diff > 0) t = diff;
                if (d == 1 && col_of(q) == cc && diff < 0) t = -diff;
                if (d == 2 && row_of(q) == r && diff < 0) t = -diff / N;
                if (d == 3 && row_of(q) == r && diff > 0) t = diff / N;
                if (t == 0 || t > drops) continue;
                // offset/length of drop t inside the carried pieces
                unsigned m = mask;
                int off = 0, dt = 0;
                for (int i = 1; i <= t; ++i) {
                    off += dt;
                    dt = __clz(m << 24) + 1;
                    m = (m << dt) & 0xFF;
                }
                const Col seg = (carry >> off) & ((Col(1) << dt) - 1);
                qc |= seg << qh;
                qh += dt;
            }
            // top-piece kinds: the source loses its kind; the last drop square takes it
            const int last = o + delta * drops;
            walls &= ~(1ull << o);
            caps &= ~(1ull << o);
            walls &= ~(1ull << last);  // a flattened wall
            if (src_wall) walls |= 1ull << last;
            if (src_cap) caps |= 1ull << last;
            reversible = (reversible + 1) & 0xFF;
        }
        ply += 1;
        to_move ^= 1;
        return 0;
    }
};

// ---- one THREAD per position: result and number of legal moves from the record's tail ---------------------------
// perft's counting levels need only Game::result (game.rs:220-267) and possible_moves().len() (move_gen.rs:7-102), both
// functions of {heights, walls, caps, occupied, black tops, scalars} = the last 96 bytes of a 6x6 record.  A warp per
// position spends its 32 lanes on identical bitboard arithmetic; a thread per position does the same work once.
template <int N>
struct ThreadPos {
    using L = StateLayout<N>;
    static constexpr int NSQ = N * N;
    static constexpr uint64_t ALL = NSQ == 64 ? ~0ull : ((1ull << NSQ) - 1);
    const uint8_t* hts;
    uint64_t walls, caps, occ, blk;
    StateScalars sc;

    __device__ __forceinline__ void load(const uint8_t* rec) {
        hts = rec + L::HTS_OFF;
        const uint4 b = *reinterpret_cast<const uint4*>(rec + L::BB_OFF);
        const uint4 d = *reinterpret_cast<const uint4*>(rec + L::DER_OFF);
        walls = uint64_t(b.x) | (uint64_t(b.y) << 32);
        caps = uint64_t(b.z) | (uint64_t(b.w) << 32);
        occ = uint64_t(d.x) | (uint64_t(d.y) << 32);
        blk = uint64_t(d.z) | (uint64_t(d.w) << 32);
        *reinterpret_cast<uint4*>(&sc) = *reinterpret_cast<const uint4*>(rec + L::SC_OFF);
    }
    // Game::result, as WarpGame::result
    __device__ __forceinline__ uint8_t result() const {
        const uint64_t wht = occ & ~blk;
        const uint64_t road_w = wht & ~walls, road_b = blk & ~walls;
        const int to_move = sc.to_move;
        if (WarpGame<N>::has_road(to_move == 0 ? road_b : road_w))
            return uint8_t((to_move == 0 ? RES_BLACK : RES_WHITE) | RES_FLAG);
        if (WarpGame<N>::has_road(to_move == 0 ? road_w : road_b))
            return uint8_t((to_move == 0 ? RES_WHITE : RES_BLACK) | RES_FLAG);
        if ((sc.wc == 0 && sc.ws == 0) || (sc.bc == 0 && sc.bs == 0) || occ == ALL) {
            const uint64_t flat = ~(walls | caps);
            const int fd = __popcll(wht & flat) - __popcll(blk & flat);
            const int k = sc.half_komi / 2;
            if (fd > k) return RES_WHITE;
            if (fd < k) return RES_BLACK;
            return (sc.half_komi % 2 == 0) ? RES_DRAW : RES_BLACK;
        }
        if (sc.reversible >= 50) return uint8_t(RES_DRAW | RES_FLAG);
        return RES_ONGOING;
    }
    // possible_moves().len(), as WarpGame::count_total
    __device__ __forceinline__ uint32_t count_moves() const {
        return count_moves_h([&](int o) { return int(hts[o]); });
    }
    template <class H>
    __device__ __forceinline__ uint32_t count_moves_h(H&& height_of) const {
        const int empties = __popcll(ALL & ~occ);
        if (sc.ply < 2) return uint32_t(empties);
        const int to_move = sc.to_move;
        const int my_st = to_move == 0 ? sc.ws : sc.bs, my_cp = to_move == 0 ? sc.wc : sc.bc;
        uint32_t total = uint32_t(empties) * uint32_t((my_st > 0 ? 2 : 0) + (my_cp > 0 ? 1 : 0));
        const uint64_t blockers = walls | caps;
        uint64_t mine = occ & (to_move == 1 ? blk : ~blk);
        while (mine) {
            const int o = __ffsll(static_cast<long long>(mine)) - 1;
            mine &= mine - 1;
            const int h = height_of(o);
            const int maxp = h < N ? h : N;
            const bool is_cap = (caps >> o) & 1;
            const int r = o % N, c = o / N;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int dist = d == 0 ? N - 1 - r : d == 1 ? r : d == 2 ? c : N - 1 - c;
                const int delta = d == 0 ? 1 : d == 1 ? -1 : d == 2 ? -N : N;
                int free = 0, q = o;
                bool wall_after = false;
                for (int st = 0; st < dist; ++st) {
                    q += delta;
                    if ((blockers >> q) & 1) {
                        wall_after = (walls >> q) & 1;
                        break;
                    }
                    ++free;
                }
                total += sum_le(maxp, free) + ((wall_after && is_cap) ? sum_flat(maxp, free) : 0);
            }
        }
        return total;
    }
};

// ---- one THREAD per child: Game::play (game.rs:121-209) as a PATCH against the parent's packed record -----------------
// A child differs from its parent in at most N stack columns (the source of a spread and its <= N-1 drop squares, or the
// one square of a placement), a few height bytes and the 48-byte tail {walls, caps | scalars | occupied, black tops}.
// perft's expansion keeps per child, in shared memory, only
//   [ the record's words from the heights to the end (heights + tail, 96 B on 6x6), patched IN PLACE | the new columns,
//     sorted by square | the bitboard of the squares whose column changed ]                  = 176 B on 6x6, not 384,
// which is what lets ~1 300 children be in flight per SM.  The thread that owns the child patches the copy from the
// parent's record (columns read through L1: ~100 siblings share them), and classifies / counts the new position from the
// tail and heights it just wrote.  When the block streams the children out, a word of stack columns comes from the parent
// unless the bitboard names one of its squares; the new column is then entry popcount(bitboard below the square).
// The move must be legal (it comes from possible_moves of the parent).
template <int N>
struct ChildPatch {
    using L = StateLayout<N>;
    using Col = typename L::Col;
    static constexpr int NSQ = N * N;
    static constexpr int HT0 = L::HTS_OFF / 16;            // first word of the heights = number of stack-column words
    static constexpr int TW = L::S / 16 - HT0;             // words from the heights to the end of the record
    static constexpr int OFF_TAIL = 0, OFF_DIRTY = TW * 16, OFF_COLS = OFF_DIRTY + 16;
    static constexpr int RAW = OFF_COLS + N * int(sizeof(Col));
    static constexpr int STRIDE = (((RAW + 15) / 16) | 1) * 16;   // an odd number of 16-byte units: conflict-free 16-byte accesses

    // `img` holds the parent's words [HT0, S/16) on entry and the child's on exit; `tp` returns the child's tail
    __device__ __forceinline__ static void build(const uint8_t* parent, uint16_t mv, uint8_t* img, ThreadPos<N>& tp) {
        const Col* pcols = reinterpret_cast<const Col*>(parent);
        uint8_t* hts = img + OFF_TAIL;                         // the heights are the first bytes of the imaged region
        uint8_t* tail = img + OFF_TAIL + (L::BB_OFF - L::HTS_OFF);
        Col* ocols = reinterpret_cast<Col*>(img + OFF_COLS);
        {   // ThreadPos::load from the image
            const uint4 bb = *reinterpret_cast<const uint4*>(tail);
            const uint4 d = *reinterpret_cast<const uint4*>(tail + 32);
            tp.walls = uint64_t(bb.x) | (uint64_t(bb.y) << 32);
            tp.caps = uint64_t(bb.z) | (uint64_t(bb.w) << 32);
            tp.occ = uint64_t(d.x) | (uint64_t(d.y) << 32);
            tp.blk = uint64_t(d.z) | (uint64_t(d.w) << 32);
            *reinterpret_cast<uint4*>(&tp.sc) = *reinterpret_cast<const uint4*>(tail + 16);
            tp.hts = hts;
        }
        StateScalars& sc = tp.sc;
        const int sq = mv & 63;
        const unsigned mask = mv >> 8;
        const int kind = (mv >> 6) & 3;
        const int o = (sq % N) * N + (sq / N);
        const uint64_t obit = 1ull << o;
        const bool swapped = sc.ply < 2;
        uint64_t dirty = obit;
        if (mask == 0) {
            // execute_place (game.rs:147-169)
            const bool black = swapped ? (sc.to_move == 0) : (sc.to_move == 1);
            ocols[0] = black ? Col(1) : Col(0);
            hts[o] = 1;
            if (kind == 1) tp.walls |= obit;
            if (kind == 2) tp.caps |= obit;
            if (kind <= 1) {
                if ((sc.to_move == 0) != swapped) sc.ws -= 1; else sc.bs -= 1;
            } else {
                if (sc.to_move == 0) sc.wc -= 1; else sc.bc -= 1;
            }
            sc.reversible = 0;
            tp.occ |= obit;
            if (black) tp.blk |= obit;
        } else {
            // execute_spread (game.rs:171-209)
            const int p = 8 - (__ffs(mask) - 1);
            const int drops = __popc(mask);
            const int delta = kind == 0 ? 1 : kind == 1 ? -1 : kind == 2 ? -N : N;
            // the new columns are stored sorted by square: along the ray for Up / Right, against it for Down / Left
            const bool asc = delta > 0;
            const Col src = pcols[o];
            const int sh = hts[o];
            const bool src_cap = (tp.caps >> o) & 1, src_wall = (tp.walls >> o) & 1;
            const unsigned carry = unsigned(src >> (sh - p)) & ((1u << p) - 1);   // bit 0 = bottom-most carried piece
            const int rem_h = sh - p;
            const Col rem = src & ((Col(1) << rem_h) - 1);
            ocols[asc ? 0 : drops] = rem;
            hts[o] = uint8_t(rem_h);
            if (rem_h == 0) {
                tp.occ &= ~obit;
                tp.blk &= ~obit;
            } else if ((rem >> (rem_h - 1)) & 1) {
                tp.blk |= obit;
            } else {
                tp.blk &= ~obit;
            }
            unsigned m = mask;
            int off = 0, q = o;
            for (int t = 1; t <= drops; ++t) {
                q += delta;
                const int dt = __clz(m << 24) + 1;
                m = (m << dt) & 0xFF;
                const unsigned seg = (carry >> off) & ((1u << dt) - 1);
                off += dt;
                const int qh = hts[q];
                ocols[asc ? t : drops - t] = pcols[q] | (Col(seg) << qh);
                hts[q] = uint8_t(qh + dt);
                const uint64_t qbit = 1ull << q;
                dirty |= qbit;
                tp.occ |= qbit;
                if ((seg >> (dt - 1)) & 1) tp.blk |= qbit; else tp.blk &= ~qbit;
            }
            // top-piece kinds: the source loses its kind; the last drop square takes it (a wall there is flattened)
            const uint64_t lbit = 1ull << q;
            tp.walls &= ~(obit | lbit);
            tp.caps &= ~obit;
            if (src_wall) tp.walls |= lbit;
            if (src_cap) tp.caps |= lbit;
            sc.reversible = uint8_t(sc.reversible + 1);
        }
        sc.ply = uint16_t(sc.ply + 1);
        sc.to_move ^= 1;
        uint4* tw = reinterpret_cast<uint4*>(tail);
        tw[0] = make_uint4(uint32_t(tp.walls), uint32_t(tp.walls >> 32), uint32_t(tp.caps), uint32_t(tp.caps >> 32));
        tw[1] = *reinterpret_cast<const uint4*>(&sc);
        tw[2] = make_uint4(uint32_t(tp.occ), uint32_t(tp.occ >> 32), uint32_t(tp.blk), uint32_t(tp.blk >> 32));
        *reinterpret_cast<uint2*>(img + OFF_DIRTY) = make_uint2(unsigned(dirty), unsigned(dirty >> 32));
    }

    // word `wd` < HT0 (stack columns) of the child's record
    __device__ __forceinline__ static uint4 cols_word(const uint4* parent, const uint8_t* img, int wd) {
        uint4 v = __ldg(parent + wd);
        const uint2 dw = *reinterpret_cast<const uint2*>(img + OFF_DIRTY);
        const uint64_t dirty = uint64_t(dw.x) | (uint64_t(dw.y) << 32);
        if constexpr (sizeof(Col) == 8) {
            const int q = 2 * wd;
            const unsigned hit = unsigned(dirty >> q) & 3u;
            if (hit) {
                const int j = __popcll(dirty & ((1ull << q) - 1));      // new columns are sorted by square
                if (hit & 1) {
                    const uint2 c = *reinterpret_cast<const uint2*>(img + OFF_COLS + 8 * j);
                    v.x = c.x; v.y = c.y;
                }
                if (hit & 2) {
                    const uint2 c = *reinterpret_cast<const uint2*>(img + OFF_COLS + 8 * (j + (hit & 1)));
                    v.z = c.x; v.w = c.y;
                }
            }
        } else {
            if ((dirty >> wd) & 1) v = *reinterpret_cast<const uint4*>(img + OFF_COLS + 16 * __popcll(dirty & ((1ull << wd) - 1)));
        }
        return v;
    }
};

}  // namespace tb
