// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// CPU oracle: a literal C++ restatement of the reference's `tak` crate (ViliamVadocz/tak), used only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker.
// Nothing under tak_b200/ may include, link or call this.
//
// Parity status: PINNED against the reference's own known answers (tests/test_oracle_golden.py):
//   tak/tests/perft.rs:20-99, tak/tests/wins.rs:5-67, tak/tests/tps.rs:5-24 (golden TPS) and :26-96,
//   alpha-tak/src/repr/tests.rs:11-111, alpha-tak/src/search/tests.rs:38-72, move_map.rs:51-201 (1575 table).
// The reference itself (Rust) cannot be compiled here (no cargo/rustc), so there is no oracle/_ref.
// Third-party semantics restated from takparse 0.5.5 (Cargo.lock:1059-1062): PTN move grammar, Pattern::mask
// bit order (UNPINNED by any reference test -- see DESIGN.md), Square/Direction stepping, TPS text.
//
// Data structures deliberately mirror the Rust ones (Vec-of-colours tiles, recursive flood fill, LIFO DFS
// move generation) so that iteration orders are identical by construction, not by cleverness.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace oracle {

enum Color : uint8_t { White = 0, Black = 1 };
inline Color opp(Color c) { return c == White ? Black : White; }
enum Piece : uint8_t { Flat = 0, Wall = 1, Cap = 2 };
// order used by move generation (move_gen.rs:62): Up, Down, Left, Right
enum Direction : uint8_t { Up = 0, Down = 1, Left = 2, Right = 3 };

// status codes shared with include/taknative.h (tak::PlayError, tak/src/error.rs:4-15)
enum Status : int {
    Ok = 0,
    OutOfBounds = -1,
    AlreadyOccupied = -2,
    NoCapstone = -3,
    NoStones = -4,
    OpeningNonFlat = -5,
    EmptySquare = -6,
    StackNotOwned = -7,
    StackWall = -8,
    StackCap = -9,
    TakeZero = -10,
    TakeCarryLimit = -11,
    TakeStackSize = -12,
    SpreadOutOfBounds = -13,
    ParseError = -35,
};

// GameResult as one byte (tak/src/game_result.rs:4-8): 0 Ongoing 1 White 2 Black 3 Draw, |0x10 = road / reversible
struct GameResult {
    uint8_t code = 0;
    bool ongoing() const { return (code & 0xF) == 0; }
    bool is_winner() const { return (code & 0xF) == 1 || (code & 0xF) == 2; }
    bool is_draw() const { return (code & 0xF) == 3; }
    Color winner() const { return (code & 0xF) == 1 ? White : Black; }
    bool operator==(const GameResult& o) const { return code == o.code; }
    bool operator!=(const GameResult& o) const { return code != o.code; }
    static GameResult Winner(Color c, bool road) { return {uint8_t((c == White ? 1 : 2) | (road ? 0x10 : 0))}; }
    static GameResult Draw(bool reversible) { return {uint8_t(3 | (reversible ? 0x10 : 0))}; }
    static GameResult Ongoing() { return {0}; }
};

// takparse Move: square + Place(piece) | Spread(direction, pattern).  `drops` lists the drop counts.
struct Move {
    uint8_t col = 0, row = 0;
    bool place = true;
    Piece piece = Flat;
    Direction dir = Up;
    std::vector<uint8_t> drops;

    bool operator==(const Move& o) const {
        return col == o.col && row == o.row && place == o.place &&
               (place ? piece == o.piece : (dir == o.dir && drops == o.drops));
    }
    int count_pieces() const {
        int s = 0;
        for (auto d : drops) s += d;
        return s;
    }
    // takparse Pattern::mask: MSB-first, each drop of c pieces = (c-1) zero bits then a one bit.
    uint8_t mask() const {
        uint32_t m = 0;
        int pos = 7;
        for (auto d : drops) {
            pos -= (d - 1);
            m |= (1u << pos);
            pos -= 1;
        }
        return uint8_t(m);
    }
    uint16_t encode(int n) const {
        uint16_t sq = uint16_t(row * n + col);
        if (place) return uint16_t(sq | (uint16_t(piece) << 6));
        return uint16_t(sq | (uint16_t(dir) << 6) | (uint16_t(mask()) << 8));
    }
    static Move decode(uint16_t m, int n) {
        Move mv;
        int sq = m & 63;
        mv.row = uint8_t(sq / n);
        mv.col = uint8_t(sq % n);
        uint8_t msk = uint8_t(m >> 8);
        if (msk == 0) {
            mv.place = true;
            mv.piece = Piece((m >> 6) & 3);
        } else {
            mv.place = false;
            mv.dir = Direction((m >> 6) & 3);
            int run = 0;
            int tz = __builtin_ctz(msk);
            for (int pos = 7; pos >= tz; --pos) {
                ++run;
                if (msk & (1u << pos)) {
                    mv.drops.push_back(uint8_t(run));
                    run = 0;
                }
            }
        }
        return mv;
    }
};

// ---- PTN (takparse Move FromStr / Display) -------------------------------------------------------------
//   place : [FSC]? file rank          spread: [count]? file rank dir [drops]?     dir: + - < >
inline bool parse_move(const std::string& s_in, int n, Move& out) {
    std::string s = s_in;
    while (!s.empty() && (s.back() == '\'' || s.back() == '!' || s.back() == '?' || s.back() == '*')) s.pop_back();
    if (s.empty()) return false;
    size_t i = 0;
    Move m;
    int count = -1;
    bool have_piece = false;
    if (s[i] >= '1' && s[i] <= '8') {
        count = s[i] - '0';
        ++i;
    } else if (s[i] == 'F' || s[i] == 'S' || s[i] == 'C') {
        m.piece = s[i] == 'F' ? Flat : (s[i] == 'S' ? Wall : Cap);
        have_piece = true;
        ++i;
    }
    if (i + 1 >= s.size() + 0 && i + 1 > s.size()) return false;
    if (i >= s.size() || s[i] < 'a' || s[i] > 'h') return false;
    m.col = uint8_t(s[i] - 'a');
    ++i;
    if (i >= s.size() || s[i] < '1' || s[i] > '8') return false;
    m.row = uint8_t(s[i] - '1');
    ++i;
    if (m.col >= n || m.row >= n) return false;
    if (i == s.size()) {
        if (count != -1) return false;
        m.place = true;
        out = m;
        return true;
    }
    if (have_piece) return false;
    char d = s[i++];
    if (d == '+') m.dir = Up;
    else if (d == '-') m.dir = Down;
    else if (d == '<') m.dir = Left;
    else if (d == '>') m.dir = Right;
    else return false;
    m.place = false;
    if (count == -1) count = 1;
    int sum = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '1' || s[i] > '8') return false;
        m.drops.push_back(uint8_t(s[i] - '0'));
        sum += s[i] - '0';
    }
    if (m.drops.empty()) {
        m.drops.push_back(uint8_t(count));
        sum = count;
    }
    if (sum != count || count > 8) return false;
    out = m;
    return true;
}

inline std::string format_move(const Move& m) {
    std::string s;
    if (m.place) {
        if (m.piece == Wall) s += 'S';
        if (m.piece == Cap) s += 'C';
        s += char('a' + m.col);
        s += char('1' + m.row);
        return s;
    }
    int count = m.count_pieces();
    if (count > 1) s += char('0' + count);
    s += char('a' + m.col);
    s += char('1' + m.row);
    s += "+-<>"[m.dir];
    if (m.drops.size() > 1)
        for (auto d : m.drops) s += char('0' + d);
    return s;
}

// ---- Tile (tak/src/tile.rs) ----------------------------------------------------------------------------
struct Tile {
    Piece piece = Flat;          // kind of the TOP piece; an empty tile has Flat (derive(Default), tile.rs:6-10)
    std::vector<Color> stack;    // bottom -> top

    bool is_empty() const { return stack.empty(); }
    size_t size() const { return stack.size(); }
    bool operator==(const Tile& o) const { return piece == o.piece && stack == o.stack; }

    // tile.rs:28-45
    int stack_on(Piece p, Color c) {
        switch (piece) {
            case Flat: break;
            case Wall:
                if (p != Cap) return StackWall;
                break;
            case Cap: return StackCap;
        }
        piece = p;
        stack.push_back(c);
        return Ok;
    }
    // tile.rs:49-63; carry is ordered top -> bottom
    int take(int n, size_t amount, Piece& out_piece, std::vector<Color>& carry) {
        if (amount == 0) return TakeZero;
        if (amount > size_t(n)) return TakeCarryLimit;
        if (amount > size()) return TakeStackSize;
        carry.clear();
        for (size_t i = 0; i < amount; ++i) {
            carry.push_back(stack.back());
            stack.pop_back();
        }
        out_piece = piece;
        piece = Flat;
        return Ok;
    }
};

inline void default_starting_stones(int n, uint8_t& stones, uint8_t& caps) {  // game.rs:10-20
    static const uint8_t S[9] = {0, 0, 0, 10, 15, 21, 30, 40, 50};
    static const uint8_t C[9] = {0, 0, 0, 0, 0, 1, 1, 2, 2};
    stones = S[n];
    caps = C[n];
}

// ---- Game (tak/src/game.rs, board.rs, move_gen.rs) -------------------------------------------------------
struct Game {
    int n = 5;
    std::vector<Tile> data;  // data[row * n + col]  == board.data[row][col]
    Color to_move = White;
    uint16_t ply = 0;
    uint8_t white_stones = 0, white_caps = 0, black_stones = 0, black_caps = 0;
    int8_t half_komi = 0;
    uint8_t reversible_plies = 0;

    explicit Game(int n_ = 5, int half_komi_ = 0) : n(n_), data(size_t(n_) * n_) {
        default_starting_stones(n, white_stones, white_caps);
        black_stones = white_stones;
        black_caps = white_caps;
        half_komi = int8_t(half_komi_);
    }
    Tile& at(int col, int row) { return data[size_t(row) * n + col]; }
    const Tile& at(int col, int row) const { return data[size_t(row) * n + col]; }

    bool is_swapped() const { return ply < 2; }                                 // game.rs:84-86
    Color color() const { return is_swapped() ? opp(to_move) : to_move; }       // game.rs:88-94
    void get_counts(uint8_t& stones, uint8_t& caps) const {                     // game.rs:96-101
        if (to_move == White) {
            stones = white_stones;
            caps = white_caps;
        } else {
            stones = black_stones;
            caps = black_caps;
        }
    }
    void dec_stones() {  // game.rs:103-109
        if ((to_move == White) ^ is_swapped()) white_stones -= 1;
        else black_stones -= 1;
    }
    void dec_caps() {  // game.rs:111-116
        if (to_move == White) white_caps -= 1;
        else black_caps -= 1;
    }

    static bool step(int& col, int& row, Direction d, int n) {  // takparse Square::checked_step
        int c = col, r = row;
        switch (d) {
            case Up: r += 1; break;
            case Down: r -= 1; break;
            case Left: c -= 1; break;
            case Right: c += 1; break;
        }
        if (c < 0 || r < 0 || c >= n || r >= n) return false;
        col = c;
        row = r;
        return true;
    }

    // game.rs:121-130
    int play(const Move& m) {
        int st = m.place ? execute_place(m) : execute_spread(m);
        if (st != Ok) return st;
        if (m.place) reversible_plies = 0;  // game.rs:211-218
        else reversible_plies = uint8_t(reversible_plies + 1);
        ply += 1;
        to_move = opp(to_move);
        return Ok;
    }

    int execute_place(const Move& m) {  // game.rs:147-169
        uint8_t stones, caps;
        get_counts(stones, caps);
        if (m.col >= n || m.row >= n) return OutOfBounds;
        if (!at(m.col, m.row).is_empty()) return AlreadyOccupied;
        if (m.piece == Cap && caps == 0) return NoCapstone;
        if ((m.piece == Flat || m.piece == Wall) && stones == 0) return NoStones;
        if (is_swapped() && (m.piece == Wall || m.piece == Cap)) return OpeningNonFlat;
        Tile t;
        t.piece = m.piece;
        t.stack.push_back(color());
        at(m.col, m.row) = t;
        if (m.piece == Flat || m.piece == Wall) dec_stones();
        else dec_caps();
        return Ok;
    }

    int execute_spread(const Move& m) {  // game.rs:171-209
        if (m.col >= n || m.row >= n) return OutOfBounds;
        Tile& src = at(m.col, m.row);
        if (src.is_empty()) return EmptySquare;
        if (src.stack.back() != color()) return StackNotOwned;
        int count = m.count_pieces();
        Piece piece;
        std::vector<Color> carry;  // top -> bottom
        int st = src.take(n, size_t(count), piece, carry);
        if (st != Ok) return st;
        // pieces: [top piece kind, Flat, Flat, ...]; pop() yields Flats first and the real top last
        std::vector<Piece> pieces;
        pieces.push_back(piece);
        for (int i = 0; i < count - 1; ++i) pieces.push_back(Flat);
        int col = m.col, row = m.row;
        for (auto drop : m.drops) {
            if (!step(col, row, m.dir, n)) return SpreadOutOfBounds;
            for (int i = 0; i < drop; ++i) {
                Piece p = pieces.back();
                pieces.pop_back();
                Color c = carry.back();
                carry.pop_back();
                int s2 = at(col, row).stack_on(p, c);
                if (s2 != Ok) return s2;
            }
        }
        return Ok;
    }

    // board.rs:61-75
    bool full() const {
        for (auto& t : data)
            if (t.is_empty()) return false;
        return true;
    }
    int8_t flat_diff() const {
        int d = 0;
        for (auto& t : data)
            if (!t.is_empty() && t.piece == Flat) d += (t.stack.back() == White) ? 1 : -1;
        return int8_t(d);
    }
    // board.rs:77-113 (recursive flood fill, literally)
    void find_paths_recursive(int x, int y, Color color, std::vector<uint8_t>& seen) const {
        if (y >= n || x >= n || y < 0 || x < 0 || seen[size_t(y) * n + x]) return;
        const Tile& t = at(x, y);
        if (!t.is_empty() && t.stack.back() == color && (t.piece == Flat || t.piece == Cap)) {
            seen[size_t(y) * n + x] = 1;
            find_paths_recursive(x + 1, y, color, seen);
            find_paths_recursive(x, y + 1, color, seen);
            if (x >= 1) find_paths_recursive(x - 1, y, color, seen);
            if (y >= 1) find_paths_recursive(x, y - 1, color, seen);
        }
    }
    bool find_paths(Color color) const {
        std::vector<uint8_t> seen(size_t(n) * n, 0);
        for (int x = 0; x < n; ++x) find_paths_recursive(x, 0, color, seen);
        for (int x = 0; x < n; ++x)
            if (seen[size_t(n - 1) * n + x]) return true;
        std::fill(seen.begin(), seen.end(), 0);
        for (int y = 0; y < n; ++y) find_paths_recursive(0, y, color, seen);
        for (int y = 0; y < n; ++y)
            if (seen[size_t(y) * n + (n - 1)]) return true;
        return false;
    }

    // game.rs:220-267
    GameResult result() const {
        if (find_paths(opp(to_move))) return GameResult::Winner(opp(to_move), true);
        if (find_paths(to_move)) return GameResult::Winner(to_move, true);
        if ((white_caps == 0 && white_stones == 0) || (black_caps == 0 && black_stones == 0) || full()) {
            int8_t fd = flat_diff();
            int8_t k = int8_t(half_komi / 2);  // truncating
            if (fd > k) return GameResult::Winner(White, false);
            if (fd < k) return GameResult::Winner(Black, false);
            if (half_komi % 2 == 0) return GameResult::Draw(false);
            return GameResult::Winner(Black, false);
        }
        if (reversible_plies >= 50) return GameResult::Draw(true);
        return GameResult::Ongoing();
    }

    // move_gen.rs:7-102
    std::vector<Move> possible_moves() const {
        std::vector<Move> moves;
        if (is_swapped()) {
            for (int x = 0; x < n; ++x)
                for (int y = 0; y < n; ++y)
                    if (at(x, y).is_empty()) moves.push_back(make_place(x, y, Flat));
            return moves;
        }
        for (int x = 0; x < n; ++x)
            for (int y = 0; y < n; ++y) {
                const Tile& t = at(x, y);
                if (!t.is_empty()) {
                    if (t.stack.back() == color()) add_spreads(x, y, moves);
                } else {
                    add_places(x, y, moves);
                }
            }
        return moves;
    }
    static Move make_place(int x, int y, Piece p) {
        Move m;
        m.col = uint8_t(x);
        m.row = uint8_t(y);
        m.place = true;
        m.piece = p;
        return m;
    }
    void add_places(int x, int y, std::vector<Move>& moves) const {
        uint8_t stones, caps;
        get_counts(stones, caps);
        if (stones > 0) {
            moves.push_back(make_place(x, y, Flat));
            moves.push_back(make_place(x, y, Wall));
        }
        if (caps > 0) moves.push_back(make_place(x, y, Cap));
    }
    void add_spreads(int x, int y, std::vector<Move>& moves) const {
        struct Spread {
            int col, row;
            int hand;
            std::vector<uint8_t> drops;
        };
        const Tile& tile = at(x, y);
        int max_carry = int(std::min(tile.size(), size_t(n)));
        for (Direction direction : {Up, Down, Left, Right}) {
            for (int pickup = 1; pickup <= max_carry; ++pickup) {
                std::vector<Spread> spreads;
                spreads.push_back({x, y, pickup, {}});
                while (!spreads.empty()) {
                    Spread spread = spreads.back();
                    spreads.pop_back();
                    if (spread.hand == 0) {
                        Move m;
                        m.col = uint8_t(x);
                        m.row = uint8_t(y);
                        m.place = false;
                        m.dir = direction;
                        m.drops = spread.drops;
                        moves.push_back(m);
                        continue;
                    }
                    int c = spread.col, r = spread.row;
                    if (step(c, r, direction, n)) {
                        bool can_drop = false;
                        switch (at(c, r).piece) {
                            case Flat: can_drop = true; break;
                            case Cap: can_drop = false; break;
                            case Wall: can_drop = spread.hand == 1 && tile.piece == Cap; break;
                        }
                        if (!can_drop) continue;
                        for (int drop = 1; drop <= spread.hand; ++drop) {
                            Spread s2{c, r, spread.hand - drop, spread.drops};
                            s2.drops.push_back(uint8_t(drop));
                            spreads.push_back(s2);
                        }
                    }
                }
            }
        }
    }

    // ---- TPS (tak/src/tps.rs:7-35 + takparse Tps Display) ------------------------------------------------
    std::string to_tps() const {
        std::string s;
        for (int row = n - 1; row >= 0; --row) {
            int empties = 0;
            bool first = true;
            auto flush = [&]() {
                if (empties > 0) {
                    if (!first) s += ',';
                    s += 'x';
                    if (empties > 1) s += std::to_string(empties);
                    first = false;
                    empties = 0;
                }
            };
            for (int col = 0; col < n; ++col) {
                const Tile& t = at(col, row);
                if (t.is_empty()) {
                    ++empties;
                    continue;
                }
                flush();
                if (!first) s += ',';
                first = false;
                for (auto c : t.stack) s += (c == White ? '1' : '2');
                if (t.piece == Wall) s += 'S';
                if (t.piece == Cap) s += 'C';
            }
            flush();
            if (row > 0) s += '/';
        }
        s += ' ';
        s += (to_move == White ? '1' : '2');
        s += ' ';
        s += std::to_string(1 + ply / 2);
        return s;
    }

    // tps.rs:37-96 : reserves recomputed from the board, komi / reversible plies reset
    static bool from_tps(const std::string& text, int n, Game& out) {
        Game g(n, 0);
        size_t sp1 = text.find(' ');
        if (sp1 == std::string::npos) return false;
        size_t sp2 = text.find(' ', sp1 + 1);
        if (sp2 == std::string::npos) return false;
        std::string board = text.substr(0, sp1);
        std::string color = text.substr(sp1 + 1, sp2 - sp1 - 1);
        std::string movenum = text.substr(sp2 + 1);
        int row = n - 1, col = 0;
        size_t i = 0;
        while (i < board.size()) {
            char ch = board[i];
            if (ch == '/') {
                if (col != n) return false;
                --row;
                col = 0;
                ++i;
            } else if (ch == ',') {
                ++i;
            } else if (ch == 'x') {
                ++i;
                int k = 1;
                if (i < board.size() && board[i] >= '1' && board[i] <= '8') {
                    k = board[i] - '0';
                    ++i;
                }
                col += k;
            } else if (ch == '1' || ch == '2') {
                if (row < 0 || col >= n) return false;
                Tile t;
                while (i < board.size() && (board[i] == '1' || board[i] == '2')) {
                    t.stack.push_back(board[i] == '1' ? White : Black);
                    ++i;
                }
                if (i < board.size() && board[i] == 'S') {
                    t.piece = Wall;
                    ++i;
                } else if (i < board.size() && board[i] == 'C') {
                    t.piece = Cap;
                    ++i;
                }
                g.at(col, row) = t;
                ++col;
            } else {
                return false;
            }
        }
        if (row != 0 || col != n) return false;
        g.to_move = color == "1" ? White : Black;
        int mv = std::atoi(movenum.c_str());
        if (mv < 1) return false;
        // takparse Tps::ply(): (full_move - 1) * 2 + (black ? 1 : 0)
        g.ply = uint16_t((mv - 1) * 2 + (g.to_move == Black ? 1 : 0));
        for (auto& t : g.data) {
            if (t.is_empty()) continue;
            if (t.piece == Cap) {
                if (t.stack.back() == White) {
                    g.white_stones += 1;
                    g.white_caps -= 1;
                } else {
                    g.black_stones += 1;
                    g.black_caps -= 1;
                }
            }
            for (auto c : t.stack) {
                if (c == White) g.white_stones -= 1;
                else g.black_stones -= 1;
            }
        }
        out = g;
        return true;
    }
};

// Game::from_ptn_moves (game.rs:73-82); returns status of the first failing play
inline int from_ptn_moves(int n, const std::vector<std::string>& moves, Game& out, int half_komi = 0) {
    Game g(n, half_komi);
    for (auto& s : moves) {
        Move m;
        if (!parse_move(s, n, m)) return ParseError;
        int st = g.play(m);
        if (st != Ok) return st;
    }
    out = g;
    return Ok;
}

// perf_count (tak/tests/perft.rs:3-18)
inline uint64_t perft(const Game& game, int depth) {
    if (depth == 0 || !game.result().ongoing()) return 1;
    if (depth == 1) return game.possible_moves().size();
    uint64_t total = 0;
    for (auto& m : game.possible_moves()) {
        Game clone = game;
        clone.play(m);
        total += perft(clone, depth - 1);
    }
    return total;
}

}  // namespace oracle
