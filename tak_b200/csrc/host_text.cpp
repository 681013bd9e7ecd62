// Host-side (cold) helpers of the C ABI: PTN move text, TPS position text, policy indexing.
// These replace the takparse 0.5.5 surface the reference re-exports (tak/src/lib.rs:15) and
// alpha_tak::search::move_index (alpha-tak/src/search/move_map.rs:19-48) / From<Game> for Tps
// (tak/src/tps.rs:7-96).  All of it works on the u16 move encoding and the POD tak_state_t.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <mutex>

#include "engine.hpp"

namespace tb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline int stones_for(int n) { return n == 3 ? 10 : n == 4 ? 15 : n == 5 ? 21 : n == 6 ? 30 : n == 7 ? 40 : 50; }
static inline int caps_for(int n) { return n <= 4 ? 0 : n <= 6 ? 1 : 2; }

int host_policy_size(int n) {
    if (n == 5) return 1575;  // repr/moves.rs:6-16 (legacy one-hot list)
    return n * n * (3 + 4 * ((1 << n) - 2));
}

// u16 spread helpers
static inline int mask_pieces(unsigned mask) { return 8 - __builtin_ctz(mask); }

// Index of a move in the 5x5 legacy list, computed in closed form from the list's generation rule:
// 75 placements (col, row, piece), then (col, row, dir in [<,-,>,+] with distance > 0, pickup 1..5,
// patterns with at most `distance` drops in ascending mask order).
static int legacy5_index(uint16_t mv) {
    const int n = 5;
    int sq = mv & 63, row = sq / n, col = sq % n;
    unsigned mask = mv >> 8;
    int kind = (mv >> 6) & 3;
    if (sq >= 25) return -1;
    if (mask == 0) return kind > 2 ? -1 : (col * n + row) * 3 + kind;
    // number of patterns of `p` pieces with at most `dist` drops
    auto npat = [](int p, int dist) {
        int c = 0;
        for (unsigned v = 1; v < (1u << p); v += 2) c += __builtin_popcount(v) <= dist;
        return c;
    };
    auto dist_of = [&](int c, int r, int d) {  // d in legacy order: 0 '<', 1 '-', 2 '>', 3 '+'
        return d == 0 ? c : d == 1 ? r : d == 2 ? n - 1 - c : n - 1 - r;
    };
    static const int abi_to_legacy[4] = {3, 1, 0, 2};  // ABI dir 0 Up 1 Down 2 Left 3 Right
    int ld = abi_to_legacy[kind];
    int p = mask_pieces(mask);
    if (p > n) return -1;
    unsigned v = mask >> (8 - p);
    if (__builtin_popcount(v) > dist_of(col, row, ld)) return -1;
    int idx = 75;
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r)
            for (int d = 0; d < 4; ++d) {
                int dist = dist_of(c, r, d);
                bool here = (c == col && r == row && d == ld);
                for (int q = 1; q <= n; ++q) {
                    if (here && q == p) {
                        for (unsigned u = 1; u < v; u += 2) idx += __builtin_popcount(u) <= dist;
                        return idx;
                    }
                    idx += dist ? npat(q, dist) : 0;
                }
            }
    return -1;
}

int host_move_index(int n, uint16_t mv) {
    if (n == 5) return legacy5_index(mv);
    int sq = mv & 63, row = sq / n, col = sq % n;
    if (sq >= n * n) return -1;
    unsigned mask = mv >> 8;
    int kind = (mv >> 6) & 3;
    int channel;
    if (mask == 0) {
        if (kind > 2) return -1;
        channel = kind;
    } else {
        int p = mask_pieces(mask);
        if (p > n) return -1;
        int pattern_offset = int(mask >> (8 - n)) - 1;
        if (pattern_offset >= (1 << n) - 2) return -1;
        static const int dir_slot[4] = {0, 2, 3, 1};  // Up 0, Right 1, Down 2, Left 3 (move_map.rs:37-42)
        channel = 3 + pattern_offset + ((1 << n) - 2) * dir_slot[kind];
    }
    return channel * n * n + row * n + col;
}

const std::vector<uint16_t>& host_move_index_table(int n) {
    static std::vector<uint16_t> tables[9];
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto& t = tables[n];
    if (t.empty()) {
        t.assign(65536, 0xFFFF);
        for (unsigned mv = 0; mv < 65536; ++mv) {
            int idx = host_move_index(n, uint16_t(mv));
            if (idx >= 0) t[mv] = uint16_t(idx);
        }
    }
    return t;
}

}  // namespace tb

using namespace tb;

extern "C" {

const char* tak_last_error(void) { return g_err; }
int32_t tak_version(void) { return 100; }

int32_t tak_move_index(int32_t n, uint16_t move, int32_t* out_index) {
    TB_CHECK(n >= 3 && n <= 8 && out_index, TAK_ERR_BAD_ARG, "tak_move_index: bad argument");
    int idx = host_move_index(n, move);
    TB_CHECK(idx >= 0, TAK_ERR_INVALID_MOVE, "could not map move 0x%04x to an index", move);
    *out_index = idx;
    return TAK_OK;
}
int32_t tak_policy_size(int32_t n, int32_t* out_size) {
    TB_CHECK(n >= 3 && n <= 8 && out_size, TAK_ERR_BAD_ARG, "tak_policy_size: bad argument");
    *out_size = host_policy_size(n);
    return TAK_OK;
}

// PTN: [FSC]?<file><rank>  |  [count]?<file><rank><dir>[drops]?
int32_t tak_ptn_parse(int32_t n, const char* text, uint16_t* out_move) {
    TB_CHECK(n >= 3 && n <= 8 && text && out_move, TAK_ERR_BAD_ARG, "tak_ptn_parse: bad argument");
    const char* p = text;
    int count = 0, piece = -1;
    if (*p >= '1' && *p <= '8') count = *p++ - '0';
    else if (*p == 'F') { piece = 0; ++p; }
    else if (*p == 'S') { piece = 1; ++p; }
    else if (*p == 'C') { piece = 2; ++p; }
    TB_CHECK(*p >= 'a' && *p < 'a' + n, TAK_ERR_PARSE, "bad file in '%s'", text);
    int col = *p++ - 'a';
    TB_CHECK(*p >= '1' && *p < '1' + n, TAK_ERR_PARSE, "bad rank in '%s'", text);
    int row = *p++ - '1';
    unsigned sq = unsigned(row * n + col);
    int dir = *p == '+' ? 0 : *p == '-' ? 1 : *p == '<' ? 2 : *p == '>' ? 3 : -1;
    if (dir < 0) {
        while (*p == '\'' || *p == '!' || *p == '?') ++p;
        TB_CHECK(*p == 0 && count == 0, TAK_ERR_PARSE, "trailing text in '%s'", text);
        *out_move = uint16_t(sq | (unsigned(piece < 0 ? 0 : piece) << 6));
        return TAK_OK;
    }
    TB_CHECK(piece < 0, TAK_ERR_PARSE, "piece prefix on a spread in '%s'", text);
    ++p;
    if (count == 0) count = 1;
    unsigned mask = 0;
    int pos = 7, sum = 0;
    bool any = false;
    for (; *p >= '1' && *p <= '8'; ++p) {
        int d = *p - '0';
        sum += d;
        TB_CHECK(sum <= 8, TAK_ERR_PARSE, "too many pieces in '%s'", text);
        pos -= d - 1;
        mask |= 1u << pos;
        pos -= 1;
        any = true;
    }
    while (*p == '\'' || *p == '!' || *p == '?' || *p == '*') ++p;
    TB_CHECK(*p == 0, TAK_ERR_PARSE, "trailing text in '%s'", text);
    if (!any) { mask = 1u << (8 - count); sum = count; }
    TB_CHECK(sum == count, TAK_ERR_PARSE, "drop counts do not add up in '%s'", text);
    *out_move = uint16_t(sq | (unsigned(dir) << 6) | (mask << 8));
    return TAK_OK;
}

int32_t tak_ptn_format(int32_t n, uint16_t move, char* out, int32_t cap) {
    TB_CHECK(n >= 3 && n <= 8 && out && cap >= 16, TAK_ERR_BAD_ARG, "tak_ptn_format: bad argument");
    int sq = move & 63, row = sq / n, col = sq % n, kind = (move >> 6) & 3;
    unsigned mask = move >> 8;
    TB_CHECK(sq < n * n, TAK_ERR_INVALID_MOVE, "square out of range");
    char* w = out;
    if (mask == 0) {
        if (kind == 1) *w++ = 'S';
        if (kind == 2) *w++ = 'C';
        *w++ = char('a' + col);
        *w++ = char('1' + row);
    } else {
        int pieces = mask_pieces(mask);
        if (pieces > 1) *w++ = char('0' + pieces);
        *w++ = char('a' + col);
        *w++ = char('1' + row);
        *w++ = "+-<>"[kind];
        if (__builtin_popcount(mask) > 1) {
            int run = 0;
            for (int b = 7; b >= 8 - pieces; --b) {
                ++run;
                if (mask & (1u << b)) { *w++ = char('0' + run); run = 0; }
            }
        }
    }
    *w = 0;
    return TAK_OK;
}

int32_t tak_state_init(int32_t n, int32_t half_komi, tak_state_t* out) {
    TB_CHECK(n >= 3 && n <= 8 && out, TAK_ERR_BAD_ARG, "tak_state_init: bad argument");
    std::memset(out, 0, sizeof(*out));
    out->n = uint8_t(n);
    out->white_stones = out->black_stones = uint8_t(stones_for(n));
    out->white_caps = out->black_caps = uint8_t(caps_for(n));
    out->half_komi = int8_t(half_komi);
    return TAK_OK;
}

int32_t tak_tps_format(const tak_state_t* s, char* out, int32_t cap) {
    TB_CHECK(s && out && cap > 0, TAK_ERR_BAD_ARG, "tak_tps_format: bad argument");
    std::string t;
    int n = s->n;
    for (int row = n - 1; row >= 0; --row) {
        int run = 0;
        std::string line;
        auto sep = [&]() { if (!line.empty()) line += ','; };
        auto flush = [&]() {
            if (!run) return;
            sep();
            line += 'x';
            if (run > 1) line += char('0' + run);
            run = 0;
        };
        for (int col = 0; col < n; ++col) {
            int i = row * n + col;
            if (s->height[i] == 0) { ++run; continue; }
            flush();
            sep();
            for (int k = 0; k < s->height[i]; ++k) {
                bool black = k < 64 ? (s->stack_lo[i] >> k) & 1 : (s->stack_hi[i] >> (k - 64)) & 1;
                line += black ? '2' : '1';
            }
            if (s->top[i] == 1) line += 'S';
            if (s->top[i] == 2) line += 'C';
        }
        flush();
        t += line;
        if (row) t += '/';
    }
    t += s->to_move ? " 2 " : " 1 ";
    t += std::to_string(1 + s->ply / 2);
    TB_CHECK(int(t.size()) < cap, TAK_ERR_CAPACITY, "tak_tps_format: buffer too small");
    std::memcpy(out, t.c_str(), t.size() + 1);
    return TAK_OK;
}

int32_t tak_tps_parse(int32_t n, const char* text, tak_state_t* out) {
    TB_CHECK(n >= 3 && n <= 8 && text && out, TAK_ERR_BAD_ARG, "tak_tps_parse: bad argument");
    tak_state_init(n, 0, out);
    const char* p = text;
    int row = n - 1, col = 0;
    while (*p && *p != ' ') {
        if (*p == '/') { TB_CHECK(col == n && row > 0, TAK_ERR_PARSE, "bad TPS row"); --row; col = 0; ++p; }
        else if (*p == ',') ++p;
        else if (*p == 'x') {
            ++p;
            int k = 1;
            if (*p >= '1' && *p <= '8') k = *p++ - '0';
            col += k;
        } else if (*p == '1' || *p == '2') {
            TB_CHECK(col < n, TAK_ERR_PARSE, "TPS row too long");
            int i = row * n + col, h = 0;
            for (; *p == '1' || *p == '2'; ++p, ++h) {
                TB_CHECK(h < 128, TAK_ERR_PARSE, "stack too tall");
                if (*p == '2') { if (h < 64) out->stack_lo[i] |= 1ull << h; else out->stack_hi[i] |= 1ull << (h - 64); }
            }
            out->height[i] = uint8_t(h);
            if (*p == 'S') { out->top[i] = 1; ++p; }
            else if (*p == 'C') { out->top[i] = 2; ++p; }
            ++col;
        } else TB_CHECK(false, TAK_ERR_PARSE, "unexpected '%c' in TPS", *p);
    }
    TB_CHECK(row == 0 && col == n && *p == ' ', TAK_ERR_PARSE, "bad TPS board");
    ++p;
    TB_CHECK(*p == '1' || *p == '2', TAK_ERR_PARSE, "bad TPS colour");
    out->to_move = *p == '2';
    ++p;
    TB_CHECK(*p == ' ', TAK_ERR_PARSE, "bad TPS move number");
    // exactly three space-separated segments (takparse's Tps::from_str refuses a wrong segment count); the move number
    // is all digits; trailing blanks are tolerated
    ++p;
    TB_CHECK(*p >= '0' && *p <= '9', TAK_ERR_PARSE, "bad TPS move number");
    long mv = 0;
    while (*p >= '0' && *p <= '9' && mv < 100000) mv = mv * 10 + (*p++ - '0');
    while (*p == ' ') ++p;
    TB_CHECK(*p == 0, TAK_ERR_PARSE, "unexpected text after the TPS move number");
    TB_CHECK(mv >= 1 && mv <= 30000, TAK_ERR_PARSE, "bad TPS move number");
    out->ply = uint16_t((mv - 1) * 2 + out->to_move);
    // reserves are inferred from the board (tps.rs:63-84)
    int ws = stones_for(n), wc = caps_for(n), bs = ws, bc = wc;
    for (int i = 0; i < n * n; ++i) {
        int h = out->height[i];
        if (!h) continue;
        auto black_at = [&](int k) { return k < 64 ? (out->stack_lo[i] >> k) & 1 : (out->stack_hi[i] >> (k - 64)) & 1; };
        if (out->top[i] == 2) {
            if (black_at(h - 1)) { bs += 1; bc -= 1; } else { ws += 1; wc -= 1; }
        }
        for (int k = 0; k < h; ++k) { if (black_at(k)) bs -= 1; else ws -= 1; }
    }
    out->white_stones = uint8_t(ws); out->white_caps = uint8_t(wc);
    out->black_stones = uint8_t(bs); out->black_caps = uint8_t(bc);
    return TAK_OK;
}

// ---- alpha_tak::Example text format (alpha-tak/src/example.rs:81-133) ---------------------------------------
// Rust's `{}` for f32: shortest decimal that round-trips, no exponent for these magnitudes, "NaN" / "inf"
static std::string rust_f32(float v) {
    if (v != v) return "NaN";
    if (v == INFINITY) return "inf";
    if (v == -INFINITY) return "-inf";
    char buf[64];
    for (int prec = 1; prec <= 9; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, double(v));
        if (strtof(buf, nullptr) == v) break;
    }
    std::string t = buf;
    if (t.find('e') != std::string::npos) {  // out of the range self-play produces; fall back to fixed notation
        snprintf(buf, sizeof buf, "%.9f", double(v));
        t = buf;
        while (!t.empty() && t.back() == '0') t.pop_back();
        if (!t.empty() && t.back() == '.') t.pop_back();
    }
    return t;
}

int32_t tak_example_format(const tak_replay_record_t* rec, char* out, int32_t cap) {
    TB_CHECK(rec && out && cap > 0, TAK_ERR_BAD_ARG, "tak_example_format: bad argument");
    TB_CHECK(rec->n_children >= 0 && rec->n_children <= TAK_REPLAY_MAX_CHILDREN, TAK_ERR_BAD_ARG,
             "tak_example_format: bad child count");
    char tps[1024];
    if (int r = tak_tps_format(&rec->state, tps, sizeof tps)) return r;
    const tak_state_t& g = rec->state;
    std::string t = tps;
    t += ';' + std::to_string(int(g.white_stones)) + ';' + std::to_string(int(g.white_caps)) + ';' +
         std::to_string(int(g.black_stones)) + ';' + std::to_string(int(g.black_caps)) + ';' +
         std::to_string(int(g.half_komi)) + ';' + rust_f32(rec->result) + ';';
    for (int i = 0; i < rec->n_children; ++i) {
        char mv[32];
        if (int r = tak_ptn_format(g.n, rec->moves[i], mv, sizeof mv)) return r;
        if (i) t += ',';
        t += mv;
        t += ':' + std::to_string(rec->visits[i]);
    }
    TB_CHECK(int(t.size()) < cap, TAK_ERR_CAPACITY, "tak_example_format: buffer too small");
    std::memcpy(out, t.c_str(), t.size() + 1);
    return TAK_OK;
}

int32_t tak_example_parse(int32_t n, const char* text, tak_replay_record_t* out) {
    TB_CHECK(n >= 3 && n <= 8 && text && out, TAK_ERR_BAD_ARG, "tak_example_parse: bad argument");
    std::memset(out, 0, sizeof *out);
    std::string s = text;
    while (!s.empty() && isspace(static_cast<unsigned char>(s.back()))) s.pop_back();
    size_t b = 0;
    while (b < s.size() && isspace(static_cast<unsigned char>(s[b]))) ++b;
    s = s.substr(b);
    std::vector<std::string> f;
    for (size_t p = 0;;) {
        const size_t q = s.find(';', p);
        f.push_back(s.substr(p, q == std::string::npos ? q : q - p));
        if (q == std::string::npos) break;
        p = q + 1;
    }
    static const char* names[] = {"tps", "white stones", "white caps", "black stones", "black caps", "half komi",
                                  "result", "policy"};
    TB_CHECK(f.size() >= 8, TAK_ERR_PARSE, "missing %s", names[f.size()]);
    if (int r = tak_tps_parse(n, f[0].c_str(), &out->state)) return r;
    auto num = [&](const std::string& t, long lo, long hi, long* v) {
        char* end = nullptr;
        *v = strtol(t.c_str(), &end, 10);
        return !t.empty() && *end == 0 && *v >= lo && *v <= hi;
    };
    long v;
    TB_CHECK(num(f[1], 0, 255, &v), TAK_ERR_PARSE, "bad white stones"); out->state.white_stones = uint8_t(v);
    TB_CHECK(num(f[2], 0, 255, &v), TAK_ERR_PARSE, "bad white caps");   out->state.white_caps = uint8_t(v);
    TB_CHECK(num(f[3], 0, 255, &v), TAK_ERR_PARSE, "bad black stones"); out->state.black_stones = uint8_t(v);
    TB_CHECK(num(f[4], 0, 255, &v), TAK_ERR_PARSE, "bad black caps");   out->state.black_caps = uint8_t(v);
    TB_CHECK(num(f[5], -128, 127, &v), TAK_ERR_PARSE, "bad half komi"); out->state.half_komi = int8_t(v);
    {
        char* end = nullptr;
        out->result = strtof(f[6].c_str(), &end);
        TB_CHECK(!f[6].empty() && *end == 0, TAK_ERR_PARSE, "bad result");
    }
    int k = 0;
    for (size_t p = 0; p <= f[7].size();) {
        size_t q = f[7].find(',', p);
        if (q == std::string::npos) q = f[7].size();
        const std::string pair = f[7].substr(p, q - p);
        const size_t c = pair.find(':');
        TB_CHECK(c != std::string::npos, TAK_ERR_PARSE, "pair has missing delimiter");
        TB_CHECK(k < TAK_REPLAY_MAX_CHILDREN, TAK_ERR_CAPACITY, "more than %d policy entries", TAK_REPLAY_MAX_CHILDREN);
        if (int r = tak_ptn_parse(n, pair.substr(0, c).c_str(), &out->moves[k])) return r;
        long long vis = 0;
        {
            const std::string t = pair.substr(c + 1);
            char* end = nullptr;
            vis = strtoll(t.c_str(), &end, 10);
            TB_CHECK(!t.empty() && *end == 0 && vis >= 0 && vis <= 0xFFFFFFFFll, TAK_ERR_PARSE, "bad visit count");
        }
        out->visits[k++] = uint32_t(vis);
        p = q + 1;
    }
    out->n_children = k;
    return TAK_OK;
}

}  // extern "C"
