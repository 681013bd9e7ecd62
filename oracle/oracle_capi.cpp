// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  C ABI over the CPU oracle so tests/ and bench.py's cpu_baseline
// leg can drive it through ctypes (oracle/liboracle.so).  See tak_oracle.hpp for the parity status.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <atomic>
#include <thread>

#include "../include/taknative.h"
#include "alphatak_oracle.hpp"

using namespace oracle;

namespace {
struct Pending {
    std::vector<size_t> path;
    Game leaf;
};
struct Search {
    Node root;
    std::deque<Pending> pending;
};

void to_state(const Game& g, tak_state_t* s) {
    std::memset(s, 0, sizeof(*s));
    s->n = uint8_t(g.n);
    s->to_move = g.to_move;
    s->ply = g.ply;
    s->white_stones = g.white_stones;
    s->white_caps = g.white_caps;
    s->black_stones = g.black_stones;
    s->black_caps = g.black_caps;
    s->half_komi = g.half_komi;
    s->reversible_plies = g.reversible_plies;
    for (int i = 0; i < g.n * g.n; ++i) {
        const Tile& t = g.data[size_t(i)];
        s->height[i] = uint8_t(t.stack.size());
        s->top[i] = t.stack.empty() ? 0 : uint8_t(t.piece);
        for (size_t k = 0; k < t.stack.size(); ++k)
            if (t.stack[k] == Black) {
                if (k < 64) s->stack_lo[i] |= (1ull << k);
                else s->stack_hi[i] |= (1ull << (k - 64));
            }
    }
}
Game from_state(const tak_state_t* s) {
    Game g(s->n, s->half_komi);
    g.to_move = Color(s->to_move);
    g.ply = s->ply;
    g.white_stones = s->white_stones;
    g.white_caps = s->white_caps;
    g.black_stones = s->black_stones;
    g.black_caps = s->black_caps;
    g.reversible_plies = s->reversible_plies;
    for (int i = 0; i < g.n * g.n; ++i) {
        Tile& t = g.data[size_t(i)];
        for (int k = 0; k < s->height[i]; ++k) {
            bool black = k < 64 ? ((s->stack_lo[i] >> k) & 1) : ((s->stack_hi[i] >> (k - 64)) & 1);
            t.stack.push_back(black ? Black : White);
        }
        t.piece = s->height[i] ? Piece(s->top[i]) : Flat;
    }
    return g;
}
}  // namespace

extern "C" {

void* orc_game_new(int n, int half_komi) { return new Game(n, half_komi); }
void orc_game_free(void* g) { delete static_cast<Game*>(g); }
void* orc_game_clone(void* g) { return new Game(*static_cast<Game*>(g)); }
int orc_game_play(void* g, uint16_t move) {
    Game* gm = static_cast<Game*>(g);
    return gm->play(Move::decode(move, gm->n));
}
int orc_game_moves(void* g, uint16_t* out, int cap) {
    Game* gm = static_cast<Game*>(g);
    auto mv = gm->possible_moves();
    for (size_t i = 0; i < mv.size() && int(i) < cap; ++i) out[i] = mv[i].encode(gm->n);
    return int(mv.size());
}
int orc_game_result(void* g) { return static_cast<Game*>(g)->result().code; }
void orc_game_get(void* g, tak_state_t* s) { to_state(*static_cast<Game*>(g), s); }
void* orc_game_from_state(const tak_state_t* s) { return new Game(from_state(s)); }
void orc_game_set_half_komi(void* g, int hk) { static_cast<Game*>(g)->half_komi = int8_t(hk); }
int orc_game_flat_diff(void* g) { return static_cast<Game*>(g)->flat_diff(); }
uint64_t orc_perft(void* g, int depth) { return perft(*static_cast<Game*>(g), depth); }

// "Play winning moves if there are any" -- the scan of train/src/self_play.rs:121-140 for ONE worker slot:
// policy = possible_moves().map(|m| (m, if clone.play(m).result() == Winner{color == inner_game.to_move} {1000} else {1}))
// Returns the number of moves; *win = whether any move wins on the spot.
int orc_selfplay_instant_win(void* g, uint16_t* out_moves, uint32_t* out_visits, int cap, int* win) {
    const Game& inner_game = *static_cast<Game*>(g);
    bool w = false;
    auto moves = inner_game.possible_moves();
    for (size_t i = 0; i < moves.size(); ++i) {
        Game clone = inner_game;
        clone.play(moves[i]);
        const GameResult r = clone.result();
        uint32_t visits = 1;  // at least one visit for all possible moves
        if (r.is_winner() && r.winner() == inner_game.to_move) {
            w = true;
            visits = 1000;    // high fake visits for winning moves
        }
        if (int(i) < cap) {
            out_moves[i] = moves[i].encode(inner_game.n);
            out_visits[i] = visits;
        }
    }
    *win = w ? 1 : 0;
    return int(moves.size());
}

// perft split over root moves on `threads` host threads (CPU-baseline leg; same counting rule)
uint64_t orc_perft_mt(void* g, int depth, int threads) {
    Game& game = *static_cast<Game*>(g);
    if (depth <= 1 || !game.result().ongoing() || threads <= 1) return perft(game, depth);
    auto moves = game.possible_moves();
    std::vector<uint64_t> counts(moves.size(), 0);
    std::vector<std::thread> pool;
    std::atomic<size_t> next{0};
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= moves.size()) break;
                Game clone = game;
                clone.play(moves[i]);
                counts[i] = perft(clone, depth - 1);
            }
        });
    for (auto& th : pool) th.join();
    uint64_t total = 0;
    for (auto c : counts) total += c;
    return total;
}

int orc_parse_move(const char* text, int n) {
    Move m;
    if (!parse_move(text, n, m)) return -1;
    return m.encode(n);
}
int orc_format_move(uint16_t move, int n, char* out, int cap) {
    std::string s = format_move(Move::decode(move, n));
    std::snprintf(out, size_t(cap), "%s", s.c_str());
    return int(s.size());
}
// Game::from_ptn_moves over a space-separated list; returns NULL and *status on failure
void* orc_game_from_ptn(int n, int half_komi, const char* moves, int* status) {
    std::vector<std::string> list;
    std::string cur;
    for (const char* p = moves; *p; ++p) {
        if (*p == ' ' || *p == ',' || *p == '\n') {
            if (!cur.empty()) list.push_back(cur);
            cur.clear();
        } else cur += *p;
    }
    if (!cur.empty()) list.push_back(cur);
    Game g(n);
    int st = from_ptn_moves(n, list, g, half_komi);
    if (status) *status = st;
    if (st != Ok) return nullptr;
    return new Game(g);
}
int orc_game_tps(void* g, char* out, int cap) {
    std::string s = static_cast<Game*>(g)->to_tps();
    std::snprintf(out, size_t(cap), "%s", s.c_str());
    return int(s.size());
}
void* orc_game_from_tps(const char* text, int n) {
    Game g(n);
    if (!Game::from_tps(text, n, g)) return nullptr;
    return new Game(g);
}

int orc_input_channels(int n) { return input_channels(n); }
int orc_board_channels(int n) { return board_channels(n); }
int orc_policy_size(int n) { return policy_size(n); }
int orc_output_size(int n) { return output_size(n); }
int orc_move_index(uint16_t move, int n) { return move_index(Move::decode(move, n), n); }
int orc_legacy_move_5(int index, char* out, int cap) {
    auto& l = legacy_moves_5();
    if (index < 0 || size_t(index) >= l.size()) return -1;
    std::string s = format_move(l[size_t(index)]);
    std::snprintf(out, size_t(cap), "%s", s.c_str());
    return int(s.size());
}
void orc_game_repr(void* g, float* out) { game_repr(*static_cast<Game*>(g), out); }
void orc_board_repr(void* g, int to_move, float* out) {
    Game* gm = static_cast<Game*>(g);
    std::fill(out, out + size_t(board_channels(gm->n)) * gm->n * gm->n, 0.0f);
    board_repr(*gm, Color(to_move), out);
}

// ---- MCTS ------------------------------------------------------------------------------------------------
void* orc_search_new() { return new Search(); }
void orc_search_free(void* s) { delete static_cast<Search*>(s); }
// Node::virtual_rollout on a clone of `game`; returns the GameResult code; Ongoing => leaf queued
int orc_search_virtual_rollout(void* s, void* game) {
    Search* se = static_cast<Search*>(s);
    Pending p;
    p.leaf = *static_cast<Game*>(game);
    GameResult r = se->root.virtual_rollout(p.leaf, p.path);
    if (r.ongoing()) se->pending.push_back(std::move(p));
    return r.code;
}
int orc_search_pending(void* s) { return int(static_cast<Search*>(s)->pending.size()); }
int orc_search_pending_state(void* s, int idx, tak_state_t* out) {
    Search* se = static_cast<Search*>(s);
    if (idx < 0 || size_t(idx) >= se->pending.size()) return -1;
    to_state(se->pending[size_t(idx)].leaf, out);
    return 0;
}
int orc_search_pending_path(void* s, int idx, int32_t* out, int cap) {
    Search* se = static_cast<Search*>(s);
    if (idx < 0 || size_t(idx) >= se->pending.size()) return -1;
    auto& p = se->pending[size_t(idx)].path;
    for (size_t i = 0; i < p.size() && int(i) < cap; ++i) out[i] = int32_t(p[i]);
    return int(p.size());
}
// Node::devirtualize_path for the oldest queued leaf
int orc_search_devirtualize(void* s, const float* policy, float eval) {
    Search* se = static_cast<Search*>(s);
    if (se->pending.empty()) return -1;
    Pending p = std::move(se->pending.front());
    se->pending.pop_front();
    se->root.devirtualize_path(p.leaf.n, p.path.data(), p.path.size(), policy, eval);
    return 0;
}
// Node::rollout with the DummyNet of alpha-tak/src/search/tests.rs:6-35 (policy = 1.0 everywhere, eval = 0)
void orc_search_rollouts_dummy(void* s, void* game, int count) {
    Search* se = static_cast<Search*>(s);
    Game* gm = static_cast<Game*>(game);
    std::vector<float> ones(size_t(policy_size(gm->n)), 1.0f);
    for (int i = 0; i < count; ++i) {
        Game clone = *gm;
        std::vector<size_t> path;
        GameResult r = se->root.virtual_rollout(clone, path);
        if (r.ongoing()) se->root.devirtualize_path(gm->n, path.data(), path.size(), ones.data(), 0.0f);
    }
}
int orc_search_children(void* s, int n, uint16_t* moves, uint32_t* visits, float* priors, float* rewards,
                        uint32_t* virtuals, int cap) {
    Search* se = static_cast<Search*>(s);
    auto& ch = se->root.children;
    for (size_t i = 0; i < ch.size() && int(i) < cap; ++i) {
        if (moves) moves[i] = ch[i].first.encode(n);
        if (visits) visits[i] = ch[i].second.visits;
        if (priors) priors[i] = ch[i].second.policy;
        if (rewards) rewards[i] = ch[i].second.expected_reward;
        if (virtuals) virtuals[i] = ch[i].second.virtual_visits;
    }
    return int(ch.size());
}
// Node::debug(depth) (debug.rs:9-24): per root child (move, visits, reward, policy, continuation), sorted by visits
// ascending (stable here; the reference's sort is unstable) and reversed.  cont_* are [cap][16].
int orc_search_debug(void* s, int n, int depth, uint16_t* moves, uint32_t* visits, float* rewards, float* policies,
                     int32_t* cont_len, uint16_t* cont_moves, uint32_t* cont_visits, int cap) {
    Search* se = static_cast<Search*>(s);
    struct Info {
        Move mov;
        uint32_t visits;
        float reward, policy;
        std::vector<std::pair<Move, uint32_t>> cont;
    };
    std::vector<Info> infos;
    for (auto& ch : se->root.children) {
        Info in{ch.first, ch.second.visits, ch.second.expected_reward, ch.second.policy, {}};
        ch.second.continuation(size_t(depth), in.cont);
        infos.push_back(std::move(in));
    }
    std::stable_sort(infos.begin(), infos.end(), [](const Info& a, const Info& b) { return a.visits < b.visits; });
    std::reverse(infos.begin(), infos.end());
    for (size_t i = 0; i < infos.size() && int(i) < cap; ++i) {
        moves[i] = infos[i].mov.encode(n);
        visits[i] = infos[i].visits;
        rewards[i] = infos[i].reward;
        policies[i] = infos[i].policy;
        cont_len[i] = int32_t(infos[i].cont.size());
        for (size_t k = 0; k < infos[i].cont.size() && k < 16; ++k) {
            cont_moves[i * 16 + k] = infos[i].cont[k].first.encode(n);
            cont_visits[i * 16 + k] = infos[i].cont[k].second;
        }
    }
    return int(infos.size());
}
void orc_search_root(void* s, uint32_t* visits, uint32_t* virtuals, float* reward) {
    Search* se = static_cast<Search*>(s);
    *visits = se->root.visits;
    *virtuals = se->root.virtual_visits;
    *reward = se->root.expected_reward;
}
int orc_search_pick(void* s, int n) {
    Search* se = static_cast<Search*>(s);
    if (se->root.children.empty()) return -1;
    return se->root.pick_move_exploit().encode(n);
}
int orc_search_play(void* s, uint16_t move, int n) {
    Search* se = static_cast<Search*>(s);
    Node next;
    if (!se->root.play(Move::decode(move, n), next)) return -1;
    se->root = std::move(next);
    se->pending.clear();
    return 0;
}
void orc_search_reset(void* s) {
    Search* se = static_cast<Search*>(s);
    se->root = Node();
    se->pending.clear();
}
void orc_search_apply_noise(void* s, const float* noise, float ratio) {
    static_cast<Search*>(s)->root.apply_noise(noise, ratio);
}
// total number of nodes in the tree (children entries incl. root) -- sizing aid for the device node pool
static size_t count_nodes(const Node& nd) {
    size_t c = 1;
    for (auto& ch : nd.children) c += count_nodes(ch.second);
    return c;
}
uint64_t orc_search_node_count(void* s) { return count_nodes(static_cast<Search*>(s)->root); }

// ---- Symmetry / Example (tak/src/symm.rs, alpha-tak/src/example.rs) ----
int orc_symmetry_move(uint16_t move, int n, int k) { return symmetries_move(Move::decode(move, n), n)[k].encode(n); }
void orc_symmetry_game(void* g, int k, tak_state_t* out) { to_state(symmetries_game(*static_cast<Game*>(g))[k], out); }
static Example make_example(void* g, const uint16_t* moves, const uint32_t* visits, int count, float result) {
    Example ex;
    ex.game = *static_cast<Game*>(g);
    ex.result = result;
    for (int i = 0; i < count; ++i) ex.policy.push_back({Move::decode(moves[i], ex.game.n), visits[i]});
    return ex;
}
void orc_example_to_tensors(void* g, const uint16_t* moves, const uint32_t* visits, int count, float result,
                            float* inputs, float* pi) {
    make_example(g, moves, visits, count, result).to_tensors(inputs, pi);
}
int orc_example_format(void* g, const uint16_t* moves, const uint32_t* visits, int count, float result, char* out,
                       int cap) {
    std::string s = make_example(g, moves, visits, count, result).to_string();
    if (int(s.size()) >= cap) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return int(s.size());
}
// returns the game handle (caller frees) or null; moves/visits/result through the out parameters
void* orc_example_parse(const char* text, int n, uint16_t* moves, uint32_t* visits, int cap, int* count, float* result) {
    Example ex;
    ex.game = Game(n, 0);
    if (!Example::from_string(text, n, ex)) return nullptr;
    if (int(ex.policy.size()) > cap) return nullptr;
    for (size_t i = 0; i < ex.policy.size(); ++i) {
        moves[i] = ex.policy[i].first.encode(n);
        visits[i] = ex.policy[i].second;
    }
    *count = int(ex.policy.size());
    *result = ex.result;
    return new Game(ex.game);
}

}  // extern "C"
