// train::self_play_parallel on the device (reference: train/src/self_play.rs:96-262): every game of the engine is a
// lock-step worker slot; per searched ply the loop is
//   forced opening (ply 0) -> instant-win shortcut -> [Dirichlet noise] -> ROLLOUTS x (virtual rollout of every game ->
//   one batched network evaluation -> devirtualise) -> pick (sample / argmax) -> replay record -> re-root + play ->
//   finished games are recorded and restarted.
// The host only sequences kernel launches; no game or tree data crosses the PCIe bus inside the loop.
#include <cmath>
#include <climits>
#include <deque>

#include "game_kernels.cuh"
#include "mcts.hpp"
#include "net.hpp"

namespace tb {

struct SpEvent {        // one finished game
    int32_t slot;
    int32_t serial;
    float white_result; // +1 white won, -1 black won, 0 draw (self_play.rs:264-275)
    int32_t plies;
};

struct SelfplayState {
    tak_selfplay_config_t cfg{};
    bool begun = false;
    int rec_cap = 0, ev_cap = 0;
    DevBuf serial;      // int [G]
    DevBuf tags;        // int [G]  per-(game, serial, ply) RNG tag
    DevBuf sample;      // u8 [G]   sample (1) or argmax (0) this ply
    DevBuf noise_on;    // u8 [G]
    DevBuf moves;       // u16 [G]
    DevBuf records;     // tak_replay_record_t [rec_cap]
    DevBuf events;      // SpEvent [ev_cap]
    DevBuf counts;      // int [4]: records, events, truncated children
    int* h_counts = nullptr;  // pinned [4]
    std::vector<std::vector<tak_replay_record_t>> held;   // per slot: records of its game that has not finished yet
    std::deque<tak_replay_record_t> ready;                // completed records not yet handed to the caller
    std::vector<tak_replay_record_t> scratch;
    uint64_t move_counter = 0;
};

static inline int warp_blocks(int warps) { return (warps + GAME_WARPS_PER_BLOCK - 1) / GAME_WARPS_PER_BLOCK; }
static inline uint64_t splitmix64_host(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// ---- kernels ---------------------------------------------------------------------------------------------------
// forced opening (self_play.rs:110-116): "a1" then "a<N>" or "<last file><N>" by a coin flip
template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_sp_opening(uint8_t* states, int n_games, const int* serial, uint64_t seed, int id_base) {
    const int w = warp_global_id();
    if (w >= n_games) return;
    constexpr int S = StateLayout<N>::S;
    WarpGame<N> g;
    g.load(states + size_t(w) * S);
    if (g.ply != 0) return;
    g.template play<false>(uint16_t(0));  // a1: row 0, col 0, flat
    const uint64_t r = splitmix64(seed ^ splitmix64((uint64_t(uint32_t(id_base + w)) << 32) | uint32_t(serial[w])));
    const int col = (r & 1) ? 0 : N - 1;
    g.template play<false>(uint16_t((N - 1) * N + col));
    g.store(states + size_t(w) * S);
}

// The lanes own squares in o-order, the ABI wants row-major order: go through shared memory.
template <int N>
__device__ __forceinline__ void export_state(const WarpGame<N>& g, tak_state_t* out, uint64_t* s_lo, uint64_t* s_hi,
                                             uint8_t* s_h) {
    const int l = threadIdx.x & 31;
#pragma unroll
    for (int half = 0; half < (WarpGame<N>::TWO ? 2 : 1); ++half) {
        const int o = l + 32 * half;
        if (o >= N * N) continue;
        const auto c = half ? g.c1 : g.c0;
        s_lo[o] = uint64_t(c);
        if constexpr (sizeof(typename WarpGame<N>::Col) == 16) s_hi[o] = uint64_t(c >> 64); else s_hi[o] = 0;
        s_h[o] = uint8_t(half ? g.h1 : g.h0);
    }
    __syncwarp();
    if (l == 0) {
        out->n = N; out->to_move = uint8_t(g.to_move); out->ply = uint16_t(g.ply);
        out->white_stones = uint8_t(g.ws); out->white_caps = uint8_t(g.wc);
        out->black_stones = uint8_t(g.bs); out->black_caps = uint8_t(g.bc);
        out->half_komi = int8_t(g.half_komi); out->reversible_plies = uint8_t(g.reversible);
        for (int i = 0; i < 6; ++i) out->_pad[i] = 0;
    }
    for (int i = l; i < 64; i += 32) {
        uint8_t h = 0, top = 0;
        uint64_t lo = 0, hi = 0;
        if (i < N * N) {
            const int row = i / N, col = i % N, o = col * N + row;
            lo = s_lo[o]; hi = s_hi[o]; h = s_h[o];
            top = h ? (((g.walls >> o) & 1) ? 1 : ((g.caps >> o) & 1) ? 2 : 0) : 0;
        }
        out->height[i] = h; out->top[i] = top; out->stack_lo[i] = lo; out->stack_hi[i] = hi;
    }
    __syncwarp();
}

struct SpView {
    uint8_t* states;
    int* serial;
    int* tags;
    uint8_t* sample;
    uint8_t* noise_on;
    uint16_t* moves;
    tak_replay_record_t* records;
    SpEvent* events;
    int* counts;  // [0] records, [1] events, [2] truncated
    int rec_cap, ev_cap;
    int n_games, id_base;
    int half_komi, exploit_ply, noise_ply, max_plies;
    uint64_t seed;
};

// "play winning moves if there are any" (self_play.rs:119-171).  The scan walks EVERY legal move (in chunks of
// TAK_REPLAY_MAX_CHILDREN through shared memory); the record keeps the first TAK_REPLAY_MAX_CHILDREN (move, visits) pairs
// and a position with more legal moves than that is counted in counts[2] (selfplay_step then fails with
// TAK_ERR_CAPACITY rather than hand out a truncated policy target).
template <int N>
__global__ void __launch_bounds__(GAME_THREADS) k_sp_instant_win(SpView sp, MctsView mv) {
    __shared__ uint64_t s_lo[GAME_WARPS_PER_BLOCK][64], s_hi[GAME_WARPS_PER_BLOCK][64];
    __shared__ uint8_t s_h[GAME_WARPS_PER_BLOCK][64];
    __shared__ uint16_t s_moves[GAME_WARPS_PER_BLOCK][TAK_REPLAY_MAX_CHILDREN];
    __shared__ uint8_t s_win[GAME_WARPS_PER_BLOCK][TAK_REPLAY_MAX_CHILDREN];
    const int w = warp_global_id();
    if (w >= sp.n_games) return;
    const int wi = (threadIdx.x >> 5);
    const int l = threadIdx.x & 31;
    constexpr int S = StateLayout<N>::S;
    constexpr int CHUNK = TAK_REPLAY_MAX_CHILDREN;
    WarpGame<N> g;
    g.load(sp.states + size_t(w) * S);
    if (g.ply < 2) return;  // fresh slot waiting for its opening (reference quirk); no move can win yet
    const int mover = g.to_move;
    bool win = false;
    int total = 0;
    // later chunks first, the record's chunk (moves 0..CHUNK-1) last so that s_moves / s_win hold it afterwards
    const int n_chunks = (g.count_total() + CHUNK - 1) / CHUNK;
    for (int c = n_chunks - 1; c >= 0; --c) {
        const int lo = c * CHUNK;
        total = g.generate([&](int k, uint16_t m) {
            if (k >= lo && k < lo + CHUNK) s_moves[wi][k - lo] = m;
        });
        __syncwarp();
        const int cnt = min(CHUNK, total - lo);
        for (int k = 0; k < cnt; ++k) {
            WarpGame<N> ch = g;
            ch.template play<false>(s_moves[wi][k]);
            const uint8_t r = ch.result();
            const bool wk = ((r & 0xF) == RES_WHITE && mover == 0) || ((r & 0xF) == RES_BLACK && mover == 1);
            if (l == 0) s_win[wi][k] = wk;
            win |= wk;
        }
        __syncwarp();
    }
    if (!win) return;
    const int n = total < CHUNK ? total : CHUNK;
    // example with 1000 fake visits on winning moves, 1 elsewhere (self_play.rs:131-140)
    int slot = 0;
    if (l == 0) slot = atomicAdd(sp.counts + 0, 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot < sp.rec_cap) {
        tak_replay_record_t* rec = sp.records + slot;
        export_state<N>(g, &rec->state, s_lo[wi], s_hi[wi], s_h[wi]);
        for (int k = l; k < n; k += 32) {
            rec->moves[k] = s_moves[wi][k];
            rec->visits[k] = s_win[wi][k] ? 1000u : 1u;
        }
        if (l == 0) {
            rec->game_id = sp.id_base + w;
            rec->game_serial = sp.serial[w];
            rec->result = nanf("");
            rec->n_children = n;
            if (total > n) atomicAdd(sp.counts + 2, 1);
        }
    }
    if (l == 0) {
        const int ev = atomicAdd(sp.counts + 1, 1);
        if (ev < sp.ev_cap) sp.events[ev] = SpEvent{w, sp.serial[w], mover == 0 ? 1.0f : -1.0f, g.ply + 1};
        sp.serial[w] += 1;
        // *node = Node::default()
        const size_t b = arena_base(mv, w, mv.half[w]);
        mv.stat[b] = make_uint4(0, 0, 0, 0);
        mv.link[b] = make_uint2(0, 0);
        mv.top[w] = 1;
        mv.pend_cnt[w] = 0;
    }
    g.reset(sp.half_komi);  // *inner_game = Game::with_komi(2): stays at ply 0 for this iteration, as in the reference
    g.store(sp.states + size_t(w) * S);
}

// per-ply flags and RNG tags
template <int N>
__global__ void k_sp_prepare(SpView sp) {
    const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= sp.n_games) return;
    const StateScalars* sc =
        reinterpret_cast<const StateScalars*>(sp.states + size_t(gidx) * StateLayout<N>::S + StateLayout<N>::SC_OFF);
    const int ply = sc->ply;
    sp.sample[gidx] = ply < sp.exploit_ply;
    sp.noise_on[gidx] = ply < sp.noise_ply;
    const uint64_t t = splitmix64((uint64_t(uint32_t(sp.id_base + gidx)) << 32) ^ (uint64_t(uint32_t(sp.serial[gidx])) << 12) ^
                                  uint64_t(ply));
    sp.tags[gidx] = int(t & 0x7FFFFFFF);
}

// IncompleteExample { game, policy: node.improved_policy() } (self_play.rs:220-223)
template <int N>
__global__ void __launch_bounds__(GAME_THREADS) k_sp_record(SpView sp, MctsView mv) {
    __shared__ uint64_t s_lo[GAME_WARPS_PER_BLOCK][64], s_hi[GAME_WARPS_PER_BLOCK][64];
    __shared__ uint8_t s_h[GAME_WARPS_PER_BLOCK][64];
    const int w = warp_global_id();
    if (w >= sp.n_games) return;
    const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
    WarpGame<N> g;
    g.load(sp.states + size_t(w) * StateLayout<N>::S);
    int slot = 0;
    if (l == 0) slot = atomicAdd(sp.counts + 0, 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot >= sp.rec_cap) return;
    tak_replay_record_t* rec = sp.records + slot;
    export_state<N>(g, &rec->state, s_lo[wi], s_hi[wi], s_h[wi]);
    const int half = mv.half[w];
    const uint4* stat = mv.stat + arena_base(mv, w, half);
    const uint2* link = mv.link + arena_base(mv, w, half);
    const uint2 lk = link[0];
    const uint32_t base = lk.x & 0xFFFFFFu;
    const int total = int(lk.y >> 16);
    const int n = total < TAK_REPLAY_MAX_CHILDREN ? total : TAK_REPLAY_MAX_CHILDREN;
    for (int k = l; k < n; k += 32) {
        rec->moves[k] = uint16_t(link[base + k].y & 0xFFFFu);
        rec->visits[k] = stat[base + k].z;
    }
    if (l == 0) {
        rec->game_id = sp.id_base + w;
        rec->game_serial = sp.serial[w];
        rec->result = nanf("");
        rec->n_children = n;
        if (total > n) atomicAdd(sp.counts + 2, 1);
    }
}

// inner_game.play(my_move); finished games are reported and the slot restarted (self_play.rs:226-258)
template <int N>
__global__ void __launch_bounds__(GAME_THREADS) k_sp_play(SpView sp, MctsView mv) {
    const int w = warp_global_id();
    if (w >= sp.n_games) return;
    const int l = threadIdx.x & 31;
    constexpr int S = StateLayout<N>::S;
    WarpGame<N> g;
    g.load(sp.states + size_t(w) * S);
    g.template play<false>(sp.moves[w]);
    if (l == 0) atomicAdd(mv.counters + 2, 1ull);   // searched plies actually played (tak_selfplay_stats_t::plies_played)
    uint8_t r = g.result();
    if (r == RES_ONGOING && sp.max_plies > 0 && g.ply >= sp.max_plies) r = RES_DRAW;  // safety cap (not in the reference)
    if (r != RES_ONGOING) {
        if (l == 0) {
            const int ev = atomicAdd(sp.counts + 1, 1);
            const float wr = (r & 0xF) == RES_WHITE ? 1.0f : (r & 0xF) == RES_BLACK ? -1.0f : 0.0f;
            if (ev < sp.ev_cap) sp.events[ev] = SpEvent{w, sp.serial[w], wr, g.ply};
            sp.serial[w] += 1;
            const size_t b = arena_base(mv, w, mv.half[w]);
            mv.stat[b] = make_uint4(0, 0, 0, 0);
            mv.link[b] = make_uint2(0, 0);
            mv.top[w] = 1;
            mv.pend_cnt[w] = 0;
        }
        g.reset(sp.half_komi);
    }
    g.store(sp.states + size_t(w) * S);
}

static SpView make_view(tak_engine* e) {
    SelfplayState& s = *e->selfplay;
    SpView v{};
    v.states = e->states.as<uint8_t>();
    v.serial = s.serial.as<int>();
    v.tags = s.tags.as<int>();
    v.sample = s.sample.as<uint8_t>();
    v.noise_on = s.noise_on.as<uint8_t>();
    v.moves = s.moves.as<uint16_t>();
    v.records = s.records.as<tak_replay_record_t>();
    v.events = s.events.as<SpEvent>();
    v.counts = s.counts.as<int>();
    v.rec_cap = s.rec_cap;
    v.ev_cap = s.ev_cap;
    v.n_games = e->max_games;
    v.id_base = s.cfg.game_id_base;
    v.half_komi = s.cfg.half_komi;
    v.exploit_ply = s.cfg.exploit_ply;
    v.noise_ply = s.cfg.noise_ply;
    v.max_plies = s.cfg.max_plies;
    v.seed = s.cfg.seed;
    return v;
}

void selfplay_destroy(tak_engine* e) {
    if (!e->selfplay) return;
    SelfplayState& s = *e->selfplay;
    for (DevBuf* b : {&s.serial, &s.tags, &s.sample, &s.noise_on, &s.moves, &s.records, &s.events, &s.counts})
        b->release();
    if (s.h_counts) cudaFreeHost(s.h_counts);
    delete e->selfplay;
    e->selfplay = nullptr;
}

template <int N>
static int step_t(tak_engine* e, int moves) {
    SelfplayState& s = *e->selfplay;
    MctsState& m = *e->mcts;
    const int G = e->max_games;
    const int wb = warp_blocks(G);
    for (int mvn = 0; mvn < moves; ++mvn, ++s.move_counter) {
        SpView sp = make_view(e);
        sp.seed = splitmix64_host(s.cfg.seed ^ s.move_counter);
        k_sp_opening<N><<<wb, GAME_THREADS, 0, e->stream>>>(sp.states, G, sp.serial, s.cfg.seed, sp.id_base);
        e->launches++;
        if (s.cfg.instant_win) {
            k_sp_instant_win<N><<<wb, GAME_THREADS, 0, e->stream>>>(sp, m.view());
            e->launches++;
        }
        k_sp_prepare<N><<<(G + 255) / 256, 256, 0, e->stream>>>(sp);
        e->launches++;
        TB_CUDA(cudaGetLastError());
        if (s.cfg.noise_ply > 0) {
            // node.rollout(game) then apply_dirichlet for plies below NOISE_PLIES (self_play.rs:174-180)
            if (int r = mcts_fast_rollouts(e, nullptr, G, 1, sp.noise_on)) return r;
            if (int r = mcts_launch_dirichlet(e, nullptr, G, sp.noise_on, s.cfg.noise_alpha, s.cfg.noise_ratio, sp.seed,
                                              sp.tags))
                return r;
        }
        if (int r = mcts_fast_rollouts(e, nullptr, G, s.cfg.rollouts, nullptr)) return r;
        if (int r = mcts_launch_pick(e, nullptr, G, sp.sample, sp.seed, sp.tags, sp.moves)) return r;
        k_sp_record<N><<<wb, GAME_THREADS, 0, e->stream>>>(sp, m.view());
        e->launches++;
        if (int r = mcts_launch_reroot(e, nullptr, sp.moves, G)) return r;
        k_sp_play<N><<<wb, GAME_THREADS, 0, e->stream>>>(sp, m.view());
        e->launches++;
        TB_CUDA(cudaGetLastError());
        if (int r = mcts_check_errors(e)) return r;
    }
    return TAK_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int32_t selfplay_begin(tak_engine_t* e, const tak_selfplay_config_t* cfg) {
    TB_CHECK(e && cfg, TAK_ERR_BAD_ARG, "selfplay_begin: null argument");
    TB_CHECK(cfg->rollouts >= 1, TAK_ERR_BAD_ARG, "rollouts must be >= 1");
    TB_CHECK(e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    TB_CHECK(e->net->arch == 0 || e->net->loaded, TAK_ERR_NO_NETWORK, "network weights not loaded");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    selfplay_destroy(e);
    SelfplayState* s = new SelfplayState();
    e->selfplay = s;
    s->cfg = *cfg;
    const int G = e->max_games;
    s->rec_cap = G * 16;
    s->ev_cap = G * 16;
    TB_CUDA(s->serial.ensure(size_t(G) * 4));
    TB_CUDA(s->tags.ensure(size_t(G) * 4));
    TB_CUDA(s->sample.ensure(size_t(G)));
    TB_CUDA(s->noise_on.ensure(size_t(G)));
    TB_CUDA(s->moves.ensure(size_t(G) * 2));
    TB_CUDA(s->records.ensure(size_t(s->rec_cap) * sizeof(tak_replay_record_t)));
    TB_CUDA(s->events.ensure(size_t(s->ev_cap) * sizeof(SpEvent)));
    TB_CUDA(s->counts.ensure(16));
    TB_CUDA(cudaMemsetAsync(s->serial.p, 0, size_t(G) * 4, e->stream));
    TB_CUDA(cudaMemsetAsync(s->counts.p, 0, 16, e->stream));
    TB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_counts), 16));
    // reserved[0] != 0 ("keep_positions"): the games keep the positions they hold (e.g. mid-game positions uploaded or
    // produced by tak_playouts) instead of starting from the empty board; slots that finish restart as usual
    if (!cfg->reserved[0])
        if (int r = tak_games_reset(e, 0, G, cfg->half_komi)) return r;
    if (int r = mcts_launch_tree_reset(e, nullptr, G)) return r;
    TB_CUDA(cudaMemsetAsync(e->mcts->counters.p, 0, 64, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    s->begun = true;
    return TAK_OK;
}

int32_t selfplay_step(tak_engine_t* e, int32_t moves, tak_selfplay_stats_t* out_stats) {
    TB_CHECK(e && moves >= 1, TAK_ERR_BAD_ARG, "selfplay_step: bad argument");
    TB_CHECK(e->selfplay && e->selfplay->begun, TAK_ERR_BAD_ARG, "selfplay_step before selfplay_begin");
    TB_CUDA(cudaSetDevice(e->device));
    SelfplayState& s = *e->selfplay;
    const int G = e->max_games;
    // room for this call's records (one per game per ply plus instant wins)?
    TB_CUDA(cudaMemcpyAsync(s.h_counts, s.counts.p, 16, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    TB_CHECK(s.h_counts[0] + 2 * G * moves <= s.rec_cap, TAK_ERR_CAPACITY,
             "replay buffer would overflow: call selfplay_drain (holds %d of %d records)", s.h_counts[0], s.rec_cap);
    unsigned long long c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
    TB_CUDA(cudaMemcpy(c0, e->mcts->counters.p, 24, cudaMemcpyDeviceToHost));
    const int ev0 = s.h_counts[1], rec0 = s.h_counts[0];
    const uint64_t launches0 = e->launches;
    cudaEvent_t t0, t1;
    TB_CUDA(cudaEventCreate(&t0));
    TB_CUDA(cudaEventCreate(&t1));
    TB_CUDA(cudaEventRecord(t0, e->stream));
    int r = TAK_ERR_BAD_ARG;
    TB_DISPATCH_N(e->n, r = step_t<N_>(e, moves));
    if (r == TAK_OK) {
        cudaEventRecord(t1, e->stream);
        cudaError_t ce = cudaEventSynchronize(t1);
        if (ce != cudaSuccess) {
            set_error("selfplay_step: %s", cudaGetErrorString(ce));
            r = TAK_ERR_CUDA;
        }
    }
    if (r == TAK_OK && out_stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        cudaMemcpy(c1, e->mcts->counters.p, 24, cudaMemcpyDeviceToHost);
        cudaMemcpy(s.h_counts, s.counts.p, 16, cudaMemcpyDeviceToHost);
        std::memset(out_stats, 0, sizeof(*out_stats));
        out_stats->plies_played = c1[2] - c0[2];          // counted on the device by k_sp_play
        out_stats->games_completed = uint64_t(s.h_counts[1] - ev0);
        out_stats->rollouts = c1[0] - c0[0];
        out_stats->evals = c1[1] - c0[1];
        out_stats->kernel_launches = e->launches - launches0;
        out_stats->records = uint64_t(s.h_counts[0] - rec0);
        out_stats->device_ms = ms;
        out_stats->records_truncated = uint64_t(s.h_counts[2]);
    }
    if (r == TAK_OK) {
        if (!out_stats) cudaMemcpy(s.h_counts, s.counts.p, 16, cudaMemcpyDeviceToHost);
        if (s.h_counts[2] != 0) {
            set_error("%d position(s) had more than %d legal moves: their replay records would be truncated",
                      s.h_counts[2], TAK_REPLAY_MAX_CHILDREN);
            r = TAK_ERR_CAPACITY;
        }
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return r;
}

int32_t selfplay_drain(tak_engine_t* e, tak_replay_record_t* out, int32_t cap, int32_t* out_count) {
    TB_CHECK(e && out_count && (out || cap == 0) && cap >= 0, TAK_ERR_BAD_ARG, "selfplay_drain: bad argument");
    TB_CHECK(e->selfplay && e->selfplay->begun, TAK_ERR_BAD_ARG, "selfplay_drain before selfplay_begin");
    TB_CUDA(cudaSetDevice(e->device));
    SelfplayState& s = *e->selfplay;
    const int G = e->max_games;
    // 1. move the device ring to the host: records into their slot's bucket, finish events into a list
    TB_CUDA(cudaMemcpy(s.h_counts, s.counts.p, 16, cudaMemcpyDeviceToHost));
    const int nrec = std::min(s.h_counts[0], s.rec_cap), nev = std::min(s.h_counts[1], s.ev_cap);
    if (s.held.size() != size_t(G)) s.held.resize(size_t(G));
    const int base = s.cfg.game_id_base;
    if (nrec) {
        s.scratch.resize(size_t(nrec));
        TB_CUDA(cudaMemcpy(s.scratch.data(), s.records.p, size_t(nrec) * sizeof(tak_replay_record_t),
                           cudaMemcpyDeviceToHost));
        for (const auto& rec : s.scratch) {
            const int slot = rec.game_id - base;
            TB_CHECK(slot >= 0 && slot < G, TAK_ERR_CUDA, "internal: replay record of slot %d", slot);
            s.held[size_t(slot)].push_back(rec);
        }
    }
    std::vector<SpEvent> events(static_cast<size_t>(nev));
    if (nev) TB_CUDA(cudaMemcpy(events.data(), s.events.p, size_t(nev) * sizeof(SpEvent), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMemset(s.counts.p, 0, 8));
    // 2. join: a record is complete once its (slot, serial) has a finish event (Example::complete, example.rs:19-25);
    //    the cost is proportional to the records of the games that finished, not to everything still held
    for (const auto& ev : events) {
        auto& bucket = s.held[size_t(ev.slot)];
        size_t keep = 0;
        for (size_t i = 0; i < bucket.size(); ++i) {
            if (bucket[i].game_serial == ev.serial) {
                bucket[i].result = bucket[i].state.to_move == 0 ? ev.white_result : -ev.white_result;
                s.ready.push_back(bucket[i]);
            } else {
                if (keep != i) bucket[keep] = bucket[i];
                ++keep;
            }
        }
        bucket.resize(keep);
    }
    // 3. hand out completed records (all of them stay queued when the caller only asks for the count)
    if (!out) {
        *out_count = int(std::min<size_t>(s.ready.size(), size_t(INT32_MAX)));
        return TAK_OK;
    }
    int produced = 0;
    while (produced < cap && !s.ready.empty()) {
        out[produced++] = s.ready.front();
        s.ready.pop_front();
    }
    *out_count = produced;
    return TAK_OK;
}

}  // extern "C"
