// Engine lifetime + the tak::Game part of the C ABI (reset / upload / download / possible_moves / play / result /
// perft).  Reference items replaced: tak/src/game.rs:37-130,220-267, tak/src/move_gen.rs:7-102,
// tak/tests/perft.rs:3-18 (counting rule).
#include <cub/device/device_scan.cuh>

#include <cstdlib>

#include "engine.hpp"
#include "game_kernels.cuh"

namespace tb {

int state_bytes_for(int n) {
    int s = 0;
    TB_DISPATCH_N(n, s = StateLayout<N_>::S);
    return s;
}

template <int N>
static void pack_t(const tak_state_t& s, uint8_t* rec) {
    using L = StateLayout<N>;
    using Col = typename L::Col;
    std::memset(rec, 0, L::S);
    Col* cols = reinterpret_cast<Col*>(rec);
    uint8_t* hts = rec + L::HTS_OFF;
    uint64_t walls = 0, caps = 0, occ = 0, blk = 0;
    for (int row = 0; row < N; ++row)
        for (int col = 0; col < N; ++col) {
            const int i = row * N + col, o = col * N + row;
            Col c = Col(s.stack_lo[i]);
            if constexpr (sizeof(Col) == 16) c |= Col(s.stack_hi[i]) << 64;
            cols[o] = c;
            hts[o] = s.height[i];
            if (s.height[i] && s.top[i] == 1) walls |= 1ull << o;
            if (s.height[i] && s.top[i] == 2) caps |= 1ull << o;
            if (s.height[i]) {
                occ |= 1ull << o;
                if ((c >> (s.height[i] - 1)) & 1) blk |= 1ull << o;
            }
        }
    uint64_t* bb = reinterpret_cast<uint64_t*>(rec + L::BB_OFF);
    bb[0] = walls;
    bb[1] = caps;
    uint64_t* der = reinterpret_cast<uint64_t*>(rec + L::DER_OFF);   // derived bitboards (tak_device.cuh)
    der[0] = occ;
    der[1] = blk;
    StateScalars sc{};
    sc.to_move = s.to_move; sc.ply = s.ply;
    sc.ws = s.white_stones; sc.wc = s.white_caps; sc.bs = s.black_stones; sc.bc = s.black_caps;
    sc.half_komi = s.half_komi; sc.reversible = s.reversible_plies;
    std::memcpy(rec + L::SC_OFF, &sc, 16);
}
template <int N>
static void unpack_t(const uint8_t* rec, tak_state_t& s) {
    using L = StateLayout<N>;
    using Col = typename L::Col;
    std::memset(&s, 0, sizeof(s));
    const Col* cols = reinterpret_cast<const Col*>(rec);
    const uint8_t* hts = rec + L::HTS_OFF;
    const uint64_t* bb = reinterpret_cast<const uint64_t*>(rec + L::BB_OFF);
    StateScalars sc;
    std::memcpy(&sc, rec + L::SC_OFF, 16);
    s.n = N; s.to_move = sc.to_move; s.ply = sc.ply;
    s.white_stones = sc.ws; s.white_caps = sc.wc; s.black_stones = sc.bs; s.black_caps = sc.bc;
    s.half_komi = sc.half_komi; s.reversible_plies = sc.reversible;
    for (int row = 0; row < N; ++row)
        for (int col = 0; col < N; ++col) {
            const int i = row * N + col, o = col * N + row;
            Col c;
            std::memcpy(&c, &cols[o], sizeof(Col));
            s.stack_lo[i] = uint64_t(c);
            if constexpr (sizeof(Col) == 16) s.stack_hi[i] = uint64_t(c >> 64);
            s.height[i] = hts[o];
            s.top[i] = hts[o] ? (((bb[0] >> o) & 1) ? 1 : ((bb[1] >> o) & 1) ? 2 : 0) : 0;
        }
}
void pack_state(int n, const tak_state_t& s, uint8_t* rec) { TB_DISPATCH_N(n, pack_t<N_>(s, rec)); }
void unpack_state(int n, const uint8_t* rec, tak_state_t& s) { TB_DISPATCH_N(n, unpack_t<N_>(rec, s)); }

static inline int warp_blocks(int warps) { return (warps + GAME_WARPS_PER_BLOCK - 1) / GAME_WARPS_PER_BLOCK; }

}  // namespace tb

using namespace tb;

static int check_ids(tak_engine_t* e, const int32_t* ids, int32_t n) {
    TB_CHECK(e && ids && n >= 0, TAK_ERR_BAD_ARG, "null engine/ids or negative count");
    for (int i = 0; i < n; ++i)
        TB_CHECK(ids[i] >= 0 && ids[i] < e->max_games, TAK_ERR_BAD_ARG, "game id %d out of range [0,%d)", ids[i],
                 e->max_games);
    return TAK_OK;
}

// ---- perft: breadth-first over packed frontiers ---------------------------------------------------------------
// Roots -> count kernel (result + move count per root).  Then per level: exclusive scan of the parents' child counts ->
// k_perft_moves (move lists + block map) -> k_perft_apply (children written in move-generation order AND classified /
// counted in the same pass, game_kernels.cuh).  The last level's counts are perf_count's answer (perft.rs:6-7), finished
// games count 1 (perft.rs:4).  When a frontier's children exceed PF_CAP states, the parents are cut into slices by
// binary search on the scanned offsets and each slice is expanded and recursed into separately, so memory is bounded
// by PF_CAP states per level.
static size_t pf_cap() {   // states per level; TAK_PERFT_CAP overrides it (the tests force the slicing path with it)
    static const size_t cap = [] {
        const char* s = std::getenv("TAK_PERFT_CAP");
        const long long v = s ? std::atoll(s) : 0;
        return v >= 4096 ? size_t(v) : size_t(1) << 24;
    }();
    return cap;
}

// `frontier` holds n parents at `depth_left` >= 2 plies above the counted level; counts[i] = children of parent i
static int perft_variant() {
    static const int v = [] {
        const char* s = std::getenv("TAK_PERFT_VARIANT");
        const int x = s ? std::atoi(s) : 3;
        return x == 0 || x == 1 || x == 5 ? x : 3;
    }();
    return v;
}

template <int N, int V>
static int perft_level(tak_engine* e, const uint8_t* frontier, size_t n, const uint32_t* d_counts, int depth_left,
                       int level, unsigned long long* d_leaves) {
    using P = PerftCfg<N, V>;
    constexpr int S = StateLayout<N>::S;
    TB_CHECK(level < tak_engine::PF_LEVELS, TAK_ERR_BAD_ARG, "perft too deep");
    TB_CHECK(n < (size_t(1) << 31), TAK_ERR_CAPACITY, "perft frontier too large");
    tak_engine::PerftLevel& lv = e->pf_level[level];
    TB_CUDA(lv.offsets.ensure(n * 8));
    uint64_t* d_off = lv.offsets.as<uint64_t>();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_off, int(n), e->stream);
    TB_CUDA(e->pf_scan_tmp.ensure(tmp_bytes + 16));
    TB_CUDA(cub::DeviceScan::ExclusiveSum(e->pf_scan_tmp.p, tmp_bytes, d_counts, d_off, int(n), e->stream));
    e->pf_launches += 2;
    TB_CUDA(e->ensure_pinned(64));
    auto offset_at = [&](size_t i, uint64_t* out) -> int {  // offsets[i], with offsets[n] = total
        uint64_t* h = static_cast<uint64_t*>(e->h_stage);
        uint32_t* hc = reinterpret_cast<uint32_t*>(h + 1);
        const size_t at = i < n ? i : n - 1;
        TB_CUDA(cudaMemcpyAsync(h, d_off + at, 8, cudaMemcpyDeviceToHost, e->stream));
        *hc = 0;
        if (i >= n) TB_CUDA(cudaMemcpyAsync(hc, d_counts + at, 4, cudaMemcpyDeviceToHost, e->stream));
        TB_CUDA(cudaStreamSynchronize(e->stream));
        *out = *h + *hc;
        return TAK_OK;
    };
    uint64_t total = 0;
    if (int r = offset_at(n, &total)) return r;
    if (total == 0) return TAK_OK;
    const bool last = depth_left == 2;   // the children of this frontier are the counted level
    const size_t PF_CAP = pf_cap();
    TB_CUDA(cudaFuncSetAttribute(k_perft_apply<N, true, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM));
    TB_CUDA(cudaFuncSetAttribute(k_perft_apply<N, false, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM));
    size_t begin = 0;
    uint64_t begin_off = 0;
    while (begin < n) {
        // largest `end` with offsets[end] - begin_off <= PF_CAP
        size_t end = n;
        uint64_t end_off = total;
        if (total - begin_off > PF_CAP) {
            size_t lo = begin + 1, hi = n;  // invariant: offsets[lo] fits (one parent never exceeds PF_CAP)
            while (lo < hi) {
                size_t mid = lo + (hi - lo + 1) / 2;
                uint64_t o = 0;
                if (int r = offset_at(mid, &o)) return r;
                if (o - begin_off <= PF_CAP) lo = mid; else hi = mid - 1;
            }
            end = lo;
            if (int r = offset_at(end, &end_off)) return r;
        }
        const size_t children = size_t(end_off - begin_off);
        if (children > 0) {
            TB_CHECK(children <= PF_CAP, TAK_ERR_CAPACITY, "one position has more children than the perft arena");
            const int n_blocks = int((children + P::CH - 1) / P::CH);
            TB_CUDA(lv.children.ensure(children * S));
            TB_CUDA(lv.moves.ensure(children * 2));
            TB_CUDA(lv.block_parent.ensure(size_t(n_blocks) * 4));
            if (!last) TB_CUDA(lv.counts.ensure(children * 4));
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            if (e->pf_n_spans < tak_engine::PF_SPANS) {
                ev0 = e->pf_span_ev[2 * e->pf_n_spans];
                ev1 = e->pf_span_ev[2 * e->pf_n_spans + 1];
                e->pf_span_children[e->pf_n_spans++] = children;
                TB_CUDA(cudaEventRecord(ev0, e->stream));
            }
            k_perft_moves<N><<<warp_blocks(int(end - begin)), GAME_THREADS, 0, e->stream>>>(
                frontier + begin * S, int(end - begin), d_counts + begin, d_off + begin, begin_off,
                lv.moves.as<uint16_t>(), lv.block_parent.as<int>(), P::CH);
            if (last)
                k_perft_apply<N, true, V><<<n_blocks, P::THREADS, P::SMEM, e->stream>>>(
                    frontier + begin * S, int(end - begin), d_off + begin, begin_off, lv.block_parent.as<int>(), n_blocks,
                    int(children), lv.moves.as<uint16_t>(), lv.children.as<uint8_t>(), nullptr, d_leaves);
            else
                k_perft_apply<N, false, V><<<n_blocks, P::THREADS, P::SMEM, e->stream>>>(
                    frontier + begin * S, int(end - begin), d_off + begin, begin_off, lv.block_parent.as<int>(), n_blocks,
                    int(children), lv.moves.as<uint16_t>(), lv.children.as<uint8_t>(), lv.counts.as<uint32_t>(), d_leaves);
            if (ev1) TB_CUDA(cudaEventRecord(ev1, e->stream));
            e->pf_launches += 2;
            e->pf_materialised += children;
            TB_CUDA(cudaGetLastError());
            if (!last)
                if (int r = perft_level<N, V>(e, lv.children.as<uint8_t>(), children, lv.counts.as<uint32_t>(),
                                           depth_left - 1, level + 1, d_leaves))
                    return r;
        }
        begin = end;
        begin_off = end_off;
    }
    return TAK_OK;
}

// roots: one count pass (the only frontier that is not produced by k_perft_apply), then the levels
template <int N>
static int perft_roots(tak_engine* e, const uint8_t* roots, size_t n, int depth, unsigned long long* d_leaves) {
    const bool last = depth == 1;
    uint32_t* d_counts = nullptr;
    if (!last) {
        TB_CUDA(e->pf_root_counts.ensure(n * 4));
        d_counts = e->pf_root_counts.as<uint32_t>();
    }
    k_perft_count<N><<<(unsigned(n) + 255) / 256, 256, 0, e->stream>>>(roots, int(n), last ? 1 : 0, d_counts, d_leaves);
    e->pf_launches++;
    TB_CUDA(cudaGetLastError());
    if (last) return TAK_OK;
    switch (perft_variant()) {
        case 0: return perft_level<N, 0>(e, roots, n, d_counts, depth, 0, d_leaves);
        case 1: return perft_level<N, 1>(e, roots, n, d_counts, depth, 0, d_leaves);
        case 5: return perft_level<N, 5>(e, roots, n, d_counts, depth, 0, d_leaves);
        default: return perft_level<N, 3>(e, roots, n, d_counts, depth, 0, d_leaves);
    }
}

extern "C" {

int32_t tak_engine_create(const tak_engine_config_t* cfg, tak_engine_t** out) {
    TB_CHECK(cfg && out, TAK_ERR_BAD_ARG, "tak_engine_create: null argument");
    TB_CHECK(cfg->n >= 3 && cfg->n <= 8, TAK_ERR_BAD_ARG, "board size %d not in 3..8", cfg->n);
    TB_CHECK(cfg->max_games > 0, TAK_ERR_BAD_ARG, "max_games must be positive");
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s): taknative has no CPU fallback", cudaGetErrorString(ce));
        return TAK_ERR_CUDA;
    }
    TB_CHECK(cfg->device >= 0 && cfg->device < count, TAK_ERR_BAD_ARG, "device %d out of range", cfg->device);
    TB_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    TB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    TB_CHECK(prop.major >= 10, TAK_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
             prop.major, prop.minor);
    tak_engine* e = new tak_engine();
    e->device = cfg->device;
    e->n = cfg->n;
    e->nsq = cfg->n * cfg->n;
    e->state_bytes = state_bytes_for(cfg->n);
    e->max_games = cfg->max_games;
    e->nodes_per_game = cfg->nodes_per_game;
    e->max_batch = cfg->max_batch > 0 ? cfg->max_batch : cfg->max_games;
    e->num_sms = prop.multiProcessorCount;
    cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = e->states.ensure(size_t(e->max_games) * e->state_bytes);
    if (err != cudaSuccess) {
        set_error("engine allocation failed: %s", cudaGetErrorString(err));
        delete e;
        return TAK_ERR_CUDA;
    }
    *out = e;
    return tak_games_reset(e, 0, e->max_games, 0);
}

int32_t tak_engine_destroy(tak_engine_t* e) {
    if (!e) return TAK_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    comm_destroy(e);
    examples_destroy(e);
    selfplay_destroy(e);
    mcts_destroy(e);
    net_destroy(e);
    for (DevBuf* b : {&e->states, &e->d_ids, &e->d_moves, &e->d_counts, &e->d_status, &e->d_results, &e->d_stage,
                      &e->pf_root, &e->pf_root_counts, &e->pf_scan_tmp, &e->pf_leaves, &e->po_plies, &e->po_result,
                      &e->po_totals})
        b->release();
    for (auto& lv : e->pf_level)
        for (DevBuf* b : {&lv.children, &lv.counts, &lv.offsets, &lv.moves, &lv.block_parent}) b->release();
    for (cudaEvent_t ev : e->pf_span_ev)
        if (ev) cudaEventDestroy(ev);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    cudaStreamDestroy(e->stream);
    delete e;
    return TAK_OK;
}

int32_t tak_engine_sync(tak_engine_t* e) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    TB_CUDA(cudaSetDevice(e->device));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t tak_games_reset(tak_engine_t* e, int32_t first, int32_t count, int32_t half_komi) {
    TB_CHECK(e && first >= 0 && count >= 0 && first + count <= e->max_games, TAK_ERR_BAD_ARG, "bad game range");
    TB_CHECK(half_komi >= -128 && half_komi <= 127, TAK_ERR_BAD_ARG, "half_komi out of i8 range");
    if (count == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    TB_DISPATCH_N(e->n, (k_reset<N_><<<warp_blocks(count), GAME_THREADS, 0, e->stream>>>(e->states.as<uint8_t>(),
                                                                                         first, count, half_komi)));
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

int32_t tak_games_upload(tak_engine_t* e, const int32_t* ids, int32_t n, const tak_state_t* states) {
    if (int r = check_ids(e, ids, n)) return r;
    TB_CHECK(states, TAK_ERR_BAD_ARG, "null states");
    if (n == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    const size_t S = e->state_bytes;
    TB_CUDA(e->ensure_pinned(S * n));
    uint8_t* rec = static_cast<uint8_t*>(e->h_stage);
    for (int i = 0; i < n; ++i) {
        TB_CHECK(states[i].n == e->n, TAK_ERR_BAD_ARG, "state %d has board size %d, engine has %d", i, states[i].n,
                 e->n);
        pack_state(e->n, states[i], rec + S * i);
    }
    // one H2D copy of the packed records + one scatter kernel
    TB_CUDA(e->d_stage.ensure(S * n));
    TB_CUDA(e->d_ids.ensure(size_t(n) * 4));
    TB_CUDA(cudaMemcpyAsync(e->d_stage.p, rec, S * n, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(e->d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, e->stream));
    k_scatter_records<<<(n * int(S / 16) + 255) / 256, 256, 0, e->stream>>>(
        e->d_stage.as<uint4>(), e->states.as<uint4>(), e->d_ids.as<int>(), n, int(S / 16), 1);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t tak_games_download(tak_engine_t* e, const int32_t* ids, int32_t n, tak_state_t* states) {
    if (int r = check_ids(e, ids, n)) return r;
    TB_CHECK(states, TAK_ERR_BAD_ARG, "null states");
    if (n == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    const size_t S = e->state_bytes;
    TB_CUDA(e->ensure_pinned(S * n));
    uint8_t* rec = static_cast<uint8_t*>(e->h_stage);
    TB_CUDA(e->d_stage.ensure(S * n));
    TB_CUDA(e->d_ids.ensure(size_t(n) * 4));
    TB_CUDA(cudaMemcpyAsync(e->d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, e->stream));
    k_scatter_records<<<(n * int(S / 16) + 255) / 256, 256, 0, e->stream>>>(
        e->d_stage.as<uint4>(), e->states.as<uint4>(), e->d_ids.as<int>(), n, int(S / 16), 0);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(rec, e->d_stage.p, S * n, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < n; ++i) unpack_state(e->n, rec + S * i, states[i]);
    return TAK_OK;
}

int32_t tak_possible_moves(tak_engine_t* e, const int32_t* ids, int32_t n, uint16_t* out_moves,
                           int32_t* out_offsets, int32_t cap) {
    if (int r = check_ids(e, ids, n)) return r;
    TB_CHECK(out_moves && out_offsets && cap >= 0, TAK_ERR_BAD_ARG, "null output");
    if (n == 0) { out_offsets[0] = 0; return TAK_OK; }
    TB_CUDA(cudaSetDevice(e->device));
    const int stride = 1024;  // per-game staging; games with more legal moves are re-run with a larger stride
    TB_CUDA(e->d_ids.ensure(size_t(n) * 4));
    TB_CUDA(e->d_counts.ensure(size_t(n) * 4));
    TB_CUDA(e->d_moves.ensure(size_t(n) * stride * 2));
    TB_CUDA(cudaMemcpyAsync(e->d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, e->stream));
    TB_DISPATCH_N(e->n, (k_moves<N_><<<warp_blocks(n), GAME_THREADS, 0, e->stream>>>(
                            e->states.as<uint8_t>(), e->d_ids.as<int>(), n, e->d_moves.as<uint16_t>(),
                            e->d_counts.as<int>(), stride)));
    e->launches++;
    TB_CUDA(cudaGetLastError());
    std::vector<int> counts(n);
    std::vector<uint16_t> staged(size_t(n) * stride);
    TB_CUDA(cudaMemcpyAsync(counts.data(), e->d_counts.p, size_t(n) * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaMemcpyAsync(staged.data(), e->d_moves.p, staged.size() * 2, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    int total = 0;
    for (int i = 0; i < n; ++i) {
        TB_CHECK(counts[i] <= stride, TAK_ERR_CAPACITY, "game %d has %d legal moves (> staging stride %d)", ids[i],
                 counts[i], stride);
        out_offsets[i] = total;
        total += counts[i];
    }
    out_offsets[n] = total;
    TB_CHECK(total <= cap, TAK_ERR_CAPACITY, "%d moves do not fit the caller's buffer of %d", total, cap);
    for (int i = 0; i < n; ++i)
        std::memcpy(out_moves + out_offsets[i], staged.data() + size_t(i) * stride, size_t(counts[i]) * 2);
    return TAK_OK;
}

int32_t tak_play(tak_engine_t* e, const int32_t* ids, const uint16_t* moves, int32_t n, int32_t* out_status) {
    if (int r = check_ids(e, ids, n)) return r;
    TB_CHECK(moves && out_status, TAK_ERR_BAD_ARG, "null moves/status");
    if (n == 0) return TAK_OK;
    // a game may appear only once per call (plays of one game are ordered by calls, like Game::play)
    TB_CUDA(cudaSetDevice(e->device));
    TB_CUDA(e->d_ids.ensure(size_t(n) * 4));
    TB_CUDA(e->d_moves.ensure(size_t(n) * 2));
    TB_CUDA(e->d_status.ensure(size_t(n) * 4));
    TB_CUDA(cudaMemcpyAsync(e->d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(e->d_moves.p, moves, size_t(n) * 2, cudaMemcpyHostToDevice, e->stream));
    TB_DISPATCH_N(e->n, (k_play<N_><<<warp_blocks(n), GAME_THREADS, 0, e->stream>>>(
                            e->states.as<uint8_t>(), e->d_ids.as<int>(), e->d_moves.as<uint16_t>(), n,
                            e->d_status.as<int>())));
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(out_status, e->d_status.p, size_t(n) * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t tak_result(tak_engine_t* e, const int32_t* ids, int32_t n, uint8_t* out_results) {
    if (int r = check_ids(e, ids, n)) return r;
    TB_CHECK(out_results, TAK_ERR_BAD_ARG, "null output");
    if (n == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    TB_CUDA(e->d_ids.ensure(size_t(n) * 4));
    TB_CUDA(e->d_results.ensure(size_t(n)));
    TB_CUDA(cudaMemcpyAsync(e->d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, e->stream));
    TB_DISPATCH_N(e->n, (k_result<N_><<<warp_blocks(n), GAME_THREADS, 0, e->stream>>>(
                            e->states.as<uint8_t>(), e->d_ids.as<int>(), n, e->d_results.as<uint8_t>())));
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(out_results, e->d_results.p, size_t(n), cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t tak_perft(tak_engine_t* e, const tak_state_t* root, int32_t depth, uint64_t* out_nodes) {
    return tak_perft_multi(e, root, 1, depth, out_nodes);
}

int32_t tak_perft_multi(tak_engine_t* e, const tak_state_t* roots, int32_t n_roots, int32_t depth, uint64_t* out_nodes) {
    TB_CHECK(e && roots && out_nodes && depth >= 0 && n_roots >= 0, TAK_ERR_BAD_ARG, "tak_perft: bad argument");
    TB_CHECK(depth <= 12, TAK_ERR_BAD_ARG, "depth %d too large", depth);
    TB_CUDA(cudaSetDevice(e->device));
    if (depth == 0 || n_roots == 0) { *out_nodes = uint64_t(n_roots); return TAK_OK; }
    std::vector<uint8_t> rec(size_t(e->state_bytes) * n_roots);
    for (int i = 0; i < n_roots; ++i) {
        TB_CHECK(roots[i].n == e->n, TAK_ERR_BAD_ARG, "root %d has board size %d, engine has %d", i, roots[i].n, e->n);
        pack_state(e->n, roots[i], rec.data() + size_t(i) * e->state_bytes);
    }
    TB_CUDA(e->pf_root.ensure(rec.size()));
    TB_CUDA(e->pf_leaves.ensure(8));
    TB_CUDA(cudaMemcpyAsync(e->pf_root.p, rec.data(), rec.size(), cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));   // `rec` is a stack-lifetime host buffer
    TB_CUDA(cudaMemsetAsync(e->pf_leaves.p, 0, 8, e->stream));
    e->pf_materialised = 0;
    e->pf_launches = 0;
    for (cudaEvent_t& ev : e->pf_span_ev)
        if (!ev) TB_CUDA(cudaEventCreate(&ev));
    cudaEvent_t e0, e1;
    TB_CUDA(cudaEventCreate(&e0));
    TB_CUDA(cudaEventCreate(&e1));
    TB_CUDA(cudaEventRecord(e0, e->stream));
    int r = TAK_OK;
    e->pf_n_spans = 0;
    TB_DISPATCH_N(e->n, r = perft_roots<N_>(e, e->pf_root.as<uint8_t>(), n_roots, depth,
                                            e->pf_leaves.as<unsigned long long>()));
    if (r == TAK_OK) {
        cudaEventRecord(e1, e->stream);
        unsigned long long total = 0;
        cudaError_t ce = cudaMemcpyAsync(&total, e->pf_leaves.p, 8, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) {
            set_error("tak_perft: %s", cudaGetErrorString(ce));
            r = TAK_ERR_CUDA;
        } else {
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            e->pf_ms = ms;
            *out_nodes = total;
            e->pf_expand_ms = 0;
            e->pf_top_span_ms = 0;
            e->pf_top_span_children = 0;
            for (int i = 0; i < e->pf_n_spans; ++i) {
                float sp = 0;
                cudaEventElapsedTime(&sp, e->pf_span_ev[2 * i], e->pf_span_ev[2 * i + 1]);
                e->pf_expand_ms += sp;
                if (e->pf_span_children[i] > e->pf_top_span_children) {
                    e->pf_top_span_children = e->pf_span_children[i];
                    e->pf_top_span_ms = sp;
                }
            }
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e->launches += e->pf_launches;
    return r;
}

int32_t tak_perft_stats(tak_engine_t* e, double* out_ms, uint64_t* out_materialised, uint64_t* out_launches) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    if (out_ms) *out_ms = e->pf_ms;
    if (out_materialised) *out_materialised = e->pf_materialised;
    if (out_launches) *out_launches = e->pf_launches;
    return TAK_OK;
}

int32_t tak_perft_profile(tak_engine_t* e, double* out6) {
    TB_CHECK(e && out6, TAK_ERR_BAD_ARG, "tak_perft_profile: bad argument");
    out6[0] = e->pf_ms;
    out6[1] = e->pf_expand_ms;
    out6[2] = double(e->pf_materialised);
    out6[3] = double(e->pf_launches);
    out6[4] = double(e->pf_top_span_children);
    out6[5] = e->pf_top_span_ms;
    return TAK_OK;
}

int32_t tak_playouts(tak_engine_t* e, int32_t first, int32_t count, uint64_t seed, int32_t game_id_base, int32_t max_plies,
                     int32_t ply_spread, int32_t* out_plies, uint8_t* out_results, uint64_t* out_totals2,
                     double* out_ms) {
    TB_CHECK(e && first >= 0 && count >= 0 && first + count <= e->max_games && max_plies >= 0 && ply_spread >= 0,
             TAK_ERR_BAD_ARG, "tak_playouts: bad argument");
    if (out_totals2) out_totals2[0] = out_totals2[1] = 0;
    if (out_ms) *out_ms = 0;
    if (count == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    TB_CUDA(e->po_plies.ensure(size_t(count) * 4));
    TB_CUDA(e->po_result.ensure(size_t(count)));
    TB_CUDA(e->po_totals.ensure(16));
    TB_CUDA(cudaMemsetAsync(e->po_totals.p, 0, 16, e->stream));
    cudaEvent_t e0, e1;
    TB_CUDA(cudaEventCreate(&e0));
    TB_CUDA(cudaEventCreate(&e1));
    TB_CUDA(cudaEventRecord(e0, e->stream));
    TB_DISPATCH_N(e->n, (k_playout<N_><<<warp_blocks(count), GAME_THREADS, 0, e->stream>>>(
                            e->states.as<uint8_t>(), first, count, seed, game_id_base, max_plies, ply_spread,
                            e->po_plies.as<int>(), e->po_result.as<uint8_t>(),
                            e->po_totals.as<unsigned long long>())));
    e->launches++;
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaEventRecord(e1, e->stream);
    unsigned long long totals[2] = {0, 0};
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(totals, e->po_totals.p, 16, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess && out_plies)
        ce = cudaMemcpyAsync(out_plies, e->po_plies.p, size_t(count) * 4, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess && out_results)
        ce = cudaMemcpyAsync(out_results, e->po_result.p, size_t(count), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    float ms = 0;
    if (ce == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ce != cudaSuccess) {
        set_error("tak_playouts: %s", cudaGetErrorString(ce));
        return TAK_ERR_CUDA;
    }
    if (out_totals2) { out_totals2[0] = totals[0]; out_totals2[1] = totals[1]; }
    if (out_ms) *out_ms = ms;
    return TAK_OK;
}

}  // extern "C"
