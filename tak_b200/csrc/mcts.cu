// alpha_tak::Node part of the C ABI: batched device MCTS (one tree per game).
// Reference: alpha-tak/src/search/{node,mcts,play,noise}.rs; schedule of train/src/self_play.rs:181-210.
#include <algorithm>
#include "mcts.hpp"

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "game_kernels.cuh"
#include "net.hpp"

namespace tb {

MctsView MctsState::view() const {
    MctsView v{};
    v.stat = stat.as<uint4>();
    v.link = link.as<uint2>();
    v.half = half.as<int>();
    v.top = top.as<uint32_t>();
    v.pend_cnt = pend_cnt.as<int>();
    v.pend_leaf = pend_leaf.as<uint32_t>();
    v.pend_plen = pend_plen.as<int>();
    v.pend_path = pend_path.as<uint32_t>();
    v.leaf_states = leaf_states.as<uint8_t>();
    v.explo = explo.as<float>();
    v.move_table = move_table.as<uint16_t>();
    v.err = err.as<int>();
    v.counters = counters.as<unsigned long long>();
    v.cap = cap;
    v.kcap = kcap;
    return v;
}

static inline int warp_blocks(int warps) { return (warps + GAME_WARPS_PER_BLOCK - 1) / GAME_WARPS_PER_BLOCK; }

int mcts_ensure(tak_engine* e, int k) {
    const int G = e->max_games;
    if (!e->mcts) {
        MctsState* m = new MctsState();
        e->mcts = m;
        m->cap = e->nodes_per_game > 0 ? e->nodes_per_game : (1 << 18);
        TB_CHECK(m->cap >= 64 && m->cap <= (1 << 24), TAK_ERR_BAD_ARG, "nodes_per_game %d out of range", m->cap);
        const size_t nodes = size_t(G) * 2 * m->cap;
        TB_CUDA(m->stat.ensure(nodes * 16));
        TB_CUDA(m->link.ensure(nodes * 8));
        TB_CUDA(m->half.ensure(size_t(G) * 4));
        TB_CUDA(m->top.ensure(size_t(G) * 4));
        TB_CUDA(m->pend_cnt.ensure(size_t(G) * 4));
        TB_CUDA(m->eval_count.ensure(16));
        TB_CUDA(m->err.ensure(16));
        TB_CUDA(m->counters.ensure(64));
        TB_CUDA(cudaMemsetAsync(m->half.p, 0, size_t(G) * 4, e->stream));
        TB_CUDA(cudaMemsetAsync(m->pend_cnt.p, 0, size_t(G) * 4, e->stream));
        TB_CUDA(cudaMemsetAsync(m->err.p, 0, 16, e->stream));
        TB_CUDA(cudaMemsetAsync(m->counters.p, 0, 64, e->stream));
        TB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&m->h_pinned), 64));
        // exploration_rate(n) = ln((1 + n + 500) / 500) + 4 in f32 (mcts.rs:7-12), tabulated with the host libm
        std::vector<float> tab(MCTS_EXPLO_TABLE);
        for (int i = 0; i < MCTS_EXPLO_TABLE; ++i) {
            volatile float nf = float(i);
            volatile float a = 1.0f + nf;
            volatile float b = a + 500.0f;
            volatile float c = b / 500.0f;
            volatile float d = logf(c);
            tab[i] = d + 4.0f;
        }
        TB_CUDA(m->explo.ensure(tab.size() * 4));
        TB_CUDA(cudaMemcpyAsync(m->explo.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, e->stream));
        const std::vector<uint16_t>& mt = host_move_index_table(e->n);
        TB_CUDA(m->move_table.ensure(mt.size() * 2));
        TB_CUDA(cudaMemcpyAsync(m->move_table.p, mt.data(), mt.size() * 2, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaStreamSynchronize(e->stream));
        m->kcap = 0;
        if (int r = mcts_ensure(e, k)) return r;
        return mcts_launch_tree_reset(e, nullptr, G);
    }
    MctsState& m = *e->mcts;
    if (k > m.kcap) {
        TB_CHECK(!m.queued, TAK_ERR_BAD_ARG, "cannot grow the pending queue (k=%d) while leaves are queued", k);
        const size_t slots = size_t(G) * k;
        TB_CUDA(m.pend_leaf.ensure(slots * 4));
        TB_CUDA(m.pend_plen.ensure(slots * 4));
        TB_CUDA(m.pend_path.ensure(slots * MCTS_MAX_DEPTH * 4));
        TB_CUDA(m.leaf_states.ensure(slots * e->state_bytes));
        TB_CUDA(m.eval_index.ensure(slots * 4));
        TB_CUDA(m.eval_slot.ensure(slots * 4));
        m.kcap = k;
    }
    return TAK_OK;
}

int mcts_launch_tree_reset(tak_engine* e, const int* d_ids, int n) {
    k_mcts_tree_reset<<<(n + 255) / 256, 256, 0, e->stream>>>(e->mcts->view(), d_ids, n);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

int mcts_launch_rollout(tak_engine* e, const int* d_ids, int n, int k, const uint8_t* d_enable) {
    MctsState& m = *e->mcts;
    // 4 warps per block (80 registers/thread): small enough to share an SM with a resident conv-tower CTA of another
    // engine replica working on the same GPU (bench.py runs two replicas per GPU so MCTS hides under the tower)
    TB_DISPATCH_N(e->n, (k_mcts_rollout<N_><<<(n + 3) / 4, 128, 0, e->stream>>>(
                            m.view(), e->states.as<uint8_t>(), d_ids, n, k, d_enable)));
    e->launches++;
    m.queued = true;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

int mcts_launch_compact(tak_engine* e, bool clamp_to_batch) {
    MctsState& m = *e->mcts;
    k_mcts_compact<<<1, 256, 0, e->stream>>>(m.pend_cnt.as<int>(), e->max_games, m.kcap, m.limits_on ? m.limits.as<int>() : nullptr,
                                             m.eval_index.as<int>(), m.eval_slot.as<int>(), m.eval_count.as<int>(),
                                             clamp_to_batch ? int(std::min<long long>(e->max_batch, (long long)e->max_games * m.kcap))
                                                            : INT_MAX,
                                             m.err.as<int>());
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

int mcts_read_eval_count(tak_engine* e, int* out) {
    MctsState& m = *e->mcts;
    TB_CUDA(cudaMemcpyAsync(m.h_pinned, m.eval_count.p, 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    *out = m.h_pinned[0];
    return TAK_OK;
}

int mcts_launch_backup(tak_engine* e, const PriorSource& ps) {
    MctsState& m = *e->mcts;
    TB_DISPATCH_N(e->n, (k_mcts_backup<N_><<<warp_blocks(e->max_games), GAME_THREADS, 0, e->stream>>>(
                            m.view(), m.eval_slot.as<int>(), e->max_games, ps,
                            m.limits_on ? m.limits.as<int>() : nullptr, m.leaf_states.as<uint8_t>())));
    e->launches++;
    m.queued = m.limits_on;  // a partial backup may leave newer leaves queued
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

int mcts_eval_and_backup(tak_engine* e) {
    TB_CHECK(e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    MctsState& m = *e->mcts;
    NetState& ns = *e->net;
    if (int r = mcts_launch_compact(e, ns.arch != 0)) return r;
    PriorSource ps{};
    ps.arch = ns.arch;
    ps.psz = ns.policy_out;
    if (ns.arch != 0) {
        // The number of queued leaves stays on the device: the forward pass is launched for the largest batch the queues
        // can hold (capped by max_batch; the compaction kernel flags an overflow, reported as TAK_ERR_CAPACITY by the
        // next mcts_check_errors) and every kernel of it reads the live count -- no stream synchronisation between
        // queueing leaves and backing them up (Player::rollout, mcts_devirtualize).
        const int bound = int(std::min<long long>(e->max_batch, (long long)e->max_games * m.kcap));
        if (int r = net_forward(e, m.leaf_states.as<uint8_t>(), m.eval_index.as<int>(), bound, nullptr, 0,
                                m.eval_count.as<int>()))
            return r;
        ps.logits = ns.logits.as<float>();
        ps.stats = ns.stats.as<float2>();
        ps.values = ns.values.as<float>();
        ps.S = ns.cap_S;
    }
    return mcts_launch_backup(e, ps);
}

// ---- the fused search loop -------------------------------------------------------------------------------------
// `reps` x Node::rollout (mcts.rs:16-23) for the listed games in lock step, one leaf per game per iteration, with NO host
// round trip: leaves take their evaluation slot with a device atomic and write their input planes from the rollout warp,
// the tower reads the leaf count from device memory, the heads are computed by the backup warp, and backup(i) + rollout
// (i+1) share a launch:   step(rollout) ; { tower ; step(backup + rollout) } x (reps - 1) ; tower ; step(backup).
// Two launches per rollout (Net6), against seven and a stream synchronisation in the stepwise path.
static int launch_step(tak_engine* e, const int* d_ids, int n, const uint8_t* d_enable, const FastEval& fe,
                       const PriorSource& ps, int do_backup, int do_rollout, int reps) {
    MctsState& m = *e->mcts;
    // The 64-register / 64-thread build was meant to run BESIDE the other engine replica's conv tower.  It does not: a
    // tower CTA (11 warps x 168 registers, allocated in units of 4 warps = 64 512 registers) leaves 1 024 registers on its
    // SM, and no block shape / register cap / stream priority made the step co-resident (profiles/r02_step_overlap.md);
    // the capped build is kept because it costs nothing measurable (3 858-3 920 moves/s for every variant).
    // TAK_STEP_REGS=128 selects the uncapped build; the DummyNet loop always uses it.
    static const int regs = [] {
        const char* s = std::getenv("TAK_STEP_REGS");
        return s ? std::atoi(s) : 64;
    }();
    const bool capped = regs != 128 && ps.arch != 0 && (e->n == 5 || e->n == 6);
    static const int cwarps = [] {
        const char* s = std::getenv("TAK_STEP_WARPS");
        return s ? std::atoi(s) : 2;
    }();
    if (capped && e->n == 6 && cwarps == 3) {      // 3 warps x 64 registers (A/B variant)
        k_mcts_step<6, 96, 10><<<(n + 2) / 3, 96, 0, e->stream>>>(m.view(), e->states.as<uint8_t>(), d_ids, n, d_enable, fe,
                                                                  ps, do_backup, do_rollout, reps);
    } else if (capped && e->n == 6) {
        k_mcts_step<6, 64, 16><<<(n + 1) / 2, 64, 0, e->stream>>>(m.view(), e->states.as<uint8_t>(), d_ids, n, d_enable, fe,
                                                                  ps, do_backup, do_rollout, reps);
    } else if (capped && e->n == 5) {
        k_mcts_step<5, 64, 16><<<(n + 1) / 2, 64, 0, e->stream>>>(m.view(), e->states.as<uint8_t>(), d_ids, n, d_enable, fe,
                                                                  ps, do_backup, do_rollout, reps);
    } else {
        TB_DISPATCH_N(e->n, (k_mcts_step<N_><<<(n + 3) / 4, 128, 0, e->stream>>>(
                                m.view(), e->states.as<uint8_t>(), d_ids, n, d_enable, fe, ps, do_backup, do_rollout, reps)));
    }
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

int mcts_fast_rollouts(tak_engine* e, const int* d_ids, int n, int reps, const uint8_t* d_enable) {
    TB_CHECK(e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    MctsState& m = *e->mcts;
    NetState& ns = *e->net;
    if (reps <= 0 || n <= 0) return TAK_OK;
    if (m.queued) {   // leaves queued by an earlier stepwise call: evaluate and back them up first (queue order)
        if (int r = mcts_eval_and_backup(e)) return r;
    }
    FastEval fe{};
    PriorSource ps{};
    ps.arch = ns.arch;
    ps.psz = ns.policy_out;
    if (ns.arch == 0) {
        // DummyNet: nothing to evaluate -- every rollout's backup can follow it at once, the whole loop is one launch
        if (int r = launch_step(e, d_ids, n, d_enable, fe, ps, 1, 1, reps)) return r;
        return launch_step(e, d_ids, n, nullptr, fe, ps, 1, 0, 1);
    }
    TB_CHECK(n <= e->max_batch, TAK_ERR_CAPACITY, "%d games searched together exceed max_batch %d", n, e->max_batch);
    if (int r = net_fast_views(e, n, fe, ps)) return r;
    int* cnt = m.eval_count.as<int>() + 2;   // [2], [3]: the two phases of the loop ([0] belongs to the compaction path)
    TB_CUDA(cudaMemsetAsync(cnt, 0, 8, e->stream));
    fe.eval_slot = m.eval_slot.as<int>();
    // TAK_STEP_TIMING=1 (profiling aid, tools/probe_overlap.py): CUDA events around the step kernel in the middle of the loop
    static const bool timing = [] { const char* v = std::getenv("TAK_STEP_TIMING"); return v && std::atoi(v) != 0; }();
    cudaEvent_t tev[2] = {nullptr, nullptr};
    const bool timed_one = timing && reps >= 4;
    if (timed_one) {
        TB_CUDA(cudaEventCreate(&tev[0]));
        TB_CUDA(cudaEventCreate(&tev[1]));
    }
    int phase = 0;
    fe.eval_count = cnt + phase;
    fe.eval_count_reset = cnt + (phase ^ 1);
    if (int r = launch_step(e, d_ids, n, d_enable, fe, ps, 0, 1, 1)) return r;
    for (int i = 1; i < reps; ++i) {
        if (int r = net_tower_fast(e, n, cnt + phase)) return r;
        phase ^= 1;
        fe.eval_count = cnt + phase;
        fe.eval_count_reset = cnt + (phase ^ 1);
        if (timed_one && i == reps / 2) TB_CUDA(cudaEventRecord(tev[0], e->stream));
        if (int r = launch_step(e, d_ids, n, d_enable, fe, ps, 1, 1, 1)) return r;
        if (timed_one && i == reps / 2) TB_CUDA(cudaEventRecord(tev[1], e->stream));
    }
    if (int r = net_tower_fast(e, n, cnt + phase)) return r;
    fe.eval_count = nullptr;
    fe.eval_count_reset = nullptr;
    if (int r = launch_step(e, d_ids, n, nullptr, fe, ps, 1, 0, 1)) return r;
    if (timed_one) {   // event to event: includes the time the kernel waits for SMs held by another replica's tower
        TB_CUDA(cudaEventSynchronize(tev[1]));
        float ms = 0;
        TB_CUDA(cudaEventElapsedTime(&ms, tev[0], tev[1]));
        fprintf(stderr, "[step timing] engine %p: k_mcts_step of iteration %d of %d took %.1f us (n = %d)\n", (void*)e,
                reps / 2, reps, ms * 1e3, n);
        cudaEventDestroy(tev[0]);
        cudaEventDestroy(tev[1]);
    }
    m.queued = false;
    return TAK_OK;
}

int mcts_check_errors(tak_engine* e) {
    MctsState& m = *e->mcts;
    TB_CUDA(cudaMemcpyAsync(m.h_pinned + 1, m.err.p, 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    const int flags = m.h_pinned[1];
    if (!flags) return TAK_OK;
    TB_CUDA(cudaMemsetAsync(m.err.p, 0, 4, e->stream));
    if (flags & MERR_POOL_FULL) {
        set_error("MCTS node pool exhausted (nodes_per_game = %d): raise tak_engine_config.nodes_per_game", m.cap);
        return TAK_ERR_CAPACITY;
    }
    if (flags & MERR_PENDING_FULL) {
        set_error("more leaves queued than one game's queue (%d) or one network batch (max_batch %d) holds", m.kcap,
                  e->max_batch);
        return TAK_ERR_CAPACITY;
    }
    if (flags & MERR_DEPTH) {
        set_error("search path deeper than %d plies", MCTS_MAX_DEPTH);
        return TAK_ERR_CAPACITY;
    }
    if (flags & MERR_VISITS) {
        set_error("visit count beyond the exploration table (%d)", MCTS_EXPLO_TABLE);
        return TAK_ERR_CAPACITY;
    }
    if (flags & MERR_QUEUED) {
        set_error("mcts_play while leaves of that game are queued: call mcts_devirtualize first");
        return TAK_ERR_BAD_ARG;
    }
    if (flags & MERR_NAN) {
        set_error("tried comparing nan (NaN upper confidence bound) or select on a node without children");
        return TAK_ERR_BAD_ARG;
    }
    set_error("tried to play an invalid move / move without a policy index / noise before a visit");
    return TAK_ERR_INVALID_MOVE;
}

int mcts_launch_pick(tak_engine* e, const int* d_ids, int n, const uint8_t* d_sample, uint64_t seed,
                     const int* d_tags, uint16_t* d_out) {
    k_mcts_pick<<<warp_blocks(n), GAME_THREADS, 0, e->stream>>>(e->mcts->view(), d_ids, n, d_sample, seed, d_tags,
                                                               d_out);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}
int mcts_launch_reroot(tak_engine* e, const int* d_ids, const uint16_t* d_moves, int n) {
    k_mcts_reroot<<<warp_blocks(n), GAME_THREADS, 0, e->stream>>>(e->mcts->view(), d_ids, d_moves, n);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}
int mcts_launch_dirichlet(tak_engine* e, const int* d_ids, int n, const uint8_t* d_enable, float alpha, float ratio,
                          uint64_t seed, const int* d_tags) {
    k_mcts_dirichlet<<<warp_blocks(n), GAME_THREADS, 0, e->stream>>>(e->mcts->view(), d_ids, n, d_enable, alpha,
                                                                    ratio, seed, d_tags);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

void mcts_destroy(tak_engine* e) {
    if (!e->mcts) return;
    MctsState& m = *e->mcts;
    for (DevBuf* b : {&m.stat, &m.link, &m.half, &m.top, &m.pend_cnt, &m.pend_leaf, &m.pend_plen, &m.pend_path,
                      &m.leaf_states, &m.eval_index, &m.eval_slot, &m.eval_count, &m.explo, &m.move_table, &m.err, &m.limits,
                      &m.counters, &m.d_ids, &m.d_moves, &m.stage_policy, &m.stage_value, &m.stage_stat,
                      &m.stage_link, &m.stage_count})
        b->release();
    if (m.h_pinned) cudaFreeHost(m.h_pinned);
    delete e->mcts;
    e->mcts = nullptr;
}

static int upload_ids(tak_engine* e, const int32_t* ids, int n, const int** d_ids) {
    MctsState& m = *e->mcts;
    for (int i = 0; i < n; ++i)
        TB_CHECK(ids[i] >= 0 && ids[i] < e->max_games, TAK_ERR_BAD_ARG, "game id %d out of range", ids[i]);
    TB_CUDA(m.d_ids.ensure(size_t(n) * 4 + 4));
    TB_CUDA(cudaMemcpyAsync(m.d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, e->stream));
    *d_ids = m.d_ids.as<int>();
    return TAK_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int32_t mcts_tree_reset(tak_engine_t* e, const int32_t* ids, int32_t n) {
    TB_CHECK(e && ids && n >= 0, TAK_ERR_BAD_ARG, "mcts_tree_reset: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    return mcts_launch_tree_reset(e, d_ids, n);
}

int32_t mcts_virtual_rollout(tak_engine_t* e, const int32_t* ids, int32_t n, int32_t k) {
    TB_CHECK(e && ids && n >= 0 && k >= 1 && k <= 4096, TAK_ERR_BAD_ARG, "mcts_virtual_rollout: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, k)) return r;
    if (n == 0) return TAK_OK;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    if (int r = mcts_launch_rollout(e, d_ids, n, k)) return r;
    return mcts_check_errors(e);
}

int32_t mcts_pending(tak_engine_t* e, int32_t* out_count, int32_t* out_game_ids, tak_state_t* out_states,
                     int32_t cap) {
    TB_CHECK(e && out_count, TAK_ERR_BAD_ARG, "mcts_pending: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    MctsState& m = *e->mcts;
    if (int r = mcts_launch_compact(e)) return r;
    int count = 0;
    if (int r = mcts_read_eval_count(e, &count)) return r;
    *out_count = count;
    if (!out_game_ids && !out_states) return TAK_OK;
    TB_CHECK(count <= cap, TAK_ERR_CAPACITY, "%d queued leaves, caller buffer holds %d", count, cap);
    std::vector<int> index(count);
    TB_CUDA(cudaMemcpyAsync(index.data(), m.eval_index.p, size_t(count) * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    std::vector<uint8_t> rec(e->state_bytes);
    for (int i = 0; i < count; ++i) {
        if (out_game_ids) out_game_ids[i] = index[i] / m.kcap;
        if (out_states) {
            TB_CUDA(cudaMemcpyAsync(rec.data(), m.leaf_states.as<uint8_t>() + size_t(index[i]) * e->state_bytes,
                                    e->state_bytes, cudaMemcpyDeviceToHost, e->stream));
            TB_CUDA(cudaStreamSynchronize(e->stream));
            unpack_state(e->n, rec.data(), out_states[i]);
        }
    }
    return TAK_OK;
}

int32_t mcts_devirtualize(tak_engine_t* e) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (int r = mcts_eval_and_backup(e)) return r;
    return mcts_check_errors(e);
}

int32_t mcts_reserve_pending(tak_engine_t* e, int32_t k) {
    TB_CHECK(e && k >= 1 && k <= 8192, TAK_ERR_BAD_ARG, "mcts_reserve_pending: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    return mcts_ensure(e, k);
}

int32_t mcts_devirtualize_first(tak_engine_t* e, const int32_t* ids, int32_t n, const int32_t* counts) {
    TB_CHECK(e && ids && counts && n >= 0, TAK_ERR_BAD_ARG, "mcts_devirtualize_first: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    MctsState& m = *e->mcts;
    std::vector<int> lim(e->max_games, 0);
    for (int i = 0; i < n; ++i) {
        TB_CHECK(ids[i] >= 0 && ids[i] < e->max_games && counts[i] >= 0, TAK_ERR_BAD_ARG,
                 "mcts_devirtualize_first: bad id / count");
        lim[ids[i]] = counts[i];
    }
    TB_CUDA(m.limits.ensure(lim.size() * 4));
    TB_CUDA(cudaMemcpyAsync(m.limits.p, lim.data(), lim.size() * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));  // `lim` is a stack-lifetime host buffer
    m.limits_on = true;
    int r = mcts_eval_and_backup(e);
    m.limits_on = false;
    if (r) return r;
    return mcts_check_errors(e);
}

int32_t mcts_player_rollouts(tak_engine_t* e, const int32_t* ids, int32_t n, int32_t batch, int32_t reps) {
    TB_CHECK(e && ids && n >= 0 && batch >= 1 && batch <= 4096 && reps >= 0, TAK_ERR_BAD_ARG,
             "mcts_player_rollouts: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 2 * batch)) return r;
    if (n == 0 || reps == 0) return TAK_OK;
    MctsState& m = *e->mcts;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    TB_CUDA(m.limits.ensure(size_t(e->max_games) * 4));
    int r = TAK_OK;
    for (int rep = 0; rep < reps && r == TAK_OK; ++rep) {
        // request_batch: the new batch is selected on trees that still carry the virtual visits of the outstanding one
        TB_CUDA(cudaMemsetAsync(m.limits.p, 0, size_t(e->max_games) * 4, e->stream));
        k_mcts_snapshot_limits<<<(n + 127) / 128, 128, 0, e->stream>>>(m.pend_cnt.as<int>(), d_ids, n, m.limits.as<int>());
        e->launches++;
        r = mcts_launch_rollout(e, d_ids, n, batch);
        if (r) break;
        // consume_batch: evaluate and back up the OLDER batch only
        m.limits_on = true;
        r = mcts_eval_and_backup(e);
        m.limits_on = false;
    }
    if (r) return r;
    return mcts_check_errors(e);
}

int32_t mcts_devirtualize_with(tak_engine_t* e, const float* policy, const float* value, int32_t count) {
    TB_CHECK(e && policy && value && count >= 0, TAK_ERR_BAD_ARG, "mcts_devirtualize_with: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    MctsState& m = *e->mcts;
    if (int r = mcts_launch_compact(e)) return r;
    int queued = 0;
    if (int r = mcts_read_eval_count(e, &queued)) return r;
    TB_CHECK(queued == count, TAK_ERR_BAD_ARG, "%d leaves are queued but %d network outputs were supplied", queued,
             count);
    const int psz = host_policy_size(e->n);
    TB_CUDA(m.stage_policy.ensure(size_t(count) * psz * 4 + 4));
    TB_CUDA(m.stage_value.ensure(size_t(count) * 4 + 4));
    TB_CUDA(cudaMemcpyAsync(m.stage_policy.p, policy, size_t(count) * psz * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(m.stage_value.p, value, size_t(count) * 4, cudaMemcpyHostToDevice, e->stream));
    PriorSource ps{};
    ps.arch = -1;
    ps.logits = m.stage_policy.as<float>();
    ps.values = m.stage_value.as<float>();
    ps.psz = psz;
    if (int r = mcts_launch_backup(e, ps)) return r;
    return mcts_check_errors(e);
}

int32_t mcts_rollouts(tak_engine_t* e, const int32_t* ids, int32_t n, int32_t n_rollouts) {
    TB_CHECK(e && ids && n >= 0 && n_rollouts >= 0, TAK_ERR_BAD_ARG, "mcts_rollouts: bad argument");
    TB_CHECK(e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    if (e->net->arch == 0 || n <= e->max_batch) {
        if (int r = mcts_fast_rollouts(e, d_ids, n, n_rollouts, nullptr)) return r;
    } else {
        for (int i = 0; i < n_rollouts; ++i) {   // more games than one network batch holds: stepwise
            if (int r = mcts_launch_rollout(e, d_ids, n, 1)) return r;
            if (int r = mcts_eval_and_backup(e)) return r;
        }
    }
    return mcts_check_errors(e);
}

static int export_root(tak_engine_t* e, int32_t id, std::vector<uint4>& st, std::vector<uint2>& lk, int* count) {
    TB_CHECK(id >= 0 && id < e->max_games, TAK_ERR_BAD_ARG, "game id %d out of range", id);
    if (int r = mcts_ensure(e, 1)) return r;
    MctsState& m = *e->mcts;
    const int cap = 4097;
    TB_CUDA(m.stage_stat.ensure(size_t(cap) * 16));
    TB_CUDA(m.stage_link.ensure(size_t(cap) * 8));
    TB_CUDA(m.stage_count.ensure(16));
    k_mcts_export_root<<<1, 256, 0, e->stream>>>(m.view(), id, m.stage_stat.as<uint4>(), m.stage_link.as<uint2>(), cap,
                                                 m.stage_count.as<int>());
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(m.h_pinned + 2, m.stage_count.p, 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    *count = m.h_pinned[2];
    TB_CHECK(*count + 1 <= cap, TAK_ERR_CAPACITY, "root has %d children", *count);
    st.resize(size_t(*count) + 1);
    lk.resize(size_t(*count) + 1);
    TB_CUDA(cudaMemcpyAsync(st.data(), m.stage_stat.p, st.size() * 16, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaMemcpyAsync(lk.data(), m.stage_link.p, lk.size() * 8, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t mcts_children(tak_engine_t* e, int32_t id, uint16_t* out_moves, uint32_t* out_visits, float* out_priors,
                      float* out_rewards, int32_t cap, int32_t* out_count) {
    TB_CHECK(e && out_count, TAK_ERR_BAD_ARG, "mcts_children: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    std::vector<uint4> st;
    std::vector<uint2> lk;
    int count = 0;
    if (int r = export_root(e, id, st, lk, &count)) return r;
    *out_count = count;
    TB_CHECK(count <= cap, TAK_ERR_CAPACITY, "root has %d children, caller buffer holds %d", count, cap);
    for (int i = 0; i < count; ++i) {
        if (out_moves) out_moves[i] = uint16_t(lk[size_t(i) + 1].y & 0xFFFFu);
        if (out_visits) out_visits[i] = st[size_t(i) + 1].z;
        if (out_priors) std::memcpy(&out_priors[i], &st[size_t(i) + 1].x, 4);
        if (out_rewards) std::memcpy(&out_rewards[i], &st[size_t(i) + 1].y, 4);
    }
    return TAK_OK;
}

int32_t mcts_children_batch(tak_engine_t* e, const int32_t* ids, int32_t n, uint16_t* out_moves, uint32_t* out_visits,
                            int32_t* out_counts, int32_t stride) {
    TB_CHECK(e && ids && out_moves && out_visits && out_counts && n >= 0 && stride > 0, TAK_ERR_BAD_ARG,
             "mcts_children_batch: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    MctsState& m = *e->mcts;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    TB_CUDA(m.stage_stat.ensure(size_t(n) * stride * 4 + 16));
    TB_CUDA(m.stage_link.ensure(size_t(n) * stride * 2 + 16));
    TB_CUDA(m.stage_count.ensure(size_t(n) * 4 + 16));
    k_mcts_export_children<<<(n + GAME_WARPS_PER_BLOCK - 1) / GAME_WARPS_PER_BLOCK, GAME_THREADS, 0, e->stream>>>(
        m.view(), d_ids, n, m.stage_link.as<uint16_t>(), m.stage_stat.as<uint32_t>(), m.stage_count.as<int>(), stride);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(out_moves, m.stage_link.p, size_t(n) * stride * 2, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaMemcpyAsync(out_visits, m.stage_stat.p, size_t(n) * stride * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaMemcpyAsync(out_counts, m.stage_count.p, size_t(n) * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < n; ++i)
        TB_CHECK(out_counts[i] <= stride, TAK_ERR_CAPACITY, "game %d has %d root children (> stride %d)", ids[i],
                 out_counts[i], stride);
    return TAK_OK;
}

int32_t mcts_debug(tak_engine_t* e, int32_t id, int32_t depth, tak_move_info_t* out, int32_t cap, int32_t* out_count) {
    static_assert(sizeof(tak_move_info_t) == sizeof(MctsMoveInfo) && TAK_DEBUG_MAX_DEPTH == MCTS_DEBUG_DEPTH,
                  "tak_move_info_t mirrors MctsMoveInfo");
    TB_CHECK(e && out && out_count && cap >= 0 && depth >= 0, TAK_ERR_BAD_ARG, "mcts_debug: bad argument");
    TB_CHECK(id >= 0 && id < e->max_games, TAK_ERR_BAD_ARG, "game id %d out of range", id);
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    MctsState& m = *e->mcts;
    const int kmax = 4096;   // >= the largest possible root fan-out the arena encodes per export (see export_root)
    TB_CUDA(m.stage_stat.ensure(size_t(kmax) * sizeof(MctsMoveInfo)));
    TB_CUDA(m.stage_count.ensure(16));
    k_mcts_debug<<<(kmax + GAME_WARPS_PER_BLOCK - 1) / GAME_WARPS_PER_BLOCK, GAME_THREADS, 0, e->stream>>>(
        m.view(), id, depth, m.stage_stat.as<MctsMoveInfo>(), kmax, m.stage_count.as<int>());
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(m.h_pinned + 2, m.stage_count.p, 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    const int count = m.h_pinned[2];
    *out_count = count;
    TB_CHECK(count <= kmax, TAK_ERR_CAPACITY, "root has %d children", count);
    TB_CHECK(count <= cap, TAK_ERR_CAPACITY, "root has %d children, caller buffer holds %d", count, cap);
    if (count == 0) return TAK_OK;
    std::vector<tak_move_info_t> tmp(static_cast<size_t>(count));
    TB_CUDA(cudaMemcpyAsync(tmp.data(), m.stage_stat.p, tmp.size() * sizeof(tak_move_info_t), cudaMemcpyDeviceToHost,
                            e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    // debug.rs:22-23: sort by visits, then reverse
    std::stable_sort(tmp.begin(), tmp.end(),
                     [](const tak_move_info_t& a, const tak_move_info_t& b) { return a.visits < b.visits; });
    std::reverse(tmp.begin(), tmp.end());
    std::memcpy(out, tmp.data(), tmp.size() * sizeof(tak_move_info_t));
    return TAK_OK;
}

int32_t mcts_root(tak_engine_t* e, int32_t id, uint32_t* out_visits, uint32_t* out_virtual, float* out_reward) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    TB_CUDA(cudaSetDevice(e->device));
    std::vector<uint4> st;
    std::vector<uint2> lk;
    int count = 0;
    if (int r = export_root(e, id, st, lk, &count)) return r;
    if (out_visits) *out_visits = st[0].z;
    if (out_virtual) *out_virtual = st[0].w;
    if (out_reward) std::memcpy(out_reward, &st[0].y, 4);
    return TAK_OK;
}

int32_t mcts_pick_move(tak_engine_t* e, const int32_t* ids, int32_t n, uint16_t* out_moves) {
    TB_CHECK(e && ids && out_moves && n >= 0, TAK_ERR_BAD_ARG, "mcts_pick_move: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    MctsState& m = *e->mcts;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    TB_CUDA(m.d_moves.ensure(size_t(n) * 2 + 2));
    if (int r = mcts_launch_pick(e, d_ids, n, nullptr, 0, nullptr, m.d_moves.as<uint16_t>())) return r;
    TB_CUDA(cudaMemcpyAsync(out_moves, m.d_moves.p, size_t(n) * 2, cudaMemcpyDeviceToHost, e->stream));
    return mcts_check_errors(e);
}

int32_t mcts_pick_move_sampled(tak_engine_t* e, const int32_t* ids, int32_t n, uint64_t seed, uint16_t* out_moves) {
    TB_CHECK(e && ids && out_moves && n >= 0, TAK_ERR_BAD_ARG, "mcts_pick_move_sampled: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    MctsState& m = *e->mcts;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    TB_CUDA(m.d_moves.ensure(size_t(n) * 2 + 2));
    TB_CUDA(m.stage_count.ensure(size_t(n) + 16));
    TB_CUDA(cudaMemsetAsync(m.stage_count.p, 1, size_t(n), e->stream));   // sample flag of every listed game
    if (int r = mcts_launch_pick(e, d_ids, n, m.stage_count.as<uint8_t>(), seed, nullptr, m.d_moves.as<uint16_t>()))
        return r;
    TB_CUDA(cudaMemcpyAsync(out_moves, m.d_moves.p, size_t(n) * 2, cudaMemcpyDeviceToHost, e->stream));
    return mcts_check_errors(e);
}

int32_t mcts_play(tak_engine_t* e, const int32_t* ids, const uint16_t* moves, int32_t n) {
    TB_CHECK(e && ids && moves && n >= 0, TAK_ERR_BAD_ARG, "mcts_play: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    MctsState& m = *e->mcts;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    TB_CUDA(m.d_moves.ensure(size_t(n) * 2 + 2));
    TB_CUDA(cudaMemcpyAsync(m.d_moves.p, moves, size_t(n) * 2, cudaMemcpyHostToDevice, e->stream));
    if (int r = mcts_launch_reroot(e, d_ids, m.d_moves.as<uint16_t>(), n)) return r;
    return mcts_check_errors(e);
}

int32_t mcts_apply_dirichlet(tak_engine_t* e, const int32_t* ids, int32_t n, float alpha, float ratio, uint64_t seed) {
    TB_CHECK(e && ids && n >= 0 && alpha > 0.f, TAK_ERR_BAD_ARG, "mcts_apply_dirichlet: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    if (int r = mcts_ensure(e, 1)) return r;
    if (n == 0) return TAK_OK;
    const int* d_ids = nullptr;
    if (int r = upload_ids(e, ids, n, &d_ids)) return r;
    if (int r = mcts_launch_dirichlet(e, d_ids, n, nullptr, alpha, ratio, seed, nullptr)) return r;
    return mcts_check_errors(e);
}

}  // extern "C"
