// alpha_tak::Node on the device: one search tree per game in a per-game node arena, one warp per game.
// Reference: alpha-tak/src/search/node.rs:4-39, mcts.rs:7-125, play.rs:13-67, noise.rs:6-16.
//
// Node pool (per game, two halves of `cap` nodes for re-root compaction):
//   stat[node] = {prior f32, expected_reward f32, visits u32, virtual_visits u32}      16 B (one float4 per child,
//                so a PUCT scan of a node's children is one coalesced read)
//   link[node] = {child_base | result << 24, move u16 | n_children u16 << 16}           8 B
// Children of a node are contiguous and in move-generation order; the root is node 0 of the current half.
//
// f32 semantics: every arithmetic step below mirrors the Rust expression tree; the translation unit is compiled
// with -fmad=false so nothing is contracted, divisions / sqrt are IEEE, and ln() comes from a host-built table
// indexed by the integer visit count.
#pragma once
#include "game_kernels.cuh"
#include "eval_views.hpp"
#include "net_kernels.cuh"
#include "tak_device.cuh"

namespace tb {

constexpr int MCTS_MAX_DEPTH = 128;
constexpr int MCTS_EXPLO_TABLE = 1 << 20;

enum MctsErr : int {
    MERR_POOL_FULL = 1,      // node arena exhausted
    MERR_DEPTH = 2,          // path longer than MCTS_MAX_DEPTH
    MERR_PENDING_FULL = 4,   // more queued leaves than kcap
    MERR_VISITS = 8,         // visit count beyond the exploration table
    MERR_BAD_MOVE = 16,      // mcts_play with a move that is not a child / unindexable move
    MERR_NAN = 32,           // NaN upper confidence bound (the reference panics)
    MERR_QUEUED = 64,        // mcts_play on a game whose leaves are still queued
};

struct MctsView {
    uint4* stat;        // [G][2][cap]
    uint2* link;
    int* half;          // [G]
    uint32_t* top;      // [G]
    int* pend_cnt;      // [G]
    uint32_t* pend_leaf;  // [G][kcap]
    int* pend_plen;       // [G][kcap]
    uint32_t* pend_path;  // [G][kcap][MCTS_MAX_DEPTH]
    uint8_t* leaf_states; // [G][kcap][S]
    const float* explo;   // [MCTS_EXPLO_TABLE]
    const uint16_t* move_table;  // [65536] move -> policy index
    int* err;
    unsigned long long* counters;  // [0] rollouts started, [1] leaves queued
    int cap;
    int kcap;
};

__device__ __forceinline__ size_t arena_base(const MctsView& v, int gid, int half) {
    return (size_t(gid) * 2 + half) * size_t(v.cap);
}
__device__ __forceinline__ float stat_prior(const uint4& s) { return __uint_as_float(s.x); }
__device__ __forceinline__ float stat_reward(const uint4& s) { return __uint_as_float(s.y); }

// Node::update_concrete (mcts.rs:120-124)
__device__ __forceinline__ void update_concrete(uint4& s, float reward) {
    const float cumulative = __fmul_rn(__uint_as_float(s.y), float(s.z));
    s.z += 1;
    s.y = __float_as_uint(__fdiv_rn(__fadd_rn(cumulative, reward), float(s.z)));
}

// Node::virtual_rollout x k (mcts.rs:26-65) incl. select (mcts.rs:94-118) for ONE game, executed by one warp
template <int N>
__device__ __forceinline__ void rollout_game(const MctsView& v, const uint8_t* states, int gid, int k,
                                             const FastEval& fe) {
    const int l = threadIdx.x & 31;
    constexpr int S = StateLayout<N>::S;
    const int half = v.half[gid];
    uint4* stat = v.stat + arena_base(v, gid, half);
    uint2* link = v.link + arena_base(v, gid, half);
    uint32_t top = v.top[gid];
    int pend = v.pend_cnt[gid];

    for (int rep = 0; rep < k; ++rep) {
        if (pend >= v.kcap) {  // the path of this rollout is built in the next free queue slot
            if (l == 0) atomicOr(v.err, MERR_PENDING_FULL);
            break;
        }
        WarpGame<N> g;
        g.load(states + size_t(gid) * S);
        const int root_color = g.to_move;
        uint32_t* path = v.pend_path + (size_t(gid) * v.kcap + pend) * MCTS_MAX_DEPTH;
        uint32_t node = 0;
        int depth = 0;
        uint8_t res = RES_ONGOING;
        bool fail = false;
        if (l == 0) path[0] = 0;
        for (;;) {
            const uint4 s = stat[node];
            const uint2 lk = link[node];
            const bool initialized = s.z != 0 || s.w != 0;
            if (!initialized) {
                // uninitialised node: evaluate the position and create the children (mcts.rs:39-51)
                res = g.result();
                uint32_t base = 0;
                int count = 0;
                if (res == RES_ONGOING) {
                    count = g.count_total();
                    if (top + uint32_t(count) > uint32_t(v.cap) || count > 0xFFFF) {
                        if (l == 0) atomicOr(v.err, MERR_POOL_FULL);
                        fail = true;
                        break;
                    }
                    base = top;
                    top += uint32_t(count);
                    const float temp_policy = __fdiv_rn(1.0f, float(count));
                    const uint4 cs = make_uint4(__float_as_uint(temp_policy), 0u, 0u, 0u);
                    g.generate([&](int i, uint16_t mv) {
                        stat[base + i] = cs;
                        link[base + i] = make_uint2(0u, uint32_t(mv));
                    });
                }
                if (l == 0) link[node] = make_uint2(base | (uint32_t(res) << 24), (lk.y & 0xFFFFu) | (uint32_t(count) << 16));
                break;
            }
            res = uint8_t(lk.x >> 24);
            if (res != RES_ONGOING) break;  // cached terminal result (mcts.rs:35-37)
            // ---- select (mcts.rs:94-118): PUCT, LAST maximum wins ties
            const uint32_t base = lk.x & 0xFFFFFFu;
            const int nchild = int(lk.y >> 16);
            const uint32_t nvis = s.z + s.w;
            if (nvis >= uint32_t(MCTS_EXPLO_TABLE)) {
                if (l == 0) atomicOr(v.err, MERR_VISITS);
                fail = true;
                break;
            }
            const float visit_count = float(nvis);
            const float explo = v.explo[nvis];
            const float sq = __fsqrt_rn(visit_count);
            float best_u = -INFINITY;
            int best_i = -1;
            for (int i = l; i < nchild; i += 32) {
                const uint4 cs = stat[base + i];
                const uint32_t cvis = cs.z + cs.w;
                float q = 0.0f;
                if (cvis != 0)
                    q = __fdiv_rn(__fsub_rn(__fmul_rn(__uint_as_float(cs.y), float(cs.z)), float(cs.w)), float(cvis));
                const float u = __fadd_rn(
                    q, __fmul_rn(__fmul_rn(explo, __uint_as_float(cs.x)), __fdiv_rn(sq, __fadd_rn(1.0f, float(cvis)))));
                if (u != u) atomicOr(v.err, MERR_NAN);
                if (u >= best_u) { best_u = u; best_i = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ou = __shfl_xor_sync(FULL, best_u, o);
                const int oi = __shfl_xor_sync(FULL, best_i, o);
                if (ou > best_u || (ou == best_u && oi > best_i)) { best_u = ou; best_i = oi; }
            }
            if (best_i < 0 || depth + 1 >= MCTS_MAX_DEPTH) {
                if (l == 0) atomicOr(v.err, best_i < 0 ? MERR_NAN : MERR_DEPTH);
                fail = true;
                break;
            }
            const uint32_t child = base + uint32_t(best_i);
            const uint16_t mv = uint16_t(link[child].y & 0xFFFFu);
            g.template play<false>(mv);
            ++depth;
            if (l == 0) path[depth] = child;
            node = child;
        }
        __syncwarp();
        if (fail) break;
        // ---- unwind (mcts.rs:53-64): every node on the path applies the same result from its own perspective
        if (l == 0) {
            for (int d = depth; d >= 0; --d) {
                const uint32_t nd = path[d];
                uint4 s = stat[nd];
                if (res == RES_ONGOING) {
                    s.w += 1;
                } else if ((res & 0xF) == RES_DRAW) {
                    update_concrete(s, 0.0f);
                } else {
                    const int winner = (res & 0xF) == RES_WHITE ? 0 : 1;
                    const int curr_color = root_color ^ (d & 1);
                    update_concrete(s, winner == curr_color ? -1.0f : 1.0f);
                }
                stat[nd] = s;
            }
            atomicAdd(v.counters, 1ull);
        }
        if (res == RES_ONGOING) {
            g.store(v.leaf_states + (size_t(gid) * v.kcap + pend) * S);
            int slot = 0;
            if (l == 0) {
                v.pend_leaf[size_t(gid) * v.kcap + pend] = node;
                v.pend_plen[size_t(gid) * v.kcap + pend] = depth;
                atomicAdd(v.counters + 1, 1ull);
                if (fe.eval_count) {
                    slot = atomicAdd(fe.eval_count, 1);
                    fe.eval_slot[size_t(gid) * v.kcap + pend] = slot;
                }
            }
            if (fe.eval_count && fe.planes) {
                slot = __shfl_sync(FULL, slot, 0);
                encode_board<N>(g, slot, fe.planes, fe.S);
            }
            ++pend;
        }
        __syncwarp();
    }
    if (l == 0) {
        v.top[gid] = top;
        v.pend_cnt[gid] = pend;
    }
}

template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_rollout(MctsView v, const uint8_t* states, const int* ids, int n, int k, const uint8_t* enable) {
    const int w = warp_global_id();
    if (w >= n) return;
    if (enable && !enable[w]) return;
    FastEval none{};
    rollout_game<N>(v, states, ids ? ids[w] : w, k, none);
}

// flat pending slot (gid*kcap + j) <-> compact evaluation index; single block
static __global__ void __launch_bounds__(1024)
    k_mcts_compact(const int* pend_cnt, int n_games, int kcap, const int* limits, int* eval_index, int* eval_slot,
                   int* eval_count, int max_eval, int* err) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    if (threadIdx.x < 32) s_warp[threadIdx.x] = 0;   // the block may have fewer than 32 warps: unused entries scan as 0
    __syncthreads();
    for (int base = 0; base < n_games; base += blockDim.x) {
        const int g = base + threadIdx.x;
        // with `limits`, only the oldest limits[g] leaves of game g are taken (Player's pipelined batches)
        const int c = g < n_games ? (limits ? min(pend_cnt[g], limits[g]) : pend_cnt[g]) : 0;
        int inc = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, s);
            if ((threadIdx.x & 31) >= s) inc += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int wv = s_warp[threadIdx.x];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(FULL, wv, s);
                if (threadIdx.x >= s) wv += t;
            }
            s_warp[threadIdx.x] = wv;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int excl = s_carry + warp_off + inc - c;
        for (int j = 0; j < c; ++j) {
            eval_slot[g * kcap + j] = excl + j;
            if (excl + j < max_eval) eval_index[excl + j] = g * kcap + j;   // beyond one network batch: flagged below
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // the evaluation is launched for at most max_eval leaves and reads this count on the device; more queued leaves
        // than one network batch holds is an error the caller reports (TAK_ERR_CAPACITY)
        if (s_carry > max_eval) atomicOr(err, MERR_PENDING_FULL);
        *eval_count = min(s_carry, max_eval);
    }
}


// Node::devirtualize_path (mcts.rs:67-91) for ONE queued leaf (flat queue slot `slot`, evaluation slot `ei`)
template <int N>
__device__ __forceinline__ void backup_leaf(const MctsView& v, uint4* stat, const uint2* link, size_t slot, int ei,
                                            const PriorSource& ps, float mx, float sum, float eval) {
    constexpr int NSQ = N * N;
    const int l = threadIdx.x & 31;
    const uint32_t leaf = v.pend_leaf[slot];
    const int plen = v.pend_plen[slot];
    const uint32_t* path = v.pend_path + slot * MCTS_MAX_DEPTH;
    // replace the temporary priors (mcts.rs:78-83): policy[move_index(mov)] -- no mask, no renormalisation
    const uint2 lk = link[leaf];
    const uint32_t base = lk.x & 0xFFFFFFu;
    const int nchild = int(lk.y >> 16);
    for (int i = l; i < nchild; i += 32) {
        const uint16_t mv = uint16_t(link[base + i].y & 0xFFFFu);
        const int idx = v.move_table[mv];
        float prior = 1.0f;
        if (idx == 0xFFFF) {
            atomicOr(v.err, MERR_BAD_MOVE);
        } else if (ps.arch == 6) {
            const int ch = idx / NSQ, sq = idx % NSQ, row = sq / N, col = sq % N;
            const float lg = ps.logits[size_t(ch) * ps.S + SlotMap<N, INFER_PF>::slot(ei, row, col)];
            prior = __fdiv_rn(expf(__fsub_rn(lg, mx)), sum);
        } else if (ps.arch == 5) {
            prior = __fdiv_rn(expf(__fsub_rn(ps.logits[size_t(ei) * ps.psz + idx], mx)), sum);
        } else if (ps.arch == -1) {
            prior = ps.logits[size_t(ei) * ps.psz + idx];
        }
        uint4 cs = stat[base + i];
        cs.x = __float_as_uint(prior);
        stat[base + i] = cs;
    }
    __syncwarp();
    if (l == 0) {
        for (int d = plen; d >= 0; --d) {
            const uint32_t nd = path[d];
            uint4 s = stat[nd];
            s.w -= 1;
            eval = -eval;
            update_concrete(s, eval);
            stat[nd] = s;
        }
    }
    __syncwarp();
}

// Node::devirtualize_path for every queued leaf of every game, in queue order
template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_backup(MctsView v, const int* eval_slot, int n_games, PriorSource ps, const int* limits,
                  uint8_t* leaf_states) {
    const int gid = warp_global_id();
    if (gid >= n_games) return;
    const int queued = v.pend_cnt[gid];
    // Player's pipelining backs up the OLDER batch while a newer one stays queued
    const int cnt = limits ? min(queued, limits[gid]) : queued;
    if (cnt == 0) return;
    const int l = threadIdx.x & 31;
    const int half = v.half[gid];
    uint4* stat = v.stat + arena_base(v, gid, half);
    uint2* link = v.link + arena_base(v, gid, half);
    for (int j = 0; j < cnt; ++j) {
        const size_t slot = size_t(gid) * v.kcap + j;
        const int ei = eval_slot[slot];
        float mx = 0.f, sum = 1.f;
        if (ps.arch == 5 || ps.arch == 6) {
            const float2 st = ps.stats[ei];
            mx = st.x;
            sum = st.y;
        }
        backup_leaf<N>(v, stat, link, slot, ei, ps, mx, sum, ps.arch == 0 ? 0.0f : ps.values[ei]);
    }
    // leaves queued after the first `limit` stay queued: move them to the front of the game's queue
    const int rest = queued - cnt;
    if (rest > 0) {
        constexpr int S = StateLayout<N>::S;
        for (int j = 0; j < rest; ++j) {
            const size_t dst = size_t(gid) * v.kcap + j, src = dst + cnt;
            const int plen = v.pend_plen[src];
            __syncwarp();
            for (int d = l; d <= plen; d += 32) v.pend_path[dst * MCTS_MAX_DEPTH + d] = v.pend_path[src * MCTS_MAX_DEPTH + d];
            for (int b = l * 16; b < S; b += 32 * 16)
                *reinterpret_cast<uint4*>(leaf_states + dst * S + b) = *reinterpret_cast<const uint4*>(leaf_states + src * S + b);
            if (l == 0) {
                v.pend_leaf[dst] = v.pend_leaf[src];
                v.pend_plen[dst] = plen;
            }
            __syncwarp();
        }
    }
    if (l == 0) v.pend_cnt[gid] = rest;
}

// One iteration of the fused search loop for every listed game: devirtualise the leaf the previous iteration queued
// (its network outputs are in place), then run the next virtual rollout and queue / encode its leaf.  Both halves touch
// only this game's tree, so they need no grid-wide ordering and live in one launch: the loop is {k_mcts_step; tower} x R.
// With the DummyNet (arch 0: prior 1, eval 0 -- nothing to evaluate) the whole loop of `reps` rollouts is ONE launch.
// MINB = minimum resident blocks per SM the compiler must allow for (register cap = 65536 / (MAXT * MINB)): the default
// build lets the kernel have its 128 registers; the capped builds (TAK_STEP_REGS) were made to measure whether a smaller
// block runs beside a resident conv-tower CTA (it does not: profiles/r02_step_overlap.md).
template <int N, int MAXT = 128, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
    k_mcts_step(MctsView v, const uint8_t* states, const int* ids, int n, const uint8_t* enable, FastEval fe,
                PriorSource ps, int do_backup, int do_rollout, int reps) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && fe.eval_count_reset) *fe.eval_count_reset = 0;
    const int w = warp_global_id();
    if (w >= n) return;
    const int gid = ids ? ids[w] : w;
    const bool roll = do_rollout && (!enable || enable[w]);
    for (int rep = 0; rep < reps; ++rep) {
        if (do_backup && v.pend_cnt[gid] > 0) {
            const int half = v.half[gid];
            uint4* stat = v.stat + arena_base(v, gid, half);
            const uint2* link = v.link + arena_base(v, gid, half);
            const size_t slot = size_t(gid) * v.kcap;
            const int ei = ps.arch != 0 ? fe.eval_slot[slot] : 0;
            float mx = 0.f, sum = 1.f, eval = 0.f;
            if (ps.arch == 6) {
                const float2 st = warp_policy_stats<N, INFER_PF>(ps.partials, ps.S, ps.groups, ei);
                mx = st.x;
                sum = st.y;
            } else if (ps.arch == 5) {
                const float2 st = ps.stats[ei];
                mx = st.x;
                sum = st.y;
            }
            if (ps.arch != 0) eval = warp_value<N, INFER_PF>(ps.trunk, ps.S, ps.value_w, ps.value_b, ei);
            backup_leaf<N>(v, stat, link, slot, ei, ps, mx, sum, eval);
            if ((threadIdx.x & 31) == 0) v.pend_cnt[gid] = 0;
            __syncwarp();
        }
        if (roll) rollout_game<N>(v, states, gid, 1, fe);
    }
}

// Node::pick_move(true) (play.rs:52-58): LAST child with the maximal visit count.  With `sample` != 0 the move is
// drawn with probability ~ visits (play.rs:60-65) from a counter-based RNG.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static __global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_pick(MctsView v, const int* ids, int n, const uint8_t* sample_flags, uint64_t seed, const int* game_tags,
                uint16_t* out_moves) {
    const int w = warp_global_id();
    if (w >= n) return;
    const int gid = ids ? ids[w] : w;
    const int l = threadIdx.x & 31;
    const int half = v.half[gid];
    const uint4* stat = v.stat + arena_base(v, gid, half);
    const uint2* link = v.link + arena_base(v, gid, half);
    const uint2 lk = link[0];
    const uint32_t base = lk.x & 0xFFFFFFu;
    const int nchild = int(lk.y >> 16);
    if (nchild == 0) {
        if (l == 0) { out_moves[w] = 0xFFFF; atomicOr(v.err, MERR_BAD_MOVE); }
        return;
    }
    int pick = -1;
    if (sample_flags && sample_flags[w]) {
        // visit-weighted sample: r uniform in [0, total) -> first child whose cumulative count exceeds r
        unsigned long long total = 0;
        for (int i = l; i < nchild; i += 32) total += stat[base + i].z;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
        if (total > 0) {
            const uint64_t r = splitmix64(seed ^ splitmix64(uint64_t(game_tags ? game_tags[w] : gid))) % total;
            unsigned long long run = 0;
            for (int i0 = 0; i0 < nchild && pick < 0; i0 += 32) {
                const int i = i0 + l;
                unsigned long long c = i < nchild ? stat[base + i].z : 0;
                unsigned long long inc = c;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const unsigned long long t = __shfl_up_sync(FULL, inc, s);
                    if (l >= s) inc += t;
                }
                const bool hit = c > 0 && run + inc > r;
                const unsigned m = __ballot_sync(FULL, hit);
                if (m) pick = i0 + __ffs(m) - 1;
                run += __shfl_sync(FULL, inc, 31);
            }
        }
    }
    if (pick < 0) {
        uint32_t best_v = 0;
        int best_i = -1;
        for (int i = l; i < nchild; i += 32) {
            const uint32_t vis = stat[base + i].z;
            if (best_i < 0 || vis >= best_v) { best_v = vis; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t ov = __shfl_xor_sync(FULL, best_v, o);
            const int oi = __shfl_xor_sync(FULL, best_i, o);
            if (oi >= 0 && (best_i < 0 || ov > best_v || (ov == best_v && oi > best_i))) { best_v = ov; best_i = oi; }
        }
        pick = best_i;
    }
    if (l == 0) out_moves[w] = uint16_t(link[base + pick].y & 0xFFFFu);
}

// Node::play (play.rs:26-43): the chosen child becomes the root; its subtree is copied breadth-first into the
// other half of the arena (tree reuse + compaction), everything else is dropped.
static __global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_reroot(MctsView v, const int* ids, const uint16_t* moves, int n) {
    const int w = warp_global_id();
    if (w >= n) return;
    const int gid = ids ? ids[w] : w;
    const int l = threadIdx.x & 31;
    if (v.pend_cnt[gid] != 0) {  // queued paths would dangle: the caller must devirtualize this game first
        if (l == 0) atomicOr(v.err, MERR_QUEUED);
        return;
    }
    const int half = v.half[gid];
    const uint4* ostat = v.stat + arena_base(v, gid, half);
    const uint2* olink = v.link + arena_base(v, gid, half);
    uint4* nstat = v.stat + arena_base(v, gid, half ^ 1);
    uint2* nlink = v.link + arena_base(v, gid, half ^ 1);
    const uint2 rlk = olink[0];
    const uint32_t rbase = rlk.x & 0xFFFFFFu;
    const int rn = int(rlk.y >> 16);
    const uint16_t mv = moves[w];
    int found = -1;
    for (int i0 = 0; i0 < rn && found < 0; i0 += 32) {
        const int i = i0 + l;
        const bool hit = i < rn && uint16_t(olink[rbase + i].y & 0xFFFFu) == mv;
        const unsigned m = __ballot_sync(FULL, hit);
        if (m) found = i0 + __ffs(m) - 1;
    }
    if (found < 0) {
        if (l == 0) atomicOr(v.err, MERR_BAD_MOVE);
        return;
    }
    if (l == 0) {
        nstat[0] = ostat[rbase + found];
        nlink[0] = olink[rbase + found];
    }
    __syncwarp();
    uint32_t q = 0, top = 1;
    while (q < top) {
        // examine up to 32 already-copied nodes at once; those with children get their block copied in order
        const uint32_t i = q + l;
        uint2 lk = make_uint2(0, 0);
        if (i < top) lk = nlink[i];
        const bool has = i < top && (lk.y >> 16) != 0;
        unsigned m = __ballot_sync(FULL, has);
        const uint32_t batch_end = (top - q) < 32u ? top : q + 32u;
        while (m) {
            const int src_lane = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t obase = __shfl_sync(FULL, lk.x, src_lane) & 0xFFFFFFu;
            const uint32_t ores = __shfl_sync(FULL, lk.x, src_lane) & 0xFF000000u;
            const uint32_t meta = __shfl_sync(FULL, lk.y, src_lane);
            const int cnt = int(meta >> 16);
            const uint32_t nbase = top;
            for (int c = l; c < cnt; c += 32) {
                nstat[nbase + c] = ostat[obase + c];
                nlink[nbase + c] = olink[obase + c];
            }
            if (l == 0) nlink[q + src_lane] = make_uint2(nbase | ores, meta);
            top += uint32_t(cnt);
        }
        __syncwarp();
        q = batch_end;
    }
    if (l == 0) {
        v.half[gid] = half ^ 1;
        v.top[gid] = top;
        v.pend_cnt[gid] = 0;
    }
}

static __global__ void k_mcts_tree_reset(MctsView v, const int* ids, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gid = ids ? ids[i] : i;
    const size_t b = arena_base(v, gid, v.half[gid]);
    v.stat[b] = make_uint4(0, 0, 0, 0);
    v.link[b] = make_uint2(0, 0);
    v.top[gid] = 1;
    v.pend_cnt[gid] = 0;
}

// root + root children of one game into a staging buffer: [0] = root stat/link, then children
static __global__ void k_mcts_export_root(MctsView v, int gid, uint4* out_stat, uint2* out_link, int cap, int* out_count) {
    const int half = v.half[gid];
    const uint4* stat = v.stat + arena_base(v, gid, half);
    const uint2* link = v.link + arena_base(v, gid, half);
    const uint2 lk = link[0];
    const uint32_t base = lk.x & 0xFFFFFFu;
    const int nchild = int(lk.y >> 16);
    if (threadIdx.x == 0) {
        out_stat[0] = stat[0];
        out_link[0] = lk;
        *out_count = nchild;
    }
    for (int i = threadIdx.x; i < nchild && i + 1 < cap; i += blockDim.x) {
        out_stat[1 + i] = stat[base + i];
        out_link[1 + i] = link[base + i];
    }
}

// Node::improved_policy (play.rs:13-22) for many games: moves / visits of the root children, `stride` per game
static __global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_export_children(MctsView v, const int* ids, int n, uint16_t* out_moves, uint32_t* out_visits,
                           int* out_counts, int stride) {
    const int w = warp_global_id();
    if (w >= n) return;
    const int gid = ids ? ids[w] : w;
    const int l = threadIdx.x & 31;
    const int half = v.half[gid];
    const uint4* stat = v.stat + arena_base(v, gid, half);
    const uint2* link = v.link + arena_base(v, gid, half);
    const uint2 lk = link[0];
    const uint32_t base = lk.x & 0xFFFFFFu;
    const int nchild = int(lk.y >> 16);
    if (l == 0) out_counts[w] = nchild;
    for (int i = l; i < nchild && i < stride; i += 32) {
        out_moves[size_t(w) * stride + i] = uint16_t(link[base + i].y & 0xFFFFu);
        out_visits[size_t(w) * stride + i] = stat[base + i].z;
    }
}

// Player::rollout (player.rs:130-133) keeps ONE batch in flight: limits[g] = leaves game g has queued right now, i.e. the
// batch requested by the previous call -- exactly what the consume half of this call must back up after the new batch
// has been selected.  (limits[] is zeroed first: games not listed keep their queues untouched.)
static __global__ void k_mcts_snapshot_limits(const int* pend_cnt, const int* ids, int n, int* limits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) limits[ids[i]] = pend_cnt[ids[i]];
}

// Node::debug / Node::continuation (search/debug.rs:9-40): one warp per root child.  The record carries the child's
// (move, visits, expected_reward, policy) and its principal continuation: from the child, repeatedly the child with the
// most visits (pick_move(true): LAST maximum, play.rs:49-58) until `depth` moves or a childless node.
constexpr int MCTS_DEBUG_DEPTH = 16;
struct MctsMoveInfo {
    uint16_t move;
    uint16_t cont_len;
    uint32_t visits;
    float reward;
    float policy;
    uint16_t cont_moves[MCTS_DEBUG_DEPTH];
    uint32_t cont_visits[MCTS_DEBUG_DEPTH];
};
static __global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_debug(MctsView v, int gid, int depth, MctsMoveInfo* out, int cap, int* out_count) {
    const int w = warp_global_id();
    const int l = threadIdx.x & 31;
    const int half = v.half[gid];
    const uint4* stat = v.stat + arena_base(v, gid, half);
    const uint2* link = v.link + arena_base(v, gid, half);
    const uint2 root = link[0];
    const int nroot = int(root.y >> 16);
    if (w == 0 && l == 0) *out_count = nroot;
    if (w >= nroot || w >= cap) return;
    uint32_t node = (root.x & 0xFFFFFFu) + uint32_t(w);
    const uint4 s = stat[node];
    if (l == 0) {
        out[w].move = uint16_t(link[node].y & 0xFFFFu);
        out[w].visits = s.z;
        out[w].reward = __uint_as_float(s.y);
        out[w].policy = __uint_as_float(s.x);
    }
    int len = 0;
    for (; len < depth && len < MCTS_DEBUG_DEPTH; ++len) {
        const uint2 lk = link[node];
        const uint32_t base = lk.x & 0xFFFFFFu;
        const int nchild = int(lk.y >> 16);
        if (nchild == 0) break;
        uint32_t best_v = 0;
        int best_i = -1;
        for (int i = l; i < nchild; i += 32) {
            const uint32_t vis = stat[base + i].z;
            if (best_i < 0 || vis >= best_v) { best_v = vis; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t ov = __shfl_xor_sync(FULL, best_v, o);
            const int oi = __shfl_xor_sync(FULL, best_i, o);
            if (oi >= 0 && (best_i < 0 || ov > best_v || (ov == best_v && oi > best_i))) { best_v = ov; best_i = oi; }
        }
        node = base + uint32_t(best_i);
        if (l == 0) {
            out[w].cont_moves[len] = uint16_t(link[node].y & 0xFFFFu);
            out[w].cont_visits[len] = best_v;
        }
    }
    if (l == 0) out[w].cont_len = uint16_t(len);
}

// Node::apply_dirichlet (noise.rs:6-16): prior = noise*ratio + prior*(1-ratio), noise ~ Dirichlet(alpha) from a
// counter-based RNG (Marsaglia-Tsang gamma with the alpha<1 boost); statistical, not bit-reproducible vs rand 0.8.
__device__ __forceinline__ float u01(uint64_t& s) {
    s = splitmix64(s);
    return (float((s >> 40) & 0xFFFFFF) + 0.5f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float sample_gamma(float alpha, uint64_t& s) {
    const float a = alpha < 1.0f ? alpha + 1.0f : alpha;
    const float d = a - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
    float x;
    for (int it = 0; it < 64; ++it) {
        const float u1 = u01(s), u2 = u01(s);
        const float nrm = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        const float t = 1.0f + c * nrm;
        if (t <= 0.0f) continue;
        const float vv = t * t * t;
        const float u = u01(s);
        x = d * vv;
        if (logf(u) < 0.5f * nrm * nrm + d - x + d * logf(vv)) break;
    }
    if (alpha < 1.0f) x *= powf(u01(s), 1.0f / alpha);
    return fmaxf(x, 1e-30f);
}
static __global__ void __launch_bounds__(GAME_THREADS)
    k_mcts_dirichlet(MctsView v, const int* ids, int n, const uint8_t* enable, float alpha, float ratio, uint64_t seed,
                     const int* game_tags) {
    const int w = warp_global_id();
    if (w >= n) return;
    if (enable && !enable[w]) return;
    const int gid = ids ? ids[w] : w;
    const int l = threadIdx.x & 31;
    const int half = v.half[gid];
    uint4* stat = v.stat + arena_base(v, gid, half);
    const uint2 lk = v.link[arena_base(v, gid, half)];
    const uint32_t base = lk.x & 0xFFFFFFu;
    const int nchild = int(lk.y >> 16);
    if (stat[0].z == 0) {  // "cannot apply dirichlet noise without initialized policy" (noise.rs:7-10)
        if (l == 0) atomicOr(v.err, MERR_BAD_MOVE);
        return;
    }
    float sum = 0.f;
    for (int i = l; i < nchild; i += 32) {
        uint64_t s = seed ^ splitmix64((uint64_t(game_tags ? game_tags[w] : gid) << 20) ^ uint64_t(i));
        sum += sample_gamma(alpha, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    for (int i = l; i < nchild; i += 32) {
        uint64_t s = seed ^ splitmix64((uint64_t(game_tags ? game_tags[w] : gid) << 20) ^ uint64_t(i));
        const float noise = sample_gamma(alpha, s) / sum;
        uint4 cs = stat[base + i];
        cs.x = __float_as_uint(noise * ratio + __uint_as_float(cs.x) * (1.0f - ratio));
        stat[base + i] = cs;
    }
}

}  // namespace tb
