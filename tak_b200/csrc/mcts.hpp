// Host-side state of the device MCTS (node pools, pending-leaf queues) and the internal entry points the
// self-play loop shares with the C ABI.
#pragma once
#include "engine.hpp"
#include "mcts_kernels.cuh"

namespace tb {

struct MctsState {
    int cap = 0;    // nodes per game per half
    int kcap = 0;   // queued leaves per game
    DevBuf stat, link, half, top, pend_cnt, pend_leaf, pend_plen, pend_path, leaf_states;
    DevBuf eval_index, eval_slot, eval_count, explo, move_table, err, counters, limits;
    bool limits_on = false;   // compact / backup take only the oldest limits[g] leaves of game g (mcts_devirtualize_first)
    DevBuf d_ids, d_moves, stage_policy, stage_value, stage_stat, stage_link, stage_count;
    int* h_pinned = nullptr;  // [4] pinned host words: eval count, error flags
    bool queued = false;      // leaves are waiting for devirtualize
    MctsView view() const;
};

int mcts_ensure(tak_engine* e, int k);
// d_ids == nullptr => games [0, n)
int mcts_launch_rollout(tak_engine* e, const int* d_ids, int n, int k, const uint8_t* d_enable = nullptr);
// fills eval_index / eval_slot / eval_count (device); clamp_to_batch: the count is capped at one network batch and an
// overflow raises the device error flag (the evaluation path reads the count on the device)
int mcts_launch_compact(tak_engine* e, bool clamp_to_batch = false);
int mcts_read_eval_count(tak_engine* e, int* out);      // syncs the stream
int mcts_launch_backup(tak_engine* e, const PriorSource& ps);
int mcts_eval_and_backup(tak_engine* e);                // compact -> network -> backup (syncs once for the count)
// `reps` x Node::rollout for the listed games, fused on the device (two launches per rollout, no host round trip)
int mcts_fast_rollouts(tak_engine* e, const int* d_ids, int n, int reps, const uint8_t* d_enable);
int mcts_check_errors(tak_engine* e);                   // syncs; maps device error flags to a status
int mcts_launch_tree_reset(tak_engine* e, const int* d_ids, int n);
int mcts_launch_pick(tak_engine* e, const int* d_ids, int n, const uint8_t* d_sample, uint64_t seed,
                     const int* d_tags, uint16_t* d_out);
int mcts_launch_reroot(tak_engine* e, const int* d_ids, const uint16_t* d_moves, int n);
int mcts_launch_dirichlet(tak_engine* e, const int* d_ids, int n, const uint8_t* d_enable, float alpha, float ratio,
                          uint64_t seed, const int* d_tags);

}  // namespace tb
