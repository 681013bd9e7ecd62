// Plain views shared by the search kernels (mcts_kernels.cuh) and the network module (net.cu): where a leaf's input planes
// go and where its network outputs are found.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace tb {

// Evaluation queue of the fused search loop (one leaf per game per iteration, mcts_rollouts / selfplay_step): the warp
// that reaches an un-evaluated leaf takes the next evaluation slot with one atomic, writes the leaf's input planes
// straight from its registers (no leaf-state round trip through HBM, no separate encode / compaction kernels), and the
// conv tower reads the slot count from device memory.  Slots are handed out in arrival order; the network's outputs do
// not depend on the slot a position sits in (tests/test_net_gpu.py), so the search results do not either.
struct FastEval {
    int* eval_count;          // the counter this step's leaves are appended to (nullptr: legacy compaction path)
    int* eval_count_reset;    // the OTHER phase's counter: consumed by the previous tower, zeroed for the next step
    int* eval_slot;           // [G * kcap] flat pending slot -> evaluation slot
    __nv_bfloat16* planes;    // input strip planes of the network (NetState::act_in), or nullptr (DummyNet)
    int S;
};

struct PriorSource {
    int arch;               // 0 dummy (prior 1, eval 0), 5 dense logits, 6 conv logits, -1 host-supplied policy
    const float* logits;    // arch 5: [B][psz]; arch 6: [ch][S]; arch -1: policy [B][psz]
    const float2* stats;    // {max, sum} per compact index (arch 5/6)
    const float* values;    // per compact index
    int S;
    int psz;
    // the fused loop computes the heads of ITS leaf inside the backup warp (same code as k_policy_stats_conv / k_value):
    const float2* partials;          // arch 6: per-slot softmax partials of the conv epilogue
    int groups;
    const __nv_bfloat16* trunk;      // trunk output strip planes (value head input)
    const float* value_w;
    float value_b;
};

}  // namespace tb
