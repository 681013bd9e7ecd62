// Network state of an engine (weights packed for conv_tc3, activation strip planes, head outputs).
#pragma once
#include <cuda_bf16.h>

#include "engine.hpp"
#include "eval_views.hpp"

namespace tb {

struct ConvLayer {
    DevBuf w;     // bf16 [slab 8][ky 3][kx 3][kchunk 2][c_out 128][8 c_in]
    DevBuf bias;  // fp32 [128]
};

struct NetProfile {
    cudaEvent_t ev[160];
    int n = 0;
};

struct NetState {
    NetProfile* profile = nullptr;         // set only by net_forward_profile
    int arch = 0;          // 0 DummyNet, 5 Net5, 6 Net6
    int n = 0;
    int c_in = 0;          // input_channels(n)
    int blocks = 0;        // residual blocks
    int policy_ch = 0;     // Net6: move_channels(6) = 251
    int policy_groups = 0; // Net6: 2 groups of 128 output channels
    int policy_out = 0;    // policy vector length (1575 / 9036)
    bool loaded = false;
    std::vector<ConvLayer> layers;         // initial conv + 2 per block (BN folded)
    std::vector<ConvLayer> policy_layers;  // Net6 policy conv, one per group
    DevBuf fc_policy_w, fc_policy_b;       // Net5 policy FC: bf16 operand image Wp[jt][pos][slab][2][128][8] (fc_tc.cuh), fp32 bias
    DevBuf fc_x;                           // Net5: trunk output repacked as X[pos][chunk][board][8] for the FC GEMM
    DevBuf value_w;                        // fp32 [128*NSQ] (NCHW flatten order)
    float value_bias = 0.f;
    // activations
    int cap_boards = 0, cap_S = 0;
    DevBuf act[3];                         // bf16 strip planes [16][S][8], rotating through the tower (each is fully
                                           // rewritten by a conv epilogue before it is read: dead tiles are L2-discarded)
    DevBuf act_in;                         // the encoded input planes (k_encode writes only real squares: never discarded)
    DevBuf logits;                         // Net6: fp32 [256][S]; Net5: fp32 [B][1575]
    DevBuf partials;                       // Net6: float2 [groups][S] per-slot softmax partials (conv epilogue)
    DevBuf stats;                          // float2 {max, sum exp} per board
    DevBuf values;                         // fp32 [B]
    const __nv_bfloat16* trunk_out = nullptr;
    // staging for the host-facing API
    DevBuf stage_states, stage_policy, stage_repr;
    std::vector<float> blob_host;          // the fp32 weight blob last loaded (net_train_begin starts from it)
    void* train = nullptr;                 // TrainState (train.cu), owned; released by train_destroy
};

// Evaluate `boards` packed states (d_states[index[i]] or d_states[i] if index == nullptr) on the engine stream.
// Leaves logits / stats / values on device; optionally writes the full softmax policy [boards][policy_out].
// d_count != nullptr: `boards` sizes the launches, the live number of boards is read from device memory
int net_forward(tak_engine* e, const uint8_t* d_states, const int* d_index, int boards, float* d_policy_out,
                int raw_logits = 0, const int* d_count = nullptr);
int net_ensure_capacity(tak_engine* e, int boards);
// fused search loop (mcts.cu): tower over the planes the rollout warps wrote, board count read on the device
int net_tower_fast(tak_engine* e, int max_boards, const int* d_count);
int net_fast_views(tak_engine* e, int max_boards, FastEval& fe, PriorSource& ps);
int net_load_blob(tak_engine* e, const float* blob, int64_t elems);
int64_t net_blob_elems(const NetState& ns);
void train_destroy(tak_engine* e);   // train.cu

}  // namespace tb
