// Engine object behind the opaque tak_engine_t handle of include/taknative.h.  Owns the CUDA stream and all
// device memory: packed game states, perft frontiers, network weights/activations, MCTS node pools.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/taknative.h"

namespace tb {

void set_error(const char* fmt, ...);

#define TB_CUDA(expr)                                                                            \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            tb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TAK_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define TB_CHECK(cond, code, ...)      \
    do {                               \
        if (!(cond)) {                 \
            tb::set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

// dispatch a template on the runtime board size
#define TB_DISPATCH_N(n, ...)                                 \
    switch (n) {                                              \
        case 3: { constexpr int N_ = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int N_ = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int N_ = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int N_ = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int N_ = 7; __VA_ARGS__; } break; \
        case 8: { constexpr int N_ = 8; __VA_ARGS__; } break; \
        default: break;                                       \
    }

// simple growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct NetState;      // net.cu
struct MctsState;     // mcts.cu
struct SelfplayState; // selfplay.cu
struct ExamplesState; // examples.cu
struct CommState;     // comm.cu

}  // namespace tb

struct tak_engine {
    int device = 0;
    int n = 0;
    int nsq = 0;
    int state_bytes = 0;  // S
    int max_games = 0;
    int nodes_per_game = 0;
    int max_batch = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;

    tb::DevBuf states;       // [max_games][S]
    // scratch for the batched Game API
    tb::DevBuf d_ids, d_moves, d_counts, d_status, d_results, d_stage;
    void* h_stage = nullptr;  // pinned host staging for upload / download
    size_t h_stage_bytes = 0;
    cudaError_t ensure_pinned(size_t need) {
        if (need <= h_stage_bytes) return cudaSuccess;
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        h_stage_bytes = 0;
        cudaError_t err = cudaMallocHost(&h_stage, need);
        if (err == cudaSuccess) h_stage_bytes = need;
        return err;
    }
    // perft
    static constexpr int PF_LEVELS = 13;
    struct PerftLevel {
        tb::DevBuf children, counts, offsets, moves, block_parent;
    };
    PerftLevel pf_level[PF_LEVELS];
    tb::DevBuf pf_root, pf_root_counts, pf_scan_tmp, pf_leaves;
    double pf_ms = 0;
    uint64_t pf_materialised = 0, pf_launches = 0;
    // CUDA-event spans around each (k_perft_moves, k_perft_apply) pair of the last tak_perft (tak_perft_profile)
    static constexpr int PF_SPANS = 64;
    cudaEvent_t pf_span_ev[2 * PF_SPANS] = {};
    uint64_t pf_span_children[PF_SPANS] = {};
    int pf_n_spans = 0;
    double pf_expand_ms = 0, pf_top_span_ms = 0;
    uint64_t pf_top_span_children = 0;
    // random playouts (tak_playouts)
    tb::DevBuf po_plies, po_result, po_totals;

    tb::NetState* net = nullptr;
    tb::MctsState* mcts = nullptr;
    tb::SelfplayState* selfplay = nullptr;
    tb::ExamplesState* examples = nullptr;
    tb::CommState* comm = nullptr;
    uint64_t launches = 0;  // kernels launched by this engine (gpu_launches in bench.py)
};

namespace tb {
// host <-> packed conversion (engine_core.cu)
void pack_state(int n, const tak_state_t& s, uint8_t* rec);
void unpack_state(int n, const uint8_t* rec, tak_state_t& s);
int state_bytes_for(int n);
// module teardown hooks
void net_destroy(tak_engine* e);
void mcts_destroy(tak_engine* e);
void selfplay_destroy(tak_engine* e);
void examples_destroy(tak_engine* e);
void comm_destroy(tak_engine* e);
// host helper shared by modules: policy index of a move (alpha_tak::search::move_index)
int host_move_index(int n, uint16_t mv);
int host_policy_size(int n);
const std::vector<uint16_t>& host_move_index_table(int n);  // [65536] move -> index (0xFFFF invalid)
}  // namespace tb
