// Batched Tak kernels over packed states: one warp per game.
//   k_reset / k_moves / k_play / k_result         -> tak_games_reset / tak_possible_moves / tak_play / tak_result
//   k_perft_count / k_perft_expand                -> tak_perft (breadth-first, perf_count rule of perft.rs:3-18)
#pragma once
#include "tak_device.cuh"

namespace tb {

constexpr int GAME_WARPS_PER_BLOCK = 8;
constexpr int GAME_THREADS = GAME_WARPS_PER_BLOCK * 32;

__device__ __forceinline__ int warp_global_id() { return (blockIdx.x * blockDim.x + threadIdx.x) >> 5; }

template <int N>
__global__ void __launch_bounds__(GAME_THREADS) k_reset(uint8_t* states, int first, int count, int half_komi) {
    const int w = warp_global_id();
    if (w >= count) return;
    WarpGame<N> g;
    g.reset(half_komi);
    g.store(states + size_t(first + w) * StateLayout<N>::S);
}

// moves of game ids[w] -> out_moves[w * stride ...], count -> out_counts[w]
template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_moves(const uint8_t* states, const int* ids, int n, uint16_t* out_moves, int* out_counts, int stride) {
    const int w = warp_global_id();
    if (w >= n) return;
    WarpGame<N> g;
    g.load(states + size_t(ids[w]) * StateLayout<N>::S);
    uint16_t* dst = out_moves + size_t(w) * stride;
    const int total = g.generate([&](int k, uint16_t mv) {
        if (k < stride) dst[k] = mv;
    });
    if ((threadIdx.x & 31) == 0) out_counts[w] = total;
}

template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_play(uint8_t* states, const int* ids, const uint16_t* moves, int n, int* out_status) {
    const int w = warp_global_id();
    if (w >= n) return;
    WarpGame<N> g;
    uint8_t* rec = states + size_t(ids[w]) * StateLayout<N>::S;
    g.load(rec);
    const int st = g.template play<true>(moves[w]);
    if (st == 0) g.store(rec);
    if ((threadIdx.x & 31) == 0) out_status[w] = st;
}

template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_result(const uint8_t* states, const int* ids, int n, uint8_t* out) {
    const int w = warp_global_id();
    if (w >= n) return;
    WarpGame<N> g;
    g.load(states + size_t(ids[w]) * StateLayout<N>::S);
    const uint8_t r = g.result();
    if ((threadIdx.x & 31) == 0) out[w] = r;
}

// staging[i] <-> states[ids[i]] in 16-byte words (to_states = 1: scatter into the engine, 0: gather out of it)
static __global__ void k_scatter_records(uint4* staging, uint4* states, const int* ids, int n, int words, int to_states) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * words) return;
    const int i = t / words, w = t % words;
    uint4* a = staging + size_t(i) * words + w;
    uint4* b = states + size_t(ids[i]) * words + w;
    if (to_states) *b = *a; else *a = *b;
}

// ---- perft -------------------------------------------------------------------------------------------------
// counts[w] = number of children parent w contributes to the next frontier (0 for finished games);
// `leaves` accumulates what perf_count returns for parents that stop here:
//   finished game -> 1 ; last level (depth_left == 1) -> number of legal moves.
template <int N>
__global__ void __launch_bounds__(256)
    k_perft_count(const uint8_t* frontier, int n, int last_level, uint32_t* counts, unsigned long long* leaves) {
    // one THREAD per parent (ThreadPos, tak_device.cuh): result + move count from the record's tail
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long add = 0;
    if (w < n) {
        ThreadPos<N> g;
        g.load(frontier + size_t(w) * StateLayout<N>::S);
        const uint8_t r = g.result();
        uint32_t c = 0;
        if (r != RES_ONGOING) {
            add = 1;
        } else {
            const uint32_t total = g.count_moves();
            if (last_level) add = total; else c = total;
        }
        if (counts) counts[w] = c;
    }
    // block reduction of `add`
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(FULL, add, o);
    __shared__ unsigned long long s_add[8];
    if ((threadIdx.x & 31) == 0) s_add[threadIdx.x >> 5] = add;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += s_add[i];
        if (t) atomicAdd(leaves, t);
    }
}

// children of parent w are written to out[offsets[w] ...] in move-generation order; the generated moves go
// through `moves_out` (also kept: it is the per-node move record of the expansion).
template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_perft_expand(const uint8_t* frontier, int n, const uint32_t* counts, const uint64_t* offsets,
                   uint64_t base_off, uint8_t* out, uint16_t* moves_out) {
    const int w = warp_global_id();
    if (w >= n) return;
    const uint32_t cnt = counts[w];
    if (cnt == 0) return;
    constexpr int S = StateLayout<N>::S;
    WarpGame<N> g;
    g.load(frontier + size_t(w) * S);
    const uint64_t base = offsets[w] - base_off;
    uint16_t* mv = moves_out + base;
    g.generate([&](int k, uint16_t m) { mv[k] = m; });
    __syncwarp();
    for (uint32_t k = 0; k < cnt; ++k) {
        WarpGame<N> child = g;
        child.template play<false>(mv[k]);
        child.store(out + (base + k) * S);
    }
}

}  // namespace tb
