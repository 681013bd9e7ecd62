// Batched Tak kernels over packed states: one warp per game.
//   k_reset / k_moves / k_play / k_result         -> tak_games_reset / tak_possible_moves / tak_play / tak_result
//   k_perft_count / k_perft_moves / k_perft_apply -> tak_perft (breadth-first, perf_count rule of perft.rs:3-18)
//   k_playout                                     -> tak_playouts (random playouts to termination, device-resident)
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

// ---- perft expansion, output-centric ----------------------------------------------------------------------------
// Children of a frontier are contiguous in the output (parent order, then move-generation order), so the expansion is
// organised by OUTPUT range: a block owns PerftCfg::CH consecutive children, whatever parents they come from.
//   k_perft_moves  one warp per parent: possible_moves -> moves_out[offsets[w] ...]; it also records, for every block
//                  boundary that falls inside its children, which parent the block starts in (block_parent).
//   k_perft_apply  per block: (1) every child finds its parent (binary search between the block's first and last
//                  parent); (2) the parents' heights + tail words are copied into per-child images in shared memory;
//                  (3) ONE THREAD per child patches its image in place and lists the (at most N) new stack columns
//                  (ChildPatch: 176 bytes per child on 6x6 instead of a 384-byte record, which is what keeps ~1 300
//                  children in flight per SM), then classifies / counts the new position from the tail it just built
//                  (ThreadPos: Game::result + possible_moves().len()); (4) the block streams the children out with
//                  16-byte stores: stack-column words come from the parent (L1) unless the patch names one of their
//                  squares, heights + tail from the images.  The child counts feed the next level's scan; at the last
//                  level they ARE perf_count's answer (perft.rs:6-7): no position is re-read by a separate count kernel.
// HBM traffic per child: S written + S/b read (the parent, once per block that touches it) + 2 B move written + read
// + 4 B count -- SURVEY.md 8(d)'s S + S/b + 2 within 2 %.
// Variant V (TAK_PERFT_VARIANT, measured on B200, 6x6 perft(5), G states/s of the depth-3 -> 4 expansion):
//   0: 256 threads, conflict-free (odd) image pitch 9.6 | 1: 256, dense pitch 10.1 | 3: 128, dense 11.1 (default) |
//   5: 64, dense.  Smaller blocks and the 160-byte pitch put more blocks on an SM (shared memory is the limiter), which
//   hides the L1 / shared-memory latency of the one-thread-per-child phase better than avoiding its 2-way bank conflicts.
template <int N, int V = 3>
struct PerftCfg {
    static constexpr int S = StateLayout<N>::S;
    static constexpr int W = S / 16;                  // 16-byte words per record
    static constexpr int THREADS = 256 >> (V >> 1);
    static constexpr int CH = THREADS;                // children per block: one patch thread each
    static constexpr int STRIDE = (V & 1) ? (ChildPatch<N>::RAW + 15) / 16 * 16 : ChildPatch<N>::STRIDE;
    static constexpr int SMEM = CH * STRIDE + CH * 4;
    static constexpr bool STREAM = false;             // st.global.cs for the children: no effect (measured)
};

template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_perft_moves(const uint8_t* frontier, int n, const uint32_t* counts, const uint64_t* offsets, uint64_t base_off,
                  uint16_t* moves_out, int* block_parent, int CH /* children per k_perft_apply block */) {
    const int w = warp_global_id();
    if (w >= n) return;
    const uint32_t cnt = counts[w];
    if (cnt == 0) return;
    WarpGame<N> g;
    g.load(frontier + size_t(w) * StateLayout<N>::S);
    const uint64_t base = offsets[w] - base_off;
    uint16_t* mv = moves_out + base;
    g.generate([&](int k, uint16_t m) { mv[k] = m; });
    for (uint64_t b = (base + CH - 1) / CH + (threadIdx.x & 31); b * CH < base + cnt; b += 32) block_parent[b] = w;
}

template <int N, bool LAST, int V = 3>
__global__ void __launch_bounds__(PerftCfg<N, V>::THREADS)
    k_perft_apply(const uint8_t* frontier, int n_parents, const uint64_t* offsets, uint64_t base_off,
                  const int* block_parent, int n_blocks, int total_children, const uint16_t* moves, uint8_t* out,
                  uint32_t* child_counts, unsigned long long* leaves) {
    using P = PerftCfg<N, V>;
    using CP = ChildPatch<N>;
    extern __shared__ __align__(16) uint8_t s_patch[];
    int* s_parent = reinterpret_cast<int*>(s_patch + P::CH * P::STRIDE);
    const int t = threadIdx.x;
    const int c0 = blockIdx.x * P::CH;
    const int nc = min(P::CH, total_children - c0);
    const uint4* src = reinterpret_cast<const uint4*>(frontier);
    // (1) owner of child c = the LAST parent whose first child is <= c (parents without children share their offset
    // with the next parent), searched between the owners of this block's and the next block's first child
    if (t < nc) {
        const uint64_t c = uint64_t(c0 + t);
        int lo = block_parent[blockIdx.x];
        int hi = int(blockIdx.x) + 1 < n_blocks ? block_parent[blockIdx.x + 1] : n_parents - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (offsets[mid] - base_off <= c) lo = mid; else hi = mid - 1;
        }
        s_parent[t] = lo;
    }
    __syncthreads();
    // (2) the parents' heights + tail words -> per-child images
    {
        constexpr int DCH = P::THREADS / CP::TW, DWD = P::THREADS % CP::TW;
        int ch = t / CP::TW, wd = t - ch * CP::TW;
        for (int idx = t; idx < nc * CP::TW; idx += P::THREADS) {
            *reinterpret_cast<uint4*>(s_patch + ch * P::STRIDE + wd * 16) =
                __ldg(src + size_t(s_parent[ch]) * P::W + CP::HT0 + wd);
            ch += DCH;
            wd += DWD;
            if (wd >= CP::TW) { wd -= CP::TW; ++ch; }
        }
    }
    __syncthreads();
    // (3) one thread per child: patch, classify, count
    unsigned long long add = 0;
    if (t < nc) {
        ThreadPos<N> tp;
        CP::build(frontier + size_t(s_parent[t]) * P::S, moves[c0 + t], s_patch + t * P::STRIDE, tp);
        uint32_t cnt = 0;
        if (tp.result() != RES_ONGOING) {
            add = 1;                                   // a finished game counts 1 wherever it ends (perft.rs:4)
        } else {
            const uint32_t total = tp.count_moves();   // heights come from the patched image
            if (LAST) add = total; else cnt = total;   // depth 1: the number of legal moves (perft.rs:6-7)
        }
        if (!LAST) child_counts[c0 + t] = cnt;
    }
    __syncthreads();
    // (4) stream the children out: stack-column words (parent's unless patched), then heights + tail from the images
    uint4* dst = reinterpret_cast<uint4*>(out) + size_t(c0) * P::W;
    {
        constexpr int DCH = P::THREADS / CP::HT0, DWD = P::THREADS % CP::HT0;
        int ch = t / CP::HT0, wd = t - ch * CP::HT0;
        for (int idx = t; idx < nc * CP::HT0; idx += P::THREADS) {
            const uint4 v = CP::cols_word(src + size_t(s_parent[ch]) * P::W, s_patch + ch * P::STRIDE, wd);
            if (P::STREAM) __stcs(dst + ch * P::W + wd, v); else dst[ch * P::W + wd] = v;
            ch += DCH;
            wd += DWD;
            if (wd >= CP::HT0) { wd -= CP::HT0; ++ch; }
        }
    }
    {
        constexpr int DCH = P::THREADS / CP::TW, DWD = P::THREADS % CP::TW;
        int ch = t / CP::TW, wd = t - ch * CP::TW;
        for (int idx = t; idx < nc * CP::TW; idx += P::THREADS) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_patch + ch * P::STRIDE + wd * 16);
            if (P::STREAM) __stcs(dst + ch * P::W + CP::HT0 + wd, v); else dst[ch * P::W + CP::HT0 + wd] = v;
            ch += DCH;
            wd += DWD;
            if (wd >= CP::TW) { wd -= CP::TW; ++ch; }
        }
    }
    // block reduction of `add`
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(FULL, add, o);
    __shared__ unsigned long long s_add[P::THREADS / 32];
    if ((t & 31) == 0) s_add[t >> 5] = add;
    __syncthreads();
    if (t == 0) {
        unsigned long long sum = 0;
        for (int i = 0; i < P::THREADS / 32; ++i) sum += s_add[i];
        if (sum) atomicAdd(leaves, sum);
    }
}

// ---- random playouts (workload A) -----------------------------------------------------------
// One warp plays game w from its current position: while the game is ongoing and fewer than max_plies[w] (or max_ply)
// plies were added, move = possible_moves()[splitmix64(seed ^ splitmix64(game_id << 32 | ply)) % len] -- the state stays
// in registers for the whole playout; HBM sees one load and one store per game.
__device__ __forceinline__ uint64_t playout_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
template <int N>
__global__ void __launch_bounds__(GAME_THREADS)
    k_playout(uint8_t* states, int first, int n, uint64_t seed, int id_base, int max_plies, int ply_spread,
              int* out_plies, uint8_t* out_result, unsigned long long* totals) {
    const int w = warp_global_id();
    if (w >= n) return;
    constexpr int S = StateLayout<N>::S;
    const int l = threadIdx.x & 31;
    const uint64_t gid = uint64_t(uint32_t(id_base + first + w));
    WarpGame<N> g;
    uint8_t* rec = states + size_t(first + w) * S;
    g.load(rec);
    // per-game ply budget: max_plies + (hash % ply_spread) so that a batch can be cut at staggered depths
    int budget = max_plies;
    if (ply_spread > 0) budget += int(playout_mix(seed ^ playout_mix(gid ^ 0xC0FFEEull)) % uint64_t(ply_spread));
    int plies = 0;
    unsigned long long generated = 0;
    uint8_t r = g.result();
    while (r == RES_ONGOING && plies < budget) {
        int total = 0;
        const uint64_t key = playout_mix(seed ^ playout_mix((gid << 32) | uint64_t(uint32_t(g.ply))));
        const uint16_t mv = g.select_move([&](int len) { return int(key % uint64_t(len)); }, total);
        generated += uint64_t(total);
        g.template play<false>(mv);
        ++plies;
        r = g.result();
    }
    g.store(rec);
    if (l == 0) {
        if (out_plies) out_plies[w] = plies;
        if (out_result) out_result[w] = r;
        atomicAdd(totals + 0, (unsigned long long)plies);
        atomicAdd(totals + 1, generated);
    }
}

}  // namespace tb
