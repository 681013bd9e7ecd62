// Network-side kernels around the tcgen05 conv tower: input encoding, heads, softmax statistics.
//   k_encode      alpha_tak::repr::game_repr (repr/game.rs:19-51, board.rs:12-54, reserves.rs:4-28) written straight
//                 into the first conv's bf16 slot-plane layout (no host tensor ops, no H2D of planes)
//   k_repr_f32    the same planes as fp32 [C][N][N] (the reference's exact tensor; parity/debug surface)
//   k_policy_stats_* / k_value / k_policy_full   heads of net6.rs:98-109 / net5.rs:106-111 (Net5's policy FC: fc_tc.cuh)
#pragma once
#include <cuda_bf16.h>

#include "conv_tc3.cuh"
#include "tak_device.cuh"

namespace tb {

__host__ __device__ constexpr int board_channels(int n) { return (n + 8) * 2; }
__host__ __device__ constexpr int input_channels_c(int n) {
    return board_channels(n) + 2 + 2 * tak_stones(n) + 2 * tak_caps(n);
}

// value of input plane `ch` on a square, given the square's column/height/kind and the game scalars
template <int N>
struct ReprCtx {
    int to_move, ws, wc, bs, bc;
    float fcd;  // (flat_diff - half_komi/2) / N^2 computed in double, stored as f32 (game.rs:40-43)
};

template <int N, class Col>
__device__ __forceinline__ float repr_plane(const ReprCtx<N>& cx, int ch, Col col, int h, int kind) {
    constexpr int BC = board_channels(N);
    constexpr int ST = tak_stones(N), CP = tak_caps(N);
    if (ch < 6) {
        if (h == 0) return 0.f;
        const int mine = (int((col >> (h - 1)) & 1) == cx.to_move) ? 0 : 1;
        return ch == kind * 2 + mine ? 1.f : 0.f;
    }
    if (ch < BC) {
        const int i = (ch - 6) >> 1;          // depth below the top: i = 0 is the piece under the top
        const int k = h - 2 - i;              // index from the bottom
        if (k < 0) return 0.f;
        const int mine = (int((col >> k) & 1) == cx.to_move) ? 0 : 1;
        return ((ch - 6) & 1) == mine ? 1.f : 0.f;
    }
    int c = ch - BC;
    const int my_st = cx.to_move == 0 ? cx.ws : cx.bs, en_st = cx.to_move == 0 ? cx.bs : cx.ws;
    const int my_cp = cx.to_move == 0 ? cx.wc : cx.bc, en_cp = cx.to_move == 0 ? cx.bc : cx.wc;
    if (c < ST) return (my_st > 0 && c == my_st - 1) ? 1.f : 0.f;
    c -= ST;
    if (c < ST) return (en_st > 0 && c == en_st - 1) ? 1.f : 0.f;
    c -= ST;
    if (c < CP) return (my_cp > 0 && c == my_cp - 1) ? 1.f : 0.f;
    c -= CP;
    if (c < CP) return (en_cp > 0 && c == en_cp - 1) ? 1.f : 0.f;
    c -= CP;
    if (c == 0) return cx.to_move == 0 ? 1.f : 0.f;
    if (c == 1) return cx.fcd;
    return 0.f;  // channel padding up to 128
}

// game_repr of the game a warp holds in registers -> strip planes of evaluation slot w.  Pad columns and the tile
// remainder are zero already: the input planes (NetState::act_in) are zero-filled at allocation and only real squares
// are ever written to them.
template <int N, bool PF = INFER_PF>
__device__ __forceinline__ void encode_board(const WarpGame<N>& g, int w, __nv_bfloat16* planes, int S) {
    ReprCtx<N> cx;
    cx.to_move = g.to_move; cx.ws = g.ws; cx.wc = g.wc; cx.bs = g.bs; cx.bc = g.bc;
    const int fcd = int(int8_t(g.flat_diff() - g.half_komi / 2));
    cx.fcd = float(double(fcd) / double(N * N));
    const int l = threadIdx.x & 31;
#pragma unroll
    for (int half = 0; half < (WarpGame<N>::TWO ? 2 : 1); ++half) {
        const int o = l + 32 * half;
        if (o >= N * N) continue;
        const auto col = half ? g.c1 : g.c0;
        const int h = half ? g.h1 : g.h0;
        const int kind = ((g.walls >> o) & 1) ? 1 : ((g.caps >> o) & 1) ? 2 : 0;
        const int x = o / N, y = o % N;
        const size_t slot = SlotMap<N, PF>::slot(w, y, x);
        // unrolled: with the channel a compile-time constant repr_plane folds to a compare or a bit test per value (rolled,
        // its chain of range checks was a fifth of the search step's instructions, profiles/r02_mcts_step_ncu_full_g4144.txt)
#pragma unroll
        for (int chunk = 0; chunk < 16; ++chunk) {
            uint4 v;
            __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = repr_plane<N>(cx, chunk * 8 + 2 * j, col, h, kind);
                const float b = repr_plane<N>(cx, chunk * 8 + 2 * j + 1, col, h, kind);
                vb[j] = __floats2bfloat162_rn(a, b);
            }
            *reinterpret_cast<uint4*>(planes + (size_t(chunk) * S + slot) * 8) = v;
        }
    }
}

// one warp per board; `states` = packed records of the boards to encode (index list optional).
// (launch bounds: at most 40 registers per thread.  The cap dates from the attempt to run this kernel beside the other
// engine replica's conv tower; nothing fits beside a tower CTA -- profiles/r02_step_overlap.md -- but the cap costs nothing)
template <int N>
__global__ void __launch_bounds__(256, 6)
    k_encode(const uint8_t* states, const int* index, int n_boards, __nv_bfloat16* planes, int S,
             const int* n_dev = nullptr) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_boards || (n_dev && w >= *n_dev)) return;   // n_dev: the live count, on the device (launch sized for n_boards)
    WarpGame<N> g;
    g.load(states + size_t(index ? index[w] : w) * StateLayout<N>::S);
    encode_board<N>(g, w, planes, S);
}

template <int N>
__global__ void __launch_bounds__(256) k_repr_f32(const uint8_t* states, int n_boards, float* out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_boards) return;
    constexpr int C = input_channels_c(N);
    WarpGame<N> g;
    g.load(states + size_t(w) * StateLayout<N>::S);
    ReprCtx<N> cx;
    cx.to_move = g.to_move; cx.ws = g.ws; cx.wc = g.wc; cx.bs = g.bs; cx.bc = g.bc;
    const int fcd = int(int8_t(g.flat_diff() - g.half_komi / 2));
    cx.fcd = float(double(fcd) / double(N * N));
    const int l = threadIdx.x & 31;
#pragma unroll
    for (int half = 0; half < (WarpGame<N>::TWO ? 2 : 1); ++half) {
        const int o = l + 32 * half;
        if (o >= N * N) continue;
        const auto col = half ? g.c1 : g.c0;
        const int h = half ? g.h1 : g.h0;
        const int kind = ((g.walls >> o) & 1) ? 1 : ((g.caps >> o) & 1) ? 2 : 0;
        const int x = o / N, y = o % N;
        for (int ch = 0; ch < C; ++ch)
            out[(size_t(w) * C + ch) * (N * N) + y * N + x] = repr_plane<N>(cx, ch, col, h, kind);
    }
}

// deterministic block reductions (fixed shuffle tree + fixed warp order)
__device__ __forceinline__ float block_reduce_max(float v, float* s_tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) s_tmp[w] = v;
    __syncthreads();
    float r = s_tmp[0];
    for (int i = 1; i < nw; ++i) r = fmaxf(r, s_tmp[i]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* s_tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) s_tmp[w] = v;
    __syncthreads();
    float r = s_tmp[0];
    for (int i = 1; i < nw; ++i) r += s_tmp[i];
    __syncthreads();
    return r;
}

// Net6-style head (policy conv): softmax statistics over ALL channels x squares of one board (net6.rs:100-103).
// The conv epilogue already reduced every slot's channels to {max, sum exp(l - max)} per 32-channel lane quarter
// (partials[group*4 + quarter][S]); one warp per board merges the N*N x parts partials: stats[b] = {max, sum of exp(l - max)}.
template <int N, bool PF>
__device__ __forceinline__ float2 warp_policy_stats(const float2* partials, int S, int groups, int b) {
    constexpr int NSQ = N * N;
    const int l = threadIdx.x & 31;
    float mx = -INFINITY;
    for (int i = l; i < NSQ * groups; i += 32) {
        const int g = i / NSQ, sq = i % NSQ;
        mx = fmaxf(mx, partials[size_t(g) * S + SlotMap<N, PF>::slot(b, sq / N, sq % N)].x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    float sum = 0.f;
    for (int i = l; i < NSQ * groups; i += 32) {
        const int g = i / NSQ, sq = i % NSQ;
        const float2 pr = partials[size_t(g) * S + SlotMap<N, PF>::slot(b, sq / N, sq % N)];
        sum += pr.y * expf(pr.x - mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    return make_float2(mx, sum);
}
// (PF: which strip layout the partials are in -- the training path uses the padded one, the default)
template <int N, bool PF = false>
__global__ void __launch_bounds__(256)
    k_policy_stats_conv(const float2* partials, int S, int groups, int n_boards, float2* stats,
                        const int* n_dev = nullptr) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= n_boards || (n_dev && b >= *n_dev)) return;
    const float2 st = warp_policy_stats<N, PF>(partials, S, groups, b);
    if ((threadIdx.x & 31) == 0) stats[b] = st;
}
// full softmax vector [b][n_ch * N*N] (index = ch*N*N + row*N + col) from the logits and the board statistics: the
// host-facing Network::policy_eval surface (network.rs:34); the search gathers only the legal moves' priors instead.
template <int N>
__global__ void __launch_bounds__(256)
    k_policy_full_conv(const float* logits, int S, int n_ch, const float2* stats, float* policy_out, int raw) {
    const int b = blockIdx.x;
    constexpr int NSQ = N * N;
    const float2 st = stats[b];
    float* dst = policy_out + size_t(b) * n_ch * NSQ;
    for (int i = threadIdx.x; i < n_ch * NSQ; i += blockDim.x) {
        const int ch = i / NSQ, sq = i % NSQ;
        const float lg = logits[size_t(ch) * S + SlotMap<N, INFER_PF>::slot(b, sq / N, sq % N)];
        dst[i] = raw ? lg : __fdiv_rn(expf(__fsub_rn(lg, st.x)), st.y);   // raw: the pre-softmax logits (parity surface)
    }
}

// softmax statistics / full policy over a dense logits row [b][n_out]
static __global__ void __launch_bounds__(256)
    k_policy_stats_dense(const float* logits, int n_out, float2* stats, float* policy_out, int raw,
                         const int* n_dev = nullptr) {
    __shared__ float s_tmp[8];
    const int b = blockIdx.x;
    if (n_dev && b >= *n_dev) return;
    const float* row = logits + size_t(b) * n_out;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < n_out; j += blockDim.x) mx = fmaxf(mx, row[j]);
    mx = block_reduce_max(mx, s_tmp);
    float sum = 0.f;
    for (int j = threadIdx.x; j < n_out; j += blockDim.x) sum += expf(row[j] - mx);
    sum = block_reduce_sum(sum, s_tmp);
    if (threadIdx.x == 0) stats[b] = make_float2(mx, sum);
    if (policy_out)
        for (int j = threadIdx.x; j < n_out; j += blockDim.x)
            policy_out[size_t(b) * n_out + j] = raw ? row[j] : __fdiv_rn(expf(__fsub_rn(row[j], mx)), sum);
}

// value head (net6.rs:104-107 / net5.rs:109): tanh(fc(flatten_NCHW(s)))  -- one warp per board
template <int N, bool PF>
__device__ __forceinline__ float warp_value(const __nv_bfloat16* act, int S, const float* wv /*[128*NSQ]*/, float bv, int w) {
    constexpr int NSQ = N * N;
    const int l = threadIdx.x & 31;
    float acc = 0.f;
    for (int pos = l; pos < NSQ; pos += 32) {
        const int y = pos / N, x = pos % N;
        const size_t slot = SlotMap<N, PF>::slot(w, y, x);
        for (int chunk = 0; chunk < 16; ++chunk) {
            const uint4 v = *reinterpret_cast<const uint4*>(act + (size_t(chunk) * S + slot) * 8);
            const __nv_bfloat162* vb = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(vb[j]);
                acc += f.x * wv[(chunk * 8 + 2 * j) * NSQ + pos];
                acc += f.y * wv[(chunk * 8 + 2 * j + 1) * NSQ + pos];
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    return tanhf(acc + bv);
}
template <int N>
__global__ void __launch_bounds__(256)
    k_value(const __nv_bfloat16* act, int S, const float* wv /*[128*NSQ]*/, float bv, int n_boards, float* out,
            const int* n_dev = nullptr) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_boards || (n_dev && w >= *n_dev)) return;
    const float val = warp_value<N, INFER_PF>(act, S, wv, bv, w);
    if ((threadIdx.x & 31) == 0) out[w] = val;
}

}  // namespace tb
