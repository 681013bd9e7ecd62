// alpha_tak::Example on the device: the replay records self-play produces become training tensors with the reference's
// 8-fold symmetry augmentation (reference: alpha-tak/src/example.rs:35-78 `to_tensors`, tak/src/symm.rs:5-97).
// SURVEY.md section 8(f) row N2 -- the data format on the consumer side of the self-play path.  Byte/integer work,
// HBM-bound: per example it reads one packed state + <= 256 (move, visits) pairs and writes 8 x (C*N*N + P) floats.
#include <vector>

#include "engine.hpp"
#include "game_kernels.cuh"
#include "net_kernels.cuh"
#include "symmetry.hpp"

namespace tb {

struct ExamplesState {
    DevBuf states, sym_states, moves, visits, counts, results, move_table, inputs, pi, z;
};

// Symmetry for Game<N> (symm.rs:57-97): boards[k][sym_k(square)] = board[square]; scalars are copied.
// One warp per (example, symmetry); every lane gathers the source square of its own output squares.
template <int N>
__global__ void __launch_bounds__(GAME_THREADS) k_symmetries(const uint8_t* states, int n_examples, uint8_t* out) {
    const int w = warp_global_id();
    if (w >= n_examples * 8) return;
    using L = StateLayout<N>;
    using Col = typename L::Col;
    const int e = w >> 3, k = w & 7;
    const uint8_t* rec = states + size_t(e) * L::S;
    const Col* cols = reinterpret_cast<const Col*>(rec);
    const uint8_t* hts = rec + L::HTS_OFF;
    WarpGame<N> g;
    g.load(rec);  // scalars (and the identity board)
    const uint64_t walls = g.walls, caps = g.caps;
    const int l = threadIdx.x & 31;
    bool w0 = false, w1 = false, k0 = false, k1 = false;
    g.c0 = 0; g.c1 = 0; g.h0 = 0; g.h1 = 0;
#pragma unroll
    for (int half = 0; half < (WarpGame<N>::TWO ? 2 : 1); ++half) {
        const int o = l + 32 * half;          // output square, move-generation order o = col*N + row
        if (o >= N * N) continue;
        int sc, sr;
        sym_square_inv(N, k, o / N, o % N, &sc, &sr);
        const int src = sc * N + sr;
        const Col c = cols[src];
        const int h = hts[src];
        const bool wl = (walls >> src) & 1, cp = (caps >> src) & 1;
        if (half) { g.c1 = c; g.h1 = h; w1 = wl; k1 = cp; } else { g.c0 = c; g.h0 = h; w0 = wl; k0 = cp; }
    }
    g.walls = g.bb(w0, w1);
    g.caps = g.bb(k0, k1);
    g.store(out + size_t(w) * L::S);
}

// pi[8e+k][move_index(sym_k(move))] = visits / total  (example.rs:64-70); pi is zero-filled beforehand
__global__ void __launch_bounds__(256)
    k_pi_scatter(const uint16_t* moves, const uint32_t* visits, const int* counts, int n_examples, int n, int psz,
                 const uint16_t* move_table, float* pi, int* err) {
    const int e = blockIdx.x;
    if (e >= n_examples) return;
    const int cnt = counts[e];
    const uint16_t* mv = moves + size_t(e) * TAK_REPLAY_MAX_CHILDREN;
    const uint32_t* vis = visits + size_t(e) * TAK_REPLAY_MAX_CHILDREN;
    __shared__ uint32_t s_total;
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < cnt; ++i) t += vis[i];   // u32 sum, then `as f32` (example.rs:65)
        s_total = t;
    }
    __syncthreads();
    const float total = float(s_total);
    for (int i = threadIdx.x; i < cnt * 8; i += blockDim.x) {
        const int c = i >> 3, k = i & 7;
        const int idx = move_table[sym_move(n, k, mv[c])];
        if (idx == 0xFFFF) {
            atomicOr(err, 1);
            continue;
        }
        pi[(size_t(e) * 8 + k) * psz + idx] = __fdiv_rn(float(vis[c]), total);
    }
}

__global__ void k_fill_z(const float* results, int n_examples, float* z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_examples * 8) z[i] = results[i >> 3];
}

void examples_destroy(tak_engine* e) {
    if (!e->examples) return;
    ExamplesState& x = *e->examples;
    for (DevBuf* b : {&x.states, &x.sym_states, &x.moves, &x.visits, &x.counts, &x.results, &x.move_table, &x.inputs,
                      &x.pi, &x.z})
        b->release();
    delete e->examples;
    e->examples = nullptr;
}

}  // namespace tb

using namespace tb;

extern "C" {

int32_t tak_symmetry_move(int32_t n, uint16_t move, int32_t k, uint16_t* out) {
    TB_CHECK(n >= 3 && n <= 8 && k >= 0 && k < 8 && out, TAK_ERR_BAD_ARG, "tak_symmetry_move: bad argument");
    TB_CHECK((move & 63) < n * n, TAK_ERR_BAD_ARG, "tak_symmetry_move: square out of range");
    *out = sym_move(n, k, move);
    return TAK_OK;
}

int32_t tak_symmetry_state(const tak_state_t* s, int32_t k, tak_state_t* out) {
    TB_CHECK(s && out && k >= 0 && k < 8 && s->n >= 3 && s->n <= 8, TAK_ERR_BAD_ARG, "tak_symmetry_state: bad argument");
    const int n = s->n;
    tak_state_t r = *s;
    for (int i = 0; i < 64; ++i) { r.height[i] = 0; r.top[i] = 0; r.stack_lo[i] = 0; r.stack_hi[i] = 0; }
    for (int row = 0; row < n; ++row)
        for (int col = 0; col < n; ++col) {
            int c, rr;
            sym_square(n, k, col, row, &c, &rr);
            const int src = row * n + col, dst = rr * n + c;
            r.height[dst] = s->height[src];
            r.top[dst] = s->top[src];
            r.stack_lo[dst] = s->stack_lo[src];
            r.stack_hi[dst] = s->stack_hi[src];
        }
    *out = r;
    return TAK_OK;
}

int32_t examples_to_tensors(tak_engine_t* e, const tak_replay_record_t* recs, int32_t count, float* inputs, float* pi,
                            float* z, int32_t on_device) {
    TB_CHECK(e && recs && count >= 0 && (count == 0 || (inputs && pi && z)), TAK_ERR_BAD_ARG,
             "examples_to_tensors: bad argument");
    if (count == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    if (!e->examples) {
        e->examples = new ExamplesState();
        const std::vector<uint16_t>& mt = host_move_index_table(e->n);
        TB_CUDA(e->examples->move_table.ensure(mt.size() * 2));
        TB_CUDA(cudaMemcpyAsync(e->examples->move_table.p, mt.data(), mt.size() * 2, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaStreamSynchronize(e->stream));
    }
    ExamplesState& x = *e->examples;
    const int S = e->state_bytes, C = input_channels_c(e->n), nsq = e->nsq, psz = host_policy_size(e->n);
    // stage the records: packed states, child lists, results
    std::vector<uint8_t> packed(size_t(count) * S);
    std::vector<uint16_t> mv(size_t(count) * TAK_REPLAY_MAX_CHILDREN);
    std::vector<uint32_t> vis(size_t(count) * TAK_REPLAY_MAX_CHILDREN);
    std::vector<int> cnt(count);
    std::vector<float> res(count);
    for (int i = 0; i < count; ++i) {
        const tak_replay_record_t& r = recs[i];
        TB_CHECK(r.state.n == e->n, TAK_ERR_BAD_ARG, "example %d has board size %d, engine has %d", i, r.state.n, e->n);
        TB_CHECK(r.n_children >= 0 && r.n_children <= TAK_REPLAY_MAX_CHILDREN, TAK_ERR_BAD_ARG,
                 "example %d: bad child count %d", i, r.n_children);
        pack_state(e->n, r.state, packed.data() + size_t(i) * S);
        std::memcpy(&mv[size_t(i) * TAK_REPLAY_MAX_CHILDREN], r.moves, sizeof(r.moves));
        std::memcpy(&vis[size_t(i) * TAK_REPLAY_MAX_CHILDREN], r.visits, sizeof(r.visits));
        cnt[i] = r.n_children;
        res[i] = r.result;
    }
    const size_t rows = size_t(count) * 8;
    TB_CUDA(x.states.ensure(packed.size()));
    TB_CUDA(x.sym_states.ensure(rows * S));
    TB_CUDA(x.moves.ensure(mv.size() * 2));
    TB_CUDA(x.visits.ensure(vis.size() * 4));
    TB_CUDA(x.counts.ensure(size_t(count) * 4 + 16));
    TB_CUDA(x.results.ensure(size_t(count) * 4));
    float *d_in = inputs, *d_pi = pi, *d_z = z;
    if (!on_device) {
        TB_CUDA(x.inputs.ensure(rows * C * nsq * 4));
        TB_CUDA(x.pi.ensure(rows * psz * 4));
        TB_CUDA(x.z.ensure(rows * 4));
        d_in = x.inputs.as<float>(); d_pi = x.pi.as<float>(); d_z = x.z.as<float>();
    }
    int* d_err = x.counts.as<int>() + count;
    TB_CUDA(cudaMemcpyAsync(x.states.p, packed.data(), packed.size(), cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(x.moves.p, mv.data(), mv.size() * 2, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(x.visits.p, vis.data(), vis.size() * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(x.counts.p, cnt.data(), size_t(count) * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(x.results.p, res.data(), size_t(count) * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemsetAsync(d_err, 0, 4, e->stream));
    TB_CUDA(cudaMemsetAsync(d_pi, 0, rows * psz * 4, e->stream));
    const int wb = (int(rows) + GAME_WARPS_PER_BLOCK - 1) / GAME_WARPS_PER_BLOCK;
    TB_DISPATCH_N(e->n, (k_symmetries<N_><<<wb, GAME_THREADS, 0, e->stream>>>(x.states.as<uint8_t>(), count,
                                                                              x.sym_states.as<uint8_t>())));
    TB_DISPATCH_N(e->n, (k_repr_f32<N_><<<(int(rows) + 7) / 8, 256, 0, e->stream>>>(x.sym_states.as<uint8_t>(),
                                                                                    int(rows), d_in)));
    k_pi_scatter<<<count, 256, 0, e->stream>>>(x.moves.as<uint16_t>(), x.visits.as<uint32_t>(), x.counts.as<int>(), count,
                                               e->n, psz, x.move_table.as<uint16_t>(), d_pi, d_err);
    k_fill_z<<<(int(rows) + 255) / 256, 256, 0, e->stream>>>(x.results.as<float>(), count, d_z);
    e->launches += 4;
    TB_CUDA(cudaGetLastError());
    int err = 0;
    TB_CUDA(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, e->stream));
    if (!on_device) {
        TB_CUDA(cudaMemcpyAsync(inputs, d_in, rows * C * nsq * 4, cudaMemcpyDeviceToHost, e->stream));
        TB_CUDA(cudaMemcpyAsync(pi, d_pi, rows * psz * 4, cudaMemcpyDeviceToHost, e->stream));
        TB_CUDA(cudaMemcpyAsync(z, d_z, rows * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    TB_CUDA(cudaStreamSynchronize(e->stream));
    TB_CHECK(err == 0, TAK_ERR_INVALID_MOVE, "examples_to_tensors: a policy move has no policy index");
    return TAK_OK;
}

}  // extern "C"
