// alpha_tak::Network for the engine: Net5 / Net6 forward (policy_eval / forward_mcts) on tcgen05 tensor cores.
// Reference: alpha-tak/src/model/net6.rs:29-138, net5.rs:29-130, res_block.rs:13-23, network.rs:26-34.
// DummyNet (alpha-tak/src/search/tests.rs:6-35) is arch 0.
#include "net.hpp"

#include <cmath>
#include <cstdlib>

#include "conv_tc3.cuh"
#include "fc_tc.cuh"
#include "net_kernels.cuh"

namespace tb {

static constexpr float BN_EPS = 1e-5f;  // tch nn::BatchNormConfig default

int64_t net_blob_elems(const NetState& ns) {
    const int64_t cin = ns.c_in, nsq = int64_t(ns.n) * ns.n;
    int64_t total = 128 * cin * 9 + 128 + 4 * 128;
    total += int64_t(ns.blocks) * (2 * (128 * 128 * 9 + 128) + 8 * 128);
    if (ns.arch == 6) total += int64_t(ns.policy_ch) * 128 * 9 + ns.policy_ch;
    else total += int64_t(ns.policy_out) * 128 * nsq + ns.policy_out;
    total += 128 * nsq + 1;
    return total;
}

// fold BN(eval) into conv weight/bias and pack for conv_tc3 (K-slab streaming, consumption order):
//   [slab = c_in/16][ky][kx][kchunk = (c_in/8)%2][c_out 128][8 c_in]
static void pack_conv(const float* w /*[co][ci][3][3]*/, const float* b, const float* bn /*gamma,beta,mean,var or null*/,
                      int c_out, int c_in, int co_base, std::vector<__nv_bfloat16>& packed, std::vector<float>& bias) {
    packed.assign(C3_W_LAYER_ELEMS, __float2bfloat16(0.f));
    bias.assign(128, 0.f);
    for (int col = 0; col < 128; ++col) {
        const int co = co_base + col;
        if (co >= c_out) continue;
        double scale = 1.0, shift = 0.0;
        if (bn) {
            const double gamma = bn[co], beta = bn[c_out + co], mean = bn[2 * c_out + co], var = bn[3 * c_out + co];
            scale = gamma / std::sqrt(var + double(BN_EPS));
            shift = beta - mean * scale;
        }
        bias[col] = float(double(b[co]) * scale + shift);
        for (int ci = 0; ci < c_in; ++ci)
            for (int tap = 0; tap < 9; ++tap) {
                const float v = float(double(w[(size_t(co) * c_in + ci) * 9 + tap]) * scale);
                const int slab = ci >> 4, kc = (ci >> 3) & 1, j = ci & 7;
                packed[((((size_t(slab) * 9 + tap) * 2 + kc) * 128) + col) * 8 + j] = __float2bfloat16(v);
            }
    }
}

static int upload_layer(tak_engine* e, ConvLayer& L, const std::vector<__nv_bfloat16>& packed,
                        const std::vector<float>& bias) {
    TB_CUDA(L.w.ensure(packed.size() * 2));
    TB_CUDA(L.bias.ensure(bias.size() * 4));
    TB_CUDA(cudaMemcpyAsync(L.w.p, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaMemcpyAsync(L.bias.p, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int net_load_blob(tak_engine* e, const float* blob, int64_t elems) {
    NetState& ns = *e->net;
    TB_CHECK(ns.arch != 0, TAK_ERR_BAD_ARG, "the DummyNet has no weights");
    TB_CHECK(elems == net_blob_elems(ns), TAK_ERR_BAD_ARG, "weight blob has %lld elements, architecture needs %lld",
             (long long)elems, (long long)net_blob_elems(ns));
    const int nsq = ns.n * ns.n;
    const float* p = blob;
    std::vector<__nv_bfloat16> packed;
    std::vector<float> bias;
    // initial conv + BN (net6.rs:39-40)
    {
        const float* w = p; p += size_t(128) * ns.c_in * 9;
        const float* b = p; p += 128;
        const float* bn = p; p += 4 * 128;
        pack_conv(w, b, bn, 128, ns.c_in, 0, packed, bias);
        if (int r = upload_layer(e, ns.layers[0], packed, bias)) return r;
    }
    // residual blocks: conv1, conv2, bn1, bn2 in creation order (net6.rs:43-54)
    for (int blk = 0; blk < ns.blocks; ++blk) {
        const float* w1 = p; p += size_t(128) * 128 * 9;
        const float* b1 = p; p += 128;
        const float* w2 = p; p += size_t(128) * 128 * 9;
        const float* b2 = p; p += 128;
        const float* bn1 = p; p += 4 * 128;
        const float* bn2 = p; p += 4 * 128;
        pack_conv(w1, b1, bn1, 128, 128, 0, packed, bias);
        if (int r = upload_layer(e, ns.layers[1 + 2 * blk], packed, bias)) return r;
        pack_conv(w2, b2, bn2, 128, 128, 0, packed, bias);
        if (int r = upload_layer(e, ns.layers[2 + 2 * blk], packed, bias)) return r;
    }
    if (ns.arch == 6) {
        const float* w = p; p += size_t(ns.policy_ch) * 128 * 9;
        const float* b = p; p += ns.policy_ch;
        for (int grp = 0; grp < ns.policy_groups; ++grp) {
            pack_conv(w, b, nullptr, ns.policy_ch, 128, grp * 128, packed, bias);
            if (int r = upload_layer(e, ns.policy_layers[grp], packed, bias)) return r;
        }
    } else {
        // FC policy: W[j][c*NSQ + pos] -> the tcgen05 operand image Wp[jt][pos][slab][kchunk][128 outputs][8 channels]
        const int K = 128 * nsq, J = ns.policy_out;
        const float* w = p; p += size_t(J) * K;
        const float* b = p; p += J;
        const int j_tiles = (J + 127) / 128;
        std::vector<__nv_bfloat16> wt(size_t(j_tiles) * nsq * 8 * 2 * 128 * 8, __float2bfloat16(0.f));
        for (int j = 0; j < J; ++j)
            for (int c = 0; c < 128; ++c)
                for (int pos = 0; pos < nsq; ++pos) {
                    const int jt = j >> 7, col = j & 127, slab = c >> 4, kc = (c >> 3) & 1, i = c & 7;
                    wt[((((size_t(jt) * nsq + pos) * 8 + slab) * 2 + kc) * 128 + col) * 8 + i] =
                        __float2bfloat16(w[size_t(j) * K + size_t(c) * nsq + pos]);
                }
        TB_CUDA(ns.fc_policy_w.ensure(wt.size() * 2));
        TB_CUDA(ns.fc_policy_b.ensure(size_t(J) * 4));
        TB_CUDA(cudaMemcpyAsync(ns.fc_policy_w.p, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaMemcpyAsync(ns.fc_policy_b.p, b, size_t(J) * 4, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaStreamSynchronize(e->stream));
    }
    {
        const float* w = p; p += size_t(128) * nsq;
        ns.value_bias = *p; p += 1;
        TB_CUDA(ns.value_w.ensure(size_t(128) * nsq * 4));
        TB_CUDA(cudaMemcpyAsync(ns.value_w.p, w, size_t(128) * nsq * 4, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaStreamSynchronize(e->stream));
    }
    TB_CHECK(p - blob == elems, TAK_ERR_BAD_ARG, "internal: blob walk mismatch");
    if (ns.blob_host.data() != blob) ns.blob_host.assign(blob, blob + elems);
    ns.loaded = true;
    return TAK_OK;
}

// tiles of `boards` boards, rounded up to the cluster granularity of the conv kernel (the extra tiles hold no boards:
// their activations stay zero)
static int tiles_for(int n, int boards) {
    const int t = n == 5 ? SlotMap<5, INFER_PF>::tiles(boards) : SlotMap<6, INFER_PF>::tiles(boards);
    return (t + C3_TILE_ALIGN - 1) / C3_TILE_ALIGN * C3_TILE_ALIGN;
}

int net_ensure_capacity(tak_engine* e, int boards) {
    NetState& ns = *e->net;
    if (boards <= ns.cap_boards) return TAK_OK;
    const int tiles = tiles_for(ns.n, boards);
    const int S = tiles * C3_TILE_M;
    for (DevBuf* b : {&ns.act[0], &ns.act[1], &ns.act[2], &ns.act_in}) {
        TB_CUDA(b->ensure(size_t(S) * 256));
        TB_CUDA(cudaMemsetAsync(b->p, 0, size_t(S) * 256, e->stream));
    }
    if (ns.arch == 6) {
        TB_CUDA(ns.logits.ensure(size_t(ns.policy_groups) * 128 * S * 4));
        TB_CUDA(ns.partials.ensure(size_t(ns.policy_groups) * 4 * S * 8));
    } else if (ns.arch == 5) {
        TB_CUDA(ns.logits.ensure(size_t(boards) * ns.policy_out * 4));
    }
    TB_CUDA(ns.stats.ensure(size_t(boards) * 8));
    TB_CUDA(ns.values.ensure(size_t(boards) * 4));
    ns.cap_boards = boards;
    ns.cap_S = S;
    return TAK_OK;
}

// The whole conv tower (initial conv, residual blocks, Net6 policy conv groups) as ONE launch of conv_tc3.cuh over the
// input planes in NetState::act_in.  `boards` sizes the launch; with d_count the kernel reads the actual number of
// boards from device memory (the fused search loop counts its leaves on the device).
template <int N>
static int launch_tower(tak_engine* e, int boards, const int* d_count) {
    NetState& ns = *e->net;
    if (int r = net_ensure_capacity(e, boards)) return r;
    const int tiles = tiles_for(N, boards);
    const int S = ns.cap_S;  // plane stride is fixed by the allocation
    __nv_bfloat16* x_in = ns.act_in.as<__nv_bfloat16>();
    __nv_bfloat16* x = ns.act[0].as<__nv_bfloat16>();
    __nv_bfloat16* t = ns.act[1].as<__nv_bfloat16>();
    __nv_bfloat16* y = ns.act[2].as<__nv_bfloat16>();
    ConvParams p{};
    conv_params_set_layout(p, N, INFER_PF);
    p.S = S; p.tile_begin = 0; p.tile_end = tiles; p.n_boards = boards; p.n_boards_dev = d_count;
    auto conv = [&](const ConvLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* res, __nv_bfloat16* out,
                    int mode, int slabs, int grp, int ch_valid, int discard = 0) {
        ConvLayerDesc& d = p.layers[p.n_layers++];
        static const bool no_discard = std::getenv("TAK_NO_L2_DISCARD") != nullptr;   // A/B switch for measurements
        d.discard = no_discard ? 0 : discard;
        d.in = in; d.res = res; d.out = out; d.out_f32 = ns.logits.as<float>(); d.partials = ns.partials.as<float2>();
        d.w = L.w.as<__nv_bfloat16>(); d.bias = L.bias.as<float>();
        d.slabs = slabs; d.mode = mode; d.out_ch_offset = grp * 128; d.out_ch_valid = ch_valid; d.group = grp;
    };
    // initial conv + BN + ReLU (net6.rs:72-76); only ceil(c_in/16) K-slabs carry input planes
    conv(ns.layers[0], x_in, nullptr, x, CONV_RELU, (ns.c_in + 15) / 16, 0, 128);
    // residual tower (res_block.rs:14-22)
    for (int blk = 0; blk < ns.blocks; ++blk) {
        conv(ns.layers[1 + 2 * blk], x, nullptr, t, CONV_RELU, C3_MAX_SLABS, 0, 128);
        conv(ns.layers[2 + 2 * blk], t, x, y, CONV_RES_RELU, C3_MAX_SLABS, 0, 128, /*t and the block input are dead*/ 3);
        std::swap(x, y);
    }
    ns.trunk_out = x;
    if (ns.arch == 6)  // policy conv (net6.rs:99-100), 128 output channels per group
        for (int grp = 0; grp < ns.policy_groups; ++grp)
            conv(ns.policy_layers[grp], x, nullptr, nullptr, CONV_LOGITS_F32, C3_MAX_SLABS, grp,
                 std::min(128, ns.policy_ch - grp * 128), std::getenv("TAK_NO_STCS") ? 0 : 4);
    NetProfile* prof = ns.profile;
    if (prof) TB_CUDA(cudaEventRecord(prof->ev[2 * prof->n], e->stream));
    TB_CUDA((conv3x3_tc3_launch<false, INFER_PF>(p, e->num_sms, e->stream)));
    e->launches++;
    if (prof) {
        TB_CUDA(cudaEventRecord(prof->ev[2 * prof->n + 1], e->stream));
        prof->n++;
    }
    if (ns.arch == 5) {
        // policy FC (net5.rs:108) on the tensor cores: repack the trunk output, then one GEMM launch (fc_tc.cuh)
        const int b_pad = (boards + FC_NT - 1) / FC_NT * FC_NT;
        TB_CUDA(ns.fc_x.ensure(size_t(N * N) * 16 * b_pad * 16));
        const size_t items = size_t(N * N) * 16 * b_pad;
        k_fc_repack<N, INFER_PF><<<unsigned((items + 255) / 256), 256, 0, e->stream>>>(x, S, boards, b_pad,
                                                                            ns.fc_x.as<__nv_bfloat16>());
        FcParams fp{};
        fp.wp = ns.fc_policy_w.as<__nv_bfloat16>(); fp.x = ns.fc_x.as<__nv_bfloat16>();
        fp.bias = ns.fc_policy_b.as<float>(); fp.logits = ns.logits.as<float>();
        fp.n_out = ns.policy_out; fp.boards = boards; fp.b_pad = b_pad; fp.n_pos = N * N;
        fp.j_tiles = (ns.policy_out + 127) / 128; fp.n_tiles = b_pad / FC_NT;
        TB_CUDA(fc_tc_launch(fp, e->num_sms, e->stream));
        e->launches += 2;
    }
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

template <int N>
static int forward_t(tak_engine* e, const uint8_t* d_states, const int* d_index, int boards, float* d_policy_out,
                     int raw_logits, const int* d_count) {
    NetState& ns = *e->net;
    if (int r = net_ensure_capacity(e, boards)) return r;
    const int S = ns.cap_S;
    const int wblocks = (boards + 7) / 8;
    // d_count: the number of boards lives on the device (the search counts its queued leaves there) and `boards` only
    // sizes the launches -- no host round trip between queueing leaves and evaluating them
    k_encode<N><<<wblocks, 256, 0, e->stream>>>(d_states, d_index, boards, ns.act_in.as<__nv_bfloat16>(), S, d_count);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    if (int r = launch_tower<N>(e, boards, d_count)) return r;
    // heads
    if (ns.arch == 6) {
        k_policy_stats_conv<N, INFER_PF><<<wblocks, 256, 0, e->stream>>>(ns.partials.as<float2>(), S, ns.policy_groups * 4, boards,
                                                                         ns.stats.as<float2>(), d_count);
        if (d_policy_out) {
            e->launches++;
            k_policy_full_conv<N><<<boards, 256, 0, e->stream>>>(ns.logits.as<float>(), S, ns.policy_ch,
                                                                 ns.stats.as<float2>(), d_policy_out, raw_logits);
        }
    } else {
        k_policy_stats_dense<<<boards, 256, 0, e->stream>>>(ns.logits.as<float>(), ns.policy_out,
                                                            ns.stats.as<float2>(), d_policy_out, raw_logits, d_count);
    }
    e->launches++;
    TB_CUDA(cudaGetLastError());
    k_value<N><<<wblocks, 256, 0, e->stream>>>(ns.trunk_out, S, ns.value_w.as<float>(), ns.value_bias, boards,
                                               ns.values.as<float>(), d_count);
    e->launches++;
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

// The fused search loop's evaluation (mcts.cu): input planes were written by the rollout warps, the number of leaves is
// on the device, the heads are computed by the backup warps.  Net5 additionally needs its FC policy GEMM and the
// softmax statistics of the dense logits (both sized for `max_boards`; rows beyond the live count are never read).
int net_tower_fast(tak_engine* e, int max_boards, const int* d_count) {
    NetState& ns = *e->net;
    int r = TAK_ERR_BAD_ARG;
    if (e->n == 5) r = launch_tower<5>(e, max_boards, d_count);
    if (e->n == 6) r = launch_tower<6>(e, max_boards, d_count);
    if (r) return r;
    if (ns.arch == 5) {
        k_policy_stats_dense<<<max_boards, 256, 0, e->stream>>>(ns.logits.as<float>(), ns.policy_out,
                                                                ns.stats.as<float2>(), nullptr, 0);
        e->launches++;
        TB_CUDA(cudaGetLastError());
    }
    return TAK_OK;
}

// where the fused loop's rollout warps write the input planes and its backup warps find the network outputs
int net_fast_views(tak_engine* e, int max_boards, FastEval& fe, PriorSource& ps) {
    TB_CHECK(e->net && e->net->arch != 0 && e->net->loaded, TAK_ERR_NO_NETWORK, "network weights not loaded");
    NetState& ns = *e->net;
    if (int r = net_ensure_capacity(e, max_boards)) return r;
    // the trunk output buffer is where the block rotation of launch_tower ends: act[0] after an even number of swaps
    ns.trunk_out = ns.act[(ns.blocks & 1) ? 2 : 0].as<__nv_bfloat16>();
    fe.planes = ns.act_in.as<__nv_bfloat16>();
    fe.S = ns.cap_S;
    ps.arch = ns.arch;
    ps.psz = ns.policy_out;
    ps.logits = ns.logits.as<float>();
    ps.stats = ns.stats.as<float2>();
    ps.values = ns.values.as<float>();
    ps.S = ns.cap_S;
    ps.partials = ns.partials.as<float2>();
    ps.groups = ns.policy_groups * 4;
    ps.trunk = ns.trunk_out;
    ps.value_w = ns.value_w.as<float>();
    ps.value_b = ns.value_bias;
    return TAK_OK;
}

int net_forward(tak_engine* e, const uint8_t* d_states, const int* d_index, int boards, float* d_policy_out,
                int raw_logits, const int* d_count) {
    TB_CHECK(e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    NetState& ns = *e->net;
    TB_CHECK(ns.arch != 0, TAK_ERR_BAD_ARG, "internal: forward on the DummyNet");
    TB_CHECK(ns.loaded, TAK_ERR_NO_NETWORK, "network weights not loaded");
    if (boards == 0) return TAK_OK;
    int r = TAK_ERR_BAD_ARG;
    if (e->n == 5) r = forward_t<5>(e, d_states, d_index, boards, d_policy_out, raw_logits, d_count);
    if (e->n == 6) r = forward_t<6>(e, d_states, d_index, boards, d_policy_out, raw_logits, d_count);
    return r;
}

void net_destroy(tak_engine* e) {
    if (!e->net) return;
    train_destroy(e);
    NetState& ns = *e->net;
    for (auto& L : ns.layers) { L.w.release(); L.bias.release(); }
    for (auto& L : ns.policy_layers) { L.w.release(); L.bias.release(); }
    for (DevBuf* b : {&ns.fc_policy_w, &ns.fc_policy_b, &ns.fc_x, &ns.value_w, &ns.act[0], &ns.act[1], &ns.act[2], &ns.act_in, &ns.logits,
                      &ns.partials, &ns.stats, &ns.values, &ns.stage_states, &ns.stage_policy, &ns.stage_repr})
        b->release();
    delete e->net;
    e->net = nullptr;
}

}  // namespace tb

using namespace tb;

extern "C" {

int32_t net_boards_per_tile(int32_t n, int32_t* out_boards) {
    TB_CHECK((n == 5 || n == 6) && out_boards, TAK_ERR_BAD_ARG, "net_boards_per_tile: n must be 5 or 6");
    *out_boards = n == 5 ? SlotMap<5, INFER_PF>::BPT : SlotMap<6, INFER_PF>::BPT;
    return TAK_OK;
}

int32_t net_input_channels(int32_t n, int32_t* out_channels) {
    TB_CHECK(n >= 3 && n <= 8 && out_channels, TAK_ERR_BAD_ARG, "net_input_channels: bad argument");
    *out_channels = input_channels_c(n);
    return TAK_OK;
}

int32_t net_create(tak_engine_t* e, int32_t arch) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    TB_CHECK(arch == 0 || arch == 5 || arch == 6, TAK_ERR_BAD_ARG, "arch must be 0 (DummyNet), 5 (Net5) or 6 (Net6)");
    TB_CHECK(arch == 0 || arch == e->n, TAK_ERR_BAD_ARG, "Net%d needs a %dx%d engine (engine is %dx%d)", arch, arch,
             arch, e->n, e->n);
    TB_CUDA(cudaSetDevice(e->device));
    net_destroy(e);
    NetState* ns = new NetState();
    ns->arch = arch;
    ns->n = e->n;
    ns->c_in = input_channels_c(e->n);
    ns->policy_out = host_policy_size(e->n);
    if (arch == 6) {
        ns->blocks = 16;                         // net6.rs:16
        ns->policy_ch = 3 + 4 * ((1 << 6) - 2);  // move_channels(6) = 251
        ns->policy_groups = 2;
    } else if (arch == 5) {
        ns->blocks = 8;                          // net5.rs:16
    }
    ns->layers.resize(arch ? 1 + 2 * ns->blocks : 0);
    ns->policy_layers.resize(ns->policy_groups);
    e->net = ns;
    return TAK_OK;
}

int32_t net_weights_size(tak_engine_t* e, int64_t* out_elems) {
    TB_CHECK(e && e->net && out_elems, TAK_ERR_NO_NETWORK, "no network");
    *out_elems = e->net->arch ? net_blob_elems(*e->net) : 0;
    return TAK_OK;
}

int32_t net_load_weights(tak_engine_t* e, const float* blob, int64_t elems) {
    TB_CHECK(e && e->net && blob, TAK_ERR_NO_NETWORK, "no network / null blob");
    TB_CUDA(cudaSetDevice(e->device));
    return net_load_blob(e, blob, elems);
}

int32_t net_load_weights_device(tak_engine_t* e, const void* device_blob, int64_t elems) {
    TB_CHECK(e && e->net && device_blob, TAK_ERR_NO_NETWORK, "no network / null blob");
    TB_CHECK(elems == net_blob_elems(*e->net), TAK_ERR_BAD_ARG, "weight blob size mismatch");
    TB_CUDA(cudaSetDevice(e->device));
    std::vector<float> host(static_cast<size_t>(elems));
    // the caller's tensor may live on another stream (torch / NCCL): a blocking copy orders after prior device work
    TB_CUDA(cudaMemcpy(host.data(), device_blob, size_t(elems) * 4, cudaMemcpyDeviceToHost));
    return net_load_blob(e, host.data(), elems);
}

static int stage_states(tak_engine_t* e, const tak_state_t* states, int b) {
    NetState& ns = *e->net;
    // packed into the engine's pinned staging buffer: the H2D copy is then a plain stream-ordered DMA (every caller
    // synchronises the stream before it returns, so the buffer is free again by the next call)
    const size_t bytes = size_t(b) * e->state_bytes;
    TB_CUDA(e->ensure_pinned(bytes));
    uint8_t* packed = static_cast<uint8_t*>(e->h_stage);
    for (int i = 0; i < b; ++i) {
        TB_CHECK(states[i].n == e->n, TAK_ERR_BAD_ARG, "state %d has board size %d, engine has %d", i, states[i].n,
                 e->n);
        pack_state(e->n, states[i], packed + size_t(i) * e->state_bytes);
    }
    TB_CUDA(ns.stage_states.ensure(bytes));
    TB_CUDA(cudaMemcpyAsync(ns.stage_states.p, packed, bytes, cudaMemcpyHostToDevice, e->stream));
    return TAK_OK;
}

int32_t net_game_repr(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out) {
    TB_CHECK(e && states && out && b >= 0, TAK_ERR_BAD_ARG, "net_game_repr: bad argument");
    if (b == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    if (!e->net) {
        int r = net_create(e, 0);
        if (r) return r;
    }
    if (int r = stage_states(e, states, b)) return r;
    NetState& ns = *e->net;
    const size_t elems = size_t(b) * input_channels_c(e->n) * e->nsq;
    TB_CUDA(ns.stage_repr.ensure(elems * 4));
    TB_DISPATCH_N(e->n, (k_repr_f32<N_><<<(b + 7) / 8, 256, 0, e->stream>>>(ns.stage_states.as<uint8_t>(), b,
                                                                           ns.stage_repr.as<float>())));
    e->launches++;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(out, ns.stage_repr.p, elems * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

static int32_t policy_eval_impl(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out_policy,
                                float* out_value, int raw_logits);
int32_t net_policy_eval(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out_policy, float* out_value) {
    return policy_eval_impl(e, states, b, out_policy, out_value, 0);
}
int32_t net_policy_logits(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out_logits, float* out_value) {
    return policy_eval_impl(e, states, b, out_logits, out_value, 1);
}
static int32_t policy_eval_impl(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out_policy,
                                float* out_value, int raw_logits) {
    TB_CHECK(e && states && out_policy && out_value && b >= 0, TAK_ERR_BAD_ARG, "net_policy_eval: bad argument");
    TB_CHECK(e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    if (b == 0) return TAK_OK;
    TB_CUDA(cudaSetDevice(e->device));
    NetState& ns = *e->net;
    const int psz = ns.policy_out;
    if (ns.arch == 0) {  // DummyNet: policy all ones, eval 0 (search/tests.rs:29-34)
        for (size_t i = 0; i < size_t(b) * psz; ++i) out_policy[i] = raw_logits ? 0.0f : 1.0f;
        for (int i = 0; i < b; ++i) out_value[i] = 0.0f;
        return TAK_OK;
    }
    const int chunk = e->max_batch;
    for (int done = 0; done < b; done += chunk) {
        const int cur = std::min(chunk, b - done);
        if (int r = stage_states(e, states + done, cur)) return r;
        TB_CUDA(ns.stage_policy.ensure(size_t(cur) * psz * 4));
        if (int r = net_forward(e, ns.stage_states.as<uint8_t>(), nullptr, cur, ns.stage_policy.as<float>(), raw_logits))
            return r;
        TB_CUDA(cudaMemcpyAsync(out_policy + size_t(done) * psz, ns.stage_policy.p, size_t(cur) * psz * 4,
                                cudaMemcpyDeviceToHost, e->stream));
        TB_CUDA(cudaMemcpyAsync(out_value + done, ns.values.p, size_t(cur) * 4, cudaMemcpyDeviceToHost, e->stream));
        TB_CUDA(cudaStreamSynchronize(e->stream));
    }
    return TAK_OK;
}

int32_t tak_host_alloc(size_t bytes, void** out) {
    TB_CHECK(out && bytes > 0, TAK_ERR_BAD_ARG, "tak_host_alloc: bad argument");
    TB_CUDA(cudaMallocHost(out, bytes));
    return TAK_OK;
}

int32_t tak_host_free(void* p) {
    if (p) TB_CUDA(cudaFreeHost(p));
    return TAK_OK;
}

int32_t net_forward_profile(tak_engine_t* e, int32_t first, int32_t count, int32_t reps, double* out4) {
    TB_CHECK(e && out4 && reps > 0 && first >= 0 && count > 0 && first + count <= e->max_games, TAK_ERR_BAD_ARG,
             "net_forward_profile: bad argument");
    TB_CHECK(e->net && e->net->arch != 0 && e->net->loaded, TAK_ERR_NO_NETWORK, "no network");
    TB_CUDA(cudaSetDevice(e->device));
    NetState& ns = *e->net;
    const uint8_t* st = e->states.as<uint8_t>() + size_t(first) * e->state_bytes;
    if (int r = net_forward(e, st, nullptr, count, nullptr)) return r;
    NetProfile prof;
    for (auto& ev : prof.ev) TB_CUDA(cudaEventCreate(&ev));
    cudaEvent_t t0, t1;
    TB_CUDA(cudaEventCreate(&t0));
    TB_CUDA(cudaEventCreate(&t1));
    double conv_ms = 0, total_ms = 0;
    int launches = 0;
    for (int i = 0; i < reps; ++i) {
        prof.n = 0;
        ns.profile = &prof;
        TB_CUDA(cudaEventRecord(t0, e->stream));
        int r = net_forward(e, st, nullptr, count, nullptr);
        ns.profile = nullptr;
        if (r) return r;
        TB_CUDA(cudaEventRecord(t1, e->stream));
        TB_CUDA(cudaEventSynchronize(t1));
        float ms = 0;
        TB_CUDA(cudaEventElapsedTime(&ms, t0, t1));
        total_ms += ms;
        for (int k = 0; k < prof.n; ++k) {
            TB_CUDA(cudaEventElapsedTime(&ms, prof.ev[2 * k], prof.ev[2 * k + 1]));
            conv_ms += ms;
        }
        launches = prof.n;
    }
    for (auto& ev : prof.ev) cudaEventDestroy(ev);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    const double nsq = double(e->nsq);
    double flop = 2.0 * nsq * 128.0 * 9.0 * (double(ns.c_in) + 2.0 * ns.blocks * 128.0);   // trunk
    if (ns.arch == 6) flop += 2.0 * nsq * double(ns.policy_ch) * 128.0 * 9.0;                // policy conv
    else flop += 2.0 * 128.0 * nsq * double(ns.policy_out);                                   // policy FC
    flop += 2.0 * 128.0 * nsq;                                                                // value FC
    out4[0] = total_ms / reps;
    out4[1] = conv_ms / reps;
    out4[2] = launches;
    out4[3] = flop * count;
    return TAK_OK;
}

int32_t net_forward_timed(tak_engine_t* e, int32_t first, int32_t count, int32_t reps, double* out_ms) {
    TB_CHECK(e && out_ms && reps > 0 && first >= 0 && count > 0 && first + count <= e->max_games, TAK_ERR_BAD_ARG,
             "net_forward_timed: bad argument");
    TB_CHECK(e->net && e->net->arch != 0, TAK_ERR_NO_NETWORK, "no network");
    TB_CUDA(cudaSetDevice(e->device));
    const uint8_t* st = e->states.as<uint8_t>() + size_t(first) * e->state_bytes;
    if (int r = net_forward(e, st, nullptr, count, nullptr)) return r;  // warm-up + allocation
    cudaEvent_t e0, e1;
    TB_CUDA(cudaEventCreate(&e0));
    TB_CUDA(cudaEventCreate(&e1));
    TB_CUDA(cudaEventRecord(e0, e->stream));
    for (int i = 0; i < reps; ++i)
        if (int r = net_forward(e, st, nullptr, count, nullptr)) return r;
    TB_CUDA(cudaEventRecord(e1, e->stream));
    TB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    TB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_ms = ms;
    return TAK_OK;
}

}  // extern "C"
