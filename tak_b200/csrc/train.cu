// Network::train for Net6 / Net5 on the device: alpha-tak/src/model/network.rs:37-97 (train / train_inner:
// loss = -sum(pi*logp)/B + sum((z-v)^2)/B accumulated over chunks, Adam lr 1e-4 wd 1e-4 every CHUNKS_IN_STEP chunks) over
// forward_training (net6.rs:111-122; BatchNorm on batch statistics) and its backward pass, which the reference gets from
// libtorch autograd.  fp32 master weights / gradients / Adam moments in the weight-blob layout (net6.rs:39-57), bf16
// operand images for the tensor-core kernels, bf16 activations saved for backward.
//   forward conv / dgrad : conv3x3_tc3_kernel<true> (CONV_LINEAR; dgrad = conv with the transposed, rotated filter)
//   wgrad                : wgrad_tc_kernel (tcgen05, MN-major operands)
//   BN / ReLU / heads / Adam : train_kernels.cuh (HBM-bound passes)
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "conv_tc3.cuh"
#include "fc_tc.cuh"
#include "net.hpp"
#include "net_kernels.cuh"
#include "train_kernels.cuh"
#include "wgrad_tc.cuh"

namespace tb {

struct TrainLayer {
    size_t w_off = 0, b_off = 0, gamma_off = 0, beta_off = 0, rm_off = 0, rv_off = 0;
    int c_in = 128;
    DevBuf w_fwd, bias_fwd, w_dgrad;
    DevBuf y, z;   // raw conv output (pre-BN) and the layer's output after BN / residual / ReLU
};

struct TrainState {
    int arch = 6, n = 6, c_in = 0, blocks = 16, policy_ch = 251, nsq = 36;
    int policy_out = 0;       // policy vector length (Net5: the FC head's 1575 outputs)
    int64_t elems = 0;
    size_t pw_off = 0, pb_off = 0, vw_off = 0, vb_off = 0;
    std::vector<TrainLayer> layers;                 // 1 + 2 * blocks
    DevBuf pol_w_fwd[2], pol_bias[2], pol_w_dgrad[2];
    DevBuf master, grad, adam_m, adam_v, zero_bias;
    DevBuf x0, g[2], dy, dy2, dt, g2, dlogits;
    DevBuf bn_sums, bn_mean, bn_rstd, bwd_sums;
    DevBuf logits, partials, stats, values, dpre, loss, wg_scratch;
    // Net5 (FC policy head, fc_tc.cuh): operand images of W, repacked activations / gradients, GEMM outputs
    DevBuf fc_wp, fc_wpt, fc_x, fc_dlx, fc_dla, fc_s2, fc_ds, fc_dwt, fc_zero_bias;
    int fc_bpad = 0, fc_kpad = 0;
    DevBuf in_stage, pi_stage, z_stage;
    int cap_boards = 0, S = 0;
    int steps = 0;          // Adam steps taken
    int chunks = 0;         // chunks accumulated since the last step
    double last_ms = 0;
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // wgrad runs on its own stream beside the dgrad / BatchNorm chain (it only feeds the gradient accumulators)
    cudaStream_t wstream = nullptr;
    cudaEvent_t ev_dy[2] = {nullptr, nullptr}, ev_w[2] = {nullptr, nullptr}, ev_join = nullptr;
    std::vector<std::pair<size_t, size_t>> trainable;   // (offset, count) of every tensor Adam updates
};

static TrainState* train_of(tak_engine* e);

static int tiles_for(int n, int boards) { return n == 5 ? SlotMap<5>::tiles(boards) : SlotMap<6>::tiles(boards); }

static void build_layout(TrainState& t) {
    size_t off = 0;
    t.layers.resize(1 + 2 * t.blocks);
    auto tensor = [&](size_t count, bool trainable) {
        const size_t o = off;
        off += count;
        if (trainable) t.trainable.emplace_back(o, count);
        return o;
    };
    auto bn = [&](TrainLayer& L) {
        L.gamma_off = tensor(128, true);
        L.beta_off = tensor(128, true);
        L.rm_off = tensor(128, false);
        L.rv_off = tensor(128, false);
    };
    {   // initial conv + BN (net6.rs:39-40)
        TrainLayer& L = t.layers[0];
        L.c_in = t.c_in;
        L.w_off = tensor(size_t(128) * t.c_in * 9, true);
        L.b_off = tensor(128, true);
        bn(L);
    }
    for (int blk = 0; blk < t.blocks; ++blk) {   // conv1, conv2, bn1, bn2 in creation order (net6.rs:43-54)
        TrainLayer& L1 = t.layers[1 + 2 * blk];
        TrainLayer& L2 = t.layers[2 + 2 * blk];
        L1.w_off = tensor(size_t(128) * 128 * 9, true);
        L1.b_off = tensor(128, true);
        L2.w_off = tensor(size_t(128) * 128 * 9, true);
        L2.b_off = tensor(128, true);
        bn(L1);
        bn(L2);
    }
    if (t.arch == 6) {   // policy conv (net6.rs:56)
        t.pw_off = tensor(size_t(t.policy_ch) * 128 * 9, true);
        t.pb_off = tensor(size_t(t.policy_ch), true);
    } else {             // policy FC (net5.rs:56-61)
        t.pw_off = tensor(size_t(t.policy_out) * 128 * t.nsq, true);
        t.pb_off = tensor(size_t(t.policy_out), true);
    }
    t.vw_off = tensor(size_t(128) * t.nsq, true);
    t.vb_off = tensor(1, true);
    t.elems = int64_t(off);
}

// fp32 master -> bf16 operand images (forward and dgrad) of every conv
static int repack(tak_engine* e, TrainState& t) {
    const float* m = t.master.as<float>();
    const int blocks = int((C3_W_LAYER_ELEMS + 255) / 256);
    for (size_t l = 0; l < t.layers.size(); ++l) {
        TrainLayer& L = t.layers[l];
        k_pack_conv_train<<<blocks, 256, 0, e->stream>>>(m + L.w_off, 128, L.c_in, 0, 0, 0, L.w_fwd.as<__nv_bfloat16>());
        k_pack_bias_train<<<1, 128, 0, e->stream>>>(m + L.b_off, 128, 0, L.bias_fwd.as<float>());
        if (l > 0)
            k_pack_conv_train<<<blocks, 256, 0, e->stream>>>(m + L.w_off, 128, 128, 0, 0, 1, L.w_dgrad.as<__nv_bfloat16>());
        t.launches += 3;
    }
    if (t.arch == 5) {
        const int J = t.policy_out, K = 128 * t.nsq, P = (J + 127) / 128;
        const size_t n_fwd = size_t(P) * 128 * t.nsq * 128, n_dg = size_t(K) * P * 128;
        k_fc_pack_fwd<<<unsigned((n_fwd + 255) / 256), 256, 0, e->stream>>>(m + t.pw_off, J, t.nsq, P,
                                                                            t.fc_wp.as<__nv_bfloat16>());
        k_fc_pack_dgrad<<<unsigned((n_dg + 255) / 256), 256, 0, e->stream>>>(m + t.pw_off, J, K, P,
                                                                             t.fc_wpt.as<__nv_bfloat16>());
        t.launches += 2;
    }
    for (int grp = 0; grp < (t.arch == 6 ? 2 : 0); ++grp) {
        k_pack_conv_train<<<blocks, 256, 0, e->stream>>>(m + t.pw_off, t.policy_ch, 128, grp * 128, 0, 0,
                                                         t.pol_w_fwd[grp].as<__nv_bfloat16>());
        k_pack_bias_train<<<1, 128, 0, e->stream>>>(m + t.pb_off, t.policy_ch, grp * 128, t.pol_bias[grp].as<float>());
        // dgrad of the policy conv: 128 outputs (trunk channels), K = policy channels grp*128 .. grp*128+127
        k_pack_conv_train<<<blocks, 256, 0, e->stream>>>(m + t.pw_off, t.policy_ch, 128, 0, grp * 128, 1,
                                                         t.pol_w_dgrad[grp].as<__nv_bfloat16>());
        t.launches += 3;
    }
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

static int ensure_capacity(tak_engine* e, TrainState& t, int boards) {
    if (boards <= t.cap_boards) return TAK_OK;
    const int tiles = tiles_for(t.n, boards);
    const int S = tiles * C3_TILE_M;
    const size_t plane = size_t(S) * 256;   // 16 chunks x S slots x 16 B
    auto planes = [&](DevBuf& b, int stacks) -> cudaError_t {
        cudaError_t r = b.ensure(plane * stacks);
        if (r != cudaSuccess) return r;
        return cudaMemsetAsync(b.p, 0, plane * stacks, e->stream);
    };
    TB_CUDA(planes(t.x0, 1));
    for (auto& L : t.layers) {
        TB_CUDA(planes(L.y, 1));
        TB_CUDA(planes(L.z, 1));
    }
    TB_CUDA(planes(t.g[0], 1));
    TB_CUDA(planes(t.g[1], 1));
    TB_CUDA(planes(t.dy, 1));
    TB_CUDA(planes(t.dy2, 1));
    TB_CUDA(planes(t.dt, 1));
    TB_CUDA(planes(t.g2, 1));
    if (t.arch == 6) {
        TB_CUDA(planes(t.dlogits, 2));
        TB_CUDA(t.logits.ensure(size_t(256) * S * 4));
        TB_CUDA(t.partials.ensure(size_t(8) * S * 8));
    } else {
        const int J = t.policy_out, K = 128 * t.nsq, P = (J + 127) / 128;
        t.fc_bpad = (boards + FC_NT - 1) / FC_NT * FC_NT;
        t.fc_kpad = (K + FC_NT - 1) / FC_NT * FC_NT;
        TB_CUDA(t.logits.ensure(size_t(boards) * J * 4));
        TB_CUDA(t.fc_x.ensure(size_t(t.nsq) * 16 * t.fc_bpad * 16));
        TB_CUDA(t.fc_dlx.ensure(size_t(P) * 16 * t.fc_bpad * 16));
        TB_CUDA(t.fc_dla.ensure(size_t(P) * (t.fc_bpad / 128) * FC_W_POS_BYTES));
        TB_CUDA(t.fc_s2.ensure(size_t(t.fc_bpad / 128) * 16 * t.fc_kpad * 16));
        TB_CUDA(t.fc_ds.ensure(size_t(boards) * K * 4));
        TB_CUDA(t.fc_dwt.ensure(size_t(K) * J * 4));
    }
    TB_CUDA(t.stats.ensure(size_t(boards) * 8));
    TB_CUDA(t.values.ensure(size_t(boards) * 4));
    TB_CUDA(t.dpre.ensure(size_t(boards) * 4));
    t.cap_boards = boards;
    t.S = S;
    return TAK_OK;
}

template <int N>
static int train_chunk_t(tak_engine* e, TrainState& t, const float* d_in, const float* d_pi, const float* d_z, int B) {
    const int tiles = tiles_for(N, B);
    const int S = t.S;
    float* master = t.master.as<float>();
    float* grad = t.grad.as<float>();
    const int nl = int(t.layers.size());
    const double count = double(B) * N * N;
    const int ew_blocks = int((size_t(16) * S + 255) / 256);
    using bf = __nv_bfloat16;
    // bnb_layer >= 0 (dgrad launches): `out` is the gradient w.r.t. the post-ReLU output of trunk layer bnb_layer; the
    // epilogue applies that layer's ReLU mask and accumulates its BatchNorm-backward sums (ConvParams::bnb_y), so the
    // separate reduction pass over g, z and y is gone and the apply pass reads the masked gradient only
    auto conv_lin = [&](const bf* in, const bf* w, const float* bias, const bf* res, bf* out, double* stats, int slabs,
                        int bnb_layer = -1) -> int {
        ConvParams p{};
        conv_params_set_layout(p, N);
        p.S = S; p.tile_begin = 0; p.tile_end = tiles; p.n_boards = B; p.n_layers = 1;
        ConvLayerDesc& d = p.layers[0];
        d.in = in; d.res = res; d.out = out; d.w = w; d.bias = bias; d.slabs = slabs; d.mode = CONV_LINEAR;
        d.out_ch_valid = 128; d.stats = stats;
        if (bnb_layer >= 0) {
            p.bnb_z = t.layers[bnb_layer].z.as<bf>();
            p.bnb_y = t.layers[bnb_layer].y.as<bf>();
            p.bnb_mean = t.bn_mean.as<float>() + bnb_layer * 128;
            p.bnb_rstd = t.bn_rstd.as<float>() + bnb_layer * 128;
            d.stats = t.bwd_sums.as<double>() + size_t(bnb_layer) * 256;
        }
        TB_CUDA(conv3x3_tc3_launch<true>(p, e->num_sms, e->stream));
        t.launches++;
        return TAK_OK;
    };
    // wgrad of a layer only feeds the gradient accumulators, so it runs on `wstream` beside the dgrad / BatchNorm chain
    // of the main stream (the HBM-bound passes fit next to a wgrad CTA on an SM).  `slot` names the dy buffer it reads:
    // ev_dy[slot] = dy written (main -> wstream), ev_w[slot] = wgrad done reading it (wstream -> main, before reuse).
    auto wgrad = [&](const bf* dy, const bf* x, int c_in, size_t w_off, int co_base, int co_valid, int slot) -> int {
        TB_CUDA(cudaEventRecord(t.ev_dy[slot], e->stream));
        TB_CUDA(cudaStreamWaitEvent(t.wstream, t.ev_dy[slot], 0));
        TB_CUDA(wgrad_tc_launch(dy, x, S, tiles, SlotMap<N>::PITCH, c_in, t.wg_scratch.as<float>(), grad + w_off, co_base,
                                co_valid, 1, e->num_sms, t.wstream));
        TB_CUDA(cudaEventRecord(t.ev_w[slot], t.wstream));
        t.launches += 2;
        return TAK_OK;
    };
    auto reuse_dy = [&](int slot) -> int {   // the main stream is about to overwrite dy buffer `slot`
        TB_CUDA(cudaStreamWaitEvent(e->stream, t.ev_w[slot], 0));
        return TAK_OK;
    };

    // ------------------------------------------------ forward_training ------------------------------------------------
    k_nchw_to_planes<N><<<ew_blocks, 256, 0, e->stream>>>(d_in, B, t.c_in, t.x0.as<bf>(), S);
    t.launches++;
    TB_CUDA(cudaMemsetAsync(t.bn_sums.p, 0, size_t(nl) * 256 * 8, e->stream));
    TB_CUDA(cudaMemsetAsync(t.bwd_sums.p, 0, size_t(nl) * 256 * 8, e->stream));
    // TAK_TRAIN_BNB=0: the two-pass BatchNorm backward everywhere (A/B and fallback)
    static const bool fuse_bnb = [] { const char* v = getenv("TAK_TRAIN_BNB"); return !v || atoi(v) != 0; }();
    for (int l = 0; l < nl; ++l) {
        TrainLayer& L = t.layers[l];
        const bf* in = l == 0 ? t.x0.as<bf>() : t.layers[l - 1].z.as<bf>();
        double* sums = t.bn_sums.as<double>() + size_t(l) * 256;
        if (int r = conv_lin(in, L.w_fwd.as<bf>(), L.bias_fwd.as<float>(), nullptr, L.y.as<bf>(), sums,
                             l == 0 ? (t.c_in + 15) / 16 : C3_MAX_SLABS))
            return r;
        float* mean = t.bn_mean.as<float>() + l * 128;
        float* rstd = t.bn_rstd.as<float>() + l * 128;
        // conv2 of a block adds the block input before the ReLU (res_block.rs:21-22)
        const bool is_conv2 = l >= 2 && (l % 2) == 0;
        const bf* res = is_conv2 ? t.layers[l - 2].z.as<bf>() : nullptr;
        k_bn_apply<N><<<ew_grid(S), 256, 0, e->stream>>>(L.y.as<bf>(), res, sums, count, master + L.gamma_off,
                                                        master + L.beta_off, master + L.rm_off, master + L.rv_off, mean,
                                                        rstd, B, S, L.z.as<bf>());
        t.launches += 1;
    }
    const bf* trunk = t.layers[nl - 1].z.as<bf>();
    bf* g = t.g[0].as<bf>();
    bf* g_alt = t.g[1].as<bf>();
    const int wblocks = (B + 7) / 8;
    bool top_premasked = false;   // the top gradient arrives masked, with the top layer's sums (Net6, fused)
    if constexpr (N == 5) {
        // ---- Net5 heads: FC policy (net5.rs:106-108) + value; loss; head gradients -- three GEMMs on fc_tc_kernel ----
        const int J = t.policy_out, K = 128 * N * N, P = (J + 127) / 128, b_pad = t.fc_bpad, k_pad = t.fc_kpad;
        auto gemm = [&](const bf* a, const bf* x, float* out, int n_out, int n, int n_pad, int n_pos, cudaStream_t st) -> int {
            FcParams fp{};
            fp.wp = a; fp.x = x; fp.bias = t.fc_zero_bias.as<float>(); fp.logits = out;
            fp.n_out = n_out; fp.boards = n; fp.b_pad = n_pad; fp.n_pos = n_pos;
            fp.j_tiles = (n_out + 127) / 128; fp.n_tiles = n_pad / FC_NT;
            TB_CUDA(fc_tc_launch(fp, e->num_sms, st));
            t.launches++;
            return TAK_OK;
        };
        {   // forward: logits[b][j] = bias[j] + W s
            const size_t items = size_t(N * N) * 16 * b_pad;
            k_fc_repack<N><<<unsigned((items + 255) / 256), 256, 0, e->stream>>>(trunk, S, B, b_pad, t.fc_x.as<bf>());
            FcParams fp{};
            fp.wp = t.fc_wp.as<bf>(); fp.x = t.fc_x.as<bf>(); fp.bias = master + t.pb_off; fp.logits = t.logits.as<float>();
            fp.n_out = J; fp.boards = B; fp.b_pad = b_pad; fp.n_pos = N * N; fp.j_tiles = P; fp.n_tiles = b_pad / FC_NT;
            TB_CUDA(fc_tc_launch(fp, e->num_sms, e->stream));
            k_policy_stats_dense<<<B, 256, 0, e->stream>>>(t.logits.as<float>(), J, t.stats.as<float2>(), nullptr, 0);
            k_value_train<N><<<wblocks, 256, 0, e->stream>>>(trunk, S, master + t.vw_off, master + t.vb_off, B,
                                                             t.values.as<float>());
            t.launches += 4;
        }
        TB_CUDA(cudaMemsetAsync(t.loss.p, 0, 16, e->stream));
        TB_CUDA(cudaMemsetAsync(t.fc_dlx.p, 0, t.fc_dlx.bytes, e->stream));
        TB_CUDA(cudaMemsetAsync(t.fc_dla.p, 0, t.fc_dla.bytes, e->stream));
        k_fc_loss_grad<<<B, 256, 0, e->stream>>>(t.logits.as<float>(), J, t.stats.as<float2>(), d_pi, B, b_pad,
                                                 t.fc_dlx.as<bf>(), t.fc_dla.as<bf>(), t.loss.as<double>());
        k_planes_colsum<<<dim3(BNR_SPLIT, P * 16), 256, 0, e->stream>>>(t.fc_dlx.as<bf>(), b_pad, J, grad + t.pb_off);
        k_value_loss_grad<<<(B + 255) / 256, 256, 0, e->stream>>>(t.values.as<float>(), d_z, B, t.dpre.as<float>(),
                                                                  t.loss.as<double>(), grad + t.vb_off);
        k_value_wgrad<N><<<dim3(16 * N * N, 16), 128, 0, e->stream>>>(t.dpre.as<float>(), trunk, B, S, grad + t.vw_off);
        t.launches += 4;
        // wgrad: dW^T[k][j] = sum_b s[b][k] dl[b][j]  (on the wgrad stream), then grad W += (dW^T)^T
        {
            TB_CUDA(cudaEventRecord(t.ev_dy[0], e->stream));
            TB_CUDA(cudaStreamWaitEvent(t.wstream, t.ev_dy[0], 0));
            const size_t items = size_t(b_pad / 8) * k_pad;
            k_fc_repack_wgrad<N><<<unsigned((items + 255) / 256), 256, 0, t.wstream>>>(trunk, S, B, b_pad, k_pad,
                                                                                     t.fc_s2.as<bf>());
            if (int r = gemm(t.fc_dla.as<bf>(), t.fc_s2.as<bf>(), t.fc_dwt.as<float>(), J, K, k_pad, b_pad / 128, t.wstream))
                return r;
            k_fc_wgrad_add<<<unsigned((size_t(J) * K + 255) / 256), 256, 0, t.wstream>>>(t.fc_dwt.as<float>(), J, K,
                                                                                        grad + t.pw_off);
            TB_CUDA(cudaEventRecord(t.ev_w[0], t.wstream));
            t.launches += 2;
        }
        // dgrad: ds[b][k] = sum_j dl[b][j] W[j][k]; trunk gradient = ds + the value head's share
        if (int r = gemm(t.fc_wpt.as<bf>(), t.fc_dlx.as<bf>(), t.fc_ds.as<float>(), K, B, b_pad, P, e->stream)) return r;
        k_fc_ds_to_planes<N><<<ew_blocks, 256, 0, e->stream>>>(t.fc_ds.as<float>(), t.dpre.as<float>(), master + t.vw_off,
                                                               B, S, g);
        t.launches++;
        TB_CUDA(cudaGetLastError());
    } else {
    {   // policy conv -> fp32 logits + per-slot softmax partials (net6.rs:113-116); value head (net6.rs:117-121)
        ConvParams p{};
        conv_params_set_layout(p, N);
        p.S = S; p.tile_begin = 0; p.tile_end = tiles; p.n_boards = B;
        for (int grp = 0; grp < 2; ++grp) {
            ConvLayerDesc& d = p.layers[p.n_layers++];
            d.in = trunk; d.out_f32 = t.logits.as<float>(); d.partials = t.partials.as<float2>();
            d.w = t.pol_w_fwd[grp].as<bf>(); d.bias = t.pol_bias[grp].as<float>();
            d.slabs = C3_MAX_SLABS; d.mode = CONV_LOGITS_F32; d.out_ch_offset = grp * 128;
            d.out_ch_valid = std::min(128, t.policy_ch - grp * 128); d.group = grp;
        }
        TB_CUDA(conv3x3_tc3_launch<false>(p, e->num_sms, e->stream));
        k_policy_stats_conv<N><<<wblocks, 256, 0, e->stream>>>(t.partials.as<float2>(), S, 8, B, t.stats.as<float2>());
        k_value_train<N><<<wblocks, 256, 0, e->stream>>>(trunk, S, master + t.vw_off, master + t.vb_off, B,
                                                         t.values.as<float>());
        t.launches += 3;
    }
    // ------------------------------------------------ loss and head gradients -----------------------------------------
    TB_CUDA(cudaMemsetAsync(t.loss.p, 0, 16, e->stream));
    TB_CUDA(cudaMemsetAsync(t.dlogits.p, 0, size_t(S) * 512, e->stream));
    k_policy_loss_grad<N><<<B, 256, 0, e->stream>>>(t.logits.as<float>(), S, t.policy_ch, t.stats.as<float2>(), d_pi, B,
                                                    t.dlogits.as<bf>(), t.loss.as<double>());
    k_planes_colsum<<<dim3(BNR_SPLIT, 32), 256, 0, e->stream>>>(t.dlogits.as<bf>(), S, t.policy_ch, grad + t.pb_off);
    k_value_loss_grad<<<(B + 255) / 256, 256, 0, e->stream>>>(t.values.as<float>(), d_z, B, t.dpre.as<float>(),
                                                              t.loss.as<double>(), grad + t.vb_off);
    k_value_wgrad<N><<<dim3(16 * N * N, 16), 128, 0, e->stream>>>(t.dpre.as<float>(), trunk, B, S, grad + t.vw_off);
    k_value_bwd_trunk<N><<<ew_blocks, 256, 0, e->stream>>>(t.dpre.as<float>(), master + t.vw_off, B, S, g);
    t.launches += 5;
    TB_CUDA(cudaGetLastError());
    // policy conv: wgrad per 128-channel group, dgrad in two K passes accumulated through the residual input
    const bf* dl0 = t.dlogits.as<bf>();
    const bf* dl1 = dl0 + size_t(16) * S * 8;
    if (int r = wgrad(dl0, trunk, 128, t.pw_off, 0, 128, 0)) return r;
    if (int r = wgrad(dl1, trunk, 128, t.pw_off, 128, t.policy_ch - 128, 1)) return r;
    if (int r = conv_lin(dl0, t.pol_w_dgrad[0].as<bf>(), t.zero_bias.as<float>(), g, g_alt, nullptr, C3_MAX_SLABS)) return r;
    if (int r = conv_lin(dl1, t.pol_w_dgrad[1].as<bf>(), t.zero_bias.as<float>(), g_alt, g, nullptr, C3_MAX_SLABS,
                         fuse_bnb ? nl - 1 : -1))
        return r;
    top_premasked = true;
    }   // N == 6 heads
    // ------------------------------------------------ trunk backward --------------------------------------------------
    // BatchNorm backward of layer l.  `gm` is the gradient w.r.t. the layer's post-ReLU output ALREADY masked by that
    // ReLU, and bwd_sums[l] holds sum g' / sum g' * xhat: both come out of the epilogue of the dgrad launch that produced
    // gm (conv_lin(..., bnb_layer = l)).  One pass is left: dy = gamma * rstd * (g' - c1 - xhat * c2), reading g' and y.
    // premasked == false (Net5's top layer, whose upstream gradient comes from the FC head's kernels): the two-pass
    // route, k_bn_bwd_reduce + masking apply, with the masked gradient written to `gmasked`.
    auto bn_backward = [&](int l, const bf* gin, bf* dy_out, bool premasked, bf* gmasked) -> int {
        TrainLayer& L = t.layers[l];
        const float* mean = t.bn_mean.as<float>() + l * 128;
        const float* rstd = t.bn_rstd.as<float>() + l * 128;
        double* sums = t.bwd_sums.as<double>() + size_t(l) * 256;
        const bf* zmask = premasked ? nullptr : L.z.as<bf>();
        if (!premasked) {
            k_bn_bwd_reduce<<<dim3(BNR_SPLIT, 16), 256, 0, e->stream>>>(gin, zmask, L.y.as<bf>(), mean, rstd, S, sums);
            t.launches++;
        }
        if (premasked)
            k_bn_bwd_apply<N, false><<<ew_grid(S), 256, 0, e->stream>>>(gin, nullptr, L.y.as<bf>(), mean, rstd,
                                                                       master + L.gamma_off, sums, count, grad + L.gamma_off,
                                                                       grad + L.beta_off, B, S, dy_out, nullptr);
        else
            k_bn_bwd_apply<N, true><<<ew_grid(S), 256, 0, e->stream>>>(gin, zmask, L.y.as<bf>(), mean, rstd,
                                                                      master + L.gamma_off, sums, count, grad + L.gamma_off,
                                                                      grad + L.beta_off, B, S, dy_out, gmasked);
        t.launches++;
        return TAK_OK;
    };
    const bool fuse = fuse_bnb;
    auto bnb = [&](int l) { return fuse ? l : -1; };
    bool premasked = fuse && top_premasked;
    for (int blk = t.blocks - 1; blk >= 0; --blk) {
        const int l1 = 1 + 2 * blk, l2 = 2 + 2 * blk;
        const bf* xin = t.layers[l1 - 1].z.as<bf>();
        // out = relu(bn2(conv2(t)) + x): the masked gradient g' = g * (out > 0) goes to both branches
        if (int r = reuse_dy(0)) return r;
        if (int r = bn_backward(l2, g, t.dy.as<bf>(), premasked, t.g2.as<bf>())) return r;
        const bf* gres = premasked ? g : t.g2.as<bf>();
        if (int r = wgrad(t.dy.as<bf>(), t.layers[l1].z.as<bf>(), 128, t.layers[l2].w_off, 0, 128, 0)) return r;
        // dt = gradient w.r.t. conv1's post-ReLU output (fused: masked, with bn1's sums)
        if (int r = conv_lin(t.dy.as<bf>(), t.layers[l2].w_dgrad.as<bf>(), t.zero_bias.as<float>(), nullptr, t.dt.as<bf>(),
                             nullptr, C3_MAX_SLABS, bnb(l1)))
            return r;
        if (int r = reuse_dy(1)) return r;
        if (int r = bn_backward(l1, t.dt.as<bf>(), t.dy2.as<bf>(), fuse, nullptr)) return r;
        if (int r = wgrad(t.dy2.as<bf>(), xin, 128, t.layers[l1].w_off, 0, 128, 1)) return r;
        // gradient w.r.t. the block input = dgrad(conv1) + the residual share (fused: masked by the ReLU of the layer
        // that produced the block input -- the previous block's conv2, or the initial conv -- with that layer's sums)
        if (int r = conv_lin(t.dy2.as<bf>(), t.layers[l1].w_dgrad.as<bf>(), t.zero_bias.as<float>(), gres, g_alt, nullptr,
                             C3_MAX_SLABS, bnb(l1 - 1)))
            return r;
        std::swap(g, g_alt);
        premasked = fuse;
    }
    if (int r = reuse_dy(0)) return r;
    if (int r = bn_backward(0, g, t.dy.as<bf>(), premasked, nullptr)) return r;
    if (int r = wgrad(t.dy.as<bf>(), t.x0.as<bf>(), t.c_in, t.layers[0].w_off, 0, 128, 0)) return r;
    // join: the chunk is complete on the main stream only when the last wgrads are
    TB_CUDA(cudaEventRecord(t.ev_join, t.wstream));
    TB_CUDA(cudaStreamWaitEvent(e->stream, t.ev_join, 0));
    // conv biases feed a BatchNorm on batch statistics: their gradient is identically zero (sum of dy over the batch
    // vanishes), so only the weight decay term reaches them in Adam -- nothing to accumulate here.
    TB_CUDA(cudaGetLastError());
    return TAK_OK;
}

}  // namespace tb

using namespace tb;

// the TrainState hangs off the NetState through an opaque pointer (net.hpp keeps no training types)
namespace tb {
static TrainState* train_of(tak_engine* e) { return e->net ? static_cast<TrainState*>(e->net->train) : nullptr; }
void train_destroy(tak_engine* e) {
    TrainState* t = train_of(e);
    if (!t) return;
    for (auto& L : t->layers)
        for (DevBuf* b : {&L.w_fwd, &L.bias_fwd, &L.w_dgrad, &L.y, &L.z}) b->release();
    for (int i = 0; i < 2; ++i)
        for (DevBuf* b : {&t->pol_w_fwd[i], &t->pol_bias[i], &t->pol_w_dgrad[i], &t->g[i]}) b->release();
    for (DevBuf* b : {&t->master, &t->grad, &t->adam_m, &t->adam_v, &t->zero_bias, &t->x0, &t->dy, &t->dt, &t->g2,
                      &t->dy2, &t->dlogits, &t->bn_sums, &t->bn_mean, &t->bn_rstd, &t->bwd_sums, &t->logits, &t->partials, &t->stats, &t->values, &t->dpre, &t->loss, &t->wg_scratch,
                      &t->in_stage, &t->pi_stage, &t->z_stage, &t->fc_wp, &t->fc_wpt, &t->fc_x, &t->fc_dlx, &t->fc_dla,
                      &t->fc_s2, &t->fc_ds, &t->fc_dwt, &t->fc_zero_bias})
        b->release();
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    for (cudaEvent_t ev : {t->ev_dy[0], t->ev_dy[1], t->ev_w[0], t->ev_w[1], t->ev_join})
        if (ev) cudaEventDestroy(ev);
    if (t->wstream) cudaStreamDestroy(t->wstream);
    delete t;
    e->net->train = nullptr;
}
}  // namespace tb

extern "C" {

int32_t net_train_begin(tak_engine_t* e, int32_t max_boards) {
    TB_CHECK(e && e->net, TAK_ERR_NO_NETWORK, "no network: call net_create first");
    NetState& ns = *e->net;
    TB_CHECK((ns.arch == 6 && e->n == 6) || (ns.arch == 5 && e->n == 5), TAK_ERR_BAD_ARG,
             "training needs a Net5 or Net6 engine (the DummyNet has no weights)");
    TB_CHECK(ns.loaded && int64_t(ns.blob_host.size()) == net_blob_elems(ns), TAK_ERR_NO_NETWORK,
             "load weights (net_load_weights) before net_train_begin");
    TB_CHECK(max_boards > 0, TAK_ERR_BAD_ARG, "max_boards must be positive");
    TB_CUDA(cudaSetDevice(e->device));
    train_destroy(e);
    TrainState* t = new TrainState();
    ns.train = t;
    t->arch = ns.arch;
    t->n = e->n;
    t->c_in = ns.c_in;
    t->blocks = ns.blocks;
    t->policy_ch = ns.policy_ch;
    t->policy_out = ns.policy_out;
    t->nsq = e->nsq;
    build_layout(*t);
    TB_CHECK(t->elems == net_blob_elems(ns), TAK_ERR_BAD_ARG, "internal: training layout does not match the blob");
    const size_t bytes = size_t(t->elems) * 4;
    for (DevBuf* b : {&t->master, &t->grad, &t->adam_m, &t->adam_v}) {
        TB_CUDA(b->ensure(bytes));
        TB_CUDA(cudaMemsetAsync(b->p, 0, bytes, e->stream));
    }
    TB_CUDA(cudaMemcpyAsync(t->master.p, ns.blob_host.data(), bytes, cudaMemcpyHostToDevice, e->stream));
    TB_CUDA(t->zero_bias.ensure(512));
    TB_CUDA(cudaMemsetAsync(t->zero_bias.p, 0, 512, e->stream));
    for (auto& L : t->layers) {
        TB_CUDA(L.w_fwd.ensure(C3_W_LAYER_ELEMS * 2));
        TB_CUDA(L.w_dgrad.ensure(C3_W_LAYER_ELEMS * 2));
        TB_CUDA(L.bias_fwd.ensure(512));
    }
    if (t->arch == 5) {
        const int J = t->policy_out, K = 128 * t->nsq, P = (J + 127) / 128;
        TB_CUDA(t->fc_wp.ensure(size_t(P) * t->nsq * FC_W_POS_BYTES));
        TB_CUDA(t->fc_wpt.ensure(size_t(K / 128) * P * FC_W_POS_BYTES));
        TB_CUDA(t->fc_zero_bias.ensure(size_t(K) * 4));
        TB_CUDA(cudaMemsetAsync(t->fc_zero_bias.p, 0, size_t(K) * 4, e->stream));
    }
    for (int i = 0; i < (t->arch == 6 ? 2 : 0); ++i) {
        TB_CUDA(t->pol_w_fwd[i].ensure(C3_W_LAYER_ELEMS * 2));
        TB_CUDA(t->pol_w_dgrad[i].ensure(C3_W_LAYER_ELEMS * 2));
        TB_CUDA(t->pol_bias[i].ensure(512));
    }
    const size_t nl = t->layers.size();
    TB_CUDA(t->bn_sums.ensure(nl * 256 * 8));
    for (DevBuf* b : {&t->bn_mean, &t->bn_rstd}) TB_CUDA(b->ensure(nl * 128 * 4));
    TB_CUDA(t->bwd_sums.ensure(nl * 256 * 8));
    TB_CUDA(t->loss.ensure(16));
    TB_CUDA(t->wg_scratch.ensure(wgrad_scratch_elems(e->num_sms) * 4));
    TB_CUDA(cudaEventCreate(&t->ev0));
    TB_CUDA(cudaEventCreate(&t->ev1));
    TB_CUDA(cudaStreamCreateWithFlags(&t->wstream, cudaStreamNonBlocking));
    for (cudaEvent_t* ev : {&t->ev_dy[0], &t->ev_dy[1], &t->ev_w[0], &t->ev_w[1], &t->ev_join})
        TB_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    // ev_w[*] must be "complete" before the first reuse_dy of a chunk
    TB_CUDA(cudaEventRecord(t->ev_w[0], t->wstream));
    TB_CUDA(cudaEventRecord(t->ev_w[1], t->wstream));
    if (int r = ensure_capacity(e, *t, max_boards)) return r;
    if (int r = repack(e, *t)) return r;
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t net_train_chunk(tak_engine_t* e, const float* inputs, const float* pi, const float* z, int32_t boards,
                        int32_t on_device, float* out_loss2) {
    TB_CHECK(e && e->net && inputs && pi && z && boards > 0, TAK_ERR_BAD_ARG, "net_train_chunk: bad argument");
    TrainState* t = train_of(e);
    TB_CHECK(t, TAK_ERR_NO_NETWORK, "call net_train_begin first");
    TB_CHECK(boards <= t->cap_boards, TAK_ERR_CAPACITY, "chunk of %d boards, net_train_begin reserved %d", boards,
             t->cap_boards);
    TB_CUDA(cudaSetDevice(e->device));
    const size_t in_elems = size_t(boards) * t->c_in * t->nsq, pi_elems = size_t(boards) * t->policy_out;
    const float *d_in = inputs, *d_pi = pi, *d_z = z;
    TB_CUDA(cudaEventRecord(t->ev0, e->stream));
    if (!on_device) {
        TB_CUDA(t->in_stage.ensure(in_elems * 4));
        TB_CUDA(t->pi_stage.ensure(pi_elems * 4));
        TB_CUDA(t->z_stage.ensure(size_t(boards) * 4));
        TB_CUDA(cudaMemcpyAsync(t->in_stage.p, inputs, in_elems * 4, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaMemcpyAsync(t->pi_stage.p, pi, pi_elems * 4, cudaMemcpyHostToDevice, e->stream));
        TB_CUDA(cudaMemcpyAsync(t->z_stage.p, z, size_t(boards) * 4, cudaMemcpyHostToDevice, e->stream));
        d_in = t->in_stage.as<float>(); d_pi = t->pi_stage.as<float>(); d_z = t->z_stage.as<float>();
    }
    const uint64_t before = t->launches;
    if (int r = t->n == 5 ? train_chunk_t<5>(e, *t, d_in, d_pi, d_z, boards) : train_chunk_t<6>(e, *t, d_in, d_pi, d_z, boards))
        return r;
    e->launches += t->launches - before;
    TB_CUDA(cudaEventRecord(t->ev1, e->stream));
    double loss[2] = {0, 0};
    TB_CUDA(cudaMemcpyAsync(loss, t->loss.p, 16, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    float ms = 0;
    TB_CUDA(cudaEventElapsedTime(&ms, t->ev0, t->ev1));
    t->last_ms = ms;
    t->chunks++;
    if (out_loss2) {
        out_loss2[0] = float(loss[0]);
        out_loss2[1] = float(loss[1]);
    }
    return TAK_OK;
}

int32_t net_train_step(tak_engine_t* e, float lr, float weight_decay) {
    TB_CHECK(e && e->net, TAK_ERR_BAD_ARG, "net_train_step: bad argument");
    TrainState* t = train_of(e);
    TB_CHECK(t, TAK_ERR_NO_NETWORK, "call net_train_begin first");
    TB_CUDA(cudaSetDevice(e->device));
    t->steps++;
    const float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
    const float bc1 = float(1.0 - std::pow(double(beta1), t->steps));
    const float bc2s = float(std::sqrt(1.0 - std::pow(double(beta2), t->steps)));
    for (auto& tr : t->trainable) {
        const int count = int(tr.second);
        k_adam<<<(count + 255) / 256, 256, 0, e->stream>>>(t->master.as<float>() + tr.first, t->grad.as<float>() + tr.first,
                                                           t->adam_m.as<float>() + tr.first,
                                                           t->adam_v.as<float>() + tr.first, count, lr, weight_decay,
                                                           beta1, beta2, eps, bc1, bc2s);
        e->launches++;
    }
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemsetAsync(t->grad.p, 0, size_t(t->elems) * 4, e->stream));   // opt.zero_grad() (network.rs:94)
    t->chunks = 0;
    if (int r = repack(e, *t)) return r;
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t net_train_get(tak_engine_t* e, int32_t what, float* out, int64_t elems) {
    TB_CHECK(e && e->net && out, TAK_ERR_BAD_ARG, "net_train_get: bad argument");
    TrainState* t = train_of(e);
    TB_CHECK(t, TAK_ERR_NO_NETWORK, "call net_train_begin first");
    TB_CHECK(elems == t->elems, TAK_ERR_BAD_ARG, "blob has %lld elements, caller passed %lld", (long long)t->elems,
             (long long)elems);
    TB_CHECK(what >= 0 && what <= 3, TAK_ERR_BAD_ARG, "what: 0 weights, 1 gradients, 2 Adam m, 3 Adam v");
    TB_CUDA(cudaSetDevice(e->device));
    const DevBuf* src[4] = {&t->master, &t->grad, &t->adam_m, &t->adam_v};
    TB_CUDA(cudaMemcpyAsync(out, src[what]->p, size_t(elems) * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t net_train_grad_ptr(tak_engine_t* e, void** out_device_ptr, int64_t* out_elems) {
    TB_CHECK(e && e->net && out_device_ptr && out_elems, TAK_ERR_BAD_ARG, "net_train_grad_ptr: bad argument");
    TrainState* t = train_of(e);
    TB_CHECK(t, TAK_ERR_NO_NETWORK, "call net_train_begin first");
    TB_CUDA(cudaSetDevice(e->device));
    TB_CUDA(cudaStreamSynchronize(e->stream));   // the caller (NCCL all-reduce on its own stream) may touch it now
    *out_device_ptr = t->grad.p;
    *out_elems = t->elems;
    return TAK_OK;
}

int32_t net_train_stats(tak_engine_t* e, double* out_ms_last_chunk, int32_t* out_chunks_pending, int32_t* out_steps) {
    TB_CHECK(e && e->net, TAK_ERR_BAD_ARG, "net_train_stats: bad argument");
    TrainState* t = train_of(e);
    TB_CHECK(t, TAK_ERR_NO_NETWORK, "call net_train_begin first");
    if (out_ms_last_chunk) *out_ms_last_chunk = t->last_ms;
    if (out_chunks_pending) *out_chunks_pending = t->chunks;
    if (out_steps) *out_steps = t->steps;
    return TAK_OK;
}

int32_t net_train_end(tak_engine_t* e) {
    TB_CHECK(e && e->net, TAK_ERR_BAD_ARG, "net_train_end: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    train_destroy(e);
    return TAK_OK;
}

}  // extern "C"
