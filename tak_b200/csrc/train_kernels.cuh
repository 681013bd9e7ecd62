// Elementwise / reduction kernels of the training step: Network::train_inner
// (alpha-tak/src/model/network.rs:59-97) with forward_training (net6.rs:111-122: BatchNorm on batch statistics,
// log_softmax policy, tanh value), loss = -sum(pi * logp)/B + sum((z - v)^2)/B, and Adam (lr 1e-4, wd 1e-4).
// The dense contractions (forward conv, dgrad, wgrad) are conv_tc3.cuh / wgrad_tc.cuh; everything here is HBM-bound
// passes over bf16 strip planes [chunk of 8 channels][slot][8] (conv_tc3.cuh).
//
// Invariant kept by every kernel: pad columns, tile remainders and boards >= n_boards hold ZERO in every activation and
// gradient plane -- the tap shifts of the conv / dgrad / wgrad kernels rely on it.
#pragma once
#include <cuda_bf16.h>

#include "conv_tc3.cuh"
#include "net_kernels.cuh"

namespace tb {

constexpr float TRAIN_BN_EPS = 1e-5f;       // tch nn::BatchNormConfig default
constexpr float TRAIN_BN_MOMENTUM = 0.1f;   // tch default

// is `slot` a real square of a board < n_boards?
template <int N>
__device__ __forceinline__ bool slot_valid(size_t slot, int n_boards) {
    using SM = SlotMap<N>;
    const int tile = int(slot >> 8), w = int(slot & 255);
    const int ry = w / SM::PITCH, rem = w - ry * SM::PITCH;
    const int bj = rem / SM::BW, rx = rem - bj * SM::BW;
    return ry < N && rx < N && tile * SM::BPT + bj < n_boards;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* vb = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(vb[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) vb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return v;
}

// ---- weights: fp32 master tensor [c_out][c_in][3][3] -> the bf16 operand image conv_tc3 streams ----------------------
// packed[(((slab*9 + tap)*2 + kc)*128 + col)*8 + j]:  out channel = col_base + col, in channel = in_base + slab*16 + kc*8 + j
//   forward image : W[out][in][tap]
//   dgrad image   : the transposed, 180-degree rotated filter  W[in][out][8 - tap]   (dX = conv(dY, W'))
static __global__ void k_pack_conv_train(const float* w, int c_out, int c_in, int col_base, int in_base, int dgrad,
                                         __nv_bfloat16* packed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= int(C3_W_LAYER_ELEMS)) return;
    const int j = idx & 7, col = (idx >> 3) & 127, kc = (idx >> 10) & 1, rest = idx >> 11;
    const int tap = rest % 9, slab = rest / 9;
    const int oc = col_base + col, ic = in_base + slab * 16 + kc * 8 + j;
    float v = 0.f;
    if (!dgrad) {
        if (oc < c_out && ic < c_in) v = w[(size_t(oc) * c_in + ic) * 9 + tap];
    } else {
        if (ic < c_out && oc < c_in) v = w[(size_t(ic) * c_in + oc) * 9 + (8 - tap)];
    }
    packed[idx] = __float2bfloat16(v);
}
// bias[128] of an output-channel group (zero beyond c_out)
static __global__ void k_pack_bias_train(const float* b, int c_out, int col_base, float* out) {
    const int i = threadIdx.x;
    out[i] = (b && col_base + i < c_out) ? b[col_base + i] : 0.f;
}

// ---- input: fp32 [B][C][N][N] (Example::to_tensors, example.rs:63-78) -> bf16 strip planes, zero elsewhere ----------
template <int N>
__global__ void __launch_bounds__(256) k_nchw_to_planes(const float* in, int n_boards, int C, __nv_bfloat16* planes, int S) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;    // (chunk, slot)
    if (idx >= size_t(16) * S) return;
    using SM = SlotMap<N>;
    const int chunk = int(idx / S);
    const size_t slot = idx % S;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (slot_valid<N>(slot, n_boards)) {
        const int tile = int(slot >> 8), w = int(slot & 255);
        const int ry = w / SM::PITCH, rem = w - ry * SM::PITCH;
        const int bj = rem / SM::BW, rx = rem - bj * SM::BW;
        const int b = tile * SM::BPT + bj;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = chunk * 8 + j;
            if (ch < C) f[j] = in[((size_t(b) * C + ch) * N + ry) * N + rx];
        }
    }
    *reinterpret_cast<uint4*>(planes + idx * 8) = pack8(f);
}

// ---- BatchNorm forward on batch statistics --------------------------------------------------------------------------
// z = relu(a*y + b (+ res)) on the real squares, 0 elsewhere  (net6.rs:72-76, res_block.rs:14-22 with train = true).
// The BatchNorm statistics are finalised here too (no separate launch): S is a multiple of 256, so a block lies inside one
// 8-channel chunk; its first 8 threads turn the conv epilogue's double sums into mean / rstd and the affine a, b, and the
// first block of every chunk also stores mean / rstd for backward and updates the running statistics (momentum 0.1,
// unbiased variance, as libtorch).
// Elementwise passes over the strip planes: grid = (ceil(S / (256 * EW_UNROLL)), 16 channel chunks); a thread handles
// EW_UNROLL slots 256 apart and issues all its 16-byte loads before it touches the first value (the one-element-per-thread
// version ran at 0.57-0.68 of the HBM peak: too few bytes in flight per SM).
constexpr int EW_UNROLL = 4;
inline dim3 ew_grid(int S) { return dim3(unsigned((S + 256 * EW_UNROLL - 1) / (256 * EW_UNROLL)), 16); }

template <int N>
__global__ void __launch_bounds__(256, 4) k_bn_apply(const __nv_bfloat16* y, const __nv_bfloat16* res, const double* sums,
                                                  double count, const float* gamma, const float* beta,
                                                  float* running_mean, float* running_var, float* mean_out,
                                                  float* rstd_out, int n_boards, int S, __nv_bfloat16* z) {
    const int chunk = blockIdx.y;
    __shared__ float s_a[8], s_b[8];
    if (threadIdx.x < 8) {
        const int c = chunk * 8 + threadIdx.x;
        const double mean = sums[c] / count;
        double var = sums[128 + c] / count - mean * mean;
        if (var < 0) var = 0;
        const double rstd = 1.0 / sqrt(var + double(TRAIN_BN_EPS));
        const double a = double(gamma[c]) * rstd;
        s_a[threadIdx.x] = float(a);
        s_b[threadIdx.x] = float(double(beta[c]) - mean * a);
        if (blockIdx.x == 0) {
            mean_out[c] = float(mean);
            rstd_out[c] = float(rstd);
            const double unbiased = count > 1 ? var * count / (count - 1) : var;
            running_mean[c] = float((1.0 - TRAIN_BN_MOMENTUM) * running_mean[c] + TRAIN_BN_MOMENTUM * mean);
            running_var[c] = float((1.0 - TRAIN_BN_MOMENTUM) * running_var[c] + TRAIN_BN_MOMENTUM * unbiased);
        }
    }
    const int s0 = blockIdx.x * (256 * EW_UNROLL) + threadIdx.x;
    uint4 yv[EW_UNROLL], rv[EW_UNROLL];
    bool ok[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        const int slot = s0 + 256 * u;
        ok[u] = slot < S && slot_valid<N>(size_t(slot), n_boards);
        yv[u] = rv[u] = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
            const size_t idx = size_t(chunk) * S + slot;
            yv[u] = *reinterpret_cast<const uint4*>(y + idx * 8);
            if (res) rv[u] = *reinterpret_cast<const uint4*>(res + idx * 8);
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        const int slot = s0 + 256 * u;
        if (slot >= S) continue;
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ok[u]) {
            float f[8], r[8];
            unpack8(yv[u], f);
            unpack8(rv[u], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = f[j] * s_a[j] + s_b[j];
                if (res) v += r[j];
                o[j] = fmaxf(v, 0.f);
            }
        }
        *reinterpret_cast<uint4*>(z + (size_t(chunk) * S + slot) * 8) = pack8(o);
    }
}

// ---- BatchNorm backward ---------------------------------------------------------------------------------------------
// g' = g * (zout > 0)  (ReLU mask; zout == nullptr: no mask);  xhat = (y - mean) * rstd
// pass 1: sums[c] += sum g', sums[128 + c] += sum g' * xhat   (gradient planes are zero off the real squares)
constexpr int BNR_SPLIT = 64;   // slot ranges per chunk
static __global__ void __launch_bounds__(256) k_bn_bwd_reduce(const __nv_bfloat16* g, const __nv_bfloat16* zout,
                                                              const __nv_bfloat16* y, const float* mean,
                                                              const float* rstd, int S, double* sums) {
    const int chunk = blockIdx.y;
    const int per = (S + BNR_SPLIT - 1) / BNR_SPLIT;
    const int s0 = blockIdx.x * per, s1 = min(S, s0 + per);
    float m[8], r[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        m[j] = mean[chunk * 8 + j];
        r[j] = rstd[chunk * 8 + j];
        a1[j] = a2[j] = 0.f;
    }
    for (int sb = s0 + threadIdx.x; sb < s1; sb += blockDim.x * EW_UNROLL) {
        uint4 gq[EW_UNROLL], yq[EW_UNROLL], zq[EW_UNROLL];
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {      // all loads of this round first
            const int sl = sb + u * blockDim.x;
            gq[u] = yq[u] = zq[u] = make_uint4(0, 0, 0, 0);
            if (sl < s1) {
                const size_t idx = size_t(chunk) * S + sl;
                gq[u] = *reinterpret_cast<const uint4*>(g + idx * 8);
                yq[u] = *reinterpret_cast<const uint4*>(y + idx * 8);
                if (zout) zq[u] = *reinterpret_cast<const uint4*>(zout + idx * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {
            if (sb + u * blockDim.x >= s1) continue;   // (a zero gradient would add m*r*0 = 0 anyway)
            float gv[8], yv[8], zv[8];
            unpack8(gq[u], gv);
            unpack8(yq[u], yv);
            unpack8(zq[u], zv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float gg = (!zout || zv[j] > 0.f) ? gv[j] : 0.f;
                a1[j] += gg;
                a2[j] += gg * ((yv[j] - m[j]) * r[j]);
            }
        }
    }
    __shared__ float sh[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o);
            a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], o);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sh[warp][j] = a1[j];
            sh[warp][8 + j] = a2[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        const int j = threadIdx.x & 7;
        atomicAdd(&sums[(threadIdx.x < 8 ? 0 : 128) + chunk * 8 + j], double(t));
    }
}
// pass 2: dy = gamma * rstd * (g' - c1 - xhat * c2) on the real squares; optionally also stores g' (the gradient that
// flows into the residual connection, res_block.rs:21)
// (the sums of pass 1 are finalised here: c1 = sum g' / n, c2 = sum g'*xhat / n per block, and the first block of every
// chunk accumulates dgamma += sum g'*xhat, dbeta += sum g')
// MASK = false: `g` is already the masked gradient g' (it came out of the dgrad epilogue with the sums, conv_tc3.cuh
// ConvParams::bnb_y): two loads and one store per slot, half the registers -- 4 blocks per SM instead of 2 (124
// registers with the mask operand: the pass ran at 3.6 TB/s against 4.9 for k_bn_apply on the same bytes).
template <int N, bool MASK>
__global__ void __launch_bounds__(256, MASK ? 2 : 4) k_bn_bwd_apply(const __nv_bfloat16* g, const __nv_bfloat16* zout,
                                                      const __nv_bfloat16* y, const float* mean, const float* rstd,
                                                      const float* gamma, const double* sums, double count,
                                                      float* grad_gamma, float* grad_beta,
                                                      int n_boards, int S, __nv_bfloat16* dy, __nv_bfloat16* gmasked) {
    const int chunk = blockIdx.y;
    __shared__ float c1[8], c2[8], s_m[8], s_r[8], s_gr[8];
    if (threadIdx.x < 8) {
        const int c = chunk * 8 + threadIdx.x;
        c1[threadIdx.x] = float(sums[c] / count);
        c2[threadIdx.x] = float(sums[128 + c] / count);
        s_m[threadIdx.x] = mean[c];
        s_r[threadIdx.x] = rstd[c];
        s_gr[threadIdx.x] = gamma[c] * rstd[c];
        if (blockIdx.x == 0) {
            grad_beta[c] += float(sums[c]);
            grad_gamma[c] += float(sums[128 + c]);
        }
    }
    const int s0 = blockIdx.x * (256 * EW_UNROLL) + threadIdx.x;
    uint4 gq[EW_UNROLL], yq[EW_UNROLL], zq[MASK ? EW_UNROLL : 1];
    bool ok[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        const int slot = s0 + 256 * u;
        ok[u] = slot < S && slot_valid<N>(size_t(slot), n_boards);
        gq[u] = yq[u] = make_uint4(0, 0, 0, 0);
        if (MASK) zq[u] = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
            const size_t idx = size_t(chunk) * S + slot;
            gq[u] = *reinterpret_cast<const uint4*>(g + idx * 8);
            yq[u] = *reinterpret_cast<const uint4*>(y + idx * 8);
            if (MASK) zq[u] = *reinterpret_cast<const uint4*>(zout + idx * 8);
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        const int slot = s0 + 256 * u;
        if (slot >= S) continue;
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ok[u]) {
            float gv[8], yv[8], zv[8];
            unpack8(gq[u], gv);
            unpack8(yq[u], yv);
            if (MASK) unpack8(zq[u], zv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float gg = (!MASK || zv[j] > 0.f) ? gv[j] : 0.f;
                const float xhat = (yv[j] - s_m[j]) * s_r[j];
                gm[j] = gg;
                o[j] = s_gr[j] * (gg - c1[j] - xhat * c2[j]);
            }
        }
        const size_t idx = size_t(chunk) * S + slot;
        *reinterpret_cast<uint4*>(dy + idx * 8) = pack8(o);
        if (MASK && gmasked) *reinterpret_cast<uint4*>(gmasked + idx * 8) = pack8(gm);
    }
}

// ---- heads ----------------------------------------------------------------------------------------------------------
// value head forward on fp32 master weights: v = tanh(fc(flatten NCHW)) (net6.rs:117-121)
template <int N>
__global__ void __launch_bounds__(256) k_value_train(const __nv_bfloat16* act, int S, const float* wv, const float* bv,
                                                     int n_boards, float* out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_boards) return;
    constexpr int NSQ = N * N;
    const int l = threadIdx.x & 31;
    float acc = 0.f;
    for (int pos = l; pos < NSQ; pos += 32) {
        const size_t slot = SlotMap<N>::slot(w, pos / N, pos % N);
        for (int chunk = 0; chunk < 16; ++chunk) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(act + (size_t(chunk) * S + slot) * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += f[j] * wv[(chunk * 8 + j) * NSQ + pos];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (l == 0) out[w] = tanhf(acc + bv[0]);
}

// policy loss and its gradient (network.rs:79-80 with log_softmax, net6.rs:113-116):
//   loss_p += -sum_i pi_i * logp_i / B ;  dlogit_i = (softmax_i * sum(pi) - pi_i) / B   -> bf16 strip planes [32 chunks]
// one block per board; logits fp32 [256][S], stats = {max, sum exp(l - max)} per board, pi [B][n_ch * N*N]
template <int N>
__global__ void __launch_bounds__(256) k_policy_loss_grad(const float* logits, int S, int n_ch, const float2* stats,
                                                          const float* pi, int n_boards, __nv_bfloat16* dlogits,
                                                          double* loss /*[0] += loss_p*/) {
    const int b = blockIdx.x;
    constexpr int NSQ = N * N;
    const float2 st = stats[b];
    const float log_sum = logf(st.y);
    const float* p = pi + size_t(b) * n_ch * NSQ;
    __shared__ float sh[8];
    __shared__ float s_sum_pi;
    float sum_pi = 0.f;
    for (int i = threadIdx.x; i < n_ch * NSQ; i += blockDim.x) sum_pi += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum_pi += __shfl_xor_sync(FULL, sum_pi, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sum_pi;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w];
        s_sum_pi = t;
    }
    __syncthreads();
    sum_pi = s_sum_pi;
    const float inv_b = 1.0f / float(n_boards);
    float nll = 0.f;
    for (int item = threadIdx.x; item < 32 * NSQ; item += blockDim.x) {     // (chunk, square)
        const int chunk = item / NSQ, sq = item % NSQ;
        const size_t slot = SlotMap<N>::slot(b, sq / N, sq % N);
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = chunk * 8 + j;
            d[j] = 0.f;
            if (ch < n_ch) {
                const float lg = logits[size_t(ch) * S + slot];
                const float logp = lg - st.x - log_sum;
                const float t = p[ch * NSQ + sq];
                nll -= t * logp;
                d[j] = (expf(logp) * sum_pi - t) * inv_b;
            }
        }
        *reinterpret_cast<uint4*>(dlogits + (size_t(chunk) * S + slot) * 8) = pack8(d);
    }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nll += __shfl_xor_sync(FULL, nll, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = nll;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w];
        atomicAdd(&loss[0], double(t) * double(inv_b));
    }
}

// per-channel sum over all slots of a stack of planes: out[chunk*8 + j] += sum_s planes[chunk][s][j]  (policy conv bias
// gradient: sum of dlogits).  grid (BNR_SPLIT, chunks)
static __global__ void __launch_bounds__(256) k_planes_colsum(const __nv_bfloat16* planes, int S, int n_valid, float* out) {
    const int chunk = blockIdx.y;
    const int per = (S + BNR_SPLIT - 1) / BNR_SPLIT;
    const int s0 = blockIdx.x * per, s1 = min(S, s0 + per);
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(planes + (size_t(chunk) * S + s) * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
    __shared__ float sh[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[warp][j] = a[j];
    }
    __syncthreads();
    if (threadIdx.x < 8 && chunk * 8 + threadIdx.x < n_valid) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        atomicAdd(&out[chunk * 8 + threadIdx.x], t);
    }
}

// value loss (network.rs:81): loss_z += sum (z - v)^2 / B ; dpre[b] = dL/d(fc output) = -2 (z - v) / B * (1 - v^2);
// grad of the fc bias += sum dpre
static __global__ void k_value_loss_grad(const float* v, const float* z, int n_boards, float* dpre, double* loss /*[1]*/,
                                         float* grad_bias) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f, d = 0.f;
    if (b < n_boards) {
        const float inv_b = 1.0f / float(n_boards);
        const float diff = z[b] - v[b];
        l = diff * diff * inv_b;
        d = -2.0f * diff * inv_b * (1.0f - v[b] * v[b]);
        dpre[b] = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l += __shfl_xor_sync(0xffffffffu, l, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&loss[1], double(l));
        atomicAdd(grad_bias, d);
    }
}
// gradient of the value head w.r.t. the trunk output: g[c][slot(b,sq)] = dpre[b] * Wv[c*NSQ + sq], zero elsewhere
template <int N>
__global__ void __launch_bounds__(256) k_value_bwd_trunk(const float* dpre, const float* wv, int n_boards, int S,
                                                         __nv_bfloat16* g) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(16) * S) return;
    using SM = SlotMap<N>;
    const int chunk = int(idx / S);
    const size_t slot = idx % S;
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (slot_valid<N>(slot, n_boards)) {
        const int tile = int(slot >> 8), w = int(slot & 255);
        const int ry = w / SM::PITCH, rem = w - ry * SM::PITCH;
        const int bj = rem / SM::BW, rx = rem - bj * SM::BW;
        const float d = dpre[tile * SM::BPT + bj];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = d * wv[(chunk * 8 + j) * (N * N) + ry * N + rx];
    }
    *reinterpret_cast<uint4*>(g + idx * 8) = pack8(o);
}
// grad Wv[c*NSQ + sq] += sum_b dpre[b] * s[b][c][sq].  grid (16 chunks * NSQ, board splits)
template <int N>
__global__ void __launch_bounds__(128) k_value_wgrad(const float* dpre, const __nv_bfloat16* act, int n_boards, int S,
                                                     float* grad_wv) {
    constexpr int NSQ = N * N;
    const int chunk = blockIdx.x / NSQ, sq = blockIdx.x % NSQ;
    const int per = (n_boards + gridDim.y - 1) / gridDim.y;
    const int b0 = blockIdx.y * per, b1 = min(n_boards, b0 + per);
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(act + (size_t(chunk) * S + SlotMap<N>::slot(b, sq / N, sq % N)) * 8), f);
        const float d = dpre[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += d * f[j];
    }
    __shared__ float sh[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[warp][j] = a[j];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
        for (int w = 0; w < 4; ++w) t += sh[w][threadIdx.x];
        atomicAdd(&grad_wv[(chunk * 8 + threadIdx.x) * NSQ + sq], t);
    }
}

// ---- Adam (tch nn::Adam {beta1 0.9, beta2 0.999, wd} over libtorch's optim::Adam: L2 weight decay folded into the
// gradient, bias-corrected moments, eps 1e-8 added to sqrt(v_hat)) ---------------------------------------------------------
static __global__ void k_adam(float* w, const float* grad, float* m, float* v, int count, float lr, float wd,
                              float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float g = grad[i] + wd * w[i];
    const float mi = beta1 * m[i] + (1.0f - beta1) * g;
    const float vi = beta2 * v[i] + (1.0f - beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    w[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace tb

// ---- Net5: fully connected policy head (net5.rs:56-62,106-108) in training ----------------------------------------------
// All three GEMMs of the head (forward, dgrad, wgrad) run on fc_tc_kernel (fc_tc.cuh), which computes
//   out[n][m] = bias[m] + sum over (p, c < 128) of  A[m][p*128 + c] * X[p][c][n]
// from the packed images  A -> Wp[m / 128][p][c / 16][(c / 8) % 2][m % 128][c % 8]   and   X -> [p][c / 8][n_pad][c % 8]:
//   forward : m = output j,          n = board,            (p, c) = (board position, trunk channel)
//   dgrad   : m = k = c*NSQ + pos,   n = board,            p*128 + c = j          (A = W^T, X = dlogits)
//   wgrad   : m = output j,          n = k = c*NSQ + pos,  p*128 + c = board      (A = dlogits^T, X = trunk output^T)
namespace tb {

__device__ __forceinline__ size_t fc_a_index(int m, int kk, int P) {      // element (m, kk = p*128 + c) of an A image
    const int mt = m >> 7, col = m & 127, p = kk >> 7, c = kk & 127;
    return ((((size_t(mt) * P + p) * 8 + (c >> 4)) * 2 + ((c >> 3) & 1)) * 128 + col) * 8 + (c & 7);
}
__device__ __forceinline__ size_t fc_x_index(int n, int kk, int n_pad) {  // element (kk, n) of an X image
    const int p = kk >> 7, c = kk & 127;
    return fc_x_offset(p, c >> 3, n, n_pad) * 8 + (c & 7);
}

// forward / dgrad operand images of the fp32 master W[J][K] (K = 128*NSQ, column k = c*NSQ + pos)
static __global__ void k_fc_pack_fwd(const float* w, int J, int NSQ, int m_tiles, __nv_bfloat16* wp) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;     // (m = j, pos, c)
    const size_t total = size_t(m_tiles) * 128 * NSQ * 128;
    if (idx >= total) return;
    const int c = int(idx % 128), pos = int((idx / 128) % NSQ), j = int(idx / (size_t(128) * NSQ));
    const float v = j < J ? w[size_t(j) * 128 * NSQ + size_t(c) * NSQ + pos] : 0.f;
    wp[fc_a_index(j, pos * 128 + c, NSQ)] = __float2bfloat16(v);
}
static __global__ void k_fc_pack_dgrad(const float* w, int J, int K, int P /*ceil(J/128)*/, __nv_bfloat16* wpt) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;     // (m = k, kk = j)
    const size_t total = size_t(K) * P * 128;
    if (idx >= total) return;
    const int j = int(idx % (size_t(P) * 128)), k = int(idx / (size_t(P) * 128));
    const float v = j < J ? w[size_t(j) * K + k] : 0.f;
    wpt[fc_a_index(k, j, P)] = __float2bfloat16(v);
}

// policy loss and gradient over dense logits [B][J] (log_softmax, network.rs:79): loss_p += -sum pi*logp / B;
// dlogit = (softmax * sum(pi) - pi) / B written as bf16 in BOTH operand layouts: dl_x (X image, contraction over j: dgrad)
// and dl_a (A image, contraction over boards: wgrad).  One block per board; both images are zeroed by the caller.
static __global__ void __launch_bounds__(256) k_fc_loss_grad(const float* logits, int J, const float2* stats,
                                                             const float* pi, int n_boards, int b_pad,
                                                             __nv_bfloat16* dl_x, __nv_bfloat16* dl_a, double* loss) {
    const int b = blockIdx.x;
    const float2 st = stats[b];
    const float log_sum = logf(st.y);
    const float* row = logits + size_t(b) * J;
    const float* p = pi + size_t(b) * J;
    __shared__ float sh[8];
    __shared__ float s_sum_pi;
    float sum_pi = 0.f;
    for (int j = threadIdx.x; j < J; j += blockDim.x) sum_pi += p[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum_pi += __shfl_xor_sync(0xffffffffu, sum_pi, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sum_pi;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w];
        s_sum_pi = t;
    }
    __syncthreads();
    sum_pi = s_sum_pi;
    const float inv_b = 1.0f / float(n_boards);
    const int P = (J + 127) / 128, PB = b_pad / 128;
    float nll = 0.f;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const float logp = row[j] - st.x - log_sum;
        nll -= p[j] * logp;
        const __nv_bfloat16 d = __float2bfloat16((expf(logp) * sum_pi - p[j]) * inv_b);
        dl_x[fc_x_index(b, j, b_pad)] = d;
        dl_a[fc_a_index(j, b, PB)] = d;
    }
    (void)P;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nll += __shfl_xor_sync(0xffffffffu, nll, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = nll;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w];
        atomicAdd(&loss[0], double(t) * double(inv_b));
    }
}

// trunk gradient = policy-head dgrad (ds[b][c*NSQ + pos], fp32 from the GEMM) + value-head gradient, as strip planes
template <int N>
__global__ void __launch_bounds__(256) k_fc_ds_to_planes(const float* ds, const float* dpre, const float* wv,
                                                         int n_boards, int S, __nv_bfloat16* g) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(16) * S) return;
    using SM = SlotMap<N>;
    constexpr int NSQ = N * N;
    const int chunk = int(idx / S);
    const size_t slot = idx % S;
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (slot_valid<N>(slot, n_boards)) {
        const int tile = int(slot >> 8), w = int(slot & 255);
        const int ry = w / SM::PITCH, rem = w - ry * SM::PITCH;
        const int bj = rem / SM::BW, rx = rem - bj * SM::BW;
        const int b = tile * SM::BPT + bj, pos = ry * N + rx;
        const float d = dpre[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = (chunk * 8 + j) * NSQ + pos;
            o[j] = ds[size_t(b) * (128 * NSQ) + k] + d * wv[k];
        }
    }
    *reinterpret_cast<uint4*>(g + idx * 8) = pack8(o);
}

// X image of the transposed trunk output for the FC wgrad: element (kk = board, n = k = c*NSQ + pos)
template <int N>
__global__ void __launch_bounds__(256) k_fc_repack_wgrad(const __nv_bfloat16* act, int S, int n_boards, int b_pad,
                                                         int k_pad, __nv_bfloat16* x) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;     // (board group of 8, k)
    constexpr int NSQ = N * N, K = 128 * NSQ;
    if (idx >= size_t(b_pad / 8) * k_pad) return;
    const int k = int(idx % k_pad), bg = int(idx / k_pad);
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (k < K) {
        const int c = k / NSQ, pos = k % NSQ;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int b = bg * 8 + i;
            if (b < n_boards)
                f[i] = __bfloat162float(act[(size_t(c >> 3) * S + SlotMap<N>::slot(b, pos / N, pos % N)) * 8 + (c & 7)]);
        }
    }
    // boards bg*8 .. bg*8+7 = contraction index kk: p = kk / 128, chunk = (kk % 128) / 8
    *reinterpret_cast<uint4*>(x + fc_x_index(k, bg * 8, k_pad)) = pack8(f);
}

// grad W[j][k] += dWt[k][j]  (dWt = GEMM output [K][J] fp32)
static __global__ void k_fc_wgrad_add(const float* dwt, int J, int K, float* grad_w) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(J) * K) return;
    const int j = int(idx % J), k = int(idx / J);
    grad_w[size_t(j) * K + k] += dwt[idx];
}

}  // namespace tb
