// 3x3 "same" convolution of the AlphaTak residual tower as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces, per layer, the cuDNN conv + BatchNorm(eval) + ReLU (+ residual add) library calls the
// reference issues through tch-rs (reference: alpha-tak/src/model/net6.rs:70-78, res_block.rs:13-23,
// policy head net6.rs:99-103).  BatchNorm is folded into the weights/bias on the host.
//
// Data layout in HBM ("slot planes"):
//   activations  act[chunk c = 0..15][slot s = 0..S-1][8 channels]   bf16   (channel = 8*c + j)
//   slot(board b, row y, col x) = GUARD + b*P*P + (y+1)*P + x,  P = N+1 (one zero pad column per row,
//   one zero pad row per board; the pad row of board b+1 doubles as the bottom pad of board b).
//   A tap (ky,kx) of the 3x3 stencil is then a constant slot shift (ky-1)*P + (kx-1): the A operand of
//   every tap is the SAME shared-memory tile addressed with a shifted start address -- no im2col copy.
//   weights  w[stage = tap*2 + khalf][kchunk 0..7][c_out 0..127][8 c_in]  bf16 (16 KiB per stage),
//   exactly the K-major no-swizzle UMMA operand image, so one 1-D bulk copy stages it.
//
// GEMM view per CTA tile: D[256 slots x 128 c_out] += A[256 x 1152] * W[128 x 1152]^T as two
// M=128,N=128 accumulators in TMEM (fp32), 9 taps x 8 K-steps of K=16 each.
//
// Warp roles (352 threads, 1 persistent CTA per SM):
//   warp 0     bulk-copy producer of the activation tile (+ halo)             -> mbarrier tx
//   warp 10    bulk-copy producer of the weight stages                         -> mbarrier tx
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer               -> tcgen05.commit
//   warps 2-9  epilogue: residual prefetch, tcgen05.ld (software pipelined) -> +bias (+residual) -> ReLU ->
//              pad mask -> bf16 slot planes
// Double-buffered activation tiles and TMEM accumulators let the epilogue of tile i overlap the MMAs of
// tile i+1; weights stream through a 5-stage ring (they stay L2 resident: 288 KiB per layer).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "ptx_sm100.cuh"

namespace tb {

constexpr int CONV_GUARD = 16;        // zero slots before the first / after the last tile
constexpr int CONV_TILE_M = 256;      // slots per CTA tile (2 x UMMA M=128)
constexpr int CONV_HALO = 8;          // max |tap shift| = P+1 <= 8 for N <= 6  (N=7,8 would need 9/10 -> see static check)
constexpr int CONV_ROWS = CONV_TILE_M + 2 * 16;  // rows staged per tile (halo of 16 covers P <= 15)
constexpr int CONV_HALO_ROWS = 16;
constexpr int CONV_CHUNKS = 16;       // 128 input channels / 8
constexpr int CONV_A_BYTES = CONV_ROWS * 16 * CONV_CHUNKS;  // 73728
constexpr int CONV_W_STAGE_BYTES = 8 * 128 * 16;            // 16384: 64 c_in x 128 c_out
constexpr int CONV_W_STAGES = 4;
constexpr int CONV_STAGES_PER_LAYER = 18;                   // 9 taps x 2 K halves
constexpr int CONV_THREADS = 352;   // 11 warps: A producer, MMA, 8 epilogue, W producer
constexpr int CONV_SMEM_BYTES = 2 * CONV_A_BYTES + CONV_W_STAGES * CONV_W_STAGE_BYTES + 1024;

enum ConvMode : int {
    CONV_RELU = 0,        // out = relu(conv + bias)                     -> bf16 slot planes
    CONV_RES_RELU = 1,    // out = relu(conv + bias + res)               -> bf16 slot planes
    CONV_LOGITS_F32 = 2,  // out = conv + bias  (no activation)          -> fp32 [channel][slot]
};

struct ConvParams {
    const __nv_bfloat16* in;    // slot planes [16][S][8]
    const __nv_bfloat16* res;   // slot planes (mode 1) or nullptr
    __nv_bfloat16* out;         // slot planes (modes 0/1)
    float* out_f32;             // [out_ch_total][S] (mode 2)
    const __nv_bfloat16* w;     // [18][8][128][8]
    const float* bias;          // [128]
    int S;                      // total slots incl. guards
    int tiles;                  // number of 256-slot tiles
    int n_boards;
    int pitch;                  // P = N+1
    int mode;
    int out_ch_offset;          // mode 2: first output channel of this 128-wide group
    int out_ch_valid;           // mode 2: number of real channels in this group (<=128)
};

__device__ __forceinline__ bool conv_slot_valid(int slot, int pitch, int n_boards) {
    int rel = slot - CONV_GUARD;
    int spb = pitch * pitch;
    int board = rel / spb;
    int loc = rel - board * spb;
    int y = loc / pitch;
    int x = loc - y * pitch;
    return rel >= 0 && board < n_boards && y >= 1 && x < pitch - 1;
}

template <int MODE>
static __global__ void __launch_bounds__(CONV_THREADS, 1) conv3x3_tc_kernel(const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_buf = smem;                                        // 2 x CONV_A_BYTES
    uint8_t* w_buf = smem + 2 * CONV_A_BYTES;                     // CONV_W_STAGES x 16 KiB
    uint8_t* tail = w_buf + CONV_W_STAGES * CONV_W_STAGE_BYTES;   // barriers etc.
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    // barrier indices
    //  0,1   a_full[2]      2,3   a_empty[2]
    //  4..7  w_full[4]      8..11 w_empty[4]
    //  12,13 acc_full[2]    14,15 acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 16 * 8);
    float* s_bias = reinterpret_cast<float*>(tail + 16 * 8 + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(0 + i), 1);
            mbar_init(BAR(2 + i), 1);
            mbar_init(BAR(12 + i), 1);
            mbar_init(BAR(14 + i), 8);  // one arrive per epilogue warp
        }
        for (int i = 0; i < CONV_W_STAGES; ++i) {
            mbar_init(BAR(4 + i), 1);
            mbar_init(BAR(8 + i), 1);
        }
        mbar_fence_init();
    }
    if (threadIdx.x < 128) s_bias[threadIdx.x] = p.bias[threadIdx.x];
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const size_t plane_bytes = static_cast<size_t>(p.S) * 16;

    if (warp == 0) {
        // ===================== activation-tile producer =====================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                const uint32_t aph = (it >> 1) & 1;
#if defined(CONV_EXP) && (CONV_EXP & 4)
                if (it >= 2) break;
#endif
                mbar_wait(BAR(2 + ab), aph ^ 1);
                mbar_expect_tx(BAR(0 + ab), CONV_A_BYTES);
                const int row0 = CONV_GUARD + tile * CONV_TILE_M - CONV_HALO_ROWS;  // >= 0
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.in) + static_cast<size_t>(row0) * 16;
                const uint32_t dst = smem_u32(a_buf + ab * CONV_A_BYTES);
                for (int c = 0; c < CONV_CHUNKS; ++c)
                    bulk_g2s(dst + c * (CONV_ROWS * 16), src + c * plane_bytes, CONV_ROWS * 16, BAR(0 + ab));
            }
        }
    } else if (warp == 10) {
        // ===================== weight-stage producer =====================
        if (lane == 0) {
            uint32_t wcount = 0;
#if defined(CONV_EXP) && (CONV_EXP & 1)
            if (false)
#endif
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                for (int st = 0; st < CONV_STAGES_PER_LAYER; ++st, ++wcount) {
                    const int s = wcount % CONV_W_STAGES;
                    const uint32_t ph = (wcount / CONV_W_STAGES) & 1;
                    mbar_wait(BAR(8 + s), ph ^ 1);
                    mbar_expect_tx(BAR(4 + s), CONV_W_STAGE_BYTES);
                    bulk_g2s(smem_u32(w_buf + s * CONV_W_STAGE_BYTES),
                             reinterpret_cast<const uint8_t*>(p.w) + static_cast<size_t>(st) * CONV_W_STAGE_BYTES,
                             CONV_W_STAGE_BYTES, BAR(4 + s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
#if defined(CONV_EXP) && (CONV_EXP & 128)
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 256);  // experiment: issue-rate probe (garbage B)
#else
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 128);
#endif
            uint32_t wcount = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                const uint32_t ph2 = (it >> 1) & 1;
                mbar_wait(BAR(14 + ab), ph2 ^ 1);  // accumulator stage drained by the epilogue
#if defined(CONV_EXP) && (CONV_EXP & 4)
                if (it < 2)
#endif
                mbar_wait(BAR(0 + ab), ph2);       // activation tile landed
                tc_fence_after();
                const uint32_t a_base = smem_u32(a_buf + ab * CONV_A_BYTES);
                const uint32_t d_base = tmem_base + ab * 256;
                for (int st = 0; st < CONV_STAGES_PER_LAYER; ++st, ++wcount) {
                    const int s = wcount % CONV_W_STAGES;
                    const uint32_t ph = (wcount / CONV_W_STAGES) & 1;
#if !(defined(CONV_EXP) && (CONV_EXP & 1))
                    mbar_wait(BAR(4 + s), ph);
#endif
                    tc_fence_after();
                    const int tap = st >> 1, half = st & 1;
                    const int shift = (tap / 3 - 1) * p.pitch + (tap % 3 - 1);
                    const uint32_t w_base = smem_u32(w_buf + s * CONV_W_STAGE_BYTES);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int chunk = half * 8 + j * 2;
                            const uint32_t a_addr =
                                a_base + chunk * (CONV_ROWS * 16) + (CONV_HALO_ROWS + t * 128 + shift) * 16;
                            const uint64_t adesc = umma_desc_kmajor_noswz(a_addr, CONV_ROWS * 16, 128);
                            const uint64_t bdesc = umma_desc_kmajor_noswz(w_base + j * 2 * (128 * 16), 128 * 16, 128);
#if defined(CONV_EXP) && (CONV_EXP & 128)
                            const uint64_t bprobe = umma_desc_kmajor_noswz(smem_u32(a_buf) + j * 8192, 256 * 16, 128);
                            umma_bf16(d_base, adesc, bprobe, idesc, (st | j) != 0);
#else
                            umma_bf16(d_base + t * 128, adesc, bdesc, idesc, (st | j) != 0);
#endif
                        }
                    }
#if !(defined(CONV_EXP) && (CONV_EXP & 1))
                    umma_commit(BAR(8 + s));  // weight stage free once these MMAs retire
#endif
                }
                umma_commit(BAR(12 + ab));  // accumulators ready
                umma_commit(BAR(2 + ab));   // activation tile free
            }
        }
    } else if (warp >= 2 && warp < 10) {
        // ===================== epilogue (8 warps, one thread per slot row) =====================
        const int ew = warp - 2;          // 0..7
        const int t = ew >> 2;            // accumulator tile 0/1
        const int quarter = warp & 3;     // TMEM lane quarter this warp may access
        const int row = t * 128 + quarter * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const uint32_t ph2 = (it >> 1) & 1;
            const int slot = CONV_GUARD + tile * CONV_TILE_M + row;
            const bool valid = conv_slot_valid(slot, p.pitch, p.n_boards);
            // the residual does not depend on the MMAs: fetch the whole row before waiting for the accumulator
            uint4 res[16];
            if (MODE == CONV_RES_RELU) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    res[c] = make_uint4(0, 0, 0, 0);
                    if (valid) res[c] = __ldg(reinterpret_cast<const uint4*>(p.res + (static_cast<size_t>(c) * p.S + slot) * 8));
                }
            }
            mbar_wait(BAR(12 + ab), ph2);
            tc_fence_after();
#if defined(CONV_EXP) && (CONV_EXP & 2)
            if (p.S != -12345) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(BAR(14 + ab)); continue; }
#endif
            const uint32_t taddr = tmem_base + ab * 256 + t * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
            uint32_t r[2][32];
            tmem_ld32(taddr, r[0]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                tmem_ld_wait();
                if (cc < 3) tmem_ld32(taddr + (cc + 1) * 32, r[(cc + 1) & 1]);  // next chunk in flight while this one is processed
                const uint32_t(&rc)[32] = r[cc & 1];
                if (MODE == CONV_LOGITS_F32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int ch = cc * 32 + j;
                        if (ch < p.out_ch_valid) {
                            float v = __uint_as_float(rc[j]) + s_bias[ch];
                            p.out_f32[static_cast<size_t>(p.out_ch_offset + ch) * p.S + slot] = valid ? v : 0.0f;
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {  // 4 chunks of 8 channels
                        const int chunk = cc * 4 + q;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(rc[q * 8 + j]) + s_bias[chunk * 8 + j];
                        if (MODE == CONV_RES_RELU) {
                            const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&res[chunk]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 f = __bfloat1622float2(rb[j]);
                                v[2 * j] += f.x;
                                v[2 * j + 1] += f.y;
                            }
                        }
                        uint4 ov;
                        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float a = valid ? fmaxf(v[2 * j], 0.0f) : 0.0f;
                            float b = valid ? fmaxf(v[2 * j + 1], 0.0f) : 0.0f;
                            ob[j] = __floats2bfloat162_rn(a, b);
                        }
                        *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(chunk) * p.S + slot) * 8) = ov;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(14 + ab));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Host-side launch. `stream` is the engine's stream; `num_sms` from the device properties.
inline cudaError_t conv3x3_tc_launch(const ConvParams& p, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<CONV_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             CONV_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv3x3_tc_kernel<CONV_RES_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     CONV_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv3x3_tc_kernel<CONV_LOGITS_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     CONV_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int grid = p.tiles < num_sms ? p.tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    if (p.mode == CONV_RELU) conv3x3_tc_kernel<CONV_RELU><<<grid, CONV_THREADS, CONV_SMEM_BYTES, stream>>>(p);
    else if (p.mode == CONV_RES_RELU) conv3x3_tc_kernel<CONV_RES_RELU><<<grid, CONV_THREADS, CONV_SMEM_BYTES, stream>>>(p);
    else conv3x3_tc_kernel<CONV_LOGITS_F32><<<grid, CONV_THREADS, CONV_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tb
