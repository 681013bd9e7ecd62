// Weight gradient of a 3x3 "same" convolution over strip planes as a tcgen05 GEMM (sm_100a) -- the wgrad half of
// Network::train's backward pass (alpha-tak/src/model/network.rs:84, autograd of net6.rs:70-78 / res_block.rs:13-23;
// the reference gets it from libtorch/cuDNN).
//
//   dW[co][ci][ky][kx] = sum over slots s of  dY[s][co] * X[s + (ky-1)*PITCH + (kx-1)][ci]
//
// Both operands live in HBM as strip planes [chunk of 8 channels][slot][8 ch] bf16 (conv_tc3.cuh): pad columns, tile
// remainders and boards beyond n_boards hold zeros in BOTH tensors, so the tap shift over a 256-slot tile with zero
// halo rows is exact, as in the forward kernel.  The contraction index is the SLOT, i.e. both GEMM operands are
// "MN-major" (8 channels contiguous in 16 B, consecutive slots 16 B apart): a core matrix is 8 slots x 8 channels =
// 128 contiguous bytes, LBO (next 8 slots) = 128 B, SBO (next 8 channels) = the chunk plane stride in shared memory.
// As in conv_tc3 the tap shift of the B operand is just `start_address += shift * 16`.
//
// GEMM per tile and tap: D_tap[co 128 x ci 32] += dY^T[128 x 16 slots] * X_shifted[16 slots x 32], 16 K-steps per tile.
// A CTA owns one QUARTER of the input channels (32 = 4 chunks) and every 9 taps: 9 x 32 = 288 TMEM columns of fp32
// accumulators that stay resident while the CTA walks all its tiles (split-K over tiles across the CTAs of a quarter);
// at the end each CTA writes its partial [128][9][32] to scratch and k_wgrad_reduce sums the partials into the fp32
// gradient tensor [co][ci][3][3] (the weight blob's own layout).  One pipeline stage = dY tile (all 128 channels,
// 64 KiB) + X tile of the quarter with halos (4 x 368 rows x 16 B = 23 KiB); 2 stages.
// An M=128,N=32,K=16 SS instruction reads 4 KiB of A per 32 math cycles -- shared-memory operand bandwidth, not the
// tensor pipe, bounds this kernel at about half the forward kernel's rate (same FLOPs): see DESIGN.md.
//
// Warp roles (192 threads): warp 0 producer (cp.async.bulk), warp 1 TMEM alloc + MMA issuer, warps 2-5 epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "conv_tc3.cuh"
#include "ptx_sm100.cuh"

namespace tb {

constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 2;
constexpr int WG_QCH = 4;                                   // channel chunks (of 8) per CTA = 32 input channels
constexpr int WG_A_BYTES = 16 * C3_TILE_M * 16;             // 65536: dY tile, 16 chunks x 256 slots x 16 B
constexpr int WG_B_PLANE = C3_ROWS * 16;                    // 5888: one chunk plane of X with zero halos
constexpr int WG_B_BYTES = WG_QCH * WG_B_PLANE;             // 23552
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;     // 89088
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024;
constexpr int WG_COLS = 9 * 32;                             // accumulator columns in use (512 allocated)
constexpr int WG_PART_ELEMS = 128 * WG_COLS;                // fp32 per CTA partial: [co 128][tap 9][ci 32]

struct WgradParams {
    const __nv_bfloat16* dy;   // strip planes, first of 16 chunks (128 output channels), plane stride S
    const __nv_bfloat16* x;    // strip planes, chunk 0 (input channels), plane stride S
    float* scratch;            // [grid][128][9][32]
    int S;
    int tiles;
    int pitch;
    int quarters;              // ceil(c_in / 32): CTA c works on quarter c % quarters
};

// kind::f16 instruction descriptor with BOTH operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_mn(uint32_t M, uint32_t N) {
    return umma_idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
}

static __global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tail = smem + WG_STAGES * WG_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);       // full[2], empty[2], acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto FULL_ = [&](int i) { return bar0 + 8u * i; };
    auto EMPTY_ = [&](int i) { return bar0 + 8u * (WG_STAGES + i); };
    const uint32_t ACC = bar0 + 8u * (2 * WG_STAGES);

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(FULL_(i), 1);
            mbar_init(EMPTY_(i), 1);
        }
        mbar_init(ACC, 1);
        mbar_fence_init();
    }
    // zero halos of every X chunk plane (never written by the bulk copies)
    for (int i = threadIdx.x; i < WG_STAGES * WG_QCH * 2 * C3_HALO; i += WG_THREADS) {
        const int row = i % C3_HALO, side = (i / C3_HALO) & 1, plane = i / (2 * C3_HALO);   // plane = stage*4 + chunk
        uint8_t* dst = smem + (plane / WG_QCH) * WG_STAGE_BYTES + WG_A_BYTES + (plane % WG_QCH) * WG_B_PLANE +
                       (side ? (C3_HALO + C3_TILE_M + row) : row) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int q = int(blockIdx.x) % p.quarters;
    const int r = int(blockIdx.x) / p.quarters;
    const int cpq = int(gridDim.x) / p.quarters;              // CTAs per quarter (host launches quarters * cpq)
    const int n_my = r < p.tiles ? (p.tiles - 1 - r) / cpq + 1 : 0;
    const size_t plane_bytes = size_t(p.S) * 16;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_my; ++it) {
                const int tile = r + it * cpq;
                const int sb = it % WG_STAGES;
                if (it >= WG_STAGES) mbar_wait(EMPTY_(sb), ((it / WG_STAGES) & 1) ^ 1);
                mbar_expect_tx(FULL_(sb), (16 + WG_QCH) * C3_TILE_M * 16);
                const uint32_t a_dst = smem_u32(smem + sb * WG_STAGE_BYTES);
                const uint8_t* a_src = reinterpret_cast<const uint8_t*>(p.dy) + size_t(tile) * (C3_TILE_M * 16);
                for (int c = 0; c < 16; ++c)
                    bulk_g2s(a_dst + c * (C3_TILE_M * 16), a_src + size_t(c) * plane_bytes, C3_TILE_M * 16, FULL_(sb));
                const uint32_t b_dst = a_dst + WG_A_BYTES + C3_HALO * 16;
                const uint8_t* b_src = reinterpret_cast<const uint8_t*>(p.x) + size_t(tile) * (C3_TILE_M * 16);
                for (int c = 0; c < WG_QCH; ++c)
                    bulk_g2s(b_dst + c * WG_B_PLANE, b_src + size_t(WG_QCH * q + c) * plane_bytes, C3_TILE_M * 16,
                             FULL_(sb));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_my > 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32_mn(128, 32);
            for (int it = 0; it < n_my; ++it) {
                const int sb = it % WG_STAGES;
                mbar_wait(FULL_(sb), (it / WG_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + sb * WG_STAGE_BYTES);
                const uint32_t b_base = a_base + WG_A_BYTES + C3_HALO * 16;
#pragma unroll 1
                for (int ks = 0; ks < C3_TILE_M / 16; ++ks) {
                    // A: dY^T, M = 128 channels (SBO = chunk plane 4096 B), K = 16 slots (LBO = 128 B per 8 slots)
                    const uint64_t adesc = umma_desc_kmajor_noswz(a_base + ks * 256, 128, C3_TILE_M * 16);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int shift = (tap / 3 - 1) * p.pitch + (tap % 3 - 1);
                        const uint64_t bdesc = umma_desc_kmajor_noswz(b_base + (ks * 16 + shift) * 16, 128, WG_B_PLANE);
                        umma_bf16(tmem_base + tap * 32, adesc, bdesc, idesc, (it | ks) != 0);
                    }
                }
                umma_commit(EMPTY_(sb));        // the stage's smem may be refilled once these MMAs have read it
            }
            umma_commit(ACC);
        }
    } else {
        // epilogue: warp w reads TMEM lane quarter w % 4 (a warp may only touch lanes 32*(warpid % 4) ..)
        const int lq = warp & 3;
        const int co = 32 * lq + lane;
        float* dst = p.scratch + (size_t(blockIdx.x) * 128 + co) * WG_COLS;
        if (n_my > 0) {
            mbar_wait(ACC, 0);
            tc_fence_after();
        }
        for (int tap = 0; tap < 9; ++tap) {
            uint32_t v[32];
            if (n_my > 0) {
                tmem_ld32(tmem_base + tap * 32 + (uint32_t(lq * 32) << 16), v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<uint4*>(dst + tap * 32 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// grad[(co_base + co) * c_in + ci][tap] (+)= sum over the CTAs of ci's quarter; co < co_valid, ci < c_in
static __global__ void k_wgrad_reduce(const float* scratch, int quarters, int cpq, float* grad, int c_in, int co_base,
                                      int co_valid, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (co, ci, tap)
    const int total = co_valid * c_in * 9;
    if (idx >= total) return;
    const int tap = idx % 9, ci = (idx / 9) % c_in, co = idx / (9 * c_in);
    const int qq = ci >> 5, j = ci & 31;
    float acc = 0.f;
    for (int r = 0; r < cpq; ++r)
        acc += scratch[(size_t(r * quarters + qq) * 128 + co) * WG_COLS + tap * 32 + j];
    float* g = grad + (size_t(co_base + co) * c_in + ci) * 9 + tap;
    *g = accumulate ? *g + acc : acc;
}

inline int wgrad_grid(int c_in, int num_sms, int* quarters_out, int* cpq_out) {
    const int quarters = (c_in + 31) / 32;
    const int cpq = num_sms / quarters;
    *quarters_out = quarters;
    *cpq_out = cpq;
    return quarters * cpq;
}
inline size_t wgrad_scratch_elems(int num_sms) { return size_t(num_sms) * WG_PART_ELEMS; }

// dW of one conv (128 output channels starting at dy's chunk 0) into grad[co_base..][c_in][3][3]
inline cudaError_t wgrad_tc_launch(const __nv_bfloat16* dy, const __nv_bfloat16* x, int S, int tiles, int pitch,
                                   int c_in, float* scratch, float* grad, int co_base, int co_valid, int accumulate,
                                   int num_sms, cudaStream_t stream) {
    {   // every launch: see conv3x3_tc3_launch
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    WgradParams p{};
    p.dy = dy; p.x = x; p.scratch = scratch; p.S = S; p.tiles = tiles; p.pitch = pitch;
    int cpq = 0;
    const int grid = wgrad_grid(c_in, num_sms, &p.quarters, &cpq);
    wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, stream>>>(p);
    const int total = co_valid * c_in * 9;
    k_wgrad_reduce<<<(total + 255) / 256, 256, 0, stream>>>(scratch, p.quarters, cpq, grad, c_in, co_base, co_valid,
                                                           accumulate);
    return cudaGetLastError();
}

}  // namespace tb
