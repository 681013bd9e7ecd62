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
// GEMM per tile part (half a tile, 128 slots) and tap: D_tap[co 128 x ci N] += dY^T[128 x 16 slots] * X_shifted[16 slots x N], N = c_in rounded up to
// 16 (128 or 96), 8 K-steps per part.  A CTA owns ONE ROW OF TAPS ky (3 taps x N <= 384 TMEM columns of fp32
// accumulators, resident while the CTA walks all its tiles: split-K over tiles across the 49 CTAs of a tap row); at the end
// it writes its partial [128][3][128] to scratch and k_wgrad_reduce sums the partials into the fp32 gradient tensor
// [co][ci][3][3] (the weight blob's own layout).  Within one tap row the three shifts are consecutive slots, so the X
// operand of a 128-slot part is just 130 rows: one pipeline stage = dY part (32 KiB) + X rows (16 x 136 x 16 B = 34 KiB);
// 2 stages (4 parts of 64 slots ran 1.55x slower: 1 KiB bulk copies and a tcgen05.commit per 12 MMAs).  Stage i always
// holds part i of a tile, so the rows that fall outside the tile (the zero halo) are the same
// rows of the same stage for the whole launch: zeroed once, never touched by the bulk copies.
// (First version: a CTA per 32 input channels x 9 taps, N = 32 instructions: 104 us per layer at 4000 positions -- an
// M=128,N=32,K=16 instruction costs ~73 cycles, the 4 KiB A-operand fetch, for 32 cycles of math.)
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
constexpr int WG_PARTS = 2;                                 // pipeline stages = parts of a tile (stage i <-> part i)
constexpr int WG_HALF = C3_TILE_M / WG_PARTS;               // 128 slots per pipeline stage
constexpr int WG_A_BYTES = 16 * WG_HALF * 16;               // 32768: dY part, 16 chunks x 128 slots x 16 B
constexpr int WG_B_ROWS = WG_HALF + 8;                      // 130 used: 128 slots + the kx = -1 / +1 neighbours
constexpr int WG_B_PLANE = WG_B_ROWS * 16;                  // 2176
constexpr int WG_B_BYTES = 16 * WG_B_PLANE;                 // 34816
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;     // 67584
constexpr int WG_SMEM_BYTES = WG_PARTS * WG_STAGE_BYTES + 1024;
constexpr int WG_COLS = 3 * 128;                            // accumulator columns per CTA (512 allocated)
constexpr int WG_PART_ELEMS = 128 * WG_COLS;                // fp32 per CTA partial: [co 128][kx 3][ci 128]

struct WgradParams {
    const __nv_bfloat16* dy;   // strip planes, first of 16 chunks (128 output channels), plane stride S
    const __nv_bfloat16* x;    // strip planes, chunk 0 (input channels), plane stride S
    float* scratch;            // [grid][128][3][128]
    int S;
    int tiles;
    int pitch;
    int n_chunks;              // input channel chunks of 8, even (N = 8 * n_chunks = 96 or 128)
};

// kind::f16 instruction descriptor with BOTH operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_mn(uint32_t M, uint32_t N) {
    return umma_idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
}

static __global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tail = smem + WG_PARTS * WG_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);       // full[PARTS], empty[PARTS], acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto FULL_ = [&](int i) { return bar0 + 8u * i; };
    auto EMPTY_ = [&](int i) { return bar0 + 8u * (WG_PARTS + i); };
    const uint32_t ACC = bar0 + 8u * (2 * WG_PARTS);

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_PARTS; ++i) {
            mbar_init(FULL_(i), 1);
            mbar_init(EMPTY_(i), 1);
        }
        mbar_init(ACC, 1);
        mbar_fence_init();
    }
    // zero every X buffer once: the rows outside the tile are never written by the bulk copies (see the header)
    for (int i = threadIdx.x; i < WG_PARTS * (WG_B_BYTES / 16); i += WG_THREADS) {
        uint8_t* dst = smem + (i / (WG_B_BYTES / 16)) * WG_STAGE_BYTES + WG_A_BYTES + (i % (WG_B_BYTES / 16)) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ky = int(blockIdx.x) % 3;
    const int r = int(blockIdx.x) / 3;
    const int cpg = int(gridDim.x) / 3;                       // CTAs per tap row (host launches 3 * cpg)
    const int n_my = r < p.tiles ? (p.tiles - 1 - r) / cpg + 1 : 0;
    const size_t plane_bytes = size_t(p.S) * 16;
    const int base_shift = (ky - 1) * p.pitch - 1;            // smem X row rr <-> tile slot 128 h + base_shift + rr
    const int N = p.n_chunks * 8;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < WG_PARTS * n_my; ++it) {
                const int tile = r + (it / WG_PARTS) * cpg, h = it % WG_PARTS;
                if (it >= WG_PARTS) mbar_wait(EMPTY_(h), ((it / WG_PARTS) & 1) ^ 1);
                const int u0 = WG_HALF * h + base_shift;                       // tile slot of X row 0
                const int r_lo = u0 < 0 ? -u0 : 0;
                const int r_hi = min(WG_HALF + 2, C3_TILE_M - u0);
                mbar_expect_tx(FULL_(h), 16 * WG_HALF * 16 + p.n_chunks * (r_hi - r_lo) * 16);
                const uint32_t a_dst = smem_u32(smem + h * WG_STAGE_BYTES);
                const uint8_t* a_src = reinterpret_cast<const uint8_t*>(p.dy) + (size_t(tile) * C3_TILE_M + WG_HALF * h) * 16;
                for (int c = 0; c < 16; ++c)
                    bulk_g2s(a_dst + c * (WG_HALF * 16), a_src + size_t(c) * plane_bytes, WG_HALF * 16, FULL_(h));
                const uint32_t b_dst = a_dst + WG_A_BYTES + r_lo * 16;
                const uint8_t* b_src = reinterpret_cast<const uint8_t*>(p.x) + (size_t(tile) * C3_TILE_M + u0 + r_lo) * 16;
                for (int c = 0; c < p.n_chunks; ++c)
                    bulk_g2s(b_dst + c * WG_B_PLANE, b_src + size_t(c) * plane_bytes, (r_hi - r_lo) * 16, FULL_(h));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_my > 0) {
            const uint32_t idesc = umma_idesc_bf16_f32_mn(128, uint32_t(N));
            for (int it = 0; it < WG_PARTS * n_my; ++it) {
                const int h = it % WG_PARTS;
                mbar_wait(FULL_(h), (it / WG_PARTS) & 1);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + h * WG_STAGE_BYTES);
                const uint32_t b_base = a_base + WG_A_BYTES;
#pragma unroll 1
                for (int ks = 0; ks < WG_HALF / 16; ++ks) {
                    // A: dY^T, M = 128 channels (SBO = chunk plane 2048 B), K = 16 slots (LBO = 128 B per 8 slots)
                    const uint64_t adesc = umma_desc_kmajor_noswz(a_base + ks * 256, 128, WG_HALF * 16);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint64_t bdesc = umma_desc_kmajor_noswz(b_base + (ks * 16 + kx) * 16, 128, WG_B_PLANE);
                        umma_bf16(tmem_base + kx * 128, adesc, bdesc, idesc, (it | ks) != 0);
                    }
                }
                umma_commit(EMPTY_(h));         // the stage's smem may be refilled once these MMAs have read it
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
        for (int kx = 0; kx < 3; ++kx)
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t v[32];
                if (n_my > 0) {
                    tmem_ld32(tmem_base + kx * 128 + c0 + (uint32_t(lq * 32) << 16), v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<uint4*>(dst + kx * 128 + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// grad[(co_base + co) * c_in + ci][tap] (+)= sum over the CTAs of the tap's row; co < co_valid, ci < c_in
static __global__ void k_wgrad_reduce(const float* scratch, int cpg, float* grad, int c_in, int co_base, int co_valid,
                                      int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (co, tap, ci): consecutive threads read consecutive ci
    const int total = co_valid * c_in * 9;
    if (idx >= total) return;
    const int ci = idx % c_in, tap = (idx / c_in) % 9, co = idx / (9 * c_in);
    const int ky = tap / 3, kx = tap % 3;
    float acc = 0.f;
    for (int r = 0; r < cpg; ++r)
        acc += scratch[(size_t(r * 3 + ky) * 128 + co) * WG_COLS + kx * 128 + ci];
    float* g = grad + (size_t(co_base + co) * c_in + ci) * 9 + tap;
    *g = accumulate ? *g + acc : acc;
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
    p.n_chunks = ((c_in + 15) / 16) * 2;
    const int cpg = num_sms / 3;
    wgrad_tc_kernel<<<3 * cpg, WG_THREADS, WG_SMEM_BYTES, stream>>>(p);
    const int total = co_valid * c_in * 9;
    k_wgrad_reduce<<<(total + 255) / 256, 256, 0, stream>>>(scratch, cpg, grad, c_in, co_base, co_valid, accumulate);
    return cudaGetLastError();
}

}  // namespace tb
