// Net5's fully connected policy head as a tcgen05 GEMM (sm_100a): net5.rs:56-62,108
//   logits[b][j] = bias[j] + sum over (pos, c) of  W[j][c*NSQ + pos] * s[b][c][pos]        j < 1575, K = 128 * 25 = 3200
// (the reference calls libtorch's linear; the first version here was a CUDA-core kernel that took 92 % of the Net5
// forward: 5.2 ms of 5.7 ms at 4096 boards against 0.43 ms for the whole conv tower).
//
// Transposed like the conv tower: D^T[128 outputs x 256 boards] += W_tile[128 x 16] * X[256 x 16]^T per K-step, the weights
// are the A operand (M = 128 outputs of one of 13 output tiles), 256 boards are the B operand (N = 256), K runs over
// (board position, 16-channel slab): 25 x 8 = 200 K-steps.  The trunk output lives in strip planes where the boards of a
// tile are NOT 16 B apart, so k_fc_repack first rewrites it as X[pos][chunk of 8 channels][board][8] bf16 -- for a fixed
// (pos, chunk) the boards are then consecutive 16-byte rows, i.e. the K-major no-swizzle operand image (core matrix =
// 8 boards x 16 B), one 4 KiB bulk copy per chunk and 256 boards.  Weights are packed on the host as
// Wp[output tile][pos][slab][kchunk 2][128 outputs][8 channels]: one 32 KiB bulk copy per (tile, pos).
// One pipeline stage = one board position = 8 K-steps (32 KiB weights + 64 KiB activations), 2 stages; one
// tcgen05.commit per stage; the fp32 accumulator (256 TMEM columns) is double buffered across work units
// (output tile, board tile).  Epilogue: TMEM lane = output j, column = board -> + bias -> logits[b][j] fp32 (32 lanes
// write 128 contiguous bytes of one board's row).
//
// Warp roles (192 threads): warp 0 producer, warp 1 TMEM alloc + MMA issuer, warps 2-5 epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "conv_tc3.cuh"
#include "ptx_sm100.cuh"

namespace tb {

constexpr int FC_THREADS = 192;
constexpr int FC_NT = 256;                                   // boards per work unit (MMA N)
constexpr int FC_W_POS_BYTES = 8 * 2 * 128 * 16;             // 32768: weights of one position, 8 slabs x [2][128][8]
constexpr int FC_X_POS_BYTES = 16 * FC_NT * 16;              // 65536: activations of one position, [16 chunks][256][8]
constexpr int FC_STAGE_BYTES = FC_W_POS_BYTES + FC_X_POS_BYTES;   // 98304
constexpr int FC_STAGES = 2;
constexpr int FC_SMEM_BYTES = FC_STAGES * FC_STAGE_BYTES + 1024;

struct FcParams {
    const __nv_bfloat16* wp;   // [jt][pos][slab 8][kc 2][128][8]
    const __nv_bfloat16* x;    // [pos][board tile of FC_NT][chunk 16][FC_NT boards][8]: one stage's B operand is ONE
                               // contiguous 64 KiB block (fc_x_offset)
    const float* bias;         // [n_out]
    float* logits;             // [boards][n_out]
    int n_out, boards, b_pad, n_pos, j_tiles, n_tiles;
};

// Element offset (in units of 8 channels = 16 B) of (position p, channel chunk, column n) in an X image of n_pad
// columns: [p][n / FC_NT][chunk][n % FC_NT].  A GEMM stage (one position, one tile of FC_NT columns, all 16 chunks) is
// contiguous, so the producer fetches it with ONE bulk copy -- as [p][chunk][n_pad] it took 16 copies of 4 KiB, and the
// copy engine's per-copy cost (~70 ns, measured on the wgrad kernel) exceeded the stage's 8 MMAs.
__host__ __device__ inline size_t fc_x_offset(int p, int chunk, int n, int n_pad) {
    return ((size_t(p) * (n_pad / FC_NT) + n / FC_NT) * 16 + chunk) * FC_NT + n % FC_NT;
}

// trunk output strip planes -> X image (fc_x_offset); boards >= n_boards (padding up to b_pad) are written as zero
template <int N, bool PF = false>
__global__ void __launch_bounds__(256) k_fc_repack(const __nv_bfloat16* act, int S, int n_boards, int b_pad,
                                                   __nv_bfloat16* x) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;     // (pos, chunk, board)
    constexpr int NSQ = N * N;
    if (idx >= size_t(NSQ) * 16 * b_pad) return;
    const int b = int(idx % b_pad), chunk = int((idx / b_pad) % 16), pos = int(idx / (size_t(b_pad) * 16));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (b < n_boards)
        v = *reinterpret_cast<const uint4*>(act + (size_t(chunk) * S + SlotMap<N, PF>::slot(b, pos / N, pos % N)) * 8);
    *reinterpret_cast<uint4*>(x + fc_x_offset(pos, chunk, b, b_pad) * 8) = v;
}

static __global__ void __launch_bounds__(FC_THREADS, 1) fc_tc_kernel(const __grid_constant__ FcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tail = smem + FC_STAGES * FC_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);   // full[2], empty[2], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto FULL_ = [&](int i) { return bar0 + 8u * i; };
    auto EMPTY_ = [&](int i) { return bar0 + 8u * (FC_STAGES + i); };
    auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * FC_STAGES + i); };
    auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * FC_STAGES + 2 + i); };
    if (threadIdx.x == 0) {
        for (int i = 0; i < FC_STAGES; ++i) {
            mbar_init(FULL_(i), 1);
            mbar_init(EMPTY_(i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(ACC_FULL(i), 1);
            mbar_init(ACC_EMPTY(i), 4);   // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int units = p.j_tiles * p.n_tiles;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int jt = u % p.j_tiles, nt = u / p.j_tiles;   // neighbouring CTAs share a board tile's activations
                for (int pos = 0; pos < p.n_pos; ++pos, ++cnt) {
                    const int sb = cnt % FC_STAGES;
                    if (cnt >= FC_STAGES) mbar_wait(EMPTY_(sb), ((cnt / FC_STAGES) & 1) ^ 1);
                    mbar_expect_tx(FULL_(sb), FC_STAGE_BYTES);
                    const uint32_t dst = smem_u32(smem + sb * FC_STAGE_BYTES);
                    bulk_g2s(dst, reinterpret_cast<const uint8_t*>(p.wp) + (size_t(jt) * p.n_pos + pos) * FC_W_POS_BYTES,
                             FC_W_POS_BYTES, FULL_(sb));
                    bulk_g2s(dst + FC_W_POS_BYTES, reinterpret_cast<const uint8_t*>(p.x) + fc_x_offset(pos, 0, nt * FC_NT, p.b_pad) * 16,
                             16 * FC_NT * 16, FULL_(sb));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, FC_NT);
            uint32_t cnt = 0;
            int it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(ACC_EMPTY(as), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + as * FC_NT;
                for (int pos = 0; pos < p.n_pos; ++pos, ++cnt) {
                    const int sb = cnt % FC_STAGES;
                    mbar_wait(FULL_(sb), (cnt / FC_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t w_base = smem_u32(smem + sb * FC_STAGE_BYTES);
                    const uint32_t x_base = w_base + FC_W_POS_BYTES;
#pragma unroll
                    for (int slab = 0; slab < 8; ++slab) {
                        const uint64_t wdesc = umma_desc_kmajor_noswz(w_base + slab * 4096, 128 * 16, 128);
                        const uint64_t xdesc = umma_desc_kmajor_noswz(x_base + slab * 2 * (FC_NT * 16), FC_NT * 16, 128);
                        umma_bf16(d_base, wdesc, xdesc, idesc, (pos | slab) != 0);
                    }
                    umma_commit(EMPTY_(sb));
                }
                umma_commit(ACC_FULL(as));
            }
        }
    } else {
        const int lq = warp & 3;
        int it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            const int jt = u % p.j_tiles, nt = u / p.j_tiles;
            const int as = it & 1;
            const int j = jt * 128 + 32 * lq + lane;
            const float bj = j < p.n_out ? p.bias[j] : 0.f;
            mbar_wait(ACC_FULL(as), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * FC_NT + (uint32_t(lq * 32) << 16);
            for (int cc = 0; cc < FC_NT / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                tmem_ld_wait();
                if (cc == FC_NT / 32 - 1) {   // accumulator drained: the issuer may start the unit after next
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ACC_EMPTY(as));
                }
                const int b0 = nt * FC_NT + cc * 32;
                if (j < p.n_out) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (b0 + i < p.boards) p.logits[size_t(b0 + i) * p.n_out + j] = __uint_as_float(v[i]) + bj;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

inline cudaError_t fc_tc_launch(const FcParams& p, int num_sms, cudaStream_t stream) {
    {   // every launch: see conv3x3_tc3_launch
        cudaError_t e = cudaFuncSetAttribute(fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    const int units = p.j_tiles * p.n_tiles;
    if (units <= 0) return cudaSuccess;
    fc_tc_kernel<<<units < num_sms ? units : num_sms, FC_THREADS, FC_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tb
