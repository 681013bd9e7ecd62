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
// operand of a 128-slot part is just 130 rows.
// Operand delivery (round 2): TILED TMA.  Both tensors are described to the TMA unit as 4-D arrays {64 elements = one
// 8-slot x 8-channel core matrix (128 B), 32 slot groups per tile, tiles, 16 chunks}; ONE cp.async.bulk.tensor per
// operand and stage fetches {64, 16 | 18 groups, 1 tile, all chunks}, and because the tile is a dimension of its own
// the slot groups that fall outside the tile (the zero halo of the tap shift) are out-of-bounds coordinates that the
// TMA unit fills with zeros -- no zeroed shared-memory rows, no stage <-> tile-part coupling, so the ring is 3 deep.
// (Round 1 issued 32 one-dimensional bulk copies of ~2 KiB per stage into 2 stages: the kernel was bound by the copy
// engine's per-copy cost, ~70 ns each -- 4 parts of 64 slots, i.e. twice the copies, ran 1.55x slower, and spreading
// the issue over 16 lanes changed nothing -- with its tensor pipe 41 % busy.)  The X box starts at the 8-slot group that
// contains the first needed slot (18 groups = 144 rows cover the 130 needed at any offset); the B descriptor adds the
// offset, as it adds the tap shift.
// (First version: a CTA per 32 input channels x 9 taps, N = 32 instructions: 104 us per layer at 4000 positions -- an
// M=128,N=32,K=16 instruction costs ~73 cycles, the 4 KiB A-operand fetch, for 32 cycles of math.)
//
// Warp roles (320 threads): warp 0 producer (TMA), warp 1 TMEM alloc + MMA issuer, warps 2-9 epilogue (two per TMEM lane
// quarter, alternating 32-column blocks: the partial is 12 dependent tcgen05.ld + store rounds per quarter, exposed at
// the end of every launch).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is resolved through the runtime, no -lcuda)

#include <mutex>
#include <tuple>
#include <map>

#include "conv_tc3.cuh"
#include "ptx_sm100.cuh"

namespace tb {

constexpr int WG_THREADS = 320;
constexpr int WG_PARTS = 2;                                 // parts of a tile (one pipeline stage holds one part)
constexpr int WG_STAGES = 3;                                // ring depth
constexpr int WG_HALF = C3_TILE_M / WG_PARTS;               // 128 slots per pipeline stage
constexpr int WG_A_BYTES = 16 * WG_HALF * 16;               // 32768: dY part, 16 chunks x 128 slots x 16 B
constexpr int WG_B_GROUPS = WG_HALF / 8 + 2;                // 18 slot groups: 130 needed rows at any offset 0..7
constexpr int WG_B_ROWS = WG_B_GROUPS * 8;                  // 144
constexpr int WG_B_PLANE = WG_B_ROWS * 16;                  // 2304
constexpr int WG_B_BYTES = 16 * WG_B_PLANE;                 // 36864
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;     // 69632
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024;   // 209920
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

static __global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ CUtensorMap map_dy,
                const __grid_constant__ CUtensorMap map_x) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tail = smem + WG_STAGES * WG_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);       // full[STAGES], empty[STAGES], acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
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
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ky = int(blockIdx.x) % 3;
    const int r = int(blockIdx.x) / 3;
    const int cpg = int(gridDim.x) / 3;                       // CTAs per tap row (host launches 3 * cpg)
    const int n_my = r < p.tiles ? (p.tiles - 1 - r) / cpg + 1 : 0;
    const int base_shift = (ky - 1) * p.pitch - 1;            // smem X row rr <-> tile slot 128 h + base_shift + rr
    const int N = p.n_chunks * 8;

    auto floor8 = [](int v) { return v >= 0 ? v / 8 : -((-v + 7) / 8); };   // slot -> 8-slot group, towards -infinity
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < WG_PARTS * n_my; ++it) {
                const int tile = r + (it / WG_PARTS) * cpg, h = it % WG_PARTS, st = it % WG_STAGES;
                if (it >= WG_STAGES) mbar_wait(EMPTY_(st), ((it / WG_STAGES) & 1) ^ 1);
                mbar_expect_tx(FULL_(st), WG_A_BYTES + p.n_chunks * WG_B_PLANE);
                const uint32_t a_dst = smem_u32(smem + st * WG_STAGE_BYTES);
                tma_load_4d(a_dst, &map_dy, 0, (WG_HALF / 8) * h, tile, 0, FULL_(st));
                // X rows u0 .. u0 + 129 of the tile (u0 = tile slot of the kx = -1 tap's first row); groups outside
                // 0..31 are out of bounds = zeros
                tma_load_4d(a_dst + WG_A_BYTES, &map_x, 0, floor8(WG_HALF * h + base_shift), tile, 0, FULL_(st));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_my > 0) {
            const uint32_t idesc = umma_idesc_bf16_f32_mn(128, uint32_t(N));
            for (int it = 0; it < WG_PARTS * n_my; ++it) {
                const int h = it % WG_PARTS, st = it % WG_STAGES;
                mbar_wait(FULL_(st), (it / WG_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + st * WG_STAGE_BYTES);
                const int u0 = WG_HALF * h + base_shift;
                const uint32_t b_base = a_base + WG_A_BYTES + (u0 - 8 * floor8(u0)) * 16;
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
                umma_commit(EMPTY_(st));        // the stage's smem may be refilled once these MMAs have read it
            }
            umma_commit(ACC);
        }
    } else {
        // epilogue: warp w reads TMEM lane quarter w % 4 (a warp may only touch lanes 32*(warpid % 4) ..)
        const int lq = warp & 3;
        const int half = (warp - 2) >> 2;
        const int co = 32 * lq + lane;
        float* dst = p.scratch + (size_t(blockIdx.x) * 128 + co) * WG_COLS;
        if (n_my > 0) {
            mbar_wait(ACC, 0);
            tc_fence_after();
        }
        int blk = 0;
        for (int kx = 0; kx < 3; ++kx)
            for (int c0 = 0; c0 < N; c0 += 32) {
                if ((blk++ & 1) != half) continue;
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

// Tensor map of a stack of 16 strip planes [chunk][S slots][8 ch] bf16 as {64 elements (8 slots x 8 ch), 32 groups per
// tile, tiles, 16 chunks} with a box of {64, box_groups, 1, box_chunks}.  Cached per (pointer, S, box): training reuses the
// same buffers every chunk.  cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query.
inline cudaError_t wgrad_tensor_map(const void* base, int S, int box_groups, int box_chunks, CUtensorMap* out) {
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::mutex mu;
    static EncodeFn encode = nullptr;
    static std::map<std::tuple<const void*, int, int, int>, CUtensorMap> cache;
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q{};
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const auto key = std::make_tuple(base, S, box_groups, box_chunks);
    auto it = cache.find(key);
    if (it == cache.end()) {
        CUtensorMap m{};
        const cuuint64_t dims[4] = {64, 32, cuuint64_t(S / C3_TILE_M), 16};
        const cuuint64_t strides[3] = {128, cuuint64_t(C3_TILE_M) * 16, cuuint64_t(S) * 16};   // bytes, dims 1..3
        const cuuint32_t box[4] = {64, cuuint32_t(box_groups), 1, cuuint32_t(box_chunks)};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
        if (cache.size() > 4096) cache.clear();   // buffers were reallocated many times: start over
        it = cache.emplace(key, m).first;
    }
    *out = it->second;
    return cudaSuccess;
}

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
    CUtensorMap map_dy, map_x;
    if (cudaError_t e = wgrad_tensor_map(dy, S, WG_HALF / 8, 16, &map_dy); e != cudaSuccess) return e;
    if (cudaError_t e = wgrad_tensor_map(x, S, WG_B_GROUPS, p.n_chunks, &map_x); e != cudaSuccess) return e;
    const int cpg = num_sms / 3;
    wgrad_tc_kernel<<<3 * cpg, WG_THREADS, WG_SMEM_BYTES, stream>>>(p, map_dy, map_x);
    const int total = co_valid * c_in * 9;
    k_wgrad_reduce<<<(total + 255) / 256, 256, 0, stream>>>(scratch, cpg, grad, c_in, co_base, co_valid, accumulate);
    return cudaGetLastError();
}

}  // namespace tb
