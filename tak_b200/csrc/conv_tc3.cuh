// 3x3 "same" convolution of the AlphaTak residual tower as a tcgen05 implicit GEMM (sm_100a), third generation.
//
// Replaces, per layer, the cuDNN conv + BatchNorm(eval) + ReLU (+ residual add) library calls the reference issues
// through tch-rs (alpha-tak/src/model/net6.rs:70-78, res_block.rs:13-23, policy head net6.rs:99-103).  BatchNorm is
// folded into the weights/bias on the host.
//
// Data layout in HBM ("strip planes"):
//   activations  act[chunk c = 0..15][slot s = 0..S-1][8 channels]  bf16   (channel = 8*c + j)
//   A tile is 256 slots = one STRIP of BPT boards laid side by side: slot(b,y,x) = (b/BPT)*256 + y*PITCH + (b%BPT)*BW
//   + x with BW = N+1 (one zero pad column per board) and PITCH = BPT*BW.  6x6: BPT 6, PITCH 42, 252 of 256 slots used;
//   5x5: BPT 8, PITCH 48.  There are NO pad rows in memory: a tap (ky,kx) is the constant slot shift (ky-1)*PITCH +
//   (kx-1), horizontal neighbours outside a board land on a pad column (or wrap onto the previous strip row's last pad
//   column), vertical neighbours outside the strip land on zero rows that exist only in shared memory.
//   Useful MMA rows: 216/256 (6x6), 200/256 (5x5) -- versus 36/49 and 25/36 for a per-board padded frame.
//   weights  w[slab k = 0..7][ky][kx][kchunk 2][c_out 128][8 c_in] bf16: the K-major no-swizzle UMMA operand image in
//   consumption order, one 36 KiB bulk copy per slab.
//
// GEMM view per tile, TRANSPOSED: D^T[128 c_out x 256 slots] += W[128 x 1152] * X[256 x 1152]^T -- the weights are the
// A operand (M = 128), the activation slab is the B operand (N = 256), one M=128,N=256 fp32 accumulator (256 TMEM
// columns, 2 stages), issued as 8 K-SLABS (16 input channels) x 9 taps of tcgen05.mma.cta_group::1.kind::f16 (K=16).
// N = 256 matters: an SS-mode instruction reads its operands from shared memory, 4+4 KiB per 64 math cycles at
// N = 128 (= all 128 B/clk: measured 79 cycles/instruction) but 4+8 KiB per 128 math cycles at N = 256; the same
// tower ran 42.6 us/layer (1.33 PFLOP/s useful) with two N=128 halves per tap and 35.9 us with one N=256 instruction.
// The slab-outer order is what makes the kernel fit and fast: an activation slab ([zero halo][256 rows][zero halo] x
// 32 B, 11.5 KiB) is dead after its 18 MMAs, so one pipeline stage = {activation slab, the slab's 9 taps of weights
// (36 KiB)} = 47.5 KiB, a 4-stage ring holds 144 KiB of weights in flight, and the issuer needs ONE tcgen05.commit per
// 18 MMAs.  Measured on B200 (profiles/r01_conv_experiments.md): each tcgen05.commit costs the issue stream ~200
// cycles, so a 12 KiB weight stage with its own commit (24 + 8 commits per tile) ran 61-67 us/layer at 4096 boards,
// this arrangement (8 + 1 commits) 41.8 us; cta_group::2 with resident weights is bound by a ~100-cycle floor per N=128
// instruction (58.8 us); the bytes of the weight stream itself are not the limit (a quarter of the bytes: same time).
//
// Warp roles (352 threads, 1 persistent CTA per SM):
//   warp 0     producer: cp.async.bulk of the stage's activation slab (2 x 4 KiB) and weight slab (36 KiB) -> full
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit -> empty / acc_full
//   warps 2-9  epilogue: residual prefetch, tcgen05.ld (software pipelined) -> 8x8 shfl transpose -> +bias (+residual)
//              -> ReLU -> pad mask -> bf16 strip planes, or fp32 logits + per-slot softmax partials (policy head)
//   warp 10    janitor: L2-discards the dead input / residual tiles of finished work items (inference tower only);
//              on the pad-free strip (template parameter PF, the inference build) it is also the MASKING warp: the strip
//              has no pad column (7 boards of 6x6 per tile, 252 of 256 MMA rows real), so the kx = -1 / +1 taps read
//              copies of the activation slab with the rows of board column N-1 / 0 zeroed, which this warp writes into
//              the stage (3 slab copies + weights = 70.5 KiB per stage, 3 stages) as soon as the slab has landed.
//              Timing-only experiment switches of that build: CONV_EXP & 512 (all taps read the unmasked slab),
//              & 1024 (no proxy fence after masking), & 2048 (no masking traffic) -- results are wrong with any of them.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "ptx_sm100.cuh"

namespace tb {

// strip layout shared by every kernel that touches activations / logits
// PF (pad-free, the inference tower): no pad column -- BW = N, 7 boards of 6x6 (10 of 5x5) per tile, 252 (250) of the
// 256 MMA rows are real squares; the horizontal taps read masked copies of the slab instead (see the kernel).  The
// training path keeps the pad column (its weight-gradient kernel relies on it).
template <int N, bool PF = false>
struct SlotMap {
    static constexpr int BW = PF ? N : N + 1;      // board width incl. its zero pad column (none when pad-free)
    static constexpr int BPT = 256 / (N * BW);     // boards per 256-slot tile (strip)
    static constexpr int PITCH = BPT * BW;         // slots per strip row
    static constexpr int USED = N * PITCH;         // slots of a tile that map to a (board, y, x or pad column)
    __host__ __device__ static constexpr int tiles(int boards) { return (boards + BPT - 1) / BPT; }
    __host__ __device__ static constexpr size_t slot(int b, int y, int x) {
        return size_t(b / BPT) * 256 + size_t(y * PITCH + (b % BPT) * BW + x);
    }
};

// The inference path (encode, tower, heads, the search's prior gather) uses the pad-free strip; -DTAK_PADFREE=0 builds it
// on the padded strip of the training path instead (A/B and fallback).
#ifndef TAK_PADFREE
#define TAK_PADFREE 1
#endif
constexpr bool INFER_PF = TAK_PADFREE != 0;

constexpr int C3_TILE_M = 256;
constexpr int C3_HALO = 56;                                // >= PITCH + 1 (43 for 6x6, 49 for 5x5)
constexpr int C3_ROWS = C3_TILE_M + 2 * C3_HALO;           // 368 rows per K chunk of a slab
constexpr int C3_SLAB_BYTES = 2 * C3_ROWS * 16;            // 11776: 16 input channels of one tile (+ zero halos)
constexpr int C3_W_SLAB_BYTES = 9 * 2 * 128 * 16;          // 36864: 9 taps x 16 c_in x 128 c_out
constexpr int C3_STAGE_BYTES = C3_SLAB_BYTES + C3_W_SLAB_BYTES;  // 48640
constexpr int C3_STAGES = 4;                               // ring depth (one stage = one K-slab: activations + weights)
constexpr int C3_MAX_SLABS = 8;                            // 128 input channels
constexpr int C3_THREADS = 352;                            // producer, MMA issuer, 8 epilogue warps, janitor
constexpr int C3_SMEM_BYTES = C3_STAGES * C3_STAGE_BYTES + 3072;
// Pad-free strip (inference tower): a stage holds THREE copies of the activation slab -- as loaded (kx = 0 taps), with the
// rows of board column N-1 zeroed (kx = -1 taps: the left neighbour of column 0 is not the previous board's last column)
// and with the rows of column 0 zeroed (kx = +1 taps) -- and the ring is 3 deep (measured: a 3-stage ring costs nothing).
constexpr int C3_PF_STAGES = 3;
constexpr int C3_PF_STAGE_BYTES = 3 * C3_SLAB_BYTES + C3_W_SLAB_BYTES;   // 72192
constexpr int C3_PF_SMEM_BYTES = C3_PF_STAGES * C3_PF_STAGE_BYTES + 3072; // 219648
constexpr size_t C3_W_LAYER_ELEMS = size_t(C3_MAX_SLABS) * 9 * 2 * 128 * 8;  // bf16 elements per packed layer
constexpr int C3_TILE_ALIGN = 1;

enum ConvMode : int {
    CONV_RELU = 0,        // out = relu(conv + bias)                     -> bf16 strip planes
    CONV_RES_RELU = 1,    // out = relu(conv + bias + res)               -> bf16 strip planes
    CONV_LOGITS_F32 = 2,  // out = conv + bias  (no activation)          -> fp32 [channel][slot] + softmax partials
    // training build of the kernel only (conv3x3_tc3_kernel<true>):
    CONV_LINEAR = 3,      // out = conv + bias (+ res if res != nullptr), no activation -> bf16 strip planes; if
                          // stats != nullptr, per-channel sum / sum of squares of the fp32 values over the real squares
                          // (forward_training's batch statistics, net6.rs:72-76 with train = true; dgrad uses it with
                          // a zero bias and no stats)
};

// one convolution of the tower
struct ConvLayerDesc {
    const __nv_bfloat16* in;    // strip planes [16][S][8]
    const __nv_bfloat16* res;   // strip planes (mode 1) or nullptr
    __nv_bfloat16* out;         // strip planes (modes 0/1)
    float* out_f32;             // [out_ch_total][S] (mode 2)
    float2* partials;           // mode 2: [group*4 + lane quarter][S] per-slot {max, sum exp(l - max)} over 32 channels
    const __nv_bfloat16* w;     // [slabs][3][3][2][128][8]
    const float* bias;          // [128]
    int slabs;                  // ceil(c_in / 16) <= 8
    int mode;
    int out_ch_offset;          // mode 2: first output channel of this 128-wide group
    int out_ch_valid;           // mode 2: number of real channels in this group (<=128)
    int group;                  // mode 2: group index for `partials`
    int discard;                // bit 2: fp32 logits are stored with the streaming hint (st.global.cs)
                                // bit 0: `in`, bit 1: `res` are dead after this layer -> their L2 lines are discarded
                                // (never written back to HBM); only the fused inference tower sets it
    double* stats;              // mode 3: [2][128] {sum, sum of squares} per output channel, accumulated (or nullptr)
};

constexpr int C3_MAX_LAYERS = 36;   // Net6: 1 + 2*16 trunk convs + 2 policy groups
constexpr int C3_GROUP = 3;         // tiles a CTA carries through the whole tower together

// The whole conv tower in ONE launch.  A strip tile never reads another tile (zero halos, pad columns), so a CTA can
// take a tile through every layer without any grid-wide synchronisation: CTA c owns tiles c, c+grid, ...; it walks
// them in groups of <= 3, running all layers over a group before moving on:  for group { for layer { for tile in group
// } }.  While the tensor core works on (layer, tile B) the epilogue of (layer, tile A) stores A's outputs and the
// producer already streams them back in for (layer+1, tile A); the only hand-off is the per-tile `ready` mbarrier.
// A group's activations (3 tiles x 3 buffers x 64 KiB per CTA, 85 MB per GPU) never leave L2, and the per-layer launch
// prologue / exposed last epilogue of a layer-per-launch schedule disappear.
struct ConvParams {
    ConvLayerDesc layers[C3_MAX_LAYERS];
    int n_layers;
    int S;                      // plane stride in slots (= allocated tiles * 256)
    int tile_begin, tile_end;   // tiles to process
    int n_boards;
    const int* n_boards_dev;    // if set: the number of boards is read from device memory when the kernel starts (the
                                // search loop counts its leaves on the device; tile_end = tile_begin + tiles(boards) and
                                // the launch is sized for the largest batch) -- no host round trip per evaluation
    int n;                      // board size N
    int bw;                     // N + 1 (padded strip) or N (pad-free strip)
    int bpt;                    // boards per tile
    int pitch;                  // bpt * bw
    // Training build, one-layer CONV_LINEAR launches that are a dgrad (train.cu): when bnb_y is set, the conv's output is
    // the gradient w.r.t. the post-ReLU output of the BatchNorm layer below, and the epilogue fuses the first pass of
    // that layer's BatchNorm backward: out = g' = (conv + bias + res) * (bnb_z > 0), and stats[c] += sum g',
    // stats[128 + c] += sum g' * xhat with xhat = (bnb_y - mean) * rstd  (what k_bn_bwd_reduce computes from HBM).
    const __nv_bfloat16* bnb_z;  // the layer's post-ReLU output (the ReLU mask)
    const __nv_bfloat16* bnb_y;  // the layer's raw conv output (pre-BN)
    const float* bnb_mean;       // [128] batch mean / 1/sqrt(var + eps) saved by k_bn_apply
    const float* bnb_rstd;
};

// barrier slots
enum : int {
    C3B_FULL = 0,                          // [C3_STAGES] bulk-copy completion (activation slab + weight slab)
    C3B_EMPTY = C3B_FULL + C3_STAGES,      // [C3_STAGES] tcgen05.commit after the stage's 18 MMAs
    C3B_ACC_FULL = C3B_EMPTY + C3_STAGES,  // [2]
    C3B_ACC_EMPTY = C3B_ACC_FULL + 2,      // [2]
    C3B_READY = C3B_ACC_EMPTY + 2,         // [C3_GROUP] outputs of (layer, tile-in-group) stored by all 8 epilogue warps
    C3B_MASKED = C3B_READY + C3_GROUP,     // [C3_STAGES] pad-free: the two masked copies of the stage's slab are written
    C3B_AFULL = C3B_MASKED + C3_STAGES,    // [C3_STAGES] pad-free: the stage's ACTIVATION slab has landed (its weights: FULL)
    C3B_COUNT = C3B_AFULL + C3_STAGES
};

// one round of the 8x8 block transpose across the 8 lanes of a channel chunk (blocks = x[8i + b], i = 0..3)
template <int M>
__device__ __forceinline__ void transpose8_stage(uint32_t (&x)[32], int j) {
    const bool up = (j & M) != 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b & M) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t send = up ? x[8 * i + b] : x[8 * i + (b | M)];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, M);
            if (up) x[8 * i + b] = recv; else x[8 * i + (b | M)] = recv;
        }
    }
}

// work-item walk shared by the three roles: groups of tiles, all layers per group
struct TowerWalk {
    int n_my, n_groups, base, rem;
    __device__ TowerWalk(const ConvParams& p, int tile_end) {
        const int first = p.tile_begin + blockIdx.x;
        n_my = first < tile_end ? (tile_end - 1 - first) / int(gridDim.x) + 1 : 0;
        n_groups = (n_my + C3_GROUP - 1) / C3_GROUP;
        base = n_groups ? n_my / n_groups : 0;
        rem = n_groups ? n_my % n_groups : 0;
    }
    __device__ int group_size(int g) const { return base + (g < rem ? 1 : 0); }
};

// TAK_TOWER_LOWREG=1: the pad-free inference build is compiled for at most 128 registers per thread (a 512-thread launch
// bound; it is still launched with 352), so that 12 allocated warps x 4 096 registers leave 16 384 on the SM: room for
// eight warps of the other engine replica's search step beside a resident tower CTA (profiles/r02_step_overlap.md).
#ifndef TAK_TOWER_LOWREG
#define TAK_TOWER_LOWREG 0
#endif
template <bool TRAIN, bool PF = false>
static __global__ void __launch_bounds__((TAK_TOWER_LOWREG && PF && !TRAIN) ? 512 : C3_THREADS, 1)
conv3x3_tc3_kernel(const __grid_constant__ ConvParams p) {
    static_assert(!(TRAIN && PF), "the training build keeps the padded strip");
    constexpr int STAGES = PF ? C3_PF_STAGES : C3_STAGES;
    constexpr int STAGE_BYTES = PF ? C3_PF_STAGE_BYTES : C3_STAGE_BYTES;
    constexpr int COPIES = PF ? 3 : 1;                       // activation slab copies per stage
    constexpr int W_OFF = COPIES * C3_SLAB_BYTES;            // the stage's weight slab follows them
#if defined(CONV_PF_BULK3)
    constexpr int PF_COPIES_LOADED = 3;   // variant: the bulk-copy engine fills all three copies, the masking warp only zeroes
#else
    constexpr int PF_COPIES_LOADED = 1;   // the masking warp reads the loaded slab and writes the two masked copies
#endif
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stage_buf = smem;                               // STAGES x {activation slab (x3 when pad-free), weight slab}
    uint8_t* tail = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + C3B_COUNT * 8);
    float* s_bias = reinterpret_cast<float*>(tail + C3B_COUNT * 8 + 16);   // [2][128]
    float* s_stats = s_bias + 256;                                         // [2][128] (TRAIN: per-CTA batch statistics)
    if (TRAIN) {
        for (int i = threadIdx.x; i < 256; i += C3_THREADS) s_stats[i] = 0.f;
    }

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(BAR(C3B_FULL + i), 1);
            mbar_init(BAR(C3B_EMPTY + i), 1);
            mbar_init(BAR(C3B_MASKED + i), 1);
            mbar_init(BAR(C3B_AFULL + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(C3B_ACC_FULL + i), 1);
            mbar_init(BAR(C3B_ACC_EMPTY + i), 8);  // one arrive per epilogue warp
        }
        for (int i = 0; i < C3_GROUP; ++i) mbar_init(BAR(C3B_READY + i), 8);
        mbar_fence_init();
    }
    // Programmatic dependent launch: the kernel after the tower may be scheduled as SMs free up; everything before
    // griddep_wait() (barrier init, halo fill, TMEM alloc, first weight slab) overlaps the previous kernel's tail.
    griddep_launch_dependents();
    // zero halos: the rows above / below the 256 loaded rows of every slab buffer are never written by the bulk copies
    for (int i = threadIdx.x; i < STAGES * COPIES * 2 * 2 * C3_HALO; i += C3_THREADS) {
        const int row = i % C3_HALO, side = (i / C3_HALO) & 1, plane = i / (2 * C3_HALO);  // plane = (stage*COPIES + copy)*2 + kchunk
        const int kc = plane & 1, copy = (plane >> 1) % COPIES, stage = (plane >> 1) / COPIES;
        uint8_t* dst = stage_buf + stage * STAGE_BYTES + copy * C3_SLAB_BYTES + kc * (C3_ROWS * 16) +
                       (side ? (C3_HALO + C3_TILE_M + row) : row) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros visible to the tensor core
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const size_t plane_bytes = static_cast<size_t>(p.S) * 16;
    int n_boards = p.n_boards, tile_end = p.tile_end;
    if (p.n_boards_dev != nullptr) {
        griddep_wait();   // the count is written by the kernel before this one
        n_boards = *reinterpret_cast<const volatile int*>(p.n_boards_dev);
        const int t = (n_boards + p.bpt - 1) / p.bpt;
        tile_end = p.tile_begin + (t < p.tile_end - p.tile_begin ? t : p.tile_end - p.tile_begin);
    }
    const TowerWalk walk(p, tile_end);
    const int tile0 = p.tile_begin + blockIdx.x;

    if (warp == 0) {
        // ===================== producer: one stage = weight slab (36 KiB) + activation slab (2 x 4 KiB) ============
        if (lane == 0) {
            uint32_t cnt = 0;
            uint32_t items_of[C3_GROUP] = {0, 0, 0};  // work items issued so far per tile-in-group = `ready` completions due
            bool first_a = true;
            for (int g = 0, j0 = 0; g < walk.n_groups; j0 += walk.group_size(g), ++g) {
                const int gs = walk.group_size(g);
                for (int L = 0; L < p.n_layers; ++L) {
                    const ConvLayerDesc& ld = p.layers[L];
                    for (int jj = 0; jj < gs; ++jj) {
                        const int tile = tile0 + (j0 + jj) * int(gridDim.x);
                        const uint8_t* src0 = reinterpret_cast<const uint8_t*>(ld.in) + static_cast<size_t>(tile) * (C3_TILE_M * 16);
                        for (int k = 0; k < ld.slabs; ++k, ++cnt) {
                            const int sb = cnt % STAGES;
#if defined(CONV_EXP) && (CONV_EXP & 4)
                            if (cnt >= STAGES) continue;
#endif
                            if (cnt >= STAGES) mbar_wait(BAR(C3B_EMPTY + sb), ((cnt / STAGES) & 1) ^ 1);
                            // pad-free: the activation slab reports on its own barrier, so that the masking warp can start
                            // on it while the (4.5x larger) weight slab is still in flight; for k > 0 it also goes out first
                            const uint32_t abar = PF ? BAR(C3B_AFULL + sb) : BAR(C3B_FULL + sb);
                            const uint32_t dst = smem_u32(stage_buf + sb * STAGE_BYTES) + C3_HALO * 16;
                            if (PF) {
                                mbar_expect_tx(BAR(C3B_AFULL + sb), PF_COPIES_LOADED * 2 * C3_TILE_M * 16);
#if defined(CONV_EXP) && (CONV_EXP & 4096)
                                mbar_expect_tx(BAR(C3B_FULL + sb), cnt < STAGES ? C3_W_SLAB_BYTES : 0);
#else
                                mbar_expect_tx(BAR(C3B_FULL + sb), C3_W_SLAB_BYTES);
#endif
                                if (k > 0) {
#pragma unroll
                                    for (int c = 0; c < PF_COPIES_LOADED; ++c) {
                                        bulk_g2s(dst + c * C3_SLAB_BYTES, src0 + static_cast<size_t>(2 * k) * plane_bytes,
                                                 C3_TILE_M * 16, abar);
                                        bulk_g2s(dst + c * C3_SLAB_BYTES + C3_ROWS * 16,
                                                 src0 + static_cast<size_t>(2 * k + 1) * plane_bytes, C3_TILE_M * 16, abar);
                                    }
                                }
                            } else {
#if defined(CONV_EXP) && (CONV_EXP & 4096)
                                mbar_expect_tx(BAR(C3B_FULL + sb), 2 * C3_TILE_M * 16 + (cnt < STAGES ? C3_W_SLAB_BYTES : 0));
#else
                                mbar_expect_tx(BAR(C3B_FULL + sb), 2 * C3_TILE_M * 16 + C3_W_SLAB_BYTES);
#endif
                            }
                            // weights never depend on anything computed here: at k == 0 they go out first
#if defined(CONV_EXP) && (CONV_EXP & 4096)
                            // experiment (wrong results, timing only): weight slabs are fetched for the first ring
                            // revolution only -- the ceiling of what sharing weight traffic between CTAs could buy
                            if (cnt < STAGES)
#endif
                            bulk_g2s(smem_u32(stage_buf + sb * STAGE_BYTES + W_OFF),
                                     reinterpret_cast<const uint8_t*>(ld.w) + static_cast<size_t>(k) * C3_W_SLAB_BYTES,
                                     C3_W_SLAB_BYTES, BAR(C3B_FULL + sb));
                            if (PF && k > 0) continue;
                            if (k == 0) {
                                if (first_a) {
                                    griddep_wait();  // the tower's input planes are written by the previous kernel
                                    first_a = false;
                                }
                                // this tile's previous layer must have been stored by the epilogue: ready[jj] completes
                                // once per work item of slot jj, the latest one being (L-1, this tile)
                                if (L > 0) mbar_wait(BAR(C3B_READY + jj), (items_of[jj] - 1) & 1);
                            }
#pragma unroll
                            for (int c = 0; c < (PF ? PF_COPIES_LOADED : 1); ++c) {
                                bulk_g2s(dst + c * C3_SLAB_BYTES, src0 + static_cast<size_t>(2 * k) * plane_bytes, C3_TILE_M * 16,
                                         abar);
                                bulk_g2s(dst + c * C3_SLAB_BYTES + C3_ROWS * 16,
                                         src0 + static_cast<size_t>(2 * k + 1) * plane_bytes, C3_TILE_M * 16, abar);
                            }
                        }
                        items_of[jj]++;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 256);  // D^T[c_out 128 x 256 slots]
            uint32_t scnt = 0;
            int it = 0;
            for (int g = 0; g < walk.n_groups; ++g) {
                const int gs = walk.group_size(g);
                for (int L = 0; L < p.n_layers; ++L) {
                    const int n_slabs = p.layers[L].slabs;
                    for (int jj = 0; jj < gs; ++jj, ++it) {
                        const int as = it & 1;
                        mbar_wait(BAR(C3B_ACC_EMPTY + as), ((it >> 1) & 1) ^ 1);  // accumulator drained by the epilogue
                        tc_fence_after();
                        const uint32_t d_base = tmem_base + as * 256;
#pragma unroll 1
                        for (int k = 0; k < n_slabs; ++k, ++scnt) {
                            const int sb = scnt % STAGES;
#if defined(CONV_EXP) && (CONV_EXP & 4)
                            if (scnt < STAGES)
#endif
                            mbar_wait(BAR(C3B_FULL + sb), (scnt / STAGES) & 1);
                            if (PF) mbar_wait(BAR(C3B_AFULL + sb), (scnt / STAGES) & 1);
                            tc_fence_after();
                            const uint32_t a_base = smem_u32(stage_buf + sb * STAGE_BYTES) + C3_HALO * 16;
                            const uint32_t w_base = smem_u32(stage_buf + sb * STAGE_BYTES + W_OFF);
                            if constexpr (!PF) {
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    const int shift = (tap / 3 - 1) * p.pitch + (tap % 3 - 1);
                                    const uint64_t wdesc = umma_desc_kmajor_noswz(w_base + tap * 4096, 128 * 16, 128);
                                    const uint64_t xdesc = umma_desc_kmajor_noswz(a_base + shift * 16, C3_ROWS * 16, 128);
                                    umma_bf16(d_base, wdesc, xdesc, idesc, (k | tap) != 0);
                                }
                            } else {
                                // the three centre-column taps read the slab as loaded; they are issued first so that the
                                // masking warp has ~450 cycles of tensor work to hide behind on top of its head start
#pragma unroll
                                for (int ky = 0; ky < 3; ++ky) {
                                    const int tap = ky * 3 + 1, shift = (ky - 1) * p.pitch;
                                    const uint64_t wdesc = umma_desc_kmajor_noswz(w_base + tap * 4096, 128 * 16, 128);
                                    const uint64_t xdesc = umma_desc_kmajor_noswz(a_base + shift * 16, C3_ROWS * 16, 128);
                                    umma_bf16(d_base, wdesc, xdesc, idesc, (k | ky) != 0);
                                }
                                mbar_wait(BAR(C3B_MASKED + sb), (scnt / STAGES) & 1);
                                tc_fence_after();
#pragma unroll
                                for (int side = 0; side < 2; ++side) {       // kx = -1 taps on copy 1, kx = +1 taps on copy 2
#pragma unroll
                                    for (int ky = 0; ky < 3; ++ky) {
                                        const int tap = ky * 3 + 2 * side, shift = (ky - 1) * p.pitch + (2 * side - 1);
                                        const uint64_t wdesc = umma_desc_kmajor_noswz(w_base + tap * 4096, 128 * 16, 128);
#if defined(CONV_EXP) && (CONV_EXP & 512)
                                        const uint64_t xdesc = umma_desc_kmajor_noswz(a_base + shift * 16, C3_ROWS * 16, 128);
#else
                                        const uint64_t xdesc = umma_desc_kmajor_noswz(
                                            a_base + (1 + side) * C3_SLAB_BYTES + shift * 16, C3_ROWS * 16, 128);
#endif
                                        umma_bf16(d_base, wdesc, xdesc, idesc, true);
                                    }
                                }
                            }
                            // ONE commit per slab: tcgen05.commit costs the issue stream ~200 cycles (measured), so
                            // the activation slab and its weights are released together
                            umma_commit(BAR(C3B_EMPTY + sb));
                        }
                        umma_commit(BAR(C3B_ACC_FULL + as));            // accumulators ready
                    }
                }
            }
        }
    } else if (warp >= 2 && warp < 10) {
        // ===================== epilogue (8 warps) =====================
        // The accumulator is D^T: TMEM lane = output channel, column = slot.  Warp w reads lane quarter q = w%4
        // (channels 32q..32q+31) and column half h (slots 128h..128h+127) in 4 chunks of 32 columns; a thread starts
        // with ONE channel x 32 slots.  The strip planes want 8 channels x 16 B per slot, so the 8 lanes of a channel
        // chunk transpose their 8x8 blocks through 3 rounds of shfl.xor (fp32, before any rounding): lane j of group g
        // ends with channels 8*(4q+g)..+7 of slots 8i+j (i = 0..3) -> 16 B vector loads of the residual and 16 B
        // stores, 128 B contiguous across the 8 lanes.
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        const int grp8 = lane >> 3, j = lane & 7;
        const int chunk = 4 * q + grp8;               // 8-channel chunk this thread stores
        const int etid = threadIdx.x - 64;            // 0..255 among the epilogue threads
        // frame validity / board index of this thread's 16 slots per tile (slot = 128h + 32cc + 8i + j)
        uint32_t frame_mask = 0;
        uint64_t board_of = 0;                        // 4 bits per slot
#pragma unroll
        for (int idx = 0; idx < 16; ++idx) {
            const int sl = 128 * h + 32 * (idx >> 2) + 8 * (idx & 3) + j;
            const int ry = sl / p.pitch, rrem = sl - ry * p.pitch;
            const int rj = rrem / p.bw, rx = rrem - rj * p.bw;
            if (ry < p.n && rx < p.n) frame_mask |= 1u << idx;   // a real square (not a pad column / tile remainder)
            board_of |= static_cast<uint64_t>(rj & 15) << (4 * idx);
        }
        griddep_wait();  // output / residual buffers belong to earlier kernels until they complete
        int it = 0;
        for (int g = 0, j0 = 0; g < walk.n_groups; j0 += walk.group_size(g), ++g) {
            const int gs = walk.group_size(g);
            for (int L = 0; L < p.n_layers; ++L) {
                const ConvLayerDesc& ld = p.layers[L];
                const int mode = ld.mode;
                for (int jj = 0; jj < gs; ++jj, ++it) {
                    const int tile = tile0 + (j0 + jj) * int(gridDim.x);
                    const int as = it & 1;
                    const uint32_t ph2 = (it >> 1) & 1;
                    const size_t slot0 = static_cast<size_t>(tile) * C3_TILE_M + 128 * h + j;
                    uint32_t valid_mask = 0;
#pragma unroll
                    for (int idx = 0; idx < 16; ++idx)
                        if (((frame_mask >> idx) & 1) && tile * p.bpt + int((board_of >> (4 * idx)) & 15) < n_boards)
                            valid_mask |= 1u << idx;
                    // bias of this layer -> shared (double buffered by accumulator stage; the named barrier keeps the
                    // 256 epilogue threads within one work item of each other)
                    // (training build: a launch is one layer, so the bias is staged once per group of tiles and the
                    // per-tile rendezvous of the eight warps is gone)
                    float* bias_s = s_bias + (TRAIN ? 0 : as * 128);
                    if (!TRAIN || jj == 0) {
                        if (TRAIN) asm volatile("bar.sync 1, 256;" ::: "memory");   // nobody still reads the previous bias
                        if (etid < 128) bias_s[etid] = ld.bias[etid];
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                    }
                    // the residual does not depend on the MMAs: the first chunk's 4 x 16 B are fetched before waiting for
                    // the accumulator, the next chunk's while the current one is processed (registers: a 352-thread CTA
                    // is allocated as 12 warps, i.e. 168 registers per thread at most).
                    // Plain (coherent) loads: these slots were stored by this very thread two layers ago.
                    auto load_res = [&](int cc, uint4 (&dst)[4]) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dst[i] = make_uint4(0, 0, 0, 0);
#if defined(CONV_EXP) && (CONV_EXP & 32)
                            if (p.S == -7)                            // experiment: no residual loads
#endif
                            if ((valid_mask >> (cc * 4 + i)) & 1)
                                dst[i] = *reinterpret_cast<const uint4*>(
                                    ld.res + (static_cast<size_t>(chunk) * p.S + slot0 + 32 * cc + 8 * i) * 8);
                        }
                    };
                    uint4 res[TRAIN ? 1 : 2][4];   // training build: single-buffered (registers), fetched one chunk ahead
                    const bool has_res = mode == CONV_RES_RELU || (TRAIN && mode == CONV_LINEAR && ld.res != nullptr);
                    const bool relu = !(TRAIN && mode == CONV_LINEAR);
                    const bool want_stats = TRAIN && mode == CONV_LINEAR && ld.stats != nullptr;
                    // fused BatchNorm-backward reduction (see ConvParams::bnb_y): the mask and the raw conv output of the
                    // layer below are fetched one 32-slot chunk ahead (single-buffered: issued when the previous chunk's
                    // values have been consumed)
                    const bool bnb = TRAIN && mode == CONV_LINEAR && p.bnb_y != nullptr;
                    uint4 bz[4], by[4];
                    auto load_zy = [&](int cc) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            bz[i] = by[i] = make_uint4(0, 0, 0, 0);
                            if ((valid_mask >> (cc * 4 + i)) & 1) {
                                const size_t off = (static_cast<size_t>(chunk) * p.S + slot0 + 32 * cc + 8 * i) * 8;
                                bz[i] = __ldg(reinterpret_cast<const uint4*>(p.bnb_z + off));
                                by[i] = __ldg(reinterpret_cast<const uint4*>(p.bnb_y + off));
                            }
                        }
                    };
                    float st_sum[8], st_sq[8];
                    if (TRAIN) {
#pragma unroll
                        for (int b = 0; b < 8; ++b) st_sum[b] = st_sq[b] = 0.f;
                        if (bnb) load_zy(0);
                    }
                    if (has_res) load_res(0, res[0]);
                    float bias8[8];
#pragma unroll
                    for (int b = 0; b < 8; ++b) bias8[b] = bias_s[chunk * 8 + b];
                    const float bias_c = bias_s[32 * q + lane];
                    mbar_wait(BAR(C3B_ACC_FULL + as), ph2);
                    tc_fence_after();
#if defined(CONV_EXP) && (CONV_EXP & 2)
                    if (p.S != -12345) {
                        tc_fence_before(); __syncwarp();
                        if (lane == 0) { mbar_arrive(BAR(C3B_ACC_EMPTY + as)); mbar_arrive(BAR(C3B_READY + jj)); }
                        continue;
                    }
#endif
                    const uint32_t taddr = tmem_base + as * 256 + 128 * h + (static_cast<uint32_t>(q * 32) << 16);
                    uint32_t x[32];
                    // training build: the chunk loop stays rolled (single-buffered operands, no register array indexed by
                    // cc) -- a quarter of the epilogue's code, which a one-layer launch runs through only 4-5 times
#pragma unroll(TRAIN ? 1 : 4)
                    for (int cc = 0; cc < 4; ++cc) {
                        tmem_ld32(taddr + cc * 32, x);   // 8 epilogue warps hide each other's TMEM latency
                        if (!TRAIN && has_res && cc < 3) load_res(cc + 1, res[TRAIN ? 0 : ((cc + 1) & 1)]);
                        tmem_ld_wait();
                        if (cc == 3) {
                            // the accumulator is in registers now: hand it back to the MMA issuer BEFORE the stores of
                            // this tile are issued and made visible (the fence below waits for them; in a one-layer
                            // launch with outputs going to HBM that wait used to sit on the tensor pipe's critical path)
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(BAR(C3B_ACC_EMPTY + as));
                        }
                        if (!TRAIN && mode == CONV_LOGITS_F32) {
                            // logits straight from the un-transposed registers: one channel, 32 consecutive slots
                            const int ch = 32 * q + lane;
                            const bool ch_ok = ch < ld.out_ch_valid;
#pragma unroll
                            for (int s4 = 0; s4 < 32; s4 += 4) {
                                float4 o;
                                o.x = __uint_as_float(x[s4 + 0]) + bias_c;
                                o.y = __uint_as_float(x[s4 + 1]) + bias_c;
                                o.z = __uint_as_float(x[s4 + 2]) + bias_c;
                                o.w = __uint_as_float(x[s4 + 3]) + bias_c;
                                x[s4 + 0] = __float_as_uint(ch_ok ? o.x : -INFINITY);
                                x[s4 + 1] = __float_as_uint(ch_ok ? o.y : -INFINITY);
                                x[s4 + 2] = __float_as_uint(ch_ok ? o.z : -INFINITY);
                                x[s4 + 3] = __float_as_uint(ch_ok ? o.w : -INFINITY);
                                if (ch_ok) {
                                    // the logits stream out (read later by other kernels, never by the tower): with the
                                    // streaming hint they do not push the tower's live activation tiles out of L2
                                    float4* dst = reinterpret_cast<float4*>(ld.out_f32 + static_cast<size_t>(ld.out_ch_offset + ch) * p.S +
                                                                            (slot0 - j) + 32 * cc + s4);
                                    if (ld.discard & 4) __stcs(dst, o); else *dst = o;
                                }
                            }
                        }
                        // 8x8 block transpose across the 8 lanes of a channel chunk: x[8i+b] (channel j, slot 8i+b) ->
                        // x[8i+b] (channel b, slot 8i+j)
#if !(defined(CONV_EXP) && (CONV_EXP & 16))
                        transpose8_stage<4>(x, j);
                        transpose8_stage<2>(x, j);
                        transpose8_stage<1>(x, j);
#endif
                        if (!TRAIN && mode == CONV_LOGITS_F32) {
                            // per-slot softmax partial over this warp's 32 channels: 8 in-thread, then the 4 lane groups
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float mx = -INFINITY;
#pragma unroll
                                for (int b = 0; b < 8; ++b) mx = fmaxf(mx, __uint_as_float(x[8 * i + b]));
                                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
                                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
                                float sum = 0.f;
                                if (mx > -INFINITY) {
#pragma unroll
                                    for (int b = 0; b < 8; ++b) sum += __expf(__uint_as_float(x[8 * i + b]) - mx);
                                }
                                sum += __shfl_xor_sync(0xffffffffu, sum, 8);
                                sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                                if (grp8 == 0)
                                    ld.partials[static_cast<size_t>(ld.group * 4 + q) * p.S + slot0 + 32 * cc + 8 * i] =
                                        make_float2(mx, sum);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int idx = cc * 4 + i;
                                const bool valid = (valid_mask >> idx) & 1;
                                float v[8];
#pragma unroll
                                for (int b = 0; b < 8; ++b) v[b] = __uint_as_float(x[8 * i + b]) + bias8[b];
                                if (has_res) {
                                    const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&res[TRAIN ? 0 : (cc & 1)][i]);
#pragma unroll
                                    for (int b = 0; b < 4; ++b) {
                                        float2 f = __bfloat1622float2(rb[b]);
                                        v[2 * b] += f.x;
                                        v[2 * b + 1] += f.y;
                                    }
                                }
                                if (TRAIN) {
                                    if (bnb) {   // ReLU mask of the layer below: the gradient passes where its output was > 0
                                        const __nv_bfloat162* zb = reinterpret_cast<const __nv_bfloat162*>(&bz[i]);
#pragma unroll
                                        for (int b = 0; b < 4; ++b) {
                                            const float2 zf = __bfloat1622float2(zb[b]);
                                            if (!(zf.x > 0.f)) v[2 * b] = 0.f;
                                            if (!(zf.y > 0.f)) v[2 * b + 1] = 0.f;
                                        }
                                    }
                                }
                                uint4 ov;
                                __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                                for (int b = 0; b < 4; ++b) {
                                    float lo = valid ? (relu ? fmaxf(v[2 * b], 0.0f) : v[2 * b]) : 0.0f;
                                    float hi = valid ? (relu ? fmaxf(v[2 * b + 1], 0.0f) : v[2 * b + 1]) : 0.0f;
                                    ob[b] = __floats2bfloat162_rn(lo, hi);
                                }
                                if (TRAIN) {
                                    if (bnb) {
                                        // sums over the bf16 values the BatchNorm-backward apply pass will read back;
                                        // the second sum is sum g' * y here: centred and scaled (xhat = (y - mean) * rstd)
                                        // in double at the CTA's flush
                                        if (valid) {
                                            const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&by[i]);
#pragma unroll
                                            for (int b = 0; b < 4; ++b) {
                                                const float2 gf = __bfloat1622float2(ob[b]);
                                                const float2 yf = __bfloat1622float2(yb[b]);
                                                st_sum[2 * b] += gf.x;
                                                st_sum[2 * b + 1] += gf.y;
                                                st_sq[2 * b] = __fmaf_rn(gf.x, yf.x, st_sq[2 * b]);   // (the library is built -fmad=false)
                                                st_sq[2 * b + 1] = __fmaf_rn(gf.y, yf.y, st_sq[2 * b + 1]);
                                            }
                                        }
                                    } else if (want_stats && valid) {
#pragma unroll
                                        for (int b = 0; b < 8; ++b) {
                                            st_sum[b] += v[b];
                                            st_sq[b] = __fmaf_rn(v[b], v[b], st_sq[b]);
                                        }
                                    }
                                }
#if defined(CONV_EXP) && (CONV_EXP & 8)
                                if (ov.x == 0x12345678u && p.S == -7)   // experiment: no output stores
#endif
                                *reinterpret_cast<uint4*>(ld.out + (static_cast<size_t>(chunk) * p.S + slot0 + 32 * cc + 8 * i) * 8) = ov;
                            }
                            if (TRAIN) {
                                if (has_res && cc < 3) load_res(cc + 1, res[0]);
                                if (bnb && cc < 3) load_zy(cc + 1);
                            }
                        }
                    }
                    if (TRAIN) {
                        if (want_stats || bnb) {   // the 8 lanes of a channel chunk hold different slots of the same 8 channels
#pragma unroll
                            for (int b = 0; b < 8; ++b) {
#pragma unroll
                                for (int o = 1; o < 8; o <<= 1) {
                                    st_sum[b] += __shfl_xor_sync(0xffffffffu, st_sum[b], o);
                                    st_sq[b] += __shfl_xor_sync(0xffffffffu, st_sq[b], o);
                                }
                            }
                            if (j == 0) {
#pragma unroll
                                for (int b = 0; b < 8; ++b) {
                                    atomicAdd(&s_stats[chunk * 8 + b], st_sum[b]);
                                    atomicAdd(&s_stats[128 + chunk * 8 + b], st_sq[b]);
                                }
                            }
                        }
                    }
                    // the stores above are read back by the bulk-copy engine (async proxy) when a later layer of THIS
                    // launch consumes them; the last layer's outputs are published by kernel completion
                    if (L + 1 < p.n_layers) asm volatile("fence.proxy.async;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(C3B_READY + jj));
                }
            }
        }
    }

    else if (warp == 10 && PF) {
        // ===================== masking warp (pad-free strip) + janitor =====================
        // For every stage: once the bulk copies have landed (FULL), write the two masked copies of the activation slab
        // -- rows of board column N-1 zeroed for the kx = -1 taps, rows of column 0 zeroed for the kx = +1 taps, the tile
        // remainder (rows >= N*PITCH) zeroed in both -- then publish them to the tensor core (fence.proxy.async) and
        // arrive on MASKED.  8 KiB read + 16 KiB written per slab: measured cost of that shared-memory traffic beside
        // the MMAs' operand fetches: +2.2 % per tile-layer, for 7 boards per tile instead of 6.
        // The L2 discards of dead activation tiles ride along one layer late: when slab 0 of (L, tile) is FULL the
        // producer has passed READY of (L-1, tile), so that item's input / residual are dead.
        uint32_t zl = 0, zr = 0;                         // bit i: row lane + 32 i is zero in the left / right copy
        const int used = p.n * p.pitch;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = lane + 32 * i;
            if (r >= used) {
                zl |= 1u << i;
                zr |= 1u << i;
            } else {
                const int x = (r % p.pitch) % p.bw;      // pad-free: bw == n
                if (x == p.n - 1) zl |= 1u << i;
                if (x == 0) zr |= 1u << i;
            }
        }
        uint32_t mcnt = 0;
        for (int g = 0, j0 = 0; g < walk.n_groups; j0 += walk.group_size(g), ++g) {
            const int gs = walk.group_size(g);
            for (int L = 0; L < p.n_layers; ++L) {
                const ConvLayerDesc& ld = p.layers[L];
                for (int jj = 0; jj < gs; ++jj) {
                    const int tile = tile0 + (j0 + jj) * int(gridDim.x);
                    for (int k = 0; k < ld.slabs; ++k, ++mcnt) {
                        const int sb = mcnt % STAGES;
                        mbar_wait(BAR(C3B_AFULL + sb), (mcnt / STAGES) & 1);
                        const uint32_t o = smem_u32(stage_buf + sb * STAGE_BYTES) + (C3_HALO + lane) * 16;
#if defined(CONV_EXP) && (CONV_EXP & 2048)
                        if (p.S == -7)                       // experiment: no masking traffic at all
#endif
#if defined(CONV_PF_BULK3)
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t at = o + kc * (C3_ROWS * 16) + i * 512;
                                if ((zl >> i) & 1)
                                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(at + C3_SLAB_BYTES), "r"(0u) : "memory");
                                if ((zr >> i) & 1)
                                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(at + 2 * C3_SLAB_BYTES), "r"(0u) : "memory");
                            }
                        }
#else
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t at = o + kc * (C3_ROWS * 16) + i * 512;
                                uint32_t a, b, c, d;
                                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(at));
                                const bool kl = !((zl >> i) & 1), kr = !((zr >> i) & 1);
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(at + C3_SLAB_BYTES), "r"(kl ? a : 0u),
                                             "r"(kl ? b : 0u), "r"(kl ? c : 0u), "r"(kl ? d : 0u) : "memory");
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(at + 2 * C3_SLAB_BYTES), "r"(kr ? a : 0u),
                                             "r"(kr ? b : 0u), "r"(kr ? c : 0u), "r"(kr ? d : 0u) : "memory");
                            }
                        }
#endif
#if !(defined(CONV_EXP) && (CONV_EXP & 1024))
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(C3B_MASKED + sb));
                        if (k == 0 && L > 0 && p.layers[L - 1].discard) {
                            const ConvLayerDesc& pd = p.layers[L - 1];
                            for (int idx = lane; idx < 16 * (C3_TILE_M / 8); idx += 32) {    // (chunk, 128-byte line of 8 slots)
                                const size_t off = (static_cast<size_t>(idx >> 5) * p.S + static_cast<size_t>(tile) * C3_TILE_M +
                                                    (idx & 31) * 8) * 8;
                                if (pd.discard & 1) l2_discard_128(pd.in + off);
                                if ((pd.discard & 2) && pd.res) l2_discard_128(pd.res + off);
                            }
                        }
                    }
                }
            }
        }
    }

    else if (warp == 10 && !TRAIN) {
        // ===================== janitor =====================
        // Dead activations: once the epilogue of a work item is done (READY), the tile's input (fully consumed by the
        // MMAs) and / or residual are never read again before they are rewritten, so their L2 lines are dropped instead
        // of being written back -- without this 70 % of the tower's intermediate activations reached HBM (1.6 GB per
        // launch at 5328 boards, 0.68 GB with it).  Issued from this otherwise idle warp: the same discards issued by
        // the epilogue warps made the (epilogue-bound) tower 5 % slower.  The warp follows every READY phase, flagged
        // layer or not, so that it is never more than one phase behind.
        uint32_t items_of[C3_GROUP] = {0, 0, 0};
        for (int g = 0, j0 = 0; g < walk.n_groups; j0 += walk.group_size(g), ++g) {
            const int gs = walk.group_size(g);
            for (int L = 0; L < p.n_layers; ++L) {
                const ConvLayerDesc& ld = p.layers[L];
                for (int jj = 0; jj < gs; ++jj) {
                    const int tile = tile0 + (j0 + jj) * int(gridDim.x);
                    mbar_wait(BAR(C3B_READY + jj), items_of[jj] & 1);
                    items_of[jj]++;
                    if (ld.discard) {
                        for (int idx = lane; idx < 16 * (C3_TILE_M / 8); idx += 32) {    // (chunk, 128-byte line of 8 slots)
                            const size_t off = (static_cast<size_t>(idx >> 5) * p.S + static_cast<size_t>(tile) * C3_TILE_M +
                                                (idx & 31) * 8) * 8;
                            if (ld.discard & 1) l2_discard_128(ld.in + off);
                            if ((ld.discard & 2) && ld.res) l2_discard_128(ld.res + off);
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    if (TRAIN) {
        // one flush of this CTA's batch statistics per launch (a training launch is ONE layer: the next layer's
        // BatchNorm needs the statistics of the whole batch first)
        double* gs = p.layers[p.n_layers - 1].stats;
        if (gs != nullptr && threadIdx.x < 256) {
            double v = double(s_stats[threadIdx.x]);
            if (p.bnb_y != nullptr && threadIdx.x >= 128) {   // sum g' * xhat = rstd * (sum g' * y - mean * sum g')
                const int c = threadIdx.x - 128;
                v = double(p.bnb_rstd[c]) * (v - double(p.bnb_mean[c]) * double(s_stats[c]));
            }
            atomicAdd(&gs[threadIdx.x], v);
        }
    }
}

// Host-side launch (layers 0..n_layers-1 over tiles [tile_begin, tile_end)). `stream` is the engine's stream.
template <bool TRAIN = false, bool PF = false>
inline cudaError_t conv3x3_tc3_launch(const ConvParams& p, int num_sms, cudaStream_t stream) {
    constexpr int SMEM = PF ? C3_PF_SMEM_BYTES : C3_SMEM_BYTES;
    if (PF != (p.bw == p.n)) return cudaErrorInvalidValue;   // the layout fields must match the kernel build
    // Set on every launch (a sub-microsecond host call): the kernel has internal linkage, so each translation unit that
    // includes this header owns its own copy of it, while a function-local `static bool` of this inline function would
    // be shared between them.
    {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc3_kernel<TRAIN, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM);
        if (e != cudaSuccess) return e;
    }
    const int tiles = p.tile_end - p.tile_begin;
    const int grid = tiles < num_sms ? tiles : num_sms;
    if (grid <= 0 || p.n_layers <= 0) return cudaSuccess;
    if (p.n_layers > C3_MAX_LAYERS) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C3_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see griddep_* in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv3x3_tc3_kernel<TRAIN, PF>, p);
}

// fill the layout fields of ConvParams for board size n (pad_free: the inference strip without pad columns)
inline void conv_params_set_layout(ConvParams& p, int n, bool pad_free = false) {
    p.n = n;
    p.bw = pad_free ? n : n + 1;
    p.bpt = 256 / (n * p.bw);
    p.pitch = p.bpt * p.bw;
}

}  // namespace tb
