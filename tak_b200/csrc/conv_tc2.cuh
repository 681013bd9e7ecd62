// 3x3 conv of the residual tower, CTA-pair version: tcgen05.mma.cta_group::2 with the layer's weights RESIDENT in the
// shared memory of the pair (each CTA holds the 64 output channels of its half, 144 KiB), so the only streaming
// operand is the activation tile.  Same slot-plane data layout and epilogue semantics as conv_tc.cuh (see there).
//
// Per pair-tile of 256 slots: CTA r owns rows [128 r, 128 r + 128): its A tile (+16-row halos) is double buffered
// (2 x 40 KiB), the leader CTA issues 72 MMAs (M=256, N=128, K=16) whose B operand is read half from each CTA, and
// each CTA's 128x128 fp32 accumulator lands in its own TMEM (4 stages x 128 columns).
//
// Warp roles per CTA (352 threads):
//   warp 0   activation-tile producer (cp.async.bulk -> a_full)
//   warp 1   TMEM alloc; in the leader CTA: single-thread MMA issuer (commit multicast -> acc_full / a_empty of both)
//   warps 2-9 epilogue: lane quarter = warp%4, column half = (warp-2)/4; residual prefetch, pipelined tcgen05.ld
//   warp 10  weight loader (once per launch) + forwarder: local a_full / w_full -> leader's a_ready / w_ready
#pragma once
#include "conv_tc.cuh"

namespace tb {

constexpr int C2_TILE_M = 128;
constexpr int C2_ROWS = C2_TILE_M + 2 * CONV_HALO_ROWS;     // 160
constexpr int C2_A_BYTES = C2_ROWS * 16 * CONV_CHUNKS;      // 40960
constexpr int C2_W_STAGE_BYTES = 8 * 64 * 16;               // 8192: 64 c_in x 64 c_out
constexpr int C2_W_BYTES = CONV_STAGES_PER_LAYER * C2_W_STAGE_BYTES;  // 147456 per CTA
constexpr int C2_ACC_STAGES = 4;
constexpr int C2_THREADS = 352;
constexpr int C2_SMEM_BYTES = C2_W_BYTES + 2 * C2_A_BYTES + 1024;

// barrier slots (identical offsets in both CTAs)
enum : int {
    C2B_A_FULL = 0,     // [2] local TMA completion
    C2B_A_EMPTY = 2,    // [2] multicast commit
    C2B_A_READY = 4,    // [2] leader: 2 arrivals
    C2B_ACC_FULL = 6,   // [4] multicast commit
    C2B_ACC_EMPTY = 10, // [4] leader: 16 arrivals (8 epilogue warps x 2 CTAs)
    C2B_W_FULL = 14,    // [9] local TMA completion, one per tap (2 stages = 16 KiB)
    C2B_W_READY = 23,   // [9] leader: 2 arrivals per tap
    C2B_COUNT = 32
};

template <int MODE>
static __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C2_THREADS, 1)
    conv3x3_tc2_kernel(const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_buf = smem;                       // C2_W_BYTES
    uint8_t* a_buf = smem + C2_W_BYTES;          // 2 x C2_A_BYTES
    uint8_t* tail = a_buf + 2 * C2_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + C2B_COUNT * 8);
    float* s_bias = reinterpret_cast<float*>(tail + C2B_COUNT * 8 + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
#if defined(CONV_EXP) && (CONV_EXP & 16)
    const int pair = (blockIdx.x >> 1) + 1000000;  // experiment: empty kernel (no tiles)
#else
    const int pair = blockIdx.x >> 1;
#endif
    const int n_pairs = gridDim.x >> 1;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(C2B_A_FULL + i), 1);
            mbar_init(BAR(C2B_A_EMPTY + i), 1);
            mbar_init(BAR(C2B_A_READY + i), 2);
        }
        for (int i = 0; i < 9; ++i) {
            mbar_init(BAR(C2B_W_FULL + i), 1);
            mbar_init(BAR(C2B_W_READY + i), 2);
        }
        for (int i = 0; i < C2_ACC_STAGES; ++i) {
            mbar_init(BAR(C2B_ACC_FULL + i), 1);
            mbar_init(BAR(C2B_ACC_EMPTY + i), 16);
        }
        mbar_fence_init();
    }
    // Programmatic dependent launch: let the next layer's CTAs be scheduled as soon as SMs free up; everything
    // before griddep_wait() below (barrier init, TMEM alloc, weight load) overlaps the previous layer's tail.
    griddep_launch_dependents();
    if (threadIdx.x < 128) s_bias[threadIdx.x] = p.bias[threadIdx.x];
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
    if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const size_t plane_bytes = static_cast<size_t>(p.S) * 16;

    if (warp == 0) {
        // ===================== activation-tile producer (both CTAs) =====================
        if (lane == 0) {
            griddep_wait();  // activations are written by the previous layer
            int it = 0;
            for (int tile = pair; tile < p.tiles; tile += n_pairs, ++it) {
                const int ab = it & 1;
                const uint32_t aph = (it >> 1) & 1;
#if defined(CONV_EXP) && (CONV_EXP & 4)
                if (it >= 2) break;
#endif
                mbar_wait_cluster(BAR(C2B_A_EMPTY + ab), aph ^ 1);
                mbar_expect_tx(BAR(C2B_A_FULL + ab), C2_A_BYTES);
                const int row0 = CONV_GUARD + tile * CONV_TILE_M + int(rank) * C2_TILE_M - CONV_HALO_ROWS;  // >= 0
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.in) + static_cast<size_t>(row0) * 16;
                const uint32_t dst = smem_u32(a_buf + ab * C2_A_BYTES);
                for (int c = 0; c < CONV_CHUNKS; ++c)
                    bulk_g2s(dst + c * (C2_ROWS * 16), src + c * plane_bytes, C2_ROWS * 16, BAR(C2B_A_FULL + ab));
            }
        }
    } else if (warp == 10) {
        // ===================== weight loader + forwarder (both CTAs) =====================
        if (lane == 0) {
            // weights never depend on the previous kernel: load them right away, one barrier per tap
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + static_cast<size_t>(rank) * C2_W_BYTES;
            for (int tap = 0; tap < 9; ++tap) {
                mbar_expect_tx(BAR(C2B_W_FULL + tap), 2 * C2_W_STAGE_BYTES);
                bulk_g2s(smem_u32(w_buf + tap * 2 * C2_W_STAGE_BYTES),
                         wsrc + static_cast<size_t>(tap) * 2 * C2_W_STAGE_BYTES, 2 * C2_W_STAGE_BYTES,
                         BAR(C2B_W_FULL + tap));
            }
            for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(BAR(C2B_W_FULL + tap), 0);
                mbar_arrive_remote(BAR(C2B_W_READY + tap), 0);
            }
            int it = 0;
            for (int tile = pair; tile < p.tiles; tile += n_pairs, ++it) {
                const int ab = it & 1;
#if defined(CONV_EXP) && (CONV_EXP & 4)
                if (it >= 2) break;
#endif
                mbar_wait(BAR(C2B_A_FULL + ab), (it >> 1) & 1);
                mbar_arrive_remote(BAR(C2B_A_READY + ab), 0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0 && lane == 0) {
#if defined(CONV_EXP) && (CONV_EXP & 128)
            constexpr uint32_t idesc = umma_idesc_bf16_f32(256, 256);  // experiment: issue-rate probe (garbage B)
#elif defined(CONV_EXP) && (CONV_EXP & 256)
            constexpr uint32_t idesc = umma_idesc_bf16_f32(256, 64);
#else
            constexpr uint32_t idesc = umma_idesc_bf16_f32(256, 128);
#endif
            int it = 0;
            for (int tile = pair; tile < p.tiles; tile += n_pairs, ++it) {
                const int ab = it & 1;
                const int as = it & (C2_ACC_STAGES - 1);
                mbar_wait_cluster(BAR(C2B_ACC_EMPTY + as), ((it >> 2) & 1) ^ 1);  // drained by both epilogues
#if defined(CONV_EXP) && (CONV_EXP & 4)
                if (it < 2)
#endif
                mbar_wait_cluster(BAR(C2B_A_READY + ab), (it >> 1) & 1);          // both A tiles landed
                tc_fence_after();
                const uint32_t a_base = smem_u32(a_buf + ab * C2_A_BYTES);
#if defined(CONV_EXP) && (CONV_EXP & 128)
                const uint32_t d_addr = tmem_base + (as & 1) * 256;
#else
                const uint32_t d_addr = tmem_base + as * 128;
#endif
#pragma unroll 1
                for (int st = 0; st < CONV_STAGES_PER_LAYER; ++st) {
                    const int tap = st >> 1, half = st & 1;
                    if (it == 0 && half == 0) {  // first tile: consume the weights tap by tap as they land
                        mbar_wait_cluster(BAR(C2B_W_READY + tap), 0);
                        tc_fence_after();
                    }
#if defined(CONV_EXP) && (CONV_EXP & 32)
                    const int shift = 0;  // experiment: every tap reads the 128 B-aligned window
#elif defined(CONV_EXP) && (CONV_EXP & 64)
                    const int shift = (tap / 3 - 1) * 8;  // experiment: aligned but distinct windows
#else
                    const int shift = (tap / 3 - 1) * p.pitch + (tap % 3 - 1);
#endif
                    const uint32_t w_base = smem_u32(w_buf + st * C2_W_STAGE_BYTES);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int chunk = half * 8 + j * 2;
                        const uint32_t a_addr = a_base + chunk * (C2_ROWS * 16) + (CONV_HALO_ROWS + shift) * 16;
                        const uint64_t adesc = umma_desc_kmajor_noswz(a_addr, C2_ROWS * 16, 128);
                        const uint64_t bdesc = umma_desc_kmajor_noswz(w_base + j * 2 * (64 * 16), 64 * 16, 128);
                        umma2_bf16(d_addr, adesc, bdesc, idesc, (st | j) != 0);
                    }
                }
                umma2_commit_mc(BAR(C2B_ACC_FULL + as), 3);  // accumulators ready in both CTAs
                umma2_commit_mc(BAR(C2B_A_EMPTY + ab), 3);   // both activation buffers free
            }
        }
    } else if (warp >= 2 && warp < 10) {
        // ===================== epilogue (8 warps: 4 lane quarters x 2 column halves) =====================
        const int quarter = warp & 3;
        const int colhalf = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        griddep_wait();  // residual input / output buffers belong to earlier layers until they complete
        int it = 0;
        for (int tile = pair; tile < p.tiles; tile += n_pairs, ++it) {
            const int as = it & (C2_ACC_STAGES - 1);
            const int slot = CONV_GUARD + tile * CONV_TILE_M + int(rank) * C2_TILE_M + row;
            const bool valid = conv_slot_valid(slot, p.pitch, p.n_boards);
            uint4 res[8];
            if (MODE == CONV_RES_RELU) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    res[c] = make_uint4(0, 0, 0, 0);
                    if (valid)
                        res[c] = __ldg(reinterpret_cast<const uint4*>(
                            p.res + (static_cast<size_t>(colhalf * 8 + c) * p.S + slot) * 8));
                }
            }
            mbar_wait_cluster(BAR(C2B_ACC_FULL + as), (it >> 2) & 1);
            tc_fence_after();
#if defined(CONV_EXP) && (CONV_EXP & 2)
            if (p.S != -12345) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive_remote(BAR(C2B_ACC_EMPTY + as), 0); continue; }
#endif
            const uint32_t taddr = tmem_base + as * 128 + colhalf * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
            uint32_t r[2][32];
            tmem_ld32(taddr, r[0]);
            tmem_ld32(taddr + 32, r[1]);
            tmem_ld_wait();
            // the accumulator stage can be recycled as soon as it sits in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(BAR(C2B_ACC_EMPTY + as), 0);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const uint32_t(&rc)[32] = r[cc];
                if (MODE == CONV_LOGITS_F32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int ch = colhalf * 64 + cc * 32 + j;
                        if (ch < p.out_ch_valid) {
                            float v = __uint_as_float(rc[j]) + s_bias[ch];
                            p.out_f32[static_cast<size_t>(p.out_ch_offset + ch) * p.S + slot] = valid ? v : 0.0f;
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int lc = cc * 4 + q;               // chunk within this column half
                        const int chunk = colhalf * 8 + lc;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(rc[q * 8 + j]) + s_bias[chunk * 8 + j];
                        if (MODE == CONV_RES_RELU) {
                            const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&res[lc]);
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
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still touch its smem / TMEM
    if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

inline cudaError_t conv3x3_tc2_launch(const ConvParams& p, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc2_kernel<CONV_RELU>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv3x3_tc2_kernel<CONV_RES_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C2_SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv3x3_tc2_kernel<CONV_LOGITS_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C2_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int pairs = num_sms / 2;
    if (p.tiles < pairs) pairs = p.tiles;
    if (pairs <= 0) return cudaSuccess;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pairs * 2);
    cfg.blockDim = dim3(C2_THREADS);
    cfg.dynamicSmemBytes = C2_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see griddep_* in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.mode == CONV_RELU) return cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<CONV_RELU>, p);
    if (p.mode == CONV_RES_RELU) return cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<CONV_RES_RELU>, p);
    return cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel<CONV_LOGITS_F32>, p);
}

}  // namespace tb
