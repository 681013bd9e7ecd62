// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (TMA engine, 1-D),
// tcgen05 (alloc / mma / commit / ld) and the UMMA descriptors used by conv_tc3.cuh.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables (cross-checked against the CuTe header
// cute/arch/mma_sm100_desc.hpp vendored in this image; nothing is included from it).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch error), never as a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("[tak_b200] mbarrier timeout: block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar,
                   parity);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------------------------------
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Tiled TMA load (tensor map in kernel-parameter space, `const __grid_constant__ CUtensorMap`): one instruction moves a
// 4-D box global -> shared, out-of-bounds elements arrive as zeros, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const void* tensor_map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            dst_smem),
        "l"(tensor_map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

// The 128-byte line at `p` will not be read again before it is overwritten: L2 may drop the dirty data instead of
// writing it back to HBM (a hint; reads after it return indeterminate data).
__device__ __forceinline__ void l2_discard_128(const void* p) {
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (lane_base+i), columns col..col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------------------
// K-major operand, no swizzle ("interleave"): in 16-byte units the canonical layout is
//   ((8 rows, n groups), 2 k-chunks) : ((1, SBO), LBO)
// i.e. a core matrix is 8 rows x 16 B stored contiguously (128 B); SBO = byte distance between
// consecutive 8-row groups, LBO = byte distance between the two 16-byte K chunks of one K=16 MMA.
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t smem_addr, uint32_t lbo_bytes,
                                                           uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);          // [0,14)  start address
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;    // [16,30) leading byte offset
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;    // [32,46) stride byte offset
    d |= static_cast<uint64_t>(1) << 46;                            // [46,48) descriptor version (Blackwell)
    // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
    return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(uint32_t M, uint32_t N) {
    return (1u << 4)            // [4,6)   D format  : 1 = f32
           | (1u << 7)          // [7,10)  A format  : 1 = bf16
           | (1u << 10)         // [10,13) B format  : 1 = bf16
           | ((N >> 3) << 17)   // [17,23) N >> 3
           | ((M >> 4) << 24);  // [24,29) M >> 4
}

}  // namespace tb

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL)
// ----------------------------------------------------------------------------------------------
namespace tb {

__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace tb
