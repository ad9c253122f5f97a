// bx_tcgen05.cuh -- PTX wrappers shared by the tensor-core convolution kernels (bx_conv_tc.cu: TF32 operands, A through
// tensor memory; bx_conv_sd.cu: fp16-split operands, A through shifted shared-memory descriptors): mbarriers, bulk copies,
// tcgen05.mma / commit / ld / st.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Issue helpers for the MMA warp.  The whole warp runs the issue loop in uniform control flow; `leader`
// is 1 in exactly one lane (elect.sync) and predicates the tcgen05 instructions themselves, so the compiler
// does not have to wrap every asm statement in a divergence loop.  The 64-bit shared-memory descriptor is
// passed as its two 32-bit halves: lo = (addr >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14.
static __device__ __forceinline__ uint32_t elect_leader() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred;
}

// A operand from tensor memory (128 lanes x 8 tf32 columns), B operand from shared memory
static __device__ __forceinline__ void mma_tf32_ts(uint32_t leader, uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %0, 0;\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], db, %5, p;\n\t"
        "}\n" ::"r"(leader),
        "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

static __device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

static __device__ __forceinline__ void mma_commit(uint32_t leader, uint32_t bar_saddr) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %0, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t"
        "}\n" ::"r"(leader),
        "r"(bar_saddr)
        : "memory");
}

static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

static __device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

static __device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

static __device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Non-blocking probe of a barrier phase (1 = completed).  The MMA warp issues it BEFORE a batch of MMAs and looks at the result
// after them: a completed mbarrier.try_wait still costs the warp ~100 cycles of round trip, which the batch then hides; only
// a phase that is not complete falls back to mbar_wait.
static __device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t r;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(r)
        : "r"(bar), "r"(parity)
        : "memory");
    return r;
}

// tcgen05.ld 32 lanes x CW consecutive 32-bit columns (CW = 8, 16 or 32) into v[0..CW)
template <int CW>
static __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[CW]) {
    if constexpr (CW == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr)
                     : "memory");
    } else if constexpr (CW == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr)
            : "memory");
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
    }
}

