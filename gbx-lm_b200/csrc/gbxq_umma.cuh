// gbxq_umma.cuh -- tcgen05 / TMEM / TMA PTX wrappers and the register-level dequantisers shared by the tensor-core
// kernels (sm_100a).  Everything is inline device code in namespace gbxq::umma.
#pragma once
#include <cuda.h>

#include "gbxq_common.cuh"

namespace gbxq {
namespace umma {

constexpr uint32_t kMagic = 0x43004300u;  // bf16x2 (128.0, 128.0): q | magic = 128 + q exactly

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst_smem, const void* tmap, int c0, int c1, uint32_t bar_smem) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar_smem)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D(TMEM) (+)= A(smem descriptor) * B(smem descriptor)
__device__ __forceinline__ void mma_ss_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// D(TMEM) (+)= A(TMEM: lane = row, 32-bit column = two consecutive k) * B(smem descriptor)
__device__ __forceinline__ void mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// the same with the shared-memory descriptor as two 32-bit halves (the issue loop keeps them in uniform registers and
// steps the low half by the K offset)
__device__ __forceinline__ void mma_ts_f16_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                                uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 bd;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 bd, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accum)
        : "memory");
}
// one lane of a converged warp (the others skip): keeps the surrounding control flow warp-uniform, so the compiler
// holds addresses / descriptors in uniform registers instead of broadcasting them lane by lane
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// registers -> TMEM: 16 consecutive 32-bit columns of the thread's lane (lane quarter of the warp: 32 * (warp % 4))
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM -> registers, NC (4, 8, 16, 32) consecutive columns of the thread's lane; waits for the data
template <int NC> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[32]) {
    if constexpr (NC == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
    } else if constexpr (NC == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr)
            : "memory");
    } else if constexpr (NC == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr)
                     : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                     : "r"(taddr)
                     : "memory");
    }
    tmem_wait_ld();
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): rows of 128 bytes, 8-row core groups 1024
// bytes apart (SBO); LBO unused for a swizzled K-major tile one atom wide.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (16 B; ignored) bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B         bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)         bits [46,48)
    d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B           bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = m, N = n.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t hsub2_bf16(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hfma2_bf16(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// Codes 8c .. 8c+7 of a run of packed codes `w` (LSB-first bitstream starting at bit 0 of w[0]) as 4 bf16x2 registers in
// natural k order (low half = even k): scale * q + bias with ONE rounding (bf16 FMA; fp32 FMA + cvt for 3/6/8-bit).
// sraw / braw: the group's bf16 scale / bias bits in the low half.
template <int BITS, int NW>
__device__ __forceinline__ void dequant8(const uint32_t (&w)[NW], int c, uint32_t sraw, uint32_t braw, uint32_t (&out)[4]) {
    if constexpr (BITS == 4 || BITS == 2) {
        const uint32_t s2 = sraw | (sraw << 16), b2 = braw | (braw << 16);
        uint32_t pr[4];  // pr[i] = (q_i, q_{i+4}) exact small integers in bf16
        if constexpr (BITS == 4) {
            const uint32_t word = w[c];
#pragma unroll
            for (int i = 0; i < 4; i++) pr[i] = hsub2_bf16(lop3_and_or(word >> (4 * i), 0x000f000fu, kMagic), kMagic);
        } else {
            const uint32_t half = w[c >> 1] >> (16 * (c & 1));  // the chunk's 8 fields at bits 0..15
#pragma unroll
            for (int i = 0; i < 4; i++)  // (field i, field i+4): bits 2i and 2i+8 -> low bits of each half
                pr[i] = hsub2_bf16(lop3_and_or(__byte_perm(half >> (2 * i), 0u, 0x4140), 0x00030003u, kMagic), kMagic);
        }
        // re-pair (0,4),(1,5),(2,6),(3,7) -> (0,1),(2,3),(4,5),(6,7)
        const uint32_t q01 = __byte_perm(pr[0], pr[1], 0x5410), q45 = __byte_perm(pr[0], pr[1], 0x7632);
        const uint32_t q23 = __byte_perm(pr[2], pr[3], 0x5410), q67 = __byte_perm(pr[2], pr[3], 0x7632);
        out[0] = hfma2_bf16(q01, s2, b2);
        out[1] = hfma2_bf16(q23, s2, b2);
        out[2] = hfma2_bf16(q45, s2, b2);
        out[3] = hfma2_bf16(q67, s2, b2);
    } else {
        const float s = __uint_as_float(sraw << 16), b = __uint_as_float(braw << 16);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int bit = (8 * c + i) * BITS;
            const int wi = bit >> 5, off = bit & 31;
            uint32_t sh;
            if (off + BITS <= 32) sh = w[wi] >> off;
            else sh = __funnelshift_r(w[wi], w[wi + 1 < NW ? wi + 1 : wi], off);
            const float q = __uint_as_float(lop3_and_or(sh, (1u << BITS) - 1u, 0x4B000000u)) - 8388608.0f;
            v[i] = fmaf(s, q, b);
        }
        out[0] = pack_bf16x2(v[0], v[1]);
        out[1] = pack_bf16x2(v[2], v[3]);
        out[2] = pack_bf16x2(v[4], v[5]);
        out[3] = pack_bf16x2(v[6], v[7]);
    }
}

}  // namespace umma
}  // namespace gbxq
