// gbxq_gemm_sm100.cu -- prefill / large-batch quantized GEMM on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM, activations by TMA) with in-kernel dequantisation.
//
// Reference call: QuantizedLinear.__call__ -> mx.quantized_matmul(transpose=True)
// (gbx_lm/models/quantized_linear_gba.py:195-203) with M = B*L rows of x; the reference prefills
// in chunks of up to 2048 tokens (gbx_lm/utils.py:226,312-319).  Above the ridge point
// (M ~ 40-140) the path is a dense contraction bounded by the tensor pipe.
//
// Swap-AB formulation: the WEIGHT tile is the MMA "A" operand (M_mma = 128 output features fill the
// 128 TMEM lanes) and the TOKENS are the MMA N dimension (BN = 32..256 accumulator columns):
//     D[n, m] (TMEM, fp32) += A[n, k] (bf16 dequantised weights, smem) * B[k, m] (bf16 x, smem)
// so one dequantised 128 x 64 weight tile feeds up to 256 tokens, and skinny batches (M = 17..64)
// use the same kernel with a narrow N.
//
// Warp roles (320 threads, one CTA per SM, 4-stage ring, BLOCK_K = 64):
//   warp 0      TMA producer: x tile [BN tokens x 64 k] via cp.async.bulk.tensor.2d, SWIZZLE_128B,
//               out-of-range tokens zero-filled by the TMA unit; also owns the TMEM allocation
//   warps 1-8   dequant producers: a thread owns half a weight row of the tile per k-block: loads its 32
//               packed codes (register ring, prefetched three k-blocks ahead), unpacks with (w >> s) & mask | 0x4300 (two bf16
//               128+q per LOP3), applies scale*q+bias with one HFMA2.BF16 per pair (single rounding),
//               and stores the row into the canonical K-major 128B-swizzled layout the UMMA
//               descriptor expects (16-byte chunk c of row r at chunk position c ^ (r & 7));
//               fence.proxy.async, then arrive.  After the main loop the same warps are the epilogue:
//               tcgen05.ld 32x32b -> one rounding to bf16 (+ bias as a second rounded add) -> y
//   warp 9      MMA issuer: one thread issues 4 x tcgen05.mma (K = 16) per stage and commits the
//               stage's smem release and, at the end, the accumulator hand-off to the epilogue
#include <cuda.h>

#include "gbxq_common.cuh"

namespace gbxq {

namespace {

constexpr int kBlockK = 64;     // bf16 elements per k-block = one 128-byte swizzle atom
constexpr int kTileN = 128;     // output features per CTA (= UMMA M = TMEM lanes)
constexpr int kStages = 4;
constexpr int kThreads = 320;
constexpr int kDequantWarps = 8;
constexpr uint32_t kMagic = 0x43004300u;

// ------------------------------------------------------------------ tcgen05 / TMA PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): rows of 128 bytes, 8-row
// core groups 1024 bytes apart (SBO), LBO unused for a swizzled K-major tile one atom wide.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (16 B; ignored) bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B         bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)         bits [46,48)
    d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B           bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N = bn.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmParams {
    const uint32_t* w;
    const uint16_t* scales;
    const uint16_t* biases;
    const __nv_bfloat16* bias;
    __nv_bfloat16* y;
    int64_t M, N, K;
    int gs_shift;
    int G;
    int words_per_row;
};

// 32 consecutive codes of one row (half h of k-block kb) -> 32 bf16 weights (4 chunks of 8) = scale*q + bias
template <int BITS> struct RowBlock {
    static constexpr int NW = BITS;  // 32-bit words holding 32 codes
    uint32_t w[NW];
    uint32_t s, b;                   // raw bf16 scale / bias of the group the 32 codes belong to
};

// hb = index of the 32-code half-block along the row (= 2*kb + h)
template <int BITS>
__device__ __forceinline__ void load_rowblock(RowBlock<BITS>& rb, const uint32_t* __restrict__ row,
                                              const uint16_t* __restrict__ srow, const uint16_t* __restrict__ brow, int hb,
                                              int gs_shift, bool valid) {
    const uint32_t* src = row + (size_t)hb * BITS;
    if (!valid) {
#pragma unroll
        for (int i = 0; i < BITS; i++) rb.w[i] = 0u;
        rb.s = rb.b = 0u;
        return;
    }
    if constexpr (BITS == 4 || BITS == 8) {
#pragma unroll
        for (int i = 0; i < BITS / 4; i++) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + i);
            rb.w[4 * i] = t.x; rb.w[4 * i + 1] = t.y; rb.w[4 * i + 2] = t.z; rb.w[4 * i + 3] = t.w;
        }
    } else if constexpr (BITS == 2 || BITS == 6) {
#pragma unroll
        for (int i = 0; i < BITS / 2; i++) {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(src) + i);
            rb.w[2 * i] = t.x; rb.w[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < BITS; i++) rb.w[i] = __ldg(src + i);
    }
    const int g = (hb * 32) >> gs_shift;
    rb.s = (uint32_t)__ldg(srow + g);
    rb.b = (uint32_t)__ldg(brow + g);
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

// chunk c (codes 8c .. 8c+7 of the 32-code half-block) as 4 packed bf16x2 registers in natural k order.
template <int BITS>
__device__ __forceinline__ uint4 dequant_chunk(const RowBlock<BITS>& rb, int c) {
    const uint32_t sraw = rb.s, braw = rb.b;
    uint4 out;
    if constexpr (BITS == 4 || BITS == 2) {
        const uint32_t s2 = sraw | (sraw << 16), b2 = braw | (braw << 16);
        uint32_t pr[4];  // pr[i] = (q_a, q_b) exact small integers in bf16
        if constexpr (BITS == 4) {
            const uint32_t word = rb.w[c];
#pragma unroll
            for (int i = 0; i < 4; i++)  // (nibble i, nibble i+4)
                pr[i] = hsub2_bf16(lop3_and_or(word >> (4 * i), 0x000f000fu, kMagic), kMagic);
            // re-pair (0,4),(1,5),(2,6),(3,7) -> (0,1),(2,3),(4,5),(6,7)
            const uint32_t q01 = __byte_perm(pr[0], pr[1], 0x5410), q45 = __byte_perm(pr[0], pr[1], 0x7632);
            const uint32_t q23 = __byte_perm(pr[2], pr[3], 0x5410), q67 = __byte_perm(pr[2], pr[3], 0x7632);
            out.x = hfma2_bf16(q01, s2, b2); out.y = hfma2_bf16(q23, s2, b2);
            out.z = hfma2_bf16(q45, s2, b2); out.w = hfma2_bf16(q67, s2, b2);
        } else {
            // 2-bit: word holds 16 codes; chunk c uses fields 8*(c&1) .. +7 of word c>>1
            const uint32_t half = rb.w[c >> 1] >> (16 * (c & 1));  // fields f .. f+7 now at bits 0..15
#pragma unroll
            for (int i = 0; i < 4; i++)  // (field i, field i+4) of the 8: bits 2i and 2i+8
                pr[i] = hsub2_bf16(lop3_and_or(__byte_perm(half >> (2 * i), 0u, 0x4140), 0x00030003u, kMagic), kMagic);
            const uint32_t q01 = __byte_perm(pr[0], pr[1], 0x5410), q45 = __byte_perm(pr[0], pr[1], 0x7632);
            const uint32_t q23 = __byte_perm(pr[2], pr[3], 0x5410), q67 = __byte_perm(pr[2], pr[3], 0x7632);
            out.x = hfma2_bf16(q01, s2, b2); out.y = hfma2_bf16(q23, s2, b2);
            out.z = hfma2_bf16(q45, s2, b2); out.w = hfma2_bf16(q67, s2, b2);
        }
    } else {
        // generic widths (3, 6, 8): exact small integers via the fp32 magic number, fp32 FMA, one rounding
        const float s = __uint_as_float(sraw << 16), b = __uint_as_float(braw << 16);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int bit = (8 * c + i) * BITS;
            const int wi = bit >> 5, off = bit & 31;
            uint32_t sh;
            if (off + BITS <= 32) sh = rb.w[wi] >> off;
            else sh = __funnelshift_r(rb.w[wi], rb.w[wi + 1], off);
            const float q = __uint_as_float(lop3_and_or(sh, (1u << BITS) - 1u, 0x4B000000u)) - 8388608.0f;
            v[i] = fmaf(s, q, b);
        }
        out.x = pack_bf16x2(v[0], v[1]); out.y = pack_bf16x2(v[2], v[3]);
        out.z = pack_bf16x2(v[4], v[5]); out.w = pack_bf16x2(v[6], v[7]);
    }
    return out;
}

template <int BITS, int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p) {
    constexpr uint32_t A_BYTES = kTileN * kBlockK * 2;  // 16 KB
    constexpr uint32_t B_BYTES = BN * kBlockK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * STAGE_BYTES);
    uint64_t* full_b = full_a + kStages;
    uint64_t* empty = full_b + kStages;
    uint64_t* tmem_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kTileN;
    const int m0 = blockIdx.y * BN;
    const int nkb = (int)(p.K / kBlockK);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_a[s], kDequantWarps);  // one arrival per dequant warp
            mbar_init(&full_b[s], 1);   // TMA producer (+ tx bytes)
            mbar_init(&empty[s], 1);    // tcgen05.commit
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (x tile) =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&empty[s], phase ^ 1u);
                mbar_arrive_expect_tx(&full_b[s], B_BYTES);
                tma_load_2d(smem + (size_t)s * STAGE_BYTES + A_BYTES, &tmap_x, kb * kBlockK, m0, &full_b[s]);
            }
        }
    } else if (warp == 1 + kDequantWarps) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kTileN, BN);
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % kStages;
                const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&full_a[s], phase);
                mbar_wait(&full_b[s], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; k++) {
                    const uint64_t ad = make_sw128_kmajor_desc(a_addr + k * 32);
                    const uint64_t bd = make_sw128_kmajor_desc(b_addr + k * 32);
                    umma_f16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
            }
            umma_commit(tmem_full);      // accumulator complete -> epilogue
        }
    } else {
        // ===================== dequant producers (warps 1-8), then epilogue =====================
        const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
        const int h = (warp - 1) >> 2;           // which 32-code half of the k-block / which half of the columns
        const int r = q4 * 32 + lane;            // tile row == TMEM lane == output feature n0 + r
        const int64_t n = (int64_t)n0 + r;
        const bool row_ok = n < p.N;
        const uint32_t* wrow = p.w + (size_t)(row_ok ? n : 0) * p.words_per_row;
        const uint16_t* srow = p.scales + (size_t)(row_ok ? n : 0) * p.G;
        const uint16_t* brow = p.biases + (size_t)(row_ok ? n : 0) * p.G;

        // register ring of packed half-blocks, fetched three k-blocks ahead (global latency ~ 3 k-blocks of MMA);
        // the loop is unrolled by 4 so that every ring slot has a fixed register name (no copies that would
        // wait for the loads in flight)
        RowBlock<BITS> ring[4];
        auto fetch = [&](int kb, RowBlock<BITS>& rb) { load_rowblock<BITS>(rb, wrow, srow, brow, 2 * kb + h, p.gs_shift, row_ok && kb < nkb); };
        auto produce = [&](int kb, const RowBlock<BITS>& rb) {
            const int s = kb % kStages;
            const uint32_t phase = (uint32_t)(kb / kStages) & 1u;
            mbar_wait(&empty[s], phase ^ 1u);
            uint8_t* a_row = smem + (size_t)s * STAGE_BYTES + (size_t)r * 128;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint4 v = dequant_chunk<BITS>(rb, c);
                *reinterpret_cast<uint4*>(a_row + (((4 * h + c) ^ (r & 7)) << 4)) = v;
            }
            fence_proxy_async();  // make the generic-proxy stores visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[s]);
        };
        fetch(0, ring[0]);
        fetch(1, ring[1]);
        fetch(2, ring[2]);
        for (int kb = 0; kb < nkb;) {
            fetch(kb + 3, ring[3]);
            produce(kb, ring[0]);
            if (++kb >= nkb) break;
            fetch(kb + 3, ring[0]);
            produce(kb, ring[1]);
            if (++kb >= nkb) break;
            fetch(kb + 3, ring[1]);
            produce(kb, ring[2]);
            if (++kb >= nkb) break;
            fetch(kb + 3, ring[2]);
            produce(kb, ring[3]);
            ++kb;
        }

        // ---- epilogue: TMEM -> registers -> bf16 -> y[m, n]; warp (q4, h) drains lanes 32*q4.. and half h of the columns
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const float bias_f = (p.bias != nullptr && row_ok) ? __bfloat162float(p.bias[n]) : 0.f;
        constexpr int HALF = BN / 2;
        constexpr int STEP = HALF >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += STEP) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0;
            if constexpr (STEP == 32) tmem_ld32(taddr, v);
            else tmem_ld16(taddr, v);
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < STEP; j++) {
                    const int64_t m = (int64_t)m0 + c0 + j;
                    if (m < p.M) {
                        float f = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j])));
                        if (p.bias != nullptr) f = __fadd_rn(f, bias_f);
                        p.y[(size_t)m * p.N + n] = __float2bfloat16_rn(f);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

template <int BITS, int BN>
int launch_inst(const CUtensorMap& tmap, const GemmParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)kStages * (kTileN * kBlockK * 2 + BN * kBlockK * 2) + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BITS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e);
        configured = true;
    }
    dim3 grid((unsigned)((p.N + kTileN - 1) / kTileN), (unsigned)((p.M + BN - 1) / BN));
    gemm_kernel<BITS, BN><<<grid, kThreads, smem, st>>>(tmap, p);
    count_launch();
    return check_cuda(cudaGetLastError());
}

template <int BITS>
int launch_bn(int bn, const CUtensorMap& tmap, const GemmParams& p, cudaStream_t st) {
    switch (bn) {
        case 32: return launch_inst<BITS, 32>(tmap, p, st);
        case 64: return launch_inst<BITS, 64>(tmap, p, st);
        case 128: return launch_inst<BITS, 128>(tmap, p, st);
        default: return launch_inst<BITS, 256>(tmap, p, st);
    }
}

int pick_bn(int64_t M) { return M <= 32 ? 32 : (M <= 64 ? 64 : (M <= 128 ? 128 : 256)); }

}  // namespace

bool gemm_supported(int64_t M, int64_t N, int64_t K, int bits, int gs, int dtype, const void* x, const void* w,
                    const void* y) {
    (void)gs;
    (void)bits;
    if (dtype != GBXQ_BF16 || M < 1 || N < 1) return false;
    if (K % kBlockK) return false;
    if (((uintptr_t)x | (uintptr_t)w) & 15) return false;
    if ((uintptr_t)y & 1) return false;
    if (M > (int64_t)1 << 24 || (N + kTileN - 1) / kTileN > 65535 * 32) return false;
    return get_encode() != nullptr;
}

int launch_gemm(const void* x, const uint32_t* w, const void* s, const void* b, const void* bias, void* y, int64_t M,
                int64_t N, int64_t K, int bits, int gs, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return GBXQ_EUNSUPPORTED;
    const int bn = pick_bn(M);
    if ((M + bn - 1) / bn > 65535) return GBXQ_EUNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return GBXQ_EUNSUPPORTED;
    GemmParams p{};
    p.w = w;
    p.scales = reinterpret_cast<const uint16_t*>(s);
    p.biases = reinterpret_cast<const uint16_t*>(b);
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.M = M;
    p.N = N;
    p.K = K;
    p.gs_shift = gs == 32 ? 5 : (gs == 64 ? 6 : 7);
    p.G = (int)(K / gs);
    p.words_per_row = (int)(K * bits / 32);
    switch (bits) {
        case 2: return launch_bn<2>(bn, tmap, p, st);
        case 3: return launch_bn<3>(bn, tmap, p, st);
        case 4: return launch_bn<4>(bn, tmap, p, st);
        case 6: return launch_bn<6>(bn, tmap, p, st);
        case 8: return launch_bn<8>(bn, tmap, p, st);
    }
    return GBXQ_EINVAL_BITS;
}

}  // namespace gbxq
