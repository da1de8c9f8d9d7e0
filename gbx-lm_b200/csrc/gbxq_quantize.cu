// gbxq_quantize.cu -- mx.quantize on the device: group-affine quantisation of fresh activations into the packed layout
// the matmul kernels read.  Reference callers: the quantized KV cache quantises the new keys / values of every step
// (gbx_lm/models/cache.py:176-177, QuantizedKVCache.update_and_fetch) before
// quantized_scaled_dot_product_attention reads them (gbx_lm/models/base.py:65-98), and KVCache.to_quantized
// (cache.py:251-263) converts a whole dense cache at once.
//
// Arithmetic (oracle/MLX_SPEC.md R6, fp32 throughout, every operation separately rounded -- no FMA can form because
// no multiply feeds an add):
//     w_max, w_min over the group;  mask = |w_min| > |w_max|
//     scale = max((w_max - w_min) / (2^bits - 1), 1e-7);  scale = mask ? scale : -scale;  edge = mask ? w_min : w_max
//     q0 = rint(edge / scale);  if (q0 != 0) { scale = edge / q0;  bias = edge; } else bias = 0
//     code = clamp(rint((w - bias) / scale), 0, 2^bits - 1)         <- from the UNROUNDED fp32 scale / bias
//     scales[g] = T(scale), biases[g] = T(bias)
// Codes, scales and biases are bit-exact against the restatement (tests/test_gpu_quantize.py).
//
// HBM-bound byte work: reads sizeof(T) bytes and writes bits/8 bytes per element (+ 2 * sizeof(T) per group).  The
// problem is flat: groups never cross rows (K % group_size == 0) and rows of codes are whole words (K % 32 == 0), so
// a warp takes 256 consecutive elements (8 per lane, one 16-byte load for the 16-bit types), reduces min / max inside
// the gs/8 lanes of a group with xor shuffles, stages its 32 * bits bytes of codes in shared memory and writes them
// back as coalesced words.  Destination rows can be remapped (gbxq_quantize_rows): the cache update writes the rows
// of every (batch, head) at the current offset of a [.., capacity, ..] buffer without a second copy.
#include "gbxq_common.cuh"

namespace gbxq {
namespace {

// destination row of source row r: (r / inner) * outer_stride + offset + r % inner;  inner <= 0: identity
struct RowMap {
    int64_t inner, outer_stride, offset;
};

__device__ __forceinline__ int64_t map_row(const RowMap& m, int64_t r) {
    const int64_t o = r / m.inner;
    return o * m.outer_stride + m.offset + (r - o * m.inner);
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        v[2 * i] = __low2float(h);
        v[2 * i + 1] = __high2float(h);
    }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

constexpr int kWarps = 8;

template <int BITS, typename T>
__global__ void __launch_bounds__(kWarps * 32) quantize_kernel(const T* __restrict__ w, uint32_t* __restrict__ q,
                                                              T* __restrict__ scales, T* __restrict__ biases,
                                                              int64_t total, int64_t K, int gs, RowMap map) {
    __shared__ __align__(16) uint8_t stage_all[kWarps][32 * BITS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all[warp];
    const int64_t nchunks = (total + 255) >> 8;
    const int64_t wpr = K * BITS / 32;  // words per row of codes
    const int64_t gpr = K / gs;         // groups per row
    const int64_t total_words = total / 32 * BITS;
    const int seg = gs >> 3;            // lanes per group: 4, 8, 16
    const float nb = (float)((1 << BITS) - 1);
    const bool dense = map.inner <= 0;

    for (int64_t c = (int64_t)blockIdx.x * kWarps + warp; c < nchunks; c += (int64_t)gridDim.x * kWarps) {
        const int64_t e0 = (c << 8) + (lane << 3);
        const bool live = e0 < total;  // total % 8 == 0: a lane's 8 elements are all inside or all outside
        float v[8];
        if (live) {
            load8(w + e0, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.f;
        }
        float mx = v[0], mn = v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) {
            mx = fmaxf(mx, v[i]);
            mn = fminf(mn, v[i]);
        }
        // groups are aligned runs of `seg` lanes; dead lanes only ever share a group with dead lanes (total % gs == 0)
        for (int o = 1; o < seg; o <<= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        const bool mask = fabsf(mn) > fabsf(mx);
        float scale = fmaxf(__fdiv_rn(__fsub_rn(mx, mn), nb), 1e-7f);
        scale = mask ? scale : -scale;
        const float edge = mask ? mn : mx;
        const float q0 = rintf(__fdiv_rn(edge, scale));
        float bias = 0.f;
        if (q0 != 0.f) {
            scale = __fdiv_rn(edge, q0);
            bias = edge;
        }
        uint64_t pack = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float t = rintf(__fdiv_rn(__fsub_rn(v[i], bias), scale));
            t = fminf(fmaxf(t, 0.f), nb);
            pack |= (uint64_t)(uint32_t)t << (i * BITS);
        }
        // the lane's 8 codes are bytes [lane * BITS, (lane + 1) * BITS) of the chunk's LSB-first stream
        if constexpr (BITS == 8) {
            *reinterpret_cast<uint64_t*>(stage + lane * 8) = pack;
        } else if constexpr (BITS == 4) {
            *reinterpret_cast<uint32_t*>(stage + lane * 4) = (uint32_t)pack;
        } else if constexpr (BITS == 2) {
            *reinterpret_cast<uint16_t*>(stage + lane * 2) = (uint16_t)pack;
        } else {
#pragma unroll
            for (int b = 0; b < BITS; b++) stage[lane * BITS + b] = (uint8_t)(pack >> (8 * b));
        }
        if (live && (lane & (seg - 1)) == 0) {
            int64_t g = e0 / gs;
            if (!dense) {
                const int64_t r = g / gpr;
                g = map_row(map, r) * gpr + (g - r * gpr);
            }
            scales[g] = from_f32<T>(scale);
            biases[g] = from_f32<T>(bias);
        }
        __syncwarp();
        const int64_t w0 = c * (8 * BITS);
        for (int j = lane; j < 8 * BITS; j += 32) {
            int64_t wi = w0 + j;
            if (wi < total_words) {
                const uint32_t word = reinterpret_cast<const uint32_t*>(stage)[j];
                if (!dense) {
                    const int64_t r = wi / wpr;
                    wi = map_row(map, r) * wpr + (wi - r * wpr);
                }
                q[wi] = word;
            }
        }
        __syncwarp();
    }
}

template <int BITS, typename T>
int launch_t(const void* w, uint32_t* q, void* s, void* b, int64_t total, int64_t K, int gs, const RowMap& map,
             cudaStream_t st) {
    const int64_t nchunks = (total + 255) >> 8;
    int64_t blocks = (nchunks + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)device_sm_count() * 8;  // 8 CTAs of 256 threads per SM, grid-stride beyond
    if (blocks > cap) blocks = cap;
    quantize_kernel<BITS, T><<<(unsigned)blocks, kWarps * 32, 0, st>>>((const T*)w, q, (T*)s, (T*)b, total, K, gs, map);
    count_launch();
    return check_cuda(cudaGetLastError());
}

template <int BITS>
int launch_b(const void* w, uint32_t* q, void* s, void* b, int64_t total, int64_t K, int gs, int dtype,
             const RowMap& map, cudaStream_t st) {
    switch (dtype) {
        case GBXQ_BF16: return launch_t<BITS, __nv_bfloat16>(w, q, s, b, total, K, gs, map, st);
        case GBXQ_F16: return launch_t<BITS, __half>(w, q, s, b, total, K, gs, map, st);
        default: return launch_t<BITS, float>(w, q, s, b, total, K, gs, map, st);
    }
}

int quantize_checked(const void* w, uint32_t* q, void* s, void* b, int64_t rows, int64_t K, int bits, int gs,
                     int dtype, const RowMap& map, cudaStream_t st) {
    if (bits != 2 && bits != 3 && bits != 4 && bits != 6 && bits != 8) return GBXQ_EINVAL_BITS;
    if (gs != 32 && gs != 64 && gs != 128) return GBXQ_EINVAL_GROUP;
    if (dtype != GBXQ_BF16 && dtype != GBXQ_F16 && dtype != GBXQ_F32) return GBXQ_EDTYPE;
    if (rows < 0 || K <= 0 || K % gs != 0) return GBXQ_ESHAPE;
    if (map.inner > 0 && (rows % map.inner != 0 || map.offset < 0 || map.outer_stride < map.offset + map.inner))
        return GBXQ_ESHAPE;
    if (rows == 0) return GBXQ_OK;
    if (!w || !q || !s || !b) return GBXQ_ENULL;
    const uintptr_t elem = dtype == GBXQ_F32 ? 4 : 2;
    if (((uintptr_t)w & 15) || ((uintptr_t)q & 3) || ((uintptr_t)s & (elem - 1)) || ((uintptr_t)b & (elem - 1)))
        return GBXQ_EALIGN;
    const int64_t total = rows * K;
    switch (bits) {
        case 2: return launch_b<2>(w, q, s, b, total, K, gs, dtype, map, st);
        case 3: return launch_b<3>(w, q, s, b, total, K, gs, dtype, map, st);
        case 4: return launch_b<4>(w, q, s, b, total, K, gs, dtype, map, st);
        case 6: return launch_b<6>(w, q, s, b, total, K, gs, dtype, map, st);
        default: return launch_b<8>(w, q, s, b, total, K, gs, dtype, map, st);
    }
}

}  // namespace
}  // namespace gbxq

extern "C" {

int gbxq_quantize(const void* w, uint32_t* qweight, void* scales, void* biases, int64_t rows, int64_t K, int bits,
                  int group_size, int dtype, void* stream) {
    return gbxq::quantize_checked(w, qweight, scales, biases, rows, K, bits, group_size, dtype, gbxq::RowMap{0, 0, 0},
                                  (cudaStream_t)stream);
}

int gbxq_quantize_rows(const void* w, uint32_t* qweight, void* scales, void* biases, int64_t rows, int64_t K, int bits,
                       int group_size, int dtype, int64_t inner_rows, int64_t out_outer_stride_rows,
                       int64_t out_row_offset, void* stream) {
    if (inner_rows <= 0) return GBXQ_ESHAPE;
    return gbxq::quantize_checked(w, qweight, scales, biases, rows, K, bits, group_size, dtype,
                                  gbxq::RowMap{inner_rows, out_outer_stride_rows, out_row_offset}, (cudaStream_t)stream);
}

}  // extern "C"
