/*
 * gbxq_oracle.c -- plain-C CPU restatement of the gbx-lm quantized-matmul path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (gbx_lm_b200 / libgbxq.so) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in the third-party `mlx`
 * wheel (requirement `mlx>=0.25.0`, /root/reference/requirements.txt:2), which is not
 * under /root/reference and not installable here.  This file restates MLX's affine
 * quantisation as documented/remembered:
 *   - packing + 3/6-bit byte rules: mlx/backend/cpu/quantized.cpp `extract_bits`
 *   - x86 scalar matmul loop:       mlx/backend/cpu/quantized.cpp `_qmm_t`
 *                                   (sum += x * (scale * T(q) + bias), every op in T)
 *   - SIMD matmul loop:             `_qmm_t_simd` (fp32 weights, fp32 accumulate)
 *   - dequantize:                   mlx/ops.cpp `dequantize` op graph (multiply, then add, in T)
 * as reached from /root/reference/gbx_lm/models/quantized_linear_gba.py:195-203
 * (quantized_matmul, transpose=True) and gbx_lm/tuner/lora.py:62-68 (dequantize).
 *
 * dtype codes: 0 = bf16, 1 = f16, 2 = f32 (activations, scales, biases, outputs share it).
 * accum codes: 0 = fp32 k-ordered, fp32 affine weights ("MLX SIMD path")
 *              1 = T-typed arithmetic throughout    ("MLX x86 scalar path")
 *              2 = fp64 truth (exact affine weights, fp64 accumulate, one final rounding)
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ scalar types */
static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint16_t f32_to_bf16(float f) { /* round-to-nearest-even */
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) {
    _Float16 v;
    memcpy(&v, &h, 2);
    return (float)v;
}
static inline uint16_t f32_to_f16(float f) {
    _Float16 v = (_Float16)f;
    uint16_t h;
    memcpy(&h, &v, 2);
    return h;
}
static inline float load_t(const void* p, size_t i, int dtype) {
    if (dtype == 0) return bf16_to_f32(((const uint16_t*)p)[i]);
    if (dtype == 1) return f16_to_f32(((const uint16_t*)p)[i]);
    return ((const float*)p)[i];
}
static inline void store_t(void* p, size_t i, float v, int dtype) {
    if (dtype == 0) ((uint16_t*)p)[i] = f32_to_bf16(v);
    else if (dtype == 1) ((uint16_t*)p)[i] = f32_to_f16(v);
    else ((float*)p)[i] = v;
}
/* round an fp32 value to T and widen back */
static inline float round_t(float v, int dtype) {
    if (dtype == 0) return bf16_to_f32(f32_to_bf16(v));
    if (dtype == 1) return f16_to_f32(f32_to_f16(v));
    return v;
}

/* ------------------------------------------------------------------ packing */
/* code k of a row occupies stream bits [k*bits, (k+1)*bits), LSB first, little-endian u32 words */
static inline uint32_t code_at(const uint32_t* row, int64_t k, int bits) {
    int64_t bit = k * bits;
    int64_t wi = bit >> 5;
    int off = (int)(bit & 31);
    uint64_t two = row[wi];
    if (off + bits > 32) two |= (uint64_t)row[wi + 1] << 32;
    return (uint32_t)((two >> off) & ((1u << bits) - 1u));
}

static int valid(int bits, int gs, int64_t K) {
    if (!(bits == 2 || bits == 3 || bits == 4 || bits == 6 || bits == 8)) return 0;
    if (!(gs == 32 || gs == 64 || gs == 128)) return 0;
    if (K <= 0 || K % gs || K % 32) return 0;
    return 1;
}

int gbxq_oracle_unpack(const uint32_t* w, uint8_t* codes, int64_t N, int64_t K, int bits) {
    if (!(bits == 2 || bits == 3 || bits == 4 || bits == 6 || bits == 8) || K % 32) return -1;
    int64_t wpr = K * bits / 32;
    for (int64_t n = 0; n < N; n++)
        for (int64_t k = 0; k < K; k++) codes[n * K + k] = (uint8_t)code_at(w + n * wpr, k, bits);
    return 0;
}

int gbxq_oracle_pack(const uint8_t* codes, uint32_t* w, int64_t N, int64_t K, int bits) {
    if (!(bits == 2 || bits == 3 || bits == 4 || bits == 6 || bits == 8) || K % 32) return -1;
    int64_t wpr = K * bits / 32;
    memset(w, 0, (size_t)(N * wpr) * 4);
    for (int64_t n = 0; n < N; n++)
        for (int64_t k = 0; k < K; k++) {
            uint64_t c = codes[n * K + k] & ((1u << bits) - 1u);
            int64_t bit = k * bits;
            int64_t wi = bit >> 5;
            int off = (int)(bit & 31);
            w[n * wpr + wi] |= (uint32_t)(c << off);
            if (off + bits > 32) w[n * wpr + wi + 1] |= (uint32_t)(c >> (32 - off));
        }
    return 0;
}

/* ------------------------------------------------------------------ dequantize */
/* W[n,k] = T( T(scale * T(q)) + bias )  -- two roundings, no FMA (mlx/ops.cpp dequantize) */
int gbxq_oracle_dequantize(const uint32_t* w, const void* scales, const void* biases, void* out,
                           int64_t N, int64_t K, int bits, int gs, int dtype) {
    if (!valid(bits, gs, K) || dtype < 0 || dtype > 2) return -1;
    int64_t wpr = K * bits / 32, G = K / gs;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++) {
        for (int64_t k = 0; k < K; k++) {
            float s = load_t(scales, n * G + k / gs, dtype);
            float b = load_t(biases, n * G + k / gs, dtype);
            float q = (float)code_at(w + n * wpr, k, bits);
            volatile float prod = s * q; /* volatile: forbid contraction into an FMA */
            float p = round_t(prod, dtype);
            volatile float sum = p + b;
            store_t(out, n * K + k, sum, dtype);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ quantized matmul */
int gbxq_oracle_qmm(const void* x, const uint32_t* w, const void* scales, const void* biases,
                    const void* bias /* nullable [N] */, void* y, int64_t M, int64_t N, int64_t K,
                    int bits, int gs, int dtype, int accum, int nthreads) {
    if (!valid(bits, gs, K) || dtype < 0 || dtype > 2 || accum < 0 || accum > 2) return -1;
    int64_t wpr = K * bits / 32, G = K / gs;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t m = 0; m < M; m++) {
        for (int64_t n = 0; n < N; n++) {
            const uint32_t* row = w + n * wpr;
            float out;
            if (accum == 0) { /* fp32 weights, fp32 accumulate, k order */
                float acc = 0.f;
                for (int64_t g = 0; g < G; g++) {
                    float s = load_t(scales, n * G + g, dtype), b = load_t(biases, n * G + g, dtype);
                    for (int64_t k = g * gs; k < (g + 1) * gs; k++) {
                        volatile float wv = s * (float)code_at(row, k, bits);
                        wv = wv + b;
                        volatile float pr = load_t(x, m * K + k, dtype) * wv;
                        acc += pr;
                    }
                }
                out = acc;
            } else if (accum == 1) { /* every op rounded to T: sum += x * (scale * T(q) + bias) */
                float acc = 0.f;
                for (int64_t g = 0; g < G; g++) {
                    float s = load_t(scales, n * G + g, dtype), b = load_t(biases, n * G + g, dtype);
                    for (int64_t k = g * gs; k < (g + 1) * gs; k++) {
                        volatile float t = s * (float)code_at(row, k, bits);
                        float t1 = round_t(t, dtype);
                        t = t1 + b;
                        float t2 = round_t(t, dtype);
                        t = load_t(x, m * K + k, dtype) * t2;
                        float t3 = round_t(t, dtype);
                        t = acc + t3;
                        acc = round_t(t, dtype);
                    }
                }
                out = acc;
            } else { /* fp64 truth */
                double acc = 0.0;
                for (int64_t g = 0; g < G; g++) {
                    double s = load_t(scales, n * G + g, dtype), b = load_t(biases, n * G + g, dtype);
                    double qs = 0.0, xs = 0.0;
                    for (int64_t k = g * gs; k < (g + 1) * gs; k++) {
                        double xv = load_t(x, m * K + k, dtype);
                        qs += xv * (double)code_at(row, k, bits);
                        xs += xv;
                    }
                    acc += s * qs + b * xs;
                }
                out = (float)acc;
            }
            float r = round_t(out, dtype);
            if (bias) { /* QuantizedLinear.__call__: x = x + bias, a second rounded add */
                volatile float t = r + load_t(bias, n, dtype);
                r = round_t(t, dtype);
            }
            store_t(y, m * N + n, r, dtype);
        }
    }
    return 0;
}

/* Timing-oriented variant for bench.py's CPU arm: the same fp32 k-ordered arithmetic as accum=0
 * restructured per 32-code block so that the compiler can keep codes in registers; threads over n.
 * Results equal accum=0 up to fp32 re-association inside a group (sum x*q and sum x are kept
 * separately per group).  Only used as a *reported baseline*. */
int gbxq_oracle_qmm_fast(const void* x, const uint32_t* w, const void* scales, const void* biases,
                         void* y, int64_t M, int64_t N, int64_t K, int bits, int gs, int dtype,
                         int nthreads) {
    if (!valid(bits, gs, K) || dtype != 0) return -1;
    int64_t wpr = K * bits / 32, G = K / gs;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const uint16_t* xb = (const uint16_t*)x;
#pragma omp parallel
    {
        float xbuf[128];
#pragma omp for schedule(static)
        for (int64_t n = 0; n < N; n++) {
            const uint32_t* row = w + n * wpr;
            for (int64_t m = 0; m < M; m++) {
                float acc = 0.f;
                for (int64_t g = 0; g < G; g++) {
                    float s = bf16_to_f32(((const uint16_t*)scales)[n * G + g]);
                    float b = bf16_to_f32(((const uint16_t*)biases)[n * G + g]);
                    float qs = 0.f, xs = 0.f;
                    for (int i = 0; i < gs; i++) xbuf[i] = bf16_to_f32(xb[m * K + g * gs + i]);
                    for (int i = 0; i < gs; i++) {
                        qs += xbuf[i] * (float)code_at(row, g * gs + i, bits);
                        xs += xbuf[i];
                    }
                    acc += s * qs + b * xs;
                }
                ((uint16_t*)y)[m * N + n] = f32_to_bf16(acc);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ quantize */
/* mx.quantize (MLX_SPEC R6), written as MLX's CPU quantiser walks a matrix (mlx/backend/cpu/quantized.cpp `quantize`,
 * as recalled): one group at a time, fp32 statistics, the edge of larger magnitude anchored exactly, the codes from the
 * UNROUNDED scale / bias, the statistics cast to T last.  Reached from the quantized KV cache,
 * /root/reference/gbx_lm/models/cache.py:176-177,251-263.  rows x K values of T in, codes [rows, K*bits/32] out. */
int gbxq_oracle_quantize(const void* w, uint32_t* q, void* scales, void* biases, int64_t rows, int64_t K, int bits,
                         int gs, int dtype) {
    if (!valid(bits, gs, K) || dtype < 0 || dtype > 2 || rows < 0) return -1;
    const int64_t wpr = K * bits / 32, G = K / gs;
    const float n_bins = (float)((1 << bits) - 1);
    const float eps = 1e-7f;
    memset(q, 0, (size_t)(rows * wpr) * 4);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t g = 0; g < G; g++) {
            const int64_t base = r * K + g * gs;
            float w_min = INFINITY, w_max = -INFINITY;
            for (int j = 0; j < gs; j++) {
                const float v = load_t(w, (size_t)(base + j), dtype);
                w_max = v > w_max ? v : w_max;
                w_min = v < w_min ? v : w_min;
            }
            const int mask = fabsf(w_min) > fabsf(w_max);
            volatile float range = w_max - w_min;
            float scale = range / n_bins;
            scale = scale > eps ? scale : eps;
            scale = mask ? scale : -scale;
            const float edge = mask ? w_min : w_max;
            const float q0 = rintf(edge / scale);
            float bias = 0.0f;
            if (q0 != 0.0f) {
                scale = edge / q0;
                bias = edge;
            }
            for (int j = 0; j < gs; j++) {
                volatile float d = load_t(w, (size_t)(base + j), dtype) - bias;
                float c = rintf(d / scale);
                c = c < 0.0f ? 0.0f : (c > n_bins ? n_bins : c);
                const uint64_t code = (uint64_t)(uint32_t)c;
                const int64_t bit = (g * gs + j) * bits;
                const int64_t wi = bit >> 5;
                const int off = (int)(bit & 31);
                q[r * wpr + wi] |= (uint32_t)(code << off);
                if (off + bits > 32) q[r * wpr + wi + 1] |= (uint32_t)(code >> (32 - off));
            }
            store_t(scales, (size_t)(r * G + g), scale, dtype);
            store_t(biases, (size_t)(r * G + g), bias, dtype);
        }
    }
    return 0;
}

int gbxq_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
