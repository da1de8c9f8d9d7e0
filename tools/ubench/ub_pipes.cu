// Micro-benchmarks that decide the decode kernel's unpack strategy (run on B200 through gpurun):
//  1. exactness of mma.sync m16n8k32 e4m3 x e5m2 / e4m3 x e4m3 with raw 4-bit codes as e4m3 bytes
//  2. issue rates per SM sub-partition: HMMA bf16 m16n8k16, QMMA fp8 m16n8k32, SHF, LOP3, IMAD, IMAD.HI, IMAD.WIDE, PRMT
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma_fp8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e5m2.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_fp8_44(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_i8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- 1. exactness: A[16][32] = codes 0..15 as raw bytes, B[32][8] = e5m2 bytes; D vs host reference
__global__ void exact_kernel(const uint8_t* A, const uint8_t* B, float* D, int mode) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    // m16n8k32 8-bit fragments: a0: row g, k = 4t..4t+3; a1: row g+8, same k; a2: row g, k = 16+4t..; a3: row g+8, k = 16+4t..
    // b0: k = 4t..4t+3, col g; b1: k = 16+4t.., col g
    auto ld4 = [](const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    uint32_t a0 = ld4(A + g * 32 + 4 * t), a1 = ld4(A + (g + 8) * 32 + 4 * t), a2 = ld4(A + g * 32 + 16 + 4 * t), a3 = ld4(A + (g + 8) * 32 + 16 + 4 * t);
    uint8_t bb[8];
    for (int i = 0; i < 4; i++) { bb[i] = B[(4 * t + i) * 8 + g]; bb[4 + i] = B[(16 + 4 * t + i) * 8 + g]; }
    uint32_t b0 = ld4(bb), b1 = ld4(bb + 4);
    float d[4] = {0, 0, 0, 0};
    if (mode == 0) mma_fp8(d, a0, a1, a2, a3, b0, b1); else mma_fp8_44(d, a0, a1, a2, a3, b0, b1);
    D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}
static float e5m2_to_f(uint8_t v) {
    int s = v >> 7, e = (v >> 2) & 31, m = v & 3;
    float f = e == 0 ? ldexpf((float)m, -16) : ldexpf(1.f + m / 4.f, e - 15);
    return s ? -f : f;
}
static float e4m3_to_f(uint8_t v) {
    int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
    float f = e == 0 ? ldexpf((float)m, -9) : ldexpf(1.f + m / 8.f, e - 7);
    return s ? -f : f;
}

// ---- 2. rates
template <int OP>
__global__ void rate_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = seed * (threadIdx.x + 1) + i;
    float d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0}, d3[4] = {0, 0, 0, 0};
    int i0[4] = {0, 0, 0, 0}, i1[4] = {0, 0, 0, 0}, i2[4] = {0, 0, 0, 0}, i3[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) { mma_bf16(d0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_bf16(d1, r[0], r[1], r[2], r[3], r[4], r[5]); mma_bf16(d2, r[0], r[1], r[2], r[3], r[4], r[5]); mma_bf16(d3, r[0], r[1], r[2], r[3], r[4], r[5]); }
            if (OP == 1) { mma_fp8(d0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_fp8(d1, r[0], r[1], r[2], r[3], r[4], r[5]); mma_fp8(d2, r[0], r[1], r[2], r[3], r[4], r[5]); mma_fp8(d3, r[0], r[1], r[2], r[3], r[4], r[5]); }
            if (OP == 10) { mma_i8(i0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_i8(i1, r[0], r[1], r[2], r[3], r[4], r[5]); mma_i8(i2, r[0], r[1], r[2], r[3], r[4], r[5]); mma_i8(i3, r[0], r[1], r[2], r[3], r[4], r[5]); }
            if (OP == 11) { mma_i8(i0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_i8(i0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_i8(i0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_i8(i0, r[0], r[1], r[2], r[3], r[4], r[5]); }
            if (OP == 12) { mma_bf16(d0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_bf16(d0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_bf16(d0, r[0], r[1], r[2], r[3], r[4], r[5]); mma_bf16(d0, r[0], r[1], r[2], r[3], r[4], r[5]); }
            if (OP == 2) { for (int i = 0; i < 8; i++) asm volatile("shf.r.clamp.b32 %0, %0, 0, 4;" : "+r"(r[i])); }
            if (OP == 3) { for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, 0x000f000f, 0x43004300, 0xEA;" : "+r"(r[i])); }
            if (OP == 4) { for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, 16, 3;" : "+r"(r[i])); }
            if (OP == 5) { for (int i = 0; i < 8; i++) asm volatile("mul.hi.u32 %0, %0, 0x10000000;" : "+r"(r[i])); }
            if (OP == 6) { for (int i = 0; i < 8; i++) asm volatile("prmt.b32 %0, %0, %0, 0x3120;" : "+r"(r[i])); }
            if (OP == 7) { for (int i = 0; i < 8; i++) asm volatile("shr.u32 %0, %0, 4;" : "+r"(r[i])); }
            if (OP == 8) { for (int i = 0; i < 4; i++) { asm volatile("shr.u32 %0, %0, 4;" : "+r"(r[i])); asm volatile("mad.lo.u32 %0, %0, 16, 3;" : "+r"(r[4 + i])); } }
            if (OP == 9) { for (int i = 0; i < 4; i++) { asm volatile("lop3.b32 %0, %0, 0x000f000f, 0x43004300, 0xEA;" : "+r"(r[i])); asm volatile("mul.hi.u32 %0, %0, 0x10000000;" : "+r"(r[4 + i])); } }
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 8; i++) acc ^= r[i];
    acc ^= __float_as_uint(d0[0] + d1[1] + d2[2] + d3[3]) ^ (uint32_t)(i0[0] + i1[1] + i2[2] + i3[3]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP> void run_rate(const char* name, int per_iter, int warps) {
    uint32_t* out; CK(cudaMalloc(&out, 148 * 1024 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    rate_kernel<OP><<<148, warps * 32>>>(out, 10, 1);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    rate_kernel<OP><<<148, warps * 32>>>(out, iters, 1);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double inst_per_smsp = (double)iters * 16 * per_iter * warps / 4.0;
    printf("%-28s warps/SM %2d: %.3f warp-inst/clk/SMSP  (%.2f clk per inst)\n", name, warps, inst_per_smsp / cycles, cycles / inst_per_smsp);
    cudaFree(out);
}

int main() {
    // exactness
    uint8_t hA[16 * 32], hB[32 * 8]; float hD[128];
    srand(1);
    for (int mode = 0; mode < 2; mode++) {
        for (int i = 0; i < 512; i++) hA[i] = rand() & 15;
        for (int i = 0; i < 256; i++) {
            if (mode == 0) { int e = 10 + rand() % 12, m = rand() & 3, s = rand() & 1; hB[i] = (uint8_t)((s << 7) | (e << 2) | m); }
            else { int e = 1 + rand() % 13, m = rand() & 7, s = rand() & 1; hB[i] = (uint8_t)((s << 7) | (e << 3) | m); }
        }
        uint8_t *dA, *dB; float* dD;
        CK(cudaMalloc(&dA, 512)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dD, 512));
        CK(cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice));
        exact_kernel<<<1, 32>>>(dA, dB, dD, mode);
        CK(cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < 16; i++) for (int j = 0; j < 8; j++) {
            double ref = 0;
            for (int k = 0; k < 32; k++) ref += (double)(hA[i * 32 + k] * ldexp(1.0, -9)) * (mode == 0 ? e5m2_to_f(hB[k * 8 + j]) : e4m3_to_f(hB[k * 8 + j]));
            maxerr = fmax(maxerr, fabs(ref - hD[i * 8 + j])); maxref = fmax(maxref, fabs(ref));
        }
        printf("exactness mode %d (A = raw nibble as e4m3, B = %s): max |err| %.3e  max |ref| %.3e  rel %.3e\n", mode, mode == 0 ? "e5m2" : "e4m3", maxerr, maxref, maxerr / maxref);
    }
    for (int w : {4, 8, 16}) {
        run_rate<0>("HMMA bf16 m16n8k16", 4, w);
        run_rate<1>("fp8 mma.sync (emulated: F2FP+HMMA)", 4, w);
        run_rate<10>("IMMA u8xs8 m16n8k32", 4, w);
    }
    run_rate<11>("IMMA dependent chain", 4, 4);
    run_rate<12>("HMMA dependent chain", 4, 4);
    for (int w : {8, 16}) {
        run_rate<2>("SHF.R funnel", 8, w);
        run_rate<7>("SHR (shr.u32)", 8, w);
        run_rate<3>("LOP3", 8, w);
        run_rate<4>("IMAD (mad.lo)", 8, w);
        run_rate<5>("IMAD.HI (mul.hi)", 8, w);
        run_rate<6>("PRMT", 8, w);
        run_rate<8>("SHR + IMAD mix", 8, w);
        run_rate<9>("LOP3 + IMAD.HI mix", 8, w);
    }
    return 0;
}
