// Micro-benchmarks behind the decode launch strategy (run on B200 through gpurun):
//  A. floor of a chain of DEPENDENT launches replayed from a CUDA graph: plain stream order vs programmatic
//     dependent launch (trigger at entry, griddepcontrol.wait before the first dependent access), at the decode
//     kernels' footprint (288 threads, 96 KB dynamic shared memory, 148 / 296 CTAs).
//  B. L2 prefetch of the NEXT layer's weights: does `prefetch.global.L2` / `cp.async.bulk.prefetch.L2` outlive the
//     issuing kernel, how long does issuing 32 MB take, and at what rate does a 32 MB buffer stream out of L2
//     compared with HBM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ub_chain ub_chain.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(288) chain_kernel(float* buf, int pdl) {
    extern __shared__ float sm[];
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    // dependent access: read what the previous launch wrote, write for the next
    const float v = buf[(blockIdx.x * 288 + threadIdx.x) & 4095];
    sm[threadIdx.x] = v;
    __syncthreads();
    buf[(blockIdx.x * 288 + threadIdx.x) & 4095] = sm[(threadIdx.x + 1) % 288] + 1.0f;
}

static float time_graph(cudaGraphExec_t ge, cudaStream_t st, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; i++) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

static void chain_floor(cudaStream_t st, float* buf, int grid, int smem_kb, int pdl) {
    const int L = 256;
    CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < L; i++) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(288);
        cfg.dynamicSmemBytes = (size_t)smem_kb * 1024;
        cfg.stream = st;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        a[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = a;
        cfg.numAttrs = pdl ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, chain_kernel, buf, pdl));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    const float ms = time_graph(ge, st, 10);
    printf("chain floor: grid %3d smem %3d KB pdl %d : %.3f us / dependent launch\n", grid, smem_kb, pdl, ms * 1e3 / L);
    CK(cudaGraphExecDestroy(ge));
    CK(cudaGraphDestroy(g));
}

// ---- B. prefetch
__global__ void pf_lsu(const uint8_t* p, size_t bytes) {
    const size_t lines = bytes / 128;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
}
__global__ void pf_tma(const uint8_t* p, size_t bytes, unsigned chunk) {
    if (threadIdx.x == 0) {
        const size_t n = bytes / chunk;
        for (size_t i = blockIdx.x; i < n; i += gridDim.x)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i * chunk), "r"(chunk) : "memory");
    }
}
__global__ void spin_kernel(long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while ((long long)(t - t0) < ns);
}
__global__ void __launch_bounds__(256) rd_kernel(const uint4* p, size_t n16, unsigned* out) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345u) *out = acc;
}

static void prefetch_exp(cudaStream_t st) {
    const size_t MB = 1 << 20;
    const int NB = 10;
    unsigned* out;
    CK(cudaMalloc(&out, 4));
    for (size_t sz : {(size_t)8 * MB, (size_t)32 * MB, (size_t)64 * MB}) {
        std::vector<uint8_t*> b(NB);
        for (int i = 0; i < NB; i++) {
            CK(cudaMalloc(&b[i], sz));
            CK(cudaMemsetAsync(b[i], i + 1, sz, st));
        }
        uint8_t* flush;
        CK(cudaMalloc(&flush, 512 * MB));
        cudaEvent_t e0, e1, e2, e3;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
        for (int mode = 0; mode < 4; mode++) {
            // mode 0: cold read (HBM); 1: LSU prefetch + 20 us spin + read; 2: TMA prefetch (16 KB pieces) + spin + read;
            // 3: read twice back to back (second = L2-resident rate)
            float t_issue = 0, t_read = 0;
            const int reps = 6;
            for (int r = 0; r < reps; r++) {
                CK(cudaMemsetAsync(flush, r, 512 * MB, st));  // evict everything
                uint8_t* p = b[r % NB];
                CK(cudaEventRecord(e0, st));
                if (mode == 1) pf_lsu<<<296, 256, 0, st>>>(p, sz);
                if (mode == 2) pf_tma<<<296, 32, 0, st>>>(p, sz, 16384u);
                if (mode == 3) rd_kernel<<<148 * 8, 256, 0, st>>>((const uint4*)p, sz / 16, out);
                CK(cudaEventRecord(e1, st));
                if (mode == 1 || mode == 2) spin_kernel<<<1, 1, 0, st>>>(20000);
                CK(cudaEventRecord(e2, st));
                rd_kernel<<<148 * 8, 256, 0, st>>>((const uint4*)p, sz / 16, out);
                CK(cudaEventRecord(e3, st));
                CK(cudaStreamSynchronize(st));
                float a, c;
                CK(cudaEventElapsedTime(&a, e0, e1));
                CK(cudaEventElapsedTime(&c, e2, e3));
                if (r > 0) { t_issue += a; t_read += c; }
            }
            t_issue /= (reps - 1); t_read /= (reps - 1);
            const char* names[4] = {"cold read (HBM)", "LSU prefetch.L2, 20us, read", "TMA bulk prefetch.L2, 20us, read", "read after read (L2 resident)"};
            printf("prefetch: %3zu MB  %-34s issue/first %.2f us   read %.2f us = %.0f GB/s\n", sz / MB, names[mode], t_issue * 1e3,
                   t_read * 1e3, sz / (t_read * 1e-3) / 1e9);
        }
        for (int i = 0; i < NB; i++) CK(cudaFree(b[i]));
        CK(cudaFree(flush));
    }
}

int main() {
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float* buf;
    CK(cudaMalloc(&buf, 4096 * 4));
    CK(cudaMemset(buf, 0, 4096 * 4));
    for (int grid : {148, 296})
        for (int smem : {48, 96})
            for (int pdl : {0, 1}) chain_floor(st, buf, grid, smem, pdl);
    prefetch_exp(st);
    return 0;
}
