// Micro-benchmark (B200): issue/pipe rates of FADD vs FADD2 (packed f32x2), alone and mixed with
// ALU / SHFL / LDS work.  Used to decide whether packing pays for the streamed stencil kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_issue fp_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N_ITER 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, float seed, int iters) {
    float a[CHAINS * 2];
    u64 p[CHAINS];
    int ia[CHAINS];
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 512] = seed;
    __syncthreads();
    for (int c = 0; c < CHAINS * 2; ++c) a[c] = seed + c + threadIdx.x;
    for (int c = 0; c < CHAINS; ++c) { p[c] = ((u64)__float_as_uint(a[2 * c]) << 32) | __float_as_uint(a[2 * c + 1]); ia[c] = c + threadIdx.x; }
    u64 inc = ((u64)__float_as_uint(seed) << 32) | __float_as_uint(seed);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {            // 16 scalar FADD
#pragma unroll
                for (int c = 0; c < CHAINS * 2; ++c) a[c] = a[c] + seed;
            } else if (MODE == 1) {     // 8 FADD2 (same flops as mode 0)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(inc));
            } else if (MODE == 2) {     // 8 FADD2 + 8 integer ALU ops
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(inc));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(ia[c]) : "r"(it));
                }
            } else if (MODE == 3) {     // 16 FADD + 8 integer ALU ops
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    a[2 * c] += seed; a[2 * c + 1] += seed;
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(ia[c]) : "r"(it));
                }
            } else if (MODE == 4) {     // 8 FADD2 + 2 SHFL + 1 LDS.128-ish
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(inc));
                a[0] = __shfl_up_sync(0xffffffffu, a[0], 1);
                a[1] = __shfl_down_sync(0xffffffffu, a[1], 1);
                a[2] += sm[(threadIdx.x + it) & 1023];
            } else if (MODE == 5) {     // 16 FADD + 2 SHFL + 1 LDS
#pragma unroll
                for (int c = 2; c < CHAINS * 2; ++c) a[c] = a[c] + seed;
                a[0] = __shfl_up_sync(0xffffffffu, a[0], 1) + seed;
                a[1] = __shfl_down_sync(0xffffffffu, a[1], 1) + seed;
                a[2] += sm[(threadIdx.x + it) & 1023];
            } else if (MODE == 6) {     // 8 FFMA2
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[c]) : "l"(inc));
            } else if (MODE == 7) {     // 8 DADD
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    double d = __longlong_as_double(p[c]);
                    d = d + (double)seed;
                    p[c] = __double_as_longlong(d);
                }
            } else if (MODE == 8) {     // 8 FADD2 + 8 MOV-like (IMAD.MOV / prmt)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(inc));
                    asm volatile("add.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(seed));
                }
            }
        }
    }
    float r = 0;
    for (int c = 0; c < CHAINS * 2; ++c) r += a[c];
    for (int c = 0; c < CHAINS; ++c) r += __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32)) + ia[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* what, double lane_ops_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 512 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512>>>(out, 1.0f, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(out, 1.0f, N_ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_groups = 148.0 * 16 * N_ITER * 8;      // executions of the inner group per warp, all SMs
    double cyc = ms * 1e-3 * 1.9e9;                     // rough, at ~1.9 GHz
    printf("%-44s %8.3f ms  %.2f cycles per group per SMSP-warp-slot (4 warps/SMSP => x4 = per SMSP)  err=%s\n",
           what, ms, cyc / (N_ITER * 8) / 4.0, cudaGetErrorString(cudaGetLastError()));
    (void)warp_groups; (void)lane_ops_per_iter;
    cudaFree(out);
}

int main() {
    run<0>("16 FADD", 16);
    run<1>("8 FADD2", 16);
    run<6>("8 FFMA2", 16);
    run<2>("8 FADD2 + 8 LOP3", 16);
    run<3>("16 FADD + 8 LOP3", 16);
    run<8>("8 FADD2 + 8 FADD", 16);
    run<4>("8 FADD2 + 2 SHFL + 1 LDS", 16);
    run<5>("16 FADD + 2 SHFL + 1 LDS", 16);
    run<7>("8 DADD", 8);
    return 0;
}
