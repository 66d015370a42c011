// Issue-rate microbenchmark for the GEMM epilogue's arithmetic on sm_100a: scalar FFMA (register and immediate form), FMUL, packed
// FFMA2 / FMUL2 (fma.rn.f32x2, mul.rn.f32x2), MUFU.RCP, MUFU.EX2.  One CTA of NW warps per SM, each thread runs 8 independent
// chains of the instruction under test; prints warp-instructions per clock per SM sub-partition and results per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/fp32x2 tools/microbench/fp32x2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int MODE>
__global__ void k(float* out, long long* cycles, float seed) {
    float a[CH], b = seed, c = seed * 0.5f;
    uint64_t A[CH], B = pk(seed, seed + 1.f), C = pk(seed * 0.5f, seed);
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + i + threadIdx.x; A[i] = pk(a[i], a[i] + 1.f); }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(a[i]) : "f"(b));
            if (MODE == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B), "l"(C));
            if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B));
            if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 7) asm volatile("mul.rn.f32 %0, %0, %0;" : "+f"(a[i]));
            if (MODE == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 9) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float x, y; upk(A[i], x, y); s += a[i] + x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int results_per_instr, int nw) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    k<MODE><<<148, nw * 32>>>(out, cyc, 1.0001f);
    k<MODE><<<148, nw * 32>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double winstr = (double)nw * ITERS * CH;                 // warp-instructions per SM
    printf("%-28s warps/SM %2d: %.3f warp-instr/clk/SMSP, %.1f results/clk/SM\n", name, nw, winstr / c / 4.0, winstr * 32 * results_per_instr / c);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int nw : {8, 16}) {
        run<0>("FFMA reg,reg,reg", 1, nw);
        run<1>("FFMA reg,reg,imm", 1, nw);
        run<2>("FMUL reg,reg", 1, nw);
        run<7>("FMUL r,r (square)", 1, nw);
        run<3>("FFMA2 (f32x2)", 2, nw);
        run<4>("FMUL2 (f32x2)", 2, nw);
        run<9>("FADD2 (f32x2)", 2, nw);
        run<5>("MUFU.RCP", 1, nw);
        run<6>("MUFU.EX2", 1, nw);
        run<8>("FMNMX", 1, nw);
    }
    return 0;
}
