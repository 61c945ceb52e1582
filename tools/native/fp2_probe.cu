// fp2_probe.cu -- does the packed FADD2 / FMUL2 path of sm_100a save issue slots? (tools, not product)
// Each thread runs CH independent dependent-chains of adds; scalar version issues 2x the instructions.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long ull;
__device__ __forceinline__ ull add2(ull a, ull b) { ull c; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ float add1(float a, float b) { float c; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(c) : "f"(a), "f"(b)); return c; }
template <int MODE> __global__ void k(float* out, int iters, float inc) {
    float a[8]; ull p[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = ((ull)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const ull inc2 = ((ull)__float_as_uint(inc) << 32) | __float_as_uint(inc);
    unsigned x = threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = add1(a[i], inc);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = add2(p[i], inc2);
        }
        if (MODE == 2 || MODE == 3) {   // mix in integer work: issue-bound case
#pragma unroll
            for (int i = 0; i < 8; i++) x = x * 3u + (unsigned)it;
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = add2(p[i], inc2);   // 3: scalar-equivalent of 16 adds as 8 FADD2
        }
    }
    float s = 0;
    if (MODE == 0) { for (int i = 0; i < 8; i++) s += a[i]; }
    else { for (int i = 0; i < 4; i++) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
}
template <int MODE> __global__ void k4(float* out, int iters, float inc) {   // scalar adds + integer mix
    float a[8]; unsigned x = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = add1(a[i], inc);
#pragma unroll
        for (int i = 0; i < 8; i++) x = x * 3u + (unsigned)it;
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    auto run = [&](const char* name, void (*f)(float*, int, float), double lane_adds_per_iter) {
        f<<<148 * 8, 256>>>(out, 100, 1.0f);
        cudaDeviceSynchronize();
        cudaEventRecord(e0); f<<<148 * 8, 256>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %8.3f ms  %.1f G lane-adds/s\n", name, ms, lane_adds_per_iter * iters * 148.0 * 8 * 256 / ms / 1e6);
    };
    run("scalar FADD x8", k<0>, 8);
    run("FADD2 x4 (=8 adds)", k<1>, 8);
    run("scalar FADD x8 + 8 IMAD", k4<0>, 8);
    run("FADD2 x4 + 8 IMAD", k<2>, 8);
    run("FADD2 x8 + 8 IMAD (16 adds)", k<3>, 16);
    return 0;
}
