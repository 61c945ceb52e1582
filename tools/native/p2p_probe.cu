// p2p_probe.cu -- how fast can a kernel on GPU 0 stage 4608-byte blocks (128 triangle records) from GPU 1's HBM?
// Methods: LDG.128 (ld.global.nc), cp.async 16 B double-buffered, cp.async.bulk (TMA engine) double-buffered;
// plus plain remote stores. Same grid shape as the remote-staging voxelizer (148 x 4 blocks of 128 threads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/native/p2p_probe tools/native/p2p_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
constexpr int BLK = 128, CHUNK_F = 128 * 9;     // floats per staging block

__global__ void __launch_bounds__(BLK, 4) k_ldg(const float* src, size_t nchunks, float* out, int spin) {
    __shared__ float4 s[CHUNK_F / 4];
    float acc = 0.f;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const float4* p = reinterpret_cast<const float4*>(src + c * CHUNK_F);
        for (int i = threadIdx.x; i < CHUNK_F / 4; i += BLK) s[i] = __ldg(p + i);
        __syncthreads();
        float v = reinterpret_cast<float*>(s)[threadIdx.x * 9];
        for (int k = 0; k < spin; k++) v = v * 1.0001f + 0.5f;
        acc += v;
        __syncthreads();
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(BLK, 4) k_cpasync(const float* src, size_t nchunks, float* out, int spin) {
    __shared__ float4 s[2][CHUNK_F / 4];
    float acc = 0.f;
    auto issue = [&](size_t c, int b) {
        const float* p = src + c * CHUNK_F;
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&s[b][0]);
        for (int i = threadIdx.x; i < CHUNK_F / 4; i += BLK)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sb + i * 16), "l"(p + i * 4) : "memory");
    };
    size_t c = blockIdx.x;
    if (c < nchunks) issue(c, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int it = 0; c < nchunks; c += gridDim.x, it++) {
        if (c + gridDim.x < nchunks) issue(c + gridDim.x, (it + 1) & 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        float v = reinterpret_cast<float*>(&s[it & 1][0])[threadIdx.x * 9];
        __syncthreads();
        for (int k = 0; k < spin; k++) v = v * 1.0001f + 0.5f;
        acc += v;
    }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" :: "r"(bar), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(BLK, 4) k_bulk(const float* src, size_t nchunks, float* out, int spin) {
    __shared__ __align__(128) float4 s[2][CHUNK_F / 4];
    __shared__ __align__(8) unsigned long long bar[2];
    const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    auto issue = [&](size_t c, int b) {
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&s[b][0]);
        const uint32_t bb = b0 + 8 * b;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bb), "r"(CHUNK_F * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sb), "l"(src + c * CHUNK_F), "r"(CHUNK_F * 4), "r"(bb) : "memory");
    };
    size_t c = blockIdx.x;
    if (c < nchunks && threadIdx.x == 0) issue(c, 0);
    for (int it = 0; c < nchunks; c += gridDim.x, it++) {
        if (c + gridDim.x < nchunks && threadIdx.x == 0) issue(c + gridDim.x, (it + 1) & 1);
        mbar_wait(b0 + 8 * (it & 1), (it >> 1) & 1);
        float v = reinterpret_cast<float*>(&s[it & 1][0])[threadIdx.x * 9];
        __syncthreads();                           // everyone has read buffer it&1 before it is refilled two iterations later
        for (int k = 0; k < spin; k++) v = v * 1.0001f + 0.5f;
        acc += v;
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(BLK, 4) k_store(float* dst, size_t nchunks) {
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        float4* p = reinterpret_cast<float4*>(dst + c * CHUNK_F);
        for (int i = threadIdx.x; i < CHUNK_F / 4; i += BLK) p[i] = make_float4(1.f, 2.f, 3.f, (float)c);
    }
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t nchunks = 15625 * 4, bytes = nchunks * CHUNK_F * 4;      // 288 MB (larger than L2)
    float *loc, *rem, *out;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, bytes)); CK(cudaMemset(rem, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&loc, bytes)); CK(cudaMemset(loc, 0, bytes)); CK(cudaMalloc(&out, 64));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grids[] = { 148 * 4, 148 * 8, 148 * 16 };
    const int spins[] = { 0, 2000 };
    for (int spin : spins) for (int g : grids) {
        for (int which = 0; which < 2; which++) {
            const float* src = which ? rem : loc;
            float ms[4] = { 0, 0, 0, 0 };
            for (int m = 0; m < 3; m++) {
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    if (m == 0) k_ldg<<<g, BLK>>>(src, nchunks, out, spin);
                    if (m == 1) k_cpasync<<<g, BLK>>>(src, nchunks, out, spin);
                    if (m == 2) k_bulk<<<g, BLK>>>(src, nchunks, out, spin);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    CK(cudaEventElapsedTime(&ms[m], e0, e1));
                }
            }
            if (spin == 0) {
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    k_store<<<g, BLK>>>(which ? rem : loc, nchunks);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    CK(cudaEventElapsedTime(&ms[3], e0, e1));
                }
            }
            printf("spin %4d grid %4d %s: ldg128 %.3f ms (%.0f GB/s)  cp.async %.3f ms (%.0f GB/s)  bulk %.3f ms (%.0f GB/s)  store %.3f ms (%.0f GB/s)\n",
                   spin, g, which ? "REMOTE" : "local ", ms[0], bytes / ms[0] / 1e6, ms[1], bytes / ms[1] / 1e6, ms[2], bytes / ms[2] / 1e6,
                   ms[3], ms[3] > 0 ? bytes / ms[3] / 1e6 : 0.0);
        }
    }
    // both directions at once (what the sharded voxelizers do): GPU 0 reads GPU 1's buffer while GPU 1 reads GPU 0's
    {
        float* out1; cudaEvent_t f0, f1;
        CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&out1, 64)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
        for (int spin : spins) for (int rep = 0; rep < 3; rep++) {
            float a = 0, b = 0;
            CK(cudaSetDevice(0)); CK(cudaEventRecord(e0)); k_cpasync<<<592, BLK>>>(rem, nchunks, out, spin); CK(cudaEventRecord(e1));
            CK(cudaSetDevice(1)); CK(cudaEventRecord(f0)); k_cpasync<<<592, BLK>>>(loc, nchunks, out1, spin); CK(cudaEventRecord(f1));
            CK(cudaEventSynchronize(f1)); CK(cudaEventElapsedTime(&b, f0, f1));
            CK(cudaSetDevice(0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&a, e0, e1));
            if (rep == 2) printf("spin %4d grid  592 BIDIR : cp.async GPU0 %.3f ms (%.0f GB/s)  GPU1 %.3f ms (%.0f GB/s)\n", spin, a, bytes / a / 1e6, b, bytes / b / 1e6);
        }
    }
    return 0;
}
