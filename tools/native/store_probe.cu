// store_probe.cu -- development microbenchmark (not part of the product): what write bandwidth do the store patterns
// of the node-record emitters reach on this GPU, and how does it move with the number of resident warps?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu && ./store_probe [GiB]
// Patterns (all write every byte of the buffer exactly once, 16-byte stores):
//   0  linear        thread i of the grid writes 16 B at 16 i, grid-stride (what a fill kernel does)
//   1  brick groups  a warp owns a 15360-byte batch (32 "bricks" of 480 B); per round 4 bricks, 8 lanes per brick, a lane
//                    writes 16 B at brick + 16 s + {0, 128, 256, 384}: the leaf emitter's pattern
//   2  warp linear   the same batches, but the 32 lanes write the round's 1920 bytes as one contiguous run
//   3  brick groups, batches handed out in order (ticket) instead of round-robin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st128(void* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b));
}
constexpr int BRICK = 480, BATCH = 32 * BRICK;

__global__ void __launch_bounds__(256) k_linear(char* buf, size_t bytes) {
    const size_t n = bytes / 16, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st128(buf + 16 * i, i, ~i);
}
template <int MODE>
__global__ void __launch_bounds__(256) k_batches(char* buf, size_t nbatch, unsigned long long* ticket) {
    const int lane = threadIdx.x & 31, g = lane >> 3, s = lane & 7;
    const size_t nwarps = (size_t)gridDim.x * 8;
    size_t b = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (MODE == 3) { unsigned long long t = 0; if (lane == 0) t = atomicAdd(ticket, 1ULL); b = __shfl_sync(0xffffffffu, t, 0); }
    while (b < nbatch) {
        char* base = buf + b * BATCH;
        for (int r = 0; r < 32; r += 4) {
            if (MODE == 2) {
                char* p = base + r * BRICK + 16 * lane;
#pragma unroll
                for (int i = 0; i < 4; i++) if (512 * i + 16 * lane < 4 * BRICK) st128(p + 512 * i, b, r);
            } else {
                char* p = base + (r + g) * BRICK + 16 * s;
                st128(p, b, r); st128(p + 128, b, r); st128(p + 256, b, r);
                if (s < 6) st128(p + 384, b, r);
            }
        }
        if (MODE == 3) { unsigned long long t = 0; if (lane == 0) t = atomicAdd(ticket, 1ULL); b = __shfl_sync(0xffffffffu, t, 0); }
        else b += nwarps;
    }
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 8.0;
    size_t nbatch = (size_t)(gib * (1ULL << 30)) / BATCH;
    const size_t bytes = nbatch * BATCH;
    char* buf; unsigned long long* ticket;
    if (cudaMalloc(&buf, bytes + 256) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&ticket, 8);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"bytes\": %zu, \"sms\": %d, \"results\": [\n", bytes, sms);
    bool first = true;
    for (int mode = 0; mode < 8; mode++) {
        // modes 4..7 = modes 0..3 with the whole buffer shifted by 16 bytes: the same stores, but every contiguous run of a
        // group now begins and ends in the middle of a 32-byte sector (what 24-byte records do to the emitters)
        char* const b0 = buf;
        char* buf = b0 + (mode >= 4 ? 16 : 0);
        const int mode_ = mode; 
        for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
            const int mode = mode_ & 3;
            float best = 1e30f;
            for (int it = 0; it < 4; it++) {
                cudaMemset(ticket, 0, 8);
                cudaEventRecord(e0);
                const unsigned grid = sms * per_sm;
                if (mode == 0) k_linear<<<grid, 256>>>(buf, bytes);
                else if (mode == 1) k_batches<1><<<grid, 256>>>(buf, nbatch, ticket);
                else if (mode == 2) k_batches<2><<<grid, 256>>>(buf, nbatch, ticket);
                else k_batches<3><<<grid, 256>>>(buf, nbatch, ticket);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            printf("%s {\"mode\": %d, \"shift16\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"GBps\": %.0f}", first ? "" : ",\n", mode, mode_ >= 4 ? 1 : 0, per_sm, best, bytes / best / 1e6);
            first = false;
        }
    }
    printf("\n]}\n");
    return cudaGetLastError() != cudaSuccess;
}
