// micro-benchmark: per-lane contiguous runs of RUN bytes written with 16-byte stores (lane stride = RUN),
// versus warp-coalesced 16-byte stores.  Prints GB/s.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define RUN 144   // bytes per lane (multiple of 16)
__global__ void scattered(uint4* out, size_t nlanes) {
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= nlanes) return;
    uint4* p = out + lane * (RUN / 16);
    uint4 v = make_uint4(lane, 1, 2, 3);
#pragma unroll 1
    for (int i = 0; i < RUN / 16; ++i) { v.x += i; p[i] = v; }
}
__global__ void coalesced(uint4* out, size_t nlanes) {
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32, l = threadIdx.x & 31;
    if (warp * 32 >= nlanes) return;
    uint4* p = out + warp * 32 * (RUN / 16);
    uint4 v = make_uint4(warp, 1, 2, 3);
#pragma unroll 1
    for (int i = 0; i < RUN / 16; ++i) { v.x += i; p[i * 32 + l] = v; }
}
int main() {
    size_t bytes = 2ull << 30, nlanes = bytes / RUN;
    uint4* d; cudaMalloc(&d, bytes + 4096);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int k = 0; k < 2; ++k) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            if (k == 0) scattered<<<(unsigned)((nlanes + 255) / 256), 256>>>(d, nlanes);
            else coalesced<<<(unsigned)((nlanes + 255) / 256), 256>>>(d, nlanes);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep == 2) printf("%s: %.3f ms, %.1f GB/s\n", k == 0 ? "scattered 16B stores (lane stride 144B)" : "coalesced 16B stores", ms, bytes / ms / 1e6);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
