// micro-benchmark: random gather of 32-byte records (the access pattern of the window update's fold)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k_gather(const uint4 *rec, const uint32_t *idx, uint32_t n, uint32_t per_thread, uint4 *out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (uint32_t b = 0; b < per_thread; b += 8) {
        uint4 v[16];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t i = (b + q) * (n / per_thread) + t; // thread-strided index list
            const uint32_t r = i < n ? idx[i] : 0;
            v[2 * q] = rec[2 * (size_t)r]; v[2 * q + 1] = rec[2 * (size_t)r + 1];
        }
#pragma unroll
        for (int q = 0; q < 16; q++) { acc.x += v[q].x; acc.y ^= v[q].y; acc.z += v[q].z; acc.w ^= v[q].w; }
    }
    if (acc.x == 0x12345678u) out[t] = acc;
}
__global__ void k_gather_seq(const uint4 *rec, uint32_t n, uint4 *out) // same bytes, coalesced
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (uint32_t i = t; i < 2 * n; i += gridDim.x * blockDim.x) { const uint4 v = rec[i]; acc.x += v.x; acc.y ^= v.y; }
    if (acc.x == 0x12345678u) out[t] = acc;
}
int main()
{
    const uint32_t n = 2u << 20; // records
    std::vector<uint32_t> h(n);
    for (uint32_t i = 0; i < n; i++) h[i] = i;
    std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
    uint4 *rec, *out, *flush; uint32_t *idx;
    CK(cudaMalloc(&rec, (size_t)n * 32)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&out, (size_t)n * 16)); CK(cudaMalloc(&flush, 256u << 20));
    CK(cudaMemset(rec, 1, (size_t)n * 32)); CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t per_thread : {8u, 16u, 32u}) for (int cold = 1; cold >= 0; cold--) {
        const uint32_t threads = n / per_thread;
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            if (cold) CK(cudaMemset(flush, rep, 256u << 20));
            cudaEventRecord(e0);
            k_gather<<<threads / 128, 128>>>(rec, idx, n, per_thread, out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("random gather: %u threads x %u records, %s L2: %.1f us  (%.2f G rec/s, %.0f GB/s useful)\n", threads, per_thread, cold ? "cold" : "warm", best * 1e3, n / best / 1e6, n * 32.0 / best / 1e6);
    }
    for (int cold = 1; cold >= 0; cold--) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            if (cold) CK(cudaMemset(flush, rep, 256u << 20));
            cudaEventRecord(e0); k_gather_seq<<<148 * 8, 256>>>(rec, n, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("sequential read of the same 64 MB, %s L2: %.1f us (%.0f GB/s)\n", cold ? "cold" : "warm", best * 1e3, n * 32.0 / best / 1e6);
    }
    return 0;
}
