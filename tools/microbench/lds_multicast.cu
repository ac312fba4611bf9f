// Microbenchmark: cost of a 16-byte-per-lane load when lanes share addresses (the access pattern of the neighbourhood walk:
// ~8 groups of ~4 lanes, each group reading one 16-byte record) -- LDS.128 from shared memory against LDG.128 from L1.
// Prints cycles per warp-wide load instruction for one resident warp per scheduler and for a full SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_multicast lds_multicast.cu && ./lds_multicast
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_lds(const int *offs, long long *out, int pattern) {
    extern __shared__ uint4 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int off = offs[pattern * 32 + lane] + warp * 97;
    unsigned int acc = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        const uint4 v = sm[(off + i) & 2047];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[1] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

__global__ void k_ldg(const uint4 *__restrict__ g, const int *offs, long long *out, int pattern) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int off = offs[pattern * 32 + lane] + warp * 97 + blockIdx.x * 4096;
    unsigned int acc = 0;
    for (int i = 0; i < 2048; ++i) acc += __ldg(g + ((off + i) & 2047) + blockIdx.x * 4096).x;   // warm L1
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        const uint4 v = __ldg(g + ((off + i) & 2047) + blockIdx.x * 4096);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[1] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    // patterns: record index per lane
    int h[6 * 32];
    const char *names[6] = {"32 distinct consecutive records", "one record for all lanes", "8 groups of 4 lanes, groups 4 records apart (cells of 4)",
                            "8 groups of 4 lanes, groups 5 records apart", "4 groups of 8 lanes, groups 3 records apart", "8 groups, irregular spacing 0,3,8,12,15,21,24,30"};
    const int irr[8] = {0, 3, 8, 12, 15, 21, 24, 30};
    for (int l = 0; l < 32; ++l) {
        h[0 * 32 + l] = l;
        h[1 * 32 + l] = 0;
        h[2 * 32 + l] = (l / 4) * 4;
        h[3 * 32 + l] = (l / 4) * 5;
        h[4 * 32 + l] = (l / 8) * 3;
        h[5 * 32 + l] = irr[l / 4];
    }
    int *d_offs;
    long long *d_out;
    uint4 *g;
    cudaMalloc(&d_offs, sizeof(h));
    cudaMemcpy(d_offs, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, 16);
    cudaMalloc(&g, sizeof(uint4) * 4096 * 256);
    cudaMemset(g, 1, sizeof(uint4) * 4096 * 256);
    for (int threads : {32, 128, 1024}) {
        for (int p = 0; p < 6; ++p) {
            long long c1 = 0, c2 = 0;
            k_lds<<<148, threads, 2048 * 16>>>(d_offs, d_out, p);
            cudaMemcpy(&c1, d_out, 8, cudaMemcpyDeviceToHost);
            k_ldg<<<148, threads>>>(g, d_offs, d_out, p);
            cudaMemcpy(&c2, d_out, 8, cudaMemcpyDeviceToHost);
            const double warps = threads / 32.0;
            printf("threads/block %4d  %-58s  LDS.128 %6.2f  LDG.128 %6.2f  cycles per warp-load per SM (elapsed/ITERS/warps)\n", threads, names[p],
                   (double)c1 / ITERS / warps, (double)c2 / ITERS / warps);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
