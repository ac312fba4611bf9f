// Small CUDA program exercising what oracle/cuda_on_host/ gives a host meaning to; tests/test_cuda_on_host.py
// rewrites its launches with oracle/build_ref.rewrite_launches, compiles it with g++ and checks the printed results.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/device_ptr.h>
#include <thrust/sort.h>
#include <cstdio>

__constant__ int d_n;
__device__ int d_counter;
texture<float, 1, cudaReadModeElementType> tex_values;

// reverse each block through dynamic shared memory: needs a real barrier between the store and the load
__global__ void reverse_blocks(const int* in, int* out) {
    extern __shared__ int sm_data[];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    sm_data[threadIdx.x] = i < d_n ? in[i] : -1;
    __syncthreads();
    int j = blockIdx.x * blockDim.x + (blockDim.x - 1 - threadIdx.x);
    if (i < d_n && j < d_n) out[i] = sm_data[blockDim.x - 1 - threadIdx.x];
}

// two barriers in a loop, static shared memory
__global__ void block_sums(const int* in, int* sums) {
    __shared__ int acc[64];
    acc[threadIdx.x] = in[blockIdx.x * blockDim.x + threadIdx.x];
    for (int stride = blockDim.x / 2; stride > 0; stride /= 2) {
        __syncthreads();
        if (threadIdx.x < stride) acc[threadIdx.x] += acc[threadIdx.x + stride];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = acc[0];
}

__global__ void count_and_fetch(float* out, unsigned int* ring) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d_n) return;
    atomicAdd(&d_counter, 1);
    atomicInc(ring, 4u);
    out[i] = tex1Dfetch(tex_values, i) * 2.0f + cos(0.0f);
}

int main() {
    const int n = 200, block = 64, grid = (n + block - 1) / block;
    int h_in[256], h_out[256] = {0}, *d_in, *d_out, *d_sums, h_sums[4];
    for (int i = 0; i < 256; ++i) h_in[i] = i < n ? 3 * i + 1 : 0;
    cudaMalloc((void**)&d_in, sizeof(h_in)); cudaMalloc((void**)&d_out, sizeof(h_out)); cudaMalloc((void**)&d_sums, sizeof(h_sums));
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(d_n, &n, sizeof(int));
    reverse_blocks<<<grid, block, block * sizeof(int)>>>(d_in, d_out);
    cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
    printf("reverse %d %d %d %d\n", h_out[0], h_out[63], h_out[64], h_out[127]);
    block_sums<<<4, 64>>>(d_in, d_sums);
    cudaMemcpy(h_sums, d_sums, sizeof(h_sums), cudaMemcpyDeviceToHost);
    printf("sums %d %d %d %d\n", h_sums[0], h_sums[1], h_sums[2], h_sums[3]);

    float h_vals[256], *d_vals, *d_res, h_res[256];
    for (int i = 0; i < 256; ++i) h_vals[i] = 0.5f * i;
    cudaMalloc((void**)&d_vals, sizeof(h_vals)); cudaMalloc((void**)&d_res, sizeof(h_res));
    cudaMemcpy(d_vals, h_vals, sizeof(h_vals), cudaMemcpyHostToDevice);
    size_t offset = 99; cudaBindTexture(&offset, tex_values, d_vals, sizeof(h_vals));
    unsigned int* d_ring; cudaMalloc((void**)&d_ring, 4);
    int zero = 0; cudaMemcpyToSymbol(d_counter, &zero, sizeof(int));
    dim3 g(grid), b(block);
    count_and_fetch<<<g, b, 0, 0>>>(d_res, d_ring);
    int counter; unsigned int ring; cudaMemcpyFromSymbol(&counter, d_counter, sizeof(int)); cudaMemcpy(&ring, d_ring, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_res, d_res, sizeof(h_res), cudaMemcpyDeviceToHost);
    printf("fetch %d %u %zu %.1f %.1f\n", counter, ring, offset, h_res[1], h_res[199]);

    // exclusive sum and a stable radix sort on the low 3 bits only
    int scan[8], h_k[8] = {5, 1, 4, 1, 5, 9, 2, 6}, *d_k, *d_s; void* tmp = nullptr; size_t bytes = 0;
    cudaMalloc((void**)&d_k, sizeof(h_k)); cudaMalloc((void**)&d_s, sizeof(scan)); cudaMemcpy(d_k, h_k, sizeof(h_k), cudaMemcpyHostToDevice);
    cub::DeviceScan::ExclusiveSum(tmp, bytes, d_k, d_s, 8); cudaMalloc(&tmp, bytes); cub::DeviceScan::ExclusiveSum(tmp, bytes, d_k, d_s, 8);
    cudaMemcpy(scan, d_s, sizeof(scan), cudaMemcpyDeviceToHost);
    printf("scan %d %d %d %d\n", scan[0], scan[1], scan[4], scan[7]);
    unsigned int keys[6] = {9, 1, 17, 2, 10, 1}, vals[6] = {0, 1, 2, 3, 4, 5}, ko[6], vo[6];
    void* tmp2 = nullptr; size_t bytes2 = 0;
    cub::DeviceRadixSort::SortPairs(tmp2, bytes2, keys, ko, vals, vo, 6, 0, 3); tmp2 = &bytes2;
    cub::DeviceRadixSort::SortPairs(tmp2, bytes2, keys, ko, vals, vo, 6, 0, 3);
    printf("sort %u%u%u%u%u%u\n", vo[0], vo[1], vo[2], vo[3], vo[4], vo[5]);
    unsigned int k2[5] = {3, 1, 3, 0, 1}, v2[5] = {0, 1, 2, 3, 4};
    thrust::sort_by_key(thrust::device_pointer_cast(k2), thrust::device_pointer_cast(k2) + 5, thrust::device_pointer_cast(v2));
    printf("thrust %u%u%u%u%u\n", v2[0], v2[1], v2[2], v2[3], v2[4]);
    return 0;
}
