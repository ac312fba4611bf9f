// cuda_on_host.h -- run CUDA C++ sources on the CPU, one host thread, for checking results.
//
// TEST INFRASTRUCTURE (oracle/): never linked into the product.  The reference's generated sources
// (FLAMEGPU/templates/*.xslt expanded by oracle/xslt_subset.py) target CUDA 8 texture references and cannot be compiled
// by CUDA 12.9; and this container has no GPU anyway.  This header gives those sources, unmodified
// apart from the <<<...>>> launch syntax (rewritten by oracle/build_ref.py into cuda_on_host::Launch), a host meaning:
//   * __global__/__device__/__constant__ vanish; device memory is host memory; cudaMemcpy* are memcpy;
//   * a launch runs every thread of every block in turn (thread (0,0,0) first, x fastest) with threadIdx/blockIdx set;
//     kernels that reach __syncthreads() run each block's threads as ucontext fibres, a barrier yielding to the next;
//   * texture<T> is a pointer, tex1Dfetch an index; atomics are plain read-modify-write (there is one host thread);
//   * cub::DeviceScan::ExclusiveSum / cub::DeviceRadixSort::SortPairs / thrust::sort_by_key, reduce, count, min/max
//     element have their specified results (exclusive prefix sum; stable sort on the selected key bits).
// Written from the CUDA runtime API's documented behaviour; it contains nothing of the reference.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>
#include <ucontext.h>
#include <sys/mman.h>

#define __host__
#define __device__
#define __global__
#define __constant__
#define __shared__ thread_local
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define CUDA_ON_HOST 1

// ---- math overloads ------------------------------------------------------------------------------------------------
// CUDA puts the float overloads of the C math functions in the global namespace (cos(float) is cosf); <cmath> alone
// leaves only the double versions there, which would silently compute in double and round.
using std::abs; using std::fabs; using std::sqrt; using std::cbrt; using std::pow; using std::exp; using std::exp2;
using std::log; using std::log2; using std::log10; using std::sin; using std::cos; using std::tan; using std::asin;
using std::acos; using std::atan; using std::atan2; using std::sinh; using std::cosh; using std::tanh; using std::floor;
using std::ceil; using std::round; using std::trunc; using std::fmod; using std::fmin; using std::fmax; using std::hypot;
using std::copysign; using std::isnan; using std::isinf; using std::min; using std::max;

// ---- vector and launch types ---------------------------------------------------------------------------------------
struct uint3 { unsigned int x, y, z; };
struct int2 { int x, y; };
struct uint2 { unsigned int x, y; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

inline thread_local uint3 threadIdx{0, 0, 0};
inline thread_local uint3 blockIdx{0, 0, 0};
inline thread_local dim3 blockDim;
inline thread_local dim3 gridDim;
static const int warpSize = 32;

// dynamic shared memory: sources declare `extern __shared__ int sm_data[];` inside device functions
#ifndef CUDA_ON_HOST_SHARED_BYTES
#define CUDA_ON_HOST_SHARED_BYTES (1 << 20)
#endif
// exactly one translation unit of a program states CUDA_ON_HOST_SHARED_MEMORY_DEFINITION at namespace scope
#define CUDA_ON_HOST_SHARED_MEMORY_DEFINITION thread_local int sm_data[CUDA_ON_HOST_SHARED_BYTES / 4] __attribute__((aligned(16)));

// ---- errors, devices, streams, events ------------------------------------------------------------------------------
enum cudaError { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
typedef cudaError cudaError_t;
static inline const char* cudaGetErrorString(cudaError_t e) {
    return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "cuda_on_host error";
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaThreadSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceReset() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount, maxThreadsPerBlock, warpSize, pciBusID, pciDeviceID, tccDriver, computeMode;
    size_t totalGlobalMem, sharedMemPerBlock;
    int maxGridSize[3], maxThreadsDim[3];
};
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "cuda_on_host (CPU)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 1; p->maxThreadsPerBlock = 1024; p->warpSize = 32;
    p->totalGlobalMem = size_t(1) << 40; p->sharedMemPerBlock = CUDA_ON_HOST_SHARED_BYTES;
    p->maxGridSize[0] = 2147483647; p->maxGridSize[1] = p->maxGridSize[2] = 65535;
    p->maxThreadsDim[0] = p->maxThreadsDim[1] = 1024; p->maxThreadsDim[2] = 64;
    return cudaSuccess;
}
typedef struct cuda_on_host_stream* cudaStream_t;
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
struct cuda_on_host_event { std::chrono::steady_clock::time_point t; };
typedef cuda_on_host_event* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cuda_on_host_event(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}

// ---- memory --------------------------------------------------------------------------------------------------------
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    *p = static_cast<T*>(std::calloc(bytes ? bytes : 1, 1));
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { std::memmove(dst, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(dst, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { std::memset(p, v, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* src, size_t n, size_t offset = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
    std::memcpy(reinterpret_cast<char*>(&symbol) + offset, src, n);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(T& symbol, const void* src, size_t n, size_t offset = 0, cudaMemcpyKind = cudaMemcpyHostToDevice, cudaStream_t = nullptr) {
    return cudaMemcpyToSymbol(symbol, src, n, offset);
}
template <class T> static inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& symbol, size_t n, size_t offset = 0, cudaMemcpyKind = cudaMemcpyDeviceToHost) {
    std::memcpy(dst, reinterpret_cast<const char*>(&symbol) + offset, n);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaGetSymbolAddress(void** p, T& symbol) { *p = &symbol; return cudaSuccess; }

// ---- textures (CUDA 8 texture references) --------------------------------------------------------------------------
enum cudaTextureReadMode { cudaReadModeElementType, cudaReadModeNormalizedFloat };
template <class T, int Dim = 1, cudaTextureReadMode Mode = cudaReadModeElementType> struct texture { const T* ptr = nullptr; };
template <class T, int D, cudaTextureReadMode M>
static inline cudaError_t cudaBindTexture(size_t* offset, texture<T, D, M>& tex, const void* p, size_t = ~size_t(0)) {
    tex.ptr = static_cast<const T*>(p);
    if (offset) *offset = 0;
    return cudaSuccess;
}
template <class T, int D, cudaTextureReadMode M> static inline cudaError_t cudaUnbindTexture(texture<T, D, M>& tex) { tex.ptr = nullptr; return cudaSuccess; }
template <class T, int D, cudaTextureReadMode M> static inline T tex1Dfetch(const texture<T, D, M>& tex, int i) { return tex.ptr[i]; }

// ---- device intrinsics ---------------------------------------------------------------------------------------------
static inline unsigned int __umul24(unsigned int a, unsigned int b) { return (a & 0xffffffu) * (b & 0xffffffu); }
static inline int __mul24(int a, int b) { return ((a << 8) >> 8) * ((b << 8) >> 8); }
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    unsigned long long u = ((unsigned long long)(unsigned int)hi << 32) | (unsigned int)lo;
    double d; std::memcpy(&d, &u, 8); return d;
}
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline unsigned int atomicInc(unsigned int* p, unsigned int limit) { unsigned int o = *p; *p = (o >= limit) ? 0u : o + 1u; return o; }
static inline unsigned int atomicDec(unsigned int* p, unsigned int limit) { unsigned int o = *p; *p = (o == 0u || o > limit) ? limit : o - 1u; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---- launches ------------------------------------------------------------------------------------------------------
namespace cuda_on_host {

struct BlockFibres {
    static constexpr size_t kStack = 256 * 1024;
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    ucontext_t scheduler;
    const std::function<void()>* body = nullptr;
    int current = -1;
    bool active = false;
    static BlockFibres& get() { static thread_local BlockFibres f; return f; }
    void reserve(size_t n) {
        while (stacks.size() < n) {
            void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) { std::fprintf(stderr, "cuda_on_host: cannot map a fibre stack\n"); std::abort(); }
            stacks.push_back(static_cast<char*>(p));
        }
        ctx.resize(std::max(ctx.size(), n));
        done.assign(n, 0);
    }
    static void trampoline() {
        BlockFibres& f = get();
        (*f.body)();
        f.done[f.current] = 1;
        swapcontext(&f.ctx[f.current], &f.scheduler);
    }
    void barrier() {
        if (!active) {
            if (blockDim.x * blockDim.y * blockDim.z == 1) return;
            std::fprintf(stderr, "cuda_on_host: __syncthreads() in a kernel launched without barrier support\n");
            std::abort();
        }
        swapcontext(&ctx[current], &scheduler);
    }
};

inline uint3 thread_coords(unsigned int t, const dim3& b) { return uint3{t % b.x, (t / b.x) % b.y, t / (b.x * b.y)}; }

struct Launch {
    dim3 grid, block;
    size_t shared;
    Launch(dim3 g, dim3 b, size_t sm = 0, cudaStream_t = nullptr) : grid(g), block(b), shared(sm) {
        if (sm > CUDA_ON_HOST_SHARED_BYTES) { std::fprintf(stderr, "cuda_on_host: %zu bytes of dynamic shared memory requested\n", sm); std::abort(); }
    }
    template <class F> void run(bool barriers, F&& f) const {
        const unsigned int threads = block.x * block.y * block.z;
        gridDim = grid;
        blockDim = block;
        for (unsigned int bz = 0; bz < grid.z; ++bz)
            for (unsigned int by = 0; by < grid.y; ++by)
                for (unsigned int bx = 0; bx < grid.x; ++bx) {
                    blockIdx = uint3{bx, by, bz};
                    if (!barriers) {
                        for (unsigned int t = 0; t < threads; ++t) { threadIdx = thread_coords(t, block); f(); }
                        continue;
                    }
                    BlockFibres& fib = BlockFibres::get();
                    fib.reserve(threads);
                    const std::function<void()> body = [&f]() { f(); };
                    fib.body = &body;
                    fib.active = true;
                    for (unsigned int t = 0; t < threads; ++t) {
                        getcontext(&fib.ctx[t]);
                        fib.ctx[t].uc_stack.ss_sp = fib.stacks[t];
                        fib.ctx[t].uc_stack.ss_size = BlockFibres::kStack;
                        fib.ctx[t].uc_link = nullptr;
                        makecontext(&fib.ctx[t], &BlockFibres::trampoline, 0);
                    }
                    unsigned int alive = threads;
                    while (alive) {
                        for (unsigned int t = 0; t < threads; ++t) {
                            if (fib.done[t]) continue;
                            threadIdx = thread_coords(t, block);
                            fib.current = int(t);
                            swapcontext(&fib.scheduler, &fib.ctx[t]);
                            if (fib.done[t]) --alive;
                        }
                    }
                    fib.active = false;
                }
        threadIdx = uint3{0, 0, 0};
        blockIdx = uint3{0, 0, 0};
    }
};

}  // namespace cuda_on_host

static inline void __syncthreads() { cuda_on_host::BlockFibres::get().barrier(); }

// occupancy queries: the block size a real device would choose is hardware-dependent.  Results do not depend on it,
// with one exception the reference has by construction: its non-partitioned message walk starts at the block's own
// tile (messages b*B .. N-1, then 0 .. b*B-1 for block b of size B), so float sums over such messages are ordered by
// block size.  CUDA_ON_HOST_BLOCK (environment, default 128, at most 1024) selects B.
static inline int cuda_on_host_block_size(int limit) {
    static const int chosen = [] {
        const char* e = std::getenv("CUDA_ON_HOST_BLOCK");
        int b = e ? std::atoi(e) : 128;
        return b < 1 ? 1 : b > 1024 ? 1024 : b;
    }();
    return (limit > 0 && limit < chosen) ? limit : chosen;
}
template <class K, class S>
static inline cudaError_t cudaOccupancyMaxPotentialBlockSizeVariableSMem(int* min_grid, int* block, K, S, int limit = 0) {
    *min_grid = 1;
    *block = cuda_on_host_block_size(limit);
    return cudaSuccess;
}
template <class K>
static inline cudaError_t cudaOccupancyMaxPotentialBlockSize(int* min_grid, int* block, K, size_t = 0, int limit = 0) {
    *min_grid = 1;
    *block = cuda_on_host_block_size(limit);
    return cudaSuccess;
}

// ---- cub -----------------------------------------------------------------------------------------------------------
namespace cub {
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void* temp, size_t& temp_bytes, In in, Out out, int n, cudaStream_t = nullptr, bool = false) {
        if (!temp) { temp_bytes = 16; return cudaSuccess; }
        auto acc = decltype(in[0] + in[0])(0);
        for (int i = 0; i < n; ++i) { auto v = in[i]; out[i] = acc; acc = acc + v; }
        return cudaSuccess;
    }
};
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, int n,
                                 int begin_bit = 0, int end_bit = int(sizeof(K) * 8), cudaStream_t = nullptr, bool = false) {
        if (!temp) { temp_bytes = 16; return cudaSuccess; }
        const int bits = end_bit - begin_bit;
        const unsigned long long mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return (((unsigned long long)keys_in[a] >> begin_bit) & mask) < (((unsigned long long)keys_in[b] >> begin_bit) & mask);
        });
        for (int i = 0; i < n; ++i) { keys_out[i] = keys_in[order[i]]; vals_out[i] = vals_in[order[i]]; }
        return cudaSuccess;
    }
};
}  // namespace cub

// ---- thrust --------------------------------------------------------------------------------------------------------
namespace thrust {
template <class T> struct device_ptr {
    T* p;
    device_ptr(T* q = nullptr) : p(q) {}
    T* get() const { return p; }
    T& operator*() const { return *p; }
    T& operator[](ptrdiff_t i) const { return p[i]; }
    device_ptr operator+(ptrdiff_t n) const { return device_ptr(p + n); }
    device_ptr operator-(ptrdiff_t n) const { return device_ptr(p - n); }
    ptrdiff_t operator-(const device_ptr& o) const { return p - o.p; }
    bool operator==(const device_ptr& o) const { return p == o.p; }
    bool operator!=(const device_ptr& o) const { return p != o.p; }
};
template <class T> static inline device_ptr<T> device_pointer_cast(T* p) { return device_ptr<T>(p); }
template <class T> static inline T* raw_pointer_cast(const device_ptr<T>& p) { return p.get(); }
// thrust::reduce sums in a tree whose shape is an implementation detail; a left-to-right sum has the same integer
// result and a floating-point result within rounding of it (callers that compare floats state a tolerance).
template <class T> static inline T reduce(device_ptr<T> a, device_ptr<T> b) { T s = T(0); for (T* p = a.get(); p != b.get(); ++p) s = s + *p; return s; }
template <class T, class V> static inline ptrdiff_t count(device_ptr<T> a, device_ptr<T> b, V v) { ptrdiff_t c = 0; for (T* p = a.get(); p != b.get(); ++p) c += (*p == v); return c; }
template <class T> static inline device_ptr<T> min_element(device_ptr<T> a, device_ptr<T> b) { return device_ptr<T>(std::min_element(a.get(), b.get())); }
template <class T> static inline device_ptr<T> max_element(device_ptr<T> a, device_ptr<T> b) { return device_ptr<T>(std::max_element(a.get(), b.get())); }
template <class K, class V> static inline void sort_by_key(device_ptr<K> a, device_ptr<K> b, device_ptr<V> v) {
    const ptrdiff_t n = b - a;
    std::vector<ptrdiff_t> order(n);
    std::iota(order.begin(), order.end(), ptrdiff_t(0));
    std::stable_sort(order.begin(), order.end(), [&](ptrdiff_t x, ptrdiff_t y) { return a[x] < a[y]; });
    std::vector<K> k(n); std::vector<V> w(n);
    for (ptrdiff_t i = 0; i < n; ++i) { k[i] = a[order[i]]; w[i] = v[order[i]]; }
    for (ptrdiff_t i = 0; i < n; ++i) { a[i] = k[i]; v[i] = w[i]; }
}
template <class T> static inline void exclusive_scan(device_ptr<T> a, device_ptr<T> b, device_ptr<T> out) {
    T acc = T(0);
    for (ptrdiff_t i = 0, n = b - a; i < n; ++i) { T v = a[i]; out[i] = acc; acc = acc + v; }
}
}  // namespace thrust

// ---- NVTX (only reached with -DPROFILE) ----------------------------------------------------------------------------
static inline int nvtxRangePushA(const char*) { return 0; }
static inline int nvtxRangePop() { return 0; }
