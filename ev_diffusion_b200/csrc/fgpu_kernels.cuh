/*
 * fgpu_kernels.cuh -- model-agnostic sm_100a kernels of the simulation step.
 *
 *   (1) spatial message build: LSD radix sort of (cell key, message index)
 *       (per pass: tile histograms, per-digit scan over tiles, stable scatter),
 *       then a fused gather of the packed message records + partition-boundary
 *       (CSR cell_start) construction.
 *       Replaces hist/reorder (krn:899-934) and hash + cub::DeviceRadixSort +
 *       reorder (krn:944-1021, sim:1860-1885); the order produced is the
 *       deterministic one: stable by cell, ties by message index.
 *   (4) stream compaction for death / birth / function conditions: warp ballot
 *       + block scan, stable.  Replaces reset_scan_input + cub ExclusiveSum +
 *       scatter_<A>_Agents (krn:222-257, sim:1763-1829).
 *   (5) append for non-swappable state moves (krn:265-280).
 *
 * No cub, no thrust.  Everything here is HBM/L2-bound integer work: kernels
 * are written for coalesced 4/16-byte accesses and sized in whole tiles.
 */
#ifndef FGPU_KERNELS_CUH
#define FGPU_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace fgpu {

/* Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
 * be scheduled while its predecessor in the stream is still draining; it must not touch anything the predecessor
 * wrote before this returns.  The iteration is a chain of 14-25 short dependent kernels, so the launch latency hidden
 * per link adds up (profiles/r02_notes.md).  A no-op for an ordinary launch. */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

/* ------------------------------------------------------------------------- */
/* radix sort                                                                 */
/* ------------------------------------------------------------------------- */
/* LSD radix sort of (cell key, message index).  One pass = three launches over
 * tiles of 4096 keys, with no inter-block waiting of any kind:
 *   k_tile_hist     digit counts of every tile            -> tile_hist[bin][tile]
 *   k_row_scan      per digit: exclusive scan over tiles  -> tile_hist (in place), totals[bin]
 *   k_radix_scatter re-reads the tile, ranks it in shared memory (one warp ballot per digit bit
 *                   + per-warp 16-bit counters, stable; match.any measured ~200 cycles on the ADU
 *                   pipe), stages it in digit order and writes contiguous runs per digit.
 * (Round 1 first used a single-sweep pass with decoupled look-back; on B200 a
 * 4096-key tile takes ~5 us while the look-back chain across the ~450 co-resident
 * tiles took ~20 us per wave -- profiles/r01_notes.md -- so the chain-free form is
 * both faster and free of spin-waits.)
 * Digits are up to 11 bits wide; the digit width is balanced over the passes. */
#ifndef FGPU_SCATTER_MINB_WIDE
#define FGPU_SCATTER_MINB_WIDE 3   /* resident blocks per SM asked of the scatter pass for 10- and 11-bit digits: 80 registers, 4 bytes spilled; 2.1M keys of 19 bits 89.6 -> 86.4 us, 4.2M of 20 bits 137.7 -> 127.7 (tools/sort_tune.py) */
#endif
constexpr int SORT_MAXBITS = 11;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; /* 4096 keys per tile */
constexpr int SORT_MAX_PASSES = 8;
constexpr int SORT_MAX_BINS = 1 << SORT_MAXBITS;

/* digit counts of one tile; tile_hist is bin-major so that k_row_scan reads rows contiguously */
__global__ void __launch_bounds__(SORT_THREADS)
k_tile_hist(const unsigned int *__restrict__ keys, int n_host, const int *__restrict__ d_n, int shift, int digit_bits,
            unsigned int *__restrict__ tile_hist, int tiles, unsigned int key_base, unsigned int key_span, int *range_error,
            unsigned int *zero2) {
    pdl_wait();
    extern __shared__ unsigned int sh[]; /* 1 << digit_bits */
    if (zero2 != nullptr && blockIdx.x == 0 && threadIdx.x == 0) zero2[0] = zero2[1] = 0u; /* gap counters of the build this sort starts */
    const int n = d_n ? *d_n : n_host; /* slab mode: the exact count lives on the device, n_host is an upper bound */
    const int nbins = 1 << digit_bits;
    for (int i = threadIdx.x; i < nbins; i += SORT_THREADS) sh[i] = 0;
    __syncthreads();
    const unsigned int mask = (unsigned int)nbins - 1u;
    const int base = blockIdx.x * SORT_TILE;
    unsigned int k[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) { /* all loads in flight before the first shared atomic */
        const int idx = base + r * SORT_THREADS + threadIdx.x;
        k[r] = (idx < n) ? keys[idx] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int idx = base + r * SORT_THREADS + threadIdx.x;
        if (idx < n) {
            /* slab mode sorts on the key relative to the first owned cell (fewer digits); a key outside
             * the owned cells would be mis-sorted silently, so the first pass reports it */
            const unsigned int rel = k[r] - key_base;
            if (range_error != nullptr && rel >= key_span) *range_error = 1;
            atomicAdd(&sh[(rel >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += SORT_THREADS) tile_hist[(size_t)i * tiles + blockIdx.x] = sh[i];
}

/* Block-wide exclusive scan of one value per thread (256 threads). */
__device__ __forceinline__ unsigned int block_exclusive_scan_256(unsigned int v, unsigned int *s_warp /*[8]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned int base = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
        if (w < warp) base += s_warp[w];
    __syncthreads();
    return base + inc - v;
}

/* one block per digit: exclusive scan of that digit's counts over the tiles, in place */
__global__ void __launch_bounds__(SORT_THREADS)
k_row_scan(unsigned int *__restrict__ tile_hist, int tiles, unsigned int *__restrict__ totals) {
    pdl_wait();
    __shared__ unsigned int s_scan[SORT_WARPS];
    unsigned int *row = tile_hist + (size_t)blockIdx.x * tiles;
    const int per = (tiles + SORT_THREADS - 1) / SORT_THREADS;
    const int lo = threadIdx.x * per, hi = min(lo + per, tiles);
    unsigned int sum = 0;
    for (int i = lo; i < hi; ++i) sum += row[i];
    unsigned int run = block_exclusive_scan_256(sum, s_scan);
    for (int i = lo; i < hi; ++i) {
        const unsigned int t = row[i];
        row[i] = run;
        run += t;
    }
    if (threadIdx.x == SORT_THREADS - 1) totals[blockIdx.x] = run;
}

/* shared memory of one scatter pass with digits of BITS bits */
template <int BITS>
struct SortSmem {
    static constexpr int BINS = 1 << BITS;
    unsigned short warp_hist[SORT_WARPS][BINS];
    unsigned int tile_excl[BINS];
    int goff[BINS];
    unsigned int keys[SORT_TILE];
    unsigned int vals[SORT_TILE];
    unsigned int scan[SORT_WARPS];
};

/*   vals_in == nullptr : values are the identity (first pass)
 * Stability: within a tile items are ranked in (warp, round, lane) order, which is
 * ascending input index; tiles are in input order.  BITS is the compile-time digit
 * capacity (bins per thread = 2^BITS / 256); digit_bits <= BITS is the width in use. */
template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS, (BITS <= 9 ? 3 : FGPU_SCATTER_MINB_WIDE))
k_radix_scatter(const unsigned int *__restrict__ keys_in, const unsigned int *__restrict__ vals_in,
                unsigned int *__restrict__ keys_out, unsigned int *__restrict__ vals_out, int n_host,
                const int *__restrict__ d_n, int shift, int digit_bits, const unsigned int *__restrict__ tile_hist,
                const unsigned int *__restrict__ totals, int tiles, unsigned int key_base,
                const uint4 *__restrict__ pay_in, uint4 *__restrict__ pay_out, int pay_quads) {
    pdl_wait();
    constexpr int BINS = 1 << BITS;
    const int n = d_n ? *d_n : n_host;
    constexpr int BPT = BINS / SORT_THREADS; /* bins per thread: 1, 2, 4 or 8 */
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem<BITS> &sm = *reinterpret_cast<SortSmem<BITS> *>(sort_smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        unsigned int *z = reinterpret_cast<unsigned int *>(&sm.warp_hist[0][0]);
        for (int i = tid; i < SORT_WARPS * BINS / 2; i += SORT_THREADS) z[i] = 0;
    }
    __syncthreads();
    const int tile = blockIdx.x;
    const int base = tile * SORT_TILE;
    const int nbins = 1 << digit_bits;
    const unsigned int mask = (unsigned int)nbins - 1u;
    const unsigned int lt = (1u << lane) - 1u;

    unsigned int key[SORT_ITEMS], val[SORT_ITEMS];
    unsigned short rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int idx = base + warp * (SORT_ITEMS * 32) + r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0u;
        val[r] = valid ? (vals_in ? vals_in[idx] : (unsigned int)idx) : 0xffffffffu;
    }
    /* Stable ranking inside the warp.  The lanes holding the same digit are found with one
     * ballot per digit bit (BITS cheap VOTEs) instead of match.any: ncu showed match.any on
     * ~30 distinct digits per warp costing ~200 cycles of the ADU pipe per call, 58 % pipe
     * utilisation and the top stall of the pass (profiles/r01_notes.md). */
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int idx = base + warp * (SORT_ITEMS * 32) + r * 32 + lane;
        const bool valid = idx < n;
        const unsigned int d = ((key[r] - key_base) >> shift) & mask;
        unsigned int peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < BITS; ++b) {
            /* m = all ones if bit b of the digit is set; lanes agreeing with me on this bit are ~(bal ^ m) */
            const int m = ((int)(d << (31 - b))) >> 31;
            const unsigned int bal = __ballot_sync(0xffffffffu, m != 0);
            peers &= ~(bal ^ (unsigned int)m);
        }
        /* the first lane of a group reads the group's counter and adds the group size; the others get the old
         * value from it by shuffle (one lane touches the counter per round: no race inside a round, and the
         * warp barrier below orders this round's update before the next round's read by another lane) */
        const unsigned int before = __popc(peers & lt);
        unsigned int prev = 0;
        if (valid && before == 0) {
            prev = sm.warp_hist[warp][d];
            sm.warp_hist[warp][d] = (unsigned short)(prev + __popc(peers));
        }
        prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
        rank[r] = (unsigned short)(prev + before);
        __syncwarp();
    }
    __syncthreads();

    /* per digit: exclusive offsets across warps and the tile count.  Thread t owns the BPT
     * consecutive bins [t*BPT, (t+1)*BPT), so one block scan of the per-thread sums yields the
     * exclusive prefix over all bins (inside the tile, and of the global digit totals). */
    unsigned int count[BPT], gcount[BPT];
    unsigned int csum = 0, gsum = 0;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        const int d = tid * BPT + k;
        unsigned int run = 0;
        if (d < nbins) {
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) {
                const unsigned int t = sm.warp_hist[w][d];
                sm.warp_hist[w][d] = (unsigned short)run;
                run += t;
            }
        }
        count[k] = run;
        gcount[k] = (d < nbins) ? totals[d] : 0u;
        csum += run;
        gsum += gcount[k];
    }
    unsigned int tile_excl = block_exclusive_scan_256(csum, sm.scan);
    unsigned int digit_base = block_exclusive_scan_256(gsum, sm.scan);
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        const int d = tid * BPT + k;
        if (d < nbins) {
            sm.tile_excl[d] = tile_excl;
            sm.goff[d] = (int)(digit_base + tile_hist[(size_t)d * tiles + tile]) - (int)tile_excl;
        }
        tile_excl += count[k];
        digit_base += gcount[k];
    }
    __syncthreads();

    /* rank -> position inside the tile, staged in shared memory so that the global writes
     * below are contiguous runs per digit */
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int idx = base + warp * (SORT_ITEMS * 32) + r * 32 + lane;
        if (idx < n) {
            const unsigned int d = ((key[r] - key_base) >> shift) & mask;
            const unsigned int lp = sm.tile_excl[d] + sm.warp_hist[warp][d] + rank[r];
            sm.keys[lp] = key[r];
            sm.vals[lp] = val[r];
        }
    }
    __syncthreads();
    const int valid_count = min(SORT_TILE, n - base);
    for (int i = tid; i < valid_count; i += SORT_THREADS) {
        const unsigned int k = sm.keys[i];
        const int dst = sm.goff[((k - key_base) >> shift) & mask] + i;
        keys_out[dst] = k;
        const unsigned int v = sm.vals[i];
        vals_out[dst] = v;
        /* MSD build: the 16-byte message records move with their key.  In a first pass v is the record's own
         * index, so these reads stay inside the tile's contiguous 64 KB of records (sequential DRAM traffic), and
         * records of one digit leave as a contiguous run. */
        for (int q = 0; q < pay_quads; ++q) pay_out[(size_t)dst * pay_quads + q] = __ldg(pay_in + (size_t)v * pay_quads + q);
    }
}

/* ------------------------------------------------------------------------- */
/* MSD build: bucket finish                                                   */
/* ------------------------------------------------------------------------- */
/* After ONE stable pass on the high digit (k_tile_hist + k_row_scan + k_radix_scatter with the records as
 * payload) the list is partitioned into buckets of 2^low_bits consecutive cells, each still in message-index
 * order.  One block per bucket finishes the job inside the bucket's own window of memory:
 *   - histogram of the low digit over the bucket, exclusive scan = the bucket's slice of the partition
 *     boundary matrix (cell_start), written for EVERY cell of the bucket, empty or not -- no boundary
 *     detection, no gap filling;
 *   - stable placement in chunks of 4096 messages (one ballot per digit bit inside a warp, per-warp counters,
 *     running per-digit offsets across chunks): message index -> order_out, record -> sorted_out.
 * All random accesses stay inside the bucket's window (a few tens of KB), so DRAM only sees sequential
 * traffic; the LSD build's final gather reads one 64-byte DRAM burst per 16-byte record instead
 * (profiles/r01_notes.md: 107 B per record at 16M messages).  Any bucket size works (more chunks). */
template <int BITS>
struct BucketSmem {
    static constexpr int BINS = 1 << BITS;
    unsigned short warp_hist[SORT_WARPS][BINS];
    unsigned int excl[BINS];
    unsigned int scan[SORT_WARPS];
    unsigned int base;
};

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
k_bucket_finish(const unsigned int *__restrict__ keys_a, const unsigned int *__restrict__ vals_a, const uint4 *__restrict__ pay_a,
                int pay_quads, unsigned int *__restrict__ order_out, uint4 *__restrict__ sorted_out,
                const unsigned int *__restrict__ totals, int nbuckets, int low_bits, unsigned int key_base,
                unsigned int *__restrict__ cell_start, unsigned int first_cell, unsigned int cells_end, unsigned int value_offset) {
    constexpr int BINS = 1 << BITS;
    constexpr int BPT = BINS / SORT_THREADS;
    extern __shared__ __align__(16) unsigned char bucket_smem_raw[];
    BucketSmem<BITS> &sm = *reinterpret_cast<BucketSmem<BITS> *>(bucket_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int nbins = 1 << low_bits;
    const unsigned int mask = (unsigned int)nbins - 1u;
    const unsigned int lt = (1u << lane) - 1u;

    /* position of the bucket in the list = messages in the buckets before it */
    {
        unsigned int part = 0;
        for (int j = tid; j < b; j += SORT_THREADS) part += totals[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) sm.scan[warp] = part;
        for (int i = tid; i < BINS; i += SORT_THREADS) sm.excl[i] = 0;
        __syncthreads();
        if (tid == 0) {
            unsigned int t = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) t += sm.scan[w];
            sm.base = t;
        }
        __syncthreads();
    }
    const unsigned int base = sm.base;
    const int cnt = (int)totals[b];
    const unsigned int *bk = keys_a + base;

    /* low-digit histogram of the bucket */
    for (int i = tid; i < cnt; i += SORT_THREADS) atomicAdd(&sm.excl[(bk[i] - key_base) & mask], 1u);
    __syncthreads();
    {
        unsigned int c[BPT], csum = 0;
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const int d = tid * BPT + k;
            c[k] = (d < nbins) ? sm.excl[d] : 0u;
            csum += c[k];
        }
        unsigned int run = block_exclusive_scan_256(csum, sm.scan);
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const int d = tid * BPT + k;
            if (d < nbins) {
                sm.excl[d] = run;
                const unsigned int cell = key_base + ((unsigned int)b << low_bits) + (unsigned int)d;
                if (cell >= first_cell && cell < cells_end) cell_start[cell] = value_offset + base + run;
            }
            run += c[k];
        }
        if (b == nbuckets - 1 && tid == 0) cell_start[cells_end] = value_offset + base + (unsigned int)cnt;
    }

    /* stable placement, a chunk of 4096 messages at a time */
    for (int c0 = 0; c0 < cnt; c0 += SORT_TILE) {
        {
            unsigned int *z = reinterpret_cast<unsigned int *>(&sm.warp_hist[0][0]);
            for (int i = tid; i < SORT_WARPS * BINS / 2; i += SORT_THREADS) z[i] = 0;
        }
        __syncthreads();
        unsigned int key[SORT_ITEMS];
        unsigned short rank[SORT_ITEMS];
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r) {
            const int idx = c0 + warp * (SORT_ITEMS * 32) + r * 32 + lane;
            key[r] = (idx < cnt) ? bk[idx] : 0u;
        }
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r) {
            const int idx = c0 + warp * (SORT_ITEMS * 32) + r * 32 + lane;
            const bool valid = idx < cnt;
            const unsigned int d = (key[r] - key_base) & mask;
            unsigned int peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int bit = 0; bit < BITS; ++bit) {
                const int m = ((int)(d << (31 - bit))) >> 31;
                const unsigned int bal = __ballot_sync(0xffffffffu, m != 0);
                peers &= ~(bal ^ (unsigned int)m);
            }
            const unsigned int before = __popc(peers & lt);
            unsigned int prev = 0;
            if (valid && before == 0) {
                prev = sm.warp_hist[warp][d];
                sm.warp_hist[warp][d] = (unsigned short)(prev + __popc(peers));
            }
            prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
            rank[r] = (unsigned short)(prev + before);
            __syncwarp();
        }
        __syncthreads();
        /* per digit: exclusive offsets across the warps of this chunk */
        unsigned int chunk_cnt[BPT];
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const int d = tid * BPT + k;
            unsigned int run = 0;
            if (d < nbins) {
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) {
                    const unsigned int t = sm.warp_hist[w][d];
                    sm.warp_hist[w][d] = (unsigned short)run;
                    run += t;
                }
            }
            chunk_cnt[k] = run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r) {
            const int idx = c0 + warp * (SORT_ITEMS * 32) + r * 32 + lane;
            if (idx < cnt) {
                const unsigned int d = (key[r] - key_base) & mask;
                const size_t dst = (size_t)base + sm.excl[d] + sm.warp_hist[warp][d] + rank[r];
                const size_t src = (size_t)base + idx;
                order_out[dst] = vals_a[src];
                for (int q = 0; q < pay_quads; ++q) sorted_out[dst * pay_quads + q] = pay_a[src * pay_quads + q];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const int d = tid * BPT + k;
            if (d < nbins) sm.excl[d] += chunk_cnt[k];
        }
        /* the barrier at the top of the next chunk orders these updates before their next use */
    }
}

/* ------------------------------------------------------------------------- */
/* gather of packed records + partition boundary matrix                       */
/* ------------------------------------------------------------------------- */
constexpr int GAP_INLINE = 32;
constexpr unsigned int GAP_LONG = 1u << 14; /* gaps longer than this are filled by the whole grid */
constexpr int GAP_LONG_MAX = 1024;           /* capacity of the long-gap queue (>= cells / GAP_LONG for 16M cells) */

__device__ __forceinline__ void fill_cells(unsigned int *cell_start, unsigned int lo, unsigned int hi /*inclusive*/,
                                           unsigned int value, unsigned int *gap_queue, unsigned int *gap_count) {
    if (hi < lo) return;
    if (hi - lo < (unsigned int)GAP_INLINE) {
        for (unsigned int c = lo; c <= hi; ++c) cell_start[c] = value;
    } else if (hi - lo < GAP_LONG) {
        const unsigned int e = atomicAdd(gap_count, 1u);
        gap_queue[3 * (GAP_LONG_MAX + e) + 0] = lo;
        gap_queue[3 * (GAP_LONG_MAX + e) + 1] = hi;
        gap_queue[3 * (GAP_LONG_MAX + e) + 2] = value;
    } else {
        /* long run of empty cells (e.g. the part of the grid another slab owns): its own queue, filled grid-wide;
         * a run that does not fit the queue is cut into GAP_LONG pieces on the ordinary queue */
        const unsigned int e = atomicAdd(gap_count + 1, 1u);
        if (e < (unsigned int)GAP_LONG_MAX) {
            gap_queue[3 * e + 0] = lo;
            gap_queue[3 * e + 1] = hi;
            gap_queue[3 * e + 2] = value;
        } else {
            for (unsigned int a = lo; a <= hi; a += GAP_LONG) {
                const unsigned int q = atomicAdd(gap_count, 1u);
                gap_queue[3 * (GAP_LONG_MAX + q) + 0] = a;
                gap_queue[3 * (GAP_LONG_MAX + q) + 1] = min(hi, a + GAP_LONG - 1u);
                gap_queue[3 * (GAP_LONG_MAX + q) + 2] = value;
            }
        }
    }
}

/* Thread i owns sorted position i: moves message order[i] there and, where the
 * cell key changes, writes cell_start for every cell in (prev_key, key].
 * cell_start[c] = first sorted index of cell c = number of messages in cells < c;
 * cell_start[cells] = n.  Empty cells get start == next start (count 0). */
template <int NQ>
__global__ void __launch_bounds__(256)
k_gather_pbm(const unsigned int *__restrict__ skeys, const unsigned int *__restrict__ svals,
             const uint4 *__restrict__ in_recs, uint4 *__restrict__ out_recs, int n_host, const int *__restrict__ d_n,
             unsigned int *__restrict__ cell_start, unsigned int first_cell, unsigned int cells, unsigned int value_offset,
             unsigned int *gap_queue, unsigned int *gap_count, const uint4 *__restrict__ carry_in, uint4 *__restrict__ carry_out,
             int carry_quads) {
    pdl_wait();
    /* boundaries are written for the cells [first_cell, cells] only (slab mode: the owned planes) and hold
     * value_offset + sorted index (slab mode: the local block sits behind the low halo in the combined list) */
    const int n = d_n ? *d_n : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0 && i == 0) fill_cells(cell_start, first_cell, cells, value_offset, gap_queue, gap_count); /* empty list: every cell empty */
    if (i >= n) return;
    const unsigned int v = svals[i];
    const unsigned int key = skeys[i];
    /* records are NQ x 16 bytes, array-of-records: one 32-byte sector serves a
     * whole 16/32-byte record of the random-access side */
#pragma unroll
    for (int q = 0; q < NQ; ++q) out_recs[(size_t)i * NQ + q] = __ldg(in_recs + (size_t)v * NQ + q);
    /* the producer's carried agent records take the same route (message v is the message of list index v) */
    for (int q = 0; q < carry_quads; ++q) carry_out[(size_t)i * carry_quads + q] = __ldg(carry_in + (size_t)v * carry_quads + q);
    if (i == 0) {
        fill_cells(cell_start, first_cell, key, value_offset, gap_queue, gap_count);
    } else {
        const unsigned int prev = skeys[i - 1];
        if (prev != key) fill_cells(cell_start, prev + 1u, key, value_offset + (unsigned int)i, gap_queue, gap_count);
    }
    if (i == n - 1) fill_cells(cell_start, key + 1u, cells, value_offset + (unsigned int)n, gap_queue, gap_count);
}

/* Runs of empty cells queued by k_gather_pbm: long runs by the whole grid, the others one block per entry. */
__global__ void __launch_bounds__(256)
k_fill_gaps(unsigned int *__restrict__ cell_start, const unsigned int *__restrict__ gap_queue,
            const unsigned int *__restrict__ gap_count) {
    pdl_wait();
    const unsigned int nlong = min(gap_count[1], (unsigned int)GAP_LONG_MAX);
    for (unsigned int e = 0; e < nlong; ++e) {
        const unsigned int lo = gap_queue[3 * e + 0], hi = gap_queue[3 * e + 1], value = gap_queue[3 * e + 2];
        for (unsigned int c = lo + blockIdx.x * blockDim.x + threadIdx.x; c <= hi; c += gridDim.x * blockDim.x) cell_start[c] = value;
    }
    const unsigned int cnt = gap_count[0];
    for (unsigned int e = blockIdx.x; e < cnt; e += gridDim.x) {
        const unsigned int *q = gap_queue + 3 * (GAP_LONG_MAX + e);
        const unsigned int lo = q[0], hi = q[1], value = q[2];
        for (unsigned int c = lo + threadIdx.x; c <= hi; c += blockDim.x) cell_start[c] = value;
    }
}

/* A column filled with one value (an agent variable the caller did not supply: its XMML defaultValue or 0) */
template <typename T>
__global__ void __launch_bounds__(256) k_fill_column(T *__restrict__ dst, T value, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = value;
}

/* Unpack one 32-bit word column of the sorted records to a dense array
 * (host read-back path only). */
__global__ void k_unpack_words(const unsigned int *__restrict__ recs, int rec_words, int n, int word, unsigned int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = recs[(size_t)i * rec_words + word];
}

/* ------------------------------------------------------------------------- */
/* stream compaction / append over SoA columns                                */
/* ------------------------------------------------------------------------- */
constexpr int MAX_COLS = 48;
struct ColSet {
    int ncols;
    unsigned int stride;              /* elements between two consecutive elements of an array variable = the list's capacity
                                         (xmachine_memory_<A>_MAX: element e of agent i lives at var[e * MAX + i], hdr:187-194) */
    int size[MAX_COLS];               /* 4 or 8 bytes */
    unsigned short len[MAX_COLS];     /* 1 for a scalar, arrayLength for an agent array variable */
    unsigned long long off[MAX_COLS]; /* byte offset of the column in the list struct */
};

/* every variable of agent `i` of one list to position `o` of another: scalars, and array variables element by element
 * (scatter_<A>_Agents / append_<A>_Agents / reorder_<A>_agents copy arrays with the same loop, krn:251-254,274-277,328-331) */
__device__ __forceinline__ void copy_cols(const char *__restrict__ src, const ColSet &sc, size_t i, char *__restrict__ dst,
                                          const ColSet &dc, size_t o) {
    for (int c = 0; c < sc.ncols; ++c) {
        const int L = sc.len[c];
        if (sc.size[c] == 4) {
            const unsigned int *sp = (const unsigned int *)(src + sc.off[c]) + i;
            unsigned int *dp = (unsigned int *)(dst + dc.off[c]) + o;
            for (int e = 0; e < L; ++e) dp[(size_t)e * dc.stride] = sp[(size_t)e * sc.stride];
        } else {
            const unsigned long long *sp = (const unsigned long long *)(src + sc.off[c]) + i;
            unsigned long long *dp = (unsigned long long *)(dst + dc.off[c]) + o;
            for (int e = 0; e < L; ++e) dp[(size_t)e * dc.stride] = sp[(size_t)e * sc.stride];
        }
    }
}

constexpr int COMPACT_THREADS = 1024;

/* number of flagged (== 1) elements per block */
__global__ void __launch_bounds__(COMPACT_THREADS)
k_flag_blocksum(const int *__restrict__ flags, int n_host, const int *__restrict__ d_n, unsigned int *__restrict__ block_sums,
                int match) {
    const int n = d_n ? *d_n : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (i < n) && (flags[i] == match);
    const int c = __syncthreads_count(f);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = (unsigned int)c;
}

/* exclusive scan of the block sums (single block), total -> *total */
__global__ void __launch_bounds__(1024) k_scan_blocksums(unsigned int *block_sums, int nb, int *total) {
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned int v = (i < nb) ? block_sums[i] : 0u;
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w; /* inclusive over warps */
        }
        __syncthreads();
        const unsigned int wbase = warp ? s_warp[warp - 1] : 0u;
        const unsigned int carry = s_carry;
        if (i < nb) block_sums[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (int)s_carry;
}

/* stable scatter of flagged agents: dst[dst_offset + rank] = src[i] for every column */
__global__ void __launch_bounds__(COMPACT_THREADS)
k_compact_scatter(const char *__restrict__ src, char *__restrict__ dst, const ColSet cols, const ColSet dcols,
                  const int *__restrict__ flags, const unsigned int *__restrict__ block_offsets, int dst_offset, int n_host,
                  const int *__restrict__ d_n, int match, int dst_cap, int *overflow) {
    __shared__ unsigned int s_warp[32];
    const int n = d_n ? *d_n : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool f = (i < n) && (flags[i] == match);
    const unsigned int ballot = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        const unsigned int v = s_warp[lane];
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        s_warp[lane] = inc - v;
    }
    __syncthreads();
    if (!f) return;
    const size_t out = (size_t)dst_offset + block_offsets[blockIdx.x] + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
    if (out >= (size_t)dst_cap) { /* fixed-capacity exchange buffers (slab mode): report, never write out of bounds */
        if (overflow) atomicExch(overflow, 1);
        return;
    }
    copy_cols(src, cols, (size_t)i, dst, dcols, out);
}

/* optional_message outputs (scatter_optional_<M>_messages, krn:411-428): stable compaction of the flagged
 * slots of the sparse staging list -- 16-byte records and, for partitioned messages, their cell keys -- to
 * dst_offset of the message list.  Uses the block offsets of k_flag_blocksum + k_scan_blocksums. */
template <int NQ>
__global__ void __launch_bounds__(COMPACT_THREADS)
k_compact_records(const uint4 *__restrict__ src_recs, const unsigned int *__restrict__ src_keys, uint4 *__restrict__ dst_recs,
                  unsigned int *__restrict__ dst_keys, const int *__restrict__ flags, const unsigned int *__restrict__ block_offsets,
                  int dst_offset, int n) {
    __shared__ unsigned int s_warp[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool f = (i < n) && (flags[i] == 1);
    const unsigned int ballot = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        const unsigned int v = s_warp[lane];
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        s_warp[lane] = inc - v;
    }
    __syncthreads();
    if (!f) return;
    const size_t out = (size_t)dst_offset + block_offsets[blockIdx.x] + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
#pragma unroll
    for (int q = 0; q < NQ; ++q) dst_recs[out * NQ + q] = src_recs[(size_t)i * NQ + q];
    if (src_keys != nullptr) dst_keys[out] = src_keys[i];
}

/* reorder_<A>_agents (krn:322-333) after the pairs of sort_<A>s_<S> were sorted: new[i] = old[values[order[i]]] */
__global__ void __launch_bounds__(256)
k_reorder_agents(const char *__restrict__ src, char *__restrict__ dst, const ColSet cols, const unsigned int *__restrict__ order,
                 const unsigned int *__restrict__ values, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t old_pos = values[order[i]];
    copy_cols(src, cols, old_pos, dst, cols, (size_t)i);
}

/* append_<A>_Agents (krn:265-280) */
__global__ void __launch_bounds__(256)
k_append(const char *__restrict__ src, char *__restrict__ dst, const ColSet cols, int dst_offset, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t out = (size_t)dst_offset + i;
    copy_cols(src, cols, (size_t)i, dst, cols, out);
}

/* ------------------------------------------------------------------------- */
/* 1-D slab decomposition: halo planes of the sorted message list, migrants    */
/* ------------------------------------------------------------------------- */
/* host-API reductions over one agent variable (simulation.xslt:1325-1357)    */
/* ------------------------------------------------------------------------- */
/* reduce_/count_/min_/max_<A>_<S>_<v>_variable(): thrust::reduce / count / min_element / max_element in the
 * reference.  Two launches: grid-stride partials per block in a fixed order (so a float sum is the same
 * from run to run, which thrust does not promise), then one block folds the partials. */
enum { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2, RED_COUNT = 3 };
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1024;

template <typename T, int OP>
__device__ __forceinline__ T red_combine(T a, T b) {
    if (OP == RED_MIN) return b < a ? b : a;
    if (OP == RED_MAX) return a < b ? b : a;
    return a + b;
}

template <typename T, int OP>
__device__ __forceinline__ T red_block(T v, T *s_part /*[RED_THREADS]*/) {
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_part[threadIdx.x] = red_combine<T, OP>(s_part[threadIdx.x], s_part[threadIdx.x + o]);
        __syncthreads();
    }
    return s_part[0];
}

/* identity: 0 for sum/count, the first element for min/max (n > 0 is checked by the caller) */
template <typename T, int OP>
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_partial(const T *__restrict__ col, int n, T count_value, T *__restrict__ partial) {
    __shared__ T s_part[RED_THREADS];
    T acc = (OP == RED_MIN || OP == RED_MAX) ? col[0] : (T)0;
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS) {
        const T v = col[i];
        if (OP == RED_COUNT) acc += (v == count_value) ? (T)1 : (T)0;
        else acc = red_combine<T, OP>(acc, v);
    }
    const T r = red_block<T, (OP == RED_COUNT ? RED_SUM : OP)>(acc, s_part);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

template <typename T, int OP>
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_final(const T *__restrict__ partial, int nparts, T *__restrict__ result) {
    __shared__ T s_part[RED_THREADS];
    T acc = (OP == RED_MIN || OP == RED_MAX) ? partial[0] : (T)0;
    for (int i = threadIdx.x; i < nparts; i += RED_THREADS) acc = red_combine<T, OP>(acc, partial[i]);
    const T r = red_block<T, OP>(acc, s_part);
    if (threadIdx.x == 0) *result = r;
}

/* ------------------------------------------------------------------------- */
/* Exchange buffers have a fixed capacity so that the NCCL message sizes are
 * known on the host without a device round trip; the true counts travel in a
 * 4-int header: {count, overflow, 0, 0}. */
constexpr int XHDR = 4;

/* ---- peer-memory transport ------------------------------------------------------------------------
 * In P2P mode the neighbours' receive buffers are mapped through CUDA IPC and one small kernel per
 * exchange (k_push, grid (PUSH_BLOCKS, 2): y = 0 towards the lower neighbour, y = 1 towards the upper)
 * stores the used part of the two local send buffers straight into them over NVLink, publishes a
 * sequence number in the neighbour's memory (every block fences its stores to system scope after a block
 * barrier; the last block writes the flag) and then waits until the two neighbours' flags for the same
 * exchange have arrived.  No collective launch, no staging copy on the receiving side, and only the
 * bytes actually used travel (the NCCL fallback has to send the fixed capacity).  The heavy producer
 * kernels stay free of system-scope fences: fencing inside them (one fence per block of a 1280-block
 * grid) cost 35-45 us per exchange (profiles/r01_notes.md).  Receive buffers are double-buffered by
 * sequence parity; a rank can never be two exchanges ahead of a neighbour because each exchange also
 * waits for that neighbour's data. */
__device__ __forceinline__ void st_flag_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_flag_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int PUSH_BLOCKS = 24;
constexpr int PUSH_MAX_SEG = 48;
/* A fixed-capacity exchange buffer = `fixed_bytes` always-used bytes (header, PBM slice) followed by
 * segments of `cap` elements of which the first hdr[0] are used. */
struct PushLayout {
    int fixed_bytes; /* multiple of 16 */
    int nseg;
    int elem[PUSH_MAX_SEG];                /* element size in bytes (4, 8 or a multiple of 16) */
    unsigned long long off[PUSH_MAX_SEG];  /* byte offset of the segment, 16-byte aligned */
};

__device__ __forceinline__ void push_bytes(const char *__restrict__ src, char *__restrict__ dst, size_t bytes, int tid, int nthreads) {
    /* bytes is a multiple of 4; bulk in 16-byte pieces, tail in words */
    const size_t q = bytes / 16;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (size_t i = tid; i < q; i += nthreads) d4[i] = s4[i];
    const size_t w0 = q * 4, w1 = bytes / 4;
    for (size_t i = w0 + tid; i < w1; i += nthreads) ((unsigned int *)dst)[i] = ((const unsigned int *)src)[i];
}

/* End of a push kernel: every block fences its peer stores; the last block to arrive publishes the
 * exchange number in both neighbours' memory and waits (bounded) until theirs have arrived here. */
__device__ __forceinline__ void push_signal_and_wait(unsigned int *done_counter, int *peer_flag_lo, int *peer_flag_hi, const int *my_flag_a,
                                                     const int *my_flag_b, int *seq_counter, int *status) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(done_counter, 1u);
        if (t == gridDim.x * gridDim.y - 1) {
            *done_counter = 0;
            /* the exchange number lives on the device (one counter per exchange stream, touched by this
             * thread only), so a captured CUDA graph of the iteration can be replayed */
            const int seq = *seq_counter + 1;
            *seq_counter = seq;
            /* every block fenced its stores (system scope) before it was counted; the release stores below are
             * cumulative over what this thread has observed, so no second fence is needed (it cost ~3 us) */
            st_flag_sys(peer_flag_lo, seq);
            st_flag_sys(peer_flag_hi, seq);
            /* ... and wait for the neighbours' data of the same exchange (bounded spin) */
            const long long t0 = clock64();
            while (ld_flag_sys(my_flag_a) < seq || ld_flag_sys(my_flag_b) < seq) {
                if (clock64() - t0 > 4000000000LL) { /* ~2 s: report instead of hanging the GPU */
                    atomicExch(status, 5);
                    break;
                }
                __nanosleep(100);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_push(const char *__restrict__ send_lo, const char *__restrict__ send_hi, char *__restrict__ peer_lo, char *__restrict__ peer_hi,
       const PushLayout L, unsigned int *done_counter, int *peer_flag_lo, int *peer_flag_hi, const int *my_flag_a,
       const int *my_flag_b, int *seq_counter, int *status) {
    pdl_wait();
    const char *src = blockIdx.y ? send_hi : send_lo;
    char *dst = blockIdx.y ? peer_hi : peer_lo;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int cnt = ((const int *)src)[0];
    push_bytes(src, dst, (size_t)L.fixed_bytes, tid, nthreads);
    for (int k = 0; k < L.nseg; ++k) push_bytes(src + L.off[k], dst + L.off[k], (size_t)cnt * L.elem[k], tid, nthreads);
    push_signal_and_wait(done_counter, peer_flag_lo, peer_flag_hi, my_flag_a, my_flag_b, seq_counter, status);
}

/* plane -> send buffer: header | (plane_cells + 1) cell_start entries rebased to 0 | records */
template <int NQ>
__global__ void __launch_bounds__(256)
k_pack_planes(const uint4 *__restrict__ sorted_local, const unsigned int *__restrict__ cell_start, int first_cell_lo,
              int first_cell_hi, int plane_cells, int cap, int *__restrict__ out_lo, int *__restrict__ out_hi) {
    const int first_cell = blockIdx.y ? first_cell_hi : first_cell_lo;
    int *out = blockIdx.y ? out_hi : out_lo;
    const int lo = (int)cell_start[first_cell], hi = (int)cell_start[first_cell + plane_cells];
    const int cnt = hi - lo;
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) {
        out[0] = min(cnt, cap);
        out[1] = cnt > cap;
        out[2] = 0;
        out[3] = 0;
    }
    int *cs = out + XHDR;
    for (int i = tid; i <= plane_cells; i += stride) cs[i] = min((int)cell_start[first_cell + i] - lo, cap);
    uint4 *recs = reinterpret_cast<uint4 *>(out + XHDR + ((plane_cells + 1 + 3) & ~3));
    const int quads = min(cnt, cap) * NQ;
    for (int i = tid; i < quads; i += stride) recs[i] = sorted_local[(size_t)lo * NQ + i];
}

/* Peer-memory transport of the two boundary planes in ONE kernel: what k_pack_planes writes into a local send
 * buffer goes straight into the neighbours' receive buffers (same layout), followed by the signal/wait of
 * k_push.  y = 0: my first owned plane -> the lower neighbour's HIGH halo; y = 1: my last -> the upper's LOW. */
template <int NQ>
__global__ void __launch_bounds__(256)
k_push_planes(const uint4 *__restrict__ sorted_combined, const unsigned int *__restrict__ cell_start, int first_cell_lo, int first_cell_hi,
              int plane_cells, int cap, int *__restrict__ peer_lo, int *__restrict__ peer_hi, unsigned int *done_counter,
              int *peer_flag_lo, int *peer_flag_hi, const int *my_flag_a, const int *my_flag_b, int *seq_counter, int *status) {
    pdl_wait();
    const int first_cell = blockIdx.y ? first_cell_hi : first_cell_lo;
    int *out = blockIdx.y ? peer_hi : peer_lo;
    const int lo = (int)cell_start[first_cell], hi = (int)cell_start[first_cell + plane_cells];
    const int cnt = hi - lo;
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) *reinterpret_cast<int4 *>(out) = make_int4(min(cnt, cap), cnt > cap, 0, 0);
    int *cs = out + XHDR;
    for (int i = tid; i <= plane_cells; i += stride) cs[i] = min((int)cell_start[first_cell + i] - lo, cap);
    uint4 *recs = reinterpret_cast<uint4 *>(out + XHDR + ((plane_cells + 1 + 3) & ~3));
    const int quads = min(cnt, cap) * NQ;
    for (int i = tid; i < quads; i += stride) recs[i] = sorted_combined[(size_t)lo * NQ + i];
    push_signal_and_wait(done_counter, peer_flag_lo, peer_flag_hi, my_flag_a, my_flag_b, seq_counter, status);
}

/* Merge: shift the owned planes' cell_start by the position of the local block
 * inside the combined buffer, splice in the two received halo planes.
 *   combined buffer: [ ... low halo | local block (n records at `local_at`) | high halo ... ] */
template <int NQ>
__global__ void __launch_bounds__(256)
k_merge_halo(uint4 *__restrict__ combined, unsigned int *__restrict__ cell_start, int local_at, int n_host,
             const int *__restrict__ d_n, int own_first_cell, int own_cells, const int *__restrict__ lo_buf,
             int lo_first_cell, const int *__restrict__ hi_buf, int hi_first_cell, int plane_cells, int *status) {
    pdl_wait();
    const int n = d_n ? *d_n : n_host;
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lo_cnt = lo_buf[0], hi_cnt = hi_buf[0];
    if (tid == 0 && (lo_buf[1] || hi_buf[1])) atomicExch(status, 1);
    const int pad = (plane_cells + 1 + 3) & ~3;
    /* the owned planes' entries (closing boundary included) were written with local_at added by the build */
    /* low halo sits immediately below the local block, high halo immediately above */
    const int lo_at = local_at - lo_cnt, hi_at = local_at + n;
    /* a halo plane adjacent in hash order shares its boundary entry with the owned block; that entry already
     * holds the right value (local_at or local_at + n), so the shared entry is skipped */
    for (int i = tid; i <= plane_cells; i += stride) {
        const int c_lo = lo_first_cell + i;
        if (!(c_lo >= own_first_cell && c_lo <= own_first_cell + own_cells)) cell_start[c_lo] = (unsigned int)(lo_at + lo_buf[XHDR + i]);
        const int c_hi = hi_first_cell + i;
        if (!(c_hi >= own_first_cell && c_hi <= own_first_cell + own_cells)) cell_start[c_hi] = (unsigned int)(hi_at + hi_buf[XHDR + i]);
    }
    const uint4 *lo_recs = reinterpret_cast<const uint4 *>(lo_buf + XHDR + pad);
    const uint4 *hi_recs = reinterpret_cast<const uint4 *>(hi_buf + XHDR + pad);
    for (int i = tid; i < lo_cnt * NQ; i += stride) combined[(size_t)lo_at * NQ + i] = lo_recs[i];
    for (int i = tid; i < hi_cnt * NQ; i += stride) combined[(size_t)hi_at * NQ + i] = hi_recs[i];
}

/* ---- fused three-way migration split (stay / down / up) ------------------------------------------ */
/* Owner class of an agent along the slab axis: 1 = stays, 2 = lower neighbour, 3 = upper neighbour,
 * 4 = further than one slab away (reported).  The plane is the cell coordinate of
 * message_<M>_grid_position + the wrap rule of message_<M>_hash (krn:860-889). */
/* pass 1: owner flag of every agent + per-block counts of the three classes: block_sums[class][block].
 * A block covers COMPACT_THREADS consecutive agents (the unit of the stable scatter below) with 256 threads
 * x 4 rounds: a quarter of the thread launches of one-thread-per-agent. */
__global__ void __launch_bounds__(256)
k_owner_count(const float *__restrict__ coord, int n_host, const int *__restrict__ d_n, float minb, float maxb, int nplanes,
              int planes_per_rank, int rank, int world, int *__restrict__ flags, unsigned int *__restrict__ block_sums,
              int nb, int *status) {
    __shared__ unsigned int s_cnt[3];
    const int n = d_n ? *d_n : n_host;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lo_rank = (rank + world - 1) % world, hi_rank = (rank + 1) % world;
    const float scale = (float)nplanes, range = __fsub_rn(maxb, minb);
    unsigned int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int r = 0; r < COMPACT_THREADS / 256; ++r) {
        const int i = blockIdx.x * COMPACT_THREADS + r * 256 + threadIdx.x;
        int f = 0;
        if (i < n) {
            int p = (int)floorf(__fdiv_rn(__fmul_rn(__fsub_rn(coord[i], minb), scale), range));
            p = (p < 0) ? nplanes - 1 : p;
            p = (p >= nplanes) ? 0 : p;
            const int owner = min(p / planes_per_rank, world - 1);
            f = 4;
            if (owner == rank) f = 1;
            else if (owner == lo_rank && owner == hi_rank) f = (p < rank * planes_per_rank) ? 2 : 3; /* world == 2 */
            else if (owner == lo_rank) f = 2;
            else if (owner == hi_rank) f = 3;
            flags[i] = f;
            if (f == 4) atomicExch(status, 2);
        }
        c1 += __popc(__ballot_sync(0xffffffffu, f == 1));
        c2 += __popc(__ballot_sync(0xffffffffu, f == 2));
        c3 += __popc(__ballot_sync(0xffffffffu, f == 3));
    }
    if ((threadIdx.x & 31) == 0) {
        if (c1) atomicAdd(&s_cnt[0], c1);
        if (c2) atomicAdd(&s_cnt[1], c2);
        if (c3) atomicAdd(&s_cnt[2], c3);
    }
    __syncthreads();
    if (threadIdx.x < 3) block_sums[threadIdx.x * nb + blockIdx.x] = s_cnt[threadIdx.x];
}

/* pass 2: stable scatter of every agent to its destination (stayers list / down buffer / up buffer).
 * The offset of a block is the sum of the counts of the blocks before it, which every block adds up for
 * itself (at most a few thousand values; no scan launch, no chain between blocks); the last block also
 * publishes the totals: *n_stay and the headers of the two send buffers. */
__global__ void __launch_bounds__(COMPACT_THREADS)
k_scatter3(const char *__restrict__ src, char *__restrict__ stay, char *__restrict__ dn, char *__restrict__ up,
           const ColSet cols, const ColSet bcols, const int *__restrict__ flags, const unsigned int *__restrict__ block_sums,
           int nb, int n_host, const int *__restrict__ d_n, int cap, int *overflow, int *n_stay, int *hdr_dn, int *hdr_up) {
    __shared__ unsigned int s_warp[3][32];
    __shared__ unsigned int s_off[3];
    const int n = d_n ? *d_n : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) s_off[threadIdx.x] = 0;
    __syncthreads();
    {
        unsigned int p0 = 0, p1 = 0, p2 = 0;
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += COMPACT_THREADS) {
            p0 += block_sums[b];
            p1 += block_sums[nb + b];
            p2 += block_sums[2 * nb + b];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
            p2 += __shfl_xor_sync(0xffffffffu, p2, o);
        }
        if (lane == 0) {
            if (p0) atomicAdd(&s_off[0], p0);
            if (p1) atomicAdd(&s_off[1], p1);
            if (p2) atomicAdd(&s_off[2], p2);
        }
    }
    __syncthreads();
    if (blockIdx.x == (unsigned int)nb - 1 && threadIdx.x < 3) {
        const int total = (int)(s_off[threadIdx.x] + block_sums[threadIdx.x * nb + blockIdx.x]);
        if (threadIdx.x == 0) *n_stay = total;
        else {
            int *h = (threadIdx.x == 1) ? hdr_dn : hdr_up;
            h[0] = min(total, cap);
            h[1] = total > cap;
            h[2] = 0;
            h[3] = 0;
        }
    }
    const int f = (i < n) ? flags[i] : 0;
    const unsigned int b1 = __ballot_sync(0xffffffffu, f == 1), b2 = __ballot_sync(0xffffffffu, f == 2),
                       b3 = __ballot_sync(0xffffffffu, f == 3);
    if (lane == 0) {
        s_warp[0][warp] = __popc(b1);
        s_warp[1][warp] = __popc(b2);
        s_warp[2][warp] = __popc(b3);
    }
    __syncthreads();
    if (warp < 3) {
        const unsigned int v = s_warp[warp][lane];
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        s_warp[warp][lane] = inc - v;
    }
    __syncthreads();
    if (f >= 1 && f <= 3) {
        const unsigned int lt = (1u << lane) - 1u;
        const unsigned int bal = (f == 1) ? b1 : (f == 2) ? b2 : b3;
        const size_t out = (size_t)s_off[f - 1] + s_warp[f - 1][warp] + __popc(bal & lt);
        char *dst = (f == 1) ? stay : (f == 2) ? dn : up;
        if (f != 1 && out >= (size_t)cap) {
            atomicExch(overflow, 1);
        } else {
            for (int c = 0; c < cols.ncols; ++c) {
                const unsigned long long doff = (f == 1) ? cols.off[c] : bcols.off[c];
                if (cols.size[c] == 4)
                    ((unsigned int *)(dst + doff))[out] = ((const unsigned int *)(src + cols.off[c]))[i];
                else
                    ((unsigned long long *)(dst + doff))[out] = ((const unsigned long long *)(src + cols.off[c]))[i];
            }
        }
    }
}

/* ---- in-place migration ------------------------------------------------------------------------------
 * Leavers are a fraction of a percent of a slab per iteration, so the list is NOT re-compacted (k_scatter3 above
 * copies every stayer, 40 bytes per agent per iteration at 5 variables): leavers are packed into the send
 * buffers and their list slots recorded as holes; after the exchange arrival j takes hole j, surplus arrivals are
 * appended, and surplus holes are filled with the live agents of the list's tail.  Every step is a function of
 * list positions only, so the resulting order does not depend on timing, transport or launch mode.
 *   k_mig_count   owner class of every agent, leavers (down / up) per block of 1024 agents
 *   k_mig_scan    exclusive offsets of the blocks, totals -> send headers + hole count
 *   k_mig_pack    leavers -> send buffers (stable), their slots -> holes[] (ascending)
 *   (transport)
 *   k_mig_fill    arrivals -> holes / end of the list;   k_mig_tail: surplus holes <- tail agents, new count */
__device__ __forceinline__ int owner_class(float c, float minb, float range, float scale, int nplanes, int ppr, int rank, int world,
                                           int lo_rank, int hi_rank) {
    int p = (int)floorf(__fdiv_rn(__fmul_rn(__fsub_rn(c, minb), scale), range));
    p = (p < 0) ? nplanes - 1 : p;
    p = (p >= nplanes) ? 0 : p;
    const int p0 = rank * ppr;
    if (p >= p0 && (p < p0 + ppr || rank == world - 1)) return 1;   /* the common case: no integer division */
    const int owner = min(p / ppr, world - 1);
    if (owner == lo_rank && owner == hi_rank) return (p < p0) ? 2 : 3; /* world == 2 */
    if (owner == lo_rank) return 2;
    if (owner == hi_rank) return 3;
    return 4;
}

/* header of the migration bookkeeping (ints): {n_old, holes, arrivals_used, -} */
constexpr int MIG_HDR = 4;

__global__ void __launch_bounds__(256)
k_mig_count(const float *__restrict__ coord, int n_host, const int *__restrict__ d_n, float minb, float maxb, int nplanes,
            int ppr, int rank, int world, unsigned int *__restrict__ block_sums, int nb, int *status) {
    pdl_wait();
    __shared__ unsigned int s_cnt[2];
    const int n = d_n ? *d_n : n_host;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lo_rank = (rank + world - 1) % world, hi_rank = (rank + 1) % world;
    const float scale = (float)nplanes, range = __fsub_rn(maxb, minb);
    unsigned int c2 = 0, c3 = 0;
#pragma unroll
    for (int r = 0; r < COMPACT_THREADS / 256; ++r) {
        const int i = blockIdx.x * COMPACT_THREADS + r * 256 + threadIdx.x;
        const int f = (i < n) ? owner_class(coord[i], minb, range, scale, nplanes, ppr, rank, world, lo_rank, hi_rank) : 0;
        if (f == 4) atomicExch(status, 2);
        c2 += __popc(__ballot_sync(0xffffffffu, f == 2));
        c3 += __popc(__ballot_sync(0xffffffffu, f == 3));
    }
    if ((threadIdx.x & 31) == 0) {
        if (c2) atomicAdd(&s_cnt[0], c2);
        if (c3) atomicAdd(&s_cnt[1], c3);
    }
    __syncthreads();
    if (threadIdx.x < 2) block_sums[threadIdx.x * nb + blockIdx.x] = s_cnt[threadIdx.x];
}

/* Pack: a block covers the same 1024 agents as in k_mig_count (256 threads x 4 rounds).  Its offsets in the two send
 * buffers are the sums of the counts of the blocks before it, which every block adds up for itself (a few thousand
 * values from L2; no scan launch, no chain between blocks); the last block also publishes the totals: the send
 * headers and {n_old, holes}.  Leavers are written in list order (stable), their slots go to holes[] (ascending). */
__device__ __forceinline__ void
mig_pack_body(const char *__restrict__ list, char *__restrict__ dn, char *__restrict__ up, const ColSet &cols, const ColSet &bcols,
              const float *__restrict__ coord, int n_host, const int *__restrict__ d_n, float minb, float maxb, int nplanes, int ppr,
              int rank, int world, const unsigned int *__restrict__ block_sums, int nb, int cap, unsigned int *__restrict__ holes,
              int *__restrict__ mig_hdr, int *__restrict__ hdr_dn, int *__restrict__ hdr_up, int *overflow) {
    constexpr int ROUNDS = COMPACT_THREADS / 256;
    __shared__ unsigned int s_off[2];
    __shared__ unsigned int s_cnt[2][ROUNDS * 8];
    const int n = d_n ? *d_n : n_host;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int mine_dn = block_sums[blockIdx.x], mine_up = block_sums[nb + blockIdx.x];
    const bool last = blockIdx.x == (unsigned int)nb - 1;
    if (mine_dn + mine_up == 0 && !last) return; /* nothing leaves from these 1024 agents */
    if (threadIdx.x < 2) s_off[threadIdx.x] = 0;
    __syncthreads();
    {
        unsigned int p0 = 0, p1 = 0;
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += 256) {
            p0 += block_sums[b];
            p1 += block_sums[nb + b];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
        }
        if (lane == 0) {
            if (p0) atomicAdd(&s_off[0], p0);
            if (p1) atomicAdd(&s_off[1], p1);
        }
    }
    __syncthreads();
    const unsigned int off_dn = s_off[0], off_up = s_off[1];
    if (last && threadIdx.x == 0) {
        const int tdn = (int)(off_dn + mine_dn), tup = (int)(off_up + mine_up);
        if (tdn > cap || tup > cap) atomicExch(overflow, 1);
        hdr_dn[0] = min(tdn, cap); hdr_dn[1] = tdn > cap; hdr_dn[2] = 0; hdr_dn[3] = 0;
        hdr_up[0] = min(tup, cap); hdr_up[1] = tup > cap; hdr_up[2] = 0; hdr_up[3] = 0;
        mig_hdr[0] = n;
        mig_hdr[1] = min(tdn + tup, 2 * cap);   /* never beyond the hole list; on overflow the status flag aborts the run */
        mig_hdr[2] = 0;
        mig_hdr[3] = 0;
    }
    if (mine_dn + mine_up == 0) return;
    const int lo_rank = (rank + world - 1) % world, hi_rank = (rank + 1) % world;
    const float scale = (float)nplanes, range = __fsub_rn(maxb, minb);
    int f[ROUNDS];
    unsigned int b2[ROUNDS], b3[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int i = blockIdx.x * COMPACT_THREADS + r * 256 + threadIdx.x;
        f[r] = (i < n) ? owner_class(coord[i], minb, range, scale, nplanes, ppr, rank, world, lo_rank, hi_rank) : 0;
        b2[r] = __ballot_sync(0xffffffffu, f[r] == 2);
        b3[r] = __ballot_sync(0xffffffffu, f[r] == 3);
        if (lane == 0) {
            s_cnt[0][r * 8 + warp] = __popc(b2[r]);
            s_cnt[1][r * 8 + warp] = __popc(b3[r]);
        }
    }
    __syncthreads();
    if (warp < 2) { /* exclusive prefix over (round, warp) = list order inside the block */
        const unsigned int v = (lane < ROUNDS * 8) ? s_cnt[warp][lane] : 0u;
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane < ROUNDS * 8) s_cnt[warp][lane] = inc - v;
    }
    __syncthreads();
    const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if (f[r] != 2 && f[r] != 3) continue;
        const int i = blockIdx.x * COMPACT_THREADS + r * 256 + threadIdx.x;
        const unsigned int dn_before = off_dn + s_cnt[0][r * 8 + warp] + __popc(b2[r] & lt);
        const unsigned int up_before = off_up + s_cnt[1][r * 8 + warp] + __popc(b3[r] & lt);
        const unsigned int hole = dn_before + up_before;             /* leavers with a smaller list index */
        if (hole < 2u * (unsigned int)cap) holes[hole] = (unsigned int)i;
        const unsigned int out = (f[r] == 2) ? dn_before : up_before;
        if (out >= (unsigned int)cap) continue;                      /* overflow was flagged above */
        char *dst = (f[r] == 2) ? dn : up;
        for (int c = 0; c < cols.ncols; ++c) {
            if (cols.size[c] == 4)
                ((unsigned int *)(dst + bcols.off[c]))[out] = ((const unsigned int *)(list + cols.off[c]))[i];
            else
                ((unsigned long long *)(dst + bcols.off[c]))[out] = ((const unsigned long long *)(list + cols.off[c]))[i];
        }
    }
}

__global__ void __launch_bounds__(256)
k_mig_pack(const char *__restrict__ list, char *__restrict__ dn, char *__restrict__ up, const ColSet cols, const ColSet bcols,
           const float *__restrict__ coord, int n_host, const int *__restrict__ d_n, float minb, float maxb, int nplanes, int ppr,
           int rank, int world, const unsigned int *__restrict__ block_sums, int nb, int cap, unsigned int *__restrict__ holes,
           int *__restrict__ mig_hdr, int *__restrict__ hdr_dn, int *__restrict__ hdr_up, int *overflow,
           /* direct mode (done_counter != nullptr): dn / up / hdr_* are the NEIGHBOURS' arrival buffers (peer memory), so the
            * leavers cross NVLink as they are packed; the last block to finish publishes the exchange number to both neighbours
            * and waits for theirs -- what k_push did in a kernel of its own */
           unsigned int *done_counter, int *peer_flag_lo, int *peer_flag_hi, const int *my_flag_a, const int *my_flag_b,
           int *seq_counter, int *status) {
    pdl_wait();
    mig_pack_body(list, dn, up, cols, bcols, coord, n_host, d_n, minb, maxb, nplanes, ppr, rank, world, block_sums, nb, cap, holes, mig_hdr,
                  hdr_dn, hdr_up, overflow);
    if (done_counter != nullptr) push_signal_and_wait(done_counter, peer_flag_lo, peer_flag_hi, my_flag_a, my_flag_b, seq_counter, status);
}


__device__ __forceinline__ void copy_agent(const char *src, const ColSet &scols, size_t si, char *dst, const ColSet &dcols, size_t di) {
    for (int c = 0; c < scols.ncols; ++c) {
        if (scols.size[c] == 4)
            ((unsigned int *)(dst + dcols.off[c]))[di] = ((const unsigned int *)(src + scols.off[c]))[si];
        else
            ((unsigned long long *)(dst + dcols.off[c]))[di] = ((const unsigned long long *)(src + scols.off[c]))[si];
    }
}

/* After the exchange, one kernel.  Arrivals (lower neighbour's first): arrival j takes hole j, those beyond the holes
 * go to the end of the list (all blocks, grid-stride).  Holes the arrivals do not fill are filled with the live agents
 * of the tail, ascending hole <- ascending tail agent, by block 0, which also publishes the new population; the two
 * jobs touch disjoint slots (the arrivals take the smallest holes, the tail job the largest ones and the tail itself). */
__global__ void __launch_bounds__(1024)
k_mig_fill(const char *__restrict__ rlo, const char *__restrict__ rhi, char *__restrict__ list, const ColSet scols, const ColSet dcols,
           const int *__restrict__ mig_hdr, const unsigned int *__restrict__ holes, int *__restrict__ n_new_out, int dst_cap, int *status) {
    pdl_wait();
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    const int *hlo = (const int *)rlo, *hhi = (const int *)rhi;
    const int clo = hlo[0], chi = hhi[0], n_old = mig_hdr[0], nh = mig_hdr[1];
    const int arrivals = clo + chi;
    if (blockIdx.x == 0 && threadIdx.x == 0 && (hlo[1] || hhi[1])) atomicExch(status, 3);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < arrivals; j += gridDim.x * blockDim.x) {
        const char *src = (j < clo) ? rlo : rhi;
        const int sj = (j < clo) ? j : j - clo;
        const long long dst = (j < nh) ? (long long)holes[j] : (long long)n_old + (j - nh);
        if (dst >= (long long)dst_cap) {
            atomicExch(status, 4);
            continue;
        }
        copy_agent(src, scols, (size_t)sj, list, dcols, (size_t)dst);
    }
    if (blockIdx.x != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (arrivals >= nh) {
        if (threadIdx.x == 0) *n_new_out = min(n_old + (arrivals - nh), dst_cap);
        return;
    }
    const int T = nh - arrivals, n_new = n_old - T;
    /* holes[arrivals .. nh) are the T largest holes, ascending; those below n_new need an agent, the others lie in the
     * tail [n_new, n_old) that is cut off */
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int t = base + threadIdx.x;
        const int pos = n_new + t;                    /* tail slot */
        bool live = false;
        if (t < T) {
            int lo = arrivals, hi = nh;               /* is `pos` a hole?  binary search in the ascending hole list */
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (holes[mid] < (unsigned int)pos) lo = mid + 1;
                else hi = mid;
            }
            live = !(lo < nh && holes[lo] == (unsigned int)pos);
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const unsigned int v = s_warp[lane];
            unsigned int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            s_warp[lane] = inc - v;
        }
        __syncthreads();
        const unsigned int carry = s_carry;
        if (live) {
            const unsigned int r = carry + s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
            copy_agent(list, dcols, (size_t)pos, list, dcols, (size_t)holes[arrivals + r]);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + __popc(bal);
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_new_out = n_new;
}

/* pass 4 (after the exchange): append both received buffers after the stayers and publish the new population */
__global__ void __launch_bounds__(256)
k_append_both(const char *__restrict__ rlo, const char *__restrict__ rhi, char *__restrict__ dst, const ColSet scols,
              const ColSet dcols, const int *__restrict__ n_stay, int *__restrict__ n_new, int dst_cap, int *status) {
    const int *hlo = (const int *)rlo, *hhi = (const int *)rhi;
    const int clo = hlo[0], chi = hhi[0], base = *n_stay;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        if (hlo[1] || hhi[1]) atomicExch(status, 3);
        if ((long long)base + clo + chi > (long long)dst_cap) atomicExch(status, 4);
        *n_new = min(base + clo + chi, dst_cap);
    }
    if (i >= clo + chi) return;
    const char *src = (i < clo) ? rlo : rhi;
    const int si = (i < clo) ? i : i - clo;
    const size_t out = (size_t)base + i;
    if (out >= (size_t)dst_cap) return;
    for (int c = 0; c < scols.ncols; ++c) {
        if (scols.size[c] == 4)
            ((unsigned int *)(dst + dcols.off[c]))[out] = ((const unsigned int *)(src + scols.off[c]))[si];
        else
            ((unsigned long long *)(dst + dcols.off[c]))[out] = ((const unsigned long long *)(src + scols.off[c]))[si];
    }
}

}  // namespace fgpu
#endif
