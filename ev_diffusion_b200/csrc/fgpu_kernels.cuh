/*
 * fgpu_kernels.cuh -- model-agnostic sm_100a kernels of the simulation step.
 *
 *   (1) spatial message build: LSD radix sort of (cell key, message index) with
 *       one global histogram pass + one "onesweep" scatter pass per digit
 *       (decoupled look-back between tiles), then a fused gather of the packed
 *       message records + partition-boundary (CSR cell_start) construction.
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

/* ------------------------------------------------------------------------- */
/* radix sort                                                                 */
/* ------------------------------------------------------------------------- */
constexpr int SORT_MAXBITS = 8;
constexpr int SORT_BINS = 1 << SORT_MAXBITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; /* 4096 keys per tile */
constexpr int SORT_MAX_PASSES = 4;

constexpr unsigned int ST_PREFIX = 0x80000000u; /* inclusive prefix published */
constexpr unsigned int ST_AGG = 0x40000000u;    /* tile aggregate published */
constexpr unsigned int ST_VAL = 0x3fffffffu;

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* Digit histograms of every pass in one read of the keys.
 * hist layout: [pass][SORT_BINS]. */
__global__ void __launch_bounds__(256) k_radix_hist(const unsigned int *__restrict__ keys, int n,
                                                    unsigned int *__restrict__ hist, int passes, int digit_bits) {
    __shared__ unsigned int sh[SORT_MAX_PASSES * SORT_BINS];
    for (int i = threadIdx.x; i < passes * SORT_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned int mask = (1u << digit_bits) - 1u;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned int k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * SORT_BINS + ((k >> (p * digit_bits)) & mask)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * SORT_BINS; i += blockDim.x) {
        const unsigned int v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
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

/* One LSD pass.  Tiles take tickets so that a tile only ever waits on tiles
 * that are already running (forward progress of the look-back).
 *   vals_in == nullptr : values are the identity (first pass)
 * Stability: within a tile items are ranked in (warp, round, lane) order, which
 * is ascending input index; tiles are ordered by ticket == input order. */
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_pass(const unsigned int *__restrict__ keys_in, const unsigned int *__restrict__ vals_in,
             unsigned int *__restrict__ keys_out, unsigned int *__restrict__ vals_out, int n, int shift,
             int digit_bits, const unsigned int *__restrict__ hist, unsigned int *status, unsigned int *ticket) {
    __shared__ unsigned int s_warp_hist[SORT_WARPS][SORT_BINS];
    __shared__ unsigned int s_tile_excl[SORT_BINS];
    __shared__ int s_goff[SORT_BINS];
    __shared__ unsigned int s_keys[SORT_TILE];
    __shared__ unsigned int s_vals[SORT_TILE];
    __shared__ unsigned int s_scan[SORT_WARPS];
    __shared__ unsigned int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * SORT_BINS; i += SORT_THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned int tile = s_tile;
    const int base = (int)tile * SORT_TILE;
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
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int idx = base + warp * (SORT_ITEMS * 32) + r * 32 + lane;
        const bool valid = idx < n;
        const unsigned int d = valid ? ((key[r] >> shift) & mask) : 0xffffffffu;
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned int prev = 0;
        if (lane == leader && valid) {
            prev = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = (unsigned short)(prev + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    /* per digit: exclusive offsets across warps, tile count */
    unsigned int count = 0;
    if (tid < nbins) {
        unsigned int run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const unsigned int t = s_warp_hist[w][tid];
            s_warp_hist[w][tid] = run;
            run += t;
        }
        count = run;
    }
    const unsigned int tile_excl = block_exclusive_scan_256(count, s_scan);
    const unsigned int gcount = (tid < nbins) ? hist[tid] : 0u;
    const unsigned int digit_base = block_exclusive_scan_256(gcount, s_scan);

    if (tid < nbins) {
        unsigned int *st = status + (size_t)tile * SORT_BINS;
        unsigned int excl = 0;
        if (tile == 0) {
            st_volatile_u32(st + tid, ST_PREFIX | count);
        } else {
            st_volatile_u32(st + tid, ST_AGG | count);
            for (int t = (int)tile - 1;; --t) {
                const unsigned int *p = status + (size_t)t * SORT_BINS + tid;
                unsigned int v;
                do {
                    v = ld_volatile_u32(p);
                } while ((v & (ST_PREFIX | ST_AGG)) == 0u);
                excl += v & ST_VAL;
                if (v & ST_PREFIX) break;
            }
            st_volatile_u32(st + tid, ST_PREFIX | (excl + count));
        }
        s_tile_excl[tid] = tile_excl;
        s_goff[tid] = (int)(digit_base + excl) - (int)tile_excl;
    }
    __syncthreads();

    /* rank -> position inside the tile, staged in shared memory so that the
     * global writes below are contiguous runs per digit */
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int idx = base + warp * (SORT_ITEMS * 32) + r * 32 + lane;
        if (idx < n) {
            const unsigned int d = (key[r] >> shift) & mask;
            const unsigned int lp = s_tile_excl[d] + s_warp_hist[warp][d] + rank[r];
            s_keys[lp] = key[r];
            s_vals[lp] = val[r];
        }
    }
    __syncthreads();
    const int valid_count = min(SORT_TILE, n - base);
    for (int i = tid; i < valid_count; i += SORT_THREADS) {
        const unsigned int k = s_keys[i];
        const int dst = s_goff[(k >> shift) & mask] + i;
        keys_out[dst] = k;
        vals_out[dst] = s_vals[i];
    }
}

/* ------------------------------------------------------------------------- */
/* gather of packed records + partition boundary matrix                       */
/* ------------------------------------------------------------------------- */
constexpr int GAP_INLINE = 32;

__device__ __forceinline__ void fill_cells(unsigned int *cell_start, unsigned int lo, unsigned int hi /*inclusive*/,
                                           unsigned int value, unsigned int *gap_queue, unsigned int *gap_count) {
    if (hi < lo) return;
    if (hi - lo < (unsigned int)GAP_INLINE) {
        for (unsigned int c = lo; c <= hi; ++c) cell_start[c] = value;
    } else {
        const unsigned int e = atomicAdd(gap_count, 1u);
        gap_queue[3 * e + 0] = lo;
        gap_queue[3 * e + 1] = hi;
        gap_queue[3 * e + 2] = value;
    }
}

/* Thread i owns sorted position i: moves message order[i] there and, where the
 * cell key changes, writes cell_start for every cell in (prev_key, key].
 * cell_start[c] = first sorted index of cell c = number of messages in cells < c;
 * cell_start[cells] = n.  Empty cells get start == next start (count 0). */
template <int NQ>
__global__ void __launch_bounds__(256)
k_gather_pbm(const unsigned int *__restrict__ skeys, const unsigned int *__restrict__ svals,
             const uint4 *__restrict__ in_recs, uint4 *__restrict__ out_recs, int n,
             unsigned int *__restrict__ cell_start, unsigned int cells, unsigned int *gap_queue,
             unsigned int *gap_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int v = svals[i];
    const unsigned int key = skeys[i];
    /* records are NQ x 16 bytes, array-of-records: one 32-byte sector serves a
     * whole 16/32-byte record of the random-access side */
#pragma unroll
    for (int q = 0; q < NQ; ++q) out_recs[(size_t)i * NQ + q] = __ldg(in_recs + (size_t)v * NQ + q);
    if (i == 0) {
        fill_cells(cell_start, 0u, key, 0u, gap_queue, gap_count);
    } else {
        const unsigned int prev = skeys[i - 1];
        if (prev != key) fill_cells(cell_start, prev + 1u, key, (unsigned int)i, gap_queue, gap_count);
    }
    if (i == n - 1) fill_cells(cell_start, key + 1u, cells, (unsigned int)n, gap_queue, gap_count);
}

/* Long runs of empty cells queued by k_gather_pbm: one block per queue entry. */
__global__ void __launch_bounds__(256)
k_fill_gaps(unsigned int *__restrict__ cell_start, const unsigned int *__restrict__ gap_queue,
            const unsigned int *__restrict__ gap_count) {
    const unsigned int cnt = *gap_count;
    for (unsigned int e = blockIdx.x; e < cnt; e += gridDim.x) {
        const unsigned int lo = gap_queue[3 * e + 0], hi = gap_queue[3 * e + 1], value = gap_queue[3 * e + 2];
        for (unsigned int c = lo + threadIdx.x; c <= hi; c += blockDim.x) cell_start[c] = value;
    }
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
    int size[MAX_COLS];               /* 4 or 8 bytes */
    unsigned long long off[MAX_COLS]; /* byte offset of the column in the list struct */
};

constexpr int COMPACT_THREADS = 1024;

/* number of flagged (== 1) elements per block */
__global__ void __launch_bounds__(COMPACT_THREADS)
k_flag_blocksum(const int *__restrict__ flags, int n, unsigned int *__restrict__ block_sums, int match) {
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
                  const int *__restrict__ flags, const unsigned int *__restrict__ block_offsets, int dst_offset, int n,
                  int match, int dst_cap, int *overflow) {
    __shared__ unsigned int s_warp[32];
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
    for (int c = 0; c < cols.ncols; ++c) {
        if (cols.size[c] == 4)
            ((unsigned int *)(dst + dcols.off[c]))[out] = ((const unsigned int *)(src + cols.off[c]))[i];
        else
            ((unsigned long long *)(dst + dcols.off[c]))[out] = ((const unsigned long long *)(src + cols.off[c]))[i];
    }
}

/* append_<A>_Agents (krn:265-280) */
__global__ void __launch_bounds__(256)
k_append(const char *__restrict__ src, char *__restrict__ dst, const ColSet cols, int dst_offset, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t out = (size_t)dst_offset + i;
    for (int c = 0; c < cols.ncols; ++c) {
        if (cols.size[c] == 4)
            ((unsigned int *)(dst + cols.off[c]))[out] = ((const unsigned int *)(src + cols.off[c]))[i];
        else
            ((unsigned long long *)(dst + cols.off[c]))[out] = ((const unsigned long long *)(src + cols.off[c]))[i];
    }
}

/* ------------------------------------------------------------------------- */
/* 1-D slab decomposition: halo planes of the sorted message list, migrants    */
/* ------------------------------------------------------------------------- */
/* Exchange buffers have a fixed capacity so that the NCCL message sizes are
 * known on the host without a device round trip; the true counts travel in a
 * 4-int header: {count, overflow, 0, 0}. */
constexpr int XHDR = 4;

/* plane -> send buffer: header | (plane_cells + 1) cell_start entries rebased to 0 | records */
template <int NQ>
__global__ void __launch_bounds__(256)
k_pack_plane(const uint4 *__restrict__ sorted_local, const unsigned int *__restrict__ cell_start, int first_cell,
             int plane_cells, int cap, int *__restrict__ out) {
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

/* Merge: shift the owned planes' cell_start by the position of the local block
 * inside the combined buffer, splice in the two received halo planes.
 *   combined buffer: [ ... low halo | local block (n records at `local_at`) | high halo ... ] */
template <int NQ>
__global__ void __launch_bounds__(256)
k_merge_halo(uint4 *__restrict__ combined, unsigned int *__restrict__ cell_start, int local_at, int n,
             int own_first_cell, int own_cells, const int *__restrict__ lo_buf, int lo_first_cell,
             const int *__restrict__ hi_buf, int hi_first_cell, int plane_cells, int *status) {
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lo_cnt = lo_buf[0], hi_cnt = hi_buf[0];
    if (tid == 0 && (lo_buf[1] || hi_buf[1])) atomicExch(status, 1);
    const int pad = (plane_cells + 1 + 3) & ~3;
    /* owned planes (including the closing boundary) */
    for (int i = tid; i <= own_cells; i += stride) cell_start[own_first_cell + i] += (unsigned int)local_at;
    __threadfence();
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

/* Owner of each agent along the slab axis: 1 = stays, 2 = lower neighbour, 3 = upper neighbour,
 * 4 = further than one slab away (reported).  The plane is the cell coordinate of
 * message_<M>_grid_position + the wrap rule of message_<M>_hash (krn:860-889). */
__global__ void __launch_bounds__(256)
k_owner_flags(const float *__restrict__ coord, int n, float minb, float maxb, int nplanes, int planes_per_rank, int rank,
              int world, int *__restrict__ flags, int *status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = (int)floorf(__fdiv_rn(__fmul_rn(__fsub_rn(coord[i], minb), (float)nplanes), __fsub_rn(maxb, minb)));
    p = (p < 0) ? nplanes - 1 : p;
    p = (p >= nplanes) ? 0 : p;
    const int owner = min(p / planes_per_rank, world - 1);
    int f = 4;
    if (owner == rank) f = 1;
    else if (owner == (rank + world - 1) % world && owner == (rank + 1) % world) f = (p < rank * planes_per_rank) ? 2 : 3; /* world == 2 */
    else if (owner == (rank + world - 1) % world) f = 2;
    else if (owner == (rank + 1) % world) f = 3;
    flags[i] = f;
    if (f == 4) atomicExch(status, 2);
}

/* header of a migrant buffer := {count (clamped), overflow} from the compaction total */
__global__ void k_set_header(int *hdr, const int *total, int cap) {
    const int t = *total;
    hdr[0] = min(t, cap);
    hdr[1] = t > cap;
    hdr[2] = 0;
    hdr[3] = 0;
}

/* append received migrants: dst[dst_base + *off_a + *off_b + i] = src[i] for i < src_hdr[0] */
__global__ void __launch_bounds__(256)
k_append_received(const char *__restrict__ src, const int *__restrict__ src_hdr, char *__restrict__ dst, const ColSet scols,
                  const ColSet dcols, const int *off_a, const int *off_b, int dst_cap, int *status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = src_hdr[0];
    if (i == 0 && src_hdr[1]) atomicExch(status, 3);
    if (i >= cnt) return;
    const size_t out = (size_t)(off_a ? *off_a : 0) + (size_t)(off_b ? *off_b : 0) + i;
    if (out >= (size_t)dst_cap) {
        atomicExch(status, 4);
        return;
    }
    for (int c = 0; c < scols.ncols; ++c) {
        if (scols.size[c] == 4)
            ((unsigned int *)(dst + dcols.off[c]))[out] = ((const unsigned int *)(src + scols.off[c]))[i];
        else
            ((unsigned long long *)(dst + dcols.off[c]))[out] = ((const unsigned long long *)(src + scols.off[c]))[i];
    }
}

/* result[0] = *a + hdr_b[0] + hdr_c[0] (new local population) */
__global__ void k_sum_counts(int *result, const int *a, const int *hdr_b, const int *hdr_c) {
    result[0] = *a + hdr_b[0] + hdr_c[0];
}

}  // namespace fgpu
#endif
