/*
 * fgpu_device.cuh -- device-side model API for sm_100a.
 *
 * The generated per-model header instantiates these templates under the
 * reference's names (add_<M>_message, get_first/next_<M>_message, rnd<>,
 * add_<A>_agent) so that functions.c compiles unchanged.  Behaviour follows
 * FLAMEGPU/templates/FLAMEGPU_kernals.xslt: cell/hash :860-889, neighbour
 * walk :105-154 + :1037-1160, message append :382-404, agent birth :287-314,
 * rand48 :1488-1547.  The mechanism is different by design:
 *
 *  - the sorted message list is an array of 16-byte aligned records (the
 *    user-visible message struct itself): one LDG.128 per 4 variables instead
 *    of one texture fetch per variable (krn:1090-1097), and get_first/next
 *    return a pointer into that array instead of copying every message
 *    through a shared-memory slot (krn:1100-1103);
 *  - the partition boundary matrix is a CSR array cell_start[cells+1]; the
 *    three x-neighbours of a row are one contiguous range, so a 27-cell walk
 *    costs 18 index loads instead of 54 (krn:1061-1062); blocks that process
 *    agents in cell order stage their PBM slices and message slices in shared
 *    memory once and walk them from there;
 *  - the cursor lives in a per-thread proxy that the compiler keeps in
 *    registers;
 *  - the rand48 state is loaded once per agent function, iterated in
 *    registers and stored once (the reference loads/stores per rnd() call,
 *    krn:1531-1540);
 *  - every index is the LOGICAL list index carried by the proxy, never
 *    blockIdx/threadIdx, so the launch shape is free.
 */
#ifndef FGPU_DEVICE_CUH
#define FGPU_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define FGPU_DEV __device__ __forceinline__

namespace fgpu {

/* ---- spatial constants of one message type (compile-time, from the XMML) -- */
/* Traits must provide:
 *   static constexpr int MAX, NQ, DIMX, DIMY, DIMZ; static constexpr bool IS2D;
 *   static constexpr float MINX..MAXZ;                                        */

/* message_<M>_grid_position, krn:860-871: float ops in this exact order.
 * The (float) casts and the single rounding per operation are kept by
 * forbidding contraction on these expressions (no a*b+c pattern is present:
 * sub, mul, div). */
template <class T>
FGPU_DEV void grid_position(float x, float y, float z, int &cx, int &cy, int &cz) {
    cx = (int)floorf(__fdiv_rn(__fmul_rn(__fsub_rn(x, T::MINX), (float)T::DIMX), __fsub_rn(T::MAXX, T::MINX)));
    cy = (int)floorf(__fdiv_rn(__fmul_rn(__fsub_rn(y, T::MINY), (float)T::DIMY), __fsub_rn(T::MAXY, T::MINY)));
    cz = (int)floorf(__fdiv_rn(__fmul_rn(__fsub_rn(z, T::MINZ), (float)T::DIMZ), __fsub_rn(T::MAXZ, T::MINZ)));
}

/* message_<M>_hash, krn:877-889: out-of-range jumps to the opposite face. */
template <class T>
FGPU_DEV unsigned int cell_hash(int cx, int cy, int cz) {
    cx = (cx < 0) ? T::DIMX - 1 : cx;
    cx = (cx >= T::DIMX) ? 0 : cx;
    cy = (cy < 0) ? T::DIMY - 1 : cy;
    cy = (cy >= T::DIMY) ? 0 : cy;
    cz = (cz < 0) ? T::DIMZ - 1 : cz;
    cz = (cz >= T::DIMZ) ? 0 : cz;
    return (unsigned int)(((cz * T::DIMY) * T::DIMX) + (cy * T::DIMX) + cx);
}

#ifdef FGPU_STAGE
/* ---- block-level staging of the neighbourhood (build option, off by default) ---------------- */
/* When a message consumer processes its agents in cell order (fgpu_launch.order),
 * the agents of one block occupy a short run of consecutive cells [c_lo, c_hi].
 * Their 27-cell neighbourhoods are then 9 windows of consecutive cells (one per
 * (dy,dz) row offset), i.e. 9 contiguous slices of the partition boundary
 * matrix and 9 contiguous slices of the sorted message list.  The block copies
 * those slices into shared memory once (coalesced 16-byte loads) and every
 * thread walks its neighbourhood out of shared memory.  Threads whose cell is
 * not fully interior (any wrap), lies outside the block's span, or whose block
 * does not fit the staging capacity use the global walk instead -- same visiting
 * order, same results.  Measured in round 1 (profiles/r01_notes.md): correct but
 * slower than the prefetching global walk (shared-memory footprint -> 20 warps/SM,
 * generic-address loads), hence a build option. */
constexpr int STAGE_CSW = 128;            /* staged PBM row window: span + 3 cell boundaries */
constexpr int STAGE_HDR_QUADS = 8;        /* StageShared fits in 128 bytes */

struct StageShared {
    int ok;         /* 1 = windows staged */
    int c_lo, c_hi; /* cell span of the block's agents (hash space) */
    int adj[9];     /* shared-memory message index = sorted message index + adj[row] */
    int cnt[9];     /* messages in the window of each row (scratch) */
};
static_assert(sizeof(StageShared) <= STAGE_HDR_QUADS * 16, "staging header overflows its slot");

template <class T>
FGPU_DEV int row_offset(int k) {
    /* rows are enumerated z-major, y-minor like next_cell3D; in 2-D mode z does not move */
    if (T::IS2D) return (k - 1) * T::DIMX;
    return ((k % 3) - 1) * T::DIMX + ((k / 3) - 1) * T::DIMX * T::DIMY;
}

template <class Msg, class T>
FGPU_DEV void stage_block(StageShared *sh, int *s_cs, Msg *s_msgs, int stage_msgs, const Msg *recs,
                          const unsigned int *cell_start, const unsigned int *sorted_keys, bool ordered, int n,
                          int count) {
    constexpr int NROWS = T::IS2D ? 3 : 9;
    constexpr int CELLS = T::DIMX * T::DIMY * T::DIMZ;
    constexpr int NQ = (int)(sizeof(Msg) / 16);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (tid == 0) {
        int ok = (stage_msgs > 0) && ordered && (sorted_keys != nullptr) && (count == n) && (n > 0);
        int c_lo = 0, c_hi = -1;
        if (ok) {
            const int t0 = blockIdx.x * nthreads;
            const int t1 = min(t0 + nthreads, n) - 1;
            c_lo = (int)sorted_keys[t0];
            c_hi = (int)sorted_keys[t1];
            const int reach = T::IS2D ? (T::DIMX + 1) : (T::DIMX * T::DIMY + T::DIMX + 1);
            ok = (c_hi - c_lo + 4 <= STAGE_CSW) && (c_lo - reach >= 0) && (c_hi + reach <= CELLS - 1);
        }
        sh->ok = ok;
        sh->c_lo = c_lo;
        sh->c_hi = c_hi;
    }
    __syncthreads();
    if (!sh->ok) return;
    const int c_lo = sh->c_lo;
    const int w = sh->c_hi - c_lo + 4; /* boundaries needed per row: cells c_lo-1 .. c_hi+1, plus one */
#pragma unroll
    for (int k = 0; k < NROWS; ++k) {
        const int first = c_lo + row_offset<T>(k) - 1;
        for (int i = tid; i < w; i += nthreads) s_cs[k * STAGE_CSW + i] = (int)__ldg(cell_start + first + i);
    }
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int k = 0; k < NROWS; ++k) {
            const int wlo = s_cs[k * STAGE_CSW], whi = s_cs[k * STAGE_CSW + w - 1];
            sh->adj[k] = total - wlo;
            sh->cnt[k] = whi - wlo;
            total += whi - wlo;
        }
        if (total > stage_msgs) sh->ok = 0;
    }
    __syncthreads();
    if (!sh->ok) return;
#pragma unroll
    for (int k = 0; k < NROWS; ++k) {
        const int wlo = s_cs[k * STAGE_CSW];
        const uint4 *src = reinterpret_cast<const uint4 *>(recs + wlo);
        uint4 *dst = reinterpret_cast<uint4 *>(s_msgs + (wlo + sh->adj[k]));
        const int quads = sh->cnt[k] * NQ;
        for (int i = tid; i < quads; i += nthreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
}
#endif /* FGPU_STAGE */

/* ---- per-thread message proxy -------------------------------------------- */
/* `Msg` is the user-visible struct xmachine_message_<M> (16-byte aligned, payload
 * only); the sorted message list in HBM is an array of exactly that struct, so
 * get_first/next hand user code a pointer straight into the list (the reference
 * copies every message through a per-thread shared-memory slot, krn:1084-1103).
 * `T` holds the partition constants.  The generated header derives
 * xmachine_message_<M>_list from this.
 *
 * Neighbourhood walk, FGPU_WALK_PIPELINE variant = a 3-deep software pipeline over the (up to) 9 row ranges:
 *   range in use      [cur, end)        being iterated by the script
 *   next range        [n1lo, n1hi)      PBM entries already loaded; its first
 *                                       message lines are prefetched into L1 when
 *                                       it becomes "next"
 *   range after next  [n2lo, n2hi)      PBM loads in flight
 * so a range switch costs no dependent load: the reference resolves each cell
 * lazily inside the message loop (krn:1048-1081), two dependent fetches on a
 * divergent path that stalls the whole warp. */
template <class Msg, class T>
struct MessageProxy {
    /* input side */
    const Msg *in_recs;
    const unsigned int *cell_start;
    int in_count;
    /* output side */
    Msg *out_recs;
    unsigned int *out_keys;
    int out_index;
    /* cursor (replaces _relative_cell/_cell_index/_cell_index_max/_agent_grid_cell, hdr:170-173) */
    int cur, end;     /* current message index, end of the open range */
#ifdef FGPU_WALK_PIPELINE
    int n1lo, n1hi;   /* next range (n1hi < 0: neighbourhood exhausted) */
    int n2lo, n2hi;   /* range after next */
#endif
    int step;         /* next group whose PBM entries are to be fetched: interior agents count rows
                         (0..8 / 0..2), others count relative cells (0..26 / 0..8) */
    int cx, cy, cz;   /* agent grid cell (non-interior walk) */
    int hrow;         /* interior walk: hash of the first cell (x-1) of the next row */
    int interior;     /* no neighbour cell wraps: rows are hrow, hrow+DIMX, ... */
#ifdef FGPU_STAGE
    const Msg *base;  /* list the open range indexes: in_recs or the staged copy */
    int fast;
    const StageShared *sh;
    const int *s_cs;
    const Msg *s_msgs;
#endif
};

FGPU_DEV void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

/* Issues the PBM loads of the next group of cells.  Cells are visited x-fastest
 * from (-1,-1,-1) to (1,1,1) exactly like next_cell3D/next_cell2D (krn:105-154),
 * each through the wrap rule of message_<M>_hash; the three x cells of a row are
 * adjacent hashes unless x wraps, and are then fetched as ONE range (same
 * visiting order).  hi < 0 marks the end of the neighbourhood. */
template <class Msg, class T>
FGPU_DEV void fetch_group(MessageProxy<Msg, T> *p, int &lo, int &hi) {
    const int s = p->step;
    if (p->interior) {
        /* interior agent: the 9 (3) rows are at fixed hash strides, no wrap tests */
        constexpr int NROWS = T::IS2D ? 3 : 9;
        if (s >= NROWS) {
            lo = 0;
            hi = -1;
            return;
        }
        const int h = p->hrow;
        lo = (int)__ldg(p->cell_start + h);
        hi = (int)__ldg(p->cell_start + h + 3);
        p->step = s + 1;
        p->hrow = h + ((s == 2 || s == 5) ? (T::DIMX * T::DIMY - 2 * T::DIMX) : T::DIMX);
        return;
    }
    constexpr int NSTEPS = T::IS2D ? 9 : 27;
    if (s >= NSTEPS) {
        lo = 0;
        hi = -1;
        return;
    }
    const int rx = (s % 3) - 1;
    const int ry = ((s / 3) % 3) - 1;
    const int rz = T::IS2D ? -1 : (s / 9) - 1;   /* 2-D: relative z stays -1 (krn:137-154) */
    if (rx == -1 && p->cx >= 1 && p->cx + 1 < T::DIMX) {
        const unsigned int h = cell_hash<T>(p->cx - 1, p->cy + ry, p->cz + rz);
        lo = (int)__ldg(p->cell_start + h);
        hi = (int)__ldg(p->cell_start + h + 3);
        p->step = s + 3;
    } else {
        const unsigned int h = cell_hash<T>(p->cx + rx, p->cy + ry, p->cz + rz);
        lo = (int)__ldg(p->cell_start + h);
        hi = (int)__ldg(p->cell_start + h + 1);
        p->step = s + 1;
    }
}

/* Opens the next non-empty range of the neighbourhood; false when exhausted. */
template <class Msg, class T>
FGPU_DEV bool open_next_range(MessageProxy<Msg, T> *p) {
#ifdef FGPU_STAGE
    if (p->fast) {
        constexpr int NROWS = T::IS2D ? 3 : 9;
        while (p->step < NROWS) {
            const int k = p->step++;
            const int lo = p->s_cs[k * STAGE_CSW + p->cx];
            const int hi = p->s_cs[k * STAGE_CSW + p->cx + 3];
            if (lo < hi) {
                const int a = p->sh->adj[k];
                p->cur = lo + a;
                p->end = hi + a;
                return true;
            }
        }
        return false;
    }
#endif
#ifdef FGPU_WALK_PIPELINE
    while (true) {
        const int lo = p->n1lo, hi = p->n1hi;
        if (hi < 0) return false;
        p->n1lo = p->n2lo;
        p->n1hi = p->n2hi;
        if (p->n1hi > p->n1lo) { /* the range that becomes "next": pull its first lines towards L1 */
            const char *q = reinterpret_cast<const char *>(p->in_recs + p->n1lo);
            prefetch_l1(q);
            prefetch_l1(q + 128);
        }
        fetch_group<Msg, T>(p, p->n2lo, p->n2hi);
        if (lo < hi) {
            p->cur = lo;
            p->end = hi;
            return true;
        }
    }
#else
    /* The kernel is instruction-issue bound (ncu: 78 % issue-slot utilisation, 64 resident warps),
     * so the cheapest switch wins: fetch the next group on demand and let the other warps hide the
     * L1 latency.  The software-pipelined variant above (FGPU_WALK_PIPELINE) removes the dependent
     * load but costs ~20 more instructions per switch and measured slower (profiles/r01_notes.md). */
    while (true) {
        int lo, hi;
        fetch_group<Msg, T>(p, lo, hi);
        if (hi < 0) return false;
        if (lo < hi) {
            p->cur = lo;
            p->end = hi;
            return true;
        }
    }
#endif
}

/* get_first: locate the agent's cell and start the range pipeline. */
template <class Msg, class T>
FGPU_DEV void begin_walk(MessageProxy<Msg, T> *p, float x, float y, float z) {
    int cx, cy, cz;
    grid_position<T>(x, y, z, cx, cy, cz);
    p->step = 0;
    p->cx = cx;
    p->cy = cy;
    p->cz = cz;
#ifdef FGPU_STAGE
    p->fast = 0;
    p->base = p->in_recs;
    if (p->sh != nullptr && p->sh->ok) {
        const bool interior = (cx >= 1) && (cx + 1 < T::DIMX) && (cy >= 1) && (cy + 1 < T::DIMY) &&
                              (T::IS2D || ((cz >= 1) && (cz + 1 < T::DIMZ)));
        const int c = (int)cell_hash<T>(cx, cy, cz);
        if (interior && c >= p->sh->c_lo && c <= p->sh->c_hi) {
            p->fast = 1;
            p->cx = c - p->sh->c_lo;
            p->base = p->s_msgs;
            return;
        }
    }
#endif
    p->interior = (cx >= 1) && (cx + 1 < T::DIMX) && (cy >= 1) && (cy + 1 < T::DIMY) &&
                  (T::IS2D || ((cz >= 1) && (cz + 1 < T::DIMZ)));
    p->hrow = (int)cell_hash<T>(cx - 1, cy - 1, cz - 1);   /* 2-D: z-1 wraps to plane 0 */
#ifdef FGPU_WALK_PIPELINE
    fetch_group<Msg, T>(p, p->n1lo, p->n1hi);
    fetch_group<Msg, T>(p, p->n2lo, p->n2hi);
#endif
}

/* Hands the current message to user code: a pointer straight into the list. */
template <class Msg, class T>
FGPU_DEV Msg *current_message(MessageProxy<Msg, T> *p) {
#ifdef FGPU_STAGE
    Msg *m = const_cast<Msg *>(p->base + p->cur);
#else
    Msg *m = const_cast<Msg *>(p->in_recs + p->cur);
#endif
    __builtin_assume(m != nullptr);
    return m;
}

/* ---- rand48 ---------------------------------------------------------------- */
/* Per-thread proxy behind the user's RNG_rand48*: constants + this slot's state. */
struct Rand48Proxy {
    unsigned int Ax, Ay, Cx, Cy;
    unsigned int sx, sy;
};

/* RNG_rand48_iterate_single, krn:1488-1514 (24-bit limb arithmetic, exact). */
FGPU_DEV void rand48_iterate(Rand48Proxy *r) {
    unsigned int R0, R1;
    const unsigned int lo00 = __umul24(r->sx, r->Ax);
    const unsigned int hi00 = __umulhi(r->sx, r->Ax);
    R0 = (lo00 & 0xFFFFFF);
    R1 = (lo00 >> 24) | (hi00 << 8);
    R0 += r->Cx;
    R1 += r->Cy;
    R1 += (R0 >> 24);
    R0 &= 0xFFFFFF;
    R1 += __umul24(r->sy, r->Ax);
    R1 += __umul24(r->sx, r->Ay);
    R1 &= 0xFFFFFF;
    r->sx = R0;
    r->sy = R1;
}

/* rnd<>, krn:1516-1543: emit the top 31 bits, then iterate. */
FGPU_DEV float rand48_draw(Rand48Proxy *r) {
    int rand = (r->sx >> 17) | (r->sy << 7);
    rand48_iterate(r);
    return (float)rand / 2147483647;
}

/* ---- newborn agents -------------------------------------------------------- */
/* Proxy behind the user's xmachine_memory_<B>_list* in functions with
 * xagentOutputs: the sparse d_<B>s_new list and the parent's logical index
 * (krn:287-314 writes the newborn at the parent's thread index). */
struct NewAgentProxy {
    void *list;
    int index;
};

FGPU_DEV unsigned int f2u(float f) { return __float_as_uint(f); }
FGPU_DEV float u2f(unsigned int u) { return __uint_as_float(u); }

}  // namespace fgpu

#endif
