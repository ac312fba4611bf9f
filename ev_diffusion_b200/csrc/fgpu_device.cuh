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
__host__ __device__ constexpr bool is_power_of_two(float v) {
    if (!(v > 0.0f) || v > 1e30f || v < 1e-30f) return false;
    while (v > 1.0f) v *= 0.5f;   /* exact for a power of two; any other value never reaches 1 */
    while (v < 1.0f) v *= 2.0f;
    return v == 1.0f;
}

/* a / extent, correctly rounded.  The extent (max - min of the partitioning bounds) is a compile-time constant; when it is a power of
 * two the quotient is an exact scaling, so the product with the (exact) reciprocal is the same correctly rounded value for every a --
 * one FMUL instead of the IEEE division sequence (FCHK + reciprocal + 3 FFMA + a slow-path call) the intrinsic compiles to. */
template <int AXIS, class T>
FGPU_DEV float div_extent(float a) {
    constexpr float ext = AXIS == 0 ? T::MAXX - T::MINX : AXIS == 1 ? T::MAXY - T::MINY : T::MAXZ - T::MINZ;
    if constexpr (is_power_of_two(ext)) return __fmul_rn(a, 1.0f / ext);
    else return __fdiv_rn(a, ext);
}

template <class T>
FGPU_DEV void grid_position(float x, float y, float z, int &cx, int &cy, int &cz) {
    cx = (int)floorf(div_extent<0, T>(__fmul_rn(__fsub_rn(x, T::MINX), (float)T::DIMX)));
    cy = (int)floorf(div_extent<1, T>(__fmul_rn(__fsub_rn(y, T::MINY), (float)T::DIMY)));
    cz = (int)floorf(div_extent<2, T>(__fmul_rn(__fsub_rn(z, T::MINZ), (float)T::DIMZ)));
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


/* ---- per-thread message proxy -------------------------------------------- */
/* `Msg` is the user-visible struct xmachine_message_<M> (16-byte aligned, payload
 * only); the sorted message list in HBM is an array of exactly that struct, so
 * get_first/next hand user code a pointer straight into the list (the reference
 * copies every message through a per-thread shared-memory slot, krn:1084-1103).
 * `T` holds the partition constants.  The generated header derives
 * xmachine_message_<M>_list from this.
 *
 *
 * Neighbourhood walk.  The 27 (9) cells are visited x-fastest from (-1,-1,-1) exactly like
 * next_cell3D/next_cell2D (krn:105-154).  The three x cells of a row are adjacent hashes unless x
 * wraps, so a row is ONE contiguous range of the sorted list: the walk has two levels, 9 (3) rows x
 * the messages of a row.  The PBM entries of row k+1 are requested when row k is opened, so they are in
 * flight while row k is iterated.  The reference resolves each cell lazily inside the flat message loop
 * (krn:1048-1081): two dependent fetches plus the hash arithmetic on a path only a few lanes of a warp
 * take at a time.  Scripts whose message loops have the canonical shape are restructured at generation
 * time (looprewrite.py) so that the two levels are two nested loops: the inner one is a counted loop over
 * contiguous 16-byte records, and the whole warp switches rows together.  Threads in the first or last
 * x column (x wraps) see a row as two ranges and stay in the same loop nest; only an agent outside the
 * grid in x walks cell by cell. */

template <class Msg, class T>
struct MessageProxy {
    /* input side */
    const Msg *in_recs;
    const unsigned int *cell_start;
    int in_count;
    /* output side */
    Msg *out_recs;
    unsigned int *out_keys;
    int *out_flags;   /* optional_message output: one flag per slot, or nullptr */
    int out_index;
    /* cursor (replaces _relative_cell/_cell_index/_cell_index_max/_agent_grid_cell, hdr:170-173) */
    const Msg *end;   /* one past the last message of the open range; the cursor itself is the pointer
                         user code holds (get_next receives it back) */
    int cnt;          /* number of messages in the open range (restructured loops count instead of comparing pointers) */
    int step;         /* ROW: next row (0..9 / 0..3); EDGE: next half-row (0..18 / 0..6); CELL: next relative cell (0..27 / 0..9) */
    int mode;         /* WALK_ROW / WALK_EDGE / WALK_CELL */
    unsigned int nlo, nhi; /* ROW: PBM entries of the next row, loaded one row ahead */
    int cx, cy, cz;   /* agent grid cell */
};

enum { WALK_ROW = 0, WALK_EDGE = 1, WALK_CELL = 2 };

/* hash of cell x = 0 of row k = (ry, rz) of the neighbourhood of cell (.,cy,cz); y and z wrap like message_<M>_hash */
template <class T>
FGPU_DEV unsigned int row_base(int cy, int cz, int k) {
    const int kz = T::IS2D ? 0 : (k * 11) >> 5;          /* k / 3 for 0 <= k < 9 */
    const int ry = k - 3 * kz - 1;                       /* k % 3 - 1 */
    int y = cy + ry;
    int z = cz + (T::IS2D ? -1 : kz - 1);                /* 2-D: relative z stays -1 (krn:137-154) */
    y = (y < 0) ? T::DIMY - 1 : y;
    y = (y >= T::DIMY) ? 0 : y;
    z = (z < 0) ? T::DIMZ - 1 : z;
    z = (z >= T::DIMZ) ? 0 : z;
    return (unsigned int)((z * T::DIMY + y) * T::DIMX);
}

/* Opens the next non-empty range of the neighbourhood and returns its first message; nullptr when
 * the neighbourhood is exhausted.
 *   ROW   1 <= cx <= DIMX-2: the three x cells of a row are adjacent hashes, one range per row, the PBM entries
 *         of row k+1 requested when row k is opened;
 *   EDGE  cx == 0 or cx == DIMX-1: x wraps to the opposite face (krn:880-885), a row is two ranges visited in
 *         the reference's order -- (DIMX-1) then (0,1), or (DIMX-2,DIMX-1) then (0) -- so that these lanes stay
 *         inside the same two-level loop as the rest of their warp;
 *   CELL  an agent outside the grid in x (cells repeat under the wrap rule): cell by cell, like the reference. */
template <class Msg, class T>
FGPU_DEV Msg *open_next_range(MessageProxy<Msg, T> *p) {
    if (p->mode == WALK_ROW) {
        constexpr int NROWS = T::IS2D ? 3 : 9;
#pragma unroll 1
        while (p->step < NROWS) {
            const unsigned int lo = p->nlo, hi = p->nhi;
            const int k = ++p->step;
            if (k < NROWS) { /* PBM entries of the row after this one: in flight while this row is iterated */
                const unsigned int *cs = p->cell_start + (row_base<T>(p->cy, p->cz, k) + (unsigned int)(p->cx - 1));
                p->nlo = __ldg(cs);
                p->nhi = __ldg(cs + 3);
            }
            if (lo < hi) {
                p->cnt = (int)(hi - lo);
                p->end = p->in_recs + hi;
                return const_cast<Msg *>(p->in_recs + lo);
            }
        }
        return nullptr;
    }
    if (p->mode == WALK_EDGE) {
        constexpr int NHALF = T::IS2D ? 6 : 18;
#pragma unroll 1
        while (p->step < NHALF) {
            const int s = p->step++;
            const unsigned int *cs = p->cell_start + row_base<T>(p->cy, p->cz, s >> 1);
            /* cx == 0:       first (DIMX-1), then (0, 1);   cx == DIMX-1: first (DIMX-2, DIMX-1), then (0) */
            const bool first = (s & 1) == 0, left = p->cx == 0;
            const int c0 = first ? (left ? T::DIMX - 1 : T::DIMX - 2) : 0;
            const int nc = (first == left) ? 1 : 2;
            const unsigned int lo = __ldg(cs + c0), hi = __ldg(cs + c0 + nc);
            if (lo < hi) {
                p->cnt = (int)(hi - lo);
                p->end = p->in_recs + hi;
                return const_cast<Msg *>(p->in_recs + lo);
            }
        }
        return nullptr;
    }
    /* cell walk, one cell at a time through the wrap rule of message_<M>_hash */
    constexpr int NSTEPS = T::IS2D ? 9 : 27;
#pragma unroll 1
    while (p->step < NSTEPS) {
        const int s = p->step++;
        const int rx = (s % 3) - 1;
        const int ry = ((s / 3) % 3) - 1;
        const int rz = T::IS2D ? -1 : (s / 9) - 1;   /* 2-D: relative z stays -1 (krn:137-154) */
        const unsigned int h = cell_hash<T>(p->cx + rx, p->cy + ry, p->cz + rz);
        const int lo = (int)__ldg(p->cell_start + h);
        const int hi = (int)__ldg(p->cell_start + h + 1);
        if (lo < hi) {
            p->cnt = hi - lo;
            p->end = p->in_recs + hi;
            return const_cast<Msg *>(p->in_recs + lo);
        }
    }
    return nullptr;
}

/* get_next: the fast path is one pointer bump and one compare. */
template <class Msg, class T>
FGPU_DEV Msg *next_message(Msg *current, MessageProxy<Msg, T> *p) {
    Msg *next = current + 1;
    if (next != p->end) {
        __builtin_assume(next != nullptr);
        return next;
    }
    return open_next_range<Msg, T>(p);
}

/* get_first: locate the agent's cell and resolve its rows. */
template <class Msg, class T>
FGPU_DEV void begin_walk(MessageProxy<Msg, T> *p, float x, float y, float z) {
    int cx, cy, cz;
    grid_position<T>(x, y, z, cx, cy, cz);
    p->step = 0;
    p->cx = cx;
    p->cy = cy;
    p->cz = cz;
    p->mode = WALK_CELL;
    if (cx >= 1 && cx + 1 < T::DIMX) {
        const unsigned int *cs = p->cell_start + (row_base<T>(cy, cz, 0) + (unsigned int)(cx - 1));
        p->nlo = __ldg(cs);
        p->nhi = __ldg(cs + 3);
        p->mode = WALK_ROW;
    } else if (T::DIMX >= 3 && (cx == 0 || cx == T::DIMX - 1)) {
        p->mode = WALK_EDGE;
    }
}

/* One message record into registers with 16-byte loads through the read-only path (the sorted list is
 * never written by a kernel that reads it). */
template <class Msg>
FGPU_DEV Msg load_message(const Msg *m) {
    constexpr int NQ = (int)(sizeof(Msg) / 16);
    union {
        Msg msg;
        uint4 q[NQ];
    } r;
#pragma unroll
    for (int i = 0; i < NQ; ++i) r.q[i] = __ldg(reinterpret_cast<const uint4 *>(m) + i);
    return r.msg;
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
