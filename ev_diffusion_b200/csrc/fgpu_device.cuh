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
 *    costs 18 index loads instead of 54 (krn:1061-1062);
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

/* ---- per-thread message proxy -------------------------------------------- */
/* `Msg` is the user-visible struct xmachine_message_<M> (16-byte aligned, payload
 * only); the sorted message list in HBM is an array of exactly that struct, so
 * get_first/next hand user code a pointer straight into the list (no staging
 * copy; the reference copies every message through a shared-memory slot,
 * krn:1084-1103).  `T` holds the partition constants.  The generated header
 * derives xmachine_message_<M>_list from this. */
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
    const Msg *end;   /* one past the last message of the open range */
    int step;         /* next relative cell to open: 0..26 (3-D) or 0..8 (2-D) */
    int cx, cy, cz;   /* agent grid cell */
};

/* Opens the next non-empty cell range and returns its first message (nullptr
 * when the neighbourhood is exhausted).  Cells are visited x-fastest from
 * (-1,-1,-1) to (1,1,1) exactly like next_cell3D/next_cell2D (krn:105-154);
 * when the three x cells of a row do not wrap they are adjacent hashes and
 * are opened as one range (same visiting order). */
template <class Msg, class T>
FGPU_DEV const Msg *open_next_range(MessageProxy<Msg, T> *p) {
    constexpr int NSTEPS = T::IS2D ? 9 : 27;
    while (p->step < NSTEPS) {
        const int s = p->step;
        const int rx = (s % 3) - 1;
        const int ry = ((s / 3) % 3) - 1;
        const int rz = T::IS2D ? -1 : (s / 9) - 1;   /* 2-D: relative z stays -1 (krn:137-154) */
        int lo, hi;
        if (rx == -1 && p->cx >= 1 && p->cx + 1 < T::DIMX) {
            /* whole row x-1..x+1 without wrap: one contiguous hash range */
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
        if (lo < hi) {
            p->end = p->in_recs + hi;
            return p->in_recs + lo;
        }
    }
    return nullptr;
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
