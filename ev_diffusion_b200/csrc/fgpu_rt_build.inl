/* fgpu_rt_build.inl -- building blocks: stream compaction, radix sort driver, spatial message build.
 * Part of the single translation unit fgpu_rt.cu (textually included there, in this order, after the
 * runtime objects it uses); not a standalone source file. */
/* ------------------------------------------------------------------------- */
/* building blocks                                                            */
/* ------------------------------------------------------------------------- */
#define LAUNCH_CHECK()                                                                       \
    do {                                                                                     \
        s->launches++;                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                 \
        if (e_ != cudaSuccess)                                                               \
            return set_err(FGPU_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

/* Stable compaction of the flagged agents of `src` into `dst` at dst_offset.
 * With count_out != nullptr the survivor count is read back (blocking, like
 * the two 4-byte D2H copies after every cub scan, e.g. simulation.xslt:1476). */
static int compact(fgpu_sim *s, AgentRT &A, const char *src, char *dst, int dst_offset, int n, const int *flags,
                   bool scatter, int *count_out, int match = 1, const ColSet *dcols = nullptr, int dst_cap = 0x7fffffff,
                   int *overflow = nullptr, const int *d_n = nullptr, int *d_total = nullptr) {
    /* d_total: where the survivor count goes on the device (default A.d_total); a caller that passes its own slot and
     * no count_out reads several counts back with one synchronisation */
    int *total = d_total ? d_total : A.d_total;
    if (n <= 0) {
        if (count_out) *count_out = 0;
        CU(cudaMemsetAsync(total, 0, sizeof(int), s->stream));
        return FGPU_OK;
    }
    const int nb = (n + COMPACT_THREADS - 1) / COMPACT_THREADS;
    k_flag_blocksum<<<nb, COMPACT_THREADS, 0, s->stream>>>(flags, n, d_n, A.block_sums, match);
    LAUNCH_CHECK();
    k_scan_blocksums<<<1, 1024, 0, s->stream>>>(A.block_sums, nb, total);
    LAUNCH_CHECK();
    if (count_out) {
        CU(cudaMemcpyAsync(count_out, total, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    if (scatter) {
        k_compact_scatter<<<nb, COMPACT_THREADS, 0, s->stream>>>(src, dst, A.cols, dcols ? *dcols : A.cols, flags, A.block_sums,
                                                                 dst_offset, n, d_n, match, dst_cap, overflow);
        LAUNCH_CHECK();
        s->touch(dst);
    }
    return FGPU_OK;
}

static int choose_digit_bits(int radix_bits, int max_bits, int *passes_out) {
    const int passes = std::max(1, sort_passes(radix_bits, max_bits));
    *passes_out = passes;
    return (radix_bits + passes - 1) / passes; /* balanced digits */
}

/* scratch layout (32-bit words): gap_count | pad[7] | totals[2048] | tile_hist[nbins][tiles] (SCR_* in fgpu_rt.cu) */

static size_t sort_scratch_words(int n, int key_bits, int max_bits) {
    int passes = 1;
    const int digit_bits = choose_digit_bits(key_bits, max_bits, &passes);
    const size_t tiles = ((size_t)n + SORT_TILE - 1) / SORT_TILE;
    return (size_t)SCR_TILEHIST + tiles * ((size_t)1 << digit_bits);
}

template <int BITS>
static cudaError_t launch_scatter(int tiles, cudaStream_t st, const unsigned int *kin, const unsigned int *vin, unsigned int *kout,
                                  unsigned int *vout, int n, const int *d_n, int shift, int digit_bits, const unsigned int *tile_hist,
                                  const unsigned int *totals, unsigned int key_base, const uint4 *pay_in = nullptr, uint4 *pay_out = nullptr,
                                  int pay_quads = 0, bool pdl = false) {
    /* the opt-in above 48 KB is per device: one flag per device of the process */
    static bool configured[64] = {};
    const size_t smem = sizeof(SortSmem<BITS>);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_radix_scatter<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    return launch_k(pdl, k_radix_scatter<BITS>, dim3(tiles), dim3(SORT_THREADS), smem, st, kin, vin, kout, vout, n, d_n, shift, digit_bits, tile_hist,
                    totals, tiles, key_base, pay_in, pay_out, pay_quads);
}

template <int BITS>
static cudaError_t launch_bucket(int nbuckets, cudaStream_t st, const unsigned int *ka, const unsigned int *va, const uint4 *pa, int pq,
                                 unsigned int *order_out, uint4 *sorted_out, const unsigned int *totals, int low_bits, unsigned int key_base,
                                 unsigned int *cell_start, unsigned int first_cell, unsigned int cells_end, unsigned int value_offset) {
    static bool configured[64] = {};
    const size_t smem = sizeof(BucketSmem<BITS>);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_bucket_finish<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    k_bucket_finish<BITS><<<nbuckets, SORT_THREADS, smem, st>>>(ka, va, pa, pq, order_out, sorted_out, totals, nbuckets, low_bits, key_base,
                                                                cell_start, first_cell, cells_end, value_offset);
    return cudaGetLastError();
}

template <class F8, class F9, class F10, class F11>
static cudaError_t by_bits(int bits, F8 f8, F9 f9, F10 f10, F11 f11) {
    if (bits <= 8) return f8();
    if (bits == 9) return f9();
    if (bits == 10) return f10();
    return f11();
}

/* Split of a key of key_bits into a high digit (one stable MSD pass over the whole list) and a low digit (finished
 * bucket by bucket): balanced, so that the pass writes runs of several records per digit and there are enough
 * buckets to fill the machine.  Keys of fewer than 12 or more than 22 bits take the LSD build. */
static bool msd_split(int key_bits, int *high_bits, int *low_bits) {
    if (key_bits < 12 || key_bits > 2 * SORT_MAXBITS) return false;
    *low_bits = key_bits / 2;
    *high_bits = key_bits - *low_bits;
    return true;
}

/* MSD build of a spatially partitioned message list (4 launches whatever the grid size):
 *   k_tile_hist + k_row_scan + k_radix_scatter on the HIGH digit, the message records travelling as payload
 *   k_bucket_finish: per bucket of 2^low cells -- low-digit order, records, message order and the bucket's slice
 *   of the partition boundary matrix. */
static int build_spatial_msd(fgpu_sim *s, MsgRT &M, int n, int key_bits, unsigned int key_base, unsigned int key_span, int *range_error) {
    int H = 0, L = 0;
    msd_split(key_bits, &H, &L);
    const int tiles = (n + SORT_TILE - 1) / SORT_TILE;
    const int nbins = 1 << H;
    const unsigned int span = key_span ? key_span : (unsigned int)M.d->grid_size;
    const int nbuckets = (int)((span + (1u << L) - 1u) >> L);
    unsigned int *totals = M.scratch + SCR_TOTALS, *tile_hist = M.scratch + SCR_TILEHIST;
    const int Q = M.d->rec_quads;
    cudaStream_t st = s->stream;
    s->prof_begin(std::string("sort:") + M.d->name);
    k_tile_hist<<<tiles, SORT_THREADS, sizeof(unsigned int) * (size_t)nbins, st>>>(M.keys_in, n, M.d_count, L, H, tile_hist, tiles, key_base, key_span,
                                                                                   range_error, nullptr);
    LAUNCH_CHECK();
    k_row_scan<<<nbins, SORT_THREADS, 0, st>>>(tile_hist, tiles, totals);
    LAUNCH_CHECK();
    unsigned int *k0 = M.keys[0], *v0 = M.vals[0];
    const unsigned int *kin = M.keys_in;
    const int *dn = M.d_count;
    const uint4 *pin = M.unsorted;
    uint4 *pa = M.pay_a;
    cudaError_t e = by_bits(
        H, [&] { return launch_scatter<8>(tiles, st, kin, nullptr, k0, v0, n, dn, L, H, tile_hist, totals, key_base, pin, pa, Q); },
        [&] { return launch_scatter<9>(tiles, st, kin, nullptr, k0, v0, n, dn, L, H, tile_hist, totals, key_base, pin, pa, Q); },
        [&] { return launch_scatter<10>(tiles, st, kin, nullptr, k0, v0, n, dn, L, H, tile_hist, totals, key_base, pin, pa, Q); },
        [&] { return launch_scatter<11>(tiles, st, kin, nullptr, k0, v0, n, dn, L, H, tile_hist, totals, key_base, pin, pa, Q); });
    if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "MSD pass launch -> %s", cudaGetErrorString(e));
    s->launches++;
    s->prof_end();
    s->prof_begin(std::string("pbm:") + M.d->name);
    unsigned int *order = M.vals[1];
    uint4 *sorted = M.sorted + (size_t)M.local_at * Q;
    unsigned int *cs = M.cell_start;
    const unsigned int cells_end = key_base + span, off = (unsigned int)M.local_at;
    e = by_bits(
        L, [&] { return launch_bucket<8>(nbuckets, st, k0, v0, pa, Q, order, sorted, totals, L, key_base, cs, key_base, cells_end, off); },
        [&] { return launch_bucket<9>(nbuckets, st, k0, v0, pa, Q, order, sorted, totals, L, key_base, cs, key_base, cells_end, off); },
        [&] { return launch_bucket<10>(nbuckets, st, k0, v0, pa, Q, order, sorted, totals, L, key_base, cs, key_base, cells_end, off); },
        [&] { return launch_bucket<11>(nbuckets, st, k0, v0, pa, Q, order, sorted, totals, L, key_base, cs, key_base, cells_end, off); });
    if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "bucket finish launch -> %s", cudaGetErrorString(e));
    s->launches++;
    s->prof_end();
    M.sorted_in = 1;
    M.carry_valid = false;
    return FGPU_OK;
}


/* The hand-written LSD radix sort of (keys_in[i], i): replaces cub::DeviceRadixSort::SortPairs (sim:1867).
 * Returns the index (0/1) of the ping-pong pair holding the sorted result; counts launches. */
static int radix_sort_pairs(cudaStream_t st, const unsigned int *keys_in, int n, int key_bits, int max_bits,
                            unsigned int *keys[2], unsigned int *vals[2], unsigned int *scratch, int *sorted_in,
                            long long *launches, const int *d_n = nullptr, unsigned int key_base = 0, unsigned int key_span = 0,
                            int *range_error = nullptr, bool pdl = false, unsigned int *zero2 = nullptr) {
    int passes = 1;
    const int digit_bits = choose_digit_bits(key_bits, max_bits, &passes);
    if (passes > SORT_MAX_PASSES) return set_err(FGPU_ERR_MODEL, "%d key bits need %d radix passes", key_bits, passes);
    const int tiles = (n + SORT_TILE - 1) / SORT_TILE;
    const int nbins = 1 << digit_bits;
    unsigned int *totals = scratch + SCR_TOTALS, *tile_hist = scratch + SCR_TILEHIST;
    for (int p = 0; p < passes; ++p) {
        const unsigned int *kin = p == 0 ? keys_in : keys[(p - 1) & 1];
        const unsigned int *vin = p == 0 ? nullptr : vals[(p - 1) & 1];
        const int shift = p * digit_bits;
        cudaError_t e = launch_k(pdl, k_tile_hist, dim3(tiles), dim3(SORT_THREADS), sizeof(unsigned int) * (size_t)nbins, st, kin, n, d_n, shift, digit_bits,
                                 tile_hist, tiles, key_base, key_span, p == 0 ? range_error : (int *)nullptr, p == 0 ? zero2 : (unsigned int *)nullptr);
        if (e == cudaSuccess) e = launch_k(pdl, k_row_scan, dim3(nbins), dim3(SORT_THREADS), 0, st, tile_hist, tiles, totals);
        if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "radix pass launch -> %s", cudaGetErrorString(e));
        if (digit_bits <= 8) e = launch_scatter<8>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base, nullptr, nullptr, 0, pdl);
        else if (digit_bits == 9) e = launch_scatter<9>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base, nullptr, nullptr, 0, pdl);
        else if (digit_bits == 10) e = launch_scatter<10>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base, nullptr, nullptr, 0, pdl);
        else e = launch_scatter<11>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base, nullptr, nullptr, 0, pdl);
        if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "radix pass launch -> %s", cudaGetErrorString(e));
        (*launches) += 3;
    }
    *sorted_in = (passes - 1) & 1;
    return FGPU_OK;
}

/* Spatial message build for h_count messages (simulation.xslt:1831-1892). */
static int build_spatial(fgpu_sim *s, MsgRT &M) {
    const int n = M.h_count;
    M.built = true;
    if (n <= 0) return FGPU_OK;
    const int max_bits = std::min(std::max(s->opt_radix_bits, 4), SORT_MAXBITS);
    unsigned int *gap_count = M.scratch + SCR_GAP;
    int in = 0;
    int key_bits = M.d->radix_bits;
    unsigned int key_base = 0, key_span = 0;
    int *range_error = nullptr;
    if (s->slab.on) {
        /* every local message lies in an owned plane (agents migrate before they output): sort on the key
         * relative to the first owned cell, which needs log2(owned cells) bits whatever the world size */
        const SlabRT &S = s->slab;
        key_base = (unsigned int)S.p0 * (unsigned int)S.plane_cells;
        key_span = (unsigned int)S.ppr * (unsigned int)S.plane_cells;
        key_bits = 1;
        while (key_bits < 32 && (1u << key_bits) < key_span) ++key_bits;
        key_bits = std::min(key_bits, M.d->radix_bits);
        range_error = S.d_status + 4;
    }
    {
        int H = 0, L = 0;
        const bool carry = s->opt_carry && M.carry_pending && M.producers == 1; /* carried records take the LSD build */
        if (s->opt_msd && !carry && M.pay_a && msd_split(key_bits, &H, &L) &&
            sizeof(unsigned int) * ((size_t)SCR_TILEHIST + (size_t)((n + SORT_TILE - 1) / SORT_TILE) * ((size_t)1 << H)) <= M.scratch_bytes) {
            return build_spatial_msd(s, M, n, key_bits, key_base, key_span, range_error);
        }
    }
    s->prof_begin(std::string("sort:") + M.d->name);
    if (sizeof(unsigned int) * sort_scratch_words(n, key_bits, max_bits) > M.scratch_bytes)
        return set_err(FGPU_ERR_MODEL, "message %s: sort scratch too small for %d keys of %d bits", M.d->name, n, key_bits);
    const bool pdl = s->opt_pdl != 0;
    int rc = radix_sort_pairs(s->stream, M.keys_in, n, key_bits, max_bits, M.keys, M.vals, M.scratch, &in, &s->launches, M.d_count, key_base,
                              key_span, range_error, pdl, gap_count /* zeroed by the first kernel of the sort */);
    if (rc) return rc;
    s->prof_end();
    M.sorted_in = in;
    s->prof_begin(std::string("pbm:") + M.d->name);
    const int g = (n + 255) / 256;
    const bool carry = M.carry_pending && M.producers == 1;
    M.carry_valid = carry;
    switch (M.d->rec_quads) {
#define GATHER_CASE(Q)                                                                                                   \
    case Q:                                                                                                              \
        launch_k(pdl, k_gather_pbm<Q>, dim3(g), dim3(256), 0, s->stream, M.keys[in], M.vals[in], M.unsorted, M.sorted + (size_t)M.local_at * Q, n, \
                 M.d_count, M.cell_start, key_base, key_base + (key_span ? key_span : (unsigned int)M.d->grid_size), (unsigned int)M.local_at,    \
                 M.gap_queue, gap_count, carry ? M.carry_unsorted : (uint4 *)nullptr, M.carry_sorted, carry ? M.carry_quads : 0);               \
        break;
        GATHER_CASE(1)
        GATHER_CASE(2)
        GATHER_CASE(3)
        GATHER_CASE(4)
        GATHER_CASE(5)
        GATHER_CASE(6)
        GATHER_CASE(7)
        GATHER_CASE(8)
#undef GATHER_CASE
        default:
            return set_err(FGPU_ERR_MODEL, "message %s: record of %d quads is not supported", M.d->name, M.d->rec_quads);
    }
    LAUNCH_CHECK();
    launch_k(pdl, k_fill_gaps, dim3(148 * 2), dim3(256), 0, s->stream, M.cell_start, M.gap_queue, gap_count);
    LAUNCH_CHECK();
    s->prof_end();
    return FGPU_OK;
}
