/* fgpu_rt_slab.inl -- 1-D slab decomposition: NCCL binding, peer-memory transport, halo exchange, migration.
 * Part of the single translation unit fgpu_rt.cu (textually included there, in this order, after the
 * runtime objects it uses); not a standalone source file. */
/* ------------------------------------------------------------------------- */
/* 1-D slab decomposition over the GPUs of one node            */
/* ------------------------------------------------------------------------- */
/* NCCL is bound at run time (dlopen) so that the runtime has no link-time
 * dependency on it and shares the copy torch already loaded, if any. */
namespace {
struct NcclUid { char internal[128]; };
typedef int (*nccl_get_uid_fn)(NcclUid *);
typedef int (*nccl_init_rank_fn)(void **, int, NcclUid, int);
typedef int (*nccl_sendrecv_fn)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_recv_fn)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_void_fn)(void);
typedef int (*nccl_destroy_fn)(void *);
typedef const char *(*nccl_errstr_fn)(int);
struct NcclApi {
    void *lib = nullptr;
    nccl_get_uid_fn GetUniqueId = nullptr;
    nccl_init_rank_fn CommInitRank = nullptr;
    nccl_sendrecv_fn Send = nullptr;
    nccl_recv_fn Recv = nullptr;
    nccl_void_fn GroupStart = nullptr, GroupEnd = nullptr;
    nccl_destroy_fn CommDestroy = nullptr;
    nccl_errstr_fn GetErrorString = nullptr;
} g_nccl;
const int NCCL_INT8 = 0; /* ncclInt8 / ncclChar */

int load_nccl() {
    if (g_nccl.lib) return FGPU_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return set_err(FGPU_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (nccl_get_uid_fn)dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (nccl_init_rank_fn)dlsym(h, "ncclCommInitRank");
    g_nccl.Send = (nccl_sendrecv_fn)dlsym(h, "ncclSend");
    g_nccl.Recv = (nccl_recv_fn)dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (nccl_void_fn)dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (nccl_void_fn)dlsym(h, "ncclGroupEnd");
    g_nccl.CommDestroy = (nccl_destroy_fn)dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd)
        return set_err(FGPU_ERR_COMM, "libnccl.so.2 lacks a required symbol");
    g_nccl.lib = h;
    return FGPU_OK;
}
}  // namespace

#define NC(call)                                                                                          \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0)                                                                                      \
            return set_err(FGPU_ERR_COMM, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                           g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error");            \
    } while (0)

static void slab_release(fgpu_sim *s) {
    SlabRT &S = s->slab;
    if (S.peer_lo) cudaIpcCloseMemHandle(S.peer_lo);
    if (S.peer_hi && S.peer_hi != S.peer_lo) cudaIpcCloseMemHandle(S.peer_hi);
    S.peer_lo = S.peer_hi = nullptr;
    if (S.xbuf) cudaFree(S.xbuf);
    if (S.d_done) cudaFree(S.d_done);
    if (S.d_seq) cudaFree(S.d_seq);
    S.d_seq = nullptr;
    S.xbuf = nullptr;
    S.d_done = nullptr;
    if (S.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(S.comm);
    S.comm = nullptr;
}

extern "C" int fgpu_nccl_unique_id(void *out128) {
    if (!out128) return set_err(FGPU_ERR_ARG, "null pointer");
    int rc = load_nccl();
    if (rc) return rc;
    NC(g_nccl.GetUniqueId((NcclUid *)out128));
    return FGPU_OK;
}

/* rand48 seeds of a slab: slot i of rank r continues the reference stream at position r*N + i, and the
 * stride constants leap over world*N slots, so that the ranks' streams never overlap. */
static int seed_rand48(fgpu_sim *s, unsigned long long skip, unsigned long long stride_slots) {
    static const unsigned long long a = 0x5DEECE66DULL, c = 0xB;
    const unsigned int N = (unsigned int)s->model->buffer_size_max;
    unsigned long long A = 1ULL, C = 0ULL;
    for (unsigned long long i = 0; i < stride_slots; ++i) {
        C += A * c;
        A *= a;
    }
    s->Ax = (unsigned int)(A & 0xFFFFFFULL);
    s->Ay = (unsigned int)((A >> 24) & 0xFFFFFFULL);
    s->Cx = (unsigned int)(C & 0xFFFFFFULL);
    s->Cy = (unsigned int)((C >> 24) & 0xFFFFFFULL);
    std::vector<uint2> h(N);
    unsigned long long x = (((unsigned long long)123) << 16) | 0x330E;
    for (unsigned long long i = 0; i < skip; ++i) x = a * x + c;
    for (unsigned int i = 0; i < N; ++i) {
        x = a * x + c;
        h[i].x = (unsigned int)(x & 0xFFFFFFULL);
        h[i].y = (unsigned int)((x >> 24) & 0xFFFFFFULL);
    }
    CU(cudaMemcpyAsync(s->seeds, h.data(), sizeof(uint2) * (size_t)N, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_slab_init(fgpu_sim *s, int rank, int world, const void *nccl_unique_id, int halo_cap, int mig_cap) {
    if (!s || world < 2 || rank < 0 || rank >= world || !nccl_unique_id) return set_err(FGPU_ERR_ARG, "bad slab arguments");
    const fgpu_model *m = s->model;
    CU(cudaSetDevice(s->device));
    int rc = load_nccl();
    if (rc) return rc;
    SlabRT &S = s->slab;
    /* every spatially partitioned message must live on the same grid */
    const fgpu_message *g = nullptr;
    for (int i = 0; i < m->nmessages; ++i) {
        const fgpu_message *M = &m->messages[i];
        if (M->partitioning != FGPU_PART_SPATIAL) continue;
        if (!g) g = M;
        else if (memcmp(g->dims, M->dims, sizeof(g->dims)) || memcmp(g->min_bounds, M->min_bounds, sizeof(g->min_bounds)) ||
                 memcmp(g->max_bounds, M->max_bounds, sizeof(g->max_bounds)))
            return set_err(FGPU_ERR_MODEL, "slab mode needs all spatial messages on one partition grid (%s differs from %s)", M->name, g->name);
    }
    if (!g) return set_err(FGPU_ERR_MODEL, "slab mode needs a spatially partitioned message");
    S.axis = g->is_2d ? 1 : 2;
    S.nplanes = g->dims[S.axis];
    S.plane_cells = (S.axis == 2) ? g->dims[0] * g->dims[1] : g->dims[0];
    if (S.nplanes % world != 0 || S.nplanes / world < 2)
        return set_err(FGPU_ERR_MODEL, "%d cell planes cannot be cut into %d slabs of >= 2 whole planes", S.nplanes, world);
    if (world == 2 && S.nplanes == 4) /* the two halo planes would be adjacent in hash order and share a boundary entry */
        return set_err(FGPU_ERR_MODEL, "2 slabs need more than 4 cell planes (the two halo planes of a rank must not be adjacent)");
    S.rank = rank;
    S.world = world;
    S.ppr = S.nplanes / world;
    S.p0 = rank * S.ppr;
    S.p1 = S.p0 + S.ppr;
    S.lo_rank = (rank + world - 1) % world;
    S.hi_rank = (rank + 1) % world;
    S.axis_min = g->min_bounds[S.axis];
    S.axis_max = g->max_bounds[S.axis];
    int maxbuf = 0;
    for (int a = 0; a < m->nagents; ++a) maxbuf = std::max(maxbuf, m->agents[a].buffer_size);
    S.halo_cap = halo_cap > 0 ? halo_cap : std::max(4096, (int)std::min<long long>(maxbuf, 4LL * maxbuf / S.ppr));
    S.mig_cap = mig_cap > 0 ? mig_cap : S.halo_cap;
    NcclUid uid;
    memcpy(&uid, nccl_unique_id, sizeof(uid));
    NC(g_nccl.CommInitRank(&S.comm, world, uid, rank));
    CU(cudaMalloc(&S.d_status, sizeof(int) * 8));
    CU(cudaMemset(S.d_status, 0, sizeof(int) * 8));
    CU(cudaMallocHost(&S.h_pinned, sizeof(int) * 256));
    /* messages: the sorted buffer gains room for one halo plane on each side of the local block */
    S.msgs.resize(m->nmessages);
    const int pad = (S.plane_cells + 1 + 3) & ~3;
    for (int i = 0; i < m->nmessages; ++i) {
        MsgRT &M = s->msgs[i];
        if (M.d->partitioning != FGPU_PART_SPATIAL) continue;
        cudaFree(M.sorted);
        M.sorted = nullptr;
        CU(cudaMalloc(&M.sorted, (size_t)M.d->rec_quads * 16 * ((size_t)M.d->buffer_size + 2 * (size_t)S.halo_cap)));
        M.local_at = S.halo_cap;
        SlabMsg &X = S.msgs[i];
        X.bytes = sizeof(int) * (size_t)(XHDR + pad) + (size_t)S.halo_cap * M.d->rec_quads * 16;
        CU(cudaMalloc(&X.send_lo, X.bytes));
        CU(cudaMalloc(&X.send_hi, X.bytes));
        CU(cudaMalloc(&X.recv_lo, X.bytes));
        CU(cudaMalloc(&X.recv_hi, X.bytes));
        CU(cudaMemset(X.recv_lo, 0, X.bytes));
        CU(cudaMemset(X.recv_hi, 0, X.bytes));
    }
    /* agents: fixed-capacity migrant buffers, one column block per variable */
    S.agents.resize(m->nagents);
    const char *axis_name = (S.axis == 2) ? "z" : "y";
    for (int a = 0; a < m->nagents; ++a) {
        AgentRT &A = s->agents[a];
        SlabAgent &X = S.agents[a];
        X.axis_var = -1;
        for (int v = 0; v < A.d->nvars; ++v)
            if (!strcmp(A.d->vars[v].name, axis_name) && A.d->vars[v].is_float && A.d->vars[v].size == 4) X.axis_var = v;
        if (X.axis_var < 0) continue; /* agents without a position never migrate */
        for (int v = 0; v < A.d->nvars; ++v)   /* the migrant buffers and their kernels move one value per variable */
            if (var_len(A.d->vars[v]) > 1)
                return set_err(FGPU_ERR_MODEL, "slab mode: agent %s has the array variable %s[%d]; migrating agents with array variables is not supported",
                               A.d->name, A.d->vars[v].name, var_len(A.d->vars[v]));
        size_t off = 16;
        X.bufcols.ncols = A.d->nvars;
        X.bufcols.stride = (unsigned int)S.mig_cap;
        for (int v = 0; v < A.d->nvars; ++v) {
            off = (off + 15) & ~(size_t)15;
            X.bufcols.size[v] = A.d->vars[v].size;
            X.bufcols.len[v] = 1;
            X.bufcols.off[v] = off;
            off += (size_t)A.d->vars[v].size * S.mig_cap;
        }
        X.bytes = (off + 15) & ~(size_t)15;
        CU(cudaMalloc(&X.send_dn, X.bytes));
        CU(cudaMalloc(&X.send_up, X.bytes));
        CU(cudaMalloc(&X.recv_lo, X.bytes));
        CU(cudaMalloc(&X.recv_hi, X.bytes));
        CU(cudaMalloc(&X.d_counts, sizeof(int) * 4 * (size_t)A.d->nstates));
        CU(cudaMemset(X.d_counts, 0, sizeof(int) * 4 * (size_t)A.d->nstates));
        CU(cudaMalloc(&X.holes, sizeof(unsigned int) * 2 * (size_t)S.mig_cap * (size_t)A.d->nstates));
        CU(cudaMalloc(&X.mig_hdr, sizeof(int) * MIG_HDR * (size_t)A.d->nstates));
        CU(cudaMemset(X.mig_hdr, 0, sizeof(int) * MIG_HDR * (size_t)A.d->nstates));
    }
    rc = seed_rand48(s, (unsigned long long)rank * (unsigned long long)m->buffer_size_max,
                     (unsigned long long)world * (unsigned long long)m->buffer_size_max);
    if (rc) return rc;
    /* peer-memory transport */
    S.msg_off.assign(m->nmessages, 0);
    S.agent_off.assign(m->nagents, 0);
    S.msg_seq.assign(m->nmessages, 0);
    S.agent_seq.assign(m->nagents, 0);
    CU(cudaMalloc(&S.d_done, sizeof(unsigned int) * 64));
    CU(cudaMemset(S.d_done, 0, sizeof(unsigned int) * 64));
    CU(cudaMalloc(&S.d_seq, sizeof(int) * 64));
    CU(cudaMemset(S.d_seq, 0, sizeof(int) * 64));
    if (s->opt_slab_p2p) {
        size_t off = 4096; /* flags */
        auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
        for (int i = 0; i < m->nmessages; ++i)
            if (S.msgs[i].bytes) {
                S.msg_off[i] = off;
                off += 4 * align(S.msgs[i].bytes);
            }
        for (int a = 0; a < m->nagents; ++a)
            if (S.agents[a].axis_var >= 0) {
                S.agent_off[a] = off;
                off += 4 * align(S.agents[a].bytes);
            }
        S.xbuf_bytes = off;
        CU(cudaMalloc(&S.xbuf, S.xbuf_bytes));
        CU(cudaMemset(S.xbuf, 0, S.xbuf_bytes));
        cudaIpcMemHandle_t mine, from_lo, from_hi;
        CU(cudaIpcGetMemHandle(&mine, S.xbuf));
        char *hbuf = nullptr;
        CU(cudaMalloc(&hbuf, 3 * sizeof(mine)));
        CU(cudaMemcpy(hbuf, &mine, sizeof(mine), cudaMemcpyHostToDevice));
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(hbuf, sizeof(mine), NCCL_INT8, S.lo_rank, S.comm, s->stream));
        NC(g_nccl.Send(hbuf, sizeof(mine), NCCL_INT8, S.hi_rank, S.comm, s->stream));
        NC(g_nccl.Recv(hbuf + 2 * sizeof(mine), sizeof(mine), NCCL_INT8, S.hi_rank, S.comm, s->stream));
        NC(g_nccl.Recv(hbuf + sizeof(mine), sizeof(mine), NCCL_INT8, S.lo_rank, S.comm, s->stream));
        NC(g_nccl.GroupEnd());
        CU(cudaStreamSynchronize(s->stream));
        CU(cudaMemcpy(&from_lo, hbuf + sizeof(mine), sizeof(mine), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&from_hi, hbuf + 2 * sizeof(mine), sizeof(mine), cudaMemcpyDeviceToHost));
        cudaFree(hbuf);
        cudaError_t e = cudaIpcOpenMemHandle((void **)&S.peer_lo, from_lo, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) {
            if (S.lo_rank == S.hi_rank) S.peer_hi = S.peer_lo;
            else e = cudaIpcOpenMemHandle((void **)&S.peer_hi, from_hi, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            fprintf(stderr, "fgpu: CUDA IPC unavailable (%s): slab exchange falls back to NCCL send/recv\n", cudaGetErrorString(e));
            S.peer_lo = S.peer_hi = nullptr;
        } else {
            S.p2p = true;
        }
    }
    S.async = s->static_kind && s->opt_slab_async;
    if (!s->stream2) {
        int lo_pri = 0, hi_pri = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CU(cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, hi_pri)); /* the exchange must not queue behind the interior agents */
        s->side_priority = hi_pri;
        CU(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    }
    for (int i = 0; i < SlabRT::RING; ++i) CU(cudaEventCreateWithFlags(&S.ring_ev[i], cudaEventDisableTiming));
    CU(cudaMallocHost(&S.h_ring, sizeof(int) * SlabRT::RING * SlabRT::RING_STRIDE));
    for (int a = 0; a < m->nagents; ++a)
        CU(cudaMemcpy(s->agents[a].d_state_count, s->agents[a].h_state_count.data(), sizeof(int) * (size_t)m->agents[a].nstates,
                      cudaMemcpyHostToDevice));
    S.on = true;
    slab_mark_exact(s);
    s->is_static = false; /* populations change every iteration: no graph replay in slab mode */
    invalidate_graph(s);
    return FGPU_OK;
}

extern "C" int fgpu_sim_slab_info(fgpu_sim *s, int *out8) {
    if (!s || !out8) return set_err(FGPU_ERR_ARG, "bad argument");
    const SlabRT &S = s->slab;
    out8[0] = S.on;
    out8[1] = S.axis;
    out8[2] = S.nplanes;
    out8[3] = S.p0;
    out8[4] = S.p1;
    out8[5] = S.halo_cap;
    out8[6] = S.mig_cap;
    out8[7] = S.plane_cells + (S.p2p ? (1 << 30) : 0); /* bit 30: peer-memory transport active */
    return FGPU_OK;
}

/* Flag-only exchange with both neighbours (peer-memory transport): the stream continues once they have reached
 * the same point.  Used by the timed stepping to line the ranks up again after the untimed L2 flush, so that
 * the jitter of the flush is not charged to the iteration that follows it. */
static int slab_neighbour_barrier(fgpu_sim *s) {
    SlabRT &S = s->slab;
    if (!S.on || !S.p2p) return FGPU_OK;
    PushLayout PLY;
    PLY.fixed_bytes = 0;
    PLY.nseg = 0;
    const size_t flag_off = 4096 - 16; /* last slot of the flag page */
    k_push<<<dim3(1, 2), 256, 0, s->stream>>>((const char *)S.d_status, (const char *)S.d_status, S.peer_lo, S.peer_hi, PLY, S.d_done + 63,
                                              (int *)(S.peer_lo + flag_off) + 1, (int *)(S.peer_hi + flag_off) + 0,
                                              (const int *)(S.xbuf + flag_off), (const int *)(S.xbuf + flag_off) + 1, S.d_seq + 63, S.d_status + 3);
    LAUNCH_CHECK();
    return FGPU_OK;
}

/* Halo exchange after a spatial message build: my two boundary planes go to the neighbours, theirs are
 * spliced around my local block (contiguous ranges of the SORTED list, so within-cell order is the owner's). */
static int slab_exchange_halo(fgpu_sim *s, int mi, cudaStream_t st) {
    SlabRT &S = s->slab;
    MsgRT &M = s->msgs[mi];
    SlabMsg &X = S.msgs[mi];
    const int n = M.h_count, PL = S.plane_cells, Q = M.d->rec_quads;
    if (n == 0) CU(cudaMemsetAsync(M.cell_start, 0, sizeof(unsigned int) * ((size_t)M.d->grid_size + 1), st));
    const uint4 *local = M.sorted; /* cell_start of the owned planes already counts from the start of the combined list */
    s->prof_begin(std::string("halo_pack:") + M.d->name);
    const int g = std::min(148 * 4, (std::max(S.halo_cap * Q, PL) + 255) / 256);
    const size_t A4 = (X.bytes + 255) & ~(size_t)255;
    const int seq = ++S.msg_seq[mi], par = seq & 1;
    const int *in_lo = X.recv_lo, *in_hi = X.recv_hi;  /* where the neighbours' planes arrive */
    if (Q < 1 || Q > 8) return set_err(FGPU_ERR_MODEL, "record of %d quads unsupported", Q);
    /* on the side stream the exchange runs beside the interior agents' kernel: its blocks must be dispatched ahead of
     * that kernel's remaining blocks, so the launches carry the priority explicitly (a captured kernel node keeps the
     * attribute) */
    const int gp = std::min(PUSH_BLOCKS * 2, g);
    cudaLaunchAttribute pri_attr[2];
    int nattr = 0;
    if (s->opt_pdl) {
        pri_attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pri_attr[nattr].val.programmaticStreamSerializationAllowed = 1;
        ++nattr;
    }
    if (st == s->stream2) {
        pri_attr[nattr].id = cudaLaunchAttributePriority;
        pri_attr[nattr].val.priority = s->side_priority;
        ++nattr;
    }
    cudaLaunchConfig_t cfg_push = {}, cfg_merge = {};
    cfg_push.gridDim = dim3(gp, 2);
    cfg_push.blockDim = dim3(256);
    cfg_push.stream = st;
    cfg_push.attrs = pri_attr;
    cfg_push.numAttrs = nattr;
    cfg_merge = cfg_push;
    cfg_merge.gridDim = dim3(g);
    if (S.p2p) {
        /* pack + push fused: my low plane is the lower neighbour's HIGH halo and vice versa */
        in_lo = (const int *)(S.xbuf + S.msg_off[mi] + (0 + par) * A4);
        in_hi = (const int *)(S.xbuf + S.msg_off[mi] + (2 + par) * A4);
        switch (Q) {
#define PUSHP_CASE(QQ)                                                                                                              \
    case QQ:                                                                                                                        \
        cudaLaunchKernelEx(&cfg_push, k_push_planes<QQ>, local, (const unsigned int *)M.cell_start, S.p0 * PL, (S.p1 - 1) * PL, PL, S.halo_cap, \
                                                              (int *)(S.peer_lo + S.msg_off[mi] + (2 + par) * A4),                 \
                                                              (int *)(S.peer_hi + S.msg_off[mi] + (0 + par) * A4), S.d_done + mi,  \
                                                              (int *)(S.peer_lo + 16 * mi) + 1, (int *)(S.peer_hi + 16 * mi) + 0,  \
                                                              (const int *)(S.xbuf + 16 * mi), (const int *)(S.xbuf + 16 * mi) + 1, \
                                                              S.d_seq + mi, S.d_status + 3);                                        \
        break;
            PUSHP_CASE(1) PUSHP_CASE(2) PUSHP_CASE(3) PUSHP_CASE(4) PUSHP_CASE(5) PUSHP_CASE(6) PUSHP_CASE(7) PUSHP_CASE(8)
#undef PUSHP_CASE
        }
        LAUNCH_CHECK();
        s->prof_end();
        s->prof_begin(std::string("halo_xfer:") + M.d->name);
    } else {
        switch (Q) {
#define PACK_CASE(QQ)                                                                                                       \
    case QQ:                                                                                                                \
        k_pack_planes<QQ><<<dim3(g, 2), 256, 0, st>>>(local, M.cell_start, S.p0 * PL, (S.p1 - 1) * PL, PL, S.halo_cap, \
                                                             X.send_lo, X.send_hi);                                         \
        break;
            PACK_CASE(1) PACK_CASE(2) PACK_CASE(3) PACK_CASE(4) PACK_CASE(5) PACK_CASE(6) PACK_CASE(7) PACK_CASE(8)
#undef PACK_CASE
        }
        LAUNCH_CHECK();
        s->prof_end();
        s->prof_begin(std::string("halo_xfer:") + M.d->name);
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(X.send_lo, X.bytes, NCCL_INT8, S.lo_rank, S.comm, st));
        NC(g_nccl.Send(X.send_hi, X.bytes, NCCL_INT8, S.hi_rank, S.comm, st));
        NC(g_nccl.Recv(X.recv_hi, X.bytes, NCCL_INT8, S.hi_rank, S.comm, st));
        NC(g_nccl.Recv(X.recv_lo, X.bytes, NCCL_INT8, S.lo_rank, S.comm, st));
        NC(g_nccl.GroupEnd());
    }
    S.halo_bytes += 2 * (long long)X.bytes;
    s->prof_end();
    s->prof_begin(std::string("halo_merge:") + M.d->name);
    const int lo_first = ((S.p0 - 1 + S.nplanes) % S.nplanes) * PL, hi_first = (S.p1 % S.nplanes) * PL;
    switch (Q) {
#define MERGE_CASE(QQ)                                                                                                      \
    case QQ:                                                                                                                \
        cudaLaunchKernelEx(&cfg_merge, k_merge_halo<QQ>, M.sorted, M.cell_start, M.local_at, n, M.d_count, S.p0 * PL, S.ppr * PL, in_lo, \
                           lo_first, in_hi, hi_first, PL, S.d_status);                                                       \
        break;
        MERGE_CASE(1) MERGE_CASE(2) MERGE_CASE(3) MERGE_CASE(4) MERGE_CASE(5) MERGE_CASE(6) MERGE_CASE(7) MERGE_CASE(8)
#undef MERGE_CASE
    }
    LAUNCH_CHECK();
    s->prof_end();
    return FGPU_OK;
}

/* Migration: agents whose plane left the slab are compacted out and sent to the owning neighbour; arrivals
 * are appended after the stayers (lower neighbour's first).  One host round trip at the end learns the new
 * populations (the reference pays two blocking D2H copies after every scan, e.g. sim:1476-1477). */
/* The host-side populations are exact right now: remember them as the base of the upper bounds. */
static void slab_mark_exact(fgpu_sim *s) {
    SlabRT &S = s->slab;
    if (!S.on) return;
    S.exact.assign(SlabRT::RING_STRIDE - 8, 0);
    int slot = 0;
    for (size_t a = 0; a < s->agents.size(); ++a) {
        if (S.agents[a].axis_var < 0) continue;
        for (int st = 0; st < s->agents[a].d->nstates; ++st, ++slot)
            if (slot < SlabRT::RING_STRIDE - 8) S.exact[slot] = s->agents[a].h_state_count[st];
    }
    S.exact_mig = S.migrations;
}

static int slab_check_status(const SlabRT &S, const int *r) {
    if (r[0] == 1 || r[1] || r[2])
        return set_err(FGPU_ERR_OVERFLOW, "slab exchange: fixed-capacity buffer overflow (halo %d, migrants %d, status %d/%d/%d)",
                       S.halo_cap, S.mig_cap, r[0], r[1], r[2]);
    if (r[0] == 2) return set_err(FGPU_ERR_COMM, "slab exchange: an agent moved further than one slab in one step");
    if (r[3]) return set_err(FGPU_ERR_COMM, "slab exchange: timed out waiting for a neighbour's data (peer-memory transport)");
    if (r[4]) return set_err(FGPU_ERR_MODEL, "slab mode: a message was output outside the planes its rank owns (messages must be emitted at the agent's position)");
    return FGPU_OK;
}

/* Collects finished read-backs of the asynchronous mode: newest exact populations + error flags.
 * `wait` blocks until everything enqueued so far has landed. */
static int slab_poll(fgpu_sim *s, bool wait) {
    SlabRT &S = s->slab;
    const fgpu_model *m = s->model;
    for (int k = 0; k < SlabRT::RING; ++k) {
        const long long mig = S.migrations - 1 - k; /* newest first */
        if (mig < 0) break;
        const int slot = (int)(mig % SlabRT::RING);
        if (S.ring_mig[slot] != mig) continue;
        if (wait) CU(cudaEventSynchronize(S.ring_ev[slot]));
        else if (cudaEventQuery(S.ring_ev[slot]) != cudaSuccess) continue;
        const int *r = S.h_ring + slot * SlabRT::RING_STRIDE;
        int rc = slab_check_status(S, r);
        if (rc) return rc;
        if (mig >= S.exact_mig) {
            S.exact.assign(r + 8, r + SlabRT::RING_STRIDE);
            S.exact_mig = mig + 1;
        }
        break;
    }
    /* upper bounds: each migration adds at most 2 * mig_cap agents to a list */
    const long long lag = S.migrations - S.exact_mig;
    int slot = 0;
    for (int a = 0; a < m->nagents; ++a) {
        if (S.agents[a].axis_var < 0) continue;
        for (int st = 0; st < s->agents[a].d->nstates; ++st, ++slot) {
            if (S.exact.empty()) continue;
            const long long ub = (long long)S.exact[slot] + lag * 2LL * S.mig_cap;
            if (lag == 0 && S.exact[slot] > s->agents[a].d->buffer_size)
                return set_err(FGPU_ERR_OVERFLOW, "slab migration: %d agents exceed the buffer of %s", S.exact[slot], s->agents[a].d->name);
            s->agents[a].h_state_count[st] = (int)std::min<long long>(ub, s->agents[a].d->buffer_size);
        }
    }
    return FGPU_OK;
}

/* Migration: agents whose plane left the slab are compacted out and sent to the owning neighbour; arrivals
 * are appended after the stayers (lower neighbour's first).  Synchronous mode: one host round trip at the
 * end learns the new populations (the reference pays two blocking D2H copies after every scan, e.g.
 * sim:1476-1477).  Asynchronous mode: none -- counts stay on the device. */
static int slab_migrate(fgpu_sim *s) {
    SlabRT &S = s->slab;
    const fgpu_model *m = s->model;
    const int ring = (int)(S.migrations % SlabRT::RING);
    int *hdst = S.async ? S.h_ring + ring * SlabRT::RING_STRIDE : S.h_pinned;
    /* a captured iteration carries no read-back: populations and status are read when the stream is settled */
    const bool ringed = S.async && !s->capturing && !S.graph_bounds;
    if (ringed && S.ring_mig[ring] >= 0) CU(cudaEventSynchronize(S.ring_ev[ring])); /* RING migrations old: long done */
    int slot = 0;
    for (int a = 0; a < m->nagents; ++a) {
        AgentRT &A = s->agents[a];
        SlabAgent &X = S.agents[a];
        if (X.axis_var < 0) continue;
        for (int st = 0; st < A.d->nstates; ++st, ++slot) {
            if (8 + slot >= SlabRT::RING_STRIDE) return set_err(FGPU_ERR_MODEL, "too many state lists for slab mode");
            const int n = A.h_state_count[st];           /* exact, or an upper bound in asynchronous mode */
            const int *d_n = S.async ? A.d_state_count + st : nullptr;
            char *list = A.state[st];
            int *flags = (int *)(list + A.d->scan_offset);
            int *cnt = X.d_counts + 4 * st; /* {n_stay, -, total down, total up} */
            const int nb = std::max(1, (n + COMPACT_THREADS - 1) / COMPACT_THREADS);
            const size_t A4 = (X.bytes + 255) & ~(size_t)255;
            const int seq = ++S.agent_seq[a], par = seq & 1;
            const char *in_lo = X.recv_lo, *in_hi = X.recv_hi;
            const float *coord = (const float *)(list + A.d->vars[X.axis_var].list_offset);
            const bool inplace = s->opt_slab_inplace != 0;
            unsigned int *holes = X.holes + 2 * (size_t)S.mig_cap * st;
            int *mhdr = X.mig_hdr + MIG_HDR * st;
            if (inplace) {
                s->prof_begin("mig:flags");
                const bool pdl = s->opt_pdl != 0;
                launch_k(pdl, k_mig_count, dim3(nb), dim3(256), 0, s->stream, coord, n, d_n, S.axis_min, S.axis_max, S.nplanes, S.ppr, S.rank, S.world,
                         A.block_sums, nb, S.d_status);
                LAUNCH_CHECK();
                s->prof_end();
                s->prof_begin("mig:scatter");
                if (S.p2p && s->opt_slab_direct) {
                    /* leavers are packed straight into the neighbours' arrival buffers; the kernel's last block signals and waits
                     * (I am the lower neighbour's upper neighbour: my "down" migrants land in its recv_hi) */
                    char *pdn = S.peer_lo + S.agent_off[a] + (2 + par) * A4, *pup = S.peer_hi + S.agent_off[a] + (0 + par) * A4;
                    launch_k(pdl, k_mig_pack, dim3(nb), dim3(256), 0, s->stream, (const char *)list, pdn, pup, A.cols, X.bufcols, coord, n, d_n,
                             S.axis_min, S.axis_max, S.nplanes, S.ppr, S.rank, S.world, (const unsigned int *)A.block_sums, nb, S.mig_cap, holes, mhdr,
                             (int *)pdn, (int *)pup, S.d_status + 1, S.d_done + 32 + a, (int *)(S.peer_lo + 2048 + 16 * a) + 1,
                             (int *)(S.peer_hi + 2048 + 16 * a) + 0, (const int *)(S.xbuf + 2048 + 16 * a), (const int *)(S.xbuf + 2048 + 16 * a) + 1,
                             S.d_seq + 32 + a, S.d_status + 3);
                } else {
                    launch_k(pdl, k_mig_pack, dim3(nb), dim3(256), 0, s->stream, (const char *)list, X.send_dn, X.send_up, A.cols, X.bufcols, coord, n, d_n,
                             S.axis_min, S.axis_max, S.nplanes, S.ppr, S.rank, S.world, (const unsigned int *)A.block_sums, nb, S.mig_cap, holes, mhdr,
                             (int *)X.send_dn, (int *)X.send_up, S.d_status + 1, (unsigned int *)nullptr, (int *)nullptr, (int *)nullptr,
                             (const int *)nullptr, (const int *)nullptr, (int *)nullptr, (int *)nullptr);
                }
                LAUNCH_CHECK();
                s->prof_end();
            } else {
                s->prof_begin("mig:flags");
                k_owner_count<<<nb, 256, 0, s->stream>>>(coord, n, d_n, S.axis_min, S.axis_max, S.nplanes, S.ppr, S.rank, S.world, flags,
                                                         A.block_sums, nb, S.d_status);
                LAUNCH_CHECK();
                s->prof_end();
                s->prof_begin("mig:scatter");
                k_scatter3<<<nb, COMPACT_THREADS, 0, s->stream>>>(list, A.swap, X.send_dn, X.send_up, A.cols, X.bufcols, flags, A.block_sums, nb, n,
                                                                  d_n, S.mig_cap, S.d_status + 1, cnt, (int *)X.send_dn, (int *)X.send_up);
                LAUNCH_CHECK();
                s->prof_end();
            }
            s->prof_begin("mig:xfer");
            if (S.p2p && inplace && s->opt_slab_direct) {
                in_lo = S.xbuf + S.agent_off[a] + (0 + par) * A4;   /* already there: k_mig_pack of the neighbours wrote and signalled */
                in_hi = S.xbuf + S.agent_off[a] + (2 + par) * A4;
            } else if (S.p2p) {
                PushLayout PLY;
                PLY.fixed_bytes = 16;
                PLY.nseg = X.bufcols.ncols;
                for (int c = 0; c < X.bufcols.ncols; ++c) {
                    PLY.elem[c] = X.bufcols.size[c];
                    PLY.off[c] = X.bufcols.off[c];
                }
                in_lo = S.xbuf + S.agent_off[a] + (0 + par) * A4;
                in_hi = S.xbuf + S.agent_off[a] + (2 + par) * A4;
                /* I am the lower neighbour's upper neighbour: my "down" migrants land in its recv_hi */
                launch_k(s->opt_pdl != 0, k_push, dim3(PUSH_BLOCKS, 2), dim3(256), 0, s->stream, (const char *)X.send_dn, (const char *)X.send_up, S.peer_lo + S.agent_off[a] + (2 + par) * A4,
                                                                    S.peer_hi + S.agent_off[a] + (0 + par) * A4, PLY, S.d_done + 32 + a,
                                                                    (int *)(S.peer_lo + 2048 + 16 * a) + 1,
                                                                    (int *)(S.peer_hi + 2048 + 16 * a) + 0,
                                                                    (const int *)(S.xbuf + 2048 + 16 * a),
                                                                    (const int *)(S.xbuf + 2048 + 16 * a) + 1, S.d_seq + 32 + a,
                                                                    S.d_status + 3);
                LAUNCH_CHECK();
            } else {
                NC(g_nccl.GroupStart());
                NC(g_nccl.Send(X.send_dn, X.bytes, NCCL_INT8, S.lo_rank, S.comm, s->stream));
                NC(g_nccl.Send(X.send_up, X.bytes, NCCL_INT8, S.hi_rank, S.comm, s->stream));
                NC(g_nccl.Recv(X.recv_hi, X.bytes, NCCL_INT8, S.hi_rank, S.comm, s->stream));
                NC(g_nccl.Recv(X.recv_lo, X.bytes, NCCL_INT8, S.lo_rank, S.comm, s->stream));
                NC(g_nccl.GroupEnd());
            }
            S.mig_bytes += 2 * (long long)X.bytes;
            s->prof_end();
            s->prof_begin("mig:append");
            if (inplace) {
                launch_k(s->opt_pdl != 0, k_mig_fill, dim3(std::min(37, (2 * S.mig_cap + 1023) / 1024)), dim3(1024), 0, s->stream, in_lo, in_hi, list, X.bufcols,
                         A.cols, (const int *)mhdr, (const unsigned int *)holes, A.d_state_count + st, A.d->buffer_size, S.d_status + 2);
                LAUNCH_CHECK();
            } else {
                k_append_both<<<(2 * S.mig_cap + 255) / 256, 256, 0, s->stream>>>(in_lo, in_hi, A.swap, X.bufcols, A.cols, cnt,
                                                                                  A.d_state_count + st, A.d->buffer_size, S.d_status + 2);
                LAUNCH_CHECK();
            }
            s->prof_end();
            if (ringed || !S.async) CU(cudaMemcpyAsync(hdst + 8 + slot, A.d_state_count + st, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            if (!inplace) std::swap(A.state[st], A.swap);
            s->touch(A.state[st]);
        }
    }
    S.dirty = false;
    if (S.async && !ringed) {
        S.migrations++;
        return FGPU_OK;
    }
    CU(cudaMemcpyAsync(hdst, S.d_status, sizeof(int) * 8, cudaMemcpyDeviceToHost, s->stream));
    if (S.async) {
        CU(cudaEventRecord(S.ring_ev[ring], s->stream));
        S.ring_mig[ring] = S.migrations;
        S.migrations++;
        return slab_poll(s, false);
    }
    S.migrations++;
    CU(cudaStreamSynchronize(s->stream));
    {
        int rc = slab_check_status(S, hdst);
        if (rc) return rc;
    }
    slot = 0;
    for (int a = 0; a < m->nagents; ++a) {
        if (S.agents[a].axis_var < 0) continue;
        for (int st = 0; st < s->agents[a].d->nstates; ++st, ++slot) {
            const int nn = hdst[8 + slot];
            if (nn > s->agents[a].d->buffer_size)
                return set_err(FGPU_ERR_OVERFLOW, "slab migration: %d agents exceed the buffer of %s", nn, s->agents[a].d->name);
            s->agents[a].h_state_count[st] = nn;
        }
    }
    return FGPU_OK;
}

/* After the stream has drained: make the host-side populations exact again (asynchronous slab mode). */
static int slab_settle(fgpu_sim *s) {
    if (!s->slab.on || !s->slab.async) return FGPU_OK;
    int rc = s->slab.graph_bounds ? FGPU_OK : slab_poll(s, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    int status[8];
    CU(cudaMemcpy(status, s->slab.d_status, sizeof(status), cudaMemcpyDeviceToHost));
    rc = slab_check_status(s->slab, status);
    if (rc) return rc;
    for (size_t a = 0; a < s->agents.size(); ++a) {
        AgentRT &A = s->agents[a];
        CU(cudaMemcpy(A.h_state_count.data(), A.d_state_count, sizeof(int) * (size_t)A.d->nstates, cudaMemcpyDeviceToHost));
        for (int st = 0; st < A.d->nstates; ++st)
            if (A.h_state_count[st] > A.d->buffer_size)
                return set_err(FGPU_ERR_OVERFLOW, "slab migration: %d agents exceed the buffer of %s", A.h_state_count[st], A.d->name);
    }
    s->slab.graph_bounds = false;
    slab_mark_exact(s);
    return FGPU_OK;
}
