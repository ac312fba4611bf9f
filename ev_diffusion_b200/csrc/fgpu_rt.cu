/*
 * fgpu_rt.cu -- host runtime of the B200-native FLAME GPU 1.5 simulation step
 * and its C ABI (include/fgpu.h).
 *
 * This is the model-agnostic replacement for the XSLT-generated simulation.cu:
 *   initialise()            simulation.xslt:472-747   -> fgpu_sim_create / load_states
 *   singleIteration()       simulation.xslt:878-975   -> fgpu_sim_step
 *   <Agent>_<func>(stream)  simulation.xslt:1402-1974 -> Sim::run_function
 * The list / pointer discipline of the reference (working list d_<A>s, swap,
 * new, one list per state; pointer swaps, stable compaction, append) is kept
 * exactly, because list order defines rand48 stream ownership and message
 * order.  What changes is the mechanism underneath:
 *   - no per-launch cudaMemcpyToSymbol / texture binding / occupancy queries;
 *     everything a kernel needs travels in one by-value argument struct;
 *   - models without conditions / death / birth replay one captured CUDA graph
 *     per iteration (zero host round trips);
 *   - spatial message build = fused cell hash in add_<M>_message, hand-written
 *     LSD radix sort, fused gather + CSR partition boundary matrix;
 *   - message consumers process agents in cell order when the consuming list
 *     is the list that produced the messages (locality only, never semantics).
 */
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <vector>

#include "fgpu.h"
#include "fgpu_kernels.cuh"

using namespace fgpu;

/* ------------------------------------------------------------------------- */
/* errors                                                                     */
/* ------------------------------------------------------------------------- */
static thread_local char g_err[1024] = "";

static int set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *fgpu_last_error(void) { return g_err; }

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return set_err(FGPU_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

/* ------------------------------------------------------------------------- */
/* runtime objects                                                            */
/* ------------------------------------------------------------------------- */
struct AgentRT {
    const fgpu_agent *d = nullptr;
    char *work = nullptr;  /* d_<A>s */
    char *swap = nullptr;  /* d_<A>s_swap */
    char *newl = nullptr;  /* d_<A>s_new */
    std::vector<char *> state; /* d_<A>s_<state> */
    int h_count = 0;           /* h_xmachine_memory_<A>_count */
    std::vector<int> h_state_count;
    ColSet cols;
    unsigned int *block_sums = nullptr;
    int *d_total = nullptr;
    int *d_state_count = nullptr; /* slab mode: exact population of each state list, on the device */
};

struct MsgRT {
    const fgpu_message *d = nullptr;
    uint4 *unsorted = nullptr; /* staging written by add_<M>_message */
    uint4 *sorted = nullptr;   /* cell-ordered list read by get_first/next */
    unsigned int *keys_in = nullptr; /* cell key of each unsorted message (written by add_<M>_message) */
    /* optional_message producers write a sparse staging list first (allocated on first use) */
    uint4 *opt_recs = nullptr;
    unsigned int *opt_keys = nullptr;
    int *opt_flags = nullptr;
    unsigned int *keys[2] = {nullptr, nullptr}; /* radix ping-pong */
    unsigned int *vals[2] = {nullptr, nullptr};
    unsigned int *cell_start = nullptr; /* CSR, grid_size + 1 */
    unsigned int *scratch = nullptr;    /* hist | tickets | gap_count | status */
    size_t scratch_bytes = 0;
    unsigned int *gap_queue = nullptr;
    int max_tiles = 0;
    int h_count = 0; /* h_message_<M>_count */
    /* which list produced the current contents (for cell-coherent consumers) */
    const char *producer_buf = nullptr;
    unsigned long long producer_epoch = 0;
    int producers = 0;
    int sorted_in = 0; /* index of keys/vals buffer holding the final sorted pairs */
    bool built = false;
    int local_at = 0;  /* record index of the local block inside `sorted` (slab mode: halo_cap) */
    const int *d_count = nullptr; /* slab mode: exact message count on the device (h_count is an upper bound) */
};

/* ---- 1-D slab decomposition state ---- */
struct SlabMsg {
    int *send_lo = nullptr, *send_hi = nullptr, *recv_lo = nullptr, *recv_hi = nullptr;
    size_t bytes = 0;
};
struct SlabAgent {
    char *send_dn = nullptr, *send_up = nullptr, *recv_lo = nullptr, *recv_hi = nullptr;
    size_t bytes = 0;
    ColSet bufcols; /* column offsets inside a migrant buffer (after the 16-byte header) */
    int axis_var = -1;
    int *d_counts = nullptr; /* per state: {n_stay, n_new} */
};
struct SlabRT {
    bool on = false;
    int rank = 0, world = 1;
    void *comm = nullptr;
    int axis = 2;          /* 2 = z, 1 = y (2-D grids) */
    int nplanes = 1, plane_cells = 1, ppr = 1, p0 = 0, p1 = 1;
    int lo_rank = 0, hi_rank = 0;
    int halo_cap = 0, mig_cap = 0;
    float axis_min = 0.f, axis_max = 1.f;
    std::vector<SlabMsg> msgs;
    std::vector<SlabAgent> agents;
    int *d_status = nullptr;
    int *h_pinned = nullptr; /* pinned mirror: status + counts */
    bool dirty = false;
    long long halo_bytes = 0, mig_bytes = 0; /* NCCL bytes sent per iteration (instrumentation) */
    /* Asynchronous mode (models whose functions are all in-place): populations stay on the device, kernels
     * read them there, the host only tracks upper bounds and learns the exact values a few iterations late
     * through a ring of pinned read-backs -- no host round trip inside an iteration. */
    /* Peer-memory transport (CUDA IPC over NVLink): kernels store straight into the neighbours' receive
     * buffers; NCCL is then only used once, to exchange the IPC handles. */
    bool p2p = false;
    char *xbuf = nullptr;      /* my receive area: flags | per message 4 buffers | per agent 4 buffers */
    size_t xbuf_bytes = 0;
    char *peer_lo = nullptr, *peer_hi = nullptr; /* the neighbours' receive areas, mapped */
    std::vector<size_t> msg_off, agent_off;      /* offset of the first of the 4 buffers (lo0, lo1, hi0, hi1) */
    std::vector<int> msg_seq, agent_seq;         /* exchange sequence numbers */
    unsigned int *d_done = nullptr;              /* last-block counters */
    int *d_seq = nullptr;                        /* exchange numbers on the device (same slots as d_done) */
    bool async = false;
    static const int RING = 4;
    cudaEvent_t ring_ev[RING] = {nullptr, nullptr, nullptr, nullptr};
    long long ring_mig[RING] = {-1, -1, -1, -1}; /* migration number each ring entry belongs to */
    int *h_ring = nullptr;                        /* pinned: RING x RING_STRIDE ints (8 status + counts) */
    static const int RING_STRIDE = 64;
    long long migrations = 0;
    bool graph_bounds = false;    /* graph replay: host-side populations are pinned to the buffer capacity until settled */
    std::vector<int> exact;       /* last exact populations seen (per migrating state list) */
    long long exact_mig = 0;      /* ... as of this migration number */
};

struct ProfEntry {
    std::string name;
    cudaEvent_t a, b;
};

struct fgpu_sim {
    const fgpu_model *model = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<AgentRT> agents;
    std::vector<MsgRT> msgs;
    std::vector<int> gc_count; /* h_<f>_condition_count */
    uint2 *seeds = nullptr;
    unsigned int Ax = 0, Ay = 0, Cx = 0, Cy = 0;
    unsigned int iteration = 0;
    long long launches = 0;
    float last_ms = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::map<const char *, unsigned long long> epoch; /* last structural modification per list buffer */
    unsigned long long epoch_counter = 1;
    /* options */
    int opt_order = 1;
    int opt_nvtx = 0;
    int opt_graph = 1;
    int opt_radix_bits = 9;
    int opt_slab_async = 1;
    int opt_slab_p2p = 1;
    /* CUDA graph of one iteration (static models) */
    bool is_static = false;
    bool static_kind = false; /* every function is an in-place pointer-swap function (regardless of slab mode) */
    cudaGraphExec_t graph = nullptr;
    std::vector<int> graph_counts; /* counts baked into the graph */
    long long graph_launches = 0;
    bool capturing = false;
    /* slab mode: captured iterations keyed by the host-side configuration they start from (list pointers
     * alternate between two buffers from one iteration to the next, so do the receive-buffer parities) */
    struct IterationGraph {
        cudaGraphExec_t exec = nullptr;
        long long launches = 0;
        /* host state after the iteration: absolute values ... */
        std::vector<char *> agent_ptrs; /* per agent: work, swap, state[0..] */
        std::vector<int> agent_h_count;
        std::vector<int> msg_ints;      /* per message: h_count, producers, sorted_in, built */
        std::vector<const char *> msg_producer;
        std::vector<const int *> msg_d_count;
        bool dirty = false;
        /* ... and increments of the monotonic counters */
        long long d_migrations = 0, d_halo_bytes = 0, d_mig_bytes = 0;
        std::vector<int> d_msg_seq, d_agent_seq;
    };
    std::map<std::vector<unsigned long long>, IterationGraph> slab_graphs;
    /* profiling */
    bool profiling = false;
    std::vector<ProfEntry> prof;
    std::vector<std::string> prof_names;
    std::vector<float> prof_ms;
    unsigned int *unpack_tmp = nullptr;
    size_t unpack_tmp_n = 0;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    SlabRT slab;

    void touch(const char *buf) { epoch[buf] = ++epoch_counter; }
    unsigned long long epoch_of(const char *buf) {
        auto it = epoch.find(buf);
        return it == epoch.end() ? 0ull : it->second;
    }
    std::vector<cudaEvent_t> ev_pool; /* events are created once and reused: no creation cost in the timed path */
    size_t ev_used = 0;
    cudaEvent_t pool_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
    /* Stage markers.  With option "nvtx" every stage is also an NVTX range on the host thread that enqueues it
     * (PROFILE_SCOPED_RANGE of the reference, header.xslt:805-863): visible in Nsight Systems, free otherwise. */
    void prof_begin(const std::string &name) {
        if (opt_nvtx) nvtxRangePushA(name.c_str());
        if (!profiling) return;
        ProfEntry e;
        e.name = name;
        e.a = pool_event();
        e.b = pool_event();
        cudaEventRecord(e.a, stream);
        prof.push_back(e);
    }
    void prof_end() {
        if (opt_nvtx) nvtxRangePop();
        if (!profiling) return;
        cudaEventRecord(prof.back().b, stream);
    }
};

static void slab_release(fgpu_sim *s);
static void slab_mark_exact(fgpu_sim *s);
static size_t msg_plane_bytes(const fgpu_message *m) { return (size_t)m->rec_quads * m->buffer_size * sizeof(uint4); }

/* ------------------------------------------------------------------------- */
/* model plugins                                                              */
/* ------------------------------------------------------------------------- */
extern "C" const fgpu_model *fgpu_model_load(const char *plugin_path) {
    if (!plugin_path) {
        set_err(FGPU_ERR_ARG, "fgpu_model_load: null path");
        return nullptr;
    }
    void *h = dlopen(plugin_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        set_err(FGPU_ERR_MODEL, "dlopen(%s): %s", plugin_path, dlerror());
        return nullptr;
    }
    typedef const fgpu_model *(*get_fn)(void);
    get_fn get = (get_fn)dlsym(h, "fgpu_model_get");
    if (!get) {
        set_err(FGPU_ERR_MODEL, "%s does not export fgpu_model_get", plugin_path);
        return nullptr;
    }
    const fgpu_model *m = get();
    if (!m || m->abi_version != FGPU_MODEL_ABI_VERSION) {
        set_err(FGPU_ERR_MODEL, "%s: plugin ABI %d, runtime ABI %d", plugin_path, m ? m->abi_version : -1,
                FGPU_MODEL_ABI_VERSION);
        return nullptr;
    }
    return m;
}

/* ------------------------------------------------------------------------- */
/* create / destroy                                                           */
/* ------------------------------------------------------------------------- */
/* Number of LSD passes for a key of radix_bits with digits of about max_bits: one bit wider digits are
 * accepted when that saves a whole pass (19- and 20-bit keys sort in 2 passes of 10 bits, not 3 of 7). */
static int sort_passes(int radix_bits, int max_bits) {
    const int p = (radix_bits + max_bits - 1) / max_bits;
    const int wide = std::min(max_bits + 1, 11);
    if (p > 1 && max_bits >= 8 && (radix_bits + wide - 1) / wide < p) return p - 1;
    return p;
}
static size_t sort_scratch_words(int n, int key_bits, int max_bits);

static int sim_alloc(fgpu_sim *s) {
    const fgpu_model *m = s->model;
    CU(cudaSetDevice(s->device));
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    s->agents.resize(m->nagents);
    for (int a = 0; a < m->nagents; ++a) {
        AgentRT &A = s->agents[a];
        A.d = &m->agents[a];
        if (A.d->nvars > MAX_COLS) return set_err(FGPU_ERR_MODEL, "agent %s has more than %d variables", A.d->name, MAX_COLS);
        CU(cudaMalloc(&A.work, A.d->list_bytes));
        CU(cudaMalloc(&A.swap, A.d->list_bytes));
        CU(cudaMalloc(&A.newl, A.d->list_bytes));
        CU(cudaMemsetAsync(A.work, 0, A.d->list_bytes, s->stream));
        CU(cudaMemsetAsync(A.swap, 0, A.d->list_bytes, s->stream));
        CU(cudaMemsetAsync(A.newl, 0, A.d->list_bytes, s->stream));
        A.state.resize(A.d->nstates);
        A.h_state_count.assign(A.d->nstates, 0);
        for (int st = 0; st < A.d->nstates; ++st) {
            CU(cudaMalloc(&A.state[st], A.d->list_bytes));
            CU(cudaMemsetAsync(A.state[st], 0, A.d->list_bytes, s->stream));
        }
        A.cols.ncols = A.d->nvars;
        for (int v = 0; v < A.d->nvars; ++v) {
            A.cols.size[v] = A.d->vars[v].size;
            A.cols.off[v] = A.d->vars[v].list_offset;
        }
        const int nb = (A.d->buffer_size + COMPACT_THREADS - 1) / COMPACT_THREADS;
        CU(cudaMalloc(&A.block_sums, sizeof(unsigned int) * 3 * (size_t)(nb + 1)));
        CU(cudaMalloc(&A.d_total, sizeof(int)));
        CU(cudaMalloc(&A.d_state_count, sizeof(int) * (size_t)A.d->nstates));
        CU(cudaMemsetAsync(A.d_state_count, 0, sizeof(int) * (size_t)A.d->nstates, s->stream));
    }
    s->msgs.resize(m->nmessages);
    for (int i = 0; i < m->nmessages; ++i) {
        MsgRT &M = s->msgs[i];
        M.d = &m->messages[i];
        CU(cudaMalloc(&M.unsorted, msg_plane_bytes(M.d)));
        if (M.d->partitioning == FGPU_PART_SPATIAL) {
            CU(cudaMalloc(&M.sorted, msg_plane_bytes(M.d)));
            CU(cudaMalloc(&M.keys_in, sizeof(unsigned int) * (size_t)M.d->buffer_size));
            for (int k = 0; k < 2; ++k) {
                CU(cudaMalloc(&M.keys[k], sizeof(unsigned int) * (size_t)M.d->buffer_size));
                CU(cudaMalloc(&M.vals[k], sizeof(unsigned int) * (size_t)M.d->buffer_size));
            }
            CU(cudaMalloc(&M.cell_start, sizeof(unsigned int) * ((size_t)M.d->grid_size + 1)));
            CU(cudaMemsetAsync(M.cell_start, 0, sizeof(unsigned int) * ((size_t)M.d->grid_size + 1), s->stream));
            M.max_tiles = (M.d->buffer_size + SORT_TILE - 1) / SORT_TILE;
            /* scratch: hist[4][2048] | tickets[4] | gap_count[1] | pad | status[passes][tiles][2048]
             * (a sort of radix_bits key bits never needs more than ceil(bits/4) <= 8... passes; the status
             * area is sized for the worst legal digit choice of this grid) */
            size_t words = 0;
            for (int mb = 4; mb <= SORT_MAXBITS; ++mb)
                if (sort_passes(M.d->radix_bits, mb) <= SORT_MAX_PASSES)
                    words = std::max(words, sort_scratch_words(M.d->buffer_size, M.d->radix_bits, mb));
            M.scratch_bytes = sizeof(unsigned int) * words;
            CU(cudaMalloc(&M.scratch, M.scratch_bytes));
            CU(cudaMalloc(&M.gap_queue, sizeof(unsigned int) * 3 * ((size_t)GAP_LONG_MAX + (size_t)M.d->grid_size / GAP_INLINE + 8)));
        }
    }
    s->gc_count.assign(m->nfunctions, 0);

    /* rand48 seeding, simulation.xslt:661-688 */
    const unsigned int N = (unsigned int)m->buffer_size_max;
    CU(cudaMalloc(&s->seeds, sizeof(uint2) * (size_t)N));
    {
        static const unsigned long long a = 0x5DEECE66DULL, c = 0xB;
        const unsigned long long M48 = (1ULL << 48) - 1ULL;
        unsigned long long A = 1ULL, C = 0ULL;
        for (unsigned int i = 0; i < N; ++i) {
            C += A * c;
            A *= a;
        }
        s->Ax = (unsigned int)(A & 0xFFFFFFULL);
        s->Ay = (unsigned int)((A >> 24) & 0xFFFFFFULL);
        s->Cx = (unsigned int)(C & 0xFFFFFFULL);
        s->Cy = (unsigned int)((C >> 24) & 0xFFFFFFULL);
        std::vector<uint2> h(N);
        unsigned long long x = (((unsigned long long)123) << 16) | 0x330E;
        for (unsigned int i = 0; i < N; ++i) {
            x = a * x + c;
            h[i].x = (unsigned int)(x & 0xFFFFFFULL);
            h[i].y = (unsigned int)((x >> 24) & 0xFFFFFFULL);
        }
        (void)M48;
        CU(cudaMemcpyAsync(s->seeds, h.data(), sizeof(uint2) * (size_t)N, cudaMemcpyHostToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    /* static model: every function is an in-place pointer-swap function with single-message output */
    s->is_static = (m->nstep_functions == 0);
    for (int f = 0; f < m->nfunctions; ++f) {
        const fgpu_function &F = m->functions[f];
        if (!F.swappable || F.output_type != FGPU_OUT_SINGLE) s->is_static = false;
    }
    s->static_kind = s->is_static;
    if (m->init_constants) m->init_constants();
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" fgpu_sim *fgpu_sim_create(const fgpu_model *model, int device) {
    if (!model) {
        set_err(FGPU_ERR_ARG, "fgpu_sim_create: null model");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_err(FGPU_ERR_CUDA, "fgpu_sim_create: no CUDA device (this runtime has no CPU path)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        set_err(FGPU_ERR_ARG, "fgpu_sim_create: device %d out of range (%d devices)", device, ndev);
        return nullptr;
    }
    fgpu_sim *s = new fgpu_sim();
    s->model = model;
    s->device = device;
    if (sim_alloc(s) != FGPU_OK) {
        fgpu_sim_destroy(s);
        return nullptr;
    }
    return s;
}

extern "C" void fgpu_sim_destroy(fgpu_sim *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->graph) cudaGraphExecDestroy(s->graph);
    for (AgentRT &A : s->agents) {
        cudaFree(A.work);
        cudaFree(A.swap);
        cudaFree(A.newl);
        for (char *p : A.state) cudaFree(p);
        cudaFree(A.block_sums);
        cudaFree(A.d_total);
        cudaFree(A.d_state_count);
    }
    for (MsgRT &M : s->msgs) {
        cudaFree(M.unsorted);
        cudaFree(M.sorted);
        cudaFree(M.keys_in);
        cudaFree(M.opt_recs);
        cudaFree(M.opt_keys);
        cudaFree(M.opt_flags);
        for (int k = 0; k < 2; ++k) {
            cudaFree(M.keys[k]);
            cudaFree(M.vals[k]);
        }
        cudaFree(M.cell_start);
        cudaFree(M.scratch);
        cudaFree(M.gap_queue);
    }
    for (SlabMsg &X : s->slab.msgs) {
        cudaFree(X.send_lo);
        cudaFree(X.send_hi);
        cudaFree(X.recv_lo);
        cudaFree(X.recv_hi);
    }
    for (SlabAgent &X : s->slab.agents) {
        cudaFree(X.send_dn);
        cudaFree(X.send_up);
        cudaFree(X.recv_lo);
        cudaFree(X.recv_hi);
        cudaFree(X.d_counts);
    }
    if (s->slab.d_status) cudaFree(s->slab.d_status);
    if (s->slab.h_pinned) cudaFreeHost(s->slab.h_pinned);
    if (s->slab.h_ring) cudaFreeHost(s->slab.h_ring);
    for (int i = 0; i < SlabRT::RING; ++i)
        if (s->slab.ring_ev[i]) cudaEventDestroy(s->slab.ring_ev[i]);
    slab_release(s);
    cudaFree(s->seeds);
    cudaFree(s->unpack_tmp);
    cudaFree(s->flush_buf);
    for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

/* ------------------------------------------------------------------------- */
/* state upload / download                                                    */
/* ------------------------------------------------------------------------- */
static void invalidate_graph(fgpu_sim *s) {
    if (s->graph) {
        cudaGraphExecDestroy(s->graph);
        s->graph = nullptr;
    }
    for (auto &kv : s->slab_graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    s->slab_graphs.clear();
}

static int upload_columns(fgpu_sim *s, int agent, int state, int offset, int count, const void *const *columns) {
    if (agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad agent index %d", agent);
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return set_err(FGPU_ERR_ARG, "bad state index %d", state);
    if (count < 0 || offset + count > A.d->buffer_size)
        return set_err(FGPU_ERR_OVERFLOW, "ERROR: MAX Buffer size (%i) for agent %s exceeded whilst reading data",
                       A.d->buffer_size, A.d->name); /* io.xslt:406, without its off-by-one */
    CU(cudaSetDevice(s->device));
    char *list = A.state[state];
    for (int v = 0; v < A.d->nvars; ++v) {
        const fgpu_var &V = A.d->vars[v];
        char *dst = list + V.list_offset + (size_t)offset * V.size;
        if (columns && columns[v]) {
            CU(cudaMemcpyAsync(dst, columns[v], (size_t)count * V.size, cudaMemcpyHostToDevice, s->stream));
        } else {
            /* defaultValue or 0 (io.xslt:276-291) */
            std::vector<char> tmp((size_t)count * V.size);
            for (int i = 0; i < count; ++i) {
                if (V.size == 8) {
                    if (V.is_float) ((double *)tmp.data())[i] = V.default_value;
                    else ((long long *)tmp.data())[i] = (long long)V.default_value;
                } else {
                    if (V.is_float) ((float *)tmp.data())[i] = (float)V.default_value;
                    else if (!strcmp(V.ctype, "unsigned int")) ((unsigned int *)tmp.data())[i] = (unsigned int)V.default_value;
                    else ((int *)tmp.data())[i] = (int)V.default_value;
                }
            }
            CU(cudaMemcpyAsync(dst, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream));
        }
    }
    const bool same_count = (A.h_state_count[state] == offset + count);
    A.h_state_count[state] = offset + count;
    if (!same_count || s->slab.on) /* d_state_count already holds this value otherwise */
        CU(cudaMemcpyAsync(A.d_state_count + state, &A.h_state_count[state], sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    slab_mark_exact(s);
    s->touch(list);
    /* captured iterations stay valid: they depend on the list buffers, which an upload does not move, and --
     * single-GPU graphs only -- on the list sizes, which step_async compares before every replay */
    return FGPU_OK;
}

extern "C" int fgpu_sim_set_agents(fgpu_sim *s, int agent, int state, int count, const void *const *columns) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    return upload_columns(s, agent, state, 0, count, columns);
}

extern "C" int fgpu_sim_add_agents(fgpu_sim *s, int agent, int state, int count, const void *const *columns) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    if (agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad agent index %d", agent);
    if (state < 0 || state >= s->agents[agent].d->nstates) return set_err(FGPU_ERR_ARG, "bad state index %d", state);
    return upload_columns(s, agent, state, s->agents[agent].h_state_count[state], count, columns);
}

extern "C" int fgpu_sim_agent_count(fgpu_sim *s, int agent, int state) {
    if (!s || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad agent index");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return set_err(FGPU_ERR_ARG, "bad state index");
    return A.h_state_count[state];
}

extern "C" int fgpu_sim_read_agent_var(fgpu_sim *s, int agent, int state, int var, void *dst, int count) {
    if (!s || !dst || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates || var < 0 || var >= A.d->nvars) return set_err(FGPU_ERR_ARG, "bad index");
    if (count < 0 || count > A.d->buffer_size) return set_err(FGPU_ERR_ARG, "bad count");
    CU(cudaSetDevice(s->device));
    const fgpu_var &V = A.d->vars[var];
    CU(cudaMemcpyAsync(dst, A.state[state] + V.list_offset, (size_t)count * V.size, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_agents(fgpu_sim *s, int agent, int state, int count, void *const *columns) {
    if (!s || !columns || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return set_err(FGPU_ERR_ARG, "bad index");
    if (count < 0 || count > A.d->buffer_size) return set_err(FGPU_ERR_ARG, "bad count");
    CU(cudaSetDevice(s->device));
    for (int v = 0; v < A.d->nvars; ++v) {
        if (!columns[v]) continue;
        const fgpu_var &V = A.d->vars[v];
        CU(cudaMemcpyAsync(columns[v], A.state[state] + V.list_offset, (size_t)count * V.size, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_agent_value(fgpu_sim *s, int agent, int state, int var, unsigned int index, void *dst) {
    if (!s || !dst || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates || var < 0 || var >= A.d->nvars) return set_err(FGPU_ERR_ARG, "bad index");
    if (index >= (unsigned int)A.h_state_count[state]) return set_err(FGPU_ERR_ARG, "agent index %u out of range", index);
    CU(cudaSetDevice(s->device));
    const fgpu_var &V = A.d->vars[var];
    CU(cudaMemcpyAsync(dst, A.state[state] + V.list_offset + (size_t)index * V.size, V.size, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" void *fgpu_sim_device_agents(fgpu_sim *s, int agent, int state) {
    if (!s || agent < 0 || agent >= (int)s->agents.size()) return nullptr;
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return nullptr;
    return A.state[state];
}

extern "C" int fgpu_sim_message_count(fgpu_sim *s, int message) {
    if (!s || message < 0 || message >= (int)s->msgs.size()) return set_err(FGPU_ERR_ARG, "bad message index");
    return s->msgs[message].h_count;
}

static int ensure_tmp(fgpu_sim *s, size_t n) {
    if (s->unpack_tmp_n >= n) return FGPU_OK;
    if (s->unpack_tmp) cudaFree(s->unpack_tmp);
    s->unpack_tmp = nullptr;
    s->unpack_tmp_n = 0;
    CU(cudaMalloc(&s->unpack_tmp, n * sizeof(unsigned int)));
    s->unpack_tmp_n = n;
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_message_var(fgpu_sim *s, int message, int var, void *dst, int count) {
    if (!s || !dst || message < 0 || message >= (int)s->msgs.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    MsgRT &M = s->msgs[message];
    if (var < 0 || var >= M.d->nvars || count < 0 || count > M.d->buffer_size) return set_err(FGPU_ERR_ARG, "bad index");
    if (count == 0) return FGPU_OK;
    CU(cudaSetDevice(s->device));
    const fgpu_var &V = M.d->vars[var];
    const uint4 *src = (M.d->partitioning == FGPU_PART_SPATIAL) ? M.sorted + (size_t)M.local_at * M.d->rec_quads : M.unsorted;
    int rc = ensure_tmp(s, (size_t)count);
    if (rc) return rc;
    const int words = V.size / 4;
    std::vector<unsigned int> tmp((size_t)count);
    for (int w = 0; w < words; ++w) {
        k_unpack_words<<<(count + 255) / 256, 256, 0, s->stream>>>((const unsigned int *)src, M.d->rec_quads * 4, count,
                                                                  (int)(V.list_offset / 4) + w, s->unpack_tmp);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(tmp.data(), s->unpack_tmp, (size_t)count * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (words == 1) {
            memcpy(dst, tmp.data(), (size_t)count * 4);
        } else {
            unsigned int *d = (unsigned int *)dst;
            for (int i = 0; i < count; ++i) d[2 * (size_t)i + w] = tmp[i];
        }
    }
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_pbm(fgpu_sim *s, int message, int *start, int *count) {
    if (!s || !start || !count || message < 0 || message >= (int)s->msgs.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    MsgRT &M = s->msgs[message];
    if (M.d->partitioning != FGPU_PART_SPATIAL) return set_err(FGPU_ERR_ARG, "message %s is not spatially partitioned", M.d->name);
    CU(cudaSetDevice(s->device));
    const int cells = M.d->grid_size;
    std::vector<unsigned int> h((size_t)cells + 1, 0u);
    if (M.h_count > 0 && M.built) {
        CU(cudaMemcpyAsync(h.data(), M.cell_start, sizeof(unsigned int) * ((size_t)cells + 1), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    for (int c = 0; c < cells; ++c) {
        start[c] = (int)h[c];
        count[c] = (int)(h[c + 1] - h[c]);
    }
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_message_keys(fgpu_sim *s, int message, unsigned int *keys, int count) {
    if (!s || !keys || message < 0 || message >= (int)s->msgs.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    MsgRT &M = s->msgs[message];
    if (M.d->partitioning != FGPU_PART_SPATIAL || count < 0 || count > M.d->buffer_size) return set_err(FGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(keys, M.keys_in, sizeof(unsigned int) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_message_order(fgpu_sim *s, int message, unsigned int *order, int count) {
    if (!s || !order || message < 0 || message >= (int)s->msgs.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    MsgRT &M = s->msgs[message];
    if (M.d->partitioning != FGPU_PART_SPATIAL || count < 0 || count > M.d->buffer_size) return set_err(FGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(order, M.vals[M.sorted_in], sizeof(unsigned int) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_read_seeds(fgpu_sim *s, unsigned int *dst_xy, int count) {
    if (!s || !dst_xy || count < 0 || count > s->model->buffer_size_max) return set_err(FGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(dst_xy, s->seeds, sizeof(uint2) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_rand48_constants(fgpu_sim *s, unsigned int *out) {
    if (!s || !out) return set_err(FGPU_ERR_ARG, "bad argument");
    out[0] = s->Ax;
    out[1] = s->Ay;
    out[2] = s->Cx;
    out[3] = s->Cy;
    return FGPU_OK;
}

extern "C" int fgpu_sim_set_constant(fgpu_sim *s, const char *name, const void *value) {
    if (!s || !name || !value) return set_err(FGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    for (int i = 0; i < s->model->nconstants; ++i)
        if (!strcmp(s->model->constants[i].name, name)) {
            CU(cudaStreamSynchronize(s->stream));
            s->model->constants[i].set(value);
            CU(cudaDeviceSynchronize());
            return FGPU_OK;
        }
    return set_err(FGPU_ERR_ARG, "unknown environment constant '%s'", name);
}

extern "C" int fgpu_sim_get_constant(fgpu_sim *s, const char *name, void *value) {
    if (!s || !name || !value) return set_err(FGPU_ERR_ARG, "bad argument");
    for (int i = 0; i < s->model->nconstants; ++i)
        if (!strcmp(s->model->constants[i].name, name)) {
            const fgpu_constant &c = s->model->constants[i];
            memcpy(value, c.get(), (size_t)c.size * c.length);
            return FGPU_OK;
        }
    return set_err(FGPU_ERR_ARG, "unknown environment constant '%s'", name);
}

/* ------------------------------------------------------------------------- */
/* 0.xml reader -- character stream, same tag rules as io.xslt:336-583        */
/* ------------------------------------------------------------------------- */
namespace {
struct XmlAgentBuf {
    std::vector<std::vector<char>> cols; /* per variable, raw bytes */
    int count = 0;
};

void store_value(const fgpu_var &V, const char *text, char *dst) {
    if (V.is_float) {
        if (V.size == 4) *(float *)dst = (float)atof(text);      /* fgpu_atof */
        else *(double *)dst = strtod(text, nullptr);             /* fpgu_strtod */
    } else if (V.size == 4) {
        if (!strcmp(V.ctype, "unsigned int")) *(unsigned int *)dst = (unsigned int)strtoul(text, nullptr, 0);
        else *(int *)dst = (int)strtol(text, nullptr, 0);        /* fpgu_strtol(.., 0) */
    } else {
        if (!strncmp(V.ctype, "unsigned", 8)) *(unsigned long long *)dst = strtoull(text, nullptr, 0);
        else *(long long *)dst = strtoll(text, nullptr, 0);
    }
}

void default_value(const fgpu_var &V, char *dst) {
    if (V.is_float) {
        if (V.size == 4) *(float *)dst = (float)V.default_value;
        else *(double *)dst = V.default_value;
    } else if (V.size == 4) {
        if (!strcmp(V.ctype, "unsigned int")) *(unsigned int *)dst = (unsigned int)V.default_value;
        else *(int *)dst = (int)V.default_value;
    } else {
        *(long long *)dst = (long long)V.default_value;
    }
}
}  // namespace

extern "C" int fgpu_sim_load_states(fgpu_sim *s, const char *xml_path) {
    if (!s || !xml_path) return set_err(FGPU_ERR_ARG, "bad argument");
    const fgpu_model *m = s->model;
    CU(cudaSetDevice(s->device));
    if (m->init_constants) m->init_constants(); /* initEnvVars(), io.xslt:254-255 */
    FILE *file = fopen(xml_path, "r");
    if (!file) return set_err(FGPU_ERR_IO, "Could not open input file %s", xml_path);

    std::vector<XmlAgentBuf> bufs(m->nagents);
    /* current values per agent type per variable, reset to defaults after each </xagent> */
    std::vector<std::vector<std::vector<char>>> cur(m->nagents);
    auto reset_cur = [&]() {
        for (int a = 0; a < m->nagents; ++a) {
            cur[a].resize(m->agents[a].nvars);
            for (int v = 0; v < m->agents[a].nvars; ++v) {
                cur[a][v].assign(8, 0);
                default_value(m->agents[a].vars[v], cur[a][v].data());
            }
        }
    };
    reset_cur();
    for (int a = 0; a < m->nagents; ++a) bufs[a].cols.resize(m->agents[a].nvars);

    std::string buffer, agentname;
    bool in_tag = false, in_comment = false, in_itno = false, in_env = false, in_xagent = false, in_name = false;
    bool reading = true;
    std::string open_var; /* the variable tag we are inside of, matched by bare name (io.xslt:458-459) */
    int rc = FGPU_OK;
    int ch;
    while (reading && (ch = fgetc(file)) != EOF) {
        const char c = (char)ch;
        if (buffer.size() >= 10000) {
            rc = set_err(FGPU_ERR_IO, "Error: XML Parsing failed Tag name or content too long (> 10000 characters)");
            break;
        }
        if (in_comment) {
            if (c == '-') buffer.push_back(c);
            else if (c == '>' && buffer.size() >= 2) { in_comment = false; buffer.clear(); }
            else buffer.clear();
        } else if (c == '>') {
            const std::string &tag = buffer;
            if (tag == "/states") reading = false;
            if (tag == "itno") in_itno = true;
            if (tag == "/itno") in_itno = false;
            if (tag == "environment") in_env = true;
            if (tag == "/environment") in_env = false;
            if (tag == "name") in_name = true;
            if (tag == "/name") in_name = false;
            if (tag == "xagent") in_xagent = true;
            if (tag == "/xagent") {
                int a = -1;
                for (int k = 0; k < m->nagents; ++k)
                    if (agentname == m->agents[k].name) a = k;
                if (a < 0) {
                    fprintf(stdout, "Warning: agent name undefined - '%s'\n", agentname.c_str());
                } else {
                    if (bufs[a].count >= m->agents[a].buffer_size) {
                        rc = set_err(FGPU_ERR_OVERFLOW, "ERROR: MAX Buffer size (%i) for agent %s exceeded whilst reading data",
                                     m->agents[a].buffer_size, m->agents[a].name);
                        break;
                    }
                    for (int v = 0; v < m->agents[a].nvars; ++v) {
                        const int sz = m->agents[a].vars[v].size;
                        bufs[a].cols[v].insert(bufs[a].cols[v].end(), cur[a][v].data(), cur[a][v].data() + sz);
                    }
                    bufs[a].count++;
                }
                reset_cur();
                in_xagent = false;
            }
            if (!tag.empty() && tag[0] == '/') {
                if (open_var == tag.substr(1)) open_var.clear();
            } else {
                open_var = tag;
            }
            in_tag = false;
            buffer.clear();
        } else if (c == '<') {
            in_tag = true;
            if (in_itno) { /* parsed into a dead local by the reference (io.xslt:478) */ }
            if (in_name) {
                agentname = buffer;
            } else if (in_xagent) {
                /* every agent type that has a variable of this name receives the value */
                for (int a = 0; a < m->nagents; ++a)
                    for (int v = 0; v < m->agents[a].nvars; ++v)
                        if (open_var == m->agents[a].vars[v].name) store_value(m->agents[a].vars[v], buffer.c_str(), cur[a][v].data());
            } else if (in_env) {
                for (int k = 0; k < m->nconstants; ++k)
                    if (open_var == m->constants[k].name && m->constants[k].length == 1) {
                        fgpu_var V;
                        V.ctype = m->constants[k].ctype;
                        V.size = m->constants[k].size;
                        V.is_float = (!strcmp(V.ctype, "float") || !strcmp(V.ctype, "double"));
                        char tmp[8];
                        store_value(V, buffer.c_str(), tmp);
                        m->constants[k].set(tmp);
                    }
            }
            buffer.clear();
        } else if (in_tag) {
            if (buffer.size() == 2 && c == '-' && buffer[1] == '-' && buffer[0] == '!') {
                in_comment = true;
                buffer.clear();
            }
            buffer.push_back(c);
        } else {
            buffer.push_back(c);
        }
    }
    fclose(file);
    if (rc != FGPU_OK) return rc;
    for (int a = 0; a < m->nagents; ++a) {
        std::vector<const void *> cols(m->agents[a].nvars);
        for (int v = 0; v < m->agents[a].nvars; ++v) cols[v] = bufs[a].cols[v].data();
        /* other states start empty */
        for (int st = 0; st < m->agents[a].nstates; ++st) s->agents[a].h_state_count[st] = 0;
        rc = upload_columns(s, a, m->agents[a].initial_state, 0, bufs[a].count, bufs[a].count ? cols.data() : nullptr);
        if (rc != FGPU_OK) return rc;
    }
    CU(cudaDeviceSynchronize());
    return FGPU_OK;
}

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
                   int *overflow = nullptr, const int *d_n = nullptr) {
    if (n <= 0) {
        if (count_out) *count_out = 0;
        CU(cudaMemsetAsync(A.d_total, 0, sizeof(int), s->stream));
        return FGPU_OK;
    }
    const int nb = (n + COMPACT_THREADS - 1) / COMPACT_THREADS;
    k_flag_blocksum<<<nb, COMPACT_THREADS, 0, s->stream>>>(flags, n, d_n, A.block_sums, match);
    LAUNCH_CHECK();
    k_scan_blocksums<<<1, 1024, 0, s->stream>>>(A.block_sums, nb, A.d_total);
    LAUNCH_CHECK();
    if (count_out) {
        CU(cudaMemcpyAsync(count_out, A.d_total, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
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

/* scratch layout (32-bit words): gap_count | pad[7] | totals[2048] | tile_hist[nbins][tiles] */
constexpr int SCR_GAP = 0, SCR_TOTALS = 8, SCR_TILEHIST = 8 + SORT_MAX_BINS;

static size_t sort_scratch_words(int n, int key_bits, int max_bits) {
    int passes = 1;
    const int digit_bits = choose_digit_bits(key_bits, max_bits, &passes);
    const size_t tiles = ((size_t)n + SORT_TILE - 1) / SORT_TILE;
    return (size_t)SCR_TILEHIST + tiles * ((size_t)1 << digit_bits);
}

template <int BITS>
static cudaError_t launch_scatter(int tiles, cudaStream_t st, const unsigned int *kin, const unsigned int *vin, unsigned int *kout,
                                  unsigned int *vout, int n, const int *d_n, int shift, int digit_bits, const unsigned int *tile_hist,
                                  const unsigned int *totals, unsigned int key_base) {
    static bool configured = false;
    const size_t smem = sizeof(SortSmem<BITS>);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_radix_scatter<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_radix_scatter<BITS><<<tiles, SORT_THREADS, smem, st>>>(kin, vin, kout, vout, n, d_n, shift, digit_bits, tile_hist, totals, tiles, key_base);
    return cudaGetLastError();
}

/* The hand-written LSD radix sort of (keys_in[i], i): replaces cub::DeviceRadixSort::SortPairs (sim:1867).
 * Returns the index (0/1) of the ping-pong pair holding the sorted result; counts launches. */
static int radix_sort_pairs(cudaStream_t st, const unsigned int *keys_in, int n, int key_bits, int max_bits,
                            unsigned int *keys[2], unsigned int *vals[2], unsigned int *scratch, int *sorted_in,
                            long long *launches, const int *d_n = nullptr, unsigned int key_base = 0, unsigned int key_span = 0,
                            int *range_error = nullptr) {
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
        k_tile_hist<<<tiles, SORT_THREADS, sizeof(unsigned int) * (size_t)nbins, st>>>(kin, n, d_n, shift, digit_bits, tile_hist, tiles, key_base,
                                                                                       key_span, p == 0 ? range_error : nullptr);
        k_row_scan<<<nbins, SORT_THREADS, 0, st>>>(tile_hist, tiles, totals);
        cudaError_t e;
        if (digit_bits <= 8) e = launch_scatter<8>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base);
        else if (digit_bits == 9) e = launch_scatter<9>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base);
        else if (digit_bits == 10) e = launch_scatter<10>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base);
        else e = launch_scatter<11>(tiles, st, kin, vin, keys[p & 1], vals[p & 1], n, d_n, shift, digit_bits, tile_hist, totals, key_base);
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
    s->prof_begin(std::string("sort:") + M.d->name);
    CU(cudaMemsetAsync(gap_count, 0, 2 * sizeof(unsigned int), s->stream));
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
    int rc = radix_sort_pairs(s->stream, M.keys_in, n, key_bits, max_bits, M.keys, M.vals, M.scratch, &in, &s->launches, M.d_count, key_base,
                              key_span, range_error);
    if (rc) return rc;
    s->prof_end();
    M.sorted_in = in;
    s->prof_begin(std::string("pbm:") + M.d->name);
    const int g = (n + 255) / 256;
    switch (M.d->rec_quads) {
#define GATHER_CASE(Q)                                                                                                   \
    case Q:                                                                                                              \
        k_gather_pbm<Q><<<g, 256, 0, s->stream>>>(M.keys[in], M.vals[in], M.unsorted, M.sorted + (size_t)M.local_at * Q, n, M.d_count, M.cell_start, \
                                                  key_base, key_base + (key_span ? key_span : (unsigned int)M.d->grid_size),        \
                                                  (unsigned int)M.local_at, M.gap_queue, gap_count);                           \
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
    k_fill_gaps<<<148 * 2, 256, 0, s->stream>>>(M.cell_start, M.gap_queue, gap_count);
    LAUNCH_CHECK();
    s->prof_end();
    return FGPU_OK;
}


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
        size_t off = 16;
        X.bufcols.ncols = A.d->nvars;
        for (int v = 0; v < A.d->nvars; ++v) {
            off = (off + 15) & ~(size_t)15;
            X.bufcols.size[v] = A.d->vars[v].size;
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
static int slab_exchange_halo(fgpu_sim *s, int mi) {
    SlabRT &S = s->slab;
    MsgRT &M = s->msgs[mi];
    SlabMsg &X = S.msgs[mi];
    const int n = M.h_count, PL = S.plane_cells, Q = M.d->rec_quads;
    if (n == 0) CU(cudaMemsetAsync(M.cell_start, 0, sizeof(unsigned int) * ((size_t)M.d->grid_size + 1), s->stream));
    const uint4 *local = M.sorted; /* cell_start of the owned planes already counts from the start of the combined list */
    s->prof_begin(std::string("halo_pack:") + M.d->name);
    const int g = std::min(148 * 4, (std::max(S.halo_cap * Q, PL) + 255) / 256);
    const size_t A4 = (X.bytes + 255) & ~(size_t)255;
    const int seq = ++S.msg_seq[mi], par = seq & 1;
    const int *in_lo = X.recv_lo, *in_hi = X.recv_hi;  /* where the neighbours' planes arrive */
    if (Q < 1 || Q > 8) return set_err(FGPU_ERR_MODEL, "record of %d quads unsupported", Q);
    if (S.p2p) {
        /* pack + push fused: my low plane is the lower neighbour's HIGH halo and vice versa */
        in_lo = (const int *)(S.xbuf + S.msg_off[mi] + (0 + par) * A4);
        in_hi = (const int *)(S.xbuf + S.msg_off[mi] + (2 + par) * A4);
        const int gp = std::min(PUSH_BLOCKS * 2, g);
        switch (Q) {
#define PUSHP_CASE(QQ)                                                                                                              \
    case QQ:                                                                                                                        \
        k_push_planes<QQ><<<dim3(gp, 2), 256, 0, s->stream>>>(local, M.cell_start, S.p0 * PL, (S.p1 - 1) * PL, PL, S.halo_cap,     \
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
        k_pack_planes<QQ><<<dim3(g, 2), 256, 0, s->stream>>>(local, M.cell_start, S.p0 * PL, (S.p1 - 1) * PL, PL, S.halo_cap, \
                                                             X.send_lo, X.send_hi);                                         \
        break;
            PACK_CASE(1) PACK_CASE(2) PACK_CASE(3) PACK_CASE(4) PACK_CASE(5) PACK_CASE(6) PACK_CASE(7) PACK_CASE(8)
#undef PACK_CASE
        }
        LAUNCH_CHECK();
        s->prof_end();
        s->prof_begin(std::string("halo_xfer:") + M.d->name);
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(X.send_lo, X.bytes, NCCL_INT8, S.lo_rank, S.comm, s->stream));
        NC(g_nccl.Send(X.send_hi, X.bytes, NCCL_INT8, S.hi_rank, S.comm, s->stream));
        NC(g_nccl.Recv(X.recv_hi, X.bytes, NCCL_INT8, S.hi_rank, S.comm, s->stream));
        NC(g_nccl.Recv(X.recv_lo, X.bytes, NCCL_INT8, S.lo_rank, S.comm, s->stream));
        NC(g_nccl.GroupEnd());
    }
    S.halo_bytes += 2 * (long long)X.bytes;
    s->prof_end();
    s->prof_begin(std::string("halo_merge:") + M.d->name);
    const int lo_first = ((S.p0 - 1 + S.nplanes) % S.nplanes) * PL, hi_first = (S.p1 % S.nplanes) * PL;
    switch (Q) {
#define MERGE_CASE(QQ)                                                                                                      \
    case QQ:                                                                                                                \
        k_merge_halo<QQ><<<g, 256, 0, s->stream>>>(M.sorted, M.cell_start, M.local_at, n, M.d_count, S.p0 * PL, S.ppr * PL, in_lo, \
                                                   lo_first, in_hi, hi_first, PL, S.d_status);                              \
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
            s->prof_begin("mig:flags");
            k_owner_count<<<nb, 256, 0, s->stream>>>((const float *)(list + A.d->vars[X.axis_var].list_offset), n, d_n,
                                                                 S.axis_min, S.axis_max, S.nplanes, S.ppr, S.rank, S.world, flags,
                                                                 A.block_sums, nb, S.d_status);
            LAUNCH_CHECK();
            s->prof_end();
            const size_t A4 = (X.bytes + 255) & ~(size_t)255;
            const int seq = ++S.agent_seq[a], par = seq & 1;
            const char *in_lo = X.recv_lo, *in_hi = X.recv_hi;
            s->prof_begin("mig:scatter");
            k_scatter3<<<nb, COMPACT_THREADS, 0, s->stream>>>(list, A.swap, X.send_dn, X.send_up, A.cols, X.bufcols, flags, A.block_sums, nb, n,
                                                              d_n, S.mig_cap, S.d_status + 1, cnt, (int *)X.send_dn, (int *)X.send_up);
            LAUNCH_CHECK();
            s->prof_end();
            s->prof_begin("mig:xfer");
            if (S.p2p) {
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
                k_push<<<dim3(PUSH_BLOCKS, 2), 256, 0, s->stream>>>(X.send_dn, X.send_up, S.peer_lo + S.agent_off[a] + (2 + par) * A4,
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
            k_append_both<<<(2 * S.mig_cap + 255) / 256, 256, 0, s->stream>>>(in_lo, in_hi, A.swap, X.bufcols, A.cols, cnt,
                                                                              A.d_state_count + st, A.d->buffer_size, S.d_status + 2);
            LAUNCH_CHECK();
            s->prof_end();
            if (ringed || !S.async) CU(cudaMemcpyAsync(hdst + 8 + slot, A.d_state_count + st, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            std::swap(A.state[st], A.swap);
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

/* ------------------------------------------------------------------------- */
/* one agent function: <Agent>_<func>(stream), simulation.xslt:1402-1974      */
/* ------------------------------------------------------------------------- */
static int run_function(fgpu_sim *s, int fi) {
    const fgpu_model *m = s->model;
    const fgpu_function &F = m->functions[fi];
    AgentRT &A = s->agents[F.agent];
    const int cur = F.current_state, nxt = F.next_state;

    if (s->slab.on) {
        /* every rank must take part in every exchange: empty lists still run their (empty) functions */
        if (s->slab.dirty && F.output_message >= 0 && s->msgs[F.output_message].d->partitioning == FGPU_PART_SPATIAL) {
            int rc = slab_migrate(s);
            if (rc) return rc;
        }
    } else if (A.h_state_count[cur] == 0)
        return FGPU_OK; /* sim:1417-1420 */
    const int state_list_size = A.h_state_count[cur];

    if (F.xagent_output >= 0) {
        /* reset_<B>_scan_input(d_<B>s_new), sim:1427-1435 */
        AgentRT &B = s->agents[F.xagent_output];
        const int nz = std::min(state_list_size, B.d->buffer_size);
        CU(cudaMemsetAsync(B.newl + B.d->scan_offset, 0, sizeof(int) * (size_t)nz, s->stream));
    }

    if (F.condition == FGPU_COND_FUNCTION) {
        /* sim:1439-1528 */
        A.h_count = A.h_state_count[cur];
        fgpu_filter fa;
        fa.current = A.state[cur];
        fa.working = A.work;
        fa.n = A.h_count;
        fa.flags_current = (int *)(A.state[cur] + A.d->scan_offset);
        fa.flags_working = (int *)(A.work + A.d->scan_offset);
        s->prof_begin(std::string("cond:") + F.name);
        F.filter(&fa, s->stream);
        LAUNCH_CHECK();
        int kept = 0, working = 0;
        int rc = compact(s, A, A.state[cur], A.swap, 0, A.h_count, (const int *)(A.state[cur] + A.d->scan_offset), true, &kept);
        if (rc) return rc;
        A.h_state_count[cur] = kept;
        std::swap(A.state[cur], A.swap);
        rc = compact(s, A, A.work, A.swap, 0, A.h_count, (const int *)(A.work + A.d->scan_offset), true, &working);
        if (rc) return rc;
        A.h_count = working;
        std::swap(A.work, A.swap);
        s->prof_end();
        if (A.h_count == 0) return FGPU_OK;
    } else if (F.condition == FGPU_COND_GLOBAL) {
        /* sim:1530-1588 */
        A.h_count = A.h_state_count[cur];
        fgpu_filter fa;
        fa.current = A.state[cur];
        fa.working = nullptr;
        fa.n = A.h_count;
        fa.flags_current = (int *)(A.state[cur] + A.d->scan_offset);
        fa.flags_working = nullptr;
        F.filter(&fa, s->stream);
        LAUNCH_CHECK();
        int ntrue = 0;
        int rc = compact(s, A, nullptr, nullptr, 0, A.h_count, (const int *)(A.state[cur] + A.d->scan_offset), false, &ntrue);
        if (rc) return rc;
        const bool mismatch = F.gc_must_evaluate_to ? (ntrue != A.h_count) : (ntrue == A.h_count);
        if (mismatch && s->gc_count[fi] < F.gc_max_iterations) {
            s->gc_count[fi]++;
            return FGPU_OK;
        }
        if (s->gc_count[fi] == F.gc_max_iterations)
            printf("Global agent condition for %s function reached the maximum number of %d conditions\n", F.name,
                   F.gc_max_iterations);
        s->gc_count[fi] = 0;
        std::swap(A.work, A.state[cur]);
        A.h_state_count[cur] = 0;
    } else {
        /* sim:1590-1601: currentState maps to working list */
        std::swap(A.work, A.state[cur]);
        A.h_count = A.h_state_count[cur];
        A.h_state_count[cur] = 0;
    }

    MsgRT *Mout = F.output_message >= 0 ? &s->msgs[F.output_message] : nullptr;
    MsgRT *Min = F.input_message >= 0 ? &s->msgs[F.input_message] : nullptr;
    if (Mout && Mout->h_count + A.h_count > Mout->d->buffer_size) /* sim:1607-1613 */
        return set_err(FGPU_ERR_OVERFLOW, "Error: Buffer size of %s message will be exceeded in function %s", Mout->d->name, F.name);
    const bool optional_out = Mout && F.output_type == FGPU_OUT_OPTIONAL;
    if (optional_out) {
        if (Mout->h_count > 0) /* the reference swaps the list away under an earlier producer (sim:1709-1716) */
            return set_err(FGPU_ERR_MODEL, "function %s: an optional_message output must be the first producer of message %s in an iteration", F.name,
                           Mout->d->name);
        if (s->slab.on) return set_err(FGPU_ERR_MODEL, "function %s: optional_message outputs are not supported in slab mode", F.name);
        if (!Mout->opt_recs) {
            CU(cudaMalloc(&Mout->opt_recs, msg_plane_bytes(Mout->d)));
            CU(cudaMalloc(&Mout->opt_keys, sizeof(unsigned int) * (size_t)Mout->d->buffer_size));
            CU(cudaMalloc(&Mout->opt_flags, sizeof(int) * (size_t)Mout->d->buffer_size));
        }
    }

    fgpu_launch L;
    memset(&L, 0, sizeof(L));
    const bool slab_async = s->slab.on && s->slab.async;
    L.agents = A.work;
    L.n = A.h_count;
    L.d_n = slab_async ? A.d_state_count + cur : nullptr; /* in-place function: the working list is the state list */
    L.order = nullptr;
    if (F.xagent_output >= 0) L.new_agents = s->agents[F.xagent_output].newl;
    if (Min) {
        L.in_recs = (Min->d->partitioning == FGPU_PART_SPATIAL) ? (const void *)Min->sorted : (const void *)Min->unsorted;
        L.in_cell_start = Min->cell_start;
        L.in_count = Min->h_count;
        L.d_in_count = slab_async ? Min->d_count : nullptr;
        /* cell-coherent processing order: only when this very list produced the
         * messages (message k <-> list index k) and has not been restructured since */
        if (s->opt_order && Min->d->partitioning == FGPU_PART_SPATIAL && Min->producers == 1 &&
            Min->producer_buf == A.work && Min->producer_epoch == s->epoch_of(A.work) && Min->h_count == A.h_count &&
            Min->built) {
            L.order = Min->vals[Min->sorted_in];
        }
    }
    if (Mout) {
        L.out_recs = Mout->unsorted;
        L.out_keys = Mout->keys_in;
        L.out_base = Mout->h_count; /* a second producer appends; the whole list is re-binned (sim:1836-1837) */
        if (optional_out) {
            /* reset_<M>_swaps (sim:1666-1672), then every agent that calls add_<M>_message flags its own slot */
            CU(cudaMemsetAsync(Mout->opt_flags, 0, sizeof(int) * (size_t)std::max(A.h_count, 1), s->stream));
            L.out_recs = Mout->opt_recs;
            L.out_keys = Mout->opt_keys;
            L.out_flags = Mout->opt_flags;
            L.out_base = 0;
        }
    }
    if (F.rng) {
        L.seeds = s->seeds;
        L.Ax = s->Ax;
        L.Ay = s->Ay;
        L.Cx = s->Cx;
        L.Cy = s->Cy;
    }
    s->prof_begin(std::string("func:") + F.name);
    F.launch(&L, s->stream);
    LAUNCH_CHECK();
    s->prof_end();

    if (optional_out) {
        /* exclusive scan of the flags + scatter_optional_<M>_messages + the count read-back the reference also
         * pays (sim:1709-1750): the message list is the stable compaction of the flagged slots */
        s->prof_begin(std::string("optional:") + F.name);
        int produced = 0;
        int rc = compact(s, A, nullptr, nullptr, 0, A.h_count, Mout->opt_flags, false, &produced);
        if (rc) return rc;
        if (A.h_count > 0) {
            const int nb = (A.h_count + COMPACT_THREADS - 1) / COMPACT_THREADS;
            const bool spatial = Mout->d->partitioning == FGPU_PART_SPATIAL;
            switch (Mout->d->rec_quads) {
#define OPT_CASE(QQ)                                                                                                                      \
    case QQ:                                                                                                                              \
        k_compact_records<QQ><<<nb, COMPACT_THREADS, 0, s->stream>>>(Mout->opt_recs, spatial ? Mout->opt_keys : nullptr, Mout->unsorted, \
                                                                     Mout->keys_in, Mout->opt_flags, A.block_sums, 0, A.h_count);        \
        break;
                OPT_CASE(1) OPT_CASE(2) OPT_CASE(3) OPT_CASE(4) OPT_CASE(5) OPT_CASE(6) OPT_CASE(7) OPT_CASE(8)
#undef OPT_CASE
                default:
                    return set_err(FGPU_ERR_MODEL, "message %s: record of %d quads is not supported", Mout->d->name, Mout->d->rec_quads);
            }
            LAUNCH_CHECK();
        }
        Mout->producers = 1;
        Mout->producer_buf = nullptr; /* message k is no longer the message of list index k: no cell-coherent order */
        Mout->producer_epoch = 0;
        Mout->d_count = nullptr;
        Mout->h_count = produced;
        s->prof_end();
    } else if (Mout) {
        /* sim:1734-1754 */
        if (Mout->h_count == 0) {
            Mout->producers = 1;
            Mout->producer_buf = A.work;
            Mout->producer_epoch = s->epoch_of(A.work);
            Mout->d_count = slab_async ? L.d_n : nullptr;
        } else {
            Mout->producers++;
            if (slab_async)
                return set_err(FGPU_ERR_MODEL, "message %s has two producers in one iteration: not supported in asynchronous slab mode", Mout->d->name);
        }
        Mout->h_count += A.h_count;
    }

    const int pre_death_count = A.h_count;
    if (F.reallocate) {
        /* sim:1763-1792 */
        s->prof_begin(std::string("death:") + F.name);
        int alive = 0;
        int rc = compact(s, A, A.work, A.swap, 0, A.h_count, (const int *)(A.work + A.d->scan_offset), true, &alive);
        if (rc) return rc;
        std::swap(A.work, A.swap);
        A.h_count = alive;
        s->prof_end();
    }
    if (F.xagent_output >= 0) {
        /* sim:1794-1829 */
        AgentRT &B = s->agents[F.xagent_output];
        const int st = F.xagent_output_state;
        s->prof_begin(std::string("birth:") + F.name);
        int born = 0;
        int rc = compact(s, B, nullptr, nullptr, 0, pre_death_count, (const int *)(B.newl + B.d->scan_offset), false, &born);
        if (rc) return rc;
        if (B.h_state_count[st] + born > B.d->buffer_size)
            return set_err(FGPU_ERR_OVERFLOW, "Error: Buffer size of %s agents in state %s will be exceeded writing new agents in function %s",
                           B.d->name, B.d->states[st], F.name);
        const int nb = (pre_death_count + COMPACT_THREADS - 1) / COMPACT_THREADS;
        k_compact_scatter<<<nb, COMPACT_THREADS, 0, s->stream>>>(B.newl, B.state[st], B.cols, B.cols,
                                                                 (const int *)(B.newl + B.d->scan_offset), B.block_sums,
                                                                 B.h_state_count[st], pre_death_count, nullptr, 1, 0x7fffffff, nullptr);
        LAUNCH_CHECK();
        s->touch(B.state[st]);
        B.h_state_count[st] += born;
        s->prof_end();
    }
    if (Mout && Mout->d->partitioning == FGPU_PART_SPATIAL) {
        int rc = build_spatial(s, *Mout);
        if (rc) return rc;
        if (s->slab.on) {
            rc = slab_exchange_halo(s, F.output_message);
            if (rc) return rc;
        }
    }
    if (s->slab.on && (F.writes_position || F.xagent_output >= 0)) s->slab.dirty = true;
    /* move agents to next state, sim:1936-1972 */
    if (A.h_state_count[nxt] + A.h_count > A.d->buffer_size)
        return set_err(FGPU_ERR_OVERFLOW, "Error: Buffer size of %s agents in state %s will be exceeded moving working agents to next state in function %s",
                       F.name, A.d->states[nxt], F.name);
    if (F.swappable) {
        std::swap(A.work, A.state[cur]);
    } else if (A.h_count > 0) {
        s->prof_begin(std::string("append:") + F.name);
        k_append<<<(A.h_count + 255) / 256, 256, 0, s->stream>>>(A.work, A.state[nxt], A.cols, A.h_state_count[nxt], A.h_count);
        LAUNCH_CHECK();
        s->touch(A.state[nxt]);
        s->prof_end();
    }
    A.h_state_count[nxt] += A.h_count;
    return FGPU_OK;
}

static int begin_iteration(fgpu_sim *s) {
    s->iteration++;
    for (MsgRT &M : s->msgs) { /* sim:888-892 */
        M.h_count = 0;
        M.producers = 0;
        M.producer_buf = nullptr;
        M.built = false;
        M.d_count = nullptr;
    }
    return FGPU_OK;
}

static int single_iteration(fgpu_sim *s) {
    const fgpu_model *m = s->model;
    begin_iteration(s);
    for (int l = 0; l < m->nlayers; ++l)
        for (int k = 0; k < m->layers[l].nfunctions; ++k) {
            int rc = run_function(s, m->layers[l].functions[k]);
            if (rc) return rc;
        }
    if (s->slab.on && s->slab.dirty) {
        int rc = slab_migrate(s);
        if (rc) return rc;
    }
    if (!s->capturing)
        for (int k = 0; k < m->nstep_functions; ++k) {
            CU(cudaStreamSynchronize(s->stream));
            if (m->bind) m->bind(s);
            m->step_functions[k]();
        }
    return FGPU_OK;
}

static std::vector<int> current_counts(fgpu_sim *s) {
    std::vector<int> c;
    for (AgentRT &A : s->agents)
        for (int v : A.h_state_count) c.push_back(v);
    c.push_back(s->opt_order);
    c.push_back(s->opt_radix_bits);
    return c;
}

/* ---- slab mode: replay of captured iterations ------------------------------------------------------
 * With device-side populations, device-side exchange numbers and peer-memory transport, an iteration of
 * a static model in slab mode is a fixed sequence of launches -- except that migration swaps each state
 * list with its swap buffer and the receive buffers alternate, so consecutive iterations start from
 * (two to four) different host-side configurations.  Each configuration gets its own captured graph; the
 * host bookkeeping an iteration performs (pointer swaps, counters) is recorded at capture and re-applied at
 * replay.  Launch grids are sized for the buffer capacity (kernels read the exact populations on the
 * device), and populations / error flags are read when the stream is settled. */
static std::vector<unsigned long long> slab_graph_key(fgpu_sim *s) {
    std::vector<unsigned long long> k;
    for (AgentRT &A : s->agents) {
        k.push_back((unsigned long long)(uintptr_t)A.work);
        k.push_back((unsigned long long)(uintptr_t)A.swap);
        for (char *p : A.state) k.push_back((unsigned long long)(uintptr_t)p);
    }
    for (int v : s->slab.msg_seq) k.push_back((unsigned long long)(v & 1));
    for (int v : s->slab.agent_seq) k.push_back((unsigned long long)(v & 1));
    k.push_back(s->slab.dirty ? 1ull : 0ull);
    k.push_back((unsigned long long)s->opt_order);
    k.push_back((unsigned long long)s->opt_radix_bits);
    return k;
}

static void slab_graph_record_after(fgpu_sim *s, fgpu_sim::IterationGraph &G) {
    G.agent_ptrs.clear();
    G.agent_h_count.clear();
    for (AgentRT &A : s->agents) {
        G.agent_ptrs.push_back(A.work);
        G.agent_ptrs.push_back(A.swap);
        for (char *p : A.state) G.agent_ptrs.push_back(p);
        G.agent_h_count.push_back(A.h_count);
    }
    G.msg_ints.clear();
    G.msg_producer.clear();
    G.msg_d_count.clear();
    for (MsgRT &M : s->msgs) {
        G.msg_ints.push_back(M.h_count);
        G.msg_ints.push_back(M.producers);
        G.msg_ints.push_back(M.sorted_in);
        G.msg_ints.push_back(M.built ? 1 : 0);
        G.msg_producer.push_back(M.producer_buf);
        G.msg_d_count.push_back(M.d_count);
    }
    G.dirty = s->slab.dirty;
}

static void slab_graph_apply_after(fgpu_sim *s, const fgpu_sim::IterationGraph &G) {
    size_t k = 0;
    for (size_t a = 0; a < s->agents.size(); ++a) {
        AgentRT &A = s->agents[a];
        A.work = G.agent_ptrs[k++];
        A.swap = G.agent_ptrs[k++];
        for (char *&p : A.state) p = G.agent_ptrs[k++];
        A.h_count = G.agent_h_count[a];
        /* list contents changed: any producer/consumer pairing recorded before the replay is void */
        s->touch(A.work);
        for (char *p : A.state) s->touch(p);
    }
    for (size_t m = 0; m < s->msgs.size(); ++m) {
        MsgRT &M = s->msgs[m];
        M.h_count = G.msg_ints[4 * m + 0];
        M.producers = G.msg_ints[4 * m + 1];
        M.sorted_in = G.msg_ints[4 * m + 2];
        M.built = G.msg_ints[4 * m + 3] != 0;
        M.producer_buf = G.msg_producer[m];
        M.producer_epoch = M.producer_buf ? s->epoch_of(M.producer_buf) : 0;
        M.d_count = G.msg_d_count[m];
    }
    SlabRT &S = s->slab;
    S.dirty = G.dirty;
    S.migrations += G.d_migrations;
    S.halo_bytes += G.d_halo_bytes;
    S.mig_bytes += G.d_mig_bytes;
    for (size_t i = 0; i < S.msg_seq.size(); ++i) S.msg_seq[i] += G.d_msg_seq[i];
    for (size_t i = 0; i < S.agent_seq.size(); ++i) S.agent_seq[i] += G.d_agent_seq[i];
    s->iteration++;
    s->launches += G.launches;
}

static int slab_graph_step(fgpu_sim *s, int n) {
    SlabRT &S = s->slab;
    if (!S.graph_bounds) {
        /* grids for the capacity; the exact populations are on the device (d_state_count) */
        for (size_t a = 0; a < s->agents.size(); ++a) {
            if (S.agents[a].axis_var < 0) continue;
            for (int &c : s->agents[a].h_state_count) c = s->agents[a].d->buffer_size;
        }
        S.graph_bounds = true;
    }
    for (int i = 0; i < n; ++i) {
        const std::vector<unsigned long long> key = slab_graph_key(s);
        auto it = s->slab_graphs.find(key);
        if (it != s->slab_graphs.end()) {
            CU(cudaGraphLaunch(it->second.exec, s->stream));
            slab_graph_apply_after(s, it->second);
            continue;
        }
        fgpu_sim::IterationGraph G;
        const long long launches0 = s->launches, mig0 = S.migrations, hb0 = S.halo_bytes, mb0 = S.mig_bytes;
        const std::vector<int> ms0 = S.msg_seq, as0 = S.agent_seq;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        s->capturing = true;
        int rc = single_iteration(s); /* performs the host bookkeeping of the iteration; the work is only recorded */
        s->capturing = false;
        cudaError_t e = cudaStreamEndCapture(s->stream, &g);
        if (rc) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "cudaStreamEndCapture -> %s", cudaGetErrorString(e));
        CU(cudaGraphInstantiate(&G.exec, g, 0));
        CU(cudaGraphDestroy(g));
        G.launches = s->launches - launches0;
        G.d_migrations = S.migrations - mig0;
        G.d_halo_bytes = S.halo_bytes - hb0;
        G.d_mig_bytes = S.mig_bytes - mb0;
        for (size_t k = 0; k < ms0.size(); ++k) G.d_msg_seq.push_back(S.msg_seq[k] - ms0[k]);
        for (size_t k = 0; k < as0.size(); ++k) G.d_agent_seq.push_back(S.agent_seq[k] - as0[k]);
        slab_graph_record_after(s, G);
        CU(cudaGraphLaunch(G.exec, s->stream)); /* the host state is already the one after this iteration */
        s->slab_graphs.emplace(key, std::move(G));
    }
    return FGPU_OK;
}

static int step_async(fgpu_sim *s, int n) {
    if (n <= 0) return FGPU_OK;
    CU(cudaSetDevice(s->device));
    const bool use_graph = s->is_static && s->opt_graph && !s->profiling;
    CU(cudaEventRecord(s->ev0, s->stream));
    if (s->slab.on && s->slab.async && s->slab.p2p && s->opt_graph && !s->profiling && s->model->nstep_functions == 0) {
        int rc = slab_graph_step(s, n);
        if (rc) return rc;
    } else if (use_graph) {
        if (!s->graph || s->graph_counts != current_counts(s)) {
            invalidate_graph(s);
            cudaGraph_t g = nullptr;
            const long long before = s->launches;
            const unsigned int it = s->iteration;
            CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            s->capturing = true;
            int rc = single_iteration(s);
            s->capturing = false;
            cudaError_t e = cudaStreamEndCapture(s->stream, &g);
            s->iteration = it;
            s->graph_launches = s->launches - before;
            s->launches = before;
            if (rc) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "cudaStreamEndCapture -> %s", cudaGetErrorString(e));
            CU(cudaGraphInstantiate(&s->graph, g, 0));
            CU(cudaGraphDestroy(g));
            s->graph_counts = current_counts(s);
            /* the event recorded before the capture is still valid */
        }
        for (int i = 0; i < n; ++i) {
            CU(cudaGraphLaunch(s->graph, s->stream));
            s->launches += s->graph_launches;
        }
        /* host-side bookkeeping of one iteration is idempotent for static models */
        s->iteration += (unsigned int)n;
    } else {
        for (int i = 0; i < n; ++i) {
            int rc = single_iteration(s);
            if (rc) return rc;
        }
    }
    CU(cudaEventRecord(s->ev1, s->stream));
    return FGPU_OK;
}

extern "C" int fgpu_sim_step_async(fgpu_sim *s, int n) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    return step_async(s, n);
}

extern "C" int fgpu_sim_sync(fgpu_sim *s) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    return slab_settle(s);
}

extern "C" int fgpu_sim_step(fgpu_sim *s, int n) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    int rc = step_async(s, n);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    if (n > 0) CU(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
    return slab_settle(s);
}

/* n iterations, each timed separately with a cudaEvent pair on the simulation
 * stream; `flush_bytes` > 0 writes a scratch buffer of that size before every
 * iteration (outside the timed span) so that each iteration starts with a cold L2. */
extern "C" int fgpu_sim_step_timed(fgpu_sim *s, int n, size_t flush_bytes, float *ms_each) {
    if (!s || n < 0) return set_err(FGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    if (flush_bytes > s->flush_bytes) {
        if (s->flush_buf) cudaFree(s->flush_buf);
        s->flush_buf = nullptr;
        s->flush_bytes = 0;
        CU(cudaMalloc(&s->flush_buf, flush_bytes));
        s->flush_bytes = flush_bytes;
    }
    std::vector<cudaEvent_t> ev(2 * (size_t)n);
    for (auto &e : ev) CU(cudaEventCreate(&e));
    int rc = FGPU_OK;
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n && rc == FGPU_OK; ++i) {
        if (flush_bytes) CU(cudaMemsetAsync(s->flush_buf, i & 0xff, flush_bytes, s->stream));
        if (flush_bytes && s->slab.on) rc = slab_neighbour_barrier(s); /* untimed: ranks start the iteration together */
        if (rc) break;
        CU(cudaEventRecord(ev[2 * i], s->stream));
        rc = step_async(s, 1);
        CU(cudaEventRecord(ev[2 * i + 1], s->stream));
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (getenv("FGPU_DEBUG_TIMING"))
        fprintf(stderr, "fgpu: enqueue of %d iterations took %.3f ms of host time (%.1f us / iteration)\n", n,
                (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6,
                ((t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3) / std::max(n, 1));
    CU(cudaStreamSynchronize(s->stream));
    float total = 0.f;
    for (int i = 0; i < n; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
        if (ms_each) ms_each[i] = t;
        total += t;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    s->last_ms = total;
    if (rc == FGPU_OK) rc = slab_settle(s);
    return rc;
}

extern "C" int fgpu_sim_begin_iteration(fgpu_sim *s) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    return begin_iteration(s);
}

extern "C" int fgpu_sim_run_function(fgpu_sim *s, int function) {
    if (!s || function < 0 || function >= s->model->nfunctions) return set_err(FGPU_ERR_ARG, "bad function index");
    CU(cudaSetDevice(s->device));
    int rc = run_function(s, function);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return slab_settle(s);
}

extern "C" unsigned int fgpu_sim_iteration(const fgpu_sim *s) { return s ? s->iteration : 0u; }
extern "C" long long fgpu_sim_launch_count(const fgpu_sim *s) { return s ? s->launches : 0; }
extern "C" float fgpu_sim_last_step_ms(const fgpu_sim *s) { return s ? s->last_ms : 0.f; }

extern "C" int fgpu_sim_run_init_functions(fgpu_sim *s) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    CU(cudaSetDevice(s->device));
    if (s->model->bind) s->model->bind(s);
    for (int k = 0; k < s->model->ninit_functions; ++k) s->model->init_functions[k]();
    CU(cudaDeviceSynchronize());
    return FGPU_OK;
}

extern "C" int fgpu_sim_run_exit_functions(fgpu_sim *s) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    if (s->model->bind) s->model->bind(s);
    for (int k = 0; k < s->model->nexit_functions; ++k) s->model->exit_functions[k]();
    return FGPU_OK;
}

/* ------------------------------------------------------------------------- */
/* state writer (io.xslt:124-200) and binary snapshot                         */
/* ------------------------------------------------------------------------- */
namespace {
/* formatSpecifier, _common_templates.xslt:25-60: integers by width and sign, everything else "%f" */
int write_value(FILE *f, const char *ctype, int size, bool is_float, const void *p) {
    if (is_float) return size == 8 ? fprintf(f, "%f", *(const double *)p) : fprintf(f, "%f", *(const float *)p);
    const bool uns = strstr(ctype, "unsigned") != nullptr;
    switch (size) {
        case 1: return uns ? fprintf(f, "%u", (unsigned)*(const unsigned char *)p) : fprintf(f, "%d", (int)*(const signed char *)p);
        case 2: return uns ? fprintf(f, "%u", (unsigned)*(const unsigned short *)p) : fprintf(f, "%d", (int)*(const short *)p);
        case 4: return uns ? fprintf(f, "%u", *(const unsigned int *)p) : fprintf(f, "%d", *(const int *)p);
        case 8: return uns ? fprintf(f, "%llu", *(const unsigned long long *)p) : fprintf(f, "%lld", *(const long long *)p);
    }
    return -1;
}

bool constant_is_float(const fgpu_constant &c) { return strstr(c.ctype, "float") || strstr(c.ctype, "double"); }

/* all state lists of one agent type on the host: cols[state][var] */
int download_states(fgpu_sim *s, int a, std::vector<std::vector<std::vector<char>>> &cols) {
    AgentRT &A = s->agents[a];
    cols.assign(A.d->nstates, std::vector<std::vector<char>>(A.d->nvars));
    for (int st = 0; st < A.d->nstates; ++st)
        for (int v = 0; v < A.d->nvars; ++v) {
            const fgpu_var &V = A.d->vars[v];
            cols[st][v].resize((size_t)A.h_state_count[st] * V.size);
            if (!cols[st][v].empty())
                CU(cudaMemcpyAsync(cols[st][v].data(), A.state[st] + V.list_offset, cols[st][v].size(), cudaMemcpyDeviceToHost, s->stream));
        }
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

const unsigned int SNAP_MAGIC = 0x50534746u; /* "FGSP" */
const unsigned int SNAP_VERSION = 1;

bool put(FILE *f, const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool get(FILE *f, void *p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }
bool put_u32(FILE *f, unsigned int v) { return put(f, &v, 4); }
bool get_u32(FILE *f, unsigned int *v) { return get(f, v, 4); }
bool put_str(FILE *f, const char *t) {
    const unsigned int n = (unsigned int)strlen(t);
    return put_u32(f, n) && put(f, t, n);
}
bool get_str(FILE *f, std::string *t) {
    unsigned int n = 0;
    if (!get_u32(f, &n) || n > 4096) return false;
    t->resize(n);
    return get(f, &(*t)[0], n);
}
}  // namespace

extern "C" int fgpu_sim_save_states(fgpu_sim *s, const char *outputpath, int iteration_number) {
    if (!s || !outputpath) return set_err(FGPU_ERR_ARG, "bad argument");
    const fgpu_model *m = s->model;
    CU(cudaSetDevice(s->device));
    int rc = slab_settle(s);
    if (rc) return rc;
    const std::string path = std::string(outputpath) + std::to_string(iteration_number) + ".xml"; /* io.xslt:143 */
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return set_err(FGPU_ERR_IO, "Error: Could not open file `%s` for output. Aborting.", path.c_str());
    fprintf(f, "<states>\n<itno>%i</itno>\n<environment>\n", iteration_number);
    for (int c = 0; c < m->nconstants; ++c) {
        const fgpu_constant &K = m->constants[c];
        const char *v = (const char *)K.get();
        fprintf(f, "\t<%s>", K.name);
        for (int j = 0; j < K.length; ++j) {
            write_value(f, K.ctype, K.size, constant_is_float(K), v + (size_t)j * K.size);
            if (j != K.length - 1) fputc(',', f);
        }
        fprintf(f, "</%s>\n", K.name);
    }
    fputs("</environment>\n", f);
    for (int a = 0; a < m->nagents; ++a) {
        std::vector<std::vector<std::vector<char>>> cols;
        rc = download_states(s, a, cols);
        if (rc) {
            fclose(f);
            return rc;
        }
        const AgentRT &A = s->agents[a];
        for (int st = 0; st < A.d->nstates; ++st)
            for (int i = 0; i < A.h_state_count[st]; ++i) {
                fprintf(f, "<xagent>\n<name>%s</name>\n", A.d->name);
                for (int v = 0; v < A.d->nvars; ++v) {
                    const fgpu_var &V = A.d->vars[v];
                    fprintf(f, "<%s>", V.name);
                    write_value(f, V.ctype, V.size, V.is_float != 0, cols[st][v].data() + (size_t)i * V.size);
                    fprintf(f, "</%s>\n", V.name);
                }
                fputs("</xagent>\n", f);
            }
    }
    fputs("</states>\n", f);
    if (fclose(f) != 0) return set_err(FGPU_ERR_IO, "write error on %s", path.c_str());
    return FGPU_OK;
}

extern "C" int fgpu_sim_save_snapshot(fgpu_sim *s, const char *path) {
    if (!s || !path) return set_err(FGPU_ERR_ARG, "bad argument");
    const fgpu_model *m = s->model;
    CU(cudaSetDevice(s->device));
    int rc = slab_settle(s);
    if (rc) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) return set_err(FGPU_ERR_IO, "could not open %s for writing", path);
    bool ok = put_u32(f, SNAP_MAGIC) && put_u32(f, SNAP_VERSION) && put_str(f, m->name) && put_u32(f, s->iteration) &&
              put_u32(f, (unsigned int)m->nagents);
    for (int a = 0; ok && a < m->nagents; ++a) {
        std::vector<std::vector<std::vector<char>>> cols;
        rc = download_states(s, a, cols);
        if (rc) {
            fclose(f);
            return rc;
        }
        const AgentRT &A = s->agents[a];
        ok = put_str(f, A.d->name) && put_u32(f, (unsigned int)A.d->nvars) && put_u32(f, (unsigned int)A.d->nstates);
        for (int v = 0; ok && v < A.d->nvars; ++v) ok = put_str(f, A.d->vars[v].name) && put_u32(f, (unsigned int)A.d->vars[v].size);
        for (int st = 0; ok && st < A.d->nstates; ++st) {
            ok = put_str(f, A.d->states[st]) && put_u32(f, (unsigned int)A.h_state_count[st]);
            for (int v = 0; ok && v < A.d->nvars; ++v) ok = put(f, cols[st][v].data(), cols[st][v].size());
        }
    }
    if (ok) {
        std::vector<uint2> seeds((size_t)m->buffer_size_max);
        CU(cudaMemcpy(seeds.data(), s->seeds, sizeof(uint2) * seeds.size(), cudaMemcpyDeviceToHost));
        ok = put_u32(f, (unsigned int)m->buffer_size_max) && put(f, seeds.data(), sizeof(uint2) * seeds.size());
    }
    ok = ok && put_u32(f, (unsigned int)m->nconstants);
    for (int c = 0; ok && c < m->nconstants; ++c) {
        const fgpu_constant &K = m->constants[c];
        ok = put_str(f, K.name) && put_u32(f, (unsigned int)(K.size * K.length)) && put(f, K.get(), (size_t)K.size * K.length);
    }
    if (fclose(f) != 0 || !ok) return set_err(FGPU_ERR_IO, "write error on %s", path);
    return FGPU_OK;
}

extern "C" int fgpu_sim_load_snapshot(fgpu_sim *s, const char *path) {
    if (!s || !path) return set_err(FGPU_ERR_ARG, "bad argument");
    const fgpu_model *m = s->model;
    if (s->slab.on) return set_err(FGPU_ERR_ARG, "snapshots are per simulation: load before fgpu_sim_slab_init");
    CU(cudaSetDevice(s->device));
    FILE *f = fopen(path, "rb");
    if (!f) return set_err(FGPU_ERR_IO, "Could not open input file %s", path);
    auto fail = [&](const char *why) {
        fclose(f);
        return set_err(FGPU_ERR_IO, "%s: %s", path, why);
    };
    unsigned int magic = 0, version = 0, iteration = 0, nagents = 0;
    std::string name;
    if (!get_u32(f, &magic) || magic != SNAP_MAGIC || !get_u32(f, &version) || version != SNAP_VERSION) return fail("not a snapshot of this format");
    if (!get_str(f, &name) || name != m->name) return fail("snapshot of another model");
    if (!get_u32(f, &iteration) || !get_u32(f, &nagents) || (int)nagents != m->nagents) return fail("agent types differ");
    for (int a = 0; a < m->nagents; ++a) {
        const AgentRT &A = s->agents[a];
        unsigned int nvars = 0, nstates = 0;
        if (!get_str(f, &name) || name != A.d->name || !get_u32(f, &nvars) || !get_u32(f, &nstates) || (int)nvars != A.d->nvars ||
            (int)nstates != A.d->nstates)
            return fail("agent layout differs");
        for (int v = 0; v < A.d->nvars; ++v) {
            unsigned int size = 0;
            if (!get_str(f, &name) || name != A.d->vars[v].name || !get_u32(f, &size) || (int)size != A.d->vars[v].size)
                return fail("variable layout differs");
        }
        for (int st = 0; st < A.d->nstates; ++st) {
            unsigned int count = 0;
            if (!get_str(f, &name) || name != A.d->states[st] || !get_u32(f, &count)) return fail("state layout differs");
            if ((int)count > A.d->buffer_size) return fail("population exceeds the buffer size of this build");
            std::vector<std::vector<char>> cols(A.d->nvars);
            std::vector<const void *> ptrs(A.d->nvars);
            for (int v = 0; v < A.d->nvars; ++v) {
                cols[v].resize((size_t)count * A.d->vars[v].size + 1);
                if (!get(f, cols[v].data(), (size_t)count * A.d->vars[v].size)) return fail("truncated");
                ptrs[v] = cols[v].data();
            }
            const int rc = upload_columns(s, a, st, 0, (int)count, ptrs.data());
            if (rc) {
                fclose(f);
                return rc;
            }
        }
    }
    unsigned int nseeds = 0, nconst = 0;
    if (!get_u32(f, &nseeds) || (int)nseeds != m->buffer_size_max) return fail("rand48 state of another buffer size");
    std::vector<uint2> seeds(nseeds);
    if (!get(f, seeds.data(), sizeof(uint2) * seeds.size())) return fail("truncated");
    CU(cudaMemcpy(s->seeds, seeds.data(), sizeof(uint2) * seeds.size(), cudaMemcpyHostToDevice));
    if (!get_u32(f, &nconst) || (int)nconst != m->nconstants) return fail("environment constants differ");
    for (int c = 0; c < m->nconstants; ++c) {
        const fgpu_constant &K = m->constants[c];
        unsigned int bytes = 0;
        if (!get_str(f, &name) || name != K.name || !get_u32(f, &bytes) || (int)bytes != K.size * K.length) return fail("environment constants differ");
        std::vector<char> v(bytes + 1);
        if (!get(f, v.data(), bytes)) return fail("truncated");
        K.set(v.data());
    }
    fclose(f);
    s->iteration = iteration;
    invalidate_graph(s);
    return FGPU_OK;
}

namespace {
template <typename T>
int reduce_typed(fgpu_sim *s, const char *col, int n, int op, const void *count_value, void *result) {
    if (s->unpack_tmp_n < (size_t)RED_MAX_BLOCKS * 2 + 4) {
        int rc = ensure_tmp(s, (size_t)RED_MAX_BLOCKS * 2 + 4);
        if (rc) return rc;
    }
    T *partial = reinterpret_cast<T *>(s->unpack_tmp); /* 8-byte elements fit: 2 words each */
    T *out = partial + RED_MAX_BLOCKS;
    const T *c = reinterpret_cast<const T *>(col);
    const int blocks = std::max(1, std::min(RED_MAX_BLOCKS, (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4)));
    T cv = (T)0;
    if (op == RED_COUNT) memcpy(&cv, count_value, sizeof(T));
    switch (op) {
#define RED_CASE(OP)                                                                            \
    case OP:                                                                                    \
        k_reduce_partial<T, OP><<<blocks, RED_THREADS, 0, s->stream>>>(c, n, cv, partial);      \
        k_reduce_final<T, (OP == RED_COUNT ? RED_SUM : OP)><<<1, RED_THREADS, 0, s->stream>>>(partial, blocks, out); \
        break;
        RED_CASE(RED_SUM) RED_CASE(RED_MIN) RED_CASE(RED_MAX) RED_CASE(RED_COUNT)
#undef RED_CASE
        default:
            return set_err(FGPU_ERR_ARG, "unknown reduction %d", op);
    }
    s->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(FGPU_ERR_CUDA, "reduction launch -> %s", cudaGetErrorString(e));
    CU(cudaMemcpyAsync(result, out, sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}
}  // namespace

extern "C" int fgpu_sim_reduce_agent_var(fgpu_sim *s, int agent, int state, int var, int op, const void *count_value, void *result) {
    if (!s || !result || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates || var < 0 || var >= A.d->nvars) return set_err(FGPU_ERR_ARG, "bad index");
    if (op == FGPU_REDUCE_COUNT && !count_value) return set_err(FGPU_ERR_ARG, "count needs a value");
    CU(cudaSetDevice(s->device));
    int rc = slab_settle(s);
    if (rc) return rc;
    const fgpu_var &V = A.d->vars[var];
    const int n = A.h_state_count[state];
    if (op == FGPU_REDUCE_COUNT && V.is_float) return set_err(FGPU_ERR_ARG, "count_ is defined for integer variables only (header.xslt:735)");
    if (n == 0) {
        if (op == FGPU_REDUCE_MIN || op == FGPU_REDUCE_MAX) return set_err(FGPU_ERR_ARG, "min/max of an empty %s list", A.d->name);
        memset(result, 0, (size_t)V.size);
        return FGPU_OK;
    }
    const char *col = A.state[state] + V.list_offset;
    const bool uns = strstr(V.ctype, "unsigned") != nullptr;
    if (V.is_float) return V.size == 8 ? reduce_typed<double>(s, col, n, op, count_value, result) : reduce_typed<float>(s, col, n, op, count_value, result);
    if (V.size == 8) return uns ? reduce_typed<unsigned long long>(s, col, n, op, count_value, result) : reduce_typed<long long>(s, col, n, op, count_value, result);
    if (V.size == 4) return uns ? reduce_typed<unsigned int>(s, col, n, op, count_value, result) : reduce_typed<int>(s, col, n, op, count_value, result);
    return set_err(FGPU_ERR_MODEL, "reduction over a %d-byte variable is not supported", V.size);
}

extern "C" int fgpu_sim_set_option(fgpu_sim *s, const char *key, int value) {
    if (!s || !key) return set_err(FGPU_ERR_ARG, "bad argument");
    if (!strcmp(key, "order")) s->opt_order = value ? 1 : 0;
    else if (!strcmp(key, "nvtx")) s->opt_nvtx = value ? 1 : 0;
    else if (!strcmp(key, "graph")) s->opt_graph = value ? 1 : 0;
    else if (!strcmp(key, "slab_p2p")) {
        if (s->slab.on) return set_err(FGPU_ERR_ARG, "slab_p2p must be set before fgpu_sim_slab_init");
        s->opt_slab_p2p = value ? 1 : 0;
    } else if (!strcmp(key, "slab_async")) {
        if (s->slab.on) return set_err(FGPU_ERR_ARG, "slab_async must be set before fgpu_sim_slab_init");
        s->opt_slab_async = value ? 1 : 0;
    } else if (!strcmp(key, "radix_bits")) {
        if (value < 4 || value > SORT_MAXBITS) return set_err(FGPU_ERR_ARG, "radix_bits must be in [4,%d]", SORT_MAXBITS);
        for (MsgRT &M : s->msgs)
            if (M.d->partitioning == FGPU_PART_SPATIAL && sort_passes(M.d->radix_bits, value) > SORT_MAX_PASSES)
                return set_err(FGPU_ERR_ARG, "radix_bits %d needs more than %d passes for message %s", value, SORT_MAX_PASSES, M.d->name);
        s->opt_radix_bits = value;
    } else
        return set_err(FGPU_ERR_ARG, "unknown option '%s'", key);
    invalidate_graph(s);
    return FGPU_OK;
}

/* `iters` iterations in stream mode with a cudaEvent pair around every stage
 * (INSTRUMENT_AGENT_FUNCTIONS, simulation.xslt:898-914, at stage granularity).  The iterations are
 * enqueued back to back without host synchronisation and the events come from a reused pool, so the
 * reported times are device times of the steady state; stages are averaged by name. */
extern "C" int fgpu_sim_profile_iterations(fgpu_sim *s, int iters, const char **names, float *ms, int cap) {
    if (!s || iters < 1) return set_err(FGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    s->prof.clear();
    s->ev_used = 0;
    s->profiling = true;
    int rc = FGPU_OK;
    for (int i = 0; i < iters && rc == FGPU_OK; ++i) rc = single_iteration(s);
    s->profiling = false;
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    rc = slab_settle(s);
    if (rc) return rc;
    s->prof_names.clear();
    s->prof_ms.clear();
    for (ProfEntry &e : s->prof) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        size_t k = 0;
        while (k < s->prof_names.size() && s->prof_names[k] != e.name) ++k;
        if (k == s->prof_names.size()) {
            s->prof_names.push_back(e.name);
            s->prof_ms.push_back(0.f);
        }
        s->prof_ms[k] += t / iters;
    }
    const int n = std::min(cap, (int)s->prof_names.size());
    for (int i = 0; i < n; ++i) {
        if (names) names[i] = s->prof_names[i].c_str();
        if (ms) ms[i] = s->prof_ms[i];
    }
    return n;
}

extern "C" int fgpu_sim_profile_iteration(fgpu_sim *s, const char **names, float *ms, int cap) {
    return fgpu_sim_profile_iterations(s, 1, names, ms, cap);
}

/* ------------------------------------------------------------------------- */
/* kernel-level entry points used by the parity tests                         */
/* ------------------------------------------------------------------------- */
extern "C" int fgpu_debug_sort_pairs(const unsigned int *keys, int n, int key_bits, int max_digit_bits,
                                     unsigned int *keys_out, unsigned int *vals_out) {
    if (n < 0 || key_bits < 1 || key_bits > 32) return set_err(FGPU_ERR_ARG, "bad argument");
    if (n == 0) return FGPU_OK;
    const int max_bits = std::min(std::max(max_digit_bits, 1), SORT_MAXBITS);
    if (sort_passes(key_bits, max_bits) > SORT_MAX_PASSES) return set_err(FGPU_ERR_ARG, "too many passes");
    unsigned int *kin, *k[2], *v[2], *scratch;
    const size_t sb = sizeof(unsigned int) * sort_scratch_words(n, key_bits, max_bits);
    CU(cudaMalloc(&kin, sizeof(unsigned int) * (size_t)n));
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&k[i], sizeof(unsigned int) * (size_t)n));
        CU(cudaMalloc(&v[i], sizeof(unsigned int) * (size_t)n));
    }
    CU(cudaMalloc(&scratch, sb));
    CU(cudaMemcpy(kin, keys, sizeof(unsigned int) * (size_t)n, cudaMemcpyHostToDevice));
    int in = 0;
    long long launches = 0;
    int rc = radix_sort_pairs(nullptr, kin, n, key_bits, max_bits, k, v, scratch, &in, &launches);
    if (rc == FGPU_OK) {
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(keys_out, k[in], sizeof(unsigned int) * (size_t)n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(vals_out, v[in], sizeof(unsigned int) * (size_t)n, cudaMemcpyDeviceToHost));
    }
    cudaFree(kin);
    for (int i = 0; i < 2; ++i) {
        cudaFree(k[i]);
        cudaFree(v[i]);
    }
    cudaFree(scratch);
    return rc;
}

/* flags -> (positions of survivors, count); columns = one 4-byte payload column */
extern "C" int fgpu_debug_compact(const int *flags, const unsigned int *payload, int n, unsigned int *out, int *count_out) {
    if (n < 0 || !count_out) return set_err(FGPU_ERR_ARG, "bad argument");
    *count_out = 0;
    if (n == 0) return FGPU_OK;
    int *d_flags, *d_total;
    unsigned int *d_in, *d_out, *d_bs;
    const int nb = (n + COMPACT_THREADS - 1) / COMPACT_THREADS;
    CU(cudaMalloc(&d_flags, sizeof(int) * (size_t)n));
    CU(cudaMalloc(&d_in, sizeof(unsigned int) * (size_t)n));
    CU(cudaMalloc(&d_out, sizeof(unsigned int) * (size_t)n));
    CU(cudaMalloc(&d_bs, sizeof(unsigned int) * (size_t)(nb + 1)));
    CU(cudaMalloc(&d_total, sizeof(int)));
    CU(cudaMemcpy(d_flags, flags, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_in, payload, sizeof(unsigned int) * (size_t)n, cudaMemcpyHostToDevice));
    ColSet cols;
    cols.ncols = 1;
    cols.size[0] = 4;
    cols.off[0] = 0;
    k_flag_blocksum<<<nb, COMPACT_THREADS>>>(d_flags, n, nullptr, d_bs, 1);
    k_scan_blocksums<<<1, 1024>>>(d_bs, nb, d_total);
    k_compact_scatter<<<nb, COMPACT_THREADS>>>((const char *)d_in, (char *)d_out, cols, cols, d_flags, d_bs, 0, n, nullptr, 1, 0x7fffffff, nullptr);
    CU(cudaGetLastError());
    CU(cudaMemcpy(count_out, d_total, sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out, d_out, sizeof(unsigned int) * (size_t)(*count_out), cudaMemcpyDeviceToHost));
    cudaFree(d_flags);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_bs);
    cudaFree(d_total);
    return FGPU_OK;
}
