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
 *
 * One translation unit, in parts (textual includes, in this order):
 *   fgpu_rt.cu             runtime objects, plugins, create/destroy, upload/download, run_function, stepping, options
 *   fgpu_rt_reader.inl     0.xml reader
 *   fgpu_rt_build.inl      compaction, radix sort driver, spatial message build
 *   fgpu_rt_slab.inl       multi-GPU slabs: transport, halo exchange, migration
 *   fgpu_rt_replay.inl     slab mode: replay of captured iterations
 *   fgpu_rt_state_io.inl   iteration-file writer, binary snapshot, host analytics reductions
 */
#include <cuda_runtime.h>
#include <charconv>
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

/* driver state of the reference's main.cu that step functions may use (process-wide there too) */
static std::string g_output_dir;
static int g_exit_early = 0;
extern "C" void fgpu_set_output_dir(const char *path) { g_output_dir = path ? path : ""; }
extern "C" const char *fgpu_get_output_dir(void) { return g_output_dir.c_str(); }
extern "C" void fgpu_set_exit_early(int flag) { g_exit_early = flag ? 1 : 0; }
extern "C" int fgpu_get_exit_early(void) { return g_exit_early; }

/* Kernel launch with the programmatic-dependent-launch attribute (see fgpu::pdl_wait): every kernel launched this way
 * starts with pdl_wait().  Captured into a CUDA graph the attribute becomes a programmatic dependency edge. */
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

/* ------------------------------------------------------------------------- */
/* runtime objects                                                            */
/* ------------------------------------------------------------------------- */
/* number of elements of an agent variable: 1 for a scalar, arrayLength for an array (fgpu_var.length) */
static inline int var_len(const fgpu_var &v) { return v.length > 1 ? v.length : 1; }

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
    /* sort_<A>s_<S> (allocated on first use): the caller's pairs, the radix ping-pong, its scratch */
    unsigned int *sort_keys = nullptr, *sort_vals = nullptr, *sort_k[2] = {nullptr, nullptr}, *sort_v[2] = {nullptr, nullptr}, *sort_scratch = nullptr;
};

struct MsgRT {
    const fgpu_message *d = nullptr;
    uint4 *unsorted = nullptr; /* staging written by add_<M>_message */
    uint4 *sorted = nullptr;   /* cell-ordered list read by get_first/next */
    uint4 *pay_a = nullptr;    /* MSD build: the records after the pass on the high digit (bucket-major) */
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
    /* carried agent records (fgpu_launch.carry_out / carry_in): written by the producer in list order, moved to cell
     * order by the build together with the message records */
    uint4 *carry_unsorted = nullptr, *carry_sorted = nullptr;
    int carry_quads = 0;          /* record size of the buffers above (allocated for the first producer's agent type) */
    bool halo_pending = false;    /* slab mode: the halo exchange of this build runs on the side stream; join before reading halo cells */
    bool carry_pending = false;   /* the producer of this iteration wrote carry_unsorted */
    bool carry_valid = false;     /* carry_sorted matches the sorted list */
    unsigned long long carry_content = 0; /* content epoch of the producer list when the records were written */
    int producer_fn = -1;                 /* the function that produced the list (single producer), for the own-message shortcut */
    unsigned long long producer_content = 0; /* content epoch of its list right after it ran */
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
    unsigned int *holes = nullptr; /* in-place migration: list slots of the leavers, ascending (2 * mig_cap per state) */
    int *mig_hdr = nullptr;        /* per state: {n_old, holes, -, -} */
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
    std::map<const char *, unsigned long long> content; /* bumped by every function launched on a list buffer */
    std::vector<char> emits_carry; /* per function: its message has a cell-order consumer right behind it (static) */
    int opt_msd = 0;   /* spatial message build: 1 = MSD pass + bucket finish where the key width allows; 0 = LSD + gather.
                          Measured slower at both 1M and 16M messages (wide-digit scatter), profiles/r02_notes.md: off */
    int opt_carry = 0; /* off by default: measured at 16M agents it trades 0.10 ms of the consumer for 0.55 ms of build (profiles/r02_notes.md) */
    /* options */
    int opt_order = 1;
    int opt_nvtx = 0;
    int opt_graph = 1;
    int opt_radix_bits = 9;
    int opt_slab_async = 1;
    int opt_slab_p2p = 1;
    int opt_slab_overlap = 0; /* 1: halo exchange on a side stream while the agents of the interior planes are processed.
                                 Off: measured no gain at 2 ranks (1.601 vs 1.592 ms; the exchange is 32 us of the step) */
    cudaStream_t stream2 = nullptr;
    int side_priority = 0;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int opt_pdl = 1;          /* programmatic dependent launch between the kernels of an iteration */
    int opt_slab_direct = 0;  /* slab mode, peer-memory transport: 1 = migrants are packed straight into the neighbours' buffers (no k_push launch).
                                 Off: measured slower at 2 ranks (1.445 vs 1.411 ms per iteration of C2-16M): the pack's scattered 4-byte stores cross
                                 NVLink one by one (pack 27 -> 75 us), which costs more than the saved launch and bulk copy (19 us) */
    int opt_mirror = 1;       /* a cell-order consumer reads the variables its own message carries from that message (fgpu_launch.mirror) */
    int opt_slab_inplace = 1; /* migration: 1 = leavers leave holes that arrivals / tail agents fill; 0 = stable three-way re-compaction */
    /* the reference's compile-time -DINSTRUMENT_* switches (simulation.xslt:370-377), as run-time options: each prints
     * the reference's own "Instrumentation: ... = %f (ms)" lines */
    int opt_instr_iterations = 0, opt_instr_agent_functions = 0, opt_instr_init_functions = 0, opt_instr_step_functions = 0,
        opt_instr_exit_functions = 0;
    cudaEvent_t instr_a = nullptr, instr_b = nullptr, instr_it_a = nullptr, instr_it_b = nullptr;
    int opt_xml_digits = 0; /* 0: the reference's "%f" iteration files; n: "%.<n>g" */
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
static int join_pending_halos(fgpu_sim *s);
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
constexpr int SCR_GAP = 0, SCR_TOTALS = 8, SCR_TILEHIST = 8 + SORT_MAX_BINS; /* scratch layout, 32-bit words */

/* Buffers of the carried agent records (option "carry"): allocated with the option, never inside a captured iteration. */
static int alloc_carry(fgpu_sim *s) {
    const fgpu_model *m = s->model;
    for (int f = 0; f < m->nfunctions; ++f) {
        if (!s->emits_carry[f]) continue;
        const fgpu_function &P = m->functions[f];
        MsgRT &M = s->msgs[P.output_message];
        if (M.carry_unsorted) continue;
        M.carry_quads = m->agents[P.agent].carry_quads;
        const size_t bytes = (size_t)M.carry_quads * 16 * (size_t)M.d->buffer_size;
        CU(cudaMalloc(&M.carry_unsorted, bytes));
        CU(cudaMalloc(&M.carry_sorted, bytes));
    }
    return FGPU_OK;
}

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
        A.cols.stride = (unsigned int)A.d->buffer_size;
        for (int v = 0; v < A.d->nvars; ++v) {
            A.cols.size[v] = A.d->vars[v].size;
            A.cols.len[v] = (unsigned short)var_len(A.d->vars[v]);
            A.cols.off[v] = A.d->vars[v].list_offset;
            if (var_len(A.d->vars[v]) > 65535) return set_err(FGPU_ERR_MODEL, "agent %s: array variable %s is longer than 65535", A.d->name, A.d->vars[v].name);
        }
        const int nb = (A.d->buffer_size + COMPACT_THREADS - 1) / COMPACT_THREADS;
        CU(cudaMalloc(&A.block_sums, sizeof(unsigned int) * 3 * (size_t)(nb + 1)));
        CU(cudaMalloc(&A.d_total, sizeof(int) * 2)); /* [1]: second count of a function with both death and birth */
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
            CU(cudaMalloc(&M.pay_a, msg_plane_bytes(M.d)));
            CU(cudaMalloc(&M.keys_in, sizeof(unsigned int) * (size_t)M.d->buffer_size));
            for (int k = 0; k < 2; ++k) {
                CU(cudaMalloc(&M.keys[k], sizeof(unsigned int) * (size_t)M.d->buffer_size));
                CU(cudaMalloc(&M.vals[k], sizeof(unsigned int) * (size_t)M.d->buffer_size));
            }
            CU(cudaMalloc(&M.cell_start, sizeof(unsigned int) * ((size_t)M.d->grid_size + 1)));
            CU(cudaMemsetAsync(M.cell_start, 0, sizeof(unsigned int) * ((size_t)M.d->grid_size + 1), s->stream));
            M.max_tiles = (M.d->buffer_size + SORT_TILE - 1) / SORT_TILE;
            /* scratch: gap_count | totals[2048] | tile_hist[bins][tiles], sized for the widest digit any key width can
             * select (slab mode sorts on a narrower, slab-relative key, which may choose wider digits than the full key) */
            const size_t words = (size_t)SCR_TILEHIST + (size_t)M.max_tiles * SORT_MAX_BINS;
            M.scratch_bytes = sizeof(unsigned int) * words;
            CU(cudaMalloc(&M.scratch, M.scratch_bytes));
            CU(cudaMalloc(&M.gap_queue, sizeof(unsigned int) * 3 * ((size_t)GAP_LONG_MAX + (size_t)M.d->grid_size / GAP_INLINE + 8)));
        }
    }
    s->gc_count.assign(m->nfunctions, 0);
    /* A producer writes carried agent records when the next function executed on the same agent type consumes its
     * spatial message from the very list the producer worked on (in place, no condition): that consumer runs in
     * cell order and would otherwise gather every variable it reads with random 4-byte loads. */
    s->emits_carry.assign(m->nfunctions, 0);
    {
        std::vector<int> seq;
        for (int l = 0; l < m->nlayers; ++l)
            for (int k = 0; k < m->layers[l].nfunctions; ++k) seq.push_back(m->layers[l].functions[k]);
        for (size_t i = 0; i < seq.size(); ++i) {
            const fgpu_function &P = m->functions[seq[i]];
            if (P.output_message < 0 || P.output_type != FGPU_OUT_SINGLE || !P.swappable) continue;
            if (m->messages[P.output_message].partitioning != FGPU_PART_SPATIAL || m->agents[P.agent].carry_quads <= 0) continue;
            bool other_producer = false;
            for (size_t j = 0; j < seq.size(); ++j)
                if (j != i && m->functions[seq[j]].output_message == P.output_message) other_producer = true;
            if (other_producer) continue;
            for (size_t j = i + 1; j < seq.size(); ++j) {
                const fgpu_function &G = m->functions[seq[j]];
                if (G.agent != P.agent) continue;
                if (G.input_message == P.output_message && G.condition == FGPU_COND_NONE && G.current_state == P.next_state)
                    s->emits_carry[seq[i]] = 1;
                break; /* only the next function on this agent type sees the variables unchanged */
            }
        }
    }
    if (s->opt_carry) {
        int rc = alloc_carry(s);
        if (rc) return rc;
    }

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
        cudaFree(A.sort_keys);
        cudaFree(A.sort_vals);
        for (int k = 0; k < 2; ++k) {
            cudaFree(A.sort_k[k]);
            cudaFree(A.sort_v[k]);
        }
        cudaFree(A.sort_scratch);
    }
    for (MsgRT &M : s->msgs) {
        cudaFree(M.unsorted);
        cudaFree(M.sorted);
        cudaFree(M.pay_a);
        cudaFree(M.keys_in);
        cudaFree(M.opt_recs);
        cudaFree(M.opt_keys);
        cudaFree(M.opt_flags);
        cudaFree(M.carry_unsorted);
        cudaFree(M.carry_sorted);
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
        cudaFree(X.holes);
        cudaFree(X.mig_hdr);
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
    for (cudaEvent_t e : {s->instr_a, s->instr_b, s->instr_it_a, s->instr_it_b})
        if (e) cudaEventDestroy(e);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->stream2) cudaStreamDestroy(s->stream2);
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

static int upload_columns(fgpu_sim *s, int agent, int state, int offset, int count, const void *const *columns, bool async = false) {
    if (agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad agent index %d", agent);
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return set_err(FGPU_ERR_ARG, "bad state index %d", state);
    if (count < 0 || offset + count > A.d->buffer_size)
        return set_err(FGPU_ERR_OVERFLOW, "ERROR: MAX Buffer size (%i) for agent %s exceeded whilst reading data",
                       A.d->buffer_size, A.d->name); /* io.xslt:406, without its off-by-one */
    CU(cudaSetDevice(s->device));
    char *list = A.state[state];
    for (int v = 0; v < A.d->nvars; ++v)
      for (int e = 0; e < var_len(A.d->vars[v]); ++e) {   /* an array variable arrives element-major: columns[v] = length blocks of `count` values */
        const fgpu_var &V = A.d->vars[v];
        char *dst = list + V.list_offset + ((size_t)e * A.d->buffer_size + (size_t)offset) * V.size;
        if (columns && columns[v]) {
            CU(cudaMemcpyAsync(dst, (const char *)columns[v] + (size_t)e * count * V.size, (size_t)count * V.size, cudaMemcpyHostToDevice, s->stream));
        } else {
            /* defaultValue or 0 (io.xslt:276-291), written on the device: no host buffer, no copy over PCIe */
            if (count > 0) {
                const int g = std::min(148 * 8, (count + 255) / 256);
                if (V.size == 8) {
                    unsigned long long bits;
                    if (V.is_float) {
                        const double d = V.default_value;
                        memcpy(&bits, &d, 8);
                    } else {
                        const long long l = (long long)V.default_value;
                        memcpy(&bits, &l, 8);
                    }
                    k_fill_column<unsigned long long><<<g, 256, 0, s->stream>>>((unsigned long long *)dst, bits, count);
                } else {
                    unsigned int bits;
                    if (V.is_float) {
                        const float f = (float)V.default_value;
                        memcpy(&bits, &f, 4);
                    } else if (!strcmp(V.ctype, "unsigned int")) {
                        bits = (unsigned int)V.default_value;
                    } else {
                        const int i = (int)V.default_value;
                        memcpy(&bits, &i, 4);
                    }
                    k_fill_column<unsigned int><<<g, 256, 0, s->stream>>>((unsigned int *)dst, bits, count);
                }
                CU(cudaGetLastError());
            }
        }
    }
    const bool same_count = (A.h_state_count[state] == offset + count);
    A.h_state_count[state] = offset + count;
    if (!same_count || s->slab.on) /* d_state_count already holds this value otherwise */
        CU(cudaMemcpyAsync(A.d_state_count + state, &A.h_state_count[state], sizeof(int), cudaMemcpyHostToDevice, s->stream));
    if (!async) CU(cudaStreamSynchronize(s->stream));
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

extern "C" int fgpu_sim_set_agents_async(fgpu_sim *s, int agent, int state, int count, const void *const *columns) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    if (s->slab.on) return set_err(FGPU_ERR_ARG, "fgpu_sim_set_agents_async is not available in slab mode");
    return upload_columns(s, agent, state, 0, count, columns, true);
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
    for (int e = 0; e < var_len(V); ++e)   /* an array variable: `length` blocks of `count` values, element-major */
        CU(cudaMemcpyAsync((char *)dst + (size_t)e * count * V.size, A.state[state] + V.list_offset + (size_t)e * A.d->buffer_size * V.size,
                           (size_t)count * V.size, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

static int read_agents(fgpu_sim *s, int agent, int state, int count, void *const *columns, bool async);
extern "C" int fgpu_sim_read_agents(fgpu_sim *s, int agent, int state, int count, void *const *columns) {
    return read_agents(s, agent, state, count, columns, false);
}
extern "C" int fgpu_sim_read_agents_async(fgpu_sim *s, int agent, int state, int count, void *const *columns) {
    return read_agents(s, agent, state, count, columns, true);
}
static int read_agents(fgpu_sim *s, int agent, int state, int count, void *const *columns, bool async) {
    if (!s || !columns || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return set_err(FGPU_ERR_ARG, "bad index");
    if (count < 0 || count > A.d->buffer_size) return set_err(FGPU_ERR_ARG, "bad count");
    CU(cudaSetDevice(s->device));
    for (int v = 0; v < A.d->nvars; ++v) {
        if (!columns[v]) continue;
        const fgpu_var &V = A.d->vars[v];
        for (int e = 0; e < var_len(V); ++e)
            CU(cudaMemcpyAsync((char *)columns[v] + (size_t)e * count * V.size, A.state[state] + V.list_offset + (size_t)e * A.d->buffer_size * V.size,
                               (size_t)count * V.size, cudaMemcpyDeviceToHost, s->stream));
    }
    if (!async) CU(cudaStreamSynchronize(s->stream));
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

extern "C" int fgpu_sim_read_agent_element(fgpu_sim *s, int agent, int state, int var, unsigned int index, unsigned int element, void *dst) {
    if (!s || !dst || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates || var < 0 || var >= A.d->nvars) return set_err(FGPU_ERR_ARG, "bad index");
    if (index >= (unsigned int)A.h_state_count[state]) return set_err(FGPU_ERR_ARG, "agent index %u out of range", index);
    const fgpu_var &V = A.d->vars[var];
    if (element >= (unsigned int)var_len(V)) return set_err(FGPU_ERR_ARG, "element %u of %s[%d] out of range", element, V.name, var_len(V));
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(dst, A.state[state] + V.list_offset + ((size_t)element * A.d->buffer_size + index) * V.size, V.size, cudaMemcpyDeviceToHost, s->stream));
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

#include "fgpu_rt_reader.inl"

#include "fgpu_rt_build.inl"

#include "fgpu_rt_slab.inl"

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
        /* The filter flags every agent in exactly one of the two lists, so ONE count (and one read-back) gives both
         * populations; and when every agent passes -- or none -- both compactions are the identity: the filter wrote
         * the passing agents at their own indices, in order (the reference scans and scatters both lists regardless,
         * sim:1459-1528). */
        const int n_all = A.h_count;
        int working = 0;
        int rc = compact(s, A, nullptr, nullptr, 0, n_all, (const int *)(A.work + A.d->scan_offset), false, &working);
        if (rc) return rc;
        const int kept = n_all - working;
        if (working == n_all) {
            s->touch(A.work);
        } else if (working > 0) {
            const int nb = (n_all + COMPACT_THREADS - 1) / COMPACT_THREADS;
            k_compact_scatter<<<nb, COMPACT_THREADS, 0, s->stream>>>(A.work, A.swap, A.cols, A.cols, (const int *)(A.work + A.d->scan_offset),
                                                                     A.block_sums, 0, n_all, nullptr, 1, 0x7fffffff, nullptr);
            LAUNCH_CHECK();
            std::swap(A.work, A.swap);
            s->touch(A.work);
        }
        if (kept > 0 && kept < n_all) {
            rc = compact(s, A, A.state[cur], A.swap, 0, n_all, (const int *)(A.state[cur] + A.d->scan_offset), true, nullptr);
            if (rc) return rc;
            std::swap(A.state[cur], A.swap);
        }
        A.h_state_count[cur] = kept;
        A.h_count = working;
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
            /* ... and the variables themselves in cell order, if the producer left them and nothing ran on the list since */
            if (s->opt_carry && Min->carry_valid && Min->carry_quads == A.d->carry_quads && Min->carry_content == s->content[A.work])
                L.carry_in = Min->carry_sorted;
            /* ... or from the agent's own message, when the producer passed them through and nothing ran on the list since */
            if (s->opt_mirror && F.mirror_producer >= 0 && Min->producer_fn == F.mirror_producer && Min->producer_content == s->content[A.work]) {
                L.mirror = 1;
                L.mirror_base = (unsigned int)Min->local_at;
            }
        }
    }
    if (Mout) {
        L.out_recs = Mout->unsorted;
        L.out_keys = Mout->keys_in;
        L.out_base = Mout->h_count; /* a second producer appends; the whole list is re-binned (sim:1836-1837) */
        if (s->opt_carry && s->emits_carry[fi] && !optional_out && Mout->h_count == 0) {
            /* buffers were allocated with the simulation (no allocation inside a captured iteration) */
            if (Mout->carry_unsorted && Mout->carry_quads == A.d->carry_quads)
            L.carry_out = Mout->carry_unsorted;
        }
        if (optional_out) {
            /* reset_<M>_swaps (sim:1666-1672), then every agent that calls add_<M>_message flags its own slot */
            CU(cudaMemsetAsync(Mout->opt_flags, 0, sizeof(int) * (size_t)std::max(A.h_count, 1), s->stream));
            L.out_recs = Mout->opt_recs;
            L.out_keys = Mout->opt_keys;
            L.out_flags = Mout->opt_flags;
            L.out_base = 0;
        }
    }
    L.pdl = s->opt_pdl;
    if (F.rng) {
        L.seeds = s->seeds;
        L.Ax = s->Ax;
        L.Ay = s->Ay;
        L.Cx = s->Cx;
        L.Cy = s->Cy;
    }
    s->prof_begin(std::string("func:") + F.name);
    if (Min && Min->halo_pending) {
        const SlabRT &S = s->slab;
        if (L.order && S.ppr > 2) {
            /* agents of the interior planes need no halo cell: run them while the exchange is in flight */
            L.win_mode = 1;
            L.win_cell_lo = (unsigned int)(S.p0 + 1) * (unsigned int)S.plane_cells;
            L.win_cell_hi = (unsigned int)(S.p1 - 1) * (unsigned int)S.plane_cells;
            L.win_base = (unsigned int)Min->local_at;
            L.launch_n = L.n;
            F.launch(&L, s->stream);
            LAUNCH_CHECK();
            L.win_mode = 2;
            L.launch_n = std::min(L.n, 2 * S.halo_cap); /* a plane never holds more than halo_cap messages (checked by the pack) */
        }
        CU(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
        Min->halo_pending = false;
    }
    F.launch(&L, s->stream);
    LAUNCH_CHECK();
    s->prof_end();
    s->content[A.work]++;
    if (Mout && L.carry_out) {
        Mout->carry_pending = true;
        Mout->carry_content = s->content[A.work];
    }

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
            Mout->producer_fn = fi;
            Mout->producer_content = s->content[A.work];
            Mout->d_count = slab_async ? L.d_n : nullptr;
        } else {
            Mout->producers++;
            if (slab_async)
                return set_err(FGPU_ERR_MODEL, "message %s has two producers in one iteration: not supported in asynchronous slab mode", Mout->d->name);
        }
        Mout->h_count += A.h_count;
    }

    const int pre_death_count = A.h_count;
    /* A function with death AND birth into its own agent type pays one count read-back, not two: the survivors'
     * scan + scatter and the newborns' scan are enqueued first, both totals come back with one synchronisation
     * (the reference blocks on two D2H copies per scan, e.g. sim:1476-1477). */
    const bool both = F.reallocate && F.xagent_output >= 0 && F.xagent_output == F.agent;
    if (F.reallocate) {
        /* sim:1763-1792 */
        s->prof_begin(std::string("death:") + F.name);
        int alive = 0;
        int rc = compact(s, A, A.work, A.swap, 0, A.h_count, (const int *)(A.work + A.d->scan_offset), true, both ? nullptr : &alive);
        if (rc) return rc;
        std::swap(A.work, A.swap);
        if (!both) A.h_count = alive;
        s->prof_end();
    }
    if (F.xagent_output >= 0) {
        /* sim:1794-1829 */
        AgentRT &B = s->agents[F.xagent_output];
        const int st = F.xagent_output_state;
        s->prof_begin(std::string("birth:") + F.name);
        int born = 0;
        if (both) {
            int rc = compact(s, B, nullptr, nullptr, 0, pre_death_count, (const int *)(B.newl + B.d->scan_offset), false, nullptr, 1, nullptr,
                             0x7fffffff, nullptr, nullptr, A.d_total + 1);
            if (rc) return rc;
            int two[2] = {0, 0};
            CU(cudaMemcpyAsync(two, A.d_total, sizeof(two), cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            A.h_count = two[0];
            born = two[1];
        } else {
            int rc = compact(s, B, nullptr, nullptr, 0, pre_death_count, (const int *)(B.newl + B.d->scan_offset), false, &born);
            if (rc) return rc;
        }
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
            const bool overlap = s->opt_slab_overlap && s->slab.p2p && !s->profiling && s->stream2 && Mout->producers == 1;
            if (overlap) {
                /* fork: pack + push + wait + merge on the side stream; the consumer joins before it touches halo cells */
                CU(cudaEventRecord(s->ev_fork, s->stream));
                CU(cudaStreamWaitEvent(s->stream2, s->ev_fork, 0));
                rc = slab_exchange_halo(s, F.output_message, s->stream2);
                if (rc) return rc;
                CU(cudaEventRecord(s->ev_join, s->stream2));
                Mout->halo_pending = true;
            } else {
                rc = slab_exchange_halo(s, F.output_message, s->stream);
                if (rc) return rc;
            }
        }
    }
    if (s->slab.on && (F.writes_position || F.xagent_output >= 0)) s->slab.dirty = true;
    /* move agents to next state, sim:1936-1972 */
    if (A.h_state_count[nxt] + A.h_count > A.d->buffer_size)
        return set_err(FGPU_ERR_OVERFLOW, "Error: Buffer size of %s agents in state %s will be exceeded moving working agents to next state in function %s",
                       F.name, A.d->states[nxt], F.name);
    if (F.swappable) {
        std::swap(A.work, A.state[cur]);
    } else if (A.h_count > 0 && A.h_state_count[nxt] == 0) {
        /* appending to an empty list is a pointer swap (same agents, same order) */
        std::swap(A.work, A.state[nxt]);
        s->touch(A.state[nxt]);
        s->touch(A.work);
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
        M.carry_pending = false;
        M.carry_valid = false;
    }
    return FGPU_OK;
}

/* joins a halo exchange nobody consumed (every fork must be joined before the iteration ends or is captured) */
static int join_pending_halos(fgpu_sim *s) {
    for (MsgRT &M : s->msgs)
        if (M.halo_pending) {
            CU(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
            M.halo_pending = false;
        }
    return FGPU_OK;
}

/* cudaEvent pair + the reference's line, simulation.xslt:898-914 */
static int instr_begin(fgpu_sim *s) {
    if (!s->instr_a) {
        CU(cudaEventCreate(&s->instr_a));
        CU(cudaEventCreate(&s->instr_b));
        CU(cudaEventCreate(&s->instr_it_a));
        CU(cudaEventCreate(&s->instr_it_b));
    }
    CU(cudaEventRecord(s->instr_a, s->stream));
    return FGPU_OK;
}
static int instr_end(fgpu_sim *s, const char *prefix, const char *name) {
    float ms = 0.f;
    CU(cudaEventRecord(s->instr_b, s->stream));
    CU(cudaEventSynchronize(s->instr_b));
    CU(cudaEventElapsedTime(&ms, s->instr_a, s->instr_b));
    printf("Instrumentation: %s%s = %f (ms)\n", prefix, name, ms);
    return FGPU_OK;
}

static int single_iteration(fgpu_sim *s) {
    const fgpu_model *m = s->model;
    const bool it_timer = s->opt_instr_iterations && !s->capturing;
    if (it_timer) {
        int rc = instr_begin(s);
        if (rc) return rc;
        CU(cudaEventRecord(s->instr_it_a, s->stream));
    }
    begin_iteration(s);
    for (int l = 0; l < m->nlayers; ++l)
        for (int k = 0; k < m->layers[l].nfunctions; ++k) {
            const int fi = m->layers[l].functions[k];
            const bool timed = s->opt_instr_agent_functions && !s->capturing;
            if (timed) {
                int rc = instr_begin(s);
                if (rc) return rc;
            }
            int rc = run_function(s, fi);
            if (rc) return rc;
            if (timed) { /* "<Agent>_<function>", simulation.xslt:913 */
                const std::string name = std::string(m->agents[m->functions[fi].agent].name) + "_" + m->functions[fi].name;
                rc = instr_end(s, "", name.c_str());
                if (rc) return rc;
            }
        }
    {
        int rc = join_pending_halos(s);
        if (rc) return rc;
    }
    if (s->slab.on && s->slab.dirty) {
        int rc = slab_migrate(s);
        if (rc) return rc;
    }
    if (!s->capturing && m->nstep_functions > 0) {
        CU(cudaStreamSynchronize(s->stream));
        if (m->ids_to_host) m->ids_to_host(); /* simulation.xslt:918-930 */
        for (int k = 0; k < m->nstep_functions; ++k) {
            if (m->bind) m->bind(s);
            if (s->opt_instr_step_functions) {
                int rc = instr_begin(s);
                if (rc) return rc;
            }
            m->step_functions[k]();
            if (s->opt_instr_step_functions) {
                int rc = instr_end(s, "", m->step_function_names ? m->step_function_names[k] : "step function");
                if (rc) return rc;
            }
            CU(cudaStreamSynchronize(s->stream));
        }
        if (m->ids_to_device) m->ids_to_device(); /* simulation.xslt:948-960 */
    }
    if (it_timer) {
        float ms = 0.f;
        CU(cudaEventRecord(s->instr_it_b, s->stream));
        CU(cudaEventSynchronize(s->instr_it_b));
        CU(cudaEventElapsedTime(&ms, s->instr_it_a, s->instr_it_b));
        printf("Instrumentation: Iteration Time = %f (ms)\n", ms); /* simulation.xslt:973 */
    }
    return FGPU_OK;
}

static std::vector<int> current_counts(fgpu_sim *s) {
    std::vector<int> c;
    for (AgentRT &A : s->agents)
        for (int v : A.h_state_count) c.push_back(v);
    c.push_back(s->opt_order);
    c.push_back(s->opt_radix_bits);
    c.push_back(s->opt_carry);
    c.push_back(s->opt_msd);
    c.push_back(s->opt_pdl);
    c.push_back(s->opt_mirror);
    return c;
}

#include "fgpu_rt_replay.inl"

static int step_async(fgpu_sim *s, int n) {
    if (n <= 0) return FGPU_OK;
    CU(cudaSetDevice(s->device));
    const bool instrumented = s->opt_instr_iterations || s->opt_instr_agent_functions;
    const bool use_graph = s->is_static && s->opt_graph && !s->profiling && !instrumented;
    CU(cudaEventRecord(s->ev0, s->stream));
    if (s->slab.on && s->slab.async && s->slab.p2p && s->opt_graph && !s->profiling && !instrumented && s->model->nstep_functions == 0) {
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
    for (int k = 0; k < s->model->ninit_functions; ++k) {
        if (s->opt_instr_init_functions) {
            int rc = instr_begin(s);
            if (rc) return rc;
        }
        s->model->init_functions[k]();
        if (s->opt_instr_init_functions) { /* simulation.xslt:714 */
            int rc = instr_end(s, "", s->model->init_function_names ? s->model->init_function_names[k] : "init function");
            if (rc) return rc;
        }
    }
    CU(cudaDeviceSynchronize());
    if (s->model->ids_to_device) s->model->ids_to_device(); /* ids handed out by init functions on the host */
    return FGPU_OK;
}

extern "C" int fgpu_sim_run_exit_functions(fgpu_sim *s) {
    if (!s) return set_err(FGPU_ERR_ARG, "null sim");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    if (s->model->bind) s->model->bind(s);
    if (s->model->ids_to_host) s->model->ids_to_host();
    for (int k = 0; k < s->model->nexit_functions; ++k) {
        if (s->opt_instr_exit_functions) {
            int rc = instr_begin(s);
            if (rc) return rc;
        }
        s->model->exit_functions[k]();
        if (s->opt_instr_exit_functions) { /* simulation.xslt:796 */
            int rc = instr_end(s, "", s->model->exit_function_names ? s->model->exit_function_names[k] : "exit function");
            if (rc) return rc;
        }
    }
    return FGPU_OK;
}

/* sort_<A>s_<S>(), simulation.xslt:750-776 */
extern "C" int fgpu_sim_sort_agents(fgpu_sim *s, int agent, int state, const void *generate_key_value_pairs) {
    if (!s || !generate_key_value_pairs || agent < 0 || agent >= (int)s->agents.size()) return set_err(FGPU_ERR_ARG, "bad argument");
    AgentRT &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return set_err(FGPU_ERR_ARG, "bad state index");
    if (s->slab.on) return set_err(FGPU_ERR_MODEL, "sort_%ss_%s is not available in slab mode", A.d->name, A.d->states[state]);
    CU(cudaSetDevice(s->device));
    const int n = A.h_state_count[state];
    if (n <= 0) return FGPU_OK;
    const int cap = (A.d->buffer_size + 1023) & ~1023; /* the caller's kernel has no bounds check: whole blocks may write */
    if (!A.sort_keys) {
        CU(cudaMalloc(&A.sort_keys, sizeof(unsigned int) * (size_t)cap));
        CU(cudaMalloc(&A.sort_vals, sizeof(unsigned int) * (size_t)cap));
        for (int k = 0; k < 2; ++k) {
            CU(cudaMalloc(&A.sort_k[k], sizeof(unsigned int) * (size_t)cap));
            CU(cudaMalloc(&A.sort_v[k], sizeof(unsigned int) * (size_t)cap));
        }
        CU(cudaMalloc(&A.sort_scratch, sizeof(unsigned int) * sort_scratch_words(A.d->buffer_size, 32, 8)));
    }
    unsigned int *keys = A.sort_keys, *values = A.sort_vals;
    void *list = A.state[state];
    void *args[3] = {&keys, &values, &list};
    const int block = 256;
    CU(cudaLaunchKernel(generate_key_value_pairs, dim3((n + block - 1) / block), dim3(block), args, 0, s->stream));
    s->launches++;
    int in = 0;
    int rc = radix_sort_pairs(s->stream, keys, n, 32, 8, A.sort_k, A.sort_v, A.sort_scratch, &in, &s->launches);
    if (rc) return rc;
    k_reorder_agents<<<(n + 255) / 256, 256, 0, s->stream>>>(A.state[state], A.swap, A.cols, A.sort_v[in], values, n);
    LAUNCH_CHECK();
    std::swap(A.state[state], A.swap);
    s->touch(A.state[state]);
    s->touch(A.swap);
    CU(cudaStreamSynchronize(s->stream));
    return FGPU_OK;
}

#include "fgpu_rt_state_io.inl"

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

/* device time of the LSD sort of n keys of key_bits (digits of at most max_digit_bits), averaged over reps: a tuning aid */
extern "C" int fgpu_debug_sort_time(int n, int key_bits, int max_digit_bits, int reps, float *ms_out) {
    if (n <= 0 || key_bits < 1 || key_bits > 32 || reps < 1 || !ms_out) return set_err(FGPU_ERR_ARG, "bad argument");
    const int max_bits = std::min(std::max(max_digit_bits, 1), SORT_MAXBITS);
    unsigned int *kin, *k[2], *v[2], *scratch;
    CU(cudaMalloc(&kin, sizeof(unsigned int) * (size_t)n));
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&k[i], sizeof(unsigned int) * (size_t)n));
        CU(cudaMalloc(&v[i], sizeof(unsigned int) * (size_t)n));
    }
    CU(cudaMalloc(&scratch, sizeof(unsigned int) * ((size_t)SCR_TILEHIST + (size_t)((n + SORT_TILE - 1) / SORT_TILE) * SORT_MAX_BINS)));
    std::vector<unsigned int> h((size_t)n);
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < n; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        h[i] = (unsigned int)(x >> 33) & (key_bits == 32 ? 0xffffffffu : ((1u << key_bits) - 1u));
    }
    CU(cudaMemcpy(kin, h.data(), sizeof(unsigned int) * (size_t)n, cudaMemcpyHostToDevice));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    int in = 0, rc = FGPU_OK;
    long long launches = 0;
    for (int r = 0; r < reps + 2 && rc == FGPU_OK; ++r) {
        if (r == 2) CU(cudaEventRecord(a, nullptr));
        rc = radix_sort_pairs(nullptr, kin, n, key_bits, max_bits, k, v, scratch, &in, &launches);
    }
    CU(cudaEventRecord(b, nullptr));
    CU(cudaEventSynchronize(b));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / reps;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
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
    cols.stride = 0;
    cols.size[0] = 4;
    cols.len[0] = 1;
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
