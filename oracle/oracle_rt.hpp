/*
 * oracle_rt.hpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A host-C++ restatement of the FLAME GPU 1.5 per-iteration step as generated
 * by the reference's templates, used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs as the checker and CPU
 * baseline.  It is never linked into, imported by or called from the product
 * path (ev_diffusion_b200/), which has no CPU fallback.
 *
 * The reference has no CPU path, no tests and no golden vectors (SURVEY.md 4,
 * 8c), and its generator (xsltproc) and its texture-reference CUDA cannot be
 * built by this image's toolchain as shipped.  What pins this restatement:
 *   - THE REFERENCE ITSELF, run here: oracle/build_ref.py expands the
 *     reference's own templates (oracle/xslt_subset.py stands in for xsltproc),
 *     rewrites only the <<<>>> launch syntax and compiles the generated
 *     simulation.cu / FLAMEGPU_kernals.cu / io.cu plus the model's functions.c
 *     against oracle/cuda_on_host/ (every CUDA thread run on the CPU) into
 *     oracle/_ref/.  tests/test_oracle_vs_reference.py steps that build and
 *     this oracle side by side -- twelve of the reference's example models
 *     from their shipped inputs, the authored fixtures (death, births,
 *     optional messages, conditions, globalCondition, host-side creation) and
 *     the Brownian workload -- and requires identical bytes in every variable
 *     of every state, identical counts and identical rand48 seeds;
 *   - the rand48 known-answer vectors derived from simulation.xslt:661-688 and
 *     cross-checked against glibc srand48(123)/lrand48();
 *   - partition dims / radix-bit vectors;
 *   - it executes the reference's OWN agent scripts (functions.c) unchanged.
 * One thing the reference leaves to the device: the order in which an agent
 * meets NON-partitioned messages starts at its block's own tile
 * (FLAMEGPU_kernals.xslt:445-520), so it depends on the block size the
 * occupancy API picks.  This oracle (and the CUDA path) use 0..N-1 for every
 * agent, the reference's order when one block holds the list.
 *
 * Every function cites the template lines it follows:
 *   sim = FLAMEGPU/templates/simulation.xslt, krn = FLAMEGPU_kernals.xslt,
 *   hdr = header.xslt, io = io.xslt.
 *
 * This file is included by the generated per-model glue (one translation
 * unit), after the host-mode header.h and the user's functions.c.
 */
#ifndef ORACLE_RT_HPP
#define ORACLE_RT_HPP

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fgpu_model.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- per-message hooks supplied by the generated glue --------------------- */
struct orc_message_hooks {
    size_t list_bytes;      /* sizeof(xmachine_message_<M>_list) */
    size_t pbm_bytes;       /* sizeof(xmachine_message_<M>_PBM) or 0 */
    size_t pbm_end_offset;  /* offsetof(PBM, end_or_count) */
    int *count_symbol;      /* d_message_<M>_count */
    int *output_type_symbol;
};
extern const orc_message_hooks orc_hooks[];

static thread_local char orc_err[1024] = "";
static int orc_set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(orc_err, sizeof(orc_err), fmt, ap);
    va_end(ap);
    return code;
}

struct OrcAgent {
    const fgpu_agent *d;
    char *work, *swap, *newl;          /* d_<A>s, d_<A>s_swap, d_<A>s_new (sim:228-230) */
    std::vector<char *> state;         /* d_<A>s_<state> */
    int h_count;                       /* h_xmachine_memory_<A>_count */
    std::vector<int> h_state_count;
};

struct OrcMessage {
    const fgpu_message *d;
    const orc_message_hooks *hooks;
    char *list, *swap;                 /* d_<M>s, d_<M>s_swap (sim:260-261) */
    char *pbm;                         /* d_<M>_partition_matrix */
    std::vector<unsigned int> keys, values, keys_unsorted;
    int h_count;
};

struct orc_sim {
    const fgpu_model *model;
    std::vector<OrcAgent> agents;
    std::vector<OrcMessage> msgs;
    std::vector<int> gc_count;
    char *rand48;                      /* RNG_rand48: uvec2 A, C, seeds[buffer_size_MAX] (hdr:334-338) */
    unsigned int iteration;
    int threads;
};

static inline int *orc_col_i(char *list, size_t off) { return (int *)(list + off); }

/* ---- create: initialise() minus the XML read, sim:472-747 ------------------ */
/* the simulation whose init/step/exit function is running (the reference's host API is global state) */
static orc_sim *orc_cur = nullptr;

static orc_sim *orc_create_impl(const fgpu_model *m) {
    orc_sim *s = new orc_sim();
    s->model = m;
    s->iteration = 0;
    s->threads = 1;
    s->agents.resize(m->nagents);
    for (int a = 0; a < m->nagents; ++a) {
        OrcAgent &A = s->agents[a];
        A.d = &m->agents[a];
        A.work = (char *)calloc(1, A.d->list_bytes);
        A.swap = (char *)calloc(1, A.d->list_bytes);
        A.newl = (char *)calloc(1, A.d->list_bytes);
        A.state.resize(A.d->nstates);
        A.h_state_count.assign(A.d->nstates, 0);
        for (int st = 0; st < A.d->nstates; ++st) A.state[st] = (char *)calloc(1, A.d->list_bytes);
        A.h_count = 0;
    }
    s->msgs.resize(m->nmessages);
    for (int i = 0; i < m->nmessages; ++i) {
        OrcMessage &M = s->msgs[i];
        M.d = &m->messages[i];
        M.hooks = &orc_hooks[i];
        M.list = (char *)calloc(1, M.hooks->list_bytes);
        M.swap = (char *)calloc(1, M.hooks->list_bytes);
        M.pbm = M.hooks->pbm_bytes ? (char *)calloc(1, M.hooks->pbm_bytes) : nullptr;
        M.keys.resize(M.d->buffer_size);
        M.values.resize(M.d->buffer_size);
        M.keys_unsorted.resize(M.d->buffer_size);
        M.h_count = 0;
    }
    s->gc_count.assign(m->nfunctions, 0);
    /* rand48 seeding, sim:661-688 */
    const unsigned int N = (unsigned int)m->buffer_size_max;
    s->rand48 = (char *)calloc(1, 16 + 8 * (size_t)N);
    unsigned int *r = (unsigned int *)s->rand48;
    static const unsigned long long a = 0x5DEECE66DLL, c = 0xB;
    int seed = 123;
    unsigned long long A, C;
    A = 1LL;
    C = 0LL;
    for (unsigned int i = 0; i < N; ++i) {
        C += A * c;
        A *= a;
    }
    r[0] = A & 0xFFFFFFLL;
    r[1] = (A >> 24) & 0xFFFFFFLL;
    r[2] = C & 0xFFFFFFLL;
    r[3] = (C >> 24) & 0xFFFFFFLL;
    unsigned long long x = (((unsigned long long)seed) << 16) | 0x330E;
    for (unsigned int i = 0; i < N; ++i) {
        x = a * x + c;
        r[4 + 2 * i] = x & 0xFFFFFFLL;
        r[4 + 2 * i + 1] = (x >> 24) & 0xFFFFFFLL;
    }
    if (m->init_constants) m->init_constants();
    return s;
}

static void orc_destroy_impl(orc_sim *s) {
    if (!s) return;
    for (OrcAgent &A : s->agents) {
        free(A.work);
        free(A.swap);
        free(A.newl);
        for (char *p : A.state) free(p);
    }
    for (OrcMessage &M : s->msgs) {
        free(M.list);
        free(M.swap);
        free(M.pbm);
    }
    free(s->rand48);
    delete s;
}

/* ---- list primitives -------------------------------------------------------- */
/* reset_<A>_scan_input, krn:222-229 (bounded to n; the reference has no bounds check) */
static void orc_reset_scan_input(OrcAgent &A, char *list, int n) {
    int *pos = orc_col_i(list, A.d->position_offset), *scan = orc_col_i(list, A.d->scan_offset);
    for (int i = 0; i < n; ++i) {
        pos[i] = 0;
        scan[i] = 0;
    }
}

/* cub::DeviceScan::ExclusiveSum(_scan_input -> _position), e.g. sim:1466-1473 */
static void orc_exclusive_scan(OrcAgent &A, char *list, int n) {
    int *pos = orc_col_i(list, A.d->position_offset), *scan = orc_col_i(list, A.d->scan_offset);
    int run = 0;
    for (int i = 0; i < n; ++i) {
        pos[i] = run;
        run += scan[i];
    }
}

/* count after a scan: scan_last_sum (+1 if scan_last_included), e.g. sim:1476-1481 */
static int orc_scan_count(OrcAgent &A, char *list, int n) {
    if (n <= 0) return 0;
    int *pos = orc_col_i(list, A.d->position_offset), *scan = orc_col_i(list, A.d->scan_offset);
    return (scan[n - 1] == 1) ? pos[n - 1] + 1 : pos[n - 1];
}

/* elements of an agent variable: 1, or arrayLength (element e of agent i at var[e * MAX + i], hdr:187-194) */
static inline int orc_var_len(const fgpu_var &V) { return V.length > 1 ? V.length : 1; }

/* every variable of one agent; array variables element by element like krn:251-254,274-277 */
static void orc_copy_agent(OrcAgent &A, char *dst, int di, const char *src, int si) {
    const size_t MAX = (size_t)A.d->buffer_size;
    for (int v = 0; v < A.d->nvars; ++v) {
        const fgpu_var &V = A.d->vars[v];
        for (int e = 0; e < orc_var_len(V); ++e)
            memcpy(dst + V.list_offset + (e * MAX + (size_t)di) * V.size, src + V.list_offset + (e * MAX + (size_t)si) * V.size, V.size);
    }
}

/* scatter_<A>_Agents, krn:239-257 */
static void orc_scatter(OrcAgent &A, char *dst, const char *src, int dst_agent_count, int number_to_scatter) {
    const int *pos = (const int *)(src + A.d->position_offset), *scan = (const int *)(src + A.d->scan_offset);
    int *dpos = orc_col_i(dst, A.d->position_offset);
    for (int index = 0; index < number_to_scatter; ++index) {
        if (scan[index] == 1) {
            const int output_index = pos[index] + dst_agent_count;
            dpos[output_index] = output_index;
            orc_copy_agent(A, dst, output_index, src, index);
        }
    }
}

/* append_<A>_Agents, krn:265-280 */
static void orc_append(OrcAgent &A, char *dst, const char *src, int dst_agent_count, int number_to_append) {
    int *dpos = orc_col_i(dst, A.d->position_offset);
    for (int index = 0; index < number_to_append; ++index) {
        const int output_index = index + dst_agent_count;
        dpos[output_index] = output_index;
        orc_copy_agent(A, dst, output_index, src, index);
    }
}

/* ---- spatial message build (deterministic variant) --------------------------- */
/* message_<M>_grid_position + message_<M>_hash, krn:860-889 */
static unsigned int orc_cell_hash(const fgpu_message *M, float px, float py, float pz) {
    int g[3];
    const float p[3] = {px, py, pz};
    for (int k = 0; k < 3; ++k)
        g[k] = (int)floorf((p[k] - M->min_bounds[k]) * (float)M->dims[k] / (M->max_bounds[k] - M->min_bounds[k]));
    for (int k = 0; k < 3; ++k) {
        g[k] = (g[k] < 0) ? M->dims[k] - 1 : g[k];
        g[k] = (g[k] >= M->dims[k]) ? 0 : g[k];
    }
    return (unsigned int)(((g[2] * M->dims[1]) * M->dims[0]) + (g[1] * M->dims[0]) + g[0]);
}

/* hash_<M>_messages + SortPairs + reorder_<M>_messages, krn:944-1021, sim:1860-1885 */
static void orc_build_spatial(OrcMessage &M) {
    const fgpu_message *d = M.d;
    const int n = M.h_count;
    int *start = (int *)M.pbm;
    int *end_or_count = (int *)(M.pbm + M.hooks->pbm_end_offset);
    memset(M.pbm, 0, M.hooks->pbm_bytes); /* sim:1835 */
    if (n > 0) {
        /* krn:944-953: glm::vec3(messages->x[index], ...) -- the variable's own type (float or double)
         * converted to float */
        auto coord = [&](int var, int i) -> float {
            const fgpu_var &V = d->vars[var];
            const char *col = M.list + V.list_offset;
            return V.size == 8 ? (float)((const double *)col)[i] : ((const float *)col)[i];
        };
        for (int i = 0; i < n; ++i) {
            M.keys[i] = orc_cell_hash(d, coord(d->xvar, i), coord(d->yvar, i), coord(d->zvar, i));
            M.keys_unsorted[i] = M.keys[i];
            M.values[i] = (unsigned int)i;
        }
        /* cub::DeviceRadixSort::SortPairs over bits [0, radix_bits): a stable sort by key
         * (every key < grid_size <= 2^radix_bits) */
        std::vector<unsigned int> perm(n);
        for (int i = 0; i < n; ++i) perm[i] = (unsigned int)i;
        const unsigned int *keys = M.keys.data();
        std::stable_sort(perm.begin(), perm.end(), [keys](unsigned int a, unsigned int b) { return keys[a] < keys[b]; });
        std::vector<unsigned int> sk(n);
        for (int i = 0; i < n; ++i) sk[i] = M.keys_unsorted[perm[i]];
        for (int i = 0; i < n; ++i) {
            M.keys[i] = sk[i];
            M.values[i] = perm[i];
        }
        memset(start, 0xff, sizeof(int) * (size_t)d->grid_size); /* sim:1880 */
        /* reorder_<M>_messages (sort variant), krn:963-1021 */
        for (int index = 0; index < n; ++index) {
            const unsigned int key = M.keys[index];
            const unsigned int old_pos = M.values[index];
            if (index == 0) {
                start[key] = index;
            } else {
                const unsigned int prev_key = M.keys[index - 1];
                if (prev_key != key) {
                    start[key] = index;
                    end_or_count[prev_key] = index;
                }
            }
            if (index == n - 1) end_or_count[key] = index + 1;
            for (int v = 0; v < d->nvars; ++v) {
                const fgpu_var &V = d->vars[v];
                memcpy(M.swap + V.list_offset + (size_t)index * V.size, M.list + V.list_offset + (size_t)old_pos * V.size, V.size);
            }
        }
    }
    std::swap(M.list, M.swap); /* sim:1888-1891 */
}

/* ---- one agent function: <Agent>_<func>(stream), sim:1402-1974 ---------------- */
static int orc_run_function(orc_sim *s, int fi) {
    const fgpu_model *m = s->model;
    const fgpu_function &F = m->functions[fi];
    OrcAgent &A = s->agents[F.agent];
    const int cur = F.current_state, nxt = F.next_state;

    if (A.h_state_count[cur] == 0) return 0; /* sim:1417 */
    int state_list_size = A.h_state_count[cur];

    if (F.xagent_output >= 0) { /* sim:1427-1435 */
        OrcAgent &B = s->agents[F.xagent_output];
        orc_reset_scan_input(B, B.newl, std::min(state_list_size, B.d->buffer_size));
    }

    fgpu_filter fa;
    if (F.condition == FGPU_COND_FUNCTION) { /* sim:1439-1528 */
        A.h_count = A.h_state_count[cur];
        orc_reset_scan_input(A, A.state[cur], state_list_size);
        orc_reset_scan_input(A, A.work, state_list_size);
        fa.current = A.state[cur];
        fa.working = A.work;
        fa.n = A.h_count;
        fa.flags_current = fa.flags_working = nullptr;
        F.filter(&fa, nullptr);
        orc_exclusive_scan(A, A.state[cur], A.h_count);
        A.h_state_count[cur] = orc_scan_count(A, A.state[cur], A.h_count);
        orc_scatter(A, A.swap, A.state[cur], 0, A.h_count);
        std::swap(A.state[cur], A.swap);
        orc_exclusive_scan(A, A.work, A.h_count);
        const int working = orc_scan_count(A, A.work, A.h_count);
        orc_scatter(A, A.swap, A.work, 0, A.h_count);
        A.h_count = working;
        std::swap(A.work, A.swap);
        if (A.h_count == 0) return 0;
        state_list_size = A.h_count;
    } else if (F.condition == FGPU_COND_GLOBAL) { /* sim:1530-1588 */
        A.h_count = A.h_state_count[cur];
        orc_reset_scan_input(A, A.state[cur], state_list_size);
        fa.current = A.state[cur];
        fa.working = nullptr;
        fa.n = A.h_count;
        fa.flags_current = fa.flags_working = nullptr;
        F.filter(&fa, nullptr);
        orc_exclusive_scan(A, A.state[cur], A.h_count);
        const int global_conditions_true = orc_scan_count(A, A.state[cur], A.h_count);
        const bool mismatch = F.gc_must_evaluate_to ? (global_conditions_true != A.h_count) : (global_conditions_true == A.h_count);
        if (mismatch && s->gc_count[fi] < F.gc_max_iterations) {
            s->gc_count[fi]++;
            return 0;
        }
        s->gc_count[fi] = 0;
        std::swap(A.work, A.state[cur]);
        A.h_state_count[cur] = 0;
    } else { /* sim:1590-1601 */
        std::swap(A.work, A.state[cur]);
        A.h_count = A.h_state_count[cur];
        A.h_state_count[cur] = 0;
    }

    OrcMessage *Mout = F.output_message >= 0 ? &s->msgs[F.output_message] : nullptr;
    OrcMessage *Min = F.input_message >= 0 ? &s->msgs[F.input_message] : nullptr;
    if (Mout && Mout->h_count + A.h_count > Mout->d->buffer_size) /* sim:1607-1613 */
        return orc_set_err(-3, "Error: Buffer size of %s message will be exceeded in function %s", Mout->d->name, F.name);
    const bool optional_out = Mout && F.output_type == FGPU_OUT_OPTIONAL;
    if (optional_out && Mout->h_count > 0) /* sim:1709-1716 swaps the list away under an earlier producer */
        return orc_set_err(-5, "function %s: an optional_message output must be the first producer of message %s in an iteration", F.name,
                           Mout->d->name);
    int *const msg_position = Mout ? (int *)Mout->list : nullptr;                       /* int _position[MAX]  */
    int *const msg_scan_input = Mout ? (int *)Mout->list + Mout->d->buffer_size : nullptr; /* int _scan_input[MAX] */
    if (optional_out) /* reset_<M>_swaps(d_<M>s), sim:1666-1672 + krn:434-441 */
        for (int i = 0; i < state_list_size; ++i) {
            msg_position[i] = 0;
            msg_scan_input[i] = 0;
        }

    if (Mout) *Mout->hooks->output_type_symbol = F.output_type; /* sim:1660-1672 */
    if (F.reallocate) orc_reset_scan_input(A, A.work, state_list_size); /* sim:1675-1681 */

    fgpu_launch L;
    memset(&L, 0, sizeof(L));
    L.agents = A.work;
    L.n = A.h_count;
    if (F.xagent_output >= 0) L.new_agents = s->agents[F.xagent_output].newl;
    if (Min) {
        L.in_recs = Min->list;
        L.in_cell_start = (const unsigned int *)Min->pbm;
        L.in_count = Min->h_count;
    }
    if (Mout) {
        L.out_recs = Mout->list;
        L.out_base = Mout->h_count;
    }
    L.seeds = s->rand48;
#ifdef _OPENMP
    omp_set_num_threads(s->threads);
#endif
    F.launch(&L, nullptr); /* GPUFLAME_<f>, sim:1688-1692 */

    if (optional_out) {
        /* sim:1709-1750: swap the lists, exclusive sum of the sparse list's flags over the agents,
         * scatter_optional_<M>_messages (krn:411-428), count from the last prefix sum */
        std::swap(Mout->list, Mout->swap);
        const int *sparse_flag = (const int *)Mout->swap + Mout->d->buffer_size;
        int *sparse_pos = (int *)Mout->swap;
        int running = 0;
        for (int i = 0; i < A.h_count; ++i) {
            sparse_pos[i] = running;
            running += sparse_flag[i];
        }
        for (int index = 0; index < A.h_count; ++index) {
            if (sparse_flag[index] != 1) continue;
            const int output_index = sparse_pos[index] + Mout->h_count;
            ((int *)Mout->list)[output_index] = output_index;
            for (int v = 0; v < Mout->d->nvars; ++v) {
                const fgpu_var &V = Mout->d->vars[v];
                memcpy(Mout->list + V.list_offset + (size_t)output_index * V.size, Mout->swap + V.list_offset + (size_t)index * V.size, V.size);
            }
        }
        if (A.h_count > 0) Mout->h_count += sparse_pos[A.h_count - 1] + (sparse_flag[A.h_count - 1] == 1 ? 1 : 0);
        *Mout->hooks->count_symbol = Mout->h_count;
    } else if (Mout) { /* sim:1734-1754 */
        Mout->h_count += A.h_count;
        *Mout->hooks->count_symbol = Mout->h_count;
    }
    const int pre_death_count = A.h_count;
    if (F.reallocate) { /* sim:1763-1792 */
        orc_exclusive_scan(A, A.work, A.h_count);
        orc_scatter(A, A.swap, A.work, 0, A.h_count);
        std::swap(A.work, A.swap);
        A.h_count = orc_scan_count(A, A.swap, A.h_count);
    }
    if (F.xagent_output >= 0) { /* sim:1794-1829 */
        OrcAgent &B = s->agents[F.xagent_output];
        const int st = F.xagent_output_state;
        orc_exclusive_scan(B, B.newl, pre_death_count);
        const int after = B.h_state_count[st] + orc_scan_count(B, B.newl, pre_death_count);
        if (after > B.d->buffer_size)
            return orc_set_err(-3, "Error: Buffer size of %s agents in state %s will be exceeded writing new agents in function %s",
                               B.d->name, B.d->states[st], F.name);
        orc_scatter(B, B.state[st], B.newl, B.h_state_count[st], pre_death_count);
        B.h_state_count[st] = after;
    }
    if (Mout && Mout->d->partitioning == FGPU_PART_SPATIAL) orc_build_spatial(*Mout); /* sim:1831-1892 */

    /* sim:1936-1972 */
    if (A.h_state_count[nxt] + A.h_count > A.d->buffer_size)
        return orc_set_err(-3, "Error: Buffer size of %s agents in state %s will be exceeded moving working agents to next state in function %s",
                           F.name, A.d->states[nxt], F.name);
    if (F.swappable) {
        std::swap(A.work, A.state[cur]);
    } else {
        orc_append(A, A.state[nxt], A.work, A.h_state_count[nxt], A.h_count);
    }
    A.h_state_count[nxt] += A.h_count;
    return 0;
}

/* singleIteration() prologue, sim:885-892 */
static void orc_begin_iteration(orc_sim *s) {
    s->iteration++;
    for (OrcMessage &M : s->msgs) {
        M.h_count = 0;
        *M.hooks->count_symbol = 0;
    }
}

static int orc_single_iteration(orc_sim *s) {
    const fgpu_model *m = s->model;
    orc_begin_iteration(s);
    for (int l = 0; l < m->nlayers; ++l)
        for (int k = 0; k < m->layers[l].nfunctions; ++k) {
            int rc = orc_run_function(s, m->layers[l].functions[k]);
            if (rc) return rc;
        }
    orc_cur = s;
    for (int k = 0; k < m->nstep_functions; ++k) m->step_functions[k]();
    return 0;
}

/* integer sums wrap like the device's (signed overflow is undefined in host C++) */
template <typename T>
static T orc_wrap_add(T a, T b) { return (T)(a + b); }
template <>
int orc_wrap_add<int>(int a, int b) { return (int)((unsigned int)a + (unsigned int)b); }
template <>
long long orc_wrap_add<long long>(long long a, long long b) { return (long long)((unsigned long long)a + (unsigned long long)b); }

template <typename T>
static void orc_reduce_typed(const char *col, int n, int op, const void *count_value, void *result) {
    const T *c = (const T *)col;
    T acc = (op == 1 || op == 2) ? c[0] : (T)0;
    T cv = (T)0;
    if (op == 3) memcpy(&cv, count_value, sizeof(T));
    for (int i = 0; i < n; ++i) {
        if (op == 0) acc = orc_wrap_add(acc, c[i]);
        else if (op == 1) acc = c[i] < acc ? c[i] : acc;
        else if (op == 2) acc = acc < c[i] ? c[i] : acc;
        else acc = (T)(acc + (c[i] == cv ? (T)1 : (T)0));
    }
    memcpy(result, &acc, sizeof(T));
}
/* reduce_/count_/min_/max_<A>_<S>_<v>_variable() on the oracle's lists; op as in FGPU_REDUCE_* */
int orc_host_reduce(int agent, int state, int var, int op, const void *count_value, void *result) {
    orc_sim *s = orc_cur;
    if (!s || agent < 0 || agent >= (int)s->agents.size()) return orc_set_err(-1, "bad index");
    OrcAgent &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates || var < 0 || var >= A.d->nvars) return orc_set_err(-1, "bad index");
    const fgpu_var &V = A.d->vars[var];
    const int n = A.h_state_count[state];
    if (n == 0) {
        if (op == 1 || op == 2) return orc_set_err(-1, "min/max of an empty list");
        memset(result, 0, (size_t)V.size);
        return 0;
    }
    const char *col = A.state[state] + V.list_offset;
    const bool uns = strstr(V.ctype, "unsigned") != nullptr;
    if (V.is_float) {
        if (V.size == 8) orc_reduce_typed<double>(col, n, op, count_value, result);
        else orc_reduce_typed<float>(col, n, op, count_value, result);
    } else if (V.size == 8) {
        if (uns) orc_reduce_typed<unsigned long long>(col, n, op, count_value, result);
        else orc_reduce_typed<long long>(col, n, op, count_value, result);
    } else {
        if (uns) orc_reduce_typed<unsigned int>(col, n, op, count_value, result);
        else orc_reduce_typed<int>(col, n, op, count_value, result);
    }
    return 0;
}
/* ---- C ABI (same shapes as include/fgpu.h, prefix orc_) ------------------------ */
extern "C" {

const char *orc_last_error(void) { return orc_err; }
void *orc_sim_create(void) { return orc_create_impl(fgpu_model_get()); }
void orc_sim_destroy(void *p) { orc_destroy_impl((orc_sim *)p); }
int orc_sim_set_threads(void *p, int n) {
    ((orc_sim *)p)->threads = n < 1 ? 1 : n;
    return 0;
}

static int orc_upload(orc_sim *s, int agent, int state, int offset, int count, const void *const *columns) {
    if (agent < 0 || agent >= (int)s->agents.size()) return orc_set_err(-1, "bad agent");
    OrcAgent &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates) return orc_set_err(-1, "bad state");
    if (count < 0 || offset + count > A.d->buffer_size) return orc_set_err(-3, "ERROR: MAX Buffer size (%i) for agent %s exceeded whilst reading data", A.d->buffer_size, A.d->name);
    for (int v = 0; v < A.d->nvars; ++v)
      for (int e = 0; e < orc_var_len(A.d->vars[v]); ++e) {   /* array variables: `length` blocks of `count` values */
        const fgpu_var &V = A.d->vars[v];
        char *dst = A.state[state] + V.list_offset + ((size_t)e * A.d->buffer_size + (size_t)offset) * V.size;
        if (columns && columns[v]) {
            memcpy(dst, (const char *)columns[v] + (size_t)e * count * V.size, (size_t)count * V.size);
        } else {
            for (int i = 0; i < count; ++i) {
                if (V.size == 8) {
                    if (V.is_float) ((double *)dst)[i] = V.default_value;
                    else ((long long *)dst)[i] = (long long)V.default_value;
                } else if (V.is_float) ((float *)dst)[i] = (float)V.default_value;
                else if (!strcmp(V.ctype, "unsigned int")) ((unsigned int *)dst)[i] = (unsigned int)V.default_value;
                else ((int *)dst)[i] = (int)V.default_value;
            }
        }
    }
    A.h_state_count[state] = offset + count;
    return 0;
}

int orc_sim_set_agents(void *p, int agent, int state, int count, const void *const *columns) {
    return orc_upload((orc_sim *)p, agent, state, 0, count, columns);
}
int orc_sim_add_agents(void *p, int agent, int state, int count, const void *const *columns) {
    orc_sim *s = (orc_sim *)p;
    if (agent < 0 || agent >= (int)s->agents.size() || state < 0 || state >= s->agents[agent].d->nstates) return orc_set_err(-1, "bad index");
    return orc_upload(s, agent, state, s->agents[agent].h_state_count[state], count, columns);
}
int orc_sim_load_states(void *, const char *) { return orc_set_err(-4, "the oracle reads states through oracle/xml_states.py"); }
int orc_sim_run_init_functions(void *p) {
    orc_sim *s = (orc_sim *)p;
    orc_cur = s;
    for (int k = 0; k < s->model->ninit_functions; ++k) s->model->init_functions[k]();
    return 0;
}
int orc_sim_step(void *p, int n) {
    for (int i = 0; i < n; ++i) {
        int rc = orc_single_iteration((orc_sim *)p);
        if (rc) return rc;
    }
    return 0;
}
int orc_sim_begin_iteration(void *p) {
    orc_begin_iteration((orc_sim *)p);
    return 0;
}
int orc_sim_run_function(void *p, int f) {
    orc_sim *s = (orc_sim *)p;
    if (f < 0 || f >= s->model->nfunctions) return orc_set_err(-1, "bad function");
    return orc_run_function(s, f);
}
unsigned int orc_sim_iteration(void *p) { return ((orc_sim *)p)->iteration; }
int orc_sim_agent_count(void *p, int agent, int state) {
    orc_sim *s = (orc_sim *)p;
    if (agent < 0 || agent >= (int)s->agents.size() || state < 0 || state >= s->agents[agent].d->nstates) return orc_set_err(-1, "bad index");
    return s->agents[agent].h_state_count[state];
}
int orc_sim_reduce_agent_var(void *p, int agent, int state, int var, int op, const void *count_value, void *result) {
    orc_cur = (orc_sim *)p;
    return orc_host_reduce(agent, state, var, op, count_value, result);
}
int orc_sim_read_agent_var(void *p, int agent, int state, int var, void *dst, int count) {
    orc_sim *s = (orc_sim *)p;
    if (agent < 0 || agent >= (int)s->agents.size()) return orc_set_err(-1, "bad index");
    OrcAgent &A = s->agents[agent];
    if (state < 0 || state >= A.d->nstates || var < 0 || var >= A.d->nvars) return orc_set_err(-1, "bad index");
    const fgpu_var &V = A.d->vars[var];
    for (int e = 0; e < orc_var_len(V); ++e)
        memcpy((char *)dst + (size_t)e * count * V.size, A.state[state] + V.list_offset + (size_t)e * A.d->buffer_size * V.size, (size_t)count * V.size);
    return 0;
}
int orc_sim_message_count(void *p, int message) {
    orc_sim *s = (orc_sim *)p;
    if (message < 0 || message >= (int)s->msgs.size()) return orc_set_err(-1, "bad index");
    return s->msgs[message].h_count;
}
int orc_sim_read_message_var(void *p, int message, int var, void *dst, int count) {
    orc_sim *s = (orc_sim *)p;
    if (message < 0 || message >= (int)s->msgs.size()) return orc_set_err(-1, "bad index");
    OrcMessage &M = s->msgs[message];
    if (var < 0 || var >= M.d->nvars) return orc_set_err(-1, "bad index");
    const fgpu_var &V = M.d->vars[var];
    memcpy(dst, M.list + V.list_offset, (size_t)count * V.size);
    return 0;
}
/* normalised PBM: count, and start where count > 0 */
int orc_sim_read_pbm(void *p, int message, int *start, int *count) {
    orc_sim *s = (orc_sim *)p;
    if (message < 0 || message >= (int)s->msgs.size()) return orc_set_err(-1, "bad index");
    OrcMessage &M = s->msgs[message];
    if (!M.pbm) return orc_set_err(-1, "not spatial");
    const int *st = (const int *)M.pbm, *en = (const int *)(M.pbm + M.hooks->pbm_end_offset);
    for (int c = 0; c < M.d->grid_size; ++c) {
        const bool empty = (M.h_count == 0) || ((unsigned int)st[c] == 0xffffffffu);
        count[c] = empty ? 0 : en[c] - st[c];
        start[c] = empty ? 0 : st[c];
    }
    return 0;
}
int orc_sim_read_message_keys(void *p, int message, unsigned int *keys, int count) {
    orc_sim *s = (orc_sim *)p;
    memcpy(keys, s->msgs[message].keys_unsorted.data(), sizeof(unsigned int) * (size_t)count);
    return 0;
}
int orc_sim_read_message_order(void *p, int message, unsigned int *order, int count) {
    orc_sim *s = (orc_sim *)p;
    memcpy(order, s->msgs[message].values.data(), sizeof(unsigned int) * (size_t)count);
    return 0;
}
int orc_sim_read_seeds(void *p, unsigned int *dst, int count) {
    memcpy(dst, ((orc_sim *)p)->rand48 + 16, 8 * (size_t)count);
    return 0;
}
int orc_sim_rand48_constants(void *p, unsigned int *out) {
    memcpy(out, ((orc_sim *)p)->rand48, 16);
    return 0;
}
int orc_sim_set_constant(void *p, const char *name, const void *value) {
    orc_sim *s = (orc_sim *)p;
    for (int i = 0; i < s->model->nconstants; ++i)
        if (!strcmp(s->model->constants[i].name, name)) {
            s->model->constants[i].set(value);
            return 0;
        }
    return orc_set_err(-1, "unknown constant %s", name);
}
int orc_sim_get_constant(void *p, const char *name, void *value) {
    orc_sim *s = (orc_sim *)p;
    for (int i = 0; i < s->model->nconstants; ++i)
        if (!strcmp(s->model->constants[i].name, name)) {
            memcpy(value, s->model->constants[i].get(), (size_t)s->model->constants[i].size * s->model->constants[i].length);
            return 0;
        }
    return orc_set_err(-1, "unknown constant %s", name);
}

} /* extern "C" */
#endif
