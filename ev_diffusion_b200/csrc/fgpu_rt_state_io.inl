/* fgpu_rt_state_io.inl -- state writer, binary snapshot, host analytics reductions.
 * Part of the single translation unit fgpu_rt.cu (textually included there, in this order, after the
 * runtime objects it uses); not a standalone source file. */
/* ------------------------------------------------------------------------- */
/* state writer (io.xslt:124-200) and binary snapshot                         */
/* ------------------------------------------------------------------------- */
namespace {
/* formatSpecifier, _common_templates.xslt:25-60: integers by width and sign, everything else "%f" */
int write_value(FILE *f, const char *ctype, int size, bool is_float, const void *p, int digits = 0) {
    /* option xml_float_digits: "%.<digits>g" instead of the reference's lossy "%f" (9 digits round-trip a float) */
    if (is_float && digits > 0) return size == 8 ? fprintf(f, "%.*g", digits, *(const double *)p) : fprintf(f, "%.*g", digits, (double)*(const float *)p);
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
            /* an array variable: `length` blocks of `count` values (element-major, the order of the list) */
            const size_t block = (size_t)A.h_state_count[st] * V.size;
            cols[st][v].resize(block * var_len(V));
            for (int e = 0; block && e < var_len(V); ++e)
                CU(cudaMemcpyAsync(cols[st][v].data() + (size_t)e * block, A.state[st] + V.list_offset + (size_t)e * A.d->buffer_size * V.size, block,
                                   cudaMemcpyDeviceToHost, s->stream));
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
    std::vector<char> iobuf(4u << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
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
                    for (int e = 0; e < var_len(V); ++e) {   /* arrays: comma-separated (io.xslt:176-186) */
                        write_value(f, V.ctype, V.size, V.is_float != 0, cols[st][v].data() + ((size_t)e * A.h_state_count[st] + i) * V.size,
                                    s->opt_xml_digits);
                        if (e != var_len(V) - 1) fputc(',', f);
                    }
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
        for (int v = 0; ok && v < A.d->nvars; ++v) ok = put_str(f, A.d->vars[v].name) && put_u32(f, (unsigned int)(A.d->vars[v].size * var_len(A.d->vars[v])));
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
            if (!get_str(f, &name) || name != A.d->vars[v].name || !get_u32(f, &size) || (int)size != A.d->vars[v].size * var_len(A.d->vars[v]))
                return fail("variable layout differs");
        }
        for (int st = 0; st < A.d->nstates; ++st) {
            unsigned int count = 0;
            if (!get_str(f, &name) || name != A.d->states[st] || !get_u32(f, &count)) return fail("state layout differs");
            if ((int)count > A.d->buffer_size) return fail("population exceeds the buffer size of this build");
            std::vector<std::vector<char>> cols(A.d->nvars);
            std::vector<const void *> ptrs(A.d->nvars);
            for (int v = 0; v < A.d->nvars; ++v) {
                const size_t bytes = (size_t)count * A.d->vars[v].size * var_len(A.d->vars[v]);
                cols[v].resize(bytes + 1);
                if (!get(f, cols[v].data(), bytes)) return fail("truncated");
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
    if (var_len(V) > 1) return set_err(FGPU_ERR_ARG, "reductions are not defined for agent array variables (header.xslt:724)");
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
    else if (!strcmp(key, "instrument_iterations")) s->opt_instr_iterations = value ? 1 : 0;
    else if (!strcmp(key, "instrument_agent_functions")) s->opt_instr_agent_functions = value ? 1 : 0;
    else if (!strcmp(key, "instrument_init_functions")) s->opt_instr_init_functions = value ? 1 : 0;
    else if (!strcmp(key, "instrument_step_functions")) s->opt_instr_step_functions = value ? 1 : 0;
    else if (!strcmp(key, "instrument_exit_functions")) s->opt_instr_exit_functions = value ? 1 : 0;
    else if (!strcmp(key, "pdl")) s->opt_pdl = value ? 1 : 0;
    else if (!strcmp(key, "msd")) s->opt_msd = value ? 1 : 0;
    else if (!strcmp(key, "mirror")) s->opt_mirror = value ? 1 : 0;
    else if (!strcmp(key, "l2_fetch")) {
        /* device-wide hint: bytes the L2 fetches from HBM per miss (32, 64 or 128).  The random 16-byte record
         * gather of the message build wastes most of a 64-byte fetch. */
        if (value != 32 && value != 64 && value != 128) return set_err(FGPU_ERR_ARG, "l2_fetch must be 32, 64 or 128");
        CU(cudaSetDevice(s->device));
        CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
        return FGPU_OK;
    }
    else if (!strcmp(key, "carry")) {
        s->opt_carry = value ? 1 : 0;
        if (s->opt_carry) {
            CU(cudaSetDevice(s->device));
            int rc = alloc_carry(s);
            if (rc) return rc;
        }
    }
    else if (!strcmp(key, "xml_float_digits")) {
        if (value < 0 || value > 17) return set_err(FGPU_ERR_ARG, "xml_float_digits must be in [0,17]");
        s->opt_xml_digits = value;
        return FGPU_OK;
    }
    else if (!strcmp(key, "slab_p2p")) {
        if (s->slab.on) return set_err(FGPU_ERR_ARG, "slab_p2p must be set before fgpu_sim_slab_init");
        s->opt_slab_p2p = value ? 1 : 0;
    } else if (!strcmp(key, "slab_direct")) {
        s->opt_slab_direct = value ? 1 : 0;
    } else if (!strcmp(key, "slab_overlap")) {
        s->opt_slab_overlap = value ? 1 : 0;
    } else if (!strcmp(key, "slab_inplace")) {
        s->opt_slab_inplace = value ? 1 : 0;
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
