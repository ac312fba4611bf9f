/* fgpu_rt_reader.inl -- 0.xml reader (readInitialStates, io.xslt:211-621).
 * Part of the single translation unit fgpu_rt.cu (textually included there, in this order, after the
 * runtime objects it uses); not a standalone source file. */
/* ------------------------------------------------------------------------- */
/* 0.xml reader -- character stream, same tag rules as io.xslt:336-583        */
/* ------------------------------------------------------------------------- */
namespace {
struct XmlAgentBuf {
    std::vector<std::vector<char>> cols; /* per variable, raw bytes; an array variable holds `length` values per agent */
    int count = 0;
};

/* element-major copy (`length` blocks of `count` values, the order of the list on the device) of an agent-major buffer */
std::vector<char> element_major(const std::vector<char> &rows, int count, int length, int size) {
    std::vector<char> out(rows.size());
    for (int i = 0; i < count; ++i)
        for (int e = 0; e < length; ++e)
            memcpy(out.data() + ((size_t)e * count + i) * size, rows.data() + ((size_t)i * length + e) * size, (size_t)size);
    return out;
}

/* atof / strtod of a value text (fgpu_atof, fpgu_strtod: io.xslt:55-64).  Plain decimals -- the only thing a states file written by
 * saveIterationData holds -- go through std::from_chars, which is correctly rounded like glibc's strtod and several times faster; anything
 * else (leading blanks, a sign '+', hex floats, inf / nan, trailing text, out-of-range magnitudes) takes strtod itself. */
double parse_double(const char *text) {
    const char *p = text;
    if (*p == '-') ++p;
    const char *digits = p;
    while (*p >= '0' && *p <= '9') ++p;
    bool any = p != digits;
    if (*p == '.') {
        const char *frac = ++p;
        while (*p >= '0' && *p <= '9') ++p;
        any = any || p != frac;
    }
    if (any && (*p == 'e' || *p == 'E')) {
        const char *e = p + 1;
        if (*e == '+' || *e == '-') ++e;
        if (*e >= '0' && *e <= '9') {
            while (*e >= '0' && *e <= '9') ++e;
            p = e;
        }
    }
    if (any && *p == '\0') {
        double v = 0.0;
        const char *first = text;
        const std::from_chars_result r = std::from_chars(first, p, v);
        if (r.ec == std::errc() && r.ptr == p) return v;
    }
    return strtod(text, nullptr);
}

void store_value(const fgpu_var &V, const char *text, char *dst) {
    if (V.is_float) {
        if (V.size == 4) *(float *)dst = (float)parse_double(text);      /* fgpu_atof */
        else *(double *)dst = parse_double(text);                        /* fpgu_strtod */
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
/* readArrayInput (io.xslt:69-93): comma-separated items; too many items is fatal in the reference (exit), a short
 * list leaves the remaining elements at their default and warns on stderr */
int store_array(const fgpu_var &V, const char *agent_name, const std::string &text, char *dst) {
    const int expected = V.length;
    std::string buf = text;
    char *save = nullptr;
    int i = 0;
    for (char *tok = strtok_r(&buf[0], ",", &save); tok != nullptr; tok = strtok_r(nullptr, ",", &save)) {
        if (i >= expected)
            return set_err(FGPU_ERR_IO, "Error: variable array %s->%s has too many items (%d), expected %d!", agent_name, V.name, i, expected);
        store_value(V, tok, dst + (size_t)i * V.size);
        ++i;
    }
    if (i != expected) fprintf(stderr, "Warning: variable array %s->%s has %d items, expected %d!\n", agent_name, V.name, i, expected);
    return FGPU_OK;
}

/* The parser proper: host code only.  Fills one column buffer per agent variable with the agents of the file (each
 * under the agent type named by its <name>); <environment> values go to the model's constant setters when
 * `apply_environment` is set (fgpu_sim_load_states) and are skipped otherwise (fgpu_states_file_*: no device). */
int parse_states_file(const fgpu_model *m, const char *xml_path, std::vector<XmlAgentBuf> &bufs, bool apply_environment) {
    FILE *file = fopen(xml_path, "r");
    if (!file) return set_err(FGPU_ERR_IO, "Could not open input file %s", xml_path);

    bufs.assign(m->nagents, XmlAgentBuf());
    /* current values per agent type per variable, reset to defaults after each </xagent> */
    std::vector<std::vector<std::vector<char>>> cur(m->nagents);
    auto reset_cur = [&]() {
        for (int a = 0; a < m->nagents; ++a) {
            cur[a].resize(m->agents[a].nvars);
            for (int v = 0; v < m->agents[a].nvars; ++v) {
                const fgpu_var &V = m->agents[a].vars[v];
                cur[a][v].assign((size_t)8 * var_len(V), 0);
                for (int e = 0; e < var_len(V); ++e) default_value(V, cur[a][v].data() + (size_t)e * V.size);
            }
        }
    };
    reset_cur();
    for (int a = 0; a < m->nagents; ++a) bufs[a].cols.resize(m->agents[a].nvars);

    /* the reference reads with fgetc (io.xslt:336); the same character stream through a 4 MB block buffer (a 16M-agent file is 2 GB) */
    std::vector<char> block(4u << 20);
    size_t block_pos = 0, block_len = 0;
    auto next_char = [&]() -> int {
        if (block_pos == block_len) {
            block_len = fread(block.data(), 1, block.size(), file);
            block_pos = 0;
            if (block_len == 0) return EOF;
        }
        return (unsigned char)block[block_pos++];
    };
    std::string buffer, agentname;
    buffer.reserve(10001);
    bool in_tag = false, in_comment = false, in_itno = false, in_env = false, in_xagent = false, in_name = false;
    bool reading = true;
    std::string open_var; /* the variable tag we are inside of, matched by bare name (io.xslt:458-459) */
    int rc = FGPU_OK;
    int ch;
    while (reading && (ch = next_char()) != EOF) {
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
                        const int sz = m->agents[a].vars[v].size * var_len(m->agents[a].vars[v]);
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
                        if (open_var == m->agents[a].vars[v].name) {
                            if (var_len(m->agents[a].vars[v]) > 1) {
                                rc = store_array(m->agents[a].vars[v], m->agents[a].name, buffer, cur[a][v].data());
                                if (rc != FGPU_OK) reading = false;
                            } else {
                                store_value(m->agents[a].vars[v], buffer.c_str(), cur[a][v].data());
                            }
                        }
            } else if (in_env && apply_environment) {
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
    return rc;
}
}  // namespace

extern "C" int fgpu_sim_load_states(fgpu_sim *s, const char *xml_path) {
    if (!s || !xml_path) return set_err(FGPU_ERR_ARG, "bad argument");
    const fgpu_model *m = s->model;
    CU(cudaSetDevice(s->device));
    if (m->init_constants) m->init_constants(); /* initEnvVars(), io.xslt:254-255 */
    std::vector<XmlAgentBuf> bufs;
    int rc = parse_states_file(m, xml_path, bufs, true);
    if (rc != FGPU_OK) return rc;
    for (int a = 0; a < m->nagents; ++a) {
        std::vector<const void *> cols(m->agents[a].nvars);
        for (int v = 0; v < m->agents[a].nvars; ++v) {
            const fgpu_var &V = m->agents[a].vars[v];
            if (var_len(V) > 1) bufs[a].cols[v] = element_major(bufs[a].cols[v], bufs[a].count, var_len(V), V.size);
            cols[v] = bufs[a].cols[v].data();
        }
        /* other states start empty */
        for (int st = 0; st < m->agents[a].nstates; ++st) s->agents[a].h_state_count[st] = 0;
        rc = upload_columns(s, a, m->agents[a].initial_state, 0, bufs[a].count, bufs[a].count ? cols.data() : nullptr);
        if (rc != FGPU_OK) return rc;
    }
    CU(cudaDeviceSynchronize());
    if (m->ids_after_load) { /* io.xslt:599-616: the first generated id = the largest id read + 1 */
        if (m->bind) m->bind(s);
        m->ids_after_load(s);
    }
    return FGPU_OK;
}

/* ---- a states file without a device: how many agents of a type it holds, and their columns ------------------------ */
extern "C" int fgpu_states_file_count(const fgpu_model *model, const char *xml_path, int agent) {
    if (!model || !xml_path || agent < 0 || agent >= model->nagents) return set_err(FGPU_ERR_ARG, "bad argument");
    std::vector<XmlAgentBuf> bufs;
    const int rc = parse_states_file(model, xml_path, bufs, false);
    return rc != FGPU_OK ? rc : bufs[agent].count;
}

extern "C" int fgpu_states_file_read_var(const fgpu_model *model, const char *xml_path, int agent, int var, void *dst, int count) {
    if (!model || !xml_path || !dst || agent < 0 || agent >= model->nagents || var < 0 || var >= model->agents[agent].nvars)
        return set_err(FGPU_ERR_ARG, "bad argument");
    std::vector<XmlAgentBuf> bufs;
    const int rc = parse_states_file(model, xml_path, bufs, false);
    if (rc != FGPU_OK) return rc;
    if (count < 0 || count > bufs[agent].count) return set_err(FGPU_ERR_ARG, "the file holds %d agents of type %s", bufs[agent].count, model->agents[agent].name);
    const fgpu_var &V = model->agents[agent].vars[var];
    if (var_len(V) > 1) {   /* element-major, like fgpu_sim_read_agent_var: `length` blocks of `count` values */
        std::vector<char> rows(bufs[agent].cols[var].begin(), bufs[agent].cols[var].begin() + (size_t)count * var_len(V) * V.size);
        const std::vector<char> em = element_major(rows, count, var_len(V), V.size);
        memcpy(dst, em.data(), em.size());
        return FGPU_OK;
    }
    memcpy(dst, bufs[agent].cols[var].data(), (size_t)count * V.size);
    return FGPU_OK;
}
