/* fgpu_rt_replay.inl -- slab mode: replay of captured iterations.
 * Part of the single translation unit fgpu_rt.cu (textually included there, in this order, after the
 * runtime objects it uses); not a standalone source file. */
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
    k.push_back((unsigned long long)s->opt_carry);
    k.push_back((unsigned long long)s->opt_msd);
    k.push_back((unsigned long long)s->opt_slab_inplace);
    k.push_back((unsigned long long)s->opt_slab_overlap);
    k.push_back((unsigned long long)s->opt_pdl);
    k.push_back((unsigned long long)s->opt_mirror);
    k.push_back((unsigned long long)s->opt_slab_direct);
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
