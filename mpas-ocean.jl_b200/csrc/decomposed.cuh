// decomposed.cuh -- domain-decomposed stepping INSIDE the library: the halo exchange (NCCL send/recv, or direct stores into
// the peers' memory), the two-stream overlap schedule, its events and the captured step graphs.  A host program -- the Julia
// shim, the Python mirror -- only calls mokab_comm_init, mokab_decomp_setup and mokab_timestep_*_decomposed.
// Included by moka_b200.cu after the staged entry points it is built from (run_stage, halo_pack, p2p_*).
//
// The reference has no multi-device path; this is BASELINE.json's north_star (e).  Schedule of one RK stage
// s on a rank (`compute` = the context's stream, `halo` = a high-priority stream of the state):
//
//     halo:     [wait I(s-1)]  BOUNDARY blocks(s)  record B(s)   pack(s) -> all-to-all -> unpack(s)     (or: push / wait kernels,
//     compute:  [wait B(s-1)]  INTERIOR blocks(s)  record I(s)                                           or nothing: PUSH blocks)
//
// BOUNDARY blocks are those whose stencils read a halo entity or that hold an entity a neighbour needs (mokab_halo_setup), so
// the message of stage s leaves while the bulk of stage s is still computing and is in place before stage s+1 touches the
// halo: stage s+1 on either stream waits for stage s of the other.  One and two consecutive steps (for either time-level
// parity) are captured -- collective calls included -- into CUDA graphs and replayed.
#pragma once

namespace mokab {

static inline cudaStream_t halo_stream(mokab_state *st) { return st->dec.overlap() ? st->dec.halo : st->ctx->stream; }

static void decomp_events_reserve(mokab_state *st, size_t n)
{
    while (st->dec.events.size() < n) {
        cudaEvent_t e = nullptr;
        MOKAB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st->dec.events.push_back(e);
    }
}

// one exchange of the output of `stage` (mokab_halo_pack's numbering) on stream s
template <class R>
static void decomp_exchange(mokab_state *st, int stage, cudaStream_t s)
{
    mokab_state::Decomp &D = st->dec;
    if (D.mode == MOKAB_HALO_P2P) {
        p2p_push<R>(st, stage, s);
        p2p_wait(st, s);
        return;
    }
    if (D.mode == MOKAB_HALO_P2P_LL) {
        p2p_push_ll<R>(st, stage, s);
        p2p_wait_ll<R>(st, stage, s);
        return;
    }
    halo_pack<R>(st, stage, D.sendBuf.p, s, false);
    comm::all_to_all(D.comm, s, D.sendBuf.p, D.recvBuf.p, D.scnt.data(), D.rcnt.data(), sizeof(R));
    halo_pack<R>(st, stage, D.recvBuf.p, s, true);
}

// Enqueue `nsteps` RungeKutta4 steps from time level st->cur; both streams are joined on the compute stream on entry and exit.
template <class R>
static void decomp_enqueue_rk4(mokab_state *st, double dt, int64_t nsteps)
{
    mokab_state::Decomp &D = st->dec;
    cudaStream_t compute = st->ctx->stream;
    const bool fused = D.mode == MOKAB_HALO_P2P_FUSED;             // the boundary launch carries the exchange itself
    const int boundary = fused ? MOKAB_PART_BOUNDARY_PUSH : MOKAB_PART_BOUNDARY;
    if (nsteps <= 0) return;
    if constexpr (sizeof(R) == 8) {
        if (st->K > 1) {   // multi-level states: whole-part launches of k_rk_stage_ml, one K + 1 plane message per stage (whatever the halo mode)
            for (int64_t i = 0; i < nsteps; ++i) {
                enqueue_rk4_step_ml(st, dt, st->cur);
                st->cur = 1 - st->cur;
            }
            return;
        }
    }
    // Experiment (option "decomp_serial_blocks", off by default): parts of a few hundred blocks (Kelvin 1024x1024 over 8 GPUs: 512)
    // are a fraction of one wave of resident blocks, a stage kernel lasts ~12 us and the two-stream schedule's chain of launches
    // and cross-stream waits is the stage time; with the exchange folded into the launch such a part can run ONE kernel per
    // stage over all its blocks (gate on the neighbours' previous stage at entry, store what they need, tick them at exit; same
    // protocol, same hazard argument).  Measured at N = 2 it is SLOWER (Kelvin 1024x1024: 0.389 vs 0.236 ms/step; 512 blocks per
    // GPU: 0.136 vs 0.111): every block of stage s + 1 then waits for the neighbours' WHOLE stage s, the interior included.
    const bool one_launch = fused && st->mesh->fusedBlocks < options().decomp_serial_blocks;
    if (one_launch) {
        for (int64_t i = 0; i < nsteps; ++i) {
            for (int s = 1; s <= 4; ++s) run_stage<R>(st, dt, s, MOKAB_PART_ALL_PUSH, compute);
            st->cur = 1 - st->cur;
        }
        p2p_wait_arrivals(st, compute);
        return;
    }
    if (!D.overlap()) {
        for (int64_t i = 0; i < nsteps; ++i) {
            for (int s = 1; s <= 4; ++s) {
                if (fused) {
                    run_stage<R>(st, dt, s, boundary, compute);
                    run_stage<R>(st, dt, s, MOKAB_PART_INTERIOR, compute);
                } else {
                    run_stage<R>(st, dt, s, MOKAB_PART_ALL, compute);
                    decomp_exchange<R>(st, s, compute);
                }
            }
            st->cur = 1 - st->cur;
        }
        if (fused) p2p_wait_arrivals(st, compute);
        return;
    }
    cudaStream_t halo = D.halo;
    decomp_events_reserve(st, 18);
    MOKAB_CUDA(cudaEventRecord(D.events[16], compute));            // fork
    MOKAB_CUDA(cudaStreamWaitEvent(halo, D.events[16], 0));
    for (int64_t i = 0; i < nsteps; ++i) {
        for (int s = 1; s <= 4; ++s) {
            // a pair of events per stage of two consecutive steps: re-recorded only after both of its waits were enqueued
            cudaEvent_t evB = D.events[(size_t)(((i & 1) * 4 + (s - 1)) * 2)], evI = D.events[(size_t)(((i & 1) * 4 + (s - 1)) * 2 + 1)];
            run_stage<R>(st, dt, s, boundary, halo);
            MOKAB_CUDA(cudaEventRecord(evB, halo));
            run_stage<R>(st, dt, s, MOKAB_PART_INTERIOR, compute);
            MOKAB_CUDA(cudaEventRecord(evI, compute));
            if (!fused) decomp_exchange<R>(st, s, halo);
            if (options().test_drop_dependency != 1)
                MOKAB_CUDA(cudaStreamWaitEvent(compute, evB, 0));  // stage s+1 interior reads stage s boundary output
            if (options().test_drop_dependency != 2)
                MOKAB_CUDA(cudaStreamWaitEvent(halo, evI, 0));     // stage s+1 boundary reads stage s interior output
        }
        st->cur = 1 - st->cur;
    }
    if (fused) p2p_wait_arrivals(st, halo);                        // the neighbours' last stores, before anything else touches the halo slots
    MOKAB_CUDA(cudaEventRecord(D.events[17], halo));               // join
    MOKAB_CUDA(cudaStreamWaitEvent(compute, D.events[17], 0));
}

// ForwardEuler (the reference driver's stepper, time_integration.jl:150-193): one launch per part, then the halo copies of
// everything the step wrote -- (h, u) and (ssh, layerThicknessEdge), two messages -- while the interior blocks run.
static void decomp_enqueue_fe(mokab_state *st, double dt, int64_t nsteps)
{
    mokab_state::Decomp &D = st->dec;
    cudaStream_t compute = st->ctx->stream;
    if (nsteps <= 0) return;
    MOKAB_REQUIRE(D.mode == MOKAB_HALO_NCCL || D.mode == MOKAB_HALO_P2P_LL,
                  "timestep_forward_euler_decomposed: ForwardEuler steps use the packed exchange (MOKAB_HALO_NCCL) or the flag-in-data one (MOKAB_HALO_P2P_LL)");
    if (!D.overlap()) {
        for (int64_t i = 0; i < nsteps; ++i) {
            run_fe_stage(st, dt, MOKAB_PART_ALL, compute);
            decomp_exchange<double>(st, 4, compute);
            decomp_exchange<double>(st, 5, compute);
            st->cur = 1 - st->cur;
        }
        return;
    }
    cudaStream_t halo = D.halo;
    decomp_events_reserve(st, 18);
    MOKAB_CUDA(cudaEventRecord(D.events[16], compute));
    MOKAB_CUDA(cudaStreamWaitEvent(halo, D.events[16], 0));
    for (int64_t i = 0; i < nsteps; ++i) {
        cudaEvent_t evB = D.events[(size_t)((i & 3) * 2)], evI = D.events[(size_t)((i & 3) * 2 + 1)];
        run_fe_stage(st, dt, MOKAB_PART_BOUNDARY, halo);
        MOKAB_CUDA(cudaEventRecord(evB, halo));
        run_fe_stage(st, dt, MOKAB_PART_INTERIOR, compute);
        MOKAB_CUDA(cudaEventRecord(evI, compute));
        decomp_exchange<double>(st, 4, halo);
        decomp_exchange<double>(st, 5, halo);
        MOKAB_CUDA(cudaStreamWaitEvent(compute, evB, 0));          // the next interior launch overwrites what this boundary launch read
        MOKAB_CUDA(cudaStreamWaitEvent(halo, evI, 0));             // the next boundary launch reads (and overwrites the inputs of) this interior launch
        st->cur = 1 - st->cur;
    }
    MOKAB_CUDA(cudaEventRecord(D.events[17], halo));
    MOKAB_CUDA(cudaStreamWaitEvent(compute, D.events[17], 0));
}

static void decomp_drop_graphs(mokab_state *st)
{
    mokab_state::Decomp &D = st->dec;
    for (int k = 0; k < 2; ++k)
        for (int p = 0; p < 2; ++p)
            for (int n = 0; n < 2; ++n)
                if (D.graph[k][p][n]) { cudaGraphExecDestroy(D.graph[k][p][n]); D.graph[k][p][n] = nullptr; }
    D.graph_dt[0] = D.graph_dt[1] = 0.0;
    D.graphs_ready[0] = D.graphs_ready[1] = false;
}

// kind 0 = RungeKutta4, 1 = ForwardEuler; graph[kind][parity][0] = one step, [1] = two steps
template <class R>
static void decomp_build_graphs(mokab_state *st, double dt, int kind)
{
    mokab_state::Decomp &D = st->dec;
    mokab_ctx *ctx = st->ctx;
    for (int p = 0; p < 2; ++p)
        for (int n = 0; n < 2; ++n)
            if (D.graph[kind][p][n]) { cudaGraphExecDestroy(D.graph[kind][p][n]); D.graph[kind][p][n] = nullptr; }
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    MOKAB_CUDA(cudaStreamSynchronize(D.halo));
    const int64_t saved_launches = ctx->launches;
    const int saved_cur = st->cur;
    for (int p = 0; p < 2; ++p)
        for (int n = 0; n < 2; ++n) {
            cudaGraph_t g = nullptr;
            st->cur = p;
            ctx->launches = 0;
            MOKAB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            try {
                if (kind == 0) decomp_enqueue_rk4<R>(st, dt, n + 1);
                else decomp_enqueue_fe(st, dt, n + 1);
            } catch (...) {
                cudaStreamEndCapture(ctx->stream, &g);
                if (g) cudaGraphDestroy(g);
                st->cur = saved_cur;
                ctx->launches = saved_launches;
                throw;
            }
            D.graph_launches[kind][n] = ctx->launches;
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            st->cur = saved_cur;
            if (e != cudaSuccess) { ctx->launches = saved_launches; MOKAB_CUDA(e); }
            cudaGraphExec_t ge = nullptr;
            // (per-node priorities -- the halo stream's launches carry theirs, moka_b200.cu: launch_ex -- count only in graphs
            //  instantiated with this flag; without it every node runs at the priority of the stream the graph is launched into)
#ifdef MOKAB_SIM
            e = cudaGraphInstantiate(&ge, g, 0);
#else
            e = cudaGraphInstantiate(&ge, g, options().launch_priority ? cudaGraphInstantiateFlagUseNodePriority : 0);
#endif
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { ctx->launches = saved_launches; MOKAB_CUDA(e); }
            D.graph[kind][p][n] = ge;
        }
    ctx->launches = saved_launches;                                // capture enqueues are not executions
    D.graph_dt[kind] = dt;
    D.graph_epoch[kind] = options().epoch;
    D.graphs_ready[kind] = true;
}

// everything the stage launches build lazily (derived mesh arrays, the weight copies of the kernel variant in use, shared-memory
// opt-ins): done here, OUTSIDE any stream capture -- these allocate, launch on the context's stream and synchronise
template <class R>
static void decomp_prepare(mokab_state *st)
{
    mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    if (st->K > 1) {
        ensure_level_buffers(st);
        // the stage kernels gather ssh of the state they read: make ssh[cur] what the current layerThickness implies, halo copies
        // included (every later step's ssh is written by the stage that produces its layerThickness, and exchanged with it)
        update_ssh(st, st->d->h[st->cur].p, st->d->ssh[st->cur].p);
    }
    ensure_fused<R>(m);
    stage_tma_prepare<R>();
    ensure_wf_block_major<R>(m);
    ensure_wf_interleaved<R>(m);
}

// what the reverse sweep needs of the input of the step about to run: RungeKutta4 (u, h) -- halo copies included, the stage
// states are recomputed from them --, ForwardEuler (u, the lagged layerThicknessEdge)
template <class R>
static void decomp_tape_record(mokab_state *st, double dt, int kind)
{
    StateT<R> *t = typed<R>(st);
    const mokab_mesh *m = st->mesh;
    cudaStream_t s = st->ctx->stream;
    if (kind == 1) {
        if constexpr (sizeof(R) == 8) fe_tape_record(st, dt, t->u[st->cur].p, t->hE[st->cur].p);
        return;
    }
    const size_t k = t->tapeDt.size(), K = (size_t)st->K;
    MOKAB_CUDA(cudaMemcpyAsync(t->tapeU.p + k * K * m->nE, t->u[st->cur].p, K * m->nE * sizeof(R), cudaMemcpyDeviceToDevice, s));
    MOKAB_CUDA(cudaMemcpyAsync(t->tapeH.p + k * K * m->nC, t->h[st->cur].p, K * m->nC * sizeof(R), cudaMemcpyDeviceToDevice, s));
    t->tapeDt.push_back(dt);
}

template <class R>
static void decomp_run(mokab_state *st, double dt, int64_t nsteps, int kind)
{
    mokab_state::Decomp &D = st->dec;
    mokab_ctx *ctx = st->ctx;
    if (typed<R>(st)->taping) {
        StateT<R> *t = typed<R>(st);
        MOKAB_REQUIRE((int64_t)t->tapeDt.size() + nsteps <= t->tapeCap, "timestep_*_decomposed: the tape is full (mokab_tape_begin max_steps)");
        MOKAB_REQUIRE(t->tapeKind == 0 || t->tapeKind == kind + 1, "timestep_*_decomposed: the tape already holds steps of the other stepper");
        t->tapeKind = kind + 1;             // (also a call with nsteps = 0: the seed then follows this stepper's state definition)
        if (kind == 0 && nsteps > 0 && t->tapeH.n < (size_t)t->tapeCap * (size_t)st->K * st->mesh->nC) t->tapeH.alloc((size_t)t->tapeCap * (size_t)st->K * st->mesh->nC);
    }
    if (nsteps <= 0) return;
    decomp_prepare<R>(st);
    if (!D.use_graph()) {
        if (typed<R>(st)->taping) {
            for (int64_t i = 0; i < nsteps; ++i) {
                decomp_tape_record<R>(st, dt, kind);
                if (kind == 0) decomp_enqueue_rk4<R>(st, dt, 1); else decomp_enqueue_fe(st, dt, 1);
            }
            return;
        }
        if (kind == 0) decomp_enqueue_rk4<R>(st, dt, nsteps); else decomp_enqueue_fe(st, dt, nsteps);
        return;
    }
    if (!D.graphs_ready[kind] || D.graph_dt[kind] != dt || D.graph_epoch[kind] != options().epoch) decomp_build_graphs<R>(st, dt, kind);
    if (typed<R>(st)->taping) {   // reverse mode: what the adjoint needs of the state before every step goes on the tape, one step per replay
        for (int64_t i = 0; i < nsteps; ++i) {
            decomp_tape_record<R>(st, dt, kind);
            MOKAB_CUDA(cudaGraphLaunch(D.graph[kind][st->cur][0], ctx->stream));
            ctx->launches += D.graph_launches[kind][0];
            st->cur = 1 - st->cur;
        }
        return;
    }
    int64_t left = nsteps;
    while (left >= 2) {                                            // (a two-step graph leaves the time-level parity where it was)
        MOKAB_CUDA(cudaGraphLaunch(D.graph[kind][st->cur][1], ctx->stream));
        ctx->launches += D.graph_launches[kind][1];
        left -= 2;
    }
    if (left) {
        MOKAB_CUDA(cudaGraphLaunch(D.graph[kind][st->cur][0], ctx->stream));
        ctx->launches += D.graph_launches[kind][0];
        st->cur = 1 - st->cur;
    }
}

// ---- set-up ---------------------------------------------------------------------------------------------------------------
// the one-off address / index exchange of the direct-store halo path (what the Python mirror did through torch.distributed)
template <class R>
static void decomp_setup_p2p(mokab_state *st)
{
    mokab_state::Decomp &D = st->dec;
    mokab_comm *c = D.comm;
    const mokab_mesh *m = st->mesh;
    const int n = c->nranks;
    // where my halo entities live in MY arrays, message order; rank q fills my segment q and needs those indices
    std::vector<int32_t> mine(std::max<size_t>(m->haloRecv.n, 1));
    {
        std::vector<int32_t> r(m->haloRecv.n);
        if (!r.empty()) {
            MOKAB_CUDA(cudaStreamSynchronize(m->ctx->stream));
            MOKAB_CUDA(cudaMemcpy(r.data(), m->haloRecv.p, r.size() * 4, cudaMemcpyDeviceToHost));
        }
        for (size_t k = 0; k < r.size(); ++k) mine[k] = r[k] < m->nC ? r[k] : -(int32_t)(r[k] - m->nC) - 1;
    }
    std::vector<int64_t> sb(n), rb(n);
    int64_t total_send = 0;
    for (int q = 0; q < n; ++q) { sb[q] = D.rcnt[q] * 4; rb[q] = D.scnt[q] * 4; total_send += D.scnt[q]; }
    std::vector<int32_t> got(std::max<int64_t>(total_send, 1));
    comm::exchange_host_v(c, mine.data(), got.data(), sb, rb);
    // everybody's addresses / IPC handles
    P2PBlob blob;
    p2p_export<R>(st, &blob);
    blob.rank = c->rank;
    std::vector<P2PBlob> send((size_t)n, blob), blobs((size_t)n);
    comm::exchange_host(c, send.data(), blobs.data(), sizeof(P2PBlob));
    // Symmetric peer relation (kernels_p2p.cuh): every rank ticks, and waits for, every rank it exchanges anything with in
    // EITHER direction.  At a corner of the decomposition a rank can own halo edges of a neighbour without holding any of that
    // neighbour's entities; with one-directional lists it would push to it and never wait for it, could run two stages ahead
    // of it, and overwrite halo slots the neighbour is still reading.  Zero-length pushes still tick.
    std::vector<int32_t> peers;
    std::vector<int64_t> counts;
    std::vector<int32_t> dst;
    {
        int64_t off = 0;
        for (int q = 0; q < n; ++q) {
            if (D.scnt[q] > 0 || D.rcnt[q] > 0) {
                peers.push_back(q);
                counts.push_back(D.scnt[q]);
                dst.insert(dst.end(), got.begin() + off, got.begin() + off + D.scnt[q]);
            }
            off += D.scnt[q];
        }
    }
    dst.push_back(0);
    p2p_setup<R>(st, c->rank, n, blobs.data(), (int)peers.size(), peers.data(), counts.data(), dst.data(), (int)peers.size(), peers.data());
    if (D.mode == MOKAB_HALO_P2P_LL) {
        // flag-in-data exchange: every rank tells each peer where that peer's segment starts in its receive list and which slot it
        // keeps for the peer's credit packet (behind the list, in peer order)
        std::vector<int64_t> tell((size_t)2 * n, -1), told((size_t)2 * n, -1);
        {
            int64_t off = 0;
            int at = 0;
            for (int q = 0; q < n; ++q) {
                if (D.scnt[q] > 0 || D.rcnt[q] > 0) { tell[2 * q] = off; tell[2 * q + 1] = (int64_t)m->haloRecv.n + at; ++at; }
                off += D.rcnt[q];
            }
        }
        comm::exchange_host(c, tell.data(), told.data(), 2 * sizeof(int64_t));
        std::vector<int64_t> base, credit;
        for (int32_t q : peers) { base.push_back(told[2 * (size_t)q]); credit.push_back(told[2 * (size_t)q + 1]); }
        base.push_back(0); credit.push_back(0);
        p2p_setup_ll<R>(st, counts.data(), base.data(), credit.data());
    }
    double one = 1.0;
    comm::allreduce_f64(c, &one, 1, 0);                            // nobody pushes before everybody is mapped
}

template <class R>
static void decomp_setup(mokab_state *st, mokab_comm *c, const int64_t *scnt, const int64_t *rcnt, int mode, unsigned flags)
{
    mokab_state::Decomp &D = st->dec;
    const mokab_mesh *m = st->mesh;
    MOKAB_REQUIRE(!D.ready, "decomp_setup: the state is already set up (mokab_decomp_close first)");
    MOKAB_REQUIRE(m->halo_ready, "decomp_setup: call mokab_halo_setup on the mesh first");
    MOKAB_REQUIRE(c->ctx == st->ctx, "decomp_setup: the communicator belongs to a different context");
    int64_t ns = 0, nr = 0;
    for (int q = 0; q < c->nranks; ++q) {
        MOKAB_REQUIRE(scnt[q] >= 0 && rcnt[q] >= 0, "decomp_setup: negative count");
        ns += scnt[q]; nr += rcnt[q];
    }
    MOKAB_REQUIRE(scnt[c->rank] == 0 && rcnt[c->rank] == 0, "decomp_setup: a rank does not exchange halo entities with itself");
    MOKAB_REQUIRE(ns == (int64_t)m->haloSend.n && nr == (int64_t)m->haloRecv.n,
                  "decomp_setup: the counts do not add up to the send / recv lists of mokab_halo_setup");
    D.comm = c; D.mode = mode; D.flags = flags;
    D.scnt.assign(scnt, scnt + c->nranks); D.rcnt.assign(rcnt, rcnt + c->nranks);
    D.sendBuf.alloc((size_t)std::max<int64_t>(ns, 1) * sizeof(R)); D.recvBuf.alloc((size_t)std::max<int64_t>(nr, 1) * sizeof(R));
    {
        std::vector<int32_t> so((size_t)c->nranks + 1, 0), ro((size_t)c->nranks + 1, 0);
        for (int q = 0; q < c->nranks; ++q) { so[q + 1] = so[q] + (int32_t)scnt[q]; ro[q + 1] = ro[q] + (int32_t)rcnt[q]; }
        D.sendOff.upload(so, st->ctx->stream); D.recvOff.upload(ro, st->ctx->stream);
    }
    D.sendBuf.zero(st->ctx->stream); D.recvBuf.zero(st->ctx->stream);
    int lo = 0, hi = 0;
#ifndef MOKAB_SIM
    MOKAB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));        // (hi is the numerically lowest = highest priority)
#endif
    MOKAB_CUDA(cudaStreamCreateWithPriority(&D.halo, cudaStreamNonBlocking, hi));
    (void)lo;
    decomp_events_reserve(st, 18);
    MOKAB_CUDA(cudaStreamSynchronize(st->ctx->stream));
    // one exchange of the message buffers outside any capture: NCCL sets its connections up on first use
    comm::all_to_all(c, D.halo, D.sendBuf.p, D.recvBuf.p, D.scnt.data(), D.rcnt.data(), sizeof(R));
    MOKAB_CUDA(cudaStreamSynchronize(D.halo));
    D.ready = true;
    if (mode != MOKAB_HALO_NCCL && st->K == 1) decomp_setup_p2p<R>(st);   // (multi-level states always use the packed exchange)
}

static void decomp_release(mokab_state *st)
{
    mokab_state::Decomp &D = st->dec;
    decomp_drop_graphs(st);
    for (cudaEvent_t e : D.events) cudaEventDestroy(e);
    D.events.clear();
    if (D.halo) { cudaStreamSynchronize(D.halo); cudaStreamDestroy(D.halo); D.halo = nullptr; }
    D.sendBuf.release(); D.recvBuf.release();
    D.sendBufML.release(); D.recvBufML.release(); D.sendOff.release(); D.recvOff.release();
    D.ready = false;
    D.comm = nullptr;
}

static void decomp_check_p2p_error(mokab_state *st)
{
    if (!st->p2p.exported) return;
    int err = 0;
    MOKAB_CUDA(cudaMemcpy(&err, st->p2p.error.p, sizeof(int), cudaMemcpyDeviceToHost));
    MOKAB_REQUIRE(err == 0, "a halo wait timed out (a peer died or the ranks' schedules diverged; MOKAB_P2P_TIMEOUT_S sets the limit)");
}

}  // namespace mokab

using namespace mokab;

extern "C" {

// ---- communicator ---------------------------------------------------------------------------------------------------------
int mokab_comm_get_unique_id(void *id_out)
{
    return guarded([&] {
        MOKAB_REQUIRE(id_out, "comm_get_unique_id: NULL argument");
        memset(id_out, 0, comm::kIdBytes);
#ifdef MOKAB_SIM
        comm::SimRegistry &R = comm::SimRegistry::get();
        std::lock_guard<std::mutex> lk(R.mu);
        const uint64_t token = R.next++;
        memcpy(id_out, &token, sizeof(token));
#else
        static_assert(sizeof(ncclUniqueId) == comm::kIdBytes, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        MOKAB_NCCL(comm::Nccl::get().GetUniqueId(&id));
        memcpy(id_out, &id, sizeof(id));
#endif
    });
}

int mokab_comm_init(mokab_ctx *ctx, const void *id, int rank, int nranks, mokab_comm **out)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && id && out, "comm_init: NULL argument");
        MOKAB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad rank / nranks");
        ctx->bind();
        auto *c = new mokab_comm();
        c->ctx = ctx; c->rank = rank; c->nranks = nranks;
        try {
#ifdef MOKAB_SIM
            memcpy(&c->token, id, sizeof(c->token));
            comm::SimRegistry &R = comm::SimRegistry::get();
            {
                std::lock_guard<std::mutex> lk(R.mu);
                auto it = R.live.find(c->token);
                if (it == R.live.end()) it = R.live.emplace(c->token, comm::SimRegistry::Entry{mokab_sim_comm_create(nranks), 0}).first;
                it->second.refs++;
                c->sim = it->second.comm;
            }
#else
            ncclUniqueId uid;
            memcpy(&uid, id, sizeof(uid));
            MOKAB_NCCL(comm::Nccl::get().CommInitRank(&c->nccl, nranks, uid, rank));
#endif
            MOKAB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        } catch (...) {
            delete c;
            throw;
        }
        *out = c;
    });
}

int mokab_comm_destroy(mokab_comm *c)
{
    return guarded([&] {
        if (!c) return;
        c->ctx->bind();
        cudaStreamSynchronize(c->stream);
#ifdef MOKAB_SIM
        comm::SimRegistry &R = comm::SimRegistry::get();
        {
            std::lock_guard<std::mutex> lk(R.mu);
            auto it = R.live.find(c->token);
            if (it != R.live.end() && --it->second.refs == 0) {
                mokab_sim_comm_destroy(it->second.comm);
                R.live.erase(it);
            }
        }
#else
        if (c->nccl) comm::Nccl::get().CommDestroy(c->nccl);
#endif
        cudaStreamDestroy(c->stream);
        delete c;
    });
}

int mokab_comm_rank(const mokab_comm *c, int *rank, int *nranks)
{
    return guarded([&] {
        MOKAB_REQUIRE(c && rank && nranks, "comm_rank: NULL argument");
        *rank = c->rank; *nranks = c->nranks;
    });
}

int mokab_comm_barrier(mokab_comm *c)
{
    return guarded([&] {
        MOKAB_REQUIRE(c, "comm_barrier: comm is NULL");
        c->ctx->bind();
        double one = 1.0;
        comm::allreduce_f64(c, &one, 1, 0);
    });
}

int mokab_comm_allreduce_f64(mokab_comm *c, double *inout, int64_t n, int op)
{
    return guarded([&] {
        MOKAB_REQUIRE(c && (inout || n == 0) && n >= 0, "comm_allreduce_f64: bad argument");
        c->ctx->bind();
        comm::allreduce_f64(c, inout, n, op);
    });
}

int mokab_comm_allgather_bytes(mokab_comm *c, const void *mine, int64_t bytes, void *all)
{
    return guarded([&] {
        MOKAB_REQUIRE(c && mine && all && bytes > 0, "comm_allgather_bytes: bad argument");
        c->ctx->bind();
        std::vector<unsigned char> send((size_t)bytes * c->nranks);
        for (int q = 0; q < c->nranks; ++q) memcpy(send.data() + (size_t)q * bytes, mine, (size_t)bytes);
        comm::exchange_host(c, send.data(), all, (size_t)bytes);
    });
}

// ---- decomposed stepping ------------------------------------------------------------------------------------------------------
int mokab_decomp_setup(mokab_state *state, mokab_comm *c, const int64_t *send_counts, const int64_t *recv_counts, int halo_mode,
                       uint32_t flags)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && c && send_counts && recv_counts, "decomp_setup: NULL argument");
        MOKAB_REQUIRE(halo_mode >= MOKAB_HALO_NCCL && halo_mode <= MOKAB_HALO_P2P_LL, "decomp_setup: unknown halo mode");
        state->ctx->bind();
        try {
            if (state->dtype == MOKAB_F64) decomp_setup<double>(state, c, send_counts, recv_counts, halo_mode, flags);
            else decomp_setup<float>(state, c, send_counts, recv_counts, halo_mode, flags);
        } catch (...) {
            decomp_release(state);
            throw;
        }
    });
}

int mokab_decomp_set_flags(mokab_state *state, uint32_t flags)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->dec.ready, "decomp_set_flags: call mokab_decomp_setup first");
        state->ctx->bind();
        if (flags != state->dec.flags) {
            MOKAB_CUDA(cudaStreamSynchronize(state->ctx->stream));
            decomp_drop_graphs(state);                              // the overlap schedule is baked into them
            state->dec.flags = flags;
        }
    });
}

int mokab_timestep_rk4_decomposed(mokab_state *state, double dt, int64_t nsteps)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->dec.ready, "timestep_rk4_decomposed: call mokab_decomp_setup first");
        MOKAB_REQUIRE(nsteps >= 0, "timestep_rk4_decomposed: nsteps must be >= 0");
        MOKAB_REQUIRE(state->K == 1 || state->dtype == MOKAB_F64, "timestep_rk4_decomposed: multi-level states are Float64");
        state->ctx->bind();
        leave_forward_euler(state);
        if (state->dtype == MOKAB_F64) { decomp_run<double>(state, dt, nsteps, 0); if (nsteps) refresh_ssh<double>(state); }
        else { decomp_run<float>(state, dt, nsteps, 0); if (nsteps) refresh_ssh<float>(state); }
    });
}

int mokab_timestep_forward_euler_decomposed(mokab_state *state, double dt, int64_t nsteps)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->dec.ready, "timestep_forward_euler_decomposed: call mokab_decomp_setup first");
        MOKAB_REQUIRE(nsteps >= 0, "timestep_forward_euler_decomposed: nsteps must be >= 0");
        require_f64(state, "timestep_forward_euler_decomposed");
        MOKAB_REQUIRE(state->K == 1, "timestep_forward_euler_decomposed: single-level states only (nVertLevels == 1)");
        state->ctx->bind();
        if (nsteps > 0) run_fe_stage_prepare(state);               // (allocations and the hand-over copy, outside any capture)
        decomp_run<double>(state, dt, nsteps, 1);
    });
}

int mokab_reduce_decomposed(mokab_state *state, int which, double *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && out && state->dec.ready, "reduce_decomposed: NULL argument or mokab_decomp_setup not called");
        state->ctx->bind();
        double v = 0.0;
        if (state->dtype == MOKAB_F64) do_reduce<double>(state, which, &v); else do_reduce<float>(state, which, &v);
        decomp_check_p2p_error(state);
        comm::allreduce_f64(state->dec.comm, &v, 1, 0);
        *out = v;
    });
}

int mokab_decomp_synchronize(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->dec.ready, "decomp_synchronize: call mokab_decomp_setup first");
        state->ctx->bind();
        MOKAB_CUDA(cudaStreamSynchronize(state->ctx->stream));
        MOKAB_CUDA(cudaStreamSynchronize(state->dec.halo));
        decomp_check_p2p_error(state);
    });
}

int mokab_decomp_close(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "decomp_close: state is NULL");
        if (!state->dec.ready) return;
        state->ctx->bind();
        MOKAB_CUDA(cudaStreamSynchronize(state->ctx->stream));
        MOKAB_CUDA(cudaStreamSynchronize(state->dec.halo));
        mokab_comm *c = state->dec.comm;
        if (state->dec.mode != MOKAB_HALO_NCCL) {                  // nobody unmaps while a neighbour may still store, nobody frees while mapped
            double one = 1.0;
            comm::allreduce_f64(c, &one, 1, 0);
            for (void *q : state->p2p.opened) MOKAB_CUDA(cudaIpcCloseMemHandle(q));
            state->p2p.opened.clear();
            state->p2p.ready = false;
            comm::allreduce_f64(c, &one, 1, 0);
        }
        decomp_release(state);
    });
}

}  // extern "C"
