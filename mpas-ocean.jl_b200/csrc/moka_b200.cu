// moka_b200.cu -- libmoka_b200.so: C ABI (include/moka_b200.h) over the sm_100a kernels.
#include "common.cuh"
#include "kernels_fused.cuh"
#include "kernels_adjoint.cuh"
#include "kernels_ref.cuh"
#include "kernels_p2p.cuh"
#include "mesh.cuh"
#include "comm.cuh"

#include <unistd.h>

namespace mokab {
thread_local std::string g_last_error;

#define LAUNCH(ctx, kernel, grid, block, ...)                          \
    do {                                                               \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);    \
        MOKAB_CUDA(cudaGetLastError());                                \
        (ctx)->launches++;                                             \
    } while (0)

static inline int nblk(int64_t n, int t = 256) { return (int)((n + t - 1) / t); }
constexpr int kP2PCounters = 1024;   // arrival counters of the direct-store halo exchange, indexed by sender rank

// ---- typed state ----------------------------------------------------------------------------------------
template <class R>
struct StateT {
    // Everything a peer GPU may write -- the halo slots of the two time levels and of the provisional buffers, and the
    // arrival counters of the direct-store halo exchange -- lives in ONE allocation, so one CUDA IPC handle maps it
    // (several small cudaMalloc blocks can share a physical chunk, which IPC cannot map twice).
    DevBuf<unsigned char> slab;
    size_t llSlots = 0;                    // slots of that receive area (0: the state was created before mokab_halo_setup)
    size_t slabOff[10] = {};               // byte offsets: u0 u1 uP0 uP1 h0 h1 hP0 hP1 counters, receive area of the flag-in-data exchange
    DevBuf<R> u[2], h[2], ssh[2];          // two time levels; `cur` holds Prog.*[end]  (u, h: views into the slab)
    DevBuf<R> uP[2], hP[2];                // RK provisional ping-pong (views into the slab)
    DevBuf<R> hEdge, flux, divC, relVort, tendU, tendH, sshProv;
    DevBuf<R> sshP[2];                     // multi-level fused RK4: ssh of the two provisional states (uP / hP)
    DevBuf<R> hE[2];                       // fused ForwardEuler: layerThicknessEdge ping-pong, indexed like the time levels
    DevBuf<R> staging;
    DevBuf<double> partial, result;
    // fused RK4 graphs: [parity] one step starting with cur == parity; pair = two steps
    cudaGraphExec_t gStep[2] = {nullptr, nullptr};
    cudaGraphExec_t gPair[2] = {nullptr, nullptr};
    double graph_dt = 0.0;
    int64_t graph_epoch = -1;              // options().epoch the graphs were captured under
    bool graphs_ready = false;
    // asynchronous host transfers (mokab_state_set_async / _get_async): rings of staging slots, one copy
    // stream per direction, events ordering copy <-> permute kernels
    static constexpr int kInSlots = 4, kOutSlots = 2;
    DevBuf<R> stIn[kInSlots], stOut[kOutSlots];
    cudaEvent_t evInCopied[kInSlots] = {}, evInFree[kInSlots] = {}, evOutReady[kOutSlots] = {}, evOutCopied[kOutSlots] = {};
    cudaStream_t h2d = nullptr, d2h = nullptr;
    int inSlot = 0, outSlot = 0;
    bool async_ready = false;
    // reverse mode (mokab_tape_* / mokab_adjoint_*): trajectory tape, stage states, adjoint ping-pong buffers
    DevBuf<R> tapeU, tapeH;                 // (tapeCap, nE), (tapeCap, nC): the state before each recorded step
    DevBuf<R> tapeE;                        // ForwardEuler: (tapeCap, nE) the lagged layerThicknessEdge each step consumed
    std::vector<double> tapeDt;             // dt of every recorded step
    int64_t tapeCap = 0;
    bool taping = false;
    int tapeKind = 0;                       // 0 = empty, 1 = RungeKutta4 steps, 2 = ForwardEuler steps (never mixed)
    DevBuf<R> lamS[2], lamE[2], lamQ[2];    // ForwardEuler adjoint: adjoints of ssh / hEdge, invArea * (lamH + lamS)
    DevBuf<R> yU[3], yH[3];                 // y_2, y_3, y_4 of the step being reversed
    DevBuf<R> yS[4], kuP;                   // multi-level states: ssh of y_1 ... y_4 (the column kernel gathers it); level sum of kbar_u
    DevBuf<R> kbU[2], kbH[2];               // kbar ping-pong (kbH carries invArea * kbar_h)
    DevBuf<R> lamU[2], lamH[2], dSsh;       // lam' / lam, swapped every reversed step; seed on ssh
    int lamCur = 0;
    bool adj_ready = false;
    void drop_async()
    {
        if (!async_ready) return;
        cudaStreamSynchronize(h2d);
        cudaStreamSynchronize(d2h);
        for (int i = 0; i < kInSlots; ++i) { cudaEventDestroy(evInCopied[i]); cudaEventDestroy(evInFree[i]); }
        for (int i = 0; i < kOutSlots; ++i) { cudaEventDestroy(evOutReady[i]); cudaEventDestroy(evOutCopied[i]); }
        cudaStreamDestroy(h2d);
        cudaStreamDestroy(d2h);
        async_ready = false;
    }
    void drop_graphs()
    {
        for (int p = 0; p < 2; ++p) {
            if (gStep[p]) cudaGraphExecDestroy(gStep[p]);
            if (gPair[p]) cudaGraphExecDestroy(gPair[p]);
            gStep[p] = gPair[p] = nullptr;
        }
        graphs_ready = false;
    }
    ~StateT()
    {
        drop_graphs();
        drop_async();
    }
};
}  // namespace mokab

struct mokab_state {
    mokab_ctx *ctx = nullptr;
    const mokab_mesh *mesh = nullptr;
    int dtype = MOKAB_F64;
    int K = 1;    // nVertLevels: layerThickness / normalVelocity (and the Diag / Tend arrays) hold K levels, level-major on the device
    int cur = 1;  // index of the time level holding Prog.*[end]
    // fused ForwardEuler leaves thicknessFlux / velocityDivCell / tend* / layerThicknessEdge to be re-created on demand
    // from the previous time level and hE[] (fe_materialize); true while those arrays are stale
    bool fe_lazy = false;
    mokab::StateT<double> *d = nullptr;
    mokab::StateT<float> *f = nullptr;
    // halo exchange by direct peer stores (kernels_p2p.cuh); set up by mokab_p2p_setup
    struct P2P {
        bool exported = false, ready = false;
        int rank = 0, nranks = 0;
        unsigned long long *arrival = nullptr;               // per sender rank: pushes arrived (in the state's slab: peers write it)
        mokab::DevBuf<unsigned long long> expect;            // per sender rank: waits completed
        mokab::DevBuf<unsigned int> done;
        mokab::DevBuf<int> error;
        std::vector<int> recvRanks, sendRanks;               // ranks this one pushes to / waits for
        mokab::DevBuf<int32_t> dst, senders;
        mokab::DevBuf<uint8_t> slot;
        mokab::DevBuf<void *> peerH, peerU;                  // [4 targets][nrecv]: hP0, hP1, h0, h1 on every receiver
        mokab::DevBuf<unsigned long long *> arrivalAt;       // [nrecv]: this rank's slot in the receiver's arrival array
        std::vector<void *> opened;                          // cudaIpcOpenMemHandle mappings to close
        int64_t nPush = 0;
        // the exchange folded into the boundary launch (fused::PushStage): CSR over the local entities + one descriptor per target
        mokab::DevBuf<int32_t> startE, startC, dstE, dstC;
        mokab::DevBuf<uint8_t> slotE, slotC;
        mokab::DevBuf<unsigned char> stageDesc;              // 4 x fused::PushStage<R>
        // flag-in-data exchange (MOKAB_HALO_P2P_LL; kernels_p2p.cuh)
        std::vector<unsigned long long *> peerLLHost;        // [nrecv]: the receivers' receive areas (mapped)
        std::vector<int64_t> peerLLSlots;                    //          and how many slots they have
        mokab::DevBuf<unsigned long long *> peerLL;
        mokab::DevBuf<int32_t> llDst;                        // per item of the send list (+ one credit per peer): its slot at the receiver
        mokab::DevBuf<uint8_t> llSlot;                       //                                                   which receiver
        mokab::DevBuf<unsigned int> llCtr;                   // [0] exchanges sent, [1] exchanges received, [2] / [3] the two ticket counters
        unsigned long long *llArea = nullptr;                // this rank's receive area (in the slab)
        int llSend = 0, llRecv = 0;                          // items sent / slots waited for per exchange (credits included)
        bool llReady = false;
    } p2p;
    // domain-decomposed stepping inside the library (csrc/decomposed.cuh); set up by mokab_decomp_setup
    struct Decomp {
        bool ready = false;
        mokab_comm *comm = nullptr;
        int mode = MOKAB_HALO_NCCL;
        uint32_t flags = 0;
        std::vector<int64_t> scnt, rcnt;                     // halo elements sent to / received from every rank
        mokab::DevBuf<unsigned char> sendBuf, recvBuf;       // packed messages (state dtype)
        mokab::DevBuf<unsigned char> sendBufML, recvBufML;   // multi-level states: K + 1 planes per message (allocated on first use)
        mokab::DevBuf<int32_t> sendOff, recvOff;             // first item of every rank's segment in the send / recv lists (nranks + 1)
        cudaStream_t halo = nullptr;                         // high-priority stream: boundary blocks + the exchange
        std::vector<cudaEvent_t> events;
        cudaGraphExec_t graph[2][2][2] = {};                 // [RungeKutta4 | ForwardEuler][time-level parity][one | two steps]
        int64_t graph_launches[2][2] = {};
        double graph_dt[2] = {0.0, 0.0};
        int64_t graph_epoch[2] = {-1, -1};
        bool graphs_ready[2] = {false, false};
        bool overlap() const { return !(flags & MOKAB_DECOMP_NO_OVERLAP); }
        bool use_graph() const { return !(flags & MOKAB_DECOMP_NO_GRAPH); }
    } dec;
    ~mokab_state()
    {
        for (int k = 0; k < 2; ++k)
            for (int p = 0; p < 2; ++p)
                for (int n = 0; n < 2; ++n)
                    if (dec.graph[k][p][n]) cudaGraphExecDestroy(dec.graph[k][p][n]);
        for (cudaEvent_t e : dec.events) cudaEventDestroy(e);
        if (dec.halo) cudaStreamDestroy(dec.halo);
        for (void *q : p2p.opened) cudaIpcCloseMemHandle(q);
        delete d;
        delete f;
    }
};

namespace mokab {

template <class R> static StateT<R> *typed(mokab_state *s);
template <> StateT<double> *typed<double>(mokab_state *s) { return s->d; }
template <> StateT<float> *typed<float>(mokab_state *s) { return s->f; }
template <class R> static FusedMesh<R> &fused_of(mokab_mesh *m);
template <> FusedMesh<double> &fused_of<double>(mokab_mesh *m) { return m->f64; }
template <> FusedMesh<float> &fused_of<float>(mokab_mesh *m) { return m->f32; }

// ---- mesh upload ----------------------------------------------------------------------------------------
static void upload_mesh(mokab_ctx *ctx, HostMesh &hm, mokab_mesh *m)
{
    cudaStream_t s = ctx->stream;
    m->ctx = ctx;
    m->nC = hm.nC; m->nE = hm.nE; m->nV = hm.nV; m->S = hm.S; m->S2 = hm.S2; m->D = hm.D;
    std::vector<int2> ce2(hm.nE);
    for (int64_t e = 0; e < hm.nE; ++e) ce2[e] = make_int2(hm.ce[2 * e], hm.ce[2 * e + 1]);
    m->ce.upload(ce2, s);
    m->eoe.upload(hm.eoe, s); m->woe.upload(hm.woe, s); m->nEoE.upload(hm.nEoE, s);
    m->dc.upload(hm.dc, s); m->dv.upload(hm.dv, s); m->fE.upload(hm.fE, s);
    m->eoc.upload(hm.eoc, s); m->sgnC.upload(hm.sgnC, s); m->nEoC.upload(hm.nEoC, s);
    m->area.upload(hm.area, s); m->H.upload(hm.H, s);
    m->eov.upload(hm.eov, s); m->sgnV.upload(hm.sgnV, s); m->areaTri.upload(hm.areaTri, s);
    m->dPermC.upload(hm.permC, s); m->dPermE.upload(hm.permE, s); m->dPermV.upload(hm.permV, s);
    m->nCo = hm.nCo; m->nEo = hm.nEo; m->uniformF = hm.uniformF; m->f0 = hm.f0;
    m->blkEdgeStart.upload(hm.blkEdgeStart, s);
    m->blkInterior.upload(hm.blkInterior, s);
    m->blkBoundary.upload(hm.blkBoundary, s);
    m->posE.upload(hm.posE, s);
    m->blkDerived.upload(hm.blkDerived, s);
    m->nDerivedBlocks = 0;
    for (uint8_t d : hm.blkDerived) m->nDerivedBlocks += d;
    m->fusedBlocks = (int)hm.blkEdgeStart.size() - 1;
    m->maxBlockEdges = 0;
    for (int b = 0; b < m->fusedBlocks; ++b) m->maxBlockEdges = std::max(m->maxBlockEdges, hm.blkEdgeStart[b + 1] - hm.blkEdgeStart[b]);
    m->nInterior = (int)hm.blkInterior.size();
    m->nBoundary = (int)hm.blkBoundary.size();
    m->hBlkEdgeStart.swap(hm.blkEdgeStart); m->hBlkInterior.swap(hm.blkInterior); m->hBlkBoundary.swap(hm.blkBoundary);
    m->permC.swap(hm.permC); m->permE.swap(hm.permE); m->permV.swap(hm.permV);
    m->hHaloEoe.swap(hm.haloEoe); m->hHaloWoe.swap(hm.haloWoe); m->haloS2 = hm.haloS2;
    MOKAB_CUDA(cudaStreamSynchronize(s));
}

template <class R>
static void ensure_fused(mokab_mesh *m)
{
    FusedMesh<R> &f = fused_of<R>(m);
    if (f.ready) return;
    mokab_ctx *ctx = m->ctx;
    const bool need_idx = m->eoeF.n == 0;
    if (need_idx) {
        m->eoeF.alloc((size_t)m->S2 * m->nE);
        m->eocF.alloc((size_t)m->S * m->nC);
    }
    f.gdc.alloc(m->nE); f.dv.alloc(m->nE);
    f.wf.alloc((size_t)m->S2 * m->nE + 16);       // + padding: the bulk copies of the TMA stage variant round their rows up to 16 bytes
    f.wf.zero(ctx->stream);
    f.invArea.alloc(m->nC); f.H.alloc(m->nC);
    LAUNCH(ctx, fused::k_build_fused_edges<R>, nblk(m->nE), 256, (int)m->nE, m->S2, m->dc.p, m->dv.p, m->fE.p, m->eoe.p,
           m->woe.p, m->nEoE.p, f.gdc.p, f.dv.p, f.wf.p, need_idx ? m->eoeF.p : nullptr, m->uniformF ? 0 : 1);
    LAUNCH(ctx, fused::k_build_fused_cells<R>, nblk(m->nC), 256, (int)m->nC, m->S, m->area.p, m->H.p, m->eoc.p, m->sgnC.p,
           m->nEoC.p, f.invArea.p, f.H.p, need_idx ? m->eocF.p : nullptr);
    // The staged entry points (mokab_rk4_stage / mokab_refresh_ssh) launch on CALLER streams that are not ordered
    // with the context's stream: the arrays built above must be complete before anybody can read them.  One host
    // synchronisation per (mesh, precision); mokab_state_create builds them up front so it never lands in a step.
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    f.ready = true;
}

// ---- state helpers ------------------------------------------------------------------------------------------
template <class R>
static void alloc_state(mokab_state *st)
{
    const mokab_mesh *m = st->mesh;
    cudaStream_t s = st->ctx->stream;
    auto *t = new StateT<R>();
    if (sizeof(R) == 8) st->d = (StateT<double> *)(void *)t; else st->f = (StateT<float> *)(void *)t;
    {
        auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t K = (size_t)st->K;
        const size_t eb = pad(K * m->nE * sizeof(R)), cb = pad(K * m->nC * sizeof(R));
        size_t off = 0;
        for (int i = 0; i < 4; ++i) { t->slabOff[i] = off; off += eb; }
        for (int i = 4; i < 8; ++i) { t->slabOff[i] = off; off += cb; }
        t->slabOff[8] = off; off += pad((size_t)kP2PCounters * sizeof(unsigned long long));
        // MOKAB_HALO_P2P_LL: 32 bytes (two parities x two packets) per entry of the receive list + one credit slot per possible peer;
        // sized from the halo lists the mesh has NOW (mokab_halo_setup before mokab_state_create, as every caller here does)
        t->llSlots = m->halo_ready && K == 1 ? (size_t)m->haloRecv.n + (size_t)p2p::kMaxPeers : 0;
        t->slabOff[9] = off; off += pad(t->llSlots * 32);
        t->slab.alloc(off); t->slab.zero(s);
        unsigned char *b = t->slab.p;
        t->u[0].view((R *)(b + t->slabOff[0]), K * m->nE); t->u[1].view((R *)(b + t->slabOff[1]), K * m->nE);
        t->uP[0].view((R *)(b + t->slabOff[2]), K * m->nE); t->uP[1].view((R *)(b + t->slabOff[3]), K * m->nE);
        t->h[0].view((R *)(b + t->slabOff[4]), K * m->nC); t->h[1].view((R *)(b + t->slabOff[5]), K * m->nC);
        t->hP[0].view((R *)(b + t->slabOff[6]), K * m->nC); t->hP[1].view((R *)(b + t->slabOff[7]), K * m->nC);
    }
    const size_t K = (size_t)st->K;
    for (int l = 0; l < 2; ++l) { t->ssh[l].alloc(m->nC); t->ssh[l].zero(s); }
    t->hEdge.alloc(K * m->nE); t->hEdge.zero(s);      // DiagnosticVars.jl:90-93
    t->flux.alloc(K * m->nE); t->flux.zero(s);
    t->divC.alloc(K * m->nC); t->divC.zero(s);
    t->relVort.alloc(K * std::max<int64_t>(m->nV, 1)); t->relVort.zero(s);
    t->tendU.alloc(K * m->nE); t->tendU.zero(s);      // TendencyVars.jl:61-62
    t->tendH.alloc(K * m->nC); t->tendH.zero(s);
    t->sshProv.alloc(m->nC); t->sshProv.zero(s);
    if (K > 1) for (int l = 0; l < 2; ++l) { t->sshP[l].alloc(m->nC); t->sshP[l].zero(s); }
    t->staging.alloc(K * std::max(std::max(m->nE, m->nC), std::max<int64_t>(m->nV, 1)));
    t->partial.alloc(reduce::kBlocks); t->result.alloc(1);
    MOKAB_CUDA(cudaStreamSynchronize(s));
}

// pert: a Float32 layerThickness -- the caller sees the whole thickness, the device array holds h - H (kernels_fused.cuh: kPert);
// alias: Float32 ssh -- the same perturbation is the prognostic variable of the fused path, a `set` also writes it there
struct FieldRef { void *p; int64_t n; const int32_t *perm; bool prognostic_end; int prev_of; bool pert = false; void *alias = nullptr; int levels = 1; };

// shadow state d_Prog (ocn_init_shadows, reference src/forward/init.jl:32-40): zeros
template <class R>
static void ensure_adj_state(mokab_state *st)
{
    StateT<R> *t = typed<R>(st);
    if (t->adj_ready) return;
    const mokab_mesh *m = st->mesh;
    cudaStream_t s = st->ctx->stream;
    const size_t K = (size_t)st->K;
    for (int i = 0; i < 3; ++i) { t->yU[i].alloc(K * m->nE); t->yH[i].alloc(K * m->nC); }
    for (int i = 0; i < 2; ++i) {
        t->kbU[i].alloc(K * m->nE); t->kbH[i].alloc(K * m->nC);
        t->lamU[i].alloc(K * m->nE); t->lamU[i].zero(s);
        t->lamH[i].alloc(K * m->nC); t->lamH[i].zero(s);
    }
    if (K > 1) {
        for (int i = 0; i < 4; ++i) t->yS[i].alloc(m->nC);
        t->kuP.alloc(m->nE);
    }
    t->dSsh.alloc(m->nC); t->dSsh.zero(s);
    t->adj_ready = true;
}

template <class R> static FieldRef field_ref1(mokab_state *st, int field);
// a field as the caller sees it: `n` entities x `levels` levels (ssh and its shadow have one level whatever nVertLevels is)
template <class R>
static FieldRef field_ref(mokab_state *st, int field)
{
    FieldRef f = field_ref1<R>(st, field);
    const bool single = field == MOKAB_SSH || field == MOKAB_SSH_PREV || field == MOKAB_D_SSH;
    f.levels = single ? 1 : st->K;
    return f;
}
template <class R>
static FieldRef field_ref1(mokab_state *st, int field)
{
    StateT<R> *t = typed<R>(st);
    const mokab_mesh *m = st->mesh;
    const int c = st->cur, o = 1 - st->cur;
    constexpr bool P = fused::kPert<R>;
    switch (field) {
    case MOKAB_SSH: return {t->ssh[c].p, m->nC, m->dPermC.p, true, field, false, P ? t->h[c].p : nullptr};
    case MOKAB_NORMAL_VELOCITY: return {t->u[c].p, m->nE, m->dPermE.p, true, field};
    case MOKAB_LAYER_THICKNESS: return {t->h[c].p, m->nC, m->dPermC.p, true, field, P};
    case MOKAB_SSH_PREV: return {t->ssh[o].p, m->nC, m->dPermC.p, false, 0, false, P ? t->h[o].p : nullptr};
    case MOKAB_NORMAL_VELOCITY_PREV: return {t->u[o].p, m->nE, m->dPermE.p, false, 0};
    case MOKAB_LAYER_THICKNESS_PREV: return {t->h[o].p, m->nC, m->dPermC.p, false, 0, P};
    case MOKAB_LAYER_THICKNESS_EDGE: return {t->hEdge.p, m->nE, m->dPermE.p, false, 0};
    case MOKAB_THICKNESS_FLUX: return {t->flux.p, m->nE, m->dPermE.p, false, 0};
    case MOKAB_VELOCITY_DIV_CELL: return {t->divC.p, m->nC, m->dPermC.p, false, 0};
    case MOKAB_RELATIVE_VORTICITY: return {t->relVort.p, m->nV, m->dPermV.p, false, 0};
    case MOKAB_TEND_NORMAL_VELOCITY: return {t->tendU.p, m->nE, m->dPermE.p, false, 0};
    case MOKAB_TEND_LAYER_THICKNESS: return {t->tendH.p, m->nC, m->dPermC.p, false, 0};
    case MOKAB_D_SSH: ensure_adj_state<R>(st); return {t->dSsh.p, m->nC, m->dPermC.p, false, 0};
    case MOKAB_D_NORMAL_VELOCITY: ensure_adj_state<R>(st); return {t->lamU[t->lamCur].p, m->nE, m->dPermE.p, false, 0};
    case MOKAB_D_LAYER_THICKNESS: ensure_adj_state<R>(st); return {t->lamH[t->lamCur].p, m->nC, m->dPermC.p, false, 0};
    default: throw Error("unknown field id " + std::to_string(field));
    }
}

// caller order -> device order into `f` (and its alias) from the staging buffer `src`; device order -> caller order into `dst`
template <class R>
static void launch_permute_in(mokab_state *st, const FieldRef &f, const R *src)
{
    mokab_ctx *ctx = st->ctx;
    if constexpr (fused::kPert<R>) {
        if (f.pert) {
            LAUNCH(ctx, k_permute_in_pert, nblk(f.n), 256, f.n, f.perm, src, (const double *)st->mesh->H.p, (float *)f.p);
            return;
        }
    }
    if (f.levels > 1) LAUNCH(ctx, k_permute_in_lv<R>, nblk(f.n), 256, f.n, f.levels, f.perm, src, (R *)f.p);
    else LAUNCH(ctx, k_permute_in<R>, nblk(f.n), 256, f.n, f.perm, src, (R *)f.p);
    if (f.alias) MOKAB_CUDA(cudaMemcpyAsync(f.alias, f.p, f.n * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
}
template <class R>
static void launch_permute_out(mokab_state *st, const FieldRef &f, R *dst)
{
    mokab_ctx *ctx = st->ctx;
    if constexpr (fused::kPert<R>) {
        if (f.pert) {
            LAUNCH(ctx, k_permute_out_pert, nblk(f.n), 256, f.n, f.perm, (const float *)f.p, (const double *)st->mesh->H.p, dst);
            return;
        }
    }
    if (f.levels > 1) LAUNCH(ctx, k_permute_out_lv<R>, nblk(f.n), 256, f.n, f.levels, f.perm, (const R *)f.p, dst);
    else LAUNCH(ctx, k_permute_out<R>, nblk(f.n), 256, f.n, f.perm, (const R *)f.p, dst);
}

template <class R>
static void state_set(mokab_state *st, int field, const void *host)
{
    mokab_ctx *ctx = st->ctx;
    StateT<R> *t = typed<R>(st);
    FieldRef f = field_ref<R>(st, field);
    if (f.n == 0) return;
    MOKAB_CUDA(cudaMemcpyAsync(t->staging.p, host, f.n * f.levels * sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
    launch_permute_in<R>(st, f, (const R *)t->staging.p);
    if (f.prognostic_end) {  // deepcopy into every time level, PrognosticVars.jl:49-53
        FieldRef prev = field_ref<R>(st, field + 3);
        MOKAB_CUDA(cudaMemcpyAsync(prev.p, f.p, f.n * f.levels * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
        if (prev.alias) MOKAB_CUDA(cudaMemcpyAsync(prev.alias, f.p, f.n * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller may reuse `host` immediately
}

template <class R>
static void state_get(mokab_state *st, int field, void *host)
{
    mokab_ctx *ctx = st->ctx;
    StateT<R> *t = typed<R>(st);
    FieldRef f = field_ref<R>(st, field);
    if (f.n == 0) return;
    launch_permute_out<R>(st, f, (R *)t->staging.p);
    MOKAB_CUDA(cudaMemcpyAsync(host, t->staging.p, f.n * f.levels * sizeof(R), cudaMemcpyDeviceToHost, ctx->stream));
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- pipelined host transfers -----------------------------------------------------------------------------------
// The H2D copy of a field runs on its own stream into a staging slot and only the permute kernel joins the
// context's stream, so the upload for step n+1 overlaps the kernels of step n; likewise the D2H copy of a
// result overlaps the kernels that follow it.  `host` must be page-locked and stay valid until
// mokab_state_synchronize.
template <class R>
static void ensure_async(mokab_state *st)
{
    StateT<R> *t = typed<R>(st);
    if (t->async_ready) return;
    const mokab_mesh *m = st->mesh;
    const size_t nmax = (size_t)st->K * (size_t)std::max(std::max(m->nE, m->nC), std::max<int64_t>(m->nV, 1));
    MOKAB_CUDA(cudaStreamCreateWithFlags(&t->h2d, cudaStreamNonBlocking));
    MOKAB_CUDA(cudaStreamCreateWithFlags(&t->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < StateT<R>::kInSlots; ++i) {
        t->stIn[i].alloc(nmax);
        MOKAB_CUDA(cudaEventCreateWithFlags(&t->evInCopied[i], cudaEventDisableTiming));
        MOKAB_CUDA(cudaEventCreateWithFlags(&t->evInFree[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < StateT<R>::kOutSlots; ++i) {
        t->stOut[i].alloc(nmax);
        MOKAB_CUDA(cudaEventCreateWithFlags(&t->evOutReady[i], cudaEventDisableTiming));
        MOKAB_CUDA(cudaEventCreateWithFlags(&t->evOutCopied[i], cudaEventDisableTiming));
    }
    t->async_ready = true;
}

template <class R>
static void state_set_async(mokab_state *st, int field, const void *host)
{
    mokab_ctx *ctx = st->ctx;
    StateT<R> *t = typed<R>(st);
    ensure_async<R>(st);
    FieldRef f = field_ref<R>(st, field);
    if (f.n == 0) return;
    const int slot = t->inSlot;
    t->inSlot = (slot + 1) % StateT<R>::kInSlots;
    MOKAB_CUDA(cudaStreamWaitEvent(t->h2d, t->evInFree[slot], 0));      // the slot's previous permute has run
    MOKAB_CUDA(cudaMemcpyAsync(t->stIn[slot].p, host, f.n * f.levels * sizeof(R), cudaMemcpyHostToDevice, t->h2d));
    MOKAB_CUDA(cudaEventRecord(t->evInCopied[slot], t->h2d));
    MOKAB_CUDA(cudaStreamWaitEvent(ctx->stream, t->evInCopied[slot], 0));
    launch_permute_in<R>(st, f, (const R *)t->stIn[slot].p);
    MOKAB_CUDA(cudaEventRecord(t->evInFree[slot], ctx->stream));
}

template <class R>
static void state_get_async(mokab_state *st, int field, void *host)
{
    mokab_ctx *ctx = st->ctx;
    StateT<R> *t = typed<R>(st);
    ensure_async<R>(st);
    FieldRef f = field_ref<R>(st, field);
    if (f.n == 0) return;
    const int slot = t->outSlot;
    t->outSlot = (slot + 1) % StateT<R>::kOutSlots;
    MOKAB_CUDA(cudaStreamWaitEvent(ctx->stream, t->evOutCopied[slot], 0));  // the slot's previous D2H has left
    launch_permute_out<R>(st, f, (R *)t->stOut[slot].p);
    MOKAB_CUDA(cudaEventRecord(t->evOutReady[slot], ctx->stream));
    MOKAB_CUDA(cudaStreamWaitEvent(t->d2h, t->evOutReady[slot], 0));
    MOKAB_CUDA(cudaMemcpyAsync(host, t->stOut[slot].p, f.n * f.levels * sizeof(R), cudaMemcpyDeviceToHost, t->d2h));
    MOKAB_CUDA(cudaEventRecord(t->evOutCopied[slot], t->d2h));
}

template <class R>
static void state_synchronize(mokab_state *st)
{
    StateT<R> *t = typed<R>(st);
    if (t->async_ready) MOKAB_CUDA(cudaStreamSynchronize(t->h2d));
    MOKAB_CUDA(cudaStreamSynchronize(st->ctx->stream));
    if (t->async_ready) MOKAB_CUDA(cudaStreamSynchronize(t->d2h));
}

// ---- reference-order operator sequences (Float64) -----------------------------------------------------------
static void require_f64(mokab_state *st, const char *what)
{
    MOKAB_REQUIRE(st->dtype == MOKAB_F64,
                  std::string(what) + ": the reference-order path is Float64 only (PrognosticVars.jl:91-93); "
                                      "Float32 states support mokab_timestep_rk4(MOKAB_RK4_FUSED), set/get and reduce");
}

// Multi-level states (K > 1): every level is a contiguous array over the entities, and the reference's kernels treat the levels
// independently (their `k` loops, e.g. Operators.jl:15,29), so the reference-order sequences below run level by level.
static void diag_compute(mokab_state *st, const double *u, const double *h)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    for (int64_t k = 0; k < st->K; ++k) {
        const double *uk = u + k * m->nE, *hk = h + k * m->nC;
        LAUNCH(ctx, ref::k_diag_edges, nblk(m->nE), 256, (int)m->nE, m->ce.p, uk, hk, t->hEdge.p + k * m->nE, t->flux.p + k * m->nE);
        LAUNCH(ctx, ref::k_divergence_on_cell, nblk(m->nC), 256, (int)m->nC, m->eoc.p, m->sgnC.p, m->nEoC.p, m->area.p, m->dv.p, uk,
               t->divC.p + k * m->nC);
        if (m->nV)
            LAUNCH(ctx, ref::k_curl_on_vertex, nblk(m->nV), 256, (int)m->nV, m->D, m->eov.p, m->sgnV.p, m->areaTri.p, m->dc.p, uk,
                   t->relVort.p + k * m->nV);
    }
}

// Diagnostics of the given state taken at face value: hEdge and flux of the SAME state (no lag) and relativeVorticity
// zeroed before the curl accumulates into it (the line the reference has commented out, Operators.jl:135).
static void diag_consistent(mokab_state *st, const double *u, const double *h)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    if (m->nV) t->relVort.zero(ctx->stream);
    for (int64_t k = 0; k < st->K; ++k) {
        const double *uk = u + k * m->nE;
        LAUNCH(ctx, ref::k_interpolate_cell2edge, nblk(m->nE), 256, (int)m->nE, m->ce.p, h + k * m->nC, t->hEdge.p + k * m->nE);
        LAUNCH(ctx, ref::k_mul, nblk(m->nE), 256, m->nE, uk, (const double *)(t->hEdge.p + k * m->nE), t->flux.p + k * m->nE);
        LAUNCH(ctx, ref::k_divergence_on_cell, nblk(m->nC), 256, (int)m->nC, m->eoc.p, m->sgnC.p, m->nEoC.p, m->area.p, m->dv.p, uk,
               t->divC.p + k * m->nC);
        if (m->nV)
            LAUNCH(ctx, ref::k_curl_on_vertex, nblk(m->nV), 256, (int)m->nV, m->D, m->eov.p, m->sgnV.p, m->areaTri.p, m->dc.p, uk,
                   t->relVort.p + k * m->nV);
    }
}

static void tend_u(mokab_state *st, const double *ssh, const double *u)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    for (int64_t k = 0; k < st->K; ++k)      // the same pressure gradient for every level (pressure_gradient.jl:61-64), Coriolis per level
        LAUNCH(ctx, ref::k_tend_normal_velocity, nblk(m->nE), 256, (int)m->nE, m->S2, m->ce.p, m->dc.p, m->eoe.p, m->woe.p, m->nEoE.p,
               m->fE.p, ssh, u + k * m->nE, st->d->tendU.p + k * m->nE);
}

static void tend_h(mokab_state *st, const double *flux)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    for (int64_t k = 0; k < st->K; ++k)
        LAUNCH(ctx, ref::k_tend_layer_thickness, nblk(m->nC), 256, (int)m->nC, m->eoc.p, m->sgnC.p, m->nEoC.p, m->area.p, m->dv.p,
               flux + k * m->nE, st->d->tendH.p + k * m->nC);
}

// Update_ssh! (time_integration.jl:205-212) for a column of st->K levels
static void update_ssh(mokab_state *st, const double *h, double *ssh, cudaStream_t s = nullptr)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    cudaStream_t q = s ? s : ctx->stream;
    if (st->K > 1) k_update_ssh_lv<<<nblk(m->nC), 256, 0, q>>>(m->nC, st->K, h, (const double *)m->H.p, ssh);
    else k_update_ssh<double><<<nblk(m->nC), 256, 0, q>>>(m->nC, h, (const double *)m->H.p, ssh);
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

// advanceTimeLevels! (time_integration.jl:10-40): previous <- new
static void advance_time_levels(mokab_state *st)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    const int c = st->cur, o = 1 - c;
    MOKAB_CUDA(cudaMemcpyAsync(t->ssh[o].p, t->ssh[c].p, m->nC * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    MOKAB_CUDA(cudaMemcpyAsync(t->u[o].p, t->u[c].p, st->K * m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    MOKAB_CUDA(cudaMemcpyAsync(t->h[o].p, t->h[c].p, st->K * m->nC * 8, cudaMemcpyDeviceToDevice, ctx->stream));
}

// ocn_timestep(::ForwardEuler), time_integration.jl:150-193
static void step_forward_euler(mokab_state *st, double dt)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    advance_time_levels(st);
    const int c = st->cur;
    diag_compute(st, t->u[c].p, t->h[c].p);
    tend_u(st, t->ssh[c].p, t->u[c].p);
    tend_h(st, t->flux.p);
    const int64_t nEk = st->K * m->nE, nCk = st->K * m->nC;
    LAUNCH(ctx, ref::k_axpy, nblk(nEk), 256, nEk, (const double *)t->u[c].p, dt, (const double *)t->tendU.p, t->u[c].p);
    LAUNCH(ctx, ref::k_axpy, nblk(nCk), 256, nCk, (const double *)t->h[c].p, dt, (const double *)t->tendH.p, t->h[c].p);
    update_ssh(st, t->h[c].p, t->ssh[c].p);
}

// ---- fused ForwardEuler ----------------------------------------------------------------------------------------------
static bool fe_fusable(const mokab_state *st)
{
    const mokab_mesh *m = st->mesh;
    const bool widths = (m->S2 == 10 && m->S == 6) || (m->S2 == 12 && m->S == 7);   // the compile-time row widths of k_fe_step
    return st->dtype == MOKAB_F64 && st->K == 1 && widths && m->nCo == m->nC && m->nEo == m->nE;
}

// Re-create the Diag / Tend arrays the unfused step would have left behind, from the state before the last fused step
// (time level 1 - cur) and the hEdge that step consumed (hE[1 - cur]); layerThicknessEdge is the hEdge it produced.
static void fe_materialize(mokab_state *st)
{
    if (!st->fe_lazy) return;
    st->fe_lazy = false;
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    const int p = st->cur, o = 1 - p;
    LAUNCH(ctx, ref::k_mul, nblk(m->nE), 256, m->nE, (const double *)t->u[o].p, (const double *)t->hE[o].p, t->flux.p);
    LAUNCH(ctx, ref::k_divergence_on_cell, nblk(m->nC), 256, (int)m->nC, m->eoc.p, m->sgnC.p, m->nEoC.p, m->area.p, m->dv.p,
           (const double *)t->u[o].p, t->divC.p);
    tend_u(st, t->ssh[o].p, t->u[o].p);
    tend_h(st, t->flux.p);
    MOKAB_CUDA(cudaMemcpyAsync(t->hEdge.p, t->hE[p].p, m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
}

// A staged RungeKutta4 call (or anything else that runs on a caller's stream) after fused / staged ForwardEuler steps: the
// Diag / Tend arrays those steps left to be re-created are re-created first, and complete before the caller's stream goes on.
static void leave_forward_euler(mokab_state *st)
{
    if (!st->fe_lazy) return;
    fe_materialize(st);
    MOKAB_CUDA(cudaStreamSynchronize(st->ctx->stream));
}

// Record what the adjoint of a ForwardEuler step needs of its input: u and the LAGGED layerThicknessEdge (`hE_lagged`).
static void fe_tape_record(mokab_state *st, double dt, const double *u, const double *hE_lagged)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    MOKAB_REQUIRE(t->tapeKind != 1, "timestep_forward_euler: the tape already holds RungeKutta4 steps");
    MOKAB_REQUIRE((int64_t)t->tapeDt.size() < t->tapeCap, "timestep_forward_euler: the tape is full (mokab_tape_begin max_steps)");
    if (t->tapeE.n < (size_t)t->tapeCap * m->nE) t->tapeE.alloc((size_t)t->tapeCap * m->nE);
    t->tapeKind = 2;
    const size_t k = t->tapeDt.size();
    MOKAB_CUDA(cudaMemcpyAsync(t->tapeU.p + k * m->nE, u, m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    MOKAB_CUDA(cudaMemcpyAsync(t->tapeE.p + k * m->nE, hE_lagged, m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    t->tapeDt.push_back(dt);
}

static void run_fe_fused(mokab_state *st, double dt, int64_t nsteps)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh); StateT<double> *t = st->d;
    if (nsteps <= 0) return;
    ensure_fused<double>(m);
    FusedMesh<double> &fm = fused_of<double>(m);
    if (t->hE[0].n == 0) { t->hE[0].alloc(m->nE); t->hE[1].alloc(m->nE); }
    if (!st->fe_lazy)   // entering from the other entry points: the canonical layerThicknessEdge is what the next flux uses
        MOKAB_CUDA(cudaMemcpyAsync(t->hE[st->cur].p, t->hEdge.p, m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    fused::FeArgs A;
    A.nE = (int)m->nE; A.nC = (int)m->nC; A.nCown = (int)m->nCo;
    A.ce = m->ce.p; A.eoe = m->eoeF.p; A.eoc = m->eocF.p; A.nEoC = m->nEoC.p; A.blkEdgeStart = m->blkEdgeStart.p;
    A.gdc = fm.gdc.p; A.woe = m->woe.p; A.fE = m->fE.p; A.dv = m->dv.p; A.invArea = fm.invArea.p; A.H = m->H.p;
    A.dt = dt; A.f0 = m->f0; A.blockList = nullptr;
    for (int64_t i = 0; i < nsteps; ++i) {
        const int p = st->cur, q = 1 - p;
        if (t->taping) fe_tape_record(st, dt, t->u[p].p, t->hE[p].p);
        A.u = t->u[p].p; A.h = t->h[p].p; A.ssh = t->ssh[p].p; A.hEold = t->hE[p].p;
        A.uNew = t->u[q].p; A.hNew = t->h[q].p; A.sshNew = t->ssh[q].p; A.hEnew = t->hE[q].p;
        if (m->S == 6) {
            if (m->uniformF) fused::k_fe_step<10, 6, true><<<m->fusedBlocks, fused::kThreads, 0, ctx->stream>>>(A);
            else             fused::k_fe_step<10, 6, false><<<m->fusedBlocks, fused::kThreads, 0, ctx->stream>>>(A);
        } else {             // pentagons / hexagons / heptagons: rows padded to 12 / 7 (index = self, weight 0)
            if (m->uniformF) fused::k_fe_step<12, 7, true><<<m->fusedBlocks, fused::kThreads, 0, ctx->stream>>>(A);
            else             fused::k_fe_step<12, 7, false><<<m->fusedBlocks, fused::kThreads, 0, ctx->stream>>>(A);
        }
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
        if (m->nV)      // relativeVorticity accumulates step by step in the reference (Operators.jl:135)
            LAUNCH(ctx, ref::k_curl_on_vertex, nblk(m->nV), 256, (int)m->nV, m->D, m->eov.p, m->sgnV.p, m->areaTri.p, m->dc.p,
                   (const double *)t->u[p].p, t->relVort.p);
        st->cur = q;
    }
    st->fe_lazy = true;
}

// before the first staged ForwardEuler launch of a state (never inside a stream capture: it allocates and synchronises)
static void run_fe_stage_prepare(mokab_state *st)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    if (t->hE[0].n == 0) { t->hE[0].alloc(m->nE); t->hE[1].alloc(m->nE); }
    if (!st->fe_lazy) {       // first staged step: the canonical layerThicknessEdge is what the first flux uses
        MOKAB_CUDA(cudaMemcpyAsync(t->hE[st->cur].p, t->hEdge.p, m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));          // the launches that follow may be on other streams
        st->fe_lazy = true;
    }
}

// One ForwardEuler step of a domain-decomposed run, staged like mokab_rk4_stage: the selected blocks read time level cur and
// write level 1 - cur for their OWNED cells and edges (u, h, ssh and the layerThicknessEdge of the state they read, which the
// next step's flux uses -- the reference's lag, DiagnosticVars.jl:108-117).  The caller then exchanges the halo copies of all
// four arrays (mokab_halo_pack/unpack stage 4: (h, u)[new]; stage 5: (ssh, layerThicknessEdge)[new]) and calls
// mokab_forward_euler_finish_step.  Same kernel, same bits as the single-domain fused step.
static void run_fe_stage(mokab_state *st, double dt, int part, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh); StateT<double> *t = st->d;
    const bool widths = (m->S2 == 10 && m->S == 6) || (m->S2 == 12 && m->S == 7);
    MOKAB_REQUIRE(widths, "forward_euler_stage: needs connectivity rows of at most (10, 6) or (12, 7) entries");
    MOKAB_REQUIRE(!t->taping || st->dec.ready, "forward_euler_stage: a tape is recorded by mokab_timestep_forward_euler_decomposed, not by staged calls");
    MOKAB_REQUIRE(m->nV == 0, "forward_euler_stage: local meshes carry no vertices (relativeVorticity is not advanced)");
    ensure_fused<double>(m);
    FusedMesh<double> &fm = fused_of<double>(m);
    cudaStream_t s = stream ? stream : ctx->stream;
    run_fe_stage_prepare(st);
    int grid = m->fusedBlocks;
    fused::FeArgs A;
    A.blockList = nullptr;
    if (part == MOKAB_PART_INTERIOR) { grid = m->nInterior; A.blockList = m->blkInterior.p; }
    if (part == MOKAB_PART_BOUNDARY) { grid = m->nBoundary; A.blockList = m->blkBoundary.p; }
    if (grid == 0) return;
    const int p = st->cur, q = 1 - p;
    A.nE = (int)m->nE; A.nC = (int)m->nC; A.nCown = (int)m->nCo;
    A.ce = m->ce.p; A.eoe = m->eoeF.p; A.eoc = m->eocF.p; A.nEoC = m->nEoC.p; A.blkEdgeStart = m->blkEdgeStart.p;
    A.gdc = fm.gdc.p; A.woe = m->woe.p; A.fE = m->fE.p; A.dv = m->dv.p; A.invArea = fm.invArea.p; A.H = m->H.p;
    A.dt = dt; A.f0 = m->f0;
    A.u = t->u[p].p; A.h = t->h[p].p; A.ssh = t->ssh[p].p; A.hEold = t->hE[p].p;
    A.uNew = t->u[q].p; A.hNew = t->h[q].p; A.sshNew = t->ssh[q].p; A.hEnew = t->hE[q].p;
#define MOKAB_FE_STAGE(S2T, ST, UNIF)                                                                   \
    do {                                                                                                \
        if (A.blockList) fused::k_fe_step<S2T, ST, UNIF, true><<<grid, fused::kThreads, 0, s>>>(A);     \
        else             fused::k_fe_step<S2T, ST, UNIF, false><<<grid, fused::kThreads, 0, s>>>(A);    \
    } while (0)
    if (m->S == 6) { if (m->uniformF) MOKAB_FE_STAGE(10, 6, true); else MOKAB_FE_STAGE(10, 6, false); }
    else           { if (m->uniformF) MOKAB_FE_STAGE(12, 7, true); else MOKAB_FE_STAGE(12, 7, false); }
#undef MOKAB_FE_STAGE
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

// ocn_timestep(::RungeKutta4) with the reference's per-stage kernel sequence (time_integration.jl:112-137)
static void step_rk4_unfused(mokab_state *st, double dt)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh; StateT<double> *t = st->d;
    const double a[3] = {dt / 2.0, dt / 2.0, dt};
    const double b[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};
    advance_time_levels(st);
    const int c = st->cur, o = 1 - c;
    const double *uCur = t->u[o].p, *hCur = t->h[o].p;
    double *uPro = t->uP[0].p, *hPro = t->hP[0].p, *uNew = t->uP[1].p, *hNew = t->hP[1].p;
    cudaStream_t s = ctx->stream;
    const int64_t nEk = st->K * m->nE, nCk = st->K * m->nC;
    MOKAB_CUDA(cudaMemcpyAsync(uPro, uCur, nEk * 8, cudaMemcpyDeviceToDevice, s));
    MOKAB_CUDA(cudaMemcpyAsync(hPro, hCur, nCk * 8, cudaMemcpyDeviceToDevice, s));
    MOKAB_CUDA(cudaMemcpyAsync(uNew, uCur, nEk * 8, cudaMemcpyDeviceToDevice, s));
    MOKAB_CUDA(cudaMemcpyAsync(hNew, hCur, nCk * 8, cudaMemcpyDeviceToDevice, s));
    for (int sg = 0; sg < 4; ++sg) {
        update_ssh(st, hPro, t->sshProv.p);
        for (int64_t k = 0; k < st->K; ++k)
            LAUNCH(ctx, ref::k_interpolate_cell2edge, nblk(m->nE), 256, (int)m->nE, m->ce.p, (const double *)(hPro + k * m->nC), t->hEdge.p + k * m->nE);
        LAUNCH(ctx, ref::k_mul, nblk(nEk), 256, nEk, (const double *)uPro, (const double *)t->hEdge.p, t->flux.p);
        tend_u(st, t->sshProv.p, uPro);
        tend_h(st, t->flux.p);
        if (sg < 3) {
            LAUNCH(ctx, ref::k_axpy, nblk(nEk), 256, nEk, uCur, a[sg], (const double *)t->tendU.p, uPro);
            LAUNCH(ctx, ref::k_axpy, nblk(nCk), 256, nCk, hCur, a[sg], (const double *)t->tendH.p, hPro);
        }
        LAUNCH(ctx, ref::k_axpy, nblk(nEk), 256, nEk, (const double *)uNew, b[sg], (const double *)t->tendU.p, uNew);
        LAUNCH(ctx, ref::k_axpy, nblk(nCk), 256, nCk, (const double *)hNew, b[sg], (const double *)t->tendH.p, hNew);
    }
    MOKAB_CUDA(cudaMemcpyAsync(t->u[c].p, uNew, nEk * 8, cudaMemcpyDeviceToDevice, s));
    MOKAB_CUDA(cudaMemcpyAsync(t->h[c].p, hNew, nCk * 8, cudaMemcpyDeviceToDevice, s));
    update_ssh(st, hNew, t->ssh[c].p);
}

static int p2p_target(const mokab_state *st, int stage);
// MOKAB_STAGE_TMA=1|2 selects a TMA variant of the fused stage kernel on hexagon meshes (not yet measured on hardware)
// (1: the slot-major rows, ten bulk copies per block; 2: a block-major copy of the weights, one bulk copy per block)
// ---- tuning options (mokab_set_option; process-wide; the environment gives the initial values) ---------------------------------
// Captured graphs bake the options into their kernel arguments: every change bumps `epoch` and the graphs are rebuilt on next use.
struct Options {
    int stage_tma = 3;            // MOKAB_STAGE_TMA: 1 / 2 = the bulk-copy variants of the stage kernel, 3 = weights through per-thread cp.async (kernels_fused.cuh)
    int stage_prefetch = 1;       // MOKAB_STAGE_PREFETCH: bit 0 = a block prefetches the streams of its own later edge iterations and of
                                  // its cell phase into L2 at entry; bit 1 = it prefetches the streams of the block launched
                                  // `stage_prefetch_distance` blocks after it
    int stage_prefetch_distance = 0;   // MOKAB_STAGE_PREFETCH_DISTANCE: 0 = one wave of resident blocks (SMs x blocks per SM)
    int stage_wf_block_major = 0; // MOKAB_STAGE_WF_BLOCK_MAJOR: the plain stage kernel reads the Coriolis weights from the block-major copy
    int stage_auto = 1;           // MOKAB_STAGE_AUTO: with stage_tma = 3, a Float64 launch of a few rounds of blocks takes the plain kernel (one
                                  // resident block per SM fewer, no prefetch): prefer_plain_variant
    int stage_auto_hi = 80;       // MOKAB_STAGE_AUTO_HI: ... up to this many blocks per SM (80 x 148 = 11 840 blocks)
    int launch_priority = 1;      // MOKAB_LAUNCH_PRIORITY: launches into a prioritised stream carry that priority as a launch attribute (launch_ex)
    int stage_pdl = 0;            // MOKAB_STAGE_PDL: stage launches carry the programmatic-stream-serialization attribute (kernels_fused.cuh: pdl_*)
    int decomp_serial_blocks = 0;  // MOKAB_DECOMP_SERIAL_BLOCKS: with MOKAB_HALO_P2P_FUSED, a rank whose part has fewer blocks than this runs ONE
                                  // launch per stage (all blocks, exchange folded in) instead of the two-stream overlap schedule.
                                  // Off by default: measured SLOWER at N = 2 (profiles/README.md r02g) -- every block of stage s + 1
                                  // then waits for the neighbours' whole stage s, where the two-stream schedule lets the interior run
    int test_drop_dependency = 0; // TEST HOOK (tests/sim: does the checker have teeth?): 1 / 2 = leave out one of the two cross-stream
                                  // waits of the decomposed RK4 schedule (interior after boundary / boundary after interior);
                                  // 3 = leave out the halo copies of kbar in the decomposed reverse sweep
    int64_t epoch = 0;
    Options()
    {
        auto geti = [](const char *n, int d) { const char *e = getenv(n); return e && *e ? atoi(e) : d; };
        // defaults = the fastest bit-identical variant measured on B200 (profiles/README.md r02d): weights through per-thread
        // cp.async (5 resident blocks instead of 4 in Float64, 6 instead of 5 in Float32) + L2 prefetch of a block's own later
        // iterations; (10, 6) meshes only, everything else runs the plain kernel whatever the switches say
        stage_tma = geti("MOKAB_STAGE_TMA", 3);
        if (stage_tma < 0 || stage_tma > 3) stage_tma = 0;
        stage_prefetch = geti("MOKAB_STAGE_PREFETCH", 1) & 3;
        stage_prefetch_distance = std::max(0, geti("MOKAB_STAGE_PREFETCH_DISTANCE", 0));
        stage_wf_block_major = geti("MOKAB_STAGE_WF_BLOCK_MAJOR", 0) ? 1 : 0;
        stage_auto = geti("MOKAB_STAGE_AUTO", 1) ? 1 : 0;
        stage_auto_hi = std::max(0, geti("MOKAB_STAGE_AUTO_HI", 80));
        stage_pdl = geti("MOKAB_STAGE_PDL", 0) ? 1 : 0;
        launch_priority = geti("MOKAB_LAUNCH_PRIORITY", 1) ? 1 : 0;
        decomp_serial_blocks = std::max(0, geti("MOKAB_DECOMP_SERIAL_BLOCKS", 0));
    }
};
static Options &options() { static Options o; return o; }
static int stage_tma_mode() { return options().stage_tma; }
static bool stage_tma_enabled() { return stage_tma_mode() != 0; }
#ifdef MOKAB_SIM
static void p2p_gate_sim(mokab_state *st, cudaStream_t s);
#endif

// The TMA stage variant needs more dynamic shared memory than the default limit: opt every instantiation in once, OUTSIDE
// any stream capture (the graphs of run_rk4_fused capture the launches themselves).
template <class R>
static void stage_tma_prepare()
{
    static bool done = false;     // (per precision: one static per instantiation)
    if (done || !stage_tma_enabled()) return;
#define MOKAB_TMA_ATTR(STAGE, FOLD, DER)                                                                                                          \
    MOKAB_CUDA(cudaFuncSetAttribute(fused::k_rk_stage<R, STAGE, 10, 6, FOLD, DER, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)); \
    MOKAB_CUDA(cudaFuncSetAttribute(fused::k_rk_stage<R, STAGE, 10, 6, FOLD, DER, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024))
    MOKAB_TMA_ATTR(1, false, false); MOKAB_TMA_ATTR(1, false, true); MOKAB_TMA_ATTR(1, true, false); MOKAB_TMA_ATTR(1, true, true);
    MOKAB_TMA_ATTR(2, false, false); MOKAB_TMA_ATTR(2, false, true); MOKAB_TMA_ATTR(2, true, false); MOKAB_TMA_ATTR(2, true, true);
    MOKAB_TMA_ATTR(4, false, false); MOKAB_TMA_ATTR(4, false, true); MOKAB_TMA_ATTR(4, true, false); MOKAB_TMA_ATTR(4, true, true);
#undef MOKAB_TMA_ATTR
    done = true;
}

// TMA = 2: the block-major copy of the weights (built once per mesh and precision, synchronised like ensure_fused)
template <class R>
static void ensure_wf_block_major(mokab_mesh *m)
{
    FusedMesh<R> &f = fused_of<R>(m);
    if ((stage_tma_mode() != 2 && !options().stage_wf_block_major) || f.wfB.n || !(m->S2 == 10 && m->S == 6)) return;
    mokab_ctx *ctx = m->ctx;
    constexpr int AL = 16 / (int)sizeof(R);
    std::vector<long long> off(m->fusedBlocks + 1, 0);
    for (int b = 0; b < m->fusedBlocks; ++b) {
        const long long nbp = (m->hBlkEdgeStart[b + 1] - m->hBlkEdgeStart[b] + AL - 1) / AL * AL;
        off[b + 1] = off[b] + (long long)m->S2 * nbp;
    }
    f.wfB.alloc((size_t)std::max<long long>(off[m->fusedBlocks], 1) + 16);
    f.wfBOff.upload(off, ctx->stream);
    LAUNCH(ctx, fused::k_build_wf_block_major<R>, m->fusedBlocks, 256, (int)m->nE, m->S2, (const int32_t *)m->blkEdgeStart.p,
           (const long long *)f.wfBOff.p, (const R *)f.wf.p, f.wfB.p);
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// TMA = 3: the slot-interleaved copy of the weights (built once per mesh and precision, synchronised like ensure_fused)
template <class R>
static void ensure_wf_interleaved(mokab_mesh *m)
{
    FusedMesh<R> &f = fused_of<R>(m);
    if (stage_tma_mode() != 3 || f.wfI.n || !((m->S2 == 10 && m->S == 6) || (m->S2 == 12 && m->S == 7))) return;
    mokab_ctx *ctx = m->ctx;
    constexpr int V = 16 / (int)sizeof(R);
    const int NG = (m->S2 + V - 1) / V;
    f.wfI.alloc((size_t)NG * m->nE * V);
    LAUNCH(ctx, fused::k_build_wf_interleaved<R>, nblk(m->nE), 256, (int)m->nE, m->S2, (const R *)f.wf.p, f.wfI.p);
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Launches that carry attributes.  (1) "stage_pdl": the programmatic-stream-serialization attribute -- the launch may become
// resident while the previous kernel of the stream drains; the kernel itself waits (griddepcontrol.wait) before it touches the
// state.  (2) The PRIORITY of the stream the launch goes to, as an attribute of the launch itself: a kernel node captured from a
// high-priority stream does not keep that priority by itself (r02l timeline: the wait kernel and the boundary blocks of the halo
// stream were dispatched only after the LAST block of the concurrently running interior launch had been dispatched, i.e. the
// boundary part ran after the interior part instead of next to it -- 13-17 us per stage exposed on 8 192-block parts); with the
// attribute the block scheduler hands freed SM slots to the halo stream's kernels first, in graphs too.
static int stream_priority_of(cudaStream_t s)
{
#ifdef MOKAB_SIM
    (void)s;
    return 0;
#else
    int p = 0;
    if (s && options().launch_priority) cudaStreamGetPriority(s, &p);
    return p;
#endif
}
template <class... P, class... A>
static void launch_ex(void (*kernel)(P...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, int priority, A &&...args)
{
#ifdef MOKAB_SIM
    (void)kernel; (void)grid; (void)block; (void)smem; (void)s; (void)pdl; (void)priority;
    throw Error("launch_ex: not available on the simulated runtime");
#else
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    if (priority != 0) { at[n].id = cudaLaunchAttributePriority; at[n].val.priority = priority; ++n; }
    cfg.attrs = at; cfg.numAttrs = (unsigned)n;
    MOKAB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(std::forward<A>(args))...));
#endif
}
// the small kernels of the halo exchange: a plain launch unless their stream has a priority to carry
#ifdef MOKAB_SIM
#define MOKAB_LAUNCH_ON(kernel, grid, block, s, ...) kernel<<<(grid), (block), 0, (s)>>>(__VA_ARGS__)
#else
#define MOKAB_LAUNCH_ON(kernel, grid, block, s, ...)                                                     \
    do {                                                                                                 \
        const int prio_ = stream_priority_of(s);                                                         \
        if (prio_ != 0) launch_ex(kernel, (grid), (block), 0, (s), false, prio_, __VA_ARGS__);           \
        else kernel<<<(grid), (block), 0, (s)>>>(__VA_ARGS__);                                           \
    } while (0)
#endif
static bool stage_pdl_enabled()
{
#if defined(MOKAB_SIM) || defined(MOKAB_STATE_LOADS_LDG)
    return false;     // (a build that reads the state through the read-only path must not start early: kernels_fused.cuh ld_state)
#else
    return options().stage_pdl != 0;
#endif
}

// Which of the two tuned stage kernels a launch of `grid` blocks takes (option "stage_auto", Float64).  Blocks of one launch
// start together and last about equally long, so a launch runs in ROUNDS of (SMs x resident blocks) and pays for a whole last
// round; a round lasts longer the more blocks are resident.  The cp.async-weights kernel (5 resident blocks) moves 1-2 % more
// bytes per second than the plain one (4) once a launch has tens of rounds, and its blocks are the quicker ones when a launch
// is a single partial round; in between the plain kernel wins by a wide margin.  Measured (cp.async vs plain, G cell-steps/s):
//   452 + 60 blocks per GPU (512 x 512 over 2 GPUs, r02j)   2.36 vs 2.19
//   1 024 blocks (512 x 512, r02i)                           2.36 vs 2.96
//   4 096 blocks (1024 x 1024 channel, r02i)                 2.48 vs 2.66
//   16 384 blocks (2048 x 2048, r02i; sustained)             2.85 vs 2.87        65 536 blocks (4096 x 4096)   2.92 vs 2.91
// Rule: the plain kernel for launches of more than one of its rounds and at most `stage_auto_hi` blocks per SM.
static bool prefer_plain_variant(int grid, int num_sms, int r_plain)
{
    return grid > num_sms * r_plain && grid <= num_sms * options().stage_auto_hi;
}

// ---- fused RK4 ------------------------------------------------------------------------------------------------
template <class R, int STAGE>
static void launch_stage(mokab_ctx *ctx, const mokab_mesh *m, fused::StageArgs<R> A, int part = MOKAB_PART_ALL,
                         cudaStream_t stream = nullptr)
{
    int grid = m->fusedBlocks;
    A.blockList = nullptr;
    if (part == MOKAB_PART_INTERIOR) { grid = m->nInterior; A.blockList = m->blkInterior.p; }
    if (part == MOKAB_PART_BOUNDARY || part == MOKAB_PART_BOUNDARY_PUSH) { grid = m->nBoundary; A.blockList = m->blkBoundary.p; }
    if (grid == 0) return;
#ifdef MOKAB_TRACE
    A.traceKind = STAGE | (part << 4);
#endif
    cudaStream_t s = stream ? stream : ctx->stream;
    const bool hex = m->S2 == 10 && m->S == 6;
    const bool hept = m->S2 == 12 && m->S == 7;   // pentagons / hexagons / heptagons (quasi-uniform MPAS meshes): rows padded to 12 / 7
    if (part == MOKAB_PART_BOUNDARY_PUSH || part == MOKAB_PART_ALL_PUSH) {   // explicit edgesOnEdge: a boundary block reads halo rows, which cannot be rebuilt
        const int prioP = stream_priority_of(s);
#define MOKAB_STAGE_PUSH(S2T, ST, FOLD)                                                                                             \
    do {                                                                                                                            \
        if (prioP != 0) launch_ex(fused::k_rk_stage<R, STAGE, S2T, ST, FOLD, false, true>, grid, fused::kThreads, 0, s, false, prioP, A, m->S2, m->S); \
        else fused::k_rk_stage<R, STAGE, S2T, ST, FOLD, false, true><<<grid, fused::kThreads, 0, s>>>(A, m->S2, m->S);               \
    } while (0)
        if (hex && m->uniformF)      MOKAB_STAGE_PUSH(10, 6, false);
        else if (hex)                MOKAB_STAGE_PUSH(10, 6, true);
        else if (m->uniformF)        MOKAB_STAGE_PUSH(0, 0, false);
        else                         MOKAB_STAGE_PUSH(0, 0, true);
#undef MOKAB_STAGE_PUSH
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
        return;
    }
    const bool der = (hex || hept) && m->nDerivedBlocks > 0;
    if (hex && !stage_tma_enabled() && options().stage_wf_block_major && part != MOKAB_PART_BOUNDARY_PUSH) {
        FusedMesh<R> &fm = fused_of<R>(const_cast<mokab_mesh *>(m));
        if (fm.wfB.n) { A.wfB = fm.wfB.p; A.wfBOff = fm.wfBOff.p; }
    }
    bool use_cpa = (hex || hept) && stage_tma_mode() == 3;
    // (Float64 only: in Float32 the cp.async kernel wins or ties on balance -- r02i: 1 024 blocks 4.14 (cp.async) vs 3.77 G
    //  (plain), 4 096 blocks 3.65 vs 3.77, 16 384 blocks 4.27 vs 4.21)
    if (use_cpa && options().stage_auto && sizeof(R) == 8) {
        const int rp = der ? fused::stage_blocks<R, true, 0>() : fused::stage_blocks<R, false, 0>();
        if (prefer_plain_variant(grid, ctx->num_sms, rp)) {
            use_cpa = false;
            A.pf = 0;
        }
    }
    if (use_cpa) {   // the default: weights through per-thread cp.async into shared memory (kernels_fused.cuh)
        FusedMesh<R> &fm = fused_of<R>(const_cast<mokab_mesh *>(m));
        if (fm.wfI.n) {
            A.wfI = fm.wfI.p;
            const size_t smem = (size_t)(hex ? fused::cpa_groups<R, 10>() : fused::cpa_groups<R, 12>()) * fused::kThreads * 16;
            const bool pdl = stage_pdl_enabled();
            const int prio = stream_priority_of(s);
#define MOKAB_STAGE_CPA(S2T, ST, FOLD, DER)                                                                                          \
    do {                                                                                                                            \
        auto k_rk_stage_cpa = fused::k_rk_stage<R, STAGE, S2T, ST, FOLD, DER, false, 3>;                                            \
        if (pdl || prio != 0) launch_ex(k_rk_stage_cpa, grid, fused::kThreads, smem, s, pdl, prio, A, m->S2, m->S);                 \
        else k_rk_stage_cpa<<<grid, fused::kThreads, smem, s>>>(A, m->S2, m->S);                                                    \
    } while (0)
            if (hex) {
                if (der && m->uniformF)      MOKAB_STAGE_CPA(10, 6, false, true);
                else if (der)                MOKAB_STAGE_CPA(10, 6, true, true);
                else if (m->uniformF)        MOKAB_STAGE_CPA(10, 6, false, false);
                else                         MOKAB_STAGE_CPA(10, 6, true, false);
            } else {
                if (der && m->uniformF)      MOKAB_STAGE_CPA(12, 7, false, true);
                else if (der)                MOKAB_STAGE_CPA(12, 7, true, true);
                else if (m->uniformF)        MOKAB_STAGE_CPA(12, 7, false, false);
                else                         MOKAB_STAGE_CPA(12, 7, true, false);
            }
#undef MOKAB_STAGE_CPA
            MOKAB_CUDA(cudaGetLastError());
            ctx->launches++;
            return;
        }
    }
    if (hex && (stage_tma_mode() == 1 || stage_tma_mode() == 2)) {   // opt-in: the weight rows of a block through bulk asynchronous copies (kernels_fused.cuh)
        constexpr int AL = fused::tma_align<R>();
        A.wStride = (m->maxBlockEdges + AL + AL - 1) / AL * AL;
        const size_t smem = (size_t)10 * A.wStride * sizeof(R);
        if (stage_tma_mode() == 2) {
            FusedMesh<R> &fm = fused_of<R>(const_cast<mokab_mesh *>(m));
            A.wfB = fm.wfB.p; A.wfBOff = fm.wfBOff.p;
        }
        if (smem <= 72 * 1024) {        // three blocks per SM; meshes whose blocks own more edges than that keep the plain kernel
#define MOKAB_STAGE_TMA(FOLD, DER)                                                                                                  \
    do {                                                                                                                            \
        if (stage_tma_mode() == 2) {                                                                                                \
            auto k_rk_stage_tma = fused::k_rk_stage<R, STAGE, 10, 6, FOLD, DER, false, 2>;                                          \
            k_rk_stage_tma<<<grid, fused::kThreads, smem, s>>>(A, m->S2, m->S);                                                     \
        } else {                                                                                                                    \
            auto k_rk_stage_tma = fused::k_rk_stage<R, STAGE, 10, 6, FOLD, DER, false, 1>;                                          \
            k_rk_stage_tma<<<grid, fused::kThreads, smem, s>>>(A, m->S2, m->S);                                                     \
        }                                                                                                                           \
    } while (0)
            if (der && m->uniformF)      MOKAB_STAGE_TMA(false, true);
            else if (der)                MOKAB_STAGE_TMA(true, true);
            else if (m->uniformF)        MOKAB_STAGE_TMA(false, false);
            else                         MOKAB_STAGE_TMA(true, false);
#undef MOKAB_STAGE_TMA
            MOKAB_CUDA(cudaGetLastError());
            ctx->launches++;
            return;
        }
        A.wStride = 0;
    }
    const bool pdl0 = stage_pdl_enabled();
    const int prio0 = stream_priority_of(s);
#define MOKAB_STAGE(S2T, ST, FOLD, DER)                                                                                              \
    do {                                                                                                                            \
        auto k_rk_stage_plain = fused::k_rk_stage<R, STAGE, S2T, ST, FOLD, DER>;                                                    \
        if (pdl0 || prio0 != 0) launch_ex(k_rk_stage_plain, grid, fused::kThreads, 0, s, pdl0, prio0, A, m->S2, m->S);              \
        else k_rk_stage_plain<<<grid, fused::kThreads, 0, s>>>(A, m->S2, m->S);                                                     \
    } while (0)
    if (hex && der && m->uniformF)  MOKAB_STAGE(10, 6, false, true);
    else if (hex && der)            MOKAB_STAGE(10, 6, true, true);
    else if (hex && m->uniformF)    MOKAB_STAGE(10, 6, false, false);
    else if (hex)                   MOKAB_STAGE(10, 6, true, false);
    else if (hept && der && m->uniformF) MOKAB_STAGE(12, 7, false, true);
    else if (hept && der)           MOKAB_STAGE(12, 7, true, true);
    else if (hept && m->uniformF)   MOKAB_STAGE(12, 7, false, false);
    else if (hept)                  MOKAB_STAGE(12, 7, true, false);
    else if (m->uniformF)          MOKAB_STAGE(0, 0, false, false);
    else                           MOKAB_STAGE(0, 0, true, false);
#undef MOKAB_STAGE
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

template <class R>
static fused::StageArgs<R> stage_args(mokab_state *st, double dt, int p, int stage)
{
    mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<R> *t = typed<R>(st);
    FusedMesh<R> &fm = fused_of<R>(m);
    fused::StageArgs<R> A;
    A.nE = (int)m->nE; A.nC = (int)m->nC; A.nCown = (int)m->nCo; A.blockList = nullptr;
    A.ce = m->ce.p; A.eoe = m->eoeF.p; A.eoc = m->eocF.p; A.nEoE = m->nEoE.p; A.nEoC = m->nEoC.p;
    A.blkEdgeStart = m->blkEdgeStart.p;
    A.posE = m->posE.p; A.blkDerived = m->nDerivedBlocks ? m->blkDerived.p : nullptr;
    A.gdc = fm.gdc.p; A.wf = fm.wf.p; A.dv = fm.dv.p; A.invArea = fm.invArea.p; A.H = fm.H.p;
    A.uCur = t->u[p].p; A.hCur = t->h[p].p; A.uAcc = t->u[1 - p].p; A.hAcc = t->h[1 - p].p;
    const double a[4] = {dt / 2.0, dt / 2.0, dt, 0.0};                  // time_integration.jl:77
    const double b[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};       // time_integration.jl:78
    A.a = (R)a[stage - 1]; A.b = (R)b[stage - 1];
    A.f0 = (R)m->f0;
    A.push = nullptr;
    A.wStride = 0; A.wfB = nullptr; A.wfBOff = nullptr; A.wfI = nullptr;
    A.pf = options().stage_prefetch;
    A.pfDist = options().stage_prefetch_distance > 0 ? options().stage_prefetch_distance
                                                       : st->ctx->num_sms * (sizeof(R) == 8 ? 4 : 5);
    switch (stage) {
    case 1: A.uOld = t->u[p].p;  A.hOld = t->h[p].p;  A.uOut = t->uP[0].p; A.hOut = t->hP[0].p; break;  // provisional == current
    case 2: A.uOld = t->uP[0].p; A.hOld = t->hP[0].p; A.uOut = t->uP[1].p; A.hOut = t->hP[1].p; break;
    case 3: A.uOld = t->uP[1].p; A.hOld = t->hP[1].p; A.uOut = t->uP[0].p; A.hOut = t->hP[0].p; break;
    default: A.uOld = t->uP[0].p; A.hOld = t->hP[0].p; A.uOut = nullptr; A.hOut = nullptr; break;
    }
    return A;
}

// the (u, h) buffers holding the output of RK stage `stage` (0 = the current state) for step parity p
template <class R>
static void stage_output(mokab_state *st, int stage, R **u, R **h)
{
    StateT<R> *t = typed<R>(st);
    const int p = st->cur;
    switch (stage) {
    case 0: *u = t->u[p].p; *h = t->h[p].p; break;
    case 1: case 3: *u = t->uP[0].p; *h = t->hP[0].p; break;
    case 2: *u = t->uP[1].p; *h = t->hP[1].p; break;
    case 5:             // staged ForwardEuler (Float64): the other two arrays a step writes -- (layerThicknessEdge, ssh)[new]
        if constexpr (sizeof(R) == 8) {
            MOKAB_REQUIRE(t->hE[0].n, "halo_pack/unpack stage 5: no staged ForwardEuler step has run");
            *u = t->hE[1 - p].p; *h = t->ssh[1 - p].p;
        } else {
            MOKAB_REQUIRE(false, "halo_pack/unpack stage 5: ForwardEuler is Float64 only");
        }
        break;
    default: *u = t->u[1 - p].p; *h = t->h[1 - p].p; break;
    }
}

template <class R>
static void run_stage(mokab_state *st, double dt, int stage, int part, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    ensure_fused<R>(const_cast<mokab_mesh *>(m));
    stage_tma_prepare<R>();
    ensure_wf_block_major<R>(const_cast<mokab_mesh *>(m));
    ensure_wf_interleaved<R>(const_cast<mokab_mesh *>(m));
    fused::StageArgs<R> A = stage_args<R>(st, dt, st->cur, stage);
    if (part == MOKAB_PART_BOUNDARY_PUSH || part == MOKAB_PART_ALL_PUSH) {
        MOKAB_REQUIRE(st->p2p.ready, "rk4_stage(MOKAB_PART_*_PUSH): call mokab_p2p_setup first");
        MOKAB_REQUIRE(part == MOKAB_PART_ALL_PUSH || m->nBoundary > 0 || (st->p2p.recvRanks.empty() && st->p2p.sendRanks.empty()),
                      "rk4_stage(MOKAB_PART_BOUNDARY_PUSH): a rank with neighbours has no boundary block");
        A.push = (const fused::PushStage<R> *)st->p2p.stageDesc.p + p2p_target(st, stage);
#ifdef MOKAB_SIM
        p2p_gate_sim(st, stream ? stream : ctx->stream);
#endif
    }
    switch (stage) {
    case 1: launch_stage<R, 1>(ctx, m, A, part, stream); break;
    case 2: case 3: launch_stage<R, 2>(ctx, m, A, part, stream); break;
    default: launch_stage<R, 4>(ctx, m, A, part, stream); break;
    }
}

template <class R>
static void halo_pack(mokab_state *st, int stage, void *buf, cudaStream_t stream, bool unpack)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    R *u, *h;
    stage_output<R>(st, stage, &u, &h);
    cudaStream_t s = stream ? stream : ctx->stream;
    if (!unpack) {
        const int n = (int)m->haloSend.n;
        if (n) MOKAB_LAUNCH_ON(k_halo_pack<R>, nblk(n), 256, s, n, (int)m->nC, (const int32_t *)m->haloSend.p, (const R *)h, (const R *)u, (R *)buf);
    } else {
        const int n = (int)m->haloRecv.n;
        if (n) MOKAB_LAUNCH_ON(k_halo_unpack<R>, nblk(n), 256, s, n, (int)m->nC, (const int32_t *)m->haloRecv.p, (const R *)buf, h, u);
    }
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

// Halo copies of an (edge array, cell array) pair that is not a stage output -- the stage states and the adjoint variables
// of the reverse sweep on a decomposed mesh -- through the packed exchange of mokab_decomp_setup, on the context's stream.
template <class R> static void p2p_push_ll_arrays(mokab_state *st, const R *u, const R *h, cudaStream_t stream);
template <class R> static void p2p_wait_ll_arrays(mokab_state *st, R *u, R *h, cudaStream_t stream);
template <class R>
static void halo_exchange_arrays(mokab_state *st, R *u, R *h)
{
    mokab_state::Decomp &D = st->dec;
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    cudaStream_t s = ctx->stream;
    if (D.mode == MOKAB_HALO_P2P_LL) {   // flag-in-data packets carry values, not addresses: any array pair goes over them
        p2p_push_ll_arrays<R>(st, u, h, s);
        p2p_wait_ll_arrays<R>(st, u, h, s);
        return;
    }
    const int ns = (int)m->haloSend.n, nr = (int)m->haloRecv.n;
    if (ns) {
        k_halo_pack<R><<<nblk(ns), 256, 0, s>>>(ns, (int)m->nC, m->haloSend.p, (const R *)h, (const R *)u, (R *)D.sendBuf.p);
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    comm::all_to_all(D.comm, s, D.sendBuf.p, D.recvBuf.p, D.scnt.data(), D.rcnt.data(), sizeof(R));
    if (nr) {
        k_halo_unpack<R><<<nblk(nr), 256, 0, s>>>(nr, (int)m->nC, m->haloRecv.p, (const R *)D.recvBuf.p, h, u);
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
}

// one RK4 step reading time level `p`, writing level 1-p (four launches)
template <class R>
static void enqueue_rk4_step(mokab_state *st, double dt, int p)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    launch_stage<R, 1>(ctx, m, stage_args<R>(st, dt, p, 1));
    launch_stage<R, 2>(ctx, m, stage_args<R>(st, dt, p, 2));
    launch_stage<R, 2>(ctx, m, stage_args<R>(st, dt, p, 3));
    launch_stage<R, 4>(ctx, m, stage_args<R>(st, dt, p, 4));
}

template <class R>
static void build_graphs(mokab_state *st, double dt)
{
    mokab_ctx *ctx = st->ctx;
    StateT<R> *t = typed<R>(st);
    t->drop_graphs();
    const int64_t saved = ctx->launches;
    for (int p = 0; p < 2; ++p)
        for (int pair = 0; pair < 2; ++pair) {
            cudaGraph_t g = nullptr;
            MOKAB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            try {
                enqueue_rk4_step<R>(st, dt, p);
                if (pair) enqueue_rk4_step<R>(st, dt, 1 - p);
            } catch (...) {
                cudaStreamEndCapture(ctx->stream, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            MOKAB_CUDA(cudaStreamEndCapture(ctx->stream, &g));
            cudaGraphExec_t ge = nullptr;
            cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
            cudaGraphDestroy(g);
            MOKAB_CUDA(e);
            (pair ? t->gPair : t->gStep)[p] = ge;
        }
    ctx->launches = saved;  // capture enqueues are not executions
    t->graph_dt = dt;
    t->graph_epoch = options().epoch;
    t->graphs_ready = true;
}

template <class R>
static void refresh_ssh(mokab_state *st, cudaStream_t stream = nullptr)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<R> *t = typed<R>(st);
    ensure_fused<R>(m);
    FusedMesh<R> &fm = fused_of<R>(m);
    cudaStream_t s = stream ? stream : ctx->stream;
    if constexpr (sizeof(R) == 8) {
        if (st->K > 1) {
            for (int l = 0; l < 2; ++l) update_ssh(st, t->h[l].p, t->ssh[l].p, s);
            return;
        }
    }
    for (int l = 0; l < 2; ++l) {
        k_update_ssh<R><<<nblk(m->nC), 256, 0, s>>>(m->nC, (const R *)t->h[l].p, (const R *)fm.H.p, t->ssh[l].p);
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
}

// Multi-level states on a decomposed mesh: one packed exchange per stage of K + 1 planes (k_halo_pack_ml), on the context's
// stream (the multi-level path has no interior / boundary split; the buffers are sized by decomp_prepare, outside any capture).
static void ensure_level_buffers(mokab_state *st)          // (allocates: never inside a stream capture)
{
    mokab_state::Decomp &D = st->dec;
    const mokab_mesh *m = st->mesh;
    const size_t ns = std::max<size_t>(m->haloSend.n, 1) * (size_t)(st->K + 1) * 8, nr = std::max<size_t>(m->haloRecv.n, 1) * (size_t)(st->K + 1) * 8;
    if (D.sendBufML.n < ns) { D.sendBufML.alloc(ns); D.sendBufML.zero(st->ctx->stream); }
    if (D.recvBufML.n < nr) { D.recvBufML.alloc(nr); D.recvBufML.zero(st->ctx->stream); }
}

static void halo_exchange_levels(mokab_state *st, double *u, double *h, double *ssh)
{
    mokab_state::Decomp &D = st->dec;
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    cudaStream_t s = ctx->stream;
    const int K = st->K, nseg = D.comm->nranks;
    const int ns = (int)m->haloSend.n, nr = (int)m->haloRecv.n;
    MOKAB_REQUIRE(D.sendBufML.n >= (size_t)std::max(ns, 1) * (K + 1) * 8 && D.recvBufML.n >= (size_t)std::max(nr, 1) * (K + 1) * 8,
                  "decomposed multi-level step: the level buffers are not allocated (internal)");
    if (ns) {
        k_halo_pack_ml<double><<<nblk(ns) * (K + 1), 256, 0, s>>>(ns, (int)m->nC, (int)m->nE, K, m->haloSend.p, D.sendOff.p, nseg,
                                                                                            (const double *)h, (const double *)u, (const double *)ssh, (double *)D.sendBufML.p);
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    comm::all_to_all(D.comm, s, D.sendBufML.p, D.recvBufML.p, D.scnt.data(), D.rcnt.data(), sizeof(double) * (size_t)(K + 1));
    if (nr) {
        k_halo_unpack_ml<double><<<nblk(nr) * (K + 1), 256, 0, s>>>(nr, (int)m->nC, (int)m->nE, K, m->haloRecv.p, D.recvOff.p, nseg,
                                                                                              (const double *)D.recvBufML.p, h, u, ssh);
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
}

// ---- fused RK4 on multi-level states (K > 1; fused::k_rk_stage_ml) ---------------------------------------------------------
template <int STAGE>
static void launch_stage_ml(mokab_ctx *ctx, const mokab_mesh *m, const fused::StageArgsML &A)
{
    const int grid = m->fusedBlocks;
    if (grid == 0) return;
    const bool hex = m->S2 == 10 && m->S == 6, hept = m->S2 == 12 && m->S == 7;
#define MOKAB_STAGE_ML(S2T, ST, FOLD) fused::k_rk_stage_ml<STAGE, S2T, ST, FOLD><<<grid, fused::kThreads, 0, ctx->stream>>>(A, m->S2, m->S)
    if (hex && m->uniformF)       MOKAB_STAGE_ML(10, 6, false);
    else if (hex)                 MOKAB_STAGE_ML(10, 6, true);
    else if (hept && m->uniformF) MOKAB_STAGE_ML(12, 7, false);
    else if (hept)                MOKAB_STAGE_ML(12, 7, true);
    else if (m->uniformF)         MOKAB_STAGE_ML(0, 0, false);
    else                          MOKAB_STAGE_ML(0, 0, true);
#undef MOKAB_STAGE_ML
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

// one RK4 step of a multi-level state reading time level p, writing level 1 - p (ssh[1 - p] included)
static void enqueue_rk4_step_ml(mokab_state *st, double dt, int p)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<double> *t = st->d;
    FusedMesh<double> &fm = fused_of<double>(m);
    fused::StageArgsML A;
    A.nE = (int)m->nE; A.nC = (int)m->nC; A.K = st->K; A.nCown = (int)m->nCo;
    A.ce = m->ce.p; A.eoe = m->eoeF.p; A.eoc = m->eocF.p; A.nEoE = m->nEoE.p; A.nEoC = m->nEoC.p; A.blkEdgeStart = m->blkEdgeStart.p;
    A.gdc = fm.gdc.p; A.wf = fm.wf.p; A.dv = fm.dv.p; A.invArea = fm.invArea.p; A.H = fm.H.p;
    A.uCur = t->u[p].p; A.hCur = t->h[p].p; A.uAcc = t->u[1 - p].p; A.hAcc = t->h[1 - p].p;
    A.f0 = m->f0;
    const double a[4] = {dt / 2.0, dt / 2.0, dt, 0.0};                  // time_integration.jl:77
    const double b[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};       // time_integration.jl:78
    for (int stage = 1; stage <= 4; ++stage) {
        A.a = a[stage - 1]; A.b = b[stage - 1];
        switch (stage) {
        case 1: A.uOld = t->u[p].p;  A.hOld = t->h[p].p;  A.sshOld = t->ssh[p].p;  A.uOut = t->uP[0].p; A.hOut = t->hP[0].p; A.sshOut = t->sshP[0].p; break;
        case 2: A.uOld = t->uP[0].p; A.hOld = t->hP[0].p; A.sshOld = t->sshP[0].p; A.uOut = t->uP[1].p; A.hOut = t->hP[1].p; A.sshOut = t->sshP[1].p; break;
        case 3: A.uOld = t->uP[1].p; A.hOld = t->hP[1].p; A.sshOld = t->sshP[1].p; A.uOut = t->uP[0].p; A.hOut = t->hP[0].p; A.sshOut = t->sshP[0].p; break;
        default: A.uOld = t->uP[0].p; A.hOld = t->hP[0].p; A.sshOld = t->sshP[0].p; A.uOut = nullptr; A.hOut = nullptr; A.sshOut = t->ssh[1 - p].p; break;
        }
        if (stage == 1) launch_stage_ml<1>(ctx, m, A);
        else if (stage == 4) launch_stage_ml<4>(ctx, m, A);
        else launch_stage_ml<2>(ctx, m, A);
        // decomposed mesh: the halo copies of what this stage wrote -- K levels of (u, h) and the free surface -- in one message
        if (st->dec.ready) halo_exchange_levels(st, stage == 4 ? A.uAcc : A.uOut, stage == 4 ? A.hAcc : A.hOut, A.sshOut);
    }
}

static void run_rk4_fused_ml(mokab_state *st, double dt, int64_t nsteps)
{
    mokab_ctx *ctx = st->ctx;
    StateT<double> *t = st->d;
    MOKAB_REQUIRE(st->dtype == MOKAB_F64, "timestep_rk4: multi-level states are Float64");
    if (t->taping) {
        MOKAB_REQUIRE((int64_t)t->tapeDt.size() + nsteps <= t->tapeCap, "timestep_rk4: the tape is full (mokab_tape_begin max_steps)");
        MOKAB_REQUIRE(t->tapeKind != 2, "timestep_rk4: the tape already holds ForwardEuler steps");
        t->tapeKind = 1;
    }
    if (nsteps <= 0) return;
    if (t->taping) {   // record the state before every step (plain launches: the tape slot changes per step)
        const mokab_mesh *m = st->mesh;
        const size_t K = (size_t)st->K;
        ensure_fused<double>(const_cast<mokab_mesh *>(m));
        if (t->tapeH.n < (size_t)t->tapeCap * K * m->nC) t->tapeH.alloc((size_t)t->tapeCap * K * m->nC);
        update_ssh(st, t->h[st->cur].p, t->ssh[st->cur].p);
        for (int64_t i = 0; i < nsteps; ++i) {
            const size_t k = t->tapeDt.size();
            MOKAB_CUDA(cudaMemcpyAsync(t->tapeU.p + k * K * m->nE, t->u[st->cur].p, K * m->nE * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            MOKAB_CUDA(cudaMemcpyAsync(t->tapeH.p + k * K * m->nC, t->h[st->cur].p, K * m->nC * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            t->tapeDt.push_back(dt);
            enqueue_rk4_step_ml(st, dt, st->cur);
            st->cur = 1 - st->cur;
        }
        update_ssh(st, t->h[1 - st->cur].p, t->ssh[1 - st->cur].p);
        return;
    }
    // the stage kernels gather ssh of the state they read: make ssh[cur] what the current layerThickness implies (every later
    // step's ssh is written by the stage that produces its layerThickness)
    update_ssh(st, t->h[st->cur].p, t->ssh[st->cur].p);
    if (!t->graphs_ready || t->graph_dt != dt || t->graph_epoch != options().epoch) {
        t->drop_graphs();
        const int64_t saved = ctx->launches;
        for (int p = 0; p < 2; ++p)
            for (int pair = 0; pair < 2; ++pair) {
                cudaGraph_t g = nullptr;
                MOKAB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                try {
                    enqueue_rk4_step_ml(st, dt, p);
                    if (pair) enqueue_rk4_step_ml(st, dt, 1 - p);
                } catch (...) {
                    cudaStreamEndCapture(ctx->stream, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                MOKAB_CUDA(cudaStreamEndCapture(ctx->stream, &g));
                cudaGraphExec_t ge = nullptr;
                cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
                cudaGraphDestroy(g);
                MOKAB_CUDA(e);
                (pair ? t->gPair : t->gStep)[p] = ge;
            }
        ctx->launches = saved;
        t->graph_dt = dt; t->graph_epoch = options().epoch; t->graphs_ready = true;
    }
    int64_t left = nsteps;
    while (left >= 2) { MOKAB_CUDA(cudaGraphLaunch(t->gPair[st->cur], ctx->stream)); ctx->launches += 8; left -= 2; }
    if (left) { MOKAB_CUDA(cudaGraphLaunch(t->gStep[st->cur], ctx->stream)); ctx->launches += 4; st->cur = 1 - st->cur; }
    // Prog.ssh[1] of the reference after its last step: the free surface of the previous state
    update_ssh(st, t->h[1 - st->cur].p, t->ssh[1 - st->cur].p);
}

template <class R>
static void run_rk4_fused(mokab_state *st, double dt, int64_t nsteps)
{
    mokab_ctx *ctx = st->ctx;
    StateT<R> *t = typed<R>(st);
    if (st->K > 1) { run_rk4_fused_ml(st, dt, nsteps); return; }
    MOKAB_REQUIRE(st->mesh->nCo == st->mesh->nC && st->mesh->nEo == st->mesh->nE,
                  "timestep_rk4: this mesh has halo entities; drive it with mokab_rk4_stage + mokab_halo_pack/unpack");
    ensure_fused<R>(const_cast<mokab_mesh *>(st->mesh));
    stage_tma_prepare<R>();
    ensure_wf_block_major<R>(const_cast<mokab_mesh *>(st->mesh));
    ensure_wf_interleaved<R>(const_cast<mokab_mesh *>(st->mesh));
    if (t->taping) {  // record the state before every step (plain launches: the tape slot changes per step)
        const mokab_mesh *m = st->mesh;
        MOKAB_REQUIRE((int64_t)t->tapeDt.size() + nsteps <= t->tapeCap, "timestep_rk4: the tape is full (mokab_tape_begin max_steps)");
        MOKAB_REQUIRE(t->tapeKind != 2, "timestep_rk4: the tape already holds ForwardEuler steps");
        t->tapeKind = 1;                    // (also a call with nsteps = 0: the seed then follows this stepper's state definition)
        if (nsteps > 0 && t->tapeH.n < (size_t)t->tapeCap * m->nC) t->tapeH.alloc((size_t)t->tapeCap * m->nC);
        for (int64_t i = 0; i < nsteps; ++i) {
            const size_t k = t->tapeDt.size();
            MOKAB_CUDA(cudaMemcpyAsync(t->tapeU.p + k * m->nE, t->u[st->cur].p, m->nE * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
            MOKAB_CUDA(cudaMemcpyAsync(t->tapeH.p + k * m->nC, t->h[st->cur].p, m->nC * sizeof(R), cudaMemcpyDeviceToDevice, ctx->stream));
            t->tapeDt.push_back(dt);
            enqueue_rk4_step<R>(st, dt, st->cur);
            st->cur = 1 - st->cur;
        }
        if (nsteps > 0) refresh_ssh<R>(st);
        return;
    }
    if (!t->graphs_ready || t->graph_dt != dt || t->graph_epoch != options().epoch) build_graphs<R>(st, dt);
    int64_t left = nsteps;
    while (left >= 2) {
        MOKAB_CUDA(cudaGraphLaunch(t->gPair[st->cur], ctx->stream));
        ctx->launches += 8;
        left -= 2;
    }
    if (left) {
        MOKAB_CUDA(cudaGraphLaunch(t->gStep[st->cur], ctx->stream));
        ctx->launches += 4;
        st->cur = 1 - st->cur;
    }
    if (nsteps > 0) refresh_ssh<R>(st);
}

template <class R>
static void do_reduce(mokab_state *st, int which, double *out)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<R> *t = typed<R>(st);
    ensure_fused<R>(m);
    FusedMesh<R> &fm = fused_of<R>(m);
    const int c = st->cur;
    MOKAB_REQUIRE(which >= 0 && which <= 2, "reduce: unknown reduction id");
    LAUNCH(ctx, reduce::k_cells<R>, reduce::kBlocks, reduce::kThreads, which, m->nCo, st->K, m->nC, (const R *)t->h[c].p, (const R *)fm.H.p,
           (const double *)m->area.p, t->partial.p);
    if (which == MOKAB_SUM_ENERGY)
        LAUNCH(ctx, reduce::k_edges_ke<R>, reduce::kBlocks, reduce::kThreads, m->nEo, st->K, m->nE, m->nC, (const int2 *)m->ce.p,
               (const double *)m->dc.p, (const double *)m->dv.p, (const R *)t->u[c].p, (const R *)t->h[c].p, (const R *)fm.H.p, t->partial.p);
    LAUNCH(ctx, reduce::k_final, 1, reduce::kThreads, (const double *)t->partial.p, t->result.p);
    MOKAB_CUDA(cudaMemcpyAsync(out, t->result.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- reverse mode -------------------------------------------------------------------------------------------
// Transpose of the Coriolis stencil: row x lists the edges e whose sum reads u[x], with w[i,e] * f[x]
// (horizontal_advection_and_coriolis.jl:70-72).  Built on the host from the device arrays, once per mesh; also
// checks the two structural facts the gather-form adjoint relies on.
static void ensure_adjoint_mesh(mokab_mesh *m)
{
    if (m->adj_ready) return;
    mokab_ctx *ctx = m->ctx;
    const int64_t nE = m->nE, nC = m->nC, nEo = m->nEo, nCo = m->nCo;
    const int S2 = m->S2, S = m->S;
    std::vector<int32_t> eoe((size_t)S2 * nE), eoc((size_t)S * nC), sgn((size_t)S * nC);
    std::vector<double> woe((size_t)S2 * nE), fE(nE);
    std::vector<uint8_t> nEoE(nE), nEoC(nC);
    std::vector<int2> ce(nE);
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    MOKAB_CUDA(cudaMemcpy(eoe.data(), m->eoe.p, eoe.size() * 4, cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(woe.data(), m->woe.p, woe.size() * 8, cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(fE.data(), m->fE.p, fE.size() * 8, cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(nEoE.data(), m->nEoE.p, nEoE.size(), cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(eoc.data(), m->eoc.p, eoc.size() * 4, cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(sgn.data(), m->sgnC.p, sgn.size() * 4, cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(nEoC.data(), m->nEoC.p, nEoC.size(), cudaMemcpyDeviceToHost));
    MOKAB_CUDA(cudaMemcpy(ce.data(), m->ce.p, ce.size() * sizeof(int2), cudaMemcpyDeviceToHost));
    // (1) edgeSignOnCell follows HorzMesh.jl:292-311 (-1 on the cellsOnEdge[1] side); (2) every edge is listed by
    // both of its cells (once by its only cell when masked).  On a decomposed mesh the rows of the owned cells are
    // checked (halo cells carry none), and (2) for the edges between two owned cells.
    std::vector<uint8_t> seen(nE, 0);
    for (int64_t c = 0; c < nCo; ++c)
        for (int i = 0; i < nEoC[c]; ++i) {
            const int32_t e = eoc[(size_t)i * nC + c];
            const bool first = ce[e].x == c;
            MOKAB_REQUIRE(first || ce[e].y == c, "adjoint: edgesOnCell lists an edge that does not border the cell");
            MOKAB_REQUIRE((sgn[(size_t)i * nC + c] > 0) == !first,
                          "adjoint: edgeSignOnCell does not follow the reference's orientation rule (HorzMesh.jl:292-311)");
            seen[e]++;
        }
    for (int64_t e = 0; e < nEo; ++e)
        if (ce[e].x < nCo && ce[e].y < nCo)
            MOKAB_REQUIRE(seen[e] == (ce[e].x == ce[e].y ? 1 : 2), "adjoint: an edge is not listed by both of its cells");
    // transpose: row x (an OWNED edge) lists every local edge e -- owned, or a halo copy, whose rows the mesh keeps on the
    // host for exactly this -- whose Coriolis sum reads u[x]
    const int HS2 = m->haloS2;
    auto for_each_entry = [&](auto &&fn) {
        for (int64_t e = 0; e < nEo; ++e)
            for (int i = 0; i < nEoE[e]; ++i) {
                const int32_t x = eoe[(size_t)i * nE + e];
                if (x >= 0 && x < nEo) fn(e, x, woe[(size_t)i * nE + e]);
            }
        for (int64_t e = nEo; e < nE; ++e)
            for (int i = 0; i < HS2; ++i) {
                const int32_t x = m->hHaloEoe[(size_t)(e - nEo) * HS2 + i];
                if (x >= 0 && x < nEo) fn(e, x, m->hHaloWoe[(size_t)(e - nEo) * HS2 + i]);
            }
    };
    std::vector<int32_t> cnt(nE, 0);
    for_each_entry([&](int64_t, int32_t x, double) { cnt[x]++; });
    int S2T = 1;
    for (int64_t e = 0; e < nEo; ++e) S2T = std::max(S2T, (int)cnt[e]);
    MOKAB_REQUIRE(S2T <= 255, "adjoint: transposed Coriolis stencil too wide");
    std::vector<int32_t> eoeT((size_t)S2T * nE, 0);
    std::vector<double> wT((size_t)S2T * nE, 0.0);
    std::vector<uint8_t> nT(nE, 0);
    for_each_entry([&](int64_t e, int32_t x, double w) {
        const int j = nT[x]++;
        eoeT[(size_t)j * nE + x] = (int32_t)e;
        wT[(size_t)j * nE + x] = w * fE[x];
    });
    m->S2T = S2T;
    m->eoeT.upload(eoeT, ctx->stream);
    m->woeT.upload(wT, ctx->stream);
    m->nEoET.upload(nT, ctx->stream);
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    m->adj_ready = true;
}

template <class R>
static void ensure_adjoint(mokab_state *st)
{
    mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    mokab_ctx *ctx = st->ctx;
    ensure_fused<R>(m);
    stage_tma_prepare<R>();
    ensure_wf_block_major<R>(m);
    ensure_wf_interleaved<R>(m);
    ensure_adjoint_mesh(m);
    ensure_adj_state<R>(st);
    FusedMesh<R> &f = fused_of<R>(m);
    if (f.wfT.n == 0) {
        f.wfT.alloc(m->woeT.n);
        LAUNCH(ctx, (k_convert<double, R>), nblk(m->woeT.n), 256, (int64_t)m->woeT.n, (const double *)m->woeT.p, f.wfT.p);
    }
}

// Reverse one recorded step: recompute y_2..y_4 from the taped state (three forward stage launches), then the
// four adjoint stages 4, 3, 2, 1.
template <class R>
static void adjoint_step(mokab_state *st, int64_t k)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<R> *t = typed<R>(st);
    FusedMesh<R> &fm = fused_of<R>(m);
    const double dt = t->tapeDt[k];
    const double a[4] = {dt / 2.0, dt / 2.0, dt, 0.0};
    const double b[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};
    const R *u0 = t->tapeU.p + (size_t)k * m->nE, *h0 = t->tapeH.p + (size_t)k * m->nC;
    // Decomposed mesh: in gather form an owned entity's adjoint reads the adjoint variables of the SAME neighbours its forward
    // stencil read, so every array a stage consumes needs plain halo copies -- the lists of the forward exchange serve as they
    // are (no transposed, accumulating exchange): 3 exchanges of the recomputed stage states + 4 of the adjoint variables.
    const bool decomposed = st->dec.ready;
    // forward recompute; the accumulator output of the forward kernel goes to a kbar buffer that is still free
    for (int s = 1; s <= 3; ++s) {
        fused::StageArgs<R> A = stage_args<R>(st, dt, st->cur, s);
        A.uCur = u0; A.hCur = h0; A.uAcc = t->kbU[1].p; A.hAcc = t->kbH[1].p;
        A.uOld = s == 1 ? u0 : t->yU[s - 2].p; A.hOld = s == 1 ? h0 : t->yH[s - 2].p;
        A.uOut = t->yU[s - 1].p; A.hOut = t->yH[s - 1].p;
        if (s == 1) launch_stage<R, 1>(ctx, m, A); else launch_stage<R, 2>(ctx, m, A);
        if (decomposed) halo_exchange_arrays<R>(st, t->yU[s - 1].p, t->yH[s - 1].p);   // the next stage, and the Jacobians, read y_s on halo entities
    }
    adjoint::AdjArgs<R> B;
    B.nE = (int)m->nE; B.nC = (int)m->nC; B.nCown = (int)m->nCo; B.S2T = m->S2T; B.S = m->S;
    B.ce = m->ce.p; B.eoeT = m->eoeT.p; B.eoc = m->eocF.p; B.nEoET = m->nEoET.p; B.nEoC = m->nEoC.p;
    B.blkEdgeStart = m->blkEdgeStart.p;
    B.gdc = fm.gdc.p; B.wfT = fm.wfT.p; B.dv = fm.dv.p; B.invArea = fm.invArea.p; B.H = fm.H.p;
    const int p = t->lamCur;
    B.lamU = t->lamU[p].p; B.lamH = t->lamH[p].p; B.accU = t->lamU[1 - p].p; B.accH = t->lamH[1 - p].p;
    B.bThis = (R)b[3];
    B.kuP = nullptr;
    const int grid = m->fusedBlocks;
    for (int s = 4; s >= 1; --s) {
        B.uY = s == 1 ? u0 : t->yU[s - 2].p; B.hY = s == 1 ? h0 : t->yH[s - 2].p;
        // kbar buffers alternate: stage 4 writes [1], 3 reads [1] writes [0], 2 reads [0] writes [1], 1 reads [1]
        B.kuIn = t->kbU[s & 1].p; B.kqIn = t->kbH[s & 1].p;
        B.kuOut = t->kbU[(s - 1) & 1].p; B.kqOut = t->kbH[(s - 1) & 1].p;
        B.aPrev = s > 1 ? (R)a[s - 2] : R(0); B.bPrev = s > 1 ? (R)b[s - 2] : R(0);
        const bool hex = m->S2T == 10 && m->S == 6;
        const bool hept = m->S2T == 12 && m->S == 7;
        const int mode = s == 4 ? 0 : s > 1 ? 1 : 2;
#define MOKAB_ADJ_LAUNCH(MODE)                                                                                         \
    do {                                                                                                               \
        if (hex) adjoint::k_rk_stage_adj<R, MODE, 10, 6><<<grid, adjoint::kThreads, 0, ctx->stream>>>(B);              \
        else if (hept) adjoint::k_rk_stage_adj<R, MODE, 12, 7><<<grid, adjoint::kThreads, 0, ctx->stream>>>(B);        \
        else adjoint::k_rk_stage_adj<R, MODE, 0, 0><<<grid, adjoint::kThreads, 0, ctx->stream>>>(B);                   \
    } while (0)
        if (mode == 0) MOKAB_ADJ_LAUNCH(0);
        else if (mode == 1) MOKAB_ADJ_LAUNCH(1);
        else MOKAB_ADJ_LAUNCH(2);
#undef MOKAB_ADJ_LAUNCH
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
        if (decomposed) {
            if (s > 1) { if (options().test_drop_dependency != 3) halo_exchange_arrays<R>(st, B.kuOut, B.kqOut); }   // kbar of the previous stage (3: test hook)
            else halo_exchange_arrays<R>(st, B.accU, B.accH);            // lam: what the next reversed step starts from
        }
    }
    t->lamCur = 1 - p;
}

// The same for a multi-level state (Float64, undecomposed): the forward recompute is the column kernel (three launches, ssh of
// every stage state kept because the next stage's single pressure gradient gathers it); every adjoint stage is one launch of the
// single-level kernel PER LEVEL -- Coriolis and thickness flux act level by level -- with the pressure term taken from the level
// sum of kbar_u (k_sum_levels), since ssh = sum_k h_k - H couples the levels of a column.
static void adjoint_step_ml(mokab_state *st, int64_t k)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<double> *t = st->d;
    FusedMesh<double> &fm = fused_of<double>(m);
    const int K = st->K;
    const size_t nE = (size_t)m->nE, nC = (size_t)m->nC;
    const double dt = t->tapeDt[k];
    const double a[4] = {dt / 2.0, dt / 2.0, dt, 0.0};
    const double b[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};
    const double *u0 = t->tapeU.p + (size_t)k * K * nE, *h0 = t->tapeH.p + (size_t)k * K * nC;
    update_ssh(st, h0, t->yS[0].p);                                    // ssh of y_1 = the taped state
    // Decomposed mesh: plain halo copies of every level of what a stage consumes, as in the single-level sweep -- one K + 1 plane
    // message per exchange (the free-surface plane is only meaningful for the stage states; for the adjoint pairs it carries a
    // scratch array).  kuP is formed locally from the exchanged kbar_u, halo edges included.
    const bool decomposed = st->dec.ready;
    fused::StageArgsML A;
    A.nE = (int)nE; A.nC = (int)nC; A.K = K; A.nCown = (int)m->nCo;
    A.ce = m->ce.p; A.eoe = m->eoeF.p; A.eoc = m->eocF.p; A.nEoE = m->nEoE.p; A.nEoC = m->nEoC.p; A.blkEdgeStart = m->blkEdgeStart.p;
    A.gdc = fm.gdc.p; A.wf = fm.wf.p; A.dv = fm.dv.p; A.invArea = fm.invArea.p; A.H = fm.H.p;
    A.uCur = u0; A.hCur = h0; A.uAcc = t->kbU[1].p; A.hAcc = t->kbH[1].p;          // (the accumulator output lands in a kbar buffer that is still free)
    A.f0 = m->f0;
    for (int s = 1; s <= 3; ++s) {
        A.a = a[s - 1]; A.b = b[s - 1];
        A.uOld = s == 1 ? u0 : t->yU[s - 2].p; A.hOld = s == 1 ? h0 : t->yH[s - 2].p; A.sshOld = t->yS[s - 1].p;
        A.uOut = t->yU[s - 1].p; A.hOut = t->yH[s - 1].p; A.sshOut = t->yS[s].p;
        if (s == 1) launch_stage_ml<1>(ctx, m, A); else launch_stage_ml<2>(ctx, m, A);
        if (decomposed) halo_exchange_levels(st, t->yU[s - 1].p, t->yH[s - 1].p, t->yS[s].p);   // y_{s+1} and its free surface on the halo entities
    }
    const int p = t->lamCur;
    const bool hex = m->S2T == 10 && m->S == 6, hept = m->S2T == 12 && m->S == 7;
    const int grid = m->fusedBlocks;
    for (int s = 4; s >= 1; --s) {
        const int mode = s == 4 ? 0 : s > 1 ? 1 : 2;
        const double *kuIn = mode == 0 ? t->lamU[p].p : t->kbU[s & 1].p;
        // level sum of this stage's kbar_u (FIRST: kbar = b_4 lam')
        LAUNCH(ctx, adjoint::k_sum_levels<double>, nblk(m->nE), 256, m->nE, K, mode == 0 ? b[3] : 1.0, kuIn, t->kuP.p);
        for (int lev = 0; lev < K; ++lev) {
            const size_t oe = (size_t)lev * nE, oc = (size_t)lev * nC;
            adjoint::AdjArgs<double> B;
            B.nE = (int)nE; B.nC = (int)nC; B.nCown = (int)m->nCo; B.S2T = m->S2T; B.S = m->S;
            B.ce = m->ce.p; B.eoeT = m->eoeT.p; B.eoc = m->eocF.p; B.nEoET = m->nEoET.p; B.nEoC = m->nEoC.p;
            B.blkEdgeStart = m->blkEdgeStart.p;
            B.gdc = fm.gdc.p; B.wfT = fm.wfT.p; B.dv = fm.dv.p; B.invArea = fm.invArea.p; B.H = fm.H.p;
            B.lamU = t->lamU[p].p + oe; B.lamH = t->lamH[p].p + oc; B.accU = t->lamU[1 - p].p + oe; B.accH = t->lamH[1 - p].p + oc;
            B.bThis = b[3];
            B.uY = (s == 1 ? u0 : t->yU[s - 2].p) + oe; B.hY = (s == 1 ? h0 : t->yH[s - 2].p) + oc;
            B.kuIn = t->kbU[s & 1].p + oe; B.kqIn = t->kbH[s & 1].p + oc;
            B.kuOut = t->kbU[(s - 1) & 1].p + oe; B.kqOut = t->kbH[(s - 1) & 1].p + oc;
            B.aPrev = s > 1 ? a[s - 2] : 0.0; B.bPrev = s > 1 ? b[s - 2] : 0.0;
            B.kuP = t->kuP.p;
#define MOKAB_ADJ_ML_LAUNCH(MODE)                                                                                              \
    do {                                                                                                                       \
        if (hex) adjoint::k_rk_stage_adj<double, MODE, 10, 6, true><<<grid, adjoint::kThreads, 0, ctx->stream>>>(B);          \
        else if (hept) adjoint::k_rk_stage_adj<double, MODE, 12, 7, true><<<grid, adjoint::kThreads, 0, ctx->stream>>>(B);    \
        else adjoint::k_rk_stage_adj<double, MODE, 0, 0, true><<<grid, adjoint::kThreads, 0, ctx->stream>>>(B);               \
    } while (0)
            if (mode == 0) MOKAB_ADJ_ML_LAUNCH(0);
            else if (mode == 1) MOKAB_ADJ_ML_LAUNCH(1);
            else MOKAB_ADJ_ML_LAUNCH(2);
#undef MOKAB_ADJ_ML_LAUNCH
            MOKAB_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        if (decomposed) {
            if (s > 1) halo_exchange_levels(st, t->kbU[(s - 1) & 1].p, t->kbH[(s - 1) & 1].p, t->yS[0].p);
            else halo_exchange_levels(st, t->lamU[1 - p].p, t->lamH[1 - p].p, t->yS[0].p);
        }
    }
    t->lamCur = 1 - p;
}

template <class R>
static void tape_begin(mokab_state *st, int64_t max_steps)
{
    StateT<R> *t = typed<R>(st);
    const mokab_mesh *m = st->mesh;
    if (t->tapeCap < max_steps) {
        t->tapeU.alloc((size_t)max_steps * (size_t)st->K * m->nE);
        t->tapeH.release();                 // RungeKutta4 tapes only: allocated by the first recorded RK4 step (run_rk4_fused)
        t->tapeE.release();                 // ForwardEuler tapes only: allocated by fe_tape_record
        t->tapeCap = max_steps;
    }
    t->tapeDt.clear();
    t->tapeKind = 0;
    t->taping = true;
}

template <class R>
static void adjoint_seed(mokab_state *st, int which)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    StateT<R> *t = typed<R>(st);
    MOKAB_REQUIRE(which == MOKAB_SUM_SSH2, "adjoint_seed: only MOKAB_SUM_SSH2 has a built-in seed; set the MOKAB_D_* fields for others");
    ensure_adjoint<R>(st);
    FusedMesh<R> &fm = fused_of<R>(m);
    t->lamU[t->lamCur].zero(ctx->stream);
    t->lamH[t->lamCur].zero(ctx->stream);
    if constexpr (sizeof(R) == 8) {
        if (st->K > 1) {   // multi-level: ssh = (sum of the column) - H, from the current layerThickness
            update_ssh(st, t->h[st->cur].p, t->ssh[st->cur].p);
            LAUNCH(ctx, adjoint::k_seed_ssh2_array<R>, nblk(m->nC), 256, m->nC, (const R *)t->ssh[st->cur].p, t->dSsh.p);
            return;
        }
    }
    if (t->tapeKind == 2)
        LAUNCH(ctx, adjoint::k_seed_ssh2_array<R>, nblk(m->nC), 256, m->nC, (const R *)t->ssh[st->cur].p, t->dSsh.p);
    else
        LAUNCH(ctx, adjoint::k_seed_ssh2<R>, nblk(m->nC), 256, m->nC, (const R *)t->h[st->cur].p, (const R *)fm.H.p, t->dSsh.p);
}

template <class R>
static void adjoint_run(mokab_state *st)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    StateT<R> *t = typed<R>(st);
    MOKAB_REQUIRE(t->tapeKind != 2, "adjoint_rk4: the tape holds ForwardEuler steps (use mokab_adjoint_forward_euler)");
    ensure_adjoint<R>(st);
    t->taping = false;
    t->tapeKind = 0;
    if constexpr (sizeof(R) == 8) {
        if (st->K > 1) {
            LAUNCH(ctx, adjoint::k_fold_dssh_levels<R>, nblk(m->nC), 256, m->nC, st->K, t->dSsh.p, t->lamH[t->lamCur].p);
            if (st->dec.ready) ensure_level_buffers(st);
            if (st->dec.ready) halo_exchange_levels(st, t->lamU[t->lamCur].p, t->lamH[t->lamCur].p, t->yS[0].p);   // the owners' seeds on the halo copies
            for (int64_t k = (int64_t)t->tapeDt.size() - 1; k >= 0; --k) adjoint_step_ml(st, k);
            t->tapeDt.clear();
            return;
        }
    }
    LAUNCH(ctx, adjoint::k_fold_dssh<R>, nblk(m->nC), 256, m->nC, t->dSsh.p, t->lamH[t->lamCur].p);
    if (st->dec.ready) halo_exchange_arrays<R>(st, t->lamU[t->lamCur].p, t->lamH[t->lamCur].p);   // the owners' seeds on the halo copies
    for (int64_t k = (int64_t)t->tapeDt.size() - 1; k >= 0; --k) adjoint_step<R>(st, k);
    t->tapeDt.clear();
}

// Reverse sweep over recorded ForwardEuler steps (adjoint::k_fe_step_adj): one launch per reversed step.
static void adjoint_run_fe(mokab_state *st)
{
    mokab_ctx *ctx = st->ctx; mokab_mesh *m = const_cast<mokab_mesh *>(st->mesh);
    MOKAB_REQUIRE(st->dtype == MOKAB_F64, "adjoint_forward_euler: ForwardEuler is Float64 only (PrognosticVars.jl:91-93)");
    StateT<double> *t = st->d;
    MOKAB_REQUIRE(t->tapeKind != 1, "adjoint_forward_euler: the tape holds RungeKutta4 steps (use mokab_adjoint_rk4)");
    MOKAB_REQUIRE(st->K == 1, "adjoint_forward_euler: single-level states only (the multi-level reverse mode is RungeKutta4)");
    ensure_adjoint<double>(st);
    FusedMesh<double> &fm = fused_of<double>(m);
    for (int i = 0; i < 2; ++i)
        if (t->lamS[i].n == 0) { t->lamS[i].alloc(m->nC); t->lamE[i].alloc(m->nE); t->lamQ[i].alloc(m->nC); }
    t->taping = false;
    int p = t->lamCur;
    const int64_t nmax = std::max(m->nC, m->nE);
    LAUNCH(ctx, adjoint::k_fe_adj_begin, nblk(nmax), 256, m->nC, m->nE, (const double *)fm.invArea.p, (const double *)t->dSsh.p,
           (const double *)t->lamH[p].p, t->lamS[p].p, t->lamE[p].p, t->lamQ[p].p);
    // Decomposed mesh: a reversed step gathers lamU (transposed Coriolis stencil, pressure term), lamE and q on the neighbours its
    // forward stencil read: plain halo copies of those three arrays per step, over the lists of the forward exchange.
    const bool decomposed = st->dec.ready;
    if (decomposed) halo_exchange_arrays<double>(st, t->lamU[p].p, t->lamQ[p].p);
    adjoint::FeAdjArgs A;
    A.nE = (int)m->nE; A.nC = (int)m->nC; A.nCown = (int)m->nCo; A.S2T = m->S2T; A.S = m->S;
    A.ce = m->ce.p; A.eoeT = m->eoeT.p; A.eoc = m->eocF.p; A.nEoET = m->nEoET.p; A.nEoC = m->nEoC.p;
    A.blkEdgeStart = m->blkEdgeStart.p;
    A.gdc = fm.gdc.p; A.wT = m->woeT.p; A.dv = m->dv.p; A.invArea = fm.invArea.p;
    const bool hex = m->S2T == 10 && m->S == 6, hept = m->S2T == 12 && m->S == 7;
    for (int64_t k = (int64_t)t->tapeDt.size() - 1; k >= 0; --k) {
        A.dt = t->tapeDt[k];
        A.uN = t->tapeU.p + (size_t)k * m->nE; A.hEN = t->tapeE.p + (size_t)k * m->nE;
        A.lamU = t->lamU[p].p; A.lamH = t->lamH[p].p; A.lamS = t->lamS[p].p; A.lamE = t->lamE[p].p; A.qIn = t->lamQ[p].p;
        A.outU = t->lamU[1 - p].p; A.outH = t->lamH[1 - p].p; A.outS = t->lamS[1 - p].p; A.outE = t->lamE[1 - p].p;
        A.qOut = t->lamQ[1 - p].p;
        if (hex) adjoint::k_fe_step_adj<10, 6><<<m->fusedBlocks, adjoint::kThreads, 0, ctx->stream>>>(A);
        else if (hept) adjoint::k_fe_step_adj<12, 7><<<m->fusedBlocks, adjoint::kThreads, 0, ctx->stream>>>(A);
        else adjoint::k_fe_step_adj<0, 0><<<m->fusedBlocks, adjoint::kThreads, 0, ctx->stream>>>(A);
        MOKAB_CUDA(cudaGetLastError());
        ctx->launches++;
        if (decomposed) {
            halo_exchange_arrays<double>(st, A.outU, A.qOut);
            halo_exchange_arrays<double>(st, A.outE, A.outH);
        }
        p = 1 - p;
    }
    t->lamCur = p;
    // d_Prog.ssh[end]: the gradient with respect to the initial ssh array (read by the first step's pressure gradient only)
    MOKAB_CUDA(cudaMemcpyAsync(t->dSsh.p, t->lamS[p].p, m->nC * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    t->tapeDt.clear();
    t->tapeKind = 0;
}

// ---- halo exchange by direct peer stores ------------------------------------------------------------------------
// What one rank tells the others (mokab_p2p_export): where its state arrays and its arrival counters live.  Ranks in
// other processes map them with CUDA IPC; ranks emulated inside one process (tests) use the addresses as they are.
struct P2PBlob {
    int64_t pid;
    int32_t rank, dtype;
    void *base;                       // the state's slab
    uint64_t off[10];                 // u0 u1 uP0 uP1 h0 h1 hP0 hP1 arrival counters, flag-in-data receive area
    uint64_t llSlots;                 // slots of that area
    cudaIpcMemHandle_t handle;
};

// How long a halo wait spins before it gives up and raises the state's error flag (MOKAB_P2P_TIMEOUT_S, default 20 s -- long
// enough for rank skew from lazy module loading, graph instantiation or host I/O; a wait that finds the flag already raised
// does not spin at all, so after one time-out the remaining launches of the call drain quickly).
static long long p2p_timeout_cycles()
{
    static const long long cycles = [] {
        const char *e = getenv("MOKAB_P2P_TIMEOUT_S");
        const double s = e && *e ? atof(e) : 20.0;
        return (long long)(std::max(0.001, s) * 1.9e9);
    }();
    return cycles;
}

template <class R>
static void p2p_export(mokab_state *st, P2PBlob *b)
{
    StateT<R> *t = typed<R>(st);
    mokab_state::P2P &x = st->p2p;
    memset(b, 0, sizeof(*b));
    b->pid = (int64_t)getpid();
    b->dtype = st->dtype;
    if (!x.exported) {
        x.arrival = (unsigned long long *)(t->slab.p + t->slabOff[8]);       // zeroed with the slab
        x.expect.alloc(kP2PCounters); x.expect.zero(st->ctx->stream);
        x.done.alloc(1); x.done.zero(st->ctx->stream);
        x.error.alloc(1); x.error.zero(st->ctx->stream);
        MOKAB_CUDA(cudaStreamSynchronize(st->ctx->stream));
        x.exported = true;
    }
    b->base = t->slab.p;
    for (int i = 0; i < 10; ++i) b->off[i] = t->slabOff[i];
    b->llSlots = t->llSlots;
#ifndef MOKAB_SIM
    MOKAB_CUDA(cudaIpcGetMemHandle(&b->handle, t->slab.p));
#endif
}

template <class R>
static void p2p_setup(mokab_state *st, int rank, int nranks, const P2PBlob *blobs, int nrecv, const int32_t *recv_ranks,
                      const int64_t *counts, const int32_t *dst_idx, int nsend, const int32_t *send_ranks)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    mokab_state::P2P &x = st->p2p;
    MOKAB_REQUIRE(x.exported, "p2p_setup: call mokab_p2p_export first");
    MOKAB_REQUIRE(nranks >= 1 && nranks <= kP2PCounters && rank >= 0 && rank < nranks, "p2p_setup: bad rank / nranks");
    MOKAB_REQUIRE(nrecv >= 0 && nrecv <= p2p::kMaxPeers && nsend >= 0 && nsend <= p2p::kMaxPeers, "p2p_setup: too many peers");
    int64_t total = 0;
    for (int i = 0; i < nrecv; ++i) total += counts[i];
    MOKAB_REQUIRE(total == (int64_t)m->haloSend.n, "p2p_setup: push counts do not add up to the send list of mokab_halo_setup");
    x.rank = rank; x.nranks = nranks; x.nPush = total;
    x.recvRanks.assign(recv_ranks, recv_ranks + nrecv);
    x.sendRanks.assign(send_ranks, send_ranks + nsend);
    x.peerLLHost.clear(); x.peerLLSlots.clear();
    // peer pointers: targets 0/1 = provisional buffers P0/P1, 2/3 = time levels 0/1
    std::vector<void *> pH((size_t)4 * std::max(nrecv, 1), nullptr), pU((size_t)4 * std::max(nrecv, 1), nullptr);
    std::vector<unsigned long long *> arr(std::max(nrecv, 1), nullptr);
    const int64_t me = (int64_t)getpid();
    for (int i = 0; i < nrecv; ++i) {
        const P2PBlob &b = blobs[recv_ranks[i]];
        MOKAB_REQUIRE(b.rank == recv_ranks[i] && b.dtype == st->dtype, "p2p_setup: blob does not belong to the rank / precision it is filed under");
        void *base = b.base;                         // same process (emulated ranks): the address is valid as it is
        if (b.pid != me) {
#ifdef MOKAB_SIM
            throw Error("p2p_setup: the simulated runtime has no inter-process mappings");
#else
            MOKAB_CUDA(cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
            x.opened.push_back(base);
#endif
        }
        void *ptr[10];
        for (int k = 0; k < 10; ++k) ptr[k] = (unsigned char *)base + b.off[k];
        x.peerLLHost.push_back((unsigned long long *)ptr[9]);
        x.peerLLSlots.push_back((int64_t)b.llSlots);
        pU[0 * nrecv + i] = ptr[2]; pU[1 * nrecv + i] = ptr[3]; pU[2 * nrecv + i] = ptr[0]; pU[3 * nrecv + i] = ptr[1];
        pH[0 * nrecv + i] = ptr[6]; pH[1 * nrecv + i] = ptr[7]; pH[2 * nrecv + i] = ptr[4]; pH[3 * nrecv + i] = ptr[5];
        arr[i] = (unsigned long long *)ptr[8] + rank;
    }
    std::vector<uint8_t> slot((size_t)total);
    {
        int64_t k = 0;
        for (int i = 0; i < nrecv; ++i)
            for (int64_t j = 0; j < counts[i]; ++j) slot[k++] = (uint8_t)i;
    }
    std::vector<int32_t> dst(dst_idx, dst_idx + total), snd(send_ranks, send_ranks + nsend);
    x.peerH.upload(pH, ctx->stream); x.peerU.upload(pU, ctx->stream); x.arrivalAt.upload(arr, ctx->stream);
    x.slot.upload(slot, ctx->stream); x.dst.upload(dst, ctx->stream); x.senders.upload(snd, ctx->stream);
    // the same list, entity-major, for the boundary launch that pushes what it has just computed
    {
        const int64_t nC = m->nC, nE = m->nE;
        std::vector<int32_t> sE(nE + 1, 0), sC(nC + 1, 0);
        for (int64_t k = 0; k < total; ++k) {
            const int32_t i = m->hHaloSend[k];
            if (i < nC) sC[i + 1]++; else sE[i - nC + 1]++;
        }
        for (int64_t i = 0; i < nC; ++i) sC[i + 1] += sC[i];
        for (int64_t i = 0; i < nE; ++i) sE[i + 1] += sE[i];
        std::vector<int32_t> dE(std::max<int64_t>(sE[nE], 1)), dC(std::max<int64_t>(sC[nC], 1)), fillE(sE.begin(), sE.end() - 1), fillC(sC.begin(), sC.end() - 1);
        std::vector<uint8_t> lE(dE.size()), lC(dC.size());
        for (int64_t k = 0; k < total; ++k) {
            const int32_t i = m->hHaloSend[k], d = dst[k];
            if (i < nC) {
                MOKAB_REQUIRE(d >= 0, "p2p_setup: a cell is sent to an edge slot");
                const int32_t at = fillC[i]++;
                dC[at] = d; lC[at] = slot[k];
            } else {
                MOKAB_REQUIRE(d < 0, "p2p_setup: an edge is sent to a cell slot");
                const int32_t at = fillE[i - nC]++;
                dE[at] = -d - 1; lE[at] = slot[k];
            }
        }
        x.startE.upload(sE, ctx->stream); x.startC.upload(sC, ctx->stream);
        x.dstE.upload(dE, ctx->stream); x.dstC.upload(dC, ctx->stream);
        x.slotE.upload(lE, ctx->stream); x.slotC.upload(lC, ctx->stream);
        std::vector<unsigned char> desc(4 * sizeof(fused::PushStage<R>));
        for (int tgt = 0; tgt < 4; ++tgt) {
            fused::PushStage<R> P;
            P.startE = x.startE.p; P.startC = x.startC.p; P.dstE = x.dstE.p; P.dstC = x.dstC.p; P.slotE = x.slotE.p; P.slotC = x.slotC.p;
            P.peerU = (R *const *)x.peerU.p + (size_t)tgt * nrecv; P.peerH = (R *const *)x.peerH.p + (size_t)tgt * nrecv;
            P.arrivalAt = (unsigned long long *const *)x.arrivalAt.p; P.nrecv = nrecv;
            P.senders = x.senders.p; P.nsend = nsend; P.arrival = x.arrival; P.expect = x.expect.p;
            P.done = x.done.p; P.error = x.error.p; P.timeout_cycles = p2p_timeout_cycles();
            memcpy(desc.data() + tgt * sizeof(P), &P, sizeof(P));
        }
        x.stageDesc.upload(desc, ctx->stream);
    }
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    x.ready = true;
}

// which of the four peer targets holds the output of RK stage `stage` (cf. stage_output)
static int p2p_target(const mokab_state *st, int stage)
{
    switch (stage) {
    case 1: case 3: return 0;
    case 2: return 1;
    case 0: return 2 + st->cur;
    default: return 2 + (1 - st->cur);
    }
}

template <class R>
static void p2p_push(mokab_state *st, int stage, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    mokab_state::P2P &x = st->p2p;
    const int nrecv = (int)x.recvRanks.size();
    if (nrecv == 0) return;
    cudaStream_t s = stream ? stream : ctx->stream;
    if (x.nPush == 0) {
        MOKAB_LAUNCH_ON(p2p::k_halo_signal, 1, 32, s, nrecv, (unsigned long long *const *)x.arrivalAt.p);
    } else {
        R *u, *h;
        stage_output<R>(st, stage, &u, &h);
        const int tgt = p2p_target(st, stage);
        p2p::PushArgs<R> A;
        A.n = (int)x.nPush; A.nC = (int)m->nC; A.src = m->haloSend.p; A.dst = x.dst.p; A.slot = x.slot.p;
        A.h = h; A.u = u;
        A.peerH = (R *const *)x.peerH.p + (size_t)tgt * nrecv; A.peerU = (R *const *)x.peerU.p + (size_t)tgt * nrecv;
        A.done = x.done.p; A.arrival = (unsigned long long *const *)x.arrivalAt.p; A.nrecv = nrecv;
        MOKAB_LAUNCH_ON(p2p::k_halo_push<R>, nblk(A.n), 256, s, A);
    }
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

#ifdef MOKAB_SIM
// the gate a PUSH launch spins in on hardware, as a stream operation of the simulated runtime
static void p2p_gate_sim(mokab_state *st, cudaStream_t s)
{
    mokab_state::P2P &x = st->p2p;
    const int nsend = (int)x.sendRanks.size();
    if (nsend == 0) return;
    const int32_t *senders = x.senders.p;
    const unsigned long long *arrival = x.arrival, *expect = x.expect.p;
    mokab_sim::enqueue_try(s, "p2p gate (arrival >= expect)", [=]() {
        for (int i = 0; i < nsend; ++i)
            if (arrival[senders[i]] < expect[senders[i]]) return false;
        return true;
    });
}
#endif

static void p2p_wait_arrivals(mokab_state *st, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx;
    mokab_state::P2P &x = st->p2p;
    const int nsend = (int)x.sendRanks.size();
    if (nsend == 0) return;
    cudaStream_t s = stream ? stream : ctx->stream;
#ifdef MOKAB_SIM
    p2p_gate_sim(st, s);
#else
    MOKAB_LAUNCH_ON(p2p::k_halo_wait_arrivals, 1, p2p::kMaxPeers, s, nsend, (const int32_t *)x.senders.p, (const unsigned long long *)x.arrival,
                    (const unsigned long long *)x.expect.p, (int *)x.error.p, p2p_timeout_cycles());
    MOKAB_CUDA(cudaGetLastError());
#endif
    ctx->launches++;
}

static void p2p_wait(mokab_state *st, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx;
    mokab_state::P2P &x = st->p2p;
    const int nsend = (int)x.sendRanks.size();
    if (nsend == 0) return;
    cudaStream_t s = stream ? stream : ctx->stream;
#ifdef MOKAB_SIM
    // a spinning kernel cannot run on a simulator that executes kernels to completion: the same predicate becomes a
    // stream operation that is retried until it holds
    const int32_t *senders = x.senders.p;
    const unsigned long long *arrival = x.arrival;
    unsigned long long *expect = x.expect.p;
    mokab_sim::enqueue_try(s, "p2p::k_halo_wait", [=]() {
        for (int i = 0; i < nsend; ++i)
            if (!p2p::sender_ready(arrival, expect, senders[i])) return false;
        for (int i = 0; i < nsend; ++i) expect[senders[i]] += 1ull;
        return true;
    });
#else
    MOKAB_LAUNCH_ON(p2p::k_halo_wait, 1, p2p::kMaxPeers, s, nsend, (const int32_t *)x.senders.p, (const unsigned long long *)x.arrival,
                    (unsigned long long *)x.expect.p, (int *)x.error.p, p2p_timeout_cycles());
    MOKAB_CUDA(cudaGetLastError());
#endif
    ctx->launches++;
}

// ---- flag-in-data exchange (MOKAB_HALO_P2P_LL) ---------------------------------------------------------------------------
// after p2p_setup (peers mapped): where this rank's items land at every receiver.  recv_base[i] = the first slot of this
// rank's segment in receiver i's receive list, credit_slot[i] = the slot receiver i keeps for this rank's credit packet.
template <class R>
static void p2p_setup_ll(mokab_state *st, const int64_t *counts, const int64_t *recv_base, const int64_t *credit_slot)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    mokab_state::P2P &x = st->p2p;
    StateT<R> *t = typed<R>(st);
    MOKAB_REQUIRE(x.ready, "p2p_setup_ll: call p2p_setup first");
    const int npeer = (int)x.recvRanks.size();
    MOKAB_REQUIRE(x.sendRanks == x.recvRanks, "p2p_setup_ll: the peer relation must be symmetric");
    MOKAB_REQUIRE(t->llSlots >= (size_t)m->haloRecv.n + (size_t)npeer,
                  "MOKAB_HALO_P2P_LL: the state has no receive area (create the state after mokab_halo_setup; single-level states)");
    std::vector<int32_t> dst;
    std::vector<uint8_t> slot;
    for (int i = 0; i < npeer; ++i)
        for (int64_t j = 0; j < counts[i]; ++j) { dst.push_back((int32_t)(recv_base[i] + j)); slot.push_back((uint8_t)i); }
    MOKAB_REQUIRE((int64_t)dst.size() == (int64_t)m->haloSend.n, "p2p_setup_ll: counts do not add up to the send list");
    for (int i = 0; i < npeer; ++i) {
        MOKAB_REQUIRE(credit_slot[i] >= 0 && credit_slot[i] < x.peerLLSlots[i] && recv_base[i] + counts[i] <= x.peerLLSlots[i],
                      "p2p_setup_ll: a slot lies outside the receiver's area");
        dst.push_back((int32_t)credit_slot[i]); slot.push_back((uint8_t)i);
    }
    x.llSend = (int)dst.size(); x.llRecv = (int)m->haloRecv.n + npeer;
    if (dst.empty()) { dst.push_back(0); slot.push_back(0); }
    std::vector<unsigned long long *> pl(x.peerLLHost.begin(), x.peerLLHost.end());
    if (pl.empty()) pl.push_back(nullptr);
    x.llDst.upload(dst, ctx->stream); x.llSlot.upload(slot, ctx->stream); x.peerLL.upload(pl, ctx->stream);
    x.llCtr.alloc(4); x.llCtr.zero(ctx->stream);
    x.llArea = (unsigned long long *)(t->slab.p + t->slabOff[9]);        // zeroed with the slab: exchange number 0 = nothing yet
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    x.llReady = true;
}

// (any (edge array, cell array) pair: the packets carry values, not addresses -- the receiver scatters into ITS arrays)
template <class R>
static void p2p_push_ll_arrays(mokab_state *st, const R *u, const R *h, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    mokab_state::P2P &x = st->p2p;
    MOKAB_REQUIRE(x.llReady, "halo exchange (MOKAB_HALO_P2P_LL): not set up");
    if (x.llSend == 0) return;
    cudaStream_t s = stream ? stream : ctx->stream;
    p2p::LLPushArgs<R> A;
    A.n = x.llSend; A.nReal = (int)m->haloSend.n; A.nC = (int)m->nC; A.src = m->haloSend.p; A.llDst = x.llDst.p; A.slot = x.llSlot.p;
    A.h = h; A.u = u; A.peerLL = (unsigned long long *const *)x.peerLL.p; A.seq = x.llCtr.p; A.done = x.llCtr.p + 2;
    MOKAB_LAUNCH_ON(p2p::k_halo_push_ll<R>, nblk(A.n), 256, s, A);
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

template <class R>
static void p2p_push_ll(mokab_state *st, int stage, cudaStream_t stream)
{
    R *u, *h;
    stage_output<R>(st, stage, &u, &h);
    p2p_push_ll_arrays<R>(st, u, h, stream);
}

template <class R>
static void p2p_wait_ll_arrays(mokab_state *st, R *u, R *h, cudaStream_t stream)
{
    mokab_ctx *ctx = st->ctx; const mokab_mesh *m = st->mesh;
    mokab_state::P2P &x = st->p2p;
    MOKAB_REQUIRE(x.llReady, "halo exchange (MOKAB_HALO_P2P_LL): not set up");
    if (x.llRecv == 0) return;
    cudaStream_t s = stream ? stream : ctx->stream;
    p2p::LLWaitArgs<R> A;
    A.n = x.llRecv; A.nReal = (int)m->haloRecv.n; A.nC = (int)m->nC; A.idx = m->haloRecv.p; A.ll = x.llArea; A.h = h; A.u = u;
    A.seq = x.llCtr.p + 1; A.done = x.llCtr.p + 3; A.error = x.error.p; A.timeout_cycles = p2p_timeout_cycles();
#ifdef MOKAB_SIM
    // a spinning kernel cannot run on a simulator that executes kernels to completion: the same predicate becomes a stream
    // operation that is retried until it holds, then the (non-spinning) kernel unpacks
    {
        const unsigned long long *ll = x.llArea;
        const unsigned int *seq = x.llCtr.p + 1;
        const int n = x.llRecv;
        mokab_sim::enqueue_try(s, "p2p::k_halo_wait_ll", [=]() {
            const unsigned int want = *seq + 1u;
            unsigned long long bits;
            for (int k = 0; k < n; ++k)
                if (!p2p::ll_arrived<R>(ll + p2p::ll_slot(k, want), want, &bits)) return false;
            return true;
        });
    }
#endif
    MOKAB_LAUNCH_ON(p2p::k_halo_wait_ll<R>, nblk(A.n), 256, s, A);
    MOKAB_CUDA(cudaGetLastError());
    ctx->launches++;
}

template <class R>
static void p2p_wait_ll(mokab_state *st, int stage, cudaStream_t stream)
{
    R *u, *h;
    stage_output<R>(st, stage, &u, &h);
    p2p_wait_ll_arrays<R>(st, u, h, stream);
}

// stand-alone operators on host arrays -------------------------------------------------------------------
struct OpBufs {
    DevBuf<double> in, in_p, out, out_p;
};

}  // namespace mokab

using namespace mokab;

// =================================================================================================================
extern "C" {

const char *mokab_last_error(void) { return g_last_error.c_str(); }
int mokab_version(void) { return MOKAB_VERSION; }

int mokab_init(int device, mokab_ctx **out)
{
    return guarded([&] {
        MOKAB_REQUIRE(out, "mokab_init: out is NULL");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw Error(std::string("mokab_init: no CUDA device available (") + cudaGetErrorString(e) +
                        "); libmoka_b200 has no CPU fallback");
        MOKAB_REQUIRE(device >= 0 && device < n, "mokab_init: device index out of range");
        MOKAB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        MOKAB_CUDA(cudaGetDeviceProperties(&prop, device));
        MOKAB_REQUIRE(prop.major >= 10, std::string("mokab_init: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                            std::to_string(prop.minor) + "; this library is built for sm_100a only");
        auto *c = new mokab_ctx();
        c->device = device;
        c->num_sms = prop.multiProcessorCount;
        MOKAB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        MOKAB_CUDA(cudaEventCreate(&c->ev0));
        MOKAB_CUDA(cudaEventCreate(&c->ev1));
        *out = c;
    });
}

int mokab_finalize(mokab_ctx *ctx)
{
    return guarded([&] {
        if (!ctx) return;
        ctx->bind();
        cudaStreamSynchronize(ctx->stream);
        cudaEventDestroy(ctx->ev0);
        cudaEventDestroy(ctx->ev1);
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
    });
}

int mokab_synchronize(mokab_ctx *ctx)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx, "synchronize: ctx is NULL");
        ctx->bind();
        MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int mokab_set_stream(mokab_ctx *ctx, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx, "set_stream: ctx is NULL");
        ctx->bind();
        MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    });
}

int mokab_timer_start(mokab_ctx *ctx)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx, "timer_start: ctx is NULL");
        ctx->bind();
        MOKAB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    });
}

int mokab_timer_stop(mokab_ctx *ctx, double *elapsed_ms)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && elapsed_ms, "timer_stop: NULL argument");
        ctx->bind();
        MOKAB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        MOKAB_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        MOKAB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *elapsed_ms = (double)ms;
    });
}

int mokab_launch_count(mokab_ctx *ctx, int64_t *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && out, "launch_count: NULL argument");
        *out = ctx->launches;
    });
}

int mokab_host_alloc(void **out, int64_t bytes)
{
    return guarded([&] {
        MOKAB_REQUIRE(out && bytes >= 0, "host_alloc: bad argument");
        MOKAB_CUDA(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    });
}

int mokab_host_free(void *p)
{
    return guarded([&] {
        if (p) MOKAB_CUDA(cudaFreeHost(p));
    });
}

// ---- mesh ------------------------------------------------------------------------------------------------------
int mokab_mesh_create(mokab_ctx *ctx, const mokab_mesh_desc *desc, uint32_t flags, mokab_mesh **out)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && desc && out, "mesh_create: NULL argument");
        ctx->bind();
        HostMesh hm;
        build_host_mesh(*desc, flags, hm);
        auto *m = new mokab_mesh();
        try {
            upload_mesh(ctx, hm, m);
        } catch (...) {
            delete m;
            throw;
        }
        *out = m;
    });
}

int mokab_mesh_destroy(mokab_mesh *mesh)
{
    return guarded([&] {
        if (!mesh) return;
        mesh->ctx->bind();
        cudaStreamSynchronize(mesh->ctx->stream);
        delete mesh;
    });
}

int mokab_mesh_get_perm(const mokab_mesh *mesh, int kind, int32_t *perm_out)
{
    return guarded([&] {
        MOKAB_REQUIRE(mesh && perm_out, "mesh_get_perm: NULL argument");
        const std::vector<int32_t> *p = kind == MOKAB_CELLS ? &mesh->permC : kind == MOKAB_EDGES ? &mesh->permE
                                        : kind == MOKAB_VERTICES ? &mesh->permV : nullptr;
        MOKAB_REQUIRE(p, "mesh_get_perm: unknown entity kind");
        if (!p->empty()) memcpy(perm_out, p->data(), p->size() * sizeof(int32_t));
    });
}

int mokab_mesh_device_bytes(const mokab_mesh *mesh, int64_t *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(mesh && out, "mesh_device_bytes: NULL argument");
        *out = mesh->device_bytes();
    });
}

// ---- state -----------------------------------------------------------------------------------------------------
int mokab_state_create(mokab_ctx *ctx, const mokab_mesh *mesh, int dtype, mokab_state **out)
{
    return mokab_state_create_levels(ctx, mesh, dtype, 1, out);
}

int mokab_state_levels(const mokab_state *state, int *nVertLevels)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && nVertLevels, "state_levels: NULL argument");
        *nVertLevels = state->K;
    });
}

int mokab_state_create_levels(mokab_ctx *ctx, const mokab_mesh *mesh, int dtype, int nVertLevels, mokab_state **out)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && mesh && out, "state_create: NULL argument");
        MOKAB_REQUIRE(nVertLevels >= 1 && nVertLevels <= 1024, "state_create: nVertLevels must be in 1..1024");
        MOKAB_REQUIRE(nVertLevels == 1 || dtype == MOKAB_F64, "state_create: multi-level states are Float64 (PrognosticVars.jl:91-93)");
        MOKAB_REQUIRE(mesh->ctx == ctx, "state_create: mesh belongs to a different context (src/Architectures.jl:27-33)");
        MOKAB_REQUIRE(dtype == MOKAB_F64 || dtype == MOKAB_F32, "state_create: dtype must be MOKAB_F64 or MOKAB_F32");
        ctx->bind();
        auto *st = new mokab_state();
        st->ctx = ctx; st->mesh = mesh; st->dtype = dtype; st->K = nVertLevels;
        try {
            if (dtype == MOKAB_F64) alloc_state<double>(st); else alloc_state<float>(st);
            // the fused-form mesh arrays of this precision, complete before any stepper can be launched on any stream
            if (dtype == MOKAB_F64) ensure_fused<double>(const_cast<mokab_mesh *>(mesh));
            else ensure_fused<float>(const_cast<mokab_mesh *>(mesh));
        } catch (...) {
            delete st;
            throw;
        }
        *out = st;
    });
}

int mokab_state_destroy(mokab_state *state)
{
    return guarded([&] {
        if (!state) return;
        state->ctx->bind();
        cudaStreamSynchronize(state->ctx->stream);
        delete state;
    });
}

int mokab_state_set(mokab_state *state, int field, const void *host)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && host, "state_set: NULL argument");
        MOKAB_REQUIRE(field >= MOKAB_SSH && field <= MOKAB_D_LAYER_THICKNESS, "state_set: unknown field id");
        state->ctx->bind();
        fe_materialize(state);
        if (state->dtype == MOKAB_F64) state_set<double>(state, field, host); else state_set<float>(state, field, host);
    });
}

int mokab_state_get(mokab_state *state, int field, void *host)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && host, "state_get: NULL argument");
        MOKAB_REQUIRE(field >= MOKAB_SSH && field <= MOKAB_D_LAYER_THICKNESS, "state_get: unknown field id");
        state->ctx->bind();
        if (field >= MOKAB_LAYER_THICKNESS_EDGE && field <= MOKAB_TEND_LAYER_THICKNESS) fe_materialize(state);
        if (state->dtype == MOKAB_F64) state_get<double>(state, field, host); else state_get<float>(state, field, host);
    });
}

int mokab_state_set_async(mokab_state *state, int field, const void *host_pinned)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && host_pinned, "state_set_async: NULL argument");
        MOKAB_REQUIRE(field >= MOKAB_SSH && field <= MOKAB_D_LAYER_THICKNESS, "state_set_async: unknown field id");
        state->ctx->bind();
        fe_materialize(state);
        if (state->dtype == MOKAB_F64) state_set_async<double>(state, field, host_pinned);
        else state_set_async<float>(state, field, host_pinned);
    });
}

int mokab_state_get_async(mokab_state *state, int field, void *host_pinned)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && host_pinned, "state_get_async: NULL argument");
        MOKAB_REQUIRE(field >= MOKAB_SSH && field <= MOKAB_D_LAYER_THICKNESS, "state_get_async: unknown field id");
        state->ctx->bind();
        if (field >= MOKAB_LAYER_THICKNESS_EDGE && field <= MOKAB_TEND_LAYER_THICKNESS) fe_materialize(state);
        if (state->dtype == MOKAB_F64) state_get_async<double>(state, field, host_pinned);
        else state_get_async<float>(state, field, host_pinned);
    });
}

int mokab_state_synchronize(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "state_synchronize: state is NULL");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) state_synchronize<double>(state); else state_synchronize<float>(state);
    });
}

// ---- src/ocn entry points -----------------------------------------------------------------------------------------
int mokab_diagnostic_compute(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "diagnostic_compute: state is NULL");
        require_f64(state, "diagnostic_compute");
        state->ctx->bind();
        fe_materialize(state);
        diag_compute(state, state->d->u[state->cur].p, state->d->h[state->cur].p);
    });
}

int mokab_diagnostic_compute_consistent(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "diagnostic_compute_consistent: state is NULL");
        require_f64(state, "diagnostic_compute_consistent");
        state->ctx->bind();
        fe_materialize(state);
        diag_consistent(state, state->d->u[state->cur].p, state->d->h[state->cur].p);
    });
}

int mokab_compute_normal_velocity_tendency(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "compute_normal_velocity_tendency: state is NULL");
        require_f64(state, "compute_normal_velocity_tendency");
        state->ctx->bind();
        fe_materialize(state);
        tend_u(state, state->d->ssh[state->cur].p, state->d->u[state->cur].p);
    });
}

int mokab_compute_layer_thickness_tendency(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "compute_layer_thickness_tendency: state is NULL");
        require_f64(state, "compute_layer_thickness_tendency");
        state->ctx->bind();
        fe_materialize(state);
        tend_h(state, state->d->flux.p);
    });
}

static void op_common(mokab_ctx *ctx, const mokab_mesh *mesh, const double *in, int64_t nin, const int32_t *perm_in,
                      double *out, int64_t nout, const int32_t *perm_out, bool out_is_inout, OpBufs &B)
{
    B.in.alloc(nin); B.in_p.alloc(nin); B.out.alloc(nout); B.out_p.alloc(nout);
    MOKAB_CUDA(cudaMemcpyAsync(B.in.p, in, nin * 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_permute_in<double>, nblk(nin), 256, nin, perm_in, (const double *)B.in.p, B.in_p.p);
    if (out_is_inout) {
        MOKAB_CUDA(cudaMemcpyAsync(B.out.p, out, nout * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, k_permute_in<double>, nblk(nout), 256, nout, perm_out, (const double *)B.out.p, B.out_p.p);
    }
}
static void op_finish(mokab_ctx *ctx, double *out, int64_t nout, const int32_t *perm_out, OpBufs &B)
{
    LAUNCH(ctx, k_permute_out<double>, nblk(nout), 256, nout, perm_out, (const double *)B.out_p.p, B.out.p);
    MOKAB_CUDA(cudaMemcpyAsync(out, B.out.p, nout * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MOKAB_CUDA(cudaStreamSynchronize(ctx->stream));
}

int mokab_gradient_on_edge(mokab_ctx *ctx, const mokab_mesh *m, const double *scalar_cell, double *grad_edge)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && m && scalar_cell && grad_edge, "gradient_on_edge: NULL argument");
        ctx->bind();
        OpBufs B;
        op_common(ctx, m, scalar_cell, m->nC, m->dPermC.p, grad_edge, m->nE, m->dPermE.p, false, B);
        LAUNCH(ctx, ref::k_gradient_on_edge, nblk(m->nE), 256, (int)m->nE, m->ce.p, m->dc.p, (const double *)B.in_p.p, B.out_p.p);
        op_finish(ctx, grad_edge, m->nE, m->dPermE.p, B);
    });
}

int mokab_divergence_on_cell(mokab_ctx *ctx, const mokab_mesh *m, const double *vec_edge, double *div_cell)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && m && vec_edge && div_cell, "divergence_on_cell: NULL argument");
        ctx->bind();
        OpBufs B;
        op_common(ctx, m, vec_edge, m->nE, m->dPermE.p, div_cell, m->nC, m->dPermC.p, false, B);
        LAUNCH(ctx, ref::k_divergence_on_cell, nblk(m->nC), 256, (int)m->nC, m->eoc.p, m->sgnC.p, m->nEoC.p, m->area.p, m->dv.p,
               (const double *)B.in_p.p, B.out_p.p);
        op_finish(ctx, div_cell, m->nC, m->dPermC.p, B);
    });
}

int mokab_curl_on_vertex(mokab_ctx *ctx, const mokab_mesh *m, const double *vec_edge, double *curl_vertex_inout)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && m && vec_edge && curl_vertex_inout, "curl_on_vertex: NULL argument");
        MOKAB_REQUIRE(m->nV > 0, "curl_on_vertex: the mesh was created without vertex arrays");
        ctx->bind();
        OpBufs B;
        op_common(ctx, m, vec_edge, m->nE, m->dPermE.p, curl_vertex_inout, m->nV, m->dPermV.p, true, B);
        LAUNCH(ctx, ref::k_curl_on_vertex, nblk(m->nV), 256, (int)m->nV, m->D, m->eov.p, m->sgnV.p, m->areaTri.p, m->dc.p,
               (const double *)B.in_p.p, B.out_p.p);
        op_finish(ctx, curl_vertex_inout, m->nV, m->dPermV.p, B);
    });
}

int mokab_interpolate_cell2edge(mokab_ctx *ctx, const mokab_mesh *m, const double *cell_value, double *edge_value)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && m && cell_value && edge_value, "interpolate_cell2edge: NULL argument");
        ctx->bind();
        OpBufs B;
        op_common(ctx, m, cell_value, m->nC, m->dPermC.p, edge_value, m->nE, m->dPermE.p, false, B);
        LAUNCH(ctx, ref::k_interpolate_cell2edge, nblk(m->nE), 256, (int)m->nE, m->ce.p, (const double *)B.in_p.p, B.out_p.p);
        op_finish(ctx, edge_value, m->nE, m->dPermE.p, B);
    });
}

// reverse mode of the two operators test/enzyme/test_Enzyme_Operators.jl differentiates
int mokab_gradient_on_edge_vjp(mokab_ctx *ctx, const mokab_mesh *m, const double *d_grad_edge, double *d_scalar_cell)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && m && d_grad_edge && d_scalar_cell, "gradient_on_edge_vjp: NULL argument");
        ctx->bind();
        OpBufs B;
        op_common(ctx, m, d_grad_edge, m->nE, m->dPermE.p, d_scalar_cell, m->nC, m->dPermC.p, false, B);
        LAUNCH(ctx, adjoint::k_gradient_on_edge_vjp, nblk(m->nC), 256, (int)m->nC, m->eoc.p, m->sgnC.p, m->nEoC.p, m->ce.p, m->dc.p,
               (const double *)B.in_p.p, B.out_p.p);
        op_finish(ctx, d_scalar_cell, m->nC, m->dPermC.p, B);
    });
}

int mokab_divergence_on_cell_vjp(mokab_ctx *ctx, const mokab_mesh *m, const double *d_div_cell, double *d_vec_edge)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && m && d_div_cell && d_vec_edge, "divergence_on_cell_vjp: NULL argument");
        ctx->bind();
        OpBufs B;
        op_common(ctx, m, d_div_cell, m->nC, m->dPermC.p, d_vec_edge, m->nE, m->dPermE.p, false, B);
        LAUNCH(ctx, adjoint::k_divergence_on_cell_vjp, nblk(m->nE), 256, (int)m->nE, m->ce.p, m->dv.p, m->area.p,
               (const double *)B.in_p.p, B.out_p.p);
        op_finish(ctx, d_vec_edge, m->nE, m->dPermE.p, B);
    });
}

// ---- src/forward entry points ----------------------------------------------------------------------------------------
int mokab_timestep_forward_euler(mokab_state *state, double dt, int64_t nsteps)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "timestep_forward_euler: state is NULL");
        MOKAB_REQUIRE(nsteps >= 0, "timestep_forward_euler: nsteps must be >= 0");
        require_f64(state, "timestep_forward_euler");
        MOKAB_REQUIRE(state->mesh->nCo == state->mesh->nC && state->mesh->nEo == state->mesh->nE,
                      "timestep_forward_euler: this mesh has halo entities; domain-decomposed runs step with "
                      "mokab_timestep_forward_euler_decomposed (or mokab_forward_euler_stage + the halo exchange)");
        state->ctx->bind();
        MOKAB_REQUIRE(!state->d->taping || state->K == 1, "timestep_forward_euler: the multi-level reverse mode is RungeKutta4");
        if (state->d->taping) {             // also a call with nsteps = 0: ssh is an input of its own for this stepper's seed
            MOKAB_REQUIRE(state->d->tapeKind != 1, "timestep_forward_euler: the tape already holds RungeKutta4 steps");
            // checked before the first step runs: a call either records all of its steps or leaves the state untouched
            MOKAB_REQUIRE((int64_t)state->d->tapeDt.size() + nsteps <= state->d->tapeCap,
                          "timestep_forward_euler: the tape is full (mokab_tape_begin max_steps)");
            state->d->tapeKind = 2;
        }
        if (fe_fusable(state)) {
            run_fe_fused(state, dt, nsteps);
        } else {
            for (int64_t i = 0; i < nsteps; ++i) {
                if (state->d->taping) fe_tape_record(state, dt, state->d->u[state->cur].p, state->d->hEdge.p);
                step_forward_euler(state, dt);
            }
        }
    });
}

int mokab_timestep_forward_euler_unfused(mokab_state *state, double dt, int64_t nsteps)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "timestep_forward_euler_unfused: state is NULL");
        MOKAB_REQUIRE(nsteps >= 0, "timestep_forward_euler_unfused: nsteps must be >= 0");
        require_f64(state, "timestep_forward_euler_unfused");
        MOKAB_REQUIRE(state->mesh->nCo == state->mesh->nC && state->mesh->nEo == state->mesh->nE,
                      "timestep_forward_euler_unfused: this mesh has halo entities; use mokab_timestep_forward_euler_decomposed");
        state->ctx->bind();
        fe_materialize(state);
        for (int64_t i = 0; i < nsteps; ++i) step_forward_euler(state, dt);
    });
}

int mokab_timestep_rk4(mokab_state *state, double dt, int64_t nsteps, int impl)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "timestep_rk4: state is NULL");
        MOKAB_REQUIRE(nsteps >= 0, "timestep_rk4: nsteps must be >= 0");
        MOKAB_REQUIRE(impl == MOKAB_RK4_FUSED || impl == MOKAB_RK4_UNFUSED, "timestep_rk4: unknown impl");
        state->ctx->bind();
        fe_materialize(state);
        if (impl == MOKAB_RK4_UNFUSED) {
            require_f64(state, "timestep_rk4(MOKAB_RK4_UNFUSED)");
            for (int64_t i = 0; i < nsteps; ++i) step_rk4_unfused(state, dt);
        } else if (state->dtype == MOKAB_F64) {
            run_rk4_fused<double>(state, dt, nsteps);
        } else {
            run_rk4_fused<float>(state, dt, nsteps);
        }
    });
}

int mokab_reduce(mokab_state *state, int which, double *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && out, "reduce: NULL argument");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) do_reduce<double>(state, which, out); else do_reduce<float>(state, which, out);
    });
}

// ---- reverse mode ------------------------------------------------------------------------------------------------
int mokab_tape_begin(mokab_state *state, int64_t max_steps)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "tape_begin: state is NULL");
        MOKAB_REQUIRE(max_steps >= 0, "tape_begin: max_steps must be >= 0");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) tape_begin<double>(state, max_steps); else tape_begin<float>(state, max_steps);
    });
}

int mokab_tape_length(mokab_state *state, int64_t *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && out, "tape_length: NULL argument");
        *out = state->dtype == MOKAB_F64 ? (int64_t)state->d->tapeDt.size() : (int64_t)state->f->tapeDt.size();
    });
}

int mokab_adjoint_seed(mokab_state *state, int which)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "adjoint_seed: state is NULL");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) adjoint_seed<double>(state, which); else adjoint_seed<float>(state, which);
    });
}

int mokab_adjoint_rk4(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "adjoint_rk4: state is NULL");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) adjoint_run<double>(state); else adjoint_run<float>(state);
    });
}

int mokab_adjoint_forward_euler(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "adjoint_forward_euler: state is NULL");
        state->ctx->bind();
        adjoint_run_fe(state);
    });
}

// ---- staged RK4 + halo messages (domain-decomposed runs) ----------------------------------------------------------
int mokab_halo_setup(mokab_mesh *m, int64_t n_send, const int32_t *send_idx, int64_t n_recv, const int32_t *recv_idx)
{
    return guarded([&] {
        MOKAB_REQUIRE(m && (n_send == 0 || send_idx) && (n_recv == 0 || recv_idx) && n_send >= 0 && n_recv >= 0,
                      "halo_setup: bad argument");
        m->ctx->bind();
        std::vector<int32_t> invC(m->nC), invE(m->nE);
        for (int64_t i = 0; i < m->nC; ++i) invC[m->permC[i]] = (int32_t)i;
        for (int64_t i = 0; i < m->nE; ++i) invE[m->permE[i]] = (int32_t)i;
        auto conv = [&](int64_t n, const int32_t *idx, bool send, std::vector<int32_t> &out) {
            out.resize(n);
            for (int64_t k = 0; k < n; ++k) {
                const int64_t i = idx[k];
                MOKAB_REQUIRE(i >= 0 && i < m->nC + m->nE, "halo_setup: index out of range");
                const bool cell = i < m->nC;
                const int32_t dev = cell ? invC[i] : invE[i - m->nC];
                const bool owned = cell ? dev < m->nCo : dev < m->nEo;
                MOKAB_REQUIRE(owned == send, send ? "halo_setup: send list names a halo entity" : "halo_setup: recv list names an owned entity");
                out[k] = cell ? dev : (int32_t)(m->nC + dev);
            }
        };
        std::vector<int32_t> s, r;
        conv(n_send, send_idx, true, s);
        conv(n_recv, recv_idx, false, r);
        // blocks holding a send entity must run in the boundary part
        std::vector<char> isB(m->fusedBlocks, 0);
        for (int32_t b : m->hBlkBoundary) isB[b] = 1;
        for (int32_t i : s) {
            int b;
            if (i < m->nC) b = i / kBlockCells;
            else b = (int)(std::upper_bound(m->hBlkEdgeStart.begin(), m->hBlkEdgeStart.end(), (int32_t)(i - m->nC)) - m->hBlkEdgeStart.begin()) - 1;
            if (b >= 0 && b < m->fusedBlocks) isB[b] = 1;
        }
        m->hBlkInterior.clear(); m->hBlkBoundary.clear();
        for (int b = 0; b < m->fusedBlocks; ++b) (isB[b] ? m->hBlkBoundary : m->hBlkInterior).push_back(b);
        cudaStream_t st = m->ctx->stream;
        m->blkInterior.upload(m->hBlkInterior, st); m->blkBoundary.upload(m->hBlkBoundary, st);
        m->nInterior = (int)m->hBlkInterior.size(); m->nBoundary = (int)m->hBlkBoundary.size();
        m->haloSend.upload(s, st);
        m->hHaloSend = s;
        m->haloRecv.upload(r, st);
        MOKAB_CUDA(cudaStreamSynchronize(st));
        m->halo_ready = true;
    });
}

int mokab_halo_pack(mokab_state *state, int stage, void *send_buf_device, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->mesh->halo_ready, "halo_pack: call mokab_halo_setup first");
        MOKAB_REQUIRE(stage >= 0 && stage <= 5, "halo_pack: stage must be 0..5");
        MOKAB_REQUIRE(state->K == 1, "halo_pack: single-level states only (nVertLevels == 1)");
        MOKAB_REQUIRE(send_buf_device || state->mesh->haloSend.n == 0, "halo_pack: NULL buffer");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) halo_pack<double>(state, stage, send_buf_device, (cudaStream_t)cuda_stream, false);
        else halo_pack<float>(state, stage, send_buf_device, (cudaStream_t)cuda_stream, false);
    });
}

int mokab_halo_unpack(mokab_state *state, int stage, const void *recv_buf_device, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->mesh->halo_ready, "halo_unpack: call mokab_halo_setup first");
        MOKAB_REQUIRE(stage >= 0 && stage <= 5, "halo_unpack: stage must be 0..5");
        MOKAB_REQUIRE(state->K == 1, "halo_unpack: single-level states only (nVertLevels == 1)");
        MOKAB_REQUIRE(recv_buf_device || state->mesh->haloRecv.n == 0, "halo_unpack: NULL buffer");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) halo_pack<double>(state, stage, const_cast<void *>(recv_buf_device), (cudaStream_t)cuda_stream, true);
        else halo_pack<float>(state, stage, const_cast<void *>(recv_buf_device), (cudaStream_t)cuda_stream, true);
    });
}

int mokab_rk4_stage(mokab_state *state, double dt, int stage, int part, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "rk4_stage: state is NULL");
        MOKAB_REQUIRE(stage >= 1 && stage <= 4, "rk4_stage: stage must be 1..4");
        MOKAB_REQUIRE(state->K == 1, "rk4_stage: single-level states only (nVertLevels == 1)");
        MOKAB_REQUIRE(part >= MOKAB_PART_ALL && part <= MOKAB_PART_ALL_PUSH, "rk4_stage: unknown part");
        state->ctx->bind();
        leave_forward_euler(state);
        if (state->dtype == MOKAB_F64) run_stage<double>(state, dt, stage, part, (cudaStream_t)cuda_stream);
        else run_stage<float>(state, dt, stage, part, (cudaStream_t)cuda_stream);
    });
}

int mokab_forward_euler_stage(mokab_state *state, double dt, int part, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "forward_euler_stage: state is NULL");
        MOKAB_REQUIRE(part >= MOKAB_PART_ALL && part <= MOKAB_PART_BOUNDARY, "forward_euler_stage: part must be ALL, INTERIOR or BOUNDARY");
        MOKAB_REQUIRE(state->K == 1, "forward_euler_stage: single-level states only (nVertLevels == 1)");
        require_f64(state, "forward_euler_stage");
        state->ctx->bind();
        run_fe_stage(state, dt, part, (cudaStream_t)cuda_stream);
    });
}

int mokab_forward_euler_finish_step(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "forward_euler_finish_step: state is NULL");
        require_f64(state, "forward_euler_finish_step");
        MOKAB_REQUIRE(state->fe_lazy, "forward_euler_finish_step: no staged ForwardEuler step has run");
        state->cur = 1 - state->cur;
    });
}

int mokab_rk4_finish_step(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "rk4_finish_step: state is NULL");
        state->cur = 1 - state->cur;
    });
}

int mokab_refresh_ssh(mokab_state *state, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "refresh_ssh: state is NULL");
        state->ctx->bind();
        leave_forward_euler(state);
        if (state->dtype == MOKAB_F64) refresh_ssh<double>(state, (cudaStream_t)cuda_stream);
        else refresh_ssh<float>(state, (cudaStream_t)cuda_stream);
    });
}

// ---- halo exchange by direct peer stores (kernels_p2p.cuh) -------------------------------------------------------------
int mokab_halo_recv_device_indices(const mokab_mesh *mesh, int32_t *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(mesh && mesh->halo_ready && (out || mesh->haloRecv.n == 0), "halo_recv_device_indices: call mokab_halo_setup first");
        mesh->ctx->bind();
        std::vector<int32_t> r(mesh->haloRecv.n);
        if (!r.empty()) {
            MOKAB_CUDA(cudaStreamSynchronize(mesh->ctx->stream));
            MOKAB_CUDA(cudaMemcpy(r.data(), mesh->haloRecv.p, r.size() * 4, cudaMemcpyDeviceToHost));
        }
        for (size_t k = 0; k < r.size(); ++k) out[k] = r[k] < mesh->nC ? r[k] : -(int32_t)(r[k] - mesh->nC) - 1;
    });
}

int mokab_p2p_blob_size(int64_t *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(out, "p2p_blob_size: NULL argument");
        *out = (int64_t)sizeof(P2PBlob);
    });
}

int mokab_p2p_export(mokab_state *state, int rank, void *blob)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && blob, "p2p_export: NULL argument");
        MOKAB_REQUIRE(state->K == 1, "p2p_export: single-level states only (nVertLevels == 1)");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) p2p_export<double>(state, (P2PBlob *)blob); else p2p_export<float>(state, (P2PBlob *)blob);
        ((P2PBlob *)blob)->rank = rank;
    });
}

int mokab_p2p_setup(mokab_state *state, int rank, int nranks, const void *blobs, int n_receivers, const int32_t *receiver_ranks,
                    const int64_t *push_counts, const int32_t *dst_idx, int n_senders, const int32_t *sender_ranks)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && blobs && state->mesh->halo_ready, "p2p_setup: NULL argument or mokab_halo_setup not called");
        MOKAB_REQUIRE((n_receivers == 0 || (receiver_ranks && push_counts)) && (n_senders == 0 || sender_ranks), "p2p_setup: NULL list");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64)
            p2p_setup<double>(state, rank, nranks, (const P2PBlob *)blobs, n_receivers, receiver_ranks, push_counts, dst_idx, n_senders, sender_ranks);
        else
            p2p_setup<float>(state, rank, nranks, (const P2PBlob *)blobs, n_receivers, receiver_ranks, push_counts, dst_idx, n_senders, sender_ranks);
    });
}

int mokab_halo_push(mokab_state *state, int stage, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->p2p.ready, "halo_push: call mokab_p2p_setup first");
        MOKAB_REQUIRE(stage >= 0 && stage <= 4, "halo_push: stage must be 0..4");
        state->ctx->bind();
        if (state->dtype == MOKAB_F64) p2p_push<double>(state, stage, (cudaStream_t)cuda_stream);
        else p2p_push<float>(state, stage, (cudaStream_t)cuda_stream);
    });
}

int mokab_halo_wait(mokab_state *state, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->p2p.ready, "halo_wait: call mokab_p2p_setup first");
        state->ctx->bind();
        p2p_wait(state, (cudaStream_t)cuda_stream);
    });
}

int mokab_halo_wait_arrivals(mokab_state *state, void *cuda_stream)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && state->p2p.ready, "halo_wait_arrivals: call mokab_p2p_setup first");
        state->ctx->bind();
        p2p_wait_arrivals(state, (cudaStream_t)cuda_stream);
    });
}

int mokab_p2p_close(mokab_state *state)
{
    return guarded([&] {
        MOKAB_REQUIRE(state, "p2p_close: state is NULL");
        state->ctx->bind();
        MOKAB_CUDA(cudaStreamSynchronize(state->ctx->stream));
        for (void *q : state->p2p.opened) MOKAB_CUDA(cudaIpcCloseMemHandle(q));
        state->p2p.opened.clear();
        state->p2p.ready = false;
    });
}

int mokab_p2p_error(mokab_state *state, int *out)
{
    return guarded([&] {
        MOKAB_REQUIRE(state && out, "p2p_error: NULL argument");
        *out = 0;
        if (!state->p2p.exported) return;
        state->ctx->bind();
        MOKAB_CUDA(cudaStreamSynchronize(state->ctx->stream));
        MOKAB_CUDA(cudaMemcpy(out, state->p2p.error.p, sizeof(int), cudaMemcpyDeviceToHost));
    });
}

// ---- stage timeline of the TRACE build (common.cuh) --------------------------------------------------------------------------
#ifdef MOKAB_TRACE
static TraceRec *g_trace_host_buf = nullptr;
static unsigned int g_trace_host_cap = 0;
#endif
int mokab_trace_begin(mokab_ctx *ctx, int64_t capacity)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && capacity >= 0, "trace_begin: bad argument");
#ifdef MOKAB_TRACE
        ctx->bind();
        MOKAB_CUDA(cudaDeviceSynchronize());
        TraceRec *none = nullptr;
        MOKAB_CUDA(cudaMemcpyToSymbol(g_trace_buf, &none, sizeof(none)));
        if (g_trace_host_buf) { cudaFree(g_trace_host_buf); g_trace_host_buf = nullptr; g_trace_host_cap = 0; }
        if (capacity == 0) return;
        MOKAB_CUDA(cudaMalloc(&g_trace_host_buf, (size_t)capacity * sizeof(TraceRec)));
        MOKAB_CUDA(cudaMemset(g_trace_host_buf, 0, (size_t)capacity * sizeof(TraceRec)));
        g_trace_host_cap = (unsigned int)capacity;
        const unsigned long long zero = 0;
        MOKAB_CUDA(cudaMemcpyToSymbol(g_trace_count, &zero, sizeof(zero)));
        MOKAB_CUDA(cudaMemcpyToSymbol(g_trace_cap, &g_trace_host_cap, sizeof(g_trace_host_cap)));
        MOKAB_CUDA(cudaMemcpyToSymbol(g_trace_buf, &g_trace_host_buf, sizeof(g_trace_host_buf)));
#else
        throw Error("trace_begin: this build carries no stage timeline (make libmoka_b200_trace.so, MOKAB_LIB=libmoka_b200_trace.so)");
#endif
    });
}

int mokab_trace_read(mokab_ctx *ctx, void *records, int64_t max_records, int64_t *count)
{
    return guarded([&] {
        MOKAB_REQUIRE(ctx && count && (records || max_records == 0), "trace_read: bad argument");
#ifdef MOKAB_TRACE
        ctx->bind();
        MOKAB_CUDA(cudaDeviceSynchronize());
        unsigned long long n = 0;
        MOKAB_CUDA(cudaMemcpyFromSymbol(&n, g_trace_count, sizeof(n)));
        const int64_t have = (int64_t)std::min<unsigned long long>(n, g_trace_host_cap);
        *count = (int64_t)n;
        const int64_t take = std::min(have, max_records);
        if (take > 0) MOKAB_CUDA(cudaMemcpy(records, g_trace_host_buf, (size_t)take * sizeof(TraceRec), cudaMemcpyDeviceToHost));
#else
        (void)records; (void)max_records;
        *count = 0;
        throw Error("trace_read: this build carries no stage timeline (make libmoka_b200_trace.so)");
#endif
    });
}

int mokab_set_option(const char *name, int64_t value)
{
    return guarded([&] {
        MOKAB_REQUIRE(name, "set_option: name is NULL");
        Options &o = options();
        const std::string n(name);
        if (n == "stage_tma") { MOKAB_REQUIRE(value >= 0 && value <= 3, "set_option: stage_tma must be 0..3"); o.stage_tma = (int)value; }
        else if (n == "stage_prefetch") { MOKAB_REQUIRE(value >= 0 && value <= 3, "set_option: stage_prefetch must be 0..3"); o.stage_prefetch = (int)value; }
        else if (n == "stage_prefetch_distance") { MOKAB_REQUIRE(value >= 0 && value < (1 << 30), "set_option: bad stage_prefetch_distance"); o.stage_prefetch_distance = (int)value; }
        else if (n == "stage_wf_block_major") o.stage_wf_block_major = value ? 1 : 0;
        else if (n == "stage_auto") o.stage_auto = value ? 1 : 0;
        else if (n == "stage_auto_hi") { MOKAB_REQUIRE(value >= 0 && value < (1 << 20), "set_option: bad stage_auto_hi"); o.stage_auto_hi = (int)value; }
        else if (n == "stage_pdl") o.stage_pdl = value ? 1 : 0;
        else if (n == "launch_priority") o.launch_priority = value ? 1 : 0;
        else if (n == "decomp_serial_blocks") o.decomp_serial_blocks = (int)value;
        else if (n == "test_drop_dependency") o.test_drop_dependency = (int)value;
        else throw Error("set_option: unknown option '" + n + "'");
        o.epoch++;
    });
}

int mokab_get_option(const char *name, int64_t *value)
{
    return guarded([&] {
        MOKAB_REQUIRE(name && value, "get_option: NULL argument");
        const Options &o = options();
        const std::string n(name);
        if (n == "stage_tma") *value = o.stage_tma;
        else if (n == "stage_prefetch") *value = o.stage_prefetch;
        else if (n == "stage_prefetch_distance") *value = o.stage_prefetch_distance;
        else if (n == "stage_wf_block_major") *value = o.stage_wf_block_major;
        else if (n == "stage_auto") *value = o.stage_auto;
        else if (n == "stage_auto_hi") *value = o.stage_auto_hi;
        else if (n == "stage_pdl") *value = o.stage_pdl;
        else if (n == "launch_priority") *value = o.launch_priority;
        else if (n == "decomp_serial_blocks") *value = o.decomp_serial_blocks;
        else throw Error("get_option: unknown option '" + n + "'");
    });
}

int mokab_mesh_derived_blocks(const mokab_mesh *mesh, int64_t *blocks, int64_t *derived)
{
    return guarded([&] {
        MOKAB_REQUIRE(mesh && blocks && derived, "mesh_derived_blocks: NULL argument");
        *blocks = mesh->fusedBlocks;
        *derived = mesh->nDerivedBlocks;
    });
}

int mokab_mesh_block_counts(const mokab_mesh *mesh, int64_t *interior, int64_t *boundary)
{
    return guarded([&] {
        MOKAB_REQUIRE(mesh && interior && boundary, "mesh_block_counts: NULL argument");
        *interior = mesh->nInterior;
        *boundary = mesh->nBoundary;
    });
}

}  // extern "C"

#include "decomposed.cuh"
