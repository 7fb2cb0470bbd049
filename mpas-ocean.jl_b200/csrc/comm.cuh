// comm.cuh -- the communicator of domain-decomposed runs: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no multi-device path; BASELINE.json's north_star asks for "halo exchange as NCCL
// send/recv (or direct P2P stores) over NVLink, overlapped with interior-cell compute", behind the C ABI.  Everything the
// library needs from a communication layer is ONE primitive -- a variable-count all-to-all of device buffers in stream order
// (ncclGroupStart; ncclSend / ncclRecv per peer; ncclGroupEnd) -- which carries the halo messages of every RK stage, the
// one-off address / index exchange of the direct-store halo path, and (with one element per rank, summed on the host in rank
// order, hence deterministic) the scalar reductions.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): single-GPU users need no NCCL at all, and inside a process that has
// already loaded a NCCL (PyTorch bundles one) the same copy is used.  The caller brings the ranks together exactly as NCCL
// itself asks: rank 0 obtains a unique id (mokab_comm_get_unique_id), hands it to the others by whatever means the host
// program has (MPI.jl, a file, a TCP store), every rank calls mokab_comm_init.
//
// The simulated runtime of tests/sim stands in for NCCL with its in-stream all-to-all between emulated ranks (host threads).
#pragma once
#include "common.cuh"

#include <map>
#include <mutex>

#ifndef MOKAB_SIM
#include <dlfcn.h>
#include <nccl.h>
#else
extern "C" {
void *mokab_sim_comm_create(int nranks);
void mokab_sim_comm_destroy(void *c);
int mokab_sim_all_to_all(void *comm, int rank, cudaStream_t s, void *send, void *recv, const int64_t *scnt, const int64_t *rcnt,
                         int64_t elem_size);
}
#endif

struct mokab_comm {
    mokab_ctx *ctx = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;                 // for the host-level collectives (barrier, scalar reductions, set-up traffic)
    mokab::DevBuf<unsigned char> bufS, bufR;       // their staging
    uint64_t token = 0;
#ifdef MOKAB_SIM
    void *sim = nullptr;
#else
    ncclComm_t nccl = nullptr;
#endif
};

namespace mokab {
namespace comm {

constexpr int kIdBytes = 128;                      // sizeof(ncclUniqueId)

#ifndef MOKAB_SIM
struct Nccl {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    static Nccl &get()
    {
        static Nccl n = [] {
            Nccl x;
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!h) throw Error(std::string("mokab_comm: cannot load libnccl.so.2 (") + dlerror() + "); domain-decomposed runs need NCCL");
            auto sym = [&](const char *name) {
                void *p = dlsym(h, name);
                if (!p) throw Error(std::string("mokab_comm: libnccl.so.2 has no symbol ") + name);
                return p;
            };
            x.GetUniqueId = (decltype(x.GetUniqueId))sym("ncclGetUniqueId");
            x.CommInitRank = (decltype(x.CommInitRank))sym("ncclCommInitRank");
            x.CommDestroy = (decltype(x.CommDestroy))sym("ncclCommDestroy");
            x.GroupStart = (decltype(x.GroupStart))sym("ncclGroupStart");
            x.GroupEnd = (decltype(x.GroupEnd))sym("ncclGroupEnd");
            x.Send = (decltype(x.Send))sym("ncclSend");
            x.Recv = (decltype(x.Recv))sym("ncclRecv");
            x.GetErrorString = (decltype(x.GetErrorString))sym("ncclGetErrorString");
            x.GetVersion = (decltype(x.GetVersion))sym("ncclGetVersion");
            return x;
        }();
        return n;
    }
};
#define MOKAB_NCCL(expr)                                                                                       \
    do {                                                                                                       \
        ncclResult_t r__ = (expr);                                                                             \
        if (r__ != ncclSuccess)                                                                                \
            throw ::mokab::Error(std::string(#expr) + " failed: " + ::mokab::comm::Nccl::get().GetErrorString(r__)); \
    } while (0)
#else
// emulated ranks are threads of one process: the "unique id" is a token, the first rank to arrive creates the shared object
struct SimRegistry {
    std::mutex mu;
    struct Entry { void *comm; int refs; };
    std::map<uint64_t, Entry> live;
    uint64_t next = 1;
    static SimRegistry &get() { static SimRegistry r; return r; }
};
#endif

// Variable-count all-to-all of device buffers on stream `s`: segment q of `send` (scnt[q] elements of `elem` bytes) goes to rank
// q, segment q of `recv` (rcnt[q] elements) comes from it.  Capturable into a CUDA graph.
static void all_to_all(mokab_comm *c, cudaStream_t s, const void *send, void *recv, const int64_t *scnt, const int64_t *rcnt, size_t elem)
{
#ifdef MOKAB_SIM
    if (mokab_sim_all_to_all(c->sim, c->rank, s, const_cast<void *>(send), recv, scnt, rcnt, (int64_t)elem) != 0)
        throw Error("mokab_comm: the simulated all-to-all failed");
#else
    Nccl &n = Nccl::get();
    MOKAB_NCCL(n.GroupStart());
    size_t so = 0, ro = 0;
    for (int q = 0; q < c->nranks; ++q) {
        if (rcnt[q] > 0) MOKAB_NCCL(n.Recv((unsigned char *)recv + ro, (size_t)rcnt[q] * elem, ncclChar, q, c->nccl, s));
        if (scnt[q] > 0) MOKAB_NCCL(n.Send((const unsigned char *)send + so, (size_t)scnt[q] * elem, ncclChar, q, c->nccl, s));
        so += (size_t)scnt[q] * elem;
        ro += (size_t)rcnt[q] * elem;
    }
    MOKAB_NCCL(n.GroupEnd());
#endif
}

// Host-level exchange (set-up traffic, reductions): every rank contributes `bytes` bytes per destination (send: nranks segments),
// receives the segments addressed to it (recv: nranks segments), synchronously.
static void exchange_host(mokab_comm *c, const void *send, void *recv, size_t bytes)
{
    const size_t total = bytes * (size_t)c->nranks;
    if (c->bufS.n < total) c->bufS.alloc(total);     // (each on its own: exchange_host_v sizes the two differently)
    if (c->bufR.n < total) c->bufR.alloc(total);
    std::vector<int64_t> cnt((size_t)c->nranks, (int64_t)bytes);
    MOKAB_CUDA(cudaMemcpyAsync(c->bufS.p, send, total, cudaMemcpyHostToDevice, c->stream));
    all_to_all(c, c->stream, c->bufS.p, c->bufR.p, cnt.data(), cnt.data(), 1);
    MOKAB_CUDA(cudaMemcpyAsync(recv, c->bufR.p, total, cudaMemcpyDeviceToHost, c->stream));
    MOKAB_CUDA(cudaStreamSynchronize(c->stream));
}

// the same with per-destination counts (bytes): send segment q has sbytes[q] bytes, recv segment q has rbytes[q]
static void exchange_host_v(mokab_comm *c, const void *send, void *recv, const std::vector<int64_t> &sbytes, const std::vector<int64_t> &rbytes)
{
    size_t ts = 0, tr = 0;
    for (int64_t b : sbytes) ts += (size_t)b;
    for (int64_t b : rbytes) tr += (size_t)b;
    if (c->bufS.n < std::max<size_t>(ts, 1)) c->bufS.alloc(std::max<size_t>(ts, 1));
    if (c->bufR.n < std::max<size_t>(tr, 1)) c->bufR.alloc(std::max<size_t>(tr, 1));
    if (ts) MOKAB_CUDA(cudaMemcpyAsync(c->bufS.p, send, ts, cudaMemcpyHostToDevice, c->stream));
    all_to_all(c, c->stream, c->bufS.p, c->bufR.p, sbytes.data(), rbytes.data(), 1);
    if (tr) MOKAB_CUDA(cudaMemcpyAsync(recv, c->bufR.p, tr, cudaMemcpyDeviceToHost, c->stream));
    MOKAB_CUDA(cudaStreamSynchronize(c->stream));
}

// every rank's `n` doubles on every rank (rank-major), then reduced on the host in rank order: the same bits on every rank
static void allreduce_f64(mokab_comm *c, double *inout, int64_t n, int op)
{
    MOKAB_REQUIRE(op >= 0 && op <= 2, "comm_allreduce_f64: op must be 0 (sum), 1 (max) or 2 (min)");
    if (n <= 0) return;
    std::vector<double> send((size_t)n * c->nranks), recv((size_t)n * c->nranks);
    for (int q = 0; q < c->nranks; ++q) memcpy(send.data() + (size_t)q * n, inout, (size_t)n * sizeof(double));
    exchange_host(c, send.data(), recv.data(), (size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; ++i) {
        double v = recv[i];
        for (int q = 1; q < c->nranks; ++q) {
            const double x = recv[(size_t)q * n + i];
            v = op == 0 ? v + x : op == 1 ? std::max(v, x) : std::min(v, x);
        }
        inout[i] = v;
    }
}

}  // namespace comm
}  // namespace mokab
