// sim_runtime.cpp -- CPU simulation of the CUDA runtime subset libmoka_b200 uses (see include/cuda_runtime.h).
// TEST INFRASTRUCTURE: linked only into tests/sim/_build/libmoka_b200_sim.so, never into the product.
//
// Model
//   * device memory = host memory, filled with 0xFF on allocation and on free (doubles read as NaN, indices as -1);
//   * a stream is a FIFO of operations that execute LATER: when a host call has to wait for them, or when the
//     scheduling policy decides to.  Nothing orders two streams except events (and joins through them), exactly as
//     on hardware; the policy picks adversarial interleavings among the ones the dependencies allow:
//        FIFO          every operation runs inside the call that enqueues it, if its dependencies allow (the order a
//                      fully synchronous device would produce);
//        LAZY          nothing runs until a host call waits for it; only the awaited stream and what it transitively
//                      waits on make progress (a producer that the consumer forgot to wait for has NOT run);
//        OTHERS_FIRST  before the awaited stream moves, every other stream runs as far as it can, newest stream
//                      first (a writer the reader forgot to hold back HAS already overwritten the data);
//        RANDOM        seeded random interleaving, at enqueue time and while waiting;
//   * stream capture records the operations and ONLY the dependencies capture semantics give them (stream order
//     within each captured stream, fork/join through events); a replay executes the nodes in a policy-chosen
//     topological order of that DAG;
//   * host-side semantics of pageable/pinned copies follow the CUDA documentation: H2D from pageable memory is
//     staged before the call returns, D2H to pageable memory returns after the copy has completed, pinned memory
//     (cudaHostAlloc) is read / written when the stream gets there;
//   * an in-stream all-to-all between emulated ranks (mokab_sim_all_to_all) stands in for the NCCL collective: it
//     completes on a rank's stream only when every rank's stream has reached the matching call;
//   * several host threads (one per emulated rank) may drive the runtime concurrently: one global lock, blocked
//     waits sleep on a condition variable until another thread enqueues what they wait for.
#include <cuda_runtime.h>

#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <string>
#include <vector>

namespace mokab_sim {
thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
}  // namespace mokab_sim

struct Comm;
struct GraphRun;

enum OpKind { OP_WORK, OP_RECORD, OP_WAIT, OP_COLL, OP_GRAPH, OP_TRY };

struct CollArgs {
    Comm *comm = nullptr;
    int rank = 0;
    uint64_t seq = 0;
    char *send = nullptr, *recv = nullptr;
    std::vector<int64_t> scnt, rcnt;
    int64_t elem = 0;
};

struct Op {
    OpKind kind = OP_WORK;
    std::function<void()> fn;
    std::function<bool()> try_fn;  // OP_TRY: retried until it returns true (a device-side wait on memory another stream writes)
    cudaEvent_t ev = nullptr;
    uint64_t ticket = 0;
    cudaStream_t src = nullptr;  // OP_WAIT: the stream the awaited record was enqueued on
    CollArgs coll;
    std::shared_ptr<GraphRun> run;
    std::string name;
};

struct mokab_sim_graph;
struct mokab_sim_stream {
    int id = 0, priority = 0;
    int device = 0;                  // the emulated rank (host thread) that created it: device-wide waits are per rank
    std::deque<Op> q;
    mokab_sim_graph *cap = nullptr;  // non-null while capturing
    std::vector<int> frontier;       // capture: the nodes the next captured operation depends on
};

struct mokab_sim_event {
    uint64_t recorded = 0, done = 0;  // ticket of the latest cudaEventRecord call / of the latest record executed
    cudaStream_t rec_stream = nullptr;
    mokab_sim_graph *cap = nullptr;   // recorded during this capture (then `frontier` is what a waiter depends on)
    std::vector<int> frontier;
    std::chrono::steady_clock::time_point when;
};

struct Node {
    OpKind kind = OP_WORK;
    std::function<void()> fn;
    std::function<bool()> try_fn;
    CollArgs coll;
    std::vector<int> deps;
    std::string name;
};

struct mokab_sim_graph {
    std::vector<Node> nodes;
    cudaStream_t origin = nullptr;
    std::vector<cudaStream_t> members;
    cudaStreamCaptureMode mode = cudaStreamCaptureModeGlobal;
    bool invalid = false;
};
struct mokab_sim_graph_exec {
    std::vector<Node> nodes;
};

struct GraphRun {
    const mokab_sim_graph_exec *g = nullptr;
    std::vector<char> done;
    std::vector<uint64_t> seq;  // per node: collective sequence number of this launch
    size_t ndone = 0;
};

struct Comm {
    int n = 0;
    std::vector<uint64_t> issued;      // per rank: collectives enqueued so far
    std::vector<uint64_t> arrived;     // per rank: sequence number its stream is waiting in
    std::vector<CollArgs> args;        // per rank: arguments of the call it is waiting in
    std::vector<cudaStream_t> stream;  // per rank: stream of its latest call
    uint64_t completed = 0;
};

namespace {
enum Policy { FIFO = 0, LAZY = 1, OTHERS_FIRST = 2, RANDOM = 3 };

struct Runtime {
    std::mutex mu;
    std::condition_variable cv;
    Policy policy = FIFO;
    std::mt19937_64 rng{1};
    std::vector<cudaStream_t> streams;
    std::map<void *, size_t> dev, pinned;
    std::set<mokab_sim_graph *> captures;
    uint64_t ticket = 0;
    int next_stream_id = 0;
    uint64_t enqueue_epoch = 0;  // bumped by every enqueue (lets blocked waiters notice new work)
    int64_t n_kernels = 0, n_ops = 0, n_graph_nodes = 0, n_colls = 0;
    std::string fatal;
    double deadlock_s = 60.0;   // a blocked wait gives other host threads (emulated ranks still setting up) this long to enqueue what it waits for
};
Runtime R;
thread_local cudaError_t t_last = cudaSuccess;
thread_local int t_capturing = 0;  // captures begun by this host thread and not yet ended
thread_local int t_device = 0;     // which emulated device this host thread drives (mokab_sim_set_thread_device)

cudaError_t fail(cudaError_t e)
{
    t_last = e;
    return e;
}

[[noreturn]] void die(const std::string &msg)
{
    fprintf(stderr, "mokab_sim: fatal: %s\n", msg.c_str());
    fflush(stderr);
    abort();
}

// ---- kernels: serial threads, fibers for the cooperative ones ----------------------------------------------------
// A cooperative block runs its threads as fibers on one pooled stack area.  Switching is a dozen instructions on
// x86-64 (callee-saved registers + stack pointer; ucontext's swapcontext costs two signal-mask system calls per switch,
// which at 1024 blocks x 256 threads x ~12 switches per reduction kernel is seconds) and ucontext elsewhere.
#if defined(__x86_64__)
extern "C" void mokab_sim_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl mokab_sim_switch
    .type mokab_sim_switch, @function
mokab_sim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size mokab_sim_switch, .-mokab_sim_switch
)");
#define MOKAB_SIM_ASM_FIBERS 1
#endif

struct CoopBlock {
    struct Warp { int live = 0, arrived = 0; uint64_t gen = 0; double buf[32]; };
    unsigned nthreads = 0;
#ifdef MOKAB_SIM_ASM_FIBERS
    std::vector<void *> sp;
    void *main_sp = nullptr;
#else
    std::vector<ucontext_t> ctx;
    ucontext_t main;
#endif
    std::vector<char> finished;
    int live = 0, arrived = 0;
    uint64_t gen = 0;
    std::vector<Warp> warps;
    unsigned cur = 0;
    const std::function<void()> *body = nullptr;
};
thread_local CoopBlock *t_coop = nullptr;
thread_local bool t_in_kernel = false;
thread_local std::vector<unsigned char> t_dyn_smem;

void coop_to_main(CoopBlock *b)
{
#ifdef MOKAB_SIM_ASM_FIBERS
    mokab_sim_switch(&b->sp[b->cur], b->main_sp);
#else
    swapcontext(&b->ctx[b->cur], &b->main);
#endif
}

void coop_yield() { coop_to_main(t_coop); }

void coop_trampoline()
{
    CoopBlock *b = t_coop;
    (*b->body)();
    const unsigned t = b->cur;
    b->finished[t] = 1;
    b->live--;
    CoopBlock::Warp &w = b->warps[t / 32];
    w.live--;
    if (b->live > 0 && b->arrived == b->live) { b->arrived = 0; b->gen++; }      // a barrier the leaver was holding up
    if (w.live > 0 && w.arrived == w.live) { w.arrived = 0; w.gen++; }
    coop_to_main(b);
    die("a finished fiber was resumed");
}

void run_block_coop(unsigned nthreads, const std::function<void()> &body)
{
    constexpr size_t kStack = 64 * 1024;
    static thread_local std::vector<char> stacks;   // one pool per host thread, reused by every block (never zeroed again)
    if (stacks.size() < (size_t)nthreads * kStack + 64) stacks.resize((size_t)nthreads * kStack + 64);
    CoopBlock b;
    b.nthreads = nthreads;
    b.finished.assign(nthreads, 0);
    b.live = (int)nthreads;
    b.warps.resize((nthreads + 31) / 32);
    for (unsigned t = 0; t < nthreads; ++t) b.warps[t / 32].live++;
    b.body = &body;
    t_coop = &b;
#ifdef MOKAB_SIM_ASM_FIBERS
    b.sp.resize(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        // initial frame: six zeroed callee-saved registers, then the entry point `ret` jumps to; 16-byte aligned so the
        // entry sees the ABI's (rsp + 8) % 16 == 0
        uintptr_t top = ((uintptr_t)stacks.data() + (size_t)(t + 1) * kStack) & ~(uintptr_t)15;
        void **f = (void **)(top - 64);
        for (int i = 0; i < 8; ++i) f[i] = nullptr;
        f[6] = (void *)coop_trampoline;
        b.sp[t] = f;
    }
#else
    b.ctx.resize(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        getcontext(&b.ctx[t]);
        b.ctx[t].uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
        b.ctx[t].uc_stack.ss_size = kStack;
        b.ctx[t].uc_link = &b.main;
        makecontext(&b.ctx[t], (void (*)())coop_trampoline, 0);
    }
#endif
    int guard = 0;
    while (b.live > 0) {
        bool any = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            if (b.finished[t]) continue;
            any = true;
            b.cur = t;
            mokab_sim::t_threadIdx = uint3{t, 0, 0};
#ifdef MOKAB_SIM_ASM_FIBERS
            mokab_sim_switch(&b.main_sp, b.sp[t]);
#else
            swapcontext(&b.main, &b.ctx[t]);
#endif
        }
        if (!any) break;
        if (++guard > 1000000) die("cooperative kernel does not terminate (barrier mismatch?)");
    }
    t_coop = nullptr;
}

void run_kernel(unsigned grid, unsigned block, size_t smem, bool coop, const std::function<void()> &body)
{
    using namespace mokab_sim;
    if (t_dyn_smem.size() < smem + 16) t_dyn_smem.resize(smem + 16);
    t_gridDim = dim3{grid, 1, 1};
    t_blockDim = dim3{block, 1, 1};
    t_in_kernel = true;
    for (unsigned b = 0; b < grid; ++b) {
        t_blockIdx = uint3{b, 0, 0};
        if (smem) memset(t_dyn_smem.data(), 0xFF, smem);                  // shared memory starts undefined in every block
        if (coop) {
            run_block_coop(block, body);
        } else {
            for (unsigned t = 0; t < block; ++t) {
                t_threadIdx = uint3{t, 0, 0};
                body();
            }
        }
    }
    t_in_kernel = false;
    R.n_kernels++;
}

// ---- scheduler ---------------------------------------------------------------------------------------------------
enum StepResult { STEP_DONE, STEP_EMPTY, STEP_BLOCKED };

void do_coll_copies(Comm *c)
{
    for (int r = 0; r < c->n; ++r) {
        const CollArgs &a = c->args[r];
        int64_t ro = 0;
        for (int q = 0; q < c->n; ++q) {
            const CollArgs &b = c->args[q];
            int64_t so = 0;
            for (int k = 0; k < r; ++k) so += b.scnt[k];
            if (a.rcnt[q] != b.scnt[r]) die("all_to_all: rank " + std::to_string(r) + " expects " + std::to_string(a.rcnt[q]) +
                                            " elements from rank " + std::to_string(q) + " which sends " + std::to_string(b.scnt[r]));
            if (a.rcnt[q]) memcpy(a.recv + ro * a.elem, b.send + so * b.elem, (size_t)(a.rcnt[q] * a.elem));
            ro += a.rcnt[q];
        }
    }
    R.n_colls++;
}

// one rank's stream has reached its collective `a`; true once the collective has completed for everybody
bool try_coll(const CollArgs &a, std::vector<cudaStream_t> *blockers)
{
    Comm *c = a.comm;
    if (c->completed >= a.seq) return true;
    if (a.seq != c->completed + 1) die("all_to_all: collectives reached out of order on rank " + std::to_string(a.rank));
    c->arrived[a.rank] = a.seq;
    c->args[a.rank] = a;
    bool all = true;
    for (int r = 0; r < c->n; ++r)
        if (c->arrived[r] != a.seq) {
            all = false;
            if (blockers && c->stream[r]) blockers->push_back(c->stream[r]);
        }
    if (!all) return false;
    do_coll_copies(c);
    c->completed = a.seq;
    return true;
}

StepResult step_graph(Op &op, std::vector<cudaStream_t> *blockers)
{
    GraphRun &run = *op.run;
    const std::vector<Node> &nodes = run.g->nodes;
    if (run.ndone == nodes.size()) return STEP_DONE;
    std::vector<int> ready;
    for (int i = 0; i < (int)nodes.size(); ++i) {
        if (run.done[i]) continue;
        bool ok = true;
        for (int d : nodes[i].deps) ok = ok && run.done[d];
        if (ok) ready.push_back(i);
    }
    if (ready.empty()) die("graph replay: no ready node (cyclic capture?)");
    if (R.policy == RANDOM) std::shuffle(ready.begin(), ready.end(), R.rng);
    else if (R.policy != FIFO) std::reverse(ready.begin(), ready.end());  // adversarial: latest-captured ready node first
    for (int i : ready) {
        const Node &n = nodes[i];
        if (n.kind == OP_COLL) {
            CollArgs a = n.coll;
            a.seq = run.seq[i];
            if (!try_coll(a, blockers)) continue;
        } else if (n.kind == OP_TRY) {
            if (!n.try_fn()) continue;
        } else {
            n.fn();
        }
        run.done[i] = 1;
        run.ndone++;
        R.n_graph_nodes++;
        return STEP_DONE;  // one node per step; the op stays at the head of its stream until all nodes are done
    }
    return STEP_BLOCKED;
}

StepResult try_step(cudaStream_t s, std::vector<cudaStream_t> *blockers)
{
    if (s->q.empty()) return STEP_EMPTY;
    Op &op = s->q.front();
    switch (op.kind) {
    case OP_WORK:
        op.fn();
        break;
    case OP_RECORD:
        op.ev->done = std::max(op.ev->done, op.ticket);
        op.ev->when = std::chrono::steady_clock::now();
        break;
    case OP_WAIT:
        if (op.ev->done < op.ticket) {
            if (blockers && op.src) blockers->push_back(op.src);
            return STEP_BLOCKED;
        }
        break;
    case OP_COLL:
        if (!try_coll(op.coll, blockers)) return STEP_BLOCKED;
        break;
    case OP_TRY:
        if (!op.try_fn()) return STEP_BLOCKED;
        break;
    case OP_GRAPH: {
        StepResult r = step_graph(op, blockers);
        if (r == STEP_BLOCKED) return r;
        if (op.run->ndone < op.run->g->nodes.size()) { R.n_ops++; return STEP_DONE; }
        break;
    }
    }
    s->q.pop_front();
    R.n_ops++;
    return STEP_DONE;
}

// one unit of progress on `s`, or on something it is blocked on (`seen`: the streams already tried in this attempt --
// each is tried once, so mutually blocked streams terminate and the search is linear)
bool progress_in(cudaStream_t s, std::vector<cudaStream_t> &seen)
{
    seen.push_back(s);
    std::vector<cudaStream_t> blockers;
    StepResult r = try_step(s, &blockers);
    if (r == STEP_DONE) return true;
    if (r == STEP_EMPTY) return false;
    if (blockers.empty())      // a wait on memory (OP_TRY): whoever writes it is unknown -- any other stream may be the one
        for (auto it = R.streams.rbegin(); it != R.streams.rend(); ++it)
            if (!(*it)->cap) blockers.push_back(*it);
    if (R.policy == RANDOM) std::shuffle(blockers.begin(), blockers.end(), R.rng);
    for (cudaStream_t b : blockers) {
        if (std::find(seen.begin(), seen.end(), b) != seen.end()) continue;
        if (progress_in(b, seen)) return true;
    }
    return false;
}

bool progress(cudaStream_t s, int)
{
    std::vector<cudaStream_t> seen;
    return progress_in(s, seen);
}

std::vector<cudaStream_t> others_order(cudaStream_t except)
{
    std::vector<cudaStream_t> v;
    for (auto it = R.streams.rbegin(); it != R.streams.rend(); ++it)
        if (*it != except && !(*it)->cap) v.push_back(*it);
    return v;
}

// run every stream but `except` as far as it can go
void drain_others(cudaStream_t except)
{
    bool moved = true;
    while (moved) {
        moved = false;
        for (cudaStream_t s : others_order(except))
            while (progress(s, 0)) moved = true;
    }
}

// work pending on the calling host thread's device (each emulated rank has its own: a device-wide wait of one rank must
// not wait for the others, which may legitimately be blocked in a collective until this rank gets there)
bool pending_anywhere()
{
    for (cudaStream_t s : R.streams)
        if (s->device == t_device && !s->q.empty()) return true;
    return false;
}

// Block the calling host thread until pred() holds, driving `target` (or everything when null).
template <class Pred>
cudaError_t wait_until(std::unique_lock<std::mutex> &lk, cudaStream_t target, Pred pred, const char *what)
{
    auto last_progress = std::chrono::steady_clock::now();
    while (!pred()) {
        bool moved = false;
        if (R.policy == OTHERS_FIRST) drain_others(target);
        if (R.policy == RANDOM) {
            std::vector<cudaStream_t> v = others_order(nullptr);
            if (!v.empty() && (R.rng() & 3) != 0) moved = progress(v[R.rng() % v.size()], 0);
        }
        if (pred()) break;
        if (target) {
            moved = progress(target, 0) || moved;
        } else {
            for (cudaStream_t s : others_order(nullptr))
                if (s->device == t_device) moved = progress(s, 0) || moved;
        }
        if (moved) {
            last_progress = std::chrono::steady_clock::now();
            continue;
        }
        // nothing can move: another host thread has to enqueue what we wait for
        const uint64_t epoch = R.enqueue_epoch;
        R.cv.wait_for(lk, std::chrono::milliseconds(20), [&] { return R.enqueue_epoch != epoch; });
        if (R.enqueue_epoch != epoch) { last_progress = std::chrono::steady_clock::now(); continue; }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - last_progress).count() > R.deadlock_s) {
            static const char *kinds[] = {"work", "record", "wait", "collective", "graph", "wait-on-memory"};
            for (cudaStream_t x : R.streams) {
                fprintf(stderr, "mokab_sim:   stream %d%s: %zu pending", x->id, x == target ? " (awaited)" : "", x->q.size());
                if (!x->q.empty()) {
                    const Op &o = x->q.front();
                    fprintf(stderr, ", head = %s %s", kinds[o.kind], o.name.c_str());
                    if (o.kind == OP_WAIT) fprintf(stderr, " ticket %llu (event done %llu) recorded on stream %d", (unsigned long long)o.ticket,
                                                   (unsigned long long)o.ev->done, o.src ? o.src->id : -1);
                    if (o.kind == OP_COLL) fprintf(stderr, " rank %d seq %llu (completed %llu)", o.coll.rank, (unsigned long long)o.coll.seq,
                                                   (unsigned long long)o.coll.comm->completed);
                }
                fprintf(stderr, "\n");
            }
            die(std::string("deadlock: ") + what + " waits for work nobody enqueues");
        }
    }
    return cudaSuccess;
}

void after_enqueue(cudaStream_t s)
{
    R.enqueue_epoch++;
    R.cv.notify_all();
    if (R.policy == FIFO) {
        // a synchronous device: everything that can run, runs now, in host order
        bool moved = true;
        while (moved) {
            moved = false;
            for (cudaStream_t x : R.streams)
                while (!x->cap && progress(x, 0)) moved = true;
        }
    } else if (R.policy == RANDOM) {
        int n = (int)(R.rng() % 4);
        std::vector<cudaStream_t> v = others_order(nullptr);
        for (int i = 0; i < n && !v.empty(); ++i) progress(v[R.rng() % v.size()], 0);
    }
    (void)s;
}

bool is_pinned(const void *p, size_t bytes)
{
    auto it = R.pinned.upper_bound(const_cast<void *>(p));
    if (it == R.pinned.begin()) return false;
    --it;
    return (const char *)p + bytes <= (const char *)it->first + it->second;
}

void enqueue(cudaStream_t s, Op &&op)
{
    if (s->cap) {
        mokab_sim_graph *g = s->cap;
        if (op.kind == OP_RECORD || op.kind == OP_WAIT || op.kind == OP_GRAPH) die("internal: record/wait/graph op reached a capturing stream");
        Node n;
        n.kind = op.kind; n.fn = std::move(op.fn); n.try_fn = std::move(op.try_fn); n.coll = op.coll; n.deps = s->frontier; n.name = op.name;
        g->nodes.push_back(std::move(n));
        s->frontier.assign(1, (int)g->nodes.size() - 1);
        return;
    }
    s->q.push_back(std::move(op));
    after_enqueue(s);
}

cudaStream_t null_stream()
{
    static mokab_sim_stream *ns = nullptr;
    if (!ns) {
        ns = new mokab_sim_stream();
        ns->device = -1;   // the legacy default stream is only reached through a null handle; nobody's device-wide wait covers it
        ns->id = R.next_stream_id++;
        R.streams.push_back(ns);
    }
    return ns;
}
cudaStream_t resolve(cudaStream_t s) { return s ? s : null_stream(); }

bool capture_blocks_this_thread()
{
    if (t_capturing > 0) return true;
    for (mokab_sim_graph *g : R.captures)
        if (g->mode == cudaStreamCaptureModeGlobal) return true;
    return false;
}
}  // namespace

// ---- device-side primitives --------------------------------------------------------------------------------------
namespace mokab_sim {
void sync_threads()
{
    CoopBlock *b = t_coop;
    if (!b) die("__syncthreads() in a kernel that was not launched cooperatively (add it to COOPERATIVE in tests/sim/build.py)");
    const uint64_t g = b->gen;
    if (++b->arrived == b->live) { b->arrived = 0; b->gen++; return; }
    while (b->gen == g) coop_yield();
}

static void warp_barrier(CoopBlock *b, CoopBlock::Warp &w)
{
    const uint64_t g = w.gen;
    if (++w.arrived == w.live) { w.arrived = 0; w.gen++; return; }
    while (w.gen == g) coop_yield();
    (void)b;
}

double shfl_down(double v, unsigned delta)
{
    CoopBlock *b = t_coop;
    if (!b) die("warp shuffle in a kernel that was not launched cooperatively (add it to COOPERATIVE in tests/sim/build.py)");
    const unsigned t = b->cur, lane = t % 32;
    CoopBlock::Warp &w = b->warps[t / 32];
    w.buf[lane] = v;
    warp_barrier(b, w);
    const unsigned src = lane + delta;
    const double r = (src < 32 && (t / 32) * 32 + src < b->nthreads) ? w.buf[src] : v;
    warp_barrier(b, w);
    return r;
}

void enqueue_try(cudaStream_t s, const char *name, std::function<bool()> ready)
{
    std::unique_lock<std::mutex> lk(R.mu);
    Op op;
    op.kind = OP_TRY;
    op.name = name;
    op.try_fn = std::move(ready);
    enqueue(resolve(s), std::move(op));
}

unsigned char *dynamic_smem()
{
    return (unsigned char *)(((uintptr_t)t_dyn_smem.data() + 15) & ~(uintptr_t)15);
}

void enqueue_kernel(cudaStream_t s, unsigned grid, unsigned block, size_t smem, bool coop, const char *name, std::function<void()> body)
{
    std::unique_lock<std::mutex> lk(R.mu);
    if (grid == 0 || block == 0 || block > 1024) {  // cudaErrorInvalidConfiguration on hardware
        t_last = cudaErrorInvalidValue;
        return;
    }
    Op op;
    op.kind = OP_WORK;
    op.name = name;
    if (smem > 227 * 1024) {                         // more than an SM has
        t_last = cudaErrorInvalidValue;
        return;
    }
    op.fn = [grid, block, smem, coop, body = std::move(body)]() { run_kernel(grid, block, smem, coop, body); };
    enqueue(resolve(s), std::move(op));
}
}  // namespace mokab_sim

// ---- runtime API -------------------------------------------------------------------------------------------------
extern "C" {

const char *cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNoDevice: return "no CUDA-capable device is detected";
    case cudaErrorStreamCaptureUnsupported: return "operation not permitted when stream is capturing";
    case cudaErrorStreamCaptureInvalidated: return "operation failed due to a previous error during capture";
    case cudaErrorStreamCaptureUnjoined: return "capturing stream has unjoined work";
    case cudaErrorStreamCaptureIsolation: return "dependency created on uncaptured work in another stream";
    default: return "unknown error";
    }
}

cudaError_t cudaGetLastError(void)
{
    cudaError_t e = t_last;
    t_last = cudaSuccess;
    return e;
}

cudaError_t cudaGetDeviceCount(int *n)
{
    *n = 1;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int device) { return device == 0 ? cudaSuccess : fail(cudaErrorInvalidValue); }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *prop, int)
{
    memset(prop, 0, sizeof(*prop));
    snprintf(prop->name, sizeof(prop->name), "simulated sm_100a device (tests/sim)");
    prop->major = 10; prop->minor = 0; prop->multiProcessorCount = 148;
    return cudaSuccess;
}

cudaError_t cudaDeviceSynchronize(void)
{
    std::unique_lock<std::mutex> lk(R.mu);
    if (capture_blocks_this_thread()) return fail(cudaErrorStreamCaptureUnsupported);
    return wait_until(lk, nullptr, [] { return !pending_anywhere(); }, "cudaDeviceSynchronize");
}

cudaError_t cudaMalloc(void **p, size_t bytes)
{
    std::unique_lock<std::mutex> lk(R.mu);
    if (capture_blocks_this_thread()) return fail(cudaErrorStreamCaptureUnsupported);
    void *q = malloc(bytes ? bytes : 1);
    if (!q) return fail(cudaErrorMemoryAllocation);
    memset(q, 0xFF, bytes);
    R.dev[q] = bytes;
    *p = q;
    return cudaSuccess;
}

cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    std::unique_lock<std::mutex> lk(R.mu);
    if (capture_blocks_this_thread()) return fail(cudaErrorStreamCaptureUnsupported);
    auto it = R.dev.find(p);
    if (it == R.dev.end()) return fail(cudaErrorInvalidValue);
    wait_until(lk, nullptr, [] { return !pending_anywhere(); }, "cudaFree");  // cudaFree synchronises the device
    it = R.dev.find(p);
    memset(p, 0xFF, it->second);
    R.dev.erase(it);
    free(p);
    return cudaSuccess;
}

cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned)
{
    std::unique_lock<std::mutex> lk(R.mu);
    void *q = malloc(bytes ? bytes : 1);
    if (!q) return fail(cudaErrorMemoryAllocation);
    memset(q, 0, bytes);
    R.pinned[q] = bytes;
    *p = q;
    return cudaSuccess;
}

cudaError_t cudaFreeHost(void *p)
{
    if (!p) return cudaSuccess;
    std::unique_lock<std::mutex> lk(R.mu);
    auto it = R.pinned.find(p);
    if (it == R.pinned.end()) return fail(cudaErrorInvalidValue);
    wait_until(lk, nullptr, [] { return !pending_anywhere(); }, "cudaFreeHost");
    R.pinned.erase(p);
    free(p);
    return cudaSuccess;
}

cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind)
{
    // legacy default stream; the library's streams are all cudaStreamNonBlocking, so nothing is implied about them
    std::unique_lock<std::mutex> lk(R.mu);
    if (capture_blocks_this_thread()) return fail(cudaErrorStreamCaptureUnsupported);
    memmove(dst, src, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s_)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    Op op;
    op.kind = OP_WORK;
    op.name = "memcpy";
    if (kind == cudaMemcpyHostToDevice && !is_pinned(src, bytes)) {
        // pageable source: staged before the call returns
        auto stage = std::make_shared<std::vector<char>>((const char *)src, (const char *)src + bytes);
        op.fn = [dst, stage]() { memcpy(dst, stage->data(), stage->size()); };
        enqueue(s, std::move(op));
        return cudaSuccess;
    }
    op.fn = [dst, src, bytes]() { memmove(dst, src, bytes); };
    if (kind == cudaMemcpyDeviceToHost && !is_pinned(dst, bytes)) {
        // pageable destination: the call returns once the copy has completed
        if (s->cap) { s->cap->invalid = true; return fail(cudaErrorStreamCaptureUnsupported); }
        enqueue(s, std::move(op));
        return wait_until(lk, s, [s] { return s->q.empty(); }, "cudaMemcpyAsync(D2H, pageable)");
    }
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t s)
{
    std::unique_lock<std::mutex> lk(R.mu);
    Op op;
    op.kind = OP_WORK;
    op.name = "memset";
    op.fn = [dst, value, bytes]() { memset(dst, value, bytes); };
    enqueue(resolve(s), std::move(op));
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int priority)
{
    std::unique_lock<std::mutex> lk(R.mu);
    auto *x = new mokab_sim_stream();
    x->id = R.next_stream_id++;
    x->device = t_device;
    x->priority = priority;
    R.streams.push_back(x);
    *s = x;
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags) { return cudaStreamCreateWithPriority(s, flags, 0); }

cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    if (!s) return fail(cudaErrorInvalidValue);
    std::unique_lock<std::mutex> lk(R.mu);
    wait_until(lk, s, [s] { return s->q.empty(); }, "cudaStreamDestroy");
    R.streams.erase(std::remove(R.streams.begin(), R.streams.end(), s), R.streams.end());
    delete s;
    return cudaSuccess;
}

cudaError_t cudaStreamSynchronize(cudaStream_t s_)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    if (s->cap) { s->cap->invalid = true; return fail(cudaErrorStreamCaptureUnsupported); }
    return wait_until(lk, s, [s] { return s->q.empty(); }, "cudaStreamSynchronize");
}

cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)
{
    *e = new mokab_sim_event();
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    // queued records / waits may still name it: events are tiny, keep them alive (test infrastructure)
    (void)e;
    return cudaSuccess;
}

cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s_)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    if (s->cap) {
        e->cap = s->cap;
        e->frontier = s->frontier;
        return cudaSuccess;
    }
    e->cap = nullptr;
    e->recorded = ++R.ticket;
    e->rec_stream = s;
    Op op;
    op.kind = OP_RECORD;
    op.ev = e;
    op.ticket = e->recorded;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaStreamWaitEvent(cudaStream_t s_, cudaEvent_t e, unsigned)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    if (e->cap && !R.captures.count(e->cap)) return fail(cudaErrorInvalidValue);  // recorded in a capture that has ended
    if (e->cap) {  // the event belongs to an ongoing capture: fork / join
        mokab_sim_graph *g = e->cap;
        if (s->cap && s->cap != g) { g->invalid = true; return fail(cudaErrorStreamCaptureIsolation); }
        if (!s->cap) {
            s->cap = g;
            s->frontier.clear();
            g->members.push_back(s);
        }
        for (int d : e->frontier)
            if (std::find(s->frontier.begin(), s->frontier.end(), d) == s->frontier.end()) s->frontier.push_back(d);
        return cudaSuccess;
    }
    if (s->cap) {  // a capturing stream may not depend on uncaptured work
        s->cap->invalid = true;
        return fail(cudaErrorStreamCaptureIsolation);
    }
    if (e->recorded == 0 || e->done >= e->recorded) return cudaSuccess;  // never recorded / already complete: no-op
    Op op;
    op.kind = OP_WAIT;
    op.ev = e;
    op.ticket = e->recorded;
    op.src = e->rec_stream;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaEventSynchronize(cudaEvent_t e)
{
    std::unique_lock<std::mutex> lk(R.mu);
    if (e->cap) return fail(cudaErrorStreamCaptureUnsupported);
    const uint64_t t = e->recorded;
    return wait_until(lk, e->rec_stream, [e, t] { return e->done >= t; }, "cudaEventSynchronize");
}

cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    std::unique_lock<std::mutex> lk(R.mu);
    if (a->done < a->recorded || b->done < b->recorded) return fail(cudaErrorInvalidValue);
    *ms = (float)std::chrono::duration<double, std::milli>(b->when - a->when).count();
    return cudaSuccess;
}

cudaError_t cudaStreamBeginCapture(cudaStream_t s_, cudaStreamCaptureMode mode)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    if (s->cap) return fail(cudaErrorStreamCaptureUnsupported);
    auto *g = new mokab_sim_graph();
    g->origin = s;
    g->mode = mode;
    g->members.push_back(s);
    s->cap = g;
    s->frontier.clear();
    R.captures.insert(g);
    t_capturing++;
    return cudaSuccess;
}

cudaError_t cudaStreamEndCapture(cudaStream_t s_, cudaGraph_t *out)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    *out = nullptr;
    mokab_sim_graph *g = s->cap;
    if (!g || g->origin != s) return fail(cudaErrorInvalidValue);
    // every captured node without a successor must be what the origin stream currently depends on: anything else is
    // work forked to another stream and never joined back
    std::vector<char> has_succ(g->nodes.size(), 0);
    for (const Node &n : g->nodes)
        for (int d : n.deps) has_succ[d] = 1;
    bool unjoined = false;
    for (int i = 0; i < (int)g->nodes.size(); ++i)
        if (!has_succ[i] && std::find(s->frontier.begin(), s->frontier.end(), i) == s->frontier.end()) unjoined = true;
    for (cudaStream_t m : g->members) {
        m->cap = nullptr;
        m->frontier.clear();
    }
    R.captures.erase(g);
    t_capturing--;
    if (g->invalid || unjoined) {
        const bool inv = g->invalid;
        delete g;
        return fail(inv ? cudaErrorStreamCaptureInvalidated : cudaErrorStreamCaptureUnjoined);
    }
    *out = g;
    return cudaSuccess;
}

cudaError_t cudaGraphInstantiate(cudaGraphExec_t *ge, cudaGraph_t g, unsigned long long)
{
    std::unique_lock<std::mutex> lk(R.mu);
    auto *x = new mokab_sim_graph_exec();
    x->nodes = g->nodes;
    *ge = x;
    return cudaSuccess;
}

cudaError_t cudaGraphDestroy(cudaGraph_t g)
{
    delete g;
    return cudaSuccess;
}

cudaError_t cudaGraphExecDestroy(cudaGraphExec_t ge)
{
    std::unique_lock<std::mutex> lk(R.mu);
    wait_until(lk, nullptr, [] { return !pending_anywhere(); }, "cudaGraphExecDestroy");
    delete ge;
    return cudaSuccess;
}

cudaError_t cudaGraphLaunch(cudaGraphExec_t ge, cudaStream_t s_)
{
    std::unique_lock<std::mutex> lk(R.mu);
    cudaStream_t s = resolve(s_);
    if (s->cap) return fail(cudaErrorStreamCaptureUnsupported);
    Op op;
    op.kind = OP_GRAPH;
    op.name = "graph";
    op.run = std::make_shared<GraphRun>();
    op.run->g = ge;
    op.run->done.assign(ge->nodes.size(), 0);
    op.run->seq.assign(ge->nodes.size(), 0);
    for (size_t i = 0; i < ge->nodes.size(); ++i)
        if (ge->nodes[i].kind == OP_COLL) {
            Comm *c = ge->nodes[i].coll.comm;
            const int r = ge->nodes[i].coll.rank;
            op.run->seq[i] = ++c->issued[r];
            c->stream[r] = s;
        }
    if (ge->nodes.empty()) return cudaSuccess;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void *, enum cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return fail(cudaErrorUnknown); }       // one process: never needed
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return fail(cudaErrorUnknown); }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

// ---- simulation controls (tests/sim/simcuda.py) ----------------------------------------------------------------------
void mokab_sim_set_policy(int policy, uint64_t seed)
{
    std::unique_lock<std::mutex> lk(R.mu);
    R.policy = (Policy)policy;
    R.rng.seed(seed);
}

void mokab_sim_stats(int64_t *out)  // kernels run, stream ops run, graph nodes run, collectives completed, pending ops, live device bytes
{
    std::unique_lock<std::mutex> lk(R.mu);
    out[0] = R.n_kernels; out[1] = R.n_ops; out[2] = R.n_graph_nodes; out[3] = R.n_colls;
    int64_t pend = 0;
    for (cudaStream_t s : R.streams) pend += (int64_t)s->q.size();
    out[4] = pend;
    int64_t bytes = 0;
    for (auto &kv : R.dev) bytes += (int64_t)kv.second;
    out[5] = bytes;
}

void mokab_sim_set_thread_device(int device) { t_device = device; }

void mokab_sim_set_deadlock_seconds(double s)
{
    std::unique_lock<std::mutex> lk(R.mu);
    R.deadlock_s = s;
}

void *mokab_sim_comm_create(int nranks)
{
    auto *c = new Comm();
    c->n = nranks;
    c->issued.assign(nranks, 0);
    c->arrived.assign(nranks, 0);
    c->args.resize(nranks);
    c->stream.assign(nranks, nullptr);
    return c;
}

void mokab_sim_comm_destroy(void *c) { delete (Comm *)c; }

// In-stream all-to-all (the NCCL collective's stand-in): rank `rank` sends scnt[q] elements to every rank q from
// consecutive segments of `send` and receives rcnt[q] from q into consecutive segments of `recv`.
int mokab_sim_all_to_all(void *comm, int rank, cudaStream_t s_, void *send, void *recv, const int64_t *scnt, const int64_t *rcnt,
                         int64_t elem_bytes)
{
    std::unique_lock<std::mutex> lk(R.mu);
    Comm *c = (Comm *)comm;
    cudaStream_t s = resolve(s_);
    Op op;
    op.kind = OP_COLL;
    op.name = "all_to_all";
    op.coll.comm = c; op.coll.rank = rank; op.coll.send = (char *)send; op.coll.recv = (char *)recv;
    op.coll.scnt.assign(scnt, scnt + c->n); op.coll.rcnt.assign(rcnt, rcnt + c->n); op.coll.elem = elem_bytes;
    if (!s->cap) {
        op.coll.seq = ++c->issued[rank];
        c->stream[rank] = s;
    }
    enqueue(s, std::move(op));
    return 0;
}

}  // extern "C"
