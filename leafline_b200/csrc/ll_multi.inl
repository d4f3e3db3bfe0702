// ll_multi.inl -- several GPUs sharing ONE tree, behind the C ABI (included at the end of ll_api.cu).
//
// The reference fans a search out inside the engine: one thread per root move, results gathered over a channel
// (src/mind.rs:399-437).  Here a group (ll_mctx) owns one ll_ctx per GPU, the communicator between them and the
// device-resident reduction of their results, so that a caller of the C ABI -- Rust, C++, Python -- gets the whole fan-out
// from one call (SURVEY.md 8(b) ownership row, 8(e)):
//   * one process, n GPUs   (ll_init_multi):      ncclCommInitAll, one persistent host thread + stream per device;
//   * one process per GPU   (ll_init_multi_rank): ncclCommInitRank with an id the caller carried to every process
//                                                 (torchrun, MPI, a pipe); the call is collective, SPMD like NCCL itself.
// The tree is dealt by subtree, no position ever crosses NVLink; what does cross it is a u64 sum of leaf counts
// (<= 8 KiB) or an element-wise max of <= a few thousand f32 keys -- latency-bound, NVLink bandwidth is irrelevant here.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2 -- in a torch process that is the copy torch already mapped), so the
// library has no link-time dependency on it and single-GPU callers never touch it.
#include <array>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <dlfcn.h>
#include <functional>
#include <mutex>
#include <thread>

namespace {

// ---- the subset of nccl.h this file needs (NCCL 2.x ABI: nccl.h:37-38, 146-186, 215, 260-285, 392, 493-503) ----------
namespace nccl {
typedef struct ncclComm *Comm;
struct UniqueId {
    char internal[128];
};
static_assert(sizeof(UniqueId) == LL_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
enum { kSuccess = 0 };
enum { kInt32 = 2, kUint64 = 5, kFloat32 = 7 }; // ncclDataType_t
enum { kSum = 0, kMax = 2 };                    // ncclRedOp_t
struct Api {
    void *handle = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommInitAll)(Comm *, int, const int *) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};
Api &api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *override_path = getenv("LEAFLINE_B200_NCCL");
        const char *names[] = {override_path, "libnccl.so.2", "libnccl.so"};
        for (const char *name : names) {
            if (!name || !*name) continue;
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
            a.error = dlerror();
        }
        if (!a.handle) return;
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(a.handle, name);
            if (!p) {
                ok = false;
                a.error = std::string("missing symbol ") + name;
            }
            return p;
        };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        if (!ok) {
            dlclose(a.handle);
            a.handle = nullptr;
        }
    });
    return a;
}
int need() {
    Api &a = api();
    if (!a.handle) return fail(LL_ERR_NCCL, "NCCL is not available (%s)", a.error.empty() ? "libnccl.so.2 not found" : a.error.c_str());
    return LL_OK;
}
} // namespace nccl

#define NCCL_TRY(expr)                                                                                          \
    do {                                                                                                        \
        int r_ = (expr);                                                                                        \
        if (r_ != nccl::kSuccess) return fail(LL_ERR_NCCL, "%s: %s", #expr, nccl::api().GetErrorString(r_));    \
    } while (0)

// result block of one collective call, u64 words: [0] leaves / leaves scored, [1] number of root moves, [2] some rank's level
// outgrew its buffer (redo level by level), [3] some rank failed, [8 + i] per-root counts (the layout of ll_perft_device)
constexpr size_t MW_LEAVES = 0, MW_ROOTS = 1, MW_OVERFLOW = 2, MW_FAILED = 3, MW_TABLE = 8, MW_WORDS = LL_PERFT_DEVICE_WORDS;
constexpr int MAX_RANKS = 32;          // unit owners travel in five descriptor bits (DescSink)
constexpr size_t PEER_SLOTS = 3;       // accumulators of the peer-memory reduction, rotated call after call
constexpr size_t PEER_WORDS = 8192;    // u64 words per accumulator (count tables, or two f32 keys a word)
constexpr u64 PEER_NO_KEYS = 0x8000000080000000ull; // two INT_MIN keys

struct Member;
struct Job {
    std::function<int(Member &)> run;
};

} // namespace

// host-side rendezvous of the member threads of one process (sense-reversing, spinning: the threads are awake anyway)
struct HostBarrier {
    std::atomic<int> count{0};
    std::atomic<unsigned> sense{0};
    int n = 1;
    bool wait() { // false: the others did not show up within 60 s (a member died on the way)
        const unsigned s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s + 1, std::memory_order_release);
            return true;
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned spins = 0; sense.load(std::memory_order_acquire) == s; spins++) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xFFFFu) == 0xFFFFu && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) return false;
        }
        return true;
    }
};

struct ll_mctx {
    int world = 0;
    HostBarrier barrier;
    bool threaded = false; // one process drives every member through its own host thread
    std::vector<Member *> members;
    int perft_shard_min_depth = 5;   // shallower trees are counted whole by every rank: nothing to gain from dealing them
    int kickoff_shard_min_depth = 6; // (and no collective is paid for)
    bool peer_reduce = false;        // counts / keys reduced by this library's own kernel over NVLink peer memory instead of NCCL
    bool have_nccl = false;
};

namespace {

// The library's own all-reduce over NVLink peer memory (NVSwitch: every GPU reaches every peer directly).  Each member owns,
// in its HBM, two windows -- one for u64 sums, one for signed 32-bit maxima -- of PEER_SLOTS accumulators, an arrival counter
// and a call counter; every member maps every other member's (one process: peer access; one process per GPU: CUDA IPC handles
// exchanged once at init through ncclAllGather).
struct PeerArgs { // one window as the kernel sees it
    u64 *acc[MAX_RANKS];
    u32 *arrived[MAX_RANKS];
    u32 *epoch;
    int rank, world;
};
constexpr size_t PEER_WINDOW_BYTES = PEER_SLOTS * PEER_WORDS * sizeof(u64);
constexpr size_t PEER_MEM_BYTES = 2 * PEER_WINDOW_BYTES + 256; // + arrival counters (2 x u32 at +0, +64) and epochs (+128, +192)

struct Member {
    ll_mctx *group = nullptr;
    ll_ctx *ctx = nullptr;
    int rank = 0;
    nccl::Comm comm = nullptr;
    DevBuf words, keys, peer_mem;
    u64 *host_words = nullptr; // pinned, MW_WORDS + a little
    std::vector<float> host_scores;
    PeerArgs sum_win{}, max_win{};
    void *opened[MAX_RANKS] = {}; // IPC mappings to close
    // the member's host thread (one process, n GPUs)
    std::thread thread;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> posted{0}, done{0};
    bool quit = false;
    const Job *job = nullptr;
    int rc = LL_OK;
    char err[sizeof g_error] = "";
};

void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}

void worker_main(Member *m) {
    cudaSetDevice(m->ctx->device);
    uint64_t seen = 0;
    for (;;) {
        int spins = 0;
        while (m->posted.load(std::memory_order_acquire) == seen) {
            if (++spins < 200000) { // back-to-back calls find the thread awake (~1 us hand-over instead of a futex wake-up)
                cpu_relax();
            } else {
                std::unique_lock<std::mutex> lk(m->mu);
                m->cv.wait(lk, [&] { return m->posted.load(std::memory_order_acquire) != seen; });
            }
        }
        seen = m->posted.load(std::memory_order_acquire);
        if (m->quit) return;
        int rc;
        try {
            rc = m->job->run(*m);
        } catch (const std::bad_alloc &) {
            rc = fail(LL_ERR_OOM, "host allocation failed");
        } catch (...) {
            rc = fail(LL_ERR_CUDA, "unexpected exception");
        }
        m->rc = rc;
        if (rc != LL_OK) memcpy(m->err, g_error, sizeof m->err);
        m->done.store(seen, std::memory_order_release);
    }
}

// run `job` on every local member (its own thread when threaded, inline otherwise); first failure wins
int run_on_all(ll_mctx *g, const Job &job) {
    if (!g->threaded) {
        for (Member *m : g->members) {
            CUDA_TRY(cudaSetDevice(m->ctx->device));
            LL_TRY(job.run(*m));
        }
        return LL_OK;
    }
    for (Member *m : g->members) {
        m->job = &job;
        {
            std::lock_guard<std::mutex> lk(m->mu);
            m->posted.fetch_add(1, std::memory_order_release);
        }
        m->cv.notify_one();
    }
    int rc = LL_OK;
    for (Member *m : g->members) {
        const uint64_t want = m->posted.load(std::memory_order_acquire);
        int spins = 0;
        while (m->done.load(std::memory_order_acquire) != want) {
            if (++spins < 4000) cpu_relax();
            else std::this_thread::yield();
        }
        if (m->rc != LL_OK && rc == LL_OK) {
            rc = m->rc;
            memcpy(g_error, m->err, sizeof g_error);
        }
    }
    return rc;
}

// ---- the reduction kernel of the peer-memory path --------------------------------------------------------------------------
// In place over `n` u64 words: every rank pushes its words into every rank's accumulator of this call (fire-and-forget
// reductions over NVLink: u64 add, or -- MAXI -- a signed 32-bit max on each half, the order-preserving keys of f32 values),
// announces itself to every rank, waits until all have announced themselves to it and reads its own accumulator back.
// The accumulator of the call after next is re-initialised on the way: every rank has finished reading it (a rank that is
// in call e has seen everybody arrive in call e - 1, i.e. everybody has left call e - 2).  One block.
template <bool MAXI> __global__ void __launch_bounds__(1024) k_peer_allreduce(const PeerArgs a, u64 *__restrict__ words, u32 n) {
    __shared__ u32 epoch_s;
    if (threadIdx.x == 0) epoch_s = *a.epoch;
    __syncthreads();
    const u32 epoch = epoch_s; // calls made so far on this window
    const size_t slot = (size_t)(epoch % PEER_SLOTS) * PEER_WORDS, next = (size_t)((epoch + 1) % PEER_SLOTS) * PEER_WORDS;
    const u64 idle = MAXI ? PEER_NO_KEYS : 0ull;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const u64 v = words[i];
        if (v == idle) continue;
        for (int r = 0; r < a.world; r++) {
            u64 *dst = a.acc[r] + slot + i;
            if (MAXI) {
                if ((u32)v != 0x80000000u) atomicMax_system((int *)dst, (int)(u32)v);
                if ((u32)(v >> 32) != 0x80000000u) atomicMax_system((int *)dst + 1, (int)(u32)(v >> 32));
            } else {
                atomicAdd_system((unsigned long long *)dst, (unsigned long long)v);
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < (u32)a.world) atomicAdd_system(a.arrived[threadIdx.x], 1u);
    if (threadIdx.x == 0) {
        const u32 want = (epoch + 1u) * (u32)a.world;
        while ((int)(*(volatile u32 *)a.arrived[a.rank] - want) < 0) __nanosleep(20);
        __threadfence_system();
    }
    __syncthreads();
    u64 *mine = a.acc[a.rank];
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) words[i] = *(volatile u64 *)(mine + slot + i);
    for (u32 i = threadIdx.x; i < (u32)PEER_WORDS; i += blockDim.x) mine[next + i] = idle;
    if (threadIdx.x == 0) *a.epoch = epoch + 1u;
}

// A collective kernel SPINS on the device until every rank's has started.  Inside one process nobody may launch it while
// another member's host thread can still reach a call that waits for the device (cudaFree / cudaMalloc in a level's
// growth synchronise a device and, with peer access enabled, its peers): the member threads meet on the host first, and
// from there to the collective they only enqueue.
int before_collective(Member &m) {
    if (m.group->threaded && !m.group->barrier.wait()) return fail(LL_ERR_CUDA, "a member of the group never reached the collective");
    return LL_OK;
}

// u64 sum over all ranks, in place, on the member's stream
int allreduce_u64_sum(Member &m, u64 *d, size_t n) {
    if (m.group->world == 1) return LL_OK;
    LL_TRY(before_collective(m));
    if (m.group->peer_reduce && n <= PEER_WORDS) {
        m.ctx->launches++;
        k_peer_allreduce<false><<<1, 1024, 0, m.ctx->stream>>>(m.sum_win, d, (u32)n);
        return check_launch("k_peer_allreduce<sum>");
    }
    NCCL_TRY(nccl::api().AllReduce(d, d, n, nccl::kUint64, nccl::kSum, m.comm, m.ctx->stream));
    return LL_OK;
}
// element-wise max of f32 keys (float_key) over all ranks, in place; n even, the buffer 8-byte aligned
int allreduce_key_max(Member &m, int *d, size_t n) {
    if (m.group->world == 1) return LL_OK;
    LL_TRY(before_collective(m));
    if (m.group->peer_reduce && n / 2 <= PEER_WORDS) {
        m.ctx->launches++;
        k_peer_allreduce<true><<<1, 1024, 0, m.ctx->stream>>>(m.max_win, (u64 *)d, (u32)(n / 2));
        return check_launch("k_peer_allreduce<max>");
    }
    NCCL_TRY(nccl::api().AllReduce(d, d, n, nccl::kInt32, nccl::kMax, m.comm, m.ctx->stream));
    return LL_OK;
}

// ---- small kernels of the collective calls ----------------------------------------------------------------------------------
__global__ void k_fill_u32(u32 *p, size_t n, u32 v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// ll_perft_batch_multi: the total of the chain that just ran -> slot `at`; an outgrown level is counted in *outgrown
__global__ void k_take_total(const u64 *leaves, const u32 *overflow, u64 *out, u32 at, u64 *outgrown) {
    out[at] = *overflow ? 0ull : *leaves;
    if (*overflow) *outgrown += 1;
}
__global__ void k_add_word(u64 *word, u64 v) { *word += v; }
// search: every unit's value, negated, folded into its parent's key (mind.rs:331-336 across the ranks)
__global__ void __launch_bounds__(BLOCK) k_fold_units(const float *__restrict__ values, const u32 *__restrict__ parent, size_t n, int *keys) {
    const size_t u = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (u < n) atomicMax(keys + parent[u], float_key(-values[u]));
}
// ... and back: the value of every parent of units -- the reduced maximum, or the leaf rule for a parent nobody could move from
__global__ void __launch_bounds__(BLOCK) k_unfold_parents(Soa parents, size_t n, const u32 *__restrict__ counts, const int *__restrict__ keys, float *__restrict__ values) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    if (counts[i]) {
        values[i] = key_float(keys[i]);
    } else { // mind.rs:281-283
        const Pos p = soa_load(parents, i);
        const float s = score(p);
        values[i] = meta_stm(p.meta) ? -s : s;
    }
}
// the tail of the key vector carries what is not a maximum through the same collective: slot pair `rank` of the first
// 2 * world keys holds this rank's leaf count (31 + 31 bits; the other ranks hold INT_MIN there, so max == gather), then a
// failure flag
__global__ void k_key_tail(int *tail, int rank, int world, const u64 *scored, int failed) {
    const u64 s = scored ? *scored : 0ull;
    tail[2 * rank] = (int)(s & 0x7FFFFFFFull);
    tail[2 * rank + 1] = (int)((s >> 31) & 0x7FFFFFFFull);
    tail[2 * world] = failed;
}

// ---- group construction -------------------------------------------------------------------------------------------------------
// rank r's window, whose memory this member sees at `base`
void fill_window(PeerArgs &w, int r, char *base, bool max_window) {
    w.acc[r] = (u64 *)(base + (max_window ? PEER_WINDOW_BYTES : 0));
    w.arrived[r] = (u32 *)(base + 2 * PEER_WINDOW_BYTES + (max_window ? 64 : 0));
}

int member_alloc(Member &m) {
    ll_ctx *ctx = m.ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    LL_TRY(ensure(ctx, m.words, (MW_WORDS + 64) * sizeof(u64)));
    LL_TRY(ensure(ctx, m.keys, 4096 * sizeof(int)));
    CUDA_TRY(cudaMallocHost((void **)&m.host_words, (MW_WORDS + 64) * sizeof(u64)));
    // the peer windows: plain cudaMalloc (IPC-exportable), sums zeroed, maxima at INT_MIN, counters zero
    CUDA_TRY(cudaMalloc(&m.peer_mem.ptr, PEER_MEM_BYTES));
    m.peer_mem.cap = PEER_MEM_BYTES;
    CUDA_TRY(cudaMemsetAsync(m.peer_mem.ptr, 0, PEER_MEM_BYTES, ctx->stream));
    k_fill_u32<<<64, 256, 0, ctx->stream>>>((u32 *)((char *)m.peer_mem.ptr + PEER_WINDOW_BYTES), PEER_WINDOW_BYTES / sizeof(u32), 0x80000000u);
    LL_TRY(check_launch("k_fill_u32"));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    char *base = (char *)m.peer_mem.ptr;
    for (PeerArgs *w : {&m.sum_win, &m.max_win}) {
        w->rank = m.rank;
        w->world = m.group->world;
    }
    m.sum_win.epoch = (u32 *)(base + 2 * PEER_WINDOW_BYTES + 128);
    m.max_win.epoch = (u32 *)(base + 2 * PEER_WINDOW_BYTES + 192);
    return LL_OK;
}

void destroy_group(ll_mctx *g) {
    if (!g) return;
    for (Member *m : g->members) {
        if (m->thread.joinable()) {
            {
                std::lock_guard<std::mutex> lk(m->mu);
                m->quit = true;
                m->posted.fetch_add(1, std::memory_order_release);
            }
            m->cv.notify_one();
            m->thread.join();
        }
    }
    for (Member *m : g->members) {
        if (m->ctx) {
            cudaSetDevice(m->ctx->device);
            cudaStreamSynchronize(m->ctx->stream);
        }
        if (m->comm) nccl::api().CommDestroy(m->comm);
        for (void *p : m->opened)
            if (p) cudaIpcCloseMemHandle(p);
        if (m->host_words) cudaFreeHost(m->host_words);
        drop_buf(m->words);
        drop_buf(m->keys);
        drop_buf(m->peer_mem);
        if (m->ctx) ll_shutdown(m->ctx);
        delete m;
    }
    delete g;
}

// "nccl" / "peer" force the reduction path (A/B runs); anything else: peer memory where every rank can map every other's
int reduce_preference() {
    const char *e = getenv("LEAFLINE_B200_REDUCE");
    if (e && !strcmp(e, "nccl")) return 0;
    if (e && !strcmp(e, "peer")) return 2;
    return 1;
}

} // namespace

// ================================================================================================ ABI (multi-GPU)

extern "C" int ll_multi_unique_id(void *id128) {
    LL_ABI_BEGIN
    if (!id128) return fail(LL_ERR_INVALID_ARG, "null id buffer");
    LL_TRY(nccl::need());
    nccl::UniqueId id;
    NCCL_TRY(nccl::api().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_shutdown_multi(ll_mctx *g) {
    LL_ABI_BEGIN
    destroy_group(g);
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_init_multi(const int *devices, int n, ll_mctx **out) {
    LL_ABI_BEGIN
    if (!out || !devices) return fail(LL_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n < 1 || n > MAX_RANKS) return fail(LL_ERR_INVALID_ARG, "%d devices out of range (1..%d)", n, MAX_RANKS);
    ll_mctx *g = new ll_mctx();
    g->world = n;
    g->threaded = n > 1;
    g->barrier.n = n;
    auto bail = [&](int rc) {
        char keep[sizeof g_error];
        memcpy(keep, g_error, sizeof keep);
        destroy_group(g);
        memcpy(g_error, keep, sizeof keep);
        return rc;
    };
    bool distinct = true;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) distinct = distinct && devices[i] != devices[j];
    for (int i = 0; i < n; i++) {
        Member *m = new Member();
        g->members.push_back(m);
        m->group = g;
        m->rank = i;
        int rc = ll_init(devices[i], &m->ctx);
        if (rc == LL_OK) rc = member_alloc(*m);
        if (rc != LL_OK) return bail(rc);
    }
    if (n > 1) {
        // every member maps every other member's windows: peer access inside one process (NVLink / NVSwitch)
        bool peers = true;
        for (int i = 0; i < n && peers; i++) {
            cudaSetDevice(devices[i]);
            for (int j = 0; j < n && peers; j++) {
                if (devices[i] == devices[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) peers = false;
                else {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peers = false;
                    cudaGetLastError();
                }
            }
        }
        for (Member *m : g->members)
            for (Member *o : g->members) {
                fill_window(m->sum_win, o->rank, (char *)o->peer_mem.ptr, false);
                fill_window(m->max_win, o->rank, (char *)o->peer_mem.ptr, true);
            }
        // the communicator (ncclCommInitAll in a single process); members sharing a device (tests on a
        // one-GPU box) cannot form one -- they reduce over peer memory only
        if (distinct && nccl::api().handle) {
            std::vector<nccl::Comm> comms(n);
            const int r = nccl::api().CommInitAll(comms.data(), n, devices);
            if (r != nccl::kSuccess) return bail(fail(LL_ERR_NCCL, "ncclCommInitAll: %s", nccl::api().GetErrorString(r)));
            for (int i = 0; i < n; i++) g->members[i]->comm = comms[i];
            g->have_nccl = true;
        }
        const int pref = reduce_preference();
        g->peer_reduce = peers && (pref >= 1 || !g->have_nccl);
        if (!g->peer_reduce && !g->have_nccl)
            return bail(fail(LL_ERR_NCCL, "no way to reduce across the devices: NCCL unavailable (%s) and no peer access", nccl::api().error.c_str()));
        for (Member *m : g->members) m->thread = std::thread(worker_main, m);
    }
    *out = g;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_init_multi_rank(int device, int rank, int world, const void *unique_id128, ll_mctx **out) {
    LL_ABI_BEGIN
    if (!out) return fail(LL_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world) return fail(LL_ERR_INVALID_ARG, "bad rank %d of %d (at most %d)", rank, world, MAX_RANKS);
    if (world > 1 && !unique_id128) return fail(LL_ERR_INVALID_ARG, "null unique id");
    if (world > 1) LL_TRY(nccl::need());
    ll_mctx *g = new ll_mctx();
    g->world = world;
    auto bail = [&](int rc) {
        char keep[sizeof g_error];
        memcpy(keep, g_error, sizeof keep);
        destroy_group(g);
        memcpy(g_error, keep, sizeof keep);
        return rc;
    };
    Member *m = new Member();
    g->members.push_back(m);
    m->group = g;
    m->rank = rank;
    int rc = ll_init(device, &m->ctx);
    if (rc == LL_OK) rc = member_alloc(*m);
    if (rc != LL_OK) return bail(rc);
    if (world > 1) {
        nccl::UniqueId id;
        memcpy(&id, unique_id128, sizeof id);
        const int r = nccl::api().CommInitRank(&m->comm, world, id, rank);
        if (r != nccl::kSuccess) return bail(fail(LL_ERR_NCCL, "ncclCommInitRank: %s", nccl::api().GetErrorString(r)));
        g->have_nccl = true;
        // exchange the windows' CUDA IPC handles through the communicator itself, map every peer's, then agree (a u64 sum
        // of failures) on whether every rank could: the peer-memory reduction is used by all ranks or by none
        cudaStream_t st = m->ctx->stream;
        struct Wire {
            cudaIpcMemHandle_t handle;
        };
        static_assert(sizeof(Wire) == 64, "cudaIpcMemHandle_t is 64 bytes");
        Wire mine{};
        u64 failures = 0;
        if (cudaIpcGetMemHandle(&mine.handle, m->peer_mem.ptr) != cudaSuccess) {
            cudaGetLastError();
            failures = 1;
        }
        DevBuf wire;
        rc = ensure(m->ctx, wire, sizeof(Wire) * (size_t)world + 64);
        if (rc != LL_OK) return bail(rc);
        std::vector<Wire> all((size_t)world);
        auto step = [&]() -> int {
            CUDA_TRY(cudaMemcpyAsync((char *)wire.ptr + sizeof(Wire) * (size_t)rank, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
            NCCL_TRY(nccl::api().AllGather((char *)wire.ptr + sizeof(Wire) * (size_t)rank, wire.ptr, sizeof(Wire), 0 /* ncclInt8 */, m->comm, st));
            CUDA_TRY(cudaMemcpyAsync(all.data(), wire.ptr, sizeof(Wire) * (size_t)world, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            return LL_OK;
        };
        rc = step();
        if (rc != LL_OK) {
            drop_buf(wire);
            return bail(rc);
        }
        char *bases[MAX_RANKS] = {};
        for (int r2 = 0; r2 < world && !failures; r2++) {
            if (r2 == rank) {
                bases[r2] = (char *)m->peer_mem.ptr;
                continue;
            }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[(size_t)r2].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                failures = 1;
                break;
            }
            m->opened[r2] = p;
            bases[r2] = (char *)p;
        }
        auto agree = [&]() -> int {
            u64 *d = (u64 *)wire.ptr;
            CUDA_TRY(cudaMemcpyAsync(d, &failures, sizeof failures, cudaMemcpyHostToDevice, st));
            NCCL_TRY(nccl::api().AllReduce(d, d, 1, nccl::kUint64, nccl::kSum, m->comm, st));
            CUDA_TRY(cudaMemcpyAsync(&failures, d, sizeof failures, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            return LL_OK;
        };
        rc = agree();
        drop_buf(wire);
        if (rc != LL_OK) return bail(rc);
        const int pref = reduce_preference();
        g->peer_reduce = failures == 0 && pref >= 1;
        if (g->peer_reduce)
            for (int r2 = 0; r2 < world; r2++) {
                fill_window(m->sum_win, r2, bases[r2], false);
                fill_window(m->max_win, r2, bases[r2], true);
            }
    }
    *out = g;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_multi_world(ll_mctx *g) { return g ? g->world : 0; }
extern "C" int ll_multi_local_count(ll_mctx *g) { return g ? (int)g->members.size() : 0; }
extern "C" ll_ctx *ll_multi_ctx(ll_mctx *g, int local_index) {
    if (!g || local_index < 0 || local_index >= (int)g->members.size()) return nullptr;
    return g->members[(size_t)local_index]->ctx;
}
/* 0 = NCCL, 1 = this library's reduction over NVLink peer memory */
extern "C" int ll_multi_reduce_path(ll_mctx *g) { return g && g->peer_reduce ? 1 : 0; }
extern "C" int ll_multi_set_shard_min_depth(ll_mctx *g, int perft_depth, int kickoff_depth) {
    if (!g) return fail(LL_ERR_INVALID_ARG, "null group");
    if (perft_depth > 0) g->perft_shard_min_depth = perft_depth;
    if (kickoff_depth > 0) g->kickoff_shard_min_depth = kickoff_depth;
    return LL_OK;
}

// ================================================================================================ collective calls
namespace {

// words (device) <-reduce-> all ranks, then to the member's pinned block; one synchronisation
int exchange_words(Member &m, size_t n) {
    u64 *d = (u64 *)m.words.ptr;
    LL_TRY(allreduce_u64_sum(m, d, n));
    CUDA_TRY(cudaMemcpyAsync(m.host_words, d, n * sizeof(u64), cudaMemcpyDeviceToHost, m.ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(m.ctx->stream));
    return LL_OK;
}
int some_rank_failed(Member &m, int own_rc) {
    if (own_rc != LL_OK) return own_rc; // g_error holds this rank's cause
    return fail(LL_ERR_CUDA, "another rank of the group failed in this call (its ll_last_error has the cause)");
}

// ---- perft of ONE tree, dealt by subtree (unit_hash) ---------------------------------------------------------------------
int member_perft_tree(Member &m, const ll_world_t *root, int depth) {
    ll_ctx *ctx = m.ctx;
    const int W = m.group->world;
    u64 *host = m.host_words;
    if (W == 1 || depth < m.group->perft_shard_min_depth || host_domain_class(root) == HOST_LITERAL) { // every rank counts the whole (small, or literal-path) tree: no collective
        uint32_t n = 0;
        LL_TRY(ll_perft(ctx, root, depth, 0, 1, (uint64_t *)&host[MW_LEAVES], (uint64_t *)(host + MW_TABLE), 1024, &n));
        host[MW_ROOTS] = n;
        return LL_OK;
    }
    u64 *d = (u64 *)m.words.ptr;
    // attempt 0: the launch chain, counts left on the device, reduced there.  attempt 1 (some rank's level outgrew its
    // buffer, or was never sized): every rank redoes its share through ll_perft, which re-sizes, and the counts are reduced again
    int own = LL_OK;
    bool launched = false;
    if (!host_well_formed(root)) own = classify_root(ctx, root, 0); // (a literal-path root never gets here)
    if (own == LL_OK) own = upload_root(ctx, root);
    if (own == LL_OK && !ctx->force_synced) own = perft_chain_launch(ctx, depth, m.rank, W, &launched);
    if (own == LL_OK && launched) {
        u64 *sc = (u64 *)ctx->scalars.ptr;
        Timed t(ctx, "pack_perft_result");
        k_pack_perft_result<<<1, 1024, 0, ctx->stream>>>(sc + SC_LEAVES, sc + SC_N + 1, (const u32 *)(sc + SC_OVERFLOW), sc + SC_PER_ROOT, d);
        own = check_launch("k_pack_perft_result");
    }
    if (own != LL_OK || !launched) {
        memset(host, 0, MW_WORDS * sizeof(u64));
        host[own != LL_OK ? MW_FAILED : MW_OVERFLOW] = 1;
        CUDA_TRY(cudaMemcpyAsync(d, host, MW_WORDS * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    }
    LL_TRY(exchange_words(m, MW_WORDS));
    if (host[MW_FAILED]) return some_rank_failed(m, own);
    if (host[MW_OVERFLOW]) {
        uint64_t leaves = 0;
        uint32_t n = 0;
        std::vector<uint64_t> table(1024, 0);
        own = ll_perft(ctx, root, depth, m.rank, W, &leaves, table.data(), 1024, &n);
        memset(host, 0, MW_WORDS * sizeof(u64));
        if (own == LL_OK) {
            host[MW_LEAVES] = leaves;
            host[MW_ROOTS] = n;
            memcpy(host + MW_TABLE, table.data(), 1024 * sizeof(u64));
        } else {
            host[MW_FAILED] = 1;
        }
        CUDA_TRY(cudaMemcpyAsync(d, host, MW_WORDS * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        LL_TRY(exchange_words(m, MW_WORDS));
        if (host[MW_FAILED]) return some_rank_failed(m, own);
    }
    host[MW_ROOTS] /= (u64)W; // every rank reported the same number of root moves
    return LL_OK;
}

// ---- perft of a BATCH of roots, dealt by root (root i -> rank i % world): weak scaling ------------------------------------
int member_perft_batch(Member &m, const ll_world_t *roots, uint32_t n, int depth) {
    ll_ctx *ctx = m.ctx;
    const int W = m.group->world;
    u64 *host = m.host_words, *d = (u64 *)m.words.ptr;
    const size_t words = (size_t)n + 2; // [n] outgrown levels, [n + 1] failures
    int own = LL_OK;
    CUDA_TRY(cudaMemsetAsync(d, 0, words * sizeof(u64), ctx->stream));
    bool unsized = depth < 3 || ctx->force_synced; // no launch chain for this depth / these buffers yet: everybody redoes below
    if (!unsized) {
        u64 *sc = (u64 *)ctx->scalars.ptr;
        for (uint32_t i = (uint32_t)m.rank; i < n && own == LL_OK && !unsized; i += (uint32_t)W) {
            bool launched = false;
            const int cls = host_domain_class(&roots[i]);
            if (cls == HOST_REJECT) own = classify_root(ctx, &roots[i], i);
            if (cls == HOST_LITERAL) { // answered by ll_perft's literal path in the redo below
                unsized = true;
                break;
            }
            if (own == LL_OK) own = upload_root(ctx, &roots[i]);
            if (own == LL_OK) own = perft_chain_launch(ctx, depth, 0, 1, &launched);
            if (own != LL_OK) break;
            if (!launched) {
                unsized = true;
                break;
            }
            ctx->launches++;
            k_take_total<<<1, 1, 0, ctx->stream>>>(sc + SC_LEAVES, (const u32 *)(sc + SC_OVERFLOW), d, i, d + n);
            own = check_launch("k_take_total");
        }
    }
    if (own == LL_OK && unsized) {
        ctx->launches++;
        k_add_word<<<1, 1, 0, ctx->stream>>>(d + n, 1ull);
        own = check_launch("k_add_word");
    }
    if (own != LL_OK) {
        memset(host, 0, words * sizeof(u64));
        host[n + 1] = 1;
        CUDA_TRY(cudaMemcpyAsync(d, host, words * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    }
    LL_TRY(exchange_words(m, words));
    if (host[n + 1]) return some_rank_failed(m, own);
    if (host[n]) { // redo this rank's roots through the path that sizes the levels
        std::vector<uint64_t> mine((size_t)n, 0);
        for (uint32_t i = (uint32_t)m.rank; i < n && own == LL_OK; i += (uint32_t)W) own = ll_perft(ctx, &roots[i], depth, 0, 1, &mine[i], nullptr, 0, nullptr);
        memset(host, 0, words * sizeof(u64));
        if (own == LL_OK) memcpy(host, mine.data(), (size_t)n * sizeof(u64));
        else host[n + 1] = 1;
        CUDA_TRY(cudaMemcpyAsync(d, host, words * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        LL_TRY(exchange_words(m, words));
        if (host[n + 1]) return some_rank_failed(m, own);
    }
    return LL_OK;
}

// ---- kickoff of ONE tree: the levels down to the unit ply on every rank (canonical order, so unit numbers agree), the units
// dealt round-robin, their values folded into their parents' keys, one max-reduction, the back-up to the root moves ----------
struct KickoffOut {
    ll_commit_t *roots;
    float *scores;
    uint32_t root_cap;
    uint32_t n_roots;
    uint64_t leaves_scored;
};
int member_kickoff(Member &m, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, KickoffOut &out) {
    ll_ctx *ctx = m.ctx;
    const int W = m.group->world;
    out.n_roots = 0;
    out.leaves_scored = 0;
    if (W == 1 || depth < 3 || depth < m.group->kickoff_shard_min_depth || (root && host_domain_class(root) == HOST_LITERAL)) // small or literal-path trees: every rank searches the whole, no collective
        return ll_kickoff(ctx, root, depth, quiet_extension, nihilistically, 0, 1, out.roots, out.scores, out.root_cap, &out.n_roots, &out.leaves_scored);
    if (depth > 9) return fail(LL_ERR_INVALID_ARG, "depth %d out of range (1..9)", depth);
    if (quiet_extension > 8) return fail(LL_ERR_INVALID_ARG, "quiet extension %d out of range", quiet_extension);
    const size_t s = depth - 1 < 3 ? (size_t)(depth - 1) : 3; // unit ply (>= 2): the units keep at least one ply, so the bottom ply is k_horizon's as on one GPU
    u64 *sc = (u64 *)ctx->scalars.ptr;
    // everything that can fail on one rank alone happens before the collective and is voted on inside it
    u64 n_roots = 0, n_parents = 0, n_units = 0;
    size_t kept = 0;
    uint64_t scored = 0;
    auto local = [&]() -> int {
        if (!root || !out.scores) return fail(LL_ERR_INVALID_ARG, "null argument");
        LL_TRY(classify_root(ctx, root, 0));
        LL_TRY(upload_root(ctx, root));
        LL_TRY(count_level(ctx, 0, nihilistically != 0, &n_roots)); // mind.rs:388-392
        if (n_roots > out.root_cap) return fail(LL_ERR_CAPACITY, "%llu root moves do not fit root_cap %u", (unsigned long long)n_roots, out.root_cap);
        if (!n_roots) return LL_OK;
        LL_TRY(emit_level(ctx, 0, nihilistically != 0, n_roots, true, 2));
        if (out.roots) {
            Frontier &g1 = ctx->levels[1];
            LL_TRY(ensure(ctx, ctx->aos_out, (size_t)n_roots * sizeof(ll_commit_t)));
            {
                Timed t(ctx, "commits_to_aos");
                k_commits_to_aos<<<(unsigned)nblocks((size_t)n_roots), BLOCK, 0, ctx->stream>>>(g1.soa(), (const u32 *)g1.cmeta.ptr, (size_t)n_roots, (u64 *)ctx->aos_out.ptr);
            }
            LL_TRY(check_launch("k_commits_to_aos"));
            CUDA_TRY(cudaMemcpyAsync(out.roots, ctx->aos_out.ptr, (size_t)n_roots * sizeof(ll_commit_t), cudaMemcpyDeviceToHost, ctx->stream));
        }
        u64 level_n = n_roots;
        for (size_t l = 1; l < s; l++) { // reckless levels 2 .. s; level s remembers every node's parent (root_mode 3)
            u64 total = 0;
            LL_TRY(count_level(ctx, l, true, &total));
            LL_TRY(emit_level(ctx, l, true, total, false, l + 1 == s ? 3 : 0));
            n_parents = level_n;
            level_n = total;
        }
        n_units = level_n;
        kept = n_units > (u64)m.rank ? (size_t)((n_units - m.rank + W - 1) / W) : 0;
        LL_TRY(reserve_level(ctx, s + 1, kept, false, true));
        Frontier &units = ctx->levels[s + 1];
        units.n = kept;
        if (kept) {
            Frontier &all = ctx->levels[s];
            Timed t(ctx, "shard_filter");
            k_deal_units<<<(unsigned)std::min<size_t>(nblocks(kept), (size_t)ctx->sm_count * 8), BLOCK, 0, ctx->stream>>>(all.soa(), (const u32 *)all.root.ptr, (size_t)n_units, m.rank, W,
                                                                                                                       units.soa(), (u32 *)units.root.ptr);
            LL_TRY(check_launch("k_deal_units"));
        }
        // the units' exact values (depth - s plies left): one launch chain when the batch is small, else level by level
        const int plies = depth - (int)s;
        bool chained = false;
        if (quiet_extension <= 0 && plies >= 2 && kept && kept <= HORIZON_CHAIN_MAX_BATCH && !ctx->force_synced) {
            bool launched = false;
            LL_TRY(horizon_chain(ctx, s + 1, plies, &launched));
            if (launched) {
                u64 flags[2] = {0, 0};
                CUDA_TRY(cudaMemcpyAsync(&flags[0], sc + SC_OVERFLOW, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(cudaMemcpyAsync(&flags[1], sc + SC_SCORED, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                if (!(u32)flags[0]) {
                    chained = true;
                    scored = flags[1];
                }
            }
        }
        if (!chained && kept) LL_TRY(horizon_levels(ctx, s + 1, plies, quiet_extension, &scored));
        return LL_OK;
    };
    int own = local();
    // n_parents is a pure function of the tree: every rank that got this far agrees on it; a rank that failed early does
    // not know it, so the vote travels in a first tiny collective of its own
    u64 *d = (u64 *)m.words.ptr;
    m.host_words[0] = own != LL_OK;
    m.host_words[1] = 0;
    CUDA_TRY(cudaMemcpyAsync(d, m.host_words, 2 * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    LL_TRY(exchange_words(m, 2));
    if (m.host_words[0]) return some_rank_failed(m, own);
    out.n_roots = (uint32_t)n_roots;
    if (!n_roots) return LL_OK;
    // fold -> reduce -> unfold -> back up
    const size_t n_keys = ((size_t)n_parents + 1) & ~(size_t)1, tail = n_keys, total_keys = tail + 2 * (size_t)W + 2;
    LL_TRY(ensure(ctx, m.keys, (total_keys + 2) * sizeof(int)));
    int *keys = (int *)m.keys.ptr;
    ctx->launches += 3;
    k_fill_u32<<<(unsigned)std::min<size_t>(nblocks(total_keys + 2, 256), 256), 256, 0, ctx->stream>>>((u32 *)keys, total_keys + 2, 0x80000000u);
    LL_TRY(check_launch("k_fill_u32"));
    Frontier &units = ctx->levels[s + 1];
    if (kept) {
        k_fold_units<<<(unsigned)nblocks(kept), BLOCK, 0, ctx->stream>>>((const float *)units.values.ptr, (const u32 *)units.root.ptr, kept, keys);
        LL_TRY(check_launch("k_fold_units"));
    }
    CUDA_TRY(cudaMemcpyAsync(sc + SC_SPARE, &scored, sizeof scored, cudaMemcpyHostToDevice, ctx->stream));
    k_key_tail<<<1, 1, 0, ctx->stream>>>(keys + tail, m.rank, W, sc + SC_SPARE, 0);
    LL_TRY(check_launch("k_key_tail"));
    LL_TRY(allreduce_key_max(m, keys, (total_keys + 1) & ~(size_t)1));
    Frontier &parents = ctx->levels[s - 1];
    LL_TRY(reserve_level(ctx, s - 1, parents.n, false, true));
    {
        Timed t(ctx, "backup");
        k_unfold_parents<<<(unsigned)nblocks((size_t)n_parents), BLOCK, 0, ctx->stream>>>(parents.soa(), (size_t)n_parents, (const u32 *)parents.counts.ptr, keys, (float *)parents.values.ptr);
        LL_TRY(check_launch("k_unfold_parents"));
    }
    for (size_t lev = s - 1; lev > 1; lev--) {
        Frontier &par = ctx->levels[lev - 1];
        Frontier &chi = ctx->levels[lev];
        LL_TRY(reserve_level(ctx, lev - 1, par.n, false, true));
        Timed t(ctx, "backup");
        k_backup<<<(unsigned)nblocks(par.n), BLOCK, 0, ctx->stream>>>(par.soa(), par.n, (const u32 *)par.counts.ptr, (const u64 *)par.block_base.ptr,
                                                                     (const float *)chi.values.ptr, (float *)par.values.ptr, 0);
        LL_TRY(check_launch("k_backup"));
    }
    std::vector<float> vals((size_t)n_roots);
    std::vector<int> tail_host(2 * (size_t)W + 2);
    CUDA_TRY(cudaMemcpyAsync(vals.data(), ctx->levels[1].values.ptr, (size_t)n_roots * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(tail_host.data(), keys + tail, tail_host.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (u64 i = 0; i < n_roots; i++) out.scores[i] = -vals[(size_t)i]; // mind.rs:425
    uint64_t all_scored = 0;
    for (int r = 0; r < W; r++) all_scored += (uint64_t)(u32)tail_host[2 * (size_t)r] | ((uint64_t)(u32)tail_host[2 * (size_t)r + 1] << 31);
    out.leaves_scored = all_scored;
    return LL_OK;
}

// ---- the host alpha-beta driver over the group: the root moves are searched concurrently, one subset per GPU (the reference
// runs one thread per root move, mind.rs:399-414); scores, variations and counters are gathered in ONE max-reduction (every
// rank writes only its own root moves' slots, the others hold INT_MIN there) -------------------------------------------------
int member_alpha_beta(Member &m, const ll_world_t *root, int depth, int device_plies, int quiet_extension, int nihilistically, ll_forecast_t *forecasts,
                      uint32_t cap, uint32_t *n_forecasts, uint64_t *stats3) {
    ll_ctx *ctx = m.ctx;
    const int W = m.group->world;
    if (W == 1) return ll_alpha_beta_forecast(ctx, root, depth, device_plies, quiet_extension, nihilistically, forecasts, cap, n_forecasts, stats3);
    uint32_t n = 0;
    uint64_t stats[3] = {0, 0, 0};
    std::vector<ll_forecast_t> mine(1024);
    int own = check_ctx(ctx);
    if (own == LL_OK) own = alpha_beta_core(ctx, root, depth, device_plies, quiet_extension, nihilistically, mine.data(), 1024, &n, stats, nullptr, nullptr, nullptr, m.rank, W, false);
    if (own == LL_OK && n > cap) own = fail(LL_ERR_CAPACITY, "%u forecasts do not fit the capacity of %u", n, cap);
    u64 *d = (u64 *)m.words.ptr;
    m.host_words[0] = own != LL_OK;
    m.host_words[1] = 0;
    CUDA_TRY(cudaMemcpyAsync(d, m.host_words, 2 * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    LL_TRY(exchange_words(m, 2));
    if (m.host_words[0]) return some_rank_failed(m, own);
    constexpr size_t PER_ROOT = 2 + LL_MAX_VARIATION; // score key, variation length, the patches
    const size_t tail = (size_t)n * PER_ROOT, total = (tail + 6 * (size_t)W + 1) & ~(size_t)1;
    std::vector<int> keys(total, INT_MIN);
    for (uint32_t i = 0; i < n; i++) {
        if ((int)(i % (uint32_t)W) != m.rank) continue;
        int *k = keys.data() + (size_t)i * PER_ROOT;
        const int bits = *(const int *)&mine[i].score;
        k[0] = bits >= 0 ? bits : bits ^ 0x7FFFFFFF; // float_key
        k[1] = (int)mine[i].variation_len;
        for (uint32_t j = 0; j < mine[i].variation_len; j++) {
            const ll_patch_t &p = mine[i].variation[j];
            k[2 + j] = (int)p.team | ((int)p.job << 8) | ((int)p.whence << 16) | ((int)p.whither << 24);
        }
    }
    for (int c = 0; c < 3; c++) {
        keys[tail + 6 * (size_t)m.rank + 2 * c] = (int)(stats[c] & 0x7FFFFFFFull);
        keys[tail + 6 * (size_t)m.rank + 2 * c + 1] = (int)((stats[c] >> 31) & 0x7FFFFFFFull);
    }
    LL_TRY(ensure(ctx, m.keys, (total + 2) * sizeof(int)));
    int *dk = (int *)m.keys.ptr;
    CUDA_TRY(cudaMemcpyAsync(dk, keys.data(), total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    LL_TRY(allreduce_key_max(m, dk, total));
    CUDA_TRY(cudaMemcpyAsync(keys.data(), dk, total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> order(n);
    std::vector<float> scores(n);
    for (uint32_t i = 0; i < n; i++) {
        order[i] = i;
        const int k = keys[(size_t)i * PER_ROOT], bits = k >= 0 ? k : k ^ 0x7FFFFFFF;
        scores[i] = *(const float *)&bits;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return scores[x] > scores[y]; }); // mind.rs:436
    for (uint32_t r = 0; r < n; r++) {
        const uint32_t i = order[r];
        ll_forecast_t &f = forecasts[r];
        f = mine[i]; // the commit (every rank listed every root move)
        f.score = scores[i];
        const int *k = keys.data() + (size_t)i * PER_ROOT;
        f.variation_len = (uint32_t)k[1];
        for (uint32_t j = 0; j < f.variation_len && j < LL_MAX_VARIATION; j++)
            f.variation[j] = ll_patch_t{(uint8_t)(k[2 + j] & 0xFF), (uint8_t)((k[2 + j] >> 8) & 0xFF), (uint8_t)((k[2 + j] >> 16) & 0xFF), (uint8_t)((k[2 + j] >> 24) & 0xFF)};
    }
    *n_forecasts = n;
    if (stats3)
        for (int c = 0; c < 3; c++) {
            stats3[c] = 0;
            for (int r = 0; r < W; r++)
                stats3[c] += (uint64_t)(u32)keys[tail + 6 * (size_t)r + 2 * c] | ((uint64_t)(u32)keys[tail + 6 * (size_t)r + 2 * c + 1] << 31);
        }
    return LL_OK;
}

int check_group(ll_mctx *g) {
    if (!g || g->members.empty()) return fail(LL_ERR_INVALID_ARG, "null group");
    return LL_OK;
}

} // namespace

extern "C" int ll_perft_multi(ll_mctx *g, const ll_world_t *root, int depth, uint64_t *leaves, uint64_t *per_root, uint32_t root_cap, uint32_t *n_roots) {
    LL_ABI_BEGIN
    LL_TRY(check_group(g));
    if (!root || !leaves) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (depth < 0 || depth > 32) return fail(LL_ERR_INVALID_ARG, "depth %d out of range", depth);
    *leaves = 0;
    if (n_roots) *n_roots = 0;
    Job job{[&](Member &m) { return member_perft_tree(m, root, depth); }};
    LL_TRY(run_on_all(g, job));
    const u64 *host = g->members[0]->host_words;
    const u64 root_count = host[MW_ROOTS];
    *leaves = host[MW_LEAVES];
    if (n_roots) *n_roots = (uint32_t)root_count;
    if (per_root) {
        if (root_count > root_cap) return fail(LL_ERR_CAPACITY, "%llu root moves do not fit root_cap %u", (unsigned long long)root_count, root_cap);
        for (u64 i = 0; i < root_count; i++) per_root[i] = host[MW_TABLE + i];
    }
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_perft_batch_multi(ll_mctx *g, const ll_world_t *roots, uint32_t n, int depth, uint64_t *leaves) {
    LL_ABI_BEGIN
    LL_TRY(check_group(g));
    if (n && (!roots || !leaves)) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (n > 1024) return fail(LL_ERR_CAPACITY, "at most 1024 roots a call");
    if (depth < 0 || depth > 32) return fail(LL_ERR_INVALID_ARG, "depth %d out of range", depth);
    if (!n) return LL_OK;
    Job job{[&](Member &m) { return member_perft_batch(m, roots, n, depth); }};
    LL_TRY(run_on_all(g, job));
    memcpy(leaves, g->members[0]->host_words, (size_t)n * sizeof(u64));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_kickoff_multi(ll_mctx *g, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, ll_commit_t *roots, float *scores,
                                uint32_t root_cap, uint32_t *n_roots, uint64_t *leaves_scored) {
    LL_ABI_BEGIN
    LL_TRY(check_group(g));
    if (!root || !scores || !n_roots) return fail(LL_ERR_INVALID_ARG, "null argument");
    *n_roots = 0;
    if (leaves_scored) *leaves_scored = 0;
    // every member searches into its own buffers; member 0's are the caller's
    std::vector<std::vector<float>> spare(g->members.size());
    std::vector<KickoffOut> outs(g->members.size());
    for (size_t i = 0; i < g->members.size(); i++) {
        if (i) spare[i].resize(root_cap);
        outs[i] = KickoffOut{i ? nullptr : roots, i ? spare[i].data() : scores, root_cap, 0, 0};
    }
    Job job{[&](Member &m) {
        size_t i = 0;
        while (g->members[i] != &m) i++;
        return member_kickoff(m, root, depth, quiet_extension, nihilistically, outs[i]);
    }};
    LL_TRY(run_on_all(g, job));
    *n_roots = outs[0].n_roots;
    if (leaves_scored) *leaves_scored = outs[0].leaves_scored;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_alpha_beta_forecast_multi(ll_mctx *g, const ll_world_t *root, int depth, int device_plies, int quiet_extension, int nihilistically,
                                            ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts, uint64_t *stats3) {
    LL_ABI_BEGIN
    LL_TRY(check_group(g));
    if (!root || !forecasts || !n_forecasts) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (depth < 1 || depth > 32) return fail(LL_ERR_INVALID_ARG, "depth %d out of range (1..32)", depth);
    if (device_plies < 0 || device_plies > 8) return fail(LL_ERR_INVALID_ARG, "device_plies %d out of range (0..8)", device_plies);
    if (quiet_extension > 8) return fail(LL_ERR_INVALID_ARG, "quiet extension %d out of range", quiet_extension);
    *n_forecasts = 0;
    std::vector<std::vector<ll_forecast_t>> spare(g->members.size());
    std::vector<uint32_t> counts(g->members.size(), 0);
    std::vector<std::array<uint64_t, 3>> stats(g->members.size());
    for (size_t i = 1; i < g->members.size(); i++) spare[i].resize(cap);
    Job job{[&](Member &m) {
        size_t i = 0;
        while (g->members[i] != &m) i++;
        return member_alpha_beta(m, root, depth, device_plies, quiet_extension, nihilistically, i ? spare[i].data() : forecasts, cap, &counts[i], stats[i].data());
    }};
    LL_TRY(run_on_all(g, job));
    *n_forecasts = counts[0];
    if (stats3) memcpy(stats3, stats[0].data(), 3 * sizeof(uint64_t));
    return LL_OK;
    LL_ABI_END
}
