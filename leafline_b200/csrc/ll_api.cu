// ll_api.cu -- the C ABI (include/leafline_b200.h) over the CUDA kernels.  No torch types, no CPU
// fallback: every entry point runs on the device or fails with an LL_ERR_* code.
#include "../../include/leafline_b200.h"
#include "ll_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <list>
#include <mutex>
#include <thread>
#include <new>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <unordered_map>
#include <vector>

using namespace ll;

#ifndef LL_TAIL_TILE
#define LL_TAIL_TILE 128 // parents per tile of the perft tail in the launch chain
#endif
#ifndef LL_NARROW_TILE_NODES
#define LL_NARROW_TILE_NODES 0 // levels at most this wide (a rank's share of a dealt tree) are expanded with 32-parent tiles
#endif
#ifndef LL_SMALL_LEVEL_NODES
#define LL_SMALL_LEVEL_NODES 0 // levels at most this wide are expanded by one warp per parent whatever their ply
#endif
#ifndef LL_SMALL_TILE_PLIES
#define LL_SMALL_TILE_PLIES 4 // plies of the perft launch chain expanded with 32-parent tiles
#endif

static thread_local char g_error[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(e_ == cudaErrorMemoryAllocation ? LL_ERR_OOM : LL_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                        \
    } while (0)
#define LL_TRY(expr)              \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != LL_OK) return rc_; \
    } while (0)

// Nothing unwinds across the C boundary (include/leafline_b200.h): every entry point that can allocate on the host runs
// its body inside this guard.
#define LL_ABI_BEGIN try {
#define LL_ABI_END                                                                                   \
    }                                                                                                \
    catch (const std::bad_alloc &) { return fail(LL_ERR_OOM, "host allocation failed"); }            \
    catch (const std::length_error &e) { return fail(LL_ERR_INVALID_ARG, "%s", e.what()); }          \
    catch (const std::exception &e) { return fail(LL_ERR_CUDA, "unexpected exception: %s", e.what()); } \
    catch (...) { return fail(LL_ERR_CUDA, "unexpected exception"); }

namespace {

struct DevBuf { // grow-only device allocation
    void *ptr = nullptr;
    size_t cap = 0;
};

struct Frontier { // one BFS level resident in HBM
    DevBuf planes, meta, root, counts, block_sums, block_base, cmeta, values, child_first, tag;
    size_t cap = 0; // positions
    size_t n = 0;
    bool borrowed = false; // planes / meta belong to the caller (ll_horizon_soa): never grown or freed here
    Soa soa() const { return Soa{(u64 *)planes.ptr, (u32 *)meta.ptr, cap}; }
};

struct TimedLaunch {
    std::string name;
    cudaEvent_t start, stop;
};

} // namespace

struct ll_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr; // uploads of the next chunk while the current one is being scored (ll_score_batch)
    cudaEvent_t slot_uploaded[2] = {nullptr, nullptr}, slot_consumed[2] = {nullptr, nullptr};
    std::vector<Frontier> levels;
    Frontier ext; // scratch (counts / scan / values) used when level 0 aliases caller-owned planes
    Frontier unit_scratch; // the whole shard level before this rank's units are picked out of it (level-by-level paths)
    DevBuf aos_in, aos_out, scalars, scores, teams, flags;
    void *pinned = nullptr; // small host staging
    size_t pinned_cap = 0;
    bool timing = false;
    int search_threads = 32;   // searcher threads of the host alpha-beta driver (one per root move up to this many; ll_set_search_threads)
    double deja_vu_gib = 0.25; // byte budget of the search drivers' memory bank (mind.rs:367-372; ll_set_deja_vu_bound)
    bool literal = false;      // this call holds a position outside domain W: the literal path answers it (ll_device.cuh: DOMAIN_LITERAL)
    bool force_synced = false; // tests: always take the level-by-level path
    bool head_cluster = false; // the chain's first plies go into one cluster launch (k_expand_head)
    int head_ctas = HEAD_CTAS, head_max_steps = HEAD_MAX_STEPS; // (experiment knobs: LEAFLINE_B200_HEAD_CLUSTER=1, LEAFLINE_B200_HEAD=ctas,steps)
    bool no_graph = false;     // tests: issue the perft launch chain launch by launch instead of replaying its CUDA graph
    uint64_t chain_sig = 0;    // the captured perft chain (same plan, same buffers, same stream -> replay)
    cudaGraphExec_t chain_exec = nullptr;
    cudaStream_t chain_stream = nullptr;
    std::vector<TimedLaunch> timed;
    uint64_t launches = 0;
};

namespace {

size_t nblocks(size_t n, size_t per = BLOCK) { return (n + per - 1) / per; }

// Grow-only: on failure the buffer is EMPTY (ptr == nullptr, cap == 0), never stale.  The new block is allocated before the
// old one is released when both fit; a level that needs most of HBM falls back to release-then-allocate.
int ensure(ll_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return LL_OK;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->chain_exec) { // the captured perft chain holds the old addresses
        cudaGraphExecDestroy(ctx->chain_exec);
        ctx->chain_exec = nullptr;
    }
    void *fresh = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&fresh, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (b.ptr) cudaFree(b.ptr);
        b.ptr = nullptr;
        b.cap = 0;
        e = cudaMalloc(&fresh, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&fresh, want);
        }
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LL_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = fresh;
    b.cap = want;
    return LL_OK;
}

void drop_buf(DevBuf &b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
}
// a level whose growth failed half-way is emptied entirely: the next call re-reserves it from scratch (LL_ERR_OOM is
// recoverable: the context stays usable)
void drop_level(Frontier &f) {
    drop_buf(f.planes); drop_buf(f.meta); drop_buf(f.root); drop_buf(f.counts); drop_buf(f.block_sums); drop_buf(f.block_base);
    drop_buf(f.cmeta); drop_buf(f.values); drop_buf(f.child_first); drop_buf(f.tag);
    f.cap = 0;
    f.n = 0;
}

// make level `l` able to hold `n` positions (contents are NOT preserved when it grows)
int reserve_level(ll_ctx *ctx, size_t l, size_t n, bool with_cmeta, bool with_values) {
    if (ctx->levels.size() <= l) ctx->levels.resize(l + 1);
    Frontier &f = ctx->levels[l];
    if (f.borrowed) { // level 0 aliasing caller-owned planes (ll_horizon_soa): never reallocated
        if (n > f.cap) return fail(LL_ERR_INVALID_ARG, "%zu positions do not fit the caller's stride %zu", n, f.cap);
    } else if (n > f.cap) {
        const size_t cap = ((n + n / 8 + 63) / 64) * 64; // even stride, 16-byte aligned planes
        int rc = ensure(ctx, f.planes, cap * 12 * sizeof(u64));
        if (rc == LL_OK) rc = ensure(ctx, f.meta, cap * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, f.root, cap * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, f.counts, cap * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, f.block_sums, (nblocks(cap) + 1) * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, f.block_base, (nblocks(cap) + 1) * sizeof(u64));
        if (rc != LL_OK) {
            drop_level(f);
            return rc;
        }
        f.cap = cap;
    }
    int rc = LL_OK;
    if (with_cmeta) rc = ensure(ctx, f.cmeta, f.cap * sizeof(u32));
    if (rc == LL_OK && with_values) rc = ensure(ctx, f.values, f.cap * sizeof(float));
    if (rc != LL_OK && !f.borrowed) drop_level(f);
    return rc;
}

struct Timed { // scope guard: counts the launch and, when timing is on, brackets it with CUDA events on the launching stream
    ll_ctx *ctx;
    int idx = -1;
    Timed(ll_ctx *c, const char *name) : ctx(c) {
        ctx->launches++;
        if (!ctx->timing) return;
        TimedLaunch t;
        t.name = name;
        if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) return;
        cudaEventRecord(t.start, ctx->stream);
        ctx->timed.push_back(t);
        idx = (int)ctx->timed.size() - 1;
    }
    ~Timed() {
        if (idx >= 0) cudaEventRecord(ctx->timed[idx].stop, ctx->stream);
    }
};

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LL_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return LL_OK;
}

int pinned(ll_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return LL_OK;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
    CUDA_TRY(cudaMallocHost(&ctx->pinned, bytes + 4096));
    ctx->pinned_cap = bytes + 4096;
    return LL_OK;
}

// scalars buffer layout (u64 slots): 4, 5 = first bad index and literal-path flag, 1 = scan total, 2 = leaf total, 3 = leaves scored,
// 8.. = per-root counts
constexpr size_t SC_BAD = 4 /* and 5: needs-the-literal-path flag, see validate_level */, SC_TOTAL = 1, SC_LEAVES = 2, SC_SCORED = 3, SC_PER_ROOT = 8, SC_SPARE = 8 + 1024,
                 SC_N = 1040 /* 64 device-resident level sizes */, SC_TILEQ = 1104 /* 64 u32 tile queues */, SC_OVERFLOW = 1136, SC_HZ_OVERFLOW = 1137, SC_SLOTS = 1152;

int upload_worlds(ll_ctx *ctx, const ll_world_t *worlds, size_t n, size_t level) {
    LL_TRY(reserve_level(ctx, level, n, false, false));
    LL_TRY(ensure(ctx, ctx->aos_in, n * sizeof(ll_world_t)));
    CUDA_TRY(cudaMemcpyAsync(ctx->aos_in.ptr, worlds, n * sizeof(ll_world_t), cudaMemcpyHostToDevice, ctx->stream));
    Frontier &f = ctx->levels[level];
    f.n = n;
    if (n) {
        Timed t(ctx, "aos_to_soa");
        k_aos_to_soa<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>((const u64 *)ctx->aos_in.ptr, n, f.soa(), 0);
    }
    return check_launch("k_aos_to_soa");
}

// level 0 := one root position, passed as a kernel argument
int upload_root(ll_ctx *ctx, const ll_world_t *root) {
    static_assert(sizeof(WorldWords) == sizeof(ll_world_t), "ll_world_t is 13 words");
    LL_TRY(reserve_level(ctx, 0, 1, false, false));
    Frontier &f = ctx->levels[0];
    f.n = 1;
    WorldWords words;
    memcpy(&words, root, sizeof words);
    {
        Timed t(ctx, "store_world");
        k_store_world<<<1, 32, 0, ctx->stream>>>(words, f.soa(), 0);
    }
    return check_launch("k_store_world");
}

// classify n resident positions: LL_ERR_DOMAIN for one outside every domain; ctx->literal if one needs the literal path
int validate_level(ll_ctx *ctx, const Soa &soa, size_t n) {
    if (!n) return LL_OK;
    u64 *sc = (u64 *)ctx->scalars.ptr;
    const u64 init[2] = {~0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(sc + SC_BAD, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timed t(ctx, "validate");
        k_validate<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>(soa, n, (unsigned long long *)(sc + SC_BAD));
    }
    LL_TRY(check_launch("k_validate"));
    u64 got[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(got, sc + SC_BAD, sizeof got, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (got[0] != ~0ull)
        return fail(LL_ERR_DOMAIN, "position %llu cannot be answered: overlapping planes, bytes out of range or an inconsistent passing-by square (see include/leafline_b200.h)",
                    (unsigned long long)got[0]);
    if (got[1]) ctx->literal = true;
    return LL_OK;
}

// count the commits of every node of level l (counts + block bases stay on the device); returns the total
int count_level(ll_ctx *ctx, size_t l, bool nihil, u64 *total_out, bool stuns_only = false) {
    Frontier &f = ctx->levels[l];
    *total_out = 0;
    if (!f.n) return LL_OK;
    const size_t nb = nblocks(f.n);
    u64 *sc = (u64 *)ctx->scalars.ptr;
    {
        Timed t(ctx, nihil ? "count_reckless" : "count_legal");
        if (nihil) k_count<true><<<(unsigned)nb, BLOCK, 0, ctx->stream>>>(f.soa(), f.n, (u32 *)f.counts.ptr, (u32 *)f.block_sums.ptr, stuns_only, ctx->literal);
        else k_count<false><<<(unsigned)nb, BLOCK, 0, ctx->stream>>>(f.soa(), f.n, (u32 *)f.counts.ptr, (u32 *)f.block_sums.ptr, stuns_only, ctx->literal);
    }
    LL_TRY(check_launch("k_count"));
    {
        Timed t(ctx, "scan_block_sums");
        k_scan_block_sums<<<1, 1024, 0, ctx->stream>>>((u32 *)f.block_sums.ptr, (u64 *)f.block_base.ptr, nb, sc + SC_TOTAL);
    }
    LL_TRY(check_launch("k_scan_block_sums"));
    CUDA_TRY(cudaMemcpyAsync(total_out, sc + SC_TOTAL, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
}

// expand level l into `g` (count_level(l) must have run).  root_mode: 0 none, 1 inherit, 2 own index, 3 parent's index.
// shard_world > 1: every child is tagged with the rank that owns it (g.tag; see unit_hash)
int emit_level_into(ll_ctx *ctx, size_t l, Frontier &g, bool nihil, u64 total, bool with_cmeta, int root_mode, bool stuns_only, int shard_world) {
    Frontier &f = ctx->levels[l];
    if ((size_t)total > g.cap || (with_cmeta && g.cmeta.cap < g.cap * sizeof(u32))) {
        // (g may be a scratch frontier outside ctx->levels: the same growth rule as reserve_level)
        const size_t cap = std::max<size_t>(g.cap, (((size_t)total + (size_t)total / 8 + 63) / 64) * 64);
        int rc = ensure(ctx, g.planes, cap * 12 * sizeof(u64));
        if (rc == LL_OK) rc = ensure(ctx, g.meta, cap * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, g.root, cap * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, g.counts, cap * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, g.block_sums, (nblocks(cap) + 1) * sizeof(u32));
        if (rc == LL_OK) rc = ensure(ctx, g.block_base, (nblocks(cap) + 1) * sizeof(u64));
        if (rc == LL_OK && with_cmeta) rc = ensure(ctx, g.cmeta, cap * sizeof(u32));
        if (rc != LL_OK) {
            drop_level(g);
            return rc;
        }
        g.cap = cap;
    }
    if (shard_world > 1) LL_TRY(ensure(ctx, g.tag, g.cap * sizeof(u32)));
    g.n = (size_t)total;
    if (!f.n || !total) return LL_OK;
    ExpandArgs a{};
    a.in = f.soa();
    a.n = f.n;
    a.counts = (const u32 *)f.counts.ptr;
    a.block_base = (const u64 *)f.block_base.ptr;
    a.root_in = root_mode == 1 ? (const u32 *)f.root.ptr : nullptr;
    a.out = g.soa();
    a.cmeta = with_cmeta ? (u32 *)g.cmeta.ptr : nullptr;
    a.root_out = root_mode == 1 || root_mode == 2 ? (u32 *)g.root.ptr : nullptr;
    a.parent_out = root_mode == 3 ? (u32 *)g.root.ptr : nullptr;
    a.root_is_index = root_mode == 2;
    if (shard_world > 1) {
        if (shard_world > 32) return fail(LL_ERR_INVALID_ARG, "at most 32 ranks share a tree");
        a.shard_world = shard_world;
        a.unit_owner = (u32 *)g.tag.ptr;
    }
    {
        Timed t(ctx, nihil ? "emit_reckless" : "emit_legal");
        if (stuns_only && !nihil) return fail(LL_ERR_INVALID_ARG, "quiescence levels are reckless");
        if (ctx->literal && !nihil) k_expand<false, MODE_EMIT, BLOCK, false, true><<<(unsigned)nblocks(f.n), BLOCK, 0, ctx->stream>>>(a); // (the reckless generator is literal as it is)
        else if (stuns_only) k_expand<true, MODE_EMIT, BLOCK, true><<<(unsigned)nblocks(f.n), BLOCK, 0, ctx->stream>>>(a);
        else if (nihil) k_expand<true, MODE_EMIT><<<(unsigned)nblocks(f.n), BLOCK, 0, ctx->stream>>>(a);
        else k_expand<false, MODE_EMIT><<<(unsigned)nblocks(f.n), BLOCK, 0, ctx->stream>>>(a);
    }
    return check_launch("k_expand<EMIT>");
}
// expand level l into level l+1
int emit_level(ll_ctx *ctx, size_t l, bool nihil, u64 total, bool with_cmeta, int root_mode, bool stuns_only = false) {
    LL_TRY(reserve_level(ctx, l + 1, (size_t)total, with_cmeta, false));
    return emit_level_into(ctx, l, ctx->levels[l + 1], nihil, total, with_cmeta, root_mode, stuns_only, 1);
}

int check_ctx(ll_ctx *ctx) {
    if (!ctx) return fail(LL_ERR_INVALID_ARG, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->literal = false; // every entry point classifies its own positions
    return LL_OK;
}
// a caller-owned ll_soa_t: device pointers present, n inside the stride, planes 16-byte aligned
int check_soa(const ll_soa_t *soa) {
    if (!soa) return fail(LL_ERR_INVALID_ARG, "null soa");
    if (soa->n > soa->stride) return fail(LL_ERR_INVALID_ARG, "soa->n %zu exceeds its stride %zu", soa->n, soa->stride);
    if (soa->n && (!soa->planes || !soa->meta)) return fail(LL_ERR_INVALID_ARG, "null planes / meta");
    if (soa->stride & 1) return fail(LL_ERR_INVALID_ARG, "stride must be even");
    return LL_OK;
}

// horizon values of level `l` nodes (n of them) into levels[l].values; uses levels l+1.. as scratch.
// The node at level l has `plies` plies of full-width reckless search left; with quiet_ext > 0 the search
// goes on below depth 0 for up to quiet_ext plies over the stunning commits only, every such node being
// free to stand pat (mind.rs:284-301).  quiet_ext <= 0: no extension (quiet = None, or Some(0)).
int horizon_levels(ll_ctx *ctx, size_t l, int plies, int quiet_ext, uint64_t *scored_total, bool with_cmeta = false, size_t *bottom_out = nullptr) {
    Frontier &f0 = ctx->levels[l];
    LL_TRY(reserve_level(ctx, l, f0.n, false, true));
    u64 *sc = (u64 *)ctx->scalars.ptr;
    if (!f0.n) return LL_OK;
    if (plies < 0) plies = 0;
    const int expanding = plies + (quiet_ext > 0 ? quiet_ext : 0); // levels whose children are generated
    if (bottom_out) *bottom_out = l;
    if (expanding == 0) {
        Timed t(ctx, "leaf_values");
        k_leaf_values<<<(unsigned)nblocks(f0.n), BLOCK, 0, ctx->stream>>>(f0.soa(), f0.n, (float *)f0.values.ptr);
        if (scored_total) *scored_total += f0.n;
        return check_launch("k_leaf_values");
    }
    // level l + j has plies - j plies left; at <= 0 it is a quiescence level.  The last expansion happens on chip
    // (k_horizon: children scored in registers) -- except on the literal path (a position outside domain W: more figurines than
    // the record lists hold), which materialises that level too and scores it like any other
    size_t bottom = l;
    for (int j = 0; j < (ctx->literal ? expanding : expanding - 1); j++) {
        const bool quiet = plies - j <= 0;
        u64 total = 0;
        LL_TRY(count_level(ctx, bottom, true, &total, quiet));
        LL_TRY(emit_level(ctx, bottom, true, total, with_cmeta, 0, quiet));
        bottom++;
    }
    Frontier &fb = ctx->levels[bottom];
    if (bottom_out) *bottom_out = ctx->literal ? bottom - 1 : bottom;
    LL_TRY(reserve_level(ctx, bottom, fb.n, false, true));
    CUDA_TRY(cudaMemsetAsync(sc + SC_SCORED, 0, sizeof(u64), ctx->stream));
    if (ctx->literal) {
        if (fb.n) {
            Timed t(ctx, "leaf_values");
            k_leaf_values<<<(unsigned)nblocks(fb.n), BLOCK, 0, ctx->stream>>>(fb.soa(), fb.n, (float *)fb.values.ptr);
            LL_TRY(check_launch("k_leaf_values"));
        }
        if (scored_total) *scored_total += fb.n;
    } else if (fb.n) {
        const bool quiet = plies - (expanding - 1) <= 0;
        ExpandArgs a{};
        a.in = fb.soa();
        a.n = fb.n;
        a.values = (float *)fb.values.ptr;
        a.scored = sc + SC_SCORED;
        a.stand_pat = quiet;
        LL_TRY(ensure(ctx, fb.tag, fb.cap * sizeof(u32))); // the list of parents whose records outgrow a tile (rarely any)
        a.overflow_list = (u32 *)fb.tag.ptr;
        a.overflow_count = (u32 *)(sc + SC_HZ_OVERFLOW);
        CUDA_TRY(cudaMemsetAsync(a.overflow_count, 0, sizeof(u32), ctx->stream));
        Timed t(ctx, "horizon_last_ply");
        if (quiet) k_horizon<true><<<(unsigned)nblocks(fb.n, HP), BLOCK, 0, ctx->stream>>>(a);
        else k_horizon<false><<<(unsigned)nblocks(fb.n, HP), BLOCK, 0, ctx->stream>>>(a);
        LL_TRY(check_launch("k_horizon"));
        ctx->launches++;
        if (quiet) k_horizon_overflow<true><<<(unsigned)ctx->sm_count, BLOCK, 0, ctx->stream>>>(a);
        else k_horizon_overflow<false><<<(unsigned)ctx->sm_count, BLOCK, 0, ctx->stream>>>(a);
    }
    LL_TRY(check_launch("k_horizon"));
    for (size_t lev = bottom; lev > l; lev--) {
        Frontier &par = ctx->levels[lev - 1];
        Frontier &chi = ctx->levels[lev];
        LL_TRY(reserve_level(ctx, lev - 1, par.n, false, true));
        const bool quiet = plies - (int)(lev - 1 - l) <= 0;
        if (!par.n) continue; // the tree died out above this level
        Timed t(ctx, "backup");
        k_backup<<<(unsigned)nblocks(par.n), BLOCK, 0, ctx->stream>>>(par.soa(), par.n, (const u32 *)par.counts.ptr, (const u64 *)par.block_base.ptr,
                                                                     (const float *)chi.values.ptr, (float *)par.values.ptr, quiet);
        LL_TRY(check_launch("k_backup"));
    }
    if (scored_total) {
        u64 s = 0;
        CUDA_TRY(cudaMemcpyAsync(&s, sc + SC_SCORED, sizeof s, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        *scored_total += s;
    }
    return LL_OK;
}

// The same values for a SMALL batch without a host round trip per level: every level is expanded order-free into the
// capacity it already has (sizes stay on the device, tiles claim their output ranges, every parent records where its
// children went), the bottom level by k_horizon<CHAIN>, the back-ups by range.  One launch chain, no synchronisation;
// the caller reads the overflow flag together with the values and, if a level outgrew its buffer, redoes the call
// level by level (which also grows the buffers).  plies >= 2, no quiescence extension.
constexpr size_t HORIZON_CHAIN_MAX_BATCH = 4096;
constexpr size_t HORIZON_CHAIN_MIN_CAP = 1u << 20; // positions per scratch level (100 MB), so that small batches fit from the start
int horizon_chain(ll_ctx *ctx, size_t l0, int plies, bool *launched) {
    *launched = false;
    const size_t bottom = l0 + (size_t)plies - 1;
    if (bottom + 2 > 60) return LL_OK;
    const size_t n = ctx->levels[l0].n;
    for (size_t j = l0 + 1; j <= bottom; j++) LL_TRY(reserve_level(ctx, j, HORIZON_CHAIN_MIN_CAP, false, true));
    LL_TRY(reserve_level(ctx, l0, n, false, true));
    for (size_t j = l0; j < bottom; j++) LL_TRY(ensure(ctx, ctx->levels[j].child_first, ctx->levels[j].cap * sizeof(u64)));
    u64 *sc = (u64 *)ctx->scalars.ptr;
    u64 *N = sc + SC_N;
    u32 *tileq = (u32 *)(sc + SC_TILEQ);
    u32 *overflow = (u32 *)(sc + SC_OVERFLOW);
    const unsigned resident = (unsigned)ctx->sm_count * LL_EXPAND_MIN_BLOCKS;
    CUDA_TRY(cudaMemsetAsync(sc + SC_LEAVES, 0, (SC_SLOTS - SC_LEAVES) * sizeof(u64), ctx->stream));
    int q = 0;
    for (size_t j = l0; j < bottom; j++) {
        Frontier &f = ctx->levels[j];
        Frontier &g = ctx->levels[j + 1];
        ExpandArgs a{};
        a.in = f.soa();
        a.n = n;
        a.n_in = j == l0 ? nullptr : N + j;
        a.tile_counter = tileq + q++;
        a.out = g.soa();
        a.n_out = N + j + 1;
        a.out_cap = g.cap;
        a.overflow = overflow;
        a.child_first = (u64 *)f.child_first.ptr;
        a.child_count = (u32 *)f.counts.ptr;
        const bool narrow = j < l0 + 2; // the first levels below a small batch: 32-parent tiles reach more SMs
        const unsigned grid = (unsigned)std::min<size_t>(nblocks(j == l0 ? n : f.cap, narrow ? 32 : BLOCK), resident);
        Timed t(ctx, "emit_reckless");
        if (narrow) k_expand<true, MODE_EMIT, 32><<<grid, BLOCK, 0, ctx->stream>>>(a);
        else k_expand<true, MODE_EMIT><<<grid, BLOCK, 0, ctx->stream>>>(a);
        LL_TRY(check_launch("k_expand<EMIT> (chain)"));
        g.n = 0; // its size lives on the device
    }
    {
        Frontier &fb = ctx->levels[bottom];
        ExpandArgs a{};
        a.in = fb.soa();
        a.n_in = N + bottom;
        a.tile_counter = tileq + q++;
        a.overflow = overflow;
        a.values = (float *)fb.values.ptr;
        a.scored = sc + SC_SCORED;
        LL_TRY(ensure(ctx, fb.tag, fb.cap * sizeof(u32)));
        a.overflow_list = (u32 *)fb.tag.ptr;
        a.overflow_count = (u32 *)(sc + SC_HZ_OVERFLOW); // (zeroed by the memset at the head of the chain)
        Timed t(ctx, "horizon_last_ply");
        k_horizon<false, true><<<(unsigned)std::min<size_t>(nblocks(fb.cap, HP), resident), BLOCK, 0, ctx->stream>>>(a);
        LL_TRY(check_launch("k_horizon<CHAIN>"));
        ctx->launches++;
        k_horizon_overflow<false><<<(unsigned)ctx->sm_count, BLOCK, 0, ctx->stream>>>(a);
        LL_TRY(check_launch("k_horizon_overflow"));
    }
    for (size_t lev = bottom; lev > l0; lev--) {
        Frontier &par = ctx->levels[lev - 1];
        Frontier &chi = ctx->levels[lev];
        Timed t(ctx, "backup");
        k_backup_ranges<<<(unsigned)std::min<size_t>(nblocks(lev == l0 + 1 ? n : par.cap), resident), BLOCK, 0, ctx->stream>>>(
            par.soa(), n, lev == l0 + 1 ? nullptr : N + lev - 1, overflow, (const u64 *)par.child_first.ptr, (const u32 *)par.counts.ptr,
            (const float *)chi.values.ptr, (float *)par.values.ptr);
        LL_TRY(check_launch("k_backup_ranges"));
    }
    *launched = true;
    return LL_OK;
}

} // namespace

// ================================================================================================ ABI

extern "C" const char *ll_last_error(void) { return g_error; }

static int init_ctx(ll_ctx *ctx) {
    ctx->levels.reserve(64); // references to levels stay valid while deeper levels are added
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->slot_uploaded[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->slot_consumed[i], cudaEventDisableTiming));
    }
    ctx->stream = ctx->own_stream;
    // generated tables, furniture/tablemaker.py:41-62
    u64 pony[64], fig[64];
    static const int pony_opt[8][2] = {{1, 2}, {-1, 2}, {1, -2}, {-1, -2}, {2, 1}, {-2, 1}, {2, -1}, {-2, -1}};
    for (int r = 0; r < 8; r++)
        for (int f = 0; f < 8; f++) {
            u64 a = 0, b = 0;
            for (int k = 0; k < 8; k++) {
                int rr = r + pony_opt[k][0], ff = f + pony_opt[k][1];
                if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) a |= 1ull << (8 * rr + ff);
            }
            for (int i = -1; i <= 1; i++)
                for (int j = -1; j <= 1; j++) {
                    if (!i && !j) continue;
                    int rr = r + i, ff = f + j;
                    if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) b |= 1ull << (8 * rr + ff);
                }
            pony[8 * r + f] = a;
            fig[8 * r + f] = b;
        }
    CUDA_TRY(cudaMemcpyToSymbol(d_pony_table, pony, sizeof pony));
    CUDA_TRY(cudaMemcpyToSymbol(d_figurehead_table, fig, sizeof fig));
    u64 lines[64 * 4]; // StagedLines: diagonal, anti-diagonal, file and rank through every square, the square excluded
    for (int sq = 0; sq < 64; sq++) {
        const int r = sq >> 3, f = sq & 7;
        u64 d = 0, a = 0, fl = 0, rk = 0;
        for (int t = 0; t < 64; t++) {
            const int tr = t >> 3, tf = t & 7;
            if (t == sq) continue;
            if (tf - tr == f - r) d |= 1ull << t;
            if (tf + tr == f + r) a |= 1ull << t;
            if (tf == f) fl |= 1ull << t;
            if (tr == r) rk |= 1ull << t;
        }
        lines[4 * sq] = d;
        lines[4 * sq + 1] = a;
        lines[4 * sq + 2] = fl;
        lines[4 * sq + 3] = rk;
    }
    CUDA_TRY(cudaMemcpyToSymbol(d_line_table, lines, sizeof lines));
    LL_TRY(ensure(ctx, ctx->scalars, SC_SLOTS * sizeof(u64)));
    { // can the head of the perft chain run as one 16-CTA cluster here?  (non-portable cluster size: opt in, then ask)
        // MEASURED SLOWER (perft(6) 0.357 ms against 0.343 ms with a launch per ply): off unless LEAFLINE_B200_HEAD_CLUSTER=1
        const char *on = getenv("LEAFLINE_B200_HEAD_CLUSTER");
        ctx->head_cluster = false;
        if (const char *knobs = getenv("LEAFLINE_B200_HEAD")) {
            int c = 0, st = 0;
            if (sscanf(knobs, "%d,%d", &c, &st) == 2 && c >= 1 && c <= HEAD_CTAS && st >= 2 && st <= HEAD_MAX_STEPS) {
                ctx->head_ctas = c;
                ctx->head_max_steps = st;
            }
        }
        if (on && *on == '1' && cudaFuncSetAttribute(k_expand_head, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3((unsigned)ctx->head_ctas, 1, 1);
            cfg.blockDim = dim3(HEAD_WARPS * 32, 1, 1);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)ctx->head_ctas;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&clusters, k_expand_head, &cfg) == cudaSuccess && clusters >= 1) ctx->head_cluster = true;
        }
        cudaGetLastError();
    }
    return LL_OK;
}

extern "C" int ll_init(int device, ll_ctx **out) {
    LL_ABI_BEGIN
    if (!out) return fail(LL_ERR_INVALID_ARG, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(LL_ERR_NO_DEVICE, "no CUDA device available (%s); leafline_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return fail(LL_ERR_INVALID_ARG, "device %d out of range (0..%d)", device, count - 1);
    CUDA_TRY(cudaSetDevice(device));
    ll_ctx *ctx = new ll_ctx();
    ctx->device = device;
    const int rc = init_ctx(ctx);
    if (rc != LL_OK) { // release whatever was created (streams, events, buffers); g_error keeps the cause
        char keep[sizeof g_error];
        memcpy(keep, g_error, sizeof keep);
        ll_shutdown(ctx);
        memcpy(g_error, keep, sizeof keep);
        return rc;
    }
    *out = ctx;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_shutdown(ll_ctx *ctx) {
    LL_ABI_BEGIN
    if (!ctx) return LL_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    auto drop = [](DevBuf &b) { drop_buf(b); };
    drop(ctx->ext.counts); drop(ctx->ext.block_sums); drop(ctx->ext.block_base); drop(ctx->ext.values); drop(ctx->ext.tag);
    for (Frontier &f : ctx->levels)
        if (!f.borrowed) drop_level(f);
    drop_level(ctx->unit_scratch);
    drop(ctx->aos_in); drop(ctx->aos_out); drop(ctx->scalars); drop(ctx->scores); drop(ctx->teams); drop(ctx->flags);
    for (TimedLaunch &t : ctx->timed) {
        cudaEventDestroy(t.start);
        cudaEventDestroy(t.stop);
    }
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->chain_exec) cudaGraphExecDestroy(ctx->chain_exec);
    for (int i = 0; i < 2; i++) {
        if (ctx->slot_uploaded[i]) cudaEventDestroy(ctx->slot_uploaded[i]);
        if (ctx->slot_consumed[i]) cudaEventDestroy(ctx->slot_consumed[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_set_stream(ll_ctx *ctx, void *cuda_stream) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_synchronize(ll_ctx *ctx) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_device_sm_count(ll_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int ll_timing_enable(ll_ctx *ctx, int on) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    ctx->timing = on != 0;
    return LL_OK;
    LL_ABI_END
}
extern "C" int ll_timing_reset(ll_ctx *ctx) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (TimedLaunch &t : ctx->timed) {
        cudaEventDestroy(t.start);
        cudaEventDestroy(t.stop);
    }
    ctx->timed.clear();
    return LL_OK;
    LL_ABI_END
}
extern "C" int ll_timing_query(ll_ctx *ctx, const char *kernel, double *ms, uint64_t *launches) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!kernel) return fail(LL_ERR_INVALID_ARG, "null kernel name");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    double total = 0;
    uint64_t count = 0;
    for (TimedLaunch &t : ctx->timed)
        if (t.name == kernel) {
            float e = 0;
            CUDA_TRY(cudaEventElapsedTime(&e, t.start, t.stop));
            total += e;
            count++;
        }
    if (ms) *ms = total;
    if (launches) *launches = count;
    return LL_OK;
    LL_ABI_END
}
extern "C" uint64_t ll_launch_count(ll_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int ll_debug_force_synced(ll_ctx *ctx, int on) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    ctx->force_synced = on == 1;
    ctx->no_graph = on == 2;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_debug_selfcheck_soa(ll_ctx *ctx, const ll_soa_t *soa, uint64_t *result4) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!soa || !result4) return fail(LL_ERR_INVALID_ARG, "null argument");
    LL_TRY(check_soa(soa));
    u64 *sc = (u64 *)ctx->scalars.ptr + SC_SPARE;
    const u64 init[4] = {0, ~0ull, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(sc, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    if (soa->n) {
        Timed t(ctx, "selfcheck");
        k_selfcheck<<<(unsigned)nblocks(soa->n), BLOCK, 0, ctx->stream>>>(Soa{(u64 *)soa->planes, soa->meta, soa->stride}, soa->n, (unsigned long long *)sc);
    }
    LL_TRY(check_launch("k_selfcheck"));
    CUDA_TRY(cudaMemcpyAsync(result4, sc, 4 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_debug_tables(ll_ctx *ctx, uint64_t *pony64, uint64_t *figurehead64) {
    LL_TRY(check_ctx(ctx));
    if (!pony64 || !figurehead64) return fail(LL_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(pony64, d_pony_table, 64 * sizeof(u64)));
    CUDA_TRY(cudaMemcpyFromSymbol(figurehead64, d_figurehead_table, 64 * sizeof(u64)));
    return LL_OK;
}

extern "C" int ll_debug_horizon_overflowed(ll_ctx *ctx, uint32_t *parents) {
    LL_TRY(check_ctx(ctx));
    if (!parents) return fail(LL_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaMemcpyAsync(parents, (u64 *)ctx->scalars.ptr + SC_HZ_OVERFLOW, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
}

extern "C" int ll_measure_int_peak(ll_ctx *ctx, int kind, double *thread_ops_per_sec) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!thread_ops_per_sec || kind < 0 || kind > 1) return fail(LL_ERR_INVALID_ARG, "bad argument");
    const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
    u32 *out = (u32 *)((u64 *)ctx->scalars.ptr + SC_SPARE); // never written (see the kernel)
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(a, ctx->stream));
        ctx->launches++;
        if (kind == 0) k_int_peak<0><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
        else k_int_peak<1><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
        CUDA_TRY(cudaEventRecord(b, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (rep && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *thread_ops_per_sec = (double)blocks * threads * iters * 16.0 / (best * 1e-3);
    return LL_OK;
    LL_ABI_END
}

// ---- scoring -----------------------------------------------------------------------------------------

static int launch_score(ll_ctx *ctx, const Soa &soa, size_t n, float *d_out) {
    if (!n) return LL_OK;
    Timed t(ctx, "score");
    k_score<<<(unsigned)nblocks((n + 1) / 2, 256), 256, 0, ctx->stream>>>(soa, n, d_out);
    return check_launch("k_score");
}

extern "C" int ll_score_batch(ll_ctx *ctx, const ll_world_t *worlds, size_t n, float *scores) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (n && (!worlds || !scores)) return fail(LL_ERR_INVALID_ARG, "null buffer");
    if (!n) return LL_OK;
    constexpr size_t CHUNK = (size_t)1 << 20; // 104 MiB of ll_world_t per chunk
    if (n <= CHUNK) {
        LL_TRY(upload_worlds(ctx, worlds, n, 0));
        LL_TRY(ensure(ctx, ctx->scores, (n + 1) * sizeof(float)));
        LL_TRY(launch_score(ctx, ctx->levels[0].soa(), n, (float *)ctx->scores.ptr));
        CUDA_TRY(cudaMemcpyAsync(scores, ctx->scores.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return LL_OK;
    }
    // Large batches are PCIe-bound (104 B in, 4 B out per position): stream them in chunks through two staging slots so
    // that the upload of chunk c+1 (copy stream) overlaps the repack + score + result download of chunk c.
    LL_TRY(reserve_level(ctx, 0, 2 * CHUNK, false, false));
    LL_TRY(ensure(ctx, ctx->aos_in, 2 * CHUNK * sizeof(ll_world_t)));
    LL_TRY(ensure(ctx, ctx->scores, 2 * CHUNK * sizeof(float)));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const Soa level = ctx->levels[0].soa();
    for (size_t first = 0, c = 0; first < n; first += CHUNK, c++) {
        const size_t m = std::min(CHUNK, n - first), slot = c & 1;
        ll_world_t *staging = (ll_world_t *)ctx->aos_in.ptr + slot * CHUNK;
        if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->slot_consumed[slot], 0)); // the slot's previous chunk has been repacked
        CUDA_TRY(cudaMemcpyAsync(staging, worlds + first, m * sizeof(ll_world_t), cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_TRY(cudaEventRecord(ctx->slot_uploaded[slot], ctx->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->slot_uploaded[slot], 0));
        {
            Timed t(ctx, "aos_to_soa");
            k_aos_to_soa<<<(unsigned)nblocks(m), BLOCK, 0, ctx->stream>>>((const u64 *)staging, m, level, slot * CHUNK);
        }
        LL_TRY(check_launch("k_aos_to_soa"));
        CUDA_TRY(cudaEventRecord(ctx->slot_consumed[slot], ctx->stream));
        float *d_scores = (float *)ctx->scores.ptr + slot * CHUNK;
        LL_TRY(launch_score(ctx, Soa{level.pl + slot * CHUNK, level.meta + slot * CHUNK, level.stride}, m, d_scores));
        CUDA_TRY(cudaMemcpyAsync(scores + first, d_scores, m * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_score_soa(ll_ctx *ctx, const ll_soa_t *soa, float *d_scores) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!soa || !d_scores) return fail(LL_ERR_INVALID_ARG, "null argument");
    LL_TRY(check_soa(soa));
    return launch_score(ctx, Soa{(u64 *)soa->planes, soa->meta, soa->stride}, soa->n, d_scores);
    LL_ABI_END
}

// ---- movegen -----------------------------------------------------------------------------------------

static bool host_well_formed(const ll_world_t *w);
static int classify_root(ll_ctx *ctx, const ll_world_t *w, size_t index);

extern "C" int ll_lookahead_batch(ll_ctx *ctx, const ll_world_t *worlds, size_t n, int nihilistically, uint32_t *offsets,
                                  ll_commit_t *commits, size_t commits_cap) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!offsets || (n && !worlds)) return fail(LL_ERR_INVALID_ARG, "null buffer");
    offsets[0] = 0;
    if (!n) return LL_OK;
    if (n == 1) { // a search driver asking for one node's commits: checked on the host, handed over as a kernel argument
        LL_TRY(classify_root(ctx, worlds, 0));
        LL_TRY(upload_root(ctx, worlds));
    } else {
        LL_TRY(upload_worlds(ctx, worlds, n, 0));
        LL_TRY(validate_level(ctx, ctx->levels[0].soa(), n));
    }
    u64 total = 0;
    LL_TRY(count_level(ctx, 0, nihilistically != 0, &total));
    if (total > 0xFFFFFFFFull) return fail(LL_ERR_CAPACITY, "more than 2^32 commits in one batch");
    Frontier &f = ctx->levels[0];
    LL_TRY(ensure(ctx, ctx->flags, (n + 1) * sizeof(u32)));
    {
        Timed t(ctx, "offsets");
        k_offsets<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>((const u32 *)f.counts.ptr, (const u64 *)f.block_base.ptr, n, (u32 *)ctx->flags.ptr);
    }
    LL_TRY(check_launch("k_offsets"));
    CUDA_TRY(cudaMemcpyAsync(offsets, ctx->flags.ptr, (n + 1) * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    if (!commits) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return LL_OK;
    }
    if (total > commits_cap) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return fail(LL_ERR_CAPACITY, "%llu commits do not fit the capacity of %zu", (unsigned long long)total, commits_cap);
    }
    LL_TRY(emit_level(ctx, 0, nihilistically != 0, total, true, 0));
    if (total) {
        Frontier &g = ctx->levels[1];
        LL_TRY(ensure(ctx, ctx->aos_out, (size_t)total * sizeof(ll_commit_t)));
        {
            Timed t(ctx, "commits_to_aos");
            k_commits_to_aos<<<(unsigned)nblocks((size_t)total), BLOCK, 0, ctx->stream>>>(g.soa(), (const u32 *)g.cmeta.ptr, (size_t)total, (u64 *)ctx->aos_out.ptr);
        }
        LL_TRY(check_launch("k_commits_to_aos"));
        CUDA_TRY(cudaMemcpyAsync(commits, ctx->aos_out.ptr, (size_t)total * sizeof(ll_commit_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_in_critical_endangerment_batch(ll_ctx *ctx, const ll_world_t *worlds, const uint8_t *teams, size_t n, uint8_t *endangered) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (n && (!worlds || !teams || !endangered)) return fail(LL_ERR_INVALID_ARG, "null buffer");
    if (!n) return LL_OK;
    for (size_t i = 0; i < n; i++)
        if (teams[i] > 1) return fail(LL_ERR_INVALID_ARG, "teams[%zu] is not a team", i);
    LL_TRY(upload_worlds(ctx, worlds, n, 0));
    LL_TRY(validate_level(ctx, ctx->levels[0].soa(), n));
    LL_TRY(ensure(ctx, ctx->teams, 2 * n));
    unsigned char *d_teams = (unsigned char *)ctx->teams.ptr, *d_out = d_teams + n;
    CUDA_TRY(cudaMemcpyAsync(d_teams, teams, n, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timed t(ctx, "endangered");
        k_endangered<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>(ctx->levels[0].soa(), n, d_teams, d_out);
    }
    LL_TRY(check_launch("k_endangered"));
    CUDA_TRY(cudaMemcpyAsync(endangered, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

// ---- perft -------------------------------------------------------------------------------------------


// host-side classification of a single position -- the mirror of domain_class (ll_device.cuh); the batch entry points
// classify on the device
enum { HOST_FAST = 0, HOST_LITERAL = 1, HOST_REJECT = 2 };
static int host_domain_class(const ll_world_t *w) {
    uint64_t acc = 0;
    for (int j = 0; j < 12; j++) {
        if (acc & w->plane[j]) return HOST_REJECT;
        acc |= w->plane[j];
    }
    if (w->initiative > 1 || w->service_eligibility > 0xF) return HOST_REJECT;
    if (w->passing_by != 0xFF) {
        const unsigned r = w->passing_by >> 4, f = w->passing_by & 0xF;
        if (r > 7 || f > 7) return HOST_REJECT;
        const uint64_t bit = 1ull << (8 * r + f);
        if (acc & bit) return HOST_REJECT;
        if (w->initiative == 0) {
            if (r != 5 || !(w->plane[6] & (bit >> 8)) || (acc & (bit << 8))) return HOST_REJECT;
        } else {
            if (r != 2 || !(w->plane[0] & (bit << 8)) || (acc & (bit >> 8))) return HOST_REJECT;
        }
    }
    if (__builtin_popcountll(w->plane[5]) > 1 || __builtin_popcountll(w->plane[11]) > 1) return HOST_LITERAL;
    uint64_t orange = 0, blue = 0;
    for (int j = 0; j < 6; j++) orange |= w->plane[j], blue |= w->plane[6 + j];
    const int n_orange = __builtin_popcountll(orange), n_blue = __builtin_popcountll(blue);
    if (n_orange > 16 || n_blue > 16) return HOST_LITERAL;
    const unsigned elig = w->service_eligibility;
    if ((elig & 0x3) && (w->plane[5] & ~(1ull << 4))) return HOST_LITERAL;
    if ((elig & 0xC) && (w->plane[11] & ~(1ull << 60))) return HOST_LITERAL;
    if (n_orange == 16 && (((elig & 0x1) && !(w->plane[3] & (1ull << 0))) || ((elig & 0x2) && !(w->plane[3] & (1ull << 7))))) return HOST_LITERAL;
    if (n_blue == 16 && (((elig & 0x4) && !(w->plane[9] & (1ull << 56))) || ((elig & 0x8) && !(w->plane[9] & (1ull << 63))))) return HOST_LITERAL;
    return HOST_FAST;
}
// classify one position handed over by the caller: LL_ERR_DOMAIN, or ctx->literal set when the literal path must answer
static int classify_root(ll_ctx *ctx, const ll_world_t *w, size_t index) {
    const int cls = host_domain_class(w);
    if (cls == HOST_REJECT)
        return fail(LL_ERR_DOMAIN, "position %zu cannot be answered: overlapping planes, bytes out of range or an inconsistent passing-by square (see include/leafline_b200.h)", index);
    if (cls == HOST_LITERAL) ctx->literal = true;
    return LL_OK;
}
static bool host_well_formed(const ll_world_t *w) { return host_domain_class(w) == HOST_FAST; } // (the multi-GPU layer deals only domain-W trees)

// Frontier sharding: with shard_world > 1 the work units are the nodes of ply `unit_ply`, dealt to the
// ranks by unit_hash (ll_device.cuh) in the expansion that creates them; 0 = the tree is too shallow to deal (rank 0
// counts all of it).  Both perft paths and every rank apply the same rule, so the ranks' counts always add up.
static int perft_unit_ply(int depth) {
    const int tail_ply = depth >= 3 ? depth - 2 : depth - 1;
    const int s = tail_ply < 3 ? tail_ply : 3;
    return s >= 2 ? s : 0;
}

// The chained path: the same levels, but every level's size stays on the device (ExpandArgs::n_in / n_out),
// the kernels are persistent tile loops and children are placed by an atomic claim per tile (perft needs no
// order) -- one launch chain, one synchronisation.  It runs inside the capacities the synced path has grown;
// *done = false (nothing returned) when a level would not fit.
static int perft_chain_launch(ll_ctx *ctx, int depth, int shard_rank, int shard_world, bool *launched) {
    *launched = false;
    const int tail_ply = depth - 2;
    const int unit_ply = shard_world > 1 ? perft_unit_ply(depth) : 0;
    if (shard_world > 1 && unit_ply == 0) return LL_OK; // shallow trees are not dealt: the level-by-level path answers
    const size_t need = (size_t)tail_ply + 1;
    if (ctx->levels.size() < need) return LL_OK;
    for (size_t l = 0; l < need; l++) {
        const Frontier &f = ctx->levels[l];
        if (!f.cap || !f.planes.ptr || !f.meta.ptr || !f.root.ptr) return LL_OK; // not sized yet (or emptied by a failed growth)
    }
    u64 *sc = (u64 *)ctx->scalars.ptr;
    u64 *N = sc + SC_N;
    u32 *tileq = (u32 *)(sc + SC_TILEQ);
    u32 *overflow = (u32 *)(sc + SC_OVERFLOW);
    const unsigned resident = (unsigned)ctx->sm_count * LL_EXPAND_MIN_BLOCKS;
    // the plan: one step per ply; each step reads its level's size from the device
    struct Plan { // zero-filled (padding included): its bytes are the signature of the captured graph
        ExpandArgs steps[CHAIN_MAX_STEPS];
        int kinds[CHAIN_MAX_STEPS];
        unsigned grids[CHAIN_MAX_STEPS];
        int n_steps;
        int head_steps; // the first plies fused into one cluster launch (k_expand_head); 0: every ply its own launch
    };
    static_assert(std::is_trivially_copyable<Plan>::value, "the plan is hashed byte-wise");
    Plan plan;
    memset(&plan, 0, sizeof plan);
    ExpandArgs *steps = plan.steps;
    int *kinds = plan.kinds;
    unsigned *grids = plan.grids;
    int q = 0, n_steps = 0;
    if (tail_ply + 1 > CHAIN_MAX_STEPS) return LL_OK; // deeper than the plan holds: the level-by-level path takes it
    for (int ply = 0; ply <= tail_ply; ply++) { // level index == ply
        Frontier &f = ctx->levels[ply];
        ExpandArgs &a = steps[n_steps];
        a.in = f.soa();
        a.n = 1;
        a.n_in = ply == 0 ? nullptr : N + ply;
        a.tile_counter = tileq + q++;
        a.root_in = ply == 0 ? nullptr : (const u32 *)f.root.ptr;
        a.overflow = overflow;
        grids[n_steps] = ply == 0 ? 1u : (unsigned)std::min<size_t>(nblocks(f.cap, ply == tail_ply ? LL_TAIL_TILE : BLOCK), resident);
        if (ply == tail_ply) {
            a.per_root = sc + SC_PER_ROOT;
            a.total = sc + SC_LEAVES;
            kinds[n_steps++] = STEP_TAIL;
            break;
        }
        Frontier &g = ctx->levels[ply + 1];
        a.out = g.soa();
        a.root_out = (u32 *)g.root.ptr;
        a.root_is_index = ply == 0;
        a.n_out = N + ply + 1;
        a.out_cap = g.cap;
        if (ply + 1 == unit_ply) { // this expansion creates the work units: every rank expands all parents, keeps its own units
            a.shard_rank = shard_rank;
            a.shard_world = shard_world;
        }
        // the first plies are narrow (1, ~30, ~900, ~27 000 nodes): one warp per parent reaches more SMs and finishes sooner
        // (so are the deeper levels of a rank's share of a dealt tree: ~25 000 nodes are 200 tiles of 128, a fifth of the blocks
        // the machine holds, each waiting on one thread's walk through the generator)
        const bool small = ply < LL_SMALL_TILE_PLIES || f.cap <= (size_t)LL_SMALL_LEVEL_NODES;
        const bool narrow_tiles = !small && f.cap <= (size_t)LL_NARROW_TILE_NODES; // 32-parent tiles: four times the blocks at work
        if (small && ply != 0) grids[n_steps] = (unsigned)std::min<size_t>(nblocks(f.cap, COOP_WARPS), resident);
        if (narrow_tiles) grids[n_steps] = (unsigned)std::min<size_t>(nblocks(f.cap, 32), resident);
        kinds[n_steps++] = small ? STEP_EMIT_SMALL : narrow_tiles ? STEP_EMIT_NARROW : STEP_EMIT;
    }
    plan.n_steps = n_steps;
    int head_steps = 0; // leading narrow plies (0, 1, 2) that go into the cluster launch
    while (ctx->head_cluster && head_steps < ctx->head_max_steps && head_steps < n_steps && kinds[head_steps] == STEP_EMIT_SMALL && head_steps < 3) head_steps++;
    if (head_steps < 2) head_steps = 0;
    plan.head_steps = head_steps;
    // (a single cooperative launch of the WHOLE chain with grid-wide barriers between the plies was measured: no faster -- DESIGN.md)
    auto issue = [&]() -> int {
        CUDA_TRY(cudaMemsetAsync(sc + SC_LEAVES, 0, (SC_SLOTS - SC_LEAVES) * sizeof(u64), ctx->stream));
        if (head_steps) {
            HeadArgs h;
            memset(&h, 0, sizeof h);
            for (int s = 0; s < head_steps; s++) h.step[s] = steps[s];
            h.n_steps = head_steps;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3((unsigned)ctx->head_ctas, 1, 1);
            cfg.blockDim = dim3(HEAD_WARPS * 32, 1, 1);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)ctx->head_ctas;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            Timed t(ctx, "emit_legal");
            CUDA_TRY(cudaLaunchKernelEx(&cfg, k_expand_head, h));
            LL_TRY(check_launch("k_expand_head"));
        }
        for (int s = head_steps; s < n_steps; s++) {
            const int kind = kinds[s];
            Timed t(ctx, kind == STEP_TAIL ? "perft_tail" : "emit_legal");
            // order-free steps expand from legal RECORDS (small code: these launches are bound by instruction fetch and by one
            // thread's serial latency); narrow levels by one warp per parent
            if (kind == STEP_TAIL) k_expand_records<MODE_PERFT, LL_TAIL_TILE><<<grids[s], BLOCK, 0, ctx->stream>>>(steps[s]);
            else if (kind == STEP_EMIT_SMALL) k_expand_coop<<<grids[s], BLOCK, 0, ctx->stream>>>(steps[s]);
            else if (kind == STEP_EMIT_NARROW) k_expand_records<MODE_EMIT, 32><<<grids[s], BLOCK, 0, ctx->stream>>>(steps[s]);
            else k_expand_records<MODE_EMIT><<<grids[s], BLOCK, 0, ctx->stream>>>(steps[s]);
            LL_TRY(check_launch("perft chain step"));
        }
        return LL_OK;
    };
    if (ctx->timing || ctx->no_graph) { // per-kernel events cannot be recorded into a replayed graph
        LL_TRY(issue());
        *launched = true;
        return LL_OK;
    }
    // the chain is the same launches with the same arguments call after call: capture it once as a CUDA graph and replay it
    uint64_t sig = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof plan; i++) sig = (sig ^ ((const unsigned char *)&plan)[i]) * 1099511628211ull;
    if (!ctx->chain_exec || ctx->chain_sig != sig || ctx->chain_stream != ctx->stream) {
        if (ctx->chain_exec) cudaGraphExecDestroy(ctx->chain_exec);
        ctx->chain_exec = nullptr;
        cudaGraph_t graph = nullptr;
        const uint64_t launches_before = ctx->launches;
        CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = issue();
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        ctx->launches = launches_before;
        if (rc != LL_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            LL_TRY(issue()); // capture unavailable: plain launches
            *launched = true;
            return LL_OK;
        }
        const cudaError_t ei = cudaGraphInstantiate(&ctx->chain_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            ctx->chain_exec = nullptr;
            cudaGetLastError();
            LL_TRY(issue());
            *launched = true;
            return LL_OK;
        }
        ctx->chain_sig = sig;
        ctx->chain_stream = ctx->stream;
    }
    CUDA_TRY(cudaGraphLaunch(ctx->chain_exec, ctx->stream));
    ctx->launches += (uint64_t)(n_steps - head_steps + (head_steps ? 1 : 0));
    *launched = true;
    return LL_OK;
}

static int perft_chained(ll_ctx *ctx, int depth, int shard_rank, int shard_world, uint64_t *leaves, uint64_t *per_root, uint32_t root_cap,
                         uint32_t *n_roots, bool *done) {
    *done = false;
    bool launched = false;
    LL_TRY(perft_chain_launch(ctx, depth, shard_rank, shard_world, &launched));
    if (!launched) return LL_OK;
    u64 *sc = (u64 *)ctx->scalars.ptr;
    LL_TRY(pinned(ctx, SC_SLOTS * sizeof(u64)));
    u64 *host = (u64 *)ctx->pinned;
    CUDA_TRY(cudaMemcpyAsync(host, sc, SC_SLOTS * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (*(u32 *)(host + SC_OVERFLOW)) return LL_OK; // a level outgrew its buffer: let the synced path size it
    const u64 root_count = host[SC_N + 1];
    if (root_count > 1024) return fail(LL_ERR_CAPACITY, "more than 1024 root moves");
    *leaves = host[SC_LEAVES];
    if (n_roots) *n_roots = (uint32_t)root_count;
    if (per_root) {
        if (root_count > root_cap) return fail(LL_ERR_CAPACITY, "%llu root moves do not fit root_cap %u", (unsigned long long)root_count, root_cap);
        for (u64 i = 0; i < root_count; i++) per_root[i] = host[SC_PER_ROOT + i];
    }
    *done = true;
    return LL_OK;
}

// the level-by-level path: every level's size comes back to the host, which sizes the next level exactly
static int perft_synced(ll_ctx *ctx, int depth, int shard_rank, int shard_world, uint64_t *leaves, uint64_t *per_root, uint32_t root_cap,
                        uint32_t *n_roots) {
    u64 *sc = (u64 *)ctx->scalars.ptr;
    CUDA_TRY(cudaMemsetAsync(sc + SC_LEAVES, 0, sizeof(u64), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(sc + SC_PER_ROOT, 0, 1024 * sizeof(u64), ctx->stream));
    // The last two plies are fused: the nodes at ply depth-2 are expanded in registers and every child's
    // legal moves are bulk-counted, so the widest materialised level is ply depth-2 (depth >= 3).
    // (the literal path -- a root outside domain W -- materialises every level down to ply depth-1 and counts the legal moves of
    // its nodes with the reference's own legality; such trees are not dealt: rank 0 counts them whole)
    const bool fused = depth >= 3 && !ctx->literal;
    const int tail_ply = fused ? depth - 2 : depth - 1;
    const int unit_ply = shard_world > 1 && !ctx->literal ? perft_unit_ply(depth) : 0;
    const bool idle = shard_world > 1 && unit_ply == 0 && shard_rank != 0; // a tree too shallow to deal is rank 0's
    u64 root_count = 0;
    for (int ply = 0;; ply++) { // level index == ply
        Frontier &f = ctx->levels[ply];
        if (ply == tail_ply) { // frontier reached: bulk-count the leaves
            if (f.n && fused && !idle) {
                ExpandArgs a{};
                a.in = f.soa();
                a.n = f.n;
                a.root_in = (const u32 *)f.root.ptr;
                a.per_root = sc + SC_PER_ROOT;
                a.total = sc + SC_LEAVES;
                Timed t(ctx, "perft_tail");
                k_expand_records<MODE_PERFT><<<(unsigned)nblocks(f.n), BLOCK, 0, ctx->stream>>>(a);
                LL_TRY(check_launch("k_expand_records<PERFT>"));
            } else if (f.n && !idle) {
                Timed t(ctx, "count_leaves");
                k_count_leaves<<<(unsigned)nblocks(f.n), BLOCK, 0, ctx->stream>>>(f.soa(), f.n, ply == 0 ? nullptr : (const u32 *)f.root.ptr,
                                                                                 ply == 0 ? nullptr : sc + SC_PER_ROOT, sc + SC_LEAVES, ctx->literal);
                LL_TRY(check_launch("k_count_leaves"));
            }
            break;
        }
        u64 total = 0;
        LL_TRY(count_level(ctx, ply, false, &total));
        if (ply == 0) {
            root_count = total;
            if (root_count > 1024) return fail(LL_ERR_CAPACITY, "more than 1024 root moves");
        }
        if (ply + 1 == unit_ply) { // this expansion creates the work units: expand canonically with owner tags, keep this rank's
            Frontier &all = ctx->unit_scratch;
            LL_TRY(emit_level_into(ctx, ply, all, false, total, false, ply == 0 ? 2 : 1, false, shard_world));
            LL_TRY(reserve_level(ctx, ply + 1, (size_t)total, false, false)); // at most all of them
            Frontier &g = ctx->levels[ply + 1];
            u64 kept = 0;
            if (total) {
                CUDA_TRY(cudaMemsetAsync(sc + SC_TOTAL, 0, sizeof(u64), ctx->stream));
                {
                    Timed t(ctx, "shard_filter");
                    k_shard_filter<<<(unsigned)nblocks((size_t)total), BLOCK, 0, ctx->stream>>>(
                        ShardArgs{all.soa(), (const u32 *)all.root.ptr, (const u32 *)all.tag.ptr, (size_t)total, g.soa(), (u32 *)g.root.ptr, sc + SC_TOTAL, shard_rank});
                }
                LL_TRY(check_launch("k_shard_filter"));
                CUDA_TRY(cudaMemcpyAsync(&kept, sc + SC_TOTAL, sizeof kept, cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            }
            g.n = (size_t)kept;
            if (!kept) break;
            continue;
        }
        LL_TRY(emit_level(ctx, ply, false, total, false, ply == 0 ? 2 : 1));
        if (!total) break;
    }
    u64 host[1 + 1024];
    CUDA_TRY(cudaMemcpyAsync(&host[0], sc + SC_LEAVES, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&host[1], sc + SC_PER_ROOT, 1024 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *leaves = host[0];
    if (depth == 1) root_count = idle ? 0 : host[0]; // the root's own move count: every root move is one leaf
    if (n_roots) *n_roots = (uint32_t)root_count;
    if (per_root) {
        if (root_count > root_cap) return fail(LL_ERR_CAPACITY, "%llu root moves do not fit root_cap %u", (unsigned long long)root_count, root_cap);
        for (u64 i = 0; i < root_count; i++) per_root[i] = depth == 1 ? 1 : host[1 + i];
    }
    return LL_OK;
}

extern "C" int ll_perft(ll_ctx *ctx, const ll_world_t *root, int depth, int shard_rank, int shard_world, uint64_t *leaves,
                        uint64_t *per_root, uint32_t root_cap, uint32_t *n_roots) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!root || !leaves) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (depth < 0 || depth > 32) return fail(LL_ERR_INVALID_ARG, "depth %d out of range", depth);
    if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) return fail(LL_ERR_INVALID_ARG, "bad shard %d/%d", shard_rank, shard_world);
    *leaves = 0;
    if (n_roots) *n_roots = 0;
    LL_TRY(classify_root(ctx, root, 0));
    LL_TRY(upload_root(ctx, root));
    if (depth == 0) {
        *leaves = shard_rank == 0 ? 1 : 0;
        return LL_OK;
    }
    if (depth >= 3 && !ctx->force_synced && !ctx->literal) {
        bool done = false;
        LL_TRY(perft_chained(ctx, depth, shard_rank, shard_world, leaves, per_root, root_cap, n_roots, &done));
        if (done) return LL_OK;
        *leaves = 0;
        if (n_roots) *n_roots = 0;
        LL_TRY(upload_root(ctx, root)); // level 0 again: a failed chain may have left the scratch levels half-written
    }
    return perft_synced(ctx, depth, shard_rank, shard_world, leaves, per_root, root_cap, n_roots);
    LL_ABI_END
}

extern "C" int ll_perft_device(ll_ctx *ctx, const ll_world_t *root, int depth, int shard_rank, int shard_world, uint64_t *d_out) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!root || !d_out) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (depth < 3 || depth > 32) return fail(LL_ERR_INVALID_ARG, "depth %d out of range (3..32)", depth);
    if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) return fail(LL_ERR_INVALID_ARG, "bad shard %d/%d", shard_rank, shard_world);
    if (!host_well_formed(root)) return fail(LL_ERR_DOMAIN, "position 0 is outside domain W (ll_perft answers the literal domain too; see include/leafline_b200.h)");
    LL_TRY(upload_root(ctx, root));
    bool launched = false;
    LL_TRY(perft_chain_launch(ctx, depth, shard_rank, shard_world, &launched));
    if (!launched) return fail(LL_ERR_CAPACITY, "level buffers are not sized yet: call ll_perft once for this depth first");
    u64 *sc = (u64 *)ctx->scalars.ptr;
    {
        Timed t(ctx, "pack_perft_result");
        k_pack_perft_result<<<1, 1024, 0, ctx->stream>>>(sc + SC_LEAVES, sc + SC_N + 1, (const u32 *)(sc + SC_OVERFLOW), sc + SC_PER_ROOT, (u64 *)d_out);
    }
    return check_launch("k_pack_perft_result");
    LL_ABI_END
}

// ---- horizon / kickoff -------------------------------------------------------------------------------

extern "C" int ll_horizon_batch(ll_ctx *ctx, const ll_world_t *frontier, size_t n, int plies, int quiet_extension, float *values) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (n && (!frontier || !values)) return fail(LL_ERR_INVALID_ARG, "null buffer");
    if (plies < 0 || plies > 8) return fail(LL_ERR_INVALID_ARG, "plies %d out of range", plies);
    if (!n) return LL_OK;
    if (quiet_extension > 8) return fail(LL_ERR_INVALID_ARG, "quiet extension %d out of range", quiet_extension);
    bool small = n <= HORIZON_CHAIN_MAX_BATCH && plies >= 2 && quiet_extension <= 0 && !ctx->force_synced;
    if (small) { // the alpha-beta driver's hand-offs: check on the host, then one launch chain and ONE synchronisation
        for (size_t i = 0; i < n; i++) LL_TRY(classify_root(ctx, &frontier[i], i));
        const bool literal = ctx->literal;
        if (literal) small = false; // the literal path goes level by level
        if (n == 1) LL_TRY(upload_root(ctx, frontier));
        else LL_TRY(upload_worlds(ctx, frontier, n, 0));
        ctx->literal = literal;
    } else {
        LL_TRY(upload_worlds(ctx, frontier, n, 0));
        LL_TRY(validate_level(ctx, ctx->levels[0].soa(), n));
    }
    if (small) {
        bool launched = false;
        LL_TRY(horizon_chain(ctx, 0, plies, &launched));
        if (launched) {
            u32 outgrown = 0;
            CUDA_TRY(cudaMemcpyAsync(values, ctx->levels[0].values.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(&outgrown, (u64 *)ctx->scalars.ptr + SC_OVERFLOW, sizeof outgrown, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (!outgrown) return LL_OK;
        }
    }
    LL_TRY(horizon_levels(ctx, 0, plies, quiet_extension, nullptr));
    CUDA_TRY(cudaMemcpyAsync(values, ctx->levels[0].values.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_horizon_soa(ll_ctx *ctx, const ll_soa_t *soa, int plies, int quiet_extension, float *d_values, uint64_t *leaves_scored) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!soa || !d_values) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (plies < 0 || plies > 8) return fail(LL_ERR_INVALID_ARG, "plies %d out of range", plies);
    if (leaves_scored) *leaves_scored = 0;
    if (!soa->n) return LL_OK;
    LL_TRY(check_soa(soa));
    LL_TRY(validate_level(ctx, Soa{(u64 *)soa->planes, soa->meta, soa->stride}, soa->n));
    // level 0 aliases the caller's planes for the duration of the call; its scratch arrays persist in ctx->ext
    if (ctx->levels.empty()) ctx->levels.resize(1);
    Frontier &ext = ctx->ext;
    LL_TRY(ensure(ctx, ext.counts, soa->stride * sizeof(u32)));
    LL_TRY(ensure(ctx, ext.block_sums, (nblocks(soa->stride) + 1) * sizeof(u32)));
    LL_TRY(ensure(ctx, ext.block_base, (nblocks(soa->stride) + 1) * sizeof(u64)));
    LL_TRY(ensure(ctx, ext.values, soa->stride * sizeof(float)));
    LL_TRY(ensure(ctx, ext.tag, soa->stride * sizeof(u32)));
    const Frontier saved = ctx->levels[0];
    Frontier view = ext;
    view.planes.ptr = soa->planes;
    view.planes.cap = soa->stride * 12 * sizeof(u64);
    view.meta.ptr = soa->meta;
    view.meta.cap = soa->stride * sizeof(u32);
    view.cap = soa->stride;
    view.n = soa->n;
    view.borrowed = true; // reserve_level never reallocates (or frees) the caller's planes
    ctx->levels[0] = view;
    int rc = horizon_levels(ctx, 0, plies, quiet_extension, leaves_scored);
    if (rc == LL_OK && cudaMemcpyAsync(d_values, ctx->levels[0].values.ptr, soa->n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
        rc = fail(LL_ERR_CUDA, "copy of horizon values failed");
    if (rc == LL_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(LL_ERR_CUDA, "synchronize failed");
    ctx->levels[0] = saved;
    return rc;
    LL_ABI_END
}

// variations (nullable): [root_cap][PV_MAX + 1] u32 -- length, then the packed commits of the line BELOW each root move
static int kickoff_core(ll_ctx *ctx, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, int shard_rank, int shard_world,
                        ll_commit_t *roots, float *scores, uint32_t root_cap, uint32_t *n_roots, uint64_t *leaves_scored, u32 *variations) {
    LL_TRY(check_ctx(ctx));
    if (!root || !scores || !n_roots) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (depth < 1 || depth > 9) return fail(LL_ERR_INVALID_ARG, "depth %d out of range (1..9)", depth);
    if (quiet_extension > 8) return fail(LL_ERR_INVALID_ARG, "quiet extension %d out of range", quiet_extension);
    if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) return fail(LL_ERR_INVALID_ARG, "bad shard %d/%d", shard_rank, shard_world);
    *n_roots = 0;
    if (leaves_scored) *leaves_scored = 0;
    LL_TRY(classify_root(ctx, root, 0));
    LL_TRY(upload_root(ctx, root));
    u64 total = 0;
    LL_TRY(count_level(ctx, 0, nihilistically != 0, &total)); // mind.rs:388-392
    if (total > root_cap) return fail(LL_ERR_CAPACITY, "%llu root moves do not fit root_cap %u", (unsigned long long)total, root_cap);
    *n_roots = (uint32_t)total;
    if (!total) return LL_OK;
    LL_TRY(emit_level(ctx, 0, nihilistically != 0, total, true, 2));
    Frontier &g = ctx->levels[1];
    if (roots) {
        LL_TRY(ensure(ctx, ctx->aos_out, (size_t)total * sizeof(ll_commit_t)));
        {
            Timed t(ctx, "commits_to_aos");
            k_commits_to_aos<<<(unsigned)nblocks((size_t)total), BLOCK, 0, ctx->stream>>>(g.soa(), (const u32 *)g.cmeta.ptr, (size_t)total, (u64 *)ctx->aos_out.ptr);
        }
        LL_TRY(check_launch("k_commits_to_aos"));
        CUDA_TRY(cudaMemcpyAsync(roots, ctx->aos_out.ptr, (size_t)total * sizeof(ll_commit_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    size_t search_level = 1;
    size_t kept = (size_t)total;
    if (shard_world > 1) { // this rank's root subtrees
        kept = total > (u64)shard_rank ? (size_t)((total - shard_rank + shard_world - 1) / shard_world) : 0;
        LL_TRY(reserve_level(ctx, 2, kept, false, false));
        Frontier &h = ctx->levels[2];
        if (kept) {
            Timed t(ctx, "shard_filter");
            k_deal_units<<<(unsigned)nblocks(kept), BLOCK, 0, ctx->stream>>>(ctx->levels[1].soa(), (const u32 *)ctx->levels[1].root.ptr, ctx->levels[1].n, shard_rank,
                                                                            shard_world, h.soa(), (u32 *)h.root.ptr);
            LL_TRY(check_launch("k_shard_filter"));
        }
        h.n = kept;
        search_level = 2;
    }
    uint64_t scored = 0;
    size_t bottom = search_level;
    std::vector<float> vals(kept);
    bool chained = false;
    if (!variations && quiet_extension <= 0 && depth - 1 >= 2 && kept && kept <= HORIZON_CHAIN_MAX_BATCH && !ctx->force_synced && !ctx->literal) {
        // scores only: the levels below the root moves as one launch chain, no host round trip per ply (see horizon_chain)
        bool launched = false;
        LL_TRY(horizon_chain(ctx, search_level, depth - 1, &launched));
        if (launched) {
            u64 *sc = (u64 *)ctx->scalars.ptr;
            u64 flags[2] = {0, 0};
            CUDA_TRY(cudaMemcpyAsync(vals.data(), ctx->levels[search_level].values.ptr, kept * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(&flags[0], sc + SC_OVERFLOW, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(&flags[1], sc + SC_SCORED, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (!(u32)flags[0]) {
                chained = true;
                scored = flags[1];
            }
        }
    }
    if (!chained) LL_TRY(horizon_levels(ctx, search_level, depth - 1, quiet_extension, &scored, variations != nullptr, &bottom)); // mind.rs:406-411: full window per root
    if (leaves_scored) *leaves_scored = scored;
    std::vector<u32> pv;
    const int expanding = (depth - 1) + (quiet_extension > 0 ? quiet_extension : 0);
    if (variations && kept && expanding > 0) {
        if (bottom - search_level + 1 > (size_t)PV_LEVELS) return fail(LL_ERR_CAPACITY, "variation deeper than %d levels", PV_LEVELS);
        PvArgs pa{};
        pa.n_levels = (int)(bottom - search_level + 1);
        for (int L = 0; L < pa.n_levels; L++) {
            Frontier &f = ctx->levels[search_level + L];
            pa.lev[L] = PvLevel{f.soa(), (const u32 *)f.counts.ptr, (const u64 *)f.block_base.ptr, (const float *)f.values.ptr, (const u32 *)f.cmeta.ptr,
                                (depth - 1) - L <= 0};
        }
        pa.n = kept;
        LL_TRY(ensure(ctx, ctx->flags, kept * (PV_MAX + 1) * sizeof(u32)));
        pa.out = (u32 *)ctx->flags.ptr;
        {
            Timed t(ctx, "variation");
            k_variation<<<(unsigned)nblocks(kept), BLOCK, 0, ctx->stream>>>(pa);
        }
        LL_TRY(check_launch("k_variation"));
        pv.resize(kept * (PV_MAX + 1));
        CUDA_TRY(cudaMemcpyAsync(pv.data(), ctx->flags.ptr, pv.size() * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (kept && !chained) CUDA_TRY(cudaMemcpyAsync(vals.data(), ctx->levels[search_level].values.ptr, kept * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (u64 i = 0; i < total; i++) scores[i] = -INFINITY;
    for (size_t o = 0; o < kept; o++) scores[(size_t)shard_rank + o * (size_t)shard_world] = -vals[o]; // mind.rs:425
    if (variations) {
        for (u64 i = 0; i < total; i++) variations[i * (PV_MAX + 1)] = 0;
        if (!pv.empty())
            for (size_t o = 0; o < kept; o++)
                memcpy(variations + ((size_t)shard_rank + o * (size_t)shard_world) * (PV_MAX + 1), pv.data() + o * (PV_MAX + 1), (PV_MAX + 1) * sizeof(u32));
    }
    return LL_OK;
}

extern "C" int ll_kickoff(ll_ctx *ctx, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, int shard_rank, int shard_world,
                          ll_commit_t *roots, float *scores, uint32_t root_cap, uint32_t *n_roots, uint64_t *leaves_scored) {
    LL_ABI_BEGIN
    return kickoff_core(ctx, root, depth, quiet_extension, nihilistically, shard_rank, shard_world, roots, scores, root_cap, n_roots, leaves_scored, nullptr);
    LL_ABI_END
}

// ---- the kickoff family as the reference's callers see it (mind.rs:441-490): forecasts sorted by score ------
static void unpack_patch(u32 m, ll_patch_t *p) {
    p->team = (uint8_t)(m & 1u);
    p->job = (uint8_t)((m >> 1) & 7u);
    const u32 from = (m >> 4) & 63u, to = (m >> 10) & 63u;
    p->whence = (uint8_t)(((from >> 3) << 4) | (from & 7u));
    p->whither = (uint8_t)(((to >> 3) << 4) | (to & 7u));
}

extern "C" int ll_forecast(ll_ctx *ctx, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, ll_forecast_t *forecasts,
                           uint32_t cap, uint32_t *n_forecasts, uint64_t *leaves_scored) {
    LL_ABI_BEGIN
    if (!forecasts || !n_forecasts) return fail(LL_ERR_INVALID_ARG, "null argument");
    static_assert(PV_MAX == LL_MAX_VARIATION, "variation capacity");
    std::vector<ll_commit_t> roots(1024);
    std::vector<float> scores(1024);
    std::vector<u32> pv((size_t)1024 * (PV_MAX + 1));
    uint32_t n = 0;
    LL_TRY(kickoff_core(ctx, root, depth, quiet_extension, nihilistically, 0, 1, roots.data(), scores.data(), 1024, &n, leaves_scored, pv.data()));
    *n_forecasts = n;
    if (n > cap) return fail(LL_ERR_CAPACITY, "%u forecasts do not fit the capacity of %u", n, cap);
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return scores[x] > scores[y]; }); // mind.rs:436
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t i = order[k];
        ll_forecast_t &f = forecasts[k];
        memset(&f, 0, sizeof f);
        f.commit = roots[i];
        f.score = scores[i];
        const u32 *line = pv.data() + (size_t)i * (PV_MAX + 1);
        f.variation[0] = ll_patch_t{roots[i].team, roots[i].job, roots[i].whence, roots[i].whither}; // T::flash(premonition.patch), mind.rs:426
        uint32_t len = 1;
        for (u32 j = 0; j < line[0] && len < LL_MAX_VARIATION; j++) unpack_patch(line[1 + j], &f.variation[len++]);
        f.variation_len = len;
    }
    return LL_OK;
    LL_ABI_END
}

// ---- the host alpha-beta driver (mind.rs:145-169, 271-364) over device frontier batches -----------------------
// The branchy part of the search stays on the host, as in the reference: negamax with alpha-beta cut-offs, MVV-LVA +
// history move ordering, a full window below every root move (mind.rs:406-411).  Where the remaining depth reaches
// `device_plies` the node is handed to the device, which returns its EXACT full-width value (horizon_levels), so
// the root scores equal TT-free exact negamax -- only the work is pruned.  The children of a node one ply above
// the hand-off are valued in two batches ("young brothers wait"): the first commit alone, then -- unless it
// already caused a cut-off -- all the others at once.
// The "memory bank" (mind.rs:312-344) is kept, made sound: an entry is keyed by the WHOLE position plus the plies left
// (no hash collisions, no deeper-entry substitution) and says what its value is -- exact, a lower bound (the node
// failed high) or an upper bound (it failed low) -- where the reference files every value as exact (mind.rs:340-343).
// A bound is used only where it settles the window, so the root scores stay those of exact negamax; frontier values
// are exact by construction and are remembered the same way.
namespace {

// the "memory bank" (mind.rs:256-267, 312-344, 367-372): least-recently-used eviction under a byte budget, like the
// reference's LruCache sized by déjà_vu_table_size_bound
struct Bank {
    struct Key { // the position and the plies left
        u64 w[13];
        bool operator==(const Key &o) const { return memcmp(w, o.w, sizeof w) == 0; }
    };
    struct KeyHash {
        size_t operator()(const Key &k) const {
            u64 h = 0x9E3779B97F4A7C15ull;
            for (int i = 0; i < 13; i++) {
                h = (h ^ k.w[i]) * 0xBF58476D1CE4E5B9ull;
                h ^= h >> 29;
            }
            return (size_t)h;
        }
    };
    enum : uint8_t { EXACT, LOWER, UPPER };
    struct Memory {
        float value;
        uint8_t kind;
        uint32_t best; // patch_key of the commit that gave the value (0xFFFFFFFF: none -- a frontier value, a leaf)
    };
    struct Entry {
        Memory memory;
        std::list<Key>::iterator at;
    };
    static constexpr size_t ENTRY_BYTES = sizeof(Key) * 2 + sizeof(Entry) + 64; // key in the map and in the list + node overheads
    std::unordered_map<Key, Entry, KeyHash> map;
    std::list<Key> order; // most recently used first
    size_t capacity = 1;
    uint64_t evicted = 0;
    explicit Bank(double gib) {
        const double bytes = gib * 1024.0 * 1024.0 * 1024.0;
        capacity = bytes / ENTRY_BYTES < 1024.0 ? 1024 : (size_t)(bytes / ENTRY_BYTES);
    }
    static Key key_of(const ll_world_t &w, int plies) {
        Key k;
        for (int j = 0; j < 12; j++) k.w[j] = w.plane[j];
        k.w[12] = (u64)w.initiative | ((u64)w.service_eligibility << 8) | ((u64)w.passing_by << 16) | ((u64)(uint32_t)plies << 32);
        return k;
    }
    const Memory *find(const Key &k) {
        auto it = map.find(k);
        if (it == map.end()) return nullptr;
        order.splice(order.begin(), order, it->second.at); // used: to the front
        return &it->second.memory;
    }
    void file(const Key &k, float value, uint8_t kind, uint32_t best) {
        auto it = map.find(k);
        if (it != map.end()) {
            it->second.memory = Memory{value, kind, best};
            order.splice(order.begin(), order, it->second.at);
            return;
        }
        if (map.size() >= capacity) { // forget the least recently used
            map.erase(order.back());
            order.pop_back();
            evicted++;
        }
        order.push_front(k);
        map.emplace(k, Entry{Memory{value, kind, best}, order.begin()});
    }
};

using Intuition = std::unordered_map<uint32_t, uint32_t>; // the "intuition bank": Patch -> credit (mind.rs:353-359)

// The searchers' single door to the context.  The reference searches every root move on a thread of its own
// (mind.rs:399-414); so does this driver, and since an ll_ctx serves one caller at a time the threads' requests -- "the
// commits of this node", "the exact values of these frontier nodes" -- are COMBINED: a thread that needs the device files its
// request and sleeps; the last thread still running finds everybody else asleep and serves the whole round with one
// ll_lookahead_batch and one ll_horizon_batch per distinct depth.  Twenty root moves mean twenty times fewer device round
// trips than a single depth-first walk -- the host's latency, not the device's work, is what bounds the pruned search.
struct DeviceGate {
    struct Request {
        int kind; // 0: reckless commits of nodes[0]; 1: values of nodes[0..n) with `plies` plies left
        const ll_world_t *nodes;
        size_t n;
        int plies;
        std::vector<ll_commit_t> *commits;
        float *values;
        int rc = LL_OK;
        bool done = false;
    };
    ll_ctx *ctx;
    int quiet_extension;
    std::mutex mu;
    std::condition_variable cv;
    int active = 0, waiting = 0;
    std::vector<Request *> pending;
    char err[sizeof g_error] = "";
    uint64_t rounds = 0;
    std::vector<ll_world_t> worlds;
    std::vector<uint32_t> offsets;
    std::vector<ll_commit_t> commits;
    std::vector<float> values;

    void serve_round() { // mu held; every other searcher is asleep on its request
        rounds++;
        std::vector<Request *> kids, vals;
        for (Request *r : pending) (r->kind == 0 ? kids : vals).push_back(r);
        if (!kids.empty()) {
            worlds.resize(kids.size());
            for (size_t i = 0; i < kids.size(); i++) worlds[i] = kids[i]->nodes[0];
            offsets.assign(kids.size() + 1, 0);
            if (commits.size() < kids.size() * 128) commits.resize(kids.size() * 128);
            int rc = ll_lookahead_batch(ctx, worlds.data(), worlds.size(), 1, offsets.data(), commits.data(), commits.size()); // reckless, mind.rs:279
            if (rc == LL_ERR_CAPACITY && offsets[kids.size()] > commits.size()) { // the counts came back: make room and ask again
                commits.resize(offsets[kids.size()]);
                rc = ll_lookahead_batch(ctx, worlds.data(), worlds.size(), 1, offsets.data(), commits.data(), commits.size());
            }
            if (rc != LL_OK) memcpy(err, g_error, sizeof err);
            for (size_t i = 0; i < kids.size(); i++) {
                kids[i]->rc = rc;
                if (rc == LL_OK) kids[i]->commits->assign(commits.begin() + offsets[i], commits.begin() + offsets[i + 1]);
            }
        }
        while (!vals.empty()) { // one batch per distinct depth (usually one)
            const int plies = vals[0]->plies;
            std::vector<Request *> now, later;
            for (Request *r : vals) (r->plies == plies ? now : later).push_back(r);
            size_t total = 0;
            for (Request *r : now) total += r->n;
            worlds.resize(total);
            values.resize(total);
            size_t at = 0;
            for (Request *r : now) {
                memcpy(worlds.data() + at, r->nodes, r->n * sizeof(ll_world_t));
                at += r->n;
            }
            // the levels below a hand-off grow ~30-fold a ply: the combined batch goes to the device in pieces whose deepest
            // materialised level stays within a few GB (and in halves of that if HBM is short all the same)
            size_t piece = total;
            {
                double grown = 1.0;
                for (int p = 1; p < plies; p++) grown *= 30.0;
                const double fit = 3.2e7 / grown;
                if (fit < (double)piece) piece = fit < 1.0 ? 1 : (size_t)fit;
            }
            int rc = LL_OK;
            for (size_t first = 0; first < total && rc == LL_OK;) {
                const size_t m = std::min(piece, total - first);
                rc = ll_horizon_batch(ctx, worlds.data() + first, m, plies, quiet_extension, values.data() + first);
                if (rc == LL_ERR_OOM && m > 1) {
                    piece = m / 2;
                    rc = LL_OK;
                    continue;
                }
                first += m;
            }
            if (rc != LL_OK) memcpy(err, g_error, sizeof err);
            at = 0;
            for (Request *r : now) {
                r->rc = rc;
                if (rc == LL_OK) memcpy(r->values, values.data() + at, r->n * sizeof(float));
                at += r->n;
            }
            vals.swap(later);
        }
        for (Request *r : pending) r->done = true;
        pending.clear();
        waiting = 0;
        cv.notify_all();
    }
    int submit(Request &r) {
        std::unique_lock<std::mutex> lk(mu);
        pending.push_back(&r);
        waiting++;
        if (waiting == active) serve_round();
        else cv.wait(lk, [&] { return r.done; });
        if (r.rc != LL_OK) memcpy(g_error, err, sizeof g_error);
        return r.rc;
    }
    void leave() { // a searcher has no more root moves: the others may all be waiting for a round nobody would serve
        std::lock_guard<std::mutex> lk(mu);
        active--;
        if (waiting && waiting == active) serve_round();
    }
};

struct SearchShared { // what the searcher threads of one kickoff share (mind.rs:381-386: the banks behind Arc<Mutex>)
    DeviceGate gate;
    Intuition &intuition;
    Bank bank;
    std::mutex banks;
    int device_plies;
    bool timebound = false;
    std::chrono::steady_clock::time_point deadline;
    std::atomic<bool> out_of_time{false};
    SearchShared(ll_ctx *ctx, int plies, int quiet, Intuition &i) : intuition(i), bank(ctx->deja_vu_gib), device_plies(plies) {
        gate.ctx = ctx;
        gate.quiet_extension = quiet;
    }
};

struct AlphaBeta { // one searcher
    SearchShared &sh;
    uint64_t expanded = 0, handed_off = 0, remembered = 0;
    int rc = LL_OK;
    explicit AlphaBeta(SearchShared &s) : sh(s) {}

    static constexpr uint32_t NO_PATCH = 0xFFFFFFFFu;
    static uint32_t patch_key(const ll_commit_t &c) { return (uint32_t)c.team | ((uint32_t)c.job << 1) | ((uint32_t)c.whence << 8) | ((uint32_t)c.whither << 16); }
    static float figurine_valuation(uint8_t agent) { // mind.rs:36-48
        static const float value[6] = {1.0f, 3.2f, 3.3f, 5.1f, 8.8f, 20000.0f};
        return ((agent >> 3) ? -1.0f : 1.0f) * value[agent & 7];
    }
    static float mvv_lva(const ll_commit_t &c) { // mind.rs:145-153
        if (c.hospitalization == LL_NONE) return 0.0f;
        return figurine_valuation(c.hospitalization) - figurine_valuation((uint8_t)((c.team << 3) | c.job));
    }
    void order(std::vector<ll_commit_t> &commits) { // mind.rs:155-169 (history first, then MVV-LVA; stable here)
        std::vector<std::tuple<uint32_t, float, uint32_t>> keys(commits.size());
        {
            std::lock_guard<std::mutex> lk(sh.banks);
            for (uint32_t i = 0; i < commits.size(); i++) {
                auto it = sh.intuition.find(patch_key(commits[i]));
                keys[i] = {it == sh.intuition.end() ? 0u : it->second + 1u, mvv_lva(commits[i]), i};
            }
        }
        std::stable_sort(keys.begin(), keys.end(), [](const auto &a, const auto &b) {
            if (std::get<0>(a) != std::get<0>(b)) return std::get<0>(a) > std::get<0>(b);
            return std::get<1>(a) > std::get<1>(b);
        });
        std::vector<ll_commit_t> sorted(commits.size());
        for (uint32_t i = 0; i < commits.size(); i++) sorted[i] = commits[std::get<2>(keys[i])];
        commits.swap(sorted);
    }
    bool recall(const Bank::Key &k, Bank::Memory *out) {
        std::lock_guard<std::mutex> lk(sh.banks);
        const Bank::Memory *m = sh.bank.find(k);
        if (m) *out = *m;
        return m != nullptr;
    }
    void file(const Bank::Key &k, float value, uint8_t kind, uint32_t best) {
        std::lock_guard<std::mutex> lk(sh.banks);
        sh.bank.file(k, value, kind, best);
    }
    bool children(const ll_world_t &node, std::vector<ll_commit_t> &out) {
        DeviceGate::Request r{0, &node, 1, 0, &out, nullptr};
        rc = sh.gate.submit(r);
        if (rc != LL_OK) return false;
        expanded++;
        return true;
    }
    // exact values of frontier nodes with `plies` plies left: the remembered ones from the bank, the others from the device
    bool device_values(const ll_world_t *nodes, size_t n, int plies, float *values) {
        std::vector<ll_world_t> unknown;
        std::vector<size_t> where;
        for (size_t i = 0; i < n; i++) {
            Bank::Memory m;
            if (recall(Bank::key_of(nodes[i], plies), &m)) { // frontier entries are always exact
                values[i] = m.value;
                remembered++;
            } else {
                unknown.push_back(nodes[i]);
                where.push_back(i);
            }
        }
        if (unknown.empty()) return true;
        std::vector<float> fresh(unknown.size());
        DeviceGate::Request r{1, unknown.data(), unknown.size(), plies, nullptr, fresh.data()};
        rc = sh.gate.submit(r);
        if (rc != LL_OK) return false;
        handed_off += unknown.size();
        for (size_t j = 0; j < unknown.size(); j++) {
            values[where[j]] = fresh[j];
            file(Bank::key_of(unknown[j], plies), fresh[j], Bank::EXACT, NO_PATCH);
        }
        return true;
    }
    // value of `node` for its side to move with `depth` plies left, window (alpha, beta); fail-soft
    float search(const ll_world_t &node, int depth, float alpha, float beta) {
        if (rc != LL_OK) return 0.0f;
        if (sh.out_of_time.load(std::memory_order_relaxed) || (sh.timebound && std::chrono::steady_clock::now() > sh.deadline)) {
            sh.out_of_time.store(true);
            rc = LL_ERR_CAPACITY; // unwinds the recursion; the caller reports the time-out, not this code
            return 0.0f;
        }
        const int device_plies = sh.device_plies;
        if (depth <= device_plies) {
            float v = 0.0f;
            device_values(&node, 1, depth < 0 ? 0 : depth, &v);
            return v;
        }
        const Bank::Key key = Bank::key_of(node, depth);
        Bank::Memory m;
        if (recall(key, &m)) { // mind.rs:316-325, with the kinds of value told apart
            if (m.kind == Bank::EXACT || (m.kind == Bank::LOWER && m.value >= beta) || (m.kind == Bank::UPPER && m.value <= alpha)) {
                remembered++;
                return m.value;
            }
        }
        const float alpha_in = alpha;
        std::vector<ll_commit_t> premonitions;
        if (!children(node, premonitions)) return 0.0f;
        if (premonitions.empty()) { // mind.rs:282-287 at depth > 0: the node is its own leaf
            float v = 0.0f;
            DeviceGate::Request r{1, &node, 1, 0, nullptr, &v};
            rc = sh.gate.submit(r); // (a leaf: plies 0; a quiescence extension finds nothing to stun here either)
            if (rc == LL_OK) file(key, v, Bank::EXACT, NO_PATCH);
            return v;
        }
        order(premonitions);
        float optimum = -INFINITY;
        uint32_t optimand = NO_PATCH;
        auto remember = [&]() { // mind.rs:340-343
            if (rc == LL_OK) file(key, optimum, optimum <= alpha_in ? Bank::UPPER : optimum >= beta ? Bank::LOWER : Bank::EXACT, optimand);
            return optimum;
        };
        auto consider = [&](const ll_commit_t &c, float value) { // mind.rs:346-361
            if (value > optimum) {
                optimum = value;
                optimand = patch_key(c);
            }
            if (value > alpha) alpha = value;
            if (alpha >= beta) {
                std::lock_guard<std::mutex> lk(sh.banks);
                sh.intuition[patch_key(c)] += 1u << (depth > 30 ? 30 : depth);
                return true;
            }
            return false;
        };
        if (depth - 1 <= device_plies) { // the children are frontier nodes: first alone, then the rest in one batch
            float first = 0.0f;
            if (!device_values(&premonitions[0].tree, 1, depth - 1, &first)) return 0.0f;
            if (consider(premonitions[0], -first)) return remember();
            const size_t rest = premonitions.size() - 1;
            if (rest) {
                std::vector<ll_world_t> trees(rest);
                std::vector<float> values(rest);
                for (size_t i = 0; i < rest; i++) trees[i] = premonitions[i + 1].tree;
                if (!device_values(trees.data(), rest, depth - 1, values.data())) return 0.0f;
                for (size_t i = 0; i < rest; i++)
                    if (consider(premonitions[i + 1], -values[i])) break;
            }
            return remember();
        }
        for (const ll_commit_t &c : premonitions) {
            const float value = -search(c.tree, depth - 1, -beta, -alpha);
            if (rc != LL_OK) return 0.0f;
            if (consider(c, value)) break;
        }
        return remember();
    }
    // The principal variation below a root move (the `Variation` memory, mind.rs:208-224): follow the bank's exact entries
    // through the plies the host searched -- every node of a principal variation has an exact entry naming its best commit --
    // then ask the device for the rest of the line below the frontier node (ll_forecast's variation kernel).  Called after
    // the searchers have finished: straight calls on the context.
    void variation_below(ll_world_t node, int depth, ll_patch_t *line, uint32_t *len) {
        ll_ctx *ctx = sh.gate.ctx;
        std::vector<ll_commit_t> kids(1024);
        while (*len < LL_MAX_VARIATION && rc == LL_OK) {
            if (depth <= sh.device_plies) {
                if (depth < 1) return;
                std::vector<ll_forecast_t> tail(1024);
                uint32_t n = 0;
                if (ll_forecast(ctx, &node, depth, sh.gate.quiet_extension, 1, tail.data(), 1024, &n, nullptr) != LL_OK || !n) return; // (reckless children, as everywhere below the root)
                for (uint32_t j = 0; j < tail[0].variation_len && *len < LL_MAX_VARIATION; j++) line[(*len)++] = tail[0].variation[j];
                return;
            }
            Bank::Memory m;
            if (!recall(Bank::key_of(node, depth), &m) || m.kind != Bank::EXACT || m.best == NO_PATCH) return;
            uint32_t offsets[2] = {0, 0};
            kids.resize(1024);
            if (ll_lookahead_batch(ctx, &node, 1, 1, offsets, kids.data(), kids.size()) != LL_OK) return;
            kids.resize(offsets[1]);
            bool found = false;
            for (const ll_commit_t &c : kids)
                if (patch_key(c) == m.best && !found) { // (ascensions share a patch: the first, like the reference's flash(patch))
                    line[(*len)++] = ll_patch_t{c.team, c.job, c.whence, c.whither};
                    node = c.tree;
                    found = true;
                }
            if (!found) return;
            depth--;
        }
    }
};

} // namespace

// `deadline` (nullable): give up once it has passed -> *timed_out = true, forecasts untouched.  `intuition` (nullable): the
// caller's intuition bank, kept across calls.  shard_world > 1: only the root moves i % shard_world == shard_rank are
// searched (the others get -INFINITY and no variation: gather with a max, like ll_kickoff).
static int alpha_beta_core(ll_ctx *ctx, const ll_world_t *root, int depth, int device_plies, int quiet_extension, int nihilistically,
                           ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts, uint64_t *stats3,
                           const std::chrono::steady_clock::time_point *deadline, bool *timed_out, Intuition *intuition = nullptr, int shard_rank = 0,
                           int shard_world = 1, bool sorted = true) {
    if (timed_out) *timed_out = false;
    uint32_t offsets[2] = {0, 0};
    std::vector<ll_commit_t> roots(1024);
    LL_TRY(ll_lookahead_batch(ctx, root, 1, nihilistically, offsets, roots.data(), roots.size())); // mind.rs:388-392
    roots.resize(offsets[1]);
    if (offsets[1] > cap) {
        *n_forecasts = offsets[1];
        return fail(LL_ERR_CAPACITY, "%u forecasts do not fit the capacity of %u", offsets[1], cap);
    }
    Intuition local;
    SearchShared shared(ctx, device_plies, quiet_extension, intuition ? *intuition : local);
    if (deadline) {
        shared.timebound = true;
        shared.deadline = *deadline;
    }
    // the root moves are taken up in the order the intuition bank suggests (mind.rs:393-396), each by a searcher thread of
    // its own (mind.rs:399-414); what is returned is ordered by score, ties in canonical order
    std::vector<uint32_t> sequence;
    {
        AlphaBeta orderer(shared);
        std::vector<ll_commit_t> ordered = roots;
        orderer.order(ordered);
        for (const ll_commit_t &c : ordered)
            for (uint32_t i = 0; i < roots.size(); i++)
                if (!memcmp(&c, &roots[i], sizeof(ll_commit_t)) && (int)(i % (uint32_t)shard_world) == shard_rank) sequence.push_back(i);
    }
    std::vector<float> scores(roots.size(), -INFINITY);
    const size_t n_threads = std::min<size_t>(sequence.size(), ctx->search_threads > 0 ? (size_t)ctx->search_threads : 1);
    std::vector<AlphaBeta> searchers(n_threads ? n_threads : 1, AlphaBeta(shared));
    std::atomic<size_t> next{0};
    shared.gate.active = (int)n_threads;
    auto work = [&](size_t t) {
        AlphaBeta &ab = searchers[t];
        try {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= sequence.size() || ab.rc != LL_OK) break;
                const uint32_t i = sequence[k];
                scores[i] = -ab.search(roots[i].tree, depth - 1, -INFINITY, INFINITY); // every root move gets the full window (mind.rs:406-411)
            }
        } catch (...) {
            ab.rc = fail(LL_ERR_OOM, "host allocation failed in a searcher thread");
        }
        shared.gate.leave();
    };
    if (n_threads <= 1) {
        if (n_threads) work(0);
    } else {
        std::vector<std::thread> threads;
        for (size_t t = 1; t < n_threads; t++) threads.emplace_back(work, t);
        work(0);
        for (std::thread &t : threads) t.join();
    }
    if (shared.out_of_time.load()) {
        if (timed_out) *timed_out = true;
        return LL_OK;
    }
    for (AlphaBeta &ab : searchers)
        if (ab.rc != LL_OK) {
            if (shared.gate.err[0]) memcpy(g_error, shared.gate.err, sizeof g_error);
            return ab.rc;
        }
    *n_forecasts = offsets[1];
    std::vector<uint32_t> order(roots.size());
    for (uint32_t i = 0; i < roots.size(); i++) order[i] = i;
    if (sorted) std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return scores[x] > scores[y]; }); // mind.rs:436
    AlphaBeta reader(shared);
    for (uint32_t k = 0; k < roots.size(); k++) {
        const uint32_t i = order[k];
        ll_forecast_t &f = forecasts[k];
        memset(&f, 0, sizeof f);
        f.commit = roots[i];
        f.score = scores[i];
        f.variation[0] = ll_patch_t{roots[i].team, roots[i].job, roots[i].whence, roots[i].whither}; // T::flash(premonition.patch), mind.rs:426
        f.variation_len = 1;
        if ((int)(i % (uint32_t)shard_world) == shard_rank) reader.variation_below(roots[i].tree, depth - 1, f.variation, &f.variation_len);
    }
    if (stats3) {
        stats3[0] = stats3[1] = stats3[2] = 0;
        for (AlphaBeta &ab : searchers) {
            stats3[0] += ab.expanded;
            stats3[1] += ab.handed_off;
            stats3[2] += ab.remembered;
        }
    }
    return LL_OK;
}

extern "C" int ll_set_search_threads(ll_ctx *ctx, int threads) {
    if (!ctx) return fail(LL_ERR_INVALID_ARG, "null context");
    if (threads < 1 || threads > 256) return fail(LL_ERR_INVALID_ARG, "%d searcher threads out of range (1..256)", threads);
    ctx->search_threads = threads;
    return LL_OK;
}

extern "C" int ll_set_deja_vu_bound(ll_ctx *ctx, double gib) {
    if (!ctx) return fail(LL_ERR_INVALID_ARG, "null context");
    if (!(gib > 0.0) || gib > 1024.0) return fail(LL_ERR_INVALID_ARG, "memory bound %g GiB out of range", gib);
    ctx->deja_vu_gib = gib;
    return LL_OK;
}

extern "C" int ll_alpha_beta_forecast(ll_ctx *ctx, const ll_world_t *root, int depth, int device_plies, int quiet_extension, int nihilistically,
                                      ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts, uint64_t *stats3) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!root || !forecasts || !n_forecasts) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (depth < 1 || depth > 32) return fail(LL_ERR_INVALID_ARG, "depth %d out of range (1..32)", depth);
    if (device_plies < 0 || device_plies > 8) return fail(LL_ERR_INVALID_ARG, "device_plies %d out of range (0..8)", device_plies);
    if (quiet_extension > 8) return fail(LL_ERR_INVALID_ARG, "quiet extension %d out of range", quiet_extension);
    return alpha_beta_core(ctx, root, depth, device_plies, quiet_extension, nihilistically, forecasts, cap, n_forecasts, stats3, nullptr, nullptr);
    LL_ABI_END
}

extern "C" int ll_fixed_depth_sequence_forecast(ll_ctx *ctx, const ll_world_t *root, const uint8_t *depths, uint32_t n_depths, int nihilistically,
                                                ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts) {
    LL_ABI_BEGIN
    if (!depths || !n_depths) return fail(LL_ERR_INVALID_ARG, "`depth_sequence` should be nonempty"); // mind.rs:480
    for (uint32_t k = 0; k < n_depths; k++) LL_TRY(ll_forecast(ctx, root, depths[k], -1, nihilistically, forecasts, cap, n_forecasts, nullptr));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_iterative_deepening_forecast(ll_ctx *ctx, const ll_world_t *root, double timeout_seconds, int nihilistically, ll_forecast_t *forecasts,
                                               uint32_t cap, uint32_t *n_forecasts, uint32_t *depth_reached) {
    LL_ABI_BEGIN
    if (!forecasts || !n_forecasts || !depth_reached) return fail(LL_ERR_INVALID_ARG, "null argument");
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    const auto deadline = t0 + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(timeout_seconds > 0 ? timeout_seconds : 0));
    // mind.rs:451-469: depth 1 always completes; then deepen while an iteration finishes before the deadline.  The
    // shallow iterations are full-width searches resident in HBM (ll_forecast: milliseconds, with variations); once the
    // tree no longer fits -- depth 8 from the opening -- the host alpha-beta driver takes over (same scores, root move
    // as the variation) and is abandoned in mid-search if it overruns, like the reference's time-bound search.
    LL_TRY(ll_forecast(ctx, root, 1, -1, nihilistically, forecasts, cap, n_forecasts, nullptr));
    *depth_reached = 1;
    std::vector<ll_forecast_t> next(cap < 1024 ? cap : 1024);
    Intuition intuition; // one intuition bank for all the iterations (mind.rs:456-464)
    bool full_width = true;
    for (int depth = 2; depth <= 32; depth++) {
        if (elapsed() >= timeout_seconds) break;
        uint32_t n = 0;
        if (full_width && depth >= 8) full_width = false;
        if (full_width) {
            const int rc = ll_forecast(ctx, root, depth, -1, nihilistically, next.data(), (uint32_t)next.size(), &n, nullptr);
            if (rc == LL_ERR_OOM) full_width = false; // the full-width tree no longer fits HBM
            else LL_TRY(rc);
        }
        if (!full_width) {
            bool timed_out = false;
            LL_TRY(alpha_beta_core(ctx, root, depth, depth >= 9 ? 5 : 4, -1, nihilistically, next.data(), (uint32_t)next.size(), &n, nullptr, &deadline, &timed_out, &intuition));
            if (timed_out) break;
        }
        if (elapsed() > timeout_seconds) break; // finished late: discarded, like the reference's `None` (mind.rs:416-420)
        memcpy(forecasts, next.data(), n * sizeof(ll_forecast_t));
        *n_forecasts = n;
        *depth_reached = (uint32_t)depth;
        if (!n) break; // nothing to play
        // (no guess at whether the next iteration can finish: like the reference, the budget is used -- an iteration that
        // overruns is abandoned in mid-search, mind.rs:416-420)
    }
    return LL_OK;
    LL_ABI_END
}

// ---- device-resident positions -----------------------------------------------------------------------

extern "C" int ll_soa_alloc(ll_ctx *ctx, size_t n, ll_soa_t *soa) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!soa) return fail(LL_ERR_INVALID_ARG, "null soa");
    memset(soa, 0, sizeof *soa);
    const size_t stride = ((n + 63) / 64) * 64;
    if (!stride) return LL_OK;
    cudaError_t e = cudaMalloc((void **)&soa->planes, stride * 12 * sizeof(u64));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LL_ERR_OOM, "cudaMalloc of %zu positions failed: %s", n, cudaGetErrorString(e));
    }
    e = cudaMalloc((void **)&soa->meta, stride * sizeof(u32));
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(soa->planes);
        soa->planes = nullptr;
        return fail(LL_ERR_OOM, "cudaMalloc of %zu positions failed: %s", n, cudaGetErrorString(e));
    }
    CUDA_TRY(cudaMemsetAsync(soa->planes, 0, stride * 12 * sizeof(u64), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(soa->meta, 0, stride * sizeof(u32), ctx->stream));
    soa->stride = stride;
    soa->n = n;
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_soa_free(ll_ctx *ctx, ll_soa_t *soa) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!soa) return LL_OK;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (soa->planes) cudaFree(soa->planes);
    if (soa->meta) cudaFree(soa->meta);
    memset(soa, 0, sizeof *soa);
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_soa_upload(ll_ctx *ctx, const ll_world_t *worlds, size_t n, ll_soa_t *dst) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!dst || (n && !worlds)) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (n > dst->stride) return fail(LL_ERR_CAPACITY, "%zu positions do not fit stride %zu", n, dst->stride);
    dst->n = n;
    if (!n) return LL_OK;
    LL_TRY(ensure(ctx, ctx->aos_in, n * sizeof(ll_world_t)));
    CUDA_TRY(cudaMemcpyAsync(ctx->aos_in.ptr, worlds, n * sizeof(ll_world_t), cudaMemcpyHostToDevice, ctx->stream));
    {
        Timed t(ctx, "aos_to_soa");
        k_aos_to_soa<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>((const u64 *)ctx->aos_in.ptr, n, Soa{(u64 *)dst->planes, dst->meta, dst->stride}, 0);
    }
    LL_TRY(check_launch("k_aos_to_soa"));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_soa_download(ll_ctx *ctx, const ll_soa_t *src, size_t first, size_t n, ll_world_t *worlds) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!src || (n && !worlds)) return fail(LL_ERR_INVALID_ARG, "null argument");
    if (first + n > src->n) return fail(LL_ERR_INVALID_ARG, "range %zu+%zu exceeds %zu positions", first, n, src->n);
    if (!n) return LL_OK;
    LL_TRY(ensure(ctx, ctx->aos_out, n * sizeof(ll_world_t)));
    {
        Timed t(ctx, "soa_to_aos");
        k_soa_to_aos<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>(Soa{(u64 *)src->planes, src->meta, src->stride}, first, n, (u64 *)ctx->aos_out.ptr);
    }
    LL_TRY(check_launch("k_soa_to_aos"));
    CUDA_TRY(cudaMemcpyAsync(worlds, ctx->aos_out.ptr, n * sizeof(ll_world_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LL_OK;
    LL_ABI_END
}

extern "C" int ll_playout_soa(ll_ctx *ctx, uint64_t seed, uint64_t first_index, size_t n, ll_soa_t *dst) {
    LL_ABI_BEGIN
    LL_TRY(check_ctx(ctx));
    if (!dst) return fail(LL_ERR_INVALID_ARG, "null soa");
    if (n > dst->stride) return fail(LL_ERR_CAPACITY, "%zu positions do not fit stride %zu", n, dst->stride);
    dst->n = n;
    if (!n) return LL_OK;
    {
        Timed t(ctx, "playout");
        k_playout<<<(unsigned)nblocks(n), BLOCK, 0, ctx->stream>>>(seed, first_index, n, Soa{(u64 *)dst->planes, dst->meta, dst->stride});
    }
    return check_launch("k_playout");
    LL_ABI_END
}
#include "ll_multi.inl"
