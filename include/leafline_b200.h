/*
 * leafline_b200.h -- C ABI of the B200-native leaf layer (movegen + static scoring) for Leafline.
 *
 * The reference (zackmdavis/Leafline) has no FFI layer: its boundary for this path is the Rust
 * method surface of `WorldState` / `Patch` / `Commit` (src/life.rs) and `mind::score` /
 * `kickoff` (src/mind.rs).  Each entry point below names the reference item it replaces; the
 * Rust-side `extern "C"` shim a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; the caller owns every host buffer, the library owns device memory,
 *    streams and scratch between ll_init() and ll_shutdown(); no pointer is retained after return.
 *  - every call returns LL_OK (0) or a negative LL_ERR_*; ll_last_error() gives a thread-local
 *    message.  Nothing unwinds or aborts across this boundary (the reference panics instead:
 *    src/sorceries.rs:2-4).
 *  - there is NO CPU fallback: without a CUDA device ll_init() fails with LL_ERR_NO_DEVICE.
 *  - an ll_ctx is NOT thread-safe; use one ll_ctx per host thread (the reference calls the leaf
 *    layer from one thread per root move, src/mind.rs:405).
 *  - DOMAINS.  Every position the reference answers and this ABI can spell is answered identically:
 *      domain W (the fast path: set-wise kernels, legality by check / pin / danger masks): no two planes overlap; at most one
 *        figurehead and at most 16 figurines per team; a set service-eligibility bit implies that team's figurehead, if any,
 *        stands on its home e-square, and that the cop stands on that corner or the team has at most 15 figurines (secret
 *        service conjures a cop where none stands, src/life.rs:621-628); passing_by is None or an empty square directly behind
 *        an enemy servant on its double-step rank.  Every state reachable from a state of W by the reference's own move
 *        generator is in W (tests/test_domain.py), so a tree rooted in W never leaves the fast path.
 *      the literal domain (any number of figureheads and figurines, any eligibility -- the reference loops over every
 *        figurehead, src/life.rs:678-690, and `--from` takes any runes, src/main.rs:199): answered by the literal path, the
 *        canonical generator with the reference's own legality (every commit made, kept iff no figurehead of the mover is
 *        endangered afterwards, src/life.rs:692-699), level by level.  Same results as the reference, at a fraction of the
 *        fast path's speed; a call takes it as soon as one of its positions is outside W.
 *      rejected with LL_ERR_DOMAIN: overlapping planes, initiative / eligibility bytes out of range, a passing_by square that
 *        is off the board, occupied or not behind an enemy servant (the reference would panic or answer nonsense).
 *    ll_score_batch / ll_score_soa / ll_soa_upload do not classify their input: mind::score is defined on any twelve planes.
 */
#ifndef LEAFLINE_B200_H
#define LEAFLINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LL_OK 0
#define LL_ERR_INVALID_ARG (-1)
#define LL_ERR_CUDA (-2)
#define LL_ERR_CAPACITY (-3) /* an output buffer is too small; nothing is written past its capacity */
#define LL_ERR_DOMAIN (-4)   /* a position cannot be answered at all (see DOMAINS above) */
#define LL_ERR_NO_DEVICE (-5)
#define LL_ERR_OOM (-6)
#define LL_ERR_NCCL (-7)     /* NCCL is missing (it is loaded at run time, by the multi-GPU calls only) or reported an error */

/* src/identity.rs:8-11, 36-43 */
enum { LL_ORANGE = 0, LL_BLUE = 1 };
enum { LL_SERVANT = 0, LL_PONY = 1, LL_SCHOLAR = 2, LL_COP = 3, LL_PRINCESS = 4, LL_FIGUREHEAD = 5 };
#define LL_NONE 0xFF
#define LL_AGENT(team, job) ((uint8_t)(((team) << 3) | (job)))

/* replaces `WorldState` (src/life.rs:157-176).  plane[6*team + job] follows the reference's field
 * order (orange servants .. orange figurehead, blue servants .. blue figurehead); bit 8*rank+file,
 * a1 = bit 0 (src/space.rs:65-71).  service_eligibility bits as src/life.rs:178-181.
 * passing_by: 0xFF = None, else rank<<4|file exactly as src/space.rs:34 packs a Locale. */
typedef struct ll_world_t {
    uint64_t plane[12];
    uint8_t initiative;
    uint8_t service_eligibility;
    uint8_t passing_by;
    uint8_t pad[5];
} ll_world_t; /* 104 bytes */

/* replaces `Patch` (src/life.rs:15-20) + `Commit` (src/life.rs:77-83).  whence / whither are
 * Locales (rank<<4|file); hospitalization / ascension are 0xFF = None, else LL_AGENT(team, job). */
typedef struct ll_commit_t {
    uint8_t team, job, whence, whither;
    uint8_t hospitalization, ascension;
    uint8_t pad[2];
    ll_world_t tree;
} ll_commit_t; /* 112 bytes */

typedef struct ll_ctx ll_ctx;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int ll_init(int device, ll_ctx **ctx); /* one context per (host thread, CUDA device) */
int ll_shutdown(ll_ctx *ctx);
const char *ll_last_error(void);
/* launch on the caller's stream (a cudaStream_t / CUstream, e.g. torch.cuda.current_stream().cuda_stream);
 * NULL restores the context's own stream. */
int ll_set_stream(ll_ctx *ctx, void *cuda_stream);
int ll_synchronize(ll_ctx *ctx);
int ll_device_sm_count(ll_ctx *ctx);

/* ---- batched leaf layer, HOST buffers (the drop-in calls) ---------------------------------- */

/* replaces `mind::score(WorldState) -> f32` (src/mind.rs:50-143), n positions at once. */
int ll_score_batch(ll_ctx *ctx, const ll_world_t *worlds, size_t n, float *scores);

/* replaces `WorldState::lookahead()` (src/life.rs:1078) when nihilistically == 0 and
 * `WorldState::reckless_lookahead()` (src/life.rs:1082) otherwise.  For world i its commits are
 * commits[offsets[i] .. offsets[i+1]) in the reference's order (SURVEY.md 8(a) "canonical move
 * order").  offsets has n+1 entries; commits may be NULL to obtain only the counts. */
int ll_lookahead_batch(ll_ctx *ctx, const ll_world_t *worlds, size_t n, int nihilistically, uint32_t *offsets,
                       ll_commit_t *commits, size_t commits_cap);

/* replaces `WorldState::in_critical_endangerment(team)` (src/life.rs:678-690) for team[i] on world i. */
int ll_in_critical_endangerment_batch(ll_ctx *ctx, const ll_world_t *worlds, const uint8_t *teams, size_t n,
                                      uint8_t *endangered);

/* perft = number of leaves of the legal-move tree (recursive `lookahead()`);
 * depth 0 -> 1.  per_root (nullable, root_cap entries) receives the leaf count below each root
 * move in canonical order and *n_roots their number ("divide").  With shard_world > 1 (<= 32) only this rank's
 * work units are counted: the units are the nodes of ply min(3, depth - 2), each named by its parent and its index c
 * among the parent's commits in canonical order, owner = (hash(parent) + c) mod shard_world -- a pure function of the
 * tree, so no rank needs to hear from another (trees of depth <= 3 are not dealt: rank 0 counts them whole).  Summing
 * the results of all ranks (a u64 allreduce; ll_perft_multi does it on the devices) gives the full answer.  The frontier is materialised in HBM up to ply depth - 2 (100 B per node + 12 B of ids and
 * counts); the last two plies are generated and counted on chip.  LL_ERR_OOM if a level does not fit. */
int ll_perft(ll_ctx *ctx, const ll_world_t *root, int depth, int shard_rank, int shard_world, uint64_t *leaves,
             uint64_t *per_root, uint32_t root_cap, uint32_t *n_roots);

/* ll_perft without the host round trip, for callers that reduce the counts on the device (NCCL): the whole
 * launch chain is enqueued on the context's stream and the results are left in DEVICE memory,
 * d_out[LL_PERFT_DEVICE_WORDS] u64: [0] leaves, [1] number of root moves, [2] non-zero if a level outgrew
 * its buffer (the other words are then meaningless: call ll_perft, which re-sizes), [8 + i] leaves below root
 * move i.  Asynchronous.  depth >= 3; needs one earlier ll_perft call of this depth (LL_ERR_CAPACITY otherwise). */
#define LL_PERFT_DEVICE_WORDS (8 + 1024)
int ll_perft_device(ll_ctx *ctx, const ll_world_t *root, int depth, int shard_rank, int shard_world, uint64_t *d_out);

/* replaces the leaf / horizon rule of `α_β_negamax_search` (src/mind.rs:279-301) for n frontier nodes:
 * values[i] = exact negamax value of frontier[i] for its side to move over `plies` plies of
 * reckless_lookahead() children with orientation * score leaves.  quiet_extension < 0 is `quiet = None`;
 * quiet_extension = e >= 0 is `quiet = Some(e)`: below depth 0 the search continues for up to e plies over the
 * commits that hospitalize somebody, every such node being free to stand pat (src/mind.rs:288-300). */
int ll_horizon_batch(ll_ctx *ctx, const ll_world_t *frontier, size_t n, int plies, int quiet_extension, float *values);

/* replaces `kickoff(world, depth, extension, nihilistically, ..)` (src/mind.rs:375-448) without TT / history /
 * threads (quiet_extension as in ll_horizon_batch): roots = lookahead()
 * (or reckless_lookahead() if nihilistically), scores[i] = -value(child_i, depth-1), canonical
 * (unsorted) order.  With shard_world > 1 only roots i % shard_world == shard_rank are searched;
 * the others get -INFINITY (gather with a max). */
int ll_kickoff(ll_ctx *ctx, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, int shard_rank, int shard_world,
               ll_commit_t *roots, float *scores, uint32_t root_cap, uint32_t *n_roots, uint64_t *leaves_scored);

/* ---- several GPUs sharing ONE tree (replaces the reference's fan-out inside the engine: a thread per root move and an
 * mpsc gather, src/mind.rs:399-437).  A group owns one ll_ctx per GPU, the NCCL communicator between them and the
 * device-resident reduction of their results.  Subtrees are dealt to the ranks, no position ever crosses NVLink; the only
 * exchange is a u64 sum of leaf counts or an element-wise max of per-parent f32 keys, done on the devices -- by NCCL, or by
 * this library's own kernel over NVLink peer memory where every rank can map every other's (ll_multi_reduce_path).
 *   one process, n GPUs:  ll_init_multi -- ncclCommInitAll, one persistent host thread and one stream per device; one call
 *                         of ll_*_multi drives them all.
 *   one process per GPU:  ll_multi_unique_id on one rank, the 128 bytes carried to every process by the caller (torchrun,
 *                         MPI, a pipe), ll_init_multi_rank everywhere; ll_*_multi is then COLLECTIVE: every rank calls it with
 *                         the same arguments and every rank gets the whole answer.
 * A group is used from one host thread at a time.  Errors on one rank are voted on inside the collective: every rank returns
 * an error, none hangs. */
typedef struct ll_mctx ll_mctx;
#define LL_NCCL_UNIQUE_ID_BYTES 128
int ll_multi_unique_id(void *id128);
int ll_init_multi(const int *devices, int n, ll_mctx **group);
int ll_init_multi_rank(int device, int rank, int world, const void *unique_id128, ll_mctx **group);
int ll_shutdown_multi(ll_mctx *group);
int ll_multi_world(ll_mctx *group);        /* ranks sharing a tree */
int ll_multi_local_count(ll_mctx *group);  /* members living in this process */
ll_ctx *ll_multi_ctx(ll_mctx *group, int local_index); /* a member's context, e.g. for ll_set_stream / ll_timing_* */
int ll_multi_reduce_path(ll_mctx *group);  /* 0 = NCCL, 1 = the library's reduction over NVLink peer memory */
/* trees shallower than these are not dealt: every rank answers them whole and no collective is paid for (defaults 5 and 6;
 * 0 keeps a value).  Collective: set the same values on every rank. */
int ll_multi_set_shard_min_depth(ll_mctx *group, int perft_depth, int kickoff_depth);

/* ll_perft over the group: ONE tree, its ply-min(3, depth - 2) nodes dealt to the ranks (the rule of ll_perft's
 * shard_rank / shard_world), counts summed on the devices.  per_root / n_roots as in ll_perft. */
int ll_perft_multi(ll_mctx *group, const ll_world_t *root, int depth, uint64_t *leaves, uint64_t *per_root, uint32_t root_cap, uint32_t *n_roots);
/* perft of n independent roots (<= 1024), root i counted by rank i % world: leaves[i] for every i on every rank. */
int ll_perft_batch_multi(ll_mctx *group, const ll_world_t *roots, uint32_t n, int depth, uint64_t *leaves);
/* ll_kickoff over the group (src/mind.rs:399-437 on n GPUs): the levels down to ply min(3, depth - 1) are expanded in canonical
 * order on every rank, the nodes of that ply dealt round-robin, searched, their values folded into their parents' keys,
 * max-reduced on the devices and backed up: scores[i] are the exact per-root scores of ll_kickoff, on every rank. */
int ll_kickoff_multi(ll_mctx *group, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, ll_commit_t *roots, float *scores,
                     uint32_t root_cap, uint32_t *n_roots, uint64_t *leaves_scored);

/* ---- the kickoff family as the reference's callers see it (src/mind.rs:441-490) ----------------------- */
#define LL_MAX_VARIATION 16
typedef struct ll_patch_t { /* `Patch` (src/life.rs:15-20) */
    uint8_t team, job, whence, whither;
} ll_patch_t;
/* one element of the `Vec<(Commit, f32, Variation)>` the kickoff functions return */
typedef struct ll_forecast_t {
    ll_commit_t commit; /* the root move */
    float score;
    uint32_t variation_len;                  /* >= 1: variation[0] is the root move's patch (mind.rs:426-427) */
    ll_patch_t variation[LL_MAX_VARIATION];  /* principal variation; ties go to the first commit in canonical order */
} ll_forecast_t;

/* replaces `kickoff::<Variation>(world, depth, extension, nihilistically, _)` (src/mind.rs:441-448): forecasts
 * sorted by score, best first (mind.rs:436).  The search is TT-free and full-width on the device. */
int ll_forecast(ll_ctx *ctx, const ll_world_t *root, int depth, int quiet_extension, int nihilistically, ll_forecast_t *forecasts,
                uint32_t cap, uint32_t *n_forecasts, uint64_t *leaves_scored);
/* the same forecasts with the reference's search STRUCTURE (src/mind.rs:145-169, 271-364): negamax with alpha-beta
 * cut-offs, MVV-LVA + history ordering and a full window below every root move run on the HOST; a node with
 * `device_plies` plies left is handed to the device, which returns its exact full-width value (quiet_extension as in
 * ll_horizon_batch).  Scores equal ll_forecast's (alpha-beta is exact at the root); only the work is pruned, so
 * depths beyond ll_forecast's HBM limit are reachable.  The memory bank (mind.rs:312-344) is sound here: entries are keyed
 * by the whole position and the plies left and say whether their value is exact or a bound, so remembering changes the
 * work, never a root score; it forgets the least recently used entry under a byte budget (mind.rs:367-372;
 * ll_set_deja_vu_bound).  variation is the principal variation (mind.rs:208-224): the bank's best commits through the plies
 * the host searched, the device's variation below the frontier node.
 * stats3 (nullable): { nodes expanded on the host, frontier nodes handed to the device, values taken from the bank }. */
int ll_alpha_beta_forecast(ll_ctx *ctx, const ll_world_t *root, int depth, int device_plies, int quiet_extension, int nihilistically,
                           ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts, uint64_t *stats3);
/* ll_alpha_beta_forecast over the group: the root moves are searched concurrently, root move i on rank i % world (the
 * reference searches every root move on a thread of its own, src/mind.rs:399-414); scores and variations are gathered in
 * one max-reduction on the devices.  stats3 are summed over the ranks. */
int ll_alpha_beta_forecast_multi(ll_mctx *group, const ll_world_t *root, int depth, int device_plies, int quiet_extension, int nihilistically,
                                 ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts, uint64_t *stats3);
/* searcher threads of the host alpha-beta driver: the reference searches every root move on a thread of its own
 * (src/mind.rs:399-414); so does ll_alpha_beta_forecast, up to this many threads (default 32; 1 = a single depth-first walk).
 * The threads' device requests are combined into one batch a round, so more threads mean fewer, larger hand-offs. */
int ll_set_search_threads(ll_ctx *ctx, int threads);
/* the `déjà_vu_bound` argument of the reference's kickoff functions (GiB of memory bank, src/mind.rs:367-372, 441-448) for the
 * searches of this context; default 0.25 */
int ll_set_deja_vu_bound(ll_ctx *ctx, double gib);
/* replaces `fixed_depth_sequence_kickoff` (src/mind.rs:473-490) */
int ll_fixed_depth_sequence_forecast(ll_ctx *ctx, const ll_world_t *root, const uint8_t *depths, uint32_t n_depths, int nihilistically,
                                     ll_forecast_t *forecasts, uint32_t cap, uint32_t *n_forecasts);
/* replaces `iterative_deepening_kickoff` (src/mind.rs:451-469): the deepest iteration that finished inside the timeout.
 * Iterations are full-width on the device (ll_forecast) while the tree fits HBM and depth < 8, then the host alpha-beta
 * driver (ll_alpha_beta_forecast's search, abandoned in mid-iteration at the deadline): same scores either way.  One
 * intuition bank serves all the iterations (mind.rs:456-464) and, like the reference, the call uses its whole budget. */
int ll_iterative_deepening_forecast(ll_ctx *ctx, const ll_world_t *root, double timeout_seconds, int nihilistically, ll_forecast_t *forecasts,
                                    uint32_t cap, uint32_t *n_forecasts, uint32_t *depth_reached);

/* ---- device-resident positions (structure-of-arrays u64 planes in HBM) ---------------------- */

/* 12 planes of `stride` u64 each (plane j of position i at planes[j*stride + i]; stride is a
 * multiple of 2 so every plane is 16-byte aligned) + one u32 of meta per position
 * (bit 0 initiative, bits 1-4 service_eligibility, bits 8-15 passing_by).  All DEVICE pointers. */
typedef struct ll_soa_t {
    uint64_t *planes;
    uint32_t *meta;
    size_t stride;
    size_t n;
} ll_soa_t;

int ll_soa_alloc(ll_ctx *ctx, size_t n, ll_soa_t *soa);
int ll_soa_free(ll_ctx *ctx, ll_soa_t *soa);
int ll_soa_upload(ll_ctx *ctx, const ll_world_t *worlds, size_t n, ll_soa_t *dst);          /* H2D + repack */
int ll_soa_download(ll_ctx *ctx, const ll_soa_t *src, size_t first, size_t n, ll_world_t *worlds);
/* scores[i] for the positions already resident in HBM; d_scores is a DEVICE pointer. Asynchronous. */
int ll_score_soa(ll_ctx *ctx, const ll_soa_t *soa, float *d_scores);
/* horizon values for resident positions; d_values is a DEVICE pointer. */
int ll_horizon_soa(ll_ctx *ctx, const ll_soa_t *soa, int plies, int quiet_extension, float *d_values, uint64_t *leaves_scored);
/* seeded random-playout positions first_index .. first_index+n-1 of the set defined in DESIGN.md
 * (SURVEY.md 8(d) config 3), generated on the device. */
int ll_playout_soa(ll_ctx *ctx, uint64_t seed, uint64_t first_index, size_t n, ll_soa_t *dst);

/* ---- per-kernel device timing (CUDA events on the launching stream) ------------------------- */
int ll_timing_enable(ll_ctx *ctx, int on);
int ll_timing_reset(ll_ctx *ctx);
/* accumulated milliseconds and launch count of the named kernel since the last reset */
int ll_timing_query(ll_ctx *ctx, const char *kernel, double *ms, uint64_t *launches);
/* total number of this library's kernel launches since ll_init (for bench.py's gpu_launches) */
uint64_t ll_launch_count(ll_ctx *ctx);
/* testing aid: ll_perft normally replays its launch chain (device-resident level sizes) as a captured CUDA graph.
 * on == 1 forces the level-by-level path (every level sized on the host), on == 2 issues the chain launch by launch;
 * 0 restores the default.  All three give identical results.  on == 1 also makes ll_horizon_batch take its level-by-level
 * path for small batches (by default <= 4096 nodes with plies >= 2 and no quiet extension run as one launch chain). */
int ll_debug_force_synced(ll_ctx *ctx, int on);
/* testing aid, full-size property check on resident positions: for every position the set-wise move counter, the
 * canonical generator (legality by check / pin / danger masks) and the reference's definition (each reckless commit
 * made, kept iff the mover's figurehead is not endangered afterwards, src/life.rs:692-699) must agree.
 * result4 = { disagreeing positions, first disagreeing index (~0 if none), legal commits, reckless commits }. */
int ll_debug_selfcheck_soa(ll_ctx *ctx, const ll_soa_t *soa, uint64_t *result4);
/* testing aid: the two generated movement tables as they are resident on the device (64 entries each), to be compared
 * with the output of the reference's generator (furniture/tablemaker.py:41-62; tests/golden/reference_tables.json). */
int ll_debug_tables(ll_ctx *ctx, uint64_t *pony64, uint64_t *figurehead64);
/* testing aid: how many parents of the LAST level-by-level horizon call (ll_horizon_batch with the chain forced off, or
 * ll_horizon_soa) had more records with stunning / ascending commits than a tile of the horizon kernel keeps, and were
 * valued by the overflow kernel instead -- so that tests can tell that path ran. */
int ll_debug_horizon_overflowed(ll_ctx *ctx, uint32_t *parents);
/* measurement utility for the integer roofline: sustained thread-level integer instructions per second
 * of a dependency-free stream; kind 0 = LOP3 only (ALU pipe), 1 = LOP3 + IMAD (ALU + FMA pipes) */
int ll_measure_int_peak(ll_ctx *ctx, int kind, double *thread_ops_per_sec);

#ifdef __cplusplus
}
#endif
#endif
