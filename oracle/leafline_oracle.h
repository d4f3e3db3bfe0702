/*
 * leafline_oracle.h -- CPU restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for leafline_b200.  It restates, in plain C, the algorithm of the
 * reference (zackmdavis/Leafline, Rust) for move generation (`src/life.rs`), static scoring
 * (`src/mind.rs:26-143`) and the leaf rule of the search (`src/mind.rs:279-287`), keeping the
 * reference's cost structure (make-move-then-full-opponent-movegen legality, attack probes by
 * full movegen) so that it can also be timed as the "reference algorithm on CPU" baseline.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this code.  The product (leafline_b200/) never calls it and has no CPU fallback.
 *
 * Parity pinning: the reference is Rust (nightly) and cannot be compiled in this image (no rustc /
 * cargo), so there is no oracle/_ref.  The restatement is pinned by every known-answer test the
 * reference holds for this path; tests/test_oracle_golden.py checks
 * them all.  Each function below cites the reference file:line it follows.
 */
#ifndef LEAFLINE_ORACLE_H
#define LEAFLINE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* identity.rs:8-11, 36-43 */
enum { LO_ORANGE = 0, LO_BLUE = 1 };
enum { LO_SERVANT = 0, LO_PONY = 1, LO_SCHOLAR = 2, LO_COP = 3, LO_PRINCESS = 4, LO_FIGUREHEAD = 5 };
#define LO_NONE 0xFF
#define LO_AGENT(team, job) ((uint8_t)(((team) << 3) | (job)))

/* life.rs:178-181 */
#define LO_ORANGE_WEST 0x1
#define LO_ORANGE_EAST 0x2
#define LO_BLUE_WEST 0x4
#define LO_BLUE_EAST 0x8

/* life.rs:157-176 WorldState.  plane[6*team + job] in the reference's field order.
 * passing_by: 0xFF = None, else rank<<4|file exactly as space.rs:34 packs a Locale. */
typedef struct {
    uint64_t plane[12];
    uint8_t initiative;
    uint8_t service_eligibility;
    uint8_t passing_by;
    uint8_t pad[5];
} lo_world_t; /* 104 bytes */

/* life.rs:15-20 Patch, life.rs:77-83 Commit.  whence/whither are Locales (rank<<4|file).
 * hospitalization / ascension: 0xFF = None, else LO_AGENT(team, job). */
typedef struct {
    uint8_t team, job, whence, whither;
    uint8_t hospitalization, ascension;
    uint8_t pad[2];
    lo_world_t tree;
} lo_commit_t; /* 112 bytes */

#define LO_MAX_COMMITS 512

/* error convention: functions that can hit a reference panic!() return LO_PANIC (< 0). */
#define LO_PANIC (-1)

void lo_world_default(lo_world_t *w);                             /* life.rs:184-215 */
void lo_world_empty(lo_world_t *w);                               /* life.rs:223-242 */
int lo_reconstruct(const char *scan, lo_world_t *w);              /* life.rs:428-507 */
int lo_preserve(const lo_world_t *w, char *book, size_t cap);     /* life.rs:293-360 */

int lo_apply(const lo_world_t *w, uint8_t team, uint8_t job, uint8_t whence, uint8_t whither,
             lo_commit_t *out);                                   /* life.rs:569-676 */
int lo_in_critical_endangerment(const lo_world_t *w, int team);   /* life.rs:678-690 */
int lo_is_being_leered_at_by(const lo_world_t *w, uint8_t locale, int team); /* life.rs:284-291 */

/* per-piece generators (life.rs:749-1050); each appends to out[*n], returns 0 or LO_PANIC */
int lo_servant_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);
int lo_pony_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);
int lo_scholar_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);
int lo_cop_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);
int lo_princess_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);
int lo_figurehead_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);
int lo_service_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n);

/* life.rs:1052-1084.  out must hold LO_MAX_COMMITS entries; returns the count or LO_PANIC. */
int lo_lookahead(const lo_world_t *w, lo_commit_t *out);
int lo_reckless_lookahead(const lo_world_t *w, lo_commit_t *out);

float lo_orientation(int team);                                   /* mind.rs:29-34 */
float lo_score(const lo_world_t *w);                              /* mind.rs:50-143 */
void lo_score_batch(const lo_world_t *worlds, size_t n, float *out);

/* perft = recursive count of lookahead() leaves; depth 0 -> 1.
 * lo_perft_divide also fills per_root[i] for root move i (canonical order), returns the root count.
 * threads > 1 runs one worker per root move, the reference's own parallel model (mind.rs:399-414). */
uint64_t lo_perft(const lo_world_t *w, int depth);
int lo_perft_divide(const lo_world_t *w, int depth, uint64_t *per_root, int cap, int threads);

/* TT-free exact negamax with the leaf rule of mind.rs:279-287 (quiet = None): value of `w` for the
 * side to move, searching `depth` plies of reckless children. */
float lo_negamax(const lo_world_t *w, int depth);
/* the same with the quiescence extension of mind.rs:284-301: quiet < 0 is None, else Some(quiet) */
float lo_negamax_quiet(const lo_world_t *w, int depth, int quiet);
/* mind.rs:375-448 kickoff without TT/history/threads: root moves = lookahead() (or reckless if
 * nihilistically), score[i] = -negamax(child_i, depth-1).  Canonical (unsorted) order. */
int lo_kickoff(const lo_world_t *w, int depth, int nihilistically, lo_commit_t *roots, float *scores,
               int threads);
int lo_kickoff_quiet(const lo_world_t *w, int depth, int quiet, int nihilistically, lo_commit_t *roots, float *scores,
                     int threads);

/* Seeded random-playout positions (defined by this build). */
uint64_t lo_mix64(uint64_t z);
int lo_playout(uint64_t seed, uint64_t index, lo_world_t *out, int *plies_out);

/* generated tables (furniture/tablemaker.py:41-62, 91-109) */
extern uint64_t LO_PONY_MOVEMENT_TABLE[64];
extern uint64_t LO_FIGUREHEAD_MOVEMENT_TABLE[64];
uint64_t lo_center_of_the_world(void);
uint64_t lo_forward_contour(int rank);
uint64_t lo_sideways_contour(int file);

#ifdef __cplusplus
}
#endif
#endif
