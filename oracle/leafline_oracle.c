/*
 * leafline_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY
 * (see leafline_oracle.h).  Every function cites the reference file:line whose behaviour it
 * restates; the control flow keeps the reference's cost structure on purpose (legality =
 * make the move, then run the opponent's full pseudo-legal movegen, life.rs:678-699; attack
 * probes = full movegen on a board with an extra figurehead, life.rs:284-291).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (see oracle/Makefile).  -ffp-contract=off
 * keeps lo_score() to individually rounded f32 mul/add, which is what rustc emits for mind.rs:50-143.
 */
#include "leafline_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- space.rs */

/* space.rs:34 -- a Locale packs rank<<4 | file into a u8 */
static inline uint8_t loc_new(int rank, int file) { return (uint8_t)((rank << 4) | file); }
static inline int loc_rank(uint8_t l) { return l >> 4; }              /* space.rs:117-119 */
static inline int loc_file(uint8_t l) { return l & 0xF; }             /* space.rs:121-123 */
static inline int loc_pindex(uint8_t l) { return 8 * loc_rank(l) + loc_file(l); } /* space.rs:65-67 */
static inline uint64_t loc_pinpoint(uint8_t l) { return 1ull << loc_pindex(l); }  /* space.rs:69-71 */

/* space.rs:82-97 displace (and 99-115 multidisplace with the factor already multiplied in):
 * the i8 -> u8 casts and the u8 shift are restated so that off-board results are rejected the
 * same way (is_legal: rank() < 8 && file() < 8, space.rs:73-75). */
static inline int loc_displace(uint8_t l, int dr, int df, uint8_t *out) {
    uint8_t r = (uint8_t)(int8_t)(loc_rank(l) + dr);
    uint8_t f = (uint8_t)(int8_t)(loc_file(l) + df);
    uint8_t raf = (uint8_t)((uint8_t)(r << 4) | f);
    if ((raf >> 4) < 8 && (raf & 0xF) < 8) {
        *out = raf;
        return 1;
    }
    return 0;
}

static inline int pincount(uint64_t bits) { return __builtin_popcountll(bits); } /* space.rs:197-200 */

/* space.rs:181-195 to_locales: ascending pindex scan of all 64 bits */
static int to_locales(uint64_t bits, uint8_t *locales) {
    int n = 0;
    uint64_t bitfield = 1;
    for (int rank = 0; rank < 8; rank++) {
        for (int file = 0; file < 8; file++) {
            if (bitfield & bits) locales[n++] = loc_new(rank, file);
            bitfield <<= 1;
        }
    }
    return n;
}

/* ---------------------------------------------------------------- generated tables */

uint64_t LO_PONY_MOVEMENT_TABLE[64];
uint64_t LO_FIGUREHEAD_MOVEMENT_TABLE[64];

/* furniture/tablemaker.py:41-62 */
static void universal_distribution(const int (*options)[2], int n_options, uint64_t *table) {
    for (int rank = 0; rank < 8; rank++) {
        for (int file = 0; file < 8; file++) {
            uint64_t acc = 0;
            for (int i = 0; i < n_options; i++) {
                int r = rank + options[i][0], f = file + options[i][1];
                if (r >= 0 && r < 8 && f >= 0 && f < 8) acc |= 1ull << (8 * r + f);
            }
            table[8 * rank + file] = acc;
        }
    }
}

__attribute__((constructor)) static void lo_init_tables(void) {
    static const int pony[8][2] = {{+1, +2}, {-1, +2}, {+1, -2}, {-1, -2},
                                   {+2, +1}, {-2, +1}, {+2, -1}, {-2, -1}};
    int fig[8][2];
    int k = 0;
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++)
            if (!(i == 0 && j == 0)) {
                fig[k][0] = i;
                fig[k][1] = j;
                k++;
            }
    universal_distribution(pony, 8, LO_PONY_MOVEMENT_TABLE);
    universal_distribution((const int(*)[2])fig, 8, LO_FIGUREHEAD_MOVEMENT_TABLE);
}

/* tablemaker.py:23-26 */
uint64_t lo_center_of_the_world(void) {
    uint64_t acc = 0;
    for (int i = 2; i < 6; i++)
        for (int j = 2; j < 6; j++) acc |= 1ull << (8 * i + j);
    return acc;
}
/* tablemaker.py:29-32 */
uint64_t lo_forward_contour(int rank) {
    uint64_t acc = 0;
    for (int j = 0; j < 8; j++) acc |= 1ull << (8 * rank + j);
    return acc;
}
/* tablemaker.py:35-38 */
uint64_t lo_sideways_contour(int file) {
    uint64_t acc = 0;
    for (int i = 0; i < 8; i++) acc |= 1ull << (8 * i + file);
    return acc;
}

/* ---------------------------------------------------------------- WorldState basics */

static inline int opposition(int team) { return team ^ 1; }          /* identity.rs:20-25 */
static inline uint64_t *plane_of(lo_world_t *w, int team, int job) { return &w->plane[6 * team + job]; }
static inline uint64_t cplane(const lo_world_t *w, int team, int job) { return w->plane[6 * team + job]; }

void lo_world_empty(lo_world_t *w) { /* life.rs:223-242 */
    memset(w, 0, sizeof *w);
    w->initiative = LO_ORANGE;
    w->service_eligibility = 0;
    w->passing_by = LO_NONE;
}

void lo_world_default(lo_world_t *w) { /* life.rs:184-215 */
    lo_world_empty(w);
    for (int f = 0; f < 8; f++) {
        *plane_of(w, LO_ORANGE, LO_SERVANT) |= loc_pinpoint(loc_new(1, f));
        *plane_of(w, LO_BLUE, LO_SERVANT) |= loc_pinpoint(loc_new(6, f));
    }
    *plane_of(w, LO_ORANGE, LO_PONY) = loc_pinpoint(loc_new(0, 1)) | loc_pinpoint(loc_new(0, 6));
    *plane_of(w, LO_ORANGE, LO_SCHOLAR) = loc_pinpoint(loc_new(0, 2)) | loc_pinpoint(loc_new(0, 5));
    *plane_of(w, LO_ORANGE, LO_COP) = loc_pinpoint(loc_new(0, 0)) | loc_pinpoint(loc_new(0, 7));
    *plane_of(w, LO_ORANGE, LO_PRINCESS) = loc_pinpoint(loc_new(0, 3));
    *plane_of(w, LO_ORANGE, LO_FIGUREHEAD) = loc_pinpoint(loc_new(0, 4)); /* space.rs:20 */
    *plane_of(w, LO_BLUE, LO_PONY) = loc_pinpoint(loc_new(7, 1)) | loc_pinpoint(loc_new(7, 6));
    *plane_of(w, LO_BLUE, LO_SCHOLAR) = loc_pinpoint(loc_new(7, 2)) | loc_pinpoint(loc_new(7, 5));
    *plane_of(w, LO_BLUE, LO_COP) = loc_pinpoint(loc_new(7, 0)) | loc_pinpoint(loc_new(7, 7));
    *plane_of(w, LO_BLUE, LO_PRINCESS) = loc_pinpoint(loc_new(7, 3));
    *plane_of(w, LO_BLUE, LO_FIGUREHEAD) = loc_pinpoint(loc_new(7, 4)); /* space.rs:21 */
    w->service_eligibility = 0xF;
    w->passing_by = LO_NONE;
}

/* life.rs:521-540 */
static uint64_t occupied_by(const lo_world_t *w, int team) {
    uint64_t acc = 0;
    for (int job = 0; job < 6; job++) acc |= cplane(w, team, job);
    return acc;
}
static uint64_t occupied(const lo_world_t *w) { return occupied_by(w, LO_ORANGE) | occupied_by(w, LO_BLUE); } /* 542-544 */
static uint64_t unoccupied(const lo_world_t *w) { return ~occupied(w); }                                        /* 546-548 */
static inline int query(uint64_t bits, uint8_t station) { return (bits & loc_pinpoint(station)) != 0; }        /* space.rs:175-179 */

/* life.rs:550-557: first agent of `team`, in dramatis_personae order (identity.rs:51-117), standing on `at` */
static uint8_t occupying_affiliated_agent(const lo_world_t *w, uint8_t at, int team) {
    for (int job = 0; job < 6; job++)
        if (query(cplane(w, team, job), at)) return LO_AGENT(team, job);
    return LO_NONE;
}
/* life.rs:559-567 */
static uint8_t occupying_agent(const lo_world_t *w, uint8_t at) {
    for (int team = 0; team < 2; team++) {
        uint8_t a = occupying_affiliated_agent(w, at, team);
        if (a != LO_NONE) return a;
    }
    return LO_NONE;
}

static uint8_t eligibility_bit(int team, int west) { /* life.rs:379-386 */
    if (team == LO_ORANGE) return west ? LO_ORANGE_WEST : LO_ORANGE_EAST;
    return west ? LO_BLUE_WEST : LO_BLUE_EAST;
}

/* ---------------------------------------------------------------- preservation runes (FEN-like) */

static const char RUNES[2][6] = {{'P', 'N', 'B', 'R', 'Q', 'K'}, {'p', 'n', 'b', 'r', 'q', 'k'}}; /* identity.rs:119-147 */

/* life.rs:428-507.  Panics of the reference (bad rune, missing field, square index out of the
 * 64-entry LOCALE_STASH) become LO_PANIC. */
int lo_reconstruct(const char *scan, lo_world_t *w) {
    char buf[256];
    size_t len = strlen(scan);
    if (len >= sizeof buf) return LO_PANIC;
    for (size_t i = 0; i <= len; i++) buf[i] = scan[i] == 'X' ? ' ' : scan[i]; /* life.rs:432 */
    /* split(' ') keeps empty fields, so walk fields by hand */
    const char *fields[4] = {0, 0, 0, 0};
    size_t flen[4] = {0, 0, 0, 0};
    {
        const char *p = buf;
        for (int k = 0; k < 4; k++) {
            const char *q = strchr(p, ' ');
            fields[k] = p;
            flen[k] = q ? (size_t)(q - p) : strlen(p);
            if (!q) {
                for (int j = k + 1; j < 4; j++) fields[j] = NULL;
                break;
            }
            p = q + 1;
        }
    }
    int rank = 7, file = 0;
    lo_world_empty(w);
    for (size_t i = 0; i < flen[0]; i++) {
        char rune = fields[0][i];
        if (rune == '/') {
            file = 0;
            rank -= 1;
        } else if (rune >= '0' && rune <= '8') {
            file += rune - '0';
        } else {
            int team = -1, job = -1;
            for (int t = 0; t < 2; t++)
                for (int j = 0; j < 6; j++)
                    if (RUNES[t][j] == rune) {
                        team = t;
                        job = j;
                    }
            if (team < 0) return LO_PANIC; /* identity.rs:203-206 */
            if (rank < 0 || rank * 8 + file >= 64 || rank * 8 + file < 0) return LO_PANIC; /* space.rs:44-47 */
            *plane_of(w, team, job) |= 1ull << (8 * rank + file);
            file += 1;
        }
    }
    if (!fields[1] || flen[1] == 0) return LO_PANIC;
    if (fields[1][0] == 'w') w->initiative = LO_ORANGE;
    else if (fields[1][0] == 'b') w->initiative = LO_BLUE;
    else return LO_PANIC; /* life.rs:471-474 */
    if (!fields[2]) return LO_PANIC;
    for (size_t i = 0; i < flen[2]; i++) {
        char e = fields[2][i];
        if (e == 'K') w->service_eligibility |= LO_ORANGE_EAST;
        else if (e == 'Q') w->service_eligibility |= LO_ORANGE_WEST;
        else if (e == 'k') w->service_eligibility |= LO_BLUE_EAST;
        else if (e == 'q') w->service_eligibility |= LO_BLUE_WEST;
        else if (e == '-') break;
        else return LO_PANIC; /* life.rs:494-496 */
    }
    if (!fields[3]) return LO_PANIC; /* life.rs:500 */
    if (flen[3] == 1 && fields[3][0] == '-') {
        w->passing_by = LO_NONE;
    } else {
        if (flen[3] < 2) return LO_PANIC;
        int r = (uint8_t)fields[3][1] - 49, f = (uint8_t)fields[3][0] - 97; /* space.rs:53-63 */
        if (r < 0 || f < 0 || r * 8 + f >= 64) return LO_PANIC;
        w->passing_by = loc_new(r, f);
    }
    return 0;
}

/* life.rs:293-360 */
int lo_preserve(const lo_world_t *w, char *book, size_t cap) {
    char tmp[128];
    int n = 0;
    int void_run_length = 0;
    for (int rank = 7; rank >= 0; rank--) {
        for (int file = 0; file < 8; file++) {
            uint8_t agent = occupying_agent(w, loc_new(rank, file));
            if (agent != LO_NONE) {
                if (void_run_length > 0) {
                    tmp[n++] = (char)('0' + void_run_length);
                    void_run_length = 0;
                }
                tmp[n++] = RUNES[agent >> 3][agent & 7];
            } else {
                void_run_length += 1;
            }
        }
        if (void_run_length > 0) {
            tmp[n++] = (char)('0' + void_run_length);
            void_run_length = 0;
        }
        if (rank > 0) tmp[n++] = '/';
    }
    tmp[n++] = ' ';
    tmp[n++] = w->initiative == LO_ORANGE ? 'w' : 'b';
    tmp[n++] = ' ';
    int any_service = 0;
    if (w->service_eligibility & LO_ORANGE_EAST) { tmp[n++] = 'K'; any_service = 1; }
    if (w->service_eligibility & LO_ORANGE_WEST) { tmp[n++] = 'Q'; any_service = 1; }
    if (w->service_eligibility & LO_BLUE_EAST) { tmp[n++] = 'k'; any_service = 1; }
    if (w->service_eligibility & LO_BLUE_WEST) { tmp[n++] = 'q'; any_service = 1; }
    if (!any_service) tmp[n++] = '-';
    tmp[n++] = ' ';
    if (w->passing_by != LO_NONE) {
        tmp[n++] = (char)('a' + loc_file(w->passing_by)); /* space.rs:49-51 */
        tmp[n++] = (char)('1' + loc_rank(w->passing_by));
    } else {
        tmp[n++] = '-';
    }
    tmp[n] = 0;
    if ((size_t)n + 1 > cap) return LO_PANIC;
    memcpy(book, tmp, (size_t)n + 1);
    return n;
}

/* ---------------------------------------------------------------- apply (make-move) */

/* life.rs:569-676 */
int lo_apply(const lo_world_t *w, uint8_t team, uint8_t job, uint8_t whence, uint8_t whither, lo_commit_t *out) {
    lo_world_t tree = *w;
    /* 571-575: transit = quench(departure).alight(destination), space.rs:169-173 */
    *plane_of(&tree, team, job) = (cplane(w, team, job) & ~loc_pinpoint(whence)) | loc_pinpoint(whither);
    if (job == LO_FIGUREHEAD) { /* 577-588 */
        tree.service_eligibility &= (uint8_t)~eligibility_bit(team, 0);
        tree.service_eligibility &= (uint8_t)~eligibility_bit(team, 1);
    } else if (job == LO_COP) { /* 589-605: file of whence only, any rank */
        if (loc_file(whence) == 0) tree.service_eligibility &= (uint8_t)~eligibility_bit(team, 1);
        else if (loc_file(whence) == 7) tree.service_eligibility &= (uint8_t)~eligibility_bit(team, 0);
    }
    /* 610-629; concerns_secret_service is life.rs:40-43 */
    if (job == LO_FIGUREHEAD && abs(loc_file(whence) - loc_file(whither)) == 2) {
        int start_file, end_file;
        if (loc_file(whither) == 6) { start_file = 7; end_file = 5; }
        else if (loc_file(whither) == 2) { start_file = 0; end_file = 3; }
        else return LO_PANIC; /* 615-618 */
        uint64_t cops = cplane(&tree, team, LO_COP);
        cops = (cops & ~loc_pinpoint(loc_new(loc_rank(whither), start_file))) |
               loc_pinpoint(loc_new(loc_rank(whither), end_file));
        *plane_of(&tree, team, LO_COP) = cops;
    }
    /* 631-651: was anyone stunned?  Looked up in the PRE-move state (`self`). */
    int opp = opposition(tree.initiative);
    uint8_t hospitalization = occupying_affiliated_agent(w, whither, opp);
    uint8_t ambulance_target = whither;
    if (hospitalization == LO_NONE && job == LO_SERVANT) {
        if (w->passing_by != LO_NONE && w->passing_by == whither) {
            int dr = (opp == LO_ORANGE) ? 1 : -1;
            if (!loc_displace(w->passing_by, dr, 0, &ambulance_target)) return LO_PANIC; /* 646 unwrap */
            hospitalization = occupying_affiliated_agent(w, ambulance_target, opp);
        }
    }
    if (hospitalization != LO_NONE) /* 652-657 */
        tree.plane[6 * (hospitalization >> 3) + (hospitalization & 7)] &= ~loc_pinpoint(ambulance_target);
    tree.initiative = (uint8_t)opp; /* 659 */
    if (job == LO_SERVANT && abs(loc_rank(whither) - loc_rank(whence)) == 2) { /* 660-669 */
        uint8_t behind;
        tree.passing_by = loc_displace(whence, team == LO_ORANGE ? 1 : -1, 0, &behind) ? behind : LO_NONE;
    } else {
        tree.passing_by = LO_NONE;
    }
    memset(out, 0, sizeof *out);
    out->team = team;
    out->job = job;
    out->whence = whence;
    out->whither = whither;
    out->hospitalization = hospitalization;
    out->ascension = LO_NONE;
    out->tree = tree;
    return 0;
}

/* ---------------------------------------------------------------- movegen */

static int underlookahead(const lo_world_t *w, int nihilistically, lo_commit_t *out);
static int lookahead_without_secret_service(const lo_world_t *w, int nihilistically, lo_commit_t *out);

/* life.rs:678-690 */
int lo_in_critical_endangerment(const lo_world_t *w, int team) {
    lo_world_t contingency = *w;
    contingency.initiative = (uint8_t)opposition(team);
    lo_commit_t premonitions[LO_MAX_COMMITS];
    int n = underlookahead(&contingency, 1, premonitions); /* reckless_lookahead */
    if (n < 0) return n;
    for (int i = 0; i < n; i++)
        if (premonitions[i].hospitalization != LO_NONE && (premonitions[i].hospitalization & 7) == LO_FIGUREHEAD)
            return 1;
    return 0;
}

/* life.rs:701-726 */
static int subpredict(lo_commit_t *out, int *n, const lo_commit_t *premonition) {
    int admirality = premonition->team == LO_ORANGE ? 7 : 0; /* life.rs:45-52 */
    if (premonition->job == LO_SERVANT && loc_rank(premonition->whither) == admirality) {
        for (int ascended = 0; ascended < 6; ascended++) {
            if (ascended == LO_SERVANT || ascended == LO_FIGUREHEAD) continue;
            if (*n >= LO_MAX_COMMITS) return LO_PANIC;
            lo_commit_t ascendency = *premonition;
            ascendency.ascension = LO_AGENT(premonition->team, ascended);
            uint64_t vessel = cplane(&premonition->tree, premonition->team, LO_SERVANT) & ~loc_pinpoint(premonition->whither);
            uint64_t ascended_pinfield = cplane(&premonition->tree, premonition->team, ascended) | loc_pinpoint(premonition->whither);
            *plane_of(&ascendency.tree, premonition->team, LO_SERVANT) = vessel;
            *plane_of(&ascendency.tree, premonition->team, ascended) = ascended_pinfield;
            out[(*n)++] = ascendency;
        }
    } else {
        if (*n >= LO_MAX_COMMITS) return LO_PANIC;
        out[(*n)++] = *premonition;
    }
    return 0;
}

/* life.rs:728-739 (predict) with careful_apply (692-699) inlined */
static int predict(const lo_world_t *w, lo_commit_t *out, int *n, int team, int job, uint8_t whence, uint8_t whither,
                   int nihilistically) {
    lo_commit_t premonition;
    int rc = lo_apply(w, (uint8_t)team, (uint8_t)job, whence, whither, &premonition);
    if (rc < 0) return rc;
    if (!nihilistically) {
        int endangered = lo_in_critical_endangerment(&premonition.tree, w->initiative);
        if (endangered < 0) return endangered;
        if (endangered) return 0;
    }
    return subpredict(out, n, &premonition);
}

#define TRY(expr) do { int rc_ = (expr); if (rc_ < 0) return rc_; } while (0)

/* life.rs:749-830 */
int lo_servant_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) {
    int initial_rank = team == LO_ORANGE ? 1 : 6;
    int step = team == LO_ORANGE ? 1 : -1;
    const int stun_files[2] = {-1, 1};
    uint8_t starts[64];
    int n_starts = to_locales(cplane(w, team, LO_SERVANT), starts);
    for (int i = 0; i < n_starts; i++) {
        uint8_t start = starts[i], destination;
        if (loc_displace(start, step, 0, &destination)) { /* 772-783 */
            if (query(unoccupied(w), destination))
                TRY(predict(w, out, n, team, LO_SERVANT, start, destination, nihilistically));
        }
        if (loc_rank(start) == initial_rank) { /* 786-803 */
            uint8_t boost, standard;
            if (!loc_displace(start, 2 * step, 0, &boost) || !loc_displace(start, step, 0, &standard)) return LO_PANIC;
            if (query(unoccupied(w), boost) && query(unoccupied(w), standard))
                TRY(predict(w, out, n, team, LO_SERVANT, start, boost, nihilistically));
        }
        for (int s = 0; s < 2; s++) { /* 805-828 */
            uint8_t stun;
            if (loc_displace(start, step, stun_files[s], &stun)) {
                if (query(occupied_by(w, opposition(team)), stun)) {
                    TRY(predict(w, out, n, team, LO_SERVANT, start, stun, nihilistically));
                } else if (w->passing_by != LO_NONE) {
                    if (w->passing_by == stun)
                        TRY(predict(w, out, n, team, LO_SERVANT, start, stun, nihilistically));
                }
            }
        }
    }
    return 0;
}

/* life.rs:832-855 */
static int ponylike_lookahead(const lo_world_t *w, int team, int job, int nihilistically, lo_commit_t *out, int *n) {
    const uint64_t *table = job == LO_PONY ? LO_PONY_MOVEMENT_TABLE : LO_FIGUREHEAD_MOVEMENT_TABLE;
    uint8_t starts[64], destinations[64];
    int n_starts = to_locales(cplane(w, team, job), starts);
    for (int i = 0; i < n_starts; i++) {
        int n_dest = to_locales(~occupied_by(w, team) & table[loc_pindex(starts[i])], destinations);
        for (int d = 0; d < n_dest; d++)
            TRY(predict(w, out, n, team, job, starts[i], destinations[d], nihilistically));
    }
    return 0;
}

static const int SCHOLAR_OFFSETS[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}}; /* life.rs:10 */
static const int COP_OFFSETS[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};       /* life.rs:11 */

/* life.rs:857-914 */
static int princesslike_lookahead(const lo_world_t *w, int team, int job, const int (*offsets)[2],
                                  const uint8_t *starts, int n_starts, int nihilistically, lo_commit_t *out, int *n) {
    for (int i = 0; i < n_starts; i++) {
        for (int o = 0; o < 4; o++) {
            int venture = 1;
            for (;;) {
                uint8_t destination;
                if (!loc_displace(starts[i], venture * offsets[o][0], venture * offsets[o][1], &destination)) break;
                int empty = query(unoccupied(w), destination);         /* 890 */
                int friend_ = query(occupied_by(w, team), destination); /* 891-892 */
                if (empty || !friend_)
                    TRY(predict(w, out, n, team, job, starts[i], destination, nihilistically));
                if (!empty) break;
                venture += 1;
            }
        }
    }
    return 0;
}

int lo_pony_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) { /* 919-925 */
    return ponylike_lookahead(w, team, LO_PONY, nihilistically, out, n);
}
int lo_scholar_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) { /* 930-938 */
    uint8_t starts[64];
    int ns = to_locales(cplane(w, team, LO_SCHOLAR), starts);
    return princesslike_lookahead(w, team, LO_SCHOLAR, SCHOLAR_OFFSETS, starts, ns, nihilistically, out, n);
}
int lo_cop_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) { /* 944-952 */
    uint8_t starts[64];
    int ns = to_locales(cplane(w, team, LO_COP), starts);
    return princesslike_lookahead(w, team, LO_COP, COP_OFFSETS, starts, ns, nihilistically, out, n);
}
/* life.rs:955-969: every princess's diagonal moves first, then every princess's orthogonal moves */
int lo_princess_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) {
    uint8_t starts[64];
    int ns = to_locales(cplane(w, team, LO_PRINCESS), starts);
    TRY(princesslike_lookahead(w, team, LO_PRINCESS, SCHOLAR_OFFSETS, starts, ns, nihilistically, out, n));
    return princesslike_lookahead(w, team, LO_PRINCESS, COP_OFFSETS, starts, ns, nihilistically, out, n);
}
int lo_figurehead_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) { /* 973-978 */
    return ponylike_lookahead(w, team, LO_FIGUREHEAD, nihilistically, out, n);
}

/* life.rs:284-291 */
int lo_is_being_leered_at_by(const lo_world_t *w, uint8_t locale, int team) {
    lo_world_t tree = *w;
    *plane_of(&tree, opposition(team), LO_FIGUREHEAD) |= loc_pinpoint(locale);
    tree.initiative = (uint8_t)team;
    lo_commit_t prems[LO_MAX_COMMITS];
    int n = lookahead_without_secret_service(&tree, 1, prems);
    if (n < 0) return n;
    for (int i = 0; i < n; i++)
        if (prems[i].whither == locale) return 1;
    return 0;
}

/* life.rs:980-1050 */
int lo_service_lookahead(const lo_world_t *w, int team, int nihilistically, lo_commit_t *out, int *n) {
    int east_service = (w->service_eligibility & eligibility_bit(team, 0)) > 0;
    int west_service = (w->service_eligibility & eligibility_bit(team, 1)) > 0;
    int home_rank = team == LO_ORANGE ? 0 : 7;
    if (!(east_service || west_service)) return 0; /* 999-1007; the debug_assert is compiled out in release */

    struct { uint8_t locales[3]; int n_locales; uint8_t whither; } to_query[2];
    int n_query = 0;
    if (west_service) { /* 1010-1019 */
        to_query[n_query].locales[0] = loc_new(home_rank, 1);
        to_query[n_query].locales[1] = loc_new(home_rank, 2);
        to_query[n_query].locales[2] = loc_new(home_rank, 3);
        to_query[n_query].n_locales = 3;
        to_query[n_query].whither = loc_new(home_rank, 2);
        n_query++;
    }
    if (east_service) { /* 1020-1029 */
        to_query[n_query].locales[0] = loc_new(home_rank, 5);
        to_query[n_query].locales[1] = loc_new(home_rank, 6);
        to_query[n_query].n_locales = 2;
        to_query[n_query].whither = loc_new(home_rank, 6);
        n_query++;
    }
    uint64_t unoc = unoccupied(w);
    int being_leered_at = -1; /* None */
    for (int q = 0; q < n_query; q++) {
        int all_unoccupied = 1;
        for (int i = 0; i < to_query[q].n_locales; i++)
            if (!query(unoc, to_query[q].locales[i])) all_unoccupied = 0;
        if (!all_unoccupied) continue;
        if (being_leered_at < 0) { /* 1034-1040 */
            being_leered_at = lo_is_being_leered_at_by(w, loc_new(home_rank, 4), opposition(team));
            if (being_leered_at < 0) return being_leered_at;
        }
        if (being_leered_at) return 0; /* 1041-1043 */
        int any = 0;
        for (int i = 0; i < to_query[q].n_locales && !any; i++) { /* 1044-1045, short-circuit any() */
            int leered = lo_is_being_leered_at_by(w, to_query[q].locales[i], opposition(team));
            if (leered < 0) return leered;
            any = leered;
        }
        if (!any)
            TRY(predict(w, out, n, team, LO_FIGUREHEAD, loc_new(home_rank, 4), to_query[q].whither, nihilistically));
    }
    return 0;
}

/* life.rs:1052-1070 */
static int lookahead_without_secret_service(const lo_world_t *w, int nihilistically, lo_commit_t *out) {
    int moving_team = w->initiative;
    int n = 0;
    TRY(lo_servant_lookahead(w, moving_team, nihilistically, out, &n));
    TRY(lo_pony_lookahead(w, moving_team, nihilistically, out, &n));
    TRY(lo_scholar_lookahead(w, moving_team, nihilistically, out, &n));
    TRY(lo_cop_lookahead(w, moving_team, nihilistically, out, &n));
    TRY(lo_princess_lookahead(w, moving_team, nihilistically, out, &n));
    TRY(lo_figurehead_lookahead(w, moving_team, nihilistically, out, &n));
    return n;
}

/* life.rs:1072-1076 */
static int underlookahead(const lo_world_t *w, int nihilistically, lo_commit_t *out) {
    int n = lookahead_without_secret_service(w, nihilistically, out);
    if (n < 0) return n;
    TRY(lo_service_lookahead(w, w->initiative, nihilistically, out, &n));
    return n;
}

int lo_lookahead(const lo_world_t *w, lo_commit_t *out) { return underlookahead(w, 0, out); }          /* 1078-1080 */
int lo_reckless_lookahead(const lo_world_t *w, lo_commit_t *out) { return underlookahead(w, 1, out); } /* 1082-1084 */

/* ---------------------------------------------------------------- mind.rs: score */

float lo_orientation(int team) { return team == LO_ORANGE ? 1.0f : -1.0f; } /* mind.rs:29-34 */

static const float FIGURINE_VALUE[6] = {1.0f, 3.2f, 3.3f, 5.1f, 8.8f, 20000.0f}; /* mind.rs:36-48 */

/* mind.rs:50-143.  One rounded f32 multiply and one rounded f32 add per reference statement, in the
 * reference's order (build with -ffp-contract=off so gcc does not fuse them). */
float lo_score(const lo_world_t *w) {
    float valuation = 0.0f;
    valuation += 0.5f * lo_orientation(w->initiative); /* :53, REWARD_FOR_INITIATIVE :26 */
    for (int team = 0; team < 2; team++) { /* :55-68 */
        for (int job = 0; job < 6; job++) {
            float figurine_valuation = lo_orientation(team) * FIGURINE_VALUE[job];
            valuation += (float)(uint8_t)pincount(cplane(w, team, job)) * figurine_valuation;
        }
        if (pincount(cplane(w, team, LO_SCHOLAR)) >= 2) valuation += lo_orientation(team) * 0.5f;
    }
    uint64_t center = lo_center_of_the_world(); /* :70-81 */
    int8_t orange_centerism = (int8_t)pincount((cplane(w, LO_ORANGE, LO_SERVANT) | cplane(w, LO_ORANGE, LO_PONY)) & center);
    int8_t blue_centerism = (int8_t)pincount((cplane(w, LO_BLUE, LO_SERVANT) | cplane(w, LO_BLUE, LO_PONY)) & center);
    valuation += 0.1f * (float)(int8_t)(orange_centerism - blue_centerism);
    uint64_t high_seventh = lo_forward_contour(6); /* :83-89 */
    valuation += 0.5f * (float)pincount(cplane(w, LO_ORANGE, LO_COP) & high_seventh);
    uint64_t low_seventh = lo_forward_contour(1);
    valuation -= 0.5f * (float)pincount(cplane(w, LO_BLUE, LO_COP) & low_seventh);
    for (int f = 0; f < 8; f++) { /* :91-111 */
        uint64_t file = lo_sideways_contour(f);
        int orange_in_line = pincount(cplane(w, LO_ORANGE, LO_SERVANT) & file);
        if (orange_in_line > 1) valuation -= 0.5f * (float)(orange_in_line - 1);
        int blue_in_line = pincount(cplane(w, LO_BLUE, LO_SERVANT) & file);
        if (blue_in_line > 1) valuation += 0.5f * (float)(blue_in_line - 1);
    }
    /* :113-131 */
    valuation += 1.8f * (float)pincount(cplane(w, LO_ORANGE, LO_SERVANT) & high_seventh);
    valuation += 0.6f * (float)pincount(cplane(w, LO_ORANGE, LO_SERVANT) & lo_forward_contour(5));
    valuation -= 1.8f * (float)pincount(cplane(w, LO_BLUE, LO_SERVANT) & low_seventh);
    valuation -= 0.6f * (float)pincount(cplane(w, LO_BLUE, LO_SERVANT) & lo_forward_contour(2));
    /* :133-140 */
    if (w->service_eligibility & (LO_ORANGE_WEST | LO_ORANGE_EAST)) valuation += 0.1f;
    if (w->service_eligibility & (LO_BLUE_WEST | LO_BLUE_EAST)) valuation -= 0.1f;
    return valuation;
}

/* ---------------------------------------------------------------- perft */

uint64_t lo_perft(const lo_world_t *w, int depth) {
    if (depth <= 0) return 1;
    lo_commit_t *prems = (lo_commit_t *)malloc(sizeof(lo_commit_t) * LO_MAX_COMMITS);
    int n = lo_lookahead(w, prems);
    uint64_t total = 0;
    if (n > 0) {
        if (depth == 1) total = (uint64_t)n;
        else
            for (int i = 0; i < n; i++) total += lo_perft(&prems[i].tree, depth - 1);
    }
    free(prems);
    return total;
}

typedef struct {
    const lo_commit_t *roots;
    int n_roots, depth;
    uint64_t *per_root;
    float *scores;
    int next;
    pthread_mutex_t lock;
    int mode; /* 0 perft, 1 negamax */
    int quiet;
} lo_job_t;

float lo_negamax_quiet(const lo_world_t *w, int depth, int quiet);
static void *root_worker(void *arg) {
    lo_job_t *job = (lo_job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&job->lock);
        int i = job->next++;
        pthread_mutex_unlock(&job->lock);
        if (i >= job->n_roots) break;
        if (job->mode == 0) job->per_root[i] = lo_perft(&job->roots[i].tree, job->depth - 1);
        else job->scores[i] = -lo_negamax_quiet(&job->roots[i].tree, job->depth - 1, job->quiet); /* mind.rs:425 */
    }
    return NULL;
}

static void run_root_job(lo_job_t *job, int threads) {
    if (threads < 1) threads = 1;
    if (threads > job->n_roots) threads = job->n_roots;
    pthread_mutex_init(&job->lock, NULL);
    job->next = 0;
    if (threads <= 1) {
        root_worker(job);
    } else {
        pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, 64u << 20);
        for (int t = 0; t < threads; t++) pthread_create(&tids[t], &attr, root_worker, job);
        for (int t = 0; t < threads; t++) pthread_join(tids[t], NULL);
        pthread_attr_destroy(&attr);
        free(tids);
    }
    pthread_mutex_destroy(&job->lock);
}

int lo_perft_divide(const lo_world_t *w, int depth, uint64_t *per_root, int cap, int threads) {
    lo_commit_t *roots = (lo_commit_t *)malloc(sizeof(lo_commit_t) * LO_MAX_COMMITS);
    int n = lo_lookahead(w, roots);
    if (n < 0 || n > cap) {
        free(roots);
        return n < 0 ? n : LO_PANIC;
    }
    lo_job_t job = {roots, n, depth, per_root, NULL, 0, PTHREAD_MUTEX_INITIALIZER, 0, -1};
    if (depth >= 1) run_root_job(&job, threads);
    free(roots);
    return n;
}

/* lo_score over an array of worlds (the full-size parity test scores 2^24 positions) */
void lo_score_batch(const lo_world_t *worlds, size_t n, float *out) {
    for (size_t i = 0; i < n; i++) out[i] = lo_score(&worlds[i]);
}

/* ---------------------------------------------------------------- mind.rs: leaf rule + negamax */

/* mind.rs:145-153 (MVV-LVA); ordering only changes how fast the exact value is found. */
static float mvv_lva(const lo_commit_t *c) {
    if (c->hospitalization == LO_NONE) return 0.0f;
    float patient = lo_orientation(c->hospitalization >> 3) * FIGURINE_VALUE[c->hospitalization & 7];
    float star = lo_orientation(c->team) * FIGURINE_VALUE[c->job];
    return patient - star;
}

/* mind.rs:271-364 without the transposition table and the history table: fail-soft alpha-beta; with a
 * full window the value is the exact negamax value.  quiet < 0 is `quiet: None`, else Some(quiet). */
static float alpha_beta(const lo_world_t *w, int depth, float alpha, float beta, int quiet) {
    lo_commit_t *prems = (lo_commit_t *)malloc(sizeof(lo_commit_t) * LO_MAX_COMMITS);
    int n = lo_reckless_lookahead(w, prems); /* :279 -- also at depth <= 0, like the reference */
    float optimum = -INFINITY;
    if (depth <= 0 || n <= 0) { /* :282-301 */
        float potential_score = lo_orientation(w->initiative) * lo_score(w);
        int stuns = 0;
        if (quiet >= 0 && abs(depth) < quiet) { /* :289-294: keep the commits that hospitalize somebody */
            for (int i = 0; i < n; i++)
                if (prems[i].hospitalization != LO_NONE) prems[stuns++] = prems[i];
        }
        if (stuns == 0) {
            free(prems);
            return potential_score;
        }
        n = stuns;
        optimum = potential_score; /* :298 stand pat */
    }
    /* stable insertion sort by MVV-LVA, descending */
    int order[LO_MAX_COMMITS];
    float key[LO_MAX_COMMITS];
    for (int i = 0; i < n; i++) {
        float k = mvv_lva(&prems[i]);
        int j = i;
        while (j > 0 && key[j - 1] < k) {
            key[j] = key[j - 1];
            order[j] = order[j - 1];
            j--;
        }
        key[j] = k;
        order[j] = i;
    }
    for (int i = 0; i < n; i++) { /* :312-362 */
        float value = -alpha_beta(&prems[order[i]].tree, depth - 1, -beta, -alpha, quiet);
        if (value > optimum) optimum = value;
        if (value > alpha) alpha = value;
        if (alpha >= beta) break;
    }
    free(prems);
    return optimum;
}

float lo_negamax(const lo_world_t *w, int depth) { return alpha_beta(w, depth, -INFINITY, INFINITY, -1); }
float lo_negamax_quiet(const lo_world_t *w, int depth, int quiet) { return alpha_beta(w, depth, -INFINITY, INFINITY, quiet); }

/* mind.rs:375-448: root moves legal unless nihilistically (388-392); each child searched with a
 * full window (406-411); root value = -child value (425). */
int lo_kickoff_quiet(const lo_world_t *w, int depth, int quiet, int nihilistically, lo_commit_t *roots, float *scores, int threads) {
    int n = nihilistically ? lo_reckless_lookahead(w, roots) : lo_lookahead(w, roots);
    if (n <= 0) return n;
    lo_job_t job = {roots, n, depth, NULL, scores, 0, PTHREAD_MUTEX_INITIALIZER, 1, quiet};
    run_root_job(&job, threads);
    return n;
}
int lo_kickoff(const lo_world_t *w, int depth, int nihilistically, lo_commit_t *roots, float *scores, int threads) {
    return lo_kickoff_quiet(w, depth, -1, nihilistically, roots, scores, threads);
}

/* ---------------------------------------------------------------- seeded random playouts */

/* splitmix64 finaliser */
uint64_t lo_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Position `index` of the seeded set (DESIGN.md "playout set"): start from the opening, play
 * L = 16 + (r(0) * 45 >> 32) uniformly random LEGAL moves, move k chosen as
 * lookahead()[r(k) * len >> 32] in canonical order, r(k) = high 32 bits of
 * mix64(mix64(mix64(seed + index) ^ attempt * 0xD1B54A32D192ED03) + k * 0x9E3779B97F4A7C15);
 * if the game ends early, restart with attempt + 1. */
int lo_playout(uint64_t seed, uint64_t index, lo_world_t *out, int *plies_out) {
    lo_commit_t *prems = (lo_commit_t *)malloc(sizeof(lo_commit_t) * LO_MAX_COMMITS);
    uint64_t h0 = lo_mix64(seed + index);
    for (uint64_t attempt = 0; attempt < 1024; attempt++) {
        uint64_t h1 = lo_mix64(h0 ^ (attempt * 0xD1B54A32D192ED03ull));
        lo_world_t w;
        lo_world_default(&w);
        uint64_t r0 = lo_mix64(h1) >> 32;
        int plies = 16 + (int)((r0 * 45) >> 32);
        int k = 0;
        for (; k < plies; k++) {
            int n = lo_lookahead(&w, prems);
            if (n < 0) {
                free(prems);
                return n;
            }
            if (n == 0) break;
            uint64_t r = lo_mix64(h1 + (uint64_t)(k + 1) * 0x9E3779B97F4A7C15ull) >> 32;
            w = prems[(r * (uint64_t)n) >> 32].tree;
        }
        if (k == plies) {
            *out = w;
            if (plies_out) *plies_out = plies;
            free(prems);
            return 0;
        }
    }
    free(prems);
    return LO_PANIC;
}
