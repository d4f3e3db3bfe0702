// ll_device.cuh -- device-side position type, bitboard primitives, the canonical move generator,
// the set-wise legal-move counter and the static evaluator.
//
// Everything here is hand-written for sm_100a.  A position lives in registers as 12 u64 planes
// (the reference's Pinfields, src/space.rs:127-200) plus a u32 of metadata; in HBM it is stored
// structure-of-arrays (see Soa below).
//
// Legality.  The reference decides legality by making the move and running the opponent's full
// reckless_lookahead, looking for a figurehead hospitalization (src/life.rs:678-699).  On
// well-formed positions (include/leafline_b200.h, domain W: disjoint planes, at most one
// figurehead per team) that is exactly "the mover's figurehead square is attacked afterwards", so
// it is decided here once per position instead of once per move:
//   check mask  -- squares a non-figurehead move must land on (capture the single checker or block it);
//   pinned set  -- own pieces that may only move along the line they share with their figurehead;
//   danger map  -- squares the opposition attacks once the figurehead is lifted off the board.
// Passing-by captures and secret service keep the make-the-move-and-test path (two pieces leave a
// rank at once; phantom cops / figureheads).
//
// The header also compiles as plain C++ (LL_HOST_CHECK) so that tests/hostcheck can run the very
// same generator logic against the oracle on the CPU box.  That build is test infrastructure: the
// shipped library contains device code only and has no CPU path.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define LL_HD __device__ __forceinline__
#define LL_HDM __device__ __forceinline__ // member functions
#else
#define LL_HD static inline
#define LL_HDM inline
#endif

typedef unsigned long long u64;
typedef unsigned int u32;

namespace ll {

enum { SERVANT = 0, PONY = 1, SCHOLAR = 2, COP = 3, PRINCESS = 4, FIGUREHEAD = 5 };
constexpr u32 JOB_NONE = 0xF;

constexpr u64 FILE_A = 0x0101010101010101ull;
constexpr u64 FILE_H = 0x8080808080808080ull;
constexpr u64 RANK_1 = 0xFFull;
constexpr u64 MAIN_DIAG = 0x8040201008040201ull;
constexpr u64 MAIN_ANTI = 0x0102040810204080ull;

// ---- intrinsics (device) / builtins (host check) ---------------------------------------------------
#ifdef __CUDACC__
LL_HD int popc64(u64 x) { return __popcll(x); }
LL_HD int lsb64(u64 x) { return __ffsll((long long)x) - 1; } // x != 0
LL_HD int msb64(u64 x) { return 63 - __clzll((long long)x); } // x != 0
LL_HD u64 brev64(u64 x) { return __brevll(x); }
LL_HD int lsb32(u32 x) { return __ffs((int)x) - 1; }
template <class V> LL_HD V ldg(const V *p) { return __ldg(p); }
LL_HD float fmul(float a, float b) { return __fmul_rn(a, b); } // never contracted into an FMA
LL_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
#else
LL_HD int popc64(u64 x) { return __builtin_popcountll(x); }
LL_HD int lsb64(u64 x) { return __builtin_ctzll(x); }
LL_HD int msb64(u64 x) { return 63 - __builtin_clzll(x); }
LL_HD u64 brev64(u64 x) {
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    return __builtin_bswap64(x);
}
LL_HD int lsb32(u32 x) { return __builtin_ctz(x); }
template <class V> LL_HD V ldg(const V *p) { return *p; }
LL_HD float fmul(float a, float b) { return a * b; } // built with -ffp-contract=off
LL_HD float fadd(float a, float b) { return a + b; }
#endif

// meta: bit0 initiative (0 Orange, 1 Blue); bits 1-4 service_eligibility (life.rs:178-181);
// bits 8-15 passing_by Locale byte (rank<<4|file, 0xFF = None) -- same byte as ll_world_t.
struct Pos {
    u64 pl[12];
    u32 meta;
};

LL_HD u32 meta_stm(u32 m) { return m & 1u; }
LL_HD u32 meta_elig(u32 m) { return (m >> 1) & 0xFu; }
LL_HD u32 meta_pb(u32 m) { return (m >> 8) & 0xFFu; }
LL_HD u32 make_meta(u32 stm, u32 elig, u32 pb) { return stm | (elig << 1) | (pb << 8); }
// bit 16: the ABI bytes this position came from were out of range (initiative > 1 or eligibility > 0xF).  Set only by
// the repacking kernels, read only by well_formed(): a malformed ll_world_t is rejected on every path alike instead of
// being silently answered as the position its masked fields spell.
constexpr u32 META_MALFORMED = 1u << 16;
// word 12 of an ll_world_t (initiative | eligibility << 8 | passing_by << 16 | padding) -> meta
LL_HD u32 meta_of_abi_word(u64 m) {
    const u32 stm = (u32)(m & 0xFFu), elig = (u32)((m >> 8) & 0xFFu);
    return make_meta(stm & 1u, elig & 0xFu, (u32)((m >> 16) & 0xFFu)) | ((stm > 1u || elig > 0xFu) ? META_MALFORMED : 0u);
}
LL_HD int loc_to_sq(u32 loc) { return (int)(((loc >> 4) << 3) | (loc & 7u)); }
LL_HD u32 sq_to_loc(int sq) { return (u32)(((sq >> 3) << 4) | (sq & 7)); }
LL_HD u64 passing_by_bit(u32 meta) { // the square a servant may stun "in passing", 0 if none
    const u32 pb = meta_pb(meta);
    return pb != 0xFFu ? 1ull << loc_to_sq(pb) : 0ull;
}

// ---- structure-of-arrays storage in HBM -------------------------------------------------------
// plane j of position i at pl[j*stride + i]; stride % 2 == 0 keeps every plane 16-byte aligned so a
// thread can move two neighbouring positions with one 128-bit access.
struct Soa {
    u64 *pl;
    u32 *meta;
    size_t stride;
};

LL_HD Pos soa_load(const Soa &s, size_t i) {
    Pos p;
#pragma unroll
    for (int j = 0; j < 12; j++) p.pl[j] = ldg(s.pl + (size_t)j * s.stride + i);
    p.meta = ldg(s.meta + i);
    return p;
}
LL_HD void soa_store(const Soa &s, size_t i, const Pos &p) {
#pragma unroll
    for (int j = 0; j < 12; j++) s.pl[(size_t)j * s.stride + i] = p.pl[j];
    s.meta[i] = p.meta;
}

// ---- generated tables (furniture/tablemaker.py:41-62), filled by the host at ll_init ----------
#ifdef __CUDACC__
__device__ u64 d_pony_table[64];
__device__ u64 d_figurehead_table[64];
__device__ u64 d_line_table[64 * 4]; // StagedLines
#else
static u64 d_pony_table[64];
static u64 d_figurehead_table[64];
static u64 d_line_table[64 * 4];
#endif
LL_HD u64 pony_moves(int sq) { return ldg(&d_pony_table[sq]); }
LL_HD u64 figurehead_moves(int sq) { return ldg(&d_figurehead_table[sq]); }

// ---- one piece's sliding attacks: hyperbola quintessence with a full 64-bit bit reversal (works
// for ranks, files and both diagonals; no tables, no memory traffic) -----------------------------
LL_HD u64 line_attacks(u64 occ, u64 bit, u64 mask_ex) {
    u64 f = occ & mask_ex;
    u64 r = brev64(f);
    f -= bit;
    r -= brev64(bit);
    f ^= brev64(r);
    return f & mask_ex;
}
LL_HD u64 rank_mask(int sq) { return RANK_1 << (sq & 56); }
LL_HD u64 file_mask(int sq) { return FILE_A << (sq & 7); }
LL_HD u64 diag_mask(int sq) { // squares with file - rank constant (step 9)
    int d = (sq & 7) - (sq >> 3);
    return d >= 0 ? MAIN_DIAG >> (8 * d) : MAIN_DIAG << (-8 * d);
}
LL_HD u64 anti_mask(int sq) { // squares with file + rank constant (step 7)
    int s = (sq & 7) + (sq >> 3) - 7;
    return s >= 0 ? MAIN_ANTI << (8 * s) : MAIN_ANTI >> (-8 * s);
}
LL_HD u64 diag_attacks(int sq, u64 occ) {
    u64 bit = 1ull << sq;
    return line_attacks(occ, bit, diag_mask(sq) & ~bit) | line_attacks(occ, bit, anti_mask(sq) & ~bit);
}
LL_HD u64 orth_attacks(int sq, u64 occ) {
    u64 bit = 1ull << sq;
    return line_attacks(occ, bit, rank_mask(sq) & ~bit) | line_attacks(occ, bit, file_mask(sq) & ~bit);
}

// ---- all sliders of a kind at once: Kogge-Stone occluded fills (branch-free, no divergence).
// Each returns the squares attacked in that direction, first blocker included. --------------------
// The four "up" fills (left shifts) and, on the bit-reversed board, the four "down" fills: brev maps square s to
// 63 - s (south <-> north, west <-> east, sw <-> ne, se <-> nw, FILE_A <-> FILE_H).  A 64-bit left shift costs one
// IMAD.SHL (FMA pipe) + one SHF (ALU pipe), a right shift two SHF: keeping every fill a left fill takes work off the
// ALU pipe that bounds the perft tail (-3 %).  (Going further -- IMAD.WIDE + IMAD for the whole shift, no ALU
// instruction at all -- was measured slower: DESIGN.md.)
template <int S> LL_HD u64 ks_up(u64 g, u64 e, u64 keep) {
    e &= keep;
    g |= e & (g << S);
    e &= e << S;
    g |= e & (g << (2 * S));
    e &= e << (2 * S);
    g |= e & (g << (4 * S));
    return (g << S) & keep;
}
// East is special: along a rank "up" is plain binary carry, so all sliders of all ranks move at once by one per-byte
// (SWAR) subtraction -- occupied - 2*sliders flips exactly the squares from each slider to its first blocker (or the
// edge: the borrow dies at the byte boundary).  14 instructions instead of the fill's 26.  Sliders must be occupied
// squares (they are: `empty` excludes them).
LL_HD u64 east_attacks(u64 sliders, u64 empty) {
    const u64 H = 0x8080808080808080ull;
    const u64 x = ~empty, y = (sliders << 1) & ~FILE_A;
    const u64 z = ((x | H) - (y & ~H)) ^ ((x ^ ~y) & H); // x - y in every byte
    return x ^ z;
}
// attacks of the orthogonal sliders `orth` and the diagonal sliders `diag` in the four up directions (north, east,
// ne, nw); called with bit-reversed arguments it yields the bit-reversed attacks in the four down directions
LL_HD u64 ks_up_attacks(u64 orth, u64 diag, u64 empty) {
    return ks_up<8>(orth, empty, ~0ull) | east_attacks(orth, empty) | ks_up<9>(diag, empty, ~FILE_A) | ks_up<7>(diag, empty, ~FILE_H);
}
LL_HD u32 ks_up_count(u64 orth, u64 diag, u64 empty, u64 ok) {
    return (u32)(popc64(ks_up<8>(orth, empty, ~0ull) & ok) + popc64(east_attacks(orth, empty) & ok) +
                 popc64(ks_up<9>(diag, empty, ~FILE_A) & ok) + popc64(ks_up<7>(diag, empty, ~FILE_H) & ok));
}

// squares from which a servant of team A attacks `bit` (life.rs:759, 765: stun offsets (+-1 rank, +-1 file))
template <int A> LL_HD u64 servant_attackers_of(u64 bit) {
    if (A == 0) return ((bit >> 7) & ~FILE_A) | ((bit >> 9) & ~FILE_H); // Orange servants stun upwards
    return ((bit << 7) & ~FILE_H) | ((bit << 9) & ~FILE_A);             // Blue servants stun downwards
}
// squares the servants `s` of team A stun towards
template <int A> LL_HD u64 servant_attacks(u64 s) {
    if (A == 0) return ((s << 7) & ~FILE_H) | ((s << 9) & ~FILE_A);
    return ((s >> 9) & ~FILE_H) | ((s >> 7) & ~FILE_A);
}

// squares the ponies `s` jump to, all of them at once (the union of PONY_MOVEMENT_TABLE rows, furniture/tablemaker.py:41-50)
LL_HD u64 pony_attacks(u64 s) {
    const u64 h1 = ((s >> 1) & ~FILE_H) | ((s << 1) & ~FILE_A);
    const u64 h2 = ((s >> 2) & ~(FILE_H | (FILE_H >> 1))) | ((s << 2) & ~(FILE_A | (FILE_A << 1)));
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}

template <int T> LL_HD u64 occupied_by(const Pos &p) { // life.rs:521-540
    return p.pl[6 * T] | p.pl[6 * T + 1] | p.pl[6 * T + 2] | p.pl[6 * T + 3] | p.pl[6 * T + 4] | p.pl[6 * T + 5];
}

// Is `sq` the destination of some pseudo-legal capture-capable move of team A (the attacker)?
// This is what both in_critical_endangerment (life.rs:678-690) and is_being_leered_at_by
// (life.rs:284-291) ask once their full movegen is boiled down, given that `sq` is occupied by the
// defender: servant stuns need only the defender on sq (life.rs:808); every other piece must not
// find one of its own team on sq (life.rs:840-844, 890-893).
template <int A> LL_HD bool attacked_by(const Pos &p, int sq, u64 occ, u64 attacker_occ) {
    u64 bit = 1ull << sq;
    if (servant_attackers_of<A>(bit) & p.pl[6 * A + SERVANT]) return true;
    if (bit & attacker_occ) return false;
    if (pony_moves(sq) & p.pl[6 * A + PONY]) return true;
    if (figurehead_moves(sq) & p.pl[6 * A + FIGUREHEAD]) return true;
    if (diag_attacks(sq, occ) & (p.pl[6 * A + SCHOLAR] | p.pl[6 * A + PRINCESS])) return true;
    if (orth_attacks(sq, occ) & (p.pl[6 * A + COP] | p.pl[6 * A + PRINCESS])) return true;
    return false;
}

// life.rs:678-690 on a well-formed position: is any figurehead of team T attacked by team 1-T?
template <int T> LL_HD bool in_critical_endangerment(const Pos &p) {
    u64 k = p.pl[6 * T + FIGUREHEAD];
    if (!k) return false;
    u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p);
    u64 occ = own | opp;
    while (k) {
        int sq = lsb64(k);
        k &= k - 1;
        if (attacked_by<1 - T>(p, sq, occ, opp)) return true;
    }
    return false;
}

// ---- move descriptors -------------------------------------------------------------------------------
// One generated move, 32 bits: whence sq | whither sq<<6 | job<<12 | ascension job<<15 (7 = none) |
// passing-by flag<<18 | secret-service flag<<19 | parent slot<<20 (filled by the expansion kernels).
constexpr u32 ASC_NONE = 7;
constexpr u32 MV_PASSING_BY = 1u << 18, MV_SERVICE = 1u << 19;
LL_HD u32 mv_pack(u32 job, int from, int to, u32 asc, u32 flags) { return (u32)from | ((u32)to << 6) | (job << 12) | (asc << 15) | flags; }
LL_HD int mv_from(u32 m) { return (int)(m & 63u); }
LL_HD int mv_to(u32 m) { return (int)((m >> 6) & 63u); }
LL_HD u32 mv_job(u32 m) { return (m >> 12) & 7u; }
LL_HD u32 mv_asc(u32 m) { return (m >> 15) & 7u; }
LL_HD u32 mv_slot(u32 m) { return (m >> 20) & 0x7Fu; }
LL_HD u32 mv_owner(u32 m) { return m >> 27; } // canonical expansion of a sharded level: the rank that owns this child

// ---- which rank owns which work unit (frontier sharding by subtree) ---------------------------------
// A unit is a node of the shard ply, named by its PARENT and its index c among the parent's commits in the reference's
// order (lookahead(), life.rs:1052-1084): owner = (hash(parent) + c) mod world.  A pure function of the tree -- every
// rank, every kernel flavour (warp-per-parent, canonical, level-by-level) and every run deals the same units to the same
// ranks without exchanging anything; a parent's commits go round the ranks, the hash only rotates where the round starts.
LL_HD u32 unit_hash(const Pos &p) {
    u64 h = (u64)(p.meta & 0xFFFFu) * 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int j = 0; j < 12; j++) h = (h ^ p.pl[j]) * 0xD6E8FEB86659FD93ull + (h >> 29);
    return (u32)(h >> 32);
}

// ---- make-move (life.rs:569-676 + ascension life.rs:701-726) ----------------------------------------
// T = mover (== initiative).  Returns the hospitalized job (JOB_NONE if nobody was stunned).
template <int T> LL_HD u32 make_child_as(const Pos &p, u32 m, Pos &c) {
    constexpr int E = 1 - T;
    const int from = mv_from(m), to = mv_to(m);
    const u32 job = mv_job(m), asc = mv_asc(m);
    const u64 fb = 1ull << from, tb = 1ull << to;
    const u32 landing = asc != ASC_NONE ? asc : job; // the plane the star ends up on
    // transit: quench whence, alight whither (space.rs:169-173); an ascending servant alights as its new job
#pragma unroll
    for (int j = 0; j < 6; j++) {
        u64 v = p.pl[6 * T + j];
        if ((u32)j == job) v &= ~fb;
        if ((u32)j == landing) v |= tb;
        c.pl[6 * T + j] = v;
    }
    u32 elig = meta_elig(p.meta);
    if (job == FIGUREHEAD) elig &= T == 0 ? ~0x3u : ~0xCu; // life.rs:577-588
    if (m & MV_SERVICE) { // secret service, life.rs:610-629: hop the cop plane whether or not a cop is there
        const int rank8 = to & 56;
        const bool east = (to & 7) == 6;
        const u64 start = 1ull << (rank8 + (east ? 7 : 0)), end = 1ull << (rank8 + (east ? 5 : 3));
        c.pl[6 * T + COP] = (c.pl[6 * T + COP] & ~start) | end;
    }
    if (job == COP) { // life.rs:589-605: file of whence only, any rank
        const int f = from & 7;
#ifdef LL_ORTHODOX_COP_CAPTURE
        const bool counts = (from >> 3) == (T == 0 ? 0 : 7); // probe build: only a cop leaving its home square (quirk 3)
#else
        const bool counts = true;
#endif
        if (counts && f == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (counts && f == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
    }
    // life.rs:631-657: victim looked up on the pre-move board; passing by removes the servant one
    // rank "behind" the passed square
    const u64 victim = (m & MV_PASSING_BY) ? (T == 0 ? tb >> 8 : tb << 8) : tb;
    u32 hosp = JOB_NONE;
#pragma unroll
    for (int j = 5; j >= 0; j--)
        if (p.pl[6 * E + j] & victim) hosp = (u32)j; // first in dramatis_personae order wins (life.rs:550-557)
#pragma unroll
    for (int j = 0; j < 6; j++) c.pl[6 * E + j] = p.pl[6 * E + j] & ~victim; // W: planes are disjoint
#ifdef LL_ORTHODOX_COP_CAPTURE
    // PROBE BUILD ONLY (profiles/tools/orthodox_probe.py): orthodox chess also withdraws the eligibility when something
    // lands on a cop's home square; the reference does not.  Together with the two other
    // switches of this macro (quirks 1 and 3) the counts must equal the published orthodox perft numbers.
    if (tb & (1ull << 0)) elig &= ~0x1u;
    if (tb & (1ull << 7)) elig &= ~0x2u;
    if (tb & (1ull << 56)) elig &= ~0x4u;
    if (tb & (1ull << 63)) elig &= ~0x8u;
#endif
    u32 new_pb = 0xFFu; // life.rs:660-669
    if (job == SERVANT) {
        const int dr = (to >> 3) - (from >> 3);
        if (dr == 2 || dr == -2) new_pb = sq_to_loc(T == 0 ? from + 8 : from - 8);
    }
    c.meta = make_meta(E, elig, new_pb);
    return hosp;
}
LL_HD u32 make_child(const Pos &p, u32 m, Pos &c) {
    return meta_stm(p.meta) == 0 ? make_child_as<0>(p, m, c) : make_child_as<1>(p, m, c);
}

// packed commit metadata handed to k_commits_to_aos: bit0 team | job<<1 | whence sq<<4 | whither sq<<10 |
// hosp job<<16 | ascension job<<20 (0xF = none)
LL_HD u32 pack_commit(u32 team, u32 m, u32 hosp) {
    const u32 asc = mv_asc(m) == ASC_NONE ? JOB_NONE : mv_asc(m);
    return team | (mv_job(m) << 1) | ((u32)mv_from(m) << 4) | ((u32)mv_to(m) << 10) | (hosp << 16) | (asc << 20);
}

// ---- legality context of one position (see the header comment) ---------------------------------------
struct Legal {
    u64 check;      // ~0: not in check; one checker: its square + the squares between; two: 0
    u64 pinned;     // own pieces pinned against the figurehead
    u64 danger;     // squares the figurehead may not step onto
    u64 kd, ka, kf, kr; // the four lines through the figurehead's square
};
LL_HD u64 pin_line(const Legal &L, u64 bit) { return (L.kd & bit) ? L.kd : (L.ka & bit) ? L.ka : (L.kf & bit) ? L.kf : L.kr; }
// may the piece on `fb` land on `tb`?  (not for the figurehead itself)
LL_HD bool lands_legally(const Legal &L, u64 fb, u64 tb) {
    if (!(tb & L.check)) return false;
    return !(fb & L.pinned) || (tb & pin_line(L, fb));
}
LL_HD u64 between_on(u64 line, u64 a, u64 b) { return ((a - 1) ^ (b - 1)) & line & ~(a | b); } // a, b on `line`

// The four lines through a square, the square itself excluded -- computed (shifts of the two main diagonals: ALU pipe) or
// read from a 2 KiB table staged in shared memory (two 128-bit loads: LSU pipe).  BASELINE.json's north_star names the
// choice ("Kogge-Stone fills or shared-memory-staged ... tables"); both are built and A/B-timed with ncu counters on the
// perft tail, whose bound is the ALU pipe.  The set-wise fills have no table form: a table answers one
// slider at a time, the fills answer all sliders of a direction at once at a cost independent of their number.
struct ComputedLines {
    LL_HDM void operator()(int sq, u64 bit, u64 &kd, u64 &ka, u64 &kf, u64 &kr) const {
        kd = diag_mask(sq) & ~bit;
        ka = anti_mask(sq) & ~bit;
        kf = file_mask(sq) & ~bit;
        kr = rank_mask(sq) & ~bit;
    }
};
struct StagedLines {
    const u64 *table; // [64][4]: diagonal, anti-diagonal, file, rank through the square, the square excluded; 32-byte rows
    LL_HDM void operator()(int sq, u64, u64 &kd, u64 &ka, u64 &kf, u64 &kr) const {
#ifdef __CUDACC__
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(table + 4 * sq), b = *reinterpret_cast<const ulonglong2 *>(table + 4 * sq + 2);
        kd = a.x;
        ka = a.y;
        kf = b.x;
        kr = b.y;
#else
        kd = table[4 * sq]; ka = table[4 * sq + 1]; kf = table[4 * sq + 2]; kr = table[4 * sq + 3];
#endif
    }
};

template <int T, class Lines = ComputedLines> LL_HD Legal legal_context(const Pos &p, u64 own, u64 opp, const Lines &lines = Lines()) {
    constexpr int E = 1 - T;
    Legal L;
    L.check = ~0ull;
    L.pinned = 0;
    L.danger = 0;
    L.kd = L.ka = L.kf = L.kr = 0;
    const u64 k = p.pl[6 * T + FIGUREHEAD];
    if (!k) return L; // nothing to endanger: every pseudo-legal move is legal (life.rs:678-690)
    const int ksq = lsb64(k);
    const u64 occ = own | opp;
    const u64 e_diag = p.pl[6 * E + SCHOLAR] | p.pl[6 * E + PRINCESS], e_orth = p.pl[6 * E + COP] | p.pl[6 * E + PRINCESS];
    // danger map: the opposition's attacks with our figurehead lifted (it cannot hide behind itself).  Only the
    // figurehead's own steps consult it, so a boxed-in figurehead (siblings in a perft frontier share that) skips it.
    const u64 escapes = figurehead_moves(ksq) & ~own;
#ifdef LL_SINGLE_ESCAPE_TEST
    // EXPERIMENT, off by default (tests/test_hostcheck.py keeps it equal to the oracle): 44 % of the perft(6)
    // tail's children have exactly ONE escape square -- ask whether that square is attacked (four lines from it, the
    // figurehead lifted) instead of filling the whole map.  Not timed on the GPU yet.
    if (escapes && !(escapes & (escapes - 1))) {
        const int sq = lsb64(escapes);
        const u64 occ_wo_k = occ ^ k;
        const u64 attackers = (servant_attackers_of<E>(escapes) & p.pl[6 * E + SERVANT]) | (pony_moves(sq) & p.pl[6 * E + PONY]) |
                              (figurehead_moves(sq) & p.pl[6 * E + FIGUREHEAD]) | (diag_attacks(sq, occ_wo_k) & e_diag) |
                              (orth_attacks(sq, occ_wo_k) & e_orth);
        L.danger = attackers ? escapes : 0ull;
    } else
#endif
    if (escapes) {
        const u64 empty_wo_k = ~(occ ^ k);
        u64 danger = servant_attacks<E>(p.pl[6 * E + SERVANT]) | pony_attacks(p.pl[6 * E + PONY]);
        if (p.pl[6 * E + FIGUREHEAD]) danger |= figurehead_moves(lsb64(p.pl[6 * E + FIGUREHEAD]));
        danger |= ks_up_attacks(e_orth, e_diag, empty_wo_k) | brev64(ks_up_attacks(brev64(e_orth), brev64(e_diag), brev64(empty_wo_k)));
        L.danger = danger;
    }
    // checkers and pins along the four lines through the figurehead
    lines(ksq, k, L.kd, L.ka, L.kf, L.kr);
    u64 checkers = (servant_attackers_of<E>(k) & p.pl[6 * E + SERVANT]) | (pony_moves(ksq) & p.pl[6 * E + PONY]) |
                   (figurehead_moves(ksq) & p.pl[6 * E + FIGUREHEAD]);
    u64 pinned = 0;
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const u64 line = l == 0 ? L.kd : l == 1 ? L.ka : l == 2 ? L.kf : L.kr;
        const u64 snipers = (l < 2 ? e_diag : e_orth) & line;
        if (!snipers) continue;
        const u64 seen = line_attacks(occ, k, line);
        checkers |= seen & snipers;
        const u64 blockers = seen & own;
        if (blockers) {
            u64 pinners = line_attacks(occ ^ blockers, k, line) & ~seen & snipers;
            for (; pinners; pinners &= pinners - 1) pinned |= between_on(line, k, pinners & (0 - pinners)) & blockers;
        }
    }
    L.pinned = pinned;
    if (checkers) {
        if (checkers & (checkers - 1)) {
            L.check = 0; // two checkers: only the figurehead can move
        } else {
            L.check = checkers | between_on(pin_line(L, checkers), k, checkers); // ponies / servants: nothing between
            if (!((L.kd | L.ka | L.kf | L.kr) & checkers)) L.check = checkers;   // a pony is on none of the lines
        }
    }
    return L;
}

// ---- secret service (life.rs:980-1050) --------------------------------------------------------------
// `leered` is is_being_leered_at_by boiled down (see attacked_by).  Sink: move(job, from, to, asc, flags).
template <int T, bool NIHIL, class Sink> LL_HD void service_lookahead(const Pos &p, u64 own, u64 opp, Sink &sink) {
    const u32 elig = meta_elig(p.meta);
    const bool west = elig & (T == 0 ? 0x1u : 0x4u), east = elig & (T == 0 ? 0x2u : 0x8u);
    if (!(west || east)) return;
    const int home = T == 0 ? 0 : 56;
    const u64 occ = own | opp;
    const bool west_clear = west && !(occ & (0x0Eull << home)), east_clear = east && !(occ & (0x60ull << home));
    if (!(west_clear || east_clear)) return;
    if (attacked_by<1 - T>(p, home + 4, occ, opp)) return; // life.rs:1034-1043
#pragma unroll 1
    for (int side = 0; side < 2; side++) { // West is asked first (life.rs:1009-1033)
        if (side == 0 ? !west_clear : !east_clear) continue;
        bool leered;
        if (side == 0)
            leered =
#ifndef LL_ORTHODOX_COP_CAPTURE
                attacked_by<1 - T>(p, home + 1, occ, opp) || // the b-file square too: SURVEY.md 8(a) quirk 1 (not in the probe build)
#endif
                attacked_by<1 - T>(p, home + 2, occ, opp) || attacked_by<1 - T>(p, home + 3, occ, opp);
        else
            leered = attacked_by<1 - T>(p, home + 5, occ, opp) || attacked_by<1 - T>(p, home + 6, occ, opp);
        if (leered) continue;
        const int to = home + (side == 0 ? 2 : 6);
        if (!NIHIL) { // careful_apply, life.rs:692-699
            Pos c;
            make_child_as<T>(p, mv_pack(FIGUREHEAD, home + 4, to, ASC_NONE, MV_SERVICE), c);
            if (in_critical_endangerment<T>(c)) continue;
        }
        sink.move(FIGUREHEAD, home + 4, to, ASC_NONE, MV_SERVICE);
    }
}

// passing by (life.rs:800-822): the make-the-move-and-test path
template <int T> LL_HD bool passing_by_is_legal(const Pos &p, int from, int to) {
    Pos c;
    make_child_as<T>(p, mv_pack(SERVANT, from, to, ASC_NONE, MV_PASSING_BY), c);
    return !in_critical_endangerment<T>(c);
}

// ---- the canonical generator (life.rs:1052-1084) -----------------------------------------------------
// servants, ponies, scholars, cops, princesses (all diagonals, then all orthogonals), figurehead, then
// secret service West, East; pieces by ascending square; NIHIL = reckless (pseudo-legal).
template <int T, bool NIHIL, class Sink> LL_HD void servant_move(const Legal &L, Sink &sink, int from, int to, u32 flags) {
    if (!NIHIL && !lands_legally(L, 1ull << from, 1ull << to)) return; // predict / careful_apply, life.rs:728-739
    if ((to >> 3) == (T == 0 ? 7 : 0)) { // subpredict, life.rs:701-726: Pony, Scholar, Cop, Princess
#pragma unroll
        for (u32 job = PONY; job <= PRINCESS; job++) sink.move(SERVANT, from, to, job, flags);
    } else {
        sink.move(SERVANT, from, to, ASC_NONE, flags);
    }
}
template <bool DESC, class Sink> LL_HD void each_target(Sink &sink, u32 job, int from, u64 targets) {
    while (targets) {
        const int to = DESC ? msb64(targets) : lsb64(targets);
        targets ^= 1ull << to;
        sink.move(job, from, to, ASC_NONE, 0);
    }
}
// one slider's rays in the reference's direction order, nearest square first (life.rs:857-914)
template <bool NIHIL, bool DIAGONAL, class Sink> LL_HD void slider_rays(const Legal &L, Sink &sink, u32 job, int from, u64 occ, u64 own) {
    const u64 bit = 1ull << from;
    u64 allowed = ~own;
    if (!NIHIL) {
        allowed &= L.check;
        if (bit & L.pinned) allowed &= pin_line(L, bit);
        if (!allowed) return;
    }
    const u64 below = bit - 1, above = ~(bit | below);
    if (DIAGONAL) { // SCHOLAR_OFFSETS (-1,-1), (-1,1), (1,-1), (1,1)  (life.rs:10)
        const u64 d = line_attacks(occ, bit, diag_mask(from) & ~bit) & allowed;
        const u64 a = line_attacks(occ, bit, anti_mask(from) & ~bit) & allowed;
        each_target<true>(sink, job, from, d & below);
        each_target<true>(sink, job, from, a & below);
        each_target<false>(sink, job, from, a & above);
        each_target<false>(sink, job, from, d & above);
    } else { // COP_OFFSETS (-1,0), (1,0), (0,-1), (0,1)  (life.rs:11)
        const u64 f = line_attacks(occ, bit, file_mask(from) & ~bit) & allowed;
        const u64 r = line_attacks(occ, bit, rank_mask(from) & ~bit) & allowed;
        each_target<true>(sink, job, from, f & below);
        each_target<false>(sink, job, from, f & above);
        each_target<true>(sink, job, from, r & below);
        each_target<false>(sink, job, from, r & above);
    }
}

template <int T, bool NIHIL, class Sink> LL_HD void lookahead_as(const Pos &p, Sink &sink) {
    const u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p), occ = own | opp, empty = ~occ;
    Legal L;
    if (!NIHIL) L = legal_context<T>(p, own, opp);
    const bool others_may_move = NIHIL || L.check != 0;
    if (others_may_move) {
        { // life.rs:749-830
            const u64 pb_bit = passing_by_bit(p.meta);
            for (u64 s = p.pl[6 * T + SERVANT]; s; s &= s - 1) {
                const int from = lsb64(s);
                const u64 fb = 1ull << from;
                const u64 one = T == 0 ? fb << 8 : fb >> 8;
                if (one & empty) servant_move<T, NIHIL>(L, sink, from, T == 0 ? from + 8 : from - 8, 0);
                if ((from >> 3) == (T == 0 ? 1 : 6)) {
                    const u64 two = T == 0 ? fb << 16 : fb >> 16;
                    if ((two & empty) && (one & empty)) servant_move<T, NIHIL>(L, sink, from, T == 0 ? from + 16 : from - 16, 0);
                }
                const u64 west = T == 0 ? (fb << 7) & ~FILE_H : (fb >> 9) & ~FILE_H;
                const u64 east = T == 0 ? (fb << 9) & ~FILE_A : (fb >> 7) & ~FILE_A;
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const u64 dest = side == 0 ? west : east;
                    const int to = side == 0 ? (T == 0 ? from + 7 : from - 9) : (T == 0 ? from + 9 : from - 7);
                    if (!dest) continue;
                    if (dest & opp) servant_move<T, NIHIL>(L, sink, from, to, 0);
                    else if (dest == pb_bit && (NIHIL || passing_by_is_legal<T>(p, from, to))) sink.move(SERVANT, from, to, ASC_NONE, MV_PASSING_BY);
                }
            }
        }
        for (u64 s = p.pl[6 * T + PONY]; s; s &= s - 1) { // life.rs:832-855
            const int from = lsb64(s);
            u64 targets = pony_moves(from) & ~own;
            if (!NIHIL) targets = ((1ull << from) & L.pinned) ? 0 : targets & L.check; // a pinned pony never stays on its line
            each_target<false>(sink, PONY, from, targets);
        }
        for (u64 s = p.pl[6 * T + SCHOLAR]; s; s &= s - 1) slider_rays<NIHIL, true>(L, sink, SCHOLAR, lsb64(s), occ, own);
        for (u64 s = p.pl[6 * T + COP]; s; s &= s - 1) slider_rays<NIHIL, false>(L, sink, COP, lsb64(s), occ, own);
        for (u64 s = p.pl[6 * T + PRINCESS]; s; s &= s - 1) slider_rays<NIHIL, true>(L, sink, PRINCESS, lsb64(s), occ, own); // life.rs:955-969
        for (u64 s = p.pl[6 * T + PRINCESS]; s; s &= s - 1) slider_rays<NIHIL, false>(L, sink, PRINCESS, lsb64(s), occ, own);
    }
    for (u64 s = p.pl[6 * T + FIGUREHEAD]; s; s &= s - 1) { // life.rs:973-978
        const int from = lsb64(s);
        u64 targets = figurehead_moves(from) & ~own;
        if (!NIHIL) targets &= ~L.danger;
        each_target<false>(sink, FIGUREHEAD, from, targets);
    }
    if (NIHIL || L.check == ~0ull) service_lookahead<T, NIHIL>(p, own, opp, sink); // never out of check (life.rs:1034-1043)
}

template <bool NIHIL, class Sink> LL_HD void lookahead(const Pos &p, Sink &sink) {
    if (meta_stm(p.meta) == 0) lookahead_as<0, NIHIL>(p, sink);
    else lookahead_as<1, NIHIL>(p, sink);
}

// ---- set-wise move count: the number of commits lookahead() / reckless_lookahead() would produce,
// by popcounts over whole piece sets (no per-move work) ---------------------------------------------------
struct CountSink {
    u32 n = 0;
    LL_HDM void move(u32, int, int, u32, u32) { n++; }
};

template <int T, bool NIHIL, class Lines = ComputedLines> LL_HD u32 count_moves_as(const Pos &p, const Lines &lines = Lines()) {
    const u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p), occ = own | opp, empty = ~occ;
    Legal L;
    if (!NIHIL) L = legal_context<T>(p, own, opp, lines);
    else { L.check = ~0ull; L.pinned = 0; L.danger = 0; L.kd = L.ka = L.kf = L.kr = 0; }
    u32 n = 0;
    if (p.pl[6 * T + FIGUREHEAD]) n += popc64(figurehead_moves(lsb64(p.pl[6 * T + FIGUREHEAD])) & ~own & ~L.danger);
    if (L.check == 0) return n;
    const u64 cm = L.check, free = ~L.pinned;
    { // servants: a pinned servant may still push along a file pin / stun along a diagonal pin
        const u64 s = p.pl[6 * T + SERVANT];
        const u64 pushers = s & (free | L.kf);
        const u64 west_stunners = s & (free | (T == 0 ? L.ka : L.kd)), east_stunners = s & (free | (T == 0 ? L.kd : L.ka));
        const u64 last = T == 0 ? RANK_1 << 56 : RANK_1;
        const u64 one = (T == 0 ? pushers << 8 : pushers >> 8) & empty;
        const u64 two = (T == 0 ? (one & (RANK_1 << 16)) << 8 : (one & (RANK_1 << 40)) >> 8) & empty & cm;
        const u64 west = (T == 0 ? (west_stunners << 7) : (west_stunners >> 9)) & ~FILE_H & opp & cm;
        const u64 east = (T == 0 ? (east_stunners << 9) : (east_stunners >> 7)) & ~FILE_A & opp & cm;
        const u64 one_ok = one & cm;
        n += popc64(one_ok) + popc64(two) + popc64(west) + popc64(east);
        n += 3 * (popc64(one_ok & last) + popc64(west & last) + popc64(east & last)); // ascension x4
        const u64 pb_bit = passing_by_bit(p.meta);
        if (pb_bit)
            for (u64 c = servant_attackers_of<T>(pb_bit) & s; c; c &= c - 1)
                if (NIHIL || passing_by_is_legal<T>(p, lsb64(c), lsb64(pb_bit))) n++;
    }
    for (u64 s = p.pl[6 * T + PONY] & free; s; s &= s - 1) n += popc64(pony_moves(lsb64(s)) & ~own & cm);
    {
        const u64 diag = p.pl[6 * T + SCHOLAR] | p.pl[6 * T + PRINCESS], orth = p.pl[6 * T + COP] | p.pl[6 * T + PRINCESS];
        const u64 ok = ~own & cm;
        const u64 df = diag & free, of = orth & free;
        // rays of different pieces in one direction never overlap (a ray stops at the next piece)
        n += ks_up_count(of, df, empty, ok) + ks_up_count(brev64(of), brev64(df), brev64(empty), brev64(ok));
        if (!NIHIL) { // pinned sliders slide along the pin only
            const u64 k = p.pl[6 * T + FIGUREHEAD];
            for (u64 s = diag & L.pinned & (L.kd | L.ka); s; s &= s - 1) {
                const u64 bit = s & (0 - s);
                n += popc64(line_attacks(occ, bit, (pin_line(L, bit) | k) & ~bit) & ok);
            }
            for (u64 s = orth & L.pinned & (L.kf | L.kr); s; s &= s - 1) {
                const u64 bit = s & (0 - s);
                n += popc64(line_attacks(occ, bit, (pin_line(L, bit) | k) & ~bit) & ok);
            }
        }
    }
    if (cm == ~0ull) {
        CountSink sink;
        service_lookahead<T, NIHIL>(p, own, opp, sink);
        n += sink.n;
    }
    return n;
}
template <bool NIHIL> LL_HD u32 count_moves(const Pos &p) {
    return meta_stm(p.meta) == 0 ? count_moves_as<0, NIHIL>(p) : count_moves_as<1, NIHIL>(p);
}
template <bool NIHIL, class Lines> LL_HD u32 count_moves(const Pos &p, const Lines &lines) {
    return meta_stm(p.meta) == 0 ? count_moves_as<0, NIHIL>(p, lines) : count_moves_as<1, NIHIL>(p, lines);
}

// ---- reckless commits as (piece or servant-set, target mask) records ------------------------------------
// For consumers that need every child but not the canonical order (the horizon kernel): one record per piece --
// and one per KIND of servant step for all servants at once -- instead of one descriptor per move.  The k-th
// commit of a position is recovered from the records by select-the-k-th-set-bit, so generation costs per piece and
// decoding runs one child per thread at full lane efficiency.  The commits are exactly those of
// reckless_lookahead() (life.rs:1082), in another order; with STUNS only those that hospitalize somebody
// (mind.rs:292-294).  At most 36 records exist: servant sets 18 (3 kinds x (plain + 4 ascensions), double step,
// 2 passing-by), up to 15 other pieces, 2 secret services, and there are at most 16 pieces.
constexpr int REC_MAX = 36;
enum { REC_PIECE = 0, REC_PUSH = 1, REC_DOUBLE = 2, REC_STUN_WEST = 3, REC_STUN_EAST = 4, REC_SERVICE = 5 };
// info: kind | job<<3 | from<<6 | asc<<12 | passing-by<<15 | first move index of the record<<16
LL_HD u32 rec_info(u32 kind, u32 job, int from, u32 asc, u32 passing_by, u32 base) {
    return kind | (job << 3) | ((u32)from << 6) | (asc << 12) | (passing_by << 15) | (base << 16);
}
LL_HD u32 rec_base(u32 info) { return info >> 16; }
LL_HD int select_bit(u64 mask, u32 t) { // index of the t-th (0-based) set bit of mask; t < popc(mask)
    u32 w = (u32)mask;
    int at = 0;
    u32 c = (u32)popc64((u64)w);
    if (t >= c) { t -= c; w = (u32)(mask >> 32); at = 32; }
    c = (u32)popc64((u64)(w & 0xFFFFu));
    if (t >= c) { t -= c; w >>= 16; at += 16; }
    c = (u32)popc64((u64)(w & 0xFFu));
    if (t >= c) { t -= c; w >>= 8; at += 8; }
    c = (u32)popc64((u64)(w & 0xFu));
    if (t >= c) { t -= c; w >>= 4; at += 4; }
    c = (u32)popc64((u64)(w & 0x3u));
    if (t >= c) { t -= c; w >>= 2; at += 2; }
    if (t >= (w & 1u)) at += 1;
    return at;
}
// the move descriptor of commit number t (0-based) of a record, for mover `team`
LL_HD u32 rec_move_to(u32 info, int to, u32 team);
LL_HD u32 rec_move(u64 mask, u32 info, u32 t, u32 team) { return rec_move_to(info, select_bit(mask, t), team); }
struct ServiceBits { // collects the secret-service destinations the exact path admits
    u64 to_bits = 0;
    LL_HDM void move(u32, int, int to, u32, u32) { to_bits |= 1ull << to; }
};
// Out: void put(u64 mask, u32 info_without_base) -- called only with mask != 0; returns nothing.  The caller keeps
// the running move count (the record's base) itself.
// PART: 0 = every record; 1 = the servants', the figurehead's and the secret service's; 2 = the ponies' and the sliders'
// (the horizon kernel generates a parent's records with two threads, one part each)
template <int T, bool STUNS, class Out, int PART = 0> LL_HD void reckless_records_as(const Pos &p, Out &out) {
    const u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p), occ = own | opp, empty = ~occ;
    const u64 allowed = STUNS ? opp : ~own;
    const u64 last = T == 0 ? RANK_1 << 56 : RANK_1;
    if (PART != 2) { // life.rs:749-830, all servants at once
        const u64 s = p.pl[6 * T + SERVANT];
        const u64 one = (T == 0 ? s << 8 : s >> 8) & empty;
        const u64 two = (T == 0 ? (one & (RANK_1 << 16)) << 8 : (one & (RANK_1 << 40)) >> 8) & empty;
        const u64 west = (T == 0 ? s << 7 : s >> 9) & ~FILE_H, east = (T == 0 ? s << 9 : s >> 7) & ~FILE_A;
        const u64 pb_bit = passing_by_bit(p.meta);
#pragma unroll
        for (int kind = REC_PUSH; kind <= REC_STUN_EAST; kind++) {
            if (kind == REC_DOUBLE) {
                if (!STUNS && two) out.put(two, rec_info(REC_DOUBLE, SERVANT, 0, ASC_NONE, 0, 0));
                continue;
            }
            if (kind == REC_PUSH && STUNS) continue;
            const u64 targets = kind == REC_PUSH ? one : (kind == REC_STUN_WEST ? west : east) & opp;
            if (targets & ~last) out.put(targets & ~last, rec_info((u32)kind, SERVANT, 0, ASC_NONE, 0, 0));
            if (targets & last) {
#pragma unroll
                for (u32 asc = PONY; asc <= PRINCESS; asc++) out.put(targets & last, rec_info((u32)kind, SERVANT, 0, asc, 0, 0));
            }
            if (kind != REC_PUSH) { // passing by: the passed square is empty, so it is never in `targets`
                const u64 passing = (kind == REC_STUN_WEST ? west : east) & pb_bit;
                if (passing) out.put(passing, rec_info((u32)kind, SERVANT, 0, ASC_NONE, 1, 0));
            }
        }
    }
    if (PART != 1) {
        for (u64 s = p.pl[6 * T + PONY]; s; s &= s - 1) { // life.rs:832-855
            const int from = lsb64(s);
            const u64 targets = pony_moves(from) & allowed;
            if (targets) out.put(targets, rec_info(REC_PIECE, PONY, from, ASC_NONE, 0, 0));
        }
        for (u64 s = p.pl[6 * T + SCHOLAR]; s; s &= s - 1) { // life.rs:857-914
            const int from = lsb64(s);
            const u64 targets = diag_attacks(from, occ) & allowed;
            if (targets) out.put(targets, rec_info(REC_PIECE, SCHOLAR, from, ASC_NONE, 0, 0));
        }
        for (u64 s = p.pl[6 * T + COP]; s; s &= s - 1) {
            const int from = lsb64(s);
            const u64 targets = orth_attacks(from, occ) & allowed;
            if (targets) out.put(targets, rec_info(REC_PIECE, COP, from, ASC_NONE, 0, 0));
        }
        for (u64 s = p.pl[6 * T + PRINCESS]; s; s &= s - 1) { // life.rs:955-969, both ray families in one record
            const int from = lsb64(s);
            const u64 targets = (diag_attacks(from, occ) | orth_attacks(from, occ)) & allowed;
            if (targets) out.put(targets, rec_info(REC_PIECE, PRINCESS, from, ASC_NONE, 0, 0));
        }
    }
    if (PART != 2) {
        for (u64 s = p.pl[6 * T + FIGUREHEAD]; s; s &= s - 1) { // life.rs:973-978
            const int from = lsb64(s);
            const u64 targets = figurehead_moves(from) & allowed;
            if (targets) out.put(targets, rec_info(REC_PIECE, FIGUREHEAD, from, ASC_NONE, 0, 0));
        }
        if (!STUNS) { // life.rs:980-1050: the exact path (leer tests), never a stun
            ServiceBits service;
            service_lookahead<T, true>(p, own, opp, service);
            if (service.to_bits) out.put(service.to_bits, rec_info(REC_SERVICE, FIGUREHEAD, (T == 0 ? 0 : 56) + 4, ASC_NONE, 0, 0));
        }
    }
}
template <bool STUNS, class Out> LL_HD void reckless_records(const Pos &p, Out &out) {
    if (meta_stm(p.meta) == 0) reckless_records_as<0, STUNS>(p, out);
    else reckless_records_as<1, STUNS>(p, out);
}
template <bool STUNS, int PART, class Out> LL_HD void reckless_records_part(const Pos &p, Out &out) {
    if (meta_stm(p.meta) == 0) reckless_records_as<0, STUNS, Out, PART>(p, out);
    else reckless_records_as<1, STUNS, Out, PART>(p, out);
}

// ---- legal commits as records (the order-free consumers of lookahead(): perft chain and tail) -----------------
// The same commits as lookahead() (life.rs:1078), in another order, produced per piece set / per piece with the
// legality masks applied to whole target sets -- a fraction of the canonical generator's code and instructions
// (the narrow perft levels are bound by instruction fetch: one thread runs the whole generator once).
template <int T, class Out> LL_HD void legal_records_as(const Pos &p, Out &out) {
    const u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p), occ = own | opp, empty = ~occ;
    const Legal L = legal_context<T>(p, own, opp);
    const u64 k = p.pl[6 * T + FIGUREHEAD];
    if (k) { // life.rs:973-978
        const int from = lsb64(k);
        const u64 targets = figurehead_moves(from) & ~own & ~L.danger;
        if (targets) out.put(targets, rec_info(REC_PIECE, FIGUREHEAD, from, ASC_NONE, 0, 0));
    }
    if (L.check == 0) return; // two checkers: only the figurehead moves
    const u64 cm = L.check, free = ~L.pinned;
    { // life.rs:749-830: a pinned servant may still push along a file pin / stun along a diagonal pin
        const u64 s = p.pl[6 * T + SERVANT];
        const u64 last = T == 0 ? RANK_1 << 56 : RANK_1;
        const u64 pushers = s & (free | L.kf);
        const u64 west_stunners = s & (free | (T == 0 ? L.ka : L.kd)), east_stunners = s & (free | (T == 0 ? L.kd : L.ka));
        const u64 one = (T == 0 ? pushers << 8 : pushers >> 8) & empty;
        const u64 two = (T == 0 ? (one & (RANK_1 << 16)) << 8 : (one & (RANK_1 << 40)) >> 8) & empty & cm;
        if (two) out.put(two, rec_info(REC_DOUBLE, SERVANT, 0, ASC_NONE, 0, 0));
#pragma unroll
        for (int kind = REC_PUSH; kind <= REC_STUN_EAST; kind++) {
            if (kind == REC_DOUBLE) continue;
            const u64 targets = (kind == REC_PUSH ? one
                                 : kind == REC_STUN_WEST ? (T == 0 ? west_stunners << 7 : west_stunners >> 9) & ~FILE_H & opp
                                                         : (T == 0 ? east_stunners << 9 : east_stunners >> 7) & ~FILE_A & opp) & cm;
            if (targets & ~last) out.put(targets & ~last, rec_info((u32)kind, SERVANT, 0, ASC_NONE, 0, 0));
            if (targets & last) {
#pragma unroll
                for (u32 asc = PONY; asc <= PRINCESS; asc++) out.put(targets & last, rec_info((u32)kind, SERVANT, 0, asc, 0, 0));
            }
        }
        const u64 pb_bit = passing_by_bit(p.meta);
        if (pb_bit) // passing by: the make-the-move-and-test path
            for (u64 c = servant_attackers_of<T>(pb_bit) & s; c; c &= c - 1)
                if (passing_by_is_legal<T>(p, lsb64(c), lsb64(pb_bit))) out.put(pb_bit, rec_info(REC_PIECE, SERVANT, lsb64(c), ASC_NONE, 1, 0));
    }
    const u64 ok = ~own & cm;
    for (u64 s = p.pl[6 * T + PONY] & free; s; s &= s - 1) { // a pinned pony never stays on its line
        const int from = lsb64(s);
        const u64 targets = pony_moves(from) & ok;
        if (targets) out.put(targets, rec_info(REC_PIECE, PONY, from, ASC_NONE, 0, 0));
    }
#pragma unroll 1
    for (int job = SCHOLAR; job <= PRINCESS; job++) { // one loop body for the three sliders: less code to fetch
        // (the plane is picked by comparisons: a runtime index would push the position out of registers)
        for (u64 s = job == SCHOLAR ? p.pl[6 * T + SCHOLAR] : job == COP ? p.pl[6 * T + COP] : p.pl[6 * T + PRINCESS]; s; s &= s - 1) {
            const int from = lsb64(s);
            const u64 bit = 1ull << from;
            u64 targets = 0;
            if (job != COP) targets |= diag_attacks(from, occ);
            if (job != SCHOLAR) targets |= orth_attacks(from, occ);
            targets &= ok;
            if (bit & L.pinned) targets &= pin_line(L, bit);
            if (targets) out.put(targets, rec_info(REC_PIECE, (u32)job, from, ASC_NONE, 0, 0));
        }
    }
    if (cm == ~0ull) { // never out of check (life.rs:1034-1043)
        ServiceBits service;
        service_lookahead<T, false>(p, own, opp, service);
        if (service.to_bits) out.put(service.to_bits, rec_info(REC_SERVICE, FIGUREHEAD, (T == 0 ? 0 : 56) + 4, ASC_NONE, 0, 0));
    }
}
template <class Out> LL_HD void legal_records(const Pos &p, Out &out) {
    if (meta_stm(p.meta) == 0) legal_records_as<0>(p, out);
    else legal_records_as<1>(p, out);
}
// the move descriptor of the commit of a record that lands on `to`
LL_HD u32 rec_move_to(u32 info, int to, u32 team) {
    const u32 kind = info & 7u, job = (info >> 3) & 7u, asc = (info >> 12) & 7u;
    int from = (int)((info >> 6) & 63u);
    if (kind == REC_PUSH) from = team == 0 ? to - 8 : to + 8;
    else if (kind == REC_DOUBLE) from = team == 0 ? to - 16 : to + 16;
    else if (kind == REC_STUN_WEST) from = team == 0 ? to - 7 : to + 9;
    else if (kind == REC_STUN_EAST) from = team == 0 ? to - 9 : to + 7;
    const u32 flags = ((info >> 15) & 1u ? MV_PASSING_BY : 0u) | (kind == REC_SERVICE ? MV_SERVICE : 0u);
    return mv_pack(job, from, to, asc, flags);
}

// The position of a commit in lookahead()'s order among the commits of ONE position, as a sortable key computed from
// the move alone (life.rs:1052-1084): servants, ponies, scholars, cops, princesses (all diagonal moves, then all
// orthogonal ones), figurehead, secret service West then East; pieces by ascending square; a servant's push, double
// push, west stun, east stun, each through Pony, Scholar, Cop, Princess when ascending; table pieces by ascending
// destination; sliders by the reference's direction lists (life.rs:10-11), nearest square first.
LL_HD u32 canonical_key(u32 m) {
    const int from = mv_from(m), to = mv_to(m);
    const u32 job = mv_job(m), asc = mv_asc(m);
    const int dr = (to >> 3) - (from >> 3), df = (to & 7) - (from & 7);
    const u32 dist = (u32)((dr < 0 ? -dr : dr) > (df < 0 ? -df : df) ? (dr < 0 ? -dr : dr) : (df < 0 ? -df : df));
    const bool diagonal = dr != 0 && df != 0;
    u32 cls = job <= COP ? job : job == PRINCESS ? (diagonal ? 4u : 5u) : (m & MV_SERVICE) ? 7u : 6u;
    u32 sub;
    if (job == SERVANT) sub = ((df == 0 ? (dist == 1 ? 0u : 1u) : df < 0 ? 2u : 3u) << 3) | (asc == ASC_NONE ? 0u : asc);
    else if (job == PONY || job == FIGUREHEAD) sub = (u32)to;
    else if (diagonal) sub = ((u32)((dr > 0 ? 2 : 0) + (df > 0 ? 1 : 0)) << 3) | dist; // (-1,-1), (-1,1), (1,-1), (1,1)
    else sub = ((dr < 0 ? 0u : dr > 0 ? 1u : df < 0 ? 2u : 3u) << 3) | dist;            // (-1,0), (1,0), (0,-1), (0,1)
    return (cls << 16) | ((u32)from << 8) | sub;
}

// ---- domains (include/leafline_b200.h) -----------------------------------------------------------------------------------
// FAST     domain W: what the set-wise kernels assume -- disjoint planes, at most one figurehead and at most 16 figurines a
//          team, a set eligibility bit means that team's figurehead (if any) stands on its home square and either the cop
//          stands on that corner or the team has at most 15 figurines (secret service conjures a cop where none stands,
//          life.rs:621-628: the sixteenth figurine must not become a seventeenth), passing-by square consistent.  Closed
//          under the reference's own generator.
// LITERAL  everything else the reference answers and this library can represent: disjoint planes and a consistent
//          passing-by square, any number of figureheads and figurines, any eligibility.  Answered by the literal path:
//          the canonical generator with the reference's own legality (every commit made, kept iff no figurehead of the
//          mover is endangered afterwards, life.rs:678-699), one thread per position, level by level.
// REJECT   overlapping planes, ABI bytes out of range, a passing-by square that is off the board, occupied or not directly
//          behind an enemy servant on its double-step rank (LL_ERR_DOMAIN).
enum { DOMAIN_FAST = 0, DOMAIN_LITERAL = 1, DOMAIN_REJECT = 2 };
LL_HD int domain_class(const Pos &p) {
    if (p.meta & META_MALFORMED) return DOMAIN_REJECT;
    u64 acc = 0, overlap = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        overlap |= acc & p.pl[j];
        acc |= p.pl[j];
    }
    if (overlap) return DOMAIN_REJECT;
    const u32 pb = meta_pb(p.meta);
    if (pb != 0xFFu) {
        if ((pb >> 4) > 7 || (pb & 0xFu) > 7) return DOMAIN_REJECT;
        const int sq = loc_to_sq(pb);
        const u64 bit = 1ull << sq;
        if (acc & bit) return DOMAIN_REJECT;
        if (meta_stm(p.meta) == 0) { // Orange to move: a Blue servant just double-stepped over rank 5
            if ((sq >> 3) != 5 || !(p.pl[6 + SERVANT] & (bit >> 8)) || (acc & (bit << 8))) return DOMAIN_REJECT;
        } else {
            if ((sq >> 3) != 2 || !(p.pl[SERVANT] & (bit << 8)) || (acc & (bit >> 8))) return DOMAIN_REJECT;
        }
    }
    if (popc64(p.pl[FIGUREHEAD]) > 1 || popc64(p.pl[6 + FIGUREHEAD]) > 1) return DOMAIN_LITERAL;
    const int n_orange = popc64(occupied_by<0>(p)), n_blue = popc64(occupied_by<1>(p));
    if (n_orange > 16 || n_blue > 16) return DOMAIN_LITERAL; // bounds the record lists (REC_MAX) and the warp-per-parent lanes
    const u32 elig = meta_elig(p.meta);
    if ((elig & 0x3u) && (p.pl[FIGUREHEAD] & ~(1ull << 4))) return DOMAIN_LITERAL;
    if ((elig & 0xCu) && (p.pl[6 + FIGUREHEAD] & ~(1ull << 60))) return DOMAIN_LITERAL;
    if (n_orange == 16 && (((elig & 0x1u) && !(p.pl[COP] & (1ull << 0))) || ((elig & 0x2u) && !(p.pl[COP] & (1ull << 7))))) return DOMAIN_LITERAL;
    if (n_blue == 16 && (((elig & 0x4u) && !(p.pl[6 + COP] & (1ull << 56))) || ((elig & 0x8u) && !(p.pl[6 + COP] & (1ull << 63))))) return DOMAIN_LITERAL;
    return DOMAIN_FAST;
}
LL_HD bool well_formed(const Pos &p) { return domain_class(p) == DOMAIN_FAST; }

// ---- the literal path: lookahead() exactly as the reference defines it (life.rs:692-699, 728-739, 1052-1084) ---------------
// The reckless generator above is already literal on any position of disjoint planes (it loops over every figurehead, asks
// no question about the number of figurines, reproduces the secret-service rules square by square); what domain W buys is
// only the legality by masks.  Here every reckless commit is made and kept iff none of the mover's figureheads is endangered
// afterwards; in_critical_endangerment / attacked_by are the opponent's whole reckless_lookahead boiled down to "is this
// square the destination of a capture-capable move", which holds on any disjoint position (the child's passing-by square, the
// only way to stun something that is not on the destination, always lies behind the servant that just stepped).
template <int T, class Sink> struct CarefulSink {
    const Pos &par;
    Sink &inner;
    LL_HDM void move(u32 job, int from, int to, u32 asc, u32 flags) {
        Pos c;
        make_child_as<T>(par, mv_pack(job, from, to, asc, flags), c);
        if (!in_critical_endangerment<T>(c)) inner.move(job, from, to, asc, flags);
    }
};
template <bool NIHIL, class Sink> LL_HD void lookahead_literal(const Pos &p, Sink &sink) {
    if (NIHIL) {
        lookahead<true>(p, sink);
    } else if (meta_stm(p.meta) == 0) {
        CarefulSink<0, Sink> careful{p, sink};
        lookahead_as<0, true>(p, careful);
    } else {
        CarefulSink<1, Sink> careful{p, sink};
        lookahead_as<1, true>(p, careful);
    }
}
template <bool NIHIL> LL_HD u32 count_moves_literal(const Pos &p) {
    CountSink sink;
    lookahead_literal<NIHIL>(p, sink);
    return sink.n;
}

// ---- static evaluation (mind.rs:50-143): strict f32, one rounded multiply and one rounded add per
// reference statement, in the reference's order; fmul/fadd are never contracted to FMA -----------------
LL_HD float score(const Pos &p) {
    const float values[6] = {1.0f, 3.2f, 3.3f, 5.1f, 8.8f, 20000.0f}; // mind.rs:36-48
    float v = 0.0f;
    v = fadd(v, fmul(0.5f, meta_stm(p.meta) ? -1.0f : 1.0f)); // mind.rs:53
#pragma unroll
    for (int team = 0; team < 2; team++) { // mind.rs:55-68
        const float orient = team ? -1.0f : 1.0f;
#pragma unroll
        for (int job = 0; job < 6; job++)
            v = fadd(v, fmul((float)popc64(p.pl[6 * team + job]), fmul(orient, values[job])));
        if (popc64(p.pl[6 * team + SCHOLAR]) >= 2) v = fadd(v, fmul(orient, 0.5f));
    }
    const u64 center = 0x00003C3C3C3C0000ull; // mind.rs:70-81
    const int oc = popc64((p.pl[SERVANT] | p.pl[PONY]) & center), bc = popc64((p.pl[6 + SERVANT] | p.pl[6 + PONY]) & center);
    v = fadd(v, fmul(0.1f, (float)(oc - bc)));
    const u64 high_seventh = 0xFFull << 48, low_seventh = 0xFFull << 8; // mind.rs:83-89
    v = fadd(v, fmul(0.5f, (float)popc64(p.pl[COP] & high_seventh)));
    v = fadd(v, -fmul(0.5f, (float)popc64(p.pl[6 + COP] & low_seventh)));
    // mind.rs:91-111: doubled servants, file by file, Orange then Blue.  Files without a second
    // servant contribute nothing, so only those files are visited (ascending, same add order).
    {
        const u64 os = p.pl[SERVANT], bs = p.pl[6 + SERVANT];
        u32 o_once = 0, o_twice = 0, b_once = 0, b_twice = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const u32 ob = (u32)(os >> (8 * r)) & 0xFFu, bb = (u32)(bs >> (8 * r)) & 0xFFu;
            o_twice |= o_once & ob;
            o_once |= ob;
            b_twice |= b_once & bb;
            b_once |= bb;
        }
        u32 files = o_twice | b_twice;
        while (files) {
            const int f = lsb32(files);
            files &= files - 1;
            const u64 file = FILE_A << f;
            const int o = popc64(os & file), b = popc64(bs & file);
            if (o > 1) v = fadd(v, -fmul(0.5f, (float)(o - 1)));
            if (b > 1) v = fadd(v, fmul(0.5f, (float)(b - 1)));
        }
    }
    // mind.rs:113-131
    v = fadd(v, fmul(1.8f, (float)popc64(p.pl[SERVANT] & high_seventh)));
    v = fadd(v, fmul(0.6f, (float)popc64(p.pl[SERVANT] & (0xFFull << 40))));
    v = fadd(v, -fmul(1.8f, (float)popc64(p.pl[6 + SERVANT] & low_seventh)));
    v = fadd(v, -fmul(0.6f, (float)popc64(p.pl[6 + SERVANT] & (0xFFull << 16))));
    const u32 elig = meta_elig(p.meta); // mind.rs:133-140
    if (elig & 0x3u) v = fadd(v, 0.1f);
    if (elig & 0xCu) v = fadd(v, -0.1f);
    return v;
}

// ---- the same evaluation from FEATURES ---------------------------------------------------------------------------
// score() reads a position only through small integers -- twelve head counts, two centre counts, the cops and
// servants on a few ranks, the servants per file -- and turns them into floats in a fixed order.  A child differs
// from its parent in at most three squares, so its integers follow from the parent's by a few +-1: the horizon kernel
// (one child per thread, never stored) keeps the parent's features in shared memory and evaluates
// score_features(child_features(...)) instead of building twelve planes and counting them again.  Same integers, same
// float sequence: bit-identical to score(make_child(...)) (tests/hostcheck checks every commit of every test position).
struct Feat {
    u32 w[7]; // w0-w2: head counts, a byte per plane (plane j in byte j&3 of word j>>2)
              // w3: bytes oc, bc (servants and ponies in the centre), orange cops on rank 6, blue cops on rank 1
              // w4: bytes orange servants on rank 6, on rank 5, blue servants on rank 1, on rank 2
              // w5, w6: orange / blue servants per file, a nibble per file
};
LL_HD Feat features_of(const Pos &p) {
    Feat f;
#pragma unroll
    for (int w = 0; w < 3; w++)
        f.w[w] = (u32)popc64(p.pl[4 * w]) | ((u32)popc64(p.pl[4 * w + 1]) << 8) | ((u32)popc64(p.pl[4 * w + 2]) << 16) | ((u32)popc64(p.pl[4 * w + 3]) << 24);
    const u64 center = 0x00003C3C3C3C0000ull, high_seventh = 0xFFull << 48, low_seventh = 0xFFull << 8;
    const u64 os = p.pl[SERVANT], bs = p.pl[6 + SERVANT];
    f.w[3] = (u32)popc64((os | p.pl[PONY]) & center) | ((u32)popc64((bs | p.pl[6 + PONY]) & center) << 8) |
             ((u32)popc64(p.pl[COP] & high_seventh) << 16) | ((u32)popc64(p.pl[6 + COP] & low_seventh) << 24);
    f.w[4] = (u32)popc64(os & high_seventh) | ((u32)popc64(os & (0xFFull << 40)) << 8) | ((u32)popc64(bs & low_seventh) << 16) |
             ((u32)popc64(bs & (0xFFull << 16)) << 24);
    u32 of = 0, bf = 0;
#pragma unroll
    for (int file = 0; file < 8; file++) {
        of |= (u32)popc64(os & (FILE_A << file)) << (4 * file);
        bf |= (u32)popc64(bs & (FILE_A << file)) << (4 * file);
    }
    f.w[5] = of;
    f.w[6] = bf;
    return f;
}
// score_features in two halves, the float sequence unchanged: the material terms (mind.rs:53-68) read the twelve head
// counts only -- a child that stuns nobody, ascends nothing and conjures nothing has its parent's, so the horizon kernel
// computes this half once per PARENT for all such children -- and the positional terms (mind.rs:70-140) continue from it.
LL_HD float score_material(const Feat &f, u32 stm) {
    const float values[6] = {1.0f, 3.2f, 3.3f, 5.1f, 8.8f, 20000.0f}; // mind.rs:36-48
    float v = 0.0f;
    v = fadd(v, fmul(0.5f, stm ? -1.0f : 1.0f)); // mind.rs:53
#pragma unroll
    for (int team = 0; team < 2; team++) { // mind.rs:55-68
        const float orient = team ? -1.0f : 1.0f;
#pragma unroll
        for (int job = 0; job < 6; job++) {
            const int j = 6 * team + job;
            v = fadd(v, fmul((float)((f.w[j >> 2] >> (8 * (j & 3))) & 0xFFu), fmul(orient, values[job])));
        }
        const int j = 6 * team + SCHOLAR;
        if (((f.w[j >> 2] >> (8 * (j & 3))) & 0xFFu) >= 2) v = fadd(v, fmul(orient, 0.5f));
    }
    return v;
}
LL_HD float score_eligibility(float v, u32 elig) { // mind.rs:133-140, the last terms of the sum
    if (elig & 0x3u) v = fadd(v, 0.1f);
    if (elig & 0xCu) v = fadd(v, -0.1f);
    return v;
}
LL_HD float score_positional(float v, u32 w3, u32 w4, u32 w5, u32 w6, u32 elig) {
    const int oc = (int)(w3 & 0xFFu), bc = (int)((w3 >> 8) & 0xFFu); // mind.rs:70-81
    v = fadd(v, fmul(0.1f, (float)(oc - bc)));
    v = fadd(v, fmul(0.5f, (float)((w3 >> 16) & 0xFFu))); // mind.rs:83-89
    v = fadd(v, -fmul(0.5f, (float)(w3 >> 24)));
    { // mind.rs:91-111: only files with a second servant add anything; ascending, Orange before Blue
        u32 files = (w5 | w6) & 0xEEEEEEEEu; // a nibble >= 2 has one of its upper three bits set
        while (files) {
            const int sh = lsb32(files) & ~3;
            files &= ~(0xFu << sh);
            const int o = (int)((w5 >> sh) & 0xFu), b = (int)((w6 >> sh) & 0xFu);
            if (o > 1) v = fadd(v, -fmul(0.5f, (float)(o - 1)));
            if (b > 1) v = fadd(v, fmul(0.5f, (float)(b - 1)));
        }
    }
    v = fadd(v, fmul(1.8f, (float)(w4 & 0xFFu))); // mind.rs:113-131
    v = fadd(v, fmul(0.6f, (float)((w4 >> 8) & 0xFFu)));
    v = fadd(v, -fmul(1.8f, (float)((w4 >> 16) & 0xFFu)));
    v = fadd(v, -fmul(0.6f, (float)(w4 >> 24)));
    return score_eligibility(v, elig);
}
LL_HD float score_features(const Feat &f, u32 stm, u32 elig) {
    return score_positional(score_material(f, stm), f.w[3], f.w[4], f.w[5], f.w[6], elig);
}
// The score of the child that a QUIET commit makes (nobody stunned, nothing ascended, no secret service, no passing by:
// the head counts and the servants' files are the parent's): `material` = score_material(parent's features, child's
// initiative), w3 / w4 / elig the parent's, updated here by the move alone.  Same integers, same float sequence as
// score_features(child_features(...)) -- tests/hostcheck compares the two on every quiet commit of every test position.
template <int T> LL_HD float quiet_child_score(float material, u32 w3, u32 w4, u32 w5, u32 w6, u32 elig, u32 job, int from, int to) {
    const u64 center = 0x00003C3C3C3C0000ull;
    if (job == FIGUREHEAD) elig &= T == 0 ? ~0x3u : ~0xCu; // life.rs:577-605
    if (job == COP) {
        const int file = from & 7;
        if (file == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (file == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
        const int cop_rank = T == 0 ? 6 : 1;
        w3 += ((u32)((to >> 3) == cop_rank) - (u32)((from >> 3) == cop_rank)) << (16 + 8 * T);
    }
    if (job <= PONY) w3 += ((u32)((center >> to) & 1u) - (u32)((center >> from) & 1u)) << (8 * T);
    if (job == SERVANT) { // a push: the file stays, the rank counts move
        const int rf = from >> 3, rt = to >> 3, near_t = T == 0 ? 6 : 1, far_t = T == 0 ? 5 : 2;
        w4 += (((u32)(rt == near_t) - (u32)(rf == near_t)) << (16 * T)) + (((u32)(rt == far_t) - (u32)(rf == far_t)) << (16 * T + 8));
    }
    return score_positional(material, w3, w4, w5, w6, elig);
}
// the same with the mover as a run-time value: ONE copy of the code in the horizon kernel (its tile loop is bound by instruction fetch)
LL_HD float quiet_child_score_rt(const int T, float material, u32 w3, u32 w4, u32 w5, u32 w6, u32 elig, u32 job, int from, int to) {
    const u64 center = 0x00003C3C3C3C0000ull;
    if (job == FIGUREHEAD) elig &= T == 0 ? ~0x3u : ~0xCu; // life.rs:577-605
    if (job == COP) {
        const int file = from & 7;
        if (file == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (file == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
        const int cop_rank = T == 0 ? 6 : 1;
        w3 += ((u32)((to >> 3) == cop_rank) - (u32)((from >> 3) == cop_rank)) << (16 + 8 * T);
    }
    if (job <= PONY) w3 += ((u32)((center >> to) & 1u) - (u32)((center >> from) & 1u)) << (8 * T);
    if (job == SERVANT) { // a push: the file stays, the rank counts move
        const int rf = from >> 3, rt = to >> 3, near_t = T == 0 ? 6 : 1, far_t = T == 0 ? 5 : 2;
        w4 += (((u32)(rt == near_t) - (u32)(rf == near_t)) << (16 * T)) + (((u32)(rt == far_t) - (u32)(rf == far_t)) << (16 * T + 8));
    }
    return score_positional(material, w3, w4, w5, w6, elig);
}
// The quiet children of ONE parent differ from it in few ways: what a quiet commit can do to w3 / w4 is one of
//   0 nothing (a scholar, a princess, the figurehead; a cop along its rank or between other ranks; a servant or pony that
//     neither enters nor leaves the centre, nor reaches the opposition's third or second rank)
//   1 / 2 a servant or pony enters / leaves the centre      3 / 4 a cop enters / leaves the opposition's second rank
//   5 a servant reaches the opposition's third rank (from the centre's last rank into it or beside it: centre count unchanged)
//   6 / 7 a servant steps from the third to the second rank, beside the centre / out of it
// -- so the horizon kernel sums mind.rs:53-131 once per (parent, way), QUIET_WAYS times a parent instead of once a child, and
// a quiet child adds the eligibility terms (the last of the sum) to the entry of its way: same integers, same float sequence.
constexpr u32 QUIET_WAYS = 8;
LL_HD u32 quiet_way_rt(const int T, u32 job, int from, int to) {
    const u64 center = 0x00003C3C3C3C0000ull;
    const int dc = job <= PONY ? (int)((center >> to) & 1u) - (int)((center >> from) & 1u) : 0;
    u32 way = dc > 0 ? 1u : dc < 0 ? 2u : 0u;
    if (job == COP) {
        const int cop_rank = T == 0 ? 6 : 1;
        const int dr = (int)((to >> 3) == cop_rank) - (int)((from >> 3) == cop_rank);
        way = dr > 0 ? 3u : dr < 0 ? 4u : 0u;
    }
    if (job == SERVANT) { // a push: towards the far side, one rank or two
        const int rt = to >> 3, near_t = T == 0 ? 6 : 1, far_t = T == 0 ? 5 : 2;
        if (rt == far_t) way = 5u;
        if (rt == near_t) way = dc < 0 ? 7u : 6u;
    }
    return way;
}
LL_HD void quiet_way_features(const int T, u32 way, u32 &w3, u32 &w4) { // the parent's w3 / w4 after a quiet commit of this way
    const u32 dc = way == 1 ? 1u : (way == 2 || way == 7) ? ~0u : 0u, dr = way == 3 ? 1u : way == 4 ? ~0u : 0u;
    w3 += (dc << (8 * T)) + (dr << (16 + 8 * T));
    if (way == 5) w4 += 1u << (16 * T + 8);
    if (way >= 6) w4 += (1u << (16 * T)) - (1u << (16 * T + 8));
}
LL_HD u32 quiet_eligibility_rt(const int T, u32 elig, u32 job, int from) { // life.rs:577-605
    if (job == FIGUREHEAD) elig &= T == 0 ? ~0x3u : ~0xCu;
    if (job == COP) {
        const int file = from & 7;
        if (file == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (file == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
    }
    return elig;
}
// the commits of a record that stun nobody, ascend nothing and are no secret service: a piece's targets on empty squares, the
// servants' plain steps
LL_HD u64 rec_quiet(u32 info, u64 mask, u64 opp) {
    const u32 kind = info & 7u, asc = (info >> 12) & 7u;
    if (kind == REC_PIECE) return (info >> 15) & 1u ? 0ull : mask & ~opp;
    return (kind == REC_PUSH || kind == REC_DOUBLE) && asc == ASC_NONE ? mask : 0ull;
}
// The ways (bit w = way w of quiet_way_rt) that the quiet commits of ONE record take, from the set of their destinations
// `quiet` (!= 0) -- a record is one piece's targets or one kind of step of all the servants, so the way depends on the
// destination alone and the question "does any commit of the record go way w" is a mask test.  All quiet commits of a
// record of one way score the same (same features, same eligibility: quiet_eligibility_rt reads the job and the origin's
// file only, and a servant's step changes no eligibility), so the best of a record's quiet children is the best of its ways.
LL_HD u32 quiet_ways_of_record(const int T, u32 info, u64 quiet) {
    const u64 center = 0x00003C3C3C3C0000ull;
    const u32 kind = info & 7u, job = (info >> 3) & 7u;
    if (kind == REC_PIECE) {
        const int from = (int)((info >> 6) & 63u);
        if (job == PONY) {
            const bool inside = (center >> from) & 1u;
            return ((quiet & center) ? (inside ? 1u : 1u << 1) : 0u) | ((quiet & ~center) ? (inside ? 1u << 2 : 1u) : 0u);
        }
        if (job == COP) {
            const int cop_rank = T == 0 ? 6 : 1;
            const u64 there = RANK_1 << (8 * cop_rank);
            const bool on = (from >> 3) == cop_rank;
            return ((quiet & there) ? (on ? 1u : 1u << 3) : 0u) | ((quiet & ~there) ? (on ? 1u << 4 : 1u) : 0u);
        }
        return 1u; // scholar, princess, figurehead
    }
    // a step of one rank (REC_PUSH) or two (REC_DOUBLE) towards the far side: `origin_inside` = the destinations whose origin is in the centre
    const int d = kind == REC_DOUBLE ? 16 : 8;
    const u64 origin_inside = T == 0 ? center << d : center >> d;
    const u64 enters = quiet & center & ~origin_inside, leaves = quiet & ~center & origin_inside;
    const u64 far = RANK_1 << (8 * (T == 0 ? 5 : 2)), near = RANK_1 << (8 * (T == 0 ? 6 : 1));
    const u64 rest = quiet & ~(far | near);
    return ((quiet & far) ? 1u << 5 : 0u) | ((quiet & near & leaves) ? 1u << 7 : 0u) | ((quiet & near & ~leaves) ? 1u << 6 : 0u) |
           ((rest & enters) ? 1u << 1 : 0u) | ((rest & leaves) ? 1u << 2 : 0u) | ((rest & ~(enters | leaves)) ? 1u : 0u);
}
// mind.rs:70-131 for N ways of one parent at once (the files' terms are the same for all of them), each sum in score_positional's order
template <int N> LL_HD void score_ways(float material, const u32 (&w3)[N], const u32 (&w4)[N], u32 w5, u32 w6, float (&v)[N]) {
#pragma unroll
    for (int u = 0; u < N; u++) {
        const int oc = (int)(w3[u] & 0xFFu), bc = (int)((w3[u] >> 8) & 0xFFu);
        v[u] = fadd(material, fmul(0.1f, (float)(oc - bc)));
        v[u] = fadd(v[u], fmul(0.5f, (float)((w3[u] >> 16) & 0xFFu)));
        v[u] = fadd(v[u], -fmul(0.5f, (float)(w3[u] >> 24)));
    }
    u32 files = (w5 | w6) & 0xEEEEEEEEu;
    while (files) {
        const int sh = lsb32(files) & ~3;
        files &= ~(0xFu << sh);
        const int o = (int)((w5 >> sh) & 0xFu), b = (int)((w6 >> sh) & 0xFu);
        if (o > 1) {
            const float t = -fmul(0.5f, (float)(o - 1));
#pragma unroll
            for (int u = 0; u < N; u++) v[u] = fadd(v[u], t);
        }
        if (b > 1) {
            const float t = fmul(0.5f, (float)(b - 1));
#pragma unroll
            for (int u = 0; u < N; u++) v[u] = fadd(v[u], t);
        }
    }
#pragma unroll
    for (int u = 0; u < N; u++) {
        v[u] = fadd(v[u], fmul(1.8f, (float)(w4[u] & 0xFFu)));
        v[u] = fadd(v[u], fmul(0.6f, (float)((w4[u] >> 8) & 0xFFu)));
        v[u] = fadd(v[u], -fmul(1.8f, (float)((w4[u] >> 16) & 0xFFu)));
        v[u] = fadd(v[u], -fmul(0.6f, (float)(w4[u] >> 24)));
    }
}
LL_HD void feat_count(Feat &f, u32 plane, u32 delta) { // head count of `plane` += delta (two's complement for -1)
    const u32 d = delta << (8 * (plane & 3u));
    f.w[0] += (plane >> 2) == 0 ? d : 0u;
    f.w[1] += (plane >> 2) == 1 ? d : 0u;
    f.w[2] += (plane >> 2) == 2 ? d : 0u;
}
// Features, eligibility and hospitalized job of the child that commit `m` (mover T) makes of a parent with features
// `f` and eligibility `elig`.  Planes: u64 operator()(int plane) gives the PARENT's planes (only the opposition's six
// are always read; the mover's cop and figurehead planes only by a secret service) -- life.rs:569-676 through the
// integers score() reads, see make_child_as for the same steps on the planes.
template <int T, class Planes> LL_HD u32 child_features(Feat &f, u32 &elig, const Planes &planes, u32 m) {
    constexpr int E = 1 - T;
    const int from = mv_from(m), to = mv_to(m);
    const u32 job = mv_job(m), asc = mv_asc(m);
    const u64 center = 0x00003C3C3C3C0000ull;
    const u64 tb = 1ull << to;
    const u64 victim = (m & MV_PASSING_BY) ? (T == 0 ? tb >> 8 : tb << 8) : tb;
    const int vs = (m & MV_PASSING_BY) ? (T == 0 ? to - 8 : to + 8) : to;
    u32 hosp = JOB_NONE;
#pragma unroll
    for (int j = 5; j >= 0; j--)
        if (planes(6 * E + j) & victim) hosp = (u32)j; // first in dramatis_personae order wins (life.rs:550-557)
    // ---- eligibility (life.rs:577-605)
    if (job == FIGUREHEAD) elig &= T == 0 ? ~0x3u : ~0xCu;
    if (job == COP) {
        const int file = from & 7;
        if (file == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (file == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
    }
    // ---- head counts
    if (hosp != JOB_NONE) feat_count(f, 6u * E + hosp, 0xFFFFFFFFu);
    if (asc != ASC_NONE) { // the servant alights as its new job (life.rs:701-726)
        feat_count(f, 6u * T + SERVANT, 0xFFFFFFFFu);
        feat_count(f, 6u * T + asc, 1u);
    }
    if (m & MV_SERVICE) { // the transits happen whether or not somebody is there (life.rs:573-575, 610-629)
        const int rank8 = to & 56;
        const u64 corner = 1ull << (rank8 + ((to & 7) == 6 ? 7 : 0));
        if (!(planes(6 * T + COP) & corner)) feat_count(f, 6u * T + COP, 1u);          // a cop out of thin air
        if (!(planes(6 * T + FIGUREHEAD) & (1ull << from))) feat_count(f, 6u * T + FIGUREHEAD, 1u); // and a figurehead
    }
    // ---- servants and ponies in the centre (never on a first or last rank: ascensions need no special case)
    const u32 in_from = (u32)(center >> from) & 1u, in_to = (u32)(center >> to) & 1u, in_vs = (u32)(center >> vs) & 1u;
    if (job <= PONY) f.w[3] += (in_to - in_from) << (8 * T);
    if (hosp <= PONY) f.w[3] -= in_vs << (8 * E);
    // ---- cops on the opposition's servant rank
    const int rf = from >> 3, rt = to >> 3, rv = vs >> 3;
    const int cop_rank_t = T == 0 ? 6 : 1, cop_rank_e = E == 0 ? 6 : 1;
    if (job == COP) f.w[3] += ((u32)(rt == cop_rank_t) - (u32)(rf == cop_rank_t)) << (16 + 8 * T);
    if (hosp == COP) f.w[3] -= (u32)(rv == cop_rank_e) << (16 + 8 * E);
    // ---- servants by rank and by file (an ascending servant leaves the board: rank 7 / 0 is not counted, its file is not re-entered)
    const int near_t = T == 0 ? 6 : 1, far_t = T == 0 ? 5 : 2, near_e = E == 0 ? 6 : 1, far_e = E == 0 ? 5 : 2;
    if (job == SERVANT) {
        f.w[4] += (((u32)(rt == near_t) - (u32)(rf == near_t)) << (16 * T)) + (((u32)(rt == far_t) - (u32)(rf == far_t)) << (16 * T + 8));
        f.w[5 + T] -= 1u << (4 * (from & 7));
        if (asc == ASC_NONE) f.w[5 + T] += 1u << (4 * (to & 7));
    }
    if (hosp == SERVANT) {
        f.w[4] -= ((u32)(rv == near_e) << (16 * E)) + ((u32)(rv == far_e) << (16 * E + 8));
        f.w[5 + E] -= 1u << (4 * (vs & 7));
    }
    return hosp;
}

// the same with the mover as a run-time value (see quiet_child_score_rt).  Planes: besides operator() (the mover's cop and
// figurehead planes, for a secret service), u32 opposition_job_at(int square) -- the job of the opposition's figurine there, or JOB_NONE
template <class Planes> LL_HD u32 child_features_rt(const int T, Feat &f, u32 &elig, const Planes &planes, u32 m) {
    const int E = 1 - T;
    const int from = mv_from(m), to = mv_to(m);
    const u32 job = mv_job(m), asc = mv_asc(m);
    const u64 center = 0x00003C3C3C3C0000ull;
    const int vs = (m & MV_PASSING_BY) ? (T == 0 ? to - 8 : to + 8) : to;
    const u32 hosp = planes.opposition_job_at(vs); // whom the commit stuns, if anybody (life.rs:550-557)
    // ---- eligibility (life.rs:577-605)
    if (job == FIGUREHEAD) elig &= T == 0 ? ~0x3u : ~0xCu;
    if (job == COP) {
        const int file = from & 7;
        if (file == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (file == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
    }
    // ---- head counts
    if (hosp != JOB_NONE) feat_count(f, 6u * E + hosp, 0xFFFFFFFFu);
    if (asc != ASC_NONE) { // the servant alights as its new job (life.rs:701-726)
        feat_count(f, 6u * T + SERVANT, 0xFFFFFFFFu);
        feat_count(f, 6u * T + asc, 1u);
    }
    if (m & MV_SERVICE) { // the transits happen whether or not somebody is there (life.rs:573-575, 610-629)
        const int rank8 = to & 56;
        const u64 corner = 1ull << (rank8 + ((to & 7) == 6 ? 7 : 0));
        if (!(planes(6 * T + COP) & corner)) feat_count(f, 6u * T + COP, 1u);          // a cop out of thin air
        if (!(planes(6 * T + FIGUREHEAD) & (1ull << from))) feat_count(f, 6u * T + FIGUREHEAD, 1u); // and a figurehead
    }
    // ---- servants and ponies in the centre (never on a first or last rank: ascensions need no special case)
    const u32 in_from = (u32)(center >> from) & 1u, in_to = (u32)(center >> to) & 1u, in_vs = (u32)(center >> vs) & 1u;
    if (job <= PONY) f.w[3] += (in_to - in_from) << (8 * T);
    if (hosp <= PONY) f.w[3] -= in_vs << (8 * E);
    // ---- cops on the opposition's servant rank
    const int rf = from >> 3, rt = to >> 3, rv = vs >> 3;
    const int cop_rank_t = T == 0 ? 6 : 1, cop_rank_e = E == 0 ? 6 : 1;
    if (job == COP) f.w[3] += ((u32)(rt == cop_rank_t) - (u32)(rf == cop_rank_t)) << (16 + 8 * T);
    if (hosp == COP) f.w[3] -= (u32)(rv == cop_rank_e) << (16 + 8 * E);
    // ---- servants by rank and by file (an ascending servant leaves the board: rank 7 / 0 is not counted, its file is not re-entered)
    const int near_t = T == 0 ? 6 : 1, far_t = T == 0 ? 5 : 2, near_e = E == 0 ? 6 : 1, far_e = E == 0 ? 5 : 2;
    if (job == SERVANT) {
        f.w[4] += (((u32)(rt == near_t) - (u32)(rf == near_t)) << (16 * T)) + (((u32)(rt == far_t) - (u32)(rf == far_t)) << (16 * T + 8));
        const u32 d = (asc == ASC_NONE ? 1u << (4 * (to & 7)) : 0u) - (1u << (4 * (from & 7))); // (no run-time index into f.w: it would live in local memory)
        f.w[5] += T == 0 ? d : 0u;
        f.w[6] += T == 0 ? 0u : d;
    }
    if (hosp == SERVANT) {
        f.w[4] -= ((u32)(rv == near_e) << (16 * E)) + ((u32)(rv == far_e) << (16 * E + 8));
        f.w[5] -= E == 0 ? 1u << (4 * (vs & 7)) : 0u;
        f.w[6] -= E == 0 ? 0u : 1u << (4 * (vs & 7));
    }
    return hosp;
}

} // namespace ll
