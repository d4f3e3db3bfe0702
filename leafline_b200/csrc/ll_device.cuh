// ll_device.cuh -- device-side position type, bitboard primitives and the canonical move generator.
//
// Everything here is hand-written for sm_100a.  A position lives in registers as 12 u64 planes
// (the reference's Pinfields, src/space.rs:127-200) plus a u32 of metadata; in HBM it is stored
// structure-of-arrays (see Soa below).  The canonical generator emits commits in exactly the
// reference's order (src/life.rs:1052-1084); legality is decided by building the
// child and testing whether the mover's figurehead square is attacked, which on well-formed
// positions equals the reference's "run the opponent's full reckless_lookahead and look for a
// figurehead hospitalization" (src/life.rs:678-699).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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

// meta: bit0 initiative (0 Orange, 1 Blue); bits 1-4 service_eligibility (life.rs:178-181);
// bits 8-15 passing_by Locale byte (rank<<4|file, 0xFF = None) -- same byte as ll_world_t.
struct Pos {
    u64 pl[12];
    u32 meta;
};

__device__ __forceinline__ u32 meta_stm(u32 m) { return m & 1u; }
__device__ __forceinline__ u32 meta_elig(u32 m) { return (m >> 1) & 0xFu; }
__device__ __forceinline__ u32 meta_pb(u32 m) { return (m >> 8) & 0xFFu; }
__device__ __forceinline__ u32 make_meta(u32 stm, u32 elig, u32 pb) { return stm | (elig << 1) | (pb << 8); }
__device__ __forceinline__ int loc_to_sq(u32 loc) { return (int)(((loc >> 4) << 3) | (loc & 7u)); }
__device__ __forceinline__ u32 sq_to_loc(int sq) { return (u32)(((sq >> 3) << 4) | (sq & 7)); }

// ---- structure-of-arrays storage in HBM -------------------------------------------------------
// plane j of position i at pl[j*stride + i]; stride % 2 == 0 keeps every plane 16-byte aligned so a
// thread can move two neighbouring positions with one 128-bit access.
struct Soa {
    u64 *pl;
    u32 *meta;
    size_t stride;
};

__device__ __forceinline__ Pos soa_load(const Soa &s, size_t i) {
    Pos p;
#pragma unroll
    for (int j = 0; j < 12; j++) p.pl[j] = __ldg(s.pl + (size_t)j * s.stride + i);
    p.meta = __ldg(s.meta + i);
    return p;
}
__device__ __forceinline__ void soa_store(const Soa &s, size_t i, const Pos &p) {
#pragma unroll
    for (int j = 0; j < 12; j++) s.pl[(size_t)j * s.stride + i] = p.pl[j];
    s.meta[i] = p.meta;
}

// ---- generated tables (furniture/tablemaker.py:41-62), filled by the host at ll_init ----------
__device__ u64 d_pony_table[64];
__device__ u64 d_figurehead_table[64];

__device__ __forceinline__ u64 pony_moves(int sq) { return __ldg(&d_pony_table[sq]); }
__device__ __forceinline__ u64 figurehead_moves(int sq) { return __ldg(&d_figurehead_table[sq]); }

// ---- sliding attacks: hyperbola quintessence with a full 64-bit bit reversal (works for ranks,
// files and both diagonals; no tables, no memory traffic) ---------------------------------------
__device__ __forceinline__ u64 line_attacks(u64 occ, u64 bit, u64 mask_ex) {
    u64 f = occ & mask_ex;
    u64 r = __brevll(f);
    f -= bit;
    r -= __brevll(bit);
    f ^= __brevll(r);
    return f & mask_ex;
}
__device__ __forceinline__ u64 rank_mask(int sq) { return RANK_1 << (sq & 56); }
__device__ __forceinline__ u64 file_mask(int sq) { return FILE_A << (sq & 7); }
__device__ __forceinline__ u64 diag_mask(int sq) { // squares with file - rank constant (step 9)
    int d = (sq & 7) - (sq >> 3);
    return d >= 0 ? MAIN_DIAG >> (8 * d) : MAIN_DIAG << (-8 * d);
}
__device__ __forceinline__ u64 anti_mask(int sq) { // squares with file + rank constant (step 7)
    int s = (sq & 7) + (sq >> 3) - 7;
    return s >= 0 ? MAIN_ANTI << (8 * s) : MAIN_ANTI >> (-8 * s);
}
__device__ __forceinline__ u64 diag_attacks(int sq, u64 occ) {
    u64 bit = 1ull << sq;
    return line_attacks(occ, bit, diag_mask(sq) & ~bit) | line_attacks(occ, bit, anti_mask(sq) & ~bit);
}
__device__ __forceinline__ u64 orth_attacks(int sq, u64 occ) {
    u64 bit = 1ull << sq;
    return line_attacks(occ, bit, rank_mask(sq) & ~bit) | line_attacks(occ, bit, file_mask(sq) & ~bit);
}

// squares from which a servant of team A attacks `bit` (life.rs:759, 765: stun offsets (+-1 rank, +-1 file))
template <int A> __device__ __forceinline__ u64 servant_attackers_of(u64 bit) {
    if (A == 0) return ((bit >> 7) & ~FILE_A) | ((bit >> 9) & ~FILE_H); // Orange servants stun upwards
    return ((bit << 7) & ~FILE_H) | ((bit << 9) & ~FILE_A);             // Blue servants stun downwards
}

template <int T> __device__ __forceinline__ u64 occupied_by(const Pos &p) { // life.rs:521-540
    return p.pl[6 * T] | p.pl[6 * T + 1] | p.pl[6 * T + 2] | p.pl[6 * T + 3] | p.pl[6 * T + 4] | p.pl[6 * T + 5];
}

// Is `sq` the destination of some pseudo-legal capture-capable move of team A (the attacker)?
// This is what both in_critical_endangerment (life.rs:678-690) and is_being_leered_at_by
// (life.rs:284-291) ask once their full movegen is boiled down, given that `sq` is occupied by the
// defender: servant stuns need only the defender on sq (life.rs:808); every other piece must not
// find one of its own team on sq (life.rs:840-844, 890-893).
template <int A> __device__ __forceinline__ bool attacked_by(const Pos &p, int sq, u64 occ, u64 attacker_occ) {
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
template <int T> __device__ __forceinline__ bool in_critical_endangerment(const Pos &p) {
    u64 k = p.pl[6 * T + FIGUREHEAD];
    if (!k) return false;
    u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p);
    u64 occ = own | opp;
    while (k) {
        int sq = __ffsll((long long)k) - 1;
        k &= k - 1;
        if (attacked_by<1 - T>(p, sq, occ, opp)) return true;
    }
    return false;
}

// ---- make-move (life.rs:569-676) ----------------------------------------------------------------
// T = mover (== initiative), JOB compile-time so that every plane index is static (no local memory).
// Returns the hospitalized job (JOB_NONE if nobody was stunned).
template <int T, int JOB> __device__ __forceinline__ u32 apply_move(const Pos &p, int from, int to, Pos &c) {
    const u64 fb = 1ull << from, tb = 1ull << to;
    c = p;
    c.pl[6 * T + JOB] = (p.pl[6 * T + JOB] & ~fb) | tb; // transit: quench then alight (space.rs:169-173)
    u32 elig = meta_elig(p.meta);
    if (JOB == FIGUREHEAD) { // life.rs:577-588
        elig &= T == 0 ? ~0x3u : ~0xCu;
        int df = (to & 7) - (from & 7);
        if (df == 2 || df == -2) { // secret service, life.rs:610-629: hop the cop plane whether or not a cop is there
            int rank8 = to & 56;
            u64 start = 1ull << (rank8 + (df == 2 ? 7 : 0)), end = 1ull << (rank8 + (df == 2 ? 5 : 3));
            c.pl[6 * T + COP] = (c.pl[6 * T + COP] & ~start) | end;
        }
    }
    if (JOB == COP) { // life.rs:589-605: file of whence only
        int f = from & 7;
        if (f == 0) elig &= T == 0 ? ~0x1u : ~0x4u;
        if (f == 7) elig &= T == 0 ? ~0x2u : ~0x8u;
    }
    // life.rs:631-657: victim looked up on the pre-move board
    constexpr int E = 1 - T;
    u64 victim = tb;
    u32 pb = meta_pb(p.meta);
    if (JOB == SERVANT && !(tb & occupied_by<E>(p)) && pb != 0xFFu && loc_to_sq(pb) == to)
        victim = T == 0 ? tb >> 8 : tb << 8; // passing by: the servant one rank "behind" the passed square
    u32 hosp = JOB_NONE;
#pragma unroll
    for (int j = 5; j >= 0; j--)
        if (p.pl[6 * E + j] & victim) hosp = j; // first in dramatis_personae order wins (life.rs:550-557)
#pragma unroll
    for (int j = 0; j < 6; j++) c.pl[6 * E + j] = p.pl[6 * E + j] & ~victim; // W: planes are disjoint
    u32 new_pb = 0xFFu; // life.rs:660-669
    if (JOB == SERVANT) {
        int dr = (to >> 3) - (from >> 3);
        if (dr == 2 || dr == -2) new_pb = sq_to_loc(T == 0 ? from + 8 : from - 8);
    }
    c.meta = make_meta(E, elig, new_pb);
    return hosp;
}

// life.rs:701-726: replace the servant on `to` by `job`
template <int T> __device__ __forceinline__ void ascend(Pos &c, int to, int job) {
    const u64 tb = 1ull << to;
    c.pl[6 * T + SERVANT] &= ~tb;
    if (job == PONY) c.pl[6 * T + PONY] |= tb;
    if (job == SCHOLAR) c.pl[6 * T + SCHOLAR] |= tb;
    if (job == COP) c.pl[6 * T + COP] |= tb;
    if (job == PRINCESS) c.pl[6 * T + PRINCESS] |= tb;
}

// packed commit metadata: bit0 team | job<<1 | whence sq<<4 | whither sq<<10 | hosp job<<16 | ascension job<<20
__device__ __forceinline__ u32 pack_commit(int team, int job, int from, int to, u32 hosp, u32 asc) {
    return (u32)team | ((u32)job << 1) | ((u32)from << 4) | ((u32)to << 10) | (hosp << 16) | (asc << 20);
}

// ---- the canonical generator ---------------------------------------------------------------------
// Sink: void operator()(const Pos& child, u32 packed_commit).  NIHIL = reckless (pseudo-legal).
template <int T, int JOB, bool NIHIL, class Sink>
__device__ __forceinline__ void predict(const Pos &p, int from, int to, Sink &sink) { // life.rs:728-739
    Pos c;
    u32 hosp = apply_move<T, JOB>(p, from, to, c);
    if (!NIHIL && in_critical_endangerment<T>(c)) return; // careful_apply, life.rs:692-699
    if (JOB == SERVANT && (to >> 3) == (T == 0 ? 7 : 0)) { // subpredict, life.rs:701-726
#pragma unroll 1
        for (int job = PONY; job <= PRINCESS; job++) {
            Pos a = c;
            ascend<T>(a, to, job);
            sink(a, pack_commit(T, JOB, from, to, hosp, (u32)job));
        }
    } else {
        sink(c, pack_commit(T, JOB, from, to, hosp, JOB_NONE));
    }
}

template <int T, int JOB, bool NIHIL, bool DESC, class Sink>
__device__ __forceinline__ void predict_each(const Pos &p, int from, u64 targets, Sink &sink) {
    while (targets) {
        int to = DESC ? 63 - __clzll((long long)targets) : __ffsll((long long)targets) - 1;
        targets ^= 1ull << to;
        predict<T, JOB, NIHIL>(p, from, to, sink);
    }
}

// one slider's rays in the reference's direction order, nearest square first (life.rs:857-914)
template <int T, int JOB, bool NIHIL, bool DIAGONAL, class Sink>
__device__ __forceinline__ void slider_rays(const Pos &p, int from, u64 occ, u64 own, Sink &sink) {
    const u64 bit = 1ull << from;
    const u64 below = bit - 1, above = ~(bit | below);
    if (DIAGONAL) { // SCHOLAR_OFFSETS (-1,-1), (-1,1), (1,-1), (1,1)  (life.rs:10)
        u64 d = line_attacks(occ, bit, diag_mask(from) & ~bit) & ~own;
        u64 a = line_attacks(occ, bit, anti_mask(from) & ~bit) & ~own;
        predict_each<T, JOB, NIHIL, true>(p, from, d & below, sink);
        predict_each<T, JOB, NIHIL, true>(p, from, a & below, sink);
        predict_each<T, JOB, NIHIL, false>(p, from, a & above, sink);
        predict_each<T, JOB, NIHIL, false>(p, from, d & above, sink);
    } else { // COP_OFFSETS (-1,0), (1,0), (0,-1), (0,1)  (life.rs:11)
        u64 f = line_attacks(occ, bit, file_mask(from) & ~bit) & ~own;
        u64 r = line_attacks(occ, bit, rank_mask(from) & ~bit) & ~own;
        predict_each<T, JOB, NIHIL, true>(p, from, f & below, sink);
        predict_each<T, JOB, NIHIL, false>(p, from, f & above, sink);
        predict_each<T, JOB, NIHIL, true>(p, from, r & below, sink);
        predict_each<T, JOB, NIHIL, false>(p, from, r & above, sink);
    }
}

// life.rs:980-1050.  `leered` is is_being_leered_at_by boiled down (see attacked_by).
template <int T, bool NIHIL, class Sink> __device__ __forceinline__ void service_lookahead(const Pos &p, Sink &sink) {
    const u32 elig = meta_elig(p.meta);
    const bool west = elig & (T == 0 ? 0x1u : 0x4u), east = elig & (T == 0 ? 0x2u : 0x8u);
    if (!(west || east)) return;
    const int home = T == 0 ? 0 : 56;
    const u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p), occ = own | opp;
    const bool west_clear = west && !(occ & (0x0Eull << home)), east_clear = east && !(occ & (0x60ull << home));
    if (!(west_clear || east_clear)) return;
    if (attacked_by<1 - T>(p, home + 4, occ, opp)) return; // life.rs:1034-1043
    if (west_clear && !attacked_by<1 - T>(p, home + 1, occ, opp) && !attacked_by<1 - T>(p, home + 2, occ, opp) &&
        !attacked_by<1 - T>(p, home + 3, occ, opp))
        predict<T, FIGUREHEAD, NIHIL>(p, home + 4, home + 2, sink);
    if (east_clear && !attacked_by<1 - T>(p, home + 5, occ, opp) && !attacked_by<1 - T>(p, home + 6, occ, opp))
        predict<T, FIGUREHEAD, NIHIL>(p, home + 4, home + 6, sink);
}

// life.rs:1052-1084: servants, ponies, scholars, cops, princesses, figurehead, then secret service
template <int T, bool NIHIL, class Sink> __device__ void lookahead_as(const Pos &p, Sink &sink) {
    const u64 own = occupied_by<T>(p), opp = occupied_by<1 - T>(p), occ = own | opp, empty = ~occ;
    { // life.rs:749-830
        const u32 pb = meta_pb(p.meta);
        const u64 pb_bit = pb != 0xFFu ? 1ull << loc_to_sq(pb) : 0ull;
        for (u64 s = p.pl[6 * T + SERVANT]; s; s &= s - 1) {
            const int from = __ffsll((long long)s) - 1;
            const u64 fb = 1ull << from;
            const u64 one = T == 0 ? fb << 8 : fb >> 8;
            if (one & empty) predict<T, SERVANT, NIHIL>(p, from, T == 0 ? from + 8 : from - 8, sink);
            if ((from >> 3) == (T == 0 ? 1 : 6)) {
                const u64 two = T == 0 ? fb << 16 : fb >> 16;
                if ((two & empty) && (one & empty)) predict<T, SERVANT, NIHIL>(p, from, T == 0 ? from + 16 : from - 16, sink);
            }
            const u64 west = T == 0 ? (fb << 7) & ~FILE_H : (fb >> 9) & ~FILE_H;
            const u64 east = T == 0 ? (fb << 9) & ~FILE_A : (fb >> 7) & ~FILE_A;
            if (west && ((west & opp) || west == pb_bit)) predict<T, SERVANT, NIHIL>(p, from, T == 0 ? from + 7 : from - 9, sink);
            if (east && ((east & opp) || east == pb_bit)) predict<T, SERVANT, NIHIL>(p, from, T == 0 ? from + 9 : from - 7, sink);
        }
    }
    for (u64 s = p.pl[6 * T + PONY]; s; s &= s - 1) { // life.rs:832-855
        const int from = __ffsll((long long)s) - 1;
        predict_each<T, PONY, NIHIL, false>(p, from, pony_moves(from) & ~own, sink);
    }
    for (u64 s = p.pl[6 * T + SCHOLAR]; s; s &= s - 1)
        slider_rays<T, SCHOLAR, NIHIL, true>(p, __ffsll((long long)s) - 1, occ, own, sink);
    for (u64 s = p.pl[6 * T + COP]; s; s &= s - 1)
        slider_rays<T, COP, NIHIL, false>(p, __ffsll((long long)s) - 1, occ, own, sink);
    for (u64 s = p.pl[6 * T + PRINCESS]; s; s &= s - 1) // life.rs:955-969: all diagonals first ...
        slider_rays<T, PRINCESS, NIHIL, true>(p, __ffsll((long long)s) - 1, occ, own, sink);
    for (u64 s = p.pl[6 * T + PRINCESS]; s; s &= s - 1) // ... then all orthogonals
        slider_rays<T, PRINCESS, NIHIL, false>(p, __ffsll((long long)s) - 1, occ, own, sink);
    for (u64 s = p.pl[6 * T + FIGUREHEAD]; s; s &= s - 1) { // life.rs:973-978
        const int from = __ffsll((long long)s) - 1;
        predict_each<T, FIGUREHEAD, NIHIL, false>(p, from, figurehead_moves(from) & ~own, sink);
    }
    service_lookahead<T, NIHIL>(p, sink);
}

template <bool NIHIL, class Sink> __device__ __forceinline__ void lookahead(const Pos &p, Sink &sink) {
    if (meta_stm(p.meta) == 0) lookahead_as<0, NIHIL>(p, sink);
    else lookahead_as<1, NIHIL>(p, sink);
}

// ---- well-formedness (domain W, include/leafline_b200.h) ----------------------------------------
__device__ __forceinline__ bool well_formed(const Pos &p) {
    u64 acc = 0, overlap = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        overlap |= acc & p.pl[j];
        acc |= p.pl[j];
    }
    if (overlap) return false;
    if (__popcll(p.pl[FIGUREHEAD]) > 1 || __popcll(p.pl[6 + FIGUREHEAD]) > 1) return false;
    const u32 elig = meta_elig(p.meta);
    if ((elig & 0x3u) && (p.pl[FIGUREHEAD] & ~(1ull << 4))) return false;
    if ((elig & 0xCu) && (p.pl[6 + FIGUREHEAD] & ~(1ull << 60))) return false;
    const u32 pb = meta_pb(p.meta);
    if (pb != 0xFFu) {
        if ((pb >> 4) > 7 || (pb & 0xFu) > 7) return false;
        const int sq = loc_to_sq(pb);
        const u64 bit = 1ull << sq;
        if (acc & bit) return false;
        if (meta_stm(p.meta) == 0) { // Orange to move: a Blue servant just double-stepped over rank 5
            if ((sq >> 3) != 5 || !(p.pl[6 + SERVANT] & (bit >> 8)) || (acc & (bit << 8))) return false;
        } else {
            if ((sq >> 3) != 2 || !(p.pl[SERVANT] & (bit << 8)) || (acc & (bit >> 8))) return false;
        }
    }
    return true;
}

// ---- static evaluation (mind.rs:50-143): strict f32, one rounded multiply and one rounded add per
// reference statement, in the reference's order; __fmul_rn/__fadd_rn are never contracted to FMA ----
__device__ __forceinline__ float score(const Pos &p) {
    const float values[6] = {1.0f, 3.2f, 3.3f, 5.1f, 8.8f, 20000.0f}; // mind.rs:36-48
    float v = 0.0f;
    v = __fadd_rn(v, __fmul_rn(0.5f, meta_stm(p.meta) ? -1.0f : 1.0f)); // mind.rs:53
#pragma unroll
    for (int team = 0; team < 2; team++) { // mind.rs:55-68
        const float orient = team ? -1.0f : 1.0f;
#pragma unroll
        for (int job = 0; job < 6; job++)
            v = __fadd_rn(v, __fmul_rn((float)__popcll(p.pl[6 * team + job]), __fmul_rn(orient, values[job])));
        if (__popcll(p.pl[6 * team + SCHOLAR]) >= 2) v = __fadd_rn(v, __fmul_rn(orient, 0.5f));
    }
    const u64 center = 0x00003C3C3C3C0000ull; // mind.rs:70-81
    const int oc = __popcll((p.pl[SERVANT] | p.pl[PONY]) & center), bc = __popcll((p.pl[6 + SERVANT] | p.pl[6 + PONY]) & center);
    v = __fadd_rn(v, __fmul_rn(0.1f, (float)(oc - bc)));
    const u64 high_seventh = 0xFFull << 48, low_seventh = 0xFFull << 8; // mind.rs:83-89
    v = __fadd_rn(v, __fmul_rn(0.5f, (float)__popcll(p.pl[COP] & high_seventh)));
    v = __fadd_rn(v, -__fmul_rn(0.5f, (float)__popcll(p.pl[6 + COP] & low_seventh)));
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
            const int f = __ffs((int)files) - 1;
            files &= files - 1;
            const u64 file = FILE_A << f;
            const int o = __popcll(os & file), b = __popcll(bs & file);
            if (o > 1) v = __fadd_rn(v, -__fmul_rn(0.5f, (float)(o - 1)));
            if (b > 1) v = __fadd_rn(v, __fmul_rn(0.5f, (float)(b - 1)));
        }
    }
    // mind.rs:113-131
    v = __fadd_rn(v, __fmul_rn(1.8f, (float)__popcll(p.pl[SERVANT] & high_seventh)));
    v = __fadd_rn(v, __fmul_rn(0.6f, (float)__popcll(p.pl[SERVANT] & (0xFFull << 40))));
    v = __fadd_rn(v, -__fmul_rn(1.8f, (float)__popcll(p.pl[6 + SERVANT] & low_seventh)));
    v = __fadd_rn(v, -__fmul_rn(0.6f, (float)__popcll(p.pl[6 + SERVANT] & (0xFFull << 16))));
    const u32 elig = meta_elig(p.meta); // mind.rs:133-140
    if (elig & 0x3u) v = __fadd_rn(v, 0.1f);
    if (elig & 0xCu) v = __fadd_rn(v, -0.1f);
    return v;
}

} // namespace ll
