// hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles leafline_b200/csrc/ll_device.cuh as plain C++ so that the
// generator / legality logic the CUDA kernels are built from can be compared with the oracle on the CPU
// box, over far more positions than a GPU call affords.  Nothing in the product links or loads this.
#include "../../include/leafline_b200.h"
#include "../../leafline_b200/csrc/ll_device.cuh"
#include <cstring>

using namespace ll;

static Pos to_pos(const ll_world_t *w) {
    Pos p;
    for (int j = 0; j < 12; j++) p.pl[j] = w->plane[j];
    p.meta = make_meta(w->initiative & 1u, w->service_eligibility & 0xFu, w->passing_by);
    return p;
}
static void from_pos(const Pos &p, ll_world_t *w) {
    for (int j = 0; j < 12; j++) w->plane[j] = p.pl[j];
    w->initiative = (uint8_t)meta_stm(p.meta);
    w->service_eligibility = (uint8_t)meta_elig(p.meta);
    w->passing_by = (uint8_t)meta_pb(p.meta);
    for (int j = 0; j < 5; j++) w->pad[j] = 0;
}

struct ListSink {
    u32 moves[1024];
    u32 n = 0;
    void move(u32 job, int from, int to, u32 asc, u32 flags) {
        if (n < 1024) moves[n] = mv_pack(job, from, to, asc, flags);
        n++;
    }
};

extern "C" void hc_init(void) {
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
                    int rr = r + i, ff = f + j;
                    if ((i || j) && rr >= 0 && rr < 8 && ff >= 0 && ff < 8) b |= 1ull << (8 * rr + ff);
                }
            d_pony_table[8 * r + f] = a;
            d_figurehead_table[8 * r + f] = b;
        }
}

extern "C" int hc_count(const ll_world_t *w, int nihil) {
    const Pos p = to_pos(w);
    return (int)(nihil ? count_moves<true>(p) : count_moves<false>(p));
}

extern "C" int hc_lookahead(const ll_world_t *w, int nihil, ll_commit_t *out, int cap) {
    const Pos p = to_pos(w);
    static thread_local ListSink sink;
    sink.n = 0;
    if (nihil) lookahead<true>(p, sink);
    else lookahead<false>(p, sink);
    const u32 team = meta_stm(p.meta);
    for (u32 i = 0; i < sink.n && (int)i < cap; i++) {
        const u32 m = sink.moves[i];
        Pos c;
        const u32 hosp = make_child(p, m, c);
        ll_commit_t *o = out + i;
        o->team = (uint8_t)team;
        o->job = (uint8_t)mv_job(m);
        o->whence = (uint8_t)sq_to_loc(mv_from(m));
        o->whither = (uint8_t)sq_to_loc(mv_to(m));
        o->hospitalization = hosp == JOB_NONE ? 0xFF : (uint8_t)(((team ^ 1u) << 3) | hosp);
        o->ascension = mv_asc(m) == ASC_NONE ? 0xFF : (uint8_t)((team << 3) | mv_asc(m));
        o->pad[0] = o->pad[1] = 0;
        from_pos(c, &o->tree);
    }
    return (int)sink.n;
}

// the literal path (positions outside domain W): lookahead() / reckless_lookahead() by the reference's own legality
extern "C" int hc_lookahead_literal(const ll_world_t *w, int nihil, ll_commit_t *out, int cap) {
    const Pos p = to_pos(w);
    static thread_local ListSink sink;
    sink.n = 0;
    if (nihil) lookahead_literal<true>(p, sink);
    else lookahead_literal<false>(p, sink);
    const u32 team = meta_stm(p.meta);
    for (u32 i = 0; i < sink.n && (int)i < cap; i++) {
        const u32 m = sink.moves[i];
        Pos c;
        const u32 hosp = make_child(p, m, c);
        ll_commit_t *o = out + i;
        o->team = (uint8_t)team;
        o->job = (uint8_t)mv_job(m);
        o->whence = (uint8_t)sq_to_loc(mv_from(m));
        o->whither = (uint8_t)sq_to_loc(mv_to(m));
        o->hospitalization = hosp == JOB_NONE ? 0xFF : (uint8_t)(((team ^ 1u) << 3) | hosp);
        o->ascension = mv_asc(m) == ASC_NONE ? 0xFF : (uint8_t)((team << 3) | mv_asc(m));
        o->pad[0] = o->pad[1] = 0;
        from_pos(c, &o->tree);
    }
    return (int)sink.n;
}
extern "C" int hc_domain_class(const ll_world_t *w) { return domain_class(to_pos(w)); }

extern "C" float hc_score(const ll_world_t *w) { return score(to_pos(w)); }
extern "C" int hc_well_formed(const ll_world_t *w) { return well_formed(to_pos(w)) ? 1 : 0; }

// the reckless commits decoded from the (piece / servant-set, target mask) records, as the horizon kernel decodes them
struct RecordList {
    u64 mask[REC_MAX];
    u32 info[REC_MAX];
    u32 n_rec = 0, n_moves = 0;
    void put(u64 m, u32 i) {
        if (n_rec < REC_MAX) {
            mask[n_rec] = m;
            info[n_rec] = i | (n_moves << 16);
        }
        n_rec++;
        n_moves += (u32)popc64(m);
    }
};
extern "C" int hc_record_moves(const ll_world_t *w, int stuns, uint32_t *moves, int cap, int *n_records) {
    const Pos p = to_pos(w);
    RecordList rl;
    if (stuns) reckless_records<true>(p, rl);
    else reckless_records<false>(p, rl);
    *n_records = (int)rl.n_rec;
    if (rl.n_rec > REC_MAX) return -1;
    const u32 team = meta_stm(p.meta);
    for (u32 k = 0; k < rl.n_moves && (int)k < cap; k++) {
        u32 r = 0;
        while (r + 1 < rl.n_rec && rec_base(rl.info[r + 1]) <= k) r++;
        moves[k] = rec_move(rl.mask[r], rl.info[r], k - rec_base(rl.info[r]), team);
    }
    return (int)rl.n_moves;
}
// k_horizon values a record's quiet commits by the set of ways they change the features (quiet_ways_of_record), not one by
// one: that set must be exactly the ways quiet_way_rt gives commit by commit (which hc_feature_check holds against score()).
// Returns the number of records that differ; *checked = quiet commits seen.
extern "C" int hc_quiet_ways_check(const ll_world_t *w, int *checked) {
    const Pos p = to_pos(w);
    RecordList rl;
    reckless_records<false>(p, rl);
    if (rl.n_rec > REC_MAX) return -1;
    const int T = (int)meta_stm(p.meta);
    u64 opp = 0;
    for (int j = 0; j < 6; j++) opp |= p.pl[6 * (1 - T) + j];
    int bad = 0;
    *checked = 0;
    for (u32 r = 0; r < rl.n_rec; r++) {
        const u64 quiet = rec_quiet(rl.info[r], rl.mask[r], opp);
        u32 want = 0;
        for (u64 q = quiet; q; q &= q - 1) {
            const int to = lsb64(q);
            const u32 m = rec_move_to(rl.info[r] & 0xFFFFu, to, (u32)T);
            if (mv_asc(m) != ASC_NONE || (m & (MV_SERVICE | MV_PASSING_BY)) || ((opp >> to) & 1u)) bad++; // not quiet after all
            want |= 1u << quiet_way_rt(T, mv_job(m), mv_from(m), to);
            ++*checked;
        }
        if (quiet && quiet_ways_of_record(T, rl.info[r], quiet) != want) bad++;
        // and every commit of the record that is NOT in `quiet` is loud: stuns, ascends, serves or passes by
        for (u64 q = rl.mask[r] & ~quiet; q; q &= q - 1) {
            const int to = lsb64(q);
            const u32 m = rec_move_to(rl.info[r] & 0xFFFFu, to, (u32)T);
            if (mv_asc(m) == ASC_NONE && !(m & (MV_SERVICE | MV_PASSING_BY)) && !((opp >> to) & 1u)) bad++;
        }
    }
    return bad;
}
extern "C" int hc_generator_moves(const ll_world_t *w, int stuns, uint32_t *moves, int cap) {
    const Pos p = to_pos(w);
    static thread_local ListSink sink;
    sink.n = 0;
    lookahead<true>(p, sink);
    const u64 opp = meta_stm(p.meta) == 0 ? occupied_by<1>(p) : occupied_by<0>(p);
    int n = 0;
    for (u32 i = 0; i < sink.n; i++) {
        const u32 m = sink.moves[i];
        if (stuns && !((1ull << mv_to(m)) & opp) && !(m & MV_PASSING_BY)) continue;
        if (n < cap) moves[n] = m;
        n++;
    }
    return n;
}

// the legal commits decoded from legal_records() the way the order-free perft kernels expand them
struct ExpandList {
    u32 moves[1024];
    u32 n = 0, n_rec = 0, team;
    void put(u64 m, u32 info) {
        n_rec++;
        for (; m; m &= m - 1) {
            if (n < 1024) moves[n] = rec_move_to(info, lsb64(m), team);
            n++;
        }
    }
};
extern "C" int hc_legal_record_moves(const ll_world_t *w, uint32_t *moves, int cap, int *n_records) {
    const Pos p = to_pos(w);
    static thread_local ExpandList el;
    el.n = el.n_rec = 0;
    el.team = meta_stm(p.meta);
    legal_records(p, el);
    *n_records = (int)el.n_rec;
    for (u32 i = 0; i < el.n && (int)i < cap; i++) moves[i] = el.moves[i];
    return (int)el.n;
}
extern "C" int hc_legal_generator_moves(const ll_world_t *w, uint32_t *moves, int cap) {
    const Pos p = to_pos(w);
    static thread_local ListSink sink;
    sink.n = 0;
    lookahead<false>(p, sink);
    for (u32 i = 0; i < sink.n && (int)i < cap; i++) moves[i] = sink.moves[i];
    return (int)sink.n;
}

extern "C" uint32_t hc_canonical_key(uint32_t move) { return canonical_key(move); }

// the horizon kernel's evaluation by features: score_features(features_of(p)) must be score(p), and for every reckless
// commit score_features(child_features(...)) must be score(make_child(...)), bit for bit, with the same eligibility and
// hospitalization.  Returns the number of disagreements (0 = fine); *checked = commits compared.
struct PosPlanes {
    const Pos &p;
    u64 operator()(int j) const { return p.pl[j]; }
};
static u32 bits_of(float x) {
    u32 b;
    memcpy(&b, &x, sizeof b);
    return b;
}
extern "C" int hc_feature_check(const ll_world_t *w, int *checked) {
    const Pos p = to_pos(w);
    int bad = 0;
    const Feat f0 = features_of(p);
    if (bits_of(score_features(f0, meta_stm(p.meta), meta_elig(p.meta))) != bits_of(score(p))) bad++;
    static thread_local ListSink sink;
    sink.n = 0;
    lookahead<true>(p, sink);
    *checked = (int)sink.n;
    const PosPlanes planes{p};
    for (u32 i = 0; i < sink.n && i < 1024; i++) {
        const u32 m = sink.moves[i];
        Pos c;
        const u32 hosp = make_child(p, m, c);
        Feat f = f0;
        u32 elig = meta_elig(p.meta);
        const u32 got_hosp = meta_stm(p.meta) == 0 ? child_features<0>(f, elig, planes, m) : child_features<1>(f, elig, planes, m);
        const Feat want = features_of(c);
        bool same = got_hosp == hosp && elig == meta_elig(c.meta);
        for (int k = 0; k < 7; k++) same = same && f.w[k] == want.w[k];
        if (bits_of(score_features(f, meta_stm(c.meta), elig)) != bits_of(score(c))) same = false;
        // the horizon kernel's quiet path: nobody stunned, nothing ascended, no secret service, no passing by
        if (hosp == JOB_NONE && mv_asc(m) == ASC_NONE && !(m & (MV_SERVICE | MV_PASSING_BY))) {
            const u32 stm_child = meta_stm(p.meta) ^ 1u;
            const float material = score_material(f0, stm_child);
            const float quiet = meta_stm(p.meta) == 0
                                    ? quiet_child_score<0>(material, f0.w[3], f0.w[4], f0.w[5], f0.w[6], meta_elig(p.meta), mv_job(m), mv_from(m), mv_to(m))
                                    : quiet_child_score<1>(material, f0.w[3], f0.w[4], f0.w[5], f0.w[6], meta_elig(p.meta), mv_job(m), mv_from(m), mv_to(m));
            if (bits_of(quiet) != bits_of(score(c))) same = false;
            // ... and the same through the table of ways a quiet commit can change the features (k_horizon)
            const int T = (int)meta_stm(p.meta);
            const u32 way = quiet_way_rt(T, mv_job(m), mv_from(m), mv_to(m));
            u32 w3[4], w4[4];
            float sums[4];
            for (u32 u = 0; u < 4; u++) {
                w3[u] = f0.w[3], w4[u] = f0.w[4];
                quiet_way_features(T, (way & 4u) + u, w3[u], w4[u]);
            }
            score_ways<4>(material, w3, w4, f0.w[5], f0.w[6], sums);
            const float by_way = score_eligibility(sums[way & 3u], quiet_eligibility_rt(T, meta_elig(p.meta), mv_job(m), mv_from(m)));
            if (way >= QUIET_WAYS || bits_of(by_way) != bits_of(score(c))) same = false;
        }
        if (!same) bad++;
    }
    return bad;
}
