// leafline.hpp -- C++ host-side mirror of the reference's interface for the leaf layer, over the C ABI of
// include/leafline_b200.h.  The reference is Rust; no Rust toolchain exists in the build image, so this header is
// what a compiled host program links against here, with the reference's names, argument meaning and error
// behaviour (a `moral_panic!` becomes a thrown leafline::MoralPanic; absence is an empty vector / std::nullopt):
//
//   reference                                                here
//   Team, JobDescription, Agent        identity.rs:8-49      Team, JobDescription, Agent
//   Locale                             space.rs:2-123        Locale
//   Patch, Commit                      life.rs:15-20, 77-83  Patch, Commit
//   WorldState::new / reconstruct / preserve  life.rs:184-507  WorldState::fresh / reconstruct / preserve
//   world.lookahead() / reckless_lookahead()  life.rs:1078-1084  Mind::lookahead / reckless_lookahead
//   world.in_critical_endangerment(team)      life.rs:678       Mind::in_critical_endangerment
//   mind::score                               mind.rs:50        Mind::score
//   mind::kickoff / iterative_deepening_kickoff / fixed_depth_sequence_kickoff  mind.rs:441-490  Mind::kickoff ...
//
// Everything that computes runs on the GPU behind the ABI; this header only converts and checks.
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../include/leafline_b200.h"

namespace leafline {

struct MoralPanic : std::runtime_error { // sorceries.rs:2-4
    using std::runtime_error::runtime_error;
};

enum class Team : uint8_t { Orange = 0, Blue = 1 };
enum class JobDescription : uint8_t { Servant = 0, Pony, Scholar, Cop, Princess, Figurehead };

struct Agent { // identity.rs:46-49
    Team team;
    JobDescription job_description;
    bool operator==(const Agent &o) const { return team == o.team && job_description == o.job_description; }
    char to_preservation_rune() const { return "PNBRQKpnbrqk"[6 * (int)team + (int)job_description]; } // identity.rs:119-147
    static Agent from_byte(uint8_t b) { return Agent{(Team)(b >> 3), (JobDescription)(b & 7)}; }
};

struct Locale { // space.rs:2-4: rank << 4 | file
    uint8_t rank_and_file;
    static Locale at(uint8_t rank, uint8_t file) { return Locale{(uint8_t)((rank << 4) | file)}; }
    static Locale from_algebraic(const std::string &n) { // space.rs:53-63
        if (n.size() != 2 || n[0] < 'a' || n[0] > 'h' || n[1] < '1' || n[1] > '8') throw MoralPanic("invalid algebraic locale " + n);
        return at((uint8_t)(n[1] - '1'), (uint8_t)(n[0] - 'a'));
    }
    uint8_t rank() const { return rank_and_file >> 4; }
    uint8_t file() const { return rank_and_file & 0xF; }
    uint32_t pindex() const { return 8u * rank() + file(); } // space.rs:65-67
    std::string to_algebraic() const { return std::string(1, (char)('a' + file())) + (char)('1' + rank()); }
    bool operator==(const Locale &o) const { return rank_and_file == o.rank_and_file; }
};

struct Patch { // life.rs:15-20
    Agent star;
    Locale whence, whither;
    bool concerns_secret_service() const { // life.rs:40-43
        return star.job_description == JobDescription::Figurehead && (whence.file() > whither.file() ? whence.file() - whither.file() : whither.file() - whence.file()) == 2;
    }
    std::string abbreviated_pagan_movement_rune() const { // life.rs:56-70
        if (concerns_secret_service()) {
            if (whither.file() == 6) return "O–O";
            if (whither.file() == 2) return "O–O–O";
            throw MoralPanic("secret service movement didn't end with figurehead on file 2 or 6");
        }
        const std::string prefix = star.job_description == JobDescription::Servant ? "" : std::string(1, "PNBRQK"[(int)star.job_description]);
        return prefix + whither.to_algebraic();
    }
};

struct WorldState { // life.rs:157-176
    ll_world_t raw{};

    static WorldState fresh() { return reconstruct("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -"); } // life.rs:184-215
    Team initiative() const { return (Team)raw.initiative; }
    std::optional<Locale> passing_by_locale() const { return raw.passing_by == LL_NONE ? std::nullopt : std::optional<Locale>(Locale{raw.passing_by}); }
    bool operator==(const WorldState &o) const {
        return !memcmp(raw.plane, o.raw.plane, sizeof raw.plane) && raw.initiative == o.raw.initiative &&
               raw.service_eligibility == o.raw.service_eligibility && raw.passing_by == o.raw.passing_by;
    }

    static WorldState reconstruct(const std::string &scan_in) { // life.rs:428-507
        std::string scan = scan_in;
        for (char &ch : scan)
            if (ch == 'X') ch = ' '; // life.rs:432
        std::vector<std::string> fields;
        size_t at = 0;
        while (at <= scan.size()) {
            size_t sp = scan.find(' ', at);
            if (sp == std::string::npos) sp = scan.size();
            fields.push_back(scan.substr(at, sp - at));
            at = sp + 1;
        }
        if (fields.size() < 4) throw MoralPanic("expected four rune fields (board, initiative, eligibility, passing-by)");
        WorldState w;
        w.raw.passing_by = LL_NONE;
        int rank = 7, file = 0;
        for (char rune : fields[0]) {
            if (rune == '/') {
                file = 0;
                rank--;
            } else if (rune >= '0' && rune <= '8') {
                file += rune - '0';
            } else {
                const char *runes = "PNBRQKpnbrqk", *hit = strchr(runes, rune);
                if (!hit || !rune) throw MoralPanic(std::string("tried to construct Agent from non-agent-preservation-rune ") + rune);
                if (rank < 0 || rank > 7 || file < 0 || file > 7) throw MoralPanic("board field runs off the world");
                w.raw.plane[hit - runes] |= 1ull << (8 * rank + file);
                file++;
            }
        }
        if (fields[1].empty() || (fields[1][0] != 'w' && fields[1][0] != 'b')) throw MoralPanic("non-initiative-preserving rune");
        w.raw.initiative = fields[1][0] == 'b';
        for (char rune : fields[2]) {
            if (rune == '-') break;
            const char *runes = "QKqk", *hit = strchr(runes, rune); // life.rs:178-181
            if (!hit || !rune) throw MoralPanic(std::string("non-eligibility rune ") + rune);
            w.raw.service_eligibility |= (uint8_t)(1u << (hit - runes));
        }
        if (fields[3] != "-") w.raw.passing_by = Locale::from_algebraic(fields[3].substr(0, 2)).rank_and_file;
        return w;
    }

    std::string preserve() const { // life.rs:293-360
        std::string book;
        for (int rank = 7; rank >= 0; rank--) {
            int run = 0;
            for (int file = 0; file < 8; file++) {
                int agent = -1;
                for (int a = 0; a < 12; a++)
                    if (raw.plane[a] >> (8 * rank + file) & 1) {
                        agent = a;
                        break;
                    }
                if (agent < 0) {
                    run++;
                } else {
                    if (run) book += (char)('0' + run), run = 0;
                    book += "PNBRQKpnbrqk"[agent];
                }
            }
            if (run) book += (char)('0' + run);
            if (rank) book += '/';
        }
        book += raw.initiative ? " b " : " w ";
        std::string eligibility;
        if (raw.service_eligibility & 2) eligibility += 'K';
        if (raw.service_eligibility & 1) eligibility += 'Q';
        if (raw.service_eligibility & 8) eligibility += 'k';
        if (raw.service_eligibility & 4) eligibility += 'q';
        book += eligibility.empty() ? "-" : eligibility;
        book += ' ';
        book += raw.passing_by == LL_NONE ? "-" : Locale{raw.passing_by}.to_algebraic();
        return book;
    }
};

struct Commit { // life.rs:77-83
    Patch patch;
    WorldState tree;
    std::optional<Agent> hospitalization, ascension;
    static Commit from_raw(const ll_commit_t &c) {
        Commit out;
        out.patch = Patch{Agent{(Team)c.team, (JobDescription)c.job}, Locale{c.whence}, Locale{c.whither}};
        out.tree.raw = c.tree;
        if (c.hospitalization != LL_NONE) out.hospitalization = Agent::from_byte(c.hospitalization);
        if (c.ascension != LL_NONE) out.ascension = Agent::from_byte(c.ascension);
        return out;
    }
};

using Variation = std::vector<Patch>;                       // mind.rs:171
using Forecast = std::tuple<Commit, float, Variation>;      // what the kickoff family returns (mind.rs:381)

// One GPU context (ll_ctx): the leaf layer and the search entry points of mind.rs.  Not thread-safe: one per host thread.
class Mind {
  public:
    explicit Mind(int device = 0) { check(ll_init(device, &ctx_)); }
    ~Mind() {
        if (owned_) ll_shutdown(ctx_);
    }
    Mind(const Mind &) = delete;
    Mind &operator=(const Mind &) = delete;

    std::vector<Commit> lookahead(const WorldState &w) { return underlookahead(w, false); }            // life.rs:1078
    std::vector<Commit> reckless_lookahead(const WorldState &w) { return underlookahead(w, true); }    // life.rs:1082
    bool in_critical_endangerment(const WorldState &w, Team team) {                                     // life.rs:678
        uint8_t t = (uint8_t)team, out = 0;
        check(ll_in_critical_endangerment_batch(ctx_, &w.raw, &t, 1, &out));
        return out != 0;
    }
    float score(const WorldState &w) { // mind.rs:50
        float s = 0;
        check(ll_score_batch(ctx_, &w.raw, 1, &s));
        return s;
    }
    std::vector<float> score(const std::vector<WorldState> &worlds) { // the batch form the driver hands off
        std::vector<ll_world_t> raw(worlds.size());
        for (size_t i = 0; i < worlds.size(); i++) raw[i] = worlds[i].raw;
        std::vector<float> out(worlds.size());
        check(ll_score_batch(ctx_, raw.data(), raw.size(), out.data()));
        return out;
    }
    uint64_t perft(const WorldState &w, int depth, std::vector<uint64_t> *divide = nullptr) {
        uint64_t leaves = 0, per_root[1024];
        uint32_t n = 0;
        check(ll_perft(ctx_, &w.raw, depth, 0, 1, &leaves, per_root, 1024, &n));
        if (divide) divide->assign(per_root, per_root + n);
        return leaves;
    }
    // mind.rs:441-448; extension = the quiescence extension (`--quiet`), nullopt = None
    std::vector<Forecast> kickoff(const WorldState &w, uint8_t depth, std::optional<uint8_t> extension, bool nihilistically) {
        std::vector<ll_forecast_t> raw(1024);
        uint32_t n = 0;
        check(ll_forecast(ctx_, &w.raw, depth, extension ? (int)*extension : -1, nihilistically, raw.data(), 1024, &n, nullptr));
        return convert(raw, n);
    }
    // the reference's search structure for depths beyond the device's full-width reach: alpha-beta with MVV-LVA +
    // history ordering on the host, frontier nodes with `device_plies` plies left valued exactly on the device
    std::vector<Forecast> alpha_beta_kickoff(const WorldState &w, uint8_t depth, uint8_t device_plies, std::optional<uint8_t> extension, bool nihilistically) {
        std::vector<ll_forecast_t> raw(1024);
        uint32_t n = 0;
        check(ll_alpha_beta_forecast(ctx_, &w.raw, depth, device_plies, extension ? (int)*extension : -1, nihilistically, raw.data(), 1024, &n, nullptr));
        return convert(raw, n);
    }
    std::pair<std::vector<Forecast>, uint8_t> iterative_deepening_kickoff(const WorldState &w, double timeout_seconds, bool nihilistically) { // mind.rs:451-469
        std::vector<ll_forecast_t> raw(1024);
        uint32_t n = 0, depth = 0;
        check(ll_iterative_deepening_forecast(ctx_, &w.raw, timeout_seconds, nihilistically, raw.data(), 1024, &n, &depth));
        return {convert(raw, n), (uint8_t)depth};
    }
    std::vector<Forecast> fixed_depth_sequence_kickoff(const WorldState &w, const std::vector<uint8_t> &depth_sequence, bool nihilistically) { // mind.rs:473-490
        std::vector<ll_forecast_t> raw(1024);
        uint32_t n = 0;
        check(ll_fixed_depth_sequence_forecast(ctx_, &w.raw, depth_sequence.data(), (uint32_t)depth_sequence.size(), nihilistically, raw.data(), 1024, &n));
        return convert(raw, n);
    }
    void set_deja_vu_bound(double gib) { check(ll_set_deja_vu_bound(ctx_, gib)); } // the kickoff functions' `déjà_vu_bound` (mind.rs:367-372)

  private:
    friend class Minds;
    ll_ctx *ctx_ = nullptr;
    bool owned_ = true;
    static void check(int rc) {
        if (rc != LL_OK) throw MoralPanic(std::string("leafline_b200: ") + ll_last_error());
    }
    std::vector<Commit> underlookahead(const WorldState &w, bool nihilistically) { // life.rs:1072-1076
        uint32_t offsets[2] = {0, 0};
        std::vector<ll_commit_t> raw(512);
        check(ll_lookahead_batch(ctx_, &w.raw, 1, nihilistically, offsets, raw.data(), raw.size()));
        std::vector<Commit> out;
        out.reserve(offsets[1]);
        for (uint32_t i = 0; i < offsets[1]; i++) out.push_back(Commit::from_raw(raw[i]));
        return out;
    }
    static std::vector<Forecast> convert(const std::vector<ll_forecast_t> &raw, uint32_t n) {
        std::vector<Forecast> out;
        for (uint32_t i = 0; i < n; i++) {
            Variation line;
            for (uint32_t j = 0; j < raw[i].variation_len; j++) {
                const ll_patch_t &p = raw[i].variation[j];
                line.push_back(Patch{Agent{(Team)p.team, (JobDescription)p.job}, Locale{p.whence}, Locale{p.whither}});
            }
            out.emplace_back(Commit::from_raw(raw[i].commit), raw[i].score, line);
        }
        return out;
    }
};

// Several GPUs sharing ONE tree (ll_mctx): the reference's fan-out inside the engine -- a thread per root move, results
// gathered over a channel, mind.rs:399-437 -- as one call.  One process drives every device (a host thread and a stream
// each, NCCL communicator and device-resident reduction owned by the group); for one process per GPU see
// ll_init_multi_rank in the header.
class Minds {
  public:
    explicit Minds(const std::vector<int> &devices) { check(ll_init_multi(devices.data(), (int)devices.size(), &group_)); }
    ~Minds() { ll_shutdown_multi(group_); }
    Minds(const Minds &) = delete;
    Minds &operator=(const Minds &) = delete;
    int world() const { return ll_multi_world(group_); }
    const char *reduce_path() const { return ll_multi_reduce_path(group_) ? "peer-memory" : "nccl"; }

    uint64_t perft(const WorldState &w, int depth, std::vector<uint64_t> *divide = nullptr) {
        uint64_t leaves = 0, per_root[1024];
        uint32_t n = 0;
        check(ll_perft_multi(group_, &w.raw, depth, &leaves, per_root, 1024, &n));
        if (divide) divide->assign(per_root, per_root + n);
        return leaves;
    }
    // root commits with their exact scores, lookahead() order (the scores of mind::kickoff before the sort, mind.rs:436)
    std::vector<std::pair<Commit, float>> root_scores(const WorldState &w, uint8_t depth, std::optional<uint8_t> extension, bool nihilistically) {
        std::vector<ll_commit_t> roots(1024);
        std::vector<float> scores(1024);
        uint32_t n = 0;
        check(ll_kickoff_multi(group_, &w.raw, depth, extension ? (int)*extension : -1, nihilistically, roots.data(), scores.data(), 1024, &n, nullptr));
        std::vector<std::pair<Commit, float>> out;
        for (uint32_t i = 0; i < n; i++) out.emplace_back(Commit::from_raw(roots[i]), scores[i]);
        return out;
    }
    // mind::kickoff by the host alpha-beta driver, the root moves searched concurrently over the GPUs (mind.rs:399-414)
    std::vector<Forecast> kickoff(const WorldState &w, uint8_t depth, uint8_t device_plies, std::optional<uint8_t> extension, bool nihilistically) {
        std::vector<ll_forecast_t> raw(1024);
        uint32_t n = 0;
        check(ll_alpha_beta_forecast_multi(group_, &w.raw, depth, device_plies, extension ? (int)*extension : -1, nihilistically, raw.data(), 1024, &n, nullptr));
        return Mind::convert(raw, n);
    }

  private:
    ll_mctx *group_ = nullptr;
    static void check(int rc) {
        if (rc != LL_OK) throw MoralPanic(std::string("leafline_b200: ") + ll_last_error());
    }
};

} // namespace leafline
