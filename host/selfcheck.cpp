// selfcheck.cpp -- a compiled host program over the C ABI (through host/leafline.hpp), written the way the
// reference's own tests are written: the known answers of src/life.rs and src/mind.rs, asked of the GPU.
// Exit code 0 and "host selfcheck ok" when everything matches.
#include <cstdio>
#include <cstdlib>

#include "leafline.hpp"

using namespace leafline;

#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            exit(1);                                                        \
        }                                                                   \
    } while (0)

int main() {
    // preservation runes survive a round trip without a GPU (life.rs:1604-1636)
    for (const char *runes : {"rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3", "rnbqkbnr/pp1ppppp/8/2p5/4P3/8/PPPP1PPP/RNBQKBNR w KQkq c6",
                              "8/8/4k3/8/8/8/8/R3K2R w KQ -"})
        EXPECT(WorldState::reconstruct(runes).preserve() == runes);
    EXPECT(Locale::from_algebraic("c4").pindex() == 26 && Locale::from_algebraic("a1").rank_and_file == 0);
    bool panicked = false;
    try {
        WorldState::reconstruct("8/8/8/8/8/8/8/7z w - -");
    } catch (const MoralPanic &) {
        panicked = true;
    }
    EXPECT(panicked);

    Mind mind(0); // throws MoralPanic without a CUDA device: there is no CPU fallback
    const WorldState opening = WorldState::fresh();
    // life.rs:1407-1432
    auto premonitions = mind.lookahead(opening);
    EXPECT(premonitions.size() == 20);
    std::vector<std::string> ponies;
    for (const Commit &c : premonitions)
        if (c.patch.star.job_description == JobDescription::Pony) ponies.push_back(c.patch.whence.to_algebraic() + c.patch.whither.to_algebraic());
    EXPECT((ponies == std::vector<std::string>{"b1a3", "b1c3", "g1f3", "g1h3"}));
    // life.rs:1309-1358: secret service availability 1, 2, 1, 0, 0, and never out of endangerment
    auto services = [&](const char *runes) {
        int n = 0;
        for (const Commit &c : mind.lookahead(WorldState::reconstruct(runes))) n += c.patch.concerns_secret_service();
        return n;
    };
    EXPECT(services("8/8/4k3/8/8/8/8/4K2R w K -") == 1);
    EXPECT(services("8/8/4k3/8/8/8/8/R3K2R w KQ -") == 2);
    EXPECT(services("8/8/4k3/8/8/8/8/R3KN1R w Q -") == 1);
    EXPECT(services("8/8/4k3/8/8/4b3/8/R3KN1R w Q -") == 0);
    EXPECT(services("8/8/4k3/8/b7/8/8/R3KN1R w Q - 0 1") == 0);
    EXPECT(services("8/8/4k3/8/4r3/8/8/4K2R w K -") == 0);
    // life.rs:1586-1602
    const WorldState fools = WorldState::reconstruct("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq -");
    EXPECT(mind.in_critical_endangerment(fools, Team::Orange) && !mind.in_critical_endangerment(fools, Team::Blue));
    EXPECT(mind.lookahead(fools).empty());
    // life.rs:1659-1682
    bool saw_passing_by = false;
    for (const Commit &c : mind.lookahead(WorldState::reconstruct("rnbqkbnr/ppp2ppp/4p3/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 3")))
        if (c.patch.whence == Locale::from_algebraic("e5") && c.patch.whither == Locale::from_algebraic("d6")) {
            saw_passing_by = true;
            EXPECT(c.tree.preserve() == "rnbqkbnr/ppp2ppp/3Pp3/8/8/8/PPPP1PPP/RNBQKBNR b KQkq -");
            EXPECT(c.hospitalization && *c.hospitalization == (Agent{Team::Blue, JobDescription::Servant}));
        }
    EXPECT(saw_passing_by);
    // mind.rs:654-660: the opening is worth exactly the reward for initiative
    EXPECT(mind.score(opening) - 0.5f == 0.0f);
    // perft to depth 5
    const uint64_t expected[] = {1, 20, 400, 8902, 197281, 4865609};
    for (int depth = 0; depth <= 5; depth++) EXPECT(mind.perft(opening, depth) == expected[depth]);
    std::vector<uint64_t> divide;
    EXPECT(mind.perft(WorldState::reconstruct("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"), 3, &divide) == 97814 && divide.size() == 48);
    // mind.rs:662-675: depth-3 search promotes to a pony
    auto forecasts = mind.kickoff(WorldState::reconstruct("8/q1P1k/8/8/8/8/6PP/7K w - -"), 3, std::nullopt, true);
    EXPECT(!forecasts.empty() && std::get<1>(forecasts[0]) > 0.0f);
    EXPECT(std::get<0>(forecasts[0]).tree.preserve() == "2N5/q3k3/8/8/8/8/6PP/7K b - -");
    EXPECT(std::get<2>(forecasts[0]).size() >= 1 && std::get<2>(forecasts[0])[0].abbreviated_pagan_movement_rune() == "c8");
    auto pruned = mind.alpha_beta_kickoff(WorldState::reconstruct("8/q1P1k/8/8/8/8/6PP/7K w - -"), 3, 1, std::nullopt, true);
    EXPECT(pruned.size() == forecasts.size() && std::get<1>(pruned[0]) == std::get<1>(forecasts[0]) &&
           std::get<0>(pruned[0]).tree == std::get<0>(forecasts[0]).tree);
    auto [deep, depth] = mind.iterative_deepening_kickoff(opening, 0.3, false);
    EXPECT(depth >= 3 && deep.size() == 20);
    // the variations are principal: a full line below every root move
    auto pv = mind.alpha_beta_kickoff(opening, 5, 2, std::nullopt, false);
    EXPECT(pv.size() == 20 && std::get<2>(pv[0]).size() == 5);
    {   // a group of one process driving two members (sharing the device on a one-GPU box): one tree, dealt by subtree
        Minds minds(std::vector<int>{0, 0});
        EXPECT(minds.world() == 2);
        std::vector<uint64_t> dealt;
        EXPECT(minds.perft(opening, 5, &dealt) == 4865609 && dealt.size() == 20);
        EXPECT(minds.perft(opening, 6) == 119060324 && minds.perft(opening, 6) == 119060324);
        auto whole = mind.kickoff(opening, 4, std::nullopt, false);
        auto many = minds.kickoff(opening, 4, 2, std::nullopt, false);
        EXPECT(many.size() == whole.size() && std::get<1>(many[0]) == std::get<1>(whole[0]));
        EXPECT(minds.root_scores(opening, 6, std::nullopt, false).size() == 20);
    }
    puts("host selfcheck ok");
    return 0;
}
