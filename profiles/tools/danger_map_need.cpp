// How often does the perft(6) tail need the danger map?  Walks the opening's legal tree on the CPU with the device
// code compiled as plain C++ (like tests/hostcheck) and counts, over the 4 865 609 ply-5 nodes whose commits the tail
// counts, those whose figurehead has a pseudo-escape square (only then legal_context fills the map).
//   g++ -O2 -std=c++17 -w -o /tmp/danger_map_need profiles/tools/danger_map_need.cpp && /tmp/danger_map_need
// Result (round 1): 53.8 % need the map; 81.6 % of those have ONE escape square; the need is uniform among the children
// of a ply-4 node (all 53.8 %, none 44.2 %, mixed 2.0 %).
#include "../../include/leafline_b200.h"
#include "../../leafline_b200/csrc/ll_device.cuh"
#include <cstdio>
#include <cstring>
using namespace ll;
struct ListSink { u32 moves[512]; u32 n = 0; void move(u32 job, int from, int to, u32 asc, u32 flags) { moves[n++] = mv_pack(job, from, to, asc, flags); } };
static unsigned long long total = 0, need = 0, parents = 0, parents_all = 0, parents_none = 0, one_escape = 0;
static void walk(const Pos &p, int depth) {
    ListSink s; lookahead<false>(p, s);
    if (depth == 1) { // p is a ply-4 node: its children are the tail's children
        unsigned k = 0;
        for (u32 i = 0; i < s.n; i++) {
            Pos c; make_child(p, s.moves[i], c);
            const int T = meta_stm(c.meta);
            const u64 own = T == 0 ? occupied_by<0>(c) : occupied_by<1>(c);
            const u64 kbit = c.pl[6 * T + FIGUREHEAD];
            total++;
            if (kbit) { const u64 esc = figurehead_moves(lsb64(kbit)) & ~own; if (esc) { need++; k++; if (!(esc & (esc - 1))) one_escape++; } }
        }
        parents++; if (k == s.n) parents_all++; if (k == 0) parents_none++;
        return;
    }
    for (u32 i = 0; i < s.n; i++) { Pos c; make_child(p, s.moves[i], c); walk(c, depth - 1); }
}
int main() {
    static const int pony_opt[8][2] = {{1, 2}, {-1, 2}, {1, -2}, {-1, -2}, {2, 1}, {-2, 1}, {2, -1}, {-2, -1}};
    for (int r = 0; r < 8; r++) for (int f = 0; f < 8; f++) { u64 a = 0, b = 0;
        for (int k = 0; k < 8; k++) { int rr = r + pony_opt[k][0], ff = f + pony_opt[k][1]; if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) a |= 1ull << (8 * rr + ff); }
        for (int i = -1; i <= 1; i++) for (int j = -1; j <= 1; j++) { int rr = r + i, ff = f + j; if ((i || j) && rr >= 0 && rr < 8 && ff >= 0 && ff < 8) b |= 1ull << (8 * rr + ff); }
        d_pony_table[8 * r + f] = a; d_figurehead_table[8 * r + f] = b; }
    Pos root; memset(&root, 0, sizeof root);
    root.pl[0] = 0xFF00ull; root.pl[1] = 0x42; root.pl[2] = 0x24; root.pl[3] = 0x81; root.pl[4] = 0x08; root.pl[5] = 0x10;
    for (int j = 0; j < 6; j++) root.pl[6 + j] = j == 0 ? 0xFFull << 48 : root.pl[j] << 56;
    root.meta = make_meta(0, 0xF, 0xFF);
    walk(root, 5);
    printf("children %llu need danger map %llu (%.1f%%), single escape square %llu (%.1f%% of those); ply-4 parents %llu: all children need it %llu (%.1f%%), none %llu (%.1f%%)\n",
           total, need, 100.0 * need / total, one_escape, 100.0 * one_escape / need, parents, parents_all, 100.0 * parents_all / parents, parents_none, 100.0 * parents_none / parents);
}
