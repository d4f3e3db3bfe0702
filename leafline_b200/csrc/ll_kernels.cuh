// ll_kernels.cuh -- the CUDA kernels of the leaf layer (sm_100a).  See DESIGN.md for the roofline
// that bounds each one.  Host-side launch wrappers live in ll_api.cu.
//
// Expansion kernels (all: a block owns a tile of parents; children are rebuilt from the parent tile in shared
// memory, one CHILD per thread, so the dominant phase is balanced however uneven the parents' move counts are):
//   k_expand          canonical order (the API's lookahead / kickoff levels): one thread per parent runs the canonical
//                     generator into a shared descriptor list at its exclusive-scan offset, then the block walks the list
//                     and stores the children coalesced.
//   k_expand_records  order-free legal expansion of the perft launch chain, and its tail (children bulk-counted instead
//                     of stored): phase 1 expands per-piece legal target-mask RECORDS (small code, one inlined generator).
//   k_expand_coop     narrow perft levels: one WARP per parent, the lanes split the position by data.
//   k_horizon         the search's bottom ply: children decoded from records by select-the-t-th-bit, scored, max-reduced.
// k_score is the HBM-bound streaming scorer; the rest are scans, back-ups, repacking and test utilities.
#pragma once
#include "ll_device.cuh"
#include <cooperative_groups.h>

namespace ll {

constexpr int BLOCK = 128;
#ifndef LL_DESC_CAP
#define LL_DESC_CAP 5120
#endif
constexpr int DESC_CAP = LL_DESC_CAP; // move descriptors per window (24 KiB); a tile with more moves takes several windows

enum { MODE_EMIT = 0, MODE_PERFT = 1 };
// The perft tail reads the four lines through the figurehead from a table staged in shared memory (StagedLines, ll_device.cuh);
// -DLL_TAIL_COMPUTED_LINES builds the variant that computes them (the A/B of DESIGN.md 7b: 171.5 M -> 165.5 M warp
// instructions per perft(6) tail, perft(7) 6.65 -> 6.25 ms).
#if !defined(LL_TAIL_COMPUTED_LINES) && !defined(LL_TAIL_STAGED_LINES)
#define LL_TAIL_STAGED_LINES
#endif

// ---- block-level helpers ---------------------------------------------------------------------------
__device__ __forceinline__ u32 warp_inclusive_scan(u32 v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
// exclusive scan of one u32 per thread over a BLOCK-thread block; *total gets the block sum
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *total) {
    __shared__ u32 warp_sums[BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    u32 base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) {
        if (w < warp) base += warp_sums[w];
        sum += warp_sums[w];
    }
    __syncthreads();
    *total = sum;
    return base + inc - v;
}
// order-preserving f32 <-> i32 map (no NaNs on this path), so that a float max can be an integer atomicMax
__device__ __forceinline__ int float_key(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

// ---- layout kernels: ll_world_t (array of 104-byte structs) <-> SoA planes -------------------------
// The 104-byte struct is 13 u64 words; word 12 holds initiative | eligibility<<8 | passing_by<<16.
__global__ void __launch_bounds__(BLOCK) k_aos_to_soa(const u64 *__restrict__ aos, size_t n, Soa out, size_t out_first) {
    __shared__ u64 tile[BLOCK * 13];
    const size_t base = (size_t)blockIdx.x * BLOCK;
    const size_t in_block = min((size_t)BLOCK, n - base);
    for (size_t w = threadIdx.x; w < in_block * 13; w += BLOCK) tile[w] = aos[base * 13 + w]; // coalesced
    __syncthreads();
    if (threadIdx.x < in_block) {
        const u64 *s = tile + threadIdx.x * 13; // stride 13 words: conflict-free (13 is odd)
        const size_t i = out_first + base + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 12; j++) out.pl[(size_t)j * out.stride + i] = s[j];
        out.meta[i] = meta_of_abi_word(s[12]);
    }
}
// one position handed over as a kernel ARGUMENT (the root of ll_perft / ll_kickoff): no staging copy, no separate
// host-to-device transfer in front of the launch chain
struct WorldWords {
    u64 w[13];
};
__global__ void __launch_bounds__(32) k_store_world(const WorldWords world, Soa out, size_t at) {
    const int j = threadIdx.x;
    if (j < 12) out.pl[(size_t)j * out.stride + at] = world.w[j];
    else if (j == 12) out.meta[at] = meta_of_abi_word(world.w[12]);
}
__device__ __forceinline__ u64 meta_to_word(u32 meta) {
    return (u64)meta_stm(meta) | ((u64)meta_elig(meta) << 8) | ((u64)meta_pb(meta) << 16);
}
__global__ void __launch_bounds__(BLOCK) k_soa_to_aos(Soa in, size_t first, size_t n, u64 *__restrict__ aos) {
    __shared__ u64 tile[BLOCK * 13];
    const size_t base = (size_t)blockIdx.x * BLOCK;
    const size_t in_block = min((size_t)BLOCK, n - base);
    if (threadIdx.x < in_block) {
        u64 *s = tile + threadIdx.x * 13;
        const size_t i = first + base + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 12; j++) s[j] = in.pl[(size_t)j * in.stride + i];
        s[12] = meta_to_word(in.meta[i]);
    }
    __syncthreads();
    for (size_t w = threadIdx.x; w < in_block * 13; w += BLOCK) aos[base * 13 + w] = tile[w];
}
// children SoA + packed commit metadata -> ll_commit_t (14 u64 words: header word + the 13-word tree)
__global__ void __launch_bounds__(BLOCK) k_commits_to_aos(Soa in, const u32 *__restrict__ cmeta, size_t n, u64 *__restrict__ aos) {
    __shared__ u64 tile[BLOCK * 15]; // 15: odd stride, conflict-free
    const size_t base = (size_t)blockIdx.x * BLOCK;
    const size_t in_block = min((size_t)BLOCK, n - base);
    if (threadIdx.x < in_block) {
        u64 *s = tile + threadIdx.x * 15;
        const size_t i = base + threadIdx.x;
        const u32 m = cmeta[i];
        const u32 team = m & 1u, job = (m >> 1) & 7u, from = (m >> 4) & 63u, to = (m >> 10) & 63u;
        const u32 hosp = (m >> 16) & 0xFu, asc = (m >> 20) & 0xFu;
        const u64 hosp_byte = hosp == JOB_NONE ? 0xFFull : (u64)(((team ^ 1u) << 3) | hosp);
        const u64 asc_byte = asc == JOB_NONE ? 0xFFull : (u64)((team << 3) | asc);
        s[0] = (u64)team | ((u64)job << 8) | ((u64)sq_to_loc((int)from) << 16) | ((u64)sq_to_loc((int)to) << 24) |
               (hosp_byte << 32) | (asc_byte << 40);
#pragma unroll
        for (int j = 0; j < 12; j++) s[1 + j] = in.pl[(size_t)j * in.stride + i];
        s[13] = meta_to_word(in.meta[i]);
    }
    __syncthreads();
    for (size_t w = threadIdx.x; w < in_block * 14; w += BLOCK) aos[base * 14 + w] = tile[(w / 14) * 15 + (w % 14)];
}

// ---- validation -------------------------------------------------------------------------------------
// first_bad[0] = first position outside every domain (atomicMin), first_bad[1] != 0 if some position needs the literal path
__global__ void __launch_bounds__(BLOCK) k_validate(Soa in, size_t n, unsigned long long *first_bad) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const int cls = domain_class(soa_load(in, i));
    if (cls == DOMAIN_REJECT) atomicMin(first_bad, (unsigned long long)i);
    else if (cls == DOMAIN_LITERAL) first_bad[1] = 1ull;
}

// ---- static scoring: HBM-bound streaming kernel, 2 positions per thread, 128-bit loads ---------------
__global__ void __launch_bounds__(256) k_score(Soa in, size_t n, float *__restrict__ out) {
    const size_t pair = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t i = pair * 2;
    if (i >= n) return;
    Pos a, b;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2 *>(in.pl + (size_t)j * in.stride + i));
        a.pl[j] = v.x;
        b.pl[j] = v.y;
    }
    const uint2 m = __ldcs(reinterpret_cast<const uint2 *>(in.meta + i));
    a.meta = m.x;
    b.meta = m.y;
    const float sa = score(a), sb = score(b);
    if (i + 1 < n && ((reinterpret_cast<uintptr_t>(out) & 7u) == 0)) {
        __stcs(reinterpret_cast<float2 *>(out + i), make_float2(sa, sb));
    } else {
        out[i] = sa;
        if (i + 1 < n) out[i + 1] = sb;
    }
}

struct StunCountSink { // commits with hospitalization.is_some()
    u64 onto;
    u32 n = 0;
    LL_HDM void move(u32, int, int to, u32, u32 flags) {
        if (((1ull << to) & onto) || (flags & MV_PASSING_BY)) n++;
    }
};
LL_HD u64 opposition_squares(const Pos &p) { return meta_stm(p.meta) == 0 ? occupied_by<1>(p) : occupied_by<0>(p); }
template <bool NIHIL> LL_HD u32 count_stuns(const Pos &p) {
    StunCountSink sink{opposition_squares(p)};
    lookahead<NIHIL>(p, sink);
    return sink.n;
}
template <bool NIHIL> LL_HD u32 count_stuns_literal(const Pos &p) {
    StunCountSink sink{opposition_squares(p)};
    lookahead_literal<NIHIL>(p, sink);
    return sink.n;
}

// ---- movegen: set-wise count -> scan -> expand ---------------------------------------------------------
// literal: positions outside domain W (ll_device.cuh: DOMAIN_LITERAL) -- counted by the generator with the reference's legality
template <bool NIHIL> __global__ void __launch_bounds__(BLOCK) k_count(Soa in, size_t n, u32 *__restrict__ counts, u32 *__restrict__ block_sums, int stuns_only, int literal) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    u32 c = 0;
    if (i < n) {
        const Pos p = soa_load(in, i);
        if (literal) c = stuns_only ? count_stuns_literal<NIHIL>(p) : count_moves_literal<NIHIL>(p);
        else c = stuns_only ? count_stuns<NIHIL>(p) : count_moves<NIHIL>(p);
        counts[i] = c;
    }
    u32 total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: block_sums[0..nb) -> exclusive prefix (u64 running total written to *total)
__global__ void __launch_bounds__(1024) k_scan_block_sums(const u32 *block_sums, u64 *block_base, size_t nb, u64 *total) {
    __shared__ u64 warp_sums[32];
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t start = 0; start < nb; start += 1024) {
        const size_t i = start + threadIdx.x;
        const u64 v = i < nb ? block_sums[i] : 0;
        u64 inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        u64 base = carry_s;
        for (int w = 0; w < warp; w++) base += warp_sums[w];
        if (i < nb) block_base[i] = base + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = base + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

// offsets[i] = exclusive prefix of counts (offsets[n] = total), as u32 for the host ABI
__global__ void __launch_bounds__(BLOCK) k_offsets(const u32 *__restrict__ counts, const u64 *__restrict__ block_base, size_t n, u32 *__restrict__ offsets) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    const u32 c = i < n ? counts[i] : 0;
    u32 total;
    const u32 ex = block_exclusive_scan(c, &total);
    const u64 base = block_base[blockIdx.x];
    if (i < n) offsets[i] = (u32)(base + ex);
    if (i == n - 1) offsets[n] = (u32)(base + ex + c);
}

// phase-1 sink: descriptors of one parent into the current window of the shared list
template <bool STUNS> struct DescSink {
    u32 *buf;
    u32 idx, lo, hi, slot_bits;
    u64 onto; // STUNS: the opposition's squares -- keep only the commits that hospitalize somebody (quiescence, mind.rs:292-294)
    u32 owner = 0, world = 0; // sharded level: owner of the NEXT commit (unit_hash(parent) mod world for the first), see unit_hash
    LL_HDM void move(u32 job, int from, int to, u32 asc, u32 flags) {
        if (STUNS && !((1ull << to) & onto) && !(flags & MV_PASSING_BY)) return;
        if (idx >= lo && idx < hi) buf[idx - lo] = mv_pack(job, from, to, asc, flags) | slot_bits | (owner << 27);
        idx++;
        if (world && ++owner == world) owner = 0;
    }
};
struct ExpandArgs {
    Soa in;
    size_t n;               // parents (ignored when n_in is set)
    const u64 *n_in;        // device-resident parent count: the launch needs no host round trip
    u32 *tile_counter;      // persistent blocks pull tiles from this queue (nullptr: tile = blockIdx.x)
    const u32 *counts;      // per-parent commit counts (nullptr: counted here)
    const u64 *block_base;  // EMIT, canonical order: first child index of each tile (from the scan)
    const u32 *root_in;     // root-move id of each parent (nullable)
    int stand_pat;          // k_horizon<STUNS>: a parent may decline every stun: value = max(orientation*score, children) (mind.rs:298)
    // MODE_EMIT
    Soa out;
    u32 *cmeta;             // packed commit metadata (nullable)
    u32 *root_out;          // nullable
    int root_is_index;      // children become the roots: root id = own index
    u64 *n_out;             // EMIT without block_base: tiles claim their output range here (order-free)
    size_t out_cap;         // ... and raise *overflow instead of writing past this many children
    u32 *overflow;
    u64 *child_first;       // EMIT, order-free (nullable): where each parent's children went ...
    u32 *child_count;       // ... and how many there are (the negamax back-up of the horizon chain reads these)
    int shard_rank, shard_world; // shard_world > 1: this expansion creates the work units (see unit_hash):
    u32 *unit_owner;        //   canonical k_expand tags every child with its owner; k_expand_coop keeps only shard_rank's
    u32 *parent_out;        // EMIT, canonical (nullable): index of each child's parent in the input level
    // MODE_PERFT
    u64 *per_root, *total;
    // k_horizon
    float *values;
    u64 *scored;
    u32 *overflow_count, *overflow_list; // parents with more records than a tile keeps (HZ_REC_CAP), for k_horizon_overflow
};

#ifndef LL_EXPAND_MIN_BLOCKS
#define LL_EXPAND_MIN_BLOCKS 6
#endif
// shared memory of one expansion block (33.8 KiB: six blocks per SM)
struct ExpandShared {
    u64 tile_pl[12][BLOCK];
    u32 tile_meta[BLOCK];
    u32 descs[DESC_CAP];
    int acc[BLOCK]; // PERFT: leaves below each parent
    u64 claim;      // next tile / claimed output base
    u32 n_desc;     // k_expand_records: descriptors claimed so far by the tile's records
#ifdef LL_TAIL_STAGED_LINES
    alignas(16) u64 lines[64 * 4]; // PERFT tail: the lines through every square (StagedLines), 2 KiB, read 128 bits at a time
#endif
};
LL_HD Pos tile_load(const ExpandShared &sh, u32 slot) {
    Pos p;
#pragma unroll
    for (int j = 0; j < 12; j++) p.pl[j] = sh.tile_pl[j][slot];
    p.meta = sh.tile_meta[slot];
    return p;
}
// P = parents per tile (BLOCK, or fewer for narrow levels: shorter tiles spread over more SMs; canonical-order
// emission needs P == BLOCK because the scan works on BLOCK-sized groups)
// STUNS = quiescence level: only stunning commits are expanded
template <bool NIHIL, int MODE, int P, bool STUNS, bool LITERAL = false> __device__ __forceinline__ void expand_tiles(const ExpandArgs &a, ExpandShared &sh) {
    const int tid = threadIdx.x;
    // chained launches: once a level has outgrown its buffer, its recorded size says nothing about what was
    // written -- everything downstream stands down and the host redoes the call level by level
    if (a.n_in && a.overflow && *(volatile u32 *)a.overflow) return;
    const size_t n = a.n_in ? (size_t)*(volatile const u64 *)a.n_in : a.n;
    for (;;) {
        size_t tile = blockIdx.x;
        if (a.tile_counter) {
            if (tid == 0) sh.claim = atomicAdd(a.tile_counter, 1u);
            __syncthreads();
            tile = (size_t)sh.claim;
        }
        if (tile * P >= n) break;
        const size_t i = tile * P + tid;
        const bool mine = tid < P && i < n; // this thread owns a parent of the tile
        u32 cnt = 0;
        if (mine) { // the parent lives in the shared tile from here on, not in registers
            const Pos p = soa_load(a.in, i);
            cnt = a.counts ? a.counts[i] : STUNS ? count_stuns<NIHIL>(p) : count_moves<NIHIL>(p);
#pragma unroll
            for (int j = 0; j < 12; j++) sh.tile_pl[j][tid] = p.pl[j];
            sh.tile_meta[tid] = p.meta;
        }
        if (MODE == MODE_PERFT && tid < P) sh.acc[tid] = 0;
        u32 total;
        const u32 ex = block_exclusive_scan(cnt, &total); // also orders the tile writes before phase 2
        size_t out_base = 0;
        bool writable = true;
        if (MODE == MODE_EMIT) {
            if (a.block_base) {
                out_base = (size_t)a.block_base[tile];
            } else { // order-free: claim a contiguous range for this tile's children
                if (tid == 0) sh.claim = atomicAdd((unsigned long long *)a.n_out, (unsigned long long)total);
                __syncthreads();
                out_base = (size_t)sh.claim;
                writable = out_base + total <= a.out_cap;
                if (!writable && tid == 0) atomicExch(a.overflow, 1u);
                if (mine && a.child_first) {
                    a.child_first[i] = (u64)out_base + ex;
                    a.child_count[i] = cnt;
                }
            }
        }

        for (u32 w0 = 0; w0 < total && writable; w0 += DESC_CAP) {
            if (cnt && ex < w0 + DESC_CAP && ex + cnt > w0) { // phase 1: this parent's commits into the window
                const Pos p = tile_load(sh, tid);
                DescSink<STUNS> sink{sh.descs, ex, w0, w0 + DESC_CAP, (u32)tid << 20, STUNS ? opposition_squares(p) : 0ull};
                if (a.shard_world > 1) {
                    sink.world = (u32)a.shard_world;
                    sink.owner = unit_hash(p) % (u32)a.shard_world;
                }
                if (LITERAL) lookahead_literal<NIHIL>(p, sink);
                else lookahead<NIHIL>(p, sink);
            }
            __syncthreads();
            const u32 wn = min((u32)DESC_CAP, total - w0);
            for (u32 k0 = 0; k0 < wn; k0 += BLOCK) { // phase 2: one child per thread
                const u32 k = k0 + tid;
                const bool live = k < wn;
                u32 slot = 0;
                Pos c;
                u32 m = 0, hosp = JOB_NONE;
                if (live) {
                    m = sh.descs[k];
                    slot = mv_slot(m);
                    hosp = make_child(tile_load(sh, slot), m, c);
                }
                if (MODE == MODE_EMIT) {
                    if (live) {
                        const size_t o = out_base + w0 + k;
                        soa_store(a.out, o, c);
                        if (a.cmeta) a.cmeta[o] = pack_commit(meta_stm(c.meta) ^ 1u, m, hosp);
                        if (a.root_out) a.root_out[o] = a.root_is_index ? (u32)o : a.root_in[tile * P + slot];
                        if (a.unit_owner) a.unit_owner[o] = mv_owner(m);
                        if (a.parent_out) a.parent_out[o] = (u32)(tile * P + slot);
                    }
                } else {
                    const int v = live ? (int)count_moves<false>(c) : 0;
                    if (live && v) atomicAdd(&sh.acc[slot], v); // (see expand_tiles_records)
                }
            }
            __syncthreads();
        }

        if (MODE == MODE_PERFT) {
            // tile order keeps a root's nodes contiguous, so a warp sees very few distinct roots
            const u32 root = mine ? (a.root_in ? a.root_in[i] : 0u) : 0xFFFFFFFFu;
            const u32 below = mine ? (u32)sh.acc[tid] : 0u;
            const unsigned peers = __match_any_sync(0xffffffffu, root);
            const u32 sum = __reduce_add_sync(peers, below);
            if ((tid & 31) == __ffs((int)peers) - 1 && root != 0xFFFFFFFFu && sum) {
                if (a.per_root) atomicAdd((unsigned long long *)a.per_root + root, (unsigned long long)sum);
                atomicAdd((unsigned long long *)a.total, (unsigned long long)sum);
            }
        }
        if (!a.tile_counter) break;
        __syncthreads(); // the tile and the claim slot are reused by the next round
    }
}
template <bool NIHIL, int MODE, int P = BLOCK, bool STUNS = false, bool LITERAL = false> __global__ void __launch_bounds__(BLOCK, LL_EXPAND_MIN_BLOCKS) k_expand(const ExpandArgs a) {
    __shared__ ExpandShared sh;
    expand_tiles<NIHIL, MODE, P, STUNS, LITERAL>(a, sh);
}

// ---- the order-free legal expansion (perft launch chain and tail): phase 1 by records ------------------------
// One sink for both passes over a parent's legal records: COUNT adds the popcounts, WRITE expands every record into
// move descriptors at the parent's offset of the current window.  The record generator is inlined ONCE (a single call
// site inside a loop) -- the narrow perft levels are bound by instruction fetch, so code size is time here.
struct RecordSink {
    u32 *buf;
    u32 idx, lo, hi, slot_bits, team;
    bool write;
    LL_HDM void put(u64 mask, u32 info) {
        if (!write) {
            idx += (u32)__popcll(mask);
            return;
        }
        for (; mask; mask &= mask - 1) {
            if (idx >= lo && idx < hi) buf[idx - lo] = rec_move_to(info, __ffsll((long long)mask) - 1, team) | slot_bits;
            idx++;
        }
    }
};
// The usual case needs no counting pass at all: nothing downstream depends on the ORDER of a tile's commits, so every record
// claims its run of the descriptor list with one shared-memory atomicAdd and writes it at once -- the generator runs once per
// parent instead of twice.  Only a tile whose commits outgrow the list (> DESC_CAP: 40 a parent on average) falls back to
// count, scan and windows.
struct ClaimSink {
    u32 *buf;
    u32 *claimed;
    u32 slot_bits, team, count;
    LL_HDM void put(u64 mask, u32 info) {
#ifdef __CUDACC__
        const u32 c = (u32)__popcll(mask);
        u32 at = atomicAdd(claimed, c);
        count += c;
        if (at + c > (u32)DESC_CAP) return; // the tile has outgrown the list: the windowed passes will redo it
        for (; mask; mask &= mask - 1) buf[at++] = rec_move_to(info, __ffsll((long long)mask) - 1, team) | slot_bits;
#endif
    }
};
template <int MODE, int P> __device__ __forceinline__ void expand_tiles_records(const ExpandArgs &a, ExpandShared &sh) {
    const int tid = threadIdx.x;
#ifdef LL_TAIL_STAGED_LINES
    if (MODE == MODE_PERFT) // (the first barrier of the tile loop orders these stores before the children's loads)
        for (int i = tid; i < 64 * 4; i += BLOCK) sh.lines[i] = d_line_table[i];
#endif
    if (a.n_in && a.overflow && *(volatile u32 *)a.overflow) return; // see expand_tiles
    const size_t n = a.n_in ? (size_t)*(volatile const u64 *)a.n_in : a.n;
    for (;;) {
        size_t tile = blockIdx.x;
        if (a.tile_counter) {
            if (tid == 0) sh.claim = atomicAdd(a.tile_counter, 1u);
            __syncthreads();
            tile = (size_t)sh.claim;
        }
        if (tile * P >= n) break;
        const size_t i = tile * P + tid;
        const bool mine = tid < P && i < n;
        if (mine) {
            const Pos p = soa_load(a.in, i);
#pragma unroll
            for (int j = 0; j < 12; j++) sh.tile_pl[j][tid] = p.pl[j];
            sh.tile_meta[tid] = p.meta;
        }
        if (MODE == MODE_PERFT && tid < P) sh.acc[tid] = 0;
        if (tid == 0) sh.n_desc = 0;
        __syncthreads();
        u32 cnt = 0, ex = 0, total = 0;
        size_t out_base = 0;
        bool writable = true, direct = false;
        // pass 0 writes the tile's descriptor list by atomic claims (and counts); if the list held them all, pass 1 consumes it
        // as it stands; otherwise pass 1 + w rewrites window w in scan order and consumes it
#pragma unroll 1
        for (u32 pass = 0;; pass++) {
            const u32 w0 = pass ? (pass - 1) * DESC_CAP : 0;
            if (pass && (w0 >= total || !writable)) break;
            if (mine && pass == 0) {
                const Pos p = tile_load(sh, tid);
                ClaimSink sink{sh.descs, &sh.n_desc, (u32)tid << 20, meta_stm(p.meta), 0u};
                legal_records(p, sink);
                cnt = sink.count;
            } else if (mine && !direct && cnt && ex < w0 + DESC_CAP && ex + cnt > w0) {
                const Pos p = tile_load(sh, tid);
                RecordSink sink{sh.descs, ex, w0, w0 + DESC_CAP, (u32)tid << 20, meta_stm(p.meta), true};
                legal_records(p, sink);
            }
            if (pass == 0) {
                ex = block_exclusive_scan(cnt, &total); // (its barriers also order the claimed writes before phase 2)
                direct = total <= (u32)DESC_CAP;
                if (MODE == MODE_EMIT) { // order-free: claim a contiguous range for this tile's children
                    if (tid == 0) sh.claim = atomicAdd((unsigned long long *)a.n_out, (unsigned long long)total);
                    __syncthreads();
                    out_base = (size_t)sh.claim;
                    writable = out_base + total <= a.out_cap;
                    if (!writable && tid == 0) atomicExch(a.overflow, 1u);
                }
                continue;
            }
            __syncthreads();
            if (MODE == MODE_EMIT && a.root_is_index) {
                // the root's children become the root moves of the divide table: put the (single) parent's commits into
                // the reference's order by their canonical keys -- a rank sort over at most a few hundred entries
                constexpr u32 S = DESC_CAP / 4;
                if (total > S) { // more root moves than the sort buffer holds: the level-by-level path takes the call
                    if (tid == 0) atomicExch(a.overflow, 1u);
                    writable = false;
                    break;
                }
                for (u32 k = tid; k < total; k += BLOCK) sh.descs[S + k] = canonical_key(sh.descs[k]);
                __syncthreads();
                for (u32 k = tid; k < total; k += BLOCK) {
                    const u32 key = sh.descs[S + k];
                    u32 rank = 0;
                    for (u32 j = 0; j < total; j++) rank += sh.descs[S + j] < key;
                    sh.descs[2 * S + rank] = sh.descs[k];
                }
                __syncthreads();
                for (u32 k = tid; k < total; k += BLOCK) sh.descs[k] = sh.descs[2 * S + k];
                __syncthreads();
            }
            const u32 wn = min((u32)DESC_CAP, total - w0);
            for (u32 k0 = 0; k0 < wn; k0 += BLOCK) { // phase 2: one child per thread
                const u32 k = k0 + tid;
                const bool live = k < wn;
                u32 slot = 0;
                Pos c;
                if (live) {
                    const u32 m = sh.descs[k];
                    slot = mv_slot(m);
                    make_child(tile_load(sh, slot), m, c);
                }
                if (MODE == MODE_EMIT) {
                    if (live) {
                        const size_t o = out_base + w0 + k;
                        soa_store(a.out, o, c);
                        if (a.root_out) a.root_out[o] = a.root_is_index ? (u32)o : a.root_in[tile * P + slot];
                    }
                } else {
#ifdef LL_TAIL_STAGED_LINES
                    const int v = live ? (int)count_moves<false>(c, StagedLines{sh.lines}) : 0;
#else
                    const int v = live ? (int)count_moves<false>(c) : 0;
#endif
                    // one shared-memory atomic a child.  (Round 1 reduced per parent inside the warp first -- __match_any_sync,
                    // __reduce_add_sync, one atomic per (warp, parent): the match alone was 7 % of the tail's warp instructions and
                    // its two hottest stall sites; -DLL_TAIL_MATCH_REDUCE builds that variant: perft(6) 0.347 against 0.316 ms.)
#ifdef LL_TAIL_MATCH_REDUCE
                    const unsigned peers = __match_any_sync(0xffffffffu, live ? slot : 0xFFFFFFFFu);
                    const int r = __reduce_add_sync(peers, v);
                    if (live && (tid & 31) == __ffs((int)peers) - 1) atomicAdd(&sh.acc[slot], r);
#else
                    if (live && v) atomicAdd(&sh.acc[slot], v);
#endif
                }
            }
            __syncthreads();
        }
        if (MODE == MODE_PERFT) {
            const u32 root = mine ? (a.root_in ? a.root_in[i] : 0u) : 0xFFFFFFFFu;
            const u32 below = mine ? (u32)sh.acc[tid] : 0u;
            const unsigned peers = __match_any_sync(0xffffffffu, root);
            const u32 sum = __reduce_add_sync(peers, below);
            if ((tid & 31) == __ffs((int)peers) - 1 && root != 0xFFFFFFFFu && sum) {
                if (a.per_root) atomicAdd((unsigned long long *)a.per_root + root, (unsigned long long)sum);
                atomicAdd((unsigned long long *)a.total, (unsigned long long)sum);
            }
        }
        if (!a.tile_counter) break;
        __syncthreads();
    }
}
template <int MODE, int P = BLOCK> __global__ void __launch_bounds__(BLOCK, LL_EXPAND_MIN_BLOCKS) k_expand_records(const ExpandArgs a) {
    __shared__ ExpandShared sh;
    expand_tiles_records<MODE, P>(a, sh);
}

// ---- narrow levels: one WARP per parent (order-free legal expansion) -------------------------------------------
// A narrow perft level (1 ... ~10^4 parents) cannot fill the machine; with one thread per parent its time is that
// thread's serial walk through the whole generator.  Here a warp owns the parent and the lanes split it by DATA, so
// every lane runs the same code: lanes 0-7 one Kogge-Stone direction each of the danger map (the four down directions
// on the bit-reversed board), lanes 8-11 one line through the figurehead each (checkers and pins), then lanes 0-15 one
// non-servant piece each, lanes 16-19 one kind of servant step each for all servants at once, lanes 20-21 the
// passing-by candidates, lane 22 secret service.  The commits are those of lookahead() in another order (the root
// ply rank-sorts them by canonical key); each lane expands its own target set into the warp's descriptor list, then
// the 32 lanes build and store the children, 32 at a time.
constexpr int COOP_WARPS = BLOCK / 32;
constexpr int COOP_CAP = 256; // commits of one position (the reference's generator never exceeds ~220)
struct CoopShared {
    u32 descs[COOP_WARPS][COOP_CAP];
    u32 keys[COOP_WARPS][COOP_CAP];
};
__device__ __forceinline__ u64 warp_or64(u64 v) {
    const u32 lo = __reduce_or_sync(0xffffffffu, (u32)v), hi = __reduce_or_sync(0xffffffffu, (u32)(v >> 32));
    return ((u64)hi << 32) | lo;
}
// an up fill whose direction is data: shift s in {8, 1, 9, 7} with its wrap mask
__device__ __forceinline__ u64 ks_up_var(u64 g, u64 e, int s, u64 keep) {
    e &= keep;
    g |= e & (g << s);
    e &= e << s;
    g |= e & (g << (2 * s));
    e &= e << (2 * s);
    g |= e & (g << (4 * s));
    return (g << s) & keep;
}
template <int T> __device__ __forceinline__ void coop_expand_parent(const ExpandArgs &a, const Pos &p, size_t parent, u32 *descs, u32 *keys) {
    constexpr int E = 1 - T;
    const int lane = threadIdx.x & 31;
    const u64 own = occupied_by<T>(p), opp = occupied_by<E>(p), occ = own | opp, empty = ~occ;
    const u64 k = p.pl[6 * T + FIGUREHEAD];
    // ---- legality context, cooperatively (cf. legal_context<T>)
    Legal L;
    L.check = ~0ull;
    L.pinned = L.danger = L.kd = L.ka = L.kf = L.kr = 0;
    if (k) { // warp-uniform
        const int ksq = lsb64(k);
        const u64 e_diag = p.pl[6 * E + SCHOLAR] | p.pl[6 * E + PRINCESS], e_orth = p.pl[6 * E + COP] | p.pl[6 * E + PRINCESS];
        u64 part = 0;
        if (lane < 8) { // one direction per lane; lanes 4-7 on the bit-reversed board
            const int d = lane & 3;
            const int s = d == 0 ? 8 : d == 1 ? 1 : d == 2 ? 9 : 7;
            const u64 keep = d == 0 ? ~0ull : d == 3 ? ~FILE_H : ~FILE_A;
            u64 g = d < 2 ? e_orth : e_diag, e = ~(occ ^ k);
            if (lane >= 4) g = brev64(g), e = brev64(e);
            part = ks_up_var(g, e, s, keep);
            if (lane >= 4) part = brev64(part);
        }
        u64 danger = warp_or64(part) | servant_attacks<E>(p.pl[6 * E + SERVANT]) | pony_attacks(p.pl[6 * E + PONY]);
        if (p.pl[6 * E + FIGUREHEAD]) danger |= figurehead_moves(lsb64(p.pl[6 * E + FIGUREHEAD]));
        L.danger = danger;
        L.kd = diag_mask(ksq) & ~k;
        L.ka = anti_mask(ksq) & ~k;
        L.kf = file_mask(ksq) & ~k;
        L.kr = rank_mask(ksq) & ~k;
        u64 chk = 0, pin = 0;
        if (lane >= 8 && lane < 12) { // one line through the figurehead per lane
            const int l = lane - 8;
            const u64 line = l == 0 ? L.kd : l == 1 ? L.ka : l == 2 ? L.kf : L.kr;
            const u64 snipers = (l < 2 ? e_diag : e_orth) & line;
            if (snipers) {
                const u64 seen = line_attacks(occ, k, line);
                chk = seen & snipers;
                const u64 blockers = seen & own;
                if (blockers) {
                    u64 pinners = line_attacks(occ ^ blockers, k, line) & ~seen & snipers;
                    for (; pinners; pinners &= pinners - 1) pin |= between_on(line, k, pinners & (0 - pinners)) & blockers;
                }
            }
        }
        const u64 checkers = warp_or64(chk) | (servant_attackers_of<E>(k) & p.pl[6 * E + SERVANT]) | (pony_moves(ksq) & p.pl[6 * E + PONY]) |
                             (figurehead_moves(ksq) & p.pl[6 * E + FIGUREHEAD]);
        L.pinned = warp_or64(pin);
        if (checkers) {
            if (checkers & (checkers - 1)) {
                L.check = 0;
            } else {
                L.check = checkers | between_on(pin_line(L, checkers), k, checkers);
                if (!((L.kd | L.ka | L.kf | L.kr) & checkers)) L.check = checkers;
            }
        }
    }
    const u64 cm = L.check, free = ~L.pinned;
    // ---- this lane's target sets: `plain`, and for ascending servants `ascending` (each target counts four times)
    u64 plain = 0, ascending = 0;
    u32 info = 0;
    const u64 others = own & ~p.pl[6 * T + SERVANT];
    if (lane < 16) { // the lane-th non-servant figurine
        if (lane < popc64(others)) {
            const int from = select_bit(others, (u32)lane);
            const u64 bit = 1ull << from;
            const u32 job = (p.pl[6 * T + PONY] & bit) ? PONY : (p.pl[6 * T + SCHOLAR] & bit) ? SCHOLAR : (p.pl[6 * T + COP] & bit) ? COP
                            : (p.pl[6 * T + PRINCESS] & bit) ? PRINCESS : FIGUREHEAD;
            u64 targets = 0;
            if (job == PONY) targets = (bit & L.pinned) ? 0ull : pony_moves(from);
            else if (job == FIGUREHEAD) targets = figurehead_moves(from) & ~L.danger;
            else {
                if (job != COP) targets |= diag_attacks(from, occ);
                if (job != SCHOLAR) targets |= orth_attacks(from, occ);
                if (bit & L.pinned) targets &= pin_line(L, bit);
            }
            targets &= ~own;
            if (job != FIGUREHEAD) targets &= cm;
            plain = targets;
            info = rec_info(REC_PIECE, job, from, ASC_NONE, 0, 0);
        }
    } else if (lane < 20) { // one kind of servant step for all servants
        if (cm) {
            const u64 s = p.pl[6 * T + SERVANT];
            const u64 last = T == 0 ? RANK_1 << 56 : RANK_1;
            const u64 pushers = s & (free | L.kf);
            const u64 one = (T == 0 ? pushers << 8 : pushers >> 8) & empty;
            const int kind = lane == 16 ? REC_PUSH : lane == 17 ? REC_DOUBLE : lane == 18 ? REC_STUN_WEST : REC_STUN_EAST;
            u64 targets;
            if (kind == REC_PUSH) targets = one;
            else if (kind == REC_DOUBLE) targets = (T == 0 ? (one & (RANK_1 << 16)) << 8 : (one & (RANK_1 << 40)) >> 8) & empty;
            else if (kind == REC_STUN_WEST) targets = (T == 0 ? (s & (free | L.ka)) << 7 : (s & (free | L.kd)) >> 9) & ~FILE_H & opp;
            else targets = (T == 0 ? (s & (free | L.kd)) << 9 : (s & (free | L.ka)) >> 7) & ~FILE_A & opp;
            targets &= cm;
            plain = targets & ~last;
            ascending = targets & last;
            info = rec_info((u32)kind, SERVANT, 0, ASC_NONE, 0, 0);
        }
    } else if (lane < 22) { // passing by: the make-the-move-and-test path, one candidate per lane
        const u64 pb_bit = passing_by_bit(p.meta);
        if (pb_bit && cm) {
            u64 c = servant_attackers_of<T>(pb_bit) & p.pl[6 * T + SERVANT];
            if (lane == 21) c &= c - 1;
            if (c && passing_by_is_legal<T>(p, lsb64(c), lsb64(pb_bit))) {
                plain = pb_bit;
                info = rec_info(REC_PIECE, SERVANT, lsb64(c), ASC_NONE, 1, 0);
            }
        }
    } else if (lane == 22) { // secret service: the exact path
        if (cm == ~0ull) {
            ServiceBits service;
            service_lookahead<T, false>(p, own, opp, service);
            plain = service.to_bits;
            info = rec_info(REC_SERVICE, FIGUREHEAD, (T == 0 ? 0 : 56) + 4, ASC_NONE, 0, 0);
        }
    }
    // ---- the lanes' commits, laid end to end
    const u32 mine = (u32)popc64(plain) + 4u * (u32)popc64(ascending);
    u32 at = warp_inclusive_scan(mine);
    const u32 total = __shfl_sync(0xffffffffu, at, 31);
    at -= mine;
    if (total == 0) return;
    // a sharded level (this expansion creates the work units): of the commits in the reference's order this rank keeps
    // number first, first + world, ... -- see unit_hash
    u32 first = 0, stride = 1, kept = total;
    if (a.shard_world > 1) {
        stride = (u32)a.shard_world;
        first = ((u32)a.shard_rank + stride - unit_hash(p) % stride) % stride;
        kept = first < total ? (total - first + stride - 1) / stride : 0u;
    }
    size_t base = 0;
    if (lane == 0 && kept) base = (size_t)atomicAdd((unsigned long long *)a.n_out, (unsigned long long)kept);
    base = (size_t)__shfl_sync(0xffffffffu, (unsigned long long)base, 0);
    if (total > (u32)COOP_CAP || base + kept > a.out_cap) { // cannot happen / the level outgrew its buffer
        if (lane == 0) atomicExch(a.overflow, 1u);
        return;
    }
    if (kept == 0) return;
    for (u64 m = plain; m; m &= m - 1) descs[at++] = rec_move_to(info, lsb64(m), T);
    for (u64 m = ascending; m; m &= m - 1)
        for (u32 asc = PONY; asc <= PRINCESS; asc++) descs[at++] = rec_move_to((info & ~(7u << 12)) | (asc << 12), lsb64(m), T);
    __syncwarp();
    if (a.root_is_index || a.shard_world > 1) { // the reference's order: the root's children are the divide table's rows, a sharded level's the units
        for (u32 i = lane; i < total; i += 32) keys[i] = canonical_key(descs[i]);
        __syncwarp();
        u32 moved[COOP_CAP / 32], rank[COOP_CAP / 32];
#pragma unroll
        for (int r = 0; r < COOP_CAP / 32; r++) {
            const u32 i = lane + 32 * r;
            rank[r] = 0;
            moved[r] = 0;
            if (i < total) {
                moved[r] = descs[i];
                const u32 key = keys[i];
                for (u32 j = 0; j < total; j++) rank[r] += keys[j] < key;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < COOP_CAP / 32; r++)
            if (lane + 32 * r < total) descs[rank[r]] = moved[r];
        __syncwarp();
    }
    const u32 root = a.root_is_index ? 0u : __ldcg(a.root_in + parent); // (possibly written earlier in the same launch: k_expand_head)
    for (u32 k = lane; k < kept; k += 32) { // 32 children at a time: coalesced plane stores
        const u32 i = first + k * stride;
        Pos c;
        make_child_as<T>(p, descs[i], c);
        const size_t o = base + k;
        soa_store(a.out, o, c);
        if (a.root_out) a.root_out[o] = a.root_is_index ? i : root;
    }
    __syncwarp();
}
__global__ void __launch_bounds__(BLOCK, LL_EXPAND_MIN_BLOCKS) k_expand_coop(const ExpandArgs a) {
    __shared__ CoopShared sh;
    u32 outgrown = 0; // see expand_tiles; the two loads are issued together (one round trip, not two)
    size_t n = a.n;
    if (a.n_in) {
        outgrown = *(volatile u32 *)a.overflow;
        n = (size_t)*(volatile const u64 *)a.n_in;
    }
    if (outgrown) return;
    const int warp = threadIdx.x >> 5;
    // parents are dealt to the warps by a grid-stride loop, not from the tile queue: these levels are a dependent chain
    // per warp, and a claim is a ~1 us atomic round trip in front of it (and another one to learn that nothing is left)
    for (size_t parent = (size_t)blockIdx.x * COOP_WARPS + warp; parent < n; parent += (size_t)gridDim.x * COOP_WARPS) {
        const Pos p = soa_load(a.in, parent); // every lane holds the whole parent
        if (meta_stm(p.meta) == 0) coop_expand_parent<0>(a, p, parent, sh.descs[warp], sh.keys[warp]);
        else coop_expand_parent<1>(a, p, parent, sh.descs[warp], sh.keys[warp]);
    }
}

// ---- horizon by records (mind.rs:279-301): the best child of every frontier node, no child ever built ------------
// Phase 1 costs per PIECE, not per move: the commits of a parent come as (piece / kind of servant step, target mask)
// records (ll_device.cuh, at most REC_MAX), made by two threads a parent together with its evaluation features.
// QUIET commits (nobody stunned, nothing ascended, no secret service, no passing by: ~85 % of all) are never decoded.
//   score() reads a position through a few small integers; a quiet commit changes them in one of QUIET_WAYS = 8 ways
//   (quiet_way_rt) and the eligibility in one of four, all quiet commits of one (way, eligibility way) score the same, and
//   which ways a record's quiet commits take is a matter of masks on its target set (quiet_ways_of_record).  So a record's
//   quiet commits are folded, as it is made, into a 32-bit set and counted (SharedRecordOut); the sum of mind.rs:53-131 is
//   computed once per (parent, way) (score_ways, four ways a thread), and the best quiet child of a half parent is the best
//   of the handful of set bits: eligibility terms added, one atomicMax.  Same integers, same float sequence as
//   score(make_child(...)) for every one of those children: tests/hostcheck holds way by way against commit by commit.
// LOUD commits are valued one CHILD per thread: only records that have any take a row of the tile (their loud targets);
//   a block scan numbers them, a map gives child -> (half parent, row) (searched for only when a tile has more loud
//   children than the map holds), the victim's job comes from three folded planes of the opposition, the features are
//   updated by its loss / the ascension / the conjured cop, and the whole evaluation runs (child_features_rt +
//   score_features).  Per parent: max over its children, a shared-memory atomicMax a child.
#ifndef LL_HP
#define LL_HP 64
#endif
constexpr int HP = LL_HP; // parents per horizon tile
constexpr u32 OWNER_CAP = 1024; // LOUD children of a tile whose parent is looked up, not searched for (a tile of 64 has ~350)
// Records WITH LOUD COMMITS a parent may keep in the tile.  The record arrays are most of the block's shared memory, which
// limits the resident warps: 36 rows (every record a position can have) = five blocks a SM, 24 rows = seven (3.24 -> 2.9 ms
// when the tile still kept every record).  A parent with more loud records in either half is handed, whole, to
// k_horizon_overflow (a block per parent): none among 2^22 playout positions.
#ifndef LL_HZ_REC_CAP
#define LL_HZ_REC_CAP 24
#endif
constexpr int HZ_REC_CAP = LL_HZ_REC_CAP;
#ifndef LL_HORIZON_MIN_BLOCKS
#define LL_HORIZON_MIN_BLOCKS 7
#endif
// A parent's records are generated by TWO threads (phase 1 would otherwise leave half the block idle):
// thread slot takes the servants, the figurehead and secret service and files them in rows [0, REC_ROWS_A), thread HP + slot the
// ponies and the sliders in rows [REC_ROWS_A, HZ_REC_CAP); each half numbers its own commits from zero and is, from the scan
// on, a parent of its own ("virtual parent" v = half * HP + slot) -- only the maximum is shared.
constexpr int REC_ROWS_A = HZ_REC_CAP * 14 / 24, REC_ROWS_B = HZ_REC_CAP - REC_ROWS_A; // 14 + 10: servant steps ascend four ways each
constexpr int HV = 2 * HP;
static_assert(HV == BLOCK, "one thread per half parent");
struct HorizonShared {
    // loud child k -> map[k]: its virtual parent (low byte) and the row of the record it comes from (high byte) -- no search
    // per child.  (Its number within the record packed in as well -- 7 + 4 + 5 bits -- spares the child a load and a
    // subtraction and costs the filling thread an OR a child: 2.59 ms against 2.48, measured when the map held every child.)
    unsigned short map[OWNER_CAP];
    u32 tile_meta[HP];
    u64 opp_job[3][HP]; // bit k of (job + 1) of the opposition's figurine on each square (0 = nobody): whom a loud commit stuns
    u64 rec_mask[HZ_REC_CAP][HP]; // the records with loud commits: those commits' destinations (transposed: neighbouring parents write neighbouring words)
    u32 rec_info[HZ_REC_CAP][HP]; // kind, job, from, ascension, passing-by (16 bits)
    unsigned char rec_loud[HZ_REC_CAP][HP]; // number of the record's first LOUD commit (<= 8 stuns a piece, <= 12 loud commits a servant: < 256)
    u32 exl[HV + 1]; // exclusive prefixes of the virtual parents' loud commit counts
    unsigned short commits[HV]; // quiet + loud commits of each virtual parent
    unsigned char n_rec[HV];
    unsigned char outgrown[HV]; // a half ran out of rows: the whole parent goes to k_horizon_overflow
    u32 warp_sums[BLOCK / 32];
    int acc[HP];
    u32 feat[7][HP]; // the parents' evaluation features (ll_device.cuh: Feat), transposed like the records
    float material[HP]; // score_material(parent's features, the CHILD's initiative)
    float quiet[QUIET_WAYS][HP]; // the sum of mind.rs:53-131 for a quiet child of each way (ll_device.cuh: quiet_way_rt)
    u32 claim; // CHAIN: the tile this block pulled
};
// What child_features_rt reads of a parent's planes.  Whom a commit stuns: the opposition's six planes, folded in phase 1 into
// three (bit k of job + 1, per square) in shared memory -- three shifts instead of six loads and compares a loud child.
// A secret service also reads the mover's cop and figurehead planes: from global memory (rare).
struct TilePlanesRt {
    const HorizonShared &sh;
    u32 slot;
    const Soa &in;
    size_t at;
    __device__ __forceinline__ u64 operator()(int j) const { return ldg(in.pl + (size_t)j * in.stride + at); }
    __device__ __forceinline__ u32 opposition_job_at(int sq) const {
        const u32 code = ((u32)(sh.opp_job[0][slot] >> sq) & 1u) | (((u32)(sh.opp_job[1][slot] >> sq) & 1u) << 1) | (((u32)(sh.opp_job[2][slot] >> sq) & 1u) << 2);
        return code ? code - 1u : (u32)JOB_NONE;
    }
};
// What phase 1 keeps of a parent's records.  QUIET commits are never decoded: a record's quiet commits are folded, as it is
// made, into `quiet_set` -- bit 8 * e + w: some quiet commit changes the features way w (quiet_ways_of_record) and the
// eligibility way e (0 not at all; 1 / 2 a cop leaves its west / east corner file; 3 the figurehead moves: what
// quiet_eligibility_rt does) -- and counted.  Only a record WITH loud commits takes a row of the tile: their destinations, its info.
struct SharedRecordOut {
    HorizonShared &sh;
    int slot;
    u64 opp;
    u32 team;
    u32 row, row_end; // this half's rows
    u32 n_rec = 0, n_quiet = 0, n_loud = 0, quiet_set = 0;
    bool outgrown = false; // more loud records than the half's rows: k_horizon_overflow takes the parent
    LL_HDM void put(u64 mask, u32 info) {
        const u64 quiet = rec_quiet(info, mask, opp), loud = mask & ~quiet;
        if (quiet) {
            const u32 job = (info >> 3) & 7u, file = (info >> 6) & 7u;
            const u32 e = (info & 7u) != REC_PIECE ? 0u : job == FIGUREHEAD ? 3u : job == COP && file == 0 ? 1u : job == COP && file == 7 ? 2u : 0u;
            quiet_set |= quiet_ways_of_record((int)team, info, quiet) << (8 * e);
            n_quiet += (u32)__popcll(quiet);
        }
        if (!loud) return;
        if (row >= row_end) {
            outgrown = true;
            return;
        }
        sh.rec_mask[row][slot] = loud;
        sh.rec_info[row][slot] = info & 0xFFFFu;
        sh.rec_loud[row][slot] = (unsigned char)n_loud;
        row++;
        n_rec++;
        n_loud += (u32)__popcll(loud);
    }
};
LL_HD u32 quiet_set_eligibility(u32 team, u32 elig, u32 e) { // the eligibility after a quiet commit of eligibility way e (SharedRecordOut)
    const u32 west = team == 0 ? 0x1u : 0x4u, east = team == 0 ? 0x2u : 0x8u;
    return elig & ~(e == 1 ? west : e == 2 ? east : e == 3 ? west | east : 0u);
}
// the virtual parent of LOUD child k, the row of the record it comes from and its number among the record's loud commits:
// looked up, or -- a tile with more loud children than the map holds -- searched for: the last virtual parent whose prefix is
// <= k, then the last record of that half whose first loud commit is <= j (it has one: records without share their successor's number)
__device__ __forceinline__ void horizon_whose(const HorizonShared &sh, bool mapped, u32 k, u32 &v, u32 &row, u32 &t) {
    if (mapped) { // block-uniform
        const u32 e = sh.map[k];
        v = e & 0xFFu;
        row = e >> 8;
        t = k - (sh.exl[v] + sh.rec_loud[row][v & (HP - 1)]);
        return;
    }
    u32 hi = HV;
    v = 0;
#pragma unroll
    for (int step = 0; step < 7; step++) { // HV = 2^7
        const u32 mid = (v + hi) >> 1;
        if (sh.exl[mid] <= k) v = mid;
        else hi = mid;
    }
    const u32 slot = v & (HP - 1), j = k - sh.exl[v], row0 = v >= (u32)HP ? (u32)REC_ROWS_A : 0u;
    u32 r = 0, rhi = sh.n_rec[v];
#pragma unroll
    for (int step = 0; step < 5; step++) { // a half has < 2^5 rows; branch-free, a converged search re-reads r
        const u32 mid = (r + rhi) >> 1;
        const bool le = (u32)sh.rec_loud[row0 + mid][slot] <= j;
        r = le ? mid : r;
        rhi = le ? rhi : mid;
    }
    row = row0 + r;
    t = j - (u32)sh.rec_loud[row][slot];
}
// CHAIN: the level's size is resident on the device (a.n_in) and persistent blocks pull tiles from a.tile_counter -- the
// launch needs no host round trip (the small-batch path of ll_horizon_batch)
template <bool STUNS, bool CHAIN = false> __global__ void __launch_bounds__(BLOCK, LL_HORIZON_MIN_BLOCKS) k_horizon(const ExpandArgs a) {
    __shared__ HorizonShared sh;
    const int tid = threadIdx.x;
    if (CHAIN && *(volatile u32 *)a.overflow) return; // see expand_tiles
    const size_t n = CHAIN ? (size_t)*(volatile const u64 *)a.n_in : a.n;
#pragma unroll 1
    for (;;) {
    size_t tile = blockIdx.x;
    if (CHAIN) {
        if (tid == 0) sh.claim = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        tile = (size_t)sh.claim;
        if (tile * HP >= n) break;
    }
    const int slot_mine = tid & (HP - 1), half = tid / HP;
    const size_t i = tile * HP + slot_mine;
    const bool mine = i < n; // (both halves of the block hold the parent)
    u32 cnt_q = 0, cnt_l = 0, n_rec = 0, quiet_set = 0;
    bool outgrown = false;
    if (mine) {
        const Pos p = soa_load(a.in, i);
        const bool blue = meta_stm(p.meta) != 0;
        u64 opp = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) opp |= blue ? p.pl[j] : p.pl[6 + j];
        SharedRecordOut out{sh, slot_mine, opp, (u32)blue, half ? (u32)REC_ROWS_A : 0u, half ? (u32)HZ_REC_CAP : (u32)REC_ROWS_A};
        if (half == 0) {
            reckless_records_part<STUNS, 1>(p, out);
            u64 o[6]; // job + 1: servant 1, pony 2, scholar 3, cop 4, princess 5, figurehead 6
#pragma unroll
            for (int j = 0; j < 6; j++) o[j] = blue ? p.pl[j] : p.pl[6 + j];
            sh.opp_job[0][slot_mine] = o[SERVANT] | o[SCHOLAR] | o[PRINCESS];
            sh.opp_job[1][slot_mine] = o[PONY] | o[SCHOLAR] | o[FIGUREHEAD];
            sh.opp_job[2][slot_mine] = o[COP] | o[PRINCESS] | o[FIGUREHEAD];
            sh.tile_meta[slot_mine] = p.meta;
            sh.acc[slot_mine] = float_key(-INFINITY);
        } else {
            reckless_records_part<STUNS, 2>(p, out);
            const Feat f = features_of(p);
#pragma unroll
            for (int w = 0; w < 7; w++) sh.feat[w][slot_mine] = f.w[w];
            sh.material[slot_mine] = score_material(f, meta_stm(p.meta) ^ 1u);
        }
        outgrown = out.outgrown;
        cnt_q = out.n_quiet;
        cnt_l = out.n_loud;
        n_rec = out.n_rec;
        quiet_set = out.quiet_set;
    }
    // (a half that ran out of rows keeps the records it has: their commits are valued like any others, for nothing -- the
    // parent's value is the overflow kernel's, see the end of the tile -- which spares the tile a barrier to tell the other half)
    sh.outgrown[tid] = outgrown;
    sh.n_rec[tid] = (unsigned char)n_rec;
    sh.commits[tid] = (unsigned short)(cnt_q + cnt_l);
    u32 total_l = 0;
    u32 ex_l = warp_inclusive_scan(cnt_l);
    if ((tid & 31) == 31) sh.warp_sums[tid >> 5] = ex_l;
    __syncthreads(); // phase 1 is over: the records, the features and the warps' loud counts are there
    ex_l -= cnt_l;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) {
        if (w < (tid >> 5)) ex_l += sh.warp_sums[w];
        total_l += sh.warp_sums[w];
    }
    sh.exl[tid] = ex_l;
    if (tid == 0) sh.exl[HV] = total_l;
    const bool mapped = total_l <= OWNER_CAP; // block-uniform
    if (mapped) { // every loud child's (virtual parent, record row), written by the half that owns it: record by record
        const u32 row0 = half ? (u32)REC_ROWS_A : 0u;
        u32 c = 0;
        for (u32 r = 0; r < n_rec; r++) {
            const u32 row = row0 + r;
            const u32 end = r + 1 < n_rec ? sh.rec_loud[row + 1][slot_mine] : cnt_l;
            const unsigned short entry = (unsigned short)(tid | row << 8);
            for (; c < end; c++) sh.map[ex_l + c] = entry;
        }
    }
    { // what a quiet child of each way scores short of the eligibility terms: four ways a thread, both halves of a parent at it
        const u32 team = meta_stm(sh.tile_meta[slot_mine]);
        u32 w3[4], w4[4];
        float sums[4];
#pragma unroll
        for (u32 u = 0; u < 4; u++) {
            w3[u] = sh.feat[3][slot_mine], w4[u] = sh.feat[4][slot_mine];
            quiet_way_features((int)team, 4u * half + u, w3[u], w4[u]);
        }
        score_ways<4>(sh.material[slot_mine], w3, w4, sh.feat[5][slot_mine], sh.feat[6][slot_mine], sums);
#pragma unroll
        for (u32 u = 0; u < 4; u++) sh.quiet[4u * half + u][slot_mine] = sums[u];
    }
    __syncthreads();
    // ---- quiet commits: all quiet commits of a parent that change the features and the eligibility the same way score the
    // same, and phase 1 noted which ways occur (SharedRecordOut::quiet_set) -- the best quiet child is the best of those
    // (three or four a half, typically); no quiet child is decoded.
    if (quiet_set) {
        const u32 meta = sh.tile_meta[slot_mine], team = meta_stm(meta);
        int best = float_key(-INFINITY);
        for (u32 q = quiet_set; q; q &= q - 1) {
            const u32 b = (u32)__ffs((int)q) - 1u;
            const float s = score_eligibility(sh.quiet[b & 7u][slot_mine], quiet_set_eligibility(team, meta_elig(meta), b >> 3));
            best = max(best, float_key(team ? -s : s)); // -(orientation(child) * score(child)), the child's initiative being the other team's (mind.rs:283, 331-336)
        }
        atomicMax(&sh.acc[slot_mine], best);
    }
    auto loud_child = [&](u32 k, u32 &slot) -> int {
        u32 v, row, t;
        horizon_whose(sh, mapped, k, v, row, t);
        slot = v & (HP - 1);
        const u32 info = sh.rec_info[row][slot];
        const u64 mask = sh.rec_mask[row][slot];
        const u32 meta = sh.tile_meta[slot];
        const u32 team = meta_stm(meta);
        const int to = select_bit(mask, t);
        const u32 m = rec_move_to(info & 0xFFFFu, to, team);
        Feat f;
#pragma unroll
        for (int w = 0; w < 7; w++) f.w[w] = sh.feat[w][slot];
        u32 elig = meta_elig(meta);
        child_features_rt((int)team, f, elig, TilePlanesRt{sh, slot, a.in, tile * HP + slot}, m);
        const float s = score_features(f, team ^ 1u, elig);
        return float_key(team ? -s : s);
    };
    auto fold = [&](bool live, u32 slot, int v) { // per parent: a shared-memory atomicMax a child (see the perft tail: the warp-level
#ifdef LL_HZ_MATCH_FOLD                           // segmented reduction in front of it cost more than the atomics it saved)
        const unsigned peers = __match_any_sync(0xffffffffu, live ? slot : 0xFFFFFFFFu);
        const int best = __reduce_max_sync(peers, v);
        if (live && (tid & 31) == __ffs((int)peers) - 1) atomicMax(&sh.acc[slot], best);
#else
        if (live) atomicMax(&sh.acc[slot], v);
#endif
    };
#ifndef LL_HZ_ILP
#define LL_HZ_ILP 1
#endif
    // ---- loud commits
    for (u32 k0 = 0; k0 < total_l; k0 += LL_HZ_ILP * BLOCK) {
        u32 slot[LL_HZ_ILP];
        int v[LL_HZ_ILP];
        bool live[LL_HZ_ILP];
#pragma unroll
        for (int u = 0; u < LL_HZ_ILP; u++) {
            const u32 k = k0 + u * BLOCK + tid;
            live[u] = k < total_l;
            slot[u] = 0;
            v[u] = loud_child(live[u] ? k : 0u, slot[u]);
            if (!live[u]) v[u] = float_key(-INFINITY);
        }
#pragma unroll
        for (int u = 0; u < LL_HZ_ILP; u++) fold(live[u], slot[u], v[u]);
    }
    __syncthreads();
    u32 scored = 0;
    if (mine && half == 0 && (sh.outgrown[slot_mine] | sh.outgrown[HP + slot_mine])) {
        a.overflow_list[atomicAdd(a.overflow_count, 1u)] = (u32)i; // the overflow kernel values it (and counts its leaves)
    } else if (mine && half == 0) {
        const u32 other = (u32)HP + (u32)slot_mine;
        const u32 cnt = cnt_q + cnt_l + sh.commits[other];
        float v = cnt ? key_float(sh.acc[slot_mine]) : -INFINITY;
        if (cnt == 0 || (STUNS && a.stand_pat)) { // nothing to play: the node is itself a leaf (mind.rs:281-283, 295-299)
            Feat f;
#pragma unroll
            for (int w = 0; w < 7; w++) f.w[w] = sh.feat[w][slot_mine];
            const u32 meta = sh.tile_meta[slot_mine];
            const float s = score_features(f, meta_stm(meta), meta_elig(meta));
            v = fmaxf(v, meta_stm(meta) ? -s : s);
        }
        a.values[i] = v;
        scored = cnt ? cnt : 1u;
    }
    scored = __reduce_add_sync(0xffffffffu, scored);
    if ((tid & 31) == 0 && scored && a.scored) atomicAdd((unsigned long long *)a.scored, (unsigned long long)scored);
    if (!CHAIN) break;
    __syncthreads(); // the tile and the claim slot are reused by the next round
    }
}

// the parents k_horizon would not keep in its tile (more than HZ_REC_CAP records: a dozen pieces AND servants stepping every
// way): one BLOCK per parent -- a single thread would take ~0.2 ms over the generator and fifty evaluations, and the launch
// waits for it.  Thread 0 lists the parent's records (all REC_MAX of them fit here), then the block makes and scores one child
// per thread: the definition itself (mind.rs:279-301), score(make_child(...)).
struct OverflowRecords {
    u64 mask[REC_MAX];
    u32 info[REC_MAX]; // rec_info with the record's first commit number
    u32 n_rec, n_moves;
    LL_HDM void put(u64 m, u32 i) {
        if (n_rec < (u32)REC_MAX) {
            mask[n_rec] = m;
            info[n_rec] = i | (n_moves << 16);
            n_rec++;
        }
        n_moves += (u32)popc64(m);
    }
};
template <bool STUNS> __global__ void __launch_bounds__(BLOCK) k_horizon_overflow(const ExpandArgs a) {
    __shared__ OverflowRecords rec;
    __shared__ int best_s;
    const u32 n = *(volatile u32 *)a.overflow_count;
    for (u32 k = blockIdx.x; k < n; k += gridDim.x) {
        const size_t i = a.overflow_list[k];
        const Pos p = soa_load(a.in, i); // every thread holds the parent
        if (threadIdx.x == 0) {
            rec.n_rec = rec.n_moves = 0;
            reckless_records<STUNS>(p, rec);
            best_s = float_key(-INFINITY);
        }
        __syncthreads();
        const u32 total = rec.n_moves, team = meta_stm(p.meta);
        int best = float_key(-INFINITY);
        for (u32 j = threadIdx.x; j < total; j += BLOCK) {
            u32 r = 0;
            while (r + 1 < rec.n_rec && rec_base(rec.info[r + 1]) <= j) r++;
            const u32 m = rec_move(rec.mask[r], rec.info[r], j - rec_base(rec.info[r]), team);
            Pos c;
            make_child(p, m, c);
            const float s = score(c);
            const int key = float_key(team ? -s : s); // -(orientation(child) * score(child)), the child's initiative being the other team's
            best = key > best ? key : best;
        }
        best = __reduce_max_sync(0xffffffffu, best);
        if ((threadIdx.x & 31) == 0) atomicMax(&best_s, best);
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = total ? key_float(best_s) : -INFINITY;
            if (total == 0 || (STUNS && a.stand_pat)) {
                const float s = score(p);
                v = fmaxf(v, team ? -s : s);
            }
            a.values[i] = v;
            if (a.scored) atomicAdd((unsigned long long *)a.scored, (unsigned long long)(total ? total : 1u));
        }
        __syncthreads();
    }
}

// ---- EXPERIMENT, off by default (LEAFLINE_B200_HEAD_CLUSTER=1 turns it on; measured slower) ---------------------
// the head of the perft chain: plies 0, 1 and 2 (1, ~30, ~900 parents) in ONE launch of one thread-block cluster
// Each of those plies is a dependent chain of ~2 000 warp instructions per parent on a machine that is otherwise idle; as
// separate launches every ply pays a launch gap and starts on cold instruction caches (~13 us a ply, of which ~4 us are
// work).  Here sixteen CTAs of sixteen warps (256 warps: one parent each, ply 2 in two or three rounds) stay resident across
// the three plies, the levels' sizes are read from the device between them and a cluster barrier (which also writes the L1s
// back) stands where the kernel boundary was.  The levels are materialised exactly as by k_expand_coop -- order-free
// placement, root ids travelling with the nodes, the shard rule of a dealt level -- so everything downstream is unchanged.
constexpr int HEAD_WARPS = 16, HEAD_CTAS = 16, HEAD_MAX_STEPS = 3;
struct HeadArgs {
    ExpandArgs step[HEAD_MAX_STEPS];
    int n_steps;
};
LL_HD Pos soa_load_cg(const Soa &s, size_t i) { // written earlier in this launch by another SM: not through the read-only path
    Pos p;
#pragma unroll
    for (int j = 0; j < 12; j++) p.pl[j] = __ldcg(s.pl + (size_t)j * s.stride + i);
    p.meta = __ldcg(s.meta + i);
    return p;
}
__global__ void __launch_bounds__(HEAD_WARPS * 32, 1) k_expand_head(const HeadArgs h) {
    __shared__ u32 descs[HEAD_WARPS][COOP_CAP];
    __shared__ u32 keys[HEAD_WARPS][COOP_CAP];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int warp = threadIdx.x >> 5;
    const size_t first = (size_t)cluster.block_rank() * HEAD_WARPS + warp, stride = (size_t)cluster.num_blocks() * HEAD_WARPS;
    for (int s = 0; s < h.n_steps; s++) {
        const ExpandArgs &a = h.step[s];
        u32 outgrown = 0;
        size_t n = a.n;
        if (a.n_in) {
            outgrown = *(volatile u32 *)a.overflow;
            n = (size_t)*(volatile const u64 *)a.n_in;
        }
        if (!outgrown)
            for (size_t parent = first; parent < n; parent += stride) {
                const Pos p = soa_load_cg(a.in, parent);
                if (meta_stm(p.meta) == 0) coop_expand_parent<0>(a, p, parent, descs[warp], keys[warp]);
                else coop_expand_parent<1>(a, p, parent, descs[warp], keys[warp]);
            }
        if (s + 1 < h.n_steps) {
            __threadfence();
            cluster.sync();
        }
    }
}

// keep the units this rank owns: the children whose owner tag (written by the canonical expansion of the shard level,
// see unit_hash) equals `rank`.  Order-free compaction (perft and the search's fold need no order: ids travel with the nodes).
struct ShardArgs {
    Soa in;
    const u32 *root_in;   // per-node id carried along (root move / parent index), nullable
    const u32 *owner;     // per-node owner tag
    size_t n;
    Soa out;
    u32 *root_out;
    u64 *n_out;           // device counter, zeroed by the caller
    int rank;
};
__global__ void __launch_bounds__(BLOCK) k_shard_filter(const ShardArgs a) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool keep = i < a.n && a.owner[i] == (u32)a.rank;
    const unsigned votes = __ballot_sync(0xffffffffu, keep);
    if (!votes) return;
    const int lane = threadIdx.x & 31, leader = __ffs((int)votes) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd((unsigned long long *)a.n_out, (unsigned long long)__popc(votes));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!keep) return;
    const size_t o = (size_t)base + (size_t)__popc(votes & ((1u << lane) - 1u));
    soa_store(a.out, o, soa_load(a.in, i));
    if (a.root_out) a.root_out[o] = a.root_in ? a.root_in[i] : 0u;
}
// index-based dealing of a level whose order is the same on every rank (the search's canonical levels): unit u -> rank u % world
__global__ void __launch_bounds__(BLOCK) k_deal_units(Soa in, const u32 *__restrict__ tag_in, size_t n, int rank, int world, Soa out, u32 *__restrict__ tag_out) {
    const size_t kept = n > (size_t)rank ? (n - rank + world - 1) / world : 0;
    for (size_t o = (size_t)blockIdx.x * BLOCK + threadIdx.x; o < kept; o += (size_t)gridDim.x * BLOCK) {
        const size_t i = (size_t)rank + o * (size_t)world;
        soa_store(out, o, soa_load(in, i));
        tag_out[o] = tag_in ? tag_in[i] : (u32)i;
    }
}

// ---- the perft launch chain: one step per ply (+ the shard filter), every size resident on the device --------
constexpr int CHAIN_MAX_STEPS = 12;
enum { STEP_EMIT_SMALL = 0, STEP_EMIT = 1, STEP_TAIL = 2, STEP_EMIT_NARROW = 3 };

// ll_perft_device: gather the chain's results into the caller's device buffer (see the header for the layout)
__global__ void __launch_bounds__(1024) k_pack_perft_result(const u64 *leaves, const u64 *n_roots, const u32 *overflow, const u64 *per_root, u64 *out) {
    const int t = threadIdx.x;
    out[8 + t] = per_root[t];
    if (t < 8) out[t] = t == 0 ? *leaves : t == 1 ? *n_roots : t == 2 ? (u64)*overflow : 0ull;
}

// ---- perft tail on an already materialised frontier: count the legal moves of every node ---------------
__global__ void __launch_bounds__(BLOCK) k_count_leaves(Soa in, size_t n, const u32 *__restrict__ root_in, u64 *__restrict__ per_root, u64 *__restrict__ total, int literal) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    u32 c = 0, root = 0xFFFFFFFFu;
    if (i < n) {
        c = literal ? count_moves_literal<false>(soa_load(in, i)) : count_moves<false>(soa_load(in, i));
        root = root_in ? root_in[i] : 0u;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    const u32 sum = __reduce_add_sync(peers, c);
    if ((threadIdx.x & 31) == __ffs((int)peers) - 1 && root != 0xFFFFFFFFu && sum) {
        if (per_root) atomicAdd(per_root + root, (u64)sum);
        atomicAdd(total, (u64)sum);
    }
}

// ---- horizon (mind.rs:279-287) -------------------------------------------------------------------------
// plies == 0: value = orientation * score
__global__ void __launch_bounds__(BLOCK) k_leaf_values(Soa in, size_t n, float *__restrict__ values) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const Pos p = soa_load(in, i);
    const float s = score(p);
    values[i] = meta_stm(p.meta) ? -s : s;
}
// negamax back-up of one level: parent = max over its children of -child; childless parent -> leaf rule
__global__ void __launch_bounds__(BLOCK) k_backup(Soa parents, size_t n, const u32 *__restrict__ counts, const u64 *__restrict__ block_base,
                                                  const float *__restrict__ child_values, float *__restrict__ values, int stand_pat) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    const u32 c = i < n ? counts[i] : 0;
    u32 total;
    const u32 ex = block_exclusive_scan(c, &total);
    if (i >= n) return;
    float v = -INFINITY;
    if (c == 0 || stand_pat) {
        const Pos p = soa_load(parents, i);
        const float s = score(p);
        v = meta_stm(p.meta) ? -s : s;
    }
    if (c) {
        const size_t first = (size_t)(block_base[blockIdx.x] + ex);
        for (u32 k = 0; k < c; k++) v = fmaxf(v, -child_values[first + k]);
    }
    values[i] = v;
}

// the same back-up where the children were placed order-free (the horizon chain): every parent knows its own range,
// the level's size stays on the device
__global__ void __launch_bounds__(BLOCK) k_backup_ranges(Soa parents, size_t n_host, const u64 *__restrict__ n_in, const u32 *__restrict__ overflow,
                                                         const u64 *__restrict__ child_first, const u32 *__restrict__ child_count,
                                                         const float *__restrict__ child_values, float *__restrict__ values) {
    if (*(volatile const u32 *)overflow) return;
    const size_t n = n_in ? (size_t)*(volatile const u64 *)n_in : n_host;
    for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLOCK) {
        const u32 c = child_count[i];
        float v = -INFINITY;
        if (c == 0) {
            const Pos p = soa_load(parents, i);
            const float s = score(p);
            v = meta_stm(p.meta) ? -s : s;
        } else {
            const size_t first = (size_t)child_first[i];
            for (u32 k = 0; k < c; k++) v = fmaxf(v, -child_values[first + k]);
        }
        values[i] = v;
    }
}

// ---- principal variation (the `Variation` memory of mind.rs:208-224) -------------------------------------
// After the back-up every materialised level holds its nodes' values; the variation below a root child is
// re-traced by following, at each node, the FIRST child (canonical order) whose negated value equals the
// node's value -- a quiescence node stands pat on ties (mind.rs:298, 346: `value > optimum` is strict).
// The reference breaks ties in its MVV-LVA / history order (unstable sort), so only the values along the
// line, not the line itself, are comparable with it.
constexpr int PV_MAX = 16;       // == LL_MAX_VARIATION
constexpr int PV_LEVELS = 12;
struct PvLevel {
    Soa soa;
    const u32 *counts;
    const u64 *block_base;
    const float *values;
    const u32 *cmeta;
    int quiet;
};
struct PvArgs {
    PvLevel lev[PV_LEVELS];
    int n_levels;   // materialised levels, the root children's level first
    size_t n;       // root children on this rank
    u32 *out;       // [n][PV_MAX + 1]: length, then packed commits (pack_commit)
};
struct BestChildSink {
    const Pos &par;
    float target;
    u64 only_onto;
    u32 found;
    LL_HDM void move(u32 job, int from, int to, u32 asc, u32 flags) {
        if (found != 0xFFFFFFFFu) return;
        if (!((1ull << to) & only_onto) && !(flags & MV_PASSING_BY)) return;
        const u32 m = mv_pack(job, from, to, asc, flags);
        Pos c;
        const u32 hosp = make_child(par, m, c);
        const float s = score(c);
        if ((meta_stm(c.meta) ? s : -s) == target) found = pack_commit(meta_stm(par.meta), m, hosp);
    }
};
__global__ void __launch_bounds__(BLOCK) k_variation(const PvArgs a) {
    const size_t r = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (r >= a.n) return;
    u32 *out = a.out + r * (PV_MAX + 1);
    u32 len = 0;
    size_t node = r;
    for (int L = 0; L < a.n_levels && len < PV_MAX; L++) {
        const PvLevel &lv = a.lev[L];
        const Pos p = soa_load(lv.soa, node);
        const float v = lv.values[node];
        if (lv.quiet) { // stand pat wins ties
            const float s = score(p);
            if ((meta_stm(p.meta) ? -s : s) == v) break;
        }
        if (L == a.n_levels - 1) { // the children of the bottom level were scored in registers, never stored
            BestChildSink sink{p, v, lv.quiet ? opposition_squares(p) : ~0ull, 0xFFFFFFFFu};
            lookahead<true>(p, sink);
            if (sink.found != 0xFFFFFFFFu) out[1 + len++] = sink.found;
            break;
        }
        const u32 c = lv.counts[node];
        if (c == 0) break;
        const size_t tile0 = (node / BLOCK) * BLOCK;
        size_t first = (size_t)lv.block_base[node / BLOCK];
        for (size_t j = tile0; j < node; j++) first += lv.counts[j];
        const PvLevel &nx = a.lev[L + 1];
        u32 k = 0;
        while (k + 1 < c && -nx.values[first + k] != v) k++;
        out[1 + len++] = nx.cmeta[first + k];
        node = first + k;
    }
    out[0] = len;
}

// ---- full-size self-consistency check (tests): the three ways this library knows a position's legal moves ------
// (1) the set-wise counter, (2) the canonical generator with legality by masks, (3) the reference's own definition:
// every reckless commit made and kept iff the mover's figurehead is not endangered afterwards (life.rs:692-699).
// out[0] += positions where they disagree, out[1] = first such index, out[2] += legal commits, out[3] += reckless commits.
struct ExactLegalSink {
    const Pos &par;
    u32 n = 0;
    LL_HDM void move(u32 job, int from, int to, u32 asc, u32 flags) {
        Pos c;
        make_child(par, mv_pack(job, from, to, asc, flags), c);
        const bool endangered = meta_stm(par.meta) == 0 ? in_critical_endangerment<0>(c) : in_critical_endangerment<1>(c);
        if (!endangered) n++;
    }
};
__global__ void __launch_bounds__(BLOCK) k_selfcheck(Soa in, size_t n, unsigned long long *out) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const Pos p = soa_load(in, i);
    const u32 fast_legal = count_moves<false>(p), fast_reckless = count_moves<true>(p);
    CountSink legal, reckless;
    lookahead<false>(p, legal);
    lookahead<true>(p, reckless);
    ExactLegalSink exact{p};
    lookahead<true>(p, exact);
    if (fast_legal != legal.n || legal.n != exact.n || fast_reckless != reckless.n) {
        atomicAdd(out, 1ull);
        atomicMin(out + 1, (unsigned long long)i);
    }
    atomicAdd(out + 2, (unsigned long long)legal.n);
    atomicAdd(out + 3, (unsigned long long)reckless.n);
}

// ---- life.rs:678-690 for an explicit team ------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_endangered(Soa in, size_t n, const unsigned char *__restrict__ teams, unsigned char *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const Pos p = soa_load(in, i);
    out[i] = teams[i] == 0 ? in_critical_endangerment<0>(p) : in_critical_endangerment<1>(p);
}

// ---- seeded random playouts (DESIGN.md "playout set"; restated on the CPU by oracle lo_playout) -------
__device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ Pos opening_position() { // life.rs:184-215
    Pos p;
    p.pl[0] = 0xFF00ull; p.pl[1] = 0x42ull; p.pl[2] = 0x24ull; p.pl[3] = 0x81ull; p.pl[4] = 0x08ull; p.pl[5] = 0x10ull;
    p.pl[6] = 0xFF00ull << 40; p.pl[7] = 0x42ull << 56; p.pl[8] = 0x24ull << 56; p.pl[9] = 0x81ull << 56;
    p.pl[10] = 0x08ull << 56; p.pl[11] = 0x10ull << 56;
    p.meta = make_meta(0, 0xF, 0xFF);
    return p;
}
struct SelectSink { // keeps the k-th commit's descriptor
    u32 n = 0, k = 0, chosen = 0;
    LL_HDM void move(u32 job, int from, int to, u32 asc, u32 flags) {
        if (n == k) chosen = mv_pack(job, from, to, asc, flags);
        n++;
    }
};
__global__ void __launch_bounds__(BLOCK) k_playout(u64 seed, u64 first_index, size_t n, Soa out) {
    const size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const u64 h0 = mix64(seed + first_index + i);
    Pos p = opening_position();
    for (u64 attempt = 0; attempt < 1024; attempt++) {
        const u64 h1 = mix64(h0 ^ (attempt * 0xD1B54A32D192ED03ull));
        p = opening_position();
        const int plies = 16 + (int)(((mix64(h1) >> 32) * 45ull) >> 32);
        int k = 0;
        for (; k < plies; k++) {
            const u32 count = count_moves<false>(p);
            if (count == 0) break;
            const u64 r = mix64(h1 + (u64)(k + 1) * 0x9E3779B97F4A7C15ull) >> 32;
            SelectSink select;
            select.k = (u32)((r * (u64)count) >> 32);
            lookahead<false>(p, select);
            Pos c;
            make_child(p, select.chosen, c);
            p = c;
        }
        if (k == plies) break;
    }
    soa_store(out, i, p);
}

} // namespace ll

// ---- integer-issue microbenchmark (measurement utility for bench.py's integer roofline) ---------------
// 16 independent chains per thread; KIND 0: one LOP3 per chain step (ALU pipe only), KIND 1: LOP3 + IMAD
// alternating (ALU + FMA pipes, the best case for the issue port).  Returns nothing useful in `out`.
namespace ll {
template <int KIND> __global__ void __launch_bounds__(256) k_int_peak(u32 *out, int iters, u32 seed) {
    u32 a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = seed * (threadIdx.x + 1) + j * 0x9E3779B9u;
    const u32 b = seed | 0x10001u, c = ~seed;
#pragma unroll 8
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) { // exactly one SASS instruction per chain step
            if (KIND == 0 || (j & 1) == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(a[(j + 1) & 15]), "r"(b));
            else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(c), "r"(a[(j + 1) & 15]));
        }
    }
    u32 r = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) r ^= a[j];
    if (r == 0x12345678u) out[0] = r; // keeps the chains alive
}
} // namespace ll
