// svo_build.cuh -- device-driven octree build ("fast path": no -levels, at least two pyramid levels per slab).
//
// Replaces OctreeBuilder::addVoxel x N + finalizeTree (src/svo_builder/OctreeBuilder.cpp:34-168) with passes that need
// NO host read-back between them: every count lives in a device-resident BuildInfo, launch grids are sized from the
// capacities of the previous build, and a build whose lists outgrow them aborts itself (pyramid intact) and is
// repeated by the host with exact sizes.
//
//   k_dense_scan    one single-pass look-back scan over a DENSE pyramid level j >= 1 (the voxelizer maintains levels 0
//                   and 1 with fire-and-forget reductions): compacts the non-zero words into the tile list
//                   (key, mask, child prefix fc), sets the words' bits in dense level j+1, counts tiles and children
//   k_small_levels  the same for all levels with <= 4096 dense words, in ONE block
//   k_brick_gather / k_scan_lookback<3> / k_brick_prefix   the brick level: gathers the <= 64 brick words of every level-1
//                   tile from dense level 0 into the brick list (Morton layout), scans (leaves, brick subtree sizes,
//                   level-1 subtree sizes) over the level-1 tiles, writes leaf ranks fc and size prefixes ps of the bricks
//   k_shard_merge   the shared upper levels from the exchanged subtree table, in one block: global counts, this
//                   rank's file range, bases of its top tiles, the upper records inside its range
//                   (same arithmetic as the host-side svo_shard_layout_from_table)
#pragma once
#include "svo_kernels.cuh"
#include "svo_dispatch.cuh"

namespace svo {

// Classic path (tiny grids, -levels): dense level j+1 from dense level j, one thread per word. dst is pre-biased
// (indexed with global word indices) and clean.
__global__ void __launch_bounds__(256) k_pyramid_up(const unsigned long long* src, unsigned long long n, unsigned long long bias, unsigned long long* dst) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (src[i] != 0ULL) {
        const unsigned long long g = bias + i;
        red_or(dst + (g >> 6), 1ULL << (g & 63));
    }
}

// ---------------------------------------------------------------------------
// k_dense_scan
// ---------------------------------------------------------------------------
constexpr int DS_THREADS = 256, DS_ITEMS = 8;              // a thread reads 8 consecutive words (64 bytes) per sub-tile
constexpr unsigned long long SMALL_LEVEL_WORDS = 4096;     // dense levels up to this size are walked by k_small_levels

struct DenseScanJob {
    const unsigned long long* dense;      // the slab's dense words of level j
    unsigned long long n, bias;           // word count; global word index = bias + i
    unsigned long long* next;             // dense level j + 1, pre-biased (indexed with the global index >> 6); NULL at the top local level
    unsigned long long* key; unsigned long long* mask; unsigned long long* fc;     // tile list of level j
    unsigned long long cap, cap_child;    // capacities of the lists of level j and level j - 1
    int j, count_only;                    // count_only: no list writes, no overflow flags (first build of a context)
    BuildInfo* info;
    unsigned long long* state; unsigned long long* ticket; unsigned long long ticket_base, epoch;
};

// K sub-tiles of 2048 words per block (K = 1 for small levels, 4 for the 10^8-word levels of the big grids: the look-back
// costs a block a fixed latency, so a block must bring enough bytes along -- at K = 1 the 1 GiB level 1 of an 8192^3 grid
// was scanned at 1 TB/s). Sweep 1 keeps only counts and one "non-zero" bit per word (the levels are > 99 % zeros); after the
// look-back, sweep 2 reads the few non-zero words again (L1 / L2 hits) and writes the list.
template <int K>
__global__ void __launch_bounds__(DS_THREADS) k_dense_scan(DenseScanJob Dj) {
    constexpr unsigned long long TILE = (unsigned long long)DS_THREADS * DS_ITEMS * K;
    __shared__ unsigned long long s_tile, s_prefix[2];
    if (threadIdx.x == 0) s_tile = atomicAdd(Dj.ticket, 1ULL) - Dj.ticket_base;
    __syncthreads();
    const unsigned long long tile = s_tile;
    if (tile * TILE >= Dj.n || build_aborted(Dj.info)) return;
    // chunk-major: sub-tile k is 2048 consecutive words, thread t owns its words [8 t, 8 t + 8) (64 bytes; a warp reads 2 KB
    // per sub-tile). Element order of the scan = (sub-tile, thread, word).
    const unsigned long long tile0 = tile * TILE;
    unsigned nzbits = 0;
    unsigned long long ex[K], run_before[K];
    unsigned long long run = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const unsigned long long b0 = tile0 + (unsigned long long)k * (DS_THREADS * DS_ITEMS) + (unsigned long long)threadIdx.x * DS_ITEMS;
        unsigned long long w[DS_ITEMS];
        if (b0 + DS_ITEMS <= Dj.n) {
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(Dj.dense + b0);      // b0 is a multiple of 8 words
#pragma unroll
            for (int i = 0; i < DS_ITEMS / 2; i++) { const ulonglong2 v = __ldg(p + i); w[2 * i] = v.x; w[2 * i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < DS_ITEMS; i++) w[i] = b0 + i < Dj.n ? __ldg(Dj.dense + b0 + i) : 0ULL;
        }
        unsigned nz = 0, pc = 0;
#pragma unroll
        for (int i = 0; i < DS_ITEMS; i++) {
            if (w[i] != 0ULL) { nzbits |= 1u << (k * DS_ITEMS + i); nz++; pc += (unsigned)__popcll(w[i]); }
        }
        unsigned long long total;
        ex[k] = block_excl_scan(((unsigned long long)pc << 16) | nz, total);    // nz <= 2048 per sub-tile
        run_before[k] = run;
        run += total;                                                            // (counts of different sub-tiles add without carry: 4 * 2048 < 65536)
    }
    if (threadIdx.x < 32) {
        const unsigned long long tot[2] = { run & 0xffffULL, run >> 16 };
        unsigned long long pre[2];
        lookback<2>(Dj.state, Dj.epoch, tile, tot, pre, &Dj.info->overflow);
        if (threadIdx.x == 0) { s_prefix[0] = pre[0]; s_prefix[1] = pre[1]; }
    }
    __syncthreads();
    unsigned long long up_word = ~0ULL, up_bits = 0ULL;
#pragma unroll
    for (int k = 0; k < K; k++) {
        unsigned bits = (nzbits >> (k * DS_ITEMS)) & 0xffu;
        if (!bits) continue;
        const unsigned long long b0 = tile0 + (unsigned long long)k * (DS_THREADS * DS_ITEMS) + (unsigned long long)threadIdx.x * DS_ITEMS;
        const unsigned long long pk = run_before[k] + ex[k];
        unsigned long long at = s_prefix[0] + (pk & 0xffffULL), cp = s_prefix[1] + (pk >> 16);
        while (bits) {
            const int i = __ffs(bits) - 1;
            bits &= bits - 1;
            const unsigned long long w = __ldg(Dj.dense + b0 + i);
            const unsigned long long g = Dj.bias + b0 + i;
            if (!Dj.count_only && at < Dj.cap) { Dj.key[at] = g; Dj.mask[at] = w; Dj.fc[at] = cp; }
            at++;
            cp += (unsigned)__popcll(w);
            if (Dj.next) {
                if ((g >> 6) != up_word) { if (up_bits) red_or(Dj.next + up_word, up_bits); up_word = g >> 6; up_bits = 0ULL; }
                up_bits |= 1ULL << (g & 63);
            }
        }
    }
    if (up_bits) red_or(Dj.next + up_word, up_bits);
    const unsigned long long total = run;
    if (threadIdx.x == 0 && (tile + 1) * TILE >= Dj.n) {       // the last tile: totals
        const unsigned long long n_tiles = s_prefix[0] + (total & 0xffffULL), n_children = s_prefix[1] + (total >> 16);
        Dj.info->count[Dj.j] = n_tiles;
        if (Dj.j == 1) Dj.info->count[0] = n_children;
        if (!Dj.count_only) {
            if (n_tiles <= Dj.cap) Dj.fc[n_tiles] = n_children;   // fc[n] = total: the lists have room for cap + 2 entries
            if (n_tiles > Dj.cap) atomicOr(&Dj.info->overflow, 1ULL << Dj.j);
            if (n_children > Dj.cap_child) atomicOr(&Dj.info->overflow, 1ULL << (Dj.j - 1));
        }
    }
}

// ---------------------------------------------------------------------------
// k_small_levels: dense levels j0..J (each <= SMALL_LEVEL_WORDS words) in one block: tile list of every level and the
// dense level above it. Dense level j0 is complete when the kernel starts (the last k_dense_scan set its bits).
// ---------------------------------------------------------------------------
struct SmallLevelsJob {
    int j0, J, count_only;
    unsigned long long* dense[MAX_LEVELS];        // local arrays (not biased)
    unsigned long long nwords[MAX_LEVELS], bias[MAX_LEVELS];
    unsigned long long* key[MAX_LEVELS]; unsigned long long* mask[MAX_LEVELS]; unsigned long long* fc[MAX_LEVELS];
    unsigned long long cap[MAX_LEVELS];
    BuildInfo* info;
};
__device__ __forceinline__ void small_levels_body(const SmallLevelsJob& S) {
    for (int j = S.j0; j <= S.J; j++) {
        const unsigned long long n = S.nwords[j];
        unsigned long long c_nz = 0, c_pc = 0;
        for (unsigned long long b = 0; b < n; b += blockDim.x) {
            const unsigned long long i = b + threadIdx.x;
            const unsigned long long w = i < n ? __ldcg(S.dense[j] + i) : 0ULL;
            unsigned long long total;
            const unsigned long long ex = block_excl_scan(((unsigned long long)__popcll(w) << 32) | (w != 0ULL ? 1ULL : 0ULL), total);
            if (w != 0ULL) {
                const unsigned long long at = c_nz + (ex & 0xffffffffULL), g = S.bias[j] + i;
                if (!S.count_only && at < S.cap[j]) { S.key[j][at] = g; S.mask[j][at] = w; S.fc[j][at] = c_pc + (ex >> 32); }
                if (j < S.J) red_or(S.dense[j + 1] + ((g >> 6) - S.bias[j + 1]), 1ULL << (g & 63));
            }
            c_nz += total & 0xffffffffULL;
            c_pc += total >> 32;
        }
        if (threadIdx.x == 0) {
            S.info->count[j] = c_nz;
            if (j == 1) S.info->count[0] = c_pc;
            if (!S.count_only) {
                if (c_nz <= S.cap[j]) S.fc[j][c_nz] = c_pc;
                if (c_nz > S.cap[j]) atomicOr(&S.info->overflow, 1ULL << j);
                if (c_pc > S.cap[j - 1]) atomicOr(&S.info->overflow, 1ULL << (j - 1));
            }
        }
        __threadfence();
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_small_levels(SmallLevelsJob S) {
    if (build_aborted(S.info)) return;
    small_levels_body(S);
}

// ---------------------------------------------------------------------------
// The brick pass: three launches.
//   k_brick_gather   one warp per level-1 tile (four tiles per warp, all gathers in flight together): reads the tile's
//                    <= 64 brick words from dense level 0 (64 GiB at 8192^3: every gather is a DRAM sector, read exactly
//                    once), converts them to the Morton layout and writes them -- with their keys -- at their final
//                    positions of the brick list (the level-1 list's child prefix fc gives them: no scan needed);
//                    leaves and brick-subtree records of the tile go to tile_tot[t]
//   k_scan_lookback<3, BrickTileTotals>   exclusive prefixes per level-1 tile: leaves, brick records, level-1 subtree
//                    sizes (L1.ps). 64 times fewer elements than bricks: the look-back chain is a few dozen tiles long
//   k_brick_prefix   one warp per level-1 tile again: reads its bricks back (coalesced, L2 hits at the sizes that fit) and
//                    writes their leaf ranks (fc) and size prefixes (ps)
// An earlier single kernel held all this in one look-back chain over the level-1 tiles; at 8192^3 it spent half of its
// time at the block barriers around the chain while the DRAM gathers of only one wave of blocks were in flight.
// ---------------------------------------------------------------------------
constexpr int BP_WARPS = 8, BP_PER_WARP = 4, BP_TILE = BP_WARPS * BP_PER_WARP;     // level-1 tiles per block
struct BrickJob {
    Level L1;                             // key, mask, fc in; ps out (n + 1)
    Level L0;                             // key, mask, fc, ps out
    const unsigned long long* dense0;     // pre-biased: indexed with global level-0 word indices
    uint32_t* tileidx;                    // pre-biased dense map word -> compact index (payload owner pass) or NULL
    unsigned long long* tile_lp;          // n1 + 1: in: leaves | records << 32 per tile (gather); out: leaf prefix (scan)
    unsigned long long* tile_sp;          // n1 + 1: brick-record prefix per level-1 tile
    BuildInfo* info;
};
__global__ void __launch_bounds__(BP_WARPS * 32) k_brick_gather(BrickJob B) {
    if (build_aborted(B.info)) return;
    const unsigned long long n1 = level_n(B.L1);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long tw0 = ((unsigned long long)blockIdx.x * BP_WARPS + wid) * BP_PER_WARP;
    if (tw0 >= n1) return;
    // lane q of the warp holds the descriptor of the warp's q-th tile
    unsigned long long myW1 = 0ULL, myK1 = 0ULL, myF1 = 0ULL;
    if (lane < BP_PER_WARP && tw0 + lane < n1) { myW1 = B.L1.mask[tw0 + lane]; myK1 = B.L1.key[tw0 + lane]; myF1 = B.L1.fc[tw0 + lane]; }
    unsigned long long m[BP_PER_WARP][2];
#pragma unroll
    for (int q = 0; q < BP_PER_WARP; q++) {                   // all the gathers first
        const unsigned long long W1 = __shfl_sync(0xffffffffu, myW1, q), K1 = __shfl_sync(0xffffffffu, myK1, q);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int bit = lane + 32 * h;
            m[q][h] = 0ULL;
            if ((W1 >> bit) & 1ULL) m[q][h] = __ldcg(B.dense0 + ((K1 << 6) | (unsigned long long)bit));
        }
    }
#pragma unroll
    for (int q = 0; q < BP_PER_WARP; q++) {
        const unsigned long long W1 = __shfl_sync(0xffffffffu, myW1, q), K1 = __shfl_sync(0xffffffffu, myK1, q), F1 = __shfl_sync(0xffffffffu, myF1, q);
        if (W1 == 0ULL) continue;                              // (warp-uniform) beyond the list
        unsigned tot = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int bit = lane + 32 * h;
            if (!((W1 >> bit) & 1ULL)) continue;
            const unsigned long long w = linear_to_morton64(m[q][h]);
            const unsigned leaves = (unsigned)__popcll(w);
            tot += leaves | ((leaves + (unsigned)__popc(nonzero_bytes(w))) << 16);      // <= 4096 leaves, <= 4608 records per tile: no carry
            const unsigned long long c = F1 + __popcll(W1 & lowmask(bit));
            if (c >= B.L0.cap) continue;                       // (the overflow flag is already set by the level-1 scan)
            const unsigned long long ck = (K1 << 6) | (unsigned long long)bit;
            B.L0.key[c] = ck;
            B.L0.mask[c] = w;
            if (B.tileidx) B.tileidx[ck] = (uint32_t)c;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
        if (lane == 0) B.tile_lp[tw0 + q] = (unsigned long long)(tot & 0xffffu) | ((unsigned long long)(tot >> 16) << 32);
    }
}
struct BrickTileTotals {   // per level-1 tile: leaves, brick records, subtree size of the tile
    const unsigned long long* tot; const unsigned long long* mask1;
    __device__ void operator()(unsigned long long t, unsigned long long (&e)[3]) const {
        const unsigned long long v = tot[t], W1 = mask1[t];
        e[0] = v & 0xffffffffULL; e[1] = v >> 32;
        e[2] = e[1] + (unsigned long long)(__popcll(W1) + __popc(nonzero_bytes(W1)));
    }
};
// A warp takes tiles_per_warp (32; fewer when the level is small) consecutive level-1 tiles: their bricks are one contiguous run of the brick list, and the prefixes of
// the run's first tile seed a plain running sum -- no per-tile logic, all loads and stores coalesced.
__global__ void __launch_bounds__(BP_WARPS * 32) k_brick_prefix(BrickJob B, int tiles_per_warp) {
    if (build_aborted(B.info)) return;
    const unsigned long long n1 = level_n(B.L1);
    const int lane = threadIdx.x & 31;
    const unsigned long long w = (unsigned long long)blockIdx.x * BP_WARPS + (threadIdx.x >> 5);
    if (w == 0 && lane == 0) {                                 // totals behind the lists (the scan wrote lp[n1], sp[n1]; both 0 when n1 == 0)
        const unsigned long long n0 = B.info->count[0], nl = B.tile_lp[n1], ns = B.tile_sp[n1];
        if (n0 <= B.L0.cap) { B.L0.fc[n0] = nl; B.L0.ps[n0] = ns; }
        B.info->n_leaves_local = nl;
        B.info->n_brick_records = ns;
    }
    const unsigned long long t0 = w * (unsigned long long)tiles_per_warp;
    if (t0 >= n1) return;
    const unsigned long long t1 = min(t0 + (unsigned long long)tiles_per_warp, n1);
    const unsigned long long c0 = B.L1.fc[t0];
    const unsigned long long c1 = min(B.L1.fc[t1], B.L0.cap);
    unsigned long long lp = B.tile_lp[t0], sp = B.tile_sp[t0];
    constexpr int U = 4;                                       // rounds of 32 bricks whose loads are in flight together
    for (unsigned long long cb = c0; cb < c1; cb += 32 * U) {
        unsigned long long m[U];
#pragma unroll
        for (int k = 0; k < U; k++) { const unsigned long long c = cb + 32 * k + lane; m[k] = c < c1 ? B.L0.mask[c] : 0ULL; }
#pragma unroll
        for (int k = 0; k < U; k++) {
            const unsigned long long c = cb + 32 * k + lane;
            const unsigned leaves = (unsigned)__popcll(m[k]);
            const unsigned v = leaves | ((leaves + (unsigned)__popc(nonzero_bytes(m[k]))) << 16);    // 32 bricks: <= 2048 leaves, <= 2304 records
            unsigned inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned a = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += a; }
            const unsigned ex = inc - v;
            if (c < c1) { B.L0.fc[c] = lp + (ex & 0xffffu); B.L0.ps[c] = sp + (ex >> 16); }
            const unsigned all = __shfl_sync(0xffffffffu, inc, 31);
            lp += all & 0xffffu; sp += all >> 16;
        }
    }
}

// ---------------------------------------------------------------------------
// k_shard_merge: the shared upper levels (above the top local level J) from the complete subtree table
// {mask, S, leaves, internals} per global level-J word. One block. The same arithmetic as shard_merge_compute()
// in svo_api.cu (host, behind svo_shard_layout_from_table) and as k_emit_upper.
// ---------------------------------------------------------------------------
struct MergeJob {
    const unsigned long long* table; unsigned long long WJ;
    int J, top, d_even, rank, world;
    unsigned long long nW[MAX_LEVELS];            // global word counts of levels J..top
    unsigned long long wj0, wj1;                  // own entries [wj0, wj1)
    unsigned long long* M[MAX_LEVELS]; unsigned long long* S[MAX_LEVELS]; unsigned long long* B[MAX_LEVELS];   // scratch, levels J..top
    unsigned long long* rpos; unsigned long long* rrec; unsigned long long rcap;      // upper records inside this rank's range
    const unsigned long long* keyJ; unsigned long long* baseJ; unsigned long long capJ; // own level-J tile list
    BuildInfo* info;
    unsigned long long nodes_cap;                 // capacity of the node buffer (speculative emission); ~0 when it is sized afterwards
};
__device__ __forceinline__ void shard_merge_body(const MergeJob& Mj) {
    __shared__ unsigned long long s_leaves, s_before, s_own, s_lo, s_hi, s_nrec, s_nodes;
    const int J = Mj.J, top = Mj.top;
    if (threadIdx.x == 0) { s_leaves = 0; s_before = 0; s_own = 0; s_lo = ~0ULL; s_hi = ~0ULL; s_nrec = 0; }
    __syncthreads();
    {   // level J: columns of the table, leaf totals
        unsigned long long a = 0, b = 0, o = 0;
        for (unsigned long long e = threadIdx.x; e < Mj.WJ; e += blockDim.x) {
            const unsigned long long mk = Mj.table[4 * e], lv = Mj.table[4 * e + 2];
            Mj.M[J][e] = mk; Mj.S[J][e] = Mj.table[4 * e + 1]; Mj.B[J][e] = 0ULL;
            if (mk) { a += lv; if (e < Mj.wj0) b += lv; else if (e < Mj.wj1) o += lv; }
        }
        if (a) atomicAdd(&s_leaves, a);
        if (b) atomicAdd(&s_before, b);
        if (o) atomicAdd(&s_own, o);
    }
    __syncthreads();
    // bottom-up: masks and subtree sizes of the upper levels
    for (int j = J + 1; j <= top; j++) {
        for (unsigned long long w = threadIdx.x; w < Mj.nW[j]; w += blockDim.x) {
            unsigned long long W = 0, sz = 0;
            for (int b = 0; b < 64; b++) {
                const unsigned long long ch = w * 64 + b;
                if (ch < Mj.nW[j - 1] && Mj.M[j - 1][ch]) { W |= 1ULL << b; sz += Mj.S[j - 1][ch]; }
            }
            if (W) sz += (unsigned long long)(__popcll(W) + __popc(nonzero_bytes(W)));
            Mj.M[j][w] = W; Mj.S[j][w] = sz; Mj.B[j][w] = 0ULL;
        }
        __syncthreads();
    }
    const unsigned long long leaves_total = s_leaves;
    const unsigned long long n_nodes = leaves_total == 0 ? 1ULL : Mj.S[top][0] + (Mj.d_even ? 1ULL : 0ULL);
    // top-down: bases (pass 0), then the upper records that fall into this rank's range (pass 1)
    for (int pass = 0; pass < 2; pass++) {
        const unsigned long long lo = s_lo, hi = s_hi;
        auto push = [&](unsigned long long pos, unsigned long long d1, unsigned long long d2) {
            if (pass == 0 || pos < lo || pos >= hi) return;
            const unsigned long long at = atomicAdd(&s_nrec, 1ULL);
            if (at < Mj.rcap) { Mj.rpos[at] = pos; Mj.rrec[3 * at] = 0ULL; Mj.rrec[3 * at + 1] = d1; Mj.rrec[3 * at + 2] = d2; }
        };
        for (int j = top; j > J; j--) {
            for (unsigned long long w = threadIdx.x; w < Mj.nW[j]; w += blockDim.x) {
                const unsigned long long W = Mj.M[j][w];
                if (!W) continue;
                const unsigned long long base = Mj.B[j][w], sz = Mj.S[j][w];
                const uint32_t nzb = nonzero_bytes(W);
                unsigned long long acc = 0;
                for (int k = 0; k < 8; k++) {
                    const uint32_t byte = (uint32_t)((W >> (8 * k)) & 0xffULL);
                    if (!byte) continue;
                    const unsigned long long before = (unsigned long long)__popcll(W & lowmask(8 * k));
                    for (int b = 0; b < 8; b++) if ((byte >> b) & 1u) {
                        const unsigned long long ch = w * 64 + 8 * k + b;
                        if (pass == 0) Mj.B[j - 1][ch] = base + acc + before;
                        acc += Mj.S[j - 1][ch];
                    }
                    const unsigned long long blk = base + acc + before;
                    unsigned long long r = 0;
                    for (int b = 0; b < 8; b++) if ((byte >> b) & 1u) {
                        const unsigned long long ch = w * 64 + 8 * k + b;
                        const uint32_t gnz = nonzero_bytes(Mj.M[j - 1][ch]);
                        push(blk + r++, Mj.B[j - 1][ch] + Mj.S[j - 1][ch] - (unsigned long long)__popc(gnz), child_offsets(gnz));
                    }
                    push(base + sz - (unsigned long long)__popc(nzb) + (unsigned long long)__popc(nzb & ((1u << k) - 1u)), blk, child_offsets(byte));
                }
                if (j == top && Mj.d_even) push(sz, base + sz - (unsigned long long)__popc(nzb), child_offsets(nzb));
            }
            __syncthreads();
        }
        if (pass == 0) {
            // this rank's range: from the base of its first tile to the base of the first tile of the next rank
            unsigned long long lo_c = ~0ULL, hi_c = ~0ULL;
            for (unsigned long long e = threadIdx.x; e < Mj.WJ; e += blockDim.x) {
                if (!Mj.M[J][e]) continue;
                const unsigned long long bse = Mj.B[J][e];
                if (e >= Mj.wj0 && bse < lo_c) lo_c = bse;
                if (e >= Mj.wj1 && bse < hi_c) hi_c = bse;
            }
            if (lo_c != ~0ULL) atomicMin(&s_lo, lo_c);
            if (hi_c != ~0ULL) atomicMin(&s_hi, hi_c);
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long nlo = Mj.rank == 0 ? 0ULL : (s_lo == ~0ULL ? n_nodes : s_lo);
                unsigned long long nhi = Mj.rank == Mj.world - 1 ? n_nodes : (s_hi == ~0ULL ? n_nodes : s_hi);
                if (leaves_total == 0) { nlo = Mj.rank == 0 ? 0ULL : 1ULL; nhi = 1ULL; }
                s_lo = nlo; s_hi = nhi; s_nodes = n_nodes;
            }
            __syncthreads();
        }
    }
    if (leaves_total == 0 && Mj.rank == 0 && threadIdx.x == 0) {
        // empty grid: finalizeTree pads everything and writes a null root (OctreeBuilder.cpp:36-42)
        const unsigned long long at = atomicAdd(&s_nrec, 1ULL);
        if (at < Mj.rcap) { Mj.rpos[at] = 0ULL; Mj.rrec[3 * at] = 0ULL; Mj.rrec[3 * at + 1] = 0ULL; Mj.rrec[3 * at + 2] = ~0ULL; }
    }
    // bases of this rank's own level-J tiles
    {
        unsigned long long nJ = Mj.info->count[J];
        if (nJ > Mj.capJ) nJ = Mj.capJ;
        for (unsigned long long i = threadIdx.x; i < nJ; i += blockDim.x) Mj.baseJ[i] = Mj.B[J][Mj.keyJ[i]];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        BuildInfo* I = Mj.info;
        I->n_voxels = leaves_total; I->n_nodes = s_nodes;
        I->leaf_offset = s_before;
        I->node_lo = s_lo; I->node_hi = s_hi;
        I->n_upper = s_nrec < Mj.rcap ? s_nrec : Mj.rcap;
        if (s_own != I->n_leaves_local) atomicOr(&I->overflow, 1ULL << 41);        // the table does not match this rank's tiles
        if (s_hi - s_lo > Mj.nodes_cap) atomicOr(&I->overflow, 1ULL << 32);
        if (s_nrec > Mj.rcap) atomicOr(&I->overflow, 1ULL << 42);
    }
    __threadfence();
    __syncthreads();
}
__global__ void __launch_bounds__(1024) k_shard_merge(MergeJob Mj) {
    if (build_aborted(Mj.info)) return;
    shard_merge_body(Mj);
}

// ---------------------------------------------------------------------------
// k_top: everything of a build that ONE block does -- the small upper levels (lists, subtree sizes), this rank's table
// entries, the merge of the shared levels, the scattered upper records and the emission of the top levels -- as stages
// of a single launch, so that a build pays one launch latency for them instead of six. The host picks the stages:
// all of them on one GPU; [lists, sizes, table] | exchange | [merge, scatter, emit] when sharded; the merge alone
// when the node buffer still has to be sized on the host.
// ---------------------------------------------------------------------------
enum { TOP_LISTS = 1, TOP_SIZES = 2, TOP_TABLE = 4, TOP_MERGE = 8, TOP_SCATTER = 16, TOP_EMIT = 32, TOP_XPUSH = 64, TOP_XWAIT = 128 };
struct TopJob {
    int stages;
    SmallLevelsJob S;
    FusedJob F;                       // lv[jB..J] views; jf_up = first level of the size pass, jf = last level NOT emitted here
    TableFillJob T; unsigned long long table_words;
    MergeJob M;
    XchgJob X; SliceCtrl* own_ctrl;   // sharded, peer-memory exchange of the table (svo_dispatch.cuh): push after TABLE, wait before MERGE
    BuildInfo* info;
};
__global__ void __launch_bounds__(1024) k_top(TopJob P) {
    if ((P.stages & TOP_LISTS) && !build_aborted(P.info)) small_levels_body(P.S);
    if ((P.stages & TOP_SIZES) && !build_aborted(P.info)) fused_up_body(P.F, P.F.jf_up);
    if ((P.stages & TOP_TABLE) && !build_aborted(P.info)) table_fill_body(P.T, P.table_words);
    if (P.stages & TOP_XPUSH) {
        // this rank's entries into every peer's exchange table, then the flag. Runs for an aborted build as well: the
        // peers must learn that this rank has nothing to offer (poisoned flag -> SVO_E_RETRY on every rank)
        const bool poisoned = build_aborted(P.info);
        if (!poisoned)
            for (int p = 0; p < P.X.world; p++) {
                unsigned long long* dst = P.X.xtable[p];
                for (unsigned long long i = threadIdx.x; i < P.X.n; i += blockDim.x) dst[P.X.lo + i] = P.X.src[P.X.lo + i];
            }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < P.X.world) {
            __threadfence_system();
            *(volatile unsigned long long*)&P.X.ctrl[threadIdx.x]->flag[2][P.X.me] = 2ULL * P.X.epoch + (poisoned ? 1ULL : 0ULL);
        }
        __syncthreads();
    }
    if (P.stages & TOP_XWAIT) {
        // ONE block spins here (like the k_xchg_wait launch it replaces): it cannot keep a peer's kernels off this GPU
        if ((int)threadIdx.x < P.X.world) {
            const volatile unsigned long long* f = &P.own_ctrl->flag[2][threadIdx.x];
            const long long t0 = clock64();
            unsigned long long v;
            while (((v = *f) >> 1) < P.X.epoch) {
                if (clock64() - t0 > 8000000000LL) { P.own_ctrl->error = 3ULL; break; }
                __nanosleep(100);
            }
            if ((v >> 1) == P.X.epoch && (v & 1ULL)) atomicOr(&P.info->overflow, 1ULL << 43);
            __threadfence_system();
        }
        __threadfence();
        __syncthreads();
    }
    if ((P.stages & TOP_MERGE) && !build_aborted(P.info)) shard_merge_body(P.M);
    if ((P.stages & TOP_SCATTER) && !build_aborted(P.info)) {
        scatter_records_body(P.M.rpos, P.M.rrec, P.M.rcap, &P.info->n_upper, P.F.E, threadIdx.x, blockDim.x);
        __threadfence();
        __syncthreads();
    }
    if ((P.stages & TOP_EMIT) && !build_aborted(P.info)) fused_emit_body(P.F);
}

}  // namespace svo
