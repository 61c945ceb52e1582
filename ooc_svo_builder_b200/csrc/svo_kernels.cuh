// svo_kernels.cuh -- the sm_100a kernels of the voxelize-and-build path.
//
// Data layout in HBM:
//   * triangles: the .tridata image as-is (AoS, 36 B or 84 B records)
//   * bit-grid pyramid: level 0 is the Morton-ordered occupancy bit-grid, one
//     64-bit word per 4x4x4 brick (= one octree node at depth D-2: its 8 bytes are
//     the child masks of its depth D-1 children). Level j+1 has one bit per
//     level-j word (non-zero), so a level-j word is the octree node at depth
//     D-2(j+1) with two tree levels packed in it. The pyramid is kept dense; the
//     voxelizer maintains the upper levels with the return value of its atomicOr.
//   * compact tile lists per level (key, mask, child prefix, subtree-size prefix,
//     file base), built top-down from the pyramid, so every pass after the
//     voxelizer is O(occupied), never O(grid).
// Kernels in this file: voxelizer (k_vox_warp, k_vox_small, k_vox_queued), partitioner (k_bin, k_owner_filter), scans
// (k_scan_small, k_scan_lookback), pyramid compaction (k_level_counts, k_compact_top, k_expand, k_fused_down / _up),
// emission (k_fused_emit, k_emit_upper, k_emit_leaf, k_emit_leaf_levels, k_levels_data, k_payload), clean-up
// (k_sparse_clear_all). The multi-GPU exchange kernels are in svo_dispatch.cuh.
#pragma once
#include <type_traits>
#include "svo_device.cuh"

namespace svo {

constexpr int MAX_LEVELS = 12;        // ceil(D/2), D <= 21 (libmorton's 21 bits per axis)
#ifndef SVO_VOX_BLOCK
#define SVO_VOX_BLOCK 128
#endif
#ifndef SVO_VOX_MINBLOCKS
#define SVO_VOX_MINBLOCKS 4
#endif
constexpr int VOX_BLOCK = SVO_VOX_BLOCK;   // threads per block of the small-bbox voxelizer
constexpr int WARPS_PER_BLOCK = 8;
// ticket counters of the voxelizer's unit scheduler: VoxJob::qcount[VOX_TICKET_BASE + (pass * VOX_TICKETS + c) * VOX_TICKET_STRIDE]
constexpr int VOX_TICKETS = 8, VOX_TICKET_STRIDE = 16, VOX_TICKET_BASE = 16, VOX_QCOUNT_WORDS = VOX_TICKET_BASE + 2 * VOX_TICKETS * VOX_TICKET_STRIDE;

constexpr int MAX_WORLD = 16;         // ranks of a sharded build that exchange through peer memory

// Sharded builds with remote staging: the triangle file stays where it was loaded -- slice s in the HBM of rank s --
// and kernels read the records they need straight from the owner's memory (NVLink loads through peer pointers).
// Global triangle index = file position = (sum of the slice sizes before s) + position in slice s.
struct TriSegs {
    const float* ptr[MAX_WORLD];              // slice s (local or peer pointer)
    const unsigned long long* nslice;         // device: triangles in every slice (published by the peers)
    int n;                                    // 0 = one local array (VoxJob::tris)
};
// smem table of segment starts: base[s] = global index of the first triangle of slice s, base[n] = total
__device__ __forceinline__ void segs_bases(const TriSegs& S, unsigned long long* s_base) {
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int s = 0; s < S.n; s++) { s_base[s] = acc; acc += S.nslice[s]; }
        s_base[S.n] = acc;
    }
    __syncthreads();
}
__device__ __forceinline__ const float* segs_record(const TriSegs& S, const unsigned long long* s_base, uint32_t tri, uint32_t fpt) {
    int s = 0;
    while (s + 1 < S.n && (unsigned long long)tri >= s_base[s + 1]) s++;
    return S.ptr[s] + ((unsigned long long)tri - s_base[s]) * fpt;
}

struct VoxJob {
    const float* tris;                // .tridata image
    uint32_t fpt;                     // floats per triangle: 9 or 21
    uint64_t q_begin, q_end;          // pair range handled by this context (its partitions)
    const uint32_t* pair_tri;         // per-partition index lists, concatenated; NULL = identity (P == 1)
    const uint64_t* part_off;         // P+1 list offsets (NULL when P == 1)
    uint32_t P, k;                    // logical partitions P = 8^k
    uint32_t side;                    // partition side in voxels = g >> k
    uint32_t g;
    float u, unit_div;                // unit length, 1/u (voxelizer.cpp:164)
    int six;                          // opt-in 6-separating variant (svo_params::separability == 6)
    int nl;                           // pyramid levels
    unsigned long long* lvl[MAX_LEVELS];
    unsigned long long* queue[2];     // medium / large work queues: (part << 32 | tri)
    unsigned long long* qcount;       // [0] medium, [1] large, [2] pairs seen (inline mode), [3] queue overflow flag
    unsigned long long small_max, medium_max;
    unsigned small_windows;           // a small pair spans at most this many 4x4x4 windows
    unsigned long long qcap;          // capacity of each work queue (entries); overflow sets qcount[3]
    // inline partition enumeration (pair_tri == NULL, P > 1): world slabs of the partition grid and the
    // Morton range of partitions this context owns
    float bmin[32], bmax[32];
    float inv_slab;                   // ~ 1 / world width of a partition slab (estimate only)
    uint32_t p_first, p_last;
    const uint32_t* subset;               // sharded: triangles that touch this rank's slab (NULL = all triangles)
    const unsigned long long* subset_count;
    // sharded with remote staging: per source rank s, subset[s * pull_cap + i] lists the staging blocks of slice s
    // that touch this rank's slab (written by rank s), pull_counts[s * MAX_WORLD] how many
    TriSegs segs;
    unsigned long long pull_cap;
    const unsigned long long* pull_counts;
    int pull_first;                       // this rank walks the sources in the order pull_first, pull_first + 1, ... (mod n)

    // sharding: level-0 word range owned by this context and the voxel bounding box of its slab.
    // lvl[] / tileidx are biased so that they are indexed with GLOBAL word indices.
    unsigned long long w_lo, w_hi;
    int sb_lo[3], sb_hi[3];
    // payload owner pass
    const unsigned long long* tilemask;   // compact level-0 masks (Morton layout)
    const uint32_t* tileidx;          // dense: level-0 word -> compact tile index
    const unsigned long long* leafprefix;  // exclusive popcount prefix over level-0 tiles
    uint32_t* owner;                  // per leaf: min triangle index that covers it
};

// ---------------------------------------------------------------------------
// hit sinks
// ---------------------------------------------------------------------------
// Binary pass: fire-and-forget 64-bit OR reductions (RED, no return value, nothing to wait for): the brick word at
// level 0 and, unconditionally, the brick's bit in the level-1 word above it. Levels >= 2 are rebuilt from the dense
// level 1 after the voxelizer (k_dense_scan / k_small_levels in svo_build.cuh, k_pyramid_up for the classic path).
__device__ __forceinline__ void red_or(unsigned long long* p, unsigned long long v) {
    asm volatile("red.global.or.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void sink_fill(const VoxJob& J, uint64_t w, uint64_t mask) {
    if (w < J.w_lo || w >= J.w_hi) return;          // another rank's slab
    red_or(&J.lvl[0][w], (unsigned long long)mask);
    if (J.nl > 1) red_or(&J.lvl[1][w >> 6], 1ULL << (w & 63));
}
// Payload owner pass: the reference's first-triangle-wins rule (voxelizer.cpp:263)
// made order independent: owner = min triangle index over all triangles passing.
// `bit` is the LINEAR in-brick index; leaf ranks follow the Morton order of the compact tile mask.
__device__ __forceinline__ void sink_owner_bit(const VoxJob& J, uint64_t w, int bit, uint32_t tri) {
    if (w < J.w_lo || w >= J.w_hi) return;
    const uint32_t t = J.tileidx[w];
    const unsigned long long r = J.leafprefix[t] + __popcll(J.tilemask[t] & lowmask(linear_to_morton_bit(bit)));
    atomicMin(&J.owner[r], tri);
}

// ---------------------------------------------------------------------------
// pair decoding
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pair_partition(const VoxJob& J, uint64_t q) {
    if (J.P == 1) return 0;
    uint32_t lo = 0, hi = J.P;            // find last p with off[p] <= q
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (J.part_off[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t pack_slab(int ix, int iy, int iz) { return (uint32_t)ix | ((uint32_t)iy << 8) | ((uint32_t)iz << 16); }
__device__ __forceinline__ void load_vertices(const VoxJob& J, const unsigned long long* s_base, uint32_t tri, float* v) {
    const float* t = J.segs.n ? segs_record(J.segs, s_base, tri, J.fpt) : J.tris + (size_t)tri * J.fpt;
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = __ldg(t + i);
}

// Restrict a clamped box to the slab this context owns (no-op on a single GPU). Returns false if empty.
__device__ __forceinline__ bool restrict_to_slab(const VoxJob& J, GridBox& b) {
    b.x0 = max(b.x0, J.sb_lo[0]); b.x1 = min(b.x1, J.sb_hi[0]);
    b.y0 = max(b.y0, J.sb_lo[1]); b.y1 = min(b.y1, J.sb_hi[1]);
    b.z0 = max(b.z0, J.sb_lo[2]); b.z1 = min(b.z1, J.sb_hi[2]);
    return b.x0 <= b.x1 && b.y0 <= b.y1 && b.z0 <= b.z1;
}

// Partition slabs touched by the interval [mn, mx]: slab i is kept unless (mx < bmin[i]) or (mn > bmax[i])
// (intersectBoxBox, intersection.h:50-53, per axis). bmin / bmax increase with i, so the kept slabs are the
// range [lo, hi] with lo = first i with !(mn > bmax[i]) and hi = last i with !(mx < bmin[i]) (empty if lo > hi).
// Both ends are found from an estimated slab (inv_w ~ 1 / slab width, a hint only) and fixed up with the
// exact comparisons, so the result is the one a scan over all slabs gives (NaN compares false, as there).
__device__ __forceinline__ void slab_range(const float* bmin, const float* bmax, int n, float inv_w, float mn, float mx, int& lo, int& hi) {
    int i = clampi(__float2int_rz(fmul(mn, inv_w)), 0, n - 1);
    // fast path (almost every triangle lies in ONE slab): slab i is the whole kept range iff it is kept, its lower
    // neighbour is dropped by mn and its upper neighbour by mx -- the same comparisons the loops below would make
    if (!(mn > bmax[i]) && !(mx < bmin[i]) && (i == 0 || mn > bmax[i - 1]) && (i == n - 1 || mx < bmin[i + 1])) { lo = hi = i; return; }
    while (i > 0 && !(mn > bmax[i - 1])) i--;
    while (i < n && (mn > bmax[i])) i++;
    lo = i;
    int j = clampi(__float2int_rz(fmul(mx, inv_w)), 0, n - 1);
    while (j < n - 1 && !(mx < bmin[j + 1])) j++;
    while (j >= 0 && (mx < bmin[j])) j--;
    hi = j;
}

// Warp-aggregated queue push; must be called by all 32 lanes.
__device__ __forceinline__ void warp_push(unsigned long long* counter, unsigned long long* q, unsigned long long cap,
                                          unsigned long long* overflow, bool pred, unsigned long long val) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
        const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
        if (at < cap) q[at] = val; else *overflow = 1ULL;
    }
}

// Morton code of the next brick along one axis: add one inside the axis' bit lane (the other lanes are filled with
// ones so that the carry ripples through them)
template <class U>
__device__ __forceinline__ U morton_inc(U m, U lane_mask) { return ((m | ~lane_mask) + (U)1) & lane_mask; }
// Split the hit mask of a box-aligned 4x4x4 window (linear layout) into the up to 8 bricks it straddles:
// per axis the low part moves up by the window's offset inside the brick, the high part moves down into
// the next brick; in the linear layout these are plain shifts once the part is masked out.
// The window lies inside the clamped box of the pair, which lies inside this rank's slab (a box): no range test.
// BIG = false: grids up to 4096^3 -- bricks have at most 10 bits per axis, word indices fit 32 bits.
template <bool OWNER, bool BIG>
__device__ __forceinline__ void emit_window(const VoxJob& J, int wx, int wy, int wz, unsigned long long hits, uint32_t tri) {
    typedef typename std::conditional<BIG, unsigned long long, uint32_t>::type idx_t;
    const int ox = wx & 3, oy = wy & 3, oz = wz & 3;
    const unsigned long long xlo = (unsigned long long)((1u << (4 - ox)) - 1u) * 0x1111111111111111ULL;
    const unsigned long long ylo = (unsigned long long)((1u << (4 * (4 - oy))) - 1u) * 0x0001000100010001ULL;
    const unsigned long long zlo = lowmask(16 * (4 - oz));
    const uint32_t bx = (uint32_t)(wx >> 2), by = (uint32_t)(wy >> 2), bz = (uint32_t)(wz >> 2);
    idx_t sx[2], sy[2], sz[2];
    if (BIG) { sx[0] = (idx_t)spread3(bx); sy[0] = (idx_t)(spread3(by) << 1); sz[0] = (idx_t)(spread3(bz) << 2); }
    else { sx[0] = (idx_t)spread3_10(bx); sy[0] = (idx_t)(spread3_10(by) << 1); sz[0] = (idx_t)(spread3_10(bz) << 2); }
    sx[1] = morton_inc<idx_t>(sx[0], (idx_t)0x1249249249249249ULL);
    sy[1] = morton_inc<idx_t>(sy[0], (idx_t)0x2492492492492492ULL);
    sz[1] = morton_inc<idx_t>(sz[0], (idx_t)0x4924924924924924ULL);
    unsigned long long* const l0 = J.lvl[0];
    unsigned long long* const l1 = J.nl > 1 ? J.lvl[1] : nullptr;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int dx = q & 1, dy = (q >> 1) & 1, dz = q >> 2;
        unsigned long long sub = hits & (dx ? ~xlo : xlo) & (dy ? ~ylo : ylo) & (dz ? ~zlo : zlo);
        if (sub) {
            const idx_t wi = sx[dx] | sy[dy] | sz[dz];
            const int sh = (dx ? ox - 4 : ox) + 4 * (dy ? oy - 4 : oy) + 16 * (dz ? oz - 4 : oz);
            sub = sh >= 0 ? (sub << sh) : (sub >> (-sh));
            if (!OWNER) {
                red_or(l0 + wi, sub);
                if (l1) red_or(l1 + (wi >> 6), 1ULL << ((unsigned)wi & 63u));
            } else {
                while (sub) {
                    const int bit = __ffsll((long long)sub) - 1;
                    sub &= sub - 1;
                    sink_owner_bit(J, (uint64_t)wi, bit, tri);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Voxelizer, small bounding boxes: one lane per triangle/partition pair.
// OWNER = false: fill the bit-grid and route bigger pairs to the queues.
// OWNER = true : payload owner pass over the same pairs (queues already built).
// vox_small_body is the per-warp work; two kernels drive it:
//   k_vox_warp  (default) persistent warps, 32-triangle units, per-warp bulk-copy staging, ticket scheduling
//   k_vox_small block-granular: a block stages 128 records through shared memory with float4 loads (identity
//               lists), or reads the vertices of gathered per-partition lists directly
// ---------------------------------------------------------------------------
// Everything after the vertices are in registers: enumerate the triangle's partitions, classify the clamped
// boxes, route big ones to the queues, voxelize small ones. Must be called by whole warps.
template <bool OWNER, bool ENUM, bool BIG>
__device__ __forceinline__ void vox_small_body(const VoxJob& J, bool active, uint32_t tri, uint32_t part, const float* v) {
    // ---- the partitions of this triangle -------------------------------------------------------------
    // list mode: one (the pair's). inline mode: every logical partition whose world box the triangle's
    // bbox touches (inclusive float test, partitioner.cpp:117-126; the boxes are a product of per-axis
    // slabs, so the test separates per axis) -- for P == 1 the reference does not test at all (:80-98).
    int lx = 0, ly = 0, lz = 0, nx = 1, ny = 1, count = 0;
    constexpr bool enumerate = ENUM;                              // (J.pair_tri == nullptr) && J.P > 1
    if (active) {
        count = 1;
        if (enumerate) {
            int hx, hy, hz;
            slab_range(J.bmin, J.bmax, 1 << J.k, J.inv_slab, stdmin(v[0], stdmin(v[3], v[6])), stdmax(v[0], stdmax(v[3], v[6])), lx, hx);
            slab_range(J.bmin, J.bmax, 1 << J.k, J.inv_slab, stdmin(v[1], stdmin(v[4], v[7])), stdmax(v[1], stdmax(v[4], v[7])), ly, hy);
            slab_range(J.bmin, J.bmax, 1 << J.k, J.inv_slab, stdmin(v[2], stdmin(v[5], v[8])), stdmax(v[2], stdmax(v[5], v[8])), lz, hz);
            nx = max(hx - lx + 1, 0); ny = max(hy - ly + 1, 0);
            count = nx * ny * max(hz - lz + 1, 0);
        } else {
            lx = (int)compact3(part); ly = (int)compact3(part >> 1); lz = (int)compact3(part >> 2);   // folds to 0 for P == 1
        }
    }
    const int most = ENUM ? __reduce_max_sync(0xffffffffu, count) : 1;   // warp-uniform trip count (1 almost always)
    TriSetup s;
    if (count > 0) tri_setup(v, J.u, s, J.six != 0);
    for (int it = 0; it < most; it++) {
        const bool valid = it < count;
        // partition = slab coordinates (ix, iy, iz); no Morton round trip. Partitions of other ranks need no test of
        // their own: a rank's slab is a box (an aligned power-of-two Morton range of chunks), so restrict_to_slab
        // leaves nothing of a partition that lies outside it.
        int ix = lx, iy = ly, iz = lz;
        if (enumerate && it > 0 && valid) {     // rare: a triangle that straddles a partition boundary
            ix = lx + it % nx; iy = ly + (it / nx) % ny; iz = lz + it / (nx * ny);
        }
        int cls = -1;
        GridBox b = { 0, -1, 0, -1, 0, -1 };
        if (valid) {
            b = clamped_box(v, J.unit_div, ix * (int)J.side, iy * (int)J.side, iz * (int)J.side, (int)J.side);   // voxelizer.cpp:149: partition origin
            if (restrict_to_slab(J, b)) {
                const unsigned long long vol = (unsigned long long)(b.x1 - b.x0 + 1) * (unsigned long long)(b.y1 - b.y0 + 1) *
                                               (unsigned long long)(b.z1 - b.z0 + 1);
                // small = a few 4x4x4 windows anchored at the box corner
                const unsigned nw = (unsigned)((b.x1 - b.x0 + 4) >> 2) * (unsigned)((b.y1 - b.y0 + 4) >> 2) * (unsigned)((b.z1 - b.z0 + 4) >> 2);
                cls = (vol <= J.small_max && nw <= J.small_windows) ? 0 : (vol <= J.medium_max ? 1 : 2);
            }
        }
        if (!OWNER && __any_sync(0xffffffffu, cls > 0)) {       // rare: some lane holds a medium / large pair
            const unsigned long long e = ((unsigned long long)pack_slab(ix, iy, iz) << 32) | tri;     // queue entry: partition as slab coordinates
            warp_push(&J.qcount[0], J.queue[0], J.qcap, &J.qcount[3], cls == 1, e);
            warp_push(&J.qcount[1], J.queue[1], J.qcap, &J.qcount[3], cls == 2, e);
        }
        // warp-uniform upper bounds of the window extents (convergent point: every lane is here)
        const int ub = min(4, __reduce_max_sync(0xffffffffu, cls == 0 ? b.y1 - b.y0 + 1 : 0));
        const int uc = min(4, __reduce_max_sync(0xffffffffu, cls == 0 ? b.z1 - b.z0 + 1 : 0));
        // the (usually one) 4x4x4 windows of the box, anchored at its corner: ONE flat loop with a warp-uniform trip count
        const int nwx = cls == 0 ? (b.x1 - b.x0 + 4) >> 2 : 0, nwy = (b.y1 - b.y0 + 4) >> 2;
        const int mine = cls == 0 ? nwx * nwy * ((b.z1 - b.z0 + 4) >> 2) : 0;
        const int trips = __reduce_max_sync(0xffffffffu, mine);
        int cx = 0, cy = 0, cz = 0;
        for (int w = 0; w < trips; w++) {
            if (w < mine) {
                const int wx = b.x0 + 4 * cx, wy = b.y0 + 4 * cy, wz = b.z0 + 4 * cz;
                const unsigned long long hits = eval_window(s, J.u, wx, wy, wz, min(4, b.x1 - wx + 1), min(4, b.y1 - wy + 1),
                                                            min(4, b.z1 - wz + 1), ub, uc);
                if (hits) emit_window<OWNER, BIG>(J, wx, wy, wz, hits, tri);
                if (++cx == nwx) { cx = 0; if (++cy == nwy) { cy = 0; cz++; } }
            }
        }
    }
}

// Stage VOX_BLOCK consecutive triangle records starting at q0 of `tris` through shared memory (float4 loads).
__device__ __forceinline__ void stage_block(const float* tris, uint32_t fpt, uint64_t q0, uint64_t q_end, float4* s_stage4, float* s_stage) {
    const uint64_t nrec = (q_end - q0 < VOX_BLOCK) ? (q_end - q0) : VOX_BLOCK;
    const uint64_t nfl = nrec * fpt;
    const float* src = tris + q0 * fpt;                        // q0 is a multiple of VOX_BLOCK: 16-byte aligned
    const uint64_t n4 = nfl >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (uint64_t i = threadIdx.x; i < n4; i += VOX_BLOCK) s_stage4[i] = __ldg(src4 + i);
    for (uint64_t i = (n4 << 2) + threadIdx.x; i < nfl; i += VOX_BLOCK) s_stage[i] = __ldg(src + i);
    __syncthreads();
}

// Bulk asynchronous staging (cp.async.bulk, the TMA engine) for the persistent remote-staging kernel: one thread
// arms an mbarrier with the byte count and issues ONE copy of a whole staging block; the copy engine moves it into
// shared memory while the block voxelizes the previous one. Measured on 2 x B200 (tools/native/p2p_probe.cu): LDGSTS
// prefetch reaches the same NVLink bandwidth in isolation, but its long-lived remote requests sit in the SM's
// load/store path and slow the voxelizer's own atomics down; the bulk engine does not go through it.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(float* s_dst, const float* g_src, unsigned bytes, unsigned long long* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(s_dst)), "l"(g_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n.reg .pred p;\nMBAR_WAIT: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MBAR_DONE;\nbra MBAR_WAIT;\nMBAR_DONE:\n}"
                 :: "r"(b), "r"(parity) : "memory");
}

// SUBSET: 0 = every triangle of [q_begin, q_end), 1 = the staging blocks listed by k_owner_filter (sharded, replicated mesh)
template <bool OWNER, bool ENUM, int SUBSET>
__global__ void __launch_bounds__(VOX_BLOCK, SVO_VOX_MINBLOCKS) k_vox_small(VoxJob J) {
    extern __shared__ float4 s_stage4[];
    float* s_stage = reinterpret_cast<float*>(s_stage4);
    float v[9];
    if (SUBSET == 1) {
        // sharded: persistent blocks walk the list of staging blocks that touch this rank's slab (k_owner_filter)
        const unsigned long long count = *J.subset_count;
        for (unsigned long long li = blockIdx.x; li < count; li += gridDim.x) {
            const uint64_t q0 = (uint64_t)J.subset[li] * VOX_BLOCK;
            const uint64_t q = q0 + threadIdx.x;
            const bool active = q < J.q_end;
            stage_block(J.tris, J.fpt, q0, J.q_end, s_stage4, s_stage);
            if (active) {
#pragma unroll
                for (int i = 0; i < 9; i++) v[i] = s_stage[threadIdx.x * J.fpt + i];
            }
            __syncthreads();                                   // the next iteration overwrites the staging buffer
            vox_small_body<OWNER, ENUM, true>(J, active, (uint32_t)q, 0u, v);
        }
        return;
    }
    const uint64_t q0 = J.q_begin + (uint64_t)blockIdx.x * VOX_BLOCK;
    const uint64_t q = q0 + threadIdx.x;
    const bool active = q < J.q_end;
    uint32_t tri = 0, part = 0;
    if (J.pair_tri == nullptr) {
        stage_block(J.tris, J.fpt, q0, J.q_end, s_stage4, s_stage);
        tri = (uint32_t)q;
        if (active) {
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = s_stage[threadIdx.x * J.fpt + i];
        }
    } else if (active) {
        tri = J.pair_tri[q];
        part = pair_partition(J, q);
        load_vertices(J, nullptr, tri, v);
    }
    vox_small_body<OWNER, ENUM, true>(J, active, tri, part, v);
}

// ---------------------------------------------------------------------------
// Warp-persistent small-box voxelizer. The unit of work is 32 consecutive triangles (one per lane). Every WARP runs its
// own loop: take a ticket (atomic counter, fetched two units ahead), bulk-copy the unit's records into its private
// double buffer (cp.async.bulk + its own mbarrier), voxelize the previous unit meanwhile. No block-wide barrier in the
// loop: a warp with expensive triangles does not hold up the other three (with block-granular staging every iteration
// cost the maximum over the four warps), tickets balance the load across the whole grid, and units that lie in
// another rank's slab are simply never listed.
//   MODE 0: units are the consecutive 32-triangle runs of J.tris (one GPU, or a compact private copy)
//   MODE 2: units listed by the source ranks for this rank (k_slice_filter), staged from the owner's HBM over NVLink
// ---------------------------------------------------------------------------
constexpr int UNIT = 32;
template <bool OWNER, bool ENUM, int MODE, bool BIG>
__global__ void __launch_bounds__(VOX_BLOCK, SVO_VOX_MINBLOCKS) k_vox_warp(VoxJob J) {
    extern __shared__ float4 s_stage4[];
    __shared__ __align__(8) unsigned long long s_bar[VOX_BLOCK / 32][2];
    __shared__ unsigned long long s_base[MAX_WORLD + 1], s_pref[MAX_WORLD + 1];
    __shared__ int s_order[MAX_WORLD];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t unit_floats = UNIT * J.fpt;
    float* wbuf = reinterpret_cast<float*>(s_stage4) + (size_t)wid * 2 * unit_floats;
    unsigned long long count;
    if (MODE == 2) {
        // Sources are walked in ROTATED order. If every rank began with the same source, all of them would pull from the
        // same GPU at the same time (its NVLink egress, ~750 GB/s, shared by N readers) and then move on together. The
        // rotation is taken among the sources that actually listed units for this rank, and it starts at (rank mod their
        // number): with a lat-long mesh every octant needs the same four latitude bands, and a rotation that started at
        // the rank's own (unneeded) slice made three ranks begin at the same source -- measured on 8 GPUs: 1.7 ms for
        // those three against 1.2 ms for the rank that had a source to itself.
        segs_bases(J.segs, s_base);
        if (threadIdx.x == 0) {
            int ne = 0, idx[MAX_WORLD];
            for (int r = 0; r < J.segs.n; r++) if (J.pull_counts[(size_t)r * MAX_WORLD] > 0ULL) idx[ne++] = r;
            unsigned long long acc = 0;
            for (int i = 0; i < J.segs.n; i++) {
                int r = 0;
                unsigned long long cnt = 0ULL;
                if (i < ne) { r = idx[(J.pull_first + i) % ne]; cnt = J.pull_counts[(size_t)r * MAX_WORLD]; }
                s_order[i] = r; s_pref[i] = acc; acc += cnt;
            }
            s_pref[J.segs.n] = acc;
        }
        __syncthreads();
        count = s_pref[J.segs.n];
    } else {
        count = (J.q_end - J.q_begin + UNIT - 1) / UNIT;
    }
    if (lane == 0) { mbar_init(&s_bar[wid][0], 1); mbar_init(&s_bar[wid][1], 1); mbar_fence_init(); }
    __syncwarp();
    // Units are handed out by VOX_TICKETS counters, each on its own 128-byte line and each owning a contiguous 1/VOX_TICKETS of
    // the units; a warp draws from its home counter and, when that is used up, from the one with the most work left. (ONE
    // counter for the whole GPU was the kernel's real limit: returning atomics on a single address retire at ~370 M/s -- at
    // 32 triangles per unit exactly the 12 G triangles/s the kernel ran at from 1024^3 to 8192^3, with 18 % of all warp
    // stalls waiting for a ticket drawn a whole unit earlier.)
    constexpr unsigned long long NONE = ~0ULL;
    unsigned long long* const tk = J.qcount + VOX_TICKET_BASE + (OWNER ? VOX_TICKETS * VOX_TICKET_STRIDE : 0);
    auto lo_of = [&](int c) -> unsigned long long { return count * (unsigned long long)c / VOX_TICKETS; };
    int tc = (int)(((unsigned)blockIdx.x * (VOX_BLOCK / 32) + (unsigned)wid) % VOX_TICKETS);     // the counter this warp draws from
    auto take = [&]() -> unsigned long long {               // raw ticket of counter tc; lane 0 holds the value until it is broadcast
        return lane == 0 ? atomicAdd(tk + tc * VOX_TICKET_STRIDE, 1ULL) : 0ULL;
    };
    // the unit behind a raw ticket of counter c, or NONE when every counter is used up
    auto resolve = [&](unsigned long long raw_lane0, int c) -> unsigned long long {
        unsigned long long raw = __shfl_sync(0xffffffffu, raw_lane0, 0);
        unsigned long long lo = lo_of(c), size = lo_of(c + 1) - lo;
        if (raw < size) return lo + raw;
        for (int tries = 0; tries < 4 * VOX_TICKETS; tries++) {
            unsigned long long rem = 0ULL;
            if (lane < VOX_TICKETS) {
                const unsigned long long v = *(volatile const unsigned long long*)(tk + lane * VOX_TICKET_STRIDE);
                const unsigned long long sz = lo_of(lane + 1) - lo_of(lane);
                rem = v < sz ? sz - v : 0ULL;
            }
            int best = lane;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, rem, d);
                const int ob = __shfl_xor_sync(0xffffffffu, best, d);
                if (o > rem || (o == rem && ob < best)) { rem = o; best = ob; }
            }
            if (rem == 0ULL) return NONE;
            tc = best;
            raw = __shfl_sync(0xffffffffu, take(), 0);
            lo = lo_of(tc); size = lo_of(tc + 1) - lo;
            if (raw < size) return lo + raw;
        }
        return NONE;
    };
    // state of the unit that was located last
    int src = 0;
    const float* uptr = nullptr;                            // first record of the unit
    uint64_t ufirst = 0;                                    // global index of its first triangle
    uint32_t uvalid = 0;                                    // triangles in it (<= 32)
    auto locate = [&](unsigned long long u) {
        if (MODE == 2) {
            if (u < s_pref[src]) src = 0;                   // (a unit of another counter's range)
            while (u >= s_pref[src + 1]) src++;             // src: position in the rotated order
            const int r = s_order[src];                     // the source rank
            const uint64_t q0 = (uint64_t)J.subset[(size_t)r * J.pull_cap + (u - s_pref[src])] * UNIT;
            const uint64_t nseg = s_base[r + 1] - s_base[r];
            uptr = J.segs.ptr[r] + q0 * J.fpt;
            ufirst = s_base[r] + q0;
            uvalid = (uint32_t)(nseg - q0 < UNIT ? nseg - q0 : UNIT);
        } else {
            const uint64_t q0 = J.q_begin + u * UNIT;
            uptr = J.tris + q0 * J.fpt;
            ufirst = q0;
            uvalid = (uint32_t)(J.q_end - q0 < UNIT ? J.q_end - q0 : UNIT);
        }
    };
    // A full unit is staged by the bulk-copy engine; the (single) ragged last unit of an array is read directly, so that
    // nothing is ever read past the end of a caller's buffer. Slices are padded, but the rule is the same for both modes.
    auto stage = [&](int b) {
        if (uvalid == UNIT && lane == 0) bulk_load(wbuf + (size_t)b * unit_floats, uptr, unit_floats * 4u, &s_bar[wid][b]);
    };
    const int c0 = tc;
    const unsigned long long r0 = take(), r1 = take();
    unsigned long long cur = resolve(r0, c0);
    unsigned long long next = cur != NONE ? resolve(r1, c0) : NONE;
    uint32_t phase = 0;
    if (cur != NONE) { locate(cur); stage(0); }
    float v[9];
    for (int it = 0; cur != NONE; it++) {
        const int b = it & 1;
        const float* my_ptr = uptr;
        const uint32_t my_valid = uvalid;
        const uint32_t tri = (uint32_t)(ufirst + lane);
        const int ca = tc;
        const unsigned long long ahead = take();            // ticket of iteration it + 2; consumed after the body
        if (next != NONE) { locate(next); stage(b ^ 1); }   // buffer b^1 was last read in iteration it - 1 (before its __syncwarp)
        const bool active = (uint32_t)lane < my_valid;
        if (my_valid == UNIT) {
            mbar_wait(&s_bar[wid][b], (phase >> b) & 1u);
            phase ^= 1u << b;
            const float* cbuf = wbuf + (size_t)b * unit_floats;
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = cbuf[lane * J.fpt + i];
        } else if (active) {
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = __ldg(my_ptr + (size_t)lane * J.fpt + i);
        }
        __syncwarp();                                       // every lane has its vertices: the buffer may be refilled
        vox_small_body<OWNER, ENUM, BIG>(J, active, tri, 0u, v);
        cur = next;
        next = cur != NONE ? resolve(ahead, ca) : NONE;
    }
}

// ---------------------------------------------------------------------------
// Voxelizer, medium and large boxes: warps walk the bricks of the box. 32 bricks
// are pruned at a time (one per lane, exact box test), the survivors are then
// evaluated by the whole warp, 64 voxels = 2 per lane, ballots build the word.
// Medium: one warp owns a pair. Large: all warps of the grid share each pair.
// ---------------------------------------------------------------------------
template <bool OWNER>
__device__ __forceinline__ void warp_voxelize_box(const VoxJob& J, const TriSetup& s, const GridBox& b, uint32_t tri,
                                                  unsigned long long chunk0, unsigned long long chunk_stride) {
    const int lane = threadIdx.x & 31;
    const int bx0 = b.x0 >> 2, by0 = b.y0 >> 2, bz0 = b.z0 >> 2;
    const unsigned long long nbx = (unsigned long long)((b.x1 >> 2) - bx0 + 1);
    const unsigned long long nby = (unsigned long long)((b.y1 >> 2) - by0 + 1);
    const unsigned long long nbz = (unsigned long long)((b.z1 >> 2) - bz0 + 1);
    const unsigned long long nb = nbx * nby * nbz;
    const int dx = lane & 3;                                  // lane = LINEAR in-brick bit index z*16 + y*4 + x (lower half: z = 0, 1)
    const int dy = (lane >> 2) & 3;
    const int dz = lane >> 4;
    const float u = J.u;
    for (unsigned long long c = chunk0; c * 32ULL < nb; c += chunk_stride) {
        const unsigned long long bi = c * 32ULL + lane;
        int bx = 0, by = 0, bz = 0;
        bool may = false;
        if (bi < nb) {
            bx = bx0 + (int)(bi % nbx);
            const unsigned long long r = bi / nbx;
            by = by0 + (int)(r % nby);
            bz = bz0 + (int)(r / nby);
            may = box_may_pass(s, u, max(b.x0, bx << 2), min(b.x1, (bx << 2) + 3), max(b.y0, by << 2),
                               min(b.y1, (by << 2) + 3), max(b.z0, bz << 2), min(b.z1, (bz << 2) + 3));
        }
        unsigned m = __ballot_sync(0xffffffffu, may);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int sbx = __shfl_sync(0xffffffffu, bx, src);
            const int sby = __shfl_sync(0xffffffffu, by, src);
            const int sbz = __shfl_sync(0xffffffffu, bz, src);
            const int x = (sbx << 2) + dx, y = (sby << 2) + dy, z = (sbz << 2) + dz;
            bool h0 = false, h1 = false;
            if (x >= b.x0 && x <= b.x1 && y >= b.y0 && y <= b.y1) {
                const float px = fmul((float)x, u), py = fmul((float)y, u);
                if (edge_pass(s, 0, px, py) && edge_pass(s, 1, px, py) && edge_pass(s, 2, px, py)) {
                    if (z >= b.z0 && z <= b.z1) {
                        const float pz = fmul((float)z, u);
                        h0 = plane_pass(s, px, py, pz) && edge_pass(s, 3, py, pz) && edge_pass(s, 4, py, pz) &&
                             edge_pass(s, 5, py, pz) && edge_pass(s, 6, pz, px) && edge_pass(s, 7, pz, px) &&
                             edge_pass(s, 8, pz, px);
                    }
                    if (z + 2 >= b.z0 && z + 2 <= b.z1) {
                        const float pz = fmul((float)(z + 2), u);
                        h1 = plane_pass(s, px, py, pz) && edge_pass(s, 3, py, pz) && edge_pass(s, 4, py, pz) &&
                             edge_pass(s, 5, py, pz) && edge_pass(s, 6, pz, px) && edge_pass(s, 7, pz, px) &&
                             edge_pass(s, 8, pz, px);
                    }
                }
            }
            const unsigned lo = __ballot_sync(0xffffffffu, h0);
            const unsigned hi = __ballot_sync(0xffffffffu, h1);
            if (lo | hi) {
                const uint64_t w = morton3((uint32_t)sbx, (uint32_t)sby, (uint32_t)sbz);
                if (!OWNER) {
                    if (lane == 0) sink_fill(J, w, ((unsigned long long)hi << 32) | lo);
                } else {
                    if (h0) sink_owner_bit(J, w, lane, tri);
                    if (h1) sink_owner_bit(J, w, lane + 32, tri);
                }
            }
        }
    }
}

__device__ __forceinline__ void queued_pair_setup(const VoxJob& J, const unsigned long long* s_base, unsigned long long e, uint32_t& tri, TriSetup& s, GridBox& b) {
    tri = (uint32_t)(e & 0xffffffffULL);
    const uint32_t slab = (uint32_t)(e >> 32);                  // pack_slab(ix, iy, iz)
    float v[9];
    load_vertices(J, s_base, tri, v);
    const int px = (int)((slab & 0xffu) * J.side), py = (int)(((slab >> 8) & 0xffu) * J.side), pz = (int)(((slab >> 16) & 0xffu) * J.side);
    b = clamped_box(v, J.unit_div, px, py, pz, (int)J.side);
    restrict_to_slab(J, b);     // queued pairs are non-empty by construction
    tri_setup(v, J.u, s, J.six != 0);
}

template <bool OWNER>
__device__ __forceinline__ void vox_medium_body(const VoxJob& J, const unsigned long long* s_base) {
    const unsigned long long n = min(J.qcount[0], J.qcap);
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    for (unsigned long long e = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); e < n; e += nwarps) {
        uint32_t tri; TriSetup s; GridBox b;
        queued_pair_setup(J, s_base, J.queue[0][e], tri, s, b);
        warp_voxelize_box<OWNER>(J, s, b, tri, 0ULL, 1ULL);
    }
}

template <bool OWNER>
__device__ __forceinline__ void vox_large_body(const VoxJob& J, const unsigned long long* s_base) {
    const unsigned long long n = min(J.qcount[1], J.qcap);
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    const unsigned long long gw = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    for (unsigned long long e = 0; e < n; e++) {
        uint32_t tri; TriSetup s; GridBox b;
        queued_pair_setup(J, s_base, J.queue[1][e], tri, s, b);
        warp_voxelize_box<OWNER>(J, s, b, tri, gw, nwarps);
    }
}

// one launch for both queues (they are usually short or empty)
template <bool OWNER>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_vox_queued(VoxJob J) {
    __shared__ unsigned long long s_base[MAX_WORLD + 1];
    if (J.segs.n) segs_bases(J.segs, s_base);
    vox_medium_body<OWNER>(J, s_base);
    vox_large_body<OWNER>(J, s_base);
}

// ---------------------------------------------------------------------------
// Partitioner: bin triangles into the logical partitions whose world box their
// bbox touches (partitioner.cpp:43-61, :117-126; BBoxBuffer.h:70-84;
// intersection.h:50-53). The boxes are a product of per-axis slabs, so the
// inclusive test separates per axis into a contiguous slab range.
// ---------------------------------------------------------------------------
struct BinJob {
    const float* tris;
    uint32_t fpt;
    uint64_t n_tris;
    uint32_t k, P;
    float bmin[32], bmax[32];         // world slab [bmin[i], bmax[i]] of slab i (same on every axis)
    float inv_slab;
    unsigned long long* counts;       // P
    unsigned long long* cursor;       // P (fill pass)
    const unsigned long long* off;    // P+1 (fill pass)
    uint32_t* pair_tri;
};

__device__ __forceinline__ void slab_range(const BinJob& B, float mn, float mx, int& lo, int& hi) {
    slab_range(B.bmin, B.bmax, 1 << B.k, B.inv_slab, mn, mx, lo, hi);
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_bin(BinJob B) {
    extern __shared__ unsigned int s_hist[];
    const bool use_smem = !FILL && B.P <= 4096;
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < B.P; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lx = 1, hx = 0, ly = 1, hy = 0, lz = 1, hz = 0;
    if (t < B.n_tris) {
        const float* v = B.tris + t * B.fpt;
        float c[9];
#pragma unroll
        for (int i = 0; i < 9; i++) c[i] = __ldg(v + i);
        slab_range(B, stdmin(c[0], stdmin(c[3], c[6])), stdmax(c[0], stdmax(c[3], c[6])), lx, hx);
        slab_range(B, stdmin(c[1], stdmin(c[4], c[7])), stdmax(c[1], stdmax(c[4], c[7])), ly, hy);
        slab_range(B, stdmin(c[2], stdmin(c[5], c[8])), stdmax(c[2], stdmax(c[5], c[8])), lz, hz);
    }
    const int nx = max(hx - lx + 1, 0), ny = max(hy - ly + 1, 0), nz = max(hz - lz + 1, 0);
    const int mine = nx * ny * nz;
    if (!FILL) {
        for (int i = 0; i < mine; i++) {
            const uint32_t part = (uint32_t)morton3(lx + i % nx, ly + (i / nx) % ny, lz + i / (nx * ny));
            if (use_smem) atomicAdd(&s_hist[part], 1u); else atomicAdd(&B.counts[part], 1ULL);
        }
        if (use_smem) {
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < B.P; i += blockDim.x)
                if (s_hist[i]) atomicAdd(&B.counts[i], (unsigned long long)s_hist[i]);
        }
    } else {
        // warp-uniform trip count so that the match/ballot aggregation is convergent
        const int most = __reduce_max_sync(0xffffffffu, mine);
        const int lane = threadIdx.x & 31;
        for (int i = 0; i < most; i++) {
            const bool valid = i < mine;
            uint32_t part = 0xffffffffu;
            if (valid) part = (uint32_t)morton3(lx + i % nx, ly + (i / nx) % ny, lz + i / (nx * ny));
            const unsigned peers = __match_any_sync(0xffffffffu, part);
            const int leader = __ffs(peers) - 1;
            unsigned long long base = 0;
            if (valid && lane == leader) base = atomicAdd(&B.cursor[part], (unsigned long long)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) B.pair_tri[B.off[part] + base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)t;
        }
    }
}

// Sharded runs: every rank holds all triangles but only voxelizes the ones that touch its slab. This light
// pass (36 B read per triangle) lists the staging blocks (VOX_BLOCK consecutive triangles) that contain at
// least one such triangle -- meshes are spatially coherent, so most blocks are all-or-nothing; the test is a
// superset test on bounding boxes (the voxelizer still checks every partition / word exactly). partition mode: per-axis slab ranges against the
// bounding box of the owned partitions; P == 1: clamped grid bbox against the slab's voxel box.
struct FilterJob {
    const float* tris; uint32_t fpt; unsigned long long n_tris;
    int use_partitions, k;
    float bmin[32], bmax[32];
    int lo[3], hi[3];                 // owned box: partition coordinates (use_partitions) or voxels
    float unit_div; int gmax;
    uint32_t* out; unsigned long long* count;
};
__global__ void __launch_bounds__(VOX_BLOCK) k_owner_filter(FilterJob Fj) {
    // one thread block = one staging block of VOX_BLOCK consecutive triangles of the voxelizer
    const unsigned long long t = (unsigned long long)blockIdx.x * VOX_BLOCK + threadIdx.x;
    bool keep = false;
    if (t < Fj.n_tris) {
        const float* v = Fj.tris + t * Fj.fpt;
        float c[9];
#pragma unroll
        for (int i = 0; i < 9; i++) c[i] = __ldg(v + i);
        keep = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float mn = stdmin(c[a], stdmin(c[3 + a], c[6 + a])), mx = stdmax(c[a], stdmax(c[3 + a], c[6 + a]));
            if (Fj.use_partitions) {
                // the kept slab range [L, H] meets the owned slabs [lo, hi] iff H >= lo and L <= hi; by monotonicity of
                // the slab tables that is (mx >= bmin[lo]) and (mn <= bmax[hi]) -- two compares instead of a slab search
                keep = keep && !(mx < Fj.bmin[Fj.lo[a]]) && !(mn > Fj.bmax[Fj.hi[a]]);
            } else {
                const int l = clampi(f2i(fmul(mn, Fj.unit_div)), 0, Fj.gmax), h = clampi(f2i(fmul(mx, Fj.unit_div)), 0, Fj.gmax);
                keep = keep && !(h < Fj.lo[a] || l > Fj.hi[a]);
            }
        }
    }
    const int any = __syncthreads_or(keep ? 1 : 0);
    if (any && threadIdx.x == 0) Fj.out[atomicAdd(Fj.count, 1ULL)] = blockIdx.x;
}

// ---------------------------------------------------------------------------
// Hand-written exclusive scans (no CUB). out has n + 1 entries (out[n] = total). Small inputs: one block
// (k_scan_small); everything else: one single-pass look-back launch (k_scan_lookback).
// ---------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 8;

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
// exclusive scan across the block; returns this thread's exclusive prefix, total via ref
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long& total) {
    __shared__ unsigned long long s_w[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const unsigned long long inc = warp_incl_scan(v);
    __syncthreads();                 // protect s_w across back-to-back calls
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long x = lane < nw ? s_w[lane] : 0ULL;
        x = warp_incl_scan(x);
        s_w[lane] = x;
    }
    __syncthreads();
    total = s_w[nw - 1];
    return inc - v + (wid ? s_w[wid - 1] : 0ULL);
}

// ---------------------------------------------------------------------------
// Device-resident result block of a build: tile counts of the local levels, global counts, this rank's file range.
// Filled by the build kernels, read back once with the final synchronisation; kernels launched with grids sized
// from CAPACITIES (remembered from the previous build) read the real counts from here, so that a steady-state build
// needs no host read-back in the middle (OctreeBuilder.cpp:34-55: the counts are only needed at finalizeTree).
// ---------------------------------------------------------------------------
struct BuildInfo {
    unsigned long long count[MAX_LEVELS];     // tiles (non-zero words) of the local levels 0..J
    unsigned long long n_leaves_local;        // voxels in this rank's slab
    unsigned long long n_brick_records;       // records of this rank's brick subtrees: leaves + depth D-1 nodes (what k_emit_leaf writes)
    unsigned long long n_voxels, n_nodes;     // global
    unsigned long long leaf_offset;           // voxels in the slabs of lower ranks
    unsigned long long node_lo, node_hi;      // this rank's records of the node file
    unsigned long long n_upper;               // shared upper-level records inside [node_lo, node_hi)
    unsigned long long leaf_ticket;           // k_emit_leaf hands out its batches in file order
    unsigned long long overflow;              // bit j: list of level j too small; bit 32: node buffer; bit 40: look-back timeout.
                                              // Once set, every later kernel of the build returns at once (the pyramid stays intact).
};
__device__ __forceinline__ bool build_aborted(const BuildInfo* info) { return info && *(volatile const unsigned long long*)&info->overflow != 0ULL; }

// ---------------------------------------------------------------------------
// Single-pass exclusive scan (decoupled look-back) of NV values per element: every element's functor is evaluated
// once, one launch. Tiles are taken in ticket order (forward progress: every predecessor of a tile is running or
// done). A tile publishes its aggregate, later its inclusive prefix, in an 8-word state record
//   [0] flag = epoch << 2 | status (1 = aggregate valid, 2 = inclusive prefix valid)   [1..NV] aggregate   [1+NV..2NV] inclusive
// Values are written once per epoch and BEFORE the flag that announces them (fence in between), so ONE look-back
// chain serves all NV values. The epoch makes the persistent state array reusable without clearing it.
// ---------------------------------------------------------------------------
constexpr int LB_THREADS = 256, LB_ITEMS = 8, LB_TILE = LB_THREADS * LB_ITEMS;
constexpr int LB_STATE = 8;                   // u64 per tile
constexpr int LB_MAXV = 3;
// Called by all 32 lanes of ONE warp. total[] = this tile's aggregate; returns the exclusive prefix of the tile in
// prefix[] (every lane) and publishes the inclusive prefix. A wait that never ends (cannot happen with ticket order;
// a guard against hanging the device) sets *err.
template <int NV>
__device__ __forceinline__ void lookback(unsigned long long* state, unsigned long long epoch, unsigned long long tile,
                                         const unsigned long long (&total)[NV], unsigned long long (&prefix)[NV], unsigned long long* err) {
    static_assert(1 + 2 * NV <= LB_STATE, "state record too small");
    volatile unsigned long long* st = state;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < NV; c++) prefix[c] = 0ULL;
    if (tile > 0) {
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < NV; c++) st[tile * LB_STATE + 1 + c] = total[c];
            __threadfence();
            st[tile * LB_STATE] = (epoch << 2) | 1ULL;
        }
        long long idx = (long long)tile - 1;
        unsigned spins = 0;
        while (true) {
            // lane l looks at tile idx - l; tiles before the first count as "inclusive prefix 0"
            const long long p = idx - lane;
            const unsigned long long f = p >= 0 ? st[p * LB_STATE] : ((epoch << 2) | 2ULL);
            const unsigned status = (unsigned)(f & 3ULL);
            const bool ok = (f >> 2) == epoch && status != 0u;
            const unsigned m_incl = __ballot_sync(0xffffffffu, ok && status == 2u);
            const unsigned m_bad = __ballot_sync(0xffffffffu, !ok);
            const int first = m_incl ? __ffs(m_incl) - 1 : 32;
            const unsigned need = first >= 31 ? 0xffffffffu : ((1u << (first + 1)) - 1u);
            if (m_bad & need) {                      // a predecessor has not published yet: look again
                if (++spins > (1u << 22)) { if (lane == 0 && err) atomicOr(err, 1ULL << 40); break; }
                __nanosleep(20);
                continue;
            }
            __threadfence();                         // the values were written before the flag
            const bool take = lane <= first && p >= 0;
#pragma unroll
            for (int c = 0; c < NV; c++) {
                unsigned long long x = take ? st[p * LB_STATE + 1 + (status == 2u ? NV : 0) + c] : 0ULL;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                prefix[c] += x;
            }
            if (first < 32) break;
            idx -= 32;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NV; c++) st[tile * LB_STATE + 1 + NV + c] = prefix[c] + total[c];
        __threadfence();
        st[tile * LB_STATE] = (epoch << 2) | 2ULL;
    }
}

// Generic exclusive scan of f over [0, n): out has n + 1 entries (out[n] = total). n comes from the host (np == NULL)
// or from device memory (np != NULL: the launch grid was sized for `n` = the capacity, the real count is min(*np, n)).
template <int NV, class F>
__global__ void __launch_bounds__(LB_THREADS) k_scan_lookback(F f, unsigned long long n, const unsigned long long* np,
                                                              unsigned long long* out0, unsigned long long* out1, unsigned long long* out2,
                                                              unsigned long long* state, unsigned long long* ticket,
                                                              unsigned long long ticket_base, unsigned long long epoch, BuildInfo* info) {
    __shared__ unsigned long long s_tile, s_prefix[NV];
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1ULL) - ticket_base;
    __syncthreads();
    if (build_aborted(info)) return;
    if (np) { const unsigned long long v = *np; n = v < n ? v : n; }
    const unsigned long long tile = s_tile;
    if (tile * LB_TILE >= n) {
        if (tile == 0 && threadIdx.x == 0) { out0[0] = 0ULL; if (NV > 1) out1[0] = 0ULL; if (NV > 2) out2[0] = 0ULL; }      // empty input: out[n] = out[0] = 0
        return;
    }
    const unsigned long long base = tile * LB_TILE + (unsigned long long)threadIdx.x * LB_ITEMS;
    unsigned long long v[NV][LB_ITEMS], acc[NV], ex[NV], total[NV];
#pragma unroll
    for (int c = 0; c < NV; c++) acc[c] = 0;
#pragma unroll
    for (int i = 0; i < LB_ITEMS; i++) {
        unsigned long long e[NV];
#pragma unroll
        for (int c = 0; c < NV; c++) e[c] = 0;
        if (base + i < n) f(base + i, e);
#pragma unroll
        for (int c = 0; c < NV; c++) { v[c][i] = e[c]; acc[c] += e[c]; }
    }
#pragma unroll
    for (int c = 0; c < NV; c++) ex[c] = block_excl_scan(acc[c], total[c]);
    if (threadIdx.x < 32) {
        unsigned long long pre[NV];
        lookback<NV>(state, epoch, tile, total, pre, info ? &info->overflow : nullptr);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int c = 0; c < NV; c++) s_prefix[c] = pre[c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NV; c++) {
        unsigned long long* out = c == 0 ? out0 : (c == 1 ? out1 : out2);
        unsigned long long run = s_prefix[c] + ex[c];
#pragma unroll
        for (int i = 0; i < LB_ITEMS; i++) {
            if (base + i < n) out[base + i] = run;
            run += v[c][i];
        }
        if (threadIdx.x == 0 && (tile + 1) * LB_TILE >= n) out[n] = s_prefix[c] + total[c];     // the last tile writes the total
    }
}
template <class F>
struct OneValue {        // adapts a functor `ull f(i)` to the NV = 1 interface
    F f;
    __device__ void operator()(unsigned long long i, unsigned long long (&e)[1]) const { e[0] = f(i); }
};
struct BrickPrefixes {   // level 0: leaf ranks (popcount) and subtree sizes popc(W) + popc8(W) of the same words
    const unsigned long long* mask;
    __device__ void operator()(unsigned long long i, unsigned long long (&e)[2]) const {
        const unsigned long long w = mask[i];
        const unsigned pc = __popcll(w);
        e[0] = pc; e[1] = pc + __popc(nonzero_bytes(w));
    }
};

// Small inputs (the upper pyramid levels): the whole exclusive scan in ONE block / one launch.
template <class F>
__global__ void __launch_bounds__(1024) k_scan_small(F f, unsigned long long n, const unsigned long long* np, unsigned long long* out, const BuildInfo* info) {
    if (build_aborted(info)) return;
    if (np) { const unsigned long long v = *np; n = v < n ? v : n; }
    // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements of each 8192-element chunk
    unsigned long long carry = 0;
    for (unsigned long long b = 0; b < n; b += (unsigned long long)blockDim.x * SCAN_ITEMS) {
        const unsigned long long base = b + (unsigned long long)threadIdx.x * SCAN_ITEMS;
        unsigned long long v[SCAN_ITEMS];
        unsigned long long acc = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            v[i] = (base + i < n) ? f(base + i) : 0ULL;
            acc += v[i];
        }
        unsigned long long total;
        unsigned long long run = carry + block_excl_scan(acc, total);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            if (base + i < n) out[base + i] = run;
            run += v[i];
        }
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// ---------------------------------------------------------------------------
// Pyramid compaction + octree construction
// ---------------------------------------------------------------------------
struct Level {
    unsigned long long* key;      // n      word index at this level (= Morton prefix of the node)
    unsigned long long* mask;     // n      the 64-bit word
    unsigned long long* fc;       // n + 1  exclusive prefix of popc(mask): index of first child tile / leaf rank
    unsigned long long* ps;       // n + 1  exclusive prefix of subtree sizes S
    unsigned long long* base;     // n      file position of the first record of the tile's subtree region
    // -levels only:
    unsigned long long* pi;       // n + 1  exclusive prefix of I = internal (non-leaf) nodes in the tile's subtree, itself included
    unsigned long long* pl;       // n + 1  exclusive prefix of the leaf counts
    unsigned long long* ibase;    // n      internal nodes completed (post-order) before the tile's subtree begins
    float* cache;                 // n * 6  averaged colour + normal of the tile's node (Node::data_cache)
    unsigned long long n;         // tile count known on the host (np == NULL) ...
    const unsigned long long* np; // ... or device-resident (BuildInfo::count[j]); then the lists hold `cap` entries
    unsigned long long cap;
    unsigned long long* clear;    // fast path: the level's dense words (pre-biased); the emitter of the level zeroes the word of
                                  // every tile it handles, which leaves a clean pyramid without a separate clearing pass
};
__device__ __forceinline__ unsigned long long level_n(const Level& L) {
    if (!L.np) return L.n;
    const unsigned long long v = *L.np;
    return v < L.cap ? v : L.cap;
}

// counts[j] = number of non-zero words of local dense level j, j = 0..J: for j < J that is the number
// of set bits of level j+1; for the top local level J the words are counted directly.
__global__ void __launch_bounds__(256) k_level_counts(unsigned long long* const* lvl, const unsigned long long* nwords, int J,
                                                      unsigned long long* counts) {
    const int j = blockIdx.y;
    if (j > J) return;
    const int src = j < J ? j + 1 : J;
    const unsigned long long n = nwords[src];
    unsigned long long acc = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long w = lvl[src][i];
        acc += j < J ? (unsigned long long)__popcll(w) : (w != 0ULL ? 1ULL : 0ULL);
    }
    unsigned long long total;
    block_excl_scan(acc, total);
    if (threadIdx.x == 0 && total) atomicAdd(&counts[j], total);
}

// Top local level: compact the non-zero words of the (small) dense level J into (key, mask). One block.
__global__ void __launch_bounds__(1024) k_compact_top(const unsigned long long* dense, unsigned long long n, unsigned long long key_bias,
                                                      unsigned long long* key, unsigned long long* mask, uint32_t* tileidx, int is_level0) {
    unsigned long long carry = 0;
    for (unsigned long long b = 0; b < n; b += blockDim.x) {
        const unsigned long long idx = b + threadIdx.x;
        const unsigned long long w = idx < n ? dense[idx] : 0ULL;
        unsigned long long total;
        const unsigned long long ex = block_excl_scan(w != 0ULL ? 1ULL : 0ULL, total);
        if (w != 0ULL) {
            key[carry + ex] = key_bias + idx; mask[carry + ex] = is_level0 ? linear_to_morton64(w) : w;
            if (tileidx) tileidx[idx] = (uint32_t)(carry + ex);     // level 0 only (payload owner pass)
        }
        carry += total;
    }
}

// The exchange table: one entry of 4 u64 per GLOBAL level-J word: {mask, subtree size S, leaves, internal nodes}.
// Every rank writes the entries of its own words (all others stay zero), the caller sums the tables of
// all ranks (NCCL all-reduce; the entries are disjoint, so the sum is the union).
struct TableFillJob {
    const unsigned long long* key; const unsigned long long* mask; const unsigned long long* ps;
    const unsigned long long* pi;                  // -levels only
    const unsigned long long* fc[MAX_LEVELS];      // child-prefix arrays of levels 0..J: chasing them gives leaf ranks
    unsigned long long n; int J;
    const unsigned long long* np;                  // device-resident count (fast path) or NULL
    unsigned long long* table;
    BuildInfo* info;
    int stride;                                    // u64 per entry: 4, or 8 with -levels (entry[4..6] = the tile's 6-float data cache)
    const float* cache;                            // -levels: Node::data_cache of the level-J tiles (k_levels_data)
};
// first leaf rank below tile i of level J: follow the first-child links down to level 0
__device__ __forceinline__ unsigned long long first_leaf_below(const TableFillJob& T, unsigned long long i) {
    for (int j = T.J; j >= 0; j--) i = T.fc[j][i];
    return i;
}
__device__ __forceinline__ void table_fill_entry(const TableFillJob& T, unsigned long long i) {
    unsigned long long* e = T.table + T.key[i] * (unsigned long long)(T.stride ? T.stride : 4);
    e[0] = T.mask[i];
    e[1] = T.ps[i + 1] - T.ps[i];
    e[2] = first_leaf_below(T, i + 1) - first_leaf_below(T, i);     // fc[j][n_j] = n_{j-1}: the chain is valid for i = n too
    e[3] = T.pi ? T.pi[i + 1] - T.pi[i] : 0ULL;
    if (T.stride == 8 && T.cache) {
        const uint32_t* cb = reinterpret_cast<const uint32_t*>(T.cache + i * 6);
        e[4] = (unsigned long long)cb[0] | ((unsigned long long)cb[1] << 32);
        e[5] = (unsigned long long)cb[2] | ((unsigned long long)cb[3] << 32);
        e[6] = (unsigned long long)cb[4] | ((unsigned long long)cb[5] << 32);
        e[7] = 0ULL;
    }
}
__global__ void __launch_bounds__(256) k_table_fill(TableFillJob T) {
    if (build_aborted(T.info)) return;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n = T.n;
    if (T.np) { const unsigned long long v = *T.np; n = v < n ? v : n; }
    if (i >= n) return;
    table_fill_entry(T, i);
}
// the same by ONE block (fused top kernel): zeroes the table first; `zero_words` u64
__device__ __forceinline__ void table_fill_body(const TableFillJob& T, unsigned long long zero_words) {
    for (unsigned long long i = threadIdx.x; i < zero_words; i += blockDim.x) T.table[i] = 0ULL;
    __syncthreads();
    unsigned long long n = T.n;
    if (T.np) { const unsigned long long v = *T.np; n = v < n ? v : n; }
    for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) table_fill_entry(T, i);
    __threadfence();
    __syncthreads();
}
// Unpack the summed table into dense columns and rebuild the (tiny, replicated) upper dense levels.
__global__ void __launch_bounds__(256) k_table_unpack(const unsigned long long* table, unsigned long long n, unsigned long long* dmask,
                                                      unsigned long long* dS, unsigned long long* dLC, unsigned long long* dI,
                                                      unsigned long long* const* lvl, int first_upper, int nl) {
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long m = table[e * 4ULL];
    dmask[e] = m; dS[e] = table[e * 4ULL + 1]; dLC[e] = table[e * 4ULL + 2]; dI[e] = table[e * 4ULL + 3];
    if (m) {
        unsigned long long w = e;
        for (int j = first_upper; j < nl; j++) {
            const unsigned long long bit = 1ULL << (w & 63);
            w >>= 6;
            if (atomicOr(&lvl[j][w], bit) != 0ULL) break;
        }
    }
}
struct DenseColOp {      // value of a dense per-word column at the key of compact tile i
    const unsigned long long* col;
    const unsigned long long* key;
    __device__ unsigned long long operator()(unsigned long long i) const { return col[key[i]]; }
};

struct CountOp {
    const unsigned long long* cnt;
    __device__ unsigned long long operator()(unsigned long long i) const { return cnt[i]; }
};
struct PopcOp {
    const unsigned long long* mask;
    __device__ unsigned long long operator()(unsigned long long i) const { return (unsigned long long)__popcll(mask[i]); }
};
// subtree size of tile i: S = popc(W) + popc8(W) + sum of the children's S
struct SizeOp {
    const unsigned long long* mask;
    const unsigned long long* fc;        // this level
    const unsigned long long* child_ps;  // level below (NULL at level 0)
    __device__ unsigned long long operator()(unsigned long long i) const {
        const unsigned long long w = mask[i];
        unsigned long long s = (unsigned long long)(__popcll(w) + __popc(nonzero_bytes(w)));
        if (child_ps) s += child_ps[fc[i + 1]] - child_ps[fc[i]];
        return s;
    }
};

// -levels: leaves / internal nodes below (and including) tile i
struct LeafCountOp {
    const unsigned long long* mask;
    const unsigned long long* fc;
    const unsigned long long* child_pl;  // NULL at level 0
    __device__ unsigned long long operator()(unsigned long long i) const {
        if (!child_pl) return (unsigned long long)__popcll(mask[i]);
        return child_pl[fc[i + 1]] - child_pl[fc[i]];
    }
};
struct InternalOp {
    const unsigned long long* mask;
    const unsigned long long* fc;
    const unsigned long long* child_pi;  // NULL at level 0
    __device__ unsigned long long operator()(unsigned long long i) const {
        unsigned long long s = 1ULL + __popc(nonzero_bytes(mask[i]));
        if (child_pi) s += child_pi[fc[i + 1]] - child_pi[fc[i]];
        return s;
    }
};

// Top-down expansion: one warp per parent tile writes its children's keys and
// gathers their words from the dense level below.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_expand(Level parent, Level child, const unsigned long long* dense_child,
                                                                uint32_t* tileidx, int child_is_level0) {
    const unsigned long long i = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= parent.n) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long W = parent.mask[i], key = parent.key[i], fc = parent.fc[i];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bit = lane + 32 * h;
        if ((W >> bit) & 1ULL) {
            const unsigned long long c = fc + __popcll(W & lowmask(bit));
            const unsigned long long ck = (key << 6) | (unsigned long long)bit;
            child.key[c] = ck;
            const unsigned long long cw = dense_child[ck];
            child.mask[c] = child_is_level0 ? linear_to_morton64(cw) : cw;     // bricks are voxelized in the linear layout
            if (tileidx) tileidx[ck] = (uint32_t)c;
        }
    }
}

struct EmitJob {
    unsigned long long* nodes;        // this rank's node records: record `pos` of the file lives at nodes + (pos - node_lo) * 3
    int is_top;                       // this level holds the single top word
    int root_here;                    // D even: the top word IS the root -> write its record at S(top)
    int leaf_data_mode;               // level 0: 0 = binary (data = 1), 1 = payload (data = 1 + leaf rank)
    int levels;                       // -levels: internal nodes carry a data index too
    int virtual_top;                  // D odd: the top word is not a node (its byte 0 is the root)
    // this rank's range [pos_lo, pos_hi) of the file and the leaves in the slabs of lower ranks: by value (info == NULL)
    // or device-resident (fast path: BuildInfo::node_lo / node_hi / leaf_offset)
    unsigned long long pos_lo, pos_hi;
    unsigned long long leaf_offset;
    unsigned long long cap;           // records the node buffer holds (speculative emission into an earlier build's buffer)
    const BuildInfo* info;
    int write_records;                // 0: only propagate the subtree bases
    unsigned long long* ticket;       // k_emit_leaf: batches in ticket order (a counter that is zero at launch), or NULL: round-robin
};
struct NodeRange { unsigned long long lo, hi, leaf_offset; };
__device__ __forceinline__ NodeRange node_range(const EmitJob& E) {
    NodeRange r;
    if (E.info) {
        r.lo = E.info->node_lo; r.hi = E.info->node_hi; r.leaf_offset = E.info->leaf_offset;
    } else { r.lo = E.pos_lo; r.hi = E.pos_hi; r.leaf_offset = E.leaf_offset; }
    if (r.hi - r.lo > E.cap) r.hi = r.lo + E.cap;
    return r;
}
// where record `pos` goes, or NULL when it belongs to another rank (or does not fit the buffer)
__device__ __forceinline__ unsigned long long* node_slot(const EmitJob& E, const NodeRange& r, unsigned long long pos) {
    if (!E.write_records || pos < r.lo || pos >= r.hi) return nullptr;
    return E.nodes + (pos - r.lo) * 3ULL;
}

// -levels: data index of an internal node = records written before it. Payload mode interleaves the
// leaf records (written at addVoxel) with the internal ones (written in post-order by groupNodes):
// 1 + leaves up to the end of its subtree + its post-order rank. Binary mode has no leaf records: 2 + rank.
// Sharded: `leaves_through` counts this rank's leaves only, E.leaf_offset adds the slabs of lower ranks; `rank` is global
// (the ibase values of a rank's top tiles come from the merge of the shared levels).
__device__ __forceinline__ unsigned long long internal_data_index(const EmitJob& E, unsigned long long leaves_through, unsigned long long rank) {
    return (E.leaf_data_mode ? 1ULL + E.leaf_offset + leaves_through : 2ULL) + rank;
}

// One tile of an upper level (tile = node at depth d with two packed levels): writes the records of its
// grandchildren (tiles of the level below) and children, and the file base of every grandchild subtree.
__device__ __forceinline__ void emit_upper_tile(const Level& L, const Level& C, const EmitJob& E, const NodeRange& R, unsigned long long i, int lane) {
    const unsigned long long W = L.mask[i], fc = L.fc[i], base = L.base[i];
    const unsigned long long S = L.ps[i + 1] - L.ps[i];
    const uint32_t nzb = nonzero_bytes(W);
    const unsigned long long ps0 = C.ps[fc];
    if (L.clear && lane == 31) L.clear[L.key[i]] = 0ULL;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bit = lane + 32 * h;
        if ((W >> bit) & 1ULL) {
            const int k = bit >> 3;
            const unsigned long long c = fc + __popcll(W & lowmask(bit));
            // subtree region of grandchild c
            const unsigned long long gbase = base + (C.ps[c] - ps0) + __popcll(W & lowmask(8 * k));
            C.base[c] = gbase;
            // its record sits in the children block of byte k
            const unsigned long long blk = base + (C.ps[fc + __popcll(W & lowmask(8 * (k + 1)))] - ps0) + __popcll(W & lowmask(8 * k));
            const unsigned long long pos = blk + __popcll(W & lowmask(bit) & ~lowmask(8 * k));
            const unsigned long long gw = C.mask[c];
            const uint32_t gnz = nonzero_bytes(gw);
            const unsigned long long gS = C.ps[c + 1] - C.ps[c];
            unsigned long long gdata = 0ULL;
            if (E.levels) {
                const unsigned long long gib = L.ibase[i] + (C.pi[c] - C.pi[fc]) + __popc(nzb & ((1u << k) - 1u));
                C.ibase[c] = gib;
                gdata = internal_data_index(E, C.pl[c + 1], gib + (C.pi[c + 1] - C.pi[c]) - 1ULL);
            }
            if (unsigned long long* o = node_slot(E, R, pos)) {
                o[0] = gdata;
                o[1] = gbase + gS - __popc(gnz);
                o[2] = child_offsets(gnz);
            }
        }
    }
    if (lane < 8 && ((nzb >> lane) & 1u)) {
        const int k = lane;
        const unsigned long long cend = fc + __popcll(W & lowmask(8 * (k + 1)));
        const unsigned long long blk = base + (C.ps[cend] - ps0) + __popcll(W & lowmask(8 * k));
        const unsigned long long pos = base + S - __popc(nzb) + __popc(nzb & ((1u << k) - 1u));
        unsigned long long cdata = 0ULL;
        if (E.levels) cdata = internal_data_index(E, C.pl[cend], L.ibase[i] + (C.pi[cend] - C.pi[fc]) + __popc(nzb & ((1u << k) - 1u)));
        if (unsigned long long* o = node_slot(E, R, pos)) {
            o[0] = cdata;
            o[1] = blk;
            o[2] = child_offsets((uint32_t)((W >> (8 * k)) & 0xffULL));
        }
    }
    if (E.root_here && lane == 8) {
        if (unsigned long long* o = node_slot(E, R, S)) {
            o[0] = E.levels ? internal_data_index(E, L.pl[i + 1], L.ibase[i] + (L.pi[i + 1] - L.pi[i]) - 1ULL) : 0ULL;
            o[1] = base + S - __popc(nzb);
            o[2] = child_offsets(nzb);
        }
    }
}
// 24-byte node record as 16 + 8 or 8 + 16 bytes, whichever is aligned
__device__ __forceinline__ void store_record(unsigned long long* o, unsigned long long d0, unsigned long long d1, unsigned long long d2) {
    if (((uintptr_t)o & 15) == 0) { asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(o), "l"(d0), "l"(d1)); o[2] = d2; }
    else { o[0] = d0; asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(o + 1), "l"(d1), "l"(d2)); }
}
// The same tile without -levels, trimmed for instruction count (the generic version spends ~400 warp instructions per
// tile, and the level above the bricks holds 10^6 tiles at 8192^3): everything that only depends on the tile word is
// computed once per warp -- per-byte child counts and their running sums with one SWAR popcount and one multiply --,
// child_offsets comes from a shared-memory table, records go out as 16 + 8 byte stores.
__device__ __forceinline__ void emit_upper_tile_fast(const Level& L, const Level& C, const EmitJob& E, const NodeRange& R, unsigned long long i, int lane) {
    const unsigned long long W = L.mask[i], fc = L.fc[i], base = L.base[i];
    const unsigned long long S = L.ps[i + 1] - L.ps[i];
    if (L.clear && lane == 31) L.clear[L.key[i]] = 0ULL;
    // per-byte popcounts of W and their inclusive running sums (each fits a byte: at most 64)
    unsigned long long bp = W - ((W >> 1) & 0x5555555555555555ULL);
    bp = (bp & 0x3333333333333333ULL) + ((bp >> 2) & 0x3333333333333333ULL);
    bp = (bp + (bp >> 4)) & 0x0f0f0f0f0f0f0f0fULL;
    const unsigned long long cum = bp * 0x0101010101010101ULL;
    const uint32_t nzb = nonzero_bytes(W);
    const unsigned long long ps0 = C.ps[fc];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bit = lane + 32 * h, k = bit >> 3;
        if ((W >> bit) & 1ULL) {
            const uint32_t byte = (uint32_t)(W >> (8 * k)) & 0xffu;
            const uint32_t through = (uint32_t)(cum >> (8 * k)) & 0xffu;           // children in bytes 0..k
            const uint32_t before = through - (uint32_t)__popc(byte);             // children in bytes 0..k-1
            const uint32_t inb = (uint32_t)__popc(byte & ((1u << (bit & 7)) - 1u)); // children of byte k below this one
            const unsigned long long c = fc + before + inb;
            const unsigned long long psc = C.ps[c], psc1 = C.ps[c + 1], pend = C.ps[fc + through];
            const uint32_t gnz = nonzero_bytes(C.mask[c]);
            const unsigned long long gbase = base + (psc - ps0) + before;         // subtree region of grandchild c
            C.base[c] = gbase;
            const unsigned long long pos = base + (pend - ps0) + before + inb;    // its record, in the children block of byte k
            if (unsigned long long* o = node_slot(E, R, pos)) store_record(o, 0ULL, gbase + (psc1 - psc) - __popc(gnz), child_offsets_lut(gnz));
        }
    }
    if (lane < 8 && ((nzb >> lane) & 1u)) {
        const int k = lane;
        const uint32_t byte = (uint32_t)(W >> (8 * k)) & 0xffu;
        const uint32_t through = (uint32_t)(cum >> (8 * k)) & 0xffu, before = through - (uint32_t)__popc(byte);
        const unsigned long long blk = base + (C.ps[fc + through] - ps0) + before;
        const unsigned long long pos = base + S - __popc(nzb) + __popc(nzb & ((1u << k) - 1u));
        if (unsigned long long* o = node_slot(E, R, pos)) store_record(o, 0ULL, blk, child_offsets_lut(byte));
    }
    if (E.root_here && lane == 8) {
        if (unsigned long long* o = node_slot(E, R, S)) store_record(o, 0ULL, base + S - __popc(nzb), child_offsets_lut(nzb));
    }
}
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_emit_upper(Level L, Level C, EmitJob E) {
    if (build_aborted(E.info)) return;
    const unsigned long long i = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= level_n(L)) return;
    const NodeRange R = node_range(E);
    if (E.levels) emit_upper_tile(L, C, E, R, i, threadIdx.x & 31);
    else emit_upper_tile_fast(L, C, E, R, i, threadIdx.x & 31);
}

// The upper levels of the device-driven build (no -levels). PERSISTENT warps walk batches of `bs` tiles (bs = 32 for the
// 10^6-tile levels of the big grids, smaller when the level has fewer tiles than the GPU has warps): lane q loads the
// descriptor of the batch's q-th tile (coalesced), and inside the batch the child loads of tile q + 1 are issued before
// the records of tile q are written -- one warp per tile spent three dependent memory latencies per tile (info -> tile ->
// children) and nothing else. Lanes map to the tile's children by RANK (the children of a tile are consecutive in the
// list below, so all loads are coalesced; 16 children per tile is typical and one round of 32 lanes covers it): the byte k
// a child belongs to follows from the running per-byte counts with one SWAR compare, its position needs no bit index.
// CHILD_RECS = false (level 1 of the device-driven build): the records of the children -- the bricks -- are written by
// k_emit_leaf, which has them in the middle of the contiguous stream it writes; this kernel then only places the bricks
// (C.base) and writes the tiles' own child records.
struct UpperChild { unsigned long long psc, psc1, pend, gm; };
template <bool CHILD_RECS>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_emit_upper_fast(Level L, Level C, EmitJob E, int bs) {
    if (build_aborted(E.info)) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long n = level_n(L);
    const unsigned long long nbatch = (n + bs - 1) / bs;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    const NodeRange R = node_range(E);
    constexpr unsigned long long ONES = 0x0101010101010101ULL, HIGH = 0x8080808080808080ULL;
    for (unsigned long long batch = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); batch < nbatch; batch += nwarps) {
        const unsigned long long t0 = batch * bs;
        const int cnt = (int)min((unsigned long long)bs, n - t0);
        unsigned long long myW = 0ULL, myFc = 0ULL, myBase = 0ULL, myS = 0ULL;
        if (lane < cnt) {
            const unsigned long long t = t0 + lane;
            myW = L.mask[t]; myFc = L.fc[t]; myBase = L.base[t]; myS = L.ps[t + 1] - L.ps[t];
            if (L.clear) L.clear[L.key[t]] = 0ULL;
        }
        // state of the tile whose loads are in flight
        unsigned long long nW = 0ULL, nFc = 0ULL, nCum = 0ULL;
        uint32_t nThrough = 0, nByte = 0;                       // of this lane's child (rank = lane)
        UpperChild nx = {0ULL, 0ULL, 0ULL, 0ULL};
        auto child_of = [&](unsigned long long W, unsigned long long cum, int r, uint32_t& through, uint32_t& byte) {
            // bytes whose running count is <= r come before the child's byte: cum is monotone, every byte <= 64
            const int k = 8 - __popcll(((cum | HIGH) - (unsigned long long)(r + 1) * ONES) & HIGH);
            through = (uint32_t)(cum >> (8 * k)) & 0xffu;
            byte = (uint32_t)(W >> (8 * k)) & 0xffu;
        };
        auto load_child = [&](unsigned long long fc, int r, uint32_t through, UpperChild& u) {
            const unsigned long long c = fc + r;
            u.psc = C.ps[c]; u.pend = C.ps[fc + through];
            if (CHILD_RECS) { u.psc1 = C.ps[c + 1]; u.gm = C.mask[c]; }
        };
        auto prepare = [&](int q) {
            nW = __shfl_sync(0xffffffffu, myW, q);
            nFc = __shfl_sync(0xffffffffu, myFc, q);
            unsigned long long bp = nW - ((nW >> 1) & 0x5555555555555555ULL);
            bp = (bp & 0x3333333333333333ULL) + ((bp >> 2) & 0x3333333333333333ULL);
            bp = (bp + (bp >> 4)) & 0x0f0f0f0f0f0f0f0fULL;
            nCum = bp * ONES;                                    // inclusive running per-byte child counts
            const int nchild = (int)(nCum >> 56);
            nx.psc = 0ULL; nx.psc1 = 0ULL; nx.pend = 0ULL; nx.gm = 0ULL;
            if (lane < nchild) { child_of(nW, nCum, lane, nThrough, nByte); load_child(nFc, lane, nThrough, nx); }
        };
        auto write_child = [&](unsigned long long base, unsigned long long ps0, int r, uint32_t through, uint32_t byte, const UpperChild& u, unsigned long long fc) {
            const uint32_t before = through - (uint32_t)__popc(byte);                 // children in the bytes below
            const unsigned long long gbase = base + (u.psc - ps0) + before;           // subtree region of this grandchild
            C.base[fc + r] = gbase;
            if (CHILD_RECS) {
                const unsigned long long pos = base + (u.pend - ps0) + r;             // its record, in the children block of its byte
                if (unsigned long long* o = node_slot(E, R, pos)) {
                    const uint32_t gnz = nonzero_bytes(u.gm);
                    store_record(o, 0ULL, gbase + (u.psc1 - u.psc) - __popc(gnz), child_offsets_lut(gnz));
                }
            }
        };
        prepare(0);
        for (int q = 0; q < cnt; q++) {
            const unsigned long long W = nW, fc = nFc, cum = nCum;
            const uint32_t through = nThrough, byte = nByte;
            const UpperChild u = nx;
            const unsigned long long base = __shfl_sync(0xffffffffu, myBase, q), S = __shfl_sync(0xffffffffu, myS, q);
            if (q + 1 < cnt) prepare(q + 1);
            const int nchild = (int)(cum >> 56);
            const unsigned long long ps0 = __shfl_sync(0xffffffffu, u.psc, 0);       // C.ps[fc]
            if (lane < nchild) write_child(base, ps0, lane, through, byte, u, fc);
            if (nchild > 32) {                                                        // second round (rare)
                const int r = lane + 32;
                if (r < nchild) {
                    uint32_t th, by; UpperChild v;
                    child_of(W, cum, r, th, by);
                    load_child(fc, r, th, v);
                    write_child(base, ps0, r, th, by, v, fc);
                }
            }
            // the tile's own children: one record per non-zero byte, behind the subtrees
            const uint32_t nzb = nonzero_bytes(W);
            const int k = lane & 7;
            const uint32_t by = (uint32_t)(W >> (8 * k)) & 0xffu;
            const uint32_t th = (uint32_t)(cum >> (8 * k)) & 0xffu, before = th - (uint32_t)__popc(by);
            // the end of byte k's subtrees: the value its last child (rank th - 1) already holds
            const unsigned long long pend_sh = __shfl_sync(0xffffffffu, u.pend, (int)(th - 1u) & 31);
            if (lane < 8 && by) {
                const unsigned long long pend = th <= 32 ? pend_sh : C.ps[fc + th];
                const unsigned long long blk = base + (pend - ps0) + before;
                const unsigned long long pos = base + S - __popc(nzb) + __popc(nzb & ((1u << k) - 1u));
                if (unsigned long long* o = node_slot(E, R, pos)) store_record(o, 0ULL, blk, child_offsets_lut(by));
            }
            if (E.root_here && lane == 8) {
                if (unsigned long long* o = node_slot(E, R, S)) store_record(o, 0ULL, base + S - __popc(nzb), child_offsets_lut(nzb));
            }
        }
    }
}

// Level 0 (bricks): the whole subtree region of a brick is contiguous in the file: popc(W) leaf records followed by
// one record per non-zero byte. A warp takes 32 consecutive bricks (one coalesced load of their words and bases) and
// walks them FOUR at a time, eight lanes per brick:
//   * the run of leaf records is streamed out as 16-BYTE stores, 128 contiguous bytes per brick and step (which field
//     of the 24-byte record a word holds is its index mod 3, independent of the brick); one 8-byte store in front /
//     behind where the run starts / ends on an odd word;
//   * lane k of the group writes the child record of byte k (16 + 8 bytes).
// Writes are guarded by this rank's range of the file and the capacity of the buffer (speculative emission).
constexpr int EMIT_TILES_PER_WARP = 32;         // bricks per batch of a warp
// 16-byte store to a 16-byte aligned address. Inline PTX on purpose: written as a C++ vector store, the two branches of
// "aligned: 16 + 8, else 8 + 16" write the same bytes and the compiler folds them into ONE (then misaligned) form.
__device__ __forceinline__ void st128(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b));
}
// PERSISTENT warps: every warp walks batches of 32 bricks with a stride of the whole grid; the descriptors (word, file
// base, leaf rank) of the NEXT batch are loaded before the current one is written, so that no warp ever sits idle on
// its loads (at 8192^3 the lists come from DRAM: without the prefetch the kernel ran at 44 % of the HBM peak).
// WHOLE SECTORS. Records are 24 bytes, so a brick's region starts anywhere on an 8-byte grid -- and a store instruction
// that leaves a 32-byte sector half written costs the L2 a read-modify-write: tools/native/store_probe.cu measures the
// bare store pattern of "eight lanes write 128 contiguous bytes of a brick" at 5.9 TB/s when the runs start on sector
// boundaries and at 2.4 TB/s when they start 16 bytes off (the rate this kernel ran at before). So every store
// instruction here covers whole sectors wherever the file allows it:
//   interior   the leaf run between its first and last sector boundary, as aligned 16-byte pairs (lanes 2 j, 2 j + 1 of a
//              group fill one sector)
//   seam       what lies between the interiors of two consecutive bricks -- the last 0..3 words of this run, this brick's
//              child records (<= 24 words), the first 0..3 words of the NEXT brick's run when it follows without a gap --
//              is composed in shared memory (<= 30 words) and written as aligned pairs too: it starts and ends on sector
//              boundaries by construction
// Partial sectors remain only where another kernel's records follow a brick (the parent's children blocks, about every
// second brick) and at the ends of a batch.
// In geometry-only mode the leaf run is a constant pattern of period three words (data 1, children base 0, offsets ~0: the
// word of field f is 1 - f).
//   BLOCKS (device-driven build): behind the last brick of a byte of the parent tile comes the parent's children block of
//              that byte -- the records OF the bricks (data 0, base + leaves, offsets of the non-zero bytes), everything a
//              lane needs is in the brick list. The seam carries it along (<= 8 more records), and all bricks of a
//              level-1 tile become one uninterrupted stream.
// What depends on the brick alone is computed ONCE, by the lane that holds the brick (32 bricks at a time), and left in
// shared memory for the eight lanes that write it; batches are handed out in file order by a ticket (E.ticket), which keeps
// the DRAM pages of the whole GPU's stores close together (store_probe: + 25 % over a fixed round-robin).
struct LeafBrick { unsigned long long W, ab, cbase, off; };      // word | flags and counts | base + leaves | offsets of the non-zero bytes
template <bool PAYLOAD, int MINB, bool BLOCKS>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) k_emit_leaf(Level L, EmitJob E) {
    constexpr int SEAM = BLOCKS ? 64 : 32;
    __shared__ __align__(16) unsigned long long s_seam[WARPS_PER_BLOCK][2][4][SEAM];
    __shared__ unsigned long long s_lut[256];                                     // child_offsets by byte
    __shared__ __align__(16) LeafBrick s_brick[WARPS_PER_BLOCK][32];
    __shared__ uint32_t s_kb[WARPS_PER_BLOCK][32];                                // key >> 3: (level-1 tile, byte) of the brick
    __shared__ ulonglong2 s_prev[WARPS_PER_BLOCK][8];                             // the same two record words of the 7 bricks before the batch
    __shared__ uint32_t s_pkb[WARPS_PER_BLOCK][8];
    s_lut[threadIdx.x & 255] = g_child_offsets.v[threadIdx.x & 255];
    __syncthreads();
    if (build_aborted(E.info)) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long n = level_n(L);
    const unsigned long long nbatch = (n + 31) / 32;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    // the first two batches of a warp are fixed, the rest are tickets, drawn two batches ahead (the atomic's round trip is
    // hidden behind a whole batch, and a small job -- two batches per warp at 1024^3 -- never waits for one)
    auto draw = [&]() -> unsigned long long { return lane == 0 ? 2ULL * nwarps + atomicAdd(E.ticket, 1ULL) : 0ULL; };
    unsigned long long batch = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + wid, batch1 = batch + nwarps;
    if (batch >= nbatch) return;
    const NodeRange R = node_range(E);
    unsigned long long* const nodes = E.nodes;
    const int g = lane >> 3, s = lane & 7;                                        // group (brick of the round), lane in the group
    const unsigned gmask = 0xffu << (8 * g);
    const int s2 = (2 * s) % 3;
    constexpr unsigned REL_BITS = 25, REL_MASK = (1u << REL_BITS) - 1u;
    constexpr unsigned F_NEXT = 1u << REL_BITS, F_HEAD = 2u << REL_BITS, F_VALID = 4u << REL_BITS, F_LOB = 8u << REL_BITS;
    unsigned long long nW = 0ULL, nBase = 0ULL, nFc = 0ULL, nKey = ~0ULL, nKeyX = ~0ULL, pW = 0ULL, pBase = 0ULL, pKey = ~0ULL;
    auto prefetch = [&](unsigned long long bt) {              // loads only: a store that depends on one of them would stall the warp here
        const unsigned long long t = bt * 32 + lane;
        nW = 0ULL; nBase = 0ULL; nFc = 0ULL; nKey = ~0ULL; nKeyX = ~0ULL; pW = 0ULL; pBase = 0ULL; pKey = ~0ULL;
        if (bt < nbatch && t < n) {
            nW = L.mask[t];
            nBase = L.base[t];
            if (PAYLOAD) nFc = L.fc[t];
            if (L.clear || BLOCKS) nKey = L.key[t];
            if (BLOCKS && lane == 31 && t + 1 < n) nKeyX = L.key[t + 1];          // the brick behind the batch
        }
        // the seven bricks before the batch (a byte of the parent tile may begin there): lane l takes brick t0 - 1 - l
        if (BLOCKS && bt < nbatch && lane < 7 && bt * 32 > (unsigned long long)lane) {
            const unsigned long long u = bt * 32 - 1 - lane;
            pKey = L.key[u]; pW = L.mask[u]; pBase = L.base[u];
        }
    };
    prefetch(batch);
    while (batch < nbatch) {
        const unsigned long long t0 = batch * 32;
        const int cnt = (int)min(32ULL, n - t0);
        const unsigned long long myW = nW, myBase = nBase, myFc = nFc, myKey = nKey, myKeyX = nKeyX, prevW = pW, prevBase = pBase, prevKey = pKey;
        unsigned long long drawn = 0ULL;
        if (E.ticket) drawn = draw();
        batch = batch1;
        prefetch(batch);                                                          // in flight while this batch is written
        if (L.clear && myKey != ~0ULL) L.clear[myKey] = 0ULL;                     // the brick's word in the bit-grid is consumed: leave it clean
        const unsigned nleaf = (unsigned)__popcll(myW);
        const uint32_t nzb = nonzero_bytes(myW);
        const unsigned nzc = (unsigned)__popc(nzb);
        const unsigned S = nleaf + nzc;                                               // records of the brick's region
        // a rank's own bricks always lie inside its range; the capacity guard drops whole bricks
        const bool ok = lane < cnt && myW != 0ULL && E.write_records && myBase >= R.lo &&
                        myBase + (unsigned long long)(S + (E.root_here ? 1u : 0u)) <= R.hi;
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (!okm) { batch1 = E.ticket ? __shfl_sync(0xffffffffu, drawn, 0) : batch + nwarps; continue; }
        const unsigned long long rel = myBase - R.lo;                                 // first record of the region in the buffer
        const unsigned long long ref = __shfl_sync(0xffffffffu, rel, __ffs(okm) - 1); // ... of the batch's first brick
        const unsigned long long leaf1 = 1ULL + R.leaf_offset + myFc;                 // payload: data index of the brick's first leaf
        // (the gridsize-4 root brick takes the plain path below; so would a brick further than 2^25 records from the batch's
        // first one, which no tree has: between two bricks that follow each other in the list lie at most the children blocks
        // of their ancestors)
        const bool in_group = ok && !E.root_here && (rel - ref) <= (unsigned long long)REL_MASK;
        if (ok && !in_group) {
            unsigned long long* const out = nodes + 3ULL * rel;
            for (unsigned q = 0; q < nleaf; q++) { out[3 * q] = PAYLOAD ? leaf1 + q : 1ULL; out[3 * q + 1] = 0ULL; out[3 * q + 2] = ~0ULL; }
            unsigned long long* o = out + 3u * nleaf;
            unsigned long long cb = myBase;                                           // first leaf of the byte
            for (uint32_t m = nzb; m; m &= m - 1u) {
                const int k = __ffs(m) - 1;
                const uint32_t byte = (uint32_t)(myW >> (8 * k)) & 0xffu;
                o[0] = 0ULL; o[1] = cb; o[2] = child_offsets_lut(byte);
                cb += (unsigned)__popc(byte);
                o += 3;
            }
            if (E.root_here) { o[0] = 0ULL; o[1] = myBase + nleaf; o[2] = child_offsets(nzb); }   // gridsize 4: the single brick is the root
        }
        // does the next / previous brick of the batch follow without a gap?
        const unsigned gm = __ballot_sync(0xffffffffu, in_group);
        bool contigN, contigP, last_of_byte = false;
        if (BLOCKS) {
            // bricks of one level-1 tile form one stream (this kernel writes the children blocks between them); the tile's
            // own records follow its last brick
            unsigned long long keyN = __shfl_down_sync(0xffffffffu, myKey, 1), keyP = __shfl_up_sync(0xffffffffu, myKey, 1);
            if (lane == 31) keyN = myKeyX;
            const unsigned long long key_before = __shfl_sync(0xffffffffu, prevKey, 0);
            if (lane == 0) keyP = key_before;
            last_of_byte = (keyN >> 3) != (myKey >> 3);
            contigN = in_group && lane < 31 && ((gm >> (lane + 1)) & 1u) && (keyN >> 6) == (myKey >> 6);
            contigP = in_group && lane > 0 && ((gm >> (lane - 1)) & 1u) && (keyP >> 6) == (myKey >> 6);
        } else {
            const unsigned long long relN = __shfl_down_sync(0xffffffffu, rel, 1), relP = __shfl_up_sync(0xffffffffu, rel, 1);
            const unsigned SP = __shfl_up_sync(0xffffffffu, S, 1);
            contigN = in_group && lane < 31 && ((gm >> (lane + 1)) & 1u) && rel + S == relN;
            contigP = in_group && lane > 0 && ((gm >> (lane - 1)) & 1u) && relP + SP == rel;
        }
        const unsigned rel32o = (unsigned)(rel - ref);
        const unsigned fa = in_group ? (rel32o | (contigN ? F_NEXT : 0u) | (contigP ? 0u : F_HEAD) | F_VALID | (last_of_byte ? F_LOB : 0u)) : 0u;
        // leaves | non-zero bytes << 8 | their count << 16 | position of the region's first word inside its 32-byte sector << 20
        // (the buffer is 256-byte aligned)
        const unsigned fb = nleaf | (nzb << 8) | (nzc << 16) | (((3u * ((unsigned)ref + rel32o)) & 3u) << 20);
        __syncwarp();                                                                 // (nobody still reads the previous batch)
        {
            LeafBrick B;
            B.W = myW; B.ab = (unsigned long long)fa | ((unsigned long long)fb << 32);
            B.cbase = myBase + nleaf; B.off = s_lut[nzb];
            s_brick[wid][lane] = B;
            if (BLOCKS) {
                s_kb[wid][lane] = (uint32_t)(myKey >> 3);
                if (lane < 7) {
                    s_prev[wid][lane].x = prevBase + (unsigned)__popcll(prevW);
                    s_prev[wid][lane].y = s_lut[nonzero_bytes(prevW)];
                    s_pkb[wid][lane] = prevKey != ~0ULL ? (uint32_t)(prevKey >> 3) : 0xffffffffu;
                }
            }
        }
        __syncwarp();
        unsigned long long* const outref = nodes + 3ULL * ref;
        const unsigned long long baseref = R.lo + ref;
        const unsigned long long room = (R.hi - R.lo) - ref;                          // records of the buffer from the batch's first brick on
        for (int r = 0; r < cnt; r += 4) {
            const int bi = r + g;
            const ulonglong2 bw = *reinterpret_cast<const ulonglong2*>(&s_brick[wid][bi].W);
            const unsigned pk = (unsigned)bw.y, pb = (unsigned)(bw.y >> 32);
            if (!(pk & F_VALID)) continue;                                            // (uniform in the group)
            const unsigned long long W = bw.x;
            unsigned long long d0 = 1ULL;
            const unsigned rel32 = pk & REL_MASK;
            const unsigned nl = pb & 0xffu, nz8 = (pb >> 8) & 0xffu, nzn = (pb >> 16) & 0xfu, a0 = (pb >> 20) & 3u;
            // the children block behind the last brick of a byte: lane s takes the byte's s-th brick from the end
            bool blk_on = false;
            unsigned long long blk_cbase = 0ULL, blk_off = 0ULL;
            if (BLOCKS && (pk & F_LOB)) {
                const int u = bi - s;
                if (u >= 0) {
                    blk_on = s_kb[wid][u] == s_kb[wid][bi];
                    if (blk_on) { const ulonglong2 rec = *reinterpret_cast<const ulonglong2*>(&s_brick[wid][u].cbase); blk_cbase = rec.x; blk_off = rec.y; }
                } else {                                                              // the byte began in the batch before this one
                    blk_on = s_pkb[wid][-u - 1] == s_kb[wid][bi];
                    if (blk_on) { const ulonglong2 rec = s_prev[wid][-u - 1]; blk_cbase = rec.x; blk_off = rec.y; }
                }
            }
            if (PAYLOAD) d0 = 1ULL + R.leaf_offset + L.fc[t0 + (unsigned)bi];
            const bool next_follows = (pk & F_NEXT) != 0u, own_head = (pk & F_HEAD) != 0u;
            unsigned long long* const out = outref + 3ULL * rel32;                    // word 0 of the brick's region
            const unsigned words = 3u * nl;
            const unsigned head = (4u - a0) & 3u;                                     // words before the first sector boundary
            const unsigned tail = (a0 + words) & 3u;                                  // words behind the last one
            const unsigned iend = words - tail;                                       // interior: [head, iend), a multiple of four words long (possibly empty)
            // ---- interior ----
            int f = (head == 3u ? 0 : (int)head) + s2; if (f >= 3) f -= 3;            // field of this lane's first word, (head + 2 s) % 3
            if (!PAYLOAD) {
                // the field advances by 16 % 3 == 1 per store: three stores are one period
                const int f1 = f == 2 ? 0 : f + 1, f2 = f1 == 2 ? 0 : f1 + 1;
                const unsigned long long x0 = (unsigned long long)(long long)(1 - f), x1 = (unsigned long long)(long long)(1 - f1),
                                         x2 = (unsigned long long)(long long)(1 - f2);
                for (unsigned q = head + 2u * s; q < iend; q += 48u) {
                    st128(out + q, x0, x1);
                    if (q + 16u < iend) st128(out + q + 16u, x1, x2);
                    if (q + 32u < iend) st128(out + q + 32u, x2, x0);
                }
            } else {
                for (unsigned q = head + 2u * s; q < iend; q += 16u) {
                    unsigned long long a, b;                    // the data index of the record in the data field
                    if (f == 0) { a = d0 + (q / 3u); b = 0ULL; }
                    else if (f == 1) { a = 0ULL; b = ~0ULL; }
                    else { a = ~0ULL; b = d0 + ((q + 1u) / 3u); }
                    st128(out + q, a, b);
                    f = f == 2 ? 0 : f + 1;
                }
            }
            // ---- head, when the brick before this one does not cover it ----
            if (own_head && (unsigned)s < head) out[s] = s == 0 ? d0 : (unsigned long long)(long long)(1 - s);
            // ---- seam ----
            unsigned long long* const sb = s_seam[wid][(r >> 2) & 1][g];
            if ((unsigned)s < tail) {                                                 // word iend + s of the run: field 3 - (tail - s), of the LAST leaf for field 0
                const int fl = (int)(3u - (tail - (unsigned)s)) % 3;
                sb[s] = fl == 0 ? (PAYLOAD ? d0 + (nl - 1u) : 1ULL) : (unsigned long long)(long long)(1 - fl);
            }
            const uint32_t byte = (uint32_t)((W >> (8 * s)) & 0xffULL);
            if (byte) {                                                               // child record of byte s
                const unsigned j0 = tail + 3u * (unsigned)__popc(nz8 & ((1u << s) - 1u));
                sb[j0] = 0ULL;
                sb[j0 + 1] = baseref + rel32 + (unsigned)__popcll(W & lowmask(8 * s));
                sb[j0 + 2] = s_lut[byte];
            }
            unsigned len = tail + 3u * nzn;
            if (BLOCKS && (pk & F_LOB)) {
                const unsigned m = (unsigned)__popc(__ballot_sync(gmask, blk_on));    // bricks of the byte (its last s + 1 .. a prefix of the lanes)
                const bool fits = (unsigned long long)rel32 + nl + nzn + m <= room;   // (speculative emission into a smaller buffer)
                if (fits) {
                    if (blk_on) {
                        const unsigned j0 = len + 3u * (m - 1u - (unsigned)s);
                        sb[j0] = 0ULL;
                        sb[j0 + 1] = blk_cbase;
                        sb[j0 + 2] = blk_off;
                    }
                    len += 3u * m;
                }
            }
            if (next_follows) {                                                       // up to the next brick's first sector boundary
                const unsigned nh = (4u - (len & 3u)) & 3u;
                if ((unsigned)s < nh) sb[len + s] = s == 0 ? (PAYLOAD ? d0 + nl : 1ULL) : (unsigned long long)(long long)(1 - s);
                len += nh;
            }
            __syncwarp(gmask);
            unsigned long long* const so = out + iend;
#pragma unroll
            for (int i = 0; i < SEAM / 16; i++) {
                const unsigned j = 2u * s + 16u * i;
                if (16u * i < len) {                                                  // (uniform in the group)
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sb + j);
                    if (j + 1u < len) st128(so + j, v.x, v.y);
                    else if (j < len) so[j] = v.x;
                }
            }
        }
        if (E.ticket) batch1 = __shfl_sync(0xffffffffu, drawn, 0); else batch1 = batch + nwarps;
    }
}

// -levels variant of the brick emitter (one warp per brick, lanes = voxels): leaf data indices
// are shifted by the internal records written before them, children carry a data index.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_emit_leaf_levels(Level L, EmitJob E) {
    const unsigned long long i = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= L.n) return;
    const NodeRange R = node_range(E);
    const int lane = threadIdx.x & 31;
    const unsigned long long W = L.mask[i], base = L.base[i], ib = L.ibase[i], lp = L.fc[i];
    const uint32_t nzb = nonzero_bytes(W);
    const int nleaf = __popcll(W);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bit = lane + 32 * h;
        if ((W >> bit) & 1ULL) {
            const int r = __popcll(W & lowmask(bit));
            if (unsigned long long* o = node_slot(E, R, base + r)) {
                o[0] = E.leaf_data_mode ? 1ULL + E.leaf_offset + lp + r + ib + __popc(nzb & ((1u << (bit >> 3)) - 1u)) : 1ULL;
                o[1] = 0ULL;
                o[2] = ~0ULL;
            }
        }
    }
    if (lane < 8 && ((nzb >> lane) & 1u)) {
        const int k = lane;
        if (unsigned long long* o = node_slot(E, R, base + nleaf + __popc(nzb & ((1u << k) - 1u)))) {
            o[0] = internal_data_index(E, lp + __popcll(W & lowmask(8 * (k + 1))), ib + __popc(nzb & ((1u << k) - 1u)));
            o[1] = base + __popcll(W & lowmask(8 * k));
            o[2] = child_offsets((uint32_t)((W >> (8 * k)) & 0xffULL));
        }
    }
    if (E.root_here && lane == 8) {   // gridsize 4
        if (unsigned long long* o = node_slot(E, R, base + nleaf + __popc(nzb))) {
            o[0] = internal_data_index(E, lp + nleaf, ib + __popc(nzb));
            o[1] = base + nleaf;
            o[2] = child_offsets(nzb);
        }
    }
}

// -levels data records, bottom-up (OctreeBuilder.cpp:82-99): a parent's colour is the sum of its
// children's cached colours in slot order divided by the number of non-null children, its normal the
// normalised mean normal. One thread per tile: the 8 byte-children first, then the tile's own node.
__device__ __forceinline__ float canon_nan(float x) { return (x != x) ? __uint_as_float(0xFFC00000u) : x; }   // x86 default NaN
__device__ __forceinline__ void finish_average(const float* sum, float notnull, float* out) {
    out[0] = fdiv(sum[0], notnull); out[1] = fdiv(sum[1], notnull); out[2] = fdiv(sum[2], notnull);
    const float tx = fdiv(sum[3], notnull), ty = fdiv(sum[4], notnull), tz = fdiv(sum[5], notnull);
    const float inv = fdiv(1.0f, fsqrt(dot3(tx, ty, tz, tx, ty, tz)));
    out[3] = fmul(tx, inv); out[4] = fmul(ty, inv); out[5] = fmul(tz, inv);
}
__device__ __forceinline__ void write_data_record(float* data, unsigned long long idx, const float* c) {
    float4* o = reinterpret_cast<float4*>(data + idx * 8ULL);
    o[0] = make_float4(0.0f, 0.0f, canon_nan(c[0]), canon_nan(c[1]));              // morton 0
    o[1] = make_float4(canon_nan(c[2]), canon_nan(c[3]), canon_nan(c[4]), canon_nan(c[5]));
}
// `data` is this rank's part of the data file biased by -data_lo records (record index = global data index).
// cache_only (sharded -levels, before the exchange): no record is written; the leaf records are read from a local
// staging array in leaf-rank order (record 1 + leaf rank) -- only the tiles' data caches are wanted, for the table.
__global__ void __launch_bounds__(128) k_levels_data(Level L, Level C, int level, EmitJob E, float* data, int real_node, int cache_only) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n) return;
    const unsigned long long W = L.mask[i], fc = L.fc[i], ib = cache_only ? 0ULL : L.ibase[i];
    const uint32_t nzb = nonzero_bytes(W);
    float wsum[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    float wn = 0.0f;
    for (int k = 0; k < 8; k++) {
        const uint32_t byte = (uint32_t)((W >> (8 * k)) & 0xffULL);
        if (!byte) continue;
        float csum[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
        float cn = 0.0f;
        const unsigned long long before = __popcll(W & lowmask(8 * k));
        const int krank = __popc(nzb & ((1u << k) - 1u));
        for (int b = 0; b < 8; b++) {
            if (!((byte >> b) & 1u)) continue;
            const unsigned long long c = fc + before + __popc(byte & ((1u << b) - 1u));   // child tile index / leaf rank
            cn = fadd(cn, 1.0f);
            if (level == 0) {
                if (E.leaf_data_mode) {     // payload leaves cache their data record (OctreeBuilder.cpp:160)
                    const float* rec = data + (cache_only ? 1ULL + c : 1ULL + E.leaf_offset + c + ib + krank) * 8ULL;
#pragma unroll
                    for (int q = 0; q < 6; q++) csum[q] = fadd(csum[q], rec[2 + q]);
                }                           // binary leaves cache zeros (OctreeBuilder.cpp:137-141)
            } else {
#pragma unroll
                for (int q = 0; q < 6; q++) csum[q] = fadd(csum[q], C.cache[c * 6 + q]);
            }
        }
        float cc[6];
        finish_average(csum, cn, cc);
        const unsigned long long cend = fc + __popcll(W & lowmask(8 * (k + 1)));
        const unsigned long long leaves_through = level == 0 ? cend : C.pl[cend];
        const unsigned long long rank = ib + (level == 0 ? 0ULL : C.pi[cend] - C.pi[fc]) + krank;
        if (!cache_only) write_data_record(data, internal_data_index(E, leaves_through, rank), cc);
        wn = fadd(wn, 1.0f);
#pragma unroll
        for (int q = 0; q < 6; q++) wsum[q] = fadd(wsum[q], cc[q]);
    }
    float wc[6];
    finish_average(wsum, wn, wc);
#pragma unroll
    for (int q = 0; q < 6; q++) L.cache[i * 6 + q] = wc[q];
    if (real_node && !cache_only) write_data_record(data, internal_data_index(E, L.pl[i + 1], ib + (L.pi[i + 1] - L.pi[i]) - 1ULL), wc);
}

// ---------------------------------------------------------------------------
// Fused single-block kernels for the SMALL upper levels (a few thousand tiles in total): one launch walks
// all of them instead of ~9 launches per level. Levels jf..J (jf >= 1) qualify when their tile counts are
// small; the big levels below keep the multi-block kernels.
// ---------------------------------------------------------------------------
struct FusedJob {
    Level lv[MAX_LEVELS];
    const unsigned long long* dense[MAX_LEVELS];   // biased dense levels (indexed by global word index)
    const unsigned long long* dense_top;           // unbiased dense level J
    unsigned long long top_words, top_bias;
    int J, jf;                                     // k_fused_emit / k_fused_down: levels above jf
    int jf_up;                                     // fused top kernel: subtree sizes of levels jf_up..J
    EmitJob E;
};

__device__ __forceinline__ void block_scan_array(const unsigned long long* mask, unsigned long long n, unsigned long long* out) {
    unsigned long long carry = 0;
    for (unsigned long long b = 0; b < n; b += blockDim.x) {
        const unsigned long long idx = b + threadIdx.x;
        const unsigned long long v = idx < n ? (unsigned long long)__popcll(mask[idx]) : 0ULL;
        unsigned long long total;
        const unsigned long long ex = block_excl_scan(v, total);
        if (idx < n) out[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
    __syncthreads();
}

// top-down: compact the top level, then (child prefix, expand) for levels J..jf; leaves fc[jf] ready so
// that the regular k_expand can produce level jf-1.
__global__ void __launch_bounds__(1024) k_fused_down(FusedJob F) {
    {   // compact dense level J
        unsigned long long carry = 0;
        for (unsigned long long b = 0; b < F.top_words; b += blockDim.x) {
            const unsigned long long idx = b + threadIdx.x;
            const unsigned long long w = idx < F.top_words ? F.dense_top[idx] : 0ULL;
            unsigned long long total;
            const unsigned long long ex = block_excl_scan(w != 0ULL ? 1ULL : 0ULL, total);
            if (w != 0ULL) { F.lv[F.J].key[carry + ex] = F.top_bias + idx; F.lv[F.J].mask[carry + ex] = w; }
            carry += total;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = F.J; j >= F.jf; j--) {
        const Level P = F.lv[j];
        block_scan_array(P.mask, P.n, P.fc);
        if (j > F.jf) {
            const Level C = F.lv[j - 1];
            for (unsigned long long i = wid; i < P.n; i += nw) {
                const unsigned long long W = P.mask[i], key = P.key[i], fc = P.fc[i];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int bit = lane + 32 * h;
                    if ((W >> bit) & 1ULL) {
                        const unsigned long long c = fc + __popcll(W & lowmask(bit));
                        const unsigned long long ck = (key << 6) | (unsigned long long)bit;
                        C.key[c] = ck;
                        C.mask[c] = F.dense[j - 1][ck];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// bottom-up: subtree-size prefixes of levels jf..J (ps of level jf-1 must be complete)
__device__ __forceinline__ void fused_up_body(const FusedJob& F, int jf) {
    for (int j = jf; j <= F.J; j++) {
        const Level L = F.lv[j];
        const unsigned long long n = level_n(L);
        const unsigned long long* cps = F.lv[j - 1].ps;
        unsigned long long carry = 0;
        for (unsigned long long b = 0; b < n; b += blockDim.x) {
            const unsigned long long idx = b + threadIdx.x;
            unsigned long long v = 0;
            if (idx < n) {
                const unsigned long long w = L.mask[idx];
                v = (unsigned long long)(__popcll(w) + __popc(nonzero_bytes(w))) + cps[L.fc[idx + 1]] - cps[L.fc[idx]];
            }
            unsigned long long total;
            const unsigned long long ex = block_excl_scan(v, total);
            if (idx < n) L.ps[idx] = carry + ex;
            carry += total;
        }
        if (threadIdx.x == 0) L.ps[n] = carry;
        if (L.pl) {                                   // leaf-count prefix (sharded table / global leaf ranks)
            const unsigned long long* cpl = F.lv[j - 1].pl;
            unsigned long long lc = 0;
            for (unsigned long long b = 0; b < n; b += blockDim.x) {
                const unsigned long long idx = b + threadIdx.x;
                const unsigned long long v = idx < n ? cpl[L.fc[idx + 1]] - cpl[L.fc[idx]] : 0ULL;
                unsigned long long total;
                const unsigned long long ex = block_excl_scan(v, total);
                if (idx < n) L.pl[idx] = lc + ex;
                lc += total;
            }
            if (threadIdx.x == 0) L.pl[n] = lc;
        }
        __threadfence();
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_fused_up(FusedJob F) {
    if (build_aborted(F.E.info)) return;
    fused_up_body(F, F.jf);
}

// top-down emission of levels J..jf+1 (each level writes the bases of the next); level jf itself is emitted
// by the regular multi-block kernel afterwards. No -levels on this path.
__device__ __forceinline__ void fused_emit_body(const FusedJob& F) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const NodeRange R = node_range(F.E);
    for (int j = F.J; j > F.jf; j--) {
        const Level L = F.lv[j], C = F.lv[j - 1];
        const unsigned long long n = level_n(L);
        EmitJob E = F.E;
        E.root_here = (j == F.J) ? F.E.root_here : 0;
        for (unsigned long long i = wid; i < n; i += nw) emit_upper_tile_fast(L, C, E, R, i, lane);      // (no -levels on this path)
        __threadfence();
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_fused_emit(FusedJob F) {
    if (build_aborted(F.E.info)) return;
    fused_emit_body(F);
}

// sharded: the records of the shared upper levels are computed on the host from the exchanged table
// (a few thousand at most) and scattered into this rank's part of the node array
__device__ __forceinline__ void scatter_records_body(const unsigned long long* pos, const unsigned long long* rec, unsigned long long n,
                                                     const unsigned long long* np, const EmitJob& E, unsigned long long first, unsigned long long stride) {
    if (np) { const unsigned long long v = *np; n = v < n ? v : n; }
    const NodeRange R = node_range(E);
    for (unsigned long long i = first; i < n; i += stride) {
        if (unsigned long long* o = node_slot(E, R, pos[i])) { o[0] = rec[3 * i]; o[1] = rec[3 * i + 1]; o[2] = rec[3 * i + 2]; }
    }
}
__global__ void __launch_bounds__(256) k_scatter_records(const unsigned long long* pos, const unsigned long long* rec, unsigned long long n,
                                                         const unsigned long long* np, EmitJob E) {
    if (build_aborted(E.info)) return;
    scatter_records_body(pos, rec, n, np, E, (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, (unsigned long long)gridDim.x * blockDim.x);
}

// sharded -levels: the data records of the shared upper levels' internal nodes, computed by the host-side merge
__global__ void __launch_bounds__(256) k_scatter_data_records(const unsigned long long* pos, const float* rec, unsigned long long n,
                                                              float* data /* biased by -data_lo records */) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    write_data_record(data, pos[i], rec + 6 * i);
}

// sparse clear of all levels in one launch (blockIdx.y = level)
struct ClearJob {
    const unsigned long long* key[MAX_LEVELS]; unsigned long long* dense[MAX_LEVELS]; unsigned long long n[MAX_LEVELS];
    const BuildInfo* info;            // fast path: the counts live on the device (n[] = capacities); an aborted build keeps its pyramid
};
__global__ void __launch_bounds__(256) k_sparse_clear_all(ClearJob Cj) {
    if (build_aborted(Cj.info)) return;
    const int j = blockIdx.y;
    unsigned long long n = Cj.n[j];
    if (Cj.info) { const unsigned long long v = Cj.info->count[j]; n = v < n ? v : n; }
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        Cj.dense[j][Cj.key[j][i]] = 0ULL;
}

// ascending Morton codes of the filled voxels
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_voxel_codes(Level L, unsigned long long* codes, unsigned long long capacity) {
    const unsigned long long i = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= level_n(L)) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long W = L.mask[i], key = L.key[i], fc = L.fc[i];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bit = lane + 32 * h;
        if ((W >> bit) & 1ULL) {
            const unsigned long long r = fc + __popcll(W & lowmask(bit));
            if (r < capacity) codes[r] = (key << 6) | (unsigned long long)bit;
        }
    }
}

// voxels per logical partition (what the reference prints per partition with -v, main.cpp:348): leaves of the bricks
// whose Morton prefix is the partition index. sh < 0: a partition is smaller than a brick (not reachable: P <= 8^5, g >= 4 side).
__global__ void __launch_bounds__(256) k_partition_voxels(const unsigned long long* key, const unsigned long long* mask, unsigned long long n, int sh,
                                                         unsigned long long* counts) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long p = sh >= 0 ? key[i] >> sh : 0ULL;
    atomicAdd(&counts[p], (unsigned long long)__popcll(mask[i]));
}

// zero exactly the words that were set, so the next run starts from a clean pyramid
__global__ void __launch_bounds__(256) k_sparse_clear(const unsigned long long* key, unsigned long long n, unsigned long long* dense) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dense[key[i]] = 0ULL;
}

// ---------------------------------------------------------------------------
// Payload: one data record per leaf, in leaf (= Morton) order
// (voxelizer.cpp:293-299, BarycentricCoords.h:4-33, main.cpp:371-384).
// ---------------------------------------------------------------------------
struct PayloadJob {
    const float* tris;        // 21-float records
    TriSegs segs;             // remote staging: the owner's record is read from the slice that holds it
    const uint32_t* owner;    // per leaf
    float* data;              // n_data * 8 floats (32-byte records)
    float unit_div;
    float gridsize_f;
    int color_mode;
    int levels;               // -levels: leaf records are interleaved with internal ones
    unsigned long long leaf_offset;   // sharding: leaves of lower ranks; `data` is biased accordingly
};

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_payload(Level L, PayloadJob Pj) {
    __shared__ unsigned long long s_base[MAX_WORLD + 1];
    if (Pj.segs.n) segs_bases(Pj.segs, s_base);
    const unsigned long long i = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= level_n(L)) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long W = L.mask[i], key = L.key[i], fc = L.fc[i];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bit = lane + 32 * h;
        if (!((W >> bit) & 1ULL)) continue;
        const unsigned long long r = fc + __popcll(W & lowmask(bit));
        const unsigned long long m = (key << 6) | (unsigned long long)bit;
        const uint32_t cx = compact3(m), cy = compact3(m >> 1), cz = compact3(m >> 2);
        const float* t = Pj.segs.n ? segs_record(Pj.segs, s_base, Pj.owner[r], 21) : Pj.tris + (size_t)Pj.owner[r] * 21;
        float v[21];
#pragma unroll
        for (int q = 0; q < 21; q++) v[q] = __ldg(t + q);
        // n = normalize(cross(e0, e1)) exactly as in the voxelizer (voxelizer.cpp:207-210)
        const float e0x = fsub(v[3], v[0]), e0y = fsub(v[4], v[1]), e0z = fsub(v[5], v[2]);
        const float e1x = fsub(v[6], v[3]), e1y = fsub(v[7], v[4]), e1z = fsub(v[8], v[5]);
        const float crx = fsub(fmul(e0y, e1z), fmul(e1y, e0z));
        const float cry = fsub(fmul(e0z, e1x), fmul(e1z, e0x));
        const float crz = fsub(fmul(e0x, e1y), fmul(e1x, e0y));
        const float inv = fdiv(1.0f, fsqrt(dot3(crx, cry, crz, crx, cry, crz)));
        const float nx = fmul(crx, inv), ny = fmul(cry, inv), nz = fmul(crz, inv);
        // ComputeBarycentricCoords(t, n, x / unit_div, y / unit_div, z / unit_div)
        const float vx = fdiv((float)cx, Pj.unit_div), vy = fdiv((float)cy, Pj.unit_div), vz = fdiv((float)cz, Pj.unit_div);
        const float coeffD = -dot3(v[0], v[1], v[2], nx, ny, nz);
        const float kk = fdiv(fadd(dot3(vx, vy, vz, nx, ny, nz), coeffD), fadd(fadd(fmul(nx, nx), fmul(ny, ny)), fmul(nz, nz)));
        const float ptx = fsub(vx, fmul(kk, nx)), pty = fsub(vy, fmul(kk, ny)), ptz = fsub(vz, fmul(kk, nz));
        // inverse(mat3(v0, v1, v2)) * point, glm compute_inverse<3,3>; m[c][r] = v[3c + r]
#define M_(c, r) v[3 * (c) + (r)]
        const float ood = fdiv(1.0f,
            fadd(fsub(fmul(M_(0, 0), fsub(fmul(M_(1, 1), M_(2, 2)), fmul(M_(2, 1), M_(1, 2)))),
                      fmul(M_(1, 0), fsub(fmul(M_(0, 1), M_(2, 2)), fmul(M_(2, 1), M_(0, 2))))),
                 fmul(M_(2, 0), fsub(fmul(M_(0, 1), M_(1, 2)), fmul(M_(1, 1), M_(0, 2))))));
        const float i00 = fmul(fsub(fmul(M_(1, 1), M_(2, 2)), fmul(M_(2, 1), M_(1, 2))), ood);
        const float i10 = fmul(-fsub(fmul(M_(1, 0), M_(2, 2)), fmul(M_(2, 0), M_(1, 2))), ood);
        const float i20 = fmul(fsub(fmul(M_(1, 0), M_(2, 1)), fmul(M_(2, 0), M_(1, 1))), ood);
        const float i01 = fmul(-fsub(fmul(M_(0, 1), M_(2, 2)), fmul(M_(2, 1), M_(0, 2))), ood);
        const float i11 = fmul(fsub(fmul(M_(0, 0), M_(2, 2)), fmul(M_(2, 0), M_(0, 2))), ood);
        const float i21 = fmul(-fsub(fmul(M_(0, 0), M_(2, 1)), fmul(M_(2, 0), M_(0, 1))), ood);
        const float i02 = fmul(fsub(fmul(M_(0, 1), M_(1, 2)), fmul(M_(1, 1), M_(0, 2))), ood);
        const float i12 = fmul(-fsub(fmul(M_(0, 0), M_(1, 2)), fmul(M_(1, 0), M_(0, 2))), ood);
        const float i22 = fmul(fsub(fmul(M_(0, 0), M_(1, 1)), fmul(M_(1, 0), M_(0, 1))), ood);
#undef M_
        const float b0 = fadd(fadd(fmul(i00, ptx), fmul(i10, pty)), fmul(i20, ptz));
        const float b1 = fadd(fadd(fmul(i01, ptx), fmul(i11, pty)), fmul(i21, ptz));
        const float b2 = fadd(fadd(fmul(i02, ptx), fmul(i12, pty)), fmul(i22, ptz));
        float col[3];
#pragma unroll
        for (int q = 0; q < 3; q++)   // InterpolateValue: b.x * c0 + b.y * c1 + b.z * c2
            col[q] = fadd(fadd(fmul(b0, v[12 + q]), fmul(b1, v[15 + q])), fmul(b2, v[18 + q]));
        const float fnx = v[9], fny = v[10], fnz = v[11];     // t.normal from the file (voxelizer.cpp:299)
        if (Pj.color_mode == 1) {                             // fixed colour, main.cpp:373-374
            col[0] = col[1] = col[2] = 1.0f;
        } else if (Pj.color_mode == 2) {                      // mortonToRGB, svo_builder_util.h:14-18
            col[0] = fdiv((float)cz, Pj.gridsize_f);
            col[1] = fdiv((float)cy, Pj.gridsize_f);
            col[2] = fdiv((float)cx, Pj.gridsize_f);
        } else if (Pj.color_mode == 3) {                      // main.cpp:379-381
            const float ninv = fdiv(1.0f, fsqrt(dot3(fnx, fny, fnz, fnx, fny, fnz)));
            col[0] = fdiv(fadd(fmul(fnx, ninv), 1.0f), 2.0f);
            col[1] = fdiv(fadd(fmul(fny, ninv), 1.0f), 2.0f);
            col[2] = fdiv(fadd(fmul(fnz, ninv), 1.0f), 2.0f);
        }
        // x86 SSE produces the "default NaN" 0xFFC00000 for invalid operations (0*inf, inf-inf, 0/0:
        // singular vertex matrix, zero normal) and propagates it; CUDA produces 0x7FFFFFFF.
#pragma unroll
        for (int q = 0; q < 3; q++) if (col[q] != col[q]) col[q] = __uint_as_float(0xFFC00000u);
        unsigned long long didx = 1ULL + Pj.leaf_offset + r;
        if (Pj.levels) didx += L.ibase[i] + __popc(nonzero_bytes(W) & ((1u << (bit >> 3)) - 1u));
        float4* o = reinterpret_cast<float4*>(Pj.data + didx * 8ULL);
        float4 a, b;
        a.x = __uint_as_float((uint32_t)(m & 0xffffffffULL));
        a.y = __uint_as_float((uint32_t)(m >> 32));
        a.z = col[0]; a.w = col[1];
        b.x = col[2]; b.y = fnx; b.z = fny; b.w = fnz;
        o[0] = a; o[1] = b;
    }
}

}  // namespace svo
