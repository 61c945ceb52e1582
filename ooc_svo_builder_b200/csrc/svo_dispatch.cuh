// svo_dispatch.cuh -- triangle dispatch of the sharded (multi-GPU) build: an all-to-all of triangle
// records written straight into the peers' HBM over NVLink / NVSwitch (peer stores), no host staging.
//
// Every rank starts with a contiguous 1/N slice of the .tridata file (TriReader reads the file front to
// back, src/libs/libtri/include/TriReader.h:41-79; rank r's slice is what it would read r-th). A triangle
// is needed by every rank whose slab its bounding box touches -- the partitioner's inclusive box test
// (partitioner.cpp:117-126, intersection.h:50-53) lifted from logical partitions to rank slabs; the test
// here is a superset test, the voxelizer on the receiving rank still applies the exact per-partition rule.
//
//   pass 1  k_dispatch_count   per staging block (128 consecutive local triangles) and destination: count
//           exscan             block offsets per destination (hand-written scan, svo_kernels.cuh)
//           k_dispatch_post    row of the count matrix M[src][dst] -> every peer's control block, flag A
//   pass 2  k_dispatch_wait    until all rows arrived (flag A of every peer)
//           k_dispatch_write   records -> peer inbox at  sum_{s < me} M[s][dst] + block offset  (stable:
//                              the inbox is ordered by (source rank, local index) = file order, so the
//                              payload rule "first triangle in file order wins" (voxelizer.cpp:263) holds
//                              on inbox positions)
//           k_dispatch_post    flag B
//   finish  k_dispatch_wait    until every peer's records landed (flag B), then the voxelizer runs on the inbox
//
// Flags are epochs (monotonic), written with a system-scope fence behind the data they publish.
#pragma once
#include "svo_kernels.cuh"

namespace svo {

// One per context, in cudaMalloc'd memory that peers map (IPC or same-process pointers).
struct DispatchCtrl {
    unsigned long long flag[2][MAX_WORLD];             // [phase][source rank] = last epoch that source finished
    unsigned long long matrix[MAX_WORLD][MAX_WORLD];   // M[src][dst] triangle counts of the current epoch
    unsigned long long error;                          // set by a wait that timed out
};

struct DispatchJob {
    const float* tris; uint32_t fpt; unsigned long long n_local;
    int world, me;
    int use_partitions, k;
    float bmin[32], bmax[32];
    int lo[MAX_WORLD][3], hi[MAX_WORLD][3];     // destination boxes: partition coordinates (use_partitions) or voxels
    float lof[MAX_WORLD][3], hif[MAX_WORLD][3]; // use_partitions: bmin[lo], bmax[hi] -- the two world-space bounds the test needs
    float unit_div; int gmax;
    unsigned int* blockcnt;                     // [world][nb]
    const unsigned long long* blockoff;         // exclusive scan of blockcnt, [world * nb + 1]
    unsigned long long nb;
    float* inbox[MAX_WORLD];                    // peer inboxes
    DispatchCtrl* ctrl[MAX_WORLD];              // peer control blocks (ctrl[me] = own)
    unsigned long long epoch;
};

// Destination ranks of one triangle as a bit mask (superset test on bounding boxes, as k_owner_filter).
__device__ __forceinline__ unsigned dispatch_mask(const DispatchJob& D, const float* c) {
    unsigned m = (1u << D.world) - 1u;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float mn = stdmin(c[a], stdmin(c[3 + a], c[6 + a])), mx = stdmax(c[a], stdmax(c[3 + a], c[6 + a]));
        if (D.use_partitions) {
            // the kept slab range [L, H] of the interval meets the destination's slabs [lo, hi] iff H >= lo and L <= hi; by
            // monotonicity of the slab tables that is !(mx < bmin[lo]) and !(mn > bmax[hi]) (intersection.h:50-53 per axis)
            for (int d = 0; d < D.world; d++)
                if ((mx < D.lof[d][a]) || (mn > D.hif[d][a])) m &= ~(1u << d);
        } else {
            const int l = clampi(f2i(fmul(mn, D.unit_div)), 0, D.gmax), h = clampi(f2i(fmul(mx, D.unit_div)), 0, D.gmax);
            for (int d = 0; d < D.world; d++)
                if (h < D.lo[d][a] || l > D.hi[d][a]) m &= ~(1u << d);
        }
    }
    return m;
}

__global__ void __launch_bounds__(VOX_BLOCK) k_dispatch_count(DispatchJob D) {
    const unsigned long long t = (unsigned long long)blockIdx.x * VOX_BLOCK + threadIdx.x;
    unsigned m = 0;
    if (t < D.n_local) {
        const float* v = D.tris + t * D.fpt;
        float c[9];
#pragma unroll
        for (int i = 0; i < 9; i++) c[i] = __ldg(v + i);
        m = dispatch_mask(D, c);
    }
    for (int d = 0; d < D.world; d++) {
        const int n = __syncthreads_count((m >> d) & 1u);
        if (threadIdx.x == 0) D.blockcnt[(unsigned long long)d * D.nb + blockIdx.x] = (unsigned)n;
    }
}

// One block, one thread per peer: publish (phase 0) this rank's row of the count matrix, then raise the flag.
__global__ void __launch_bounds__(MAX_WORLD) k_dispatch_post(DispatchJob D, int phase) {
    const int p = threadIdx.x;
    if (p >= D.world) return;
    if (phase == 0) {
        for (int d = 0; d < D.world; d++) {
            const unsigned long long n = D.nb ? D.blockoff[(unsigned long long)(d + 1) * D.nb] - D.blockoff[(unsigned long long)d * D.nb] : 0ULL;
            *(volatile unsigned long long*)&D.ctrl[p]->matrix[D.me][d] = n;
        }
    }
    __threadfence_system();
    *(volatile unsigned long long*)&D.ctrl[p]->flag[phase][D.me] = D.epoch;
}

// One block, one thread per peer: spin until that peer's flag reaches the epoch. Gives up after ~4 s of GPU clock
// (a peer that never arrives must not hang the device) and records the failure in ctrl->error.
__global__ void __launch_bounds__(MAX_WORLD) k_dispatch_wait(DispatchCtrl* own, int world, int phase, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p >= world) return;
    const volatile unsigned long long* f = &own->flag[phase][p];
    const long long t0 = clock64();
    while (*f < epoch) {
        if (clock64() - t0 > 8000000000LL) { own->error = 1ULL + (unsigned long long)phase; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// Pass 2: staging block -> per destination, the selected records compacted in shared memory (stable) and
// streamed to the peer inbox as one contiguous run of floats (coalesced NVLink stores).
__global__ void __launch_bounds__(VOX_BLOCK) k_dispatch_write(DispatchJob D, unsigned long long inbox_cap) {
    extern __shared__ float4 s_dyn4[];
    float* s_in = reinterpret_cast<float*>(s_dyn4);
    float* s_out = s_in + (size_t)VOX_BLOCK * D.fpt;
    __shared__ unsigned s_warp[VOX_BLOCK / 32];
    const unsigned long long q0 = (unsigned long long)blockIdx.x * VOX_BLOCK;
    const unsigned long long nrec = D.n_local - q0 < VOX_BLOCK ? D.n_local - q0 : VOX_BLOCK;
    const unsigned long long nfl = nrec * D.fpt;
    {   // stage the block's records (q0 * fpt * 4 bytes is a multiple of 16)
        const float* src = D.tris + q0 * D.fpt;
        const unsigned long long n4 = nfl >> 2;
        const float4* src4 = reinterpret_cast<const float4*>(src);
        for (unsigned long long i = threadIdx.x; i < n4; i += VOX_BLOCK) s_dyn4[i] = __ldg(src4 + i);
        for (unsigned long long i = (n4 << 2) + threadIdx.x; i < nfl; i += VOX_BLOCK) s_in[i] = __ldg(src + i);
    }
    __syncthreads();
    unsigned m = 0;
    if (threadIdx.x < nrec) m = dispatch_mask(D, s_in + (size_t)threadIdx.x * D.fpt);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const DispatchCtrl* own = D.ctrl[D.me];
    for (int d = 0; d < D.world; d++) {
        const unsigned cnt = D.blockcnt[(unsigned long long)d * D.nb + blockIdx.x];
        if (cnt == 0) continue;                                  // block-uniform
        unsigned long long at = D.blockoff[(unsigned long long)d * D.nb + blockIdx.x] - D.blockoff[(unsigned long long)d * D.nb];
        for (int s = 0; s < D.me; s++) at += own->matrix[s][d];
        if (at + cnt > inbox_cap) {                              // block-uniform: never write past a peer's inbox
            if (threadIdx.x == 0) *(volatile unsigned long long*)&D.ctrl[D.me]->error = 3ULL;
            continue;
        }
        float* dst = D.inbox[d] + at * D.fpt;
        const float* from = s_in;
        if (cnt != nrec) {
            // stable compaction of the selected records
            const bool sel = (m >> d) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, sel);
            if (lane == 0) s_warp[wid] = __popc(b);
            __syncthreads();
            unsigned pos = __popc(b & ((1u << lane) - 1u));
            for (int w = 0; w < wid; w++) pos += s_warp[w];
            if (sel) {
                const float* r = s_in + (size_t)threadIdx.x * D.fpt;
                float* o = s_out + (size_t)pos * D.fpt;
                for (uint32_t i = 0; i < D.fpt; i++) o[i] = r[i];
            }
            __syncthreads();
            from = s_out;
        }
        const unsigned total = cnt * D.fpt;
        for (unsigned i = threadIdx.x; i < total; i += VOX_BLOCK) dst[i] = from[i];
        __syncthreads();                                         // s_out / s_warp are reused by the next destination
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------
// Remote staging ("slices"): nothing is copied. Rank s lists, for every destination rank, the staging blocks of
// ITS slice that touch the destination's slab and stores the list into the destination's list buffer (4 bytes per
// block over NVLink); the destination's voxelizer then stages exactly those blocks from rank s's HBM with NVLink
// loads, overlapped with the math of its other resident blocks (k_vox_small, SUBSET + segs).
//   flag[0][s] = epoch : rank s's lists and slice size for this job are published
//   flag[1][s] = epoch : rank s has finished reading its peers' slices and lists (they may be rewritten)
// ---------------------------------------------------------------------------
struct SliceCtrl {
    unsigned long long flag[3][MAX_WORLD];             // [0] lists published, [1] done reading, [2] table entries pushed
    unsigned long long count[MAX_WORLD][MAX_WORLD];    // count[src][dst]: blocks of slice src listed for dst (row dst's own column is what it reads)
    unsigned long long nslice[MAX_WORLD];              // triangles in slice s
    unsigned long long error;
};

struct SliceJob {
    DispatchJob D;                              // geometry + local slice (tris, fpt, n_local, nb); inbox / blockcnt unused
    uint32_t* list[MAX_WORLD];                  // peer list buffers: region [src * cap, (src + 1) * cap) belongs to source src
    SliceCtrl* ctrl[MAX_WORLD];
    unsigned long long cap;                     // list entries (32-triangle units) per region
    unsigned long long* cursor;                 // local: MAX_WORLD list cursors + [MAX_WORLD] the count of finished blocks (all zeroed by the posting block)
    float4* ubox;                               // per 32-triangle unit of the slice, two float4: {min x, y, z, max x}, {max y, z, largest |coordinate|, NaN seen}
    int ubox_mode;                              // 0: test the triangles (no cache), 1: fill the cache, nothing else (k_slice_boxes), 2: test the cached boxes
};
constexpr int FILTER_WARPS = 8;
// order-preserving float <-> int maps (for the integer warp reductions): a < b  <=>  ord(a) < ord(b) for non-NaN floats
__device__ __forceinline__ int float_to_ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ord_to_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
// One WARP per 32-triangle unit, one triangle per lane (nine direct loads of the lane's vertices: a unit is 1152 / 2688
// contiguous bytes, every line is used by the warp), several units in flight. One test per UNIT and destination instead of
// one per triangle: the unit's bounding box (six warp reductions) against the destination box, lane d testing destination
// d. A superset of the union of the per-triangle tests (per axis it is exactly their union), which is all the list needs:
// the voxelizer on the destination applies the exact per-partition rule to every triangle. Units with a NaN or an
// absurdly large coordinate (float -> int conversion no longer monotone) go to every destination. Every warp owns a
// contiguous run of units and reserves list space once per 64 of them (lane d keeps the hit mask of destination d), so the
// list cursors see a fraction of the atomics a per-unit append would issue.
// Bounding box of unit `u` of a slice (all 32 lanes; lane = triangle). any_odd: a triangle of the unit has a NaN or an
// absurdly large coordinate. With `rec` != NULL lane 0 also writes the unit's cache record: the largest |coordinate| of the
// unit decides the "absurdly large" test for ANY unit_div (|x| * unit_div is monotone in |x|), NaNs are flagged.
__device__ __forceinline__ void unit_box(const float* tris, uint32_t fpt, unsigned long long n, unsigned long long u, int lane, float unit_div,
                                         float (&umn)[3], float (&umx)[3], bool& any_odd, float4* rec) {
    const unsigned long long t = u * UNIT + lane;
    float mn[3], mx[3];
    bool odd = false, nan = false;
    float big = 0.0f;
    if (t < n) {
        const float* c = tris + t * fpt;
        float v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = __ldg(c + i);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            mn[a] = stdmin(v[a], stdmin(v[3 + a], v[6 + a])); mx[a] = stdmax(v[a], stdmax(v[3 + a], v[6 + a]));
            odd = odd || !(fabsf(fmul(mn[a], unit_div)) < 1.0e9f) || !(fabsf(fmul(mx[a], unit_div)) < 1.0e9f);   // NaN compares false
            nan = nan || mn[a] != mn[a] || mx[a] != mx[a];
            if (mn[a] == mn[a]) big = fmaxf(big, fabsf(mn[a]));
            if (mx[a] == mx[a]) big = fmaxf(big, fabsf(mx[a]));
        }
    } else {
#pragma unroll
        for (int a = 0; a < 3; a++) { mn[a] = __int_as_float(0x7f800000); mx[a] = __int_as_float(0xff800000); }
    }
    any_odd = __any_sync(0xffffffffu, odd);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        umn[a] = ord_to_float(__reduce_min_sync(0xffffffffu, float_to_ord(mn[a])));
        umx[a] = ord_to_float(__reduce_max_sync(0xffffffffu, float_to_ord(mx[a])));
    }
    if (rec) {
        const bool any_nan = __any_sync(0xffffffffu, nan);
        const float ubig = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(big)));     // big >= 0: the bit patterns order like the values
        if (lane == 0) {
            rec[0] = make_float4(umn[0], umn[1], umn[2], umx[0]);
            rec[1] = make_float4(umx[1], umx[2], ubig, any_nan ? 1.0f : 0.0f);
        }
    }
}
// Fills the box cache of a slice: once per LOAD of the slice, not per job (svo_shard_slice_publish launches it when the slice
// has changed since the last time).
__global__ void __launch_bounds__(FILTER_WARPS * 32) k_slice_boxes(SliceJob S) {
    const int lane = threadIdx.x & 31;
    const unsigned long long n = S.D.n_local, n_units = (n + UNIT - 1) / UNIT;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * FILTER_WARPS;
    for (unsigned long long u = (unsigned long long)blockIdx.x * FILTER_WARPS + (threadIdx.x >> 5); u < n_units; u += nwarps) {
        float umn[3], umx[3];
        bool any_odd;
        unit_box(S.D.tris, S.D.fpt, n, u, lane, S.D.unit_div, umn, umx, any_odd, S.ubox + 2 * u);
    }
}
__global__ void __launch_bounds__(FILTER_WARPS * 32) k_slice_filter(SliceJob S) {
    __shared__ int s_last;
    // (The wait for the peers to have finished reading the previous job's lists stays a ONE-block kernel in front of this
    // one, k_slice_wait: a whole grid that spins for a peer can fill the GPU, and with several contexts per device -- two
    // steps in flight, or all ranks of a test on one GPU -- the kernel it is waiting for might never get an SM.)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t fpt = S.D.fpt;
    const unsigned long long n = S.D.n_local;
    const unsigned long long n_units = (n + UNIT - 1) / UNIT;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * FILTER_WARPS;
    const unsigned long long per = (n_units + nwarps - 1) / nwarps;
    const unsigned long long gw = (unsigned long long)blockIdx.x * FILTER_WARPS + wid;
    const unsigned long long u1 = min(n_units, (gw + 1) * per);
    // lane d tests destination d: its box, once, in registers
    const int dd = lane < S.D.world ? lane : 0;
    float d_lof[3], d_hif[3];
    int d_lo[3], d_hi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { d_lof[a] = S.D.lof[dd][a]; d_hif[a] = S.D.hif[dd][a]; d_lo[a] = S.D.lo[dd][a]; d_hi[a] = S.D.hi[dd][a]; }
    for (unsigned long long base = gw * per; base < u1; base += 64) {
        unsigned long long hits = 0;                            // lane d: bit j = unit base + j touches destination d
        const int nj = (int)min(64ULL, u1 - base);
        if (S.ubox_mode == 2) {
            // The units' boxes were computed when the slice was loaded (k_slice_boxes): they do not depend on the job. Here the
            // roles are turned around: a lane holds the boxes of two units (coalesced loads, nothing waits for anything), the
            // destinations are walked by the whole warp, and a ballot hands destination d's hit mask to lane d.
            float4 bx[2][2];
            bool valid[2], odd[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                valid[h] = lane + 32 * h < nj;
                bx[h][0] = bx[h][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[h]) { bx[h][0] = __ldg(S.ubox + 2 * (base + lane + 32 * h)); bx[h][1] = __ldg(S.ubox + 2 * (base + lane + 32 * h) + 1); }
                odd[h] = bx[h][1].w != 0.0f || !(fmul(bx[h][1].z, S.D.unit_div) < 1.0e9f);
            }
            for (int d = 0; d < S.D.world; d++) {
                unsigned m[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float umn[3] = { bx[h][0].x, bx[h][0].y, bx[h][0].z }, umx[3] = { bx[h][0].w, bx[h][1].x, bx[h][1].y };
                    bool touch = true;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        if (S.D.use_partitions) {
                            touch = touch && !(umx[a] < S.D.lof[d][a]) && !(umn[a] > S.D.hif[d][a]);
                        } else {
                            const int l = clampi(f2i(fmul(umn[a], S.D.unit_div)), 0, S.D.gmax), hh = clampi(f2i(fmul(umx[a], S.D.unit_div)), 0, S.D.gmax);
                            touch = touch && !(hh < S.D.lo[d][a] || l > S.D.hi[d][a]);
                        }
                    }
                    m[h] = __ballot_sync(0xffffffffu, valid[h] && (touch || odd[h]));
                }
                if (lane == d) hits = (unsigned long long)m[0] | ((unsigned long long)m[1] << 32);
            }
        }
#pragma unroll 4
        for (int j = 0; j < (S.ubox_mode == 2 ? 0 : nj); j++) {
            float umn_[3], umx_[3];
            bool any_odd;
            unit_box(S.D.tris, fpt, n, base + j, lane, S.D.unit_div, umn_, umx_, any_odd, nullptr);
            bool touch = true;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float umn = umn_[a], umx = umx_[a];
                if (lane < S.D.world) {
                    if (S.D.use_partitions) {
                        touch = touch && !(umx < d_lof[a]) && !(umn > d_hif[a]);
                    } else {
                        const int l = clampi(f2i(fmul(umn, S.D.unit_div)), 0, S.D.gmax), h = clampi(f2i(fmul(umx, S.D.unit_div)), 0, S.D.gmax);
                        touch = touch && !(h < d_lo[a] || l > d_hi[a]);
                    }
                }
            }
            if (lane < S.D.world && (touch || any_odd)) hits |= 1ULL << j;      // (a unit of the run holds at least one triangle)
        }
        if (lane < S.D.world && hits) {
            unsigned long long pos = atomicAdd(&S.cursor[lane], (unsigned long long)__popcll(hits));
            uint32_t* out = S.list[lane] + (unsigned long long)S.D.me * S.cap;
            while (hits) {
                const int j = __ffsll((long long)hits) - 1;
                hits &= hits - 1;
                out[pos++] = (uint32_t)(base + j);              // 32-triangle unit index; pos < n_units <= cap
            }
        }
    }
    // the block that finishes last publishes: counts + slice size to every peer, then flag 0 behind a system-scope fence
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&S.cursor[MAX_WORLD], 1ULL) == (unsigned long long)gridDim.x - 1ULL;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        const int p = threadIdx.x;
        if (p < S.D.world) {
            *(volatile unsigned long long*)&S.ctrl[p]->count[S.D.me][p] = *(volatile unsigned long long*)&S.cursor[p];
            *(volatile unsigned long long*)&S.ctrl[p]->nslice[S.D.me] = S.D.n_local;
            S.cursor[p] = 0ULL;
            __threadfence_system();
            *(volatile unsigned long long*)&S.ctrl[p]->flag[0][S.D.me] = S.D.epoch;
        }
        if (threadIdx.x == 0) S.cursor[MAX_WORLD] = 0ULL;
    }
}

// phase 0: publish counts + slice size, raise flag 0, reset the cursors. phase 1: raise flag 1 (done reading).
__global__ void __launch_bounds__(MAX_WORLD) k_slice_post(SliceJob S, int phase) {
    const int p = threadIdx.x;
    if (p >= S.D.world) return;
    if (phase == 0) {
        *(volatile unsigned long long*)&S.ctrl[p]->count[S.D.me][p] = S.cursor[p];
        *(volatile unsigned long long*)&S.ctrl[p]->nslice[S.D.me] = S.D.n_local;
        S.cursor[p] = 0ULL;
    }
    __threadfence_system();
    *(volatile unsigned long long*)&S.ctrl[p]->flag[phase][S.D.me] = S.D.epoch;
}

__global__ void __launch_bounds__(MAX_WORLD) k_slice_wait(SliceCtrl* own, int world, int phase, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p >= world) return;
    const volatile unsigned long long* f = &own->flag[phase][p];
    const long long t0 = clock64();
    while (*f < epoch) {
        if (clock64() - t0 > 8000000000LL) { own->error = 1ULL + (unsigned long long)phase; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// The one exchange of the sharded build -- the table of top-of-shard subtree records -- without a library collective:
// the entries of different ranks are disjoint contiguous ranges, so every rank stores its own range into every peer's
// exchange table (NVLink stores) and raises flag 2; after k_xchg_wait each rank holds the complete table.
// Flag value = 2 * exchange epoch + poison. A rank whose speculative local build was aborted (its lists outgrew the
// capacities of the previous build, BuildInfo::overflow) has no valid entries to offer: it raises the flag with the
// poison bit, every peer's wait marks its own build as "repeat" (overflow bit 43) and ALL ranks return SVO_E_RETRY
// from svo_shard_emit in the same step -- no host read-back is needed to keep the ranks in agreement.
struct XchgJob {
    const unsigned long long* src;              // local table (only [lo, lo + n) is read)
    unsigned long long lo, n;                   // own range, in u64
    unsigned long long* xtable[MAX_WORLD];
    SliceCtrl* ctrl[MAX_WORLD];
    int world, me;
    unsigned long long epoch;
    const BuildInfo* info;                      // NULL: the entries are always valid (sized builds)
};
__global__ void __launch_bounds__(256) k_xchg_push(XchgJob X) {
    const int p = blockIdx.x;                   // one block per destination
    const bool poisoned = build_aborted(X.info);
    unsigned long long* dst = X.xtable[p];
    if (!poisoned) for (unsigned long long i = threadIdx.x; i < X.n; i += blockDim.x) dst[X.lo + i] = X.src[X.lo + i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *(volatile unsigned long long*)&X.ctrl[p]->flag[2][X.me] = 2ULL * X.epoch + (poisoned ? 1ULL : 0ULL);
    }
}
__global__ void __launch_bounds__(MAX_WORLD) k_xchg_wait(SliceCtrl* own, int world, unsigned long long epoch, BuildInfo* info) {
    const int p = threadIdx.x;
    if (p >= world) return;
    const volatile unsigned long long* f = &own->flag[2][p];
    const long long t0 = clock64();
    unsigned long long v;
    while (((v = *f) >> 1) < epoch) {
        if (clock64() - t0 > 8000000000LL) { own->error = 3ULL; break; }
        __nanosleep(200);
    }
    if ((v >> 1) == epoch && (v & 1ULL) && info) atomicOr(&info->overflow, 1ULL << 43);
    __threadfence_system();
}

struct U32Op {
    const unsigned int* v;
    __device__ unsigned long long operator()(unsigned long long i) const { return (unsigned long long)v[i]; }
};

}  // namespace svo
