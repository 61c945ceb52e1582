// svo_api.cu -- C ABI (include/svo_b200.h) and host-side orchestration of the
// sm_100a kernels in svo_kernels.cuh. No CPU fallback: every compute entry
// point needs a compute-capability 10.x device.
#include "../../include/svo_b200.h"
#include "svo_kernels.cuh"
#include "svo_dispatch.cuh"
#include "svo_build.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

using namespace svo;
typedef unsigned long long ull;

namespace {

std::string g_create_error;          // svo_ctx_create failures (process-wide: the CLI creates the context on a helper thread)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;      // a little slack so steady-state runs never reallocate
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct LevelBufs {
    DevBuf key, mask, fc, ps, base, pi, pl, ibase, cache;
    ull n = 0;
    Level view() const {
        Level L;
        L.key = key.as<ull>(); L.mask = mask.as<ull>(); L.fc = fc.as<ull>(); L.ps = ps.as<ull>(); L.base = base.as<ull>();
        L.pi = pi.as<ull>(); L.pl = pl.as<ull>(); L.ibase = ibase.as<ull>(); L.cache = cache.as<float>();
        L.n = n; L.np = nullptr; L.cap = n; L.clear = nullptr;
        return L;
    }
};

constexpr size_t XTABLE_BYTES = 1u << 20;   // exchange table inside the slice window: up to 32768 top-of-shard entries

enum { EV_UP0, EV_UP1, EV_PART0, EV_PART1, EV_VOX0, EV_VOX1, EV_BUILD0, EV_EMIT0, EV_EMIT1, EV_BUILD1, EV_VS0, EV_VS1, EV_EL0, EV_EL1, EV_CMP1, EV_CLR0, EV_CLR1, EV_DN0, EV_DN1, EV_DSP0, EV_DSP1, EV_PW0, EV_PW1, EV_COUNT };

}  // namespace

namespace {
// SVO_TIMELINE=1: host wall-clock stamps (us since the stamp named "partition") printed at the end of every build
struct Timeline {
    bool on = false; int n = 0; const char* name[32]; double t[32];
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
    void stamp(const char* w) { if (on && n < 32) { name[n] = w; t[n++] = now(); } }
    void dump() { if (!on) return; for (int i = 0; i < n; i++) fprintf(stderr, "%s %.1f%s", name[i], t[i] - t[0], i + 1 < n ? " | " : "\n"); n = 0; }
};
}  // namespace

struct svo_ctx {
    Timeline tl;
    int device = 0;
    int sm_count = 148;
    int res_emit_leaf[2] = {4, 4}, res_emit_upper = 4;      // resident blocks per SM of the persistent emitters (occupancy API)
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::string err;
    cudaEvent_t ev[EV_COUNT];
    bool ev_set[EV_COUNT];
    ull* h_pinned = nullptr;          // 64 u64 of pinned scratch for small read-backs

    // triangles
    DevBuf tri_own;
    const float* d_tris = nullptr;
    uint64_t n_tris = 0;
    int fpt = 0;
    bool have_tris = false;
    uint64_t stream_fill = 0;          // svo_triangles_begin / _append: records copied so far
    cudaEvent_t up_ev[2] = { nullptr, nullptr };
    bool up_pending[2] = { false, false };
    int up_slot = 0;

    // job
    svo_params prm;
    bool partitioned = false, voxelized = false, built = false;
    int D = 0, nl = 0, k = 0;
    uint64_t P = 1;
    uint32_t side = 0;
    float unit_vox = 0, unit_div = 0;

    // dense pyramid
    DevBuf dense[MAX_LEVELS];
    ull nwords[MAX_LEVELS];
    uint64_t dense_grid = 0;          // gridsize the pyramid is allocated for
    bool dense_clean = false;
    DevBuf d_lvlptrs, d_nwords, d_counts;

    // partition lists
    DevBuf part_counts, part_cursor, part_off, pair_tri;
    uint64_t n_pairs = 0;
    std::vector<uint64_t> h_part_counts;

    // work queues
    DevBuf queue[2], qcount, subset;
    bool use_subset = false;
    bool warp_kernel = true;           // SVO_VOX_KERNEL=block selects the block-granular small-box voxelizer

    // compact levels
    LevelBufs lv[MAX_LEVELS];
    DevBuf lb_state, lb_ticket;
    ull lb_epoch = 0, lb_tickets = 0;

    // device-driven build (svo_build.cuh): counts live in a device-resident BuildInfo; list capacities are remembered
    // from the previous build so that a steady-state build needs no host read-back before the final one
    DevBuf info_buf, merge_scratch, merge_rpos, merge_rrec;
    DevBuf brick_lp, brick_sp;                 // per level-1 tile: leaf / brick-record prefixes (brick pass)
    BuildInfo* h_info = nullptr;       // pinned copy
    bool fast = false, spec = false, fast_caps_ok = false;
    int jB = 1;                        // levels 1..jB are scanned by k_dense_scan, jB+1..J by k_small_levels
    int jE = 1;                        // levels jE+1..J are emitted by the single-block top kernel
    int top_pending = 0;               // k_top stages of phase A whose launch was deferred (one GPU: to phase B; sharded with the
                                       // peer-memory exchange: to svo_shard_exchange, which adds the push of the table entries)
    bool xwait_pending = false;        // the wait for the peers' table entries is folded into phase B's k_top
    ull* xchg_user_table = nullptr;    // caller's table buffer: receives a copy of the complete table at the end of svo_shard_emit
    ull fcap[MAX_LEVELS];              // entries the tile lists of each level hold

    // outputs
    DevBuf nodes, data, owner, tileidx, codes;
    uint64_t n_voxels = 0, n_nodes = 0, n_data = 0;

    // sharding / two-phase build
    int world = 1, rank = 0;
    int dc = 0, J = 0;
    ull c0 = 0, c1 = 1, slab_start = 0, slab_end = 0, WJ = 1;
    ull bias[MAX_LEVELS];
    int sb_lo[3], sb_hi[3];
    ull q_begin = 0, q_end = 0, qcap = 1;
    bool want_pl = false, phase_a_done = false;
    int jf = 0;                        // fused single-block kernels handle local levels jf..J (0 = not fused)
    bool use_lists = false;            // SVO_PARTITION_LISTS=1: build per-partition index lists (count / scan / fill)
    float inv_slab = 0.f;
    float slab_min[32], slab_max[32];  // world slabs of the partition grid (partitioner.cpp:54-59)
    ull p_first = 0, p_last = 0;
    DevBuf table_own, dcol[4];
    LevelBufs glv;                     // global level-J tile list (sharded / odd depth)
    std::vector<ull> h_table, h_rpos, h_rrec, h_ownbase, h_ownibase, h_dpos;
    std::vector<float> h_drec;
    DevBuf d_rpos, d_rrec, d_dpos, d_drec, stage_data;
    bool owner_ready = false;           // sharded -levels: the owner pass already ran before the exchange
    int tstride = 4;                    // u64 per exchange-table entry: 4, or 8 with -levels (+ the tile's 6-float data cache)
    ull n_upper_records = 0;
    ull leaf_offset = 0, n_voxels_local = 0;
    ull node_lo = 0, node_hi = 0, data_lo = 0, data_hi = 0;

    // triangle dispatch over peer memory (svo_dispatch.cuh)
    DevBuf inbox, ctrl_buf, blockcnt, blockoff;
    uint64_t inbox_cap = 0;
    int inbox_fpt = 0;
    float* peer_inbox[MAX_WORLD];
    DispatchCtrl* peer_ctrl[MAX_WORLD];
    DispatchCtrl* h_ctrl = nullptr;   // pinned read-back copy
    bool attached = false, dispatched = false;
    int dispatch_phase = 0;           // 0 idle, 1 counted, 2 sent
    ull dispatch_epoch = 0;
    uint32_t dispatch_launches = 0;
    DispatchJob dj;

    // remote staging of triangle slices (svo_dispatch.cuh)
    DevBuf sl_ubox;                    // per-unit bounding boxes of this rank's slice (k_slice_boxes); valid until the slice changes
    bool sl_boxes_valid = false;
    DevBuf window, sl_cursor;          // window = [SliceCtrl | exchange table | block lists | slice], one allocation peers map
    struct View { void* p = nullptr; template <class T> T* as() const { return static_cast<T*>(p); } } slice, sl_list, sl_ctrl, sl_xtable;
    ull* peer_xtable[MAX_WORLD];
    uint64_t slice_cap = 0, sl_cap_blocks = 0, slice_n_local = 0;
    int slice_fpt = 0;
    float* peer_slice[MAX_WORLD];
    uint32_t* peer_list[MAX_WORLD];
    SliceCtrl* peer_slctrl[MAX_WORLD];
    bool sl_attached = false, sliced = false, filter_attr_set = false;
    ull sl_epoch = 0, xchg_epoch = 0;
    int sl_world = 0;                        // world size the window was laid out for
    bool exchanged_by_peer_memory = false;   // the table of this job went through svo_shard_exchange (poison-aware)
    bool uses_peer_exchange = false;         // ... and so did an earlier job of this context: local builds may be speculative
    SliceJob sj;

    svo_stats stats;
    uint32_t launches = 0;
};

namespace {

int fail(svo_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                              \
            return fail(c, e_ == cudaErrorMemoryAllocation ? SVO_E_NOMEM : SVO_E_CUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
        }                                                                                          \
    } while (0)

#define LAUNCHED()                                                                                 \
    do {                                                                                           \
        c->launches++;                                                                             \
        CK(cudaGetLastError());                                                                    \
    } while (0)



inline unsigned blocks_for(ull n, unsigned per) { return (unsigned)((n + per - 1) / per); }

void mark(svo_ctx* c, int e) {
    cudaEventRecord(c->ev[e], c->stream);
    c->ev_set[e] = true;
}
float span(svo_ctx* c, int a, int b) {
    float ms = 0.f;
    if (c->ev_set[a] && c->ev_set[b] && cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) == cudaSuccess) return ms;
    (void)cudaGetLastError();
    return 0.f;
}

// state of the single-pass scans: persistent, epoch-tagged (no clearing between launches)
int lookback_prepare(svo_ctx* c, ull nt) {
    const size_t need = (size_t)(nt + 1) * LB_STATE * sizeof(ull);
    if (need > c->lb_state.cap) {
        CK(c->lb_state.ensure(need));
        CK(cudaMemsetAsync(c->lb_state.p, 0, c->lb_state.cap, c->stream));
    }
    if (!c->lb_ticket.p) {
        CK(c->lb_ticket.ensure(sizeof(ull)));
        CK(cudaMemsetAsync(c->lb_ticket.p, 0, sizeof(ull), c->stream));
        c->lb_tickets = 0;
    }
    ++c->lb_epoch;                           // 62 bits: never wraps
    return SVO_OK;
}

constexpr ull SCAN_ONE_BLOCK_MAX = 4096;     // up to here one block is faster than the look-back chain

// Exclusive scan of f over [0, n) into out[0..n]. np != NULL: the count lives on the device (n is the capacity the grid is sized for).
template <class F>
int exscan(svo_ctx* c, F f, ull n, ull* out, const ull* np = nullptr, BuildInfo* info = nullptr) {
    if (n == 0) {
        CK(cudaMemsetAsync(out, 0, sizeof(ull), c->stream));
        return SVO_OK;
    }
    if (n <= SCAN_ONE_BLOCK_MAX) {
        k_scan_small<<<1, 1024, 0, c->stream>>>(f, n, np, out, info); LAUNCHED();
        return SVO_OK;
    }
    const ull nt = (n + LB_TILE - 1) / LB_TILE;
    int rc = lookback_prepare(c, nt);
    if (rc) return rc;
    OneValue<F> g{ f };
    k_scan_lookback<1><<<(unsigned)nt, LB_THREADS, 0, c->stream>>>(g, n, np, out, (ull*)nullptr, (ull*)nullptr, c->lb_state.as<ull>(), c->lb_ticket.as<ull>(), c->lb_tickets, c->lb_epoch, info); LAUNCHED();
    c->lb_tickets += nt;
    return SVO_OK;
}

// child prefix (fc) and subtree-size prefix (ps) of the brick level in one single-pass scan
int exscan_level0(svo_ctx* c, const ull* mask, ull n, ull* fc, ull* ps) {
    if (n == 0) {
        CK(cudaMemsetAsync(fc, 0, sizeof(ull), c->stream));
        CK(cudaMemsetAsync(ps, 0, sizeof(ull), c->stream));
        return SVO_OK;
    }
    const ull nt = (n + LB_TILE - 1) / LB_TILE;
    int rc = lookback_prepare(c, nt);
    if (rc) return rc;
    BrickPrefixes g{ mask };
    k_scan_lookback<2><<<(unsigned)nt, LB_THREADS, 0, c->stream>>>(g, n, (const ull*)nullptr, fc, ps, (ull*)nullptr, c->lb_state.as<ull>(), c->lb_ticket.as<ull>(), c->lb_tickets, c->lb_epoch, (BuildInfo*)nullptr); LAUNCHED();
    c->lb_tickets += nt;
    return SVO_OK;
}

int ilog2u(uint64_t v) { int r = -1; while (v) { v >>= 1; r++; } return r; }

// Allocates (and zeroes) the dense pyramid for the current geometry (gridsize, shard).
int ensure_pyramid(svo_ctx* c) {
    const uint64_t g = c->prm.gridsize;
    // everything nwords[] depends on: gridsize, shard layout (world, rank) and the chunk depth / top local level (which
    // change with -l on a sharded context)
    const uint64_t geom_key = (g << 32) | ((uint64_t)c->dc << 24) | ((uint64_t)c->J << 16) | ((uint64_t)c->world << 8) | (uint64_t)c->rank;
    if (c->dense_grid != geom_key) {
        size_t total = 0;
        for (int j = 0; j < c->nl; j++) total += (size_t)c->nwords[j] * 8;
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t have = 0;
        for (int j = 0; j < MAX_LEVELS; j++) have += c->dense[j].cap;
        if (total > free_b + have) {
            char m[256];
            snprintf(m, sizeof m, "bit-grid pyramid for gridsize %llu needs %.1f GiB, device has %.1f GiB free",
                     (ull)g, total / 1073741824.0, (free_b + have) / 1073741824.0);
            return fail(c, SVO_E_NOMEM, m);
        }
        for (int j = 0; j < MAX_LEVELS; j++) c->dense[j].release();
        for (int j = 0; j < c->nl; j++) CK(c->dense[j].ensure((size_t)c->nwords[j] * 8));
        c->dense_grid = geom_key;
        c->dense_clean = false;
        ull* ptrs[MAX_LEVELS] = { nullptr };
        for (int j = 0; j < c->nl; j++) ptrs[j] = c->dense[j].as<ull>();
        CK(c->d_lvlptrs.ensure(sizeof ptrs));
        CK(c->d_nwords.ensure(sizeof c->nwords));
        CK(c->d_counts.ensure(MAX_LEVELS * sizeof(ull)));
        CK(cudaMemcpyAsync(c->d_lvlptrs.p, ptrs, sizeof ptrs, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_nwords.p, c->nwords, sizeof c->nwords, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));      // ptrs / nwords are stack / member memory
    }
    if (!c->dense_clean) {
        for (int j = 0; j < c->nl; j++) CK(cudaMemsetAsync(c->dense[j].p, 0, (size_t)c->nwords[j] * 8, c->stream));
        c->dense_clean = true;
    }
    return SVO_OK;
}

VoxJob make_voxjob(svo_ctx* c) {
    VoxJob J;
    memset(&J, 0, sizeof J);
    J.tris = c->d_tris;
    J.fpt = (uint32_t)c->fpt;
    J.q_begin = c->q_begin;
    J.q_end = c->q_end;
    const bool lists = c->P > 1 && c->use_lists;
    J.pair_tri = lists ? c->pair_tri.as<uint32_t>() : nullptr;
    J.part_off = lists ? c->part_off.as<uint64_t>() : nullptr;
    for (int i = 0; i < 32; i++) { J.bmin[i] = c->slab_min[i]; J.bmax[i] = c->slab_max[i]; }
    J.p_first = (uint32_t)c->p_first; J.p_last = (uint32_t)c->p_last;
    J.inv_slab = c->inv_slab;
    J.qcap = c->qcap;
    J.P = (uint32_t)c->P;
    J.k = (uint32_t)c->k;
    J.side = c->side;
    J.g = (uint32_t)c->prm.gridsize;
    J.u = c->unit_vox;
    J.unit_div = c->unit_div;
    J.six = c->prm.separability == 6 ? 1 : 0;
    J.nl = c->J + 1;                                             // the cascade stops at the top local level
    for (int j = 0; j <= c->J; j++) J.lvl[j] = c->dense[j].as<ull>() - c->bias[j];   // indexed with global word indices
    J.w_lo = c->bias[0];
    J.w_hi = c->bias[0] + c->nwords[0];
    for (int a = 0; a < 3; a++) { J.sb_lo[a] = c->sb_lo[a]; J.sb_hi[a] = c->sb_hi[a]; }
    J.queue[0] = c->queue[0].as<ull>();
    J.queue[1] = c->queue[1].as<ull>();
    J.qcount = c->qcount.as<ull>();
    J.small_max = 512;
    J.medium_max = 32768;
    J.small_windows = 8;
    if (const char* e = getenv("SVO_SMALL_WINDOWS")) J.small_windows = (unsigned)strtoul(e, nullptr, 10);
    J.tilemask = c->lv[0].mask.as<ull>();
    if (const char* e = getenv("SVO_SMALL_MAX")) J.small_max = strtoull(e, nullptr, 10);
    if (const char* e = getenv("SVO_MEDIUM_MAX")) J.medium_max = strtoull(e, nullptr, 10);
    J.tileidx = c->tileidx.p ? c->tileidx.as<uint32_t>() - c->bias[0] : nullptr;
    J.leafprefix = c->lv[0].fc.as<ull>();
    J.owner = c->owner.as<uint32_t>();
    if (c->sliced) {
        J.segs.n = c->world;
        for (int r = 0; r < c->world; r++) J.segs.ptr[r] = c->peer_slice[r];
        const SliceCtrl* own = (const SliceCtrl*)c->sl_ctrl.p;
        J.segs.nslice = own->nslice;
        J.subset = c->sl_list.as<uint32_t>();
        J.pull_cap = c->sl_cap_blocks * 4;
        J.pull_counts = &own->count[0][c->rank];
        J.pull_first = c->rank;

    }
    return J;
}

template <bool OWNER>
int launch_voxelizer(svo_ctx* c) {
    if (c->q_end == c->q_begin) return SVO_OK;
    VoxJob J = make_voxjob(c);
    const size_t smem = J.pair_tri ? 0 : (size_t)VOX_BLOCK * c->fpt * sizeof(float);
    if (OWNER)      // the owner pass has its own unit tickets; a build may be repeated on the same voxelization
        CK(cudaMemsetAsync(c->qcount.as<ull>() + VOX_TICKET_BASE + VOX_TICKETS * VOX_TICKET_STRIDE, 0, (size_t)VOX_TICKETS * VOX_TICKET_STRIDE * sizeof(ull), c->stream));
    if (!OWNER) mark(c, EV_VS0);
    if (c->sliced) {
        // remote staging: wait (on the device) until every peer has published its block lists, then walk them
        if (!OWNER) {
            mark(c, EV_PW0);
            k_slice_wait<<<1, MAX_WORLD, 0, c->stream>>>((SliceCtrl*)c->sl_ctrl.p, c->world, 0, c->sl_epoch); LAUNCHED();
            mark(c, EV_PW1);
            mark(c, EV_VS0);                 // ms_vox_small excludes the wait for the peers' lists
        }
        const unsigned g2 = (unsigned)c->sm_count * SVO_VOX_MINBLOCKS;
        const size_t smem2 = 2 * (size_t)VOX_BLOCK * c->fpt * sizeof(float);      // per warp: two 32-triangle buffers
        const bool big = c->prm.gridsize > 4096;
        if (J.P > 1) {
            if (big) { k_vox_warp<OWNER, true, 2, true><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
            else { k_vox_warp<OWNER, true, 2, false><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
        } else {
            if (big) { k_vox_warp<OWNER, false, 2, true><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
            else { k_vox_warp<OWNER, false, 2, false><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
        }
    } else if (c->use_subset) {
        // sharded: compact the triangles that touch this rank's slab once (the owner pass reuses the list)
        if (!OWNER) {
            FilterJob Fj;
            memset(&Fj, 0, sizeof Fj);
            Fj.tris = c->d_tris; Fj.fpt = (uint32_t)c->fpt; Fj.n_tris = c->n_tris;
            Fj.use_partitions = c->P > 1 ? 1 : 0; Fj.k = c->k;
            for (int i = 0; i < 32; i++) { Fj.bmin[i] = c->slab_min[i]; Fj.bmax[i] = c->slab_max[i]; }
            for (int a = 0; a < 3; a++) {
                Fj.lo[a] = c->P > 1 ? c->sb_lo[a] / (int)c->side : c->sb_lo[a];
                Fj.hi[a] = c->P > 1 ? c->sb_hi[a] / (int)c->side : c->sb_hi[a];
            }
            Fj.unit_div = c->unit_div; Fj.gmax = (int)c->prm.gridsize - 1;
            Fj.out = c->subset.as<uint32_t>(); Fj.count = c->qcount.as<ull>() + 4;
            k_owner_filter<<<blocks_for(c->n_tris, VOX_BLOCK), VOX_BLOCK, 0, c->stream>>>(Fj); LAUNCHED();
        }
        J.subset = c->subset.as<uint32_t>(); J.subset_count = c->qcount.as<ull>() + 4;
        const unsigned g2 = (unsigned)c->sm_count * SVO_VOX_MINBLOCKS;
        const size_t smem2 = (size_t)VOX_BLOCK * c->fpt * sizeof(float);
        if (J.P > 1) { k_vox_small<OWNER, true, 1><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
        else { k_vox_small<OWNER, false, 1><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
    } else if (J.pair_tri == nullptr && c->warp_kernel) {
        // one GPU (or a compact private copy): warp-persistent kernel over all 32-triangle units
        const ull units = (c->q_end - c->q_begin + UNIT - 1) / UNIT;
        const unsigned g2 = (unsigned)std::min<ull>((units + VOX_BLOCK / 32 - 1) / (VOX_BLOCK / 32), (ull)c->sm_count * SVO_VOX_MINBLOCKS);
        const size_t smem2 = 2 * (size_t)VOX_BLOCK * c->fpt * sizeof(float);
        const bool big = c->prm.gridsize > 4096;
        if (J.P > 1) {
            if (big) { k_vox_warp<OWNER, true, 0, true><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
            else { k_vox_warp<OWNER, true, 0, false><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
        } else {
            if (big) { k_vox_warp<OWNER, false, 0, true><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
            else { k_vox_warp<OWNER, false, 0, false><<<g2, VOX_BLOCK, smem2, c->stream>>>(J); LAUNCHED(); }
        }
    } else if (J.pair_tri == nullptr && J.P > 1) { k_vox_small<OWNER, true, 0><<<blocks_for(c->q_end - c->q_begin, VOX_BLOCK), VOX_BLOCK, smem, c->stream>>>(J); LAUNCHED(); }
    else { k_vox_small<OWNER, false, 0><<<blocks_for(c->q_end - c->q_begin, VOX_BLOCK), VOX_BLOCK, smem, c->stream>>>(J); LAUNCHED(); }
    if (!OWNER) mark(c, EV_VS1);
    const unsigned grid = (unsigned)c->sm_count * 4;
    k_vox_queued<OWNER><<<grid, WARPS_PER_BLOCK * 32, 0, c->stream>>>(J); LAUNCHED();
    return SVO_OK;
}

int validate_params(svo_ctx* c, const svo_params* p) {
    if (!p) return fail(c, SVO_E_INVALID, "params is NULL");
    const uint64_t g = p->gridsize;
    if (g < 2 || (g & (g - 1)) != 0) return fail(c, SVO_E_INVALID, "gridsize must be a power of two >= 2");
    if (g > (1ull << 20)) return fail(c, SVO_E_INVALID, "gridsize above 2^20 is not supported");
    if (p->memory_limit_mb < 1) return fail(c, SVO_E_INVALID, "memory_limit_mb must be >= 1");
    if (!(p->bbox_max0 > p->bbox_min0)) return fail(c, SVO_E_INVALID, "bbox_max0 must exceed bbox_min0");
    if (p->color_mode < 0 || p->color_mode > 3) return fail(c, SVO_E_INVALID, "unknown color_mode");
    if (p->separability != 0 && p->separability != 6 && p->separability != 26) return fail(c, SVO_E_INVALID, "separability must be 0 / 26 (conservative, the reference's test) or 6 (thin)");
    return SVO_OK;
}

}  // namespace

extern "C" {

const char* svo_version(void) { return "ooc_svo_builder_b200 0.1 (reference ooc_svo_builder 1.6.4 byte layout)"; }

const char* svo_last_error(const svo_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int svo_ctx_create(int device, svo_ctx** out) {
    svo_ctx* c = nullptr;
    if (!out) return fail(c, SVO_E_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(c, SVO_E_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                       " (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(c, SVO_E_INVALID, "device index out of range");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char m[512];
        snprintf(m, sizeof m, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", device, prop.name, prop.major, prop.minor);
        return fail(c, SVO_E_CUDA, m);
    }
    CK(cudaSetDevice(device));
    svo_ctx* n = new svo_ctx();
    n->device = device;
    n->sm_count = prop.multiProcessorCount;
    {
        int r = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_emit_leaf<false, 4, true>, WARPS_PER_BLOCK * 32, 0) == cudaSuccess && r > 0) n->res_emit_leaf[0] = r;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_emit_leaf<true, 4, true>, WARPS_PER_BLOCK * 32, 0) == cudaSuccess && r > 0) n->res_emit_leaf[1] = r;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_emit_upper_fast<true>, WARPS_PER_BLOCK * 32, 0) == cudaSuccess && r > 0) n->res_emit_upper = r;
    }
    memset(&n->stats, 0, sizeof n->stats);
    memset(&n->prm, 0, sizeof n->prm);
    memset(n->nwords, 0, sizeof n->nwords);
    memset(n->bias, 0, sizeof n->bias);
    memset(n->fcap, 0, sizeof n->fcap);
    for (int i = 0; i < EV_COUNT; i++) n->ev_set[i] = false;
    c = n;
    cudaError_t e2 = cudaStreamCreateWithFlags(&n->own_stream, cudaStreamNonBlocking);
    n->stream = n->own_stream;
    for (int i = 0; i < EV_COUNT && e2 == cudaSuccess; i++) e2 = cudaEventCreate(&n->ev[i]);
    if (e2 == cudaSuccess) e2 = cudaHostAlloc((void**)&n->h_pinned, 64 * sizeof(ull), cudaHostAllocDefault);
    if (e2 == cudaSuccess) e2 = cudaHostAlloc((void**)&n->h_info, sizeof(BuildInfo), cudaHostAllocDefault);
    if (e2 != cudaSuccess) {
        std::string m = std::string("context setup: ") + cudaGetErrorString(e2);
        delete n;
        return fail(nullptr, SVO_E_CUDA, m);
    }
    *out = n;
    return SVO_OK;
}

void svo_ctx_destroy(svo_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->tri_own.release();
    for (int j = 0; j < MAX_LEVELS; j++) {
        c->dense[j].release();
        c->lv[j].key.release(); c->lv[j].mask.release(); c->lv[j].fc.release(); c->lv[j].ps.release(); c->lv[j].base.release();
        c->lv[j].pi.release(); c->lv[j].pl.release(); c->lv[j].ibase.release(); c->lv[j].cache.release();
    }
    c->glv.key.release(); c->glv.mask.release(); c->glv.fc.release(); c->glv.ps.release(); c->glv.base.release();
    c->glv.pi.release(); c->glv.pl.release(); c->glv.ibase.release(); c->glv.cache.release();
    c->table_own.release();
    for (int q = 0; q < 4; q++) c->dcol[q].release();
    c->d_lvlptrs.release(); c->d_nwords.release(); c->d_counts.release();
    c->part_counts.release(); c->part_cursor.release(); c->part_off.release(); c->pair_tri.release();
    c->queue[0].release(); c->queue[1].release(); c->qcount.release(); c->subset.release();
    c->window.release(); c->sl_cursor.release(); c->sl_ubox.release();
    c->inbox.release(); c->ctrl_buf.release(); c->blockcnt.release(); c->blockoff.release();
    if (c->h_ctrl) cudaFreeHost(c->h_ctrl);
    c->lb_state.release(); c->lb_ticket.release();
    c->nodes.release(); c->data.release(); c->owner.release(); c->tileidx.release(); c->codes.release();
    for (int i = 0; i < EV_COUNT; i++) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; i++) if (c->up_ev[i]) cudaEventDestroy(c->up_ev[i]);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_info) cudaFreeHost(c->h_info);
    c->info_buf.release(); c->merge_scratch.release(); c->merge_rpos.release(); c->merge_rrec.release(); c->brick_lp.release(); c->brick_sp.release();
    c->d_rpos.release(); c->d_rrec.release(); c->d_dpos.release(); c->d_drec.release(); c->stage_data.release();
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int svo_ctx_set_stream(svo_ctx* c, void* cuda_stream) {
    if (!c) return SVO_E_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return SVO_OK;
}

uint64_t svo_estimate_partitions(uint64_t gridsize, uint64_t memory_limit) {
    // partitioner.cpp:12-28
    uint64_t required = (gridsize * gridsize * gridsize) / 1024 / 1024;
    if (required <= memory_limit) return 1;
    uint64_t numpartitions = 1, required_partition = required;
    while (required_partition > memory_limit) {
        required_partition /= 8;
        numpartitions *= 8;
    }
    return numpartitions;
}

float svo_text_roundtrip_float(float v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)v);
    return strtof(buf, nullptr);
}

void* svo_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
void svo_host_free(void* p) { if (p) cudaFreeHost(p); }

int svo_synchronize(svo_ctx* c) {
    if (!c) return SVO_E_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

static int set_tris_common(svo_ctx* c, uint64_t n_tris, int fpt) {
    if (fpt != 9 && fpt != 21) return fail(c, SVO_E_INVALID, "floats_per_tri must be 9 (binary) or 21 (payload)");
    if (n_tris > 0xffffffffULL) return fail(c, SVO_E_INVALID, "more than 2^32-1 triangles");
    c->n_tris = n_tris;
    c->fpt = fpt;
    c->have_tris = true;
    c->dispatched = false;
    c->sliced = false;
    c->partitioned = c->voxelized = c->built = false;
    return SVO_OK;
}

int svo_set_triangles(svo_ctx* c, const float* tris, uint64_t n_tris, int fpt) {
    if (!c) return SVO_E_INVALID;
    if (n_tris && !tris) return fail(c, SVO_E_INVALID, "tris is NULL");
    CK(cudaSetDevice(c->device));
    int rc = set_tris_common(c, n_tris, fpt);
    if (rc) return rc;
    const size_t bytes = (size_t)n_tris * fpt * sizeof(float);
    CK(c->tri_own.ensure(bytes ? bytes : 16));
    for (int i = 0; i < EV_COUNT; i++) c->ev_set[i] = false;
    mark(c, EV_UP0);
    if (bytes) CK(cudaMemcpyAsync(c->tri_own.p, tris, bytes, cudaMemcpyHostToDevice, c->stream));
    mark(c, EV_UP1);
    c->d_tris = c->tri_own.as<float>();
    return SVO_OK;
}

int svo_triangles_begin(svo_ctx* c, uint64_t n_tris, int fpt) {
    if (!c) return SVO_E_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = set_tris_common(c, n_tris, fpt);
    if (rc) return rc;
    c->have_tris = false;                              // complete only after the last append
    const size_t bytes = (size_t)n_tris * fpt * sizeof(float);
    CK(c->tri_own.ensure(bytes ? bytes : 16));
    c->d_tris = c->tri_own.as<float>();
    c->stream_fill = 0;
    for (int i = 0; i < EV_COUNT; i++) c->ev_set[i] = false;
    mark(c, EV_UP0);
    if (!c->up_ev[0]) { CK(cudaEventCreateWithFlags(&c->up_ev[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->up_ev[1], cudaEventDisableTiming)); }
    c->up_slot = 0; c->up_pending[0] = c->up_pending[1] = false;
    if (n_tris == 0) { c->have_tris = true; mark(c, EV_UP1); }
    return SVO_OK;
}

int svo_triangles_append(svo_ctx* c, const float* host_chunk, uint64_t n) {
    if (!c) return SVO_E_INVALID;
    if (c->have_tris || !c->d_tris || c->d_tris != c->tri_own.as<float>()) return fail(c, SVO_E_INVALID, "svo_triangles_append without svo_triangles_begin");
    if (c->stream_fill + n > c->n_tris) return fail(c, SVO_E_RANGE, "more triangles appended than announced");
    if (n && !host_chunk) return fail(c, SVO_E_INVALID, "host_chunk is NULL");
    CK(cudaSetDevice(c->device));
    const size_t rec = (size_t)c->fpt * sizeof(float);
    if (n) CK(cudaMemcpyAsync((char*)c->tri_own.p + c->stream_fill * rec, host_chunk, n * rec, cudaMemcpyHostToDevice, c->stream));
    // double-buffer contract: when this call returns, every EARLIER chunk has been copied (its buffer may be refilled);
    // the chunk just passed is still in flight
    const int slot = c->up_slot;
    CK(cudaEventRecord(c->up_ev[slot], c->stream));
    c->up_pending[slot] = true;
    if (c->up_pending[slot ^ 1]) { CK(cudaEventSynchronize(c->up_ev[slot ^ 1])); c->up_pending[slot ^ 1] = false; }
    c->up_slot = slot ^ 1;
    c->stream_fill += n;
    if (c->stream_fill == c->n_tris) { c->have_tris = true; mark(c, EV_UP1); }
    return SVO_OK;
}

int svo_set_triangles_device(svo_ctx* c, const float* tris, uint64_t n_tris, int fpt) {
    if (!c) return SVO_E_INVALID;
    if (n_tris && !tris) return fail(c, SVO_E_INVALID, "tris is NULL");
    if (((uintptr_t)tris & 15) != 0) return fail(c, SVO_E_INVALID, "device triangle pointer must be 16-byte aligned");
    CK(cudaSetDevice(c->device));
    int rc = set_tris_common(c, n_tris, fpt);
    if (rc) return rc;
    for (int i = 0; i < EV_COUNT; i++) c->ev_set[i] = false;
    c->d_tris = tris;
    return SVO_OK;
}

static int setup_geometry(svo_ctx* c);

// Grid constants of a job (host arithmetic only): depth, logical partitions, unit lengths, partition slabs.
static int derive_grid(svo_ctx* c, const svo_params* params) {
    c->prm = *params;
    const uint64_t g = params->gridsize;
    c->D = ilog2u(g);
    c->nl = (c->D + 1) / 2;
    c->P = svo_estimate_partitions(g, params->memory_limit_mb);               // main.cpp:298
    c->k = ilog2u(c->P) / 3;
    if (c->k > 5) return fail(c, SVO_E_INVALID, "more than 8^5 logical partitions are not supported");
    c->side = (uint32_t)(g >> c->k);
    // main.cpp:304-311: the voxelizer's unit length comes from the bbox re-read from the .trip text header
    const float rmin = svo_text_roundtrip_float(params->bbox_min0), rmax = svo_text_roundtrip_float(params->bbox_max0);
    c->unit_vox = (rmax - rmin) / (float)g;
    c->unit_div = 1.0f / c->unit_vox;                                          // voxelizer.cpp:164
    if (c->P > 1) {
        const float unit_part = (params->bbox_max0 - params->bbox_min0) / (float)g;   // partitioner.cpp:45
        c->inv_slab = 1.0f / ((float)c->side * unit_part);
        for (uint32_t i = 0; i < (1u << c->k); i++) {
            c->slab_min[i] = (float)(uint32_t)(i * c->side) * unit_part;                    // :54-56
            c->slab_max[i] = (float)(uint32_t)((i + 1) * c->side - 1 + 1u) * unit_part;     // :57-59
        }
    }
    return SVO_OK;
}

int svo_partition(svo_ctx* c, const svo_params* params, uint64_t* n_partitions, uint64_t* part_tricounts, uint64_t cap) {
    if (!c) return SVO_E_INVALID;
    int rc = validate_params(c, params);
    if (rc) return rc;
    if (!c->have_tris) return fail(c, SVO_E_INVALID, "svo_partition before svo_set_triangles");
    if ((params->payload ? 21 : 9) != c->fpt) return fail(c, SVO_E_INVALID, "params.payload does not match floats_per_tri");
    CK(cudaSetDevice(c->device));
    c->voxelized = c->built = false;
    c->launches = 0;
    c->tl.on = getenv("SVO_TIMELINE") != nullptr; c->tl.n = 0; c->tl.stamp("partition");
    rc = derive_grid(c, params);
    if (rc) return rc;
    c->h_part_counts.assign(c->P, 0);
    mark(c, EV_PART0);
    if (c->P == 1) {
        // partition_one, partitioner.cpp:80-98: the single partition holds every triangle, untested
        c->n_pairs = c->n_tris;
        c->h_part_counts[0] = c->n_tris;
    } else {
        BinJob B;
        memset(&B, 0, sizeof B);
        B.tris = c->d_tris; B.fpt = (uint32_t)c->fpt; B.n_tris = c->n_tris; B.k = (uint32_t)c->k; B.P = (uint32_t)c->P;
        B.inv_slab = c->inv_slab;
        for (uint32_t i = 0; i < (1u << c->k); i++) { B.bmin[i] = c->slab_min[i]; B.bmax[i] = c->slab_max[i]; }
        const char* lists_env = getenv("SVO_PARTITION_LISTS");
        c->use_lists = lists_env && lists_env[0] == '1';
        c->n_pairs = 0;
        // Default: the voxelizer enumerates each triangle's partitions inline (no lists, no read-back). The
        // per-partition counts (what the reference writes to the .trip header) are computed only on request.
        if (c->sliced && (c->use_lists || part_tricounts))
            return fail(c, SVO_E_INVALID, "per-partition lists / counts are not available with remote triangle slices");
        if (c->use_lists || part_tricounts) {
            CK(c->part_counts.ensure(c->P * sizeof(ull)));
            CK(c->part_cursor.ensure(c->P * sizeof(ull)));
            CK(c->part_off.ensure((c->P + 1) * sizeof(ull)));
            CK(cudaMemsetAsync(c->part_counts.p, 0, c->P * sizeof(ull), c->stream));
            CK(cudaMemsetAsync(c->part_cursor.p, 0, c->P * sizeof(ull), c->stream));
            B.counts = c->part_counts.as<ull>(); B.cursor = c->part_cursor.as<ull>(); B.off = c->part_off.as<ull>();
            if (c->n_tris) {
                const size_t smem = c->P <= 4096 ? c->P * sizeof(unsigned) : 0;
                k_bin<false><<<blocks_for(c->n_tris, 256), 256, smem, c->stream>>>(B); LAUNCHED();
            }
            CountOp op{ c->part_counts.as<ull>() };
            rc = exscan(c, op, c->P, c->part_off.as<ull>());
            if (rc) return rc;
            std::vector<ull> off(c->P + 1);
            CK(cudaMemcpyAsync(off.data(), c->part_off.p, (c->P + 1) * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            for (uint64_t i = 0; i < c->P; i++) c->h_part_counts[i] = off[i + 1] - off[i];
            c->n_pairs = off[c->P];
            if (c->use_lists) {
                CK(c->pair_tri.ensure((c->n_pairs ? c->n_pairs : 1) * sizeof(uint32_t)));
                B.pair_tri = c->pair_tri.as<uint32_t>();
                if (c->n_tris) { k_bin<true><<<blocks_for(c->n_tris, 256), 256, 0, c->stream>>>(B); LAUNCHED(); }
            }
        }
    }
    mark(c, EV_PART1);
    rc = setup_geometry(c);
    if (rc) return rc;
    c->partitioned = true;
    if (n_partitions) *n_partitions = c->P;
    if (part_tricounts) {
        if (cap < c->P) return fail(c, SVO_E_RANGE, "part_tricounts capacity is smaller than the partition count");
        for (uint64_t i = 0; i < c->P; i++) part_tricounts[i] = c->h_part_counts[i];
    }
    return SVO_OK;
}

int svo_shard_configure(svo_ctx* c, int rank, int world) {
    if (!c) return SVO_E_INVALID;
    if (world < 1 || rank < 0 || rank >= world || (world & (world - 1)) != 0)
        return fail(c, SVO_E_INVALID, "shard world size must be a power of two and 0 <= rank < world");
    if (world != c->world || rank != c->rank) { c->sl_attached = false; c->attached = false; c->sliced = false; c->dispatched = false; }
    c->world = world;
    c->rank = rank;
    c->partitioned = c->voxelized = c->built = false;
    return SVO_OK;
}

// Shard geometry: the grid is cut into 8^dc chunks (subtrees at depth dc >= k), rank r owns a contiguous
// Morton range of them. Pyramid levels 0..J live inside a chunk (local, dense per slab); levels above J are
// tiny and replicated on every rank.
static int shard_chunk_depth(const svo_ctx* c) {
    if (c->world <= 1) return 0;
    int need = 0;
    while ((1 << (3 * need)) < c->world) need++;
    return c->k > need ? c->k : need;
}
// voxel bounding box of the slab of `rank` (exact when the slab is a box, a superset otherwise)
static void shard_box(const svo_ctx* c, int rank, int dc, int lo[3], int hi[3]) {
    if (c->world == 1) {
        for (int a = 0; a < 3; a++) { lo[a] = 0; hi[a] = (int)c->prm.gridsize - 1; }
        return;
    }
    const ull nchunks = 1ULL << (3 * dc);
    const ull c0 = nchunks * (ull)rank / (ull)c->world, c1 = nchunks * (ull)(rank + 1) / (ull)c->world;
    const uint32_t cs = (uint32_t)(c->prm.gridsize >> dc);
    for (int a = 0; a < 3; a++) { lo[a] = 0x7fffffff; hi[a] = -1; }
    for (ull ch = c0; ch < c1; ch++) {
        const uint32_t cc[3] = { compact3(ch), compact3(ch >> 1), compact3(ch >> 2) };
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], (int)(cc[a] * cs));
            hi[a] = std::max(hi[a], (int)(cc[a] * cs + cs - 1));
        }
    }
}

static int setup_geometry(svo_ctx* c) {
    const int D = c->D;
    const int dc = shard_chunk_depth(c);
    if (c->world > 1) {
        if (D - dc < 2) return fail(c, SVO_E_INVALID, "gridsize too small for this many shards");
    }
    c->tstride = (c->world > 1 && c->prm.generate_levels) ? 8 : 4;
    c->dc = dc;
    c->J = (D - dc) / 2 - 1;
    if (c->J < 0) c->J = 0;                                     // gridsize 2: the single level-0 word is the (virtual) top
    const ull nchunks = 1ULL << (3 * dc);
    c->c0 = nchunks * (ull)c->rank / (ull)c->world;
    c->c1 = nchunks * (ull)(c->rank + 1) / (ull)c->world;
    const int chunk_bits = 3 * (D - dc);
    c->slab_start = c->c0 << chunk_bits;
    c->slab_end = c->c1 << chunk_bits;
    for (int j = 0; j < MAX_LEVELS; j++) {
        const int sh = 6 * (j + 1);
        if (j >= c->nl) { c->nwords[j] = 0; c->bias[j] = 0; continue; }
        if (j <= c->J) {                                        // local level: the slab's words, global index = local + bias
            const ull lo = c->slab_start >> sh, hi = c->slab_end >> sh;
            c->bias[j] = lo;
            c->nwords[j] = hi > lo ? hi - lo : 1;
        } else {                                                // replicated upper level
            c->bias[j] = 0;
            c->nwords[j] = 3 * D >= sh ? (1ULL << (3 * D - sh)) : 1ULL;
        }
    }
    const int shJ = 6 * (c->J + 1);
    c->WJ = 3 * D >= shJ ? (1ULL << (3 * D - shJ)) : 1ULL;
    // voxel bounding box of the slab (exact when the slab is a box, a superset otherwise; the sinks filter by word range)
    shard_box(c, c->rank, dc, c->sb_lo, c->sb_hi);
    // logical partitions that intersect the slab, and the work items this context walks
    c->p_first = 0; c->p_last = c->P - 1;
    if (c->world > 1 && c->P > 1) {
        const int sh = 3 * (dc - c->k);
        c->p_first = c->c0 >> sh; c->p_last = (c->c1 - 1) >> sh;
    }
    const bool lists = c->P > 1 && c->use_lists;
    if (!lists) { c->q_begin = 0; c->q_end = c->n_tris; }        // one thread per triangle, partitions enumerated inline
    else {
        ull acc = 0;
        c->q_begin = c->q_end = 0;
        for (ull p = 0; p < c->P; p++) {
            if (p == c->p_first) c->q_begin = acc;
            acc += c->h_part_counts[p];
            if (p == c->p_last) c->q_end = acc;
        }
    }
    return SVO_OK;
}

int svo_voxelize(svo_ctx* c) {
    if (!c) return SVO_E_INVALID;
    if (!c->partitioned) return fail(c, SVO_E_INVALID, "svo_voxelize before svo_partition");
    CK(cudaSetDevice(c->device));
    int rc = ensure_pyramid(c);
    if (rc) return rc;
    CK(c->qcount.ensure(VOX_QCOUNT_WORDS * sizeof(ull)));
    CK(cudaMemsetAsync(c->qcount.p, 0, VOX_QCOUNT_WORDS * sizeof(ull), c->stream));
    c->use_subset = c->world > 1 && !c->dispatched && !c->sliced && !(c->P > 1 && c->use_lists);
    { const char* e = getenv("SVO_VOX_KERNEL"); c->warp_kernel = !(e && strcmp(e, "block") == 0); }
    if (c->use_subset) CK(c->subset.ensure((size_t)(c->n_tris / VOX_BLOCK + 2) * sizeof(uint32_t)));
    // queue capacity: exact with lists; with inline enumeration a triangle may appear once per partition it
    // touches, so leave headroom and detect overflow (qcount[3]) instead of trusting a bound
    const ull npairs = c->q_end - c->q_begin;
    c->qcap = (c->P > 1 && !c->use_lists) ? npairs + npairs / 2 + (1u << 20) : (npairs ? npairs : 1);
    const size_t qbytes = (size_t)c->qcap * sizeof(ull);
    CK(c->queue[0].ensure(qbytes));
    CK(c->queue[1].ensure(qbytes));
    mark(c, EV_VOX0);
    c->dense_clean = false;
    c->tl.stamp("vox_launch");
    rc = launch_voxelizer<false>(c);
    if (rc) return rc;
    mark(c, EV_VOX1);
    c->tl.stamp("vox_launched");
    c->voxelized = true;
    c->built = false;
    return SVO_OK;
}

// ---------------------------------------------------------------------------
// Build, phase A (local): compact tile lists of levels J..0 from the slab's pyramid, subtree sizes
// bottom-up, and this rank's entries of the exchange table.
// ---------------------------------------------------------------------------
static int alloc_level(svo_ctx* c, LevelBufs& L, ull n, bool pl, bool levels) {
    L.n = n;
    CK(L.key.ensure((n + 1) * sizeof(ull)));
    CK(L.mask.ensure((n + 1) * sizeof(ull)));
    CK(L.fc.ensure((n + 2) * sizeof(ull)));
    CK(L.ps.ensure((n + 2) * sizeof(ull)));
    CK(L.base.ensure((n + 1) * sizeof(ull)));
    if (pl) CK(L.pl.ensure((n + 2) * sizeof(ull)));
    if (levels) {
        CK(L.pi.ensure((n + 2) * sizeof(ull)));
        CK(L.ibase.ensure((n + 1) * sizeof(ull)));
        CK(L.cache.ensure((n + 1) * 6 * sizeof(float)));
    }
    return SVO_OK;
}

static int size_scans(svo_ctx* c, LevelBufs& L, const LevelBufs* child, bool pl, bool levels) {
    SizeOp op{ L.mask.as<ull>(), L.fc.as<ull>(), child ? child->ps.as<ull>() : nullptr };
    int rc = exscan(c, op, L.n, L.ps.as<ull>());
    if (rc) return rc;
    if (pl) {
        LeafCountOp lop{ L.mask.as<ull>(), L.fc.as<ull>(), child ? child->pl.as<ull>() : nullptr };
        if ((rc = exscan(c, lop, L.n, L.pl.as<ull>()))) return rc;
    }
    if (levels) {
        InternalOp iop{ L.mask.as<ull>(), L.fc.as<ull>(), child ? child->pi.as<ull>() : nullptr };
        if ((rc = exscan(c, iop, L.n, L.pi.as<ull>()))) return rc;
    }
    return SVO_OK;
}

static int build_phase_a(svo_ctx* c, ull* table) {
    const int J = c->J;
    const bool payload = c->prm.payload != 0, levels = c->prm.generate_levels != 0;
    const bool want_pl = levels;                        // leaf-count prefixes: only the -levels data indices need them
    c->want_pl = want_pl;
    mark(c, EV_BUILD0);
    // the voxelizer maintains dense levels 0 and 1 only: rebuild the levels above from level 1
    for (int j = 1; j < J; j++) {
        k_pyramid_up<<<blocks_for(c->nwords[j], 256), 256, 0, c->stream>>>(c->dense[j].as<ull>(), c->nwords[j], c->bias[j], c->dense[j + 1].as<ull>() - c->bias[j + 1]); LAUNCHED();
    }
    // ---- sync #1: how many non-zero words does every local level hold? ----
    CK(cudaMemsetAsync(c->d_counts.p, 0, MAX_LEVELS * sizeof(ull), c->stream));
    {
        dim3 grid((unsigned)c->sm_count * 4, (unsigned)(J + 1));
        k_level_counts<<<grid, 256, 0, c->stream>>>((ull* const*)c->d_lvlptrs.p, c->d_nwords.as<ull>(), J, c->d_counts.as<ull>()); LAUNCHED();
    }
    CK(cudaMemcpyAsync(c->h_pinned, c->d_counts.p, MAX_LEVELS * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
    c->tl.stamp("sync1_wait");
    CK(cudaStreamSynchronize(c->stream));
    c->tl.stamp("sync1_done");
    for (int j = 0; j <= J; j++) {
        int rc = alloc_level(c, c->lv[j], c->h_pinned[j], want_pl, levels);
        if (rc) return rc;
    }
    if (payload) CK(c->tileidx.ensure((size_t)c->nwords[0] * sizeof(uint32_t)));
    // levels jf..J with few tiles are walked by single-block fused kernels (one launch instead of ~9 per level)
    c->jf = 0;
    if (!levels && J >= 1) {
        int jf = J + 1;
        while (jf > 1 && c->lv[jf - 1].n <= 4096) jf--;
        if (jf <= J) c->jf = jf;
    }
    const int jf = c->jf;
    FusedJob F;
    if (jf) {
        memset(&F, 0, sizeof F);
        for (int j = 0; j <= J; j++) {
            F.lv[j] = c->lv[j].view(); F.dense[j] = c->dense[j].as<ull>() - c->bias[j];
            if (!want_pl) F.lv[j].pl = nullptr;
        }
        F.dense_top = c->dense[J].as<ull>(); F.top_words = c->nwords[J]; F.top_bias = c->bias[J];
        F.J = J; F.jf = jf;
    }
    // ---- top-down: compact tile lists ----
    if (jf) {
        if (c->lv[J].n) { k_fused_down<<<1, 1024, 0, c->stream>>>(F); LAUNCHED(); }
        else for (int j = jf; j <= J; j++) CK(cudaMemsetAsync(c->lv[j].fc.p, 0, sizeof(ull), c->stream));
    } else if (c->lv[J].n) {
        k_compact_top<<<1, 1024, 0, c->stream>>>(c->dense[J].as<ull>(), c->nwords[J], c->bias[J], c->lv[J].key.as<ull>(), c->lv[J].mask.as<ull>(),
                                                (payload && J == 0) ? c->tileidx.as<uint32_t>() : nullptr, J == 0 ? 1 : 0); LAUNCHED();
    }
    for (int j = J; j >= 0; j--) {
        if (j == 0 && !levels && (!jf || jf > 0)) {
            int rc = exscan_level0(c, c->lv[0].mask.as<ull>(), c->lv[0].n, c->lv[0].fc.as<ull>(), c->lv[0].ps.as<ull>());
            if (rc) return rc;
        } else if (!jf || j < jf) {
            PopcOp op{ c->lv[j].mask.as<ull>() };
            int rc = exscan(c, op, c->lv[j].n, c->lv[j].fc.as<ull>());
            if (rc) return rc;
        }
        if (j > 0 && c->lv[j].n && (!jf || j <= jf)) {
            k_expand<<<blocks_for(c->lv[j].n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(
                c->lv[j].view(), c->lv[j - 1].view(), c->dense[j - 1].as<ull>() - c->bias[j - 1],
                (payload && j == 1) ? c->tileidx.as<uint32_t>() - c->bias[0] : nullptr, j == 1 ? 1 : 0); LAUNCHED();
        }
    }
    // ---- bottom-up: subtree sizes (+ leaf / internal counts) ----
    for (int j = 0; j <= J; j++) {
        if (jf && j >= jf) break;
        if (j == 0 && !levels) continue;                 // done together with fc by exscan_level0
        int rc = size_scans(c, c->lv[j], j ? &c->lv[j - 1] : nullptr, want_pl, levels);
        if (rc) return rc;
    }
    if (jf) {
        if (c->lv[J].n) { k_fused_up<<<1, 1024, 0, c->stream>>>(F); LAUNCHED(); }
        else for (int j = jf; j <= J; j++) {
            CK(cudaMemsetAsync(c->lv[j].ps.p, 0, sizeof(ull), c->stream));
            if (want_pl) CK(cudaMemsetAsync(c->lv[j].pl.p, 0, sizeof(ull), c->stream));
        }
    }
    // ---- sharded -levels: the data caches (averaged colour + normal, Node::data_cache) of this rank's top tiles go into the
    // table, so that every rank can average the shared upper levels (OctreeBuilder.cpp:82-99). They only depend on leaf
    // VALUES, not on data indices: payload leaves are computed into a staging array in leaf-rank order (the owner pass is
    // kept for phase B), binary leaves cache zeros; k_levels_data then runs in cache-only mode.
    c->owner_ready = false;
    if (levels && c->world > 1) {
        CK(cudaMemcpyAsync(c->h_pinned + 36, c->lv[0].fc.as<ull>() + c->lv[0].n, sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const ull nleaf = c->lv[0].n ? c->h_pinned[36] : 0;
        EmitJob Ec;
        memset(&Ec, 0, sizeof Ec);
        Ec.leaf_data_mode = payload ? 1 : 0; Ec.levels = 1;
        float* stage = nullptr;
        if (payload && nleaf) {
            CK(c->stage_data.ensure((size_t)(nleaf + 1) * SVO_DATA_BYTES));
            CK(c->owner.ensure((size_t)nleaf * sizeof(uint32_t)));
            CK(cudaMemsetAsync(c->owner.p, 0xff, (size_t)nleaf * sizeof(uint32_t), c->stream));
            int rc = launch_voxelizer<true>(c);
            if (rc) return rc;
            PayloadJob Pj;
            memset(&Pj, 0, sizeof Pj);
            Pj.tris = c->d_tris; Pj.owner = c->owner.as<uint32_t>();
            if (c->sliced) {
                Pj.segs.n = c->world;
                for (int r = 0; r < c->world; r++) Pj.segs.ptr[r] = c->peer_slice[r];
                Pj.segs.nslice = ((const SliceCtrl*)c->sl_ctrl.p)->nslice;
            }
            Pj.data = c->stage_data.as<float>();
            Pj.unit_div = c->unit_div; Pj.gridsize_f = (float)c->prm.gridsize; Pj.color_mode = c->prm.color_mode;
            Pj.levels = 0; Pj.leaf_offset = 0;
            k_payload<<<blocks_for(c->lv[0].n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(c->lv[0].view(), Pj); LAUNCHED();
            c->owner_ready = true;
            stage = c->stage_data.as<float>();
        }
        for (int j = 0; j <= J; j++) {
            if (!c->lv[j].n) continue;
            const Level L = c->lv[j].view();
            k_levels_data<<<blocks_for(L.n, 128), 128, 0, c->stream>>>(L, j ? c->lv[j - 1].view() : L, j, Ec, stage, 0, 1); LAUNCHED();
        }
    }
    // ---- this rank's table entries ----
    if (table) {
        CK(cudaMemsetAsync(table, 0, (size_t)c->WJ * c->tstride * sizeof(ull), c->stream));
        if (c->lv[J].n) {
            TableFillJob Tf;
            memset(&Tf, 0, sizeof Tf);
            Tf.key = c->lv[J].key.as<ull>(); Tf.mask = c->lv[J].mask.as<ull>(); Tf.ps = c->lv[J].ps.as<ull>();
            Tf.pi = levels ? c->lv[J].pi.as<ull>() : nullptr;
            for (int j = 0; j <= J; j++) Tf.fc[j] = c->lv[j].fc.as<ull>();
            Tf.n = c->lv[J].n; Tf.J = J; Tf.table = table;      // np, info: NULL (memset)
            Tf.stride = c->tstride; Tf.cache = (c->tstride == 8) ? c->lv[J].cache.as<float>() : nullptr;
            k_table_fill<<<blocks_for(c->lv[J].n, 256), 256, 0, c->stream>>>(Tf); LAUNCHED();
        }
    }
    mark(c, EV_CMP1);
    c->tl.stamp("phaseA_launched");
    c->phase_a_done = true;
    return SVO_OK;
}

// ---------------------------------------------------------------------------
// Sharded merge of the shared upper levels, on the host: the summed table holds {mask, S, leaves} of every
// top-of-shard subtree (<= a few thousand entries). From it every rank derives, identically, the upper
// 64-tree levels, the global record counts, the file base of each subtree, its own file range and the
// (few) upper-level records that fall into that range. Same formulas as k_emit_upper.
// ---------------------------------------------------------------------------
static inline ull h_lowmask(int n) { return n >= 64 ? ~0ULL : ((1ULL << n) - 1ULL); }
// Pure host arithmetic (no CUDA): T = the summed table, geometry from derive_grid / setup_geometry. Fills the global counts,
// this rank's file range and leaf offsets, the bases of its own level-J tiles (h_ownbase) and the upper-level records that
// fall into its range (h_rpos / h_rrec). Also behind svo_shard_layout_from_table, which the CPU tests check against the oracle.
// -levels on the host: Node::data_cache arithmetic of OctreeBuilder.cpp:82-99 in the reference's float op order (the same as
// finish_average / k_levels_data on the device): colour = sum / n, normal = normalize(sum / n).
static inline void h_finish_average(const float* sum, float notnull, float* out) {
    out[0] = sum[0] / notnull; out[1] = sum[1] / notnull; out[2] = sum[2] / notnull;
    const float tx = sum[3] / notnull, ty = sum[4] / notnull, tz = sum[5] / notnull;
    const float d = (tx * tx + ty * ty) + tz * tz;
    const float inv = 1.0f / sqrtf(d);
    out[3] = tx * inv; out[4] = ty * inv; out[5] = tz * inv;
}
static void shard_merge_compute(svo_ctx* c, const ull* T) {
    const int J = c->J, top = c->nl - 1;
    const ull WJ = c->WJ;
    const bool d_even = (c->D % 2) == 0;
    const bool levels = c->prm.generate_levels != 0, payload = c->prm.payload != 0;
    const int ts = levels ? 8 : 4;                          // u64 per table entry
    std::vector<std::vector<ull>> M(top + 1), S(top + 1), B(top + 1);
    M[J].resize(WJ); S[J].resize(WJ); B[J].assign(WJ, 0);
    ull leaves_total = 0;
    for (ull e = 0; e < WJ; e++) { M[J][e] = T[ts * e]; S[J][e] = T[ts * e + 1]; leaves_total += T[ts * e + 2]; }
    for (int j = J + 1; j <= top; j++) {
        const size_t n = (size_t)c->nwords[j];
        M[j].assign(n, 0); S[j].assign(n, 0); B[j].assign(n, 0);
        for (size_t ch = 0; ch < M[j - 1].size(); ch++) if (M[j - 1][ch]) M[j][ch >> 6] |= 1ULL << (ch & 63);
        for (size_t w = 0; w < n; w++) {
            const ull W = M[j][w];
            if (!W) continue;
            ull sz = (ull)__builtin_popcountll(W) + (ull)__builtin_popcount(nonzero_bytes(W));
            for (int b = 0; b < 64; b++) if ((W >> b) & 1ULL) sz += S[j - 1][w * 64 + b];
            S[j][w] = sz;
        }
    }
    c->n_voxels = leaves_total;
    c->n_nodes = leaves_total == 0 ? 1 : S[top][0] + (d_even ? 1 : 0);
    // ---- -levels: internal-node counts, leaves through every subtree, data caches bottom-up; ranks and data indices top-down ----
    std::vector<std::vector<ull>> I(top + 1), LT(top + 1), IB(top + 1);      // internals in the subtree; leaves up to its END; post-order rank before it
    std::vector<std::vector<float>> CA(top + 1), CC(top + 1);                // tile cache (6 floats), byte-children caches (8 x 6 floats)
    if (levels) {
        I[J].assign(WJ, 0); LT[J].assign(WJ, 0); IB[J].assign(WJ, 0); CA[J].assign(WJ * 6, 0.f);
        ull run = 0;
        for (ull e = 0; e < WJ; e++) {
            run += T[ts * e + 2];
            LT[J][e] = run;
            I[J][e] = T[ts * e + 3];
            uint32_t w[6] = { (uint32_t)T[ts * e + 4], (uint32_t)(T[ts * e + 4] >> 32), (uint32_t)T[ts * e + 5], (uint32_t)(T[ts * e + 5] >> 32),
                              (uint32_t)T[ts * e + 6], (uint32_t)(T[ts * e + 6] >> 32) };
            memcpy(&CA[J][e * 6], w, sizeof w);
        }
        for (int j = J + 1; j <= top; j++) {
            const size_t n = M[j].size();
            I[j].assign(n, 0); LT[j].assign(n, 0); IB[j].assign(n, 0); CA[j].assign(n * 6, 0.f); CC[j].assign(n * 48, 0.f);
            for (size_t w = 0; w < n; w++) {
                const ull W = M[j][w];
                // leaves through the end of this word's range even when it is empty (prefix of the last child)
                LT[j][w] = LT[j - 1][std::min(w * 64 + 63, LT[j - 1].size() - 1)];
                if (!W) continue;
                ull inter = 1ULL + (ull)__builtin_popcount(nonzero_bytes(W));
                float wsum[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, wn = 0.0f;
                for (int k = 0; k < 8; k++) {
                    const uint32_t byte = (uint32_t)((W >> (8 * k)) & 0xffULL);
                    if (!byte) continue;
                    float csum[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f }, cn = 0.0f;
                    for (int b = 0; b < 8; b++) if ((byte >> b) & 1u) {
                        const size_t ch = w * 64 + 8 * k + b;
                        inter += I[j - 1][ch];
                        cn = cn + 1.0f;
                        for (int q = 0; q < 6; q++) csum[q] = csum[q] + CA[j - 1][ch * 6 + q];
                    }
                    float* cc = &CC[j][(w * 8 + k) * 6];
                    h_finish_average(csum, cn, cc);
                    wn = wn + 1.0f;
                    for (int q = 0; q < 6; q++) wsum[q] = wsum[q] + cc[q];
                }
                h_finish_average(wsum, wn, &CA[j][w * 6]);
                I[j][w] = inter;
            }
        }
    }
    const ull n_internal = (levels && leaves_total) ? I[top][0] - (d_even ? 0ULL : 1ULL) : 0ULL;     // the virtual top word of an odd depth is no node
    c->n_data = (payload ? 1 + leaves_total : 2) + n_internal;                                       // OctreeBuilder.cpp:25-29 (+ one record per internal node)
    auto data_index = [&](ull leaves_through, ull rank) -> ull { return (payload ? 1ULL + leaves_through : 2ULL) + rank; };
    // top-down: bases + upper records
    std::vector<ull> rpos, rrec, dpos;
    std::vector<float> drec;
    auto push = [&](ull pos, ull d0, ull d1, ull d2) { rpos.push_back(pos); rrec.push_back(d0); rrec.push_back(d1); rrec.push_back(d2); };
    auto push_data = [&](ull idx, const float* six) { dpos.push_back(idx); for (int q = 0; q < 6; q++) drec.push_back(six[q]); };
    for (int j = top; j > J; j--) {
        for (size_t w = 0; w < M[j].size(); w++) {
            const ull W = M[j][w];
            if (!W) continue;
            const ull base = B[j][w], sz = S[j][w];
            const uint32_t nzb = nonzero_bytes(W);
            const ull ib = levels ? IB[j][w] : 0ULL;
            ull acc = 0, iacc = 0;
            for (int k = 0; k < 8; k++) {
                const uint32_t byte = (uint32_t)((W >> (8 * k)) & 0xffULL);
                if (!byte) continue;
                const ull before = (ull)__builtin_popcountll(W & h_lowmask(8 * k));
                const ull krank = (ull)__builtin_popcount(nzb & ((1u << k) - 1u));
                std::vector<ull> gdata(8, 0ULL);
                size_t last_ch = 0;
                for (int b = 0; b < 8; b++) if ((byte >> b) & 1u) {
                    const size_t ch = w * 64 + 8 * k + b;
                    B[j - 1][ch] = base + acc + before;
                    acc += S[j - 1][ch];
                    if (levels) {
                        const ull gib = ib + iacc + krank;                       // internals completed before this grandchild's subtree
                        IB[j - 1][ch] = gib;
                        gdata[b] = data_index(LT[j - 1][ch], gib + I[j - 1][ch] - 1ULL);
                        iacc += I[j - 1][ch];
                        last_ch = ch;
                    }
                }
                const ull blk = base + acc + before;
                ull r = 0;
                for (int b = 0; b < 8; b++) if ((byte >> b) & 1u) {
                    const size_t ch = w * 64 + 8 * k + b;
                    const uint32_t gnz = nonzero_bytes(M[j - 1][ch]);
                    push(blk + r++, gdata[b], B[j - 1][ch] + S[j - 1][ch] - (ull)__builtin_popcount(gnz), child_offsets(gnz));
                }
                ull cdata = 0ULL;
                if (levels) {
                    cdata = data_index(LT[j - 1][last_ch], ib + iacc + krank);   // the byte's node completes right behind its last grandchild
                    push_data(cdata, &CC[j][(w * 8 + k) * 6]);
                }
                push(base + sz - (ull)__builtin_popcount(nzb) + krank, cdata, blk, child_offsets(byte));
            }
            const bool real_node = !(j == top && !d_even);
            ull own = 0ULL;
            if (levels && real_node) {
                own = data_index(LT[j][w], ib + I[j][w] - 1ULL);
                push_data(own, &CA[j][w * 6]);
            }
            if (j == top && d_even) push(sz, own, base + sz - (ull)__builtin_popcount(nzb), child_offsets(nzb));
        }
    }
    // this rank's tiles, offsets and file range
    const ull wj0 = c->bias[J], wj1 = c->bias[J] + c->nwords[J];
    c->leaf_offset = 0; c->n_voxels_local = 0;
    c->h_ownbase.clear(); c->h_ownibase.clear();
    for (ull e = 0; e < WJ; e++) {
        if (!M[J][e]) continue;
        if (e < wj0) c->leaf_offset += T[ts * e + 2];
        else if (e < wj1) { c->h_ownbase.push_back(B[J][e]); if (levels) c->h_ownibase.push_back(IB[J][e]); c->n_voxels_local += T[ts * e + 2]; }
    }
    auto first_base_from = [&](ull e0) -> ull { for (ull e = e0; e < WJ; e++) if (M[J][e]) return B[J][e]; return c->n_nodes; };
    c->node_lo = c->rank == 0 ? 0 : first_base_from(wj0);
    c->node_hi = c->rank == c->world - 1 ? c->n_nodes : first_base_from(wj1);
    if (leaves_total == 0) { c->node_lo = c->rank == 0 ? 0 : 1; c->node_hi = 1; }
    c->h_rpos.clear(); c->h_rrec.clear();
    for (size_t i = 0; i < rpos.size(); i++) {
        if (rpos[i] >= c->node_lo && rpos[i] < c->node_hi) {
            c->h_rpos.push_back(rpos[i]);
            c->h_rrec.push_back(rrec[3 * i]); c->h_rrec.push_back(rrec[3 * i + 1]); c->h_rrec.push_back(rrec[3 * i + 2]);
        }
    }
    c->n_upper_records = c->h_rpos.size();
    // ---- data file range of this rank ----
    c->h_dpos.clear(); c->h_drec.clear();
    if (levels) {
        // first data record of a top tile's subtree: its first leaf (payload) or its first internal node (binary), i.e. the
        // record with `leaves before` leaves and IB internal nodes in front of it
        auto first_data_from = [&](ull e0) -> ull {
            for (ull e = e0; e < WJ; e++) if (M[J][e]) return data_index(LT[J][e] - T[ts * e + 2], IB[J][e]);
            return c->n_data;
        };
        c->data_lo = c->rank == 0 ? 0 : first_data_from(wj0);
        c->data_hi = c->rank == c->world - 1 ? c->n_data : first_data_from(wj1);
        if (leaves_total == 0) { c->data_lo = c->rank == 0 ? 0 : c->n_data; c->data_hi = c->n_data; }
        for (size_t i = 0; i < dpos.size(); i++) {
            if (dpos[i] >= c->data_lo && dpos[i] < c->data_hi) {
                c->h_dpos.push_back(dpos[i]);
                for (int q = 0; q < 6; q++) c->h_drec.push_back(drec[6 * i + q]);
            }
        }
    } else if (payload) {
        c->data_lo = c->rank == 0 ? 0 : 1 + c->leaf_offset;
        c->data_hi = 1 + c->leaf_offset + c->n_voxels_local;
    } else {
        c->data_lo = 0; c->data_hi = c->n_data;
        if (c->rank != 0) c->data_lo = c->data_hi = c->n_data;
    }
}

static int shard_host_merge(svo_ctx* c, const ull* table) {
    const int J = c->J;
    const ull WJ = c->WJ;
    c->h_table.resize((size_t)WJ * c->tstride);
    CK(cudaMemcpyAsync(c->h_table.data(), table, (size_t)WJ * c->tstride * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    shard_merge_compute(c, c->h_table.data());
    if (c->h_ownbase.size() != c->lv[J].n) return fail(c, SVO_E_INVALID, "sharded merge: table does not match this rank's tiles (was the table summed over all ranks?)");
    if (!c->h_ownbase.empty())
        CK(cudaMemcpyAsync(c->lv[J].base.p, c->h_ownbase.data(), c->h_ownbase.size() * sizeof(ull), cudaMemcpyHostToDevice, c->stream));
    if (!c->h_ownibase.empty())
        CK(cudaMemcpyAsync(c->lv[J].ibase.p, c->h_ownibase.data(), c->h_ownibase.size() * sizeof(ull), cudaMemcpyHostToDevice, c->stream));
    if (c->n_upper_records) {
        CK(c->d_rpos.ensure(c->n_upper_records * sizeof(ull)));
        CK(c->d_rrec.ensure(c->n_upper_records * 3 * sizeof(ull)));
        CK(cudaMemcpyAsync(c->d_rpos.p, c->h_rpos.data(), c->n_upper_records * sizeof(ull), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_rrec.p, c->h_rrec.data(), c->n_upper_records * 3 * sizeof(ull), cudaMemcpyHostToDevice, c->stream));
    }
    return SVO_OK;
}

// Tail of every build: tell the peers this rank is done reading, fetch the queue statistics, the one final
// synchronisation, stage times. `post_done`: false when the caller may still have to repeat the build (the peers'
// slices must stay readable).
static int finish_build_sync(svo_ctx* c, bool post_done) {
    c->h_pinned[48] = 0;
    if (c->sliced && post_done) {
        // this rank has finished reading its peers' slices and lists: they may be rewritten for the next job
        k_slice_post<<<1, MAX_WORLD, 0, c->stream>>>(c->sj, 1); LAUNCHED();
    }
    if (c->sliced) CK(cudaMemcpyAsync(c->h_pinned + 48, &((SliceCtrl*)c->sl_ctrl.p)->error, sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
    // queue statistics
    CK(cudaMemcpyAsync(c->h_pinned + 40, c->qcount.p, 4 * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
    c->tl.stamp("sync3_wait");
    CK(cudaStreamSynchronize(c->stream));
    c->tl.stamp("sync3_done");
    if (c->h_pinned[48]) {
        CK(cudaMemsetAsync(&((SliceCtrl*)c->sl_ctrl.p)->error, 0, sizeof(ull), c->stream));
        c->dense_clean = false;
        return fail(c, SVO_E_CUDA, "remote triangle slices: timed out waiting for a peer rank");
    }
    if (c->h_pinned[43]) {
        c->dense_clean = false;
        return fail(c, SVO_E_NOMEM, "work queue overflow: too many medium/large triangle-partition pairs for inline enumeration; "
                                    "set SVO_PARTITION_LISTS=1 to build exact per-partition lists");
    }
    return SVO_OK;
}
static void finish_build_stats(svo_ctx* c) {
    c->tl.dump();
    c->phase_a_done = false;
    c->built = true;
    c->voxelized = false;            // the pyramid has been consumed (and cleared): a new build needs a new svo_voxelize
    c->stats.n_partitions = c->P;
    // inline enumeration does not count pairs on the device; the sum of the per-partition counts is known when they were requested
    c->stats.n_pairs = (c->P > 1 && !c->use_lists) ? c->n_pairs : c->q_end - c->q_begin;
    c->stats.n_voxels = c->n_voxels; c->stats.n_nodes = c->n_nodes; c->stats.n_data = c->n_data;
    c->stats.n_bricks = c->lv[0].n; c->stats.n_tiles1 = c->J >= 1 ? c->lv[1].n : 0;
    c->stats.speculative = (c->fast && c->spec) ? 1u : 0u;
    c->stats.n_brick_records = c->fast ? c->h_info->n_brick_records : 0;
    c->stats.n_medium = c->h_pinned[40]; c->stats.n_large = c->h_pinned[41];
    const ull queued = c->stats.n_medium + c->stats.n_large;
    c->stats.n_small = c->stats.n_pairs >= queued ? c->stats.n_pairs - queued : 0;   // n_pairs is 0 when the pairs were not counted
    c->stats.ms_upload = span(c, EV_UP0, EV_UP1);
    c->stats.ms_partition = span(c, EV_PART0, EV_PART1);
    c->stats.ms_voxelize = span(c, EV_VOX0, EV_VOX1);
    c->stats.ms_build = span(c, EV_BUILD0, EV_BUILD1);
    c->stats.ms_emit = span(c, EV_EMIT0, EV_EMIT1);
    c->stats.ms_clear = span(c, EV_CLR0, EV_CLR1);
    c->stats.ms_download = 0.f;
    c->stats.ms_vox_small = span(c, EV_VS0, EV_VS1);
    c->stats.ms_emit_leaf = c->lv[0].n ? span(c, EV_EL0, EV_EL1) : 0.f;
    c->stats.ms_compact = span(c, EV_BUILD0, EV_CMP1);
    c->stats.ms_peer_wait = c->sliced ? span(c, EV_PW0, EV_PW1) : 0.f;
    c->stats.ms_dispatch = (c->dispatched || c->sliced) ? span(c, EV_DSP0, EV_DSP1) : 0.f;
    c->stats.kernel_launches = c->launches + ((c->dispatched || c->sliced) ? c->dispatch_launches : 0);
}

// ---------------------------------------------------------------------------
// Build, phase B: replicated upper levels from the (summed) table, file bases top-down, emission.
// ---------------------------------------------------------------------------
static int build_phase_b(svo_ctx* c, const ull* table) {
    const int J = c->J, nl = c->nl, top = nl - 1;
    const bool payload = c->prm.payload != 0, levels = c->prm.generate_levels != 0;
    const bool want_pl = c->want_pl;
    const bool host_merge = c->world > 1;                 // sharded: upper levels merged on the host from the table
    const bool upper = J < top && !host_merge;
    const bool d_even = (c->D % 2) == 0;
    ull goff = 0, n_gJ = c->lv[J].n;
    c->leaf_offset = 0;
    if (host_merge) {
        int rc = shard_host_merge(c, table);
        if (rc) return rc;
    }
    if (upper) {
        // ---- dense columns of the global level-J words + replicated upper pyramid ----
        for (int q = 0; q < 4; q++) CK(c->dcol[q].ensure((size_t)c->WJ * sizeof(ull)));
        for (int j = J + 1; j < nl; j++) CK(cudaMemsetAsync(c->dense[j].p, 0, (size_t)c->nwords[j] * sizeof(ull), c->stream));
        k_table_unpack<<<blocks_for(c->WJ, 256), 256, 0, c->stream>>>(table, c->WJ, c->dcol[0].as<ull>(), c->dcol[1].as<ull>(), c->dcol[2].as<ull>(),
                                                                       c->dcol[3].as<ull>(), (ull* const*)c->d_lvlptrs.p, J + 1, nl); LAUNCHED();
        // the host needs the global tile count of every upper level and this rank's offsets: the table is tiny
        if (c->world > 1) {
            c->h_table.resize((size_t)c->WJ * 4);
            CK(cudaMemcpyAsync(c->h_table.data(), table, (size_t)c->WJ * 4 * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            n_gJ = 0;
            const ull wj0 = c->bias[J], wj1 = c->bias[J] + c->nwords[J];
            for (ull e = 0; e < c->WJ; e++) {
                if (!c->h_table[e * 4]) continue;
                if (e < wj0) { goff++; c->leaf_offset += c->h_table[e * 4 + 2]; }
                (void)wj1;
                n_gJ++;
            }
        }
        // tile counts of the replicated upper levels
        if (c->world == 1) {
            // one GPU has at most ONE upper level: the virtual top word of an odd depth
            for (int j = J + 1; j < nl; j++) { int rc = alloc_level(c, c->lv[j], c->lv[J].n ? 1 : 0, want_pl, levels); if (rc) return rc; }
        } else {
            std::vector<ull> cur((size_t)c->WJ);
            for (ull e = 0; e < c->WJ; e++) cur[e] = c->h_table[e * 4] ? 1 : 0;
            for (int j = J + 1; j < nl; j++) {
                std::vector<ull> nxt((size_t)c->nwords[j], 0);
                for (size_t e = 0; e < cur.size(); e++) if (cur[e]) nxt[e >> 6] = 1;
                ull n = 0; for (ull v : nxt) n += v;
                int rc = alloc_level(c, c->lv[j], n, want_pl, levels); if (rc) return rc;
                cur.swap(nxt);
            }
        }
        int rc = alloc_level(c, c->glv, n_gJ, want_pl, levels);
        if (rc) return rc;
        // ---- upper compact lists (tiny), ending in the GLOBAL level-J list ----
        for (int j = top; j > J; j--) {
            if (c->lv[j].n) { k_compact_top<<<1, 1024, 0, c->stream>>>(c->dense[j].as<ull>(), c->nwords[j], 0, c->lv[j].key.as<ull>(), c->lv[j].mask.as<ull>(), nullptr, 0); LAUNCHED(); }
        }
        for (int j = top; j > J; j--) {
            PopcOp op{ c->lv[j].mask.as<ull>() };
            if ((rc = exscan(c, op, c->lv[j].n, c->lv[j].fc.as<ull>()))) return rc;
        }
        // global level-J list: keys and masks from the dense mask column, in key order
        if (c->glv.n) { k_compact_top<<<1, 1024, 0, c->stream>>>(c->dcol[0].as<ull>(), c->WJ, 0, c->glv.key.as<ull>(), c->glv.mask.as<ull>(), nullptr, 0); LAUNCHED(); }
        {
            DenseColOp sop{ c->dcol[1].as<ull>(), c->glv.key.as<ull>() };
            if ((rc = exscan(c, sop, c->glv.n, c->glv.ps.as<ull>()))) return rc;
            if (want_pl) { DenseColOp lop{ c->dcol[2].as<ull>(), c->glv.key.as<ull>() }; if ((rc = exscan(c, lop, c->glv.n, c->glv.pl.as<ull>()))) return rc; }
            if (levels) { DenseColOp iop{ c->dcol[3].as<ull>(), c->glv.key.as<ull>() }; if ((rc = exscan(c, iop, c->glv.n, c->glv.pi.as<ull>()))) return rc; }
        }
        for (int j = J + 1; j < nl; j++) {
            if ((rc = size_scans(c, c->lv[j], j == J + 1 ? &c->glv : &c->lv[j - 1], want_pl, levels))) return rc;
        }
    }
    // ---- sync #2: record counts ----
    LevelBufs& topL = c->lv[top];
    if (!host_merge) {
        CK(cudaMemcpyAsync(c->h_pinned + 32, c->lv[0].fc.as<ull>() + c->lv[0].n, sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h_pinned + 33, topL.ps.as<ull>() + topL.n, sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
        if (levels) CK(cudaMemcpyAsync(c->h_pinned + 34, topL.pi.as<ull>() + topL.n, sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
        if (want_pl) CK(cudaMemcpyAsync(c->h_pinned + 35, topL.pl.as<ull>() + topL.n, sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
        c->tl.stamp("sync2_wait");
        CK(cudaStreamSynchronize(c->stream));
        c->tl.stamp("sync2_done");
        c->n_voxels_local = c->h_pinned[32];
        c->n_voxels = c->n_voxels_local;
        const ull s_top = c->h_pinned[33];
        c->n_nodes = c->n_voxels == 0 ? 1 : s_top + (d_even ? 1 : 0);
    }
    if (!host_merge) {
        c->n_data = payload ? 1 + c->n_voxels : 2;          // OctreeBuilder.cpp:25-29
        if (levels && c->n_voxels) c->n_data += c->h_pinned[34] - (d_even ? 0 : 1);   // one record per internal node (the virtual top word is no node)
    }                                                       // (sharded: shard_merge_compute)

    // level-J view of this rank's tiles inside the global list
    Level LJ = c->lv[J].view();
    Level GJ = c->glv.view();
    if (upper) {
        LJ.key = GJ.key + goff; LJ.mask = GJ.mask + goff; LJ.ps = GJ.ps + goff; LJ.base = GJ.base + goff;
        LJ.pl = want_pl ? GJ.pl + goff : nullptr;
        LJ.pi = levels ? GJ.pi + goff : nullptr;
        LJ.ibase = levels ? GJ.ibase + goff : nullptr;
        LJ.cache = levels ? GJ.cache + goff * 6 : nullptr;
    }
    mark(c, EV_EMIT0);
    EmitJob E;
    memset(&E, 0, sizeof E);
    E.leaf_data_mode = payload ? 1 : 0;
    E.levels = levels ? 1 : 0;
    E.virtual_top = d_even ? 0 : 1;
    E.leaf_offset = c->leaf_offset;
    E.pos_lo = 0; E.pos_hi = ~0ULL; E.cap = ~0ULL;
    E.write_records = 1;
    if (!host_merge) {
        c->node_lo = 0; c->node_hi = c->n_nodes;
        if (c->n_voxels) {
            CK(cudaMemsetAsync(topL.base.p, 0, sizeof(ull), c->stream));
            if (levels) CK(cudaMemsetAsync(topL.ibase.p, 0, sizeof(ull), c->stream));
        }
    }
    auto emit_upper_levels = [&](void) -> int {
        if (host_merge) {
            // the shared upper levels were merged on the host: scatter the records that fall into this rank's range
            if (c->n_upper_records) {
                k_scatter_records<<<blocks_for(c->n_upper_records, 256), 256, 0, c->stream>>>(c->d_rpos.as<ull>(), c->d_rrec.as<ull>(), c->n_upper_records, (const ull*)nullptr, E); LAUNCHED();
            }
            return SVO_OK;
        }
        for (int j = top; j > J; j--) {
            if (!c->lv[j].n) continue;
            E.is_top = (j == top);
            E.root_here = (j == top) && d_even;
            k_emit_upper<<<blocks_for(c->lv[j].n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(c->lv[j].view(), j == J + 1 ? GJ : c->lv[j - 1].view(), E); LAUNCHED();
        }
        return SVO_OK;
    };
    const ull n_local_nodes = c->node_hi - c->node_lo;
    CK(c->nodes.ensure((size_t)(n_local_nodes ? n_local_nodes : 1) * SVO_NODE_BYTES));
    E.nodes = c->nodes.as<ull>();
    E.pos_lo = c->node_lo; E.pos_hi = c->node_hi; E.cap = n_local_nodes;
    auto emit_all = [&](void) -> int {
    if (c->n_voxels == 0) {
        // empty grid: finalizeTree pads everything and writes a null root (OctreeBuilder.cpp:36-42)
        static const ull null_root[3] = { 0ULL, 0ULL, ~0ULL };
        if (n_local_nodes) CK(cudaMemcpyAsync(c->nodes.p, null_root, sizeof null_root, cudaMemcpyHostToDevice, c->stream));
    } else {
        int rc = emit_upper_levels();
        if (rc) return rc;
        int j_start = J;
        if (c->jf && c->jf < J) {
            // fused single-block emission of the small local levels J .. jf+1
            FusedJob F;
            memset(&F, 0, sizeof F);
            for (int j = 0; j <= J; j++) F.lv[j] = (j == J) ? LJ : c->lv[j].view();
            F.J = J; F.jf = c->jf;
            F.E = E; F.E.is_top = (J == top); F.E.root_here = (J == top) && d_even;
            k_fused_emit<<<1, 1024, 0, c->stream>>>(F); LAUNCHED();
            j_start = c->jf;
        }
        for (int j = j_start; j >= 1; j--) {
            if (!c->lv[j].n) continue;
            E.is_top = (j == top);
            E.root_here = (j == top) && d_even;
            k_emit_upper<<<blocks_for(c->lv[j].n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(j == J ? LJ : c->lv[j].view(), c->lv[j - 1].view(), E); LAUNCHED();
        }
        if (c->lv[0].n) {
            E.is_top = (top == 0);
            E.root_here = (top == 0) && d_even;
            const Level L0 = (J == 0) ? LJ : c->lv[0].view();
            mark(c, EV_EL0);
            if (levels) { k_emit_leaf_levels<<<blocks_for(L0.n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(L0, E); LAUNCHED(); }
            else if (payload) { k_emit_leaf<true, 4, false><<<blocks_for(L0.n, WARPS_PER_BLOCK * EMIT_TILES_PER_WARP), WARPS_PER_BLOCK * 32, 0, c->stream>>>(L0, E); LAUNCHED(); }
            else { k_emit_leaf<false, 4, false><<<blocks_for(L0.n, WARPS_PER_BLOCK * EMIT_TILES_PER_WARP), WARPS_PER_BLOCK * 32, 0, c->stream>>>(L0, E); LAUNCHED(); }
            mark(c, EV_EL1);
        }
    }
    return SVO_OK;
    };
    {
        int rc = emit_all();
        if (rc) return rc;
    }
    mark(c, EV_EMIT1);
    // ---- data records ----
    // This rank's range of the data file: everything on one GPU; sharded: shard_merge_compute set it (payload: the leaf
    // records of its slab; -levels: leaves and internal records interleaved in post-order, the shared upper levels'
    // records fall to the rank whose range holds their index).
    if (!host_merge) {
        c->data_lo = 0; c->data_hi = c->n_data;
    }
    const ull n_local_data = c->data_hi - c->data_lo;
    CK(c->data.ensure((size_t)(n_local_data ? n_local_data : 1) * SVO_DATA_BYTES));
    float* const data_biased = c->data.as<float>() - c->data_lo * 8;          // record index = global data index
    if (!payload) {
        static const uint32_t white[16] = { 0, 0, 0, 0, 0, 0, 0, 0,                         // record 0: NULL
                                            0, 0, 0x3f800000u, 0x3f800000u, 0x3f800000u, 0, 0, 0 };  // record 1: white voxel
        if (c->rank == 0) CK(cudaMemcpyAsync(c->data.p, white, sizeof white, cudaMemcpyHostToDevice, c->stream));
    } else {
        if (c->rank == 0) CK(cudaMemsetAsync(c->data.p, 0, SVO_DATA_BYTES, c->stream));
        if (c->n_voxels_local) {
            if (!c->owner_ready) {
                CK(c->owner.ensure((size_t)c->n_voxels_local * sizeof(uint32_t)));
                CK(cudaMemsetAsync(c->owner.p, 0xff, (size_t)c->n_voxels_local * sizeof(uint32_t), c->stream));
                int rc = launch_voxelizer<true>(c);
                if (rc) return rc;
            }
            PayloadJob Pj;
            memset(&Pj, 0, sizeof Pj);
            Pj.tris = c->d_tris; Pj.owner = c->owner.as<uint32_t>();
            if (c->sliced) {
                Pj.segs.n = c->world;
                for (int r = 0; r < c->world; r++) Pj.segs.ptr[r] = c->peer_slice[r];
                Pj.segs.nslice = ((const SliceCtrl*)c->sl_ctrl.p)->nslice;
            }
            Pj.data = data_biased;
            Pj.unit_div = c->unit_div; Pj.gridsize_f = (float)c->prm.gridsize; Pj.color_mode = c->prm.color_mode;
            Pj.levels = levels ? 1 : 0;
            Pj.leaf_offset = c->leaf_offset;
            const Level L0 = (J == 0) ? LJ : c->lv[0].view();
            k_payload<<<blocks_for(L0.n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(L0, Pj); LAUNCHED();
        }
    }
    c->owner_ready = false;
    if (levels && c->n_voxels) {
        // internal-node data records, bottom-up (needs the leaf records and the ibase values written above); sharded: the
        // local levels here, the shared upper levels' records come from the host-side merge
        E.is_top = 0; E.root_here = 0;
        for (int j = 0; j < (host_merge ? J + 1 : nl); j++) {
            if (!c->lv[j].n) continue;
            const int real_node = !(j == top && !d_even);
            const Level L = (j == J) ? LJ : c->lv[j].view();
            const Level C = j == 0 ? L : (j - 1 == J ? (upper ? GJ : LJ) : c->lv[j - 1].view());
            k_levels_data<<<blocks_for(L.n, 128), 128, 0, c->stream>>>(L, C, j, E, data_biased, real_node, 0); LAUNCHED();
        }
        if (host_merge && !c->h_dpos.empty()) {
            const size_t nrec = c->h_dpos.size();
            CK(c->d_dpos.ensure(nrec * sizeof(ull)));
            CK(c->d_drec.ensure(nrec * 6 * sizeof(float)));
            CK(cudaMemcpyAsync(c->d_dpos.p, c->h_dpos.data(), nrec * sizeof(ull), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_drec.p, c->h_drec.data(), nrec * 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            k_scatter_data_records<<<blocks_for(nrec, 256), 256, 0, c->stream>>>(c->d_dpos.as<ull>(), c->d_drec.as<float>(), nrec, data_biased); LAUNCHED();
        }
    }
    mark(c, EV_BUILD1);
    // ---- leave a clean pyramid behind: zero exactly the words that were set ----
    mark(c, EV_CLR0);
    if (c->lv[0].n) {
        ClearJob Cj;
        memset(&Cj, 0, sizeof Cj);
        for (int j = 0; j <= J; j++) { Cj.key[j] = c->lv[j].key.as<ull>(); Cj.dense[j] = c->dense[j].as<ull>() - c->bias[j]; Cj.n[j] = c->lv[j].n; }
        dim3 grid((unsigned)std::min<ull>(blocks_for(c->lv[0].n, 256), (ull)c->sm_count * 16), (unsigned)(J + 1));
        k_sparse_clear_all<<<grid, 256, 0, c->stream>>>(Cj); LAUNCHED();
    }
    c->dense_clean = true;
    mark(c, EV_CLR1);
    {
        int rc = finish_build_sync(c, true);
        if (rc) return rc;
    }
    finish_build_stats(c);
    return SVO_OK;
}

// ---------------------------------------------------------------------------
// Device-driven build (svo_build.cuh). Phase A: tile lists of the local levels straight from the dense pyramid, subtree
// sizes, this rank's table entries. Phase B: merged upper levels (k_shard_merge), emission, clear.
// In steady state (list and node-buffer capacities known from the previous build, binary mode) nothing between the
// voxelizer launch and the final synchronisation waits for the host. The first build of a context, payload builds
// and builds that outgrow the capacities take the same kernels with two read-backs (counts, record range).
// ---------------------------------------------------------------------------
static bool fast_path_applies(const svo_ctx* c) {
    if (c->prm.generate_levels || c->J < 1) return false;
    const char* e = getenv("SVO_BUILD_PATH");
    return !(e && strcmp(e, "classic") == 0);
}
static Level fast_view(svo_ctx* c, int j) {
    Level L = c->lv[j].view();
    L.np = &c->info_buf.as<BuildInfo>()->count[j];
    L.cap = c->fcap[j];
    L.n = c->fcap[j];
    L.clear = c->dense[j].as<ull>() - c->bias[j];      // the emitter of every level clears the words of its tiles
    return L;
}
static int read_info(svo_ctx* c, const char* stamp) {
    CK(cudaMemcpyAsync(c->h_info, c->info_buf.p, sizeof(BuildInfo), cudaMemcpyDeviceToHost, c->stream));
    c->tl.stamp(stamp);
    CK(cudaStreamSynchronize(c->stream));
    return SVO_OK;
}
static int fast_check_info(svo_ctx* c) {
    const ull o = c->h_info->overflow;
    if (o & (1ULL << 40)) { c->dense_clean = false; return fail(c, SVO_E_CUDA, "octree build: a look-back scan timed out"); }
    if (o & (1ULL << 41)) { c->dense_clean = false; return fail(c, SVO_E_INVALID, "sharded merge: table does not match this rank's tiles (was the table exchanged / summed over all ranks?)"); }
    if (o & (1ULL << 42)) { c->dense_clean = false; return fail(c, SVO_E_RANGE, "sharded merge: upper-level record buffer too small"); }
    return SVO_OK;
}

// The single-block stages of a build (k_top, svo_build.cuh) for the current geometry and buffers.
static int make_topjob(svo_ctx* c, ull* table, TopJob& P) {
    const int J = c->J, nl = c->nl, top = nl - 1, jB = c->jB;
    const bool d_even = (c->D % 2) == 0;
    BuildInfo* dinfo = c->info_buf.as<BuildInfo>();
    memset(&P, 0, sizeof P);
    P.info = dinfo;
    // small levels jB+1..J: lists
    P.S.j0 = jB + 1; P.S.J = J; P.S.count_only = 0; P.S.info = dinfo;
    for (int j = jB; j <= J; j++) {
        P.S.dense[j] = c->dense[j].as<ull>(); P.S.nwords[j] = c->nwords[j]; P.S.bias[j] = c->bias[j];
        P.S.key[j] = c->lv[j].key.as<ull>(); P.S.mask[j] = c->lv[j].mask.as<ull>(); P.S.fc[j] = c->lv[j].fc.as<ull>();
        P.S.cap[j] = c->fcap[j];
    }
    // sizes of jB+1..J, emission of the levels with at most a few hundred tiles (jE+1..J); the rest is emitted by one
    // multi-block launch per level
    int jE = jB;
    for (int j = jB + 1; j <= J; j++) if (c->nwords[j] > 256) jE = j;
    c->jE = jE;
    for (int j = jB; j <= J; j++) { P.F.lv[j] = fast_view(c, j); P.F.lv[j].pl = nullptr; }
    P.F.J = J; P.F.jf = jE; P.F.jf_up = jB + 1;
    EmitJob& E = P.F.E;
    E.nodes = c->nodes.as<ull>();
    E.cap = c->nodes.cap / SVO_NODE_BYTES;
    E.info = dinfo;
    E.leaf_data_mode = c->prm.payload ? 1 : 0;
    E.virtual_top = d_even ? 0 : 1;
    E.write_records = 1;
    E.is_top = (J == top); E.root_here = ((J == top) && d_even) ? 1 : 0;
    // table entries of this rank
    P.T.key = c->lv[J].key.as<ull>(); P.T.mask = c->lv[J].mask.as<ull>(); P.T.ps = c->lv[J].ps.as<ull>();
    for (int j = 0; j <= J; j++) P.T.fc[j] = c->lv[j].fc.as<ull>();
    P.T.n = c->fcap[J]; P.T.np = &dinfo->count[J]; P.T.J = J; P.T.table = table; P.T.info = dinfo;
    P.table_words = c->WJ * 4;
    // merge of the shared upper levels
    ull n_scratch = 0, rcap = 2;
    for (int j = J; j <= top; j++) n_scratch += (j == J ? c->WJ : c->nwords[j]);
    for (int j = J + 1; j <= top; j++) rcap += c->nwords[j] * 73;
    CK(c->merge_scratch.ensure((size_t)n_scratch * 3 * sizeof(ull)));
    CK(c->merge_rpos.ensure((size_t)rcap * sizeof(ull)));
    CK(c->merge_rrec.ensure((size_t)rcap * 3 * sizeof(ull)));
    MergeJob& Mj = P.M;
    Mj.table = table; Mj.WJ = c->WJ;
    Mj.J = J; Mj.top = top; Mj.d_even = d_even ? 1 : 0; Mj.rank = c->rank; Mj.world = c->world;
    ull off = 0;
    for (int j = J; j <= top; j++) {
        Mj.nW[j] = j == J ? c->WJ : c->nwords[j];
        Mj.M[j] = c->merge_scratch.as<ull>() + off; Mj.S[j] = Mj.M[j] + n_scratch; Mj.B[j] = Mj.S[j] + n_scratch;
        off += Mj.nW[j];
    }
    Mj.wj0 = c->bias[J]; Mj.wj1 = c->bias[J] + c->nwords[J];
    Mj.rpos = c->merge_rpos.as<ull>(); Mj.rrec = c->merge_rrec.as<ull>(); Mj.rcap = rcap;
    Mj.keyJ = c->lv[J].key.as<ull>(); Mj.baseJ = c->lv[J].base.as<ull>(); Mj.capJ = c->fcap[J];
    Mj.info = dinfo;
    Mj.nodes_cap = c->spec ? c->nodes.cap / SVO_NODE_BYTES : ~0ULL;
    if (c->sliced && c->sl_attached) {
        XchgJob& X = P.X;
        X.src = table;
        X.lo = c->bias[J] * (ull)c->tstride; X.n = c->nwords[J] * (ull)c->tstride;
        X.world = c->world; X.me = c->rank; X.epoch = c->xchg_epoch;
        X.info = dinfo;
        for (int r = 0; r < c->world; r++) { X.xtable[r] = c->peer_xtable[r]; X.ctrl[r] = c->peer_slctrl[r]; }
        P.own_ctrl = (SliceCtrl*)c->sl_ctrl.p;
        if (c->xwait_pending) Mj.table = c->sl_xtable.as<ull>();      // the complete table lives in this rank's exchange window
    }
    return SVO_OK;
}
static int launch_top(svo_ctx* c, ull* table, int stages) {
    if (!stages) return SVO_OK;
    TopJob P;
    int rc = make_topjob(c, table, P);
    if (rc) return rc;
    P.stages = stages;
    k_top<<<1, 1024, 0, c->stream>>>(P); LAUNCHED();
    return SVO_OK;
}

static int fast_phase_a(svo_ctx* c, ull* table, bool fill_table, bool force_sync) {
    const int J = c->J;
    const bool payload = c->prm.payload != 0;
    c->want_pl = false;
    c->jf = 0;
    int jB = 1;
    for (int j = 2; j <= J; j++) if (c->nwords[j] > SMALL_LEVEL_WORDS) jB = j;
    c->jB = jB;
    CK(c->info_buf.ensure(sizeof(BuildInfo)));
    BuildInfo* dinfo = c->info_buf.as<BuildInfo>();
    // sharded: a speculative local build that aborts leaves this rank without table entries; only the peer-memory
    // exchange can tell the peers (poisoned flag, svo_dispatch.cuh), so other exchanges get sized builds
    // (a context whose previous job went through svo_shard_exchange is taken to keep doing so)
    bool spec = c->fast_caps_ok && !payload && !force_sync && c->nodes.cap > 0 && (c->world == 1 || (c->sliced && c->uses_peer_exchange));
    c->exchanged_by_peer_memory = false;
    if (const char* e = getenv("SVO_SPECULATIVE_BUILD")) spec = spec && e[0] != '0';
    c->spec = spec;
    mark(c, EV_BUILD0);
    CK(cudaMemsetAsync(dinfo, 0, sizeof(BuildInfo), c->stream));
    auto dense_pass = [&](int count_only) -> int {
        for (int j = 1; j <= jB; j++) {
            const bool big = c->nwords[j] > (4ULL << 20);                     // > 32 MiB of words: four sub-tiles per block
            const ull tile_words = (ull)DS_THREADS * DS_ITEMS * (big ? 4 : 1);
            const ull nt = (c->nwords[j] + tile_words - 1) / tile_words;
            int rc = lookback_prepare(c, nt);
            if (rc) return rc;
            DenseScanJob Dj;
            memset(&Dj, 0, sizeof Dj);
            Dj.dense = c->dense[j].as<ull>(); Dj.n = c->nwords[j]; Dj.bias = c->bias[j];
            Dj.next = j < J ? c->dense[j + 1].as<ull>() - c->bias[j + 1] : nullptr;
            Dj.key = c->lv[j].key.as<ull>(); Dj.mask = c->lv[j].mask.as<ull>(); Dj.fc = c->lv[j].fc.as<ull>();
            Dj.cap = c->fcap[j]; Dj.cap_child = c->fcap[j - 1];
            Dj.j = j; Dj.count_only = count_only; Dj.info = dinfo;
            Dj.state = c->lb_state.as<ull>(); Dj.ticket = c->lb_ticket.as<ull>(); Dj.ticket_base = c->lb_tickets; Dj.epoch = c->lb_epoch;
            if (big) { k_dense_scan<4><<<(unsigned)nt, DS_THREADS, 0, c->stream>>>(Dj); LAUNCHED(); }
            else { k_dense_scan<1><<<(unsigned)nt, DS_THREADS, 0, c->stream>>>(Dj); LAUNCHED(); }
            c->lb_tickets += nt;
        }
        if (count_only && jB < J) {
            SmallLevelsJob S;
            memset(&S, 0, sizeof S);
            S.j0 = jB + 1; S.J = J; S.count_only = 1; S.info = dinfo;
            for (int j = jB; j <= J; j++) { S.dense[j] = c->dense[j].as<ull>(); S.nwords[j] = c->nwords[j]; S.bias[j] = c->bias[j]; }
            k_small_levels<<<1, 1024, 0, c->stream>>>(S); LAUNCHED();
        }
        return SVO_OK;
    };
    if (!spec) {
        // ---- counts first (read-back #1), then lists of exactly the right size (+ slack for the builds to come) ----
        int rc = dense_pass(1);
        if (rc) return rc;
        if ((rc = read_info(c, "sync1_wait"))) return rc;
        c->tl.stamp("sync1_done");
        if ((rc = fast_check_info(c))) return rc;
        for (int j = 0; j <= J; j++) {
            const ull n = c->h_info->count[j];
            const ull want = n + n / 8 + 64;
            if (want > c->fcap[j]) c->fcap[j] = want;
        }
        CK(cudaMemsetAsync(dinfo, 0, sizeof(BuildInfo), c->stream));
    }
    for (int j = 0; j <= J; j++) {
        int rc = alloc_level(c, c->lv[j], c->fcap[j], false, false);      // no-op once the lists are large enough
        if (rc) return rc;
        c->lv[j].n = 0;
    }
    if (payload) CK(c->tileidx.ensure((size_t)c->nwords[0] * sizeof(uint32_t)));
    int rc = dense_pass(0);
    if (rc) return rc;
    {   // bricks: lists, leaf ranks, subtree sizes of levels 0 and 1
        const ull n1 = spec ? std::min<ull>(c->fcap[1], c->nwords[1]) : c->h_info->count[1];
        CK(c->brick_lp.ensure((size_t)(c->fcap[1] + 2) * sizeof(ull)));
        CK(c->brick_sp.ensure((size_t)(c->fcap[1] + 2) * sizeof(ull)));
        BrickJob B;
        memset(&B, 0, sizeof B);
        B.L1 = fast_view(c, 1); B.L0 = fast_view(c, 0);
        B.dense0 = c->dense[0].as<ull>() - c->bias[0];
        B.tileidx = payload ? c->tileidx.as<uint32_t>() - c->bias[0] : nullptr;
        B.tile_lp = c->brick_lp.as<ull>(); B.tile_sp = c->brick_sp.as<ull>();
        B.info = dinfo;
        if (n1) { k_brick_gather<<<blocks_for(n1, BP_TILE), BP_WARPS * 32, 0, c->stream>>>(B); LAUNCHED(); }
        const ull nt = std::max<ull>((n1 + LB_TILE - 1) / LB_TILE, 1);
        if ((rc = lookback_prepare(c, nt))) return rc;
        BrickTileTotals g{ B.tile_lp, c->lv[1].mask.as<ull>() };
        k_scan_lookback<3><<<(unsigned)nt, LB_THREADS, 0, c->stream>>>(g, n1, &dinfo->count[1], B.tile_lp, B.tile_sp, c->lv[1].ps.as<ull>(),
                                                                      c->lb_state.as<ull>(), c->lb_ticket.as<ull>(), c->lb_tickets, c->lb_epoch, dinfo); LAUNCHED();
        c->lb_tickets += nt;
        int tpw = 32;                                  // tiles per warp: fewer when the level does not fill the GPU
        while (tpw > 2 && n1 / tpw < (ull)c->sm_count * 32) tpw >>= 1;
        k_brick_prefix<<<std::max(blocks_for(n1, BP_WARPS * tpw), 1u), BP_WARPS * 32, 0, c->stream>>>(B, tpw); LAUNCHED();
    }
    // subtree sizes of the big levels above: look-back scans (the small levels follow in k_top)
    for (int j = 2; j <= jB; j++) {
        SizeOp op{ c->lv[j].mask.as<ull>(), c->lv[j].fc.as<ull>(), c->lv[j - 1].ps.as<ull>() };
        const ull n = spec ? std::min<ull>(c->fcap[j], c->nwords[j]) : c->h_info->count[j];
        if ((rc = exscan(c, op, n, c->lv[j].ps.as<ull>(), &dinfo->count[j], dinfo))) return rc;
    }
    // ---- everything one block does: small-level lists and sizes, this rank's table entries. On one GPU the launch is
    // deferred to phase B, where the merge, the scattered upper records and the top-level emission join it ----
    c->top_pending = TOP_LISTS | TOP_SIZES | (fill_table ? TOP_TABLE : 0);
    if (fill_table) c->xwait_pending = false;              // (a repeated build keeps reading the exchanged table from the window)
    if (c->world > 1 && !(fill_table && c->sliced && c->uses_peer_exchange)) {
        // (a context that exchanges over peer memory keeps the stages pending: svo_shard_exchange launches them together with
        // the push of the table entries)
        if ((rc = launch_top(c, table, c->top_pending))) return rc;
        c->top_pending = 0;
    }
    mark(c, EV_CMP1);
    c->tl.stamp("phaseA_launched");
    c->phase_a_done = true;
    return SVO_OK;
}

static int fast_phase_b(svo_ctx* c, ull* table) {
    const int J = c->J;
    const bool payload = c->prm.payload != 0;
    BuildInfo* dinfo = c->info_buf.as<BuildInfo>();
    const bool spec = c->spec;
    int rc;
    if (c->world > 1 && c->top_pending)
        return fail(c, SVO_E_INVALID, "svo_shard_emit: this context exchanges its table with svo_shard_exchange, which was not called for this job");
    const int xw = c->xwait_pending ? TOP_XWAIT : 0;
    if (!spec) {
        // ---- merged upper levels; read-back #2: record counts and this rank's range ----
        if ((rc = launch_top(c, table, c->top_pending | xw | TOP_MERGE))) return rc;
        c->top_pending = 0;
        if (c->world == 1) mark(c, EV_CMP1);
        if ((rc = read_info(c, "sync2_wait"))) return rc;
        c->tl.stamp("sync2_done");
        if ((rc = fast_check_info(c))) return rc;
        if (c->h_info->overflow) { c->dense_clean = false; return fail(c, SVO_E_CUDA, "octree build: list capacity exceeded in a sized build (internal error)"); }
        const ull n_local = c->h_info->node_hi - c->h_info->node_lo;
        CK(c->nodes.ensure((size_t)(n_local ? n_local : 1) * SVO_NODE_BYTES));
        mark(c, EV_EMIT0);
        if ((rc = launch_top(c, table, TOP_SCATTER | TOP_EMIT))) return rc;
    } else {
        // ---- one launch: [small-level lists and sizes, table,] merge, scattered upper records, top-level emission ----
        mark(c, EV_EMIT0);
        if ((rc = launch_top(c, table, c->top_pending | xw | TOP_MERGE | TOP_SCATTER | TOP_EMIT))) return rc;
        c->top_pending = 0;
    }
    EmitJob E;
    {
        TopJob P;
        if ((rc = make_topjob(c, table, P))) return rc;
        E = P.F.E;
    }
    const int top = c->nl - 1;
    const bool d_even = (c->D % 2) == 0;
    auto launch_n = [&](int j) -> ull { return spec ? std::min<ull>(c->fcap[j], c->nwords[j]) : c->h_info->count[j]; };
    const int root_level_here = (J == top) && d_even;
    for (int j = c->jE; j >= 1; j--) {
        const ull n = launch_n(j);
        if (!n) continue;
        E.is_top = (j == top);
        E.root_here = (j == J) && root_level_here;
        // persistent warps over batches of bs tiles: full batches once the level has more tiles than the GPU holds warps
        const ull resident = (ull)c->sm_count * c->res_emit_upper;
        int bs = 32;
        while (bs > 1 && n / bs < resident * WARPS_PER_BLOCK) bs >>= 1;
        const unsigned grid = (unsigned)std::min<ull>(blocks_for(blocks_for(n, bs), WARPS_PER_BLOCK), resident);
        // (the records of the bricks -- the children of level 1 -- are written by k_emit_leaf)
        if (j == 1) { k_emit_upper_fast<false><<<grid, WARPS_PER_BLOCK * 32, 0, c->stream>>>(fast_view(c, j), fast_view(c, j - 1), E, bs); LAUNCHED(); }
        else { k_emit_upper_fast<true><<<grid, WARPS_PER_BLOCK * 32, 0, c->stream>>>(fast_view(c, j), fast_view(c, j - 1), E, bs); LAUNCHED(); }
    }
    E.is_top = 0; E.root_here = 0;
    mark(c, EV_EL0);
    if (launch_n(0)) {
        // persistent warps: at most the resident set (a sixth block per SM would run alone after the others finished)
        const unsigned grid = (unsigned)std::min<ull>(blocks_for(launch_n(0), WARPS_PER_BLOCK * EMIT_TILES_PER_WARP), (ull)c->sm_count * c->res_emit_leaf[payload ? 1 : 0]);
        const Level L0 = fast_view(c, 0);
        E.ticket = &dinfo->leaf_ticket;          // (zeroed with the rest of the BuildInfo when the build began)
        if (payload) k_emit_leaf<true, 4, true><<<grid, WARPS_PER_BLOCK * 32, 0, c->stream>>>(L0, E);
        else k_emit_leaf<false, 4, true><<<grid, WARPS_PER_BLOCK * 32, 0, c->stream>>>(L0, E);
        LAUNCHED();
    }
    mark(c, EV_EL1);
    mark(c, EV_EMIT1);
    // ---- data records ----
    if (!payload) {
        // 2 records, all on rank 0 (n_data is known without the counts)
        CK(c->data.ensure(2 * SVO_DATA_BYTES));
        static const uint32_t white[16] = { 0, 0, 0, 0, 0, 0, 0, 0,                         // record 0: NULL
                                            0, 0, 0x3f800000u, 0x3f800000u, 0x3f800000u, 0, 0, 0 };  // record 1: white voxel
        if (c->rank == 0) CK(cudaMemcpyAsync(c->data.p, white, sizeof white, cudaMemcpyHostToDevice, c->stream));
    } else {
        // payload builds are sized builds: the counts are on the host
        c->leaf_offset = c->h_info->leaf_offset;
        c->n_voxels_local = c->h_info->n_leaves_local;
        c->n_voxels = c->h_info->n_voxels;
        c->data_lo = c->rank == 0 ? 0 : 1 + c->leaf_offset;
        c->data_hi = 1 + c->leaf_offset + c->n_voxels_local;
        if (c->world == 1) { c->data_lo = 0; c->data_hi = 1 + c->n_voxels; }
        const ull n_local_data = c->data_hi - c->data_lo;
        CK(c->data.ensure((size_t)(n_local_data ? n_local_data : 1) * SVO_DATA_BYTES));
        if (c->rank == 0) CK(cudaMemsetAsync(c->data.p, 0, SVO_DATA_BYTES, c->stream));
        if (c->n_voxels_local) {
            CK(c->owner.ensure((size_t)c->n_voxels_local * sizeof(uint32_t)));
            CK(cudaMemsetAsync(c->owner.p, 0xff, (size_t)c->n_voxels_local * sizeof(uint32_t), c->stream));
            c->lv[0].n = c->h_info->count[0];
            rc = launch_voxelizer<true>(c);
            if (rc) return rc;
            PayloadJob Pj;
            memset(&Pj, 0, sizeof Pj);
            Pj.tris = c->d_tris; Pj.owner = c->owner.as<uint32_t>();
            if (c->sliced) {
                Pj.segs.n = c->world;
                for (int r = 0; r < c->world; r++) Pj.segs.ptr[r] = c->peer_slice[r];
                Pj.segs.nslice = ((const SliceCtrl*)c->sl_ctrl.p)->nslice;
            }
            Pj.data = c->data.as<float>() - c->data_lo * 8;
            Pj.unit_div = c->unit_div; Pj.gridsize_f = (float)c->prm.gridsize; Pj.color_mode = c->prm.color_mode;
            Pj.levels = 0;
            Pj.leaf_offset = c->leaf_offset;
            k_payload<<<blocks_for(launch_n(0), WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(fast_view(c, 0), Pj); LAUNCHED();
        }
    }
    mark(c, EV_BUILD1);
    // (the pyramid is clean again: every emitter zeroed the dense words of the tiles it handled; an aborted build
    // launched no emitter and keeps its pyramid)
    mark(c, EV_CLR0);
    mark(c, EV_CLR1);
    if (c->xwait_pending && c->xchg_user_table)      // the caller's table buffer ends up holding the complete table (off the critical path)
        CK(cudaMemcpyAsync(c->xchg_user_table, c->sl_xtable.p, (size_t)c->WJ * c->tstride * sizeof(ull), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->h_info, c->info_buf.p, sizeof(BuildInfo), cudaMemcpyDeviceToHost, c->stream));
    return SVO_OK;
}

// One fast build, phases A (unless already done by svo_shard_count) and B, with the repeat of a speculative build that
// outgrew its capacities. `table`: the subtree table (own entries on one GPU; complete after the exchange when sharded).
static int fast_build(svo_ctx* c, ull* table, bool phase_a_needed) {
    int rc;
    if (phase_a_needed && (rc = fast_phase_a(c, table, true, false))) return rc;
    if ((rc = fast_phase_b(c, table))) return rc;
    // (a speculative build that has to be repeated only re-reads this rank's own pyramid and the complete table: the
    // peers may be told right away that their slices are no longer needed)
    if ((rc = finish_build_sync(c, true))) return rc;
    if ((rc = fast_check_info(c))) return rc;
    if (c->h_info->overflow && c->world > 1 && (c->h_info->overflow & ((1ULL << 43) | 0xffffffffULL))) {
        // This rank's (bits 0..31) or a peer's (bit 43) speculative LOCAL build was aborted before the table exchange:
        // the exchanged table is incomplete on every rank, and every rank knows (poisoned exchange flag). All of them
        // return SVO_E_RETRY now; the pyramids are intact, the ranks that overflowed take a sized build next time.
        if (!c->spec && !(c->h_info->overflow & (1ULL << 43))) { c->dense_clean = false; return fail(c, SVO_E_CUDA, "octree build: list capacity exceeded in a sized build (internal error)"); }
        if (!c->exchanged_by_peer_memory) { c->dense_clean = false; return fail(c, SVO_E_CUDA, "octree build: speculative local build aborted without a poison-aware exchange (internal error)"); }
        if (c->h_info->overflow & 0xffffffffULL) c->fast_caps_ok = false;
        c->phase_a_done = false;
        return fail(c, SVO_E_RETRY, "a rank's tile lists outgrew the capacities of the previous build: repeat svo_shard_count / svo_shard_exchange / svo_shard_emit (the voxelized grid is kept)");
    }
    if (c->h_info->overflow) {
        if (!c->spec) { c->dense_clean = false; return fail(c, SVO_E_CUDA, "octree build: list capacity exceeded in a sized build (internal error)"); }
        // the tree outgrew the lists / node buffer of the earlier builds: every kernel behind the detection returned at
        // once, the pyramid is intact. Repeat with counts read back and exact sizes (the table is complete already).
        c->fast_caps_ok = false;
        if ((rc = fast_phase_a(c, table, c->world == 1, true))) return rc;
        if ((rc = fast_phase_b(c, table))) return rc;
        if ((rc = finish_build_sync(c, false))) return rc;
        if ((rc = fast_check_info(c))) return rc;
        if (c->h_info->overflow) { c->dense_clean = false; return fail(c, SVO_E_CUDA, "octree build: list capacity exceeded in a sized build (internal error)"); }
    }
    c->dense_clean = true;
    c->fast_caps_ok = true;
    const BuildInfo& I = *c->h_info;
    for (int j = 0; j <= c->J; j++) c->lv[j].n = I.count[j];
    c->n_voxels = I.n_voxels; c->n_nodes = I.n_nodes;
    c->n_voxels_local = I.n_leaves_local; c->leaf_offset = I.leaf_offset;
    c->node_lo = I.node_lo; c->node_hi = I.node_hi;
    c->n_upper_records = I.n_upper;
    const bool payload = c->prm.payload != 0;
    c->n_data = payload ? 1 + c->n_voxels : 2;          // OctreeBuilder.cpp:25-29
    if (!payload) {
        c->data_lo = 0; c->data_hi = c->n_data;
        if (c->rank != 0) c->data_lo = c->data_hi = c->n_data;
    }
    finish_build_stats(c);
    return SVO_OK;
}

int svo_build(svo_ctx* c, uint64_t* n_voxels, uint64_t* n_nodes, uint64_t* n_data) {
    if (!c) return SVO_E_INVALID;
    if (!c->voxelized) return fail(c, SVO_E_INVALID, "svo_build before svo_voxelize");
    if (c->world != 1) return fail(c, SVO_E_INVALID, "sharded context: use svo_shard_count / svo_shard_emit");
    CK(cudaSetDevice(c->device));
    c->fast = fast_path_applies(c);
    int rc;
    if (c->fast) {
        CK(c->table_own.ensure((size_t)c->WJ * 4 * sizeof(ull)));
        rc = fast_build(c, c->table_own.as<ull>(), true);
        if (rc) return rc;
    } else {
        const bool upper = c->J < c->nl - 1;
        if (upper) CK(c->table_own.ensure((size_t)c->WJ * 4 * sizeof(ull)));
        rc = build_phase_a(c, upper ? c->table_own.as<ull>() : nullptr);
        if (rc) return rc;
        rc = build_phase_b(c, upper ? c->table_own.as<ull>() : nullptr);
        if (rc) return rc;
    }
    if (n_voxels) *n_voxels = c->n_voxels;
    if (n_nodes) *n_nodes = c->n_nodes;
    if (n_data) *n_data = c->n_data;
    return SVO_OK;
}

int svo_shard_table_size(svo_ctx* c, uint64_t* n_u64) {
    if (!c || !n_u64) return SVO_E_INVALID;
    if (!c->partitioned) return fail(c, SVO_E_INVALID, "svo_shard_table_size before svo_partition");
    *n_u64 = c->WJ * c->tstride;
    return SVO_OK;
}

int svo_shard_count(svo_ctx* c, uint64_t* dev_table) {
    if (!c) return SVO_E_INVALID;
    if (!c->voxelized) return fail(c, SVO_E_INVALID, "svo_shard_count before svo_voxelize");
    CK(cudaSetDevice(c->device));
    if (!dev_table) { CK(c->table_own.ensure((size_t)c->WJ * c->tstride * sizeof(ull))); dev_table = c->table_own.as<uint64_t>(); }      // library-owned table
    c->fast = fast_path_applies(c);
    if (c->fast) return fast_phase_a(c, (ull*)dev_table, true, false);
    return build_phase_a(c, (ull*)dev_table);
}

int svo_shard_exchange(svo_ctx* c, uint64_t* dev_table) {
    if (!c) return SVO_E_INVALID;
    if (!c->phase_a_done) return fail(c, SVO_E_INVALID, "svo_shard_exchange before svo_shard_count");
    if (!c->sliced) return fail(c, SVO_E_INVALID, "svo_shard_exchange needs the peer windows of svo_shard_slice_*; otherwise sum the table with your own collective");
    if (!dev_table) dev_table = c->table_own.as<uint64_t>();
    if (!dev_table) return fail(c, SVO_E_INVALID, "dev_table is NULL");
    const size_t bytes = (size_t)c->WJ * c->tstride * sizeof(ull);
    if (bytes > XTABLE_BYTES) return fail(c, SVO_E_RANGE, "subtree table is larger than the exchange window; sum it with your own collective");
    CK(cudaSetDevice(c->device));
    if (c->fast) {
        // device-driven build: the push rides in the single-block top kernel (with the deferred stages of the local build, if
        // any), the wait for the peers' entries rides in svo_shard_emit's: no kernel of its own, no copy of the table
        ++c->xchg_epoch;                                    // every rank calls the exchange the same number of times
        c->xchg_user_table = (ull*)dev_table;
        int rc = launch_top(c, (ull*)dev_table, c->top_pending | TOP_XPUSH);
        if (rc) return rc;
        c->top_pending = 0;
        c->xwait_pending = true;
        c->exchanged_by_peer_memory = true;
        c->uses_peer_exchange = true;
        return SVO_OK;
    }
    // own entries = the level-J words of this rank's slab: a contiguous range of the table; the ranges of all ranks tile it
    XchgJob X;
    memset(&X, 0, sizeof X);
    X.src = (const ull*)dev_table;
    X.lo = c->bias[c->J] * (ull)c->tstride; X.n = c->nwords[c->J] * (ull)c->tstride;
    X.world = c->world; X.me = c->rank; X.epoch = ++c->xchg_epoch;      // every rank calls the exchange the same number of times
    X.info = c->fast ? c->info_buf.as<BuildInfo>() : nullptr;
    for (int r = 0; r < c->world; r++) { X.xtable[r] = c->peer_xtable[r]; X.ctrl[r] = c->peer_slctrl[r]; }
    k_xchg_push<<<c->world, 256, 0, c->stream>>>(X); LAUNCHED();
    k_xchg_wait<<<1, MAX_WORLD, 0, c->stream>>>((SliceCtrl*)c->sl_ctrl.p, c->world, c->xchg_epoch, c->fast ? c->info_buf.as<BuildInfo>() : nullptr); LAUNCHED();
    c->exchanged_by_peer_memory = true;
    c->uses_peer_exchange = true;
    CK(cudaMemcpyAsync(dev_table, c->sl_xtable.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return SVO_OK;
}

int svo_shard_emit(svo_ctx* c, const uint64_t* dev_table, uint64_t* n_voxels, uint64_t* n_nodes, uint64_t* n_data) {
    if (!c) return SVO_E_INVALID;
    if (!c->phase_a_done) return fail(c, SVO_E_INVALID, "svo_shard_emit before svo_shard_count");
    if (!dev_table) dev_table = c->table_own.as<uint64_t>();
    if (!dev_table) return fail(c, SVO_E_INVALID, "dev_table is NULL");
    CK(cudaSetDevice(c->device));
    int rc = c->fast ? fast_build(c, (ull*)dev_table, false) : build_phase_b(c, (const ull*)dev_table);
    if (rc) return rc;
    if (n_voxels) *n_voxels = c->n_voxels;
    if (n_nodes) *n_nodes = c->n_nodes;
    if (n_data) *n_data = c->n_data;
    return SVO_OK;
}

int svo_shard_ranges(svo_ctx* c, uint64_t* node_lo, uint64_t* node_hi, uint64_t* data_lo, uint64_t* data_hi) {
    if (!c) return SVO_E_INVALID;
    if (!c->built) return fail(c, SVO_E_INVALID, "svo_shard_ranges before the build finished");
    if (node_lo) *node_lo = c->node_lo;
    if (node_hi) *node_hi = c->node_hi;
    if (data_lo) *data_lo = c->data_lo;
    if (data_hi) *data_hi = c->data_hi;
    return SVO_OK;
}

// ---------------------------------------------------------------------------
// Triangle dispatch over peer memory (svo_dispatch.cuh)
// ---------------------------------------------------------------------------
int svo_ipc_export(const void* dev_ptr, void* handle64) {
    if (!dev_ptr || !handle64) return SVO_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)) != cudaSuccess) { (void)cudaGetLastError(); return SVO_E_CUDA; }
    memcpy(handle64, &h, 64);
    return SVO_OK;
}
int svo_ipc_open(const void* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return SVO_E_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    if (cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); *dev_ptr = nullptr; return SVO_E_CUDA; }
    return SVO_OK;
}
int svo_ipc_close(void* dev_ptr) {
    if (!dev_ptr) return SVO_OK;
    if (cudaIpcCloseMemHandle(dev_ptr) != cudaSuccess) { (void)cudaGetLastError(); return SVO_E_CUDA; }
    return SVO_OK;
}

int svo_shard_dispatch_create(svo_ctx* c, uint64_t capacity_tris, int fpt, void** dev_inbox, void** dev_ctrl) {
    if (!c) return SVO_E_INVALID;
    if (fpt != 9 && fpt != 21) return fail(c, SVO_E_INVALID, "floats_per_tri must be 9 (binary) or 21 (payload)");
    if (capacity_tris > 0xffffffffULL) return fail(c, SVO_E_INVALID, "more than 2^32-1 triangles");
    if (c->world > MAX_WORLD) return fail(c, SVO_E_INVALID, "triangle dispatch supports at most 16 ranks");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->attached = false; c->dispatched = false; c->dispatch_phase = 0;
    // exact-size cudaMalloc allocations of their own: these are the buffers peers map
    c->inbox.release(); c->ctrl_buf.release();
    CK(cudaMalloc(&c->inbox.p, (size_t)(capacity_tris ? capacity_tris : 1) * fpt * sizeof(float) + 256));
    c->inbox.cap = (size_t)(capacity_tris ? capacity_tris : 1) * fpt * sizeof(float) + 256;
    CK(cudaMalloc(&c->ctrl_buf.p, sizeof(DispatchCtrl)));
    c->ctrl_buf.cap = sizeof(DispatchCtrl);
    CK(cudaMemset(c->ctrl_buf.p, 0, sizeof(DispatchCtrl)));
    if (!c->h_ctrl) CK(cudaHostAlloc((void**)&c->h_ctrl, sizeof(DispatchCtrl), cudaHostAllocDefault));
    c->inbox_cap = capacity_tris;
    c->inbox_fpt = fpt;
    c->dispatch_epoch = 0;
    if (dev_inbox) *dev_inbox = c->inbox.p;
    if (dev_ctrl) *dev_ctrl = c->ctrl_buf.p;
    return SVO_OK;
}

int svo_shard_dispatch_attach(svo_ctx* c, void* const* inbox_ptrs, void* const* ctrl_ptrs) {
    if (!c) return SVO_E_INVALID;
    if (!c->inbox.p) return fail(c, SVO_E_INVALID, "svo_shard_dispatch_attach before svo_shard_dispatch_create");
    if (!inbox_ptrs || !ctrl_ptrs) return fail(c, SVO_E_INVALID, "peer pointer arrays are NULL");
    for (int r = 0; r < c->world; r++) {
        if (!inbox_ptrs[r] || !ctrl_ptrs[r]) return fail(c, SVO_E_INVALID, "a peer pointer is NULL");
        c->peer_inbox[r] = (float*)inbox_ptrs[r];
        c->peer_ctrl[r] = (DispatchCtrl*)ctrl_ptrs[r];
    }
    if (c->peer_inbox[c->rank] != c->inbox.p || (void*)c->peer_ctrl[c->rank] != c->ctrl_buf.p)
        return fail(c, SVO_E_INVALID, "entry [rank] of the peer arrays must be this context's own buffers");
    c->attached = true;
    return SVO_OK;
}

int svo_shard_dispatch_count(svo_ctx* c, const svo_params* params, const float* dev_local_tris, uint64_t n_local, int fpt) {
    if (!c) return SVO_E_INVALID;
    if (!c->attached) return fail(c, SVO_E_INVALID, "svo_shard_dispatch_count before svo_shard_dispatch_attach");
    int rc = validate_params(c, params);
    if (rc) return rc;
    if (fpt != c->inbox_fpt || (params->payload ? 21 : 9) != fpt) return fail(c, SVO_E_INVALID, "floats_per_tri does not match the inbox / params.payload");
    if (n_local && !dev_local_tris) return fail(c, SVO_E_INVALID, "dev_local_tris is NULL");
    if (((uintptr_t)dev_local_tris & 15) != 0) return fail(c, SVO_E_INVALID, "device triangle pointer must be 16-byte aligned");
    if (n_local > c->inbox_cap) return fail(c, SVO_E_RANGE, "local slice is larger than the inbox capacity");
    CK(cudaSetDevice(c->device));
    rc = derive_grid(c, params);
    if (rc) return rc;
    const int dc = shard_chunk_depth(c);
    if (c->D - dc < 2) return fail(c, SVO_E_INVALID, "gridsize too small for this many shards");
    c->partitioned = c->voxelized = c->built = false;
    c->have_tris = false; c->dispatched = false;
    for (int i = 0; i < EV_COUNT; i++) c->ev_set[i] = false;
    c->dispatch_launches = 0;
    const uint32_t launches_before = c->launches;
    mark(c, EV_DSP0);
    DispatchJob& D = c->dj;
    memset(&D, 0, sizeof D);
    D.tris = dev_local_tris; D.fpt = (uint32_t)fpt; D.n_local = n_local;
    D.world = c->world; D.me = c->rank;
    D.use_partitions = c->P > 1 ? 1 : 0; D.k = c->k;
    for (int i = 0; i < 32; i++) { D.bmin[i] = c->slab_min[i]; D.bmax[i] = c->slab_max[i]; }
    for (int r = 0; r < c->world; r++) {
        int lo[3], hi[3];
        shard_box(c, r, dc, lo, hi);
        for (int a = 0; a < 3; a++) {
            D.lo[r][a] = c->P > 1 ? lo[a] / (int)c->side : lo[a];
            D.hi[r][a] = c->P > 1 ? hi[a] / (int)c->side : hi[a];
            if (c->P > 1) { D.lof[r][a] = c->slab_min[D.lo[r][a]]; D.hif[r][a] = c->slab_max[D.hi[r][a]]; }
        }
        D.inbox[r] = c->peer_inbox[r];
        D.ctrl[r] = c->peer_ctrl[r];
    }
    D.unit_div = c->unit_div; D.gmax = (int)c->prm.gridsize - 1;
    D.nb = (n_local + VOX_BLOCK - 1) / VOX_BLOCK;
    D.epoch = ++c->dispatch_epoch;
    const ull ncnt = (ull)c->world * D.nb;
    CK(c->blockcnt.ensure((size_t)(ncnt + 1) * sizeof(unsigned)));
    CK(c->blockoff.ensure((size_t)(ncnt + 2) * sizeof(ull)));
    D.blockcnt = c->blockcnt.as<unsigned>();
    D.blockoff = c->blockoff.as<ull>();
    if (D.nb) {
        k_dispatch_count<<<(unsigned)D.nb, VOX_BLOCK, 0, c->stream>>>(D); LAUNCHED();
        U32Op op{ D.blockcnt };
        rc = exscan(c, op, ncnt, c->blockoff.as<ull>());
        if (rc) return rc;
    }
    k_dispatch_post<<<1, MAX_WORLD, 0, c->stream>>>(D, 0); LAUNCHED();
    c->dispatch_launches += c->launches - launches_before;
    c->dispatch_phase = 1;
    return SVO_OK;
}

int svo_shard_dispatch_send(svo_ctx* c) {
    if (!c) return SVO_E_INVALID;
    if (c->dispatch_phase != 1) return fail(c, SVO_E_INVALID, "svo_shard_dispatch_send before svo_shard_dispatch_count");
    CK(cudaSetDevice(c->device));
    const uint32_t launches_before = c->launches;
    DispatchJob& D = c->dj;
    k_dispatch_wait<<<1, MAX_WORLD, 0, c->stream>>>((DispatchCtrl*)c->ctrl_buf.p, c->world, 0, D.epoch); LAUNCHED();
    if (D.nb) {
        const size_t smem = 2 * (size_t)VOX_BLOCK * D.fpt * sizeof(float);
        k_dispatch_write<<<(unsigned)D.nb, VOX_BLOCK, smem, c->stream>>>(D, c->inbox_cap); LAUNCHED();
    }
    k_dispatch_post<<<1, MAX_WORLD, 0, c->stream>>>(D, 1); LAUNCHED();
    c->dispatch_launches += c->launches - launches_before;
    c->dispatch_phase = 2;
    return SVO_OK;
}

int svo_shard_dispatch_finish(svo_ctx* c, uint64_t* n_received) {
    if (!c) return SVO_E_INVALID;
    if (c->dispatch_phase != 2) return fail(c, SVO_E_INVALID, "svo_shard_dispatch_finish before svo_shard_dispatch_send");
    CK(cudaSetDevice(c->device));
    const uint32_t launches_before = c->launches;
    k_dispatch_wait<<<1, MAX_WORLD, 0, c->stream>>>((DispatchCtrl*)c->ctrl_buf.p, c->world, 1, c->dj.epoch); LAUNCHED();
    c->dispatch_launches += c->launches - launches_before;
    mark(c, EV_DSP1);
    CK(cudaMemcpyAsync(c->h_ctrl, c->ctrl_buf.p, sizeof(DispatchCtrl), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->dispatch_phase = 0;
    if (c->h_ctrl->error) {
        const ull e = c->h_ctrl->error;
        CK(cudaMemsetAsync(&((DispatchCtrl*)c->ctrl_buf.p)->error, 0, sizeof(ull), c->stream));
        return fail(c, e == 3 ? SVO_E_RANGE : SVO_E_CUDA,
                    e == 3 ? "triangle dispatch: a peer inbox is too small for the triangles routed to it"
                           : "triangle dispatch: timed out waiting for a peer rank");
    }
    // every rank holds the whole count matrix, so all of them see an overflowing column and fail together
    ull n_in = 0;
    for (int d = 0; d < c->world; d++) {
        ull col = 0;
        for (int s = 0; s < c->world; s++) col += c->h_ctrl->matrix[s][d];
        if (col > c->inbox_cap) return fail(c, SVO_E_RANGE, "triangle dispatch: a peer inbox is too small for the triangles routed to it");
        if (d == c->rank) n_in = col;
    }
    int rc = set_tris_common(c, n_in, c->inbox_fpt);
    if (rc) return rc;
    c->d_tris = c->inbox.as<float>();
    c->dispatched = true;
    if (n_received) *n_received = n_in;
    return SVO_OK;
}

// ---------------------------------------------------------------------------
// Remote staging of triangle slices (svo_dispatch.cuh)
// ---------------------------------------------------------------------------
// Layout of the window every rank allocates (identical offsets on all ranks: same capacity, fpt, world)
struct WindowLayout { size_t ctrl, xtable, list, slice, total; };
static WindowLayout window_layout(uint64_t cap_blocks, int fpt, int world) {
    WindowLayout L;
    auto up = [](size_t v) { return (v + 4095) & ~(size_t)4095; };
    L.ctrl = 0;
    L.xtable = up(sizeof(SliceCtrl));
    L.list = L.xtable + XTABLE_BYTES;
    L.slice = L.list + up((size_t)world * cap_blocks * 4 * sizeof(uint32_t));      // one entry per 32-triangle unit
    L.total = L.slice + up((size_t)cap_blocks * VOX_BLOCK * fpt * sizeof(float));
    return L;
}

int svo_shard_slice_create(svo_ctx* c, uint64_t capacity_tris, int fpt, void** dev_window) {
    if (!c) return SVO_E_INVALID;
    if (fpt != 9 && fpt != 21) return fail(c, SVO_E_INVALID, "floats_per_tri must be 9 (binary) or 21 (payload)");
    if (c->world > MAX_WORLD) return fail(c, SVO_E_INVALID, "remote triangle slices support at most 16 ranks");
    if (capacity_tris * (uint64_t)c->world > 0xffffffffULL) return fail(c, SVO_E_INVALID, "more than 2^32-1 triangles");
    if (capacity_tris >= (1ULL << 28) * VOX_BLOCK) return fail(c, SVO_E_INVALID, "slice capacity above 2^35 triangles");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->sl_attached = false; c->sliced = false;
    c->window.release();
    const uint64_t cap = capacity_tris ? capacity_tris : 1;
    c->sl_cap_blocks = (cap + VOX_BLOCK - 1) / VOX_BLOCK;
    const WindowLayout L = window_layout(c->sl_cap_blocks, fpt, c->world);
    // an exact cudaMalloc allocation of its own: this is the buffer peers map
    CK(cudaMalloc(&c->window.p, L.total)); c->window.cap = L.total;
    char* w = (char*)c->window.p;
    c->sl_ctrl.p = w + L.ctrl; c->sl_xtable.p = w + L.xtable; c->sl_list.p = w + L.list; c->slice.p = w + L.slice;
    CK(cudaMemset(c->sl_ctrl.p, 0, sizeof(SliceCtrl)));
    CK(c->sl_cursor.ensure((MAX_WORLD + 1) * sizeof(ull)));
    CK(cudaMemset(c->sl_cursor.p, 0, (MAX_WORLD + 1) * sizeof(ull)));
    c->slice_cap = capacity_tris;
    c->sl_world = c->world;
    c->slice_fpt = fpt;
    c->slice_n_local = 0;
    c->sl_epoch = 0; c->xchg_epoch = 0; c->uses_peer_exchange = false;
    if (dev_window) *dev_window = c->window.p;
    return SVO_OK;
}

int svo_shard_slice_attach(svo_ctx* c, void* const* windows) {
    if (!c) return SVO_E_INVALID;
    if (!c->window.p) return fail(c, SVO_E_INVALID, "svo_shard_slice_attach before svo_shard_slice_create");
    if (!windows) return fail(c, SVO_E_INVALID, "peer window array is NULL");
    if (c->world != c->sl_world) return fail(c, SVO_E_INVALID, "svo_shard_configure changed the world size after svo_shard_slice_create: create the window again");
    CK(cudaSetDevice(c->device));
    const WindowLayout L = window_layout(c->sl_cap_blocks, c->slice_fpt, c->world);
    for (int r = 0; r < c->world; r++) {
        if (!windows[r]) return fail(c, SVO_E_INVALID, "a peer window pointer is NULL");
        // a raw pointer of another device of THIS process (several contexts in one process): map it. Pointers that came
        // through svo_ipc_open are mapped already.
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, windows[r]) == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != c->device) {
            const cudaError_t pe = cudaDeviceEnablePeerAccess(pa.device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return fail(c, SVO_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)); }
        }
        (void)cudaGetLastError();
        char* w = (char*)windows[r];
        c->peer_slctrl[r] = (SliceCtrl*)(w + L.ctrl);
        c->peer_xtable[r] = (ull*)(w + L.xtable);
        c->peer_list[r] = (uint32_t*)(w + L.list);
        c->peer_slice[r] = (float*)(w + L.slice);
    }
    if (windows[c->rank] != c->window.p) return fail(c, SVO_E_INVALID, "entry [rank] of the window array must be this context's own window");
    c->sl_attached = true;
    return SVO_OK;
}

static int slice_fence(svo_ctx* c) {
    if (!c) return SVO_E_INVALID;
    if (!c->sl_attached) return fail(c, SVO_E_INVALID, "svo_shard_slice_fence before svo_shard_slice_attach");
    CK(cudaSetDevice(c->device));
    if (c->sl_epoch) { k_slice_wait<<<1, MAX_WORLD, 0, c->stream>>>((SliceCtrl*)c->sl_ctrl.p, c->world, 1, c->sl_epoch); LAUNCHED(); }
    return SVO_OK;
}

// Public form: the caller is about to write into the slice itself, so the cached unit boxes go with the old contents.
int svo_shard_slice_fence(svo_ctx* c) {
    if (!c) return SVO_E_INVALID;
    c->sl_boxes_valid = false;
    return slice_fence(c);
}

int svo_shard_slice_upload(svo_ctx* c, const float* src, uint64_t n_local) {
    if (!c) return SVO_E_INVALID;
    if (!c->sl_attached) return fail(c, SVO_E_INVALID, "svo_shard_slice_upload before svo_shard_slice_attach");
    if (n_local > c->slice_cap) return fail(c, SVO_E_RANGE, "slice is larger than the capacity given to svo_shard_slice_create");
    if (n_local && !src) return fail(c, SVO_E_INVALID, "src is NULL");
    int rc = slice_fence(c);                      // peers may still be reading the previous contents
    if (rc) return rc;
    for (int i = 0; i < EV_COUNT; i++) c->ev_set[i] = false;
    mark(c, EV_UP0);
    if (n_local) CK(cudaMemcpyAsync(c->slice.p, src, (size_t)n_local * c->slice_fpt * sizeof(float), cudaMemcpyDefault, c->stream));
    mark(c, EV_UP1);
    c->slice_n_local = n_local;
    c->sl_boxes_valid = false;
    return SVO_OK;
}

int svo_shard_slice_begin(svo_ctx* c, uint64_t n_local) {
    if (!c) return SVO_E_INVALID;
    if (!c->sl_attached) return fail(c, SVO_E_INVALID, "svo_shard_slice_begin before svo_shard_slice_attach");
    if (n_local > c->slice_cap) return fail(c, SVO_E_RANGE, "slice is larger than the capacity given to svo_shard_slice_create");
    int rc = slice_fence(c);                      // peers may still be reading the previous contents
    if (rc) return rc;
    for (int i = 0; i < EV_COUNT; i++) c->ev_set[i] = false;
    mark(c, EV_UP0);
    if (!c->up_ev[0]) { CK(cudaEventCreateWithFlags(&c->up_ev[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->up_ev[1], cudaEventDisableTiming)); }
    c->up_slot = 0; c->up_pending[0] = c->up_pending[1] = false;
    c->slice_n_local = n_local;
    c->stream_fill = 0;
    c->sl_boxes_valid = false;                              // the slice is about to change
    if (n_local == 0) mark(c, EV_UP1);
    return SVO_OK;
}

int svo_shard_slice_append(svo_ctx* c, const float* host_chunk, uint64_t n) {
    if (!c) return SVO_E_INVALID;
    if (!c->sl_attached) return fail(c, SVO_E_INVALID, "svo_shard_slice_append before svo_shard_slice_attach");
    if (c->stream_fill + n > c->slice_n_local) return fail(c, SVO_E_RANGE, "more triangles appended than announced");
    if (n && !host_chunk) return fail(c, SVO_E_INVALID, "host_chunk is NULL");
    CK(cudaSetDevice(c->device));
    const size_t rec = (size_t)c->slice_fpt * sizeof(float);
    if (n) CK(cudaMemcpyAsync((char*)c->slice.p + c->stream_fill * rec, host_chunk, n * rec, cudaMemcpyHostToDevice, c->stream));
    // same double-buffer contract as svo_triangles_append: on return every EARLIER chunk has been copied
    const int slot = c->up_slot;
    CK(cudaEventRecord(c->up_ev[slot], c->stream));
    c->up_pending[slot] = true;
    if (c->up_pending[slot ^ 1]) { CK(cudaEventSynchronize(c->up_ev[slot ^ 1])); c->up_pending[slot ^ 1] = false; }
    c->up_slot = slot ^ 1;
    c->stream_fill += n;
    c->sl_boxes_valid = false;
    if (c->stream_fill == c->slice_n_local) mark(c, EV_UP1);
    return SVO_OK;
}

int svo_shard_slice_publish(svo_ctx* c, const svo_params* params, uint64_t n_total) {
    if (!c) return SVO_E_INVALID;
    if (!c->sl_attached) return fail(c, SVO_E_INVALID, "svo_shard_slice_publish before svo_shard_slice_attach");
    if (c->world != c->sl_world) return fail(c, SVO_E_INVALID, "svo_shard_configure changed the world size after svo_shard_slice_create: create the window again");
    int rc = validate_params(c, params);
    if (rc) return rc;
    if ((params->payload ? 21 : 9) != c->slice_fpt) return fail(c, SVO_E_INVALID, "params.payload does not match the slices' floats_per_tri");
    if (n_total > 0xffffffffULL) return fail(c, SVO_E_INVALID, "more than 2^32-1 triangles");
    CK(cudaSetDevice(c->device));
    rc = derive_grid(c, params);
    if (rc) return rc;
    const int dc = shard_chunk_depth(c);
    if (c->D - dc < 2) return fail(c, SVO_E_INVALID, "gridsize too small for this many shards");
    rc = slice_fence(c);                                    // the peers' list buffers are about to be rewritten (one-block wait kernel)
    if (rc) return rc;
    const uint32_t launches_before = c->launches;
    mark(c, EV_DSP0);
    SliceJob& S = c->sj;
    memset(&S, 0, sizeof S);
    DispatchJob& D = S.D;
    D.tris = c->slice.as<float>(); D.fpt = (uint32_t)c->slice_fpt; D.n_local = c->slice_n_local;
    D.world = c->world; D.me = c->rank;
    D.use_partitions = c->P > 1 ? 1 : 0; D.k = c->k;
    for (int i = 0; i < 32; i++) { D.bmin[i] = c->slab_min[i]; D.bmax[i] = c->slab_max[i]; }
    for (int r = 0; r < c->world; r++) {
        int lo[3], hi[3];
        shard_box(c, r, dc, lo, hi);
        for (int a = 0; a < 3; a++) {
            D.lo[r][a] = c->P > 1 ? lo[a] / (int)c->side : lo[a];
            D.hi[r][a] = c->P > 1 ? hi[a] / (int)c->side : hi[a];
            if (c->P > 1) { D.lof[r][a] = c->slab_min[D.lo[r][a]]; D.hif[r][a] = c->slab_max[D.hi[r][a]]; }
        }
        S.list[r] = c->peer_list[r];
        S.ctrl[r] = c->peer_slctrl[r];
    }
    D.unit_div = c->unit_div; D.gmax = (int)c->prm.gridsize - 1;
    D.nb = (D.n_local + VOX_BLOCK - 1) / VOX_BLOCK;
    D.epoch = ++c->sl_epoch;
    S.cap = c->sl_cap_blocks * 4;
    S.cursor = c->sl_cursor.as<ull>();
    if (D.n_local) {
        const ull n_units = (D.n_local + UNIT - 1) / UNIT;
        const unsigned grid = (unsigned)std::min<ull>((n_units + FILTER_WARPS * 4 - 1) / (FILTER_WARPS * 4), (ull)c->sm_count * 8);
        // The bounding boxes of the slice's 32-triangle units do not depend on the job: they are computed when the slice has
        // changed (svo_shard_slice_begin / _append) and reused by every job on the same slice, whose filter then reads 32
        // bytes per unit instead of the unit's 1152 / 2688. (SVO_SLICE_BOXES=0: test the triangles every time.)
        static const bool use_boxes = !(getenv("SVO_SLICE_BOXES") && getenv("SVO_SLICE_BOXES")[0] == '0');
        if (use_boxes) {
            CK(c->sl_ubox.ensure((size_t)c->sl_cap_blocks * 4 * 2 * sizeof(float4)));
            S.ubox = c->sl_ubox.as<float4>();
            if (!c->sl_boxes_valid) {
                S.ubox_mode = 1;
                k_slice_boxes<<<grid, FILTER_WARPS * 32, 0, c->stream>>>(S); LAUNCHED();
                c->sl_boxes_valid = true;
            }
            S.ubox_mode = 2;
        }
        // one launch: list this slice's units per destination and (the block that finishes last) publish counts + flag
        k_slice_filter<<<grid, FILTER_WARPS * 32, 0, c->stream>>>(S); LAUNCHED();
    } else {
        k_slice_post<<<1, MAX_WORLD, 0, c->stream>>>(S, 0); LAUNCHED();
    }
    mark(c, EV_DSP1);
    c->dispatch_launches = c->launches - launches_before;
    // the triangle set of this context is now the union of all slices, addressed by file position
    c->n_tris = n_total;
    c->fpt = c->slice_fpt;
    c->d_tris = c->slice.as<float>();
    c->have_tris = true;
    c->dispatched = false;
    c->sliced = true;
    c->partitioned = c->voxelized = c->built = false;
    return SVO_OK;
}

int svo_shard_layout_from_table(const svo_params* params, int rank, int world, const uint64_t* host_table, uint64_t n_u64,
                                svo_shard_layout* out, uint64_t* rec_pos, uint64_t* rec_words, uint64_t rec_capacity) {
    if (!params || !host_table || !out) return fail(nullptr, SVO_E_INVALID, "svo_shard_layout_from_table: NULL argument");
    if (world < 1 || rank < 0 || rank >= world || (world & (world - 1)) != 0)
        return fail(nullptr, SVO_E_INVALID, "shard world size must be a power of two and 0 <= rank < world");
    svo_ctx tmp;                                   // geometry only: no CUDA call is made on it
    svo_ctx* c = &tmp;
    memset(c->nwords, 0, sizeof c->nwords);
    memset(c->bias, 0, sizeof c->bias);
    int rc = validate_params(c, params);
    if (rc == SVO_OK) { c->world = world; c->rank = rank; rc = derive_grid(c, params); }
    if (rc == SVO_OK) rc = setup_geometry(c);
    if (rc != SVO_OK) return fail(nullptr, rc, tmp.err);
    if (n_u64 != c->WJ * c->tstride) return fail(nullptr, SVO_E_RANGE, "table size does not match svo_shard_table_size for this geometry");
    shard_merge_compute(c, (const ull*)host_table);
    out->n_voxels = c->n_voxels; out->n_nodes = c->n_nodes;
    out->node_lo = c->node_lo; out->node_hi = c->node_hi;
    out->leaf_offset = c->leaf_offset; out->n_voxels_local = c->n_voxels_local;
    out->n_upper_records = c->n_upper_records;
    if (rec_pos && rec_words) {
        if (rec_capacity < c->n_upper_records) return fail(nullptr, SVO_E_RANGE, "rec_capacity is smaller than n_upper_records");
        for (ull i = 0; i < c->n_upper_records; i++) {
            rec_pos[i] = c->h_rpos[i];
            for (int q = 0; q < 3; q++) rec_words[3 * i + q] = c->h_rrec[3 * i + q];
        }
    }
    return SVO_OK;
}

static int fetch_common(svo_ctx* c, const DevBuf& src, uint64_t lo, uint64_t hi, uint64_t rec, uint64_t first, uint64_t count, void* dst) {
    if (!c->built) return fail(c, SVO_E_INVALID, "fetch before svo_build");
    if (first < lo || first > hi || count > hi - first) return fail(c, SVO_E_RANGE, "record range outside this context's part of the file");
    first -= lo;
    if (count && !dst) return fail(c, SVO_E_INVALID, "dst is NULL");
    CK(cudaSetDevice(c->device));
    mark(c, EV_DN0);
    if (count) CK(cudaMemcpyAsync(dst, (const char*)src.p + first * rec, count * rec, cudaMemcpyDefault, c->stream));    // dst: host, or device memory (UVA)
    mark(c, EV_DN1);
    CK(cudaStreamSynchronize(c->stream));
    c->stats.ms_download += span(c, EV_DN0, EV_DN1);
    return SVO_OK;
}

int svo_fetch_nodes(svo_ctx* c, uint64_t first, uint64_t count, void* dst) {
    if (!c) return SVO_E_INVALID;
    return fetch_common(c, c->nodes, c->node_lo, c->node_hi, SVO_NODE_BYTES, first, count, dst);
}
int svo_fetch_data(svo_ctx* c, uint64_t first, uint64_t count, void* dst) {
    if (!c) return SVO_E_INVALID;
    return fetch_common(c, c->data, c->data_lo, c->data_hi, SVO_DATA_BYTES, first, count, dst);
}

int svo_device_nodes(svo_ctx* c, const void** p, uint64_t* n) {
    if (!c) return SVO_E_INVALID;
    if (!c->built) return fail(c, SVO_E_INVALID, "svo_device_nodes before svo_build");
    if (p) *p = c->nodes.p;
    if (n) *n = c->node_hi - c->node_lo;
    return SVO_OK;
}
int svo_device_data(svo_ctx* c, const void** p, uint64_t* n) {
    if (!c) return SVO_E_INVALID;
    if (!c->built) return fail(c, SVO_E_INVALID, "svo_device_data before svo_build");
    if (p) *p = c->data.p;
    if (n) *n = c->data_hi - c->data_lo;
    return SVO_OK;
}

int svo_fetch_voxel_codes(svo_ctx* c, uint64_t* dst, uint64_t capacity, uint64_t* n_written) {
    if (!c) return SVO_E_INVALID;
    if (!c->built) return fail(c, SVO_E_INVALID, "svo_fetch_voxel_codes before svo_build");
    CK(cudaSetDevice(c->device));
    const uint64_t nloc = c->world > 1 ? c->n_voxels_local : c->n_voxels;
    const uint64_t n = nloc < capacity ? nloc : capacity;
    if (n_written) *n_written = n;
    if (n == 0) return SVO_OK;
    if (!dst) return fail(c, SVO_E_INVALID, "dst is NULL");
    CK(c->codes.ensure((size_t)n * sizeof(ull)));
    k_voxel_codes<<<blocks_for(c->lv[0].n, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, 0, c->stream>>>(c->lv[0].view(), c->codes.as<ull>(), n); LAUNCHED();
    CK(cudaMemcpyAsync(dst, c->codes.p, (size_t)n * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

int svo_partition_voxel_counts(svo_ctx* c, uint64_t* counts, uint64_t capacity) {
    if (!c) return SVO_E_INVALID;
    if (!c->built) return fail(c, SVO_E_INVALID, "svo_partition_voxel_counts before svo_build");
    if (!counts || capacity < c->P) return fail(c, SVO_E_RANGE, "counts capacity is smaller than the partition count");
    CK(cudaSetDevice(c->device));
    CK(c->part_counts.ensure(c->P * sizeof(ull)));
    CK(cudaMemsetAsync(c->part_counts.p, 0, c->P * sizeof(ull), c->stream));
    if (c->lv[0].n) {
        // a level-0 word covers 6 Morton bits, a partition 3 * (D - k): partition of a brick = key >> (3 (D - k) - 6)
        const int sh = 3 * (c->D - c->k) - 6;
        k_partition_voxels<<<blocks_for(c->lv[0].n, 256), 256, 0, c->stream>>>(c->lv[0].key.as<ull>(), c->lv[0].mask.as<ull>(), c->lv[0].n, sh, c->part_counts.as<ull>()); LAUNCHED();
    }
    CK(cudaMemcpyAsync(counts, c->part_counts.p, c->P * sizeof(ull), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

int svo_get_stats(svo_ctx* c, svo_stats* s) {
    if (!c || !s) return SVO_E_INVALID;
    *s = c->stats;
    return SVO_OK;
}

int svo_run(svo_ctx* c, const svo_params* params, const float* host_tris, uint64_t n_tris,
            void* nodes_dst, uint64_t nodes_cap, void* data_dst, uint64_t data_cap, svo_stats* stats) {
    if (!c) return SVO_E_INVALID;
    if (!params) return fail(c, SVO_E_INVALID, "params is NULL");
    int rc = svo_set_triangles(c, host_tris, n_tris, params->payload ? 21 : 9);
    if (rc) return rc;
    rc = svo_partition(c, params, nullptr, nullptr, 0);
    if (rc) return rc;
    rc = svo_voxelize(c);
    if (rc) return rc;
    uint64_t nv, nn, nd;
    rc = svo_build(c, &nv, &nn, &nd);
    if (rc) return rc;
    int short_rc = SVO_OK;
    if (nodes_dst) {
        if (nodes_cap < nn) short_rc = fail(c, SVO_E_RANGE, "nodes_dst too small");
        else if ((rc = svo_fetch_nodes(c, 0, nn, nodes_dst))) return rc;
    }
    if (data_dst) {
        if (data_cap < nd) short_rc = fail(c, SVO_E_RANGE, "data_dst too small");
        else if ((rc = svo_fetch_data(c, 0, nd, data_dst))) return rc;
    }
    if (stats) *stats = c->stats;
    return short_rc;
}

}  // extern "C"
