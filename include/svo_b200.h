/* svo_b200.h -- C ABI of the B200-native voxelize-and-build path.
 *
 * Drop-in boundary for Forceflow/ooc_svo_builder's hot path. The reference has
 * no plugin / FFI interface; its in-process seams are the four
 * calls main() makes (src/svo_builder/main.cpp:298-389). Each entry point below
 * replaces one of them and cites it. All reference paths are relative to
 * /root/reference/.
 *
 * Conventions: extern "C", plain pointers and sizes, int status (0 = ok, see
 * SVO_E_*), svo_last_error() for the message. The library owns device memory
 * and streams; callers own every host buffer they pass. A context is bound to
 * one CUDA device and is not thread-safe; use one context per device. There is
 * NO CPU fallback: every compute entry point fails with SVO_E_CUDA when no
 * sm_100 device is usable.
 */
#ifndef SVO_B200_H_
#define SVO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_OK            0
#define SVO_E_INVALID     1   /* bad argument / call order                         */
#define SVO_E_CUDA        2   /* CUDA runtime error or no usable device            */
#define SVO_E_NOMEM       3   /* device or host allocation failed                  */
#define SVO_E_RANGE       4   /* destination buffer too small / range out of bounds*/
#define SVO_E_RETRY       5   /* sharded build only: repeat svo_shard_count / _exchange / _emit (see svo_shard_emit)   */

/* main.cpp:23 `enum ColorType`, selected by `-c` (main.cpp:152-178). */
#define SVO_COLOR_MODEL   0
#define SVO_COLOR_FIXED   1
#define SVO_COLOR_LINEAR  2
#define SVO_COLOR_NORMAL  3

#define SVO_NODE_BYTES    24  /* octree_io.h:62-66: u64 data, u64 children_base, i8 children_offset[8] */
#define SVO_DATA_BYTES    32  /* VoxelData.h:10-17: u64 morton, f32 color[3], f32 normal[3]            */

typedef struct svo_ctx svo_ctx;

/* The program parameters of main.cpp:28-36 that reach the hot path. */
typedef struct svo_params {
    uint64_t gridsize;          /* -s, power of two >= 2                       (main.cpp:30)  */
    uint64_t memory_limit_mb;   /* -l, decides the logical partition count     (main.cpp:31)  */
    float    bbox_min0;         /* .tri header bbox.min[0]                     (tri_tools.h:109) */
    float    bbox_max0;         /* .tri header bbox.max[0]; only the x extent is used (main.cpp:311) */
    int32_t  payload;           /* 0 = svo_builder_binary (9 floats / triangle),
                                   1 = svo_builder        (21 floats / triangle)  (tri_util.h:7-11) */
    int32_t  generate_levels;   /* -levels                                     (main.cpp:35)  */
    int32_t  color_mode;        /* -c, SVO_COLOR_*                             (main.cpp:33)  */
    float    sparseness_limit;  /* -d as a fraction; accepted for CLI compatibility. It only
                                   switches the reference between two routes that produce the
                                   same files (voxelizer.cpp:176-186, main.cpp:355-368).       */
    int32_t  separability;      /* 0 or 26: the reference's conservative (26-separating) Schwarz-Seidel test
                                   (voxelizer.cpp:138-307). 6: OPT-IN 6-separating ("thin") variant of the same
                                   paper -- plane test against the voxel's centre segment along the dominant normal
                                   axis + the one projection orthogonal to it at the voxel centre. The reference has
                                   no such mode (its only test is the conservative one); the definition is restated in
                                   the C file of the test oracle (oracle/), which is what this mode is tested against. */
} svo_params;

/* Counters and device-side stage timings (CUDA events) of the last run. */
typedef struct svo_stats {
    uint64_t n_partitions;      /* logical partitions P = 8^k                                  */
    uint64_t n_pairs;           /* sum of per-partition triangle counts                        */
    uint64_t n_voxels;          /* "Total amount of voxels" (main.cpp:391)                     */
    uint64_t n_nodes;           /* .octree n_nodes                                             */
    uint64_t n_data;            /* .octree n_data                                              */
    uint64_t n_small, n_medium, n_large; /* triangle/partition pairs per bbox work class       */
    float ms_upload;            /* host -> device triangle copy                                */
    float ms_partition;         /* binning pass                                                */
    float ms_voxelize;          /* Schwarz-Seidel kernels (all classes) + payload owner pass   */
    float ms_build;             /* pyramid compaction + subtree sizes + node/data emission     */
    float ms_emit;              /* the node emission kernels alone (subset of ms_build)        */
    float ms_clear;             /* sparse clear of the bit-grid pyramid for the next run       */
    float ms_download;          /* device -> host copies issued by svo_fetch_*                 */
    float ms_vox_small;         /* k_vox_small alone (subset of ms_voxelize)                   */
    float ms_emit_leaf;         /* k_emit_leaf alone (subset of ms_emit)                       */
    float ms_compact;           /* level counts + top-down tile-list expansion + subtree sizes */
    float ms_dispatch;          /* multi-GPU: slice block-list pass (remote staging) or triangle dispatch */
    float ms_peer_wait;         /* multi-GPU, remote staging: device-side wait for the peers' block lists  */
    uint32_t kernel_launches;   /* kernels launched by the last run                            */
    uint32_t speculative;       /* 1: the build ran without a host read-back before the final one (capacities of the previous build) */
    uint64_t n_bricks;          /* occupied 4x4x4 bricks (octree nodes at depth D-2) of this context's slab               */
    uint64_t n_tiles1;          /* occupied 16^3 tiles (non-zero words of pyramid level 1) of this context's slab         */
    uint64_t n_brick_records;   /* node records of this context's brick subtrees (leaves + depth D-1 nodes): what the
                                   brick-level emitter writes; 0 on the host-driven (classic) build path                  */
} svo_stats;

/* ---- lifetime ------------------------------------------------------------ */

/* Creates a context on CUDA device `device`. Fails (SVO_E_CUDA) if the device is
 * missing or is not compute capability 10.x. */
int  svo_ctx_create(int device, svo_ctx** out);
void svo_ctx_destroy(svo_ctx* ctx);
/* Makes ctx issue all its work on `cuda_stream` (a cudaStream_t of ctx's device,
 * e.g. the caller's framework stream) instead of its own stream; NULL restores
 * the context's own stream. The caller keeps the stream alive. */
int  svo_ctx_set_stream(svo_ctx* ctx, void* cuda_stream);
/* Message of the last failure on `ctx`; ctx may be NULL for svo_ctx_create failures. */
const char* svo_last_error(const svo_ctx* ctx);
const char* svo_version(void);

/* ---- partitioner --------------------------------------------------------- */

/* Replaces `size_t estimate_partitions(gridsize, memory_limit)`
 * (src/svo_builder/partitioner.cpp:12-28). Pure host arithmetic. */
uint64_t svo_estimate_partitions(uint64_t gridsize, uint64_t memory_limit_mb);

/* Replaces the .trip text round trip main.cpp applies to the bbox before it
 * computes the voxelizer's unit length (trip_tools.h:110-111 written, :78 read,
 * main.cpp:304-311): formats like `ostream << float`, parses like `istream >> float`. */
float svo_text_roundtrip_float(float v);

/* Stages the triangle records. Replaces TriReader's 8192-triangle fread loop
 * (src/libs/libtri/include/TriReader.h:41-79). `tris` is n_tris * (9 or 21)
 * packed little-endian float32 (tri_util.h:29-55).
 *   svo_set_triangles        : `tris` is HOST memory (pageable or pinned); copied to the device.
 *   svo_set_triangles_device : `tris` is DEVICE memory on ctx's device; borrowed, not copied,
 *                              and must stay valid until the next svo_set_triangles* call. */
int svo_set_triangles(svo_ctx* ctx, const float* tris, uint64_t n_tris, int floats_per_tri);
int svo_set_triangles_device(svo_ctx* ctx, const float* tris, uint64_t n_tris, int floats_per_tri);

/* Streamed variant of svo_set_triangles for files larger than the host budget (-l): the caller reads the
 * .tridata file in chunks (the role of TriReader's buffer, TriReader.h:9-40, with megabytes instead of 8192
 * triangles) into TWO alternating pinned buffers (svo_host_alloc) and appends them; the copy of one chunk
 * overlaps the fread of the next. When svo_triangles_append returns, every EARLIER chunk has reached the
 * device (its buffer may be refilled); the chunk just passed is still in flight. The triangle set is complete
 * when n_tris records have been appended. */
int svo_triangles_begin(svo_ctx* ctx, uint64_t n_tris, int floats_per_tri);
int svo_triangles_append(svo_ctx* ctx, const float* host_chunk, uint64_t n_chunk_tris);

/* Replaces `TripInfo partition(tri_info, n_partitions, gridsize)`
 * (partitioner.cpp:101-149, BBoxBuffer.h:70-84, intersection.h:50-53): bins every
 * triangle into every logical partition whose world box its bbox touches
 * (inclusive float test). Produces device index lists instead of .tripdata
 * files. `part_tricounts` (may be NULL) receives n_partitions counts -- the
 * values the reference writes to the .trip header (trip_tools.h:118-120). */
int svo_partition(svo_ctx* ctx, const svo_params* params,
                  uint64_t* n_partitions, uint64_t* part_tricounts, uint64_t tricounts_capacity);

/* ---- voxelizer ----------------------------------------------------------- */

/* Replaces the per-partition loop around `voxelize_schwarz_method(...)`
 * (main.cpp:329-352, voxelizer.cpp:138-307) for ALL partitions: conservative
 * Schwarz-Seidel triangle/box overlap into a Morton-ordered bit-grid. Requires
 * svo_partition. n_voxels may be NULL (the count is final after svo_build). */
int svo_voxelize(svo_ctx* ctx);

/* ---- octree builder ------------------------------------------------------ */

/* Replaces OctreeBuilder::addVoxel x N + finalizeTree (OctreeBuilder.cpp:34-168,
 * main.cpp:354-389): builds the node array (and the payload data array) on the
 * device in the reference's file order. Outputs the record counts that go into
 * the .octree header (octree_io.h:74-83). */
int svo_build(svo_ctx* ctx, uint64_t* n_voxels, uint64_t* n_nodes, uint64_t* n_data);

/* Replaces writeNode / writeVoxelData (octree_io.h:49-66): copies records
 * [first, first+count) of the .octreenodes / .octreedata image into caller
 * memory (count * 24 resp. count * 32 bytes). Chunk the calls to stream the
 * result within a host memory budget (-l). */
int svo_fetch_nodes(svo_ctx* ctx, uint64_t first, uint64_t count, void* dst);   /* dst: host memory, or device memory of ctx's device */
int svo_fetch_data(svo_ctx* ctx, uint64_t first, uint64_t count, void* dst);

/* Device-resident views of the same images (valid until the next svo_partition /
 * svo_build on ctx) for callers that keep the octree on the GPU. */
int svo_device_nodes(svo_ctx* ctx, const void** dev_ptr, uint64_t* n_nodes);
int svo_device_data(svo_ctx* ctx, const void** dev_ptr, uint64_t* n_data);

/* Ascending Morton codes of the filled voxels (the stream the reference feeds to
 * addVoxel, main.cpp:355-368). dst holds `capacity` uint64; *n_written <= capacity. */
int svo_fetch_voxel_codes(svo_ctx* ctx, uint64_t* dst, uint64_t capacity, uint64_t* n_written);

/* ---- multi-GPU: partitions sharded across contexts -------------------------
 *
 * One context per GPU (one process per GPU, or several contexts in one process).
 * Rank r of `world` owns a contiguous Morton range of the grid's 8^dc chunks
 * (dc >= log8 P), i.e. whole logical partitions when P >= world. It voxelizes and
 * builds only the subtrees of its range; the only exchange is a small table with
 * one 4 x u64 entry {mask, subtree size, leaves, internal nodes} per top-of-shard
 * subtree, which the CALLER sums across ranks (NCCL all-reduce over NVLink; the
 * entries of different ranks are disjoint). Every rank then merges the shared
 * upper levels itself and emits its own contiguous range of the output files.
 *
 *   svo_shard_configure(rank, world)  before svo_partition
 *   svo_partition, svo_voxelize       as on one GPU (every rank sees all triangles, or the ones the
 *                                     triangle dispatch below routed to it)
 *   svo_shard_table_size              u64 count of the table
 *   svo_shard_count(dev_table)        local build phase; zeroes the table, writes own entries
 *   <all-reduce(sum) dev_table>       the caller's collective
 *   svo_shard_emit(dev_table, ...)    merged upper levels + emission; returns GLOBAL counts
 *   svo_shard_ranges                  [node_lo, node_hi) / [data_lo, data_hi): the records this
 *                                     context holds; svo_fetch_* take global record positions
 *                                     inside these ranges. The ranges of all ranks tile the files.
 * -levels: the table entries grow to 8 x u64 (the tile's 6-float data cache rides along) and the shared upper levels
 * are averaged by every rank on the host from the table (the reference's float op order); builds are sized.
 *
 * SVO_E_RETRY. From its second job on a context builds speculatively: tile lists and node buffer keep the
 * capacities of the previous job and nothing waits for the host between svo_voxelize and the end of
 * svo_shard_emit. When the table goes through svo_shard_exchange (peer memory) and some rank's tile lists
 * turn out too small, that rank has no table entries to offer; it says so inside the exchange, and EVERY
 * rank's svo_shard_emit returns SVO_E_RETRY in that step. All ranks then repeat svo_shard_count ->
 * svo_shard_exchange -> svo_shard_emit (no new svo_voxelize: the voxelized grids are kept); the ranks that
 * overflowed take a sized build. Local builds are speculative only on contexts whose previous job used
 * svo_shard_exchange (such a context is expected to keep using it); with the caller's own collective builds
 * are always sized and SVO_E_RETRY never occurs. */
/* dev_table may be NULL in svo_shard_count / svo_shard_exchange / svo_shard_emit: the library then uses a table of its
 * own (callers without a device allocator, e.g. the CLI; only meaningful with svo_shard_exchange). */
int svo_shard_configure(svo_ctx* ctx, int rank, int world);
int svo_shard_table_size(svo_ctx* ctx, uint64_t* n_u64);
int svo_shard_count(svo_ctx* ctx, uint64_t* dev_table);
int svo_shard_emit(svo_ctx* ctx, const uint64_t* dev_table, uint64_t* n_voxels, uint64_t* n_nodes, uint64_t* n_data);
int svo_shard_ranges(svo_ctx* ctx, uint64_t* node_lo, uint64_t* node_hi, uint64_t* data_lo, uint64_t* data_hi);

/* The host-side merge svo_shard_emit performs, as a pure function (no GPU, no context): from the summed table of a job
 * it derives the global counts, rank `rank`'s range of the .octreenodes file and the records of the shared upper octree
 * levels that fall into that range (rec_pos[i] = file position, rec_words[3i..3i+2] = the 24-byte record, octree_io.h:62-66).
 * For callers that size or pre-allocate the output files before emission, and for CPU tests of the exchange protocol.
 * rec_pos / rec_words may be NULL (counts only). Errors: svo_last_error(NULL). */
typedef struct svo_shard_layout {
    uint64_t n_voxels, n_nodes;          /* global: "Total amount of voxels", .octree n_nodes          */
    uint64_t node_lo, node_hi;           /* this rank's records of .octreenodes                          */
    uint64_t leaf_offset;                /* voxels in the slabs of lower ranks (payload: data index offset) */
    uint64_t n_voxels_local;             /* voxels in this rank's slab                                    */
    uint64_t n_upper_records;            /* shared-level records inside [node_lo, node_hi)               */
} svo_shard_layout;
int svo_shard_layout_from_table(const svo_params* params, int rank, int world, const uint64_t* host_table, uint64_t n_u64,
                                svo_shard_layout* out, uint64_t* rec_pos, uint64_t* rec_words, uint64_t rec_capacity);

/* ---- multi-GPU: triangle dispatch over NVLink peer memory (copying alternative) --
 *
 * For callers that want every rank to end up with a private, compact copy of
 * its triangles (e.g. the slices cannot stay resident). Replaces, for the sharded build, the reference's partition files as the way
 * triangles reach the worker that voxelizes them (partitioner.cpp:101-149 writes
 * every triangle into the .tripdata file of each partition it touches; here it
 * is written into the HBM of each RANK whose slab it touches). Every rank starts
 * with a contiguous slice of the .tridata file in its own HBM (rank r holds the
 * r-th slice, in file order); one kernel pass counts, a second writes the
 * records straight into the peers' inboxes with NVLink stores -- no host
 * staging, no NCCL on the data path. The inbox ends up ordered by (source rank,
 * position in the slice) = file order, so the payload rule "first triangle in
 * file order wins" (voxelizer.cpp:263) is kept. After svo_shard_dispatch_finish
 * the inbox IS the context's triangle set (as after svo_set_triangles_device):
 * continue with svo_partition / svo_voxelize / svo_shard_count / ... as usual.
 *
 *   svo_shard_dispatch_create(capacity, fpt, &inbox, &ctrl)   once; capacity = total triangle count of the
 *                                                             mesh is always enough; same value on every rank
 *   <share the two device pointers with the peers>            svo_ipc_export / svo_ipc_open across processes,
 *                                                             the raw pointers inside one process
 *   svo_shard_dispatch_attach(inbox_ptrs, ctrl_ptrs)          `world` entries each, entry [rank] = own buffers
 *   per job:  svo_shard_dispatch_count -> svo_shard_dispatch_send -> svo_shard_dispatch_finish
 *             (three calls so that one host thread can drive several ranks: issue each phase for all
 *              ranks before the next; the cross-rank waits happen on the device, in stream order)
 * Errors: SVO_E_RANGE if an inbox is too small (reported by every rank), SVO_E_CUDA if a peer never
 * arrives (device-side wait gives up after a few seconds). At most 16 ranks. */
int svo_shard_dispatch_create(svo_ctx* ctx, uint64_t capacity_tris, int floats_per_tri, void** dev_inbox, void** dev_ctrl);
int svo_shard_dispatch_attach(svo_ctx* ctx, void* const* inbox_ptrs, void* const* ctrl_ptrs);
int svo_shard_dispatch_count(svo_ctx* ctx, const svo_params* params, const float* dev_local_tris, uint64_t n_local, int floats_per_tri);
int svo_shard_dispatch_send(svo_ctx* ctx);
int svo_shard_dispatch_finish(svo_ctx* ctx, uint64_t* n_received);
/* cudaIpcGetMemHandle / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle as plain bytes (64-byte handle). */
int svo_ipc_export(const void* dev_ptr, void* handle64);
int svo_ipc_open(const void* handle64, void** dev_ptr);
int svo_ipc_close(void* dev_ptr);

/* ---- multi-GPU: remote staging of triangle slices (no copy at all) ----------
 *
 * The default multi-GPU input path. The triangle file stays where it was
 * loaded: rank r keeps the r-th slice (file order) in a library-owned buffer in
 * its own HBM, mapped by every peer. Per job each rank lists, for every
 * destination rank, the 128-triangle staging blocks of ITS slice that touch the
 * destination's slab (one light pass over the local slice; 4 bytes per block
 * stored into the destination's list buffer over NVLink). The destination's
 * voxelizer then stages exactly those blocks straight from the owner's HBM with
 * NVLink loads, overlapped with the math of its other resident blocks -- the
 * transfer is fused into the voxelizer kernel, triangle records are never
 * copied, and per-rank work does not grow with the number of ranks. Global
 * triangle index = file position, so the payload rule "first triangle in file
 * order wins" (voxelizer.cpp:263) is kept.
 *
 *   svo_shard_slice_create(capacity, fpt, &window)   once; capacity = max triangles per slice, the SAME value on
 *                                                    every rank. The window is ONE device allocation holding the
 *                                                    control block, exchange table, block lists and the slice.
 *   <share the window pointer with the peers>        svo_ipc_export / svo_ipc_open across processes, the raw
 *                                                    pointer inside one process (svo_shard_slice_attach enables peer
 *                                                    access between the devices itself)
 *   svo_shard_slice_attach(windows)                  `world` entries, entry [rank] = own window
 *   svo_shard_slice_upload(src, n_local)     host or device source -> this rank's slice buffer (stream ordered;
 *                                            waits on the device until the peers have finished reading the old one)
 *   per job: svo_shard_slice_publish(params, n_total)            n_total = triangles in all slices (sizes work queues)
 *            svo_partition / svo_voxelize / svo_shard_count / svo_shard_exchange / svo_shard_emit
 *            (svo_shard_exchange replaces the caller's all-reduce of the table: every rank stores its own, disjoint
 *             entries into the peers' windows with NVLink stores; no library collective on the path)
 *   svo_shard_slice_fence                    stream-ordered wait until every peer has finished reading this rank's
 *                                            slice for the last published job (call before writing into `slice`
 *                                            yourself; upload and publish do it for you). The library keeps the bounding
 *                                            boxes of the slice's 32-triangle units from one job to the next; they are
 *                                            recomputed after svo_shard_slice_upload / _begin / _append / _fence, so a
 *                                            caller that writes the slice itself MUST call svo_shard_slice_fence first.
 * Per-partition counts (svo_partition's part_tricounts) are not available in this mode. At most 16 ranks. */
int svo_shard_slice_create(svo_ctx* ctx, uint64_t capacity_tris, int floats_per_tri, void** dev_window);
int svo_shard_slice_attach(svo_ctx* ctx, void* const* windows);
int svo_shard_exchange(svo_ctx* ctx, uint64_t* dev_table);
int svo_shard_slice_upload(svo_ctx* ctx, const float* src, uint64_t n_local);
/* Streamed variant of svo_shard_slice_upload (what svo_triangles_begin / _append are to svo_set_triangles): announce
 * the slice size, then append chunks from two alternating pinned buffers; same double-buffer contract. */
int svo_shard_slice_begin(svo_ctx* ctx, uint64_t n_local);
int svo_shard_slice_append(svo_ctx* ctx, const float* host_chunk, uint64_t n_chunk_tris);
int svo_shard_slice_publish(svo_ctx* ctx, const svo_params* params, uint64_t n_total);
int svo_shard_slice_fence(svo_ctx* ctx);

/* ---- whole path ---------------------------------------------------------- */

/* main.cpp:298-389 in one call: partition + voxelize + build from HOST triangle
 * records, results copied into caller buffers when they are non-NULL
 * (nodes_capacity / data_capacity in records; SVO_E_RANGE if too small -- the
 * counts are still returned so the caller can retry with larger buffers via
 * svo_fetch_*). */
int svo_run(svo_ctx* ctx, const svo_params* params,
            const float* host_tris, uint64_t n_tris,
            void* nodes_dst, uint64_t nodes_capacity,
            void* data_dst, uint64_t data_capacity,
            svo_stats* stats);

int svo_get_stats(svo_ctx* ctx, svo_stats* stats);

/* Voxels found in every logical partition (the "found N new voxels" line of the reference's -v output, main.cpp:348;
 * voxelizer.cpp:289 counts them in `nfilled`). After svo_build; one GPU (a sharded context reports its own slab). */
int svo_partition_voxel_counts(svo_ctx* ctx, uint64_t* counts, uint64_t capacity);

/* Blocks until all work queued on ctx's stream has finished. */
int svo_synchronize(svo_ctx* ctx);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers that
 * want svo_set_triangles / svo_fetch_* to run at full PCIe speed. */
void* svo_host_alloc(size_t bytes);
void  svo_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H_ */
