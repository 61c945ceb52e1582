/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for the voxelize-and-build path.
 *
 * Plain-C restatement of Forceflow/ooc_svo_builder's hot path. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (libsvo_b200.so) never does and has no CPU
 * fallback.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py runs this restatement
 * against the unmodified reference compiled into oracle/_ref/ (oracle/Makefile),
 * and tests/golden/ holds digests of reference outputs generated in the build
 * container by tests/golden/make_golden.py.
 */
#ifndef SVO_ORACLE_H_
#define SVO_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* main.cpp:23 `enum ColorType` */
enum { ORACLE_COLOR_MODEL = 0, ORACLE_COLOR_FIXED = 1, ORACLE_COLOR_LINEAR = 2, ORACLE_COLOR_NORMAL = 3 };

typedef struct {
    uint64_t n_partitions;
    uint64_t n_voxels;   /* "Total amount of voxels" (main.cpp:391) */
    uint64_t n_nodes;    /* .octree n_nodes */
    uint64_t n_data;     /* .octree n_data  */
    uint8_t* nodes;      /* n_nodes * 24 bytes, .octreenodes image */
    uint8_t* data;       /* n_data  * 32 bytes, .octreedata image  */
} svo_oracle_result;

/* partitioner.cpp:12-28 */
uint64_t svo_oracle_estimate_partitions(uint64_t gridsize, uint64_t memory_limit_mb);

/* Formats a float the way `ostream << float` does and parses it back the way
 * `istream >> float` does: the .trip header round trip that main.cpp:304-311
 * applies to the mesh bbox before computing the voxelizer's unit length. */
float svo_oracle_text_roundtrip(float v);

/* partitioner.cpp:101-149 + BBoxBuffer.h:70-84: per-partition triangle counts.
 * counts has n_partitions entries. If lists != NULL it must hold n_partitions
 * pointers that receive malloc'd ascending triangle-index arrays. */
int svo_oracle_partition(const float* tris, uint64_t n_tris, int floats_per_tri,
                         float bbox_min0, float bbox_max0, uint64_t gridsize, uint64_t n_partitions,
                         uint64_t* counts, uint64_t** lists);

/* voxelizer.cpp:138-307 for one partition. `tri_ids` (may be NULL = all
 * triangles, in order) selects the partition's triangle list. voxels must hold
 * (morton_end - morton_start) bytes; it is cleared first. owner (may be NULL)
 * receives, per filled voxel, the index into tri_ids' order of the triangle that
 * claimed it. Returns the number of voxels set. */
uint64_t svo_oracle_voxelize(const float* tris, int floats_per_tri, const uint64_t* tri_ids, uint64_t n_ids,
                             uint64_t morton_start, uint64_t morton_end, float unitlength,
                             uint8_t* voxels, uint32_t* owner);

/* Whole pipeline, main.cpp:281-399: partition -> per partition voxelize + feed
 * the streaming OctreeBuilder -> finalize. floats_per_tri 9 = svo_builder_binary,
 * 21 = svo_builder. Returns 0 on success. */
int svo_oracle_build(const float* tris, uint64_t n_tris, int floats_per_tri,
                     float bbox_min0, float bbox_max0,
                     uint64_t gridsize, uint64_t memory_limit_mb,
                     int generate_levels, int color_mode,
                     svo_oracle_result* out);

/* OctreeBuilder only (OctreeBuilder.cpp), fed with ascending Morton codes
 * (binary mode: addVoxel(uint64)). */
int svo_oracle_build_from_codes(const uint64_t* codes, uint64_t n, uint64_t gridsize,
                                int generate_levels, svo_oracle_result* out);

void svo_oracle_free(svo_oracle_result* r);

/* 26 (default) = the reference's conservative test; 6 = the 6-separating variant (NOT in the reference: a restatement
 * of the published definition, see svo_oracle.c). Process-wide switch of svo_oracle_voxelize / svo_oracle_build. */
void svo_oracle_set_separability(int separability);

/* libmorton morton3D_64_encode / decode (x -> bit 0, y -> bit 1, z -> bit 2). */
uint64_t svo_oracle_morton_encode(uint32_t x, uint32_t y, uint32_t z);
void svo_oracle_morton_decode(uint64_t m, uint32_t* x, uint32_t* y, uint32_t* z);

#ifdef __cplusplus
}
#endif
#endif
