/* TEST INFRASTRUCTURE ONLY -- see svo_oracle.h.
 *
 * Plain-C restatement of the reference's voxelize-and-build path. Every function
 * cites the reference file:line it follows (paths relative to /root/reference/).
 * Built with -ffp-contract=off so every float operation rounds once, like the
 * reference built without -mfma.
 *
 * Parity status: PINNED against oracle/_ref (the unmodified reference binaries)
 * by tests/test_oracle_vs_reference.py and against tests/golden/.
 */
#include "svo_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Morton codes: src/libs/libmorton/include/morton3D.h:36-47, 163-178.        */
/* The reference uses LUTs; the bit layout (x lowest) is what matters.        */
/* ------------------------------------------------------------------------- */
static uint64_t spread3(uint64_t a) {
    a &= 0x1fffffULL;
    a = (a | a << 32) & 0x1f00000000ffffULL;
    a = (a | a << 16) & 0x1f0000ff0000ffULL;
    a = (a | a << 8) & 0x100f00f00f00f00fULL;
    a = (a | a << 4) & 0x10c30c30c30c30c3ULL;
    a = (a | a << 2) & 0x1249249249249249ULL;
    return a;
}
static uint32_t compact3(uint64_t a) {
    a &= 0x1249249249249249ULL;
    a = (a ^ (a >> 2)) & 0x10c30c30c30c30c3ULL;
    a = (a ^ (a >> 4)) & 0x100f00f00f00f00fULL;
    a = (a ^ (a >> 8)) & 0x1f0000ff0000ffULL;
    a = (a ^ (a >> 16)) & 0x1f00000000ffffULL;
    a = (a ^ (a >> 32)) & 0x1fffffULL;
    return (uint32_t)a;
}
uint64_t svo_oracle_morton_encode(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
void svo_oracle_morton_decode(uint64_t m, uint32_t* x, uint32_t* y, uint32_t* z) {
    *x = compact3(m);
    *y = compact3(m >> 1);
    *z = compact3(m >> 2);
}

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */
/* std::min / std::max as libstdc++ defines them (NaN behaviour included). */
static float stdminf(float a, float b) { return (b < a) ? b : a; }
static float stdmaxf(float a, float b) { return (a < b) ? b : a; }

/* static_cast<int>(float) as x86-64 cvttss2si does it: truncation, and the
 * "integer indefinite" value for NaN / out of range (voxelizer.cpp:191-196). */
static int f2i(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return (int)0x80000000;
    return (int)f;
}
/* svo_builder_util.h:50-52 */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* glm scalar semantics, see oracle/glm_shim/glm/glm.hpp */
static float dot3(const float* a, const float* b) {
    float px = a[0] * b[0], py = a[1] * b[1], pz = a[2] * b[2];
    return px + py + pz;
}
static float dot2(float ax, float ay, float bx, float by) {
    float px = ax * bx, py = ay * by;
    return px + py;
}
static void normalize3(const float* v, float* o) {
    float inv = 1.0f / sqrtf(dot3(v, v));
    o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv;
}

float svo_oracle_text_roundtrip(float v) {
    /* trip_tools.h:110-111 writes `outfile << float` (precision 6, %g);      */
    /* trip_tools.h:78 reads it back with `file >> float`.                    */
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)v);
    return strtof(buf, NULL);
}

/* partitioner.cpp:12-28 */
uint64_t svo_oracle_estimate_partitions(uint64_t gridsize, uint64_t memory_limit) {
    uint64_t required = (gridsize * gridsize * gridsize) / 1024 / 1024;
    if (required <= memory_limit) return 1;
    uint64_t numpartitions = 1, required_partition = required;
    while (required_partition > memory_limit) {
        required_partition /= 8;
        numpartitions *= 8;
    }
    return numpartitions;
}

/* intersection.h:9-18 */
static void tri_bbox(const float* t, float* mn, float* mx) {
    for (int k = 0; k < 3; k++) {
        mn[k] = stdminf(t[k], stdminf(t[3 + k], t[6 + k]));
        mx[k] = stdmaxf(t[k], stdmaxf(t[3 + k], t[6 + k]));
    }
}

/* partitioner.cpp:43-77 (createBuffers) + :117-126 + BBoxBuffer.h:70-84 + intersection.h:50-53 */
int svo_oracle_partition(const float* tris, uint64_t n_tris, int fpt,
                         float bbox_min0, float bbox_max0, uint64_t gridsize, uint64_t P,
                         uint64_t* counts, uint64_t** lists) {
    if (P == 1) { /* partition_one, partitioner.cpp:80-98: plain copy */
        counts[0] = n_tris;
        if (lists) {
            lists[0] = (uint64_t*)malloc(sizeof(uint64_t) * (n_tris ? n_tris : 1));
            for (uint64_t i = 0; i < n_tris; i++) lists[0][i] = i;
        }
        return 0;
    }
    float unitlength = (bbox_max0 - bbox_min0) / (float)gridsize;     /* partitioner.cpp:45 */
    uint64_t morton_part = (gridsize * gridsize * gridsize) / P;        /* :46 */
    float* bmin = (float*)malloc(sizeof(float) * 3 * P);
    float* bmax = (float*)malloc(sizeof(float) * 3 * P);
    for (uint64_t i = 0; i < P; i++) {
        uint32_t gmin[3], gmax[3];
        svo_oracle_morton_decode(morton_part * i, &gmin[0], &gmin[1], &gmin[2]);            /* :52 */
        svo_oracle_morton_decode(morton_part * (i + 1) - 1, &gmax[0], &gmax[1], &gmax[2]);  /* :53 */
        for (int k = 0; k < 3; k++) {
            bmin[3 * i + k] = (float)gmin[k] * unitlength;         /* :54-56 */
            bmax[3 * i + k] = (float)(gmax[k] + 1u) * unitlength;  /* :57-59 */
        }
        counts[i] = 0;
    }
    uint64_t* cap = NULL;
    if (lists) {
        cap = (uint64_t*)calloc(P, sizeof(uint64_t));
        for (uint64_t i = 0; i < P; i++) lists[i] = NULL;
    }
    for (uint64_t t = 0; t < n_tris; t++) {
        float mn[3], mx[3];
        tri_bbox(tris + t * fpt, mn, mx);
        for (uint64_t j = 0; j < P; j++) {
            const float* a = bmin + 3 * j; const float* b = bmax + 3 * j;
            /* intersectBoxBox(bbox, bbox_world): reject only on strict < / > */
            if (mx[0] < a[0] || mx[1] < a[1] || mx[2] < a[2] || mn[0] > b[0] || mn[1] > b[1] || mn[2] > b[2]) continue;
            if (lists) {
                if (counts[j] == cap[j]) {
                    cap[j] = cap[j] ? cap[j] * 2 : 1024;
                    lists[j] = (uint64_t*)realloc(lists[j], cap[j] * sizeof(uint64_t));
                }
                lists[j][counts[j]] = t;
            }
            counts[j]++;
        }
    }
    free(bmin); free(bmax); free(cap);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* voxelizer.cpp:138-307  voxelize_schwarz_method                            */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t morton; float color[3]; float normal[3]; } voxeldata; /* VoxelData.h:10-17, 32 bytes */

typedef struct { voxeldata* v; uint64_t n, cap; } vdvec;
static void vd_push(vdvec* a, const voxeldata* x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 4096; a->v = (voxeldata*)realloc(a->v, a->cap * sizeof(voxeldata)); }
    a->v[a->n++] = *x;
}

/* BarycentricCoords.h:4-33 */
static void barycentric_colour(const float* t /*21 floats*/, const float* n, float vx, float vy, float vz, float* out) {
    const float *v0 = t, *v1 = t + 3, *v2 = t + 6;
    float voxel[3] = { vx, vy, vz };
    float coeffD = -dot3(v0, n);                                                  /* :25 */
    float k = (dot3(voxel, n) + coeffD) / (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]); /* :28 */
    float point[3] = { voxel[0] - k * n[0], voxel[1] - k * n[1], voxel[2] - k * n[2] }; /* :29 */
    /* mat3(v0,v1,v2) column major: m[c][r]; glm::inverse */
    float m[3][3] = { { v0[0], v0[1], v0[2] }, { v1[0], v1[1], v1[2] }, { v2[0], v2[1], v2[2] } };
    float ood = 1.0f / (
        + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
        - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
        + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    float I[3][3];
    I[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
    I[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
    I[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
    I[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
    I[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
    I[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
    I[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
    I[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
    I[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
    float b[3];
    for (int r = 0; r < 3; r++) b[r] = I[0][r] * point[0] + I[1][r] * point[1] + I[2][r] * point[2];
    /* InterpolateValue, :4-7 */
    const float *c0 = t + 12, *c1 = t + 15, *c2 = t + 18;
    for (int r = 0; r < 3; r++) out[r] = b[0] * c0[r] + b[1] * c1[r] + b[2] * c2[r];
}

/* Separability of the voxelization. 26 = the reference's conservative test (voxelizer.cpp:138-307). 6 = the
 * 6-separating ("thin") variant of the same paper (Schwarz & Seidel 2010, "Fast parallel surface and solid
 * voxelization on GPUs", sec. 4.2). The reference does NOT implement it: there is nothing to pin this
 * restatement to but the published definition, so parity for this variant is "CUDA path == this restatement", not
 * "== reference". Definition used (same setup arithmetic as the conservative test, glm op order):
 *   k      = dominant axis of the triangle normal n (|n_x| >= |n_y| and |n_x| >= |n_z| -> x; else |n_y| >= |n_z| -> y; else z)
 *   target = the axis-parallel segment through the voxel centre along k (its two face centres)
 *   plane  : d1 = n . (c1 - v0), d2 = n . (c2 - v0) with c1 = (h, h, h) but c1[k] = 0, c2 = (h, h, h) but c2[k] = u, h = u / 2;
 *            reject iff (n . p + d1) (n . p + d2) > 0                    (the plane separates the segment's end points)
 *   edges  : ONLY the projection orthogonal to k, evaluated at the projected voxel CENTRE:
 *            d_e = -(n_e . v_i) + h n_e.x + h n_e.y ; reject iff n_e . p + d_e < 0
 * The candidate box (clamped bounding box) is the conservative one. */
static int g_separability = 26;
void svo_oracle_set_separability(int s) { g_separability = (s == 6) ? 6 : 26; }

/* Shared body. If `pay` != NULL payload records are appended (svo_builder);
 * otherwise this is the BINARY_VOXELIZATION build. */
static uint64_t voxelize_partition(const float* tris, int fpt, const uint64_t* ids, uint64_t n_ids,
                                   uint64_t morton_start, uint64_t morton_end, float unitlength,
                                   uint8_t* voxels, uint32_t* owner, vdvec* pay) {
    uint64_t nfilled = 0;
    memset(voxels, 0, (size_t)(morton_end - morton_start));                       /* :144 */
    uint32_t pmin[3], pmax[3];
    svo_oracle_morton_decode(morton_start, &pmin[0], &pmin[1], &pmin[2]);          /* :149 */
    svo_oracle_morton_decode(morton_end - 1, &pmax[0], &pmax[1], &pmax[2]);        /* :150 */
    float unit_div = 1.0f / unitlength;                                            /* :164 */
    float delta_p[3] = { unitlength, unitlength, unitlength };

    for (uint64_t q = 0; q < n_ids; q++) {
        const float* t = tris + (ids ? ids[q] : q) * fpt;
        const float *v0 = t, *v1 = t + 3, *v2 = t + 6;
        float wmn[3], wmx[3];
        tri_bbox(t, wmn, wmx);                                                     /* :189 */
        int gmn[3], gmx[3];
        for (int k = 0; k < 3; k++) {
            gmn[k] = clampi(f2i(wmn[k] * unit_div), (int)pmin[k], (int)pmax[k]);   /* :191-204 */
            gmx[k] = clampi(f2i(wmx[k] * unit_div), (int)pmin[k], (int)pmax[k]);
        }
        float e0[3], e1[3], e2[3], cr[3], n[3];
        for (int k = 0; k < 3; k++) { e0[k] = v1[k] - v0[k]; e1[k] = v2[k] - v1[k]; e2[k] = v0[k] - v2[k]; } /* :207-209 */
        cr[0] = e0[1] * e1[2] - e1[1] * e0[2];
        cr[1] = e0[2] * e1[0] - e1[2] * e0[0];
        cr[2] = e0[0] * e1[1] - e1[0] * e0[1];
        normalize3(cr, n);                                                         /* :210 */
        float c[3] = { 0.0f, 0.0f, 0.0f };                                         /* :212-215 */
        if (n[0] > 0) c[0] = unitlength;
        if (n[1] > 0) c[1] = unitlength;
        if (n[2] > 0) c[2] = unitlength;
        float a1[3], a2[3];
        for (int k = 0; k < 3; k++) { a1[k] = c[k] - v0[k]; a2[k] = (delta_p[k] - c[k]) - v0[k]; }
        float d1 = dot3(n, a1);                                                    /* :216 */
        float d2 = dot3(n, a2);                                                    /* :217 */
        /* projection planes: (A,B) index pairs XY=(0,1), YZ=(1,2), ZX=(2,0); flipped by the third normal component */
        static const int PA[3] = { 0, 1, 2 }, PB[3] = { 1, 2, 0 }, PN[3] = { 2, 0, 1 };
        float ne[3][3][2], de[3][3];
        const float* E[3] = { e0, e1, e2 };
        const float* V[3] = { v0, v1, v2 };
        for (int p = 0; p < 3; p++) {
            int A = PA[p], B = PB[p];
            for (int j = 0; j < 3; j++) {
                float nx = -1.0f * E[j][B], ny = E[j][A];                          /* :220-222 etc. */
                if (n[PN[p]] < 0.0f) { nx = -1.0f * nx; ny = -1.0f * ny; }         /* :223-227 */
                ne[p][j][0] = nx; ne[p][j][1] = ny;
                de[p][j] = (-1.0f * dot2(nx, ny, V[j][A], V[j][B]))
                           + stdmaxf(0.0f, unitlength * nx) + stdmaxf(0.0f, unitlength * ny); /* :228-230 */
            }
        }
        if (g_separability == 6) {
            const float ax = fabsf(n[0]), ay = fabsf(n[1]), az = fabsf(n[2]);
            const int k = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
            const float h = unitlength * 0.5f;
            float c1[3] = { h, h, h }, c2[3] = { h, h, h };
            c1[k] = 0.0f; c2[k] = unitlength;
            for (int q2 = 0; q2 < 3; q2++) { a1[q2] = c1[q2] - v0[q2]; a2[q2] = c2[q2] - v0[q2]; }
            d1 = dot3(n, a1);
            d2 = dot3(n, a2);
            for (int p = 0; p < 3; p++) {
                int A = PA[p], B = PB[p];
                for (int j = 0; j < 3; j++) {
                    if (PN[p] != k) { ne[p][j][0] = 0.0f; ne[p][j][1] = 0.0f; de[p][j] = 0.0f; continue; }
                    const float nx = ne[p][j][0], ny = ne[p][j][1];
                    de[p][j] = ((-1.0f * dot2(nx, ny, V[j][A], V[j][B])) + h * nx) + h * ny;
                }
            }
        }
        for (int x = gmn[0]; x <= gmx[0]; x++) {                                   /* :257-259 */
            for (int y = gmn[1]; y <= gmx[1]; y++) {
                for (int z = gmn[2]; z <= gmx[2]; z++) {
                    uint64_t index = svo_oracle_morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z);
                    if (voxels[index - morton_start] == 1) continue;               /* :263 */
                    float p[3] = { x * unitlength, y * unitlength, z * unitlength };
                    float nDOTp = dot3(n, p);
                    if ((nDOTp + d1) * (nDOTp + d2) > 0.0f) continue;              /* :268 */
                    int rej = 0;
                    for (int pl = 0; pl < 3 && !rej; pl++) {
                        float pa = p[PA[pl]], pb = p[PB[pl]];
                        for (int j = 0; j < 3; j++) {
                            if ((dot2(ne[pl][j][0], ne[pl][j][1], pa, pb) + de[pl][j]) < 0.0f) { rej = 1; break; } /* :273-287 */
                        }
                    }
                    if (rej) continue;
                    voxels[index - morton_start] = 1;                              /* :290 / :293 */
                    if (owner) owner[index - morton_start] = (uint32_t)q;
                    if (pay) {
                        voxeldata vd;
                        vd.morton = index;
                        barycentric_colour(t, n, x / unit_div, y / unit_div, z / unit_div, vd.color); /* :295-296 */
                        vd.normal[0] = t[9]; vd.normal[1] = t[10]; vd.normal[2] = t[11];              /* :299 t.normal */
                        vd_push(pay, &vd);
                    }
                    nfilled++;
                }
            }
        }
    }
    return nfilled;
}

uint64_t svo_oracle_voxelize(const float* tris, int fpt, const uint64_t* ids, uint64_t n_ids,
                             uint64_t morton_start, uint64_t morton_end, float unitlength,
                             uint8_t* voxels, uint32_t* owner) {
    return voxelize_partition(tris, fpt, ids, n_ids, morton_start, morton_end, unitlength, voxels, owner, NULL);
}

/* ------------------------------------------------------------------------- */
/* OctreeBuilder.{h,cpp}, Node.h, octree_io.h                                */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t data, children_base;     /* Node.h:16-17 */
    int8_t off[8];                    /* Node.h:18 */
    voxeldata cache;                  /* Node.h:20 data_cache */
} onode;

typedef struct {
    onode (*buf)[8];                  /* b_buffers[depth][slot] */
    int* cnt;
    int maxdepth;
    uint64_t current_morton, max_morton;
    uint64_t node_pos, data_pos;
    int levels;
    uint8_t* nodes; uint64_t nodes_cap;
    uint8_t* data;  uint64_t data_cap;
} builder;

static void node_init(onode* n) {     /* Node.h:33-35 */
    memset(n, 0, sizeof *n);
    memset(n->off, 0xff, 8);
}
static int node_is_null(const onode* n) { /* Node.h:50-65 */
    static const int8_t leaf[8] = { -1, -1, -1, -1, -1, -1, -1, -1 };
    return memcmp(n->off, leaf, 8) == 0 && n->data == 0;
}
static uint64_t write_node(builder* b, const onode* n) { /* octree_io.h:62-66 */
    if ((b->node_pos + 1) * 24 > b->nodes_cap) {
        b->nodes_cap = b->nodes_cap ? b->nodes_cap * 2 : (1u << 16);
        b->nodes = (uint8_t*)realloc(b->nodes, b->nodes_cap);
    }
    uint8_t* p = b->nodes + b->node_pos * 24;
    memcpy(p, &n->data, 8); memcpy(p + 8, &n->children_base, 8); memcpy(p + 16, n->off, 8);
    return b->node_pos++;
}
static uint64_t write_data(builder* b, const voxeldata* v) { /* octree_io.h:49-53 */
    if ((b->data_pos + 1) * 32 > b->data_cap) {
        b->data_cap = b->data_cap ? b->data_cap * 2 : (1u << 16);
        b->data = (uint8_t*)realloc(b->data, b->data_cap);
    }
    uint8_t* p = b->data + b->data_pos * 32;
    memcpy(p, &v->morton, 8); memcpy(p + 8, v->color, 12); memcpy(p + 20, v->normal, 12);
    return b->data_pos++;
}

static unsigned ilog2(uint64_t v) { unsigned r = (unsigned)-1; while (v) { v >>= 1; r++; } return r; } /* svo_builder_util.h:37-44 */
static unsigned find_power_of_8(uint64_t n) { /* svo_builder_util.h:28-35 */
    if (n == 0) return 0;
    unsigned hi = 0;
    while (n >>= 1) hi++;
    return hi / 3;
}

static void builder_init(builder* b, uint64_t gridlength, int levels, int binary) { /* OctreeBuilder.cpp:4-32 */
    memset(b, 0, sizeof *b);
    b->levels = levels;
    b->maxdepth = (int)ilog2(gridlength);
    b->buf = calloc((size_t)b->maxdepth + 1, sizeof *b->buf);
    b->cnt = calloc((size_t)b->maxdepth + 1, sizeof(int));
    uint32_t maxm = (uint32_t)(gridlength - 1);
    b->max_morton = svo_oracle_morton_encode(maxm, maxm, maxm);
    voxeldata z; memset(&z, 0, sizeof z);
    write_data(b, &z);                                  /* :25 first data point is NULL */
    if (binary) {                                       /* :26-29 */
        voxeldata w; memset(&w, 0, sizeof w);
        w.color[0] = w.color[1] = w.color[2] = 1.0f;
        write_data(b, &w);
    }
}

static onode group_nodes(builder* b, const onode* buffer) { /* OctreeBuilder.cpp:58-102 */
    onode parent; node_init(&parent);
    int first = 1;
    for (int k = 0; k < 8; k++) {
        if (!node_is_null(&buffer[k])) {
            if (first) {
                parent.children_base = write_node(b, &buffer[k]);
                parent.off[k] = 0;
                first = 0;
            } else {
                parent.off[k] = (int8_t)(write_node(b, &buffer[k]) - parent.children_base);
            }
        } else {
            parent.off[k] = -1;
        }
    }
    if (b->levels) {                                   /* :82-99 */
        voxeldata d; memset(&d, 0, sizeof d);
        float notnull = 0.0f;
        for (int i = 0; i < 8; i++) {
            if (!node_is_null(&buffer[i])) notnull++;
            for (int r = 0; r < 3; r++) { d.color[r] += buffer[i].cache.color[r]; d.normal[r] += buffer[i].cache.normal[r]; }
        }
        float tn[3];
        for (int r = 0; r < 3; r++) { d.color[r] = d.color[r] / notnull; tn[r] = d.normal[r] / notnull; }
        normalize3(tn, d.normal);
        parent.data = write_data(b, &d);
        parent.cache = d;
    }
    return parent;
}

static void refine_buffers(builder* b, int start_depth) { /* OctreeBuilder.cpp:112-128 */
    for (int d = start_depth; d >= 0; d--) {
        if (b->cnt[d] == 8) {
            int empty = 1;
            for (int k = 0; k < 8; k++) if (!node_is_null(&b->buf[d][k])) { empty = 0; break; } /* OctreeBuilder.h:50-57 */
            onode up;
            if (empty) node_init(&up); else up = group_nodes(b, b->buf[d]);
            b->buf[d - 1][b->cnt[d - 1]++] = up;
            b->cnt[d] = 0;
        } else break;
    }
}

static uint64_t pow8(int e) { return (uint64_t)pow(8.0, e); } /* the reference uses pow(8.0, ..) in double */

static void add_empty_voxel(builder* b, int buffer) { /* OctreeBuilder.cpp:105-109 */
    onode n; node_init(&n);
    b->buf[buffer][b->cnt[buffer]++] = n;
    refine_buffers(b, buffer);
    b->current_morton = (uint64_t)(b->current_morton + pow(8.0, b->maxdepth - buffer));
}
static int highest_non_empty_buffer(const builder* b) { /* OctreeBuilder.h:60-70 */
    int highest = b->maxdepth;
    for (int k = b->maxdepth; k >= 0; k--) {
        if (b->cnt[k] == 0) highest--; else return highest;
    }
    return highest;
}
static int best_fill_buffer(const builder* b, uint64_t budget) { /* OctreeBuilder.h:73-80 */
    int s = b->maxdepth - (int)find_power_of_8(budget);
    if (s == b->maxdepth) return b->maxdepth;
    int h = highest_non_empty_buffer(b);
    return s > h ? s : h;
}
static void fast_add_empty(builder* b, uint64_t budget) { /* OctreeBuilder.h:83-91 */
    uint64_t r = budget;
    while (r > 0) {
        unsigned buffer = (unsigned)best_fill_buffer(b, r);
        add_empty_voxel(b, (int)buffer);
        r -= pow8(b->maxdepth - (int)buffer);
    }
}
static void add_voxel_binary(builder* b, uint64_t m) { /* OctreeBuilder.cpp:131-146 */
    if (m != b->current_morton) fast_add_empty(b, m - b->current_morton);
    onode n; node_init(&n);
    n.data = 1;
    b->buf[b->maxdepth][b->cnt[b->maxdepth]++] = n;
    refine_buffers(b, b->maxdepth);
    b->current_morton++;
}
static void add_voxel_payload(builder* b, const voxeldata* v) { /* OctreeBuilder.cpp:149-168 */
    if (v->morton != b->current_morton) fast_add_empty(b, v->morton - b->current_morton);
    onode n; node_init(&n);
    n.data = write_data(b, v);
    n.cache = *v;
    b->buf[b->maxdepth][b->cnt[b->maxdepth]++] = n;
    refine_buffers(b, b->maxdepth);
    b->current_morton++;
}
static void finalize_tree(builder* b) { /* OctreeBuilder.cpp:34-55 */
    if (b->current_morton < b->max_morton) fast_add_empty(b, (b->max_morton - b->current_morton) + 1);
    write_node(b, &b->buf[0][0]);
}
static void builder_release(builder* b, svo_oracle_result* out) {
    out->n_nodes = b->node_pos; out->n_data = b->data_pos;
    out->nodes = b->nodes; out->data = b->data;
    free(b->buf); free(b->cnt);
}

static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
static int cmp_vd(const void* a, const void* b) { uint64_t x = ((const voxeldata*)a)->morton, y = ((const voxeldata*)b)->morton; return x < y ? -1 : x > y; }

int svo_oracle_build_from_codes(const uint64_t* codes, uint64_t n, uint64_t gridsize, int levels, svo_oracle_result* out) {
    builder b; builder_init(&b, gridsize, levels, 1);
    for (uint64_t i = 0; i < n; i++) add_voxel_binary(&b, codes[i]);
    finalize_tree(&b);
    memset(out, 0, sizeof *out);
    out->n_partitions = 1; out->n_voxels = n;
    builder_release(&b, out);
    return 0;
}

/* main.cpp:281-399 */
int svo_oracle_build(const float* tris, uint64_t n_tris, int fpt, float bbox_min0, float bbox_max0,
                     uint64_t gridsize, uint64_t memory_limit, int levels, int color_mode, svo_oracle_result* out) {
    if (fpt != 9 && fpt != 21) return -1;
    int binary = (fpt == 9);
    memset(out, 0, sizeof *out);
    uint64_t P = svo_oracle_estimate_partitions(gridsize, memory_limit);      /* :298 */
    uint64_t* counts = (uint64_t*)calloc(P, sizeof(uint64_t));
    uint64_t** lists = (uint64_t**)calloc(P, sizeof(uint64_t*));
    svo_oracle_partition(tris, n_tris, fpt, bbox_min0, bbox_max0, gridsize, P, counts, lists); /* :300 */
    /* :304-311: the bbox is re-read from the .trip text header */
    float rmin = svo_oracle_text_roundtrip(bbox_min0), rmax = svo_oracle_text_roundtrip(bbox_max0);
    float unitlength = (rmax - rmin) / (float)gridsize;                       /* :311 */
    uint64_t morton_part = (gridsize * gridsize * gridsize) / P;              /* :312 */
    uint8_t* voxels = (uint8_t*)malloc((size_t)morton_part);                  /* :314 */
    builder b; builder_init(&b, gridsize, levels, binary);                    /* :326 */
    uint64_t nfilled = 0;
    vdvec pay = { 0, 0, 0 };
    for (uint64_t i = 0; i < P; i++) {                                        /* :329 */
        if (counts[i] == 0) continue;                                         /* :330 */
        uint64_t start = i * morton_part, end = (i + 1) * morton_part;
        pay.n = 0;
        nfilled += voxelize_partition(tris, fpt, lists[i], counts[i], start, end, unitlength, voxels, NULL, binary ? NULL : &pay); /* :347 */
        if (binary) {                                                         /* :355-368: both routes feed ascending codes */
            for (uint64_t j = 0; j < morton_part; j++) if (voxels[j]) add_voxel_binary(&b, start + j);
        } else {                                                              /* :371-384 */
            qsort(pay.v, (size_t)pay.n, sizeof(voxeldata), cmp_vd);
            for (uint64_t j = 0; j < pay.n; j++) {
                voxeldata* it = &pay.v[j];
                if (color_mode == ORACLE_COLOR_FIXED) {
                    it->color[0] = it->color[1] = it->color[2] = 1.0f;        /* main.cpp:33 fixed_color */
                } else if (color_mode == ORACLE_COLOR_LINEAR) {               /* svo_builder_util.h:14-18 */
                    uint32_t a, bb, c;                                        /* decode(m, z, y, x): first output -> "z" */
                    svo_oracle_morton_decode(it->morton, &a, &bb, &c);
                    it->color[0] = (float)c / (float)gridsize;
                    it->color[1] = (float)bb / (float)gridsize;
                    it->color[2] = (float)a / (float)gridsize;
                } else if (color_mode == ORACLE_COLOR_NORMAL) {               /* :379-381 */
                    float nn[3]; normalize3(it->normal, nn);
                    for (int r = 0; r < 3; r++) it->color[r] = (nn[r] + 1.0f) / 2.0f;
                }
                add_voxel_payload(&b, it);
            }
        }
    }
    finalize_tree(&b);                                                        /* :389 */
    out->n_partitions = P; out->n_voxels = nfilled;
    builder_release(&b, out);
    for (uint64_t i = 0; i < P; i++) free(lists[i]);
    free(lists); free(counts); free(voxels); free(pay.v);
    (void)cmp_u64;
    return 0;
}

void svo_oracle_free(svo_oracle_result* r) {
    free(r->nodes); free(r->data);
    r->nodes = r->data = NULL;
}
