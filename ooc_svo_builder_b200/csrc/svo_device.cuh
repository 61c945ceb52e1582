// svo_device.cuh -- device-side arithmetic shared by the kernels.
//
// Float parity: the reference's accept/reject predicates are float32 comparisons
// against zero (src/svo_builder/voxelizer.cpp:268-287), built for x86-64 without
// FMA. Every float operation here is an explicit round-to-nearest intrinsic
// (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fsqrt_rn), which nvcc never
// contracts into FMA, in exactly the reference's (and glm's scalar) operation
// order. The file is additionally compiled with -fmad=false.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace svo {

// ---------------------------------------------------------------------------
// exact float ops
// ---------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
// libstdc++ std::min / std::max (NaN behaviour included), intersection.h:11-16
__device__ __forceinline__ float stdmin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float stdmax(float a, float b) { return (a < b) ? b : a; }
// glm::dot: products first, summed left to right
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) {
    return fadd(fmul(ax, bx), fmul(ay, by));
}
// static_cast<int>(float) as x86-64 cvttss2si (voxelizer.cpp:191-196): truncate;
// NaN / out of range -> 0x80000000.
__device__ __forceinline__ int f2i(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(f);
}
// svo_builder_util.h:50-52
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------------------
// Morton codes, libmorton layout: x -> bit 0, y -> bit 1, z -> bit 2
// (src/libs/libmorton/include/morton3D.h:36-47). Bit tricks instead of LUTs.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t spread3(uint64_t a) {
    a &= 0x1fffffULL;
    a = (a | a << 32) & 0x1f00000000ffffULL;
    a = (a | a << 16) & 0x1f0000ff0000ffULL;
    a = (a | a << 8) & 0x100f00f00f00f00fULL;
    a = (a | a << 4) & 0x10c30c30c30c30c3ULL;
    a = (a | a << 2) & 0x1249249249249249ULL;
    return a;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint64_t a) {
    a &= 0x1249249249249249ULL;
    a = (a ^ (a >> 2)) & 0x10c30c30c30c30c3ULL;
    a = (a ^ (a >> 4)) & 0x100f00f00f00f00fULL;
    a = (a ^ (a >> 8)) & 0x1f0000ff0000ffULL;
    a = (a ^ (a >> 16)) & 0x1f00000000ffffULL;
    a = (a ^ (a >> 32)) & 0x1fffffULL;
    return (uint32_t)a;
}
__host__ __device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
// 10-bit-per-axis variant (bricks of grids up to 4096^3): 32-bit arithmetic only.
__host__ __device__ __forceinline__ uint32_t spread3_10(uint32_t a) {
    a &= 0x3ffu;
    a = (a | a << 16) & 0x030000ffu;
    a = (a | a << 8) & 0x0300f00fu;
    a = (a | a << 4) & 0x030c30c3u;
    a = (a | a << 2) & 0x09249249u;
    return a;
}

__host__ __device__ __forceinline__ uint64_t lowmask(int n) { return n >= 64 ? ~0ULL : ((1ULL << n) - 1ULL); }

// 8-bit mask of the non-zero bytes of w: bit k <-> byte k != 0.
__host__ __device__ __forceinline__ uint32_t nonzero_bytes(uint64_t w) {
    uint64_t t = w | (w >> 4);
    t |= t >> 2;
    t |= t >> 1;
    t &= 0x0101010101010101ULL;
    return (uint32_t)((t * 0x0102040810204080ULL) >> 56);
}

// Bit index inside a 4x4x4 brick word = low 6 Morton bits of (x,y,z).
__host__ __device__ __forceinline__ int brick_bit(int x, int y, int z) {
    return (x & 1) | ((y & 1) << 1) | ((z & 1) << 2) | ((x & 2) << 2) | ((y & 2) << 3) | ((z & 2) << 4);
}

// children_offset[8] of a node whose child mask is m (OctreeBuilder.cpp:61-79):
// byte c = rank of child c among the present children, 0xFF when absent.
__host__ __device__ __forceinline__ uint64_t child_offsets(uint32_t m) {
    uint64_t r = 0;
    uint32_t rank = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint64_t b = ((m >> c) & 1u) ? (uint64_t)rank++ : 0xFFULL;
        r |= b << (8 * c);
    }
    return r;
}

// The same as a 256-entry table in global memory, built at compile time (one cached 8-byte load instead of ~40
// instructions; the emitters look it up once per record).
struct ChildOffsetTable { unsigned long long v[256]; };
constexpr ChildOffsetTable make_child_offset_table() {
    ChildOffsetTable t{};
    for (unsigned m = 0; m < 256; m++) {
        unsigned long long r = 0;
        unsigned rank = 0;
        for (int c = 0; c < 8; c++) {
            const unsigned long long b = ((m >> c) & 1u) ? (unsigned long long)rank++ : 0xFFULL;
            r |= b << (8 * c);
        }
        t.v[m] = r;
    }
    return t;
}
__device__ const ChildOffsetTable g_child_offsets = make_child_offset_table();
__device__ __forceinline__ unsigned long long child_offsets_lut(uint32_t m) { return __ldg(&g_child_offsets.v[m & 0xffu]); }

// ---------------------------------------------------------------------------
// Schwarz-Seidel conservative test, voxelizer.cpp:206-254 (setup) and :266-287
// ---------------------------------------------------------------------------
struct TriSetup {
    float nx, ny, nz, d1, d2;
    // nine 2D edge functions. i = 3*plane + edge; plane 0 = XY (a=x, b=y),
    // 1 = YZ (a=y, b=z), 2 = ZX (a=z, b=x).
    float ea[9], eb[9], ed[9];
};

// v = v0 v1 v2 (9 floats); u = unitlength. six: the opt-in 6-separating variant (not in the reference; definition and
// float op order restated in the C file of the test oracle under oracle/): the plane test runs against the voxel's centre segment along the
// dominant normal axis, only the projection orthogonal to that axis is tested, at the projected voxel centre. It is
// expressed through the SAME setup record -- other d1 / d2, the two unused projections get (0, 0, +0) edge functions
// that always pass -- so that every kernel behind the setup (window evaluation, exact box pruning) serves both variants.
__device__ __forceinline__ void tri_setup(const float* v, float u, TriSetup& s, bool six = false) {
    float e0x = fsub(v[3], v[0]), e0y = fsub(v[4], v[1]), e0z = fsub(v[5], v[2]);   // :207
    float e1x = fsub(v[6], v[3]), e1y = fsub(v[7], v[4]), e1z = fsub(v[8], v[5]);   // :208
    float e2x = fsub(v[0], v[6]), e2y = fsub(v[1], v[7]), e2z = fsub(v[2], v[8]);   // :209
    // n = normalize(cross(e0, e1))                                                   :210
    float cx = fsub(fmul(e0y, e1z), fmul(e1y, e0z));
    float cy = fsub(fmul(e0z, e1x), fmul(e1z, e0x));
    float cz = fsub(fmul(e0x, e1y), fmul(e1x, e0y));
    float inv = fdiv(1.0f, fsqrt(dot3(cx, cy, cz, cx, cy, cz)));
    s.nx = fmul(cx, inv); s.ny = fmul(cy, inv); s.nz = fmul(cz, inv);
    // critical point                                                                 :212-217
    float ccx = (s.nx > 0.0f) ? u : 0.0f;
    float ccy = (s.ny > 0.0f) ? u : 0.0f;
    float ccz = (s.nz > 0.0f) ? u : 0.0f;
    s.d1 = dot3(s.nx, s.ny, s.nz, fsub(ccx, v[0]), fsub(ccy, v[1]), fsub(ccz, v[2]));
    s.d2 = dot3(s.nx, s.ny, s.nz, fsub(fsub(u, ccx), v[0]), fsub(fsub(u, ccy), v[1]), fsub(fsub(u, ccz), v[2]));
    const float ex[3] = { e0x, e1x, e2x }, ey[3] = { e0y, e1y, e2y }, ez[3] = { e0z, e1z, e2z };
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float vx = v[3 * j], vy = v[3 * j + 1], vz = v[3 * j + 2];
        // XY: n = (-e.y, e.x), flipped when n.z < 0                                  :220-230
        float a = fmul(-1.0f, ey[j]), b = ex[j];
        if (s.nz < 0.0f) { a = fmul(-1.0f, a); b = fmul(-1.0f, b); }
        s.ea[j] = a; s.eb[j] = b;
        s.ed[j] = fadd(fadd(fmul(-1.0f, dot2(a, b, vx, vy)), stdmax(0.0f, fmul(u, a))), stdmax(0.0f, fmul(u, b)));
        // YZ: n = (-e.z, e.y), flipped when n.x < 0                                  :232-242
        a = fmul(-1.0f, ez[j]); b = ey[j];
        if (s.nx < 0.0f) { a = fmul(-1.0f, a); b = fmul(-1.0f, b); }
        s.ea[3 + j] = a; s.eb[3 + j] = b;
        s.ed[3 + j] = fadd(fadd(fmul(-1.0f, dot2(a, b, vy, vz)), stdmax(0.0f, fmul(u, a))), stdmax(0.0f, fmul(u, b)));
        // ZX: n = (-e.x, e.z), flipped when n.y < 0                                  :244-254
        a = fmul(-1.0f, ex[j]); b = ez[j];
        if (s.ny < 0.0f) { a = fmul(-1.0f, a); b = fmul(-1.0f, b); }
        s.ea[6 + j] = a; s.eb[6 + j] = b;
        s.ed[6 + j] = fadd(fadd(fmul(-1.0f, dot2(a, b, vz, vx)), stdmax(0.0f, fmul(u, a))), stdmax(0.0f, fmul(u, b)));
    }
    if (six) {
        const float ax = fabsf(s.nx), ay = fabsf(s.ny), az = fabsf(s.nz);
        const int k = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);            // dominant axis
        const float h = fmul(u, 0.5f);
        const float c1x = k == 0 ? 0.0f : h, c1y = k == 1 ? 0.0f : h, c1z = k == 2 ? 0.0f : h;
        const float c2x = k == 0 ? u : h, c2y = k == 1 ? u : h, c2z = k == 2 ? u : h;
        s.d1 = dot3(s.nx, s.ny, s.nz, fsub(c1x, v[0]), fsub(c1y, v[1]), fsub(c1z, v[2]));
        s.d2 = dot3(s.nx, s.ny, s.nz, fsub(c2x, v[0]), fsub(c2y, v[1]), fsub(c2z, v[2]));
        // plane p (0 = XY, 1 = YZ, 2 = ZX) is orthogonal to axis (p + 2) % 3; its first / second coordinate are axis p, (p + 1) % 3
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const bool keep = ((p + 2) % 3) == k;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int i = 3 * p + j;
                if (!keep) { s.ea[i] = 0.0f; s.eb[i] = 0.0f; s.ed[i] = 0.0f; continue; }
                const float va = v[3 * j + p], vb = v[3 * j + (p + 1) % 3];
                s.ed[i] = fadd(fadd(fmul(-1.0f, dot2(s.ea[i], s.eb[i], va, vb)), fmul(h, s.ea[i])), fmul(h, s.eb[i]));
            }
        }
    }
}

__device__ __forceinline__ bool edge_pass(const TriSetup& s, int i, float pa, float pb) {
    return !(fadd(dot2(s.ea[i], s.eb[i], pa, pb), s.ed[i]) < 0.0f);            // :273-287
}
__device__ __forceinline__ bool plane_pass(const TriSetup& s, float px, float py, float pz) {
    float nd = dot3(s.nx, s.ny, s.nz, px, py, pz);
    return !(fmul(fadd(nd, s.d1), fadd(nd, s.d2)) > 0.0f);                     // :268
}
// full test at world-space voxel min corner p = (x*u, y*u, z*u)
__device__ __forceinline__ bool voxel_pass(const TriSetup& s, float px, float py, float pz) {
    if (!plane_pass(s, px, py, pz)) return false;
    if (!edge_pass(s, 0, px, py) || !edge_pass(s, 1, px, py) || !edge_pass(s, 2, px, py)) return false;
    if (!edge_pass(s, 3, py, pz) || !edge_pass(s, 4, py, pz) || !edge_pass(s, 5, py, pz)) return false;
    if (!edge_pass(s, 6, pz, px) || !edge_pass(s, 7, pz, px) || !edge_pass(s, 8, pz, px)) return false;
    return true;
}

// Exact early-out for a box of voxels [xa..xb] x [ya..yb] x [za..zb]: returns false
// only if NO voxel of the box can pass. Rounded float multiply and add are
// monotone, so each edge function attains its maximum over the box at the corner
// picked by the signs of its coefficients, and n.p attains min / max at opposite
// corners; a box is dropped only when that extreme already fails. NaN
// coefficients compare false everywhere and therefore never prune.
__device__ __forceinline__ bool box_may_pass(const TriSetup& s, float u,
                                             int xa, int xb, int ya, int yb, int za, int zb) {
    const float lx = fmul((float)xa, u), hx = fmul((float)xb, u);
    const float ly = fmul((float)ya, u), hy = fmul((float)yb, u);
    const float lz = fmul((float)za, u), hz = fmul((float)zb, u);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int pl = i / 3;
        const float la = pl == 0 ? lx : (pl == 1 ? ly : lz), ha = pl == 0 ? hx : (pl == 1 ? hy : hz);
        const float lb = pl == 0 ? ly : (pl == 1 ? lz : lx), hb = pl == 0 ? hy : (pl == 1 ? hz : hx);
        const float pa = (s.ea[i] >= 0.0f) ? ha : la;
        const float pb = (s.eb[i] >= 0.0f) ? hb : lb;
        if (fadd(dot2(s.ea[i], s.eb[i], pa, pb), s.ed[i]) < 0.0f) return false;
    }
    const float sx0 = (s.nx >= 0.0f) ? lx : hx, sx1 = (s.nx >= 0.0f) ? hx : lx;
    const float sy0 = (s.ny >= 0.0f) ? ly : hy, sy1 = (s.ny >= 0.0f) ? hy : ly;
    const float sz0 = (s.nz >= 0.0f) ? lz : hz, sz1 = (s.nz >= 0.0f) ? hz : lz;
    const float smin = dot3(s.nx, s.ny, s.nz, sx0, sy0, sz0);
    const float smax = dot3(s.nx, s.ny, s.nz, sx1, sy1, sz1);
    const float a0 = fadd(smin, s.d1), b0 = fadd(smin, s.d2);
    if (a0 > 0.0f && b0 > 0.0f && fmul(a0, b0) > 0.0f) return false;   // everything strictly above both planes
    const float a1 = fadd(smax, s.d1), b1 = fadd(smax, s.d2);
    if (a1 < 0.0f && b1 < 0.0f && fmul(a1, b1) > 0.0f) return false;   // everything strictly below both planes
    return true;
}

// During voxelization a brick word uses the LINEAR in-brick layout (bit = z*16 + y*4 + x): box-aligned
// windows can then be split into bricks with plain shifts. The octree wants the Morton layout (a byte =
// one 2x2x2 child); the permutation is applied once per occupied brick when the tile lists are built.
__host__ __device__ __forceinline__ uint64_t linear_to_morton64(uint64_t w) {
    uint64_t t;
    t = ((w >> 2) ^ w) & 0x0C0C0C0C0C0C0C0CULL;  w ^= t ^ (t << 2);     // swap index bits 1,2
    t = ((w >> 12) ^ w) & 0x0000F0F00000F0F0ULL; w ^= t ^ (t << 12);    // swap index bits 2,4
    t = ((w >> 8) ^ w) & 0x0000FF000000FF00ULL;  w ^= t ^ (t << 8);     // swap index bits 3,4
    return w;
}
__host__ __device__ __forceinline__ int linear_to_morton_bit(int i) { return brick_bit(i & 3, (i >> 2) & 3, i >> 4); }

// Conservative test of a whole 4x4x4 window of voxels anchored at (wx, wy, wz).
// Returns the hit mask in the linear layout (bit = lz*16 + ly*4 + lx), restricted to the extents
// (ea, eb, ec) <= 4 of the box inside the window. The nine edge functions only depend on two
// coordinates each, so they are evaluated on three 4x4 projections (products hoisted per axis) and
// expanded with multiplies; every rounded operation is the one the per-voxel test performs.
//
// (ub, uc) are WARP-UNIFORM upper bounds of (eb, ec): the row (y) and slice (z) loops run to them, so there is
// no divergence; rows / slices beyond them lie outside every box of the warp. The loops are deliberately NOT
// unrolled (only the 4 cells of a row are): fully unrolled, the kernel exceeded the instruction cache and
// stalled on instruction fetch (ncu: no_instruction was the top stall reason).
//
// Mask assembly from sign bits: an edge value e = fadd(., ed) fails iff it is negative. ed is never -0 (its last
// addend is std::max(0.0f, .) >= +0), so e is never -0, and a NaN result of a CUDA add is the canonical
// 0x7FFFFFFF (sign clear, like the reference's "not < 0"). So fail = sign(e0)|sign(e1)|sign(e2), pushed into
// the mask with one funnel shift per cell (cells visited from the highest bit down). The plane test rejects
// iff fmul(a, b) > 0  <=>  sign(0 - fmul(a, b)) set (+-0 and NaN give a clear sign).
__device__ __forceinline__ uint64_t eval_window(const TriSetup& s, float u, int wx, int wy, int wz, int ea, int eb, int ec,
                                                int ub, int uc) {
    float px[4], py[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { px[i] = fmul((float)(wx + i), u); py[i] = fmul((float)(wy + i), u); }
    // ---- XY: edges 0..2, a = x, b = y; bit ly*4 + lx; rows = ly ----
    uint32_t fxy = 0;
    {
        float A[3][4];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) A[j][i] = fmul(s.ea[j], px[i]);
#pragma unroll 1
        for (int ly = ub - 1; ly >= 0; ly--) {
            const float pyr = fmul((float)(wy + ly), u);
            const float b0 = fmul(s.eb[0], pyr), b1 = fmul(s.eb[1], pyr), b2 = fmul(s.eb[2], pyr);
#pragma unroll
            for (int lx = 3; lx >= 0; lx--) {
                const uint32_t f = __float_as_uint(fadd(fadd(A[0][lx], b0), s.ed[0])) | __float_as_uint(fadd(fadd(A[1][lx], b1), s.ed[1])) |
                                   __float_as_uint(fadd(fadd(A[2][lx], b2), s.ed[2]));
                fxy = __funnelshift_l(f, fxy, 1);
            }
        }
    }
    // ---- YZ (edges 3..5, a = y, b = z, bit lz*4 + ly), ZX (edges 6..8, a = z, b = x, bit lz*4 + lx) and the
    //      plane test (voxelizer.cpp:266-268; n.p summed x, then y, then z like dot3): rows = lz ----
    uint32_t fyz = 0, fzx = 0;
    uint64_t reject = 0;
    {
        float Ayz[3][4], Bzx[3][4], nxp[4];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) { Ayz[j][i] = fmul(s.ea[3 + j], py[i]); Bzx[j][i] = fmul(s.eb[6 + j], px[i]); }
#pragma unroll
        for (int i = 0; i < 4; i++) nxp[i] = fmul(s.nx, px[i]);
#pragma unroll 1
        for (int lz = uc - 1; lz >= 0; lz--) {
            const float pzr = fmul((float)(wz + lz), u);
            const float by0 = fmul(s.eb[3], pzr), by1 = fmul(s.eb[4], pzr), by2 = fmul(s.eb[5], pzr);
            const float az0 = fmul(s.ea[6], pzr), az1 = fmul(s.ea[7], pzr), az2 = fmul(s.ea[8], pzr);
            const float nzp = fmul(s.nz, pzr);
#pragma unroll
            for (int ly = 3; ly >= 0; ly--) {
                const uint32_t f = __float_as_uint(fadd(fadd(Ayz[0][ly], by0), s.ed[3])) | __float_as_uint(fadd(fadd(Ayz[1][ly], by1), s.ed[4])) |
                                   __float_as_uint(fadd(fadd(Ayz[2][ly], by2), s.ed[5]));
                fyz = __funnelshift_l(f, fyz, 1);
            }
#pragma unroll
            for (int lx = 3; lx >= 0; lx--) {
                const uint32_t f = __float_as_uint(fadd(fadd(az0, Bzx[0][lx]), s.ed[6])) | __float_as_uint(fadd(fadd(az1, Bzx[1][lx]), s.ed[7])) |
                                   __float_as_uint(fadd(fadd(az2, Bzx[2][lx]), s.ed[8]));
                fzx = __funnelshift_l(f, fzx, 1);
            }
            uint32_t rs = 0;
#pragma unroll 1
            for (int ly = ub - 1; ly >= 0; ly--) {
                const float nyp = fmul(s.ny, fmul((float)(wy + ly), u));
#pragma unroll
                for (int lx = 3; lx >= 0; lx--) {
                    const float nd = fadd(fadd(nxp[lx], nyp), nzp);
                    rs = __funnelshift_l(__float_as_uint(fsub(0.0f, fmul(fadd(nd, s.d1), fadd(nd, s.d2)))), rs, 1);
                }
            }
            reject |= (uint64_t)(rs & 0xffffu) << (16 * lz);
        }
    }
    const uint32_t mxy = ~fxy & 0xffffu, myz = ~fyz & 0xffffu, mzx = ~fzx & 0xffffu;
    // expand the projections to the 64 voxels and restrict to the box
    uint64_t cand = (uint64_t)mxy * 0x0001000100010001ULL;
    {
        uint64_t x = myz;                                   // bit k -> nibble k
        x = (x | x << 24) & 0x000000FF000000FFULL;
        x = (x | x << 12) & 0x000F000F000F000FULL;
        x = (x | x << 6) & 0x0303030303030303ULL;
        x = (x | x << 3) & 0x1111111111111111ULL;
        cand &= x * 0xFULL;
        uint64_t y = mzx;                                   // nibble lz -> 16-bit slice lz, repeated over ly
        y = (y | y << 24) & 0x000000FF000000FFULL;
        y = (y | y << 12) & 0x000F000F000F000FULL;
        cand &= y * 0x1111ULL;
    }
    cand &= (uint64_t)((1u << ea) - 1u) * 0x1111111111111111ULL;
    cand &= (uint64_t)((1u << (4 * eb)) - 1u) * 0x0001000100010001ULL;
    cand &= lowmask(16 * ec);
    return cand & ~reject;
}

struct GridBox { int x0, x1, y0, y1, z0, z1; };

// Triangle bbox in grid coordinates, clamped INTO the partition box
// (voxelizer.cpp:189-204, intersection.h:9-18).
__device__ __forceinline__ GridBox clamped_box(const float* v, float unit_div,
                                               int px0, int py0, int pz0, int side) {
    GridBox b;
    float mn, mx;
    mn = stdmin(v[0], stdmin(v[3], v[6])); mx = stdmax(v[0], stdmax(v[3], v[6]));
    b.x0 = clampi(f2i(fmul(mn, unit_div)), px0, px0 + side - 1);
    b.x1 = clampi(f2i(fmul(mx, unit_div)), px0, px0 + side - 1);
    mn = stdmin(v[1], stdmin(v[4], v[7])); mx = stdmax(v[1], stdmax(v[4], v[7]));
    b.y0 = clampi(f2i(fmul(mn, unit_div)), py0, py0 + side - 1);
    b.y1 = clampi(f2i(fmul(mx, unit_div)), py0, py0 + side - 1);
    mn = stdmin(v[2], stdmin(v[5], v[8])); mx = stdmax(v[2], stdmax(v[5], v[8]));
    b.z0 = clampi(f2i(fmul(mn, unit_div)), pz0, pz0 + side - 1);
    b.z1 = clampi(f2i(fmul(mx, unit_div)), pz0, pz0 + side - 1);
    return b;
}

}  // namespace svo
