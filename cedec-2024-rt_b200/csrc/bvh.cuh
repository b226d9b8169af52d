// bvh.cuh — compressed 8-wide BVH: node format, ray/triangle test and the traversal loop.
//
// Replaces HIPRT's hiprtGeomTraversalClosestCustomStack behind raytrace()/check_visibility()
// (common/raytrace.hpp:18-52; 06_ao_hiprt.cu:13-33).  B200 has no RT cores, so this is a software
// walk: 80-byte nodes holding eight children as 8-bit boxes on a per-node power-of-two grid (the
// node stream for 10 M triangles stays L2-resident: 126 MB L2), children pre-sorted into octant
// slots so that `slot ^ octant` is a front-to-back priority, one (base, bitmask) stack entry per
// tree level, triangles in 48-byte records fetched as 3 x LDG.128.
//
// The tree only culls.  Accept/reject and the t,u,v bits come from ray_triangle(), which is the
// reference's intersect_ray_triangle (common/core.hpp:91-136) operation for operation; ties on t are
// resolved towards the larger primitive id, which is what the reference's own brute-force loop does
// (04_ao.cu:14-24).  Boxes are padded at build time (bvh_build.cuh) by more than the rounding error
// of both that test and the slab arithmetic below, so no hit the exhaustive loop finds is lost.
#pragma once
#include "vecmath.cuh"

namespace crt
{
constexpr int kLeafMaxTris = 3;  // triangles per leaf child (3 x 8 slots = 24 hit-mask bits)
constexpr int kStackSize = 48;   // >= depth of the wide tree (checked at build time, CRT_ESTACK)

// 80 bytes, read as five 16-byte words.
struct alignas(16) WideNode
{
    float px, py, pz;            // grid origin (min corner of the node box)
    uint8_t ex, ey, ez;          // biased exponents: cell size on axis a is 2^(e_a - 127)
    uint8_t imask;               // bit s set: slot s is an inner node
    uint32_t child_base;         // first inner child; child in slot s is child_base + popc(imask & ((1<<s)-1))
    uint32_t tri_base;           // first triangle record of this node's leaf children
    uint8_t meta[8];             // 0: empty slot; inner: 0xff; leaf: (count << 5) | offset from tri_base
    uint8_t qlo[3][8];           // per axis, per slot: box min in grid cells (rounded down)
    uint8_t qhi[3][8];           // box max in grid cells (rounded up)
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

// 48 bytes: v0.xyz + primitive id, v1.xyz, v2.xyz (original vertex bits, untouched)
struct alignas(16) WideTri
{
    float v0x, v0y, v0z;
    int32_t prim;
    float v1x, v1y, v1z, pad1;
    float v2x, v2y, v2z, pad2;
};
static_assert(sizeof(WideTri) == 48, "WideTri must be 48 bytes");

struct Bvh
{
    const WideNode* nodes;
    const WideTri* tris;
};

struct Hit
{
    float t, u, v;
    int prim;  // -1: miss
};

// common/core.hpp:91-136: t from the plane, then the three signed sub-areas; u = area(p,v2,v0)/A,
// v = area(p,v0,v1)/A.  Accepts tmin <= t <= tmax (NaN fails).
CRT_HD bool ray_triangle(f3 ro, f3 rd, float tmin, float tmax, f3 v0, f3 v1, f3 v2, float& t_out, float& u_out,
                         float& v_out)
{
    const f3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    const f3 n = cross(e0, e1);
    const float t = dot(v0 - ro, n) / dot(n, rd);
    if (!(tmin <= t && t <= tmax)) return false;
    const f3 p = ro + rd * t;
    const float a0 = dot(n, cross(e0, p - v0));
    const float a1 = dot(n, cross(e1, p - v1));
    const float a2 = dot(n, cross(e2, p - v2));
    if (a0 < 0.0f || a1 < 0.0f || a2 < 0.0f) return false;
    const float a = a0 + a1 + a2;
    t_out = t;
    u_out = a2 / a;
    v_out = a0 / a;
    return true;
}

struct u4 { uint32_t x, y, z, w; };

CRT_HD u4 load_u4(const void* p)
{
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg((const uint4*)p);
    return {v.x, v.y, v.z, v.w};
#else
    u4 v;
    memcpy(&v, p, 16);
    return v;
#endif
}

// byte k of a 32-bit word as float.  Device: PRMT builds the bits of 2^23 + byte, one FADD removes 2^23.
CRT_HD float byte_to_float(uint32_t w, int k)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 | k)) - 8388608.0f;
#else
    return (float)((w >> (8 * k)) & 0xffu);
#endif
}

// Closest hit (ANY = false) in [tmin, tmax]: smallest t, ties -> larger primitive id.
// Any hit (ANY = true): returns at the first accepted triangle; only hit.prim >= 0 is meaningful.
template <bool ANY>
CRT_HD bool trace(const Bvh& bvh, f3 ro, f3 rd, float tmin, float tmax, Hit& hit)
{
    hit.prim = -1;
    hit.t = tmax;
    hit.u = hit.v = 0.0f;

    // reciprocal direction; an exactly axis-parallel component becomes +-1e-20 so that the slab
    // arithmetic never sees 0 * inf (the padded boxes make the perturbation harmless)
    const float dx = fabsf(rd.x) > 1e-20f ? rd.x : copysignf(1e-20f, rd.x);
    const float dy = fabsf(rd.y) > 1e-20f ? rd.y : copysignf(1e-20f, rd.y);
    const float dz = fabsf(rd.z) > 1e-20f ? rd.z : copysignf(1e-20f, rd.z);
    const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
    const bool nx = dx < 0.0f, ny = dy < 0.0f, nz = dz < 0.0f;
    const uint32_t octinv = 7u ^ ((nx ? 1u : 0u) | (ny ? 2u : 0u) | (nz ? 4u : 0u));

    uint32_t stack_base[kStackSize], stack_mask[kStackSize];
    int sp = 0;
    uint32_t node_idx = 0;
    uint32_t ng_base = 0, ng_mask = 0;  // current node group: inner-child hits in bits 24..31, imask in bits 0..7

    for (;;)
    {
        // ---- intersect the eight child boxes of node_idx
        const char* np = (const char*)(bvh.nodes + node_idx);
        const u4 n0 = load_u4(np), n1 = load_u4(np + 16), n2 = load_u4(np + 32), n3 = load_u4(np + 48),
                 n4 = load_u4(np + 64);
        const uint32_t e_imask = n0.w;
        const float sx = u2f((e_imask & 0xffu) << 23), sy = u2f(((e_imask >> 8) & 0xffu) << 23),
                    sz = u2f(((e_imask >> 16) & 0xffu) << 23);
        const uint32_t imask = e_imask >> 24;
        const float ax = sx * idx, ay = sy * idy, az = sz * idz;  // exact power-of-two scaling
        const float ox = (u2f(n0.x) - ro.x) * idx, oy = (u2f(n0.y) - ro.y) * idy, oz = (u2f(n0.z) - ro.z) * idz;
        // word layout: n2 = qlo.x[0..3] qlo.x[4..7] qlo.y[0..3] qlo.y[4..7]; n3 = qlo.z.. qhi.x..; n4 = qhi.y.. qhi.z..
        const uint32_t lox[2] = {n2.x, n2.y}, loy[2] = {n2.z, n2.w}, loz[2] = {n3.x, n3.y};
        const uint32_t hix[2] = {n3.z, n3.w}, hiy[2] = {n4.x, n4.y}, hiz[2] = {n4.z, n4.w};
        const uint32_t meta[2] = {n1.z, n1.w};
        uint32_t hits = 0;
#pragma unroll
        for (int s = 0; s < 8; s++)
        {
            const int w = s >> 2, k = s & 3;
            const uint32_t m = (meta[w] >> (8 * k)) & 0xffu;
            const float nearx = byte_to_float(nx ? hix[w] : lox[w], k), farx = byte_to_float(nx ? lox[w] : hix[w], k);
            const float neary = byte_to_float(ny ? hiy[w] : loy[w], k), fary = byte_to_float(ny ? loy[w] : hiy[w], k);
            const float nearz = byte_to_float(nz ? hiz[w] : loz[w], k), farz = byte_to_float(nz ? loz[w] : hiz[w], k);
            const float t0 = fmaxf(fmaxf(fmaf(nearx, ax, ox), fmaf(neary, ay, oy)), fmaxf(fmaf(nearz, az, oz), tmin));
            const float t1 = fminf(fminf(fmaf(farx, ax, ox), fmaf(fary, ay, oy)), fminf(fmaf(farz, az, oz), hit.t));
            if (m != 0 && t0 <= t1)
            {
                if ((imask >> s) & 1u) hits |= 1u << (24 + (s ^ octinv));
                else hits |= ((1u << (m >> 5)) - 1u) << (m & 31u);
            }
        }
        ng_base = n1.x;
        ng_mask = (hits & 0xff000000u) | imask;
        uint32_t tmask = hits & 0x00ffffffu;

        // ---- triangles of this node's leaf children that survived the box test
        const WideTri* tp = bvh.tris + n1.y;
        while (tmask)
        {
            const int i = 31 - clz32(tmask & (0u - tmask));
            tmask &= tmask - 1u;
            const char* q = (const char*)(tp + i);
            const u4 a = load_u4(q), b = load_u4(q + 16), c = load_u4(q + 32);
            float t, u, v;
            if (ray_triangle(ro, rd, tmin, hit.t, f3{u2f(a.x), u2f(a.y), u2f(a.z)}, f3{u2f(b.x), u2f(b.y), u2f(b.z)},
                             f3{u2f(c.x), u2f(c.y), u2f(c.z)}, t, u, v))
            {
                const int prim = (int)a.w;
                if (t < hit.t || hit.prim < 0 || prim > hit.prim)
                {
                    hit.t = t;
                    hit.u = u;
                    hit.v = v;
                    hit.prim = prim;
                    if (ANY) return true;
                }
            }
        }

        // ---- next node: nearest remaining child of the current group, else pop
        if ((ng_mask >> 24) == 0)
        {
            if (sp == 0) break;
            --sp;
            ng_base = stack_base[sp];
            ng_mask = stack_mask[sp];
        }
        const int bit = 31 - clz32(ng_mask);  // bits 24..31 are non-empty here
        ng_mask &= ~(1u << bit);
        const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
        node_idx = ng_base + (uint32_t)popc(ng_mask & 0xffu & ((1u << slot) - 1u));
        if (ng_mask >> 24)
        {
            stack_base[sp] = ng_base;
            stack_mask[sp] = ng_mask;
            ++sp;
        }
    }
    return hit.prim >= 0;
}
}  // namespace crt
