// TEST INFRASTRUCTURE — not part of the product path.
//
// Conservative CPU BVH used by both oracles (oracle/_ref and oracle/port) in
// place of HIPRT 2.4.6b6daf9, whose traversal lives in a closed binary
// (/root/reference/libs/hiprt/hiprt/linux64/libhiprt0200464.so; call sites
// common/raytrace.hpp:25-35, examples/06_ao_hiprt/06_ao_hiprt.cu:20-28).
//
// The tree only *culls*; every accept/reject decision and every t/u/v bit comes
// from the leaf functor, which is the reference's intersect_ray_triangle
// (common/core.hpp:91-136).  Because the boxes are padded and the slab test has
// slack, the result equals the reference's brute-force loop
// (examples/04_ao/04_ao.cu:14-24): smallest t, and on an exact t tie the
// larger triangle index wins (the loop shrinks tmax to t and accepts t<=tmax,
// so a later index overwrites an equal-t hit).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <vector>

namespace cpubvh
{
struct Node
{
    float lo[3], hi[3];
    int left;   // internal: index of left child (right = left + 1); leaf: first prim slot
    int count;  // 0 = internal, else number of prims
};

struct Bvh
{
    std::vector<Node> nodes;
    std::vector<int> prims;  // permutation of triangle indices
    int n_tris = 0;
    std::atomic<int> next{1};  // node allocator (nodes are pre-sized, tasks build in parallel)
};

struct BuildItem
{
    float c[3];
    float lo[3], hi[3];
    int idx;
};

inline void build_rec(Bvh& bvh, std::vector<BuildItem>& items, int node, int first, int last, int depth)
{
    Node& nd = bvh.nodes[node];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = first; i < last; i++)
        for (int a = 0; a < 3; a++)
        {
            lo[a] = std::min(lo[a], items[i].lo[a]);
            hi[a] = std::max(hi[a], items[i].hi[a]);
            clo[a] = std::min(clo[a], items[i].c[a]);
            chi[a] = std::max(chi[a], items[i].c[a]);
        }
    for (int a = 0; a < 3; a++)
    {
        // pad: absorbs the rounding of the reference's plane/area arithmetic
        nd.lo[a] = lo[a] - 1e-4f * std::max(1.0f, std::fabs(lo[a]));
        nd.hi[a] = hi[a] + 1e-4f * std::max(1.0f, std::fabs(hi[a]));
    }
    const int n = last - first;
    if (n <= 4)
    {
        nd.left = first;
        nd.count = n;
        return;
    }
    int axis = 0;
    for (int a = 1; a < 3; a++)
        if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
    const int mid = first + n / 2;
    std::nth_element(items.begin() + first, items.begin() + mid, items.begin() + last,
                     [axis](const BuildItem& x, const BuildItem& y) { return x.c[axis] < y.c[axis]; });
    const int left = bvh.next.fetch_add(2);
    nd.left = left;
    nd.count = 0;
    if (depth < 6)
    {
#pragma omp task shared(bvh, items)
        build_rec(bvh, items, left, first, mid, depth + 1);
#pragma omp task shared(bvh, items)
        build_rec(bvh, items, left + 1, mid, last, depth + 1);
#pragma omp taskwait
    }
    else
    {
        build_rec(bvh, items, left, first, mid, depth + 1);
        build_rec(bvh, items, left + 1, mid, last, depth + 1);
    }
}

// verts: 9 floats per triangle, `stride_floats` floats between triangles.
inline Bvh* build(const float* verts, int stride_floats, int n)
{
    Bvh* bvh = new Bvh;
    bvh->n_tris = n;
    std::vector<BuildItem> items(n);
    for (int i = 0; i < n; i++)
    {
        const float* v = verts + (size_t)i * stride_floats;
        BuildItem& it = items[i];
        it.idx = i;
        for (int a = 0; a < 3; a++)
        {
            it.lo[a] = std::min(v[a], std::min(v[3 + a], v[6 + a]));
            it.hi[a] = std::max(v[a], std::max(v[3 + a], v[6 + a]));
            it.c[a] = (v[a] + v[3 + a] + v[6 + a]) * (1.0f / 3.0f);
        }
    }
    bvh->nodes.resize(2 * (size_t)n + 2);
    if (n > 0)
    {
#pragma omp parallel
#pragma omp single
        build_rec(*bvh, items, 0, 0, n, 0);
    }
    bvh->nodes.resize(bvh->next.load());
    bvh->prims.resize(n);
    for (int i = 0; i < n; i++) bvh->prims[i] = items[i].idx;
    return bvh;
}

struct Hit
{
    int prim = -1;
    float t = 0.0f, u = 0.0f, v = 0.0f;
};

inline bool slab(const Node& nd, const float o[3], const float inv[3], float tmin, float tmax, float& tnear)
{
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++)
    {
        float ta = (nd.lo[a] - o[a]) * inv[a];
        float tb = (nd.hi[a] - o[a]) * inv[a];
        if (ta != ta || tb != tb) continue;  // 0 * inf: origin on the slab plane, axis-parallel ray
        if (ta > tb) std::swap(ta, tb);
        ta = ta - std::fabs(ta) * 4e-6f - 1e-6f;
        tb = tb + std::fabs(tb) * 4e-6f + 1e-6f;
        t0 = std::max(t0, ta);
        t1 = std::min(t1, tb);
    }
    tnear = t0;
    return t0 <= t1;
}

// Leaf: bool(int prim, float tmin, float tmax, float& t, float& u, float& v)  -- the reference test.
// any_hit=true returns at the first accepted triangle (boolean-equivalent for
// check_visibility, common/raytrace.hpp:45-52).
template <class Leaf>
inline Hit trace(const Bvh& bvh, const float o[3], const float d[3], float tmin, float tmax, Leaf&& leaf,
                 bool any_hit = false)
{
    Hit best;
    if (bvh.n_tris == 0) return best;
    float inv[3];
    for (int a = 0; a < 3; a++) inv[a] = 1.0f / d[a];
    float best_t = tmax;
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    float tn;
    while (sp)
    {
        const Node& nd = bvh.nodes[stack[--sp]];
        if (!slab(nd, o, inv, tmin, best_t, tn)) continue;
        if (nd.count)
        {
            for (int k = 0; k < nd.count; k++)
            {
                const int prim = bvh.prims[nd.left + k];
                float t, u, v;
                if (leaf(prim, tmin, best_t, t, u, v))
                {
                    // leaf() accepts t <= best_t; tie rule: larger index wins
                    if (t < best_t || best.prim < 0 || prim > best.prim)
                    {
                        best.prim = prim;
                        best.t = t;
                        best.u = u;
                        best.v = v;
                        best_t = t;
                        if (any_hit) return best;
                    }
                }
            }
        }
        else
        {
            float tl, tr;
            const bool hl = slab(bvh.nodes[nd.left], o, inv, tmin, best_t, tl);
            const bool hr = slab(bvh.nodes[nd.left + 1], o, inv, tmin, best_t, tr);
            if (hl && hr)
            {
                if (tl <= tr)
                {
                    stack[sp++] = nd.left + 1;
                    stack[sp++] = nd.left;
                }
                else
                {
                    stack[sp++] = nd.left;
                    stack[sp++] = nd.left + 1;
                }
            }
            else if (hl) stack[sp++] = nd.left;
            else if (hr) stack[sp++] = nd.left + 1;
        }
    }
    return best;
}
}  // namespace cpubvh
