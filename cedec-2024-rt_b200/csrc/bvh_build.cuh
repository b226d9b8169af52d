// bvh_build.cuh — the per-element steps of the GPU BVH build (replaces hiprtBuildGeometry,
// common/loader.hpp:68-112).  bvh_build.cu launches one thread per element for each step:
//
//   1 tri_bounds      triangle AABB + scene bounds (atomic min/max)
//   2 morton_key      63-bit Morton code of the centroid            -> radix sort (cub)
//   3 ploc_*          binary tree by parallel locally-ordered clustering (default), or
//     lbvh_node       Karras 2012 binary radix tree, one inner node per thread (CRT_BVH_BUILDER=lbvh)
//   4 lbvh_refit      bottom-up boxes (second arrival at a node continues) and, in the same sweep, the
//                     SAH-optimal way to cut every subtree into at most i = 1..7 wide-BVH children
//                     (dynamic programme of Ylitie et al. 2017, section 3.1)
//   5 collapse_item   top-down, level by level: a wide node takes the <= 8 children the programme chose;
//                     subtrees of <= 3 triangles may become leaf children; children are matched to octant
//                     slots, boxes are quantised outward to the node's 8-bit grid, triangles are written
//                     as 48-byte records next to each other
//
// The functions are host-callable too, so tests/emu can run the identical build sequentially on the
// CPU; the library itself only ever runs them on the device.
#pragma once
#include "bvh.cuh"

namespace crt
{
#if defined(__CUDA_ARCH__)
#define CRT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define CRT_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define CRT_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define CRT_FENCE() __threadfence()
#else
template <class T>
inline T crt_host_fetch_add(T* p, T v)
{
    const T old = *p;
    *p = old + v;
    return old;
}
#define CRT_ATOMIC_ADD(p, v) crt_host_fetch_add((p), (v))
#define CRT_ATOMIC_MIN(p, v) (*(p) = *(p) < (v) ? *(p) : (v))
#define CRT_ATOMIC_MAX(p, v) (*(p) > (v) ? *(p) : (*(p) = (v)))
#define CRT_FENCE()
#endif

// order-preserving float <-> uint mapping for atomic min/max on floats
CRT_HD uint32_t float_to_ordered(float f)
{
    const uint32_t u = f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
CRT_HD float ordered_to_float(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct BuildTri  // the nine vertex floats of a reference Triangle (60-byte stride, 4-byte aligned)
{
    f3 v0, v1, v2;
};
CRT_HD BuildTri load_build_tri(const float* tris60, uint32_t i)
{
    const float* p = tris60 + (size_t)i * 15;
    return {{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}};
}

struct Aabb
{
    f3 lo, hi;
};
CRT_HD float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
CRT_HD float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
CRT_HD Aabb tri_aabb(const BuildTri& t)
{
    return {{fmin3(t.v0.x, t.v1.x, t.v2.x), fmin3(t.v0.y, t.v1.y, t.v2.y), fmin3(t.v0.z, t.v1.z, t.v2.z)},
            {fmax3(t.v0.x, t.v1.x, t.v2.x), fmax3(t.v0.y, t.v1.y, t.v2.y), fmax3(t.v0.z, t.v1.z, t.v2.z)}};
}
CRT_HD Aabb aabb_union(const Aabb& a, const Aabb& b)
{
    return {{fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z)},
            {fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z)}};
}
CRT_HD float aabb_half_area(const Aabb& b)
{
    const float x = b.hi.x - b.lo.x, y = b.hi.y - b.lo.y, z = b.hi.z - b.lo.z;
    return x * y + y * z + z * x;
}

// ---- step 1: scene bounds; bounds6 = ordered-uint {lo.xyz, hi.xyz}
CRT_HD void tri_bounds(uint32_t i, const float* tris60, uint32_t* bounds6)
{
    const Aabb b = tri_aabb(load_build_tri(tris60, i));
    CRT_ATOMIC_MIN(bounds6 + 0, float_to_ordered(b.lo.x));
    CRT_ATOMIC_MIN(bounds6 + 1, float_to_ordered(b.lo.y));
    CRT_ATOMIC_MIN(bounds6 + 2, float_to_ordered(b.lo.z));
    CRT_ATOMIC_MAX(bounds6 + 3, float_to_ordered(b.hi.x));
    CRT_ATOMIC_MAX(bounds6 + 4, float_to_ordered(b.hi.y));
    CRT_ATOMIC_MAX(bounds6 + 5, float_to_ordered(b.hi.z));
}

// ---- step 2: 63-bit Morton key (21 bits per axis) of the box centre
CRT_HD uint64_t spread21(uint64_t x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
CRT_HD uint64_t morton_key(uint32_t i, const float* tris60, f3 scene_lo, f3 scene_inv_extent)
{
    const Aabb b = tri_aabb(load_build_tri(tris60, i));
    const float cx = ((b.lo.x + b.hi.x) * 0.5f - scene_lo.x) * scene_inv_extent.x;
    const float cy = ((b.lo.y + b.hi.y) * 0.5f - scene_lo.y) * scene_inv_extent.y;
    const float cz = ((b.lo.z + b.hi.z) * 0.5f - scene_lo.z) * scene_inv_extent.z;
    const float s = 2097151.0f;  // 2^21 - 1
    const uint64_t qx = (uint64_t)fminf(fmaxf(cx * s, 0.0f), s);
    const uint64_t qy = (uint64_t)fminf(fmaxf(cy * s, 0.0f), s);
    const uint64_t qz = (uint64_t)fminf(fmaxf(cz * s, 0.0f), s);
    return (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
}

// ---- step 3: binary radix tree.  Node ids: inner i in [0, n-2]; leaf j is (n-1) + j.
struct BinTree
{
    uint32_t n;         // triangles (= leaves)
    uint32_t* left;     // [n-1]
    uint32_t* right;    // [n-1]
    uint32_t* parent;   // [2n-1]
    uint32_t* first;    // [n-1]  first sorted position covered by inner node
    uint32_t* count;    // [n-1]  triangles under inner node
    float* box;         // [2n-1][6]  padded AABB
    uint32_t* visits;   // [n-1]  refit arrival counters (zeroed)
    float* cost;        // [2n-1][7]  c(node, i): least SAH cost of the subtree as a forest of <= i wide-BVH children
    uint8_t* split;     // [2n-1][8]  [0]: 1 = c(node,1) is a leaf; [j-1], j = 2..8: triangles... see sah_plan_node
};

CRT_HD uint32_t bin_tri_count(const BinTree& bt, uint32_t id) { return id >= bt.n - 1 ? 1u : bt.count[id]; }

// SAH constants: one node step (8 child boxes) vs one triangle test, as in Ylitie et al. 2017
#ifndef CRT_COST_TRI
#define CRT_COST_TRI 0.3f
#endif
constexpr float kCostNode = 1.0f, kCostTri = CRT_COST_TRI, kCostInf = 1.0e30f;

// c(n,1) = min(leaf, internal);  internal = distribute(n,8) + area * kCostNode
// c(n,i) = min(distribute(n,i), c(n,i-1)), i = 2..7;  distribute(n,j) = min_k c(left,k) + c(right,j-k)
// split[j-1] (j = 2..8) = the k of the best distribution, or 0 when c(n,j) = c(n,j-1) ("use fewer roots")
CRT_HD void sah_plan_leaf(const BinTree& bt, uint32_t id, float half_area)
{
    float* c = bt.cost + (size_t)id * 7;
    uint8_t* sp = bt.split + (size_t)id * 8;
    for (int i = 0; i < 7; i++) c[i] = half_area * kCostTri;
    sp[0] = 1;
    for (int j = 1; j < 8; j++) sp[j] = 0;
}
CRT_HD void sah_plan_node(const BinTree& bt, uint32_t id, uint32_t l, uint32_t r, float half_area, uint32_t tri_count)
{
    float cl[8], cr[8];
    for (int i = 1; i <= 7; i++)
    {
#if defined(__CUDA_ARCH__)
        // a child's table may have been written by another thread: read it through L2
        cl[i] = __ldcg(bt.cost + (size_t)l * 7 + i - 1);
        cr[i] = __ldcg(bt.cost + (size_t)r * 7 + i - 1);
#else
        cl[i] = bt.cost[(size_t)l * 7 + i - 1];
        cr[i] = bt.cost[(size_t)r * 7 + i - 1];
#endif
    }
    float dist[9];
    uint8_t arg[9];
    for (int j = 2; j <= 8; j++)
    {
        float best = kCostInf * 4.0f;
        int bk = 1;
        for (int k = 1; k < j; k++)
        {
            if (k > 7 || j - k > 7) continue;
            const float v = cl[k] + cr[j - k];
            if (v < best)
            {
                best = v;
                bk = k;
            }
        }
        dist[j] = best;
        arg[j] = (uint8_t)bk;
    }
    float* c = bt.cost + (size_t)id * 7;
    uint8_t* sp = bt.split + (size_t)id * 8;
    const float c_internal = dist[8] + half_area * kCostNode;
    const float c_leaf = tri_count <= (uint32_t)kLeafMaxTris ? half_area * (float)tri_count * kCostTri : kCostInf;
    c[0] = c_leaf <= c_internal ? c_leaf : c_internal;
    sp[0] = c_leaf <= c_internal ? 1 : 0;
    sp[7] = arg[8];
    for (int i = 2; i <= 7; i++)
    {
        if (dist[i] < c[i - 2])
        {
            c[i - 1] = dist[i];
            sp[i - 1] = arg[i];
        }
        else
        {
            c[i - 1] = c[i - 2];
            sp[i - 1] = 0;
        }
    }
}

CRT_HD int lbvh_delta(const uint64_t* keys, uint32_t n, int i, int j)
{
    if (j < 0 || j >= (int)n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);  // duplicate keys: fall back to the position
    return clz64(a ^ b);
}

CRT_HD void lbvh_node(uint32_t idx, const uint64_t* keys, const BinTree& bt)
{
    const uint32_t n = bt.n;
    const int i = (int)idx;
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2)
    {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t l_id = (lo == gamma) ? (n - 1) + (uint32_t)gamma : (uint32_t)gamma;
    const uint32_t r_id = (hi == gamma + 1) ? (n - 1) + (uint32_t)(gamma + 1) : (uint32_t)(gamma + 1);
    bt.left[idx] = l_id;
    bt.right[idx] = r_id;
    bt.parent[l_id] = idx;
    bt.parent[r_id] = idx;
    bt.first[idx] = (uint32_t)lo;
    bt.count[idx] = (uint32_t)(hi - lo + 1);
    if (idx == 0) bt.parent[0] = 0xffffffffu;
}

// ---- step 4: bottom-up boxes.  One thread per leaf; `pad` is added on every side of every leaf box.
CRT_HD void store_box(float* box, uint32_t id, const Aabb& b)
{
    float* p = box + (size_t)id * 6;
    p[0] = b.lo.x; p[1] = b.lo.y; p[2] = b.lo.z;
    p[3] = b.hi.x; p[4] = b.hi.y; p[5] = b.hi.z;
}
CRT_HD Aabb load_box(const float* box, uint32_t id)
{
    const float* p = box + (size_t)id * 6;
    return {{p[0], p[1], p[2]}, {p[3], p[4], p[5]}};
}
CRT_HD void lbvh_refit(uint32_t leaf, const float* tris60, const uint32_t* sorted_idx, float pad, const BinTree& bt)
{
    Aabb b = tri_aabb(load_build_tri(tris60, sorted_idx[leaf]));
    b.lo = b.lo - f3{pad, pad, pad};
    b.hi = b.hi + f3{pad, pad, pad};
    uint32_t id = (bt.n - 1) + leaf;
    store_box(bt.box, id, b);
    sah_plan_leaf(bt, id, aabb_half_area(b));
    if (bt.n == 1) return;
    uint32_t p = bt.parent[id];
    while (p != 0xffffffffu)
    {
        CRT_FENCE();
        if (CRT_ATOMIC_ADD(bt.visits + p, 1u) == 0u) return;  // first arrival: the sibling will finish
        CRT_FENCE();
#if defined(__CUDA_ARCH__)
        // the sibling's box was written by another thread: read it through L2
        const volatile float* vb = bt.box;
        const uint32_t l = bt.left[p], r = bt.right[p];
        Aabb lb, rb;
        lb.lo = {vb[(size_t)l * 6 + 0], vb[(size_t)l * 6 + 1], vb[(size_t)l * 6 + 2]};
        lb.hi = {vb[(size_t)l * 6 + 3], vb[(size_t)l * 6 + 4], vb[(size_t)l * 6 + 5]};
        rb.lo = {vb[(size_t)r * 6 + 0], vb[(size_t)r * 6 + 1], vb[(size_t)r * 6 + 2]};
        rb.hi = {vb[(size_t)r * 6 + 3], vb[(size_t)r * 6 + 4], vb[(size_t)r * 6 + 5]};
        b = aabb_union(lb, rb);
#else
        b = aabb_union(load_box(bt.box, bt.left[p]), load_box(bt.box, bt.right[p]));
#endif
        store_box(bt.box, p, b);
        sah_plan_node(bt, p, bt.left[p], bt.right[p], aabb_half_area(b), bt.count[p]);
        p = bt.parent[p];
    }
}

// ---- steps 3'/4' (default builder): PLOC — parallel locally-ordered clustering (Meister & Bittner 2018).
// Clusters start as the Morton-sorted triangles.  Each round every cluster looks kPlocRadius positions to either
// side for the neighbour whose union box has the smallest area; mutual nearest neighbours merge into a new inner
// node (box, triangle count and SAH plan are final at that moment), the array is compacted, repeat until one
// cluster is left.  Agglomerative clustering follows surface area, so the topology is much closer to a SAH
// build than the Morton-prefix splits of the LBVH; the binary tree feeds the same collapse.
constexpr int kPlocRadius = 16;

struct PlocRound
{
    uint32_t m;              // clusters this round
    const uint32_t* node_in; // [m] binary node id of each cluster
    const float* box_in;     // [m][8] lo.xyz, -, hi.xyz, -
    uint32_t* nn;            // [m] nearest neighbour position
    unsigned long long* flag;// [m] low word: cluster survives (1/0); high word: it leads a merge (1/0)
    uint32_t* node_out;      // [<= m]
    float* box_out;
};

CRT_HD void ploc_init_leaf(uint32_t j, const float* tris60, const uint32_t* sorted_idx, float pad, const BinTree& bt,
                           uint32_t* node_out, float* box_out)
{
    Aabb b = tri_aabb(load_build_tri(tris60, sorted_idx[j]));
    b.lo = b.lo - f3{pad, pad, pad};
    b.hi = b.hi + f3{pad, pad, pad};
    const uint32_t id = (bt.n - 1) + j;
    store_box(bt.box, id, b);
    sah_plan_leaf(bt, id, aabb_half_area(b));
    node_out[j] = id;
    float* q = box_out + (size_t)j * 8;
    q[0] = b.lo.x; q[1] = b.lo.y; q[2] = b.lo.z; q[3] = 0.0f;
    q[4] = b.hi.x; q[5] = b.hi.y; q[6] = b.hi.z; q[7] = 0.0f;
}
CRT_HD Aabb ploc_box(const float* box, uint32_t i)
{
    const float* q = box + (size_t)i * 8;
    return {{q[0], q[1], q[2]}, {q[4], q[5], q[6]}};
}
// nearest neighbour by (union area, i ^ j): the key is symmetric in (i, j), so the globally best pair is always
// mutual (every round merges at least one pair) and runs of identical boxes pair up (i, i^1) instead of chaining
CRT_HD void ploc_nn(uint32_t i, const PlocRound& p)
{
    const Aabb bi = ploc_box(p.box_in, i);
    const uint32_t lo = i > (uint32_t)kPlocRadius ? i - kPlocRadius : 0u;
    const uint32_t hi = i + kPlocRadius < p.m - 1 ? i + kPlocRadius : p.m - 1;
    float best = 3.0e38f;
    uint32_t best_j = i, best_x = 0xffffffffu;
    for (uint32_t j = lo; j <= hi; j++)
    {
        if (j == i) continue;
        const float a = aabb_half_area(aabb_union(bi, ploc_box(p.box_in, j)));
        const uint32_t x = i ^ j;
        if (a < best || (a == best && x < best_x))
        {
            best = a;
            best_j = j;
            best_x = x;
        }
    }
    p.nn[i] = best_j;
}
CRT_HD void ploc_flag(uint32_t i, const PlocRound& p)
{
    const uint32_t j = p.nn[i];
    const bool mutual = j != i && p.nn[j] == i;
    const unsigned long long survive = (mutual && j < i) ? 0ull : 1ull;
    const unsigned long long lead = (mutual && i < j) ? 1ull : 0ull;
    p.flag[i] = survive | (lead << 32);
}
// scan[i] = exclusive prefix sum of flag; id_top = id of the first node created this round (ids count down to 0)
CRT_HD void ploc_apply(uint32_t i, const PlocRound& p, const unsigned long long* scan, uint32_t id_top, const BinTree& bt)
{
    const unsigned long long f = p.flag[i];
    if (!(f & 1ull)) return;
    const uint32_t pos = (uint32_t)(scan[i] & 0xffffffffull);
    uint32_t node = p.node_in[i];
    Aabb b = ploc_box(p.box_in, i);
    if (f >> 32)
    {
        const uint32_t j = p.nn[i];
        const uint32_t id = id_top - (uint32_t)(scan[i] >> 32);
        const uint32_t l = node, r = p.node_in[j];
        b = aabb_union(b, ploc_box(p.box_in, j));
        bt.left[id] = l;
        bt.right[id] = r;
        const uint32_t cnt = bin_tri_count(bt, l) + bin_tri_count(bt, r);
        bt.count[id] = cnt;
        store_box(bt.box, id, b);
        sah_plan_node(bt, id, l, r, aabb_half_area(b), cnt);
        node = id;
    }
    p.node_out[pos] = node;
    float* q = p.box_out + (size_t)pos * 8;
    q[0] = b.lo.x; q[1] = b.lo.y; q[2] = b.lo.z; q[3] = 0.0f;
    q[4] = b.hi.x; q[5] = b.hi.y; q[6] = b.hi.z; q[7] = 0.0f;
}

// ---- step 5: collapse to the wide tree
struct CollapseItem
{
    uint32_t bnode;  // binary node to expand
    uint32_t wnode;  // wide node to write
};
struct WideOut
{
    WideNode* nodes;
    WideTri* tris;
    uint32_t* node_count;   // allocator, starts at 1 (root)
    uint32_t* tri_count;    // allocator, starts at 0
    CollapseItem* next;     // next level's work
    uint32_t* next_count;
};

// sorted positions of the (<= 3) leaves under a binary node; PLOC subtrees are not contiguous ranges
CRT_HD int bin_leaves(const BinTree& bt, uint32_t id, uint32_t out[kLeafMaxTris])
{
    uint32_t st[kLeafMaxTris + 1];
    int sp = 0, n = 0;
    st[sp++] = id;
    while (sp)
    {
        const uint32_t x = st[--sp];
        if (x >= bt.n - 1) out[n++] = x - (bt.n - 1);
        else
        {
            st[sp++] = bt.right[x];
            st[sp++] = bt.left[x];
        }
    }
    return n;
}

CRT_HD uint8_t quant_exponent(float extent)
{
    // smallest e with extent / 2^e <= 250, as a biased exponent; extent >= 0.  The grid starts one cell below
    // the node box and child boxes get kQuantMargin cells of slack, so 250 cells of extent use q = 0..252.
    if (!(extent > 0.0f)) return 1;
    int e = (int)((f2u(extent) >> 23) & 0xffu) - 127 - 8;  // 2^e ~ extent / 256 .. extent / 512
    if (e < -126) e = -126;
    while (extent / u2f((uint32_t)(e + 127) << 23) > 250.0f) ++e;
    return (uint8_t)(e + 127);
}

// the node's quantisation grid from its box: power-of-two cell per axis, origin one cell below the box so that the
// lower margin never needs clamping
struct NodeGrid
{
    float lo[3], inv_cell[3];
};
CRT_HD NodeGrid set_node_grid(WideNode& wn, const Aabb& nb)
{
    wn.ex = quant_exponent(nb.hi.x - nb.lo.x);
    wn.ey = quant_exponent(nb.hi.y - nb.lo.y);
    wn.ez = quant_exponent(nb.hi.z - nb.lo.z);
    const float cell[3] = {u2f((uint32_t)wn.ex << 23), u2f((uint32_t)wn.ey << 23), u2f((uint32_t)wn.ez << 23)};
    wn.px = nb.lo.x - cell[0];
    wn.py = nb.lo.y - cell[1];
    wn.pz = nb.lo.z - cell[2];
    return NodeGrid{{wn.px, wn.py, wn.pz}, {1.0f / cell[0], 1.0f / cell[1], 1.0f / cell[2]}};
}
// child box -> the slot's six bytes: outward rounding plus kQuantMargin cells of slack (see bvh.cuh: byte_to_unit_float)
CRT_HD void quantise_slot(WideNode& wn, int s, const Aabb& box, const NodeGrid& g)
{
    const float clo[3] = {box.lo.x, box.lo.y, box.lo.z}, chi[3] = {box.hi.x, box.hi.y, box.hi.z};
    for (int a = 0; a < 3; a++)
    {
        const float ql = floorf((clo[a] - g.lo[a]) * g.inv_cell[a] - kQuantMargin);
        const float qh = ceilf((chi[a] - g.lo[a]) * g.inv_cell[a] + kQuantMargin);
        wn.qlo[a][s] = (uint8_t)fminf(fmaxf(ql, 0.0f), 255.0f);
        wn.qhi[a][s] = (uint8_t)fminf(fmaxf(qh, 0.0f), 255.0f);
    }
}

CRT_HD void collapse_item(const CollapseItem it, const float* tris60, const uint32_t* sorted_idx, const BinTree& bt,
                          const WideOut& out)
{
    // the <= 8 children the SAH plan chose for this subtree (sah_plan_node): walk the recorded splits
    uint32_t ch[8];
    bool ch_leaf[8];
    int cnt = 0;
    if (bt.split[(size_t)it.bnode * 8] == 1)
    {
        ch[0] = it.bnode;  // the whole tree is one leaf (n <= 3)
        ch_leaf[0] = true;
        cnt = 1;
    }
    else
    {
        uint32_t st_node[16];
        int st_j[16], sp = 0;
        const int k8 = bt.split[(size_t)it.bnode * 8 + 7];
        st_node[sp] = bt.right[it.bnode];
        st_j[sp++] = 8 - k8;
        st_node[sp] = bt.left[it.bnode];
        st_j[sp++] = k8;
        while (sp)
        {
            --sp;
            const uint32_t m = st_node[sp];
            int j = st_j[sp];
            const uint8_t* ms = bt.split + (size_t)m * 8;
            while (j > 1 && ms[j - 1] == 0) --j;  // c(m,j) = c(m,j-1): fewer roots are at least as good
            if (j == 1)
            {
                ch[cnt] = m;
                ch_leaf[cnt] = ms[0] == 1;
                ++cnt;
            }
            else
            {
                const int k = ms[j - 1];
                st_node[sp] = bt.right[m];
                st_j[sp++] = j - k;
                st_node[sp] = bt.left[m];
                st_j[sp++] = k;
            }
        }
    }

    // node box = union of the (already padded) child boxes
    Aabb cb[8];
    Aabb nb = load_box(bt.box, ch[0]);
    cb[0] = nb;
    for (int k = 1; k < cnt; k++)
    {
        cb[k] = load_box(bt.box, ch[k]);
        nb = aabb_union(nb, cb[k]);
    }

    // octant slots: slot bit a set = child lies on the + side of axis a.  Greedy maximum of
    // sum_a (+-)(centre_child - centre_node)_a over the free (child, slot) pairs.
    int slot_of[8];
    {
        const f3 nc = (nb.lo + nb.hi) * 0.5f;
        float cost[8][8];
        for (int k = 0; k < cnt; k++)
        {
            const f3 d = (cb[k].lo + cb[k].hi) * 0.5f - nc;
            for (int s = 0; s < 8; s++)
                cost[k][s] = ((s & 1) ? d.x : -d.x) + ((s & 2) ? d.y : -d.y) + ((s & 4) ? d.z : -d.z);
        }
        uint32_t child_free = (1u << cnt) - 1u, slot_free = 0xffu;
        for (int round = 0; round < cnt; round++)
        {
            int bk = -1, bs = -1;
            float bc = -3.0e38f;
            for (int k = 0; k < cnt; k++)
            {
                if (!((child_free >> k) & 1u)) continue;
                for (int s = 0; s < 8; s++)
                    if (((slot_free >> s) & 1u) && cost[k][s] > bc)
                    {
                        bc = cost[k][s];
                        bk = k;
                        bs = s;
                    }
            }
            slot_of[bk] = bs;
            child_free &= ~(1u << bk);
            slot_free &= ~(1u << bs);
        }
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    for (int k = 0; k < cnt; k++) child_in_slot[slot_of[k]] = k;

    WideNode wn;
    const NodeGrid grid = set_node_grid(wn, nb);

    uint32_t n_inner = 0, n_leaf_tris = 0;
    uint8_t imask = 0;
    for (int s = 0; s < 8; s++)
    {
        const int k = child_in_slot[s];
        if (k < 0) continue;
        if (!ch_leaf[k])
        {
            imask |= (uint8_t)(1u << s);
            n_inner++;
        }
        else n_leaf_tris += bin_tri_count(bt, ch[k]);
    }
    const uint32_t child_base = n_inner ? CRT_ATOMIC_ADD(out.node_count, n_inner) : 0u;
    const uint32_t tri_base = n_leaf_tris ? CRT_ATOMIC_ADD(out.tri_count, n_leaf_tris) : 0u;
    const uint32_t next_base = n_inner ? CRT_ATOMIC_ADD(out.next_count, n_inner) : 0u;
    wn.imask = imask;
    wn.child_base = child_base;
    wn.tri_base = tri_base;

    uint32_t inner_rank = 0, tri_off = 0;
    for (int s = 0; s < 8; s++)
    {
        const int k = child_in_slot[s];
        if (k < 0)
        {
            wn.meta[s] = 0;
            for (int a = 0; a < 3; a++)
            {
                wn.qlo[a][s] = 255;
                wn.qhi[a][s] = 0;
            }
            continue;
        }
        quantise_slot(wn, s, cb[k], grid);
        const uint32_t tc = bin_tri_count(bt, ch[k]);
        if (!ch_leaf[k])
        {
            wn.meta[s] = (uint8_t)(0x20u | (24u + (uint32_t)s));
            out.next[next_base + inner_rank] = CollapseItem{ch[k], child_base + inner_rank};
            inner_rank++;
        }
        else
        {
            wn.meta[s] = (uint8_t)((((1u << tc) - 1u) << 5) | tri_off);  // unary count
            uint32_t leaf_pos[kLeafMaxTris];
            bin_leaves(bt, ch[k], leaf_pos);
            for (uint32_t j = 0; j < tc; j++)
            {
                const uint32_t prim = sorted_idx[leaf_pos[j]];
                const BuildTri t = load_build_tri(tris60, prim);
                WideTri wt;
                wt.v0x = t.v0.x; wt.v0y = t.v0.y; wt.v0z = t.v0.z; wt.prim = (int32_t)prim;
                wt.v1x = t.v1.x; wt.v1y = t.v1.y; wt.v1z = t.v1.z; wt.pad1 = 0.0f;
                wt.v2x = t.v2.x; wt.v2y = t.v2.y; wt.v2z = t.v2.z; wt.pad2 = 0.0f;
                out.tris[tri_base + tri_off + j] = wt;
            }
            tri_off += tc;
        }
    }
    out.nodes[it.wnode] = wn;
}

// ---- refit (crt_refit_geometry; HIPRT's hiprtBuildOperationUpdate, hiprt_types.h:131-135): the vertices moved, the
// topology stays.  Triangle records reload their vertices; nodes recompute their grids and child bytes bottom-up, one
// launch per level (the collapse allocated node indices level by level, so a level is an index range).  Slot
// assignment, child indices and triangle ranges are kept: traversal order may be less front-to-back than a fresh
// build's, results are the same (the tree only culls).
CRT_HD void refit_tri(uint32_t i, const float* tris60, WideTri* tris)
{
    WideTri wt = tris[i];
    const BuildTri t = load_build_tri(tris60, (uint32_t)wt.prim);
    wt.v0x = t.v0.x; wt.v0y = t.v0.y; wt.v0z = t.v0.z;
    wt.v1x = t.v1.x; wt.v1y = t.v1.y; wt.v1z = t.v1.z;
    wt.v2x = t.v2.x; wt.v2y = t.v2.y; wt.v2z = t.v2.z;
    tris[i] = wt;
}
// node_box: [n_nodes][6] exact (padded) boxes; the children's entries are final when their parent's level runs
CRT_HD void refit_node(uint32_t node, WideNode* nodes, const WideTri* tris, float* node_box, float pad)
{
    WideNode wn = nodes[node];
    Aabb cb[8], nb;
    bool any = false;
    for (int s = 0; s < 8; s++)
    {
        const uint32_t meta = wn.meta[s];
        if (meta == 0) continue;
        Aabb b;
        if ((meta & 0x1fu) >= 24u)  // inner child
            b = load_box(node_box, wn.child_base + (uint32_t)popc((uint32_t)wn.imask & ((1u << s) - 1u)));
        else
        {
            const uint32_t count = (uint32_t)popc(meta >> 5), first = wn.tri_base + (meta & 0x1fu);
            for (uint32_t j = 0; j < count; j++)
            {
                const WideTri& w = tris[first + j];
                const Aabb tb = tri_aabb(BuildTri{{w.v0x, w.v0y, w.v0z}, {w.v1x, w.v1y, w.v1z}, {w.v2x, w.v2y, w.v2z}});
                b = j ? aabb_union(b, tb) : tb;
            }
            b.lo = b.lo - f3{pad, pad, pad};
            b.hi = b.hi + f3{pad, pad, pad};
        }
        cb[s] = b;
        nb = any ? aabb_union(nb, b) : b;
        any = true;
    }
    if (!any) return;  // the empty tree's root
    const NodeGrid grid = set_node_grid(wn, nb);
    for (int s = 0; s < 8; s++)
        if (wn.meta[s] != 0) quantise_slot(wn, s, cb[s], grid);
    nodes[node] = wn;
    store_box(node_box, node, nb);
}
}  // namespace crt
