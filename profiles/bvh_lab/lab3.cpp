// TEST / TUNING INFRASTRUCTURE — never part of the product.
// Third host-side lab: how much would a better binary tree buy?  Builds the binary BVH top-down with a binned
// surface-area heuristic (the classic high-quality CPU build) instead of PLOC, feeds it through the library's own plan /
// collapse steps (bvh_build.cuh: lbvh_refit, collapse_item) and walks the recorded rays of lab.cpp through both wide trees.
//
//   g++ -std=c++17 -O2 -fopenmp -ffp-contract=off -DCRT_COUNT -I cedec-2024-rt_b200/csrc -o /tmp/lab/lab3 profiles/bvh_lab/lab3.cpp
//   /tmp/lab/lab3 /tmp/lab/blocks_restir_x6.tri /tmp/lab/rays_x6.bin 480 270
#define CRT_COUNT 1
#include "../../tests/emu/emu.cpp"

#include <chrono>

struct Ray
{
    f3 o, d;
    int cls, pix, own;
};
static std::vector<char> read_file(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> b(n);
    if (fread(b.data(), 1, n, f) != (size_t)n) exit(1);
    fclose(f);
    return b;
}

struct SahBuilder
{
    const float* t60;
    uint32_t n;
    std::vector<Aabb> box;
    std::vector<f3> cen;
    std::vector<uint32_t> order;
    uint32_t *left, *right, *parent, *count;
    int bins = 32;

    static float comp(const f3& v, int a) { return a == 0 ? v.x : a == 1 ? v.y : v.z; }
    // returns the node id of the subtree over order[lo, hi); inner ids are [id, id + hi - lo - 1)
    uint32_t build(uint32_t lo, uint32_t hi, uint32_t id, uint32_t par, int depth)
    {
        if (hi - lo == 1)
        {
            const uint32_t leaf = (n - 1) + lo;
            parent[leaf] = par;
            return leaf;
        }
        parent[id] = par;
        count[id] = hi - lo;
        Aabb cb{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}};
        for (uint32_t i = lo; i < hi; i++)
        {
            const f3 c = cen[order[i]];
            cb.lo = {fminf(cb.lo.x, c.x), fminf(cb.lo.y, c.y), fminf(cb.lo.z, c.z)};
            cb.hi = {fmaxf(cb.hi.x, c.x), fmaxf(cb.hi.y, c.y), fmaxf(cb.hi.z, c.z)};
        }
        uint32_t mid = lo;
        float best = 1e30f;
        int best_axis = -1;
        float best_pos = 0;
        const uint32_t cnt = hi - lo;
        if (cnt <= 16)
        {
            // full sweep on every axis
            std::vector<uint32_t> tmp(order.begin() + lo, order.begin() + hi), best_order;
            std::vector<float> right_area(cnt);
            for (int a = 0; a < 3; a++)
            {
                std::sort(tmp.begin(), tmp.end(), [&](uint32_t x, uint32_t y) { return comp(cen[x], a) < comp(cen[y], a); });
                Aabb acc = box[tmp[cnt - 1]];
                for (int i = (int)cnt - 1; i >= 1; i--)
                {
                    acc = aabb_union(acc, box[tmp[i]]);
                    right_area[i] = aabb_half_area(acc);
                }
                acc = box[tmp[0]];
                for (uint32_t i = 1; i < cnt; i++)
                {
                    const float c = aabb_half_area(acc) * i + right_area[i] * (cnt - i);
                    if (c < best) { best = c; best_axis = a; mid = lo + i; best_order = tmp; }
                    acc = aabb_union(acc, box[tmp[i]]);
                }
            }
            if (best_axis >= 0) std::copy(best_order.begin(), best_order.end(), order.begin() + lo);
        }
        else
        {
            for (int a = 0; a < 3; a++)
            {
                const float c0 = comp(cb.lo, a), c1 = comp(cb.hi, a);
                if (!(c1 > c0)) continue;
                const float scale = bins / (c1 - c0);
                std::vector<Aabb> bb(bins, Aabb{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}});
                std::vector<uint32_t> bc(bins, 0);
                for (uint32_t i = lo; i < hi; i++)
                {
                    int b = (int)((comp(cen[order[i]], a) - c0) * scale);
                    b = b < 0 ? 0 : b >= bins ? bins - 1 : b;
                    bb[b] = aabb_union(bb[b], box[order[i]]);
                    bc[b]++;
                }
                std::vector<float> ra(bins, 0);
                std::vector<uint32_t> rc(bins, 0);
                Aabb acc{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}};
                uint32_t c = 0;
                for (int b = bins - 1; b >= 1; b--)
                {
                    if (bc[b]) acc = aabb_union(acc, bb[b]);
                    c += bc[b];
                    ra[b] = c ? aabb_half_area(acc) : 0;
                    rc[b] = c;
                }
                acc = Aabb{{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}};
                c = 0;
                for (int b = 1; b < bins; b++)
                {
                    if (bc[b - 1]) acc = aabb_union(acc, bb[b - 1]);
                    c += bc[b - 1];
                    if (c == 0 || rc[b] == 0) continue;
                    const float cost = aabb_half_area(acc) * c + ra[b] * rc[b];
                    if (cost < best) { best = cost; best_axis = a; best_pos = c0 + b / scale; }
                }
            }
            if (best_axis >= 0)
            {
                auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t x) { return comp(cen[x], best_axis) < best_pos; });
                mid = (uint32_t)(it - order.begin());
            }
        }
        if (mid <= lo || mid >= hi) mid = lo + cnt / 2;  // coincident centroids: any split
        uint32_t l, r;
        const uint32_t lid = id + 1, rid = id + (mid - lo);
        if (cnt > 4096 && depth < 12)
        {
#pragma omp task shared(l)
            l = build(lo, mid, lid, id, depth + 1);
#pragma omp task shared(r)
            r = build(mid, hi, rid, id, depth + 1);
#pragma omp taskwait
        }
        else
        {
            l = build(lo, mid, lid, id, depth + 1);
            r = build(mid, hi, rid, id, depth + 1);
        }
        left[id] = l;
        right[id] = r;
        return id;
    }
};

// the emu's build with the binary tree from the SAH builder (everything after the topology is the library's code)
static EmuGeom* build_sah(const float* tris60, uint32_t n)
{
    EmuGeom* g = new EmuGeom;
    g->tris60 = tris60;
    uint32_t b6[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (uint32_t i = 0; i < n; i++) tri_bounds(i, tris60, b6);
    float max_abs = 0;
    for (int a = 0; a < 6; a++) max_abs = fmaxf(max_abs, fabsf(ordered_to_float(b6[a])));
    g->pad = 64.0f * 5.9604645e-8f * fmaxf(max_abs, 1.0f);
    const uint32_t ni = n - 1;
    std::vector<uint32_t> left(ni + 1), right(ni + 1), parent(2 * (size_t)n - 1), first(ni + 1), count(ni + 1), visits(ni + 1, 0u);
    std::vector<float> box((2 * (size_t)n - 1) * 6), cost((2 * (size_t)n - 1) * 7);
    std::vector<uint8_t> split((2 * (size_t)n - 1) * 8);
    BinTree bt{n, left.data(), right.data(), parent.data(), first.data(), count.data(), box.data(), visits.data(), cost.data(), split.data()};
    SahBuilder sb;
    sb.t60 = tris60;
    sb.n = n;
    sb.box.resize(n);
    sb.cen.resize(n);
    sb.order.resize(n);
    for (uint32_t i = 0; i < n; i++)
    {
        sb.box[i] = tri_aabb(load_build_tri(tris60, i));
        sb.cen[i] = (sb.box[i].lo + sb.box[i].hi) * 0.5f;
        sb.order[i] = i;
    }
    sb.left = left.data(); sb.right = right.data(); sb.parent = parent.data(); sb.count = count.data();
    if (const char* e = getenv("LAB_BINS")) sb.bins = atoi(e);
#pragma omp parallel
#pragma omp single
    sb.build(0, n, 0, 0xffffffffu, 0);
    const std::vector<uint32_t>& idx = sb.order;
    for (uint32_t i = 0; i < n; i++) lbvh_refit(i, tris60, idx.data(), g->pad, bt);
    g->nodes.resize((size_t)n + 1);
    g->tris.resize(n);
    std::vector<CollapseItem> q0((size_t)n + 1), q1((size_t)n + 1);
    uint32_t counters[3] = {1u, 0u, 0u};
    WideOut out{g->nodes.data(), g->tris.data(), &counters[0], &counters[1], nullptr, &counters[2]};
    q0[0] = CollapseItem{0u, 0u};
    uint32_t n_items = 1;
    CollapseItem *qi = q0.data(), *qo = q1.data();
    g->level_begin.assign(1, 0u);
    while (n_items)
    {
        g->depth++;
        g->level_begin.push_back(g->level_begin.back() + n_items);
        out.next = qo;
        counters[2] = 0;
        for (uint32_t i = 0; i < n_items; i++) collapse_item(qi[i], tris60, idx.data(), bt, out);
        n_items = counters[2];
        std::swap(qi, qo);
    }
    if (counters[1] != n) fprintf(stderr, "build_sah: %u triangle records for %u triangles\n", counters[1], n);
    g->nodes.resize(counters[0]);
    return g;
}

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "/tmp/lab/blocks_restir_x6.tri";
    const char* dump = argc > 2 ? argv[2] : "/tmp/lab/rays_x6.bin";
    const int W = argc > 3 ? atoi(argv[3]) : 480, H = argc > 4 ? atoi(argv[4]) : 270;
    std::vector<char> file = read_file(path);
    const uint32_t n_tris = (uint32_t)(file.size() / 60);
    const float* t60 = (const float*)file.data();
    std::vector<char> rd = read_file(dump);
    std::vector<Ray> rays;
    for (size_t off = 0; off + 36 <= rd.size(); off += 36)
    {
        const float* f = (const float*)(rd.data() + off);
        Ray r;
        r.o = f3{f[0], f[1], f[2]};
        r.d = f3{f[3], f[4], f[5]};
        r.cls = (int)f[6];
        r.pix = (int)f[7];
        memcpy(&r.own, f + 8, 4);
        rays.push_back(r);
    }
    const float eye_a[3] = {-0.579885f, 22.194597f, -6.567105f}, at_a[3] = {5.224952f, 20.847435f, 1.431192f}, up_a[3] = {0, 1, 0};
    crt_raygen rg;
    orc_lookat(eye_a, at_a, up_a, kPi / 4.0f, W, H, &rg);
    for (int which = 0; which < 2; which++)
    {
        auto t0 = std::chrono::steady_clock::now();
        EmuGeom* g = which ? build_sah(t60, n_tris) : build(t60, n_tris);
        auto t1 = std::chrono::steady_clock::now();
        const Bvh bvh = g->view();
        size_t leaf[4] = {0, 0, 0, 0}, inner = 0;
        for (const WideNode& wn : g->nodes)
            for (int s = 0; s < 8; s++)
            {
                const uint32_t m = wn.meta[s];
                if (!m) continue;
                if ((m & 0x1f) >= 24) inner++;
                else leaf[popc(m >> 5)]++;
            }
        printf("== %s: %zu wide nodes, depth %d, leaves 1/2/3: %zu %zu %zu, build %.1f s\n", which ? "binned SAH, top-down" : "PLOC", g->nodes.size(), g->depth,
               leaf[1], leaf[2], leaf[3], std::chrono::duration<double>(t1 - t0).count());
        double n[3] = {0, 0, 0}, nodes[3] = {0, 0, 0}, tris[3] = {0, 0, 0};
        size_t mism = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : mism)
        for (size_t i = 0; i < rays.size(); i++)
        {
            const Ray& r = rays[i];
            const unsigned long long n0 = crt::g_count_nodes, t0c = crt::g_count_tris;
            Hit h;
            if (r.cls == 0)
            {
                const int row = r.pix / W, xi = r.pix % W, yi = H - 1 - row;
                f3 ro, rdir;
                primary_ray(rg, make_pix(xi, yi, W, H), W, H, ro, rdir);
                trace<false>(bvh, ro, rdir, 0.0f, kFltMax, h);
                if (h.prim != r.own) mism++;
            }
            else if (r.cls == 1) trace<true, true>(bvh, r.o, r.d, 0.0f, 0.99f, h);
            else trace<true, false>(bvh, r.o, r.d, 0.0f, 0.99f, h);
            const double dn = (double)(crt::g_count_nodes - n0), dt = (double)(crt::g_count_tris - t0c);
#pragma omp critical
            {
                n[r.cls] += 1; nodes[r.cls] += dn; tris[r.cls] += dt;
            }
        }
        const char* names[3] = {"primary", "visibility reuse", "resolve"};
        for (int c = 0; c < 3; c++) printf("   %-18s %9.0f rays: nodes %.2f tris %.2f\n", names[c], n[c], nodes[c] / n[c], tris[c] / n[c]);
        printf("   primary-hit mismatches against the recorded frame: %zu\n", mism);
        delete g;
    }
    return 0;
}
