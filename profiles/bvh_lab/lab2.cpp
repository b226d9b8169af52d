// TEST / TUNING INFRASTRUCTURE — never part of the product.
// Second host-side lab: reads the ray dump of lab.cpp (LAB_DUMP) and measures, with the library's own traversal
// headers, what ENTRY HINTS would buy the any-hit walks: before the walk from the root, walk only the subtree k levels
// above the leaf that holds (a) the light triangle the ray aims at, (b) the triangle the ray starts on.  An any-hit
// answer does not depend on the visiting order, so a hit found there is the walk's answer, bit for bit.
//
//   g++ -std=c++17 -O2 -fopenmp -ffp-contract=off -DCRT_COUNT -I cedec-2024-rt_b200/csrc -o /tmp/lab/lab2 profiles/bvh_lab/lab2.cpp
//   /tmp/lab/lab2 /tmp/lab/blocks_restir_x6.tri /tmp/lab/rays_x6.bin
#define CRT_COUNT 1
#include "../../tests/emu/emu.cpp"

#include <chrono>
#include <map>

struct Ray
{
    f3 o, d;
    int cls, pix, own;
    int light = -1;
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

struct Counts
{
    double nodes = 0, tris = 0;
};

// any-hit walk of the subtree under `root_node` (0 = whole tree); `skip` = a node whose subtree was walked before
static bool walk_any(const Bvh& bvh, const Ray& r, bool far_first, uint32_t root_node, int64_t skip, Counts& c)
{
    RaySetup rsu = setup_ray(r.o, r.d, far_first);
    Walk w;
    WalkStack wst;
    walk_begin(w, rsu);
    w.ng_base = root_node;
    for (;;)
    {
        if ((w.ng_mask >> 24) == 0)
        {
            if (w.sp == 0) return false;
            --w.sp;
            w.ng_base = wst.e[w.sp].base;
            w.ng_mask = wst.e[w.sp].mask;
        }
        const int bit = 31 - clz32(w.ng_mask);
        w.ng_mask &= ~(1u << bit);
        const uint32_t slot = (uint32_t)(bit - 24) ^ rsu.octinv;
        const uint32_t node_idx = w.ng_base + (uint32_t)popc(w.ng_mask & 0xffu & ((1u << slot) - 1u));
        if (w.ng_mask >> 24) { wst.e[w.sp] = WalkEntry{w.ng_base, w.ng_mask}; ++w.sp; }
        if ((int64_t)node_idx == skip)
        {
            w.ng_mask = 0;
            continue;
        }
        uint32_t imask;
        const uint32_t hits = intersect_node(bvh, node_idx, rsu, 0.0f, 0.99f, w.ng_base, w.tri_base, imask);
        c.nodes += 1;
        w.ng_mask = (hits & 0xff000000u) | imask;
        uint32_t tm = hits & 0x00ffffffu;
        while (tm)
        {
            const int i = 31 - clz32(tm & (0u - tm));
            tm &= tm - 1u;
            Hit h; h.prim = -1; h.t = 0.99f; h.u = h.v = 0;
            c.tris += 1;
            if (intersect_wide_tri(bvh.tris + w.tri_base + i, rsu, 0.0f, h)) return true;
        }
    }
}

// BOTTOM-UP any-hit walk: starts at the node that holds the leaf of the triangle the ray starts on, walks that node's
// subtree, then climbs: parent, its other children's subtrees, grandparent ... up to the root.  Visits the nodes a walk from
// the root visits (every ancestor's box holds the origin), nearest first, without the descent before the first triangle.
static bool walk_bottom_up(const Bvh& bvh, const std::vector<uint32_t>& parent, uint32_t start, const Ray& r, bool far_first, Counts& c, int* climbs = nullptr)
{
    RaySetup rsu = setup_ray(r.o, r.d, far_first);
    uint32_t cur = start;
    int64_t skip = -1;
    for (;;)
    {
        Walk w;
        WalkStack wst;
        walk_begin(w, rsu);
        // node step on `cur`
        uint32_t imask;
        uint32_t hits = intersect_node(bvh, cur, rsu, 0.0f, 0.99f, w.ng_base, w.tri_base, imask);
        c.nodes += 1;
        if (skip >= 0)
        {
            const uint32_t rank = (uint32_t)skip - w.ng_base;
            uint32_t m = imask;
            for (uint32_t k = 0; k < rank; k++) m &= m - 1u;
            const uint32_t slot = (uint32_t)(31 - clz32(m & (0u - m)));
            hits &= ~(1u << (24 + (slot ^ rsu.octinv)));
        }
        w.ng_mask = (hits & 0xff000000u) | imask;
        uint32_t tm = hits & 0x00ffffffu;
        w.sp = 0;
        for (;;)
        {
            while (tm)
            {
                const int i = 31 - clz32(tm & (0u - tm));
                tm &= tm - 1u;
                Hit h; h.prim = -1; h.t = 0.99f; h.u = h.v = 0;
                c.tris += 1;
                if (intersect_wide_tri(bvh.tris + w.tri_base + i, rsu, 0.0f, h)) return true;
            }
            if ((w.ng_mask >> 24) == 0)
            {
                if (w.sp == 0) break;
                --w.sp;
                w.ng_base = wst.e[w.sp].base;
                w.ng_mask = wst.e[w.sp].mask;
            }
            const int bit = 31 - clz32(w.ng_mask);
            w.ng_mask &= ~(1u << bit);
            const uint32_t slot = (uint32_t)(bit - 24) ^ rsu.octinv;
            const uint32_t node_idx = w.ng_base + (uint32_t)popc(w.ng_mask & 0xffu & ((1u << slot) - 1u));
            if (w.ng_mask >> 24) { wst.e[w.sp] = WalkEntry{w.ng_base, w.ng_mask}; ++w.sp; }
            const uint32_t h2 = intersect_node(bvh, node_idx, rsu, 0.0f, 0.99f, w.ng_base, w.tri_base, imask);
            c.nodes += 1;
            w.ng_mask = (h2 & 0xff000000u) | imask;
            tm = h2 & 0x00ffffffu;
        }
        if (cur == 0) return false;
        skip = cur;
        cur = parent[cur];
        if (climbs) ++*climbs;
    }
}

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "/tmp/lab/blocks_restir_x6.tri";
    const char* dump = argc > 2 ? argv[2] : "/tmp/lab/rays_x6.bin";
    std::vector<char> file = read_file(path);
    const uint32_t n_tris = (uint32_t)(file.size() / 60);
    const float* t60 = (const float*)file.data();
    EmuGeom* g = build(t60, n_tris);
    const Bvh bvh = g->view();
    printf("scene %s: %u tris, %zu nodes, depth %d\n", path, n_tris, g->nodes.size(), g->depth);
    // parents and the node that holds each primitive's leaf
    const size_t nn = g->nodes.size();
    std::vector<uint32_t> parent(nn, 0), tri_node(n_tris, 0);
    for (size_t i = 0; i < nn; i++)
    {
        const WideNode& wn = g->nodes[i];
        int k = 0;
        for (int s = 0; s < 8; s++)
        {
            const uint32_t m = wn.meta[s];
            if (!m) continue;
            if ((m & 0x1f) >= 24) parent[wn.child_base + k++] = (uint32_t)i;
            else
                for (int j = 0; j < popc(m >> 5); j++) tri_node[g->tris[wn.tri_base + (m & 0x1f) + j].prim] = (uint32_t)i;
        }
    }
    auto ancestor = [&](uint32_t node, int k)
    {
        while (k-- > 0) node = parent[node];
        return node;
    };
    // subtree sizes (triangles) per level of ancestry, for the report
    std::vector<char> rd = read_file(dump);
    std::vector<Ray> rays, prim_recs;
    for (size_t off = 0; off + 36 <= rd.size(); off += 36)
    {
        const float* f = (const float*)(rd.data() + off);
        Ray r;
        r.o = f3{f[0], f[1], f[2]};
        r.d = f3{f[3], f[4], f[5]};
        r.cls = (int)f[6];
        r.pix = (int)f[7];
        memcpy(&r.own, f + 8, 4);
        if (r.cls == 1 || r.cls == 2) rays.push_back(r);
        else prim_recs.push_back(r);
    }
    printf("%zu shadow rays\n", rays.size());
    // the light triangle a ray aims at: the closest hit just beyond the segment's end
    size_t found = 0;
#pragma omp parallel for reduction(+ : found) schedule(dynamic, 256)
    for (size_t i = 0; i < rays.size(); i++)
    {
        Hit h;
        if (trace<false>(bvh, rays[i].o, rays[i].d, 0.99f, 1.02f, h) && has_emission(tri_at(t60, h.prim).emissive()))
        {
            rays[i].light = h.prim;
            found++;
        }
    }
    printf("light triangle identified for %.1f%% of the rays\n", 100.0 * found / rays.size());

    // primary rays: what a hint for the closest hit would buy.  (a) the walk as it is; (b) the triangle the same pixel hit
    // in the previous frame tested first (here: the true hit, i.e. a static camera), which only presets the bound t
    if (!prim_recs.empty())
    {
        const int W = argc > 3 ? atoi(argv[3]) : 480, H = argc > 4 ? atoi(argv[4]) : 270;
        const float eye_a[3] = {-0.579885f, 22.194597f, -6.567105f}, at_a[3] = {5.224952f, 20.847435f, 1.431192f}, up_a[3] = {0, 1, 0};
        crt_raygen rg;
        orc_lookat(eye_a, at_a, up_a, kPi / 4.0f, W, H, &rg);
        Counts a, b, c2;
        size_t n = 0, mism = 0;
        for (const Ray& r : prim_recs)
        {
            const int row = r.pix / W, xi = r.pix % W, yi = H - 1 - row;
            const Pix px = make_pix(xi, yi, W, H);
            f3 ro, rd;
            primary_ray(rg, px, W, H, ro, rd);
            Hit h;
            unsigned long long n0 = crt::g_count_nodes, t0 = crt::g_count_tris;
            trace<false>(bvh, ro, rd, 0.0f, kFltMax, h);
            a.nodes += crt::g_count_nodes - n0; a.tris += crt::g_count_tris - t0;
            if (h.prim != r.own) mism++;
            n++;
            // hinted: test the known triangle first, then walk with that bound
            Hit hh; hh.prim = -1; hh.t = kFltMax; hh.u = hh.v = 0;
            const RaySetup rsu = setup_ray(ro, rd);
            n0 = crt::g_count_nodes; t0 = crt::g_count_tris;
            if (r.own >= 0)
            {
                const TriRef t = tri_at(t60, r.own);
                float tt, u, v;
                if (ray_triangle(ro, rd, 0.0f, kFltMax, t.v(0), t.v(1), t.v(2), tt, u, v)) { hh.t = tt; hh.u = u; hh.v = v; hh.prim = r.own; }
            }
            Walk w; WalkStack wst;
            walk_begin(w, rsu);
            for (;;)
                if (walk_step<false, false>(bvh, w, wst, rsu, 0.0f, hh, 32) != kWalkContinue) break;
            b.nodes += crt::g_count_nodes - n0; b.tris += crt::g_count_tris - t0;
            if (hh.prim != h.prim || hh.t != h.t || hh.u != h.u || hh.v != h.v) mism++;
        }
        printf("== primary rays: %zu | walk: nodes %.2f tris %.2f | previous hit tested first: nodes %.2f tris %.2f (+1 test) | mismatches %zu\n", n,
               a.nodes / n, a.tris / n, b.nodes / n, b.tris / n, mism);
    }
    // config 4 (09_ris with the shadowed target function): 32 uniformly drawn light samples per path vertex, one shadow ray each.
    // Here: the camera vertices of the recorded frame, 8 candidates each (any random numbers serve the statistics)
    if (!prim_recs.empty() && getenv("LAB_UNIFORM"))
    {
        const int W = argc > 3 ? atoi(argv[3]) : 480, H = argc > 4 ? atoi(argv[4]) : 270;
        const float eye_a[3] = {-0.579885f, 22.194597f, -6.567105f}, at_a[3] = {5.224952f, 20.847435f, 1.431192f}, up_a[3] = {0, 1, 0};
        crt_raygen rg;
        orc_lookat(eye_a, at_a, up_a, kPi / 4.0f, W, H, &rg);
        std::vector<uint32_t> lights;
        for (uint32_t i = 0; i < n_tris; i++)
            if (has_emission(tri_at(t60, (int)i).emissive())) lights.push_back(i);
        const LightsIndexed L{t60, lights.data(), (uint32_t)lights.size()};
        std::vector<Ray> cand;
        for (size_t k = 0; k < prim_recs.size(); k += 7)
        {
            const Ray& r = prim_recs[k];
            if (r.own < 0 || has_emission(tri_at(t60, r.own).emissive())) continue;
            const int row = r.pix / W, xi = r.pix % W, yi = H - 1 - row;
            const Pix px = make_pix(xi, yi, W, H);
            f3 ro, rd;
            primary_ray(rg, px, W, H, ro, rd);
            Hit h;
            trace<false>(bvh, ro, rd, 0.0f, kFltMax, h);
            if (h.prim < 0) continue;
            const Surf surf = surface_from_hit(tri_at(t60, h.prim), ro, rd, h.t);
            Pcg rng(hash_pcg4(xi, yi, 1, 7), 0);
            for (int c = 0; c < 8; c++)
            {
                const float r0 = rng.next_f(), r1 = rng.next_f(), r2 = rng.next_f();
                const LightSample ls = L.sample(r0, r1, r2);
                Ray q;
                q.o = surf.p + 0.001f * surf.n;
                q.d = ls.p - surf.p;
                q.cls = 3; q.pix = r.pix; q.own = h.prim; q.light = -1;
                cand.push_back(q);
            }
        }
        size_t n = cand.size(), occ = 0, own_hit = 0;
        Counts far, near, bu, rest_far, rest_near;
        size_t nrest = 0;
        for (const Ray& r : cand)
        {
            Counts a, b, c;
            const bool h = walk_any(bvh, r, true, 0, -1, a);
            walk_any(bvh, r, false, 0, -1, b);
            far.nodes += a.nodes; far.tris += a.tris; near.nodes += b.nodes; near.tris += b.tris;
            occ += h;
            const TriRef t = tri_at(t60, r.own);
            float tt, u, v;
            if (ray_triangle(r.o, r.d, 0.0f, 0.99f, t.v(0), t.v(1), t.v(2), tt, u, v)) { own_hit++; continue; }
            nrest++;
            rest_far.nodes += a.nodes; rest_far.tris += a.tris; rest_near.nodes += b.nodes; rest_near.tris += b.tris;
            walk_bottom_up(bvh, parent, tri_node[r.own], r, false, c);
            bu.nodes += c.nodes; bu.tris += c.tris;
        }
        printf("== uniform candidates from the camera vertices: %zu rays, occluded %.1f%%, own triangle stops %.1f%%\n", n, 100.0 * occ / n, 100.0 * own_hit / n);
        printf("   all rays: far first nodes %.2f tris %.2f | near first nodes %.2f tris %.2f\n", far.nodes / n, far.tris / n, near.nodes / n, near.tris / n);
        printf("   rays left after the own-triangle test: far first %.2f / %.2f | near first %.2f / %.2f | bottom-up from the own node %.2f / %.2f\n",
               rest_far.nodes / nrest, rest_far.tris / nrest, rest_near.nodes / nrest, rest_near.tris / nrest, bu.nodes / nrest, bu.tris / nrest);
    }
    for (int cls = 1; cls <= 2; cls++)
    {
        const bool far_first = cls == 1;
        size_t n = 0, occ = 0;
        Counts base, base_occ, base_clear;
        std::vector<char> occluded(rays.size(), 0);
        for (size_t i = 0; i < rays.size(); i++)
        {
            if (rays[i].cls != cls) continue;
            n++;
            Counts c;
            const bool h = walk_any(bvh, rays[i], far_first, 0, -1, c);
            occluded[i] = h;
            base.nodes += c.nodes; base.tris += c.tris;
            if (h) { occ++; base_occ.nodes += c.nodes; base_occ.tris += c.tris; }
            else { base_clear.nodes += c.nodes; base_clear.tris += c.tris; }
        }
        printf("== %s rays: %zu, occluded %.1f%% | walk from the root: nodes %.2f tris %.2f (occluded %.2f / %.2f, clear %.2f / %.2f)\n",
               cls == 1 ? "visibility-reuse" : "resolve", n, 100.0 * occ / n, base.nodes / n, base.tris / n, base_occ.nodes / occ,
               base_occ.tris / occ, base_clear.nodes / (n - occ), base_clear.tris / (n - occ));
        // persistent-kernel drain: a simulation of k_trace_shadow_queue's scheduling (warps of 32 lanes fetch rays in queue order,
        // a lane takes a new ray when fewer than 24 lanes of its warp walk; one node step per iteration) with the rays in queue
        // order against the rays binned by segment length, longest first
        {
            std::vector<std::pair<float, unsigned>> q;  // (length, node steps)
            for (size_t i = 0; i < rays.size(); i++)
            {
                const Ray& r = rays[i];
                if (r.cls != cls) continue;
                if (cls == 1)
                {
                    const TriRef t = tri_at(t60, r.own);
                    float tt, u, v;
                    if (ray_triangle(r.o, r.d, 0.0f, 0.99f, t.v(0), t.v(1), t.v(2), tt, u, v)) continue;
                }
                Counts c;
                walk_any(bvh, r, far_first, 0, -1, c);
                q.push_back({sqrtf(dot(r.d, r.d)), (unsigned)c.nodes});
            }
            // length statistics
            {
                std::vector<float> len;
                for (auto& e : q) len.push_back(e.first);
                std::sort(len.begin(), len.end());
                double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
                for (auto& e : q) { const double x = log(e.first), y = e.second; sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y; }
                const double nq = (double)q.size(), cov = sxy / nq - sx / nq * sy / nq, vx = sxx / nq - sx / nq * sx / nq, vy = syy / nq - sy / nq * sy / nq;
                printf("   segment length: p10 %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f | correlation of log length with node steps %.2f\n", len[len.size() / 10],
                       len[len.size() / 2], len[len.size() * 9 / 10], len[len.size() * 99 / 100], len.back(), cov / sqrt(vx * vy));
            }
            auto simulate = [&](const std::vector<unsigned>& steps, int warps)
            {
                size_t next = 0;
                std::vector<std::vector<unsigned>> lane(warps, std::vector<unsigned>(32, 0));
                std::vector<char> done(warps, 0);
                unsigned long long iters = 0, busy_iters = 0, lane_steps = 0;
                int left = warps;
                unsigned long long span = 0;
                while (left)
                {
                    ++span;
                    for (int w = 0; w < warps; w++)
                    {
                        if (done[w]) continue;
                        int active = 0;
                        for (unsigned x : lane[w]) active += x > 0;
                        if (active < 24 && next < steps.size())
                            for (unsigned& x : lane[w])
                                if (x == 0 && next < steps.size()) x = steps[next++];
                        active = 0;
                        for (unsigned& x : lane[w])
                            if (x > 0) { --x; ++active; }
                        if (active == 0 && next >= steps.size()) { done[w] = 1; --left; continue; }
                        ++iters;
                        lane_steps += active;
                    }
                }
                (void)busy_iters;
                printf("      %d warps: %llu warp iterations, %.1f lanes per iteration, span %llu iterations (ideal %.0f)\n", warps, iters,
                       (double)lane_steps / iters, span, (double)lane_steps / (warps * 28.0));
            };
            std::vector<unsigned> in_order, binned;
            for (auto& e : q) in_order.push_back(e.second);
            float lo = 1e30f, hi = 0;
            for (auto& e : q) { lo = std::min(lo, e.first); hi = std::max(hi, e.first); }
            for (int nb : {4, 8})
            {
                std::vector<float> len;
                for (auto& e : q) len.push_back(e.first);
                std::sort(len.begin(), len.end());
                binned.clear();
                for (int b = nb - 1; b >= 0; b--)
                {
                    const float t0 = len[len.size() * b / nb], t1 = b == nb - 1 ? 1e30f : len[len.size() * (b + 1) / nb];
                    for (auto& e : q)
                        if (e.first >= t0 && e.first < t1) binned.push_back(e.second);
                }
                for (int warps : {64, 128, 256})
                {
                    printf("   %d length bins (quantiles), longest first:\n", nb);
                    simulate(binned, warps);
                    printf("   queue order:\n");
                    simulate(in_order, warps);
                }
            }
            // fixed bins: factor-2 classes of the segment length below a quarter of the scene diagonal (what the kernel can compute)
            {
                const float D = 470.0f;
                for (int nb : {4, 6, 8})
                {
                    binned.clear();
                    std::vector<size_t> cnt(nb, 0);
                    auto bin_of = [&](float len) { int b = 0; float th = D / 4; while (b < nb - 1 && len < th) { ++b; th *= 0.5f; } return b; };
                    for (int b = 0; b < nb; b++)
                        for (auto& e : q)
                            if (bin_of(e.first) == b) { binned.push_back(e.second); cnt[b]++; }
                    printf("   %d fixed factor-2 bins from D/4 down, longest first (", nb);
                    for (int b = 0; b < nb; b++) printf("%.0f%% ", 100.0 * cnt[b] / q.size());
                    printf("):\n");
                    for (int warps : {64, 128, 256, 512}) simulate(binned, warps);
                }
                for (float div : {8.0f, 16.0f, 32.0f, 64.0f})
                {
                    binned.clear();
                    size_t c0 = 0;
                    for (auto& e : q) if (e.first >= D / div) { binned.push_back(e.second); c0++; }
                    std::vector<unsigned> tail;
                    for (auto& e : q) if (e.first < D / div) tail.push_back(e.second);
                    binned.insert(binned.end(), tail.rbegin(), tail.rend());  // the short rays are stored from the buffer's end downwards
                    printf("   2 bins, long (>= D/%.0f: %.0f%%) first:\n", div, 100.0 * c0 / q.size());
                    for (int warps : {64, 256, 512}) simulate(binned, warps);
                }
                printf("   queue order:\n");
                simulate(in_order, 512);
            }
            // oracle: sorted by the true step count (the bound of any ordering)
            std::vector<unsigned> sorted = in_order;
            std::sort(sorted.rbegin(), sorted.rend());
            printf("   sorted by true node steps (bound):\n");
            for (int warps : {64, 128, 256}) simulate(sorted, warps);
        }
        // the cheapest hint: the ray's own triangle (the one its origin lies on), tested by the emitting pixel kernel
        {
            size_t caught1 = 0, caught2 = 0, below = 0, below_caught = 0;
            Counts rest, rest2;
            size_t nrest = 0, nrest2 = 0;
            for (size_t i = 0; i < rays.size(); i++)
            {
                const Ray& r = rays[i];
                if (r.cls != cls) continue;
                auto hits_prim = [&](int prim)
                {
                    if (prim < 0 || prim >= (int)n_tris) return false;
                    const TriRef t = tri_at(t60, prim);
                    float tt, u, v;
                    return ray_triangle(r.o, r.d, 0.0f, 0.99f, t.v(0), t.v(1), t.v(2), tt, u, v);
                };
                const bool h1 = hits_prim(r.own);
                const bool h2 = h1 || hits_prim(r.own ^ 1);
                // below the horizon of the own triangle's plane?
                const TriRef t = tri_at(t60, r.own);
                const f3 nrm = cross(t.v(1) - t.v(0), t.v(2) - t.v(0));
                const float side_o = dot(nrm, r.o - t.v(0)), side_e = dot(nrm, (r.o + r.d * 0.99f) - t.v(0));
                const bool crosses = (side_o > 0) != (side_e > 0);
                if (crosses) below++;
                if (crosses && h2) below_caught++;
                if (h1) caught1++;
                if (h2) caught2++;
                Counts c;
                if (!h1) { walk_any(bvh, r, far_first, 0, -1, c); rest.nodes += c.nodes; rest.tris += c.tris; nrest++; }
                if (!h2) { Counts c2; walk_any(bvh, r, far_first, 0, -1, c2); rest2.nodes += c2.nodes; rest2.tris += c2.tris; nrest2++; }
            }
            printf("   own triangle: hit by %.1f%% of all rays (own or own^1: %.1f%%); segment crosses the own plane for %.1f%%, of which %.1f%% caught\n",
                   100.0 * caught1 / n, 100.0 * caught2 / n, 100.0 * below / n, 100.0 * below_caught / (below ? below : 1));
            printf("   remaining rays after the own-triangle test: %.1f%% of the rays, nodes %.2f tris %.2f each -> work left: nodes %.1f%% tris %.1f%% of the base\n",
                   100.0 * nrest / n, rest.nodes / nrest, rest.tris / nrest, 100.0 * rest.nodes / base.nodes, 100.0 * rest.tris / base.tris);
            printf("   remaining after own and own^1: %.1f%% of the rays -> work left: nodes %.1f%% tris %.1f%%\n", 100.0 * nrest2 / n,
                   100.0 * rest2.nodes / base.nodes, 100.0 * rest2.tris / base.tris);
        }
        // bottom-up walk from the own triangle's node (after the own-triangle test)
        for (int order = 0; order < 2; order++)
        {
            Counts all, rest;
            size_t nrest = 0, wrong = 0;
            double climbs = 0;
            for (size_t i = 0; i < rays.size(); i++)
            {
                const Ray& r = rays[i];
                if (r.cls != cls) continue;
                const TriRef t = tri_at(t60, r.own);
                float tt, u, v;
                const bool own_hit = ray_triangle(r.o, r.d, 0.0f, 0.99f, t.v(0), t.v(1), t.v(2), tt, u, v);
                Counts c;
                int cl = 0;
                const bool h = walk_bottom_up(bvh, parent, tri_node[r.own], r, order == 1, c, &cl);
                if (h != (bool)occluded[i]) wrong++;
                all.nodes += c.nodes; all.tris += c.tris;
                if (!own_hit) { rest.nodes += c.nodes; rest.tris += c.tris; nrest++; climbs += cl; }
            }
            printf("   bottom-up walk (%s children first): all rays nodes %.2f tris %.2f | rays left after the own-triangle test: nodes %.2f tris %.2f, %.1f climbs (per ray of the class: nodes %.2f tris %.2f)%s\n",
                   order ? "far" : "near", all.nodes / n, all.tris / n, rest.nodes / nrest, rest.tris / nrest, climbs / nrest, rest.nodes / n, rest.tris / n,
                   wrong ? "  RESULT MISMATCH" : "");
        }
        // bottom-up walk from the LIGHT triangle's node (rays whose light was identified; the others walk from the root)
        for (int order = 0; order < 2; order++)
        {
            Counts rest, rest_base;
            size_t nrest = 0, wrong = 0;
            double climbs = 0;
            for (size_t i = 0; i < rays.size(); i++)
            {
                const Ray& r = rays[i];
                if (r.cls != cls) continue;
                const TriRef t = tri_at(t60, r.own);
                float tt, u, v;
                if (ray_triangle(r.o, r.d, 0.0f, 0.99f, t.v(0), t.v(1), t.v(2), tt, u, v)) continue;
                Counts c, cb;
                int cl = 0;
                const bool h = r.light >= 0 ? walk_bottom_up(bvh, parent, tri_node[r.light], r, order == 1, c, &cl) : walk_any(bvh, r, far_first, 0, -1, c);
                walk_any(bvh, r, far_first, 0, -1, cb);
                if (h != (bool)occluded[i]) wrong++;
                rest.nodes += c.nodes; rest.tris += c.tris; nrest++; climbs += cl;
                rest_base.nodes += cb.nodes; rest_base.tris += cb.tris;
            }
            printf("   bottom-up from the light's node (%s children first), rays left after the own-triangle test: nodes %.2f tris %.2f, %.1f climbs (from the root: %.2f / %.2f)%s\n",
                   order ? "far" : "near", rest.nodes / nrest, rest.tris / nrest, climbs / nrest, rest_base.nodes / nrest, rest_base.tris / nrest,
                   wrong ? "  RESULT MISMATCH" : "");
        }
        if (getenv("LAB_HINTS"))
        for (int side = 0; side < 3; side++)  // 0: light side, 1: origin side, 2: light side then origin side
            for (int k = 0; k <= 3; k++)
            {
                Counts hint, total;
                size_t caught = 0, hinted = 0, wrong = 0;
                for (size_t i = 0; i < rays.size(); i++)
                {
                    const Ray& r = rays[i];
                    if (r.cls != cls) continue;
                    bool hit = false;
                    int64_t skip = -1;
                    Counts c;
                    const int prims[2] = {side == 1 ? r.own : r.light, side == 2 ? r.own : -1};
                    bool any_hint = false;
                    for (int p = 0; p < 2 && !hit; p++)
                    {
                        if (prims[p] < 0) continue;
                        const uint32_t e = ancestor(tri_node[prims[p]], k);
                        if (e == 0 || (int64_t)e == skip) continue;  // the root itself: no hint
                        any_hint = true;
                        hit = walk_any(bvh, r, far_first, e, -1, c);
                        if (p == 0) skip = e;
                    }
                    if (any_hint) hinted++;
                    hint.nodes += c.nodes; hint.tris += c.tris;
                    if (hit) caught++;
                    else hit = walk_any(bvh, r, far_first, 0, side == 2 ? -1 : skip, c);
                    if (hit != (bool)occluded[i]) wrong++;
                    total.nodes += c.nodes; total.tris += c.tris;
                }
                printf("   hint %-22s k=%d: hinted %.1f%%, caught %.1f%% of the occluded | hint walk nodes %.2f tris %.2f | total nodes %.2f tris %.2f (base %.2f / %.2f)%s\n",
                       side == 0 ? "light side" : side == 1 ? "origin side" : "light then origin side", k, 100.0 * hinted / n,
                       100.0 * caught / (occ ? occ : 1), hint.nodes / n, hint.tris / n, total.nodes / n, total.tris / n, base.nodes / n,
                       base.tris / n, wrong ? "  RESULT MISMATCH" : "");
            }
    }
    return 0;
}
