// TEST / TUNING INFRASTRUCTURE — never part of the product.
// Host-side lab for the traversal work of the fused ReSTIR DI frame: builds the wide BVH with the very same
// headers the library uses (through tests/emu/emu.cpp, -DCRT_COUNT), renders a few frames of config 4/5's camera
// at a reduced size, records the three ray populations of the last frames (primary, visibility reuse, resolve)
// and reports node steps / triangle tests per ray, occlusion rates and what an occluder cache would catch.
//
//   g++ -std=c++17 -O2 -fopenmp -ffp-contract=off -DCRT_COUNT -I cedec-2024-rt_b200/csrc -o /tmp/lab/lab profiles/bvh_lab/lab.cpp
//   /tmp/lab/lab /tmp/lab/blocks_restir.tri 960 540 4
#define CRT_COUNT 1
#include "../../tests/emu/emu.cpp"

#include <chrono>
#include <map>

struct RayRec
{
    f3 o, d;
    int pix;
    int occl_tri;  // index into the BVH triangle records of the occluder found (-1: unoccluded)
    unsigned nodes, tris;
    int flag = 0;  // resolve rays: the final sample's stored visibility bit
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

// any-hit walk that also reports which triangle record stopped it
static int any_hit_record(const Bvh& bvh, f3 ro, f3 rd, unsigned& nodes, unsigned& tris)
{
    const unsigned long long n0 = crt::g_count_nodes, t0 = crt::g_count_tris;
    Hit h;
    const bool hit = trace<true>(bvh, ro, rd, 0.0f, 0.99f, h);
    nodes = (unsigned)(crt::g_count_nodes - n0);
    tris = (unsigned)(crt::g_count_tris - t0);
    return hit ? h.prim : -1;
}

static void stats(const char* name, const std::vector<RayRec>& rays)
{
    if (rays.empty()) { printf("%-18s none\n", name); return; }
    double sn = 0, st = 0, sno = 0, sto = 0, snu = 0, stu = 0;
    size_t occ = 0;
    for (const RayRec& r : rays)
    {
        sn += r.nodes; st += r.tris;
        if (r.occl_tri >= 0) { occ++; sno += r.nodes; sto += r.tris; }
        else { snu += r.nodes; stu += r.tris; }
    }
    const size_t n = rays.size(), un = n - occ;
    printf("%-18s %9zu rays  nodes %.2f tris %.2f | occluded %.1f%%: nodes %.2f tris %.2f | clear: nodes %.2f tris %.2f\n", name, n,
           sn / n, st / n, 100.0 * occ / n, occ ? sno / occ : 0.0, occ ? sto / occ : 0.0, un ? snu / un : 0.0, un ? stu / un : 0.0);
}

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "/tmp/lab/blocks_restir.tri";
    const int W = argc > 2 ? atoi(argv[2]) : 960, H = argc > 3 ? atoi(argv[3]) : 540;
    const int frames = argc > 4 ? atoi(argv[4]) : 4;
    std::vector<char> file = read_file(path);
    const uint32_t n_tris = (uint32_t)(file.size() / 60);
    const float* t60 = (const float*)file.data();
    auto t_a = std::chrono::steady_clock::now();
    EmuGeom* g = build(t60, n_tris);
    auto t_b = std::chrono::steady_clock::now();
    printf("scene %s: %u tris, %zu wide nodes, depth %d, build %.1f s\n", path, n_tris, g->nodes.size(), g->depth,
           std::chrono::duration<double>(t_b - t_a).count());
    // node statistics: children per node, leaf sizes
    {
        size_t inner = 0, leaf[4] = {0, 0, 0, 0}, empty = 0;
        for (const WideNode& wn : g->nodes)
            for (int s = 0; s < 8; s++)
            {
                const uint32_t m = wn.meta[s];
                if (m == 0) empty++;
                else if ((m & 0x1f) >= 24) inner++;
                else leaf[popc(m >> 5)]++;
            }
        printf("slots: inner %zu, leaf1 %zu leaf2 %zu leaf3 %zu, empty %zu (%.2f children per node)\n", inner, leaf[1], leaf[2],
               leaf[3], empty, (double)(inner + leaf[1] + leaf[2] + leaf[3]) / g->nodes.size());
    }
    std::vector<uint32_t> lights;
    for (uint32_t i = 0; i < n_tris; i++)
        if (has_emission(tri_at(t60, (int)i).emissive())) lights.push_back(i);
    printf("lights %zu\n", lights.size());

    const float eye_a[3] = {-0.579885f, 22.194597f, -6.567105f}, at_a[3] = {5.224952f, 20.847435f, 1.431192f}, up_a[3] = {0, 1, 0};
    crt_raygen rg;
    orc_lookat(eye_a, at_a, up_a, kPi / 4.0f, W, H, &rg);
    crt_options options;
    memset(&options, 0, sizeof options);
    options.accumulate = 1;
    options.max_depth = 6;
    options.ris_sample_count = 32;
    options.use_temporal_resampling = 1;
    options.use_spatial_resampling = 1;
    options.spatial_resampling_sample_count = 5;
    options.spatial_resampling_radius = 30.0f;
    options.spatial_resampling_passes = 3;
    options.use_visibility_reuse = 1;
    const Opt opt = make_opt(options);
    const f3 eye = v3(eye_a);
    const size_t n = (size_t)W * H;
    std::vector<crt_visibility> vis(n);
    std::vector<char> T(n * 76, 0), A(n * 76, 0), B(n * 76, 0), g0(n * 16), g1(n * 8);
    std::vector<uint8_t> cls(n, 0);
    std::vector<crt_float4> accum(n, crt_float4{0, 0, 0, 0});
    const GBuf gb{g0.data(), g1.data(), cls.data()};
    const SoaStore sT{T.data(), n}, sA{A.data(), n}, sB{B.data(), n};
    const LightsIndexed L{t60, lights.data(), (uint32_t)lights.size()};
    const Bvh bvh = g->view();
    g_math_mode = 0;

    std::vector<RayRec> prim, vr, rs, vr_prev, rs_prev;
    for (int frame = 1; frame <= frames; frame++)
    {
        const bool rec = frame >= frames - 1;
        vr_prev.swap(vr);
        rs_prev.swap(rs);
        prim.assign(rec ? n : 0, RayRec{});
        vr.assign(n, RayRec{{0, 0, 0}, {0, 0, 0}, -1, -1, 0, 0});
        rs.assign(n, RayRec{{0, 0, 0}, {0, 0, 0}, -1, -1, 0, 0});
        launch(W, H, [&](Pix p)
               {
                   const unsigned long long n0 = crt::g_count_nodes, t0 = crt::g_count_tris;
                   px_raycast(p, W, H, bvh, rg, vis.data());
                   if (rec)
                   {
                       prim[p.idx].pix = p.idx;
                       prim[p.idx].nodes = (unsigned)(crt::g_count_nodes - n0);
                       prim[p.idx].tris = (unsigned)(crt::g_count_tris - t0);
                       prim[p.idx].occl_tri = vis[p.idx].index;
                   }
               });
        launch(W, H, [&](Pix p)
               {
                   const CandPixel cp = classify_pixel(p, t60, vis.data());
                   const DeferredRay d = px_candidate_temporal<Math<0>>(p, cp, frame, bvh, t60, eye, L, opt, sT, gb);
                   if (d.want)
                   {
                       RayRec& r = vr[p.idx];
                       r.o = d.org; r.d = d.dir; r.pix = p.idx;
                       r.occl_tri = any_hit_record(bvh, d.org, d.dir, r.nodes, r.tris);
                       if (r.occl_tri < 0) *sT.mword(p.idx) |= kVisBit;
                   }
               });
        SoaStore in = sT, out = sA;
        for (int pass = 0; pass < options.spatial_resampling_passes; pass++)
        {
            if (pass == 1) { in = sA; out = sB; }
            else if (pass > 1) std::swap(in, out);
            launch(W, H, [&](Pix p) { px_spatial_fast<Math<0>>(p, W, H, frame, pass, bvh, eye, opt, in, out, gb); });
        }
        const SoaStore fin = out;
        launch(W, H, [&](Pix p)
               {
                   DeferredShade sh{{0, 0, 0}, {0, 0, 0}, 0.0f};
                   const DeferredRay d = px_resolve_fast(p, accum.data(), t60, vis.data(), fin, gb, sh, true, true);
                   if (!d.want) return;
                   RayRec& r = rs[p.idx];
                   r.o = d.org; r.d = d.dir; r.pix = p.idx;
                   r.flag = (int)(fin.load(p.idx).s.vis & 1u);
                   r.occl_tri = any_hit_record(bvh, d.org, d.dir, r.nodes, r.tris);
                   const float V = r.occl_tri >= 0 ? 0.0f : 1.0f;
                   write_accum(accum.data(), p.idx, sh.bg * V * sh.rad * sh.ucw, true);
               });
        printf("frame %d done\n", frame);
    }
    auto compact = [](std::vector<RayRec>& v)
    {
        std::vector<RayRec> o;
        for (const RayRec& r : v)
            if (r.pix >= 0) o.push_back(r);
        v.swap(o);
    };
    std::vector<RayRec> vr_full = vr, rs_full = rs, vrp_full = vr_prev, rsp_full = rs_prev;
    compact(vr); compact(rs);
    {
        // primary: occl_tri is the primitive id here
        double sn = 0, st = 0;
        for (const RayRec& r : prim) { sn += r.nodes; st += r.tris; }
        printf("%-18s %9zu rays  nodes %.2f tris %.2f\n", "primary", prim.size(), sn / prim.size(), st / prim.size());
    }
    stats("visibility reuse", vr);
    stats("resolve", rs);

    // occluder cache: a ray of the last frame first tests the triangle that stopped (a) the same pixel's ray of the same
    // class in the previous frame, (b) the same pixel's ray of the other class in this/previous frame
    auto tri_hits = [&](int prim_id, const RayRec& r)
    {
        if (prim_id < 0) return false;
        const TriRef t = tri_at(t60, prim_id);
        float tt, u, v;
        return ray_triangle(r.o, r.d, 0.0f, 0.99f, t.v(0), t.v(1), t.v(2), tt, u, v);
    };
    auto cache_report = [&](const char* name, const std::vector<RayRec>& cur, const std::vector<RayRec>& a, const std::vector<RayRec>* b)
    {
        size_t rays = 0, occ = 0, hit_a = 0, hit_ab = 0;
        double saved_nodes = 0, saved_tris = 0, all_nodes = 0, all_tris = 0;
        for (size_t i = 0; i < cur.size(); i++)
        {
            const RayRec& r = cur[i];
            if (r.pix < 0) continue;
            rays++;
            all_nodes += r.nodes; all_tris += r.tris;
            if (r.occl_tri < 0) continue;
            occ++;
            const bool ha = a[i].pix >= 0 && tri_hits(a[i].occl_tri, r);
            const bool hb = b && (*b)[i].pix >= 0 && tri_hits((*b)[i].occl_tri, r);
            if (ha) hit_a++;
            if (ha || hb) { hit_ab++; saved_nodes += r.nodes; saved_tris += r.tris; }
        }
        printf("cache %-26s rays %zu occluded %zu | hit by A %.1f%% of occluded, by A|B %.1f%% | walk work saved: nodes %.1f%% tris %.1f%%\n",
               name, rays, occ, 100.0 * hit_a / (occ ? occ : 1), 100.0 * hit_ab / (occ ? occ : 1), 100.0 * saved_nodes / all_nodes,
               100.0 * saved_tris / all_tris);
    };
    if (frames >= 2)
    {
        cache_report("vr <- prev vr (+prev rs)", vr_full, vrp_full, &rsp_full);
        cache_report("rs <- prev rs (+this vr)", rs_full, rsp_full, &vr_full);
    }
    // ray identity across frames: a resolve (visibility-reuse) ray of the last frame whose origin and direction are, bit for
    // bit, those of the same pixel's ray of the same class one frame earlier has that ray's answer
    if (frames >= 2)
    {
        auto same = [](const RayRec& a, const RayRec& b) { return a.pix >= 0 && b.pix >= 0 && memcmp(&a.o, &b.o, 12) == 0 && memcmp(&a.d, &b.d, 12) == 0; };
        size_t n_rs = 0, id_rs = 0, n_vr = 0, id_vr = 0, id_rs_vr = 0;
        double w_all = 0, w_same = 0;
        for (size_t i = 0; i < rs_full.size(); i++)
        {
            if (rs_full[i].pix >= 0)
            {
                n_rs++;
                w_all += rs_full[i].nodes;
                if (same(rs_full[i], rsp_full[i])) { id_rs++; w_same += rs_full[i].nodes; }
                else if (same(rs_full[i], vrp_full[i])) id_rs_vr++;
            }
            if (vr_full[i].pix >= 0) { n_vr++; if (same(vr_full[i], vrp_full[i])) id_vr++; }
        }
        printf("ray identity: resolve rays equal to the previous frame's resolve ray of the pixel %.1f%% (%.1f%% of their node steps), to its visibility-reuse ray %.1f%%; visibility-reuse rays equal to the previous one %.1f%%\n",
               100.0 * id_rs / (n_rs ? n_rs : 1), 100.0 * w_same / (w_all ? w_all : 1), 100.0 * id_rs_vr / (n_rs ? n_rs : 1), 100.0 * id_vr / (n_vr ? n_vr : 1));
    }
    // neighbourhood cache: occluder of the pixel to the left in the same frame and class
    {
        size_t occ = 0, hit = 0;
        for (size_t i = 1; i < vr_full.size(); i++)
        {
            const RayRec& r = vr_full[i];
            if (r.pix < 0 || r.occl_tri < 0) continue;
            occ++;
            if (vr_full[i - 1].pix >= 0 && tri_hits(vr_full[i - 1].occl_tri, r)) hit++;
        }
        printf("cache vr <- left neighbour's occluder: %.1f%% of occluded\n", 100.0 * hit / (occ ? occ : 1));
    }
    // histogram of triangle tests per ray (resolve)
    {
        std::map<unsigned, size_t> h;
        for (const RayRec& r : rs) h[r.tris > 40 ? 40 : r.tris]++;
        printf("resolve tri-test histogram:");
        for (auto& kv : h) printf(" %u:%zu", kv.first, kv.second);
        printf("\n");
    }
    // why are triangles tested that are not hit?  For the clear resolve rays: every tested triangle is classified by
    // whether the ray segment touches the triangle's own exact box (geometry: a triangle fills half its box) or only the
    // quantised / padded / leaf-union box (tree slack)
    {
        size_t tests = 0, own_box = 0, rays_n = 0, rej_plane = 0, rej_sphere = 0, rej_quadsphere = 0;
        auto seg_hits_box = [](f3 o, f3 d, const float lo[3], const float hi[3])
        {
            double t0 = 0.0, t1 = 0.99;
            const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
            for (int a = 0; a < 3; a++)
            {
                if (dd[a] == 0.0) { if (oo[a] < lo[a] || oo[a] > hi[a]) return false; continue; }
                double ta = (lo[a] - oo[a]) / dd[a], tb = (hi[a] - oo[a]) / dd[a];
                if (ta > tb) std::swap(ta, tb);
                t0 = std::max(t0, ta); t1 = std::min(t1, tb);
            }
            return t0 <= t1;
        };
        std::map<int, size_t> by_leaf;
        for (const RayRec& r : rs)
        {
            if (r.occl_tri >= 0) continue;
            rays_n++;
            const RaySetup rsu = setup_ray(r.o, r.d);
            Walk w;
            WalkStack wst;
            walk_begin(w, rsu);
            for (;;)
            {
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
                uint32_t imask;
                const uint32_t hits = intersect_node(bvh, node_idx, rsu, 0.0f, 0.99f, w.ng_base, w.tri_base, imask);
                w.ng_mask = (hits & 0xff000000u) | imask;
                uint32_t tm = hits & 0x00ffffffu;
                // leaf sizes of this node
                const WideNode& wn = bvh.nodes[node_idx];
                while (tm)
                {
                    const int i = 31 - clz32(tm & (0u - tm));
                    tm &= tm - 1u;
                    const WideTri& t = bvh.tris[w.tri_base + i];
                    const float lo[3] = {fmin3(t.v0x, t.v1x, t.v2x), fmin3(t.v0y, t.v1y, t.v2y), fmin3(t.v0z, t.v1z, t.v2z)};
                    const float hi[3] = {fmax3(t.v0x, t.v1x, t.v2x), fmax3(t.v0y, t.v1y, t.v2y), fmax3(t.v0z, t.v1z, t.v2z)};
                    tests++;
                    {
                        // candidate pre-rejects, evaluated in double (effectiveness only; margins come later)
                        const double v0[3] = {t.v0x, t.v0y, t.v0z}, v1[3] = {t.v1x, t.v1y, t.v1z}, v2[3] = {t.v2x, t.v2y, t.v2z};
                        double e0[3], e1[3], nn[3];
                        for (int a = 0; a < 3; a++) { e0[a] = v1[a] - v0[a]; e1[a] = v2[a] - v0[a]; }
                        nn[0] = e0[1] * e1[2] - e0[2] * e1[1]; nn[1] = e0[2] * e1[0] - e0[0] * e1[2]; nn[2] = e0[0] * e1[1] - e0[1] * e1[0];
                        const double oo[3] = {r.o.x, r.o.y, r.o.z}, dd[3] = {r.d.x, r.d.y, r.d.z};
                        double num = 0, den = 0;
                        for (int a = 0; a < 3; a++) { num += nn[a] * (v0[a] - oo[a]); den += nn[a] * dd[a]; }
                        const double tt = num / den;
                        if (!(tt >= 0.0 && tt <= 0.99)) rej_plane++;
                        else
                        {
                            // bounding sphere: midpoint of the longest edge if that covers the third vertex, else circumcentre
                            const double* vs[3] = {v0, v1, v2};
                            double best = -1; int bi = 0;
                            for (int k = 0; k < 3; k++)
                            {
                                double l2 = 0;
                                for (int a = 0; a < 3; a++) { const double d = vs[(k + 1) % 3][a] - vs[k][a]; l2 += d * d; }
                                if (l2 > best) { best = l2; bi = k; }
                            }
                            double c[3], r2 = best / 4;
                            for (int a = 0; a < 3; a++) c[a] = 0.5 * (vs[bi][a] + vs[(bi + 1) % 3][a]);
                            double d3 = 0;
                            for (int a = 0; a < 3; a++) { const double d = vs[(bi + 2) % 3][a] - c[a]; d3 += d * d; }
                            if (d3 > r2) r2 = d3;  // crude for acute triangles (not minimal, still bounding)
                            double dp = 0;
                            for (int a = 0; a < 3; a++) { const double d = oo[a] + tt * dd[a] - c[a]; dp += d * d; }
                            if (dp > r2) rej_sphere++;
                        }
                    }
                    const bool ob = seg_hits_box(r.o, r.d, lo, hi);
                    if (ob) own_box++;
                    int leaf_n = 0;
                    for (int s = 0; s < 8; s++)
                    {
                        const uint32_t m = wn.meta[s];
                        if (m && (m & 0x1f) < 24 && i >= (int)(m & 0x1f) && i < (int)(m & 0x1f) + popc(m >> 5)) leaf_n = popc(m >> 5);
                    }
                    by_leaf[leaf_n * 2 + (ob ? 1 : 0)]++;
                }
            }
        }
        printf("clear resolve rays %zu: %.2f failed triangle tests per ray, %.1f%% of them touch the triangle's own exact box\n", rays_n,
               (double)tests / rays_n, 100.0 * own_box / (tests ? tests : 1));
        printf("   pre-reject by plane side %.1f%%, then by bounding sphere at the crossing %.1f%% (of all failed tests)\n", 100.0 * rej_plane / tests, 100.0 * rej_sphere / tests);
        for (auto& kv : by_leaf) printf("   leaf size %d, own box %s: %.2f per ray\n", kv.first / 2, kv.first & 1 ? "hit " : "miss", (double)kv.second / rays_n);
    }
    // any-hit child order: front-to-back (the walk's) against back-to-front (the octant of the reversed direction)
    {
        auto walk_any = [&](const RayRec& r, bool reversed, unsigned& nodes, unsigned& tris)
        {
            RaySetup rsu = setup_ray(r.o, r.d);
            if (reversed) rsu.octinv ^= 7u;
            Walk w;
            WalkStack wst;
            walk_begin(w, rsu);
            nodes = tris = 0;
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
                uint32_t imask;
                const uint32_t hits = intersect_node(bvh, node_idx, rsu, 0.0f, 0.99f, w.ng_base, w.tri_base, imask);
                nodes++;
                w.ng_mask = (hits & 0xff000000u) | imask;
                uint32_t tm = hits & 0x00ffffffu;
                while (tm)
                {
                    const int i = 31 - clz32(tm & (0u - tm));
                    tm &= tm - 1u;
                    Hit h; h.prim = -1; h.t = 0.99f; h.u = h.v = 0;
                    tris++;
                    if (intersect_wide_tri(bvh.tris + w.tri_base + i, rsu, 0.0f, h)) return true;
                }
            }
        };
        for (int cls = 0; cls < 2; cls++)
        {
            const std::vector<RayRec>& rays = cls ? rs : vr;
            double n0 = 0, t0 = 0, n1 = 0, t1 = 0;
            size_t cnt = 0;
            for (const RayRec& r : rays)
            {
                unsigned a, b, c, d;
                const bool h0 = walk_any(r, false, a, b), h1 = walk_any(r, true, c, d);
                if (h0 != h1) printf("ORDER CHANGES RESULT?!\n");
                n0 += a; t0 += b; n1 += c; t1 += d; cnt++;
            }
            printf("any-hit order, %s rays: front-to-back nodes %.2f tris %.2f | back-to-front nodes %.2f tris %.2f\n", cls ? "resolve" : "visibility-reuse",
                   n0 / cnt, t0 / cnt, n1 / cnt, t1 / cnt);
            if (cls)
                for (int fl = 0; fl < 2; fl++)
                {
                    double a0 = 0, b0 = 0, a1 = 0, b1 = 0; size_t c2 = 0, occ = 0;
                    for (const RayRec& r : rays)
                    {
                        if (r.flag != fl) continue;
                        unsigned a, b, c, d;
                        walk_any(r, false, a, b); walk_any(r, true, c, d);
                        a0 += a; b0 += b; a1 += c; b1 += d; c2++; occ += r.occl_tri >= 0;
                    }
                    printf("   resolve rays whose sample is stored %s: %zu rays, %.1f%% occluded | front-to-back nodes %.2f tris %.2f | back-to-front nodes %.2f tris %.2f\n",
                           fl ? "visible" : "occluded", c2, 100.0 * occ / (c2 ? c2 : 1), a0 / c2, b0 / c2, a1 / c2, b1 / c2);
                }
        }
    }
    // dump rays for tree experiments
    if (const char* dump = getenv("LAB_DUMP"))
    {
        FILE* f = fopen(dump, "wb");
        auto put = [&](const std::vector<RayRec>& v, int cls)
        {
            for (const RayRec& r : v)
            {
                float rec[8] = {r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, (float)cls, (float)r.pix};
                fwrite(rec, 4, 8, f);
                const int own = vis[r.pix].index;  // the primitive the pixel's primary ray hit (the ray's origin lies on it)
                fwrite(&own, 4, 1, f);
            }
        };
        put(vr, 1);
        put(rs, 2);
        for (const RayRec& r : prim)  // primary rays: origin = eye, direction recomputed by the reader; only pix and prim kept
        {
            float rec[8] = {0, 0, 0, 0, 0, 0, 0.0f, (float)r.pix};
            fwrite(rec, 4, 8, f);
            fwrite(&r.occl_tri, 4, 1, f);
        }
        fclose(f);
    }
    return 0;
}
