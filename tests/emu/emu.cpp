// TEST INFRASTRUCTURE — host build of the CUDA path's device functions.
//
// There is no GPU in the build container, so the *logic* of the kernels (BVH build steps, the wide-BVH
// walk, every per-pixel body) is compiled here for the host from the very same headers the .cu files
// include (cedec-2024-rt_b200/csrc/*.cuh) and run sequentially, pixel by pixel, against the oracle
// (tests/test_emu_parity.py).  It is not part of libcedecrt.so, nothing in the product path can reach it,
// and it is never timed or shipped: it only lets `pytest -m "not gpu"` catch logic errors before GPU time
// is spent.  It exports the oracle's `orc_*` C API so the same Python driver (oracle/orc.py) runs it.
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "bvh_build.cuh"
#include "restir_fast.cuh"

using namespace crt;

namespace crt
{
int g_emu_postpone = 0;
}
extern "C" void emu_set_postpone(int v) { crt::g_emu_postpone = v; }

#if defined(CRT_COUNT)
namespace crt
{
thread_local unsigned long long g_count_nodes = 0, g_count_tris = 0;
}
static unsigned long long g_total_nodes = 0, g_total_tris = 0;
extern "C" void emu_counters(unsigned long long* out, int reset)
{
    // fold every OpenMP thread's counters into the totals
#pragma omp parallel
    {
#pragma omp critical
        {
            g_total_nodes += crt::g_count_nodes;
            g_total_tris += crt::g_count_tris;
            crt::g_count_nodes = crt::g_count_tris = 0;
        }
    }
    out[0] = g_total_nodes;
    out[1] = g_total_tris;
    if (reset) g_total_nodes = g_total_tris = 0;
}
#endif

namespace
{
struct EmuGeom
{
    std::vector<WideNode> nodes;
    std::vector<WideTri> tris;
    const float* tris60 = nullptr;
    int depth = 0;
    float pad = 0;
    std::vector<uint32_t> level_begin;  // csrc/geometry.cu: crt_geometry_t::level_begin
    Bvh view() const { return Bvh{nodes.data(), tris.data(), kPostponeRatio}; }
};
int g_math_mode = 0;
int g_resolve_reuse = 1;
long g_resolve_rays = 0;
long g_tid_begin = 0, g_tid_end = -1;

// same sequence as csrc/geometry.cu:build(), one loop per kernel
EmuGeom* build(const float* tris60, uint32_t n)
{
    EmuGeom* g = new EmuGeom;
    g->tris60 = tris60;
    if (n == 0)
    {
        WideNode root;
        memset(&root, 0, sizeof root);
        for (int s = 0; s < 8; s++)
            for (int a = 0; a < 3; a++) root.qlo[a][s] = 255;
        root.ex = root.ey = root.ez = 127;
        g->nodes.push_back(root);
        g->tris.resize(1);
        g->depth = 1;
        return g;
    }
    uint32_t b6[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    for (uint32_t i = 0; i < n; i++) tri_bounds(i, tris60, b6);
    float lo[3], hi[3], max_abs = 0;
    for (int a = 0; a < 3; a++)
    {
        lo[a] = ordered_to_float(b6[a]);
        hi[a] = ordered_to_float(b6[3 + a]);
        max_abs = fmaxf(max_abs, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
    }
    float pad_scale = 64.0f;
    if (const char* s = getenv("CRT_BVH_PAD_ULPS")) pad_scale = (float)atof(s);
    g->pad = pad_scale * 5.9604645e-8f * fmaxf(max_abs, 1.0f);
    f3 inv{hi[0] > lo[0] ? 1.0f / (hi[0] - lo[0]) : 0.0f, hi[1] > lo[1] ? 1.0f / (hi[1] - lo[1]) : 0.0f,
           hi[2] > lo[2] ? 1.0f / (hi[2] - lo[2]) : 0.0f};
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; i++) keys[i] = morton_key(i, tris60, f3{lo[0], lo[1], lo[2]}, inv);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> skeys(n);
    for (uint32_t i = 0; i < n; i++) skeys[i] = keys[idx[i]];

    const uint32_t ni = n - 1;
    std::vector<uint32_t> left(ni + 1), right(ni + 1), parent(2 * (size_t)n - 1), first(ni + 1), count(ni + 1),
        visits(ni + 1, 0u);
    std::vector<float> box((2 * (size_t)n - 1) * 6), cost((2 * (size_t)n - 1) * 7);
    std::vector<uint8_t> split((2 * (size_t)n - 1) * 8);
    BinTree bt{n, left.data(), right.data(), parent.data(), first.data(), count.data(), box.data(), visits.data(),
               cost.data(), split.data()};
    const char* builder = getenv("CRT_BVH_BUILDER");
    if (builder && !strcmp(builder, "lbvh"))
    {
        for (uint32_t i = 0; i < ni; i++) lbvh_node(i, skeys.data(), bt);
        for (uint32_t i = 0; i < n; i++) lbvh_refit(i, tris60, idx.data(), g->pad, bt);
    }
    else
    {
        // same round structure as csrc/geometry.cu:build_ploc()
        std::vector<uint32_t> node_a(n), node_b(n), nn(n);
        std::vector<float> box_a((size_t)n * 8), box_b((size_t)n * 8);
        std::vector<unsigned long long> flag(n), scan(n);
        for (uint32_t j = 0; j < n; j++) ploc_init_leaf(j, tris60, idx.data(), g->pad, bt, node_a.data(), box_a.data());
        uint32_t m = n, id_top = n >= 2 ? n - 2 : 0;
        uint32_t *ni_ = node_a.data(), *no_ = node_b.data();
        float *bi_ = box_a.data(), *bo_ = box_b.data();
        while (m > 1)
        {
            PlocRound pr{m, ni_, bi_, nn.data(), flag.data(), no_, bo_};
            for (uint32_t i = 0; i < m; i++) ploc_nn(i, pr);
            for (uint32_t i = 0; i < m; i++) ploc_flag(i, pr);
            unsigned long long acc = 0;
            for (uint32_t i = 0; i < m; i++)
            {
                scan[i] = acc;
                acc += flag[i];
            }
            for (uint32_t i = 0; i < m; i++) ploc_apply(i, pr, scan.data(), id_top, bt);
            id_top -= (uint32_t)(acc >> 32);
            m = (uint32_t)(acc & 0xffffffffull);
            std::swap(ni_, no_);
            std::swap(bi_, bo_);
        }
    }

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
    if (counters[1] != n) fprintf(stderr, "emu build: %u triangle records for %u triangles\n", counters[1], n);
    g->nodes.resize(counters[0]);
    return g;
}

template <class F>
void launch(int W, int H, F&& f)
{
    const long n = (long)W * H;
    const long t0 = g_tid_begin, t1 = g_tid_end < 0 ? n : (g_tid_end < n ? g_tid_end : n);
#pragma omp parallel for schedule(dynamic, 1024)
    for (long tid = t0; tid < t1; tid++) f(make_pix((int)(tid % W), (int)(tid / W), W, H));
}
f3 v3(const float* p) { return {p[0], p[1], p[2]}; }
}  // namespace

extern "C"
{
    int orc_example() { return 0; }
    const char* orc_kind() { return "emu"; }
    int orc_threads() { return omp_get_max_threads(); }
    void orc_set_threads(int n) { omp_set_num_threads(n); }
    void orc_set_math_mode(int m) { g_math_mode = m; }
    void orc_set_arg_order(int) {}
    void orc_set_range(long b, long e) { g_tid_begin = b; g_tid_end = e; }
    void* orc_geom_build(const crt_triangle* tris, int n) { return build((const float*)tris, (uint32_t)n); }
    void orc_geom_free(void* g) { delete (EmuGeom*)g; }
    // same sequence as csrc/geometry.cu:refit() — the caller has changed vertices of the array the tree was built over
    void emu_geom_refit(void* gp)
    {
        EmuGeom* g = (EmuGeom*)gp;
        const uint32_t n = (uint32_t)g->tris.size();
        if (g->level_begin.empty()) return;  // the empty tree
        float max_abs = 0;
        for (size_t i = 0; i < (size_t)n * 15; i++)
            if (i % 15 < 9) max_abs = fmaxf(max_abs, fabsf(g->tris60[i]));
        g->pad = 64.0f * 5.9604645e-8f * fmaxf(max_abs, 1.0f);
        std::vector<float> node_box(g->nodes.size() * 6);
        for (uint32_t i = 0; i < n; i++) refit_tri(i, g->tris60, g->tris.data());
        for (int l = (int)g->level_begin.size() - 2; l >= 0; l--)
            for (uint32_t i = g->level_begin[l]; i < g->level_begin[l + 1]; i++)
                refit_node(i, g->nodes.data(), g->tris.data(), node_box.data(), g->pad);
    }
    void emu_geom_stats(void* gp, double* out)
    {
        EmuGeom* g = (EmuGeom*)gp;
        out[0] = (double)g->tris.size();
        out[1] = (double)g->nodes.size();
        out[2] = g->depth;
        out[3] = g->pad;
    }
    void orc_lookat(const float* eye, const float* center, const float* up, float fovy, int W, int H, crt_raygen* rg)
    {
        // same arithmetic as crt_raygen_lookat (csrc/api.cu)
        const f3 e = v3(eye), c = v3(center), u0 = v3(up);
        const f3 f = normalize(c - e);
        const f3 s = normalize(cross(f, u0));
        const f3 u = cross(s, f);
        const float tan_y = tanf(fovy * 0.5f);
        const float tan_x = tan_y / (float)H * (float)W;
        const f3 r = s * tan_x, up2 = u * tan_y;
        rg->m_origin = {e.x, e.y, e.z};
        rg->m_right = {r.x, r.y, r.z};
        rg->m_up = {up2.x, up2.y, up2.z};
    }
    int orc_closest_hit(void* gp, const float* o, const float* d, float tmin, float tmax, float* tuv)
    {
        Hit h;
        trace<false>(((EmuGeom*)gp)->view(), v3(o), v3(d), tmin, tmax, h);
        if (h.prim < 0) return -1;
        tuv[0] = h.t; tuv[1] = h.u; tuv[2] = h.v;
        return h.prim;
    }
    // closest hit of a walk seeded the way px_raycast_hinted seeds it: the reference's test on triangle `hint` of `tris`
    // (any id; out of range = no seed), then bvh.cuh: trace_seeded
    int emu_closest_hit_seeded(void* gp, const crt_triangle* tris, int n_tris, const float* o, const float* d, int hint, float* tuv)
    {
        Hit h;
        h.prim = -1;
        h.t = kFltMax;
        h.u = h.v = 0.0f;
        if ((uint32_t)hint < (uint32_t)n_tris)
        {
            const TriRef tri = tri_at((const float*)tris, hint);
            float t, u, v;
            if (ray_triangle(v3(o), v3(d), 0.0f, kFltMax, tri.v(0), tri.v(1), tri.v(2), t, u, v))
            {
                h.t = t; h.u = u; h.v = v; h.prim = hint;
            }
        }
        trace_seeded<false>(((EmuGeom*)gp)->view(), v3(o), v3(d), 0.0f, h);
        if (h.prim < 0) return -1;
        tuv[0] = h.t; tuv[1] = h.u; tuv[2] = h.v;
        return h.prim;
    }
    int emu_any_hit(void* gp, const float* o, const float* d, float tmin, float tmax)
    {
        Hit h;
        return trace<true>(((EmuGeom*)gp)->view(), v3(o), v3(d), tmin, tmax, h) ? 1 : 0;
    }
    // the same question with the children nearest to the ray's end visited first (bvh.cuh: setup_ray far_first)
    int emu_any_hit_far_first(void* gp, const float* o, const float* d, float tmin, float tmax)
    {
        Hit h;
        return trace<true, true>(((EmuGeom*)gp)->view(), v3(o), v3(d), tmin, tmax, h) ? 1 : 0;
    }
    void orc_clear(crt_float4* buf, int W, int H)
    {
        launch(W, H, [&](Pix p) { buf[p.idx] = {0, 0, 0, 0}; });
    }
    void orc_tone_mapping(uint32_t* pixels, const crt_float4* accum, int W, int H)
    {
        launch(W, H, [&](Pix p)
               {
                   const crt_float4 a = accum[p.idx];
                   pixels[p.idx] = g_math_mode ? tone_map_rgba8<Math<1>>(f4{a.x, a.y, a.z, a.w})
                                               : tone_map_rgba8<Math<0>>(f4{a.x, a.y, a.z, a.w});
               });
    }
    void orc_raycast(int W, int H, void* gp, const crt_triangle*, int, const crt_raygen* rg, crt_visibility* vis)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p) { px_raycast(p, W, H, bvh, *rg, vis); });
    }
    // the hinted form (restir_pixel.cuh: px_raycast_hinted): `vis` holds the hints on entry, the answers on return
    void emu_raycast_hinted(int W, int H, void* gp, const crt_triangle* tris, int n_tris, const crt_raygen* rg, crt_visibility* vis)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p) { px_raycast_hinted(p, W, H, bvh, *rg, vis, (const float*)tris, (uint32_t)n_tris); });
    }
    void orc_generate_candidate(int W, int H, int frame, void* gp, const crt_triangle* tris, int,
                                const crt_visibility* vis, const float* eye, const uint32_t* lights, int nlights,
                                const crt_options* opt, crt_reservoir* res)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p)
               {
                   px_generate_candidate(p, frame, bvh, (const float*)tris, vis, v3(eye),
                                         LightsIndexed{(const float*)tris, lights, (uint32_t)nlights}, make_opt(*opt),
                                         AosStore{res});
               });
    }
    void orc_temporal_resampling(int W, int H, int frame, void* gp, const crt_triangle* tris, int,
                                 const crt_visibility* vis, const float* eye, const crt_options* opt,
                                 const crt_reservoir* prev, crt_reservoir* res)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p)
               {
                   if (g_math_mode)
                       px_temporal<Math<1>>(p, frame, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt),
                                            AosStore{(crt_reservoir*)prev}, AosStore{res});
                   else
                       px_temporal<Math<0>>(p, frame, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt),
                                            AosStore{(crt_reservoir*)prev}, AosStore{res});
               });
    }
    void orc_temporal_resampling_reprojected(int W, int H, int frame, void* gp, const crt_triangle* tris, int,
                                             const crt_visibility* vis, const float* eye, const crt_options* opt,
                                             const crt_raygen* prev_cam, const crt_reservoir* prev, crt_reservoir* res)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p)
               {
                   if (g_math_mode)
                       px_temporal<Math<1>>(p, frame, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt),
                                            AosStore{(crt_reservoir*)prev}, AosStore{res}, prev_cam, W, H);
                   else
                       px_temporal<Math<0>>(p, frame, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt),
                                            AosStore{(crt_reservoir*)prev}, AosStore{res}, prev_cam, W, H);
               });
    }
    void orc_save_temporal_reservoir(int W, int H, const crt_reservoir* src, crt_reservoir* dst)
    {
        memcpy(dst, src, (size_t)W * H * sizeof(crt_reservoir));
    }
    void orc_spatial_resampling(int W, int H, int frame, int pass, void* gp, const crt_triangle* tris, int,
                                const crt_visibility* vis, const float* eye, const crt_options* opt,
                                const crt_reservoir* in, crt_reservoir* out)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p)
               {
                   if (g_math_mode)
                       px_spatial<Math<1>>(p, W, H, frame, pass, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt),
                                           AosStore{(crt_reservoir*)in}, AosStore{out});
                   else
                       px_spatial<Math<0>>(p, W, H, frame, pass, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt),
                                           AosStore{(crt_reservoir*)in}, AosStore{out});
               });
    }
    void orc_resolve(crt_float4* accum, int W, int H, void* gp, const crt_triangle* tris, int,
                     const crt_visibility* vis, const float* eye, const crt_options* opt, const crt_reservoir* res)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p)
               { px_resolve(p, accum, bvh, (const float*)tris, vis, v3(eye), make_opt(*opt), AosStore{(crt_reservoir*)res}); });
    }
    static int g_example = 9;
    void orc_set_example(int e) { g_example = e; }
    void orc_path_trace(int W, int H, int frame, void* gp, const crt_triangle* tris, int, const uint32_t* lights,
                        int nlights, const crt_raygen* rg, const crt_options* opt, crt_float4* accum)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        const float* t = (const float*)tris;
        const Opt o = make_opt(*opt);
        const uint32_t nl = (uint32_t)nlights;
        launch(W, H, [&](Pix p)
               {
                   if (g_example == 7)
                       g_math_mode ? px_path_trace<7, Math<1>>(p, W, H, frame, bvh, t, lights, nl, *rg, o, accum)
                                   : px_path_trace<7, Math<0>>(p, W, H, frame, bvh, t, lights, nl, *rg, o, accum);
                   else if (g_example == 8)
                       g_math_mode ? px_path_trace<8, Math<1>>(p, W, H, frame, bvh, t, lights, nl, *rg, o, accum)
                                   : px_path_trace<8, Math<0>>(p, W, H, frame, bvh, t, lights, nl, *rg, o, accum);
                   else
                       g_math_mode ? px_path_trace<9, Math<1>>(p, W, H, frame, bvh, t, lights, nl, *rg, o, accum)
                                   : px_path_trace<9, Math<0>>(p, W, H, frame, bvh, t, lights, nl, *rg, o, accum);
               });
    }
    int orc_ao(uint32_t* pixels, const crt_raygen* rg, int W, int H, void* gp, const crt_triangle* tris, int,
               int n_rays)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        launch(W, H, [&](Pix p)
               {
                   uint32_t ao_rays = 0;
                   pixels[p.idx] = g_math_mode ? px_ao<Math<1>>(p, *rg, W, H, bvh, (const float*)tris, n_rays, ao_rays)
                                               : px_ao<Math<0>>(p, *rg, W, H, bvh, (const float*)tris, n_rays, ao_rays);
               });
        return 0;
    }

    // ---- the fused frame (csrc/kernels_fast.cu: crt_restir_di_frame), same kernel sequence, rays traced inline.
    // temporal / res_a / res_b: sector-planar storage of W*H*76 bytes each (restir_fast.cuh: SoaStore).
    // Final spatial output: res_a for an odd number of passes, res_b for an even one, `temporal` if spatial is off.
    // prev_rg != nullptr: temporal reprojection inside the fused frame (crt_restir_set_previous_camera): the history is a
    // snapshot of `temporal`, read at the reprojected pixel
    void emu_restir_frame_fast_reprojected(int W, int H, int frame, void* gp, const crt_triangle* tris, const crt_raygen* rg,
                                           const float* eye_p, const uint32_t* lights, int nlights, const crt_options* options,
                                           crt_visibility* vis, char* temporal, char* res_a, char* res_b, crt_float4* accum,
                                           uint32_t* pixels, const crt_raygen* prev_rg);
    void emu_restir_frame_fast(int W, int H, int frame, void* gp, const crt_triangle* tris, const crt_raygen* rg,
                               const float* eye_p, const uint32_t* lights, int nlights, const crt_options* options,
                               crt_visibility* vis, char* temporal, char* res_a, char* res_b, crt_float4* accum,
                               uint32_t* pixels)
    {
        emu_restir_frame_fast_reprojected(W, H, frame, gp, tris, rg, eye_p, lights, nlights, options, vis, temporal, res_a, res_b,
                                          accum, pixels, nullptr);
    }
    void emu_restir_frame_fast_reprojected(int W, int H, int frame, void* gp, const crt_triangle* tris, const crt_raygen* rg,
                                           const float* eye_p, const uint32_t* lights, int nlights, const crt_options* options,
                                           crt_visibility* vis, char* temporal, char* res_a, char* res_b, crt_float4* accum,
                                           uint32_t* pixels, const crt_raygen* prev_rg)
    {
        const Bvh bvh = ((EmuGeom*)gp)->view();
        const float* t60 = (const float*)tris;
        const f3 eye = v3(eye_p);
        const Opt opt = make_opt(*options);
        const size_t n = (size_t)W * H;
        std::vector<char> g0(n * 16), g1(n * 8);
        std::vector<uint8_t> cls(n, 0);
        const GBuf g{g0.data(), g1.data(), cls.data()};
        const SoaStore T{temporal, n}, A{res_a, n}, B{res_b, n};
        const LightsIndexed L{t60, lights, (uint32_t)nlights};
        launch(W, H, [&](Pix p) { px_raycast(p, W, H, bvh, *rg, vis); });
        std::vector<char> snapshot;
        Reprojection rp;
        if (prev_rg && opt.temporal)
        {
            snapshot.assign(temporal, temporal + n * 72);
            rp.history = SoaStore{snapshot.data(), n};
            rp.prev_cam = *prev_rg;
            rp.W = W;
            rp.H = H;
        }
        const bool reproject = prev_rg && opt.temporal;
        launch(W, H, [&](Pix p)
               {
                   const CandPixel cp = classify_pixel(p, t60, vis);
                   DeferredRay d;
                   if (reproject)
                       d = g_math_mode ? px_candidate_temporal<Math<1>, LightsIndexed, true>(p, cp, frame, bvh, t60, eye, L, opt, T, g, HaloPeers(), 0u, rp)
                                       : px_candidate_temporal<Math<0>, LightsIndexed, true>(p, cp, frame, bvh, t60, eye, L, opt, T, g, HaloPeers(), 0u, rp);
                   else
                       d = g_math_mode ? px_candidate_temporal<Math<1>>(p, cp, frame, bvh, t60, eye, L, opt, T, g)
                                       : px_candidate_temporal<Math<0>>(p, cp, frame, bvh, t60, eye, L, opt, T, g);
                   if (d.want)
                   {
                       Hit h;
                       if (!trace<true, true>(bvh, d.org, d.dir, 0.0f, 0.99f, h)) *T.mword(p.idx) |= kVisBit;  // k_trace_shadow_queue<2>
                   }
               });
        SoaStore in = T, out = A;
        if (opt.spatial)
            for (int pass = 0; pass < options->spatial_resampling_passes; pass++)
            {
                if (pass == 1) { in = A; out = B; }
                else if (pass > 1) std::swap(in, out);
                launch(W, H, [&](Pix p)
                       {
                           g_math_mode ? px_spatial_fast<Math<1>>(p, W, H, frame, pass, bvh, eye, opt, in, out, g)
                                       : px_spatial_fast<Math<0>>(p, W, H, frame, pass, bvh, eye, opt, in, out, g);
                       });
            }
        const SoaStore fin = opt.spatial && options->spatial_resampling_passes > 0 ? out : T;
        launch(W, H, [&](Pix p)
               {
                   DeferredShade sh{{0, 0, 0}, {0, 0, 0}, 0.0f};
                   const DeferredRay d = px_resolve_fast(p, accum, t60, vis, fin, g, sh, opt.accumulate, g_resolve_reuse && opt.reuse);
                   if (!d.want) return;
#pragma omp atomic
                   g_resolve_rays++;
                   Hit h;
                   const float V = trace<true>(bvh, d.org, d.dir, 0.0f, 0.99f, h) ? 0.0f : 1.0f;
                   write_accum(accum, p.idx, sh.bg * V * sh.rad * sh.ucw, opt.accumulate);  // shadow_epilogue<kEpiResolve>
               });
        orc_tone_mapping(pixels, accum, W, H);
    }
    // resolve: 1 (default) uses the stored answer of an already traced shadow ray, 0 traces every ray (kernels_fast.cu:
    // crt_ctx::resolve_reuse); emu_resolve_rays: shadow rays traced by resolve since the last call
    void emu_set_resolve_reuse(int v) { g_resolve_reuse = v; }
    long emu_resolve_rays()
    {
        const long r = g_resolve_rays;
        g_resolve_rays = 0;
        return r;
    }
    void emu_soa_to_aos(const char* soa, crt_reservoir* aos, long n)
    {
        const SoaStore s{const_cast<char*>(soa), (size_t)n};
        for (long i = 0; i < n; i++) soa_to_aos(s, aos, (int)i);
    }
    void emu_aos_to_soa(const crt_reservoir* aos, char* soa, long n)
    {
        const SoaStore s{soa, (size_t)n};
        for (long i = 0; i < n; i++) aos_to_soa(aos, s, (int)i);
    }
}
