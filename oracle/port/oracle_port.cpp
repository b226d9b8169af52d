// TEST INFRASTRUCTURE — CPU restatement ("port") of the ReSTIR DI hot path of
// yumcyaWiz/CEDEC-2024-RT (examples 04, 06-10).  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the
// product path (cedec-2024-rt_b200/) never does.
//
// Parity status: PINNED.  tests/test_oracle_pinning.py checks this restatement
// bit for bit (math mode 0) against oracle/_ref — the reference's own unmodified
// .cu files compiled as host C++ — wherever /root/reference exists, and against
// golden fixtures generated from oracle/_ref (tests/golden/, script
// tests/golden/make_golden.py) everywhere else.
// What is *not* pinned by the reference: exact t/uv/edge ownership of HIPRT
// 2.4.6b6daf9's closed-source traversal.  Both oracles use the reference's own
// intersect_ray_triangle (common/core.hpp:91-136) with the brute-force tie rule
// of examples/04_ao/04_ao.cu:14-24 instead (see oracle/cpu_bvh.h).
#include <omp.h>

#include <cstdio>
#include <cstdlib>

#include "../cpu_bvh.h"
#include "port_math.h"

namespace port
{
int g_math_mode = 0;

// The reference draws several randoms inside one argument list, e.g.
// sample_light(..., random.uniformf(), random.uniformf(), random.uniformf()) (10_restir_di.cu:88-90).
// C++ leaves the evaluation order of arguments unspecified: the GPU compilers the reference targets
// (NVRTC / hiprtc, clang-style) evaluate left to right — that is the canonical order (0) and what the
// CUDA path implements; g++ on x86-64 evaluates right to left, so oracle/_ref, the reference compiled
// as host C++, sees the draws reversed.  Order 1 reproduces that and exists only so that
// tests/test_oracle_pinning.py can pin this restatement bit for bit against oracle/_ref.
int g_arg_order = 0;

// ---------------------------------------------------------------- structs (reference layouts)
struct Triangle  // common/core.hpp:38-43, 60 B
{
    V3 v[3];
    V3 color;
    V3 emissive;
};
struct Visibility  // core.hpp:167-172, 16 B
{
    V2 uv;
    int index;
    int pad;
};
struct Sample  // common/reservoir.hpp:5-13, 64 B
{
    V3 origin_position, origin_normal, hit_position, hit_normal, radiance;
    bool visibility;
};
struct Reservoir  // reservoir.hpp:15-38, 76 B
{
    Sample s;
    float w_sum;
    float ucw;
    int M;
};
struct Options  // common/options.hpp:4-23, 48 B
{
    bool accumulate;
    int max_depth;
    V3 sky_color;
    int ris_sample_count;
    float rejection_heuristics_threshold;
    bool use_temporal_resampling;
    bool use_spatial_resampling;
    int spatial_resampling_sample_count;
    float spatial_resampling_radius;
    int spatial_resampling_passes;
    bool use_shadowed_target_function;
    bool use_visibility_reuse;
};
struct RayGen  // common/camera.hpp:5-9, 36 B
{
    V3 origin, right, up;
};
static_assert(sizeof(Triangle) == 60 && sizeof(Visibility) == 16 && sizeof(Sample) == 64, "layout");
static_assert(sizeof(Reservoir) == 76 && sizeof(Options) == 48 && sizeof(RayGen) == 36, "layout");

static Reservoir empty_reservoir()
{
    Reservoir r;
    memset(&r, 0, sizeof r);  // Reservoir{}: all-zero members (reservoir.hpp:7-21)
    return r;
}

// ---------------------------------------------------------------- camera (common/camera.hpp)
static RayGen lookat(V3 eye, V3 center, V3 up, float fovy, int W, int H)  // camera.hpp:11-25
{
    const V3 f = normalize(center - eye);
    const V3 s = normalize(cross(f, up));
    const V3 u = cross(s, f);
    const float tan_y = tanf(fovy * 0.5f);
    const float tan_x = tan_y / (float)H * (float)W;
    return {eye, s * tan_x, u * tan_y};
}
static void shoot(const RayGen& rg, float u, float v, V3& ro, V3& rd)  // camera.hpp:27-35
{
    const V3 forward = normalize(cross(rg.up, rg.right));
    const V3 to = rg.origin + forward + mix(-rg.right, rg.right, u) + mix(rg.up, -rg.up, v);
    ro = rg.origin;
    rd = normalize(to - rg.origin);
}

// ---------------------------------------------------------------- triangle helpers (core.hpp:45-69)
static V3 tangent_of(const Triangle& t) { return normalize(t.v[1] - t.v[0]); }
static V3 normal_of(const Triangle& t) { return normalize(cross(t.v[1] - t.v[0], t.v[2] - t.v[0])); }
static float area_of(const Triangle& t) { return 0.5f * length(cross(t.v[1] - t.v[0], t.v[2] - t.v[0])); }
static bool has_emission(const Triangle& t) { return t.emissive.x > 0.0f || t.emissive.y > 0.0f || t.emissive.z > 0.0f; }

// core.hpp:91-136: plane hit, then three signed sub-areas; u = area(p,v2,v0)/A, v = area(p,v0,v1)/A
static bool ray_triangle(float& t_out, float& u_out, float& v_out, V3 ro, V3 rd, float tmin, float tmax, V3 v0,
                         V3 v1, V3 v2)
{
    const V3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    const V3 n = cross(e0, e1);
    const float t = dot(v0 - ro, n) / dot(n, rd);
    if (!(tmin <= t && t <= tmax)) return false;  // also false for NaN
    const V3 p = ro + rd * t;
    const float a0 = dot(n, cross(e0, p - v0));
    const float a1 = dot(n, cross(e1, p - v1));
    const float a2 = dot(n, cross(e2, p - v2));
    if (a0 < 0.0f || a1 < 0.0f || a2 < 0.0f) return false;
    const float a = a0 + a1 + a2;
    t_out = t;
    u_out = a2 / a;  // core.hpp:127-133: uOut = bV, vOut = bW
    v_out = a0 / a;
    return true;
}

// ---------------------------------------------------------------- traversal stand-in (raytrace.hpp:18-52)
struct Geom
{
    cpubvh::Bvh* bvh;
    const Triangle* tris;
};
struct Isect
{
    float t = 0.0f;
    V2 uv = {0.0f, 0.0f};
    int index = -1;
};
static bool raytrace(const Geom& g, V3 ro, V3 rd, float tmin, float tmax, Isect& is, bool any_hit = false)
{
    const float o[3] = {ro.x, ro.y, ro.z}, d[3] = {rd.x, rd.y, rd.z};
    const Triangle* tris = g.tris;
    cpubvh::Hit h = cpubvh::trace(
        *g.bvh, o, d, tmin, tmax,
        [&](int prim, float t0, float t1, float& t, float& u, float& v)
        { return ray_triangle(t, u, v, ro, rd, t0, t1, tris[prim].v[0], tris[prim].v[1], tris[prim].v[2]); },
        any_hit);
    if (h.prim < 0) return false;
    is.t = h.t;
    is.uv = {h.u, h.v};
    is.index = h.prim;
    return true;
}
// raytrace.hpp:45-52 — origin p0 + 1e-3 n0 (core.hpp:32-36), direction p1 - p0 (not renormalised), t in [0, 0.99].
// The reference asks for the closest hit and only uses hit/no-hit, so an any-hit walk is result-equivalent.
static float check_visibility(const Geom& g, V3 p0, V3 n0, V3 p1)
{
    Isect is;
    return raytrace(g, p0 + 0.001f * n0, p1 - p0, 0.0f, 0.99f, is, true) ? 0.0f : 1.0f;
}

// ---------------------------------------------------------------- surfaces (core.hpp:145-207)
struct Surf
{
    V3 p, n;
};
static V3 bary_point(const Triangle& t, V2 uv) { return (1.0f - uv.x - uv.y) * t.v[0] + uv.x * t.v[1] + uv.y * t.v[2]; }
static Surf surface_from_visibility(const Visibility& vis, const Triangle* tris, V3 eye)  // core.hpp:188-207
{
    const Triangle& t = tris[vis.index];
    Surf s{bary_point(t, vis.uv), normal_of(t)};
    const V3 view = normalize(eye - s.p);
    if (dot(view, s.n) < 0.0f) s.n = -s.n;
    return s;
}
static Surf surface_from_hit(V3 ro, V3 rd, const Isect& is, const Triangle* tris)  // core.hpp:152-165
{
    Surf s{ro + is.t * rd, normal_of(tris[is.index])};
    if (dot(-rd, s.n) < 0.0f) s.n = -s.n;
    return s;
}

// ---------------------------------------------------------------- sampling (core.hpp:76-89, 237-295)
static V3 sample_hemisphere(float r0, float r1, float r2)
{
    const float theta = r0 * 2.0f * kPi;
    float radius = r1 + r2;
    if (1.0f < radius) radius = 2.0f - radius;
    const float x = m_cos(theta) * radius;
    const float z = m_sin(theta) * radius;
    const float yy = 1.0f - radius * radius;
    return {x, sqrtf(yy < 0.0f ? 0.0f : yy), z};
}
static V2 warp_unit_triangle(float x, float y)  // core.hpp:237-252 (Heitz 2019)
{
    if (y > x) { x *= 0.5f; y -= x; }
    else { y *= 0.5f; x -= y; }
    return {x, y};
}
struct LightSample
{
    V3 p, n;
    int index;
};
static LightSample sample_light(const Triangle* tris, const uint32_t* lights, size_t n_lights, float rv0, float rv1,
                                float rv2)  // core.hpp:261-285
{
    uint32_t nth = (uint32_t)(rv0 * (float)n_lights);
    if (nth == n_lights) nth = (uint32_t)n_lights - 1;
    LightSample ls;
    ls.index = (int)lights[nth];
    const Triangle& t = tris[ls.index];
    ls.p = bary_point(t, warp_unit_triangle(rv1, rv2));
    ls.n = normal_of(t);
    return ls;
}
static float geometry_term(V3 p0, V3 n0, V3 p1, V3 n1)  // core.hpp:287-295
{
    V3 v = p1 - p0;
    const float sqr = dot(v, v);
    v = normalize(v);
    return fabsf(dot(v, n0)) * fabsf(dot(-v, n1)) / sqr;
}

// ---------------------------------------------------------------- reservoir.hpp
static void res_update(Reservoir& r, const Sample& s, float w, float u)  // reservoir.hpp:22-29
{
    r.w_sum += w;
    r.M += 1;
    if (u < w / r.w_sum) r.s = s;
}
static void res_merge(Reservoir& r, const Reservoir& o, float w, float u)  // reservoir.hpp:31-37
{
    r.w_sum += w;
    r.M += o.M;
    if (u < w / r.w_sum) r.s = o.s;
}
static float target_function(const Geom& g, V3 p0, V3 n0, V3 p1, V3 n1, V3 radiance, bool shadowed)  // 42-59
{
    const float brdf = 1.0f / kPi;
    const float G = geometry_term(p0, n0, p1, n1);
    if (shadowed) return brdf * G * check_visibility(g, p0, n0, p1) * luminance(radiance);
    return brdf * G * luminance(radiance);
}
static float rejection_heuristics(const Reservoir& r0, const Reservoir& r1, V3 eye)  // reservoir.hpp:61-87
{
    const float d0 = length(r0.s.origin_position - eye);
    const float d1 = length(r1.s.origin_position - eye);
    const float diff = (d1 - d0) * (d1 - d0) / d0;
    float w = 1.0f;
    w *= m_exp(-32.0f * diff);
    const float c = dot(r0.s.origin_normal, r1.s.origin_normal);
    w *= m_pow(c > 0.0f ? c : 0.0f, 8.0f);
    return w;
}
static V2 sample_2d_gaussian(float rv0, float rv1)  // reservoir.hpp:89-95 (Box-Muller)
{
    const float a = -2.0f * m_log(rv0);
    const float radius = sqrtf(a > 0.0f ? a : 0.0f);
    const float phi = 2.0f * kPi * rv1;
    return {radius * m_cos(phi), radius * m_sin(phi)};
}
static float ucw_of(const Reservoir& r, float p_hat) { return p_hat > 0.0f ? r.w_sum / ((float)r.M * p_hat) : 0.0f; }

// float -> int the way x86 cvttss2si does it for the values the reference meets (inf/NaN -> INT_MIN)
static int to_int(float f)
{
    if (!(f > -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}

static void draw3(Pcg& rng, float& a, float& b, float& c)
{
    if (g_arg_order == 0) { a = rng.next_f(); b = rng.next_f(); c = rng.next_f(); }
    else { c = rng.next_f(); b = rng.next_f(); a = rng.next_f(); }
}
static void draw2(Pcg& rng, float& a, float& b)
{
    if (g_arg_order == 0) { a = rng.next_f(); b = rng.next_f(); }
    else { b = rng.next_f(); a = rng.next_f(); }
}

// ---------------------------------------------------------------- per-pixel kernels
struct Px
{
    int xi, yi, idx;
};
static Px pixel_of(long tid, int W, int H)  // e.g. 10_restir_di.cu:14-20: buffers are stored bottom-up
{
    Px p;
    p.xi = (int)(tid % W);
    p.yi = (int)(tid / W);
    p.idx = p.xi + (H - p.yi - 1) * W;
    return p;
}

// RIS over the emissive triangles, shared by generate_candidate (10_restir_di.cu:78-111) and 09_ris.cu:66-100
static Reservoir ris_candidates(const Geom& g, const Surf& surf, const uint32_t* lights, size_t n_lights, int count,
                                bool shadowed, Pcg& rng)
{
    Reservoir r = empty_reservoir();
    for (int i = 0; i < count; ++i)
    {
        Sample s;
        memset(&s, 0, sizeof s);
        s.origin_position = surf.p;
        s.origin_normal = surf.n;
        float r0, r1, r2;
        draw3(rng, r0, r1, r2);
        const LightSample ls = sample_light(g.tris, lights, n_lights, r0, r1, r2);
        s.hit_position = ls.p;
        s.hit_normal = ls.n;
        const Triangle& lt = g.tris[ls.index];
        s.radiance = lt.emissive;
        const float light_pdf = 1.0f / (float)n_lights * 1.0f / area_of(lt);
        const float p_hat = target_function(g, surf.p, surf.n, s.hit_position, s.hit_normal, s.radiance, shadowed);
        res_update(r, s, p_hat / light_pdf, rng.next_f());
    }
    return r;
}

static void k_raycast(long tid, int W, int H, const Geom& g, const RayGen& rg, Visibility* vis)  // 10_restir_di.cu:9-34
{
    const Px px = pixel_of(tid, W, H);
    V3 ro, rd;
    shoot(rg, (float)px.xi / (float)W, (float)px.yi / (float)H, ro, rd);
    Isect is;
    raytrace(g, ro, rd, 0.0f, kFltMax, is);
    vis[px.idx] = Visibility{is.uv, is.index, 0};
}

static void k_generate_candidate(long tid, int W, int H, int frame, const Geom& g, const Visibility* vis, V3 eye,
                                 const uint32_t* lights, size_t n_lights, const Options& opt,
                                 Reservoir* out)  // 10_restir_di.cu:36-135
{
    const Px px = pixel_of(tid, W, H);
    const Visibility v = vis[px.idx];
    if (v.index == -1 || has_emission(g.tris[v.index]))
    {
        out[px.idx] = empty_reservoir();
        return;
    }
    Pcg rng(hash_pcg4(px.xi, px.yi, frame, 0), 0);
    const Surf surf = surface_from_visibility(v, g.tris, eye);
    Reservoir r = ris_candidates(g, surf, lights, n_lights, opt.ris_sample_count, false, rng);
    r.ucw = ucw_of(r, target_function(g, surf.p, surf.n, r.s.hit_position, r.s.hit_normal, r.s.radiance,
                                      opt.use_shadowed_target_function));
    if (opt.use_visibility_reuse) r.s.visibility = check_visibility(g, surf.p, surf.n, r.s.hit_position) != 0.0f;
    out[px.idx] = r;
}

// ---- temporal reprojection — NOT in the reference: 10_restir_di.cu:174-180 reads the previous
// reservoir at the same pixel_idx and the host only clears the accumulation when the camera moves (10_restir_di.cpp:257-267).
// This function is therefore the *specification* of the extension; the CUDA path (csrc/restir_core.cuh: reproject_pixel)
// restates it operation for operation.
//   The surface point of the current pixel is projected into the previous frame's camera — the inverse of
//   RayGenerator::shoot (camera.hpp:27-35): to = o + forward + right (2u - 1) + up (1 - 2v) — and the pixel whose sample
//   point (u, v) = (xi / W, yi / H) is nearest is taken: xi = floor(u W + 1/2), yi = floor(v H + 1/2).  Behind the previous
//   camera or outside its image there is no history.  Whether the surface found there is the same one is left to the
//   reference's own rejection heuristics (reservoir.hpp:61-87), which scale M by depth and normal agreement.
//   With an unmoved camera this returns the pixel itself (the hit point lies on the pixel's own ray up to rounding), so
//   temporal_resampling_reprojected == temporal_resampling bit for bit (tests/test_reprojection.py).
static bool reproject_pixel(const RayGen& prev, int W, int H, V3 p, int& xp, int& yp)
{
    const V3 forward = normalize(cross(prev.up, prev.right));
    const V3 d = p - prev.origin;
    const float s = dot(d, forward);
    if (!(s > 0.0f)) return false;
    const V3 q = d / s;
    const float a = dot(q, prev.right) / dot(prev.right, prev.right);
    const float b = dot(q, prev.up) / dot(prev.up, prev.up);
    const float u = (a + 1.0f) * 0.5f;
    const float v = (1.0f - b) * 0.5f;
    xp = to_int(floorf(u * (float)W + 0.5f));
    yp = to_int(floorf(v * (float)H + 0.5f));
    return xp >= 0 && xp < W && yp >= 0 && yp < H;
}

// prev_cam == nullptr: the reference's kernel (same pixel_idx)
static void k_temporal(long tid, int W, int H, int frame, const Geom& g, const Visibility* vis, V3 eye,
                       const Options& opt, const Reservoir* prev_buf, Reservoir* cur,
                       const RayGen* prev_cam = nullptr)  // 10_restir_di.cu:137-237
{
    const Px px = pixel_of(tid, W, H);
    const Visibility v = vis[px.idx];
    if (v.index == -1 || has_emission(g.tris[v.index])) return;
    if (!opt.use_temporal_resampling) return;
    Pcg rng(hash_pcg4(px.xi, px.yi, frame, 1), 0);
    const Surf surf = surface_from_visibility(v, g.tris, eye);
    Reservoir prev = prev_buf[px.idx];
    if (prev_cam)
    {
        int xp, yp;
        // no history: the merge below runs with Reservoir{} (M = 0, weight 0), so the random stream does not depend on it
        prev = reproject_pixel(*prev_cam, W, H, surf.p, xp, yp) ? prev_buf[xp + (H - yp - 1) * W] : empty_reservoir();
    }
    Reservoir r = cur[px.idx];
    const int cap = 20 * opt.ris_sample_count;  // M-cap, 10_restir_di.cu:186-188
    prev.M = prev.M < cap ? prev.M : cap;
    float p_hat_y = target_function(g, surf.p, surf.n, prev.s.hit_position, prev.s.hit_normal, prev.s.radiance,
                                    opt.use_shadowed_target_function);
    if (opt.use_visibility_reuse) p_hat_y *= (float)prev.s.visibility;
    prev.M = to_int((float)prev.M * rejection_heuristics(r, prev, eye));  // int *= float truncates (211-212)
    const float weight = p_hat_y * prev.ucw * (float)prev.M;
    res_merge(r, prev, weight, rng.next_f());
    r.ucw = ucw_of(r, target_function(g, surf.p, surf.n, r.s.hit_position, r.s.hit_normal, r.s.radiance,
                                      opt.use_shadowed_target_function));
    cur[px.idx] = r;
}

static void k_spatial(long tid, int W, int H, int frame, int pass, const Geom& g, const Visibility* vis, V3 eye,
                      const Options& opt, const Reservoir* in, Reservoir* out)  // 10_restir_di.cu:256-388
{
    const Px px = pixel_of(tid, W, H);
    const Visibility v = vis[px.idx];
    if (v.index == -1 || has_emission(g.tris[v.index])) return;  // output left untouched
    Pcg rng(hash_pcg4(px.xi, px.yi, frame, 2 + pass), 0);
    const Surf surf = surface_from_visibility(v, g.tris, eye);
    Reservoir r = in[px.idx];
    if (!opt.use_spatial_resampling)
    {
        out[px.idx] = r;
        return;
    }
    for (int k = 0; k < opt.spatial_resampling_sample_count; ++k)
    {
        float rv0, rv1;
        draw2(rng, rv0, rv1);
        const V2 gs = sample_2d_gaussian(rv0, rv1);
        const int x = to_int((float)px.xi + opt.spatial_resampling_radius / 1.96f * gs.x);
        const int y = to_int((float)px.yi + opt.spatial_resampling_radius / 1.96f * gs.y);
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        if (x == px.xi && y == px.yi) continue;
        const int pid = x + (H - y - 1) * W;
        const Visibility nv = vis[pid];
        if (nv.index == -1 || has_emission(g.tris[nv.index])) continue;
        Reservoir nb = in[pid];
        float p_hat_y = target_function(g, surf.p, surf.n, nb.s.hit_position, nb.s.hit_normal, nb.s.radiance,
                                        opt.use_shadowed_target_function);
        if (opt.use_visibility_reuse) p_hat_y *= (float)nb.s.visibility;
        nb.M = to_int((float)nb.M * rejection_heuristics(r, nb, eye));  // compares against the *running* reservoir
        const float weight = p_hat_y * nb.ucw * (float)nb.M;
        res_merge(r, nb, weight, rng.next_f());  // third random only for surviving neighbours
    }
    r.ucw = ucw_of(r, target_function(g, surf.p, surf.n, r.s.hit_position, r.s.hit_normal, r.s.radiance,
                                      opt.use_shadowed_target_function));
    out[px.idx] = r;
}

static void accumulate(V4* accum, int idx, V3 c, bool add)  // e.g. 10_restir_di.cu:451-458
{
    if (add) accum[idx] = {accum[idx].x + c.x, accum[idx].y + c.y, accum[idx].z + c.z, accum[idx].w + 1.0f};
    else accum[idx] = {c.x, c.y, c.z, 1.0f};
}

static void k_resolve(long tid, V4* accum, int W, int H, const Geom& g, const Visibility* vis, V3 eye,
                      const Options& opt, const Reservoir* res)  // 10_restir_di.cu:390-459
{
    const Px px = pixel_of(tid, W, H);
    const Visibility v = vis[px.idx];
    if (v.index == -1)
    {
        accum[px.idx] = {0.0f, 0.0f, 0.0f, 1.0f};  // assigned even when accumulating
        return;
    }
    const Triangle& tri = g.tris[v.index];
    if (has_emission(tri))
    {
        accum[px.idx] = {tri.emissive.x, tri.emissive.y, tri.emissive.z, 1.0f};
        return;
    }
    const Surf surf = surface_from_visibility(v, g.tris, eye);
    const Reservoir& r = res[px.idx];
    const V3 brdf = 1.0f / kPi * tri.color;
    const float G = geometry_term(surf.p, surf.n, r.s.hit_position, r.s.hit_normal);
    const float V = check_visibility(g, surf.p, surf.n, r.s.hit_position);
    const V3 radiance = brdf * G * V * r.s.radiance * r.ucw;
    accumulate(accum, px.idx, radiance, opt.accumulate);
}

static void k_clear(long tid, V4* buf, int W, int H)  // common/kernels/common.cu:4-17
{
    buf[pixel_of(tid, W, H).idx] = {0.0f, 0.0f, 0.0f, 0.0f};
}
static float aces(float x)  // common.cu:19-28 (Narkowicz 2015)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return (x * (a * x + b)) / (x * (c * x + d) + e);
}
static uint8_t to_u8(float v)  // clamp(v*255, 0, 255) then truncation, common.cu:66-72
{
    float s = v * 255.0f;
    s = s > 0.0f ? s : 0.0f;  // max(x, 0): NaN -> 0
    s = s < 255.0f ? s : 255.0f;
    return (uint8_t)s;
}
static void k_tone_mapping(long tid, uint8_t* pixels, const V4* accum, int W, int H)  // common.cu:30-74
{
    const int idx = pixel_of(tid, W, H).idx;
    const V4 a = accum[idx];
    const float gamma = 1.0f / 2.2f;
    pixels[4 * idx + 0] = to_u8(m_pow(aces(a.x / a.w), gamma));
    pixels[4 * idx + 1] = to_u8(m_pow(aces(a.y / a.w), gamma));
    pixels[4 * idx + 2] = to_u8(m_pow(aces(a.z / a.w), gamma));
    pixels[4 * idx + 3] = 255;
}

// next bounce direction, shared by 07/08/09 (e.g. 08_nee.cu:98-106; core.hpp:209-235)
static V3 bounce_direction(const Surf& surf, const Triangle& tri, Pcg& rng)
{
    const V3 t = tangent_of(tri);
    const V3 b = normalize(cross(t, surf.n));
    float r0, r1, r2;
    draw3(rng, r0, r1, r2);
    const V3 l = sample_hemisphere(r0, r1, r2);
    return l.x * t + l.y * surf.n + l.z * b;
}

// mode 7: 07_pt.cu:11-90; mode 8: 08_nee.cu:11-129; mode 9: 09_ris.cu:11-166
static void k_path_trace(long tid, int mode, int W, int H, int frame, const Geom& g, const uint32_t* lights,
                         size_t n_lights, const RayGen& rg, const Options& opt, V4* accum)
{
    const Px px = pixel_of(tid, W, H);
    Pcg rng(hash_pcg3(px.xi, px.yi, frame), 0);
    V3 ro, rd;
    shoot(rg, (float)px.xi / (float)W, (float)px.yi / (float)H, ro, rd);
    V3 radiance = {0, 0, 0}, throughput = {1, 1, 1};
    for (int depth = 0; depth < opt.max_depth; ++depth)
    {
        Isect is;
        if (!raytrace(g, ro, rd, 0.0f, kFltMax, is))
        {
            if (mode == 7) radiance = radiance + throughput * opt.sky_color;
            break;
        }
        const Triangle& tri = g.tris[is.index];
        if (has_emission(tri))
        {
            if (mode == 7 || depth == 0) radiance = radiance + throughput * tri.emissive;
            break;
        }
        const Surf surf = surface_from_hit(ro, rd, is, g.tris);
        if (mode == 8)
        {
            float r0, r1, r2;
            draw3(rng, r0, r1, r2);
            const LightSample ls = sample_light(g.tris, lights, n_lights, r0, r1, r2);
            const Triangle& lt = g.tris[ls.index];
            const float V = check_visibility(g, surf.p, surf.n, ls.p);
            const V3 brdf = 1.0f / kPi * tri.color;
            const float G = geometry_term(surf.p, surf.n, ls.p, ls.n);
            const float light_pdf = 1.0f / (float)n_lights * 1.0f / area_of(lt);
            radiance = radiance + throughput * brdf * G * V * lt.emissive / light_pdf;
        }
        else if (mode == 9)
        {
            const Reservoir r = ris_candidates(g, surf, lights, n_lights, opt.ris_sample_count,
                                               opt.use_shadowed_target_function, rng);
            const V3 brdf = 1.0f / kPi * tri.color;
            const float G = geometry_term(surf.p, surf.n, r.s.hit_position, r.s.hit_normal);
            const float V = check_visibility(g, surf.p, surf.n, r.s.hit_position);
            const float p_hat = target_function(g, surf.p, surf.n, r.s.hit_position, r.s.hit_normal, r.s.radiance,
                                                opt.use_shadowed_target_function);
            radiance = radiance + throughput * brdf * G * V * r.s.radiance * ucw_of(r, p_hat);
        }
        const V3 wo = bounce_direction(surf, tri, rng);
        throughput = throughput * tri.color;
        ro = surf.p + 0.001f * surf.n;
        rd = wo;
    }
    accumulate(accum, px.idx, radiance, opt.accumulate);
}

// brute-force closest hit, 04_ao.cu:8-29
static bool brute_closest(const Triangle* tris, int n, V3 ro, V3 rd, Isect& is)
{
    float t = kFltMax, u, v;
    int index = -1;
    for (int i = 0; i < n; i++)
        if (ray_triangle(t, u, v, ro, rd, 0.0f, t, tris[i].v[0], tris[i].v[1], tris[i].v[2])) index = i;
    if (index < 0) return false;
    is.t = t;
    is.index = index;
    return true;
}

// 06_ao_hiprt.cu:35-91 (brute=false) and 04_ao.cu:31-88 (brute=true); N_Rays is a parameter here
static void k_ao(long tid, uint8_t* pixels, const RayGen& rg, int W, int H, const Geom& g, int n_tris, int n_rays,
                 bool brute)
{
    const Px px = pixel_of(tid, W, H);
    Pcg rng(0, hash_pcg3(px.xi, px.yi, 42));
    V3 ro, rd;
    shoot(rg, (float)px.xi / (float)W, (float)px.yi / (float)H, ro, rd);
    auto closest = [&](V3 o, V3 d, Isect& is)
    { return brute ? brute_closest(g.tris, n_tris, o, d, is) : raytrace(g, o, d, 0.0f, kFltMax, is); };
    uint8_t* out = pixels + 4 * (size_t)px.idx;
    Isect is;
    if (!closest(ro, rd, is))
    {
        out[0] = out[1] = out[2] = 32;
        out[3] = 255;
        return;
    }
    const Triangle& tri = g.tris[is.index];
    V3 n = normal_of(tri);
    if (0.0f < dot(n, rd)) n = -n;
    const V3 t0 = tangent_of(tri);
    const V3 t1 = cross(t0, n);
    const V3 ao_ro = ro + rd * is.t + n * 0.0001f;
    int n_visible = 0;
    for (int i = 0; i < n_rays; i++)
    {
        float r0, r1, r2;
        draw3(rng, r0, r1, r2);
        const V3 s = sample_hemisphere(r0, r1, r2);
        const V3 ao_rd = t0 * s.x + t1 * s.z + n * s.y;
        Isect ao;
        if (!closest(ao_ro, ao_rd, ao)) n_visible++;
    }
    const float ao = (float)n_visible / (float)n_rays;
    const uint8_t c = (uint8_t)(m_pow(ao, 1.0f / 2.2f) * 255.0f);
    out[0] = out[1] = out[2] = c;
    out[3] = 255;
}
}  // namespace port

// ======================================================================= C API (same as ref_driver.cpp)
using namespace port;
static long g_tid_begin = 0, g_tid_end = -1;

template <class F>
static void launch(int W, int H, F&& f)
{
    const long n = (long)W * H;
    const long t0 = g_tid_begin, t1 = g_tid_end < 0 ? n : (g_tid_end < n ? g_tid_end : n);
#pragma omp parallel for schedule(dynamic, 2048)
    for (long tid = t0; tid < t1; tid++) f(tid);
}
static V3 f3(const float* p) { return {p[0], p[1], p[2]}; }

extern "C"
{
    int orc_example() { return 0; }
    const char* orc_kind() { return "port"; }
    int orc_threads() { return omp_get_max_threads(); }
    void orc_set_threads(int n) { omp_set_num_threads(n); }
    void orc_set_math_mode(int m) { g_math_mode = m; }
    void orc_set_arg_order(int o) { g_arg_order = o; }
    void orc_set_range(long begin, long end) { g_tid_begin = begin; g_tid_end = end; }

    uint64_t orc_fnv1a64(const uint8_t* p, size_t n)
    {
        uint64_t h = 0xcbf29ce484222325ULL;
        for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ULL;
        return h;
    }
    // the same walk from another offset basis.  SURVEY.md sections 2.3 / 4 quote their goldens from a probe whose
    // basis was 1469598103934665603 — the standard basis 14695981039346656037 with its last digit dropped — so
    // checking those values needs that start value (tests/test_oracle_pinning.py: SURVEY_FNV_BASIS).
    uint64_t orc_fnv1a64_from(const uint8_t* p, size_t n, uint64_t basis)
    {
        uint64_t h = basis;
        for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ULL;
        return h;
    }

    uint32_t orc_hash_pcg3(uint32_t x, uint32_t y, uint32_t z) { return hash_pcg3(x, y, z); }
    uint32_t orc_hash_pcg4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return hash_pcg4(x, y, z, w); }
    void orc_pcg_probe(uint64_t seed, uint64_t seq, uint32_t* three_u32)
    {
        Pcg r(seed, seq);
        for (int i = 0; i < 3; i++) three_u32[i] = r.next_u32();
    }

    void* orc_geom_build(const Triangle* tris, int n)
    {
        Geom* g = new Geom;
        g->tris = tris;
        g->bvh = cpubvh::build((const float*)tris, sizeof(Triangle) / 4, n);
        return g;
    }
    void orc_geom_free(void* g)
    {
        if (!g) return;
        delete ((Geom*)g)->bvh;
        delete (Geom*)g;
    }
    void orc_lookat(const float* eye, const float* center, const float* up, float fovy, int W, int H, RayGen* out)
    {
        *out = lookat(f3(eye), f3(center), f3(up), fovy, W, H);
    }
    int orc_closest_hit(void* geom, const float* o, const float* d, float tmin, float tmax, float* tuv)
    {
        Isect is;
        if (!raytrace(*(Geom*)geom, f3(o), f3(d), tmin, tmax, is)) return -1;
        tuv[0] = is.t; tuv[1] = is.uv.x; tuv[2] = is.uv.y;
        return is.index;
    }
    // brute-force reference for the BVH (04_ao.cu:8-29 loop with the uv of the winner)
    int orc_closest_hit_brute(const Triangle* tris, int n, const float* o, const float* d, float tmin, float tmax,
                              float* tuv)
    {
        float t = tmax, u = 0, v = 0, bu = 0, bv = 0;
        int index = -1;
        for (int i = 0; i < n; i++)
            if (ray_triangle(t, u, v, f3(o), f3(d), tmin, t, tris[i].v[0], tris[i].v[1], tris[i].v[2]))
            {
                index = i; bu = u; bv = v;
            }
        if (index < 0) return -1;
        tuv[0] = t; tuv[1] = bu; tuv[2] = bv;
        return index;
    }

    void orc_clear(V4* buf, int W, int H) { launch(W, H, [&](long tid) { k_clear(tid, buf, W, H); }); }
    void orc_tone_mapping(uint8_t* pixels, const V4* accum, int W, int H)
    {
        launch(W, H, [&](long tid) { k_tone_mapping(tid, pixels, accum, W, H); });
    }
    void orc_raycast(int W, int H, void* geom, const Triangle*, int, const RayGen* rg, Visibility* vis)
    {
        launch(W, H, [&](long tid) { k_raycast(tid, W, H, *(Geom*)geom, *rg, vis); });
    }
    void orc_generate_candidate(int W, int H, int frame, void* geom, const Triangle*, int, const Visibility* vis,
                                const float* eye, const uint32_t* lights, int nlights, const Options* opt,
                                Reservoir* res)
    {
        launch(W, H, [&](long tid)
               { k_generate_candidate(tid, W, H, frame, *(Geom*)geom, vis, f3(eye), lights, nlights, *opt, res); });
    }
    void orc_temporal_resampling(int W, int H, int frame, void* geom, const Triangle*, int, const Visibility* vis,
                                 const float* eye, const Options* opt, const Reservoir* prev, Reservoir* res)
    {
        launch(W, H, [&](long tid) { k_temporal(tid, W, H, frame, *(Geom*)geom, vis, f3(eye), *opt, prev, res); });
    }
    // temporal resampling with reprojection into the previous frame's camera (extension, see reproject_pixel)
    void orc_temporal_resampling_reprojected(int W, int H, int frame, void* geom, const Triangle*, int, const Visibility* vis,
                                             const float* eye, const Options* opt, const RayGen* prev_cam,
                                             const Reservoir* prev, Reservoir* res)
    {
        launch(W, H, [&](long tid) { k_temporal(tid, W, H, frame, *(Geom*)geom, vis, f3(eye), *opt, prev, res, prev_cam); });
    }
    int orc_reproject_pixel(const RayGen* prev_cam, int W, int H, const float* p, int* xy)
    {
        return reproject_pixel(*prev_cam, W, H, f3(p), xy[0], xy[1]) ? 1 : 0;
    }
    void orc_save_temporal_reservoir(int W, int H, const Reservoir* src, Reservoir* dst)  // 10_restir_di.cu:239-254
    {
        launch(W, H, [&](long tid) { const int i = pixel_of(tid, W, H).idx; dst[i] = src[i]; });
    }
    void orc_spatial_resampling(int W, int H, int frame, int pass, void* geom, const Triangle*, int,
                                const Visibility* vis, const float* eye, const Options* opt, const Reservoir* in,
                                Reservoir* out)
    {
        launch(W, H, [&](long tid) { k_spatial(tid, W, H, frame, pass, *(Geom*)geom, vis, f3(eye), *opt, in, out); });
    }
    void orc_resolve(V4* accum, int W, int H, void* geom, const Triangle*, int, const Visibility* vis,
                     const float* eye, const Options* opt, const Reservoir* res)
    {
        launch(W, H, [&](long tid) { k_resolve(tid, accum, W, H, *(Geom*)geom, vis, f3(eye), *opt, res); });
    }
    // mode = 7, 8 or 9 selects the example
    void orc_path_trace_mode(int mode, int W, int H, int frame, void* geom, const uint32_t* lights, int nlights,
                             const RayGen* rg, const Options* opt, V4* accum)
    {
        launch(W, H, [&](long tid)
               { k_path_trace(tid, mode, W, H, frame, *(Geom*)geom, lights, nlights, *rg, *opt, accum); });
    }
    static int g_pt_mode = 9;
    void orc_set_example(int e) { g_pt_mode = e; }
    void orc_path_trace(int W, int H, int frame, void* geom, const Triangle*, int, const uint32_t* lights,
                        int nlights, const RayGen* rg, const Options* opt, V4* accum)
    {
        orc_path_trace_mode(g_pt_mode, W, H, frame, geom, lights, nlights, rg, opt, accum);
    }
    // brute != 0: example 04 (no BVH); else example 06
    int orc_ao_mode(int brute, uint8_t* pixels, const RayGen* rg, int W, int H, void* geom, const Triangle* tris,
                    int ntris, int n_rays)
    {
        Geom tmp{nullptr, tris};
        const Geom& g = brute ? tmp : *(Geom*)geom;
        launch(W, H, [&](long tid) { k_ao(tid, pixels, *rg, W, H, g, ntris, n_rays, brute != 0); });
        return 0;
    }
    int orc_ao(uint8_t* pixels, const RayGen* rg, int W, int H, void* geom, const Triangle* tris, int ntris,
               int n_rays)
    {
        return orc_ao_mode(g_pt_mode == 4, pixels, rg, W, H, geom, tris, ntris, n_rays);
    }
}
