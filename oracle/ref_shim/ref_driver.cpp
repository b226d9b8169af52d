// TEST INFRASTRUCTURE (oracle/_ref) — host driver for the reference's own kernels.
//
// The kernel TU is the reference's *unmodified* examples/NN/NN.cu compiled as
// host C++ (see oracle/Makefile); this TU includes the reference headers in
// plain host mode, provides cpu_closest_hit() behind the fake HIPRT traversal
// class, and exposes the common `orc_*` C API that tests/ and bench.py's
// cpu_baseline leg call through ctypes.  One shared object per example because
// 07/08/09 all define `path_trace` and 04/06 both define `kernelMain`.
//
// ABI note: device-side TypedBuffer<T> is
// non-copyable, so the Itanium ABI passes it by invisible reference -> the
// kernels are declared here with `const TB&`; PODs go by value.
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common/camera.hpp"
#include "common/core.hpp"
#include "common/options.hpp"
#include "common/reservoir.hpp"

#include "../cpu_bvh.h"

#ifndef REF_EXAMPLE
#error "define REF_EXAMPLE=4|6|7|8|9|10"
#endif

struct shim_uint3 { unsigned x, y, z; };
thread_local shim_uint3 threadIdx, blockIdx, blockDim;

struct TB { void* p; size_t n; };  // {T* m_data; size_t m_size:63, m_isDevice:1} (typedbuffer.hpp:16-20)
static inline TB tb(const void* p, size_t n) { return TB{(void*)p, n | (1ull << 63)}; }

static_assert(sizeof(Triangle) == 60, "Triangle");
static_assert(sizeof(Visibility) == 16, "Visibility");
static_assert(sizeof(Reservoir) == 76, "Reservoir");
static_assert(sizeof(Options) == 48, "Options");
static_assert(sizeof(RayGenerator) == 36, "RayGenerator");

// ---- math hooks (cuda_shim.h): 0 = glibc float functions (what the reference
// compiled as host C++ really does), 1 = correctly rounded via double (the mode
// the CUDA path can reproduce bit for bit).
static int g_math_mode = 0;
extern "C" float shim_logf(float x) { return g_math_mode ? (float)log((double)x) : logf(x); }
extern "C" float shim_expf(float x) { return g_math_mode ? (float)exp((double)x) : expf(x); }
extern "C" float shim_powf(float x, float y) { return g_math_mode ? (float)pow((double)x, (double)y) : powf(x, y); }
extern "C" float shim_sinf(float x) { return g_math_mode ? (float)sin((double)x) : sinf(x); }
extern "C" float shim_cosf(float x) { return g_math_mode ? (float)cos((double)x) : cosf(x); }

// ---- geometry handle: what sits in the hiprtGeometry argument slot
struct Geom
{
    cpubvh::Bvh* bvh;
    const Triangle* tris;
};

// fake-HIPRT types, same field offsets as oracle/ref_shim/hiprt/hiprt_device.h
struct ShimRay { float3 origin; float minT; float3 direction; float maxT; };
struct ShimHit { unsigned primID; float2 uv; float3 normal; float t; };

ShimHit cpu_closest_hit(void* g, const ShimRay& r) asm("_Z15cpu_closest_hitPvRK8hiprtRay");
ShimHit cpu_closest_hit(void* g, const ShimRay& r)
{
    const Geom* geom = (const Geom*)g;
    const Triangle* tris = geom->tris;
    const float3 ro = r.origin, rd = r.direction;
    const float o[3] = {ro.x, ro.y, ro.z}, d[3] = {rd.x, rd.y, rd.z};
    cpubvh::Hit h = cpubvh::trace(*geom->bvh, o, d, r.minT, r.maxT,
                                  [&](int prim, float tmin, float tmax, float& t, float& u, float& v)
                                  {
                                      const Triangle& tri = tris[prim];
                                      return intersect_ray_triangle(&t, &u, &v, ro, rd, tmin, tmax, tri.vertices[0],
                                                                    tri.vertices[1], tri.vertices[2]);
                                  });
    ShimHit out;
    out.primID = h.prim < 0 ? ~0u : (unsigned)h.prim;
    out.uv = {h.u, h.v};
    out.normal = {0, 0, 0};
    out.t = h.prim < 0 ? -1.0f : h.t;
    return out;
}

// ---- the reference kernels (defined in the kernel TU)
extern "C"
{
#if REF_EXAMPLE == 10
    void clear(const TB&, int, int);
    void tone_mapping(const TB& pixels, const TB& accum, int, int);
    void raycast(int, int, void* geom, const TB& tris, RayGenerator, const TB& vis);
    void generate_candidate(int, int, int frame, void*, const TB& tris, const TB& vis, float3 eye, const TB& lights,
                            Options, const TB& res);
    void temporal_resampling(int, int, int, void*, const TB&, const TB&, float3, Options, const TB& prev,
                             const TB& res);
    void save_temporal_reservoir(int, int, const TB& src, const TB& dst);
    void spatial_resampling(int, int, int frame, int pass, void*, const TB&, const TB&, float3, Options,
                            const TB& in, const TB& out);
    void resolve(const TB& accum, int, int, void*, const TB&, const TB&, float3, Options, const TB& res);
#elif REF_EXAMPLE == 9 || REF_EXAMPLE == 8
    void clear(const TB&, int, int);
    void tone_mapping(const TB& pixels, const TB& accum, int, int);
    void path_trace(int, int, int frame, void*, const TB& tris, const TB& lights, RayGenerator, Options,
                    const TB& accum);
#elif REF_EXAMPLE == 7
    void clear(const TB&, int, int);
    void tone_mapping(const TB& pixels, const TB& accum, int, int);
    void path_trace(int, int, int frame, void*, const TB& tris, RayGenerator, Options, const TB& accum);
#elif REF_EXAMPLE == 6
    void kernelMain(const TB& pixels, RayGenerator, int, int, void* geom, const TB& tris);
#elif REF_EXAMPLE == 4
    void kernelMain(const TB& pixels, RayGenerator, int, int, const TB& tris);
#endif
}

static long g_tid_begin = 0, g_tid_end = -1;

template <class F>
static void launch(int W, int H, F&& f)
{
    // reference launch shape: grid = ceil(W*H/256), block = 256 (10_restir_di.cpp:278-279)
    const long n = (long)W * H;
    const long t0 = g_tid_begin, t1 = g_tid_end < 0 ? n : g_tid_end;
    const int b0 = (int)(t0 / 256), b1 = (int)((t1 + 255) / 256);
#pragma omp parallel for schedule(dynamic, 8)
    for (int b = b0; b < b1; b++)
    {
        blockDim = {256, 1, 1};
        blockIdx = {(unsigned)b, 0, 0};
        for (unsigned t = 0; t < 256; t++)
        {
            const long tid = (long)b * 256 + t;
            if (tid < t0 || tid >= t1) continue;
            threadIdx = {t, 0, 0};
            f();
        }
    }
}

static float3 f3(const float* p) { return {p[0], p[1], p[2]}; }

extern "C"
{
    int orc_example() { return REF_EXAMPLE; }
    const char* orc_kind() { return "reference"; }
    int orc_threads() { return omp_get_max_threads(); }
    void orc_set_threads(int n) { omp_set_num_threads(n); }
    void orc_set_math_mode(int m) { g_math_mode = m; }
    // restrict the following launches to thread ids [begin, end) (end < 0: all)
    void orc_set_range(long begin, long end) { g_tid_begin = begin; g_tid_end = end; }

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

    // RayGenerator::lookat (common/camera.hpp:11-25) evaluated by the reference code itself
    void orc_lookat(const float* eye, const float* center, const float* up, float fovy, int W, int H,
                    RayGenerator* out)
    {
        RayGenerator rg;
        rg.lookat(f3(eye), f3(center), f3(up), fovy, W, H);
        *out = rg;
    }

    // single-ray probes (closest hit through the same path the kernels use)
    int orc_closest_hit(void* geom, const float* o, const float* d, float tmin, float tmax, float* tuv)
    {
        ShimRay r{f3(o), tmin, f3(d), tmax};
        ShimHit h = cpu_closest_hit(geom, r);
        if (h.primID == ~0u) return -1;
        tuv[0] = h.t; tuv[1] = h.uv.x; tuv[2] = h.uv.y;
        return (int)h.primID;
    }

#if REF_EXAMPLE == 10 || REF_EXAMPLE == 9 || REF_EXAMPLE == 8 || REF_EXAMPLE == 7
    void orc_clear(float4* buf, int W, int H)
    {
        TB b = tb(buf, (size_t)W * H);
        launch(W, H, [&] { clear(b, W, H); });
    }
    void orc_tone_mapping(uint8_t* pixels, const float4* accum, int W, int H)
    {
        TB p = tb(pixels, (size_t)W * H * 4), a = tb(accum, (size_t)W * H);
        launch(W, H, [&] { tone_mapping(p, a, W, H); });
    }
#endif

#if REF_EXAMPLE == 10
    void orc_raycast(int W, int H, void* geom, const Triangle* tris, int ntris, const RayGenerator* rg,
                     Visibility* vis)
    {
        TB t = tb(tris, ntris), v = tb(vis, (size_t)W * H);
        launch(W, H, [&] { raycast(W, H, geom, t, *rg, v); });
    }
    void orc_generate_candidate(int W, int H, int frame, void* geom, const Triangle* tris, int ntris,
                                const Visibility* vis, const float* eye, const uint32_t* lights, int nlights,
                                const Options* opt, Reservoir* res)
    {
        TB t = tb(tris, ntris), v = tb(vis, (size_t)W * H), l = tb(lights, nlights), r = tb(res, (size_t)W * H);
        launch(W, H, [&] { generate_candidate(W, H, frame, geom, t, v, f3(eye), l, *opt, r); });
    }
    void orc_temporal_resampling(int W, int H, int frame, void* geom, const Triangle* tris, int ntris,
                                 const Visibility* vis, const float* eye, const Options* opt, const Reservoir* prev,
                                 Reservoir* res)
    {
        TB t = tb(tris, ntris), v = tb(vis, (size_t)W * H), p = tb(prev, (size_t)W * H), r = tb(res, (size_t)W * H);
        launch(W, H, [&] { temporal_resampling(W, H, frame, geom, t, v, f3(eye), *opt, p, r); });
    }
    void orc_save_temporal_reservoir(int W, int H, const Reservoir* src, Reservoir* dst)
    {
        TB s = tb(src, (size_t)W * H), d = tb(dst, (size_t)W * H);
        launch(W, H, [&] { save_temporal_reservoir(W, H, s, d); });
    }
    void orc_spatial_resampling(int W, int H, int frame, int pass, void* geom, const Triangle* tris, int ntris,
                                const Visibility* vis, const float* eye, const Options* opt, const Reservoir* in,
                                Reservoir* out)
    {
        TB t = tb(tris, ntris), v = tb(vis, (size_t)W * H), i = tb(in, (size_t)W * H), o = tb(out, (size_t)W * H);
        launch(W, H, [&] { spatial_resampling(W, H, frame, pass, geom, t, v, f3(eye), *opt, i, o); });
    }
    void orc_resolve(float4* accum, int W, int H, void* geom, const Triangle* tris, int ntris, const Visibility* vis,
                     const float* eye, const Options* opt, const Reservoir* res)
    {
        TB a = tb(accum, (size_t)W * H), t = tb(tris, ntris), v = tb(vis, (size_t)W * H), r = tb(res, (size_t)W * H);
        launch(W, H, [&] { resolve(a, W, H, geom, t, v, f3(eye), *opt, r); });
    }
#elif REF_EXAMPLE == 9 || REF_EXAMPLE == 8
    void orc_path_trace(int W, int H, int frame, void* geom, const Triangle* tris, int ntris, const uint32_t* lights,
                        int nlights, const RayGenerator* rg, const Options* opt, float4* accum)
    {
        TB t = tb(tris, ntris), l = tb(lights, nlights), a = tb(accum, (size_t)W * H);
        launch(W, H, [&] { path_trace(W, H, frame, geom, t, l, *rg, *opt, a); });
    }
#elif REF_EXAMPLE == 7
    void orc_path_trace(int W, int H, int frame, void* geom, const Triangle* tris, int ntris, const uint32_t*, int,
                        const RayGenerator* rg, const Options* opt, float4* accum)
    {
        TB t = tb(tris, ntris), a = tb(accum, (size_t)W * H);
        launch(W, H, [&] { path_trace(W, H, frame, geom, t, *rg, *opt, a); });
    }
#elif REF_EXAMPLE == 6
    // N_Rays is hard-coded to 64 in the reference (06_ao_hiprt.cu:71); n_rays must be 64 here.
    int orc_ao(uint8_t* pixels, const RayGenerator* rg, int W, int H, void* geom, const Triangle* tris, int ntris,
               int n_rays)
    {
        if (n_rays != 64) return -1;
        TB p = tb(pixels, (size_t)W * H * 4), t = tb(tris, ntris);
        launch(W, H, [&] { kernelMain(p, *rg, W, H, geom, t); });
        return 0;
    }
#elif REF_EXAMPLE == 4
    // brute force over all triangles (04_ao.cu:8-29); N_Rays hard-coded to 64 (04_ao.cu:60)
    int orc_ao(uint8_t* pixels, const RayGenerator* rg, int W, int H, void*, const Triangle* tris, int ntris,
               int n_rays)
    {
        if (n_rays != 64) return -1;
        TB p = tb(pixels, (size_t)W * H * 4), t = tb(tris, ntris);
        launch(W, H, [&] { kernelMain(p, *rg, W, H, t); });
        return 0;
    }
#endif
}
