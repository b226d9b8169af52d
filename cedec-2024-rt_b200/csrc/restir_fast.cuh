// restir_fast.cuh — per-pixel bodies of the fused frame ("fast mode", crt_restir_di_frame):
// the same arithmetic as restir_pixel.cuh, in the same order, on a different data flow.
//
//   * reservoirs live in planar SoA form inside the TypedBuffer<Reservoir> allocations (SoaStore): four float4
//     planes and one float2 plane, 72 B/px, every access a 128-bit (64-bit) load or store, a warp's 8x4 pixel
//     tile touching full 128-byte lines;
//   * generate_candidate + temporal_resampling are one body: the candidate reservoir stays in registers, the
//     merged reservoir is written in place over last frame's (same pixel: no reprojection in the reference,
//     10_restir_di.cu:178), which also makes save_temporal_reservoir (:239-254) a no-op;
//   * the visibility-reuse ray (:127-131) is only emitted when this frame's candidate survives the temporal
//     merge — otherwise Reservoir::merge (reservoir.hpp:31-37) overwrites the sample, visibility included, and
//     the reference's traced value is never read;
//   * the surface point/normal the reference rebuilds from Visibility + Triangle in every kernel
//     (make_surface_info, core.hpp:188-207) is computed once and kept in a small G-buffer together with a
//     1-byte pixel class; the spatial pass rejects sky/emissive neighbours (:317-326) from that byte;
//   * tone mapping (common.cu:30-74) stays a separate sweep: inside the ray tracer's epilogue its three powf calls
//     ran once per finished ray with a handful of live lanes and cost 1.0 ms per 4K frame (profiles/r1, state h)
//     against 0.1 ms as a kernel of its own.
//
// Host-callable like restir_pixel.cuh, so tests/emu runs these bodies on the CPU against the oracle.
#pragma once
#include "restir_pixel.cuh"

namespace crt
{
struct u2 { uint32_t x, y; };
CRT_HD void store_u4(void* p, u4 v)
{
#if defined(__CUDA_ARCH__)
    *(uint4*)p = make_uint4(v.x, v.y, v.z, v.w);
#else
    memcpy(p, &v, 16);
#endif
}
CRT_HD u2 load_u2(const void* p)
{
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg((const uint2*)p);
    return {v.x, v.y};
#else
    u2 v;
    memcpy(&v, p, 8);
    return v;
#endif
}
CRT_HD void store_u2(void* p, u2 v)
{
#if defined(__CUDA_ARCH__)
    *(uint2*)p = make_uint2(v.x, v.y);
#else
    memcpy(p, &v, 8);
#endif
}

// Planar reservoir storage over n pixels (n * 72 bytes used of the n * 76 the reference allocates):
//   plane 0 @ 0      float4  hit_position.xyz, ucw
//   plane 1 @ 16 n   float4  hit_normal.xyz, radiance.x
//   plane 2 @ 32 n   float4  radiance.y, radiance.z, M | visibility << 31, w_sum
//   plane 3 @ 48 n   float4  origin_position.xyz, origin_normal.x
//   plane 4 @ 64 n   float2  origin_normal.y, origin_normal.z
// resolve reads planes 0-2 only; a neighbour merge reads everything but uses no w_sum.
constexpr int kSoaPlanes = 5;
constexpr int kHaloRows = 87;  // rows a spatial pass can reach beyond a slab: |offset| <= 86.4 px (10_restir_di.cu:309-313)
CRT_HD size_t soa_plane_offset(int plane, size_t n) { return (size_t)plane * 16u * n; }
CRT_HD size_t soa_plane_elem(int plane) { return plane < 4 ? 16u : 8u; }
constexpr uint32_t kVisBit = 0x80000000u;

struct SoaStore
{
    char* base;
    size_t n;
    CRT_HD char* plane(int p, int idx) const { return base + soa_plane_offset(p, n) + (size_t)idx * soa_plane_elem(p); }
    CRT_HD Res load(int idx) const
    {
        const u4 a = load_u4(plane(0, idx)), b = load_u4(plane(1, idx)), c = load_u4(plane(2, idx)), d = load_u4(plane(3, idx));
        const u2 e = load_u2(plane(4, idx));
        Res r;
        r.s.hp = {u2f(a.x), u2f(a.y), u2f(a.z)};
        r.ucw = u2f(a.w);
        r.s.hn = {u2f(b.x), u2f(b.y), u2f(b.z)};
        r.s.rad = {u2f(b.w), u2f(c.x), u2f(c.y)};
        r.M = (int)(c.z & ~kVisBit);
        r.s.vis = c.z >> 31;
        r.w_sum = u2f(c.w);
        r.s.op = {u2f(d.x), u2f(d.y), u2f(d.z)};
        r.s.on = {u2f(d.w), u2f(e.x), u2f(e.y)};
        return r;
    }
    // what resolve reads (10_restir_di.cu:431-447): hit position/normal, radiance, ucw
    CRT_HD Res load_shading(int idx) const
    {
        const u4 a = load_u4(plane(0, idx)), b = load_u4(plane(1, idx)), c = load_u4(plane(2, idx));
        Res r = empty_res();
        r.s.hp = {u2f(a.x), u2f(a.y), u2f(a.z)};
        r.ucw = u2f(a.w);
        r.s.hn = {u2f(b.x), u2f(b.y), u2f(b.z)};
        r.s.rad = {u2f(b.w), u2f(c.x), u2f(c.y)};
        return r;
    }
    CRT_HD void store(int idx, const Res& r) const
    {
        store_u4(plane(0, idx), u4{f2u(r.s.hp.x), f2u(r.s.hp.y), f2u(r.s.hp.z), f2u(r.ucw)});
        store_u4(plane(1, idx), u4{f2u(r.s.hn.x), f2u(r.s.hn.y), f2u(r.s.hn.z), f2u(r.s.rad.x)});
        store_u4(plane(2, idx), u4{f2u(r.s.rad.y), f2u(r.s.rad.z), ((uint32_t)r.M & ~kVisBit) | (r.s.vis ? kVisBit : 0u), f2u(r.w_sum)});
        store_u4(plane(3, idx), u4{f2u(r.s.op.x), f2u(r.s.op.y), f2u(r.s.op.z), f2u(r.s.on.x)});
        store_u2(plane(4, idx), u2{f2u(r.s.on.y), f2u(r.s.on.z)});
    }
    // word holding M | visibility << 31 (the shadow-ray kernel sets the bit for unoccluded candidates)
    CRT_HD uint32_t* mvis_word(int idx) const { return (uint32_t*)plane(2, idx) + 2; }
};

// G-buffer of the fused frame: surface point and shading normal of the primary hit, and the pixel class
enum : uint8_t { kPixSkip = 0, kPixDiffuse = 1 };  // skip = sky or emissive: no reservoir work (10_restir_di.cu:62-70)
struct GBuf
{
    char* g0;      // float4 per pixel: p.xyz, n.x
    char* g1;      // float2 per pixel: n.y, n.z
    uint8_t* cls;  // kPix*
    CRT_HD Surf load(int idx) const
    {
        const u4 a = load_u4(g0 + (size_t)idx * 16);
        const u2 b = load_u2(g1 + (size_t)idx * 8);
        return Surf{{u2f(a.x), u2f(a.y), u2f(a.z)}, {u2f(a.w), u2f(b.x), u2f(b.y)}};
    }
    CRT_HD void store(int idx, const Surf& s) const
    {
        store_u4(g0 + (size_t)idx * 16, u4{f2u(s.p.x), f2u(s.p.y), f2u(s.p.z), f2u(s.n.x)});
        store_u2(g1 + (size_t)idx * 8, u2{f2u(s.n.y), f2u(s.n.z)});
    }
    CRT_HD uint8_t pixel_class(int idx) const
    {
#if defined(__CUDA_ARCH__)
        return __ldg(cls + idx);
#else
        return cls[idx];
#endif
    }
};

// Multi-GPU row slabs (slab_p2p.cu): the neighbouring slabs' copies of the buffer a kernel writes.  A pixel whose
// row lies within kHaloRows of the slab boundary is stored there too — the halo rows travel over NVLink from the
// producing kernel's own epilogue, overlapped with the rest of the kernel, instead of in a copy afterwards.
struct HaloPeers
{
    char* up = nullptr;         // reservoir storage of the slab above (rows yi < up_end are mirrored there)
    char* down = nullptr;       // ... of the slab below (rows yi >= down_begin)
    uint8_t* up_cls = nullptr;  // pixel-class planes (candidate kernel only)
    uint8_t* down_cls = nullptr;
    int up_end = 0, down_begin = 0;
    CRT_HD void mirror(const Pix& px, size_t n, const Res& r) const
    {
        if (up && px.yi < up_end) SoaStore{up, n}.store(px.idx, r);
        if (down && px.yi >= down_begin) SoaStore{down, n}.store(px.idx, r);
    }
    CRT_HD void mirror_class(const Pix& px, uint8_t c) const
    {
        if (up_cls && px.yi < up_end) up_cls[px.idx] = c;
        if (down_cls && px.yi >= down_begin) down_cls[px.idx] = c;
    }
};

// ---- generate_candidate (10_restir_di.cu:36-135) + temporal_resampling (:137-237) + save (:239-254).
// `temporal` holds last frame's post-temporal reservoirs on entry and this frame's on exit.
// Requires !opt.shadowed (the fused frame falls back to the per-kernel path otherwise).
// Returns the visibility-reuse ray if one has to be traced; the reservoir is then stored with
// visibility = false and the tracer sets the bit for an unoccluded ray.
template <class M, class L>
CRT_HD DeferredRay px_candidate_temporal(const Pix& px, int frame, const Bvh& bvh, const float* tris60,
                                         const crt_visibility* vis, f3 eye, const L& lights, const Opt& opt_in,
                                         const SoaStore& temporal, const GBuf& g, const HaloPeers& peers = HaloPeers())
{
    Opt opt = opt_in;
    opt.shadowed = false;  // compile-time constant here: no traversal code inside the target function
    DeferredRay ray{false, {0, 0, 0}, {0, 0, 0}};
    const Vis v = load_vis(vis, px.idx);
    bool skip = v.index == -1;
    TriRef tri{tris60};
    if (!skip)
    {
        tri = tri_at(tris60, v.index);
        skip = has_emission(tri.emissive());
    }
    if (skip)
    {
        g.cls[px.idx] = kPixSkip;
        peers.mirror_class(px, kPixSkip);     // (the reservoirs of skipped pixels are never read by a neighbour)
        temporal.store(px.idx, empty_res());  // what the reference leaves there: Reservoir{} copied by save_temporal
        return ray;
    }
    g.cls[px.idx] = kPixDiffuse;
    peers.mirror_class(px, kPixDiffuse);
    Pcg rng(hash_pcg4(px.xi, px.yi, frame, 0), 0);
    const Surf surf = surface_from_visibility(tri, v.u, v.v, eye);
    g.store(px.idx, surf);
    Res r = ris_candidates(bvh, lights, surf, opt.ris_count, false, rng);
    r.ucw = ucw_of(r, target_function(bvh, surf.p, surf.n, r.s.hp, r.s.hn, r.s.rad, false));
    bool candidate_survives = true;
    if (opt.temporal)
    {
        Pcg rng_t(hash_pcg4(px.xi, px.yi, frame, 1), 0);
        candidate_survives = !temporal_merge<M>(bvh, surf, eye, opt, temporal.load(px.idx), r, rng_t);
    }
    if (opt.reuse && candidate_survives) ray = visibility_ray(surf.p, surf.n, r.s.hp);
    temporal.store(px.idx, r);
    peers.mirror(px, temporal.n, r);  // visibility still pending for a deferred ray: the tracer mirrors the final word
    return ray;
}

// ---- spatial_resampling (10_restir_di.cu:256-388) with opt.spatial == true (the disabled form is a copy the
// fused frame skips)
template <class M>
CRT_HD void px_spatial_fast(const Pix& px, int W, int H, int frame, int pass, const Bvh& bvh, f3 eye, const Opt& opt_in,
                            const SoaStore& in, const SoaStore& out, const GBuf& g, const HaloPeers& peers = HaloPeers())
{
    Opt opt = opt_in;
    opt.shadowed = false;
    if (g.pixel_class(px.idx) == kPixSkip) return;  // output left untouched, as in the reference
    Pcg rng(hash_pcg4(px.xi, px.yi, frame, 2 + pass), 0);
    const Surf surf = g.load(px.idx);
    Res r = in.load(px.idx);
    for (int k = 0; k < opt.spatial_count; ++k)
    {
        int x, y;
        spatial_neighbour<M>(px.xi, px.yi, opt.radius, rng, x, y);
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        if (x == px.xi && y == px.yi) continue;
        const int pid = x + (H - y - 1) * W;
        if (g.pixel_class(pid) == kPixSkip) continue;
        spatial_merge<M>(bvh, surf, eye, opt, in.load(pid), r, rng);
    }
    r.ucw = ucw_of(r, target_function(bvh, surf.p, surf.n, r.s.hp, r.s.hn, r.s.rad, false));
    out.store(px.idx, r);
    peers.mirror(px, out.n, r);
}

// ---- resolve (10_restir_di.cu:390-459).  Sky/emissive pixels are finished here; for the others the shadow ray
// and the shading factors are returned (see px_resolve).
template <class RS>
CRT_HD DeferredRay px_resolve_fast(const Pix& px, crt_float4* accum, const float* tris60,
                                   const crt_visibility* vis, const RS& res, const GBuf& g, DeferredShade& shade)
{
    DeferredRay ray{false, {0, 0, 0}, {0, 0, 0}};
    const Vis v = load_vis(vis, px.idx);
    if (g.pixel_class(px.idx) == kPixSkip)
    {
        f3 c{0.0f, 0.0f, 0.0f};
        if (v.index != -1) c = tri_at(tris60, v.index).emissive();
        accum[px.idx] = {c.x, c.y, c.z, 1.0f};  // assigned even when accumulating
        return ray;
    }
    const Surf surf = g.load(px.idx);
    const Res r = res.load_shading(px.idx);
    shade.bg = (kInvPi * tri_at(tris60, v.index).color()) * geometry_term(surf.p, surf.n, r.s.hp, r.s.hn);
    shade.rad = r.s.rad;
    shade.ucw = r.ucw;
    return visibility_ray(surf.p, surf.n, r.s.hp);
}
// AoS <-> SoA conversion of one reservoir (crt_reservoir_export_aos / crt_reservoir_import_aos)
CRT_HD void soa_to_aos(const SoaStore& s, crt_reservoir* aos, int idx) { AosStore{aos}.store(idx, s.load(idx)); }
CRT_HD void aos_to_soa(const crt_reservoir* aos, const SoaStore& s, int idx)
{
    s.store(idx, AosStore{const_cast<crt_reservoir*>(aos)}.load(idx));
}
}  // namespace crt
