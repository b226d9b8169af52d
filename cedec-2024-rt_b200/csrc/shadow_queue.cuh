// shadow_queue.cuh — wavefront shadow rays: a compacted ray queue and a persistent-warp any-hit kernel.
//
// check_visibility() (common/raytrace.hpp:45-52) is called once per diffuse pixel by generate_candidate
// (visibility reuse, 10_restir_di.cu:127-131) and once by resolve (:443-444).  The rays go from coherent
// pixels to randomly chosen lights anywhere in the scene, so their walks have very different lengths:
// traced inside the per-pixel kernels, a warp ran at 9 of 32 lanes (profiles/r1/source_a_k_resolve.txt).
// Here the per-pixel kernel only *emits* the ray (warp-aggregated append, 64-byte record), and a
// persistent kernel — one resident set of warps per SM, rays fetched from a global counter — walks them;
// a lane that finishes takes the next ray as soon as fewer than kRefillThreshold lanes of its warp are
// still busy, so the warps stay full.  The result is written where the per-pixel code would have put it
// (the reservoir's visibility byte, or the shaded pixel), bit for bit.
#pragma once
#include "restir_fast.cuh"

namespace crt
{
// 64-byte ray record: four 16-byte words
struct alignas(16) ShadowRay
{
    float ox, oy, oz;   // origin  p0 + 1e-3 n0
    uint32_t pix;       // pixel index the result belongs to
    float dx, dy, dz;   // direction p1 - p0 (t in [0, 0.99])
    float ucw;          // resolve: reservoir.ucw
    float bgx, bgy, bgz;  // resolve: brdf * G
    float pad0;
    float rx, ry, rz;   // resolve: reservoir.sample.radiance
    float pad1;
};
static_assert(sizeof(ShadowRay) == 64, "ShadowRay must be 64 bytes");

struct ShadowQueue
{
    ShadowRay* rays;
    uint32_t* count;   // rays appended so far
    uint32_t* next;    // next ray to fetch (persistent kernel)
    uint32_t capacity;
    unsigned long long* total;  // rays traced since crt_init: [0] visibility reuse, [1] resolve (crt_shadow_rays_traced);
                                // [2] visibility-reuse rays settled at emission by the own-triangle pre-test
    float tmax = 0.99f;         // the segment is t in [0, tmax]: 0.99 for check_visibility (raytrace.hpp:45-52), FLT_MAX for AO rays
    uint32_t stride4 = 4;       // record size in 16-byte words: 4 = ShadowRay, 2 = compact (origin + pixel, direction) for AO rays
};

// A record whose pixel word is kHoleRecord holds no ray: a producer that reserves a fixed block of records per warp and
// then finds some of its rays decided already (kernels_paths.cu) marks the unused ones; the tracers drop them at the fetch.
constexpr uint32_t kHoleRecord = 0xffffffffu;

#ifndef CRT_REFILL
#define CRT_REFILL 24
#endif
constexpr int kRefillThreshold = CRT_REFILL;  // refetch when fewer lanes than this are still walking
// Short queues (one GPU's slab of a multi-GPU frame) refill earlier: measured on a 3840x272 frame 0.245 -> 0.234 ms
// (visibility reuse) and 0.295 -> 0.282 ms (resolve) at 28 against 24; at 4K 28 is no better (profiles/r2/tuning.txt)
constexpr int kRefillThresholdShort = 28;
constexpr uint32_t kShortQueue = 2000000u;

#if defined(__CUDACC__)
// warp-aggregated append; every lane of the warp must call it (has = whether this lane emits a ray)
// decided: this lane's ray was settled by the own-triangle pre-test (counted in total[2], crt_rays_decided_at_emission)
__device__ __forceinline__ void queue_push(const ShadowQueue& q, bool has, const ShadowRay& ray, bool decided = false)
{
    const unsigned mask = __ballot_sync(0xffffffffu, has);
    const unsigned dmask = __ballot_sync(0xffffffffu, decided);
    const int lane = threadIdx.x & 31;
    if (dmask && lane == 0) atomicAdd(q.total + 2, (unsigned long long)__popc(dmask));
    if (mask == 0) return;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(q.count, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (has)
    {
        const uint32_t slot = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        float4* dst = (float4*)(q.rays + slot);
        dst[0] = make_float4(ray.ox, ray.oy, ray.oz, __uint_as_float(ray.pix));
        dst[1] = make_float4(ray.dx, ray.dy, ray.dz, ray.ucw);
        dst[2] = make_float4(ray.bgx, ray.bgy, ray.bgz, 0.0f);
        dst[3] = make_float4(ray.rx, ray.ry, ray.rz, 0.0f);
    }
}

// The same append with one atomic per 256-thread block instead of one per warp (every thread of the block must call it, the
// block must be 8 warps): the warps' counts meet in shared memory.  259 k warps appending to one counter within the 0.37 ms
// of k_resolve_fast are one same-address atomic every 1.4 ns, which is as fast as the L2 takes them: with one per block the
// kernel needs 0.25 ms (profiles/r2/tuning.txt, batch 30).
__device__ __forceinline__ void queue_push_block(const ShadowQueue& q, bool has, const ShadowRay& ray, bool decided = false)
{
    __shared__ uint32_t s_cnt[8], s_dec[8], s_base;
    const unsigned mask = __ballot_sync(0xffffffffu, has);
    const unsigned dmask = __ballot_sync(0xffffffffu, decided);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        s_cnt[warp] = (uint32_t)__popc(mask);
        s_dec[warp] = (uint32_t)__popc(dmask);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t total = 0, dec = 0;
#pragma unroll
        for (int w = 0; w < 8; w++)
        {
            const uint32_t c = s_cnt[w];
            s_cnt[w] = total;  // exclusive prefix
            total += c;
            dec += s_dec[w];
        }
        s_base = total ? atomicAdd(q.count, total) : 0u;
        if (dec) atomicAdd(q.total + 2, (unsigned long long)dec);
    }
    __syncthreads();
    if (has)
    {
        const uint32_t slot = s_base + s_cnt[warp] + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        float4* dst = (float4*)(q.rays + slot);
        dst[0] = make_float4(ray.ox, ray.oy, ray.oz, __uint_as_float(ray.pix));
        dst[1] = make_float4(ray.dx, ray.dy, ray.dz, ray.ucw);
        dst[2] = make_float4(ray.bgx, ray.bgy, ray.bgz, 0.0f);
        dst[3] = make_float4(ray.rx, ray.ry, ray.rz, 0.0f);
    }
}

enum
{
    kEpiReservoirVisibility = 0,  // reservoirs[pix].sample.visibility = !occluded   (generate_candidate, AoS)
    kEpiResolve = 1,              // accumulation[pix] (+)= brdf*G*V*radiance*ucw     (resolve)
    kEpiSoaVisibility = 2,        // fused frame: set the visibility bit of the reservoir record (restir_fast.cuh)
    kEpiCountVisible = 3,         // 06_ao: visible_count[pix] += 1 for an unoccluded ray (06_ao_hiprt.cu:78-82)
    kEpiBitmask = 4               // 08_nee / 09_ris (kernels_paths.cu): the record's word packs path | ray << 27; bit `ray` of
                                  // visible_count[path] is set for an unoccluded ray
};

struct ShadowSink
{
    crt_reservoir* reservoirs;  // kEpiReservoirVisibility
    crt_float4* accumulation;   // kEpiResolve*
    int accumulate;
    uint32_t* soa_plane0;       // kEpiSoaVisibility: plane 0 of the reservoir storage (flags | M in word kMWord of 8)
    // multi-GPU slabs: plane 0 of the neighbours' copies; pixel indices below up_end_idx belong... see HaloPeers.
    // Bottom-up storage: rows yi < up_end are the pixel indices >= up_first_idx, rows yi >= down_begin those < down_end_idx
    uint32_t* up_plane0 = nullptr;
    uint32_t* down_plane0 = nullptr;
    uint32_t up_first_idx = 0, down_end_idx = 0;
    uint32_t* visible_count = nullptr;  // kEpiCountVisible
};

// the shading factors stay in the queue record until the ray is decided (keeps the walk's register count down)
template <int EPI>
__device__ __forceinline__ void shadow_epilogue(const ShadowSink& sink, const ShadowRay* rec, uint32_t pix,
                                                bool occluded)
{
    if (EPI == kEpiCountVisible)
    {
        if (!occluded) atomicAdd(sink.visible_count + pix, 1u);
    }
    else if (EPI == kEpiBitmask)
    {
        if (!occluded) atomicOr(sink.visible_count + (pix & 0x07ffffffu), 1u << (pix >> 27));
    }
    else if (EPI == kEpiReservoirVisibility)
    {
        // word 15 of the 19-word Reservoir holds `bool visibility` (+ 3 padding bytes)
        ((uint32_t*)(sink.reservoirs + pix))[15] = occluded ? 0u : 1u;
    }
    else if (EPI == kEpiSoaVisibility)
    {
        // the reservoir was stored with visibility = false; only this thread touches the word now
        if (!occluded)
        {
            const size_t at = (size_t)pix * 8 + kMWord;
            const uint32_t w = sink.soa_plane0[at] | kVisBit;
            sink.soa_plane0[at] = w;
            if (sink.up_plane0 && pix >= sink.up_first_idx) sink.up_plane0[at] = w;
            if (sink.down_plane0 && pix < sink.down_end_idx) sink.down_plane0[at] = w;
        }
    }
    else
    {
        // radiance = brdf * G * V * sample.radiance * ucw, evaluated left to right (10_restir_di.cu:446-447)
        const float4* src = (const float4*)rec;
        const float4 w1 = __ldcs(src + 1), w2 = __ldcs(src + 2), w3 = __ldcs(src + 3);
        const float V = occluded ? 0.0f : 1.0f;
        const f3 bg{w2.x, w2.y, w2.z}, rad{w3.x, w3.y, w3.z};
        const f3 c = bg * V * rad * w1.w;
        write_accum(sink.accumulation, (int)pix, c, sink.accumulate != 0);
    }
}

// One node step of the any-hit walk (the node half of walk_step): pops / descends and leaves the hit
// triangles of the visited node in w.tmask.  Returns false when the walk has nothing left (a miss).
__device__ __forceinline__ bool shadow_node_step(const Bvh& bvh, Walk& w, WalkStack& st, const RaySetup& r, float tmax)
{
    if ((w.ng_mask >> 24) == 0)
    {
        if (w.sp == 0) return false;
        --w.sp;
        const WalkEntry top = st.e[w.sp];
        w.ng_base = top.base;
        w.ng_mask = top.mask;
    }
    const int bit = 31 - __clz((int)w.ng_mask);
    w.ng_mask &= ~(1u << bit);
    const uint32_t slot = (uint32_t)(bit - 24) ^ r.octinv;
    const uint32_t node_idx = w.ng_base + (uint32_t)__popc(w.ng_mask & 0xffu & ((1u << slot) - 1u));
    if (w.ng_mask >> 24)
    {
        st.e[w.sp] = WalkEntry{w.ng_base, w.ng_mask};
        ++w.sp;
    }
    uint32_t imask;
    const uint32_t hits = intersect_node(bvh, node_idx, r, 0.0f, tmax, w.ng_base, w.tri_base, imask);
    w.ng_mask = (hits & 0xff000000u) | imask;
    w.tmask = hits & 0x00ffffffu;
    return true;
}

constexpr int kShadowWarps = 4;    // warps per block of the persistent kernel

// Persistent any-hit walk over the queue.  Launch with (resident blocks per SM) x (SM count) blocks.
//
// Each iteration has two warp-wide phases:
//   node phase      every walking lane takes one node step of its own ray (8 child boxes);
//   triangle phase  the (ray, triangle) pairs the node steps produced — a few lanes own a few triangles each —
//                   are dealt out evenly to all 32 lanes: pair g belongs to the lane whose running pair count
//                   first exceeds g (a five-step binary search over the warp's inclusive scan, by shuffles), the
//                   lane picks the owner's (g - first pair)-th triangle, tests it against the *owner's* ray
//                   (fetched with shuffles) and the owners read the outcome of their pairs from one ballot.
//                   Everything stays in registers: no shared memory, no atomics.  (Without pooling the
//                   ~100-instruction triangle test ran with 3 of 32 lanes, profiles/r1/source_c_*; the first pooled
//                   form published pairs through shared memory with a per-lane loop that cost 19 % of the kernel's
//                   instructions at 7 lanes each, profiles/r1/lines_k_k_trace_shadow_queue_1.txt.)
// A lane whose ray is decided takes the next ray from the global counter as soon as fewer than
// kRefillThreshold lanes of its warp are still walking.
#ifndef CRT_SHADOW_MINBLOCKS
#define CRT_SHADOW_MINBLOCKS 8  // 64 registers, 32 warps/SM: measured 5 % faster than 6 (profiles/r1/tuning_j.txt)
#endif
template <int EPI>
__global__ void __launch_bounds__(kShadowWarps * 32, CRT_SHADOW_MINBLOCKS) k_trace_shadow_queue(Bvh bvh, ShadowQueue q, ShadowSink sink)
{
    const uint32_t n_rays = *q.count;
    const int refill_below = n_rays < kShortQueue ? kRefillThresholdShort : kRefillThreshold;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (EPI != kEpiCountVisible && EPI != kEpiBitmask && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(q.total + (EPI == kEpiResolve ? 1 : 0), (unsigned long long)n_rays);
    const float tmax = q.tmax;

    bool active = false, exhausted = false;
    RaySetup r;
    r.ro = r.rd = f3{0.0f, 0.0f, 0.0f};
    uint32_t pix = 0, ray_idx = 0;
    Walk w;
    WalkStack stack;
    w.sp = 0;
    w.ng_base = w.ng_mask = w.tri_base = w.tmask = 0;

    for (;;)
    {
        // ---- refill idle lanes from the global counter
        if (!exhausted)
        {
            const unsigned idle = __ballot_sync(full, !active);
            if (idle)
            {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(q.next, (uint32_t)__popc(idle));
                base = __shfl_sync(full, base, leader);
                if (!active)
                {
                    const uint32_t my = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                    if (my < n_rays)
                    {
                        const float4* src = (const float4*)q.rays + (size_t)my * q.stride4;
                        const float4 w0 = __ldg(src), w1 = __ldg(src + 1);
                        pix = __float_as_uint(w0.w);
                        ray_idx = my;
                        if (pix != kHoleRecord)
                        {
                        // visibility-reuse rays aim at freshly sampled lights: 93 % are occluded, mostly next to the light, so
                        // their walk starts at the far end (bvh.cuh: setup_ray); resolve rays aim at samples that survived
                        // the resampling — three quarters are clear, and for the rest the near end finds the blocker sooner
                        r = setup_ray(f3{w0.x, w0.y, w0.z}, f3{w1.x, w1.y, w1.z}, EPI == kEpiSoaVisibility || EPI == kEpiReservoirVisibility || EPI == kEpiBitmask);
                        walk_begin(w, r);
                        active = true;
                        }
                    }
                }
                if (base + (uint32_t)__popc(idle) >= n_rays) exhausted = true;  // warp-uniform
            }
        }
        unsigned act = __ballot_sync(full, active);
        if (act == 0)
        {
            if (exhausted) break;
            continue;  // the fetch brought nothing but hole records: fetch again
        }

        // ---- walk until the warp thins out
        for (;;)
        {
            // node phase
            bool missed = false;
            if (active) missed = !shadow_node_step(bvh, w, stack, r, tmax);
            // triangle phase: inclusive scan of the per-lane pair counts
            const uint32_t own_mask = active ? w.tmask : 0u;
            const uint32_t cnt = (uint32_t)__popc(own_mask);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t v = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t total = __shfl_sync(full, incl, 31);
            const uint32_t excl = incl - cnt;  // this lane's first pair
            bool occluded = false;
            for (uint32_t base = 0; base < total; base += 32)
            {
                const uint32_t g = base + (uint32_t)lane;  // the pair this lane tests
                int owner = 0;                             // = number of lanes whose pairs all lie before g
#pragma unroll
                for (int step = 16; step; step >>= 1)
                {
                    const uint32_t v = __shfl_sync(full, incl, owner + step - 1);
                    if (v <= g) owner += step;
                }
                const bool valid = g < total;
                const uint32_t first = __shfl_sync(full, excl, owner);
                uint32_t m = __shfl_sync(full, own_mask, owner);
                const uint32_t tri_base = __shfl_sync(full, w.tri_base, owner);
                RaySetup o;
                o.ro.x = __shfl_sync(full, r.ro.x, owner);
                o.ro.y = __shfl_sync(full, r.ro.y, owner);
                o.ro.z = __shfl_sync(full, r.ro.z, owner);
                o.rd.x = __shfl_sync(full, r.rd.x, owner);
                o.rd.y = __shfl_sync(full, r.rd.y, owner);
                o.rd.z = __shfl_sync(full, r.rd.z, owner);
                bool hit = false;
                if (valid)
                {
                    // drop the owner's earlier triangles: the (g - first)-th set bit of its 24-bit mask, found by
                    // halving (branch-free; a clear-lowest-bit loop ran up to 23 times for the warp's slowest lane)
                    uint32_t k = g - first, shift = 0;
#pragma unroll
                    for (int width = 16; width; width >>= 1)
                    {
                        const uint32_t c = (uint32_t)__popc((m >> shift) & ((1u << width) - 1u));
                        if (k >= c)
                        {
                            k -= c;
                            shift += (uint32_t)width;
                        }
                    }
                    Hit h;
                    h.prim = -1;
                    h.t = tmax;
                    h.u = h.v = 0.0f;
                    hit = intersect_wide_tri(bvh.tris + tri_base + shift, o, 0.0f, h);  // bit `shift` of m is that triangle
                }
                // lanes [excl - base, incl - base) of this round tested this lane's pairs
                const uint32_t hits = __ballot_sync(full, hit);
                const int lo = (int)excl - (int)base, hi = (int)incl - (int)base;
                if (hi > 0 && lo < 32)
                {
                    const uint32_t below_hi = hi >= 32 ? full : ((1u << hi) - 1u);
                    const uint32_t below_lo = lo <= 0 ? 0u : ((1u << lo) - 1u);
                    occluded |= (hits & below_hi & ~below_lo) != 0u;
                }
            }
            w.tmask = 0;
            if (active && (occluded || missed))
            {
                shadow_epilogue<EPI>(sink, (const ShadowRay*)((const float4*)q.rays + (size_t)ray_idx * q.stride4), pix, occluded);
                active = false;
            }
            act = __ballot_sync(full, active);
            if (act == 0) break;
            if (!exhausted && __popc(act) < refill_below) break;
        }
    }
}

// ---- the same persistent kernel for CLOSEST hits (kernels_paths.cu: the camera and bounce rays of examples 07-09, whose
// per-thread walk ran at 7 lanes per instruction).  Records are the compact 32-byte form (origin + path, direction); the
// segment is t in [0, FLT_MAX].  Differences from the any-hit kernel: the node phase culls with the lane's current best t,
// children near the origin first; in the triangle phase every (ray, triangle) pair is tested against the best hit its
// owner had when the round began, and the owner then takes the best of the pairs that beat it — smallest t, then largest
// primitive id, the order of the reference's exhaustive loop (04_ao.cu:14-24), which does not depend on the order of the
// tests — so the result is the per-thread walk's, bit for bit.  The hit goes to the path state of kernels_paths.cu.
struct ClosestSink
{
    float4* ro_t;     // .w <- hit distance
    float4* rd_prim;  // .w <- primitive id (int bits), -1 for a miss
};
#ifndef CRT_CLOSEST_MINBLOCKS
#define CRT_CLOSEST_MINBLOCKS 6
#endif
static __global__ void __launch_bounds__(kShadowWarps * 32, CRT_CLOSEST_MINBLOCKS) k_trace_closest_queue(Bvh bvh, ShadowQueue q, ClosestSink sink)
{
    const uint32_t n_rays = *q.count;
    const int refill_below = n_rays < kShortQueue ? kRefillThresholdShort : kRefillThreshold;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;

    bool active = false, exhausted = false;
    RaySetup r;
    r.ro = r.rd = f3{0.0f, 0.0f, 0.0f};
    uint32_t pix = 0;
    Walk w;
    WalkStack stack;
    w.sp = 0;
    w.ng_base = w.ng_mask = w.tri_base = w.tmask = 0;
    Hit best;
    best.t = kFltMax;
    best.u = best.v = 0.0f;
    best.prim = -1;

    for (;;)
    {
        if (!exhausted)
        {
            const unsigned idle = __ballot_sync(full, !active);
            if (idle)
            {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(q.next, (uint32_t)__popc(idle));
                base = __shfl_sync(full, base, leader);
                if (!active)
                {
                    const uint32_t my = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                    if (my < n_rays)
                    {
                        const float4* src = (const float4*)q.rays + (size_t)my * 2;
                        const float4 w0 = __ldg(src), w1 = __ldg(src + 1);
                        pix = __float_as_uint(w0.w);
                        r = setup_ray(f3{w0.x, w0.y, w0.z}, f3{w1.x, w1.y, w1.z});
                        walk_begin(w, r);
                        best.t = kFltMax;
                        best.u = best.v = 0.0f;
                        best.prim = -1;
                        active = true;
                    }
                }
                if (base + (uint32_t)__popc(idle) >= n_rays) exhausted = true;  // warp-uniform
            }
        }
        unsigned act = __ballot_sync(full, active);
        if (act == 0) break;

        for (;;)
        {
            // node phase: one node step per walking lane, culled by its best hit so far
            bool finished = false;
            if (active) finished = !shadow_node_step(bvh, w, stack, r, best.t);
            // triangle phase: the warp's (ray, triangle) pairs dealt out over its lanes (see k_trace_shadow_queue)
            const uint32_t own_mask = active ? w.tmask : 0u;
            const uint32_t cnt = (uint32_t)__popc(own_mask);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t v = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t total = __shfl_sync(full, incl, 31);
            const uint32_t excl = incl - cnt;
            for (uint32_t base = 0; base < total; base += 32)
            {
                const uint32_t g = base + (uint32_t)lane;
                int owner = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1)
                {
                    const uint32_t v = __shfl_sync(full, incl, owner + step - 1);
                    if (v <= g) owner += step;
                }
                const bool valid = g < total;
                const uint32_t first = __shfl_sync(full, excl, owner);
                const uint32_t m = __shfl_sync(full, own_mask, owner);
                const uint32_t tri_base = __shfl_sync(full, w.tri_base, owner);
                RaySetup o;
                o.ro.x = __shfl_sync(full, r.ro.x, owner);
                o.ro.y = __shfl_sync(full, r.ro.y, owner);
                o.ro.z = __shfl_sync(full, r.ro.z, owner);
                o.rd.x = __shfl_sync(full, r.rd.x, owner);
                o.rd.y = __shfl_sync(full, r.rd.y, owner);
                o.rd.z = __shfl_sync(full, r.rd.z, owner);
                Hit h;  // the owner's best hit as the round begins: a pair only reports a hit that beats it
                h.t = __shfl_sync(full, best.t, owner);
                h.prim = __shfl_sync(full, best.prim, owner);
                h.u = h.v = 0.0f;
                bool hit = false;
                if (valid)
                {
                    uint32_t k = g - first, shift = 0;
#pragma unroll
                    for (int width = 16; width; width >>= 1)
                    {
                        const uint32_t c = (uint32_t)__popc((m >> shift) & ((1u << width) - 1u));
                        if (k >= c)
                        {
                            k -= c;
                            shift += (uint32_t)width;
                        }
                    }
                    hit = intersect_wide_tri(bvh.tris + tri_base + shift, o, 0.0f, h);
                }
                // the owner scans the lanes that tested its pairs in this round, [lo, hi), for the best of their hits:
                // smallest t, then largest primitive id.  (t >= 0 here, and -0 is folded into +0 for the comparison only.)
                const int lo = (int)excl - (int)base, hi = (int)incl - (int)base;
                const int seg_lo = lo < 0 ? 0 : lo, seg_hi = hi > 32 ? 32 : hi;
                const int seg = seg_hi > seg_lo ? seg_hi - seg_lo : 0;
                const int longest = __reduce_max_sync(full, seg);
                const uint32_t key_t = hit ? __float_as_uint(h.t + 0.0f) : 0xffffffffu;
                const uint32_t key_p = hit ? (uint32_t)h.prim : 0u;
                uint32_t win_t = 0xffffffffu, win_p = 0u;
                int win_lane = -1;
                for (int it = 0; it < longest; ++it)
                {
                    const int src = seg_lo + it < 32 ? seg_lo + it : 31;
                    const uint32_t kt = __shfl_sync(full, key_t, src), kp = __shfl_sync(full, key_p, src);
                    if (it < seg && (kt < win_t || (kt == win_t && kt != 0xffffffffu && kp > win_p)))
                    {
                        win_t = kt;
                        win_p = kp;
                        win_lane = src;
                    }
                }
                const int from = win_lane >= 0 ? win_lane : lane;
                const float wt = __shfl_sync(full, h.t, from), wu = __shfl_sync(full, h.u, from), wv = __shfl_sync(full, h.v, from);
                if (win_lane >= 0)
                {
                    best.t = wt;
                    best.u = wu;
                    best.v = wv;
                    best.prim = (int)win_p;
                }
            }
            w.tmask = 0;
            if (active && finished)
            {
                sink.ro_t[pix].w = best.t;
                sink.rd_prim[pix].w = __int_as_float(best.prim);
                active = false;
            }
            act = __ballot_sync(full, active);
            if (act == 0) break;
            if (!exhausted && __popc(act) < refill_below) break;
        }
    }
}
#endif  // __CUDACC__
}  // namespace crt
