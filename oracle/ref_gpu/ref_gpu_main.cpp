// ref_gpu_main.cpp — TEST / MEASUREMENT INFRASTRUCTURE: the reference GPU comparator.
//
// A headless host for the reference's *own* Orochi/HIPRT CUDA build of examples/10_restir_di: it drives the
// reference's unmodified classes (common/shader.hpp Shader/ShaderArgument, common/typedbuffer.hpp, common/loader.hpp
// buildHiprtGeometry) and its unmodified kernel file (examples/10_restir_di/10_restir_di.cu, compiled at run time by
// hiprtBuildTraceKernels exactly as 10_restir_di.cpp:82-85 does) with the launch list of 10_restir_di.cpp:270-380,
// timed with OroStopwatch around the kernel list like 10_restir_di.cpp:254-255,382-383.  What differs from the
// reference's main(): no GLFW window (common/misc.hpp is not included), the scene comes from the staged Triangle[]
// cache (bytes of the reference loader's output) optionally tiled like BASELINE config 5, resolution and options
// come from argv, and the Visibility buffer can be dumped for the primitive-id comparison.
//
// Nothing of the product links against or includes this file; it is built by oracle/Makefile (target refgpu) into
// the git-ignored baseline/_ref/ from the sources where they lie under /root/reference.
#include <Orochi/Orochi.h>
#include <Orochi/OrochiUtils.h>
#include <hiprt/hiprt.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common/camera.hpp"
#include "common/core.hpp"
#include "common/loader.hpp"
#include "common/math.hpp"
#include "common/options.hpp"
#include "common/reservoir.hpp"
#include "common/shader.hpp"
#include "common/typedbuffer.hpp"

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
    std::string base = "./", tri_path, dump_vis, dump_candidates;
    int width = 3840, height = 2160, frames = 8, warmup = 3, nx = 1, nz = 1;
    float pitch_x = 130.0f, pitch_z = 82.0f;
    float3 cameraOrig{-0.579885f, 22.194597f, -6.567105f}, cameraLookat{5.224952f, 20.847435f, 1.431192f};
    for (int i = 1; i < argc; i++)
    {
        auto is = [&](const char* k) { return !strcmp(argv[i], k) && i + 1 < argc; };
        if (is("--base")) base = argv[++i];
        else if (is("--tri")) tri_path = argv[++i];
        else if (is("--width")) width = atoi(argv[++i]);
        else if (is("--height")) height = atoi(argv[++i]);
        else if (is("--frames")) frames = atoi(argv[++i]);
        else if (is("--warmup")) warmup = atoi(argv[++i]);
        else if (is("--tiles-x")) nx = atoi(argv[++i]);
        else if (is("--tiles-z")) nz = atoi(argv[++i]);
        else if (is("--dump-vis")) dump_vis = argv[++i];
        // Reservoir[] exactly as generate_candidate leaves it on frame 1 (10_restir_di.cu:36-135), before any other kernel
        // touches it: pins the order in which NVRTC evaluates the uniformf() arguments of sample_light (:88-90)
        else if (is("--dump-candidates")) dump_candidates = argv[++i];
        else if (!strcmp(argv[i], "--eye") && i + 3 < argc) { cameraOrig.x = (float)atof(argv[++i]); cameraOrig.y = (float)atof(argv[++i]); cameraOrig.z = (float)atof(argv[++i]); }
        else if (!strcmp(argv[i], "--lookat") && i + 3 < argc) { cameraLookat.x = (float)atof(argv[++i]); cameraLookat.y = (float)atof(argv[++i]); cameraLookat.z = (float)atof(argv[++i]); }
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (base.back() != '/') base += '/';
    auto fail = [](const char* what, int code) {
        printf("{\"impl\": \"reference-hiprt\", \"unavailable\": \"%s (code %d)\"}\n", what, code);
        fflush(stdout);
        return 0;
    };

    if (int e = oroInitialize((oroApi)(ORO_API_HIP | ORO_API_CUDA), 0)) return fail("oroInitialize failed", e);
    if (oroError e = oroInit(0)) return fail("oroInit failed", (int)e);
    oroDevice device;
    if (oroError e = oroDeviceGet(&device, 0)) return fail("oroDeviceGet failed", (int)e);
    oroCtx ctx;
    if (oroError e = oroCtxCreate(&ctx, 0, device)) return fail("oroCtxCreate failed", (int)e);
    oroCtxSetCurrent(ctx);
    oroStream stream = 0;
    oroStreamCreate(&stream);
    oroDeviceProp props;
    oroGetDeviceProperties(&props, device);
    const bool isNvidia = oroGetCurAPI(0) & ORO_API_CUDADRIVER;
    fprintf(stderr, "Device: %s  Cuda: %s\n", props.name, isNvidia ? "Yes" : "No");

    std::vector<std::string> shader_options;
    shader_options.push_back("-I" + base);
    shader_options.push_back("-I" + base + "libs/hiprt");
    shader_options.push_back("-DNO_VECTOR_OP_OVERLOAD");
    shader_options.push_back(isNvidia ? NV_ARG_LINE_INFO : AMD_ARG_LINE_INFO);

    hiprtContext hContext = 0;
    double t0 = now_s();
    hiprtError herr = hiprtCreateContext(
        HIPRT_API_VERSION, {oroGetRawCtx(ctx), oroGetRawDevice(device), isNvidia ? hiprtDeviceNVIDIA : hiprtDeviceAMD},
        hContext);
    if (herr != hiprtSuccess || !hContext) return fail("hiprtCreateContext failed", (int)herr);
    const double t_ctx = now_s() - t0;

    t0 = now_s();
    const std::string shader_path = base + "examples/10_restir_di/10_restir_di.cu";
    Shader shader(shader_path.c_str(), "10_restir_di.cu", shader_options, UseHiprt(hContext).func("raycast"));
    const double t_compile = now_s() - t0;
    fprintf(stderr, "hiprt context %.1f s, trace-kernel compile %.1f s\n", t_ctx, t_compile);

    // scene: raw Triangle[] (what loadTrianglesFromObj returns), tiled nx x nz in tile-major order
    std::vector<Triangle> triangles;
    {
        FILE* f = fopen(tri_path.c_str(), "rb");
        if (!f) return fail("cannot open the triangle cache", 0);
        fseek(f, 0, SEEK_END);
        const size_t n = (size_t)ftell(f) / sizeof(Triangle);
        fseek(f, 0, SEEK_SET);
        std::vector<Triangle> one(n);
        if (fread(one.data(), sizeof(Triangle), n, f) != n) return fail("short read of the triangle cache", 0);
        fclose(f);
        triangles.reserve(n * nx * nz);
        for (int iz = 0; iz < nz; iz++)
            for (int ix = 0; ix < nx; ix++)
                for (size_t i = 0; i < n; i++)
                {
                    Triangle t = one[i];
                    for (int k = 0; k < 3; k++)
                    {
                        t.vertices[k].x += (float)ix * pitch_x;
                        t.vertices[k].z += (float)iz * pitch_z;
                    }
                    triangles.push_back(t);
                }
    }
    std::vector<uint32_t> light_indices;
    for (size_t i = 0; i < triangles.size(); ++i)
    {
        const Triangle& t = triangles[i];
        if (t.emissive.x > 0.0f || t.emissive.y > 0.0f || t.emissive.z > 0.0f) light_indices.push_back((uint32_t)i);
    }
    fprintf(stderr, "triangles: %zu  lights: %zu\n", triangles.size(), light_indices.size());

    const size_t n_px = (size_t)width * height;
    TypedBuffer<uint8_t> pixel_buffer(TYPED_BUFFER_DEVICE);
    pixel_buffer.allocate(4 * n_px);
    TypedBuffer<float4> accumulation_buffer(TYPED_BUFFER_DEVICE);
    accumulation_buffer.allocate(n_px);
    TypedBuffer<Visibility> visibility_buffer(TYPED_BUFFER_DEVICE);
    visibility_buffer.allocate(n_px);
    TypedBuffer<Reservoir> reservoir_buffer0(TYPED_BUFFER_DEVICE), reservoir_buffer1(TYPED_BUFFER_DEVICE),
        temporal_reservoir_buffer(TYPED_BUFFER_DEVICE);
    reservoir_buffer0.allocate(n_px);
    reservoir_buffer1.allocate(n_px);
    temporal_reservoir_buffer.allocate(n_px);
    // the reference leaves these uninitialised (10_restir_di.cpp:121-122); zero = "no history" 
    oroMemsetD8((oroDeviceptr)temporal_reservoir_buffer.data(), 0, temporal_reservoir_buffer.bytes());
    oroMemsetD8((oroDeviceptr)reservoir_buffer0.data(), 0, reservoir_buffer0.bytes());
    oroMemsetD8((oroDeviceptr)reservoir_buffer1.data(), 0, reservoir_buffer1.bytes());

    TypedBuffer<Triangle> triangle_buffer(TYPED_BUFFER_DEVICE);
    triangle_buffer.allocate(triangles.size());
    oroMemcpyHtoD((oroDeviceptr)triangle_buffer.data(), triangles.data(), triangle_buffer.bytes());
    TypedBuffer<uint32_t> light_buffer(TYPED_BUFFER_DEVICE);
    light_buffer.allocate(light_indices.size());
    oroMemcpyHtoD((oroDeviceptr)light_buffer.data(), light_indices.data(), light_buffer.bytes());

    t0 = now_s();
    hiprtGeometry geom = buildHiprtGeometry(hContext, triangles);
    const double t_build = now_s() - t0;
    if (!geom) return fail("buildHiprtGeometry returned a null geometry", 0);
    fprintf(stderr, "hiprt geometry build %.2f s\n", t_build);

    Options options;  // BASELINE config 5: temporal + spatial (5 neighbours, r = 30, 3 passes) + visibility reuse, accumulate
    options.accumulate = true;
    options.use_temporal_resampling = true;
    options.use_spatial_resampling = true;
    const int grid = ceiling_div(width * height, 256);

    shader.launch("clear", ShaderArgument().ptr(&accumulation_buffer).value(width).value(height), grid, 1, 1, 256, 1, 1,
                  stream);
    oroStreamSynchronize(stream);

    std::vector<float> ms;
    int frame = 0;
    for (int it = 0; it < warmup + frames; it++)
    {
        frame++;
        RayGenerator rayGen;
        rayGen.lookat(cameraOrig, cameraLookat, float3{0.0f, 1.0f, 0.0f}, 3.14159265358979323846f / 4.0f, width, height);
        OroStopwatch sw(stream);
        sw.start();
        shader.launch("raycast",
                      ShaderArgument().value(width).value(height).value(geom).ptr(&triangle_buffer).ptr(&rayGen).ptr(&visibility_buffer),
                      grid, 1, 1, 256, 1, 1, stream);
        shader.launch("generate_candidate",
                      ShaderArgument().value(width).value(height).value(frame).value(geom).ptr(&triangle_buffer).ptr(&visibility_buffer)
                          .value(cameraOrig).ptr(&light_buffer).value(options).ptr(&reservoir_buffer0),
                      grid, 1, 1, 256, 1, 1, stream);
        if (it == 0 && !dump_candidates.empty())
        {
            oroStreamSynchronize(stream);
            TypedBuffer<Reservoir> cand = reservoir_buffer0.toHost();
            TypedBuffer<Visibility> v0 = visibility_buffer.toHost();
            FILE* f = fopen(dump_candidates.c_str(), "wb");
            if (f) { fwrite(cand.data(), sizeof(Reservoir), n_px, f); fwrite(v0.data(), sizeof(Visibility), n_px, f); fclose(f); }
        }
        shader.launch("temporal_resampling",
                      ShaderArgument().value(width).value(height).value(frame).value(geom).ptr(&triangle_buffer).ptr(&visibility_buffer)
                          .value(cameraOrig).value(options).ptr(&temporal_reservoir_buffer).ptr(&reservoir_buffer0),
                      grid, 1, 1, 256, 1, 1, stream);
        shader.launch("save_temporal_reservoir",
                      ShaderArgument().value(width).value(height).ptr(&reservoir_buffer0).ptr(&temporal_reservoir_buffer), grid, 1, 1,
                      256, 1, 1, stream);
        TypedBuffer<Reservoir>* buf_input = &reservoir_buffer0;
        TypedBuffer<Reservoir>* buf_output = &reservoir_buffer1;
        for (int k = 0; k < options.spatial_resampling_passes; ++k)
        {
            if (k != 0) std::swap(buf_input, buf_output);
            shader.launch("spatial_resampling",
                          ShaderArgument().value(width).value(height).value(frame).value(k).value(geom).ptr(&triangle_buffer)
                              .ptr(&visibility_buffer).value(cameraOrig).value(options).ptr(buf_input).ptr(buf_output),
                          grid, 1, 1, 256, 1, 1, stream);
        }
        shader.launch("resolve",
                      ShaderArgument().ptr(&accumulation_buffer).value(width).value(height).value(geom).ptr(&triangle_buffer)
                          .ptr(&visibility_buffer).value(cameraOrig).value(options).ptr(buf_output),
                      grid, 1, 1, 256, 1, 1, stream);
        shader.launch("tone_mapping", ShaderArgument().ptr(&pixel_buffer).ptr(&accumulation_buffer).value(width).value(height), grid,
                      1, 1, 256, 1, 1, stream);
        sw.stop();
        const float t = sw.getMs();
        fprintf(stderr, "frame %d: %.3f ms\n", frame, t);
        if (it >= warmup) ms.push_back(t);
    }
    oroError last = oroStreamSynchronize(stream);
    if (last != oroSuccess) return fail("a kernel of the frame loop failed", (int)last);

    // sanity of the output: mean accumulated radiance and the primitive-id image
    TypedBuffer<float4> acc = accumulation_buffer.toHost();
    double sum[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n_px; i++)
    {
        sum[0] += acc[i].x; sum[1] += acc[i].y; sum[2] += acc[i].z; sum[3] += acc[i].w;
    }
    TypedBuffer<Visibility> vis = visibility_buffer.toHost();
    size_t sky = 0;
    uint64_t h = 1469598103934665603ull;  // FNV-1a-64 of the int32 index image
    for (size_t i = 0; i < n_px; i++)
    {
        const int idx = vis[i].index;
        sky += idx == -1;
        for (int b = 0; b < 4; b++) { h ^= (uint8_t)((uint32_t)idx >> (8 * b)); h *= 1099511628211ull; }
    }
    if (!dump_vis.empty())
    {
        FILE* f = fopen(dump_vis.c_str(), "wb");
        if (f) { fwrite(vis.data(), sizeof(Visibility), n_px, f); fclose(f); }
    }
    double total = 0;
    for (float t : ms) total += t;
    const double per = ms.empty() ? 0.0 : total / ms.size();
    printf("{\"impl\": \"reference-hiprt\", \"metric\": \"ReSTIR DI 4K Mpix/s\", \"value\": %.3f, \"unit\": \"Mpix/s\", "
           "\"ms_per_step\": %.4f, \"steps\": %d, \"warmup\": %d, \"device\": \"%s\", \"width\": %d, \"height\": %d, "
           "\"triangles\": %zu, \"lights\": %zu, \"hiprt_context_s\": %.2f, \"trace_kernel_compile_s\": %.2f, "
           "\"geometry_build_s\": %.3f, \"mean_radiance\": [%.6f, %.6f, %.6f], \"mean_w\": %.4f, \"sky_pixels\": %zu, "
           "\"fnv64_primid\": \"%016llx\"}\n",
           per > 0 ? n_px / per / 1e3 : 0.0, per, (int)ms.size(), warmup, props.name, width, height, triangles.size(),
           light_indices.size(), t_ctx, t_compile, t_build, sum[0] / sum[3], sum[1] / sum[3], sum[2] / sum[3],
           sum[3] / n_px, sky, (unsigned long long)h);
    fflush(stdout);
    return 0;
}
