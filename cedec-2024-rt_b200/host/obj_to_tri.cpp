// obj_to_tri — writes the Triangle[] bytes obj_scene.hpp produces for an OBJ (test tool; no CUDA calls).
#include <cstdio>

#include "obj_scene.hpp"

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        fprintf(stderr, "usage: %s scene.obj out.tri [mtl_basedir]\n", argv[0]);
        return 2;
    }
    try
    {
        const std::vector<crt_triangle> tris = crt::loadTrianglesFromObj(argv[1], argc > 3 ? argv[3] : "");
        FILE* f = fopen(argv[2], "wb");
        if (!f) return 1;
        fwrite(tris.data(), sizeof(crt_triangle), tris.size(), f);
        fclose(f);
        printf("%zu triangles\n", tris.size());
    }
    catch (const crt::Error& e)
    {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
