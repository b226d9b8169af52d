// TEST INFRASTRUCTURE (oracle/_ref) — the reference's own OBJ loader
// (common/loader.hpp:11-66 + tinyobjloader 1.0.6) behind a C entry point, so
// primitive IDs and vertex bits of every scene come from the reference itself.
#include <cstdlib>
#include <cstring>

#include "common/loader.hpp"

extern "C"
{
    // returns the triangle count and a malloc'ed Triangle[] (60 B each) in *out
    long orc_load_obj(const char* obj_path, const char* mtl_dir, void** out)
    {
        std::vector<Triangle> tris = loadTrianglesFromObj(obj_path, mtl_dir);
        void* p = malloc(tris.size() * sizeof(Triangle) + 1);
        memcpy(p, tris.data(), tris.size() * sizeof(Triangle));
        *out = p;
        return (long)tris.size();
    }
    void orc_free(void* p) { free(p); }
}
