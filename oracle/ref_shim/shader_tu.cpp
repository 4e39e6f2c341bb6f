// shader_tu.cpp - one translation unit per reference shader.  TEST INFRASTRUCTURE ONLY.
//
// Compiled once per shader by oracle/Makefile with
//   -DSHADER_FILE="</root/reference/.../shader.X.comp>"   the UNMODIFIED reference source, included where it lies
//   -DSHADER_NAME="shader.X.comp"                          the name load_shader() is called with
//   -DSHADER_NS=ns_X                                       a namespace, so that every shader keeps its own
//                                                          `_Size`, `main`, `GetNeighbours`
//   -DUNIFORMS_INC="<oracle/_ref/gen/X.uniforms.inc>"      `U(type, name)` lines produced by sed from the
//                                                          shader's own `uniform type name;` declarations
// The shader's `#include "bgfx_compute.sh"` finds oracle/ref_shim/bgfx_compute.sh (already included
// below, so the guard makes it a no-op inside the namespace); `constants.sh` and `common.sh` resolve
// next to the shader, i.e. to the reference's own files.
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "bgfx_compute.sh"
#include "runtime.h"

namespace SHADER_NS {
#include SHADER_FILE
}  // namespace SHADER_NS

namespace {
// Rows [y0, y1) of one z-slice of a dispatch: gl_GlobalInvocationID is set per invocation and the shader's
// main() is called once per invocation (inlined here, so the oracle is not dominated by call overhead).
void run_rows(long y0, long y1, long nx, unsigned z) {
    natrix_ref::uvec3& id = natrix_ref::tl_global_invocation_id;
    id.z = z;
    for (long y = y0; y < y1; ++y) {
        id.y = (uint32_t)y;
        id.xy.y = (uint32_t)y;
        for (long x = 0; x < nx; ++x) {
            id.x = (uint32_t)x;
            id.xy.x = (uint32_t)x;
            SHADER_NS::main();
        }
    }
}
struct Registration {
    Registration() {
        natrix_ref::begin_program(SHADER_NAME, &run_rows);
#define U(_type, _name) natrix_ref::register_uniform(#_name, reinterpret_cast<float*>(&SHADER_NS::_name), (int)(sizeof(_type) / sizeof(float)));
#include UNIFORMS_INC
#undef U
    }
} registration;
}  // namespace
