// runtime.cpp - the handful of bgfx calls the reference's simulator makes, for the shim.
// TEST INFRASTRUCTURE ONLY (see bgfx_compute.sh).
//
// Mirrors, call for call, what natrix/core/fluid_simulator.py and demo/smooth_particles_area.py ask of
// bgfx-python (reference call sites in brackets):
//   nref_create_uniform(name)            bgfx.createUniform(name, Vec4)            [fluid_simulator.py:292-313]
//   nref_set_uniform(h, values, n)       bgfx.setUniform(h, as_void_ptr(c_float*n)) [:119-130, 315-336]
//   nref_create_buffer(bytes)            bgfx.createDynamicVertexBuffer(...)        [utils/shaders_utils.py:7-12]
//   nref_set_buffer(slot, h)             bgfx.setBuffer(slot, h, access)            [:338-355, 444-474]
//   nref_create_program(file)            bgfx.createProgram(load_shader(file, COMPUTE, root))  [:370-442]
//   nref_dispatch(prog, gx, gy, gz)      bgfx.dispatch(0, prog, gx, gy, gz)         [:181-280]
//   nref_destroy_*                       bgfx.destroy(h)                            [:476-515]
// Semantics follow the evident intent the reference relies on: uniform values and
// slot bindings persist across dispatches, dispatches execute in submission order, each one to
// completion, over gx*gy*gz work groups of the shader's NUM_THREADS size.  Buffers are zero-filled at
// creation (the reference never uploads).  Threads: OpenMP over rows of invocations -
// the shaders have no shared memory, barriers or atomics, and within one dispatch no invocation reads
// a cell another one writes (InitBoundaries writes slot 1 in place but reads nothing).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "bgfx_compute.sh"
#include "runtime.h"

namespace natrix_ref {

thread_local uvec3 tl_global_invocation_id;

struct UniformTarget {
    std::string name;
    float* dst;
    int components;
};
struct Program {
    std::string name;
    RowsFn rows;
    int lx, ly, lz;
    std::vector<BufferView*> buffers;
    std::vector<UniformTarget> uniforms;
};
struct BufferObject {
    void* data;
    size_t bytes;
};
struct State {
    std::vector<Program> programs;
    std::vector<BufferView*> pending_buffers;
    int pending_group[3] = {0, 0, 0};
    std::vector<std::string> uniform_names;                 // handle -> name
    std::map<std::string, std::vector<float>> uniform_values;   // name -> 4 floats
    std::vector<BufferObject> buffers;                      // handle -> storage
    int slots[32];
    State() {
        for (int& s : slots) s = -1;
    }
};
static State& state() {
    static State s;
    return s;
}

void register_buffer(BufferView* b) { state().pending_buffers.push_back(b); }
GroupSize::GroupSize(int x, int y, int z) {
    State& s = state();
    s.pending_group[0] = x;
    s.pending_group[1] = y;
    s.pending_group[2] = z;
}
void begin_program(const char* name, RowsFn rows) {
    State& s = state();
    Program p;
    p.name = name;
    p.rows = rows;
    p.lx = s.pending_group[0];
    p.ly = s.pending_group[1];
    p.lz = s.pending_group[2];
    p.buffers.swap(s.pending_buffers);
    s.pending_group[0] = s.pending_group[1] = s.pending_group[2] = 0;
    s.programs.push_back(p);
}
void register_uniform(const char* name, float* dst, int components) {
    state().programs.back().uniforms.push_back({name, dst, components});
}
}  // namespace natrix_ref

using natrix_ref::state;

#ifndef NATRIX_REF_BUILD_INFO
#define NATRIX_REF_BUILD_INFO "unknown"
#endif

extern "C" {

const char* nref_build_info(void) { return NATRIX_REF_BUILD_INFO; }

int nref_program_count(void) { return (int)state().programs.size(); }
const char* nref_program_name(int i) { return state().programs[i].name.c_str(); }

int nref_create_uniform(const char* name) {
    auto& s = state();
    for (size_t i = 0; i < s.uniform_names.size(); ++i)
        if (s.uniform_names[i] == name) return (int)i;          // bgfx: same name, same uniform
    s.uniform_names.push_back(name);
    s.uniform_values[name] = std::vector<float>(4, 0.0f);
    return (int)s.uniform_names.size() - 1;
}

int nref_set_uniform(int handle, const float* values, int n) {
    auto& s = state();
    if (handle < 0 || handle >= (int)s.uniform_names.size() || n < 0 || n > 4) return -1;
    std::vector<float>& v = s.uniform_values[s.uniform_names[handle]];
    for (int i = 0; i < n; ++i) v[i] = values[i];
    return 0;
}

int nref_create_buffer(size_t bytes) {
    auto& s = state();
    void* p = calloc(bytes ? bytes : 1, 1);
    if (!p) return -1;
    for (size_t i = 0; i < s.buffers.size(); ++i)
        if (!s.buffers[i].data) {
            s.buffers[i] = {p, bytes};
            return (int)i;
        }
    s.buffers.push_back({p, bytes});
    return (int)s.buffers.size() - 1;
}
void* nref_buffer_ptr(int handle) {
    auto& s = state();
    return (handle >= 0 && handle < (int)s.buffers.size()) ? s.buffers[handle].data : nullptr;
}
size_t nref_buffer_bytes(int handle) {
    auto& s = state();
    return (handle >= 0 && handle < (int)s.buffers.size()) ? s.buffers[handle].bytes : 0;
}
void nref_destroy_buffer(int handle) {
    auto& s = state();
    if (handle < 0 || handle >= (int)s.buffers.size()) return;
    free(s.buffers[handle].data);
    s.buffers[handle] = {nullptr, 0};
    for (int& slot : s.slots)
        if (slot == handle) slot = -1;
}
int nref_set_buffer(int slot, int handle) {
    auto& s = state();
    if (slot < 0 || slot >= 32) return -1;
    s.slots[slot] = handle;
    return 0;
}

int nref_create_program(const char* shader_file) {
    auto& s = state();
    for (size_t i = 0; i < s.programs.size(); ++i)
        if (s.programs[i].name == shader_file) return (int)i;
    return -1;
}

int nref_dispatch(int program, int gx, int gy, int gz) {
    auto& s = state();
    if (program < 0 || program >= (int)s.programs.size()) return -1;
    natrix_ref::Program& p = s.programs[program];
    for (natrix_ref::BufferView* b : p.buffers) {
        int h = (b->slot >= 0 && b->slot < 32) ? s.slots[b->slot] : -1;
        b->base = (h >= 0) ? s.buffers[h].data : nullptr;
    }
    for (const natrix_ref::UniformTarget& u : p.uniforms) {
        auto it = s.uniform_values.find(u.name);
        for (int c = 0; c < u.components; ++c) u.dst[c] = (it == s.uniform_values.end()) ? 0.0f : it->second[c];
    }
    const long nx = (long)gx * p.lx, ny = (long)gy * p.ly, nz = (long)gz * p.lz;
    natrix_ref::RowsFn rows = p.rows;
    const long chunk = 4;                                  // rows per task
    for (long z = 0; z < nz; ++z) {
#pragma omp parallel for schedule(static)
        for (long y0 = 0; y0 < ny; y0 += chunk) rows(y0, (y0 + chunk < ny) ? y0 + chunk : ny, nx, (unsigned)z);
    }
    return 0;
}

void nref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int nref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}  // extern "C"
