// runtime.h - what a shader translation unit needs from the shim runtime.  TEST INFRASTRUCTURE ONLY.
#ifndef NATRIX_REF_RUNTIME_H
#define NATRIX_REF_RUNTIME_H
namespace natrix_ref {
// Claims every BUFFER_* view and the NUM_THREADS size registered since the previous call for
// the program `name` (static initialisation runs a translation unit's objects in order).
typedef void (*RowsFn)(long y0, long y1, long nx, unsigned z);   // invocations of rows [y0, y1) of slice z
void begin_program(const char* name, RowsFn rows);
void register_uniform(const char* name, float* dst, int components);
}  // namespace natrix_ref
#endif
