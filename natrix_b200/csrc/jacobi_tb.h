// jacobi_tb.h - temporally blocked pressure-Jacobi solver (the hot loop).
#pragma once
#include "common.cuh"

namespace natrix {

constexpr int JACOBI_TB_MAX_DEPTH = 8;

struct JacobiTB;   // caches TMA descriptors and launch geometry

JacobiTB* jacobi_tb_create();
void jacobi_tb_destroy(JacobiTB* tb);
const char* jacobi_tb_error(JacobiTB* tb);
// The TMA path needs 16-byte row pitches (width % 16 == 0) and width >= 256; other grids use
// the shared-memory kernel (jacobi_smem.cu).
bool jacobi_tb_supported(const Geom& g);

// Runs `depth` (1..JACOBI_TB_MAX_DEPTH) Jacobi sweeps pin -> pout for local rows [r0, r1).
// div4 is the PRE-SCALED divergence (0.25 b, or b itself where the mask carries NB_RAW: common.cuh).
// pin / div4 / nbmask must be valid on rows [r0-depth, r1+depth) clipped to the global domain.
// p_is_zero: pin is known to be all zero (first block of a step), so it is not read.
// boxes: nboxes x (x0, x1, y0, y1) bounding boxes (global columns, local rows, half-open), or circles
// encoded as (cx, -1 - radius, cy, 0), of the obstacles stamped this step - a scheduling hint only: tiles are cut shorter where the select body
// will run; results never depend on it.
// Returns the number of kernels launched, or -1 on error (see jacobi_tb_error).
int jacobi_tb_launch(JacobiTB* tb, const float* pin, const float* div4, const uint8_t* nbmask,
                     float* pout, Geom g, int depth, int r0, int r1, bool p_is_zero, int packed,
                     const int* boxes, int nboxes, cudaStream_t st);

// The following launches leave n SMs without a block (0 = use every SM): room for a halo-exchange kernel that
// runs beside an interior launch.
void jacobi_tb_reserve_sms(JacobiTB* tb, int n);

// Tile-plan cache counters: launches that found their plan / that had to cut and upload a new one.
void jacobi_tb_plan_stats(JacobiTB* tb, unsigned long long* hits, unsigned long long* misses);

// Host-only: the tile plan jacobi_tb_launch would use for these arguments (no CUDA call), as
// (strip, first row, end row, 0) quadruples; a strip is SW - 2*hx output columns wide, hx = 4 for depth <= 4
// else 8.  Returns the number of tiles (which may exceed `cap`; only `cap` are written).  For tests.
int jacobi_tb_plan_debug(int w, int depth, int r0, int r1, const int* boxes, int nboxes, int max_tiles,
                         int* out4, int cap);

}  // namespace natrix
