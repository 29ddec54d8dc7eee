/*
 * ORACLE -- CPU restatement of the lmbspecialops geometry ops on the DeMoN
 * inference hot path.  TEST INFRASTRUCTURE ONLY: nothing in demon_b200/ may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * The reference itself (TensorFlow 1.4 custom ops on Eigen) cannot be built in
 * this environment (no TensorFlow, no Eigen), see DESIGN.md "Oracle".  The op
 * bodies are restated in geometry_ops_impl.h for float and double; pinning
 * status per op is listed in DESIGN.md (median3x3: reference KATs, exact;
 * depth_to_flow/flow_to_depth: the reference's round-trip test; warp2d,
 * scale_invariant_gradient and leaky_relu forward values: PARITY UNPINNED by
 * the reference's tests -- pinned only by this restatement).
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off -shared -fPIC)
 */
#include <limits.h>
#include <math.h>
#include <float.h>

#define CAT_(a,b) a##b

/* ---- float ---- */
#define T float
#define FN(x) CAT_(x,_f32)
static float sqrt__f32(float x) { return sqrtf(x); }
static float sin__f32(float x) { return sinf(x); }
static float cos__f32(float x) { return cosf(x); }
static float abs__f32(float x) { return fabsf(x); }
static int isfinite__f32(float x) { return isfinite(x); }
static float nan__f32(void) { return NAN; }
static float eps__f32(void) { return FLT_EPSILON; }
#include "geometry_ops_impl.h"
#undef T
#undef FN

/* ---- double ---- */
#define T double
#define FN(x) CAT_(x,_f64)
static double sqrt__f64(double x) { return sqrt(x); }
static double sin__f64(double x) { return sin(x); }
static double cos__f64(double x) { return cos(x); }
static double abs__f64(double x) { return fabs(x); }
static int isfinite__f64(double x) { return isfinite(x); }
static double nan__f64(void) { return (double)NAN; }
static double eps__f64(void) { return DBL_EPSILON; }
#include "geometry_ops_impl.h"
#undef T
#undef FN
