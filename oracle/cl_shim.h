/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Shim that lets g++ compile the reference's OpenCL C kernel sources
 * (/root/reference/chimeraCL/kernels/*.cl) UNMODIFIED as host C++.
 * The .cl files are #included from where they lie; nothing is copied.
 * Build recipe: oracle/Makefile -> oracle/_ref/*.so (git-ignored).
 *
 * Semantics chosen so results are reproducible:
 *   - IEEE double, no FMA contraction (-ffp-contract=off), no fast-math
 *   - a work-item == one call of the kernel body with get_global_id(0)
 *     returning a thread-local counter set by the NDRange drivers
 *   - atom_add -> __atomic_fetch_add (relaxed)
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

typedef unsigned int uint;
struct double2 { double s0, s1; };
static inline double2 operator+(double2 a, double2 b) {
  return double2{a.s0 + b.s0, a.s1 + b.s1};
}

#define __kernel extern "C"
#define __global
#define __constant const

static thread_local size_t chimera_ref_gid0;
static inline size_t get_global_id(int) { return chimera_ref_gid0; }
static inline uint atom_add(uint *p, uint v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

#ifndef BLOCK_SIZE
#define BLOCK_SIZE 32
#endif

using std::cos;
using std::floor;
using std::sin;
using std::sqrt;
