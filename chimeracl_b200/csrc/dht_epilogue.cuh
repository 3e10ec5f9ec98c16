// chimera-b200: epilogue shared by the FP64 (dht.cu) and int8-sliced (dht_i8.cu)
// contraction kernels: C (op)= alpha * (v0, v1) for one pair of adjacent columns.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace chb {

__device__ __forceinline__ void gemm_store(double* c, double v0, double v1, double are,
                                           double aim, bool cplx, bool accumulate, bool pair) {
  if (!(are == 1.0 && aim == 0.0)) {
    if (cplx) {
      const double re = are * v0 - aim * v1;
      const double im = are * v1 + aim * v0;
      v0 = re; v1 = im;
    } else {
      v0 *= are; v1 *= are;
    }
  }
  if (pair && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
    double2 o = make_double2(v0, v1);
    if (accumulate) { double2 old = *reinterpret_cast<double2*>(c); o.x += old.x; o.y += old.y; }
    *reinterpret_cast<double2*>(c) = o;
  } else {
    if (accumulate) { v0 += c[0]; if (pair) v1 += c[1]; }
    c[0] = v0;
    if (pair) c[1] = v1;
  }
}

}  // namespace chb
