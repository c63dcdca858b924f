// fp64 tensor-core primitive shared by the MH and the covariance kernels.
#pragma once

namespace ptm {

// D(8x8) += A(8x4) . B(4x8), mma.sync m8n8k4 f64 (DMMA).  Lane l holds A[l>>2][l&3], B[l&3][l>>2] and
// C/D[l>>2][2(l&3) + {0,1}].  Measured 37.1 TFLOP/s on B200 (scripts/micro/dmma_rate.cu), the DFMA peak.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

}  // namespace ptm
