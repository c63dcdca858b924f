// Counter-based random draws for the PT-MCMC engine (sm_100a).
//
// Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit seed with counter
// (iter, purpose<<24 | block, walker, temperature).  A chain consumes 64-bit words strictly in
// order inside one iteration, in the reference's own draw order: proposal index
// (ref PTMCMCSampler.py:1058), the proposal's draws (:839-873, :897-930, :955-976), then the
// accept uniform (:616).  No generator state lives in HBM.
#pragma once
#include <cstdint>

namespace ptm {

enum : uint32_t { PURPOSE_MH = 0, PURPOSE_SWAP = 1 };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct Stream {
    uint32_t c0, c1b, c2, c3, k0, k1, j;
    uint4 blk;
    __device__ __forceinline__ Stream(uint64_t seed, uint32_t purpose, uint64_t iter, uint32_t walker,
                                      uint32_t temp)
        : c0((uint32_t)iter), c1b(purpose << 24), c2(walker), c3(temp), k0((uint32_t)seed),
          k1((uint32_t)(seed >> 32)), j(0)
    {
    }
    __device__ __forceinline__ uint64_t next()
    {
        if (!(j & 1u)) blk = philox4x32_10(c0, c1b | (j >> 1), c2, c3, k0, k1);
        const uint64_t w = (j & 1u) ? ((uint64_t)blk.z | ((uint64_t)blk.w << 32))
                                    : ((uint64_t)blk.x | ((uint64_t)blk.y << 32));
        ++j;
        return w;
    }
    // skip to word index jj (used by the accept kernel of the host-callback path)
    __device__ __forceinline__ void seek(uint32_t jj)
    {
        j = jj;
        if (j & 1u) blk = philox4x32_10(c0, c1b | (j >> 1), c2, c3, k0, k1);
    }
};

// integer in [0, n): high 64 bits of word*n (stands in for Generator.integers)
__device__ __forceinline__ uint64_t word_to_int(uint64_t w, uint64_t n) { return __umul64hi(w, n); }
// double in [0,1) with 53 random bits (stands in for Generator.random / uniform)
__device__ __forceinline__ double word_to_unit(uint64_t w)
{
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}
// Box-Muller pair: radius from the high 32 bits, angle from the low 32 bits with the quadrant
// taken from the integer so the reduced angle is exact (stands in for standard_normal)
__device__ __forceinline__ void word_to_normals(uint64_t w, double &z0, double &z1)
{
    const uint32_t hi = (uint32_t)(w >> 32), lo = (uint32_t)w;
    const double u1 = ((double)hi + 0.5) * (1.0 / 4294967296.0);
    const double r = sqrt(-2.0 * log(u1));
    const uint32_t quad = lo >> 30;
    const double f = (double)(lo & 0x3FFFFFFFu) * (1.0 / 2147483648.0);
    double s0, c0;
    sincospi(f, &s0, &c0);
    const double s = (quad == 0) ? s0 : (quad == 1) ? c0 : (quad == 2) ? -s0 : -c0;
    const double c = (quad == 0) ? c0 : (quad == 1) ? -s0 : (quad == 2) ? -c0 : s0;
    z0 = r * c;
    z1 = r * s;
}

}  // namespace ptm
