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
// Box-Muller pair (stands in for standard_normal): radius from the high 32 bits, angle from the low 32
// bits.  Evaluated in SINGLE precision with explicitly rounded operations only (fma, mul, sub, IEEE sqrt,
// integer ops; no MUFU approximations, no contraction), so the CPU oracle reproduces every bit, at a
// quarter of the instructions of the double-precision log / sqrt / sincospi -- the normals were a third of
// all instructions of an MH step.  u1 = (2 hi + 1) 2^-33 is split exactly into m 2^e with a 24-bit m, so the
// radius is the exact radius of a u1 rounded to 24 significant bits (|z| reaches 6.76); the angle keeps 24
// of its 30 bits after the exact quadrant reduction.  Polynomials: Chebyshev fits, errors below 1e-8.
__device__ __forceinline__ void word_to_normals(uint64_t w, double &z0, double &z1)
{
    const uint32_t hi = (uint32_t)(w >> 32), lo = (uint32_t)w;
    const uint64_t n = ((uint64_t)hi << 1) | 1ull;
    const int lz = __clzll((long long)n);
    const uint32_t mant = (uint32_t)((n << lz) >> 40);  // [2^23, 2^24)
    int e = 30 - lz;
    float m = __fmul_rn(__uint2float_rn(mant), 1.1920929e-07f);  // mant 2^-23 in [1, 2), exact
    if (mant > 11863283u) {  // m > sqrt(2)
        m = __fmul_rn(m, 0.5f);
        e += 1;
    }
    const float t = __fsub_rn(m, 1.0f);  // exact
    float p = 0.0874394551f;             // ln(1 + t) / t on [sqrt(1/2) - 1, sqrt(2) - 1]
    p = __fmaf_rn(p, t, -0.143773302f);
    p = __fmaf_rn(p, t, 0.149490952f);
    p = __fmaf_rn(p, t, -0.165606961f);
    p = __fmaf_rn(p, t, 0.199569777f);
    p = __fmaf_rn(p, t, -0.250021547f);
    p = __fmaf_rn(p, t, 0.333341837f);
    p = __fmaf_rn(p, t, -0.499999881f);
    p = __fmaf_rn(p, t, 1.0f);
    const float lnu = __fmaf_rn(__int2float_rn(e), 0.693147182f, __fmul_rn(p, t));
    const float r = __fsqrt_rn(__fmul_rn(-2.0f, lnu));
    const uint32_t quad = lo >> 30, g = (lo >> 6) & 0xFFFFFFu;
    const float f = __fmul_rn(__uint2float_rn(g), 2.98023224e-08f);  // g 2^-25 in [0, 1/2), exact
    const bool sw = g > 0x800000u;                                    // f > 1/4: use the co-function
    const float x = sw ? __fsub_rn(0.5f, f) : f;
    const float y = __fmul_rn(x, x);
    float sp = -0.589076877f;  // sin(pi x) / x in y = x^2, |x| <= 1/4
    sp = __fmaf_rn(sp, y, 2.54976702f);
    sp = __fmaf_rn(sp, y, -5.16770792f);
    sp = __fmaf_rn(sp, y, 3.14159274f);
    const float sx = __fmul_rn(sp, x);
    float cx = 0.231329247f;   // cos(pi x) in y
    cx = __fmaf_rn(cx, y, -1.3350445f);
    cx = __fmaf_rn(cx, y, 4.05870724f);
    cx = __fmaf_rn(cx, y, -4.93480206f);
    cx = __fmaf_rn(cx, y, 1.0f);
    const float s0 = sw ? cx : sx, c0 = sw ? sx : cx;  // sin(pi f), cos(pi f)
    const float s = (quad == 0) ? s0 : (quad == 1) ? c0 : (quad == 2) ? -s0 : -c0;
    const float c = (quad == 0) ? c0 : (quad == 1) ? -s0 : (quad == 2) ? -c0 : s0;
    z0 = (double)__fmul_rn(r, c);
    z1 = (double)__fmul_rn(r, s);
}

}  // namespace ptm
