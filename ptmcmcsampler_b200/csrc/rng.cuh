// Counter-based random draws for the PT-MCMC engine (sm_100a).
//
// Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit seed with counter
// (iter, purpose<<24 | block, walker, temperature).  A chain consumes 64-bit words strictly in
// order inside one iteration, in the reference's own draw order: proposal index
// (ref PTMCMCSampler.py:1058), the proposal's draws (:839-873, :897-930, :955-976), then the
// accept uniform (:616).  No generator state lives in HBM.
#pragma once
#ifdef __CUDACC_RTC__  // NVRTC ships no standard headers
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
#else
#include <cstdint>
#endif

#include "params.h"

namespace ptm {

enum : uint32_t { PURPOSE_MH = 0, PURPOSE_SWAP = 1 };

// The ten round keys (k0 + r W0, k1 + r W1) are the same for every draw of a run: the host expands
// them once into DevParams::rk (philox_round_keys in params.h), so that a round is two wide multiplies
// and two three-input XORs whose key operand comes straight from the constant bank.
__device__ __forceinline__ uint4 philox4x32_10(const DevParams &p, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ p.rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ p.rk[2 * r + 1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct Stream {
    const DevParams &p;
    uint32_t c0, c1b, c2, c3, j;
    uint4 blk;
    __device__ __forceinline__ Stream(const DevParams &p_, uint32_t purpose, uint64_t iter, uint32_t walker, uint32_t temp)
        : p(p_), c0((uint32_t)iter), c1b(purpose << 24), c2(walker), c3(temp), j(0)
    {
    }
    __device__ __forceinline__ uint64_t next()
    {
        if (!(j & 1u)) blk = philox4x32_10(p, c0, c1b | (j >> 1), c2, c3);
        const uint64_t w = (j & 1u) ? ((uint64_t)blk.z | ((uint64_t)blk.w << 32))
                                    : ((uint64_t)blk.x | ((uint64_t)blk.y << 32));
        ++j;
        return w;
    }
    // skip to word index jj
    __device__ __forceinline__ void seek(uint32_t jj)
    {
        j = jj;
        if (j & 1u) blk = philox4x32_10(p, c0, c1b | (j >> 1), c2, c3);
    }
    // Philox block b of this (iteration, walker, temperature): words 2b and 2b + 1
    __device__ __forceinline__ uint4 block(uint32_t b) const { return philox4x32_10(p, c0, c1b | b, c2, c3); }
};

__device__ __forceinline__ uint64_t lo_word(const uint4 &b) { return (uint64_t)b.x | ((uint64_t)b.y << 32); }
__device__ __forceinline__ uint64_t hi_word(const uint4 &b) { return (uint64_t)b.z | ((uint64_t)b.w << 32); }
// word j of a stream through ONE out-of-line copy of the generator (rare paths of the unrolled kernels)
static __device__ __noinline__ uint64_t stream_word(const Stream &st, uint32_t j)
{
    const uint4 b = st.block(j >> 1);
    return (j & 1u) ? hi_word(b) : lo_word(b);
}

// integer in [0, n): high 64 bits of word*n (stands in for Generator.integers)
__device__ __forceinline__ uint64_t word_to_int(uint64_t w, uint64_t n) { return __umul64hi(w, n); }
// the same value for n < 2^32 with two 32 x 32 -> 64 multiplies
__device__ __forceinline__ uint32_t word_to_int32(uint64_t w, uint32_t n)
{
    const uint64_t lo = (uint64_t)(uint32_t)w * n, hi = (uint64_t)(uint32_t)(w >> 32) * n;
    return (uint32_t)((hi + (lo >> 32)) >> 32);
}
// double in [0,1) with 53 random bits (stands in for Generator.random / uniform)
__device__ __forceinline__ double word_to_unit(uint64_t w)
{
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

// Hastings test  diff > log(u)  (ref :615-616) with u = word_to_unit(w).  A single-precision log2 decides
// unless diff lies within its error bound of log(u) (about one draw in 10^5), where the double-precision
// logarithm is evaluated: the decision is exactly that of the double-precision comparison, at a fifth of
// its instructions.  Error budget of the filter: MUFU.LG2 2^-22 absolute on [1/2, 2], 2 ulp elsewhere,
// the float conversion of u and the product with ln 2 2^-24 relative each; the margin is > 6x that.
static __device__ __noinline__ bool hastings_exact(double diff, double u) { return diff > log(u); }

__device__ __forceinline__ bool hastings_accept(double diff, uint64_t w)
{
    const double u = word_to_unit(w);
    const double la = (double)(__log2f((float)u) * 0.693147182f);
    const double margin = fma(fabs(la), 2e-6, 2e-6);
    if (diff > la + margin) return true;
    if (diff < la - margin) return false;
    return hastings_exact(diff, u);  // also u = 0 (la = -inf) and NaN differences (compare false, ref :616)
}

// Box-Muller pair (stands in for standard_normal): radius from the high 32 bits, angle from the low 32
// bits.  Evaluated in SINGLE precision with explicitly rounded operations only (fma, mul, sub, IEEE sqrt,
// integer ops; no MUFU approximations, no contraction), so the CPU oracle reproduces every bit, at a
// quarter of the instructions of the double-precision log / sqrt / sincospi.
//   u1 = (2 hi + 1) 2^-33 truncated to 24 significant bits: the round-toward-zero conversion and fma below
//   produce exactly the oracle's  mant 2^e  (its 64-bit normalisation, orc_word_to_normals) in two
//   instructions; |z| reaches 6.76.  The split  u1 = m 2^e, m in [sqrt(1/2), sqrt(2))  is the usual
//   integer add on the float's bits.  The angle keeps 24 of its 30 bits after the exact quadrant reduction;
//   quadrant and co-function selection are a select and a sign-bit XOR.
// Polynomials: Chebyshev fits, errors below 1e-8.
__device__ __forceinline__ void word_to_normals(uint32_t hi, uint32_t lo, double &z0, double &z1)
{
    const float u1 = __fmaf_rz(__uint2float_rz(hi), __uint_as_float(0x2F800000u), __uint_as_float(0x2F000000u));
    const uint32_t ix = __float_as_uint(u1) + 0x004AFB0Cu;  // carries into the exponent when m > sqrt(2)
    const int e = (int)(ix >> 23) - 127;
    const float m = __uint_as_float((ix & 0x007FFFFFu) + 0x3F3504F4u);
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
    // (sin, cos)(pi (f + quad / 2)): odd quadrants swap the pair once more, signs follow the quadrant
    const bool flip = sw != ((quad & 1u) != 0u);
    const uint32_t sb = __float_as_uint(flip ? cx : sx) ^ ((quad & 2u) << 30);
    const uint32_t cb = __float_as_uint(flip ? sx : cx) ^ (((quad + 1u) & 2u) << 30);
    z0 = (double)__fmul_rn(r, __uint_as_float(cb));
    z1 = (double)__fmul_rn(r, __uint_as_float(sb));
}
__device__ __forceinline__ void word_to_normals(uint64_t w, double &z0, double &z1)
{
    word_to_normals((uint32_t)(w >> 32), (uint32_t)w, z0, z1);
}

}  // namespace ptm
