// Device-side LLG building blocks: virtual force, counter-based thermal field, and the
// per-spin update rules of the solvers. Everything is evaluated in registers inside the fused
// stage kernels (device/kernels.cu) -- the reference runs one full-field sweep per primitive
// (~75 sweeps per Depondt iteration).
#pragma once

#include "stencil.cuh"

namespace sb
{
namespace dev
{

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11). Counter = (site_lo, site_hi, iter_lo, iter_hi), key = seed.
// The thermal field of the reference is a serial std::mt19937 + std::normal_distribution stream
// (Method_LLG.cpp:65-110); it cannot and need not be reproduced bit-wise -- parity at T>0 is
// statistical. What is kept: one xi per (iteration, site, component), shared by predictor and
// corrector, scaled by epsilon*sqrt(T/mu_s). Keyed by the GLOBAL site index, the noise is
// independent of the multi-GPU decomposition.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10( unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4] )
{
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for( int r = 0; r < 10; ++r )
    {
        const unsigned hi0 = __umulhi( M0, c0 ), lo0 = M0 * c0;
        const unsigned hi1 = __umulhi( M1, c2 ), lo1 = M1 * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Three standard normal variates for (seed, iteration, global site): Box-Muller in fp64 on
// uniforms of 32-bit resolution, u in (0,1).
__device__ __forceinline__ D3 gaussian3( std::uint64_t seed, std::uint64_t iteration, std::uint64_t site )
{
    unsigned r[4];
    philox4x32_10(
        unsigned( site ), unsigned( site >> 32 ), unsigned( iteration ), unsigned( iteration >> 32 ), unsigned( seed ),
        unsigned( seed >> 32 ), r );
    const double scale = 2.3283064365386963e-10; // 2^-32
    const double u0 = ( double( r[0] ) + 0.5 ) * scale, u1 = ( double( r[1] ) + 0.5 ) * scale;
    const double u2 = ( double( r[2] ) + 0.5 ) * scale, u3 = ( double( r[3] ) + 0.5 ) * scale;
    const double rad0 = sqrt( -2.0 * log( u0 ) ), rad1 = sqrt( -2.0 * log( u2 ) );
    double s0, c0, s1;
    sincospi( 2.0 * u1, &s0, &c0 );
    s1 = sinpi( 2.0 * u3 );
    return make_d3( rad0 * c0, rad0 * s0, rad1 * s1 );
}

template<int NB_T>
__device__ __forceinline__ D3 thermal_field( const StencilParams & p, const LLGParams & l, const Site & site )
{
    // global site index in the reference's order
    const std::uint64_t gsite
        = std::uint64_t( site.a ) * p.NB + site.ib
          + std::uint64_t( p.Na ) * p.NB * ( std::uint64_t( site.b ) + std::uint64_t( p.Nb ) * ( p.c_begin + site.c ) );
    const D3 n      = gaussian3( l.seed, l.iteration, gsite );
    const double sc = l.thermal_scale[NB_T == 1 ? 0 : site.ib];
    return make_d3( sc * n.x, sc * n.y, sc * n.z );
}

// Virtual force (Method_LLG.cpp:131-226): F = -gradient.
//   dynamics:      Fv = dtg/mu_s (F + alpha s x F) [+ STT monolayer] [+ xi + alpha s x xi]
//   minimisation:  Fv = dtg' s x F
template<int NB_T>
__device__ __forceinline__ D3
virtual_force( const LLGParams & l, const Site & site, const D3 & s, const D3 & F, const D3 & xi )
{
    if( l.direct_minimization )
    {
        const D3 c = cross3( s, F );
        return make_d3( l.dtg * c.x, l.dtg * c.y, l.dtg * c.z );
    }
    const D3 sxF     = cross3( s, F );
    const double da  = l.dtg * l.damping;
    const double ims = l.inv_mu_s[NB_T == 1 ? 0 : site.ib];
    D3 fv            = make_d3(
        ( l.dtg * F.x + da * sxF.x ) * ims, ( l.dtg * F.y + da * sxF.y ) * ims, ( l.dtg * F.z + da * sxF.z ) * ims );
    if( l.has_stt )
    {
        // monolayer approximation: Fv += c1 * pol + c2 * (pol x s)   (Method_LLG.cpp:207-212)
        const D3 pol = make_d3( l.stt_pol[0], l.stt_pol[1], l.stt_pol[2] );
        const D3 pxs = cross3( pol, s );
        fv.x += l.stt_c1 * pol.x + l.stt_c2 * pxs.x;
        fv.y += l.stt_c1 * pol.y + l.stt_c2 * pxs.y;
        fv.z += l.stt_c1 * pol.z + l.stt_c2 * pxs.z;
    }
    if( l.has_thermal )
    {
        const D3 sxxi = cross3( s, xi );
        fv.x += xi.x + l.damping * sxxi.x;
        fv.y += xi.y + l.damping * sxxi.y;
        fv.z += xi.z + l.damping * sxxi.z;
    }
    return fv;
}

// Rodrigues rotation of v about H by the angle |H| (Depondt; Vectormath.cpp:474-485 with
// axis = H/|H|, angle = |H|, Solver_Depondt.hpp:43-52). Written in terms of theta^2 = |H|^2:
//   R v = v cos(t) + (H x v) sin(t)/t + H (H.v) (1-cos(t))/t^2
// so that no normalisation of the axis (and no division by zero for H = 0) is needed.
// For t^2 < 0.25 the three even functions are evaluated as polynomials in t^2 (truncation < 1e-17);
// otherwise through sincos.
__device__ __forceinline__ D3 rotate_about( const D3 & v, const D3 & H )
{
    const double t2 = dot3( H, H );
    double c, sinc, omc; // cos t, sin t / t, (1 - cos t)/t^2
    if( t2 < 0.25 )
    {
        // Horner in x = t^2: sinc = sum (-x)^k/(2k+1)!, omc = sum (-x)^k/(2k+2)!
        const double x = t2;
        sinc = 1.0
               - x / 6.0
                     * ( 1.0
                         - x / 20.0
                               * ( 1.0
                                   - x / 42.0
                                         * ( 1.0
                                             - x / 72.0
                                                   * ( 1.0
                                                       - x / 110.0
                                                             * ( 1.0 - x / 156.0 * ( 1.0 - x / 210.0 * ( 1.0 - x / 272.0 ) ) ) ) ) ) );
        omc = 0.5
              * ( 1.0
                  - x / 12.0
                        * ( 1.0
                            - x / 30.0
                                  * ( 1.0
                                      - x / 56.0
                                            * ( 1.0
                                                - x / 90.0
                                                      * ( 1.0
                                                          - x / 132.0
                                                                * ( 1.0 - x / 182.0 * ( 1.0 - x / 240.0 * ( 1.0 - x / 306.0 ) ) ) ) ) ) ) );
        c = 1.0 - x * omc;
    }
    else
    {
        const double t = sqrt( t2 );
        double sn;
        sincos( t, &sn, &c );
        sinc = sn / t;
        omc  = ( 1.0 - c ) / t2;
    }
    const D3 Hxv    = cross3( H, v );
    const double hv = dot3( H, v ) * omc;
    return make_d3( v.x * c + Hxv.x * sinc + H.x * hv, v.y * c + Hxv.y * sinc + H.y * hv, v.z * c + Hxv.z * sinc + H.z * hv );
}

// Eigen's normalize(): leaves the zero vector untouched
__device__ __forceinline__ D3 normalized3( const D3 & v )
{
    const double n2 = dot3( v, v );
    if( n2 > 0 )
    {
        const double inv = 1.0 / sqrt( n2 );
        return make_d3( v.x * inv, v.y * inv, v.z * inv );
    }
    return v;
}

// Cayley-type transform of the semi-implicit B solver (Solver_Kernels.cpp:15-48)
__device__ __forceinline__ D3 sib_transform( const D3 & s, const D3 & force )
{
    const D3 A         = make_d3( 0.5 * force.x, 0.5 * force.y, 0.5 * force.z );
    const double detAi = 1.0 / ( 1.0 + dot3( A, A ) );
    const D3 sxA       = cross3( s, A );
    const D3 a2        = make_d3( s.x - sxA.x, s.y - sxA.y, s.z - sxA.z );
    D3 o;
    o.x = ( a2.x * ( A.x * A.x + 1 ) + a2.y * ( A.x * A.y - A.z ) + a2.z * ( A.x * A.z + A.y ) ) * detAi;
    o.y = ( a2.x * ( A.y * A.x + A.z ) + a2.y * ( A.y * A.y + 1 ) + a2.z * ( A.y * A.z - A.x ) ) * detAi;
    o.z = ( a2.x * ( A.z * A.x - A.y ) + a2.y * ( A.z * A.y + A.x ) + a2.z * ( A.z * A.z + 1 ) ) * detAi;
    return o;
}

} // namespace dev
} // namespace sb
