// Device-side LLG building blocks: virtual force, counter-based thermal field, and the
// per-spin update rules of the solvers. Everything is evaluated in registers inside the fused
// stage kernels (device/kernels.cuh, device/sc6.cuh) -- the reference runs one full-field sweep per
// primitive (~75 sweeps per Depondt iteration).
//
// The step is close to the fp64-pipe / HBM balance point of B200, so the arithmetic here is
// written for a minimal number of fp64 instructions: no divisions, polynomial coefficients as literals.
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

// MUFU-based fp32 primitives (approx, flush-to-zero): one SFU instruction each instead of the ~100-instruction fp64
// library routines. They are used ONLY to shape the random variates of the thermal field.
__device__ __forceinline__ float sfu_lg2( float x )
{
    float y;
    asm( "lg2.approx.ftz.f32 %0, %1;" : "=f"( y ) : "f"( x ) );
    return y;
}
__device__ __forceinline__ float sfu_sqrt( float x )
{
    float y;
    asm( "sqrt.approx.ftz.f32 %0, %1;" : "=f"( y ) : "f"( x ) );
    return y;
}
__device__ __forceinline__ float sfu_sin( float x )
{
    float y;
    asm( "sin.approx.ftz.f32 %0, %1;" : "=f"( y ) : "f"( x ) );
    return y;
}
__device__ __forceinline__ float sfu_cos( float x )
{
    float y;
    asm( "cos.approx.ftz.f32 %0, %1;" : "=f"( y ) : "f"( x ) );
    return y;
}
// 23 random bits -> float in [1, 2): the angle of a Box-Muller pair in turns (sin / cos are periodic, no subtraction needed)
__device__ __forceinline__ float unit_1_2( unsigned r )
{
    return __uint_as_float( ( r >> 9 ) | 0x3f800000u );
}
// Uniform in (0, 1) for the radius of a Box-Muller pair.
//   default: 23 random bits, 1.mantissa - (1 - 2^-24): one integer and one fp32 instruction. The smallest value is 2^-24, so the
//     radius stops at sqrt(-2 ln 2^-24) = 5.77 sigma; a normal variate lies beyond that with probability 8e-9
//     (tests/test_thermal_variates_gpu.py: moments up to the sixth, distribution function and tails to 5 sigma on 1.2e8 samples).
//   SB_THERMAL_TAIL32 = 1: all 32 bits, (r + 1/2) 2^-32 through an integer -> float conversion (exact for small r, the far
//     tail): radius up to 6.76 sigma. The conversion runs on the quarter-rate pipe next to the 7 SFU instructions of the
//     shaping: measured + 1.3 % on the whole Depondt step (profiles/r2o), which is why it is not the default.
#ifndef SB_THERMAL_TAIL32
#define SB_THERMAL_TAIL32 0
#endif
__device__ __forceinline__ float unit_open( unsigned r )
{
#if SB_THERMAL_TAIL32
    return fmaf( __uint2float_rn( r ), 2.3283064365386963e-10f, 1.1641532182693481e-10f );
#else
    return __uint_as_float( ( r >> 9 ) | 0x3f800000u ) - 0.99999994f;
#endif
}

#ifndef SB_THERMAL_FP64
#define SB_THERMAL_FP64 0 // 1: Box-Muller shaped in fp64 (log, sqrt, sincospi): the variant bench.py times beside the product
#endif

// Three normal variates of standard deviation sigma from four Philox words: Box-Muller shaped in fp32 with SFU
// instructions (|error| of sin / cos / lg2 ~ 1e-6 absolute): they are random numbers whose distribution, not whose
// digits, matters (T > 0 parity is statistical; tests/test_thermal_variates_gpu.py checks the first four
// moments and the distribution function on 1e8 samples), and an fp64 Box-Muller (log, sqrt, sincospi in software)
// costs more instructions than the whole rest of a solver stage (measured: bench.py `checks.thermal_fp64`). Everything
// downstream of the variates is fp64.
//   radius  sigma sqrt(-2 ln u) = sqrt(k lg2 u),  k = -2 ln2 sigma^2 (host constant, LLGParams::thermal_k)
//   angle   2 pi v with v = 1.mantissa in [1, 2): the same point of the circle as v - 1
__device__ __forceinline__ float3 scaled_gaussian3f( unsigned r0, unsigned r1, unsigned r2, unsigned r3, float k )
{
#if SB_THERMAL_FP64
    const double u0 = ( double( r0 ) + 0.5 ) * 2.3283064365386963e-10, u2 = ( double( r2 ) + 0.5 ) * 2.3283064365386963e-10;
    const double kd = double( k ) * 1.4426950408889634; // k lg2 u = (k / ln 2) ln u
    const double rad0 = sqrt( kd * log( u0 ) ), rad1 = sqrt( kd * log( u2 ) );
    double s0, c0, s1, c1;
    sincospi( 2.0 * ( double( r1 ) + 0.5 ) * 2.3283064365386963e-10, &s0, &c0 );
    sincospi( 2.0 * ( double( r3 ) + 0.5 ) * 2.3283064365386963e-10, &s1, &c1 );
    return make_float3( float( rad0 * c0 ), float( rad0 * s0 ), float( rad1 * s1 ) );
#else
    const float rad0 = sfu_sqrt( k * sfu_lg2( unit_open( r0 ) ) );
    const float rad1 = sfu_sqrt( k * sfu_lg2( unit_open( r2 ) ) );
    const float ang0 = 6.2831853071795865f * unit_1_2( r1 );
    const float ang1 = 6.2831853071795865f * unit_1_2( r3 );
    return make_float3( rad0 * sfu_cos( ang0 ), rad0 * sfu_sin( ang0 ), rad1 * sfu_sin( ang1 ) );
#endif
}
__device__ __forceinline__ D3 scaled_gaussian3( unsigned r0, unsigned r1, unsigned r2, unsigned r3, float k )
{
    const float3 v = scaled_gaussian3f( r0, r1, r2, r3, k );
    return make_d3( double( v.x ), double( v.y ), double( v.z ) );
}

// xi of the site `plane_site` (index inside its plane, reference order: ib + NB (a + Na b)) of the GLOBAL plane `gplane`
// for basis atom ib. Philox counter = (plane_site, gplane, iteration lo, iteration hi): independent of the
// decomposition over GPUs and of the kernel that evaluates it.
__device__ __forceinline__ D3 thermal_field_at( const LLGParams & l, unsigned plane_site, unsigned gplane, int ib )
{
    unsigned r[4];
    philox4x32_10(
        plane_site, gplane, unsigned( l.iteration ), unsigned( l.iteration >> 32 ), unsigned( l.seed ), unsigned( l.seed >> 32 ), r );
    return scaled_gaussian3( r[0], r[1], r[2], r[3], l.thermal_k[ib] );
}

template<int NB_T>
__device__ __forceinline__ D3 thermal_field( const StencilParams & p, const LLGParams & l, const Site & site )
{
    const unsigned plane_site = unsigned( site.a * p.NB + site.ib + p.Na * p.NB * site.b );
    const int ib              = NB_T == 1 ? 0 : site.ib;
    if( l.has_tgrad )
    {
        // site temperature of the linear gradient, cut off at zero (get_gradient_distribution, Vectormath.cpp:633-652)
        double T = l.tgrad_T0 + l.tgrad_cell[0] * site.a + l.tgrad_cell[1] * site.b + l.tgrad_cell[2] * ( p.c_begin + site.c )
                   + l.tgrad_basis[ib];
        T = fmin( fmax( T, 0.0 ), 1e30 );
        unsigned r[4];
        philox4x32_10(
            plane_site, unsigned( p.c_begin + site.c ), unsigned( l.iteration ), unsigned( l.iteration >> 32 ), unsigned( l.seed ),
            unsigned( l.seed >> 32 ), r );
        return scaled_gaussian3( r[0], r[1], r[2], r[3], l.thermal_k_per_T[ib] * float( T ) );
    }
    return thermal_field_at( l, plane_site, unsigned( p.c_begin + site.c ), ib );
}

// Virtual force (Method_LLG.cpp:131-226): F = -gradient.
//   dynamics:      Fv = dtg/mu_s (F + alpha s x F) [+ STT monolayer] [+ xi + alpha s x xi]
//   minimisation:  Fv = dtg' s x F
// evaluated as Fv = c1 F + xi + s x (c2 F + alpha xi), c1 = dtg/mu_s, c2 = alpha c1: one cross product instead of two.
__device__ __forceinline__ D3 virtual_force_ib( const LLGParams & l, int ib, const D3 & s, const D3 & F, const D3 & xi )
{
    if( l.direct_minimization )
    {
        const D3 c = cross3( s, F );
        return make_d3( l.dtg * c.x, l.dtg * c.y, l.dtg * c.z );
    }
    const double c1 = l.c1[ib], c2 = l.c2[ib];
    D3 w            = make_d3( c2 * F.x, c2 * F.y, c2 * F.z );
    D3 fv           = make_d3( c1 * F.x, c1 * F.y, c1 * F.z );
    if( l.has_thermal )
    {
        w  = make_d3( fma( l.damping, xi.x, w.x ), fma( l.damping, xi.y, w.y ), fma( l.damping, xi.z, w.z ) );
        fv = make_d3( fv.x + xi.x, fv.y + xi.y, fv.z + xi.z );
    }
    const D3 sxw = cross3( s, w );
    fv           = make_d3( fv.x + sxw.x, fv.y + sxw.y, fv.z + sxw.z );
    if( l.has_stt == 1 )
    {
        // monolayer approximation: Fv += c1 * pol + c2 * (pol x s)   (Method_LLG.cpp:207-212)
        const D3 pol = make_d3( l.stt_pol[0], l.stt_pol[1], l.stt_pol[2] );
        const D3 pxs = cross3( pol, s );
        fv.x += l.stt_c1 * pol.x + l.stt_c2 * pxs.x;
        fv.y += l.stt_c1 * pol.y + l.stt_c2 * pxs.y;
        fv.z += l.stt_c1 * pol.z + l.stt_c2 * pxs.z;
    }
    return fv;
}

template<int NB_T>
__device__ __forceinline__ D3
virtual_force( const LLGParams & l, const Site & site, const D3 & s, const D3 & F, const D3 & xi )
{
    return virtual_force_ib( l, NB_T == 1 ? 0 : site.ib, s, F, xi );
}

// Spin-transfer torque in the gradient approximation (Method_LLG.cpp:184-205 with Vectormath::jacobian, Vectormath.cpp:816-903):
// finite differences of the configuration `conf` along the three lattice translations of the SAME basis atom, central where
// both neighbours exist, one-sided (and doubled) at an open boundary, contracted with the current direction.
template<int NB_T>
__device__ __forceinline__ D3
stt_gradient_term( const StencilParams & p, const LLGParams & l, const ConstField3 & conf, const Site & site, const D3 & s )
{
    const int NB = NB_T > 0 ? NB_T : p.NB;
    const int x  = site.a * NB + ( NB_T == 1 ? 0 : site.ib );
    D3 grad      = make_d3( 0, 0, 0 );
#pragma unroll
    for( int t = 0; t < 3; ++t )
    {
        const int n   = t == 0 ? p.Na : ( t == 1 ? p.Nb : p.Nc );
        const int pos = t == 0 ? site.a : ( t == 1 ? site.b : site.c );
        D3 m0 = s, m1 = s;
        double factor = 0.5;
        for( int side = 0; side < 2; ++side )
        {
            int q      = pos + ( side == 0 ? 1 : -1 );
            bool valid = true;
            if( q < 0 || q >= n )
            {
                valid = p.bc[t] != 0;
                q     = q < 0 ? q + n : q - n;
            }
            if( valid )
            {
                const int j = t == 0 ? storage_index( p, q * NB + ( x - site.a * NB ), site.b, site.c )
                                     : ( t == 1 ? storage_index( p, x, q, site.c ) : storage_index( p, x, site.b, q ) );
                ( side == 0 ? m0 : m1 ) = load3( conf, j );
            }
            else
                factor *= 2;
        }
        const double w = l.stt_w[t] * factor;
        grad.x += w * ( m0.x - m1.x );
        grad.y += w * ( m0.y - m1.y );
        grad.z += w * ( m0.z - m1.z );
    }
    const D3 gxs = cross3( grad, s );
    return make_d3( l.stt_g1 * grad.x + l.stt_g2 * gxs.x, l.stt_g1 * grad.y + l.stt_g2 * gxs.y, l.stt_g1 * grad.z + l.stt_g2 * gxs.z );
}

// Taylor coefficients of sin(t)/t and (1-cos t)/t^2 in x = t^2: (-1)^k/(2k+1)! and (-1)^k/(2k+2)!. In constant
// memory so that the DFMAs take them as constant-bank operands (64-bit literals would each cost two UMOVs per use).
static __constant__ double ROT_SINC[7] = { 1.0, -1.6666666666666666e-01, 8.3333333333333332e-03, -1.9841269841269841e-04,
                                           2.7557319223985893e-06, -2.5052108385441720e-08, 1.6059043836821613e-10 };
static __constant__ double ROT_OMC[7]  = { 0.5, -4.1666666666666664e-02, 1.3888888888888889e-03, -2.4801587301587302e-05,
                                          2.7557319223985888e-07, -2.0876756987868100e-09, 1.1470745597729725e-11 };

// Rodrigues rotation of v about H by the angle |H| (Depondt; Vectormath.cpp:474-485 with
// axis = H/|H|, angle = |H|, Solver_Depondt.hpp:43-52). Written in terms of x = |H|^2:
//   R v = v cos(t) + (H x v) sin(t)/t + H (H.v) (1-cos(t))/t^2
// so that no normalisation of the axis (and no division by zero for H = 0) is needed.
// For x < 1/16 the even functions sin(t)/t and (1-cos t)/t^2 are degree-6 polynomials in x (truncation
// < 1.1e-17 relative); otherwise through sincos. dt-sized steps always take the polynomial branch.
__device__ __forceinline__ D3 rotate_about( const D3 & v, const D3 & H )
{
    const double x = dot3( H, H );
    double c, sinc, omc; // cos t, sin t / t, (1 - cos t)/t^2
    if( x < 0.0625 )
    {
        sinc = ROT_SINC[6];
        sinc = fma( sinc, x, ROT_SINC[5] );
        sinc = fma( sinc, x, ROT_SINC[4] );
        sinc = fma( sinc, x, ROT_SINC[3] );
        sinc = fma( sinc, x, ROT_SINC[2] );
        sinc = fma( sinc, x, ROT_SINC[1] );
        sinc = fma( sinc, x, 1.0 );
        omc  = ROT_OMC[6];
        omc  = fma( omc, x, ROT_OMC[5] );
        omc  = fma( omc, x, ROT_OMC[4] );
        omc  = fma( omc, x, ROT_OMC[3] );
        omc  = fma( omc, x, ROT_OMC[2] );
        omc  = fma( omc, x, ROT_OMC[1] );
        omc  = fma( omc, x, 0.5 );
        c    = fma( -x, omc, 1.0 );
    }
    else
    {
        const double t = sqrt( x );
        double sn;
        sincos( t, &sn, &c );
        sinc = sn / t;
        omc  = ( 1.0 - c ) / x;
    }
    const D3 Hxv    = cross3( H, v );
    const double hv = dot3( H, v ) * omc;
    return make_d3(
        fma( v.x, c, fma( Hxv.x, sinc, H.x * hv ) ), fma( v.y, c, fma( Hxv.y, sinc, H.y * hv ) ),
        fma( v.z, c, fma( Hxv.z, sinc, H.z * hv ) ) );
}

// Eigen's normalize(): leaves the zero vector untouched
__device__ __forceinline__ D3 normalized3( const D3 & v )
{
    const double n2 = dot3( v, v );
    if( n2 > 0 )
    {
        const double inv = rsqrt( n2 );
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

// ---------------------------------------------------------------------------------------------
// Per-site update rules of the explicit solvers, shared by the generic and the specialised stage kernels.
//   Depondt (Solver_Depondt.hpp:29-77)  stage 1: s' = R(Fv(s)) s          stage 2: s <- R((Fv(s)+Fv(s'))/2) s
//   Heun    (Solver_Heun.hpp:30-81)     stage 1: s' = |s - s x Fv(s)|     stage 2: s <- |s + k1/2 + k2/2|
//   SIB     (Solver_SIB.hpp:22-50)      stage 1: s' = (s + T(s,Fv(s)))/2  stage 2: s <- T(s, Fv(s'))
//   RK4     (Solver_RK4.hpp:41-147)     stages 1-4 with a running accumulator acc = k1/6 + k2/3 + k3/3
// si / Fv: configuration at the start of the iteration and its virtual force (where the stage needs it);
// spi / Fvp: the current predictor and its virtual force (stages >= 2). acc: RK4 accumulator, updated in place.
// ---------------------------------------------------------------------------------------------
template<int SOLVER, int STAGE>
__device__ __forceinline__ D3 solver_update( const D3 & si, const D3 & Fv, const D3 & spi, const D3 & Fvp, D3 & acc )
{
    if( SOLVER == Solver_Depondt )
    {
        if( STAGE == 1 )
            return rotate_about( si, Fv );
        return rotate_about( si, make_d3( 0.5 * ( Fv.x + Fvp.x ), 0.5 * ( Fv.y + Fvp.y ), 0.5 * ( Fv.z + Fvp.z ) ) );
    }
    else if( SOLVER == Solver_Heun )
    {
        const D3 k1 = cross3( Fv, si ); // -(s x Fv)
        if( STAGE == 1 )
            return normalized3( make_d3( si.x + k1.x, si.y + k1.y, si.z + k1.z ) );
        const D3 k2 = cross3( Fvp, spi ); // -(s' x Fv')
        return normalized3( make_d3(
            si.x + 0.5 * k1.x + 0.5 * k2.x, si.y + 0.5 * k1.y + 0.5 * k2.y, si.z + 0.5 * k1.z + 0.5 * k2.z ) );
    }
    else if( SOLVER == Solver_SIB )
    {
        if( STAGE == 1 )
        {
            const D3 t = sib_transform( si, Fv );
            return make_d3( 0.5 * ( t.x + si.x ), 0.5 * ( t.y + si.y ), 0.5 * ( t.z + si.z ) );
        }
        return sib_transform( si, Fvp );
    }
    else // RK4
    {
        // k_n = -(conf_n x Fv_n); intermediates are |s + c k_n| with c = 1/2, 1/2, 1
        const D3 k     = STAGE == 1 ? cross3( Fv, si ) : cross3( Fvp, spi );
        const double w = ( STAGE == 1 || STAGE == 4 ) ? 1.0 / 6.0 : 1.0 / 3.0;
        acc            = make_d3( acc.x + w * k.x, acc.y + w * k.y, acc.z + w * k.z );
        if( STAGE < 4 )
        {
            const double c = STAGE == 3 ? 1.0 : 0.5;
            return normalized3( make_d3( si.x + c * k.x, si.y + c * k.y, si.z + c * k.z ) );
        }
        return normalized3( make_d3( si.x + acc.x, si.y + acc.y, si.z + acc.z ) );
    }
}

// Which virtual forces a stage needs
template<int SOLVER, int STAGE>
struct StageNeeds
{
    static constexpr bool two_stage = SOLVER == Solver_Depondt || SOLVER == Solver_Heun || SOLVER == Solver_SIB;
    static constexpr bool last      = ( two_stage && STAGE == 2 ) || ( SOLVER == Solver_RK4 && STAGE == 4 );
    static constexpr bool Fv_s      = STAGE == 1 || ( STAGE == 2 && ( SOLVER == Solver_Depondt || SOLVER == Solver_Heun ) );
    static constexpr bool Fv_sp     = STAGE >= 2;
};

} // namespace dev
} // namespace sb
