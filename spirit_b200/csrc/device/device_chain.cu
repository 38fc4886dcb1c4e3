// GNEB on the device: the whole chain of images lives in HBM as [noi][field] allocations and every kernel is batched
// over images (blockIdx.y = image). Reference: Method_GNEB::Calculate_Force / Calculate_Force_Virtual /
// Hook_Post_Iteration (core/src/engine/Method_GNEB.cpp:87-456), Manifoldmath::Tangents / Geodesic_Tangent /
// dist_geodesic / project_* (core/src/engine/Manifoldmath.cpp:23-213), the solver templates over `noi` images
// (core/include/engine/Solver_VP.hpp:29-114, Solver_Depondt.hpp:29-77, Solver_Heun.hpp:30-81, Solver_SIB.hpp:22-50).
//
// One force evaluation of a configuration set {s_img}:
//   k_chain_gradient   F_g = -grad E(s_img) for every image, partial sums of E and of the squared geodesic
//                      distance to the previous image
//   k_reduce_rows      E[img], D2[img]                                   (deterministic two-level reductions)
//   k_chain_rx         Rx[img] = Rx[img-1] + sqrt(D2[img])
//   k_chain_tangent    t_img (energy-weighted secants, projected to the tangent planes; geodesic tangent at the two
//                      end images), partial sums of t.t and of P(F_g).t
//   k_reduce_rows      TT[img], FT[img]
//   k_chain_coeffs     per image: F = P(F_g) + c_t t   with c_t from the image type (normal: -FT/TT + spring; climbing:
//                      -2 FT/TT; falling: 0), F = 0 for the end images and stationary images
// and the total force is assembled inside the solver kernel that consumes it (VP velocity update, or the stage of
// Depondt / Heun / SIB), so F is written once and the projected gradient force, the spring force and the tangent
// normalisation never exist as separate fields (the reference keeps F_gradient, F_spring, F_total, tangents and two
// secant temporaries per image and sweeps each of them several times).
#include "device_buffers.cuh"

#include "../core/hamiltonian.hpp"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include "oso.cuh"

#include <cstring>
#include <stdexcept>

namespace sb
{
namespace dev
{

namespace
{

// Scalar slots on the device, each an array [noi]
enum Slot
{
    S_E = 0,  // energy of the last evaluated configuration
    S_D2,     // squared geodesic distance to the previous image
    S_RX,     // reaction coordinate
    S_TT,     // t.t
    S_FT,     // P(F_g).t = F_g.t
    S_PG2,    // |P(F_g)|^2                       (the three after S_FT are reduced together with it)
    S_TP2,    // |s(i+1) - s(i)|^2   of the images' spins (path shortening)
    S_TM2,    // |s(i) - s(i-1)|^2
    S_SHG,    // f . F_go   f = secant difference, F_go = gradient force orthogonal to the tangent (path shortening)
    S_SHT,    // f . t
    S_SHF,    // f . f
    S_CG,     // coefficient of P(F_g) in the total force (1; the rotational coefficient at moving end images)
    S_CT,     // coefficient of t in the total force
    S_CSH,    // path shortening: F += csh ( f - sha F_go - shb t )
    S_SHA,
    S_SHB,
    S_LEN,    // length of the path segment ending at this image in normalised (Rx, E) (energy-weighted springs)
    S_ZERO,   // 1: the total force of this image is zero (end image / stationary)
    S_TQ,     // max torque^2 (hook)
    S_VP,     // partial v.F per image
    S_VP2,    // partial F.F per image
    S_N_SLOTS
};
// global scalars behind the per-image slots: [0] ratio_prev, [1] ratio, [2] degenerate flag, [3] / [4] translation of the
// left / right end image along its tangent (translating endpoints)
constexpr int G_RATIO_PREV = 0, G_RATIO = 1, G_DEGENERATE = 2, G_TRANS_L = 3, G_TRANS_R = 4, G_N = 6;

struct ChainView
{
    double * S;   // configurations
    double * P;   // predictor configurations
    double * Fg;  // effective field -grad E (unprojected) of the last evaluation
    double * T;   // tangents (not normalised)
    double * F;   // total force of the first evaluation of the iteration ("forces")
    double * F2;  // total force of the predictor evaluation / new force of VP
    double * Fpr; // VP: force projected by the hook (F_prev of the next iteration)
    double * P2;  // RK4: second predictor configuration
    double * Acc; // RK4: running sum k1/6 + k2/3 + k3/3
    std::size_t stride; // doubles per image and field
    double * scal;      // [S_N_SLOTS][noi] + G_N   (noi = GLOBAL number of images)
    int noi;            // global number of images of the chain
    int n_local;        // images held by this rank: global indices [i_begin, i_begin + n_local)
    int i_begin;
    int moving_endpoints; // the end images feel a force too (Method_GNEB.cpp:261-355)
    int shortening;       // path shortening force on the normal images
    double * Ddi;         // dipolar gradient field of the evaluated configurations (null without dipolar interaction)
    const unsigned char * site_flags; // pinned sites / defects of the lattice (StencilParams::site_flags), or null
};
// Field pointers are offset by one image, so that local index -1 / n_local address the halo images received from the
// neighbouring ranks (images sharded over GPUs); unsharded chains have no halo and i_begin = 0.

__device__ __forceinline__ ConstField3 cfield( const double * base, std::size_t stride, int img )
{
    ConstField3 f;
    f.base = base + stride * img;
    return f;
}
__device__ __forceinline__ Field3 field( double * base, std::size_t stride, int img )
{
    Field3 f;
    f.base = base + stride * img;
    return f;
}
__device__ __forceinline__ double * slot( const ChainView & v, int s )
{
    return v.scal + std::size_t( s ) * v.noi;
}
__device__ __forceinline__ double * globals( const ChainView & v )
{
    return v.scal + std::size_t( S_N_SLOTS ) * v.noi;
}

// row r of slot k of `partials` ([slots][rows][n]) -> out[k * out_stride + r]; blockIdx = (row, slot): the slots of one
// evaluation are reduced by ONE launch (the reductions are a sixth of a GNEB iteration at 64 x 256^2, profiles/r1zz)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_reduce_rows( const double * __restrict__ partials, int n, double * __restrict__ out, int out_stride )
{
    const double * row = partials + ( std::size_t( blockIdx.y ) * gridDim.x + blockIdx.x ) * n;
    double v           = 0;
    for( int i = threadIdx.x; i < n; i += BLOCK_THREADS )
        v += row[i];
    v = block_sum( v );
    if( threadIdx.x == 0 )
        out[std::size_t( blockIdx.y ) * out_stride + blockIdx.x] = v;
}
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_reduce_rows_max( const double * __restrict__ partials, int n, double * __restrict__ out, int out_stride )
{
    const double * row = partials + ( std::size_t( blockIdx.y ) * gridDim.x + blockIdx.x ) * n;
    double v           = 0;
    for( int i = threadIdx.x; i < n; i += BLOCK_THREADS )
        v = fmax( v, row[i] );
    v = block_max( v );
    if( threadIdx.x == 0 )
        out[std::size_t( blockIdx.y ) * out_stride + blockIdx.x] = v;
}

// Gradient and energy of every image + squared geodesic distance to the previous image (Method_GNEB.cpp:95-128)
template<int NB_T>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_gradient(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v,
    const double * __restrict__ conf, double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;      // local image
    const int gi  = v.i_begin + img; // global image
    Site site;
    const bool active = locate_site( p, lg, site, NB_T > 0 ? NB_T : p.NB );
    double e = 0, d2 = 0;
    if( active )
    {
        const ConstField3 s  = cfield( conf, v.stride, img );
        // dipolar field of this image's configuration (computed before this kernel, one convolution per image); unused without
        const ConstField3 dd = cfield( v.Ddi ? v.Ddi : conf, v.stride, img );
        const D3 si          = load3( s, site.idx );
        if( NB_T == 1 && p.sc6 && !p.sc6_extras && !v.Ddi )
        {
            // nearest-neighbour structure (the images of the BASELINE chain): the six neighbours by index arithmetic, pairs +-r
            // combined (sc6.cuh) instead of the walk over the neighbour table; an open boundary contributes a zero spin
            const D3 zero = make_d3( 0, 0, 0 );
            auto neighbour = [&]( int axis, int dir ) -> D3
            {
                if( !p.sc6_axis[axis] )
                    return zero;
                const int n = axis == 0 ? p.Na : ( axis == 1 ? p.Nb : p.Nc );
                int q       = ( axis == 0 ? site.a : ( axis == 1 ? site.b : site.c ) ) + dir;
                if( q < 0 || q >= n )
                {
                    if( !p.bc[axis] )
                        return zero;
                    q = q < 0 ? q + n : q - n;
                }
                return load3( s, axis == 0 ? storage_index( p, q, site.b, site.c )
                                           : ( axis == 1 ? storage_index( p, site.a, q, site.c ) : storage_index( p, site.a, site.b, q ) ) );
            };
            const D3 g0 = make_d3( p.sc6_g0[0], p.sc6_g0[1], p.sc6_g0[2] );
            D3 g        = g0;
            sc6_axis_gradient<0, true>( p, neighbour( 0, -1 ), neighbour( 0, 1 ), g );
            sc6_axis_gradient<1, true>( p, neighbour( 1, -1 ), neighbour( 1, 1 ), g );
            if( p.sc6_axis[2] )
                sc6_axis_gradient<2, true>( p, neighbour( 2, -1 ), neighbour( 2, 1 ), g );
            g.x = fma( p.sc6_A[0], si.x, g.x );
            g.y = fma( p.sc6_A[1], si.y, g.y );
            g.z = fma( p.sc6_A[2], si.z, g.z );
            if( p.sc6_aniso_full )
            {
                g.x = fma( p.sc6_A[3], si.y, fma( p.sc6_A[4], si.z, g.x ) );
                g.y = fma( p.sc6_A[3], si.x, fma( p.sc6_A[5], si.z, g.y ) );
                g.z = fma( p.sc6_A[4], si.x, fma( p.sc6_A[5], si.y, g.z ) );
            }
            store3( field( v.Fg, v.stride, img ), site.idx, make_d3( -g.x, -g.y, -g.z ) );
            // E = 1/2 (g - g0) . s + g0 . s: the bilinear terms count half, the Zeeman term (g0 = -mu_s B n) whole
            e = 0.5 * ( ( g.x + g0.x ) * si.x + ( g.y + g0.y ) * si.y + ( g.z + g0.z ) * si.z );
        }
        else
        {
        const SiteGradient g = site_gradient<NB_T>( p, s, dd, site, si );
        const D3 gt          = total( g );
        store3( field( v.Fg, v.stride, img ), site.idx, make_d3( -gt.x, -gt.y, -gt.z ) );
        e = site_energy<NB_T>( p, site, si, g );
        }
        if( gi > 0 )
        {
            const D3 sp = load3( cfield( conf, v.stride, img - 1 ), site.idx );
            double r    = dot3( si, sp );
            r           = fmax( -1.0, fmin( 1.0, r ) ); // Vectormath::angle, Vectormath.cpp:434-441
            const double a = acos( r );
            d2             = a * a;
        }
    }
    e = block_sum( e );
    if( threadIdx.x == 0 )
        partials[( std::size_t( 0 ) * v.n_local + img ) * nblocks + blockIdx.x] = e;
    d2 = block_sum( d2 );
    if( threadIdx.x == 0 )
        partials[( std::size_t( 1 ) * v.n_local + img ) * nblocks + blockIdx.x] = d2;
}

// Rx and the degenerate-chain check (Method_GNEB.cpp:116-126)
static __global__ void k_chain_rx( const __grid_constant__ ChainView v )
{
    double * Rx       = slot( v, S_RX );
    const double * D2 = slot( v, S_D2 );
    Rx[0]             = 0;
    for( int i = 1; i < v.noi; ++i )
    {
        const double d = sqrt( D2[i] );
        Rx[i]          = Rx[i - 1] + d;
        if( d < 1e-10 )
            globals( v )[G_DEGENERATE] = 1.0;
    }
}

// Tangents (Manifoldmath.cpp:77-213)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_tangent(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v,
    const double * __restrict__ conf, double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;
    const int g   = v.i_begin + img;
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double tt = 0, ft = 0, pg2 = 0, tp2 = 0, tm2 = 0;
    if( active )
    {
        const D3 s = load3( cfield( conf, v.stride, img ), site.idx );
        D3 t;
        if( g == 0 || g == v.noi - 1 )
        {
            // Geodesic_Tangent at the end images: t = mid x (plus x minus)
            const D3 minus = g == 0 ? s : load3( cfield( conf, v.stride, img - 1 ), site.idx );
            const D3 plus  = g == 0 ? load3( cfield( conf, v.stride, img + 1 ), site.idx ) : s;
            D3 axis        = cross3( plus, minus );
            if( fabs( dot3( minus, plus ) + 1.0 ) < 1e-15 )
                axis = fabs( s.x - 1.0 ) > 1e-15 ? make_d3( 1, 0, 0 ) : make_d3( 0, 1, 0 );
            t = cross3( s, axis );
        }
        else
        {
            const D3 sp      = load3( cfield( conf, v.stride, img + 1 ), site.idx );
            const D3 sm      = load3( cfield( conf, v.stride, img - 1 ), site.idx );
            const D3 tp      = make_d3( sp.x - s.x, sp.y - s.y, sp.z - s.z );
            const D3 tm      = make_d3( s.x - sm.x, s.y - sm.y, s.z - sm.z );
            const double * E = slot( v, S_E );
            const double Em = E[g], Ep = E[g + 1], Emi = E[g - 1];
            double wp, wm;
            if( ( Ep < Em && Em > Emi ) || ( Ep > Em && Em < Emi ) )
            {
                const double Emax = fmax( fabs( Ep - Em ), fabs( Emi - Em ) );
                const double Emin = fmin( fabs( Ep - Em ), fabs( Emi - Em ) );
                wp                = Ep > Emi ? Emax : Emin;
                wm                = Ep > Emi ? Emin : Emax;
            }
            else if( Ep > Em && Em > Emi )
            {
                wp = 1;
                wm = 0;
            }
            else if( Ep < Em && Em < Emi )
            {
                wp = 0;
                wm = 1;
            }
            else
            {
                wp = 1;
                wm = 1;
            }
            t = make_d3( wp * tp.x + wm * tm.x, wp * tp.y + wm * tm.y, wp * tp.z + wm * tm.z );
            // project into the tangent plane of the spin
            const double d = dot3( t, s );
            t              = make_d3( t.x - d * s.x, t.y - d * s.y, t.z - d * s.z );
        }
        store3( field( v.T, v.stride, img ), site.idx, t );
        tt          = dot3( t, t );
        const D3 Fg = load3( cfield( v.Fg, v.stride, img ), site.idx );
        // projected gradient force . tangent (for interior images t is perpendicular to s, so this is also F_g.t)
        const double d = dot3( Fg, s );
        const D3 Pg    = make_d3( Fg.x - d * s.x, Fg.y - d * s.y, Fg.z - d * s.z );
        ft             = dot3( Pg, t );
        pg2            = dot3( Pg, Pg );
        if( g == 0 || g == v.noi - 1 )
            ft = dot3( Fg, t ); // dE/dRx at the end images uses the unprojected effective field (Method_GNEB.cpp:433-437)
        else if( v.shortening )
        {
            // secants of the IMAGES' spins (not of a predictor configuration: Method_GNEB.cpp:209-214 reads chain->images)
            const D3 s0 = load3( cfield( v.S, v.stride, img ), site.idx );
            const D3 sp = load3( cfield( v.S, v.stride, img + 1 ), site.idx );
            const D3 sm = load3( cfield( v.S, v.stride, img - 1 ), site.idx );
            const D3 a  = make_d3( sp.x - s0.x, sp.y - s0.y, sp.z - s0.z );
            const D3 b  = make_d3( s0.x - sm.x, s0.y - sm.y, s0.z - sm.z );
            tp2         = dot3( a, a );
            tm2         = dot3( b, b );
        }
    }
    const double vals[5] = { tt, ft, pg2, tp2, tm2 };
    for( int k = 0; k < ( v.shortening || v.moving_endpoints ? 5 : 2 ); ++k )
    {
        const double r = block_sum( vals[k] );
        if( threadIdx.x == 0 )
            partials[( std::size_t( k ) * v.n_local + img ) * nblocks + blockIdx.x] = r;
    }
}

// Path shortening (Method_GNEB.cpp:206-233), second pass: with the norms of the secants and of the gradient force known, the
// scalar products of f = t+/|t+| - t-/|t-| with the gradient force orthogonal to the tangent, with the tangent and with itself
__device__ __forceinline__ D3 chain_secant_difference( const ChainView & v, int img, std::size_t idx )
{
    const int g     = v.i_begin + img;
    const double ip = rsqrt( slot( v, S_TP2 )[g] ), im = rsqrt( slot( v, S_TM2 )[g] );
    const D3 s0     = load3( cfield( v.S, v.stride, img ), idx );
    const D3 sp     = load3( cfield( v.S, v.stride, img + 1 ), idx );
    const D3 sm     = load3( cfield( v.S, v.stride, img - 1 ), idx );
    return make_d3(
        ( sp.x - s0.x ) * ip - ( s0.x - sm.x ) * im, ( sp.y - s0.y ) * ip - ( s0.y - sm.y ) * im, ( sp.z - s0.z ) * ip - ( s0.z - sm.z ) * im );
}
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_shrink(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v,
    const double * __restrict__ conf, double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;
    const int g   = v.i_begin + img;
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double vals[3]    = { 0, 0, 0 };
    if( active && g > 0 && g < v.noi - 1 )
    {
        const D3 s      = load3( cfield( conf, v.stride, img ), site.idx );
        const D3 Fg     = load3( cfield( v.Fg, v.stride, img ), site.idx );
        const D3 t      = load3( cfield( v.T, v.stride, img ), site.idx );
        const double d  = dot3( Fg, s );
        const double c  = slot( v, S_FT )[g] / slot( v, S_TT )[g];
        const D3 Fgo    = make_d3( Fg.x - d * s.x - c * t.x, Fg.y - d * s.y - c * t.y, Fg.z - d * s.z - c * t.z );
        const D3 f      = chain_secant_difference( v, img, site.idx );
        vals[0]         = dot3( f, Fgo );
        vals[1]         = dot3( f, t );
        vals[2]         = dot3( f, f );
    }
    for( int k = 0; k < 3; ++k )
    {
        const double r = block_sum( vals[k] );
        if( threadIdx.x == 0 )
            partials[( std::size_t( k ) * v.n_local + img ) * nblocks + blockIdx.x] = r;
    }
}

// Translating endpoints (Method_GNEB.cpp:268-305): the two end images are pushed along their tangents by the mean of their
// gradient forces, the force of the other end rotated site by site from the other end's spin onto this one. One launch over
// the sites; partial sums of F_translation_left . t_0 and F_translation_right . t_last (t not normalised here).
__device__ __forceinline__ D3 rotate_between( const D3 & from, const D3 & to, const D3 & x, bool transpose )
{
    // rotation about from x to by the angle between them (Eigen::AngleAxis); identity for collinear spins
    const double c = dot3( from, to );
    if( fabs( c ) >= 1.0 )
        return x;
    D3 axis        = cross3( from, to );
    const double n = sqrt( dot3( axis, axis ) );
    if( n == 0 )
        return x;
    axis            = make_d3( axis.x / n, axis.y / n, axis.z / n );
    const double an = transpose ? -acos( c ) : acos( c );
    const double sn = sin( an ), cs = cos( an );
    const D3 axx    = cross3( axis, x );
    const double ad = dot3( axis, x ) * ( 1 - cs );
    return make_d3( x.x * cs + axx.x * sn + axis.x * ad, x.y * cs + axx.y * sn + axis.y * ad, x.z * cs + axx.z * sn + axis.z * ad );
}
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_translation(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v,
    const double * __restrict__ conf, double * __restrict__ partials, int nblocks )
{
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double vals[2]    = { 0, 0 };
    if( active )
    {
        const int last = v.noi - 1;
        const D3 cl = load3( cfield( conf, v.stride, 0 ), site.idx ), cr = load3( cfield( conf, v.stride, last ), site.idx );
        D3 Fl = load3( cfield( v.Fg, v.stride, 0 ), site.idx ), Fr = load3( cfield( v.Fg, v.stride, last ), site.idx );
        const double dl = dot3( Fl, cl ), dr = dot3( Fr, cr ); // gradient forces projected to the tangent planes of the configurations
        Fl = make_d3( Fl.x - dl * cl.x, Fl.y - dl * cl.y, Fl.z - dl * cl.z );
        Fr = make_d3( Fr.x - dr * cr.x, Fr.y - dr * cr.y, Fr.z - dr * cr.z );
        const D3 sl = load3( cfield( v.S, v.stride, 0 ), site.idx ), sr = load3( cfield( v.S, v.stride, last ), site.idx ); // images' spins
        const D3 Frr = rotate_between( sl, sr, Fr, false ), Flr = rotate_between( sl, sr, Fl, true );
        const D3 tl = load3( cfield( v.T, v.stride, 0 ), site.idx ), tr = load3( cfield( v.T, v.stride, last ), site.idx );
        vals[0]     = -0.5 * ( ( Fl.x + Frr.x ) * tl.x + ( Fl.y + Frr.y ) * tl.y + ( Fl.z + Frr.z ) * tl.z );
        vals[1]     = -0.5 * ( ( Flr.x + Fr.x ) * tr.x + ( Flr.y + Fr.y ) * tr.y + ( Flr.z + Fr.z ) * tr.z );
    }
    for( int k = 0; k < 2; ++k )
    {
        const double r = block_sum( vals[k] );
        if( threadIdx.x == 0 )
            partials[std::size_t( k ) * nblocks + blockIdx.x] = r;
    }
}

struct ChainTypes
{
    int type[256];
};

struct ChainForceParams
{
    double spring_constant, spring_force_ratio, path_shortening_constant;
    int moving_endpoints, translating_endpoints, escape_first;
    double delta_Rx0_left, delta_Rx0_right;
    int nos;
};

// Energy-weighted springs (Method_GNEB.cpp:137-170): lengths of the path segments in the normalised (Rx, E) plane from a cubic
// Hermite interpolation of the energy along the path (Cubic_Hermite_Spline.cpp:11-48) with 20 points per segment. One thread.
static __global__ void k_chain_lengths( const __grid_constant__ ChainView v, double ratio )
{
    const double * Rx = slot( v, S_RX );
    const double * E  = slot( v, S_E );
    double * len      = slot( v, S_LEN );
    const int noi = v.noi, n_int = 20;
    const double ratio_E = fmin( 1.0, ratio ), ratio_Rx = 1 - ratio_E;
    auto slope = [&]( int i ) { // dE/dRx as the reference takes it: effective field . normalised tangent
        const double tt = slot( v, S_TT )[i];
        return tt > 0 ? slot( v, S_FT )[i] / sqrt( tt ) : 0.0;
    };
    auto point = [&]( int i, int j, double & x, double & e ) { // interpolation point j of segment i (j = 0: image i)
        const double t   = j / double( n_int + 1 );
        const double t2 = t * t, t3 = t2 * t;
        const double h00 = 2 * t3 - 3 * t2 + 1, h10 = -2 * t3 + 3 * t2, h01 = t3 - 2 * t2 + t, h11 = t3 - t2;
        x                = Rx[i] + t * ( Rx[i + 1] - Rx[i] );
        e                = h00 * E[i] + h10 * E[i + 1] + h01 * slope( i ) * ( Rx[i] - Rx[i + 1] ) + h11 * slope( i + 1 ) * ( Rx[i] - Rx[i + 1] );
    };
    // ranges of the interpolated curve
    double xmin = Rx[noi - 1], xmax = Rx[noi - 1], emin = E[noi - 1], emax = E[noi - 1];
    for( int i = 0; i < noi - 1; ++i )
        for( int j = 0; j <= n_int; ++j )
        {
            double x, e;
            point( i, j, x, e );
            xmin = fmin( xmin, x ), xmax = fmax( xmax, x ), emin = fmin( emin, e ), emax = fmax( emax, e );
        }
    const double range_Rx = xmax - xmin, range_E = emax - emin;
    len[0] = 0;
    for( int img = 1; img < noi; ++img )
    {
        double l = 0, x0, e0;
        point( img - 1, 0, x0, e0 );
        for( int j = 1; j <= n_int + 1; ++j )
        {
            double x1, e1;
            if( j <= n_int )
                point( img - 1, j, x1, e1 );
            else if( img < noi - 1 )
                point( img, 0, x1, e1 );
            else
                x1 = Rx[noi - 1], e1 = E[noi - 1];
            // the reference sums i = 1 .. n_interpolations of its flat index: the last sub-interval of a segment (up to the
            // next image) belongs to the sum of that segment only through idx = (img - 1)(n + 1) + i <= img (n + 1) - 1
            if( j <= n_int )
            {
                const double dRx = ratio_Rx * ( x1 - x0 ) / range_Rx, dE = ratio_E * ( e1 - e0 ) / range_E;
                l += sqrt( dRx * dRx + dE * dE );
            }
            x0 = x1, e0 = e1;
        }
        len[img] = l * range_Rx;
    }
}

// Per-image coefficients of the total force (Method_GNEB.cpp:175-355), one thread per image:
//   F = cg P(F_g) + ct t + csh ( f - sha F_go - shb t )
static __global__ void k_chain_coeffs( const __grid_constant__ ChainView v, const __grid_constant__ ChainTypes types, const __grid_constant__ ChainForceParams P )
{
    const int img = blockIdx.x * blockDim.x + threadIdx.x; // local image; types are indexed locally
    if( img >= v.n_local )
        return;
    const int g = v.i_begin + img;
    double cg = 1, ct = 0, zero = 0, csh = 0, sha = 0, shb = 0;
    const double tt = slot( v, S_TT )[g], ft = slot( v, S_FT )[g];
    const double * Rx = slot( v, S_RX );
    const bool end    = g == 0 || g == v.noi - 1;
    if( end && P.moving_endpoints )
    {
        // rotational part of the gradient force + spring towards the equilibrium distance to the neighbour + translation
        double rc = 1;
        if( P.escape_first )
        {
            const double pl = slot( v, S_FT )[0] / sqrt( slot( v, S_TT )[0] );
            const double pr = slot( v, S_FT )[v.noi - 1] / sqrt( slot( v, S_TT )[v.noi - 1] );
            if( pl > pr )
                rc = 0;
        }
        const double nt         = sqrt( tt );
        const double projection = ft / nt; // F_gradient . normalised tangent
        const double delta_Rx0  = g == 0 ? P.delta_Rx0_left : P.delta_Rx0_right;
        const double delta_Rx   = g == 0 ? Rx[1] - Rx[0] : Rx[v.noi - 1] - Rx[v.noi - 2];
        const double k          = ( g == 0 ? 1.0 : -1.0 ) * P.spring_constant;
        // project_parallel( F_translation, normalised tangent ): (sum F_translation . t / |t|) t / |t|
        const double alpha = P.translating_endpoints ? globals( v )[g == 0 ? G_TRANS_L : G_TRANS_R] / tt : 0.0;
        cg                 = rc;
        ct                 = ( -rc * projection + k * ( delta_Rx - delta_Rx0 ) ) / nt + alpha;
    }
    else if( end || types.type[img] == 3 )
        zero = 1;
    else if( types.type[img] == 1 ) // climbing: invert the component along the tangent
        ct = -2.0 * ft / tt;
    else if( types.type[img] == 2 ) // falling: gradient force only
        ct = 0;
    else
    {
        // normal: orthogonal to the tangent + spring force along it
        const double d = P.spring_force_ratio > 0 ? P.spring_constant * ( slot( v, S_LEN )[g + 1] - slot( v, S_LEN )[g] )
                                                  : P.spring_constant * ( Rx[g + 1] - 2 * Rx[g] + Rx[g - 1] );
        ct             = -ft / tt + d / sqrt( tt );
        if( P.path_shortening_constant > 0 )
        {
            // f projected orthogonally to the normalised gradient force and to the normalised tangent (orthogonal to each
            // other), normalised, scaled by max( |F_go|, nos * constant )
            const double go2      = slot( v, S_PG2 )[g] - ft * ft / tt; // |F_go|^2
            const double gradnorm = sqrt( go2 );
            const double fg = slot( v, S_SHG )[g], fT = slot( v, S_SHT )[g], ff = slot( v, S_SHF )[g];
            sha               = fg / go2; // ( f . F_go / |F_go| ) F_go / |F_go|
            shb               = fT / tt;
            const double n2   = ff - fg * fg / go2 - fT * fT / tt;
            csh               = fmax( gradnorm, P.nos * P.path_shortening_constant ) / sqrt( n2 );
        }
    }
    slot( v, S_CG )[g]   = cg;
    slot( v, S_CT )[g]   = ct;
    slot( v, S_CSH )[g]  = csh;
    slot( v, S_SHA )[g]  = sha;
    slot( v, S_SHB )[g]  = shb;
    slot( v, S_ZERO )[g] = zero;
}

// Total force of a site from the effective field, the tangent and the image coefficients
__device__ __forceinline__ D3 chain_total_force( const ChainView & v, int img, std::size_t idx, const D3 & s )
{
    const int g = v.i_begin + img;
    if( slot( v, S_ZERO )[g] != 0.0 )
        return make_d3( 0, 0, 0 );
    if( v.site_flags && ( __ldg( v.site_flags + idx ) & FLAG_PINNED ) ) // pinning mask on F_total (Method_GNEB.cpp:251-254)
        return make_d3( 0, 0, 0 );
    const D3 Fg     = load3( cfield( v.Fg, v.stride, img ), idx );
    const D3 t      = load3( cfield( v.T, v.stride, img ), idx );
    const double d  = dot3( Fg, s );
    const double cg = slot( v, S_CG )[g], ct = slot( v, S_CT )[g];
    const D3 Pg     = make_d3( Fg.x - d * s.x, Fg.y - d * s.y, Fg.z - d * s.z );
    D3 F            = make_d3( cg * Pg.x + ct * t.x, cg * Pg.y + ct * t.y, cg * Pg.z + ct * t.z );
    const double csh = slot( v, S_CSH )[g];
    if( csh != 0.0 )
    {
        const double c = slot( v, S_FT )[g] / slot( v, S_TT )[g], sha = slot( v, S_SHA )[g], shb = slot( v, S_SHB )[g];
        const D3 f     = chain_secant_difference( v, img, idx );
        F.x += csh * ( f.x - sha * ( Pg.x - c * t.x ) - shb * t.x );
        F.y += csh * ( f.y - sha * ( Pg.y - c * t.y ) - shb * t.y );
        F.z += csh * ( f.z - sha * ( Pg.z - c * t.z ) - shb * t.z );
    }
    return F;
}

// VP, part A (Solver_VP.hpp:29-79): F_new, v = ratio_prev F_old + (F_prev + F_new)/2, partial sums of v.F_new, F_new.F_new
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_vp_a(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v,
    const double * __restrict__ F_prev, double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double proj = 0, norm2 = 0;
    if( active )
    {
        const double ratio_prev = globals( v )[G_RATIO_PREV];
        const D3 s              = load3( cfield( v.S, v.stride, img ), site.idx );
        const D3 Fn             = chain_total_force( v, img, site.idx, s );
        const D3 Fo             = load3( cfield( v.F, v.stride, img ), site.idx );    // raw force of the last iteration
        const D3 Fp             = load3( cfield( F_prev, v.stride, img ), site.idx ); // the same, or projected by the hook
        const D3 vel            = make_d3(
            ratio_prev * Fo.x + 0.5 * ( Fp.x + Fn.x ), ratio_prev * Fo.y + 0.5 * ( Fp.y + Fn.y ),
            ratio_prev * Fo.z + 0.5 * ( Fp.z + Fn.z ) );
        proj  = dot3( vel, Fn );
        norm2 = dot3( Fn, Fn );
        store3( field( v.F2, v.stride, img ), site.idx, Fn );
    }
    proj = block_sum( proj );
    if( threadIdx.x == 0 )
        partials[( std::size_t( 0 ) * v.n_local + img ) * nblocks + blockIdx.x] = proj;
    norm2 = block_sum( norm2 );
    if( threadIdx.x == 0 )
        partials[( std::size_t( 1 ) * v.n_local + img ) * nblocks + blockIdx.x] = norm2;
}

// sums over images -> ratio (Solver_VP.hpp:75-103)
static __global__ void k_chain_vp_ratio( const __grid_constant__ ChainView v )
{
    double proj = 0, norm2 = 0;
    for( int i = 0; i < v.noi; ++i )
    {
        proj += slot( v, S_VP )[i];
        norm2 += slot( v, S_VP2 )[i];
    }
    const double ratio        = proj > 0 ? proj / norm2 : 0.0;
    globals( v )[G_RATIO]      = ratio;
    globals( v )[G_RATIO_PREV] = ratio;
}

// VP, part B: s <- |s + dt ratio F + dt F / 2| for every image; HOOK: max torque and the in-place projection of the
// forces (Method_GNEB.cpp:416-428)
template<bool HOOK>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_vp_b(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v, double dt,
    double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double t2         = 0;
    if( active )
    {
        const double ratio = globals( v )[G_RATIO];
        const Field3 S     = field( v.S, v.stride, img );
        const D3 s         = load3( S, site.idx );
        D3 F               = load3( cfield( v.F2, v.stride, img ), site.idx );
        const double c     = dt * ratio + 0.5 * dt;
        const D3 sn        = normalized3( make_d3( s.x + c * F.x, s.y + c * F.y, s.z + c * F.z ) );
        store3( S, site.idx, sn );
        if( HOOK )
        {
            const double f = dot3( F, sn );
            F              = make_d3( F.x - f * sn.x, F.y - f * sn.y, F.z - f * sn.z );
            t2             = dot3( F, F );
            store3( field( v.Fpr, v.stride, img ), site.idx, F );
        }
    }
    if( HOOK )
    {
        t2 = block_max( t2 );
        if( threadIdx.x == 0 )
            partials[std::size_t( img ) * nblocks + blockIdx.x] = t2;
    }
}

// Total force of every image at configuration `conf` into F2. Needed before a kernel that overwrites the images' spins when the
// force reads the NEIGHBOURING images' spins (path shortening: the secants are taken between the images, Method_GNEB.cpp:209-214):
// computed inside the updating kernel, a thread could read a neighbour's spin that is already updated.
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_force_at(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v, const double * __restrict__ conf )
{
    const int img = blockIdx.y;
    Site site;
    if( !locate_site( p, lg, site, p.NB ) )
        return;
    const D3 c = load3( cfield( conf, v.stride, img ), site.idx );
    store3( field( v.F2, v.stride, img ), site.idx, chain_total_force( v, img, site.idx, c ) );
}

// Stages of the two-stage solvers over all images. Virtual force: Fv = dtg s x F, zero for the end images
// (Method_GNEB.cpp:359-391). Stage 1 assembles F(s) and writes the predictor; stage 2 assembles F(s') and updates s.
template<int SOLVER, int STAGE>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_stage(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v, double dtg )
{
    const int img = blockIdx.y;
    Site site;
    if( !locate_site( p, lg, site, p.NB ) )
        return;
    const int g    = v.i_begin + img;
    const bool end = !v.moving_endpoints && ( g == 0 || g == v.noi - 1 );
    const D3 s     = load3( cfield( v.S, v.stride, img ), site.idx );
    D3 acc         = make_d3( 0, 0, 0 );
    if( STAGE == 1 )
    {
        const D3 F = chain_total_force( v, img, site.idx, s );
        store3( field( v.F, v.stride, img ), site.idx, F );
        D3 Fv = make_d3( 0, 0, 0 );
        if( !end )
        {
            const D3 c = cross3( s, F );
            Fv         = make_d3( dtg * c.x, dtg * c.y, dtg * c.z );
        }
        store3( field( v.P, v.stride, img ), site.idx, solver_update<SOLVER, 1>( s, Fv, s, Fv, acc ) );
    }
    else
    {
        const D3 sp = load3( cfield( v.P, v.stride, img ), site.idx );
        // with path shortening F2 was computed by k_chain_force_at before this launch (it reads the neighbours' spins)
        const D3 F2 = v.shortening ? load3( cfield( v.F2, v.stride, img ), site.idx ) : chain_total_force( v, img, site.idx, sp );
        if( !v.shortening )
            store3( field( v.F2, v.stride, img ), site.idx, F2 );
        D3 Fv = make_d3( 0, 0, 0 ), Fvp = make_d3( 0, 0, 0 );
        if( !end )
        {
            const D3 F1 = load3( cfield( v.F, v.stride, img ), site.idx );
            const D3 c1 = cross3( s, F1 );
            Fv          = make_d3( dtg * c1.x, dtg * c1.y, dtg * c1.z );
            const D3 c2 = cross3( sp, F2 );
            Fvp         = make_d3( dtg * c2.x, dtg * c2.y, dtg * c2.z );
        }
        store3( field( v.S, v.stride, img ), site.idx, solver_update<SOLVER, 2>( s, Fv, sp, Fvp, acc ) );
    }
}

// RK4 over all images (Solver_RK4.hpp:41-147): stage n evaluates the force at configuration n (s, s + k1/2, s + k2/2, s + k3),
// k_n = -(conf_n x Fv_n) with Fv = dtg conf x F (zero at the end images), and the new spins are |s + k1/6 + k2/3 + k3/3 + k4/6|.
// `conf` is the configuration the force was just evaluated at, `out` the configuration this stage writes.
template<int STAGE>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_rk4(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v, double dtg,
    const double * __restrict__ conf, double * __restrict__ out )
{
    const int img = blockIdx.y;
    Site site;
    if( !locate_site( p, lg, site, p.NB ) )
        return;
    const int g    = v.i_begin + img;
    const bool end = !v.moving_endpoints && ( g == 0 || g == v.noi - 1 );
    const D3 s     = load3( cfield( v.S, v.stride, img ), site.idx );
    const D3 c     = STAGE == 1 ? s : load3( cfield( conf, v.stride, img ), site.idx );
    const D3 F     = ( STAGE == 4 && v.shortening ) ? load3( cfield( v.F2, v.stride, img ), site.idx ) // k_chain_force_at ran before
                                                    : chain_total_force( v, img, site.idx, c );
    if( STAGE == 1 )
        store3( field( v.F, v.stride, img ), site.idx, F ); // "forces": the hook projects the force of the first evaluation
    if( STAGE == 4 && !v.shortening )
        store3( field( v.F2, v.stride, img ), site.idx, F ); // F_total of the last evaluation: max torque
    D3 Fv = make_d3( 0, 0, 0 );
    if( !end )
    {
        const D3 x = cross3( c, F );
        Fv         = make_d3( dtg * x.x, dtg * x.y, dtg * x.z );
    }
    D3 acc = make_d3( 0, 0, 0 );
    if( STAGE > 1 )
        acc = load3( cfield( v.Acc, v.stride, img ), site.idx );
    const D3 o = solver_update<Solver_RK4, STAGE>( s, Fv, c, Fv, acc );
    if( STAGE < 4 )
        store3( field( v.Acc, v.stride, img ), site.idx, acc );
    store3( field( out, v.stride, img ), site.idx, o );
}

// OSO / atlas minimisers over the chain: total force of every image into F2 (hook, atlas gradient) and s x F into F
// (the OSO gradient is T(+-s x F), oso.cuh). End / stationary images have F = 0 and do not move.
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_chain_cross( const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v )
{
    const int img = blockIdx.y;
    Site site;
    if( !locate_site( p, lg, site, p.NB ) )
        return;
    const D3 s = load3( cfield( v.S, v.stride, img ), site.idx );
    const D3 F = chain_total_force( v, img, site.idx, s );
    store3( field( v.F2, v.stride, img ), site.idx, F );
    store3( field( v.F, v.stride, img ), site.idx, cross3( s, F ) );
}

// Hook of the two-stage solvers: max_i |F_total - (F_total.s)s| per image with the NEW spins, F_total = force of the last
// (predictor) evaluation (Method_GNEB.cpp:416-424)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_chain_hook(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ ChainView v,
    double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double t2         = 0;
    if( active )
    {
        const D3 s     = load3( cfield( v.S, v.stride, img ), site.idx );
        const D3 F     = load3( cfield( v.F2, v.stride, img ), site.idx );
        const double f = dot3( F, s );
        const D3 q     = make_d3( F.x - f * s.x, F.y - f * s.y, F.z - f * s.z );
        t2             = dot3( q, q );
    }
    t2 = block_max( t2 );
    if( threadIdx.x == 0 )
        partials[std::size_t( img ) * nblocks + blockIdx.x] = t2;
}

} // namespace

// ---------------------------------------------------------------------------------------------
struct DeviceChainBuffers
{
    cudaStream_t stream = nullptr;
    double * fields     = nullptr; // 7 x [noi][stride]
    double * ddi        = nullptr; // dipolar gradient fields of the evaluated configurations [noi + 2][stride] (on demand)
    double * rk4        = nullptr; // RK4: second predictor and accumulator, 2 x [noi + 2][stride] (on demand)
    double * partials   = nullptr; // [5][noi][nblocks]
    double * scal       = nullptr;
    double * h_scal     = nullptr; // pinned mirror
    double * staging    = nullptr; // AoS [nos][3]
    std::size_t stride  = 0;
    std::size_t n_scal  = 0;
    ChainView view{};

    ~DeviceChainBuffers()
    {
        if( fields )
            cudaFree( fields );
        if( ddi )
            cudaFree( ddi );
        if( rk4 )
            cudaFree( rk4 );
        if( partials )
            cudaFree( partials );
        if( scal )
            cudaFree( scal );
        if( staging )
            cudaFree( staging );
        if( h_scal )
            cudaFreeHost( h_scal );
        if( stream )
            cudaStreamDestroy( stream );
    }
};

DeviceChain::DeviceChain( const Geometry & geometry, int noi, int i_begin, int noi_global )
        : noi_( noi ), i_begin_( i_begin ), noi_global_( noi_global < 0 ? noi : noi_global )
{
    require_device();
    if( noi_global_ > 256 )
        throw std::runtime_error( "spirit_b200: chains of more than 256 images are not supported" );
    sharded_ = noi_global_ != noi_;
    if( sharded_ && !comm_active() )
        throw std::runtime_error( "spirit_b200: a sharded chain needs an initialised communicator (SpiritB200_Comm_Init)" );
    if( i_begin_ < 0 || i_begin_ + noi_ > noi_global_ )
        throw std::runtime_error( "spirit_b200: chain shard outside of the global chain" );
    table_ = std::make_unique<DeviceImage>( geometry );
    nos_   = table_->nos();
    buf_   = std::make_unique<DeviceChainBuffers>();
    auto & b       = *buf_;
    const auto & T = *table_->buffers();
    b.stride       = 3 * T.n_storage;
    b.n_scal       = std::size_t( S_N_SLOTS ) * noi_global_ + G_N;
    // every field holds the local images plus one halo image on each side (neighbour images of the adjacent ranks)
    const std::size_t fs = std::size_t( noi + 2 ) * b.stride;
    SB_CUDA_CHECK( cudaStreamCreateWithFlags( &b.stream, cudaStreamNonBlocking ) );
    SB_CUDA_CHECK( cudaMalloc( &b.fields, 7 * fs * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMemset( b.fields, 0, 7 * fs * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMalloc( &b.partials, 5 * std::size_t( noi ) * T.nblocks * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMalloc( &b.scal, b.n_scal * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMemset( b.scal, 0, b.n_scal * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaHostAlloc( &b.h_scal, b.n_scal * sizeof( double ), cudaHostAllocDefault ) );
    SB_CUDA_CHECK( cudaMalloc( &b.staging, 3 * std::size_t( nos_ ) * sizeof( double ) ) );
    b.view.S       = b.fields + 0 * fs + b.stride; // + stride: local image 0 (index -1 is the lower halo image)
    b.view.P       = b.fields + 1 * fs + b.stride;
    b.view.Fg      = b.fields + 2 * fs + b.stride;
    b.view.T       = b.fields + 3 * fs + b.stride;
    b.view.F       = b.fields + 4 * fs + b.stride;
    b.view.F2      = b.fields + 5 * fs + b.stride;
    b.view.Fpr     = b.fields + 6 * fs + b.stride;
    b.view.stride  = b.stride;
    b.view.scal    = b.scal;
    b.view.noi     = noi_global_;
    b.view.n_local = noi_;
    b.view.i_begin = i_begin_;
    b.view.moving_endpoints = 0;
    b.view.shortening       = 0;
    b.view.Ddi              = nullptr;
    b.view.site_flags       = nullptr;
    b.view.P2               = nullptr;
    b.view.Acc              = nullptr;
}

void DeviceChain::ensure_ddi_field()
{
    auto & b = *buf_;
    if( b.ddi )
        return;
    const std::size_t fs = std::size_t( noi_ + 2 ) * b.stride;
    SB_CUDA_CHECK( cudaMalloc( &b.ddi, fs * sizeof( double ) ) );
    SB_CUDA_CHECK( cudaMemset( b.ddi, 0, fs * sizeof( double ) ) );
}

DeviceChain::~DeviceChain() = default;

void DeviceChain::set_hamiltonian( const Hamiltonian & ham )
{
    // every image is evaluated with this Hamiltonian (Method_GNEB.cpp:99-100 calls each image's own; they are copies of one
    // another in a chain); with a dipolar term one convolution per image and evaluation fills the field the stencil adds
    table_->set_hamiltonian( ham );
    buf_->view.site_flags = table_->stencil().site_flags; // (all images of the chain have image 0's pinned sites and defects)
    if( table_->stencil().has_ddi )
    {
        if( sharded_ )
            throw std::runtime_error( "spirit_b200: GNEB with dipole-dipole interaction on a chain sharded over GPUs is not implemented" );
        ensure_ddi_field();
        buf_->view.Ddi = buf_->ddi + buf_->stride;
    }
    else
        buf_->view.Ddi = nullptr;
}

void DeviceChain::synchronize()
{
    SB_CUDA_CHECK( cudaStreamSynchronize( buf_->stream ) );
}

void DeviceChain::upload_image( int img, const double * host_aos )
{
    auto & b       = *buf_;
    const auto & p = table_->stencil();
    SB_CUDA_CHECK( cudaMemcpyAsync( b.staging, host_aos, 3 * std::size_t( nos_ ) * sizeof( double ), cudaMemcpyHostToDevice, b.stream ) );
    Field3 f;
    f.base = b.view.S + b.stride * img;
    k_aos_to_soa<<<( nos_ + BLOCK_THREADS - 1 ) / BLOCK_THREADS, BLOCK_THREADS, 0, b.stream>>>(
        b.staging, f, nos_, table_->buffers()->plane_sites, p.plane_stride, p.halo );
    ++launches_;
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) ); // the staging buffer is reused by the next image
}

static void chain_download( DeviceChainBuffers & b, const DeviceImage & table, const double * base, int img, double * host_aos, int nos )
{
    const auto & p = table.stencil();
    ConstField3 f;
    f.base = base + b.stride * img;
    k_soa_to_aos<<<( nos + BLOCK_THREADS - 1 ) / BLOCK_THREADS, BLOCK_THREADS, 0, b.stream>>>(
        f, b.staging, nos, const_cast<DeviceImage &>( table ).buffers()->plane_sites, p.plane_stride, p.halo, 1.0 );
    SB_CUDA_CHECK( cudaGetLastError() );
    SB_CUDA_CHECK( cudaMemcpyAsync( host_aos, b.staging, 3 * std::size_t( nos ) * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
}

void DeviceChain::download_image( int img, double * host_aos )
{
    chain_download( *buf_, *table_, buf_->view.S, img, host_aos, nos_ );
    ++launches_;
}

void DeviceChain::download_effective_field( int img, double * host_aos )
{
    chain_download( *buf_, *table_, buf_->view.Fg, img, host_aos, nos_ );
    ++launches_;
}

void DeviceChain::vp_reset()
{
    oso_.reset();
    auto & b = *buf_;
    // velocity = 0, F = F_prev = 0 (Method_GNEB's constructor evaluates no force, Method_GNEB.cpp:23-73)
    const std::size_t fs = std::size_t( noi_ ) * b.stride * sizeof( double );
    SB_CUDA_CHECK( cudaMemsetAsync( b.view.F, 0, fs, b.stream ) );
    SB_CUDA_CHECK( cudaMemsetAsync( b.view.F2, 0, fs, b.stream ) );
    SB_CUDA_CHECK( cudaMemsetAsync( b.view.Fpr, 0, fs, b.stream ) );
    SB_CUDA_CHECK( cudaMemsetAsync( b.scal + std::size_t( S_N_SLOTS ) * noi_global_, 0, G_N * sizeof( double ), b.stream ) );
    vp_prev_projected_ = false;
}

// Sharded chain: the first / last local image of `field` goes to the lower / upper rank, whose last / first image arrives
// in the halo slots (local index -1 / n_local). A chain is open: the end ranks have one neighbour only.
void DeviceChain::exchange_halo_images( double * field_base )
{
    if( !sharded_ )
        return;
    auto & b          = *buf_;
    const int rank = comm_rank(), world = comm_world();
    const int lower = rank > 0 ? rank - 1 : -1, upper = rank < world - 1 ? rank + 1 : -1;
    comm_sendrecv(
        field_base, field_base + b.stride * ( noi_ - 1 ), field_base - std::ptrdiff_t( b.stride ), field_base + b.stride * noi_, b.stride,
        lower, upper, b.stream );
}

// rows [slot][local image] written by a reduction -> every rank gets the values of ALL images
void DeviceChain::share_slots( int first_slot, int n_slots, bool max )
{
    if( !sharded_ )
        return;
    auto & b = *buf_;
    comm_allreduce( b.scal + std::size_t( first_slot ) * noi_global_, std::size_t( n_slots ) * noi_global_, max, b.stream );
}

// which_configuration: 0 = S, 1 = P. Leaves F_g, T and the per-image coefficients on the device.
void DeviceChain::evaluate_force( const GNEBParams & params, int which_configuration, int )
{
    auto & b          = *buf_;
    const auto & T    = *table_->buffers();
    const auto & p    = table_->stencil();
    double * cf       = which_configuration == 0 ? b.view.S : ( which_configuration == 1 ? b.view.P : b.view.P2 );
    const dim3 grid( T.nblocks, noi_ );
    b.view.moving_endpoints = params.moving_endpoints ? 1 : 0;
    b.view.shortening       = params.path_shortening_constant > 0 ? 1 : 0;
    if( sharded_ && ( params.moving_endpoints || params.path_shortening_constant > 0 || params.spring_force_ratio > 0 ) )
        throw std::runtime_error( "spirit_b200: moving endpoints, path shortening and energy-weighted springs are not implemented on a chain "
                                  "sharded over GPUs" );
    exchange_halo_images( cf );
    if( b.view.Ddi )
        for( int img = 0; img < noi_; ++img ) // one dipolar convolution per image (shared plan: one after the other)
            table_->ddi_gradient_of( cf + b.stride * img, b.view.Ddi + b.stride * img, b.stream );
    if( p.NB == 1 )
        k_chain_gradient<1><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, cf, b.partials, T.nblocks );
    else
        k_chain_gradient<0><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, cf, b.partials, T.nblocks );
    // two adjacent slots (E, D2): rows [slot][local image] -> scal[slot][i_begin + image]
    reduce_to_slots( S_E, 2, false );
    k_chain_rx<<<1, 1, 0, b.stream>>>( b.view );
    k_chain_tangent<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, cf, b.partials, T.nblocks );
    reduce_to_slots( S_TT, ( b.view.shortening || b.view.moving_endpoints ) ? 5 : 2, false );
    launches_ += 6;
    if( b.view.shortening )
    {
        k_chain_shrink<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, cf, b.partials, T.nblocks );
        reduce_to_slots( S_SHG, 3, false );
        launches_ += 2;
    }
    if( params.moving_endpoints && params.translating_endpoints )
    {
        k_chain_translation<<<dim3( T.nblocks, 1 ), BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, cf, b.partials, T.nblocks );
        k_reduce_rows<<<dim3( 1, 2 ), BLOCK_THREADS, 0, b.stream>>>(
            b.partials, T.nblocks, b.scal + std::size_t( S_N_SLOTS ) * noi_global_ + G_TRANS_L, 1 );
        launches_ += 2;
    }
    if( params.spring_force_ratio > 0 )
    {
        k_chain_lengths<<<1, 1, 0, b.stream>>>( b.view, params.spring_force_ratio );
        ++launches_;
    }
    ChainTypes types{};
    for( int i = 0; i < noi_; ++i )
        types.type[i] = i < int( params.image_type.size() ) ? params.image_type[i] : 0;
    ChainForceParams P{};
    P.spring_constant          = params.spring_constant;
    P.spring_force_ratio       = params.spring_force_ratio;
    P.path_shortening_constant = params.path_shortening_constant;
    P.moving_endpoints         = params.moving_endpoints ? 1 : 0;
    P.translating_endpoints    = params.translating_endpoints ? 1 : 0;
    P.escape_first             = params.escape_first ? 1 : 0;
    P.delta_Rx0_left           = params.equilibrium_delta_Rx_left;
    P.delta_Rx0_right          = params.equilibrium_delta_Rx_right;
    P.nos                      = nos_;
    k_chain_coeffs<<<( noi_ + 63 ) / 64, 64, 0, b.stream>>>( b.view, types, P );
    SB_CUDA_CHECK( cudaGetLastError() );
}

// partial rows [n_slots][n_local][nblocks] -> slots first_slot .. first_slot + n_slots - 1, entries of the local images;
// sharded: zero the rest and all-reduce so that every rank holds all images' values
void DeviceChain::reduce_to_slots( int first_slot, int n_slots, bool max )
{
    auto & b       = *buf_;
    const auto & T = *table_->buffers();
    if( sharded_ )
        SB_CUDA_CHECK( cudaMemsetAsync( b.scal + std::size_t( first_slot ) * noi_global_, 0, std::size_t( n_slots ) * noi_global_ * sizeof( double ), b.stream ) );
    double * out = b.scal + std::size_t( first_slot ) * noi_global_ + i_begin_;
    if( max )
        k_reduce_rows_max<<<dim3( noi_, n_slots ), BLOCK_THREADS, 0, b.stream>>>( b.partials, T.nblocks, out, noi_global_ );
    else
        k_reduce_rows<<<dim3( noi_, n_slots ), BLOCK_THREADS, 0, b.stream>>>( b.partials, T.nblocks, out, noi_global_ );
    share_slots( first_slot, n_slots, max );
}

namespace
{
template<int SOLVER>
void launch_chain_stages(
    DeviceChain & chain, DeviceChainBuffers & b, const DeviceBuffers & T, const StencilParams & p, int noi, double dtg, int stage )
{
    const dim3 grid( T.nblocks, noi );
    if( stage == 1 )
        k_chain_stage<SOLVER, 1><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, dtg );
    else
        k_chain_stage<SOLVER, 2><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, dtg );
}
} // namespace

void DeviceChain::iterate( int solver, const GNEBParams & params, int n_iterations, bool hook, ChainHookResult * result )
{
    auto & b       = *buf_;
    const auto & T = *table_->buffers();
    const auto & p = table_->stencil();
    const dim3 grid( T.nblocks, noi_ );
    const bool oso = solver == Solver_VP_OSO || solver == Solver_LBFGS_OSO || solver == Solver_LBFGS_Atlas;
    if( solver != Solver_VP && solver != Solver_Depondt && solver != Solver_Heun && solver != Solver_SIB && solver != Solver_RK4 && !oso )
        throw std::runtime_error( "spirit_b200: GNEB solver id " + std::to_string( solver ) + " is not implemented" );
    if( solver == Solver_RK4 && !b.rk4 )
    {
        const std::size_t fs = std::size_t( noi_ + 2 ) * b.stride;
        SB_CUDA_CHECK( cudaMalloc( &b.rk4, 2 * fs * sizeof( double ) ) );
        SB_CUDA_CHECK( cudaMemset( b.rk4, 0, 2 * fs * sizeof( double ) ) );
        b.view.P2  = b.rk4 + b.stride;
        b.view.Acc = b.rk4 + fs + b.stride;
    }
    // images sharded over GPUs: the dot products of the minimisers run over all images of the chain (all-reduce). The atlas
    // solver's chart check compares the spins of IMAGE 0 with the charts of every image (Solver_Kernels.cpp:155-184), which only
    // the rank holding image 0 could do: refused.
    if( solver == Solver_LBFGS_Atlas && sharded_ )
        throw std::runtime_error( "spirit_b200: LBFGS_Atlas is not implemented on a chain sharded over GPUs" );
    // all local images back to back: one long field for the element-wise passes of oso.cuh
    const OsoLayout L{ std::size_t( noi_ ) * T.n_storage, p.plane_stride, T.plane_sites };
    if( oso && !oso_ )
    {
        oso_.reset( new OsoState );
        oso_->distributed = sharded_;
        oso_->allocate( solver, L.n_sites, noi_, ConstField3{ b.view.S }, L, b.stream, launches_ );
    }

    for( int it = 0; it < n_iterations; ++it )
    {
        const bool hk = hook && it == n_iterations - 1;
        if( oso )
        {
            // Solver_VP_OSO.hpp:34-115 / Solver_LBFGS_OSO.hpp:39-77 / Solver_LBFGS_Atlas.hpp:50-107 with noi images
            evaluate_force( params, 0, 0 );
            k_chain_cross<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view );
            ++launches_;
            oso_update( *oso_, solver, Field3{ b.view.S }, ConstField3{ b.view.F }, 1.0, ConstField3{ b.view.F2 }, L, nos_, params.dt, b.stream, launches_ );
            if( hk )
            {
                k_chain_hook<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, b.partials, T.nblocks );
                reduce_to_slots( S_TQ, 1, true );
                launches_ += 2;
            }
        }
        else if( solver == Solver_VP )
        {
            evaluate_force( params, 0, 0 );
            const double * F_prev = vp_prev_projected_ ? b.view.Fpr : b.view.F;
            k_chain_vp_a<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, F_prev, b.partials, T.nblocks );
            reduce_to_slots( S_VP, 2, false );
            k_chain_vp_ratio<<<1, 1, 0, b.stream>>>( b.view );
            if( hk )
            {
                k_chain_vp_b<true><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, params.dt, b.partials, T.nblocks );
                reduce_to_slots( S_TQ, 1, true );
                ++launches_;
            }
            else
                k_chain_vp_b<false><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, params.dt, b.partials, T.nblocks );
            launches_ += 4;
            // the new force becomes "the force of the last iteration"
            std::swap( b.view.F, b.view.F2 );
            vp_prev_projected_ = hk;
        }
        else if( solver == Solver_RK4 )
        {
            // configurations: s -> P (s + k1/2) -> P2 (s + k2/2) -> P (s + k3) -> s
            evaluate_force( params, 0, 0 );
            k_chain_rk4<1><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, params.dtg, b.view.S, b.view.P );
            evaluate_force( params, 1, 0 );
            k_chain_rk4<2><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, params.dtg, b.view.P, b.view.P2 );
            evaluate_force( params, 2, 0 );
            k_chain_rk4<3><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, params.dtg, b.view.P2, b.view.P );
            evaluate_force( params, 1, 0 );
            if( b.view.shortening )
            {
                k_chain_force_at<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, b.view.P );
                ++launches_;
            }
            k_chain_rk4<4><<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, params.dtg, b.view.P, b.view.S );
            launches_ += 4;
            if( hk )
            {
                k_chain_hook<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, b.partials, T.nblocks );
                reduce_to_slots( S_TQ, 1, true );
                launches_ += 2;
            }
        }
        else
        {
            for( int stage = 1; stage <= 2; ++stage )
            {
                evaluate_force( params, stage - 1, 0 );
                if( stage == 2 && b.view.shortening )
                {
                    k_chain_force_at<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, b.view.P );
                    ++launches_;
                }
                if( solver == Solver_Depondt )
                    launch_chain_stages<Solver_Depondt>( *this, b, T, p, noi_, params.dtg, stage );
                else if( solver == Solver_Heun )
                    launch_chain_stages<Solver_Heun>( *this, b, T, p, noi_, params.dtg, stage );
                else
                    launch_chain_stages<Solver_SIB>( *this, b, T, p, noi_, params.dtg, stage );
                ++launches_;
            }
            if( hk )
            {
                k_chain_hook<<<grid, BLOCK_THREADS, 0, b.stream>>>( p, T.lg, b.view, b.partials, T.nblocks );
                reduce_to_slots( S_TQ, 1, true );
                launches_ += 2;
            }
        }
    }
    SB_CUDA_CHECK( cudaGetLastError() );
    if( hook )
    {
        SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scal, b.scal, b.n_scal * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
        if( result )
        {
            auto get = [&]( int s, int i ) { return b.h_scal[std::size_t( s ) * noi_global_ + i_begin_ + i]; };
            result->energy.assign( noi_, 0.0 );
            result->Rx.assign( noi_, 0.0 );
            result->max_torque.assign( noi_, 0.0 );
            result->dE_dRx.assign( noi_, 0.0 );
            for( int i = 0; i < noi_; ++i )
            {
                result->energy[i]     = get( S_E, i );
                result->Rx[i]         = get( S_RX, i );
                result->max_torque[i] = std::sqrt( get( S_TQ, i ) );
                const double tt       = get( S_TT, i );
                result->dE_dRx[i]     = tt > 0 ? get( S_FT, i ) / std::sqrt( tt ) : 0.0;
            }
            result->degenerate = b.h_scal[std::size_t( S_N_SLOTS ) * noi_global_ + G_DEGENERATE] != 0.0;
            // convergence is decided on the whole chain, identically on every rank
            result->max_torque_chain = 0;
            for( int i = 0; i < noi_global_; ++i )
                result->max_torque_chain = std::max( result->max_torque_chain, std::sqrt( b.h_scal[std::size_t( S_TQ ) * noi_global_ + i] ) );
        }
    }
}

} // namespace dev
} // namespace sb
