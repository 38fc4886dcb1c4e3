// Specialised solver-stage kernels for the nearest-neighbour ("7-point") stencil: one basis atom, neighbours only
// at +-a, +-b, +-c (StencilParams::sc6). This is the structure of every BASELINE.json configuration (simple cubic,
// first-shell exchange + DMI) and the one the roofline target is quoted on; any other pair list runs through the
// generic gather kernels of kernels.cuh.
//
// Mapping: a CTA owns BX consecutive sites of BY consecutive rows and MARCHES along c over a segment of `lc`
// planes. Each thread keeps its own column (c-1, c, c+1) in registers, so per plane it reads one new own-column
// value (the only access that has to come from HBM) and the four in-plane neighbours, which are the own-column
// values of neighbouring threads of the same plane step and hit L1 (+-a, +-b inside the CTA) or L2 (the two
// halo rows above / below the CTA's rows). All index arithmetic, boundary handling and the Philox round keys
// are hoisted out of the plane loop. L2->SM traffic per field is (BY+2)/BY x 24 B per site instead of 5-7 x.
//
// HBM traffic per site: stage 1 reads s (24 B) and writes s' (24 B); stage 2 of Depondt/Heun reads s and s'
// and writes the new configuration (72 B), recomputing the stage-1 virtual force instead of storing it.
#pragma once

#include "llg.cuh"

namespace sb
{
namespace dev
{

constexpr int SC6_MAX_THREADS = 512;

// load or zero: open boundaries / absent axes contribute a zero spin
__device__ __forceinline__ D3 ld3v( const ConstField3 & f, std::size_t idx, bool valid )
{
    D3 r = make_d3( 0.0, 0.0, 0.0 );
    if( valid )
        r = make_d3( __ldg( f.x + idx ), __ldg( f.y + idx ), __ldg( f.z + idx ) );
    return r;
}

// Contribution of the neighbour pair (minus, plus) along one axis:
//   g -= J (s+ + s-) + (s+ - s-) x D        [ D(+) = D, D(-) = -D ]
// which is Gradient_Exchange + Gradient_DMI (Hamiltonian_Heisenberg.cpp:822-864) for the two redundant pairs.
template<int AXIS>
__device__ __forceinline__ void sc6_axis_gradient( const StencilParams & p, const D3 & m, const D3 & pl, D3 & g )
{
    if( !p.sc6_axis[AXIS] )
        return;
    const double J = p.sc6_J[AXIS];
    g.x            = fma( -J, m.x + pl.x, g.x );
    g.y            = fma( -J, m.y + pl.y, g.y );
    g.z            = fma( -J, m.z + pl.z, g.z );
    const int fl   = p.sc6_dflags[AXIS];
    if( fl )
    {
        const D3 d = make_d3( pl.x - m.x, pl.y - m.y, pl.z - m.z );
        // d x D = (d.y Dz - d.z Dy, d.z Dx - d.x Dz, d.x Dy - d.y Dx), zero components of D skipped
        if( fl & 1 )
        {
            const double D = p.sc6_D[AXIS][0];
            g.y            = fma( -D, d.z, g.y );
            g.z            = fma( D, d.y, g.z );
        }
        if( fl & 2 )
        {
            const double D = p.sc6_D[AXIS][1];
            g.x            = fma( D, d.z, g.x );
            g.z            = fma( -D, d.x, g.z );
        }
        if( fl & 4 )
        {
            const double D = p.sc6_D[AXIS][2];
            g.x            = fma( -D, d.y, g.x );
            g.y            = fma( D, d.x, g.y );
        }
    }
}

// Full site gradient from the six neighbour spins (same split as site_gradient of stencil.cuh)
__device__ __forceinline__ SiteGradient sc6_site_gradient(
    const StencilParams & p, const D3 & si, const D3 & xm, const D3 & xp, const D3 & bm, const D3 & bp, const D3 & cm,
    const D3 & cp, const ConstField3 & ddi, std::size_t idx )
{
    SiteGradient out;
    D3 g = make_d3( 0.0, 0.0, 0.0 );
    sc6_axis_gradient<0>( p, xm, xp, g );
    sc6_axis_gradient<1>( p, bm, bp, g );
    sc6_axis_gradient<2>( p, cm, cp, g );
    // Uniaxial anisotropy: g -= 2 K (n.s) n   (Hamiltonian_Heisenberg.cpp:785-800)
    for( int i = 0; i < p.n_aniso; ++i )
    {
        const Anisotropy & an = p.aniso[i];
        if( an.ib != 0 )
            continue;
        double d = 0.0;
        if( an.flags & 1 )
            d = an.nx * si.x;
        if( an.flags & 2 )
            d = fma( an.ny, si.y, d );
        if( an.flags & 4 )
            d = fma( an.nz, si.z, d );
        d *= -2.0 * an.K;
        if( an.flags & 1 )
            g.x = fma( d, an.nx, g.x );
        if( an.flags & 2 )
            g.y = fma( d, an.ny, g.y );
        if( an.flags & 4 )
            g.z = fma( d, an.nz, g.z );
    }
    if( p.has_ddi )
    {
        g.x += __ldg( ddi.x + idx );
        g.y += __ldg( ddi.y + idx );
        g.z += __ldg( ddi.z + idx );
    }
    out.bilinear = g;
    D3 r         = make_d3( 0.0, 0.0, 0.0 );
    if( p.has_cubic )
    {
        const double k = 2.0 * p.K4[0];
        r.x -= k * si.x * si.x * si.x;
        r.y -= k * si.y * si.y * si.y;
        r.z -= k * si.z * si.z * si.z;
    }
    if( p.has_zeeman )
    {
        r.x -= p.zeeman[0][0];
        r.y -= p.zeeman[0][1];
        r.z -= p.zeeman[0][2];
    }
    out.rest = r;
    return out;
}

// In-plane neighbour offsets of a thread's column and the plane bookkeeping of the march
struct SC6Column
{
    int oc, oxm, oxp, obm, obp; // offsets inside a plane
    bool vxm, vxp, vbm, vbp;
    std::size_t plane_stride;
};

__device__ __forceinline__ SC6Column sc6_column( const StencilParams & p, int x, int b )
{
    SC6Column col;
    int xm = x - 1, xp = x + 1, bm = b - 1, bp = b + 1;
    col.vxm = col.vxp = p.sc6_axis[0] != 0;
    col.vbm = col.vbp = p.sc6_axis[1] != 0;
    if( xm < 0 )
    {
        xm += p.Na;
        col.vxm = col.vxm && p.bc[0];
    }
    if( xp >= p.Na )
    {
        xp -= p.Na;
        col.vxp = col.vxp && p.bc[0];
    }
    if( bm < 0 )
    {
        bm += p.Nb;
        col.vbm = col.vbm && p.bc[1];
    }
    if( bp >= p.Nb )
    {
        bp -= p.Nb;
        col.vbp = col.vbp && p.bc[1];
    }
    const int row    = p.Na * b;
    col.oc           = row + x;
    col.oxm          = row + xm;
    col.oxp          = row + xp;
    col.obm          = p.Na * bm + x;
    col.obp          = p.Na * bp + x;
    col.plane_stride = std::size_t( p.Na ) * p.Nb;
    return col;
}

// Storage offset of the plane that holds the c-neighbour `cc` (= c-1 or c+1, local index). The plane always exists in
// storage (periodic wrap on one device, halo planes on a slab); whether it CONTRIBUTES is sc6_c_valid.
__device__ __forceinline__ std::size_t sc6_c_plane( const StencilParams & p, const SC6Column & col, int cc )
{
    if( p.halo == 0 )
    {
        if( cc < 0 )
            cc += p.Nc;
        else if( cc >= p.Nc )
            cc -= p.Nc;
        return std::size_t( cc ) * col.plane_stride;
    }
    return std::size_t( cc + p.halo ) * col.plane_stride;
}
__device__ __forceinline__ bool sc6_c_valid( const StencilParams & p, int cc )
{
    const int gc = p.c_begin + cc;
    return p.sc6_axis[2] != 0 && ( p.bc[2] || ( gc >= 0 && gc < p.Nc ) );
}

// Virtual force of the configuration `f` at the thread's site of the current plane. (below, center, above) is the
// thread's column of `f`; the in-plane neighbours are gathered here.
__device__ __forceinline__ D3 sc6_virtual_force(
    const StencilParams & p, const LLGParams & l, const SC6Column & col, const ConstField3 & f, const ConstField3 & ddi,
    std::size_t base, const D3 & below_raw, bool vb, const D3 & center, const D3 & above_raw, bool va, const D3 & xi )
{
    const D3 zero        = make_d3( 0.0, 0.0, 0.0 );
    const D3 below       = vb ? below_raw : zero;
    const D3 above       = va ? above_raw : zero;
    const D3 xm          = ld3v( f, base + col.oxm, col.vxm );
    const D3 xp          = ld3v( f, base + col.oxp, col.vxp );
    const D3 bm          = ld3v( f, base + col.obm, col.vbm );
    const D3 bp          = ld3v( f, base + col.obp, col.vbp );
    const SiteGradient g = sc6_site_gradient( p, center, xm, xp, bm, bp, below, above, ddi, base + col.oc );
    const D3 gt          = total( g );
    return virtual_force_ib( l, 0, center, make_d3( -gt.x, -gt.y, -gt.z ), xi );
}

template<int SOLVER, int STAGE>
static __global__ void __launch_bounds__( SC6_MAX_THREADS ) k_sc6_stage(
    const __grid_constant__ StencilParams p, const int lc, const __grid_constant__ LLGParams l,
    const __grid_constant__ StageArgs a )
{
    using Needs = StageNeeds<SOLVER, STAGE>;

    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y * blockDim.y + threadIdx.y;
    if( x >= p.Na || b >= p.Nb )
        return;
    const int c0 = blockIdx.z * lc;
    const int c1 = min( c0 + lc, p.nc_local );

    const SC6Column col = sc6_column( p, x, b );
    const bool thermal  = l.has_thermal && !l.direct_minimization;
    // global index of the site (x, b, c_begin + c0) in the reference's order: Philox counter
    std::uint64_t gsite = std::uint64_t( col.oc ) + col.plane_stride * std::uint64_t( p.c_begin + c0 );

    // the thread's columns: s (needed with neighbours only if this stage evaluates Fv(s)) and the predictor
    D3 s_below = make_d3( 0, 0, 0 ), s_center, s_above = make_d3( 0, 0, 0 );
    D3 p_below = make_d3( 0, 0, 0 ), p_center = make_d3( 0, 0, 0 ), p_above = make_d3( 0, 0, 0 );
    {
        const std::size_t pb   = sc6_c_plane( p, col, c0 - 1 );
        const std::size_t base = std::size_t( c0 + p.halo ) * col.plane_stride;
        s_center               = ld3v( a.s, base + col.oc, true );
        if( Needs::Fv_s )
            s_below = ld3v( a.s, pb + col.oc, true );
        if( Needs::Fv_sp )
        {
            p_center = ld3v( a.sp, base + col.oc, true );
            p_below  = ld3v( a.sp, pb + col.oc, true );
        }
    }

    for( int c = c0; c < c1; ++c )
    {
        const std::size_t base = std::size_t( c + p.halo ) * col.plane_stride;
        const std::size_t pa   = sc6_c_plane( p, col, c + 1 );
        const bool vb = sc6_c_valid( p, c - 1 ), va = sc6_c_valid( p, c + 1 );
        // own column, next plane. Without neighbours of s (SIB stage 2, RK4 stages 2-4) only the centre is needed.
        if( Needs::Fv_s || c + 1 < c1 )
            s_above = ld3v( a.s, pa + col.oc, true );
        if( Needs::Fv_sp )
            p_above = ld3v( a.sp, pa + col.oc, true );

        D3 xi = make_d3( 0, 0, 0 );
        if( thermal )
            xi = thermal_field_at( l, gsite, 0 );

        D3 Fv = make_d3( 0, 0, 0 ), Fvp = make_d3( 0, 0, 0 );
        if( Needs::Fv_s )
            Fv = sc6_virtual_force( p, l, col, a.s, a.ddi_s, base, s_below, vb, s_center, s_above, va, xi );
        if( Needs::Fv_sp )
            Fvp = sc6_virtual_force( p, l, col, a.sp, a.ddi_sp, base, p_below, vb, p_center, p_above, va, xi );

        D3 acc = make_d3( 0, 0, 0 );
        if( SOLVER == Solver_RK4 && STAGE > 1 )
            acc = make_d3( a.acc.x[base + col.oc], a.acc.y[base + col.oc], a.acc.z[base + col.oc] );
        const D3 out = solver_update<SOLVER, STAGE>( s_center, Fv, p_center, Fvp, acc );
        if( SOLVER == Solver_RK4 && STAGE < 4 )
        {
            a.acc.x[base + col.oc] = acc.x;
            a.acc.y[base + col.oc] = acc.y;
            a.acc.z[base + col.oc] = acc.z;
        }
        a.out.x[base + col.oc] = out.x;
        a.out.y[base + col.oc] = out.y;
        a.out.z[base + col.oc] = out.z;

        // march
        s_below  = s_center;
        s_center = s_above;
        p_below  = p_center;
        p_center = p_above;
        gsite += col.plane_stride;
    }
}

} // namespace dev
} // namespace sb
