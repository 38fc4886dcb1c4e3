// CUDA kernels of the hot path (sm_100a, fp64, SoA). Included by device/device_image.cu only.
//
// Kernel inventory (one line each; rooflines and byte counts are in DESIGN.md):
//   k_aos_to_soa / k_soa_to_aos      layout conversion at the C-API boundary
//   k_gradient                       gradient (+ energy partials) of one configuration
//   k_energy_contributions           per-spin, per-term energies
//   k_llg_stage<SOLVER,STAGE,..>     fused "gradient -> virtual force -> spin update" solver stages
//   k_vp_a / k_vp_b                  velocity-projection minimiser (two global sums between them)
//   k_hook                           max-torque + tangential projection after an iteration block
//   k_reduce_sum / k_reduce_max      second level of the deterministic two-level reductions
#pragma once

#include "llg.cuh"

namespace sb
{
namespace dev
{

constexpr int BLOCK_THREADS = 256;

// Launch geometry: a CTA covers BX consecutive sites of a row times BY consecutive rows (so the
// +-b neighbour rows are shared inside the CTA through L1), grid is 1-D.
struct LaunchGeom
{
    int bx, by;           // block shape, bx*by = BLOCK_THREADS
    int blocks_x, blocks_y;
    int rowlen, rows;     // Na*NB, Nb*nc_local
};

__device__ __forceinline__ bool locate_site( const StencilParams & p, const LaunchGeom & lg, Site & site, int NB )
{
    const int bxi = blockIdx.x % lg.blocks_x;
    const int byi = blockIdx.x / lg.blocks_x;
    const int tx  = threadIdx.x % lg.bx;
    const int ty  = threadIdx.x / lg.bx;
    const int x   = bxi * lg.bx + tx;
    const int row = byi * lg.by + ty;
    if( x >= lg.rowlen || row >= lg.rows )
        return false;
    site.b = row % p.Nb;
    site.c = row / p.Nb;
    if( NB == 1 )
    {
        site.a  = x;
        site.ib = 0;
    }
    else
    {
        site.a  = x / NB;
        site.ib = x - site.a * NB;
    }
    site.idx = storage_index( p, x, site.b, site.c );
    return true;
}

// ---------------------------------------------------------------------------------------------
// Block-level reductions (deterministic: fixed tree shape)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum( double v )
{
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 )
        v += __shfl_down_sync( 0xffffffffu, v, o );
    return v;
}
__device__ __forceinline__ double warp_max( double v )
{
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 )
        v = fmax( v, __shfl_down_sync( 0xffffffffu, v, o ) );
    return v;
}
// All threads of the CTA must call. Result valid in thread 0.
__device__ __forceinline__ double block_sum( double v )
{
    __shared__ double sh[BLOCK_THREADS / 32];
    v = warp_sum( v );
    __syncthreads(); // protect sh against a previous use
    if( ( threadIdx.x & 31 ) == 0 )
        sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if( threadIdx.x < 32 )
    {
        v = threadIdx.x < BLOCK_THREADS / 32 ? sh[threadIdx.x] : 0.0;
        v = warp_sum( v );
    }
    return v;
}
__device__ __forceinline__ double block_max( double v )
{
    __shared__ double shm[BLOCK_THREADS / 32];
    v = warp_max( v );
    __syncthreads();
    if( ( threadIdx.x & 31 ) == 0 )
        shm[threadIdx.x >> 5] = v;
    __syncthreads();
    if( threadIdx.x < 32 )
    {
        v = threadIdx.x < BLOCK_THREADS / 32 ? shm[threadIdx.x] : 0.0;
        v = warp_max( v );
    }
    return v;
}

// Second level: one CTA folds `n` partials in a fixed order. out[0] = result.
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_reduce_sum( const double * __restrict__ partials, int n, double * __restrict__ out )
{
    double v = 0;
    for( int i = threadIdx.x; i < n; i += BLOCK_THREADS )
        v += partials[i];
    v = block_sum( v );
    if( threadIdx.x == 0 )
        out[0] = v;
}
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_reduce_max( const double * __restrict__ partials, int n, double * __restrict__ out )
{
    double v = 0;
    for( int i = threadIdx.x; i < n; i += BLOCK_THREADS )
        v = fmax( v, partials[i] );
    v = block_max( v );
    if( threadIdx.x == 0 )
        out[0] = v;
}

// Test probe: the normal variates the stage kernels draw, with unit standard deviation, for `count` consecutive Philox counters
// (site in plane = i mod 2^20, plane = i / 2^20) of iteration l.iteration: out[3 i ..] = the three variates of counter i
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_dump_variates( const __grid_constant__ LLGParams l, std::size_t count, float * __restrict__ out )
{
    const std::size_t i = std::size_t( blockIdx.x ) * BLOCK_THREADS + threadIdx.x;
    if( i >= count )
        return;
    unsigned r[4];
    philox4x32_10(
        unsigned( i & 0xfffffu ), unsigned( i >> 20 ), unsigned( l.iteration ), unsigned( l.iteration >> 32 ), unsigned( l.seed ),
        unsigned( l.seed >> 32 ), r );
    const float3 v = scaled_gaussian3f( r[0], r[1], r[2], r[3], -1.3862943611198906f ); // k = -2 ln 2: sigma = 1
    out[3 * i + 0] = v.x;
    out[3 * i + 1] = v.y;
    out[3 * i + 2] = v.z;
}

// The two hook scalars of an iteration block in one launch: out[0] = sum of sums[0..n), out[1] = max of maxs[0..n)
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_reduce_hook( const double * __restrict__ sums, const double * __restrict__ maxs, int n, double * __restrict__ out )
{
    double v = 0, m = 0;
    for( int i = threadIdx.x; i < n; i += BLOCK_THREADS )
    {
        v += sums[i];
        m = fmax( m, maxs[i] );
    }
    v = block_sum( v );
    m = block_max( m );
    if( threadIdx.x == 0 )
    {
        out[0] = v;
        out[1] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// Layout conversion at the C-API boundary: host arrays are AoS [nos][3] in the reference's site
// order; device fields are planar with (optional) halo planes.
// ---------------------------------------------------------------------------------------------
// i: index in the reference's site order (no padding, no halo) -> storage index
__device__ __forceinline__ std::size_t storage_of_linear( int i, int plane_sites, int plane_stride, int halo )
{
    const int c = i / plane_sites;
    return std::size_t( i - c * plane_sites ) + std::size_t( plane_stride ) * ( c + halo );
}
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_aos_to_soa( const double * __restrict__ aos, Field3 f, int n, int plane_sites, int plane_stride, int halo )
{
    const int i = blockIdx.x * BLOCK_THREADS + threadIdx.x;
    if( i < n )
        store3(
            f, storage_of_linear( i, plane_sites, plane_stride, halo ),
            make_d3( aos[3 * std::size_t( i ) + 0], aos[3 * std::size_t( i ) + 1], aos[3 * std::size_t( i ) + 2] ) );
}
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_soa_to_aos( ConstField3 f, double * __restrict__ aos, int n, int plane_sites, int plane_stride, int halo, double scale )
{
    const int i = blockIdx.x * BLOCK_THREADS + threadIdx.x;
    if( i < n )
    {
        const D3 v                    = load3( f, storage_of_linear( i, plane_sites, plane_stride, halo ) );
        aos[3 * std::size_t( i ) + 0] = scale * v.x;
        aos[3 * std::size_t( i ) + 1] = scale * v.y;
        aos[3 * std::size_t( i ) + 2] = scale * v.z;
    }
}

// out = in with the sites without a magnetic moment zeroed (element e of an AoSoA-32 field belongs to storage site
// (e / 96) * 32 + e % 32): operand of the dipolar convolution on a lattice with defects
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_mask_moments( const double * __restrict__ in, double * __restrict__ out, const unsigned char * __restrict__ flags, std::size_t n )
{
    const std::size_t e = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x;
    if( e < n )
    {
        const std::size_t site = ( e / ( 3 * FIELD_BLOCK ) ) * FIELD_BLOCK + ( e % FIELD_BLOCK );
        out[e]                 = ( flags[site] & FLAG_NO_MU_S ) ? 0.0 : in[e];
    }
}

// ---------------------------------------------------------------------------------------------
// Gradient (+ energy) of one configuration. Hamiltonian_Heisenberg.cpp:670-766.
// out = sign * gradient (sign = -1 gives the effective field / force).
// ---------------------------------------------------------------------------------------------
template<int NB_T, bool WITH_ENERGY>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_gradient(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, ConstField3 s, ConstField3 ddi,
    Field3 out, double sign, double * __restrict__ energy_partials )
{
    Site site;
    const bool active = locate_site( p, lg, site, NB_T > 0 ? NB_T : p.NB );
    double e          = 0;
    if( active )
    {
        const D3 si          = load3( s, site.idx );
        const SiteGradient g = site_gradient<NB_T>( p, s, ddi, site, si );
        const D3 gt          = total( g );
        store3( out, site.idx, make_d3( sign * gt.x, sign * gt.y, sign * gt.z ) );
        if( WITH_ENERGY )
            e = site_energy<NB_T>( p, site, si, g );
    }
    if( WITH_ENERGY )
    {
        e = block_sum( e );
        if( threadIdx.x == 0 )
            energy_partials[blockIdx.x] = e;
    }
}

// Per-spin energies of each term (Hamiltonian_Heisenberg.cpp:307-404). `terms` holds up to 6 output
// arrays (null = term inactive), indexed 0 Zeeman, 1 anisotropy, 2 cubic, 3 exchange, 4 DMI, 5 DDI,
// each of length nos in the reference's site order (no halo).
struct EnergyTermPointers
{
    double * term[6];
};
template<int NB_T>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_energy_contributions(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, ConstField3 s, ConstField3 ddi,
    EnergyTermPointers out )
{
    Site site;
    const int NB = NB_T > 0 ? NB_T : p.NB;
    if( !locate_site( p, lg, site, NB ) )
        return;
    const int ib  = NB_T == 1 ? 0 : site.ib;
    const D3 si   = load3( s, site.idx );
    const int lin = site.a * NB + ib + p.Na * NB * ( site.b + p.Nb * site.c ); // index without halo
    const unsigned flags = p.site_flags ? __ldg( p.site_flags + site.idx ) : 0u;
    if( flags & FLAG_VACANT )
    {
        for( int t = 0; t < 6; ++t )
            if( out.term[t] )
                out.term[t][lin] = 0.0;
        return;
    }

    if( out.term[0] )
        out.term[0][lin] = ( flags & FLAG_NO_MU_S ) ? 0.0 : -( p.zeeman[ib][0] * si.x + p.zeeman[ib][1] * si.y + p.zeeman[ib][2] * si.z );
    if( out.term[1] )
    {
        double e = 0;
        for( int i = 0; i < p.n_aniso; ++i )
            if( p.aniso[i].ib == ib )
            {
                const double d = p.aniso[i].nx * si.x + p.aniso[i].ny * si.y + p.aniso[i].nz * si.z;
                e -= p.aniso[i].K * d * d;
            }
        out.term[1][lin] = e;
    }
    if( out.term[2] )
    {
        const double x2 = si.x * si.x, y2 = si.y * si.y, z2 = si.z * si.z;
        out.term[2][lin] = -0.5 * p.K4[ib] * ( x2 * x2 + y2 * y2 + z2 * z2 );
    }
    if( out.term[3] || out.term[4] )
    {
        double e_ex = 0, e_dmi = 0;
        for( int n = p.neigh_begin[ib]; n < p.neigh_begin[ib + 1]; ++n )
        {
            const Neighbour & nb = p.neigh[n];
            int ja = site.a + nb.da, jb = site.b + nb.db, jc = site.c + nb.dc;
            bool valid = true;
            if( ja < 0 ) { ja += p.Na; valid = valid && p.bc[0]; }
            else if( ja >= p.Na ) { ja -= p.Na; valid = valid && p.bc[0]; }
            if( jb < 0 ) { jb += p.Nb; valid = valid && p.bc[1]; }
            else if( jb >= p.Nb ) { jb -= p.Nb; valid = valid && p.bc[1]; }
            if( p.halo == 0 )
            {
                if( jc < 0 ) { jc += p.Nc; valid = valid && p.bc[2]; }
                else if( jc >= p.Nc ) { jc -= p.Nc; valid = valid && p.bc[2]; }
            }
            else
            {
                const int gc = p.c_begin + jc;
                valid        = valid && ( p.bc[2] || ( gc >= 0 && gc < p.Nc ) );
            }
            if( valid && p.site_flags && ( __ldg( p.site_flags + storage_index( p, ja * NB + nb.jb, jb, jc ) ) & FLAG_VACANT ) )
                valid = false;
            if( valid )
            {
                const D3 sj = load3( s, storage_index( p, ja * NB + nb.jb, jb, jc ) );
                e_ex -= 0.5 * nb.J * dot3( si, sj );
                // -1/2 D . (s_i x s_j)
                const D3 c = cross3( si, sj );
                e_dmi -= 0.5 * ( nb.Dx * c.x + nb.Dy * c.y + nb.Dz * c.z );
            }
        }
        if( out.term[3] )
            out.term[3][lin] = e_ex;
        if( out.term[4] )
            out.term[4][lin] = e_dmi;
    }
    if( out.term[5] )
        out.term[5][lin] = ( flags & FLAG_NO_MU_S ) ? 0.0 : 0.5 * dot3( si, load3( ddi, site.idx ) );
}

// ---------------------------------------------------------------------------------------------
// Fused solver stages. One launch = gradient stencil + virtual force + spin update for every site.
//
//   Depondt (Solver_Depondt.hpp:29-77)  stage 1: s' = R(Fv(s)) s          stage 2: s <- R((Fv(s)+Fv(s'))/2) s
//   Heun    (Solver_Heun.hpp:30-81)     stage 1: s' = |s - s x Fv(s)|     stage 2: s <- |s + k1/2 + k2/2|
//   SIB     (Solver_SIB.hpp:22-50)      stage 1: s' = (s + T(s,Fv(s)))/2  stage 2: s <- T(s, Fv(s'))
//   RK4     (Solver_RK4.hpp:41-147)     stages 1-4 with a running accumulator acc = k1/6 + k2/3 + k3/3
//
// Stage 2 of Depondt / Heun RECOMPUTES the stage-1 virtual force from s instead of storing it:
// 120 B of HBM traffic per spin and iteration (R s | W s' | R s, s' | W s_new) instead of 168 B.
// The new configuration goes to a third buffer because neighbours still read the old one.
//
// HOOK variants (last iteration before Method_LLG::Hook_Post_Iteration, Method_LLG.cpp:246-301):
// stage 1 also stores F(s) and Fv(s); the last stage also reduces the energy of the configuration
// of its (last) force evaluation -- the reference's `current_energy`.
// ---------------------------------------------------------------------------------------------
struct StageArgs
{
    ConstField3 s;      // configuration at the start of the iteration
    ConstField3 sp;     // current predictor configuration (stages >= 2)
    ConstField3 ddi_s;  // DDI gradient of s  (if has_ddi)
    ConstField3 ddi_sp; // DDI gradient of sp (if has_ddi)
    Field3 out;         // configuration written by this stage
    Field3 acc;         // RK4 accumulator (read-modify-write, own site only)
    Field3 F_out;       // HOOK, stage 1: force -gradient(s)
    Field3 Fv_out;      // HOOK, stage 1: virtual force Fv(s)
    double * energy_partials; // HOOK, last stage
    float * xi;         // marching kernels at T > 0: the thermal field of the iteration as the fp32 variates it is made of
                        // (AoSoA-32 like the fields): written by stage 1, read by the later stages
};

template<int SOLVER, int STAGE, int NB_T, bool HOOK>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_llg_stage(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ LLGParams l,
    const __grid_constant__ StageArgs a )
{
    using Needs               = StageNeeds<SOLVER, STAGE>;
    constexpr bool last_stage = Needs::last;

    Site site;
    const bool active = locate_site( p, lg, site, NB_T > 0 ? NB_T : p.NB );
    double e          = 0;
    if( active )
    {
        const D3 si = load3( a.s, site.idx );
        D3 xi       = make_d3( 0, 0, 0 );
        if( l.has_thermal && !l.direct_minimization )
            xi = thermal_field<NB_T>( p, l, site );

        D3 Fv = make_d3( 0, 0, 0 ), Fvp = make_d3( 0, 0, 0 ), spi = si;
        if( Needs::Fv_s )
        {
            const SiteGradient g = site_gradient<NB_T>( p, a.s, a.ddi_s, site, si );
            const D3 gt          = total( g );
            const bool frozen    = site_frozen( g );
            const D3 F           = frozen ? make_d3( 0, 0, 0 ) : make_d3( -gt.x, -gt.y, -gt.z );
            Fv                   = virtual_force<NB_T>( l, site, si, F, xi );
            if( l.has_stt == 2 )
                Fv = add3( Fv, stt_gradient_term<NB_T>( p, l, a.s, site, si ) );
            if( frozen )
                Fv = make_d3( 0, 0, 0 );
            if( HOOK && STAGE == 1 )
            {
                store3( a.F_out, site.idx, F );
                store3( a.Fv_out, site.idx, Fv );
            }
        }
        if( Needs::Fv_sp )
        {
            spi                  = load3( a.sp, site.idx );
            const SiteGradient g = site_gradient<NB_T>( p, a.sp, a.ddi_sp, site, spi );
            const D3 gt          = total( g );
            Fvp                  = virtual_force<NB_T>( l, site, spi, make_d3( -gt.x, -gt.y, -gt.z ), xi );
            if( l.has_stt == 2 )
                Fvp = add3( Fvp, stt_gradient_term<NB_T>( p, l, a.sp, site, spi ) );
            if( site_frozen( g ) )
                Fvp = make_d3( 0, 0, 0 );
            if( HOOK && last_stage )
                e = site_energy<NB_T>( p, site, spi, g );
        }

        D3 acc = make_d3( 0, 0, 0 );
        if( SOLVER == Solver_RK4 && STAGE > 1 )
            acc = load3( a.acc, site.idx );
        const D3 out = solver_update<SOLVER, STAGE>( si, Fv, spi, Fvp, acc );
        if( SOLVER == Solver_RK4 && STAGE < 4 )
            store3( a.acc, site.idx, acc );
        store3( a.out, site.idx, out );
    }
    if( HOOK && last_stage )
    {
        e = block_sum( e );
        if( threadIdx.x == 0 )
            a.energy_partials[blockIdx.x] = e;
    }
}

// ---------------------------------------------------------------------------------------------
// Velocity projection (Solver_VP.hpp:29-114), mass m = 1. The projected velocity after an
// iteration is always ratio*F (or 0), so the velocity field is never stored:
//   A:  F = -grad(s);  v = ratio_prev F_prev + (F_prev + F)/2;  partial sums of v.F and F.F;  F_prev <- F
//   (two-level reduce, ratio = proj/|F|^2 or 0)
//   B:  s <- |s + dt ratio F + dt F/2|
// `scal` layout: [0] ratio_prev, [1] sum v.F, [2] sum F.F, [3] ratio of this iteration.
// HOOK: A reduces the energy E(s); B also produces max |Fv - (Fv.s_new)s_new| with
// Fv = dtg' s x F, and writes F projected tangentially to the NEW spins into a second buffer -- the
// reference's hook projects `forces` in place and the projected force is what the next iteration
// sees as F_prev, while its stored velocity still derives from the raw force (SURVEY.md 8c hazard 6).
// ---------------------------------------------------------------------------------------------
template<int NB_T, bool HOOK>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_vp_a(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, ConstField3 s, ConstField3 ddi,
    Field3 F, ConstField3 F_prev, const double * __restrict__ scal, double * __restrict__ part_proj,
    double * __restrict__ part_norm2, double * __restrict__ energy_partials )
{
    Site site;
    const bool active = locate_site( p, lg, site, NB_T > 0 ? NB_T : p.NB );
    double proj = 0, norm2 = 0, e = 0;
    if( active )
    {
        const double ratio_prev = scal[0];
        const D3 si             = load3( s, site.idx );
        const SiteGradient g    = site_gradient<NB_T>( p, s, ddi, site, si );
        const D3 gt             = total( g );
        const D3 Fn             = site_frozen( g ) ? make_d3( 0, 0, 0 ) : make_d3( -gt.x, -gt.y, -gt.z );
        // velocity of the last iteration = ratio_prev * (raw force of the last iteration); F_prev is the same force,
        // or its tangential projection if a post-iteration hook ran in between (SURVEY.md 8c hazard 6)
        const D3 Fr             = load3( F, site.idx );
        const D3 Fp             = load3( F_prev, site.idx );
        const D3 v              = make_d3(
            ratio_prev * Fr.x + 0.5 * ( Fp.x + Fn.x ), ratio_prev * Fr.y + 0.5 * ( Fp.y + Fn.y ),
            ratio_prev * Fr.z + 0.5 * ( Fp.z + Fn.z ) );
        proj  = dot3( v, Fn );
        norm2 = dot3( Fn, Fn );
        store3( F, site.idx, Fn );
        if( HOOK )
            e = site_energy<NB_T>( p, site, si, g );
    }
    proj = block_sum( proj );
    if( threadIdx.x == 0 )
        part_proj[blockIdx.x] = proj;
    norm2 = block_sum( norm2 );
    if( threadIdx.x == 0 )
        part_norm2[blockIdx.x] = norm2;
    if( HOOK )
    {
        e = block_sum( e );
        if( threadIdx.x == 0 )
            energy_partials[blockIdx.x] = e;
    }
}

// scal[3] = ratio (0 if the projection is not positive); also becomes ratio_prev of the next iteration
static __global__ void k_vp_ratio( double * __restrict__ scal )
{
    const double proj = scal[1], norm2 = scal[2];
    const double ratio = proj > 0 ? proj / norm2 : 0.0;
    scal[3]            = ratio;
    scal[0]            = ratio;
}

template<bool HOOK>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_vp_b(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, Field3 s, ConstField3 F, Field3 F_projected,
    const double * __restrict__ scal, double dt, double dtg, double * __restrict__ torque_partials )
{
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double t2         = 0;
    if( active )
    {
        const double ratio = scal[3];
        const D3 si        = load3( s, site.idx );
        D3 Fi              = load3( F, site.idx );
        const double c     = dt * ratio + 0.5 * dt;
        const D3 sn        = normalized3( make_d3( si.x + c * Fi.x, si.y + c * Fi.y, si.z + c * Fi.z ) );
        store3( s, site.idx, sn );
        if( HOOK )
        {
            const D3 sxF = cross3( si, Fi );
            D3 Fv        = make_d3( dtg * sxF.x, dtg * sxF.y, dtg * sxF.z );
            const double d = dot3( Fv, sn );
            Fv             = make_d3( Fv.x - d * sn.x, Fv.y - d * sn.y, Fv.z - d * sn.z );
            t2             = dot3( Fv, Fv );
            const double f = dot3( Fi, sn );
            Fi             = make_d3( Fi.x - f * sn.x, Fi.y - f * sn.y, Fi.z - f * sn.z );
            store3( F_projected, site.idx, Fi );
        }
    }
    if( HOOK )
    {
        t2 = block_max( t2 );
        if( threadIdx.x == 0 )
            torque_partials[blockIdx.x] = t2;
    }
}

// ---------------------------------------------------------------------------------------------
// Post-iteration hook (Method_LLG.cpp:246-301, Method_Solver.hpp:224-230):
//   max torque = max_i |Fv_i - (Fv_i.s_i) s_i|  (Fv: stage-1 virtual force, s: the NEW spins)
//   effective field = F - (F.s) s               (F: stage-1 force)  -- in place
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_hook(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, ConstField3 s, Field3 F, ConstField3 Fv,
    double * __restrict__ torque_partials )
{
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    double t2         = 0;
    if( active )
    {
        const D3 si    = load3( s, site.idx );
        const D3 fv    = load3( Fv, site.idx );
        const double d = dot3( fv, si );
        const D3 tq    = make_d3( fv.x - d * si.x, fv.y - d * si.y, fv.z - d * si.z );
        t2             = dot3( tq, tq );
        const D3 Fi    = load3( F, site.idx );
        const double f = dot3( Fi, si );
        store3( F, site.idx, make_d3( Fi.x - f * si.x, Fi.y - f * si.y, Fi.z - f * si.z ) );
    }
    t2 = block_max( t2 );
    if( threadIdx.x == 0 )
        torque_partials[blockIdx.x] = t2;
}

// Force, virtual force and energy of the current spins (constructor-time evaluation, Method_LLG.cpp:57-62)
template<int NB_T>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_force_and_virtual(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, const __grid_constant__ LLGParams l,
    ConstField3 s, ConstField3 ddi, Field3 F_out, Field3 Fv_out, double * __restrict__ energy_partials )
{
    Site site;
    const bool active = locate_site( p, lg, site, NB_T > 0 ? NB_T : p.NB );
    double e          = 0;
    if( active )
    {
        const D3 si = load3( s, site.idx );
        D3 xi       = make_d3( 0, 0, 0 );
        if( l.has_thermal && !l.direct_minimization )
            xi = thermal_field<NB_T>( p, l, site );
        const SiteGradient g = site_gradient<NB_T>( p, s, ddi, site, si );
        const D3 gt          = total( g );
        const bool frozen    = site_frozen( g );
        const D3 F           = frozen ? make_d3( 0, 0, 0 ) : make_d3( -gt.x, -gt.y, -gt.z );
        store3( F_out, site.idx, F );
        D3 Fv = virtual_force<NB_T>( l, site, si, F, xi );
        if( l.has_stt == 2 )
            Fv = add3( Fv, stt_gradient_term<NB_T>( p, l, s, site, si ) );
        if( frozen )
            Fv = make_d3( 0, 0, 0 );
        store3( Fv_out, site.idx, Fv );
        e = site_energy<NB_T>( p, site, si, g );
    }
    e = block_sum( e );
    if( threadIdx.x == 0 )
        energy_partials[blockIdx.x] = e;
}

// sum_i mu_s[ib] * s_i per component (Magnetization, Vectormath.cpp:495-502): partials [3][nblocks]
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_magnetization(
    const __grid_constant__ StencilParams p, const __grid_constant__ LaunchGeom lg, ConstField3 s, double * __restrict__ partials,
    int nblocks, int weighted )
{
    Site site;
    const bool active = locate_site( p, lg, site, p.NB );
    D3 m              = make_d3( 0, 0, 0 );
    if( active )
    {
        const D3 si     = load3( s, site.idx );
        double mu       = weighted ? p.mu_s[site.ib] : 1.0;
        if( weighted && p.site_flags && ( __ldg( p.site_flags + site.idx ) & FLAG_NO_MU_S ) )
            mu = 0.0;
        m               = make_d3( mu * si.x, mu * si.y, mu * si.z );
    }
    double v = block_sum( m.x );
    if( threadIdx.x == 0 )
        partials[blockIdx.x] = v;
    v = block_sum( m.y );
    if( threadIdx.x == 0 )
        partials[nblocks + blockIdx.x] = v;
    v = block_sum( m.z );
    if( threadIdx.x == 0 )
        partials[2 * nblocks + blockIdx.x] = v;
}

// Topological charge of a planar lattice with one basis atom (TopologicalChargeDensity, Vectormath.cpp:504-631): every cell
// (a, b) carries the two triangles of its parallelogram (0, a, b, a+b), cut along the diagonal the Delaunay triangulation of
// the reference picks (diag 0: a-b, 1: 0-(a+b)); a triangle counts when its translations are inside the lattice or allowed
// by the boundary conditions. Charge of a triangle: sign / (4 pi) * 2 atan2( s1.(s2 x s3), 1 + s1.s2 + s1.s3 + s2.s3 ).
// density (nullable): [2][Na*Nb], 0 for triangles that do not count; partials [nblocks].
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_topological_charge(
    const __grid_constant__ StencilParams p, ConstField3 s, int diag, double sign0, double sign1, double * __restrict__ density,
    double * __restrict__ partials )
{
    const int cell  = blockIdx.x * blockDim.x + threadIdx.x;
    const int cells = p.Na * p.Nb;
    double q0 = 0, q1 = 0;
    if( cell < cells )
    {
        const int a = cell % p.Na, b = cell / p.Na;
        const bool allowed = ( a + 1 < p.Na || p.bc[0] ) && ( b + 1 < p.Nb || p.bc[1] );
        if( allowed )
        {
            const std::size_t plane = std::size_t( p.plane_stride ) * p.halo;
            const int an = ( a + 1 ) % p.Na, bn = ( b + 1 ) % p.Nb;
            const D3 s0 = load3( s, plane + a + std::size_t( p.Na ) * b ), sa = load3( s, plane + an + std::size_t( p.Na ) * b );
            const D3 sb = load3( s, plane + a + std::size_t( p.Na ) * bn ), sab = load3( s, plane + an + std::size_t( p.Na ) * bn );
            auto solid = []( const D3 & v1, const D3 & v2, const D3 & v3 )
            {
                const double x = dot3( v1, cross3( v2, v3 ) );
                const double y = 1 + dot3( v1, v2 ) + dot3( v1, v3 ) + dot3( v2, v3 );
                return 2 * atan2( x, y );
            };
            const double f = 1.0 / ( 4.0 * 3.14159265358979323846 );
            if( diag == 0 )
            {
                q0 = sign0 * f * solid( sa, sb, sab );
                q1 = sign1 * f * solid( sa, sb, s0 );
            }
            else
            {
                q0 = sign0 * f * solid( s0, sa, sab );
                q1 = sign1 * f * solid( s0, sab, sb );
            }
        }
        if( density )
        {
            density[cell]         = q0;
            density[cells + cell] = q1;
        }
    }
    const double v = block_sum( q0 + q1 );
    if( threadIdx.x == 0 )
        partials[blockIdx.x] = v;
}

// The same for any number of basis atoms: the triangles of one cell come as a table (Delaunay triangulation of the basis
// atoms and the three neighbouring corners of the cell, made on the host as in Vectormath.cpp:516-575). Vertex v < NB is
// basis atom v of the cell (a, b); NB + 2 / NB + 1 / NB are atom 0 of the cells (a+1, b) / (a, b+1) / (a+1, b+1), which
// count only when that translation stays inside the lattice or is allowed by the boundary conditions.
struct TopologyTable
{
    int n;
    int vertex[16][3];
    double sign[16];
};
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_topological_charge_table(
    const __grid_constant__ StencilParams p, ConstField3 s, const __grid_constant__ TopologyTable t, double * __restrict__ density,
    double * __restrict__ partials )
{
    const int cell  = blockIdx.x * blockDim.x + threadIdx.x;
    const int cells = p.Na * p.Nb;
    double sum      = 0;
    if( cell < cells )
    {
        const int a = cell % p.Na, b = cell / p.Na;
        const bool a_ok = a + 1 < p.Na || p.bc[0], b_ok = b + 1 < p.Nb || p.bc[1];
        const int an = ( a + 1 ) % p.Na, bn = ( b + 1 ) % p.Nb;
        const std::size_t plane = std::size_t( p.plane_stride ) * p.halo;
        for( int k = 0; k < t.n; ++k )
        {
            bool valid = true;
            D3 v[3];
            for( int c = 0; c < 3; ++c )
            {
                const int id = t.vertex[k][c];
                std::size_t site;
                if( id < p.NB )
                    site = std::size_t( id ) + std::size_t( p.NB ) * ( a + std::size_t( p.Na ) * b );
                else if( id == p.NB + 2 )
                {
                    valid = valid && a_ok;
                    site  = std::size_t( p.NB ) * ( an + std::size_t( p.Na ) * b );
                }
                else if( id == p.NB + 1 )
                {
                    valid = valid && b_ok;
                    site  = std::size_t( p.NB ) * ( a + std::size_t( p.Na ) * bn );
                }
                else
                {
                    valid = valid && a_ok && b_ok;
                    site  = std::size_t( p.NB ) * ( an + std::size_t( p.Na ) * bn );
                }
                v[c] = load3( s, plane + site );
            }
            double q = 0;
            if( valid )
            {
                const double x = dot3( v[0], cross3( v[1], v[2] ) );
                const double y = 1 + dot3( v[0], v[1] ) + dot3( v[0], v[2] ) + dot3( v[1], v[2] );
                q              = t.sign[k] * ( 1.0 / ( 4.0 * 3.14159265358979323846 ) ) * 2 * atan2( x, y );
            }
            if( density )
                density[std::size_t( k ) * cells + cell] = q;
            sum += q;
        }
    }
    const double total = block_sum( sum );
    if( threadIdx.x == 0 )
        partials[blockIdx.x] = total;
}

} // namespace dev
} // namespace sb
