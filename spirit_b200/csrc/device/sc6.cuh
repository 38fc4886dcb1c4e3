// Specialised solver-stage kernels for the nearest-neighbour ("7-point") stencil: one basis atom, neighbours only
// at +-a, +-b, +-c (StencilParams::sc6). This is the structure of every BASELINE.json configuration (simple cubic,
// first-shell exchange + DMI) and the one the roofline target is quoted on; any other pair list runs through the
// generic gather kernels of kernels.cuh.
//
// Mapping: a CTA owns BX consecutive sites of BY consecutive rows and MARCHES along c over a segment of `lc`
// planes. Each thread keeps its own column (c-1, c, c+1) in registers, so per plane it reads one new own-column
// value (the only access that has to come from HBM) and the four in-plane neighbours, which are the own-column
// values of neighbouring threads of the same plane step and hit L1 (+-a, +-b inside the CTA) or L2 (the two
// halo rows above / below the CTA's rows). L2->SM traffic per field is (BY+2)/BY x 24 B per site instead of 5-7 x.
//
// The step sits near the balance point of B200's fp64 pipe (16 lanes per SM sub-partition: one DFMA warp
// instruction per 2 cycles) and HBM, so the plane loop is written for a minimal instruction count:
//   * all index arithmetic and boundary handling is hoisted out of the loop; inside it a neighbour costs one
//     64-bit address (plane pointer + 32-bit element offset) and three loads at immediate offsets (AoSoA-32);
//   * the Hamiltonian's structure is a template parameter (SPEC): no flag tests or constant reloads per plane;
//   * CTAs that touch no open boundary run a loop without predicated loads or zero-selects (BOUNDARY = false);
//   * signs and prefactors are folded into host-computed constants, the Philox round keys come from the host.
//
// HBM traffic per site: stage 1 reads s (24 B) and writes s' (24 B); stage 2 of Depondt/Heun reads s and s'
// and writes the new configuration (72 B), recomputing the stage-1 virtual force instead of storing it.
#pragma once

#include "kernels.cuh"

namespace sb
{
namespace dev
{

// Launch shapes. A stage that evaluates the gradient of ONE configuration (every stage 1, SIB stage 2, RK4) holds one
// register window and compiles to <= 80 registers: 256-thread CTAs, three per SM (24 warps). Stage 2 of Depondt and
// Heun holds two windows (s and s'), needs ~128 registers and runs as one 512-thread CTA per SM (16 warps).
// Measured on B200 (256^3, profiles/): stage 1 0.228 ms at 16 warps/SM -> 0.203 ms at 24 warps/SM.
// (SB_SC6_* macros: tuning builds, spirit_b200/build.py SPIRIT_B200_DEFINES)
#ifndef SB_SC6_THREADS_1W
#define SB_SC6_THREADS_1W 256
#define SB_SC6_MINB_1W 3
#endif
#ifndef SB_SC6_STORE_XI
#define SB_SC6_STORE_XI 0 // 1: at T > 0 stage 1 stores the fp32 noise variates and the later stages read them instead of
                          // regenerating them. Measured (profiles/r1i): 77 fewer instructions in stage 2 and NO gain (0.2977 vs
                          // 0.2926 ms; stage 1 +6 %): the stages are not bound by their instruction count
#endif
#ifndef SB_SC6_L2_PREFETCH
#define SB_SC6_L2_PREFETCH 0 // > 0: prefetch.global.L2 of the own-column sites this many planes ahead (no registers)
#endif
// Tuning builds only (profiles/r1zz_ptxas_study_plain_variants.txt): compile the rare terms' uniform tests out of the plane loop.
// A library built with any of these set serves only Hamiltonians without that term (no run-time check).
#ifndef SB_SC6_NO_ANISO_FULL
#define SB_SC6_NO_ANISO_FULL 0
#endif
#ifndef SB_SC6_NO_EXTRAS
#define SB_SC6_NO_EXTRAS 0
#endif
#ifndef SB_SC6_NO_STT
#define SB_SC6_NO_STT 0
#endif
#ifndef SB_SC6_THREADS_2W
#define SB_SC6_THREADS_2W 512
#define SB_SC6_MINB_2W 1
#endif
constexpr int SC6_THREADS_1W = SB_SC6_THREADS_1W, SC6_MINB_1W = SB_SC6_MINB_1W;
constexpr int SC6_THREADS_2W = SB_SC6_THREADS_2W, SC6_MINB_2W = SB_SC6_MINB_2W;
template<int SOLVER, int STAGE>
struct SC6Shape
{
    static constexpr bool two_windows = StageNeeds<SOLVER, STAGE>::Fv_s && StageNeeds<SOLVER, STAGE>::Fv_sp;
    static constexpr int threads      = two_windows ? SC6_THREADS_2W : SC6_THREADS_1W;
    static constexpr int min_blocks   = two_windows ? SC6_MINB_2W : SC6_MINB_1W;
};

// SPEC bits
constexpr int SC6_HAS_C       = 1; // neighbours along c exist (3-D system)
constexpr int SC6_DMI_GENERAL = 2; // DMI vectors are not parallel to their bonds (else: axis a -> Dx, b -> Dy, c -> Dz only)
constexpr int SC6_N_SPECS     = 4;
// MODE: what the virtual force is made of (Method_LLG.cpp:131-226)
constexpr int SC6_DYNAMICS = 0; // Fv = dtg/mu_s (F + alpha s x F) [+ STT]
constexpr int SC6_THERMAL  = 1; // ... + xi + alpha s x xi
constexpr int SC6_MINIMISE = 2; // Fv = dtg' s x F (direct minimisation)
constexpr int SC6_N_MODES  = 3;

struct SC6Geometry
{
    dim3 grid, block;
    int lc          = 1; // planes per CTA (march length)
    int seg_first   = 0; // CTA z index -> c-segment: seg_first + blockIdx.z * seg_stride (lets a launch cover only the
    int seg_stride  = 1; // two segments at the slab ends, or only the interior ones: halo-exchange overlap)
};
struct SC6Launch
{
    SC6Geometry one_window, two_windows; // see SC6Shape
    int spec = 0;                        // SC6_HAS_C | SC6_DMI_GENERAL
};

__device__ __forceinline__ D3 ld3p( const double * __restrict__ plane, unsigned e )
{
    const double * q = plane + e;
    return make_d3( __ldg( q ), __ldg( q + FIELD_BLOCK ), __ldg( q + 2 * FIELD_BLOCK ) );
}
// load or zero: open boundaries contribute a zero spin
__device__ __forceinline__ D3 ld3pv( const double * __restrict__ plane, unsigned e, bool valid )
{
    D3 r = make_d3( 0.0, 0.0, 0.0 );
    if( valid )
        r = ld3p( plane, e );
    return r;
}

__device__ __forceinline__ void sc6_prefetch3( const double * q )
{
    asm volatile( "prefetch.global.L2 [%0];" ::"l"( q ) );
    asm volatile( "prefetch.global.L2 [%0];" ::"l"( q + FIELD_BLOCK ) );
    asm volatile( "prefetch.global.L2 [%0];" ::"l"( q + 2 * FIELD_BLOCK ) );
}

// Contribution of the neighbour pair (minus, plus) along one axis:
//   g -= J (s+ + s-) + (s+ - s-) x D        [ D(+) = D, D(-) = -D ]
// which is Gradient_Exchange + Gradient_DMI (Hamiltonian_Heisenberg.cpp:822-864) for the two redundant pairs.
template<int AXIS, bool DMI_GENERAL>
__device__ __forceinline__ void sc6_axis_gradient( const StencilParams & p, const D3 & m, const D3 & pl, D3 & g )
{
    const double nJ = p.sc6_nJ[AXIS]; // -J
    g.x             = fma( nJ, m.x + pl.x, g.x );
    g.y             = fma( nJ, m.y + pl.y, g.y );
    g.z             = fma( nJ, m.z + pl.z, g.z );
    const D3 d      = make_d3( pl.x - m.x, pl.y - m.y, pl.z - m.z );
    // g -= d x D,  d x D = (d.y Dz - d.z Dy, d.z Dx - d.x Dz, d.x Dy - d.y Dx)
    if( DMI_GENERAL || AXIS == 0 )
    {
        const double D = p.sc6_D[AXIS][0];
        g.y            = fma( -D, d.z, g.y );
        g.z            = fma( D, d.y, g.z );
    }
    if( DMI_GENERAL || AXIS == 1 )
    {
        const double D = p.sc6_D[AXIS][1];
        g.x            = fma( D, d.z, g.x );
        g.z            = fma( -D, d.x, g.z );
    }
    if( DMI_GENERAL || AXIS == 2 )
    {
        const double D = p.sc6_D[AXIS][2];
        g.x            = fma( -D, d.y, g.x );
        g.y            = fma( D, d.x, g.y );
    }
}

// Gradient of all terms at one site from its six neighbour spins.
//   start value: -mu_s B n (Zeeman, Hamiltonian_Heisenberg.cpp:768-783; zero without a field)
//   on-site quadratic form: g += A s with A = -2 sum_k K_k n_k n_k^T (uniaxial anisotropies, :785-800)
//   rare terms behind one uniform flag: cubic anisotropy (:802-820), precomputed dipolar field
// RARE: how the rare terms (off-diagonal anisotropy elements, cubic anisotropy, dipolar field) are reached.
//   0  two uniform tests (two-pass kernels)
//   1  nested: the off-diagonal anisotropy elements sit behind the same flag as the others (sc6_extras includes
//      sc6_aniso_full): one test in the common case (fused kernels)
//   2  none: the caller guarantees !sc6_extras (plain march of the fused kernels: no branch inside the gradient)
template<int SPEC, int RARE = 0>
__device__ __forceinline__ D3 sc6_gradient(
    const StencilParams & p, const D3 & si, const D3 & xm, const D3 & xp, const D3 & bm, const D3 & bp, const D3 & cm,
    const D3 & cp, const double * __restrict__ ddi_plane, unsigned e )
{
    D3 g = make_d3( p.sc6_g0[0], p.sc6_g0[1], p.sc6_g0[2] );
    sc6_axis_gradient<0, ( SPEC & SC6_DMI_GENERAL ) != 0>( p, xm, xp, g );
    sc6_axis_gradient<1, ( SPEC & SC6_DMI_GENERAL ) != 0>( p, bm, bp, g );
    if( SPEC & SC6_HAS_C )
        sc6_axis_gradient<2, ( SPEC & SC6_DMI_GENERAL ) != 0>( p, cm, cp, g );
    g.x = fma( p.sc6_A[0], si.x, g.x );
    g.y = fma( p.sc6_A[1], si.y, g.y );
    g.z = fma( p.sc6_A[2], si.z, g.z );
    if( RARE == 0 && !SB_SC6_NO_ANISO_FULL && p.sc6_aniso_full ) // off-diagonal elements: anisotropy axes that are not lattice axes
    {
        g.x = fma( p.sc6_A[3], si.y, fma( p.sc6_A[4], si.z, g.x ) );
        g.y = fma( p.sc6_A[3], si.x, fma( p.sc6_A[5], si.z, g.y ) );
        g.z = fma( p.sc6_A[4], si.x, fma( p.sc6_A[5], si.y, g.z ) );
    }
    if( RARE != 2 && !SB_SC6_NO_EXTRAS && p.sc6_extras )
    {
        if( RARE == 1 && p.sc6_aniso_full )
        {
            g.x = fma( p.sc6_A[3], si.y, fma( p.sc6_A[4], si.z, g.x ) );
            g.y = fma( p.sc6_A[3], si.x, fma( p.sc6_A[5], si.z, g.y ) );
            g.z = fma( p.sc6_A[4], si.x, fma( p.sc6_A[5], si.y, g.z ) );
        }
        if( p.has_cubic )
        {
            const double k = -2.0 * p.K4[0];
            g.x            = fma( k * si.x, si.x * si.x, g.x );
            g.y            = fma( k * si.y, si.y * si.y, g.y );
            g.z            = fma( k * si.z, si.z * si.z, g.z );
        }
        if( p.has_ddi )
        {
            const D3 gd = ld3p( ddi_plane, e );
            g.x += gd.x;
            g.y += gd.y;
            g.z += gd.z;
        }
    }
    return g;
}

// Virtual force from the gradient g = -F (Method_LLG.cpp:131-226), signs folded into nc1 = -dtg/mu_s, nc2 = alpha nc1:
//   dynamics:      Fv = nc1 g + xi + s x (nc2 g + alpha xi)
//   minimisation:  Fv = -dtg' s x g
template<int MODE, bool WITH_STT = true>
__device__ __forceinline__ D3 sc6_virtual_force( const LLGParams & l, const D3 & s, const D3 & g, const D3 & xi )
{
    D3 w, fv;
    if( MODE == SC6_MINIMISE )
    {
        w  = make_d3( -l.dtg * g.x, -l.dtg * g.y, -l.dtg * g.z );
        fv = make_d3( 0.0, 0.0, 0.0 );
    }
    else
    {
        const double nc1 = l.nc1[0], nc2 = l.nc2[0];
        if( MODE == SC6_THERMAL )
        {
            w  = make_d3( fma( nc2, g.x, l.damping * xi.x ), fma( nc2, g.y, l.damping * xi.y ), fma( nc2, g.z, l.damping * xi.z ) );
            fv = make_d3( fma( nc1, g.x, xi.x ), fma( nc1, g.y, xi.y ), fma( nc1, g.z, xi.z ) );
        }
        else
        {
            w  = make_d3( nc2 * g.x, nc2 * g.y, nc2 * g.z );
            fv = make_d3( nc1 * g.x, nc1 * g.y, nc1 * g.z );
        }
    }
    fv.x = fma( s.y, w.z, fma( -s.z, w.y, fv.x ) );
    fv.y = fma( s.z, w.x, fma( -s.x, w.z, fv.y ) );
    fv.z = fma( s.x, w.y, fma( -s.y, w.x, fv.z ) );
    if( WITH_STT && !SB_SC6_NO_STT && MODE != SC6_MINIMISE && l.has_stt )
    {
        const D3 pol = make_d3( l.stt_pol[0], l.stt_pol[1], l.stt_pol[2] );
        const D3 pxs = cross3( pol, s );
        fv.x += l.stt_c1 * pol.x + l.stt_c2 * pxs.x;
        fv.y += l.stt_c1 * pol.y + l.stt_c2 * pxs.y;
        fv.z += l.stt_c1 * pol.z + l.stt_c2 * pxs.z;
    }
    return fv;
}

// xi of one site: Philox4x32-10 with the host-expanded round keys (LLGParams::philox_key) -> Box-Muller shaped in fp32
// (llg.cuh, gaussian3). The amplitude epsilon sqrt(T/mu_s) is folded into the radius: sqrt(k lg2 u) with the host
// constant k = -2 ln2 scale^2, and the angle uniforms stay in [1, 2) turns (sin / cos are periodic): 7 MUFU + 9 fp32
// instructions + 3 conversions per site.
__device__ __forceinline__ float3 sc6_thermal_field( const LLGParams & l, unsigned plane_site, unsigned gplane )
{
    unsigned c0 = plane_site, c1 = gplane, c2 = unsigned( l.iteration ), c3 = unsigned( l.iteration >> 32 );
#pragma unroll
    for( int r = 0; r < 10; ++r )
    {
        const std::uint64_t p0 = std::uint64_t( 0xD2511F53u ) * c0;
        const std::uint64_t p1 = std::uint64_t( 0xCD9E8D57u ) * c2;
        const unsigned n0 = unsigned( p1 >> 32 ) ^ c1 ^ l.philox_key[r][0];
        const unsigned n2 = unsigned( p0 >> 32 ) ^ c3 ^ l.philox_key[r][1];
        c1                = unsigned( p1 );
        c3                = unsigned( p0 );
        c0                = n0;
        c2                = n2;
    }
    return scaled_gaussian3f( c0, c1, c2, c3, l.thermal_k[0] );
}

// Element offset (inside the plane pointer of a field) of the plane that holds the c-neighbour `cc` (= c-1 or c+1,
// local index). The plane always exists in storage (periodic wrap on one device, halo planes on a slab); whether it
// CONTRIBUTES is sc6_c_valid.
__device__ __forceinline__ std::size_t sc6_c_plane( const StencilParams & p, int cc )
{
    if( p.halo == 0 )
    {
        if( cc < 0 )
            cc += p.Nc;
        else if( cc >= p.Nc )
            cc -= p.Nc;
        return std::size_t( cc ) * ( 3 * std::size_t( p.plane_stride ) );
    }
    return std::size_t( cc + p.halo ) * ( 3 * std::size_t( p.plane_stride ) );
}
__device__ __forceinline__ bool sc6_c_valid( const StencilParams & p, int cc )
{
    const int gc = p.c_begin + cc;
    return p.bc[2] || ( gc >= 0 && gc < p.Nc );
}

// What a thread holds of one configuration while it marches: its own column and the in-plane neighbours of the
// current plane.
struct SC6Window
{
    D3 below, center, above; // own column at c-1, c, c+1
    D3 xm, xp, bm, bp;       // in-plane neighbours at c
};

struct SC6Offsets
{
    unsigned ec, exm, exp_, ebm, ebp; // element offsets inside a plane
    bool vxm, vxp, vbm, vbp;          // neighbour contributes (open boundaries)
};

template<bool BOUNDARY>
__device__ __forceinline__ void sc6_load_inplane( SC6Window & w, const double * __restrict__ plane, const SC6Offsets & o )
{
    if( BOUNDARY )
    {
        w.xm = ld3pv( plane, o.exm, o.vxm );
        w.xp = ld3pv( plane, o.exp_, o.vxp );
        w.bm = ld3pv( plane, o.ebm, o.vbm );
        w.bp = ld3pv( plane, o.ebp, o.vbp );
    }
    else
    {
        w.xm = ld3p( plane, o.exm );
        w.xp = ld3p( plane, o.exp_ );
        w.bm = ld3p( plane, o.ebm );
        w.bp = ld3p( plane, o.ebp );
    }
}

// One plane step of a thread. (below, center, above) are its own-column registers for plane c; `below` is dead once
// the gradient has consumed it and receives the own-column value of plane c + 2, so that three calls with cyclically
// rotated arguments advance the march without a single register move.
// Software-pipelined by hand: right after the gradients of plane c have consumed the neighbour registers, the loads
// of plane c+1 are issued into them and travel while the thread does the remaining ~200 instructions of plane c
// (noise, virtual force, rotation, store).
template<int SOLVER, int STAGE, int SPEC, int MODE, bool BOUNDARY>
__device__ __forceinline__ void sc6_plane_step(
    const StencilParams & p, const LLGParams & l, const StageArgs & a, const SC6Offsets & o, const int c, const int c1,
    const std::size_t plane_elems, const unsigned plane_site, const unsigned gplane, D3 & s_below, const D3 & s_center, const D3 & s_above,
    D3 & p_below, const D3 & p_center, const D3 & p_above, SC6Window & ws, SC6Window & wp, float3 & xi_pipe )
{
    using Needs          = StageNeeds<SOLVER, STAGE>;
    constexpr bool HAS_C = ( SPEC & SC6_HAS_C ) != 0;
    constexpr bool COL_S = HAS_C && Needs::Fv_s;  // s needs its c-neighbours
    // T > 0: stage 1 makes the noise of the iteration (Philox + Box-Muller, ~70 instructions) and stores the three
    // fp32 variates it consists of (12 B per site); the later stages read them back one plane ahead instead of
    // regenerating them. Bit-identical to regenerating (the variates ARE fp32), 24 B more traffic per spin-step.
    constexpr bool XI_LOAD = MODE == SC6_THERMAL && STAGE > 1 && SB_SC6_STORE_XI;
    const float3 xi_f      = xi_pipe; // of plane c (XI_LOAD)
    constexpr bool COL_P = HAS_C && Needs::Fv_sp; // s' needs its c-neighbours
    const D3 zero        = make_d3( 0.0, 0.0, 0.0 );

    const std::size_t base = std::size_t( c + p.halo ) * plane_elems;
    bool vb = true, va = true;
    if( BOUNDARY && HAS_C )
    {
        vb = sc6_c_valid( p, c - 1 );
        va = sc6_c_valid( p, c + 1 );
    }

    // 1. gradients of plane c (consume the neighbour registers)
    D3 gs = zero, gp = zero;
    if( Needs::Fv_s )
        gs = sc6_gradient<SPEC>(
            p, s_center, ws.xm, ws.xp, ws.bm, ws.bp, ( BOUNDARY && !vb ) ? zero : s_below, ( BOUNDARY && !va ) ? zero : s_above,
            a.ddi_s.base + base, o.ec );
    if( Needs::Fv_sp )
        gp = sc6_gradient<SPEC>(
            p, p_center, wp.xm, wp.xp, wp.bm, wp.bp, ( BOUNDARY && !vb ) ? zero : p_below, ( BOUNDARY && !va ) ? zero : p_above,
            a.ddi_sp.base + base, o.ec );

    // 2. issue the loads of plane c + 1: in-plane neighbours, and the own-column value of plane c + 2 into `below`
    if( c + 1 < c1 )
    {
        const std::size_t base1 = base + plane_elems;
        const std::size_t pa2   = HAS_C ? sc6_c_plane( p, c + 2 ) : base1 + plane_elems;
        if( COL_S || c + 2 < c1 )
            s_below = ld3p( a.s.base + pa2, o.ec );
        if( Needs::Fv_s )
            sc6_load_inplane<BOUNDARY>( ws, a.s.base + base1, o );
        if( Needs::Fv_sp )
        {
            if( COL_P || c + 2 < c1 )
                p_below = ld3p( a.sp.base + pa2, o.ec );
            sc6_load_inplane<BOUNDARY>( wp, a.sp.base + base1, o );
        }
        if( SB_SC6_L2_PREFETCH > 0 && c + SB_SC6_L2_PREFETCH < c1 )
        {
            // the own-column load of plane c + 2 above is the one access of the step that has to come from HBM, with one
            // plane step of lookahead; pull the lines it will need into L2 well before
            const std::size_t far = base + std::size_t( SB_SC6_L2_PREFETCH ) * plane_elems + o.ec;
            if( Needs::Fv_s || true )
                sc6_prefetch3( a.s.base + far );
            if( Needs::Fv_sp )
                sc6_prefetch3( a.sp.base + far );
        }
        if( XI_LOAD )
        {
            const float * q = a.xi + base1 + o.ec;
            xi_pipe         = make_float3( __ldg( q ), __ldg( q + FIELD_BLOCK ), __ldg( q + 2 * FIELD_BLOCK ) );
        }
    }

    // 3. the rest of plane c
    D3 xi = zero;
    if( MODE == SC6_THERMAL )
    {
        float3 f = xi_f;
        if( !XI_LOAD )
        {
            f = sc6_thermal_field( l, plane_site, gplane );
            if( SB_SC6_STORE_XI && STAGE == 1 )
            {
                float * q          = a.xi + base + o.ec;
                q[0]               = f.x;
                q[FIELD_BLOCK]     = f.y;
                q[2 * FIELD_BLOCK] = f.z;
            }
        }
        xi = make_d3( double( f.x ), double( f.y ), double( f.z ) );
    }
    D3 Fv = zero, Fvp = zero;
    if( Needs::Fv_s )
        Fv = sc6_virtual_force<MODE>( l, s_center, gs, xi );
    if( Needs::Fv_sp )
        Fvp = sc6_virtual_force<MODE>( l, p_center, gp, xi );

    D3 acc = zero;
    if( SOLVER == Solver_RK4 && STAGE > 1 )
    {
        const double * q = a.acc.base + base + o.ec;
        acc              = make_d3( q[0], q[FIELD_BLOCK], q[2 * FIELD_BLOCK] );
    }
    const D3 out = solver_update<SOLVER, STAGE>( s_center, Fv, p_center, Fvp, acc );
    if( SOLVER == Solver_RK4 && STAGE < 4 )
    {
        double * q         = a.acc.base + base + o.ec;
        q[0]               = acc.x;
        q[FIELD_BLOCK]     = acc.y;
        q[2 * FIELD_BLOCK] = acc.z;
    }
    double * q         = a.out.base + base + o.ec;
    q[0]               = out.x;
    q[FIELD_BLOCK]     = out.y;
    q[2 * FIELD_BLOCK] = out.z;
}

// The march of one thread over the planes [c0, c1) of its column (x, b).
template<int SOLVER, int STAGE, int SPEC, int MODE, bool BOUNDARY>
__device__ __forceinline__ void sc6_march(
    const StencilParams & p, const LLGParams & l, const StageArgs & a, const int x, const int b, const int c0, const int c1 )
{
    using Needs          = StageNeeds<SOLVER, STAGE>;
    constexpr bool HAS_C = ( SPEC & SC6_HAS_C ) != 0;
    constexpr bool COL_S = HAS_C && Needs::Fv_s;
    constexpr bool COL_P = HAS_C && Needs::Fv_sp;

    // in-plane neighbours: site offsets -> element offsets (AoSoA-32)
    SC6Offsets o;
    {
        int xm = x - 1, xp = x + 1, bm = b - 1, bp = b + 1;
        o.vxm = o.vxp = o.vbm = o.vbp = true;
        if( xm < 0 )
        {
            xm += p.Na;
            o.vxm = p.bc[0];
        }
        if( xp >= p.Na )
        {
            xp -= p.Na;
            o.vxp = p.bc[0];
        }
        if( bm < 0 )
        {
            bm += p.Nb;
            o.vbm = p.bc[1];
        }
        if( bp >= p.Nb )
        {
            bp -= p.Nb;
            o.vbp = p.bc[1];
        }
        const int row = p.Na * b;
        o.ec          = unsigned( elem_offset( row + x ) );
        o.exm         = unsigned( elem_offset( row + xm ) );
        o.exp_        = unsigned( elem_offset( row + xp ) );
        o.ebm         = unsigned( elem_offset( p.Na * bm + x ) );
        o.ebp         = unsigned( elem_offset( p.Na * bp + x ) );
    }
    const std::size_t plane_elems = 3 * std::size_t( p.plane_stride );
    // Philox counter: (site inside the plane, global plane)
    const unsigned plane_site = unsigned( p.Na * b + x );
    unsigned gplane           = unsigned( p.c_begin + c0 );

    const D3 zero = make_d3( 0.0, 0.0, 0.0 );
    SC6Window ws, wp; // in-plane neighbours of s and of the predictor s'
    ws.xm = ws.xp = ws.bm = ws.bp = wp.xm = wp.xp = wp.bm = wp.bp = zero;
    // own-column rings: at plane c0 + k the roles (below, center, above) are ring[k % 3], ring[(k+1) % 3], ring[(k+2) % 3]
    D3 rs[3] = { zero, zero, zero }, rp[3] = { zero, zero, zero };
    float3 xi_pipe = make_float3( 0.f, 0.f, 0.f ); // stored noise of the next plane

    // prologue: plane c0 and the own-column value of plane c0 + 1
    {
        const std::size_t base = std::size_t( c0 + p.halo ) * plane_elems;
        const std::size_t pa   = HAS_C ? sc6_c_plane( p, c0 + 1 ) : base + plane_elems;
        rs[1]                  = ld3p( a.s.base + base, o.ec );
        if( COL_S || c0 + 1 < c1 )
            rs[2] = ld3p( a.s.base + pa, o.ec );
        if( COL_S )
            rs[0] = ld3p( a.s.base + sc6_c_plane( p, c0 - 1 ), o.ec );
        if( Needs::Fv_s )
            sc6_load_inplane<BOUNDARY>( ws, a.s.base + base, o );
        if( Needs::Fv_sp )
        {
            rp[1] = ld3p( a.sp.base + base, o.ec );
            if( COL_P || c0 + 1 < c1 )
                rp[2] = ld3p( a.sp.base + pa, o.ec );
            if( COL_P )
                rp[0] = ld3p( a.sp.base + sc6_c_plane( p, c0 - 1 ), o.ec );
            sc6_load_inplane<BOUNDARY>( wp, a.sp.base + base, o );
        }
        if( MODE == SC6_THERMAL && STAGE > 1 && SB_SC6_STORE_XI )
        {
            const float * q = a.xi + base + o.ec;
            xi_pipe         = make_float3( __ldg( q ), __ldg( q + FIELD_BLOCK ), __ldg( q + 2 * FIELD_BLOCK ) );
        }
    }

    for( int cb = c0; cb < c1; cb += 3 )
    {
#pragma unroll
        for( int k = 0; k < 3; ++k )
        {
            const int c = cb + k;
            if( k > 0 && c >= c1 )
                break;
            sc6_plane_step<SOLVER, STAGE, SPEC, MODE, BOUNDARY>(
                p, l, a, o, c, c1, plane_elems, plane_site, gplane, rs[k], rs[( k + 1 ) % 3], rs[( k + 2 ) % 3], rp[k], rp[( k + 1 ) % 3],
                rp[( k + 2 ) % 3], ws, wp, xi_pipe );
            ++gplane;
        }
    }
}

template<int SOLVER, int STAGE, int SPEC, int MODE>
static __global__ void __launch_bounds__( SC6Shape<SOLVER, STAGE>::threads, SC6Shape<SOLVER, STAGE>::min_blocks ) k_sc6_stage(
    const __grid_constant__ StencilParams p, const int lc, const int seg_first, const int seg_stride,
    const __grid_constant__ LLGParams l, const __grid_constant__ StageArgs a )
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y * blockDim.y + threadIdx.y;
    if( x >= p.Na || b >= p.Nb )
        return;
    const int c0 = ( seg_first + int( blockIdx.z ) * seg_stride ) * lc;
    const int c1 = min( c0 + lc, p.nc_local );

    // Does this CTA touch an open boundary? (uniform) Interior CTAs run the loop without predicates.
    const int x0 = blockIdx.x * blockDim.x, x1 = min( x0 + int( blockDim.x ), p.Na ) - 1;
    const int b0 = blockIdx.y * blockDim.y, b1 = min( b0 + int( blockDim.y ), p.Nb ) - 1;
    bool boundary = ( !p.bc[0] && ( x0 == 0 || x1 == p.Na - 1 ) ) || ( !p.bc[1] && ( b0 == 0 || b1 == p.Nb - 1 ) );
    if( ( SPEC & SC6_HAS_C ) && !p.bc[2] )
        boundary = boundary || ( p.c_begin + c0 == 0 ) || ( p.c_begin + c1 == p.Nc );
    if( boundary )
        sc6_march<SOLVER, STAGE, SPEC, MODE, true>( p, l, a, x, b, c0, c1 );
    else
        sc6_march<SOLVER, STAGE, SPEC, MODE, false>( p, l, a, x, b, c0, c1 );
}

// One launcher per solver, each in its own translation unit (sc6_<solver>.cu) so that the 8 SPEC x stage
// instantiations compile in parallel.
template<int SOLVER, int STAGE>
void sc6_launch_stage(
    const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a )
{
    const int mode = l.direct_minimization ? SC6_MINIMISE : ( l.has_thermal ? SC6_THERMAL : SC6_DYNAMICS );
#define SB_SC6_CASE( S, M )                                                                                            \
    case S * SC6_N_MODES + M: k_sc6_stage<SOLVER, STAGE, S, M><<<G.grid, G.block, 0, stream>>>( p, G.lc, G.seg_first, G.seg_stride, l, a ); break;
    const SC6Geometry & G = SC6Shape<SOLVER, STAGE>::two_windows ? L.two_windows : L.one_window;
    switch( L.spec * SC6_N_MODES + mode )
    {
        SB_SC6_CASE( 0, 0 )
        SB_SC6_CASE( 0, 1 )
        SB_SC6_CASE( 0, 2 )
        SB_SC6_CASE( 1, 0 )
        SB_SC6_CASE( 1, 1 )
        SB_SC6_CASE( 1, 2 )
        SB_SC6_CASE( 2, 0 )
        SB_SC6_CASE( 2, 1 )
        SB_SC6_CASE( 2, 2 )
        SB_SC6_CASE( 3, 0 )
        SB_SC6_CASE( 3, 1 )
        SB_SC6_CASE( 3, 2 )
    }
#undef SB_SC6_CASE
}

// defined in sc6_depondt.cu, sc6_heun.cu, sc6_sib.cu, sc6_rk4.cu
void sc6_launch_depondt( int stage, const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a );
void sc6_launch_heun( int stage, const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a );
void sc6_launch_sib( int stage, const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a );
void sc6_launch_rk4( int stage, const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a );

} // namespace dev
} // namespace sb
