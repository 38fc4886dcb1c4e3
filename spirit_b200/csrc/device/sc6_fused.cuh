// Both stages of a predictor-corrector solver (Depondt, Heun, SIB) in ONE kernel for the nearest-neighbour stencil:
// every iteration reads the spins once and writes them once (48 B per spin-step instead of the 120 B of the two-pass
// scheme of sc6.cuh), and the corrector re-uses the predictor's virtual force and noise instead of recomputing them.
//
// Reference semantics: Solver_Depondt.hpp:29-77, Solver_Heun.hpp:30-81, Solver_SIB.hpp:22-50 (predictor from Fv(s),
// corrector from Fv(s')), Method_LLG.cpp:131-226 (virtual force), :246-301 (hook quantities).
//
// The corrector at a site needs the predictor s' at the site's six neighbours, so s' cannot stay thread-private:
//   * a CTA owns a tile of TX x TY sites of a plane and MARCHES along c like the kernels of sc6.cuh;
//   * it computes the predictor on the tile PLUS its one-site rim (TY + 2 row warps and one "side" warp for the two rim
//     columns; 19 warps for 16 rows of results = 19 % redundant predictor work) and one plane before / after its
//     c-segment, and passes s' between threads through shared memory (a ring of 4 plane buffers, one CTA barrier per
//     plane);
//   * the corrector of plane c runs one plane step behind the predictor of plane c + 1. Own-column values of s (ring of
//     4 planes: the corrector still needs s(c) when the predictor has moved on), the predictor's virtual force, noise
//     and s' centre / above stay in registers; s' below and the four in-plane s' neighbours come from shared memory.
//   * the thermal field is a pure function of (site in plane, global plane, iteration) (Philox counter), so the rim
//     sites recomputed by neighbouring CTAs get the same noise: results are bit-identical to the two-pass kernels.
// Not served here (the two-pass kernels remain): dipolar field (needs a global transform of s' between the stages),
// RK4, spin-transfer torque, temperature gradients, Nc == 1 with c-neighbours.
#pragma once

#include "sc6.cuh"

namespace sb
{
namespace dev
{

#ifndef SB_FUSED_TY
#define SB_FUSED_TY 16
#endif
#ifndef SB_FUSED_PREFETCH
#define SB_FUSED_PREFETCH 0 // 1: in-plane neighbours of s one plane ahead in registers (as sc6.cuh); 0: loaded when needed
#endif
#ifndef SB_FUSED_NOISE_FIRST
#define SB_FUSED_NOISE_FIRST 1 // the noise of a plane is generated between the issue of the plane's loads and their first use
#endif
#ifndef SB_FUSED_RARE_SPLIT
#define SB_FUSED_RARE_SPLIT 1 // interior CTAs of a Hamiltonian without rare terms run a march without the uniform test for them
#endif
#ifndef SB_FUSED_S1_CONST
#define SB_FUSED_S1_CONST 0 // interior CTAs with 16 rows: "this thread predicts" is a compile-time true
#endif
#ifndef SB_FUSED_PIN_SLOT
#define SB_FUSED_PIN_SLOT 1
#endif
#ifndef SB_FUSED_PBELOW_REG
#define SB_FUSED_PBELOW_REG 0 // 1: s'(c-1) of the own column kept in registers instead of re-read from shared memory
#endif

constexpr int FUSED_TX      = 32;
constexpr int FUSED_TY      = SB_FUSED_TY; // <= 16: the two rim columns must fit one warp
constexpr int FUSED_WARPS   = FUSED_TY + 3;
constexpr int FUSED_THREADS = 32 * FUSED_WARPS;
constexpr int FUSED_RS      = FUSED_TX + 2;                // row stride of a shared plane buffer (doubles)
constexpr int FUSED_CS      = ( FUSED_TY + 2 ) * FUSED_RS; // component stride
constexpr int FUSED_BS      = 3 * FUSED_CS;                // buffer stride
constexpr int FUSED_NBUF    = 4;
constexpr int FUSED_RED          = FUSED_NBUF * FUSED_BS;           // hook reduction scratch: 2 doubles per warp
constexpr int FUSED_STASH        = FUSED_RED + 2 * FUSED_WARPS + 2; // hook: gradient of s, two plane buffers (own slots only)
constexpr std::size_t FUSED_SMEM = std::size_t( FUSED_STASH ) * sizeof( double );
constexpr std::size_t FUSED_SMEM_HOOK = std::size_t( FUSED_STASH + 2 * FUSED_BS ) * sizeof( double );
static_assert( 2 * FUSED_TY <= 32, "rim columns must fit one warp" );

struct FusedArgs
{
    ConstField3 s;            // configuration at the start of the iteration
    Field3 out;               // configuration at the end of the iteration
    Field3 F_out;             // HOOK: effective field -gradient(s), projected tangentially to the NEW spins
    double * energy_partials; // HOOK: per CTA, energy of the predictor configuration (the last force evaluation)
    double * torque_partials; // HOOK: per CTA, max |Fv - (Fv.s_new) s_new|^2 with Fv the predictor's virtual force
    // Slab decomposition with peer-mapped memory (NVLink): the CTAs that produce the first / last `halo` planes of the slab
    // also store them into the halo planes of the neighbouring ranks' `out` field -- the halo exchange is part of this
    // kernel. Null: no neighbour on that side (or the exchange is done by the host: NCCL send / recv).
    double * peer_lo_out; // `out` field of the rank below: my planes c < halo are its planes halo + peer_lo_nc + c
    double * peer_hi_out; // `out` field of the rank above: my planes c >= nc_local - halo are its planes c - (nc_local - halo)
    int peer_lo_nc;       // nc_local of the rank below
};

struct FusedGeometry
{
    dim3 grid;
    int lc         = 1;
    int seg_first  = 0;
    int seg_stride = 1;
};

// storage plane (inside a field) of the local plane m, m in [-2, nc_local + 1]
__device__ __forceinline__ unsigned fused_plane( const StencilParams & p, int m )
{
    if( p.halo == 0 )
    {
        if( m < 0 )
            m += p.Nc;
        else if( m >= p.Nc )
            m -= p.Nc;
        return unsigned( m );
    }
    return unsigned( m + p.halo );
}
// first element of a storage plane: a 32 x 32 -> 64 bit product of two uniform values (one uniform-datapath instruction;
// with a 64-bit plane size the compiler does this arithmetic per thread)
__device__ __forceinline__ std::uint64_t fused_plane_offset( unsigned plane, unsigned plane_elems )
{
    return std::uint64_t( plane ) * plane_elems;
}

// Philox counter word of the local plane q: its GLOBAL plane (periodic wrap for the planes recomputed beyond the lattice ends)
__device__ __forceinline__ unsigned fused_global_plane( const StencilParams & p, int q )
{
    int gq = p.c_begin + q;
    if( gq < 0 )
        gq += p.Nc;
    else if( gq >= p.Nc )
        gq -= p.Nc;
    return unsigned( gq );
}

template<bool BOUNDARY>
__device__ __forceinline__ D3 fused_lds3( const double * buf, int so, bool valid )
{
    D3 r = make_d3( 0.0, 0.0, 0.0 );
    if( !BOUNDARY || valid )
        r = make_d3( buf[so], buf[so + FUSED_CS], buf[so + 2 * FUSED_CS] );
    return r;
}

// Deterministic two-level reduction of the hook quantities of a CTA: warp tree, then thread 0 folds the warps in order.
__device__ __forceinline__ void fused_reduce_hook( double * sm, double e_acc, double t_acc, const FusedArgs & a, const int cta )
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double * red = sm + FUSED_RED;
    e_acc        = warp_sum( e_acc );
    t_acc        = warp_max( t_acc );
    __syncthreads();
    if( lane == 0 )
    {
        red[warp]               = e_acc;
        red[FUSED_WARPS + warp] = t_acc;
    }
    __syncthreads();
    if( threadIdx.x == 0 )
    {
        double e = 0.0, t = 0.0;
        for( int w = 0; w < FUSED_WARPS; ++w )
        {
            e += red[w];
            t = fmax( t, red[FUSED_WARPS + w] );
        }
        a.energy_partials[cta] = e;
        a.torque_partials[cta] = t;
    }
}

// RARE: how sc6_gradient reaches the rare terms (1: behind one uniform flag, 2: the Hamiltonian has none -- no test at all)
template<int SOLVER, int SPEC, int MODE, bool HOOK, bool BOUNDARY, int RARE>
__device__ __forceinline__ void sc6_fused_march(
    const StencilParams & p, const LLGParams & l, const FusedArgs & a, double * __restrict__ sm, const int c0, const int c1, const int cta )
{
    constexpr bool HAS_C = ( SPEC & SC6_HAS_C ) != 0;
    constexpr int LAG    = HAS_C ? 1 : 0; // the corrector runs LAG planes behind the predictor
    const D3 zero        = make_d3( 0.0, 0.0, 0.0 );

    // ---- role of this thread: tile position (i, j), i in [-1, TX], j in [-1, TY] ------------------------------------------
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = lane, j = warp - 1;
    bool s1 = true; // computes the predictor of its site (interior CTAs with 16 rows: every thread, known at compile time)
    if( warp >= FUSED_TY + 2 )
    {
        i = lane < FUSED_TY ? -1 : FUSED_TX;
        j = lane < FUSED_TY ? lane : lane - FUSED_TY;
        if( BOUNDARY || FUSED_TY != 16 || !SB_FUSED_S1_CONST )
            s1 = lane < 2 * FUSED_TY;
    }
    int x = int( blockIdx.x ) * FUSED_TX + i, b = int( blockIdx.y ) * FUSED_TY + j;
    // computes the corrector (the result) of its site
    const bool s2 = warp < FUSED_TY + 2 && j >= 0 && j < FUSED_TY && x < p.Na && b < p.Nb;
    if( x < 0 )
    {
        x += p.Na;
        if( BOUNDARY || !SB_FUSED_S1_CONST )
            s1 = s1 && p.bc[0];
    }
    else if( x >= p.Na )
    {
        if( BOUNDARY || !SB_FUSED_S1_CONST )
            s1 = s1 && p.bc[0] && x == p.Na;
        x = 0;
    }
    if( b < 0 )
    {
        b += p.Nb;
        if( BOUNDARY || !SB_FUSED_S1_CONST )
            s1 = s1 && p.bc[1];
    }
    else if( b >= p.Nb )
    {
        if( BOUNDARY || !SB_FUSED_S1_CONST )
            s1 = s1 && p.bc[1] && b == p.Nb;
        b = 0;
    }
    int so = ( j + 1 ) * FUSED_RS + ( i + 1 ); // own slot in a shared plane buffer
#if SB_FUSED_PIN_SLOT
    // opaque to the compiler from here on: it would otherwise recompute the slot from the thread index (a dozen instructions)
    // before every use instead of keeping it in a register
    asm volatile( "" : "+r"( so ) );
#endif

    // ---- in-plane neighbours of the site in global memory (as sc6_march) ---------------------------------------------
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
    const unsigned plane_elems = 3u * unsigned( p.plane_stride );
    const unsigned plane_site  = unsigned( p.Na * b + x );

    // ---- the march ----------------------------------------------------------------------------------------------------
    // body k handles the predictor of plane q = q0 + k and the corrector of plane q - LAG
    const int q0      = c0 - LAG;
    const int nbodies = c1 - c0 + 2 * LAG;

    SC6Window ws;
    ws.xm = ws.xp = ws.bm = ws.bp = zero;
    D3 r[4] = { zero, zero, zero, zero }; // own column of s: plane m in slot (m - (q0 - 1)) & 3
    if( s1 )
    {
        const double * sb = a.s.base;
        if( HAS_C )
        {
            r[0] = ld3p( sb + fused_plane_offset( fused_plane( p, q0 - 1 ), plane_elems ), o.ec );
            r[2] = ld3p( sb + fused_plane_offset( fused_plane( p, q0 + 1 ), plane_elems ), o.ec );
        }
        const double * pl = sb + fused_plane_offset( fused_plane( p, q0 ), plane_elems );
        r[1]              = ld3p( pl, o.ec );
        if( SB_FUSED_PREFETCH )
            sc6_load_inplane<BOUNDARY>( ws, pl, o );
    }
    D3 Fv_prev = zero, p_center = zero, p_below = zero;
    float3 xi_prev = make_float3( 0.f, 0.f, 0.f );
    double e_acc = 0.0, t_acc = 0.0;

    for( int kb = 0; kb < nbodies; kb += 4 )
    {
#pragma unroll
        for( int u = 0; u < 4; ++u )
        {
            const int k = kb + u;
            if( u > 0 && k >= nbodies )
                break;
            const int q = q0 + k;
            D3 & s_below        = r[u];
            const D3 & s_center = r[( u + 1 ) & 3];
            const D3 & s_above  = r[( u + 2 ) & 3];
            D3 & s_incoming     = r[( u + 3 ) & 3];
            double * buf_q      = sm + ( ( u + 1 ) & 3 ) * FUSED_BS; // s'(q)
            const double * buf_c = sm + ( ( u + 1 - LAG ) & 3 ) * FUSED_BS; // s'(q - LAG): the corrector's plane
            const double * buf_b = sm + ( ( u - LAG ) & 3 ) * FUSED_BS;     // s'(q - LAG - 1)

            // ---- predictor of plane q ------------------------------------------------------------------------------------
            bool q_valid = true, vb = true, va = true;
            if( BOUNDARY && HAS_C )
            {
                q_valid = sc6_c_valid( p, q );
                vb      = sc6_c_valid( p, q - 1 );
                va      = sc6_c_valid( p, q + 1 );
            }
            D3 Fv = zero, gs = zero, spn = zero;
            float3 xi = make_float3( 0.f, 0.f, 0.f );
            if( s1 )
            {
                const double * sb = a.s.base;
                // Order matters to ptxas: (1) the loads, (2) the noise of the plane (Philox + Box-Muller, ~75 integer / SFU
                // instructions that depend on nothing), (3) the gradient that consumes the loads.
                if( !SB_FUSED_PREFETCH )
                    sc6_load_inplane<BOUNDARY>( ws, sb + fused_plane_offset( fused_plane( p, q ), plane_elems ), o );
                D3 xid = zero;
                if( SB_FUSED_NOISE_FIRST && MODE == SC6_THERMAL && q_valid )
                {
                    xi  = sc6_thermal_field( l, plane_site, fused_global_plane( p, q ) );
                    xid = make_d3( double( xi.x ), double( xi.y ), double( xi.z ) );
                }
                if( q_valid )
                    gs = sc6_gradient<SPEC, RARE>(
                        p, s_center, ws.xm, ws.xp, ws.bm, ws.bp, ( BOUNDARY && !vb ) ? zero : s_below,
                        ( BOUNDARY && !va ) ? zero : s_above, nullptr, 0u );
                // loads of the next planes travel under the rest of this body (also when plane q itself lies beyond an open
                // end of the lattice and is skipped)
                if( k + 1 < nbodies )
                {
                    if( SB_FUSED_PREFETCH )
                        sc6_load_inplane<BOUNDARY>( ws, sb + fused_plane_offset( fused_plane( p, q + 1 ), plane_elems ), o );
                    // (the own-column value fetched ahead comes from HBM: issued after the gradient so that it does not share
                    // a scoreboard with the in-plane loads the gradient waits for)
                    if( HAS_C )
                        s_incoming = ld3p( sb + fused_plane_offset( fused_plane( p, q + 2 ), plane_elems ), o.ec );
                    else
                        r[( u + 2 ) & 3] = ld3p( sb + fused_plane_offset( fused_plane( p, q + 1 ), plane_elems ), o.ec );
                }
                if( q_valid )
                {
                    if( !SB_FUSED_NOISE_FIRST && MODE == SC6_THERMAL )
                    {
                        xi  = sc6_thermal_field( l, plane_site, fused_global_plane( p, q ) );
                        xid = make_d3( double( xi.x ), double( xi.y ), double( xi.z ) );
                    }
                    Fv = sc6_virtual_force<MODE, false>( l, s_center, gs, xid );
                    D3 acc = zero;
                    spn    = solver_update<SOLVER, 1>( s_center, Fv, s_center, zero, acc );
                    buf_q[so]                = spn.x;
                    buf_q[so + FUSED_CS]     = spn.y;
                    buf_q[so + 2 * FUSED_CS] = spn.z;
                    if( HOOK && HAS_C )
                    {
                        // the corrector of this plane (next body, same thread) turns the gradient into the projected
                        // effective field: through shared memory instead of 6 registers held across the body
                        double * st         = sm + FUSED_STASH + ( u & 1 ) * FUSED_BS + so;
                        st[0]               = gs.x;
                        st[FUSED_CS]        = gs.y;
                        st[2 * FUSED_CS]    = gs.z;
                    }
                }
            }
            __syncthreads();

            // ---- corrector of plane c = q - LAG ------------------------------------------------------------------------
            if( s2 && k >= 2 * LAG )
            {
                const int c = q - LAG;
                bool vb2 = true, va2 = true;
                if( BOUNDARY && HAS_C )
                {
                    vb2 = sc6_c_valid( p, c - 1 );
                    va2 = sc6_c_valid( p, c + 1 );
                }
                const D3 pc = HAS_C ? p_center : spn;
                const D3 nxm = fused_lds3<BOUNDARY>( buf_c, so - 1, o.vxm );
                const D3 nxp = fused_lds3<BOUNDARY>( buf_c, so + 1, o.vxp );
                const D3 nbm = fused_lds3<BOUNDARY>( buf_c, so - FUSED_RS, o.vbm );
                const D3 nbp = fused_lds3<BOUNDARY>( buf_c, so + FUSED_RS, o.vbp );
                D3 pb = zero, pa = zero;
                if( HAS_C )
                {
                    pb = SB_FUSED_PBELOW_REG ? ( ( BOUNDARY && !vb2 ) ? zero : p_below ) : fused_lds3<BOUNDARY>( buf_b, so, vb2 );
                    pa = ( BOUNDARY && !va2 ) ? zero : spn;
                }
                const D3 gp = sc6_gradient<SPEC, RARE>( p, pc, nxm, nxp, nbm, nbp, pb, pa, nullptr, 0u );
                if( HOOK )
                {
                    // energy of the predictor configuration from its total gradient (site_energy, stencil.cuh):
                    //   1/2 g_bilinear . s + E_cubic + E_zeeman  =  1/2 (g + g0) . s + K4/2 sum s^4      (g0 = -mu_s B)
                    // (first thing after the gradient: nothing of it has to stay in registers for the end of the corrector)
                    double e = 0.5 * ( ( gp.x + p.sc6_g0[0] ) * pc.x + ( gp.y + p.sc6_g0[1] ) * pc.y + ( gp.z + p.sc6_g0[2] ) * pc.z );
                    if( RARE != 2 && p.has_cubic )
                    {
                        const double x2 = pc.x * pc.x, y2 = pc.y * pc.y, z2 = pc.z * pc.z;
                        e += 0.5 * p.K4[0] * ( x2 * x2 + y2 * y2 + z2 * z2 );
                    }
                    e_acc += e;
                }
                const float3 xf = HAS_C ? xi_prev : xi;
                D3 xid          = zero;
                if( MODE == SC6_THERMAL )
                    xid = make_d3( double( xf.x ), double( xf.y ), double( xf.z ) );
                const D3 Fvp  = sc6_virtual_force<MODE, false>( l, pc, gp, xid );
                const D3 si   = HAS_C ? s_below : s_center; // s(c)
                const D3 Fv_s = HAS_C ? Fv_prev : Fv;
                D3 acc        = zero;
                const D3 out  = solver_update<SOLVER, 2>( si, Fv_s, pc, Fvp, acc );
                double * qo   = a.out.base + fused_plane_offset( unsigned( c + p.halo ), plane_elems ) + o.ec;
                qo[0]               = out.x;
                qo[FIELD_BLOCK]     = out.y;
                qo[2 * FIELD_BLOCK] = out.z;
                if( HAS_C && p.halo > 0 )
                {
                    // halo exchange by peer stores (uniform tests: only the CTAs at the slab ends get here with a pointer)
                    if( a.peer_lo_out && c < p.halo )
                    {
                        double * qp = a.peer_lo_out + fused_plane_offset( unsigned( p.halo + a.peer_lo_nc + c ), plane_elems ) + o.ec;
                        qp[0]               = out.x;
                        qp[FIELD_BLOCK]     = out.y;
                        qp[2 * FIELD_BLOCK] = out.z;
                    }
                    if( a.peer_hi_out && c >= p.nc_local - p.halo )
                    {
                        double * qp = a.peer_hi_out + fused_plane_offset( unsigned( c - ( p.nc_local - p.halo ) ), plane_elems ) + o.ec;
                        qp[0]               = out.x;
                        qp[FIELD_BLOCK]     = out.y;
                        qp[2 * FIELD_BLOCK] = out.z;
                    }
                }
                if( HOOK )
                {
                    const double d = dot3( Fv_s, out );
                    const D3 tq    = make_d3( Fv_s.x - d * out.x, Fv_s.y - d * out.y, Fv_s.z - d * out.z );
                    t_acc          = fmax( t_acc, dot3( tq, tq ) );
                    D3 g1 = gs; // F = -g
                    if( HAS_C )
                    {
                        const double * st = sm + FUSED_STASH + ( ( u + 1 ) & 1 ) * FUSED_BS + so;
                        g1                = make_d3( st[0], st[FUSED_CS], st[2 * FUSED_CS] );
                    }
                    const double f = dot3( g1, out );
                    double * qf    = a.F_out.base + fused_plane_offset( unsigned( c + p.halo ), plane_elems ) + o.ec;
                    qf[0]               = f * out.x - g1.x;
                    qf[FIELD_BLOCK]     = f * out.y - g1.y;
                    qf[2 * FIELD_BLOCK] = f * out.z - g1.z;
                }
            }
            if( HAS_C )
            {
                Fv_prev = Fv;
                xi_prev = xi;
                if( SB_FUSED_PBELOW_REG )
                    p_below = p_center;
                p_center = spn;
            }
        }
    }

    if( HOOK )
        fused_reduce_hook( sm, e_acc, t_acc, a, cta );
}


template<int SOLVER, int SPEC, int MODE, bool HOOK>
static __global__ void __launch_bounds__( FUSED_THREADS, 1 ) k_sc6_fused(
    const __grid_constant__ StencilParams p, const int lc, const int seg_first, const int seg_stride,
    const __grid_constant__ LLGParams l, const __grid_constant__ FusedArgs a )
{
    extern __shared__ double fused_smem[];
    const int seg = seg_first + int( blockIdx.z ) * seg_stride;
    const int c0  = seg * lc;
    const int c1  = min( c0 + lc, p.nc_local );
    const int cta = blockIdx.x + gridDim.x * ( blockIdx.y + gridDim.y * seg ); // slot of the hook partials (over ALL segments)

    // Does this CTA touch an open boundary or the ragged edge of the lattice? (uniform)
    const int x0 = blockIdx.x * FUSED_TX, b0 = blockIdx.y * FUSED_TY;
    bool boundary = x0 + FUSED_TX > p.Na || b0 + FUSED_TY > p.Nb;
    boundary      = boundary || ( !p.bc[0] && ( x0 == 0 || x0 + FUSED_TX == p.Na ) ) || ( !p.bc[1] && ( b0 == 0 || b0 + FUSED_TY == p.Nb ) );
    if( ( SPEC & SC6_HAS_C ) && !p.bc[2] )
        boundary = boundary || ( p.c_begin + c0 <= 1 ) || ( p.c_begin + c1 >= p.Nc - 1 );
    if( boundary )
        sc6_fused_march<SOLVER, SPEC, MODE, HOOK, true, 1>( p, l, a, fused_smem, c0, c1, cta );
    else if( !SB_FUSED_RARE_SPLIT || p.sc6_extras )
        sc6_fused_march<SOLVER, SPEC, MODE, HOOK, false, 1>( p, l, a, fused_smem, c0, c1, cta );
    else
        sc6_fused_march<SOLVER, SPEC, MODE, HOOK, false, 2>( p, l, a, fused_smem, c0, c1, cta );
}

template<int SOLVER, int SPEC, int MODE, bool HOOK>
void sc6_fused_launch_one( const FusedGeometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const FusedArgs & a )
{
    static bool configured = false;
    if( !configured )
    {
        cudaFuncSetAttribute(
            k_sc6_fused<SOLVER, SPEC, MODE, HOOK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int( HOOK ? FUSED_SMEM_HOOK : FUSED_SMEM ) );
        configured = true;
    }
    k_sc6_fused<SOLVER, SPEC, MODE, HOOK><<<G.grid, FUSED_THREADS, HOOK ? FUSED_SMEM_HOOK : FUSED_SMEM, stream>>>(
        p, G.lc, G.seg_first, G.seg_stride, l, a );
}

template<int SOLVER, bool HOOK>
void sc6_fused_launch_solver( int spec, const FusedGeometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const FusedArgs & a )
{
    const int mode = l.direct_minimization ? SC6_MINIMISE : ( l.has_thermal ? SC6_THERMAL : SC6_DYNAMICS );
#define SB_FUSED_CASE( S, M )                                                                                          \
    case S * SC6_N_MODES + M: sc6_fused_launch_one<SOLVER, S, M, HOOK>( G, stream, p, l, a ); break;
    switch( spec * SC6_N_MODES + mode )
    {
        SB_FUSED_CASE( 0, 0 )
        SB_FUSED_CASE( 0, 1 )
        SB_FUSED_CASE( 0, 2 )
        SB_FUSED_CASE( 1, 0 )
        SB_FUSED_CASE( 1, 1 )
        SB_FUSED_CASE( 1, 2 )
        SB_FUSED_CASE( 2, 0 )
        SB_FUSED_CASE( 2, 1 )
        SB_FUSED_CASE( 2, 2 )
        SB_FUSED_CASE( 3, 0 )
        SB_FUSED_CASE( 3, 1 )
        SB_FUSED_CASE( 3, 2 )
    }
#undef SB_FUSED_CASE
}

// defined in sc6_fused_<solver>.cu
void sc6_fused_depondt( bool hook, int spec, const FusedGeometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const FusedArgs & a );
void sc6_fused_heun( bool hook, int spec, const FusedGeometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const FusedArgs & a );
void sc6_fused_sib( bool hook, int spec, const FusedGeometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const FusedArgs & a );

} // namespace dev
} // namespace sb
