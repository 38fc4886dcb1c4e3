// Tile-staged variant of the nearest-neighbour marching kernels (included at the end of sc6.cuh).
//
// The register march of sc6.cuh prefetches the four in-plane neighbours of the next plane into registers. That
// holds 24 registers per configuration in flight for a whole plane step: stage 2 of Depondt / Heun (two
// configurations) needs 128 registers, i.e. 16 warps per SM, and the step is latency-bound at that occupancy
// (ncu, profiles/r1f: 3.4-4.0 warps stalled on the long scoreboard per issued instruction, issue slots 50-57 % busy;
// 12 warps/SM: +34 % time). It also fetches every plane twice from L2 (once as own-column prefetch, once as
// in-plane neighbours): ~120 B per site cross the L2->SM fabric in stage 2, 60 % of its measured ceiling.
//
// Here the planes travel through shared memory instead, moved by the bulk-copy engine (TMA, cp.async.bulk), so a
// load in flight holds no register and every plane crosses L2->SM once (plus the halo rows):
//   * a CTA owns a tile of BX x BY sites (BX = a whole number of AoSoA-32 blocks that divides the row, so a tile row
//     of one plane is ONE contiguous run of BX/32 x 768 bytes in HBM) and marches along c;
//   * per plane and configuration, one warp issues BY + 2 row copies (the tile rows plus the rows above and below it;
//     if the tile is narrower than the lattice row, each row carries one extra AoSoA block per side) into a ring of
//     NS stage buffers; completion is signalled on an mbarrier per stage (expect-tx byte count);
//   * at plane c the threads read their four in-plane neighbours from tile(c) and their +c neighbour from
//     tile(c+1) with LDS (29 cycles, short scoreboard), keep (c-1, c) of their own column in registers, and after a
//     CTA barrier the producer warp refills the buffer of tile(c) with tile(c+NS).
// Arithmetic (gradient, noise, virtual force, solver update) is shared with the register march.
//
// Requires Na % 32 == 0 (rows start on AoSoA block boundaries); everything else runs through sc6_march.
#pragma once

#include <stdexcept>
#include <string>

namespace sb
{
namespace dev
{

// ---- helpers shared with nothing: arithmetic of a site, written for the tile march --------------------------------
// Addresses inside the march are a UNIFORM 64-bit plane pointer plus a per-thread 32-bit BYTE offset: one
// IADD3 + IADD3.X per neighbour, the three components at immediate offsets +0 / +256 / +512 (AoSoA-32).
__device__ __forceinline__ D3 ld3pb( const char * __restrict__ plane, unsigned off )
{
    const double * q = reinterpret_cast<const double *>( plane + off );
    return make_d3( __ldg( q ), __ldg( q + FIELD_BLOCK ), __ldg( q + 2 * FIELD_BLOCK ) );
}
// the same with an ELEMENT offset
__device__ __forceinline__ D3 ld3p( const char * __restrict__ plane, unsigned e )
{
    const double * q = reinterpret_cast<const double *>( plane ) + e;
    return make_d3( __ldg( q ), __ldg( q + FIELD_BLOCK ), __ldg( q + 2 * FIELD_BLOCK ) );
}

__device__ __forceinline__ const char * bytes( const double * q )
{
    return reinterpret_cast<const char *>( q );
}


// Gradient of all terms at one site, in two parts so that a caller can retire the in-plane neighbours before the
// c-neighbours arrive: everything except the pairs along c, then those.
//   start value: -mu_s B n (Zeeman, Hamiltonian_Heisenberg.cpp:768-783; zero without a field)
//   on-site quadratic form: g += A s with A = -2 sum_k K_k n_k n_k^T (uniaxial anisotropies, :785-800)
//   GENERAL only: off-diagonal A, cubic anisotropy (:802-820), precomputed dipolar field. The host routes a
//   Hamiltonian with any of these to the GENERAL variant of the march, so the plain variant tests no flags.
template<int SPEC, bool GENERAL, bool DDI_BYTES = false>
__device__ __forceinline__ D3 sc6t_gradient_inplane(
    const StencilParams & p, const D3 & si, const D3 & xm, const D3 & xp, const D3 & bm, const D3 & bp,
    const char * __restrict__ ddi_plane, unsigned off )
{
    D3 g = make_d3( p.sc6_g0[0], p.sc6_g0[1], p.sc6_g0[2] );
    sc6_axis_gradient<0, ( SPEC & SC6_DMI_GENERAL ) != 0>( p, xm, xp, g );
    sc6_axis_gradient<1, ( SPEC & SC6_DMI_GENERAL ) != 0>( p, bm, bp, g );
    g.x = fma( p.sc6_A[0], si.x, g.x );
    g.y = fma( p.sc6_A[1], si.y, g.y );
    g.z = fma( p.sc6_A[2], si.z, g.z );
    if( GENERAL && p.sc6_extras )
    {
        if( p.sc6_aniso_full ) // off-diagonal elements: anisotropy axes that are not lattice axes
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
            const D3 gd = DDI_BYTES ? ld3pb( ddi_plane, off ) : ld3p( ddi_plane, off );
            g.x += gd.x;
            g.y += gd.y;
            g.z += gd.z;
        }
    }
    return g;
}

// Virtual force from the gradient g = -F (Method_LLG.cpp:131-226), signs folded into nc1 = -dtg/mu_s, nc2 = alpha nc1:
//   dynamics:      Fv = nc1 g + xi + s x (nc2 g + alpha xi)
//   minimisation:  Fv = -dtg s x g
// (spin-transfer torque never reaches these kernels: launch_stage sends it to the generic ones)
template<int MODE>
__device__ __forceinline__ D3 sc6t_virtual_force( const LLGParams & l, const D3 & s, const D3 & g, const D3 & xi )
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
    return fv;
}

// Depondt corrector axis (Solver_Depondt.hpp:60-70): H = (Fv(s) + Fv(s'))/2 in one pass, with the factor 1/2 folded
// into host constants (h1 = nc1/2, h2 = nc2/2, ah = alpha/2; hd = -dtg/2):
//   dynamics:      H = h1 (g + g') + xi + s x (h2 g + ah xi) + s' x (h2 g' + ah xi)
//   minimisation:  H = s x (hd g) + s' x (hd g')
template<int MODE>
__device__ __forceinline__ D3
sc6t_depondt_mean_force( const LLGParams & l, const D3 & s, const D3 & g, const D3 & sp, const D3 & gp, const D3 & xi )
{
    D3 w, wp, H;
    if( MODE == SC6_MINIMISE )
    {
        const double hd = l.half_ndtg;
        w  = make_d3( hd * g.x, hd * g.y, hd * g.z );
        wp = make_d3( hd * gp.x, hd * gp.y, hd * gp.z );
        H  = make_d3( s.y * w.z, s.z * w.x, s.x * w.y );
    }
    else
    {
        const double h1 = l.half_nc1[0], h2 = l.half_nc2[0];
        const D3 gs     = make_d3( g.x + gp.x, g.y + gp.y, g.z + gp.z );
        if( MODE == SC6_THERMAL )
        {
            const D3 ax = make_d3( l.half_damping * xi.x, l.half_damping * xi.y, l.half_damping * xi.z );
            w  = make_d3( fma( h2, g.x, ax.x ), fma( h2, g.y, ax.y ), fma( h2, g.z, ax.z ) );
            wp = make_d3( fma( h2, gp.x, ax.x ), fma( h2, gp.y, ax.y ), fma( h2, gp.z, ax.z ) );
            H  = make_d3( fma( h1, gs.x, xi.x ), fma( h1, gs.y, xi.y ), fma( h1, gs.z, xi.z ) );
        }
        else
        {
            w  = make_d3( h2 * g.x, h2 * g.y, h2 * g.z );
            wp = make_d3( h2 * gp.x, h2 * gp.y, h2 * gp.z );
            H  = make_d3( h1 * gs.x, h1 * gs.y, h1 * gs.z );
        }
        H.x = fma( s.y, w.z, H.x );
        H.y = fma( s.z, w.x, H.y );
        H.z = fma( s.x, w.y, H.z );
    }
    H.x = fma( sp.y, wp.z, fma( -sp.z, wp.y, fma( -s.z, w.y, H.x ) ) );
    H.y = fma( sp.z, wp.x, fma( -sp.x, wp.z, fma( -s.x, w.z, H.y ) ) );
    H.z = fma( sp.x, wp.y, fma( -sp.y, wp.x, fma( -s.y, w.x, H.z ) ) );
    return H;
}


// BYTE offset (from a field's base pointer) of the plane that holds the c-neighbour `cc` (= c-1 or c+1, local
// index). The plane always exists in storage (periodic wrap on one device, halo planes on a slab); whether it
// CONTRIBUTES is sc6_c_valid.
__device__ __forceinline__ std::size_t sc6t_c_plane( const StencilParams & p, int cc, std::size_t plane_bytes )
{
    if( p.halo == 0 )
    {
        if( cc < 0 )
            cc += p.Nc;
        else if( cc >= p.Nc )
            cc -= p.Nc;
        return std::size_t( cc ) * plane_bytes;
    }
    return std::size_t( cc + p.halo ) * plane_bytes;
}

// Uniform state of the march: where the planes are (byte offsets from a field's base, equal for all fields)
struct SC6TPlanes
{
    std::size_t cur;    // plane c
    std::size_t stride; // bytes per plane
    std::size_t last_above; // plane holding the +c neighbour of the segment's last plane (periodic wrap / halo)
};


// The rest of a site once its gradients are known: virtual forces, solver update, store. `off` is the byte offset
// of the site from a field's base pointer.
template<int SOLVER, int STAGE, int MODE>
__device__ __forceinline__ void sc6t_finish_site(
    const LLGParams & l, const StageArgs & a, const std::size_t off, const D3 & s_center, const D3 & gs, const D3 & p_center,
    const D3 & gp, const D3 & xi )
{
    using Needs   = StageNeeds<SOLVER, STAGE>;
    const D3 zero = make_d3( 0.0, 0.0, 0.0 );
    D3 out;
    if( SOLVER == Solver_Depondt && STAGE == 2 )
        out = rotate_about( s_center, sc6t_depondt_mean_force<MODE>( l, s_center, gs, p_center, gp, xi ) );
    else
    {
        D3 Fv = zero, Fvp = zero;
        if( Needs::Fv_s )
            Fv = sc6t_virtual_force<MODE>( l, s_center, gs, xi );
        if( Needs::Fv_sp )
            Fvp = sc6t_virtual_force<MODE>( l, p_center, gp, xi );
        D3 acc = zero;
        if( SOLVER == Solver_RK4 && STAGE > 1 )
        {
            const double * q = reinterpret_cast<const double *>( bytes( a.acc.base ) + off );
            acc              = make_d3( q[0], q[FIELD_BLOCK], q[2 * FIELD_BLOCK] );
        }
        out = solver_update<SOLVER, STAGE>( s_center, Fv, p_center, Fvp, acc );
        if( SOLVER == Solver_RK4 && STAGE < 4 )
        {
            double * q         = reinterpret_cast<double *>( reinterpret_cast<char *>( a.acc.base ) + off );
            q[0]               = acc.x;
            q[FIELD_BLOCK]     = acc.y;
            q[2 * FIELD_BLOCK] = acc.z;
        }
    }
    double * q         = reinterpret_cast<double *>( reinterpret_cast<char *>( a.out.base ) + off );
    q[0]               = out.x;
    q[FIELD_BLOCK]     = out.y;
    q[2 * FIELD_BLOCK] = out.z;
}


constexpr int SC6T_MAX_STAGES  = 4;
constexpr int SC6T_BLOCK_BYTES = 3 * FIELD_BLOCK * int( sizeof( double ) ); // one AoSoA-32 block: 768 B

__device__ __forceinline__ unsigned smem_u32( const void * q )
{
    return unsigned( __cvta_generic_to_shared( q ) );
}
__device__ __forceinline__ void mbar_init( unsigned bar, unsigned count )
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( bar ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void mbar_expect_tx( unsigned bar, unsigned bytes )
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( bar ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void mbar_arrive( unsigned bar )
{
    asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( bar ) : "memory" );
}
__device__ __forceinline__ bool mbar_try_wait( unsigned bar, unsigned parity )
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"( ok )
        : "r"( bar ), "r"( parity ), "r"( 10000000u ) // suspend-time hint (ns): sleep in hardware instead of spinning
        : "memory" );
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait( unsigned bar, unsigned parity )
{
    while( !mbar_try_wait( bar, parity ) )
    {
    }
}
// global -> shared bulk copy (TMA engine, no registers, no LSU slots); bytes and both addresses multiples of 16
__device__ __forceinline__ void bulk_g2s( unsigned dst, const void * src, unsigned bytes, unsigned bar )
{
    asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( dst ),
                  "l"( src ), "r"( bytes ), "r"( bar )
                  : "memory" );
}

__device__ __forceinline__ D3 lds3( unsigned addr )
{
    D3 r;
    asm volatile( "ld.shared.f64 %0, [%1];" : "=d"( r.x ) : "r"( addr ) );
    asm volatile( "ld.shared.f64 %0, [%1+256];" : "=d"( r.y ) : "r"( addr ) );
    asm volatile( "ld.shared.f64 %0, [%1+512];" : "=d"( r.z ) : "r"( addr ) );
    return r;
}
__device__ __forceinline__ D3 lds3v( unsigned addr, bool valid )
{
    const D3 r = lds3( addr );
    return valid ? r : make_d3( 0.0, 0.0, 0.0 );
}

// The bulk copies that bring one plane's tile into a stage buffer, as a table in shared memory: entry i (one per
// (configuration, tile row), at most 32) holds up to three copies -- the main run of AoSoA blocks, extended by the
// halo blocks where they are contiguous with it, and the wrapped halo blocks of a periodic row end. Offsets are
// relative to the plane (source) and to the stage buffer (destination); bytes == 0: no copy. Rows / halo blocks that
// do not exist (open boundary, beyond the lattice) are skipped: the consumers never use them.
struct SC6TileCopies
{
    unsigned src[3], dst[3], bytes[3];
    unsigned field; // 0: first configuration of the stage, 1: second
};
constexpr int SC6T_MAX_ITEMS = 32;

__device__ __forceinline__ void sc6t_plan_copies(
    const StencilParams & p, const SC6TileShape & t, const int nf, SC6TileCopies * table, unsigned * total_bytes, const int lane )
{
    const int nblk = p.Na / FIELD_BLOCK, bxb = t.bx / FIELD_BLOCK, rows = t.by + 2;
    const int b0 = blockIdx.y * t.by, xblk0 = blockIdx.x * bxb;
    SC6TileCopies e;
    for( int k = 0; k < 3; ++k )
        e.src[k] = e.dst[k] = e.bytes[k] = 0;
    e.field = 0;
    if( lane < nf * rows )
    {
        const int f = lane / rows, r = lane - f * rows; // tile row r holds lattice row b0 + r - 1
        int gb      = b0 + r - 1;
        if( gb == -1 )
            gb = p.bc[1] ? p.Nb - 1 : -1;
        else if( gb == p.Nb )
            gb = p.bc[1] ? 0 : -1;
        else if( gb > p.Nb )
            gb = -1;
        if( gb >= 0 )
        {
            const unsigned row = unsigned( gb ) * unsigned( nblk ) * SC6T_BLOCK_BYTES;
            const unsigned dst = unsigned( f * t.field_bytes + r * t.pitch );
            int first = xblk0, count = bxb;
            unsigned d = dst;
            if( t.xhalo )
            {
                d += SC6T_BLOCK_BYTES;
                if( xblk0 > 0 )
                {
                    --first;
                    ++count;
                    d -= SC6T_BLOCK_BYTES;
                }
                else if( p.bc[0] )
                {
                    e.src[1]   = row + unsigned( nblk - 1 ) * SC6T_BLOCK_BYTES;
                    e.dst[1]   = dst;
                    e.bytes[1] = SC6T_BLOCK_BYTES;
                }
                if( xblk0 + bxb < nblk )
                    ++count;
                else if( p.bc[0] )
                {
                    e.src[2]   = row;
                    e.dst[2]   = dst + unsigned( bxb + 1 ) * SC6T_BLOCK_BYTES;
                    e.bytes[2] = SC6T_BLOCK_BYTES;
                }
            }
            e.src[0]   = row + unsigned( first ) * SC6T_BLOCK_BYTES;
            e.dst[0]   = d;
            e.bytes[0] = unsigned( count ) * SC6T_BLOCK_BYTES;
            e.field    = unsigned( f );
        }
    }
    table[lane] = e;
    const unsigned sum = __reduce_add_sync( 0xffffffffu, e.bytes[0] + e.bytes[1] + e.bytes[2] );
    if( lane == 0 )
        *total_bytes = sum;
}

// One warp: copy the tile of the plane at byte offset `plane_off` into the stage buffer `stage`
__device__ __forceinline__ void sc6t_issue_tile(
    const SC6TileCopies * table, const unsigned total_bytes, const char * base0, const char * base1, const std::size_t plane_off,
    const unsigned stage, const unsigned bar, const int lane )
{
    if( lane == 0 )
    {
        // the buffer was read through the generic proxy; the copies write it through the async proxy
        asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
        mbar_expect_tx( bar, total_bytes );
    }
    __syncwarp();
    const SC6TileCopies e = table[lane];
    const char * plane    = ( e.field ? base1 : base0 ) + plane_off;
#pragma unroll
    for( int k = 0; k < 3; ++k )
        if( e.bytes[k] )
            bulk_g2s( stage + e.dst[k], plane + e.src[k], e.bytes[k], bar );
}

// Per-thread addresses inside a tile (bytes from the start of one configuration's tile) and boundary flags
struct SC6TileSite
{
    unsigned c, xm, xp; // own site and the +-a neighbours; +-b are c -+ pitch
    unsigned ec;        // global byte offset of the own site inside a plane (stores, own-site loads)
    bool vxm, vxp, vbm, vbp;
};

// Shared bookkeeping of a CTA
struct SC6TileControl
{
    SC6TileCopies copies[SC6T_MAX_ITEMS];
    unsigned total_bytes;               // of one tile (expect-tx count)
    unsigned warps;                     // warps of the CTA that own lattice sites
    unsigned released[SC6T_MAX_STAGES]; // warps that are done with the tile in a stage buffer
};

struct SC6TileRing // stage / phase bookkeeping of tile(c), uniform
{
    int stage;
    unsigned phase;
};
__device__ __forceinline__ SC6TileRing next_tile( const SC6TileRing & r, int nstage )
{
    SC6TileRing n = r;
    if( ++n.stage == nstage )
    {
        n.stage = 0;
        n.phase ^= 1u;
    }
    return n;
}

template<int SOLVER, int STAGE, int SPEC, int MODE, bool GENERAL>
__device__ __forceinline__ void sc6t_march(
    const StencilParams & p, const LLGParams & l, const StageArgs & a, const SC6TileShape & t, const unsigned tiles,
    const unsigned full, SC6TileControl & ctl, const int c0, const int c1 )
{
    using Needs          = StageNeeds<SOLVER, STAGE>;
    constexpr bool HAS_C = ( SPEC & SC6_HAS_C ) != 0;
    constexpr int NF     = ( Needs::Fv_s && Needs::Fv_sp ) ? 2 : 1;
    // tile configuration 0 is s where its gradient is needed, else s'; configuration 1 (two-window stages) is s'
    constexpr int TILE_S = 0, TILE_P = Needs::Fv_s ? 1 : 0;
    const D3 zero        = make_d3( 0.0, 0.0, 0.0 );

    SC6TPlanes pl;
    pl.stride     = 3 * std::size_t( p.plane_stride ) * sizeof( double );
    pl.cur        = std::size_t( c0 + p.halo ) * pl.stride;
    pl.last_above = HAS_C ? sc6t_c_plane( p, c1, pl.stride ) : 0;
    const int last_tile = HAS_C ? c1 : c1 - 1; // planes c0 .. last_tile are staged

    const char * base0  = Needs::Fv_s ? bytes( a.s.base ) : bytes( a.sp.base );
    const char * base1  = bytes( a.sp.base );

    const int tid = threadIdx.x, lane = tid & 31;
    const int ty = tid / t.bx, tx = tid - ty * t.bx;
    const int x = blockIdx.x * t.bx + tx, b = blockIdx.y * t.by + ty;
    if( b >= p.Nb )
        return; // whole warps (a warp covers 32 sites of one row); ctl.warps counts the remaining ones

    // prologue (warp 0): fill the ring
    if( tid < 32 )
        for( int k = 0; k < t.nstage && c0 + k <= last_tile; ++k )
            sc6t_issue_tile(
                ctl.copies, ctl.total_bytes, base0, base1, ( c0 + k == c1 ) ? pl.last_above : pl.cur + std::size_t( k ) * pl.stride,
                tiles + unsigned( k * t.stage_bytes ), full + 8u * k, lane );

    SC6TileSite o;
    {
        const int shift = t.xhalo ? FIELD_BLOCK : 0;
        const int pc = tx + shift;
        int pm = pc - 1, pp = pc + 1;
        if( !t.xhalo )
        {
            pm = tx == 0 ? t.bx - 1 : tx - 1;
            pp = tx == t.bx - 1 ? 0 : tx + 1;
        }
        const unsigned row = unsigned( ( ty + 1 ) * t.pitch );
        o.c   = row + unsigned( ( pc >> 5 ) * SC6T_BLOCK_BYTES + ( pc & 31 ) * 8 );
        o.xm  = row + unsigned( ( pm >> 5 ) * SC6T_BLOCK_BYTES + ( pm & 31 ) * 8 );
        o.xp  = row + unsigned( ( pp >> 5 ) * SC6T_BLOCK_BYTES + ( pp & 31 ) * 8 );
        o.ec  = unsigned( elem_offset( p.Na * b + x ) ) * 8u;
        o.vxm = p.bc[0] || x > 0;
        o.vxp = p.bc[0] || x < p.Na - 1;
        o.vbm = p.bc[1] || b > 0;
        o.vbp = p.bc[1] || b < p.Nb - 1;
    }
    // Philox counter: (site inside the plane, global plane) -- sc6_thermal_field
    const unsigned plane_site = unsigned( p.Na * b + x );
    unsigned gplane           = unsigned( p.c_begin + c0 );

    // own-column rings: at plane c0 + k the roles (below, center, above) are ring[k % 3], ring[(k+1) % 3], ring[(k+2) % 3]
    D3 rs[3] = { zero, zero, zero }, rp[3] = { zero, zero, zero };
    if( HAS_C )
    {
        const std::size_t below = sc6t_c_plane( p, c0 - 1, pl.stride );
        if( Needs::Fv_s )
            rs[0] = ld3pb( bytes( a.s.base ) + below, o.ec );
        if( Needs::Fv_sp )
            rp[0] = ld3pb( bytes( a.sp.base ) + below, o.ec );
    }
    if( !Needs::Fv_s )
        rs[1] = ld3pb( bytes( a.s.base ) + pl.cur, o.ec );
    SC6TileRing cur{ 0, 0u };

    for( int cb = c0; cb < c1; cb += 3 )
    {
#pragma unroll
        for( int k = 0; k < 3; ++k )
        {
            const int c = cb + k;
            if( k > 0 && c >= c1 )
                break;
            D3 & s_below = rs[k], &s_center = rs[( k + 1 ) % 3], &s_above = rs[( k + 2 ) % 3];
            D3 & p_below = rp[k], &p_center = rp[( k + 1 ) % 3], &p_above = rp[( k + 2 ) % 3];
            const unsigned tile = tiles + unsigned( cur.stage * t.stage_bytes );

            // s without a gradient in this stage (SIB stage 2, RK4 stages 2-4): own site straight from HBM, one plane ahead
            if( !Needs::Fv_s && c + 1 < c1 )
                s_above = ld3pb( bytes( a.s.base ) + pl.cur + pl.stride, o.ec );
            // tile(c): waited for as tile(c+1) of the previous step, except on the first plane / in 2-D
            if( !HAS_C || c == c0 )
            {
                mbar_wait( full + 8u * cur.stage, cur.phase );
                if( Needs::Fv_s )
                    s_center = lds3( tile + TILE_S * t.field_bytes + o.c );
                if( Needs::Fv_sp )
                    p_center = lds3( tile + TILE_P * t.field_bytes + o.c );
            }
            // in-plane neighbours from tile(c), folded into the gradient right away (24 registers -> 6)
            D3 gs = zero, gp = zero;
            if( Needs::Fv_s )
            {
                const unsigned q = tile + TILE_S * t.field_bytes;
                const D3 xm = GENERAL ? lds3v( q + o.xm, o.vxm ) : lds3( q + o.xm );
                const D3 xp = GENERAL ? lds3v( q + o.xp, o.vxp ) : lds3( q + o.xp );
                const D3 bm = GENERAL ? lds3v( q + o.c - t.pitch, o.vbm ) : lds3( q + o.c - t.pitch );
                const D3 bp = GENERAL ? lds3v( q + o.c + t.pitch, o.vbp ) : lds3( q + o.c + t.pitch );
                gs = sc6t_gradient_inplane<SPEC, GENERAL, true>( p, s_center, xm, xp, bm, bp, bytes( a.ddi_s.base ) + pl.cur, o.ec );
            }
            if( Needs::Fv_sp )
            {
                const unsigned q = tile + TILE_P * t.field_bytes;
                const D3 xm = GENERAL ? lds3v( q + o.xm, o.vxm ) : lds3( q + o.xm );
                const D3 xp = GENERAL ? lds3v( q + o.xp, o.vxp ) : lds3( q + o.xp );
                const D3 bm = GENERAL ? lds3v( q + o.c - t.pitch, o.vbm ) : lds3( q + o.c - t.pitch );
                const D3 bp = GENERAL ? lds3v( q + o.c + t.pitch, o.vbp ) : lds3( q + o.c + t.pitch );
                gp = sc6t_gradient_inplane<SPEC, GENERAL, true>( p, p_center, xm, xp, bm, bp, bytes( a.ddi_sp.base ) + pl.cur, o.ec );
            }
            // This warp is done with tile(c). The last warp of the CTA to get here refills the buffer with
            // tile(c + NS): no CTA-wide barrier, no warp ever waits for anything but data.
            {
                unsigned last = 0;
                __syncwarp();
                if( lane == 0 )
                {
                    // (shared-memory operations of a warp are performed in order: the reads above precede the release)
                    last = atomicAdd( &ctl.released[cur.stage], 1u ) == ctl.warps - 1u;
                    if( last )
                        ctl.released[cur.stage] = 0;
                }
                last = __shfl_sync( 0xffffffffu, last, 0 );
                if( last && c + t.nstage <= last_tile )
                    sc6t_issue_tile(
                        ctl.copies, ctl.total_bytes, base0, base1,
                        ( c + t.nstage == c1 ) ? pl.last_above : pl.cur + std::size_t( t.nstage ) * pl.stride, tile,
                        full + 8u * cur.stage, lane );
            }
            // noise next: it needs no data and covers what is left of the wait for tile(c+1)
            D3 xi = zero;
            if( MODE == SC6_THERMAL )
                xi = sc6_thermal_field( l, plane_site, gplane );
            if( HAS_C )
            {
                const SC6TileRing nxt = next_tile( cur, t.nstage );
                mbar_wait( full + 8u * nxt.stage, nxt.phase );
                const unsigned q = tiles + unsigned( nxt.stage * t.stage_bytes );
                bool vb = true, va = true;
                if( GENERAL )
                {
                    vb = sc6_c_valid( p, c - 1 );
                    va = sc6_c_valid( p, c + 1 );
                }
                if( Needs::Fv_s )
                {
                    s_above = lds3( q + TILE_S * t.field_bytes + o.c );
                    sc6_axis_gradient<2, ( SPEC & SC6_DMI_GENERAL ) != 0>(
                        p, ( GENERAL && !vb ) ? zero : s_below, ( GENERAL && !va ) ? zero : s_above, gs );
                }
                if( Needs::Fv_sp )
                {
                    p_above = lds3( q + TILE_P * t.field_bytes + o.c );
                    sc6_axis_gradient<2, ( SPEC & SC6_DMI_GENERAL ) != 0>(
                        p, ( GENERAL && !vb ) ? zero : p_below, ( GENERAL && !va ) ? zero : p_above, gp );
                }
            }
            sc6t_finish_site<SOLVER, STAGE, MODE>( l, a, pl.cur + o.ec, s_center, gs, p_center, gp, xi );

            cur = next_tile( cur, t.nstage );
            pl.cur += pl.stride;
            ++gplane;
        }
    }
}

template<int SOLVER, int STAGE>
struct SC6TShape
{
    static constexpr bool two_windows = StageNeeds<SOLVER, STAGE>::Fv_s && StageNeeds<SOLVER, STAGE>::Fv_sp;
    static constexpr int threads = two_windows ? SC6T_THREADS_2W : SC6T_THREADS_1W;
};

// Block: bx * by threads, thread -> site (tid % bx, tid / bx) of the tile.
template<int SOLVER, int STAGE, int SPEC, int MODE>
static __global__ void __launch_bounds__( SC6TShape<SOLVER, STAGE>::threads, 1 ) k_sc6t_stage(
    const __grid_constant__ StencilParams p, const int lc, const int seg_first, const int seg_stride,
    const __grid_constant__ SC6TileShape t, const __grid_constant__ LLGParams l, const __grid_constant__ StageArgs a )
{
    extern __shared__ __align__( 128 ) unsigned char sc6t_smem[];
    __shared__ __align__( 8 ) unsigned long long sc6t_full[SC6T_MAX_STAGES];
    __shared__ SC6TileControl ctl;

    const int c0 = ( seg_first + int( blockIdx.z ) * seg_stride ) * lc;
    const int c1 = min( c0 + lc, p.nc_local );
    const int b0 = blockIdx.y * t.by, b1 = min( b0 + t.by, p.Nb ) - 1;
    if( threadIdx.x < 32 )
    {
        constexpr int NF = SC6TShape<SOLVER, STAGE>::two_windows ? 2 : 1;
        sc6t_plan_copies( p, t, NF, ctl.copies, &ctl.total_bytes, threadIdx.x );
        if( threadIdx.x == 0 )
        {
            ctl.warps = unsigned( ( b1 - b0 + 1 ) * ( t.bx / 32 ) );
            for( int k = 0; k < t.nstage; ++k )
            {
                ctl.released[k] = 0;
                mbar_init( smem_u32( &sc6t_full[k] ), 1 );
            }
            asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
        }
    }
    __syncthreads();

    // Does this CTA touch an open boundary, or does the Hamiltonian have one of the rare terms? (uniform)
    const int x0 = blockIdx.x * t.bx, x1 = x0 + t.bx - 1;
    bool general = ( !p.bc[0] && ( x0 == 0 || x1 == p.Na - 1 ) ) || ( !p.bc[1] && ( b0 == 0 || b1 == p.Nb - 1 ) ) || p.sc6_extras;
    if( ( SPEC & SC6_HAS_C ) && !p.bc[2] )
        general = general || ( p.c_begin + c0 == 0 ) || ( p.c_begin + c1 == p.Nc );
    const unsigned tiles = smem_u32( sc6t_smem ), full = smem_u32( sc6t_full );
    if( general )
        sc6t_march<SOLVER, STAGE, SPEC, MODE, true>( p, l, a, t, tiles, full, ctl, c0, c1 );
    else
        sc6t_march<SOLVER, STAGE, SPEC, MODE, false>( p, l, a, t, tiles, full, ctl, c0, c1 );
}

// Opt in to > 48 KB of dynamic shared memory (per kernel instantiation, whenever a launch needs more than any
// before), then launch
template<typename Kernel>
void sc6t_launch( Kernel kernel, const SC6Geometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a )
{
    static int configured_bytes = 48 * 1024;
    if( G.tile.smem_bytes > configured_bytes )
    {
        const cudaError_t err = cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G.tile.smem_bytes );
        if( err != cudaSuccess )
            throw std::runtime_error( std::string( "spirit_b200: cudaFuncSetAttribute(MaxDynamicSharedMemorySize): " ) + cudaGetErrorString( err ) );
        configured_bytes = G.tile.smem_bytes;
    }
    kernel<<<G.grid, G.block, G.tile.smem_bytes, stream>>>( p, G.lc, G.seg_first, G.seg_stride, G.tile, l, a );
}

} // namespace dev
} // namespace sb
