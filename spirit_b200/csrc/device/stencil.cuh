// Device-side gather stencil for the Heisenberg gradient (fp64, SoA).
//
// One thread owns one spin and only ever writes its own site (gather formulation, race-free by
// construction -- the same convention as the reference's OpenMP/CUDA builds, which double every
// pair, Hamiltonian_Heisenberg.cpp:103-109,130-139,165-175).
//
// Data layout in HBM: AoSoA-32 (below). The contiguous index of a row is x = ib + NB*a, rows are ordered by b,
// planes by c_local + halo. A warp reads 32 consecutive doubles per component and neighbour: fully coalesced
// 256-B requests; the +-a neighbours hit the same lines (L1), the +-b rows are re-used inside the CTA, the
// +-c planes are re-used out of L2 (generic kernels) or kept in registers (marching kernels, sc6.cuh).
#pragma once

#include "params.hpp"

#include <cuda_runtime.h>

#include <cstddef>

namespace sb
{
namespace dev
{

struct D3
{
    double x, y, z;
};

__device__ __forceinline__ D3 make_d3( double x, double y, double z )
{
    D3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
// Explicit fused multiply-adds in a fixed order: which product the compiler would contract into an FMA depends on the
// code around the call, and every kernel (and every code path inside a kernel) must produce the same bits for a site.
__device__ __forceinline__ double dot3( const D3 & a, const D3 & b )
{
    return fma( a.z, b.z, fma( a.y, b.y, a.x * b.x ) );
}
__device__ __forceinline__ D3 add3( const D3 & a, const D3 & b )
{
    return make_d3( a.x + b.x, a.y + b.y, a.z + b.z );
}
__device__ __forceinline__ D3 cross3( const D3 & a, const D3 & b )
{
    return make_d3( fma( a.y, b.z, -( a.z * b.y ) ), fma( a.z, b.x, -( a.x * b.z ) ), fma( a.x, b.y, -( a.y * b.x ) ) );
}

// Field layout in HBM: "AoSoA-32". Sites are grouped in blocks of 32 consecutive storage indices; a block stores
// its 32 x components, then its 32 y, then its 32 z (768 B). A warp reading one component of 32 consecutive sites
// reads 256 contiguous bytes (two full lines), and the three components of a site sit at fixed byte offsets
// +0 / +256 / +512 from one address, so a neighbour costs ONE address computation and three loads with immediate
// offsets (planar x[], y[], z[] arrays cost three 64-bit address computations per neighbour).
constexpr int FIELD_BLOCK = 32;

__device__ __host__ __forceinline__ std::size_t elem_offset( std::size_t idx )
{
    return ( idx >> 5 ) * ( 3 * FIELD_BLOCK ) + ( idx & ( FIELD_BLOCK - 1 ) );
}

struct ConstField3
{
    const double * __restrict__ base;
};
struct Field3
{
    double * __restrict__ base;
};

__device__ __forceinline__ D3 load3( const ConstField3 & f, std::size_t idx )
{
    const double * q = f.base + elem_offset( idx );
    return make_d3( __ldg( q ), __ldg( q + FIELD_BLOCK ), __ldg( q + 2 * FIELD_BLOCK ) );
}
// plain (coherent) load from a buffer that other kernels of the same stream write
__device__ __forceinline__ D3 load3( const Field3 & f, std::size_t idx )
{
    const double * q = f.base + elem_offset( idx );
    return make_d3( q[0], q[FIELD_BLOCK], q[2 * FIELD_BLOCK] );
}
__device__ __forceinline__ void store3( const Field3 & f, std::size_t idx, const D3 & v )
{
    double * q          = f.base + elem_offset( idx );
    q[0]               = v.x;
    q[FIELD_BLOCK]     = v.y;
    q[2 * FIELD_BLOCK] = v.z;
}

// Coordinates of the site a thread owns
struct Site
{
    int a, ib;   // cell index along a and basis atom
    int b, c;    // row: b, and LOCAL plane index c (0 .. nc_local)
    int idx;     // storage index into the planar arrays
};

// Storage index of (x, b, c_local): planes are padded to a multiple of 32 sites (p.plane_stride) so that a step in
// c moves every site by the same number of field elements.
__device__ __forceinline__ int storage_index( const StencilParams & p, int x, int b, int c_local )
{
    return x + p.Na * p.NB * b + p.plane_stride * ( c_local + p.halo );
}

// Gradient of the pair terms + uniaxial anisotropy ("bilinear" terms: E = 1/2 g.s), and of the
// cubic anisotropy and Zeeman terms (energies need their own expressions).
// Reference: Gradient_and_Energy, Hamiltonian_Heisenberg.cpp:704-766.
struct SiteGradient
{
    D3 bilinear; // anisotropy + exchange + DMI (+ DDI field if present)
    D3 rest;     // cubic anisotropy + Zeeman
    unsigned flags; // FLAG_* of the site (0 on a lattice without pinned sites / defects)
};

// WITH_FLAGS: the lattice has vacancies to test for (a second copy of the loop, so that the loop of a plain lattice keeps its
// independent, predicated neighbour loads: with the flag test in the one loop the film iteration went 5.3 -> 7.1 ms, profiles/r2zi)
template<int NB_T, bool WITH_FLAGS>
__device__ __forceinline__ D3 pair_gradient_loop( const StencilParams & p, const ConstField3 & s, const Site & site )
{
    D3 g           = make_d3( 0, 0, 0 );
    const int NB   = NB_T > 0 ? NB_T : p.NB;
    const int ib   = NB_T == 1 ? 0 : site.ib;
    const int nbeg = p.neigh_begin[ib];
    const int nend = p.neigh_begin[ib + 1];
#pragma unroll 2
    for( int n = nbeg; n < nend; ++n )
    {
        const Neighbour & nb = p.neigh[n];
        // Translate and apply boundary conditions (idx_from_pair, Vectormath.hpp:437-528).
        // |translation| <= N is guaranteed by the host (larger ones are dropped there), so a
        // single wrap suffices.
        int ja = site.a + nb.da, jb = site.b + nb.db, jc = site.c + nb.dc;
        bool valid = true;
        if( ja < 0 )
        {
            ja += p.Na;
            valid = valid && p.bc[0];
        }
        else if( ja >= p.Na )
        {
            ja -= p.Na;
            valid = valid && p.bc[0];
        }
        if( jb < 0 )
        {
            jb += p.Nb;
            valid = valid && p.bc[1];
        }
        else if( jb >= p.Nb )
        {
            jb -= p.Nb;
            valid = valid && p.bc[1];
        }
        if( p.halo == 0 )
        {
            // whole lattice on this device: wrap locally
            if( jc < 0 )
            {
                jc += p.Nc;
                valid = valid && p.bc[2];
            }
            else if( jc >= p.Nc )
            {
                jc -= p.Nc;
                valid = valid && p.bc[2];
            }
        }
        else
        {
            // slab: the halo planes hold the neighbour's data; only the global range decides validity
            const int gc = p.c_begin + jc;
            valid        = valid && ( p.bc[2] || ( gc >= 0 && gc < p.Nc ) );
        }
        if( valid )
        {
            const int j = storage_index( p, ja * NB + nb.jb, jb, jc );
            if( WITH_FLAGS && ( __ldg( p.site_flags + j ) & FLAG_VACANT ) )
                continue; // idx_from_pair: the pair does not exist
            const D3 sj = load3( s, j );
            // g -= J s_j + s_j x D
            g.x -= nb.J * sj.x + ( sj.y * nb.Dz - sj.z * nb.Dy );
            g.y -= nb.J * sj.y + ( sj.z * nb.Dx - sj.x * nb.Dz );
            g.z -= nb.J * sj.z + ( sj.x * nb.Dy - sj.y * nb.Dx );
        }
    }
    return g;
}

template<int NB_T>
__device__ __forceinline__ D3 pair_gradient( const StencilParams & p, const ConstField3 & s, const Site & site )
{
    if( p.site_flags )
        return pair_gradient_loop<NB_T, true>( p, s, site );
    return pair_gradient_loop<NB_T, false>( p, s, site );
}

template<int NB_T>
__device__ __forceinline__ SiteGradient
site_gradient( const StencilParams & p, const ConstField3 & s, const ConstField3 & ddi, const Site & site, const D3 & si )
{
    SiteGradient out;
    const int ib = NB_T == 1 ? 0 : site.ib;
    // Site flags enter as factors, not as branches: an early return for vacancies keeps ptxas from issuing the loads of a
    // site together (k_vp_a on the 2048 x 2048 x 4 film: 0.53 -> 2.36 ms, long-scoreboard stalls per issue 5 -> 47, profiles/r2zk)
    const unsigned flags = p.site_flags ? __ldg( p.site_flags + site.idx ) : 0u;
    out.flags            = flags;
    const double present = ( flags & FLAG_VACANT ) ? 0.0 : 1.0;                    // vacancy: every term vanishes
    const double moment  = ( flags & ( FLAG_VACANT | FLAG_NO_MU_S ) ) ? 0.0 : 1.0; // no moment: no Zeeman, no dipolar term

    D3 g = pair_gradient<NB_T>( p, s, site );

    // Uniaxial anisotropy: g -= 2 K (n.s) n   (Hamiltonian_Heisenberg.cpp:785-800)
    for( int i = 0; i < p.n_aniso; ++i )
    {
        const Anisotropy & an = p.aniso[i];
        if( an.ib == ib )
        {
            const double c = 2.0 * an.K * ( an.nx * si.x + an.ny * si.y + an.nz * si.z );
            g.x -= c * an.nx;
            g.y -= c * an.ny;
            g.z -= c * an.nz;
        }
    }
    // Dipole-dipole field, precomputed by the FFT convolution for this configuration
    if( p.has_ddi )
    {
        const D3 gd = load3( ddi, site.idx );
        if( p.site_flags )
        {
            g.x += moment * gd.x;
            g.y += moment * gd.y;
            g.z += moment * gd.z;
        }
        else
        {
            g.x += gd.x;
            g.y += gd.y;
            g.z += gd.z;
        }
    }
    if( p.site_flags )
        g = make_d3( present * g.x, present * g.y, present * g.z );
    out.bilinear = g;

    D3 r = make_d3( 0, 0, 0 );
    // Cubic anisotropy: g_c -= 2 K4 s_c^3   (Hamiltonian_Heisenberg.cpp:802-820)
    if( p.has_cubic )
    {
        const double k = 2.0 * p.K4[ib];
        r.x -= k * si.x * si.x * si.x;
        r.y -= k * si.y * si.y * si.y;
        r.z -= k * si.z * si.z * si.z;
    }
    // Zeeman: g -= mu_s B n   (Hamiltonian_Heisenberg.cpp:768-783)
    if( p.site_flags )
        r = make_d3( present * r.x, present * r.y, present * r.z );
    if( p.has_zeeman )
    {
        if( p.site_flags )
        {
            r.x -= moment * p.zeeman[ib][0];
            r.y -= moment * p.zeeman[ib][1];
            r.z -= moment * p.zeeman[ib][2];
        }
        else
        {
            r.x -= p.zeeman[ib][0];
            r.y -= p.zeeman[ib][1];
            r.z -= p.zeeman[ib][2];
        }
    }
    out.rest = r;
    return out;
}

// Energy of one site consistent with Gradient_and_Energy (Hamiltonian_Heisenberg.cpp:732-751):
//   1/2 g_bilinear . s  -  K4/2 (sx^4 + sy^4 + sz^4)  -  mu_s B n . s
template<int NB_T>
__device__ __forceinline__ double
site_energy( const StencilParams & p, const Site & site, const D3 & si, const SiteGradient & g )
{
    const int ib = NB_T == 1 ? 0 : site.ib;
    double e = 0.5 * dot3( g.bilinear, si );
    if( p.has_cubic )
    {
        const double x2 = si.x * si.x, y2 = si.y * si.y, z2 = si.z * si.z;
        e -= 0.5 * p.K4[ib] * ( x2 * x2 + y2 * y2 + z2 * z2 );
    }
    const unsigned flags = g.flags;
    if( p.has_zeeman && !( flags & ( FLAG_VACANT | FLAG_NO_MU_S ) ) )
        e -= p.zeeman[ib][0] * si.x + p.zeeman[ib][1] * si.y + p.zeeman[ib][2] * si.z;
    return ( flags & FLAG_VACANT ) ? 0.0 : e;
}

// force and virtual force of these sites are zero: pinned ones (mask_unpinned, Method_LLG.cpp:122-124, 222-224) and vacancies
// (their gradient is zero; the reference's dynamics divide their virtual force by mu_s = 0). Taken from the flags site_gradient
// loaded: a second look at p.site_flags after the gradient cost k_vp_a a factor 4.5 on the film (ptxas serialised the loads of
// the site behind it; profiles/r2zk, r2zl)
__device__ __forceinline__ bool site_frozen( const SiteGradient & g )
{
    return ( g.flags & ( FLAG_VACANT | FLAG_PINNED ) ) != 0u;
}

__device__ __forceinline__ D3 total( const SiteGradient & g )
{
    return make_d3( g.bilinear.x + g.rest.x, g.bilinear.y + g.rest.y, g.bilinear.z + g.rest.z );
}

} // namespace dev
} // namespace sb
