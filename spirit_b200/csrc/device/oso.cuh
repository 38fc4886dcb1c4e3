// Device building blocks of the OSO / atlas minimisers, shared by the single-image path (device_oso.cu) and the chain
// path (device_chain.cu). See device_oso.cu for the overview.
#pragma once

#include "device_buffers.cuh"

#include <algorithm>
#include <cmath>

namespace sb
{
namespace dev
{

namespace
{
constexpr int OSO_BLOCKS_MAX = 4096;

// Planes are padded to a multiple of 32 storage sites; the padding entries of the work fields (F, Fv) are never
// written, so site-wise passes must not read them: they keep the padding of everything they write at zero.
struct OsoLayout
{
    std::size_t n_sites; // storage sites
    int plane_stride, plane_sites;
};
__device__ __forceinline__ bool oso_real_site( const OsoLayout & L, std::size_t i )
{
    return int( i % std::size_t( L.plane_stride ) ) < L.plane_sites;
}

// g = T v with T = [[0,0,1],[0,-1,0],[1,0,0]] (Solver_Kernels.cpp:52): (v.z, -v.y, v.x)
__device__ __forceinline__ D3 oso_t( const D3 & v )
{
    return make_d3( v.z, -v.y, v.x );
}

// OSO gradient of every site from the virtual force Fv = scale_fv * (s x F): g = sign * T(s x F).
// LBFGS_OSO: g = T(-s x F) (sign -1); VP_OSO: g = -T(-s x F) = T(s x F) (sign +1, Solver_VP_OSO.hpp:68-70).
// VP_OSO also advances the velocity, v += (g_prev + g) / 2m, and accumulates v.g and g.g (Solver_VP_OSO.hpp:73-87).
template<bool VP>
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_oso_gradient(
    ConstField3 Fv, Field3 grad, Field3 vel, const OsoLayout L, double factor, double half_inv_m, double * __restrict__ p_vg,
    double * __restrict__ p_gg )
{
    double vg = 0, gg = 0;
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue; // grad and vel stay zero there
        const D3 f = load3( Fv, i );
        const D3 g = oso_t( make_d3( factor * f.x, factor * f.y, factor * f.z ) );
        if( VP )
        {
            const D3 gp = load3( grad, i );
            D3 v        = load3( vel, i );
            v           = make_d3( v.x + half_inv_m * ( gp.x + g.x ), v.y + half_inv_m * ( gp.y + g.y ), v.z + half_inv_m * ( gp.z + g.z ) );
            store3( vel, i, v );
            vg += dot3( v, g );
            gg += dot3( g, g );
        }
        store3( grad, i, g );
    }
    if( VP )
    {
        vg = block_sum( vg );
        if( threadIdx.x == 0 )
            p_vg[blockIdx.x] = vg;
        gg = block_sum( gg );
        if( threadIdx.x == 0 )
            p_gg[blockIdx.x] = gg;
    }
}

// Rotation of every spin by its search direction (oso_rotate, Solver_Kernels.cpp:62-93): theta = |sd|, axis -sd/theta.
__device__ __forceinline__ D3 oso_rotated( const D3 & s, const D3 & sd )
{
    const double theta = sqrt( dot3( sd, sd ) );
    if( !( theta > 1.0e-20 ) )
        return s;
    double sn, q;
    sincos( theta, &sn, &q );
    const double w = 1 - q, x = -sd.x / theta, y = -sd.y / theta, z = -sd.z / theta;
    const double s1 = -y * z * w, s2 = x * z * w, s3 = -x * y * w, p1 = x * sn, p2 = y * sn, p3 = z * sn;
    return make_d3(
        ( q + z * z * w ) * s.x + ( s1 + p1 ) * s.y + ( s2 + p2 ) * s.z, ( s1 - p1 ) * s.x + ( q + y * y * w ) * s.y + ( s3 + p3 ) * s.z,
        ( s2 - p2 ) * s.x + ( s3 - p3 ) * s.y + ( q + x * x * w ) * s.z );
}

// VP_OSO, second half (Solver_VP_OSO.hpp:93-113): v = g * ratio (or 0), sd = dt v + dt g / 2m, rotate.
// scalars[0] = v.g, scalars[1] = g.g (all sites)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_vp_oso_update(
    Field3 s, ConstField3 grad, Field3 vel, const OsoLayout L, const double * __restrict__ scalars, double dt, double half_inv_m )
{
    const double proj = scalars[0], ratio = proj / scalars[1];
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue;
        const D3 g = load3( grad, i );
        D3 v       = make_d3( 0, 0, 0 );
        if( proj > 0 )
            v = make_d3( g.x * ratio, g.y * ratio, g.z * ratio );
        store3( vel, i, v );
        const D3 sd = make_d3( dt * v.x + half_inv_m * dt * g.x, dt * v.y + half_inv_m * dt * g.y, dt * v.z + half_inv_m * dt * g.z );
        store3( s, i, oso_rotated( load3( s, i ), sd ) );
    }
}

// sd *= scaling (the reference scales the stored search direction, Solver_LBFGS_OSO.hpp:66-69: it is the next
// iteration's delta_a), then rotate
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_oso_rotate( Field3 s, Field3 sd, const OsoLayout L, double scaling )
{
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue;
        D3 d = load3( sd, i );
        if( scaling != 1.0 )
        {
            d = make_d3( scaling * d.x, scaling * d.y, scaling * d.z );
            store3( sd, i, d );
        }
        store3( s, i, oso_rotated( load3( s, i ), d ) );
    }
}

// ---- stereographic atlas (LBFGS_Atlas): two-component fields are flat [2][storage sites] ----------------------------
// chart of every spin: a3 = sign(s_z) (Solver_LBFGS_Atlas.hpp:34-41)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_atlas_init( ConstField3 s, double * __restrict__ a3, const OsoLayout L )
{
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
        a3[i] = ( oso_real_site( L, i ) && !( load3( s, i ).z > 0 ) ) ? -1.0 : 1.0;
}
// atlas_calc_gradients (Solver_Kernels.cpp:128-153)
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_atlas_gradient( ConstField3 s, ConstField3 F, const double * __restrict__ a3, double * __restrict__ resid, const OsoLayout L )
{
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue;
        const D3 si = load3( s, i ), f = load3( F, i );
        const double a = a3[i];
        const double J00 = si.y * si.y + si.z * ( si.z + a ), J01 = -si.x * si.y, J11 = si.x * si.x + si.z * ( si.z + a );
        const double J02 = -si.x * ( si.z + a ), J12 = -si.y * ( si.z + a );
        resid[i]             = -( J00 * f.x + J01 * f.y + J02 * f.z );
        resid[L.n_sites + i] = -( J01 * f.x + J11 * f.y + J12 * f.z );
    }
}
// dirs *= scaling, atlas_rotate (Solver_Kernels.cpp:105-126), ncg_atlas_check_coordinates (:155-184): *flag = 1 if any
// spin has left the trusted part of its chart (s_z a3 < tol)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_atlas_rotate(
    Field3 s, const double * __restrict__ a3, double * __restrict__ dirs, const OsoLayout L, double scaling, double tol, int * __restrict__ flag )
{
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue;
        const double d0 = scaling * dirs[i], d1 = scaling * dirs[L.n_sites + i];
        dirs[i]             = d0;
        dirs[L.n_sites + i] = d1;
        const D3 si         = load3( s, i );
        const double a      = a3[i];
        const double gamma  = 1 + si.z * a;
        const double denom  = ( si.x * si.x + si.y * si.y ) / gamma + 2 * ( d0 * si.x + d1 * si.y ) + gamma * ( d0 * d0 + d1 * d1 );
        const double inv    = 1 / ( gamma + denom );
        const D3 so         = make_d3( 2 * ( si.x + d0 * gamma ) * inv, 2 * ( si.y + d1 * gamma ) * inv, a * ( gamma - denom ) * inv );
        store3( s, i, so );
        if( flag && so.z * a < tol )
            *flag = 1;
    }
}
// lbfgs_atlas_transform_direction (Solver_Kernels.cpp:186-246): spins in the wrong half of their chart change chart;
// direction, previous residual and the L-BFGS memory are rescaled, and 1/rho_n changes by sum (factor^2 - 1) a_n . g_n,
// accumulated here as a deterministic reduction (the reference updates rho inside its parallel loop).
struct AtlasMemory
{
    double * upd[3];
    double * gupd[3];
};
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_atlas_transform(
    ConstField3 s, double * __restrict__ a3, double * __restrict__ dirs, double * __restrict__ g_pr, const AtlasMemory m, const OsoLayout L,
    double * __restrict__ partials, int nblocks )
{
    double acc[3] = { 0, 0, 0 };
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < L.n_sites; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue;
        const double sz = load3( s, i ).z;
        if( sz * a3[i] < 0 )
        {
            const double a      = sz > 0 ? 1.0 : -1.0;
            a3[i]               = a;
            const double factor = ( 1 - a * sz ) / ( 1 + a * sz );
            const std::size_t j = L.n_sites + i;
            dirs[i] *= factor;
            dirs[j] *= factor;
            g_pr[i] *= factor;
            g_pr[j] *= factor;
#pragma unroll
            for( int n = 0; n < 3; ++n )
            {
                const double a0 = m.upd[n][i], a1 = m.upd[n][j], g0 = m.gupd[n][i], g1 = m.gupd[n][j];
                acc[n] += ( factor * factor - 1 ) * ( a0 * g0 + a1 * g1 );
                m.upd[n][i]  = a0 * factor;
                m.upd[n][j]  = a1 * factor;
                m.gupd[n][i] = g0 * factor;
                m.gupd[n][j] = g1 * factor;
            }
        }
    }
#pragma unroll
    for( int n = 0; n < 3; ++n )
    {
        const double v = block_sum( acc[n] );
        if( threadIdx.x == 0 )
            partials[n * nblocks + blockIdx.x] = v;
    }
}

// ---- flat element-wise passes of the L-BFGS recursion (n = 3 * storage sites doubles) ------------------------------
// da = sd, dg = g - g_pr; partial sums of dg.da and dg.dg   (Solver_Kernels.hpp:86-103,134-136)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_lbfgs_memorise(
    double * __restrict__ da, double * __restrict__ dg, const double * __restrict__ sd, const double * __restrict__ g,
    const double * __restrict__ g_pr, std::size_t n, double * __restrict__ p_dgda, double * __restrict__ p_dgdg )
{
    double a = 0, b = 0;
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < n; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        const double x = sd[i], y = g[i] - g_pr[i];
        da[i] = x;
        dg[i] = y;
        a += y * x;
        b += y * y;
    }
    a = block_sum( a );
    if( threadIdx.x == 0 )
        p_dgda[blockIdx.x] = a;
    b = block_sum( b );
    if( threadIdx.x == 0 )
        p_dgdg[blockIdx.x] = b;
}
// y = (first ? src : y) + c * x ; partial sum of z.y  (z may be null). One step of either loop of the recursion.
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_lbfgs_axpy_dot(
    double * __restrict__ y, const double * __restrict__ src, double c, const double * __restrict__ x, const double * __restrict__ z,
    std::size_t n, double * __restrict__ partials )
{
    double a = 0;
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < n; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        double v = src ? src[i] : y[i];
        if( x )
            v += c * x[i];
        y[i] = v;
        if( z )
            a += z[i] * v;
    }
    if( z )
    {
        a = block_sum( a );
        if( threadIdx.x == 0 )
            partials[blockIdx.x] = a;
    }
}
// y = c * x ; partial sum of z.y
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_lbfgs_scale_dot(
    double * __restrict__ y, double c, const double * __restrict__ x, const double * __restrict__ z, std::size_t n, double * __restrict__ partials )
{
    double a = 0;
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < n; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        const double v = c * x[i];
        y[i]           = v;
        a += z[i] * v;
    }
    a = block_sum( a );
    if( threadIdx.x == 0 )
        partials[blockIdx.x] = a;
}
// sd = (first ? -g : -(sd + c * x)), g_pr = g; partial sum of sd.sd   (Solver_Kernels.hpp:66-70,175-186 + maximum_rotation)
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_lbfgs_finish(
    double * __restrict__ sd, double c, const double * __restrict__ x, const double * __restrict__ g, double * __restrict__ g_pr,
    int gradient_descent, std::size_t n, double * __restrict__ partials )
{
    double a = 0;
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < n; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        const double gi = g[i];
        double v;
        if( gradient_descent )
            v = -gi;
        else
            v = -( sd[i] + c * x[i] );
        sd[i]   = v;
        g_pr[i] = gi;
        a += v * v;
    }
    a = block_sum( a );
    if( threadIdx.x == 0 )
        partials[blockIdx.x] = a;
}
} // namespace

// L-BFGS two-loop recursion (lbfgs_get_searchdir, core/include/engine/Solver_Kernels.hpp:44-190) over flat device
// arrays of n doubles (one image, or all images of a chain back to back: the reference sums its dot products over the
// images). The scalars are folded on the device in a fixed order and read back by the host, which owns the control
// flow (restart on a non-positive curvature) exactly as the reference does. Each launch applies the pending update of
// one loop of the recursion and accumulates the next dot product.
struct LbfgsEngine
{
    static constexpr int MEM = 3; // n_lbfgs_memory (Solver_LBFGS_OSO.hpp:15, Solver_LBFGS_Atlas.hpp:15)
    double *g = nullptr, *g_pr = nullptr, *sd = nullptr, *q = nullptr;
    double * da[MEM] = { nullptr, nullptr, nullptr };
    double * dg[MEM] = { nullptr, nullptr, nullptr };
    std::size_t n    = 0;
    int nb           = 0;
    double * partials = nullptr; // [>= 3][nb]
    double * scalars = nullptr, *h_scalars = nullptr; // device [4], pinned host [4]
    cudaStream_t stream       = nullptr;
    std::uint64_t * launches  = nullptr;
    double rho[MEM] = { 0, 0, 0 }, alpha[MEM] = { 0, 0, 0 };
    int local_iter  = 0;
    bool distributed = false; // the fields are one rank's part (slab / image shard): every dot product is summed over the ranks

    // fold `count` partial arrays into scalars[0..count) and bring them to the host
    void fetch( int count )
    {
        for( int k = 0; k < count; ++k )
            k_reduce_sum<<<1, BLOCK_THREADS, 0, stream>>>( partials + std::size_t( k ) * nb, nb, scalars + k );
        if( distributed )
            comm_allreduce( scalars, count, false, stream );
        SB_CUDA_CHECK( cudaMemcpyAsync( h_scalars, scalars, count * sizeof( double ), cudaMemcpyDeviceToHost, stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        *launches += count;
    }

    // New search direction in sd from the gradient in g; returns sum sd.sd
    double direction()
    {
        const double epsilon = 1e-300; // Solver_Kernels.hpp:56
        double * p0 = partials, *p1 = partials + nb;
        double sd_c         = 0; // the last update of the second loop is fused into k_lbfgs_finish
        const double * sd_x = nullptr;
        bool descent        = local_iter == 0;
        if( !descent )
        {
            const int m_index = local_iter % MEM;
            k_lbfgs_memorise<<<nb, BLOCK_THREADS, 0, stream>>>( da[m_index], dg[m_index], sd, g, g_pr, n, p0, p1 );
            ++*launches;
            fetch( 2 );
            const double rinv = h_scalars[0], dy2 = h_scalars[1];
            if( rinv > epsilon )
                rho[m_index] = 1.0 / rinv;
            else
            {
                local_iter = 0; // restart with a gradient-descent step (Solver_Kernels.hpp:108-114)
                descent    = true;
            }
            if( !descent )
            {
                // first loop: q = g; for k: alpha_c = rho_c (da_c . q); q -= alpha_c dg_c
                const double * src    = g;
                double c_prev         = 0;
                const double * x_prev = nullptr;
                for( int k = MEM - 1; k > -1; --k )
                {
                    const int c_ind = ( k + m_index + 1 ) % MEM;
                    k_lbfgs_axpy_dot<<<nb, BLOCK_THREADS, 0, stream>>>( q, src, c_prev, x_prev, da[c_ind], n, p0 );
                    ++*launches;
                    fetch( 1 );
                    alpha[c_ind] = rho[c_ind] * h_scalars[0];
                    src          = nullptr;
                    c_prev       = -alpha[c_ind];
                    x_prev       = dg[c_ind];
                }
                k_lbfgs_axpy_dot<<<nb, BLOCK_THREADS, 0, stream>>>( q, nullptr, c_prev, x_prev, nullptr, n, p0 );
                ++*launches;
                // sd = q / (rho_m dy2); second loop: for k: rhopdg = rho_c (dg_c . sd); sd += (alpha_c - rhopdg) da_c
                const double rhody2     = dy2 * rho[m_index];
                const double inv_rhody2 = rhody2 > epsilon ? 1.0 / rhody2 : 1.0 / epsilon;
                for( int k = 0; k < MEM; ++k )
                {
                    const int c_ind = local_iter < MEM ? k : ( k + m_index + 1 ) % MEM;
                    if( k == 0 )
                        k_lbfgs_scale_dot<<<nb, BLOCK_THREADS, 0, stream>>>( sd, inv_rhody2, q, dg[c_ind], n, p0 );
                    else
                        k_lbfgs_axpy_dot<<<nb, BLOCK_THREADS, 0, stream>>>( sd, nullptr, sd_c, sd_x, dg[c_ind], n, p0 );
                    ++*launches;
                    fetch( 1 );
                    const double rhopdg = rho[c_ind] * h_scalars[0];
                    sd_c                = alpha[c_ind] - rhopdg;
                    sd_x                = da[c_ind];
                }
            }
        }
        if( descent )
        {
            // Solver_Kernels.hpp:61-84: sd = -g, g_pr = g, memory cleared
            for( int i = 0; i < MEM; ++i )
            {
                rho[i] = 0;
                SB_CUDA_CHECK( cudaMemsetAsync( da[i], 0, n * sizeof( double ), stream ) );
                SB_CUDA_CHECK( cudaMemsetAsync( dg[i], 0, n * sizeof( double ), stream ) );
            }
        }
        k_lbfgs_finish<<<nb, BLOCK_THREADS, 0, stream>>>( sd, sd_c, sd_x, g, g_pr, descent ? 1 : 0, n, p0 );
        ++*launches;
        ++local_iter;
        fetch( 1 );
        return h_scalars[0];
    }
};

// Fields and scalars of the minimisers for `n_images` images stored back to back (n_sites storage sites in total)
struct OsoState
{
    static constexpr int MEM = LbfgsEngine::MEM;
    DeviceField grad, grad_pr, sd, q, vel;
    DeviceField da[MEM], dg[MEM];
    LbfgsEngine lbfgs;
    int nblocks        = 0;
    int n_images       = 1;
    bool distributed   = false; // slab of a lattice / shard of a chain: scalars are reduced over the ranks (set before allocate)
    double * partials  = nullptr; // [max(3, n_images)][nblocks]
    double * scalars   = nullptr; // device [max(4, n_images)]
    double * h_scalars = nullptr;

    // solver: Solver_VP_OSO / Solver_LBFGS_OSO / Solver_LBFGS_Atlas. `spins`: the images' configurations (atlas charts)
    void allocate( int solver, std::size_t n_sites, int n_images_, ConstField3 spins, const OsoLayout & L, cudaStream_t stream, std::uint64_t & launches )
    {
        const bool lbfgs_solver = solver != Solver_VP_OSO, atlas = solver == Solver_LBFGS_Atlas;
        n_images = n_images_;
        nblocks  = int( std::min<std::size_t>( OSO_BLOCKS_MAX, ( n_sites + BLOCK_THREADS - 1 ) / BLOCK_THREADS ) );
        const int rows = std::max( 3, n_images ), n_scal = std::max( 4, n_images );
        SB_CUDA_CHECK( cudaMalloc( &partials, std::size_t( rows ) * nblocks * sizeof( double ) ) );
        SB_CUDA_CHECK( cudaMalloc( &scalars, n_scal * sizeof( double ) ) );
        SB_CUDA_CHECK( cudaHostAlloc( &h_scalars, n_scal * sizeof( double ), cudaHostAllocDefault ) );
        auto zero = [&]( DeviceField & f ) {
            f.allocate( n_sites );
            SB_CUDA_CHECK( cudaMemsetAsync( f.base, 0, 3 * n_sites * sizeof( double ), stream ) );
        };
        zero( grad );
        if( lbfgs_solver )
        {
            zero( grad_pr );
            zero( sd );
            zero( q );
            for( int i = 0; i < MEM; ++i )
            {
                zero( da[i] );
                zero( dg[i] );
            }
            lbfgs.g = grad.base, lbfgs.g_pr = grad_pr.base, lbfgs.sd = sd.base, lbfgs.q = q.base;
            for( int i = 0; i < MEM; ++i )
                lbfgs.da[i] = da[i].base, lbfgs.dg[i] = dg[i].base;
            lbfgs.n         = ( atlas ? 2 : 3 ) * n_sites;
            lbfgs.nb        = nblocks;
            lbfgs.partials  = partials;
            lbfgs.scalars   = scalars;
            lbfgs.h_scalars = h_scalars;
            lbfgs.stream    = stream;
            lbfgs.launches  = &launches;
            lbfgs.distributed = distributed;
        }
        if( !lbfgs_solver || atlas )
            zero( vel ); // VP_OSO: velocity; atlas: charts a3 (first n_sites doubles) and the chart-change flag behind them
        if( atlas )
        {
            k_atlas_init<<<nblocks, BLOCK_THREADS, 0, stream>>>( spins, vel.base, L );
            ++launches;
        }
    }

    ~OsoState()
    {
        for( DeviceField * f : { &grad, &grad_pr, &sd, &q, &vel } )
            f->release();
        for( int i = 0; i < MEM; ++i )
        {
            da[i].release();
            dg[i].release();
        }
        if( partials )
            cudaFree( partials );
        if( scalars )
            cudaFree( scalars );
        if( h_scalars )
            cudaFreeHost( h_scalars );
    }
};

// sum of x^2 per image: image img owns n_segs segments of seg_len doubles at x + s * seg_stride + img * seg_len
static __global__ void __launch_bounds__( BLOCK_THREADS ) k_image_sumsq(
    const double * __restrict__ x, std::size_t seg_len, int n_segs, std::size_t seg_stride, double * __restrict__ partials, int nblocks )
{
    const int img = blockIdx.y;
    double a      = 0;
    for( int sgm = 0; sgm < n_segs; ++sgm )
    {
        const double * q = x + std::size_t( sgm ) * seg_stride + std::size_t( img ) * seg_len;
        for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < seg_len; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
            a += q[i] * q[i];
    }
    a = block_sum( a );
    if( threadIdx.x == 0 )
        partials[std::size_t( img ) * nblocks + blockIdx.x] = a;
}
// ncg_atlas_check_coordinates over a chain as the reference evaluates it (Solver_Kernels.cpp:155-184: the spins of
// image 0 against the charts of every image)
static __global__ void __launch_bounds__( BLOCK_THREADS )
    k_atlas_check_chain( ConstField3 s0, const double * __restrict__ a3, std::size_t sites_per_image, int n_images, const OsoLayout L, double tol, int * __restrict__ flag )
{
    for( std::size_t i = blockIdx.x * std::size_t( BLOCK_THREADS ) + threadIdx.x; i < sites_per_image; i += std::size_t( gridDim.x ) * BLOCK_THREADS )
    {
        if( !oso_real_site( L, i ) )
            continue;
        const double sz = load3( s0, i ).z;
        for( int img = 0; img < n_images; ++img )
            if( sz * a3[std::size_t( img ) * sites_per_image + i] < tol )
                *flag = 1;
    }
}

// One update of the configurations from the force: `C` = c_factor^-1 * (s x F) of every site (c_factor = the prefactor
// the virtual force carries), `F` the force itself (atlas). Gradient in the solver's coordinates, search direction,
// step limit (per image: the smallest factor of all images, Solver_LBFGS_OSO.hpp:59-69 / Solver_LBFGS_Atlas.hpp:79-97),
// rotation.
inline void oso_update(
    OsoState & o, int solver, Field3 S, ConstField3 C, double inv_c_factor, ConstField3 F, const OsoLayout & L, int nos_per_image,
    double dt, cudaStream_t stream, std::uint64_t & launches )
{
    const bool lbfgs = solver != Solver_VP_OSO, atlas = solver == Solver_LBFGS_Atlas;
    const int nb     = o.nblocks;
    double * p0 = o.partials, *p1 = o.partials + nb;
    const double half_inv_m = 0.5 / 1.0;                                     // m = 1 (Method_Solver.hpp:174)
    const double maxmove    = atlas ? 0.05 : 3.14159265358979323846 / 200.0; // Solver_LBFGS_Atlas.hpp:30, Solver_LBFGS_OSO.hpp:31
    const std::size_t sites_per_image = L.n_sites / std::size_t( o.n_images );
    if( !lbfgs )
    {
        k_oso_gradient<true><<<nb, BLOCK_THREADS, 0, stream>>>( C, o.grad.f(), o.vel.f(), L, inv_c_factor, half_inv_m, p0, p1 );
        k_reduce_sum<<<1, BLOCK_THREADS, 0, stream>>>( p0, nb, o.scalars );
        k_reduce_sum<<<1, BLOCK_THREADS, 0, stream>>>( p1, nb, o.scalars + 1 );
        if( o.distributed )
            comm_allreduce( o.scalars, 2, false, stream ); // the projection is taken over the whole lattice / chain
        k_vp_oso_update<<<nb, BLOCK_THREADS, 0, stream>>>( S, o.grad.c(), o.vel.f(), L, o.scalars, dt, half_inv_m );
        launches += 4;
        return;
    }
    double * a3      = o.vel.base;
    int * chart_flag = atlas ? reinterpret_cast<int *>( o.vel.base + L.n_sites ) : nullptr;
    if( atlas )
        k_atlas_gradient<<<nb, BLOCK_THREADS, 0, stream>>>( ConstField3{ S.base }, F, a3, o.grad.base, L );
    else
        k_oso_gradient<false><<<nb, BLOCK_THREADS, 0, stream>>>( C, o.grad.f(), o.vel.f(), L, -inv_c_factor, 0.0, p0, p1 );
    ++launches;
    double sumsq = o.lbfgs.direction();
    if( o.n_images > 1 )
    {
        // the largest root-mean-square step of any image decides
        const int nbi = std::max( 1, nb / o.n_images );
        if( atlas )
            k_image_sumsq<<<dim3( nbi, o.n_images ), BLOCK_THREADS, 0, stream>>>( o.sd.base, sites_per_image, 2, L.n_sites, o.partials, nbi );
        else
            k_image_sumsq<<<dim3( nbi, o.n_images ), BLOCK_THREADS, 0, stream>>>( o.sd.base, 3 * sites_per_image, 1, 0, o.partials, nbi );
        for( int i = 0; i < o.n_images; ++i )
            k_reduce_sum<<<1, BLOCK_THREADS, 0, stream>>>( o.partials + std::size_t( i ) * nbi, nbi, o.scalars + i );
        SB_CUDA_CHECK( cudaMemcpyAsync( o.h_scalars, o.scalars, o.n_images * sizeof( double ), cudaMemcpyDeviceToHost, stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        launches += 1 + o.n_images;
        sumsq = 0;
        for( int i = 0; i < o.n_images; ++i )
            sumsq = std::max( sumsq, o.h_scalars[i] );
        if( o.distributed )
        {
            // images sharded over ranks: the largest step of ANY image of the chain
            SB_CUDA_CHECK( cudaMemcpyAsync( o.scalars, &sumsq, sizeof( double ), cudaMemcpyHostToDevice, stream ) );
            comm_allreduce( o.scalars, 1, true, stream );
            SB_CUDA_CHECK( cudaMemcpyAsync( &sumsq, o.scalars, sizeof( double ), cudaMemcpyDeviceToHost, stream ) );
            SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        }
    }
    const double rms     = std::sqrt( sumsq / double( nos_per_image ) );
    const double scaling = rms > maxmove ? maxmove / rms : 1.0;
    if( !atlas )
    {
        k_oso_rotate<<<nb, BLOCK_THREADS, 0, stream>>>( S, o.sd.f(), L, scaling );
        ++launches;
        return;
    }
    SB_CUDA_CHECK( cudaMemsetAsync( chart_flag, 0, sizeof( int ), stream ) );
    k_atlas_rotate<<<nb, BLOCK_THREADS, 0, stream>>>( S, a3, o.sd.base, L, scaling, -0.6, o.n_images == 1 ? chart_flag : nullptr );
    ++launches;
    if( o.n_images > 1 )
    {
        OsoLayout L1 = L;
        L1.n_sites   = sites_per_image;
        k_atlas_check_chain<<<nb, BLOCK_THREADS, 0, stream>>>( ConstField3{ S.base }, a3, sites_per_image, o.n_images, L1, -0.6, chart_flag );
        ++launches;
    }
    int flag = 0;
    SB_CUDA_CHECK( cudaMemcpyAsync( &flag, chart_flag, sizeof( int ), cudaMemcpyDeviceToHost, stream ) );
    SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
    if( o.distributed )
    {
        // a chart change anywhere changes the charts everywhere (Solver_Kernels.cpp:155-184 decides on the whole field)
        double any = flag ? 1.0 : 0.0;
        SB_CUDA_CHECK( cudaMemcpyAsync( o.scalars, &any, sizeof( double ), cudaMemcpyHostToDevice, stream ) );
        comm_allreduce( o.scalars, 1, true, stream );
        SB_CUDA_CHECK( cudaMemcpyAsync( &any, o.scalars, sizeof( double ), cudaMemcpyDeviceToHost, stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( stream ) );
        flag = any != 0.0;
    }
    if( flag )
    {
        AtlasMemory mem;
        for( int i = 0; i < OsoState::MEM; ++i )
        {
            mem.upd[i]  = o.da[i].base;
            mem.gupd[i] = o.dg[i].base;
        }
        k_atlas_transform<<<nb, BLOCK_THREADS, 0, stream>>>( ConstField3{ S.base }, a3, o.sd.base, o.grad_pr.base, mem, L, o.partials, nb );
        ++launches;
        o.lbfgs.fetch( 3 );
        for( int i = 0; i < OsoState::MEM; ++i )
            o.lbfgs.rho[i] = 1.0 / ( 1.0 / o.lbfgs.rho[i] + o.h_scalars[i] ); // Solver_Kernels.cpp:195-198,241-244
    }
}

} // namespace dev
} // namespace sb
