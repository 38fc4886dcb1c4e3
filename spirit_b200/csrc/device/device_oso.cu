// Minimisers on the device for one image: VP_OSO and LBFGS_OSO (orthogonal spin optimisation) and LBFGS_Atlas
// (stereographic atlas; Solver_LBFGS_Atlas.hpp:13-120, Solver_Kernels.cpp:105-246).
// Reference: Method_Solver<VP_OSO>::Iteration (core/include/engine/Solver_VP_OSO.hpp:34-115),
// Method_Solver<LBFGS_OSO>::Iteration (Solver_LBFGS_OSO.hpp:39-77), Solver_Kernels::oso_calc_gradients / oso_rotate /
// maximum_rotation (core/src/engine/Solver_Kernels.cpp:50-103), Solver_Kernels::lbfgs_get_searchdir
// (core/include/engine/Solver_Kernels.hpp:44-190).
//
// The reference runs one field sweep per BLAS-1 primitive (LBFGS: ~40 per iteration). Here an iteration is
//   1 stencil pass   force F = -grad E, virtual force (for the hook), OSO gradient g = T(-s x F), energy
//   VP_OSO           + 1 fused pass (velocity update with both dot products) + 1 fused pass (projection, search
//                    direction, rotation of the spins)
//   LBFGS_OSO        + the two-loop recursion, each of its 8 dot products fused with the update that precedes it, and
//                    1 fused pass (sign flip, memory of the gradient, norm for the step limit) + the rotation
// All element-wise passes run over the flat storage of the fields (AoSoA blocks; padding entries are zero and stay
// zero). The scalars of the recursion (rho, alpha, ...) are folded on the device in a fixed order and read back by the
// host, which owns the control flow of the recursion (restart on a non-positive curvature) exactly as the reference
// does. Single GPU only: on a slab decomposition the solvers refuse (the dot products would need an all-reduce each).
#include "device_buffers.cuh"

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
        if( so.z * a < tol )
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

struct OsoState
{
    static constexpr int MEM = 3; // n_lbfgs_memory (Solver_LBFGS_OSO.hpp:15)
    DeviceField grad, grad_pr, sd, q, vel;
    DeviceField da[MEM], dg[MEM];
    double rho[MEM] = { 0, 0, 0 }, alpha[MEM] = { 0, 0, 0 };
    int local_iter  = 0;
    int nblocks     = 0;
    double * partials = nullptr; // [2][nblocks]
    double * scalars  = nullptr; // device [4]
    double * h_scalars = nullptr;

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

void OsoStateDeleter::operator()( OsoState * p ) const
{
    delete p;
}

void DeviceImage::oso_reset()
{
    oso_.reset();
}

void DeviceImage::oso_iterate( int solver, LLGParams & llg, int n_iterations, bool hook, HookResult * result )
{
    if( solver != Solver_LBFGS_OSO && solver != Solver_VP_OSO && solver != Solver_LBFGS_Atlas )
        throw std::runtime_error( "spirit_b200: solver id " + std::to_string( solver ) + " is not an OSO / atlas solver" );
    if( slab_ )
        throw std::runtime_error( "spirit_b200: VP_OSO / LBFGS_OSO / LBFGS_Atlas are not implemented on a slab decomposition" );
    auto & b = *buf_;
    ensure_work_fields( Solver_VP );
    const bool lbfgs = solver != Solver_VP_OSO, atlas = solver == Solver_LBFGS_Atlas;
    const std::size_t n_sites = b.n_storage, n3 = 3 * b.n_storage;
    const std::size_t n       = atlas ? 2 * n_sites : n3; // doubles of a direction / gradient field
    const OsoLayout L{ n_sites, stencil_.plane_stride, b.plane_sites };
    const int M = OsoState::MEM;
    if( !oso_ )
    {
        oso_.reset( new OsoState );
        auto & o  = *oso_;
        o.nblocks = int( std::min<std::size_t>( OSO_BLOCKS_MAX, ( n_sites + BLOCK_THREADS - 1 ) / BLOCK_THREADS ) );
        SB_CUDA_CHECK( cudaMalloc( &o.partials, 3 * std::size_t( o.nblocks ) * sizeof( double ) ) );
        SB_CUDA_CHECK( cudaMalloc( &o.scalars, 4 * sizeof( double ) ) );
        SB_CUDA_CHECK( cudaHostAlloc( &o.h_scalars, 4 * sizeof( double ), cudaHostAllocDefault ) );
        auto zero = [&]( DeviceField & f ) {
            f.allocate( n_sites );
            SB_CUDA_CHECK( cudaMemsetAsync( f.base, 0, n3 * sizeof( double ), b.stream ) );
        };
        zero( o.grad );
        if( lbfgs )
        {
            zero( o.grad_pr );
            zero( o.sd );
            zero( o.q );
            for( int i = 0; i < M; ++i )
            {
                zero( o.da[i] );
                zero( o.dg[i] );
            }
        }
        if( !lbfgs || atlas )
            zero( o.vel ); // VP_OSO: velocity; atlas: chart a3 (first n_sites doubles) and the chart-change flag
        if( atlas )
        {
            k_atlas_init<<<o.nblocks, BLOCK_THREADS, 0, b.stream>>>( b.spins.c(), o.vel.base, L );
            ++launches_;
        }
    }
    auto & o     = *oso_;
    const int nb = o.nblocks;
    double * p0  = o.partials;
    double * p1  = o.partials + nb;
    // fold `count` partial arrays into scalars[0..count) and bring them to the host
    auto fetch = [&]( int count ) {
        for( int k = 0; k < count; ++k )
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( o.partials + std::size_t( k ) * nb, nb, o.scalars + k );
        SB_CUDA_CHECK( cudaMemcpyAsync( o.h_scalars, o.scalars, count * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
        launches_ += count;
    };
    const double epsilon    = 1e-300;                                        // Solver_Kernels.hpp:56
    const double maxmove    = atlas ? 0.05 : 3.14159265358979323846 / 200.0; // Solver_LBFGS_Atlas.hpp:30, Solver_LBFGS_OSO.hpp:31
    const double half_inv_m = 0.5 / 1.0;                                     // m = 1 (Method_Solver.hpp:174)
    double * a3 = o.vel.base;
    int * chart_flag = atlas ? reinterpret_cast<int *>( o.vel.base + n_sites ) : nullptr;

    // lbfgs_get_searchdir (Solver_Kernels.hpp:44-190) on the gradient in o.grad; leaves the direction in o.sd and returns
    // the factor that limits its root-mean-square length to maxmove (maximum_rotation, Solver_Kernels.cpp:95-103).
    // Each launch applies the pending update of one loop of the recursion and accumulates the next dot product.
    auto lbfgs_direction = [&]() -> double {
        double * g = o.grad.base, *g_pr = o.grad_pr.base, *sd = o.sd.base, *q = o.q.base;
        double sd_c         = 0; // the last update of the second loop is fused into k_lbfgs_finish
        const double * sd_x = nullptr;
        bool descent        = o.local_iter == 0;
        if( !descent )
        {
            const int m_index = o.local_iter % M;
            k_lbfgs_memorise<<<nb, BLOCK_THREADS, 0, b.stream>>>( o.da[m_index].base, o.dg[m_index].base, sd, g, g_pr, n, p0, p1 );
            ++launches_;
            fetch( 2 );
            const double rinv = o.h_scalars[0], dy2 = o.h_scalars[1];
            if( rinv > epsilon )
                o.rho[m_index] = 1.0 / rinv;
            else
            {
                o.local_iter = 0; // restart with a gradient-descent step (Solver_Kernels.hpp:108-114)
                descent      = true;
            }
            if( !descent )
            {
                // first loop: q = g; for k: alpha_c = rho_c (da_c . q); q -= alpha_c dg_c
                const double * src    = g;
                double c_prev         = 0;
                const double * x_prev = nullptr;
                for( int k = M - 1; k > -1; --k )
                {
                    const int c_ind = ( k + m_index + 1 ) % M;
                    k_lbfgs_axpy_dot<<<nb, BLOCK_THREADS, 0, b.stream>>>( q, src, c_prev, x_prev, o.da[c_ind].base, n, p0 );
                    ++launches_;
                    fetch( 1 );
                    o.alpha[c_ind] = o.rho[c_ind] * o.h_scalars[0];
                    src            = nullptr;
                    c_prev         = -o.alpha[c_ind];
                    x_prev         = o.dg[c_ind].base;
                }
                k_lbfgs_axpy_dot<<<nb, BLOCK_THREADS, 0, b.stream>>>( q, nullptr, c_prev, x_prev, nullptr, n, p0 );
                ++launches_;
                // sd = q / (rho_m dy2); second loop: for k: rhopdg = rho_c (dg_c . sd); sd += (alpha_c - rhopdg) da_c
                const double rhody2     = dy2 * o.rho[m_index];
                const double inv_rhody2 = rhody2 > epsilon ? 1.0 / rhody2 : 1.0 / epsilon;
                for( int k = 0; k < M; ++k )
                {
                    const int c_ind = o.local_iter < M ? k : ( k + m_index + 1 ) % M;
                    if( k == 0 )
                        k_lbfgs_scale_dot<<<nb, BLOCK_THREADS, 0, b.stream>>>( sd, inv_rhody2, q, o.dg[c_ind].base, n, p0 );
                    else
                        k_lbfgs_axpy_dot<<<nb, BLOCK_THREADS, 0, b.stream>>>( sd, nullptr, sd_c, sd_x, o.dg[c_ind].base, n, p0 );
                    ++launches_;
                    fetch( 1 );
                    const double rhopdg = o.rho[c_ind] * o.h_scalars[0];
                    sd_c                = o.alpha[c_ind] - rhopdg;
                    sd_x                = o.da[c_ind].base;
                }
            }
        }
        if( descent )
        {
            // Solver_Kernels.hpp:61-84: sd = -g, g_pr = g, memory cleared
            for( int i = 0; i < M; ++i )
            {
                o.rho[i] = 0;
                SB_CUDA_CHECK( cudaMemsetAsync( o.da[i].base, 0, n * sizeof( double ), b.stream ) );
                SB_CUDA_CHECK( cudaMemsetAsync( o.dg[i].base, 0, n * sizeof( double ), b.stream ) );
            }
        }
        k_lbfgs_finish<<<nb, BLOCK_THREADS, 0, b.stream>>>( sd, sd_c, sd_x, g, g_pr, descent ? 1 : 0, n, p0 );
        ++launches_;
        ++o.local_iter;
        fetch( 1 );
        const double rms = std::sqrt( o.h_scalars[0] / double( nos_ ) );
        return rms > maxmove ? maxmove / rms : 1.0;
    };

    for( int it = 0; it < n_iterations; ++it )
    {
        const bool hk = hook && ( it == n_iterations - 1 );
        // force, virtual force (LBFGS: s x F, VP_OSO: dt gamma/mu_B s x F -- `llg` carries the prefactor), energy
        compute_ddi_gradient( 0 );
        SB_DISPATCH_NB(
            ( k_force_and_virtual<1><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
                stencil_, b.lg, llg, b.spins.c(), b.ddi_s.c(), b.F.f(), b.Fv.f(), b.partials ) ),
            ( k_force_and_virtual<0><<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>(
                stencil_, b.lg, llg, b.spins.c(), b.ddi_s.c(), b.F.f(), b.Fv.f(), b.partials ) ) );
        if( hk )
        {
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials, b.nblocks, b.scalars + 4 );
            ++launches_;
        }
        // Fv = dtg (s x F)  ->  s x F = Fv / dtg
        const double inv_dtg = 1.0 / llg.dtg;
        if( !lbfgs )
        {
            k_oso_gradient<true><<<nb, BLOCK_THREADS, 0, b.stream>>>( b.Fv.c(), o.grad.f(), o.vel.f(), L, inv_dtg, half_inv_m, p0, p1 );
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( p0, nb, o.scalars );
            k_reduce_sum<<<1, BLOCK_THREADS, 0, b.stream>>>( p1, nb, o.scalars + 1 );
            k_vp_oso_update<<<nb, BLOCK_THREADS, 0, b.stream>>>( b.spins.f(), o.grad.c(), o.vel.f(), L, o.scalars, llg.dt, half_inv_m );
            launches_ += 4;
        }
        else if( !atlas )
        {
            k_oso_gradient<false><<<nb, BLOCK_THREADS, 0, b.stream>>>( b.Fv.c(), o.grad.f(), o.vel.f(), L, -inv_dtg, 0.0, p0, p1 );
            ++launches_;
            const double scaling = lbfgs_direction();
            k_oso_rotate<<<nb, BLOCK_THREADS, 0, b.stream>>>( b.spins.f(), o.sd.f(), L, scaling );
            ++launches_;
        }
        else
        {
            k_atlas_gradient<<<nb, BLOCK_THREADS, 0, b.stream>>>( b.spins.c(), b.F.c(), a3, o.grad.base, L );
            ++launches_;
            const double scaling = lbfgs_direction();
            SB_CUDA_CHECK( cudaMemsetAsync( chart_flag, 0, sizeof( int ), b.stream ) );
            k_atlas_rotate<<<nb, BLOCK_THREADS, 0, b.stream>>>( b.spins.f(), a3, o.sd.base, L, scaling, -0.6, chart_flag );
            ++launches_;
            int flag = 0;
            SB_CUDA_CHECK( cudaMemcpyAsync( &flag, chart_flag, sizeof( int ), cudaMemcpyDeviceToHost, b.stream ) );
            SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
            if( flag )
            {
                AtlasMemory mem;
                for( int i = 0; i < M; ++i )
                {
                    mem.upd[i]  = o.da[i].base;
                    mem.gupd[i] = o.dg[i].base;
                }
                k_atlas_transform<<<nb, BLOCK_THREADS, 0, b.stream>>>( b.spins.c(), a3, o.sd.base, o.grad_pr.base, mem, L, o.partials, nb );
                ++launches_;
                fetch( 3 );
                for( int i = 0; i < M; ++i )
                    o.rho[i] = 1.0 / ( 1.0 / o.rho[i] + o.h_scalars[i] ); // Solver_Kernels.cpp:195-198,241-244
            }
        }
        if( hk )
        {
            k_hook<<<b.nblocks, BLOCK_THREADS, 0, b.stream>>>( stencil_, b.lg, b.spins.c(), b.F.f(), b.Fv.c(), b.partials + b.nblocks );
            k_reduce_max<<<1, BLOCK_THREADS, 0, b.stream>>>( b.partials + b.nblocks, b.nblocks, b.scalars + 5 );
            launches_ += 2;
        }
        ++llg.iteration;
    }
    effective_field_in_Fv_ = false;
    SB_CUDA_CHECK( cudaGetLastError() );
    if( hook )
    {
        SB_CUDA_CHECK( cudaMemcpyAsync( b.h_scalars + 4, b.scalars + 4, 2 * sizeof( double ), cudaMemcpyDeviceToHost, b.stream ) );
        SB_CUDA_CHECK( cudaStreamSynchronize( b.stream ) );
        if( result )
        {
            result->energy     = b.h_scalars[4];
            result->max_torque = std::sqrt( b.h_scalars[5] );
        }
    }
}

} // namespace dev
} // namespace sb
