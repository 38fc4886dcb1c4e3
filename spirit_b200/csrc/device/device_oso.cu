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
// does. On a slab decomposition every one of these scalars is summed over the ranks before the host reads it.
#include "oso.cuh"

namespace sb
{
namespace dev
{

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
    auto & b = *buf_;
    ensure_work_fields( Solver_VP );
    // Slab of a lattice: the element-wise passes run over the OWNED planes (the halo planes in front of them are skipped by
    // offsetting the field views: whole planes are whole AoSoA blocks), every dot product is summed over the ranks
    // (Solver_LBFGS_OSO.hpp:39-77 takes them over the whole field) and the rotated spins refresh the neighbours' halo planes.
    const std::size_t skip = slab_ ? 3 * std::size_t( stencil_.halo ) * stencil_.plane_stride : 0;
    const std::size_t n_owned = slab_ ? std::size_t( stencil_.plane_stride ) * stencil_.nc_local : b.n_storage;
    const OsoLayout L{ n_owned, stencil_.plane_stride, b.plane_sites };
    const Field3 S{ b.spins.base + skip };
    const int nos_lattice = slab_ ? int( std::size_t( nos_ ) / stencil_.nc_local * stencil_.Nc ) : nos_;
    if( !oso_ )
    {
        oso_.reset( new OsoState );
        oso_->distributed = slab_ && comm_world() > 1;
        oso_->allocate( solver, n_owned, 1, ConstField3{ S.base }, L, b.stream, launches_ );
    }
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
        oso_update(
            *oso_, solver, S, ConstField3{ b.Fv.base + skip }, 1.0 / llg.dtg, ConstField3{ b.F.base + skip }, L, nos_lattice, llg.dt, b.stream,
            launches_ );
        if( slab_ )
            exchange_halo( &b.spins );
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
        allreduce_scalars( 4, 1, false ); // energy
        allreduce_scalars( 5, 1, true );  // max torque^2
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
