// Instantiations of the fused two-stage kernels (sc6_fused.cuh) for one solver: 4 Hamiltonian structures x 3 force modes,
// with and without the hook quantities. One translation unit per solver so that they compile in parallel.
#include "sc6_fused.cuh"

namespace sb
{
namespace dev
{

void sc6_fused_sib( bool hook, int spec, const FusedGeometry & G, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const FusedArgs & a )
{
    if( hook )
        sc6_fused_launch_solver<Solver_SIB, true>( spec, G, stream, p, l, a );
    else
        sc6_fused_launch_solver<Solver_SIB, false>( spec, G, stream, p, l, a );
}

} // namespace dev
} // namespace sb
