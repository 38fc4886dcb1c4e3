// Instantiations of the nearest-neighbour marching kernels (sc6.cuh) for one solver: 2 stages x 8 Hamiltonian
// structures. One translation unit per solver so that they compile in parallel.
#include "sc6.cuh"

namespace sb
{
namespace dev
{

void sc6_launch_depondt( int stage, const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a )
{
    switch( stage )
    {
        case 1: sc6_launch_stage<Solver_Depondt, 1>( L, stream, p, l, a ); break;
        case 2: sc6_launch_stage<Solver_Depondt, 2>( L, stream, p, l, a ); break;
    }
}

} // namespace dev
} // namespace sb
