// Instantiations of the nearest-neighbour marching kernels (sc6.cuh) for one solver: 4 stages x 8 Hamiltonian
// structures. One translation unit per solver so that they compile in parallel.
#include "sc6.cuh"

namespace sb
{
namespace dev
{

void sc6_launch_rk4( int stage, const SC6Launch & L, cudaStream_t stream, const StencilParams & p, const LLGParams & l, const StageArgs & a )
{
    switch( stage )
    {
        case 1: sc6_launch_stage<Solver_RK4, 1>( L, stream, p, l, a ); break;
        case 2: sc6_launch_stage<Solver_RK4, 2>( L, stream, p, l, a ); break;
        case 3: sc6_launch_stage<Solver_RK4, 3>( L, stream, p, l, a ); break;
        case 4: sc6_launch_stage<Solver_RK4, 4>( L, stream, p, l, a ); break;
    }
}

} // namespace dev
} // namespace sb
