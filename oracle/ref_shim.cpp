// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Engine-level probes into the UNMODIFIED reference (linked into oracle/_ref/libSpirit_ref.so
// by oracle/Makefile). The reference's C API returns energies / torques as `float`
// (core/include/Spirit/System.h:55, Simulation.h:58-71), so 1e-12 parity checks need the
// double-precision engine objects -- reached here exactly as the reference's own tests do
// (core/test/test_anisotropy.cpp:137-149, core/test/test_physics.cpp:13,110-130 include
// <data/State.hpp> and call state->active_image->hamiltonian->Gradient_and_Energy()).
//
// Every probe has a twin with the same signature in the product library
// (include/spirit_b200.h, prefix SpiritB200_) so a parity test can call both uniformly.

#include <Spirit/State.h>
#include <data/Spin_System.hpp>
#include <data/Spin_System_Chain.hpp>
#include <data/State.hpp>
#include <engine/Hamiltonian_Heisenberg.hpp>
#include <engine/Manifoldmath.hpp>
#include <engine/Vectormath.hpp>

#include <cstring>
#include <memory>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{

std::shared_ptr<Data::Spin_System> image_of( State * state, int idx_image )
{
    std::shared_ptr<Data::Spin_System> image;
    std::shared_ptr<Data::Spin_System_Chain> chain;
    int idx_chain = -1;
    from_indices( state, idx_image, idx_chain, image, chain );
    return image;
}

// Either the image's own spins or a caller-provided AoS [nos][3] array
vectorfield spins_from( const Data::Spin_System & image, const double * spins )
{
    if( !spins )
        return *image.spins;
    vectorfield vf( image.nos );
    std::memcpy( vf.data(), spins, sizeof( double ) * 3 * image.nos );
    return vf;
}

} // namespace

extern "C"
{

int refshim_version()
{
    return 2;
}

int refshim_num_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void refshim_set_num_threads( int n )
{
#ifdef _OPENMP
    omp_set_num_threads( n );
#endif
}

// gradient[nos][3], energy: hamiltonian->Gradient_and_Energy (Hamiltonian_Heisenberg.cpp:704-766)
int refshim_Gradient_and_Energy( State * state, const double * spins, double * gradient, double * energy, int idx_image )
try
{
    auto image = image_of( state, idx_image );
    auto s     = spins_from( *image, spins );
    vectorfield g( image->nos, Vector3::Zero() );
    scalar E = 0;
    image->hamiltonian->Gradient_and_Energy( s, g, E );
    std::memcpy( gradient, g.data(), sizeof( double ) * 3 * image->nos );
    *energy = E;
    return image->nos;
}
catch( ... )
{
    return -1;
}

// hamiltonian->Gradient (Hamiltonian_Heisenberg.cpp:670-702)
int refshim_Gradient( State * state, const double * spins, double * gradient, int idx_image )
try
{
    auto image = image_of( state, idx_image );
    auto s     = spins_from( *image, spins );
    vectorfield g( image->nos, Vector3::Zero() );
    image->hamiltonian->Gradient( s, g );
    std::memcpy( gradient, g.data(), sizeof( double ) * 3 * image->nos );
    return image->nos;
}
catch( ... )
{
    return -1;
}

// Energy_Contributions_per_Spin (Hamiltonian_Heisenberg.cpp:262-305). names: [max_terms][32] chars,
// totals[max_terms], per_spin (nullable) [n_terms][nos]. Returns the number of terms.
int refshim_Energy_Contributions(
    State * state, const double * spins, int max_terms, char * names, double * totals, double * per_spin, int idx_image )
try
{
    auto image = image_of( state, idx_image );
    auto s     = spins_from( *image, spins );
    std::vector<std::pair<std::string, scalarfield>> contributions;
    image->hamiltonian->Energy_Contributions_per_Spin( s, contributions );
    int n = std::min<int>( max_terms, contributions.size() );
    for( int t = 0; t < n; ++t )
    {
        std::strncpy( names + 32 * t, contributions[t].first.c_str(), 31 );
        names[32 * t + 31] = 0;
        double sum         = 0;
        for( auto e : contributions[t].second )
            sum += e;
        totals[t] = sum;
        if( per_spin )
            std::memcpy( per_spin + std::size_t( t ) * image->nos, contributions[t].second.data(), sizeof( double ) * image->nos );
    }
    return n;
}
catch( ... )
{
    return -1;
}

// image->E in double (System_Get_Energy narrows to float, System.h:55)
double refshim_Get_Energy( State * state, int idx_image )
try
{
    return image_of( state, idx_image )->E;
}
catch( ... )
{
    return 0;
}

// Pair lists as the engine holds them after Update_Interactions (Hamiltonian_Heisenberg.cpp:101-198).
// kind 0: exchange, 1: DMI. ijt: [max][5] = i, j, da, db, dc; magnitudes[max]; normals[max][3] (DMI only).
int refshim_Get_Pairs( State * state, int kind, int max_pairs, int * ijt, double * magnitudes, double * normals, int idx_image )
try
{
    auto image = image_of( state, idx_image );
    auto * ham = dynamic_cast<Engine::Hamiltonian_Heisenberg *>( image->hamiltonian.get() );
    if( !ham )
        return -1;
    const auto & pairs = kind == 0 ? ham->exchange_pairs : ham->dmi_pairs;
    const auto & mags  = kind == 0 ? ham->exchange_magnitudes : ham->dmi_magnitudes;
    int n              = pairs.size();
    for( int p = 0; p < n && p < max_pairs; ++p )
    {
        ijt[5 * p + 0] = pairs[p].i;
        ijt[5 * p + 1] = pairs[p].j;
        for( int d = 0; d < 3; ++d )
            ijt[5 * p + 2 + d] = pairs[p].translations[d];
        magnitudes[p] = mags[p];
        if( kind == 1 && normals )
            for( int d = 0; d < 3; ++d )
                normals[3 * p + d] = ham->dmi_normals[p][d];
    }
    return n;
}
catch( ... )
{
    return -1;
}

// Maximum torque of the running / last method on the image (or chain if idx_image == -2), in double
double refshim_Get_MaxTorque( State * state, int idx_image )
try
{
    if( idx_image == -2 )
        return state->method_chain ? state->method_chain->getTorqueMaxNorm() : 0;
    std::shared_ptr<Data::Spin_System> image;
    std::shared_ptr<Data::Spin_System_Chain> chain;
    int idx_chain = -1;
    from_indices( state, idx_image, idx_chain, image, chain );
    auto & m = state->method_image[idx_image];
    return m ? m->getTorqueMaxNorm() : 0;
}
catch( ... )
{
    return 0;
}

// Chain reaction coordinate and energies in double (Chain_Get_Rx / Chain_Get_Energy are float, Chain.h)
int refshim_Chain_Get_Rx_E( State * state, double * Rx, double * E )
try
{
    check_state( state );
    int noi = state->chain->noi;
    for( int i = 0; i < noi; ++i )
    {
        Rx[i] = state->chain->Rx[i];
        E[i]  = state->chain->images[i]->E;
    }
    return noi;
}
catch( ... )
{
    return -1;
}

// Magnetization in double (Quantity_Get_Magnetization is float, Quantities.h:21)
int refshim_Get_Magnetization( State * state, double * m, int idx_image )
try
{
    auto image = image_of( state, idx_image );
    auto M     = Engine::Vectormath::Magnetization( *image->spins, image->geometry->mu_s );
    m[0]       = M[0];
    m[1]       = M[1];
    m[2]       = M[2];
    return 0;
}
catch( ... )
{
    return -1;
}

} // extern "C"
