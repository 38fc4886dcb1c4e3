#include "state.hpp"

#include <cstring>
#include <stdexcept>

namespace sb
{

// ---------------------------------------------------------------------------------------------
HostField::HostField( std::size_t n )
{
    resize( n );
}

HostField::HostField( const HostField & other )
{
    resize( other.n_ );
    if( n_ )
        std::memcpy( data_, other.data_, n_ * sizeof( Vec3 ) );
}

HostField & HostField::operator=( const HostField & other )
{
    if( this != &other )
    {
        if( n_ != other.n_ )
            resize( other.n_ );
        if( n_ )
            std::memcpy( data_, other.data_, n_ * sizeof( Vec3 ) );
    }
    return *this;
}

HostField::~HostField()
{
    dev::host_free( data_, pinned_ );
}

void HostField::resize( std::size_t n )
{
    dev::host_free( data_, pinned_ );
    data_ = nullptr;
    n_    = n;
    if( n )
    {
        data_ = static_cast<Vec3 *>( dev::host_alloc( n * sizeof( Vec3 ), pinned_ ) );
        std::memset( static_cast<void *>( data_ ), 0, n * sizeof( Vec3 ) );
    }
}

// ---------------------------------------------------------------------------------------------
Spin_System::Spin_System(
    std::shared_ptr<Hamiltonian> hamiltonian_, std::shared_ptr<Geometry> geometry_, std::shared_ptr<Parameters_LLG> llg )
        : nos( geometry_->nos ),
          spins( geometry_->nos ),
          effective_field( geometry_->nos ),
          hamiltonian( std::move( hamiltonian_ ) ),
          geometry( std::move( geometry_ ) ),
          llg_parameters( std::move( llg ) )
{
}

Spin_System::Spin_System( const Spin_System & other )
        : nos( other.nos ),
          spins( other.spins ),
          effective_field( ( const_cast<Spin_System &>( other ).refresh_effective_field_mirror(), other.effective_field ) ),
          E( other.E ),
          E_array( other.E_array ),
          M( other.M )
{
    geometry              = std::make_shared<Geometry>( *other.geometry );
    hamiltonian           = std::make_shared<Hamiltonian>( *other.hamiltonian );
    hamiltonian->geometry = geometry;
    llg_parameters        = std::make_shared<Parameters_LLG>( *other.llg_parameters );
    iteration_allowed     = false;
    singleshot_allowed    = false;
}

dev::DeviceImage & Spin_System::device()
{
    if( !device_ )
        device_ = std::make_unique<dev::DeviceImage>( *geometry );
    return *device_;
}

void Spin_System::sync_to_device()
{
    auto & d = device();
    d.set_hamiltonian( *hamiltonian );
    if( !device_is_newer )
        d.upload_spins( spins.scalars() );
}

// Spin_System.cpp:115-129: per-term energies and their sum
void Spin_System::UpdateEnergy()
{
    sync_to_device();
    std::vector<double> totals( hamiltonian->contribution_names.size(), 0.0 );
    device().energy_contributions( *hamiltonian, totals.data(), nullptr );
    E_array.clear();
    double sum = 0;
    for( std::size_t t = 0; t < totals.size(); ++t )
    {
        E_array.emplace_back( hamiltonian->contribution_names[t], totals[t] );
        sum += totals[t];
    }
    E = sum;
}

// Spin_System.cpp:131-141: effective field = -gradient
void Spin_System::UpdateEffectiveField()
{
    sync_to_device();
    device().update_effective_field();
    device().download_effective_field( effective_field.scalars() );
    effective_field_stale = false;
}

void Spin_System::refresh_effective_field_mirror()
{
    if( !effective_field_stale )
        return;
    effective_field_stale = false;
    if( device_ )
        device_->download_effective_field( effective_field.scalars() );
}

// ---------------------------------------------------------------------------------------------
void from_indices(
    const State * state, int & idx_image, int & idx_chain, std::shared_ptr<Spin_System> & image,
    std::shared_ptr<Chain> & chain )
{
    if( state == nullptr )
        throw std::runtime_error( "The State pointer is invalid" );
    if( state->chain == nullptr )
        throw std::runtime_error( "The State seems to not be initialised correctly" );
    idx_chain = 0;
    chain     = state->chain;
    if( idx_image >= state->chain->noi )
        throw std::out_of_range(
            "Index " + std::to_string( idx_image ) + " points to non-existent image (NOI="
            + std::to_string( state->chain->noi ) + "). No action taken." );
    if( idx_image < 0 )
    {
        image     = state->active_image;
        idx_image = state->idx_active_image;
    }
    else
        image = chain->images[idx_image];
}

} // namespace sb
