// C API: Chain_* and Transition_*. Reference behaviour: core/src/Spirit/Chain.cpp, Transitions.cpp,
// core/src/utility/Configuration_Chain.cpp:28-75.
#include "api_common.hpp"

#include <Spirit/Chain.h>
#include <Spirit/Simulation.h>
#include <Spirit/State.h>
#include <Spirit/Transitions.h>

#include <cmath>

using namespace sb;

namespace
{
// Locks only the chain's own mutex; the images that are added / removed are private to this thread
struct ChainListLock
{
    explicit ChainListLock( Chain & c ) : c_( c )
    {
        c_.mutex_.lock();
    }
    ~ChainListLock()
    {
        c_.mutex_.unlock();
    }
    Chain & c_;
};

void stop_chain_simulation( State * state, int idx_image, int idx_chain )
{
    if( Simulation_Running_On_Chain( state, idx_chain ) )
    {
        state->chain->iteration_allowed = false;
        Simulation_Stop( state, idx_image, idx_chain );
    }
}

std::shared_ptr<Spin_System> copy_of_clipboard( State * state )
{
    ImageLock lock( *state->clipboard_image );
    return std::make_shared<Spin_System>( *state->clipboard_image );
}

double angle( const Vec3 & a, const Vec3 & b )
{
    return std::acos( std::fmax( -1.0, std::fmin( 1.0, a.dot( b ) ) ) );
}

// Rodrigues rotation (Vectormath.cpp:456-461)
Vec3 rotate( const Vec3 & v, const Vec3 & axis, double ang )
{
    return v * std::cos( ang ) + axis.cross( v ) * std::sin( ang ) + axis * ( axis.dot( v ) * ( 1 - std::cos( ang ) ) );
}
} // namespace

int Chain_Get_NOI( State * state, int idx_chain ) noexcept
try
{
    return state->chain->noi;
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
    return 0;
}

bool Chain_next_Image( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( idx_image + 1 >= chain->noi )
        return false;
    ++chain->idx_active_image;
    state->idx_active_image = chain->idx_active_image;
    State_Update( state );
    return true;
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
    return false;
}

bool Chain_prev_Image( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( idx_image <= 0 )
        return false;
    --chain->idx_active_image;
    state->idx_active_image = chain->idx_active_image;
    State_Update( state );
    return true;
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
    return false;
}

bool Chain_Jump_To_Image( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto chain              = resolve( state, idx_image, idx_chain ).chain;
    chain->idx_active_image = idx_image;
    state->idx_active_image = idx_image;
    State_Update( state );
    return true;
}
SB_API_CATCH_RET( false )

// Chain.cpp:104-215
void Chain_Set_Length( State * state, int n_images, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( n_images < 1 )
    {
        Log( Log_Level::Warning, Log_Sender::API, "Tried to reduce length of chain below 1. No action taken.", -1, idx_chain );
        return;
    }
    if( n_images == chain->noi )
        return;
    stop_chain_simulation( state, idx_image, idx_chain );

    if( n_images > chain->noi )
    {
        if( !state->clipboard_image )
            Chain_Image_to_Clipboard( state, -1, idx_chain );
        while( chain->noi < n_images )
        {
            auto copy = copy_of_clipboard( state );
            ChainListLock lock( *chain );
            chain->noi++;
            chain->images.push_back( copy );
            chain->image_type.push_back( GNEB_Image_Type::Normal );
            state->method_image.push_back( nullptr );
        }
    }
    else
    {
        for( int img = chain->noi - 1; img > n_images - 1; --img )
        {
            Simulation_Stop( state, img, idx_chain );
            ChainListLock lock( *chain );
            chain->noi--;
            if( chain->idx_active_image == chain->noi )
            {
                --chain->idx_active_image;
                state->idx_active_image = chain->idx_active_image;
            }
            chain->images.pop_back();
            chain->image_type.pop_back();
            state->method_image.pop_back();
        }
    }
    State_Update( state );
    Chain_Setup_Data( state, idx_chain );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

void Chain_Image_to_Clipboard( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    state->clipboard_image = std::make_shared<Spin_System>( *image );
}
SB_API_CATCH_VOID

void Chain_Replace_Image( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto r = resolve( state, idx_image, idx_chain );
    if( !state->clipboard_image )
    {
        Log( Log_Level::Warning, Log_Sender::API, "Tried to replace image, but clipboard was empty.", idx_image, idx_chain );
        return;
    }
    stop_chain_simulation( state, idx_image, idx_chain );
    Simulation_Stop( state, idx_image, idx_chain );
    auto copy = copy_of_clipboard( state );
    {
        ChainListLock lock( *r.chain );
        r.chain->images[idx_image] = copy;
    }
    State_Update( state );
    Chain_Update_Data( state, idx_chain );
}
SB_API_CATCH_VOID

namespace
{
void insert_image( State * state, int position, int idx_image, int idx_chain )
{
    auto chain = state->chain;
    if( !state->clipboard_image )
    {
        Log( Log_Level::Warning, Log_Sender::API, "Tried to insert image, but clipboard was empty.", idx_image, idx_chain );
        return;
    }
    stop_chain_simulation( state, idx_image, idx_chain );
    auto copy = copy_of_clipboard( state );
    {
        ChainListLock lock( *chain );
        chain->noi++;
        chain->images.insert( chain->images.begin() + position, copy );
        chain->image_type.insert( chain->image_type.begin() + position, GNEB_Image_Type::Normal );
        state->method_image.insert( state->method_image.begin() + position, nullptr );
    }
    State_Update( state );
    Chain_Setup_Data( state, idx_chain );
}
} // namespace

void Chain_Insert_Image_Before( State * state, int idx_image, int idx_chain ) noexcept
try
{
    resolve( state, idx_image, idx_chain );
    insert_image( state, idx_image, idx_image, idx_chain );
    // the active image keeps pointing at the same configuration (Chain.cpp:330-333)
    if( state->clipboard_image && state->idx_active_image >= idx_image )
    {
        ++state->chain->idx_active_image;
        state->idx_active_image = state->chain->idx_active_image;
        State_Update( state );
    }
}
SB_API_CATCH_VOID

void Chain_Insert_Image_After( State * state, int idx_image, int idx_chain ) noexcept
try
{
    resolve( state, idx_image, idx_chain );
    insert_image( state, idx_image + 1, idx_image, idx_chain );
}
SB_API_CATCH_VOID

void Chain_Push_Back( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    insert_image( state, chain->noi, idx_image, idx_chain );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

bool Chain_Delete_Image( State * state, int idx_image, int idx_chain ) noexcept
try
{
    auto chain = resolve( state, idx_image, idx_chain ).chain;
    if( chain->noi <= 1 )
    {
        Log( Log_Level::Warning, Log_Sender::API, "Tried to delete last image.", idx_image, idx_chain );
        return false;
    }
    stop_chain_simulation( state, idx_image, idx_chain );
    Simulation_Stop( state, idx_image, idx_chain );
    {
        ChainListLock lock( *chain );
        chain->noi--;
        if( chain->idx_active_image == chain->noi )
        {
            --chain->idx_active_image;
            state->idx_active_image = chain->idx_active_image;
        }
        chain->images.erase( chain->images.begin() + idx_image );
        chain->image_type.erase( chain->image_type.begin() + idx_image );
        state->method_image.erase( state->method_image.begin() + idx_image );
    }
    State_Update( state );
    Chain_Setup_Data( state, idx_chain );
    return true;
}
SB_API_CATCH_RET( false )

bool Chain_Pop_Back( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    return Chain_Delete_Image( state, chain->noi - 1, idx_chain );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
    return false;
}

void Chain_Get_Rx( State * state, float * Rx, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    for( std::size_t i = 0; i < chain->Rx.size(); ++i )
        Rx[i] = float( chain->Rx[i] );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

void Chain_Get_Rx_Interpolated( State * state, float * Rx_interpolated, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    for( std::size_t i = 0; i < chain->Rx_interpolated.size(); ++i )
        Rx_interpolated[i] = float( chain->Rx_interpolated[i] );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

void Chain_Get_Energy( State * state, float * energy, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    for( int i = 0; i < chain->noi; ++i )
        energy[i] = float( chain->images[i]->E );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

void Chain_Get_Energy_Interpolated( State * state, float * E_interpolated, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    for( std::size_t i = 0; i < chain->E_interpolated.size(); ++i )
        E_interpolated[i] = float( chain->E_interpolated[i] );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

// Chain.cpp:692-727: energies of all images and the reaction coordinate (geodesic distances). The energies come from
// the GPU; the distance between two host-resident configurations is a plain host sum (setup path, not timed).
void Chain_Update_Data( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( int( chain->Rx.size() ) != chain->noi )
        chain->Rx.assign( chain->noi, 0.0 );
    for( int i = 0; i < chain->noi; ++i )
    {
        auto & image = *chain->images[i];
        ImageLock lock( image );
        image.UpdateEnergy();
        if( i > 0 )
        {
            const auto & prev = *chain->images[i - 1];
            double d2         = 0;
            for( int s = 0; s < image.nos; ++s )
            {
                const double a = angle( prev.spins[s], image.spins[s] );
                d2 += a * a;
            }
            chain->Rx[i] = chain->Rx[i - 1] + std::sqrt( d2 );
        }
    }
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

void Chain_Setup_Data( State * state, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    {
        ChainListLock lock( *chain );
        chain->Setup_Interpolation();
    }
    Chain_Update_Data( state, idx_chain );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

// ---------------------------------------------------------------------------------------------
// Transitions
// ---------------------------------------------------------------------------------------------
void Transition_Homogeneous( State * state, int idx_1, int idx_2, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( idx_2 <= idx_1 || idx_1 < 0 || idx_2 >= chain->noi )
    {
        Log( Log_Level::Error, Log_Sender::API,
             "Cannot set homogeneous transition between images " + std::to_string( idx_1 + 1 ) + " and " + std::to_string( idx_2 + 1 ), -1, idx_chain );
        return;
    }
    chain->Lock();
    try
    {
        // Configuration_Chain.cpp:28-75
        auto & s1 = chain->images[idx_1]->spins;
        auto & s2 = chain->images[idx_2]->spins;
        const Vec3 ex{ 1, 0, 0 }, ey{ 0, 1, 0 };
        bool antiparallel = false;
        for( int i = 0; i < chain->images[0]->nos; ++i )
        {
            const double rot_angle = angle( s1[i], s2[i] );
            Vec3 rot_axis          = s1[i].cross( s2[i] ).normalized();
            if( std::abs( rot_angle - constants::Pi ) < 1e-4 )
            {
                antiparallel = true;
                rot_axis     = ( std::abs( s1[i].dot( ex ) ) - 1 > 1e-4 ) ? ex : ey;
            }
            for( int img = idx_1 + 1; img < idx_2; ++img )
            {
                if( rot_angle > 1e-8 )
                    chain->images[img]->spins[i] = rotate( s1[i], rot_axis, rot_angle * double( img - idx_1 ) / double( idx_2 - idx_1 ) );
                else
                    chain->images[img]->spins[i] = s1[i];
            }
        }
        if( antiparallel )
            Log( Log_Level::Warning, Log_Sender::All, "For the interpolation of antiparallel spins an arbitrary rotation axis has been chosen." );
        for( int img = idx_1 + 1; img < idx_2; ++img ) // Transitions.cpp:42
            chain->images[img]->geometry->apply_pinning( chain->images[img]->spins.data() );
    }
    catch( ... )
    {
        handle_exception_api( __func__, -1, idx_chain );
    }
    chain->Unlock();
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

// Transitions.cpp:59-78
void Transition_Homogeneous_Insert_Interpolated( State * state, int n_interpolate, int idx_chain ) noexcept
try
{
    int noi = Chain_Get_NOI( state, idx_chain );
    if( n_interpolate == 0 || noi < 2 )
        return;
    for( int img = 0; img < noi - 1; ++img )
    {
        const int idx = img * ( n_interpolate + 1 );
        Chain_Image_to_Clipboard( state, idx, idx_chain );
        for( int k = 0; k < n_interpolate; ++k )
            Chain_Insert_Image_After( state, idx, idx_chain );
        Transition_Homogeneous( state, idx, idx + n_interpolate + 1, idx_chain );
    }
    Chain_Update_Data( state, idx_chain );
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}

// Transitions.cpp:80-123
void Transition_Add_Noise_Temperature( State * state, float temperature, int idx_1, int idx_2, int idx_chain ) noexcept
try
{
    int idx_image = -1;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    if( idx_2 <= idx_1 || idx_1 < 0 || idx_2 >= chain->noi )
    {
        Log( Log_Level::Error, Log_Sender::API, "Cannot add noise between these images", -1, idx_chain );
        return;
    }
    chain->Lock();
    auto all = []( const Vec3 &, const Vec3 & ) { return true; };
    for( int img = idx_1 + 1; img <= idx_2 - 1; ++img )
    {
        configurations::Add_Noise_Temperature( *chain->images[img], temperature, img, all );
        chain->images[img]->geometry->apply_pinning( chain->images[img]->spins.data() ); // Transitions.cpp:107
    }
    chain->Unlock();
}
catch( ... )
{
    handle_exception_api( __func__, -1, idx_chain );
}
