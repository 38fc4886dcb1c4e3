#include "configurations.hpp"
#include "constants.hpp"
#include "logging.hpp"

#include <algorithm>
#include <cmath>

namespace sb
{
namespace configurations
{

using constants::Pi;

// core/src/Spirit/Configurations.cpp:15-63: a negative cut-off disables that criterion
filterfunction get_filter(
    const Vec3 & position, const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical, bool inverted )
{
    const double rx = r_cut_rectangular[0], ry = r_cut_rectangular[1], rz = r_cut_rectangular[2];
    const double rc = r_cut_cylindrical, rs = r_cut_spherical;
    return [=]( const Vec3 &, const Vec3 & p ) {
        const Vec3 d         = p - position;
        const double r_cyl   = std::sqrt( d.x * d.x + d.y * d.y );
        const double r_sph   = d.norm();
        const bool inside    = ( rx < 0 || std::abs( d.x ) < rx ) && ( ry < 0 || std::abs( d.y ) < ry )
                            && ( rz < 0 || std::abs( d.z ) < rz ) && ( rc < 0 || r_cyl < rc ) && ( rs < 0 || r_sph < rs );
        return inverted ? !inside : inside;
    };
}

Vec3 random_unit_vector( std::mt19937 & prng )
{
    std::uniform_real_distribution<double> distribution( -1, 1 );
    const double v_z  = distribution( prng );
    const double phi  = distribution( prng ) * Pi;
    const double r_xy = std::sqrt( 1 - v_z * v_z );
    return Vec3{ r_xy * std::cos( phi ), r_xy * std::sin( phi ), v_z };
}

void Domain( Spin_System & s, Vec3 v, const filterfunction & filter )
{
    if( v.norm() < 1e-8 )
    {
        Log( Log_Level::Warning, Log_Sender::All, "Homogeneous vector was zero and got set to (0, 0, 1)" );
        v = { 0, 0, 1 };
    }
    else
        v.normalize();
    const auto & positions = s.geometry->positions();
    for( int i = 0; i < s.nos; ++i )
        if( filter( s.spins[i], positions[i] ) )
            s.spins[i] = v;
}

void Random( Spin_System & s, const filterfunction & filter )
{
    const auto & positions = s.geometry->positions();
    for( int i = 0; i < s.nos; ++i )
        if( filter( s.spins[i], positions[i] ) )
            s.spins[i] = random_unit_vector( s.llg_parameters->prng );
}

// Configurations.cpp:131-162: xi = sqrt(T k_B) * random unit vectors, added on the filtered sites, then all normalised
void Add_Noise_Temperature( Spin_System & s, double temperature, int delta_seed, const filterfunction & filter )
{
    if( temperature == 0.0 )
        return;
    const auto & positions = s.geometry->positions();
    const double epsilon   = std::sqrt( temperature * constants::k_B );
    std::mt19937 local( 123456789 + delta_seed );
    std::mt19937 & prng = delta_seed != 0 ? local : s.llg_parameters->prng;
    for( int i = 0; i < s.nos; ++i )
    {
        // the reference draws a vector for every site and masks afterwards
        const Vec3 xi = random_unit_vector( prng ) * epsilon;
        if( filter( s.spins[i], positions[i] ) )
            s.spins[i] += xi;
    }
    for( int i = 0; i < s.nos; ++i )
        s.spins[i].normalize();
}

void Skyrmion(
    Spin_System & s, Vec3 pos, double r, double order, double phase, bool upDown, bool achiral, bool rl,
    const filterfunction & filter )
{
    const auto & positions = s.geometry->positions();
    const int ksi = int( rl ) * 2 - 1, dir = int( upDown ) * 2 - 1;
    for( int i = 0; i < s.nos; ++i )
    {
        const double dx = positions[i].x - pos.x, dy = positions[i].y - pos.y;
        const double distance = std::sqrt( dx * dx + dy * dy ) / r;
        if( filter( s.spins[i], positions[i] ) )
        {
            const double x = dx / distance / r;
            double phi_i   = std::acos( std::max( -1.0, std::min( 1.0, x ) ) );
            if( distance == 0 )
                phi_i = 0;
            if( dy < 0.0 )
                phi_i = -phi_i;
            phi_i += phase / 180 * Pi;
            const double theta_i = Pi - Pi * distance;
            s.spins[i].x         = ksi * std::sin( theta_i ) * std::cos( order * phi_i );
            s.spins[i].y         = ksi * std::sin( theta_i ) * std::sin( order * ( phi_i + achiral * Pi ) );
            s.spins[i].z         = std::cos( theta_i ) * -dir;
        }
    }
    for( int i = 0; i < s.nos; ++i )
        s.spins[i].normalize();
}

void DW_Skyrmion(
    Spin_System & s, Vec3 pos, double dw_radius, double dw_width, double order, double phase, bool upDown, bool achiral,
    bool rl, const filterfunction & filter )
{
    const auto & positions = s.geometry->positions();
    const int ksi = int( rl ) * 2 - 1, dir = int( upDown ) * 2 - 1;
    for( int i = 0; i < s.nos; ++i )
    {
        const double dx = positions[i].x - pos.x, dy = positions[i].y - pos.y;
        const double distance = std::sqrt( dx * dx + dy * dy );
        if( filter( s.spins[i], positions[i] ) )
        {
            const double theta_i = std::asin( std::tanh( -2 * ( distance + dw_radius ) / dw_width ) )
                                   + std::asin( std::tanh( -2 * ( distance - dw_radius ) / dw_width ) ) + Pi;
            const double x = dx / distance;
            double phi_i   = std::acos( std::max( -1.0, std::min( 1.0, x ) ) );
            if( distance == 0 )
                phi_i = 0;
            if( dy < 0.0 )
                phi_i = -phi_i;
            phi_i += phase / 180 * Pi;
            s.spins[i].x = ksi * std::sin( theta_i ) * std::cos( order * phi_i );
            s.spins[i].y = ksi * std::sin( theta_i ) * std::sin( order * phi_i + achiral * Pi );
            s.spins[i].z = std::cos( theta_i ) * -dir;
        }
    }
    for( int i = 0; i < s.nos; ++i )
        s.spins[i].normalize();
}

// Configurations.cpp:164-229 (toroidal hopfion about `normal`)
void Hopfion( Spin_System & s, Vec3 pos, double r, int order, Vec3 normal, const filterfunction & filter )
{
    if( r == 0.0 )
        return;
    // Frame with `normal` as z axis (Vectormath::dreibein): rows ex, ey, ez
    normal.normalize();
    Vec3 ex, ey, ez = normal;
    {
        const Vec3 zaxis{ 0, 0, 1 };
        if( std::abs( ez.z ) > 1 - 1e-12 )
        {
            ex = { 1, 0, 0 };
            ey = ez.z > 0 ? Vec3{ 0, 1, 0 } : Vec3{ 0, -1, 0 };
        }
        else
        {
            ex = zaxis.cross( ez ).normalized();
            ey = ez.cross( ex );
        }
    }
    const auto & positions = s.geometry->positions();
    for( int n = 0; n < s.nos; ++n )
    {
        // position in the rotated frame, about pos
        const Vec3 dp = positions[n] - pos;
        const Vec3 p  = Vec3{ ex.dot( dp ), ey.dot( dp ), ez.dot( dp ) } + pos;
        if( !filter( s.spins[n], p ) )
            continue;
        const double d = ( p - pos ).norm();
        double T       = d == 0 ? 0 : ( p.z - pos.z ) / d;
        T              = std::acos( T );
        double t       = d / r;
        t              = 1.0 + 4.22 / ( t * t );
        const double tmp = Pi * ( 1.0 - 1.0 / std::sqrt( t ) );
        t                = std::sin( tmp ) * std::sin( T );
        t                = std::acos( 1.0 - 2.0 * t * t );
        const double F   = std::atan2( p.y - pos.y, p.x - pos.x );
        double f         = F + std::atan( 1.0 / ( std::tan( tmp ) * std::cos( T ) ) );
        if( !( T > Pi / 2.0 ) )
            f += Pi;
        s.spins[n] = Vec3{ std::sin( t ) * std::cos( order * f ), std::sin( t ) * std::sin( order * f ), std::cos( t ) };
    }
}

// Configurations.cpp:314-429
void SpinSpiral( Spin_System & s, const std::string & direction_type, Vec3 q, Vec3 axis, double theta, const filterfunction & filter )
{
    const Vec3 vx{ 1, 0, 0 }, vy{ 0, 1, 0 }, vz{ 0, 0, 1 };
    const Vec3 a1 = s.geometry->bravais_vectors[0], a2 = s.geometry->bravais_vectors[1], a3 = s.geometry->bravais_vectors[2];
    axis.normalize();
    Vec3 e1, e2;
    if( axis.z == 0 )
    {
        e2 = axis.cross( vz );
        e1 = vz;
    }
    else if( axis.z > 0 )
    {
        e1 = vx;
        e2 = vy;
    }
    else
    {
        e1 = vx;
        e2 = -vy;
    }
    theta           = theta / 180.0 * Pi;
    const Vec3 v1   = ( e1 - e1.dot( axis ) * axis ).normalized();
    const Vec3 v2   = ( e2 - e2.dot( axis ) * axis - e2.dot( v1 ) * v1 ).normalized();
    if( direction_type == "Reciprocal Lattice" )
    {
        const Vec3 b1 = ( 2.0 * Pi / a1.dot( a2.cross( a3 ) ) ) * a2.cross( a3 );
        const Vec3 b2 = ( 2.0 * Pi / a2.dot( a3.cross( a1 ) ) ) * a3.cross( a1 );
        const Vec3 b3 = ( 2.0 * Pi / a3.dot( a1.cross( a2 ) ) ) * a1.cross( a2 );
        q             = q.x * b1 + q.y * b2 + q.z * b3;
    }
    else if( direction_type == "Real Lattice" )
        q = Vec3{ q.dot( a1 ), q.dot( a2 ), q.dot( a3 ) };
    else if( direction_type != "Real Space" )
        Log( Log_Level::Warning, Log_Sender::All, "Got passed invalid type for SS: " + direction_type );
    const auto & positions = s.geometry->positions();
    for( int i = 0; i < s.nos; ++i )
        if( filter( s.spins[i], positions[i] ) )
        {
            const double phase = positions[i].dot( q );
            s.spins[i] = axis * std::cos( theta ) + v1 * ( std::cos( phase ) * std::sin( theta ) ) + v2 * ( std::sin( phase ) * std::sin( theta ) );
            s.spins[i].normalize();
        }
}

void Insert( Spin_System & s, const std::vector<Vec3> & configuration, int shift, const filterfunction & filter )
{
    const int nos = s.nos;
    if( shift < 0 )
        shift += nos;
    if( std::size_t( nos ) != configuration.size() )
    {
        Log( Log_Level::Warning, Log_Sender::All, "Tried to insert spin configuration with NOS != NOS_system" );
        return;
    }
    const auto & positions = s.geometry->positions();
    for( int i = 0; i < nos; ++i )
        if( filter( s.spins[i], positions[i] ) )
            s.spins[i] = configuration[( i + shift ) % nos];
}

} // namespace configurations
} // namespace sb
