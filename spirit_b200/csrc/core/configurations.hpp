// Host-side initialisers of spin configurations (not on the timed path; they build the inputs).
// Behaviour follows Utility::Configurations (core/src/utility/Configurations.cpp:73-330) and the
// position filter of the API layer (core/src/Spirit/Configurations.cpp:15-63).
#pragma once

#include "state.hpp"

#include <functional>

namespace sb
{
namespace configurations
{

using filterfunction = std::function<bool( const Vec3 & spin, const Vec3 & position )>;

filterfunction get_filter(
    const Vec3 & position, const float r_cut_rectangular[3], float r_cut_cylindrical, float r_cut_spherical, bool inverted );

void Domain( Spin_System & s, Vec3 direction, const filterfunction & filter );
void Random( Spin_System & s, const filterfunction & filter );
void Add_Noise_Temperature( Spin_System & s, double temperature, int delta_seed, const filterfunction & filter );
void Skyrmion(
    Spin_System & s, Vec3 pos, double r, double order, double phase, bool upDown, bool achiral, bool rl,
    const filterfunction & filter );
void DW_Skyrmion(
    Spin_System & s, Vec3 pos, double dw_radius, double dw_width, double order, double phase, bool upDown, bool achiral,
    bool rl, const filterfunction & filter );
void Hopfion( Spin_System & s, Vec3 pos, double r, int order, Vec3 normal, const filterfunction & filter );
void SpinSpiral( Spin_System & s, const std::string & direction_type, Vec3 q, Vec3 axis, double theta, const filterfunction & filter );
void Insert( Spin_System & s, const std::vector<Vec3> & configuration, int shift, const filterfunction & filter );

// One unit vector, uniform on the sphere, from two draws (Vectormath.cpp:41-52)
Vec3 random_unit_vector( std::mt19937 & prng );

} // namespace configurations
} // namespace sb
