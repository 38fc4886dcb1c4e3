// Basic host-side value types of the spirit_b200 host library.
//
// The reference stores fields as std::vector<Eigen::Vector3d> (AoS, 24 B per spin;
// core/include/engine/Vectormath_Defines.hpp:30-108). The host side of this library keeps the
// same AoS layout for everything that crosses the C API (System_Get_Spin_Directions returns a
// live `scalar*` of shape [nos][3]); the device side is SoA (see device/device_image.hpp).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

typedef double scalar; // Spirit_Defines.h: `typedef SPIRIT_SCALAR_TYPE scalar;` -- this build is double only

namespace sb
{

struct Vec3
{
    double x = 0, y = 0, z = 0;

    double & operator[]( int i )
    {
        return ( &x )[i];
    }
    const double & operator[]( int i ) const
    {
        return ( &x )[i];
    }
    Vec3 operator+( const Vec3 & o ) const
    {
        return { x + o.x, y + o.y, z + o.z };
    }
    Vec3 operator-( const Vec3 & o ) const
    {
        return { x - o.x, y - o.y, z - o.z };
    }
    Vec3 operator-() const
    {
        return { -x, -y, -z };
    }
    Vec3 operator*( double c ) const
    {
        return { x * c, y * c, z * c };
    }
    Vec3 operator/( double c ) const
    {
        return { x / c, y / c, z / c };
    }
    Vec3 & operator+=( const Vec3 & o )
    {
        x += o.x;
        y += o.y;
        z += o.z;
        return *this;
    }
    Vec3 & operator-=( const Vec3 & o )
    {
        x -= o.x;
        y -= o.y;
        z -= o.z;
        return *this;
    }
    double dot( const Vec3 & o ) const
    {
        return x * o.x + y * o.y + z * o.z;
    }
    Vec3 cross( const Vec3 & o ) const
    {
        return { y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x };
    }
    double squaredNorm() const
    {
        return x * x + y * y + z * z;
    }
    double norm() const
    {
        return std::sqrt( squaredNorm() );
    }
    // Eigen 3.3.90 semantics (core/thirdparty/Eigen/src/Core/Dot.h:145-151): a zero vector is left untouched
    void normalize()
    {
        double z2 = squaredNorm();
        if( z2 > 0 )
        {
            double n = std::sqrt( z2 );
            x /= n;
            y /= n;
            z /= n;
        }
    }
    Vec3 normalized() const
    {
        Vec3 v = *this;
        v.normalize();
        return v;
    }
};
inline Vec3 operator*( double c, const Vec3 & v )
{
    return v * c;
}

static_assert( sizeof( Vec3 ) == 24, "Vec3 must be layout-compatible with scalar[3]" );

// Interaction pair: basis indices i, j and the cell translation of j (Vectormath_Defines.hpp:85-89)
struct Pair
{
    int i = 0, j = 0;
    std::array<int, 3> translations{ 0, 0, 0 };
};

using intfield    = std::vector<int>;
using scalarfield = std::vector<double>;
using vectorfield = std::vector<Vec3>;
using pairfield   = std::vector<Pair>;

} // namespace sb
