#include "method_gneb.hpp"
#include "logging.hpp"

#include <cmath>
#include <stdexcept>

namespace sb
{

// core/src/utility/Cubic_Hermite_Spline.cpp:11-48
std::vector<std::vector<double>> cubic_hermite_interpolate(
    const std::vector<double> & x, const std::vector<double> & p, const std::vector<double> & m, int n_interpolations )
{
    const std::size_t n_points = p.size() + ( p.size() - 1 ) * n_interpolations;
    std::vector<std::vector<double>> result( 2, std::vector<double>( n_points ) );
    for( std::size_t i = 0; i + 1 < p.size(); ++i )
    {
        const double x0 = x[i], x1 = x[i + 1], p0 = p[i], p1 = p[i + 1], m0 = m[i], m1 = m[i + 1];
        for( int j = 0; j < n_interpolations + 1; ++j )
        {
            const double t   = j / double( n_interpolations + 1 );
            const double t2  = t * t, t3 = t2 * t;
            const double h00 = 2 * t3 - 3 * t2 + 1, h10 = -2 * t3 + 3 * t2, h01 = t3 - 2 * t2 + t, h11 = t3 - t2;
            const std::size_t idx = i * ( n_interpolations + 1 ) + j;
            result[0][idx]        = x0 + t * ( x1 - x0 );
            result[1][idx]        = h00 * p0 + h10 * p1 + h01 * m0 * ( x0 - x1 ) + h11 * m1 * ( x0 - x1 );
        }
    }
    result[0].back() = x.back();
    result[1].back() = p.back();
    return result;
}


// ---- temporary: the image-batched device chain is not built yet ----
namespace dev
{
class DeviceChain
{
};
} // namespace dev

Method_GNEB::Method_GNEB( std::shared_ptr<Chain> chain_, int solver_, int idx_chain_ )
        : Method( chain_->gneb_parameters, -1, idx_chain_ ), chain( std::move( chain_ ) )
{
    solver = solver_;
    throw std::runtime_error( "spirit_b200: GNEB is not implemented yet" );
}
Method_GNEB::~Method_GNEB() = default;
void Method_GNEB::Iteration( bool ) {}
void Method_GNEB::Hook_Post_Iteration() {}
void Method_GNEB::Finalize() {}
void Method_GNEB::Save_Current( bool, bool ) {}
bool Method_GNEB::Converged()
{
    return true;
}
void Method_GNEB::Sync_Host() {}
void Method_GNEB::Sync_Device() {}

} // namespace sb
