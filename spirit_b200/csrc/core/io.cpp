#include "io.hpp"

#include <chrono>
#include <ctime>

namespace sb
{
namespace io
{

ovf::Segment segment_of( const Spin_System & system )
{
    const Geometry & g = *system.geometry;
    ovf::Segment seg;
    seg.meshtype   = "rectangular";
    seg.meshunit   = "nm";
    seg.n_cells[0] = g.n_cells[0] * g.n_cell_atoms;
    seg.n_cells[1] = g.n_cells[1];
    seg.n_cells[2] = g.n_cells[2];
    seg.N          = system.nos;
    for( int i = 0; i < 3; ++i )
    {
        seg.bounds_min[i] = g.bounds_min[i] * 0.1;
        seg.bounds_max[i] = g.bounds_max[i] * 0.1;
        seg.origin[i]     = 0;
        seg.step_size[i]  = g.lattice_constant * g.bravais_vectors[i][i] * 0.1;
    }
    return seg;
}

ovf::Segment spin_segment( const Spin_System & system, const std::string & comment )
{
    ovf::Segment seg = segment_of( system );
    seg.title        = std::string( "SPIRIT Version " ) + version_full();
    seg.comment      = comment;
    seg.valuedim     = 3;
    seg.valuelabels  = "spin_x spin_y spin_z";
    seg.valueunits   = "none none none";
    return seg;
}

std::string current_date_time()
{
    const std::time_t t = std::chrono::system_clock::to_time_t( std::chrono::system_clock::now() );
    std::tm parts{};
    localtime_r( &t, &parts );
    char buf[64];
    std::strftime( buf, sizeof( buf ), "%Y-%m-%d_%H-%M-%S", &parts );
    return buf;
}

} // namespace io
} // namespace sb
