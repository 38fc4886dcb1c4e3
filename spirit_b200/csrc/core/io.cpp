#include "io.hpp"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <stdexcept>

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

std::vector<double> read_column_text( const std::string & file )
{
    std::ifstream in( file );
    if( !in )
        throw std::runtime_error( "Unable open file \"" + file + "\", are you sure it exists?" );
    std::vector<double> rows;
    std::string line;
    while( std::getline( in, line ) )
    {
        const std::size_t remark = line.find( '#' );
        if( remark != std::string::npos )
            line.erase( remark );
        const char * p = line.c_str();
        double v[3];
        int have = 0;
        while( *p && have < 3 )
        {
            while( *p && ( std::isspace( static_cast<unsigned char>( *p ) ) || *p == ',' ) )
                ++p;
            if( !*p )
                break;
            char * end = nullptr;
            v[have]    = std::strtod( p, &end );
            if( end == p )
                break;
            ++have;
            p = end;
        }
        if( have == 3 )
            rows.insert( rows.end(), v, v + 3 );
    }
    return rows;
}

std::string centred( const std::string & text, std::size_t width )
{
    if( text.size() >= width )
        return text;
    const std::size_t left = ( width - text.size() ) / 2;
    return std::string( left, ' ' ) + text + std::string( width - text.size() - left, ' ' );
}
std::string fixed10( double v )
{
    char buf[64];
    std::snprintf( buf, sizeof( buf ), "%.10f", v );
    return buf;
}
std::string shortest( double v )
{
    char buf[64];
    for( int prec = 1; prec <= 17; ++prec )
    {
        std::snprintf( buf, sizeof( buf ), "%.*g", prec, v );
        if( std::strtod( buf, nullptr ) == v )
            break;
    }
    return buf;
}

void write_energy_header( const Spin_System & s, const std::string & file, const std::vector<std::string> & columns, bool readability )
{
    std::string separator, line;
    for( const auto & column : columns )
    {
        if( readability )
            separator += "----------------------++";
        line += " " + centred( column ) + " ||";
    }
    bool first = true;
    for( const auto & pair : s.E_array )
    {
        if( !first )
        {
            line += "|";
            if( readability )
                separator += "+";
        }
        first = false;
        line += " " + centred( pair.first ) + " ";
        if( readability )
            separator += "----------------------";
    }
    line += "\n";
    separator += "\n";
    std::string header = readability ? separator + line + separator : line;
    if( !readability )
        std::replace( header.begin(), header.end(), '|', ' ' );
    std::ofstream( file, std::ios::trunc ) << header;
}

void append_image_energy( const Spin_System & s, long iteration, const std::string & file, bool normalize, bool readability )
{
    const double norm = normalize ? 1.0 / double( s.nos ) : 1.0;
    std::string line   = " " + centred( std::to_string( iteration ) ) + " || " + centred( fixed10( s.E * norm ) ) + " |";
    for( const auto & pair : s.E_array )
        line += "| " + centred( fixed10( pair.second * norm ) ) + " ";
    line += "\n";
    if( !readability )
        std::replace( line.begin(), line.end(), '|', ' ' );
    std::ofstream( file, std::ios::app ) << line;
}

void write_image_energy( const Spin_System & s, const std::string & file, bool normalize, bool readability )
{
    const double norm = normalize ? 1.0 / double( s.nos ) : 1.0;
    write_energy_header( s, file, { "E_tot" }, true );
    std::string line = " " + centred( fixed10( s.E * norm ) ) + " |";
    for( const auto & pair : s.E_array )
        line += "| " + centred( fixed10( pair.second * norm ) ) + " ";
    line += "\n";
    if( !readability )
        std::replace( line.begin(), line.end(), '|', ' ' );
    std::ofstream( file, std::ios::app ) << line;
}

void write_chain_energies( const Chain & chain, const std::string & file, bool normalize, bool readability )
{
    const double norm = normalize ? 1.0 / double( chain.images[0]->nos ) : 1.0;
    write_energy_header( *chain.images[0], file, { "image", "Rx", "E_tot" }, true );
    std::ofstream out( file, std::ios::app );
    for( int i = 0; i < chain.noi; ++i )
    {
        const Spin_System & s = *chain.images[i];
        std::string line = " " + centred( std::to_string( i ) ) + " || " + centred( fixed10( chain.Rx[i] ) ) + " || " + centred( fixed10( s.E * norm ) ) + " |";
        for( const auto & pair : s.E_array )
            line += "| " + centred( fixed10( pair.second * norm ) ) + " ";
        line += "\n";
        if( !readability )
            std::replace( line.begin(), line.end(), '|', ' ' );
        out << line;
    }
}

} // namespace io
} // namespace sb
