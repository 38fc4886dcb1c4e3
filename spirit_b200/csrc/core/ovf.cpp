#include "ovf.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace sb
{
namespace ovf
{

namespace
{
constexpr int COUNT_DIGITS     = 6;
constexpr float CHECK_4        = 1234567.0f;
constexpr double CHECK_8       = 123456789012345.0;

std::string lower( std::string s )
{
    std::transform( s.begin(), s.end(), s.begin(), []( unsigned char c ) { return char( std::tolower( c ) ); } );
    return s;
}
std::string trim( const std::string & s )
{
    std::size_t b = 0, e = s.size();
    while( b < e && std::isspace( static_cast<unsigned char>( s[b] ) ) )
        ++b;
    while( e > b && std::isspace( static_cast<unsigned char>( s[e - 1] ) ) )
        --e;
    return s.substr( b, e - b );
}
// "# key: value   ## remark" -> (lower-case key, value). False for anything else (plain comments, data lines).
bool key_value( const std::string & line, std::string & key, std::string & value )
{
    std::size_t i = 0;
    while( i < line.size() && std::isspace( static_cast<unsigned char>( line[i] ) ) )
        ++i;
    if( i >= line.size() || line[i] != '#' )
        return false;
    if( i + 1 < line.size() && line[i + 1] == '#' )
        return false; // "##": a remark
    const std::size_t colon = line.find( ':', i );
    if( colon == std::string::npos )
        return false;
    key                 = lower( trim( line.substr( i + 1, colon - i - 1 ) ) );
    std::string rest    = line.substr( colon + 1 );
    const std::size_t r = rest.find( "##" );
    if( r != std::string::npos )
        rest = rest.substr( 0, r );
    value = trim( rest );
    return !key.empty();
}
// line [pos, eol) of `s`; returns the position after the line break
std::size_t next_line( const std::string & s, std::size_t pos, std::string & line )
{
    const std::size_t eol = s.find( '\n', pos );
    if( eol == std::string::npos )
    {
        line = s.substr( pos );
        return s.size();
    }
    line = s.substr( pos, eol - pos );
    if( !line.empty() && line.back() == '\r' )
        line.pop_back();
    return eol + 1;
}
void put_le( std::string & out, const void * value, std::size_t bytes )
{
    // the hosts this library runs on are little-endian (x86-64, aarch64): bytes go out as they are
    out.append( static_cast<const char *>( value ), bytes );
}
std::string number( double v )
{
    // shortest representation that reads back exactly (the reference prints with fmt's default, which is the same idea)
    char buf[64];
    for( int prec = 1; prec <= 17; ++prec )
    {
        std::snprintf( buf, sizeof( buf ), "%.*g", prec, v );
        if( std::strtod( buf, nullptr ) == v )
            break;
    }
    return buf;
}
std::string segment_text( const Segment & seg, const double * data, int format )
{
    const std::string meshtype = seg.meshtype.empty() ? "rectangular" : seg.meshtype;
    int n_rows                 = 0;
    if( meshtype == "rectangular" )
        n_rows = seg.n_cells[0] * seg.n_cells[1] * seg.n_cells[2];
    else if( meshtype == "irregular" )
        n_rows = seg.pointcount;
    else
        throw std::runtime_error( "invalid meshtype \"" + seg.meshtype + "\"" );
    const int n_cols = seg.valuedim;
    if( n_cols * n_rows <= 0 )
        throw std::runtime_error( "segment without data (columns x rows <= 0)" );
    if( format == BIN )
        format = BIN8;
    const char * type = format == BIN8 ? "Binary 8" : format == BIN4 ? "Binary 4" : format == TEXT ? "Text" : format == CSV ? "CSV" : nullptr;
    if( !type )
        throw std::runtime_error( "invalid file format index " + std::to_string( format ) );

    auto labels = [&]( const std::string & given )
    {
        if( !given.empty() )
            return " " + given;
        std::string s = " "; // "# valueunits:  unspecified unspecified" as the reference writes it
        for( int i = 0; i < n_cols; ++i )
            s += " unspecified";
        return s;
    };
    std::string out;
    out.reserve( 1024 + std::size_t( n_rows ) * n_cols * ( format == BIN4 ? 4 : format == BIN8 ? 8 : 23 ) );
    out += "#\n# Begin: Segment\n# Begin: Header\n#\n";
    out += "# Title: " + seg.title + "\n#\n";
    out += "# Desc: " + seg.comment + "\n#\n";
    out += "# valuedim: " + std::to_string( n_cols ) + "   ## field dimensionality\n";
    out += "# valueunits:" + labels( seg.valueunits ) + "\n";
    out += "# valuelabels:" + labels( seg.valuelabels ) + "\n";
    out += "#\n## Fundamental mesh measurement unit. Treated as a label:\n";
    out += "# meshunit: " + ( seg.meshunit.empty() ? std::string( "unspecified" ) : seg.meshunit ) + "\n#\n";
    const char * axes = "xyz";
    for( int i = 0; i < 3; ++i )
        out += std::string( "# " ) + axes[i] + "min: " + number( seg.bounds_min[i] ) + "\n";
    for( int i = 0; i < 3; ++i )
        out += std::string( "# " ) + axes[i] + "max: " + number( seg.bounds_max[i] ) + "\n";
    out += "#\n# meshtype: " + meshtype + "\n";
    if( meshtype == "rectangular" )
    {
        for( int i = 0; i < 3; ++i )
            out += std::string( "# " ) + axes[i] + "base: " + number( seg.origin[i] ) + "\n";
        for( int i = 0; i < 3; ++i )
            out += std::string( "# " ) + axes[i] + "stepsize: " + number( seg.step_size[i] ) + "\n";
        for( int i = 0; i < 3; ++i )
            out += std::string( "# " ) + axes[i] + "nodes: " + std::to_string( seg.n_cells[i] ) + "\n";
    }
    else
        out += "# pointcount: " + std::to_string( seg.pointcount ) + "\n";
    out += "#\n# End: Header\n#\n";
    out += std::string( "# Begin: Data " ) + type + "\n";
    if( format == BIN8 )
    {
        put_le( out, &CHECK_8, 8 );
        put_le( out, data, std::size_t( n_rows ) * n_cols * 8 );
        out += "\n";
    }
    else if( format == BIN4 )
    {
        put_le( out, &CHECK_4, 4 );
        std::vector<float> narrow( std::size_t( n_rows ) * n_cols );
        for( std::size_t i = 0; i < narrow.size(); ++i )
            narrow[i] = float( data[i] );
        put_le( out, narrow.data(), narrow.size() * 4 );
        out += "\n";
    }
    else
    {
        char buf[64];
        for( int row = 0; row < n_rows; ++row )
        {
            for( int col = 0; col < n_cols; ++col )
            {
                std::snprintf( buf, sizeof( buf ), "%22.12f", data[std::size_t( row ) * n_cols + col] );
                out += buf;
                if( format == CSV )
                    out += ",";
            }
            out += "\n";
        }
    }
    out += std::string( "# End: Data " ) + type + "\n# End: Segment\n";
    return out;
}
std::string top_header( int n_segments )
{
    char digits[16];
    std::snprintf( digits, sizeof( digits ), "%0*d", COUNT_DIGITS, n_segments );
    return std::string( "# OOMMF OVF 2.0\n#\n# Segment count: " ) + digits + "\n";
}
} // namespace

File::File( const std::string & file_name ) : name( file_name )
{
    scan();
}

File::File( const std::string & file_name, ForWriting ) : name( file_name )
{
    scan_head();
}

// What a writer needs to know of an existing file, from its first bytes: is it there, is it empty, is it OVF 2.0, how many
// segments does its header declare and where do the digits of that count stand. A header without a count line (or anything
// else unexpected) falls back to the full scan.
void File::scan_head()
{
    found = is_ovf = false;
    n_segments     = 0;
    count_pos_     = std::string::npos;
    head_only_     = true;
    empty_         = false;
    segments_.clear();
    contents_.clear();
    std::ifstream in( name, std::ios::binary );
    if( !in )
    {
        message = "file not found";
        return;
    }
    found = true;
    std::string head( 4096, '\0' );
    in.read( &head[0], std::streamsize( head.size() ) );
    head.resize( std::size_t( in.gcount() ) );
    empty_ = head.empty();
    if( empty_ )
        return;
    std::size_t pos = 0;
    std::string line, key, value;
    bool version_ok = false;
    while( pos < head.size() )
    {
        pos = next_line( head, pos, line );
        if( trim( line ).empty() )
            continue;
        const std::string l = lower( line );
        version_ok          = l.find( "oommf" ) != std::string::npos && l.find( "ovf" ) != std::string::npos && l.find( "2.0" ) != std::string::npos;
        break;
    }
    while( version_ok && pos < head.size() )
    {
        const std::size_t line_start = pos;
        pos                          = next_line( head, pos, line );
        if( !key_value( line, key, value ) )
            continue;
        if( key == "segment count" && pos < head.size() ) // (the whole line lies inside the bytes read)
        {
            const std::size_t c = head.find( ':', line_start );
            count_pos_          = head.find_first_of( "0123456789", c );
            n_segments          = std::atoi( value.c_str() );
            is_ovf              = count_pos_ != std::string::npos;
            contents_           = head;
            if( is_ovf )
                return;
        }
        break;
    }
    scan(); // not the header this library and the reference write: look at the whole file
    head_only_ = false;
}

void File::scan()
{
    found = is_ovf = false;
    n_segments     = 0;
    count_pos_     = std::string::npos; // a rescan of a replaced file must not patch the counter at a stale offset
    segments_.clear();
    contents_.clear();
    std::ifstream in( name, std::ios::binary );
    if( !in )
    {
        message = "file not found";
        return;
    }
    found = true;
    std::ostringstream ss;
    ss << in.rdbuf();
    contents_ = ss.str();

    // top header: "# OOMMF OVF 2.0" (first non-empty line), then "# Segment count: n"
    std::size_t pos = 0;
    std::string line;
    bool version_ok = false;
    while( pos < contents_.size() )
    {
        pos = next_line( contents_, pos, line );
        if( trim( line ).empty() )
            continue;
        const std::string l = lower( line );
        version_ok          = l.find( "oommf" ) != std::string::npos && l.find( "ovf" ) != std::string::npos && l.find( "2.0" ) != std::string::npos;
        break;
    }
    if( !version_ok )
    {
        message = "not an OVF 2.0 file (version line missing)";
        return;
    }
    int declared = -1;
    while( pos < contents_.size() )
    {
        const std::size_t line_start = pos;
        pos                          = next_line( contents_, pos, line );
        std::string key, value;
        if( !key_value( line, key, value ) )
            continue;
        if( key == "segment count" )
        {
            declared            = std::atoi( value.c_str() );
            const std::size_t c = contents_.find( ':', line_start );
            count_pos_          = contents_.find_first_of( "0123456789", c );
            break;
        }
        if( key == "begin" )
        {
            pos = line_start; // no count line: count the segments below
            break;
        }
    }
    // segments: from "# Begin: Segment" to "# End: Segment". Binary blocks are skipped by their length, not scanned.
    while( pos < contents_.size() )
    {
        const std::size_t line_start = pos;
        pos                          = next_line( contents_, pos, line );
        std::string key, value;
        if( !key_value( line, key, value ) || key != "begin" || lower( value ) != "segment" )
            continue;
        Span span;
        span.begin    = line_start;
        int valuedim = 0, rows = 1, pointcount = -1;
        bool closed = false;
        while( pos < contents_.size() )
        {
            pos = next_line( contents_, pos, line );
            if( !key_value( line, key, value ) )
                continue;
            if( key == "valuedim" )
                valuedim = std::atoi( value.c_str() );
            else if( key == "xnodes" || key == "ynodes" || key == "znodes" )
                rows *= std::atoi( value.c_str() );
            else if( key == "pointcount" )
                pointcount = std::atoi( value.c_str() );
            else if( key == "begin" && lower( value ).rfind( "data binary", 0 ) == 0 )
            {
                const int width     = lower( value ).find( '4' ) != std::string::npos ? 4 : 8;
                const std::size_t n = std::size_t( pointcount >= 0 ? pointcount : rows ) * std::size_t( std::max( valuedim, 0 ) );
                pos += std::size_t( width ) * ( n + 1 );
                pos = std::min( pos, contents_.size() );
            }
            else if( key == "end" && lower( value ) == "segment" )
            {
                span.end = pos;
                closed   = true;
                break;
            }
        }
        if( !closed )
        {
            message = "segment " + std::to_string( segments_.size() + 1 ) + " is not closed";
            return;
        }
        segments_.push_back( span );
    }
    n_segments = int( segments_.size() );
    if( declared >= 0 && declared != n_segments )
        message = "segment count in the header (" + std::to_string( declared ) + ") differs from the segments found (" + std::to_string( n_segments ) + ")";
    is_ovf = true;
}

Segment File::read_segment_header( int index ) const
{
    if( head_only_ )
        throw std::logic_error( "ovf::File opened for writing cannot be read" );
    if( !is_ovf )
        throw std::runtime_error( "file \"" + name + "\" is not in OVF format: " + message );
    if( index < 0 || index >= n_segments )
        throw std::runtime_error( "segment index " + std::to_string( index ) + " out of range, file \"" + name + "\" has " + std::to_string( n_segments ) );
    Segment seg;
    seg.meshunit.clear();
    seg.meshtype.clear();
    std::size_t pos = segments_[index].begin;
    std::string line, key, value;
    bool in_header = false, nodes[3] = { false, false, false };
    while( pos < segments_[index].end )
    {
        pos = next_line( contents_, pos, line );
        if( !key_value( line, key, value ) )
            continue;
        if( key == "begin" && lower( value ) == "header" )
            in_header = true;
        else if( key == "end" && lower( value ) == "header" )
            break;
        else if( !in_header )
            continue;
        else if( key == "title" )
            seg.title = value;
        else if( key == "desc" )
            seg.comment = seg.comment.empty() ? value : seg.comment + "\n" + value;
        else if( key == "valuedim" )
            seg.valuedim = std::atoi( value.c_str() );
        else if( key == "valueunits" )
            seg.valueunits = value;
        else if( key == "valuelabels" )
            seg.valuelabels = value;
        else if( key == "meshunit" )
            seg.meshunit = value;
        else if( key == "meshtype" )
            seg.meshtype = lower( value );
        else if( key == "pointcount" )
            seg.pointcount = std::atoi( value.c_str() );
        else if( key.size() == 4 && key.compare( 1, 3, "min" ) == 0 && key[0] >= 'x' && key[0] <= 'z' )
            seg.bounds_min[key[0] - 'x'] = std::strtod( value.c_str(), nullptr );
        else if( key.size() == 4 && key.compare( 1, 3, "max" ) == 0 && key[0] >= 'x' && key[0] <= 'z' )
            seg.bounds_max[key[0] - 'x'] = std::strtod( value.c_str(), nullptr );
        else if( key.size() == 5 && key.compare( 1, 4, "base" ) == 0 && key[0] >= 'x' && key[0] <= 'z' )
            seg.origin[key[0] - 'x'] = std::strtod( value.c_str(), nullptr );
        else if( key.size() == 9 && key.compare( 1, 8, "stepsize" ) == 0 && key[0] >= 'x' && key[0] <= 'z' )
            seg.step_size[key[0] - 'x'] = std::strtod( value.c_str(), nullptr );
        else if( key.size() == 6 && key.compare( 1, 5, "nodes" ) == 0 && key[0] >= 'x' && key[0] <= 'z' )
        {
            seg.n_cells[key[0] - 'x'] = std::atoi( value.c_str() );
            nodes[key[0] - 'x']       = true;
        }
    }
    if( seg.valuedim <= 0 )
        throw std::runtime_error( "segment header without a valid valuedim" );
    if( seg.meshtype == "rectangular" )
    {
        if( !( nodes[0] && nodes[1] && nodes[2] ) )
            throw std::runtime_error( "rectangular mesh without xnodes / ynodes / znodes" );
        seg.N = seg.n_cells[0] * seg.n_cells[1] * seg.n_cells[2];
    }
    else if( seg.meshtype == "irregular" )
        seg.N = seg.pointcount;
    else
        throw std::runtime_error( "segment header without a valid meshtype" );
    return seg;
}

void File::read_segment_data( int index, const Segment & seg, int n_rows, double * data ) const
{
    if( index < 0 || index >= n_segments )
        throw std::runtime_error( "segment index out of range" );
    const Segment in_file = read_segment_header( index );
    n_rows                = std::min( n_rows, in_file.N );
    const int n_cols      = in_file.valuedim;
    if( seg.valuedim != n_cols )
        throw std::runtime_error( "segment has " + std::to_string( n_cols ) + " columns, " + std::to_string( seg.valuedim ) + " expected" );
    std::size_t pos = segments_[index].begin;
    std::string line, key, value;
    while( pos < segments_[index].end )
    {
        pos = next_line( contents_, pos, line );
        if( key_value( line, key, value ) && key == "begin" && lower( value ).rfind( "data", 0 ) == 0 )
            break;
    }
    const std::string type = lower( value ); // "data text", "data csv", "data binary 4", "data binary 8"
    const std::size_t n    = std::size_t( n_rows ) * n_cols;
    if( type.rfind( "data binary", 0 ) == 0 )
    {
        const int width = type.find( '4' ) != std::string::npos ? 4 : 8;
        if( pos + std::size_t( width ) * ( std::size_t( in_file.N ) * n_cols + 1 ) > contents_.size() )
            throw std::runtime_error( "binary data block is shorter than its header says" );
        if( width == 8 )
        {
            double check;
            std::memcpy( &check, contents_.data() + pos, 8 );
            if( check != CHECK_8 )
                throw std::runtime_error( "wrong check value of the Binary 8 block (byte order?)" );
            std::memcpy( data, contents_.data() + pos + 8, n * 8 );
        }
        else
        {
            float check;
            std::memcpy( &check, contents_.data() + pos, 4 );
            if( check != CHECK_4 )
                throw std::runtime_error( "wrong check value of the Binary 4 block (byte order?)" );
            std::vector<float> narrow( n );
            std::memcpy( narrow.data(), contents_.data() + pos + 4, n * 4 );
            for( std::size_t i = 0; i < n; ++i )
                data[i] = double( narrow[i] );
        }
        return;
    }
    if( type != "data text" && type != "data csv" )
        throw std::runtime_error( "unknown data block \"" + value + "\"" );
    // numbers separated by blanks and / or commas; '#' lines are remarks until "# End: Data"
    std::size_t have = 0;
    while( pos < segments_[index].end && have < n )
    {
        pos = next_line( contents_, pos, line );
        if( key_value( line, key, value ) && key == "end" )
            break;
        const char * p = line.c_str();
        while( *p && have < n )
        {
            while( *p && ( std::isspace( static_cast<unsigned char>( *p ) ) || *p == ',' ) )
                ++p;
            if( !*p || *p == '#' )
                break;
            char * end     = nullptr;
            const double v = std::strtod( p, &end );
            if( end == p )
                throw std::runtime_error( "text data block: cannot read a number from \"" + line + "\"" );
            data[have++] = v;
            p            = end;
        }
    }
    if( have < n )
        throw std::runtime_error( "text data block holds " + std::to_string( have ) + " values, " + std::to_string( n ) + " expected" );
}

void File::write_segment( const Segment & segment, const double * data, int format )
{
    const std::string body = segment_text( segment, data, format );
    std::ofstream out( name, std::ios::binary | std::ios::trunc );
    if( !out )
        throw std::runtime_error( "cannot open \"" + name + "\" for writing" );
    out << top_header( 1 ) << body;
    out.close();
    if( head_only_ )
    {
        // what scan_head() would find, without reading the file back
        found = is_ovf = true;
        empty_         = false;
        n_segments     = 1;
        contents_      = top_header( 1 );
        const std::size_t c = contents_.find( ':', lower( contents_ ).find( "segment count" ) );
        count_pos_          = contents_.find_first_of( "0123456789", c );
        return;
    }
    scan();
}

void File::append_segment( const Segment & segment, const double * data, int format )
{
    if( !head_only_ )
        scan();
    if( !found || ( head_only_ ? empty_ : contents_.empty() ) )
    {
        write_segment( segment, data, format );
        return;
    }
    if( !is_ovf )
        throw std::runtime_error( "cannot append to \"" + name + "\": " + message );
    const std::string body = segment_text( segment, data, format );
    {
        std::ofstream out( name, std::ios::binary | std::ios::app );
        if( !out )
            throw std::runtime_error( "cannot open \"" + name + "\" for appending" );
        out << body;
    }
    if( count_pos_ != std::string::npos && count_pos_ > 0 )
    {
        // patch the segment count in place (same number of digits as it was written with)
        std::size_t digits = 0;
        while( count_pos_ + digits < contents_.size() && std::isdigit( static_cast<unsigned char>( contents_[count_pos_ + digits] ) ) )
            ++digits;
        char buf[32];
        std::snprintf( buf, sizeof( buf ), "%0*d", int( digits ), n_segments + 1 );
        if( std::strlen( buf ) == digits )
        {
            std::fstream patch( name, std::ios::binary | std::ios::in | std::ios::out );
            patch.seekp( std::streamoff( count_pos_ ) );
            patch.write( buf, std::streamsize( digits ) );
            if( head_only_ )
                contents_.replace( count_pos_, digits, buf );
        }
    }
    if( head_only_ )
        ++n_segments; // (the next append of a chain continues from here: no pass over the file)
    else
        scan();
}

} // namespace ovf
} // namespace sb
