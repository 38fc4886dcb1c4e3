// Spirit/IO.h: spin configurations and chains as OVF 2.0 files (core/src/Spirit/IO.cpp:86-860). The configurations are
// the host copies the API works on between simulations; nothing here touches the device.
#include "api_common.hpp"

#include "../core/io.hpp"
#include "../core/ovf.hpp"

#include <Spirit/Chain.h>
#include <Spirit/Configurations.h>
#include <Spirit/IO.h>
#include <Spirit/State.h>
#include <Spirit/Version.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <string>

using namespace sb;

namespace
{
using sb::io::segment_of;
using sb::io::spin_segment;
void check_format( int format )
{
    if( format < ovf::BIN || format > ovf::CSV )
        throw std::runtime_error( "Invalid file format index " + std::to_string( format ) );
}
void warn_extension( const char * file, int idx_image, int idx_chain )
{
    const std::string name( file );
    const std::size_t dot = name.rfind( '.' );
    if( dot == std::string::npos || name.substr( dot ) != ".ovf" )
        Log( Log_Level::Warning, Log_Sender::API,
             "The file \"" + name + "\" is written in OVF format but has different extension. It is recommend to use the appropriate \".ovf\" extension",
             idx_image, idx_chain );
}
// One segment of `file` into the spins of `image` (IO.cpp:224-277): at most nos rows, 3 columns, then every spin is
// normalised; (near-)zero vectors become +z (vacancies of a defect build)
void read_spins( const ovf::File & file, int idx_in_file, Spin_System & image, int idx_image, int idx_chain )
{
    ovf::Segment seg = file.read_segment_header( idx_in_file );
    if( seg.N < image.nos )
        Log( Log_Level::Warning, Log_Sender::API,
             "OVF file \"" + file.name + "\": segment " + std::to_string( idx_in_file + 1 ) + "/" + std::to_string( file.n_segments ) + " contains only "
                 + std::to_string( seg.N ) + " spins while the system contains " + std::to_string( image.nos ) + ".",
             idx_image, idx_chain );
    else if( seg.N > image.nos )
        Log( Log_Level::Warning, Log_Sender::API,
             "OVF file \"" + file.name + "\": segment " + std::to_string( idx_in_file + 1 ) + "/" + std::to_string( file.n_segments ) + " contains "
                 + std::to_string( seg.N ) + " spins while the system contains only " + std::to_string( image.nos )
                 + ". Reading only part of the segment data.",
             idx_image, idx_chain );
    if( seg.valuedim != 3 )
        throw std::runtime_error(
            "Segment " + std::to_string( idx_in_file + 1 ) + "/" + std::to_string( file.n_segments ) + " in OVF file \"" + file.name
            + "\" should have 3 columns, but only has " + std::to_string( seg.valuedim ) + ". Will not read." );
    static_assert( sizeof( Vec3 ) == 3 * sizeof( double ), "spins are packed triples of doubles" );
    file.read_segment_data( idx_in_file, seg, std::min( seg.N, image.nos ), &image.spins[0].x );
    for( int i = 0; i < image.nos; ++i )
    {
        Vec3 & s = image.spins[i];
        if( s.norm() < 1e-5 )
        {
            s = Vec3{ 0, 0, 1 };
            image.geometry->set_vacancy_read_from_file( i ); // IO.cpp:268-272 (defects are always built here)
        }
        else
            s.normalize();
    }
}
// image `idx_in_file` of a plain column file (Dataparser.cpp:23-51): rows idx nos .. (idx + 1) nos - 1, as many as there are
void read_spins_columns( const std::vector<double> & rows, int idx_in_file, Spin_System & image )
{
    const std::size_t n_rows = rows.size() / 3, first = std::size_t( image.nos ) * std::size_t( std::max( idx_in_file, 0 ) );
    for( int i = 0; i < image.nos && first + i < n_rows; ++i )
        image.spins[i] = Vec3{ rows[3 * ( first + i )], rows[3 * ( first + i ) + 1], rows[3 * ( first + i ) + 2] };
    for( int i = 0; i < image.nos; ++i )
    {
        Vec3 & s = image.spins[i];
        if( s.norm() < 1e-5 )
        {
            s = Vec3{ 0, 0, 1 };
            image.geometry->set_vacancy_read_from_file( i ); // Dataparser.cpp:39-46
        }
        else
            s.normalize();
    }
}
} // namespace

int IO_N_Images_In_File( State *, const char * file, int idx_image, int idx_chain ) noexcept
try
{
    ovf::File f( file );
    if( f.is_ovf )
        return f.n_segments;
    Log( Log_Level::Warning, Log_Sender::API, std::string( "File \"" ) + file + "\" is not OVF. Cannot measure number of images.", idx_image, idx_chain );
    return -1;
}
SB_API_CATCH_RET( -1 )

void IO_Positions_Write( State * state, const char * file, int format, const char * comment, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    check_format( format );
    warn_extension( file, idx_image, idx_chain );
    ovf::Segment seg = segment_of( *image );
    seg.title        = std::string( "SPIRIT Version " ) + io::version_full();
    seg.comment      = comment;
    seg.valuedim     = 3;
    seg.valuelabels  = "position_x position_y position_z";
    seg.valueunits   = "none none none";
    ovf::File( file, ovf::File::ForWriting{} ).write_segment( seg, &image->geometry->positions()[0].x, format );
    Log( Log_Level::Info, Log_Sender::API, std::string( "Wrote positions to file \"" ) + file + "\"", idx_image, idx_chain );
}
SB_API_CATCH_VOID

void IO_Image_Read( State * state, const char * file, int idx_image_infile, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    ovf::File f( file );
    if( !f.found )
        throw std::runtime_error( std::string( "Unable open file \"" ) + file + "\", are you sure it exists?" );
    if( !f.is_ovf )
    {
        Log( Log_Level::Error, Log_Sender::API,
             std::string( "File \"" ) + file + "\" does not seem to be in valid OVF format. Message: " + f.message
                 + ". Will try to read as data column text format file.",
             idx_image, idx_chain );
        read_spins_columns( io::read_column_text( file ), idx_image_infile, *image );
        return;
    }
    read_spins( f, idx_image_infile, *image, idx_image, idx_chain );
    Log( Log_Level::Info, Log_Sender::API, std::string( "Read image from file \"" ) + file + "\"", idx_image, idx_chain );
}
SB_API_CATCH_VOID

void IO_Image_Write( State * state, const char * file, int format, const char * comment, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    check_format( format );
    warn_extension( file, idx_image, idx_chain );
    ovf::File( file, ovf::File::ForWriting{} ).write_segment( spin_segment( *image, comment ), &image->spins[0].x, format );
    Log( Log_Level::Info, Log_Sender::API, std::string( "Wrote spins to file \"" ) + file + "\"", idx_image, idx_chain );
}
SB_API_CATCH_VOID

void IO_Image_Append( State * state, const char * file, int format, const char * comment, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    check_format( format );
    warn_extension( file, idx_image, idx_chain );
    ovf::File( file, ovf::File::ForWriting{} ).append_segment( spin_segment( *image, comment ), &image->spins[0].x, format );
    Log( Log_Level::Info, Log_Sender::API, std::string( "Appended spins to file \"" ) + file + "\"", idx_image, idx_chain );
}
SB_API_CATCH_VOID

void IO_Chain_Read( State * state, const char * file, int start_image_infile, int end_image_infile, int insert_idx, int idx_chain ) noexcept
{
    int idx_image = insert_idx;
    try
    {
        auto chain = resolve( state, idx_image, idx_chain ).chain;
        int noi    = chain->noi;
        if( insert_idx < 0 )
            insert_idx = 0;
        if( insert_idx > noi )
            Log( Log_Level::Error, Log_Sender::API,
                 "IO_Chain_Read: Tried to start reading chain on invalid index(insert_idx=" + std::to_string( insert_idx ) + ", but chain has "
                     + std::to_string( noi ) + " images)",
                 insert_idx, idx_chain );
        ovf::File f( file );
        if( !f.found )
            throw std::runtime_error( std::string( "Unable open file \"" ) + file + "\", are you sure it exists?" );
        std::vector<double> columns; // plain column file: nos rows per image (Dataparser.cpp:53-110)
        if( !f.is_ovf )
        {
            Log( Log_Level::Warning, Log_Sender::API, std::string( "IO_Chain_Read: File \"" ) + file + "\" seems to not be OVF. Trying to read column data",
                 insert_idx, idx_chain );
            columns = io::read_column_text( file );
        }
        const int noi_infile = f.is_ovf ? f.n_segments : int( columns.size() / 3 / std::size_t( chain->images[0]->nos ) );
        if( start_image_infile < 0 )
            start_image_infile = 0;
        if( end_image_infile < 0 )
            end_image_infile = noi_infile - 1;
        if( end_image_infile < start_image_infile || end_image_infile >= noi_infile )
        {
            Log( Log_Level::Warning, Log_Sender::API,
                 "IO_Chain_Read: specified invalid reading range (start_image_infile=" + std::to_string( start_image_infile )
                     + ", end_image_infile=" + std::to_string( end_image_infile ) + "). Set to read entire file \"" + file + "\" ("
                     + std::to_string( noi_infile ) + " images).",
                 insert_idx, idx_chain );
            end_image_infile = noi_infile - 1;
        }
        if( start_image_infile >= noi_infile )
            throw std::runtime_error(
                "Specified starting index " + std::to_string( start_image_infile ) + ", but file \"" + file + "\" contains only "
                + std::to_string( noi_infile ) + " images." );
        const int noi_to_read = end_image_infile - start_image_infile + 1;
        const int noi_to_add  = noi_to_read - ( noi - insert_idx );
        if( noi_to_add > 0 )
        {
            // the chain grows by copies of its last image (IO.cpp:514-520)
            Chain_Image_to_Clipboard( state, noi - 1, idx_chain );
            Chain_Set_Length( state, noi + noi_to_add, idx_chain );
        }
        {
            chain->Lock();
            try
            {
                // (the reference's loop bounds, IO.cpp:523: images insert_idx .. noi_to_read - 1)
                for( int i = insert_idx; i < noi_to_read; ++i )
                {
                    if( f.is_ovf )
                        read_spins( f, start_image_infile, *chain->images[i], i, idx_chain );
                    else
                        read_spins_columns( columns, start_image_infile, *chain->images[i] );
                    ++start_image_infile;
                }
            }
            catch( ... )
            {
                chain->Unlock();
                throw;
            }
            chain->Unlock();
        }
        Chain_Setup_Data( state, idx_chain );
        Log( Log_Level::Info, Log_Sender::API, std::string( "Read chain from file \"" ) + file + "\"", insert_idx, idx_chain );
    }
    catch( ... )
    {
        sb::handle_exception_api( __func__, idx_image, idx_chain );
    }
}

namespace
{
void chain_to_file( State * state, const char * file, int format, const char * comment, int idx_chain, bool append )
{
    int idx_image = 0;
    auto chain    = resolve( state, idx_image, idx_chain ).chain;
    check_format( format );
    chain->Lock();
    try
    {
        ovf::File f( file, ovf::File::ForWriting{} );
        for( int i = 0; i < chain->noi; ++i )
        {
            const std::string desc = "Image " + std::to_string( i + 1 ) + " of " + std::to_string( chain->noi ) + ". " + comment;
            const ovf::Segment seg = spin_segment( *chain->images[i], desc );
            if( i == 0 && !append )
                f.write_segment( seg, &chain->images[i]->spins[0].x, format );
            else
                f.append_segment( seg, &chain->images[i]->spins[0].x, format );
        }
    }
    catch( ... )
    {
        chain->Unlock();
        throw;
    }
    chain->Unlock();
    Log( Log_Level::Info, Log_Sender::API, std::string( append ? "Appended chain to file \"" : "Wrote chain to file \"" ) + file + "\"", 0, idx_chain );
}
} // namespace

void IO_Chain_Write( State * state, const char * file, int format, const char * comment, int idx_chain ) noexcept
{
    int idx_image = 0;
    try
    {
        chain_to_file( state, file, format, comment, idx_chain, false );
    }
    SB_API_CATCH_VOID
}

void IO_Chain_Append( State * state, const char * file, int format, const char * comment, int idx_chain ) noexcept
{
    int idx_image = 0;
    try
    {
        chain_to_file( state, file, format, comment, idx_chain, true );
    }
    SB_API_CATCH_VOID
}

// IO.cpp:851-950: total and per-term energy of every spin as an OVF field with 1 + n_terms columns
void IO_Image_Write_Energy_per_Spin( State * state, const char * file, int format, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    ImageLock lock( *image );
    check_format( format );
    warn_extension( file, idx_image, idx_chain );
    Spin_System & sys = *image;
    sys.UpdateEnergy();
    const int n_terms = int( sys.E_array.size() );
    std::vector<double> totals( static_cast<std::size_t>( n_terms ), 0.0 );
    std::vector<double> per_term( static_cast<std::size_t>( n_terms ) * sys.nos, 0.0 );
    sys.device().energy_contributions( *sys.hamiltonian, totals.data(), per_term.data() );
    const int width = 1 + n_terms;
    std::vector<double> data( std::size_t( width ) * sys.nos, 0.0 );
    for( int i = 0; i < sys.nos; ++i )
    {
        double e = 0;
        for( int t = 0; t < n_terms; ++t )
        {
            const double v                         = per_term[std::size_t( t ) * sys.nos + i];
            data[std::size_t( i ) * width + 1 + t] = v;
            e += v;
        }
        data[std::size_t( i ) * width] = e;
    }
    ovf::Segment seg = segment_of( sys );
    seg.title        = std::string( "SPIRIT Version " ) + io::version_full();
    seg.comment      = "Energy per spin. Total=" + io::shortest( sys.E ) + "meV";
    seg.valuelabels  = "Total";
    seg.valueunits   = "meV";
    for( const auto & pair : sys.E_array )
    {
        seg.comment += ", " + pair.first + "=" + io::shortest( pair.second ) + "meV";
        seg.valuelabels += " " + pair.first;
        seg.valueunits += " meV";
    }
    seg.valuedim = width;
    ovf::File( file, ovf::File::ForWriting{} ).write_segment( seg, data.data(), format );
    Log( Log_Level::Info, Log_Sender::API, std::string( "Wrote energies per spin to file \"" ) + file + "\"", idx_image, idx_chain );
}
SB_API_CATCH_VOID

// IO.cpp:952-972
void IO_Image_Write_Energy( State * state, const char * file, int idx_image, int idx_chain ) noexcept
try
{
    auto image = resolve( state, idx_image, idx_chain ).image;
    io::write_image_energy( *image, file, true, true );
}
SB_API_CATCH_VOID

// IO.cpp:975-992
void IO_Chain_Write_Energies( State * state, const char * file, int idx_chain ) noexcept
{
    int idx_image = -1;
    try
    {
        auto chain = resolve( state, idx_image, idx_chain ).chain;
        io::write_chain_energies( *chain, file, true, true );
    }
    SB_API_CATCH_VOID
}

// IO.cpp:730-790 / Datawriter.cpp:22-113: the pair lists the stencil kernels work on, as text tables. The lists are the
// redundant ones (both directions of every pair), like those of the reference's OpenMP and CUDA builds: no mirrored lines.
namespace
{
std::string fixed8( double v )
{
    char buf[64];
    std::snprintf( buf, sizeof( buf ), "%.8f", v );
    return buf;
}
std::string pair_columns( const Pair & p )
{
    return io::centred( std::to_string( p.i ), 3 ) + " " + io::centred( std::to_string( p.j ), 3 ) + "    " + io::centred( std::to_string( p.translations[0] ), 3 )
           + " " + io::centred( std::to_string( p.translations[1] ), 3 ) + " " + io::centred( std::to_string( p.translations[2] ), 3 ) + "    ";
}
} // namespace

void IO_Image_Write_Neighbours_Exchange( State * state, const char * file, int idx_image, int idx_chain ) noexcept
try
{
    auto image            = resolve( state, idx_image, idx_chain ).image;
    const Hamiltonian & h = *image->hamiltonian;
    std::string out       = "###    Interaction neighbours:\n";
    out += "n_neighbours_exchange " + std::to_string( h.exchange_pairs.size() ) + "\n";
    if( !h.exchange_pairs.empty() )
    {
        out += io::centred( "i", 3 ) + " " + io::centred( "j", 3 ) + "    " + io::centred( "da", 3 ) + " " + io::centred( "db", 3 ) + " " + io::centred( "dc", 3 )
               + "    " + io::centred( "Jij", 15 ) + "\n";
        for( std::size_t k = 0; k < h.exchange_pairs.size(); ++k )
            out += pair_columns( h.exchange_pairs[k] ) + io::centred( fixed8( h.exchange_magnitudes[k] ), 15 ) + "\n";
    }
    std::ofstream( file, std::ios::trunc ) << out;
}
SB_API_CATCH_VOID

void IO_Image_Write_Neighbours_DMI( State * state, const char * file, int idx_image, int idx_chain ) noexcept
try
{
    auto image            = resolve( state, idx_image, idx_chain ).image;
    const Hamiltonian & h = *image->hamiltonian;
    std::string out       = "###    Interaction neighbours:\n";
    out += "n_neighbours_dmi " + std::to_string( h.dmi_pairs.size() ) + "\n";
    if( !h.dmi_pairs.empty() )
    {
        out += io::centred( "i", 3 ) + " " + io::centred( "j", 3 ) + "    " + io::centred( "da", 3 ) + " " + io::centred( "db", 3 ) + " " + io::centred( "dc", 3 )
               + "    " + io::centred( "Dij", 15 ) + " " + io::centred( "Dijx", 15 ) + " " + io::centred( "Dijy", 15 ) + " " + io::centred( "Dijz", 15 ) + "\n";
        for( std::size_t k = 0; k < h.dmi_pairs.size(); ++k )
            out += pair_columns( h.dmi_pairs[k] ) + io::centred( fixed8( h.dmi_magnitudes[k] ), 15 ) + " " + io::centred( fixed8( h.dmi_normals[k].x ), 15 ) + " "
                   + io::centred( fixed8( h.dmi_normals[k].y ), 15 ) + " " + io::centred( fixed8( h.dmi_normals[k].z ), 15 ) + "\n";
    }
    std::ofstream( file, std::ios::trunc ) << out;
}
SB_API_CATCH_VOID

// IO.cpp:36-85: the image becomes the system described by another input file (same number of spins as every image of the
// chain), and starts from a random configuration
int IO_System_From_Config( State * state, const char * file, int idx_image, int idx_chain ) noexcept
try
{
    auto ref  = resolve( state, idx_image, idx_chain );
    auto image = ref.image;
    std::shared_ptr<Spin_System> system = config::Spin_System_from_Config( file );
    for( const auto & other : ref.chain->images )
        if( other->nos != system->nos )
            return 0;
    {
        ImageLock lock( *image );
        image->drop_device();
        image->nos             = system->nos;
        image->spins           = system->spins;
        image->effective_field = system->effective_field;
        image->E               = system->E;
        image->E_array         = system->E_array;
        image->M               = system->M;
        image->geometry        = std::make_shared<Geometry>( *system->geometry );
        image->hamiltonian     = std::make_shared<Hamiltonian>( *system->hamiltonian );
        image->hamiltonian->geometry = image->geometry;
        image->llg_parameters        = std::make_shared<Parameters_LLG>( *system->llg_parameters );
    }
    const float position[3] = { 0, 0, 0 }, rectangular[3] = { -1, -1, -1 };
    Configuration_Random( state, position, rectangular, -1, -1, false, false, idx_image, idx_chain );
    return 1;
}
SB_API_CATCH_RET( 0 )
