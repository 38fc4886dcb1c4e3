#include "config.hpp"
#include "constants.hpp"
#include "logging.hpp"

#include <algorithm>
#include <cctype>
#include <fstream>
#include <random>
#include <stdexcept>

namespace sb
{
namespace config
{

namespace
{
std::string lower( std::string s )
{
    std::transform( s.begin(), s.end(), s.begin(), []( unsigned char c ) { return char( std::tolower( c ) ); } );
    return s;
}

int count_words( const std::string & s )
{
    std::istringstream stream( s );
    std::string w;
    int n = 0;
    while( stream >> w )
        ++n;
    return n;
}
} // namespace

// Filter_File_Handle.cpp:79-110: '|' and '+' are removed, a line that starts with '#' is skipped,
// anything after a '#' is cut. Empty lines are kept (GetLine returns them).
ConfigFile::ConfigFile( const std::string & filename )
{
    std::ifstream in( filename, std::ios::in | std::ios::binary );
    if( !in.is_open() )
        throw std::runtime_error( "Could not open file \"" + filename + "\"" );
    std::string line;
    while( std::getline( in, line ) )
    {
        line.erase( std::remove( line.begin(), line.end(), '|' ), line.end() );
        line.erase( std::remove( line.begin(), line.end(), '+' ), line.end() );
        const auto hash = line.find( '#' );
        if( hash == 0 )
            continue;
        if( hash != std::string::npos )
            line.erase( hash );
        lines_.push_back( line );
    }
}

// Filter_File_Handle.cpp:115-160: prefix match on the (lower-cased) line, then skip the keyword's words
bool ConfigFile::Find( const std::string & keyword )
{
    const std::string key = lower( keyword );
    for( std::size_t i = 0; i < lines_.size(); ++i )
    {
        const std::string l = lower( lines_[i] );
        if( l.compare( 0, key.size(), key ) == 0 )
        {
            iss.clear();
            iss.str( lines_[i] );
            std::string dump;
            for( int w = 0, n = count_words( keyword ); w < n; ++w )
                iss >> dump;
            cursor_ = i + 1;
            return true;
        }
    }
    return false;
}

bool ConfigFile::GetLine()
{
    if( cursor_ >= lines_.size() )
        return false;
    iss.clear();
    iss.str( lines_[cursor_++] );
    return true;
}

bool ConfigFile::Read_String( std::string & var, const std::string & keyword )
{
    if( !Find( keyword ) )
        return false;
    std::getline( iss, var );
    const auto start = var.find_first_not_of( " \t\n\r\f\v" );
    const auto end   = var.find_last_not_of( " \t\n\r\f\v" );
    if( start != std::string::npos )
        var = var.substr( start, end - start + 1 );
    return true;
}

long duration_from_string( const std::string & s )
{
    long hours = 0, minutes = 0, seconds = 0;
    std::istringstream iss( s );
    std::string token;
    if( std::getline( iss, token, ':' ) && !token.empty() )
        hours = std::stol( token );
    if( std::getline( iss, token, ':' ) && !token.empty() )
        minutes = std::stol( token );
    if( std::getline( iss, token, ':' ) && !token.empty() )
        seconds = std::stol( token );
    return hours * 3600 + minutes * 60 + seconds;
}

// ---------------------------------------------------------------------------------------------
// Log (Configparser.cpp:27-152): only the console switches matter here; no log file is written.
// ---------------------------------------------------------------------------------------------
void Log_from_Config( const std::string & config_file, bool quiet )
{
    bool to_console   = true;
    int console_level = int( Log_Level::Parameter );
    if( !config_file.empty() )
    {
        try
        {
            ConfigFile f( config_file );
            f.Read_Single( to_console, "log_to_console" );
            f.Read_Single( console_level, "log_console_level" );
        }
        catch( ... )
        {
        }
    }
    Log.messages_to_console = to_console;
    Log.level_console       = Log_Level( std::max( 0, std::min( 6, console_level ) ) );
    if( quiet )
    {
        // State.cpp / Configparser.cpp:120-130: quiet mode only lets errors through to the console
        Log.level_console = Log_Level::Error;
    }
}

// ---------------------------------------------------------------------------------------------
// Geometry (Configparser.cpp:186-700, Dataparser.cpp:259-287)
// ---------------------------------------------------------------------------------------------
std::shared_ptr<Geometry> Geometry_from_Config( const std::string & config_file )
{
    std::vector<Vec3> bravais_vectors = Geometry::BravaisVectorsSC();
    std::vector<Vec3> cell_atoms      = { Vec3{ 0, 0, 0 } };
    std::vector<double> cell_mu_s     = { 1.0 };
    double lattice_constant           = 1;
    std::array<int, 3> n_cells        = { 100, 100, 1 };

    if( !config_file.empty() )
    {
        try
        {
            ConfigFile f( config_file );
            f.Read_Single( lattice_constant, "lattice_constant" );

            // Bravais lattice type or explicit vectors / matrix (Configparser.cpp:186-268)
            if( f.Find( "bravais_lattice" ) )
            {
                std::string type;
                f.iss >> type;
                type = lower( type );
                if( type == "sc" )
                    bravais_vectors = Geometry::BravaisVectorsSC();
                else if( type == "fcc" )
                    bravais_vectors = Geometry::BravaisVectorsFCC();
                else if( type == "bcc" )
                    bravais_vectors = Geometry::BravaisVectorsBCC();
                else if( type == "hex2d" || type == "hex2d60" )
                    bravais_vectors = Geometry::BravaisVectorsHex2D60();
                else if( type == "hex2d120" )
                    bravais_vectors = Geometry::BravaisVectorsHex2D120();
                else
                    Log( Log_Level::Warning, Log_Sender::IO, "Bravais lattice \"" + type + "\" unknown. Using simple cubic..." );
            }
            else if( f.Find( "bravais_vectors" ) )
            {
                for( int v = 0; v < 3; ++v )
                {
                    f.GetLine();
                    f.iss >> bravais_vectors[v][0] >> bravais_vectors[v][1] >> bravais_vectors[v][2];
                }
            }
            else if( f.Find( "bravais_matrix" ) )
            {
                for( int row = 0; row < 3; ++row )
                {
                    f.GetLine();
                    f.iss >> bravais_vectors[0][row] >> bravais_vectors[1][row] >> bravais_vectors[2][row];
                }
            }

            f.Read_3( n_cells, "n_basis_cells" );

            // Basis cell, either in this file or in a separate one
            std::string basis_file;
            if( f.Find( "basis_file" ) )
                f.iss >> basis_file;
            else if( f.Find( "basis" ) )
                basis_file = config_file;
            if( !basis_file.empty() )
            {
                ConfigFile b( basis_file );
                if( b.Find( "basis" ) )
                {
                    std::size_t n_cell_atoms = 0;
                    b.GetLine();
                    b.iss >> n_cell_atoms;
                    cell_atoms.assign( n_cell_atoms, Vec3{} );
                    cell_mu_s.assign( n_cell_atoms, 1.0 );
                    for( std::size_t iatom = 0; iatom < n_cell_atoms; ++iatom )
                    {
                        b.GetLine();
                        b.iss >> cell_atoms[iatom][0] >> cell_atoms[iatom][1] >> cell_atoms[iatom][2];
                    }
                }
            }
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::IO,
                 std::string( "Failed to read Geometry parameters: " ) + e.what() + ". Leaving values at default." );
        }

        // Magnetic moments: `mu_s m0 m1 ...`; missing values repeat mu_s[0] (Configparser.cpp:343-361)
        try
        {
            ConfigFile f( config_file );
            if( f.Find( "mu_s" ) )
            {
                for( std::size_t iatom = 0; iatom < cell_atoms.size(); ++iatom )
                    if( !( f.iss >> cell_mu_s[iatom] ) )
                        cell_mu_s[iatom] = cell_mu_s[0];
            }
            else
                Log( Log_Level::Error, Log_Sender::IO, "Keyword 'mu_s' not found. Using Default: 1" );
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::IO, std::string( "Unable to read mu_s: " ) + e.what() );
        }
    }

    // The reference stores the Bravais vectors scaled by nothing and multiplies positions by the lattice constant
    auto geometry = std::make_shared<Geometry>( bravais_vectors, n_cells, cell_atoms, cell_mu_s, lattice_constant );

    // Pinning (Pinning_from_Config, Configparser.cpp:573-699) and defects (Configparser.cpp:389-441; the tables are read by
    // Pinned_from_File / Defects_from_File, Dataparser.cpp:681-775) -- compile-time options of the reference, always built here
    if( !config_file.empty() )
    {
        Pinning pinning;
        Defects defects;
        pinning.pinned_cell.assign( cell_atoms.size(), Vec3{ 0, 0, 1 } );
        ConfigFile f( config_file );
        if( f.Find( "atom_types" ) ) // (propagates: State_Setup fails, as for every Hamiltonian / lattice this library cannot build)
            throw std::runtime_error( "spirit_b200: disordered basis cells ('atom_types' with concentrations) are not supported" );
        try
        {
            int na = 0, nb = 0, nc = 0;
            f.Read_Single( pinning.na_left, "pin_na_left" );
            f.Read_Single( pinning.na_right, "pin_na_right" );
            f.Read_Single( na, "pin_na " );
            if( na > 0 && ( pinning.na_left == 0 || pinning.na_right == 0 ) )
                pinning.na_left = pinning.na_right = na;
            f.Read_Single( pinning.nb_left, "pin_nb_left" );
            f.Read_Single( pinning.nb_right, "pin_nb_right" );
            f.Read_Single( nb, "pin_nb " );
            if( nb > 0 && ( pinning.nb_left == 0 || pinning.nb_right == 0 ) )
                pinning.nb_left = pinning.nb_right = nb;
            f.Read_Single( pinning.nc_left, "pin_nc_left" );
            f.Read_Single( pinning.nc_right, "pin_nc_right" );
            f.Read_Single( nc, "pin_nc " );
            if( nc > 0 && ( pinning.nc_left == 0 || pinning.nc_right == 0 ) )
                pinning.nc_left = pinning.nc_right = nc;
            if( pinning.na_left > 0 || pinning.na_right > 0 || pinning.nb_left > 0 || pinning.nb_right > 0 || pinning.nc_left > 0
                || pinning.nc_right > 0 )
            {
                if( f.Find( "pinning_cell" ) )
                    for( std::size_t i = 0; i < cell_atoms.size(); ++i )
                    {
                        f.GetLine();
                        f.iss >> pinning.pinned_cell[i][0] >> pinning.pinned_cell[i][1] >> pinning.pinned_cell[i][2];
                    }
                else
                {
                    pinning.na_left = pinning.na_right = pinning.nb_left = pinning.nb_right = pinning.nc_left = pinning.nc_right = 0;
                    Log( Log_Level::Warning, Log_Sender::IO, "Pinning specified, but keyword 'pinning_cell' not found. Won't pin any spins!" );
                }
            }
            // tables: `n_pinned N` / `n_defects N` followed by N lines in this file, or a file of such lines
            auto read_table = [&]( const std::string & count_key, const std::string & file_key, auto && row )
            {
                std::string file;
                if( f.Find( count_key ) )
                    file = config_file;
                else if( f.Find( file_key ) )
                    f.iss >> file;
                if( file.empty() )
                    return;
                ConfigFile t( file );
                int n = int( 1e8 ), read = 0;
                if( t.Find( count_key ) )
                    t.iss >> n;
                else
                    t.To_Start();
                while( read < n && t.GetLine() )
                {
                    row( t.iss );
                    ++read;
                }
            };
            read_table( "n_pinned", "pinned_from_file", [&]( std::istringstream & in ) {
                LatticeSite site;
                Vec3 o{ 0, 0, 0 };
                in >> site.i >> site.translations[0] >> site.translations[1] >> site.translations[2] >> o[0] >> o[1] >> o[2];
                pinning.sites.push_back( site );
                pinning.spins.push_back( o );
            } );
            read_table( "n_defects", "defects_from_file", [&]( std::istringstream & in ) {
                LatticeSite site;
                int type = 0;
                in >> site.i >> site.translations[0] >> site.translations[1] >> site.translations[2] >> type;
                defects.sites.push_back( site );
                defects.types.push_back( type );
            } );
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::IO,
                 std::string( "Failed to read pinning / defects: " ) + e.what() + ". Leaving values at default." );
        }
        geometry->set_pinning_and_defects( pinning, defects );
    }
    return geometry;
}

// ---------------------------------------------------------------------------------------------
// LLG parameters (Configparser.cpp:701-824)
// ---------------------------------------------------------------------------------------------
std::shared_ptr<Parameters_LLG> Parameters_LLG_from_Config( const std::string & config_file )
{
    auto p = std::make_shared<Parameters_LLG>();
    std::random_device random;
    p->rng_seed = int( random() );
    p->prng     = std::mt19937( p->rng_seed );

    if( !config_file.empty() )
    {
        try
        {
            ConfigFile f( config_file );
            std::string str_max_walltime = "0";
            f.Read_Single( p->output_file_tag, "output_file_tag" );
            f.Read_Single( p->output_folder, "llg_output_folder" );
            f.Read_Single( p->output_any, "llg_output_any" );
            f.Read_Single( p->output_initial, "llg_output_initial" );
            f.Read_Single( p->output_final, "llg_output_final" );
            f.Read_Single( p->output_energy_spin_resolved, "llg_output_energy_spin_resolved" );
            f.Read_Single( p->output_energy_step, "llg_output_energy_step" );
            f.Read_Single( p->output_energy_archive, "llg_output_energy_archive" );
            f.Read_Single( p->output_energy_divide_by_nspins, "llg_output_energy_divide_by_nspins" );
            f.Read_Single( p->output_energy_add_readability_lines, "llg_output_energy_add_readability_lines" );
            f.Read_Single( p->output_configuration_step, "llg_output_configuration_step" );
            f.Read_Single( p->output_configuration_archive, "llg_output_configuration_archive" );
            f.Read_Single( p->output_vf_filetype, "llg_output_configuration_filetype" );
            f.Read_Single( str_max_walltime, "llg_max_walltime" );
            p->max_walltime_sec = duration_from_string( str_max_walltime );
            f.Read_Single( p->rng_seed, "llg_seed" );
            p->prng = std::mt19937( p->rng_seed );
            f.Read_Single( p->n_iterations, "llg_n_iterations" );
            f.Read_Single( p->n_iterations_log, "llg_n_iterations_log" );
            f.Read_Single( p->n_iterations_amortize, "llg_n_iterations_amortize" );
            f.Read_Single( p->dt, "llg_dt" );
            f.Read_Single( p->temperature, "llg_temperature" );
            f.Read_3( p->temperature_gradient_direction, "llg_temperature_gradient_direction" );
            p->temperature_gradient_direction.normalize();
            f.Read_Single( p->temperature_gradient_inclination, "llg_temperature_gradient_inclination" );
            f.Read_Single( p->damping, "llg_damping" );
            f.Read_Single( p->beta, "llg_beta" );
            f.Read_Single( p->stt_use_gradient, "llg_stt_use_gradient" );
            f.Read_Single( p->stt_magnitude, "llg_stt_magnitude" );
            f.Read_3( p->stt_polarisation_normal, "llg_stt_polarisation_normal" );
            p->stt_polarisation_normal.normalize();
            f.Read_Single( p->force_convergence, "llg_force_convergence" );
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::IO, std::string( "Unable to read LLG parameters: " ) + e.what() );
        }
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// GNEB parameters (Configparser.cpp:1010-1104)
// ---------------------------------------------------------------------------------------------
std::shared_ptr<Parameters_GNEB> Parameters_GNEB_from_Config( const std::string & config_file )
{
    auto p = std::make_shared<Parameters_GNEB>();
    if( !config_file.empty() )
    {
        try
        {
            ConfigFile f( config_file );
            std::string str_max_walltime = "0";
            f.Read_Single( p->output_file_tag, "output_file_tag" );
            f.Read_Single( p->output_folder, "gneb_output_folder" );
            f.Read_Single( p->output_any, "gneb_output_any" );
            f.Read_Single( p->output_initial, "gneb_output_initial" );
            f.Read_Single( p->output_final, "gneb_output_final" );
            f.Read_Single( p->output_energies_step, "gneb_output_energies_step" );
            f.Read_Single( p->output_energies_add_readability_lines, "gneb_output_energies_add_readability_lines" );
            f.Read_Single( p->output_energies_interpolated, "gneb_output_energies_interpolated" );
            f.Read_Single( p->output_energies_divide_by_nspins, "gneb_output_energies_divide_by_nspins" );
            f.Read_Single( p->output_chain_step, "gneb_output_chain_step" );
            f.Read_Single( p->output_vf_filetype, "gneb_output_chain_filetype" );
            f.Read_Single( str_max_walltime, "gneb_max_walltime" );
            p->max_walltime_sec = duration_from_string( str_max_walltime );
            f.Read_Single( p->spring_constant, "gneb_spring_constant" );
            f.Read_Single( p->force_convergence, "gneb_force_convergence" );
            f.Read_Single( p->n_iterations, "gneb_n_iterations" );
            f.Read_Single( p->n_iterations_log, "gneb_n_iterations_log" );
            f.Read_Single( p->n_iterations_amortize, "gneb_n_iterations_amortize" );
            f.Read_Single( p->n_E_interpolations, "gneb_n_energy_interpolations" );
            f.Read_Single( p->moving_endpoints, "gneb_moving_endpoints" );
            f.Read_Single( p->equilibrium_delta_Rx_left, "gneb_equilibrium_delta_Rx_left" );
            f.Read_Single( p->equilibrium_delta_Rx_right, "gneb_equilibrium_delta_Rx_right" );
            f.Read_Single( p->translating_endpoints, "gneb_translating_endpoints" );
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::IO, std::string( "Unable to read GNEB parameters: " ) + e.what() );
        }
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// Pairs table (Dataparser.cpp:290-514): header columns i j da db dc Jij Dij Dijx Dijy Dijz (or Dija/b/c)
// ---------------------------------------------------------------------------------------------
static void Pairs_from_File( const std::string & pairs_file, const Geometry & geometry, Hamiltonian & ham )
{
    ConfigFile f( pairs_file );
    int n_pairs = 0;
    if( f.Find( "n_interaction_pairs" ) )
        f.iss >> n_pairs;
    else
    {
        n_pairs = int( 1e8 );
        f.To_Start();
    }

    std::vector<std::string> columns( 20 );
    int col_i = -1, col_j = -1, col_da = -1, col_db = -1, col_dc = -1, col_J = -1, col_Dx = -1, col_Dy = -1, col_Dz = -1,
        col_Dij = -1;
    bool has_J = false, has_Dij = false, dmi_abc = false;
    f.GetLine();
    for( std::size_t i = 0; i < columns.size(); ++i )
    {
        f.iss >> columns[i];
        const std::string c = lower( columns[i] );
        if( c == "i" )
            col_i = int( i );
        else if( c == "j" )
            col_j = int( i );
        else if( c == "da" )
            col_da = int( i );
        else if( c == "db" )
            col_db = int( i );
        else if( c == "dc" )
            col_dc = int( i );
        else if( c == "jij" )
        {
            col_J = int( i );
            has_J = true;
        }
        else if( c == "dij" )
        {
            col_Dij = int( i );
            has_Dij = true;
        }
        else if( c == "dijx" || c == "dija" )
            col_Dx = int( i );
        else if( c == "dijy" || c == "dijb" )
            col_Dy = int( i );
        else if( c == "dijz" || c == "dijc" )
            col_Dz = int( i );
    }
    // Note: the reference maps Dija/b/c onto the x/y/z columns as well and never sets its DMI_abc flag
    // (Dataparser.cpp:352-363), so the components are always taken as Cartesian.
    (void)dmi_abc;
    (void)geometry;
    const bool dmi_xyz = col_Dx >= 0 && col_Dy >= 0 && col_Dz >= 0;
    if( !has_J && !dmi_xyz )
        Log( Log_Level::Warning, Log_Sender::IO, "No interactions could be found in pairs file \"" + pairs_file + "\"" );

    int i_pair = 0;
    while( f.GetLine() && i_pair < n_pairs )
    {
        int pi = 0, pj = 0, da = 0, db = 0, dc = 0;
        double Jij = 0, Dij = 0, D1 = 0, D2 = 0, D3 = 0;
        std::string sdump;
        for( int i = 0; i < int( columns.size() ); ++i )
        {
            if( i == col_i )
                f.iss >> pi;
            else if( i == col_j )
                f.iss >> pj;
            else if( i == col_da )
                f.iss >> da;
            else if( i == col_db )
                f.iss >> db;
            else if( i == col_dc )
                f.iss >> dc;
            else if( i == col_J && has_J )
                f.iss >> Jij;
            else if( i == col_Dij && has_Dij )
                f.iss >> Dij;
            else if( i == col_Dx && dmi_xyz )
                f.iss >> D1;
            else if( i == col_Dy && dmi_xyz )
                f.iss >> D2;
            else if( i == col_Dz && dmi_xyz )
                f.iss >> D3;
            else
                f.iss >> sdump;
        }
        const double dnorm = std::sqrt( D1 * D1 + D2 * D2 + D3 * D3 );
        if( dnorm != 0 )
        {
            D1 /= dnorm;
            D2 /= dnorm;
            D3 /= dnorm;
        }
        if( !has_Dij )
            Dij = dnorm;

        const std::array<int, 3> tnew{ da, db, dc };
        const std::array<int, 3> tneg{ -da, -db, -dc };
        if( Jij != 0 )
        {
            int at = -1;
            for( std::size_t k = 0; k < ham.exchange_pairs_in.size(); ++k )
            {
                const auto & p = ham.exchange_pairs_in[k];
                if( ( pi == p.i && pj == p.j && tnew == p.translations ) || ( pi == p.j && pj == p.i && tneg == p.translations ) )
                {
                    at = int( k );
                    break;
                }
            }
            if( at >= 0 )
                ham.exchange_magnitudes_in[at] += Jij;
            else
            {
                ham.exchange_pairs_in.push_back( Pair{ pi, pj, tnew } );
                ham.exchange_magnitudes_in.push_back( Jij );
            }
        }
        if( Dij != 0 )
        {
            int at = -1, dfact = 1;
            for( std::size_t k = 0; k < ham.dmi_pairs_in.size(); ++k )
            {
                const auto & p = ham.dmi_pairs_in[k];
                if( pi == p.i && pj == p.j && tnew == p.translations )
                {
                    at = int( k );
                    break;
                }
                if( pi == p.j && pj == p.i && tneg == p.translations )
                {
                    at    = int( k );
                    dfact = -1; // pseudo-vector: the reversed pair carries the mirrored D
                    break;
                }
            }
            if( at >= 0 )
            {
                const Vec3 newD = ham.dmi_magnitudes_in[at] * ham.dmi_normals_in[at] + ( dfact * Dij ) * Vec3{ D1, D2, D3 };
                const double n  = std::sqrt( newD.x * newD.x + newD.y * newD.y + newD.z * newD.z );
                ham.dmi_magnitudes_in[at] = n;
                ham.dmi_normals_in[at]    = newD / n;
            }
            else
            {
                ham.dmi_pairs_in.push_back( Pair{ pi, pj, tnew } );
                ham.dmi_magnitudes_in.push_back( Dij );
                ham.dmi_normals_in.push_back( Vec3{ D1, D2, D3 } );
            }
        }
        ++i_pair;
    }
}


// Per-atom anisotropy table (Dataparser.cpp:97-260, Anisotropy_from_File): a header line with the columns
// i, K, Kx Ky Kz (or Ka Kb Kc, in units of the Bravais vectors), K4 in any order (the first six columns count), then one
// line per basis atom. Without a K column the magnitude is the norm of the vector. Entries with K == 0 / K4 == 0 are
// dropped, like in the reference.
static void Anisotropy_from_File( const std::string & file, const Geometry & geometry, Hamiltonian & ham )
{
    ConfigFile f( file );
    int n_anisotropy = 0;
    if( f.Find( "n_anisotropy" ) )
        f.iss >> n_anisotropy;
    else
    {
        n_anisotropy = int( 1e8 );
        f.To_Start();
    }
    std::vector<std::string> columns( 6 );
    int col_i = -1, col_K = -1, col_Kx = -1, col_Ky = -1, col_Kz = -1, col_Ka = -1, col_Kb = -1, col_Kc = -1, col_K4 = -1;
    f.GetLine();
    for( std::size_t i = 0; i < columns.size(); ++i )
    {
        f.iss >> columns[i];
        const std::string c = lower( columns[i] );
        if( c == "i" )
            col_i = int( i );
        else if( c == "k" )
            col_K = int( i );
        else if( c == "kx" )
            col_Kx = int( i );
        else if( c == "ky" )
            col_Ky = int( i );
        else if( c == "kz" )
            col_Kz = int( i );
        else if( c == "ka" )
            col_Ka = int( i );
        else if( c == "kb" )
            col_Kb = int( i );
        else if( c == "kc" )
            col_Kc = int( i );
        else if( c == "k4" )
            col_K4 = int( i );
    }
    const bool K_magnitude = col_K >= 0;
    const bool K_xyz = col_Kx >= 0 && col_Ky >= 0 && col_Kz >= 0, K_abc = col_Ka >= 0 && col_Kb >= 0 && col_Kc >= 0;
    if( !K_xyz && !K_abc )
        Log( Log_Level::Warning, Log_Sender::IO, "No anisotropy data could be found in header of file \"" + file + "\"" );

    ham.anisotropy_indices.clear();
    ham.anisotropy_magnitudes.clear();
    ham.anisotropy_normals.clear();
    ham.cubic_anisotropy_indices.clear();
    ham.cubic_anisotropy_magnitudes.clear();
    int spin_i = 0, i_anisotropy = 0;
    double spin_K = 0, k1 = 0, k2 = 0, k3 = 0, spin_K4 = 0; // values persist over lines that do not set them, as in the reference
    while( f.GetLine() && i_anisotropy < n_anisotropy )
    {
        std::string sdump;
        for( int i = 0; i < int( columns.size() ); ++i )
        {
            if( i == col_i )
                f.iss >> spin_i;
            else if( i == col_K )
                f.iss >> spin_K;
            else if( ( i == col_Kx && K_xyz ) || ( i == col_Ka && K_abc ) )
                f.iss >> k1;
            else if( ( i == col_Ky && K_xyz ) || ( i == col_Kb && K_abc ) )
                f.iss >> k2;
            else if( ( i == col_Kz && K_xyz ) || ( i == col_Kc && K_abc ) )
                f.iss >> k3;
            else if( i == col_K4 )
                f.iss >> spin_K4;
            else
                f.iss >> sdump;
        }
        Vec3 K_temp{ k1, k2, k3 };
        if( K_abc )
        {
            const double lc = geometry.lattice_constant;
            k1              = K_temp.dot( lc * geometry.bravais_vectors[0] );
            k2              = K_temp.dot( lc * geometry.bravais_vectors[1] );
            k3              = K_temp.dot( lc * geometry.bravais_vectors[2] );
            K_temp          = Vec3{ k1, k2, k3 };
        }
        if( K_magnitude )
        {
            K_temp.normalize();
            if( K_temp.norm() == 0 )
                K_temp = Vec3{ 0, 0, 1 };
        }
        else
        {
            spin_K = K_temp.norm();
            if( spin_K != 0 )
                K_temp.normalize();
        }
        if( spin_K != 0 )
        {
            ham.anisotropy_indices.push_back( spin_i );
            ham.anisotropy_magnitudes.push_back( spin_K );
            ham.anisotropy_normals.push_back( K_temp );
        }
        if( spin_K4 != 0 )
        {
            ham.cubic_anisotropy_indices.push_back( spin_i );
            ham.cubic_anisotropy_magnitudes.push_back( spin_K4 );
        }
        ++i_anisotropy;
    }
}

// ---------------------------------------------------------------------------------------------
// Heisenberg Hamiltonian (Configparser.cpp:1195-1633)
// ---------------------------------------------------------------------------------------------
std::shared_ptr<Hamiltonian> Hamiltonian_from_Config( const std::string & config_file, std::shared_ptr<Geometry> geometry )
{
    auto ham                     = std::make_shared<Hamiltonian>( geometry );
    std::string hamiltonian_type = "heisenberg_neighbours";
    double B = 0, K = 0, K4 = 0;
    Vec3 B_normal{ 0, 0, 1 }, K_normal{ 0, 0, 1 };
    std::size_t n_shells_exchange = 0, n_shells_dmi = 0;
    int dm_chirality = 1;
    bool anisotropy_from_file = false;

    if( !config_file.empty() )
    {
        try
        {
            ConfigFile f( config_file );
            f.Read_Single( hamiltonian_type, "hamiltonian" );
            if( hamiltonian_type != "heisenberg_neighbours" && hamiltonian_type != "heisenberg_pairs" )
                throw Unsupported(
                    "Hamiltonian type \"" + hamiltonian_type
                    + "\" is outside the hot path of spirit_b200 (only heisenberg_neighbours / heisenberg_pairs)" );

            std::array<int, 3> bc{ 0, 0, 0 };
            f.Read_3( bc, "boundary_conditions" );
            for( int d = 0; d < 3; ++d )
                ham->boundary_conditions[d] = bc[d] != 0;

            f.Read_Single( B, "external_field_magnitude" );
            f.Read_3( B_normal, "external_field_normal" );
            B_normal.normalize();
            if( B_normal.norm() < 1e-8 )
                B_normal = { 0, 0, 1 };

            // per-atom anisotropy table in the config file itself or in a file of its own (Configparser.cpp:1352-1385)
            std::string anisotropy_file;
            if( f.Find( "n_anisotropy" ) )
                anisotropy_file = config_file;
            else if( f.Find( "anisotropy_file" ) )
                f.iss >> anisotropy_file;
            if( !anisotropy_file.empty() )
            {
                Anisotropy_from_File( anisotropy_file, *geometry, *ham );
                anisotropy_from_file = true;
            }
            else
            {
                f.Read_Single( K, "anisotropy_magnitude" );
                f.Read_3( K_normal, "anisotropy_normal" );
                K_normal.normalize();
                f.Read_Single( K4, "cubic_anisotropy_magnitude" );
            }

            if( hamiltonian_type == "heisenberg_pairs" )
            {
                std::string pairs_file;
                if( f.Find( "n_interaction_pairs" ) )
                    pairs_file = config_file;
                else if( f.Find( "interaction_pairs_file" ) )
                    f.iss >> pairs_file;
                if( !pairs_file.empty() )
                    Pairs_from_File( pairs_file, *geometry, *ham );
            }
            else
            {
                f.Read_Single( n_shells_exchange, "n_shells_exchange" );
                ham->exchange_shell_magnitudes.assign( n_shells_exchange, 0.0 );
                if( n_shells_exchange > 0 && f.Find( "jij" ) )
                    for( auto & j : ham->exchange_shell_magnitudes )
                        f.iss >> j;
                f.Read_Single( n_shells_dmi, "n_shells_dmi" );
                ham->dmi_shell_magnitudes.assign( n_shells_dmi, 0.0 );
                if( n_shells_dmi > 0 && f.Find( "dij" ) )
                    for( auto & d : ham->dmi_shell_magnitudes )
                        f.iss >> d;
                f.Read_Single( dm_chirality, "dm_chirality" );
            }

            std::string ddi_method = "none";
            f.Read_String( ddi_method, "ddi_method" );
            if( ddi_method == "none" )
                ham->ddi_method = DDI_Method::None;
            else if( ddi_method == "fft" )
                ham->ddi_method = DDI_Method::FFT;
            else if( ddi_method == "fmm" )
                ham->ddi_method = DDI_Method::FMM;
            else if( ddi_method == "cutoff" )
                ham->ddi_method = DDI_Method::Cutoff;
            else
                Log( Log_Level::Warning, Log_Sender::IO, "Keyword 'ddi_method' got passed invalid method \"" + ddi_method + "\". Setting to \"none\"." );
            f.Read_3( ham->ddi_n_periodic_images, "ddi_n_periodic_images" );
            f.Read_Single( ham->ddi_pb_zero_padding, "ddi_pb_zero_padding" );
            f.Read_Single( ham->ddi_cutoff_radius, "ddi_radius" );

            if( f.Find( "n_interaction_quadruplets" ) )
            {
                int nq = 0;
                f.iss >> nq;
                if( nq > 0 )
                    throw Unsupported( "quadruplet interactions are outside the hot path of spirit_b200" );
            }
        }
        catch( const Unsupported & )
        {
            // physics this library does not compute: the State must not come up with a silently different Hamiltonian
            // (the reference fails State_Setup for a Hamiltonian it cannot build); State_Setup returns nullptr
            throw;
        }
        catch( const std::exception & e )
        {
            Log( Log_Level::Error, Log_Sender::IO, std::string( "Unable to read Hamiltonian from config file: " ) + e.what() );
        }
    }

    // Hamiltonian_Heisenberg.cpp:35 -- the field is stored in meV per mu_B
    ham->external_field_magnitude = B * constants::mu_B;
    ham->external_field_normal    = B_normal;
    if( !anisotropy_from_file && K != 0 )
        for( int i = 0; i < geometry->n_cell_atoms; ++i )
        {
            ham->anisotropy_indices.push_back( i );
            ham->anisotropy_magnitudes.push_back( K );
            ham->anisotropy_normals.push_back( K_normal );
        }
    if( !anisotropy_from_file && K4 != 0 )
        for( int i = 0; i < geometry->n_cell_atoms; ++i )
        {
            ham->cubic_anisotropy_indices.push_back( i );
            ham->cubic_anisotropy_magnitudes.push_back( K4 );
        }
    ham->dmi_shell_chirality = dm_chirality;
    ham->Update_Interactions();
    return ham;
}

std::shared_ptr<Spin_System> Spin_System_from_Config( const std::string & config_file )
{
    auto geometry    = Geometry_from_Config( config_file );
    auto llg         = Parameters_LLG_from_Config( config_file );
    auto hamiltonian = Hamiltonian_from_Config( config_file, geometry );
    return std::make_shared<Spin_System>( hamiltonian, geometry, llg );
}

} // namespace config
} // namespace sb
