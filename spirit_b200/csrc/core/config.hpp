// input.cfg parsing for the hot-path keys (geometry, Heisenberg Hamiltonian, LLG and GNEB parameters).
// Semantics follow the reference's keyword scanner and *_from_Config functions
// (core/src/io/Filter_File_Handle.cpp:115-160, core/src/io/Configparser.cpp:154-1633,
// core/src/io/Dataparser.cpp:259-514): free-form `keyword value...` lines, first line whose start
// matches the keyword (case-insensitive) wins, '#' starts a comment, '|' and '+' are ignored, missing
// keys leave the hard-coded defaults.
#pragma once

#include "state.hpp"

#include <sstream>

namespace sb
{
namespace config
{

// Input that asks for physics outside this library (another Hamiltonian type, quadruplets): State_Setup fails with it
// instead of coming up with a Hamiltonian that silently lacks terms.
struct Unsupported : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

// Keyword scanner over one file, loaded once
class ConfigFile
{
public:
    explicit ConfigFile( const std::string & filename ); // throws if the file cannot be opened
    // Position on the first line starting with `keyword`; the stream then holds the rest of that line
    bool Find( const std::string & keyword );
    // Advance to the next (non-comment) line
    bool GetLine();
    void To_Start()
    {
        cursor_ = 0;
    }
    std::istringstream iss;

    template<typename T>
    bool Read_Single( T & var, const std::string & keyword )
    {
        if( !Find( keyword ) )
            return false;
        iss >> var;
        return true;
    }
    template<typename V>
    bool Read_3( V & v, const std::string & keyword )
    {
        if( !Find( keyword ) )
            return false;
        iss >> v[0] >> v[1] >> v[2];
        return true;
    }
    bool Read_String( std::string & var, const std::string & keyword );

private:
    std::vector<std::string> lines_;
    std::size_t cursor_ = 0;
};

std::shared_ptr<Geometry> Geometry_from_Config( const std::string & config_file );
std::shared_ptr<Parameters_LLG> Parameters_LLG_from_Config( const std::string & config_file );
std::shared_ptr<Parameters_GNEB> Parameters_GNEB_from_Config( const std::string & config_file );
std::shared_ptr<Hamiltonian> Hamiltonian_from_Config( const std::string & config_file, std::shared_ptr<Geometry> geometry );
std::shared_ptr<Spin_System> Spin_System_from_Config( const std::string & config_file );
void Log_from_Config( const std::string & config_file, bool quiet );

// "hh:mm:ss" -> seconds (core/src/utility/Timing.cpp DurationFromString)
long duration_from_string( const std::string & s );

} // namespace config
} // namespace sb
