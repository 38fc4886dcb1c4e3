// What the file-writing callers share: the OVF segment header of a spin system (core/src/io/OVF_File.cpp:14-43) and the
// time stamp of output file names (core/src/utility/Timing.cpp:19-27).
#pragma once

#include "ovf.hpp"
#include "state.hpp"

#include <string>
#include <vector>

namespace sb
{
namespace io
{

inline const char * version_full()
{
    return "2.2.0 (spirit_b200)";
}
// rectangular mesh, basis atoms folded into the x axis, lengths in nm; no value columns yet
ovf::Segment segment_of( const Spin_System & system );
// + title, comment and the three spin columns
ovf::Segment spin_segment( const Spin_System & system, const std::string & comment );
// "%Y-%m-%d_%H-%M-%S" of now (local time)
std::string current_date_time();

// Plain column files (no OVF header; core/src/io/Dataparser.cpp:23-51): everything after a '#' is a remark, a data line holds
// three numbers separated by blanks and / or commas. Returns the rows [n][3] of the whole file.
std::vector<double> read_column_text( const std::string & file );

// Energy tables (core/src/io/Datawriter.cpp:116-234): fmt's "{:^20}" / "{:^20.10f}" columns
std::string centred( const std::string & text, std::size_t width = 20 );
std::string fixed10( double v );
std::string shortest( double v ); // shortest decimal string that reads back exactly (fmt's "{}")
// header line: the given first columns, then one column per energy contribution of `s`
void write_energy_header( const Spin_System & s, const std::string & file, const std::vector<std::string> & columns, bool readability );
void append_image_energy( const Spin_System & s, long iteration, const std::string & file, bool normalize, bool readability );
// one line: E_tot, contributions, under a header with separator lines (Write_Image_Energy, Datawriter.cpp:186-206)
void write_image_energy( const Spin_System & s, const std::string & file, bool normalize, bool readability );
// one line per image: image, Rx, E_tot, contributions (the header always carries the separator lines, as in the reference)
void write_chain_energies( const Chain & chain, const std::string & file, bool normalize, bool readability );

} // namespace io
} // namespace sb
