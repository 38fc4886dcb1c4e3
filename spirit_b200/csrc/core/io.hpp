// What the file-writing callers share: the OVF segment header of a spin system (core/src/io/OVF_File.cpp:14-43) and the
// time stamp of output file names (core/src/utility/Timing.cpp:19-27).
#pragma once

#include "ovf.hpp"
#include "state.hpp"

#include <string>

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

} // namespace io
} // namespace sb
