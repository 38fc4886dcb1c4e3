// OVF 2.0 vector-field files (the "OOMMF vector field" format with the reference's conventions): what the reference
// reads and writes through its bundled ovf library (core/thirdparty/ovf/include/detail/write.hpp:75-330 for the layout
// of a written file, parse_rules.hpp for what a reader accepts; core/src/io/OVF_File.cpp:14-43 for the header values of a
// spin system). A file is a top header with a six-digit segment count that is patched in place on every append, followed by
// segments; a segment is a "# key: value" header and a data block (Text, CSV, Binary 4, Binary 8; binary blocks are
// little-endian and start with the check value 1234567.0f / 123456789012345.0).
// Host-side only: spins live in host memory between API calls (the device holds them during a simulation).
#pragma once

#include <string>
#include <vector>

namespace sb
{
namespace ovf
{

// IO_Fileformat_* of Spirit/IO.h
enum Format
{
    BIN  = 0, // binary in the library's precision (double)
    BIN4 = 1,
    BIN8 = 2,
    TEXT = 3,
    CSV  = 4
};

struct Segment
{
    std::string title, comment, valueunits, valuelabels, meshunit = "nm", meshtype = "rectangular";
    int valuedim = 0;
    int n_cells[3] = { 0, 0, 0 };
    int pointcount = 0;
    double bounds_min[3] = { 0, 0, 0 }, bounds_max[3] = { 0, 0, 0 }, origin[3] = { 0, 0, 0 }, step_size[3] = { 0, 0, 0 };
    int N = 0; // number of value rows: product of n_cells (rectangular) or pointcount (irregular)
};

// Parsed view of a file: the positions of its segments
class File
{
public:
    explicit File( const std::string & name );
    // A file that is only going to be written or appended to: only its top header is read (version line and segment count), so
    // that appending to an archive does not cost a pass over everything written so far. The read_* calls are not available.
    struct ForWriting
    {
    };
    File( const std::string & name, ForWriting );
    bool found = false, is_ovf = false;
    int n_segments = 0;
    std::string name, message;

    // header of segment `index`; throws std::runtime_error with the reason
    Segment read_segment_header( int index ) const;
    // the first min(n_rows, segment rows) rows of `valuedim` columns into data[row * valuedim + col]
    void read_segment_data( int index, const Segment & segment, int n_rows, double * data ) const;

    // write: a new file with this one segment; append: add a segment (creates the file when it does not exist)
    void write_segment( const Segment & segment, const double * data, int format );
    void append_segment( const Segment & segment, const double * data, int format );

private:
    struct Span
    {
        std::size_t begin = 0, end = 0; // byte range [begin, end) of "# Begin: Segment" .. "# End: Segment" line
    };
    std::vector<Span> segments_;
    std::size_t count_pos_ = 0; // position of the six digits of the segment count
    std::string contents_;
    bool head_only_ = false; // ForWriting: contents_ holds the first bytes of the file only, segments_ stays empty
    bool empty_     = false;
    void scan();
    void scan_head();
    void adopt_head( const std::string & head, int segments );
};

} // namespace ovf
} // namespace sb
