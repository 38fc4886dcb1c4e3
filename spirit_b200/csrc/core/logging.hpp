// Minimal logging with the reference's levels and senders (core/include/Spirit/Log.h:24-60,
// core/include/utility/Logging.hpp). Messages go to the console; errors and warnings are counted
// (Log_Get_N_Errors / Log_Get_N_Warnings). No log file is written.
#pragma once

#include <string>
#include <vector>

namespace sb
{

enum class Log_Level
{
    All       = 0,
    Severe    = 1,
    Error     = 2,
    Warning   = 3,
    Parameter = 4,
    Info      = 5,
    Debug     = 6
};

enum class Log_Sender
{
    All  = 0,
    IO   = 1,
    GNEB = 2,
    LLG  = 3,
    MC   = 4,
    MMF  = 5,
    EMA  = 6,
    API  = 7,
    UI   = 8,
    HTST = 9
};

struct Logger
{
    bool messages_to_console = true;
    Log_Level level_console  = Log_Level::Parameter;
    bool messages_to_file    = false;
    Log_Level level_file     = Log_Level::Parameter;
    std::string output_folder = ".";
    std::string file_tag      = "";
    int n_entries = 0, n_errors = 0, n_warnings = 0;
    // formatted entries (level <= level_file) waiting for / already written to the log file (Logging.cpp:160-225)
    std::vector<std::string> file_lines;
    std::size_t n_lines_written = 0;
    std::string file_name() const; // Log_<tag>.txt, Log.txt without a tag, Log_<start time>.txt for the tag "<time>"
    void Append_to_File();         // lines not yet written; no-op unless messages_to_file
    void Dump_to_File();           // the whole log

    void operator()( Log_Level level, Log_Sender sender, const std::string & message, int idx_image = -1, int idx_chain = -1 );
    void SendBlock( Log_Level level, Log_Sender sender, const std::vector<std::string> & messages, int idx_image = -1, int idx_chain = -1 );
};

extern Logger Log;

// Called from the catch(...) of every API function: logs the in-flight exception and returns
// (the C API never lets an exception cross the ABI, core/include/utility/Exception.hpp:119-121).
void handle_exception_api( const char * function, int idx_image = -1, int idx_chain = -1 );

} // namespace sb
