#include "logging.hpp"

#include <cstdio>
#include <ctime>
#include <fstream>
#include <exception>
#include <mutex>
#include <stdexcept>

namespace sb
{

Logger Log;

namespace
{
std::mutex log_mutex;
const char * level_names[]  = { "  ALL  ", "SEVERE ", " ERROR ", "WARNING", " PARAM ", " INFO  ", " DEBUG " };
const char * sender_names[] = { "ALL ", "IO  ", "GNEB", "LLG ", "MC  ", "MMF ", "EMA ", "API ", "UI  ", "HTST" };
} // namespace

void Logger::operator()( Log_Level level, Log_Sender sender, const std::string & message, int idx_image, int idx_chain )
{
    std::lock_guard<std::mutex> guard( log_mutex );
    ++n_entries;
    if( level == Log_Level::Error || level == Log_Level::Severe )
        ++n_errors;
    if( level == Log_Level::Warning )
        ++n_warnings;
    if( messages_to_console && int( level ) <= int( level_console ) )
    {
        char idx[16] = "  ";
        if( idx_image >= 0 )
            std::snprintf( idx, sizeof( idx ), "%02d", idx_image + 1 );
        std::fprintf(
            stderr, "[%s] [%s] [%s]  %s\n", level_names[int( level )], sender_names[int( sender )], idx, message.c_str() );
    }
    (void)idx_chain;
    if( int( level ) <= int( level_file ) )
    {
        char idx[16] = "--";
        if( idx_image >= 0 )
            std::snprintf( idx, sizeof( idx ), "%02d", idx_image + 1 );
        file_lines.push_back( std::string( "[" ) + level_names[int( level )] + "] [" + sender_names[int( sender )] + "] [" + idx + "]  " + message + "\n" );
    }
}

std::string Logger::file_name() const
{
    static const std::string started = []
    {
        const std::time_t t = std::time( nullptr );
        std::tm parts{};
        localtime_r( &t, &parts );
        char buf[64];
        std::strftime( buf, sizeof( buf ), "%Y-%m-%d_%H-%M-%S", &parts );
        return std::string( buf );
    }();
    if( file_tag == "<time>" )
        return "Log_" + started + ".txt";
    return file_tag.empty() ? "Log.txt" : "Log_" + file_tag + ".txt";
}

void Logger::Append_to_File()
{
    std::lock_guard<std::mutex> guard( log_mutex );
    if( !messages_to_file )
        return;
    std::ofstream out( output_folder + "/" + file_name(), std::ios::app );
    for( ; n_lines_written < file_lines.size(); ++n_lines_written )
        out << file_lines[n_lines_written];
}

void Logger::Dump_to_File()
{
    std::lock_guard<std::mutex> guard( log_mutex );
    if( !messages_to_file )
        return;
    std::ofstream out( output_folder + "/" + file_name(), std::ios::trunc );
    for( const auto & line : file_lines )
        out << line;
    n_lines_written = file_lines.size();
}

void Logger::SendBlock( Log_Level level, Log_Sender sender, const std::vector<std::string> & messages, int idx_image, int idx_chain )
{
    for( const auto & m : messages )
        ( *this )( level, sender, m, idx_image, idx_chain );
}

void handle_exception_api( const char * function, int idx_image, int idx_chain )
{
    try
    {
        throw;
    }
    catch( const std::out_of_range & e )
    {
        Log( Log_Level::Warning, Log_Sender::API, std::string( function ) + ": " + e.what(), idx_image, idx_chain );
    }
    catch( const std::exception & e )
    {
        Log( Log_Level::Error, Log_Sender::API, std::string( function ) + ": " + e.what(), idx_image, idx_chain );
    }
    catch( ... )
    {
        Log( Log_Level::Error, Log_Sender::API, std::string( function ) + ": unknown exception", idx_image, idx_chain );
    }
}

} // namespace sb
