/* Linkage of the C API: plain C symbols, `noexcept` for C++ callers
 * (same convention as the reference's core/include/Spirit/DLL_Define_Export.h:18-26). */
#ifndef SPIRIT_B200_EXPORT_H
#define SPIRIT_B200_EXPORT_H
#ifdef __cplusplus
#define SPIRIT_API extern "C" __attribute__( ( visibility( "default" ) ) )
#define SPIRIT_NOEXCEPT noexcept
#define SPIRIT_DEFAULT( x ) = x
#else
#include <stdbool.h>
#define SPIRIT_API
#define SPIRIT_NOEXCEPT
#define SPIRIT_DEFAULT( x )
#endif
#endif
