/* Build information. Replaces core/include/Spirit/Version.h:9-27. */
#ifndef SPIRIT_B200_VERSION_H
#define SPIRIT_B200_VERSION_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

SPIRIT_API const int Spirit_Version_Major(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const int Spirit_Version_Minor(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const int Spirit_Version_Patch(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Version(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Version_Revision(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Version_Full(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Compiler(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Compiler_Version(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Compiler_Full(  ) SPIRIT_NOEXCEPT;
/* always "double" */
SPIRIT_API const char * Spirit_Scalar_Type(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Defects(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Pinning(  ) SPIRIT_NOEXCEPT;
/* "ON": there is no other backend */
SPIRIT_API const char * Spirit_Cuda(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_OpenMP(  ) SPIRIT_NOEXCEPT;
SPIRIT_API int Spirit_OpenMP_Get_Num_Threads(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_Threads(  ) SPIRIT_NOEXCEPT;
SPIRIT_API const char * Spirit_FFTW(  ) SPIRIT_NOEXCEPT;
#endif
