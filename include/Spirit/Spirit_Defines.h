/* Scalar type of the library. The reference generates this header from
 * core/CMake/Spirit_Defines.h.in; this build is double precision only. */
#ifndef SPIRIT_B200_DEFINES_H
#define SPIRIT_B200_DEFINES_H
#define SPIRIT_SCALAR_TYPE_DOUBLE
#define SPIRIT_SCALAR_TYPE double
typedef SPIRIT_SCALAR_TYPE scalar;
#endif
