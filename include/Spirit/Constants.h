/* Physical constants, digit for digit the reference's (core/include/Spirit/Constants.h:12-40,
 * core/include/utility/Constants.hpp:18-46). */
#ifndef SPIRIT_B200_CONSTANTS_H
#define SPIRIT_B200_CONSTANTS_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

/* [meV/T] */
SPIRIT_API scalar Constants_mu_B(  ) SPIRIT_NOEXCEPT;
/* [T^2 m^3 / meV] */
SPIRIT_API scalar Constants_mu_0(  ) SPIRIT_NOEXCEPT;
/* [meV/K] */
SPIRIT_API scalar Constants_k_B(  ) SPIRIT_NOEXCEPT;
/* [meV ps / rad] */
SPIRIT_API scalar Constants_hbar(  ) SPIRIT_NOEXCEPT;
/* [mRy/meV] */
SPIRIT_API scalar Constants_mRy(  ) SPIRIT_NOEXCEPT;
/* [rad/(ps T)] */
SPIRIT_API scalar Constants_gamma(  ) SPIRIT_NOEXCEPT;
SPIRIT_API scalar Constants_g_e(  ) SPIRIT_NOEXCEPT;
SPIRIT_API scalar Constants_Pi(  ) SPIRIT_NOEXCEPT;
#endif
