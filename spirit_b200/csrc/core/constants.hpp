// Physical constants. Values are the reference's, digit for digit
// (core/include/utility/Constants.hpp:18-46) -- they pin numerical parity.
// Conventions: energy meV, time ps, field T.
#pragma once

namespace sb
{
namespace constants
{
constexpr double mu_B  = 0.057883817555;  // Bohr magneton [meV/T]
constexpr double mu_0  = 2.0133545 * 1e-28; // vacuum permeability [T^2 m^3 / meV]
constexpr double k_B   = 0.08617330350;   // Boltzmann constant [meV/K]
constexpr double hbar  = 0.6582119514;    // [meV ps / rad]
constexpr double gamma = 0.1760859644;    // gyromagnetic ratio of the electron [rad/(ps T)]
constexpr double g_e   = 2.00231930436182;
constexpr double mRy   = 1.0 / 13.605693009;
constexpr double erg   = 6.2415091 * 1e14;
constexpr double Pi    = 3.141592653589793238462643383279502884197169399375105820974;
constexpr double Pi_2  = 1.570796326794896619231321691639751442098584699687552910487;
} // namespace constants
} // namespace sb
