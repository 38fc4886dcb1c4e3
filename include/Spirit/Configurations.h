/* Spin-configuration initialisers (host side; they build the inputs of the accelerated path).
 * Replaces core/include/Spirit/Configurations.h:42-150. Positions are relative to the geometry centre; negative
 * cut-offs disable the respective filter. */
#ifndef SPIRIT_B200_CONFIGURATIONS_H
#define SPIRIT_B200_CONFIGURATIONS_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

#ifdef __cplusplus
static const float defaultPos[3]    = { 0, 0, 0 };
static const float defaultRect[3]   = { -1, -1, -1 };
static const float defaultNormal[3] = { 0, 0, 1 };
#endif

/* Configurations.h:49 */
SPIRIT_API void Configuration_To_Clipboard( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:52 */
SPIRIT_API void Configuration_From_Clipboard( State * state, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:58 */
SPIRIT_API bool Configuration_From_Clipboard_Shift( State * state, const float shift[3], const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:72 */
SPIRIT_API void Configuration_Domain( State * state, const float direction[3], const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:78 */
SPIRIT_API void Configuration_PlusZ( State * state, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:83 */
SPIRIT_API void Configuration_MinusZ( State * state, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:88: libstdc++ mt19937 of the LLG parameters, so the same llg_seed gives the reference's state */
SPIRIT_API void Configuration_Random( State * state, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), bool external SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:94 */
SPIRIT_API void Configuration_SpinSpiral( State * state, const char * direction_type, float q[3], float axis[3], float theta, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:100 -- not implemented (logs an error) */
SPIRIT_API void Configuration_SpinSpiral_2q( State * state, const char * direction_type, float q1[3], float q2[3], float axis[3], float theta, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:113 */
SPIRIT_API void Configuration_Add_Noise_Temperature( State * state, float temperature, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:119 -- eigenmodes are out of scope (logs an error) */
SPIRIT_API void Configuration_Displace_Eigenmode( State * state, int idx_mode, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:122 */
SPIRIT_API void Configuration_Skyrmion( State * state, float r, float order, float phase, bool upDown, bool achiral, bool rl, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:128 */
SPIRIT_API void Configuration_DW_Skyrmion( State * state, float dw_radius, float dw_width, float order, float phase, bool upDown, bool achiral, bool rl, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:134 */
SPIRIT_API void Configuration_Hopfion( State * state, float r, int order SPIRIT_DEFAULT( 1 ), const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), const float normal[3] SPIRIT_DEFAULT( defaultNormal ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:143 -- pinning is a compile-time feature that is off (as in the reference default build): no-op */
SPIRIT_API void Configuration_Set_Pinned( State * state, bool pinned, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Configurations.h:148 -- defects are off: no-op */
SPIRIT_API void Configuration_Set_Atom_Type( State * state, int type, const float position[3] SPIRIT_DEFAULT( defaultPos ), const float r_cut_rectangular[3] SPIRIT_DEFAULT( defaultRect ), float r_cut_cylindrical SPIRIT_DEFAULT( -1 ), float r_cut_spherical SPIRIT_DEFAULT( -1 ), bool inverted SPIRIT_DEFAULT( false ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
