/* Initial paths for GNEB. Replaces core/include/Spirit/Transitions.h:29-46. */
#ifndef SPIRIT_B200_TRANSITIONS_H
#define SPIRIT_B200_TRANSITIONS_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

/* Transitions.h:29: geodesic interpolation between two images */
SPIRIT_API void Transition_Homogeneous( State * state, int idx_1, int idx_2, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Transitions.h:37 */
SPIRIT_API void Transition_Homogeneous_Insert_Interpolated( State * state, int n_interpolate, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Transitions.h:44 */
SPIRIT_API void Transition_Add_Noise_Temperature( State * state, float temperature, int idx_1, int idx_2, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
