/* Derived quantities. Replaces core/include/Spirit/Quantities.h:18-21 (topological charge and the MMF
 * helpers are out of scope). */
#ifndef SPIRIT_B200_QUANTITIES_H
#define SPIRIT_B200_QUANTITIES_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

/* Quantities.h:18 */
SPIRIT_API void Quantity_Get_Average_Spin( State * state, float s[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Quantities.h:21: mean of mu_s * s, reduced on the GPU */
SPIRIT_API void Quantity_Get_Magnetization( State * state, float m[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Quantities.h:24 -- planar lattices with one basis atom */
SPIRIT_API float Quantity_Get_Topological_Charge( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Quantities.h:27 */
SPIRIT_API int Quantity_Get_Topological_Charge_Density( State * state, float * charge_density, int * triangle_indices, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
