/* Per-image data access. Replaces core/include/Spirit/System.h:22-76.
 * The returned scalar* are borrowed AoS [nos][3] views into live (pinned) host storage: callers may read and
 * write them between calls; the library re-uploads the spins to the GPU at the start of every
 * Simulation_*_Start / SingleShot / System_Update_Data and downloads after every iteration block. */
#ifndef SPIRIT_B200_SYSTEM_H
#define SPIRIT_B200_SYSTEM_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

SPIRIT_API int System_Get_Index( State * state ) SPIRIT_NOEXCEPT;                                                      /* System.h:25 */
SPIRIT_API int System_Get_NOS( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :28 */
SPIRIT_API scalar * System_Get_Spin_Directions( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :36 */
SPIRIT_API scalar * System_Get_Effective_Field( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :43 */
SPIRIT_API float System_Get_Rx( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :52 */
SPIRIT_API float System_Get_Energy( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :55 */
/* names separated by '|'; returns the required buffer length when names == NULL (System.h:58) */
SPIRIT_API int System_Get_Energy_Array_Names( State * state, char * names, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* returns the number of contributions when energies == NULL (System.h:61) */
SPIRIT_API int System_Get_Energy_Array( State * state, float * energies, bool divide_by_nspins SPIRIT_DEFAULT( true ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
SPIRIT_API void System_Print_Energy_Array( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :70 */
/* energies + effective field of the image, evaluated on the GPU (System.h:73) */
SPIRIT_API void System_Update_Data( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
