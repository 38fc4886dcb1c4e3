/* GNEB parameters. Replaces core/include/Spirit/Parameters_GNEB.h:20-200. */
#ifndef SPIRIT_B200_PARAMETERS_GNEB_H
#define SPIRIT_B200_PARAMETERS_GNEB_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

#define GNEB_IMAGE_NORMAL 0
#define GNEB_IMAGE_CLIMBING 1
#define GNEB_IMAGE_FALLING 2
#define GNEB_IMAGE_STATIONARY 3

/* Parameters_GNEB.h:48 */
SPIRIT_API void Parameters_GNEB_Set_Output_Tag( State * state, const char * tag, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :51 */
SPIRIT_API void Parameters_GNEB_Set_Output_Folder( State * state, const char * folder, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :54 */
SPIRIT_API void Parameters_GNEB_Set_Output_General( State * state, bool any, bool initial, bool final, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :65 */
SPIRIT_API void Parameters_GNEB_Set_Output_Energies( State * state, bool step, bool interpolated, bool divide_by_nos, bool add_readability_lines, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :75 */
SPIRIT_API void Parameters_GNEB_Set_Output_Chain( State * state, bool step, int filetype SPIRIT_DEFAULT( 3 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :86 */
SPIRIT_API void Parameters_GNEB_Set_N_Iterations( State * state, int n_iterations, int n_iterations_log, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :96 */
SPIRIT_API void Parameters_GNEB_Set_Convergence( State * state, float convergence, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :99 */
SPIRIT_API void Parameters_GNEB_Set_Spring_Constant( State * state, float spring_constant, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :103 */
SPIRIT_API void Parameters_GNEB_Set_Spring_Force_Ratio( State * state, float ratio, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :106 */
SPIRIT_API void Parameters_GNEB_Set_Path_Shortening_Constant( State * state, float path_shortening_constant, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :109 */
SPIRIT_API void Parameters_GNEB_Set_Moving_Endpoints( State * state, bool moving_endpoints, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :112 */
SPIRIT_API void Parameters_GNEB_Set_Translating_Endpoints( State * state, bool translating_endpoints, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :116 */
SPIRIT_API void Parameters_GNEB_Set_Equilibrium_Delta_Rx( State * state, float delta_Rx_left, float delta_Rx_right, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :126 */
SPIRIT_API void Parameters_GNEB_Set_Climbing_Falling( State * state, int image_type, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :130: maxima climb, minima fall (core/src/Spirit/Parameters_GNEB.cpp:338-367) */
SPIRIT_API void Parameters_GNEB_Set_Image_Type_Automatically( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :133 */
SPIRIT_API void Parameters_GNEB_Set_N_Energy_Interpolations( State * state, int n, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :144 */
SPIRIT_API const char * Parameters_GNEB_Get_Output_Tag( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :147 */
SPIRIT_API const char * Parameters_GNEB_Get_Output_Folder( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :150 */
SPIRIT_API void Parameters_GNEB_Get_Output_General( State * state, bool * any, bool * initial, bool * final, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :154 */
SPIRIT_API void Parameters_GNEB_Get_Output_Energies( State * state, bool * step, bool * interpolated, bool * divide_by_nos, bool * add_readability_lines, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :159 */
SPIRIT_API void Parameters_GNEB_Get_Output_Chain( State * state, bool * step, int * filetype, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :169 */
SPIRIT_API void Parameters_GNEB_Get_N_Iterations( State * state, int * iterations, int * iterations_log, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :179 */
SPIRIT_API float Parameters_GNEB_Get_Convergence( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :182 */
SPIRIT_API float Parameters_GNEB_Get_Spring_Constant( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :185 */
SPIRIT_API float Parameters_GNEB_Get_Spring_Force_Ratio( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :188 */
SPIRIT_API float Parameters_GNEB_Get_Path_Shortening_Constant( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :191 */
SPIRIT_API bool Parameters_GNEB_Get_Moving_Endpoints( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :194 */
SPIRIT_API bool Parameters_GNEB_Get_Translating_Endpoints( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :197 */
SPIRIT_API void Parameters_GNEB_Get_Equilibrium_Delta_Rx( State * state, float * delta_Rx_left, float * delta_Rx_right, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :202 */
SPIRIT_API int Parameters_GNEB_Get_Climbing_Falling( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :205 */
SPIRIT_API int Parameters_GNEB_Get_N_Energy_Interpolations( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
