/* LLG parameters. Replaces core/include/Spirit/Parameters_LLG.h:25-198. */
#ifndef SPIRIT_B200_PARAMETERS_LLG_H
#define SPIRIT_B200_PARAMETERS_LLG_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

/* Parameters_LLG.h:25 */
SPIRIT_API void Parameters_LLG_Set_Output_Tag( State * state, const char * tag, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :29 */
SPIRIT_API void Parameters_LLG_Set_Output_Folder( State * state, const char * folder, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :33 */
SPIRIT_API void Parameters_LLG_Set_Output_General( State * state, bool any, bool initial, bool final, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :44 */
SPIRIT_API void Parameters_LLG_Set_Output_Energy( State * state, bool energy_step, bool energy_archive, bool energy_spin_resolved, bool energy_divide_by_nos, bool energy_add_readability_lines, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :55 */
SPIRIT_API void Parameters_LLG_Set_Output_Configuration( State * state, bool configuration_step, bool configuration_archive, int configuration_filetype SPIRIT_DEFAULT( 3 ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :70 */
SPIRIT_API void Parameters_LLG_Set_N_Iterations( State * state, int n_iterations, int n_iterations_log, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :82 */
SPIRIT_API void Parameters_LLG_Set_Direct_Minimization( State * state, bool direct, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :90 */
SPIRIT_API void Parameters_LLG_Set_Convergence( State * state, float convergence, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :94 [ps] */
SPIRIT_API void Parameters_LLG_Set_Time_Step( State * state, float dt, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :98 */
SPIRIT_API void Parameters_LLG_Set_Damping( State * state, float damping, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :101 */
SPIRIT_API void Parameters_LLG_Set_Non_Adiabatic_Damping( State * state, float beta, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :111 (monolayer approximation only) */
SPIRIT_API void Parameters_LLG_Set_STT( State * state, bool use_gradient, float magnitude, const float normal[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :117 [K] */
SPIRIT_API void Parameters_LLG_Set_Temperature( State * state, float T, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :121 */
SPIRIT_API void Parameters_LLG_Set_Temperature_Gradient( State * state, float inclination, const float direction[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :132 */
SPIRIT_API const char * Parameters_LLG_Get_Output_Tag( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :135 */
SPIRIT_API const char * Parameters_LLG_Get_Output_Folder( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :138 */
SPIRIT_API void Parameters_LLG_Get_Output_General( State * state, bool * any, bool * initial, bool * final, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :142 */
SPIRIT_API void Parameters_LLG_Get_Output_Energy( State * state, bool * energy_step, bool * energy_archive, bool * energy_spin_resolved, bool * energy_divide_by_nos, bool * energy_add_readability_lines, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :148 */
SPIRIT_API void Parameters_LLG_Get_Output_Configuration( State * state, bool * configuration_step, bool * configuration_archive, int * configuration_filetype, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :159 */
SPIRIT_API void Parameters_LLG_Get_N_Iterations( State * state, int * iterations, int * iterations_log, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :170 */
SPIRIT_API bool Parameters_LLG_Get_Direct_Minimization( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :178 */
SPIRIT_API float Parameters_LLG_Get_Convergence( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :181 */
SPIRIT_API float Parameters_LLG_Get_Time_Step( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :184 */
SPIRIT_API float Parameters_LLG_Get_Damping( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :187 */
SPIRIT_API float Parameters_LLG_Get_Non_Adiabatic_Damping( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :190 */
SPIRIT_API float Parameters_LLG_Get_Temperature( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :193 */
SPIRIT_API void Parameters_LLG_Get_Temperature_Gradient( State * state, float * inclination, float direction[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* :197 */
SPIRIT_API void Parameters_LLG_Get_STT( State * state, bool * use_gradient, float * magnitude, float normal[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
