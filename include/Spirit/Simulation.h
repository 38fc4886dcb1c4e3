/* Start / stop / query iterative methods. Replaces core/include/Spirit/Simulation.h:33-218.
 * In scope: LLG and GNEB with solvers VP, SIB, Depondt, Heun, RK4. MC / MMF / EMA and the
 * LBFGS / OSO solvers log an error and return (they are not part of the accelerated path). */
#ifndef SPIRIT_B200_SIMULATION_H
#define SPIRIT_B200_SIMULATION_H
#include "Export.h"
struct State;
typedef struct State State;

#define Solver_VP 0
#define Solver_SIB 1
#define Solver_Depondt 2
#define Solver_Heun 3
#define Solver_RungeKutta4 4
#define Solver_LBFGS_OSO 5
#define Solver_LBFGS_Atlas 6
#define Solver_VP_OSO 7

/* Simulation.h:58-71; the arrays are allocated by the library and released by free_run_info */
struct Simulation_Run_Info
{
    int total_iterations;
    int total_walltime;
    float total_ips;
    float max_torque;
    int n_history_iteration;
    int * history_iteration;
    int n_history_max_torque;
    float * history_max_torque;
    int n_history_energy;
    float * history_energy;
};
#ifndef __cplusplus
typedef struct Simulation_Run_Info Simulation_Run_Info;
#endif

SPIRIT_API void free_run_info( Simulation_Run_Info info ) SPIRIT_NOEXCEPT; /* Simulation.h:73 */

SPIRIT_API void Simulation_MC_Start( State * state, int n_iterations SPIRIT_DEFAULT( -1 ), int n_iterations_log SPIRIT_DEFAULT( -1 ), bool singleshot SPIRIT_DEFAULT( false ), Simulation_Run_Info * info SPIRIT_DEFAULT( nullptr ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :81, out of scope */
/* Simulation.h:86, core/src/Spirit/Simulation.cpp:134-217 */
SPIRIT_API void Simulation_LLG_Start( State * state, int solver_type, int n_iterations SPIRIT_DEFAULT( -1 ), int n_iterations_log SPIRIT_DEFAULT( -1 ), bool singleshot SPIRIT_DEFAULT( false ), Simulation_Run_Info * info SPIRIT_DEFAULT( nullptr ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Simulation.h:91, Simulation.cpp:219-315 */
SPIRIT_API void Simulation_GNEB_Start( State * state, int solver_type, int n_iterations SPIRIT_DEFAULT( -1 ), int n_iterations_log SPIRIT_DEFAULT( -1 ), bool singleshot SPIRIT_DEFAULT( false ), Simulation_Run_Info * info SPIRIT_DEFAULT( nullptr ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
SPIRIT_API void Simulation_MMF_Start( State * state, int solver_type, int n_iterations SPIRIT_DEFAULT( -1 ), int n_iterations_log SPIRIT_DEFAULT( -1 ), bool singleshot SPIRIT_DEFAULT( false ), Simulation_Run_Info * info SPIRIT_DEFAULT( nullptr ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :96, out of scope */
SPIRIT_API void Simulation_EMA_Start( State * state, int n_iterations SPIRIT_DEFAULT( -1 ), int n_iterations_log SPIRIT_DEFAULT( -1 ), bool singleshot SPIRIT_DEFAULT( false ), Simulation_Run_Info * info SPIRIT_DEFAULT( nullptr ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :101, out of scope */

/* One iteration followed by the post-iteration hook (Simulation.h:113, Simulation.cpp:449-540) */
SPIRIT_API void Simulation_SingleShot( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
SPIRIT_API void Simulation_N_Shot( State * state, int N, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :122 */
SPIRIT_API void Simulation_Stop( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;     /* :125 */
SPIRIT_API void Simulation_Stop_All( State * state ) SPIRIT_NOEXCEPT;                                                                           /* :128 */

SPIRIT_API float Simulation_Get_MaxTorqueComponent( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :142 */
SPIRIT_API void Simulation_Get_Chain_MaxTorqueComponents( State * state, float * torques, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;               /* :150 */
SPIRIT_API float Simulation_Get_MaxTorqueNorm( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;      /* :157 */
SPIRIT_API void Simulation_Get_Chain_MaxTorqueNorms( State * state, float * torques, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;                    /* :165 */
SPIRIT_API float Simulation_Get_IterationsPerSecond( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT; /* :173 */
SPIRIT_API int Simulation_Get_Iteration( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;             /* :176 */
SPIRIT_API float Simulation_Get_Time( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;                /* :183 */
SPIRIT_API int Simulation_Get_Wall_Time( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;             /* :186 */
SPIRIT_API const char * Simulation_Get_Solver_Name( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;  /* :195 */
SPIRIT_API const char * Simulation_Get_Method_Name( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;  /* :204 */
SPIRIT_API bool Simulation_Running_On_Image( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;         /* :212 */
SPIRIT_API bool Simulation_Running_On_Chain( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;                                              /* :215 */
SPIRIT_API bool Simulation_Running_Anywhere_On_Chain( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;                                     /* :218 */
#endif
