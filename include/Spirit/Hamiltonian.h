/* Heisenberg Hamiltonian setters/getters. Replaces core/include/Spirit/Hamiltonian.h:31-172.
 * Setters take float (as the reference does) and rebuild the pair tables the GPU stencil consumes. */
#ifndef SPIRIT_B200_HAMILTONIAN_H
#define SPIRIT_B200_HAMILTONIAN_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

#define SPIRIT_CHIRALITY_BLOCH 1
#define SPIRIT_CHIRALITY_NEEL 2
#define SPIRIT_CHIRALITY_BLOCH_INVERSE -1
#define SPIRIT_CHIRALITY_NEEL_INVERSE -2
#define SPIRIT_DDI_METHOD_NONE 0
#define SPIRIT_DDI_METHOD_FFT 1
#define SPIRIT_DDI_METHOD_FMM 2
#define SPIRIT_DDI_METHOD_CUTOFF 3

/* Hamiltonian.h:65, core/src/Spirit/Hamiltonian.cpp:30-62 */
SPIRIT_API void Hamiltonian_Set_Boundary_Conditions( State * state, const bool * periodical, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:69 [T] */
SPIRIT_API void Hamiltonian_Set_Field( State * state, float magnitude, const float * normal, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:73 [meV], same K for every basis atom */
SPIRIT_API void Hamiltonian_Set_Anisotropy( State * state, float magnitude, const float * normal, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:77 */
SPIRIT_API void Hamiltonian_Set_Cubic_Anisotropy( State * state, float magnitude, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:81 neighbour shells */
SPIRIT_API void Hamiltonian_Set_Exchange( State * state, int n_shells, const float * jij, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:86 */
SPIRIT_API void Hamiltonian_Set_DMI( State * state, int n_shells, const float * dij, int chirality SPIRIT_DEFAULT( SPIRIT_CHIRALITY_BLOCH ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:98; methods NONE and FFT are implemented */
SPIRIT_API void Hamiltonian_Set_DDI( State * state, int ddi_method, int n_periodic_images[3], float cutoff_radius SPIRIT_DEFAULT( 0 ), bool pb_zero_padding SPIRIT_DEFAULT( true ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:110: always "Heisenberg" */
SPIRIT_API const char * Hamiltonian_Get_Name( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:113 */
SPIRIT_API void Hamiltonian_Get_Boundary_Conditions( State * state, bool * periodical, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:117 */
SPIRIT_API void Hamiltonian_Get_Field( State * state, float * magnitude, float * normal, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:121 */
SPIRIT_API void Hamiltonian_Get_Anisotropy( State * state, float * magnitude, float * normal, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:125 */
SPIRIT_API void Hamiltonian_Get_Cubic_Anisotropy( State * state, float * magnitude, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:133 */
SPIRIT_API void Hamiltonian_Get_Exchange_Shells( State * state, int * n_shells, float * jij, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:137 (redundant list: both directions) */
SPIRIT_API int Hamiltonian_Get_Exchange_N_Pairs( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:140 */
SPIRIT_API void Hamiltonian_Get_Exchange_Pairs( State * state, int idx[][2], int translations[][3], float * Jij, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:149 */
SPIRIT_API void Hamiltonian_Get_DMI_Shells( State * state, int * n_shells, float * dij, int * chirality, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:153 */
SPIRIT_API int Hamiltonian_Get_DMI_N_Pairs( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Hamiltonian.h:163 */
SPIRIT_API void Hamiltonian_Get_DDI( State * state, int * ddi_method, int n_periodic_images[3], float * cutoff_radius, bool * pb_zero_padding, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
