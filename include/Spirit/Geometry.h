/* Lattice geometry. Replaces core/include/Spirit/Geometry.h:30-186 (triangulation / tetrahedra are
 * visualisation helpers and not provided). */
#ifndef SPIRIT_B200_GEOMETRY_H
#define SPIRIT_B200_GEOMETRY_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

typedef enum
{
    Bravais_Lattice_Irregular   = 0,
    Bravais_Lattice_Rectilinear = 1,
    Bravais_Lattice_SC          = 2,
    Bravais_Lattice_Hex2D       = 3,
    Bravais_Lattice_Hex2D_60    = 4,
    Bravais_Lattice_Hex2D_120   = 5,
    Bravais_Lattice_HCP         = 6,
    Bravais_Lattice_BCC         = 7,
    Bravais_Lattice_FCC         = 8
} Bravais_Lattice_Type;

/* Geometry.h:51. Setters apply to every image of the chain and reset spins like the reference (new sites get +z) */
SPIRIT_API void Geometry_Set_Bravais_Lattice_Type( State * state, Bravais_Lattice_Type lattice_type ) SPIRIT_NOEXCEPT;
/* Geometry.h:54 */
SPIRIT_API void Geometry_Set_N_Cells( State * state, int n_cells[3] ) SPIRIT_NOEXCEPT;
/* Geometry.h:57 */
SPIRIT_API void Geometry_Set_Cell_Atoms( State * state, int n_atoms, float ** atoms ) SPIRIT_NOEXCEPT;
/* Geometry.h:60 */
SPIRIT_API void Geometry_Set_mu_s( State * state, float mu_s, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:63 */
SPIRIT_API void Geometry_Set_Cell_Atom_Types( State * state, int n_atoms, int * atom_types ) SPIRIT_NOEXCEPT;
/* Geometry.h:66 */
SPIRIT_API void Geometry_Set_Bravais_Vectors( State * state, float ta[3], float tb[3], float tc[3] ) SPIRIT_NOEXCEPT;
/* Geometry.h:69 */
SPIRIT_API void Geometry_Set_Lattice_Constant( State * state, float lattice_constant ) SPIRIT_NOEXCEPT;
/* Geometry.h:78 */
SPIRIT_API int Geometry_Get_NOS( State * state ) SPIRIT_NOEXCEPT;
/* Geometry.h:87: borrowed view [nos][3] */
SPIRIT_API scalar * Geometry_Get_Positions( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:97 */
SPIRIT_API int * Geometry_Get_Atom_Types( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:100 */
SPIRIT_API void Geometry_Get_Bounds( State * state, float min[3], float max[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:104 */
SPIRIT_API void Geometry_Get_Center( State * state, float center[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:118 */
SPIRIT_API Bravais_Lattice_Type Geometry_Get_Bravais_Lattice_Type( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:122 */
SPIRIT_API void Geometry_Get_Bravais_Vectors( State * state, float a[3], float b[3], float c[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:126 */
SPIRIT_API int Geometry_Get_Dimensionality( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:129: one value per basis atom */
SPIRIT_API void Geometry_Get_mu_s( State * state, float * mu_s, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:132 */
SPIRIT_API void Geometry_Get_N_Cells( State * state, int n_cells[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:138 */
SPIRIT_API void Geometry_Get_Cell_Bounds( State * state, float min[3], float max[3], int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:142 */
SPIRIT_API int Geometry_Get_N_Cell_Atoms( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Geometry.h:145 */
SPIRIT_API int Geometry_Get_Cell_Atoms( State * state, scalar ** atoms, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
