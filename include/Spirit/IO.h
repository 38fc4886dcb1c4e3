/* Spin configurations and chains on disk: OVF 2.0 files as the reference reads and writes them.
 * Replaces the vector-field part of core/include/Spirit/IO.h:26-103 (implementation: core/src/Spirit/IO.cpp:165-860,
 * core/src/io/OVF_File.cpp:14-43, core/thirdparty/ovf) and the energy files (IO.h:118-130). The interpolated chain energies and the eigenmode files are in Compat.h
 * (not provided). */
#ifndef SPIRIT_B200_IO_H
#define SPIRIT_B200_IO_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

/* IO.h:26-38 */
#define IO_Fileformat_OVF_bin 0
#define IO_Fileformat_OVF_bin4 1
#define IO_Fileformat_OVF_bin8 2
#define IO_Fileformat_OVF_text 3
#define IO_Fileformat_OVF_csv 4

/* IO.h:46 */
SPIRIT_API int IO_System_From_Config( State * state, const char * file, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:50 */
SPIRIT_API void IO_Positions_Write( State * state, const char * file, int format SPIRIT_DEFAULT( IO_Fileformat_OVF_bin ), const char * comment SPIRIT_DEFAULT( "-" ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:60 */
SPIRIT_API int IO_N_Images_In_File( State * state, const char * file, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:63 */
SPIRIT_API void IO_Image_Read( State * state, const char * file, int idx_image_infile SPIRIT_DEFAULT( 0 ), int idx_image_inchain SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:67 */
SPIRIT_API void IO_Image_Write( State * state, const char * file, int format SPIRIT_DEFAULT( IO_Fileformat_OVF_bin ), const char * comment SPIRIT_DEFAULT( "-" ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:72 */
SPIRIT_API void IO_Image_Append( State * state, const char * file, int format SPIRIT_DEFAULT( IO_Fileformat_OVF_bin ), const char * comment SPIRIT_DEFAULT( "-" ), int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:87 */
SPIRIT_API void IO_Chain_Read( State * state, const char * file, int start_image_infile SPIRIT_DEFAULT( 0 ), int end_image_infile SPIRIT_DEFAULT( -1 ), int insert_idx SPIRIT_DEFAULT( 0 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:92 */
SPIRIT_API void IO_Chain_Write( State * state, const char * file, int format SPIRIT_DEFAULT( IO_Fileformat_OVF_text ), const char * comment SPIRIT_DEFAULT( "-" ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:97 */
SPIRIT_API void IO_Chain_Append( State * state, const char * file, int format SPIRIT_DEFAULT( IO_Fileformat_OVF_text ), const char * comment SPIRIT_DEFAULT( "-" ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:105 */
SPIRIT_API void IO_Image_Write_Neighbours_Exchange( State * state, const char * file, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:109 */
SPIRIT_API void IO_Image_Write_Neighbours_DMI( State * state, const char * file, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:118 */
SPIRIT_API void IO_Image_Write_Energy_per_Spin( State * state, const char * file, int format, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:121 */
SPIRIT_API void IO_Image_Write_Energy( State * state, const char * file, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* IO.h:124 */
SPIRIT_API void IO_Chain_Write_Energies( State * state, const char * file, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
