/* Chain of images (GNEB). Replaces core/include/Spirit/Chain.h:22-156. */
#ifndef SPIRIT_B200_CHAIN_H
#define SPIRIT_B200_CHAIN_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

/* Chain.h:22 */
SPIRIT_API int Chain_Get_NOI( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:34 */
SPIRIT_API bool Chain_next_Image( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:37 */
SPIRIT_API bool Chain_prev_Image( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:40 */
SPIRIT_API bool Chain_Jump_To_Image( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:43 */
SPIRIT_API void Chain_Set_Length( State * state, int n_images, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:56 */
SPIRIT_API void Chain_Image_to_Clipboard( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:59 */
SPIRIT_API void Chain_Replace_Image( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:62 */
SPIRIT_API void Chain_Insert_Image_Before( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:65 */
SPIRIT_API void Chain_Insert_Image_After( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:68 */
SPIRIT_API void Chain_Push_Back( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:71 */
SPIRIT_API bool Chain_Delete_Image( State * state, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:74 */
SPIRIT_API bool Chain_Pop_Back( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:86 */
SPIRIT_API void Chain_Get_Rx( State * state, float * Rx, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:93 */
SPIRIT_API void Chain_Get_Rx_Interpolated( State * state, float * Rx_interpolated, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:96 */
SPIRIT_API void Chain_Get_Energy( State * state, float * energy, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:103 */
SPIRIT_API void Chain_Get_Energy_Interpolated( State * state, float * E_interpolated, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:148 */
SPIRIT_API void Chain_Update_Data( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Chain.h:153 */
SPIRIT_API void Chain_Setup_Data( State * state, int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
#endif
