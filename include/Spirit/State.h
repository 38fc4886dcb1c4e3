/* State: the opaque handle every API function takes.
 * Replaces core/include/Spirit/State.h:61-93 of the reference (same symbols, same signatures). */
#ifndef SPIRIT_B200_STATE_H
#define SPIRIT_B200_STATE_H
#include "Export.h"

struct State;
typedef struct State State;
typedef struct State State;

/* Build a State from an input.cfg (empty string: defaults). Heap-allocated, freed only by State_Delete.
 * reference: State.h:71, core/src/Spirit/State.cpp:18-197 */
SPIRIT_API State * State_Setup( const char * config_file SPIRIT_DEFAULT( "" ), bool quiet SPIRIT_DEFAULT( false ) ) SPIRIT_NOEXCEPT;
/* reference: State.h:76 */
SPIRIT_API void State_Delete( State * state ) SPIRIT_NOEXCEPT;
/* reference: State.h:81 (re-synchronises counters after chain edits) */
SPIRIT_API void State_Update( State * state ) SPIRIT_NOEXCEPT;
/* reference: State.h:86. Config writing is outside the hot path: logs a warning, writes nothing. */
SPIRIT_API void State_To_Config( State * state, const char * config_file, const char * comment SPIRIT_DEFAULT( "" ) ) SPIRIT_NOEXCEPT;
/* reference: State.h:91 */
SPIRIT_API const char * State_DateTime( State * state ) SPIRIT_NOEXCEPT;
#endif
