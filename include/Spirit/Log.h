/* Logging controls. Replaces the subset of core/include/Spirit/Log.h:25-120 that drivers use. */
#ifndef SPIRIT_B200_LOG_H
#define SPIRIT_B200_LOG_H
#include "Export.h"
#include "Spirit_Defines.h"
struct State;
typedef struct State State;

typedef enum { Log_Level_All = 0, Log_Level_Severe = 1, Log_Level_Error = 2, Log_Level_Warning = 3, Log_Level_Parameter = 4, Log_Level_Info = 5, Log_Level_Debug = 6 } Spirit_Log_Level;
typedef enum { Log_Sender_All = 0, Log_Sender_IO = 1, Log_Sender_GNEB = 2, Log_Sender_LLG = 3, Log_Sender_MC = 4, Log_Sender_MMF = 5, Log_Sender_EMA = 6, Log_Sender_API = 7, Log_Sender_UI = 8, Log_Sender_HTST = 9 } Spirit_Log_Sender;

/* Log.h:62 */
SPIRIT_API void Log_Send( State * state, Spirit_Log_Level level, Spirit_Log_Sender sender, const char * message, int idx_image SPIRIT_DEFAULT( -1 ), int idx_chain SPIRIT_DEFAULT( -1 ) ) SPIRIT_NOEXCEPT;
/* Log.h:92 */
SPIRIT_API void Log_Set_Output_To_Console( State * state, bool output, int level ) SPIRIT_NOEXCEPT;
/* Log.h:104 */
SPIRIT_API int Log_Get_N_Entries( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:107 */
SPIRIT_API int Log_Get_N_Errors( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:110 */
SPIRIT_API int Log_Get_N_Warnings( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:70 */
SPIRIT_API void Log_Append( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:73 */
SPIRIT_API void Log_Dump( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:90 */
SPIRIT_API void Log_Set_Output_File_Tag( State * state, const char * tag ) SPIRIT_NOEXCEPT;
/* Log.h:93 */
SPIRIT_API void Log_Set_Output_Folder( State * state, const char * folder ) SPIRIT_NOEXCEPT;
/* Log.h:99 */
SPIRIT_API void Log_Set_Output_To_File( State * state, bool output, int level ) SPIRIT_NOEXCEPT;
/* Log.h:107 */
SPIRIT_API const char * Log_Get_Output_File_Tag( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:110 */
SPIRIT_API const char * Log_Get_Output_Folder( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:113 */
SPIRIT_API bool Log_Get_Output_To_Console( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:116 */
SPIRIT_API int Log_Get_Output_Console_Level( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:119 */
SPIRIT_API bool Log_Get_Output_To_File( State * state ) SPIRIT_NOEXCEPT;
/* Log.h:122 */
SPIRIT_API int Log_Get_Output_File_Level( State * state ) SPIRIT_NOEXCEPT;
#endif
