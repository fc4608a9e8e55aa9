/* oracle/shim/windows.h — build-environment stub so the UNMODIFIED reference sources that
 * `#include <windows.h>` (monte_cpp/CBCT_real2.cpp:2) compile on Linux.  Nothing from the
 * reference is copied here.  It also pins the reference's time-based RNG seed
 * (CBCT_real2.cpp:102: init_genrand((unsigned long)time(NULL))) to ORACLE_REF_SEED so the
 * as-shipped binary becomes a deterministic known-answer generator.                         */
#ifndef ORACLE_SHIM_WINDOWS_H
#define ORACLE_SHIM_WINDOWS_H
#include <time.h>
typedef unsigned long DWORD;
static inline DWORD timeGetTime(void) { return (DWORD)(clock() * 1000 / CLOCKS_PER_SEC); }
#ifndef ORACLE_REF_SEED
#define ORACLE_REF_SEED 5489
#endif
#define time(x) ((time_t)ORACLE_REF_SEED)
#endif
