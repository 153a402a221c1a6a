/* TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored <bbcat-base/EnhancedFile.h>.
 * /root/reference/src/BiQuad.cpp includes it but only uses EnhancedFile inside `#if BBCDEBUG_LEVEL >= 3`
 * (BiQuad.cpp:5 sets the level to 1), so an empty header is enough to compile the file unmodified. */
#ifndef ORACLE_SHIM_BBCAT_BASE_ENHANCEDFILE_H
#define ORACLE_SHIM_BBCAT_BASE_ENHANCEDFILE_H
#include "misc.h"
#endif
