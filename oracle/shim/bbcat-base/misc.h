/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Stand-in for the un-vendored <bbcat-base/misc.h> (bbcat-base >= 0.1.2.0,
 * required by /root/reference/configure.ac:40-44) so that the reference's own
 * in-tree sources compile unmodified into oracle/_ref/ (see oracle/Makefile).
 *
 * Three semantics are not visible in the mounted tree and are fixed here:
 *   - Sample_t is float   (forced by the overload sets in
 *                          SoundDelayBuffer.h:24-31 / SoundFormatConversions.h:60-68)
 *   - limited::limit      is a plain clamp (only reading consistent with
 *                          SoundFormatRawConversions.cpp:701 + debian/changelog:5)
 *   - MACHINE_IS_BIG_ENDIAN is false on x86-64
 */
#ifndef ORACLE_SHIM_BBCAT_BASE_MISC_H
#define ORACLE_SHIM_BBCAT_BASE_MISC_H

#include <stdint.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include <algorithm>
#include <cmath>

#define BBC_AUDIOTOOLBOX_START namespace bbcat {
#define BBC_AUDIOTOOLBOX_END   }

typedef unsigned int uint_t;
typedef signed int   sint_t;
typedef int16_t      sint16_t;
typedef int32_t      sint32_t;
typedef int64_t      sint64_t;
typedef float        Sample_t;

#define UNUSED_PARAMETER(x) ((void)(x))
#define MACHINE_IS_BIG_ENDIAN false
#define MEMALIGNED(n, x) x __attribute__((aligned(n)))

#define BBCERROR(...)  do { fprintf(stderr, "bbcat-ref error: "); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define BBCDEBUG(...)  do { } while (0)
#define BBCDEBUG1(x)   do { } while (0)
#define BBCDEBUG2(x)   do { } while (0)
#define BBCDEBUG3(x)   do { } while (0)
#define BBCDEBUG4(x)   do { } while (0)

namespace limited {
template <typename T> inline T limit(T v, T lo, T hi) { return std::min(std::max(v, lo), hi); }
template <typename T> inline T subz(T a, T b) { return (a >= b) ? (a - b) : T(); }
}  // namespace limited

#endif
