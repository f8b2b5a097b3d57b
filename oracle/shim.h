/* oracle/shim.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Force-included (-include shim.h) in front of every UNMODIFIED reference
 * translation unit when oracle/_ref/libref.so is built from the sources under
 * /root/reference/programs.  It exists because programs/random.h:7 evaluates
 * `RAND_MAX + 1` in int, which overflows on glibc (RAND_MAX == INT_MAX) and
 * makes random_double() negative, so vec3::random_in_unit_sphere()
 * (programs/vec3.h:83-95) never terminates.  The reference was evidently
 * written where RAND_MAX == 0x7fff; this header restores that environment and
 * routes rand() to a seedable, thread-local generator so that renders are
 * reproducible and can run under OpenMP.  No reference source is edited.
 */
#ifndef ORACLE_SHIM_H
#define ORACLE_SHIM_H

/* Pull in every std header the reference uses BEFORE `rand` becomes a macro,
 * so the macro cannot rewrite declarations inside <cstdlib>. */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <vector>

#undef RAND_MAX
#define RAND_MAX 0x7fff

extern "C" int oracle_rand(void);          /* 15-bit uniform, oracle/ref_rng.cc */
extern "C" void oracle_seed(uint64_t s);   /* seeds the calling thread's stream */
#define rand oracle_rand

#endif /* ORACLE_SHIM_H */
