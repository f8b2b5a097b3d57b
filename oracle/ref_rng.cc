/* oracle/ref_rng.cc -- TEST INFRASTRUCTURE ONLY.
 * Seedable replacement for libc rand() used by the shimmed reference build
 * (oracle/shim.h).  splitmix64 state step, top 15 bits out (RAND_MAX 0x7fff).
 * The same generator is restated in oracle/rt_oracle.c (orc_rand15) so that
 * the C restatement can be driven by the identical stream.
 */
#include <cstdint>

static thread_local uint64_t g_state = 0x9E3779B97F4A7C15ull;

extern "C" void oracle_seed(uint64_t s) { g_state = s; }

extern "C" int oracle_rand(void) {
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (int)(z >> 49);
}
