// libc_rand.cpp -- TEST INFRASTRUCTURE, linked into oracle/_ref only.
// The reference's sampleHWt draws from lrand48(), which NumbTh.h:32-35 maps to libc rand().  To make
// a reference run reproducible by the oracle, rand() is interposed here and draws 31 bits from the
// same SplitMix64 stream that stands in for NTL's generator (neither stream is
// pinned by the reference).  srand() is a no-op: the stream is seeded through NTL::SetSeed.
#include "ntl_compat.h"

extern "C" int rand(void) noexcept { return (int)(NTL::GlobalRandomStream().next64() >> 33); }
extern "C" void srand(unsigned) noexcept {}
