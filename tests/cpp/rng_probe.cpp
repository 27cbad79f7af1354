// rng_probe.cpp -- the host layer's random stream (ntl_shim.h): prints 4 words after an optional SetSeed.
//   rng_probe            unseeded draw (ChaCha20 keyed from the OS unless FHESI_TEST_RNG=splitmix64)
//   rng_probe <decimal>  SetSeed(<decimal>) first; the decimal may be wider than 64 bits
// Run by tests/test_host_cpp.py.
#include <iostream>
#include <sstream>
#include "NTL/ZZ.h"
using namespace NTL;
int main(int argc, char **argv) {
  if (argc > 1) {
    ZZ seed;
    std::istringstream(argv[1]) >> seed;
    SetSeed(seed);
  }
  for (int i = 0; i < 4; ++i) std::cout << GlobalRandomStream().next64() << (i < 3 ? " " : "\n");
  return 0;
}
