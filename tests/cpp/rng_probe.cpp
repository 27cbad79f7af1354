// rng_probe.cpp -- the host layer's random stream (ntl_shim.h): prints 4 (or argv[2]) words after an optional SetSeed.
//   rng_probe            unseeded draw (ChaCha20 keyed from the OS unless FHESI_TEST_RNG=splitmix64)
//   rng_probe <decimal>  SetSeed(<decimal>) first; the decimal may be wider than 64 bits
// Run by tests/test_host_cpp.py.
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include "NTL/ZZ.h"
using namespace NTL;
int main(int argc, char **argv) {
  if (argc > 1) {
    ZZ seed;
    if (std::string(argv[1]) != "-") {
      std::istringstream(argv[1]) >> seed;
      SetSeed(seed);
    }
  }
  const int count = argc > 2 ? atoi(argv[2]) : 4;
  for (int i = 0; i < count; ++i) std::cout << GlobalRandomStream().next64() << (i < count - 1 ? " " : "\n");
  return 0;
}
