// import_probe.cpp -- Import(ifstream&, ZZX&) on a caller-supplied file; prints the degree.  A hostile or
// truncated file must end in Error() (non-zero exit), never in a huge allocation or an out-of-bounds write.
// Run by tests/test_host_cpp.py.
#include <fstream>
#include <iostream>
#include "Serialization.h"
int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  ZZX poly;
  Import(in, poly);
  std::cout << "degree " << deg(poly) << "\n";
  return 0;
}
