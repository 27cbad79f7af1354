// zz_probe.cpp -- the big-integer type that stands under BOTH the product's host layer (fhe-si_b200/host/
// ntl_shim.h) and the reference build (oracle/ntl_compat includes the same file), exercised against an
// independent implementation: Python integers (tests/test_host_cpp.py::test_zz_matches_python_integers).
// Reads "op a b" lines (decimal operands), prints one decimal result per line, NTL semantics:
//   add sub mul   div mod (floor division, result of % non-negative for a positive modulus)
//   shl shr (b = shift; >> shifts the magnitude and keeps the sign)
//   nbits nbytes   bytes (BytesFromZZ -> ZZFromBytes round trip through b bytes)   powmod (a^b mod 2^61-1)
//   invmod (a^-1 mod b)   cmp (-1, 0, 1)   modl (a % (long)b)
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "NTL/ZZ.h"
using namespace NTL;
int main() {
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream is(line);
    std::string op;
    ZZ a, b;
    is >> op >> a >> b;
    if (op == "add") std::cout << a + b;
    else if (op == "sub") std::cout << a - b;
    else if (op == "mul") std::cout << a * b;
    else if (op == "div") std::cout << a / b;
    else if (op == "mod") std::cout << a % b;
    else if (op == "shl") std::cout << (a << to_long(b));
    else if (op == "shr") std::cout << (a >> to_long(b));
    else if (op == "nbits") std::cout << NumBits(a);
    else if (op == "nbytes") std::cout << NumBytes(a);
    else if (op == "bytes") {
      std::vector<unsigned char> buf(to_long(b) + 1, 0);
      BytesFromZZ(buf.data(), a, to_long(b));
      ZZ r;
      ZZFromBytes(r, buf.data(), to_long(b));
      std::cout << r;
    } else if (op == "powmod") std::cout << PowerMod(a % ((ZZ(1L) << 61) - 1L), b, (ZZ(1L) << 61) - 1L);
    else if (op == "invmod") std::cout << InvMod(a % b, b);
    else if (op == "cmp") std::cout << (a < b ? -1 : (a == b ? 0 : 1));
    else if (op == "modl") std::cout << (a % to_long(b));
    else return 2;
    std::cout << "\n";
  }
  return 0;
}
