// dcrt_ops.cpp -- DoubleCRT /=, Exp and randomize of the host layer (DoubleCRT.cpp:406-435,466-480):
// division undoes multiplication, Exp(3) equals two multiplications.  Run by tests/test_host_cpp.py.
#include <iostream>
#include "DoubleCRT.h"
#include "FHEContext.h"
int main(){ FHEcontext context(22,80,to_ZZ(23L),7,3); activeContext=&context; context.SetUpSIContext();
 SetSeed(to_ZZ(5L)); DoubleCRT a(context); a.randomize(); DoubleCRT b=a; b*=to_ZZ(12345L); b/=to_ZZ(12345L);
 ZZX pa,pb; a.toPoly(pa); b.toPoly(pb); if(!(pa==pb)) {std::cout<<"div FAIL\n"; return 1;}
 DoubleCRT c=a; c.Exp(3); DoubleCRT d=a; d*=a; d*=a; ZZX pc,pd; c.toPoly(pc); d.toPoly(pd); if(!(pc==pd)){std::cout<<"exp FAIL\n";return 1;}
 std::cout<<"dcrt ok\n"; }
