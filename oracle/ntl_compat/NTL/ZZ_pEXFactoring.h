// NTL/ZZ_pEXFactoring.h -- forwarding header of the NTL stand-in used to compile the reference's own sources
// (oracle/_ref).  Test infrastructure.
#pragma once
#include "../ntl_compat.h"
