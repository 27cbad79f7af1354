// NTL/tools.h -- client-facing include name; everything lives in ntl_shim.h
#pragma once
#include "../ntl_shim.h"
