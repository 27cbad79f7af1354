// NumbTh.h -- include name used by dwu4/fhe-si clients; the classes live in fhesi_host.h
#pragma once
#include "fhesi_host.h"
