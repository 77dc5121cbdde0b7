// oracle/ref_tls_rng.h -- TEST INFRASTRUCTURE: forced include for UWBNetwork.cpp ONLY.
// The reference keeps one file-scope `std::mt19937 rng;` (UWBNetwork.cpp:4) shared by every
// UWBNetwork instance.  The CPU-baseline harness steps independent vehicles on several threads;
// making that one object thread_local removes the data race without touching the source.
// Single-threaded behaviour is unchanged.
#pragma once
#include <random>
#define mt19937 mt19937 thread_local
