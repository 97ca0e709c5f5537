// TEST INFRASTRUCTURE ONLY. Force-included before every reference TU built into oracle/_ref/.
// The reference was written for MSVC6/STLport and leans on transitive includes and on
// mixed-type min/max that libstdc++ templates reject; these overloads restore them without
// touching any arithmetic (each returns exactly what the MSVC template call would).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdarg>
#include <cfloat>
#include <algorithm>
#include <functional>
#include <string>
#include <map>
#include <list>
#include <vector>
#include <memory>
#include <stdexcept>
namespace std {
inline double min(double a, int b) { return a < (double)b ? a : (double)b; }   // HaploBuilder.cpp:320-321
inline int max(int a, size_t b) { return a < (int)b ? (int)b : a; }            // PatternManager.cpp:57
inline int min(int a, size_t b) { return (int)b < a ? (int)b : a; }            // PatternManager.cpp:328
}
