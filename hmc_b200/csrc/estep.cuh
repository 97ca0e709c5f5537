#pragma once
#include "common.cuh"
namespace hmc {
struct EstepTrie { void release() {} };
}
